"""Schedules, solvers, weight fillings, snapshot codecs and decision rules - the parts of the
reference that are plain host logic (/root/reference/lr_adjust.py:183-302, gd.py:111-170,395-419,
nn_units.py:430-520, decision.py:276-550, core snapshotter) checked against closed forms."""
import os
import pickle

import numpy
import pytest

from veles.znicz_b200.core.config import root
from veles.znicz_b200.core.memory import Array
from veles.znicz_b200.core.workflow import DummyWorkflow
from veles.znicz_b200.ops import all2all, gd
from veles.znicz_b200.workflow import lr_adjust
from veles.znicz_b200.workflow.lr_adjust import LRAdjustPolicyRegistry


def test_lr_policies_closed_forms():
    reg = LRAdjustPolicyRegistry.registry
    assert {"exp", "fixed", "step_exp", "inv", "arbitrary_step"} <= set(reg)
    assert reg["fixed"](0.3)(12345) == 0.3
    exp = reg["exp"](0.1, gamma=0.5, a_ratio=0.01)
    assert abs(exp(200) - 0.1 * 0.5 ** 2) < 1e-12
    step = reg["step_exp"](0.1, gamma=0.1, step=100)
    assert [round(step(i), 6) for i in (0, 99, 100, 250)] == [0.1, 0.1, 0.01, 0.001]
    inv = reg["inv"](0.01, gamma=0.0001, pow_ratio=0.75)
    assert abs(inv(10000) - 0.01 * 2.0 ** -0.75) < 1e-12
    arb = reg["arbitrary_step"](0.001, lrs_with_lengths=[(1, 60000), (0.1, 5000), (0.01, 100)])
    assert [arb(i) for i in (0, 59999, 60000, 64999, 65000, 65100)] == \
        [0.001, 0.001, 0.0001, 0.0001, 0.00001, 0.0]
    with pytest.raises(ValueError):
        reg["arbitrary_step"](0.001, lrs_with_lengths=[])
    with pytest.raises(ValueError):
        reg["arbitrary_step"](0.001, lrs_with_lengths=[(1, 0)])


def test_learning_rate_adjust_unit_rewrites_every_gd_unit_per_minibatch():
    wf = DummyWorkflow()
    adj = lr_adjust.LearningRateAdjust(
        wf, lr_policy_name="step_exp", lr_parameters={"gamma": 0.5, "step": 2},
        bias_lr_policy_name="inv", bias_lr_parameters={"gamma": 1.0, "pow_ratio": 1.0})
    units = []
    for lr, lrb in ((0.1, 0.2), (0.4, 0.2)):
        g = gd.GradientDescent(wf, learning_rate=lr, learning_rate_bias=lrb)
        units.append(g)
        adj.add_gd_unit(g)
    adj.initialize()
    seen = []
    for _ in range(5):
        adj.run()
        seen.append((units[0].learning_rate, units[1].learning_rate, units[0].learning_rate_bias))
    assert [round(s[0], 6) for s in seen] == [0.1, 0.1, 0.05, 0.05, 0.025]
    assert [round(s[1], 6) for s in seen] == [0.4, 0.4, 0.2, 0.2, 0.1]
    assert [round(s[2], 6) for s in seen] == [0.2, 0.1, round(0.2 / 3, 6), 0.05, 0.04]
    with pytest.raises(TypeError):
        adj.add_gd_unit(object())


def _gd_pair(solvers=(), **gkw):
    wf = DummyWorkflow()
    f = all2all.All2All(wf, output_sample_shape=3, weights_stddev=0.1)
    rs = numpy.random.RandomState(3)
    f.input = Array(rs.uniform(-1, 1, (4, 5)).astype(numpy.float32))
    f.initialize(device=None)
    f.run()
    kw = dict(learning_rate=0.1, learning_rate_bias=0.1, weights_decay=0.0, gradient_moment=0.5,
              gradient_moment_bias=0.5, solvers=set(solvers))
    kw.update(gkw)
    g = gd.GradientDescent(wf, **kw)
    g.err_output = Array(rs.uniform(-1, 1, (4, 3)).astype(numpy.float32))
    g.input, g.output, g.weights, g.bias = f.input, f.output, f.weights, f.bias
    g.forward_unit = f
    g.initialize(device=None)
    return f, g


def test_momentum_step_matches_closed_form():
    f, g = _gd_pair(weights_decay=0.01, l1_vs_l2=0.25)
    w0 = f.weights.mem.copy()
    x, e = f.input.mem, g.err_output.mem
    grad = e.T.dot(x)
    g.run()
    reg = 0.01 * ((1 - 0.25) * w0 + 0.5 * 0.25 * numpy.sign(w0))
    v1 = -0.1 * (grad + reg)
    assert numpy.allclose(f.weights.mem, w0 + v1, atol=1e-6)
    w1 = f.weights.mem.copy()
    g.run()                                          # same err / input: momentum carries v1 over
    reg = 0.01 * ((1 - 0.25) * w1 + 0.5 * 0.25 * numpy.sign(w1))
    v2 = -0.1 * (grad + reg) + 0.5 * v1
    assert numpy.allclose(f.weights.mem, w1 + v2, atol=1e-6)


def test_adagrad_adadelta_fast_solvers():
    f, g = _gd_pair(("momentum", "adagrad"))      # (extra solvers need the moment vectors)
    w0 = f.weights.mem.copy()
    grad = g.err_output.mem.T.dot(f.input.mem)
    g.run()
    step = -0.1 * grad
    expect = step / numpy.sqrt(step ** 2 + g.adagrad_epsilon)     # first step: ~ sign(step)
    assert numpy.allclose(f.weights.mem, w0 + expect, atol=1e-5)
    assert numpy.abs(f.weights.mem - w0).max() <= 1.0 + 1e-6
    f2, g2 = _gd_pair(("momentum", "adadelta"))
    w0 = f2.weights.mem.copy()
    g2.run()
    rho, eps = g2.adadelta_momentum, g2.adadelta_epsilon
    eg = (1 - rho) * step ** 2
    expect = step * numpy.sqrt(eps) / numpy.sqrt(eg + eps)
    assert numpy.allclose(f2.weights.mem, w0 + expect, atol=1e-6)
    f3, g3 = _gd_pair(("momentum", "fast"))
    w0 = f3.weights.mem.copy()
    g3.run()
    assert numpy.isfinite(f3.weights.mem).all() and not numpy.allclose(f3.weights.mem, w0)
    with pytest.raises(ValueError):
        _gd_pair(("adagrad", "adadelta"))
    with pytest.raises(ValueError):
        _gd_pair(("bogus",))


@pytest.mark.parametrize("filling", ["uniform", "gaussian", "constant"])
def test_weight_fillings(filling):
    wf = DummyWorkflow()
    f = all2all.All2All(wf, output_sample_shape=64, weights_filling=filling, weights_stddev=0.05,
                        bias_filling="constant", bias_stddev=0.5)
    f.input = Array(numpy.zeros((2, 100), numpy.float32))
    f.initialize(device=None)
    w = f.weights.mem
    assert w.shape == (64, 100) and numpy.isfinite(w).all()
    if filling == "uniform":
        assert numpy.abs(w).max() <= 0.05 + 1e-7 and w.std() > 0.02
    elif filling == "gaussian":
        assert 0.04 < w.std() < 0.06 and numpy.abs(w).max() > 0.1
    else:
        assert (w == 0.05).all()
    assert (f.bias.mem == 0.5).all()


def test_gabor_filling_of_conv_kernels():
    from veles.znicz_b200.ops import conv
    wf = DummyWorkflow()
    c = conv.Conv(wf, n_kernels=8, kx=9, ky=9, weights_filling="gabor", weights_stddev=0.1)
    c.input = Array(numpy.zeros((1, 16, 16, 1), numpy.float32))
    c.initialize(device=None)
    w = c.weights.mem.reshape(8, 9, 9)
    assert numpy.isfinite(w).all() and numpy.abs(w).max() > 0
    # oriented band-pass filters: (close to) zero mean, different orientations differ
    assert numpy.abs(w.mean(axis=(1, 2))).max() < 0.2 * numpy.abs(w).max()
    assert not numpy.allclose(w[0], w[1])


@pytest.mark.parametrize("codec", ["", "gz", "bz2", "xz"])
def test_snapshot_codecs_round_trip(tmp_path, codec):
    from veles.znicz_b200.core.snapshotter import SnapshotterToFile
    from veles.znicz_b200.models import mnist
    old = root.common.disable.snapshotting
    root.common.disable.snapshotting = False
    try:
        wf = mnist.build(
            layers=mnist.fc_layers(), loader_name="synthetic_mnist",
            loader_config={"minibatch_size": 10, "n_train": 40, "n_valid": 20, "noise": 0.3,
                           "normalization_type": "linear"},
            decision_config={"max_epochs": 1, "fail_iterations": 5},
            snapshotter_config={"prefix": "cdc", "interval": 1, "time_interval": 0,
                                "compression": codec, "directory": str(tmp_path)})
        wf.initialize(device="numpy")
        wf.run()
    finally:
        root.common.disable.snapshotting = old
    files = [n for n in os.listdir(tmp_path) if n.startswith("cdc_") and ".pickle" in n]
    assert files and all(n.endswith(".pickle" + ("." + codec if codec else "")) for n in files)
    back = SnapshotterToFile.import_file(str(tmp_path / "cdc_current.lnk"))
    assert type(back).__name__ == type(wf).__name__
    w0 = wf.forwards[0].weights.mem
    assert numpy.array_equal(back.forwards[0].weights.mem, w0)
    assert int(back.loader.epoch_number) == int(wf.loader.epoch_number)


def test_snapshot_to_database(tmp_path):
    import sqlite3
    from veles.znicz_b200.core.snapshotter import SnapshotterToDB
    wf = DummyWorkflow()
    wf.payload = {"answer": 42}
    snap = SnapshotterToDB(wf, prefix="dbsnap", odbc=str(tmp_path / "s.sqlite"), table="veles",
                           compression="gz")
    snap.suffix = "epoch3"
    snap.initialize()
    snap.export()
    con = sqlite3.connect(str(tmp_path / "s.sqlite"))
    rows = con.execute("select * from veles").fetchall()
    con.close()
    assert len(rows) == 1 and any("dbsnap" in str(c) for c in rows[0])
    blob = [c for c in rows[0] if isinstance(c, (bytes, memoryview))][0]
    import gzip
    back = pickle.loads(gzip.decompress(bytes(blob)))
    assert back.payload == {"answer": 42}


def test_decision_improvement_rules_and_stop_conditions():
    """DecisionGD: validation improves -> `improved`; `fail_iterations` epochs without improvement
    or `max_epochs` -> complete (/root/reference/decision.py:276-293,478-550)."""
    from veles.znicz_b200.models import mnist
    wf = mnist.build(
        layers=mnist.fc_layers(), loader_name="synthetic_mnist",
        loader_config={"minibatch_size": 10, "n_train": 80, "n_valid": 40, "noise": 0.3,
                       "normalization_type": "linear"},
        decision_config={"max_epochs": 50, "fail_iterations": 3})
    wf.initialize(device="numpy")
    wf.run()
    dec = wf.decision
    assert bool(dec.complete)
    # stopped either by the epoch limit or by the patience window
    assert dec.epoch_number <= 50
    stalled = dec.epoch_number - max(e for e in dec.best_n_err_pt_epoch_number if e is not None)
    assert dec.epoch_number == 50 or stalled >= 3 or dec.best_n_err_pt[1] == 0
    assert dec.best_n_err_pt[1] is not None and dec.best_n_err_pt[1] < 50.0
    names = dec.get_metric_names()
    vals = dec.get_metric_values()
    assert set(names) <= set(vals) or len(vals) >= 1


def test_trivial_decision_counts_epochs():
    from veles.znicz_b200.workflow.decision import TrivialDecision
    from veles.znicz_b200.core.mutable import Bool
    wf = DummyWorkflow()
    d = TrivialDecision(wf, max_epochs=2)
    d.minibatch_class, d.last_minibatch, d.class_lengths = 2, Bool(True), [0, 0, 10]
    d.epoch_number, d.epoch_ended, d.minibatch_size = 0, Bool(True), 10
    d.initialize()
    d.run()
    assert not bool(d.complete)
    d.epoch_number = 2
    d.run()
    assert bool(d.complete)
