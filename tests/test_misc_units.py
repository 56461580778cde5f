"""Deconv/GDDeconv, Kohonen, RBM, RProp, ResizableAll2All, ZeroFiller, Depooling, Cutter1D,
Multiplier/Summator, stochastic pooling (numpy oracles; mirrors the reference's
tests/unit/test_deconv.py, test_kohonen.py, test_rbm.py, test_resizable_all2all.py,
test_zero_filling.py, test_multiplier.py, test_summator.py, test_cutter.py)."""
import numpy
import pytest

from veles.znicz_b200.core.config import root
from veles.znicz_b200.core.memory import Array
from veles.znicz_b200.core.workflow import DummyWorkflow
from veles.znicz_b200.ops import (conv, deconv, gd_deconv, kohonen, rbm_units, rprop_gd,
                                  resizable_all2all, weights_zerofilling, depooling, pooling,
                                  cutter, multiplier, summator, all2all, gd)
from numdiff import numeric_grad

RS = numpy.random.RandomState(21)


def test_deconv_padding_algebra():
    """/root/reference/tests/unit/test_deconv.py:97-125."""
    assert deconv.Deconv.compute_padding(8, 8, 4, 4, (2, 2)) == (2, 2, 2, 2)
    assert deconv.Deconv.compute_padding(9, 7, 4, 4, (2, 2)) == (2, 2, 3, 3)
    deconv.Deconv.check_padding_is_safe(4, 4, (2, 2))
    with pytest.raises(ValueError):
        deconv.Deconv.check_padding_is_safe(3, 3, (2, 2))


@pytest.mark.parametrize("unsafe", [False, True])
def test_deconv_and_gd_numdiff(unsafe):
    root.common.engine.precision_type = "double"
    try:
        wf = DummyWorkflow()
        n, sy, sx, c, f = 2, 8, 8, 3, 4
        k, sl = (4, (2, 2)) if not unsafe else (3, (2, 2))
        pad = deconv.Deconv.compute_padding(sx, sy, k, k, sl)
        cv = conv.Conv(wf, n_kernels=f, kx=k, ky=k, padding=pad, sliding=sl,
                       weights_stddev=0.3, include_bias=False)
        cv.input = Array(RS.uniform(-1, 1, (n, sy, sx, c)))
        cv.initialize(device=None)
        cv.run()
        dc = deconv.Deconv(wf, n_kernels=f, kx=k, ky=k, sliding=sl, padding=pad,
                           unsafe_padding=unsafe)
        dc.input = cv.output
        dc.weights = cv.weights
        dc.output_shape_source = cv.input
        dc.initialize(device=None)
        dc.run()
        assert dc.output.shape == cv.input.shape
        assert bool(dc.hits) == unsafe
        r = RS.uniform(-1, 1, dc.output.shape)
        x = cv.output

        def loss():
            dc.run()
            return float((dc.output.mem * r).sum())

        g = gd_deconv.GDDeconv(wf, n_kernels=f, kx=k, ky=k, sliding=sl, padding=pad,
                               learning_rate=1.0, weights_decay=0.0, apply_gradient=False)
        g.input = x
        g.weights = cv.weights
        g.err_output = Array(r.copy())
        g.hits = dc.hits if unsafe else None
        g.initialize(device=None)
        g.run()
        ng = numeric_grad(loss, x.mem)
        assert numpy.abs(ng - g.err_input.mem).max() < 1e-6
        ngw = numeric_grad(loss, cv.weights.mem)
        assert numpy.abs(ngw - g.gradient_weights.mem).max() < 1e-6
    finally:
        root.common.engine.precision_type = "float"


def test_kohonen_trainer_converges_and_forward_matches():
    wf = DummyWorkflow()
    centers = numpy.array([[-0.7, -0.7], [0.7, 0.7], [-0.7, 0.7], [0.7, -0.7]], numpy.float32)
    data = (centers[RS.randint(0, 4, 200)] + RS.normal(0, 0.05, (200, 2))).astype(numpy.float32)
    tr = kohonen.KohonenTrainer(wf, shape=(4, 4), weights_stddev=0.05)
    tr.input = Array(data[:10].copy())
    tr.initialize(device=None)
    for it in range(300):
        tr.input.mem[...] = data[(it * 10) % 200:(it * 10) % 200 + 10]
        tr.run()
    fw = kohonen.KohonenForward(wf)
    fw.input = Array(data)
    fw.weights = tr.weights
    fw.initialize(device=None)
    fw.run()
    w = tr.weights.mem
    d = ((data[:, None, :] - w[None]) ** 2).sum(2)
    assert numpy.array_equal(fw.output.mem, d.argmin(1))
    assert numpy.sqrt(d.min(1)).mean() < 0.15        # neurons moved onto the clusters
    assert tr.winners.mem.sum() == 10 * 300


def test_kohonen_validator_and_decision():
    wf = DummyWorkflow()
    v = kohonen.KohonenValidator(wf)
    v.shape = (2, 2)
    v.input = Array(numpy.array([0, 0, 1, 3, 3, 3], numpy.int32))
    v.minibatch_indices = Array(numpy.arange(6, dtype=numpy.int32))
    v.minibatch_size = 6
    v.samples_by_label = {"a": {0, 1, 2}, "b": {3, 4, 5}}
    v.labels_mapping = {"a": 0, "b": 1}
    v.reversed_labels_mapping = ["a", "b"]
    v.initialize()
    v.run()
    assert v.result["b"] == {3}
    assert v.result["a"] == {0, 1}
    assert abs(v.fitness - 1.0) < 1e-9


def test_rbm_cd1_improves_reconstruction():
    wf = DummyWorkflow()
    n, vsz, hsz = 40, 12, 6
    patterns = (RS.rand(4, vsz) > 0.5).astype(numpy.float32)
    data = patterns[RS.randint(0, 4, n)]
    w = Array(RS.normal(0, 0.05, (hsz, vsz)).astype(numpy.float32))
    hb = Array(numpy.zeros(hsz, numpy.float32))
    vb = Array(numpy.zeros(vsz, numpy.float32))
    make_h = all2all.All2AllSigmoid(wf, output_sample_shape=hsz, weights_stddev=0.05)
    make_h.input = Array(data.copy())
    make_h.weights, make_h.bias = w, hb
    make_h.initialize(device=None)
    grad = rbm_units.GradientRBM(wf, stddev=0.05, v_size=vsz, h_size=hsz, cd_k=1)
    grad.input = make_h.output
    grad.weights, grad.hbias, grad.vbias = w, hb, vb
    grad.batch_size = n
    bw0 = rbm_units.BatchWeights(wf)
    bw0.v, bw0.h, bw0.batch_size = make_h.input, make_h.output, n
    bw1 = rbm_units.BatchWeights2(wf)
    bw1.batch_size = n
    gc = rbm_units.GradientsCalculator(wf)
    upd = rbm_units.WeightsUpdater(wf, learning_rate=0.2)

    def recon_err():
        h = 1 / (1 + numpy.exp(-(data.dot(w.mem.T) + hb.mem)))
        v = 1 / (1 + numpy.exp(-(h.dot(w.mem) + vb.mem)))
        return float(((v - data) ** 2).mean())

    grad.initialize(device=None)
    make_h.run()
    grad.run()
    bw1.v, bw1.h = grad.v1, grad.h1
    bw0.initialize(device=None)
    bw1.initialize(device=None)
    gc.weights0, gc.vbias0, gc.hbias0 = bw0.weights_batch, bw0.vbias_batch, bw0.hbias_batch
    gc.weights1, gc.vbias1, gc.hbias1 = bw1.weights_batch, bw1.vbias_batch, bw1.hbias_batch
    bw0.run()
    bw1.run()
    gc.initialize(device=None)
    upd.weights_grad, upd.vbias_grad, upd.hbias_grad = gc.weights_grad, gc.vbias_grad, gc.hbias_grad
    upd.weights, upd.hbias, upd.vbias = w, hb, vb
    e0 = recon_err()
    for _ in range(150):
        make_h.run()
        grad.run()
        bw0.run()
        bw1.run()
        gc.run()
        upd.run()
    assert recon_err() < 0.7 * e0


def test_rprop_decreases_on_sign_flip():
    wf = DummyWorkflow()
    f = all2all.All2All(wf, output_sample_shape=3, weights_stddev=0.1)
    f.input = Array(RS.uniform(-1, 1, (5, 4)).astype(numpy.float32))
    f.initialize(device=None)
    f.run()
    g = rprop_gd.GDRProp(wf)
    g.input, g.output, g.weights, g.bias = f.input, f.output, f.weights, f.bias
    g.err_output = Array(RS.uniform(-1, 1, (5, 3)).astype(numpy.float32))
    g.need_err_input = False
    g.initialize(device=None)
    g.run()
    lr1 = g.weight_lrs.mem.copy()
    g.run()                      # same gradient sign -> increase
    assert (g.weight_lrs.mem >= lr1).all()
    g.err_output.mem[...] *= -1  # flipped sign -> decrease (the reference never does)
    lr2 = g.weight_lrs.mem.copy()
    g.run()
    assert (g.weight_lrs.mem < lr2).any()


def test_resizable_all2all_preserves_weights():
    wf = DummyWorkflow()
    u = resizable_all2all.ResizableAll2All(wf, output_sample_shape=4, weights_stddev=0.1)
    u.input = Array(RS.uniform(-1, 1, (3, 5)).astype(numpy.float32))
    u.initialize(device=None)
    w0 = u.weights.mem.copy()
    u.output_sample_shape = 6
    assert u.weights.shape == (6, 5) and u.output.shape == (3, 6)
    assert numpy.array_equal(u.weights.mem[:4], w0)
    u.run()
    u.output_sample_shape = 2
    assert u.weights.shape == (2, 5)
    assert numpy.array_equal(u.weights.mem, w0[:2])


def test_zero_filler_mask():
    wf = DummyWorkflow()
    z = weights_zerofilling.ZeroFiller(wf, grouping=2)
    z.weights = Array(numpy.ones((4, 6), numpy.float32))
    z.initialize(device=None)
    z.run()
    k = numpy.arange(4)[:, None] % 2
    c = numpy.arange(6)[None, :] % 2
    assert numpy.array_equal(z.weights.mem, (k != c).astype(numpy.float32))
    with pytest.raises(ValueError):
        weights_zerofilling.ZeroFiller(wf, grouping=1)


def test_depooling_inverts_max_pooling_positions():
    wf = DummyWorkflow()
    p = pooling.MaxPooling(wf, kx=2, ky=2, sliding=(2, 2))
    p.input = Array(RS.uniform(-1, 1, (2, 6, 6, 3)).astype(numpy.float32))
    p.initialize(device=None)
    p.run()
    d = depooling.Depooling(wf)
    d.input = p.output
    d.output_offset = p.input_offset
    d.output_shape_source = p.input
    d.initialize(device=None)
    d.run()
    nz = d.output.mem != 0
    assert nz.sum() == p.output.size
    assert numpy.allclose(d.output.mem[nz], p.input.mem[nz])


def test_cutter1d_multiplier_summator():
    wf = DummyWorkflow()
    c = cutter.Cutter1D(wf, alpha=2.0, beta=0.5, input_offset=1, output_offset=2, length=3)
    c.input = Array(numpy.arange(12, dtype=numpy.float32).reshape(2, 6))
    c.output = Array(numpy.ones((2, 6), numpy.float32))
    c.initialize(device=None)
    c.run()
    exp = numpy.ones((2, 6), numpy.float32)
    exp[:, 2:5] = 2.0 * c.input.mem[:, 1:4] + 0.5
    assert numpy.allclose(c.output.mem, exp)
    x = Array(RS.uniform(-1, 1, (3, 4)).astype(numpy.float32))
    y = Array(RS.uniform(-1, 1, (3, 4)).astype(numpy.float32))
    m = multiplier.Multiplier(wf)
    m.x, m.y = x, y
    m.initialize(device=None)
    m.run()
    assert numpy.allclose(m.output.mem, x.mem * y.mem)
    s = summator.Summator(wf)
    s.x, s.y = x, y
    s.initialize(device=None)
    s.run()
    assert numpy.allclose(s.output.mem, x.mem + y.mem)
    gm = multiplier.GDMultiplier(wf)
    gm.x, gm.y = x, y
    gm.err_output = Array(numpy.ones((3, 4), numpy.float32))
    gm.initialize(device=None)
    gm.run()
    assert numpy.allclose(gm.err_x.mem, y.mem) and numpy.allclose(gm.err_y.mem, x.mem)


def test_stochastic_pooling_statistics_and_pool_depool():
    wf = DummyWorkflow()
    x = numpy.zeros((1, 2, 2, 1), numpy.float32)
    x[0, 0, 0, 0] = 3.0
    x[0, 1, 1, 0] = 1.0
    big = numpy.tile(x, (400, 1, 1, 1))
    p = pooling.StochasticPooling(wf, kx=2, ky=2, sliding=(2, 2), seed=5)
    p.input = Array(big)
    p.initialize(device=None)
    p.run()
    frac3 = float((p.output.mem == 3.0).mean())
    assert 0.65 < frac3 < 0.85                      # P = 3 / (3 + 1)
    assert set(numpy.unique(p.output.mem)) <= {1.0, 3.0}
    q = pooling.StochasticPoolingDepooling(wf, kx=2, ky=2, sliding=(2, 2), seed=6)
    q.input = Array(RS.uniform(0.1, 1, (3, 4, 4, 2)).astype(numpy.float32))
    before = q.input.mem.copy()
    q.initialize(device=None)
    q.run()
    after = q.input.mem
    assert (after != 0).sum() == 3 * 2 * 2 * 2      # one survivor per window per channel
    assert numpy.allclose(after[after != 0], before[after != 0])


def test_nn_rollback_raises_and_lowers_learning_rates_and_restores_weights():
    """/root/reference/nn_rollback.py:44-181: improved -> lr x lr_plus and the weights are stashed;
    `minus_steps` epochs without improvement (or NaNs at once) -> lr x lr_minus and the stashed
    weights come back."""
    from veles.znicz_b200.core.memory import Array
    from veles.znicz_b200.core.mutable import Bool
    from veles.znicz_b200.workflow.nn_rollback import NNRollback

    class _GD(object):
        def __init__(self):
            self.learning_rate, self.learning_rate_bias = 0.1, 0.2
            self.weights = Array(numpy.ones((2, 3), numpy.float32))
            self.bias = Array(numpy.zeros(2, numpy.float32))
            self.gradient_weights = Array()
            self.gradient_bias = Array()
            self.forward_unit = None

    wf = DummyWorkflow()
    rb = NNRollback(wf, lr_plus=1.5, lr_minus=0.5, minus_steps=2)
    rb.improved = Bool(True)
    gd = _GD()
    rb.add_gd(gd)
    rb.initialize()
    rb.run()                                         # improved: raise the rate, stash weights
    assert abs(gd.learning_rate - 0.15) < 1e-9 and abs(gd.learning_rate_bias - 0.3) < 1e-9
    gd.weights.map_write()
    gd.weights.mem[...] = 7.0                        # the epoch moved the weights somewhere bad
    rb.improved <<= False
    rb.run()                                         # 1st bad epoch: nothing yet
    assert abs(gd.learning_rate - 0.15) < 1e-9
    rb.run()                                         # 2nd: lower the rate, restore
    assert abs(gd.learning_rate - 0.075) < 1e-9
    gd.weights.map_read()
    assert (gd.weights.mem == 1.0).all()
    gd.weights.map_write()
    gd.weights.mem[0, 0] = numpy.nan
    rb.run()                                         # NaNs: immediate rollback
    gd.weights.map_read()
    assert numpy.isfinite(gd.weights.mem).all() and abs(gd.learning_rate - 0.0375) < 1e-9
    # per-unit factors and the elastic hooks of the reference's IDistributable surface
    rb.add_gd(gd, lr_plus=2.0, lr_minus=0.25)
    rb.improved <<= True
    rb.run()
    assert abs(gd.learning_rate - 0.075) < 1e-9
    rb.generate_data_for_slave("s1")
    rb.drop_slave("s1")
    assert not rb.slaves_


def test_reference_style_imports_resolve_to_this_package():
    """`from veles.znicz.conv import Conv`, `from veles.config import root` keep working through
    the alias finder (compat.py; the reference's import surface is SURVEY §1.3)."""
    import subprocess
    import sys
    code = (
        "import veles.znicz_b200.compat as c; c.install()\n"
        "from veles.config import root\n"
        "from veles.znicz.conv import Conv\n"
        "from veles.znicz.all2all import All2AllSoftmax\n"
        "from veles.znicz.gd_conv import GradientDescentConv\n"
        "from veles.memory import Array\n"
        "from veles.mutable import Bool\n"
        "import veles.znicz_b200.ops.conv as m, veles.znicz_b200.core.config as k\n"
        "assert Conv is m.Conv and root is k.root\n"
        "print('ok')\n")
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and r.stdout.strip() == "ok", r.stderr[-2000:]
