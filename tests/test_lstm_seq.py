"""LSTMSequence / GDLSTMSequence: numeric differentiation of the numpy oracle, equivalence
with the chained single-step ``LSTM`` cells of the reference design, and a sequence
classification workflow through the layer DSL."""
import numpy

from numdiff import numeric_grad
from veles.znicz_b200.core.config import root
from veles.znicz_b200.core.memory import Array
from veles.znicz_b200.core.workflow import DummyWorkflow
from veles.znicz_b200.ops.lstm_seq import LSTMSequence, GDLSTMSequence


def _build(seq, rs, batch=3, t=4, n_in=5, n_hid=4):
    wf = DummyWorkflow()
    u = LSTMSequence(wf, output_sample_shape=n_hid, weights_stddev=0.4, return_sequences=seq)
    u.input = Array(rs.uniform(-1, 1, (batch, t, n_in)))
    u.initialize(device="numpy")
    return wf, u


def test_lstm_seq_gradients_numdiff():
    root.common.engine.precision_type = "double"
    try:
        for seq in (True, False):
            rs = numpy.random.RandomState(7)
            wf, u = _build(seq, rs)
            u.run()
            r = rs.uniform(-1, 1, u.output.shape)

            def loss():
                u.run()
                return float((u.output.mem * r).sum())
            gd = GDLSTMSequence(wf, learning_rate=0.0, weights_decay=0.0, apply_gradient=False,
                                gradient_moment=0.0)
            gd.err_output = Array(r.copy())
            gd.link_attrs(u, "input", "output", "weights", "bias", "gates", "cells", "hidden",
                          "xh")
            gd.initialize(device="numpy")
            u.run()
            gd.run()
            for arr, grad in ((u.input, gd.err_input.mem), (u.weights, gd.gradient_weights.mem),
                              (u.bias, gd.gradient_bias.mem)):
                num = numeric_grad(loss, arr.mem)
                assert numpy.abs(num - grad.reshape(num.shape)).max() < 1e-6, seq
    finally:
        root.common.engine.precision_type = "float"


def test_lstm_seq_matches_chained_cells():
    """Same math as T chained reference-style LSTM cells sharing weights."""
    from veles.znicz_b200.ops.lstm import LSTM
    root.common.engine.precision_type = "double"
    try:
        rs = numpy.random.RandomState(11)
        batch, t, n_in, n_hid = 2, 3, 4, 3
        wf, u = _build(True, rs, batch, t, n_in, n_hid)
        u.bias.mem[...] = rs.uniform(-0.3, 0.3, u.bias.shape)
        u.run()
        w, b = u.weights.mem, u.bias.mem
        h = numpy.zeros((batch, n_hid))
        c = numpy.zeros((batch, n_hid))
        wf2 = DummyWorkflow()
        for s in range(t):
            cell = LSTM(wf2, output_sample_shape=n_hid, weights_stddev=0.1, simple=True)
            cell.input = Array(u.input.mem[:, s, :].copy())
            cell.prev_output = Array(h.copy())
            cell.prev_memory = Array(c.copy())
            cell.link_from(wf2.start_point)
            cell.initialize(device="numpy")
            for k, gate in enumerate((cell.input_gate, cell.forget_gate, cell.memory_maker,
                                      cell.output_gate)):
                gate.weights.mem[...] = w[k * n_hid:(k + 1) * n_hid]
                gate.bias.mem[...] = b[k * n_hid:(k + 1) * n_hid]
            cell.run()
            h, c = cell.output.mem.copy(), cell.memory.mem.copy()
            numpy.testing.assert_allclose(u.output.mem[:, s, :], h, rtol=1e-9, atol=1e-12)
    finally:
        root.common.engine.precision_type = "float"


def test_lstm_seq_in_standard_workflow():
    """Sequence classification: which half of the sequence carries the pulse."""
    from veles.znicz_b200.loader.fullbatch import FullBatchLoader
    from veles.znicz_b200.workflow.standard_workflow import StandardWorkflow
    from veles.znicz_b200.core.workflow import DummyLauncher

    class PulseLoader(FullBatchLoader):
        MAPPING = "pulse_sequences"

        def load_data(self):
            rs = numpy.random.RandomState(5)
            n, t, f = 240, 8, 3
            x = rs.randn(n, t, f).astype(numpy.float32) * 0.1
            y = rs.randint(0, 2, n)
            for k in range(n):
                pos = rs.randint(0, t // 2) + (t // 2) * y[k]
                x[k, pos, :] += 1.5
            self.original_data.reset(x)
            self.original_labels = y.tolist()
            self.class_lengths[:] = [0, 60, 180]

    root.common.disable.snapshotting = True
    wf = StandardWorkflow(
        DummyLauncher(), loader_name="pulse_sequences",
        loader_config={"minibatch_size": 20, "normalization_type": "none"},
        layers=[{"name": "lstm", "type": "lstm_seq",
                 "->": {"output_sample_shape": 12, "weights_stddev": 0.3},
                 "<-": {"learning_rate": 0.1, "gradient_moment": 0.5, "weights_decay": 0.0}},
                {"name": "out", "type": "softmax",
                 "<-": {"learning_rate": 0.1, "gradient_moment": 0.5, "weights_decay": 0.0}}],
        loss_function="softmax",
        decision_config={"max_epochs": 12, "fail_iterations": 20},
        snapshotter_config={"prefix": "lstmseq", "interval": 1000, "time_interval": 1e9})
    wf.initialize(device="numpy")
    assert type(wf.gds[0]).__name__ == "GDLSTMSequence"
    wf.run()
    assert wf.decision.best_n_err_pt[1] < 20.0, wf.decision.best_n_err_pt


def test_lstm_seq_model_builds_and_trains():
    """models/lstm_seq.py (north-star config 5) on the numpy backend."""
    from veles.znicz_b200.models import lstm_seq
    root.common.disable.snapshotting = True
    try:
        wf = lstm_seq.build(seq_len=6, features=8, hidden=12, n_classes=3,
                            loader_config={"minibatch_size": 16, "n_train": 96, "n_valid": 32,
                                           "noise": 0.2},
                            decision_config={"max_epochs": 3, "fail_iterations": 10})
        wf.initialize(device="numpy")
        assert [type(f).__name__ for f in wf.forwards] == ["LSTMSequence", "All2AllSoftmax"]
        w0 = wf.forwards[0].weights.mem.copy()
        wf.run()
        assert bool(wf.decision.complete)
        assert numpy.isfinite(wf.forwards[0].weights.mem).all()
        assert numpy.abs(wf.forwards[0].weights.mem - w0).max() > 0
    finally:
        root.common.disable.snapshotting = False
