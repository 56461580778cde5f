"""LSTM cell + GDLSTM: numeric differentiation of the nested workflows
(/root/reference/tests/unit/test_lstm.py:61-177 checks closed forms; here every output
gradient incl. the err_memory path is verified numerically)."""
import numpy
import pytest

from veles.znicz_b200.core.config import root
from veles.znicz_b200.core.memory import Array
from veles.znicz_b200.core.workflow import DummyWorkflow
from veles.znicz_b200.ops.lstm import LSTM, GDLSTM
from numdiff import numeric_grad


@pytest.mark.parametrize("simple", [True, False])
def test_lstm_cell_gradients(simple):
    root.common.engine.precision_type = "double"
    try:
        rs = numpy.random.RandomState(3)
        batch, n_in, n_hid = 4, 5, 3
        wf = DummyWorkflow()
        cell = LSTM(wf, output_sample_shape=n_hid, weights_stddev=0.5, simple=simple)
        cell.input = Array(rs.uniform(-1, 1, (batch, n_in)))
        cell.prev_output = Array(rs.uniform(-1, 1, (batch, n_hid)))
        cell.prev_memory = Array(rs.uniform(-1, 1, (batch, n_hid)))
        cell.link_from(wf.start_point)
        wf.end_point.link_from(cell)
        wf.initialize(device="numpy")
        wf.run()
        assert cell.output.shape == (batch, n_hid) and cell.memory.shape == (batch, n_hid)
        r1 = rs.uniform(-1, 1, (batch, n_hid))
        r2 = rs.uniform(-1, 1, (batch, n_hid))

        def loss():
            cell.run()
            return float((cell.output.mem * r1).sum() + (cell.memory.mem * r2).sum())

        gd = GDLSTM(wf, cell, learning_rate=0.0, weights_decay=0.0, apply_gradient=False)
        gd.err_output = Array(r1.copy())
        gd.err_memory = Array(r2.copy())
        gd.initialize(device=wf.device)
        cell.run()
        gd.err_output.mem[...] = r1
        gd.run()
        for name, arr, got in (("input", cell.input, gd.err_input),
                               ("prev_output", cell.prev_output, gd.err_prev_output),
                               ("prev_memory", cell.prev_memory, gd.err_prev_memory)):
            ng = numeric_grad(loss, arr.mem)
            got.map_read()
            assert numpy.abs(ng - got.mem.reshape(ng.shape)).max() < 1e-6, name
        # weight gradients of one gate
        ng = numeric_grad(loss, cell.forget_gate.weights.mem)
        assert numpy.abs(ng - gd.gd_forget_gate.gradient_weights.mem).max() < 1e-6
    finally:
        root.common.engine.precision_type = "float"


def test_lstm_sequence_shares_weights():
    rs = numpy.random.RandomState(4)
    wf = DummyWorkflow()
    cells = []
    zeros = Array(numpy.zeros((2, 3), numpy.float32))
    prev_o, prev_m = zeros, zeros
    prev_unit = wf.start_point
    for t in range(3):
        c = LSTM(wf, output_sample_shape=3, weights_stddev=0.3)
        c.input = Array(rs.uniform(-1, 1, (2, 4)).astype(numpy.float32))
        if cells:
            c.link_weights(cells[0])
            c.link_attrs(cells[-1], ("prev_output", "output"), ("prev_memory", "memory"))
        else:
            c.prev_output, c.prev_memory = prev_o, prev_m
        c.link_from(prev_unit)
        prev_unit = c
        cells.append(c)
    wf.end_point.link_from(prev_unit)
    wf.initialize(device="numpy")
    wf.run()
    assert cells[2].input_gate.weights is cells[0].input_gate.weights
    assert numpy.isfinite(cells[2].output.mem).all()
    assert numpy.abs(cells[2].output.mem).max() > 0
