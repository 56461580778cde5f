"""Numeric differentiation of every forward/GD pair on the numpy oracle path
(/root/reference/tests/unit/gd_numdiff.py, test_gd.py, test_gd_conv.py,
test_gd_workflow.py)."""
import numpy
import pytest

from numdiff import check_pair
from veles.znicz_b200.ops import (all2all, gd, conv, gd_conv, pooling, gd_pooling,
                                  activation, normalization, cutter)

RS = numpy.random.RandomState(11)


@pytest.mark.parametrize("fwd,bwd", [
    (all2all.All2All, gd.GradientDescent), (all2all.All2AllTanh, gd.GDTanh),
    (all2all.All2AllRELU, gd.GDRELU), (all2all.All2AllStrictRELU, gd.GDStrictRELU),
    (all2all.All2AllSigmoid, gd.GDSigmoid)])
@pytest.mark.parametrize("transposed", [False, True])
def test_fc(fwd, bwd, transposed):
    x = RS.uniform(-1, 1, (5, 7))
    check_pair(fwd, bwd, x,
               fwd_kwargs={"output_sample_shape": 4, "weights_stddev": 0.5,
                           "weights_transposed": transposed},
               gd_kwargs={"weights_transposed": transposed})


@pytest.mark.parametrize("fwd,bwd", [
    (conv.Conv, gd_conv.GradientDescentConv), (conv.ConvTanh, gd_conv.GDTanhConv),
    (conv.ConvRELU, gd_conv.GDRELUConv), (conv.ConvStrictRELU, gd_conv.GDStrictRELUConv),
    (conv.ConvSigmoid, gd_conv.GDSigmoidConv)])
@pytest.mark.parametrize("padding,sliding", [((0, 0, 0, 0), (1, 1)),
                                             ((2, 1, 1, 2), (2, 1)),
                                             ((1, 1, 1, 1), (2, 2))])
def test_conv(fwd, bwd, padding, sliding):
    x = RS.uniform(-1, 1, (2, 6, 7, 3))
    kw = {"n_kernels": 4, "kx": 3, "ky": 2, "padding": padding, "sliding": sliding,
          "weights_stddev": 0.5}
    gkw = {k: kw[k] for k in ("n_kernels", "kx", "ky", "padding", "sliding")}
    gkw["unpack_size"] = 16
    check_pair(fwd, bwd, x, fwd_kwargs=kw, gd_kwargs=gkw)


@pytest.mark.parametrize("fwd,bwd,links", [
    (pooling.MaxPooling, gd_pooling.GDMaxPooling, ("input_offset",)),
    (pooling.MaxAbsPooling, gd_pooling.GDMaxAbsPooling, ("input_offset",)),
    (pooling.AvgPooling, gd_pooling.GDAvgPooling, ())])
@pytest.mark.parametrize("k,s", [((2, 2), (2, 2)), ((3, 3), (2, 2)), ((3, 2), (1, 2))])
def test_pooling(fwd, bwd, links, k, s):
    x = RS.uniform(-1, 1, (2, 7, 6, 3))
    kw = {"kx": k[0], "ky": k[1], "sliding": s}
    check_pair(fwd, bwd, x, fwd_kwargs=kw, gd_kwargs=kw, link=(), extra_links=links)


def test_lrn():
    x = RS.uniform(-1, 1, (2, 3, 3, 8))
    kw = {"alpha": 0.05, "beta": 0.75, "n": 3, "k": 1}
    check_pair(normalization.LRNormalizerForward, normalization.LRNormalizerBackward,
               x, fwd_kwargs=kw, gd_kwargs=kw, link=())
    kw = {"alpha": 0.01, "beta": 0.75, "n": 5, "k": 2}
    check_pair(normalization.LRNormalizerForward, normalization.LRNormalizerBackward,
               x, fwd_kwargs=kw, gd_kwargs=kw, link=())


@pytest.mark.parametrize("name", ["Tanh", "Sigmoid", "RELU", "StrictRELU", "Log",
                                  "TanhLog", "SinCos", "Mul"])
def test_activation(name):
    x = RS.uniform(-4, 4, (3, 10))
    fkw = {"factor": 0.37} if name == "Mul" else {}
    check_pair(getattr(activation, "Forward" + name),
               getattr(activation, "Backward" + name), x, fwd_kwargs=fkw,
               gd_kwargs=fkw, link=())


def test_cutter():
    x = RS.uniform(-1, 1, (2, 8, 9, 3))
    kw = {"padding": (1, 2, 3, 1)}
    check_pair(cutter.Cutter, cutter.GDCutter, x, fwd_kwargs=kw, gd_kwargs=kw, link=())


def test_err_input_alpha_beta():
    """err_input = alpha * new + beta * old (/root/reference/nn_units.py:400-402)."""
    from veles.znicz_b200.core.memory import Array
    from veles.znicz_b200.core.workflow import DummyWorkflow
    wf = DummyWorkflow()
    f = all2all.All2All(wf, output_sample_shape=3, weights_stddev=0.5)
    f.input = Array(RS.uniform(-1, 1, (4, 5)).astype(numpy.float32))
    f.initialize(device=None)
    f.run()
    g = gd.GradientDescent(wf, err_input_alpha=0.5, err_input_beta=2.0,
                           apply_gradient=False)
    g.err_output = Array(RS.uniform(-1, 1, (4, 3)).astype(numpy.float32))
    g.input, g.output, g.weights, g.bias = f.input, f.output, f.weights, f.bias
    g.initialize(device=None)
    old = RS.uniform(-1, 1, (4, 5)).astype(numpy.float32)
    g.err_input.mem[...] = old
    g.run()
    expect = 0.5 * g.err_output.mem.dot(f.weights.mem) + 2.0 * old
    assert numpy.abs(g.err_input.mem - expect).max() < 1e-5
