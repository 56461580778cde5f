"""Oracle style (a) of the reference's unit tests: fixed, hand-computed vectors
(/root/reference/tests/unit/test_all2all.py:56-93, libZnicz/tests/all2all_tanh.cc:39-49) for the
numpy paths that serve as oracles of the device kernels."""
import math

import numpy

from veles.znicz_b200.core.memory import Array
from veles.znicz_b200.core.workflow import DummyWorkflow
from veles.znicz_b200.ops import all2all, gd, normalization, pooling

X = numpy.array([[1.0, 2.0, 3.0], [0.0, -1.0, 2.0]], numpy.float32)
W = numpy.array([[1.0, 0.0, -1.0], [0.5, 0.5, 0.5]], numpy.float32)
B = numpy.array([0.1, -0.2], numpy.float32)
Y = numpy.array([[-1.9, 2.8], [-1.9, 0.3]], numpy.float32)            # X . W^T + B by hand


def _fc(cls):
    wf = DummyWorkflow()
    f = cls(wf, output_sample_shape=2, weights_stddev=0.1)
    f.input = Array(X.copy())
    f.initialize(device=None)
    f.weights.map_write()
    f.bias.map_write()
    f.weights.mem[...] = W
    f.bias.mem[...] = B
    f.run()
    return f


def test_all2all_linear_tanh_sigmoid_strict_relu():
    assert numpy.allclose(_fc(all2all.All2All).output.mem, Y, atol=1e-6)
    t = _fc(all2all.All2AllTanh).output.mem
    assert abs(t[0, 1] - 1.7159 * math.tanh(0.6666 * 2.8)) < 1e-5
    assert abs(t[0, 0] + 1.7159 * math.tanh(0.6666 * 1.9)) < 1e-5
    s = _fc(all2all.All2AllSigmoid).output.mem
    assert abs(s[1, 1] - 1.0 / (1.0 + math.exp(-0.3))) < 1e-6
    r = _fc(all2all.All2AllStrictRELU).output.mem
    assert numpy.allclose(r, [[0.0, 2.8], [0.0, 0.3]], atol=1e-6)
    sp = _fc(all2all.All2AllRELU).output.mem                       # "RELU" = softplus
    assert abs(sp[0, 1] - math.log(1.0 + math.exp(2.8))) < 1e-5


def test_softmax_rows_and_argmax():
    f = _fc(all2all.All2AllSoftmax)
    e = numpy.exp(Y - Y.max(axis=1, keepdims=True))
    assert numpy.allclose(f.output.mem, e / e.sum(axis=1, keepdims=True), atol=1e-6)
    assert f.max_idx.mem.tolist() == [1, 1]
    assert numpy.allclose(f.output.mem.sum(axis=1), 1.0, atol=1e-6)


def test_gradient_descent_step_by_hand():
    f = _fc(all2all.All2All)
    wf = f.workflow
    g = gd.GradientDescent(wf, learning_rate=0.5, learning_rate_bias=0.25, weights_decay=0.0,
                           gradient_moment=0.0, gradient_moment_bias=0.0)
    err = numpy.array([[1.0, 0.0], [0.0, 2.0]], numpy.float32)
    g.err_output = Array(err.copy())
    g.input, g.output, g.weights, g.bias = f.input, f.output, f.weights, f.bias
    g.initialize(device=None)
    g.run()
    # err_input = err . W; gradW = err^T . X; gradB = column sums of err
    assert numpy.allclose(g.err_input.mem, [[1.0, 0.0, -1.0], [1.0, 1.0, 1.0]], atol=1e-6)
    grad_w = numpy.array([[1.0, 2.0, 3.0], [0.0, -2.0, 4.0]])
    assert numpy.allclose(f.weights.mem, W - 0.5 * grad_w, atol=1e-6)
    assert numpy.allclose(f.bias.mem, B - 0.25 * numpy.array([1.0, 2.0]), atol=1e-6)


def test_max_and_avg_pooling_with_partial_windows():
    x = numpy.arange(25, dtype=numpy.float32).reshape(1, 5, 5, 1)
    x[0, 1, 1, 0] = 100.0
    wf = DummyWorkflow()
    mp = pooling.MaxPooling(wf, kx=2, ky=2, sliding=(2, 2))
    mp.input = Array(x.copy())
    mp.initialize(device=None)
    mp.run()
    # ceil mode: 5 -> 3 outputs, the last window is a single row / column
    assert mp.output.mem.shape == (1, 3, 3, 1)
    assert mp.output.mem[0, :, :, 0].tolist() == [[100.0, 8.0, 9.0], [16.0, 18.0, 19.0],
                                                   [21.0, 23.0, 24.0]]
    assert int(mp.input_offset.mem[0, 0, 0, 0]) == 6                # flat index of the 100
    ap = pooling.AvgPooling(wf, kx=2, ky=2, sliding=(2, 2))
    ap.input = Array(x.copy())
    ap.initialize(device=None)
    ap.run()
    assert abs(ap.output.mem[0, 0, 0, 0] - (0 + 1 + 5 + 100) / 4.0) < 1e-5
    assert abs(ap.output.mem[0, 2, 2, 0] - 24.0) < 1e-6             # 1 x 1 window at the corner
    assert abs(ap.output.mem[0, 0, 2, 0] - (4 + 9) / 2.0) < 1e-6    # 2 x 1 window at the edge


def test_lrn_by_hand():
    x = numpy.array([1.0, 2.0, 3.0, 4.0], numpy.float32).reshape(1, 1, 1, 4)
    wf = DummyWorkflow()
    n = normalization.LRNormalizerForward(wf, alpha=0.5, beta=0.75, k=2.0, n=3)
    n.input = Array(x.copy())
    n.initialize(device=None)
    n.run()
    sums = [1 + 4, 1 + 4 + 9, 4 + 9 + 16, 9 + 16]                   # window of 3 channels, clipped
    expect = [v * (2.0 + 0.5 * s) ** -0.75 for v, s in zip([1, 2, 3, 4], sums)]
    assert numpy.allclose(n.output.mem.ravel(), expect, rtol=1e-5)


def test_reference_fc_fixed_vectors_on_every_backend_available():
    """The reference's own fixed vectors for the linear FC layer (data of
    /root/reference/tests/unit/test_all2all.py:60-88): 5 x 5 input, 3 x 5
    weights, 15 expected outputs - numpy path here, the device path in
    tests/test_gpu_units.py::test_fc* compares against this oracle."""
    x = numpy.array([[1, 2, 3, 2, 1], [0, 1, 2, 1, 0], [0, 1, 0, 1, 0], [2, 0, 1, 0, 2],
                     [1, 0, 1, 0, 1]], numpy.float32)
    w = numpy.array([[1, 0, 2, 1, -1], [3, 1, 0, 2, 3], [-1, 2, 0, 1, 3]], numpy.float32)
    b = numpy.array([10, -10, 5], numpy.float32)
    expect = numpy.array([18, 2, 13, 15, -7, 8, 11, -7, 8, 12, 2, 9, 12, -4, 7], numpy.float32)
    for transposed in (False, True):
        wf = DummyWorkflow()
        f = all2all.All2All(wf, output_sample_shape=[3], weights_stddev=0.05,
                            weights_transposed=transposed)
        f.input = Array(x.copy())
        f.initialize(device=None)
        f.weights.map_write()
        f.bias.map_write()
        f.weights.mem[...] = w.T if transposed else w
        f.bias.mem[...] = b
        f.run()
        assert numpy.abs(f.output.mem.ravel() - expect).max() < 1e-4
