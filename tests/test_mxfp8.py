"""Block-scaled fp8 numerics reference (utils/mxfp8.py) and the CPU emulation of fp8 fully
connected layers: format properties and a convergence comparison against full precision."""
import numpy

from veles.znicz_b200.core.config import root
from veles.znicz_b200.utils import mxfp8

RS = numpy.random.RandomState(7)


def test_e4m3_block_format_properties():
    x = (RS.randn(5, 100) * numpy.array([1e-3, 1.0, 50.0, 1e4, 1e-6])[:, None]).astype(numpy.float32)
    q, e, n = mxfp8.quantize(x, axis=1)
    assert q.shape == (5, 128) and e.shape == (5, 4) and n == 100 and e.dtype == numpy.int8
    assert numpy.abs(q).max() <= 448.0                       # saturating e4m3 range
    # every quantised value is an e4m3 number: re-quantising changes nothing
    assert numpy.array_equal(mxfp8._e4m3(q), q)
    back = mxfp8.dequantize(q, e, n, axis=1)
    assert back.shape == x.shape
    # 3 mantissa bits: relative error of a normal value <= 2^-4; elements far below the block's
    # maximum lose more (they share its scale) - bound against the block maximum instead
    blk = numpy.abs(numpy.pad(x, ((0, 0), (0, 28))).reshape(5, 4, 32)).max(axis=2)
    err = numpy.abs(numpy.pad(back - x, ((0, 0), (0, 28))).reshape(5, 4, 32)).max(axis=2)
    assert (err <= blk * 2.0 ** -4 + 1e-30).all()
    # scales are powers of two chosen so that the block maximum lands in (224, 448]
    top = numpy.abs(q.reshape(5, 4, 32)).max(axis=2)
    assert ((top > 223.9) & (top <= 448.0))[blk > 0].all()
    z = mxfp8.fake_quant(numpy.zeros((2, 40), numpy.float32))
    assert (z == 0).all()
    col = mxfp8.fake_quant(x, axis=0)                         # quantisation along the other axis
    assert col.shape == x.shape and not numpy.array_equal(col, back)


def test_block_scaled_matmul_error_is_small_and_unbiased():
    a = RS.randn(64, 256).astype(numpy.float32)
    b = RS.randn(256, 48).astype(numpy.float32)
    ref = a.dot(b)
    got = mxfp8.matmul(a, b)
    rel = numpy.linalg.norm(got - ref) / numpy.linalg.norm(ref)
    assert rel < 0.05                                         # ~3 % for gaussian operands
    assert abs(float((got - ref).mean())) < 0.05 * float(numpy.abs(ref).mean())
    # per-block scaling keeps a row with a huge dynamic range usable (per-tensor scaling would
    # flush the small block to zero)
    a2 = a.copy()
    a2[:, :32] *= 1e-4
    a2[:, 32:64] *= 1e4
    ref2 = a2[:, :32].dot(b[:32])
    got2 = mxfp8.fake_quant(a2, axis=1)[:, :32].dot(b[:32])
    assert numpy.linalg.norm(got2 - ref2) / numpy.linalg.norm(ref2) < 0.05


def _train(emulate):
    from veles.znicz_b200.core import prng
    from veles.znicz_b200.models import mnist
    root.common.engine.fp8_emulation = emulate
    prng.get(1).seed(1234)
    prng.get(2).seed(5678)
    try:
        wf = mnist.build(
            layers=mnist.fc_layers(), loader_name="synthetic_mnist",
            loader_config={"minibatch_size": 20, "n_train": 400, "n_valid": 200, "noise": 0.6,
                           "normalization_type": "linear"},
            decision_config={"max_epochs": 6, "fail_iterations": 10})
        wf.initialize(device="numpy")
        wf.run()
        return wf.decision.best_n_err_pt[1], [f.weights.mem.copy() for f in wf.forwards]
    finally:
        root.common.engine.fp8_emulation = False


def test_fp8_emulated_training_converges_like_full_precision():
    """FC net on synthetic MNIST: block-scaled fp8 GEMM operands in forward, err_input and weight
    gradient (fp32 master weights and accumulation) reach the same validation error region."""
    err_ref, w_ref = _train(False)
    err_q, w_q = _train(True)
    assert err_ref < 15.0, err_ref
    assert err_q < 15.0 and abs(err_q - err_ref) <= 5.0, (err_q, err_ref)
    # it really took another path: the weights differ, but only by the quantisation noise
    d = max(float(numpy.abs(a - b).max()) for a, b in zip(w_ref, w_q))
    scale = max(float(numpy.abs(a).max()) for a in w_ref)
    assert 0 < d < 0.25 * scale
