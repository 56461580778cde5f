"""Host-side decisions of the device paths that need no GPU: shape gates of the persistent LSTM
kernels, space-to-depth geometry, channel padding of the split-bf16 operands, collective choice."""
import os

import pytest

from veles.znicz_b200.kernels import load_extension


@pytest.fixture(scope="module")
def ext():
    e = load_extension(required=False)
    if e is None:
        pytest.skip("extension not built")
    return e


def test_lstm_persistent_shape_gates_and_buffers(ext):
    # 4 clusters x 8 CTAs of 32 rows; 8 epilogue warps x 12 vectors x 32 lanes x 16 B per CTA and step
    assert ext.lstm_state_floats(32, 128, 128, 256) == 32 * 32 * 8 * 12 * 32 * 4
    assert ext.lstm_part_floats(128, 256) == 2 * 32 * 8 * 16 * 32 * 4
    os.environ["ZNICZ_LSTM_ROWS"] = "128"
    try:
        assert ext.lstm_state_floats(32, 128, 128, 256) == 32 * 8 * 8 * 12 * 32 * 4
    finally:
        del os.environ["ZNICZ_LSTM_ROWS"]
    assert ext.lstm_state_floats(32, 128, 100, 256) == 0      # I % 64 != 0: per-step path
    assert ext.lstm_state_floats(32, 128, 128, 512) == 0      # H / 32 > 8 CTAs per cluster
    assert ext.lstm_state_floats(32, 128, 128, 96) == 0       # H % 64 != 0
    assert ext.lstm_state_floats(7, 40, 64, 64) > 0           # smallest supported layer
    # large batches fall back to fatter clusters so that the grid stays within one wave
    assert ext.lstm_state_floats(1, 4096, 64, 256) == 1 * (4096 // 128) * 8 * 8 * 12 * 32 * 4


def test_space_to_depth_geometry():
    from veles.znicz_b200.kernels import api

    class _U(object):
        pass
    g = [128, 227, 227, 3, 55, 55, 96, 11, 11, 4, 4, 0, 0]          # AlexNet conv1
    sd = api._s2d_geom(_U(), g)
    assert sd["g"] == [128, 57, 57, 64, 55, 55, 96, 3, 3, 1, 1, 0, 0] and sd["kw"] == 9 * 64
    assert api._s2d_geom(_U(), [100, 32, 32, 3, 32, 32, 32, 5, 5, 1, 1, 2, 2]) is None   # stride 1
    assert api._s2d_geom(_U(), [8, 64, 64, 16, 31, 31, 8, 4, 4, 2, 2, 0, 0])["g"][3] == 64  # 2*2*16
    assert api._s2d_geom(_U(), [8, 64, 64, 32, 31, 31, 8, 4, 4, 2, 2, 0, 0]) is None     # 128 > 64 channels
    assert api._s2d_geom(_U(), [8, 64, 64, 3, 21, 21, 8, 2, 2, 3, 3, 0, 0]) is None      # kernel < stride
    sd = api._s2d_geom(_U(), [2, 30, 30, 4, 9, 10, 8, 9, 7, 3, 3, 1, 0])
    assert (sd["kyp"], sd["kxp"]) == (3, 3) and sd["g"][1:3] == [11, 12] and (sd["pt"], sd["pl"]) == (1, 0)
    os.environ["ZNICZ_CONV_S2D"] = "0"
    try:
        assert api._s2d_geom(_U(), g) is None
    finally:
        del os.environ["ZNICZ_CONV_S2D"]


def test_split_bf16_channel_padding_rules():
    from veles.znicz_b200.kernels import fp32x, api
    # 4 parts per pixel must give the tap-mode gather 32 / 64 channels per tap or a multiple of 64
    for c in (1, 3, 8, 9, 16, 24, 32, 64, 87, 96, 256):
        cp = fp32x._cpad4(c)
        assert cp >= c and (4 * cp in (32, 64) or (4 * cp) % 64 == 0), (c, cp)
    assert [fp32x._cpad4(c) for c in (1, 8, 9, 87)] == [8, 8, 16, 96]
    # the row-stacked wgrad operand keeps the bf16 path's channel padding rules
    assert [api._fprop_cpad(c) for c in (1, 3, 8, 24, 64, 96, 128)] == [8, 8, 0, 32, 0, 128, 0]
    assert fp32x._wgrad_cpad(3) == 8 and fp32x._wgrad_cpad(64) == 64 and fp32x._wgrad_cpad(96) == 128


def test_collective_choice():
    from veles.znicz_b200.ops.fused_step import FusedStep

    class _Symm(object):
        def __init__(self, mc):
            self.mc = mc

        def reduction_buffer(self, key, numel, with_multicast=False):
            return [11, 22], self.mc

    def pick(numel, multicast, env=None):
        fs = FusedStep.__new__(FusedStep)
        fs.mc_red = 0x1000 if multicast else 0
        if env:
            os.environ["ZNICZ_DP_ALGO"] = env
        else:
            os.environ.pop("ZNICZ_DP_ALGO", None)
        try:
            fs._pick_algo(_Symm(0x2000 if multicast else 0), numel)
        finally:
            os.environ.pop("ZNICZ_DP_ALGO", None)
        return fs

    small, big = 89578, 61_000_000
    a = pick(small, True)
    assert (a.algo_name, a.algo, a.mc_red_used) == ("nvls1", 2, 0x1000)   # one barrier, in-switch reduce
    b = pick(big, True)
    assert (b.algo_name, b.algo, b.mc_red_used, b.mc_sum) == ("twoshot", 1, 0x1000, 0x2000)
    assert pick(small, False).algo_name == "oneshot" and pick(small, False).mc_red_used == 0
    c = pick(big, False)
    assert (c.algo_name, c.algo, c.mc_sum, c.sum_ptrs) == ("twoshot_peer", 1, 0, [11, 22])
    assert pick(big, True, "oneshot").algo_name == "oneshot"
    assert pick(small, False, "twoshot").algo_name == "twoshot_peer"      # no multicast on the platform
    with pytest.raises(RuntimeError):
        pick(small, False, "nvls1")


def test_fused_step_leaves_tied_weights_to_immediate_updates():
    """Two GD units updating one tensor (tied auto-encoder weights) must not become two entries of
    the single whole-network update launch."""
    import numpy
    from veles.znicz_b200.core.memory import Array
    from veles.znicz_b200.ops.fused_step import FusedStep

    class _GD(object):
        def __init__(self, w, b=None):
            self.weights, self.bias = w, b

    w_tied, w1, w2 = (Array(numpy.ones((2, 2), numpy.float32)) for _ in range(3))
    b1 = Array(numpy.ones(2, numpy.float32))
    empty = Array()
    a, b, c, d = _GD(w_tied, b1), _GD(w1, empty), _GD(w_tied, empty), _GD(w2, None)
    shared = FusedStep.units_sharing_weights([a, b, c, d])
    assert shared == {id(a), id(c)}
    assert FusedStep.units_sharing_weights([b, d]) == set()
    e = _GD(w2, b1)                                   # shares only the bias with `a`
    assert FusedStep.units_sharing_weights([a, e]) == {id(a), id(e)}
