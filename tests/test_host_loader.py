"""Native host-side minibatch assembly (kernels/csrc/host_loader.h): synchronous gather and the
prefetch pool. These are plain C++ (no GPU needed), exercised through the torch extension."""
import numpy
import pytest
import torch


@pytest.fixture(scope="module")
def ext():
    from veles.znicz_b200.kernels import load_extension
    e = load_extension(required=False)
    if e is None or not hasattr(e, "host_gather_rows"):
        pytest.skip("extension not built")
    return e


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_host_gather_rows(ext, dtype):
    src = torch.randn(300, 5, 7, 3)
    src[3, 0, 0, 0] = float("nan")
    src[4, 0, 0, 0] = float("inf")
    idx = torch.randint(0, 300, (40,), dtype=torch.int32)
    idx[0], idx[1], idx[2] = 3, 4, 299
    dst = torch.full((40, 5, 7, 3), 7.0).to(dtype)
    ext.host_gather_rows(src, idx, dst, 33)
    ref = src[idx.long()].to(dtype)
    ref[33:] = 0                                  # rows beyond the minibatch are zeroed
    assert torch.equal(dst.nan_to_num(9.0), ref.nan_to_num(9.0))


def test_host_gather_clamps_bad_indices(ext):
    src = torch.arange(20, dtype=torch.float32).view(10, 2)
    idx = torch.tensor([-5, 3, 99], dtype=torch.int32)
    dst = torch.zeros(3, 2)
    ext.host_gather_rows(src, idx, dst, 3)
    assert dst.tolist() == [[0.0, 1.0], [6.0, 7.0], [18.0, 19.0]]


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_prefetch_pool_matches_synchronous_gather(ext, dtype):
    if not hasattr(ext, "host_prefetch_submit"):
        pytest.skip("no prefetcher in this build")
    src = torch.randn(500, 64)
    rs = numpy.random.RandomState(3)
    for n in (50, 1, 37):
        idx = torch.from_numpy(rs.randint(0, 500, 50).astype(numpy.int32))
        a = torch.empty(50, 64, dtype=dtype)
        b = torch.empty(50, 64, dtype=dtype)
        ticket = ext.host_prefetch_submit(src, idx, a, n)
        idx_copy = idx.clone()
        idx.zero_()                               # the pool works on its own copy of the indices
        ext.host_prefetch_wait(ticket)
        ext.host_gather_rows(src, idx_copy, b, n)
        assert torch.equal(a, b)


def test_peek_next_indices_predicts_the_next_minibatch():
    """The prefetcher assembles minibatch N+1 from ``peek_next_indices()``: it must name exactly
    the indices the next run() serves, across class boundaries, and decline at epoch wraps."""
    from veles.znicz_b200.core.workflow import DummyWorkflow
    from veles.znicz_b200.loader.synthetic import SyntheticImageLoader
    wf = DummyWorkflow()
    ld = SyntheticImageLoader(wf, minibatch_size=32, shape=(4, 4, 3), n_classes=5,
                              n_train=80, n_valid=40, n_test=10)
    ld.initialize(device="numpy")
    ld.run()
    hits = wraps = 0
    for _ in range(40):
        peek = ld.peek_next_indices()
        expected = None
        if peek is not None:
            start, n = peek
            expected = ld.shuffled_indices.mem[start:start + n].copy()
        ld.run()
        got = ld.minibatch_indices.mem[:int(ld.minibatch_size)]
        if expected is None:
            wraps += 1
        else:
            hits += 1
            numpy.testing.assert_array_equal(got, expected)
    assert hits > 25 and wraps >= 5          # 130 samples / 32: 6 minibatches per epoch
