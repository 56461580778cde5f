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
