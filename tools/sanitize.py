"""Small workloads for ``compute-sanitizer --tool {memcheck,racecheck,synccheck}``: every
hand-rolled synchronisation protocol of the repo on sizes a sanitizer run can finish -
gemm_umma_k (mbarrier ring, TMA + LDGSTS producers, TMEM), gemm_pair_k (cluster barriers,
cta_group::2 TMA / commit multicast, remote mbarrier arrive), multi_update_k (grid barrier;
fake-peer flag protocol with 2 ranks sharing the GPU), fc_small, pooling / LRN / evaluator."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from veles.znicz_b200.kernels import load_extension  # noqa: E402

ext = load_extension(required=True)
dev = torch.device("cuda:0")
torch.manual_seed(0)


def gemm_single():
    m, n, k = 256, 96, 192
    a = torch.randn(m, k, device=dev).bfloat16()
    b = torch.randn(n, k, device=dev).bfloat16()
    o = torch.empty(m, n, device=dev, dtype=torch.bfloat16)
    assert ext.gemm(a, k, False, b, k, True, o, n, False, m, n, k, None, 0, 1.0, 0.0, 1, 0, 1) == 0
    torch.cuda.synchronize()
    assert ((o.float() - a.float() @ b.float().t()).abs().max() < 0.5)


def conv_umma():
    n, h, w, c, f = 4, 12, 12, 16, 32
    x = torch.randn(n, h, w, c, device=dev).bfloat16()
    wt = torch.randn(f, 3 * 3 * c, device=dev).bfloat16().contiguous()
    out = torch.empty(n, h, w, f, device=dev, dtype=torch.bfloat16)
    g = [n, h, w, c, h, w, f, 3, 3, 1, 1, 1, 1]
    assert ext.conv_fprop(x, wt, wt.shape[1], False, None, out, g, 0, 1) == 0
    torch.cuda.synchronize()


def gemm_pair():
    m, n, k = 512, 256, 128
    a = torch.randn(m, k, device=dev).bfloat16()
    b = torch.randn(n, k, device=dev).bfloat16()
    o = torch.empty(m, n, device=dev, dtype=torch.bfloat16)
    assert ext.gemm_pair(a, b, o, None, 0, 1.0) == 0
    torch.cuda.synchronize()
    assert ((o.float() - a.float() @ b.float().t()).abs().max() < 0.5)


def multi_update():
    sys.argv = ["fake_peer_worker.py", "2", "1"]
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
    import fake_peer_worker
    fake_peer_worker.main()


if __name__ == "__main__":
    which = sys.argv[1:] or ["gemm_single", "conv_umma", "gemm_pair", "multi_update"]
    for name in which:
        globals()[name]()
        print("ok", name, flush=True)
