"""Wide-row tensor search (d = 64 NS) against the exact wide-row kernel: labels must be equal.  python tools/check_wide.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "audio-tokens_b200"))
import torch
from at_b200 import FlatL2, _lib, row_l2norm

g = torch.Generator(device="cuda").manual_seed(5)
shapes = [(3000, 128, 64), (5000, 640, 128), (20000, 640, 500), (4099, 320, 1000), (700, 1024, 96), (100000, 640, 512)]
if len(sys.argv) > 1:
    shapes = shapes[: int(sys.argv[1])]
for n, d, k in shapes:
    base = torch.rand(n, 64, device="cuda", generator=g)
    x = row_l2norm((base.repeat(1, d // 64) * (1.0 + 0.3 * torch.rand(n, d, device="cuda", generator=g))).contiguous())
    c = (x[torch.randperm(n, device="cuda", generator=g)[:k]] + 0.01 * torch.randn(k, d, device="cuda", generator=g)).contiguous()
    ix = FlatL2(d)
    ix.set_centroids(c)
    ls, _ = ix.search(x, algo=_lib.ALGO_SIMT, want_dist=False)
    torch.cuda.synchronize()
    lt, _ = ix.search(x, algo=_lib.ALGO_TENSOR, want_dist=False)
    torch.cuda.synchronize()
    mism = int((ls != lt).sum())
    _, full0 = ix.tc_stats()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        ix.search(x, algo=_lib.ALGO_TENSOR, want_dist=False)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    _, full1 = ix.tc_stats()
    e0.record()
    for _ in range(3):
        ix.search(x, algo=_lib.ALGO_SIMT, want_dist=False)
    e1.record()
    torch.cuda.synchronize()
    ms_exact = e0.elapsed_time(e1) / 3
    print(f"n={n} d={d} k={k}: mismatches {mism}; rows scanned exactly {(full1 - full0) / 3 / n:.2%}; tensor search {ms:.3f} ms = "
          f"{2.0 * n * k * d / ms / 1e9:.1f} TFLOP/s algorithmic; exact kernel {ms_exact:.3f} ms", flush=True)
    assert mism == 0
print("wide tensor search == exact wide-row kernel")
