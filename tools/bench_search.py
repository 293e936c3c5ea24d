"""Micro-benchmark of the search kernel alone (CUDA events, L2-exceeding input), all operand modes."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "audio-tokens_b200"))
import torch
from at_b200 import FlatL2, MelPlan, _lib, synth_clips

n_clips = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
ks = [int(v) for v in sys.argv[2].split(",")] if len(sys.argv) > 2 else [1024]
plan = MelPlan(22050, 1024, 512, 64, True)
l2s = []
for b0 in range(0, n_clips, 2000):
    w = synth_clips(4242, b0, min(2000, n_clips - b0), 220500)
    _, _, l2 = plan.forward(w, want_l2=True)
    l2s.append(l2.reshape(-1, 64))
    del w
x = torch.cat(l2s).contiguous()
del l2s
n = x.shape[0]
print("rows", n, flush=True)
for k in ks:
    c = x[torch.randperm(n, device="cuda")[:k]].contiguous()
    for mode, name in ((1, "stream"), (2, "resident")):
        ix = FlatL2(64)
        ix.set_tc_mode(mode)
        ix.set_centroids(c)
        lab = torch.empty(n, dtype=torch.int32, device="cuda")
        dist = torch.empty(n, dtype=torch.float32, device="cuda")
        for want in (True, False):
            for _ in range(3):
                ix.search(x, algo=_lib.ALGO_TENSOR, labels=lab, dist=dist, want_dist=want)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5):
                ix.search(x, algo=_lib.ALGO_TENSOR, labels=lab, dist=dist, want_dist=want)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 5
            r8, rfull = ix.tc_stats()
            print(f"k={k} mode={name} dist={want}: {ms:.3f} ms  {2.0 * n * k * 64 / ms / 1e9:.1f} TFLOP/s algorithmic "
                  f"({2.25 * 2.0 * n * k * 64 / ms / 1e9:.1f} executed)  cumulative re-check rows {r8} full {rfull}", flush=True)
