"""use_convolution branch alone (processors/cluster_creator.py:28-34,68-81): Conv1d expansion 64 -> 640 values per frame,
row normalisation, exact wide-row search and one Lloyd iteration at K = vocab_size (reference default 500), CUDA events.

    python tools/bench_conv.py [n_clips] [k]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "audio-tokens_b200"))
import torch
import at_b200
from at_b200 import FlatL2, LloydTrainer, MelPlan, synth_clips

n_clips = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
k = int(sys.argv[2]) if len(sys.argv) > 2 else 500
plan = MelPlan(22050, 1024, 512, 64, True)
spec, _ = plan.forward(synth_clips(4242, 0, n_clips, 220500))
x = spec.reshape(-1, 64).contiguous()
n = x.shape[0]
torch.manual_seed(42)
conv = torch.nn.Conv1d(1, 10, 3, padding=1)
w, b = conv.weight.detach().reshape(10, 3).cuda().contiguous(), conv.bias.detach().cuda()


def timed(fn, reps=5):
    for _ in range(2):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, out


ms, wide = timed(lambda: at_b200.conv_expand(x, w, b))
print(f"rows {n}: at_conv_expand {ms:.3f} ms = {n * (64 + 640) * 4 / ms / 1e6:.0f} GB/s (algorithmic: 256 B read + 2,560 B written per row)")
ms, wn = timed(lambda: at_b200.row_l2norm(wide))
print(f"at_row_l2norm (d=640) {ms:.3f} ms = {n * 640 * 8 / ms / 1e6:.0f} GB/s")
ix = FlatL2(640)
ix.set_centroids(wn[torch.randperm(n, device='cuda')[:k]].contiguous())
from at_b200 import _lib
ms, _ = timed(lambda: ix.search(wn, want_dist=False, algo=_lib.ALGO_SIMT), reps=3)
print(f"exact wide-row search (fp32 tile kernel) K={k}: {ms:.2f} ms = {2.0 * n * k * 640 / ms / 1e9:.1f} TFLOP/s fp32 (2NKd)")
lab_e, _ = ix.search(wn, want_dist=False, algo=_lib.ALGO_SIMT)
ms, _ = timed(lambda: ix.search(wn, want_dist=False), reps=3)
lab_t, _ = ix.search(wn, want_dist=False)
print(f"wide-row tensor search (tcgen05, slices accumulated, image built per call) K={k}: {ms:.2f} ms = "
      f"{2.0 * n * k * 640 / ms / 1e9:.1f} TFLOP/s algorithmic; labels equal to the exact kernel's: {bool((lab_e == lab_t).all())}")
tr = LloydTrainer(640, k)
tr.begin(wn)
tr.set_centroids(wn[:k].contiguous())
ms, _ = timed(lambda: tr.step(wn, None), reps=3)
print(f"Lloyd iteration d=640 K={k}: {ms:.2f} ms")
