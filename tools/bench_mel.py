"""Mel kernel alone: 4,000 clips per launch (inputs larger than L2), CUDA events.  python tools/bench_mel.py [clips]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "audio-tokens_b200"))
import torch
from at_b200 import MelPlan, _lib, synth_clips

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4000
plan = MelPlan(22050, 1024, 512, 64, True)
w = synth_clips(4242, 0, n, 220500)
for _ in range(3):
    plan.forward(w, want_l2=True)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    plan.forward(w, want_l2=True)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
import hashlib
spec, bad, l2 = plan.forward(w[:64].contiguous(), want_l2=True)
print("digest spec", hashlib.sha1(spec.cpu().numpy().tobytes()).hexdigest()[:16], "l2", hashlib.sha1(l2.cpu().numpy().tobytes()).hexdigest()[:16])
print(f"lib {_lib.LIB_PATH}: mel {n} clips: {ms:.3f} ms -> {n * 431 / ms / 1e3:.1f} M frames/s; {ms * 20000 / n:.2f} ms per 20,000 clips; "
      f"{n * 992336 / ms / 1e6:.0f} GB/s algorithmic")
