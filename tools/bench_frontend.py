"""Front-end step (channel mean + sample-rate conversion) on 10 s clips: device kernel vs the reference's way
(torch.mean + a new torchaudio Resample module per clip, on the GPU and on the CPU)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "audio-tokens_b200"))
import torch
from torchaudio.transforms import Resample
from at_b200 import ResamplePlan

for orig, ch in ((44100, 2), (48000, 2), (16000, 1)):
    L = orig * 10
    n_clips = 200
    wave = (torch.randint(-20000, 20000, (n_clips, ch, L), device="cuda").float() / 32768.0)
    plan = ResamplePlan(orig, 22050)
    out = torch.empty((1, plan.out_len(L)), device="cuda")
    for i in range(10):
        plan.forward(wave[i], out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n_clips):
        plan.forward(wave[i], out)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n_clips
    byt = ch * L * 4 + plan.out_len(L) * 4
    outb = torch.empty((n_clips, plan.out_len(L)), device="cuda")
    for _ in range(3):
        plan.forward_batch(wave, outb)
    e0.record()
    for _ in range(5):
        plan.forward_batch(wave, outb)
    e1.record()
    torch.cuda.synchronize()
    msb = e0.elapsed_time(e1) / 5
    print(f"{orig} Hz x{ch}: batch of {n_clips} clips in one launch: {msb:.3f} ms = {byt * n_clips / msb / 1e6:.0f} GB/s algorithmic", flush=True)
    # the reference's way on the same device
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(20):
        w = torch.mean(wave[i], dim=0, keepdim=True)
        y = Resample(orig, 22050).to("cuda")(w)
    torch.cuda.synchronize()
    ref_gpu = (time.perf_counter() - t0) / 20 * 1e3
    wc = wave[:5].cpu()
    t0 = time.perf_counter()
    for i in range(5):
        y = Resample(orig, 22050)(torch.mean(wc[i], dim=0, keepdim=True))
    ref_cpu = (time.perf_counter() - t0) / 5 * 1e3
    print(f"{orig} Hz x{ch} -> 22050: {ms * 1e3:.1f} us per 10 s clip = {byt / ms / 1e6:.1f} GB/s algorithmic; "
          f"reference way (new Resample per clip): {ref_gpu:.2f} ms on this GPU, {ref_cpu:.1f} ms on the CPU", flush=True)
