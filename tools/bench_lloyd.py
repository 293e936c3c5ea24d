"""Lloyd-loop micro-benchmark of one library build (AT_B200_LIB selects an experiment build): 20 iterations at the C2 shape
with the per-kernel CUDA-event profile (search / update / finalize), the tail statistics, and a digest of the final
centroids so two builds can be compared for bit-equality.

    python tools/bench_lloyd.py [n_clips] [k] [iters] [check]

check = 1 also runs the loop with the exact SIMT search and asserts bit-identical centroids (labels equal in every
iteration, since the update is exact integer arithmetic).
"""
import hashlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "audio-tokens_b200"))
import ctypes

import torch
from at_b200 import FlatL2, LloydTrainer, MelPlan, _lib, synth_clips
from at_b200.kmeans import rand_perm

n_clips = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
k = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 20
check = len(sys.argv) > 4 and sys.argv[4] == "1"
lib = _lib.load()
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
init = x[torch.from_numpy(rand_perm(n, 1235)[:k].astype("int64")).cuda()].contiguous()


def prof(tag):
    cnt, ms = ctypes.c_int64(), ctypes.c_double()
    _lib.check(lib.at_profile_summary(tag, ctypes.byref(cnt), ctypes.byref(ms)))
    return cnt.value, ms.value


def run(algo, profile):
    tr = LloydTrainer(64, k, algo=algo)
    tr.begin(x)
    tr.set_centroids(init)
    st = torch.zeros(iters, 4, device="cuda")
    for it in range(3):   # warm-up (also builds the row image)
        tr.step(x, st[it])
    tr.set_centroids(init)
    torch.cuda.synchronize()
    if profile:
        lib.at_profile_enable(1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for it in range(iters):
        tr.step(x, st[it])
    e1.record()
    torch.cuda.synchronize()
    out = {"loop_ms": e0.elapsed_time(e1)}
    if profile:
        lib.at_profile_enable(0)
        for name, tag in (("search", 0), ("update", 2), ("finalize", 3)):
            c, ms = prof(tag)
            out[name] = (c, ms / max(c, 1))
    cents = tr.get_centroids()
    out["digest"] = hashlib.sha256(cents.cpu().numpy().tobytes()).hexdigest()[:16]
    out["obj"] = float(st[iters - 1, 0])
    return out, cents


print(f"lib {_lib.LIB_PATH}  rows {n}  k {k}  iters {iters}", flush=True)
plain, _ = run(_lib.ALGO_TENSOR, False)
r, cents = run(_lib.ALGO_TENSOR, True)
flops = 2.0 * n * k * 64
print(f"loop (no profiling events): {plain['loop_ms']:.2f} ms = {plain['loop_ms'] / iters:.3f} ms/iter = "
      f"{iters / plain['loop_ms'] * 1e3:.1f} iter/s", flush=True)
print(f"search avg {r['search'][1]:.4f} ms = {flops / r['search'][1] / 1e9:.1f} TFLOP/s algorithmic "
      f"(frac of 1379.1 sustained: {flops / r['search'][1] / 1e9 / 1379.1:.3f}); update avg {r['update'][1]:.4f} ms; "
      f"finalize avg {r['finalize'][1]:.4f} ms; objective {r['obj']:.1f}; centroids {r['digest']}", flush=True)
# one-off search (row image built per call) + tail statistics
ix = FlatL2(64)
ix.set_centroids(cents)
lab = torch.empty(n, dtype=torch.int32, device="cuda")
for _ in range(2):
    ix.search(x, algo=_lib.ALGO_TENSOR, labels=lab, want_dist=False)
s0 = ix.tc_stats()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    ix.search(x, algo=_lib.ALGO_TENSOR, labels=lab, want_dist=False)
e1.record()
torch.cuda.synchronize()
s1 = ix.tc_stats()
print(f"one-off search incl. row image: {e0.elapsed_time(e1) / 5:.3f} ms; per search candidate rows "
      f"{(s1[0] - s0[0]) / 5 / n * 100:.2f} %, exact-scan rows {(s1[1] - s0[1]) / 5 / n * 100:.3f} %", flush=True)
if check:
    ex, _ = run(_lib.ALGO_SIMT, False)
    print(f"exact SIMT loop: {ex['loop_ms'] / iters:.2f} ms/iter, centroids {ex['digest']}", flush=True)
    assert ex["digest"] == r["digest"], "tensor-path centroids differ from the exact search's"
    print("tensor loop == exact loop (bit-identical centroids)", flush=True)
