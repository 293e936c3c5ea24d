"""How many rows change cluster per Lloyd iteration on the benchmark data (decides the incremental-update threshold)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "audio-tokens_b200"))
import torch
from at_b200 import MelPlan, LloydTrainer, synth_clips
from at_b200.kmeans import rand_perm

n_clips = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
k = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
plan = MelPlan(22050, 1024, 512, 64, True)
l2s = []
for b0 in range(0, n_clips, 2000):
    w = synth_clips(4242, b0, min(2000, n_clips - b0), 220500)
    _, _, l2 = plan.forward(w, want_l2=True)
    l2s.append(l2.reshape(-1, 64))
x = torch.cat(l2s).contiguous()
n = x.shape[0]
perm = torch.from_numpy(rand_perm(n, 1235)[:k].astype("int64")).cuda()
tr = LloydTrainer(64, k)
tr.begin(x)
tr.set_centroids(x[perm].contiguous())
prev = torch.full((n,), -1, dtype=torch.int32, device="cuda")
lab = torch.empty(n, dtype=torch.int32, device="cuda")
st = torch.zeros(4, device="cuda")
for it in range(20):
    tr.step(x, st, labels=lab)
    ch = int((lab != prev).sum())
    print(f"iter {it}: changed {ch} ({ch / n:.4%}) obj {float(st[0]):.1f} nsplit {int(st[1])}", flush=True)
    prev.copy_(lab)
