"""One mel pass + a few tensor searches on benchmark-shaped data: the target of `ncu --set full -k regex:...`."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "audio-tokens_b200"))
import torch
from at_b200 import FlatL2, MelPlan, _lib, synth_clips

n_clips = int(sys.argv[1]) if len(sys.argv) > 1 else 4000
k = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
plan = MelPlan(22050, 1024, 512, 64, True)
l2s = []
for b0 in range(0, n_clips, 2000):
    w = synth_clips(4242, b0, min(2000, n_clips - b0), 220500)
    for _ in range(2):
        _, _, l2 = plan.forward(w, want_l2=True)
    l2s.append(l2.reshape(-1, 64))
x = torch.cat(l2s).contiguous()
n = x.shape[0]
c = x[torch.randperm(n, device="cuda")[:k]].contiguous()
ix = FlatL2(64)
ix.set_centroids(c)
lab = torch.empty(n, dtype=torch.int32, device="cuda")
for _ in range(3):
    ix.search(x, algo=_lib.ALGO_TENSOR, labels=lab, want_dist=False)
torch.cuda.synchronize()
print("rows", n, "tail", ix.tc_stats())
