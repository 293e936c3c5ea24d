"""Benchmark-shaped launches of every hot kernel, the target of the ncu captures kept under profiles/:

    ncu --set full --clock-control none --import-source on -k regex:k_mel --launch-skip 4 --launch-count 1 ... prof_kernels.py 4000 1024 mel
    ncu --set full --clock-control none --import-source on \
        -k regex:"k_tc_rows|k_assign_tc|k_tc_tail|k_tc_full|k_gather_sum|k_diff" --launch-skip 12 --launch-count 10 ... prof_kernels.py 20000 1024

Order of the second form's matching launches: three one-off searches (k_tc_rows, k_assign_tc, k_tc_tail, k_tc_full each),
then Lloyd step 1 (row image, search, full regroup) and Lloyd step 2 (search, incremental update).
"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "audio-tokens_b200"))
import torch
from at_b200 import FlatL2, LloydTrainer, MelPlan, _lib, synth_clips
from at_b200.kmeans import rand_perm

n_clips = int(sys.argv[1]) if len(sys.argv) > 1 else 4000
k = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
mel_only = len(sys.argv) > 3 and sys.argv[3] == "mel"
plan = MelPlan(22050, 1024, 512, 64, True)
l2s = []
for b0 in range(0, n_clips, 2000):
    w = synth_clips(4242, b0, min(2000, n_clips - b0), 220500)
    for _ in range(3 if mel_only else 1):
        _, _, l2 = plan.forward(w, want_l2=True)
    l2s.append(l2.reshape(-1, 64))
    del w
torch.cuda.synchronize()
if mel_only:
    sys.exit(0)
x = torch.cat(l2s).contiguous()
del l2s
n = x.shape[0]
init = x[torch.from_numpy(rand_perm(n, 1235)[:k].astype("int64")).cuda()].contiguous()
ix = FlatL2(64)
ix.set_centroids(init)
lab = torch.empty(n, dtype=torch.int32, device="cuda")
for _ in range(3):
    ix.search(x, algo=_lib.ALGO_TENSOR, labels=lab, want_dist=False)
tr = LloydTrainer(64, k)
tr.begin(x)
tr.set_centroids(init)
st = torch.zeros(4, device="cuda")
for _ in range(2):
    tr.step(x, st)
torch.cuda.synchronize()
print("rows", n, "tail (candidate rows, exact-scan rows), cumulative over 3 searches:", ix.tc_stats())
