"""How many rows an exact bound test (Hamerly: one upper bound to the assigned centroid, one lower bound to every other)
would let a Lloyd iteration skip on the benchmark data.  Decides whether row skipping is worth building: the search is
paid per row scanned, the bounds cost one pass over 12 B per row.

    python tools/bound_skip_sim.py [n_clips] [k] [iters]

Per iteration: share of rows (a) skipped by the drifted bounds alone, (b) skipped after tightening the upper bound with one
exact distance to the assigned centroid, (c) needing the full search; plus the label churn for comparison.
"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "audio-tokens_b200"))
import torch
from at_b200 import MelPlan, LloydTrainer, synth_clips
from at_b200.kmeans import rand_perm

n_clips = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
k = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 20
plan = MelPlan(22050, 1024, 512, 64, True)
l2s = []
for b0 in range(0, n_clips, 2000):
    w = synth_clips(4242, b0, min(2000, n_clips - b0), 220500)
    _, _, l2 = plan.forward(w, want_l2=True)
    l2s.append(l2.reshape(-1, 64))
x = torch.cat(l2s).contiguous()
del l2s
n = x.shape[0]


def top2(c):
    """true first / second smallest Euclidean distance and the arg-min of every row (fp32 matmul in blocks)."""
    d1 = torch.empty(n, device="cuda"); d2 = torch.empty(n, device="cuda"); a = torch.empty(n, dtype=torch.int64, device="cuda")
    cn = (c * c).sum(1)
    for b0 in range(0, n, 1 << 19):
        xb = x[b0:b0 + (1 << 19)]
        d = ((xb * xb).sum(1, keepdim=True) + cn[None, :] - 2.0 * xb @ c.T).clamp_min_(0)
        v, i = torch.topk(d, 2, dim=1, largest=False)
        d1[b0:b0 + xb.shape[0]] = v[:, 0].sqrt(); d2[b0:b0 + xb.shape[0]] = v[:, 1].sqrt(); a[b0:b0 + xb.shape[0]] = i[:, 0]
    return d1, d2, a


torch.backends.cuda.matmul.allow_tf32 = False
perm = torch.from_numpy(rand_perm(n, 1235)[:k].astype("int64")).cuda()
tr = LloydTrainer(64, k)
tr.begin(x)
c_old = x[perm].contiguous()
tr.set_centroids(c_old)
st = torch.zeros(4, device="cuda")
lab = torch.empty(n, dtype=torch.int32, device="cuda")
u = l = assigned = None
for it in range(iters):
    c = tr.get_centroids() if it else c_old
    if it == 0:
        u, l, assigned = top2(c)
        full = torch.ones(n, dtype=torch.bool, device="cuda")
        s1 = s2 = 0.0
    else:
        delta = (c - c_prev).norm(dim=1)
        top = torch.topk(delta, 2).values
        other = torch.where(delta[assigned] >= top[0], top[1], top[0])   # largest movement among the OTHER centroids
        u = u + delta[assigned]
        l = l - other
        skip1 = u < l
        exact_u = (x - c[assigned]).norm(dim=1)
        u = torch.where(skip1, u, exact_u)
        skip2 = (~skip1) & (u < l)
        full = ~(skip1 | skip2)
        s1, s2 = float(skip1.float().mean()), float(skip2.float().mean())
        t1, t2, ta = top2(c)
        # sanity: a skipped row keeps its label
        assert bool((ta[~full] == assigned[~full]).all())
        u = torch.where(full, t1, u); l = torch.where(full, t2, l); assigned = torch.where(full, ta, assigned)
    c_prev = c.clone()
    tr.step(x, st, labels=lab)
    print(f"iter {it}: skip by bounds {s1:.3f}, after one exact distance {s2:.3f}, full search {float(full.float().mean()):.3f}; "
          f"max centroid move {float((tr.get_centroids() - c_prev).norm(dim=1).max()):.4f} nsplit {int(st[1])}", flush=True)
