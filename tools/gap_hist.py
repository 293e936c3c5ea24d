"""Top-2 / top-3 gap statistics of the benchmark rows against k-means centroids, in the tensor kernel's accumulator
units, and the share of rows a single-product (fp16 x fp16) certification threshold would send to the tail."""
import os, sys, math
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "audio-tokens_b200"))
import torch
from at_b200 import MelPlan, LloydTrainer, synth_clips
from at_b200.kmeans import rand_perm

n_clips = int(sys.argv[1]) if len(sys.argv) > 1 else 4000
k = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 10
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
st = torch.zeros(4, device="cuda")
for it in range(iters + 1):
    if it in (0, 2, iters):
        c = tr.get_centroids()
        absmax = float(x.abs().max())
        sx = 2.0 ** (6 - math.frexp(absmax)[1] + 1)   # Sx * max|x| in [64, 128)
        cmax = float(c.norm(dim=1).max()) * 1.0009765625
        e = 6 - math.frexp(cmax)[1]
        if math.ldexp(cmax, e + 1) <= 96.0: e += 1
        S = min(2.0 ** e, sx)
        P = S * S
        Sc = P / sx
        xt = (x * sx).half().float()
        ch = (-2.0 * Sc * c).half().float()
        ecol = (ch - (-2.0 * Sc * c)).norm(dim=1)          # |e_j|
        ecol_inf = (ch - (-2.0 * Sc * c)).abs().max(dim=1).values
        drow = ((x * sx) - xt).norm(dim=1) * (S / sx)       # S |delta|
        xtn = xt.norm(dim=1)
        xt1 = xt.abs().sum(dim=1)
        cn = (c.double() ** 2).sum(1)
        g2s, g3s, taus_old, taus_new, taus_new1, err_act = [], [], [], [], [], []
        for r0 in range(0, n, 1 << 18):
            xb = x[r0:r0 + (1 << 18)].double()
            d = ((xb ** 2).sum(1, keepdim=True) + cn[None, :] - 2.0 * xb @ c.double().T) * P
            v, ix = torch.topk(d, 3, dim=1, largest=False)
            g2s.append((v[:, 1] - v[:, 0]).float()); g3s.append((v[:, 2] - v[:, 0]).float())
            eb = drow[r0:r0 + (1 << 18)]
            t_old = 1.0625 * (0.07 + 4.0 * eb * torch.sqrt(v[:, 0].float().clamp_min(0) + 1.0))
            ej = ecol[ix[:, 0]] + ecol[ix[:, 1]]
            t_new = t_old + 1.0625 * xtn[r0:r0 + (1 << 18)] * 2 * ecol.max()
            taus_old.append(t_old); taus_new.append(t_new)
            # actual error of the single product for the best two
            xtb = xt[r0:r0 + (1 << 18)].double()
            a = (xtb[:, None, :] * (ch.double()[ix[:, :2]] - (-2.0 * Sc * c.double())[ix[:, :2]])).sum(2)
            err_act.append((a[:, 0] - a[:, 1]).abs().float())
        g2, g3, to, tn, ea = map(torch.cat, (g2s, g3s, taus_old, taus_new, err_act))
        print(f"iter {it}: S={S} Sx={sx} P={P} mean P*d1 n/a; max|e_j|={float(ecol.max()):.4f} mean|e_j|={float(ecol.mean()):.4f} "
              f"|x~| mean {float(xtn.mean()):.1f}; tau_old mean {float(to.mean()):.3f} tau_new mean {float(tn.mean()):.3f}; "
              f"actual c-rounding diff error: mean {float(ea.mean()):.4f} max {float(ea.max()):.4f}")
        print(f"   tail share old {(g2 <= to).float().mean():.4%} (third {(g3 <= to).float().mean():.4%});  new {(g2 <= tn).float().mean():.4%} (third {(g3 <= tn).float().mean():.4%})")
        for t in (0.5, 1, 2, 4, 8, 16):
            print(f"   P(gap2 <= {t}) = {(g2 <= t).float().mean():.4%}   P(gap3 <= {t}) = {(g3 <= t).float().mean():.4%}")
    tr.step(x, st)
