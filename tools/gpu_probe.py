"""Diagnostic: run each piece of the hot path once with timestamps (flush after every line)."""
import faulthandler
import os
import sys
import time

faulthandler.dump_traceback_later(100, repeat=True)
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "audio-tokens_b200"))
T0 = time.time()


def log(*a):
    print(f"[{time.time() - T0:7.2f}s]", *a, flush=True)


log("import torch")
import torch

log("torch", torch.__version__, "cuda", torch.cuda.is_available(), torch.cuda.get_device_name(0))
import numpy as np
from at_b200 import FlatL2, LloydTrainer, MelPlan, _lib, row_l2norm, synth_clips

lib = _lib.load()
log("lib loaded", lib.at_version())
w = synth_clips(4242, 0, 40, 22050)
torch.cuda.synchronize()
log("synth ok", float(w.abs().max()))
plan = MelPlan(22050, 1024, 512, 64, True)
log("plan ok")
spec, bad, l2 = plan.forward(w, want_l2=True)
torch.cuda.synchronize()
log("mel ok", spec.shape, int(bad.sum()), float(spec.min()), float(spec.max()))
from oracle import mel_ref

ref = mel_ref.mel_db_torchaudio(w[0].cpu().numpy(), 22050, 1024, 512, 64, True).T
log("mel err vs torchaudio", float(np.abs(spec[0].cpu().numpy() - ref).max()))
x = l2.reshape(-1, 64).contiguous()
y = row_l2norm(spec.reshape(-1, 64).contiguous())
torch.cuda.synchronize()
log("row_l2norm ok; fused-l2 == row_l2norm:", bool(torch.equal(x, y)))
for k in (64, 256):
    c = x[:: x.shape[0] // k][:k].contiguous()
    ix = FlatL2(64)
    ix.set_centroids(c)
    torch.cuda.synchronize()
    log("set_centroids ok", k)
    ls, ds = ix.search(x, algo=_lib.ALGO_SIMT)
    torch.cuda.synchronize()
    log("simt search ok", k, int(ls.max()))
    if "notc" not in sys.argv:
        lt, dt = ix.search(x, algo=_lib.ALGO_TENSOR)
        torch.cuda.synchronize()
        mism = int((ls != lt).sum())
        log("tensor search ok", k, "label mismatches", mism, "dist equal", bool(torch.equal(ds, dt)))
        if mism:
            bad_rows = torch.nonzero(ls != lt)[:5, 0].tolist()
            log("first mismatching rows", bad_rows, ls[bad_rows].tolist(), lt[bad_rows].tolist(), ds[bad_rows].tolist(), dt[bad_rows].tolist())
tr = LloydTrainer(64, 64, algo=_lib.ALGO_SIMT)
tr.begin(x)
log("begin ok")
tr.set_centroids(x[:: x.shape[0] // 64][:64].contiguous())
st = torch.zeros(4, device="cuda")
for it in range(3):
    tr.step(x, st)
    torch.cuda.synchronize()
    log("lloyd step", it, st.tolist())
from oracle import faiss_ref

lab, d1, d2 = faiss_ref.assign_l2_scalar(x.cpu().numpy(), tr.get_centroids().cpu().numpy())
log("oracle assign ok")
log("DONE")
