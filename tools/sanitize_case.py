"""The smallest cases of every hand-written kernel family, for compute-sanitizer (memcheck / racecheck / synccheck):

    compute-sanitizer --tool memcheck  python tools/sanitize_case.py
    compute-sanitizer --tool racecheck python tools/sanitize_case.py

mel (uniform + ragged + short clips), tcgen05 search (streamed and resident operand tiles, padded K), both tail kernels,
one Lloyd step (full regroup + incremental), resampler, token collate.  Prints a digest so a clean run can be told from
a silent failure."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "audio-tokens_b200"))
import torch
from at_b200 import FlatL2, LloydTrainer, MelPlan, ResamplePlan, _lib, synth_clips
from at_b200.token_dataset import TokenStore

wave = synth_clips(4242, 0, 12, 22050)
plan = MelPlan(22050, 1024, 512, 64, True)
spec, bad, l2 = plan.forward(wave, want_l2=True)
rag = plan.forward_ragged([wave[0, :9000], wave[1, :300], wave[2]], want_l2=True)
x = l2.reshape(-1, 64).contiguous()
n = x.shape[0]
tot = float(spec.sum()) + float(rag[0].sum())
for k, mode in ((96, 2), (700, 1)):
    c = x[torch.randperm(n, device="cuda")[:k]].contiguous() + 1e-3
    ix = FlatL2(64)
    ix.set_tc_mode(mode)
    ix.set_centroids(c)
    a, da = ix.search(x, algo=_lib.ALGO_TENSOR)
    b, db = ix.search(x, algo=_lib.ALGO_SIMT)
    assert torch.equal(a, b), "tensor search != exact search"
    tot += float(da.sum())
tr = LloydTrainer(64, 64, algo=_lib.ALGO_TENSOR)
tr.begin(x)
tr.set_centroids(x[:64].contiguous())
st = torch.zeros(4, device="cuda")
for _ in range(3):
    tr.step(x, st)
tot += float(tr.get_centroids().sum())
rp = ResamplePlan(44100, 22050)
tot += float(rp.forward(torch.randn(2, 5000, device="cuda")).sum())
ts = TokenStore(torch.arange(100, device="cuda"), [0, 10, 35, 100])
tot += float(ts.collate(torch.tensor([2, 0], device="cuda"))[0].sum())
torch.cuda.synchronize()
print("sanitize_case ok", tot)
