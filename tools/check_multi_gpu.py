"""Hardware check of the multi-GPU k-means (run under torchrun, one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 \
        tools/check_multi_gpu.py [total_clips] [k] [iters]

1. the N-rank result (rows sharded by rank, per-iteration exchange over peer memory, and again over NCCL) is BIT-EQUAL to the
   single-GPU result over the same rows (rank 0 recomputes it alone);
2. the cost of the exchange step: Lloyd loop time per iteration with the peer-memory kernel vs dist.all_reduce.
"""
import hashlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "audio-tokens_b200"))
import torch
import torch.distributed as dist

from at_b200 import LloydTrainer, MelPlan, synth_clips
from at_b200.kmeans import rand_perm

total_clips = int(sys.argv[1]) if len(sys.argv) > 1 else 4000
k = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 20
rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
L = 220500
plan = MelPlan(22050, 1024, 512, 64, True)


def rows_of(first, count):
    parts = []
    for b0 in range(first, first + count, 1000):
        w = synth_clips(4242, b0, min(1000, first + count - b0), L)
        parts.append(plan.forward(w, want_l2=True)[2].reshape(-1, 64))
    return torch.cat(parts).contiguous()


per = total_clips // world
x = rows_of(rank * per, per)
n_total = per * world * 431
off = rank * per * 431
init_rows = rand_perm(n_total, 1235)[:k].astype("int64")


def initial(xl, offset):
    import numpy as np

    mine = (init_rows >= offset) & (init_rows < offset + xl.shape[0])
    c = torch.zeros((k, 64), device="cuda")
    c.index_copy_(0, torch.from_numpy(np.nonzero(mine)[0]).cuda(), xl.index_select(0, torch.from_numpy(init_rows[mine] - offset).cuda()))
    return c


def run(xl, offset, group, reduce):
    tr = LloydTrainer(64, k, group=group, reduce=reduce)
    c0 = initial(xl, offset)
    if group is not False and world > 1:
        dist.all_reduce(c0)
    tr.begin(xl, n_total)
    tr.set_centroids(c0)
    for _ in range(2):
        tr.step(xl, None)
    tr.set_centroids(c0)
    if world > 1 and group is not False:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        tr.step(xl, None)
    e1.record()
    torch.cuda.synchronize()
    if tr.peer is not None:
        tr.peer.status()
    ms = torch.tensor([e0.elapsed_time(e1) / iters], device="cuda", dtype=torch.float64)
    if world > 1 and group is not False:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    c = tr.get_centroids()
    return tr.reduce, float(ms), hashlib.sha256(c.cpu().numpy().tobytes()).hexdigest()[:16]


res = {}
if world > 1:
    for mode in ("peer", "nccl"):
        try:
            res[mode] = run(x, off, None, mode)
        except RuntimeError as ex:
            res[mode] = ("unavailable", float("nan"), repr(ex)[:100])
if rank == 0:
    full = x if world == 1 else rows_of(0, per * world)
    single = run(full, 0, False, "auto")
    print(f"rows {n_total} k {k} iters {iters} world {world}")
    print(f"single GPU        : {single[1]:.3f} ms/iter  centroids {single[2]}")
    for mode, r in res.items():
        print(f"{world} ranks, {mode:5s} ({r[0]}): {r[1]:.3f} ms/iter  centroids {r[2]}  {'== single GPU' if r[2] == single[2] else '!= single GPU'}")
    bad = [m for m, r in res.items() if r[0] != "unavailable" and r[2] != single[2]]
    assert not bad, f"multi-rank centroids differ from the single-GPU result: {bad}"
    if res:
        print("multi-rank centroids are bit-equal to the single-GPU result")
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
