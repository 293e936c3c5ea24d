"""World-size-2 coverage of the row-sharded k-means host logic on CPU (gloo): global FAISS subsample / init over
the global row index, assembly of the initial centroids by all-reduce, exchange of the exact int64 accumulators.
The CUDA LloydTrainer is replaced by an injected CPU backend with the same interface (tests may use the oracle)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class CpuExactBackend:
    """begin / set_centroids / step / get_centroids like at_b200.kmeans.LloydTrainer, on CPU tensors:
    oracle search + exact int64 fixed-point sums + all-reduce + oracle split_clusters."""

    def __init__(self, d, k):
        self.d, self.k = d, k

    def begin(self, x_local, n_total):
        m = x_local.abs().max().reshape(1) if x_local.numel() else torch.zeros(1)
        if dist.is_initialized() and dist.get_world_size() > 1:
            dist.all_reduce(m, op=dist.ReduceOp.MAX)
        self.n_total = int(n_total)
        nb = int(np.frexp(self.n_total + 1.0)[1])
        mb = int(np.frexp(float(m))[1]) if float(m) > 0 else 0
        self.e = 61 - nb - mb

    def set_centroids(self, c):
        self.c = c.clone()

    def get_centroids(self):
        return self.c.clone()

    def step(self, x_local, stats_out=None, labels=None):
        from oracle import faiss_ref

        x = x_local.numpy()
        lab, d1, _ = faiss_ref.assign_l2_scalar(x, self.c.numpy()) if len(x) else (np.zeros(0, np.int64), None, None)
        acc = np.zeros(self.k * self.d + self.k + 1, dtype=np.int64)
        fx = np.rint(x.astype(np.float64) * 2.0 ** self.e).astype(np.int64)
        np.add.at(acc[: self.k * self.d].reshape(self.k, self.d), lab, fx)
        np.add.at(acc[self.k * self.d: self.k * self.d + self.k], lab, 1)
        t = torch.from_numpy(acc)
        if dist.is_initialized() and dist.get_world_size() > 1:
            dist.all_reduce(t)
        acc = t.numpy()
        counts = acc[self.k * self.d: self.k * self.d + self.k]
        sums = (acc[: self.k * self.d].reshape(self.k, self.d).astype(np.float64) * 2.0 ** (-self.e)).astype(np.float32)
        cen = np.zeros((self.k, self.d), dtype=np.float32)
        nz = counts > 0
        cen[nz] = sums[nz] * (np.float32(1.0) / counts[nz].astype(np.float32))[:, None]
        cen, _, nsplit = faiss_ref.split_clusters(cen, counts.astype(np.float32), self.n_total)
        self.c = torch.from_numpy(cen)
        if stats_out is not None:
            stats_out[1] = float(nsplit)


def _data(n=6000, d=16, seed=0):
    rng = np.random.default_rng(seed)
    cent = rng.random((24, d), dtype=np.float32)
    x = cent[rng.integers(0, 24, n)] + 0.05 * rng.standard_normal((n, d)).astype(np.float32)
    return np.abs(x).astype(np.float32)


def _worker(rank, world, port, kwargs, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "audio-tokens_b200"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from at_b200.kmeans import Kmeans

    x = _data()
    cuts = [0, 2500, len(x)]  # uneven shards
    shard = torch.from_numpy(x[cuts[rank]:cuts[rank + 1]])
    km = Kmeans(16, 20, group=None, backend=CpuExactBackend(16, 20), **kwargs)
    km.train(shard)
    np.save(os.path.join(out_dir, f"c{rank}.npy"), km.centroids)
    dist.barrier()
    dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("kwargs", [dict(niter=4), dict(niter=3, max_points_per_centroid=100000)])
def test_two_rank_kmeans_equals_single_process(tmp_path, kwargs):
    sys.path.insert(0, os.path.join(ROOT, "audio-tokens_b200"))
    from at_b200.kmeans import Kmeans

    mp.spawn(_worker, args=(2, _free_port(), kwargs, str(tmp_path)), nprocs=2, join=True)
    c0 = np.load(tmp_path / "c0.npy")
    c1 = np.load(tmp_path / "c1.npy")
    assert np.array_equal(c0, c1)  # every rank ends with the same centroids
    single = Kmeans(16, 20, group=False, backend=CpuExactBackend(16, 20), **kwargs)
    single.train(torch.from_numpy(_data()))
    # exact integer sums -> the sharded result is bit-identical to the single-process one
    assert np.array_equal(single.centroids, c0)
    # and it is FAISS's trajectory: the oracle with the same subsample / init (6000 > 20*256 -> 5120 rows by default)
    from oracle import faiss_ref

    ref = faiss_ref.Kmeans(16, 20, **kwargs)
    ref.exact_search = True
    ref.train(_data())
    rel = np.linalg.norm(c0 - ref.centroids, axis=1) / np.linalg.norm(ref.centroids, axis=1)
    assert (rel <= 1e-4).mean() >= 0.9
