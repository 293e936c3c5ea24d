"""k-means with faiss::Clustering semantics on the device.

* ``LloydTrainer``  -- device-level: one Lloyd iteration = at_kmeans_accumulate (+ all-reduce of the exact int64
  sums / counts / objective across ranks) + at_kmeans_finalize.
* ``Kmeans``        -- the faiss.Kmeans look-alike the reference constructs in
  processors/cluster_creator.py:42-48 and trains in :53-56 (same kwargs, ``train`` / ``centroids`` / ``obj`` /
  ``iteration_stats``), including FAISS's seeded subsample (rand_perm(seed)) and random-point initialisation
  (rand_perm(seed + 1)).  With a torch.distributed process group every rank passes its own shard of the rows
  and all ranks end with the same centroids; because per-cluster sums are exact integers the result does not
  depend on the number of ranks.
"""
from __future__ import annotations

import ctypes
import sys
import time

import numpy as np

from . import _lib


class ClusteringParameters:
    """faiss/Clustering.h ClusteringParameters defaults (the reference overrides niter, verbose, gpu only)."""

    def __init__(self):
        self.niter = 25
        self.nredo = 1
        self.verbose = False
        self.spherical = False
        self.int_centroids = False
        self.update_index = False
        self.frozen_centroids = False
        self.min_points_per_centroid = 39
        self.max_points_per_centroid = 256
        self.seed = 1234
        self.decode_block_size = 32768


def rand_perm(n: int, seed: int) -> np.ndarray:
    """faiss::rand_perm through the library's host routine (std::mt19937)."""
    perm = np.empty(n, dtype=np.int32)
    _lib.check(_lib.load().at_rand_perm_host(_lib.ptr(perm), n, seed))
    return perm


def _world(group):
    import torch.distributed as dist

    if group is None and not (dist.is_available() and dist.is_initialized()):
        return None, 0, 1
    if group is False:
        return None, 0, 1
    return dist, dist.get_rank(group), dist.get_world_size(group)


class PeerReducer:
    """The per-iteration exchange over peer memory (at_peer_*): every rank's accumulator window is mapped by all the
    others through CUDA IPC and one kernel signals, waits and sums them over NVLink.  torch.distributed is used once, to
    exchange the 64-byte handles."""

    def __init__(self, group, words: int):
        import torch
        import torch.distributed as dist

        self.lib = _lib.load()
        g = group or None
        self.rank, self.world = dist.get_rank(g), dist.get_world_size(g)
        self.h = None
        h = ctypes.c_void_p()
        ok = self.world <= 16 and self.lib.at_peer_create(self.rank, self.world, words, ctypes.byref(h)) == 0
        handle = torch.zeros(64, dtype=torch.uint8)
        if ok:
            ok = self.lib.at_peer_export(h, _lib.ptr(handle)) == 0
        # a collective decision at every step: a rank must never be alone in (or out of) the peer protocol
        ok = self._all_ok(ok, g)
        handles = [torch.zeros(64, dtype=torch.uint8, device="cuda") for _ in range(self.world)]
        dist.all_gather(handles, handle.cuda(), group=g)
        if ok:
            for r, hd in enumerate(handles):
                hb = hd.cpu().contiguous()
                if r != self.rank and self.lib.at_peer_import(h, r, _lib.ptr(hb)) != 0:
                    ok = False
                    break
        ok = self._all_ok(ok, g)   # doubles as the barrier at_peer_connect asks for
        if ok:
            ok = self.lib.at_peer_connect(h) == 0
        ok = self._all_ok(ok, g)
        if not ok:
            if h:
                self.lib.at_peer_destroy(h)
            raise RuntimeError("peer-memory reduction unavailable: " + self.lib.at_last_error().decode("utf-8", "replace"))
        self.h = h

    @staticmethod
    def _all_ok(ok, g):
        import torch
        import torch.distributed as dist

        t = torch.tensor([1 if ok else 0], dtype=torch.int32, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MIN, group=g)
        return bool(t.item())

    def local_buffer(self):
        return ctypes.c_void_p(self.lib.at_peer_local_buffer(self.h))

    def total(self):
        return ctypes.c_void_p(self.lib.at_peer_total(self.h))

    def reduce(self):
        _lib.check(self.lib.at_peer_reduce(self.h, _lib.stream_ptr()))

    def status(self):
        _lib.check(self.lib.at_peer_status(self.h, _lib.stream_ptr()))

    def __del__(self):
        try:
            if getattr(self, "h", None):
                self.lib.at_peer_destroy(self.h)
                self.h = None
        except Exception:
            pass


class LloydTrainer:
    """Device-level Lloyd loop state for (k, d).  All tensors are CUDA tensors on the current device.

    reduce: how the ranks' accumulators are summed each iteration when ``group`` spans several ranks -- "peer" (one kernel
    over CUDA-IPC-mapped peer memory, at_peer_*), "nccl" (torch.distributed all_reduce) or "auto" (peer when it can be set
    up on every rank, else nccl; AT_B200_REDUCE overrides)."""

    def __init__(self, d: int, k: int, group=False, algo: int = _lib.ALGO_AUTO, reduce: str = "auto"):
        import os

        import torch

        _lib.require_cuda()
        self.lib = _lib.load()
        self.d, self.k, self.algo = int(d), int(k), algo
        self.group = group
        h = ctypes.c_void_p()
        _lib.check(self.lib.at_kmeans_create(self.d, self.k, ctypes.byref(h)))
        self.h = h
        words = self.lib.at_kmeans_accum_words(self.h)
        self.accum = torch.zeros(words, dtype=torch.int64, device="cuda")
        self._absmax = torch.zeros(1, dtype=torch.float32, device="cuda")
        self.n_total = 0
        self._rows_key = None   # identity of the rows the library's caches (row image, previous labels) were built from
        self.peer = None
        self.reduce = "none"
        dist, _, world = _world(self.group)
        if world > 1:
            reduce = os.environ.get("AT_B200_REDUCE", reduce)
            self.reduce = "nccl"
            if reduce in ("auto", "peer"):
                try:
                    self.peer = PeerReducer(self.group, words)
                    self.reduce = "peer"
                except RuntimeError:
                    if reduce == "peer":
                        raise

    def __del__(self):
        try:
            if getattr(self, "h", None):
                self.lib.at_kmeans_destroy(self.h)
                self.h = None
        except Exception:
            pass

    def begin(self, x_local, n_total: int | None = None, absmax=None):
        """Fix the fixed-point scale from the global max |x| and row count. One host sync.  absmax: CUDA float32[1] already
        holding max |x_local| (the mel kernel produces it on the way out, MelPlan.set_absmax_out); None = scan the rows."""
        import torch

        dist, _, world = _world(self.group)
        self._rows_key = self._key(x_local)
        if x_local.numel() == 0:   # a shard left empty by the subsample: nothing to scan (at_absmax rejects a null pointer)
            self._absmax.zero_()
        elif absmax is not None:
            self._absmax.copy_(absmax.reshape(self._absmax.shape))
        else:
            _lib.check(self.lib.at_absmax(_lib.ptr(x_local), x_local.numel(), _lib.ptr(self._absmax), _lib.stream_ptr()))
        if world > 1:
            dist.all_reduce(self._absmax, op=dist.ReduceOp.MAX, group=self.group or None)
            if n_total is None:
                nt = torch.tensor([x_local.shape[0]], dtype=torch.int64, device="cuda")
                dist.all_reduce(nt, group=self.group or None)
                n_total = int(nt.item())
        elif n_total is None:
            n_total = x_local.shape[0]
        self.n_total = int(n_total)
        # (.item() waits for the current stream only; the stream-ordered begin keeps other streams' work in flight)
        _lib.check(self.lib.at_kmeans_begin_on(self.h, float(self._absmax.item()), self.n_total, _lib.stream_ptr()))

    @staticmethod
    def _key(x):
        # data pointer + shape + torch's in-place version counter: a different tensor at a recycled address, or the same
        # tensor modified in place, changes the key
        return (x.data_ptr(), tuple(x.shape), getattr(x, "_version", 0))

    def invalidate(self):
        """Drop the library's per-training-set caches (at_kmeans_invalidate)."""
        _lib.check(self.lib.at_kmeans_invalidate(self.h))
        self._rows_key = None

    def set_centroids(self, c):
        import torch

        assert c.is_cuda and c.dtype == torch.float32 and tuple(c.shape) == (self.k, self.d) and c.is_contiguous()
        _lib.check(self.lib.at_kmeans_set_centroids(self.h, _lib.ptr(c), _lib.stream_ptr()))

    def set_incremental(self, on: bool):
        """Incremental update (only rows whose label changed move between the exact integer sums) on / off."""
        _lib.check(self.lib.at_kmeans_set_incremental(self.h, int(bool(on))))

    def get_centroids(self):
        import torch

        out = torch.empty((self.k, self.d), dtype=torch.float32, device="cuda")
        _lib.check(self.lib.at_kmeans_get_centroids(self.h, _lib.ptr(out), _lib.stream_ptr()))
        return out

    def step(self, x_local, stats_out=None, labels=None):
        """One Lloyd iteration over this rank's rows.  stats_out: CUDA float32[4] (objective, nsplit,
        imbalance factor, empty clusters) or None.  No host synchronisation."""
        dist, _, world = _world(self.group)
        key = self._key(x_local)
        if key != self._rows_key:
            # not the rows begin() / the previous step saw (another tensor, possibly at a recycled address, or an in-place
            # update): the cached fp16 image and the incremental sums would be stale
            _lib.check(self.lib.at_kmeans_invalidate(self.h))
            self._rows_key = key
        if self.peer is not None:
            # accumulate straight into this rank's window; one kernel signals, waits and adds the windows over NVLink
            _lib.check(self.lib.at_kmeans_accumulate(self.h, _lib.ptr(x_local), x_local.shape[0], 0, self.algo,
                                                     self.peer.local_buffer(), _lib.ptr(labels), _lib.stream_ptr()))
            self.peer.reduce()
            _lib.check(self.lib.at_kmeans_finalize(self.h, self.peer.total(), self.n_total, _lib.ptr(stats_out),
                                                   _lib.stream_ptr()))
            return
        _lib.check(self.lib.at_kmeans_accumulate(self.h, _lib.ptr(x_local), x_local.shape[0], 0, self.algo,
                                                 _lib.ptr(self.accum), _lib.ptr(labels), _lib.stream_ptr()))
        if world > 1:
            dist.all_reduce(self.accum, group=self.group or None)
        _lib.check(self.lib.at_kmeans_finalize(self.h, _lib.ptr(self.accum), self.n_total, _lib.ptr(stats_out),
                                               _lib.stream_ptr()))


class Kmeans:
    """faiss.Kmeans(d, k, **kwargs) subset (faiss/python/extra_wrappers.py).

    kwargs are copied onto ClusteringParameters; unknown names raise AttributeError like FAISS.  ``gpu`` is
    accepted and ignored (this implementation only runs on the GPU).  Extra keyword-only knobs that do not exist
    in FAISS and default to its behaviour: ``group`` (torch.distributed process group for row-sharded training),
    ``backend`` (object with begin/set_centroids/step/get_centroids; the CUDA LloydTrainer by default),
    ``algo``.
    """

    def __init__(self, d, k, *, group=False, backend=None, algo=_lib.ALGO_AUTO, **kwargs):
        self.d, self.k = int(d), int(k)
        self.gpu = False
        self.cp = ClusteringParameters()
        for key, v in kwargs.items():
            if key == "gpu":
                self.gpu = v
            else:
                getattr(self.cp, key)  # AttributeError for non-existent fields, like FAISS
                setattr(self.cp, key, v)
        self.group, self.algo = group, algo
        self._backend = backend
        self.centroids = None
        self.obj = None
        self.iteration_stats = None
        self.index = None

    # -- helpers -------------------------------------------------------------------------------
    def _trainer(self):
        if self._backend is None:
            self._backend = LloydTrainer(self.d, self.k, group=self.group, algo=self.algo)
        return self._backend

    def train(self, x, weights=None, init_centroids=None):
        """x: (n_local, d) float32 numpy array (uploaded) or CUDA tensor (this rank's shard).  Returns the final
        objective like FAISS (Kmeans.train returns obj[-1])."""
        import torch

        if weights is not None:
            raise NotImplementedError("weights are not used by the reference and are not implemented")
        if self.cp.spherical or self.cp.int_centroids or self.cp.frozen_centroids or self.cp.nredo != 1:
            raise NotImplementedError("only FAISS's default spherical/int_centroids/frozen_centroids/nredo are implemented")
        dist, rank, world = _world(self.group)
        cp, k, d = self.cp, self.k, self.d
        backend = self._trainer()
        dev = "cuda" if isinstance(backend, LloydTrainer) else "cpu"
        if isinstance(x, np.ndarray):
            x = torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32)).to(dev)
        x = x.contiguous()
        assert x.dim() == 2 and x.shape[1] == d and x.dtype == torch.float32
        n_local = x.shape[0]
        if world > 1:
            sizes = torch.zeros(world, dtype=torch.int64, device=x.device)
            sizes[rank] = n_local
            dist.all_reduce(sizes, group=self.group or None)
            sizes = sizes.cpu().numpy()
        else:
            sizes = np.array([n_local], dtype=np.int64)
        off = int(sizes[:rank].sum())
        n = int(sizes.sum())
        if n < k:
            raise RuntimeError(
                "Error: 'nx >= k' failed: Number of training points (%d) should be at least as large as number "
                "of clusters (%d)" % (n, k))
        finite = torch.isfinite(x).all().to(torch.int32)
        if world > 1:
            dist.all_reduce(finite, op=dist.ReduceOp.MIN, group=self.group or None)
        if not bool(finite.item()):
            raise RuntimeError("Error: 'std::isfinite(x[i])' failed: input contains NaN's or Inf's")

        # Clustering::train_encoded -> subsample_training_set: perm = rand_perm(nx, seed)[: k * max_ppc]
        perm = None
        if n > k * cp.max_points_per_centroid:
            nx = k * cp.max_points_per_centroid
            if cp.verbose:
                print("Sampling a subset of %d / %d for training" % (nx, n))
            perm = rand_perm(n, cp.seed)[:nx].astype(np.int64)
            mine = (perm >= off) & (perm < off + n_local)
            sel = torch.from_numpy(perm[mine] - off).to(x.device)
            x_train = x.index_select(0, sel)
        else:
            nx = n
            x_train = x
            if n < k * cp.min_points_per_centroid:
                print("WARNING clustering %d points to %d centroids: please provide at least %d training points"
                      % (n, k, k * cp.min_points_per_centroid), file=sys.stderr)

        # initial centroids: given ones first, the rest random training points (rand_perm(nx, seed + 1))
        cent = torch.zeros((k, d), dtype=torch.float32, device=x.device)
        n_input = 0
        if init_centroids is not None:
            ic = init_centroids
            if isinstance(ic, np.ndarray):
                ic = torch.from_numpy(np.ascontiguousarray(ic, dtype=np.float32))
            assert ic.shape[1] == d
            n_input = min(k, ic.shape[0])
        perm2 = rand_perm(nx, cp.seed + 1)  # FAISS draws it even when every centroid is given
        if n_input < k:
            pos = perm2[n_input:k].astype(np.int64)        # positions in the (perm-ordered) training set
            rows = perm[pos] if perm is not None else pos   # global row ids
            mine = (rows >= off) & (rows < off + n_local)
            idx = torch.from_numpy(np.nonzero(mine)[0] + n_input).to(x.device)
            src = torch.from_numpy(rows[mine] - off).to(x.device)
            cent.index_copy_(0, idx, x.index_select(0, src))
            if world > 1:
                dist.all_reduce(cent, group=self.group or None)  # other ranks hold zeros in those rows
        if n_input:
            cent[:n_input] = ic[:n_input].to(x.device)

        stats = None
        t0 = time.time()
        if nx == k:
            # faiss::Clustering::train_encoded: with exactly k training points the centroids ARE the training set, in input
            # order (init_centroids and the permutation are ignored), and one all-zero iteration stat is pushed
            full = torch.zeros((k, d), dtype=torch.float32, device=x.device)
            if perm is None:
                full[off:off + n_local] = x
            else:   # (unreachable with FAISS's parameters: subsampling to exactly k rows needs max_points_per_centroid = 1)
                pos = torch.from_numpy(np.nonzero(mine)[0]).to(x.device)
                full.index_copy_(0, pos, x_train)
            if world > 1:
                dist.all_reduce(full, group=self.group or None)
            self._final = full
            stats = torch.zeros((1, 4), dtype=torch.float32, device=x.device)
        else:
            backend.begin(x_train, nx)
            backend.set_centroids(cent)
            stats = torch.zeros((max(cp.niter, 1), 4), dtype=torch.float32, device=x.device)
            for it in range(cp.niter):
                backend.step(x_train, stats[it])
            self._final = backend.get_centroids()
        self.centroids = self._final.cpu().numpy()
        st = stats[: (1 if nx == k else cp.niter)].cpu().numpy() if stats is not None else np.zeros((0, 4), dtype=np.float32)
        elapsed = time.time() - t0
        self.iteration_stats = [
            dict(obj=float(r[0]), time=elapsed * (i + 1) / max(len(st), 1), time_search=0.0,
                 imbalance_factor=float(r[2]), nsplit=int(r[1]))
            for i, r in enumerate(st)
        ]
        self.obj = np.array([s["obj"] for s in self.iteration_stats])
        if cp.verbose and rank == 0:
            for i, s in enumerate(self.iteration_stats):
                print("  Iteration %d (%.2f s, search %.2f s): objective=%g imbalance=%.3f nsplit=%d"
                      % (i, s["time"], s["time_search"], s["obj"], s["imbalance_factor"], s["nsplit"]))
        return self.obj[-1] if self.obj.size > 0 else 0.0

    def centroids_device(self):
        return self._final

    def assign(self, x):
        from .index import IndexFlatL2

        ix = IndexFlatL2(self.d)
        ix.add(self.centroids)
        D, I = ix.search(np.ascontiguousarray(x, dtype=np.float32), 1)
        return D.ravel(), I.ravel()
