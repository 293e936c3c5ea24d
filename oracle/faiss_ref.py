"""oracle/faiss_ref.py -- TEST INFRASTRUCTURE ONLY.

CPU restatement of the FAISS 1.8.0 subset that audio-tokens uses
(reference call sites: processors/cluster_creator.py:26,42-58; processors/spec_tokenizer.py:77,123-127;
tools/manual_tester.py:78,86-87).  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import this module.  The product
(``audio-tokens_b200/``) never does.

FAISS (conda ``faiss-gpu=1.8.0``, ``libfaiss=1.8.0``, MKL 2023.0; environment.yml:69,137,195) is a
third-party dependency that is absent from /root/reference and from this image, so the algorithm is
restated from the published upstream sources (facebookresearch/faiss tag v1.8.0):

* ``faiss/python/extra_wrappers.py``  class Kmeans  (kwargs -> ClusteringParameters, train(), .centroids, .obj)
* ``faiss/Clustering.{h,cpp}``        defaults, train_encoded, subsample_training_set, compute_centroids,
                                       split_clusters, imbalance_factor
* ``faiss/utils/random.cpp``          RandomGenerator = std::mt19937, rand_perm
* ``faiss/utils/distances.cpp``       knn_L2sqr -> exhaustive_L2sqr_blas (4096 x 1024 blocks, sgemm,
                                       dis = |x|^2 + |y|^2 - 2 ip, clamp at 0, strict '<' top-1)

PARITY UNPINNED against real FAISS outputs: there is no FAISS binary here and the reference holds no
golden vectors or tests for this path.  What *is* pinned: the mt19937 streams,
rand_perm and rand_float against known answers from g++ 13.3 ``std::mt19937`` (tests/test_oracle_faiss.py),
and every numeric routine against an fp64 brute-force evaluation.

Not reproducible bit-for-bit versus real FAISS (documented, and the reason the parity gates carry a
near-tie carve-out): MKL sgemm summation order and the AVX2 per-lane top-1 of
``exhaustive_L2sqr_fused_cmax``.
"""
from __future__ import annotations

import ctypes
import os
import sys
import time

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def _lib():
    """The C restatement (oracle/faiss_ref.c), built by ``make -C oracle`` / ``__graft_entry__.build()``."""
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "_ref", "libfaiss_ref.so")
        if not os.path.exists(path):
            import subprocess

            subprocess.check_call(["make", "-C", _HERE], stdout=subprocess.DEVNULL)
        lib = ctypes.CDLL(path)
        i64, f32p, i64p = ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p
        lib.ref_mt19937_words.argtypes = [i64, ctypes.c_int, ctypes.c_void_p]
        lib.ref_rand_floats.argtypes = [i64, ctypes.c_int, ctypes.c_void_p]
        lib.ref_rand_perm.argtypes = [ctypes.c_void_p, i64, i64]
        lib.ref_assign_l2.argtypes = [f32p, i64, f32p, i64, i64, i64p, f32p, f32p]
        lib.ref_assign_l2_f64.argtypes = [f32p, i64, f32p, i64, i64, i64p, ctypes.c_void_p, ctypes.c_void_p]
        lib.ref_compute_centroids.argtypes = [i64, i64, i64, f32p, i64p, f32p, f32p]
        lib.ref_l2_block_argmin.argtypes = [f32p, i64, i64, i64, f32p, f32p, i64, f32p, i64p]
        lib.ref_split_clusters.argtypes = [i64, i64, i64, f32p, f32p]
        lib.ref_split_clusters.restype = ctypes.c_int
        _LIB = lib
    return _LIB


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


# --------------------------------------------------------------------------- RNG (faiss/utils/random.cpp)
def mt19937_words(seed: int, n: int) -> np.ndarray:
    out = np.empty(n, dtype=np.uint32)
    _lib().ref_mt19937_words(seed, n, _p(out))
    return out


def rand_floats(seed: int, n: int) -> np.ndarray:
    out = np.empty(n, dtype=np.float32)
    _lib().ref_rand_floats(seed, n, _p(out))
    return out


def rand_perm(n: int, seed: int) -> np.ndarray:
    """faiss::rand_perm(perm, n, seed): forward Fisher-Yates driven by std::mt19937(seed)."""
    perm = np.empty(n, dtype=np.int32)
    _lib().ref_rand_perm(_p(perm), n, seed)
    return perm


# --------------------------------------------------------------------------- distances
def assign_l2_scalar(x, c):
    """Scalar fp32 FAISS formula, lowest index wins ties. Returns labels(int64), best, second."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    c = np.ascontiguousarray(c, dtype=np.float32)
    n, d = x.shape
    k = c.shape[0]
    labels = np.empty(n, dtype=np.int64)
    d1 = np.empty(n, dtype=np.float32)
    d2 = np.empty(n, dtype=np.float32)
    _lib().ref_assign_l2(_p(x), n, _p(c), k, d, _p(labels), _p(d1), _p(d2))
    return labels, d1, d2


def assign_l2_f64(x, c):
    """fp64 sum (x-c)^2 argmin with the top-2 distances (truth for near-tie analysis)."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    c = np.ascontiguousarray(c, dtype=np.float32)
    n, d = x.shape
    k = c.shape[0]
    labels = np.empty(n, dtype=np.int64)
    d1 = np.empty(n, dtype=np.float64)
    d2 = np.empty(n, dtype=np.float64)
    _lib().ref_assign_l2_f64(_p(x), n, _p(c), k, d, _p(labels), _p(d1), _p(d2))
    return labels, d1, d2


def knn_l2sqr_blas(x, c, bs_x=4096, bs_y=1024):
    """exhaustive_L2sqr_blas restated with an sgemm per (4096 x 1024) block (faiss/utils/distances.cpp).

    Uses torch (MKL sgemm, all host threads) when importable, else numpy.  This is the flavour timed as the
    CPU baseline because FAISS's own cost is one sgemm per block plus the top-1 scan.
    """
    x = np.ascontiguousarray(x, dtype=np.float32)
    c = np.ascontiguousarray(c, dtype=np.float32)
    n = x.shape[0]
    k = c.shape[0]
    try:
        import torch

        xt, ct = torch.from_numpy(x), torch.from_numpy(c)
        x_norms = (xt * xt).sum(1).numpy()
        c_norms = (ct * ct).sum(1).numpy()
        best = np.full(n, np.inf, dtype=np.float32)
        labels = np.zeros(n, dtype=np.int64)
        lib = _lib()
        ip = torch.empty((min(bs_x, n), min(bs_y, k)), dtype=torch.float32)
        for i0 in range(0, n, bs_x):
            i1 = min(n, i0 + bs_x)
            for j0 in range(0, k, bs_y):
                j1 = min(k, j0 + bs_y)
                blk = ip[: i1 - i0, : j1 - j0]
                torch.mm(xt[i0:i1], ct[j0:j1].T, out=blk)   # the sgemm of the block (MKL, all host threads)
                # fused epilogue in C (OpenMP): distances, clamp, strict '<' top-1 -- what FAISS does after its sgemm
                lib.ref_l2_block_argmin(blk.data_ptr(), i1 - i0, j1 - j0, ip.stride(0), _p(x_norms[i0:i1]),
                                        _p(c_norms[j0:j1]), j0, _p(best[i0:i1]), _p(labels[i0:i1]))
        return best, labels
    except ImportError:  # pragma: no cover
        x_norms = (x * x).sum(1)
        c_norms = (c * c).sum(1)
        best = np.full(n, np.inf, dtype=np.float32)
        labels = np.zeros(n, dtype=np.int64)
        for i0 in range(0, n, bs_x):
            i1 = min(n, i0 + bs_x)
            for j0 in range(0, k, bs_y):
                j1 = min(k, j0 + bs_y)
                dis = x_norms[i0:i1, None] + c_norms[None, j0:j1] - 2 * (x[i0:i1] @ c[j0:j1].T)
                np.maximum(dis, 0, out=dis)
                bidx = dis.argmin(1)
                bmin = dis[np.arange(i1 - i0), bidx]
                upd = bmin < best[i0:i1]
                best[i0:i1] = np.where(upd, bmin, best[i0:i1])
                labels[i0:i1] = np.where(upd, bidx + j0, labels[i0:i1])
        return best, labels


# --------------------------------------------------------------------------- Clustering pieces
def compute_centroids(x, assign, k):
    """faiss::compute_centroids: in-order fp32 sums, hassign as float, empty clusters stay zero."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    assign = np.ascontiguousarray(assign, dtype=np.int64)
    n, d = x.shape
    hassign = np.empty(k, dtype=np.float32)
    centroids = np.empty((k, d), dtype=np.float32)
    _lib().ref_compute_centroids(d, k, n, _p(x), _p(assign), _p(hassign), _p(centroids))
    return centroids, hassign


def split_clusters(centroids, hassign, n):
    """faiss::split_clusters (in place on copies). Returns centroids, hassign, nsplit."""
    centroids = np.array(centroids, dtype=np.float32, order="C")
    hassign = np.array(hassign, dtype=np.float32)
    k, d = centroids.shape
    nsplit = _lib().ref_split_clusters(d, k, n, _p(hassign), _p(centroids))
    return centroids, hassign, nsplit


def imbalance_factor(hist) -> float:
    """faiss::imbalance_factor(k, hist): sum(hist^2) * k / tot^2."""
    hist = np.asarray(hist, dtype=np.float64)
    tot = hist.sum()
    return float((hist * hist).sum() * len(hist) / (tot * tot))


def lloyd_step(x, centroids, exact: bool = False):
    """One FAISS iteration from given centroids: search, obj, compute_centroids, split_clusters.

    Returns dict(centroids, hassign_before_split, labels, dist, obj, nsplit).
    """
    k = centroids.shape[0]
    if exact:
        labels, dist, _ = assign_l2_scalar(x, centroids)
    else:
        dist, labels = knn_l2sqr_blas(x, centroids)
    obj = np.float32(0)
    # FAISS accumulates the objective in a float, in point order
    obj = float(np.cumsum(dist, dtype=np.float32)[-1]) if len(dist) else 0.0
    new_c, hassign = compute_centroids(x, labels, k)
    counts = hassign.copy()
    new_c, hassign, nsplit = split_clusters(new_c, hassign, x.shape[0])
    return dict(centroids=new_c, counts=counts, labels=labels, dist=dist, obj=obj, nsplit=nsplit)


# --------------------------------------------------------------------------- python API subset
def get_num_gpus() -> int:
    return 0


class IndexFlatL2:
    """faiss.IndexFlatL2 subset: add / reset / search(x, 1) / ntotal."""

    def __init__(self, d: int):
        self.d = int(d)
        self.xb = np.zeros((0, self.d), dtype=np.float32)
        self.is_trained = True

    @property
    def ntotal(self) -> int:
        return self.xb.shape[0]

    def add(self, x):
        x = np.ascontiguousarray(x, dtype=np.float32)
        assert x.ndim == 2 and x.shape[1] == self.d
        self.xb = np.concatenate([self.xb, x], axis=0)

    def reset(self):
        self.xb = np.zeros((0, self.d), dtype=np.float32)

    def search(self, x, k: int):
        if k != 1:
            raise NotImplementedError("oracle restates k=1 only (the only value the reference uses)")
        x = np.ascontiguousarray(x, dtype=np.float32)
        assert x.ndim == 2 and x.shape[1] == self.d
        if x.shape[0] < 20:  # distance_compute_blas_threshold: FAISS uses direct sum (x-y)^2 below it
            labels, d1, _ = assign_l2_f64(x, self.xb)
            return d1.astype(np.float32)[:, None], labels[:, None]
        dist, labels = knn_l2sqr_blas(x, self.xb)
        return dist[:, None], labels[:, None]


class ClusteringParameters:
    """faiss/Clustering.h defaults."""

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


class Kmeans:
    """faiss.Kmeans subset (faiss/python/extra_wrappers.py).  kwargs are copied onto ClusteringParameters;
    unknown names raise AttributeError like FAISS does."""

    def __init__(self, d, k, **kwargs):
        self.d = int(d)
        self.k = int(k)
        self.gpu = False
        self.cp = ClusteringParameters()
        for key, v in kwargs.items():
            if key == "gpu":
                if v is True or v == -1:
                    v = get_num_gpus()
                self.gpu = v
            else:
                getattr(self.cp, key)
                setattr(self.cp, key, v)
        self.centroids = None
        self.obj = None
        self.iteration_stats = None
        self.index = None
        self.exact_search = False  # oracle-only switch: scalar fp32 search instead of blocked sgemm
        self.perm_cache = None     # oracle-only: {(n, seed): rand_perm(n, seed)} drawn ahead of a timed train() call

    def train(self, x, weights=None, init_centroids=None):
        assert weights is None, "weights are not used by the reference"
        x = np.ascontiguousarray(x, dtype=np.float32)
        n, d = x.shape
        assert d == self.d
        cp = self.cp
        k = self.k
        if n < k:
            raise RuntimeError(
                "Error: 'nx >= k' failed: Number of training points (%d) should be at least as large as "
                "number of clusters (%d)" % (n, k)
            )
        if not np.isfinite(x).all():
            raise RuntimeError("Error: 'std::isfinite(x[i])' failed: input contains NaN's or Inf's")
        centroids_in = None
        if init_centroids is not None:
            nc, d2 = init_centroids.shape
            assert d2 == d
            centroids_in = np.ascontiguousarray(init_centroids, dtype=np.float32)

        # subsample_training_set
        if n > k * cp.max_points_per_centroid:
            perm = rand_perm(n, cp.seed)
            nx = k * cp.max_points_per_centroid
            if cp.verbose:
                print("Sampling a subset of %d / %d for training" % (nx, n))
            x = np.ascontiguousarray(x[perm[:nx]])
            n = nx
        elif n < k * cp.min_points_per_centroid:
            print(
                "WARNING clustering %d points to %d centroids: please provide at least %d training points"
                % (n, k, k * cp.min_points_per_centroid),
                file=sys.stderr,
            )

        self.index = IndexFlatL2(d)
        stats = []
        best_obj = None
        best_centroids = None
        for redo in range(cp.nredo):
            n_input = 0 if centroids_in is None else centroids_in.shape[0]
            centroids = np.zeros((k, d), dtype=np.float32)
            if n_input:
                centroids[:n_input] = centroids_in[:n_input]
            pseed = cp.seed + 1 + redo * 15486557
            perm = self.perm_cache.get((n, pseed)) if self.perm_cache else None
            if perm is None:
                perm = rand_perm(n, pseed)
            if n_input < k:
                centroids[n_input:] = x[perm[n_input:k]]
            if n == k:
                # train_encoded's corner case: the training set IS the centroids, in input order (init_centroids and the
                # permutation are ignored), and one all-zero iteration stat is pushed
                best_centroids = x.copy()
                stats.append(dict(obj=0.0, time=0.0, time_search=0.0, imbalance_factor=0.0, nsplit=0))
                break
            t0 = time.time()
            for it in range(cp.niter):
                step = lloyd_step(x, centroids, exact=self.exact_search)
                centroids = step["centroids"]
                stats.append(
                    dict(
                        obj=step["obj"],
                        time=time.time() - t0,
                        time_search=0.0,
                        imbalance_factor=imbalance_factor(step["counts"]),
                        nsplit=step["nsplit"],
                    )
                )
                if cp.verbose:
                    print(
                        "  Iteration %d (%.2f s) objective=%g imbalance=%.3f nsplit=%d"
                        % (it, stats[-1]["time"], step["obj"], stats[-1]["imbalance_factor"], step["nsplit"])
                    )
            if best_obj is None or (stats and stats[-1]["obj"] < best_obj):
                best_obj = stats[-1]["obj"] if stats else 0.0
                best_centroids = centroids
        self.centroids = best_centroids
        self.index.reset()
        self.index.add(self.centroids)
        self.iteration_stats = stats
        self.obj = np.array([s["obj"] for s in stats])
        return self.obj[-1] if self.obj.size > 0 else 0.0

    def assign(self, x):
        D, I = self.index.search(np.ascontiguousarray(x, dtype=np.float32), 1)
        return D.ravel(), I.ravel()
