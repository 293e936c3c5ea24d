"""Flat L2 index: device-level FlatL2 and the faiss.IndexFlatL2 look-alike used by SpecTokenizer
(processors/spec_tokenizer.py:123-127,77)."""
from __future__ import annotations

import ctypes

import numpy as np

from . import _lib


class FlatL2:
    """Device-level index over CUDA tensors (at_index_*)."""

    def __init__(self, d: int):
        _lib.require_cuda()
        self.lib = _lib.load()
        self.d = int(d)
        h = ctypes.c_void_p()
        _lib.check(self.lib.at_index_create(self.d, ctypes.byref(h)))
        self.h = h
        self.k = 0

    def __del__(self):
        try:
            if getattr(self, "h", None):
                self.lib.at_index_destroy(self.h)
                self.h = None
        except Exception:
            pass

    def set_tc_mode(self, mode: int):
        """0 auto (resident when the tiles fit), 1 always stream centroid tiles (at_index_set_tc_mode)."""
        _lib.check(self.lib.at_index_set_tc_mode(self.h, int(mode)))

    def tc_stats(self):
        """(rows decided by the fp32 re-check, rows scanned exactly) of the tcgen05 kernel so far (at_index_tc_stats)."""
        out = (ctypes.c_uint64 * 2)()
        _lib.check(self.lib.at_index_tc_stats(self.h, out))
        return int(out[0]), int(out[1])

    def set_centroids(self, c):
        import torch

        assert c.is_cuda and c.dtype == torch.float32 and c.dim() == 2 and c.shape[1] == self.d and c.is_contiguous()
        _lib.check(self.lib.at_index_set_centroids(self.h, _lib.ptr(c), c.shape[0], _lib.stream_ptr()))
        self.k = c.shape[0]

    def search(self, x, l2norm_rows: bool = False, algo: int = _lib.ALGO_AUTO, want_dist: bool = True,
               labels_dtype=None, labels=None, dist=None):
        """x (n, d) fp32 CUDA -> (labels, dist).  labels int32 by default (int64 on request)."""
        import torch

        assert x.is_cuda and x.dtype == torch.float32 and x.dim() == 2 and x.shape[1] == self.d and x.is_contiguous()
        n = x.shape[0]
        labels_dtype = labels_dtype or (labels.dtype if labels is not None else torch.int32)
        if labels is None:
            labels = torch.empty(n, dtype=labels_dtype, device=x.device)
        if want_dist and dist is None:
            dist = torch.empty(n, dtype=torch.float32, device=x.device)
        l32 = labels if labels_dtype == torch.int32 else None
        l64 = labels if labels_dtype == torch.int64 else None
        _lib.check(self.lib.at_index_search(self.h, _lib.ptr(x), n, int(l2norm_rows), algo, _lib.ptr(l32),
                                            _lib.ptr(l64), _lib.ptr(dist if want_dist else None),
                                            _lib.stream_ptr()))
        return labels, dist


    def search_trained_rows(self, trainer, x, l2norm_rows: bool = False, want_dist: bool = False, labels_dtype=None,
                            labels=None, dist=None):
        """search(x, 1) for the rows ``trainer`` (a LloydTrainer) was just trained on, re-using the operand image its Lloyd
        iterations built (at_index_search_trained_rows).  Same results as search(..., algo=ALGO_TENSOR)."""
        import torch

        assert x.is_cuda and x.dtype == torch.float32 and x.dim() == 2 and x.shape[1] == self.d and x.is_contiguous()
        n = x.shape[0]
        labels_dtype = labels_dtype or (labels.dtype if labels is not None else torch.int32)
        if labels is None:
            labels = torch.empty(n, dtype=labels_dtype, device=x.device)
        if want_dist and dist is None:
            dist = torch.empty(n, dtype=torch.float32, device=x.device)
        l32 = labels if labels_dtype == torch.int32 else None
        l64 = labels if labels_dtype == torch.int64 else None
        _lib.check(self.lib.at_index_search_trained_rows(self.h, trainer.h, _lib.ptr(x), n, int(l2norm_rows), _lib.ptr(l32),
                                                         _lib.ptr(l64), _lib.ptr(dist if want_dist else None),
                                                         _lib.stream_ptr()))
        return labels, dist


class IndexFlatL2:
    """faiss.IndexFlatL2(d) subset: add / reset / search(x, 1) / ntotal, numpy in, numpy out
    ((n, 1) float32 distances, (n, 1) int64 labels -- spec_tokenizer.py:77-78 squeezes axis 1)."""

    def __init__(self, d: int):
        self.d = int(d)
        self._ix = FlatL2(self.d)
        self._xb = np.zeros((0, self.d), dtype=np.float32)
        self.is_trained = True

    @property
    def ntotal(self) -> int:
        return self._xb.shape[0]

    def add(self, x):
        import torch

        x = np.ascontiguousarray(x, dtype=np.float32)
        assert x.ndim == 2 and x.shape[1] == self.d
        self._xb = np.concatenate([self._xb, x], axis=0)
        self._ix.set_centroids(torch.from_numpy(self._xb).cuda())

    def reset(self):
        self._xb = np.zeros((0, self.d), dtype=np.float32)
        self._ix.k = 0

    def search(self, x, k: int = 1, l2norm_rows: bool = False, chunk_rows: int = 1 << 22, want_dist: bool = True):
        """FAISS's search(x, k).  want_dist=False (not a FAISS argument; the stage classes discard D) returns (None, I) and
        lets the library skip the exact-distance pass -- and take the tensor path for rows wider than 64 values."""
        import torch

        if k != 1:
            raise NotImplementedError("only k=1 (the value the reference uses) is implemented")
        if self.ntotal == 0:
            raise RuntimeError("search on an empty index")
        on_device = torch.is_tensor(x)   # rows already in HBM (e.g. the conv-expanded batch): no host round trip
        if not on_device:
            x = np.ascontiguousarray(x, dtype=np.float32)
        assert x.ndim == 2 and x.shape[1] == self.d
        n = x.shape[0]
        D = np.empty((n, 1), dtype=np.float32) if want_dist else None
        I = np.empty((n, 1), dtype=np.int64)
        for a in range(0, n, chunk_rows):
            b = min(n, a + chunk_rows)
            xd = x[a:b].contiguous() if on_device else torch.from_numpy(x[a:b]).cuda()
            fused = l2norm_rows and self.d <= 128   # the fused normalisation covers the register-resident kernels
            if l2norm_rows and not fused:
                from . import row_l2norm

                xd = row_l2norm(xd)
            lab, dist = self._ix.search(xd, l2norm_rows=fused, labels_dtype=torch.int64, want_dist=want_dist)
            I[a:b, 0] = lab.cpu().numpy()
            if want_dist:
                D[a:b, 0] = dist.cpu().numpy()
        return D, I
