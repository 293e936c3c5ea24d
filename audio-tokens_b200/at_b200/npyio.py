"""Batched reader for the spectrogram files the stages exchange (float32 (n_mels, T), fortran_order: physically
frame-major [T][n_mels]).

The reference reads a batch with ``[np.load(f).T for f in files]`` + ``np.concatenate`` on one thread
(processors/cluster_creator.py:88-102, processors/spec_tokenizer.py:67-73): 10,000 small reads back to back.  Here the
reads of a batch run on a small thread pool (file reads release the GIL) and the frames land directly in one
preallocated (N, n_mels) float32 matrix, in file order.
"""
from __future__ import annotations

from concurrent.futures import ThreadPoolExecutor

import numpy as np


def _load_frames(path):
    a = np.load(path)
    return a.T   # (T, n_mels): a C-contiguous view for the files the spectrogram stage writes


def load_spec_batch(files, threads: int = 8):
    """files: sequence of paths -> (frames (sum T, n_mels) float32 C-contiguous, lengths [T_i]) in the order given.
    Equal to ``np.concatenate([np.load(f).T for f in files], axis=0).astype(np.float32)``."""
    files = list(files)
    if not files:
        return np.zeros((0, 0), dtype=np.float32), []
    if threads <= 1 or len(files) == 1:
        parts = [_load_frames(f) for f in files]
    else:
        with ThreadPoolExecutor(max_workers=threads) as pool:
            parts = list(pool.map(_load_frames, files))
    lengths = [int(p.shape[0]) for p in parts]
    width = {int(p.shape[1]) for p in parts}
    if len(width) != 1:
        # same failure as np.concatenate in the reference
        raise ValueError("all the input array dimensions except for the concatenation axis must match exactly")
    out = np.empty((sum(lengths), width.pop()), dtype=np.float32)
    pos = 0
    for p, n in zip(parts, lengths):
        out[pos:pos + n] = p
        pos += n
    return out, lengths
