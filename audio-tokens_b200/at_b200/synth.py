"""Synthetic 16-bit-PCM clips generated on the device (at_synth_clips)."""
from __future__ import annotations

import numpy as np

from . import _lib


def sine_table() -> np.ndarray:
    """The shared 4096-entry int16 sine table: round(32767 sin(2 pi k / 4096))."""
    k = np.arange(4096, dtype=np.float64)
    return np.round(32767.0 * np.sin(2.0 * np.pi * k / 4096.0)).astype(np.int16)


_TABLE_DEV = {}


def synth_clips(seed: int, first: int, count: int, n_samples: int, out=None):
    """(count, n_samples) fp32 CUDA tensor of clips first..first+count-1 of the run with base seed `seed`."""
    import torch

    _lib.require_cuda()
    dev = torch.cuda.current_device()
    if dev not in _TABLE_DEV:
        _TABLE_DEV[dev] = torch.from_numpy(sine_table()).cuda()
    if out is None:
        out = torch.empty((count, n_samples), dtype=torch.float32, device="cuda")
    assert out.is_cuda and out.is_contiguous() and out.numel() == count * n_samples
    _lib.check(_lib.load().at_synth_clips(seed & 0xFFFFFFFF, first, count, n_samples, _lib.ptr(_TABLE_DEV[dev]),
                                          _lib.ptr(out), _lib.stream_ptr()))
    return out
