"""ResamplePlan: channel mean + torchaudio.transforms.Resample(orig, new) (defaults) on the device.

Mirrors SpectrogramGenerator.convert_to_mono + SpectrogramGenerator.resample
(processors/spectrogram_generator.py:109-121); one plan per (orig_freq, new_freq) pair instead of a new Resample module
(= a new filter bank) per clip.
"""
from __future__ import annotations

import ctypes

from . import _lib


class ResamplePlan:
    def __init__(self, orig_freq: int, new_freq: int):
        _lib.require_cuda()
        self.lib = _lib.load()
        self.orig_freq, self.new_freq = int(orig_freq), int(new_freq)
        h = ctypes.c_void_p()
        _lib.check(self.lib.at_resample_plan_create(self.orig_freq, self.new_freq, ctypes.byref(h)))
        self.h = h

    def __del__(self):
        try:
            if getattr(self, "h", None):
                self.lib.at_resample_plan_destroy(self.h)
                self.h = None
        except Exception:
            pass

    def out_len(self, n_in: int) -> int:
        return int(self.lib.at_resample_out_len(self.h, int(n_in)))

    def forward(self, wave, out=None):
        """wave (C, L) fp32 CUDA (torchaudio.load's layout) -> (1, ceil(new * L / orig)) mono waveform at new_freq."""
        import torch

        assert wave.is_cuda and wave.dtype == torch.float32 and wave.dim() == 2 and wave.is_contiguous()
        C, L = wave.shape
        n_out = self.out_len(L)
        if out is None:
            out = torch.empty((1, n_out), dtype=torch.float32, device=wave.device)
        assert out.is_cuda and out.dtype == torch.float32 and out.is_contiguous() and out.numel() >= n_out
        _lib.check(self.lib.at_resample_mono(self.h, _lib.ptr(wave), C, L, _lib.ptr(out), _lib.stream_ptr()))
        return out

    def forward_batch(self, wave, out=None):
        """wave (B, C, L) fp32 CUDA -> (B, ceil(new * L / orig)): B clips of the same length in one launch."""
        import torch

        assert wave.is_cuda and wave.dtype == torch.float32 and wave.dim() == 3 and wave.is_contiguous()
        B, C, L = wave.shape
        n_out = self.out_len(L)
        if out is None:
            out = torch.empty((B, n_out), dtype=torch.float32, device=wave.device)
        assert out.is_cuda and out.dtype == torch.float32 and out.is_contiguous() and out.numel() >= B * n_out
        _lib.check(self.lib.at_resample_mono_batch(self.h, _lib.ptr(wave), C, L, B, _lib.ptr(out), _lib.stream_ptr()))
        return out
