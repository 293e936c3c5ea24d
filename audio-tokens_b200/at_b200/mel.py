"""MelPlan: the device-side MelSpectrogram + AmplitudeToDB (+ min-max) of the reference's stage 1.

Mirrors processors/spectrogram_generator.py:28-34,123-131 of danavery/audio-tokens, batched over clips.
"""
from __future__ import annotations

import ctypes

import numpy as np

from . import _lib


class MelPlan:
    """MelSpectrogram(sample_rate, n_mels, n_fft, hop_length) + AmplitudeToDB() [+ normalize_spectrogram].

    ``torch_constants=True`` uploads the very window / filterbank tensors torch and torchaudio build
    (torch.hann_window, torchaudio.functional.melscale_fbanks) so the constants are bit-identical to the
    reference's; otherwise the library's own double-precision evaluation is used.
    """

    def __init__(self, sample_rate: int, n_fft: int, hop_length: int, n_mels: int, normalize: bool,
                 torch_constants: bool = True):
        _lib.require_cuda()
        self.lib = _lib.load()
        self.sample_rate, self.n_fft, self.hop, self.n_mels = sample_rate, n_fft, hop_length, n_mels
        self.normalize = bool(normalize)
        h = ctypes.c_void_p()
        _lib.check(self.lib.at_mel_plan_create(sample_rate, n_fft, hop_length, n_mels, int(self.normalize),
                                               ctypes.byref(h)))
        self.h = h
        if torch_constants:
            self._upload_torch_constants()

    def _upload_torch_constants(self):
        import torch

        try:
            import torchaudio.functional as AF
        except Exception:  # torchaudio not importable: keep the built-in constants
            return
        win = torch.hann_window(self.n_fft, periodic=True, dtype=torch.float32).numpy()
        import warnings

        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            fb = AF.melscale_fbanks(self.n_fft // 2 + 1, 0.0, float(self.sample_rate // 2), self.n_mels,
                                    self.sample_rate, None, "htk").numpy()
        win = np.ascontiguousarray(win, dtype=np.float32)
        fb = np.ascontiguousarray(fb, dtype=np.float32)
        _lib.check(self.lib.at_mel_plan_set_constants_host(self.h, _lib.ptr(win), _lib.ptr(fb)))

    def __del__(self):
        try:
            if getattr(self, "h", None):
                self.lib.at_mel_plan_destroy(self.h)
                self.h = None
        except Exception:
            pass

    def num_frames(self, n_samples: int) -> int:
        return 1 + n_samples // self.hop

    def set_absmax_out(self, t):
        """Attach (or detach with None) a CUDA float32[1] tensor that every forward writing the L2-normalised copy raises to
        the largest |element| written (at_mel_plan_set_absmax_out); the caller zeroes it."""
        _lib.check(self.lib.at_mel_plan_set_absmax_out(self.h, _lib.ptr(t)))

    def work_groups(self) -> int:
        """Independent clip streams of one launch (at_mel_work_groups): size streaming chunks as a multiple of it."""
        return int(self.lib.at_mel_work_groups(self.h))

    def forward(self, wave, out=None, out_l2=None, want_l2: bool = False):
        """Uniform batch: wave (B, L) fp32 CUDA -> (spec [B, T, n_mels] frame-major, bad_flags [B] int32[, l2])."""
        import torch

        assert wave.is_cuda and wave.dtype == torch.float32 and wave.dim() == 2 and wave.is_contiguous()
        B, L = wave.shape
        T = self.num_frames(L)
        if out is None:
            out = torch.empty((B, T, self.n_mels), dtype=torch.float32, device=wave.device)
        if want_l2 and out_l2 is None:
            out_l2 = torch.empty((B, T, self.n_mels), dtype=torch.float32, device=wave.device)
        bad = torch.zeros(B, dtype=torch.int32, device=wave.device)
        _lib.check(self.lib.at_mel_forward(self.h, _lib.ptr(wave), None, None, L, B, _lib.ptr(out),
                                           _lib.ptr(out_l2), _lib.ptr(bad), _lib.stream_ptr()))
        return (out, bad, out_l2) if (want_l2 or out_l2 is not None) else (out, bad)

    def forward_ragged(self, waves, want_l2: bool = False):
        """List of 1-D fp32 tensors (any device) -> (spec [sum T, n_mels], frame_offsets (B+1,) int64 host,
        bad_flags [B] int32[, l2])."""
        import torch

        lens = [int(w.numel()) for w in waves]
        B = len(lens)
        so = np.zeros(B + 1, dtype=np.int64)
        so[1:] = np.cumsum(lens)
        fo = np.zeros(B + 1, dtype=np.int64)
        fo[1:] = np.cumsum([self.num_frames(n) for n in lens])
        flat = torch.empty(int(so[-1]), dtype=torch.float32, device="cuda")
        for w, a, b in zip(waves, so[:-1], so[1:]):
            flat[a:b].copy_(w.reshape(-1), non_blocking=True)
        so_d = torch.from_numpy(so).cuda()
        fo_d = torch.from_numpy(fo).cuda()
        out = torch.empty((int(fo[-1]), self.n_mels), dtype=torch.float32, device="cuda")
        out_l2 = torch.empty_like(out) if want_l2 else None
        bad = torch.zeros(B, dtype=torch.int32, device="cuda")
        _lib.check(self.lib.at_mel_forward(self.h, _lib.ptr(flat), _lib.ptr(so_d), _lib.ptr(fo_d), 0, B,
                                           _lib.ptr(out), _lib.ptr(out_l2), _lib.ptr(bad), _lib.stream_ptr()))
        return (out, fo, bad, out_l2) if want_l2 else (out, fo, bad)

    def forward_host(self, wave: np.ndarray):
        """HOST buffers end to end (at_mel_forward_host): wave (B, L) fp32 numpy / pinned tensor ->
        (spec (B, T, n_mels) numpy, bad (B,) int32 numpy)."""
        arr = wave
        if hasattr(wave, "numpy"):
            arr = wave.numpy()
        assert arr.dtype == np.float32 and arr.ndim == 2 and arr.flags.c_contiguous
        B, L = arr.shape
        T = self.num_frames(L)
        out = np.empty((B, T, self.n_mels), dtype=np.float32)
        bad = np.zeros(B, dtype=np.int32)
        _lib.check(self.lib.at_mel_forward_host(self.h, _lib.ptr(arr), L, B, _lib.ptr(out), _lib.ptr(bad)))
        return out, bad
