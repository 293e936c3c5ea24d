"""``nn.Module`` mirrors of the two torchaudio transforms the reference constructs
(processors/spectrogram_generator.py:28-34 of danavery/audio-tokens):

    self.spec_transformer = MelSpectrogram(sample_rate=..., n_mels=..., n_fft=..., hop_length=...).to(self.device)
    self.amplitude_to_db_transformer = AmplitudeToDB().to(self.device)
    mel_spec = self.spec_transformer(audio).squeeze(0)             # :124
    mel_spec_db = self.amplitude_to_db_transformer(mel_spec)       # :125

Same constructor keywords, same call shapes ((..., L) waveform -> (..., n_mels, T) power; (...) -> (...) dB), ``.to(device)``
like any module, backed by the fused sm_100a mel kernel in its power-output mode (at_mel_plan_set_output) and by
at_amplitude_to_db.  With ``at_b200.dropin.install(operators=True)`` they replace the torchaudio classes, so the
reference's own spectrogram_generator.py runs unchanged on the B200 kernels (level L-B).
Only torchaudio's defaults (the reference overrides nothing else) are implemented; anything else raises
NotImplementedError rather than computing something different.
"""
from __future__ import annotations

import math

import torch
from torch import nn

from . import _lib
from .mel import MelPlan


class MelSpectrogram(nn.Module):
    """torchaudio.transforms.MelSpectrogram with its defaults: win_length = n_fft, periodic Hann window, power 2, center +
    reflect pad, onesided, HTK mel scale, norm None, f_min 0, f_max sample_rate // 2."""

    def __init__(self, sample_rate: int = 16000, n_fft: int = 400, win_length=None, hop_length=None, f_min: float = 0.0,
                 f_max=None, pad: int = 0, n_mels: int = 128, window_fn=torch.hann_window, power: float = 2.0,
                 normalized: bool = False, wkwargs=None, center: bool = True, pad_mode: str = "reflect", onesided=None,
                 norm=None, mel_scale: str = "htk"):
        super().__init__()
        hop_length = hop_length if hop_length is not None else (win_length or n_fft) // 2
        unsupported = []
        if win_length not in (None, n_fft):
            unsupported.append("win_length != n_fft")
        if f_min != 0.0 or f_max not in (None, float(sample_rate // 2)):
            unsupported.append("f_min / f_max")
        if pad != 0 or window_fn is not torch.hann_window or wkwargs or power != 2.0 or normalized:
            unsupported.append("pad / window_fn / wkwargs / power / normalized")
        if not center or pad_mode != "reflect" or onesided not in (None, True) or norm is not None or mel_scale != "htk":
            unsupported.append("center / pad_mode / onesided / norm / mel_scale")
        if unsupported:
            raise NotImplementedError("at_b200 MelSpectrogram covers torchaudio's defaults only (" + ", ".join(unsupported) + ")")
        self.sample_rate, self.n_fft, self.hop_length, self.n_mels = int(sample_rate), int(n_fft), int(hop_length), int(n_mels)
        self.win_length = self.n_fft
        self._plan = None

    def _get_plan(self):
        if self._plan is None:
            plan = MelPlan(self.sample_rate, self.n_fft, self.hop_length, self.n_mels, normalize=False)
            _lib.check(plan.lib.at_mel_plan_set_output(plan.h, 1))   # AT_MEL_OUT_POWER
            self._plan = plan
        return self._plan

    def forward(self, waveform: torch.Tensor) -> torch.Tensor:
        """(..., L) -> (..., n_mels, T) mel power, T = 1 + L // hop_length."""
        if not waveform.is_cuda:
            raise RuntimeError("at_b200 MelSpectrogram needs a CUDA tensor (there is no CPU fallback); call .to('cuda') on the input")
        lead = waveform.shape[:-1]
        w = waveform.reshape(-1, waveform.shape[-1]).to(torch.float32).contiguous()
        spec, bad = self._get_plan().forward(w)          # (B, T, n_mels) frame-major
        if bool((bad == 2).any()):
            raise RuntimeError("Argument #4: Padding size should be less than the corresponding input dimension, "
                               f"but got: padding ({self.n_fft // 2}, {self.n_fft // 2}) at dimension 2 of input "
                               f"{[1, w.shape[0], w.shape[1]]}")
        # (n_mels, T) view of each frame-major tile: the layout torch's own matmul + transpose hands back (and np.save then
        # writes with fortran_order=True, like the reference's files)
        return spec.transpose(1, 2).reshape(*lead, self.n_mels, spec.shape[1])


class AmplitudeToDB(nn.Module):
    """torchaudio.transforms.AmplitudeToDB: multiplier * log10(clamp(x, amin)) - multiplier * db_multiplier, optional
    top_db clamp per (channel, freq, time) tile like torchaudio.functional.amplitude_to_DB."""

    def __init__(self, stype: str = "power", top_db=None):
        super().__init__()
        if top_db is not None and top_db < 0:
            raise ValueError("top_db must be positive value")
        self.stype = stype
        self.top_db = top_db
        self.multiplier = 10.0 if stype == "power" else 20.0
        self.amin = 1e-10
        self.ref_value = 1.0
        self.db_multiplier = math.log10(max(self.amin, self.ref_value))

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if not x.is_cuda:
            raise RuntimeError("at_b200 AmplitudeToDB needs a CUDA tensor (there is no CPU fallback)")
        src = x.to(torch.float32)
        # keep the memory order of the input (the mel tiles arrive as transposed views): work on the dense storage
        if src.is_contiguous():
            dense = src
        elif src.dim() >= 2 and src.transpose(-1, -2).is_contiguous():
            dense = src.transpose(-1, -2)
        else:
            dense = src.contiguous()
            src = dense
        out_dense = torch.empty_like(dense)
        lib = _lib.load()
        _lib.check(lib.at_amplitude_to_db(_lib.ptr(dense), dense.numel(), self.multiplier, self.amin, self.db_multiplier,
                                          _lib.ptr(out_dense), _lib.stream_ptr()))
        out = out_dense if dense is src else out_dense.transpose(-1, -2)
        if self.top_db is not None:
            # functional.amplitude_to_DB: clamp to (max over each (channel, freq, time) tile) - top_db
            shape = out.shape
            packed = out.reshape(-1, shape[-3], shape[-2], shape[-1]) if out.dim() > 2 else out.reshape(1, 1, *shape[-2:])
            floor = packed.amax(dim=(-3, -2, -1), keepdim=True) - self.top_db
            out = torch.max(packed, floor).reshape(shape)
        return out
