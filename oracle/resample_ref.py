"""Oracle (TEST INFRASTRUCTURE ONLY) for the step before the path: channel mean + sample-rate conversion.

Restates SpectrogramGenerator.convert_to_mono + SpectrogramGenerator.resample
(/root/reference/processors/spectrogram_generator.py:109-121), i.e. torch.mean(dim=0, keepdim=True) followed by
torchaudio.transforms.Resample(sr, common_sr) with its defaults (torchaudio/functional/functional.py:
_get_sinc_resample_kernel, _apply_sinc_resample_kernel: sinc_interp_hann, lowpass_filter_width 6, rolloff 0.99).

Two flavours: ``resample_torchaudio`` makes the very library calls the reference makes (CPU); ``resample_numpy`` is an
fp64 restatement of the published arithmetic.  tests/test_oracle_resample.py pins the second against the first.
"""
import math

import numpy as np


def resample_torchaudio(wave: np.ndarray, orig_freq: int, new_freq: int) -> np.ndarray:
    """wave (C, L) float32 -> (1, L') float32 through torch.mean + torchaudio.transforms.Resample on the CPU."""
    import torch
    from torchaudio.transforms import Resample

    w = torch.from_numpy(np.ascontiguousarray(wave, dtype=np.float32))
    if w.shape[0] > 1:
        w = torch.mean(w, dim=0, keepdim=True)
    if orig_freq != new_freq:
        w = Resample(orig_freq, new_freq)(w)
    return w.numpy()


def sinc_kernel(orig_freq: int, new_freq: int):
    """(kernel float32 [new][2 width + orig], width, orig, new) exactly as _get_sinc_resample_kernel builds it."""
    g = math.gcd(int(orig_freq), int(new_freq))
    o, n = int(orig_freq) // g, int(new_freq) // g
    base = min(o, n) * 0.99
    width = math.ceil(6 * o / base)
    idx = np.arange(-width, width + o, dtype=np.float64)[None, :] / o
    phase = (np.arange(0, -n, -1).astype(np.float32) / np.float32(n)).astype(np.float64)[:, None]   # int64 / int -> float32
    t = (phase + idx) * base
    t = np.clip(t, -6, 6)
    window = np.cos(t * math.pi / 6 / 2) ** 2
    t = t * math.pi
    scale = base / o
    with np.errstate(invalid="ignore", divide="ignore"):
        k = np.where(t == 0, 1.0, np.sin(t) / t)
    k = k * (window * scale)
    return k.astype(np.float32), width, o, n


def resample_numpy(wave: np.ndarray, orig_freq: int, new_freq: int) -> np.ndarray:
    """fp64 restatement: y[j n + p] = sum_k kern[p][k] xpad[j o + k]."""
    w = np.asarray(wave, dtype=np.float32)
    if w.shape[0] > 1:
        w = (w.sum(axis=0, dtype=np.float32) / np.float32(w.shape[0]))[None, :]
    if orig_freq == new_freq:
        return w
    kern, width, o, n = sinc_kernel(orig_freq, new_freq)
    L = w.shape[1]
    xpad = np.concatenate([np.zeros(width), w[0].astype(np.float64), np.zeros(width + o)])
    taps = kern.shape[1]
    nj = (len(xpad) - taps) // o + 1
    frames = np.lib.stride_tricks.sliding_window_view(xpad, taps)[::o][:nj]      # (nj, taps)
    y = frames @ kern.astype(np.float64).T                                        # (nj, n)
    target = int(math.ceil(n * L / o))
    return y.reshape(-1)[:target].astype(np.float32)[None, :]
