"""oracle/mel_ref.py -- TEST INFRASTRUCTURE ONLY.

CPU restatement of stage 1 of the audio-tokens hot path:

    SpectrogramGenerator.generate_mel_spectrogram   processors/spectrogram_generator.py:123-126
    SpectrogramGenerator.normalize_spectrogram      processors/spectrogram_generator.py:128-131
    SpectrogramGenerator.check_for_nan_inf          processors/spectrogram_generator.py:133-146

whose arithmetic lives in torchaudio (pinned 2.4.1 in environment.yml:265; 2.11.0 installed here, same
functional code path):

    MelSpectrogram(sample_rate, n_mels, n_fft, hop_length) with every other argument at its default
        -> Spectrogram: periodic Hann, center=True, pad_mode="reflect", power=2, onesided
           (torchaudio/transforms/_transforms.py:566-632, functional/functional.py:54-146)
        -> MelScale: HTK scale, norm=None, f_min=0, f_max=sr//2
           (torchaudio/transforms/_transforms.py:380-420, functional/functional.py:425-587)
    AmplitudeToDB() : 10*log10(clamp(x, 1e-10)), db_multiplier = log10(max(1e-10, 1.0)) = 0, no top_db
           (torchaudio/transforms/_transforms.py:324-347, functional/functional.py:356-404)

Two flavours are provided:

* ``mel_db_numpy``  -- an fp64 numpy restatement (independent arithmetic, used for tolerance analysis).
* ``mel_db_torchaudio`` -- the same library calls the reference class makes (used as the CPU baseline on the
  GPU box, where /root/reference does not exist, and to cross-check the numpy restatement).

PINNED: tests/golden/mel_*.npz were generated in the build container by importing the reference's own
``SpectrogramGenerator`` from /root/reference (tests/golden/make_golden.py); both flavours are checked
against them in tests/test_oracle_mel.py.
"""
from __future__ import annotations

import math

import numpy as np


# --------------------------------------------------------------------------- constants
def hann_periodic(n_fft: int, dtype=np.float64) -> np.ndarray:
    """torch.hann_window(n_fft, periodic=True): 0.5 - 0.5 cos(2 pi n / n_fft)."""
    n = np.arange(n_fft, dtype=np.float64)
    return (0.5 - 0.5 * np.cos(2.0 * np.pi * n / n_fft)).astype(dtype)


def melscale_fbanks_htk(n_freqs: int, f_min: float, f_max: float, n_mels: int, sample_rate: int,
                        dtype=np.float64) -> np.ndarray:
    """torchaudio.functional.melscale_fbanks(norm=None, mel_scale="htk") -> (n_freqs, n_mels).

    functional.py:518-587: all_freqs = linspace(0, sample_rate // 2, n_freqs);
    m_pts = linspace(hz2mel(f_min), hz2mel(f_max), n_mels + 2); f_pts = 700 (10^(m/2595) - 1);
    fb = max(0, min(down_slopes, up_slopes)).
    """
    all_freqs = np.linspace(0.0, float(sample_rate // 2), n_freqs)
    m_min = 2595.0 * math.log10(1.0 + f_min / 700.0)
    m_max = 2595.0 * math.log10(1.0 + f_max / 700.0)
    m_pts = np.linspace(m_min, m_max, n_mels + 2)
    f_pts = 700.0 * (10.0 ** (m_pts / 2595.0) - 1.0)
    f_diff = f_pts[1:] - f_pts[:-1]
    slopes = f_pts[None, :] - all_freqs[:, None]
    down = (-1.0 * slopes[:, :-2]) / f_diff[:-1]
    up = slopes[:, 2:] / f_diff[1:]
    fb = np.maximum(0.0, np.minimum(down, up))
    return fb.astype(dtype)


def num_frames(n_samples: int, hop: int) -> int:
    """center=True STFT: 1 + floor(L / hop)."""
    return 1 + n_samples // hop


# --------------------------------------------------------------------------- fp64 restatement
def power_spectrogram_numpy(wave: np.ndarray, n_fft: int, hop: int) -> np.ndarray:
    """|STFT|^2, shape (n_fft//2+1, T), fp64. Reflect pad n_fft//2, periodic Hann, win_length=n_fft."""
    wave = np.asarray(wave, dtype=np.float64).reshape(-1)
    pad = n_fft // 2
    if wave.shape[0] <= pad:
        # torch's reflect pad requires pad < input length
        raise RuntimeError(
            "Argument #4: Padding size should be less than the corresponding input dimension, "
            "but got: padding (%d, %d) at dimension 2 of input" % (pad, pad)
        )
    padded = np.pad(wave, (pad, pad), mode="reflect")
    T = num_frames(wave.shape[0], hop)
    idx = np.arange(n_fft)[None, :] + hop * np.arange(T)[:, None]
    frames = padded[idx] * hann_periodic(n_fft)[None, :]
    spec = np.fft.rfft(frames, axis=1)
    return (spec.real ** 2 + spec.imag ** 2).T


def mel_db_numpy(wave: np.ndarray, sample_rate: int, n_fft: int, hop: int, n_mels: int,
                 normalize: bool = False, fb: np.ndarray | None = None) -> np.ndarray:
    """(n_mels, T) fp64 dB mel spectrogram (optionally min-max normalised over the whole clip)."""
    power = power_spectrogram_numpy(wave, n_fft, hop)  # (F, T)
    if fb is None:
        fb = melscale_fbanks_htk(n_fft // 2 + 1, 0.0, float(sample_rate // 2), n_mels, sample_rate)
    mel = fb.T.astype(np.float64) @ power  # (n_mels, T)
    db = 10.0 * np.log10(np.maximum(mel, 1e-10))
    if normalize:
        db = normalize_numpy(db)
    return db


def normalize_numpy(spec: np.ndarray) -> np.ndarray:
    """(s - min) / (max - min) over the whole tile; max == min gives NaN (0/0) like the reference."""
    lo, hi = spec.min(), spec.max()
    with np.errstate(invalid="ignore", divide="ignore"):
        return (spec - lo) / (hi - lo)


def is_bad(spec: np.ndarray) -> bool:
    """check_for_nan_inf: any NaN or Inf -> the clip is dropped."""
    return bool(np.isnan(spec).any() or np.isinf(spec).any())


# --------------------------------------------------------------------------- library-call flavour
class TorchaudioMel:
    """The exact library calls of the reference class (spectrogram_generator.py:28-34,123-131), CPU."""

    def __init__(self, sample_rate: int, n_fft: int, hop: int, n_mels: int, normalize: bool):
        import torch
        from torchaudio.transforms import AmplitudeToDB, MelSpectrogram

        self.torch = torch
        self.spec = MelSpectrogram(sample_rate=sample_rate, n_mels=n_mels, n_fft=n_fft, hop_length=hop)
        self.to_db = AmplitudeToDB()
        self.normalize = normalize

    def __call__(self, wave):
        torch = self.torch
        audio = torch.as_tensor(wave, dtype=torch.float32).reshape(1, -1)
        s = self.to_db(self.spec(audio).squeeze(0))
        if self.normalize:
            s = (s - torch.min(s)) / (torch.max(s) - torch.min(s))
        return s

    def is_bad(self, s) -> bool:
        torch = self.torch
        return bool(torch.isnan(s).any() or torch.isinf(s).any())


def mel_db_torchaudio(wave, sample_rate, n_fft, hop, n_mels, normalize=False) -> np.ndarray:
    return TorchaudioMel(sample_rate, n_fft, hop, n_mels, normalize)(wave).numpy()


def normalize_rows(v: np.ndarray) -> np.ndarray:
    """ClusterCreator.normalize_vectors / SpecTokenizer.normalize_vectors
    (cluster_creator.py:64-66, spec_tokenizer.py:106-109): v / (||v||_2 + 1e-10), fp32 in, fp32 out."""
    norms = np.linalg.norm(v, axis=1, keepdims=True)
    return v / (norms + 1e-10)
