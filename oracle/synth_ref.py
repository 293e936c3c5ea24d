"""oracle/synth_ref.py -- TEST INFRASTRUCTURE ONLY.

CPU (numpy) twin of the device-side synthetic clip generator (audio-tokens_b200/csrc/at_synth.cu).
Integer-only arithmetic, so both produce bit-identical 16-bit-PCM-valued waveforms (stored as fp32 =
int16 / 32768, which is what torchaudio.load yields for the 16-bit FLAC files the reference reads,
processors/spectrogram_generator.py:99).  There is no network and no AudioSet audio in this image, so
every benchmark and parity test runs on these clips.

Clip ``i`` of a run with base seed ``S``: 3-8 sine partials (shared 4096-entry int16 sine table) with hashed
integer phase increments, amplitudes and per-partial piecewise-linear envelopes (8192-sample segments, some
silent), plus hashed white noise at a per-clip level; clipped to int16.  Never constant.
"""
from __future__ import annotations

import numpy as np

SINE_TABLE_BITS = 12
ENV_SEG_BITS = 13
_M32 = np.uint64(0xFFFFFFFF)


def sine_table() -> np.ndarray:
    """round(32767 sin(2 pi k / 4096)) as int16; the device generator is handed this same table."""
    k = np.arange(1 << SINE_TABLE_BITS, dtype=np.float64)
    return np.round(32767.0 * np.sin(2.0 * np.pi * k / (1 << SINE_TABLE_BITS))).astype(np.int16)


def hash32(x):
    """lowbias32 finaliser on uint32 (scalar or array)."""
    x = np.asarray(x, dtype=np.uint64) & _M32
    x ^= x >> np.uint64(16)
    x = (x * np.uint64(0x7FEB352D)) & _M32
    x ^= x >> np.uint64(15)
    x = (x * np.uint64(0x846CA68B)) & _M32
    x ^= x >> np.uint64(16)
    return x


def _h(x) -> int:
    return int(hash32(np.uint64(int(x) & 0xFFFFFFFF)))


def clip_params(seed: int, index: int) -> dict:
    h0 = _h((seed * 0x9E3779B1 + index) & 0xFFFFFFFF)
    h0 = _h(h0 ^ 0x85EBCA6B)
    P = 3 + _h(h0 + 1) % 6
    noise_amp = 16 + _h(h0 + 2) % 240
    gain = 96 + _h(h0 + 3) % 160
    partials = []
    for p in range(P):
        hp = _h(h0 + 16 + p)
        hq = _h(hp)
        inc = ((15600000 + ((hq >> 8) & 0xFFFFFF)) << (hp % 7)) & 0xFFFFFFFF
        phi0 = _h(hp + 0x1234567)
        amp = 512 + _h(hq + 7) % 3584
        partials.append(dict(hp=hp, inc=inc, phi0=phi0, amp=amp))
    return dict(h0=h0, noise_amp=noise_amp, gain=gain, partials=partials)


def make_clip_int16(seed: int, index: int, n_samples: int, table: np.ndarray | None = None) -> np.ndarray:
    if table is None:
        table = sine_table()
    tab = table.astype(np.int64)
    prm = clip_params(seed, index)
    n = np.arange(n_samples, dtype=np.uint64)
    total = np.zeros(n_samples, dtype=np.int64)
    seg = (n >> np.uint64(ENV_SEG_BITS)).astype(np.uint64)
    r = (n & np.uint64((1 << ENV_SEG_BITS) - 1)).astype(np.int64)
    for part in prm["partials"]:
        phase = (np.uint64(part["phi0"]) + n * np.uint64(part["inc"])) & _M32
        s = tab[(phase >> np.uint64(32 - SINE_TABLE_BITS)).astype(np.int64)]
        hp = np.uint64(part["hp"])

        def env_at(j):
            h = hash32((hp ^ ((j * np.uint64(0x9E3779B1)) & _M32)) + np.uint64(0x55))
            return np.maximum(0, (h % np.uint64(384)).astype(np.int64) - 128)

        e0 = env_at(seg)
        e1 = env_at(seg + np.uint64(1))
        env = (e0 * ((1 << ENV_SEG_BITS) - r) + e1 * r) >> ENV_SEG_BITS
        total += (((s * part["amp"]) >> 12) * env) >> 8
    total = (total * prm["gain"]) >> 9
    hn = hash32(np.uint64(prm["h0"]) ^ hash32(n + np.uint64(0x68E31DA4)))
    na = prm["noise_amp"]
    nz = (hn % np.uint64(2 * na + 1)).astype(np.int64) - na
    return np.clip(total + nz, -32768, 32767).astype(np.int16)


def make_clip(seed: int, index: int, n_samples: int, table=None) -> np.ndarray:
    """fp32 waveform in [-1, 1): int16 / 32768 (exact)."""
    return make_clip_int16(seed, index, n_samples, table).astype(np.float32) / np.float32(32768.0)


def make_clips(seed: int, first: int, count: int, n_samples: int) -> np.ndarray:
    table = sine_table()
    out = np.empty((count, n_samples), dtype=np.float32)
    for i in range(count):
        out[i] = make_clip(seed, first + i, n_samples, table)
    return out


_CLIB = None


def make_clips_c(seed: int, first: int, count: int, n_samples: int, pcm16: bool = False) -> np.ndarray:
    """Same clips through the C twin (oracle/synth_ref.c, OpenMP over clips): what bench.py's CPU arm uses to build its
    input in seconds.  Bit-identical to make_clips (tests/test_oracle_synth.py).  pcm16=True returns the int16 samples."""
    import ctypes
    import os

    global _CLIB
    if _CLIB is None:
        here = os.path.dirname(os.path.abspath(__file__))
        path = os.path.join(here, "_ref", "libsynth_ref.so")
        if not os.path.exists(path):
            import subprocess

            subprocess.check_call(["make", "-C", here], stdout=subprocess.DEVNULL)
        _CLIB = ctypes.CDLL(path)
        _CLIB.ref_synth_clips.argtypes = [ctypes.c_uint32, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64,
                                          ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
        _CLIB.ref_synth_clips.restype = None
    table = np.ascontiguousarray(sine_table())
    out = np.empty((count, n_samples), dtype=np.int16 if pcm16 else np.float32)
    ptr = out.ctypes.data_as(ctypes.c_void_p)
    _CLIB.ref_synth_clips(seed & 0xFFFFFFFF, first, count, n_samples, table.ctypes.data_as(ctypes.c_void_p),
                          None if pcm16 else ptr, ptr if pcm16 else None)
    return out
