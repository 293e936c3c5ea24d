"""The resampling oracle's numpy restatement against the library calls the reference makes (torchaudio, CPU)."""
import numpy as np
import pytest

from oracle import resample_ref


@pytest.mark.parametrize("orig,new,channels", [(44100, 22050, 2), (48000, 22050, 1), (16000, 22050, 1), (32000, 22050, 2)])
def test_numpy_restatement_matches_torchaudio(orig, new, channels):
    rng = np.random.default_rng(orig + channels)
    L = orig // 2 + 37
    wave = (rng.integers(-20000, 20000, size=(channels, L)).astype(np.float32) / 32768.0)
    a = resample_ref.resample_torchaudio(wave, orig, new)
    b = resample_ref.resample_numpy(wave, orig, new)
    assert a.shape == b.shape == (1, int(np.ceil(new * L / orig)))
    assert np.abs(a - b).max() <= 2e-6


def test_kernel_known_answers():
    """44100 -> 22050 reduces to orig 2, new 1: width ceil(12 / 0.99) = 13, 28 taps, centre tap = 0.99 / 2."""
    k, width, o, n = resample_ref.sinc_kernel(44100, 22050)
    assert (width, o, n) == (13, 2, 1) and k.shape == (1, 28)
    assert abs(float(k[0, width]) - 0.495) < 1e-7
    k2, width2, o2, n2 = resample_ref.sinc_kernel(48000, 22050)
    assert (o2, n2) == (320, 147) and width2 == 14 and k2.shape == (147, 348)


@pytest.mark.parametrize("name", ["frontend_44100_stereo", "frontend_48000_mono", "frontend_16000_mono"])
def test_oracle_matches_the_reference_methods_goldens(golden_dir, name):
    """tests/golden/frontend_*.npz were produced by the reference's own convert_to_mono + resample
    (tests/golden/make_golden_frontend.py): both oracle flavours reproduce them."""
    import os

    g = np.load(os.path.join(golden_dir, name + ".npz"))
    wave = g["pcm"].astype(np.float32) / np.float32(32768.0)
    sr, common = int(g["source_rate"]), int(g["common_sr"])
    a = resample_ref.resample_torchaudio(wave, sr, common)
    b = resample_ref.resample_numpy(wave, sr, common)
    assert a.shape == g["out"].shape
    assert np.abs(a - g["out"]).max() <= 1e-6
    assert np.abs(b - g["out"]).max() <= 2e-6


@pytest.mark.parametrize("orig,new", [(44100, 22050), (48000, 22050), (16000, 22050), (32000, 22050), (8000, 22050)])
def test_library_filter_bank_equals_torchaudio(orig, new):
    """The filter bank the CUDA library builds on the host (at_resample_bank_host, no device needed) against
    torchaudio's own _get_sinc_resample_kernel: same shape, values within one float32 ulp of the centre tap."""
    import ctypes
    import math

    import torchaudio.functional.functional as AF
    from at_b200 import _lib

    lib = _lib.load()
    ph, taps, width = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    _lib.check(lib.at_resample_bank_host(orig, new, None, ctypes.byref(ph), ctypes.byref(taps), ctypes.byref(width)))
    bank = np.empty((ph.value, taps.value), dtype=np.float32)
    _lib.check(lib.at_resample_bank_host(orig, new, bank.ctypes.data_as(ctypes.c_void_p), None, None, None))
    kern, w = AF._get_sinc_resample_kernel(orig, new, math.gcd(orig, new))
    kern = kern.numpy()[:, 0, :]
    assert w == width.value and kern.shape == bank.shape
    assert np.abs(bank - kern).max() <= 6e-8
    ours, _, _, _ = resample_ref.sinc_kernel(orig, new)
    assert np.abs(ours - kern).max() <= 6e-8
