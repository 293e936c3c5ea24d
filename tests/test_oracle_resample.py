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
