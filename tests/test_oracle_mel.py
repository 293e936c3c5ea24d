"""Pins oracle/mel_ref.py against golden outputs of the reference's own SpectrogramGenerator
(tests/golden/make_golden.py ran processors/spectrogram_generator.py from /root/reference)."""
import glob
import os

import numpy as np
import pytest

from oracle import mel_ref, synth_ref


def _cases(golden_dir):
    return sorted(glob.glob(os.path.join(golden_dir, "mel_*.npz")))


def test_goldens_exist(golden_dir):
    assert len(_cases(golden_dir)) >= 3


def test_synth_generator_reproduces_golden_pcm(golden_dir):
    for path in _cases(golden_dir):
        g = np.load(path)
        for j, (idx, n) in enumerate(zip(g["clip_index"], g["n_samples"])):
            assert np.array_equal(synth_ref.make_clip_int16(int(g["seed"]), int(idx), int(n)), g[f"pcm_{j}"])


@pytest.mark.parametrize("flavour", ["numpy", "torchaudio"])
def test_mel_oracle_matches_reference_goldens(golden_dir, flavour):
    for path in _cases(golden_dir):
        g = np.load(path)
        sr, n_fft, hop, n_mels, norm = (int(g[k]) for k in ("sample_rate", "n_fft", "hop_length", "n_mels", "normalize"))
        for j in range(len(g["clip_index"])):
            wave = g[f"pcm_{j}"].astype(np.float32) / np.float32(32768.0)
            ref = g[f"spec_{j}"]
            fn = mel_ref.mel_db_numpy if flavour == "numpy" else mel_ref.mel_db_torchaudio
            got = fn(wave, sr, n_fft, hop, n_mels, bool(norm))
            assert got.shape == ref.shape == (n_mels, 1 + len(wave) // hop)
            if norm:
                # parity gate: <= 1e-4 absolute after min-max
                assert np.abs(got - ref).max() <= 1e-4
            else:
                rng = ref.max() - ref.min()
                assert (np.abs(got - ref) <= 1e-4 * np.maximum(np.abs(ref), rng)).all()


def test_silent_clip_is_dropped(golden_dir):
    g = np.load(os.path.join(golden_dir, "mel_1024_512_norm.npz"))
    assert int(g["silent_is_bad"]) == 1
    s = mel_ref.mel_db_numpy(np.zeros(4096, dtype=np.float32), 22050, 1024, 512, 64, True)
    assert mel_ref.is_bad(s)
    s = mel_ref.mel_db_numpy(np.zeros(4096, dtype=np.float32), 22050, 1024, 512, 64, False)
    assert not mel_ref.is_bad(s) and np.allclose(s, -100.0)


def test_npy_contract_is_fortran_order(golden_dir):
    g = np.load(os.path.join(golden_dir, "mel_1024_512_norm.npz"))
    header = bytes(g["npy_header"])
    assert b"'fortran_order': True" in header and b"'descr': '<f4'" in header


def test_too_short_clip_raises_like_reflect_pad():
    with pytest.raises(RuntimeError):
        mel_ref.mel_db_numpy(np.ones(512, dtype=np.float32), 22050, 1024, 512, 64)


def test_filterbank_and_window_against_torch():
    import torch
    import torchaudio

    fb_t = torchaudio.functional.melscale_fbanks(513, 0.0, 11025.0, 64, 22050).numpy()
    fb_n = mel_ref.melscale_fbanks_htk(513, 0.0, 11025.0, 64, 22050)
    assert np.abs(fb_t - fb_n).max() < 2e-5
    assert ((fb_t > 0) == (fb_n > 0)).mean() > 0.9999
    assert (fb_n > 0).sum() == 998 and ((fb_n > 0).sum(1) <= 2).all()
    w = torch.hann_window(1024, periodic=True).numpy()
    assert np.abs(w - mel_ref.hann_periodic(1024)).max() < 1e-6


def test_normalize_rows_matches_reference_formula():
    v = np.random.default_rng(0).random((100, 64), dtype=np.float32)
    v[5] = 0
    out = mel_ref.normalize_rows(v)
    assert out.dtype == np.float32 and (out[5] == 0).all()
    np.testing.assert_allclose(np.linalg.norm(out[[0, 1, 2]], axis=1), 1.0, rtol=1e-6)
