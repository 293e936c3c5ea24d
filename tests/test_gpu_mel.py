"""GPU parity of stage 1 (at_mel_* through the C ABI) against the reference goldens and the oracle."""
import glob
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _torch():
    import torch

    return torch


def _plan(sr, n_fft, hop, n_mels, norm):
    from at_b200 import MelPlan

    return MelPlan(sr, n_fft, hop, n_mels, norm)


def test_synth_device_matches_numpy_twin():
    torch = _torch()
    from at_b200 import synth_clips
    from oracle import synth_ref

    dev = synth_clips(4242, 5, 3, 30000).cpu().numpy()
    for j in range(3):
        ref = synth_ref.make_clip(4242, 5 + j, 30000)
        assert np.array_equal(dev[j], ref)


def test_mel_matches_reference_goldens(golden_dir):
    torch = _torch()
    for path in sorted(glob.glob(os.path.join(golden_dir, "mel_*.npz"))):
        g = np.load(path)
        sr, n_fft, hop, n_mels, norm = (int(g[k]) for k in ("sample_rate", "n_fft", "hop_length", "n_mels", "normalize"))
        plan = _plan(sr, n_fft, hop, n_mels, norm)
        waves = [torch.from_numpy(g[f"pcm_{j}"].astype(np.float32) / np.float32(32768.0)) for j in range(len(g["clip_index"]))]
        out, fo, bad = plan.forward_ragged(waves)
        out = out.cpu().numpy()
        assert (bad.cpu().numpy() == 0).all()
        for j in range(len(waves)):
            ref = g[f"spec_{j}"]  # (n_mels, T)
            got = out[fo[j]:fo[j + 1]].T
            assert got.shape == ref.shape
            if norm:
                assert np.abs(got - ref).max() <= 1e-4, (path, j, np.abs(got - ref).max())
            else:
                rng = ref.max() - ref.min()
                assert (np.abs(got - ref) <= 1e-4 * np.maximum(np.abs(ref), rng)).all(), (path, j)


@pytest.mark.parametrize("n_fft,hop,norm", [(1024, 512, True), (512, 128, False), (256, 64, True)])
def test_mel_matches_oracle_on_full_clips(n_fft, hop, norm):
    torch = _torch()
    from at_b200 import synth_clips
    from oracle import mel_ref

    B, L = 5, 220500 if n_fft == 1024 else 40000
    wave = synth_clips(4242, 100, B, L)
    plan = _plan(22050, n_fft, hop, 64, norm)
    spec, bad, l2 = plan.forward(wave, want_l2=True)
    spec, l2 = spec.cpu().numpy(), l2.cpu().numpy()
    assert (bad.cpu().numpy() == 0).all()
    w = wave.cpu().numpy()
    for b in range(B):
        ref = mel_ref.mel_db_torchaudio(w[b], 22050, n_fft, hop, 64, norm).T  # (T, n_mels)
        assert spec[b].shape == ref.shape
        if norm:
            assert np.abs(spec[b] - ref).max() <= 1e-4
        else:
            rng = ref.max() - ref.min()
            assert (np.abs(spec[b] - ref) <= 1e-4 * np.maximum(np.abs(ref), rng)).all()
        np.testing.assert_allclose(l2[b], mel_ref.normalize_rows(spec[b]), rtol=2e-6, atol=1e-7)


def test_mel_edge_cases_silent_short_and_uniform_vs_ragged():
    torch = _torch()
    from at_b200 import synth_clips

    plan = _plan(22050, 1024, 512, 64, True)
    w = synth_clips(1, 0, 3, 5000)
    w[1] = 0.0  # silent clip -> max == min -> NaN -> dropped by the reference
    spec, bad = plan.forward(w)
    assert bad.cpu().tolist() == [0, 1, 0]
    assert torch.isnan(spec[1]).all()
    # ragged path == uniform path, bit for bit
    out, fo, bad2 = plan.forward_ragged([w[0].cpu(), w[2].cpu()])
    assert torch.equal(out[fo[0]:fo[1]], spec[0]) and torch.equal(out[fo[1]:fo[2]], spec[2])
    # too short for reflect padding (torch raises): flag 2
    out, fo, bad3 = plan.forward_ragged([torch.ones(512), w[0].cpu()])
    assert bad3.cpu().tolist() == [2, 0]
    assert torch.equal(out[fo[1]:fo[2]], spec[0])
    # not normalised: silent clip is finite (-100 dB everywhere)
    plan2 = _plan(22050, 1024, 512, 64, False)
    s2, b2 = plan2.forward(torch.zeros(1, 4096, device="cuda"))
    assert b2.item() == 0 and torch.allclose(s2, torch.full_like(s2, -100.0))


def test_mel_host_entry_point_matches_device_path():
    torch = _torch()
    from at_b200 import synth_clips

    plan = _plan(22050, 1024, 512, 64, True)
    w = synth_clips(7, 0, 40, 22050)
    spec, bad = plan.forward(w)
    out, badh = plan.forward_host(w.cpu().numpy())
    assert np.array_equal(out, spec.cpu().numpy()) and (badh == 0).all()


def test_builtin_constants_close_to_torch_constants():
    torch = _torch()
    from at_b200 import MelPlan, synth_clips

    w = synth_clips(3, 0, 2, 22050)
    a, _ = MelPlan(22050, 1024, 512, 64, True, torch_constants=True).forward(w)
    b, _ = MelPlan(22050, 1024, 512, 64, True, torch_constants=False).forward(w)
    assert (a - b).abs().max().item() < 1e-4


@pytest.mark.parametrize("orig,new,channels", [(44100, 22050, 2), (48000, 22050, 1), (16000, 22050, 1), (32000, 22050, 2), (22050, 22050, 2)])
def test_mono_resample_matches_torchaudio(orig, new, channels):
    """Channel mean + Resample on the device against torch.mean + torchaudio.transforms.Resample on the CPU
    (processors/spectrogram_generator.py:109-121).  Tolerance: 2e-6 absolute on samples in [-1, 1)."""
    import torch
    from at_b200 import ResamplePlan
    from oracle import resample_ref

    rng = np.random.default_rng(orig + channels)
    L = orig + 123
    wave = (rng.integers(-20000, 20000, size=(channels, L)).astype(np.float32) / 32768.0)
    want = resample_ref.resample_torchaudio(wave, orig, new)
    plan = ResamplePlan(orig, new)
    got = plan.forward(torch.from_numpy(wave).cuda()).cpu().numpy()
    assert got.shape == want.shape
    assert np.abs(got - want).max() <= 2e-6


def test_mono_resample_batch_equals_per_clip():
    import torch
    from at_b200 import ResamplePlan

    g = torch.Generator("cuda").manual_seed(3)
    wave = (torch.randint(-20000, 20000, (5, 2, 48000 + 17), device="cuda", generator=g).float() / 32768.0)
    plan = ResamplePlan(48000, 22050)
    batch = plan.forward_batch(wave)
    for i in range(5):
        assert torch.equal(batch[i:i + 1], plan.forward(wave[i]))


@pytest.mark.parametrize("name", ["frontend_44100_stereo", "frontend_48000_mono", "frontend_16000_mono"])
def test_mono_resample_matches_reference_goldens(golden_dir, name):
    """Device channel mean + resampler against the outputs of the reference's own convert_to_mono + resample methods
    (tests/golden/frontend_*.npz)."""
    import os
    import torch
    from at_b200 import ResamplePlan, pcm16_to_f32

    g = np.load(os.path.join(golden_dir, name + ".npz"))
    wave = pcm16_to_f32(torch.from_numpy(g["pcm"]).cuda())
    got = ResamplePlan(int(g["source_rate"]), int(g["common_sr"])).forward(wave).cpu().numpy()
    assert got.shape == g["out"].shape
    assert np.abs(got - g["out"]).max() <= 2e-6
