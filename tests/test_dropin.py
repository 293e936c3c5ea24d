"""The drop-in boundary at level L-A / L-B: the reference's ``run_pipeline.py`` runs unmodified and
its three hot-path stages resolve to audio-tokens_b200/processors/, while everything else keeps resolving to the
checkout -- whatever the order of sys.path.

CPU tests check the import routing (a stand-in checkout, and the real /root/reference when it is mounted); the GPU test
drives the whole stand-in pipeline through the launcher and compares the three output trees with the oracle."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "audio-tokens_b200")
SITE = os.path.join(PKG, "dropin_site")
STUBS = os.path.join(ROOT, "tests", "fixtures", "stubs")
REF = "/root/reference"

PROBE = r"""
import json, runpy, sys
ck = sys.argv[1]
sys.path.insert(0, ck)                       # what `python <checkout>/run_pipeline.py` does before running the script
ns = runpy.run_path(ck + "/run_pipeline.py", run_name="probe")   # imports only: main() is guarded by __name__
import audio_tokens_config, processors, processors.model_trainer as mt
out = {n: sys.modules[ns[n].__module__].__file__ for n in ("SpectrogramGenerator", "ClusterCreator", "SpecTokenizer", "ModelTrainer")}
out["config"] = audio_tokens_config.__file__
out["processors"] = processors.__file__
print(json.dumps(out))
"""


def _env(*extra_path):
    env = dict(os.environ)
    env["PYTHONPATH"] = os.pathsep.join([SITE, *extra_path])
    env.pop("AUDIO_TOKENS_REFERENCE", None)
    return env


def _assert_routing(out, checkout):
    for name, mod in (("SpectrogramGenerator", "spectrogram_generator"), ("ClusterCreator", "cluster_creator"),
                      ("SpecTokenizer", "spec_tokenizer")):
        assert out[name] == os.path.join(PKG, "processors", f"{mod}.py"), (name, out[name])
    for name in ("ModelTrainer", "config", "processors"):
        assert out[name].startswith(checkout + os.sep), (name, out[name])   # everything else: the checkout's own


def test_sitecustomize_routes_only_the_three_stage_modules(tmp_path):
    """PYTHONPATH=<repo>/audio-tokens_b200/dropin_site + the checkout at sys.path[0] (the documented command line)."""
    from standin_checkout import make_checkout

    ck, _ = make_checkout(str(tmp_path), [f"clip{i:04d}" for i in range(8)])
    r = subprocess.run([sys.executable, "-c", PROBE, ck], env=_env(), capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    _assert_routing(json.loads(r.stdout.strip().splitlines()[-1]), ck)


def test_documented_command_reaches_the_b200_stage_classes(tmp_path):
    """`python <checkout>/run_pipeline.py` itself: without a GPU the run must stop inside THIS repo's SpectrogramGenerator
    (its no-CPU-fallback error), not inside the checkout's decoy; with a GPU it is covered by the test below."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("covered by test_launcher_runs_the_whole_pipeline on a GPU box")
    from standin_checkout import make_checkout

    ck, _ = make_checkout(str(tmp_path), [f"clip{i:04d}" for i in range(8)])
    r = subprocess.run([sys.executable, os.path.join(ck, "run_pipeline.py")], env=_env(), capture_output=True, text=True,
                       timeout=600, cwd=ck)
    assert r.returncode != 0
    assert "SpectrogramGenerator (B200 build) needs a CUDA device" in r.stderr, r.stderr[-2000:]


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "processors")), reason="the reference checkout is not mounted")
def test_routing_against_the_real_reference_checkout():
    """Same probe against the UNMODIFIED /root/reference/run_pipeline.py (matplotlib, absent from this image, is stubbed)."""
    r = subprocess.run([sys.executable, "-c", PROBE, REF], env=_env(STUBS), capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-3000:]
    _assert_routing(json.loads(r.stdout.strip().splitlines()[-1]), REF)


def test_install_is_idempotent_and_reversible(tmp_path):
    from at_b200 import dropin

    n0 = len(sys.meta_path)
    f1 = dropin.install()
    f2 = dropin.install(operators=False)
    assert f1 is f2 and len(sys.meta_path) == n0 + 1 and sys.meta_path[0] is f1
    assert f1.find_spec("processors.model_trainer") is None and f1.find_spec("numpy") is None
    spec = f1.find_spec("processors.cluster_creator")
    assert spec is not None and spec.origin == os.path.join(PKG, "processors", "cluster_creator.py")
    dropin.uninstall()
    assert len(sys.meta_path) == n0 and not dropin.installed()


def test_faiss_name_resolves_to_the_compat_module_only_when_faiss_is_absent():
    import importlib.util

    from at_b200 import dropin

    had_real = importlib.util.find_spec("faiss") is not None
    dropin.install(operators=True, stages=False)
    try:
        sys.modules.pop("faiss", None)
        import faiss

        if had_real:
            assert not getattr(faiss, "__file__", "").endswith("faiss_compat.py")
        else:
            import at_b200.faiss_compat as fc

            assert faiss is fc and hasattr(faiss, "Kmeans") and hasattr(faiss, "IndexFlatL2") and hasattr(faiss, "get_num_gpus")
        import torchaudio.transforms as T
        from at_b200 import torchaudio_compat as C

        assert T.MelSpectrogram is C.MelSpectrogram and T.AmplitudeToDB is C.AmplitudeToDB
    finally:
        dropin.uninstall()
        sys.modules.pop("faiss", None)
    import torchaudio.transforms as T

    assert T.MelSpectrogram.__module__.startswith("torchaudio")


# ------------------------------------------------------------------------------------------------------- GPU
SR, L = 22050, 22050 * 3


def _wave(ytid):
    import torch
    from oracle import synth_ref

    idx = int(ytid[4:])
    if idx == 3:
        return torch.zeros(1, L)          # silent clip: NaN after min-max -> dropped like the reference does
    if idx == 6:                          # a stereo clip at another rate: channel mean + resample in front of the path
        w = synth_ref.make_clip(4242, idx, 2 * 44100).reshape(1, -1)
        return torch.from_numpy(np.concatenate([w, 0.5 * w], axis=0))
    return torch.from_numpy(synth_ref.make_clip(4242, idx, L if idx % 5 else L - 777)).reshape(1, -1)


@pytest.mark.gpu
def test_launcher_runs_the_whole_pipeline(tmp_path, monkeypatch):
    """python -m at_b200.run_pipeline --reference <checkout>: the checkout's run_pipeline.py executes as __main__, the
    three stages are this repo's, the config and the trainer are the checkout's; outputs equal the oracle's."""
    import torch
    import torchaudio
    from at_b200 import dropin, run_pipeline
    from oracle import faiss_ref, mel_ref, resample_ref
    from standin_checkout import make_checkout

    ytids = [f"clip{i:04d}" for i in range(24)]
    ck, cfg = make_checkout(str(tmp_path), ytids)
    os.remove(os.path.join(cfg["audio_source_path"], "bal_train", "cl", "clip0005.flac"))   # a missing file is skipped

    def fake_load(path, *a, **k):   # torchaudio.load needs torchcodec, absent from this image
        y = os.path.basename(str(path))[:-5]
        return _wave(y), (44100 if y == "clip0006" else SR)

    monkeypatch.setattr(torchaudio, "load", fake_load)
    saved_path, saved_argv = list(sys.path), list(sys.argv)
    # a fresh interpreter as far as the checkout's module names are concerned (other tests import this repo's
    # ``processors`` package directly); put everything back afterwards
    ours = {m: sys.modules.pop(m) for m in list(sys.modules)
            if m == "audio_tokens_config" or m == "processors" or m.startswith("processors.")}
    monkeypatch.chdir(ck)
    try:
        run_pipeline.main(["--reference", ck])
    finally:
        dropin.uninstall()
        sys.path[:] = saved_path
        sys.argv[:] = saved_argv
        for m in list(sys.modules):
            if m == "audio_tokens_config" or m == "processors" or m.startswith("processors."):
                del sys.modules[m]
        sys.modules.update(ours)

    # the checkout's trainer ran last, with the checkout's config
    tr = json.load(open(os.path.join(ck, "output", "trainer.json")))
    assert tr["marker"] == "checkout-config" and tr["config_file"].startswith(ck)
    # ---- stage 1 files vs the oracle (reference torchaudio calls on the CPU)
    spec_dir = os.path.join(ck, "spectrograms")
    train = sorted(f for f in os.listdir(os.path.join(spec_dir, "train")))
    assert "clip0003.npy" not in train and "clip0005.npy" not in train and len(train) == 18
    assert tr["tokens"] == train
    for f in train[:8]:
        with open(os.path.join(spec_dir, "train", f), "rb") as fh:
            header = fh.read(128)
        assert b"'fortran_order': True" in header and b"'descr': '<f4'" in header
        s = np.load(os.path.join(spec_dir, "train", f))
        w = _wave(f[:-4])
        if f == "clip0006.npy":
            w = torch.from_numpy(resample_ref.resample_torchaudio(w.numpy(), 44100, SR)).reshape(1, -1)
        ref = mel_ref.mel_db_torchaudio(w[0].numpy(), SR, 1024, 512, 64, True)
        assert s.shape == ref.shape and np.abs(s - ref).max() <= 1e-4
    # ---- stage 2 vs the FAISS restatement
    cents = np.load(cfg["centroids_path"])
    assert cents.shape == (32, 64) and cents.dtype == np.float32 and not np.isfortran(cents)
    x = np.concatenate([np.load(os.path.join(spec_dir, "train", f)).T for f in train], axis=0).astype(np.float32)
    km = faiss_ref.Kmeans(64, 32, niter=6)
    km.exact_search = True
    km.train(mel_ref.normalize_rows(x))
    ref_c = mel_ref.normalize_rows(km.centroids)
    rel = np.linalg.norm(cents - ref_c, axis=1) / np.linalg.norm(ref_c, axis=1)
    assert (rel <= 1e-4).mean() >= 0.9, rel.max()
    # ---- stage 3 vs the scalar fp32 formula on the saved centroids
    for split in ("train", "validation"):
        for f in sorted(os.listdir(os.path.join(spec_dir, split))):
            t = np.load(os.path.join(ck, "tokenized_audio", split, f))
            s = np.load(os.path.join(spec_dir, split, f)).T.astype(np.float32)
            assert t.dtype == np.int64 and t.shape == (s.shape[0],)
            lab, d1, d2 = faiss_ref.assign_l2_scalar(mel_ref.normalize_rows(s), cents)
            mism = t != lab
            assert ((d2 - d1)[mism] <= 1e-6 * np.maximum(d1[mism], 1e-30) + 1e-9).all()


@pytest.mark.gpu
def test_operator_mirrors_run_the_reference_style_generator():
    """Level L-B: code written like the reference's generate_mel_spectrogram (spectrogram_generator.py:28-34,123-126)
    over torchaudio.transforms, with the operators routed to the B200 mirrors, equals the oracle."""
    import torch
    from at_b200 import dropin
    from oracle import mel_ref, synth_ref

    # the oracle makes the reference's torchaudio calls on the CPU: evaluate it BEFORE the operators are routed
    cases = []
    for n_fft, hop, n in ((1024, 512, 22050), (512, 128, 8000)):
        wave = torch.from_numpy(synth_ref.make_clip(4242, 11, n)).reshape(1, -1)
        ref_db = mel_ref.mel_db_torchaudio(wave[0].numpy(), SR, n_fft, hop, 64, False)
        import torchaudio.transforms as T

        ref_pow = T.MelSpectrogram(sample_rate=SR, n_mels=64, n_fft=n_fft, hop_length=hop)(wave)[0].numpy()
        cases.append((n_fft, hop, n, wave, ref_db, ref_pow))
    dropin.install(operators=True, stages=False)
    try:
        from torchaudio.transforms import AmplitudeToDB, MelSpectrogram
        from at_b200 import torchaudio_compat

        assert MelSpectrogram is torchaudio_compat.MelSpectrogram
        device = torch.device("cuda")
        for n_fft, hop, n, wave, ref_db, ref_pow in cases:
            spec_transformer = MelSpectrogram(sample_rate=SR, n_mels=64, n_fft=n_fft, hop_length=hop).to(device)
            amplitude_to_db_transformer = AmplitudeToDB().to(device)
            mel_spec = spec_transformer(wave.to(device)).squeeze(0)
            mel_spec_db = amplitude_to_db_transformer(mel_spec)
            assert tuple(mel_spec_db.shape) == (64, 1 + n // hop)
            got = mel_spec_db.cpu().numpy()
            assert np.isfortran(got)   # the transposed view torch's own MelScale hands back: np.save writes fortran_order
            assert np.abs(got - ref_db).max() <= 1e-4 * max(np.abs(ref_db).max(), ref_db.max() - ref_db.min())
            # the power output itself, against torchaudio's MelSpectrogram on the CPU
            assert np.abs(mel_spec.cpu().numpy() - ref_pow).max() <= 2e-5 * ref_pow.max() + 1e-12
    finally:
        dropin.uninstall()
