"""Regenerates tests/golden/mel_*.npz by running the REFERENCE's own class, imported unchanged from
/root/reference (processors/spectrogram_generator.py), on synthetic clips from oracle/synth_ref.py.

Run in the build container only (the GPU box has no /root/reference):

    python tests/golden/make_golden.py

The .npz files are committed; tests never import the reference.
"""
import json
import os
import sys
import tempfile
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("AUDIO_TOKENS_REFERENCE", "/root/reference")
sys.path.insert(0, ROOT)
sys.path.insert(0, REF)

from oracle import synth_ref  # noqa: E402

CASES = {
    # name: (config overrides, [(clip index, n_samples)])
    "mel_1024_512_norm": (dict(n_fft=1024, hop_length=512, normalize=True), [(0, 22050), (1, 11111), (2, 5000), (3, 513)]),
    "mel_512_128_raw": (dict(n_fft=512, hop_length=128, normalize=False), [(4, 8000), (5, 4097)]),
    "mel_256_64_norm": (dict(n_fft=256, hop_length=64, normalize=True), [(6, 3000)]),
}


def main():
    from audio_tokens_config import AudioTokensConfig
    from processors.spectrogram_generator import SpectrogramGenerator

    torch.set_num_threads(1)
    with tempfile.TemporaryDirectory() as tmp:
        split = os.path.join(tmp, "split.json")
        with open(split, "w") as f:
            json.dump({"train": [], "validation": []}, f)
        for name, (over, clips) in CASES.items():
            cfg = AudioTokensConfig()
            cfg.split_file = split
            for k, v in over.items():
                setattr(cfg, k, v)
            gen = SpectrogramGenerator(cfg)
            gen.device = torch.device("cpu")
            out = {"sample_rate": cfg.common_sr, "n_fft": cfg.n_fft, "hop_length": cfg.hop_length,
                   "n_mels": cfg.n_mels, "normalize": int(cfg.normalize), "seed": 4242,
                   "clip_index": np.array([c[0] for c in clips]), "n_samples": np.array([c[1] for c in clips])}
            for j, (idx, n) in enumerate(clips):
                pcm = synth_ref.make_clip_int16(4242, idx, n)
                wave = torch.from_numpy(pcm.astype(np.float32) / np.float32(32768.0)).reshape(1, -1)
                spec = gen.generate_mel_spectrogram(wave)
                if cfg.normalize:
                    spec = gen.normalize_spectrogram(spec)
                assert not gen.check_for_nan_inf(spec)
                out[f"pcm_{j}"] = pcm
                out[f"spec_{j}"] = spec.numpy().astype(np.float32)
            # silent clip: reference yields NaN after min-max and drops the clip
            if cfg.normalize:
                silent = torch.zeros(1, 4096)
                s = gen.normalize_spectrogram(gen.generate_mel_spectrogram(silent))
                out["silent_is_bad"] = int(gen.check_for_nan_inf(s))
            # the saved-file contract (fortran_order header) for one clip
            path = os.path.join(tmp, "x.npy")
            np.save(path, spec.cpu())
            with open(path, "rb") as f:
                out["npy_header"] = np.frombuffer(f.read(128), dtype=np.uint8)
            np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
            print(name, {k: (v.shape if hasattr(v, "shape") else v) for k, v in out.items() if k.startswith("spec")})


if __name__ == "__main__":
    main()
