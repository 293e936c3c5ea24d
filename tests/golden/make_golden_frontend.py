"""Regenerates tests/golden/frontend_*.npz by running the REFERENCE's own methods, imported unchanged from
/root/reference: SpectrogramGenerator.convert_to_mono + SpectrogramGenerator.resample
(processors/spectrogram_generator.py:109-121) on synthetic multi-channel 16-bit-PCM-valued clips.

Run in the build container only (the GPU box has no /root/reference):

    python tests/golden/make_golden_frontend.py

The .npz files are committed; tests never import the reference.
"""
import json
import os
import sys
import tempfile

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("AUDIO_TOKENS_REFERENCE", "/root/reference")
sys.path.insert(0, ROOT)
sys.path.insert(0, REF)

CASES = [  # (name, source rate, channels, samples)
    ("frontend_44100_stereo", 44100, 2, 30011),
    ("frontend_48000_mono", 48000, 1, 24000),
    ("frontend_16000_mono", 16000, 1, 9000),
]


def main():
    from audio_tokens_config import AudioTokensConfig
    from processors.spectrogram_generator import SpectrogramGenerator

    torch.set_num_threads(1)
    with tempfile.TemporaryDirectory() as tmp:
        split = os.path.join(tmp, "split.json")
        with open(split, "w") as f:
            json.dump({"train": [], "validation": []}, f)
        cfg = AudioTokensConfig()
        cfg.split_file = split
        gen = SpectrogramGenerator(cfg)
        gen.device = torch.device("cpu")
        for name, sr, ch, n in CASES:
            rng = np.random.default_rng(sr + ch)
            t = np.arange(n) / sr
            pcm = np.stack([(8000 * np.sin(2 * np.pi * (220.0 * (c + 1)) * t) + rng.integers(-3000, 3000, n)).astype(np.int16)
                            for c in range(ch)])
            wave = torch.from_numpy(pcm.astype(np.float32) / np.float32(32768.0))
            out = gen.resample(gen.convert_to_mono(wave), sr)
            np.savez_compressed(os.path.join(HERE, name + ".npz"), pcm=pcm, source_rate=sr, common_sr=cfg.common_sr,
                                out=out.numpy().astype(np.float32))
            print(name, tuple(out.shape))


if __name__ == "__main__":
    main()
