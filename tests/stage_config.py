"""TEST FIXTURE: a stand-in for the reference's ``AudioTokensConfig`` (audio_tokens_config.py:14-81 of
danavery/audio-tokens) carrying the fields the three hot-path stages read, with the reference's defaults, plus the knobs
that only exist in the B200 stages (read there with ``getattr(config, name, default)``).  The product never ships a config
class: the drop-in stages take the reference's own object (at_b200.dropin leaves ``audio_tokens_config`` to the reference
checkout).  /root/reference does not exist on the GPU box, hence this fixture.
"""
from __future__ import annotations

import os
from dataclasses import dataclass, field
from pathlib import Path
from typing import List, Optional

_BASE = os.path.dirname(os.path.abspath(__file__))


def _here(*parts) -> str:
    return os.path.join(_BASE, *parts)


@dataclass
class StageConfig:
    # shared
    random_seed: int = 4242

    # stage 1: SpectrogramGenerator
    split_file: str = _here("output", "bal_train_data_split.json")
    audio_source_path: str = "/media/davery/audioset"
    audio_source_sets: List[str] = field(default_factory=lambda: ["bal_train"])
    dest_spec_path: Path = Path(_here("spectrograms"))
    common_sr: int = 22050
    normalize: bool = False
    n_mels: int = 64
    n_fft: int = 512
    hop_length: int = 128
    spectrogram_batch_size: int = 5000

    # stage 2: ClusterCreator
    vocab_size: int = 500
    niter: int = 20
    use_convolution: bool = False
    num_kernels: int = 10
    kernel_size: int = 3
    clustering_batch_size: int = 10000
    centroids_path: Path = Path(_here("output", "centroids.npy"))
    source_spec_path: Path = Path(_here("spectrograms"))

    # stage 3: SpecTokenizer
    dest_tokenized_path: str = _here("tokenized_audio")
    tokenizer_batch_size: int = 10000

    # ---- knobs that exist only in the B200 implementation (reference behaviour by default) ----
    # faiss.ClusteringParameters.max_points_per_centroid; None keeps FAISS's 256 (subsample to 256*K rows)
    max_points_per_centroid: Optional[int] = None
    # sort the globbed .npy lists (the reference does not: its result depends on directory order)
    sort_files: bool = False
