"""at_b200 -- host side of the B200-native audio-tokens hot path.

Python mirrors of the library operators the reference calls (torchaudio MelSpectrogram + AmplitudeToDB,
faiss.Kmeans, faiss.IndexFlatL2) on top of libat_b200.so (hand-written sm_100a kernels behind the C ABI in
include/audio_tokens_b200.h).  torch is used for device memory, streams and torch.distributed only.
"""
from . import _lib
from .index import FlatL2, IndexFlatL2
from .kmeans import ClusteringParameters, Kmeans, LloydTrainer
from .mel import MelPlan
from .resample import ResamplePlan
from .synth import sine_table, synth_clips

__all__ = [
    "_lib", "MelPlan", "ResamplePlan", "FlatL2", "IndexFlatL2", "Kmeans", "ClusteringParameters", "LloydTrainer",
    "synth_clips", "sine_table", "get_num_gpus", "row_l2norm", "pcm16_to_f32",
]


def pcm16_to_f32(pcm, out=None):
    """16-bit PCM CUDA tensor -> fp32 waveform in [-1, 1) (sample / 32768: what torchaudio.load yields for a 16-bit
    file, processors/spectrogram_generator.py:99)."""
    import torch

    _lib.require_cuda()
    assert pcm.is_cuda and pcm.dtype == torch.int16 and pcm.is_contiguous()
    if out is None:
        out = torch.empty(pcm.shape, dtype=torch.float32, device=pcm.device)
    assert out.is_cuda and out.dtype == torch.float32 and out.is_contiguous() and out.numel() >= pcm.numel()
    _lib.check(_lib.load().at_pcm16_to_f32(_lib.ptr(pcm), pcm.numel(), _lib.ptr(out), _lib.stream_ptr()))
    return out


def get_num_gpus() -> int:
    """faiss.get_num_gpus() (processors/cluster_creator.py:26)."""
    import torch

    return torch.cuda.device_count() if torch.cuda.is_available() else 0


def row_l2norm(x):
    """normalize_vectors (cluster_creator.py:64-66) on a CUDA tensor (n, d) fp32 -> new tensor."""
    import torch

    _lib.require_cuda()
    assert x.is_cuda and x.dtype == torch.float32 and x.dim() == 2 and x.is_contiguous()
    out = torch.empty_like(x)
    _lib.check(_lib.load().at_row_l2norm(_lib.ptr(x), x.shape[0], x.shape[1], _lib.ptr(out), _lib.stream_ptr()))
    return out
