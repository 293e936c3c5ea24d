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
    "synth_clips", "sine_table", "get_num_gpus", "row_l2norm", "pcm16_to_f32", "conv_expand", "make_conv_layer",
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


def make_conv_layer(config):
    """The never-trained nn.Conv1d(1 -> num_kernels, kernel_size, padding = kernel_size // 2) the reference builds right
    after set_seed(config.random_seed) (cluster_creator.py:28-34, spec_tokenizer.py:115-121): created on the CPU exactly as
    there so the seeded default initialisation yields the same weights, then its (weight, bias) are moved to the device as
    plain fp32 tensors for at_conv_expand.  Returns (weight (num_kernels, kernel_size), bias (num_kernels,))."""
    import torch

    _lib.require_cuda()
    conv = torch.nn.Conv1d(in_channels=1, out_channels=config.num_kernels, kernel_size=config.kernel_size,
                           padding=config.kernel_size // 2)
    w = conv.weight.detach().reshape(config.num_kernels, config.kernel_size).to(device="cuda", dtype=torch.float32)
    b = conv.bias.detach().to(device="cuda", dtype=torch.float32)
    return w.contiguous(), b.contiguous()


def conv_expand(x, weight, bias, out=None):
    """apply_convolution (cluster_creator.py:68-81, spec_tokenizer.py:92-104) on the device: x (n, n_mels) fp32 CUDA ->
    (n, n_mels * num_kernels), column m * num_kernels + c = conv channel c at mel bin m (at_conv_expand)."""
    import torch

    _lib.require_cuda()
    assert x.is_cuda and x.dtype == torch.float32 and x.dim() == 2 and x.is_contiguous()
    assert weight.is_cuda and weight.dtype == torch.float32 and weight.dim() == 2 and weight.is_contiguous()
    assert bias.is_cuda and bias.dtype == torch.float32 and bias.shape == (weight.shape[0],) and bias.is_contiguous()
    n, n_mels = x.shape
    kc, ks = weight.shape
    if out is None:
        out = torch.empty((n, n_mels * kc), dtype=torch.float32, device=x.device)
    assert out.is_cuda and out.dtype == torch.float32 and out.is_contiguous() and out.numel() == n * n_mels * kc
    _lib.check(_lib.load().at_conv_expand(_lib.ptr(x), n, n_mels, _lib.ptr(weight), _lib.ptr(bias), kc, ks, _lib.ptr(out),
                                          _lib.stream_ptr()))
    return out
