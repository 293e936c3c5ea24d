"""ctypes binding of libat_b200.so (the C ABI declared in include/audio_tokens_b200.h).

The shared library is built in-tree by ``audio-tokens_b200/csrc/build.sh`` (``__graft_entry__.build()``).
There is no fallback: if the library is missing, or a call fails, a RuntimeError is raised.
"""
from __future__ import annotations

import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# AT_B200_LIB: an experiment build of the same library (tools/ only; the product path is the in-tree file)
LIB_PATH = os.environ.get("AT_B200_LIB") or os.path.join(_HERE, "libat_b200.so")

c_i64 = ctypes.c_int64
c_int = ctypes.c_int
c_ptr = ctypes.c_void_p
c_f32 = ctypes.c_float

# name -> (restype, argtypes); mirrors include/audio_tokens_b200.h one to one
PROTOTYPES = {
    "at_version": (c_int, []),
    "at_last_error": (ctypes.c_char_p, []),
    "at_device_info": (c_int, [c_ptr, c_ptr, c_ptr]),
    "at_kernel_launches": (c_i64, []),
    "at_profile_enable": (c_int, [c_int]),
    "at_profile_summary": (c_int, [c_int, c_ptr, c_ptr]),
    "at_mel_plan_create": (c_int, [c_int, c_int, c_int, c_int, c_int, c_ptr]),
    "at_mel_plan_set_constants_host": (c_int, [c_ptr, c_ptr, c_ptr]),
    "at_mel_plan_set_output": (c_int, [c_ptr, c_int]),
    "at_mel_plan_destroy": (c_int, [c_ptr]),
    "at_mel_num_frames": (c_i64, [c_ptr, c_i64]),
    "at_mel_work_groups": (c_int, [c_ptr]),
    "at_mel_plan_set_absmax_out": (c_int, [c_ptr, c_ptr]),
    "at_mel_forward": (c_int, [c_ptr, c_ptr, c_ptr, c_ptr, c_i64, c_int, c_ptr, c_ptr, c_ptr, c_ptr]),
    "at_mel_forward_host": (c_int, [c_ptr, c_ptr, c_i64, c_int, c_ptr, c_ptr]),
    "at_amplitude_to_db": (c_int, [c_ptr, c_i64, c_f32, c_f32, c_f32, c_ptr, c_ptr]),
    "at_row_l2norm": (c_int, [c_ptr, c_i64, c_int, c_ptr, c_ptr]),
    "at_conv_expand": (c_int, [c_ptr, c_i64, c_int, c_ptr, c_ptr, c_int, c_int, c_ptr, c_ptr]),
    "at_index_create": (c_int, [c_int, c_ptr]),
    "at_index_destroy": (c_int, [c_ptr]),
    "at_index_set_centroids": (c_int, [c_ptr, c_ptr, c_int, c_ptr]),
    "at_index_ntotal": (c_int, [c_ptr]),
    "at_index_set_tc_mode": (c_int, [c_ptr, c_int]),
    "at_index_centroids": (c_ptr, [c_ptr]),
    "at_index_tc_stats": (c_int, [c_ptr, c_ptr]),
    "at_index_search": (c_int, [c_ptr, c_ptr, c_i64, c_int, c_int, c_ptr, c_ptr, c_ptr, c_ptr]),
    "at_kmeans_create": (c_int, [c_int, c_int, c_ptr]),
    "at_kmeans_destroy": (c_int, [c_ptr]),
    "at_kmeans_set_centroids": (c_int, [c_ptr, c_ptr, c_ptr]),
    "at_kmeans_centroids": (c_ptr, [c_ptr]),
    "at_kmeans_get_centroids": (c_int, [c_ptr, c_ptr, c_ptr]),
    "at_absmax": (c_int, [c_ptr, c_i64, c_ptr, c_ptr]),
    "at_kmeans_begin": (c_int, [c_ptr, c_f32, c_i64]),
    "at_kmeans_begin_on": (c_int, [c_ptr, c_f32, c_i64, c_ptr]),
    "at_kmeans_accum_words": (c_i64, [c_ptr]),
    "at_kmeans_accumulate": (c_int, [c_ptr, c_ptr, c_i64, c_int, c_int, c_ptr, c_ptr, c_ptr]),
    "at_kmeans_finalize": (c_int, [c_ptr, c_ptr, c_i64, c_ptr, c_ptr]),
    "at_kmeans_set_incremental": (c_int, [c_ptr, c_int]),
    "at_kmeans_invalidate": (c_int, [c_ptr]),
    "at_index_search_trained_rows": (c_int, [c_ptr, c_ptr, c_ptr, c_i64, c_int, c_ptr, c_ptr, c_ptr, c_ptr]),
    "at_peer_create": (c_int, [c_int, c_int, c_i64, c_ptr]),
    "at_peer_export": (c_int, [c_ptr, c_ptr]),
    "at_peer_import": (c_int, [c_ptr, c_int, c_ptr]),
    "at_peer_connect": (c_int, [c_ptr]),
    "at_peer_destroy": (c_int, [c_ptr]),
    "at_peer_local_buffer": (c_ptr, [c_ptr]),
    "at_peer_total": (c_ptr, [c_ptr]),
    "at_peer_reduce": (c_int, [c_ptr, c_ptr]),
    "at_peer_status": (c_int, [c_ptr, c_ptr]),
    "at_pcm16_to_f32": (c_int, [c_ptr, c_i64, c_ptr, c_ptr]),
    "at_resample_plan_create": (c_int, [c_int, c_int, c_ptr]),
    "at_resample_bank_host": (c_int, [c_int, c_int, c_ptr, c_ptr, c_ptr, c_ptr]),
    "at_resample_plan_destroy": (c_int, [c_ptr]),
    "at_resample_out_len": (c_i64, [c_ptr, c_i64]),
    "at_resample_mono": (c_int, [c_ptr, c_ptr, c_int, c_i64, c_ptr, c_ptr]),
    "at_resample_mono_batch": (c_int, [c_ptr, c_ptr, c_int, c_i64, c_int, c_ptr, c_ptr]),
    "at_bincount": (c_int, [c_ptr, c_i64, c_int, c_ptr, c_ptr]),
    "at_tokens_collate": (c_int, [c_ptr, c_int, c_ptr, c_ptr, c_int, c_int, c_int, c_ptr, c_ptr, c_ptr]),
    "at_tokens_multihot": (c_int, [c_ptr, c_ptr, c_ptr, c_int, c_int, c_ptr, c_ptr]),
    "at_tokens_batch_max_len": (c_int, [c_ptr, c_ptr, c_int, c_ptr, c_ptr]),
    "at_synth_clips": (c_int, [ctypes.c_uint32, c_i64, c_i64, c_i64, c_ptr, c_ptr, c_ptr]),
    "at_rand_perm_host": (c_int, [c_ptr, c_i64, c_i64]),
}

ALGO_AUTO, ALGO_SIMT, ALGO_TENSOR = 0, 1, 2

_lib = None


def load():
    """Load the library (once).  Fails loudly when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `bash audio-tokens_b200/csrc/build.sh` "
                "(or `python -c 'import __graft_entry__ as g; g.build()'`). There is no CPU fallback."
            )
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in PROTOTYPES.items():
            if os.environ.get("AT_B200_LIB") and not hasattr(lib, name):
                continue  # an older experiment build (tools/ only)
            fn = getattr(lib, name)  # AttributeError if the .so does not export a declared symbol
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def check(rc: int):
    if rc != 0:
        msg = load().at_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"libat_b200 error {rc}: {msg}")


def stream_ptr():
    """The current torch CUDA stream as a void* for the C ABI."""
    import torch

    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def ctypes_stream(stream):
    """A torch.cuda.Stream as a void* for the C ABI."""
    return ctypes.c_void_p(stream.cuda_stream)


def ptr(t):
    """Device (or host) pointer of a torch tensor / numpy array, or NULL for None."""
    if t is None:
        return ctypes.c_void_p(0)
    if hasattr(t, "data_ptr"):
        return ctypes.c_void_p(t.data_ptr())
    return ctypes.c_void_p(t.ctypes.data)


def require_cuda():
    import torch

    if not torch.cuda.is_available():
        raise RuntimeError("audio-tokens_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
