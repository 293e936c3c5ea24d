"""CPU-side checks of the boundary: the C-ABI library loads and exports every symbol the header declares;
compute entry points fail loudly without a device (no CPU fallback)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "audio_tokens_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(at_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from at_b200 import _lib

    lib = _lib.load()
    names = _declared_symbols()
    assert len(names) >= 25
    for name in names:
        assert hasattr(lib, name), name
    # and the binding table covers the header one to one
    assert sorted(_lib.PROTOTYPES) == names


def test_version_and_host_rand_perm():
    from at_b200 import _lib
    from at_b200.kmeans import rand_perm

    assert _lib.load().at_version() == 100
    assert rand_perm(10, 1234).tolist() == [5, 4, 8, 2, 6, 9, 1, 7, 0, 3]
    from oracle import faiss_ref

    assert (rand_perm(5000, 1235) == faiss_ref.rand_perm(5000, 1235)).all()


def test_no_cpu_fallback():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from at_b200 import _lib

    lib = _lib.load()
    h = ctypes.c_void_p()
    assert lib.at_index_create(64, ctypes.byref(h)) == -2  # AT_ERR_CUDA
    assert b"cudaGetDevice" in lib.at_last_error()
    assert lib.at_mel_plan_create(22050, 1024, 512, 64, 1, ctypes.byref(h)) == -2
    with pytest.raises(RuntimeError):
        from at_b200 import MelPlan

        MelPlan(22050, 1024, 512, 64, True)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "audio-tokens_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src, os.path.join(dirpath, f)


def test_library_holds_blackwell_tensor_and_bulk_copy_instructions():
    """The built library is sm_100a code whose SASS carries the 5th-generation tensor-core path (tcgen05.mma -> UTC*MMA,
    tcgen05.ld -> LDTM), the bulk-copy engine (cp.async.bulk -> UBLKCP) and packed fp32 (FFMA2 / FADD2), and none of the
    legacy tensor instructions (mma.sync -> HMMA): a changed build flag or a silent fallback would show here."""
    import collections
    import re
    import shutil
    import subprocess

    from at_b200 import _lib

    tool = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(tool) and not shutil.which("cuobjdump"):
        pytest.skip("cuobjdump not available")
    out = subprocess.run([tool, "-sass", _lib.LIB_PATH], capture_output=True, text=True, timeout=300).stdout
    archs = set(re.findall(r"arch = (sm_\w+)", out))
    assert archs == {"sm_100a"}, archs
    ops = collections.Counter(m.group(1).split(".")[0] for m in re.finditer(r"^\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]+)", out, re.M))
    assert ops["UTCHMMA"] >= 10 and ops["LDTM"] >= 8 and ops["UBLKCP"] >= 4, {k: ops[k] for k in ("UTCHMMA", "LDTM", "UBLKCP")}
    assert ops["FFMA2"] > 0 and ops["FADD2"] > 0
    assert ops["HMMA"] == 0 and ops["HGMMA"] == 0
    fns = set(re.findall(r"Function : (\S+)", out))
    for name in ("k_mel", "k_assign_tc", "k_assign_tc_wide", "k_tc_tail", "k_tc_full", "k_gather_sum", "k_finalize_split",
                 "k_conv_expand", "k_assign_gemm", "k_peer", "k_tokens"):
        assert any(name in f for f in fns), name
