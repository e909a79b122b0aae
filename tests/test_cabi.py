"""The C-ABI library: it builds for sm_100a without a GPU, loads, and exports exactly the symbols include/ccedit_b200.h
declares; the ctypes mirror (ccedit_b200/_lib.py) agrees with the header; no compute call is made here."""
import ctypes
import os
import re
import subprocess

import pytest

from conftest import ROOT

HEADER = os.path.join(ROOT, "include", "ccedit_b200.h")


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ccedit_[a-z0-9_]+)\s*\(", src)))


@pytest.fixture(scope="module")
def lib_path():
    from ccedit_b200 import build
    return build.build()          # no-op when the in-tree .so is newer than its sources


def test_header_declares_the_expected_entry_points():
    syms = declared_symbols()
    for s in ("ccedit_gemm", "ccedit_attention", "ccedit_temporal_attention", "ccedit_groupnorm_spatial",
              "ccedit_groupnorm_temporal", "ccedit_layernorm", "ccedit_last_error", "ccedit_launch_count"):
        assert s in syms


def test_library_exports_every_declared_symbol(lib_path):
    lib = ctypes.CDLL(lib_path)
    for s in declared_symbols():
        assert hasattr(lib, s), f"{s} declared in include/ccedit_b200.h but not exported by {lib_path}"
    out = subprocess.run(["nm", "-D", "--defined-only", lib_path], capture_output=True, text=True).stdout
    exported = sorted(set(re.findall(r" T (ccedit_[a-z0-9_]+)$", out, flags=re.M)))
    assert exported == declared_symbols(), "exported C symbols and header declarations differ"


def test_ctypes_mirror_matches_header(lib_path):
    from ccedit_b200 import _lib
    assert sorted(_lib.SIGNATURES) == declared_symbols()
    lib = _lib.load()
    assert lib.ccedit_abi_version() == 3
    assert lib.ccedit_launch_count() >= 0
    # struct sizes agree with the C compiler's layout of the header
    prog = r'''
#include <stdio.h>
#include "ccedit_b200.h"
int main(void) { printf("%zu %zu\n", sizeof(ccedit_gemm_desc), sizeof(ccedit_attn_desc)); return 0; }
'''
    exe = os.path.join(ROOT, "build", "sizeof_check")
    os.makedirs(os.path.dirname(exe), exist_ok=True)
    src = exe + ".c"
    with open(src, "w") as f:
        f.write(prog)
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), src, "-o", exe], check=True)
    g, a = map(int, subprocess.run([exe], capture_output=True, text=True, check=True).stdout.split())
    assert ctypes.sizeof(_lib.GemmDesc) == g and ctypes.sizeof(_lib.AttnDesc) == a


def test_argument_errors_reach_python_without_a_gpu(lib_path):
    """Argument validation happens before any CUDA call: a null descriptor returns CCEDIT_ERR_INVALID + message."""
    from ccedit_b200 import _lib
    lib = _lib.load()
    assert lib.ccedit_gemm(None, None) == 1
    assert b"null descriptor" in lib.ccedit_last_error()
    with pytest.raises(RuntimeError, match="null descriptor"):
        _lib.check(lib.ccedit_attention(None, None), "ccedit_attention")


def test_sass_uses_blackwell_tensor_and_tma_paths(lib_path):
    """The GEMM kernel is tcgen05 + TMA: UTC*MMA / LDTM / UTMALDG must appear in the sm_100a SASS."""
    out = subprocess.run(["cuobjdump", "-sass", lib_path], capture_output=True, text=True).stdout
    assert "sm_100a" in out or "SM100a" in out.replace("_", "") or "arch = sm_100" in out
    # tcgen05.mma / tcgen05.ld / tcgen05.st (P of the attention kernel lives in TMEM) / TMA load (2-D weights, 3-D K/V,
    # 5-D activations) / TMA store (staged GEMM epilogue)
    for mnemonic in ("UTCHMMA", "LDTM", "STTM", "UTMALDG.2D", "UTMALDG.3D", "UTMALDG.5D", "UTMASTG.5D"):
        assert mnemonic in out, f"{mnemonic} missing from SASS"
