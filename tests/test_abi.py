"""The C-ABI shared library loads on a CPU-only box and exports every symbol include/gvl.h declares."""
import ctypes
import os
import re

from gvl import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "gvl.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(gvl_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported_and_mirrored():
    names = _declared()
    assert len(names) >= 30
    lib = ctypes.CDLL(_lib.lib_path())
    for n in names:
        assert hasattr(lib, n), "libgvl.so does not export %s" % n
    assert set(names) == set(_lib.EXPORTS), set(names) ^ set(_lib.EXPORTS)


def test_loader_is_strict_and_versioned():
    lib = _lib.load()
    assert b"sm_100a" in lib.gvl_version()
    assert lib.gvl_launch_count() == 0
    # argument validation happens before any CUDA call, so it is checkable without a GPU
    assert lib.gvl_gemm_bf16(None, 0, None, 0, None, 0, 1, 8, 8, None, None, None, 0, 0, 0, 0, 0, None) == -1
    # [heads][splits of 128 tokens][D+2] fp32 partials + [heads] int32 arrival counters
    assert lib.gvl_decode_attention_workspace(32, 96, 4096) == 32 * 32 * 98 * 4 + 32 * 4


def test_no_oracle_import_in_product():
    pkg = os.path.join(ROOT, "grounded-video-llm_b200", "gvl")
    for f in os.listdir(pkg):
        if f.endswith(".py"):
            assert "oracle" not in open(os.path.join(pkg, f)).read().replace("The oracle under\n/oracle", "").replace("oracle is test", ""), f
