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


def test_decode_attention_split_covers_every_token_once_and_fits_the_workspace():
    """Host-side arithmetic of the single-kernel decode step's attention phase (csrc/decode_mega.cu att_split, through the C ABI, no
    GPU): for every (context, heads, grid) every token of every head belongs to exactly one warp, ranges are 8-aligned above the
    one-warp-per-head regime, the split is balanced, and no head spreads over more CTAs than the merge workspace has records for."""
    import ctypes
    from gvl import _lib
    lib = _lib.load()
    out = [ctypes.c_int() for _ in range(4)]

    def split(ctx, H, G):
        rc = lib.gvl_lm_attention_split(ctx, H, G, *[ctypes.byref(o) for o in out])
        assert rc == 0, (rc, ctx, H, G)
        return [o.value for o in out]

    ctxs = list(range(1, 300)) + [383, 384, 385, 1000, 2379, 2380, 3483, 3484, 3485, 4095, 4096, 4097, 7679, 7680, 8191, 8192, 16384]
    for H in (1, 2, 4, 8, 12, 16, 24, 32, 40, 64):
        for G in (8, 16, 60, 108, 132, 148, 160):
            for ctx in ctxs:
                wph, lw, mp, cap = split(ctx, H, G)
                assert wph >= 1 and wph * H <= G * 8
                assert wph * lw >= ctx                                   # the last warp of a head reaches the end of the context
                assert mp <= cap, (ctx, H, G, mp, cap)                   # partial records per head fit MegaPlan::att_maxp
                if ctx <= 128:
                    assert (wph, lw) == (1, ctx)
                else:
                    assert wph == (G * 8) // H and lw % 8 == 0 and lw - (ctx + wph - 1) // wph < 8
                # the warp that owns the newest token (position ctx - 1) exists and owns it alone
                owner = (ctx - 1) // lw
                assert owner < wph and owner * lw <= ctx - 1 < min(ctx, (owner + 1) * lw)
    # argument errors come back as status codes
    for bad in ((0, 32, 148), (10, 0, 148), (10, 65, 148), (10, 32, 0), (10, 32, 3)):
        assert lib.gvl_lm_attention_split(*bad, *[ctypes.byref(o) for o in out]) == _lib.GVL_ERR_ARG
    # the production shapes: 37 warps per head, 96 tokens per warp, <= 6 partial records per head
    assert split(3484, 32, 148) == [37, 96, 6, 6]
