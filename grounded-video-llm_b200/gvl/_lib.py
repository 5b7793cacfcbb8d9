"""ctypes loader for libgvl.so (the C ABI in include/gvl.h).

There is NO CPU fallback: if the shared library is missing or a call fails, we raise. The oracle under
/oracle is test infrastructure and is never imported from here.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libgvl.so")

STATUS = {
    0: "GVL_OK",
    -1: "GVL_ERR_ARG (bad shape / unsupported combination)",
    -2: "GVL_ERR_ALIGN (pointer or leading dimension not 16-byte aligned)",
    -3: "GVL_ERR_CUDA (kernel launch failed)",
    -4: "GVL_ERR_DRIVER (cuTensorMapEncodeTiled unavailable or failed)",
    -5: "GVL_ERR_NOMEM",
    -6: "GVL_ERR_STATE",
}

GVL_OK, GVL_ERR_ARG, GVL_ERR_ALIGN, GVL_ERR_CUDA, GVL_ERR_DRIVER, GVL_ERR_NOMEM, GVL_ERR_STATE = 0, -1, -2, -3, -4, -5, -6

_lib = None

c_vp = ctypes.c_void_p
c_i = ctypes.c_int
c_ll = ctypes.c_longlong
c_f = ctypes.c_float
c_sz = ctypes.c_size_t


class ClipLayer(ctypes.Structure):
    _fields_ = [(n, c_vp) for n in (
        "ln1_w", "ln1_b", "qkv_w", "qkv_b", "out_w", "out_b", "ln2_w", "ln2_b",
        "fc1_w", "fc1_b", "fc2_w", "fc2_b")]


class ClipWeights(ctypes.Structure):
    _fields_ = [("n_layers", c_i), ("dim", c_i), ("heads", c_i), ("ffn", c_i), ("n_patch", c_i),
                ("image", c_i), ("kpad", c_i),
                ("patch_w", c_vp), ("cls", c_vp), ("pos", c_vp), ("pre_ln_w", c_vp), ("pre_ln_b", c_vp),
                ("layers", ctypes.POINTER(ClipLayer))]


class Iv2Block(ctypes.Structure):
    _fields_ = [(n, c_vp) for n in (
        "norm1_w", "qkv_w", "q_norm_w", "k_norm_w", "proj_w", "proj_b", "ls1", "norm2_w",
        "fc1_w", "fc1_b", "fc2_w", "fc2_b", "ls2")]


class Iv2Weights(ctypes.Structure):
    _fields_ = [("n_blocks", c_i), ("dim", c_i), ("heads", c_i), ("ffn", c_i), ("frames", c_i), ("kpad", c_i),
                ("head_dim_pad", c_i),
                ("patch_w", c_vp), ("patch_b", c_vp), ("cls", c_vp), ("pos", c_vp),
                ("blocks", ctypes.POINTER(Iv2Block))]


class LmLayer(ctypes.Structure):
    _fields_ = [(n, c_vp) for n in ("in_norm_w", "qkv_w", "o_w", "post_norm_w", "gate_up_w", "down_w")]


class LmWeights(ctypes.Structure):
    _fields_ = [("n_layers", c_i), ("dim", c_i), ("heads", c_i), ("kv_heads", c_i), ("head_dim", c_i),
                ("ffn", c_i), ("vocab", c_i), ("max_ctx", c_i), ("rms_eps", c_f),
                ("final_norm_w", c_vp), ("lm_head_w", c_vp), ("lm_head_b", c_vp), ("embed", c_vp),
                ("rope_cos", c_vp), ("rope_sin", c_vp),
                ("layers", ctypes.POINTER(LmLayer))]


_SIGS = {
    "gvl_version": (ctypes.c_char_p, []),
    "gvl_launch_count": (c_ll, []),
    "gvl_gemm_bf16": (c_i, [c_vp, c_i, c_vp, c_i, c_vp, c_i, c_i, c_i, c_i, c_vp, c_vp, c_vp, c_i, c_i, c_i, c_i,
                            c_i, c_vp]),
    "gvl_attention": (c_i, [c_vp, c_vp, c_vp, c_vp, ctypes.POINTER(c_ll), ctypes.POINTER(c_ll),
                            ctypes.POINTER(c_ll), ctypes.POINTER(c_ll), c_i, c_i, c_i, c_i, c_i, c_i, c_f, c_i, c_i,
                            c_i, c_vp]),
    "gvl_layernorm_f32": (c_i, [c_vp, c_vp, c_vp, c_vp, c_i, c_i, c_f, c_vp]),
    "gvl_rmsnorm_bf16": (c_i, [c_vp, c_ll, c_vp, c_vp, c_ll, c_i, c_i, c_f, c_vp]),
    "gvl_iv2_qk_rmsnorm": (c_i, [c_vp, c_vp, c_vp, c_i, c_i, c_f, c_vp]),
    "gvl_im2col_patch14": (c_i, [c_vp, c_i, c_vp, c_i, c_i, c_i, c_i, c_i, c_vp]),
    "gvl_clip_assemble": (c_i, [c_vp, c_vp, c_vp, c_vp, c_i, c_i, c_i, c_vp]),
    "gvl_iv2_assemble": (c_i, [c_vp, c_vp, c_vp, c_vp, c_i, c_i, c_i, c_vp]),
    "gvl_hd_merge_newline": (c_i, [c_vp, c_vp, c_vp, c_i, c_vp]),
    "gvl_iv2_pool": (c_i, [c_vp, c_vp, c_i, c_i, c_i, c_vp]),
    "gvl_clip_pool3": (c_i, [c_vp, c_vp, c_i, c_vp]),
    "gvl_embed_splice": (c_i, [c_vp, c_i, c_i, c_vp, c_vp, c_i, c_vp, c_i, c_i, c_vp]),
    "gvl_rope_qkv_cache": (c_i, [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_i, c_i, c_i, c_i, c_i, c_i, c_vp]),
    "gvl_visual_concat": (c_i, [c_vp, c_i, c_vp, c_i, c_vp, c_vp, c_i, c_i, c_vp]),
    "gvl_layernorm_f32_out_f32": (c_i, [c_vp, c_vp, c_vp, c_vp, c_i, c_i, c_f, c_vp]),
    "gvl_gemv_bf16": (c_i, [c_vp, c_i, c_vp, c_i, c_vp, c_i, c_i, c_i, c_i, c_vp, c_f, c_vp, c_vp, c_i, c_i, c_i,
                            c_vp]),
    "gvl_decode_attention": (c_i, [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_i, c_i, c_i, c_i, c_f, c_vp]),
    "gvl_decode_attention_workspace": (c_sz, [c_i, c_i, c_i]),
    "gvl_argmax_f32": (c_i, [c_vp, c_i, c_vp, c_vp]),
    "gvl_clip_workspace": (c_sz, [ctypes.POINTER(ClipWeights), c_i]),
    "gvl_clip_encode": (c_i, [ctypes.POINTER(ClipWeights), c_vp, c_vp, c_i, c_vp, c_sz, c_vp]),
    "gvl_iv2_workspace": (c_sz, [ctypes.POINTER(Iv2Weights), c_i]),
    "gvl_iv2_encode": (c_i, [ctypes.POINTER(Iv2Weights), c_vp, c_vp, c_i, c_vp, c_sz, c_vp]),
    "gvl_lm_create": (c_i, [ctypes.POINTER(LmWeights), ctypes.POINTER(c_vp)]),
    "gvl_lm_create_ex": (c_i, [ctypes.POINTER(LmWeights), c_i, ctypes.POINTER(c_vp)]),
    "gvl_lm_destroy": (None, [c_vp]),
    "gvl_lm_prefill": (c_i, [c_vp, c_vp, c_i, c_vp, c_vp, c_vp]),
    "gvl_lm_decode": (c_i, [c_vp, c_i, c_vp, c_vp, c_ll, c_ll, c_vp]),
    "gvl_lm_first_token": (c_vp, [c_vp]),
    "gvl_lm_decode_kind": (c_i, [c_vp]),
    "gvl_lm_attention_split": (c_i, [c_i, c_i, c_i] + [ctypes.POINTER(c_i)] * 4),
    "gvl_lm_decode_batch": (c_i, [ctypes.POINTER(c_vp), c_i, c_i, c_vp, c_vp, c_ll, c_ll, c_vp]),
    "gvl_lm_set_next_token": (c_i, [c_vp, c_vp, c_vp]),
    "gvl_lm_set_graph": (c_i, [c_vp, c_i]),
    "gvl_lm_mega_trace": (c_i, [c_vp, c_vp, c_i, ctypes.POINTER(c_i), ctypes.POINTER(c_i)]),
    "gvl_frame_transform_workspace": (ctypes.c_size_t, [c_i, c_i, c_i, c_i, c_i]),
    "gvl_frame_transform": (c_i, [c_vp, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_i, ctypes.POINTER(ctypes.c_float),
                                  ctypes.POINTER(ctypes.c_float), c_vp, c_vp, ctypes.c_size_t, c_vp]),
    "gvl_profile_enable": (c_i, [c_i]),
    "gvl_profile_collect": (c_i, [c_i, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_double),
                                  ctypes.POINTER(c_ll)]),
}

EXPORTS = tuple(_SIGS.keys())


def lib_path():
    return _LIB_PATH


def load():
    """Load libgvl.so; raise (never fall back) if it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_LIB_PATH):
        raise RuntimeError(
            "libgvl.so not found at %s -- build it with `python grounded-video-llm_b200/build.py` "
            "(there is no CPU fallback for the gvl hot path)" % _LIB_PATH)
    lib = ctypes.CDLL(_LIB_PATH)
    missing = []
    for name, (res, args) in _SIGS.items():
        fn = getattr(lib, name, None)
        if fn is None:
            missing.append(name)
            continue
        fn.restype = res
        fn.argtypes = args
    if missing and not os.environ.get("GVL_ALLOW_MISSING"):  # bring-up only; never set in tests/bench
        raise RuntimeError("libgvl.so does not export %s (ABI drift vs include/gvl.h)" % ", ".join(missing))
    _lib = lib
    return lib


def check(rc, what):
    if rc != 0:
        raise RuntimeError("%s failed: %s" % (what, STATUS.get(rc, "status %d" % rc)))
