"""Operator-level Python mirror of include/gvl.h: torch tensors in, C-ABI calls underneath.

torch is used for device memory and the current CUDA stream only; every computation below happens in
libgvl.so. Shape / dtype errors raise ValueError (the reference's own error class for these, e.g.
modeling_clip.py:276-320); library failures raise RuntimeError.
"""
import ctypes

import torch

from . import _lib

ACT_NONE, ACT_GELU_ERF, ACT_QUICK_GELU, ACT_SWIGLU = 0, 1, 2, 3
RES_NONE, RES_BF16, RES_F32 = 0, 1, 2


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _p(t):
    return ctypes.c_void_p(0 if t is None else t.data_ptr())


def _req(t, dtype, name):
    if not t.is_cuda:
        raise ValueError("%s must be a CUDA tensor (gvl has no CPU path)" % name)
    if t.dtype != dtype:
        raise ValueError("%s must be %s, got %s" % (name, dtype, t.dtype))


def gemm(a, w, bias=None, act=ACT_NONE, gamma=None, residual=None, out_dtype=torch.bfloat16, out=None, bn=0):
    """out = epilogue(a @ w.T); a [M,K] bf16 (row stride allowed), w [N,K] bf16."""
    lib = _lib.load()
    _req(a, torch.bfloat16, "a")
    _req(w, torch.bfloat16, "w")
    if a.dim() != 2 or w.dim() != 2 or a.shape[1] != w.shape[1]:
        raise ValueError("gemm shapes: a %s, w %s" % (tuple(a.shape), tuple(w.shape)))
    if a.stride(1) != 1 or w.stride(1) != 1:
        raise ValueError("gemm operands must be K-major (unit inner stride)")
    M, K = a.shape
    N = w.shape[0]
    n_out = N // 2 if act == ACT_SWIGLU else N
    if out is None:
        out = torch.empty((M, n_out), dtype=out_dtype, device=a.device)
    res_kind = RES_NONE
    ldr = 0
    if residual is not None:
        res_kind = RES_F32 if residual.dtype == torch.float32 else RES_BF16
        ldr = residual.stride(0)
    if bias is not None:
        _req(bias, torch.bfloat16, "bias")
    if gamma is not None:
        _req(gamma, torch.float32, "gamma")
    rc = lib.gvl_gemm_bf16(_p(a), a.stride(0), _p(w), w.stride(0), _p(out), out.stride(0), M, N, K, _p(bias),
                           _p(gamma), _p(residual), ldr, act, res_kind, 1 if out.dtype == torch.float32 else 0, bn,
                           _stream())
    _lib.check(rc, "gvl_gemm_bf16")
    return out


def _strides3(t_bs, t_ts, t_hs):
    arr = (ctypes.c_longlong * 3)(t_bs, t_ts, t_hs)
    return arr


def attention(q, k, v, scale, causal=False, round_scores=False, out=None, o_dim=0):
    """q [B,Sq,H,D], k/v [B,Skv,KVH,D] (any strides with unit stride on D) -> o [B,Sq,H,D] bf16."""
    lib = _lib.load()
    for t, n in ((q, "q"), (k, "k"), (v, "v")):
        _req(t, torch.bfloat16, n)
        if t.dim() != 4 or t.stride(3) != 1:
            raise ValueError("%s must be [B,S,H,D] with unit stride on D" % n)
    B, Sq, H, D = q.shape
    Skv, KVH = k.shape[1], k.shape[2]
    if out is None:
        out = torch.empty((B, Sq, H, o_dim if o_dim else D), dtype=torch.bfloat16, device=q.device)
    rc = lib.gvl_attention(_p(q), _p(k), _p(v), _p(out),
                           _strides3(q.stride(0), q.stride(1), q.stride(2)),
                           _strides3(k.stride(0), k.stride(1), k.stride(2)),
                           _strides3(v.stride(0), v.stride(1), v.stride(2)),
                           _strides3(out.stride(0), out.stride(1), out.stride(2)),
                           B, H, KVH, Sq, Skv, D, float(scale), int(causal), int(round_scores), int(o_dim), _stream())
    _lib.check(rc, "gvl_attention")
    return out


def layernorm(x, w, b, eps=1e-5):
    lib = _lib.load()
    _req(x, torch.float32, "x")
    rows = x.numel() // x.shape[-1]
    y = torch.empty(x.shape, dtype=torch.bfloat16, device=x.device)
    rc = lib.gvl_layernorm_f32(_p(x), _p(w), _p(b), _p(y), rows, x.shape[-1], float(eps), _stream())
    _lib.check(rc, "gvl_layernorm_f32")
    return y


def rmsnorm(x, w, eps):
    lib = _lib.load()
    _req(x, torch.bfloat16, "x")
    x2 = x.reshape(-1, x.shape[-1])
    y = torch.empty_like(x2)
    rc = lib.gvl_rmsnorm_bf16(_p(x2), x2.stride(0), _p(w), _p(y), y.stride(0), x2.shape[0], x2.shape[1], float(eps),
                              _stream())
    _lib.check(rc, "gvl_rmsnorm_bf16")
    return y.reshape(x.shape)


def iv2_qk_rmsnorm_(qkv, wq, wk, eps=1e-6):
    lib = _lib.load()
    _req(qkv, torch.bfloat16, "qkv")
    dim = qkv.shape[-1] // 3
    rows = qkv.numel() // qkv.shape[-1]
    rc = lib.gvl_iv2_qk_rmsnorm(_p(qkv), _p(wq), _p(wk), rows, dim, float(eps), _stream())
    _lib.check(rc, "gvl_iv2_qk_rmsnorm")
    return qkv


def im2col_patch14(pix, frames, kpad):
    """pix [N,C,(T,)H,W] -> [N*T*g*g, kpad] bf16."""
    lib = _lib.load()
    n, c = pix.shape[0], pix.shape[1]
    hw = pix.shape[-1]
    g = hw // 14
    out = torch.empty((n * frames * g * g, kpad), dtype=torch.bfloat16, device=pix.device)
    rc = lib.gvl_im2col_patch14(_p(pix), 1 if pix.dtype == torch.float32 else 0, _p(out), n, c, frames, hw, kpad,
                                _stream())
    _lib.check(rc, "gvl_im2col_patch14")
    return out


def clip_assemble(patch, cls, pos, n_img):
    lib = _lib.load()
    n_patch = patch.shape[0] // n_img
    dim = patch.shape[1]
    x = torch.empty((n_img, n_patch + 1, dim), dtype=torch.float32, device=patch.device)
    rc = lib.gvl_clip_assemble(_p(patch), _p(cls), _p(pos), _p(x), n_img, n_patch, dim, _stream())
    _lib.check(rc, "gvl_clip_assemble")
    return x


def iv2_assemble(patch, cls, pos, n_seg):
    lib = _lib.load()
    n_patch = patch.shape[0] // n_seg
    dim = patch.shape[1]
    x = torch.empty((n_seg, n_patch + 1, dim), dtype=torch.bfloat16, device=patch.device)
    rc = lib.gvl_iv2_assemble(_p(patch), _p(cls), _p(pos), _p(x), n_seg, n_patch, dim, _stream())
    _lib.check(rc, "gvl_iv2_assemble")
    return x


def hd_merge_newline(hs, sub_gn):
    """hs fp32 [N,577,1024] -> bf16 [N,156,4096] (llava_next_video.py:454-489)."""
    lib = _lib.load()
    _req(hs, torch.float32, "hs")
    if hs.shape[1:] != (577, 1024):
        raise ValueError("hd_merge expects [N,577,1024]")  # reference asserts L==576, C==1024 (:460)
    out = torch.empty((hs.shape[0], 156, 4096), dtype=torch.bfloat16, device=hs.device)
    rc = lib.gvl_hd_merge_newline(_p(hs), _p(sub_gn), _p(out), hs.shape[0], _stream())
    _lib.check(rc, "gvl_hd_merge_newline")
    return out


def iv2_pool(x, frames):
    lib = _lib.load()
    _req(x, torch.bfloat16, "x")
    n, _, dim = x.shape
    out = torch.empty((n, frames * 16, dim), dtype=torch.bfloat16, device=x.device)
    rc = lib.gvl_iv2_pool(_p(x), _p(out), n, frames, dim, _stream())
    _lib.check(rc, "gvl_iv2_pool")
    return out


def clip_pool3(hs):
    lib = _lib.load()
    _req(hs, torch.float32, "hs")
    out = torch.empty((hs.shape[0], 64, 1024), dtype=torch.bfloat16, device=hs.device)
    rc = lib.gvl_clip_pool3(_p(hs), _p(out), hs.shape[0], _stream())
    _lib.check(rc, "gvl_clip_pool3")
    return out


def embed_splice(ids, img_pos, table, visual, vis_last=False):
    lib = _lib.load()
    _req(ids, torch.int64, "ids")
    t_text = ids.shape[0]
    n_vis, dim = visual.shape
    out = torch.empty((t_text - 1 + n_vis, dim), dtype=torch.bfloat16, device=table.device)
    rc = lib.gvl_embed_splice(_p(ids), t_text, int(img_pos), _p(table), _p(visual), n_vis, _p(out), dim,
                              int(vis_last), _stream())
    _lib.check(rc, "gvl_embed_splice")
    return out


def rope_qkv_cache(qkv, k_cache, v_cache, cos, sin, heads, kv_heads, head_dim, pos0, positions=None):
    lib = _lib.load()
    tokens = qkv.shape[0]
    q_out = torch.empty((tokens, heads * head_dim), dtype=torch.bfloat16, device=qkv.device)
    max_ctx = k_cache.shape[1]
    rc = lib.gvl_rope_qkv_cache(_p(qkv), _p(q_out), _p(k_cache), _p(v_cache), _p(cos), _p(sin), _p(positions),
                                tokens, heads, kv_heads, head_dim, pos0, max_ctx, _stream())
    _lib.check(rc, "gvl_rope_qkv_cache")
    return q_out


def gemv(x, w, norm_w=None, norm_eps=1e-5, bias=None, residual=None, act=ACT_NONE, out_dtype=torch.bfloat16):
    lib = _lib.load()
    _req(x, torch.bfloat16, "x")
    M, K = x.shape
    N = w.shape[0]
    n_out = N // 2 if act == ACT_SWIGLU else N
    out = torch.empty((M, n_out), dtype=out_dtype, device=x.device)
    rc = lib.gvl_gemv_bf16(_p(x), x.stride(0), _p(w), w.stride(0), _p(out), out.stride(0), M, N, K, _p(norm_w),
                           float(norm_eps), _p(bias), _p(residual), 0 if residual is None else residual.stride(0), act,
                           1 if out_dtype == torch.float32 else 0, _stream())
    _lib.check(rc, "gvl_gemv_bf16")
    return out


def decode_attention(q, k_cache, v_cache, ctx_len_dev, scale):
    """q [H*D] bf16, caches [KVH, max_ctx, D]; ctx_len_dev int32[1] on device."""
    lib = _lib.load()
    kvh, max_ctx, d = k_cache.shape
    heads = q.numel() // d
    ws_bytes = lib.gvl_decode_attention_workspace(heads, d, max_ctx)
    ws = torch.zeros((ws_bytes // 4,), dtype=torch.float32, device=q.device)   # arrival counters must start at 0
    o = torch.empty_like(q)
    rc = lib.gvl_decode_attention(_p(q), _p(k_cache), _p(v_cache), _p(o), _p(ws), _p(ctx_len_dev), heads, kvh, d,
                                  max_ctx, float(scale), _stream())
    _lib.check(rc, "gvl_decode_attention")
    return o


def argmax(logits):
    lib = _lib.load()
    _req(logits, torch.float32, "logits")
    out = torch.empty((1,), dtype=torch.int64, device=logits.device)
    rc = lib.gvl_argmax_f32(_p(logits), logits.numel(), _p(out), _stream())
    _lib.check(rc, "gvl_argmax_f32")
    return out


def launch_count():
    return int(_lib.load().gvl_launch_count())
