"""ORACLE -- TEST INFRASTRUCTURE ONLY.  Never imported by the product (grounded-video-llm_b200/gvl).

A plain-torch restatement of the Grounded-VideoLLM forward path (reference WHB139426/Grounded-Video-LLM
@ e26da4e), written against the reference's own files; every function cites the file:line it follows.
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may use it.

Pinning (see oracle/make_golden.py, tests/test_oracle_golden.py): every function here is checked in
fp32 mode against the IMPORTED reference modules (tiny random-init configs, committed as
tests/golden/*.npz) and against the README known-answer timestamps (README.md:90-94).

Two arithmetic modes:
  mode="fp32"  every op in fp32 -- equals the reference's CPU fp32 path (BASELINE config 1).
  mode="bf16"  mirrors the reference's CUDA `autocast(bf16)` forward: tensors that are bf16 in the
               reference are kept as fp32 storage holding bf16-representable values; every matmul takes
               bf16-rounded operands, accumulates in fp32 and is rounded once (tensor-core semantics);
               the rounding points are the ones listed in SURVEY.md 8a "numerics contract".
Tensors are fp32 storage throughout, so the functions run on CPU (default) or on any torch device.
"""
import math
import re

import torch
import torch.nn.functional as F

IMAGE_TOKEN_INDEX = -200            # datasets/chat/base_template.py (constant used at llava_next_video.py:579)
GROUNDING_TOKEN = "<timestamp_grounding>"
DEFAULT_IMAGE_TOKEN = "<image>"


# ----------------------------------------------------------------------------- rounding helpers
def bf(x):
    """Round to bf16 (RNE) and return as fp32 storage."""
    return x.to(torch.bfloat16).to(torch.float32)


def _r(x, mode):
    return bf(x) if mode == "bf16" else x


def linear(x, w, b, mode):
    """nn.Linear under autocast: bf16 operands, fp32 accumulate (+bias), one rounding."""
    if mode == "bf16":
        y = bf(x) @ bf(w).t()
        if b is not None:
            y = y + bf(b)
        return bf(y)
    y = x @ w.t()
    return y if b is None else y + b


def gelu_erf(x, mode):
    return _r(F.gelu(x), mode)


def quick_gelu(x, mode):
    # HF QuickGELUActivation: input * sigmoid(1.702 * input) -- three bf16 ops on a bf16 tensor
    if mode == "bf16":
        t = bf(1.702 * x)
        return bf(x * bf(torch.sigmoid(t)))
    return x * torch.sigmoid(1.702 * x)


def rmsnorm(x, w, eps, mode):
    """internvideo2.py:437-448 / modeling_phi3.py:310-324 / modeling_llama.py:74-88."""
    xf = x.float()
    var = xf.pow(2).mean(-1, keepdim=True)
    y = xf * torch.rsqrt(var + eps)
    if mode == "bf16":
        return bf(bf(w) * bf(y))
    return w * y


def layernorm(x, w, b, eps=1e-5):
    """nn.LayerNorm; autocast runs it in fp32 (modeling_clip.py:351-353)."""
    return F.layer_norm(x.float(), (x.shape[-1],), w.float(), b.float(), eps)


def attention_core(q, k, v, scale, causal, mode, style="flash"):
    """q,k,v: [B,H,S,D].
    style="flash": flash_attn semantics (internvideo2.py:514, modeling_phi3.py:857): fp32 scores, fp32 softmax
                   statistics, unnormalised probabilities rounded to bf16 before P@V, fp32 accumulate.
    style="eager": CLIPAttention (modeling_clip.py:274-314): bmm rounded to bf16, softmax in fp32,
                   normalised probabilities rounded to bf16, bmm rounded.
    """
    s = q @ k.transpose(-1, -2)
    Sq, Sk = q.shape[-2], k.shape[-2]
    if mode == "bf16" and style == "eager":
        s = bf(s)
    s = s * scale
    if causal:
        i = torch.arange(Sq, device=q.device)[:, None]
        j = torch.arange(Sk, device=q.device)[None, :]
        s = s.masked_fill(j > i + (Sk - Sq), float("-inf"))
    if mode != "bf16":
        return torch.softmax(s, dim=-1) @ v
    if style == "eager":
        p = bf(torch.softmax(s, dim=-1))
        return bf(p @ v)
    m = s.max(dim=-1, keepdim=True).values
    p = torch.exp(s - m)
    l = p.sum(dim=-1, keepdim=True)
    return bf((bf(p) @ v) / l)


# ----------------------------------------------------------------------------- CLIP ViT (spatial stream)
def clip_patch_embed(pix, P, mode):
    """CLIPVisionEmbeddings.forward (modeling_clip.py:182-191)."""
    w = P["vision_model.embeddings.patch_embedding.weight"]        # [D,3,14,14], no bias
    D = w.shape[0]
    ps = w.shape[-1]
    cols = F.unfold(_r(pix.float(), mode), kernel_size=ps, stride=ps).transpose(1, 2)   # [N, L, 3*ps*ps]
    patch = cols @ _r(w.reshape(D, -1), mode).t()
    patch = _r(patch, mode)
    cls = P["vision_model.embeddings.class_embedding"].float().expand(pix.shape[0], 1, D)
    emb = torch.cat([cls, patch], dim=1)                            # promoted to fp32
    return emb + P["vision_model.embeddings.position_embedding.weight"].float()[None]


def clip_layer(x, P, pre, heads, mode):
    """CLIPEncoderLayer.forward (modeling_clip.py:355-393) with CLIPAttention (:252-328), CLIPMLP (:339-343)."""
    D = x.shape[-1]
    hd = D // heads
    scale = hd ** -0.5
    h = layernorm(x, P[pre + "layer_norm1.weight"], P[pre + "layer_norm1.bias"])
    q = linear(h, P[pre + "self_attn.q_proj.weight"], P[pre + "self_attn.q_proj.bias"], mode)
    q = _r(q * scale, mode)
    k = linear(h, P[pre + "self_attn.k_proj.weight"], P[pre + "self_attn.k_proj.bias"], mode)
    v = linear(h, P[pre + "self_attn.v_proj.weight"], P[pre + "self_attn.v_proj.bias"], mode)
    B, S, _ = x.shape

    def sh(t):
        return t.view(B, S, heads, hd).transpose(1, 2)
    o = attention_core(sh(q), sh(k), sh(v), 1.0, False, mode, style="eager")
    o = o.transpose(1, 2).reshape(B, S, D)
    x = x + linear(o, P[pre + "self_attn.out_proj.weight"], P[pre + "self_attn.out_proj.bias"], mode)
    h = layernorm(x, P[pre + "layer_norm2.weight"], P[pre + "layer_norm2.bias"])
    f = quick_gelu(linear(h, P[pre + "mlp.fc1.weight"], P[pre + "mlp.fc1.bias"], mode), mode)
    return x + linear(f, P[pre + "mlp.fc2.weight"], P[pre + "mlp.fc2.bias"], mode)


def clip_hidden_states(pix, P, heads, n_layers, mode="bf16", upto=None):
    """CLIPVisionTransformer.forward + CLIPEncoder.forward (modeling_clip.py:830-872, 578-657).
    Returns the hidden_states tuple as a list; the consumer reads [-2] (llava_next_video.py:505)."""
    x = clip_patch_embed(pix, P, mode)
    x = layernorm(x, P["vision_model.pre_layrnorm.weight"], P["vision_model.pre_layrnorm.bias"])
    hs = [x]
    n_run = n_layers if upto is None else upto
    for l in range(n_run):
        x = clip_layer(x, P, "vision_model.encoder.layers.%d." % l, heads, mode)
        hs.append(x)
    return hs


# ----------------------------------------------------------------------------- InternVideo2 (temporal stream)
def iv2_forward(pix, P, heads, depth, mode="bf16", x_vis_return_idx=-2, style="flash"):
    """PretrainInternVideo2.forward(x, None, False, x_vis_return_idx, x_vis_only=True)
    (internvideo2.py:970-1040), blocks = Block._inner_forward (:680-684), Attention (:564-605), Mlp (:630-636),
    LayerScale (:451-466), PatchEmbed (:721-725). pix: [N,3,T,H,W]."""
    w = P["patch_embed.proj.weight"]                                # [D,3,1,14,14]
    D = w.shape[0]
    N, C, T, Hh, Ww = pix.shape
    ps = w.shape[-1]
    fr = pix.float().permute(0, 2, 1, 3, 4).reshape(N * T, C, Hh, Ww)
    cols = F.unfold(_r(fr, mode), kernel_size=ps, stride=ps).transpose(1, 2)            # [N*T, L, C*ps*ps]
    patch = cols @ _r(w.reshape(D, -1), mode).t() + _r(P["patch_embed.proj.bias"].float(), mode)
    patch = _r(patch, mode).reshape(N, T * cols.shape[1], D)
    cls = _r(P["cls_token"].float(), mode).expand(N, -1, -1)
    x = torch.cat([cls, patch], dim=1)
    x = _r(x + _r(P["pos_embed"].float(), mode), mode)
    hd = D // heads
    scale = hd ** -0.5
    last = depth + x_vis_return_idx
    for i in range(depth):
        pre = "blocks.%d." % i
        h = rmsnorm(x, P[pre + "norm1.weight"].float(), 1e-6, mode)
        qkv = linear(h, P[pre + "attn.qkv.weight"].float(), None, mode)
        B, S, _ = qkv.shape
        qkv = qkv.view(B, S, 3, heads, hd)
        q, k, v = qkv.unbind(2)
        q = rmsnorm(q.flatten(-2), P[pre + "attn.q_norm.weight"].float(), 1e-6, mode).view(B, S, heads, hd)
        k = rmsnorm(k.flatten(-2), P[pre + "attn.k_norm.weight"].float(), 1e-6, mode).view(B, S, heads, hd)
        if mode == "bf16" and style == "naive":
            # _naive_attn (:564-583): (q*scale) rounded, scores rounded, softmax in bf16-out
            o = attention_core(_r(q.transpose(1, 2) * scale, mode), k.transpose(1, 2), v.transpose(1, 2), 1.0, False,
                               mode, style="eager")
        else:
            o = attention_core(q.transpose(1, 2), k.transpose(1, 2), v.transpose(1, 2), scale, False, mode)
        o = o.transpose(1, 2).reshape(B, S, D)
        a = linear(o, P[pre + "attn.proj.weight"].float(), P[pre + "attn.proj.bias"].float(), mode)
        x = _r(x + _r(a * _r(P[pre + "ls1.gamma"].float(), mode), mode), mode)
        h = rmsnorm(x, P[pre + "norm2.weight"].float(), 1e-6, mode)
        f = gelu_erf(linear(h, P[pre + "mlp.fc1.weight"].float(), P[pre + "mlp.fc1.bias"].float(), mode), mode)
        m = linear(f, P[pre + "mlp.fc2.weight"].float(), P[pre + "mlp.fc2.bias"].float(), mode)
        x = _r(x + _r(m * _r(P[pre + "ls2.gamma"].float(), mode), mode), mode)
        if i == last:
            break
    return x


# ----------------------------------------------------------------------------- index maps / projectors
def hd_merge_newline(image_features, sub_gn):
    """reshape_hd_patches_2x2merge_phi3(h_crop=w_crop=1) + add_image_newline_phi3 (llava_next_video.py:454-489).
    image_features [N,576,C] -> [N,156,4C]."""
    N, L, C = image_features.shape
    H = int(L ** 0.5)
    x = image_features.reshape(N, H // 2, 2, H // 2, 2, C).permute(0, 1, 3, 2, 4, 5).reshape(N, H // 2, H // 2, 4 * C)
    nl = sub_gn.reshape(1, 1, 1, 4 * C).expand(N, H // 2, 1, 4 * C)
    return torch.cat([x, nl.to(x.dtype)], dim=2).reshape(N, -1, 4 * C)


def pool_temporal(x_vis, frames):
    """llava_next_video.py:532-549: drop cls, AdaptiveAvgPool3d([T,4,4]) on the 16x16 grid -> [N, T*16, D]."""
    N, _, D = x_vis.shape
    g = int(math.isqrt((x_vis.shape[1] - 1) // frames))
    t = x_vis[:, 1:].reshape(N, frames, g, g, D).permute(0, 4, 1, 2, 3)
    p = F.adaptive_avg_pool3d(t, (frames, 4, 4))
    return p.permute(0, 2, 3, 4, 1).reshape(N, frames * 16, D)


def pool_spatial_llama(image_features):
    """llava_next_video.py:509-517: AdaptiveAvgPool3d([segs,8,8]) == exact 3x3 means on the 24x24 grid."""
    N, L, C = image_features.shape
    g = int(L ** 0.5)
    t = image_features.reshape(N, g, g, C).permute(0, 3, 1, 2)
    return F.adaptive_avg_pool2d(t, (8, 8)).permute(0, 2, 3, 1).reshape(N, 64, C)


def mlp2(x, w0, b0, w1, b1, mode):
    """Phi3_5_Projecter / Video_Projecter / LlavaMultiModalProjector (llava_next_video.py:26-54): Linear-GELU-Linear."""
    return linear(gelu_erf(linear(x, w0, b0, mode), mode), w1, b1, mode)


def encode_images_phi(spatial, temporal, P, cfg, mode="bf16"):
    """LLAVA_NEXT_VIDEO.encode_images, phi3.5 branch (llava_next_video.py:491-566).
    spatial [B,segs,3,336,336], temporal [B,frames,3,224,224] -> [B, segs*(156+16*fps+1), D]."""
    B, segs = spatial.shape[:2]
    frames = temporal.shape[1]
    fps = frames // segs
    hs = clip_hidden_states(spatial.flatten(0, 1), P["clip"], cfg["clip_heads"], cfg["clip_layers"], mode,
                            upto=cfg["clip_layers"] - 1)
    feat = hs[-1][:, 1:]                                            # hidden_states[-2][:, 1:]
    feat = hd_merge_newline(feat, P["sub_GN"].float())
    sp = mlp2(feat, P["mm.linear_0.weight"], P["mm.linear_0.bias"], P["mm.linear_1.weight"], P["mm.linear_1.bias"], mode)
    tp = temporal.reshape(B, segs, fps, *temporal.shape[2:]).permute(0, 1, 3, 2, 4, 5).flatten(0, 1)
    xv = iv2_forward(tp, P["iv2"], cfg["iv2_heads"], cfg["iv2_depth"], mode)
    pooled = _r(pool_temporal(xv, fps), mode)
    tm = mlp2(pooled, P["vp.up_proj.weight"], P["vp.up_proj.bias"], P["vp.down_proj.weight"], P["vp.down_proj.bias"], mode)
    nl = mlp2(P["glb_GN"].float().reshape(1, -1), P["mm.linear_0.weight"], P["mm.linear_0.bias"],
              P["mm.linear_1.weight"], P["mm.linear_1.bias"], mode)
    Dm = sp.shape[-1]
    vid = torch.cat([sp.reshape(B, segs, -1, Dm), tm.reshape(B, segs, -1, Dm),
                     nl.reshape(1, 1, 1, Dm).expand(B, segs, 1, Dm)], dim=2)
    return vid.reshape(B, -1, Dm)


def encode_images_llama(spatial, temporal, P, cfg, mode="bf16"):
    """LLAVA_NEXT_VIDEO.encode_images, llama3 / vicuna branch (llava_next_video.py:507-518, 530-566):
    3x3 adaptive pooling of the CLIP grid -> LlavaMultiModalProjector (linear_1, GELU, linear_2), temporal stream as for
    phi3.5, newline = the learned `image_newline` vector. -> [B, segs*(64+16*fps+1), D]."""
    B, segs = spatial.shape[:2]
    frames = temporal.shape[1]
    fps = frames // segs
    hs = clip_hidden_states(spatial.flatten(0, 1), P["clip"], cfg["clip_heads"], cfg["clip_layers"], mode,
                            upto=cfg["clip_layers"] - 1)
    feat = pool_spatial_llama(hs[-1][:, 1:])                        # fp32 means of the fp32 hidden state
    sp = mlp2(feat, P["mm.linear_1.weight"], P["mm.linear_1.bias"], P["mm.linear_2.weight"], P["mm.linear_2.bias"], mode)
    tp = temporal.reshape(B, segs, fps, *temporal.shape[2:]).permute(0, 1, 3, 2, 4, 5).flatten(0, 1)
    xv = iv2_forward(tp, P["iv2"], cfg["iv2_heads"], cfg["iv2_depth"], mode)
    pooled = _r(pool_temporal(xv, fps), mode)
    tm = mlp2(pooled, P["vp.up_proj.weight"], P["vp.up_proj.bias"], P["vp.down_proj.weight"], P["vp.down_proj.bias"], mode)
    Dm = sp.shape[-1]
    nl = _r(P["image_newline"].float(), mode).reshape(1, 1, 1, Dm).expand(B, segs, 1, Dm)
    vid = torch.cat([sp.reshape(B, segs, -1, Dm), tm.reshape(B, segs, -1, Dm), nl], dim=2)
    return vid.reshape(B, -1, Dm)


def splice_embeds(ids, embed_table, visual, vis_last=False):
    """prepare_multimodal_inputs for one sample (llava_next_video.py:568-596)."""
    pos = int((ids == IMAGE_TOKEN_INDEX).nonzero()[0])
    pre = embed_table[ids[:pos]]
    post = embed_table[ids[pos + 1:]]
    if vis_last:
        return torch.cat([pre, post, visual], dim=0)
    return torch.cat([pre, visual, post], dim=0)


# ----------------------------------------------------------------------------- LLM
def phi3_rope_tables(positions, head_dim, base, short_factor, long_factor, max_pos, orig_max_pos, seq_len=None,
                     long_from=None):
    """Phi3LongRoPEScaledRotaryEmbedding.forward (modeling_phi3.py:371-409). Returns fp32 cos, sin [S, head_dim]
    BEFORE the cast to the activation dtype (the caller rounds to bf16 in bf16 mode).
    One forward call uses ONE factor set, picked from seq_len (= kv_seq_len at the call sites, :562-563, 680-686).
    `long_from` restates the KV-CACHED generate path in a no-cache forward: position p carries the rotation of the call that
    produced it -- long_factor iff p >= long_from (a cached decode step at position p has kv_seq_len = p + 1; the keys already
    in the cache are not re-rotated, and prepare_inputs_for_generation's reset (:1557-1562) never fires when generating from
    inputs_embeds because its input_ids holds only the generated tokens)."""
    positions = positions.to(torch.float32)
    shape = torch.arange(0, head_dim, 2, dtype=torch.int64, device=positions.device).float() / head_dim
    scale = max_pos / orig_max_pos
    sf = 1.0 if scale <= 1.0 else math.sqrt(1 + math.log(scale) / math.log(orig_max_pos))

    def table(factor):
        ext = torch.tensor(factor, dtype=torch.float32, device=positions.device)
        inv_freq = 1.0 / (ext * base ** shape)
        freqs = positions[:, None] * inv_freq[None, :]
        emb = torch.cat([freqs, freqs], dim=-1)
        return emb.cos() * sf, emb.sin() * sf

    if long_from is not None:
        (cs, ss), (cl, sl) = table(short_factor), table(long_factor)
        is_long = (positions >= long_from)[:, None]
        return torch.where(is_long, cl, cs), torch.where(is_long, sl, ss)
    if seq_len is None:
        seq_len = int(positions.max().item()) + 1
    return table(long_factor if seq_len > orig_max_pos else short_factor)


def plain_rope_tables(positions, head_dim, base, bf16_matmul_quirk=False):
    """Phi3RotaryEmbedding (modeling_phi3.py:345-368, rope_scaling=None) / LlamaRotaryEmbedding (modeling_llama.py:94-133).
    The Llama forward is NOT wrapped in autocast(enabled=False): under the reference's CUDA bf16 autocast the
    `inv_freq @ position_ids` matmul runs with bf16 operands and a bf16 result (bf16_matmul_quirk=True)."""
    inv_freq = 1.0 / (base ** (torch.arange(0, head_dim, 2, dtype=torch.int64, device=positions.device).float() / head_dim))
    pos = positions.to(torch.float32)
    if bf16_matmul_quirk:
        freqs = bf(bf(pos)[:, None] * bf(inv_freq)[None, :])
        emb = torch.cat([freqs, freqs], dim=-1)
        return bf(emb.cos()), bf(emb.sin())
    freqs = pos[:, None] * inv_freq[None, :]
    emb = torch.cat([freqs, freqs], dim=-1)
    return emb.cos(), emb.sin()


def rotate_half(x):
    h = x.shape[-1] // 2
    return torch.cat([-x[..., h:], x[..., :h]], dim=-1)


def apply_rope(q, cos, sin, mode):
    """apply_rotary_pos_emb (modeling_phi3.py:413-445): (q*cos) + (rotate_half(q)*sin) on bf16 tensors."""
    if mode == "bf16":
        return bf(bf(q * cos) + bf(rotate_half(q) * sin))
    return q * cos + rotate_half(q) * sin


def lm_forward(embeds, P, cfg, mode="bf16", positions=None, return_hidden=False):
    """Phi3Model.forward / LlamaModel.forward without cache + lm_head (modeling_phi3.py:1249-1383, 1034-1095,
    629-775 (attention), 458-464 (MLP), 1525-1526 (logits.float()); modeling_llama.py:934-1044, 699-760, 218-238).
    embeds [S,D] (one unpadded sequence). P uses the reference's parameter names ("model.layers.N...").
    cfg: dict(arch='phi3'|'llama', layers, heads, kv_heads, head_dim, eps, rope=dict(...))."""
    S, D = embeds.shape
    H, KVH, hd = cfg["heads"], cfg["kv_heads"], cfg["head_dim"]
    if positions is None:
        positions = torch.arange(S, device=embeds.device)
    rp = cfg["rope"]
    if rp["type"] == "longrope":
        cos, sin = phi3_rope_tables(positions, hd, rp["base"], rp["short_factor"], rp["long_factor"], rp["max_pos"],
                                    rp["orig_max_pos"], seq_len=rp.get("seq_len"), long_from=rp.get("long_from"))
    else:
        cos, sin = plain_rope_tables(positions, hd, rp["base"], bf16_matmul_quirk=(mode == "bf16" and rp.get("bf16_quirk", False)))
    cos, sin = _r(cos, mode), _r(sin, mode)
    x = _r(embeds.float(), mode)
    scale = hd ** -0.5
    for l in range(cfg["layers"]):
        pre = "model.layers.%d." % l
        h = rmsnorm(x, P[pre + "input_layernorm.weight"].float(), cfg["eps"], mode)
        if cfg["arch"] == "phi3":
            qkv = linear(h, P[pre + "self_attn.qkv_proj.weight"].float(), None, mode)
            q, k, v = qkv[:, :H * hd], qkv[:, H * hd:(H + KVH) * hd], qkv[:, (H + KVH) * hd:]
        else:
            q = linear(h, P[pre + "self_attn.q_proj.weight"].float(), None, mode)
            k = linear(h, P[pre + "self_attn.k_proj.weight"].float(), None, mode)
            v = linear(h, P[pre + "self_attn.v_proj.weight"].float(), None, mode)
        q = q.reshape(S, H, hd).transpose(0, 1)
        k = k.reshape(S, KVH, hd).transpose(0, 1)
        v = v.reshape(S, KVH, hd).transpose(0, 1)
        q = apply_rope(q, cos[None], sin[None], mode)
        k = apply_rope(k, cos[None], sin[None], mode)
        rep = H // KVH
        kk = k.repeat_interleave(rep, dim=0)
        vv = v.repeat_interleave(rep, dim=0)
        o = attention_core(q[None], kk[None], vv[None], scale, True, mode)[0]
        o = o.transpose(0, 1).reshape(S, H * hd)
        x = _r(x + linear(o, P[pre + "self_attn.o_proj.weight"].float(), None, mode), mode)
        h = rmsnorm(x, P[pre + "post_attention_layernorm.weight"].float(), cfg["eps"], mode)
        if cfg["arch"] == "phi3":
            gu = linear(h, P[pre + "mlp.gate_up_proj.weight"].float(), None, mode)
            gate, up = gu.chunk(2, dim=-1)
        else:
            gate = linear(h, P[pre + "mlp.gate_proj.weight"].float(), None, mode)
            up = linear(h, P[pre + "mlp.up_proj.weight"].float(), None, mode)
        if cfg["arch"] == "phi3":
            m = _r(up * _r(F.silu(gate), mode), mode)            # up * act(gate)   (modeling_phi3.py:461-462)
        else:
            m = _r(_r(F.silu(gate), mode) * up, mode)            # act(gate) * up   (modeling_llama.py:236)
        x = _r(x + linear(m, P[pre + "mlp.down_proj.weight"].float(), None, mode), mode)
    hidden = x
    hN = rmsnorm(x, P["model.norm.weight"].float(), cfg["eps"], mode)
    b = P.get("lm_head.bias")
    logits = linear(hN, P["lm_head.weight"].float(), None if b is None else b.float(), mode).float()
    if return_hidden:
        return logits, hidden
    return logits


def greedy_decode(embeds, P, cfg, n_new, mode="bf16", eos_id=None, pad_id=0):
    """Teacher-forced restatement of HF GenerationMixin greedy search driven by inputs_embeds
    (transformers==4.40.1, called at llava_next_video.py:655-661): step 0 consumes inputs_embeds, every later
    step consumes the embedding of the previous argmax; finished rows emit pad_id. Recomputes the full
    no-cache forward each step (mathematically identical to the KV-cached path) -- with LongRoPE that needs
    the per-position factor choice of the cached path (phi3_rope_tables `long_from`): a prompt of S <= original_max rotates
    positions < original_max with short_factor and later ones with long_factor; a longer prompt uses long_factor throughout."""
    table = P["model.embed_tokens.weight"].float()
    seq = _r(embeds.float(), mode)
    if cfg["rope"]["type"] == "longrope" and "long_from" not in cfg["rope"]:
        rp = dict(cfg["rope"])
        rp["long_from"] = rp["orig_max_pos"] if seq.shape[0] <= rp["orig_max_pos"] else 0
        cfg = dict(cfg, rope=rp)
    toks, all_logits = [], []
    finished = False
    for _ in range(n_new):
        logits = lm_forward(seq, P, cfg, mode)[-1]
        all_logits.append(logits)
        t = int(torch.argmax(logits).item())
        if finished:
            t = pad_id
        elif eos_id is not None and t == eos_id:
            finished = True
        toks.append(t)
        seq = torch.cat([seq, _r(table[t][None], mode)], dim=0)
    return toks, torch.stack(all_logits)


# ----------------------------------------------------------------------------- host-side integer / string logic
def tokenizer_image_token(prompt, tokenizer, image_token_index=IMAGE_TOKEN_INDEX):
    """LLAVA_NEXT_VIDEO.tokenizer_image_token (llava_next_video.py:409-426)."""
    chunks = [tokenizer(chunk).input_ids for chunk in prompt.split(DEFAULT_IMAGE_TOKEN)]
    ids = []
    offset = 0
    if len(chunks) > 0 and len(chunks[0]) > 0 and chunks[0][0] == tokenizer.bos_token_id:
        offset = 1
        ids.append(chunks[0][0])
    sep = [image_token_index] * (offset + 1)
    inter = [e for pair in zip(chunks, [sep] * len(chunks)) for e in pair][:-1]
    for x in inter:
        ids.extend(x[offset:])
    return ids


def left_pad_batch(id_lists, pad_id, max_txt_len):
    """generate() pre-amble (llava_next_video.py:622-647): flip / pad_sequence / truncate / flip == left padding."""
    L = min(max(len(x) for x in id_lists), max_txt_len)
    ids = torch.full((len(id_lists), L), pad_id, dtype=torch.long)
    mask = torch.zeros((len(id_lists), L), dtype=torch.long)
    for r, x in enumerate(id_lists):
        x = list(x)[::-1][:L][::-1]          # flipped sequence truncated at max_txt_len keeps the TAIL
        ids[r, L - len(x):] = torch.tensor(x, dtype=torch.long)
        mask[r, L - len(x):] = 1
    return ids, mask


def parse_time_interval(text, duration, num_temporal_tokens=300, llm="phi3.5"):
    """inference.py:125-134."""
    def rep(m):
        x = int(m.group(1))
        sec = duration * x / num_temporal_tokens
        return (" %.2f seconds" % sec) if llm == "phi3.5" else ("%.2f seconds" % sec)
    return re.sub(r"<(\d+)>", rep, text)


def quantize_referring(query, duration, num_temporal_tokens=300):
    """inference.py:107: k = int(float(sec) / duration * num_temporal_tokens)."""
    return re.sub(r"(\d+) seconds", lambda m: "<%d>" % int(float(m.group(1)) / duration * num_temporal_tokens), query)


def quantize_training(t, duration, num_temporal_tokens=300):
    """datasets/mix_grounded.py:78-91: k = min(int(num_temporal_tokens * t / duration), num_temporal_tokens)."""
    return min(int(num_temporal_tokens * t / duration), num_temporal_tokens)


def get_frame_indices_middle(num_frames, vlen):
    """mm_utils/video_utils.py:13-51 with sample='middle'."""
    import numpy as np
    acc = min(num_frames, vlen)
    intervals = np.linspace(start=0, stop=vlen, num=acc + 1).astype(int)
    idx = [(int(intervals[i]) + int(intervals[i + 1]) - 1) // 2 for i in range(acc)]
    if len(idx) < num_frames:
        idx = idx + [idx[-1]] * (num_frames - len(idx))
    return idx


def spatial_keyframe_indices(num_frames, num_segs):
    """inference.py:81-83."""
    per = int(num_frames // num_segs)
    return [(i * per) + int(per / 2) for i in range(num_segs)]


# ----------------------------------------------------------------------------- random-init parameter factories
def _tn(shape, std, gen):
    return torch.nn.init.trunc_normal_(torch.empty(shape), std=std, a=-2 * std, b=2 * std, generator=gen)


def make_clip_params(dim=1024, heads=16, ffn=4096, layers=24, image=336, seed=0):
    """Random-init CLIP ViT parameters with the reference's names and initialiser scales (modeling_clip.py:406-430)."""
    g = torch.Generator().manual_seed(seed)
    n_pos = (image // 14) ** 2 + 1
    P = {
        "vision_model.embeddings.class_embedding": torch.randn(dim, generator=g) * dim ** -0.5,
        "vision_model.embeddings.patch_embedding.weight": torch.randn(dim, 3, 14, 14, generator=g) * 0.02,
        "vision_model.embeddings.position_embedding.weight": torch.randn(n_pos, dim, generator=g) * 0.02,
        "vision_model.pre_layrnorm.weight": 1 + 0.1 * torch.randn(dim, generator=g),
        "vision_model.pre_layrnorm.bias": 0.1 * torch.randn(dim, generator=g),
    }
    in_std = dim ** -0.5 * (2 * layers) ** -0.5
    for l in range(layers):
        pre = "vision_model.encoder.layers.%d." % l
        for n in ("q_proj", "k_proj", "v_proj"):
            P[pre + "self_attn.%s.weight" % n] = torch.randn(dim, dim, generator=g) * in_std
            P[pre + "self_attn.%s.bias" % n] = 0.02 * torch.randn(dim, generator=g)
        P[pre + "self_attn.out_proj.weight"] = torch.randn(dim, dim, generator=g) * dim ** -0.5
        P[pre + "self_attn.out_proj.bias"] = 0.02 * torch.randn(dim, generator=g)
        P[pre + "mlp.fc1.weight"] = torch.randn(ffn, dim, generator=g) * (2 * dim) ** -0.5
        P[pre + "mlp.fc1.bias"] = 0.02 * torch.randn(ffn, generator=g)
        P[pre + "mlp.fc2.weight"] = torch.randn(dim, ffn, generator=g) * in_std
        P[pre + "mlp.fc2.bias"] = 0.02 * torch.randn(dim, generator=g)
        for n in ("layer_norm1", "layer_norm2"):
            P[pre + n + ".weight"] = 1 + 0.1 * torch.randn(dim, generator=g)
            P[pre + n + ".bias"] = 0.1 * torch.randn(dim, generator=g)
    return P


def make_iv2_params(dim=1408, heads=16, ffn=6144, depth=40, frames=8, image=224, seed=0, gamma=None):
    """Random-init InternVideo2 parameters (internvideo2.py:929-944 initialisers incl. fix_init_weight).
    gamma=None keeps the constructed LayerScale 1e-5; a float pair (lo, hi) draws U(lo,hi) so the branch is visible."""
    g = torch.Generator().manual_seed(seed)
    n_tok = frames * (image // 14) ** 2 + 1
    P = {
        "patch_embed.proj.weight": torch.randn(dim, 3, 1, 14, 14, generator=g) * 0.02,
        "patch_embed.proj.bias": 0.02 * torch.randn(dim, generator=g),
        "cls_token": _tn((1, 1, dim), 0.02, g),
        "pos_embed": 0.02 * torch.randn(1, n_tok, dim, generator=g),
    }
    for i in range(depth):
        pre = "blocks.%d." % i
        P[pre + "norm1.weight"] = 1 + 0.1 * torch.randn(dim, generator=g)
        P[pre + "norm2.weight"] = 1 + 0.1 * torch.randn(dim, generator=g)
        P[pre + "attn.qkv.weight"] = _tn((3 * dim, dim), 0.02, g)
        P[pre + "attn.q_norm.weight"] = 1 + 0.1 * torch.randn(dim, generator=g)
        P[pre + "attn.k_norm.weight"] = 1 + 0.1 * torch.randn(dim, generator=g)
        P[pre + "attn.proj.weight"] = _tn((dim, dim), 0.02, g) / math.sqrt(2.0 * (i + 1))
        P[pre + "attn.proj.bias"] = 0.02 * torch.randn(dim, generator=g)
        P[pre + "mlp.fc1.weight"] = _tn((ffn, dim), 0.02, g)
        P[pre + "mlp.fc1.bias"] = 0.02 * torch.randn(ffn, generator=g)
        P[pre + "mlp.fc2.weight"] = _tn((dim, ffn), 0.02, g) / math.sqrt(2.0 * (i + 1))
        P[pre + "mlp.fc2.bias"] = 0.02 * torch.randn(dim, generator=g)
        for n in ("ls1.gamma", "ls2.gamma"):
            if gamma is None:
                P[pre + n] = 1e-5 * torch.ones(dim)
            else:
                P[pre + n] = gamma[0] + (gamma[1] - gamma[0]) * torch.rand(dim, generator=g)
    return P


def make_lm_params(arch="phi3", dim=3072, heads=32, kv_heads=32, head_dim=96, ffn=8192, layers=32, vocab=32366,
                   seed=0, lm_head_bias=True, std=0.02):
    """Random-init decoder parameters, HF normal_(0, 0.02) (modeling_phi3.py:1131-1140); lm_head gets a bias as in
    reset_embeddings (llava_next_video.py:263)."""
    g = torch.Generator().manual_seed(seed)
    P = {"model.embed_tokens.weight": torch.randn(vocab, dim, generator=g) * std,
         "model.norm.weight": 1 + 0.1 * torch.randn(dim, generator=g),
         "lm_head.weight": torch.randn(vocab, dim, generator=g) * std}
    if lm_head_bias:
        P["lm_head.bias"] = (torch.rand(vocab, generator=g) * 2 - 1) * dim ** -0.5
    for l in range(layers):
        pre = "model.layers.%d." % l
        P[pre + "input_layernorm.weight"] = 1 + 0.1 * torch.randn(dim, generator=g)
        P[pre + "post_attention_layernorm.weight"] = 1 + 0.1 * torch.randn(dim, generator=g)
        if arch == "phi3":
            P[pre + "self_attn.qkv_proj.weight"] = torch.randn((heads + 2 * kv_heads) * head_dim, dim, generator=g) * std
            P[pre + "mlp.gate_up_proj.weight"] = torch.randn(2 * ffn, dim, generator=g) * std
        else:
            P[pre + "self_attn.q_proj.weight"] = torch.randn(heads * head_dim, dim, generator=g) * std
            P[pre + "self_attn.k_proj.weight"] = torch.randn(kv_heads * head_dim, dim, generator=g) * std
            P[pre + "self_attn.v_proj.weight"] = torch.randn(kv_heads * head_dim, dim, generator=g) * std
            P[pre + "mlp.gate_proj.weight"] = torch.randn(ffn, dim, generator=g) * std
            P[pre + "mlp.up_proj.weight"] = torch.randn(ffn, dim, generator=g) * std
        P[pre + "self_attn.o_proj.weight"] = torch.randn(dim, heads * head_dim, generator=g) * std
        P[pre + "mlp.down_proj.weight"] = torch.randn(dim, ffn, generator=g) * std
    return P


# Phi-3.5-mini-instruct LongRoPE factors are part of the published config.json, which is not in this container
# (no network). These are smooth stand-ins with the published structure (48 values = head_dim/2, short ~1..1.3,
# long ~1..64) so that BOTH branches (seq_len <= 4096 / > 4096, modeling_phi3.py:380-385) are exercised.
def phi35_rope_cfg(head_dim=96, seq_len=None):
    n = head_dim // 2
    short = [1.0 + 0.3 * (i / max(n - 1, 1)) ** 2 for i in range(n)]
    long = [1.0 + 63.0 * (i / max(n - 1, 1)) ** 3 for i in range(n)]
    return dict(type="longrope", base=10000.0, short_factor=short, long_factor=long, max_pos=131072,
                orig_max_pos=4096, seq_len=seq_len)


def make_projector_params(kind, d_in, d_out, seed=0):
    g = torch.Generator().manual_seed(seed)
    n0, n1 = {"mm": ("linear_0", "linear_1"), "vp": ("up_proj", "down_proj")}[kind]
    return {
        "%s.%s.weight" % (kind, n0): (torch.rand(d_out, d_in, generator=g) * 2 - 1) * d_in ** -0.5,
        "%s.%s.bias" % (kind, n0): (torch.rand(d_out, generator=g) * 2 - 1) * d_in ** -0.5,
        "%s.%s.weight" % (kind, n1): (torch.rand(d_out, d_out, generator=g) * 2 - 1) * d_out ** -0.5,
        "%s.%s.bias" % (kind, n1): (torch.rand(d_out, generator=g) * 2 - 1) * d_out ** -0.5,
    }
