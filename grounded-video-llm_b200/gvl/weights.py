"""Weight ingest: reference state_dicts -> packed bf16/fp32 device buffers + the C-ABI weight tables.

Input dictionaries use the REFERENCE's parameter names (CLIPVisionModel.state_dict(), PretrainInternVideo2
.state_dict(), Phi3ForCausalLM / LlamaForCausalLM.state_dict(), projector state_dicts), so a maintainer can feed
`torch.load(ckpt)['model'][...]` (inference.py:156-162) straight in. Load-time only; not on the hot path.

Packing decisions (all exact unless noted):
  * CLIP q/k/v are fused into one [3D, D] operand; the q rows and q bias are pre-multiplied by head_dim^-0.5
    (modeling_clip.py:263 multiplies the projected q; exact when the scale is a power of two, which it is for the
    real model: 64^-0.5 = 2^-3).
  * Patch-embed conv weights are flattened to [D, 3*14*14] and zero-padded along K to a multiple of 64 so the
    TMA row stride is 16-byte aligned (modeling_clip.py:169-175, internvideo2.py:714-718).
  * Phi-3 gate_up_proj / Llama gate_proj+up_proj rows are interleaved per 256-row block (128 gate rows, 128 up
    rows) so one GEMM tile holds gate and up of the same output column and SwiGLU runs in the epilogue
    (modeling_phi3.py:458-464).
  * Llama q/k/v are concatenated to one qkv operand (modeling_llama.py:283-286).
  * InternVideo2 parameters are rounded to bf16 first, because the reference casts the whole tower with
    `.to(torch.bfloat16)` (llava_next_video.py:134); LayerScale gamma is then widened back to fp32 (:451-466).
"""
import ctypes
import math

import torch

from . import _lib

KPAD = 640  # 3*14*14 = 588 -> 640


def _dev(t, dtype, device):
    return t.detach().to(device=device, dtype=dtype).contiguous()


def _is_pow2(x):
    m, _ = math.frexp(x)
    return m == 0.5


class Packed:
    """Keeps device tensors alive for as long as the C structs that point at them."""

    def __init__(self):
        self.tensors = []

    def keep(self, t):
        self.tensors.append(t)
        return ctypes.c_void_p(t.data_ptr())


def _pad_k(w2d, kpad):
    out = torch.zeros((w2d.shape[0], kpad), dtype=w2d.dtype, device=w2d.device)
    out[:, : w2d.shape[1]] = w2d
    return out


def pack_clip(sd, heads, n_layers_run, device="cuda", image=336):
    """sd: CLIPVisionModel.state_dict() (keys 'vision_model....'). n_layers_run = layers needed for hidden_states[-2]."""
    pk = Packed()
    bf, f32 = torch.bfloat16, torch.float32
    pw = sd["vision_model.embeddings.patch_embedding.weight"]
    D = pw.shape[0]
    hd = D // heads
    scale = hd ** -0.5
    prescale = _is_pow2(scale)
    ffn = sd["vision_model.encoder.layers.0.mlp.fc1.weight"].shape[0]
    n_pos = sd["vision_model.embeddings.position_embedding.weight"].shape[0]
    if n_pos != (image // 14) ** 2 + 1 or tuple(pw.shape[1:]) != (3, 14, 14):
        raise ValueError("CLIP tables do not match image=%d: %d position rows (want %d), patch kernel %s" % (
            image, n_pos, (image // 14) ** 2 + 1, tuple(pw.shape[1:])))
    layers = (_lib.ClipLayer * n_layers_run)()
    for l in range(n_layers_run):
        p = "vision_model.encoder.layers.%d." % l
        qs = scale if prescale else 1.0
        qkv_w = torch.cat([sd[p + "self_attn.q_proj.weight"].float() * qs, sd[p + "self_attn.k_proj.weight"].float(),
                           sd[p + "self_attn.v_proj.weight"].float()], dim=0)
        qkv_b = torch.cat([sd[p + "self_attn.q_proj.bias"].float() * qs, sd[p + "self_attn.k_proj.bias"].float(),
                           sd[p + "self_attn.v_proj.bias"].float()], dim=0)
        L = layers[l]
        L.ln1_w = pk.keep(_dev(sd[p + "layer_norm1.weight"], f32, device))
        L.ln1_b = pk.keep(_dev(sd[p + "layer_norm1.bias"], f32, device))
        L.qkv_w = pk.keep(_dev(qkv_w, bf, device))
        L.qkv_b = pk.keep(_dev(qkv_b, bf, device))
        L.out_w = pk.keep(_dev(sd[p + "self_attn.out_proj.weight"], bf, device))
        L.out_b = pk.keep(_dev(sd[p + "self_attn.out_proj.bias"], bf, device))
        L.ln2_w = pk.keep(_dev(sd[p + "layer_norm2.weight"], f32, device))
        L.ln2_b = pk.keep(_dev(sd[p + "layer_norm2.bias"], f32, device))
        L.fc1_w = pk.keep(_dev(sd[p + "mlp.fc1.weight"], bf, device))
        L.fc1_b = pk.keep(_dev(sd[p + "mlp.fc1.bias"], bf, device))
        L.fc2_w = pk.keep(_dev(sd[p + "mlp.fc2.weight"], bf, device))
        L.fc2_b = pk.keep(_dev(sd[p + "mlp.fc2.bias"], bf, device))
    w = _lib.ClipWeights()
    w.n_layers, w.dim, w.heads, w.ffn = n_layers_run, D, heads, ffn
    w.image = image
    w.n_patch = (image // 14) ** 2
    kreal = pw[0].numel()
    w.kpad = (kreal + 63) // 64 * 64
    w.patch_w = pk.keep(_dev(_pad_k(pw.reshape(D, -1).float(), w.kpad), bf, device))
    w.cls = pk.keep(_dev(sd["vision_model.embeddings.class_embedding"], f32, device))
    w.pos = pk.keep(_dev(sd["vision_model.embeddings.position_embedding.weight"], f32, device))
    w.pre_ln_w = pk.keep(_dev(sd["vision_model.pre_layrnorm.weight"], f32, device))
    w.pre_ln_b = pk.keep(_dev(sd["vision_model.pre_layrnorm.bias"], f32, device))
    w.layers = ctypes.cast(layers, ctypes.POINTER(_lib.ClipLayer))
    pk.layers = layers
    pk.struct = w
    pk.prescaled = prescale
    if not prescale:
        raise ValueError("CLIP head_dim^-0.5 must be a power of two for the fused q pre-scale (got head_dim=%d)" % hd)
    return pk


def pack_iv2(sd, heads, n_blocks_run, frames, device="cuda"):
    """sd: PretrainInternVideo2.state_dict() (keys 'patch_embed.proj.weight', 'blocks.N....')."""
    pk = Packed()
    bf, f32 = torch.bfloat16, torch.float32
    pw = sd["patch_embed.proj.weight"]
    D = pw.shape[0]
    ffn = sd["blocks.0.mlp.fc1.weight"].shape[0]
    # gvl_iv2_assemble indexes pos_embed by token: a checkpoint that was not temporally interpolated to `frames` (4-frame
    # checkpoint = 1025 rows, internvideo2.py:260-320) would be read out of bounds instead of failing
    if sd["pos_embed"].shape[-2] != 1 + frames * 256 or sd["pos_embed"].shape[-1] != D:
        raise ValueError("InternVideo2 pos_embed has %d rows, want 1 + %d*256 (interpolate it first: gvl.ingest)" % (
            sd["pos_embed"].shape[-2], frames))
    if sd["blocks.0.attn.q_norm.weight"].numel() != D or sd["blocks.0.attn.k_norm.weight"].numel() != D or D % heads:
        raise ValueError("InternVideo2 q_norm / k_norm must span the flattened %d-wide q / k rows" % D)
    blocks = (_lib.Iv2Block * n_blocks_run)()
    hd = D // heads
    hdp = (hd + 31) // 32 * 32          # 88 -> 96: TMA / tcgen05 friendly head stride; pad rows are exact zeros

    def pad_heads(t, lead):
        """[lead*heads*hd, ...] -> [lead*heads*hdp, ...] with zeros in each head's pad rows."""
        if hdp == hd:
            return t
        tail = t.shape[1:]
        t = t.reshape(lead, heads, hd, *tail)
        out = torch.zeros((lead, heads, hdp) + tuple(tail), dtype=t.dtype, device=t.device)
        out[:, :, :hd] = t
        return out.reshape(lead * heads * hdp, *tail)

    for i in range(n_blocks_run):
        p = "blocks.%d." % i
        B = blocks[i]
        B.norm1_w = pk.keep(_dev(sd[p + "norm1.weight"], bf, device))
        B.qkv_w = pk.keep(_dev(pad_heads(sd[p + "attn.qkv.weight"], 3), bf, device))
        B.q_norm_w = pk.keep(_dev(pad_heads(sd[p + "attn.q_norm.weight"], 1), bf, device))
        B.k_norm_w = pk.keep(_dev(pad_heads(sd[p + "attn.k_norm.weight"], 1), bf, device))
        B.proj_w = pk.keep(_dev(sd[p + "attn.proj.weight"], bf, device))
        B.proj_b = pk.keep(_dev(sd[p + "attn.proj.bias"], bf, device))
        B.ls1 = pk.keep(_dev(sd[p + "ls1.gamma"].to(bf), f32, device))
        B.norm2_w = pk.keep(_dev(sd[p + "norm2.weight"], bf, device))
        B.fc1_w = pk.keep(_dev(sd[p + "mlp.fc1.weight"], bf, device))
        B.fc1_b = pk.keep(_dev(sd[p + "mlp.fc1.bias"], bf, device))
        B.fc2_w = pk.keep(_dev(sd[p + "mlp.fc2.weight"], bf, device))
        B.fc2_b = pk.keep(_dev(sd[p + "mlp.fc2.bias"], bf, device))
        B.ls2 = pk.keep(_dev(sd[p + "ls2.gamma"].to(bf), f32, device))
    w = _lib.Iv2Weights()
    w.n_blocks, w.dim, w.heads, w.ffn, w.frames = n_blocks_run, D, heads, ffn, frames
    w.head_dim_pad = hdp
    kreal = pw[0].numel()
    w.kpad = (kreal + 63) // 64 * 64
    w.patch_w = pk.keep(_dev(_pad_k(pw.reshape(D, -1).float(), w.kpad), bf, device))
    w.patch_b = pk.keep(_dev(sd["patch_embed.proj.bias"], bf, device))
    w.cls = pk.keep(_dev(sd["cls_token"].reshape(-1), bf, device))
    w.pos = pk.keep(_dev(sd["pos_embed"].reshape(-1, D), bf, device))
    w.blocks = ctypes.cast(blocks, ctypes.POINTER(_lib.Iv2Block))
    pk.blocks = blocks
    pk.struct = w
    return pk


def interleave_gate_up(gate, up):
    """[F,D],[F,D] -> [2F,D] with 128 gate rows then 128 up rows per 256-row block."""
    F, D = gate.shape
    if F % 128 != 0:
        raise ValueError("intermediate size must be a multiple of 128")
    return torch.stack([gate.reshape(F // 128, 128, D), up.reshape(F // 128, 128, D)], dim=1).reshape(2 * F, D)


def pack_lm(sd, arch, heads, kv_heads, head_dim, rms_eps, max_ctx, rope_cos, rope_sin, device="cuda"):
    """sd: Phi3ForCausalLM / LlamaForCausalLM state_dict ('model.layers.N....', 'lm_head.weight' [, 'lm_head.bias']).
    rope_cos / rope_sin: fp32 or bf16 [max_ctx, head_dim] tables built by gvl.rope (rounded to bf16 here)."""
    pk = Packed()
    bf = torch.bfloat16
    n_layers = 0
    while ("model.layers.%d.input_layernorm.weight" % n_layers) in sd:
        n_layers += 1
    D = sd["model.embed_tokens.weight"].shape[1]
    layers = (_lib.LmLayer * n_layers)()
    ffn = None
    for l in range(n_layers):
        p = "model.layers.%d." % l
        L = layers[l]
        if arch == "phi3":
            qkv = sd[p + "self_attn.qkv_proj.weight"]
            gu = sd[p + "mlp.gate_up_proj.weight"]
            ffn = gu.shape[0] // 2
            gate, up = gu[:ffn], gu[ffn:]
        else:
            qkv = torch.cat([sd[p + "self_attn.q_proj.weight"], sd[p + "self_attn.k_proj.weight"],
                             sd[p + "self_attn.v_proj.weight"]], dim=0)
            gate, up = sd[p + "mlp.gate_proj.weight"], sd[p + "mlp.up_proj.weight"]
            ffn = gate.shape[0]
        L.in_norm_w = pk.keep(_dev(sd[p + "input_layernorm.weight"], bf, device))
        L.qkv_w = pk.keep(_dev(qkv, bf, device))
        L.o_w = pk.keep(_dev(sd[p + "self_attn.o_proj.weight"], bf, device))
        L.post_norm_w = pk.keep(_dev(sd[p + "post_attention_layernorm.weight"], bf, device))
        L.gate_up_w = pk.keep(_dev(interleave_gate_up(gate, up), bf, device))
        L.down_w = pk.keep(_dev(sd[p + "mlp.down_proj.weight"], bf, device))
    w = _lib.LmWeights()
    w.n_layers, w.dim, w.heads, w.kv_heads, w.head_dim = n_layers, D, heads, kv_heads, head_dim
    w.ffn, w.vocab, w.max_ctx, w.rms_eps = ffn, sd["lm_head.weight"].shape[0], max_ctx, rms_eps
    w.final_norm_w = pk.keep(_dev(sd["model.norm.weight"], bf, device))
    w.lm_head_w = pk.keep(_dev(sd["lm_head.weight"], bf, device))
    w.lm_head_b = pk.keep(_dev(sd["lm_head.bias"], bf, device)) if "lm_head.bias" in sd else None
    pk.embed = _dev(sd["model.embed_tokens.weight"], bf, device)
    w.embed = pk.keep(pk.embed)
    w.rope_cos = pk.keep(_dev(rope_cos, bf, device))
    w.rope_sin = pk.keep(_dev(rope_sin, bf, device))
    w.layers = ctypes.cast(layers, ctypes.POINTER(_lib.LmLayer))
    pk.layers = layers
    pk.struct = w
    return pk


def pack_mlp2(w0, b0, w1, b1, device="cuda"):
    """Linear-GELU-Linear projector (Phi3_5_Projecter / Video_Projecter / LlavaMultiModalProjector)."""
    bf = torch.bfloat16
    return tuple(_dev(t, bf, device) for t in (w0, b0, w1, b1))
