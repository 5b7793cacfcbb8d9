"""Synthetic (random-init) parameters of the named architectures, created directly on the target device under the
reference's state_dict names. Used by bench.py / smoke (there is no network for checkpoints). Not on the hot path.
Init scales follow the reference initialisers (normal(0, 0.02) for the decoders, modeling_phi3.py:1131-1140;
trunc-normal 0.02 for InternVideo2, internvideo2.py:929-944; LayerScale 1e-5 as constructed).
"""
import torch

PHI35 = dict(arch="phi3", dim=3072, heads=32, kv_heads=32, head_dim=96, ffn=8192, layers=32, vocab=32064 + 302, eps=1e-5)
LLAMA3_8B = dict(arch="llama", dim=4096, heads=32, kv_heads=8, head_dim=128, ffn=14336, layers=32, vocab=128256 + 302, eps=1e-5)
CLIP_L336 = dict(dim=1024, heads=16, ffn=4096, layers=24, image=336)
IV2_1B = dict(dim=1408, heads=16, ffn=6144, depth=40)


def phi35_rope(head_dim=96):
    """Stand-in LongRoPE factors with the published structure (the real config.json is not available offline)."""
    n = head_dim // 2
    short = [1.0 + 0.3 * (i / max(n - 1, 1)) ** 2 for i in range(n)]
    long = [1.0 + 63.0 * (i / max(n - 1, 1)) ** 3 for i in range(n)]
    return dict(type="longrope", base=10000.0, short_factor=short, long_factor=long, max_pos=131072, orig_max_pos=4096)


def llama3_rope():
    return dict(type="plain", base=500000.0, bf16_quirk=True)


def make_params(llm="phi3.5", device="cuda", seed=0, lm=None, clip=None, iv2=None, frames_per_seg=8, lm_dtype=torch.bfloat16):
    lm = dict(lm or (PHI35 if llm == "phi3.5" else LLAMA3_8B))
    clip = dict(clip or CLIP_L336)
    iv2 = dict(iv2 or IV2_1B)
    g = torch.Generator(device=device).manual_seed(seed)

    def rn(*shape, std=0.02, dtype=torch.float32):
        return (torch.randn(*shape, device=device, generator=g) * std).to(dtype)

    D, F = clip["dim"], clip["ffn"]
    npos = (clip["image"] // 14) ** 2 + 1
    cp = {"vision_model.embeddings.class_embedding": rn(D),
          "vision_model.embeddings.patch_embedding.weight": rn(D, 3, 14, 14),
          "vision_model.embeddings.position_embedding.weight": rn(npos, D),
          "vision_model.pre_layrnorm.weight": 1 + rn(D), "vision_model.pre_layrnorm.bias": rn(D)}
    for l in range(clip["layers"]):
        p = "vision_model.encoder.layers.%d." % l
        for n in ("q_proj", "k_proj", "v_proj", "out_proj"):
            cp[p + "self_attn.%s.weight" % n] = rn(D, D)
            cp[p + "self_attn.%s.bias" % n] = rn(D)
        cp[p + "mlp.fc1.weight"], cp[p + "mlp.fc1.bias"] = rn(F, D), rn(F)
        cp[p + "mlp.fc2.weight"], cp[p + "mlp.fc2.bias"] = rn(D, F), rn(D)
        for n in ("layer_norm1", "layer_norm2"):
            cp[p + n + ".weight"], cp[p + n + ".bias"] = 1 + rn(D), rn(D)

    D, F = iv2["dim"], iv2["ffn"]
    vp_ = {"patch_embed.proj.weight": rn(D, 3, 1, 14, 14), "patch_embed.proj.bias": rn(D),
           "cls_token": rn(1, 1, D), "pos_embed": rn(1, 1 + frames_per_seg * 256, D)}
    for i in range(iv2["depth"]):
        p = "blocks.%d." % i
        vp_[p + "norm1.weight"], vp_[p + "norm2.weight"] = 1 + rn(D), 1 + rn(D)
        vp_[p + "attn.qkv.weight"] = rn(3 * D, D)
        vp_[p + "attn.q_norm.weight"], vp_[p + "attn.k_norm.weight"] = 1 + rn(D), 1 + rn(D)
        vp_[p + "attn.proj.weight"], vp_[p + "attn.proj.bias"] = rn(D, D), rn(D)
        vp_[p + "mlp.fc1.weight"], vp_[p + "mlp.fc1.bias"] = rn(F, D), rn(F)
        vp_[p + "mlp.fc2.weight"], vp_[p + "mlp.fc2.bias"] = rn(D, F), rn(D)
        vp_[p + "ls1.gamma"] = iv2.get("gamma", 1e-5) * torch.ones(D, device=device)
        vp_[p + "ls2.gamma"] = iv2.get("gamma", 1e-5) * torch.ones(D, device=device)

    Dm, H, KVH, hd, Fm, V = lm["dim"], lm["heads"], lm["kv_heads"], lm["head_dim"], lm["ffn"], lm["vocab"]
    dt = lm_dtype
    lp = {"model.embed_tokens.weight": rn(V, Dm, dtype=dt), "model.norm.weight": (1 + rn(Dm)).to(dt),
          "lm_head.weight": rn(V, Dm, dtype=dt), "lm_head.bias": rn(V, dtype=dt)}
    for l in range(lm["layers"]):
        p = "model.layers.%d." % l
        lp[p + "input_layernorm.weight"] = (1 + rn(Dm)).to(dt)
        lp[p + "post_attention_layernorm.weight"] = (1 + rn(Dm)).to(dt)
        if lm["arch"] == "phi3":
            lp[p + "self_attn.qkv_proj.weight"] = rn((H + 2 * KVH) * hd, Dm, dtype=dt)
            lp[p + "mlp.gate_up_proj.weight"] = rn(2 * Fm, Dm, dtype=dt)
        else:
            lp[p + "self_attn.q_proj.weight"] = rn(H * hd, Dm, dtype=dt)
            lp[p + "self_attn.k_proj.weight"] = rn(KVH * hd, Dm, dtype=dt)
            lp[p + "self_attn.v_proj.weight"] = rn(KVH * hd, Dm, dtype=dt)
            lp[p + "mlp.gate_proj.weight"] = rn(Fm, Dm, dtype=dt)
            lp[p + "mlp.up_proj.weight"] = rn(Fm, Dm, dtype=dt)
        lp[p + "self_attn.o_proj.weight"] = rn(Dm, H * hd, dtype=dt)
        lp[p + "mlp.down_proj.weight"] = rn(Dm, Fm, dtype=dt)

    out = {"vision_tower": cp, "video_encoder": vp_, "language_model": lp,
           "video_projecter": {"up_proj.weight": rn(Dm, iv2["dim"]), "up_proj.bias": rn(Dm),
                               "down_proj.weight": rn(Dm, Dm), "down_proj.bias": rn(Dm)}}
    if llm == "phi3.5":
        out["multi_modal_projector"] = {"linear_0.weight": rn(Dm, 4 * clip["dim"]), "linear_0.bias": rn(Dm),
                                        "linear_1.weight": rn(Dm, Dm), "linear_1.bias": rn(Dm)}
        out["glb_GN"] = rn(1, 1, 4 * clip["dim"])
        out["sub_GN"] = rn(1, 1, 1, 4 * clip["dim"])
    else:
        out["multi_modal_projector"] = {"linear_1.weight": rn(Dm, clip["dim"]), "linear_1.bias": rn(Dm),
                                        "linear_2.weight": rn(Dm, Dm), "linear_2.bias": rn(Dm)}
        out["image_newline"] = rn(Dm)
    lm_cfg = dict(arch=lm["arch"], heads=H, kv_heads=KVH, head_dim=hd, eps=lm["eps"],
                  rope=phi35_rope(hd) if lm["arch"] == "phi3" else llama3_rope())
    return out, lm_cfg, dict(heads=clip["heads"], layers=clip["layers"], image=clip["image"]), dict(heads=iv2["heads"], depth=iv2["depth"])


def make_clip_inputs(batch, num_frames=96, num_segs=12, seed=1234, device="cpu", pin=False):
    """Synthetic pre-normalised inputs of the final shapes: N(0,1) fp32, seed 1234; 64 text ids with the
    <image> sentinel at position 20 (seed 7)."""
    # one generator per clip (seed + b): clip b holds the same values whatever the batch size, so that the tokens of clip 0 can be
    # compared across 1 / 2 / 4 / 8 GPUs and clips-per-GPU settings (bench.py extra.tokens_clip0_sha256_16)
    sp = torch.empty(batch, num_segs, 3, 336, 336)
    tp = torch.empty(batch, num_frames, 3, 224, 224)
    for b in range(batch):
        g = torch.Generator().manual_seed(seed + b)
        sp[b] = torch.randn(num_segs, 3, 336, 336, generator=g)
        tp[b] = torch.randn(num_frames, 3, 224, 224, generator=g)
    ids = torch.randint(3, 32000, (64,), generator=torch.Generator().manual_seed(7))
    ids[20] = -200
    if pin:
        sp, tp = sp.pin_memory(), tp.pin_memory()
    if device != "cpu":
        sp, tp = sp.to(device), tp.to(device)
    return {"spatial_pixel_values": sp, "temporal_pixel_values": tp, "input_ids": [ids.tolist()] * batch,
            "video_ids": ["synthetic"] * batch, "pad_token_id": 0, "eos_token_id": None}
