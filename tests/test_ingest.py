"""Checkpoint ingest: host logic vs golden vectors produced by the reference's own functions, and vs the
published PEFT / reset_embeddings semantics on small synthetic checkpoints laid out like the reference's files."""
import os

import numpy as np
import pytest
import torch

from gvl import ingest


def test_pos_embed_interpolation_bit_exact_vs_reference_golden(gold_dir):
    z = np.load(os.path.join(gold_dir, "ingest_pos_embed.npz"))      # oracle/make_golden.py: interpolate_pos_embed_internvideo2_new
    sd = {k[3:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("in_")}
    ingest.interpolate_pos_embed_internvideo2(sd, num_frames=8, num_patches_per_frame=16, orig_t_size=4)
    for k, v in sd.items():
        assert torch.equal(v, torch.from_numpy(z["out_" + k])), k
    assert sd["pos_embed"].shape == (1, 1 + 8 * 16, 24) and sd["img_pos_embed"].shape == (1, 17, 24)
    with pytest.raises(KeyError):
        ingest.interpolate_pos_embed_internvideo2({"x": torch.zeros(1)}, 8)


def test_merge_lora_matches_runtime_lora_and_key_layouts():
    g = torch.Generator().manual_seed(0)
    w = torch.randn(48, 32, generator=g)
    a, b = torch.randn(8, 32, generator=g) * 0.1, torch.randn(48, 8, generator=g) * 0.1
    x = torch.randn(5, 32, generator=g)
    want = x @ w.T + ingest.LORA_SCALE * ((x @ a.T) @ b.T)             # peft: result += lora_B(lora_A(dropout(x))) * scaling
    for layout in ("peft03", "base_layer"):
        base_key = "base_model.model.model.layers.0.self_attn.qkv_proj." + ("weight" if layout == "peft03" else "base_layer.weight")
        sd = {base_key: w.clone(),
              "base_model.model.model.layers.0.self_attn.qkv_proj.lora_A.default.weight": a,
              "base_model.model.model.layers.0.self_attn.qkv_proj.lora_B.default.weight": b,
              "base_model.model.model.norm.weight": torch.ones(32),
              "base_model.model.lm_head.bias": torch.zeros(7)}
        out = ingest.merge_lora(sd)
        assert set(out) == {"model.layers.0.self_attn.qkv_proj.weight", "model.norm.weight", "lm_head.bias"}
        got = x @ out["model.layers.0.self_attn.qkv_proj.weight"].T
        assert torch.allclose(got, want, atol=1e-5)
    plain = {"model.norm.weight": torch.ones(4)}
    assert ingest.merge_lora(plain) == plain
    with pytest.raises(KeyError):
        ingest.merge_lora({"m.lora_A.default.weight": a})


def test_reset_embeddings_and_token_ids():
    g = torch.Generator().manual_seed(1)
    e, h = torch.randn(50, 16, generator=g), torch.randn(50, 16, generator=g)
    e2, h2, b2 = ingest.reset_embeddings(e, h)
    assert e2.shape == (352, 16) and h2.shape == (352, 16) and b2.shape == (352,)
    assert torch.equal(e2[:50], e) and torch.equal(e2[50], torch.mean(e, dim=0)) and torch.equal(e2[351], e2[50])
    assert torch.equal(h2[60], torch.mean(h, dim=0))
    ids = ingest.temporal_token_ids(32011)
    assert ids["<0>"] == 32011 and ids["<300>"] == 32311 and ids["<timestamp_grounding>"] == 32312 and len(ids) == 302


def test_load_params_from_reference_file_layout(tmp_path):
    """A miniature of the reference's weight directory (README.md:59-80 layout), written with torch.save / safetensors."""
    from safetensors.torch import save_file
    root = tmp_path / "Phi-3.5-vision-instruct-seperated"
    (root / "language_model_seperated").mkdir(parents=True)
    g = torch.Generator().manual_seed(2)
    torch.save({"vision_model.embeddings.class_embedding": torch.randn(8, generator=g)}, root / "vision_model.pth")
    torch.save({"glb_GN": torch.randn(1, 1, 32, generator=g), "sub_GN": torch.randn(1, 1, 1, 32, generator=g)}, root / "image_newlines.pth")
    torch.save({"linear_0.weight": torch.randn(4, 4, generator=g)}, root / "multi_modal_projector.pth")
    lm = {"model.embed_tokens.weight": torch.randn(20, 8, generator=g), "model.layers.0.self_attn.qkv_proj.weight": torch.randn(24, 8, generator=g),
          "lm_head.weight": torch.randn(20, 8, generator=g)}
    save_file(lm, str(root / "language_model_seperated" / "model-00001-of-00001.safetensors"))
    video = tmp_path / "internvideo2-1B.pt"
    torch.save({"pos_embed": torch.randn(1, 1 + 4 * 256, 8, generator=g), "cls_token": torch.zeros(1, 1, 8)}, video)
    a, b = torch.randn(2, 8, generator=g), torch.randn(24, 2, generator=g)
    tuned = {"base_model.model.model.embed_tokens.weight": torch.randn(322, 8, generator=g),
             "base_model.model.lm_head.weight": torch.randn(322, 8, generator=g), "base_model.model.lm_head.bias": torch.randn(322, generator=g),
             "base_model.model.model.layers.0.self_attn.qkv_proj.weight": lm["model.layers.0.self_attn.qkv_proj.weight"].clone(),
             "base_model.model.model.layers.0.self_attn.qkv_proj.lora_A.default.weight": a,
             "base_model.model.model.layers.0.self_attn.qkv_proj.lora_B.default.weight": b}
    ckpt = tmp_path / "sft.pth"
    torch.save({"model": {"video_projecter": {"up_proj.weight": torch.randn(4, 4, generator=g)}, "language_model": tuned,
                          "multi_modal_projector": {"linear_0.weight": torch.ones(4, 4)}}}, ckpt)
    with pytest.raises(KeyError):
        ingest.load_params("phi3.5", str(root), str(video), None)
    p = ingest.load_params("phi3.5", str(root), str(video), str(ckpt), num_frames=96, num_segs=12)
    assert p["video_encoder"]["pos_embed"].shape == (1, 1 + 8 * 256, 8)            # 4 -> 8 frames per segment
    assert torch.equal(p["multi_modal_projector"]["linear_0.weight"], torch.ones(4, 4))
    assert p["language_model"]["model.embed_tokens.weight"].shape == (322, 8) and "lm_head.bias" in p["language_model"]
    want = lm["model.layers.0.self_attn.qkv_proj.weight"] + 2.0 * (b @ a)
    assert torch.allclose(p["language_model"]["model.layers.0.self_attn.qkv_proj.weight"], want, atol=1e-6)
    assert p["glb_GN"].shape == (1, 1, 32) and p["sub_GN"].shape == (1, 1, 1, 32)
