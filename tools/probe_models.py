"""GPU bring-up probe for the stage-level entry points (CLIP / IV2 / LM / full pipeline) against the oracle.
Each case runs in its own subprocess. Not a pytest file.
"""
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "grounded-video-llm_b200"))
sys.path.insert(0, ROOT)


def _stats(a, b):
    d = (a.float() - b.float()).abs()
    return "max_abs=%.4g mean_abs=%.4g ref_absmax=%.4g" % (d.max().item(), d.mean().item(), b.float().abs().max().item())


def case_clip_tiny():
    import torch
    from gvl import model
    from oracle import gvl_oracle as O
    P = O.make_clip_params(dim=64, heads=4, ffn=128, layers=4, image=56, seed=1)
    pix = torch.randn(3, 3, 56, 56, generator=torch.Generator().manual_seed(2))
    ref = O.clip_hidden_states(pix, P, 4, 4, mode="bf16", upto=3)[-1]
    m = model.CLIPVisionModel(P, num_heads=4, num_layers=4, image_size=56)
    out = m(pix.cuda(), output_hidden_states=True).hidden_states[-2]
    torch.cuda.synchronize()
    return _stats(out.cpu(), ref)


def case_clip_full_1img():
    import torch
    from gvl import model
    from oracle import gvl_oracle as O
    P = O.make_clip_params(seed=1)
    pix = torch.randn(1, 3, 336, 336, generator=torch.Generator().manual_seed(2))
    Pd = {k: v.cuda() for k, v in P.items()}
    ref = O.clip_hidden_states(pix.cuda(), Pd, 16, 24, mode="bf16", upto=23)[-1]     # oracle on the GPU (torch ops)
    m = model.CLIPVisionModel(P, num_heads=16, num_layers=24)
    out = m(pix.cuda(), output_hidden_states=True).hidden_states[-2]
    torch.cuda.synchronize()
    return _stats(out, ref)


def case_iv2_tiny():
    import torch
    from gvl import model
    from oracle import gvl_oracle as O
    P = O.make_iv2_params(dim=64, heads=4, ffn=128, depth=4, frames=2, seed=3, gamma=(0.5, 1.5))
    pix = torch.randn(2, 3, 2, 224, 224, generator=torch.Generator().manual_seed(4))
    ref = O.iv2_forward(pix, P, 4, 4, mode="bf16", x_vis_return_idx=-2)
    m = model.PretrainInternVideo2(P, num_heads=4, depth=4, num_frames=2)
    out = m(pix.cuda(), None, False, x_vis_return_idx=-2, x_vis_only=True)
    torch.cuda.synchronize()
    return _stats(out.cpu(), ref)


def case_iv2_full_1seg():
    import torch
    from gvl import model
    from oracle import gvl_oracle as O
    P = O.make_iv2_params(depth=40, seed=3, gamma=(0.05, 0.15))
    pix = torch.randn(1, 3, 8, 224, 224, generator=torch.Generator().manual_seed(4))
    Pd = {k: v.cuda() for k, v in P.items()}
    ref = O.iv2_forward(pix.cuda(), Pd, 16, 40, mode="bf16", x_vis_return_idx=-2)
    m = model.PretrainInternVideo2(P, num_heads=16, depth=40, num_frames=8)
    out = m(pix.cuda(), None, False, x_vis_return_idx=-2, x_vis_only=True)
    torch.cuda.synchronize()
    return _stats(out, ref)


def _lm_tiny(arch):
    import torch
    from gvl import model
    from oracle import gvl_oracle as O
    kvh = 4 if arch == "phi3" else 2
    P = O.make_lm_params(arch=arch, dim=256, heads=4, kv_heads=kvh, head_dim=64, ffn=512, layers=2, vocab=1000, seed=5,
                         std=0.05)
    if arch == "phi3":
        rope = O.phi35_rope_cfg(64)
        rope["orig_max_pos"] = 4096
    else:
        rope = dict(type="plain", base=500000.0, bf16_quirk=True)
    cfg = dict(arch=arch, layers=2, heads=4, kv_heads=kvh, head_dim=64, eps=1e-5, rope=rope)
    emb = torch.randn(40, 256, generator=torch.Generator().manual_seed(6)) * 0.5
    ref_logits = O.lm_forward(emb, P, cfg, mode="bf16")
    toks_ref, lg_ref = O.greedy_decode(emb, P, cfg, 6, mode="bf16")
    lm = model.CausalLM(P, arch, 4, kvh, 64, 1e-5, rope, max_ctx=256)
    out = lm(inputs_embeds=emb.cuda()[None]).logits[0]
    toks, lg = lm.generate(inputs_embeds=emb.cuda()[None], max_new_tokens=6, return_logits=True)
    torch.cuda.synchronize()
    return "prefill_logits[%s] decode_logits[%s] tokens gvl=%s oracle=%s" % (
        _stats(out.cpu(), ref_logits), _stats(lg[0].cpu(), lg_ref), toks[0].tolist(), toks_ref)


def case_lm_tiny_phi3():
    return _lm_tiny("phi3")


def case_lm_tiny_llama():
    return _lm_tiny("llama")


def case_lm_wide_4layers():
    import torch
    from gvl import model
    from oracle import gvl_oracle as O
    P = O.make_lm_params(arch="phi3", layers=4, vocab=32366, seed=7)
    rope = O.phi35_rope_cfg(96)
    cfg = dict(arch="phi3", layers=4, heads=32, kv_heads=32, head_dim=96, eps=1e-5, rope=rope)
    emb = torch.randn(700, 3072, generator=torch.Generator().manual_seed(8)) * 0.05
    Pd = {k: v.cuda() for k, v in P.items()}
    ref_logits, ref_hidden = O.lm_forward(emb.cuda(), Pd, cfg, mode="bf16", return_hidden=True)
    lm = model.CausalLM(P, "phi3", 32, 32, 96, 1e-5, rope, max_ctx=1024)
    o = lm(inputs_embeds=emb.cuda()[None])
    toks, lg = lm.generate(inputs_embeds=emb.cuda()[None], max_new_tokens=4, return_logits=True)
    torch.cuda.synchronize()
    toks_ref, lg_ref = O.greedy_decode(emb.cuda(), Pd, cfg, 4, mode="bf16")
    return "hidden[%s] logits[%s] decode_logits[%s] tokens %s vs %s" % (
        _stats(o.hidden_states, ref_hidden), _stats(o.logits[0], ref_logits), _stats(lg[0], lg_ref), toks[0].tolist(),
        toks_ref)


def _rand_params_gpu():
    """Full-size random-init parameters created directly on the GPU (bf16-representable fp32 is unnecessary here)."""
    import torch
    g = torch.Generator(device="cuda").manual_seed(0)

    def rn(*shape, std=0.02):
        return torch.randn(*shape, device="cuda", generator=g) * std
    clip = {"vision_model.embeddings.class_embedding": rn(1024),
            "vision_model.embeddings.patch_embedding.weight": rn(1024, 3, 14, 14),
            "vision_model.embeddings.position_embedding.weight": rn(577, 1024),
            "vision_model.pre_layrnorm.weight": 1 + rn(1024), "vision_model.pre_layrnorm.bias": rn(1024)}
    for l in range(24):
        p = "vision_model.encoder.layers.%d." % l
        for n in ("q_proj", "k_proj", "v_proj", "out_proj"):
            clip[p + "self_attn.%s.weight" % n] = rn(1024, 1024)
            clip[p + "self_attn.%s.bias" % n] = rn(1024)
        clip[p + "mlp.fc1.weight"], clip[p + "mlp.fc1.bias"] = rn(4096, 1024), rn(4096)
        clip[p + "mlp.fc2.weight"], clip[p + "mlp.fc2.bias"] = rn(1024, 4096), rn(1024)
        for n in ("layer_norm1", "layer_norm2"):
            clip[p + n + ".weight"], clip[p + n + ".bias"] = 1 + rn(1024), rn(1024)
    iv2 = {"patch_embed.proj.weight": rn(1408, 3, 1, 14, 14), "patch_embed.proj.bias": rn(1408),
           "cls_token": rn(1, 1, 1408), "pos_embed": rn(1, 2049, 1408)}
    for i in range(40):
        p = "blocks.%d." % i
        iv2[p + "norm1.weight"], iv2[p + "norm2.weight"] = 1 + rn(1408), 1 + rn(1408)
        iv2[p + "attn.qkv.weight"] = rn(4224, 1408)
        iv2[p + "attn.q_norm.weight"], iv2[p + "attn.k_norm.weight"] = 1 + rn(1408), 1 + rn(1408)
        iv2[p + "attn.proj.weight"], iv2[p + "attn.proj.bias"] = rn(1408, 1408), rn(1408)
        iv2[p + "mlp.fc1.weight"], iv2[p + "mlp.fc1.bias"] = rn(6144, 1408), rn(6144)
        iv2[p + "mlp.fc2.weight"], iv2[p + "mlp.fc2.bias"] = rn(1408, 6144), rn(1408)
        iv2[p + "ls1.gamma"] = 1e-5 * torch.ones(1408, device="cuda")
        iv2[p + "ls2.gamma"] = 1e-5 * torch.ones(1408, device="cuda")
    lm = {"model.embed_tokens.weight": rn(32366, 3072).bfloat16(), "model.norm.weight": (1 + rn(3072)).bfloat16(),
          "lm_head.weight": rn(32366, 3072).bfloat16(), "lm_head.bias": rn(32366).bfloat16()}
    for l in range(32):
        p = "model.layers.%d." % l
        lm[p + "input_layernorm.weight"] = (1 + rn(3072)).bfloat16()
        lm[p + "post_attention_layernorm.weight"] = (1 + rn(3072)).bfloat16()
        lm[p + "self_attn.qkv_proj.weight"] = rn(9216, 3072).bfloat16()
        lm[p + "self_attn.o_proj.weight"] = rn(3072, 3072).bfloat16()
        lm[p + "mlp.gate_up_proj.weight"] = rn(16384, 3072).bfloat16()
        lm[p + "mlp.down_proj.weight"] = rn(3072, 8192).bfloat16()
    mm = {"linear_0.weight": rn(3072, 4096), "linear_0.bias": rn(3072), "linear_1.weight": rn(3072, 3072), "linear_1.bias": rn(3072)}
    vp = {"up_proj.weight": rn(3072, 1408), "up_proj.bias": rn(3072), "down_proj.weight": rn(3072, 3072), "down_proj.bias": rn(3072)}
    return {"vision_tower": clip, "video_encoder": iv2, "multi_modal_projector": mm, "video_projecter": vp,
            "language_model": lm, "glb_GN": rn(1, 1, 4096), "sub_GN": rn(1, 1, 1, 4096)}


def case_pipeline_timing():
    import torch
    from gvl import model, ops
    from oracle import gvl_oracle as O
    params = _rand_params_gpu()
    lm_cfg = dict(arch="phi3", heads=32, kv_heads=32, head_dim=96, eps=1e-5, rope=O.phi35_rope_cfg(96))
    m = model.LLAVA_NEXT_VIDEO(params, llm="phi3.5", lm_cfg=lm_cfg)
    del params
    g = torch.Generator(device="cuda").manual_seed(1234)
    sp = torch.randn(1, 12, 3, 336, 336, device="cuda", generator=g)
    tp = torch.randn(1, 96, 3, 224, 224, device="cuda", generator=g)
    ids = torch.randint(3, 32000, (64,), generator=torch.Generator().manual_seed(7))
    ids[20] = -200
    samples = {"spatial_pixel_values": sp, "temporal_pixel_values": tp, "input_ids": [ids.tolist()]}
    ev = lambda: torch.cuda.Event(enable_timing=True)
    res = []
    for it in range(3):
        e = [ev() for _ in range(6)]
        e[0].record()
        hs = m.vision_tower(sp[0], output_hidden_states=True).hidden_states[-2]
        e[1].record()
        xv = m.video_encoder(tp.reshape(12, 8, 3, 224, 224).permute(0, 2, 1, 3, 4).contiguous(), None, False,
                             x_vis_return_idx=-2, x_vis_only=True)
        e[2].record()
        feats = m.encode_images(samples)
        e[3].record()
        idt, mask = __import__("gvl.hostlogic", fromlist=["x"]).left_pad([ids.tolist()], 0, 2048)
        emb, _, masks = m.prepare_multimodal_inputs(idt, None, mask, feats, ["v"])
        logits, _ = m.language_model.prefill(emb[0], n_new=16)
        e[4].record()
        toks = m.language_model.generate(inputs_embeds=emb, attention_mask=masks, max_new_tokens=16)
        e[5].record()
        torch.cuda.synchronize()
        res.append("clip=%.2fms iv2=%.2fms encode_images(all)=%.2fms splice+prefill=%.2fms prefill+16decode=%.2fms S=%d" % (
            e[0].elapsed_time(e[1]), e[1].elapsed_time(e[2]), e[2].elapsed_time(e[3]), e[3].elapsed_time(e[4]),
            e[4].elapsed_time(e[5]), emb.shape[1]))
    return " || ".join(res) + " launches=%d tokens=%s" % (ops.launch_count(), toks[0].tolist())


def case_gemm_perf():
    import torch
    from gvl import ops
    out = []
    for (M, N, K, act, bn) in ((24588, 6144, 1408, 1, 0), (24588, 1408, 6144, 0, 0), (24588, 4224, 1408, 0, 0),
                               (24588, 1408, 1408, 0, 0), (6924, 4096, 1024, 2, 0), (6924, 1024, 4096, 0, 0),
                               (3484, 16384, 3072, 3, 0), (3484, 3072, 8192, 0, 0), (3484, 9216, 3072, 0, 0),
                               (8192, 8192, 8192, 0, 256), (8192, 8192, 8192, 0, 128)):
        a = torch.randn(M, K, device="cuda").bfloat16()
        w = torch.randn(N, K, device="cuda").bfloat16()
        b = torch.randn(N, device="cuda").bfloat16() if act in (1, 2) else None
        o = torch.empty(M, N // 2 if act == 3 else N, device="cuda", dtype=torch.bfloat16)
        for _ in range(3):
            ops.gemm(a, w, bias=b, act=act, out=o, bn=bn)
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(10):
            ops.gemm(a, w, bias=b, act=act, out=o, bn=bn)
        e.record()
        torch.cuda.synchronize()
        ms = s.elapsed_time(e) / 10
        tf = 2.0 * M * N * K / ms / 1e9
        # cuBLAS for comparison
        for _ in range(3):
            torch.matmul(a, w.t())
        s.record()
        for _ in range(10):
            torch.matmul(a, w.t())
        e.record()
        torch.cuda.synchronize()
        ms2 = s.elapsed_time(e) / 10
        out.append("[%dx%dx%d act%d bn%d] gvl %.3fms %.0fTF | cublas %.3fms %.0fTF" % (M, N, K, act, bn, ms, tf, ms2, 2.0 * M * N * K / ms2 / 1e9))
    return "\n   ".join(out)


def case_attn_perf():
    import torch
    from gvl import ops
    out = []
    for (B, H, S, D, causal) in ((12, 16, 2049, 88, False), (12, 16, 577, 64, False), (1, 32, 3484, 96, True)):
        q = torch.randn(B, S, H, D, device="cuda").bfloat16()
        k = torch.randn(B, S, H, D, device="cuda").bfloat16()
        v = torch.randn(B, S, H, D, device="cuda").bfloat16()
        for _ in range(3):
            ops.attention(q, k, v, D ** -0.5, causal=causal)
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(10):
            ops.attention(q, k, v, D ** -0.5, causal=causal)
        e.record()
        torch.cuda.synchronize()
        ms = s.elapsed_time(e) / 10
        fl = 4.0 * B * H * S * S * D * (0.5 if causal else 1.0)
        out.append("[B%d H%d S%d D%d causal%d] %.3fms %.0fTF" % (B, H, S, D, causal, ms, fl / ms / 1e9))
    return " | ".join(out)


CASES = {k[5:]: v for k, v in globals().items() if k.startswith("case_")}


def main():
    if os.environ.get("PROBE_CHILD") and len(sys.argv) > 1 and sys.argv[1] in CASES:
        t0 = time.time()
        print("RESULT %s :: %s (%.1fs)" % (sys.argv[1], CASES[sys.argv[1]](), time.time() - t0))
        return
    names = sys.argv[1:] if len(sys.argv) > 1 else list(CASES)
    for name in names:
        try:
            r = subprocess.run([sys.executable, __file__, name], capture_output=True, text=True, timeout=600, env=dict(os.environ, PROBE_CHILD="1"))
            txt = r.stdout + r.stderr
            if "RESULT" in txt:
                print(txt[txt.index("RESULT"):].strip())
            else:
                print("FAIL %s rc=%d :: %s" % (name, r.returncode, " | ".join(txt.strip().splitlines()[-8:])))
        except subprocess.TimeoutExpired:
            print("TIMEOUT %s" % name)
        sys.stdout.flush()


if __name__ == "__main__":
    main()
