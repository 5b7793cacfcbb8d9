"""Generate tests/golden/* by RUNNING THE REFERENCE (imported from /root/reference through oracle/ref_shims.py).

    python oracle/make_golden.py

Each fixture holds seeded inputs, the (tiny, random-init) parameters under the reference's own state_dict names and
the reference's fp32 CPU outputs; tests/test_oracle_golden.py replays them through oracle/gvl_oracle.py on any box
(the reference itself does not travel to the GPU box). The script also prints oracle-vs-reference errors so a
drift is visible at generation time.
"""
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
from oracle import gvl_oracle as O  # noqa: E402
from oracle import ref_shims as R   # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def _np(d):
    return {k: v.detach().float().numpy() for k, v in d.items()}


def _save(name, **arrs):
    os.makedirs(GOLD, exist_ok=True)
    np.savez_compressed(os.path.join(GOLD, name), **arrs)
    print("wrote", name, "%.1f KB" % (os.path.getsize(os.path.join(GOLD, name)) / 1024))


def gold_clip(mods):
    from transformers import CLIPVisionConfig
    torch.manual_seed(11)
    cfg = CLIPVisionConfig(hidden_size=64, intermediate_size=128, num_hidden_layers=4, num_attention_heads=4,
                           image_size=56, patch_size=14, hidden_act="quick_gelu", layer_norm_eps=1e-5,
                           attention_dropout=0.0)
    cfg._attn_implementation = "eager"
    m = mods["clip"].CLIPVisionModel(cfg).eval()
    for p in m.parameters():                       # make biases / LN affine non-trivial
        if p.dim() == 1:
            p.data.add_(0.1 * torch.randn_like(p))
    pix = torch.randn(2, 3, 56, 56)
    with torch.no_grad():
        out = m(pix, output_hidden_states=True)
    hs_m2 = out.hidden_states[-2]
    P = {k: v for k, v in m.state_dict().items() if "position_ids" not in k}
    mine = O.clip_hidden_states(pix, P, 4, 4, mode="fp32")
    e = (mine[-2] - hs_m2).abs().max().item()
    e_all = max((a - b).abs().max().item() for a, b in zip(mine, out.hidden_states))
    print("clip  oracle(fp32) vs reference: hidden_states[-2] %.3g, all %.3g" % (e, e_all))
    assert e_all < 1e-4
    _save("clip_tiny.npz", pix=pix.numpy(), hs_m2=hs_m2.numpy(), hs_last=out.hidden_states[-1].numpy(),
          **{"P:" + k: v for k, v in _np(P).items()})


def gold_iv2(mods):
    torch.manual_seed(12)
    iv = mods["iv2"]
    m = iv.PretrainInternVideo2(
        in_chans=3, img_size=28, patch_size=14, embed_dim=64, depth=4, num_heads=4, mlp_ratio=2.0, clip_embed_dim=32,
        attn_pool_num_heads=4, qkv_bias=False, drop_path_rate=0.0, init_values=0.1, qk_normalization=True,
        use_flash_attn=False, use_fused_rmsnorm=False, use_fused_mlp=False, layerscale_no_force_fp32=False,
        num_frames=2, tubelet_size=1, sep_pos_embed=False, sep_image_video_pos_embed=True, use_checkpoint=False,
        checkpoint_num=0, clip_teacher_embed_dim=32, clip_teacher_final_dim=16, clip_return_layer=1).eval()
    for n, p in m.named_parameters():
        if p.dim() == 1 and "gamma" not in n:
            p.data.add_(0.1 * torch.randn_like(p))
        if "gamma" in n:
            p.data.copy_(0.5 + torch.rand_like(p))
    pix = torch.randn(2, 3, 2, 28, 28)
    with torch.no_grad():
        xv = m(pix, None, False, x_vis_return_idx=-2, x_vis_only=True)
        xv_full = m(pix, None, False, x_vis_return_idx=-1, x_vis_only=True)
    keep = ("patch_embed.", "cls_token", "pos_embed", "blocks.")
    P = {k: v for k, v in m.state_dict().items() if k.startswith(keep) and not k.startswith("clip_")}
    mine = O.iv2_forward(pix, P, 4, 4, mode="fp32", x_vis_return_idx=-2)
    mine_full = O.iv2_forward(pix, P, 4, 4, mode="fp32", x_vis_return_idx=-1)
    e1, e2 = (mine - xv).abs().max().item(), (mine_full - xv_full).abs().max().item()
    print("iv2   oracle(fp32) vs reference: idx=-2 %.3g, idx=-1 %.3g" % (e1, e2))
    assert max(e1, e2) < 1e-4
    _save("iv2_tiny.npz", pix=pix.numpy(), x_vis_m2=xv.numpy(), x_vis_m1=xv_full.numpy(),
          **{"P:" + k: v for k, v in _np(P).items()})


def _rope_cfg_tiny(hd):
    n = hd // 2
    return dict(short=[1.0 + 0.05 * i for i in range(n)], long=[1.0 + 0.5 * i for i in range(n)])


def gold_phi3(mods):
    torch.manual_seed(13)
    phi = mods["phi3"]
    Cfg = phi.Phi3Config
    hd = 16
    rc = _rope_cfg_tiny(hd)
    cfg = Cfg(vocab_size=97, hidden_size=64, intermediate_size=128, num_hidden_layers=2, num_attention_heads=4,
              num_key_value_heads=4, rms_norm_eps=1e-5, max_position_embeddings=64, original_max_position_embeddings=16,
              rope_theta=10000.0, sliding_window=None, attention_dropout=0.0, resid_pdrop=0.0, embd_pdrop=0.0,
              pad_token_id=0, bos_token_id=1, eos_token_id=2)
    cfg.rope_scaling = {"type": "longrope", "short_factor": rc["short"], "long_factor": rc["long"]}
    cfg._attn_implementation = "eager"
    m = phi.Phi3ForCausalLM(cfg).eval()
    # reset_embeddings (llava_next_video.py:231-268): lm_head replaced by a Linear WITH bias
    m.lm_head = torch.nn.Linear(64, 97, bias=True)
    for n, p in m.named_parameters():
        if p.dim() == 1 and "lm_head" not in n:
            p.data.add_(0.1 * torch.randn_like(p))
    outs = {}
    for name, S in (("short", 12), ("long", 24)):     # 24 > original_max_position_embeddings=16 -> long_factor branch
        emb = torch.randn(1, S, 64) * 0.5
        with torch.no_grad():
            o = m(inputs_embeds=emb, use_cache=False, return_dict=True)
        P = dict(m.state_dict())
        ocfg = dict(arch="phi3", layers=2, heads=4, kv_heads=4, head_dim=hd, eps=1e-5,
                    rope=dict(type="longrope", base=10000.0, short_factor=rc["short"], long_factor=rc["long"],
                              max_pos=64, orig_max_pos=16))
        mine = O.lm_forward(emb[0], P, ocfg, mode="fp32")
        e = (mine - o.logits[0]).abs().max().item()
        print("phi3  oracle(fp32) vs reference (%s, S=%d): logits %.3g" % (name, S, e))
        assert e < 1e-4
        outs["emb_" + name] = emb[0].numpy()
        outs["logits_" + name] = o.logits[0].numpy()
    _save("phi3_tiny.npz", short_factor=np.array(rc["short"], dtype=np.float64),
          long_factor=np.array(rc["long"], dtype=np.float64), **outs,
          **{"P:" + k: v for k, v in _np(m.state_dict()).items()})


def gold_llama(mods):
    torch.manual_seed(14)
    ll = mods["llama"]
    from transformers import LlamaConfig
    cfg = LlamaConfig(vocab_size=89, hidden_size=64, intermediate_size=160, num_hidden_layers=2, num_attention_heads=4,
                      num_key_value_heads=2, rms_norm_eps=1e-5, max_position_embeddings=64, attention_bias=False,
                      attention_dropout=0.0, pretraining_tp=1, pad_token_id=0, bos_token_id=1, eos_token_id=2)
    cfg.rope_theta = 500000.0
    cfg.rope_scaling = None
    cfg.pretraining_tp = 1
    cfg.attention_bias = False
    cfg.mlp_bias = False
    cfg._attn_implementation = "eager"
    m = ll.LlamaForCausalLM(cfg).eval()
    m.lm_head = torch.nn.Linear(64, 89, bias=True)
    for n, p in m.named_parameters():
        if p.dim() == 1 and "lm_head" not in n:
            p.data.add_(0.1 * torch.randn_like(p))
    emb = torch.randn(1, 14, 64) * 0.5
    with torch.no_grad():
        o = m(inputs_embeds=emb, use_cache=False, return_dict=True)
    P = dict(m.state_dict())
    ocfg = dict(arch="llama", layers=2, heads=4, kv_heads=2, head_dim=16, eps=1e-5, rope=dict(type="plain", base=500000.0))
    mine = O.lm_forward(emb[0], P, ocfg, mode="fp32")
    e = (mine - o.logits[0]).abs().max().item()
    print("llama oracle(fp32) vs reference: logits %.3g" % e)
    assert e < 1e-4
    _save("llama_tiny.npz", emb=emb[0].numpy(), logits=o.logits[0].numpy(),
          **{"P:" + k: v for k, v in _np(P).items() if "rotary_emb" not in k})


class _Self:
    """Stand-in for the LLAVA_NEXT_VIDEO instance the extracted methods expect."""
    device = "cpu"


def gold_index_maps():
    ns = R.std_namespace()
    f = "models/llava_next_video.py"
    merge = R.extract(f, "reshape_hd_patches_2x2merge_phi3", "LLAVA_NEXT_VIDEO", ns)
    newline = R.extract(f, "add_image_newline_phi3", "LLAVA_NEXT_VIDEO", ns)
    me = _Self()
    # integer-coded features: value = token*1024 + channel, so the output IS the source index map
    L, C = 576, 1024
    feat = (torch.arange(L)[:, None] * C + torch.arange(C)[None, :]).double()[None].repeat(2, 1, 1)
    feat[1] += L * C
    me.sub_GN = -(torch.arange(4 * C).double() + 1).reshape(1, 1, 1, -1)   # negative codes mark newline entries
    hd = merge(me, feat, 1, 1)
    out = newline(me, hd)
    mine = O.hd_merge_newline(feat, me.sub_GN)
    assert torch.equal(out, mine), "hd_merge_newline oracle != reference"
    # keep the full map of image 1 (it is regular, compresses to a few KB)
    _save("hd_merge_map.npz", index_map=out[1].to(torch.int64).numpy())

    # encode_images pooling / concat semantics, driven through the reference's own encode_images with stub towers
    enc = R.extract(f, "encode_images", "LLAVA_NEXT_VIDEO", ns)
    segs, fps, B = 3, 2, 2

    class Tower:
        def __call__(self, x, output_hidden_states=True):
            n = x.shape[0]
            hs = (torch.arange(n * 577 * 1024).double().reshape(n, 577, 1024),) * 3
            return type("O", (), {"hidden_states": hs})()

    class Video:
        def __call__(self, x, mask, use_image, x_vis_return_idx=-1, x_vis_only=False):
            n, T = x.shape[0], x.shape[2]
            return (torch.arange(n * (1 + T * 256) * 8).double() * 0.25).reshape(n, 1 + T * 256, 8)

    me2 = _Self()
    me2.llm = "phi3.5"
    me2.vision_tower = Tower()
    me2.video_encoder = Video()
    me2.multi_modal_projector = lambda t: t[..., :8] if t.shape[-1] > 8 else t      # keep shapes small, order intact
    me2.video_projecter = lambda t: t
    me2.sub_GN = -(torch.arange(4096).double() + 1).reshape(1, 1, 1, -1)
    me2.glb_GN = -(torch.arange(4096).double() + 5000).reshape(1, 1, -1)
    me2.reshape_hd_patches_2x2merge_phi3 = lambda a, b, c: merge(me2, a, b, c)
    me2.add_image_newline_phi3 = lambda a: newline(me2, a)
    samples = {"spatial_pixel_values": torch.zeros(B, segs, 3, 4, 4), "temporal_pixel_values": torch.zeros(B, segs * fps, 3, 4, 4)}
    vid = enc(me2, samples)
    # the same thing with the oracle's pieces
    hs = Tower()(torch.zeros(B * segs, 1))
    feat2 = hs.hidden_states[-2][:, 1:]
    sp = O.hd_merge_newline(feat2, me2.sub_GN)[..., :8].reshape(B, segs, 156, 8)
    xv = Video()(torch.zeros(B * segs, 3, fps, 4, 4), None, False)
    tm = O.pool_temporal(xv, fps).reshape(B, segs, fps * 16, 8)
    nlr = me2.glb_GN[0, 0, :8].reshape(1, 1, 1, 8).expand(B, segs, 1, 8)
    mine2 = torch.cat([sp, tm, nlr], dim=2).reshape(B, -1, 8)
    assert vid.shape == (B, segs * (156 + 16 * fps + 1), 8), vid.shape
    assert torch.equal(vid, mine2), "encode_images ordering oracle != reference"
    # llama3 / vicuna branch of the same method (3x3 pooled CLIP grid, learned image_newline)
    me4 = _Self()
    me4.llm = "llama3"
    me4.vision_tower = Tower()
    me4.video_encoder = Video()
    me4.multi_modal_projector = lambda t: t[..., :8]
    me4.video_projecter = lambda t: t
    me4.image_newline = -(torch.arange(8).double() + 7000)
    me4.config = type("C", (), {"hidden_size": 8})()
    vid_l = enc(me4, samples)
    sp_l = O.pool_spatial_llama(feat2)[..., :8].reshape(B, segs, 64, 8)
    nl_l = me4.image_newline.reshape(1, 1, 1, 8).expand(B, segs, 1, 8)
    mine_l = torch.cat([sp_l, tm, nl_l], dim=2).reshape(B, -1, 8)
    assert vid_l.shape == (B, segs * (64 + 16 * fps + 1), 8), vid_l.shape
    assert torch.allclose(vid_l, mine_l, rtol=0, atol=1e-9), "encode_images (llama) ordering oracle != reference"
    _save("encode_images_order.npz", video_features=vid.numpy(), video_features_llama=vid_l.numpy(),
          segs=np.int64(segs), fps=np.int64(fps))

    # prepare_multimodal_inputs
    prep = R.extract(f, "prepare_multimodal_inputs", "LLAVA_NEXT_VIDEO", ns)
    table = torch.arange(50 * 4).double().reshape(50, 4)
    me3 = _Self()
    me3.get_input_embeddings = lambda: (lambda ids: table[ids])
    ids = torch.tensor([[5, 7, -200, 9, 11, 13]])
    vis = -(torch.arange(3 * 4).double() + 1).reshape(1, 3, 4)
    res = {}
    for tag, vid_ids in (("video", ["v.mp4"]), ("text", ["text"])):
        emb, _, mask = prep(me3, ids, ids.clone(), torch.ones_like(ids), vis, vid_ids)
        mine3 = O.splice_embeds(ids[0], table, vis[0], vis_last=(tag == "text"))
        assert torch.equal(emb[0], mine3)
        res["embeds_" + tag] = emb[0].numpy()
        res["mask_" + tag] = mask[0].numpy()
    _save("splice.npz", ids=ids[0].numpy(), table=table.numpy(), visual=vis[0].numpy(), **res)


def gold_host_logic():
    ns = R.std_namespace()
    parse = R.extract("inference.py", "parse_time_interval", None, ns)
    gfi = R.extract("mm_utils/video_utils.py", "get_frame_indices", None, ns)
    tok = R.extract("models/llava_next_video.py", "tokenizer_image_token", "LLAVA_NEXT_VIDEO", ns)
    duration = 4257 / 29.97002997002997
    out = {"duration": duration, "parse": [], "frame_indices": [], "referring": [], "training_quant": [], "tokenize": []}
    for llm in ("phi3.5", "llama3"):
        for txt in ("From <30> to <53>.", "<228> <239>", "no tokens", "<0> and <300>"):
            out["parse"].append({"text": txt, "llm": llm, "duration": duration, "result": parse(txt, duration, 300, llm)})
    for n, vlen in ((96, 4257), (96, 300), (8, 4257), (16, 10), (96, 96), (12, 1000)):
        out["frame_indices"].append({"n": n, "vlen": vlen, "result": [int(x) for x in gfi(n, vlen, sample="middle")]})
    # inference.py:107 (inline lambda; restated literally here and pinned by the README answers 147 / 168)
    import re
    for q in ("What happens between 70 seconds and 80 seconds?", "at 0 seconds", "until 142 seconds"):
        r = re.sub(r"(\d+) seconds", lambda m: f"<{int(float(m.group(1))/duration*300)}>", q)
        out["referring"].append({"query": q, "duration": duration, "result": r})
        assert O.quantize_referring(q, duration) == r
    for t, d in ((14.2, 142.04), (142.04, 142.04), (200.0, 142.04), (0.0, 10.0), (3.3333, 10.0)):
        k = min(int(300 * t / d), 300)              # datasets/mix_grounded.py:83-84
        out["training_quant"].append({"t": t, "duration": d, "result": k})

    class Tok:
        bos_token_id = 1

        def __init__(self, bos):
            self.bos = bos

        def __call__(self, s):
            ids = ([1] if self.bos else []) + [3 + (ord(c) % 50) for c in s]
            return type("E", (), {"input_ids": ids})()

    me = _Self()
    for bos in (True, False):
        for prompt in ("<image> <timestamp_grounding>\nfind it", "hello <image>\nworld", "no image here", "<image>"):
            out["tokenize"].append({"prompt": prompt, "bos": bos, "result": tok(me, prompt, Tok(bos))})
            assert O.tokenizer_image_token(prompt, Tok(bos)) == out["tokenize"][-1]["result"]
    for e in out["parse"]:
        assert O.parse_time_interval(e["text"], e["duration"], 300, e["llm"]) == e["result"]
    for e in out["frame_indices"]:
        assert O.get_frame_indices_middle(e["n"], e["vlen"]) == e["result"]
    # README.md:90-94 known answers
    readme = {"<30>": "14.20", "<53>": "25.09", "<228>": "107.95", "<239>": "113.16"}
    for k, v in readme.items():
        assert parse(k, duration, 300, "phi3.5") == " %s seconds" % v, (k, parse(k, duration, 300, "phi3.5"))
    out["readme_kat"] = readme
    with open(os.path.join(GOLD, "host_logic.json"), "w") as fh:
        json.dump(out, fh, indent=1)
    print("wrote host_logic.json")

def golden_prompts(out_dir):
    """inference.py:90-116 prompt construction with the reference's own chat templates (datasets/chat/base_template.py, exec'd from
    its file with the dataclass decorators made hashable -- Python >= 3.11 rejects its mutable dataclass defaults)."""
    import re
    src = open(os.path.join(R.REF, "datasets", "chat", "base_template.py")).read()
    src = src.replace("@dataclass\n", "@dataclass(unsafe_hash=True)\n")
    src = src.replace("sys.path.append(os.path.abspath(os.path.join(__file__, \"..\", \"..\", \"..\")))", "")
    import sys
    import types
    mod = types.ModuleType("ref_base_template")          # dataclasses looks the defining module up in sys.modules
    mod.__file__ = os.path.join(R.REF, "datasets", "chat", "base_template.py")
    sys.modules["ref_base_template"] = mod
    ns = mod.__dict__
    exec(compile(src, mod.__file__, "exec"), ns)
    duration = 4257 / 29.97002997002997
    out = []
    for llm, T in (("phi3.5", ns["Phi_3_5_Template"]), ("llama3", ns["LLaMA3_Template"]), ("vicuna", ns["Vicuna_Template"])):
        tpl = T()
        for mode, text in (("grounding", "Give you a textual query: \"the man is cooking\". When does the described content occur in the video?"),
                           ("qa", "What is the man doing?"), ("referring", "What happens between 70 seconds and 80 seconds?")):
            if mode == "grounding":
                q = "<image>" + " " + "<timestamp_grounding>" + "\n" + text
            elif mode == "qa":
                q = "<image>" + "\n" + text
            else:
                q = "<image>" + "\n" + re.sub(r"(\d+) seconds", lambda m: f"<{int(float(m.group(1))/duration*300)}>", text)
            conv = [{"from": "human", "value": q}, {"from": "gpt", "value": ""}]
            sep, eos = tpl.separator.apply()
            out.append({"llm": llm, "mode": mode, "text": text, "duration": duration, "prompt": tpl.encode(conv).replace(eos, "")})
    with open(os.path.join(out_dir, "prompts.json"), "w") as fh:
        json.dump(out, fh, indent=1)
    print("wrote prompts.json", len(out))


def golden_ingest(out_dir):
    """the reference's own interpolate_pos_embed_internvideo2_new (internvideo2.py:260-320) on a small synthetic
    checkpoint (orig_t_size 4 -> 8 frames, 4 x 4 spatial grid) -> tests/golden/ingest_pos_embed.npz."""
    import types
    R.import_models()
    iv2 = R._mods["iv2"]
    g = torch.Generator().manual_seed(77)
    c, hw = 24, 16
    ckpt = {"pos_embed": torch.randn(1, 1 + 4 * hw, c, generator=g), "clip_pos_embed": torch.randn(1, 1 + 4 * hw, c, generator=g),
            "img_pos_embed": torch.randn(1, 1 + hw, c, generator=g)}
    src = {k: v.clone() for k, v in ckpt.items()}
    model = types.SimpleNamespace(patch_embed=types.SimpleNamespace(num_patches=8 * hw), pos_embed=torch.zeros(1, 1 + 8 * hw, c),
                                  num_frames=8, tubelet_size=1)
    iv2.interpolate_pos_embed_internvideo2_new(ckpt, model, orig_t_size=4)
    np.savez_compressed(os.path.join(out_dir, "ingest_pos_embed.npz"), **{"in_" + k: v.numpy() for k, v in src.items()},
                        **{"out_" + k: v.numpy() for k, v in ckpt.items()})
    print("ingest_pos_embed: pos_embed", tuple(ckpt["pos_embed"].shape))


def main():
    if not R.available():
        raise SystemExit("reference not found at %s" % R.REF)
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    mods = R.import_models()
    gold_clip(mods)
    gold_iv2(mods)
    gold_phi3(mods)
    gold_llama(mods)
    gold_index_maps()
    gold_host_logic()
    golden_ingest(GOLD)
    golden_prompts(GOLD)


if __name__ == "__main__":
    main()

