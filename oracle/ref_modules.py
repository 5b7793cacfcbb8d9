"""ORACLE SUPPORT -- TEST INFRASTRUCTURE ONLY. Builds the REFERENCE's own modules (imported through oracle/ref_shims.py
from /root/reference or oracle/_ref) with random-init weights of the named architecture, on any device, and drives the
reference's own `encode_images` / `prepare_multimodal_inputs` on a light stand-in object (LLAVA_NEXT_VIDEO.__init__
hard-requires weight files and a tokenizer directory that do not exist offline).

Used by tests/test_gpu_vs_reference.py (parity against the reference's CUDA bf16 forward), bench.py's reference arm
(`--impl reference`, `extra.gpu_reference`) and oracle/make_golden.py. Never imported by the product.

Construction mirrors llava_next_video.py:113-151: CLIP tower fp32 parameters (run under autocast), InternVideo2 cast to
bf16 with `.to(dtype)` (:134), projectors fp32 parameters, language model in bf16 (`torch_dtype=self.dtype`), lm_head
replaced by a Linear WITH bias and the vocabulary grown by 302 rows (reset_embeddings, :231-268).
"""
import contextlib
import copy

import torch
from torch import nn

from . import ref_shims as R


def _no_tf32():
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False


@contextlib.contextmanager
def no_init():
    """Construct reference modules WITHOUT running their random initialisers (single-threaded RNG: ~2 minutes for the fp32 3.8B
    decoder on the host): torch.nn.init.* and the modules' reset_parameters / _init_weights become no-ops for the duration;
    the caller fills the parameters with fast_fill_() afterwards. Timing infrastructure only (oracle/ref_bench.py)."""
    import torch.nn.init as I
    mods = R.import_models()
    saved = []

    def patch(obj, name, fn):
        saved.append((obj, name, getattr(obj, name)))
        setattr(obj, name, fn)

    ident = lambda t, *a, **k: t
    for name in ("trunc_normal_", "normal_", "uniform_", "xavier_uniform_", "xavier_normal_", "kaiming_uniform_", "kaiming_normal_",
                 "constant_", "zeros_", "ones_"):
        patch(I, name, ident)
    for cls in (nn.Linear, nn.Embedding, nn.Conv2d, nn.Conv3d, nn.LayerNorm):
        patch(cls, "reset_parameters", lambda self: None)
    for key, cname in (("phi3", "Phi3PreTrainedModel"), ("llama", "LlamaPreTrainedModel"), ("clip", "CLIPPreTrainedModel")):
        patch(getattr(mods[key], cname), "_init_weights", lambda self, module: None)
    iv = mods["iv2"]
    patch(iv.PretrainInternVideo2, "_init_weights", lambda self, m: None)
    patch(iv.PretrainInternVideo2, "fix_init_weight", lambda self: None)
    patch(iv, "trunc_normal_", ident)
    try:
        yield
    finally:
        for obj, name, old in reversed(saved):
            setattr(obj, name, old)


def fast_fill_(module, seed=0, std=0.02, threads=None):
    """Seeded normal(0, std) for every matrix / embedding / conv weight, drawn in parallel chunks (one generator per chunk; torch
    releases the GIL), ones for norm weights, zeros for biases and the rest, 1e-5 for LayerScale gammas (as constructed)."""
    import os
    from concurrent.futures import ThreadPoolExecutor
    threads = threads or os.cpu_count() or 1
    jobs = []
    for n, (name, p) in enumerate(module.named_parameters()):
        d = p.data
        if d.dim() >= 2 and "pos" not in name and "cls" not in name:
            flat = d.view(-1)
            step = max(1 << 22, (flat.numel() + threads - 1) // threads)
            for c, i in enumerate(range(0, flat.numel(), step)):
                jobs.append((flat[i:i + step], seed * 1000003 + n * 1009 + c))
        elif name.endswith("gamma"):
            d.fill_(1e-5)
        elif name.endswith("weight") and d.dim() == 1:
            d.fill_(1.0)
        elif d.dim() >= 2:
            jobs.append((d.view(-1), seed * 1000003 + n * 1009))
        else:
            d.zero_()

    def work(job):
        t, sd = job
        t.normal_(0.0, std, generator=torch.Generator().manual_seed(sd))

    with ThreadPoolExecutor(max_workers=threads) as ex:
        list(ex.map(work, jobs))
    return module


def flash_attn_usable(device="cuda"):
    """True if the installed flash_attn wheel runs on this device (2.8.3 may lack sm_100 kernels)."""
    try:
        from flash_attn import flash_attn_func
        q = torch.randn(1, 128, 2, 64, device=device, dtype=torch.bfloat16)
        o = flash_attn_func(q, q, q, causal=True)
        torch.cuda.synchronize()
        return bool(torch.isfinite(o.float()).all())
    except Exception:                                   # noqa: BLE001 -- any failure means "use the eager twin"
        return False


def clip_config(layers=24, dim=1024, heads=16, ffn=4096, image=336):
    from transformers import CLIPVisionConfig
    cfg = CLIPVisionConfig(attention_dropout=0.0, dropout=0.0, hidden_act="quick_gelu", hidden_size=dim, image_size=image,
                           initializer_factor=1.0, initializer_range=0.02, intermediate_size=ffn, layer_norm_eps=1e-5,
                           num_attention_heads=heads, num_channels=3, num_hidden_layers=layers, patch_size=14,
                           projection_dim=768)                       # llava_next_video.py:56-72
    cfg._attn_implementation = "eager"                                # the reference's CLIPAttention is the eager bmm path
    return cfg


def build_clip(state_dict=None, device="cpu", **kw):
    mods = R.import_models()
    with torch.device(device):
        m = mods["clip"].CLIPVisionModel(clip_config(**kw)).eval()
    if state_dict is not None:
        missing, unexpected = m.load_state_dict(state_dict, strict=False)
        assert not unexpected and all("position_ids" in k or "post_layernorm" in k for k in missing), (missing, unexpected)
    return m


def build_iv2(state_dict=None, frames=8, depth=40, flash=False, device="cpu", dim=1408, heads=16, img=224,
              mlp_ratio=48 / 11, dtype=torch.bfloat16):
    """pretrain_internvideo2_1b_patch14_224 (internvideo2.py:1089-1114) with depth / width overridable for small cases."""
    mods = R.import_models()
    with torch.device(device):
        m = mods["iv2"].PretrainInternVideo2(
            in_chans=3, img_size=img, patch_size=14, embed_dim=dim, depth=depth, num_heads=heads, mlp_ratio=mlp_ratio,
            clip_embed_dim=768, attn_pool_num_heads=16, qkv_bias=False, drop_path_rate=0.25, init_values=0.00001,
            qk_normalization=True, use_flash_attn=flash, use_fused_rmsnorm=False, use_fused_mlp=False, fused_mlp_heuristic=1,
            layerscale_no_force_fp32=False, num_frames=frames, tubelet_size=1, sep_pos_embed=False,
            sep_image_video_pos_embed=True, use_checkpoint=False, checkpoint_num=40, clip_teacher_embed_dim=3200,
            clip_teacher_final_dim=768, clip_norm_type="l2", clip_return_layer=6, clip_student_return_interval=1).eval()
    if state_dict is not None:
        missing, unexpected = m.load_state_dict(state_dict, strict=False)
        assert not unexpected, unexpected
        on_path = [k for k in missing if k.startswith(("patch_embed.", "cls_token", "pos_embed", "blocks."))]
        assert not on_path, on_path
    return m.to(dtype)                                               # llava_next_video.py:134


def phi3_config(layers=32, dim=3072, heads=32, kv_heads=32, ffn=8192, vocab=32064 + 302, rope=None, max_pos=131072,
                orig_max_pos=4096, attn="eager"):
    mods = R.import_models()
    cfg = mods["phi3"].Phi3Config(
        vocab_size=vocab, hidden_size=dim, intermediate_size=ffn, num_hidden_layers=layers, num_attention_heads=heads,
        num_key_value_heads=kv_heads, rms_norm_eps=1e-5, max_position_embeddings=max_pos,
        original_max_position_embeddings=orig_max_pos, rope_theta=10000.0, sliding_window=None, attention_dropout=0.0,
        resid_pdrop=0.0, embd_pdrop=0.0, pad_token_id=0, bos_token_id=1, eos_token_id=2)
    if rope is not None:                                             # transformers 5.x normalises rope_scaling; set after
        cfg.rope_scaling = {"type": "longrope", "short_factor": list(rope["short_factor"]),
                            "long_factor": list(rope["long_factor"])}
    else:
        cfg.rope_scaling = None
    cfg._attn_implementation = attn
    return cfg


def llama_config(layers=32, dim=4096, heads=32, kv_heads=8, ffn=14336, vocab=128256 + 302, theta=500000.0, max_pos=8192,
                 attn="eager"):
    from transformers import LlamaConfig
    cfg = LlamaConfig(vocab_size=vocab, hidden_size=dim, intermediate_size=ffn, num_hidden_layers=layers,
                      num_attention_heads=heads, num_key_value_heads=kv_heads, rms_norm_eps=1e-5,
                      max_position_embeddings=max_pos, attention_bias=False, attention_dropout=0.0, pad_token_id=0,
                      bos_token_id=1, eos_token_id=2)
    cfg.rope_theta, cfg.rope_scaling, cfg.pretraining_tp, cfg.attention_bias, cfg.mlp_bias = theta, None, 1, False, False
    cfg._attn_implementation = attn
    return cfg


def build_lm(arch, cfg, state_dict=None, device="cpu", dtype=torch.bfloat16):
    """Phi3ForCausalLM / LlamaForCausalLM after reset_embeddings (lm_head = Linear(D, V, bias=True))."""
    mods = R.import_models()
    cls = mods["phi3"].Phi3ForCausalLM if arch == "phi3" else mods["llama"].LlamaForCausalLM
    prev = torch.get_default_dtype()
    torch.set_default_dtype(dtype)
    try:
        with torch.device(device):
            m = cls(cfg).eval()
            m.lm_head = nn.Linear(cfg.hidden_size, cfg.vocab_size, bias=True)
    finally:
        torch.set_default_dtype(prev)
    if state_dict is not None:
        missing, unexpected = m.load_state_dict(state_dict, strict=False)
        assert not unexpected and all("rotary_emb" in k for k in missing), (missing, unexpected)
    return m.to(dtype)


def use_flash_attention(m):
    """Switch a constructed Phi3ForCausalLM / LlamaForCausalLM to the reference's own FlashAttention2 classes
    (modeling_phi3.py:613-920, modeling_llama.py:402-594) -- what `attn_implementation="flash_attention_2"` selects in the
    reference (llava_next_video.py:146-148). Done after construction because transformers 5.x refuses the string at
    __init__ for model classes that only carry the old `_supports_flash_attn_2` attribute. The FlashAttention2 classes
    subclass the eager ones and add no parameters, so re-classing the modules is exact."""
    mods = R.import_models()
    for sub in m.modules():
        if type(sub) is mods["phi3"].Phi3Attention:
            sub.__class__ = mods["phi3"].Phi3FlashAttention2
            sub._flash_attn_uses_top_left_mask = False           # flash_attn >= 2.1 (modeling_phi3.py:625-627)
        elif type(sub) is mods["llama"].LlamaAttention:
            sub.__class__ = mods["llama"].LlamaFlashAttention2
            sub._flash_attn_uses_top_left_mask = False
        if isinstance(getattr(sub, "_attn_implementation", None), str) and not hasattr(type(sub), "_attn_implementation"):
            sub._attn_implementation = "flash_attention_2"       # Phi3Model keeps its own copy (modeling_phi3.py:1235)
    m.config._attn_implementation_internal = "flash_attention_2"  # LlamaModel reads the config (modeling_llama.py:1047)
    return m


class RefVLM:
    """Stand-in for a constructed LLAVA_NEXT_VIDEO: holds the reference's sub-modules and exposes the reference's own
    encode_images / prepare_multimodal_inputs / reshape_hd_patches_2x2merge_phi3 / add_image_newline_phi3, extracted
    from models/llava_next_video.py with `ast` (ref_shims.extract) and bound to this object."""

    def __init__(self, llm, vision_tower, video_encoder, multi_modal_projector, video_projecter, language_model, extras,
                 device, dtype=torch.bfloat16):
        self.llm, self.device, self.dtype = llm, torch.device(device), dtype
        self.vision_tower, self.video_encoder = vision_tower, video_encoder
        self.multi_modal_projector, self.video_projecter = multi_modal_projector, video_projecter
        self.language_model = language_model
        for k, v in extras.items():
            setattr(self, k, v)
        self.config = type("C", (), {"hidden_size": language_model.config.hidden_size if language_model is not None else
                                     video_projecter.down_proj.weight.shape[0]})()
        ns = R.std_namespace()
        ns["math"] = __import__("math")
        f = "models/llava_next_video.py"
        for name in ("encode_images", "prepare_multimodal_inputs", "reshape_hd_patches_2x2merge_phi3",
                     "add_image_newline_phi3"):
            fn = R.extract(f, name, "LLAVA_NEXT_VIDEO", ns)
            setattr(self, name, fn.__get__(self))

    def get_input_embeddings(self):
        return self.language_model.get_input_embeddings()

    def autocast(self):
        if self.dtype == torch.float32 or self.device.type == "cpu":
            return contextlib.nullcontext()
        return torch.autocast("cuda", dtype=self.dtype)

    def modules(self):
        return [self.vision_tower, self.video_encoder, self.multi_modal_projector, self.video_projecter,
                self.language_model]

    def float_copy(self, parts=("vision_tower", "video_encoder", "multi_modal_projector", "video_projecter", "language_model")):
        """The same parameter VALUES run in fp32 without autocast (TF32 off) and with the eager attention twins (fp32 cannot
        run FlashAttention): the 'exact' side of the three-number comparison. InternVideo2 / LM parameters stay the
        bf16-rounded values the reference holds. Modules not named in `parts` are left out (None) to bound memory."""
        _no_tf32()
        c = copy.copy(self)
        c.dtype = torch.float32
        for name in ("vision_tower", "video_encoder", "multi_modal_projector", "video_projecter", "language_model"):
            m = getattr(self, name)
            setattr(c, name, fp32_eager(m) if (m is not None and name in parts) else None)
        for name in ("glb_GN", "sub_GN", "image_newline"):
            if hasattr(self, name):
                setattr(c, name, getattr(self, name).float())
        for name in ("encode_images", "prepare_multimodal_inputs", "reshape_hd_patches_2x2merge_phi3",
                     "add_image_newline_phi3"):
            setattr(c, name, getattr(self, name).__func__.__get__(c))
        return c


def fp32_eager(module):
    """Deep copy of a reference module in fp32 with every attention switched to the reference's own eager twin
    (internvideo2.py:564-583 `_naive_attn`; Phi3Attention / LlamaAttention, which the FlashAttention2 classes subclass)."""
    mods = R.import_models()
    m = copy.deepcopy(module).float()
    for sub in m.modules():
        if hasattr(sub, "use_flash_attn"):
            sub.use_flash_attn = False
        if isinstance(sub, mods["phi3"].Phi3Attention):
            sub.__class__ = mods["phi3"].Phi3Attention
        if isinstance(sub, mods["llama"].LlamaAttention):
            sub.__class__ = mods["llama"].LlamaAttention
        if hasattr(sub, "_attn_implementation") and isinstance(getattr(sub, "_attn_implementation"), str):
            sub._attn_implementation = "eager"
    if hasattr(m, "config") and hasattr(m.config, "_attn_implementation"):
        m.config = copy.deepcopy(m.config)
        m.config._attn_implementation = "eager"
        for sub in m.modules():
            if hasattr(sub, "config") and sub is not m and getattr(sub.config, "_attn_implementation", None) is not None:
                sub.config = m.config
    return m


def build_vlm(params, llm="phi3.5", lm_cfg=None, frames_per_seg=8, clip_kw=None, iv2_kw=None, lm_kw=None, device="cpu",
              flash=False, with_lm=True):
    """params: the dict gvl.synth.make_params returns (reference state_dict names). Returns a RefVLM whose modules hold
    exactly those values (loaded with load_state_dict, so a naming mismatch fails here)."""
    mods = R.import_models()
    ns = R.std_namespace()
    src_cls = {}
    for name in ("Phi3_5_Projecter", "Video_Projecter"):
        src_cls[name] = _extract_class("models/llava_next_video.py", name, ns)
    vt = build_clip(params["vision_tower"], device=device, **(clip_kw or {}))
    ve = build_iv2(params["video_encoder"], frames=frames_per_seg, flash=flash, device=device, **(iv2_kw or {}))
    lm = None
    D = params["video_projecter"]["down_proj.weight"].shape[0]
    if with_lm:
        arch = "phi3" if llm == "phi3.5" else "llama"
        kw = dict(lm_kw or {})
        cfg = phi3_config(rope=lm_cfg["rope"], **kw) if arch == "phi3" else llama_config(**kw)
        lm = build_lm(arch, cfg, params["language_model"], device=device)
        if flash:
            use_flash_attention(lm)
    with torch.device(device):
        vp = src_cls["Video_Projecter"](params["video_projecter"]["up_proj.weight"].shape[1], D).eval()
        vp.load_state_dict(params["video_projecter"])
        extras = {}
        if llm == "phi3.5":
            mm = src_cls["Phi3_5_Projecter"]().eval()
            if mm.linear_0.weight.shape != params["multi_modal_projector"]["linear_0.weight"].shape:
                w0, w1 = params["multi_modal_projector"]["linear_0.weight"], params["multi_modal_projector"]["linear_1.weight"]
                mm.linear_0 = nn.Linear(w0.shape[1], w0.shape[0])
                mm.linear_1 = nn.Linear(w1.shape[1], w1.shape[0])
            mm.load_state_dict(params["multi_modal_projector"])
            extras["glb_GN"] = params["glb_GN"].to(device).float()
            extras["sub_GN"] = params["sub_GN"].to(device).float()
        else:
            mm = _LlavaProjector(params["multi_modal_projector"]).eval()
            extras["image_newline"] = params["image_newline"].to(device, torch.bfloat16)
    return RefVLM(llm, vt, ve, mm.to(device), vp.to(device), lm, extras, device)


class _LlavaProjector(nn.Module):
    """transformers' LlavaMultiModalProjector (third-party, HF llava; used at llava_next_video.py:138): Linear -> GELU ->
    Linear with biases, parameter names linear_1 / linear_2."""

    def __init__(self, sd):
        super().__init__()
        w1, w2 = sd["linear_1.weight"], sd["linear_2.weight"]
        self.linear_1 = nn.Linear(w1.shape[1], w1.shape[0])
        self.act = nn.GELU()
        self.linear_2 = nn.Linear(w2.shape[1], w2.shape[0])
        self.load_state_dict(sd)

    def forward(self, x):
        return self.linear_2(self.act(self.linear_1(x)))


def _extract_class(rel_path, name, namespace):
    import ast
    import os
    path = os.path.join(R.REF, rel_path)
    tree = ast.parse(open(path).read())
    cls = [n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == name][0]
    ns = dict(namespace)
    exec(compile(ast.Module(body=[cls], type_ignores=[]), path, "exec"), ns)
    return ns[name]


def ShimCache():
    """The slice of transformers==4.40.1's DynamicCache the reference's decoders call (modeling_phi3.py:562-569, 680-721,
    1288-1292; modeling_llama.py:969-972), restated because transformers 5.x removed `get_usable_length` /
    `from_legacy_cache` / `to_legacy_cache`. Patched onto transformers' class in the TEST process only."""
    from transformers.cache_utils import DynamicCache
    if not hasattr(DynamicCache, "get_usable_length"):
        DynamicCache.get_usable_length = lambda self, new_seq_length, layer_idx=0: self.get_seq_length(layer_idx)
    if not hasattr(DynamicCache, "to_legacy_cache"):
        DynamicCache.to_legacy_cache = lambda self: self
    if not hasattr(DynamicCache, "from_legacy_cache"):
        DynamicCache.from_legacy_cache = classmethod(lambda cls, past=None: past if isinstance(past, DynamicCache) else cls())
    return DynamicCache()


@torch.no_grad()
def greedy_generate(lm, inputs_embeds, max_new_tokens, eos_token_id=None, pad_token_id=0):
    """HF greedy search restated (GenerationMixin.generate, transformers==4.40.1; call site llava_next_video.py:655-661):
    step 0 feeds inputs_embeds, later steps feed the last token with the KV cache; stops at EOS, pads after it.
    Returns (tokens int64 [n], logits fp32 [n, V]) for ONE unpadded sequence."""
    cache = ShimCache()
    emb = inputs_embeds
    toks, logs = [], []
    done = False
    for t in range(max_new_tokens):
        out = lm(inputs_embeds=emb, past_key_values=cache, use_cache=True, return_dict=True)
        cache = out.past_key_values
        lg = out.logits[:, -1].float()
        nxt = lg.argmax(-1)
        if done:
            nxt = torch.full_like(nxt, pad_token_id)
        toks.append(nxt)
        logs.append(lg[0])
        if eos_token_id is not None and int(nxt) == eos_token_id:
            done = True
        emb = lm.get_input_embeddings()(nxt)[:, None]
    return torch.stack(toks, 1)[0], torch.stack(logs, 0)
