"""Checkpoint ingest: the reference's on-disk layout -> the `params` dictionary gvl.model.LLAVA_NEXT_VIDEO takes.

Load-time host logic only (torch is the container format of the checkpoints); nothing here runs on the hot path.
Mirrors, step by step:
  * `LLAVA_NEXT_VIDEO.__init__` (llava_next_video.py:100-151): vision_model.pth, image_newlines.pth / image_newline.pth,
    multi_modal_projector.pth, the InternVideo2 checkpoint with its temporal position-embedding interpolation 4 -> T
    (`interpolate_pos_embed_internvideo2_new`, internvideo2.py:260-320), language_model_seperated/ (HF directory);
  * `reset_embeddings` (llava_next_video.py:231-268): +302 vocabulary rows (`<0>`..`<300>`, `<timestamp_grounding>`) initialised
    to the mean row, new lm_head WITH bias;
  * `lora_model` (llava_next_video.py:212-229; third-party peft==0.3.0): state-dict key layout of a PEFT-wrapped causal LM and
    y = W x + (lora_alpha / r) * B (A x), alpha / r = 256 / 128 = 2 -- merged into W at load;
  * `inference.py:156-162`: the fine-tuned `ckpt['model']` sub-dictionaries override projectors and language model.
"""
import os
import re

import torch

NUM_TEMPORAL_TOKENS = 300
LORA_SCALE = 256.0 / 128.0                       # lora_alpha / r (llava_next_video.py:216, 221)


# ----------------------------------------------------------------------------------------------- InternVideo2 pos_embed
def interpolate_pos_embed_temporal(pos_embed, orig_t_size, new_t_size, num_extra_tokens=1):
    """internvideo2.py:290-303: the position tokens are interpolated linearly along T, every (patch, channel) series on its
    own; class token untouched. pos_embed: [1, extra + T0 * HW, C] -> [1, extra + T1 * HW, C]. The interpolation itself is
    torch's `interpolate(mode='linear', align_corners=False)` (src = (i + 0.5) * T0 / T1 - 0.5 clamped at 0, two-tap blend) --
    called rather than re-derived so that the loaded table is bit-identical to the reference's (its CPU kernel fuses the blend
    into an FMA, which a hand-written `w0 * x0 + w1 * x1` does not reproduce: 2.4e-7 off)."""
    if orig_t_size == new_t_size:
        return pos_embed
    c = pos_embed.shape[-1]
    extra = pos_embed[:, :num_extra_tokens]
    series = pos_embed[:, num_extra_tokens:].reshape(orig_t_size, -1, c).permute(1, 2, 0)        # [HW, C, T0]
    series = torch.nn.functional.interpolate(series, size=new_t_size, mode="linear")               # [HW, C, T1]
    tokens = series.permute(2, 0, 1).reshape(1, -1, c)                                             # [1, T1 * HW, C]
    return torch.cat((extra, tokens), dim=1)


def interpolate_pos_embed_internvideo2(state_dict, num_frames, tubelet_size=1, num_patches_per_frame=256, orig_t_size=4):
    """In-place on the checkpoint dict, like the reference: every '*pos_embed*' key except 'img_pos_embed'. Only the temporal
    branch is implemented: the shipped checkpoint and the model share the 16 x 16 spatial grid (orig_size == new_size)."""
    new_t = num_frames // tubelet_size
    names = [k for k in state_dict if ("pos_embed" in k or "clip_pos_embed" in k) and "img_pos_embed" not in k]
    if not names:
        raise KeyError("no pos_embed in the InternVideo2 checkpoint")
    if "pos_embed_spatial" in state_dict or "pos_embed_temporal" in state_dict:
        raise NotImplementedError
    for k in names:
        pe = state_dict[k]
        extra = pe.shape[-2] - orig_t_size * num_patches_per_frame
        if extra not in (0, 1):
            raise ValueError("%s: %d tokens is not extra + %d x %d" % (k, pe.shape[-2], orig_t_size, num_patches_per_frame))
        state_dict[k] = interpolate_pos_embed_temporal(pe, orig_t_size, new_t, extra)
    return state_dict


# ----------------------------------------------------------------------------------------------- vocabulary extension
def reset_embeddings(embed_weight, lm_head_weight, num_new_tokens=NUM_TEMPORAL_TOKENS + 2, lm_head_bias=None):
    """llava_next_video.py:231-268 on raw tensors: new rows = the mean row (computed in the weights' dtype, like torch.mean on
    the module weight); the new lm_head has a bias (nn.Linear default), which only a checkpoint gives meaningful values to --
    zeros are used when none is supplied."""
    e_new = torch.cat([embed_weight, torch.mean(embed_weight, dim=0)[None].expand(num_new_tokens, -1)], dim=0)
    h_new = torch.cat([lm_head_weight, torch.mean(lm_head_weight, dim=0)[None].expand(num_new_tokens, -1)], dim=0)
    if lm_head_bias is None:
        lm_head_bias = torch.zeros(h_new.shape[0], dtype=h_new.dtype)
    return e_new.contiguous(), h_new.contiguous(), lm_head_bias


def temporal_token_ids(tokenizer_len_before_add):
    """`tokenizer.add_tokens` appends: id(<k>) = len(tokenizer) + k, id(<timestamp_grounding>) = len(tokenizer) + 301 (SURVEY L9:
    the ids start at len(tokenizer), not at config.vocab_size)."""
    base = int(tokenizer_len_before_add)
    return {"<%d>" % k: base + k for k in range(NUM_TEMPORAL_TOKENS + 1)} | {"<timestamp_grounding>": base + NUM_TEMPORAL_TOKENS + 1}


# ----------------------------------------------------------------------------------------------- PEFT / LoRA
_PEFT_PREFIX = "base_model.model."


def merge_lora(state_dict, scale=LORA_SCALE, adapter="default"):
    """State dict of a PEFT-wrapped LM (keys 'base_model.model.<hf name>', '<module>.lora_A.<adapter>.weight',
    '<module>.lora_B.<adapter>.weight', newer peft: '<module>.base_layer.weight') -> plain HF state dict with
    W' = W + scale * B @ A (accumulated in float32, stored in W's dtype). A dict without LoRA keys passes through."""
    plain, lora_a, lora_b = {}, {}, {}
    for k, v in state_dict.items():
        if k.startswith(_PEFT_PREFIX):
            k = k[len(_PEFT_PREFIX):]
        m = re.match(r"(.*)\.lora_([AB])\.(?:%s\.)?weight$" % re.escape(adapter), k)
        if m:
            (lora_a if m.group(2) == "A" else lora_b)[m.group(1)] = v
            continue
        if ".lora_" in k:                     # lora_dropout, other adapters, embedding adapters: not used by the reference
            continue
        plain[k.replace(".base_layer.", ".")] = v
    if set(lora_a) != set(lora_b):
        raise KeyError("unpaired LoRA factors: %s" % sorted(set(lora_a) ^ set(lora_b)))
    for mod, a in lora_a.items():
        wk = mod + ".weight"
        if wk not in plain:
            raise KeyError("LoRA factors for %s but no base weight" % mod)
        w = plain[wk]
        plain[wk] = (w.float() + scale * (lora_b[mod].float() @ a.float())).to(w.dtype)
    return plain


# ----------------------------------------------------------------------------------------------- files -> params
def _load_pth(path):
    return torch.load(path, map_location="cpu", weights_only=True)


def load_hf_language_model(directory):
    """language_model_seperated/: *.safetensors shards (or pytorch_model*.bin) -> one state dict."""
    sd = {}
    names = sorted(os.listdir(directory))
    st = [n for n in names if n.endswith(".safetensors")]
    if st:
        from safetensors.torch import load_file
        for n in st:
            sd.update(load_file(os.path.join(directory, n)))
        return sd
    bins = [n for n in names if re.match(r"pytorch_model.*\.bin$", n)]
    if not bins:
        raise FileNotFoundError("no safetensors / pytorch_model*.bin under %s" % directory)
    for n in bins:
        sd.update(_load_pth(os.path.join(directory, n)))
    return sd


def load_params(llm, pretrained_vision_proj_llm_path, pretrained_video_path, ckpt_path=None, num_frames=96, num_segs=12):
    """The reference's construction + checkpoint loading (llava_next_video.py:100-151, inference.py:156-162) as one function.
    Returns the `params` dict of gvl.model.LLAVA_NEXT_VIDEO: vision_tower / video_encoder / multi_modal_projector /
    video_projecter / language_model state dicts (+ sub_GN, glb_GN for phi3.5, image_newline otherwise)."""
    root = pretrained_vision_proj_llm_path
    params = {"vision_tower": _load_pth(os.path.join(root, "vision_model.pth"))}
    if llm == "phi3.5":
        nl = _load_pth(os.path.join(root, "image_newlines.pth"))
        params["glb_GN"], params["sub_GN"] = nl["glb_GN"], nl["sub_GN"]
    else:
        params["image_newline"] = _load_pth(os.path.join(root, "image_newline.pth"))["image_newline"]
    video = _load_pth(pretrained_video_path)
    interpolate_pos_embed_internvideo2(video, num_frames // num_segs, orig_t_size=4)
    params["video_encoder"] = video
    params["multi_modal_projector"] = _load_pth(os.path.join(root, "multi_modal_projector.pth"))
    lm = load_hf_language_model(os.path.join(root, "language_model_seperated"))
    params["video_projecter"] = None
    if ckpt_path is not None:
        ckpt = _load_pth(ckpt_path)["model"]
        if "multi_modal_projector" in ckpt:
            params["multi_modal_projector"] = ckpt["multi_modal_projector"]
        if "video_projecter" in ckpt:
            params["video_projecter"] = ckpt["video_projecter"]
        if "language_model" in ckpt:
            # the fine-tuned LM is the PEFT-wrapped, vocabulary-extended model: its dict carries embed_tokens / lm_head (+bias) with
            # 302 extra rows and the LoRA factors; everything else falls back to the base weights
            tuned = merge_lora(ckpt["language_model"])
            base = merge_lora(lm)
            base.update(tuned)
            lm = base
    if params["video_projecter"] is None:
        raise KeyError("video_projecter weights come from the training checkpoint (inference.py:159-160): pass ckpt_path")
    params["language_model"] = lm
    return params
