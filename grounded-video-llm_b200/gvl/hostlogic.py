"""Host-side integer / string logic of the path, kept call-compatible with the reference:
tokenizer_image_token, left-padding, temporal-token <-> seconds arithmetic, frame-index sampling and the RoPE
tables (built once at load; fp32 trig then cast to bf16 exactly as the reference does per forward).
Bit-exactness matters here: the temporal-token expressions are evaluated literally in IEEE double, in the
reference's own operation order.
"""
import math
import re

import numpy as np
import torch

IMAGE_TOKEN_INDEX = -200
IGNORE_INDEX = -100
DEFAULT_IMAGE_TOKEN = "<image>"
GROUNDING_TOKEN = "<timestamp_grounding>"


def tokenizer_image_token(prompt, tokenizer, image_token_index=IMAGE_TOKEN_INDEX, return_tensors=None, device="cpu"):
    """llava_next_video.py:409-426."""
    prompt_chunks = [tokenizer(chunk).input_ids for chunk in prompt.split(DEFAULT_IMAGE_TOKEN)]
    input_ids = []
    offset = 0
    if len(prompt_chunks) > 0 and len(prompt_chunks[0]) > 0 and prompt_chunks[0][0] == tokenizer.bos_token_id:
        offset = 1
        input_ids.append(prompt_chunks[0][0])
    sep = [image_token_index] * (offset + 1)
    pieces = []
    for i, chunk in enumerate(prompt_chunks):
        pieces.append(chunk)
        if i + 1 < len(prompt_chunks):
            pieces.append(sep)
    for x in pieces:
        input_ids.extend(x[offset:])
    if return_tensors is not None:
        if return_tensors == "pt":
            return torch.tensor(input_ids, dtype=torch.long, device=device)
        raise ValueError(f"Unsupported tensor type: {return_tensors}")
    return input_ids


def left_pad(id_lists, pad_id, max_txt_len):
    """generate() pre-amble, llava_next_video.py:622-647 (flip / pad / truncate / flip)."""
    L = min(max(len(x) for x in id_lists), max_txt_len)
    ids = torch.full((len(id_lists), L), pad_id, dtype=torch.long)
    mask = torch.zeros((len(id_lists), L), dtype=torch.long)
    for r, x in enumerate(id_lists):
        x = list(x)
        if len(x) > L:
            x = x[len(x) - L:]
        ids[r, L - len(x):] = torch.tensor(x, dtype=torch.long)
        mask[r, L - len(x):] = 1
    return ids, mask


def parse_time_interval(text, duration, num_temporal_tokens=300, llm="phi3.5"):
    """inference.py:125-134:  seconds = duration * k / num_temporal_tokens, formatted %.2f."""
    pattern = r"<(\d+)>"

    def replace_func(match):
        x = int(match.group(1))
        m = duration * x / num_temporal_tokens
        if llm == "phi3.5":
            return f" {m:.2f} seconds"
        elif llm == "llama3":
            return f"{m:.2f} seconds"
        return None

    return re.sub(pattern, replace_func, text)


def seconds_to_token_inference(query, duration, num_temporal_tokens=300):
    """inference.py:107:  k = int(float(sec) / duration * num_temporal_tokens)."""
    return re.sub(r"(\d+) seconds", lambda m: f"<{int(float(m.group(1)) / duration * num_temporal_tokens)}>", query)


def seconds_to_token_training(time, duration, num_temporal_tokens=300):
    """datasets/mix_grounded.py:78-91:  k = min(int(num_temporal_tokens * time / duration), num_temporal_tokens)."""
    return min(int(num_temporal_tokens * time / duration), num_temporal_tokens)


def get_frame_indices(num_frames, vlen, sample="middle"):
    """mm_utils/video_utils.py:13-51 (the 'middle' branch used by inference.py:71-75)."""
    if sample != "middle":
        raise NotImplementedError("only sample='middle' is on the inference path")
    acc_samples = min(num_frames, vlen)
    intervals = np.linspace(start=0, stop=vlen, num=acc_samples + 1).astype(int)
    frame_indices = [int((intervals[i] + intervals[i + 1] - 1) // 2) for i in range(acc_samples)]
    if len(frame_indices) < num_frames:
        frame_indices = frame_indices + [frame_indices[-1]] * (num_frames - len(frame_indices))
    return frame_indices


def spatial_keyframes(num_frames, num_segs):
    """inference.py:81-83."""
    per = int(num_frames // num_segs)
    return [(i * per) + int(per / 2) for i in range(num_segs)]


# --------------------------------------------------------------------------------------- RoPE tables
def longrope_tables(max_ctx, head_dim, base, short_factor, long_factor, max_pos, orig_max_pos, use_long, long_from=None):
    """Phi3LongRoPEScaledRotaryEmbedding.forward (modeling_phi3.py:371-409) for positions 0..max_ctx-1 -> bf16.

    The factor set is chosen per FORWARD CALL from `seq_len = kv_seq_len` (modeling_phi3.py:562-563, 680-686): a prefill of S
    tokens rotates every position with long_factor iff S > original_max_position_embeddings; a cached decode step at
    position p (kv_seq_len = p + 1) rotates ITS q / k with long_factor iff p >= original_max_position_embeddings, while the
    keys already in the cache keep the rotation they were stored with. On the reference's inference path (generate from
    inputs_embeds) the cache reset of prepare_inputs_for_generation (:1557-1562) never fires, because its `input_ids` holds
    only the generated tokens. Hence one table serves a whole generate call: rows < long_from short, rows >= long_from long,
    with long_from = 0 for a long prefill (use_long) and original_max_position_embeddings otherwise."""
    if long_from is None:
        long_from = 0 if use_long else max_ctx
    shape = torch.arange(0, head_dim, 2, dtype=torch.int64).float() / head_dim
    pos = torch.arange(max_ctx, dtype=torch.float32)
    scale = max_pos / orig_max_pos
    sf = 1.0 if scale <= 1.0 else math.sqrt(1 + math.log(scale) / math.log(orig_max_pos))
    out = []
    for factor in (short_factor, long_factor):
        inv_freq = 1.0 / (torch.tensor(factor, dtype=torch.float32) * base ** shape)
        freqs = pos[:, None] * inv_freq[None, :]
        emb = torch.cat((freqs, freqs), dim=-1)
        out.append(((emb.cos() * sf).to(torch.bfloat16), (emb.sin() * sf).to(torch.bfloat16)))
    is_long = (torch.arange(max_ctx) >= long_from)[:, None]
    return torch.where(is_long, out[1][0], out[0][0]), torch.where(is_long, out[1][1], out[0][1])


def plain_rope_tables(max_ctx, head_dim, base, bf16_matmul_quirk=False):
    """Phi3RotaryEmbedding (modeling_phi3.py:345-368) / LlamaRotaryEmbedding (modeling_llama.py:94-133).
    bf16_matmul_quirk reproduces the reference's CUDA-autocast behaviour for Llama: the inv_freq @ position matmul
    is not excluded from autocast there, so both operands and the product are bf16."""
    inv_freq = 1.0 / (base ** (torch.arange(0, head_dim, 2, dtype=torch.int64).float() / head_dim))
    pos = torch.arange(max_ctx, dtype=torch.float32)
    if bf16_matmul_quirk:
        freqs = (pos.to(torch.bfloat16).float()[:, None] * inv_freq.to(torch.bfloat16).float()[None, :]).to(torch.bfloat16)
        emb = torch.cat((freqs, freqs), dim=-1)
        return emb.float().cos().to(torch.bfloat16), emb.float().sin().to(torch.bfloat16)
    freqs = pos[:, None] * inv_freq[None, :]
    emb = torch.cat((freqs, freqs), dim=-1)
    return emb.cos().to(torch.bfloat16), emb.sin().to(torch.bfloat16)


def warp_logits(scores, temperature=None, top_k=50, top_p=None, min_tokens_to_keep=1):
    """HF logits warpers in GenerationMixin._get_logits_warper order (third-party transformers==4.40.1, used by the reference's
    `.generate(do_sample=True, temperature=0.2, top_p=None)`, inference.py:170-176): TemperatureLogitsWarper (scores / T),
    TopKLogitsWarper (keep the k largest; ties with the k-th value survive), TopPLogitsWarper (drop the ascending-sorted tail
    whose cumulative probability is <= 1 - top_p, always keep the last `min_tokens_to_keep`). scores: float [B, V]."""
    import torch
    if temperature is not None and temperature != 1.0:
        scores = scores / temperature
    if top_k is not None and top_k > 0:
        k = min(max(int(top_k), min_tokens_to_keep), scores.shape[-1])
        kth = torch.topk(scores, k)[0][..., -1, None]
        scores = scores.masked_fill(scores < kth, -float("inf"))
    if top_p is not None and top_p < 1.0:
        sorted_logits, sorted_indices = torch.sort(scores, descending=False)
        cumulative = sorted_logits.softmax(dim=-1).cumsum(dim=-1)
        remove_sorted = cumulative <= (1 - top_p)
        remove_sorted[..., -min_tokens_to_keep:] = False
        remove = remove_sorted.scatter(1, sorted_indices, remove_sorted)
        scores = scores.masked_fill(remove, -float("inf"))
    return scores


# ----------------------------------------------------------------------------------------------- prompts (inference.py:90-116)
# One-turn prompt of the reference's chat templates (datasets/chat/base_template.py:114-140) with an empty assistant slot and the
# template's end-of-turn string removed, as create_inputs does (`chat_template.encode(conv).replace(eos, '')`).
# (system, user prefix, assistant prefix, eos)
_TEMPLATES = {
    "phi3.5": ("<|system|>\nYou are a helpful AI assistant that can generate responses based on visual inputs.",
               "\n<|user|>\n", "\n<|assistant|>\n", "<|endoftext|>"),
    "llama3": ("<|start_header_id|>system<|end_header_id|>You are a helpful language and vision assistant. You are able to understand "
               "the visual content that the user provides, and assist the user with a variety of tasks using natural language.",
               "<|start_header_id|>user<|end_header_id|>", "<|start_header_id|>assistant<|end_header_id|>", "<|eot_id|>"),
    "vicuna": ("You are a helpful language and vision assistant. You are able to understand the visual content that the user "
               "provides, and assist the user with a variety of tasks using natural language.",
               "\nUSER: ", "\nASSISTANT: ", "</s>"),
}


def build_prompt(llm, mode, text, duration=None, num_temporal_tokens=300):
    """mode 'grounding' | 'qa' | 'referring' (inference.py:93-110). For 'referring' the "<n> seconds" mentions of `text` are
    quantised to temporal tokens with the inference-side expression (inference.py:107)."""
    system, user, assistant, eos = _TEMPLATES[llm]
    if mode == "grounding":
        question = DEFAULT_IMAGE_TOKEN + " " + GROUNDING_TOKEN + "\n" + text
    elif mode == "qa":
        question = DEFAULT_IMAGE_TOKEN + "\n" + text
    elif mode == "referring":
        question = DEFAULT_IMAGE_TOKEN + "\n" + seconds_to_token_inference(text, duration, num_temporal_tokens)
    else:
        raise ValueError("mode must be grounding, qa or referring")
    if DEFAULT_IMAGE_TOKEN in question and GROUNDING_TOKEN not in question:
        # Template._prompt re-inserts the image token in front of the stripped question (base_template.py:106-108)
        question = (DEFAULT_IMAGE_TOKEN + "\n" + question.replace(DEFAULT_IMAGE_TOKEN, "").strip()).strip()
    prompt = system + user + question + assistant + "" + eos
    return prompt.replace(eos, "")
