"""BASELINE config 4: Llama-3-8B LLaVA-Next backbone, 96 frames (193 visual tokens / segment -> S = 2316 + 63), dense-video-caption
style decode of 256 tokens, 1 GPU. Prints one JSON line (videos/s, stage times, decode GB/s against the 15.01 GB/token weight
stream + 131 KB x ctx of K/V).

    python tools/bench_cfg4.py [--decode 256] [--steps 3]
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "grounded-video-llm_b200"))
from gvl import _lib, hostlogic, model, synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--decode", type=int, default=256)
    ap.add_argument("--steps", type=int, default=3)
    a = ap.parse_args()
    dev = "cuda:0"
    params, lm_cfg, clip_cfg, iv2_cfg = synth.make_params("llama3", device=dev, seed=0)
    m = model.LLAVA_NEXT_VIDEO(params, llm="llama3", lm_cfg=lm_cfg, clip_cfg=clip_cfg, iv2_cfg=iv2_cfg, max_ctx=4096, device=dev)
    del params
    torch.cuda.empty_cache()
    s = synth.make_clip_inputs(1, device=dev)
    for _ in range(2):
        m.generate(s, max_new_tokens=a.decode)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    torch.cuda.synchronize()
    ev[0].record()
    for _ in range(a.steps):
        m.generate(s, max_new_tokens=a.decode)
    ev[1].record()
    torch.cuda.synchronize()
    total = ev[0].elapsed_time(ev[1]) / a.steps
    # stage split on one more pass
    st = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    st[0].record()
    feats = m.encode_images(s)
    st[1].record()
    ids, mask = hostlogic.left_pad(s["input_ids"], 0, 2048)
    emb, _, masks = m.prepare_multimodal_inputs(ids, None, mask, feats, ["v"])
    m.language_model.prefill(emb[0], n_new=a.decode)
    st[2].record()
    m.language_model.generate(inputs_embeds=emb[:1], attention_mask=masks[:1], max_new_tokens=a.decode)
    st[3].record()
    torch.cuda.synchronize()
    enc, pre, both = st[0].elapsed_time(st[1]), st[1].elapsed_time(st[2]), st[2].elapsed_time(st[3])
    S = emb.shape[1]
    dec_ms = (both - pre) / (a.decode - 1)
    byts = 15.01e9 + (S + a.decode / 2.0) * 131072.0
    peak = 6555.2
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        peak = json.load(open(pk))["hbm_gbs"]
    print(json.dumps({"config": "BASELINE configs[3]: Llama-3-8B, 96 frames, S=%d, %d decode tokens, 1 GPU" % (S, a.decode),
                      "videos_per_s": 1e3 / total, "ms_per_video": total, "encode_images_ms": enc, "splice+prefill_ms": pre,
                      "decode_ms_per_token": dec_ms, "decode_gbs": byts / dec_ms / 1e6, "decode_frac_of_hbm_peak": byts / dec_ms / 1e6 / peak,
                      "decode_kernel": "decode_mega_kernel<128> (single persistent kernel)" if _lib.load().gvl_lm_decode_kind(m.language_model._active[0])
                                       else "per-op chain (gemv3_kernel + decode_attn_kernel, CUDA graph)"}))


if __name__ == "__main__":
    main()
