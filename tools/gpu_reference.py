"""The GPU "kernel-to-beat" (SURVEY 8d last line): the REFERENCE's own modules (oracle/_ref), bf16
autocast, torch eager + cuBLAS + FlashAttention-2 (when the wheel runs on this device), cfg2, on the same B200.
Times encode_images, splice + prefill (KV-cached forward) and 16 greedy decode steps with CUDA events.

    python tools/gpu_reference.py [--reps 5] [--new 16] [--out gpurun_out/r2_gpu_reference.json]
Test / measurement infrastructure: uses oracle/, never imported by the product."""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "grounded-video-llm_b200"))


def measure(reps=5, n_new=16, llm="phi3.5", num_frames=96, num_segs=12, seed=0, batch=1):
    from gvl import synth
    from oracle import ref_modules as RM
    dev = "cuda"
    flash = RM.flash_attn_usable()
    params, lm_cfg, _, _ = synth.make_params(llm, device=dev, seed=seed, frames_per_seg=num_frames // num_segs)
    ref = RM.build_vlm(params, llm, lm_cfg, frames_per_seg=num_frames // num_segs, device=dev, flash=flash)
    del params
    torch.cuda.empty_cache()
    s = synth.make_clip_inputs(batch, num_frames, num_segs, device=dev)
    ids = torch.tensor(s["input_ids"][0])[None].to(dev).repeat(batch, 1)
    samples = {"spatial_pixel_values": s["spatial_pixel_values"], "temporal_pixel_values": s["temporal_pixel_values"]}
    ev = lambda: torch.cuda.Event(enable_timing=True)
    t_enc, t_pre, t_dec = [], [], []
    toks = None
    with torch.inference_mode(), ref.autocast():
        for r in range(reps + 2):
            e = [ev() for _ in range(4)]
            e[0].record()
            feats = ref.encode_images(samples)
            e[1].record()
            emb, _, mask = ref.prepare_multimodal_inputs(ids, ids.clone(), torch.ones_like(ids), feats, ["video"] * batch)
            cache = RM.ShimCache()
            out = ref.language_model(inputs_embeds=emb, past_key_values=cache, use_cache=True, return_dict=True)
            nxt = out.logits[:, -1].float().argmax(-1)
            e[2].record()
            got = [nxt]
            for t in range(n_new - 1):
                step = ref.language_model.get_input_embeddings()(nxt)[:, None]
                out = ref.language_model(inputs_embeds=step, past_key_values=out.past_key_values, use_cache=True, return_dict=True)
                nxt = out.logits[:, -1].float().argmax(-1)
                got.append(nxt)
            e[3].record()
            torch.cuda.synchronize()
            if r >= 2:
                t_enc.append(e[0].elapsed_time(e[1]))
                t_pre.append(e[1].elapsed_time(e[2]))
                t_dec.append(e[2].elapsed_time(e[3]) / max(n_new - 1, 1))
            toks = torch.stack(got, 1)
    med = lambda x: sorted(x)[len(x) // 2]
    total = med(t_enc) + med(t_pre) + med(t_dec) * (n_new - 1)
    return dict(what="reference modules (oracle/_ref), torch eager bf16 autocast on this GPU", llm=llm, batch=batch,
                attention="flash_attention_2" if flash else "eager", encode_ms=med(t_enc), splice_prefill_ms=med(t_pre),
                decode_ms_per_token=med(t_dec), new_tokens=n_new, ms_per_video=total / batch, videos_per_s=1000.0 * batch / total,
                S=int(emb.shape[1]), tokens=toks[0].tolist(), gpu=torch.cuda.get_device_name(0))


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--new", type=int, default=16)
    ap.add_argument("--llm", default="phi3.5")
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    res = measure(a.reps, a.new, a.llm)
    line = json.dumps(res)
    print(line)
    if a.out:
        os.makedirs(os.path.dirname(a.out), exist_ok=True)
        open(a.out, "w").write(line + "\n")
