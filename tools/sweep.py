"""BASELINE config 5: frame-count sweep (8 frames per segment, segs = frames/8) x batch, PREFILL ONLY
(encode_images + splice + decoder prefill), achieved TFLOP/s against the algorithmic FLOPs of BASELINE.md section 3.
frames >= 128 give S > 4096 and exercise the LongRoPE long-factor branch (modeling_phi3.py:380-385).

    python tools/sweep.py [--frames 16,32,64,96,128,192,256] [--batch 1,4]
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "grounded-video-llm_b200"))
from gvl import hostlogic, model, synth  # noqa: E402


def flops(segs, T_text=64):
    S = 285 * segs + T_text - 1
    lm = 32 * (S * 226.5e6 + 6144.0 * S * S) + 2 * 3072 * 32366
    return segs * (0.366e12 + 4.958e12 + 0.0104e12) + lm, S


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", default="16,32,64,96,128,192,256")
    ap.add_argument("--batch", default="1,4")
    a = ap.parse_args()
    peak = 1407.6
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        peak = json.load(open(pk))["bf16_tflops_sustained"]
    params, lm_cfg, clip_cfg, iv2_cfg = synth.make_params("phi3.5", device="cuda", seed=0)
    m = model.LLAVA_NEXT_VIDEO(params, llm="phi3.5", lm_cfg=lm_cfg, clip_cfg=clip_cfg, iv2_cfg=iv2_cfg, max_ctx=9344)
    del params
    torch.cuda.empty_cache()
    print("| frames | segs | S | batch | prefill ms/clip | TFLOP/clip | TFLOP/s | frac of %.0f (sustained cuBLAS bf16) | rope |" % peak)
    print("|---|---|---|---|---|---|---|---|---|")
    for frames in [int(x) for x in a.frames.split(",")]:
        segs = frames // 8
        fl, S = flops(segs)
        for B in [int(x) for x in a.batch.split(",")]:
            if B * frames > 1024:
                continue
            g = torch.Generator(device="cuda").manual_seed(1234)
            sp = torch.randn(B, segs, 3, 336, 336, device="cuda", generator=g)
            tp = torch.randn(B, frames, 3, 224, 224, device="cuda", generator=g)
            ids = torch.randint(3, 32000, (64,), generator=torch.Generator().manual_seed(7))
            ids[20] = -200
            samples = {"spatial_pixel_values": sp, "temporal_pixel_values": tp}
            idt, mask = hostlogic.left_pad([ids.tolist()] * B, 0, 2048)

            def step():
                feats = m.encode_images(samples)
                emb, _, _ = m.prepare_multimodal_inputs(idt, None, mask, feats, ["v"] * B)
                for b in range(B):
                    m.language_model.prefill(emb[b])

            for _ in range(2):
                step()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(3):
                step()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 3 / B
            tf = fl / ms / 1e9
            print("| %d | %d | %d | %d | %.2f | %.1f | %.0f | %.2f | %s |" % (frames, segs, S, B, ms, fl / 1e12, tf, tf / peak,
                                                                             "long" if S > 4096 else "short"))
            sys.stdout.flush()


if __name__ == "__main__":
    main()
