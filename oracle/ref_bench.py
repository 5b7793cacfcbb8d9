"""ORACLE SUPPORT -- MEASUREMENT INFRASTRUCTURE ONLY (bench.py `--impl reference` and `cpu_baseline`).

Times THE REFERENCE'S OWN MODULES (imported from oracle/_ref or /root/reference through oracle/ref_shims.py, random-init, named
architecture, fp32 -- the reference's CPU mode, BASELINE configs[0]) on the host cores:

  RefCpuSampler.sample()   one bounded sample of the configs[1] workload (96 frames, S = 3483, 16 greedy tokens) through the stock
                           code path of every stage, full width, reduced depth, scaled by the unit / layer counts it skips:
                             CLIPVisionModel, 1 image, all 24 layers (the reference runs layer 24 + post_layernorm too)      x 12 images
                             PretrainInternVideo2 (depth 3 -> blocks 0, 1 with x_vis_return_idx=-2), 1 segment of 8 frames  x 12 x 39/2
                             Phi3ForCausalLM with 1 and with 2 decoder layers: KV-cached prefill (logits for ALL positions, as the
                             reference computes them) and one cached decode step -> per-layer = T(2) - T(1), fixed = T(1) - per-layer
                                                                                                                             x 32 layers, 15 steps
  cfg1_end_to_end()        BASELINE configs[0] MEASURED END TO END, full depth: 8 frames, 1 segment, the reference's own encode_images
                           -> prepare_multimodal_inputs -> KV-cached greedy decode of 16 tokens, fp32 (about a minute on 16 cores,
                           ~22 GB of host memory for the random-init 3.8B decoder).
Never imported by the product."""
import os
import time

import torch

from . import ref_modules as RM


def _t(fn, reps=1):
    best = None
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return best


class RefCpuSampler:
    S = 3420 + 63
    NEW = 16

    def __init__(self, threads=None):
        from gvl import synth  # parameter NAMES / shapes only (state_dict keys of the reference); no kernels involved
        self.threads = threads or os.cpu_count() or 1
        torch.set_num_threads(self.threads)
        rope = synth.phi35_rope(96)
        with torch.no_grad(), RM.no_init():
            self.clip = RM.build_clip(device="cpu").float()
            self.iv2 = RM.build_iv2(frames=8, depth=3, flash=False, device="cpu", dtype=torch.float32)
            self.lm = {n: RM.build_lm("phi3", RM.phi3_config(layers=n, rope=rope), device="cpu", dtype=torch.float32) for n in (1, 2)}
        for i, m in enumerate([self.clip, self.iv2] + list(self.lm.values())):
            RM.fast_fill_(m, seed=i, threads=self.threads)
        g = torch.Generator().manual_seed(1234)
        self.img = torch.randn(1, 3, 336, 336, generator=g)
        self.seg = torch.randn(1, 3, 8, 224, 224, generator=g)
        self.emb = torch.randn(1, self.S, 3072, generator=g) * 0.05

    @torch.no_grad()
    def sample(self):
        t = {}
        t["clip_1_image_24_layers"] = _t(lambda: self.clip(self.img, output_hidden_states=True))
        t["iv2_1_segment_2_blocks"] = _t(lambda: self.iv2(self.seg, None, False, x_vis_return_idx=-2, x_vis_only=True))
        pre, dec = {}, {}
        for n, lm in self.lm.items():
            cache = RM.ShimCache()
            t0 = time.perf_counter()
            out = lm(inputs_embeds=self.emb, past_key_values=cache, use_cache=True, return_dict=True)
            pre[n] = time.perf_counter() - t0
            nxt = out.logits[:, -1].argmax(-1)
            step = lm.get_input_embeddings()(nxt)[:, None]
            t0 = time.perf_counter()
            lm(inputs_embeds=step, past_key_values=out.past_key_values, use_cache=True, return_dict=True)
            dec[n] = time.perf_counter() - t0
        layer_p, layer_d = max(pre[2] - pre[1], 0.0), max(dec[2] - dec[1], 0.0)
        fixed_p, fixed_d = max(pre[1] - layer_p, 0.0), max(dec[1] - layer_d, 0.0)
        t.update(lm_prefill_per_layer=layer_p, lm_prefill_fixed_embed_norm_lm_head_all_rows=fixed_p, lm_decode_step_per_layer=layer_d,
                 lm_decode_step_fixed=fixed_d)
        sec = (12 * t["clip_1_image_24_layers"] + 12 * 39 / 2.0 * t["iv2_1_segment_2_blocks"] + fixed_p + 32 * layer_p
               + (self.NEW - 1) * (fixed_d + 32 * layer_d))
        return sec, t

    SAMPLE = ("the reference's own modules (oracle/_ref), fp32, all host cores: CLIPVisionModel 1 image x 24 layers (x12), "
              "PretrainInternVideo2 1 segment x 2 blocks (x12x39/2), Phi3ForCausalLM 1- and 2-layer models, KV-cached prefill S=3483 with "
              "all-row logits + 1 cached decode step (per-layer = difference, x32 layers; x15 steps): scaled to the full clip")


@torch.no_grad()
def cfg1_end_to_end(threads=None, new_tokens=16, seed=0):
    """BASELINE configs[0], measured: reference modules at FULL depth on the CPU, fp32, eager attention, greedy, 8 frames / 1 segment."""
    from gvl import synth
    threads = threads or os.cpu_count() or 1
    torch.set_num_threads(threads)
    t0 = time.perf_counter()
    # random-init modules of the named architecture, built directly in fp32 (the reference's CPU dtype); parameters drawn with the
    # parallel filler instead of the modules' single-threaded initialisers
    small = dict(lm=dict(synth.PHI35, layers=1, vocab=8), clip=dict(synth.CLIP_L336, layers=1), iv2=dict(synth.IV2_1B, depth=1))
    params, lm_cfg, _, _ = synth.make_params("phi3.5", device="cpu", seed=seed, lm_dtype=torch.float32, **small)
    with RM.no_init():
        vt = RM.build_clip(device="cpu").float()
        ve = RM.build_iv2(frames=8, depth=40, flash=False, device="cpu", dtype=torch.float32)
        lm = RM.build_lm("phi3", RM.phi3_config(rope=lm_cfg["rope"]), device="cpu", dtype=torch.float32)
        f = RM.build_vlm(params, "phi3.5", lm_cfg, frames_per_seg=8, device="cpu", flash=False, with_lm=False,
                         clip_kw=dict(layers=1), iv2_kw=dict(depth=1))          # projectors, glb_GN / sub_GN, bound methods
    for i, m in enumerate((vt, ve, lm)):
        RM.fast_fill_(m, seed=seed * 10 + i, threads=threads)
    f.vision_tower, f.video_encoder, f.language_model = vt, ve, lm
    f.multi_modal_projector, f.video_projecter = f.multi_modal_projector.float(), f.video_projecter.float()
    f.dtype = torch.float32
    f.config = type("C", (), {"hidden_size": 3072})()
    del params
    t_build = time.perf_counter() - t0
    s = synth.make_clip_inputs(1, num_frames=8, num_segs=1)
    ids = torch.tensor(s["input_ids"][0])[None]
    t = {}
    t0 = time.perf_counter()
    feats = f.encode_images({"spatial_pixel_values": s["spatial_pixel_values"], "temporal_pixel_values": s["temporal_pixel_values"]})
    t["encode_images_s"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    emb, _, _ = f.prepare_multimodal_inputs(ids, ids.clone(), torch.ones_like(ids), feats, ["video"])
    toks, _ = RM.greedy_generate(f.language_model, emb, new_tokens)
    t["splice_prefill_decode_s"] = time.perf_counter() - t0
    total = t["encode_images_s"] + t["splice_prefill_decode_s"]
    return dict(config="BASELINE configs[0]: Phi-3.5-3.8B, 8 frames (1 segment: 1 key-frame 336^2 + 8 frames 224^2), 285 visual tokens, "
                       "prefill S=%d, %d greedy tokens, fp32, host CPU, the reference's own modules end to end (measured, not extrapolated)"
                       % (emb.shape[1], new_tokens),
                seconds_per_video=total, videos_per_s=1.0 / total, cores=threads, build_s=t_build, detail_s=t, tokens=toks.tolist())
