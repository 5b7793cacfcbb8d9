#!/usr/bin/env python
"""bench.py -- videos/sec (prefill + decode) of the Grounded-VideoLLM forward path, Phi-3.5-3.8B, 96-frame clips.

    python bench.py --gpus 1 --steps 5 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...      # the reference's own modules (oracle/_ref) on the host CPU cores

One "step" = one pass of the whole hot path over one batch of synthetic clips (BASELINE.json configs[1] at N=1:
1 clip = 12 CLIP key-frames 336^2 + 96 frames 224^2 -> 3420 visual tokens -> 3483-token prefill -> 16 greedy tokens).
At N>1 every rank owns `--clips-per-gpu` clips (weak scaling): (clip, segment) units are block-partitioned over the
ranks, projected visual tokens are exchanged with one NCCL all-gather, the decoder runs clip-sharded.
Prints ONE JSON line on rank 0 (see DESIGN.md "Measurement").
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "grounded-video-llm_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

METRIC = "videos/sec (prefill+decode) Phi3.5-3.8B 96-frame"
UNIT = "videos/s"
DECODE_TOKENS = 16
T_TEXT = 64                      # 64 ids incl. the <image> sentinel -> S = 3420 + 63
S_PREFILL = 3420 + T_TEXT - 1
# algorithmic work per video (identical counting rules)
FLOPS_CLIP, FLOPS_IV2, FLOPS_PROJ = 4.39e12, 59.50e12, 0.13e12
FLOPS_LM = 32 * (S_PREFILL * 226.5e6 + 6144.0 * S_PREFILL ** 2) + 2 * 3072 * 32366
FLOPS_PREFILL = FLOPS_CLIP + FLOPS_IV2 + FLOPS_PROJ + FLOPS_LM
DECODE_BYTES_WEIGHTS = 7.447e9
KV_BYTES_PER_CTX_TOKEN = 393216.0


_REAL_STDOUT = None


def _capture_stdout():
    """Rank 0 must print exactly ONE JSON line on stdout. Libraries write there too (NCCL prints its version banner with
    printf at communicator creation): park the real stdout on a spare descriptor and point fd 1 at stderr for the whole run."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def _emit(line):
    sys.stdout.flush()
    os.write(_REAL_STDOUT if _REAL_STDOUT is not None else 1, (line + "\n").encode())


def _peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sustained=d["bf16_tflops_sustained"], src="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sustained=1400.0, src="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index):
        self.rows, self.proc, self.idx = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            pass
        sm = sorted(float(r[1]) for r in self.rows if len(r) >= 8 and r[1].replace(".", "").isdigit())
        mx = [float(r[2]) for r in self.rows if len(r) >= 8 and r[2].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 8 for i in range(4) if r[4 + i].lower().startswith("active")})
        pw = [float(r[3]) for r in self.rows if len(r) >= 8 and r[3].replace(".", "").isdigit()]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "power_w_max": max(pw) if pw else None, "samples": len(sm)}


# ------------------------------------------------------------------------------------------------ CPU (reference) arm
REF_ARM_BUDGET_S = 100.0      # timed samples of the reference arm stop after this much CPU time (the run must end within minutes)


def _ref_sampler():
    """The REFERENCE'S OWN modules on the host cores (oracle/ref_bench.py; kind "reference"): test / measurement infrastructure,
    the one place besides tests/ and smoke() where bench.py executes anything under oracle/."""
    from oracle import ref_bench, ref_shims
    if not ref_shims.available():
        raise RuntimeError("reference files not installed: run `python -c 'import __graft_entry__ as g; g.build()'` where "
                           "/root/reference exists (oracle/build_ref.py copies them to the git-ignored oracle/_ref/)")
    return ref_bench


def cpu_reference_sample(sampler=None):
    """One bounded sample of the configs[1] workload through the reference's stock modules (fp32, all host cores), scaled to the
    full clip. Returns (videos_per_s, seconds_per_video, detail, sampler)."""
    rb = _ref_sampler()
    sampler = sampler or rb.RefCpuSampler()
    sec, detail = sampler.sample()
    return 1.0 / sec, sec, detail, sampler


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    rb = _ref_sampler()
    cores = os.cpu_count() or 1
    sampler = rb.RefCpuSampler(cores)
    vals, spent, walls = [], 0.0, []
    for i in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        sec, detail = sampler.sample()
        spent += time.perf_counter() - t0
        if i >= args.warmup:
            vals.append((sec, detail))
            walls.append(time.perf_counter() - t0)
        if spent > REF_ARM_BUDGET_S and (vals or i + 1 >= args.warmup):
            if not vals:
                vals.append((sec, detail))                         # the warm-up alone used the budget: keep its last sample
            break
    sec = sum(x[0] for x in vals) / len(vals)
    v = 1.0 / sec
    cfg1 = None
    if os.environ.get("GVL_REF_CFG1", "1") != "0":
        try:
            cfg1 = rb.cfg1_end_to_end(cores)
        except Exception as e:                                     # noqa: BLE001 -- e.g. not enough host memory for the fp32 3.8B decoder
            cfg1 = {"unavailable": "%s: %s" % (type(e).__name__, e)}
    out = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
           "dtype": "f32", "data": "synthetic",
           "config": {"workload": "BASELINE configs[1]: Phi-3.5-3.8B grounding inference, 1 clip = 96 frames (12x336^2 + 96x224^2), prefill "
                                  "S=%d + %d greedy tokens" % (S_PREFILL, DECODE_TOKENS),
                      "where": "host CPU, the reference's own modules (oracle/_ref), torch fp32, %d threads" % cores,
                      "samples_timed": len(vals), "sample_budget_s": REF_ARM_BUDGET_S,
                      "wall_s_per_sample": (sum(walls) / len(walls)) if walls else None,
                      "scaling_of_a_step": "ms_per_step = one bounded sample SCALED to a full clip (units and layers it skips multiplied "
                                           "back in, cpu_baseline.sample); the wall time one sample takes is wall_s_per_sample, so "
                                           "steps x ms_per_step is not this run's duration"},
           "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "reference", "sample": sampler.SAMPLE},
           "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0, "detail_s": vals[-1][1],
           "extra": {"cfg1_cpu_end_to_end": cfg1}}
    _emit(json.dumps(out))
    return 0


# ------------------------------------------------------------------------------------------------ B200 arm
def run_gvl_arm(args):
    import torch
    import torch.distributed as dist
    from gvl import _lib, model, ops, synth
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: the gvl hot path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = "cuda:%d" % local
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"      # the version banner goes to stdout; rank 0 must print exactly ONE JSON line
        import datetime
        # a collective that one rank never enters must fail in minutes, not hold the box for the default 10-minute watchdog
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device(dev), timeout=datetime.timedelta(seconds=180))
    lib = _lib.load()
    B = world * args.clips_per_gpu
    params, lm_cfg, clip_cfg, iv2_cfg = synth.make_params("phi3.5", device=dev, seed=0)
    m = model.LLAVA_NEXT_VIDEO(params, llm="phi3.5", lm_cfg=lm_cfg, clip_cfg=clip_cfg, iv2_cfg=iv2_cfg, max_ctx=4096, device=dev)
    del params
    torch.cuda.empty_cache()
    host = synth.make_clip_inputs(B, pin=True)                       # pinned host buffers (e2e path)
    resident = dict(host)
    resident["spatial_pixel_values"] = host["spatial_pixel_values"].to(dev)
    resident["temporal_pixel_values"] = host["temporal_pixel_values"].to(dev)
    h2d = host["spatial_pixel_values"].numel() * 4 + host["temporal_pixel_values"].numel() * 4 + B * T_TEXT * 8
    d2h = B * DECODE_TOKENS * 8

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def one_step(samples, to_host):
        toks = m.generate(samples, max_new_tokens=DECODE_TOKENS)
        if to_host:
            return [t.cpu() for t in toks]                             # device -> host read of the step's result
        return toks

    def timed(samples, to_host, steps):
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        barrier()
        for a, b in evs:
            a.record()
            one_step(samples, to_host)
            b.record()
        barrier()
        ms = sum(a.elapsed_time(b) for a, b in evs)
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)                   # max over ranks
        return float(t.item())

    for _ in range(max(args.warmup, 3)):
        one_step(resident, False)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = ops.launch_count()
    total_ms = timed(resident, False, args.steps)
    launches = ops.launch_count() - l0
    for _ in range(2):
        one_step(host, True)
    e2e_ms = timed(host, True, args.steps)
    # ---- e2e from RAW frames (SURVEY 8d synthetic clip: uint8 [96,3,336,336], seed 1234): H2D of the uint8 clip, GPU
    # frame_transform (bit-exact Pillow bicubic 336 -> 224 + normalisation, csrc/preprocess.cu), then the same generate call
    from gvl import preprocess
    raw_host = torch.randint(0, 256, (B, 96, 3, 336, 336), dtype=torch.uint8, generator=torch.Generator().manual_seed(1234)).pin_memory()

    def raw_step():
        raw = raw_host.to(dev, non_blocking=True)
        px = [preprocess.create_pixel_inputs(raw[b], 96, 12) for b in range(B)]
        smp = dict(host)
        smp["spatial_pixel_values"] = torch.cat([p_["spatial_pixel_values"] for p_ in px])
        smp["temporal_pixel_values"] = torch.cat([p_["temporal_pixel_values"] for p_ in px])
        return [t.cpu() for t in m.generate(smp, max_new_tokens=DECODE_TOKENS)]

    for _ in range(2):
        raw_step()
    rv = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    for a_, b_ in rv:
        a_.record()
        raw_step()
        b_.record()
    barrier()
    raw_t = torch.tensor([sum(a_.elapsed_time(b_) for a_, b_ in rv)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(raw_t, op=dist.ReduceOp.MAX)
    raw_ms = float(raw_t.item())
    clocks = sampler.stop() if rank == 0 else None

    # ---- roofline pass: per-kernel-family CUDA events on the launching stream, one extra step, graph replay off
    import ctypes
    roof, extra = None, {}
    for lmh in m.language_model._lms.values():
        lib.gvl_lm_set_graph(lmh, 0)
    lib.gvl_profile_enable(1)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    barrier()
    ev[0].record()
    feats = m.encode_images(resident)
    ev[1].record()
    from gvl import hostlogic
    ids, mask = hostlogic.left_pad(resident["input_ids"], 0, 2048)
    mine = [b for b in range(B) if b % world == rank]
    emb, _, masks = m.prepare_multimodal_inputs(ids[mine], None, mask[mine], feats[mine], ["v"] * len(mine))
    for i in range(len(mine)):
        m.language_model.prefill(emb[i], n_new=DECODE_TOKENS)
    ev[2].record()
    torch.cuda.synchronize()
    fam = {}
    for kind, name in ((0, "gemm"), (1, "attn"), (2, "gemv")):     # families of exactly ONE encode + prefill per owned clip
        ms, work, n = ctypes.c_double(), ctypes.c_double(), ctypes.c_longlong()
        lib.gvl_profile_collect(kind, ctypes.byref(ms), ctypes.byref(work), ctypes.byref(n))
        fam[name] = (ms.value, work.value, n.value)
    lib.gvl_profile_enable(0)
    m.language_model.generate(inputs_embeds=emb[:1], attention_mask=masks[:1], max_new_tokens=DECODE_TOKENS)
    ev[3].record()
    torch.cuda.synchronize()
    for lmh in m.language_model._lms.values():
        lib.gvl_lm_set_graph(lmh, 1)
    # ---- decode roofline: clean (profiling off) prefill vs prefill + 16-token generate on the same embeddings; the
    # difference is the 15 decode steps of the single-kernel decode path (one launch runs all steps of a generate call)
    dv = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
    m.language_model.generate(inputs_embeds=emb[:1], attention_mask=masks[:1], max_new_tokens=DECODE_TOKENS)
    torch.cuda.synchronize()
    dv[0].record()
    m.language_model.prefill(emb[0], n_new=DECODE_TOKENS)
    dv[1].record()
    m.language_model.generate(inputs_embeds=emb[:1], attention_mask=masks[:1], max_new_tokens=DECODE_TOKENS)
    dv[2].record()
    # clean prefill-side timing (the per-launch events of the family pass above cost ~10 % of the prefill)
    feats2 = m.encode_images(resident)
    dv[3].record()
    emb2, _, _ = m.prepare_multimodal_inputs(ids[mine], None, mask[mine], feats2[mine], ["v"] * len(mine))
    for i in range(len(mine)):
        m.language_model.prefill(emb2[i], n_new=DECODE_TOKENS)
    dv[4].record()
    torch.cuda.synchronize()
    clean_enc_ms, clean_pre_ms = dv[2].elapsed_time(dv[3]), dv[3].elapsed_time(dv[4])
    dec_steps = DECODE_TOKENS - 1                                     # the first token comes out of the prefill
    dec_step_diff_ms = (dv[1].elapsed_time(dv[2]) - dv[0].elapsed_time(dv[1])) / dec_steps
    # The difference above is ONE sample taken right after a tensor-bound prefill; it moves by +-10 % with the power state (2.3 - 2.9 ms
    # for the same binary inside one job, profiles/r2_decode.md). The roofline figure is therefore the decode launch timed BY ITSELF:
    # prefill (untimed) -> CUDA events around gvl_lm_decode(15 steps), 5 repetitions, median.
    dec_samples = []
    dtoks = torch.zeros(DECODE_TOKENS, dtype=torch.int64, device=dev)
    cur_stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    for _ in range(5):
        m.language_model.prefill(emb[0], n_new=DECODE_TOKENS)
        handle = m.language_model._active[0]
        torch.cuda.synchronize()
        d0, d1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        d0.record()
        rc = lib.gvl_lm_decode(handle, dec_steps, ctypes.c_void_p(dtoks.data_ptr()), ctypes.c_void_p(0), -1, 0, cur_stream)
        d1.record()
        torch.cuda.synchronize()
        assert rc == 0, rc
        dec_samples.append(d0.elapsed_time(d1) / dec_steps)
    dec_samples.sort()
    dec_step_ms = dec_samples[len(dec_samples) // 2]
    dec_bytes = DECODE_BYTES_WEIGHTS + (emb.shape[1] + dec_steps / 2.0) * KV_BYTES_PER_CTX_TOKEN
    pk = _peaks()
    if rank == 0:
        g_ms, g_fl, g_n = fam["gemm"]
        a_ms, a_fl, a_n = fam["attn"]
        v_ms, v_by, v_n = fam["gemv"]
        ach = g_fl / (g_ms * 1e-3) / 1e12 if g_ms > 0 else 0.0
        prof_ms = ev[0].elapsed_time(ev[2])                         # one encode_images + one splice/prefill per owned clip
        # dram__bytes_read + dram__bytes_write of the dominant launch of this family (InternVideo2 fc1, 24588x6144x1408, algorithmic
        # 388 MB) from the committed `ncu --set full` capture of the SAME kernel (profiles/r2_gemm.md); not measured live
        roof = {"kernel": "gvl::gemm_bf16_tcgen05_2cta_kernel (cta_group::2, all epilogue variants; the 1-CTA kernel for small shapes)",
                "bound": "tensor", "achieved": ach, "peak": pk["tf_sustained"], "unit": "TFLOP/s", "frac": ach / pk["tf_sustained"],
                "traffic": 337.0e6, "traffic_launch": "IV2 fc1 24588x6144x1408: 88.0 MB read + 249.1 MB written (ncu --set full, "
                                                       "profiles/r2_gemm_iv2_fc1_ncu_metrics.csv); algorithmic 69 + 17 + 302 MB",
                "launches": g_n, "avg_launch_ms": g_ms / max(g_n, 1), "peak_source": pk["src"] + " (sustained cuBLAS bf16)",
                "share_of_step": g_ms / prof_ms if prof_ms > 0 else None, "family_ms": g_ms, "profiled_region_ms": prof_ms,
                "profiled_region": "one encode_images + one splice/prefill (per-launch CUDA events on; decode excluded)"}
        enc_ms, pre_ms, dec_ms = ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2]), ev[2].elapsed_time(ev[3])
        per_rank_clips = len(mine)
        prefill_s = (clean_enc_ms + clean_pre_ms) * 1e-3
        units_here = 12 * B / world
        prefill_flops = units_here * (FLOPS_CLIP + FLOPS_IV2 + FLOPS_PROJ) / 12 + per_rank_clips * FLOPS_LM
        extra = {
            "stage_ms_profiled": {"encode_images": enc_ms, "splice+prefill": pre_ms, "prefill+%d_decode" % DECODE_TOKENS: dec_ms},
            "stage_ms": {"encode_images": clean_enc_ms, "splice+prefill": clean_pre_ms, "decode_step": dec_step_ms,
                         "note": "CUDA events without the per-launch profiling events; prefill_tflops_achieved uses these"},
            "prefill_tflops_achieved": prefill_flops / prefill_s / 1e12,
            "prefill_frac_of_tensor_peak": prefill_flops / prefill_s / 1e12 / pk["tf_sustained"],
            "attention": {"ms": a_ms, "tflops": a_fl / (a_ms * 1e-3) / 1e12 if a_ms > 0 else None, "launches": a_n,
                          "kernel": "gvl::attn_tc2_kernel (tcgen05; attn_tc_kernel for single-tile shapes)"},
            "decode": {"bound": "hbm", "ms_per_step": dec_step_ms, "bytes_per_step": dec_bytes,
                       "achieved_gbs": dec_bytes / (dec_step_ms * 1e-3) / 1e9, "peak_gbs": pk["hbm"],
                       "frac": dec_bytes / (dec_step_ms * 1e-3) / 1e9 / pk["hbm"],
                       "traffic_bytes_per_step_ncu": 8.86e9,   # dram__bytes_read of one captured launch / its 8 steps (profiles/r2_decode_mega_final_ncu_metrics.csv)
                       "kernel": "gvl::decode_mega_kernel<96> (one persistent launch per generate call; GVL_DECODE_MEGA=0: per-op chain)"
                                 if os.environ.get("GVL_DECODE_MEGA", "1") != "0" else "per-op chain: gemv3_kernel + decode_attn_kernel (CUDA graph)",
                       "samples_ms_per_step": dec_samples, "ms_per_step_generate_minus_prefill": dec_step_diff_ms,
                       "how": "CUDA events around the decode launch alone (gvl_lm_decode, %d steps, after an untimed prefill), median of 5; "
                              "(prefill + %d-token generate) - prefill is reported beside it; bytes = 7.447 GB weights + ctx x 393 KB K/V"
                              % (dec_steps, DECODE_TOKENS)},
            "gemv_launches_profiled": {"ms": v_ms, "launches": v_n},
            "e2e_from_raw_frames": {"value": B * args.steps / (raw_ms * 1e-3), "unit": UNIT, "ms_per_step": raw_ms / args.steps,
                                    "h2d_bytes_per_step": B * 96 * 3 * 336 * 336 + B * T_TEXT * 8,
                                    "what": "uint8 clip [96,3,336,336] in pinned host memory -> H2D -> GPU frame_transform (Pillow-bicubic "
                                            "bit-exact resize 336->224, key-frame selection, normalisation) -> generate -> tokens to host"},
            "peaks": pk,
        }
    # ---- result check across N: sha256 of clip 0's greedy tokens (clip 0 is the same synthetic clip at every N and every
    # clips-per-gpu: make_clip_inputs draws clips in order from one seeded generator), printed so that N=1 / 2 / 4 / 8 can be compared
    import hashlib
    toks0 = one_step(resident, True)[0]
    tok_sha = hashlib.sha256(",".join(str(int(x)) for x in toks0.tolist()).encode()).hexdigest()[:16]
    # ---- strong scaling of ONE clip over the N ranks (12 units block-partitioned, decoder on rank 0): latency in ms
    strong_ms = None
    if world > 1:
        one = synth.make_clip_inputs(1, device=dev)
        for _ in range(2):
            m.generate(one, max_new_tokens=DECODE_TOKENS)
        strong_ms = timed(one, False, 3) / 3
        t1 = m.generate(one, max_new_tokens=DECODE_TOKENS)[0]          # every rank: generate issues collectives
        strong_sha = hashlib.sha256(",".join(str(int(x)) for x in t1.tolist()).encode()).hexdigest()[:16]
    if rank == 0:
        extra["tokens_clip0_sha256_16"] = tok_sha
        extra["strong_scaling_1_clip_ms"] = strong_ms
        if world > 1:
            extra["strong_scaling_tokens_sha256_16"] = strong_sha      # clip 0 with its 12 units encoded on `world` ranks
    cpu, gpu_ref = None, None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            sys.path.insert(0, os.path.join(ROOT, "tools"))
            import gpu_reference
            gpu_ref = gpu_reference.measure(reps=3, n_new=DECODE_TOKENS)
        except Exception as e:                                     # noqa: BLE001 -- a baseline leg must not take the bench line down
            gpu_ref = {"unavailable": "%s: %s" % (type(e).__name__, e)}
        v, sec, detail, sampler = cpu_reference_sample()
        cpu = {"value": v, "unit": UNIT, "cores": os.cpu_count(), "kind": "reference",
               "sample": sampler.SAMPLE + "; %.1f s/video" % sec, "detail_s": detail}
    if rank == 0:
        value = B * args.steps / (total_ms * 1e-3)
        out = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
               "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
               "dtype": "bf16", "data": "synthetic",
               "config": {"workload": "BASELINE configs[1] x %d clip(s)/GPU: Phi-3.5-3.8B, 96 frames (12 key-frames 336^2 + 96 frames 224^2), "
                                      "3420 visual tokens, prefill S=%d, %d greedy decode tokens, random-init weights" %
                                      (args.clips_per_gpu, S_PREFILL, DECODE_TOKENS),
                          "global_batch": B, "parallelism": "units block-partitioned over %d rank(s); one all-gather of visual tokens; "
                                                             "decoder clip-sharded" % world,
                          "l2": "no explicit flush: each step streams 10.9 GB of weights + >1 GB of activations (>> 126 MB L2)",
                          "timing": "CUDA events per step on the launching stream, barrier+synchronize on both sides, max over ranks"},
               "clocks": clocks,
               "e2e": {"value": B * args.steps / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                       "ms_per_step": e2e_ms / args.steps, "api": "gvl.model.LLAVA_NEXT_VIDEO.generate(samples) with pinned host tensors"},
               "gpu_launches": launches, "roofline": roof, "cpu_baseline": cpu, "extra": extra}
        out["extra"]["gpu_reference"] = gpu_ref          # the reference's modules, torch eager + FA2, same GPU (tools/gpu_reference.py)
        _emit(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="gvl", choices=["gvl", "reference"])
    ap.add_argument("--clips-per-gpu", type=int, default=1)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    _capture_stdout()
    if args.impl == "reference":
        return run_reference_arm(args)
    return run_gvl_arm(args)


if __name__ == "__main__":
    sys.exit(main())
