"""GPU probe: full-size Phi-3.5 decode step. Times N graph-replayed steps against a ctx-token cache with CUDA events and,
with GVL_MEGA_TRACE=1, prints the per-phase clock64 breakdown of the single-kernel step (decode_mega.cu).

    python tools/probe_decode.py [ctx] [steps] [layers]
"""
import ctypes
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "grounded-video-llm_b200"))
sys.path.insert(0, ROOT)
from gvl import _lib, model, synth  # noqa: E402


def main():
    ctx = int(sys.argv[1]) if len(sys.argv) > 1 else 3483
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 32
    layers = int(sys.argv[3]) if len(sys.argv) > 3 else 32
    dev = "cuda:0"
    lib = _lib.load()
    cfg = dict(synth.PHI35, layers=layers)
    params, lm_cfg, _, _ = synth.make_params("phi3.5", device=dev, seed=0, lm=cfg, clip=dict(synth.CLIP_L336, layers=1),
                                             iv2=dict(synth.IV2_1B, depth=1))
    sd = params["language_model"]
    lm = model.CausalLM(sd, lm_cfg["arch"], lm_cfg["heads"], lm_cfg["kv_heads"], lm_cfg["head_dim"], lm_cfg["eps"],
                        lm_cfg["rope"], max_ctx=4096, device=dev)
    emb = (torch.randn(ctx, lm.dim, device=dev) * 0.05).bfloat16()
    for _ in range(2):
        lm.generate(inputs_embeds=emb[None], max_new_tokens=steps + 1)
    torch.cuda.synchronize()
    # time prefill and prefill+decode separately (decode = difference)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    ev[0].record()
    lm.prefill(emb, n_new=steps + 1)
    ev[1].record()
    lm.generate(inputs_embeds=emb[None], max_new_tokens=steps + 1)
    ev[2].record()
    torch.cuda.synchronize()
    pre, both = ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2])
    per_step = (both - pre) / steps
    wbytes = layers * 113.25e6 * 2 + 32366 * 3072 * 2
    kvb = (ctx + steps / 2) * 2 * 96 * 32 * 2 * layers
    peak = 6555.2
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        peak = json.load(open(p))["hbm_gbs"]
    print("mode mega=%s  ctx=%d layers=%d: prefill %.2f ms, decode %.3f ms/step  -> %.0f GB/s (%.1f%% of %.0f)" % (
        os.environ.get("GVL_DECODE_MEGA", "1"), ctx, layers, pre, per_step, (wbytes + kvb) / per_step / 1e6,
        100 * (wbytes + kvb) / per_step / 1e6 / peak, peak))
    # the decode launch alone, repeated: (generate - prefill) above is ONE sample taken right after a tensor-bound prefill and moves
    # by +-10 % with the power state; here every repetition re-prefills, lets the GPU settle, and times gvl_lm_decode by itself
    reps = int(os.environ.get("GVL_PROBE_REPS", "7"))
    toks = torch.zeros(steps + 1, dtype=torch.int64, device=dev)
    stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    samples = []
    for _ in range(reps):
        lm.prefill(emb, n_new=steps + 1)
        handle = lm._active[0]
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = lib.gvl_lm_decode(handle, steps, ctypes.c_void_p(toks.data_ptr()), ctypes.c_void_p(0), -1, 0, stream)
        e1.record()
        torch.cuda.synchronize()
        assert rc == 0, rc
        samples.append(e0.elapsed_time(e1) / steps)
    samples.sort()
    med, best = samples[len(samples) // 2], samples[0]
    nbytes = wbytes + kvb
    print("decode launch alone, %d repetitions of %d steps: median %.3f ms/step (%.1f%% of %.0f GB/s), best %.3f (%.1f%%), all %s" % (
        reps, steps, med, 100 * nbytes / med / 1e6 / peak, peak, best, 100 * nbytes / best / 1e6 / peak,
        " ".join("%.3f" % v for v in samples)))
    if os.environ.get("GVL_MEGA_TRACE"):
        handle = lm._active[0]
        buf = np.zeros((160, 1024), dtype=np.int64)
        n, st = ctypes.c_int(), ctypes.c_int()
        rc = lib.gvl_lm_mega_trace(handle, buf.ctypes.data_as(ctypes.c_void_p), 160, ctypes.byref(n), ctypes.byref(st))
        if rc != 0:
            print("trace unavailable rc=%d" % rc)
            return
        t = buf[: n.value].astype(np.float64)
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        np.save(os.path.join(ROOT, "gpurun_out", "decode_trace_raw.npy"), buf[: n.value])
        L = layers
        # marks: 0 start | embed: done, passed | per layer 5 phases x (staged, done, passed) | lm_head x 3
        names = ["qkv", "attn", "o_proj", "gate_up", "down"]
        agg = {k: np.zeros(3) for k in names + ["lm_head", "embed"]}
        worst = {k: 0.0 for k in agg}
        agg["embed"] += [0, (t[:, 1] - t[:, 0]).mean(), (t[:, 2] - t[:, 1]).mean()]
        base = 3
        for l in range(L):
            for i, k in enumerate(names):
                j = base + (l * 5 + i) * 3
                prev = t[:, j - 1]
                agg[k] += [(t[:, j] - prev).mean(), (t[:, j + 1] - t[:, j]).mean(), (t[:, j + 2] - t[:, j + 1]).mean()]
                worst[k] += (t[:, j + 1] - prev).max()
        j = base + L * 15
        agg["lm_head"] += [(t[:, j] - t[:, j - 1]).mean(), (t[:, j + 1] - t[:, j]).mean(), (t[:, j + 2] - t[:, j + 1]).mean()]
        total = (t[:, j + 2] - t[:, 0]).mean()
        print("trace of the last step (clock64 cycles, mean over %d CTAs; per-step totals):" % n.value)
        print("  %-8s %12s %12s %12s %14s" % ("phase", "stage_x", "work", "barrier_wait", "slowest_CTA_sum"))
        for k in ["embed"] + names + ["lm_head"]:
            a = agg[k]
            print("  %-8s %12.0f %12.0f %12.0f %14.0f" % (k, a[0], a[1], a[2], worst[k]))
        g0, g1 = buf[0, 958], buf[0, 959]
        if g1 > g0 > 0:
            print("  kernel wall time (globaltimer, all %d steps of the last launch): %.3f ms = %.3f ms/step" % (
                steps, (g1 - g0) / 1e6, (g1 - g0) / 1e6 / steps))
        occ = buf[: n.value, 768:800].reshape(n.value, 4, 8).astype(np.float64)
        print("  ring slots already landed when the phase starts (last layer, mean over CTAs x warps, of 3): "
              + ", ".join("%s %.2f" % (k, occ[:, i].mean()) for i, k in enumerate(["qkv", "o_proj", "gate_up", "down"])))
        wt = buf[: n.value, 800:832].reshape(n.value, 4, 8).astype(np.float64)
        ni = buf[: n.value, 832:864].reshape(n.value, 4, 8).astype(np.float64)
        print("  last layer, per consumer warp: cycles spent WAITING for ring items / items consumed / phase work cycles (mean over CTAs x warps): "
              + ", ".join("%s %.0f / %.1f" % (k, wt[:, i].mean(), ni[:, i].mean()) for i, k in enumerate(["qkv", "o_proj", "gate_up", "down"])))
        sm = buf[: n.value, 864:888].reshape(n.value, 3, 8)[:, :, :4].astype(np.float64)
        print("  staging of the last layer (cycles, mean over CTAs): entry -> loads in registers -> CTA barrier -> written: "
              + ", ".join("%s %s" % (k, np.diff(sm[:, i], axis=1).mean(0).astype(int).tolist()) for i, k in enumerate(["qkv", "attn_merge", "gate_up"])))
        gm = buf[: n.value, 896:928].reshape(n.value, 4, 8)[:, :, :4].astype(np.float64)
        print("  GEMV phases of the last layer, warp 0 (cycles, mean over CTAs): entry -> item loop done -> CTA barrier -> epilogue done: "
              + ", ".join("%s %s" % (k, np.diff(gm[:, i], axis=1).mean(0).astype(int).tolist()) for i, k in enumerate(["qkv", "o_proj", "gate_up", "down"])))
        a = buf[: n.value, 960:1024].reshape(n.value, 8, 8).astype(np.float64)
        d = np.diff(a, axis=2)                                   # [cta, warp, 7]
        tot = a[:, :, 7] - a[:, :, 0]
        labels = ["rope", "append", "kv_loop", "new_tok", "reduce+write", "cta_bar", "merge"]
        print("attention phase of the last layer, per warp (cycles): mean / max over %d warps" % (n.value * 8))
        for i, k in enumerate(labels):
            print("    %-14s %9.0f %9.0f" % (k, d[:, :, i].mean(), d[:, :, i].max()))
        w = np.unravel_index(np.argmax(a[:, :, 5] - a[:, :, 0]), tot.shape)
        print("    slowest warp before the CTA barrier: cta %d warp %d: %s" % (w[0], w[1], d[w[0], w[1]].astype(int).tolist()))
        print("  total cycles/step %.0f  (= %.3f ms at %.0f MHz if the SM clock was that)" % (total, total / 1.9e6, 1900))


if __name__ == "__main__":
    main()
