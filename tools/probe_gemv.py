"""GPU probe: decode GEMV / decode-attention bandwidth per shape, measured from a CUDA-graph replay (no host launch
overhead between kernels), weights rotated through > L2 worth of copies so every launch streams from HBM."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "grounded-video-llm_b200"))
from gvl import ops  # noqa: E402

SHAPES = [("qkv+norm", 9216, 3072, 0, True), ("o_proj+res", 3072, 3072, 0, False), ("gate_up+norm+swiglu", 16384, 3072, 3, True),
          ("down+res", 3072, 8192, 0, False), ("lm_head+norm+bias", 32366, 3072, 0, True)]


def timed_graph(fn, reps):
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        fn()                                   # warm-up (func attributes etc.)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=s):
        for i in range(reps):
            fn(i)
    g.replay()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    g.replay()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e3


def main():
    peak = 6486.1
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        peak = json.load(open(p))["hbm_gbs"]
    print("mode: GVL_GEMV_BULK=%s" % os.environ.get("GVL_GEMV_BULK", "unset"))
    for name, N, K, act, norm in SHAPES:
        copies = max(2, int(400e6 // (N * K * 2)) + 1)
        ws = [(torch.randn(N, K, device="cuda") * 0.02).bfloat16() for _ in range(copies)]
        x = (torch.randn(1, K, device="cuda") * 0.5).bfloat16()
        nw = torch.ones(K, device="cuda").bfloat16() if norm else None
        res = torch.zeros(1, N, device="cuda").bfloat16() if "res" in name else None

        def fn(i=0):
            ops.gemv(x, ws[i % copies], norm_w=nw, residual=res, act=act)
        us = timed_graph(fn, 40)
        gbs = N * K * 2 / us / 1e3
        print("gemv %-22s N=%6d K=%5d: %7.1f us  %6.0f GB/s  (%.0f%% of %.0f)" % (name, N, K, us, gbs, 100 * gbs / peak, peak))
        del ws
    H, D, ctx, maxc = 32, 96, 3484, 4096
    caches = [(torch.randn(H, maxc, D, device="cuda").bfloat16(), torch.randn(H, maxc, D, device="cuda").bfloat16()) for _ in range(12)]
    q = torch.randn(H * D, device="cuda").bfloat16()
    cl = torch.tensor([ctx], dtype=torch.int32, device="cuda")

    def fa(i=0):
        kc, vc = caches[i % len(caches)]
        ops.decode_attention(q, kc, vc, cl, D ** -0.5)
    us = timed_graph(fa, 36)
    byts = 2 * H * ctx * D * 2
    print("decode_attention H=32 d=96 ctx=%d: %.1f us  %.0f GB/s (%.0f%%)  [+ one memset per call for the workspace]" % (
        ctx, us, byts / us / 1e3, 100 * byts / us / 1e3 / peak))


if __name__ == "__main__":
    main()
