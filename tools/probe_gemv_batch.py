"""GPU probe: what a BATCHED decode step on the per-op chain would cost -- the chain's GEMV for M = 1..4 activation rows on the
Phi-3.5 shapes (weights read once for all rows) and the q_len = 1 attention per sequence at ctx 3483.
    python tools/probe_gemv_batch.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "grounded-video-llm_b200"))
from gvl import ops  # noqa: E402

SHAPES = [("qkv", 9216, 3072, 0, True), ("o_proj", 3072, 3072, 0, False), ("gate_up", 16384, 3072, 3, True), ("down", 3072, 8192, 0, False)]


def t_ms(fn, reps=40):
    for _ in range(5):
        fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def main():
    dev = "cuda"
    g = torch.Generator(device=dev).manual_seed(0)
    # several copies of every weight so that consecutive launches do not hit L2
    W = {n: [(torch.randn(N, K, device=dev, generator=g) * 0.02).bfloat16() for _ in range(6)] for n, N, K, _, _ in SHAPES}
    lmh = [(torch.randn(32366, 3072, device=dev, generator=g) * 0.02).bfloat16() for _ in range(2)]
    for M in (1, 2, 3, 4):
        per_layer = 0.0
        out = []
        for n, N, K, act, norm in SHAPES:
            x = (torch.randn(M, K, device=dev, generator=g) * 0.5).bfloat16()
            nw = torch.ones(K, device=dev, dtype=torch.bfloat16) if norm else None
            it = [0]

            def run():
                w = W[n][it[0] % 6]
                it[0] += 1
                ops.gemv(x, w, norm_w=nw, act=act)
            ms = t_ms(run)
            per_layer += ms
            out.append("%s %.1f us (%.0f GB/s)" % (n, ms * 1e3, 2.0 * N * K / ms / 1e6))
        x = (torch.randn(M, 3072, device=dev, generator=g) * 0.5).bfloat16()
        it = [0]

        def run_h():
            ops.gemv(x, lmh[it[0] % 2], out_dtype=torch.float32)
            it[0] += 1
        head = t_ms(run_h)
        print("M=%d: %s | lm_head %.1f us | 32 layers + head = %.3f ms" % (M, ", ".join(out), head * 1e3, 32 * per_layer + head))
    ctx = 3483
    kc = (torch.randn(32, 4096, 96, device=dev, generator=g)).bfloat16()
    vc = (torch.randn(32, 4096, 96, device=dev, generator=g)).bfloat16()
    q = torch.randn(32, 96, device=dev, generator=g).bfloat16()
    cl = torch.tensor([ctx], dtype=torch.int32, device=dev)
    ms = t_ms(lambda: ops.decode_attention(q, kc, vc, cl, 96 ** -0.5))
    print("decode attention, one sequence, ctx %d: %.1f us per layer -> %.3f ms per step and sequence" % (ctx, ms * 1e3, 32 * ms))


if __name__ == "__main__":
    main()
