"""GPU probe: run ONE GEMM shape of the cfg2 step a few times (for `ncu --set full -k regex:gemm_bf16_tcgen05` captures and
CUDA-event timing).   python tools/probe_gemm_shape.py <name> [reps]
names: iv2_fc1 iv2_fc2 iv2_qkv iv2_proj phi_gate_up phi_down phi_qkv phi_o clip_fc1 clip_fc2"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "grounded-video-llm_b200"))
from gvl import ops  # noqa: E402

SHAPES = {  # M, N, K, act, bias, gamma, residual
    "iv2_fc1": (24588, 6144, 1408, 1, True, False, None),
    "iv2_fc2": (24588, 1408, 6144, 0, True, True, "bf16"),
    "iv2_qkv": (24588, 4608, 1408, 0, False, False, None),
    "iv2_proj": (24588, 1408, 1408, 0, True, True, "bf16"),
    "phi_gate_up": (3483, 16384, 3072, 3, False, False, None),
    "phi_down": (3483, 3072, 8192, 0, False, False, "bf16"),
    "phi_qkv": (3483, 9216, 3072, 0, False, False, None),
    "phi_o": (3483, 3072, 3072, 0, False, False, "bf16"),
    "clip_fc1": (6924, 4096, 1024, 2, True, False, None),
    "clip_fc2": (6924, 1024, 4096, 0, True, False, "f32"),
}


def run(name, reps=10):
    M, N, K, act, bias, gamma, res = SHAPES[name]
    g = torch.Generator(device="cuda").manual_seed(1)
    a = (torch.randn(M, K, device="cuda", generator=g) * 0.5).bfloat16()
    w = (torch.randn(N, K, device="cuda", generator=g) * 0.02).bfloat16()
    b = torch.randn(N, device="cuda", generator=g).bfloat16() if bias else None
    gm = torch.rand(N, device="cuda", generator=g) if gamma else None
    No = N // 2 if act == 3 else N
    out_f32 = res == "f32"
    o = torch.zeros(M, No, device="cuda", dtype=torch.float32 if out_f32 else torch.bfloat16)
    kw = dict(bias=b, act=act, out=o)
    if gamma:
        kw["gamma"] = gm
    if res is not None:
        kw["residual"] = o
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for _ in range(3):
        ops.gemm(a, w, **kw)
    ts = []
    for _ in range(reps):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        ops.gemm(a, w, **kw)
        e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    ms = sorted(ts)[len(ts) // 2]
    print("%s %dx%dx%d act%d: %.4f ms  %.0f TFLOP/s (median of %d, L2 flushed)" % (name, M, N, K, act, ms, 2.0 * M * N * K / ms / 1e9, reps))


if __name__ == "__main__":
    names = sys.argv[1].split(",") if len(sys.argv) > 1 else list(SHAPES)
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
    for n in names:
        run(n, reps)
