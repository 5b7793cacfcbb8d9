"""GPU probe: 2-CTA (cta_group::2) GEMM vs the 1-CTA kernel -- correctness of every epilogue and TFLOP/s per shape.
Each (mode, case) runs in its own subprocess; GVL_GEMM_2CTA selects the kernel."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "grounded-video-llm_b200"))
sys.path.insert(0, os.path.join(ROOT, "tools"))

CORRECT = ["gemm_iv2_fc1_big", "gemm_iv2_proj_big", "gemm_clip_fc2_big", "gemm_quickgelu_big", "gemm_swiglu_big", "gemm_k640_big",
           "gemm_ragged_n"]
SHAPES = [(24588, 6144, 1408, 1), (24588, 1408, 6144, 0), (24588, 4608, 1408, 0), (24588, 1408, 1408, 0), (6924, 4096, 1024, 2),
          (6924, 1024, 4096, 0), (6924, 3072, 1024, 0), (3484, 16384, 3072, 3), (3484, 3072, 8192, 0), (3484, 9216, 3072, 0),
          (8192, 8192, 8192, 0)]


def correctness(name):
    import probe_ops as P
    table = {
        "gemm_iv2_fc1_big": dict(M=24588, N=6144, K=1408, act=1, bias=True),
        "gemm_iv2_proj_big": dict(M=24588, N=1408, K=1408, bias=True, gamma=True, res="bf16"),
        "gemm_clip_fc2_big": dict(M=6924, N=1024, K=4096, bias=True, res="f32", out_f32=True),
        "gemm_quickgelu_big": dict(M=6924, N=4096, K=1024, act=2, bias=True),
        "gemm_swiglu_big": dict(M=3484, N=16384, K=3072, act=3),
        "gemm_k640_big": dict(M=24576, N=1408, K=640, bias=True),
        "gemm_ragged_n": dict(M=20000, N=4224, K=1408),
    }
    err, scale = P._gemm_case(**table[name])
    return "err=%.4g scale=%.4g" % (err, scale)


def perf():
    import torch
    from gvl import ops
    out = []
    for (M, N, K, act) in SHAPES:
        a = (torch.randn(M, K, device="cuda") * 0.5).bfloat16()
        w = (torch.randn(N, K, device="cuda") * 0.02).bfloat16()
        b = torch.randn(N, device="cuda").bfloat16() if act in (1, 2) else None
        o = torch.empty(M, N // 2 if act == 3 else N, device="cuda", dtype=torch.bfloat16)
        for _ in range(3):
            ops.gemm(a, w, bias=b, act=act, out=o)
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(20):
            ops.gemm(a, w, bias=b, act=act, out=o)
        e.record()
        torch.cuda.synchronize()
        ms = s.elapsed_time(e) / 20
        out.append("%dx%dx%d act%d: %.3fms %.0fTF" % (M, N, K, act, ms, 2.0 * M * N * K / ms / 1e9))
    return " | ".join(out)


def main():
    if os.environ.get("PROBE_CHILD"):
        name = sys.argv[1]
        print("RESULT[2cta=%s] %s :: %s" % (os.environ.get("GVL_GEMM_2CTA", "-"), name, perf() if name == "perf" else correctness(name)))
        return
    for mode in ("1", "0"):
        for name in (CORRECT if mode == "1" else []) + ["perf"]:
            try:
                r = subprocess.run([sys.executable, __file__, name], capture_output=True, text=True, timeout=300,
                                   env=dict(os.environ, PROBE_CHILD="1", GVL_GEMM_2CTA=mode))
                txt = r.stdout + r.stderr
                if "RESULT" in txt:
                    print(txt[txt.index("RESULT"):].strip().splitlines()[0])
                else:
                    print("FAIL[2cta=%s] %s rc=%d :: %s" % (mode, name, r.returncode, " | ".join(txt.strip().splitlines()[-6:])))
            except subprocess.TimeoutExpired:
                print("TIMEOUT[2cta=%s] %s" % (mode, name))
            sys.stdout.flush()


if __name__ == "__main__":
    main()
