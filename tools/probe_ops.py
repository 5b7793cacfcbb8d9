"""GPU probe: run each operator check in its own subprocess (a trap/hang in one kernel must not take the
others down) and print a one-line verdict per case. Not a pytest file; used during bring-up.

    python tools/probe_ops.py            # all cases
    python tools/probe_ops.py gemm_basic # one case, in-process
"""
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "grounded-video-llm_b200"))


def _gemm_case(M, N, K, act=0, bias=False, gamma=False, res=None, out_f32=False, bn=0, seed=0):
    import torch
    from gvl import ops
    g = torch.Generator(device="cuda").manual_seed(seed)
    a = (torch.randn(M, K, device="cuda", generator=g) * 0.5).bfloat16()
    w = (torch.randn(N, K, device="cuda", generator=g) * 0.05).bfloat16()
    b = (torch.randn(N, device="cuda", generator=g) * 0.1).bfloat16() if bias else None
    n_out = N // 2 if act == 3 else N
    gm = (torch.rand(n_out, device="cuda", generator=g) + 0.5) if gamma else None
    r = None
    if res == "bf16":
        r = torch.randn(M, n_out, device="cuda", generator=g).bfloat16()
    elif res == "f32":
        r = torch.randn(M, n_out, device="cuda", generator=g)
    out = ops.gemm(a, w, bias=b, act=act, gamma=gm, residual=r,
                   out_dtype=torch.float32 if out_f32 else torch.bfloat16, bn=bn)
    torch.cuda.synchronize()
    # fp32 reference with the same rounding points
    acc = a.float() @ w.float().t()
    if b is not None:
        acc = acc + b.float()
    if act == 3:
        wi = acc.view(M, N // 256, 2, 128)
        gate = wi[:, :, 0, :].reshape(M, -1).bfloat16().float()
        up = wi[:, :, 1, :].reshape(M, -1).bfloat16().float()
        y = (up * torch.nn.functional.silu(gate).bfloat16().float()).bfloat16().float()
    else:
        y = acc.bfloat16().float()
        if act == 1:
            y = torch.nn.functional.gelu(y).bfloat16().float()
        elif act == 2:
            t = (1.702 * y).bfloat16().float()
            y = (y * torch.sigmoid(t).bfloat16().float()).bfloat16().float()
    if gm is not None:
        y = (y * gm).bfloat16().float()
    if r is not None:
        y = y + r.float()
    if not out_f32:
        y = y.bfloat16().float()
    err = (out.float() - y).abs().max().item()
    scale = y.abs().max().item()
    return err, scale


def case_gemm_basic():
    return _gemm_case(256, 256, 128)


def case_gemm_bn128():
    return _gemm_case(384, 384, 256, bn=128)


def case_gemm_ragged():
    return _gemm_case(577 * 2, 1024, 1024, bias=True)


def case_gemm_iv2_fc1():
    return _gemm_case(2049 * 2, 6144, 1408, act=1, bias=True)


def case_gemm_iv2_proj():
    return _gemm_case(2049, 1408, 1408, bias=True, gamma=True, res="bf16")


def case_gemm_iv2_qkv():
    return _gemm_case(2049, 4224, 1408)


def case_gemm_clip_fc2():
    return _gemm_case(577 * 3, 1024, 4096, bias=True, res="f32", out_f32=True)


def case_gemm_quickgelu():
    return _gemm_case(577, 4096, 1024, act=2, bias=True)


def case_gemm_swiglu():
    return _gemm_case(700, 16384, 3072, act=3)


def case_gemm_k640():
    return _gemm_case(2048, 1408, 640, bias=True)


def case_gemm_big():
    return _gemm_case(24588, 6144, 1408, act=1, bias=True)


def _attn_case(B, H, KVH, Sq, Skv, D, causal, round_scores=False, seed=0):
    import torch
    from gvl import ops
    g = torch.Generator(device="cuda").manual_seed(seed)
    q = torch.randn(B, Sq, H, D, device="cuda", generator=g).bfloat16()
    k = torch.randn(B, Skv, KVH, D, device="cuda", generator=g).bfloat16()
    v = torch.randn(B, Skv, KVH, D, device="cuda", generator=g).bfloat16()
    scale = D ** -0.5
    o = ops.attention(q, k, v, scale, causal=causal, round_scores=round_scores)
    torch.cuda.synchronize()
    rep = H // KVH
    qf = q.float().permute(0, 2, 1, 3)
    kf = k.float().permute(0, 2, 1, 3).repeat_interleave(rep, dim=1)
    vf = v.float().permute(0, 2, 1, 3).repeat_interleave(rep, dim=1)
    s = qf @ kf.transpose(-1, -2)
    if round_scores:
        s = s.bfloat16().float()
    s = s * scale
    if causal:
        i = torch.arange(Sq, device="cuda")[:, None]
        j = torch.arange(Skv, device="cuda")[None, :]
        s = s.masked_fill(j > i + (Skv - Sq), float("-inf"))
    p = torch.softmax(s, dim=-1)
    ref = (p @ vf).permute(0, 2, 1, 3)
    err = (o.float() - ref).abs().max().item()
    return err, ref.abs().max().item()


def case_attn_clip():
    return _attn_case(2, 16, 16, 577, 577, 64, False, round_scores=True)


def case_attn_iv2():
    return _attn_case(1, 16, 16, 2049, 2049, 88, False)


def case_attn_phi_causal():
    return _attn_case(1, 8, 8, 1000, 1000, 96, True)


def case_attn_llama_gqa():
    return _attn_case(1, 8, 2, 300, 300, 128, True)


def case_attn_small():
    return _attn_case(1, 2, 2, 50, 50, 64, False)


def case_norms():
    import torch
    from gvl import ops
    torch.manual_seed(0)
    x = torch.randn(577 * 2, 1024, device="cuda")
    w = torch.randn(1024, device="cuda")
    b = torch.randn(1024, device="cuda")
    y = ops.layernorm(x, w, b, 1e-5)
    ref = torch.nn.functional.layer_norm(x, (1024,), w, b, 1e-5).bfloat16()
    e1 = (y.float() - ref.float()).abs().max().item()
    xb = torch.randn(2049, 1408, device="cuda").bfloat16()
    wb = torch.randn(1408, device="cuda").bfloat16()
    y2 = ops.rmsnorm(xb, wb, 1e-6)
    xf = xb.float()
    ref2 = wb * (xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + 1e-6)).bfloat16()
    e2 = (y2.float() - ref2.float()).abs().max().item()
    return max(e1, e2), 1.0


CASES = {k[5:]: v for k, v in globals().items() if k.startswith("case_")}


def main():
    if os.environ.get("PROBE_CHILD") and len(sys.argv) > 1 and sys.argv[1] in CASES:
        t0 = time.time()
        err, scale = CASES[sys.argv[1]]()
        print("RESULT %s err=%.4g scale=%.4g time=%.1fs" % (sys.argv[1], err, scale, time.time() - t0))
        return
    for name in CASES:
        try:
            r = subprocess.run([sys.executable, __file__, name], capture_output=True, text=True, timeout=180, env=dict(os.environ, PROBE_CHILD="1"))
            lines = [l for l in (r.stdout + r.stderr).splitlines() if l.strip()]
            res = [l for l in lines if l.startswith("RESULT")]
            if res:
                print(res[-1])
            else:
                print("FAIL %s rc=%d :: %s" % (name, r.returncode, " | ".join(lines[-6:])))
        except subprocess.TimeoutExpired:
            print("TIMEOUT %s" % name)
        sys.stdout.flush()


if __name__ == "__main__":
    main()
