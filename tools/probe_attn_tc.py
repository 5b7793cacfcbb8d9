"""GPU probe for the tcgen05 attention kernel (attention_tc.cu): correctness vs an fp32 torch reference and timing,
each case in its own subprocess (a bad descriptor traps instead of hanging, see ptx::mbar_wait)."""
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "grounded-video-llm_b200"))

CASES = {
    # name: (B, H, KVH, Sq, Skv, D, causal, round_scores, o_dim)
    "d64_one_tile": (1, 1, 1, 128, 128, 64, False, False, 0),
    "d64_partial": (1, 2, 2, 50, 50, 64, False, False, 0),
    "d64_clip": (2, 16, 16, 577, 577, 64, False, True, 0),
    "d96_one_tile": (1, 1, 1, 128, 128, 96, False, False, 0),
    "d96_iv2_pad": (1, 16, 16, 2049, 2049, 96, False, False, 88),
    "d96_causal": (1, 8, 8, 1000, 1000, 96, True, False, 0),
    "d96_causal_long": (1, 32, 32, 3484, 3484, 96, True, False, 0),
    "d128_gqa_causal": (1, 8, 2, 300, 300, 128, True, False, 0),
    "d128_two_tiles": (1, 2, 2, 256, 256, 128, False, False, 0),
    "d96_q1": (1, 4, 4, 1, 130, 96, True, False, 0),
    "d96_iv2_b12": (12, 16, 16, 2049, 2049, 96, False, False, 88),
    "d64_clip_b12": (12, 16, 16, 577, 577, 64, False, True, 0),
    "d128_llama_long": (1, 32, 8, 2380, 2380, 128, True, False, 0),
    "d96_odd_sub": (1, 4, 4, 700, 700, 96, False, False, 0),
    "d96_causal_offset": (1, 4, 4, 300, 1000, 96, True, False, 0),
}


def run_case(name):
    import torch
    from gvl import ops
    B, H, KVH, Sq, Skv, D, causal, rs, o_dim = CASES[name]
    g = torch.Generator(device="cuda").manual_seed(1)
    q = torch.randn(B, Sq, H, D, device="cuda", generator=g).bfloat16()
    k = torch.randn(B, Skv, KVH, D, device="cuda", generator=g).bfloat16()
    v = torch.randn(B, Skv, KVH, D, device="cuda", generator=g).bfloat16()
    d_real = o_dim if o_dim else D
    if o_dim:
        q[..., o_dim:] = 0
        k[..., o_dim:] = 0
        v[..., o_dim:] = 0
    scale = d_real ** -0.5
    o = ops.attention(q, k, v, scale, causal=causal, round_scores=rs, o_dim=o_dim)
    torch.cuda.synchronize()
    rep = H // KVH
    qf = q.float().permute(0, 2, 1, 3)
    kf = k.float().permute(0, 2, 1, 3).repeat_interleave(rep, dim=1)
    vf = v.float().permute(0, 2, 1, 3).repeat_interleave(rep, dim=1)
    s = qf @ kf.transpose(-1, -2)
    if rs:
        s = s.bfloat16().float()
    s = s * scale
    if causal:
        i = torch.arange(Sq, device="cuda")[:, None]
        j = torch.arange(Skv, device="cuda")[None, :]
        s = s.masked_fill(j > i + (Skv - Sq), float("-inf"))
    ref = (torch.softmax(s, dim=-1) @ vf).permute(0, 2, 1, 3)[..., :d_real]
    err = (o.float() - ref).abs()
    # where is the error? per q-row-block / per d-chunk maxima help to diagnose descriptor mistakes
    e_rows = err.amax(dim=(0, 2, 3))
    e_cols = err.amax(dim=(0, 1, 2))
    worst_rows = [float(e_rows[i:i + 32].max()) for i in range(0, min(Sq, 256), 32)]
    worst_cols = [float(e_cols[i:i + 16].max()) for i in range(0, d_real, 16)]
    msg = "err=%.4g ref_absmax=%.4g nan=%d rows32=%s cols16=%s" % (
        err.max().item(), ref.abs().max().item(), int(torch.isnan(o.float()).sum()),
        ["%.2g" % x for x in worst_rows], ["%.2g" % x for x in worst_cols])
    if Sq >= 512:
        for _ in range(3):
            ops.attention(q, k, v, scale, causal=causal, round_scores=rs, o_dim=o_dim)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(10):
            ops.attention(q, k, v, scale, causal=causal, round_scores=rs, o_dim=o_dim)
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / 10
        fl = 4.0 * B * H * Sq * Skv * d_real * (0.5 if causal else 1.0)
        msg += " time=%.3fms %.0fTF" % (ms, fl / ms / 1e9)
    return msg


def main():
    if os.environ.get("PROBE_CHILD"):
        print("RESULT %s :: %s" % (sys.argv[1], run_case(sys.argv[1])))
        return
    names = sys.argv[1:] or list(CASES)
    for name in names:
        try:
            r = subprocess.run([sys.executable, __file__, name], capture_output=True, text=True, timeout=120,
                               env=dict(os.environ, PROBE_CHILD="1"))
            txt = r.stdout + r.stderr
            if "RESULT" in txt:
                print(txt[txt.index("RESULT"):].strip().splitlines()[0])
            else:
                print("FAIL %s rc=%d :: %s" % (name, r.returncode, " | ".join(txt.strip().splitlines()[-6:])))
        except subprocess.TimeoutExpired:
            print("TIMEOUT %s" % name)
        sys.stdout.flush()


if __name__ == "__main__":
    main()
