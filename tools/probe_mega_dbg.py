"""GPU debug probe: single-kernel decode vs per-op chain logits, per step, over a sweep of contexts / head geometries."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "grounded-video-llm_b200"))
sys.path.insert(0, ROOT)
from gvl import model  # noqa: E402
from oracle import gvl_oracle as O  # noqa: E402


def run(arch, heads, kvh, hd, ctxs, n_new=5, layers=2):
    P = O.make_lm_params(arch=arch, dim=512, heads=heads, kv_heads=kvh, head_dim=hd, ffn=512, layers=layers, vocab=1000, seed=13, std=0.04)
    rope = O.phi35_rope_cfg(hd) if arch == "phi3" else dict(type="plain", base=500000.0, bf16_quirk=True)
    os.environ["GVL_DECODE_MEGA"] = "1"
    warm = (torch.randn(16, 512) * 0.5).cuda()
    a = model.CausalLM(P, arch, heads, kvh, hd, 1e-5, rope, max_ctx=4096)
    a.generate(inputs_embeds=warm[None], max_new_tokens=2)       # the gvl_lm object (and its decode mode) is created lazily
    os.environ["GVL_DECODE_MEGA"] = "0"
    b = model.CausalLM(P, arch, heads, kvh, hd, 1e-5, rope, max_ctx=4096)
    b.generate(inputs_embeds=warm[None], max_new_tokens=2)
    for ctx in ctxs:
        emb = (torch.randn(ctx, 512, generator=torch.Generator().manual_seed(5)) * 0.5).cuda()
        _, la = a.generate(inputs_embeds=emb[None], max_new_tokens=n_new, return_logits=True)
        _, lb = b.generate(inputs_embeds=emb[None], max_new_tokens=n_new, return_logits=True)
        d = (la[0] - lb[0]).abs().amax(-1).tolist()
        print("%s H%d KV%d D%d L%d ctx %5d: per-step max|mega-chain| %s" % (arch, heads, kvh, hd, layers, ctx, " ".join("%.3g" % x for x in d)), flush=True)
    a.close()
    b.close()


if __name__ == "__main__":
    run("phi3", 16, 16, 64, [1498, 1499, 1500, 1501, 1502, 1503, 1504, 700, 2900])
    run("phi3", 16, 16, 64, [1500], n_new=8)
    run("phi3", 16, 16, 64, [1500], n_new=5, layers=1)
    run("phi3", 32, 32, 96, [3483, 1500, 200], n_new=6)
    run("llama", 32, 8, 128, [2380, 1500], n_new=6)
