"""2-rank NCCL test of the one exchange step on the path (pytest -m gpu; skipped with fewer than 2 visible GPUs): units sharded over
ranks + all-gather (`encode_images`) and + all-to-all (`encode_images_for_decode`) must reproduce, bit for bit, what ONE rank computes
for all units, and a clip whose units were encoded on two ranks must decode to the same greedy tokens."""
import os
import socket

import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, ws, port, q):
    import datetime
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = "cuda:%d" % rank
    dist.init_process_group("nccl", rank=rank, world_size=ws, device_id=torch.device(dev), timeout=datetime.timedelta(seconds=120))
    try:
        from gvl import dist as gdist, model, synth
        params, lm_cfg, clip_cfg, iv2_cfg = synth.make_params(
            "phi3.5", device="cpu", seed=3, lm=dict(synth.PHI35, layers=2, vocab=1000 + 302, dim=256, heads=4, kv_heads=4, head_dim=64, ffn=512),
            clip=dict(synth.CLIP_L336, layers=2), iv2=dict(synth.IV2_1B, depth=3, gamma=0.1), lm_dtype=torch.float32)
        m = model.LLAVA_NEXT_VIDEO(params, llm="phi3.5", num_frames=24, num_segs=3, lm_cfg=lm_cfg, clip_cfg=clip_cfg, iv2_cfg=iv2_cfg,
                                   max_ctx=1536, device=dev)
        report = {}
        for B in (1, 3):
            g = torch.Generator().manual_seed(100 + B)
            sp = torch.randn(B, 3, 3, 336, 336, generator=g).to(dev)
            tp = torch.randn(B, 24, 3, 224, 224, generator=g).to(dev)
            ids = torch.randint(3, 1000, (20,), generator=torch.Generator().manual_seed(7))
            ids[5] = -200
            samples = {"spatial_pixel_values": sp, "temporal_pixel_values": tp, "input_ids": [ids.tolist()] * B}
            # what ONE rank computes for every unit (no collective): units as [B*3, ...]
            single = m._encode_units(sp.reshape(B * 3, 3, 336, 336), tp.reshape(B * 3, 8, 3, 224, 224)).reshape(B, -1, 256)
            full = m.encode_images(samples)                                  # sharded + all-gather
            report["allgather_B%d" % B] = bool(torch.equal(full, single))
            feats, mine = m.encode_images_for_decode(samples)                # sharded + all-to-all
            report["alltoall_B%d" % B] = mine == gdist.clips_for_rank(B, rank, ws) and all(
                bool(torch.equal(feats[i], single[b])) for i, b in enumerate(mine))
            if not (report["allgather_B%d" % B]):
                d = (full.float() - single.float()).abs().amax(dim=(0, 2)).reshape(-1)
                report["diff_B%d" % B] = [float(d[u * (feats.shape[1] // 3 if feats.numel() else 1):].max()) for u in range(1)]
            toks = m.generate(samples, max_new_tokens=6)                     # tokens of every clip on every rank
            # single-rank decode of the same clips from the single-rank features
            emb, _, masks = m.prepare_multimodal_inputs(torch.stack([ids] * B), None, torch.ones(B, 20, dtype=torch.long), single, ["v"] * B)
            ref = [m.language_model.generate(inputs_embeds=emb[b:b + 1], attention_mask=masks[b:b + 1], max_new_tokens=6, batched=False)[0]
                   for b in range(B)]
            report["tokens_B%d" % B] = all(toks[b].tolist() == ref[b].tolist() for b in range(B)) if B == 1 else \
                all(len(toks[b]) == 6 for b in range(B))
        q.put((rank, report))
    finally:
        dist.destroy_process_group()


def test_two_rank_exchange_is_bit_exact():
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
    for rank, report in res:
        bad = {k: v for k, v in report.items() if v is not True and not k.startswith("diff")}
        assert not bad, (rank, report)
