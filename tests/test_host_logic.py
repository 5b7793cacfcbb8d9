"""Product-side host logic (gvl/hostlogic.py, gvl/dist.py, gvl/weights.py) -- no GPU, no oracle in the product path;
the oracle / goldens are only the checker here."""
import json
import os

import pytest
import torch

from gvl import dist as gdist
from gvl import hostlogic as H
from gvl import weights
from oracle import gvl_oracle as O


class Tok:
    bos_token_id = 1

    def __init__(self, bos):
        self.bos = bos

    def __call__(self, s):
        ids = ([1] if self.bos else []) + [3 + (ord(c) % 50) for c in s]
        return type("E", (), {"input_ids": ids})()


def test_golden_host_logic(gold_dir):
    g = json.load(open(os.path.join(gold_dir, "host_logic.json")))
    for e in g["parse"]:
        assert H.parse_time_interval(e["text"], e["duration"], 300, e["llm"]) == e["result"]
    for e in g["frame_indices"]:
        assert H.get_frame_indices(e["n"], e["vlen"]) == e["result"]
    for e in g["referring"]:
        assert H.seconds_to_token_inference(e["query"], e["duration"]) == e["result"]
    for e in g["training_quant"]:
        assert H.seconds_to_token_training(e["t"], e["duration"]) == e["result"]
    for e in g["tokenize"]:
        assert H.tokenizer_image_token(e["prompt"], Tok(e["bos"])) == e["result"]
    dur = g["duration"]
    for k, v in g["readme_kat"].items():
        assert H.parse_time_interval(k, dur) == " %s seconds" % v
    assert H.spatial_keyframes(96, 12) == [4 + 8 * i for i in range(12)]


def test_temporal_token_roundtrip_all_tokens():
    # decode(k) then re-encode never drifts by more than one token, and k=0 / k=300 are fixed points
    dur = 142.04187520854188
    for k in range(301):
        txt = H.parse_time_interval("<%d>" % k, dur)
        sec = float(txt.split()[0])
        back = H.seconds_to_token_training(sec, dur)
        assert abs(back - k) <= 1
    assert H.seconds_to_token_training(dur * 2, dur) == 300


def test_left_pad_matches_oracle_and_truncates_tail():
    lists = [[1, 5, -200, 7], [1, 2, 3, 4, 5, 6, -200, 8, 9], [4]]
    ids, mask = H.left_pad(lists, 0, 2048)
    oi, om = O.left_pad_batch(lists, 0, 2048)
    assert torch.equal(ids, oi) and torch.equal(mask, om)
    ids, mask = H.left_pad(lists, 0, 5)
    oi, om = O.left_pad_batch(lists, 0, 5)
    assert torch.equal(ids, oi) and torch.equal(mask, om)
    assert ids[1].tolist() == [5, 6, -200, 8, 9]
    with pytest.raises(ValueError):
        H.tokenizer_image_token("x", Tok(True), return_tensors="np")


def test_rope_tables_match_oracle():
    r = O.phi35_rope_cfg(96)
    for use_long in (False, True):
        cos, sin = H.longrope_tables(300, 96, r["base"], r["short_factor"], r["long_factor"], r["max_pos"], r["orig_max_pos"], use_long)
        oc, osn = O.phi3_rope_tables(torch.arange(300), 96, r["base"], r["short_factor"], r["long_factor"], r["max_pos"],
                                     r["orig_max_pos"], seq_len=5000 if use_long else 300)
        assert torch.equal(cos.float(), O.bf(oc)) and torch.equal(sin.float(), O.bf(osn))
    cos, sin = H.plain_rope_tables(3000, 128, 500000.0, bf16_matmul_quirk=True)
    oc, osn = O.plain_rope_tables(torch.arange(3000), 128, 500000.0, bf16_matmul_quirk=True)
    assert torch.equal(cos.float(), oc) and torch.equal(sin.float(), osn)


def test_partition_units():
    for n in (12, 384, 5, 1):
        for ws in (1, 2, 4, 8):
            parts = gdist.partition_units(n, ws)
            assert sum(c for _, c in parts) == n
            assert [s for s, _ in parts] == [sum(c for _, c in parts[:i]) for i in range(ws)]
            assert max(c for _, c in parts) - min(c for _, c in parts) <= 1
    assert [c for _, c in gdist.partition_units(12, 8)] == [2, 2, 2, 2, 1, 1, 1, 1]
    assert gdist.clips_for_rank(32, 3, 8) == [3, 11, 19, 27]


def test_gate_up_interleave_and_clip_pack_cpu():
    g = torch.arange(256 * 2).float().reshape(256, 2)
    u = -torch.arange(256 * 2).float().reshape(256, 2)
    w = weights.interleave_gate_up(g, u)
    assert torch.equal(w[:128], g[:128]) and torch.equal(w[128:256], u[:128]) and torch.equal(w[256:384], g[128:])
    P = O.make_clip_params(dim=64, heads=4, ffn=128, layers=2, image=56, seed=1)
    pk = weights.pack_clip(P, 4, 1, device="cpu", image=56)
    assert pk.struct.kpad == 640 and pk.struct.n_patch == 16 and pk.struct.n_layers == 1
    with pytest.raises(ValueError):
        weights.pack_clip(O.make_clip_params(dim=96, heads=4, ffn=128, layers=1, image=56), 4, 1, device="cpu", image=56)


def test_warp_logits_matches_transformers_warpers():
    """hostlogic.warp_logits vs the third-party library the reference calls (transformers LogitsWarper classes)."""
    lp = pytest.importorskip("transformers.generation.logits_process")
    g = torch.Generator().manual_seed(5)
    scores = torch.randn(3, 500, generator=g) * 3
    scores[0, 10] = scores[0, 11]                                   # a tie
    ids = torch.zeros(3, 1, dtype=torch.long)
    for temperature, top_k, top_p in [(0.2, 50, None), (1.0, 50, 0.9), (0.7, 0, 0.5), (0.2, 5, 0.95), (None, 1, None)]:
        ref = scores.clone()
        if temperature is not None and temperature != 1.0:
            ref = lp.TemperatureLogitsWarper(temperature)(ids, ref)
        if top_k:
            ref = lp.TopKLogitsWarper(top_k=top_k)(ids, ref)
        if top_p is not None:
            ref = lp.TopPLogitsWarper(top_p=top_p)(ids, ref)
        got = H.warp_logits(scores.clone(), temperature, top_k, top_p)
        assert torch.equal(torch.isinf(got), torch.isinf(ref)), (temperature, top_k, top_p)
        assert torch.equal(torch.nan_to_num(got, neginf=0.0), torch.nan_to_num(ref, neginf=0.0))


def test_read_frames_on_the_reference_demo_video():
    """README.md:90-94 transcript video: 4257 frames at 29.97003 fps -> duration 142.0419 s, 96 'middle' indices 21, 65, ..., 4234
    (the same numbers the README timestamps are computed from). Runs where the reference checkout is mounted."""
    path = "/root/reference/experiments/_3klvlS4W7A.mp4"
    if not os.path.exists(path):
        pytest.skip("reference demo video not present")
    pytest.importorskip("cv2")
    from gvl import video
    frames, idx, fps, vlen, duration = video.read_frames(path, 8, sample="middle")
    assert vlen == 4257 and abs(fps - 29.97003) < 1e-4 and abs(duration - 142.0419) < 1e-3
    assert idx == H.get_frame_indices(8, 4257, sample="middle") and len(idx) == 8
    assert frames.dtype == torch.uint8 and frames.shape[:2] == (8, 3) and frames.float().std() > 1.0
    full = H.get_frame_indices(96, 4257, sample="middle")
    assert full[:4] == [21, 65, 110, 154] and full[-2:] == [4189, 4234]
    assert "%.2f" % (duration * 30 / 300) == "14.20" and "%.2f" % (duration * 239 / 300) == "113.16"


def test_build_prompt_matches_reference_templates(gold_dir):
    """Goldens: the reference's own Template classes exec'd from datasets/chat/base_template.py (oracle/make_golden.py)."""
    cases = json.load(open(os.path.join(gold_dir, "prompts.json")))
    assert len(cases) == 9
    for e in cases:
        assert H.build_prompt(e["llm"], e["mode"], e["text"], e["duration"]) == e["prompt"], (e["llm"], e["mode"])
    with pytest.raises(ValueError):
        H.build_prompt("phi3.5", "caption", "x")


def test_stacked_rows_have_the_shape_and_padding_of_hf_generate():
    """CausalLM decodes the rows of a batch independently (each until its own EOS) and `_stack_rows` assembles HF generate's [B, L]:
    L = the step at which the LAST row hit EOS (or max_new_tokens), finished rows padded (GenerationMixin greedy loop as driven at
    llava_next_video.py:655-661: next = next * unfinished + pad * (1 - unfinished); stop when every row is finished)."""
    import types
    from gvl import model
    eos, pad, V = 2, 0, 11

    def hf_loop(script, max_new):
        B = len(script)
        unfinished = torch.ones(B, dtype=torch.long)
        cols = []
        for t in range(max_new):
            nxt = torch.tensor([script[b][t] for b in range(B)])
            nxt = nxt * unfinished + pad * (1 - unfinished)
            cols.append(nxt)
            unfinished = unfinished * (nxt != eos).long()
            if unfinished.max() == 0:
                break
        return torch.stack(cols, dim=1)

    g = torch.Generator().manual_seed(5)
    fake = types.SimpleNamespace(device="cpu", vocab=V)
    for case in range(40):
        B, max_new = int(torch.randint(1, 5, (1,), generator=g)), int(torch.randint(1, 9, (1,), generator=g))
        script = [[int(x) for x in torch.randint(1, 6, (max_new,), generator=g)] for _ in range(B)]      # eos = 2 appears often
        want = hf_loop(script, max_new)
        outs = []
        for b in range(B):                                   # what one row's decode returns: its tokens up to and including its EOS
            row = script[b][:max_new]
            n = row.index(eos) + 1 if eos in row else len(row)
            outs.append(torch.tensor(row[:n]))
        got = model.CausalLM._stack_rows(fake, outs, [None] * B, eos, pad, False)
        assert got.tolist() == want.tolist(), (case, script)
    # without an EOS id every row runs to max_new_tokens
    outs = [torch.tensor([4, 2, 5]), torch.tensor([2, 2, 2])]
    assert model.CausalLM._stack_rows(fake, outs, [None, None], None, pad, False).tolist() == [[4, 2, 5], [2, 2, 2]]
