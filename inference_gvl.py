#!/usr/bin/env python
"""inference.py of the reference (WHB139426/Grounded-Video-LLM, inference.py:14-215) on the gvl-b200 path: same flags (--dtype accepts bfloat16 only; --model / --stage /
--lora / --attn_implementation are accepted and have one implementation here), same three
examples (temporal grounding, referring, video QA), same prompt construction and timestamp decoding.

    video file --(gvl.video.read_frames: 96 'middle' frames)--> uint8 frames on the GPU
               --(gvl.preprocess.create_pixel_inputs: Pillow-bit-exact resize / crop / normalise kernels)--> pixel tensors
               --(gvl.model.LLAVA_NEXT_VIDEO.generate: CLIP + InternVideo2 + projectors + LLM prefill / decode)--> text
               --(gvl.hostlogic.parse_time_interval)--> "From 14.20 seconds to 25.09 seconds."

With the real weights:   python inference_gvl.py --video_path clip.mp4 --ckpt_path ... (paths as in the reference README)
Without (this container has neither weights nor a tokenizer): --synthetic runs the same flow on a reduced-depth random-init model
with a byte-level stand-in tokenizer, which exercises every stage but produces meaningless text.
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "grounded-video-llm_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)


def _dtype_flag(s):
    name = str(s).replace("torch.", "").lower()
    if name not in ("bfloat16", "bf16"):
        raise argparse.ArgumentTypeError("the gvl-b200 path computes in bfloat16 only (reference default, inference.py:18); got %r" % s)
    return "bfloat16"


def parse_args(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--device", default="cuda")
    ap.add_argument("--seed", type=int, default=42)
    # flags of the reference CLI that select things this path fixes (inference.py:18-29): accepted and validated, so a reference
    # command line keeps working; bf16 is the only arithmetic the kernels implement, attention is always the fused kernel
    ap.add_argument("--dtype", default="torch.bfloat16", type=_dtype_flag)
    ap.add_argument("--model", default="llava_next_video", choices=["llava_next_video"])
    ap.add_argument("--stage", default="sft", choices=["pretrain", "grounded", "sft"])
    ap.add_argument("--lora", type=lambda s: str(s).lower() not in ("0", "false", "no", ""), default=True)
    ap.add_argument("--attn_implementation", default="flash_attention_2", choices=["eager", "flash_attention_2"])
    ap.add_argument("--llm", default="phi3.5", choices=["llama3", "vicuna", "phi3.5"])
    ap.add_argument("--max_txt_len", type=int, default=2048)
    ap.add_argument("--num_temporal_tokens", type=int, default=300)
    ap.add_argument("--num_frames", type=int, default=96)
    ap.add_argument("--num_segs", type=int, default=12)
    ap.add_argument("--config_path", default="weight_path/Phi-3.5-vision-instruct")
    ap.add_argument("--tokenizer_path", default="weight_path/Phi-3.5-mini-instruct")
    ap.add_argument("--pretrained_video_path", default="weight_path/internvideo/vision-encoder-InternVideo2-stage2_1b-224p-f4.pt")
    ap.add_argument("--pretrained_vision_proj_llm_path", default="weight_path/Phi-3.5-vision-instruct-seperated/")
    ap.add_argument("--ckpt_path", default="weight_path/ckpt/sft_llava_next_video_phi3.5_mix_sft_multi_modal_projector_video_projecter_language_model.pth")
    ap.add_argument("--prompt_grounding", default="Give you a textual query: 'The female host wearing purple clothes is reporting news in the "
                                                  "studio'. When does the described content occur in the video? Please return the start and end timestamps.")
    ap.add_argument("--prompt_videoqa", default="Question: What does this TV news report about?\nOptions:\n(A) thievery\n(B) community "
                                                "violence incidents\n(C) fashion show\n(D) aging population")
    ap.add_argument("--prompt_referring", default="What is happening from 70 seconds to 80 seconds?")
    ap.add_argument("--video_path", default="./experiments/_3klvlS4W7A.mp4")
    ap.add_argument("--do_sample", type=lambda s: str(s).lower() not in ("0", "false", "no"), default=True)
    ap.add_argument("--num_beams", type=int, default=1)
    ap.add_argument("--max_new_tokens", type=int, default=2048)
    ap.add_argument("--temperature", type=float, default=0.2)
    ap.add_argument("--top_p", type=float, default=None)
    ap.add_argument("--synthetic", action="store_true", help="random-init reduced-depth model + stand-in tokenizer (no weights needed)")
    args = ap.parse_args(argv)
    if args.llm == "vicuna":
        ap.error("--llm vicuna: only the phi3.5 and llama3 checkpoints of the reference README are on the gvl-b200 path")
    return args


class ByteTokenizer:
    """Stand-in for --synthetic: bytes -> ids 3..258, temporal tokens <k> appended after them like tokenizer.add_tokens does."""
    bos_token_id, eos_token_id, pad_token_id = 1, 2, 0

    def __init__(self, n_temporal=300):
        self.base = 259
        self.n_temporal = n_temporal

    def __len__(self):
        return self.base + self.n_temporal + 2

    def __call__(self, text):
        return type("Enc", (), {"input_ids": [self.bos_token_id] + [3 + b for b in text.encode("utf-8")]})()

    def batch_decode(self, ids, skip_special_tokens=True):
        out = []
        for row in (ids.tolist() if hasattr(ids, "tolist") else ids):
            s = []
            for t in row:
                if 3 <= t < self.base:
                    s.append(chr(t - 3) if t - 3 < 128 else "?")
                elif self.base <= t <= self.base + self.n_temporal:
                    s.append("<%d>" % (t - self.base))
            out.append("".join(s))
        return out


def build_model(args):
    import torch
    from gvl import ingest, model, synth
    if args.synthetic:
        tok = ByteTokenizer(args.num_temporal_tokens)
        params, lm_cfg, clip_cfg, iv2_cfg = synth.make_params(
            "phi3.5", device="cpu", seed=args.seed,
            lm=dict(synth.PHI35, layers=2, vocab=len(tok), dim=512, heads=8, kv_heads=8, head_dim=64, ffn=1024),
            clip=dict(synth.CLIP_L336, layers=3), iv2=dict(synth.IV2_1B, depth=3, gamma=0.1), lm_dtype=torch.float32)
        return model.LLAVA_NEXT_VIDEO(params, llm="phi3.5", tokenizer=tok, num_frames=args.num_frames, num_segs=args.num_segs,
                                      max_txt_len=args.max_txt_len, lm_cfg=lm_cfg, clip_cfg=clip_cfg, iv2_cfg=iv2_cfg,
                                      max_new_tokens=args.max_new_tokens, device=args.device), tok
    from transformers import AutoTokenizer
    tok = AutoTokenizer.from_pretrained(args.tokenizer_path, use_fast=args.llm == "phi3.5")
    if args.llm == "phi3.5":
        tok.pad_token = "<|end|>"                                            # llava_next_video.py:113
    else:
        tok.eos_token_id, tok.pad_token_id = 128009, 128001                 # llava_next_video.py:102-103
    tok.add_tokens(["<%d>" % i for i in range(args.num_temporal_tokens + 1)] + ["<timestamp_grounding>"])   # :234-236
    # the LM is built from <pretrained_vision_proj_llm_path>/language_model_seperated/config.json (llava_next_video.py:146-148);
    # config_path only supplies the vision config
    lm_cfg_path = os.path.join(args.pretrained_vision_proj_llm_path, "language_model_seperated", "config.json")
    cfg = json.load(open(lm_cfg_path if os.path.exists(lm_cfg_path) else os.path.join(args.config_path, "config.json")))
    if args.llm == "phi3.5":
        rs = cfg["rope_scaling"]
        lm_cfg = dict(arch="phi3", heads=cfg["num_attention_heads"], kv_heads=cfg["num_key_value_heads"],
                      head_dim=cfg["hidden_size"] // cfg["num_attention_heads"], eps=cfg["rms_norm_eps"],
                      rope=dict(type="longrope", base=cfg["rope_theta"], short_factor=rs["short_factor"], long_factor=rs["long_factor"],
                                max_pos=cfg["max_position_embeddings"], orig_max_pos=cfg["original_max_position_embeddings"]))
    else:
        t = cfg.get("text_config", cfg)
        lm_cfg = dict(arch="llama", heads=t["num_attention_heads"], kv_heads=t["num_key_value_heads"],
                      head_dim=t["hidden_size"] // t["num_attention_heads"], eps=t["rms_norm_eps"],
                      rope=dict(type="plain", base=t.get("rope_theta", 500000.0), bf16_quirk=True))
    params = ingest.load_params(args.llm, args.pretrained_vision_proj_llm_path, args.pretrained_video_path, args.ckpt_path,
                                args.num_frames, args.num_segs)
    return model.LLAVA_NEXT_VIDEO(params, llm=args.llm, tokenizer=tok, num_frames=args.num_frames, num_segs=args.num_segs,
                                  max_txt_len=args.max_txt_len, lm_cfg=lm_cfg, max_new_tokens=args.max_new_tokens,
                                  device=args.device), tok


def create_inputs(args, mode, pixels, duration):
    """inference.py:65-123: the pixel tensors are shared by the three modes, only the prompt differs."""
    from gvl import hostlogic
    text = {"grounding": args.prompt_grounding, "qa": args.prompt_videoqa, "referring": args.prompt_referring}[mode]
    prompt = hostlogic.build_prompt(args.llm, mode, text, duration, args.num_temporal_tokens)
    return {"video_ids": [args.video_path], "question_ids": [args.video_path], "prompts": [prompt],
            "temporal_pixel_values": pixels["temporal_pixel_values"], "spatial_pixel_values": pixels["spatial_pixel_values"]}


def _synthetic_clip(path, n=300, fps=25.0, size=(480, 360)):
    """--synthetic without a video file: write a small moving-pattern mp4 so that decode + frame sampling still run."""
    import cv2
    import numpy as np
    w = cv2.VideoWriter(path, cv2.VideoWriter_fourcc(*"mp4v"), fps, size)
    if not w.isOpened():
        raise RuntimeError("cannot create the synthetic clip %s" % path)
    for i in range(n):
        f = np.zeros((size[1], size[0], 3), np.uint8)
        f[:, :, 0] = (i * 255) // n
        cv2.circle(f, (40 + (i * (size[0] - 80)) // n, size[1] // 2), 30, (0, 255, 255), -1)
        w.write(f)
    w.release()
    return path


def main(argv=None):
    args = parse_args(argv)
    import torch
    from gvl import hostlogic, preprocess, video
    torch.manual_seed(args.seed)
    if not torch.cuda.is_available():
        raise RuntimeError("inference_gvl.py needs a CUDA device: the gvl hot path has no CPU fallback")
    if args.synthetic and not os.path.exists(args.video_path):
        import tempfile
        args.video_path = _synthetic_clip(os.path.join(tempfile.gettempdir(), "gvl_synthetic_clip.mp4"))
    model, _tok = build_model(args)
    frames, _idx, _fps, _vlen, duration = video.read_frames(args.video_path, args.num_frames, sample="middle")
    pixels = preprocess.create_pixel_inputs(frames.to(args.device), args.num_frames, args.num_segs)
    gen = {"do_sample": args.do_sample, "num_beams": args.num_beams, "max_new_tokens": args.max_new_tokens,
           "temperature": args.temperature, "top_p": args.top_p}
    results = {}
    for mode in ("grounding", "qa", "referring"):
        samples = create_inputs(args, mode, pixels, duration)
        results[mode] = (samples["prompts"][0], model.generate(samples, **gen)[0])
    print("\n******grounding example******")
    print(results["grounding"][0])
    print(hostlogic.parse_time_interval(results["grounding"][1], duration, args.num_temporal_tokens, args.llm))
    print("\n******referring example******")
    print(results["referring"][0])
    print(results["referring"][1])
    print("\n******videoqa example******")
    print(results["qa"][0])
    print(results["qa"][1])
    return results


if __name__ == "__main__":
    main()
