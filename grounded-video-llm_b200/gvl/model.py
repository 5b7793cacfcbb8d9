"""Host-side mirror of the reference's module API for the inference forward path.

Same class / method names and argument meaning as the reference:
    CLIPVisionModel.forward(pixel_values, output_hidden_states=True).hidden_states[-2]   modeling_clip.py:904-940
    PretrainInternVideo2.forward(x, mask, use_image, x_vis_return_idx, x_vis_only)        internvideo2.py:970-1040
    Phi3_5_Projecter / Video_Projecter .forward                                           llava_next_video.py:26-54
    CausalLM.forward(inputs_embeds=...) / .generate(inputs_embeds=..., ...)               modeling_phi3.py:1466-1551
    LLAVA_NEXT_VIDEO.encode_images / prepare_multimodal_inputs / generate                 llava_next_video.py:491-666
All arithmetic happens in libgvl.so through the C ABI (include/gvl.h); torch provides device memory, streams and
torch.distributed. There is no CPU path: inputs that are not CUDA tensors are moved to the device first.
"""
import ctypes
from types import SimpleNamespace

import torch

from . import _lib, dist as gdist, hostlogic, ops, weights


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _hf_output(kind, **fields):
    """The reference returns transformers' ModelOutput classes (BaseModelOutputWithPooling, modeling_clip.py:868-872;
    CausalLMOutputWithPast, modeling_phi3.py:1545-1551). Use them when transformers is importable (it always is next to the reference),
    a plain namespace with the same attribute names otherwise."""
    try:
        from transformers import modeling_outputs as mo
        return getattr(mo, kind)(**fields)
    except Exception:                                            # noqa: BLE001 -- transformers absent or incompatible field set
        return SimpleNamespace(**fields)


class CLIPVisionModel:
    """Spatial stream. Only hidden_states[-2] is materialised (the one consumer reads, llava_next_video.py:505):
    the 24th layer and post_layernorm the reference also runs are dead work and are skipped."""

    def __init__(self, state_dict, num_heads=16, num_layers=24, image_size=336, device="cuda"):
        self.device = torch.device(device)
        self.num_layers = num_layers
        self.pk = weights.pack_clip(state_dict, num_heads, num_layers - 1, device=self.device, image=image_size)
        self.w = self.pk.struct
        self.dim = self.w.dim
        self.tokens = self.w.n_patch + 1
        self._ws = None

    def forward(self, pixel_values, output_attentions=None, output_hidden_states=None, return_dict=None):
        if pixel_values is None:
            raise ValueError("You have to specify pixel_values")     # modeling_clip.py:846-847
        if not output_hidden_states:
            raise ValueError("gvl CLIPVisionModel serves hidden_states[-2] only; call with output_hidden_states=True")
        lib = _lib.load()
        pix = pixel_values.to(self.device, torch.float32).contiguous()
        n = pix.shape[0]
        if pix.shape[1:] != (3, self.w.image, self.w.image):
            raise ValueError("pixel_values must be [N,3,%d,%d]" % (self.w.image, self.w.image))
        need = lib.gvl_clip_workspace(ctypes.byref(self.w), n)
        if self._ws is None or self._ws.numel() < need:
            self._ws = torch.empty((need,), dtype=torch.uint8, device=self.device)
        hs = torch.empty((n, self.tokens, self.dim), dtype=torch.float32, device=self.device)
        rc = lib.gvl_clip_encode(ctypes.byref(self.w), ctypes.c_void_p(pix.data_ptr()), ctypes.c_void_p(hs.data_ptr()),
                                 n, ctypes.c_void_p(self._ws.data_ptr()), self._ws.numel(), _stream())
        _lib.check(rc, "gvl_clip_encode")
        states = [None] * (self.num_layers + 1)
        states[-2] = hs
        return _hf_output("BaseModelOutputWithPooling", last_hidden_state=None, pooler_output=None, hidden_states=tuple(states),
                          attentions=None)

    __call__ = forward


class PretrainInternVideo2:
    """Temporal stream; serves the x_vis_only=True call the VLM makes (llava_next_video.py:532)."""

    def __init__(self, state_dict, num_heads=16, depth=40, num_frames=8, device="cuda"):
        self.device = torch.device(device)
        self.depth = depth
        self.num_frames = num_frames
        self.heads = num_heads
        self.sd = state_dict
        self._packs = {}
        self._ws = None

    def _pack(self, n_blocks):
        if n_blocks not in self._packs:
            self._packs[n_blocks] = weights.pack_iv2(self.sd, self.heads, n_blocks, self.num_frames, device=self.device)
        return self._packs[n_blocks]

    def forward(self, x, mask=None, use_image=False, x_vis_return_idx=-1, x_vis_only=False):
        if mask is not None or use_image or not x_vis_only:
            raise NotImplementedError("only forward(x, None, False, x_vis_return_idx, x_vis_only=True) is on the path")
        lib = _lib.load()
        pk = self._pack(self.depth + x_vis_return_idx + 1)          # blocks 0..depth+idx (internvideo2.py:1028-1030)
        w = pk.struct
        pix = x.to(self.device, torch.float32).contiguous()
        n = pix.shape[0]
        if pix.shape[1:] != (3, self.num_frames, 224, 224):
            raise ValueError("x must be [N,3,%d,224,224]" % self.num_frames)
        need = lib.gvl_iv2_workspace(ctypes.byref(w), n)
        if self._ws is None or self._ws.numel() < need:
            self._ws = torch.empty((need,), dtype=torch.uint8, device=self.device)
        out = torch.empty((n, 1 + self.num_frames * 256, w.dim), dtype=torch.bfloat16, device=self.device)
        rc = lib.gvl_iv2_encode(ctypes.byref(w), ctypes.c_void_p(pix.data_ptr()), ctypes.c_void_p(out.data_ptr()), n,
                                ctypes.c_void_p(self._ws.data_ptr()), self._ws.numel(), _stream())
        _lib.check(rc, "gvl_iv2_encode")
        return out

    __call__ = forward


class MLP2Projector:
    """Linear -> GELU(erf) -> Linear with biases: Phi3_5_Projecter / Video_Projecter / LlavaMultiModalProjector."""

    def __init__(self, w0, b0, w1, b1, device="cuda"):
        self.w0, self.b0, self.w1, self.b1 = weights.pack_mlp2(w0, b0, w1, b1, device)

    def forward(self, x):
        shp = x.shape
        x2 = x.reshape(-1, shp[-1])
        if x2.dtype != torch.bfloat16:
            x2 = x2.to(torch.bfloat16)                      # autocast cast at the first Linear
        h = ops.gemm(x2.contiguous(), self.w0, bias=self.b0, act=ops.ACT_GELU_ERF)
        y = ops.gemm(h, self.w1, bias=self.b1)
        return y.reshape(*shp[:-1], y.shape[-1])

    __call__ = forward


class CausalLM:
    """Phi3ForCausalLM / LlamaForCausalLM replacement for one unpadded sequence at a time."""

    def __init__(self, state_dict, arch, num_heads, num_kv_heads, head_dim, rms_eps, rope, max_ctx=4096, device="cuda"):
        self.device = torch.device(device)
        self.arch = arch
        self.rope = dict(rope)
        self.max_ctx = max_ctx
        self.sd = state_dict
        self.cfg = (num_heads, num_kv_heads, head_dim, rms_eps)
        self._lms = {}
        self.vocab = state_dict["lm_head.weight"].shape[0]
        self.dim = state_dict["model.embed_tokens.weight"].shape[1]
        self._shared = None

    def _tables(self, long_from):
        r = self.rope
        hd = self.cfg[2]
        if r["type"] == "longrope":
            return hostlogic.longrope_tables(self.max_ctx, hd, r["base"], r["short_factor"], r["long_factor"],
                                             r["max_pos"], r["orig_max_pos"], False, long_from=long_from)
        return hostlogic.plain_rope_tables(self.max_ctx, hd, r["base"], bf16_matmul_quirk=r.get("bf16_quirk", False))

    def _get(self, long_from, slot=0):
        """One C-side LM object per RoPE table; the weight buffers are shared. LongRoPE (modeling_phi3.py:371-409) picks the
        factor set per forward call from kv_seq_len: `long_from` = first position rotated with long_factor -- 0 after a prompt
        longer than original_max_position_embeddings, original_max_position_embeddings otherwise (cached decode steps past it
        switch to long_factor while the keys already cached keep their rotation; hostlogic.longrope_tables)."""
        key = long_from if slot == 0 else (long_from, slot)       # slot > 0: further sequences of a batch (own KV cache and state)
        if key not in self._lms:
            lib = _lib.load()
            cos, sin = self._tables(long_from)
            h, kvh, hd, eps = self.cfg
            if self._shared is None:
                pk = weights.pack_lm(self.sd, self.arch, h, kvh, hd, eps, self.max_ctx, cos, sin, device=self.device)
                self._shared = pk
            else:
                pk = self._shared
                cos_d = cos.to(self.device)
                sin_d = sin.to(self.device)
                pk.tensors += [cos_d, sin_d]
                pk.struct.rope_cos = ctypes.c_void_p(cos_d.data_ptr())
                pk.struct.rope_sin = ctypes.c_void_p(sin_d.data_ptr())
            handle = ctypes.c_void_p()
            rc = lib.gvl_lm_create_ex(ctypes.byref(pk.struct), 0 if slot == 0 else 1, ctypes.byref(handle))   # 1 = GVL_LM_NO_SINGLE_KERNEL
            _lib.check(rc, "gvl_lm_create_ex")
            self._lms[key] = handle
        return self._lms[key]

    @property
    def embed_table(self):
        self._get(self._long_from(1))
        return self._shared.embed

    def get_input_embeddings(self):
        table = self.embed_table
        return lambda ids: table[ids]

    def _long_from(self, prompt_len):
        r = self.rope
        if r["type"] != "longrope":
            return 0
        return 0 if prompt_len > r["orig_max_pos"] else r["orig_max_pos"]

    def prefill(self, inputs_embeds, want_hidden=False, n_new=0, slot=0):
        """inputs_embeds [S,D] bf16. Returns (last-position fp32 logits [V], hidden [S,D] or None). `n_new` is only a
        capacity hint: generate() clamps the number of steps to the cache (max_new_tokens is a ceiling, not a reservation)."""
        lib = _lib.load()
        emb = inputs_embeds.to(self.device, torch.bfloat16).contiguous()
        S = emb.shape[0]
        if S > self.max_ctx:
            raise ValueError("sequence of %d tokens exceeds max_ctx=%d" % (S, self.max_ctx))
        lm = self._get(self._long_from(S), slot)
        logits = torch.empty((self.vocab,), dtype=torch.float32, device=self.device)
        hidden = torch.empty((S, self.dim), dtype=torch.bfloat16, device=self.device) if want_hidden else None
        rc = lib.gvl_lm_prefill(lm, ctypes.c_void_p(emb.data_ptr()), S, ctypes.c_void_p(logits.data_ptr()),
                                ctypes.c_void_p(hidden.data_ptr() if want_hidden else 0), _stream())
        _lib.check(rc, "gvl_lm_prefill")
        self._active = (lm, S)
        return logits, hidden

    def forward(self, inputs_embeds=None, input_ids=None, **unused):
        """CausalLMOutputWithPast-like: .logits fp32 [1,S,V] for ALL positions (test / parity use; the generate path
        only computes the last row). Batch size 1."""
        if inputs_embeds is None:
            inputs_embeds = self.embed_table[input_ids.to(self.device)]
        emb = inputs_embeds.reshape(-1, inputs_embeds.shape[-1])
        _, hidden = self.prefill(emb, want_hidden=True)
        hn = ops.rmsnorm(hidden, self._final_norm(), self.cfg[3])
        logits = ops.gemm(hn, self._lm_head_padded()[0], bias=self._lm_head_padded()[1])
        logits = logits[:, : self.vocab].float()
        return _hf_output("CausalLMOutputWithPast", loss=None, logits=logits[None], past_key_values=None, hidden_states=(hidden,),
                          attentions=None)

    __call__ = forward

    def _final_norm(self):
        if not hasattr(self, "_fn"):
            self._fn = self.sd["model.norm.weight"].to(self.device, torch.bfloat16).contiguous()
        return self._fn

    def _lm_head_padded(self):
        """vocab (32366) is not a multiple of 8 -> pad rows for the GEMM used by the all-position test path."""
        if not hasattr(self, "_lmh"):
            V = self.vocab
            Vp = (V + 7) // 8 * 8
            w = torch.zeros((Vp, self.dim), dtype=torch.bfloat16, device=self.device)
            w[:V] = self.sd["lm_head.weight"].to(self.device, torch.bfloat16)
            b = None
            if "lm_head.bias" in self.sd:
                b = torch.zeros((Vp,), dtype=torch.bfloat16, device=self.device)
                b[:V] = self.sd["lm_head.bias"].to(self.device, torch.bfloat16)
            self._lmh = (w, b)
        return self._lmh

    EOS_CHECK_EVERY = 64      # decode steps between host-side EOS checks (one 8-byte D2H read per chunk)

    def _cap(self, S, max_new_tokens):
        """max_new_tokens is a ceiling (HF stops at EOS long before the reference's default 2048, inference.py:48): token 0
        comes from the prefill, step t consumes cache slot S+t-1 < max_ctx."""
        n = min(int(max_new_tokens), self.max_ctx - S + 1)
        if n < 1:
            raise ValueError("no room to generate: prompt of %d tokens fills max_ctx=%d" % (S, self.max_ctx))
        return n

    def _sample(self, inputs_embeds, attention_mask, eos_token_id, pad_token_id, max_new_tokens, temperature, top_k, top_p,
                generator, return_logits):
        """HF sampling loop (GenerationMixin._sample, transformers==4.40.1; call site llava_next_video.py:655-661 with
        inference.py's do_sample=True, temperature=0.2): per step logits -> warpers -> softmax -> multinomial. The logits
        come from the library one step at a time; the pick happens on the device with torch (no host synchronisation), is
        handed back with gvl_lm_set_next_token, and finished rows emit pad_token_id like HF's unfinished_sequences mask.
        Every EOS_CHECK_EVERY steps the `unfinished` flag is read back and the row stops early."""
        lib = _lib.load()
        B = inputs_embeds.shape[0]
        outs, logs = [], []
        for b in range(B):
            emb = inputs_embeds[b]
            if attention_mask is not None:
                emb = emb[attention_mask[b].to(emb.device).bool()]
            logits, _ = self.prefill(emb)
            lm, S = self._active
            n_cap = self._cap(S, max_new_tokens)
            toks = torch.full((n_cap,), int(pad_token_id), dtype=torch.int64, device=self.device)
            lg = torch.zeros((n_cap, self.vocab), dtype=torch.float32, device=self.device) if return_logits else None
            step_logits = torch.empty((1, self.vocab), dtype=torch.float32, device=self.device)
            scratch = torch.empty((1,), dtype=torch.int64, device=self.device)
            unfinished = torch.ones((), dtype=torch.int64, device=self.device)
            n_done = n_cap
            for t in range(n_cap):
                if return_logits:
                    lg[t].copy_(logits.reshape(-1))
                scores = hostlogic.warp_logits(logits.reshape(1, -1), temperature, top_k, top_p)
                nxt = torch.multinomial(torch.softmax(scores, dim=-1), num_samples=1, generator=generator).reshape(())
                nxt = nxt * unfinished + int(pad_token_id) * (1 - unfinished)
                toks[t] = nxt
                if eos_token_id is not None:
                    unfinished = unfinished * (nxt != int(eos_token_id)).to(torch.int64)
                    if (t + 1) % self.EOS_CHECK_EVERY == 0 and int(unfinished) == 0:
                        n_done = t + 1
                        break
                if t + 1 < n_cap:
                    rc = lib.gvl_lm_set_next_token(lm, ctypes.c_void_p(toks[t:t + 1].data_ptr()), _stream())
                    _lib.check(rc, "gvl_lm_set_next_token")
                    rc = lib.gvl_lm_decode(lm, 1, ctypes.c_void_p(scratch.data_ptr()), ctypes.c_void_p(step_logits.data_ptr()),
                                           -1, int(pad_token_id), _stream())
                    _lib.check(rc, "gvl_lm_decode")
                    logits = step_logits
            outs.append(toks[:n_done])
            logs.append(lg[:n_done] if return_logits else None)
        return self._stack_rows(outs, logs, eos_token_id, pad_token_id, return_logits)

    def _stack_rows(self, outs, logs, eos_token_id, pad_token_id, return_logits):
        """HF generate returns [B, L]: L = length at which the LAST row finished (EOS included), earlier rows padded."""
        if eos_token_id is not None:
            lens = []
            for t in outs:
                hit = (t == int(eos_token_id)).nonzero()
                lens.append(int(hit[0]) + 1 if hit.numel() > 0 else t.shape[0])
            L = max(lens)
        else:
            L = max(t.shape[0] for t in outs)
        rows = []
        for t in outs:
            r = torch.full((L,), int(pad_token_id), dtype=torch.int64, device=self.device)
            n = min(L, t.shape[0])
            r[:n] = t[:n]
            rows.append(r)
        out = torch.stack(rows, dim=0)
        if not return_logits:
            return out
        lrows = []
        for lg in logs:
            r = torch.zeros((L, self.vocab), dtype=torch.float32, device=self.device)
            n = min(L, lg.shape[0])
            r[:n] = lg[:n]
            lrows.append(r)
        return out, torch.stack(lrows, dim=0)

    MAX_BATCH = 4             # sequences per gvl_lm_decode_batch call (GEMV kernel: up to 4 activation rows per weight pass)

    def _greedy_group(self, embs, eos, pad_token_id, max_new_tokens, return_logits):
        """Greedy decode of 2..4 unpadded sequences TOGETHER: each is prefilled on its own gvl_lm object (own KV cache), then all
        advance in lock-step through gvl_lm_decode_batch, which streams every weight matrix once per step for the whole group.
        Returns ([tokens_b], [logits_b or None])."""
        lib = _lib.load()
        nb = len(embs)
        handles, firsts, first_logits, caps = [], [], [], []
        for slot, emb in enumerate(embs):
            fl, _ = self.prefill(emb, slot=slot)
            lm, S = self._active
            handles.append(lm)
            caps.append(self._cap(S, max_new_tokens))
            firsts.append(_wrap_device_i64(ctypes.c_void_p(lib.gvl_lm_first_token(lm)), self.device))
            first_logits.append(fl)
        n_cap = min(caps)
        toks = torch.full((nb, n_cap), int(pad_token_id), dtype=torch.int64, device=self.device)
        lg = torch.zeros((nb, n_cap, self.vocab), dtype=torch.float32, device=self.device) if return_logits else None
        for b in range(nb):
            toks[b, 0:1].copy_(firsts[b])
            if return_logits:
                lg[b, 0].copy_(first_logits[b])
        arr = (ctypes.c_void_p * nb)(*[h.value for h in handles])
        done = 1
        chunk = (n_cap - 1) if eos < 0 else self.EOS_CHECK_EVERY
        finished = (toks[:, 0] == eos) if eos >= 0 else None
        # a row whose FIRST token is EOS keeps stepping with the others (its later tokens are cut off below, like HF pads them)
        while done < n_cap and not (finished is not None and bool(finished.all())):
            n = min(chunk, n_cap - done)
            tbuf = torch.empty((nb, n), dtype=torch.int64, device=self.device)
            lbuf = torch.empty((nb, n, self.vocab), dtype=torch.float32, device=self.device) if return_logits else None
            rc = lib.gvl_lm_decode_batch(arr, nb, n, ctypes.c_void_p(tbuf.data_ptr()),
                                         ctypes.c_void_p(lbuf.data_ptr() if return_logits else 0), eos, int(pad_token_id), _stream())
            _lib.check(rc, "gvl_lm_decode_batch")
            toks[:, done:done + n] = tbuf
            if return_logits:
                lg[:, done:done + n] = lbuf
            done += n
            if eos >= 0:
                finished = finished | (tbuf == eos).any(dim=1)
        outs, logs = [], []
        for b in range(nb):
            t = toks[b, :done]
            if eos >= 0:
                hit = (t == eos).nonzero()
                if hit.numel() > 0:
                    first = int(hit[0])
                    t = t.clone()
                    t[first + 1:] = int(pad_token_id)
            outs.append(t)
            logs.append(lg[b, :done] if return_logits else None)
        return outs, logs

    def generate(self, inputs_embeds=None, attention_mask=None, eos_token_id=None, pad_token_id=0, do_sample=False,
                 num_beams=1, max_new_tokens=16, temperature=None, top_p=None, top_k=50, generator=None, return_logits=False,
                 batched=True, **unused):
        """generate for a batch of left-padded sequences (each row is compacted with its attention_mask and run
        as an unpadded sequence: identical to the reference's varlen path because padding carries mask 0 and
        position ids are mask-cumsum, modeling_phi3.py:1593-1599). Returns int64 [B, L] like HF generate from inputs_embeds
        (new tokens only): L = max_new_tokens when eos_token_id is None, else the step at which the last row hit EOS.
        Greedy with more than one row: the rows are prefilled one by one on their own KV caches and then decoded TOGETHER in groups of
        up to 4 (gvl_lm_decode_batch: one pass over the weights per step for the whole group; batched=False keeps them separate).
        do_sample=False: greedy, all steps of a chunk inside the library (one launch per EOS_CHECK_EVERY steps; one launch
        in total without EOS). do_sample=True: HF sampling (temperature, top_k -- HF's GenerationConfig default 50 --, top_p,
        multinomial), one library step per token. Decoding across Phi-3.5's LongRoPE switch (position 4096) follows the
        reference's cached path: see _get()."""
        if num_beams != 1:
            raise NotImplementedError("beam search is not on the reference's inference path (num_beams=1, inference.py:172)")
        if inputs_embeds.dim() == 2:
            inputs_embeds = inputs_embeds[None]
        if do_sample:
            return self._sample(inputs_embeds, attention_mask, eos_token_id, pad_token_id, max_new_tokens, temperature, top_k,
                                top_p, generator, return_logits)
        lib = _lib.load()
        B = inputs_embeds.shape[0]
        outs, logs = [], []
        eos = -1 if eos_token_id is None else int(eos_token_id)
        rows = []
        for b in range(B):
            emb = inputs_embeds[b]
            if attention_mask is not None:
                emb = emb[attention_mask[b].to(emb.device).bool()]
            rows.append(emb)
        if B > 1 and batched:
            # several clips on this GPU: decode them in groups that share every weight pass (decode is HBM-bound on the weights)
            for g0 in range(0, B, self.MAX_BATCH):
                grp = rows[g0:g0 + self.MAX_BATCH]
                if len(grp) == 1:
                    o, l = self.generate(inputs_embeds=grp[0][None], eos_token_id=eos_token_id, pad_token_id=pad_token_id,
                                         max_new_tokens=max_new_tokens, return_logits=True, batched=False)
                    o, l = [o[0]], [l[0] if return_logits else None]
                else:
                    o, l = self._greedy_group(grp, eos, pad_token_id, max_new_tokens, return_logits)
                outs += o
                logs += l
            return self._stack_rows(outs, logs, eos_token_id, pad_token_id, return_logits)
        for b in range(B):
            emb = rows[b]
            first_logits, _ = self.prefill(emb)
            lm, S = self._active
            n_cap = self._cap(S, max_new_tokens)
            toks = torch.full((n_cap,), int(pad_token_id), dtype=torch.int64, device=self.device)
            lg = torch.zeros((n_cap, self.vocab), dtype=torch.float32, device=self.device) if return_logits else None
            # token 0 is the argmax of the prefill logits; decode steps produce tokens 1..n-1
            first = ctypes.c_void_p(lib.gvl_lm_first_token(lm))
            toks[0:1].copy_(_wrap_device_i64(first, self.device))   # first token: one device int64 owned by the lib
            if return_logits:
                lg[0].copy_(first_logits)
            done = 1
            chunk = (n_cap - 1) if eos < 0 else self.EOS_CHECK_EVERY
            if eos >= 0 and int(toks[0]) == eos:
                n_cap = 1                                             # HF: everything after EOS is pad
            while done < n_cap:
                n = min(chunk, n_cap - done)
                rc = lib.gvl_lm_decode(lm, n, ctypes.c_void_p(toks[done:].data_ptr()),
                                       ctypes.c_void_p(lg[done:].data_ptr() if return_logits else 0), eos, int(pad_token_id),
                                       _stream())
                _lib.check(rc, "gvl_lm_decode")
                done += n
                if eos >= 0 and done < n_cap and bool((toks[done - n:done] == eos).any()):
                    break
            outs.append(toks[:done])
            logs.append(lg[:done] if return_logits else None)
        return self._stack_rows(outs, logs, eos_token_id, pad_token_id, return_logits)

    def close(self):
        lib = _lib.load()
        for h in self._lms.values():
            lib.gvl_lm_destroy(h)
        self._lms = {}


def _wrap_device_i64(ptr, device):
    """View one device int64 owned by the library as a torch tensor (no copy)."""
    class _Holder:
        pass
    h = _Holder()
    h.__cuda_array_interface__ = {"shape": (1,), "typestr": "<i8", "data": (ptr.value, False), "version": 2}
    return torch.as_tensor(h, device=device)


class LLAVA_NEXT_VIDEO:
    """Drop-in for models/llava_next_video.py::LLAVA_NEXT_VIDEO on the inference path (phi3.5 and llama3 variants).

    params: dict with the reference's sub-module state_dicts
        'vision_tower'            CLIPVisionModel.state_dict()
        'video_encoder'           PretrainInternVideo2.state_dict()
        'multi_modal_projector'   {'linear_0.weight', ...} (phi3.5) or {'linear_1.weight', 'linear_2.weight', ...} (llama3)
        'video_projecter'         {'up_proj.weight', 'up_proj.bias', 'down_proj.weight', 'down_proj.bias'}
        'language_model'          Phi3ForCausalLM / LlamaForCausalLM state_dict (after reset_embeddings; LoRA merged)
        'glb_GN', 'sub_GN'        (phi3.5)   or 'image_newline' (llama3)
    """

    def __init__(self, params, llm="phi3.5", tokenizer=None, num_frames=96, num_segs=12, max_txt_len=2048,
                 lm_cfg=None, clip_cfg=None, iv2_cfg=None, max_ctx=None, max_new_tokens=2048, device="cuda"):
        """max_ctx (KV-cache rows) defaults to what the reference CLI's defaults can reach: visual tokens + max_txt_len +
        max_new_tokens (inference.py:44-48: 2048 / 2048), rounded up to 256 -- 7680 rows = 3.0 GB for Phi-3.5 at 96 frames."""
        if max_ctx is None:
            tps = (156 if llm == "phi3.5" else 64) + 16 * (num_frames // num_segs) + 1
            max_ctx = (num_segs * tps + max_txt_len + max_new_tokens + 255) // 256 * 256
        self.llm = llm
        self.device = torch.device(device)
        self.tokenizer = tokenizer
        self.num_frames, self.num_segs = num_frames, num_segs
        self.max_txt_len = max_txt_len
        clip_cfg = clip_cfg or dict(heads=16, layers=24, image=336)
        iv2_cfg = iv2_cfg or dict(heads=16, depth=40)
        self.vision_tower = CLIPVisionModel(params["vision_tower"], clip_cfg["heads"], clip_cfg["layers"],
                                            clip_cfg.get("image", 336), device)
        self.video_encoder = PretrainInternVideo2(params["video_encoder"], iv2_cfg["heads"], iv2_cfg["depth"],
                                                  num_frames // num_segs, device)
        mm = params["multi_modal_projector"]
        if llm == "phi3.5":
            self.multi_modal_projector = MLP2Projector(mm["linear_0.weight"], mm["linear_0.bias"], mm["linear_1.weight"],
                                                       mm["linear_1.bias"], device)
            self.sub_GN = params["sub_GN"].to(self.device, torch.float32).reshape(-1).contiguous()
            self.glb_GN = params["glb_GN"].to(self.device, torch.float32).reshape(1, -1).contiguous()
        else:
            self.multi_modal_projector = MLP2Projector(mm["linear_1.weight"], mm["linear_1.bias"], mm["linear_2.weight"],
                                                       mm["linear_2.bias"], device)
            self.image_newline = params["image_newline"].to(self.device, torch.bfloat16).reshape(-1).contiguous()
        vp = params["video_projecter"]
        self.video_projecter = MLP2Projector(vp["up_proj.weight"], vp["up_proj.bias"], vp["down_proj.weight"],
                                             vp["down_proj.bias"], device)
        c = lm_cfg
        self.language_model = CausalLM(params["language_model"], c["arch"], c["heads"], c["kv_heads"], c["head_dim"],
                                       c["eps"], c["rope"], max_ctx=max_ctx, device=device)
        self._newline_tok = None

    # ---------------------------------------------------------------- encode_images (llava_next_video.py:491-566)
    def _encode_units(self, spatial_units, temporal_units):
        """spatial_units [U,3,336,336], temporal_units [U,fps,3,224,224] -> [U, tokens_per_seg, D] bf16."""
        U = spatial_units.shape[0]
        fps = temporal_units.shape[1]
        lib = _lib.load()
        hs = self.vision_tower(spatial_units, output_hidden_states=True).hidden_states[-2]       # fp32 [U,577,1024]
        if self.llm == "phi3.5":
            feat = ops.hd_merge_newline(hs, self.sub_GN)                                         # [U,156,4096] bf16
        else:
            feat = ops.clip_pool3(hs)                                                             # [U,64,1024] bf16
        sp = self.multi_modal_projector(feat)                                                     # [U,156|64,D]
        tpix = temporal_units.permute(0, 2, 1, 3, 4).contiguous()                                 # (b s) c t h w
        xv = self.video_encoder(tpix, None, False, x_vis_return_idx=-2, x_vis_only=True)          # [U,1+fps*256,1408]
        pooled = ops.iv2_pool(xv, fps)                                                            # [U,fps*16,1408]
        tm = self.video_projecter(pooled)                                                         # [U,fps*16,D]
        if self._newline_tok is None:
            if self.llm == "phi3.5":
                self._newline_tok = self.multi_modal_projector(self.glb_GN).reshape(-1).contiguous()   # :560-561
            else:
                self._newline_tok = self.image_newline
        D = sp.shape[-1]
        out = torch.empty((U, sp.shape[1] + tm.shape[1] + 1, D), dtype=torch.bfloat16, device=self.device)
        rc = lib.gvl_visual_concat(ctypes.c_void_p(sp.data_ptr()), sp.shape[1], ctypes.c_void_p(tm.data_ptr()),
                                   tm.shape[1], ctypes.c_void_p(self._newline_tok.data_ptr()),
                                   ctypes.c_void_p(out.data_ptr()), U, D, _stream())
        _lib.check(rc, "gvl_visual_concat")
        return out

    def _encode_local_units(self, samples, unit_chunk=48):
        """This rank's block-partition share of the (clip, segment) units -> [units_local, tokens_per_seg, D]."""
        spatial = samples["spatial_pixel_values"]
        temporal = samples["temporal_pixel_values"]
        B, segs = spatial.shape[:2]
        frames = temporal.shape[1]
        if frames % segs != 0:
            raise ValueError("num_frames must be divisible by num_segs")
        fps = frames // segs
        n_units = B * segs
        sp_u = spatial.reshape(n_units, *spatial.shape[2:])
        tp_u = temporal.reshape(n_units, fps, *temporal.shape[2:])
        rank, ws = gdist.world()
        start, cnt = gdist.partition_units(n_units, ws)[rank]
        blocks = []
        for s0 in range(start, start + cnt, unit_chunk):
            s1 = min(s0 + unit_chunk, start + cnt)
            blocks.append(self._encode_units(sp_u[s0:s1].to(self.device, non_blocking=True),
                                             tp_u[s0:s1].to(self.device, non_blocking=True)))
        if blocks:
            local = torch.cat(blocks, dim=0) if len(blocks) > 1 else blocks[0]
        else:
            tps = (156 if self.llm == "phi3.5" else 64) + 16 * fps + 1
            local = torch.empty((0, tps, self.language_model.dim), dtype=torch.bfloat16, device=self.device)
        return local, B, segs

    def encode_images(self, samples, unit_chunk=48):
        """Reference semantics (llava_next_video.py:491-566): every caller gets the full [B, segs*tokens_per_seg, D]."""
        local, B, segs = self._encode_local_units(samples, unit_chunk)
        full = gdist.allgather_units(local, B * segs)               # the ONE collective on the path
        return full.reshape(B, segs * full.shape[1], full.shape[2])

    def encode_images_for_decode(self, samples, unit_chunk=48):
        """What `generate` needs: only the clips THIS rank decodes (clip b -> rank b % world). Same single collective, as an
        all-to-all with per-peer splits (gvl.dist.exchange_units). Returns (feats [n_mine, ...], clip indices)."""
        local, B, segs = self._encode_local_units(samples, unit_chunk)
        return gdist.exchange_units(local, B, segs)

    # ---------------------------------------------------------------- prepare_multimodal_inputs (:568-596)
    def get_input_embeddings(self):
        return self.language_model.get_input_embeddings()

    def prepare_multimodal_inputs(self, batch_input_ids, batch_labels, batch_attention_mask, batch_image_features,
                                  batch_image_ids):
        table = self.language_model.embed_table
        embeds, masks = [], []
        for feats, ids, mask, image_ids in zip(batch_image_features, batch_input_ids, batch_attention_mask, batch_image_ids):
            where = (ids == hostlogic.IMAGE_TOKEN_INDEX).nonzero()
            if where.numel() != 1:
                raise ValueError("each prompt must contain exactly one <image> sentinel")
            pos = int(where[0])
            vis_last = image_ids == "text"
            # padding rows (mask 0) hold pad ids: valid table rows, gathered like any other id
            ids_dev = ids.to(self.device)
            embeds.append(ops.embed_splice(ids_dev, pos, table, feats.contiguous(), vis_last=vis_last))
            m = mask.to(self.device)
            ones = torch.ones(feats.shape[0], dtype=m.dtype, device=self.device)
            if vis_last:
                masks.append(torch.cat([m[:pos], m[pos + 1:], torch.zeros_like(ones)]))
            else:
                masks.append(torch.cat([m[:pos], ones, m[pos + 1:]]))
        return torch.stack(embeds, 0), None, torch.stack(masks, 0)

    def tokenizer_image_token(self, prompt, tokenizer, image_token_index=hostlogic.IMAGE_TOKEN_INDEX, return_tensors=None):
        return hostlogic.tokenizer_image_token(prompt, tokenizer, image_token_index, return_tensors, device="cpu")

    # ---------------------------------------------------------------- generate (:616-666)
    @torch.inference_mode()
    def generate(self, samples, **generate_kwargs):
        if "input_ids" in samples:                       # tokenizer-free entry used by bench / tests
            id_lists = [list(map(int, x)) for x in samples["input_ids"]]
            pad_id = int(samples.get("pad_token_id", 0))
            eos_id = samples.get("eos_token_id")
        else:
            id_lists = [self.tokenizer_image_token(t, self.tokenizer) for t in samples["prompts"]]
            pad_id, eos_id = self.tokenizer.pad_token_id, self.tokenizer.eos_token_id
        ids, mask = hostlogic.left_pad(id_lists, pad_id, self.max_txt_len)
        feats, mine = self.encode_images_for_decode(samples)
        B = len(id_lists)
        video_ids = samples.get("video_ids", ["video"] * B)
        gk = dict(generate_kwargs)
        gk.setdefault("do_sample", False)
        local = {}
        if mine:
            sel = torch.tensor(mine, dtype=torch.long)
            embeds, _, masks = self.prepare_multimodal_inputs(ids[sel], None, mask[sel], feats, [video_ids[b] for b in mine])
            # ONE call for the clips this rank owns: greedy rows are decoded together (CausalLM.generate, gvl_lm_decode_batch)
            out = self.language_model.generate(inputs_embeds=embeds, attention_mask=masks, eos_token_id=eos_id, pad_token_id=pad_id, **gk)
            for i, b in enumerate(mine):
                row = out[i]
                if eos_id is not None:
                    hit = (row == int(eos_id)).nonzero()
                    if hit.numel() > 0:
                        row = row[: int(hit[0]) + 1]
                local[b] = row
        rank, ws = gdist.world()
        if ws > 1:
            width = int(gk.get("max_new_tokens", 16))
            gathered = gdist.gather_tokens(local, B, width, pad_id if pad_id is not None else 0, self.device)
            local = dict(enumerate(gathered))
        toks = [local[b] for b in range(B)]
        if self.tokenizer is not None and "input_ids" not in samples:
            text = self.tokenizer.batch_decode([t.cpu().tolist() for t in toks], skip_special_tokens=True)
            return [t.strip() for t in text]
        return toks
