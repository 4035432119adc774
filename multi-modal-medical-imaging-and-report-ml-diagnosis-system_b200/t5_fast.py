"""KV-cached T5 decoder step behind the bundle's own `generate`.

The reference produces `report_text` with `fusion_model.generate(...)` -> HF `T5ForConditionalGeneration.generate`
(beam search, 150-180 new tokens; training_pipeline.py:613-618, inference_pipeline.py:190-196): ~98 % of the request
latency, spent in hundreds of tiny eager kernels per token.  Here the SEARCH stays HF's own code - beam bookkeeping,
length penalty, early stopping, min-length and no-repeat-n-gram processors are untouched, so the tokens are HF's by
construction - and only the model call inside the loop is replaced: `FastT5Generator` swaps the report model's
`forward` for one decoder step over a private KV cache.

The step backend of the product is `MmdxStep`: hand-written CUDA kernels through the C ABI (csrc/t5_decoder.cu,
`mmdx_t5_*`); there is no CPU step here.  Its checker - a plain fp32 torch restatement of the HF T5 decoder step with the
same interface, which is also what the CPU tests drive `FastT5Generator` with - lives in oracle/t5_step_ref.py (test
infrastructure).
"""
from __future__ import annotations

import math
import threading

import torch

_PATCH_LOCK = threading.RLock()      # one generation at a time swaps a model's forward (and owns the step backend's cache)


def lm_head_setup(model):
    """(weights tied?, decoder output scaled by d_model^-0.5?) of a T5ForConditionalGeneration.  transformers 4.x ties and
    scales together (`tie_word_embeddings`); 5.x always ties and keeps the scaling as `scale_decoder_outputs`."""
    cfg = model.config
    sd = model.state_dict()
    tied = "lm_head.weight" not in sd or sd["lm_head.weight"].data_ptr() == sd["shared.weight"].data_ptr() \
        or bool(torch.equal(sd["lm_head.weight"], sd["shared.weight"]) and cfg.tie_word_embeddings)
    scaled = bool(getattr(cfg, "scale_decoder_outputs", cfg.tie_word_embeddings))
    return tied, scaled


def relative_position_bias_table(cfg, rel_weight, n, device="cpu"):
    """Relative-position bias of the decoder self-attention by distance d = query_pos - key_pos in [0, n): rows of
    `rel_weight` [buckets, heads] picked with HF's bucket arithmetic (T5Attention._relative_position_bucket with
    bidirectional=False), evaluated with the same torch float32 ops so that bucket boundaries fall where HF puts them."""
    nb, md = cfg.relative_attention_num_buckets, cfg.relative_attention_max_distance
    rp = torch.arange(n, device=device)
    max_exact = nb // 2
    large = max_exact + (torch.log(rp.float() / max_exact) / math.log(md / max_exact) * (nb - max_exact)).to(torch.long)
    large = torch.min(large, torch.full_like(large, nb - 1))
    bucket = torch.where(rp < max_exact, rp, large)
    return rel_weight.to(device)[bucket]                              # [n, heads]


class MmdxStep:
    """The same step on the hand-written CUDA kernels of libmmdx.so (csrc/t5_decoder.cu, `mmdx_t5_*`)."""

    def __init__(self, model, device=None):
        import ctypes as C
        from ._lib import MmdxError, lib
        self._C, self._lib, self._err = C, lib(), MmdxError
        cfg = model.config
        if cfg.feed_forward_proj != "relu":
            raise ValueError("only the ReLU feed-forward of t5-small / T5Config() is implemented")
        dev = torch.device(device) if device is not None else next(model.parameters()).device
        if dev.type != "cuda":
            raise MmdxError("MmdxStep needs a CUDA device (there is no CPU fallback; TorchStep is the CPU checker)")
        self.dev = dev
        self.cfg = cfg
        self.blocks = range(cfg.num_decoder_layers)
        self._h = C.c_void_p()
        tied, scaled = lm_head_setup(model)
        self._check(self._lib.mmdx_t5_create(dev.index or 0, cfg.d_model, cfg.num_heads, cfg.d_kv, cfg.d_ff, cfg.num_decoder_layers,
                                             cfg.vocab_size, float(cfg.layer_norm_epsilon), (1 if scaled else 2) if tied else 0,
                                             C.byref(self._h)))
        if not tied and scaled:
            raise ValueError("untied LM head with scaled decoder output is not a T5 configuration")
        for k, v in model.state_dict().items():
            if k == "shared.weight" or k.startswith("decoder.") or (k == "lm_head.weight" and not tied):
                if "relative_attention_bias" in k:
                    continue
                t = v.detach().to("cpu", torch.float32).contiguous()
                self._check(self._lib.mmdx_t5_load_tensor(self._h, k.encode(), C.c_void_p(t.data_ptr()), t.numel()))
        self._check(self._lib.mmdx_t5_finalize(self._h))
        self._rel = model.state_dict()["decoder.block.0.layer.0.SelfAttention.relative_attention_bias.weight"].detach().float().cpu()
        self.t = 0

    def _check(self, rc):
        if rc != 0:
            raise self._err(self._lib.mmdx_t5_last_error().decode())

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self._lib.mmdx_t5_destroy(self._h)
            self._h = self._C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:      # noqa: BLE001
            pass

    @property
    def launch_count(self):
        return int(self._lib.mmdx_t5_launch_count(self._h))

    def _stream(self):
        return self._C.c_void_p(torch.cuda.current_stream(self.dev).cuda_stream)

    def begin(self, enc, rows, max_steps):
        bias = relative_position_bias_table(self.cfg, self._rel, max_steps).contiguous()      # [max_steps, heads] fp32, host
        enc = enc.to(self.dev, torch.float32).contiguous()
        self._vocab = self.cfg.vocab_size
        self._rows = rows
        with torch.cuda.device(self.dev):
            self._check(self._lib.mmdx_t5_begin(self._h, self._C.c_void_p(enc.data_ptr()), rows, enc.shape[1], int(max_steps),
                                                self._C.c_void_p(bias.data_ptr()), self._stream()))
        self._keep = [enc, bias]       # stream-ordered use: keep the operands alive instead of synchronising
        self.t = 0

    def reorder(self, beam_idx):
        idx = beam_idx.to(self.dev, torch.int32).contiguous()
        with torch.cuda.device(self.dev):
            self._check(self._lib.mmdx_t5_reorder(self._h, self._C.c_void_p(idx.data_ptr()), self._stream()))
        self._keep_idx = idx

    def step(self, tokens):
        tok = tokens.to(self.dev, torch.int32).contiguous()
        logits = torch.empty(self._rows, self._vocab, dtype=torch.float32, device=self.dev)
        with torch.cuda.device(self.dev):
            self._check(self._lib.mmdx_t5_step(self._h, self._C.c_void_p(tok.data_ptr()), self._C.c_void_p(logits.data_ptr()),
                                               self._stream()))
        self._keep_tok = tok
        self.t += 1
        return logits


    def score_topk(self, logits, beam_scores, banned, ban_eos, eos_id, num_beams, k):
        """Per study the k best continuations of log_softmax(logits) + beam_scores with the masks applied
        (`mmdx_t5_score_topk`): -> (scores fp32 [B, k], flat index int64 [B, k] = row_in_study * vocab + token), on the host."""
        C = self._C
        B = self._rows // num_beams
        bs = beam_scores.to(self.dev, torch.float32).contiguous()
        max_ban = 0 if banned is None else int(banned.shape[1])
        bn = None if not max_ban else banned.to(self.dev, torch.int32).contiguous()
        out_s = torch.empty(B, k, dtype=torch.float32, device=self.dev)
        out_i = torch.empty(B, k, dtype=torch.int32, device=self.dev)
        with torch.cuda.device(self.dev):
            self._check(self._lib.mmdx_t5_score_topk(self._h, C.c_void_p(logits.data_ptr()), C.c_void_p(bs.data_ptr()),
                                                     C.c_void_p(bn.data_ptr() if bn is not None else 0), max_ban, 1 if ban_eos else 0,
                                                     int(eos_id), int(num_beams), int(k), C.c_void_p(out_s.data_ptr()),
                                                     C.c_void_p(out_i.data_ptr()), self._stream()))
        return out_s.cpu(), out_i.cpu().to(torch.int64)


    def generate_native(self, cond, max_new_tokens, min_new_tokens, num_beams, no_repeat_ngram_size, length_penalty,
                        early_stopping, eos_token_id, pad_token_id, decoder_start_token_id):
        """The whole search in one C call (`mmdx_t5_generate`: the same algorithm as NativeBeamSearch, bookkeeping in C++)."""
        C = self._C
        B = cond.shape[0]
        cond = cond.to(self.dev, torch.float32).contiguous()
        bias = relative_position_bias_table(self.cfg, self._rel, int(max_new_tokens) + 1).contiguous()
        out = torch.empty(B, int(max_new_tokens) + 1, dtype=torch.int32)
        n = C.c_int32(0)
        es = 1 if early_stopping is True else (2 if early_stopping == "never" else 0)
        start = self.cfg.decoder_start_token_id if decoder_start_token_id is None else decoder_start_token_id
        with torch.cuda.device(self.dev):
            self._check(self._lib.mmdx_t5_generate(self._h, C.c_void_p(cond.data_ptr()), B, cond.shape[1], int(num_beams),
                                                   int(max_new_tokens), int(min_new_tokens), int(no_repeat_ngram_size),
                                                   float(length_penalty), es, int(eos_token_id), int(pad_token_id), int(start),
                                                   C.c_void_p(bias.data_ptr()), C.c_void_p(out.data_ptr()), C.byref(n),
                                                   self._stream()))
        return out[:, :n.value].to(torch.int64)


class NativeBeamSearch:
    """Beam search driven from here instead of from HF's Python loop: per token one decoder step, one scoring / top-k
    launch pair on the device and ~30 tiny host tensor ops, instead of the ~2.7 ms of eager-mode bookkeeping HF spends per
    token.  The algorithm is HF's `GenerationMixin._beam_search` (transformers 5.x) for `do_sample=False`,
    `num_return_sequences=1`, one EOS id, with the two logits processors the reference's settings enable
    (inference_pipeline.py:190: `min_new_tokens`, `no_repeat_ngram_size`) and its stopping rule (`early_stopping`,
    `length_penalty`, `max_new_tokens`): top-2K continuations per study, unfinished ones continue, finished ones compete
    on score / length ** length_penalty.  tests/test_t5_cpu.py and test_t5_gpu.py hold it to token identity with HF's own
    `generate`.  Opt-in (`model_bundle["fast_report"] = "native"`): HF's loop over the CUDA step stays the default because
    it is HF's search by construction, while this one agrees with it up to fp32 rounding of near-tied scores."""

    def __init__(self, backend, cfg, host_loop="auto"):
        """host_loop: "python" keeps the bookkeeping in this file (what the CPU tests pin against HF); "auto" hands the whole
        search to the backend's C++ loop when it has one (`MmdxStep.generate_native` -> `mmdx_t5_generate`, num_beams <= 4)."""
        self.be, self.cfg, self.host_loop = backend, cfg, host_loop

    @staticmethod
    def _banned(seqs, cur_len, n):
        """no_repeat_ngram_size = n: for every row the tokens that would repeat an n-gram already in seqs[row, :cur_len]
        (HF NoRepeatNGramLogitsProcessor), as an int32 array padded with -1."""
        import numpy as np
        R = seqs.shape[0]
        if n <= 0 or cur_len + 1 < n:
            return None
        S = seqs[:, :cur_len]
        if n == 1:
            return S.astype(np.int32)
        m = np.ones((R, cur_len - n + 1), bool)
        for j in range(n - 1):                                      # windows whose first n-1 tokens equal the current suffix
            m &= S[:, j:cur_len - n + 1 + j] == S[:, cur_len - n + 1 + j:cur_len - n + 2 + j]
        nxt = S[:, n - 1:]
        width = max(int(m.sum(1).max()), 1)
        out = np.full((R, width), -1, np.int32)
        for r in range(R):
            t = nxt[r][m[r]]
            out[r, :len(t)] = t
        return out

    _KNOWN = ("max_new_tokens", "min_new_tokens", "num_beams", "no_repeat_ngram_size", "length_penalty", "early_stopping",
              "eos_token_id", "pad_token_id", "decoder_start_token_id", "use_cache")

    @classmethod
    def supports(cls, gen_kwargs):
        """True when these `generate` arguments are the ones this search restates (deterministic beam / greedy search,
        one EOS id, at most 4 beams for the C++ host loop's top-8 kernel)."""
        extra = {k: v for k, v in gen_kwargs.items() if k not in cls._KNOWN and v not in (None, False)}
        eos = gen_kwargs.get("eos_token_id", 1)
        return (not extra and bool(gen_kwargs.get("max_new_tokens")) and isinstance(eos, int)
                and 1 <= int(gen_kwargs.get("num_beams", 1)) <= 4)

    @torch.no_grad()
    def generate(self, cond, max_new_tokens, min_new_tokens=0, num_beams=1, no_repeat_ngram_size=0, length_penalty=1.0,
                 early_stopping=False, eos_token_id=1, pad_token_id=0, decoder_start_token_id=None, **unused):
        import numpy as np
        unsupported = {k: v for k, v in unused.items() if v not in (None, False) and k not in ("use_cache",)}
        if unsupported:
            raise ValueError(f"NativeBeamSearch does not implement {sorted(unsupported)}")
        with _PATCH_LOCK:                 # the step backend holds ONE generation's KV cache
            if self.host_loop == "auto" and hasattr(self.be, "generate_native") and int(num_beams) <= 4:
                return self.be.generate_native(cond, max_new_tokens, min_new_tokens, num_beams, no_repeat_ngram_size,
                                               length_penalty, early_stopping, eos_token_id, pad_token_id,
                                               decoder_start_token_id)
            return self._generate(cond, max_new_tokens, min_new_tokens, num_beams, no_repeat_ngram_size, length_penalty,
                                  early_stopping, eos_token_id, pad_token_id, decoder_start_token_id)

    def _generate(self, cond, max_new_tokens, min_new_tokens, num_beams, no_repeat_ngram_size, length_penalty, early_stopping,
                  eos_token_id, pad_token_id, decoder_start_token_id):
        import numpy as np
        be, K = self.be, int(num_beams)
        B = cond.shape[0]
        V = self.cfg.vocab_size
        start = self.cfg.decoder_start_token_id if decoder_start_token_id is None else decoder_start_token_id
        eos, lp = int(eos_token_id), float(length_penalty)
        prompt, max_length = 1, 1 + int(max_new_tokens)
        K2 = 2 * K
        fill = pad_token_id if pad_token_id else eos               # HF: `pad_token_id or eos_token_id[0]`
        NEG = np.float32(-1.0e9)
        be.begin(cond.repeat_interleave(K, 0), B * K, max_length)
        run_seq = np.full((B, K, max_length), fill, np.int64)
        run_seq[:, :, 0] = start
        seqs = run_seq.copy()
        run_len = np.zeros((B, K), np.int64)                        # generated tokens of every stored hypothesis
        fin_len = np.zeros((B, K), np.int64)
        run_scores = np.zeros((B, K), np.float32)
        run_scores[:, 1:] = NEG
        beam_scores = np.full((B, K), NEG, np.float32)
        finished = np.zeros((B, K), bool)
        unsat = np.ones((B, 1), bool)
        top_mask = np.arange(K2) < K
        cur_len = 1
        bidx = np.arange(B)[:, None]
        while True:
            tokens = torch.from_numpy(run_seq[:, :, cur_len - 1].reshape(-1))
            logits = be.step(tokens)
            flat = run_seq.reshape(B * K, max_length)
            banned = self._banned(flat, cur_len, int(no_repeat_ngram_size))
            ban_eos = (cur_len - prompt) < int(min_new_tokens)
            tk_scores, tk_idx = be.score_topk(logits, torch.from_numpy(run_scores.reshape(-1)),
                                              None if banned is None else torch.from_numpy(banned), ban_eos, eos, K, K2)
            tk_scores = tk_scores.numpy().astype(np.float32)
            tk_idx = tk_idx.numpy()
            src_beam, tok = tk_idx // V, tk_idx % V                 # [B, K2]
            tk_seq = run_seq[bidx, src_beam]                        # [B, K2, max_length]
            tk_seq[:, :, cur_len] = tok
            tk_len = np.full((B, K2), cur_len + 1 - prompt, np.int64)
            hits = (tok == eos) | (cur_len + 1 >= max_length)
            # running beams of the next step: the best K continuations that did not stop
            run_pool = tk_scores + hits.astype(np.float32) * NEG
            nxt = torch.topk(torch.from_numpy(run_pool), K).indices.numpy()        # the selection HF makes, ties included
            run_seq = tk_seq[bidx, nxt]
            run_scores = run_pool[bidx, nxt]
            run_len = tk_len[bidx, nxt]
            beam_src = (src_beam[bidx, nxt] + np.arange(B)[:, None] * K).reshape(-1)
            # finished hypotheses: only the top K of the 2K continuations may finish
            just = hits & top_mask[None, :]
            fs = tk_scores / np.float32((cur_len + 1 - prompt) ** lp)
            full = finished.all(axis=1, keepdims=True) & (early_stopping is True)
            fs = fs + full.astype(np.float32) * NEG
            fs = fs + (~unsat).astype(np.float32) * NEG
            fs = fs + (~just).astype(np.float32) * NEG
            m_scores = np.concatenate([beam_scores, fs], 1)
            m_seq = np.concatenate([seqs, tk_seq], 1)
            m_fin = np.concatenate([finished, just], 1)
            m_len = np.concatenate([fin_len, tk_len], 1)
            top = torch.topk(torch.from_numpy(m_scores), K).indices.numpy()
            seqs, beam_scores, finished, fin_len = m_seq[bidx, top], m_scores[bidx, top], m_fin[bidx, top], m_len[bidx, top]
            be.reorder(torch.from_numpy(beam_src))
            cur_len += 1
            # can a running beam still beat the worst finished one?
            best_len = (max_length - prompt) if (early_stopping == "never" and lp > 0.0) else (cur_len - prompt)
            best_possible = run_scores[:, :1] / np.float32(best_len ** lp)
            worst = np.where(finished, beam_scores.min(axis=1, keepdims=True), NEG)
            unsat = unsat & (best_possible > worst).any(axis=1, keepdims=True)
            go_on = unsat.any() and not (finished.all() and early_stopping is True) and not hits.all()
            if not go_on:
                break
        out_len = prompt + int(fin_len[:, 0].max())
        return torch.from_numpy(seqs[:, 0, :out_len].copy())


class _StepCache:
    """What HF's generation loop needs from `past_key_values`: a length and beam reordering."""

    def __init__(self, backend):
        self.backend = backend

    def get_seq_length(self, layer_idx=0):
        return self.backend.t

    def reorder_cache(self, beam_idx):
        self.backend.reorder(beam_idx)

    # attributes / methods generation utilities probe on cache objects
    is_compileable = False

    def get_max_cache_shape(self, layer_idx=0):
        return -1

    def __len__(self):
        return len(self.backend.blocks) if hasattr(self.backend, "blocks") else 1


class FastT5Generator:
    """`generate(cond, **gen_kwargs)`: HF's generate on `model` with the model call replaced by `backend.step`."""

    def __init__(self, model, backend):
        """backend: an object with begin(enc, rows, max_steps) / step(tokens) -> logits / reorder(beam_idx) and a step
        counter `t` - `MmdxStep` in the product, oracle.t5_step_ref.TorchStep in the CPU tests."""
        self.model = model
        self.backend = backend

    @torch.no_grad()
    def generate(self, cond, **gen_kwargs):
        from transformers.modeling_outputs import BaseModelOutput, Seq2SeqLMOutput
        model, be = self.model, self.backend
        beams = int(gen_kwargs.get("num_beams", 1) or 1)
        B = cond.shape[0]
        max_new = gen_kwargs.get("max_new_tokens")
        if max_new is None:
            raise ValueError("pass max_new_tokens (the reference does: inference_pipeline.py:190)")
        dev = cond.device
        state = {"started": False}

        def fast_forward(decoder_input_ids=None, encoder_outputs=None, past_key_values=None, **kw):
            if not state["started"]:
                enc = encoder_outputs.last_hidden_state if hasattr(encoder_outputs, "last_hidden_state") else encoder_outputs[0]
                be.begin(enc, enc.shape[0], int(max_new) + 1)
                state["started"] = True
                state["cache"] = _StepCache(be)
            if decoder_input_ids.shape[1] != 1 and be.t != 0:
                raise RuntimeError("fast T5 step expects one new token per call")
            logits = be.step(decoder_input_ids[:, -1])
            return Seq2SeqLMOutput(logits=logits[:, None, :].to(dev), past_key_values=state["cache"])

        with _PATCH_LOCK:
            orig = model.forward
            model.forward = fast_forward
            try:
                return model.generate(encoder_outputs=BaseModelOutput(last_hidden_state=cond), use_cache=True, **gen_kwargs)
            finally:
                model.forward = orig
