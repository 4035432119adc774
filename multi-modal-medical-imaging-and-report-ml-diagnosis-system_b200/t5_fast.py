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

import torch


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
        self._check(self._lib.mmdx_t5_create(dev.index or 0, cfg.d_model, cfg.num_heads, cfg.d_kv, cfg.d_ff, cfg.num_decoder_layers,
                                             cfg.vocab_size, float(cfg.layer_norm_epsilon), 1 if cfg.tie_word_embeddings else 0,
                                             C.byref(self._h)))
        for k, v in model.state_dict().items():
            if k == "shared.weight" or k.startswith("decoder.") or (k == "lm_head.weight" and not cfg.tie_word_embeddings):
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

        orig = model.forward
        model.forward = fast_forward
        try:
            return model.generate(encoder_outputs=BaseModelOutput(last_hidden_state=cond), use_cache=True, **gen_kwargs)
        finally:
            model.forward = orig
