"""KV-cached T5 decoder step behind the bundle's own `generate`.

The reference produces `report_text` with `fusion_model.generate(...)` -> HF `T5ForConditionalGeneration.generate`
(beam search, 150-180 new tokens; training_pipeline.py:613-618, inference_pipeline.py:190-196): ~98 % of the request
latency, spent in hundreds of tiny eager kernels per token.  Here the SEARCH stays HF's own code - beam bookkeeping,
length penalty, early stopping, min-length and no-repeat-n-gram processors are untouched, so the tokens are HF's by
construction - and only the model call inside the loop is replaced: `FastT5Generator` swaps the report model's
`forward` for one decoder step over a private KV cache.

Two step backends with the same interface:
  * `TorchStep`  - plain fp32 torch restatement of the HF T5 decoder step (CPU or GPU); the checker of the CUDA backend
                   and what the CPU tests run (tests/test_t5_cpu.py: token-identical to stock HF generate);
  * `MmdxStep`   - hand-written CUDA kernels through the C ABI (csrc/t5_decoder.cu, `mmdx_t5_*`).
"""
from __future__ import annotations

import math

import torch


class TorchStep:
    """One T5 decoder step for R = batch * beams rows, fp32, KV-cached.  Mirrors HF modeling_t5 (T5Stack decoder):
    x = E[tok]; per block: x += SelfAttn(RMSNorm(x)) ; x += CrossAttn(RMSNorm(x), enc) ; x += Wo relu(Wi RMSNorm(x));
    logits = (RMSNorm(x) * d_model^-0.5) E^T (tied embeddings).  T5 attention has no 1/sqrt(d) scaling; the
    self-attention adds the learned relative-position bias of block 0 in every block; cross-attention adds none."""

    def __init__(self, model, device=None):
        cfg = model.config
        self.cfg = cfg
        dev = torch.device(device) if device is not None else next(model.parameters()).device
        self.dev = dev
        sd = {k: v.detach().to(dev, torch.float32) for k, v in model.state_dict().items()}
        self.E = sd["shared.weight"]
        self.lm = sd.get("lm_head.weight", self.E)
        self.tied = bool(cfg.tie_word_embeddings)
        self.blocks = []
        for i in range(cfg.num_decoder_layers):
            p = f"decoder.block.{i}.layer."
            self.blocks.append({
                "ln0": sd[p + "0.layer_norm.weight"],
                "sq": sd[p + "0.SelfAttention.q.weight"], "sk": sd[p + "0.SelfAttention.k.weight"],
                "sv": sd[p + "0.SelfAttention.v.weight"], "so": sd[p + "0.SelfAttention.o.weight"],
                "ln1": sd[p + "1.layer_norm.weight"],
                "cq": sd[p + "1.EncDecAttention.q.weight"], "ck": sd[p + "1.EncDecAttention.k.weight"],
                "cv": sd[p + "1.EncDecAttention.v.weight"], "co": sd[p + "1.EncDecAttention.o.weight"],
                "ln2": sd[p + "2.layer_norm.weight"],
                "wi": sd[p + "2.DenseReluDense.wi.weight"], "wo": sd[p + "2.DenseReluDense.wo.weight"],
            })
        self.final_ln = sd["decoder.final_layer_norm.weight"]
        self.rel = sd["decoder.block.0.layer.0.SelfAttention.relative_attention_bias.weight"]     # [buckets, heads]
        self.eps = cfg.layer_norm_epsilon
        self.H, self.dk = cfg.num_heads, cfg.d_kv
        if cfg.feed_forward_proj != "relu":
            raise ValueError("only the ReLU feed-forward of t5-small / T5Config() is implemented")
        self.t = 0

    # -- relative position bias for distance d = query_pos - key_pos >= 0 (decoder: bidirectional=False)
    def bias_table(self, n):
        nb, md = self.cfg.relative_attention_num_buckets, self.cfg.relative_attention_max_distance
        rp = torch.arange(n, device=self.dev)
        max_exact = nb // 2
        large = max_exact + (torch.log(rp.float() / max_exact) / math.log(md / max_exact) * (nb - max_exact)).to(torch.long)
        large = torch.min(large, torch.full_like(large, nb - 1))
        bucket = torch.where(rp < max_exact, rp, large)
        return self.rel[bucket]                                   # [n, heads]

    def _rms(self, x, w):
        var = x.pow(2).mean(-1, keepdim=True)
        return x * torch.rsqrt(var + self.eps) * w

    def begin(self, enc, rows, max_steps):
        """enc: [rows, n_enc, d_model] encoder states per row (already expanded over beams)."""
        R, H, dk = rows, self.H, self.dk
        self.t = 0
        self.bias = self.bias_table(max_steps + 1)
        self.sk = [torch.zeros(R, H, max_steps + 1, dk, device=self.dev) for _ in self.blocks]
        self.sv = [torch.zeros(R, H, max_steps + 1, dk, device=self.dev) for _ in self.blocks]
        enc = enc.to(self.dev, torch.float32)
        self.ck = [(enc @ b["ck"].t()).view(R, -1, H, dk).transpose(1, 2) for b in self.blocks]
        self.cv = [(enc @ b["cv"].t()).view(R, -1, H, dk).transpose(1, 2) for b in self.blocks]

    def reorder(self, beam_idx):
        idx = beam_idx.to(self.dev, torch.long)
        self.sk = [k.index_select(0, idx) for k in self.sk]
        self.sv = [v.index_select(0, idx) for v in self.sv]
        self.ck = [k.index_select(0, idx) for k in self.ck]
        self.cv = [v.index_select(0, idx) for v in self.cv]

    def step(self, tokens):
        """tokens: [R] int64 -> logits [R, vocab] fp32; appends this position to the cache."""
        R, H, dk, t = tokens.shape[0], self.H, self.dk, self.t
        x = self.E[tokens.to(self.dev)]
        for i, b in enumerate(self.blocks):
            h = self._rms(x, b["ln0"])
            q = (h @ b["sq"].t()).view(R, H, 1, dk)
            self.sk[i][:, :, t] = (h @ b["sk"].t()).view(R, H, dk)
            self.sv[i][:, :, t] = (h @ b["sv"].t()).view(R, H, dk)
            s = q @ self.sk[i][:, :, :t + 1].transpose(-1, -2)                    # [R,H,1,t+1]
            s = s + self.bias[torch.arange(t, -1, -1, device=self.dev)].t().view(1, H, 1, t + 1)
            a = torch.softmax(s.float(), -1) @ self.sv[i][:, :, :t + 1]
            x = x + a.transpose(1, 2).reshape(R, H * dk) @ b["so"].t()
            h = self._rms(x, b["ln1"])
            q = (h @ b["cq"].t()).view(R, H, 1, dk)
            a = torch.softmax((q @ self.ck[i].transpose(-1, -2)).float(), -1) @ self.cv[i]
            x = x + a.transpose(1, 2).reshape(R, H * dk) @ b["co"].t()
            h = self._rms(x, b["ln2"])
            x = x + torch.relu(h @ b["wi"].t()) @ b["wo"].t()
        x = self._rms(x, self.final_ln)
        if self.tied:
            x = x * (self.cfg.d_model ** -0.5)
        self.t += 1
        return x @ self.lm.t()


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
        self._bias_src = TorchStep.bias_table                     # HF's bucket arithmetic, evaluated once per generation
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
        class _B:      # bias_table() only needs cfg / rel / dev
            pass
        b = _B(); b.cfg, b.rel, b.dev = self.cfg, self._rel, torch.device("cpu")
        bias = TorchStep.bias_table(b, max_steps).contiguous()            # [max_steps, heads] fp32 on the host
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

    def __init__(self, model, backend=None):
        self.model = model
        self.backend = backend if backend is not None else TorchStep(model)

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
