"""ORACLE - test infrastructure, not product code.

fp32 torch restatement of ONE KV-cached decoder step of HF `T5ForConditionalGeneration` (transformers
models/t5/modeling_t5.py: T5Stack decoder -> T5Block -> T5LayerSelfAttention / T5LayerCrossAttention / T5LayerFF), the
model call inside the reference's report generation (`FusionTransformerModel.generate`, training_pipeline.py:613-618).
It has the interface of the CUDA backend (`mmdx_b200.t5_fast.MmdxStep`) and is its checker: tests/test_t5_cpu.py drives
HF's own beam search with it and asserts token identity with stock `generate`; tests/test_t5_gpu.py compares the CUDA
step's logits with it.  Only tests may import this file.
"""
from __future__ import annotations

import torch

from mmdx_b200.t5_fast import relative_position_bias_table


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
        self.tied = bool(getattr(cfg, "scale_decoder_outputs", cfg.tie_word_embeddings))     # = "scale the decoder output"
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

    def bias_table(self, n):
        return relative_position_bias_table(self.cfg, self.rel, n, self.dev)

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

    def score_topk(self, logits, beam_scores, banned, ban_eos, eos_id, num_beams, k):
        """torch restatement of mmdx_t5_score_topk: log_softmax, masks, + beam score, top-k per study."""
        R, V = logits.shape
        lp = torch.log_softmax(logits.float(), -1)
        if ban_eos:
            lp[:, eos_id] = -float("inf")
        if banned is not None:
            for r in range(R):
                b = banned[r][banned[r] >= 0].long()
                lp[r, b] = -float("inf")
        lp = (lp + beam_scores.to(lp.device).view(R, 1)).view(R // num_beams, num_beams * V)
        s, i = torch.topk(lp, k)
        return s.cpu(), i.cpu()
