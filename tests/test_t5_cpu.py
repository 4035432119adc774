"""the report decoder.  `t5_fast.FastT5Generator` keeps HF's beam search and replaces only the model call
inside it with a KV-cached decoder step; here the step is the fp32 torch restatement (`TorchStep`, the checker of the
CUDA backend) and the bar is token identity with stock `T5ForConditionalGeneration.generate` on the same weights -
the reference's call (FusionTransformerModel.generate, training_pipeline.py:613-618) with its own generation settings
(inference_pipeline.py:190), shortened so the CPU suite stays fast."""
import numpy as np
import pytest
import torch
from transformers import T5Config, T5ForConditionalGeneration
from transformers.modeling_outputs import BaseModelOutput

from conftest import load_golden
from mmdx_b200.t5_fast import FastT5Generator
from oracle.t5_step_ref import TorchStep


@pytest.fixture(scope="module")
def t5():
    torch.manual_seed(0)
    return T5ForConditionalGeneration(T5Config(decoder_start_token_id=0)).eval()


REF_KW = dict(num_beams=4, no_repeat_ngram_size=3, length_penalty=1.1, early_stopping=True, eos_token_id=1, pad_token_id=0)


@pytest.mark.parametrize("kw", [
    dict(REF_KW, max_new_tokens=24, min_new_tokens=18),                         # the reference's settings, shortened
    dict(max_new_tokens=16, num_beams=1, eos_token_id=1, pad_token_id=0),       # greedy
    dict(max_new_tokens=12, num_beams=2, length_penalty=0.8, early_stopping=False, eos_token_id=1, pad_token_id=0),
])
def test_generate_token_identical_to_hf(t5, kw):
    torch.manual_seed(1)
    cond = torch.randn(3, 4, 512)
    with torch.no_grad():
        want = t5.generate(encoder_outputs=BaseModelOutput(last_hidden_state=cond), **kw)
    got = FastT5Generator(t5, TorchStep(t5)).generate(cond, **kw)
    assert torch.equal(want, got)
    # the model is untouched afterwards (forward restored): a second stock call gives the same tokens
    with torch.no_grad():
        assert torch.equal(t5.generate(encoder_outputs=BaseModelOutput(last_hidden_state=cond), **kw), want)


def test_generate_from_the_reference_conditioning_tokens(t5):
    """cond = the golden conditioning tokens written by the reference's own _make_encoder_outputs (tests/golden)."""
    g = load_golden("g2_B8_L128_ragged")
    cond = torch.from_numpy(g["cond"][:2])
    kw = dict(REF_KW, max_new_tokens=16, min_new_tokens=12)
    with torch.no_grad():
        want = t5.generate(encoder_outputs=BaseModelOutput(last_hidden_state=cond), **kw)
    assert torch.equal(FastT5Generator(t5, TorchStep(t5)).generate(cond, **kw), want)


def test_step_logits_match_hf_decoder(t5):
    """One level down: the step's logits against HF's own forward with its cache, including a beam reorder."""
    torch.manual_seed(2)
    R = 4
    cond = torch.randn(R, 4, 512)
    be = TorchStep(t5)
    be.begin(cond, R, 8)
    toks = torch.zeros(R, 1, dtype=torch.long)
    past = None
    seq = toks
    for t in range(5):
        with torch.no_grad():
            out = t5(decoder_input_ids=seq[:, -1:] if past is not None else seq, encoder_outputs=BaseModelOutput(last_hidden_state=cond),
                     past_key_values=past, use_cache=True)
        past = out.past_key_values
        lg = be.step(seq[:, -1])
        assert np.abs(lg.numpy() - out.logits[:, -1].numpy()).max() < 2e-5
        nxt = out.logits[:, -1].argmax(-1, keepdim=True)
        seq = torch.cat([seq, nxt], 1)
        if t == 2:                                         # swap two beams in both caches
            idx = torch.tensor([1, 0, 2, 3])
            past.reorder_cache(idx)
            be.reorder(idx)
            seq = seq[idx]
            cond = cond[idx]


@pytest.mark.parametrize("seed,kw", [
    (1, dict(REF_KW, max_new_tokens=24, min_new_tokens=18)),
    (2, dict(REF_KW, max_new_tokens=30, min_new_tokens=0)),
    (3, dict(max_new_tokens=14, num_beams=1, eos_token_id=1, pad_token_id=0)),
    (4, dict(max_new_tokens=12, num_beams=2, length_penalty=0.8, early_stopping=False, eos_token_id=1, pad_token_id=0)),
    (5, dict(max_new_tokens=16, num_beams=3, no_repeat_ngram_size=2, length_penalty=2.0, early_stopping="never", eos_token_id=1,
             pad_token_id=0)),
])
def test_native_beam_search_token_identical_to_hf(t5, seed, kw):
    """The native search loop (NativeBeamSearch: HF's beam-search algorithm restated over the step + score/top-k
    interface) against stock HF generate - here with the torch step and torch scoring, so this pins the search LOGIC."""
    from mmdx_b200.t5_fast import NativeBeamSearch
    torch.manual_seed(seed)
    cond = torch.randn(3, 4, 512)
    with torch.no_grad():
        want = t5.generate(encoder_outputs=BaseModelOutput(last_hidden_state=cond), **kw)
    got = NativeBeamSearch(TorchStep(t5), t5.config).generate(cond, **kw)
    assert want.shape == got.shape and torch.equal(want, got), (want.tolist(), got.tolist())


class _EosBiased:
    """Step backend wrapper that raises the EOS logit by `b` (the HF side gets the same through a forward hook on lm_head):
    with random-init weights EOS never wins on its own, and the interesting part of a beam search is what happens when
    hypotheses finish at different lengths."""

    def __init__(self, be, b):
        self.be, self.b = be, b

    def __getattr__(self, k):
        return getattr(self.be, k)

    def step(self, tok):
        lg = self.be.step(tok)
        lg[:, 1] += self.b
        return lg


@pytest.mark.parametrize("bias", [5.0, 6.5])
def test_native_beam_search_with_finishing_hypotheses(t5, bias):
    """EOS made likely: hypotheses finish at different steps, the finished pool fills up and competes on
    score / length ** length_penalty, early stopping ends the search, shorter outputs are padded the way HF pads them
    (with EOS when pad_token_id is 0).  Native search == HF generate == HF's loop over the same step, shapes included."""
    from mmdx_b200.t5_fast import NativeBeamSearch

    def hook(mod, inp, out):
        out[..., 1] += bias
        return out

    n_eos = 0
    for seed in range(6):
        torch.manual_seed(100 + seed)
        cond = torch.randn(2, 4, 512)
        kw = dict(max_new_tokens=24, min_new_tokens=(0 if seed % 2 else 3), num_beams=4, length_penalty=1.1,
                  early_stopping=(True if seed % 3 else False), eos_token_id=1, pad_token_id=0, no_repeat_ngram_size=3)
        h = t5.lm_head.register_forward_hook(hook)
        try:
            with torch.no_grad():
                want = t5.generate(encoder_outputs=BaseModelOutput(last_hidden_state=cond), **kw)
        finally:
            h.remove()
        got = NativeBeamSearch(_EosBiased(TorchStep(t5), bias), t5.config).generate(cond, **kw)
        assert want.shape == got.shape and torch.equal(want, got), (seed, want.tolist(), got.tolist())
        got2 = FastT5Generator(t5, _EosBiased(TorchStep(t5), bias)).generate(cond, **kw)
        assert torch.equal(want, got2)
        n_eos += int((want == 1).any())
    assert n_eos >= 3


def test_native_search_argument_support():
    """inference() sends a request to the native search only when every generation argument is one it restates
    (NativeBeamSearch.supports); anything else keeps HF's own search code in the loop."""
    from mmdx_b200.t5_fast import NativeBeamSearch
    ref = dict(max_new_tokens=180, min_new_tokens=150, num_beams=4, no_repeat_ngram_size=3, length_penalty=1.1,
               early_stopping=True, eos_token_id=1, pad_token_id=0)          # inference_pipeline.py:190
    assert NativeBeamSearch.supports(ref)
    assert NativeBeamSearch.supports(dict(ref, num_beams=1, early_stopping="never", do_sample=False, repetition_penalty=None))
    for extra in (dict(do_sample=True), dict(repetition_penalty=1.2), dict(num_beams=5), dict(eos_token_id=[1, 2]),
                  dict(max_new_tokens=None), dict(num_return_sequences=2), dict(top_k=50)):
        assert not NativeBeamSearch.supports(dict(ref, **extra)), extra
