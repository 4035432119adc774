"""SURVEY.md 8f N1 on the GPU: the CUDA decoder-step kernels (csrc/t5_decoder.cu through mmdx_t5_*) against the fp32 torch
step and against stock HF generate - logits within 2e-5 relative, generated tokens identical - and the drop-in
inference() producing its report through them."""
import time

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from transformers import T5Config, T5ForConditionalGeneration          # noqa: E402
from transformers.modeling_outputs import BaseModelOutput               # noqa: E402

from conftest import load_golden                                        # noqa: E402
from mmdx_b200 import synth                                             # noqa: E402
from mmdx_b200 import inference_pipeline as ip                          # noqa: E402
from mmdx_b200.t5_fast import FastT5Generator, MmdxStep                 # noqa: E402
from oracle.t5_step_ref import TorchStep                                # noqa: E402

REF_KW = dict(num_beams=4, no_repeat_ngram_size=3, length_penalty=1.1, early_stopping=True, eos_token_id=1, pad_token_id=0)


@pytest.fixture(scope="module")
def t5():
    torch.manual_seed(0)
    return T5ForConditionalGeneration(T5Config(decoder_start_token_id=0)).eval().cuda()


def test_step_logits_match_torch_step_with_reorder(t5):
    """Ten steps of R = studies x beams rows, with two beam reorders (a hypothesis dropped, one duplicated - always inside
    a study, as beam search does): logits of the CUDA step against the fp32 torch step on the same GPU."""
    torch.manual_seed(3)
    for studies, beams in ((1, 1), (1, 4), (3, 4)):
        R = studies * beams
        cond = torch.randn(studies, 4, 512, device="cuda").repeat_interleave(beams, 0)
        a, b = TorchStep(t5), MmdxStep(t5)
        a.begin(cond, R, 16)
        b.begin(cond, R, 16)
        tok = torch.zeros(R, dtype=torch.long, device="cuda")
        for t in range(10):
            la, lb = a.step(tok), b.step(tok)
            torch.cuda.synchronize()
            err = float((la - lb).abs().max() / la.abs().max())
            assert err < 2e-5, (R, t, err)
            tok = la.topk(2, -1).indices[torch.arange(R), torch.arange(R) % 2]      # different continuations per beam
            if t in (3, 6) and beams > 1:
                idx = torch.cat([s0 * beams + torch.tensor([1, 1, 0, 3]) for s0 in range(studies)]).cuda()
                a.reorder(idx)
                b.reorder(idx)
                tok = tok[idx]
        b.close()


@pytest.mark.parametrize("kw", [
    dict(REF_KW, max_new_tokens=40, min_new_tokens=30),
    dict(max_new_tokens=24, num_beams=1, eos_token_id=1, pad_token_id=0),
])
def test_generate_token_identical_to_hf_on_gpu(t5, kw):
    g = load_golden("g2_B8_L128_ragged")
    cond = torch.from_numpy(g["cond"][:3]).cuda()
    with torch.no_grad():
        t0 = time.perf_counter()
        want = t5.generate(encoder_outputs=BaseModelOutput(last_hidden_state=cond), **kw)
        torch.cuda.synchronize()
        t_hf = time.perf_counter() - t0
    gen = FastT5Generator(t5, MmdxStep(t5))
    gen.generate(cond, **dict(kw, max_new_tokens=4, min_new_tokens=0))          # warm-up
    t0 = time.perf_counter()
    got = gen.generate(cond, **kw)
    torch.cuda.synchronize()
    t_fast = time.perf_counter() - t0
    print(f"generate {kw['max_new_tokens']} tokens x {kw['num_beams']} beams x 3 studies: HF eager {t_hf * 1e3:.0f} ms, "
          f"mmdx step {t_fast * 1e3:.0f} ms ({gen.backend.launch_count} kernel launches)")
    assert torch.equal(want, got)


def test_inference_report_text_through_the_cuda_decoder(state_bundle, g1):
    """inference() with a fusion module that carries a T5 report model: report_text comes from HF's beam search over the
    CUDA decoder step and equals what the stock HF path (fast_report=False) returns."""
    from PIL import Image
    fs = state_bundle["fusion_state"]

    class Fusion(torch.nn.Module):
        def __init__(self):
            super().__init__()
            torch.manual_seed(0)
            self.n_cond, self.h_dec = 4, 512
            self.report_model = T5ForConditionalGeneration(T5Config(decoder_start_token_id=0)).eval()

        def state_dict(self, *a, **k):
            return fs

    class Tok:
        eos_token_id, pad_token_id = 1, 0

        def batch_decode(self, ids, skip_special_tokens=True):
            return [" ".join(str(int(t)) for t in row) for row in ids]

    b = dict(state_bundle)
    b["bert_tok"] = synth.make_bert_tokenizer()
    b["fusion_model"], b["t5_tok"] = Fusion(), Tok()
    pil = Image.fromarray(np.repeat(g1["gray"][0][..., None], 3, axis=-1))
    kw = dict(max_new_tokens=20, min_new_tokens=16)
    fast = ip.inference(b, pil, str(g1["details"][0]), device="cuda", gen_kwargs=kw)
    assert ip.fast_report_generator(b["fusion_model"], torch.device("cuda", 0)) is not None
    slow = ip.inference(dict(b, fast_report=False), pil, str(g1["details"][0]), device="cuda", gen_kwargs=kw)
    assert fast["report_text"] == slow["report_text"] and len(fast["report_text"].split()) >= 17
    assert fast["disease_vector"] == g1["inf_vector"][0].tolist()
