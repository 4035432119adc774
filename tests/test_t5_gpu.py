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


@pytest.mark.parametrize("studies,beams,n_enc,nofold", [(1, 4, 7, False), (2, 4, 4, True), (5, 4, 4, False), (2, 3, 16, False)])
def test_step_variants_match_torch_step(t5, monkeypatch, studies, beams, n_enc, nofold):
    """The step kernel's other shapes against the fp32 torch step: a conditioning length whose folded cross-attention
    matrix needs zero padding (1 x 4 rows x 8 heads x 7 tokens = 224 -> 256 columns), the unfolded cross-attention path
    (forced, and chosen by itself at 20 rows where the folded score matrix would be too wide), three row chunks, and a
    beam count that is not a power of two."""
    if nofold:
        monkeypatch.setenv("MMDX_T5_NOFOLD", "1")
    torch.manual_seed(11)
    R = studies * beams
    cond = torch.randn(studies, n_enc, 512, device="cuda").repeat_interleave(beams, 0)
    a, b = TorchStep(t5), MmdxStep(t5)
    a.begin(cond, R, 12)
    b.begin(cond, R, 12)
    tok = torch.zeros(R, dtype=torch.long, device="cuda")
    for t in range(8):
        la, lb = a.step(tok), b.step(tok)
        torch.cuda.synchronize()
        err = float((la - lb).abs().max() / la.abs().max())
        assert err < 2e-5, (R, t, err)
        tok = la.topk(2, -1).indices[torch.arange(R), torch.arange(R) % 2]
        if t == 4 and beams > 1:
            idx = torch.cat([s0 * beams + torch.tensor(([1, 1, 0, 3] if beams == 4 else [2, 0, 0])) for s0 in range(studies)]).cuda()
            a.reorder(idx)
            b.reorder(idx)
            tok = tok[idx]
    b.close()


def test_step_random_shapes_match_torch_step(t5):
    """Fuzz: 20 random shapes of the decoder step (1-6 studies, 1-4 beams, 1-16 conditioning tokens, 4-24 positions, random
    beam reorders inside every study) against the fp32 torch step; a 60-case run is kept in profiles/r02_t5_fuzz.log."""
    g = torch.Generator().manual_seed(5)
    a, b = TorchStep(t5), MmdxStep(t5)
    for c in range(20):
        studies = int(torch.randint(1, 7, (1,), generator=g))
        beams = int(torch.randint(1, 5, (1,), generator=g))
        n_enc = [1, 2, 3, 4, 4, 4, 5, 7, 9, 16][int(torch.randint(0, 10, (1,), generator=g))]
        steps = int(torch.randint(4, 25, (1,), generator=g))
        R = studies * beams
        cond = torch.randn(studies, n_enc, 512, generator=g).cuda().repeat_interleave(beams, 0)
        a.begin(cond, R, steps)
        b.begin(cond, R, steps)
        tok = torch.randint(0, 32128, (R,), generator=g).cuda()
        for t in range(steps):
            la, lb = a.step(tok), b.step(tok)
            torch.cuda.synchronize()
            err = float((la - lb).abs().max() / la.abs().max())
            assert err < 2e-5, (c, studies, beams, n_enc, t, err)
            tok = la.topk(3, -1).indices[torch.arange(R), torch.randint(0, 3, (R,), generator=g)]
            if beams > 1 and t < steps - 1 and int(torch.randint(0, 3, (1,), generator=g)) == 0:
                idx = torch.cat([s0 * beams + torch.randint(0, beams, (beams,), generator=g) for s0 in range(studies)]).cuda()
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
    # the default is the native search (one C call per report); "hf" keeps HF's own beam search over the CUDA step, "native"
    # insists on the native search; arguments the native search does not restate fall back to the HF loop by default
    def launches():
        return sum(g[0].backend.launch_count for g in ip._REPORT_GENS.values() if g[0] is not None)
    launches0 = launches()
    hf = ip.inference(dict(b, fast_report="hf"), pil, str(g1["details"][0]), device="cuda", gen_kwargs=kw)
    native = ip.inference(dict(b, fast_report="native"), pil, str(g1["details"][0]), device="cuda", gen_kwargs=kw)
    assert hf["report_text"] == slow["report_text"] and native["report_text"] == slow["report_text"]
    assert native["disease_probs"] == fast["disease_probs"]
    assert launches() > launches0
    # a whole request at the reference's default generation settings (gen_kwargs=None: 180 new tokens, 4 beams)
    ip.inference(b, pil, str(g1["details"][0]), device="cuda")
    ts = []
    for _ in range(3):
        t0 = time.perf_counter()
        full = ip.inference(b, pil, str(g1["details"][0]), device="cuda")
        ts.append(time.perf_counter() - t0)
    print(f"inference() with the default report settings: {min(ts) * 1e3:.1f} ms per request "
          f"({len(full['report_text'].split())} report tokens)")
    sampled = ip.inference(b, pil, str(g1["details"][0]), device="cuda", gen_kwargs=dict(kw, repetition_penalty=1.3))
    want = ip.inference(dict(b, fast_report=False), pil, str(g1["details"][0]), device="cuda", gen_kwargs=dict(kw, repetition_penalty=1.3))
    assert sampled["report_text"] == want["report_text"]
    with pytest.raises(ValueError):
        ip.inference(dict(b, fast_report="native"), pil, "x", device="cuda", gen_kwargs=dict(kw, do_sample=True))


def test_native_beam_search_on_the_cuda_kernels(t5):
    """NativeBeamSearch over the CUDA step + the scoring / top-k kernels (mmdx_t5_score_topk) against stock HF generate on
    the GPU, at the reference's full generation settings (inference_pipeline.py:190: 180 new tokens, 150 minimum, 4 beams,
    no-repeat 3-gram, length penalty 1.1, early stopping) and at short ones; prints the three timings."""
    from mmdx_b200.t5_fast import NativeBeamSearch
    g = load_golden("g2_B8_L128_ragged")
    for n_new, n_min, B in ((180, 150, 2), (32, 0, 3)):
        cond = torch.from_numpy(g["cond"][:B]).cuda()
        kw = dict(REF_KW, max_new_tokens=n_new, min_new_tokens=n_min)
        with torch.no_grad():
            t0 = time.perf_counter()
            want = t5.generate(encoder_outputs=BaseModelOutput(last_hidden_state=cond), **kw)
            torch.cuda.synchronize()
            t_hf = time.perf_counter() - t0
        step = MmdxStep(t5)
        loop = FastT5Generator(t5, step)
        t0 = time.perf_counter()
        got_loop = loop.generate(cond, **kw)
        torch.cuda.synchronize()
        t_loop = time.perf_counter() - t0
        nat = NativeBeamSearch(step, t5.config, host_loop="python")
        nat.generate(cond, **dict(kw, max_new_tokens=4, min_new_tokens=0))
        t0 = time.perf_counter()
        got = nat.generate(cond, **kw)
        torch.cuda.synchronize()
        t_nat = time.perf_counter() - t0
        cpp = NativeBeamSearch(step, t5.config)                      # bookkeeping in C++ (mmdx_t5_generate)
        cpp.generate(cond, **dict(kw, max_new_tokens=4, min_new_tokens=0))
        t0 = time.perf_counter()
        got_cpp = cpp.generate(cond, **kw)
        t_cpp = time.perf_counter() - t0
        print(f"generate {n_new} tokens x 4 beams x {B} studies: HF eager {t_hf * 1e3:.0f} ms, HF loop over the CUDA step "
              f"{t_loop * 1e3:.0f} ms, native search (python bookkeeping) {t_nat * 1e3:.0f} ms, (C++ bookkeeping) {t_cpp * 1e3:.0f} ms")
        assert torch.equal(want, got_loop)
        assert want.shape == got.shape and torch.equal(want.cpu(), got), (want.tolist(), got.tolist())
        assert want.shape == got_cpp.shape and torch.equal(want.cpu(), got_cpp), (want.tolist(), got_cpp.tolist())
        step.close()


def test_cpp_search_with_finishing_hypotheses(t5):
    """mmdx_t5_generate against the Python bookkeeping (pinned against HF on the CPU) with EOS made likely - hypotheses
    finishing at different lengths, early stopping, HF's padding - and across beam counts / penalties / stopping modes."""
    import copy
    from mmdx_b200.t5_fast import NativeBeamSearch
    m = copy.deepcopy(t5)
    torch.manual_seed(7)
    with torch.no_grad():
        # random-init decoder states share a strong common direction; an EOS embedding along it makes EOS competitive at
        # every step (tied LM head), so hypotheses finish at different lengths
        ids = torch.randint(0, 32128, (4, 12), device="cuda")
        out = m(decoder_input_ids=ids, encoder_outputs=BaseModelOutput(last_hidden_state=torch.randn(4, 4, 512, device="cuda")),
                output_hidden_states=True)
        u = out.decoder_hidden_states[-1].reshape(-1, 512)
        v = (u / u.norm(dim=-1, keepdim=True)).mean(0)
        m.shared.weight[1] = 14.0 * v / v.norm()
    step = MmdxStep(m)
    n_short = 0
    for seed in range(8):
        torch.manual_seed(200 + seed)
        cond = torch.randn(2, 4, 512, device="cuda")
        kw = dict(max_new_tokens=28, min_new_tokens=(0 if seed % 2 else 3), num_beams=(4 if seed % 4 else 2),
                  length_penalty=(1.1 if seed % 3 else 0.7), early_stopping=(True, False, "never")[seed % 3], eos_token_id=1,
                  pad_token_id=0, no_repeat_ngram_size=(3 if seed % 2 else 2))
        a = NativeBeamSearch(step, m.config, host_loop="python").generate(cond, **kw)
        b = NativeBeamSearch(step, m.config).generate(cond, **kw)
        assert a.shape == b.shape and torch.equal(a, b), (seed, a.tolist(), b.tolist())
        n_short += int(a.shape[1] < 29 or bool((a[:, 1:] == 1).any()))
    print("runs with an EOS / early finish:", n_short)
    assert n_short >= 3
    step.close()


def test_score_topk_kernel_vs_torch(t5):
    """mmdx_t5_score_topk against the torch restatement: same candidates in the same order, scores to fp32 rounding."""
    torch.manual_seed(5)
    R, K, V = 12, 4, t5.config.vocab_size
    a, b = TorchStep(t5), MmdxStep(t5)
    cond = torch.randn(R // K, 4, 512, device="cuda").repeat_interleave(K, 0)
    a.begin(cond, R, 4)
    b.begin(cond, R, 4)
    tok = torch.randint(0, V, (R,), device="cuda")
    la, lb = a.step(tok), b.step(tok)
    bs = torch.randn(R) * 3
    banned = torch.full((R, 5), -1, dtype=torch.int32)
    banned[:, 0] = la.argmax(-1).cpu().int()                    # ban every row's best token
    banned[3, 1] = 17
    for ban_eos in (False, True):
        sa, ia = a.score_topk(la.clone(), bs, banned, ban_eos, 1, K, 8)
        sb, ib = b.score_topk(lb.clone(), bs, banned, ban_eos, 1, K, 8)
        assert torch.equal(ia, ib)
        assert float((sa - sb).abs().max()) < 1e-4
    b.close()


def test_decoder_c_abi_error_behaviour(t5):
    """mmdx_t5_* misuse fails with a message instead of touching memory: steps before begin / beyond the reserved length,
    more beams than the top-8 kernel serves, rows that are not studies x beams, an unfinished engine."""
    import ctypes as C
    from mmdx_b200._lib import MmdxError, lib
    step = MmdxStep(t5)
    tok = torch.zeros(4, dtype=torch.long, device="cuda")
    with pytest.raises(MmdxError, match="before mmdx_t5_begin"):
        step._rows, step._vocab = 4, 32128
        step.step(tok)
    step.begin(torch.randn(4, 4, 512, device="cuda"), 4, 2)
    step.step(tok)
    step.step(tok)
    with pytest.raises(MmdxError, match="more steps"):
        step.step(tok)
    cond = torch.randn(1, 4, 512, device="cuda")
    kw = dict(max_new_tokens=4, min_new_tokens=0, no_repeat_ngram_size=0, length_penalty=1.0, early_stopping=False,
              eos_token_id=1, pad_token_id=0, decoder_start_token_id=0)
    with pytest.raises(MmdxError, match="num_beams"):
        step.generate_native(cond, num_beams=5, **kw)
    step.begin(torch.randn(4, 4, 512, device="cuda"), 4, 4)
    lg = step.step(tok)
    with pytest.raises(MmdxError, match="studies x beams"):
        step.score_topk(lg, torch.zeros(4), None, False, 1, 3, 6)
    with pytest.raises(MmdxError, match="<= 8"):
        step.score_topk(lg, torch.zeros(4), None, False, 1, 4, 9)
    step.close()
    h = C.c_void_p()
    assert lib().mmdx_t5_create(0, 512, 8, 64, 2048, 6, 32128, 1e-6, 1, C.byref(h)) == 0
    assert lib().mmdx_t5_finalize(h) != 0 and b"missing" in lib().mmdx_t5_last_error()
    enc = torch.zeros(1, 4, 512, device="cuda")
    bias = torch.zeros(4, 8)
    assert lib().mmdx_t5_begin(h, C.c_void_p(enc.data_ptr()), 1, 4, 4, C.c_void_p(bias.data_ptr()), None) != 0
    assert b"finalized" in lib().mmdx_t5_last_error()
    assert lib().mmdx_t5_create(0, 500, 8, 64, 2048, 6, 32128, 1e-6, 1, C.byref(C.c_void_p())) != 0     # d_model % 128
    lib().mmdx_t5_destroy(h)


def test_large_report_batches_are_row_independent(t5):
    """33 studies x 4 beams = 132 rows (nine passes of 16 rows, unfolded cross-attention) through mmdx_t5_generate: the
    reports of the first two studies equal what a 2-study call (8 rows, folded cross-attention) generates."""
    torch.manual_seed(21)
    step = MmdxStep(t5)
    cond = torch.randn(33, 4, 512, device="cuda")
    kw = dict(max_new_tokens=16, min_new_tokens=12, num_beams=4, no_repeat_ngram_size=3, length_penalty=1.1,
              early_stopping=True, eos_token_id=1, pad_token_id=0, decoder_start_token_id=0)
    big = step.generate_native(cond, **kw)
    small = step.generate_native(cond[:2], **kw)
    assert big.shape[0] == 33 and torch.equal(big[:2, :small.shape[1]], small)
    step.close()
