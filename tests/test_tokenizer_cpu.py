"""the native WordPiece tokenizer (csrc/tokenizer.cpp, host-only part of libmmdx.so) against the HF
BertTokenizer it replaces (tokenize_patient_details, training_pipeline.py:323,335-342) - integer rows, so the bar is
bit-exact: same input_ids, attention_mask and token_type_ids for every string."""
import string

import numpy as np
import pytest

from conftest import load_golden
from mmdx_b200 import synth
from mmdx_b200.tokenizer import NativeBertTokenizer


@pytest.fixture(scope="module", autouse=True)
def _built():
    from mmdx_b200 import _lib
    _lib.build()                         # no-op when libmmdx.so is up to date (host-only code: no GPU needed)


@pytest.fixture(scope="module")
def toks():
    hf = synth.make_bert_tokenizer()
    return hf, NativeBertTokenizer(hf)


def _same(hf, nt, texts, max_len):
    a = hf(texts, padding="max_length", truncation=True, return_tensors="np", max_length=max_len)
    b = nt(texts, max_length=max_len)
    for k in ("input_ids", "attention_mask", "token_type_ids"):
        if not np.array_equal(a[k], b[k]):
            bad = int(np.nonzero((a[k] != b[k]).any(1))[0][0])
            raise AssertionError((k, texts[bad], a[k][bad].tolist(), b[k][bad].tolist()))


def test_patient_details_grammar_and_reference_samples(toks):
    hf, nt = toks
    texts = synth.synth_details(2000, seed=5)
    g1 = load_golden("g1_samples")
    texts += [str(t) for t in g1["details"]]
    texts.append("44 year old female PA view , hypertension , cough")            # inference_pipeline.py:223
    for L in (96, 128, 512, 16):
        _same(hf, nt, texts, L)
    assert nt.fallbacks == 0
    # the goldens' own ids (made by the reference's tokenize_patient_details) come back exactly
    out = nt([str(t) for t in g1["details"]], max_length=96)
    assert np.array_equal(out["input_ids"], g1["input_ids"]) and np.array_equal(out["attention_mask"], g1["attention_mask"])


def test_adversarial_ascii(toks):
    hf, nt = toks
    texts = [
        "", " ", "   \t\n  ", "a", "A", ".", "...", "a.b,c;d", "cough,fever!!", "  leading and trailing  ",
        "MiXeD CaSe CoUgH", "x" * 99, "x" * 100, "x" * 101, "x" * 300, "cough" * 30, "co\x01ugh", "co\x00ugh fe\x7fver",
        "tab\tseparated\nlines\r\nhere", "\x0b\x0cvertical", "unknownword zzzqqq", "123 45 6789 120 121 0042",
        "e-mail: a@b.c (test) [x] {y} <z> 50% #1 $5 ^ ~ ` | \\ / _under_score_", "it's \"quoted\"", "a" + "." * 150,
        " ".join(["cough"] * 600), "-".join(["fever"] * 300), "year" + "\x1f" + "old", "##cough ##", "# # #a", "[", "]", "[x]",
        "[ SEP ]", "[sep]", "[Sep]",
    ]
    for L in (2, 3, 4, 8, 96, 512):
        _same(hf, nt, texts, L)


def test_fallback_strings_go_through_hf(toks):
    hf, nt = toks
    n0 = nt.fallbacks
    texts = ["café cough", "patient [SEP] cough", "fever 中文 pain", "plain ascii", "[CLS]", "naïve [MASK] x",
             "Édema", "[UNK]", "[PAD] [PAD]"]
    _same(hf, nt, texts, 32)
    assert nt.fallbacks - n0 == 8


def test_random_ascii_strings(toks):
    hf, nt = toks
    rng = np.random.Generator(np.random.PCG64(123))
    words = ["cough", "fever", "year", "old", "male", "PA", "view", "chest", "pain", "zz", "q", "12", "7", "smoking"]
    alphabet = string.ascii_letters + string.digits + string.punctuation + "     \t\n"
    texts = []
    for _ in range(3000):
        if rng.random() < 0.5:
            n = int(rng.integers(0, 60))
            texts.append("".join(alphabet[int(i)] for i in rng.integers(0, len(alphabet), n)))
        else:
            n = int(rng.integers(0, 40))
            sep = [" ", ", ", "-", "  ", ".", ""]
            texts.append("".join(words[int(rng.integers(len(words)))] + sep[int(rng.integers(len(sep)))] for _ in range(n)))
    texts = [t for t in texts if "[" not in t or True]
    _same(hf, nt, texts, 48)


def test_longest_match_with_a_subword_vocabulary():
    """A vocabulary with overlapping multi-character pieces: greedy longest-match-first, '##' continuations, and the
    whole word collapsing to one [UNK] when a remainder cannot be matched."""
    from transformers import BertTokenizer
    base = ["[PAD]"] + [f"[unused{i}]" for i in range(99)] + ["[UNK]", "[CLS]", "[SEP]", "[MASK]"]
    pieces = ["a", "b", "c", "ab", "abc", "abcd", "##a", "##b", "##c", "##d", "##cd", "##bcd", "##de", "##e", "card", "##io",
              "cardio", "##meg", "##aly", "##megaly", "##m", "##eg", "x", "##xy", "1", "##1", "11", ",", ".", "##."]
    vocab = {w: i for i, w in enumerate(base + pieces)}
    hf = BertTokenizer(vocab=vocab, do_lower_case=True)
    nt = NativeBertTokenizer(hf)
    texts = ["abcd", "abcde", "abcdf", "ab cd", "abab", "cardiomegaly", "Cardiomegaly, cardio.", "cardiomeg", "cardiox",
             "xxy", "xy", "111", "1111", "a1", "1a", "abcdabcd", "abcdeabc", "e", "d", "cde", "a,b.c", "ABCD ABCDE"]
    rng = np.random.Generator(np.random.PCG64(9))
    for _ in range(2000):
        n = int(rng.integers(1, 14))
        texts.append("".join("abcdex1 ,."[int(i)] for i in rng.integers(0, 10, n)))
    _same(hf, nt, texts, 24)
    nt.close()


def test_inference_pipeline_uses_the_native_tokenizer():
    """inference_pipeline.tokenize(): native for the bundle's BERT tokenizer, identical arrays to the HF call."""
    from mmdx_b200 import inference_pipeline as ip
    hf = synth.make_bert_tokenizer()
    bundle = {"bert_tok": hf}
    texts = synth.synth_details(64, seed=8)
    a = hf(texts, padding="max_length", truncation=True, return_tensors="np", max_length=96)
    b = ip.tokenize(bundle, texts, 96)
    assert isinstance(ip.native_tokenizer(bundle), NativeBertTokenizer)
    for k in ("input_ids", "attention_mask", "token_type_ids"):
        assert np.array_equal(a[k], b[k])
    assert ip.native_tokenizer(bundle) is ip.native_tokenizer(bundle)
