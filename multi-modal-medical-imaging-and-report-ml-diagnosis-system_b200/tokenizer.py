"""Native WordPiece tokenizer behind `tokenize_patient_details`.

`NativeBertTokenizer` wraps the C++ tokenizer of libmmdx.so (csrc/tokenizer.cpp, host-only): it is built from the
bundle's HF `BertTokenizer` (same vocabulary, same lower-casing) and returns the same `input_ids` / `attention_mask` /
`token_type_ids` arrays, bit for bit, for 7-bit ASCII text - which is all the reference's patient-details grammar
produces (backend/ml/data_prep/raw_data_pre_preparation.py:114-163).  Strings it does not handle (non-ASCII bytes,
literal special-token text) are sent through the HF tokenizer itself, one call for all of them, so the result is
always what HF would have returned.  HF's Python wrapper sustains ~5-18 k strings/s; this does > 1 M/s per core."""
from __future__ import annotations

import ctypes as C
import json

import numpy as np

from ._lib import MmdxError, lib


def _supported(hf_tok):
    """The restated pipeline is exactly BertNormalizer(clean, lowercase = strip accents) -> BertPreTokenizer ->
    WordPiece('##', 100, [UNK]) -> [CLS] $A [SEP]; anything else stays with HF."""
    try:
        cfg = json.loads(hf_tok.backend_tokenizer.to_str())
        nz, pt, md, pp = cfg["normalizer"], cfg["pre_tokenizer"], cfg["model"], cfg["post_processor"]
        ok = nz["type"] == "BertNormalizer" and nz["clean_text"] and nz["handle_chinese_chars"]
        ok = ok and (nz["strip_accents"] is None or nz["strip_accents"] == nz["lowercase"])
        ok = ok and pt["type"] == "BertPreTokenizer"
        ok = ok and md["type"] == "WordPiece" and md["continuing_subword_prefix"] == "##" \
            and md["max_input_chars_per_word"] == 100 and md["unk_token"] == "[UNK]"
        ok = ok and pp["type"] == "TemplateProcessing"
        specials = {t["content"] for t in cfg.get("added_tokens", [])}
        ok = ok and specials <= {"[PAD]", "[UNK]", "[CLS]", "[SEP]", "[MASK]"}
        return bool(ok), bool(nz["lowercase"]), md["vocab"]
    except Exception:      # noqa: BLE001 - not a tokenizers-backed BERT tokenizer
        return False, True, None


class NativeBertTokenizer:
    def __init__(self, hf_tok, n_threads: int = 0):
        ok, lower, vocab = _supported(hf_tok)
        if not ok:
            raise MmdxError("the bundle's tokenizer is not a plain BERT WordPiece tokenizer; keep using it directly")
        self.hf = hf_tok
        self.n_threads = n_threads
        toks = [None] * (max(vocab.values()) + 1)
        for w, i in vocab.items():
            toks[i] = w
        blob = "\n".join("[unused-hole]" if w is None else w for w in toks).encode("utf-8")
        self._h = C.c_void_p()
        if lib().mmdx_tokenizer_create(blob, len(blob), 1 if lower else 0, C.byref(self._h)) != 0:
            raise MmdxError(lib().mmdx_tokenizer_last_error().decode())
        self.fallbacks = 0           # strings that went through HF (telemetry for the bench / tests)

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            lib().mmdx_tokenizer_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:      # noqa: BLE001
            pass

    def encode_batch(self, texts, max_len: int = 96):
        """-> (ids int32 [n,max_len], lens int32 [n]); pad id where position >= len."""
        n = len(texts)
        enc = [t.encode("utf-8") for t in texts]
        offs = np.zeros(n + 1, np.int64)
        if n:
            np.cumsum(np.fromiter((len(e) for e in enc), np.int64, n), out=offs[1:])
        blob = b"".join(enc)
        ids = np.empty((n, max_len), np.int32)
        lens = np.empty(n, np.int32)
        fb = np.empty(n, np.uint8)
        rc = lib().mmdx_tokenize_batch(self._h, blob, offs.ctypes.data, n, int(max_len), int(self.n_threads), ids.ctypes.data,
                                       lens.ctypes.data, fb.ctypes.data)
        if rc != 0:
            raise MmdxError(lib().mmdx_tokenizer_last_error().decode())
        if fb.any():
            idx = np.nonzero(fb)[0]
            self.fallbacks += len(idx)
            o = self.hf([texts[i] for i in idx], padding="max_length", truncation=True, return_tensors="np", max_length=max_len)
            ids[idx] = o["input_ids"]
            lens[idx] = o["attention_mask"].sum(1)
        return ids, lens

    def __call__(self, texts, padding="max_length", truncation=True, return_tensors="np", max_length=96):
        """Same call shape as tokenize_patient_details uses on the HF tokenizer (training_pipeline.py:338-342)."""
        if padding != "max_length" or not truncation or return_tensors != "np":
            raise ValueError("NativeBertTokenizer implements padding='max_length', truncation=True, return_tensors='np'")
        ids, lens = self.encode_batch(list(texts), max_length)
        mask = (np.arange(max_length, dtype=np.int32)[None, :] < lens[:, None]).astype(np.int64)
        return {"input_ids": ids.astype(np.int64), "token_type_ids": np.zeros_like(mask), "attention_mask": mask}
