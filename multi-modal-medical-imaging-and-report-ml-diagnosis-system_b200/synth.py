"""Seeded synthetic studies and random-init weights for the hot path.

There is no network on the build or GPU boxes, so no trained bundle
(`backend/ml/model/model_bundle.pt`, git-ignored in the reference) and no
pretrained BERT / ResNet weights exist.  Everything here is generated from a
seed with numpy's PCG64 so that this container (where the golden fixtures are
made from the reference's own modules) and the GPU box produce bit-identical
inputs and weights.

State-dict key names and shapes follow the reference's three modules
(`backend/ml/pipelines/training_pipeline.py:157-311` ImageEncoderCNN,
`:348-508` TextEncoderTransformer, `:516-618` FusionTransformerModel) - they
are the weight-import contract of the engine.
"""
from __future__ import annotations

import zlib

import numpy as np
import torch

# Same order as the reference's default class list
# (backend/ml/pipelines/inference_pipeline.py:121-125).
CLASS_NAMES = [
    "No Finding", "Enlarged Cardiomediastinum", "Cardiomegaly", "Lung Opacity",
    "Lung Lesion", "Edema", "Consolidation", "Pneumonia", "Atelectasis",
    "Pneumothorax", "Pleural Effusion", "Pleural Other", "Fracture",
]

RESNET_STAGES = ((64, 3, 1), (128, 4, 2), (256, 6, 2), (512, 3, 2))  # (planes, blocks, stride)


def _rng(seed: int, key: str) -> np.random.Generator:
    return np.random.Generator(np.random.PCG64([seed, zlib.crc32(key.encode())]))


def _normal(seed, key, shape, std, mean=0.0):
    a = _rng(seed, key).standard_normal(size=shape, dtype=np.float32)
    return torch.from_numpy(a * np.float32(std) + np.float32(mean))


def _uniform(seed, key, shape, lo, hi):
    a = _rng(seed, key).random(size=shape, dtype=np.float32)
    return torch.from_numpy(a * np.float32(hi - lo) + np.float32(lo))


def _bn(sd, seed, prefix, c, gamma=(0.5, 1.5)):
    sd[prefix + ".weight"] = _uniform(seed, prefix + ".weight", (c,), *gamma)
    sd[prefix + ".bias"] = _normal(seed, prefix + ".bias", (c,), 0.1)
    sd[prefix + ".running_mean"] = _normal(seed, prefix + ".running_mean", (c,), 0.1)
    sd[prefix + ".running_var"] = _uniform(seed, prefix + ".running_var", (c,), 0.5, 1.5)
    sd[prefix + ".num_batches_tracked"] = torch.tensor(0, dtype=torch.long)


def _conv(sd, seed, key, cout, cin, k):
    std = (2.0 / (cin * k * k)) ** 0.5
    sd[key] = _normal(seed, key, (cout, cin, k, k), std)


def _linear(sd, seed, prefix, nout, nin, std=None, bias_std=0.02):
    std = std if std is not None else (1.0 / nin) ** 0.5
    sd[prefix + ".weight"] = _normal(seed, prefix + ".weight", (nout, nin), std)
    sd[prefix + ".bias"] = _normal(seed, prefix + ".bias", (nout,), bias_std)


def _ln(sd, seed, prefix, n):
    sd[prefix + ".weight"] = _uniform(seed, prefix + ".weight", (n,), 0.8, 1.2)
    sd[prefix + ".bias"] = _normal(seed, prefix + ".bias", (n,), 0.05)


def image_state(seed: int = 0, d_img: int = 1024, n_disease: int = 13) -> dict:
    """ImageEncoderCNN.state_dict(): `backbone.{0,1,4..7}.*`, `proj.*`, `classifier.*`.

    backbone = Sequential(conv1, bn1, relu, maxpool, layer1..4, avgpool)
    (training_pipeline.py:178-183), hence indices 0,1 and 4..7.
    BN statistics are randomised so that a wrong fold shows up in parity.
    """
    sd: dict = {}
    _conv(sd, seed, "backbone.0.weight", 64, 3, 7)
    _bn(sd, seed, "backbone.1", 64)
    inplanes = 64
    for li, (planes, blocks, stride) in enumerate(RESNET_STAGES):
        for b in range(blocks):
            p = f"backbone.{4 + li}.{b}"
            _conv(sd, seed, p + ".conv1.weight", planes, inplanes, 1)
            _bn(sd, seed, p + ".bn1", planes)
            _conv(sd, seed, p + ".conv2.weight", planes, planes, 3)
            _bn(sd, seed, p + ".bn2", planes)
            _conv(sd, seed, p + ".conv3.weight", planes * 4, planes, 1)
            _bn(sd, seed, p + ".bn3", planes * 4, gamma=(0.2, 0.6))
            if b == 0:
                _conv(sd, seed, p + ".downsample.0.weight", planes * 4, inplanes, 1)
                _bn(sd, seed, p + ".downsample.1", planes * 4, gamma=(0.4, 0.9))
            inplanes = planes * 4
    _linear(sd, seed, "proj", d_img, 2048)
    _linear(sd, seed, "classifier", n_disease, d_img)
    return sd


def text_state(seed: int = 0, d_txt: int = 512, n_disease: int = 13, layers: int = 12,
               hidden: int = 768, ffn: int = 3072, vocab: int = 30522, max_pos: int = 512) -> dict:
    """TextEncoderTransformer.state_dict(): `encoder.*` (HF BertModel), `proj.*`, `classifier.*`."""
    sd: dict = {}
    e = "encoder.embeddings."
    sd[e + "word_embeddings.weight"] = _normal(seed, e + "word", (vocab, hidden), 0.05)
    sd[e + "position_embeddings.weight"] = _normal(seed, e + "pos", (max_pos, hidden), 0.05)
    sd[e + "token_type_embeddings.weight"] = _normal(seed, e + "type", (2, hidden), 0.05)
    _ln(sd, seed, e + "LayerNorm", hidden)
    for l in range(layers):
        p = f"encoder.encoder.layer.{l}."
        for n in ("query", "key", "value"):
            _linear(sd, seed, p + "attention.self." + n, hidden, hidden, std=0.03)
        _linear(sd, seed, p + "attention.output.dense", hidden, hidden, std=0.03)
        _ln(sd, seed, p + "attention.output.LayerNorm", hidden)
        _linear(sd, seed, p + "intermediate.dense", ffn, hidden, std=0.03)
        _linear(sd, seed, p + "output.dense", hidden, ffn, std=0.02)
        _ln(sd, seed, p + "output.LayerNorm", hidden)
    _linear(sd, seed, "encoder.pooler.dense", hidden, hidden, std=0.03)
    _linear(sd, seed, "proj", d_txt, hidden)
    _linear(sd, seed, "classifier", n_disease, d_txt)
    return sd


def fusion_state(seed: int = 0, d_img: int = 1024, d_txt: int = 512, d_fuse_hidden: int = 1024,
                 n_disease: int = 13, n_cond: int = 4, h_dec: int = 512) -> dict:
    """FusionTransformerModel.state_dict() minus `report_model.*` (T5, off the named path)."""
    sd: dict = {}
    _linear(sd, seed, "fusion_mlp.0", d_fuse_hidden, d_img + d_txt)
    _ln(sd, seed, "fusion_mlp.3", d_fuse_hidden)
    _linear(sd, seed, "disease_head", n_disease, d_fuse_hidden, std=0.06, bias_std=0.3)
    _linear(sd, seed, "cond_proj.0", n_cond * h_dec, d_fuse_hidden)
    return sd


def make_state_bundle(seed: int = 0, d_img: int = 1024, d_txt: int = 512, d_fuse_hidden: int = 1024,
                      n_disease: int = 13, bert_tok=None, version: int = 999) -> dict:
    """A bundle in the on-disk layout of `model_bundle.pt`
    (training_pipeline.py:783-791: cfg / fusion_state / image_state / text_state),
    extended with the serving keys `inference()` reads (api/views.py:247-257)."""
    return {
        "cfg": {"fusion": {"d_img": d_img, "d_txt": d_txt, "d_fuse_hidden": d_fuse_hidden,
                           "n_disease": n_disease, "n_cond_tokens": 4}},
        "image_state": image_state(seed, d_img, n_disease),
        "text_state": text_state(seed, d_txt, n_disease),
        "fusion_state": fusion_state(seed, d_img, d_txt, d_fuse_hidden, n_disease),
        "bert_tok": bert_tok,
        "t5_tok": None,
        "class_names": list(CLASS_NAMES[:n_disease]),
        "thresholds": [0.5] * n_disease,
        "version": version,
    }


# ---------------------------------------------------------------------------
# synthetic studies
# ---------------------------------------------------------------------------

def synth_images(batch: int, size: int = 224, seed: int = 1234) -> np.ndarray:
    """Chest-X-ray-shaped uint8 images [B, size, size, 3], R=G=B: a smooth
    low-frequency field (bilinear-upsampled 14x14 U[0,255]) plus N(0,8) noise.
    Study i depends only on (seed, i), so any batch is a prefix of a larger one."""
    g = 14
    pos = (np.arange(size, dtype=np.float32) + 0.5) * (g / size) - 0.5
    i0 = np.clip(np.floor(pos).astype(np.int64), 0, g - 1)
    i1 = np.clip(i0 + 1, 0, g - 1)
    f = np.clip(pos - np.floor(pos), 0, 1).astype(np.float32)
    out = np.empty((batch, size, size, 3), np.uint8)
    for b in range(batch):
        rng = np.random.Generator(np.random.PCG64([seed, b]))
        base = rng.random(size=(g, g), dtype=np.float32) * 255.0
        rows = base[i0, :] * (1 - f)[:, None] + base[i1, :] * f[:, None]
        img = rows[:, i0] * (1 - f)[None, :] + rows[:, i1] * f[None, :]
        img = img + rng.standard_normal(size=img.shape, dtype=np.float32) * 8.0
        out[b] = np.clip(np.rint(img), 0, 255).astype(np.uint8)[..., None]
    return out


def synth_token_ids(batch: int, seq_len: int = 128, seed: int = 1235, ragged: bool = False):
    """Synthetic WordPiece ids: `[CLS] ... [SEP]`, ids U{1000..30521}, token_type 0.

    Returns (input_ids int64 [B,L], attention_mask int64 [B,L]).  With `ragged`
    each row gets a valid length U{16..40} (the size the reference's
    patient-details grammar produces) and is padded with [PAD]=0 to L.
    Row i depends only on (seed, i): independent of batch size and seq_len."""
    ids = np.zeros((batch, seq_len), np.int64)
    mask = np.zeros((batch, seq_len), np.int64)
    for b in range(batch):
        rng = np.random.Generator(np.random.PCG64([seed, b]))
        n = int(rng.integers(16, 41))
        row = rng.integers(1000, 30522, size=512, dtype=np.int64)
        n = min(n, seq_len) if ragged else seq_len
        ids[b, :n] = row[:n]
        ids[b, 0] = 101
        ids[b, n - 1] = 102
        mask[b, :n] = 1
    return ids, mask


_SYMPTOMS = [
    "cough", "productive cough", "dry cough", "chronic cough", "fever", "low grade fever",
    "shortness of breath", "chest pain", "pleuritic chest pain", "leg swelling", "fatigue",
    "weight loss", "chest tightness", "malaise", "tenderness", "sudden chest pain",
    "swelling of ankles", "chest discomfort", "nighttime breathlessness",
]


def synth_details(batch: int, seed: int = 1236) -> list:
    """Patient-details strings with the shape of the reference's generator
    (backend/ml/data_prep/raw_data_pre_preparation.py:114-163):
    "<age> year old <sex> <view> view , <risk factors> , <symptoms>"."""
    rng = np.random.Generator(np.random.PCG64(seed))
    out = []
    for _ in range(batch):
        parts = [f"{int(rng.integers(18, 91))} year old {('male', 'female')[int(rng.integers(2))]}",
                 f"{('AP', 'PA')[int(rng.integers(2))]} view"]
        risk = []
        if rng.random() < 0.25:
            risk.append(f"smoking history of {int(rng.integers(5, 41))} pack years")
        if rng.random() < 0.10:
            risk.append("diabetes")
        if rng.random() < 0.30:
            risk.append("hypertension")
        if risk:
            parts.append(", " + ", ".join(risk))
        k = int(rng.integers(0, 4))
        if k:
            sel = rng.choice(len(_SYMPTOMS), size=k, replace=False)
            parts.append(", " + ", ".join(_SYMPTOMS[int(i)] for i in sel))
        else:
            parts.append(", routine evaluation")
        out.append(" ".join(parts))
    return out


def build_vocab(size: int = 30522) -> list:
    """Deterministic stand-in for `bert-base-uncased/vocab.txt` (not on the box):
    the special tokens sit at their real indices ([PAD]=0, [UNK]=100, [CLS]=101,
    [SEP]=102, [MASK]=103); words of the patient-details grammar, digits and
    characters follow from index 1000.  Token ids are an INPUT of the hot path, so
    oracle and engine see exactly the same ids."""
    vocab = [f"[unused{i}]" for i in range(size)]
    vocab[0], vocab[100], vocab[101], vocab[102], vocab[103] = "[PAD]", "[UNK]", "[CLS]", "[SEP]", "[MASK]"
    words = []
    seen = set()

    def add(w):
        if w not in seen:
            seen.add(w)
            words.append(w)

    for c in ",.;:!?()/-'\"%":
        add(c)
    for c in "abcdefghijklmnopqrstuvwxyz0123456789":
        add(c)
    for c in "abcdefghijklmnopqrstuvwxyz0123456789":
        add("##" + c)
    for n in range(0, 121):
        add(str(n))
    text = ("year old male female ap pa view smoking history of pack years diabetes hypertension "
            "routine evaluation asymptomatic screening on exertion breathlessness difficulty "
            "breathing when lying down acute localized wall pain with deep the a and " + " ".join(_SYMPTOMS))
    for w in text.split():
        add(w)
    for i, w in enumerate(words):
        vocab[1000 + i] = w
    return vocab


def make_bert_tokenizer():
    """HF `BertTokenizer` (WordPiece, lower-casing) over `build_vocab()` - the offline replacement for
    `AutoTokenizer.from_pretrained("bert-base-uncased")` (training_pipeline.py:323).  It is the checker and the non-ASCII
    fallback: the product path tokenises with the native `csrc/tokenizer.cpp` (tokenizer.py), which
    tests/test_tokenizer_cpu.py holds to this tokenizer's ids bit for bit."""
    from transformers import BertTokenizer
    return BertTokenizer(vocab={w: i for i, w in enumerate(build_vocab())}, do_lower_case=True)
