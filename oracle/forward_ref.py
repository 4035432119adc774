"""ORACLE - test infrastructure, not product code.

CPU restatement of the reference's batched multimodal inference forward
(`backend/ml/pipelines/inference_pipeline.py:150-206` and the three modules in
`backend/ml/pipelines/training_pipeline.py`).  Only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` legs
may import this file; the shipped engine never does.

Integer work (Pillow's 8-bit bilinear resample, crop) is restated in numpy and
is bit-exact; floating-point work is restated as plain fp32 torch functional
ops on CPU, one op per reference op, in the reference's order (no folding, no
fusion), driven directly by the reference's state_dict tensors.

PARITY PINNING: the reference has no tests and no golden vectors (SURVEY.md
section 4), so the oracle is pinned against outputs of the reference's own
modules, imported in the build container by `oracle/ref_import.py` and dumped
by `oracle/make_golden.py` into `tests/golden/` (see tests/test_oracle.py), and
against Pillow / torchvision themselves, which are installed wherever the tests
run.  The arithmetic lives in un-vendored third-party packages pinned by
`backend/requirements.txt` (pillow 11.3.0, torchvision 0.23.0, transformers
4.56.1, torch 2.8.0); the versions installed in this image (12.2.0 / 0.26.0 /
5.5.0 / 2.11.0) are the oracle of record.
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn.functional as F

IMAGENET_MEAN = (0.485, 0.456, 0.406)   # training_pipeline.py:117
IMAGENET_STD = (0.229, 0.224, 0.225)
PRECISION_BITS = 22                      # Pillow ImagingResample 8bpc fixed point


# ---------------------------------------------------------------------------
# P1-P5: image_transfom_into_tensor (training_pipeline.py:112-119)
# ---------------------------------------------------------------------------

def resize_output_size(h: int, w: int, size: int = 256):
    """T.Resize(int): shorter side -> size, longer -> int(size*long/short)
    (torchvision/transforms/functional.py `_compute_resized_output_size`)."""
    short, long = (w, h) if w <= h else (h, w)
    new_short, new_long = size, int(size * long / short)
    new_w, new_h = (new_short, new_long) if w <= h else (new_long, new_short)
    return new_h, new_w


def bilinear_coeffs(in_size: int, out_size: int):
    """Pillow `precompute_coeffs` + `normalize_coeffs_8bpc` for the triangle filter
    (support 1.0): per output index the first tap, tap count and int32 weights."""
    scale = in_size / out_size
    fs = max(scale, 1.0)
    support = 1.0 * fs
    ksize = int(math.ceil(support)) * 2 + 1
    xmin = np.zeros(out_size, np.int32)
    xcnt = np.zeros(out_size, np.int32)
    wts = np.zeros((out_size, ksize), np.int32)
    for x in range(out_size):
        center = (x + 0.5) * scale
        lo = int(center - support + 0.5)
        lo = max(lo, 0)
        hi = int(center + support + 0.5)
        hi = min(hi, in_size)
        n = hi - lo
        w = np.zeros(n, np.float64)
        for k in range(n):
            v = 1.0 - abs((k + lo - center + 0.5) * (1.0 / fs))
            w[k] = v if v > 0.0 else 0.0
        tot = w.sum()
        if tot != 0.0:
            w = w / tot
        xmin[x], xcnt[x] = lo, n
        for k in range(n):
            wts[x, k] = int(0.5 + w[k] * (1 << PRECISION_BITS)) if w[k] >= 0 else int(-0.5 + w[k] * (1 << PRECISION_BITS))
    return xmin, xcnt, wts


def _resample_axis(a: np.ndarray, axis: int, out_size: int) -> np.ndarray:
    in_size = a.shape[axis]
    if in_size == out_size:
        return a
    xmin, xcnt, wts = bilinear_coeffs(in_size, out_size)
    a = np.moveaxis(a, axis, 0).astype(np.int64)
    out = np.empty((out_size,) + a.shape[1:], np.int64)
    for x in range(out_size):
        acc = np.full(a.shape[1:], 1 << (PRECISION_BITS - 1), np.int64)
        for k in range(int(xcnt[x])):
            acc += a[xmin[x] + k] * int(wts[x, k])
        out[x] = acc >> PRECISION_BITS
    out = np.clip(out, 0, 255).astype(np.uint8)
    return np.moveaxis(out, 0, axis)


def pil_resize_bilinear_u8(img: np.ndarray, out_h: int, out_w: int) -> np.ndarray:
    """PIL.Image.resize((out_w,out_h), BILINEAR) on an 8-bit HWC image: horizontal
    pass first, uint8 intermediate, then vertical (Pillow ImagingResample)."""
    a = _resample_axis(img, 1, out_w)
    return _resample_axis(a, 0, out_h)


def center_crop_offsets(h: int, w: int, size: int = 224):
    """torchvision center_crop: top=int(round((h-size)/2.0)) (functional.py:592-594)."""
    return int(round((h - size) / 2.0)), int(round((w - size) / 2.0))


def preprocess_u8(img: np.ndarray, resize: int = 256, crop: int = 224) -> np.ndarray:
    """Resize(256) + CenterCrop(224) on an HWC uint8 image -> uint8 [crop,crop,C].
    resize=0 / crop=0 skip that stage (BASELINE config C4(i): the CNN at the image's own resolution, which the
    reference's modules allow - adaptive avgpool, training_pipeline.py:183 - with ToTensor + Normalize only)."""
    r = img
    if resize:
        h, w = img.shape[:2]
        oh, ow = resize_output_size(h, w, resize)
        r = pil_resize_bilinear_u8(img, oh, ow)
    if crop:
        top, left = center_crop_offsets(r.shape[0], r.shape[1], crop)
        r = r[top:top + crop, left:left + crop]
    return np.ascontiguousarray(r)


def preprocess_f32(img: np.ndarray, resize: int = 256, crop: int = 224) -> torch.Tensor:
    """Whole transform: u8 HWC -> fp32 CHW, /255, gray->3ch, (x-mean)/std."""
    u8 = preprocess_u8(img, resize, crop)
    x = torch.from_numpy(u8).permute(2, 0, 1).to(torch.float32).div(255)
    if x.size(0) == 1:
        x = x.repeat(3, 1, 1)
    mean = torch.tensor(IMAGENET_MEAN).view(3, 1, 1)
    std = torch.tensor(IMAGENET_STD).view(3, 1, 1)
    return (x - mean) / std


def preprocess_f32_pillow(img: np.ndarray) -> torch.Tensor:
    """The same transform through the libraries the reference itself calls (training_pipeline.py:112-119: torchvision
    Resize/CenterCrop/ToTensor/Normalize over Pillow's C resampler) - used by the timed CPU baseline so that the
    reference arm is not slowed down by this file's numpy restatement of Pillow.  Bit-identical to preprocess_f32
    (tests/test_oracle.py)."""
    import torchvision.transforms as T
    from PIL import Image
    t = T.Compose([T.Resize(256, antialias=True), T.CenterCrop(224), T.ToTensor(),
                   T.Lambda(lambda x: x.repeat(3, 1, 1) if x.size(0) == 1 else x),
                   T.Normalize(mean=list(IMAGENET_MEAN), std=list(IMAGENET_STD))])
    a = img[..., 0] if img.ndim == 3 and img.shape[2] == 1 else img
    return t(Image.fromarray(a))


# ---------------------------------------------------------------------------
# I1-I2: ImageEncoderCNN.forward (training_pipeline.py:276-311)
# ---------------------------------------------------------------------------

def _bn(x, sd, p):
    return F.batch_norm(x, sd[p + ".running_mean"], sd[p + ".running_var"], sd[p + ".weight"], sd[p + ".bias"],
                        training=False, eps=1e-5)


def resnet50_features(x: torch.Tensor, sd: dict, return_stages: bool = False):
    """torchvision ResNet-50 v1.5 minus fc, eval mode: [B,3,H,W] -> [B,2048]."""
    stages = {}
    x = F.relu(_bn(F.conv2d(x, sd["backbone.0.weight"], stride=2, padding=3), sd, "backbone.1"))
    stages["stem"] = x
    x = F.max_pool2d(x, 3, 2, 1)
    stages["pool"] = x
    for li, (blocks, stride) in enumerate(((3, 1), (4, 2), (6, 2), (3, 2))):
        for b in range(blocks):
            p = f"backbone.{4 + li}.{b}"
            s = stride if b == 0 else 1
            idt = x
            o = F.relu(_bn(F.conv2d(x, sd[p + ".conv1.weight"]), sd, p + ".bn1"))
            o = F.relu(_bn(F.conv2d(o, sd[p + ".conv2.weight"], stride=s, padding=1), sd, p + ".bn2"))
            o = _bn(F.conv2d(o, sd[p + ".conv3.weight"]), sd, p + ".bn3")
            if (p + ".downsample.0.weight") in sd:
                idt = _bn(F.conv2d(x, sd[p + ".downsample.0.weight"], stride=s), sd, p + ".downsample.1")
            x = F.relu(o + idt)
        stages[f"layer{li + 1}"] = x
    feats = F.adaptive_avg_pool2d(x, 1).flatten(1)
    return (feats, stages) if return_stages else feats


def image_encode(x: torch.Tensor, sd: dict):
    """ImageEncoderCNN.encode: backbone -> flatten -> proj (training_pipeline.py:291-302).
    The warm-up `classifier` (:309-310) is never read by inference() and is skipped."""
    feats = resnet50_features(x, sd)
    return feats, F.linear(feats, sd["proj.weight"], sd["proj.bias"])


# ---------------------------------------------------------------------------
# T1-T8: TextEncoderTransformer.forward (training_pipeline.py:452-508) over HF BertModel
# ---------------------------------------------------------------------------

def bert_last_hidden(input_ids, attention_mask, token_type_ids, sd: dict, heads: int = 12,
                     return_layers: bool = False):
    """HF BertModel forward, eval mode (transformers/models/bert/modeling_bert.py):
    embeddings word+type+position -> LN(1e-12); 12x {QKV, softmax(QK^T/8 + key mask) V,
    dense+residual+LN, dense+erf-GELU, dense+residual+LN}.  The pooler is never read."""
    e = "encoder.embeddings."
    B, L = input_ids.shape
    if token_type_ids is None:
        token_type_ids = torch.zeros_like(input_ids)
    h = sd[e + "word_embeddings.weight"][input_ids] + sd[e + "token_type_embeddings.weight"][token_type_ids] \
        + sd[e + "position_embeddings.weight"][:L].unsqueeze(0)
    h = F.layer_norm(h, (h.size(-1),), sd[e + "LayerNorm.weight"], sd[e + "LayerNorm.bias"], 1e-12)
    H = h.size(-1)
    dh = H // heads
    add_mask = torch.zeros(B, 1, 1, L, dtype=torch.float32)
    add_mask.masked_fill_(attention_mask.view(B, 1, 1, L) == 0, torch.finfo(torch.float32).min)
    layers = []
    n_layers = 1 + max(int(k.split(".")[3]) for k in sd if k.startswith("encoder.encoder.layer."))
    for l in range(n_layers):
        p = f"encoder.encoder.layer.{l}."

        def lin(x, name):
            return F.linear(x, sd[p + name + ".weight"], sd[p + name + ".bias"])

        q = lin(h, "attention.self.query").view(B, L, heads, dh).transpose(1, 2)
        k = lin(h, "attention.self.key").view(B, L, heads, dh).transpose(1, 2)
        v = lin(h, "attention.self.value").view(B, L, heads, dh).transpose(1, 2)
        s = torch.matmul(q, k.transpose(-1, -2)) * (dh ** -0.5) + add_mask
        ctx = torch.matmul(torch.softmax(s, dim=-1), v).transpose(1, 2).reshape(B, L, H)
        a = F.layer_norm(lin(ctx, "attention.output.dense") + h, (H,),
                         sd[p + "attention.output.LayerNorm.weight"], sd[p + "attention.output.LayerNorm.bias"], 1e-12)
        f = F.gelu(lin(a, "intermediate.dense"))
        h = F.layer_norm(lin(f, "output.dense") + a, (H,),
                         sd[p + "output.LayerNorm.weight"], sd[p + "output.LayerNorm.bias"], 1e-12)
        if return_layers:
            layers.append(h)
    return (h, layers) if return_layers else h


def mean_pool(last_hidden, attention_mask):
    """TextEncoderTransformer.mean_pool (training_pipeline.py:452-459)."""
    mask = attention_mask.unsqueeze(-1).type_as(last_hidden)
    return (last_hidden * mask).sum(dim=1) / mask.sum(dim=1).clamp(min=1e-6)


def text_encode(input_ids, attention_mask, token_type_ids, sd: dict):
    """encode(): BERT -> masked mean pool -> proj (training_pipeline.py:465-498)."""
    h = bert_last_hidden(input_ids, attention_mask, token_type_ids, sd)
    pooled = mean_pool(h, attention_mask)
    return pooled, F.linear(pooled, sd["proj.weight"], sd["proj.bias"])


# ---------------------------------------------------------------------------
# F1 + O1: FusionTransformerModel.forward (:584-610) and inference() post-processing
# ---------------------------------------------------------------------------

def fusion_head(z_img, z_txt, sd: dict):
    """cat -> Linear -> GELU(erf) -> Dropout(eval: identity) -> LayerNorm(1e-5) -> Linear."""
    z = torch.cat([z_img, z_txt], dim=-1)
    hdn = F.gelu(F.linear(z, sd["fusion_mlp.0.weight"], sd["fusion_mlp.0.bias"]))
    z_fuse = F.layer_norm(hdn, (hdn.size(-1),), sd["fusion_mlp.3.weight"], sd["fusion_mlp.3.bias"], 1e-5)
    return z_fuse, F.linear(z_fuse, sd["disease_head.weight"], sd["disease_head.bias"])


def cond_tokens(z_fuse, sd: dict, n_cond: int = 4):
    """FusionTransformerModel._make_encoder_outputs (training_pipeline.py:553-558, 574-578): Linear -> GELU(erf),
    viewed as [B, n_cond, h_dec] - the "encoder output" the T5 decoder cross-attends to."""
    c = F.gelu(F.linear(z_fuse, sd["cond_proj.0.weight"], sd["cond_proj.0.bias"]))
    return c.view(z_fuse.size(0), n_cond, -1)


@torch.no_grad()
def forward_batch(bundle: dict, x_img: torch.Tensor, input_ids, attention_mask, token_type_ids=None) -> dict:
    """The named path on preprocessed fp32 images and token ids; every intermediate kept."""
    feats, z_img = image_encode(x_img, bundle["image_state"])
    pooled, z_txt = text_encode(input_ids, attention_mask, token_type_ids, bundle["text_state"])
    z_fuse, logits = fusion_head(z_img, z_txt, bundle["fusion_state"])
    probs = torch.sigmoid(logits)
    thr = torch.tensor(bundle["thresholds"], dtype=torch.float32)
    vector = (probs >= thr).int()          # inference_pipeline.py:186 (>=)
    out = {"feats": feats, "z_img": z_img, "pooled": pooled, "z_txt": z_txt, "z_fuse": z_fuse,
           "logits": logits, "probs": probs, "vector": vector}
    if "cond_proj.0.weight" in bundle["fusion_state"]:
        out["cond"] = cond_tokens(z_fuse, bundle["fusion_state"])
    return out


@torch.no_grad()
def inference_batch(bundle: dict, images_u8: list, input_ids, attention_mask, token_type_ids=None,
                    resize: int = 256, crop: int = 224, pillow: bool = False) -> dict:
    """images_u8: list of HWC uint8 arrays (already decoded, as after `.convert("RGB")`).
    pillow=True: preprocessing through Pillow / torchvision themselves (the timed CPU baseline)."""
    if pillow:
        assert resize == 256 and crop == 224
        x = torch.stack([preprocess_f32_pillow(im) for im in images_u8])
    else:
        x = torch.stack([preprocess_f32(im, resize, crop) for im in images_u8])
    return forward_batch(bundle, x, input_ids, attention_mask, token_type_ids)
