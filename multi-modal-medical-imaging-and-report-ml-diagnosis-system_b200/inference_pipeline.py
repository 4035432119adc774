"""Drop-in for the reference's `backend/ml/pipelines/inference_pipeline.py` hot path.

`inference(model_bundle, image_pil, patient_details, device=None, gen_kwargs=None)` keeps the
reference's signature, argument meaning, error behaviour and result dict
(inference_pipeline.py:150-206); the work between `image_transfom_into_tensor` (:174) and the
sigmoid/threshold (:185-186) runs in hand-written sm_100a kernels through the C ABI (include/mmdx.h).
`inference_batch` is the batch extension (the reference is hard-wired to B=1: `.unsqueeze(0)` :174).

Differences a maintainer must know (also in INTEGRATION.md):
  * there is no CPU path: `device=None` means the current CUDA device (the reference means "cpu");
  * report generation (T5 beam search, :190-196) is off the accelerated path: when the bundle carries
    the reference's `fusion_model` module and `t5_tok`, `report_text` is produced by that module's own
    `generate` from our embeddings; pass `gen_kwargs=False` to skip it (`report_text` = "").
"""
from __future__ import annotations

import threading

import numpy as np
import torch

from .engine import Engine, pack_tokens

_ENGINES: dict = {}
_LOCK = threading.Lock()


def _parse_device(device):
    # same accepted types / same TypeError as inference_pipeline.py:152-159
    if isinstance(device, torch.device):
        dev = device
    elif isinstance(device, str):
        dev = torch.device(device)
    elif device is None:
        dev = torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else torch.device("cuda")
    else:
        raise TypeError(f"'device' must be str|torch.device|None, got {type(device)}")
    if dev.type != "cuda":
        raise RuntimeError("mmdx_b200 runs on CUDA (B200) only - there is no CPU fallback; got device=%r" % (device,))
    return dev


def _states_from_bundle(model_bundle: dict) -> dict:
    """Accepts the serving bundle of api/views.py:247-257 (nn.Modules) or the on-disk
    `model_bundle.pt` layout of training_pipeline.py:783-791 (state dicts)."""
    out = {}
    for prefix, mod_key, sd_key in (("image", "image_encoder", "image_state"), ("text", "text_encoder", "text_state"),
                                    ("fusion", "fusion_model", "fusion_state")):
        if model_bundle.get(mod_key) is not None:
            out[prefix] = model_bundle[mod_key].state_dict()
        elif model_bundle.get(sd_key) is not None:
            out[prefix] = model_bundle[sd_key]
        else:
            raise ValueError(f"Bundle missing '{mod_key}' / '{sd_key}'")
    return out


def _precision(model_bundle: dict, precision=None) -> str:
    p = precision or model_bundle.get("precision") or "bf16"
    if p not in ("bf16", "fp32"):
        raise ValueError("precision must be 'bf16' (tcgen05 path) or 'fp32' (parity mode)")
    return p


def get_engine(model_bundle: dict, device=None, precision=None) -> Engine:
    """Weights are packed to the device once per bundle identity, device and precision, never per call."""
    dev = _parse_device(device)
    idx = dev.index if dev.index is not None else torch.cuda.current_device()
    fp32 = _precision(model_bundle, precision) == "fp32"
    key = (id(model_bundle), idx, fp32)
    with _LOCK:
        hit = _ENGINES.get(key)
        if hit is not None and hit[1] is model_bundle:
            return hit[0]
        with torch.cuda.device(idx):
            if model_bundle.get("packed_weights"):       # serving bundle without torch modules (save_packed_bundle)
                eng = Engine.from_packed(model_bundle["packed_weights"], device=idx)
            else:
                eng = Engine(_states_from_bundle(model_bundle), device=idx, fp32=fp32)
        _ENGINES[key] = (eng, model_bundle)
        return eng


def save_packed_bundle(model_bundle: dict, path: str, device=None) -> dict:
    """Writes the bundle's weights as one packed file (kernel-ready bf16 arena, Engine.save_packed) and returns the
    light serving bundle that refers to it: {"packed_weights": path, class_names, thresholds, version, bert_tok,
    t5_tok}.  `inference()` on that bundle needs no torch modules; `report_text` is "" unless `fusion_model` is kept."""
    eng = get_engine(model_bundle, "cuda" if device is None else device)
    eng.save_packed(path)
    light = {k: model_bundle.get(k) for k in ("class_names", "thresholds", "version", "bert_tok", "t5_tok")}
    light["packed_weights"] = path
    return light


def clear_engines():
    """Counterpart of api/views.py:260-263 `clear_model_bundle`."""
    with _LOCK:
        for eng, _ in _ENGINES.values():
            eng.close()
        _ENGINES.clear()


def _to_u8(img) -> np.ndarray:
    if isinstance(img, np.ndarray):
        a = img
    elif isinstance(img, torch.Tensor):
        a = img.cpu().numpy()
    else:                                  # PIL image (api/views.py:70 hands over `.convert("RGB")`)
        a = np.asarray(img)
    if a.ndim == 2:
        a = a[..., None]
    if a.dtype != np.uint8 or a.ndim != 3 or a.shape[2] not in (1, 3):
        raise ValueError("images must be 8-bit with 1 or 3 channels (PIL mode L/RGB or uint8 HWC)")
    return np.ascontiguousarray(a)


_TOKENIZERS: dict = {}
_REPORT_GENS: dict = {}


def fast_report_generator(fusion, dev):
    """`t5_fast.FastT5Generator` on the CUDA decoder-step kernels for the bundle's report model (built once per module and
    device), or None when that model is not a ReLU T5 (then HF's own generate runs unchanged)."""
    rm = getattr(fusion, "report_model", None)
    if rm is None:
        return None
    key = (id(rm), dev.index)
    with _LOCK:
        hit = _REPORT_GENS.get(key)
        if hit is not None and hit[1] is rm:
            return hit[0]
        try:
            from .t5_fast import FastT5Generator, MmdxStep
            gen = FastT5Generator(rm, MmdxStep(rm, dev))
        except Exception:      # noqa: BLE001 - unsupported decoder: not an error, HF does the work
            gen = None
        _REPORT_GENS[key] = (gen, rm)
        return gen


def native_tokenizer(model_bundle):
    """The C++ WordPiece tokenizer for the bundle's BERT tokenizer (built once per tokenizer object), or None when that
    tokenizer is not a plain tokenizers-backed BERT WordPiece tokenizer (then HF is used as is)."""
    tok = model_bundle.get("bert_tok")
    if tok is None:
        return None
    with _LOCK:
        hit = _TOKENIZERS.get(id(tok))
        if hit is not None and hit[1] is tok:
            return hit[0]
        try:
            from .tokenizer import NativeBertTokenizer
            nt = NativeBertTokenizer(tok)
        except Exception:      # noqa: BLE001 - unsupported tokenizer type: not an error, HF does the work
            nt = None
        _TOKENIZERS[id(tok)] = (nt, tok)
        return nt


def tokenize(model_bundle, text_list, max_len=96):
    """tokenize_patient_details (training_pipeline.py:335-342) with the bundle's BERT tokenizer - through the native
    WordPiece tokenizer (csrc/tokenizer.cpp, bit-identical ids) when the tokenizer is a plain BERT one."""
    tok = model_bundle.get("bert_tok")
    if tok is None:
        raise ValueError("Bundle missing 'bert_tok'")
    nt = native_tokenizer(model_bundle)
    if nt is not None:
        return nt(list(text_list), max_length=max_len)
    return tok(list(text_list), padding="max_length", truncation=True, return_tensors="np", max_length=max_len)


@torch.no_grad()
def inference_batch(model_bundle, images, details=None, device=None, gen_kwargs=False, max_len=96, tokens=None,
                    precision=None):
    """Batched `inference`: images = list of PIL / uint8 HWC arrays (any sizes), details = list[str]
    (or `tokens` = dict with input_ids / attention_mask / token_type_ids [B,L] already tokenized).
    Returns a list of result dicts in input order.
    precision (or model_bundle["precision"]): "bf16" = the tcgen05 path (probabilities within 1e-2 of the reference),
    "fp32" = the parity mode (within 1e-5, labels exact; ~20x slower; classification only)."""
    dev = _parse_device(device)
    fp32 = _precision(model_bundle, precision) == "fp32"
    eng = get_engine(model_bundle, dev, precision)
    class_names = model_bundle["class_names"]
    imgs = [_to_u8(im) for im in images]
    B = len(imgs)
    if tokens is None:
        if details is None or len(details) != B:
            raise ValueError("need one patient_details string per image")
        tokens = tokenize(model_bundle, details, max_len)
    ids_all = np.asarray(tokens["input_ids"])
    mask_all = np.asarray(tokens["attention_mask"])
    tt_all = np.asarray(tokens["token_type_ids"]) if tokens.get("token_type_ids") is not None else np.zeros_like(ids_all)
    if ids_all.shape[0] != B:
        raise ValueError("tokens and images disagree on the batch size")
    thr = torch.tensor(model_bundle["thresholds"], dtype=torch.float32)
    if thr.numel() != eng.n_cls or len(class_names) != eng.n_cls:
        raise ValueError(f"bundle has {thr.numel()} thresholds / {len(class_names)} class names for {eng.n_cls} classes")

    probs_out = np.zeros((B, eng.n_cls), np.float32)
    vec_out = np.zeros((B, eng.n_cls), np.uint8)
    z_img_out = np.zeros((B, eng.d_img), np.float32)
    z_txt_out = np.zeros((B, eng.d_txt), np.float32)
    groups: dict = {}
    for i, a in enumerate(imgs):
        groups.setdefault(a.shape, []).append(i)
    want_report = gen_kwargs is not False and model_bundle.get("fusion_model") is not None \
        and model_bundle.get("t5_tok") is not None and not fp32
    fusion_mod = model_bundle.get("fusion_model")
    use_cond = bool(want_report and eng.cond_width and hasattr(fusion_mod, "report_model")
                    and getattr(fusion_mod, "n_cond", 0) * getattr(fusion_mod, "h_dec", 0) == eng.cond_width)
    cond_out = np.zeros((B, eng.cond_width), np.float32) if use_cond else None
    with torch.cuda.device(dev):
        thr_d = thr.to(dev)
        for shape, idxs in groups.items():
            ids, pos, tt, cu, mlen = pack_tokens(ids_all[idxs], mask_all[idxs], tt_all[idxs], eng.table_sizes)
            if fp32:
                t = [torch.from_numpy(x).to(dev) for x in (np.stack([imgs[i] for i in idxs]), ids, pos, tt, cu)]
                o = eng.forward_f32(t[0], t[1], t[2], t[3], t[4], mlen, thresholds=thr_d)
                probs_out[idxs] = o["probs"].cpu().numpy()
                vec_out[idxs] = o["vector"].cpu().numpy()
                continue
            if not want_report or use_cond:
                # one C call per shape group: H2D, forward, D2H (small groups replay a captured CUDA graph)
                host = [torch.from_numpy(x).pin_memory() for x in (np.stack([imgs[i] for i in idxs]), ids, pos, tt, cu)]
                with eng.lock:   # forward + cond_tokens hand z_fuse over inside the engine: one request at a time
                    _, probs, vec = eng.forward_host(host[0], host[1], host[2], host[3], host[4], mlen, thresholds=thr)
                    if want_report:  # cond_proj on the engine: the decoder only needs these tokens
                        cond_out[idxs] = eng.cond_tokens(len(idxs)).cpu().numpy()
                probs_out[idxs] = probs.numpy()
                vec_out[idxs] = vec.numpy()
                continue
            batch = torch.from_numpy(np.stack([imgs[i] for i in idxs])).pin_memory().to(dev, non_blocking=True)
            t = [torch.from_numpy(x).pin_memory().to(dev, non_blocking=True) for x in (ids, pos, tt, cu)]
            with eng.lock:       # the three staged calls meet in the engine's shared head buffers
                _, z_img = eng.image_encode(batch, want_feats=False)
                _, z_txt = eng.text_encode(t[0], t[1], t[2], t[3], mlen, want_pooled=False)
                _, _, probs, vec = eng.head(len(idxs), thr_d, want_z_fuse=False)
                probs_out[idxs] = probs.cpu().numpy()
                vec_out[idxs] = vec.cpu().numpy()
                z_img_out[idxs] = z_img.cpu().numpy()
                z_txt_out[idxs] = z_txt.cpu().numpy()

    reports = [""] * B
    if want_report:
        # inference_pipeline.py:190-196, unchanged, on the reference's own module (off the accelerated path)
        t5_tok = model_bundle["t5_tok"]
        fusion = model_bundle["fusion_model"].to(dev)
        gen_attributes = dict(max_new_tokens=180, min_new_tokens=150, num_beams=4, no_repeat_ngram_size=3,
                              length_penalty=1.1, early_stopping=True, eos_token_id=t5_tok.eos_token_id,
                              pad_token_id=t5_tok.pad_token_id)
        if gen_kwargs:
            gen_attributes.update(gen_kwargs)
        if use_cond:
            # FusionTransformerModel.generate (training_pipeline.py:613-618) minus the two steps the engine has already
            # done (fusion_mlp, cond_proj): the T5 decoder is driven with the conditioning tokens directly
            from transformers.modeling_outputs import BaseModelOutput
            cond = torch.from_numpy(cond_out).to(dev).view(B, fusion.n_cond, fusion.h_dec)
            cond = cond.to(next(fusion.report_model.parameters()).dtype)
            fast = fast_report_generator(fusion, dev) if model_bundle.get("fast_report", True) else None
            if fast is not None and gen_attributes.get("max_new_tokens"):
                # The KV-cached CUDA decoder step (csrc/t5_decoder.cu) under the report search.  Default:
                # the whole beam search in one C call (mmdx_t5_generate - HF's algorithm restated, token-identical in the
                # tests, ~20x faster than eager HF) when the generation arguments are the ones it restates, else HF's own
                # beam search with only the model call replaced (same tokens by construction, ~3x).  fast_report = "hf"
                # forces the latter, "native" the former (raising on arguments it does not implement).
                from .t5_fast import NativeBeamSearch
                mode = model_bundle.get("fast_report", True)
                native = mode == "native" or (mode is True and NativeBeamSearch.supports(gen_attributes))
                with eng.lock:
                    if native:
                        gen_ids = NativeBeamSearch(fast.backend, fusion.report_model.config).generate(cond, **gen_attributes)
                    else:
                        gen_ids = fast.generate(cond, **gen_attributes)
            else:
                gen_ids = fusion.report_model.generate(encoder_outputs=BaseModelOutput(last_hidden_state=cond), **gen_attributes)
        else:
            gen_ids = fusion.generate(torch.from_numpy(z_img_out).to(dev), torch.from_numpy(z_txt_out).to(dev),
                                      **gen_attributes)
        reports = t5_tok.batch_decode(gen_ids, skip_special_tokens=True)

    return [{
        "report_text": reports[i],
        "disease_probs": {class_names[j]: float(probs_out[i, j]) for j in range(len(class_names))},
        "disease_vector": [int(v) for v in vec_out[i]],
        "model_version": model_bundle["version"],
    } for i in range(B)]


@torch.no_grad()
def inference_batch_jpeg(model_bundle, jpeg_blobs, details=None, device=None, max_len=96, tokens=None):
    """Batched inference straight from JPEG bytes (what api/views.py:67-70 receives): the images are decoded on the GPU
    by nvJPEG and never exist on the host.  Classification only (report_text ""), same result dicts as
    inference_batch.  nvJPEG is not bit-identical to Pillow (about 2 % of the bytes differ by one), so probabilities
    move in the 4th decimal; use inference_batch with PIL images where bit parity with the reference matters."""
    import io
    from PIL import Image
    dev = _parse_device(device)
    eng = get_engine(model_bundle, dev)
    class_names = model_bundle["class_names"]
    B = len(jpeg_blobs)
    if tokens is None:
        if details is None or len(details) != B:
            raise ValueError("need one patient_details string per image")
        tokens = tokenize(model_bundle, details, max_len)
    ids_all = np.asarray(tokens["input_ids"])
    mask_all = np.asarray(tokens["attention_mask"])
    tt_all = np.asarray(tokens["token_type_ids"]) if tokens.get("token_type_ids") is not None else np.zeros_like(ids_all)
    if ids_all.shape[0] != B:
        raise ValueError("tokens and images disagree on the batch size")
    groups: dict = {}
    for i, b in enumerate(jpeg_blobs):
        w, h = Image.open(io.BytesIO(b)).size                    # header only
        groups.setdefault((h, w), []).append(i)
    probs_out = np.zeros((B, eng.n_cls), np.float32)
    vec_out = np.zeros((B, eng.n_cls), np.uint8)
    with torch.cuda.device(dev):
        thr_d = torch.tensor(model_bundle["thresholds"], dtype=torch.float32).to(dev)
        for (h, w), idxs in groups.items():
            imgs = eng.decode_jpeg_batch([jpeg_blobs[i] for i in idxs], h, w)
            ids, pos, tt, cu, mlen = pack_tokens(ids_all[idxs], mask_all[idxs], tt_all[idxs], eng.table_sizes)
            t = [torch.from_numpy(x).pin_memory().to(dev, non_blocking=True) for x in (ids, pos, tt, cu)]
            _, probs, vec = eng.forward(imgs, t[0], t[1], t[2], t[3], mlen, thresholds=thr_d)
            probs_out[idxs] = probs.cpu().numpy()
            vec_out[idxs] = vec.cpu().numpy()
    return [{
        "report_text": "",
        "disease_probs": {class_names[j]: float(probs_out[i, j]) for j in range(len(class_names))},
        "disease_vector": [int(v) for v in vec_out[i]],
        "model_version": model_bundle["version"],
    } for i in range(B)]


@torch.no_grad()
def inference(model_bundle, image_pil, patient_details, device=None, gen_kwargs=None):
    """Same contract as the reference's inference() (inference_pipeline.py:150-206):
    returns {report_text, disease_probs{class: float}, disease_vector[0/1], model_version}.
    model_bundle["precision"] = "fp32" selects the parity mode (see inference_batch)."""
    _parse_device(device)   # TypeError before any work, like the reference
    return inference_batch(model_bundle, [image_pil], [patient_details], device=device, gen_kwargs=gen_kwargs,
                           max_len=96)[0]
