"""What the host-side stages in front of the accelerated path sustain on this box: the
GPU path takes ~29 k studies/s per B200, so JPEG decode (api/views.py:70, Pillow) and tokenisation
(training_pipeline.py:335-342, HF tokenizer) decide what a serving process can actually feed it.

  python tools/bench_host_stages.py            # prints one JSON object; results are kept under profiles/

Measures, on synthetic 512x512 chest-X-ray-shaped JPEGs (the size of backend/sample_images) and patient-detail strings:
  * Pillow decode + convert("RGB") on one core and on all cores (process pool),
  * nvJPEG through torchvision.io.decode_jpeg(device="cuda") in batches, and how many decoded bytes differ from Pillow's
    (libjpeg-turbo vs nvJPEG IDCT - the reason the parity goldens keep Pillow),
  * the BERT tokenizer on batches of 256 strings."""
import io
import json
import os
import sys
import time
from concurrent.futures import ProcessPoolExecutor

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def make_jpegs(n, hw=512, seed=0):
    from PIL import Image
    from mmdx_b200 import synth
    imgs = synth.synth_images(n, hw, seed=seed)
    out = []
    for a in imgs:
        buf = io.BytesIO()
        Image.fromarray(a).save(buf, format="JPEG", quality=90)
        out.append(buf.getvalue())
    return out


def _decode_many(blobs):
    from PIL import Image
    s = 0
    for b in blobs:
        a = np.asarray(Image.open(io.BytesIO(b)).convert("RGB"))
        s += int(a[0, 0, 0])
    return s


def main():
    res = {"cores": os.cpu_count()}
    blobs = make_jpegs(256)
    res["jpeg_bytes_avg"] = sum(map(len, blobs)) / len(blobs)
    t = time.perf_counter(); _decode_many(blobs); dt = time.perf_counter() - t
    res["pillow_decode_1core_img_s"] = len(blobs) / dt
    nproc = os.cpu_count() or 1
    chunks = [blobs[i::nproc] * 4 for i in range(nproc)]
    with ProcessPoolExecutor(nproc) as ex:
        list(ex.map(_decode_many, chunks))                      # warm the pool
        t = time.perf_counter(); list(ex.map(_decode_many, chunks)); dt = time.perf_counter() - t
    res["pillow_decode_allcores_img_s"] = sum(map(len, chunks)) / dt

    try:
        import torch
        from torchvision.io import decode_jpeg
        if torch.cuda.is_available():
            from PIL import Image
            ts = [torch.frombuffer(bytearray(b), dtype=torch.uint8) for b in blobs]
            out = decode_jpeg(ts, device="cuda")                # warm-up (creates the nvJPEG handle)
            torch.cuda.synchronize()
            t = time.perf_counter()
            for _ in range(4):
                out = decode_jpeg(ts, device="cuda")
            torch.cuda.synchronize()
            dt = time.perf_counter() - t
            res["nvjpeg_decode_img_s"] = 4 * len(ts) / dt
            ref = np.asarray(Image.open(io.BytesIO(blobs[0])).convert("RGB"))
            got = out[0].permute(1, 2, 0).cpu().numpy()
            if got.shape[2] == 1:
                got = np.repeat(got, 3, axis=2)
            d = np.abs(got.astype(int) - ref.astype(int))
            res["nvjpeg_vs_pillow"] = {"bytes_differing_frac": float((d > 0).mean()), "max_abs_diff": int(d.max())}
    except Exception as e:                                      # noqa: BLE001 - report, do not hide
        res["nvjpeg_error"] = repr(e)[:200]

    try:                                                        # the engine's own batched nvJPEG entry point
        import torch
        from mmdx_b200 import synth as _synth
        from mmdx_b200 import inference_pipeline as ip
        from mmdx_b200._lib import lib
        eng = ip.get_engine(_synth.make_state_bundle(seed=0), "cuda")
        for nb in (64, 256):
            eng.decode_jpeg_batch(blobs[:nb], 512, 512)
            torch.cuda.synchronize()
            t = time.perf_counter()
            for _ in range(4):
                eng.decode_jpeg_batch(blobs[:nb], 512, 512)
            torch.cuda.synchronize()
            res[f"mmdx_nvjpeg_batch{nb}_img_s"] = 4 * nb / (time.perf_counter() - t)
        res["mmdx_nvjpeg_backend"] = int(lib().mmdx_jpeg_backend(eng.handle))
    except Exception as e:                                      # noqa: BLE001
        res["mmdx_nvjpeg_error"] = repr(e)[:200]

    from mmdx_b200 import synth
    tok = synth.make_bert_tokenizer()
    rng = np.random.default_rng(0)
    words = ["patient", "male", "female", "age", "view", "pa", "ap", "follow", "up", "finding", "history", "of", "cough", "fever"]
    texts = [" ".join(rng.choice(words, size=int(rng.integers(12, 30)))) for _ in range(256)]
    tok(texts, padding="max_length", truncation=True, return_tensors="np", max_length=128)
    t = time.perf_counter()
    for _ in range(8):
        tok(texts, padding="max_length", truncation=True, return_tensors="np", max_length=128)
    dt = time.perf_counter() - t
    res["tokenizer_class"] = type(tok).__name__ + (" (tokenizers-backed)" if getattr(tok, "is_fast", False) else " (python)")
    res["tokenize_batch256_studies_s"] = 8 * 256 / dt
    # the native WordPiece tokenizer of libmmdx.so (csrc/tokenizer.cpp), same strings, same ids
    from mmdx_b200.tokenizer import NativeBertTokenizer
    nt = NativeBertTokenizer(tok)
    a = tok(texts, padding="max_length", truncation=True, return_tensors="np", max_length=128)
    b = nt(texts, max_length=128)
    res["native_tokenizer_ids_identical"] = bool(all((a[k] == b[k]).all() for k in a))
    details = synth.synth_details(4096, seed=2)
    for name, tx in (("batch256", texts), ("batch4096_details", details)):
        nt(tx, max_length=128)
        t = time.perf_counter()
        reps = 50 if len(tx) <= 256 else 10
        for _ in range(reps):
            nt(tx, max_length=128)
        res[f"native_tokenize_{name}_studies_s"] = reps * len(tx) / (time.perf_counter() - t)
    print(json.dumps(res))


if __name__ == "__main__":
    main()
