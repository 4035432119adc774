"""ORACLE - test infrastructure, build-container only.

Generates tests/golden/*.npz by running the REFERENCE's own modules
(/root/reference/backend, imported through oracle/ref_import.py) on
  G1: backend/sample_images/e{1,2}.jpg + backend/sample_details/patient_details.json,
      B=1 each, L=96 (inference_pipeline.py:174-186), seed-0 weights; each study also goes through the
      reference's real entry point `inference()` (inference_pipeline.py:150-206, T5 beam search stubbed: off the
      named path) and its `disease_probs` / `disease_vector` are stored as inf_probs / inf_vector;
  G2: seeded synthetic studies, B=8, L in {96,128}, full and ragged valid lengths;
  G3: Pillow/torchvision Resize(256)+CenterCrop(224) uint8 known answers for
      512x512, 224x224, 300x400 and 1024x768 inputs (crc32 of the cropped bytes).
Run:  python -m oracle.make_golden      (about a minute on 8 cores)
"""
import json
import os
import zlib

import numpy as np
import torch
from PIL import Image

from mmdx_b200 import synth
from oracle import ref_import

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
REF = "/root/reference/backend"


@torch.no_grad()
def run_reference(tp, rb, pil_images, tok):
    """The steps of inference() :174-186 with the reference's modules, batched, keeping intermediates."""
    x = torch.stack([tp.image_transfom_into_tensor(p) for p in pil_images])
    ie, te, fm = rb["image_encoder"], rb["text_encoder"], rb["fusion_model"]
    feats = ie.backbone(x).flatten(1)
    z_img = ie(x)["embeddings"]
    enc = te.encoder(input_ids=tok["input_ids"], attention_mask=tok["attention_mask"],
                     token_type_ids=tok["token_type_ids"], return_dict=True)
    pooled = te.mean_pool(enc.last_hidden_state, tok["attention_mask"])
    z_txt = te(**tok)["embeddings"]
    out = fm(z_img=z_img, z_txt=z_txt, report_labels=None)
    logits = out["disease_logits"]
    probs = torch.sigmoid(logits)
    vector = (probs >= torch.tensor(rb["thresholds"])).int()
    pre_u8 = np.stack([np.asarray(T_crop(tp, p)) for p in pil_images])
    # conditioning tokens of the report decoder from the reference's own helper (training_pipeline.py:574-578)
    cond = fm._make_encoder_outputs(out["z_fuse"]).last_hidden_state
    return {"pre_u8": pre_u8, "x_checksum": np.float64(x.double().sum().item()),
            "feats": feats.numpy(), "z_img": z_img.numpy(), "pooled": pooled.numpy(), "z_txt": z_txt.numpy(),
            "z_fuse": out["z_fuse"].numpy(), "logits": logits.numpy(), "probs": probs.numpy(),
            "vector": vector.numpy().astype(np.uint8), "cond": cond.numpy()}


class _StubT5Tok:
    """`inference()` only reads eos/pad ids from the T5 tokenizer and decodes what `generate` returned."""
    eos_token_id, pad_token_id = 1, 0

    def batch_decode(self, ids, skip_special_tokens=True):
        return ["" for _ in ids]


def run_reference_entry_point(ip, rb, pil, details):
    """The reference's real `inference()` (inference_pipeline.py:150-206) on one study.  Report generation (T5 beam
    search, :190-196) is off the named path: `fusion_model.generate` is replaced by a stub for the duration of the call;
    everything else - device ladder, transform, tokenizer call, the three modules, sigmoid, `[0]`, the threshold tensor,
    `>=`, the result dict - is the reference's own code."""
    fm = rb["fusion_model"]
    b = dict(rb)
    b["t5_tok"] = _StubT5Tok()
    fm.generate = lambda z_img, z_txt, **kw: torch.zeros(1, 1, dtype=torch.long)
    try:
        return ip.inference(b, pil, details)
    finally:
        del fm.generate


def T_crop(tp, pil):
    """First two stages of the reference's own transform object (Resize, CenterCrop) -> PIL u8."""
    t = tp.image_transfom_into_tensor.transforms
    return t[1](t[0](pil))


def main():
    os.makedirs(OUT, exist_ok=True)
    torch.manual_seed(0)
    tp, ip = ref_import.import_reference()
    sb = synth.make_state_bundle(seed=0)
    rb = ref_import.build_reference_bundle(sb)

    # ---- G1: the reference's two sample studies, one at a time (B=1) like inference()
    det = json.load(open(f"{REF}/sample_details/patient_details.json"))
    names = sorted(det)
    grays, rows, texts = [], [], []
    for n in names:
        pil = Image.open(f"{REF}/sample_images/{n}").convert("RGB")        # api/views.py:70
        a = np.asarray(pil)
        assert (a[..., 0] == a[..., 1]).all() and (a[..., 0] == a[..., 2]).all()
        grays.append(a[..., 0].copy())
        tok = tp.tokenize_patient_details([det[n]], max_len=96)             # inference_pipeline.py:175
        r = run_reference(tp, rb, [pil], tok)
        res = run_reference_entry_point(ip, rb, pil, det[n])
        assert list(res["disease_probs"]) == list(rb["class_names"]) and res["model_version"] == rb["version"]
        inf_probs = np.array([[res["disease_probs"][c] for c in rb["class_names"]]], np.float32)
        inf_vector = np.array([res["disease_vector"]], np.uint8)
        # the staged run above (which also yields the intermediates) and the entry point must agree exactly
        assert np.array_equal(inf_probs, r["probs"]) and np.array_equal(inf_vector, r["vector"]), n
        r.update(input_ids=tok["input_ids"].numpy(), attention_mask=tok["attention_mask"].numpy(),
                 inf_probs=inf_probs, inf_vector=inf_vector)
        rows.append(r)
        texts.append(det[n])
    g1 = {k: np.concatenate([r[k] for r in rows]) if rows[0][k].ndim else np.array([r[k] for r in rows])
          for k in rows[0]}
    np.savez_compressed(os.path.join(OUT, "g1_samples.npz"), gray=np.stack(grays), names=np.array(names),
                        details=np.array(texts), **g1)
    print("G1", {k: v.shape for k, v in g1.items()}, g1["vector"].tolist())

    # ---- G2: synthetic batches
    for L, ragged in ((96, True), (128, False), (128, True)):
        B = 8
        imgs = synth.synth_images(B, 224, seed=1234)
        ids, mask = synth.synth_token_ids(B, L, seed=1235, ragged=ragged)
        tok = {"input_ids": torch.from_numpy(ids), "attention_mask": torch.from_numpy(mask),
               "token_type_ids": torch.zeros(B, L, dtype=torch.long)}
        r = run_reference(tp, rb, [Image.fromarray(im) for im in imgs], tok)
        r.pop("pre_u8_full", None)
        r["pre_crc"] = np.array([zlib.crc32(p.tobytes()) for p in r.pop("pre_u8")], dtype=np.uint64)
        tag = f"g2_B{B}_L{L}_{'ragged' if ragged else 'full'}"
        np.savez_compressed(os.path.join(OUT, tag + ".npz"), **r)
        print(tag, r["probs"][0].round(3).tolist())

    # ---- G3: resize+crop known answers straight from Pillow/torchvision
    kat = {}
    for (h, w) in ((512, 512), (224, 224), (300, 400), (1024, 768), (257, 640)):
        rng = np.random.Generator(np.random.PCG64([77, h, w]))
        a = rng.integers(0, 256, size=(h, w, 3), dtype=np.uint8)
        out = np.asarray(T_crop(tp, Image.fromarray(a)))
        kat[f"{h}x{w}"] = {"crc32": zlib.crc32(out.tobytes()), "sum": int(out.sum()), "seed": [77, h, w]}
    json.dump(kat, open(os.path.join(OUT, "g3_resize_kat.json"), "w"), indent=1)
    print("G3", kat)


if __name__ == "__main__":
    main()
