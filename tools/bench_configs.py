"""BASELINE.json configs[3] and configs[4] on one B200 (configs[1] is bench.py's line; configs[2] is bench.py --gpus N):

  C4  high-res stress: 512x512 images + 512-token reports, batch 64
        (i)  CNN run AT 512x512 (engine with resize_short=0, crop=0: preprocessing = normalise only) - 139.35 GFLOP/study
        (ii) the reference's faithful transform (Resize 256 + CenterCrop 224) on 512x512 inputs - 104.81 GFLOP/study
  C5  image-only branch (K_pre + stem + 52 convs + avgpool + proj), batch 1 ... 1024: latency and images/s

Timing: CUDA events on the launch stream, 3 warm-up passes, then every point is timed for the SAME duration (>= 1 s, at least
`reps` passes): the board is power-limited, a 40 ms burst runs at higher clocks than a 90 ms one, and round 1's sweep - a
fixed number of passes per point - showed that as a throughput drop from B = 512 to B = 1024.
Inputs rotate over distinct device buffers larger than L2 where the batch is small.  Prints one JSON object."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mmdx_b200 import engine, synth  # noqa: E402
from mmdx_b200 import inference_pipeline as ip  # noqa: E402

PEAK_TF = 1389.5
if os.path.exists(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")):
    PEAK_TF = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                          "MEASURED_PEAKS.json"))).get("bf16_tflops_sustained", PEAK_TF)


MIN_SECONDS = 1.0


def timed(fn, reps, warm=3):
    for i in range(warm):
        fn(i)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(3):
        fn(i)
    b.record()
    torch.cuda.synchronize()
    reps = max(reps, int(MIN_SECONDS * 1e3 / max(a.elapsed_time(b) / 3, 1e-3)))
    a.record()
    for i in range(reps):
        fn(i)
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def c4(states):
    out = {}
    B, L, HW = 64, 512, 512
    ids, mask = synth.synth_token_ids(B, L, seed=77, ragged=False)
    pi, pp, pt, cu, mlen = engine.pack_tokens(ids, mask)
    tok = [torch.from_numpy(x).cuda() for x in (pi, pp, pt, cu)]
    sets = [torch.from_numpy(synth.synth_images(B, HW, seed=100 + s)).cuda() for s in range(3)]    # 3 x 50 MB
    bert = 169_869_312 * L + 36_864 * L * L + 786_432 + 3_145_728 + 26_624
    for name, kw, cnn in (("C4i_cnn_at_512", dict(resize_short=0, crop=0), 42_706_403_328 + 4_194_304),
                          ("C4ii_faithful_resize_224", dict(), 8_174_272_512 + 4_194_304)):
        eng = engine.Engine(states, **kw)
        ms = timed(lambda i: eng.forward(sets[i % 3], tok[0], tok[1], tok[2], tok[3], mlen), 10)
        fl = (cnn + bert) * B
        out[name] = {"batch": B, "seq_len": L, "image": HW, "ms_per_batch": ms, "studies_per_s": B / ms * 1e3,
                     "gflop_per_study": (cnn + bert) / 1e9, "tflops": fl / ms / 1e9,
                     "frac_of_measured_sustained_bf16": fl / ms / 1e9 / PEAK_TF}
        eng.close()
    return out


def c5(states):
    eng = engine.Engine(states)
    rows = []
    for B in (1, 2, 4, 8, 16, 32, 64, 128, 256, 512, 1024):
        nset = max(2, min(8, (160 << 20) // (B * 224 * 224 * 3) + 1))
        sets = [torch.from_numpy(synth.synth_images(B, 224, seed=500 + s)).cuda() for s in range(nset)]
        ms = timed(lambda i: eng.image_encode(sets[i % nset], want_feats=False), 20 if B <= 256 else 8)
        rows.append({"batch": B, "ms": ms, "images_per_s": B / ms * 1e3,
                     "frac_of_tensor_roofline": 8_178_466_816 * B / ms / 1e9 / PEAK_TF})
    eng.close()
    return rows


if __name__ == "__main__":
    torch.cuda.set_device(0)
    states = ip._states_from_bundle(synth.make_state_bundle(seed=0))
    which = sys.argv[1:] or ["c4", "c5"]
    res = {}
    if "c4" in which:
        res["C4"] = c4(states)
    if "c5" in which:
        res["C5_image_only_sweep"] = c5(states)
    print(json.dumps(res, indent=1))
