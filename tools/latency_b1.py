"""B = 1 request latency through mmdx_forward_host (the reference's real call shape: one 512x512 image, ~34 tokens; CUDA graph
replay): median / p99 over 400 calls.  A/B switches are read from the environment (e.g. MMDX_HEAD_FUSED=0)."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mmdx_b200 import engine, synth
from mmdx_b200 import inference_pipeline as ip

bundle = synth.make_state_bundle(seed=0)
eng = ip.get_engine(bundle, "cuda")
im1 = synth.synth_images(1, 512, seed=77)
ids1, mask1 = synth.synth_token_ids(1, 96, seed=78, ragged=True)
p1 = engine.pack_tokens(ids1, mask1, None, eng.table_sizes)
h1 = [torch.from_numpy(x).pin_memory() for x in (im1, p1[0], p1[1], p1[2], p1[3])]
for _ in range(20):
    eng.forward_host(h1[0], h1[1], h1[2], h1[3], h1[4], p1[4])
n0 = eng.launch_count
ts = []
for _ in range(400):
    t = time.perf_counter()
    eng.forward_host(h1[0], h1[1], h1[2], h1[3], h1[4], p1[4])
    ts.append(time.perf_counter() - t)
print(f"B=1 latency: median {1e3 * np.median(ts):.4f} ms, p99 {1e3 * np.percentile(ts, 99):.4f} ms, "
      f"{(eng.launch_count - n0) // 400} launches per request, MMDX_HEAD_FUSED={os.environ.get('MMDX_HEAD_FUSED', '1')}")
