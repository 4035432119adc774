"""Scratch: which sequences of a text batch come back NaN / wrong with the folded LayerNorm path."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mmdx_b200 import engine, synth
from mmdx_b200 import inference_pipeline as ip
from oracle import forward_ref as R

bundle = synth.make_state_bundle(seed=0)
eng = ip.get_engine(bundle, "cuda")
def run(B, L, short=None, tag=""):
    ids, mask = synth.synth_token_ids(B, L, seed=72, ragged=False)
    if short is not None:
        mask[1, short:] = 0; ids[1, short - 1] = 102; ids[1, short:] = 0
    pi, pp, pt, cu, mlen = engine.pack_tokens(ids, mask)
    t = [torch.from_numpy(x).cuda() for x in (pi, pp, pt, cu)]
    pooled, z = eng.text_encode(t[0], t[1], t[2], t[3], mlen)
    torch.cuda.synchronize()
    p = pooled.cpu().numpy()
    bad = [int(i) for i in np.nonzero(~np.isfinite(p).all(1))[0]]
    print(f"{tag} B={B} L={L} short={short} T={len(pi)} nan_rows={bad} absmax={np.nanmax(np.abs(p)):.3f}", flush=True)
for (B, L, short) in ((8, 512, 300), (8, 512, None), (4, 512, None), (2, 512, None), (8, 256, None), (16, 128, None), (31, 128, None), (8, 384, None)):
    run(B, L, short, os.environ.get("TAG", ""))
