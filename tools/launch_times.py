"""Per-launch device times of one C2 step (B=256, L=128), in launch order, from the engine's event-bracketed profile
mode (branches serialised).  Text branch pattern with the folded LayerNorm: embed, then per layer qkv, attention,
attention-out, ffn1, ffn2.   python tools/launch_times.py [B] [L]"""
import ctypes as C, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mmdx_b200 import engine, synth
from mmdx_b200 import inference_pipeline as ip
from mmdx_b200._lib import lib

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
L = int(sys.argv[2]) if len(sys.argv) > 2 else 128
CLS = ["pre", "stem", "pool", "conv", "tgemm", "attn", "ln", "head", "misc"]
bundle = synth.make_state_bundle(seed=0)
eng = ip.get_engine(bundle, "cuda")
imgs = synth.synth_images(B, 224, seed=1)
ids, mask = synth.synth_token_ids(B, L, seed=2, ragged=False)
pk = engine.pack_tokens(ids, mask)
d = [torch.from_numpy(np.ascontiguousarray(x)).cuda() for x in (imgs, pk[0], pk[1], pk[2], pk[3])]
for _ in range(3):
    eng.forward(*d, pk[4])
torch.cuda.synchronize()
reps = 5
acc = None
for _ in range(reps):
    lib().mmdx_profile_begin(eng.handle)
    eng.forward(*d, pk[4])
    ms = (C.c_float * 1024)(); cl = (C.c_int32 * 1024)()
    n = lib().mmdx_profile_end_list(eng.handle, ms, cl, 1024)
    a = np.array(ms[:n]); c = np.array(cl[:n])
    acc = a if acc is None else np.minimum(acc, a)
txt = [i for i in range(n) if CLS[c[i]] in ("tgemm", "attn", "ln")]
print("launches", n, "sum_ms", round(float(acc.sum()), 3))
tg = [(CLS[c[i]], float(acc[i]) * 1e3) for i in txt]
print("text branch, first launches (us):", [(k, round(v, 1)) for k, v in tg[:16]])
# per-role averages over the 12 layers
seq = [k for k, _ in tg]
per = 7 if seq.count("ln") > 12 else 5
body = tg[1:1 + 12 * per]
names = ["qkv", "attn", "ao", "ln1", "ff1", "ff2", "ln2"] if per == 7 else ["qkv", "attn", "ao", "ff1", "ff2"]
for j, nm in enumerate(names):
    v = [body[l * per + j][1] for l in range(12)]
    print(f"{nm:5s} avg {np.mean(v):7.1f} us  min {np.min(v):7.1f}  max {np.max(v):7.1f}")
print("layer total us", round(sum(v for _, v in body) / 12, 1))
conv = [float(acc[i]) * 1e3 for i in range(n) if CLS[c[i]] == "conv"]
print("conv launches (us):", [round(v, 1) for v in conv])
