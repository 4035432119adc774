"""Small engine run for compute-sanitizer (memcheck / racecheck): one B=2 forward through the device
entry point, the host entry point (ordinary path, then captured graph + replay), the fp32 mode and three T5 decoder steps.
No oracle: the sanitizer slows kernels 10-50x.   compute-sanitizer --tool memcheck python tools/sanitize_smoke.py"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mmdx_b200 import engine, synth                    # noqa: E402
from mmdx_b200 import inference_pipeline as ip         # noqa: E402

bundle = synth.make_state_bundle(seed=0)
prec = os.environ.get("SAN_PRECISION", "bf16")
eng = ip.get_engine(bundle, "cuda:0", precision=prec)
B, L = 2, 64
imgs = synth.synth_images(B, 224, seed=1)
ids, mask = synth.synth_token_ids(B, L, seed=2, ragged=True)
pk = engine.pack_tokens(ids, mask, None, eng.table_sizes)
host = [torch.from_numpy(np.ascontiguousarray(x)).pin_memory() for x in (imgs, pk[0], pk[1], pk[2], pk[3])]
dev = [x.cuda() for x in host]
a = eng.forward(*dev, pk[4])
torch.cuda.synchronize()
outs = [eng.forward_host(*host, pk[4]) for _ in range(3)]        # ordinary, capture + replay, replay
assert all(torch.equal(o[1], a[1].cpu()) for o in outs)
if prec == "fp32":
    o = eng.forward_f32(*dev, pk[4])
    torch.cuda.synchronize()
    assert float((o["probs"] - a[1]).abs().max()) < 1e-2
if os.environ.get("SAN_T5", "1") == "1":
    from transformers import T5Config, T5ForConditionalGeneration
    from mmdx_b200.t5_fast import MmdxStep
    torch.manual_seed(0)
    t5 = T5ForConditionalGeneration(T5Config(decoder_start_token_id=0)).eval().cuda()
    st = MmdxStep(t5)
    st.begin(torch.randn(4, 4, 512, device="cuda"), 4, 8)
    tok = torch.zeros(4, dtype=torch.long, device="cuda")
    for i in range(3):
        tok = st.step(tok).argmax(-1)
        st.reorder(torch.tensor([1, 0, 2, 2]))
    torch.cuda.synchronize()
print("sanitize_smoke: ok", a[1][0, :4].tolist())
