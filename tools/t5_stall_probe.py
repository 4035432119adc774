"""Repeats mmdx_t5_generate at a few shapes and prints every call's wall time (looking for intermittent stalls)."""
import sys
import time

import torch
from transformers import T5Config, T5ForConditionalGeneration

sys.path.insert(0, ".")
from mmdx_b200.t5_fast import MmdxStep  # noqa: E402

torch.manual_seed(0)
m = T5ForConditionalGeneration(T5Config(decoder_start_token_id=0)).eval().cuda()
step = MmdxStep(m)
for studies, tokens, nmin in ((3, 32, 0), (3, 4, 0), (3, 32, 0), (2, 180, 150), (3, 32, 0), (1, 180, 150), (3, 32, 0)):
    cond = torch.randn(studies, 4, 512, device="cuda")
    kw = dict(max_new_tokens=tokens, min_new_tokens=nmin, num_beams=4, no_repeat_ngram_size=3, length_penalty=1.1,
              early_stopping=True, eos_token_id=1, pad_token_id=0, decoder_start_token_id=0)
    ts = []
    for _ in range(6):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        out = step.generate_native(cond, **kw)
        ts.append((time.perf_counter() - t0) * 1e3)
    print(studies, tokens, tuple(out.shape), " ".join(f"{t:.1f}" for t in ts), flush=True)
step.close()
