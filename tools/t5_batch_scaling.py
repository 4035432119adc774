import sys, time, torch
from transformers import T5Config, T5ForConditionalGeneration
sys.path.insert(0, ".")
from mmdx_b200.t5_fast import MmdxStep
torch.manual_seed(0)
m = T5ForConditionalGeneration(T5Config(decoder_start_token_id=0)).eval().cuda()
step = MmdxStep(m)
for studies in (1, 2, 4, 8, 16):
    cond = torch.randn(studies, 4, 512, device="cuda")
    kw = dict(max_new_tokens=180, min_new_tokens=150, num_beams=4, no_repeat_ngram_size=3, length_penalty=1.1,
              early_stopping=True, eos_token_id=1, pad_token_id=0, decoder_start_token_id=0)
    step.generate_native(cond, **dict(kw, max_new_tokens=8, min_new_tokens=4))
    ts = []
    for _ in range(3):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        out = step.generate_native(cond, **kw)
        ts.append(time.perf_counter() - t0)
    print(f"{studies:3d} studies x 4 beams: {min(ts)*1e3:7.1f} ms per batch, {min(ts)*1e3/studies:6.1f} ms per report", flush=True)
step.close()
