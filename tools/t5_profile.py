"""Times the native report search (mmdx_t5_generate) on a random-init t5-small and, under ncu, lists its kernels:
  python tools/t5_profile.py [tokens] [studies]
  ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file out.csv python tools/t5_profile.py 24 1"""
import sys
import time

import torch
from transformers import T5Config, T5ForConditionalGeneration

sys.path.insert(0, ".")
from mmdx_b200.t5_fast import MmdxStep  # noqa: E402


def main():
    tokens = int(sys.argv[1]) if len(sys.argv) > 1 else 180
    studies = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    torch.manual_seed(0)
    m = T5ForConditionalGeneration(T5Config(decoder_start_token_id=0)).eval().cuda()
    step = MmdxStep(m)
    cond = torch.randn(studies, 4, 512, device="cuda")
    kw = dict(max_new_tokens=tokens, min_new_tokens=tokens, num_beams=4, no_repeat_ngram_size=3, length_penalty=1.1,
              early_stopping=True, eos_token_id=1, pad_token_id=0, decoder_start_token_id=0)
    step.generate_native(cond, **dict(kw, max_new_tokens=4, min_new_tokens=4))
    best = 1e9
    for _ in range(3):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        out = step.generate_native(cond, **kw)
        torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t0)
    print(f"native search, {studies} studies x 4 beams x {tokens} tokens: {best * 1e3:.1f} ms "
          f"({best / tokens * 1e6:.0f} us per token), output {tuple(out.shape)}")
    # the decoder step alone: 64 steps enqueued back to back (no beam search, no host round trip per token)
    rows = studies * 4
    step.begin(cond.repeat_interleave(4, 0), rows, 200)
    tok = torch.zeros(rows, dtype=torch.long, device="cuda")
    for _ in range(8):
        step.step(tok)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(64):
        step.step(tok)
    e1.record()
    torch.cuda.synchronize()
    print(f"decoder step alone, back to back: {e0.elapsed_time(e1) / 64 * 1e3:.0f} us per step (positions 8..71)")
    import os
    if os.environ.get("MMDX_T5_PROF") == "1":
        import ctypes as C
        import numpy as np
        buf = np.zeros(1024, np.uint64)
        n = C.c_int(0)
        step._check(step._lib.mmdx_t5_step_profile(step._h, C.c_void_p(buf.ctypes.data), 1024, C.byref(n)))
        ts = buf[:50].astype(np.int64)
        dt = np.diff(ts) / 1e3
        names = [f"L{l}.{p}" for l in range(6) for p in ("qkv", "self", "so", "cq", "cross", "co", "wi", "wo")] + ["lm"]   # L0.qkv includes the embedding
        print("phase times of the last step (us, CTA 0):")
        for l in range(0, len(dt), 8):
            print("  " + "  ".join(f"{nm}={v:.1f}" for nm, v in zip(names[l:l + 8], dt[l:l + 8])))
        print(f"  total {(ts[-1] - ts[0]) / 1e3:.1f} us")
    step.close()


if __name__ == "__main__":
    main()
