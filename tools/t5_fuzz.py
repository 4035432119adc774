"""Random shapes of the decoder step against the fp32 torch step (oracle/t5_step_ref.TorchStep): studies 1-6, beams 1-4,
1-16 conditioning tokens, 4-24 positions, random beam reorders inside every study.  Prints the worst relative logit error.
Checker only (imports oracle/): python tools/t5_fuzz.py [cases] [seed]"""
import sys
import torch
from transformers import T5Config, T5ForConditionalGeneration

sys.path.insert(0, ".")
from mmdx_b200.t5_fast import MmdxStep  # noqa: E402
from oracle.t5_step_ref import TorchStep  # noqa: E402

cases = int(sys.argv[1]) if len(sys.argv) > 1 else 40
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 0
torch.manual_seed(0)
m = T5ForConditionalGeneration(T5Config(decoder_start_token_id=0)).eval().cuda()
g = torch.Generator().manual_seed(seed)
a, b = TorchStep(m), MmdxStep(m)
worst = 0.0
for c in range(cases):
    studies = int(torch.randint(1, 7, (1,), generator=g))
    beams = int(torch.randint(1, 5, (1,), generator=g))
    n_enc = [1, 2, 3, 4, 4, 4, 5, 7, 9, 16][int(torch.randint(0, 10, (1,), generator=g))]
    steps = int(torch.randint(4, 25, (1,), generator=g))
    R = studies * beams
    cond = torch.randn(studies, n_enc, 512, generator=g).cuda().repeat_interleave(beams, 0)
    a.begin(cond, R, steps)
    b.begin(cond, R, steps)
    tok = torch.randint(0, 32128, (R,), generator=g).cuda()
    err_c = 0.0
    for t in range(steps):
        la, lb = a.step(tok), b.step(tok)
        torch.cuda.synchronize()
        err_c = max(err_c, float((la - lb).abs().max() / la.abs().max()))
        tok = la.topk(3, -1).indices[torch.arange(R), torch.randint(0, 3, (R,), generator=g)]
        if beams > 1 and t < steps - 1 and int(torch.randint(0, 3, (1,), generator=g)) == 0:
            idx = torch.cat([s0 * beams + torch.randint(0, beams, (beams,), generator=g) for s0 in range(studies)]).cuda()
            a.reorder(idx)
            b.reorder(idx)
            tok = tok[idx]
    worst = max(worst, err_c)
    print(f"case {c:3d}: {studies} studies x {beams} beams, n_enc {n_enc:2d}, {steps:2d} steps: max rel err {err_c:.2e}", flush=True)
    assert err_c < 2e-5, "mismatch"
b.close()
print(f"worst relative logit error over {cases} cases: {worst:.2e}")
