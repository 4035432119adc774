"""Multi-GPU results checked ON GPUs (the N>1 path was only ever exercised under gloo with the oracle as
compute): two processes, one per GPU, NCCL; `dist.inference_batch_sharded` with the engine as `run_local`.  Every rank
must end up with the whole batch's logits / probabilities / label vectors, its own block bit-identical to what it
computed, and the whole batch equal (within bf16 batch-composition noise) to one GPU running the unsharded batch.
Needs two GPUs: skipped on the driver's 1-GPU test box, run with `gpurun --gpus 2` (log under profiles/)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

from conftest import ROOT

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, n, q):
    try:
        sys.path.insert(0, ROOT)
        os.environ["MASTER_ADDR"] = "127.0.0.1"
        os.environ["MASTER_PORT"] = str(port)
        import torch.distributed as dist
        torch.cuda.set_device(rank)
        dev = torch.device("cuda", rank)
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
        from mmdx_b200 import dist as mdist
        from mmdx_b200 import engine, synth
        from mmdx_b200 import inference_pipeline as ip
        bundle = synth.make_state_bundle(seed=0)
        eng = ip.get_engine(bundle, dev)
        imgs = synth.synth_images(n, 224, seed=2024)
        ids, mask = synth.synth_token_ids(n, 128, seed=2025, ragged=True)

        def run(images, toks):
            pi, pp, pt, cu, mlen = engine.pack_tokens(toks["input_ids"], toks["attention_mask"])
            t = [torch.from_numpy(x).to(dev) for x in (np.ascontiguousarray(images), pi, pp, pt, cu)]
            logits, probs, vec = eng.forward(t[0], t[1], t[2], t[3], t[4], mlen)
            return {"logits": logits, "probs": probs, "vector": vec}

        got = mdist.inference_batch_sharded(run, imgs, {"input_ids": ids, "attention_mask": mask})
        lo, hi = mdist.shard_range(n, world, rank)
        mine = run(imgs[lo:hi], {"input_ids": ids[lo:hi], "attention_mask": mask[lo:hi]})
        full = run(imgs, {"input_ids": ids, "attention_mask": mask})
        torch.cuda.synchronize()
        ok = all(got[k].shape[0] == n for k in got)
        ok = ok and all(torch.equal(got[k][lo:hi], mine[k]) for k in got)                 # own block: bit-identical
        ok = ok and float((got["probs"] - full["probs"]).abs().max()) < 4e-3             # whole batch vs one GPU
        decided = (full["probs"] - 0.5).abs() > 4e-3
        ok = ok and bool((got["vector"][decided] == full["vector"][decided]).all())
        # every rank holds the same gathered tensor
        chk = got["logits"].double().sum().reshape(1)
        lst = [torch.zeros_like(chk) for _ in range(world)]
        dist.all_gather(lst, chk)
        ok = ok and all(float(x) == float(lst[0]) for x in lst)
        q.put((rank, bool(ok), ""))
        dist.destroy_process_group()
    except Exception as ex:      # noqa: BLE001
        q.put((rank, False, repr(ex)))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs (NCCL refuses two ranks on one device)")
@pytest.mark.parametrize("n", [32, 13])
def test_sharded_inference_nccl(n):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + (os.getpid() % 2000) + n
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=600) for _ in procs]
    for p in procs:
        p.join(120)
    assert sorted(r[:2] for r in res) == [(0, True), (1, True)], res
