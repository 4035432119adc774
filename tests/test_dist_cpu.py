"""world_size-2 gloo test of the batch sharding + logits gather used at N>1 GPUs (CPU, oracle as the per-rank
compute): the gathered result must equal the unsharded one, for even and ragged splits."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT
from mmdx_b200 import dist as mdist


def test_shard_range_partitions():
    for n in (1, 2, 7, 8, 256, 257):
        for world in (1, 2, 3, 8):
            spans = [mdist.shard_range(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, n, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    from mmdx_b200 import synth
    from oracle import forward_ref as R
    sd = synth.fusion_state(seed=0)
    rng = np.random.Generator(np.random.PCG64(3))
    z_img = torch.from_numpy(rng.standard_normal((n, 1024), dtype=np.float32))
    z_txt = torch.from_numpy(rng.standard_normal((n, 512), dtype=np.float32))

    def run_local(imgs, toks):
        _, logits = R.fusion_head(imgs, toks["z_txt"], sd)
        return {"logits": logits, "vector": (torch.sigmoid(logits) >= 0.5).to(torch.uint8)}

    got = mdist.inference_batch_sharded(run_local, z_img, {"z_txt": z_txt})
    _, full = R.fusion_head(z_img, z_txt, sd)
    ok = torch.allclose(got["logits"], full, atol=1e-5) and got["logits"].shape == full.shape \
        and torch.equal(got["vector"], (torch.sigmoid(full) >= 0.5).to(torch.uint8))
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


@pytest.mark.parametrize("n", [8, 7])
def test_sharded_gather_equals_unsharded_gloo(n):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000) + n
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(60)
    assert sorted(res) == [(0, True), (1, True)]
