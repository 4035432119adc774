"""Library-kernel baseline on the same GPU: the reference architecture (torchvision ResNet-50 minus fc + proj,
HF BertModel + masked mean pool + proj, fusion MLP + 13-way head; training_pipeline.py:178-189, 358-367, 534-542)
run by stock PyTorch (cuDNN / cuBLAS / SDPA) in bf16, channels_last, random-init weights, on the C2 workload
(B=256, 224x224 already-normalised images, 128 all-valid tokens).  SURVEY.md section 8(d): "the stock-PyTorch GPU
path as the library kernel to beat".  Preprocessing is NOT included in this arm (it would run on the host in the
reference), so the number flatters the baseline.  Usage on the GPU box:  python tools/bench_stock_gpu.py [--compile]
Prints one JSON line; results are kept under profiles/."""
import argparse
import json

import torch
import torch.nn as nn


def run(batch=256, seq=128, steps=20, warmup=5, device="cuda"):
    """Times the stock-PyTorch arm and returns its result dict (bench.py embeds it as `gpu_library_baseline`)."""
    a = argparse.Namespace(batch=batch, seq=seq, steps=steps, warmup=warmup)
    import torchvision
    from transformers import BertConfig, BertModel
    torch.manual_seed(0)
    dev = torch.device(device)
    cnn = torchvision.models.resnet50(weights=None)
    cnn.fc = nn.Linear(2048, 1024)                       # backbone + proj (d_img = 1024)
    bert = BertModel(BertConfig(), add_pooling_layer=False)
    proj_t = nn.Linear(768, 512)
    fuse = nn.Sequential(nn.Linear(1536, 1024), nn.GELU(), nn.Dropout(0.1), nn.LayerNorm(1024))
    head = nn.Linear(1024, 13)
    mods = [cnn, bert, proj_t, fuse, head]
    for m in mods:
        m.eval().to(dev, dtype=torch.bfloat16)
    cnn.to(memory_format=torch.channels_last)
    B, L = a.batch, a.seq
    x = torch.randn(B, 3, 224, 224, device=dev, dtype=torch.bfloat16).contiguous(memory_format=torch.channels_last)
    ids = torch.randint(1000, 30522, (B, L), device=dev)
    mask = torch.ones(B, L, device=dev, dtype=torch.long)

    @torch.no_grad()
    def step():
        z_img = cnn(x)
        h = bert(input_ids=ids, attention_mask=mask).last_hidden_state
        m = mask.unsqueeze(-1).to(h.dtype)
        z_txt = proj_t((h * m).sum(1) / m.sum(1).clamp(min=1e-6))
        return torch.sigmoid(head(fuse(torch.cat([z_img, z_txt], -1))).float())

    for _ in range(a.warmup):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        p = step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.steps
    out = {"impl": "stock PyTorch bf16 (cuDNN/cuBLAS/SDPA), eager, channels_last", "torch": torch.__version__,
           "batch": B, "seq_len": L, "steps": a.steps, "ms_per_step": ms, "studies_per_s": B / ms * 1e3,
           "finite": bool(torch.isfinite(p).all()),
           "note": "preprocessing and H2D not included (they run on the host in the reference): flatters this arm"}
    del mods, cnn, bert, x
    torch.cuda.empty_cache()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--seq", type=int, default=128)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    a = ap.parse_args()
    print(json.dumps(run(a.batch, a.seq, a.steps, a.warmup)))


if __name__ == "__main__":
    main()
