#!/usr/bin/env python
"""bench.py - studies/sec of the batched multimodal inference forward on B200.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
  (N>1: launched by `python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...`)

A "step" is one pass of the hot path over one batch of synthetic studies: BASELINE.json configs[1]
(224x224 chest-X-ray-shaped images + 128-token reports, batch 256 per GPU, bf16).  Prints ONE JSON line.

  value     studies/s, device-resident inputs (uint8 images + packed int32 ids already in HBM)
  e2e       studies/s through the public host-buffer call (pinned host -> H2D -> forward -> D2H each step)
  roofline  the dominant kernels (gemm_tcgen05_kernel + the fused layer-1 bottleneck kernel: all convs but conv1, every Linear) against the measured
            bf16 peak; kernel_rooflines = the same arithmetic for every other kernel class (tensor or HBM bound)
  cpu_baseline  the CPU oracle port (oracle/forward_ref.py) on this box's host cores, bounded sample

`--impl reference` times the reference's algorithm on the host CPU (the oracle port: the reference is
Python and /root/reference does not exist on the GPU box) for the same metric/config.
"""
import argparse
import datetime
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np
import torch

BATCH_PER_GPU = 256
SEQ_LEN = 128
IMG = 224
N_INPUT_SETS = 4           # 4 x 38.5 MB of distinct uint8 images rotate through the steps (> 126 MB L2)
CLASSES = ["preprocess", "stem_conv", "pooling", "bottleneck_convs", "text_gemms", "attention", "layernorm_embed",
           "head", "other"]


STEM_FLOPS = 2 * 64 * 3 * 49 * 112 * 112          # conv1 7x7/2 @224: 236,027,904 FLOP per study (stem_pool_tcgen05_kernel)


def flops_per_study(L=SEQ_LEN):
    """ResNet-50 convs @224 + img proj + BERT(L) + txt proj + fusion + head (2*MAC)."""
    gemm_kernel = (8_174_272_512 - STEM_FLOPS) + 4_194_304 + 169_869_312 * L + 786_432 + 3_145_728   # gemm_tcgen05_kernel
    attention = 36_864 * L * L                                    # attention_*_tcgen05_kernel, all 12 layers
    head13 = 26_624
    return gemm_kernel, STEM_FLOPS, attention, head13


def gemm_bytes_per_step(B, L, hidden=768, ffn=3072, layers=12, d_img=1024, d_txt=512, d_fuse=1024):
    """Algorithmic bytes one step moves through gemm_tcgen05_kernel launches (bf16 operands: A read once, weights once,
    residual once, output written once) - the denominator `roofline.traffic` (ncu DRAM bytes) is compared with."""
    T = B * L
    text = layers * (T * hidden * 2 * (1 + 3)            # QKV: A in, 3H out
                     + T * hidden * 2 * 3                # attention output: A, residual, out
                     + T * (hidden + ffn) * 2            # FFN1
                     + T * (ffn + 2 * hidden) * 2        # FFN2: A, residual, out
                     + (4 * hidden * hidden + 2 * hidden * ffn) * 2)
    img, cin, hw = 0, 64, 56
    for li, (mid, cout, blocks, stride) in enumerate(((64, 256, 3, 1), (128, 512, 4, 2), (256, 1024, 6, 2), (512, 2048, 3, 2))):
        for b in range(blocks):
            s = stride if b == 0 else 1
            ho = hw // s
            if li == 0:
                # layer 1: conv1 of block 0, then one bneck64_tcgen05_kernel per block: t1 in, shortcut in (block 0: the
                # 64-channel block input - its downsample conv runs in the kernel), y out, the next block's conv1 out
                nxt = 64 if b + 1 < blocks else 128
                if b == 0:
                    img += B * hw * hw * (cin + mid) * 2 + cin * mid * 2
                img += B * hw * hw * (mid + (cin if b == 0 else cout) + cout + nxt) * 2 + (9 * mid * mid + mid * cout + cout * nxt) * 2
                if b == 0:
                    img += cin * cout * 2
            else:
                if not (li == 1 and b == 0):                                                        # layer2.0 conv1: done above
                    img += B * hw * hw * (cin + mid) * 2 + cin * mid * 2                            # conv1 1x1
                img += B * (hw * hw + ho * ho) * mid * 2 + 9 * mid * mid * 2                        # conv2 3x3 (stride s)
                if b == 0:      # conv3 + downsample as one GEMM over [t2 | x strided]
                    img += B * ho * ho * (mid + cin + cout) * 2 + (mid + cin) * cout * 2
                else:           # conv3 1x1 + residual
                    img += B * ho * ho * (mid + 2 * cout) * 2 + mid * cout * 2
            cin, hw = cout, ho
    head = B * (2048 + d_img + hidden + d_txt + d_img + d_txt) * 2 + B * d_fuse * 4 \
        + (2048 * d_img + hidden * d_txt + (d_img + d_txt) * d_fuse) * 2
    return text + img + head


def measured_traffic():
    """DRAM bytes per launch of the dominant kernel from the committed ncu pass (profiles/r01_traffic.json), or None."""
    try:
        with open(os.path.join(ROOT, "profiles", "r01_traffic.json")) as f:
            t = json.load(f)
        ks = [t[k] for k in ("gemm_tcgen05_kernel", "bneck64_tcgen05_kernel") if k in t]
        n = sum(k["launches_per_step"] for k in ks)
        return sum(k["dram_bytes_read"] + k["dram_bytes_write"] for k in ks) / n, t["gemm_tcgen05_kernel"].get("source")
    except (OSError, KeyError, ValueError):
        return None, None


def class_rooflines(by_class, B, L, peaks, hidden=768, layers=12):
    """Algorithmic FLOPs or bytes per step of every non-GEMM kernel class / its device time, against the measured
    peak that bounds it (DESIGN.md section 4 lists the per-unit figures)."""
    T = B * L
    work = {
        # uint8 HWC in, bf16 NHWC4 out
        "preprocess": ("hbm", B * (IMG * IMG * 3 + IMG * IMG * 4 * 2)),
        # conv1+bn+relu+maxpool fused: padded bf16 image in, pooled 56x56x64 out; tensor-bound by FLOPs
        "stem_conv": ("tensor", B * STEM_FLOPS),
        # global avgpool (7x7x2048 bf16 in) + masked mean pool (T x 768 bf16 in)
        "pooling": ("hbm", B * 49 * 2048 * 2 + T * hidden * 2),
        "attention": ("tensor", B * 36_864 * L * L),
        # 24 LayerNorms (read + write T x 768 bf16) + embedding gather/LN (3 table rows read, 1 row written per token)
        "layernorm_embed": ("hbm", 2 * layers * 2 * T * hidden * 2 + 4 * T * hidden * 2),
    }
    out = {}
    for name, (bound, amount) in work.items():
        if name not in by_class or by_class[name]["ms_per_step"] <= 0:
            continue
        sec = by_class[name]["ms_per_step"] * 1e-3
        if bound == "tensor":
            ach, peak, unit = amount / sec / 1e12, peaks["bf16_sustained"], "TFLOP/s"
        else:
            ach, peak, unit = amount / sec / 1e9, peaks["hbm"], "GB/s"
        out[name] = {"bound": bound, "achieved": ach, "peak": peak, "unit": unit, "frac": ach / peak}
    if "attention" in out:      # attention also moves qkv in / ctx out once per layer: report the HBM view as well
        sec = by_class["attention"]["ms_per_step"] * 1e-3
        out["attention"]["hbm_gbs"] = layers * T * hidden * 2 * 4 / sec / 1e9
    return out


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"bf16_sustained": d.get("bf16_tflops_sustained", 1389.5), "bf16_burst": d.get("bf16_tflops", 1670.0),
                "hbm": d.get("hbm_gbs", 6551.0), "source": "measured"}
    return {"bf16_sustained": 1400.0, "bf16_burst": 1590.0, "hbm": 6650.0, "source": "fallback"}


class ClockSampler:
    """nvidia-smi samples of SM clock / throttle reasons during the timed region."""
    FIELDS = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                                       "-lms", "50", "-i", str(gpu_index)], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self, t0, t1):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.p is None:
            return out
        time.sleep(0.12)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = open(self.f.name).read().strip().splitlines()
        os.unlink(self.f.name)
        clocks, reasons, mx, power = [], set(), None, []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            c = [x.strip() for x in r.split(",")]
            if len(c) < 8:
                continue
            try:
                ts = datetime.datetime.strptime(c[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
                if ts < t0 - 0.05 or ts > t1 + 0.05:
                    continue
                clocks.append(float(c[1])); mx = float(c[2]); power.append(float(c[3]))
            except ValueError:
                continue
            for n, v in zip(names, c[4:8]):
                if v == "Active":
                    reasons.add(n)
        if clocks:
            out.update(sm_mhz=float(np.median(clocks)), sm_max_mhz=mx, reasons=sorted(reasons), samples=len(clocks),
                       power_w_max=max(power))
        return out


def cpu_forward_sample(n_studies, passes, threads):
    """The CPU oracle port on `n_studies` studies of the bench workload; best of `passes` (after one warm-up)."""
    from mmdx_b200 import synth
    from oracle import forward_ref as R
    torch.set_num_threads(threads)
    bundle = synth.make_state_bundle(seed=0)
    imgs = synth.synth_images(n_studies, IMG, seed=1234)
    ids, mask = synth.synth_token_ids(n_studies, SEQ_LEN, seed=1235, ragged=False)
    ids_t, mask_t = torch.from_numpy(ids), torch.from_numpy(mask)
    best = float("inf")
    for i in range(passes + 1):
        t = time.perf_counter()
        R.inference_batch(bundle, list(imgs), ids_t, mask_t)
        dt = time.perf_counter() - t
        if i > 0:
            best = min(best, dt)
    return n_studies / best, best


def run_reference(args, rank):
    """Reference arm: the reference's own CPU algorithm (oracle port) on all host threads."""
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    n = 16
    from mmdx_b200 import synth
    from oracle import forward_ref as R
    torch.set_num_threads(threads)
    bundle = synth.make_state_bundle(seed=0)
    imgs = synth.synth_images(n, IMG, seed=1234)
    ids, mask = synth.synth_token_ids(n, SEQ_LEN, seed=1235, ragged=False)
    ids_t, mask_t = torch.from_numpy(ids), torch.from_numpy(mask)
    steps = max(1, min(args.steps, 10))
    warm = max(1, min(args.warmup, 2))
    for _ in range(warm):
        R.inference_batch(bundle, list(imgs), ids_t, mask_t)
    t = time.perf_counter()
    for _ in range(steps):
        R.inference_batch(bundle, list(imgs), ids_t, mask_t)
    dt = time.perf_counter() - t
    v = n * steps / dt
    sample = f"{n} studies/step of the C2 workload (224x224 + {SEQ_LEN} tokens), fp32, torch CPU, {steps} steps"
    print(json.dumps({
        "impl": "reference", "metric": "studies/sec (image+report) batched inference", "value": v, "unit": "studies/s",
        "n_gpus": args.gpus, "steps": steps, "warmup": warm, "ms_per_step": 1e3 * dt / steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"C2: synthetic 224x224 chest X-ray + {SEQ_LEN}-token report, bounded sample of {n} studies per step",
                   "note": "reference algorithm on host CPU (oracle port of backend/ml inference forward; random-init weights)"},
        "cpu_baseline": {"value": v, "unit": "studies/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "studies/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--batch", type=int, default=BATCH_PER_GPU)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--profile-only", action="store_true", help="one warm pass + few steps, for ncu")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    args.warmup = max(args.warmup, 3)

    import torch.distributed as dist
    from mmdx_b200 import engine, synth
    from mmdx_b200 import inference_pipeline as ip
    from mmdx_b200._lib import lib
    import ctypes as C

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    B, L = args.batch, SEQ_LEN
    bundle = synth.make_state_bundle(seed=0)
    eng = ip.get_engine(bundle, dev)

    # ---- synthetic inputs: N_INPUT_SETS distinct batches per rank, host-pinned and device-resident copies
    host_sets, dev_sets = [], []
    for s in range(N_INPUT_SETS):
        imgs = synth.synth_images(B, IMG, seed=1234 + 1000 * s + 17 * rank)
        ids, mask = synth.synth_token_ids(B, L, seed=1235 + 1000 * s + 17 * rank, ragged=False)
        pi, pp, pt, cu, mlen = engine.pack_tokens(ids, mask)
        hs = [torch.from_numpy(x).pin_memory() for x in (imgs, pi, pp, pt, cu)]
        host_sets.append(hs)
        dev_sets.append([x.to(dev) for x in hs])
    h2d = sum(x.numel() * x.element_size() for x in host_sets[0])
    d2h = B * eng.n_cls * (4 + 4 + 1)
    host_out = (torch.empty(B, eng.n_cls, dtype=torch.float32).pin_memory(),
                torch.empty(B, eng.n_cls, dtype=torch.float32).pin_memory(),
                torch.empty(B, eng.n_cls, dtype=torch.uint8).pin_memory())
    gathered = torch.empty(world * B, eng.n_cls, dtype=torch.float32, device=dev) if world > 1 else None

    def step_device(i):
        d = dev_sets[i % N_INPUT_SETS]
        logits, probs, vec = eng.forward(d[0], d[1], d[2], d[3], d[4], mlen)
        if world > 1:
            dist.all_gather_into_tensor(gathered, logits)      # the path's only collective: [B,13] logits
        return logits

    def step_host(i):
        h = host_sets[i % N_INPUT_SETS]
        logits, probs, vec = eng.forward_host(h[0], h[1], h[2], h[3], h[4], mlen, out=host_out)
        if world > 1:
            dist.all_gather_into_tensor(gathered, logits.to(dev, non_blocking=True))
        return logits

    # secondary e2e figure: the host-buffer call as a two-deep request pipeline (mmdx_forward_host_submit / _wait): every
    # step still copies its inputs from pinned host memory and its results back; the copy of step i+1 runs under step i
    host_outs = [tuple(torch.empty_like(t).pin_memory() for t in host_out) for _ in range(2)]

    def step_host_pipelined(i):
        h = host_sets[i % N_INPUT_SETS]
        eng.forward_host_submit(i % 2, h[0], h[1], h[2], h[3], h[4], mlen, host_outs[i % 2])
        if i >= 1:
            eng.forward_host_wait((i - 1) % 2)
            if world > 1:
                dist.all_gather_into_tensor(gathered, host_outs[(i - 1) % 2][0].to(dev, non_blocking=True))

    def drain_host_pipeline():
        eng.forward_host_wait(0)
        eng.forward_host_wait(1)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def timed(fn, steps, warmup, sample_clocks=False, drain=None):
        for i in range(warmup):
            fn(i)
        if drain:
            drain()
        barrier()
        sampler = ClockSampler(local_rank) if sample_clocks else None
        if sampler:
            time.sleep(0.25)
        n0 = eng.launch_count
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.time()
        e0.record()
        for i in range(steps):
            fn(i)
        if drain:
            drain()                # every request of the timed region has delivered its results to host memory
        e1.record()
        barrier()
        t1 = time.time()
        ms = e0.elapsed_time(e1)
        launches = eng.launch_count - n0
        clocks = sampler.stop(t0, t1) if sampler else None
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, launches, clocks

    if args.profile_only:
        timed(step_device, args.steps, args.warmup)
        return

    ms_dev, launches, clocks = timed(step_device, args.steps, args.warmup, sample_clocks=True)
    ms_e2e, _, _ = timed(step_host, args.steps, 3)
    ms_e2e_pipe, _, _ = timed(step_host_pipelined, args.steps, 3, drain=drain_host_pipeline)
    value = world * B * args.steps / (ms_dev * 1e-3)
    e2e = world * B * args.steps / (ms_e2e * 1e-3)

    # ---- per-kernel-class device time (CUDA events around every launch) for the roofline of the dominant kernel
    prof_steps = 3
    step_device(0); torch.cuda.synchronize()
    lib().mmdx_profile_begin(eng.handle)
    for i in range(prof_steps):
        step_device(i)
    ms_cls = (C.c_float * 16)(); n_cls = (C.c_int64 * 16)()
    lib().mmdx_profile_end(eng.handle, ms_cls, n_cls, 16)
    by_class = {CLASSES[i]: {"ms_per_step": ms_cls[i] / prof_steps, "launches_per_step": n_cls[i] // prof_steps}
                for i in range(len(CLASSES)) if n_cls[i]}
    gemm_classes = ("bottleneck_convs", "text_gemms", "head")
    gemm_ms = sum(by_class[c]["ms_per_step"] for c in gemm_classes if c in by_class)
    gemm_launches = sum(by_class[c]["launches_per_step"] for c in gemm_classes if c in by_class) - 1   # head_tail is not a GEMM
    f_gemm, f_stem, f_attn, f_head = flops_per_study(L)
    peaks = load_peaks()
    achieved = f_gemm * B / (gemm_ms * 1e-3) / 1e12
    traffic, traffic_src = measured_traffic() if B == BATCH_PER_GPU else (None, None)
    roofline = {"bound": "tensor", "kernel": "gemm_tcgen05_kernel (39 bottleneck convs + every Linear layer) + bneck64_tcgen05_kernel (3 fused layer-1 bottlenecks)",
                "achieved": achieved, "peak": peaks["bf16_sustained"], "unit": "TFLOP/s",
                "frac": achieved / peaks["bf16_sustained"], "peak_source": peaks["source"] + " sustained bf16",
                "traffic": traffic, "traffic_unit": "bytes of DRAM read+write per launch (ncu, B=256)",
                "traffic_source": traffic_src,
                "algorithmic_bytes_per_launch": gemm_bytes_per_step(B, L) / max(gemm_launches, 1),
                "launches_per_step": gemm_launches,
                "avg_launch_ms": gemm_ms / max(gemm_launches, 1),
                "algorithmic_flops_per_launch": f_gemm * B / max(gemm_launches, 1),
                "whole_path_frac_of_tensor_roofline": (sum(flops_per_study(L)) * value / world) / (peaks["bf16_sustained"] * 1e12)}

    out = {
        "metric": "studies/sec (image+report) batched inference", "value": value, "unit": "studies/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_dev / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": f"C2: synthetic {IMG}x{IMG} chest X-ray + {L}-token report, batch {B} per GPU, bf16, "
                               "random-init ResNet-50 + BERT-base + fusion head",
                   "global_batch": world * B, "seq_len": L, "parallelism": f"dp{world} (batch-sharded, NCCL all_gather of logits)",
                   "l2": f"inputs rotate over {N_INPUT_SETS} distinct batches ({N_INPUT_SETS * h2d / 1e6:.0f} MB > 126 MB L2); "
                         "activations (GBs per step) sweep L2 between steps"},
        "e2e": {"value": e2e, "unit": "studies/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": ms_e2e / args.steps,
                "api": "mmdx_forward_host (synchronous: H2D, forward, D2H, wait - one call per step)",
                "pipelined": {"value": world * B * args.steps / (ms_e2e_pipe * 1e-3), "ms_per_step": ms_e2e_pipe / args.steps,
                              "api": "mmdx_forward_host_submit/_wait, two requests in flight (H2D of step i+1 under the "
                                     "kernels of step i); no faster: the step is power-bound, not idle-bound"}},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": roofline,
        "kernel_classes": by_class,
        "kernel_rooflines": class_rooflines(by_class, B, L, peaks),
    }
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        n = 32
        v, secs = cpu_forward_sample(n, 2, threads)
        out["cpu_baseline"] = {"value": v, "unit": "studies/s", "cores": threads, "kind": "port",
                               "sample": f"{n} studies of the same workload, fp32 torch CPU oracle, best of 2 passes ({secs:.1f} s each)"}
    if rank == 0:
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
