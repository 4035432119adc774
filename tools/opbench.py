"""Single-kernel timings at the C2 sizes (B=256, L=128) through the C ABI: CUDA events on the launch stream,
L2 flushed between repetitions (a 256 MB memset).  Usage on the GPU box:
    python tools/opbench.py attention gemm conv ln            # any subset; no argument = everything
    MMDX_CG=1 python tools/opbench.py gemm                    # pin the CTA-group size
Prints one line per case: name, microseconds (median of reps), achieved TFLOP/s or GB/s."""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mmdx_b200 import _lib, engine  # noqa: E402


def P(t):
    return C.c_void_p(0 if t is None else t.data_ptr())


def S():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


_flush = None


def timeit(fn, reps=7):
    global _flush
    if _flush is None:
        _flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for _ in range(2):
        fn()
    ts = []
    for _ in range(reps):
        _flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda._sleep(400000)     # ~0.2 ms of GPU spin: hides fn()'s host-side work (tensor-map encode, launch)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    return float(np.median(ts))


def bf(x):
    return x.to(torch.bfloat16)


def bench_attention(h, lib):
    for B, L in [(256, 128), (64, 512), (256, 32)]:
        T, hid = B * L, 768
        qkv = bf(torch.randn(T, 3 * hid, device="cuda"))
        cu = torch.arange(0, T + 1, L, dtype=torch.int32, device="cuda")
        ctx = torch.empty(T, hid, device="cuda", dtype=torch.bfloat16)
        us = timeit(lambda: _lib.check(lib.mmdx_op_attention(h.handle, P(qkv), P(cu), B, T, L, 12, hid, P(ctx), S())))
        fl = 4.0 * B * 12 * L * L * 64
        by = T * hid * 2 * 4
        print(f"attention B={B} L={L}: {us:8.1f} us  {fl / us / 1e6:7.1f} TFLOP/s  {by / us / 1e3:7.1f} GB/s", flush=True)


def bench_gemm(h, lib):
    M = 32768
    cases = [("qkv", 2304, 768, 0, False, 0), ("attn_out+res", 768, 768, 0, True, 0),
             ("ffn1+gelu", 3072, 768, 2, False, 0), ("ffn1 (no act)", 3072, 768, 0, False, 0),
             ("ffn2+res", 768, 3072, 0, True, 0), ("ffn2 (no res)", 768, 3072, 0, False, 0)]
    if os.environ.get("OPBENCH_SWEEP"):        # tile-width sweep on the short-K residual GEMM
        cases += [(f"attn_out{'+res' if r else ''} bn{bn}", 768, 768, 0, r, bn) for r in (False, True) for bn in (256, 192, 128)]
    only = os.environ.get("OPBENCH_ONLY")
    for name, N, K, act, res, bn in cases:
        if only and name != only:
            continue
        a = bf(torch.randn(M, K, device="cuda"))
        w = bf(torch.randn(N, K, device="cuda") * K ** -0.5)
        bias = torch.randn(N, device="cuda")
        r = bf(torch.randn(M, N, device="cuda")) if res else None
        out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
        us = timeit(lambda: _lib.check(lib.mmdx_op_gemm(h.handle, P(a), K, P(w), P(bias), P(r), N, P(out), N, M, N, K, act, 0, bn, S())))
        print(f"gemm {name:18s} M={M} N={N} K={K}: {us:8.1f} us  {2.0 * M * N * K / us / 1e6:7.1f} TFLOP/s", flush=True)


def bench_gemm_ln(h, lib):
    """dense + bias + residual + LayerNorm: one fused launch against GEMM + layernorm_kernel."""
    M, N = 32768, 768
    for name, K in [("attn_out+res+LN", 768), ("ffn2+res+LN", 3072)]:
        a = bf(torch.randn(M, K, device="cuda"))
        w = bf(torch.randn(N, K, device="cuda") * K ** -0.5)
        bias = torch.randn(N, device="cuda")
        r = bf(torch.randn(M, N, device="cuda"))
        g = torch.ones(N, device="cuda")
        b = torch.zeros(N, device="cuda")
        pre = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
        out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)

        def two():
            _lib.check(lib.mmdx_op_gemm(h.handle, P(a), K, P(w), P(bias), P(r), N, P(pre), N, M, N, K, 0, 0, 0, S()))
            _lib.check(lib.mmdx_op_layernorm(h.handle, P(pre), M, N, P(g), P(b), 1e-12, P(out), S()))
        us2 = timeit(two)
        us1 = timeit(lambda: _lib.check(lib.mmdx_op_gemm_ln(h.handle, P(a), K, P(w), P(bias), P(r), N, P(pre), N, P(g), P(b), 1e-12,
                                                            P(out), N, M, N, K, S())))
        usi = timeit(lambda: _lib.check(lib.mmdx_op_gemm_ln(h.handle, P(a), K, P(w), P(bias), P(r), N, P(pre), N, P(g), P(b), 1e-12,
                                                            P(pre), N, M, N, K, S())))
        print(f"gemm+ln {name:16s} M={M} K={K}: two launches {us2:7.1f} us   fused {us1:7.1f} us   fused in place {usi:7.1f} us", flush=True)


def bench_conv(h, lib):
    NB = 256
    cases = [("l1.0.c1", 56, 64, 64, 1, 1, False), ("l1.0.ds", 56, 64, 256, 1, 1, False),
             ("l1.c1", 56, 256, 64, 1, 1, False), ("l1.c2", 56, 64, 64, 3, 1, False), ("l1.c3+res", 56, 64, 256, 1, 1, True),
             ("l2.0.c1", 56, 256, 128, 1, 1, False), ("l2.0.c2s2", 56, 128, 128, 3, 2, False), ("l2.0.ds", 56, 256, 512, 1, 2, False),
             ("l2.c1", 28, 512, 128, 1, 1, False), ("l2.c2", 28, 128, 128, 3, 1, False), ("l2.c3+res", 28, 128, 512, 1, 1, True),
             ("l3.0.c1", 28, 512, 256, 1, 1, False), ("l3.0.c2s2", 28, 256, 256, 3, 2, False), ("l3.0.ds", 28, 512, 1024, 1, 2, False),
             ("l3.c1", 14, 1024, 256, 1, 1, False), ("l3.c2", 14, 256, 256, 3, 1, False), ("l3.c3+res", 14, 256, 1024, 1, 1, True),
             ("l4.0.c1", 14, 1024, 512, 1, 1, False), ("l4.0.c2s2", 14, 512, 512, 3, 2, False), ("l4.0.ds", 14, 1024, 2048, 1, 2, False),
             ("l4.c1", 7, 2048, 512, 1, 1, False), ("l4.c2", 7, 512, 512, 3, 1, False), ("l4.c3+res", 7, 512, 2048, 1, 1, True)]
    only = os.environ.get("OPBENCH_ONLY")
    for name, HW, Cin, Cout, k, s, res in cases:
        if only and name != only:
            continue
        x = bf(torch.randn(NB, HW, HW, Cin, device="cuda"))
        w = bf(torch.randn(Cout, k * k, Cin, device="cuda") * (k * k * Cin) ** -0.5)
        bias = torch.randn(Cout, device="cuda")
        OH = (HW + 2 * (k // 2) - k) // s + 1
        r = bf(torch.randn(NB, OH, OH, Cout, device="cuda")) if res else None
        out = torch.empty(NB, OH, OH, Cout, device="cuda", dtype=torch.bfloat16)
        us = timeit(lambda: _lib.check(lib.mmdx_op_conv(h.handle, P(x), NB, HW, HW, Cin, P(w), P(bias), P(r), P(out), Cout, k, s, 1, S())))
        fl = 2.0 * NB * OH * OH * Cout * Cin * k * k
        by = 2.0 * (x.numel() + out.numel() * (2 if res else 1))
        print(f"conv {name:10s} {HW}x{HW} {Cin}->{Cout} k{k} s{s}: {us:8.1f} us  {fl / us / 1e6:7.1f} TFLOP/s  {by / us / 1e3:7.1f} GB/s", flush=True)


def bench_bneck(h, lib):
    """Fused layer-1 bottleneck against its three separate launches (conv2 3x3, conv3 + shortcut, next conv1)."""
    NB, HW = 256, 56
    for C1 in (64, 128, 0):
        CN = max(C1, 64)
        t1 = bf(torch.randn(NB, HW, HW, 64, device="cuda").relu())
        res = bf(torch.randn(NB, HW, HW, 256, device="cuda"))
        w2 = bf(torch.randn(64, 9, 64, device="cuda") / 24); b2 = torch.randn(64, device="cuda")
        w3 = bf(torch.randn(256, 1, 64, device="cuda") / 8); b3 = torch.randn(256, device="cuda")
        w1 = bf(torch.randn(CN, 1, 256, device="cuda") / 16); b1 = torch.randn(CN, device="cuda")
        t2 = torch.empty(NB, HW, HW, 64, device="cuda", dtype=torch.bfloat16)
        y = torch.empty(NB, HW, HW, 256, device="cuda", dtype=torch.bfloat16)
        t1n = torch.empty(NB, HW, HW, CN, device="cuda", dtype=torch.bfloat16)

        def three():
            _lib.check(lib.mmdx_op_conv(h.handle, P(t1), NB, HW, HW, 64, P(w2), P(b2), None, P(t2), 64, 3, 1, 1, S()))
            _lib.check(lib.mmdx_op_conv(h.handle, P(t2), NB, HW, HW, 64, P(w3), P(b3), P(res), P(y), 256, 1, 1, 1, S()))
            if C1:
                _lib.check(lib.mmdx_op_conv(h.handle, P(y), NB, HW, HW, 256, P(w1), P(b1), None, P(t1n), C1, 1, 1, 1, S()))
        us3 = timeit(three)
        us1 = timeit(lambda: _lib.check(lib.mmdx_op_bneck64(h.handle, P(t1), P(res), P(w2), P(b2), P(w3), P(b3), P(w1), P(b1), C1,
                                                            None, None, None, P(y), P(t1n), NB, HW, HW, S())))
        by = 2.0 * NB * HW * HW * (64 + 256 + 256 + C1)
        print(f"bottleneck 56x56 next conv1 {C1:3d}: separate launches {us3:7.1f} us   fused {us1:7.1f} us  ({by / us1 / 1e3:6.1f} GB/s algorithmic)",
              flush=True)


def bench_ln(h, lib):
    T, N = 32768, 768
    x = bf(torch.randn(T, N, device="cuda"))
    g = torch.rand(N, device="cuda") + 0.5
    b = torch.randn(N, device="cuda")
    y = torch.empty_like(x)
    us = timeit(lambda: _lib.check(lib.mmdx_op_layernorm(h.handle, P(x), T, N, P(g), P(b), 1e-12, P(y), S())))
    print(f"layernorm T={T} N={N}: {us:8.1f} us  {4.0 * T * N / us / 1e3:7.1f} GB/s", flush=True)


def bench_pre(h, lib):
    for B, HW in [(256, 224), (64, 512)]:
        img = torch.randint(0, 256, (B, HW, HW, 3), dtype=torch.uint8, device="cuda")
        hp, wp = C.c_int(), C.c_int()
        lib.mmdx_padded_dims(224, 224, C.byref(hp), C.byref(wp))
        out = torch.zeros(B, hp.value, wp.value, 4, device="cuda", dtype=torch.bfloat16)
        oh, ow = C.c_int(), C.c_int()
        us = timeit(lambda: _lib.check(lib.mmdx_op_preprocess(h.handle, P(img), B, HW, HW, 3, P(out), C.byref(oh), C.byref(ow), S())))
        by = img.numel() + B * 224 * 224 * 4 * 2
        print(f"preprocess B={B} {HW}x{HW}: {us:8.1f} us  {by / us / 1e3:7.1f} GB/s", flush=True)


def bench_stem(h, lib):
    B, HW = 256, 224
    hp, wp = C.c_int(), C.c_int()
    lib.mmdx_padded_dims(HW, HW, C.byref(hp), C.byref(wp))
    x = torch.zeros(B, hp.value, wp.value, 4, device="cuda", dtype=torch.bfloat16)
    x[:, 3:3 + HW, 3:3 + HW, :3] = torch.randn(B, HW, HW, 3, device="cuda").to(torch.bfloat16)
    w = torch.randn(64, 3, 7, 7) * 147 ** -0.5
    pk = np.zeros(7 * 4 * 8 * 8 * 8, dtype=np.uint16)
    wf = np.ascontiguousarray(w.numpy())
    lib.mmdx_pack_stem_weights(wf.ctypes.data_as(C.c_void_p), None, pk.ctypes.data_as(C.c_void_p))
    wd = torch.from_numpy(pk.view(np.int16)).cuda()
    bias = torch.randn(64, device="cuda")
    for pool in (1, 0):
        shape = (B, 56, 56, 64) if pool else (B, 112, 112, 64)
        out = torch.empty(shape, device="cuda", dtype=torch.bfloat16)
        us = timeit(lambda: _lib.check(lib.mmdx_op_stem_pool(h.handle, P(x), B, HW, HW, P(wd), P(bias), P(out), pool, S())))
        fl = 2.0 * B * 112 * 112 * 64 * 147
        by = 2.0 * (x.numel() + out.numel())
        print(f"stem pool={pool} B={B} {HW}x{HW}: {us:8.1f} us  {fl / us / 1e6:7.1f} TFLOP/s  {by / us / 1e3:7.1f} GB/s", flush=True)


ALL = {"stem": bench_stem, "attention": bench_attention, "gemm": bench_gemm, "gemm_ln": bench_gemm_ln, "bneck": bench_bneck, "conv": bench_conv, "ln": bench_ln, "pre": bench_pre}

if __name__ == "__main__":
    torch.cuda.set_device(0)
    h = engine.RawHandle()
    lib = _lib.lib()
    for name in (sys.argv[1:] or list(ALL)):
        ALL[name](h, lib)
    h.close()
