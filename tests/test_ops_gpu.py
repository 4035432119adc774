"""Single-kernel parity tests (B200 only): every hot kernel is called through the C ABI
(`mmdx_op_*`, include/mmdx.h) and compared with a plain fp32 torch evaluation of the same op on the
same bf16-rounded inputs, or - for the integer resample - bit-exactly with the numpy oracle / Pillow."""
import ctypes as C

import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

from mmdx_b200 import _lib, engine            # noqa: E402
from oracle import forward_ref as R           # noqa: E402


@pytest.fixture(scope="module")
def h():
    hd = engine.RawHandle()
    yield hd
    hd.close()


def P(t):
    return C.c_void_p(0 if t is None else t.data_ptr())


def S():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def rel_err(got, ref):
    return float((got.float() - ref.float()).abs().max() / ref.float().abs().max().clamp_min(1e-6))


def bf(x):
    return x.to(torch.bfloat16)


# ---------------------------------------------------------------------------------------- GEMM
@pytest.mark.parametrize("M,N,K,act,res,f32,bn", [
    (128, 64, 64, 0, False, False, 64),
    (128, 128, 64, 0, False, False, 128),
    (128, 256, 128, 0, False, False, 256),
    (256, 128, 256, 0, False, False, 0),
    (1000, 768, 768, 0, True, False, 0),
    (4096, 2304, 768, 0, False, False, 256),
    (300, 3072, 768, 2, False, False, 0),
    (513, 768, 3072, 0, True, False, 128),
    (20000, 768, 768, 0, True, False, 256),
    (20000, 768, 768, 0, True, False, 192),
    (32768, 768, 3072, 0, True, False, 0),
    (4099, 1536, 128, 1, False, False, 192),
    (1, 1024, 2048, 0, False, False, 0),
    (5, 512, 768, 0, False, True, 0),
    (256, 1024, 1536, 2, False, True, 0),
    (777, 64, 64, 1, False, False, 0),
])
def test_gemm_tcgen05(h, M, N, K, act, res, f32, bn):
    g = torch.Generator(device="cuda").manual_seed(M * 7 + N + K)
    a = bf(torch.randn(M, K, device="cuda", generator=g))
    w = bf(torch.randn(N, K, device="cuda", generator=g) * (K ** -0.5))
    bias = torch.randn(N, device="cuda", generator=g)
    r = bf(torch.randn(M, N, device="cuda", generator=g)) if res else None
    out = torch.full((M, N), float("nan"), device="cuda", dtype=torch.float32 if f32 else torch.bfloat16)
    _lib.check(_lib.lib().mmdx_op_gemm(h.handle, P(a), K, P(w), P(bias), P(r), N, P(out), N, M, N, K, act, int(f32), bn, S()))
    torch.cuda.synchronize()
    ref = a.float() @ w.float().t() + bias
    if res:
        ref = ref + r.float()
    if act == 1:
        ref = F.relu(ref)
    elif act == 2:
        ref = F.gelu(ref)
    assert torch.isfinite(out.float()).all()
    assert rel_err(out, ref) < (2e-5 if f32 else 1e-2) + 2e-3


def test_gelu_epilogue_matches_erf_form(h):
    """The epilogue's branch-free GELU against torch's erf GELU, isolated with an identity weight
    (fp32 output, so only the activation's own error is seen): |err| < 1e-6 absolute."""
    M, K = 4096, 64
    x = bf(torch.linspace(-9, 9, M * K, device="cuda").view(M, K))
    w = bf(torch.eye(K, device="cuda"))
    out = torch.empty(M, K, device="cuda", dtype=torch.float32)
    _lib.check(_lib.lib().mmdx_op_gemm(h.handle, P(x), K, P(w), None, None, 0, P(out), K, M, K, K, 2, 1, 0, S()))
    torch.cuda.synchronize()
    ref = F.gelu(x.double()).float()
    assert float((out - ref).abs().max()) < 1e-6
    ob = torch.empty(M, K, device="cuda", dtype=torch.bfloat16)      # TMA epilogue path, bf16 output
    _lib.check(_lib.lib().mmdx_op_gemm(h.handle, P(x), K, P(w), None, None, 0, P(ob), K, M, K, K, 2, 0, 0, S()))
    torch.cuda.synchronize()
    assert torch.equal(ob, out.to(torch.bfloat16))


@pytest.mark.parametrize("M,K,inplace", [(256, 768, False), (1000, 768, False), (128 * 150 + 77, 768, True), (4096, 3072, True)])
def test_gemm_fused_layernorm(h, M, K, inplace):
    """BertSelfOutput / BertOutput as one launch (dense + bias + residual, then LayerNorm eps 1e-12 of the rows the
    CTA pair has just stored): pre-LN rows equal the plain GEMM's bit for bit, LN output within bf16 rounding of
    fp32 LayerNorm over those rows.  Ragged M, more items than CTA pairs, in-place output."""
    N = 768
    g = torch.Generator(device="cuda").manual_seed(M + K)
    a = bf(torch.randn(M, K, device="cuda", generator=g))
    w = bf(torch.randn(N, K, device="cuda", generator=g) * (K ** -0.5))
    bias = torch.randn(N, device="cuda", generator=g)
    res = bf(torch.randn(M, N, device="cuda", generator=g))
    gamma = 1.0 + 0.1 * torch.randn(N, device="cuda", generator=g)
    beta = 0.1 * torch.randn(N, device="cuda", generator=g)
    pre_ref = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    _lib.check(_lib.lib().mmdx_op_gemm(h.handle, P(a), K, P(w), P(bias), P(res), N, P(pre_ref), N, M, N, K, 0, 0, 256, S()))
    pre = torch.full((M, N), float("nan"), device="cuda", dtype=torch.bfloat16)
    out = pre if inplace else torch.full((M, N), float("nan"), device="cuda", dtype=torch.bfloat16)
    _lib.check(_lib.lib().mmdx_op_gemm_ln(h.handle, P(a), K, P(w), P(bias), P(res), N, P(pre), N, P(gamma), P(beta), 1e-12,
                                          P(out), N, M, N, K, S()))
    torch.cuda.synchronize()
    if not inplace:
        assert torch.equal(pre, pre_ref)
    ref = F.layer_norm(pre_ref.float(), (N,), gamma, beta, 1e-12)
    assert torch.isfinite(out.float()).all()
    assert float((out.float() - ref).abs().max()) < 0.04          # bf16 rounding of values up to ~5
    assert rel_err(out, ref) < 4e-3


@pytest.mark.parametrize("NB,H,W,C1,DS", [(3, 16, 8, 64, False), (2, 56, 56, 64, False), (1, 20, 24, 0, False),
                                            (3, 56, 56, 128, False), (40, 56, 56, 64, False), (3, 56, 56, 64, True),
                                            (33, 24, 40, 64, True), (33, 56, 56, 128, False)])
def test_bneck64_fused_block(h, NB, H, W, C1, DS):
    """Fused layer-1 bottleneck (conv2 3x3 -> conv3 1x1 + shortcut + ReLU -> next conv1 1x1 + ReLU, intermediates
    on chip; DS: the shortcut is the downsample conv of the 64-channel block input, computed in the kernel) against
    fp32 torch convs with the same bf16 rounding points as the unfused kernels.  Odd tile counts (the pair's second
    CTA runs past the end), partial tiles, several items per CTA pair."""
    g = torch.Generator(device="cuda").manual_seed(NB * 1000 + H + W + C1)
    rn = lambda *sh: torch.randn(*sh, device="cuda", generator=g)
    t1 = bf(rn(NB, H, W, 64).relu())
    res = bf(rn(NB, H, W, 256))
    x0 = bf(rn(NB, H, W, 64).relu())
    wd = bf(rn(256, 64) * (64 ** -0.5)); bd = rn(256) * 0.5
    w2 = bf(rn(64, 64, 3, 3) * (576 ** -0.5)); b2 = rn(64) * 0.5
    w3 = bf(rn(256, 64) * (64 ** -0.5)); b3 = rn(256) * 0.5
    CN = max(C1, 64)
    w1 = bf(rn(CN, 256) * (256 ** -0.5)); b1 = rn(CN) * 0.5
    w2p = w2.permute(0, 2, 3, 1).contiguous()                       # [Cout][kh][kw][Cin]
    y = torch.full((NB, H, W, 256), float("nan"), device="cuda", dtype=torch.bfloat16)
    t1n = torch.full((NB, H, W, CN), float("nan"), device="cuda", dtype=torch.bfloat16)
    _lib.check(_lib.lib().mmdx_op_bneck64(h.handle, P(t1), None if DS else P(res), P(w2p), P(b2), P(w3), P(b3), P(w1), P(b1), C1,
                                          P(x0) if DS else None, P(wd) if DS else None, P(bd) if DS else None, P(y), P(t1n),
                                          NB, H, W, S()))
    torch.cuda.synchronize()
    x = t1.float().permute(0, 3, 1, 2)
    t2 = bf(F.relu(F.conv2d(x, w2.float(), b2, padding=1))).float()
    if DS:      # the unfused path rounds the downsample output to bf16 before the add; here it stays in the fp32 accumulator
        short = F.conv2d(x0.float().permute(0, 3, 1, 2), wd.float()[:, :, None, None], bd)
    else:
        short = res.float().permute(0, 3, 1, 2)
    yr = bf(F.relu(F.conv2d(t2, w3.float()[:, :, None, None], b3) + short))
    assert torch.isfinite(y.float()).all()
    assert rel_err(y.permute(0, 3, 1, 2), yr.float()) < 1e-2
    if C1:
        tr = F.relu(F.conv2d(yr.float(), w1.float()[:, :, None, None], b1))
        assert torch.isfinite(t1n.float()).all()
        assert rel_err(t1n.permute(0, 3, 1, 2), tr) < 1e-2


@pytest.mark.parametrize("NB,H,W,Cin,Cmid,Cout,stride", [(4, 56, 56, 256, 128, 512, 2), (5, 15, 13, 64, 64, 256, 2),
                                                          (3, 14, 14, 1024, 512, 2048, 2), (2, 20, 12, 64, 64, 256, 1)])
def test_conv3_plus_downsample_as_one_gemm(h, NB, H, W, Cin, Cmid, Cout, stride):
    """conv3 + the stride-s downsample conv of the block input as one implicit GEMM over two tensors (K-concatenated
    weights) against fp32 torch convs; odd sizes, stride 1 and 2."""
    g = torch.Generator(device="cuda").manual_seed(NB + H + W + Cin + Cout)
    rn = lambda *sh: torch.randn(*sh, device="cuda", generator=g)
    OH, OW = (H - 1) // stride + 1, (W - 1) // stride + 1
    t2 = bf(rn(NB, OH, OW, Cmid).relu())
    x = bf(rn(NB, H, W, Cin).relu())
    w3 = bf(rn(Cout, Cmid) * (Cmid ** -0.5)); wd = bf(rn(Cout, Cin) * (Cin ** -0.5))
    b = rn(Cout) * 0.5
    wcat = torch.cat([w3, wd], dim=1).contiguous()
    out = torch.full((NB, OH, OW, Cout), float("nan"), device="cuda", dtype=torch.bfloat16)
    _lib.check(_lib.lib().mmdx_op_conv3_ds(h.handle, P(t2), P(x), P(wcat), P(b), P(out), NB, H, W, Cin, Cmid, Cout, stride, S()))
    torch.cuda.synchronize()
    ref = F.relu(F.conv2d(t2.float().permute(0, 3, 1, 2), w3.float()[:, :, None, None], b) +
                 F.conv2d(x.float().permute(0, 3, 1, 2), wd.float()[:, :, None, None], stride=stride))
    assert torch.isfinite(out.float()).all()
    assert rel_err(out.permute(0, 3, 1, 2), ref) < 1e-2


def test_gemm_strided_output_and_no_bias(h):
    M, N, K = 200, 512, 768
    a = bf(torch.randn(M, K, device="cuda"))
    w = bf(torch.randn(N, K, device="cuda") * (K ** -0.5))
    buf = torch.zeros(M, 1536, device="cuda", dtype=torch.bfloat16)
    out = buf[:, 1024:]
    _lib.check(_lib.lib().mmdx_op_gemm(h.handle, P(a), K, P(w), None, None, 0, P(out), 1536, M, N, K, 0, 0, 0, S()))
    torch.cuda.synchronize()
    assert rel_err(out, a.float() @ w.float().t()) < 1e-2
    assert float(buf[:, :1024].abs().max()) == 0.0


# ---------------------------------------------------------------------------------------- conv
def _conv_case(h, NB, H, W, Cin, Cout, k, stride, act, res, seed=0):
    g = torch.Generator(device="cuda").manual_seed(seed + NB + H + Cin + Cout + k + stride)
    x = bf(torch.randn(NB, H, W, Cin, device="cuda", generator=g))
    w = bf(torch.randn(Cout, Cin, k, k, device="cuda", generator=g) * ((Cin * k * k) ** -0.5))
    bias = torch.randn(Cout, device="cuda", generator=g)
    pad = k // 2
    OH, OW = (H + 2 * pad - k) // stride + 1, (W + 2 * pad - k) // stride + 1
    r = bf(torch.randn(NB, OH, OW, Cout, device="cuda", generator=g)) if res else None
    wp = w.permute(0, 2, 3, 1).contiguous()       # [Cout][k][k][Cin]
    out = torch.full((NB, OH, OW, Cout), float("nan"), device="cuda", dtype=torch.bfloat16)
    _lib.check(_lib.lib().mmdx_op_conv(h.handle, P(x), NB, H, W, Cin, P(wp), P(bias), P(r), P(out), Cout, k, stride, act, S()))
    torch.cuda.synchronize()
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), w.float(), bias, stride=stride, padding=pad).permute(0, 2, 3, 1)
    if res:
        ref = ref + r.float()
    if act == 1:
        ref = F.relu(ref)
    assert torch.isfinite(out.float()).all()
    assert rel_err(out, ref) < 1.2e-2


@pytest.mark.parametrize("NB,H,W,Cin,Cout,k,stride,act,res", [
    (2, 56, 56, 64, 64, 1, 1, 1, False),       # layer1 conv1
    (2, 56, 56, 64, 64, 3, 1, 1, False),       # layer1 conv2
    (3, 56, 56, 64, 256, 1, 1, 1, True),       # layer1 conv3 + residual
    (2, 56, 56, 128, 128, 3, 2, 1, False),     # layer2.0 conv2 (stride 2)
    (2, 56, 56, 256, 512, 1, 2, 0, False),     # layer2.0 downsample
    (4, 28, 28, 128, 128, 3, 1, 1, False),     # layer2 conv2
    (5, 14, 14, 256, 256, 3, 1, 1, False),     # layer3 conv2
    (2, 14, 14, 512, 512, 3, 2, 1, False),     # layer4.0 conv2
    (3, 7, 7, 512, 512, 3, 1, 1, False),       # layer4 conv2, small batch
    (130, 7, 7, 512, 512, 3, 1, 1, False),     # layer4 conv2, (1,1,128)-style tiles
    (1, 7, 7, 512, 2048, 1, 1, 1, True),       # layer4 conv3, batch 1
    (2, 15, 13, 64, 64, 3, 1, 0, False),       # odd sizes
    (3, 128, 128, 64, 64, 3, 1, 1, False),     # layer1 conv2 at 512x512 input (halo-tile kernel, 8 x 8 tiles per image)
    (150, 9, 20, 64, 64, 3, 1, 1, False),      # halo-tile kernel, more tiles than SMs, ragged both ways
    (2, 15, 13, 64, 128, 3, 2, 0, False),      # odd sizes, stride 2
    (2, 15, 13, 64, 128, 1, 2, 0, False),
])
def test_conv_implicit_gemm(h, NB, H, W, Cin, Cout, k, stride, act, res):
    _conv_case(h, NB, H, W, Cin, Cout, k, stride, act, res)


def _pad_nhwc4(x_nchw):
    """fp32 [B,3,H,W] -> bf16 zero-bordered [B,hp,wp,4] with the image at (3,3) (what K_pre writes)."""
    B, _, H, W = x_nchw.shape
    hp, wp = C.c_int(), C.c_int()
    _lib.lib().mmdx_padded_dims(H, W, C.byref(hp), C.byref(wp))
    buf = torch.zeros(B, hp.value, wp.value, 4, device="cuda", dtype=torch.bfloat16)
    buf[:, 3:3 + H, 3:3 + W, :3] = bf(x_nchw.permute(0, 2, 3, 1))
    return buf


def _pack_stem(w):
    out = np.zeros(7 * 4 * 8 * 8 * 8, dtype=np.uint16)
    wf = np.ascontiguousarray(w.float().cpu().numpy())
    assert _lib.lib().mmdx_pack_stem_weights(wf.ctypes.data_as(C.c_void_p), None, out.ctypes.data_as(C.c_void_p)) == 0
    return torch.from_numpy(out.view(np.int16)).cuda().view(torch.bfloat16)


@pytest.mark.parametrize("B,H,W", [(2, 224, 224), (1, 64, 96), (3, 30, 34), (5, 224, 224), (1, 512, 512), (2, 225, 231)])
@pytest.mark.parametrize("pool", [0, 1])
def test_stem_fused_kernel(h, B, H, W, pool):
    """conv1 + bias + ReLU (+ MaxPool 3x3/2) straight from raw image rows in shared memory (stem_tcgen05.cuh)
    against torch on the same bf16 inputs; the pooled output must equal max-pooling the bf16-rounded conv rows."""
    g = torch.Generator(device="cuda").manual_seed(B + H + pool)
    x = bf(torch.randn(B, 3, H, W, device="cuda", generator=g)).float()
    w = bf(torch.randn(64, 3, 7, 7, device="cuda", generator=g) * (147 ** -0.5))
    bias = torch.randn(64, device="cuda", generator=g)
    wpk = _pack_stem(w)
    OH, OW = (H - 1) // 2 + 1, (W - 1) // 2 + 1
    PH, PW = (OH - 1) // 2 + 1, (OW - 1) // 2 + 1
    shape = (B, PH, PW, 64) if pool else (B, OH, OW, 64)
    out = torch.full(shape, float("nan"), device="cuda", dtype=torch.bfloat16)
    xin = _pad_nhwc4(x)
    _lib.check(_lib.lib().mmdx_op_stem_pool(h.handle, P(xin), B, H, W, P(wpk), P(bias), P(out), pool, S()))
    torch.cuda.synchronize()
    ref = F.relu(F.conv2d(x, w.float(), bias, stride=2, padding=3))
    if pool:
        ref = F.max_pool2d(ref, 3, 2, 1)
    ref = ref.permute(0, 2, 3, 1)
    assert torch.isfinite(out.float()).all()
    assert rel_err(out, ref) < 1.2e-2


# ---------------------------------------------------------------------------------------- preprocessing
@pytest.mark.parametrize("B,H,W,Cc", [(2, 512, 512, 3), (3, 224, 224, 3), (1, 300, 400, 3), (1, 1024, 768, 3),
                                      (2, 257, 640, 1), (1, 256, 256, 3), (1, 256, 300, 3)])
def test_resample_u8_bit_exact(h, B, H, W, Cc):
    rng = np.random.Generator(np.random.PCG64([3, H, W]))
    a = rng.integers(0, 256, size=(B, H, W, Cc), dtype=np.uint8)
    d = torch.from_numpy(a).cuda()
    out = torch.zeros(B, 224, 224, Cc, dtype=torch.uint8, device="cuda")
    _lib.check(_lib.lib().mmdx_op_resample_u8(h.handle, P(d), B, H, W, Cc, P(out), S()))
    torch.cuda.synchronize()
    want = np.stack([R.preprocess_u8(a[i]) for i in range(B)])
    assert np.array_equal(out.cpu().numpy(), want)


@pytest.mark.parametrize("B,H,W,Cc", [(2, 512, 512, 3), (2, 224, 224, 3), (1, 300, 400, 1)])
def test_preprocess_normalized_bf16(h, B, H, W, Cc):
    rng = np.random.Generator(np.random.PCG64([4, H, W]))
    a = rng.integers(0, 256, size=(B, H, W, Cc), dtype=np.uint8)
    d = torch.from_numpy(a).cuda()
    hp, wp = C.c_int(), C.c_int()
    _lib.lib().mmdx_padded_dims(224, 224, C.byref(hp), C.byref(wp))
    out = torch.zeros(B, hp.value, wp.value, 4, dtype=torch.bfloat16, device="cuda")
    oh, ow = C.c_int(), C.c_int()
    _lib.check(_lib.lib().mmdx_op_preprocess(h.handle, P(d), B, H, W, Cc, P(out), C.byref(oh), C.byref(ow), S()))
    torch.cuda.synchronize()
    assert (oh.value, ow.value) == (224, 224)
    want = torch.stack([R.preprocess_f32(a[i]) for i in range(B)]).permute(0, 2, 3, 1)       # fp32 NHWC
    got = out[:, 3:227, 3:227, :3].float().cpu()
    assert torch.equal(got, want.to(torch.bfloat16).float())       # exactly the bf16 rounding of the fp32 reference
    o = out.float()
    assert float(o[..., 3].abs().max()) == 0 and float(o[:, :3].abs().max()) == 0 and float(o[:, :, :3].abs().max()) == 0
    assert float(o[:, 227:].abs().max()) == 0 and float(o[:, :, 227:].abs().max()) == 0


# ---------------------------------------------------------------------------------------- pooling / norms
def test_avgpool(h):
    y = bf(torch.randn(5, 49, 2048, device="cuda"))
    ob = torch.zeros(5, 2048, device="cuda", dtype=torch.bfloat16)
    of = torch.zeros(5, 2048, device="cuda")
    _lib.check(_lib.lib().mmdx_op_avgpool(h.handle, P(y), 5, 49, 2048, P(ob), P(of), S()))
    torch.cuda.synchronize()
    ref = y.float().mean(1)
    assert (of - ref).abs().max() < 1e-5 and rel_err(ob, ref) < 5e-3


@pytest.mark.parametrize("rows,N,eps", [(1000, 768, 1e-12), (7, 1024, 1e-5), (64, 256, 1e-12)])
def test_layernorm(h, rows, N, eps):
    x = bf(torch.randn(rows, N, device="cuda") * 3 + 0.5)
    g = torch.rand(N, device="cuda") + 0.5
    b = torch.randn(N, device="cuda") * 0.1
    y = torch.zeros_like(x)
    _lib.check(_lib.lib().mmdx_op_layernorm(h.handle, P(x), rows, N, P(g), P(b), eps, P(y), S()))
    torch.cuda.synchronize()
    ref = F.layer_norm(x.float(), (N,), g, b, eps)
    assert rel_err(y, ref) < 6e-3


def test_embed_layernorm(h):
    T, N = 777, 768
    word = bf(torch.randn(30522, N, device="cuda") * 0.05)
    ptab = bf(torch.randn(512, N, device="cuda") * 0.05)
    ttab = bf(torch.randn(2, N, device="cuda") * 0.05)
    ids = torch.randint(0, 30522, (T,), device="cuda", dtype=torch.int32)
    pos = torch.randint(0, 512, (T,), device="cuda", dtype=torch.int32)
    tt = torch.randint(0, 2, (T,), device="cuda", dtype=torch.int32)
    g = torch.rand(N, device="cuda") + 0.5
    b = torch.randn(N, device="cuda") * 0.1
    y = torch.zeros(T, N, device="cuda", dtype=torch.bfloat16)
    _lib.check(_lib.lib().mmdx_op_embed_ln(h.handle, P(ids), P(pos), P(tt), T, P(word), P(ptab), P(ttab), P(g), P(b), 1e-12, P(y), S()))
    torch.cuda.synchronize()
    ref = F.layer_norm(word[ids.long()].float() + ttab[tt.long()].float() + ptab[pos.long()].float(), (N,), g, b, 1e-12)
    assert rel_err(y, ref) < 6e-3


@pytest.mark.parametrize("lens", [[128, 128, 128], [17, 64, 65, 1, 40], [512, 300], [96] * 4, [129, 5, 128, 33, 257],
                                  [31, 32, 33, 1, 16, 15, 127] * 30])
def test_attention_varlen(h, lens):
    heads, hid = 12, 768
    T = sum(lens)
    qkv = bf(torch.randn(T, 3 * hid, device="cuda"))
    cu = torch.tensor(np.concatenate([[0], np.cumsum(lens)]), dtype=torch.int32, device="cuda")
    ctx = torch.full((T, hid), float("nan"), device="cuda", dtype=torch.bfloat16)
    _lib.check(_lib.lib().mmdx_op_attention(h.handle, P(qkv), P(cu), len(lens), T, max(lens), heads, hid, P(ctx), S()))
    torch.cuda.synchronize()
    ref = torch.empty(T, hid, device="cuda")
    o = 0
    for n in lens:
        q, k, v = [qkv[o:o + n, i * hid:(i + 1) * hid].float().view(n, heads, 64).transpose(0, 1) for i in range(3)]
        s = torch.softmax(q @ k.transpose(-1, -2) * 0.125, -1)
        ref[o:o + n] = (s @ v).transpose(0, 1).reshape(n, hid)
        o += n
    assert torch.isfinite(ctx.float()).all()
    assert rel_err(ctx, ref) < 1.5e-2


def test_attention_general_kernel_on_short_sequences(monkeypatch):
    """Sequences of <= 128 tokens normally take the 4-deep TMEM-resident kernel; MMDX_ATTN=general pins the
    flash-style kernel so that it is also checked on ragged short inputs (both must agree with torch)."""
    monkeypatch.setenv("MMDX_ATTN", "general")
    hd = engine.RawHandle()
    try:
        test_attention_varlen(hd, [17, 64, 65, 1, 40, 128])
    finally:
        hd.close()


def test_seq_mean_pool_and_head_tail(h):
    lens = [5, 128, 33]
    T, hid = sum(lens), 768
    x = bf(torch.randn(T, hid, device="cuda"))
    cu = torch.tensor(np.concatenate([[0], np.cumsum(lens)]), dtype=torch.int32, device="cuda")
    ob = torch.zeros(3, hid, device="cuda", dtype=torch.bfloat16)
    of = torch.zeros(3, hid, device="cuda")
    _lib.check(_lib.lib().mmdx_op_seq_mean_pool(h.handle, P(x), P(cu), 3, hid, P(ob), P(of), S()))
    torch.cuda.synchronize()
    ref = torch.stack([x[cu[i]:cu[i + 1]].float().mean(0) for i in range(3)])
    assert (of - ref).abs().max() < 1e-5
    # head tail
    B, D, nc = 6, 1024, 13
    hd = torch.randn(B, D, device="cuda")
    g = torch.rand(D, device="cuda") + 0.5
    b = torch.randn(D, device="cuda") * 0.1
    w = torch.randn(nc, D, device="cuda") * 0.06
    bb = torch.randn(nc, device="cuda") * 0.3
    thr = torch.full((nc,), 0.5, device="cuda")
    zf = torch.zeros(B, D, device="cuda"); lg = torch.zeros(B, nc, device="cuda"); pr = torch.zeros(B, nc, device="cuda")
    vec = torch.zeros(B, nc, device="cuda", dtype=torch.uint8)
    _lib.check(_lib.lib().mmdx_op_head_tail(h.handle, P(hd), B, D, P(g), P(b), 1e-5, P(w), P(bb), nc, P(thr), P(zf), P(lg), P(pr), P(vec), S()))
    torch.cuda.synchronize()
    z = F.layer_norm(hd, (D,), g, b, 1e-5)
    l = F.linear(z, w, bb)
    assert (zf - z).abs().max() < 1e-4 and (lg - l).abs().max() < 1e-4
    assert (pr - torch.sigmoid(l)).abs().max() < 1e-5
    assert torch.equal(vec, (pr >= thr).to(torch.uint8))


# ---------------------------------------------------------------------------------------- conv3 + next conv1 (two-GEMM launch)
@pytest.mark.parametrize("M,K1,N1,N2", [
    (256, 128, 512, 128),          # one item
    (1000, 128, 512, 128),         # ragged: odd number of m-tiles, partial last tile (layer-2 shape, BN 128)
    (50176, 256, 1024, 256),       # layer 3 at B = 256 (BN 256): 196 items on 74 CTA pairs
    (40000, 128, 512, 256),        # layer2.3 -> layer3.0 conv1 (BN 256 with N1 = 512)
    (30011, 256, 1024, 512),       # layer3.5 -> layer4.0 conv1 (two G2 tiles per item), ragged
    (100352, 128, 512, 128),       # layer 2 at B = 128: 392 items
])
def test_conv3_conv1_two_gemm_launch(h, M, K1, N1, N2):
    """gemm2_tcgen05_kernel: y = relu(t2 W3^T + b3 + res), t1n = relu(y W1n^T + b1n) in one launch, against fp32 torch on
    the same bf16 inputs (the second GEMM consumes the bf16-rounded y, like the two-launch path)."""
    g = torch.Generator(device="cuda").manual_seed(M + K1 + N1 + N2)
    t2 = bf(torch.randn(M, K1, generator=g, device="cuda").abs())
    res = bf(torch.randn(M, N1, generator=g, device="cuda"))
    w3 = bf(torch.randn(N1, K1, generator=g, device="cuda") * (K1 ** -0.5))
    w1 = bf(torch.randn(N2, N1, generator=g, device="cuda") * (N1 ** -0.5))
    b3 = torch.randn(N1, generator=g, device="cuda") * 0.1
    b1 = torch.randn(N2, generator=g, device="cuda") * 0.1
    y = torch.full((M, N1), float("nan"), dtype=torch.bfloat16, device="cuda")
    t1n = torch.full((M, N2), float("nan"), dtype=torch.bfloat16, device="cuda")
    for rep in range(2):                      # twice: the second launch walks the items in the other direction (zigzag)
        _lib.check(_lib.lib().mmdx_op_conv3_conv1(h.handle, P(t2), P(w3), P(b3), P(res), P(y), P(w1), P(b1), P(t1n), M, K1, N1, N2, S()))
        torch.cuda.synchronize()
        y_ref = torch.relu(t2.float() @ w3.float().t() + b3 + res.float())
        assert rel_err(y, y_ref) < 1e-2, (rep, rel_err(y, y_ref))
        t_ref = torch.relu(y.float() @ w1.float().t() + b1)          # from the kernel's own (bf16) y
        assert rel_err(t1n, t_ref) < 1e-2, (rep, rel_err(t1n, t_ref))
        y.fill_(float("nan")); t1n.fill_(float("nan"))
