"""CPU-side checks of the C-ABI library: it loads without a GPU, exports every symbol that
include/mmdx.h declares, the host-only helpers agree with the oracle, and the product refuses
to run without CUDA instead of falling back."""
import ctypes as C
import os
import re

import numpy as np
import pytest
import torch

from conftest import ROOT
from mmdx_b200 import _lib, engine
from oracle import forward_ref as R


@pytest.fixture(scope="module")
def lib():
    _lib.build()
    return _lib.lib()


def test_header_symbols_all_exported(lib):
    hdr = open(os.path.join(ROOT, "include", "mmdx.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    names = set(re.findall(r"\b(mmdx_[a-z0-9_]+)\s*\(", hdr))
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/mmdx.h but not exported by libmmdx.so"
    assert names == set(_lib.SIGNATURES), names ^ set(_lib.SIGNATURES)
    assert b"sm_100a" in lib.mmdx_version()


@pytest.mark.parametrize("in_size,out_size", [(512, 256), (224, 256), (300, 256), (1024, 341), (640, 637), (257, 256)])
def test_resample_coeffs_match_oracle(lib, in_size, out_size):
    xmin, xcnt, wts = R.bilinear_coeffs(in_size, out_size)
    first, n = 7, min(224, out_size - 7)
    ks = wts.shape[1]
    f = np.zeros(n, np.int32); c = np.zeros(n, np.int32); w = np.zeros((n, ks), np.int32)
    got = lib.mmdx_resample_coeffs(in_size, out_size, first, n, f.ctypes.data, c.ctypes.data, w.ctypes.data, w.size)
    assert got == ks
    assert np.array_equal(f, xmin[first:first + n]) and np.array_equal(c, xcnt[first:first + n])
    assert np.array_equal(w, wts[first:first + n])


@pytest.mark.parametrize("hw", [(512, 512), (224, 224), (300, 400), (1024, 768), (257, 640), (1000, 999)])
def test_resize_geometry_matches_oracle(lib, hw):
    oh, ow, top, left = C.c_int(), C.c_int(), C.c_int(), C.c_int()
    assert lib.mmdx_resize_geometry(hw[0], hw[1], 256, 224, C.byref(oh), C.byref(ow), C.byref(top), C.byref(left)) == 0
    eh, ew = R.resize_output_size(hw[0], hw[1], 256)
    assert (oh.value, ow.value) == (eh, ew)
    assert (top.value, left.value) == R.center_crop_offsets(eh, ew, 224)


def test_pack_tokens_unpads():
    ids = np.array([[101, 5, 6, 102, 0, 0], [101, 7, 102, 0, 0, 0]])
    mask = (ids != 0).astype(np.int64)
    i, p, t, cu, mlen = engine.pack_tokens(ids, mask)
    assert i.tolist() == [101, 5, 6, 102, 101, 7, 102] and p.tolist() == [0, 1, 2, 3, 0, 1, 2]
    assert cu.tolist() == [0, 4, 7] and mlen == 4 and t.sum() == 0
    with pytest.raises(ValueError):
        engine.pack_tokens(ids, np.zeros_like(mask))


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_no_cpu_fallback(lib):
    with pytest.raises(_lib.MmdxError):
        engine.RawHandle()
    from mmdx_b200 import inference_pipeline as ip
    with pytest.raises(TypeError):
        ip.inference({}, None, "x", device=3)
    with pytest.raises(RuntimeError):
        ip.inference({}, None, "x", device="cpu")
    cfg = _lib.Config(0, 256, 224, 12, (C.c_float * 3)(0, 0, 0), (C.c_float * 3)(1, 1, 1))
    h = C.c_void_p()
    assert lib.mmdx_create(C.byref(cfg), C.byref(h)) != 0
    assert b"no CUDA device" in lib.mmdx_last_error()


@pytest.mark.parametrize("num_items,nt1,nt2,groups,reverse", [(784, 4, 1, 74, 0), (196, 4, 1, 74, 1), (196, 4, 2, 74, 0),
                                                              (5, 2, 1, 74, 0), (1, 4, 1, 1, 1), (75, 3, 2, 74, 1), (0, 4, 1, 4, 0)])
def test_two_gemm_schedule_covers_every_job_once(lib, num_items, nt1, nt2, groups, reverse):
    """The static schedule of gemm2_tcgen05_kernel (conv3 + next conv1 in one launch), enumerated by the same iterator the
    device runs: every (GEMM, item, n-tile) job exactly once over all CTA pairs, an item's second GEMM after ALL tiles of
    its first one in the same pair, and one item of look-ahead between them (so the stores have landed)."""
    seen = {}
    g_used = min(groups, max(num_items, 1))
    for g in range(g_used):
        buf = np.zeros(3 * 4096, np.int32)
        n = lib.mmdx_gemm2_schedule(num_items, nt1, nt2, reverse, g, g_used, buf.ctypes.data, 4096)
        jobs = buf[:3 * n].reshape(n, 3).tolist()
        pos = {}
        for i, (t, item, nt) in enumerate(jobs):
            assert (t, item, nt) not in seen
            seen[(t, item, nt)] = g
            pos[(t, item, nt)] = i
        items = sorted({j[1] for j in jobs}, key=lambda it: min(p for (t, i2, _), p in pos.items() if i2 == it))
        for k, it in enumerate(items):
            last_g1 = max(pos[(0, it, nt)] for nt in range(nt1))
            first_g2 = min(pos[(1, it, nt)] for nt in range(nt2))
            assert first_g2 > last_g1
            if k + 1 < len(items):          # the next item's first GEMM is issued before this item's second one
                assert pos[(0, items[k + 1], nt1 - 1)] < first_g2
    assert len(seen) == num_items * (nt1 + nt2)
    assert {k[1] for k in seen} == set(range(num_items))
