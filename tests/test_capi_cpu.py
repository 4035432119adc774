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
