"""ctypes binding of libmmdx.so (C ABI declared in include/mmdx.h).

The library is built in-tree by csrc/build.sh (`__graft_entry__.build()`); there is no
CPU fallback - if it is missing the import of this module fails loudly."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MMDX_LIB") or os.path.join(_HERE, "libmmdx.so")     # MMDX_LIB: A/B builds of the same source


class MmdxError(RuntimeError):
    pass


class Config(C.Structure):
    _fields_ = [("device", C.c_int32), ("resize_short", C.c_int32), ("crop", C.c_int32), ("n_heads", C.c_int32),
                ("mean", C.c_float * 3), ("std", C.c_float * 3), ("keep_fp32", C.c_int32)]


def build(force: bool = False) -> str:
    """Compile libmmdx.so for sm_100a (nvcc cross-compiles without a GPU)."""
    env = dict(os.environ)
    if force:
        env["FORCE"] = "1"
    r = subprocess.run(["bash", os.path.join(_HERE, "csrc", "build.sh")], env=env, capture_output=True, text=True)
    if r.returncode != 0:
        raise MmdxError("building libmmdx.so failed:\n" + r.stdout + r.stderr)
    return LIB_PATH


_p, _i, _i64, _f = C.c_void_p, C.c_int, C.c_int64, C.c_float
_ip = C.POINTER(C.c_int)

# name -> argtypes (restype is int unless listed in _RESTYPE)
SIGNATURES = {
    "mmdx_last_error": [],
    "mmdx_version": [],
    "mmdx_resample_coeffs": [_i, _i, _i, _i, _p, _p, _p, _i],
    "mmdx_resize_geometry": [_i, _i, _i, _i, _ip, _ip, _ip, _ip],
    "mmdx_padded_dims": [_i, _i, _ip, _ip],
    "mmdx_create": [C.POINTER(Config), C.POINTER(_p)],
    "mmdx_destroy": [_p],
    "mmdx_load_tensor": [_p, C.c_char_p, _p, _i, C.POINTER(C.c_int64)],
    "mmdx_finalize_weights": [_p],
    "mmdx_save_packed": [_p, C.c_char_p],
    "mmdx_load_packed": [_p, C.c_char_p],
    "mmdx_num_sms": [_p],
    "mmdx_dims": [_p, C.POINTER(C.c_int32)],
    "mmdx_table_sizes": [_p, C.POINTER(C.c_int32)],
    "mmdx_image_encode": [_p, _p, _i, _i, _i, _i, _p, _p, _p],
    "mmdx_text_encode": [_p, _p, _p, _p, _p, _i, _i, _i, _p, _p, _p],
    "mmdx_head": [_p, _i, _p, _p, _p, _p, _p, _p],
    "mmdx_cond_tokens": [_p, _i, _p, _p],
    "mmdx_forward": [_p, _p, _i, _i, _i, _i, _p, _p, _p, _p, _i, _i, _p, _p, _p, _p, _p],
    "mmdx_forward_f32": [_p, _p, _i, _i, _i, _i, _p, _p, _p, _p, _i, _i, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p],
    "mmdx_forward_host": [_p, _p, _i, _i, _i, _i, _p, _p, _p, _p, _i, _i, _p, _p, _p, _p, _p],
    "mmdx_forward_host_submit": [_p, _i, _p, _i, _i, _i, _i, _p, _p, _p, _p, _i, _i, _p, _p, _p, _p, _p],
    "mmdx_forward_host_wait": [_p, _i],
    "mmdx_decode_jpeg_batch": [_p, _p, _p, _i, _i, _i, _p, _p],
    "mmdx_jpeg_backend": [_p],
    "mmdx_launch_count": [_p],
    "mmdx_profile_begin": [_p],
    "mmdx_profile_end": [_p, _p, _p, _i],
    "mmdx_profile_end_list": [_p, _p, _p, _i],
    "mmdx_tokenizer_create": [C.c_char_p, C.c_size_t, _i, C.POINTER(_p)],
    "mmdx_tokenizer_destroy": [_p],
    "mmdx_tokenize_batch": [_p, _p, _p, _i, _i, _i, _p, _p, _p],
    "mmdx_tokenizer_last_error": [],
    "mmdx_t5_create": [_i, _i, _i, _i, _i, _i, _i, _f, _i, C.POINTER(_p)],
    "mmdx_t5_destroy": [_p],
    "mmdx_t5_load_tensor": [_p, C.c_char_p, _p, _i64],
    "mmdx_t5_finalize": [_p],
    "mmdx_t5_begin": [_p, _p, _i, _i, _i, _p, _p],
    "mmdx_t5_reorder": [_p, _p, _p],
    "mmdx_t5_step": [_p, _p, _p, _p],
    "mmdx_t5_score_topk": [_p, _p, _p, _p, _i, _i, _i, _i, _i, _p, _p, _p],
    "mmdx_t5_generate": [_p, _p, _i, _i, _i, _i, _i, _i, _f, _i, _i, _i, _i, _p, _p, _p, _p],
    "mmdx_t5_launch_count": [_p],
    "mmdx_t5_step_profile": [_p, _p, C.c_int, _p],
    "mmdx_t5_last_error": [],
    "mmdx_op_gemm": [_p, _p, _i64, _p, _p, _p, _i64, _p, _i64, _i, _i, _i, _i, _i, _i, _p],
    "mmdx_op_gemm_ln": [_p, _p, _i64, _p, _p, _p, _i64, _p, _i64, _p, _p, _f, _p, _i64, _i, _i, _i, _p],
    "mmdx_op_conv": [_p, _p, _i, _i, _i, _i, _p, _p, _p, _p, _i, _i, _i, _i, _p],
    "mmdx_op_bneck64": [_p, _p, _p, _p, _p, _p, _p, _p, _p, _i, _p, _p, _p, _p, _p, _i, _i, _i, _p],
    "mmdx_op_conv3_ds": [_p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _p],
    "mmdx_op_conv3_conv1": [_p, _p, _p, _p, _p, _p, _p, _p, _p, _i64, _i, _i, _i, _p],
    "mmdx_gemm2_schedule": [_i, _i, _i, _i, _i, _i, _p, _i],
    "mmdx_pack_stem_weights": [_p, _p, _p],
    "mmdx_op_stem_pool": [_p, _p, _i, _i, _i, _p, _p, _p, _i, _p],
    "mmdx_op_preprocess": [_p, _p, _i, _i, _i, _i, _p, _ip, _ip, _p],
    "mmdx_op_resample_u8": [_p, _p, _i, _i, _i, _i, _p, _p],
    "mmdx_op_avgpool": [_p, _p, _i, _i, _i, _p, _p, _p],
    "mmdx_op_layernorm": [_p, _p, _i, _i, _p, _p, _f, _p, _p],
    "mmdx_op_embed_ln": [_p, _p, _p, _p, _i, _p, _p, _p, _p, _p, _f, _p, _p],
    "mmdx_op_attention": [_p, _p, _p, _i, _i, _i, _i, _i, _p, _p],
    "mmdx_op_seq_mean_pool": [_p, _p, _p, _i, _i, _p, _p, _p],
    "mmdx_op_head_tail": [_p, _p, _i, _i, _p, _p, _f, _p, _p, _i, _p, _p, _p, _p, _p, _p],
}
_RESTYPE = {"mmdx_last_error": C.c_char_p, "mmdx_version": C.c_char_p, "mmdx_destroy": None,
            "mmdx_tokenizer_last_error": C.c_char_p, "mmdx_tokenizer_destroy": None,
            "mmdx_t5_last_error": C.c_char_p, "mmdx_t5_destroy": None, "mmdx_t5_launch_count": C.c_int64,
            "mmdx_launch_count": C.c_int64}

_lib = None


def lib():
    """Loads libmmdx.so (once) and declares every entry point of include/mmdx.h."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise MmdxError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                            "(there is no CPU fallback for the mmdx hot path)")
        l = C.CDLL(LIB_PATH)
        for name, args in SIGNATURES.items():
            fn = getattr(l, name)
            fn.argtypes = args
            fn.restype = _RESTYPE.get(name, C.c_int)
        _lib = l
    return _lib


def check(rc: int):
    if rc != 0:
        raise MmdxError(lib().mmdx_last_error().decode())
