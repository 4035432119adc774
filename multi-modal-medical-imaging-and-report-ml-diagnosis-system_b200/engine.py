"""Python host side of the engine: weight import from the reference's state_dicts, token packing,
and thin wrappers over the C ABI (`include/mmdx.h`).  PyTorch is used only for device memory and streams."""
from __future__ import annotations

import ctypes as C
import os
import threading

import numpy as np
import torch

from . import _lib
from ._lib import Config, MmdxError, check, lib

IMAGENET_MEAN = (0.485, 0.456, 0.406)   # training_pipeline.py:117
IMAGENET_STD = (0.229, 0.224, 0.225)


def _ptr(t):
    return C.c_void_p(0 if t is None else t.data_ptr())


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def pack_tokens(input_ids, attention_mask, token_type_ids=None, table_sizes=None):
    """Host-side unpadding: keep the tokens with attention_mask==1 (padded keys are masked to -inf and
    padded query rows are dropped by the masked mean in the reference, training_pipeline.py:452-459, so
    computing only valid tokens is results-identical).  Returns int32 numpy arrays
    (ids[T], pos[T], tt[T], cu_seqlens[B+1]) and max_len.
    `table_sizes` = (vocab, max_positions, type_vocab) of the engine (Engine.table_sizes): ids outside the
    embedding tables raise IndexError, like nn.Embedding in the reference."""
    ids = np.asarray(input_ids)
    mask = np.asarray(attention_mask).astype(bool)
    B, L = ids.shape
    tt = np.zeros_like(ids) if token_type_ids is None else np.asarray(token_type_ids)
    if table_sizes is not None:
        vocab, max_pos, type_vocab = table_sizes
        if ids[mask].size and (ids[mask].min() < 0 or ids[mask].max() >= vocab):
            raise IndexError(f"token id out of range for the word embedding table ({vocab} rows)")
        if tt[mask].size and (tt[mask].min() < 0 or tt[mask].max() >= type_vocab):
            raise IndexError(f"token_type_id out of range for the token-type embedding table ({type_vocab} rows)")
        if L > max_pos and mask[:, max_pos:].any():
            raise IndexError(f"sequence longer than the position embedding table ({max_pos} rows)")
    lens = mask.sum(1)
    if (lens == 0).any():
        raise ValueError("every study needs at least one unmasked token")
    pos = np.broadcast_to(np.arange(L, dtype=np.int32), (B, L))
    cu = np.zeros(B + 1, np.int32)
    np.cumsum(lens, out=cu[1:])
    return (np.ascontiguousarray(ids[mask], dtype=np.int32), np.ascontiguousarray(pos[mask], dtype=np.int32),
            np.ascontiguousarray(tt[mask], dtype=np.int32), cu, int(lens.max()))


class Engine:
    """One engine per process/GPU.  `states` = {"image": sd, "text": sd, "fusion": sd} with the reference's
    state_dict key names."""

    def __init__(self, states: dict | None, device: int | None = None, resize_short: int = 256, crop: int = 224,
                 n_heads: int = 12, packed: str | None = None, fp32: bool = False):
        """fp32=True also keeps an fp32 copy of the weights (+452 MB) for `forward_f32`, the parity mode held to 1e-5."""
        if not torch.cuda.is_available():
            raise MmdxError("mmdx needs a CUDA device (B200, sm_100a); there is no CPU fallback")
        self.device = torch.cuda.current_device() if device is None else int(device)
        self.has_fp32 = bool(fp32) and packed is None
        cfg = Config(self.device, resize_short, crop, n_heads, (C.c_float * 3)(*IMAGENET_MEAN),
                     (C.c_float * 3)(*IMAGENET_STD), 1 if self.has_fp32 else 0)
        self._h = C.c_void_p()
        check(lib().mmdx_create(C.byref(cfg), C.byref(self._h)))
        if packed is not None:
            if states is not None:
                raise ValueError("pass either state dicts or a packed weight file, not both")
            check(lib().mmdx_load_packed(self._h, os.fsencode(packed)))
        else:
            skip = ("num_batches_tracked", "report_model.", "classifier.", "pooler.", "position_ids")
            for prefix, sd in states.items():
                for k, v in sd.items():
                    if any(s in k for s in skip):
                        continue   # dead at inference (I3, T6, T9) or off the named path (T5)
                    t = v.detach().to(dtype=torch.float32, device="cpu").contiguous()
                    shape = (C.c_int64 * max(t.dim(), 1))(*t.shape)
                    check(lib().mmdx_load_tensor(self._h, f"{prefix}.{k}".encode(), C.c_void_p(t.data_ptr()), t.dim(),
                                                 shape))
            check(lib().mmdx_finalize_weights(self._h))
        d = (C.c_int32 * 8)()
        check(lib().mmdx_dims(self._h, d))
        self.d_img, self.d_txt, self.d_fuse, self.n_cls, self.hidden, self.n_layers, self.cond_width, self.max_pos = list(d)
        self.feat_dim = 2048
        t = (C.c_int32 * 3)()
        check(lib().mmdx_table_sizes(self._h, t))
        self.vocab, _, self.type_vocab = list(t)
        # A request is often several C calls (forward, then cond_tokens; or image_encode / text_encode / head) that
        # hand results to each other through engine-owned buffers: callers that may run concurrently (Django's threaded
        # dev server) hold this lock around the whole sequence.  Each C call also takes the engine's
        # own mutex, so single calls are safe without it.
        self.lock = threading.RLock()

    @property
    def table_sizes(self):
        return self.vocab, self.max_pos, self.type_vocab

    @classmethod
    def from_packed(cls, path: str, device: int | None = None, resize_short: int = 256, crop: int = 224,
                    n_heads: int = 12) -> "Engine":
        """Engine from a packed weight file written by save_packed(): one file read + one H2D copy, no torch modules
        (SURVEY.md section 8f N2 - what a serving process does instead of views.py:188-258)."""
        return cls(None, device, resize_short, crop, n_heads, packed=path)

    def save_packed(self, path: str) -> None:
        """Writes the finalized weight arena (BN folded, QKV fused, bf16 kernel layouts) and its table to `path`."""
        check(lib().mmdx_save_packed(self._h, os.fsencode(path)))

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            lib().mmdx_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def handle(self):
        return self._h

    @property
    def launch_count(self) -> int:
        return int(lib().mmdx_launch_count(self._h))

    def _dev(self):
        return torch.device("cuda", self.device)

    # ---- the hot path on device tensors -------------------------------------------------------
    def image_encode(self, images_u8: torch.Tensor, want_feats=True):
        """images_u8: uint8 [B,H,W,C] on the device -> (feats fp32 [B,2048] | None, z_img fp32 [B,d_img])."""
        assert images_u8.dtype == torch.uint8 and images_u8.dim() == 4 and images_u8.is_cuda and images_u8.is_contiguous()
        B, H, W, Cc = images_u8.shape
        feats = torch.empty(B, self.feat_dim, dtype=torch.float32, device=self._dev()) if want_feats else None
        z = torch.empty(B, self.d_img, dtype=torch.float32, device=self._dev())
        check(lib().mmdx_image_encode(self._h, _ptr(images_u8), B, H, W, Cc, _ptr(feats), _ptr(z), _stream()))
        return feats, z

    def text_encode(self, ids, pos, tt, cu, max_len, want_pooled=True):
        """Packed int32 device tensors -> (pooled fp32 [B,hidden] | None, z_txt fp32 [B,d_txt])."""
        B, T = cu.numel() - 1, ids.numel()
        pooled = torch.empty(B, self.hidden, dtype=torch.float32, device=self._dev()) if want_pooled else None
        z = torch.empty(B, self.d_txt, dtype=torch.float32, device=self._dev())
        check(lib().mmdx_text_encode(self._h, _ptr(ids), _ptr(pos), _ptr(tt), _ptr(cu), B, T, int(max_len), _ptr(pooled),
                                     _ptr(z), _stream()))
        return pooled, z

    def head(self, B, thresholds=None, want_z_fuse=True):
        dev = self._dev()
        z_fuse = torch.empty(B, self.d_fuse, dtype=torch.float32, device=dev) if want_z_fuse else None
        logits = torch.empty(B, self.n_cls, dtype=torch.float32, device=dev)
        probs = torch.empty_like(logits)
        vec = torch.empty(B, self.n_cls, dtype=torch.uint8, device=dev)
        check(lib().mmdx_head(self._h, B, _ptr(thresholds), _ptr(z_fuse), _ptr(logits), _ptr(probs), _ptr(vec), _stream()))
        return z_fuse, logits, probs, vec

    def cond_tokens(self, B, n_cond: int | None = None):
        """GELU(cond_proj(z_fuse)) for the batch head() has just processed: fp32 [B, n_cond*h_dec], or
        [B, n_cond, h_dec] when n_cond is given - the conditioning tokens of the report decoder
        (training_pipeline.py:574-578)."""
        if not self.cond_width:
            raise MmdxError("the bundle has no cond_proj weights")
        out = torch.empty(B, self.cond_width, dtype=torch.float32, device=self._dev())
        check(lib().mmdx_cond_tokens(self._h, B, _ptr(out), _stream()))
        return out if n_cond is None else out.view(B, n_cond, -1)

    def decode_jpeg_batch(self, blobs, H, W):
        """Baseline JPEGs of one size (list of bytes) -> uint8 [n,H,W,3] on the device, decoded by nvJPEG (optional path,
        SURVEY.md 8f N3; ~2 % of the bytes differ by one from Pillow / libjpeg-turbo)."""
        n = len(blobs)
        bufs = [(C.c_ubyte * len(b)).from_buffer_copy(b) for b in blobs]
        ptrs = (C.c_void_p * n)(*[C.addressof(b) for b in bufs])
        sizes = (C.c_size_t * n)(*[len(b) for b in blobs])
        out = torch.empty(n, H, W, 3, dtype=torch.uint8, device=self._dev())
        check(lib().mmdx_decode_jpeg_batch(self._h, ptrs, sizes, n, int(H), int(W), _ptr(out), _stream()))
        return out

    def forward(self, images_u8, ids, pos, tt, cu, max_len, thresholds=None):
        """Whole path, device in / device out: (logits, probs, vector)."""
        B, H, W, Cc = images_u8.shape
        dev = self._dev()
        logits = torch.empty(B, self.n_cls, dtype=torch.float32, device=dev)
        probs = torch.empty_like(logits)
        vec = torch.empty(B, self.n_cls, dtype=torch.uint8, device=dev)
        check(lib().mmdx_forward(self._h, _ptr(images_u8), B, H, W, Cc, _ptr(ids), _ptr(pos), _ptr(tt), _ptr(cu),
                                 ids.numel(), int(max_len), _ptr(thresholds), _ptr(logits), _ptr(probs), _ptr(vec),
                                 _stream()))
        return logits, probs, vec

    def forward_f32(self, images_u8, ids, pos, tt, cu, max_len, thresholds=None):
        """The whole path in fp32 (parity mode, engine created with fp32=True): device tensors in, dict of device
        tensors out - every intermediate the goldens hold plus logits / probs / vector."""
        if not self.has_fp32:
            raise MmdxError("engine was created without fp32 weights (Engine(..., fp32=True))")
        B, H, W, Cc = images_u8.shape
        dev = self._dev()
        f = lambda n: torch.empty(B, n, dtype=torch.float32, device=dev)      # noqa: E731
        o = {"feats": f(self.feat_dim), "z_img": f(self.d_img), "pooled": f(self.hidden), "z_txt": f(self.d_txt),
             "z_fuse": f(self.d_fuse), "logits": f(self.n_cls), "probs": f(self.n_cls),
             "vector": torch.empty(B, self.n_cls, dtype=torch.uint8, device=dev)}
        check(lib().mmdx_forward_f32(self._h, _ptr(images_u8), B, H, W, Cc, _ptr(ids), _ptr(pos), _ptr(tt), _ptr(cu),
                                     ids.numel(), int(max_len), _ptr(thresholds), _ptr(o["feats"]), _ptr(o["z_img"]),
                                     _ptr(o["pooled"]), _ptr(o["z_txt"]), _ptr(o["z_fuse"]), _ptr(o["logits"]),
                                     _ptr(o["probs"]), _ptr(o["vector"]), _stream()))
        return o

    def forward_host(self, images_u8, ids, pos, tt, cu, max_len, thresholds=None, out=None):
        """Whole path with HOST tensors (pinned recommended): H2D, forward, D2H, sync inside the C call."""
        B, H, W, Cc = images_u8.shape
        if out is None:
            out = (torch.empty(B, self.n_cls, dtype=torch.float32).pin_memory(),
                   torch.empty(B, self.n_cls, dtype=torch.float32).pin_memory(),
                   torch.empty(B, self.n_cls, dtype=torch.uint8).pin_memory())
        logits, probs, vec = out
        check(lib().mmdx_forward_host(self._h, _ptr(images_u8), B, H, W, Cc, _ptr(ids), _ptr(pos), _ptr(tt), _ptr(cu),
                                      ids.numel(), int(max_len), _ptr(thresholds), _ptr(logits), _ptr(probs), _ptr(vec),
                                      _stream()))
        return logits, probs, vec

    def forward_host_submit(self, slot, images_u8, ids, pos, tt, cu, max_len, out, thresholds=None):
        """Pipelined form of forward_host for throughput serving: enqueue request `slot` (0/1) and return at once;
        `out` = (logits, probs, vector) pinned host tensors that forward_host_wait(slot) fills.  With two slots in
        flight the H2D copy of one request runs under the kernels of the other."""
        B, H, W, Cc = images_u8.shape
        logits, probs, vec = out
        check(lib().mmdx_forward_host_submit(self._h, int(slot), _ptr(images_u8), B, H, W, Cc, _ptr(ids), _ptr(pos), _ptr(tt),
                                             _ptr(cu), ids.numel(), int(max_len), _ptr(thresholds), _ptr(logits), _ptr(probs),
                                             _ptr(vec), _stream()))

    def forward_host_wait(self, slot):
        check(lib().mmdx_forward_host_wait(self._h, int(slot)))


class RawHandle:
    """An engine handle without weights: enough for the single-kernel entry points (`mmdx_op_*`)."""

    def __init__(self, device: int | None = None, resize_short: int = 256, crop: int = 224, n_heads: int = 12):
        if not torch.cuda.is_available():
            raise MmdxError("mmdx needs a CUDA device (B200, sm_100a); there is no CPU fallback")
        self.device = torch.cuda.current_device() if device is None else int(device)
        cfg = Config(self.device, resize_short, crop, n_heads, (C.c_float * 3)(*IMAGENET_MEAN),
                     (C.c_float * 3)(*IMAGENET_STD))
        self._h = C.c_void_p()
        check(lib().mmdx_create(C.byref(cfg), C.byref(self._h)))

    @property
    def handle(self):
        return self._h

    def close(self):
        if self._h.value:
            lib().mmdx_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
