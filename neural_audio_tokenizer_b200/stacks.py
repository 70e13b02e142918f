"""The quantiser stacks of one tokenizer (S0-S3 semantic, A0-A3 acoustic) encoded in one native call.

`NeuralAudioTokenizer.forward` calls `self.semantic_quantizer(...)` and `self.acoustic_quantizer(...)` one after the
other (nat.py:3239-3240). When only the index streams are wanted (the tokenise path, `encode`, nat.py:1422-1426) the
two stacks are independent pieces of work over the same frame range: `encode_stacks` hands both to
`nat_rvq_encode_stacks_f32`, which shares the layer-0 preparation between stacks fed the same tensor and runs both in
one persistent launch per chunk of frames; `encode_stacks_host` is the end-to-end form on HOST buffers
(`nat_tokenize_host_f32`: every chunk is uploaded once, all stacks run on it, the index streams of all stacks come
back as one array). Argmin contract only (every layer `use_stochastic=False`, eval mode): sampling and training go
through the per-stack modules.
"""
from __future__ import annotations

import ctypes
from typing import Optional, Sequence, Union

import torch

from . import _lib
from .quantizers import ResidualVectorQuantizer, _require_cuda

_CODE_DTYPES = {torch.int64: _lib.CODES_I64, torch.int32: _lib.CODES_I32, torch.int16: _lib.CODES_I16}


def _check_stacks(stacks: Sequence[ResidualVectorQuantizer]):
    if not 1 <= len(stacks) <= 2:
        raise ValueError(f"expected one or two stacks (semantic, acoustic), got {len(stacks)}")
    for s in stacks:
        if not s._argmin_mode():
            raise NotImplementedError("encode_stacks(): a layer samples (use_stochastic=True or training); the argmin "
                                      "contract needs use_stochastic=False on every layer (install(force_argmin=True))")
    dev = stacks[0].quantizers[0].codebook.device
    _require_cuda(stacks[0].quantizers[0].codebook, "codebook")
    for s in stacks:
        if s.quantizers[0].codebook.device != dev:
            raise RuntimeError("all stacks must live on one device")
    return dev


def _handles(stacks):
    hs = [s._pack.get(s._codebooks()) for s in stacks]
    return hs, (ctypes.c_void_p * len(hs))(*[h.value if isinstance(h, ctypes.c_void_p) else h for h in hs])


def encode_stacks(stacks: Sequence[ResidualVectorQuantizer], x: Union[torch.Tensor, Sequence[torch.Tensor]],
                  code_dtype: torch.dtype = torch.int16, out: Optional[torch.Tensor] = None,
                  workspace: Optional[torch.Tensor] = None) -> torch.Tensor:
    """x: one [B, C, T] CUDA tensor for all stacks, or one per stack (same B; streams of different length are aligned
    to the shortest one by linear interpolation first, as nat.py:3225-3236 does). Returns [sum L, B, T] codes."""
    lib = _lib.load()
    dev = _check_stacks(stacks)
    xs = [x] * len(stacks) if isinstance(x, torch.Tensor) else list(x)
    if len(xs) != len(stacks):
        raise ValueError(f"{len(stacks)} stacks but {len(xs)} inputs")
    cont = []
    for i, (s, t) in enumerate(zip(stacks, xs)):
        prior = next((j for j in range(i) if xs[j] is t), None)
        if prior is not None:                       # the same tensor for several stacks: one layer-0 preparation
            s._validate(t)
            cont.append(cont[prior])
            continue
        t = s._validate(t)
        _require_cuda(t, "input")
        if t.dtype != torch.float32:
            raise TypeError(f"expected float32 features, got {t.dtype}")
        cont.append(t if t.is_contiguous() else t.contiguous())
    B = cont[0].shape[0]
    t_ins = [int(t.shape[2]) for t in cont]
    T = min(t_ins)                                   # the common time base, nat.py:3227
    for t in cont:
        if t.device != dev or t.shape[0] != B:
            raise ValueError("inputs must share the batch extent and the stacks' device")
    L_total = sum(len(s.quantizers) for s in stacks)
    if out is None:
        out = torch.empty((L_total, B, T), dtype=code_dtype, device=dev)
    elif out.dtype != code_dtype or tuple(out.shape) != (L_total, B, T) or not out.is_contiguous() or out.device != dev:
        raise ValueError(f"out must be a contiguous {code_dtype} tensor of shape {(L_total, B, T)} on {dev}")
    if B * T == 0:
        return out
    with torch.cuda.device(dev):
        _, harr = _handles(stacks)
        need = lib.nat_rvq_stacks_workspace_bytes(harr, len(stacks), B * T)
        ws = workspace if workspace is not None and workspace.numel() >= need else torch.empty(need, dtype=torch.uint8, device=dev)
        stream = torch.cuda.current_stream(dev).cuda_stream
        if any(t != T for t in t_ins):
            # Streams of different length are brought to the shorter time base first (nat.py:3225-3236). Where the
            # stacks take the shared-preparation form the interpolation is folded into the preparation's loads;
            # elsewhere it is a kernel of its own. Either way the reference's CPU arithmetic, bit for bit.
            if lib.nat_rvq_stacks_fused(harr, len(stacks), B * T):
                xarr = (ctypes.c_void_p * len(cont))(*[t.data_ptr() for t in cont])
                tarr = (ctypes.c_int64 * len(cont))(*t_ins)
                _lib.check(lib.nat_rvq_encode_stacks_aligned_f32(harr, len(stacks), xarr, tarr, B, T, out.data_ptr(),
                                                                 _CODE_DTYPES[code_dtype], ws.data_ptr(), ws.numel(), 0,
                                                                 stream))
                return out
            from .align import interpolate_linear
            cont = [t if t.shape[2] == T else interpolate_linear(t, T) for t in cont]
        xarr = (ctypes.c_void_p * len(cont))(*[t.data_ptr() for t in cont])
        _lib.check(lib.nat_rvq_encode_stacks_f32(harr, len(stacks), xarr, _lib.LAYOUT_BCT, B, T, out.data_ptr(),
                                                 _CODE_DTYPES[code_dtype], ws.data_ptr(), ws.numel(), 0, stream))
    return out


class HostContext:
    """Owner of a `nat_host_ctx`: the device staging arena, copy stream and events of the host-buffer calls. One per
    concurrent caller (it serves one call at a time); nothing of it lives on the shared codebook handles."""

    def __init__(self, device: Union[torch.device, int, str]):
        self.device = torch.device(device)
        self._h = ctypes.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(_lib.load().nat_host_ctx_create(ctypes.byref(self._h)))

    def close(self):
        if self._h:
            try:
                _lib.load().nat_host_ctx_destroy(self._h)
            except Exception:
                pass
            self._h = ctypes.c_void_p()

    def __del__(self):
        self.close()


def encode_stacks_host(stacks: Sequence[ResidualVectorQuantizer], x_cpu: torch.Tensor,
                       code_dtype: torch.dtype = torch.int16, out: Optional[torch.Tensor] = None,
                       ctx: Optional[HostContext] = None) -> torch.Tensor:
    """HOST features [B, C, T] in (pinned memory makes the copies asynchronous), HOST index streams [sum L, B, T] out:
    the end-to-end call. All stacks quantise the same frames; every chunk crosses PCIe once."""
    lib = _lib.load()
    dev = _check_stacks(stacks)
    if x_cpu.is_cuda:
        raise ValueError("encode_stacks_host() takes host tensors; use encode_stacks() for device tensors")
    x = stacks[0]._validate(x_cpu)
    for s in stacks[1:]:
        s._validate(x_cpu)
    if x.dtype != torch.float32:
        raise TypeError(f"expected float32 features, got {x.dtype}")
    x = x if x.is_contiguous() else x.contiguous()
    B, _, T = x.shape
    L_total = sum(len(s.quantizers) for s in stacks)
    if out is None:
        out = torch.empty((L_total, B, T), dtype=code_dtype, pin_memory=True)
    elif out.is_cuda or out.dtype != code_dtype or tuple(out.shape) != (L_total, B, T) or not out.is_contiguous():
        raise ValueError(f"out must be a contiguous host {code_dtype} tensor of shape {(L_total, B, T)}")
    if B * T == 0:
        return out
    own = ctx is None
    if own:
        ctx = HostContext(dev)
    try:
        with torch.cuda.device(dev):
            _, harr = _handles(stacks)
            _lib.check(lib.nat_tokenize_host_f32(ctx._h, harr, len(stacks), x.data_ptr(), _lib.LAYOUT_BCT, B, T,
                                                 out.data_ptr(), _CODE_DTYPES[code_dtype],
                                                 torch.cuda.current_stream(dev).cuda_stream))
    finally:
        if own:
            ctx.close()
    return out
