"""Drop-in `VectorQuantizer` / `ResidualVectorQuantizer` backed by libnat_b200.so.

Mirrors the reference's Python surface for this path (nat.py = /root/reference/neural_audio_tokenizer.py):
  * constructor signatures, attributes and buffer names   nat.py:1337-1356, 2099-2117
  * `ResidualVectorQuantizer.forward/encode/decode`       nat.py:1358-1446
  * `VectorQuantizer.forward/decode/_update_ema`          nat.py:2119-2221
  * ValueError conditions for bad rank / channel count    nat.py:1378-1391, 2126-2138

Contract: the tensor-core path implements the ARGMIN branch (nat.py:2155-2157). When a layer has
`use_stochastic=True` in eval mode the reference samples (nat.py:2150-2154, its default); this module then follows
`sampling_mode`:
  * "host_noise" (default): the Exp(1) draws behind `torch.multinomial(probs, 1)` are made on the host from torch's CPU
    generator, in the reference's order, and shipped to the device (nat_rvq_sample_f32) -- seeded runs reproduce the
    reference's codes except at near-ties of probs / q. Meant for the reference's own clip sizes (N*K floats per layer);
  * "philox": device-side noise, equal to the reference in distribution only (stated, never silent); distances from
    the tensor-core pass, bulk throughput. "philox_exact" keeps the exact per-frame scan with the same noise;
  * "delegate": call `stochastic_delegate` (e.g. the unmodified reference module).
Training mode (nat.py:2150, 2179-2181: the layer samples, then `_update_ema`) goes to `stochastic_delegate` when one
is set; otherwise it runs here, layer by layer: native sampling on the device, then the EMA update of
nat.py:2205-2221 in PyTorch on the same device (cold path; the codebook `copy_` it ends with bumps `_version`, so
the next layer-call re-derives the device-side codebook state). The module never returns argmin codes when sampling
was asked for. There is no CPU path: tensors must live on a CUDA device.
"""
from __future__ import annotations

import ctypes
import inspect
import types
from typing import List, Optional, Sequence

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _lib


class _CodebookPack:
    """Owns the native `nat_rvq_codebooks` handle for a list of codebook tensors and refreshes it when any of them
    changed: detected from (data_ptr, `_version`, shape, device). `tensor.copy_()` bumps `_version`; writes through
    `.data` (`codebook.data.copy_(...)`, nat.py:1926, 2073) do NOT, so every initializer of the drop-in calls
    `invalidate()` afterwards, `ResidualVectorQuantizer.invalidate_codebooks()` is the public form for callers that
    write through `.data` themselves, and `verify=True` compares the device contents on every call (one host sync)."""

    def __init__(self):
        self.handle = None
        self.signature = None
        self.device = None
        self.verify = False          # compare contents with the uploaded snapshot on every get() (debugging aid)
        self._snapshot = None

    def invalidate(self):
        """Forget what was uploaded: the next call re-derives the device-side codebook state."""
        self.signature = None

    @staticmethod
    def _sig(codebooks: Sequence[torch.Tensor]):
        return tuple((cb.data_ptr(), cb._version, tuple(cb.shape), str(cb.device)) for cb in codebooks)

    def get(self, codebooks: Sequence[torch.Tensor]):
        lib = _lib.load()
        sig = self._sig(codebooks)
        if self.handle is not None and sig == self.signature:
            if not self.verify or all(torch.equal(a, b) for a, b in zip(self._snapshot, codebooks)):
                return self.handle
        dev = codebooks[0].device
        K, D = codebooks[0].shape
        for cb in codebooks:
            if cb.device != dev or tuple(cb.shape) != (K, D) or cb.dtype != torch.float32:
                raise ValueError("all codebooks of a stack must share device, shape [K, D] and dtype float32")
        tensors = [cb if cb.is_contiguous() else cb.contiguous() for cb in codebooks]
        ptrs = (ctypes.c_void_p * len(tensors))(*[t.data_ptr() for t in tensors])
        stream = torch.cuda.current_stream(dev).cuda_stream
        same_geometry = (self.handle is not None and self.device == dev and self.signature is not None
                         and len(self.signature) == len(sig) and self.signature[0][2] == sig[0][2])
        with torch.cuda.device(dev):
            if same_geometry:
                _lib.check(lib.nat_rvq_codebooks_update(self.handle, ptrs, stream))
            else:
                self.close()
                out = ctypes.c_void_p()
                _lib.check(lib.nat_rvq_codebooks_create(ptrs, len(tensors), K, D, stream, ctypes.byref(out)))
                self.handle = out
        self.signature, self.device = sig, dev
        self._snapshot = [cb.detach().clone() for cb in codebooks] if self.verify else None
        return self.handle

    def close(self):
        if self.handle is not None:
            try:
                _lib.load().nat_rvq_codebooks_destroy(self.handle)
            except Exception:
                pass
            self.handle = None
            self.signature = None

    def __del__(self):
        self.close()


def _require_cuda(t: torch.Tensor, what: str) -> None:
    if not t.is_cuda:
        raise RuntimeError(f"{what} is on {t.device}: the B200 RVQ path has no CPU fallback; move the module and its "
                           "inputs to a CUDA device")


def _native_encode(pack: _CodebookPack, codebooks: Sequence[torch.Tensor], x_bct: torch.Tensor,
                   commitment_weight: float, want_quantized: bool, want_loss: bool, exact_scan: bool = False,
                   stats: Optional[torch.Tensor] = None, code_dtype: torch.dtype = torch.int64):
    """x_bct [B, C, T] fp32 CUDA -> (codes [L, B, T], quantized [B, C, T] or None, loss [L] or None)."""
    lib = _lib.load()
    _require_cuda(x_bct, "input")
    if x_bct.dtype != torch.float32:
        raise TypeError(f"expected float32 features, got {x_bct.dtype}")
    dev = x_bct.device
    for cb in codebooks:
        if cb.device != dev:
            raise RuntimeError(f"input is on {dev} but the codebook is on {cb.device}")
    x = x_bct if x_bct.is_contiguous() else x_bct.contiguous()
    B, C, T = x.shape
    L = len(codebooks)
    dt = {torch.int64: _lib.CODES_I64, torch.int32: _lib.CODES_I32, torch.int16: _lib.CODES_I16}[code_dtype]
    codes = torch.empty((L, B, T), dtype=code_dtype, device=dev)
    quantized = torch.empty_like(x) if want_quantized else None
    loss = torch.empty(L, dtype=torch.float32, device=dev) if want_loss else None
    if B * T == 0:
        if loss is not None:
            loss.fill_(float("nan"))           # mean over zero elements, as F.mse_loss gives
        return codes, quantized, loss
    with torch.cuda.device(dev):
        handle = pack.get(codebooks)
        ws_bytes = lib.nat_rvq_workspace_bytes(handle, B * T)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        _lib.check(lib.nat_rvq_encode_f32(
            handle, x.data_ptr(), _lib.LAYOUT_BCT, B, T, codes.data_ptr(), dt,
            quantized.data_ptr() if quantized is not None else None,
            loss.data_ptr() if loss is not None else None, float(commitment_weight),
            stats.data_ptr() if stats is not None else None, ws.data_ptr(), ws_bytes,
            _lib.RVQ_EXACT_SCAN if exact_scan else _lib.RVQ_DEFAULT,
            torch.cuda.current_stream(dev).cuda_stream))
        # `ws` may be freed when this frame returns: the caching allocator keeps it tied to the current stream.
    return codes, quantized, loss


MAX_HOST_NOISE_BYTES = 1 << 30


def _native_sample(pack: _CodebookPack, codebooks: Sequence[torch.Tensor], x_bct: torch.Tensor,
                   commitment_weight: float, want_quantized: bool, want_loss: bool, temperatures: Sequence[float],
                   mode: str, philox_draw: int = 0, code_dtype: torch.dtype = torch.int64):
    """Sampling form of _native_encode. temperatures[l] <= 0 marks an argmin layer. Returns the same triple."""
    lib = _lib.load()
    _require_cuda(x_bct, "input")
    if x_bct.dtype != torch.float32:
        raise TypeError(f"expected float32 features, got {x_bct.dtype}")
    dev = x_bct.device
    x = x_bct if x_bct.is_contiguous() else x_bct.contiguous()
    B, C, T = x.shape
    L, K = len(codebooks), codebooks[0].shape[0]
    N = B * T
    dt = {torch.int64: _lib.CODES_I64, torch.int32: _lib.CODES_I32, torch.int16: _lib.CODES_I16}[code_dtype]
    codes = torch.empty((L, B, T), dtype=code_dtype, device=dev)
    quantized = torch.empty_like(x) if want_quantized else None
    loss = torch.empty(L, dtype=torch.float32, device=dev) if want_loss else None
    noise = None
    if mode == "host_noise":
        # the device array is indexed by layer ([L, N, K], argmin layers' slices stay unused): that is what is allocated
        if N * K * 4 * L > MAX_HOST_NOISE_BYTES:
            raise RuntimeError(
                f"sampling_mode='host_noise' would ship {N * K * 4 * L / 2**30:.1f} GiB of host-drawn noise "
                f"({N} frames x {K} codes x {L} layers); use sampling_mode='philox' (distributional parity) "
                "or use_stochastic=False")
        # one draw per sampling layer, in layer order, exactly as torch.multinomial makes it on the CPU
        # (`at::empty_like(probs).exponential_(1)`); argmin layers consume nothing
        noise = torch.empty((L, N, K), dtype=torch.float32, device=dev)
        for l, t in enumerate(temperatures):
            if t > 0 and N * K > 0:
                noise[l].copy_(torch.empty((N, K), dtype=torch.float32).exponential_(1))
    elif mode not in ("philox", "philox_exact"):
        raise ValueError(f"unknown sampling_mode {mode!r}")
    if N == 0:
        if loss is not None:
            loss.fill_(float("nan"))
        return codes, quantized, loss
    temps = (ctypes.c_float * L)(*[float(t) for t in temperatures])
    with torch.cuda.device(dev):
        handle = pack.get(codebooks)
        ws_bytes = lib.nat_rvq_workspace_bytes(handle, N)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        _lib.check(lib.nat_rvq_sample_f32(
            handle, x.data_ptr(), _lib.LAYOUT_BCT, B, T, codes.data_ptr(), dt,
            quantized.data_ptr() if quantized is not None else None,
            loss.data_ptr() if loss is not None else None, float(commitment_weight), temps,
            noise.data_ptr() if noise is not None else None, int(torch.initial_seed()) & (2 ** 64 - 1),
            int(philox_draw), ws.data_ptr(), ws_bytes, _lib.RVQ_EXACT_SCAN if mode == "philox_exact" else 0,
            torch.cuda.current_stream(dev).cuda_stream))
    return codes, quantized, loss


def _native_decode(pack: _CodebookPack, codebooks: Sequence[torch.Tensor], codes: torch.Tensor, n_lists: int,
                   B: int, T: int) -> torch.Tensor:
    """codes [n_lists, B*T] int64 CUDA -> [B, D, T] fp32 (sum of per-layer gathers, nat.py:1438-1444)."""
    lib = _lib.load()
    dev = codebooks[0].device
    _require_cuda(codebooks[0], "codebook")
    K, D = codebooks[0].shape
    out = torch.empty((B, D, T), dtype=torch.float32, device=dev)
    if B * T == 0:
        return out
    if n_lists:
        # F.embedding raises IndexError on the reference path (nat.py:2195); an unchecked gather would read out of bounds
        lo, hi = torch.aminmax(codes)
        lo, hi = int(lo), int(hi)
        if lo < 0 or hi >= K:
            raise IndexError(f"code index out of range: codes span [{lo}, {hi}], the codebook has {K} entries")
    with torch.cuda.device(dev):
        handle = pack.get(codebooks)
        _lib.check(lib.nat_rvq_decode_f32(handle, codes.data_ptr() if n_lists else None, _lib.CODES_I64, n_lists, B, T,
                                          _lib.LAYOUT_BCT, out.data_ptr(),
                                          torch.cuda.current_stream(dev).cuda_stream))
    return out


class VectorQuantizer(nn.Module):
    """One VQ layer; same constructor, buffers and return values as nat.py:2092-2221."""

    def __init__(self, input_dim: int, codebook_size: int, commitment_weight: float = 0.25, ema_decay: float = 0.99,
                 temperature: float = 0.5, use_stochastic: bool = True):
        super().__init__()
        self.input_dim = input_dim
        self.codebook_size = codebook_size
        self.commitment_weight = commitment_weight
        self.ema_decay = ema_decay
        self.temperature = temperature
        self.use_stochastic = use_stochastic
        # same draw from the global generator as nat.py:2115-2117, so seeded construction yields the same codebooks
        self.register_buffer("codebook", torch.randn(codebook_size, input_dim))
        self.register_buffer("ema_count", torch.zeros(codebook_size))
        self.register_buffer("ema_weight", self.codebook.clone())
        self.stochastic_delegate = None
        self.sampling_mode = "host_noise"     # "host_noise" | "philox" | "delegate" (module docstring)
        self.exact_scan = False
        self._pack = _CodebookPack()
        self._draws = 0

    def _argmin_mode(self) -> bool:
        return not (self.training or self.use_stochastic)

    def invalidate_codebooks(self) -> None:
        """Call after writing `codebook` through `.data` (which bypasses `_version`): the next call re-uploads."""
        self._pack.invalidate()

    def forward(self, x):
        if x.dim() not in [2, 3]:
            raise ValueError(f"VectorQuantizer expects 2D or 3D input, got {x.dim()}D tensor with shape {x.shape}")
        original_shape = x.shape
        if x.dim() == 2:
            x = x.unsqueeze(0)
        B, C, T = x.shape
        if C != self.input_dim:
            raise ValueError(f"Expected {self.input_dim} feature dimensions, got {C}")
        if not self._argmin_mode():
            if self.stochastic_delegate is not None and (self.training or self.sampling_mode == "delegate"):
                return self.stochastic_delegate(x if len(original_shape) == 3 else x.squeeze(0))
            if self.sampling_mode == "delegate":
                raise NotImplementedError(
                    "VectorQuantizer: sampling_mode='delegate' needs `stochastic_delegate` (e.g. the reference module); "
                    "sampling runs natively with sampling_mode 'host_noise' or 'philox'.")
            codes, quantized, loss = _native_sample(self._pack, [self.codebook], x, self.commitment_weight, True, True,
                                                    [self.temperature], self.sampling_mode, self._draws)
            self._draws += 1
            if self.training:                                   # nat.py:2179-2181
                flat_input = x.transpose(1, 2).contiguous().view(-1, self.input_dim)
                self._update_ema(flat_input, codes[0].reshape(-1))
        else:
            codes, quantized, loss = _native_encode(self._pack, [self.codebook], x, self.commitment_weight, True, True,
                                                    exact_scan=self.exact_scan)
        codes = codes[0]
        loss = loss[0]
        if len(original_shape) == 2:
            quantized = quantized.squeeze(0)
            codes = codes.squeeze(0)
        return quantized, codes, loss

    def decode(self, codes):
        original_shape = codes.shape
        if codes.dim() == 1:
            codes = codes.unsqueeze(0)
        B, T = codes.shape
        flat = codes.reshape(1, -1).to(device=self.codebook.device, dtype=torch.int64).contiguous()
        quantized = _native_decode(self._pack, [self.codebook], flat, 1, B, T)
        if len(original_shape) == 1:
            quantized = quantized.squeeze(0)
        return quantized

    def _update_ema(self, flat_input, codes_flat):
        """EMA codebook update (nat.py:2205-2221), called by the training-mode forward; never on the tokenise path,
        kept in PyTorch. `codebook.copy_` bumps `_version`: the device-side codebook state follows."""
        with torch.no_grad():
            onehot = F.one_hot(codes_flat, self.codebook_size).float()
            self.ema_count.mul_(self.ema_decay).add_(onehot.sum(dim=0), alpha=1 - self.ema_decay)
            self.ema_weight.mul_(self.ema_decay).add_(torch.matmul(onehot.t(), flat_input), alpha=1 - self.ema_decay)
            self.codebook.copy_(self.ema_weight / (self.ema_count + 1e-5).unsqueeze(1))


class ResidualVectorQuantizer(nn.Module):
    """L chained VQ layers; same constructor, attributes and return values as nat.py:1329-1446."""

    def __init__(self, input_dim: int = 512, codebook_size: int = 4096, num_quantizers: int = 8,
                 commitment_weight: float = 0.25, ema_decay: float = 0.99, temperature: float = 0.5,
                 use_stochastic: bool = True):
        super().__init__()
        self.input_dim = input_dim
        self.codebook_size = codebook_size
        self.num_quantizers = num_quantizers
        self.commitment_weight = commitment_weight
        self.quantizers = nn.ModuleList([
            VectorQuantizer(input_dim, codebook_size, commitment_weight, ema_decay, temperature=temperature,
                            use_stochastic=use_stochastic)
            for _ in range(num_quantizers)
        ])
        self.stochastic_delegate = None      # e.g. the reference module itself (training, or sampling_mode "delegate")
        self.sampling_mode = "host_noise"    # how eval-mode sampling layers get their noise (module docstring)
        self._draws = 0
        self.codes_on_cpu = False            # one bulk D2H instead of per-element reads in the NDJSON emitter
        self.exact_scan = False              # debugging aid: exact fp64 full scan for every frame
        self.collect_stats = False
        self.last_stats = None
        self._pack = _CodebookPack()

    # -- codebook sourcing (cold path, stays with the reference) -------------------------------------------------
    # The reference fills the codebooks on the first forward of the tokenizer through methods of ITS quantizer class
    # (`initialize_from_mert_model` / `initialize_from_encodec_weights` / `initialize_from_encodec`, called at
    # nat.py:3075, 3086, 3166, 3176) and their helpers (`_validate_*`, `_prepare_encodec_features`). Those only touch
    # `self.quantizers[i].codebook / ema_*`, `self.input_dim`, `self.codebook_size`, `self.num_quantizers`, all of
    # which this class keeps, so they are run unmodified on the drop-in: `install()` / `patch_reference_module()`
    # record the reference class here and unknown attributes resolve to its functions bound to this object. The
    # initializers write through `.data` (nat.py:1926, 2073), which `_version` does not see: they are wrapped to
    # invalidate the device-side codebook state when they return.
    _reference_class = None
    _INITIALIZERS = ("initialize_from_mert_model", "initialize_from_encodec_weights", "initialize_from_encodec")

    def __getattr__(self, name):
        try:
            return super().__getattr__(name)
        except AttributeError:
            ref = self.__dict__.get("_reference_class") or type(self)._reference_class
            if ref is None or name.startswith("__") or not hasattr(ref, name):
                if name in self._INITIALIZERS:
                    raise RuntimeError(
                        f"{name}: codebook sourcing stays with the reference implementation (SURVEY.md L1b); graft this "
                        "module with install() / patch_reference_module() so that its class is known, or fill "
                        "`quantizers[i].codebook` yourself and call invalidate_codebooks()") from None
                raise
            static = inspect.getattr_static(ref, name)
            fn = getattr(ref, name)
            if isinstance(static, staticmethod) or not callable(fn):
                return fn
            bound = types.MethodType(fn, self)
            if name not in self._INITIALIZERS:
                return bound

            def initializer(*args, **kwargs):
                try:
                    return bound(*args, **kwargs)
                finally:
                    self.invalidate_codebooks()
            return initializer

    def invalidate_codebooks(self) -> None:
        """Call after writing codebooks through `.data` (which bypasses `_version`): the next call re-uploads."""
        self._pack.invalidate()
        for q in self.quantizers:
            if isinstance(q, VectorQuantizer):
                q._pack.invalidate()

    # -- helpers ---------------------------------------------------------------------------------------------
    def _codebooks(self) -> List[torch.Tensor]:
        return [q.codebook for q in self.quantizers]

    def _argmin_mode(self) -> bool:
        return all(q._argmin_mode() if isinstance(q, VectorQuantizer) else not (q.training or q.use_stochastic)
                   for q in self.quantizers)

    def _needs_delegate(self) -> bool:
        return self.sampling_mode == "delegate"

    def _training(self) -> bool:
        return any(q.training for q in self.quantizers)

    def _forward_training(self, x):
        """The layer loop of nat.py:1393-1415 with every layer in training mode: sample, update the layer's EMA
        statistics and codebook, subtract. One native call per layer (the codebooks change between layers' calls)."""
        residual = x
        quantized_layers, codes = [], []
        total_loss = 0
        for quantizer in self.quantizers:
            quantized, code, loss = quantizer(residual)
            quantized_layers.append(quantized)
            codes.append(code.cpu() if self.codes_on_cpu else code)
            total_loss = total_loss + loss
            residual = residual - quantized.detach()
        return sum(quantized_layers), codes, {"vq_loss": total_loss, "num_layers": len(quantized_layers)}

    def _sample(self, x, want_quantized: bool, want_loss: bool):
        temps = [float(q.temperature) if q.use_stochastic else 0.0 for q in self.quantizers]
        out = _native_sample(self._pack, self._codebooks(), x, self.commitment_weight, want_quantized, want_loss,
                             temps, self.sampling_mode, self._draws)
        self._draws += len(self.quantizers)
        return out

    def _validate(self, x):
        if x.dim() not in [2, 3]:
            raise ValueError(f"Expected 2D or 3D input tensor, got {x.shape}")
        if x.dim() == 2:
            x = x.unsqueeze(0)
        if x.dim() != 3:
            raise ValueError(f"Expected 3D input tensor [B, C, T], got {x.shape}")
        if x.shape[1] != self.input_dim:
            raise ValueError(f"Expected {self.input_dim} feature dimensions, got {x.shape[1]}")
        return x

    def _stats_tensor(self, dev):
        if not self.collect_stats:
            return None
        return torch.zeros((len(self.quantizers), _lib.STAT_FIELDS), dtype=torch.int64, device=dev)

    def _finish_codes(self, codes: torch.Tensor) -> List[torch.Tensor]:
        if self.codes_on_cpu:
            codes = codes.cpu()
        return [codes[l] for l in range(codes.shape[0])]

    # -- reference surface -----------------------------------------------------------------------------------
    def forward(self, x, training_mode: bool = None):
        original_training = self.training
        if training_mode is not None:
            self.train(training_mode)
        try:
            x = self._validate(x)
            if not self._argmin_mode():
                if self.stochastic_delegate is not None and (self._needs_delegate() or self._training()):
                    return self.stochastic_delegate(x)
                if self._needs_delegate():
                    raise NotImplementedError(
                        "ResidualVectorQuantizer: sampling_mode is 'delegate' and no `stochastic_delegate` (the "
                        "reference module) was supplied. Sampling runs natively with sampling_mode 'host_noise' or "
                        "'philox'; install(tokenizer, force_argmin=True) gives the argmin contract (nat.py:2155-2157).")
                if self._training():
                    for q in self.quantizers:           # layers sample with the stack's noise source
                        if isinstance(q, VectorQuantizer):
                            q.sampling_mode = self.sampling_mode
                    return self._forward_training(x)
                codes, quantized, loss = self._sample(x, True, True)
            else:
                stats = self._stats_tensor(x.device)
                codes, quantized, loss = _native_encode(self._pack, self._codebooks(), x, self.commitment_weight, True,
                                                        True, exact_scan=self.exact_scan, stats=stats)
                self.last_stats = stats
            total = loss[0]
            for l in range(1, loss.shape[0]):                # total_loss += loss, layer by layer (nat.py:1402)
                total = total + loss[l]
            losses = {"vq_loss": total, "num_layers": len(self.quantizers)}
            return quantized, self._finish_codes(codes), losses
        finally:
            if training_mode is not None:
                self.train(original_training)

    def encode(self, x):
        """Codes only (nat.py:1422-1426): skips the quantised sum and the losses the reference computes and drops."""
        with torch.no_grad():
            original_training = self.training
            self.train(False)
            try:
                x = self._validate(x)
                if not self._argmin_mode():
                    if self._needs_delegate():
                        if self.stochastic_delegate is not None:
                            return self.stochastic_delegate.encode(x)
                        raise NotImplementedError("encode(): sampling_mode is 'delegate' without a delegate; see forward()")
                    # the reference's encode() runs the whole forward and drops the rest (nat.py:1422-1426); the losses
                    # and the quantised sum draw nothing from the generator, so skipping them keeps the RNG stream
                    codes, _, _ = self._sample(x, False, False)
                    return self._finish_codes(codes)
                stats = self._stats_tensor(x.device)
                codes, _, _ = _native_encode(self._pack, self._codebooks(), x, self.commitment_weight, False, False,
                                             exact_scan=self.exact_scan, stats=stats)
                self.last_stats = stats
                return self._finish_codes(codes)
            finally:
                self.train(original_training)

    def encode_host(self, x_cpu: torch.Tensor, code_dtype: torch.dtype = torch.int16,
                    out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """End-to-end form of encode(): HOST features [B, C, T] in, HOST index streams [L, B, T] out.

        One native call (nat_rvq_encode_host_f32) streams the frames through a two-slot device arena so the H2D copy
        of chunk i+1 overlaps the kernels of chunk i, and brings the int16 codes back. Pinned inputs/outputs make the
        copies asynchronous; pageable memory works too. The module's codebooks must be on a CUDA device."""
        lib = _lib.load()
        if x_cpu.is_cuda:
            raise ValueError("encode_host() takes host tensors; use encode() for device tensors")
        x = self._validate(x_cpu)
        if x.dtype != torch.float32:
            raise TypeError(f"expected float32 features, got {x.dtype}")
        if not self._argmin_mode():
            raise NotImplementedError("encode_host(): a layer is in sampling mode; see forward()")
        x = x if x.is_contiguous() else x.contiguous()
        B, C, T = x.shape
        L = len(self.quantizers)
        dev = self.quantizers[0].codebook.device
        _require_cuda(self.quantizers[0].codebook, "codebook")
        dt = {torch.int64: _lib.CODES_I64, torch.int32: _lib.CODES_I32, torch.int16: _lib.CODES_I16}[code_dtype]
        if out is None:
            out = torch.empty((L, B, T), dtype=code_dtype, pin_memory=True)
        if B * T:
            with torch.cuda.device(dev):
                handle = self._pack.get(self._codebooks())
                _lib.check(lib.nat_rvq_encode_host_f32(handle, x.data_ptr(), _lib.LAYOUT_BCT, B, T, out.data_ptr(), dt,
                                                       torch.cuda.current_stream(dev).cuda_stream))
        return out

    def decode(self, codes):
        if not codes:
            return torch.zeros(1, self.input_dim, 1)
        B, T = codes[0].shape
        dev = self.quantizers[0].codebook.device
        used = codes[:len(self.quantizers)]
        stacked = torch.stack([c.reshape(-1).to(device=dev, dtype=torch.int64) for c in used]).contiguous()
        return _native_decode(self._pack, self._codebooks(), stacked, len(used), B, T)
