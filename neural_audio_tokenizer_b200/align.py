"""Time-base alignment of the two feature streams before quantisation: the host mirror of
nat.py:3225-3236, `F.interpolate(x, size=T_target, mode='linear', align_corners=False)` on `[B, C, T]`.

On a CUDA tensor torch would run its own CUDA upsampling kernel, whose rounding differs from the CPU reference's;
`interpolate_linear` reproduces the reference's CPU arithmetic on the device (nat_interp_linear_f32), so the features
entering the quantisers -- and hence the token streams -- are the ones the CPU reference would have produced.
"""
from __future__ import annotations

import types
from typing import Tuple

import torch

from . import _lib


def interpolate_linear(x: torch.Tensor, size: int) -> torch.Tensor:
    """F.interpolate(x, size=size, mode='linear', align_corners=False) for a 3-D fp32 CUDA tensor."""
    if x.dim() != 3:
        raise ValueError(f"linear interpolation expects [B, C, T], got {tuple(x.shape)}")
    if not x.is_cuda:
        raise RuntimeError(f"input is on {x.device}: the B200 alignment path has no CPU fallback")
    if x.dtype != torch.float32:
        raise TypeError(f"expected float32 features, got {x.dtype}")
    B, C, T = x.shape
    size = int(size)
    if T < 1 or size < 1:
        raise ValueError("Input and output sizes should be greater than 0")       # torch's own condition
    xc = x if x.is_contiguous() else x.contiguous()
    out = torch.empty((B, C, size), dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        _lib.check(_lib.load().nat_interp_linear_f32(xc.data_ptr(), B * C, T, size, out.data_ptr(),
                                                     torch.cuda.current_stream(x.device).cuda_stream))
    return out


def align_time_bases(semantic_features: torch.Tensor, acoustic_features: torch.Tensor
                     ) -> Tuple[torch.Tensor, torch.Tensor]:
    """nat.py:3225-3236: bring both streams to min(T_sem, T_acc) frames."""
    t_target = min(semantic_features.shape[-1], acoustic_features.shape[-1])
    if semantic_features.shape[-1] != t_target:
        semantic_features = interpolate_linear(semantic_features, t_target)
    if acoustic_features.shape[-1] != t_target:
        acoustic_features = interpolate_linear(acoustic_features, t_target)
    return semantic_features, acoustic_features


def install(nat_module) -> None:
    """Rebind the name `F` inside the imported reference module to a namespace whose `interpolate` routes the call at
    nat.py:3230-3236 (3-D fp32 CUDA input, mode='linear', align_corners=False, `size=` given) to the device kernel
    and leaves every other call with torch's own function."""
    F = nat_module.F
    torch_interpolate = F.interpolate

    def interpolate(input, size=None, scale_factor=None, mode="nearest", align_corners=None, **kw):
        if (mode == "linear" and align_corners is False and scale_factor is None and size is not None and not kw and
                torch.is_tensor(input) and input.is_cuda and input.dim() == 3 and input.dtype == torch.float32):
            return interpolate_linear(input, size[0] if isinstance(size, (tuple, list)) else size)
        return torch_interpolate(input, size=size, scale_factor=scale_factor, mode=mode, align_corners=align_corners, **kw)

    shim = types.SimpleNamespace(**{k: getattr(F, k) for k in dir(F) if not k.startswith("__")})
    shim.interpolate = interpolate
    nat_module.F = shim
