"""``torch.ops.lorastencil.{stencil1d, stencil2d, stencil3d}`` -- the stencil operators on device-resident tensors.

The reference has no Python surface; its host operators are ``gpu_X(in, out, params, times, dims...)`` on padded HOST
arrays (src/1d/1d_utils.h:45-47, src/2d/2d_utils.h:47-51, src/3d/3d_utils.h:44-48).  These ops are the same operators
for callers whose grids already live in HBM:

    out = torch.ops.lorastencil.stencil2d(x, "box2d3r", times, params=None, mode=0, boundary=0)

``x``: contiguous float64 CUDA tensor of the PADDED shape (1-D ``n+8``; 2-D ``(m+8, n+8)``; 3-D ``(h+2, m+4, n+8)``;
the radius-2 extensions ``box3d2r`` / ``star3d2r``: ``(h+4, m+4, n+8)`` and 125 weights), halo included.  Returns a new tensor of the same shape holding the whole padded buffer ``times % 2`` of the reference's
ping-pong (S2/S3: halo = the caller's for even ``times``, zero for odd).  ``params``: 9 / 49 / 27 weights (None = the
reference CLI's table for the shape); ``mode``: 0 = what the reference GPU operator does with ``params``, 1 = every
weight honoured; ``boundary``: 0 = the reference's ping-pong halo (S2), 1 = Dirichlet (the caller's halo at every
launch), 2 = zero halo, 3 = periodic (include/lorastencil.h: LORA_BOUNDARY_*).  Registered through ``torch.library`` (dispatch key CUDA + a Meta kernel for shape propagation); there
is no CPU kernel: a CPU tensor raises, like everything else in this package.
"""
from __future__ import annotations

from typing import Optional

import numpy as np
import torch

from . import _lib
from .plan import BOUNDARY_NAMES, HALO, Plan

_PLANS: dict = {}
_DIM_SHAPES = {1: ("1d1r", "1d2r"), 2: ("star2d1r", "box2d1r", "star2d3r", "box2d3r"),
               3: ("box3d1r", "star3d1r", "box3d2r", "star3d2r")}


def _plan_for(x: torch.Tensor, dim: int, shape: str, params: Optional[torch.Tensor], mode: int) -> Plan:
    if shape not in _DIM_SHAPES[dim]:
        raise ValueError(f"stencil{dim}d: shape must be one of {_DIM_SHAPES[dim]}, got {shape!r}")
    if x.dim() != dim or x.dtype != torch.float64 or not x.is_contiguous():
        raise TypeError(f"stencil{dim}d: expected a contiguous float64 tensor with {dim} dimension(s) (padded grid)")
    dims = tuple(int(s) - 2 * h for s, h in zip(x.shape, _lib.halo_of(shape)))
    if min(dims) < 1:
        raise ValueError(f"stencil{dim}d: padded shape {tuple(x.shape)} leaves no interior")
    p = None if params is None else np.ascontiguousarray(params.detach().cpu().numpy().astype(np.float64).reshape(-1))
    key = (shape, dims, None if p is None else p.tobytes(), int(mode), x.device.index)
    plan = _PLANS.get(key)
    if plan is None:
        if len(_PLANS) >= 64:
            _PLANS.clear()
        with torch.cuda.device(x.device):
            plan = Plan(shape, dims, params=p, mode=int(mode))
        _PLANS[key] = plan
    return plan


def _run(x: torch.Tensor, dim: int, shape: str, times: int, params: Optional[torch.Tensor], mode: int,
         boundary: int = 0) -> torch.Tensor:
    if times < 0:
        raise ValueError("times must be >= 0")
    if not 0 <= int(boundary) < len(BOUNDARY_NAMES):
        raise ValueError(f"boundary must be 0..{len(BOUNDARY_NAMES) - 1} ({', '.join(BOUNDARY_NAMES)})")
    plan = _plan_for(x, dim, shape, params, mode)
    plan.boundary = BOUNDARY_NAMES[int(boundary)]  # plans are cached per (shape, size, weights): set it on every call
    with torch.cuda.device(x.device):
        b0 = x.clone()            # the operator never writes its input (the reference's `in` is const)
        b1 = torch.zeros_like(x)  # S2: the second ping-pong buffer starts as zeros
        res = plan.run(b0, b1, int(times))
    return res


_lib_def = torch.library.Library("lorastencil", "DEF")
for _d in (1, 2, 3):
    _lib_def.define(f"stencil{_d}d(Tensor x, str shape, int times, Tensor? params=None, int mode=0, int boundary=0) -> Tensor")


def _make_cuda(dim):
    def impl(x, shape, times, params=None, mode=0, boundary=0):
        return _run(x, dim, shape, times, params, mode, boundary)
    return impl


def _meta(x, shape, times, params=None, mode=0, boundary=0):
    return torch.empty_like(x)


def _make_cpu(dim):
    def impl(x, shape, times, params=None, mode=0, boundary=0):
        raise _lib.LoraError(f"torch.ops.lorastencil.stencil{dim}d has no CPU kernel: pass a CUDA tensor "
                             "(this package has no CPU fallback)")
    return impl


for _d in (1, 2, 3):
    _lib_def.impl(f"stencil{_d}d", _make_cuda(_d), "CUDA")
    _lib_def.impl(f"stencil{_d}d", _meta, "Meta")
    _lib_def.impl(f"stencil{_d}d", _make_cpu(_d), "CPU")
