"""The reference's host operators, by name, on halo-padded host arrays.

Mirrors ``src/1d/1d_utils.h:45-47``, ``src/2d/2d_utils.h:47-51`` and ``src/3d/3d_utils.h:44-48`` of
zondie17/LoRAStencil: ``gpu_X(in, out, params, times, dims...)``.  ``in`` / ``out`` are C-contiguous
float64 numpy arrays of the PADDED size (1-D ``n+8``; 2-D ``(m+8, n+8)``; 3-D ``(h+2, m+4, n+8)``) or
pinned torch CPU tensors of that size; ``out`` receives the whole padded buffer ``times % 2``.  Each call
goes through the C ABI (``lora_gpu_*``): H2D copy, ``times`` kernel launches on the GPU, D2H copy.
"""
from __future__ import annotations

import ctypes
from ctypes import POINTER, c_double, c_longlong

import numpy as np

from . import _lib


def _ptr(a):
    if isinstance(a, np.ndarray):
        if a.dtype != np.float64 or not a.flags["C_CONTIGUOUS"]:
            raise TypeError("expected a C-contiguous float64 array")
        return ctypes.c_void_p(a.ctypes.data), a.size
    # torch CPU tensor (possibly pinned)
    if a.dtype.__str__() != "torch.float64" or not a.is_contiguous() or a.device.type != "cpu":
        raise TypeError("expected a contiguous float64 CPU tensor")
    return ctypes.c_void_p(a.data_ptr()), a.numel()


def _params(params, n):
    p = np.ascontiguousarray(np.asarray(params, dtype=np.float64).reshape(-1))
    if p.size != n:
        raise ValueError(f"expected {n} weights, got {p.size}")
    return p


def _call(name, nparams, padded, in_, out, params, times, dims):
    L = _lib.lib()
    pin, nin = _ptr(in_)
    pout, nout = _ptr(out)
    need = int(np.prod(padded))
    if nin < need or nout < need:
        raise ValueError(f"{name}: arrays must hold the padded grid of {need} doubles")
    p = _params(params, nparams)
    getattr(L, "lora_" + name)(pin, pout, p.ctypes.data_as(POINTER(c_double)), int(times), *[int(d) for d in dims])
    return out


def gpu_1d1r(in_, out, params, times, input_n):
    return _call("gpu_1d1r", 9, (input_n + 8,), in_, out, params, times, (input_n,))


def gpu_1d2r(in_, out, params, times, input_n):
    return _call("gpu_1d2r", 9, (input_n + 8,), in_, out, params, times, (input_n,))


def gpu_star_2d1r(in_, out, params, times, input_m, input_n):
    return _call("gpu_star_2d1r", 49, (input_m + 8, input_n + 8), in_, out, params, times, (input_m, input_n))


def gpu_star_2d3r(in_, out, params, times, input_m, input_n):
    return _call("gpu_star_2d3r", 49, (input_m + 8, input_n + 8), in_, out, params, times, (input_m, input_n))


def gpu_box_2d3r(in_, out, params, times, input_m, input_n):
    return _call("gpu_box_2d3r", 49, (input_m + 8, input_n + 8), in_, out, params, times, (input_m, input_n))


def gpu_box_3d1r(in_, out, params, times, input_h, input_m, input_n):
    return _call("gpu_box_3d1r", 27, (input_h + 2, input_m + 4, input_n + 8), in_, out, params, times,
                 (input_h, input_m, input_n))


def gpu_star_3d1r(in_, out, params, times, input_h, input_m, input_n):
    return _call("gpu_star_3d1r", 27, (input_h + 2, input_m + 4, input_n + 8), in_, out, params, times,
                 (input_h, input_m, input_n))


def gpu_box_3d2r(in_, out, params, times, input_h, input_m, input_n):
    """Radius-2 box (125 weights) -- not in the reference; arrays of (h+4, m+4, n+8) doubles."""
    return _call("gpu_box_3d2r", 125, (input_h + 4, input_m + 4, input_n + 8), in_, out, params, times,
                 (input_h, input_m, input_n))


def gpu_star_3d2r(in_, out, params, times, input_h, input_m, input_n):
    """Radius-2 13-point star (125 weights, off-axis entries zero) -- not in the reference; (h+4, m+4, n+8) doubles."""
    return _call("gpu_star_3d2r", 125, (input_h + 4, input_m + 4, input_n + 8), in_, out, params, times,
                 (input_h, input_m, input_n))


# CLI shape name -> operator, as dispatched by the reference drivers
# (src/1d/main.cu:126-133, src/2d/main.cu:268-280, src/3d/main.cu:192-199)
BY_SHAPE = {"1d1r": gpu_1d1r, "1d2r": gpu_1d2r, "star2d1r": gpu_star_2d1r, "star2d3r": gpu_star_2d3r,
            "box2d1r": gpu_box_2d3r, "box2d3r": gpu_box_2d3r, "box3d1r": gpu_box_3d1r, "star3d1r": gpu_star_3d1r,
            "box3d2r": gpu_box_3d2r, "star3d2r": gpu_star_3d2r}  # the last two: new, not reference operators


def run_host(shape: str, in_, out, params, times: int, dims, mode: int = _lib.WEIGHTS_REFERENCE):
    """``lora_gpu_run_host``: any shape by CLI name, with a weight mode."""
    L = _lib.lib()
    sid = _lib.SHAPE_IDS[shape]
    pin, _ = _ptr(in_)
    pout, _ = _ptr(out)
    p = _params(params, _lib.nparams(shape))
    d = (c_longlong * 3)(*[int(x) for x in dims], *([0] * (3 - len(dims))))
    L.lora_gpu_run_host(sid, int(mode), pin, pout, p.ctypes.data_as(POINTER(c_double)), int(times), d)
    return out


def set_verbose(on: bool) -> bool:
    return bool(_lib.lib().lora_set_verbose(1 if on else 0))


def last_loop_ms() -> float:
    return float(_lib.lib().lora_last_loop_ms())


def last_total_ms() -> float:
    return float(_lib.lib().lora_last_total_ms())


def last_chunks() -> int:
    """Chunks the last 1-D drop-in call was cut into to overlap its copies with its launches (1 = none)."""
    return int(_lib.lib().lora_last_chunks())


def set_gpus(k: int) -> int:
    """Spread every following drop-in call over k GPUs of this process (slabs along the outermost axis, ghost zones
    exchanged inside the kernels over NVLink); same as the environment variable LORA_NGPU.  Returns the previous value."""
    return int(_lib.lib().lora_set_gpus(int(k)))


def last_gpus() -> int:
    """GPUs (slabs) the last drop-in call actually ran on."""
    return int(_lib.lib().lora_last_gpus())


def release_workspace() -> None:
    """Free the device buffers the drop-in operators cache between calls."""
    _lib.lib().lora_release_workspace()


def last_bands() -> int:
    """Time-skewed bands the last drop-in call was cut into to overlap its copies with its launches (1 = none)."""
    return int(_lib.lib().lora_last_bands())
