"""Device-resident plans (layer 2 of include/lorastencil.h) on torch CUDA tensors."""
from __future__ import annotations

import ctypes
from ctypes import POINTER, byref, c_double, c_longlong, c_void_p

import numpy as np

from . import _lib

HALO = {1: (4,), 2: (4, 4), 3: (1, 2, 4)}  # S1: src/1d/main.cu:96, src/2d/main.cu:217-218, src/3d/main.cu:21-23
MAX_TB_1D = 15     # deepest temporal block of the 1-D kernel (kMaxTb1 in csrc/kernels.h)
DEFAULT_TB_1D = 15  # kDefaultTb1
BOUNDARY_NAMES = ["reference", "dirichlet", "zero", "periodic"]
FORM_NAMES = {0: "taps9", 1: "cross", 2: "pyramid", 3: "diamond", 4: "direct49", 5: "sep3", 6: "star7", 7: "direct27",
              8: "pyramid_pruned", 9: "rank2", 10: "rank3", 11: "star13", 12: "hsep5", 13: "direct125", 14: "sep5"}


def _dp(a: np.ndarray):
    return a.ctypes.data_as(POINTER(c_double))


def reference_table(shape: str) -> np.ndarray:
    """The weight table the reference CLI passes for ``shape``."""
    sid = _lib.SHAPE_IDS[shape]
    out = np.zeros(_lib.nparams(shape), dtype=np.float64)
    _lib.check(_lib.lib().lora_reference_table(sid, _dp(out)), "lora_reference_table")
    return out


def effective_weights(shape: str, mode: int = _lib.WEIGHTS_REFERENCE, params=None) -> np.ndarray:
    """Direct-tap weights a plan built from (shape, mode, params) applies."""
    sid = _lib.SHAPE_IDS[shape]
    out = np.zeros(_lib.nparams(shape), dtype=np.float64)
    p = None if params is None else np.ascontiguousarray(np.asarray(params, dtype=np.float64).reshape(-1))
    _lib.check(_lib.lib().lora_effective_weights(sid, int(mode), None if p is None else _dp(p), _dp(out)),
               "lora_effective_weights")
    return out


def decompose_2d(shape: str, params, mode: int = _lib.WEIGHTS_GENERAL) -> dict:
    """Host low-rank decomposition of a 7x7 table (layer 3)."""
    d = _lib.Decomp2D()
    p = np.ascontiguousarray(np.asarray(params, dtype=np.float64).reshape(-1))
    _lib.check(_lib.lib().lora_decompose_2d(_lib.SHAPE_IDS[shape], int(mode), _dp(p), byref(d)), "lora_decompose_2d")
    return {"form": FORM_NAMES[d.form], "nterms": d.nterms,
            "vert": np.array([[d.vert[t][k] for k in range(7)] for t in range(3)]),
            "horiz": np.array([[d.horiz[t][k] for k in range(7)] for t in range(3)]),
            "centre": d.centre, "residual": np.array(list(d.residual)), "recon_err": d.recon_err,
            "macs_per_cell": d.macs_per_cell}


def decompose_3d_r2(shape: str, params) -> dict:
    """Structure the host finds in a 125-weight table of a radius-2 3-D shape (layer 3)."""
    d = _lib.Decomp3DR2()
    p = np.ascontiguousarray(np.asarray(params, dtype=np.float64).reshape(-1))
    if p.size != 125:
        raise ValueError(f"expected 125 weights, got {p.size}")
    _lib.check(_lib.lib().lora_decompose_3d_r2(_lib.SHAPE_IDS[shape], _dp(p), byref(d)), "lora_decompose_3d_r2")
    return {"form": FORM_NAMES[d.form], "a": np.array(list(d.a)), "b": np.array(list(d.b)), "c": np.array(list(d.c)),
            "q": np.array(list(d.q)).reshape(5, 5), "recon_err": d.recon_err, "macs_per_cell": d.macs_per_cell}


class Plan:
    """A device-resident stencil plan: weights factored on the host once, launches on CUDA tensors.

    ``dims`` are the interior sizes of the grid this device holds.  Buffers are float64 CUDA tensors
    with the padded shape ``padded_shape`` (allocate with ``new_buffer``)."""

    def __init__(self, shape: str, dims, params=None, mode: int = _lib.WEIGHTS_REFERENCE):
        self.shape = shape
        self.dims = tuple(int(d) for d in dims)
        self.dim = len(self.dims)
        if len(_lib.halo_of(shape)) != self.dim:
            raise ValueError(f"{shape} takes {len(_lib.halo_of(shape))} sizes, got {self.dims}")
        self.padded_shape = tuple(d + 2 * h for d, h in zip(self.dims, _lib.halo_of(shape)))
        self._h = c_void_p()
        p = None if params is None else np.ascontiguousarray(np.asarray(params, dtype=np.float64).reshape(-1))
        d = (c_longlong * 3)(*self.dims, *([0] * (3 - self.dim)))
        _lib.check(_lib.lib().lora_plan_create(byref(self._h), _lib.SHAPE_IDS[shape], int(mode),
                                               None if p is None else _dp(p), d), "lora_plan_create")

    @classmethod
    def borrowed(cls, handle, shape: str, dims):
        """A view of a plan that something else owns (a native slab's plan): never destroyed from here."""
        self = cls.__new__(cls)
        self.shape = shape
        self.dims = tuple(int(d) for d in dims)
        self.dim = len(self.dims)
        self.padded_shape = tuple(d + 2 * h for d, h in zip(self.dims, _lib.halo_of(shape)))
        self._h = c_void_p(handle)
        self._borrowed = True
        return self

    def __del__(self):
        try:
            if self._h and not getattr(self, "_borrowed", False):
                _lib.lib().lora_plan_destroy(self._h)
                self._h = c_void_p()
        except Exception:
            pass

    @property
    def describe(self) -> str:
        return _lib.lib().lora_plan_describe(self._h).decode()

    @property
    def launches(self) -> int:
        return int(_lib.lib().lora_plan_launch_count(self._h))

    @property
    def temporal_block(self) -> int:
        """Deepest temporal block `run` fuses (1 = one kernel launch per time step)."""
        return int(_lib.lib().lora_plan_temporal_block(self._h))

    @temporal_block.setter
    def temporal_block(self, tb: int):
        _lib.check(_lib.lib().lora_plan_set_temporal_block(self._h, int(tb)), "lora_plan_set_temporal_block")

    @property
    def boundary(self) -> str:
        """'reference' (the reference's alternating caller's / zero halo, S2), 'dirichlet' (the caller's halo is the
        boundary condition of every launch), 'zero' (zero halo for every launch) or 'periodic' (the grid is a torus:
        `run` rewrites the halo ring from the interior before every launch and on the result; one launch per step)."""
        return BOUNDARY_NAMES[int(_lib.lib().lora_plan_boundary(self._h))]

    @boundary.setter
    def boundary(self, mode: str):
        _lib.check(_lib.lib().lora_plan_set_boundary(self._h, BOUNDARY_NAMES.index(mode)), "lora_plan_set_boundary")

    def wrap_ring(self, buf, stream=None):
        """Halo ring of ``buf`` <- the periodic image of its interior (what `run` does before every launch of a
        periodic plan), for callers that drive `step` themselves.  Asynchronous."""
        import torch
        self._check_buf(buf)
        s = torch.cuda.current_stream(buf.device) if stream is None else stream
        _lib.check(_lib.lib().lora_plan_wrap_ring(self._h, c_void_p(buf.data_ptr()), c_void_p(s.cuda_stream)),
                   "lora_plan_wrap_ring")

    def step_fused(self, src, dst, halo_src, lo, hi, tb, launches_before, virt_lo, virt_hi, stream=None, mirror=None):
        """One fused launch of `tb` time steps: see lora_plan_step_fused in include/lorastencil.h.  ``mirror``: device
        address (int) in a neighbour's buffer corresponding to dst[0] -- the launch stores there too."""
        import torch
        s = torch.cuda.current_stream(src.device) if stream is None else stream
        _lib.check(_lib.lib().lora_plan_step_fused_mirror(
            self._h, c_void_p(src.data_ptr()), c_void_p(dst.data_ptr()),
            c_void_p(halo_src.data_ptr()) if halo_src is not None else None, int(lo), int(hi), int(tb),
            int(launches_before), int(bool(virt_lo)), int(bool(virt_hi)), c_void_p(mirror) if mirror else None,
            c_void_p(s.cuda_stream)), "lora_plan_step_fused")

    @property
    def cells(self) -> int:
        return int(np.prod(self.dims))

    def new_buffer(self, device="cuda", zero: bool = True):
        import torch
        return (torch.zeros if zero else torch.empty)(self.padded_shape, dtype=torch.float64, device=device)

    def _check_buf(self, t):
        if not t.is_cuda or t.dtype.__str__() != "torch.float64" or not t.is_contiguous() or \
                tuple(t.shape) != self.padded_shape:
            raise TypeError(f"expected a contiguous float64 CUDA tensor of shape {self.padded_shape}")

    def step(self, src, dst, lo: int = 0, hi: int | None = None, stream=None, mirror=None):
        """One launch: dst[interior, outermost index in [lo, hi)] = stencil(src).  Asynchronous.  ``mirror``: device
        address (int) in a neighbour's buffer corresponding to dst[0] -- the launch stores there too."""
        import torch
        self._check_buf(src)
        self._check_buf(dst)
        hi = self.dims[0] if hi is None else hi
        s = torch.cuda.current_stream(src.device) if stream is None else stream
        _lib.check(_lib.lib().lora_plan_step_mirror(self._h, c_void_p(src.data_ptr()), c_void_p(dst.data_ptr()), int(lo),
                                                    int(hi), c_void_p(mirror) if mirror else None,
                                                    c_void_p(s.cuda_stream)), "lora_plan_step")

    def run(self, buf0, buf1, times: int, stream=None):
        """``times`` launches, launch i reads buf[i%2]; returns the tensor holding the result."""
        import torch
        self._check_buf(buf0)
        self._check_buf(buf1)
        s = torch.cuda.current_stream(buf0.device) if stream is None else stream
        _lib.check(_lib.lib().lora_plan_run(self._h, c_void_p(buf0.data_ptr()), c_void_p(buf1.data_ptr()), int(times),
                                            c_void_p(s.cuda_stream)), "lora_plan_run")
        return buf0 if times % 2 == 0 else buf1
