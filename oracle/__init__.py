"""CPU oracle for the LoRAStencil hot path -- TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package.  ``lorastencil_b200`` never does.

It wraps three things (paths relative to /root/reference):

* ``liboracle.so`` (oracle/oracle.c): the C restatement of ``test_cpu``
  (src/1d/main.cu:34-40, src/2d/main.cu:38-93, src/3d/main.cu:33-68), of the ping-pong /
  halo semantics of the ``gpu_*`` host operators (src/2d/gpu.cu:392-421 and siblings) and of
  the unseeded ``rand()`` fill (src/2d/main.cu:229-236);
* numpy restatements of the reference's hard-coded weight tables (``reference_params``) and
  of the weights its GPU operators *actually* apply for arbitrary ``params``
  (``effective_params``: the pyramidal rank-1 peel of src/2d/gpu.cu:280-350, the ignored
  ``params`` of star2d1r / star3d1r, the ``params[0..2]``-only 3-D box);
* when built (``make -C oracle ref``; needs /root/reference at build time, not at run time)
  the reference's own code under ``oracle/_ref``: ``ref_cpu(dim)`` = its verbatim
  ``test_cpu``; ``ref_gpu(dim)`` = its GPU operators recompiled for sm_100a.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from ctypes import POINTER, c_double, c_int, c_longlong

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

SHAPES_1D = ("1d1r", "1d2r")
SHAPES_2D = ("star2d1r", "box2d1r", "star2d3r", "box2d3r")
SHAPES_3D = ("box3d1r", "star3d1r")
ALL_SHAPES = SHAPES_1D + SHAPES_2D + SHAPES_3D

# halo widths per axis (S1): src/1d/main.cu:96, src/2d/main.cu:217-218, src/3d/main.cu:21-23
HALO = {1: (4,), 2: (4, 4), 3: (1, 2, 4)}

# the artifact's "fused time steps per launch" multiplier K in its GStencil/s printout
# src/1d/gpu_1r.cu:132, src/1d/gpu_2r.cu:134, src/2d/gpu.cu:419,478,553,
# src/3d/gpu_box.cu:221, src/3d/gpu_star.cu:190
ARTIFACT_K = {"1d1r": 3, "1d2r": 2, "star2d1r": 3, "box2d1r": 3, "star2d3r": 1, "box2d3r": 3,
              "box3d1r": 1, "star3d1r": 1}


def dim_of(shape: str) -> int:
    if shape in SHAPES_1D:
        return 1
    if shape in SHAPES_2D:
        return 2
    if shape in SHAPES_3D:
        return 3
    raise ValueError(f"unknown shape {shape!r}")


def build(force: bool = False) -> None:
    """Compile liboracle.so (and oracle/_ref when /root/reference is present)."""
    so, src = os.path.join(_HERE, "liboracle.so"), os.path.join(_HERE, "oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "liboracle.so"], check=True, capture_output=True)


def build_ref() -> bool:
    """Build oracle/_ref from /root/reference if it is there; True when _ref exists after."""
    if os.path.exists("/root/reference/src/2d/main.cu"):
        subprocess.run(["make", "-C", _HERE, "ref"], check=True, capture_output=True)
    return os.path.exists(os.path.join(_HERE, "_ref", "libref_cpu_2d.so"))


def lib() -> ctypes.CDLL:
    global _LIB
    if _LIB is None:
        build()
        L = ctypes.CDLL(os.path.join(_HERE, "liboracle.so"))
        dp = POINTER(c_double)
        L.oracle_fill_rand.argtypes = [dp, c_longlong, c_int]
        L.oracle_fill_rand.restype = None
        L.oracle_max_threads.restype = c_int
        L.oracle_set_threads.argtypes = [c_int]
        L.oracle_step_1d.argtypes = [dp, dp, dp, c_longlong]
        L.oracle_step_2d.argtypes = [dp, dp, dp, c_longlong, c_longlong]
        L.oracle_step_3d.argtypes = [dp, dp, dp, c_longlong, c_longlong, c_longlong]
        L.oracle_run_1d.argtypes = [dp, dp, dp, c_int, c_longlong]
        L.oracle_run_2d.argtypes = [dp, dp, dp, c_int, c_longlong, c_longlong]
        L.oracle_run_3d.argtypes = [dp, dp, dp, c_int, c_longlong, c_longlong, c_longlong]
        L.oracle_step_3d_r2.argtypes = [dp, dp, dp, c_longlong, c_longlong, c_longlong]
        L.oracle_run_3d_r2.argtypes = [dp, dp, dp, c_int, c_longlong, c_longlong, c_longlong]
        for f in (L.oracle_run_1d, L.oracle_run_2d, L.oracle_run_3d, L.oracle_run_3d_r2):
            f.restype = c_int
        _LIB = L
    return _LIB


def _p(a: np.ndarray):
    assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(POINTER(c_double))


def padded_shape(shape: str, dims) -> tuple:
    d = dim_of(shape)
    assert len(dims) == d
    return tuple(int(x) + 2 * h for x, h in zip(dims, HALO[d]))


def fill_rand(shape: str, dims) -> np.ndarray:
    """The reference's default (FILL_RANDOM) input for ``lorastencil_Xd shape dims...``."""
    ps = padded_shape(shape, dims)
    a = np.empty(ps, dtype=np.float64)
    lib().oracle_fill_rand(_p(a), a.size, 10000 if dim_of(shape) == 1 else 100)
    return a


def step(shape_or_dim, a: np.ndarray, params: np.ndarray) -> np.ndarray:
    """One direct-tap step (test_cpu): returns a zero array with the interior written."""
    d = shape_or_dim if isinstance(shape_or_dim, int) else dim_of(shape_or_dim)
    params = np.ascontiguousarray(params, dtype=np.float64)
    out = np.zeros_like(a)
    L = lib()
    if d == 1:
        L.oracle_step_1d(_p(a), _p(out), _p(params), a.shape[0])
    elif d == 2:
        L.oracle_step_2d(_p(a), _p(out), _p(params), a.shape[0], a.shape[1])
    else:
        L.oracle_step_3d(_p(a), _p(out), _p(params), a.shape[0], a.shape[1], a.shape[2])
    return out


def run(shape_or_dim, a: np.ndarray, params: np.ndarray, times: int, out: np.ndarray | None = None) -> np.ndarray:
    """``times`` launches with the gpu_* buffer semantics (S2/S3).  ``params`` are the weights
    the operator applies (use ``effective_params`` to mimic a reference GPU operator)."""
    d = shape_or_dim if isinstance(shape_or_dim, int) else dim_of(shape_or_dim)
    params = np.ascontiguousarray(params, dtype=np.float64)
    a = np.ascontiguousarray(a, dtype=np.float64)
    if out is None:
        out = np.zeros_like(a)
    L = lib()
    h = HALO[d]
    if d == 1:
        rc = L.oracle_run_1d(_p(a), _p(out), _p(params), times, a.shape[0] - 2 * h[0])
    elif d == 2:
        rc = L.oracle_run_2d(_p(a), _p(out), _p(params), times, a.shape[0] - 2 * h[0], a.shape[1] - 2 * h[1])
    else:
        rc = L.oracle_run_3d(_p(a), _p(out), _p(params), times, a.shape[0] - 2 * h[0], a.shape[1] - 2 * h[1],
                             a.shape[2] - 2 * h[2])
    if rc != 0:
        raise MemoryError("oracle work buffers")
    return out


# --------------------------------------------------------------------------------------------
# radius-2 3-D shapes (box3d2r / star3d2r): NOT in the reference; checker of the product's extension
# --------------------------------------------------------------------------------------------
HALO_R2 = (2, 2, 4)
R2_SHAPES = ("box3d2r", "star3d2r")


def padded_shape_r2(dims) -> tuple:
    return tuple(d + 2 * k for d, k in zip(dims, HALO_R2))


def reference_params_r2(shape: str) -> np.ndarray:
    """The default tables of the product for these shapes (include/lorastencil.h: lora_reference_table), restated."""
    if shape == "box3d2r":
        a = np.array([1.0, 2.0, 3.0, 2.0, 1.0])
        return np.einsum("i,j,k->ijk", a, a, a).reshape(-1)
    if shape == "star3d2r":
        w = np.zeros((5, 5, 5))
        for d, v in ((0, 3.0), (1, 2.0), (2, 1.0)):
            for ax in range(3):
                for sgn in (-1, 1):
                    idx = [2, 2, 2]
                    idx[ax] += sgn * d
                    w[tuple(idx)] = v
        return w.reshape(-1)
    raise ValueError(shape)


def step_r2(a: np.ndarray, params: np.ndarray) -> np.ndarray:
    """One direct-tap step over the interior of a (h+4, m+4, n+8) array: a zero array with the interior written."""
    a = np.ascontiguousarray(a, dtype=np.float64)
    params = np.ascontiguousarray(params, dtype=np.float64)
    assert a.ndim == 3 and params.size == 125
    out = np.zeros_like(a)
    lib().oracle_step_3d_r2(_p(a), _p(out), _p(params), a.shape[0], a.shape[1], a.shape[2])
    return out


def run_r2(a: np.ndarray, params: np.ndarray, times: int) -> np.ndarray:
    """``times`` launches with the gpu_* buffer semantics (S2/S3) on the radius-2 layout."""
    a = np.ascontiguousarray(a, dtype=np.float64)
    params = np.ascontiguousarray(params, dtype=np.float64)
    assert a.ndim == 3 and params.size == 125
    out = np.zeros_like(a)
    h, m, n = (s - 2 * k for s, k in zip(a.shape, HALO_R2))
    if lib().oracle_run_3d_r2(_p(a), _p(out), _p(params), int(times), h, m, n) != 0:
        raise MemoryError("oracle work buffers")
    return out


def wrap_ring(a: np.ndarray) -> np.ndarray:
    """The padded array whose halo ring is the periodic image of ``a``'s interior (numpy's wrap padding of the
    interior by the storage halo of S1).  The reference has no boundary update (S2); this and ``run_periodic`` are
    the checker of the product's LORA_BOUNDARY_PERIODIC mode."""
    h = HALO[a.ndim]
    inner = tuple(slice(k, -k) for k in h)
    return np.pad(a[inner], [(k, k) for k in h], mode="wrap")


def run_periodic(shape_or_dim, a: np.ndarray, params: np.ndarray, times: int) -> np.ndarray:
    """``times`` direct-tap steps (``step`` = test_cpu) on a torus: the ring is rewritten from the interior before
    every step and once more on the result; the caller's halo values are never read."""
    d = shape_or_dim if isinstance(shape_or_dim, int) else dim_of(shape_or_dim)
    cur = wrap_ring(np.ascontiguousarray(a, dtype=np.float64))
    for _ in range(times):
        cur = wrap_ring(step(d, np.ascontiguousarray(cur), params))
    return cur


# --------------------------------------------------------------------------------------------
# weight tables of the reference CLIs
# --------------------------------------------------------------------------------------------

def reference_params(shape: str) -> np.ndarray:
    """Weights the reference ``main`` passes for ``shape`` (9 / 49 / 27 doubles)."""
    if shape == "1d1r":  # src/1d/main.cu:77
        return np.array([0, 1, 2, 3, 4, 3, 2, 1, 0], dtype=np.float64)
    if shape == "1d2r":  # src/1d/main.cu:78
        return np.array([1, 2, 3, 4, 5, 4, 3, 2, 1], dtype=np.float64)
    if shape in ("box2d1r", "box2d3r"):  # src/2d/main.cu:150-174 (one table for both, :209-211)
        w = np.zeros((7, 7))
        num = 1
        for i in range(-3, 1):
            for j in range(-3, 1):
                if i <= j:
                    for a, b in ((i, j), (-i, j), (i, -j), (-i, -j), (j, i), (-j, i), (j, -i), (-j, -i)):
                        w[a + 3, b + 3] = num
                    num += 1
        w[3, 3] = 8
        return w.reshape(49)
    if shape == "star2d3r":  # src/2d/main.cu:176-184
        w = np.zeros((7, 7))
        for k, i in enumerate(range(-3, 1), start=1):
            w[i + 3, 3] = w[-i + 3, 3] = w[3, i + 3] = w[3, -i + 3] = k
        return w.reshape(49)
    if shape == "star2d1r":  # src/2d/main.cu:186-195
        return np.array([0, 0, 0, 1, 0, 0, 0,
                         0, 0, 2, 4, 2, 0, 0,
                         0, 2, 4, 8, 4, 2, 0,
                         1, 4, 8, 16, 8, 4, 1,
                         0, 2, 4, 8, 4, 2, 0,
                         0, 0, 2, 4, 2, 0, 0,
                         0, 0, 0, 1, 0, 0, 0], dtype=np.float64)
    if shape == "box3d1r":  # src/3d/main.cu:112-119
        return np.array([[1, 2, 1][i % 3] for i in range(27)], dtype=np.float64)
    if shape == "star3d1r":  # src/3d/main.cu:121-125
        return np.array([0, 0, 0, 0, 1, 0, 0, 0, 0,
                         0, 1, 0, 1, 2, 1, 0, 1, 0,
                         0, 0, 0, 0, 1, 0, 0, 0, 0], dtype=np.float64)
    raise ValueError(shape)


def reference_peel_box2d(params: np.ndarray):
    """Restatement of the host factorisation in gpu_box_2d3r (src/2d/gpu.cu:280-350).

    Returns (u, v, centre): u[t], v[t] (t = 0..2, 7 entries each) as the reference uploads
    them, plus the 1x1 remainder ``fact_param_matrix_h[3][3*7+3]`` it computes and drops."""
    P = np.asarray(params, dtype=np.float64).reshape(7, 7)
    F = np.zeros((4, 7, 7))
    T = np.zeros((3, 7, 7))
    # level 0 (:283-299)
    F[0, 0, :] = P[0, :]
    F[0, 6, :] = P[6, :]
    for r in (1, 2, 3):
        prop = P[r, 0] / P[0, 0]
        F[0, r, :] = prop * P[0, :]
        F[0, 6 - r, :] = F[0, r, :]
        T[0, r, :] = P[r, :] - F[0, r, :]
        T[0, 6 - r, :] = T[0, r, :]
    # level 1 (:300-316)
    F[1, 1, 1:6] = T[0, 1, 1:6]
    F[1, 5, 1:6] = F[1, 1, 1:6]
    for r in (2, 3):
        prop = T[0, r, 1] / T[0, 1, 1]
        F[1, r, 1:6] = prop * T[0, 1, 1:6]
        F[1, 6 - r, 1:6] = F[1, r, 1:6]
        T[1, r, 1:6] = T[0, r, 1:6] - F[1, r, 1:6]
        T[1, 6 - r, 1:6] = T[1, r, 1:6]
    # level 2 (:317-332)
    F[2, 2, 2:5] = T[1, 2, 2:5]
    F[2, 4, 2:5] = T[1, 2, 2:5]
    prop = T[1, 3, 2] / T[1, 2, 2]
    F[2, 3, 2:5] = prop * T[1, 2, 2:5]
    F[3, 3, 2:5] = T[1, 3, 2:5] - F[2, 3, 2:5]
    # vectors (:334-350)
    u = np.zeros((3, 7))
    v = np.zeros((3, 7))
    u[0, :] = F[0, 0, :]
    v[0, :] = F[0, :, 0] / F[0, 0, 0]
    u[1, 1:6] = F[1, 1, 1:6]
    v[1, 1:6] = F[1, 1:6, 1] / F[1, 1, 1]
    u[2, 2:5] = F[2, 2, 2:5]
    v[2, 2:5] = F[2, 2:5, 2] / F[2, 2, 2]
    return u, v, F[3, 3, 3]


def effective_params(shape: str, params: np.ndarray | None = None) -> np.ndarray:
    """Direct-tap weights equal to what the reference GPU operator for ``shape`` applies to
    ``params`` (quirks 2-3 of SURVEY.md appendix B).  For the reference's own tables this is
    the table itself."""
    if params is None:
        params = reference_params(shape)
    params = np.asarray(params, dtype=np.float64)
    if shape in SHAPES_1D:  # band P[r+c][c]=params[r], src/1d/gpu_1r.cu:95-99
        return params.copy()
    if shape in ("box2d1r", "box2d3r"):
        # u_t applied along rows (vertical), v_t along columns (horizontal); the 1x1 remainder
        # is not applied: src/2d/gpu.cu:358-369 (t<3), :68-101
        u, v, _ = reference_peel_box2d(params)
        w = np.zeros((7, 7))
        for t in range(3):
            w += np.outer(u[t], v[t])
        return w.reshape(49)
    if shape == "star2d3r":  # src/2d/gpu.cu:433-444
        P = params.reshape(7, 7)
        w = np.zeros((7, 7))
        w[:, 3] = P[:, 3]
        for c in range(7):
            if c != 3:
                w[3, c] = P[3, c]
        return w.reshape(49)
    if shape == "star2d1r":  # params ignored: src/2d/gpu.cu:486-487 + residual :249-264
        uv = np.array([0, 1, 2, 4, 2, 1, 0], dtype=np.float64)
        w = np.outer(uv, uv)
        for dr, dc in ((0, -3), (0, 3), (-3, 0), (3, 0)):
            w[3 + dr, 3 + dc] += 1.0
        for dr, dc in ((-2, -2), (-2, 2), (2, -2), (2, 2)):
            w[3 + dr, 3 + dc] -= 1.0
        return w.reshape(49)
    if shape == "box3d1r":  # ones (h) x ones (m) x params[0..2] (n): src/3d/gpu_box.cu:151-164
        w = np.zeros((3, 3, 3))
        w[:, :, :] = params[:3][None, None, :]
        return w.reshape(27)
    if shape == "star3d1r":  # params ignored: src/3d/gpu_star.cu:51,142-151
        w = np.zeros((3, 3, 3))
        w[0, 1, 1] = w[2, 1, 1] = 1
        w[1, 0, 1] = w[1, 2, 1] = 1
        w[1, 1, 0] = w[1, 1, 2] = 1
        w[1, 1, 1] = 2
        return w.reshape(27)
    raise ValueError(shape)


# --------------------------------------------------------------------------------------------
# the reference itself (oracle/_ref), when present
# --------------------------------------------------------------------------------------------

_REF_CPU_SYMBOL = {1: "_Z8test_cpuPdS_S_i", 2: "_Z8test_cpuPdS_S_ii", 3: "_Z8test_cpuPdS_S_iii"}
_REF_GPU_SYMBOL = {
    "1d1r": "_Z8gpu_1d1rPKdPdS0_ii", "1d2r": "_Z8gpu_1d2rPKdPdS0_ii",
    "star2d1r": "_Z13gpu_star_2d1rPKdPdS0_iii", "star2d3r": "_Z13gpu_star_2d3rPKdPdS0_iii",
    "box2d1r": "_Z12gpu_box_2d3rPKdPdS0_iii", "box2d3r": "_Z12gpu_box_2d3rPKdPdS0_iii",
    "box3d1r": "_Z12gpu_box_3d1rPKdPdS0_iiii", "star3d1r": "_Z13gpu_star_3d1rPKdPdS0_iiii",
}


def ref_available(kind: str = "cpu", dim: int = 2) -> bool:
    return os.path.exists(os.path.join(_HERE, "_ref", f"libref_{kind}_{dim}d.so"))


def ref_cpu_step(dim: int, a: np.ndarray, params: np.ndarray) -> np.ndarray:
    """The reference's verbatim test_cpu (single step) on padded array ``a``."""
    L = ctypes.CDLL(os.path.join(_HERE, "_ref", f"libref_cpu_{dim}d.so"))
    f = getattr(L, _REF_CPU_SYMBOL[dim])
    dp = POINTER(c_double)
    f.argtypes = [dp, dp, dp] + [c_int] * dim
    f.restype = None
    params = np.ascontiguousarray(params, dtype=np.float64).copy()
    a = np.ascontiguousarray(a, dtype=np.float64).copy()
    out = np.zeros_like(a)
    f(_p(a), _p(out), _p(params), *[int(s) for s in a.shape])
    return out


def ref_cpu_fn(dim: int):
    """ctypes handle of the reference test_cpu (for timing; releases the GIL)."""
    L = ctypes.CDLL(os.path.join(_HERE, "_ref", f"libref_cpu_{dim}d.so"))
    f = getattr(L, _REF_CPU_SYMBOL[dim])
    f.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p] + [c_int] * dim
    f.restype = None
    return f


def ref_gpu_run(shape: str, a: np.ndarray, params: np.ndarray, times: int) -> np.ndarray:
    """The reference GPU operator (recompiled for sm_100a) -- needs a GPU.  Prints the
    reference's banner to stdout like the original."""
    d = dim_of(shape)
    L = ctypes.CDLL(os.path.join(_HERE, "_ref", f"libref_gpu_{d}d.so"))
    f = getattr(L, _REF_GPU_SYMBOL[shape])
    dp = POINTER(c_double)
    f.argtypes = [dp, dp, dp, c_int] + [c_int] * d
    f.restype = None
    params = np.ascontiguousarray(params, dtype=np.float64).copy()
    a = np.ascontiguousarray(a, dtype=np.float64)
    out = np.zeros_like(a)
    interior = [int(s) - 2 * h for s, h in zip(a.shape, HALO[d])]
    f(_p(a), _p(out), _p(params), int(times), *interior)
    return out
