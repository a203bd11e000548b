"""ctypes binding of include/lorastencil.h.  Fails loudly when the CUDA library is missing."""
from __future__ import annotations

import ctypes
import os
import subprocess
from ctypes import POINTER, Structure, c_char_p, c_double, c_int, c_longlong, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

SHAPES = ("1d1r", "1d2r", "star2d1r", "box2d1r", "star2d3r", "box2d3r", "box3d1r", "star3d1r")  # the reference's
R2_SHAPES = ("box3d2r", "star3d2r")  # radius-2 3-D shapes (new): own layout (h+4, m+4, n+8), 125 weights
SHAPE_IDS = {name: i for i, name in enumerate(SHAPES + R2_SHAPES)}  # == lora_shape_t


def nparams(shape: str) -> int:
    """Length of the weight table of ``shape``: 9 (1-D), 49 (2-D), 27 (3-D), 125 (3-D radius 2)."""
    sid = SHAPE_IDS[shape]
    return 9 if sid < 2 else (49 if sid < 6 else (27 if sid < 8 else 125))


def halo_of(shape: str) -> tuple:
    """Storage halo per axis (S1: src/1d/main.cu:96, src/2d/main.cu:217-218, src/3d/main.cu:21-23; radius-2: ours)."""
    sid = SHAPE_IDS[shape]
    return (4,) if sid < 2 else ((4, 4) if sid < 6 else ((1, 2, 4) if sid < 8 else (2, 2, 4)))


WEIGHTS_REFERENCE = 0
WEIGHTS_GENERAL = 1

# every symbol include/lorastencil.h declares (tests/test_abi.py checks the library exports them all)
C_ABI_SYMBOLS = (
    "lora_gpu_1d1r", "lora_gpu_1d2r", "lora_gpu_star_2d1r", "lora_gpu_star_2d3r", "lora_gpu_box_2d3r",
    "lora_gpu_box_3d1r", "lora_gpu_star_3d1r", "lora_gpu_box_3d2r", "lora_gpu_star_3d2r", "lora_gpu_run_host", "lora_set_verbose", "lora_last_loop_ms",
    "lora_last_total_ms", "lora_last_chunks", "lora_last_bands", "lora_release_workspace", "lora_plan_create", "lora_plan_destroy",
    "lora_plan_padded_elems", "lora_plan_step", "lora_plan_run", "lora_plan_set_temporal_block",
    "lora_plan_temporal_block", "lora_plan_set_boundary", "lora_plan_boundary", "lora_plan_wrap_ring", "lora_plan_step_fused", "lora_plan_step_mirror", "lora_plan_step_fused_mirror",
    "lora_peer_alloc", "lora_peer_free", "lora_peer_open", "lora_peer_close", "lora_stream_write_flag",
    "lora_stream_wait_flag_geq", "lora_debug_temporal_schedule", "lora_debug_pair_schedule", "lora_debug_tasks_2dtb", "lora_debug_tasks_2dtb_pairs", "lora_debug_wrap_ring_host", "lora_debug_r2_grid", "lora_debug_tb2_probe", "lora_plan_launch_count", "lora_plan_describe",
    "lora_last_error", "lora_decompose_2d", "lora_decompose_3d_r2", "lora_reference_table", "lora_effective_weights",
    "lora_set_gpus", "lora_last_gpus",
    "lora_slab_create", "lora_slab_destroy", "lora_slab_info", "lora_slab_buffer", "lora_slab_export",
    "lora_slab_connect_ipc", "lora_slab_connect_local", "lora_slab_reset", "lora_slab_sweep", "lora_slab_run",
    "lora_slab_schedule", "lora_slab_result_index", "lora_slab_launch_count", "lora_slab_plan", "lora_slab_geometry",
    "lora_slabset_create", "lora_slabset_destroy", "lora_slabset_load", "lora_slabset_run", "lora_slabset_sync",
    "lora_slabset_store", "lora_slabset_launch_count", "lora_slabset_temporal_block",
)
# the reference's own C++ symbols (include/lorastencil_dropin.hpp)
CXX_DROPIN_SYMBOLS = (
    "_Z8gpu_1d1rPKdPdS0_ii", "_Z8gpu_1d2rPKdPdS0_ii", "_Z13gpu_star_2d1rPKdPdS0_iii",
    "_Z13gpu_star_2d3rPKdPdS0_iii", "_Z12gpu_box_2d3rPKdPdS0_iii", "_Z12gpu_box_3d1rPKdPdS0_iiii",
    "_Z13gpu_star_3d1rPKdPdS0_iiii",
)


class LoraError(RuntimeError):
    pass


class Decomp2D(Structure):
    _fields_ = [("form", c_int), ("nterms", c_int), ("vert", c_double * 7 * 3), ("horiz", c_double * 7 * 3),
                ("centre", c_double), ("residual", c_double * 8), ("recon_err", c_double), ("macs_per_cell", c_int)]


class Decomp3DR2(Structure):
    _fields_ = [("form", c_int), ("a", c_double * 5), ("b", c_double * 5), ("c", c_double * 5), ("q", c_double * 25),
                ("recon_err", c_double), ("macs_per_cell", c_int)]


def lib_path() -> str:
    """The in-tree library; LORASTENCIL_LIB may name another build of it (kernel tuning experiments)."""
    return os.environ.get("LORASTENCIL_LIB") or os.path.join(_HERE, "lib", "liblorastencil_b200.so")


def library_built() -> bool:
    return os.path.exists(lib_path())


def build(force: bool = False) -> None:
    """Compile the CUDA library and the CLI drivers for sm_100a (nvcc cross-compiles without a GPU)."""
    if force:
        subprocess.run(["make", "-C", os.path.join(_HERE, "csrc"), "clean"], check=True, capture_output=True)
    r = subprocess.run(["make", "-C", os.path.join(_HERE, "csrc"), "-j8"], capture_output=True, text=True)
    if r.returncode != 0:
        raise LoraError("building liblorastencil_b200.so failed:\n" + r.stdout[-4000:] + r.stderr[-4000:])


def lib() -> ctypes.CDLL:
    global _LIB
    if _LIB is not None:
        return _LIB
    if not library_built():
        raise LoraError(f"{lib_path()} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                        "(there is no CPU fallback)")
    L = ctypes.CDLL(lib_path())
    dp = POINTER(c_double)
    for name, nd in (("lora_gpu_1d1r", 1), ("lora_gpu_1d2r", 1), ("lora_gpu_star_2d1r", 2), ("lora_gpu_star_2d3r", 2),
                     ("lora_gpu_box_2d3r", 2), ("lora_gpu_box_3d1r", 3), ("lora_gpu_star_3d1r", 3),
                     ("lora_gpu_box_3d2r", 3), ("lora_gpu_star_3d2r", 3)):
        f = getattr(L, name)
        f.argtypes = [c_void_p, c_void_p, dp, c_int] + [c_int] * nd
        f.restype = None
    L.lora_gpu_run_host.argtypes = [c_int, c_int, c_void_p, c_void_p, dp, c_int, POINTER(c_longlong)]
    L.lora_gpu_run_host.restype = None
    L.lora_set_verbose.argtypes = [c_int]
    L.lora_set_verbose.restype = c_int
    L.lora_last_loop_ms.restype = c_double
    L.lora_last_total_ms.restype = c_double
    L.lora_last_chunks.restype = c_int
    L.lora_last_bands.restype = c_int
    L.lora_release_workspace.restype = None
    L.lora_plan_create.argtypes = [POINTER(c_void_p), c_int, c_int, dp, POINTER(c_longlong)]
    L.lora_plan_create.restype = c_int
    L.lora_plan_destroy.argtypes = [c_void_p]
    L.lora_plan_destroy.restype = None
    L.lora_plan_padded_elems.argtypes = [c_void_p]
    L.lora_plan_padded_elems.restype = c_longlong
    L.lora_plan_step.argtypes = [c_void_p, c_void_p, c_void_p, c_longlong, c_longlong, c_void_p]
    L.lora_plan_step.restype = c_int
    L.lora_plan_run.argtypes = [c_void_p, c_void_p, c_void_p, c_int, c_void_p]
    L.lora_plan_run.restype = c_int
    L.lora_plan_set_temporal_block.argtypes = [c_void_p, c_int]
    L.lora_plan_set_temporal_block.restype = c_int
    L.lora_plan_temporal_block.argtypes = [c_void_p]
    L.lora_plan_temporal_block.restype = c_int
    L.lora_plan_set_boundary.argtypes = [c_void_p, c_int]
    L.lora_plan_set_boundary.restype = c_int
    L.lora_plan_boundary.argtypes = [c_void_p]
    L.lora_plan_boundary.restype = c_int
    L.lora_plan_wrap_ring.argtypes = [c_void_p, c_void_p, c_void_p]
    L.lora_plan_wrap_ring.restype = c_int
    L.lora_plan_step_fused.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_longlong, c_longlong, c_int, c_int,
                                       c_int, c_int, c_void_p]
    L.lora_plan_step_fused.restype = c_int
    L.lora_plan_step_mirror.argtypes = [c_void_p, c_void_p, c_void_p, c_longlong, c_longlong, c_void_p, c_void_p]
    L.lora_plan_step_mirror.restype = c_int
    L.lora_plan_step_fused_mirror.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_longlong, c_longlong, c_int, c_int,
                                              c_int, c_int, c_void_p, c_void_p]
    L.lora_plan_step_fused_mirror.restype = c_int
    L.lora_peer_alloc.argtypes = [POINTER(c_void_p), ctypes.c_ulonglong, c_void_p]
    L.lora_peer_alloc.restype = c_int
    L.lora_peer_free.argtypes = [c_void_p]
    L.lora_peer_free.restype = c_int
    L.lora_peer_open.argtypes = [c_void_p, POINTER(c_void_p)]
    L.lora_peer_open.restype = c_int
    L.lora_peer_close.argtypes = [c_void_p]
    L.lora_peer_close.restype = c_int
    L.lora_stream_write_flag.argtypes = [c_void_p, c_void_p, ctypes.c_ulonglong]
    L.lora_stream_write_flag.restype = c_int
    L.lora_stream_wait_flag_geq.argtypes = [c_void_p, c_void_p, ctypes.c_ulonglong]
    L.lora_stream_wait_flag_geq.restype = c_int
    L.lora_debug_temporal_schedule.argtypes = [c_int, c_int, POINTER(c_int), c_int]
    L.lora_debug_temporal_schedule.restype = c_int
    L.lora_debug_pair_schedule.argtypes = [c_int, POINTER(c_int), c_int]
    L.lora_debug_pair_schedule.restype = c_int
    L.lora_debug_tb2_probe.argtypes = [POINTER(c_double), POINTER(c_double)]
    L.lora_debug_tb2_probe.restype = c_int
    L.lora_debug_tasks_2dtb.argtypes = [c_int, c_int, c_int, c_int, c_int, POINTER(c_int), c_int]
    L.lora_debug_tasks_2dtb.restype = c_int
    L.lora_debug_tasks_2dtb_pairs.argtypes = [c_int, c_int, c_int, c_int, c_int, POINTER(c_int), c_int]
    L.lora_debug_tasks_2dtb_pairs.restype = c_int
    L.lora_debug_wrap_ring_host.argtypes = [c_int, POINTER(c_longlong), POINTER(c_double)]
    L.lora_debug_wrap_ring_host.restype = c_int
    L.lora_debug_r2_grid.argtypes = [c_int, c_int, c_longlong, c_int, c_int, c_int, POINTER(c_longlong)]
    L.lora_debug_r2_grid.restype = c_int
    L.lora_decompose_3d_r2.argtypes = [c_int, dp, POINTER(Decomp3DR2)]
    L.lora_decompose_3d_r2.restype = c_int
    L.lora_plan_launch_count.argtypes = [c_void_p]
    L.lora_plan_launch_count.restype = c_longlong
    L.lora_plan_describe.argtypes = [c_void_p]
    L.lora_plan_describe.restype = c_char_p
    L.lora_last_error.restype = c_char_p
    L.lora_decompose_2d.argtypes = [c_int, c_int, dp, POINTER(Decomp2D)]
    L.lora_decompose_2d.restype = c_int
    L.lora_reference_table.argtypes = [c_int, dp]
    L.lora_reference_table.restype = c_int
    L.lora_effective_weights.argtypes = [c_int, c_int, dp, dp]
    L.lora_effective_weights.restype = c_int
    L.lora_set_gpus.argtypes = [c_int]
    L.lora_set_gpus.restype = c_int
    L.lora_last_gpus.restype = c_int
    llp = POINTER(c_longlong)
    L.lora_slab_create.argtypes = [POINTER(c_void_p), c_int, c_int, dp, llp, c_int, c_int, c_int]
    L.lora_slab_create.restype = c_int
    L.lora_slab_destroy.argtypes = [c_void_p]
    L.lora_slab_destroy.restype = None
    L.lora_slab_info.argtypes = [c_void_p, llp]
    L.lora_slab_info.restype = c_int
    L.lora_slab_buffer.argtypes = [c_void_p, c_int]
    L.lora_slab_buffer.restype = c_void_p
    L.lora_slab_export.argtypes = [c_void_p, c_void_p]
    L.lora_slab_export.restype = c_int
    L.lora_slab_connect_ipc.argtypes = [c_void_p, c_int, c_void_p]
    L.lora_slab_connect_ipc.restype = c_int
    L.lora_slab_connect_local.argtypes = [c_void_p, c_int, c_void_p]
    L.lora_slab_connect_local.restype = c_int
    L.lora_slab_reset.argtypes = [c_void_p]
    L.lora_slab_reset.restype = c_int
    L.lora_slab_sweep.argtypes = [c_void_p, c_int, c_void_p]
    L.lora_slab_sweep.restype = c_int
    L.lora_slab_run.argtypes = [c_void_p, c_int, c_void_p]
    L.lora_slab_run.restype = c_int
    L.lora_slab_schedule.argtypes = [c_void_p, c_int, POINTER(c_int), c_int]
    L.lora_slab_schedule.restype = c_int
    L.lora_slab_result_index.argtypes = [c_void_p]
    L.lora_slab_result_index.restype = c_int
    L.lora_slab_launch_count.argtypes = [c_void_p]
    L.lora_slab_launch_count.restype = c_longlong
    L.lora_slab_plan.argtypes = [c_void_p]
    L.lora_slab_plan.restype = c_void_p
    L.lora_slab_geometry.argtypes = [c_int, llp, c_int, c_int, c_longlong, llp]
    L.lora_slab_geometry.restype = c_int
    L.lora_slabset_create.argtypes = [POINTER(c_void_p), c_int, c_int, dp, llp, c_int, POINTER(c_int)]
    L.lora_slabset_create.restype = c_int
    L.lora_slabset_destroy.argtypes = [c_void_p]
    L.lora_slabset_destroy.restype = None
    L.lora_slabset_load.argtypes = [c_void_p, c_void_p]
    L.lora_slabset_load.restype = c_int
    L.lora_slabset_run.argtypes = [c_void_p, c_int]
    L.lora_slabset_run.restype = c_int
    L.lora_slabset_sync.argtypes = [c_void_p]
    L.lora_slabset_sync.restype = c_int
    L.lora_slabset_store.argtypes = [c_void_p, c_void_p]
    L.lora_slabset_store.restype = c_int
    L.lora_slabset_launch_count.argtypes = [c_void_p]
    L.lora_slabset_launch_count.restype = c_longlong
    L.lora_slabset_temporal_block.argtypes = [c_void_p]
    L.lora_slabset_temporal_block.restype = c_int
    _LIB = L
    return L


def check(rc: int, what: str) -> None:
    if rc != 0:
        raise LoraError(f"{what}: error {rc}: {lib().lora_last_error().decode()}")
