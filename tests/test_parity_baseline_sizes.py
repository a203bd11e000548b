"""GPU parity at the sizes BASELINE.json names (configs b-e), on the paths that ship: the drop-in operators (chunked 1-D,
fused 2-D star forms, unfused box / 3-D) against the REFERENCE GPU operators recompiled unmodified for sm_100a
(oracle/_ref/libref_gpu_*.so; the reference's own check is one launch vs test_cpu, src/2d/main.cu:302-326 -- here 1, 3, 4
and `exact_upto` launches, full padded output, np.array_equal) plus a CPU-oracle band.  Every case runs in its own
process (tests/baseline_size_check.py): the reference operators leak their device buffers."""
import json
import os
import subprocess
import sys

import pytest

import oracle

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))

# (shape, dims, launch counts): 1, 3, 4 launches and the last count at which integer data stays exact
CASES = [
    ("1d2r", (1 << 28,), (1, 3, 4, 8)),                 # config b (chunked, copy-overlapped, 15-deep sweeps cut to 8)
    ("1d1r", (1 << 28,), (1, 8)),
    ("star2d3r", (10240, 10240), (1, 3, 4, 9)),         # config c (3 launches fused per sweep)
    ("box2d3r", (10240, 10240), (1, 3, 4, 5)),
    ("star2d1r", (10240, 10240), (1, 3, 4, 6)),
    ("box3d1r", (512, 512, 512), (1, 3, 4, 8)),         # config d
    ("star3d1r", (512, 512, 512), (1, 3, 4, 15)),
    ("box2d1r", (40960, 40960), (1, 4)),                # config e on one GPU (13.4 GB per buffer, 64-bit indexing)
    ("box3d1r", (1024, 1024, 1024), (1, 4)),
]
# the same grids as slabs (LORA_DEVICES: on a 1-GPU box the slabs share device 0 and still run the whole exchange
# protocol -- band tasks, mirror stores, flags; on a multi-GPU box they sit on different GPUs)
SLAB_CASES = [
    ("1d2r", (1 << 26,), (16, 31), 4),
    ("star2d3r", (10240, 10240), (4, 7), 4),
    ("box2d1r", (10240, 10240), (4,), 4),
    ("box3d1r", (512, 512, 512), (4,), 4),
    ("star3d1r", (512, 512, 512), (5,), 3),
]


def _need_host_gb(dims):
    import psutil
    n = 1
    for d in dims:
        n *= d
    need = 3.5 * n * 8 / 2**30 + 4
    have = psutil.virtual_memory().available / 2**30
    if have < need:
        pytest.skip(f"needs {need:.0f} GB of host memory, {have:.0f} GB available")


def _run(shape, dims, times, extra=()):
    cmd = [sys.executable, os.path.join(HERE, "baseline_size_check.py"), shape, ",".join(map(str, dims)),
           ",".join(map(str, times)), *extra]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert lines, r.stdout[-2000:] + r.stderr[-2000:]
    rep = json.loads(lines[-1])
    assert r.returncode == 0 and rep["ok"], rep
    return rep


@pytest.mark.parametrize("shape,dims,times", CASES, ids=lambda v: v if isinstance(v, str) else "x".join(map(str, v)))
def test_baseline_size_equals_reference_gpu_operator(shape, dims, times):
    if not oracle.ref_available("gpu", oracle.dim_of(shape)):
        pytest.skip("oracle/_ref GPU libraries not built")
    _need_host_gb(dims)
    rep = _run(shape, dims, times, ("--band", "16" if len(dims) < 3 else "4"))
    assert rep["reference_gpu"]
    for c in rep["cases"]:
        assert c["gpus"] == 1
        if c["times"] <= {"1d1r": 8, "1d2r": 8, "box2d1r": 5, "box2d3r": 5, "star2d1r": 6, "star2d3r": 9, "box3d1r": 8,
                          "star3d1r": 15}[shape]:
            assert c["vs_reference_gpu"] == "bit-identical", c


@pytest.mark.parametrize("shape,dims,times,k", SLAB_CASES, ids=lambda v: v if isinstance(v, str) else str(v))
def test_slabs_at_scale_equal_reference_gpu_operator(shape, dims, times, k):
    """The multi-GPU path (lora_gpu_* under LORA_DEVICES / LORA_NGPU: slabs + in-kernel ghost exchange) at large sizes,
    against the reference GPU operator."""
    import torch
    if not oracle.ref_available("gpu", oracle.dim_of(shape)):
        pytest.skip("oracle/_ref GPU libraries not built")
    _need_host_gb(dims)
    ndev = torch.cuda.device_count()
    devices = ",".join(str(i % ndev) for i in range(k))
    rep = _run(shape, dims, times, ("--devices", devices))
    for c in rep["cases"]:
        assert c["gpus"] == k, c
