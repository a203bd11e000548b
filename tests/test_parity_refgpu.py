"""GPU parity against the REFERENCE GPU operators themselves, recompiled for sm_100a from
/root/reference by oracle/Makefile into oracle/_ref/libref_gpu_*.so (travels to the GPU box as a
prebuilt file).  This is what pins the multi-launch semantics (S2/S3), which no reference test covers.
Sizes obey the reference's implicit constraints (S4): n%1024 (1-D); m%32, n%64 (2-D); m%8, n%64 (3-D)."""
import numpy as np
import pytest

import oracle
from lorastencil_b200 import ops

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not oracle.ref_available("gpu", 2), reason="oracle/_ref GPU libraries not built")]

CASES = [("1d1r", (4096,)), ("1d2r", (8192,)), ("box2d1r", (64, 128)), ("box2d3r", (96, 64)), ("star2d1r", (64, 64)),
         ("star2d3r", (32, 192)), ("box3d1r", (8, 16, 64)), ("star3d1r", (6, 8, 128))]


@pytest.mark.parametrize("shape,dims", CASES, ids=lambda v: v if isinstance(v, str) else "x".join(map(str, v)))
def test_same_results_as_reference_gpu_operator(shape, dims, capfd):
    prev = ops.set_verbose(False)
    try:
        a = oracle.fill_rand(shape, dims)
        p = oracle.reference_params(shape)
        # integer data: both sides are exact while values stay below 2^53
        exact_upto = {"1d1r": 8, "1d2r": 8, "box2d1r": 5, "box2d3r": 5, "star2d1r": 6, "star2d3r": 9,
                      "box3d1r": 8, "star3d1r": 15}[shape]
        for times in (1, 2, 3, 5, 6, 12):
            ref = oracle.ref_gpu_run(shape, a, p, times)
            out = np.zeros_like(a)
            ops.BY_SHAPE[shape](a, out, p, times, *dims)
            if times <= exact_upto:
                assert np.array_equal(out, ref), (shape, times)
            else:
                assert np.abs(out - ref).max() <= 1e-12 * np.abs(ref).max(), (shape, times)
        rng = np.random.default_rng(2)
        b = rng.uniform(-1, 1, a.shape)
        for times in (1, 5):
            ref = oracle.ref_gpu_run(shape, b, p, times)
            out = np.zeros_like(b)
            ops.BY_SHAPE[shape](b, out, p, times, *dims)
            assert np.abs(out - ref).max() <= 1e-12 * np.abs(ref).max(), (shape, times)
    finally:
        ops.set_verbose(prev)


def test_banner_matches_reference(capfd):
    """Same stdout lines as the reference operator (src/2d/gpu.cu:415-419), numbers aside."""
    a = oracle.fill_rand("box2d1r", (64, 64))
    p = oracle.reference_params("box2d1r")
    oracle.ref_gpu_run("box2d1r", a, p, 2)
    ref_lines = [l for l in capfd.readouterr().out.splitlines() if l.strip()]
    prev = ops.set_verbose(True)
    try:
        ops.gpu_box_2d3r(a, np.zeros_like(a), p, 2, 64, 64)
    finally:
        ops.set_verbose(prev)
    our_lines = [l for l in capfd.readouterr().out.splitlines() if l.strip()]
    assert len(ref_lines) == len(our_lines) == 3
    assert our_lines[0] == ref_lines[0] == "LoRAStencil(2D box_2d3r): "
    assert our_lines[1].startswith("Time = ") and our_lines[1].endswith("[ms]")
    assert our_lines[2].startswith("GStencil/s = ")


@pytest.mark.parametrize("dim,args", [(1, ["1d1r", "10240", "3"]), (1, ["1d2r", "4096", "1"]),
                                      (2, ["box2d1r", "1024", "1024", "10"]), (2, ["box2d3r", "64", "128", "2"]),
                                      (2, ["star2d1r", "96", "64", "3"]), (2, ["star2d3r", "32", "192", "1"]),
                                      (3, ["box3d1r", "16", "16", "64", "2"]), (3, ["star3d1r", "8", "24", "128", "3"])])
def test_reference_driver_links_against_our_library_and_passes_its_own_check(dim, args):
    """The reference's UNMODIFIED main.cu (compiled with -DCHECK_ERROR by oracle/Makefile, linked against
    liblorastencil_b200.so instead of its gpu_*.cu) runs its own verification loop -- test_cpu vs one launch,
    every interior cell whose |difference| > 1e-7 is printed (src/2d/main.cu:282-328) -- and prints none."""
    import os
    import subprocess
    exe = os.path.join(os.path.dirname(oracle.__file__), "_ref", f"refmain_{dim}d")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/refmain_* not built")
    r = subprocess.run([exe, *args], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    out = r.stdout
    assert "Comparing naive and lora" in out and "Correct!" in out
    assert "naive = " not in out, out[-2000:]           # no mismatch lines
    assert out.count("GStencil/s = ") == 2              # timed run + the times=1 verification run
