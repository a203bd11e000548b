"""Multi-GPU slabs through the drop-in surface (LORA_NGPU / LORA_DEVICES -> lora_slabset_*, csrc/slab.cu): the grid is
cut along its outermost axis, every sweep is one kernel launch per slab whose band tasks store into the neighbour's
ghost zone and raise its flag.  On a 1-GPU box the slabs share device 0 (LORA_DEVICES=0,0,...): same kernels, same
mirror stores, same flag protocol -- only the wire is missing; on a multi-GPU box they sit on different GPUs.
Results must be BIT-IDENTICAL to the single-GPU operator, and match the CPU oracle."""
import os
import subprocess
import sys

import numpy as np
import pytest

import lorastencil_b200 as ls
import oracle
from lorastencil_b200 import ops

pytestmark = pytest.mark.gpu
RTOL = 1e-12


@pytest.fixture(scope="module", autouse=True)
def _quiet():
    prev = ops.set_verbose(False)
    yield
    ops.set_verbose(prev)


def _devices(k):
    import torch
    n = torch.cuda.device_count()
    return ",".join(str(i % n) for i in range(k))


CASES = [("1d2r", (1 << 20,), 47, 2), ("1d1r", (100000,), 4, 3), ("1d2r", (40000,), 16, 4), ("box2d1r", (512, 640), 6, 2),
         ("box2d3r", (300, 258), 5, 3), ("star2d3r", (300, 258), 5, 2), ("star2d1r", (256, 256), 25, 4),
         ("star2d3r", (2048, 1024), 31, 8), ("star2d1r", (90, 70), 7, 3), ("box3d1r", (64, 64, 128), 5, 2),
         ("star3d1r", (33, 40, 136), 4, 3), ("box3d1r", (96, 32, 64), 21, 8), ("star3d1r", (8, 20, 30), 6, 4),
         ("box2d3r", (120, 71), 5, 3), ("star2d3r", (64, 129), 4, 2), ("box3d1r", (12, 9, 31), 4, 3),  # odd columns
         # enough planes per slab for several plane chunks: the bands are folded into the first / last chunk
         ("box3d1r", (200, 64, 128), 5, 2), ("star3d1r", (150, 40, 136), 4, 3), ("box3d1r", (260, 32, 64), 7, 4)]


@pytest.mark.parametrize("shape,dims,times,k", CASES, ids=lambda v: v if isinstance(v, str) else str(v).replace(" ", ""))
def test_slabs_identical_to_single_gpu_and_oracle(shape, dims, times, k, monkeypatch):
    rng = np.random.default_rng(42)
    a = rng.integers(0, 100, size=oracle.padded_shape(shape, dims)).astype(np.float64)
    p = oracle.reference_params(shape)
    monkeypatch.delenv("LORA_DEVICES", raising=False)
    monkeypatch.delenv("LORA_NGPU", raising=False)
    one = np.full_like(a, -7.0)
    ops.BY_SHAPE[shape](a, one, p, times, *dims)
    assert ops.last_gpus() == 1
    monkeypatch.setenv("LORA_DEVICES", _devices(k))
    many = np.full_like(a, -7.0)
    ops.BY_SHAPE[shape](a, many, p, times, *dims)
    assert ops.last_gpus() == k
    if not np.array_equal(many, one):
        bad = np.argwhere(many != one)
        raise AssertionError((shape, dims, times, k, "first / last mismatching index", bad[0].tolist(), bad[-1].tolist(),
                              "count", len(bad), "rows", sorted(set(bad[:, 0].tolist()))[:40]))
    ref = oracle.run(shape, a, oracle.effective_params(shape, p), times)
    if oracle.dim_of(shape) == 1:
        assert many[-1] == -7.0
        many, ref = many[:-1], ref[:-1]
    scale = np.abs(ref).max()
    assert np.abs(many - ref).max() <= RTOL * scale


def test_repeated_calls_and_odd_even_launch_counts(monkeypatch):
    """The flags only ever grow: back-to-back calls (new slab sets each time) and launch counts of both parities."""
    monkeypatch.setenv("LORA_DEVICES", _devices(3))
    shape, dims = "box2d3r", (96, 128)
    a = oracle.fill_rand(shape, dims)
    p = oracle.reference_params(shape)
    for times in (0, 1, 2, 3, 4):
        out = np.full_like(a, -7.0)
        ops.BY_SHAPE[shape](a, out, p, times, *dims)
        assert ops.last_gpus() == 3
        assert np.array_equal(out, oracle.run(shape, a, oracle.effective_params(shape, p), times)), times


def test_too_thin_for_that_many_slabs_runs_on_one_gpu(monkeypatch):
    monkeypatch.setenv("LORA_DEVICES", _devices(8))
    shape, dims = "star2d3r", (40, 64)  # 8 slabs of 5 rows < 2 x 9 ghost rows
    a = oracle.fill_rand(shape, dims)
    out = np.zeros_like(a)
    ops.BY_SHAPE[shape](a, out, oracle.reference_params(shape), 3, *dims)
    assert ops.last_gpus() == 1
    assert np.array_equal(out, oracle.run(shape, a, oracle.effective_params(shape), 3))


def test_native_slab_runner_single_rank_matches_plan():
    """lorastencil_b200.slab.SlabRunner in its native mode (lora_slab_*), world size 1."""
    import torch
    from lorastencil_b200.slab import SlabRunner
    for shape, dims, times in (("1d2r", (5000,), 17), ("star2d3r", (70, 200), 7), ("box3d1r", (9, 34, 130), 4)):
        a = oracle.fill_rand(shape, dims)
        r = SlabRunner(shape, dims, device="cuda")
        assert r.halo_mode == "p2p" and r.native is not None
        r.load_global(a)
        r.run(times)
        torch.cuda.synchronize()
        got = r.gather_global(a.shape)
        r.close()
        ref = oracle.run(shape, a, oracle.effective_params(shape), times)
        if len(dims) == 1:
            got, ref = got[:-1], ref[:-1]
        assert np.abs(got - ref).max() <= RTOL * np.abs(ref).max()


@pytest.mark.parametrize("dim,args", [(1, ["1d2r", "1048576", "20"]), (2, ["box2d1r", "1024", "1024", "10"]),
                                      (2, ["star2d3r", "512", "192", "7"]), (3, ["box3d1r", "64", "32", "128", "5"])])
def test_reference_driver_on_several_gpus_passes_its_own_check(dim, args):
    """The reference's UNMODIFIED main.cu (-DCHECK_ERROR) linked on our library, with LORA_NGPU / LORA_DEVICES in the
    environment: its timed run and its verification launch both go through the slab path; its own CHECK_ERROR loop
    (test_cpu vs one launch, src/2d/main.cu:282-328) prints no mismatch."""
    import torch
    exe = os.path.join(os.path.dirname(oracle.__file__), "_ref", f"refmain_{dim}d")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/refmain_* not built")
    env = dict(os.environ)
    if torch.cuda.device_count() >= 2:
        env["LORA_NGPU"] = "2"
        env.pop("LORA_DEVICES", None)
    else:
        env["LORA_DEVICES"] = "0,0"
    r = subprocess.run([exe, *args], capture_output=True, text=True, timeout=300, env=env)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "Comparing naive and lora" in r.stdout and "Correct!" in r.stdout
    assert "naive = " not in r.stdout, r.stdout[-2000:]


def test_cli_gpus_flag():
    """bin/lorastencil_2d ... --gpus k --check"""
    import torch
    exe = os.path.join(os.path.dirname(ls.__file__), "bin", "lorastencil_2d")
    env = dict(os.environ)
    if torch.cuda.device_count() < 2:
        env["LORA_DEVICES"] = "0,0"
    r = subprocess.run([exe, "box2d1r", "512", "256", "4", "--gpus", "2", "--check"], capture_output=True, text=True,
                       timeout=300, env=env)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "Correct!" in r.stdout and "naive = " not in r.stdout


def test_real_multi_gpu_ranks_identical_to_one_gpu():
    """One process per GPU (torchrun, CUDA IPC between ranks); skipped on a 1-GPU box."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29519",
                        os.path.join(os.path.dirname(__file__), "multigpu_check.py")], capture_output=True, text=True,
                       timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
