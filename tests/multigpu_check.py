"""Multi-GPU parity check, run under torchrun with one rank per GPU:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tests/multigpu_check.py

Every rank runs its slab through lorastencil_b200.slab.SlabRunner, once per halo-exchange mode ("p2p": edge bands
stored straight into the neighbour's ghost rows over NVLink peer memory, flags in stream order; "nccl": send/recv);
rank 0 additionally runs the whole grid on its own GPU with a plain Plan and compares BITWISE.
Exit code 0 = all cases identical."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import lorastencil_b200 as ls  # noqa: E402
from lorastencil_b200.slab import SlabRunner  # noqa: E402

CASES = [("1d2r", (1 << 20,), 7), ("1d2r", (1 << 22,), 47), ("1d1r", (100000,), 4), ("box2d1r", (512, 640), 6),
         ("star2d3r", (300, 258), 5), ("star2d1r", (256, 256), 25), ("star2d3r", (2048, 1024), 31), ("box3d1r", (64, 64, 128), 5),
         ("star3d1r", (33, 40, 136), 4), ("box3d1r", (96, 32, 64), 21),
         # sweeps of two launches (pyramid / diamond forms) in the slabs: ring of buffer 1 borrowed, ghost rows of 6
         ("box2d1r", (1024, 640), 9), ("star2d1r", (600, 516), 8), ("box2d3r", (700, 258), 4)]


def main():
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    rank, world = dist.get_rank(), dist.get_world_size()
    bad = 0
    for mode, (shape, dims, times) in [(m, c) for m in ("p2p", "nccl") for c in CASES]:
        os.environ["LORA_HALO"] = mode
        runner = SlabRunner(shape, dims, device=dev)
        rng = np.random.default_rng(42)
        d = len(dims)
        padded = tuple(x + 2 * h for x, h in zip(dims, ls.plan.HALO[d]))
        a = rng.integers(0, 100, size=padded).astype(np.float64)
        runner.load_global(a)
        runner.run(times)
        torch.cuda.synchronize()
        got = runner.gather_global(a.shape)
        used = runner.halo_mode
        runner.close()
        if rank == 0:
            plan = ls.Plan(shape, dims)
            b0, b1 = torch.from_numpy(a).to(dev), plan.new_buffer(dev)
            ref = plan.run(b0, b1, times).cpu().numpy()
            ok = np.array_equal(got, ref)
            print(f"[{world} GPUs, halo {used}] {shape} {dims} x{times}: {'identical' if ok else 'MISMATCH'}", flush=True)
            bad += 0 if ok else 1
    # host-resident 1-D job without any exchange: every rank runs the drop-in operator on its slab + a margin of the
    # dependency-cone width, the slabs put together must equal the single-GPU operator call on the whole line
    from lorastencil_b200 import ops
    from lorastencil_b200.slab import host_segment, run_host_segment
    ops.set_verbose(False)
    for shape, n, times in (("1d2r", 1 << 20, 37), ("1d1r", 300000, 8)):
        rng = np.random.default_rng(7)
        a = rng.integers(0, 10, size=(n + 8,)).astype(np.float64)
        p = ls.reference_table(shape)
        lo, hi, gl, gr = host_segment(n, world, rank, times)
        seg = np.ascontiguousarray(a[lo - gl:hi + gr + 8])
        out = np.zeros_like(seg)
        run_host_segment(shape, seg, out, p, times)
        mine = torch.from_numpy(out[4 + gl:4 + gl + hi - lo].copy())
        pieces = [None] * world
        dist.all_gather_object(pieces, (lo, hi, mine.numpy()))
        if rank == 0:
            whole = np.zeros_like(a)
            ops.BY_SHAPE[shape](a, whole, p, times, n)
            got = np.concatenate([x[2] for x in sorted(pieces, key=lambda t: t[0])])
            ok = np.array_equal(got, whole[4:4 + n])
            print(f"[{world} GPUs, no exchange] {shape} ({n},) x{times} host segments with margins: {'identical' if ok else 'MISMATCH'}",
                  flush=True)
            bad += 0 if ok else 1
    flag = torch.tensor([bad], device=dev)
    dist.broadcast(flag, 0)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(1 if int(flag.item()) else 0)


if __name__ == "__main__":
    main()
