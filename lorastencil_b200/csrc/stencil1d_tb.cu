// stencil1d_tb.cu -- 1-D 9-tap stencil with TEMPORAL BLOCKING: TB launches of the reference's
// kernel_1d1r / kernel_1d2r (src/1d/gpu_1r.cu:21-87) fused into one sweep, intermediate time levels
// living in registers only.  New functionality: the reference has no temporal blocking (it only counts
// its composite 9-tap kernels as 3 / 2 steps, src/1d/gpu_1r.cu:132).
//
// Mapping (one warp = one independent worker, as in stencil1d.cu):
//   * the line is cut into rows of 256 cells, lane l owns 8 consecutive cells of a row;
//   * level s (the grid after s of the TB launches) is held SKEWED: row j of level s covers cells
//     [B + 256 j - 4 s, +256).  With that skew the 16-cell window a lane needs from level s-1 is exactly
//     lane l-1's 8 cells followed by its own 8 cells, so one level costs 8 FP64 warp shuffles (16 SHFL.32)
//     + 72 DFMA per lane, no shared-memory round trip, and a row never needs data from the row after it;
//   * lane 0 takes lane 31's cells of the previous row from a 64-byte per-level mailbox in shared memory;
//   * level 0 arrives through the warp's private TMA ring (cp.async.bulk, 2 rows = 4 KB per stage); level TB
//     leaves through a 2 KB staging row and a TMA store (cp.async.bulk shared -> global), row-aligned because
//     B = s0 + 4 TB makes the final skew vanish;
//   * a lane's 8 cells are 64 contiguous bytes, i.e. a 64-byte lane stride in shared memory, which is a 4-way
//     bank conflict for plain 128-bit accesses.  Lanes therefore touch their four 16-byte pieces in a rotated
//     order (piece (k + lane/2) mod 4 in instruction k): every quarter-warp then covers all 8 bank groups, the
//     access is conflict-free, and a 2-level select network un-rotates the registers (ALU pipe, which is idle).
//     The LSU data pipe is this kernel's bottleneck (profiles/), so wavefronts are what is being saved.
//
// Reference semantics (S2) under fusion: launch i of the reference sees the caller's
// halo when i is even and zeros when i is odd.  Inside a fused sweep the halo cells of every level are
// therefore VIRTUAL: (time of the level even) ? caller's halo (read from `halo_src`, the padded buffer 0
// whose halo is never written) : 0.  Only rows touching an end of the global line take that path.
#include "common.cuh"
#include "kernels.h"

namespace lora {

namespace {

constexpr int kTbRow = 256;                 // cells per row
constexpr int kTbStageRows = 2;             // rows per bulk copy
constexpr int kTbStage = kTbStageRows * kTbRow;  // 512 doubles <= kStageElems

// piece order used by lane `lane` for its k-th 128-bit access to a 64-byte chunk: conflict-free in shared memory
__device__ __forceinline__ int rot_of(int lane) { return (lane >> 1) & 3; }

// out[j] = in[(j - rot) & 3] for 16-byte pieces (2-level barrel of selects)
__device__ __forceinline__ void unrotate(const double2 (&in)[4], int rot, double2 (&out)[4]) {
    double2 a[4];
#pragma unroll
    for (int j = 0; j < 4; j++) {
        a[j].x = (rot & 1) ? in[(j + 3) & 3].x : in[j].x;
        a[j].y = (rot & 1) ? in[(j + 3) & 3].y : in[j].y;
    }
#pragma unroll
    for (int j = 0; j < 4; j++) {
        out[j].x = (rot & 2) ? a[(j + 2) & 3].x : a[j].x;
        out[j].y = (rot & 2) ? a[(j + 2) & 3].y : a[j].y;
    }
}

// lane's 8 cells (64 contiguous bytes at chunk) -> registers, conflict-free
__device__ __forceinline__ void load_rotated(const double *chunk, int lane, double (&cur)[8]) {
    const int rot = rot_of(lane);
    double2 r[4], pc[4];
#pragma unroll
    for (int k = 0; k < 4; k++) r[k] = *reinterpret_cast<const double2 *>(chunk + 2 * ((k + rot) & 3));
    unrotate(r, rot, pc);  // r[k] holds piece (k + rot) & 3  =>  piece j = r[(j - rot) & 3]
#pragma unroll
    for (int j = 0; j < 4; j++) {
        cur[2 * j] = pc[j].x;
        cur[2 * j + 1] = pc[j].y;
    }
}

// registers -> lane's 64-byte chunk of the staging row, conflict-free
__device__ __forceinline__ void store_rotated(double *chunk, int lane, const double (&cur)[8]) {
    const int rot = rot_of(lane);
    double2 pc[4], q[4];
#pragma unroll
    for (int j = 0; j < 4; j++) pc[j] = make_double2(cur[2 * j], cur[2 * j + 1]);
    // instruction k writes piece (k + rot) & 3: q[k] = pc[(k + rot) & 3] = pc[(k - (4 - rot)) & 3]
    unrotate(pc, (4 - rot) & 3, q);
#pragma unroll
    for (int k = 0; k < 4; k++) *reinterpret_cast<double2 *>(chunk + 2 * ((k + rot) & 3)) = q[k];
}

// virtual halo of one level: cells -4..-1 and n..n+3 take (level time even ? caller's halo : 0)
__device__ __forceinline__ void fix_halo(double (&v)[8], long long p, int level, const Geom1DTB &g) {
    const bool use_h = ((g.par0 + level) & 1) == 0;
#pragma unroll
    for (int q = 0; q < 8; q++) {
        const long long x = p + q;
        const bool in_left = g.virt_left && x >= -4 && x < 0;
        const bool in_right = g.virt_right && x >= g.n && x < g.n + 4;
        if (in_left || in_right) v[q] = use_h ? g.halo_src[x + 4] : 0.0;
    }
}

// one row through all TB levels.  `p` = interior coordinate of this lane's first level-0 cell.
template <int TB, bool FIX>
__device__ __forceinline__ void sweep_row(const double *rowp, double *mailbox, int i, int lane, long long p,
                                          const Geom1DTB &g, const Weights1D &w, double (&cur)[8]) {
    load_rotated(rowp, lane, cur);
    if (FIX) fix_halo(cur, p, 0, g);
#pragma unroll
    for (int s = 1; s <= TB; s++) {
        // mailbox[level][row parity][8]: lane 31 posts its cells for lane 0 of the next row
        double *mb_wr = mailbox + ((s - 1) * 2 + (i & 1)) * 8;
        const double *mb_rd = mailbox + ((s - 1) * 2 + ((i + 1) & 1)) * 8;
        if (lane == 31) {
#pragma unroll
            for (int q = 0; q < 8; q += 2) *reinterpret_cast<double2 *>(mb_wr + q) = make_double2(cur[q], cur[q + 1]);
        }
        double win[16];
#pragma unroll
        for (int q = 0; q < 8; q += 2) {
            const double2 m = *reinterpret_cast<const double2 *>(mb_rd + q);  // broadcast read, used by lane 0 only
            const double a = __shfl_up_sync(0xffffffffu, cur[q], 1);
            const double b = __shfl_up_sync(0xffffffffu, cur[q + 1], 1);
            win[q] = lane == 0 ? m.x : a;
            win[q + 1] = lane == 0 ? m.y : b;
        }
#pragma unroll
        for (int q = 0; q < 8; q++) win[8 + q] = cur[q];
        // level s, cell q sits at (level s-1 position of win[0]) + 4 + q: taps win[q .. q+8]
#pragma unroll
        for (int q = 0; q < 8; q++) {
            double a = w.w[0] * win[q];
#pragma unroll
            for (int k = 1; k < 9; k++) a = fma(w.w[k], win[q + k], a);
            cur[q] = a;
        }
        p -= 4;
        if (FIX && s < TB) fix_halo(cur, p, s, g);
    }
}

template <int TB>
__global__ void __launch_bounds__(32 * kWarpsPerCta, 3)
k_stencil1d_tb(const __grid_constant__ Geom1DTB g, const __grid_constant__ Weights1D w) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long task = (long long)blockIdx.x * kWarpsPerCta + warp;
    if (task >= g.ntasks) return;

    unsigned char *wsm = smem_raw + warp * kTbWarpSmem;
    double *ring = reinterpret_cast<double *>(wsm);                            // 3 stages x 512 doubles
    double *outbuf = reinterpret_cast<double *>(wsm + kTbRing);                // 2 staging rows x 256 doubles
    double *mailbox = reinterpret_cast<double *>(wsm + kTbRing + kTbOut);      // [level][row parity][8]
    uint64_t *bars = reinterpret_cast<uint64_t *>(wsm + kTbRing + kTbOut + kTbMail);

    const long long s0 = g.lo + task * (long long)g.rows_per_task * kTbRow;  // first output cell (interior coords)
    const long long len = min((long long)g.rows_per_task * kTbRow, g.hi - s0);
    const int J = (int)((len + kTbRow - 1) / kTbRow);  // output rows; iteration i handles row j = i - 1
    const int niter = J + 1;
    const int nst = (niter + kTbStageRows - 1) / kTbStageRows;
    const long long task_start = s0 + 4 * TB - kTbRow;  // interior coordinate of level-0 row -1, lane 0
    const long long padded_len = g.n + 8;

    auto issue = [&](int k, int slot) {
        const long long start = task_start + 4 + (long long)k * kTbStage;  // padded index of the stage
        const long long skip = start < 0 ? -start : 0;                     // cells left of the array: never needed
        const long long cnt = min((long long)kTbStage, padded_len - start) - skip;
        if (cnt <= 0) {
            mbar_arrive(&bars[slot]);
            return;
        }
        const long long even = cnt & ~1LL;
        mbar_arrive_expect_tx(&bars[slot], (uint32_t)(even * 8));
        if (even > 0)
            tma_load_1d(ring + slot * kTbStage + skip, g.in + start + skip, (uint32_t)(even * 8), &bars[slot]);
        if (cnt != even) ring[slot * kTbStage + skip + even] = g.in[start + skip + even];
    };

    if (lane == 0) {
#pragma unroll
        for (int k = 0; k < kStages; k++) mbar_init(&bars[k], 1);
        fence_barrier_init();
#pragma unroll
        for (int k = 0; k < kStages; k++)
            if (k < nst) issue(k, k);
    }
    __syncwarp();

    const long long end = s0 + len;
    for (int i = 0; i < niter; i++) {
        const int st = i / kTbStageRows, rr = i % kTbStageRows, slot = st % kStages;
        if (rr == 0) mbar_wait(&bars[slot], (st / kStages) & 1);
        const double *rowp = ring + slot * kTbStage + rr * kTbRow + 8 * lane;
        const long long p0 = task_start + (long long)i * kTbRow;  // interior coordinate of lane 0's first level-0 cell
        // rows whose skewed levels can touch a virtual halo zone take the (rare) patched path; warp-uniform
        const bool edge = (g.virt_left && p0 - 4 * TB < 0) || (g.virt_right && p0 + kTbRow > g.n);
        double cur[8];
        if (edge)
            sweep_row<TB, true>(rowp, mailbox, i, lane, p0 + 8 * lane, g, w, cur);
        else
            sweep_row<TB, false>(rowp, mailbox, i, lane, p0 + 8 * lane, g, w, cur);
        __syncwarp();  // mailbox hand-over between consecutive rows; every lane has consumed the stage row

        // level TB, row j = i - 1: cells s0 + 256 j + 8 lane .. +7
        if (i >= 1) {
            const long long prow = p0 - 4 * TB;  // first cell of the output row
            if (prow + kTbRow <= end) {
                // whole row: stage it (conflict-free STS) and let the TMA write the 2 KB line segment
                double *stage = outbuf + (i & 1) * kTbRow;
                if (lane == 0) tma_store_wait_read<1>();  // the store issued from this staging row two rows ago has drained
                __syncwarp();
                store_rotated(stage + 8 * lane, lane, cur);
                fence_proxy_async();  // generic-proxy writes -> visible to the async proxy
                __syncwarp();
                if (lane == 0) {
                    tma_store_1d(g.out + 4 + prow, stage, kTbRow * 8);
                    tma_store_commit();
                }
            } else {
                const long long p = prow + 8 * lane;
                double *o = g.out + 4 + p;
#pragma unroll
                for (int q = 0; q < 8; q++)
                    if (p + q < end) o[q] = cur[q];
            }
        }
        if (rr == kTbStageRows - 1 || i == niter - 1) {
            if (lane == 0 && st + kStages < nst) issue(st + kStages, slot);
        }
    }
    if (lane == 0) tma_store_wait_read<0>();  // shared memory must outlive the last TMA stores
}

template <int TB>
cudaError_t launch_tb(const Geom1DTB &g, const Weights1D &w, cudaStream_t s) {
    const long long ctas = (g.ntasks + kWarpsPerCta - 1) / kWarpsPerCta;
    k_stencil1d_tb<TB><<<(unsigned)ctas, 32 * kWarpsPerCta, kSmem1Tb, s>>>(g, w);
    return cudaGetLastError();
}

template <int TB>
cudaError_t opt_in_tb() {
    return cudaFuncSetAttribute(k_stencil1d_tb<TB>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem1Tb);
}

}  // namespace

cudaError_t kernels_init_1d_tb() {
    cudaError_t e;
    if ((e = opt_in_tb<1>()) != cudaSuccess) return e;
    if ((e = opt_in_tb<2>()) != cudaSuccess) return e;
    if ((e = opt_in_tb<3>()) != cudaSuccess) return e;
    return opt_in_tb<4>();
}

cudaError_t launch_1d_tb(int tb, const Geom1DTB &g, const Weights1D &w, cudaStream_t s) {
    if (g.ntasks <= 0) return cudaSuccess;
    switch (tb) {
        case 1: return launch_tb<1>(g, w, s);
        case 2: return launch_tb<2>(g, w, s);
        case 3: return launch_tb<3>(g, w, s);
        case 4: return launch_tb<4>(g, w, s);
        default: return cudaErrorInvalidValue;
    }
}

}  // namespace lora
