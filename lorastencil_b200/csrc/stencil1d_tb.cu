// stencil1d_tb.cu -- 1-D 9-tap stencil with TEMPORAL BLOCKING: TB launches of the reference's
// kernel_1d1r / kernel_1d2r (src/1d/gpu_1r.cu:21-87) fused into one sweep, intermediate time levels
// living in registers only.  New functionality: the reference has no temporal blocking (it only counts
// its composite 9-tap kernels as 3 / 2 steps, src/1d/gpu_1r.cu:132).
//
// The kernel is built to be bound by the FP64 pipe (9 DFMA per cell per level is the floor for general
// weights), i.e. everything else has to cost (much) less than one issue slot per DFMA:
//
//   * one warp = one independent worker with a private TMA ring; the line is cut, in PADDED coordinates X,
//     into rows of 512 cells [512 r, 512 r + 512), lane l owns 16 consecutive cells (128 bytes) of a row;
//   * level s (the line after s of the TB launches) is held SKEWED by -4 s cells: the 24-cell window a
//     lane needs from level s-1 is lane l-1's last 8 cells followed by its own 16, so a level costs 8 FP64
//     warp shuffles (16 SHFL.32) against 144 DFMA per lane, never a shared-memory round trip, and a row
//     never needs the row after it.  The shuffle is circular: lane 0 receives lane 31's cells, which are
//     what it needs one row LATER -- it parks them in a private 64-byte mailbox per level and row parity (one
//     thread writes and reads it, so there is no cross-lane synchronisation);
//   * level 0 arrives by cp.async.bulk.tensor.2d through a tensor map that views the line as rows of 16
//     doubles with the 128-byte hardware swizzle: a lane's 128 contiguous bytes are then read with eight
//     conflict-free 128-bit LDS (16-byte chunk k sits at chunk k ^ (lane & 7)) -- no padding, no register
//     un-shuffling;
//   * level TB leaves through a swizzled 4 KB staging row and a tensor-map TMA store.  The skew moves the
//     output row to X - 4 TB, so the store map's base is shifted by (-4 TB mod 16) cells; whatever the maps
//     cannot reach (the first / last few cells of the array, partial rows of a sub-range) goes through plain
//     predicated global accesses on the rare edge path.
//
// Reference semantics (S2) under fusion: launch i of the reference sees the caller's
// halo when i is even and zeros when i is odd.  Inside a fused sweep the halo cells of every level are
// therefore VIRTUAL: (time of the level even) ? caller's halo (read from `halo_src`, the padded buffer 0
// whose halo is never written) : 0.  Only rows touching an end of the global line take that path.
#include "common.cuh"
#include "kernels.h"

namespace lora {

namespace {

constexpr int kCpl = kTbCellsPerLane;  // 16
constexpr int kRow = kTbRowCells;      // 512
constexpr unsigned kFull = 0xffffffffu;
#ifndef LORA_TB_UNROLL
#define LORA_TB_UNROLL 1
#endif
constexpr int kTbUnroll = LORA_TB_UNROLL;  // levels per iteration of the level loop (tuning experiments only)

// virtual halo of one level: padded cells 0..3 and n+4..n+7 take (level time even ? caller's halo : 0)
__device__ __forceinline__ void fix_halo(double (&v)[kCpl], long long X, int level, const Geom1DTB &g) {
    const bool use_h = ((g.par0 + (level & g.par_mask)) & 1) == 0;
#pragma unroll
    for (int q = 0; q < kCpl; q++) {
        const long long x = X + q;
        const bool in_left = g.virt_left && x >= 0 && x < 4;
        const bool in_right = g.virt_right && x >= g.n + 4 && x < g.n + 8;
        if (in_left || in_right) v[q] = use_h ? g.halo_src[x] : 0.0;
    }
}

// One row through all TB levels.  X = padded coordinate of this lane's first level-0 cell; on return cur[]
// holds level TB of the cells X - 4 TB .. X - 4 TB + 15.
// NT = taps actually computed: 9, or 7 when the host found w[0] == w[8] == 0 (the reference's 1d1r table
// {0,1,2,3,4,3,2,1,0}): structural zeros cost nothing -- 7 instead of 9 FP64 operations per cell per level.
template <bool FIX, int NT>
__device__ __forceinline__ void sweep_row(const unsigned char *stage, double *mailbox, int par, int lane, long long X,
                                          const Geom1DTB &g, const Weights1D &w, double (&cur)[kCpl]) {
#pragma unroll
    for (int j = 0; j < kTbLaneRows; j++) {  // the lane's rows of 16 doubles: tensor-map row kTbLaneRows * lane + j
        const int row = kTbLaneRows * lane + j;
        const unsigned char *rowp = stage + row * 128;
        const int sw = row & 7;
#pragma unroll
        for (int k = 0; k < 8; k++) {
            const double2 v = *reinterpret_cast<const double2 *>(rowp + ((k ^ sw) << 4));
            cur[16 * j + 2 * k] = v.x;
            cur[16 * j + 2 * k + 1] = v.y;
        }
    }
    if (FIX) {
        // cells the load map does not cover (the tail of the array; everything when there is no map)
#pragma unroll
        for (int q = 0; q < kCpl; q++) {
            const long long x = X + q;
            if (x >= g.xcov || x < 0) cur[q] = (x >= 0 && x < g.n + 8) ? g.in[x] : 0.0;
        }
        fix_halo(cur, X, 0, g);
    }
    const int src_lane = (lane + 31) & 31;
    const int TB = g.tb;
    // the level loop is deliberately NOT unrolled: one level is ~200 instructions, a fully unrolled row would
    // not fit the instruction cache (measured: 1349 -> 1431 GStencil/s at TB = 8 when rolled)
#pragma unroll(kTbUnroll)
    for (int s = 1; s <= TB; s++) {
        double win[kCpl + 8];
#pragma unroll
        for (int q = 0; q < 8; q++) win[q] = __shfl_sync(kFull, cur[kCpl - 8 + q], src_lane);
        if (lane == 0) {
            // lane 31's cells belong to the NEXT row's lane 0: park them, take what was parked one row ago
            // (two mailboxes per level alternate by row parity, so the load can land in the window registers)
            double *mb_wr = mailbox + ((s - 1) * 2 + par) * 8;
            const double *mb_rd = mailbox + ((s - 1) * 2 + (par ^ 1)) * 8;
#pragma unroll
            for (int q = 0; q < 8; q += 2) *reinterpret_cast<double2 *>(mb_wr + q) = make_double2(win[q], win[q + 1]);
#pragma unroll
            for (int q = 0; q < 8; q += 2) {
                const double2 t = *reinterpret_cast<const double2 *>(mb_rd + q);
                win[q] = t.x;
                win[q + 1] = t.y;
            }
        }
#pragma unroll
        for (int q = 0; q < kCpl; q++) win[8 + q] = cur[q];
        // level s, cell q sits at (level s-1 position of win[0]) + 4 + q: taps win[q .. q+8].  In DESCENDING order cell q
        // reads the old cells q - 8 .. q only, so every new value can take the register of the old one -- measured: the
        // 7-tap variant gains 7 % that way (1998 -> 2143 GStencil/s), the 9-tap one loses 1.5 % (1740 -> 1713) and keeps
        // the ascending order
#ifndef LORA_TB_DESC
#define LORA_TB_DESC (NT == 7)
#endif
#pragma unroll
        for (int qq = 0; qq < kCpl; qq++) {
            const int q = (LORA_TB_DESC) ? kCpl - 1 - qq : qq;
            constexpr int k0 = (9 - NT) / 2;
            double a = w.w[k0] * win[q + k0];
#pragma unroll
            for (int k = k0 + 1; k < 9 - k0; k++) a = fma(w.w[k], win[q + k], a);
            cur[q] = a;
        }
        X -= 4;
        if (FIX && s < TB) fix_halo(cur, X, s, g);
    }
}

template <int NT>
__global__ void __launch_bounds__(32 * kWarpsPerCta, kTbCtasPerSm)
k_stencil1d_tb(const __grid_constant__ CUtensorMap imap, const __grid_constant__ CUtensorMap omap,
               const __grid_constant__ Geom1DTB g, const __grid_constant__ Weights1D w) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    const int warp = uniform_warp_id(), lane = threadIdx.x & 31;
    const long long task = (long long)blockIdx.x * kWarpsPerCta + warp;
    if (task >= g.ntasks) return;  // warps never synchronise with each other
    const int TB = g.tb;

    unsigned char *wsm = smem_raw + warp * kTbWarpSmem;
    unsigned char *ring = wsm;                                                 // kTbStages x 4 KB, swizzled rows
    unsigned char *outbuf = wsm + kTbRing;                                     // 2 staging rows x 4 KB, swizzled
    double *mailbox = reinterpret_cast<double *>(wsm + kTbRing + kTbOut);      // [level][row parity][8]
    uint64_t *bars = reinterpret_cast<uint64_t *>(wsm + kTbRing + kTbOut + kTbMail);

    // segment of this task (band segments first), the cells it may write, its output rows
    const int seg = seg_of(g.sg, task);
    const long long xlo = g.sg.lo[seg], xhi = g.sg.hi[seg], mirror = g.sg.mirror[seg];
    const long long rho0 = (xlo + 4 * TB) / kTbRowCells;
    const long long nrows = (xhi - 1 + 4 * TB) / kTbRowCells - rho0 + 1;
    const long long rfirst = rho0 + (task - g.sg.first[seg]) * g.sg.chunk[seg];  // first output row of the task
    const int J = (int)min(g.sg.chunk[seg], rho0 + nrows - rfirst);
    const int niter = J + 1;  // iteration 0 is the warm-up row rfirst - 1: it only fills the mailboxes

    auto issue = [&](int k, int slot) {
        if (g.use_tma) {
            mbar_arrive_expect_tx(&bars[slot], kRow * 8);
            tma_load_2d(ring + slot * (kRow * 8), &imap, 0, (int)((kRow / 16) * (rfirst - 1 + k)), &bars[slot]);
        } else {
            mbar_arrive(&bars[slot]);
        }
    };

    if (lane == 0) {
#pragma unroll
        for (int k = 0; k < kTbStages; k++) mbar_init(&bars[k], 1);
        fence_barrier_init();
#pragma unroll
        for (int k = 0; k < kTbStages; k++)
            if (k < niter) issue(k, k);
    }
    __syncwarp();

    for (int i = 0; i < niter; i++) {
        const int slot = i % kTbStages;
        mbar_wait(&bars[slot], (i / kTbStages) & 1);
        const long long X0 = (rfirst - 1 + i) * (long long)kRow;  // padded coordinate of lane 0's first level-0 cell
        // rows with cells the load map cannot reach, or whose skewed levels touch a virtual halo zone, take the
        // (rare) patched path; warp-uniform
        const bool edge = (X0 + kRow >= g.xcov) || (g.virt_left && X0 - 4 * (TB - 1) < 4) ||
                          (g.virt_right && X0 + kRow > g.n + 4);
        double cur[kCpl];
        if (edge)
            sweep_row<true, NT>(ring + slot * (kRow * 8), mailbox, i & 1, lane, X0 + kCpl * lane, g, w, cur);
        else
            sweep_row<false, NT>(ring + slot * (kRow * 8), mailbox, i & 1, lane, X0 + kCpl * lane, g, w, cur);
        __syncwarp();  // every lane has consumed the stage
        if (lane == 0 && i + kTbStages < niter) issue(i + kTbStages, slot);

        if (i >= 1) {
            const long long Xr = X0 - 4 * TB;  // first cell of the output row (level TB)
            const long long rc = (Xr - g.out_off) >> 4;  // row of 16 in the store map (exact when Xr >= out_off)
            if (g.use_tma && mirror == 0 && Xr >= xlo && Xr + kRow <= xhi && Xr >= g.out_off && rc + kRow / 16 <= g.out_rows) {
                // whole row: stage it (conflict-free STS) and let the TMA write the 4 KB line segment
                unsigned char *stage = outbuf + (i % kTbOutBufs) * (kRow * 8);
                if (lane == 0) tma_store_wait_read<kTbOutBufs - 1>();  // the last store issued from this staging row has drained
                __syncwarp();
#pragma unroll
                for (int j = 0; j < kTbLaneRows; j++) {
                    const int row = kTbLaneRows * lane + j;
                    unsigned char *rowp = stage + row * 128;
                    const int sw = row & 7;
#pragma unroll
                    for (int k = 0; k < 8; k++)
                        *reinterpret_cast<double2 *>(rowp + ((k ^ sw) << 4)) =
                            make_double2(cur[16 * j + 2 * k], cur[16 * j + 2 * k + 1]);
                }
                fence_proxy_async();  // generic-proxy writes -> visible to the async proxy
                __syncwarp();
                if (lane == 0) {
                    tma_store_2d(&omap, stage, 0, (int)rc);
                    tma_store_commit();
                }
            } else {
                const long long x = Xr + kCpl * lane;
#pragma unroll
                for (int q = 0; q < kCpl; q++)
                    if (x + q >= xlo && x + q < xhi) {
                        g.out[x + q] = cur[q];
                        if (mirror != 0) g.out[x + q + mirror] = cur[q];  // neighbour slab's ghost zone (peer memory)
                    }
            }
        }
    }
    if (lane == 0) tma_store_wait_all();  // shared memory must outlive the last TMA stores
    const int seg_done = seg_of(g.sg, task);  // recomputed: not kept live across the sweep
    if (g.sg.flag[seg_done] != nullptr) {  // a band task: tell the neighbour once every task of the band has stored
        __threadfence_system();
        __syncwarp();
        if (lane == 0) seg_arrive(g.sg, seg_done);
    }
}

}  // namespace

cudaError_t kernels_init_1d_tb() {
    cudaError_t e = cudaFuncSetAttribute(k_stencil1d_tb<9>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem1Tb);
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(k_stencil1d_tb<7>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem1Tb);
}

cudaError_t launch_1d_tb(const CUtensorMap &imap, const CUtensorMap &omap, const Geom1DTB &g, const Weights1D &w,
                         cudaStream_t s) {
    if (g.ntasks <= 0) return cudaSuccess;
    if (g.tb < 1 || g.tb > kMaxTb1) return cudaErrorInvalidValue;
    const long long ctas = (g.ntasks + kWarpsPerCta - 1) / kWarpsPerCta;
    if (w.w[0] == 0.0 && w.w[8] == 0.0)  // outer taps are structural zeros: the 7-tap variant
        k_stencil1d_tb<7><<<(unsigned)ctas, 32 * kWarpsPerCta, kSmem1Tb, s>>>(imap, omap, g, w);
    else
        k_stencil1d_tb<9><<<(unsigned)ctas, 32 * kWarpsPerCta, kSmem1Tb, s>>>(imap, omap, g, w);
    return cudaGetLastError();
}

}  // namespace lora
