// stencil2d_tb.cu -- 2-D stencils with TEMPORAL BLOCKING: TB (3, or 2) launches of the reference's 2-D kernels
// (src/2d/gpu.cu:31-273) fused into one sweep; the intermediate grids never leave the register file.
// New functionality (the reference launches one kernel per time step).
//
// Same worker model as stencil2d.cu -- a warp owns a strip of 128 columns, lane l four of them, and sweeps a
// chunk of rows top to bottom through its private TMA ring -- extended into a register pipeline of TB levels:
//
//   input row i  --push-->  A_1 (6 row accumulators: a shift register, stencil2d_push.cuh)  --completed row-->  v_1
//   v_1 + 3 columns from either neighbour lane (6 FP64 warp shuffles)  --push-->  A_2  --retire-->  v_2 ...
//   ... v_TB leaves with one 256-bit store.
//
// Every level lags the one before by 3 rows (the radius), so one loop iteration advances all TB levels by one
// row: a plain row loop (the accumulators shift by register renaming inside the push, nothing is unrolled).
// Lanes 0 and 31 have no neighbour on one side, so the valid strip shrinks by one lane (4 columns >= radius 3)
// per level and side: a warp reads 136 columns, computes 128 at every level and writes 128 - 8 (TB - 1)
// (overlapped tiling across strips: 14 % redundant FP64 work at TB = 3, no inter-warp synchronisation at all).
// Vertically a chunk re-computes 3 (TB - s) rows of level s above and below its output rows.
//
// Reference semantics under fusion (S2): the halo ring of the grid launch t reads is the
// caller's halo when t is even and zero when t is odd, and no launch ever writes it.  Intermediate levels are
// therefore PATCHED before they feed the next level: a retired cell outside the interior takes
// (its time is even) ? caller's halo (buffer 0 of the ping-pong) : 0.  With TB = 3 the time parity equals the buffer
// parity at level 0 and the source buffer's own halo ring is already the right one; sweeps of TB = 2 all start at even
// times, so their callers put the caller's ring around BOTH buffers while they run (plan.cu: lora_plan_run).
#include "common.cuh"
#include "kernels.h"
#include "stencil2d_push.cuh"
#include "../../include/lorastencil.h"

namespace lora {

namespace {

constexpr unsigned kFull = 0xffffffffu;

struct SweepTB {
    const CUtensorMap *tmap;
    double *ring;
    uint64_t *bars;
    double *orow;           // output row pointer (this lane's first column), advances by pitch once rows retire
    const double *hsrc;     // caller's halo: padded buffer 0, pointing at (row 0, this lane's first column)
    const double *hal;      // shared memory: caller's halo columns of the task's rows, [nin][4 left + 4 right]
    volatile int *scratch;  // shared memory: one word per warp (the guard store of the early refill)
    long long pitch, mirror;
    int nin, nst, boxcol, row0_padded, lane;
    int rho0;               // interior row of input row 0 of the chunk (= r0 - 3 TB)
    int c0;                 // interior column of this lane's first cell
    int m, n;
    int out_lo, out_hi;     // interior rows this chunk writes: [out_lo, out_hi)
    bool store_lane;        // this lane's columns are valid at level TB
    bool col_edge;          // the strip has cells outside [0, n): intermediate levels need column patching
    bool hal_ok;            // ... and the task is short enough for its halo columns to be staged in s.hal
    bool virt_top, virt_bot;  // rows above 0 / below m are the global halo ring (virtual), not a neighbour slab's rows
    int par0, par_mask;     // time parity before level 0; 1 = the halo alternates with the level (reference), 0 = fixed
    bool vec4;
};

// Virtual halo for a retired row of level `level` (time par0 + level): cells outside the interior are not computed
// values but (time even) ? caller's halo : 0.  Only groups of rows that can meet the ring run this (EDGE phases).
//   * a whole row outside [0, m) (at most 3 per level at the top / bottom of the grid): loaded on the spot;
//   * cells of an inside row whose column is outside [0, n) (first / last strip, every row): the caller's-halo
//     columns of all the task's rows were staged in shared memory when the task started (s.hal: 4 left + 4 right
//     doubles per row), so the patch is a predicated LDS, not an L2 / DRAM round trip per row.
__device__ __forceinline__ void patch_row(double (&v)[4], int i_row, int rho, int level, const SweepTB &s) {
    const bool caller = ((s.par0 + (level & s.par_mask)) & 1) == 0;  // warp-uniform: at even times the ring holds the caller's halo
    const bool row_out = (s.virt_top && rho < 0) || (s.virt_bot && rho >= s.m);  // warp-uniform
    if (row_out) {
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const int c = s.c0 + q;
            double h = 0.0;
            if (caller && rho >= -4 && rho < s.m + 4 && c >= -4 && c < s.n + 4) h = s.hsrc[(long long)rho * s.pitch + q];
            v[q] = h;
        }
    } else if (s.col_edge && s.hal_ok) {
        const double *hr = s.hal + 8 * min(max(i_row, 0), s.nin - 1);
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const int c = s.c0 + q;
            if (c < 0)
                v[q] = (caller && c >= -4) ? hr[c + 4] : 0.0;
            else if (c >= s.n)
                v[q] = (caller && c < s.n + 4) ? hr[4 + c - s.n] : 0.0;
        }
    } else if (s.col_edge) {
        // an INNER strip whose 128-column window crosses column n (the last strip is narrower than 8 columns) runs
        // as a long task, too long for the staging area: its few halo cells come straight from global memory
        const bool row_ok = rho >= -4 && rho < s.m + 4;
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const int c = s.c0 + q;
            if (c < 0 || c >= s.n)
                v[q] = (caller && row_ok && c >= -4 && c < s.n + 4) ? s.hsrc[(long long)rho * s.pitch + q] : 0.0;
        }
    }
}

template <int FORM, int TB, bool EDGE>
__device__ __forceinline__ void row_phase(int i, SweepTB &s, double (&A)[TB][kAcc][4], const Weights2D &w,
                                          const WeightsDirect49 &wd) {
    const int st = i / kRowsPerStage, rr = i % kRowsPerStage, slot = st % kStages;
    if (rr == 0) mbar_wait(&s.bars[slot], (st / kStages) & 1);
    const double2 *rowp = reinterpret_cast<const double2 *>(s.ring + slot * kStageElems + rr * kBoxCols + 4 * s.lane);
    double x[12];
#pragma unroll
    for (int k = 0; k < 6; k++) {
        const double2 v = rowp[k];
        x[2 * k] = v.x;
        x[2 * k + 1] = v.y;
    }
    // Refilling a stage needs care: the TMA unit is NOT ordered behind the warp's LDS queue.  Issuing the refill right
    // after the stage's last loads were *issued* (round 1) let a refill that hits in L2 land before a delayed LDS had
    // read its row -- observed about once per 10^6 refills when several grids share the SMs (multi-GPU slabs on one
    // device): the lanes then saw the row 16 rows further down.  Two safe orders:
    //   * late: refill after the push, behind the same guard store (arithmetic is free to move across the refill's asm,
    //     so "after the push" alone orders nothing);
    //   * early (cross form: +6.6 %, 450 -> 480 GStencil/s; the diamond form loses 11 % to extra spills): refill before
    //     the push, but behind a shared-memory store of the XOR of every value of the row -- that store cannot issue
    //     before the loads have completed, and the TMA is issued after it in program order.
    constexpr bool kEarlyRefill = FORM == LORA_FORM_CROSS;
    const bool refill_row = rr == kRowsPerStage - 1 || i == s.nin - 1;
    int guard = 0;
    if (refill_row) {
#pragma unroll
        for (int k = 0; k < 12; k++) guard ^= __double2hiint(x[k]);
    }
    if (kEarlyRefill && refill_row) {
        __syncwarp();  // every lane has its values of this stage
        if (s.lane == 0 && st + kStages < s.nst) {
            *s.scratch = guard;
            mbar_arrive_expect_tx(&s.bars[slot], kStageElems * 8);
            tma_load_2d(s.ring + slot * kStageElems, s.tmap, s.boxcol, s.row0_padded + (st + kStages) * kRowsPerStage,
                        &s.bars[slot]);
        }
    }
    double v[4];  // the row level 1 completes: rho0 + i - 3
    push_row<FORM>(x, A[0], v, w, wd);
    if (!kEarlyRefill && refill_row) {
        __syncwarp();  // every lane has consumed this stage
        if (s.lane == 0 && st + kStages < s.nst) {
            *s.scratch = guard;  // behind the push in program order is not enough for ptxas: the same guard store
            mbar_arrive_expect_tx(&s.bars[slot], kStageElems * 8);
            tma_load_2d(s.ring + slot * kStageElems, s.tmap, s.boxcol, s.row0_padded + (st + kStages) * kRowsPerStage,
                        &s.bars[slot]);
        }
    }

#pragma unroll
    for (int lv = 1; lv < TB; lv++) {
        // v = the row level lv has just completed: rho0 + i - 3 lv
        if (EDGE) patch_row(v, i - 3 * lv, s.rho0 + i - 3 * lv, lv, s);
        // window of level lv: own 4 columns + 3 from either neighbour lane
        x[4] = v[0];
        x[5] = v[1];
        x[6] = v[2];
        x[7] = v[3];
        x[1] = __shfl_up_sync(kFull, v[1], 1);
        x[2] = __shfl_up_sync(kFull, v[2], 1);
        x[3] = __shfl_up_sync(kFull, v[3], 1);
        x[8] = __shfl_down_sync(kFull, v[0], 1);
        x[9] = __shfl_down_sync(kFull, v[1], 1);
        x[10] = __shfl_down_sync(kFull, v[2], 1);
        x[0] = 0.0;
        x[11] = 0.0;
        push_row<FORM>(x, A[lv], v, w, wd);
    }

    // v = level TB, row rho0 + i - 3 TB
    const int rout = s.rho0 + i - 3 * TB;
    if (rout >= s.out_lo && rout < s.out_hi) {
        if (s.store_lane) {
            const int left = s.n - s.c0;
            if (left >= 4) {
                if (s.vec4) {
                    st_global_v4(s.orow, v[0], v[1], v[2], v[3]);
                } else {
                    st_global_v2(s.orow, v[0], v[1]);
                    st_global_v2(s.orow + 2, v[2], v[3]);
                }
            } else {
#pragma unroll
                for (int q = 0; q < 4; q++)
                    if (q < left) s.orow[q] = v[q];
            }
            if (s.mirror != 0) {  // the same row into the neighbour slab's ghost rows (peer memory over NVLink)
                double *om = s.orow + s.mirror;
                if (left >= 4) {
                    if (s.vec4) {
                        st_global_v4(om, v[0], v[1], v[2], v[3]);
                    } else {
                        st_global_v2(om, v[0], v[1]);
                        st_global_v2(om + 2, v[2], v[3]);
                    }
                } else {
#pragma unroll
                    for (int q = 0; q < 4; q++)
                        if (q < left) om[q] = v[q];
                }
            }
        }
        s.orow += s.pitch;
    }
}

// rows per trip of the row loop: the light cross form wants a few rows in flight for the scheduler to interleave, the
// FP64-heavy forms want the smallest loop body (instruction cache)
#ifndef LORA_ROWS_UNROLL
#define LORA_ROWS_UNROLL 0
#endif
template <int FORM>
constexpr int kRowsUnroll = LORA_ROWS_UNROLL ? LORA_ROWS_UNROLL : 1;

// The row loop.  Rows whose retired rows can meet the halo ring run the phase WITH the patch code, all others the lean
// one (a warp-uniform branch per row): a task at the top or bottom of the grid pays for the patch code during its first /
// last few rows only; tasks of the first / last strip pay for it throughout.
template <int FORM, int TB>
__device__ __forceinline__ void sweep_rows(SweepTB &s, const Weights2D &w, const WeightsDirect49 &wd) {
    double A[TB][kAcc][4];
#pragma unroll
    for (int lv = 0; lv < TB; lv++)
#pragma unroll
        for (int j = 0; j < kAcc; j++)
#pragma unroll
            for (int q = 0; q < 4; q++) A[lv][j][q] = 0.0;

    constexpr int U = kRowsUnroll<FORM>;
    for (int base = 0; base < s.nin; base += U) {
        // rows retired by levels 1 .. TB-1 during these input rows: rho0 + base - 3 (TB - 1) .. rho0 + base + U - 1 - 3
        const bool edge = s.col_edge || (s.virt_top && s.rho0 + base - 3 * (TB - 1) < 0) ||
                          (s.virt_bot && s.rho0 + base + U - 4 >= s.m);
        if (edge) {
#pragma unroll
            for (int k = 0; k < U; k++)
                if (base + k < s.nin) row_phase<FORM, TB, true>(base + k, s, A, w, wd);
        } else {
#pragma unroll
            for (int k = 0; k < U; k++)
                if (base + k < s.nin) row_phase<FORM, TB, false>(base + k, s, A, w, wd);
        }
    }
}

// Sweeps of two launches never need the caller's halo at their one intermediate level under the reference's
// alternating halo (level 1 sits at an odd time: zeros), so they run without the halo staging area: 70 KB instead of
// 115 KB of shared memory per CTA, i.e. THREE CTAs (12 warps) per SM where their register count allows it -- these
// kernels are latency-bound at two warps per scheduler (profiles/r2_ncu_kernels.md).  (A Dirichlet boundary reads its
// few halo cells from global memory instead.)
template <int TB>
struct Smem2DTB {
    static constexpr int scratch = TB == 2 ? kSmem12 : kSmem2TbScratch;  // one guard word per warp
    static constexpr int bytes = scratch + 64;
};

template <int FORM, int TB>
#ifndef LORA_TB2_CTAS
#define LORA_TB2_CTAS 3
#endif
__global__ void __launch_bounds__(32 * kWarpsPerCta, TB == 2 ? LORA_TB2_CTAS : 2)
k_stencil2d_tb(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ Geom2DTB g,
               const __grid_constant__ Weights2D w, const __grid_constant__ WeightsDirect49 wd) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    const int warp = uniform_warp_id(), lane = threadIdx.x & 31;
    const int task = blockIdx.x * kWarpsPerCta + warp;
    if (task >= g.ntasks) return;  // warps never synchronise with each other

    constexpr int kStripOut = kWarpCols - 8 * (TB - 1);  // columns a strip writes
    int strip, r0, R, seg;
    if (!decode_task_2dtb(g, task, strip, r0, R, seg)) return;  // task order and lengths: kernels.h
    const int cs = strip * kStripOut;                    // first interior column the strip writes
    const int cw = cs - 4 * (TB - 1);                    // first interior column the warp computes

    SweepTB s;
    s.tmap = &tmap;
    s.ring = reinterpret_cast<double *>(smem_raw) + warp * (kStages * kStageElems);
    s.bars = reinterpret_cast<uint64_t *>(smem_raw + kWarpsPerCta * kStages * kStageElems * 8) + warp * kStages;
    s.nin = R + 6 * TB;  // input rows r0 - 3 TB .. r0 + R + 3 TB - 1
    s.nst = (s.nin + kRowsPerStage - 1) / kRowsPerStage;
    s.boxcol = cw;                       // padded column of the box origin = interior column cw - 4, + 4
    s.row0_padded = r0 - 3 * TB + 4;     // padded row of input row 0 (negative rows are zero-filled by the TMA)
    s.lane = lane;
    s.rho0 = r0 - 3 * TB;
    s.c0 = cw + 4 * lane;
    s.m = g.m;
    s.n = g.n;
    s.out_lo = r0;
    s.out_hi = r0 + R;
    s.store_lane = lane >= TB - 1 && lane <= 32 - TB && s.c0 < g.n && s.c0 < cs + kStripOut;
    s.col_edge = cw < 0 || cw + kWarpCols > g.n;
    s.hal_ok = TB != 2 && R + 6 * TB <= kHalRows2Tb;
    s.virt_top = g.virt_top != 0;
    s.virt_bot = g.virt_bot != 0;
    s.par0 = g.par0 & 1;
    s.par_mask = g.par_mask & 1;
    s.vec4 = g.vec4 != 0;
    s.pitch = g.pitch;
    s.mirror = g.sg.mirror[seg];
    s.orow = g.out + (long long)(r0 + 4) * g.pitch + 4 + s.c0;
    s.hsrc = g.halo_src + 4 * g.pitch + 4 + s.c0;
    double *hal = reinterpret_cast<double *>(smem_raw + kSmem12) + warp * (kHalRows2Tb * 8);
    s.hal = hal;
    s.scratch = reinterpret_cast<volatile int *>(smem_raw + Smem2DTB<TB>::scratch) + 2 * warp;
    if (s.col_edge && s.hal_ok) {
        // stage the caller's halo columns (4 left of column 0, 4 right of column n-1) of the task's rows
        for (int idx = lane; idx < s.nin; idx += 32) {
            const int rho = s.rho0 + idx;
            const bool row_ok = rho >= -4 && rho < g.m + 4;
            const double *rowp = g.halo_src + (long long)(rho + 4) * g.pitch;
#pragma unroll
            for (int q = 0; q < 4; q++) {
                hal[8 * idx + q] = row_ok ? rowp[q] : 0.0;
                hal[8 * idx + 4 + q] = row_ok ? rowp[g.n + 4 + q] : 0.0;
            }
        }
    }

    if (lane == 0) {
#pragma unroll
        for (int k = 0; k < kStages; k++) mbar_init(&s.bars[k], 1);
        fence_barrier_init();
#pragma unroll
        for (int k = 0; k < kStages; k++)
            if (k < s.nst) {
                mbar_arrive_expect_tx(&s.bars[k], kStageElems * 8);
                tma_load_2d(s.ring + k * kStageElems, &tmap, s.boxcol, s.row0_padded + k * kRowsPerStage, &s.bars[k]);
            }
    }
    __syncwarp();

    sweep_rows<FORM, TB>(s, w, wd);
    const int seg_done = seg_of(g.sg, blockIdx.x * kWarpsPerCta + uniform_warp_id());  // recomputed: not kept live
    if (g.sg.flag[seg_done] != nullptr) {  // a band task: tell the neighbour once every task of the band has stored
        __threadfence_system();
        __syncwarp();
        if ((threadIdx.x & 31) == 0) seg_arrive(g.sg, seg_done);
    }
}

template <int FORM, int TB>
cudaError_t launch_form(const CUtensorMap &tmap, const Geom2DTB &g, const Weights2D &w, const WeightsDirect49 &wd,
                        cudaStream_t st) {
    if (g.ntasks <= 0) return cudaSuccess;
    const int ctas = (g.ntasks + kWarpsPerCta - 1) / kWarpsPerCta;
    k_stencil2d_tb<FORM, TB><<<ctas, 32 * kWarpsPerCta, Smem2DTB<TB>::bytes, st>>>(tmap, g, w, wd);
    return cudaGetLastError();
}

template <int FORM, int TB>
cudaError_t opt_in() {
    return cudaFuncSetAttribute(k_stencil2d_tb<FORM, TB>, cudaFuncAttributeMaxDynamicSharedMemorySize, Smem2DTB<TB>::bytes);
}

}  // namespace

cudaError_t kernels_init_2d_tb() {
    cudaError_t e;
    if ((e = opt_in<LORA_FORM_PYRAMID, 3>()) != cudaSuccess) return e;
    if ((e = opt_in<LORA_FORM_PYRAMID_PRUNED, 3>()) != cudaSuccess) return e;
    if ((e = opt_in<LORA_FORM_CROSS, 3>()) != cudaSuccess) return e;
    if ((e = opt_in<LORA_FORM_DIAMOND, 3>()) != cudaSuccess) return e;
    if ((e = opt_in<LORA_FORM_PYRAMID, 2>()) != cudaSuccess) return e;
    if ((e = opt_in<LORA_FORM_PYRAMID_PRUNED, 2>()) != cudaSuccess) return e;
    if ((e = opt_in<LORA_FORM_DIAMOND, 2>()) != cudaSuccess) return e;
    return cudaSuccess;
}

int strip_out_cols_2d_tb(int tb) { return kWarpCols - 8 * (tb - 1); }

cudaError_t launch_2d_tb(int form, int tb, const CUtensorMap &tmap, const Geom2DTB &g, const Weights2D &w,
                         const WeightsDirect49 &wd, cudaStream_t s) {
    if (tb == 2) {  // two launches per sweep: no spills for the forms that are heavy at three (stencil2d_tb.cu header)
        switch (form) {
            case LORA_FORM_PYRAMID: return launch_form<LORA_FORM_PYRAMID, 2>(tmap, g, w, wd, s);
            case LORA_FORM_PYRAMID_PRUNED: return launch_form<LORA_FORM_PYRAMID_PRUNED, 2>(tmap, g, w, wd, s);
            case LORA_FORM_DIAMOND: return launch_form<LORA_FORM_DIAMOND, 2>(tmap, g, w, wd, s);
            default: return cudaErrorInvalidValue;
        }
    }
    if (tb != 3) return cudaErrorInvalidValue;
    switch (form) {
        case LORA_FORM_PYRAMID: return launch_form<LORA_FORM_PYRAMID, 3>(tmap, g, w, wd, s);
        case LORA_FORM_PYRAMID_PRUNED: return launch_form<LORA_FORM_PYRAMID_PRUNED, 3>(tmap, g, w, wd, s);
        case LORA_FORM_CROSS: return launch_form<LORA_FORM_CROSS, 3>(tmap, g, w, wd, s);
        case LORA_FORM_DIAMOND: return launch_form<LORA_FORM_DIAMOND, 3>(tmap, g, w, wd, s);
        default: return cudaErrorInvalidValue;
    }
}

}  // namespace lora
