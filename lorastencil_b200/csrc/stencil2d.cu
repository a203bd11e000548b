// stencil2d.cu -- 2-D low-rank stencil kernels for sm_100a.
//
// Replaces kernel2d_box2d3r / kernel2d_star2d3r / kernel2d_star2d1r of the reference
// (src/2d/gpu.cu:31-273).  Same mathematics -- the 7x7 table applied as a sum of rank-1 terms
// (vertical profile (x) horizontal profile), a cross, or one rank-1 term + residual taps -- but a
// different machine mapping:
//
//   * every warp is an independent worker: it owns a strip of 128 columns and sweeps a chunk of rows
//     top to bottom.  Lane l owns 4 consecutive columns.
//   * input rows reach shared memory through the warp's private TMA ring (cp.async.bulk.tensor.2d,
//     boxes of 136 columns x 8 rows, 2 stages, one mbarrier per stage); no CTA-wide barrier exists.
//   * per input row a lane reads its 12-double window with six 128-bit LDS, forms the horizontal
//     profile sums h_t (FP64 FMA, weights as uniform-register operands) and PUSHES u_t[dr] * h_t
//     into the per-column register accumulators of the seven output rows it touches -- a shift
//     register of six (stencil2d_push.cuh): the first FMA of every output row reads the neighbouring
//     accumulator and writes this one, the completed row leaves with one 256-bit store.  A plain
//     row loop, no register moves.
//
// MACs per cell: pyramid 31, cross 13, diamond 18, direct 49 (reference: 108 / 32 / 36+8 DMMA MACs).
// FP64 DMMA (mma.sync m8n8k4, the only FP64 tensor shape on sm_100a) is not used here: the operands
// are banded Toeplitz matrices (<= 7/16 dense), so the tensor pipe would spend >2x the FP64 work of
// the FMA form; see DESIGN.md and profiles/ for the measured pipe rates.
#include "common.cuh"
#include "kernels.h"
#include "stencil2d_push.cuh"
#include "../../include/lorastencil.h"

namespace lora {

namespace {

struct Sweep2D {
    const CUtensorMap *tmap;
    double *ring;
    uint64_t *bars;
    double *orow;  // next output row, this lane's first column
    long long pitch, mirror;
    int nin, nst, boxcol, row0_padded, lane, ncols_left;  // ncols_left = n - c0 (how many of the 4 columns exist)
    bool vec4;
};

// one input row: wait for its stage if it opens one, read the window, push, retire the oldest
// accumulator, refill the ring if the row closes a stage
template <int FORM>
__device__ __forceinline__ void row_phase(int i, Sweep2D &s, double (&A)[kAcc][4], const Weights2D &w,
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
    double done[4];  // output row i - 6 of the chunk
    push_row<FORM>(x, A, done, w, wd);
    if (i >= 6) {
        if (s.ncols_left >= 4) {
            if (s.vec4) {
                st_global_v4(s.orow, done[0], done[1], done[2], done[3]);
            } else {
                st_global_v2(s.orow, done[0], done[1]);
                st_global_v2(s.orow + 2, done[2], done[3]);
            }
        } else {
#pragma unroll
            for (int q = 0; q < 4; q++)
                if (q < s.ncols_left) s.orow[q] = done[q];
        }
        if (s.mirror != 0) {  // the same row into the neighbour slab's ghost rows (peer memory over NVLink)
            double *om = s.orow + s.mirror;
            if (s.ncols_left >= 4) {
                if (s.vec4) {
                    st_global_v4(om, done[0], done[1], done[2], done[3]);
                } else {
                    st_global_v2(om, done[0], done[1]);
                    st_global_v2(om + 2, done[2], done[3]);
                }
            } else {
#pragma unroll
                for (int q = 0; q < 4; q++)
                    if (q < s.ncols_left) om[q] = done[q];
            }
        }
        s.orow += s.pitch;
    }

    if (rr == kRowsPerStage - 1 || i == s.nin - 1) {
        __syncwarp();  // every lane has consumed this stage
        if (s.lane == 0 && st + kStages < s.nst) {
            mbar_arrive_expect_tx(&s.bars[slot], kStageElems * 8);
            tma_load_2d(s.ring + slot * kStageElems, s.tmap, s.boxcol, s.row0_padded + (st + kStages) * kRowsPerStage,
                        &s.bars[slot]);
        }
    }
}

template <int FORM>
__global__ void __launch_bounds__(32 * kWarpsPerCta,
                                  (FORM == LORA_FORM_PYRAMID || FORM == LORA_FORM_PYRAMID_PRUNED || FORM == LORA_FORM_DIRECT49 ||
                                   FORM == LORA_FORM_RANK2 || FORM == LORA_FORM_RANK3) ? 3 : 4)
k_stencil2d(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ Geom2D g,
            const __grid_constant__ Weights2D w, const __grid_constant__ WeightsDirect49 wd) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    const int warp = uniform_warp_id(), lane = threadIdx.x & 31;
    const int task = blockIdx.x * kWarpsPerCta + warp;
    if (task >= g.ntasks) return;  // warps never synchronise with each other

    const int seg = seg_of(g.sg, task);  // band segments come first, i.e. are dispatched first
    const int tt = task - (int)g.sg.first[seg];
    const int strip = tt % g.nstrips, chunk = tt / g.nstrips;
    const int r0 = (int)g.sg.lo[seg] + chunk * (int)g.sg.chunk[seg];  // first interior row of this chunk
    const int R = min((int)g.sg.chunk[seg], (int)g.sg.hi[seg] - r0);
    const int c0 = strip * kWarpCols + 4 * lane;         // first interior column of this lane

    Sweep2D s;
    s.tmap = &tmap;
    s.ring = reinterpret_cast<double *>(smem_raw) + warp * (kStages * kStageElems);
    s.bars = reinterpret_cast<uint64_t *>(smem_raw + kWarpsPerCta * kStages * kStageElems * 8) + warp * kStages;
    s.nin = R + 6;  // input rows r0-3 .. r0+R+2  ==  padded rows r0+1 .. r0+R+6
    s.nst = (s.nin + kRowsPerStage - 1) / kRowsPerStage;
    s.boxcol = strip * kWarpCols;  // padded column of the box origin = interior column - 4
    s.row0_padded = r0 + 1;
    s.lane = lane;
    s.ncols_left = g.n - c0;
    s.vec4 = g.vec4 != 0;
    s.pitch = g.pitch;
    s.mirror = g.sg.mirror[seg];
    s.orow = g.out + (long long)(r0 + 4) * g.pitch + 4 + c0;

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

    double A[kAcc][4];
#pragma unroll
    for (int j = 0; j < kAcc; j++)
#pragma unroll
        for (int q = 0; q < 4; q++) A[j][q] = 0.0;

    for (int i = 0; i < s.nin; i++) row_phase<FORM>(i, s, A, w, wd);
    const int seg_done = seg_of(g.sg, task);  // recomputed: not kept live across the sweep
    if (g.sg.flag[seg_done] != nullptr) {  // a band task: tell the neighbour once every task of the band has stored
        __threadfence_system();
        __syncwarp();
        if (lane == 0) seg_arrive(g.sg, seg_done);
    }
}

template <int FORM>
cudaError_t launch_form(const CUtensorMap &tmap, const Geom2D &g, const Weights2D &w, const WeightsDirect49 &wd,
                        cudaStream_t st) {
    if (g.ntasks <= 0) return cudaSuccess;
    const int ctas = (g.ntasks + kWarpsPerCta - 1) / kWarpsPerCta;
    k_stencil2d<FORM><<<ctas, 32 * kWarpsPerCta, kSmem12, st>>>(tmap, g, w, wd);
    return cudaGetLastError();
}

template <int FORM>
cudaError_t opt_in() {
    return cudaFuncSetAttribute(k_stencil2d<FORM>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem12);
}

}  // namespace

cudaError_t kernels_init_2d() {
    cudaError_t e;
    if ((e = opt_in<LORA_FORM_PYRAMID>()) != cudaSuccess) return e;
    if ((e = opt_in<LORA_FORM_PYRAMID_PRUNED>()) != cudaSuccess) return e;
    if ((e = opt_in<LORA_FORM_CROSS>()) != cudaSuccess) return e;
    if ((e = opt_in<LORA_FORM_DIAMOND>()) != cudaSuccess) return e;
    if ((e = opt_in<LORA_FORM_DIRECT49>()) != cudaSuccess) return e;
    if ((e = opt_in<LORA_FORM_RANK2>()) != cudaSuccess) return e;
    if ((e = opt_in<LORA_FORM_RANK3>()) != cudaSuccess) return e;
    return cudaSuccess;
}

cudaError_t launch_2d(int form, const CUtensorMap &tmap, const Geom2D &g, const Weights2D &w,
                      const WeightsDirect49 &wd, cudaStream_t s) {
    switch (form) {
        case LORA_FORM_PYRAMID: return launch_form<LORA_FORM_PYRAMID>(tmap, g, w, wd, s);
        case LORA_FORM_PYRAMID_PRUNED: return launch_form<LORA_FORM_PYRAMID_PRUNED>(tmap, g, w, wd, s);
        case LORA_FORM_CROSS: return launch_form<LORA_FORM_CROSS>(tmap, g, w, wd, s);
        case LORA_FORM_DIAMOND: return launch_form<LORA_FORM_DIAMOND>(tmap, g, w, wd, s);
        case LORA_FORM_DIRECT49: return launch_form<LORA_FORM_DIRECT49>(tmap, g, w, wd, s);
        case LORA_FORM_RANK2: return launch_form<LORA_FORM_RANK2>(tmap, g, w, wd, s);
        case LORA_FORM_RANK3: return launch_form<LORA_FORM_RANK3>(tmap, g, w, wd, s);
        default: return cudaErrorInvalidValue;
    }
}

}  // namespace lora
