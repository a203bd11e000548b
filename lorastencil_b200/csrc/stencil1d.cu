// stencil1d.cu -- 1-D 9-tap stencil kernel for sm_100a (shapes 1d1r, 1d2r).
//
// Replaces kernel_1d1r / kernel_1d2r of the reference (src/1d/gpu_1r.cu:21-87,
// src/1d/gpu_2r.cu:22-88), which fold 1024 outputs into an 8x128 matrix so the convolution becomes
// X(8x16) . P(16x8) on the FP64 tensor cores (16 MACs per output, 9 useful).  Here the line is cut
// into "rows" of 128 outputs; every warp is an independent worker that sweeps a run of consecutive
// rows.  Its private TMA ring (cp.async.bulk, no tensor map needed in 1-D) stages 8 rows + 8 halo
// doubles per transaction; lane l reads its 12-double window with six 128-bit LDS, evaluates the
// 9 taps for its 4 outputs with FP64 FMAs (weights from the constant bank / uniform registers) and
// writes them with one 256-bit store.  9 MACs per output, no CTA-wide synchronisation.
#include "common.cuh"
#include "kernels.h"

namespace lora {

namespace {

constexpr int kRowElems = kWarpCols;                       // 128 outputs per row
constexpr int kStageRows1 = kRowsPerStage;                 // rows per bulk copy
constexpr int kStageLoad1 = kStageRows1 * kRowElems + 8;   // doubles fetched per stage (520)

__global__ void __launch_bounds__(32 * kWarpsPerCta, 4)
k_stencil1d(const __grid_constant__ Geom1D g, const __grid_constant__ Weights1D w) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    const int warp = uniform_warp_id(), lane = threadIdx.x & 31;
    const long long task = (long long)blockIdx.x * kWarpsPerCta + warp;
    if (task >= g.ntasks) return;

    double *ring = reinterpret_cast<double *>(smem_raw) + warp * (kStages * kStageElems);
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem_raw + kWarpsPerCta * kStages * kStageElems * 8) + warp * kStages;

    const long long s0 = g.lo + task * (long long)g.rows_per_task * kRowElems;  // first interior index
    const long long len = min((long long)g.rows_per_task * kRowElems, g.hi - s0);
    const int nrows = (int)((len + kRowElems - 1) / kRowElems);
    const int nst = (nrows + kStageRows1 - 1) / kStageRows1;
    const long long padded_len = g.n + 8;

    // stage k covers padded indices [s0 + k*512, s0 + k*512 + 520); clip to the array, keep 16-byte multiples
    auto issue = [&](int k, int slot) {
        const long long start = s0 + (long long)k * kStageRows1 * kRowElems;
        long long cnt = min((long long)kStageLoad1, padded_len - start);
        const long long even = cnt & ~1LL;
        mbar_arrive_expect_tx(&bars[slot], (uint32_t)(even * 8));
        tma_load_1d(ring + slot * kStageElems, g.in + start, (uint32_t)(even * 8), &bars[slot]);
        if (cnt != even) ring[slot * kStageElems + even] = g.in[start + even];  // odd tail element (n odd)
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

    for (int r = 0; r < nrows; r++) {
        const int st = r / kStageRows1, rr = r % kStageRows1, slot = st % kStages;
        if (rr == 0) mbar_wait(&bars[slot], (st / kStages) & 1);
        const double2 *rowp =
            reinterpret_cast<const double2 *>(ring + slot * kStageElems + rr * kRowElems + 4 * lane);
        double x[12];
#pragma unroll
        for (int k = 0; k < 6; k++) {
            const double2 v = rowp[k];
            x[2 * k] = v.x;
            x[2 * k + 1] = v.y;
        }
        // x[j] = padded[s + j], output q (interior s + q, padded s + q + 4) = sum_k w[k] * x[q + k]
        double y[4];
#pragma unroll
        for (int q = 0; q < 4; q++) {
            double a = w.w[0] * x[q];
#pragma unroll
            for (int k = 1; k < 9; k++) a = fma(w.w[k], x[q + k], a);
            y[q] = a;
        }
        const long long s = s0 + (long long)r * kRowElems + 4 * lane;  // interior index of y[0]
        double *o = g.out + 4 + s;
        if (s + 3 < g.hi) {
            if (g.vec4) {
                st_global_v4(o, y[0], y[1], y[2], y[3]);
            } else {
                st_global_v2(o, y[0], y[1]);
                st_global_v2(o + 2, y[2], y[3]);
            }
        } else {
#pragma unroll
            for (int q = 0; q < 4; q++)
                if (s + q < g.hi) o[q] = y[q];
        }
        if (g.mirror != 0) {  // the same cells into the neighbour slab's halo (peer memory)
#pragma unroll
            for (int q = 0; q < 4; q++)
                if (s + q < g.hi) o[g.mirror + q] = y[q];
        }
        if (rr == kStageRows1 - 1 || r == nrows - 1) {
            __syncwarp();
            if (lane == 0 && st + kStages < nst) issue(st + kStages, slot);
            __syncwarp();  // lane 0's plain store of an odd tail element is ordered before the other lanes' reads
        }
    }
}

}  // namespace

cudaError_t kernels_init_1d() {
    return cudaFuncSetAttribute(k_stencil1d, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem12);
}

cudaError_t launch_1d(const Geom1D &g, const Weights1D &w, cudaStream_t s) {
    if (g.ntasks <= 0) return cudaSuccess;
    const long long ctas = (g.ntasks + kWarpsPerCta - 1) / kWarpsPerCta;
    k_stencil1d<<<(unsigned)ctas, 32 * kWarpsPerCta, kSmem12, s>>>(g, w);
    return cudaGetLastError();
}

}  // namespace lora
