// stencil3d_r2.cu -- radius-2 3-D shapes (box3d2r: 5 x 5 x 5 = 125 taps, star3d2r: 13 taps).
//
// The reference stops at radius 1 in 3-D (src/3d/3d_utils.h:39-42 is its whole shape list; SURVEY.md section 8(f)-4 asks
// for the radius-2 members of the paper family).  Layout: (h + 4) x (m + 4) x (n + 8) doubles, i.e. the reference's 3-D
// layout (src/3d/main.cu:21-23) with the plane halo widened to the radius; same launch semantics as every other kernel
// here: reads the padded source, writes the INTERIOR of the destination (S2).
//
// 2.5-D streaming in the "push" formulation of stencil2d_push.cuh, along the plane axis: a thread owns one (row,
// column) and walks the planes of its chunk.  Every plane p is read ONCE (its 5 x 5 in-plane neighbourhood, or the 9
// cells of the in-plane cross) and pushed into the five outputs it contributes to -- accumulators of planes p-2 .. p+2
// held in registers as a shift register; after plane p the accumulator of plane p-2 is complete and is stored.  Three
// forms, chosen by the host from the structure of the table (decompose.cpp: decompose_3d_r2), in the spirit of the
// low-rank adaptation of the 2-D forms:
//   STAR13   13 FMA per cell: in-plane cross into the middle accumulator, 4 centre taps into the others
//   HSEP5    30 FMA: w[dh][dr][dc] = a[dh] * Q[dr][dc] (rank 1 along the plane axis, ANY in-plane 5 x 5 table Q):
//            one in-plane sum T = Q . plane, then a[dh] * T into the five accumulators
//   DIRECT125  125 FMA: five in-plane sums, one per accumulator
// In-plane neighbours come through L1 (25 loads per cell and plane, 5 x 5 threads share them); no shared memory, no
// barriers, so rows / columns outside the grid simply retire.  A chunk of planes costs 4 extra plane reads of warm-up.
#include "common.cuh"
#include "kernels.h"
#include "../../include/lorastencil.h"

namespace lora {

namespace {

constexpr int kR2Cols = 128;  // threads along the columns (one cell each)
constexpr int kR2Rows = 2;    // rows per CTA

template <int FORM>
__global__ void __launch_bounds__(kR2Cols * kR2Rows)
k_stencil3d_r2(const __grid_constant__ Geom3DR2 g, const __grid_constant__ WeightsR2 w) {
    const int c = blockIdx.x * kR2Cols + threadIdx.x;
    const int r = blockIdx.y * kR2Rows + threadIdx.y;
    if (c >= g.n || r >= g.m) return;
    const long long q_lo = g.lo + (long long)blockIdx.z * g.planes_per_chunk;     // first output plane of this chunk
    const long long q_hi = min(q_lo + (long long)g.planes_per_chunk, g.hi);       // one past the last
    // cell (plane p, r, c) of the interior sits at padded (p + 2, r + 2, c + 4)
    const long long cell = (long long)(r + 2) * g.row_pitch + 4 + c;
    const double *in = g.in + cell;
    double *out = g.out + cell;

    double acc[5] = {0.0, 0.0, 0.0, 0.0, 0.0};  // acc[k]: output plane p - 2 + k while plane p is being pushed
    for (long long p = q_lo - 2; p <= q_hi + 1; p++) {
        const double *pl = in + (p + 2) * g.plane_pitch;
        if (FORM == LORA_FORM_STAR13) {
            const double ctr = pl[0];
            double t = w.w[62] * ctr;  // (0, 0, 0)
#pragma unroll
            for (int d = 1; d <= 2; d++) {
                t = fma(w.w[62 - d], pl[-d], t);                      // (0, 0, -d)
                t = fma(w.w[62 + d], pl[d], t);                       // (0, 0, +d)
                t = fma(w.w[62 - 5 * d], pl[-d * g.row_pitch], t);    // (0, -d, 0)
                t = fma(w.w[62 + 5 * d], pl[d * g.row_pitch], t);     // (0, +d, 0)
            }
            acc[2] += t;
            // plane p is plane q + dh of output q = p - dh: tap (dh, 0, 0) = w[62 + 25 dh], accumulator k = 2 - dh
            acc[0] = fma(w.w[62 + 50], ctr, acc[0]);
            acc[1] = fma(w.w[62 + 25], ctr, acc[1]);
            acc[3] = fma(w.w[62 - 25], ctr, acc[3]);
            acc[4] = fma(w.w[62 - 50], ctr, acc[4]);
        } else {
            double v[25];
#pragma unroll
            for (int dr = -2; dr <= 2; dr++)
#pragma unroll
                for (int dc = -2; dc <= 2; dc++) v[(dr + 2) * 5 + dc + 2] = pl[dr * g.row_pitch + dc];
            if (FORM == LORA_FORM_HSEP5) {
                double t = w.q[0] * v[0];
#pragma unroll
                for (int i = 1; i < 25; i++) t = fma(w.q[i], v[i], t);
#pragma unroll
                for (int k = 0; k < 5; k++) acc[k] = fma(w.a[4 - k], t, acc[k]);  // dh = 2 - k, a[dh + 2]
            } else {
#pragma unroll
                for (int k = 0; k < 5; k++) {
                    const double *wk = w.w + (4 - k) * 25;  // the in-plane table of dh = 2 - k
                    double t = acc[k];
#pragma unroll
                    for (int i = 0; i < 25; i++) t = fma(wk[i], v[i], t);
                    acc[k] = t;
                }
            }
        }
        if (p - 2 >= q_lo) out[p * g.plane_pitch] = acc[0];  // output plane p - 2 at padded plane p
        acc[0] = acc[1];
        acc[1] = acc[2];
        acc[2] = acc[3];
        acc[3] = acc[4];
        acc[4] = 0.0;
    }
}

}  // namespace

cudaError_t launch_3d_r2(int form, const Geom3DR2 &g, const WeightsR2 &w, cudaStream_t s) {
    if (g.hi <= g.lo) return cudaSuccess;
    const long long chunks = (g.hi - g.lo + g.planes_per_chunk - 1) / g.planes_per_chunk;
    const long long by = (g.m + kR2Rows - 1) / kR2Rows;
    if (chunks > 65535 || by > 65535) return cudaErrorInvalidConfiguration;
    const dim3 grid((unsigned)((g.n + kR2Cols - 1) / kR2Cols), (unsigned)by, (unsigned)chunks);
    const dim3 block(kR2Cols, kR2Rows);
    switch (form) {
        case LORA_FORM_STAR13: k_stencil3d_r2<LORA_FORM_STAR13><<<grid, block, 0, s>>>(g, w); break;
        case LORA_FORM_HSEP5: k_stencil3d_r2<LORA_FORM_HSEP5><<<grid, block, 0, s>>>(g, w); break;
        case LORA_FORM_DIRECT125: k_stencil3d_r2<LORA_FORM_DIRECT125><<<grid, block, 0, s>>>(g, w); break;
        default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

int r2_rows_per_cta() { return kR2Rows; }
int r2_cols_per_cta() { return kR2Cols; }

}  // namespace lora
