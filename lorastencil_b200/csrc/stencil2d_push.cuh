// stencil2d_push.cuh -- the per-row "push" operators shared by the 2-D kernels (stencil2d.cu, stencil2d_tb.cu).
//
// One input row (a lane's 12-double window x[]: 4 own columns + 4 either side) is folded into the per-column register
// accumulators of the seven output rows it contributes to.  The accumulators are a SHIFT REGISTER: A[n] holds the
// output row that sees the NEXT input row at row offset dr = 3 - n, and a push writes
//     out   = (last contribution, dr = +3) + A[0]          -- that output row is complete
//     A[n]  = (contribution at dr = 2 - n) + A[n + 1]      -- n = 0 .. 4
//     A[5]  = (first contribution, dr = -3)                -- an assignment: the row 3 below is born
// An FMA has a destination of its own, so the shift costs nothing: the first FMA of every output row reads the
// neighbouring accumulator and writes this one.  (Round 1 kept the accumulators in place and rotated their ROLES with
// the input row modulo 7, which needs the row loop unrolled 7x: 34-86 KB of code per kernel, no_instruction stalls of
// 0.4-0.5 per issue in the fused kernels; the shift form is a plain loop of 5-12 KB and one accumulator less per
// column and level.)  Forms = the host decomposition's output (decompose.cpp), replacing the reference's
// banded-matrix DMMA chains (src/2d/gpu.cu:68-101, :146-171, :225-264).  Operation order per output cell is unchanged
// (contributions in input-row order, terms in order within a row): same bits as before.
#pragma once
#include "kernels.h"
#include "../../include/lorastencil.h"

namespace lora {

constexpr int kAcc = 6;  // accumulators per column between two pushes (the seventh output row leaves as `out`)

// x[4 + q + dc] is the input at (own column q) + dc.  T[dr + 3] is the accumulator of the output row that sees this input
// row at row offset dr; every form gives each of them at least one contribution per row (none is a pure move), and the
// one at dr = -3 is ASSIGNED by its first term (w * h, not fma(w, h, 0)).  Accumulators are first touched in DESCENDING dr:
// A[n]'s new value is then born after A[n]'s old value has been consumed (by A[n - 1]'s first FMA), so the register
// allocator can keep every accumulator in place -- in ascending order it needs two register moves per double and row.
template <int FORM>
__device__ __forceinline__ void push_row(const double (&x)[12], double (&A)[kAcc][4], double (&out)[4], const Weights2D &w,
                                         const WeightsDirect49 &wd) {
    double T[7][4];
#pragma unroll
    for (int d = 1; d < 7; d++)
#pragma unroll
        for (int q = 0; q < 4; q++) T[d][q] = A[6 - d][q];
#define ACC(dr) T[(dr) + 3]
    if constexpr (FORM == LORA_FORM_PYRAMID_PRUNED) {
        // PYRAMID without the taps the host found to be structurally zero: middle term at offsets -2, 0, +2 only,
        // no centre remainder (decompose.cpp: prune_pyramid)
#pragma unroll
        for (int t = 0; t < 3; t++) {
            const int rad = 3 - t;
#pragma unroll
            for (int q = 0; q < 4; q++) {
                double h = w.horiz[t][3 - rad] * x[4 + q - rad];
#pragma unroll
                for (int dc = -rad + 1; dc <= rad; dc++)
                    if (!(t == 1 && (dc == -1 || dc == 1))) h = fma(w.horiz[t][3 + dc], x[4 + q + dc], h);
#pragma unroll
                for (int dr = rad; dr >= -rad; dr--) {
                    if (t == 1 && (dr == -1 || dr == 1)) continue;
                    if (dr == -3)
                        ACC(dr)[q] = w.vert[t][3 + dr] * h;  // birth of the output row 3 below (t == 0 only)
                    else
                        ACC(dr)[q] = fma(w.vert[t][3 + dr], h, ACC(dr)[q]);
                }
            }
        }
    } else if constexpr (FORM == LORA_FORM_RANK2 || FORM == LORA_FORM_RANK3) {
        // sum of 2 / 3 rank-1 terms of full support (host: lowrank_lu): 14 FP64 operations per term and cell
        constexpr int NT = FORM == LORA_FORM_RANK2 ? 2 : 3;
#pragma unroll
        for (int t = 0; t < NT; t++) {
#pragma unroll
            for (int q = 0; q < 4; q++) {
                double h = w.horiz[t][0] * x[4 + q - 3];
#pragma unroll
                for (int dc = -2; dc <= 3; dc++) h = fma(w.horiz[t][3 + dc], x[4 + q + dc], h);
#pragma unroll
                for (int dr = 3; dr >= -3; dr--) {
                    if (dr == -3 && t == 0)
                        ACC(dr)[q] = w.vert[t][3 + dr] * h;  // birth of the output row 3 below
                    else
                        ACC(dr)[q] = fma(w.vert[t][3 + dr], h, ACC(dr)[q]);
                }
            }
        }
    } else if constexpr (FORM == LORA_FORM_PYRAMID) {
#pragma unroll
        for (int t = 0; t < 3; t++) {
            const int rad = 3 - t;
#pragma unroll
            for (int q = 0; q < 4; q++) {
                double h = w.horiz[t][3 - rad] * x[4 + q - rad];
#pragma unroll
                for (int dc = -rad + 1; dc <= rad; dc++) h = fma(w.horiz[t][3 + dc], x[4 + q + dc], h);
#pragma unroll
                for (int dr = rad; dr >= -rad; dr--) {
                    if (dr == -3)
                        ACC(dr)[q] = w.vert[t][3 + dr] * h;  // birth of the output row 3 below (t == 0 only)
                    else
                        ACC(dr)[q] = fma(w.vert[t][3 + dr], h, ACC(dr)[q]);
                }
            }
        }
#pragma unroll
        for (int q = 0; q < 4; q++) ACC(0)[q] = fma(w.centre, x[4 + q], ACC(0)[q]);
    } else if constexpr (FORM == LORA_FORM_CROSS) {
#pragma unroll
        for (int q = 0; q < 4; q++) {
#pragma unroll
            for (int dr = 3; dr >= -3; dr--) {
                if (dr == -3)
                    ACC(dr)[q] = w.vert[0][3 + dr] * x[4 + q];  // birth of the output row 3 below
                else
                    ACC(dr)[q] = fma(w.vert[0][3 + dr], x[4 + q], ACC(dr)[q]);
            }
            double h = w.horiz[1][0] * x[4 + q - 3];
#pragma unroll
            for (int dc = -2; dc <= 3; dc++)
                if (dc != 0) h = fma(w.horiz[1][3 + dc], x[4 + q + dc], h);
            ACC(0)[q] += h;
        }
    } else if constexpr (FORM == LORA_FORM_DIAMOND) {
#pragma unroll
        for (int q = 0; q < 4; q++) {
            double h = w.horiz[0][1] * x[4 + q - 2];
#pragma unroll
            for (int dc = -1; dc <= 2; dc++) h = fma(w.horiz[0][3 + dc], x[4 + q + dc], h);
            ACC(3)[q] = fma(w.residual[3], x[4 + q], ACC(3)[q]);
#pragma unroll
            for (int dr = 2; dr >= -2; dr--) ACC(dr)[q] = fma(w.vert[0][3 + dr], h, ACC(dr)[q]);
            ACC(2)[q] = fma(w.residual[6], x[4 + q - 2], ACC(2)[q]);
            ACC(2)[q] = fma(w.residual[7], x[4 + q + 2], ACC(2)[q]);
            ACC(0)[q] = fma(w.residual[0], x[4 + q - 3], ACC(0)[q]);
            ACC(0)[q] = fma(w.residual[1], x[4 + q + 3], ACC(0)[q]);
            ACC(-2)[q] = fma(w.residual[4], x[4 + q - 2], ACC(-2)[q]);
            ACC(-2)[q] = fma(w.residual[5], x[4 + q + 2], ACC(-2)[q]);
            ACC(-3)[q] = w.residual[2] * x[4 + q];  // birth of the output row 3 below
        }
    } else {  // DIRECT49
#pragma unroll
        for (int dr = 3; dr >= -3; dr--)
#pragma unroll
            for (int q = 0; q < 4; q++)
#pragma unroll
                for (int dc = -3; dc <= 3; dc++) {
                    if (dr == -3 && dc == -3)
                        ACC(dr)[q] = wd.w[(dr + 3) * 7 + dc + 3] * x[4 + q + dc];  // birth of the output row 3 below
                    else
                        ACC(dr)[q] = fma(wd.w[(dr + 3) * 7 + dc + 3], x[4 + q + dc], ACC(dr)[q]);
                }
    }
#undef ACC
#pragma unroll
    for (int q = 0; q < 4; q++) out[q] = T[6][q];
#pragma unroll
    for (int d = 0; d < 6; d++)
#pragma unroll
        for (int q = 0; q < 4; q++) A[5 - d][q] = T[d][q];
}

}  // namespace lora
