// stencil2d_push.cuh -- the per-row "push" operators shared by the 2-D kernels (stencil2d.cu, stencil2d_tb.cu).
//
// One input row (a lane's 12-double window x[]: 4 own columns + 4 either side) is folded into the seven
// per-column register accumulators of the output rows it contributes to; which accumulator belongs to which
// output row rotates with the phase PH (= input row index mod 7), so the caller unrolls its row loop 7x and
// no register is ever moved.  Forms = the host decomposition's output (decompose.cpp), replacing the reference's
// banded-matrix DMMA chains (src/2d/gpu.cu:68-101, :146-171, :225-264).
#pragma once
#include "kernels.h"
#include "../../include/lorastencil.h"

namespace lora {

constexpr int NACC = 7;

// x[4 + q + dc] is the input at (own column q) + dc; accumulator (3 - dr + PH) % 7 belongs to the
// output row that sees this input row at row offset dr.  An output row's FIRST contribution is the one at
// dr = -3, and every form has exactly one such term per column: it ASSIGNS the accumulator (w * h instead of
// fma(w, h, 0)), so a retired accumulator needs neither zeroing nor a copy -- its registers are simply reborn.
template <int FORM, int PH>
__device__ __forceinline__ void push_row(const double (&x)[12], double (&A)[NACC][4], const Weights2D &w,
                                         const WeightsDirect49 &wd) {
#define ACC(dr) A[((3 - (dr)) + PH) % NACC]
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
                for (int dr = -rad; dr <= rad; dr++) {
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
                for (int dr = -3; dr <= 3; dr++) {
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
                for (int dr = -rad; dr <= rad; dr++) {
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
            for (int dr = -3; dr <= 3; dr++) {
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
#pragma unroll
            for (int dr = -2; dr <= 2; dr++) ACC(dr)[q] = fma(w.vert[0][3 + dr], h, ACC(dr)[q]);
            ACC(0)[q] = fma(w.residual[0], x[4 + q - 3], ACC(0)[q]);
            ACC(0)[q] = fma(w.residual[1], x[4 + q + 3], ACC(0)[q]);
            ACC(-3)[q] = w.residual[2] * x[4 + q];  // birth of the output row 3 below
            ACC(3)[q] = fma(w.residual[3], x[4 + q], ACC(3)[q]);
            ACC(-2)[q] = fma(w.residual[4], x[4 + q - 2], ACC(-2)[q]);
            ACC(-2)[q] = fma(w.residual[5], x[4 + q + 2], ACC(-2)[q]);
            ACC(2)[q] = fma(w.residual[6], x[4 + q - 2], ACC(2)[q]);
            ACC(2)[q] = fma(w.residual[7], x[4 + q + 2], ACC(2)[q]);
        }
    } else {  // DIRECT49
#pragma unroll
        for (int dr = -3; dr <= 3; dr++)
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
}

}  // namespace lora
