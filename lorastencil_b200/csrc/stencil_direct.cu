// stencil_direct.cu -- direct-tap 2-D / 3-D kernels WITHOUT TMA, for grids whose padded row length is odd.
//
// A tensor map needs a row pitch that is a multiple of 16 bytes, i.e. an even number of FP64 columns; every BASELINE
// size satisfies that and runs the TMA kernels (stencil2d*.cu, stencil3d.cu).  The reference itself needs much more
// (n % 64 == 0, src/2d/gpu.cu:402-404, src/3d/gpu_box.cu:201-202, and silently reads out of bounds otherwise); this
// kernel removes the last size constraint: any m, n, h >= 1.  It applies the EFFECTIVE weights of the plan (the 49 /
// 27 direct taps equal to whatever form the host decomposition chose), one thread per cell, plain coalesced loads
// through L1 -- a correctness path for odd sizes, not a roofline kernel.  Same launch semantics as the others:
// interior only, segments / mirror stores / band flags of a multi-GPU slab (kernels.h: Segs).
#include "common.cuh"
#include "kernels.h"

namespace lora {

namespace {

constexpr int kDirectThreads = 256;

// task = one row (2-D) or one (plane, row) pair (3-D) x one block of 256 columns
template <int DIM>
__global__ void __launch_bounds__(kDirectThreads)
k_stencil_direct(const __grid_constant__ GeomDirect g, const __grid_constant__ WeightsDirect49 w) {
    const long long blk = blockIdx.x;
    const int cb = (int)(blk % g.col_blocks);
    const long long line = blk / g.col_blocks;          // row (2-D) or plane-row (3-D) index within the launch
    const long long outer_t = DIM == 2 ? line : line / g.m;  // index along the swept axis, in segment order
    const int seg = seg_of(g.sg, outer_t);
    const long long outer = g.sg.lo[seg] + (outer_t - g.sg.first[seg]);  // interior row (2-D) / plane (3-D)
    const int c = cb * kDirectThreads + threadIdx.x;
    if (c < g.n) {
        double acc = 0.0;
        long long o;
        if (DIM == 2) {
            const double *p = g.in + (outer + 4) * g.pitch + 4 + c;
#pragma unroll
            for (int dr = -3; dr <= 3; dr++)
#pragma unroll
                for (int dc = -3; dc <= 3; dc++) acc = fma(w.w[(dr + 3) * 7 + dc + 3], p[dr * g.pitch + dc], acc);
            o = (outer + 4) * g.pitch + 4 + c;
        } else {
            const long long r = line % g.m;
            const double *p = g.in + (outer + 1) * g.plane_pitch + (r + 2) * g.pitch + 4 + c;
#pragma unroll
            for (int dh = -1; dh <= 1; dh++)
#pragma unroll
                for (int dr = -1; dr <= 1; dr++)
#pragma unroll
                    for (int dc = -1; dc <= 1; dc++)
                        acc = fma(w.w[(dh + 1) * 9 + (dr + 1) * 3 + dc + 1], p[dh * g.plane_pitch + dr * g.pitch + dc], acc);
            o = (outer + 1) * g.plane_pitch + (r + 2) * g.pitch + 4 + c;
        }
        g.out[o] = acc;
        if (g.sg.mirror[seg] != 0) g.out[o + g.sg.mirror[seg]] = acc;  // neighbour slab's ghost zone (peer memory)
    }
    if (g.sg.flag[seg] != nullptr) {
        __threadfence_system();
        __syncthreads();
        if (threadIdx.x == 0) seg_arrive(g.sg, seg);
    }
}

}  // namespace

cudaError_t launch_direct(int dim, const GeomDirect &g, const WeightsDirect49 &w, cudaStream_t s) {
    const long long lines = g.sg.first[g.sg.nseg] * (dim == 3 ? g.m : 1);
    const long long blocks = lines * g.col_blocks;
    if (blocks <= 0) return cudaSuccess;
    if (blocks > 0x7fffffffLL) return cudaErrorInvalidConfiguration;
    if (dim == 2)
        k_stencil_direct<2><<<(unsigned)blocks, kDirectThreads, 0, s>>>(g, w);
    else
        k_stencil_direct<3><<<(unsigned)blocks, kDirectThreads, 0, s>>>(g, w);
    return cudaGetLastError();
}

}  // namespace lora
