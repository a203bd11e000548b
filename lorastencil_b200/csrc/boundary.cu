// boundary.cu -- the halo ring of a padded grid as the PERIODIC image of its interior (LORA_BOUNDARY_PERIODIC).
//
// The reference has no boundary update at all: its kernels write the interior only and the halo a launch sees
// alternates between the caller's values and zeros (S2: src/2d/gpu.cu:396-400, store offsets :106).  SURVEY.md section
// 8(f)-4 asks for the boundary conditions the paper family uses instead; a periodic one is a copy, before every launch,
// of the `halo` interior cells next to the far face of each axis into the halo cells in front of the near face, and
// vice versa.  One launch per axis, innermost axis first: a later axis copies whole padded lines / planes, wrapped
// halo cells of the earlier axes included, which is what puts the right values into the edges and corners of the ring.
// 24 to 64 bytes per boundary line against 16 bytes per interior cell of the sweep that follows -- grid-stride copies
// through L2, no staging.
#include "common.cuh"
#include "kernels.h"

namespace lora {

namespace {

constexpr int kWrapThreads = 256;

// The array seen as [outer][len + 2 halo][inner], row-major.  Work item t = (o, j, i), j in [0, 2 halo): j < halo fills
// the leading halo cell j from interior cell len + j - halo (padded index len + j), the others fill the trailing halo
// cell halo + len + (j - halo) from interior cell j - halo (padded index j).  Shared by the kernel and by the host
// restatement below that the CPU tests check against numpy's wrap padding.
__host__ __device__ inline void wrap_item(long long t, long long len, int halo, long long inner, long long *dst, long long *src) {
    const long long per_outer = 2LL * halo * inner;
    const long long line = (len + 2LL * halo) * inner;
    const long long o = t / per_outer, rem = t % per_outer;
    const long long j = rem / inner, i = rem % inner;
    const long long base = o * line + i;
    *dst = base + (j < halo ? j : len + j) * inner;
    *src = base + (j < halo ? len + j : j) * inner;
}

__global__ void __launch_bounds__(kWrapThreads)
k_wrap_axis(double *buf, long long outer, long long len, int halo, long long inner) {
    const long long total = outer * 2LL * halo * inner;
    for (long long t = blockIdx.x * (long long)kWrapThreads + threadIdx.x; t < total; t += (long long)gridDim.x * kWrapThreads) {
        long long dst, src;
        wrap_item(t, len, halo, inner, &dst, &src);
        buf[dst] = buf[src];
    }
}

}  // namespace

// one axis of the ring: `outer` lines of `len` interior cells (+ `halo` either side), each cell `inner` doubles wide
cudaError_t launch_wrap_axis(double *buf, long long outer, long long len, int halo, long long inner, int sm_count,
                             cudaStream_t s) {
    const long long total = outer * 2LL * halo * inner;
    if (total <= 0) return cudaSuccess;
    long long blocks = (total + kWrapThreads - 1) / kWrapThreads;
    const long long cap = (long long)sm_count * 8;  // whole waves of 8 CTAs per SM, grid-stride beyond that
    if (blocks > cap) blocks = cap;
    k_wrap_axis<<<(unsigned)blocks, kWrapThreads, 0, s>>>(buf, outer, len, halo, inner);
    return cudaGetLastError();
}

// the same work items on host memory, one after the other (lora_debug_wrap_ring_host: CPU tests, no GPU needed)
void wrap_axis_host(double *buf, long long outer, long long len, int halo, long long inner) {
    const long long total = outer * 2LL * halo * inner;
    for (long long t = 0; t < total; t++) {
        long long dst, src;
        wrap_item(t, len, halo, inner, &dst, &src);
        buf[dst] = buf[src];
    }
}

}  // namespace lora
