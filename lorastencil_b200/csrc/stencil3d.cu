// stencil3d.cu -- 3-D 27-point / 7-point stencil kernels for sm_100a (shapes box3d1r, star3d1r).
//
// Replaces kernel_box3d1r / kernel_star3d1r of the reference (src/3d/gpu_box.cu:21-140,
// src/3d/gpu_star.cu:22-133).  Like the reference this is 2.5-D streaming over the outermost axis,
// but built as a producer/consumer pipeline instead of load -> wait -> 4 barriers per plane:
//
//   * a CTA owns a 32-row x 128-column tile and a chunk of planes.  34 x 132 halo tiles of consecutive
//     planes stream through a 4-deep shared-memory ring (cp.async.bulk.tensor.3d), gated by full/empty
//     mbarriers; lane 0 of warp 0 refills the slot freed one plane earlier, so three planes are always
//     in flight and nothing in the plane loop is a CTA-wide barrier.
//   * consumer warp w owns rows 4w..4w+3 of the tile, lane l owns 4 consecutive columns: a 4x4
//     register micro-tile.  Per plane it reads 6 rows x 8 doubles (128-bit LDS), releases the stage,
//     evaluates the in-plane operator in registers and PUSHES the result into three plane
//     accumulators (output planes p-1, p, p+1).  The oldest accumulator is complete after every plane
//     and leaves with 256-bit stores; the accumulator ring is rotated by a 3x unrolled plane loop.
//
// In-plane operators: SEP3  s = c (*) x along n, t = b (*) s along m, push a[.] * t      (~10.5 FP64 ops/cell)
//                     STAR7 5 in-plane taps + 2 pushes                                  (7 ops/cell)
//                     DIRECT27 all 27 taps                                             (27 ops/cell)
// (reference: 80 / 24 DMMA MACs per cell, results bounced through shared memory.)
#include "common.cuh"
#include "kernels.h"
#include "../../include/lorastencil.h"

namespace lora {

namespace {

// X[rr][j]: rr = 0..5 <-> tile rows 4w-1 .. 4w+4, j = 0..7 <-> interior columns c0-2 .. c0+5
// micro-tile cell (r, q): row rr = r + 1, column j = q + 2
template <int FORM, int PH>
__device__ __forceinline__ void push_plane(const double (&X)[6][8], double (&A)[3][4][4], const Weights3D &w) {
#define ACC(dh) A[((1 - (dh)) + PH) % 3]
    if constexpr (FORM == LORA_FORM_SEP3) {
        double t[4][4];
#pragma unroll
        for (int rr = 0; rr < 6; rr++) {
#pragma unroll
            for (int q = 0; q < 4; q++) {
                double s = w.c[0] * X[rr][q + 1];
                s = fma(w.c[1], X[rr][q + 2], s);
                s = fma(w.c[2], X[rr][q + 3], s);
                // row rr is the (dr = rr - 1 - r) neighbour of micro-tile row r
#pragma unroll
                for (int r = 0; r < 4; r++) {
                    const int dr = rr - 1 - r;
                    if (dr == -1) t[r][q] = w.b[0] * s;
                    else if (dr == 0) t[r][q] = fma(w.b[1], s, t[r][q]);
                    else if (dr == 1) t[r][q] = fma(w.b[2], s, t[r][q]);
                }
            }
        }
#pragma unroll
        for (int r = 0; r < 4; r++)
#pragma unroll
            for (int q = 0; q < 4; q++) {
                ACC(1)[r][q] = fma(w.a[2], t[r][q], ACC(1)[r][q]);
                ACC(0)[r][q] = fma(w.a[1], t[r][q], ACC(0)[r][q]);
                ACC(-1)[r][q] = fma(w.a[0], t[r][q], ACC(-1)[r][q]);
            }
    } else if constexpr (FORM == LORA_FORM_STAR7) {
#pragma unroll
        for (int r = 0; r < 4; r++)
#pragma unroll
            for (int q = 0; q < 4; q++) {
                const double xc = X[r + 1][q + 2];
                double v = w.star[0] * xc;
                v = fma(w.star[1], X[r + 1][q + 1], v);
                v = fma(w.star[2], X[r + 1][q + 3], v);
                v = fma(w.star[3], X[r][q + 2], v);
                v = fma(w.star[4], X[r + 2][q + 2], v);
                ACC(0)[r][q] += v;
                ACC(1)[r][q] = fma(w.star[6], xc, ACC(1)[r][q]);    // this plane is h+1 of output plane p-1
                ACC(-1)[r][q] = fma(w.star[5], xc, ACC(-1)[r][q]);  // and h-1 of output plane p+1
            }
    } else {  // DIRECT27
#pragma unroll
        for (int dh = -1; dh <= 1; dh++)
#pragma unroll
            for (int r = 0; r < 4; r++)
#pragma unroll
                for (int q = 0; q < 4; q++)
#pragma unroll
                    for (int dr = -1; dr <= 1; dr++)
#pragma unroll
                        for (int dc = -1; dc <= 1; dc++)
                            ACC(dh)[r][q] = fma(w.direct[(dh + 1) * 9 + (dr + 1) * 3 + dc + 1],
                                                X[r + 1 + dr][q + 2 + dc], ACC(dh)[r][q]);
    }
#undef ACC
}

struct Sweep3D {
    const unsigned char *ring;
    const CUtensorMap *tmap;
    uint64_t *full, *empty;
    int box_c, box_r, box_h;  // TMA box origin (padded coordinates) of input plane 0 of the chunk
    double *optr;        // output plane pointer for this lane's micro-tile origin (advances by plane_pitch)
    long long row_pitch, plane_pitch, mirror;
    int nin, warp, lane;
    volatile int *guard;       // shared memory: one word per warp (the guard store of the stage release)
    int rows_left, cols_left;  // how many of the 4 micro-tile rows / columns exist
    bool vec4;
    int hout;                  // interior plane index of the next plane to be stored
    int mlo, mhi;              // planes [mlo, mhi) are stored a second time at + mirror (a neighbour slab's ghost planes)
    int early_plane;           // >= 0: after this plane has been stored the CTA reports to the segment's flag (lo band)
    const Segs *sg;
    int seg;
};

template <int FORM, int PH>
__device__ __forceinline__ void plane_phase(int i, Sweep3D &s, double (&A)[3][4][4], const Weights3D &w) {
    const int slot = i % k3Stages;
    mbar_wait(&s.full[slot], (i / k3Stages) & 1);
    const double *tile = reinterpret_cast<const double *>(s.ring + slot * k3StageBytes);
    double X[6][8];
#pragma unroll
    for (int rr = 0; rr < 6; rr++) {
        const double2 *rowp = reinterpret_cast<const double2 *>(tile + (4 * s.warp + rr) * k3BoxCols + 4 * s.lane);
#pragma unroll
        for (int k = 0; k < 4; k++) {
            if (FORM == LORA_FORM_STAR7 && (rr == 0 || rr == 5) && (k == 0 || k == 3)) {
                X[rr][2 * k] = 0.0;  // corner columns of the halo rows are not part of a star
                X[rr][2 * k + 1] = 0.0;
                continue;
            }
            const double2 v = rowp[k];
            X[rr][2 * k] = v.x;
            X[rr][2 * k + 1] = v.y;
        }
    }
    // Release the stage -- but only once its values have ARRIVED in the registers, not merely once the loads have been
    // issued: the mbarrier arrive does not queue behind the warp's shared-memory loads, and with eight warps' loads
    // backed up in the load / store unit the refill (which needs nothing but the eight arrivals and an L2 hit) was seen
    // landing on rows a delayed LDS had yet to read -- a handful of wrong rows per 10^5 planes, run-to-run differences
    // in profiles/debug/tb3_stress.py.  A shared-memory store of a word computed from every loaded value cannot issue
    // before the loads have completed, and the arrive follows it in program order.
    {
        int gw = 0;
#pragma unroll
        for (int rr = 0; rr < 6; rr++)
#pragma unroll
            for (int k = 0; k < 4; k++) gw ^= __double2hiint(X[rr][2 * k]);
        *s.guard = gw;
    }
    __syncwarp();
    if (s.lane == 0) mbar_arrive(&s.empty[slot]);  // this warp no longer needs the stage
    // producer duty: refill the slot every warp released one plane ago with plane i - 1 + k3Stages.  All of warp 0 waits
    // (a warp-uniform branch): a spin loop under `lane == 0` makes ptxas treat the plane loop as divergent, and the
    // weights then sit in vector registers instead of uniform ones
    const int nx = i - 1 + k3Stages;
    if (s.warp == 0 && i >= 1 && nx < s.nin) {
        const int ps = (i - 1) % k3Stages;
        mbar_wait(&s.empty[ps], ((i - 1) / k3Stages) & 1);
        if (s.lane == 0) {
            mbar_arrive_expect_tx(&s.full[ps], k3BoxRows * k3BoxCols * 8);
            tma_load_3d(const_cast<unsigned char *>(s.ring) + ps * k3StageBytes, s.tmap, s.box_c, s.box_r,
                        s.box_h + nx, &s.full[ps]);
        }
    }

    push_plane<FORM, PH>(X, A, w);

    double(&done)[4][4] = A[PH % 3];  // logical accumulator 0: output plane i - 2 of the chunk
    if (i >= 2) {
#pragma unroll
        for (int r = 0; r < 4; r++) {
            if (r < s.rows_left) {
                double *o = s.optr + r * s.row_pitch;
                if (s.cols_left >= 4) {
                    if (s.vec4) {
                        st_global_v4(o, done[r][0], done[r][1], done[r][2], done[r][3]);
                    } else {
                        st_global_v2(o, done[r][0], done[r][1]);
                        st_global_v2(o + 2, done[r][2], done[r][3]);
                    }
                } else {
#pragma unroll
                    for (int q = 0; q < 4; q++)
                        if (q < s.cols_left) o[q] = done[r][q];
                }
                if (s.mirror != 0 && s.hout >= s.mlo && s.hout < s.mhi) {  // the same row into the neighbour slab's ghost plane (NVLink)
                    double *om = o + s.mirror;
                    if (s.cols_left >= 4) {
                        if (s.vec4) {
                            st_global_v4(om, done[r][0], done[r][1], done[r][2], done[r][3]);
                        } else {
                            st_global_v2(om, done[r][0], done[r][1]);
                            st_global_v2(om + 2, done[r][2], done[r][3]);
                        }
                    } else {
#pragma unroll
                        for (int q = 0; q < 4; q++)
                            if (q < s.cols_left) om[q] = done[r][q];
                    }
                }
            }
        }
        s.optr += s.plane_pitch;
        if (s.hout == s.early_plane) {  // CTA-uniform: the lo band is complete -- tell the neighbour now, not at the end
            __threadfence_system();
            __syncthreads();
            if (threadIdx.x == 0) seg_arrive(*s.sg, s.seg);
        }
        s.hout++;
    }
#pragma unroll
    for (int r = 0; r < 4; r++)
#pragma unroll
        for (int q = 0; q < 4; q++) done[r][q] = 0.0;
}

template <int FORM>
__global__ void __launch_bounds__(k3Threads, 1)
k_stencil3d(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ Geom3D g,
            const __grid_constant__ Weights3D w) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    uint64_t *full = reinterpret_cast<uint64_t *>(smem_raw + k3Stages * k3StageBytes);
    uint64_t *empty = full + k3Stages;
    const int warp = uniform_warp_id(), lane = threadIdx.x & 31;

    const int tile_n = blockIdx.x % g.tiles_n, tile_m = blockIdx.x / g.tiles_n;
    const int seg = seg_of(g.sg, blockIdx.y);  // band segments come first in blockIdx.y, i.e. in dispatch order
    const int h0 = (int)(g.sg.lo[seg] + (blockIdx.y - g.sg.first[seg]) * g.sg.chunk[seg]);  // first interior plane of the chunk
    const int H = (int)min(g.sg.chunk[seg], g.sg.hi[seg] - h0);
    const int nin = H + 2;                                    // input planes h0-1 .. h0+H == padded h0 ..
    const int r_tile = tile_m * k3TileRows, c_tile = tile_n * k3TileCols;

    if (threadIdx.x == 0) {
#pragma unroll
        for (int k = 0; k < k3Stages; k++) {
            mbar_init(&full[k], 1);
            mbar_init(&empty[k], k3Warps);
        }
        fence_barrier_init();
        // box origin: padded col = c_tile + 4 - 2, padded row = r_tile + 2 - 1, padded plane = h0 + i
#pragma unroll
        for (int k = 0; k < k3Stages; k++)
            if (k < nin) {
                mbar_arrive_expect_tx(&full[k], k3BoxRows * k3BoxCols * 8);
                tma_load_3d(smem_raw + k * k3StageBytes, &tmap, c_tile + 2, r_tile + 1, h0 + k, &full[k]);
            }
    }
    __syncthreads();  // the only CTA-wide barrier: barrier objects are initialised

    Sweep3D s;
    s.ring = smem_raw;
    s.tmap = &tmap;
    s.box_c = c_tile + 2;
    s.box_r = r_tile + 1;
    s.box_h = h0;
    s.full = full;
    s.empty = empty;
    s.guard = reinterpret_cast<volatile int *>(empty + k3Stages) + warp;
    s.nin = nin;
    s.warp = warp;
    s.lane = lane;
    const int r0 = r_tile + 4 * warp, c0 = c_tile + 4 * lane;
    s.rows_left = g.m - r0;
    s.cols_left = g.n - c0;
    s.vec4 = g.vec4 != 0;
    s.row_pitch = g.row_pitch;
    s.plane_pitch = g.plane_pitch;
    s.mirror = g.sg.mirror[seg];
    s.hout = h0;
    s.mlo = (int)g.sg.mlo[seg];
    s.mhi = (int)g.sg.mhi[seg];
    const bool early = g.sg.flag[seg] != nullptr && g.sg.early[seg] != 0;
    s.early_plane = early ? s.mhi - 1 : -1;
    s.sg = &g.sg;
    s.seg = seg;
    s.optr = g.out + (long long)(h0 + 1) * g.plane_pitch + (long long)(r0 + 2) * g.row_pitch + 4 + c0;

    double A[3][4][4];
#pragma unroll
    for (int j = 0; j < 3; j++)
#pragma unroll
        for (int r = 0; r < 4; r++)
#pragma unroll
            for (int q = 0; q < 4; q++) A[j][r][q] = 0.0;

    for (int base = 0; base < nin; base += 3) {
        if (base + 0 < nin) plane_phase<FORM, 0>(base + 0, s, A, w);
        if (base + 1 < nin) plane_phase<FORM, 1>(base + 1, s, A, w);
        if (base + 2 < nin) plane_phase<FORM, 2>(base + 2, s, A, w);
    }
    if (g.sg.flag[seg] != nullptr && !early) {  // a band chunk: tell the neighbour once every CTA of the band has stored
        __threadfence_system();
        __syncthreads();
        if (threadIdx.x == 0) seg_arrive(g.sg, seg);
    }
}

template <int FORM>
cudaError_t launch_form(const CUtensorMap &tmap, const Geom3D &g, const Weights3D &w, cudaStream_t st) {
    const int chunks = (int)g.sg.first[g.sg.nseg];
    if (chunks <= 0) return cudaSuccess;
    dim3 grid(g.tiles_m * g.tiles_n, chunks);
    k_stencil3d<FORM><<<grid, k3Threads, k3Smem, st>>>(tmap, g, w);
    return cudaGetLastError();
}

template <int FORM>
cudaError_t opt_in() {
    return cudaFuncSetAttribute(k_stencil3d<FORM>, cudaFuncAttributeMaxDynamicSharedMemorySize, k3Smem);
}

}  // namespace

cudaError_t kernels_init_3d() {
    cudaError_t e;
    if ((e = opt_in<LORA_FORM_SEP3>()) != cudaSuccess) return e;
    if ((e = opt_in<LORA_FORM_STAR7>()) != cudaSuccess) return e;
    if ((e = opt_in<LORA_FORM_DIRECT27>()) != cudaSuccess) return e;
    return cudaSuccess;
}

cudaError_t launch_3d(int form, const CUtensorMap &tmap, const Geom3D &g, const Weights3D &w, cudaStream_t s) {
    switch (form) {
        case LORA_FORM_SEP3: return launch_form<LORA_FORM_SEP3>(tmap, g, w, s);
        case LORA_FORM_STAR7: return launch_form<LORA_FORM_STAR7>(tmap, g, w, s);
        case LORA_FORM_DIRECT27: return launch_form<LORA_FORM_DIRECT27>(tmap, g, w, s);
        default: return cudaErrorInvalidValue;
    }
}

cudaError_t kernels_init() {
    cudaError_t e;
    if ((e = kernels_init_1d()) != cudaSuccess) return e;
    if ((e = kernels_init_1d_tb()) != cudaSuccess) return e;
    if ((e = kernels_init_2d()) != cudaSuccess) return e;
    if ((e = kernels_init_2d_tb()) != cudaSuccess) return e;
    if ((e = kernels_init_3d_tb()) != cudaSuccess) return e;
    return kernels_init_3d();
}

}  // namespace lora
