// stencil3d_tb.cu -- 3-D 7-point / 27-point stencils with TEMPORAL BLOCKING: two launches of the reference's 3-D
// kernels (src/3d/gpu_star.cu:101-133, src/3d/gpu_box.cu:105-140) fused into one sweep; the intermediate grid never
// reaches HBM.  New functionality (the reference launches one kernel per time step).
//
// Same 2.5-D streaming as stencil3d.cu -- a CTA owns a tile and a chunk of planes, halo tiles of consecutive planes
// stream through a 4-deep TMA ring gated by full / empty mbarriers -- with a second level behind the first:
//
//   level-0 plane q --(in-plane taps + the two plane taps, carried in two register accumulators)--> level-1 plane q-1
//   level-1 plane j --(same operator)--> level-2 plane j-1 --> global memory
//
// What keeps it inside the register file and off the shared-memory pipe:
//   * a lane owns RM = 3 rows x 4 columns (not 4 x 4): two levels x two carried accumulators x 12 cells = 48 doubles
//     of state; a 7-point (or separable) operator needs only TWO accumulators per level between planes -- the plane
//     that completes is handed on at once;
//   * level 1 reaches level 2 without a tile round trip: a lane's own 3 x 4 values stay in registers, the columns
//     left / right come from the neighbour lanes by warp shuffle, and only the first / last row of every warp goes
//     through shared memory (2 x 128 doubles per warp and plane, double-buffered, ONE __syncthreads per plane);
//   * overlapped tiling: every level is computed on the full 24 x 128 thread tile, the valid region shrinks by one
//     cell per level and side; a CTA writes 22 rows x 120 columns (lanes 0 and 31 and the outer rows only feed their
//     neighbours), i.e. it reads a 26 x 132 box per 22 x 120 outputs: 18.4 B of DRAM traffic per cell per TWO launches
//     against 2 x 16.8 unfused.
//
// Reference semantics (S2) under fusion: a fused sweep starts at an even time -- level 0 sees the caller's halo, which
// is physically in the source buffer's ring (the host copies the ring of buffer 0 into buffer 1 before the first sweep
// that reads buffer 1 and clears it again afterwards: lora_plan_run) -- and level 1 sits at an odd time, whose halo is
// ZERO: level-1 cells outside the interior are forced to zero before they feed level 2.  Same operations in the same
// order as two unfused launches: results are bit-identical.
#include "common.cuh"
#include "kernels.h"
#include "../../include/lorastencil.h"

namespace lora {

namespace {

constexpr unsigned kFull = 0xffffffffu;
constexpr int RM = kT3Rm;

// one lane's level state: the accumulator of the plane that is one plane short of complete, and the one just born
struct LevelState {
    double full[RM][4];
    double next[RM][4];
};

// 7-point star, in-plane part: centre, n-1, n+1, m-1, m+1 -- the operation order of stencil3d.cu's STAR7 push
__device__ __forceinline__ double star_inplane(const Weights3D &w, double c, double l, double r, double u, double d) {
    double v = w.star[0] * c;
    v = fma(w.star[1], l, v);
    v = fma(w.star[2], r, v);
    v = fma(w.star[3], u, v);
    v = fma(w.star[4], d, v);
    return v;
}

// separable 27-point operator a (x) b (x) c, one level, with the window delivered row by row (only one window row is
// live at a time): s = c * row along n, t = b * s along m, then the two carried plane accumulators -- the operation
// order of stencil3d.cu's SEP3 push.  rowfn(rr, row): window row rr (region row RM*warp - 1 + rr), row[j] = column
// 4*lane - 2 + j; only row[1..6] are read.
template <class RowFn>
__device__ __forceinline__ void sep3_level(const Weights3D &w, RowFn rowfn, LevelState &L, double (&out)[RM][4]) {
    double t[RM][4];
#pragma unroll
    for (int rr = 0; rr < RM + 2; rr++) {
        double row[8];
        rowfn(rr, row);
#pragma unroll
        for (int q = 0; q < 4; q++) {
            double s = w.c[0] * row[q + 1];
            s = fma(w.c[1], row[q + 2], s);
            s = fma(w.c[2], row[q + 3], s);
#pragma unroll
            for (int r = 0; r < RM; r++) {
                const int dr = rr - 1 - r;  // row rr is the dr neighbour of micro-tile row r
                if (dr == -1) t[r][q] = w.b[0] * s;
                else if (dr == 0) t[r][q] = fma(w.b[1], s, t[r][q]);
                else if (dr == 1) t[r][q] = fma(w.b[2], s, t[r][q]);
            }
        }
    }
#pragma unroll
    for (int r = 0; r < RM; r++)
#pragma unroll
        for (int q = 0; q < 4; q++) {
            out[r][q] = fma(w.a[2], t[r][q], L.full[r][q]);      // this plane is h+1 of the plane that completes
            L.full[r][q] = fma(w.a[1], t[r][q], L.next[r][q]);   // and h of the next one
            L.next[r][q] = fma(w.a[0], t[r][q], 0.0);            // and h-1 of the one after
        }
}

template <int FORM>
__global__ void __launch_bounds__(k3Threads, 1)
k_stencil3d_tb(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ Geom3DTB g,
               const __grid_constant__ Weights3D w) {
    static_assert(FORM == LORA_FORM_STAR7 || FORM == LORA_FORM_SEP3, "fused 3-D sweeps: 7-point and separable forms");
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    double *edge = reinterpret_cast<double *>(smem_raw + k3Stages * kT3StageBytes);  // [2][k3Warps][2][k3TileCols]
    uint64_t *full = reinterpret_cast<uint64_t *>(smem_raw + k3Stages * kT3StageBytes + kT3EdgeBytes);
    uint64_t *empty = full + k3Stages;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    const int tile_n = blockIdx.x % g.tiles_n, tile_m = blockIdx.x / g.tiles_n;
    const int h0 = (int)(g.h_lo + (long long)blockIdx.y * g.planes_per_chunk);  // first output plane of the chunk
    const int H = (int)min((long long)g.planes_per_chunk, g.h_hi - h0);
    const int nin = H + 4;                       // level-0 planes h0-2 .. h0+H+1
    const int R0 = tile_m * kT3OutRows, C0 = tile_n * kT3OutCols;  // first output row / column of the tile
    // TMA box origin in padded coordinates: region row -1 = interior row R0 - 2 = padded row R0; region column -2 =
    // interior column C0 - 6 = padded column C0 - 2 (out-of-bounds coordinates are zero-filled)
    const int box_c = C0 - 2, box_r = R0, box_h = h0 - 1;

    if (threadIdx.x == 0) {
#pragma unroll
        for (int k = 0; k < k3Stages; k++) {
            mbar_init(&full[k], 1);
            mbar_init(&empty[k], k3Warps);
        }
        fence_barrier_init();
#pragma unroll
        for (int k = 0; k < k3Stages; k++)
            if (k < nin) {
                mbar_arrive_expect_tx(&full[k], kT3BoxRows * k3BoxCols * 8);
                tma_load_3d(smem_raw + k * kT3StageBytes, &tmap, box_c, box_r, box_h + k, &full[k]);
            }
    }
    __syncthreads();

    // this lane's cells: region rows RM*warp .. +RM-1, region columns 4*lane .. +3
    const int gr0 = R0 + RM * warp - 1;  // interior row of the lane's first row
    const int gc0 = C0 + 4 * lane - 4;   // interior column of its first column
    bool rowin[RM], colin[4];
#pragma unroll
    for (int r = 0; r < RM; r++) rowin[r] = gr0 + r >= 0 && gr0 + r < g.m;
#pragma unroll
    for (int q = 0; q < 4; q++) colin[q] = gc0 + q >= 0 && gc0 + q < g.n;
    // what this lane stores: region rows 1 .. kT3Rows-2, lanes 1 .. 30, inside the grid
    const bool lane_stores = lane >= 1 && lane <= 30 && gc0 < g.n;
    const int cols_left = g.n - gc0;
    double *optr = g.out + (long long)(h0 + 1) * g.plane_pitch + (long long)(gr0 + 2) * g.row_pitch + 4 + gc0;

    LevelState L1, L2;
#pragma unroll
    for (int r = 0; r < RM; r++)
#pragma unroll
        for (int q = 0; q < 4; q++) L1.full[r][q] = L1.next[r][q] = L2.full[r][q] = L2.next[r][q] = 0.0;

    // consumer release + producer duty (as in stencil3d.cu): the stage's values are in registers
    auto release_stage = [&](int i, int slot) {
        __syncwarp();
        if (lane == 0) {
            mbar_arrive(&empty[slot]);  // this warp no longer needs the stage
            const int nx = i - 1 + k3Stages;
            if (warp == 0 && i >= 1 && nx < nin) {  // refill the slot every warp released one plane ago
                const int ps = (i - 1) % k3Stages;
                mbar_wait(&empty[ps], ((i - 1) / k3Stages) & 1);
                mbar_arrive_expect_tx(&full[ps], kT3BoxRows * k3BoxCols * 8);
                tma_load_3d(smem_raw + ps * kT3StageBytes, &tmap, box_c, box_r, box_h + nx, &full[ps]);
            }
        }
    };

    for (int i = 0; i < nin; i++) {
        // ---- level 0 -> level 1: plane q = h0 - 2 + i arrives, level-1 plane q - 1 completes
        const int slot = i % k3Stages;
        mbar_wait(&full[slot], (i / k3Stages) & 1);
        const double *tile = reinterpret_cast<const double *>(smem_raw + slot * kT3StageBytes);
        const int j1 = h0 - 3 + i;  // the level-1 plane that completes now
        const bool plane1_in = j1 >= 0 && j1 < g.h;
        double V[RM][4];
        if constexpr (FORM == LORA_FORM_STAR7) {
            double X[RM + 2][8];  // rows RM*warp-1 .. RM*warp+RM, columns 4*lane-2 .. 4*lane+5 (region coordinates)
#pragma unroll
            for (int rr = 0; rr < RM + 2; rr++) {
                const double2 *rowp = reinterpret_cast<const double2 *>(tile + (RM * warp + rr) * k3BoxCols + 4 * lane);
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    if ((rr == 0 || rr == RM + 1) && (k == 0 || k == 3)) continue;  // corners are not part of a star
                    const double2 v = rowp[k];
                    X[rr][2 * k] = v.x;
                    X[rr][2 * k + 1] = v.y;
                }
            }
            release_stage(i, slot);
#pragma unroll
            for (int r = 0; r < RM; r++)
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    const double xc = X[r + 1][q + 2];
                    const double v = star_inplane(w, xc, X[r + 1][q + 1], X[r + 1][q + 3], X[r][q + 2], X[r + 2][q + 2]);
                    V[r][q] = fma(w.star[6], xc, L1.full[r][q]);  // this plane is h+1 of level-1 plane q-1
                    L1.full[r][q] = L1.next[r][q] + v;             // plane q: born one plane ago + its in-plane part
                    L1.next[r][q] = fma(w.star[5], xc, 0.0);       // this plane is h-1 of level-1 plane q+1
                }
        } else {
            sep3_level(w, [&](int rr, double (&row)[8]) {
                const double2 *rowp = reinterpret_cast<const double2 *>(tile + (RM * warp + rr) * k3BoxCols + 4 * lane);
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const double2 v = rowp[k];
                    row[2 * k] = v.x;
                    row[2 * k + 1] = v.y;
                }
            }, L1, V);
            release_stage(i, slot);
        }
        // level 1 lives at an odd time: its halo is zero (S2) -- outside the interior nothing is computed
#pragma unroll
        for (int r = 0; r < RM; r++)
#pragma unroll
            for (int q = 0; q < 4; q++)
                if (!(plane1_in && rowin[r] && colin[q])) V[r][q] = 0.0;
        if (i < 2) continue;  // CTA-uniform: level-1 planes before h0 - 1 are not needed

        // ---- level 1 -> level 2: rows above / below through shared memory, columns left / right by shuffle
        double *eb = edge + (size_t)(i & 1) * (k3Warps * 2 * k3TileCols);
        {
            double2 *top = reinterpret_cast<double2 *>(eb + (warp * 2 + 0) * k3TileCols + 4 * lane);
            double2 *bot = reinterpret_cast<double2 *>(eb + (warp * 2 + 1) * k3TileCols + 4 * lane);
            top[0] = make_double2(V[0][0], V[0][1]);
            top[1] = make_double2(V[0][2], V[0][3]);
            bot[0] = make_double2(V[RM - 1][0], V[RM - 1][1]);
            bot[1] = make_double2(V[RM - 1][2], V[RM - 1][3]);
        }
        __syncthreads();  // one barrier per plane: the other buffer is not touched before everybody has passed this one again
        double lf[RM], rt[RM];
#pragma unroll
        for (int r = 0; r < RM; r++) {
            lf[r] = __shfl_up_sync(kFull, V[r][3], 1);    // lane 0 gets its own value back: its columns are never stored
            rt[r] = __shfl_down_sync(kFull, V[r][0], 1);  // likewise lane 31
        }
        // warp 0 has nobody above and warp 7 nobody below: their outer rows are never stored, any value will do
        const int wa = warp > 0 ? warp - 1 : 0, wb = warp < k3Warps - 1 ? warp + 1 : k3Warps - 1;
        const bool emit = i >= 4;  // level-2 plane h0 + i - 4 completes
        double O[RM][4];
        if constexpr (FORM == LORA_FORM_STAR7) {
            double up[4], dn[4];
            {
                const double2 *a = reinterpret_cast<const double2 *>(eb + (wa * 2 + 1) * k3TileCols + 4 * lane);
                const double2 *b = reinterpret_cast<const double2 *>(eb + (wb * 2 + 0) * k3TileCols + 4 * lane);
                const double2 a0 = a[0], a1 = a[1], b0 = b[0], b1 = b[1];
                up[0] = a0.x, up[1] = a0.y, up[2] = a1.x, up[3] = a1.y;
                dn[0] = b0.x, dn[1] = b0.y, dn[2] = b1.x, dn[3] = b1.y;
            }
#pragma unroll
            for (int r = 0; r < RM; r++)
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    const double xc = V[r][q];
                    const double v = star_inplane(w, xc, q > 0 ? V[r][q - 1] : lf[r], q < 3 ? V[r][q + 1] : rt[r],
                                                  r > 0 ? V[r - 1][q] : up[q], r < RM - 1 ? V[r + 1][q] : dn[q]);
                    O[r][q] = fma(w.star[6], xc, L2.full[r][q]);
                    L2.full[r][q] = L2.next[r][q] + v;
                    L2.next[r][q] = fma(w.star[5], xc, 0.0);
                }
        } else {
            // window rows of level 1: the neighbour warps' edge rows (8 columns from 4*lane - 2; lanes 0 and 31 read a
            // shifted window -- their columns are never stored), own rows from registers + the shuffled columns
            const int cbase = min(max(4 * lane - 2, 0), k3TileCols - 8);
            sep3_level(w, [&](int rr, double (&row)[8]) {
                if (rr == 0 || rr == RM + 1) {
                    const double2 *e = reinterpret_cast<const double2 *>(
                        eb + ((rr == 0 ? wa * 2 + 1 : wb * 2 + 0)) * k3TileCols + cbase);
#pragma unroll
                    for (int k = 0; k < 4; k++) {
                        const double2 v = e[k];
                        row[2 * k] = v.x;
                        row[2 * k + 1] = v.y;
                    }
                } else {
                    row[0] = 0.0;
                    row[1] = lf[rr - 1];
#pragma unroll
                    for (int q = 0; q < 4; q++) row[2 + q] = V[rr - 1][q];
                    row[6] = rt[rr - 1];
                    row[7] = 0.0;
                }
            }, L2, O);
        }
#pragma unroll
        for (int r = 0; r < RM; r++) {
            const int rr = RM * warp + r;  // region row
            if (emit && lane_stores && rr >= 1 && rr <= kT3Rows - 2 && gr0 + r < g.m) {
                double *op = optr + r * g.row_pitch;
                if (cols_left >= 4) {
                    if (g.vec4) {
                        st_global_v4(op, O[r][0], O[r][1], O[r][2], O[r][3]);
                    } else {
                        st_global_v2(op, O[r][0], O[r][1]);
                        st_global_v2(op + 2, O[r][2], O[r][3]);
                    }
                } else {
#pragma unroll
                    for (int q = 0; q < 4; q++)
                        if (q < cols_left) op[q] = O[r][q];
                }
            }
        }
        if (emit) optr += g.plane_pitch;
    }
}

}  // namespace

cudaError_t kernels_init_3d_tb() {
    cudaError_t e = cudaFuncSetAttribute(k_stencil3d_tb<LORA_FORM_STAR7>, cudaFuncAttributeMaxDynamicSharedMemorySize, kT3Smem);
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(k_stencil3d_tb<LORA_FORM_SEP3>, cudaFuncAttributeMaxDynamicSharedMemorySize, kT3Smem);
}

cudaError_t launch_3d_tb(int form, const CUtensorMap &tmap, const Geom3DTB &g, const Weights3D &w, cudaStream_t s) {
    if (form != LORA_FORM_STAR7 && form != LORA_FORM_SEP3) return cudaErrorInvalidValue;
    const long long planes = g.h_hi - g.h_lo;
    if (planes <= 0) return cudaSuccess;
    const int chunks = (int)((planes + g.planes_per_chunk - 1) / g.planes_per_chunk);
    dim3 grid(g.tiles_m * g.tiles_n, chunks);
    if (form == LORA_FORM_STAR7)
        k_stencil3d_tb<LORA_FORM_STAR7><<<grid, k3Threads, kT3Smem, s>>>(tmap, g, w);
    else
        k_stencil3d_tb<LORA_FORM_SEP3><<<grid, k3Threads, kT3Smem, s>>>(tmap, g, w);
    return cudaGetLastError();
}

}  // namespace lora
