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
//   * a lane owns RM = 4 rows x 4 columns: two levels x two carried accumulators x 16 cells = 64 doubles of state; a
//     7-point (or separable) operator needs only TWO accumulators per level between planes -- the plane that completes
//     is handed on at once;
//   * level 1 reaches level 2 without a tile round trip: a lane's own 4 x 4 values stay in registers, the columns
//     left / right come from the neighbour lanes by warp shuffle, and only the first / last row of every warp goes
//     through shared memory (2 x 128 doubles per warp and plane, double-buffered, ONE __syncthreads per plane);
//   * overlapped tiling along the rows only: every level is computed on the full 32 x 128 thread tile and the outer
//     rows only feed their neighbours (a CTA writes 30 rows); along the columns a tile writes ALL 128 columns it owns
//     (512 columns = 4 tiles, not 5 of 120): the two level-1 columns just outside the tile, which level 2 needs, are
//     computed as ONE extra cell per lane (lanes 0 .. 2 RM - 1: RM rows x {left, right}) from the level-0 box, which
//     covers them anyway, and handed to lanes 0 / 31 by shuffle.  A CTA reads a 34 x 132 box per 30 x 128 outputs:
//     17.4 B of DRAM traffic per cell per TWO launches against 2 x 16.8 unfused.
//
// Reference semantics (S2) under fusion: a fused sweep starts at an even time -- level 0 sees the caller's halo, which
// is physically in the source buffer's ring (the host copies the ring of buffer 0 into buffer 1 before the first sweep
// that reads buffer 1 and clears it again afterwards: lora_plan_run) -- and level 1 sits at an odd time, whose halo is
// ZERO: level-1 cells outside the interior are forced to zero before they feed level 2.  Same operations in the same
// order as two unfused launches: results are bit-identical.
#include <cstdlib>

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

// the extra level-1 cell of a lane (region column -1 or 128): same operator, same operation order, operands straight
// from the level-0 box.  c points at the cell's centre in the box, `pitch` doubles per box row.
struct EdgeState {
    double full, next;
};
template <int FORM>
__device__ __forceinline__ double edge_cell(const Weights3D &w, const double *c, int pitch, EdgeState &E) {
    double out;
    if constexpr (FORM == LORA_FORM_SEP3) {
        double s[3];
#pragma unroll
        for (int k = 0; k < 3; k++) {
            const double *r = c + (k - 1) * pitch;
            s[k] = w.c[0] * r[-1];
            s[k] = fma(w.c[1], r[0], s[k]);
            s[k] = fma(w.c[2], r[1], s[k]);
        }
        double t = w.b[0] * s[0];
        t = fma(w.b[1], s[1], t);
        t = fma(w.b[2], s[2], t);
        out = fma(w.a[2], t, E.full);
        E.full = fma(w.a[1], t, E.next);
        E.next = fma(w.a[0], t, 0.0);
    } else {
        const double xc = c[0];
        double a = w.star[0] * xc;
        a = fma(w.star[1], c[-1], a);
        a = fma(w.star[2], c[1], a);
        a = fma(w.star[3], c[-pitch], a);
        a = fma(w.star[4], c[pitch], a);
        out = fma(w.star[6], xc, E.full);
        E.full = E.next + a;
        E.next = fma(w.star[5], xc, 0.0);
    }
    return out;
}

// Column ownership: lane l owns region columns {2l, 2l+1} (pair A) and {64+2l, 64+2l+1} (pair B) of every row.  A
// 128-bit shared-memory access then has a lane stride of 16 bytes -- conflict-free -- where 4 consecutive columns per
// lane (stride 32 bytes, stencil3d.cu) cost two wavefronts per quarter-warp; the columns left / right of a pair come
// from the neighbour lanes by CIRCULAR warp shuffle: lane 31's pair A ends at column 63, whose right neighbour is lane
// 0's pair B, and vice versa; only column -1 (lane 0) and column 128 (lane 31) are not owned by anybody in the warp.
// A window row is {LA, A0, A1, RA, LB, B0, B1, RB}.
__device__ __forceinline__ void window_row(double a0, double a1, double b0, double b1, int lane, double edge_l, double edge_r,
                                           double (&row)[8]) {
    const int prev = (lane + 31) & 31, next = (lane + 1) & 31;
    const double la = __shfl_sync(kFull, a1, prev), lb = __shfl_sync(kFull, b1, prev);
    const double ra = __shfl_sync(kFull, a0, next), rb = __shfl_sync(kFull, b0, next);
    row[0] = lane == 0 ? edge_l : la;
    row[1] = a0;
    row[2] = a1;
    row[3] = lane == 31 ? rb : ra;
    row[4] = lane == 0 ? la : lb;
    row[5] = b0;
    row[6] = b1;
    row[7] = lane == 31 ? edge_r : rb;
}
// own cell q (0, 1: pair A; 2, 3: pair B) sits at row[kCell[q]]
#define LORA_CELL(q) ((q) < 2 ? 1 + (q) : 3 + (q))

// One level of the operator over the lane's RM x 4 cells, the window delivered row by row (only one window row is
// live at a time).  rowfn(rr, row): window row rr = region row RM*warp - 1 + rr; `full_row` false: only the own cells
// of that row are needed (rows above / below of a star).  Operation order = stencil3d.cu's pushes, so fused and unfused
// launches give the same bits.
template <int FORM, class RowFn>
__device__ __forceinline__ void level(const Weights3D &w, RowFn rowfn, LevelState &L, double (&out)[RM][4]) {
    if constexpr (FORM == LORA_FORM_SEP3) {
        // s = c * row along n, t = b * s along m, then the two carried plane accumulators (a along h)
        double t[RM][4];
#pragma unroll
        for (int rr = 0; rr < RM + 2; rr++) {
            double row[8];
            rowfn(rr, row, true);
#pragma unroll
            for (int q = 0; q < 4; q++) {
                double s = w.c[0] * row[LORA_CELL(q) - 1];
                s = fma(w.c[1], row[LORA_CELL(q)], s);
                s = fma(w.c[2], row[LORA_CELL(q) + 1], s);
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
    } else {
        // 7-point star: centre, n-1, n+1, m-1, m+1 in-plane; the plane taps through the carried accumulators
        double v[RM][4], up[4], cur[8], nxt[8];
        rowfn(0, cur, false);
#pragma unroll
        for (int q = 0; q < 4; q++) up[q] = cur[LORA_CELL(q)];
        rowfn(1, cur, true);
#pragma unroll
        for (int r = 0; r < RM; r++) {
            rowfn(r + 2, nxt, r + 2 <= RM);
#pragma unroll
            for (int q = 0; q < 4; q++) {
                const double xc = cur[LORA_CELL(q)];
                double a = w.star[0] * xc;
                a = fma(w.star[1], cur[LORA_CELL(q) - 1], a);
                a = fma(w.star[2], cur[LORA_CELL(q) + 1], a);
                a = fma(w.star[3], up[q], a);
                a = fma(w.star[4], nxt[LORA_CELL(q)], a);
                v[r][q] = a;
                out[r][q] = fma(w.star[6], xc, L.full[r][q]);  // this plane is h+1 of the plane that completes
                L.full[r][q] = L.next[r][q] + a;                // plane q: born one plane ago + its in-plane part
                L.next[r][q] = fma(w.star[5], xc, 0.0);         // this plane is h-1 of the plane after
                up[q] = xc;
            }
#pragma unroll
            for (int k = 0; k < 8; k++) cur[k] = nxt[k];
        }
        (void)v;
    }
}

// SLAB: the launch serves neighbouring slabs (mirror stores, flags); the plain instantiation carries none of that code
// (3-4 % on a single GPU: 502 vs 481 GStencil/s for the separable form at 512^3)
template <int FORM, bool SLAB>
__global__ void __launch_bounds__(k3Threads, 1)
k_stencil3d_tb(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ Geom3DTB g,
               const __grid_constant__ Weights3D w) {
    static_assert(FORM == LORA_FORM_STAR7 || FORM == LORA_FORM_SEP3, "fused 3-D sweeps: 7-point and separable forms");
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    double *edge = reinterpret_cast<double *>(smem_raw + k3Stages * kT3StageBytes);  // [2][k3Warps][2][k3TileCols]
    uint64_t *full = reinterpret_cast<uint64_t *>(smem_raw + k3Stages * kT3StageBytes + kT3EdgeBytes);
    uint64_t *empty = full + k3Stages;
    const int warp = uniform_warp_id(), lane = threadIdx.x & 31;
    volatile int *guard = reinterpret_cast<volatile int *>(empty + k3Stages) + warp;  // stage release, see level1

    const int tile_n = blockIdx.x % g.tiles_n, tile_m = blockIdx.x / g.tiles_n;
    const int seg = seg_of(g.sg, blockIdx.y);  // a slab's hi-band chunk comes first in blockIdx.y, i.e. in dispatch order
    const int h0 = (int)(g.sg.lo[seg] + (blockIdx.y - g.sg.first[seg]) * g.sg.chunk[seg]);  // first output plane of the chunk
    const int H = (int)min(g.sg.chunk[seg], g.sg.hi[seg] - h0);
    const int nin = H + 4;                       // level-0 planes h0-2 .. h0+H+1
    const int R0 = tile_m * kT3OutRows, C0 = tile_n * kT3OutCols;  // first output row / column of the tile
    // TMA box origin in padded coordinates: region row -1 = interior row R0 - 2 = padded row R0; region column -2 =
    // interior column C0 - 2 = padded column C0 + 2 (out-of-bounds coordinates are zero-filled)
    const int box_c = C0 + 2, box_r = R0, box_h = h0 - 1;

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

    // this lane's cells: region rows RM*warp .. +RM-1; region columns 2*lane, 2*lane+1 (A) and 64+2*lane, +1 (B)
    const int gr0 = R0 + RM * warp - 1;         // interior row of the lane's first row
    const int gcA = C0 + 2 * lane;              // interior column of A0
    const int gcB = gcA + 64;                   // interior column of B0
    bool rowin[RM], colin[4];
#pragma unroll
    for (int r = 0; r < RM; r++) rowin[r] = gr0 + r >= 0 && gr0 + r < g.m;
    colin[0] = gcA < g.n;
    colin[1] = gcA + 1 < g.n;
    colin[2] = gcB < g.n;
    colin[3] = gcB + 1 < g.n;
    // level-1 cells outside the interior are zero (S2); only tiles that touch the rim of the grid have any
    const bool rim = R0 == 0 || R0 + kT3Rows - 1 > g.m || C0 + k3TileCols > g.n;
    // the lane's extra level-1 cell: lane e < 2 RM owns row e % RM of region column -1 (e < RM) or 128
    const int er = lane % RM, eside = (lane / RM) & 1;
    const int ebox = ((RM * warp + er + 1) * k3BoxCols) + (eside ? k3TileCols + 2 : 1);  // its centre in the box
    const int egc = eside ? C0 + k3TileCols : C0 - 1;
    const bool ein = gr0 + er >= 0 && gr0 + er < g.m && egc >= 0 && egc < g.n;
    double *optr = g.out + (long long)(h0 + 1) * g.plane_pitch + (long long)(gr0 + 2) * g.row_pitch + 4 + gcA;
    // multi-GPU slabs: planes [mlo, mhi) of this chunk are stored a second time at + mirror (the neighbour slab's ghost
    // planes, peer memory over NVLink).  Both band chunks report to their flag when their CTAs finish: an early report of
    // the lo band right after its planes are stored (as the unfused kernel does) puts a conditional CTA barrier into the
    // plane loop, which cost this kernel 3-4 % (523 vs 551 GStencil/s for the separable form at 512^3)
    const long long mirror = SLAB ? g.sg.mirror[seg] : 0;
    const int mlo = SLAB ? (int)g.sg.mlo[seg] : 0, mhi = SLAB ? (int)g.sg.mhi[seg] : 0;

    LevelState L1, L2;
    EdgeState E1{0.0, 0.0};
#pragma unroll
    for (int r = 0; r < RM; r++)
#pragma unroll
        for (int q = 0; q < 4; q++) L1.full[r][q] = L1.next[r][q] = L2.full[r][q] = L2.next[r][q] = 0.0;

    // ---- level 0 -> level 1: plane q = h0 - 2 + i arrives, level-1 plane q - 1 completes (V: the lane's cells, EV: its
    // extra cell)
    auto level1 = [&](int i, double (&V)[RM][4], double &EV) {
        const int slot = i % k3Stages;
        mbar_wait(&full[slot], (i / k3Stages) & 1);
        const double *tile = reinterpret_cast<const double *>(smem_raw + slot * kT3StageBytes);
        const int j1 = h0 - 3 + i;  // the level-1 plane that completes now
        const bool plane1_in = j1 >= 0 && j1 < g.h;
        int gw = 0;  // a word computed from every value loaded from the stage (the guard store below)
        level<FORM>(w, [&](int rr, double (&row)[8], bool full_row) {
            // box row RM*warp + rr; box column = region column + 2
            const double *rowp = tile + (RM * warp + rr) * k3BoxCols + 2;
            const double2 a = *reinterpret_cast<const double2 *>(rowp + 2 * lane);
            const double2 b = *reinterpret_cast<const double2 *>(rowp + 64 + 2 * lane);
            gw ^= __double2hiint(a.x) ^ __double2hiint(b.x);
            if (full_row) {
                double e = 0.0;
                if (lane == 0) e = rowp[-1];       // region column -1
                if (lane == 31) e = rowp[128];     // region column 128
                gw ^= __double2hiint(e);
                window_row(a.x, a.y, b.x, b.y, lane, e, e, row);
            } else {
                row[1] = a.x, row[2] = a.y, row[5] = b.x, row[6] = b.y;
            }
        }, L1, V);
        EV = edge_cell<FORM>(w, tile + ebox, k3BoxCols, E1);
        // Release the stage and refill the one released a plane ago -- once the stage's values have ARRIVED in registers
        // (stencil3d.cu, plane_phase): the store of a word computed from all of them cannot issue before the loads have
        // completed, and the arrive follows it in program order
        *guard = gw ^ __double2hiint(EV);
        // (the producer's wait is done by all of warp 0, a warp-uniform branch: a spin loop under `lane == 0` makes
        // ptxas treat the whole plane loop as divergent, and the weights then live in vector registers)
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[slot]);
        const int nx = i - 1 + k3Stages;
        if (warp == 0 && i >= 1 && nx < nin) {
            const int ps = (i - 1) % k3Stages;
            mbar_wait(&empty[ps], ((i - 1) / k3Stages) & 1);
            if (lane == 0) {
                mbar_arrive_expect_tx(&full[ps], kT3BoxRows * k3BoxCols * 8);
                tma_load_3d(smem_raw + ps * kT3StageBytes, &tmap, box_c, box_r, box_h + nx, &full[ps]);
            }
        }
        // level 1 lives at an odd time: its halo is zero (S2) -- outside the interior nothing is computed
        if (rim || !plane1_in) {  // CTA-uniform
#pragma unroll
            for (int r = 0; r < RM; r++)
#pragma unroll
                for (int q = 0; q < 4; q++)
                    if (!(plane1_in && rowin[r] && colin[q])) V[r][q] = 0.0;
        }
        if (!(plane1_in && ein)) EV = 0.0;
    };

    // The first two planes only warm level 1 up (level-1 planes before h0 - 1 are not needed).  Peeled, not skipped with
    // a `continue`: with the early exit in the loop ptxas treats the loop as divergent and keeps the weights in vector
    // registers; peeled, the weights are uniform-register operands and the kernel needs ~30 registers less.
    {
        double V[RM][4], EV;
        level1(0, V, EV);
        level1(1, V, EV);
    }
    for (int i = 2; i < nin; i++) {
        double V[RM][4], EV;
        level1(i, V, EV);

        // ---- level 1 -> level 2: the rows above / below come through shared memory (first / last row of every warp),
        // everything else from the lane's own registers and its neighbours' by shuffle
        double *eb = edge + (size_t)(i & 1) * (k3Warps * 2 * kT3EdgePitch) + 2;  // column c of an edge row at [c]
        {
            double *top = eb + (warp * 2 + 0) * kT3EdgePitch, *bot = eb + (warp * 2 + 1) * kT3EdgePitch;
            *reinterpret_cast<double2 *>(top + 2 * lane) = make_double2(V[0][0], V[0][1]);
            *reinterpret_cast<double2 *>(top + 64 + 2 * lane) = make_double2(V[0][2], V[0][3]);
            *reinterpret_cast<double2 *>(bot + 2 * lane) = make_double2(V[RM - 1][0], V[RM - 1][1]);
            *reinterpret_cast<double2 *>(bot + 64 + 2 * lane) = make_double2(V[RM - 1][2], V[RM - 1][3]);
            if constexpr (FORM == LORA_FORM_SEP3) {  // the window rows above / below are full rows: columns -1, 128 too
                if (lane < 2 * RM && (er == 0 || er == RM - 1))
                    (er == 0 ? top : bot)[eside ? k3TileCols : -1] = EV;
            }
        }
        __syncthreads();  // one barrier per plane: the other buffer is not touched before everybody has passed this one again
        // warp 0 has nobody above and warp 7 nobody below: their outer rows are never stored, any value will do
        const int wa = warp > 0 ? warp - 1 : 0, wb = warp < k3Warps - 1 ? warp + 1 : k3Warps - 1;
        double O[RM][4];
        level<FORM>(w, [&](int rr, double (&row)[8], bool full_row) {
            double a0, a1, b0, b1, el = 0.0, er_ = 0.0;
            if (rr == 0 || rr == RM + 1) {
                const double *e = eb + (rr == 0 ? wa * 2 + 1 : wb * 2 + 0) * kT3EdgePitch;
                const double2 a = *reinterpret_cast<const double2 *>(e + 2 * lane);
                const double2 b = *reinterpret_cast<const double2 *>(e + 64 + 2 * lane);
                a0 = a.x, a1 = a.y, b0 = b.x, b1 = b.y;
                if (full_row) {
                    if (lane == 0) el = e[-1];
                    if (lane == 31) er_ = e[k3TileCols];
                }
            } else {
                a0 = V[rr - 1][0], a1 = V[rr - 1][1], b0 = V[rr - 1][2], b1 = V[rr - 1][3];
                if (full_row) {  // the lanes that own this row's extra cells hand them over
                    el = __shfl_sync(kFull, EV, rr - 1);
                    er_ = __shfl_sync(kFull, EV, RM + rr - 1);
                }
            }
            if (full_row) {
                window_row(a0, a1, b0, b1, lane, el, er_, row);
            } else {
                row[1] = a0, row[2] = a1, row[5] = b0, row[6] = b1;
            }
        }, L2, O);
        if (i >= 4) {  // level-2 plane h0 + i - 4 is complete
            const int hout = h0 + i - 4;
            const bool mirrored = SLAB && mirror != 0 && hout >= mlo && hout < mhi;  // CTA-uniform
#pragma unroll
            for (int r = 0; r < RM; r++) {
                const int rr = RM * warp + r;  // region row
                if (rr >= 1 && rr <= kT3Rows - 2 && gr0 + r < g.m) {
                    double *op = optr + r * g.row_pitch;
                    if (colin[0]) {
                        if (gcA + 1 < g.n) st_global_v2(op, O[r][0], O[r][1]);
                        else op[0] = O[r][0];
                    }
                    if (colin[2]) {
                        if (gcB + 1 < g.n) st_global_v2(op + 64, O[r][2], O[r][3]);
                        else op[64] = O[r][2];
                    }
                    if (mirrored) {
                        double *om = op + mirror;
                        if (colin[0]) {
                            if (gcA + 1 < g.n) st_global_v2(om, O[r][0], O[r][1]);
                            else om[0] = O[r][0];
                        }
                        if (colin[2]) {
                            if (gcB + 1 < g.n) st_global_v2(om + 64, O[r][2], O[r][3]);
                            else om[64] = O[r][2];
                        }
                    }
                }
            }
            optr += g.plane_pitch;
        }
    }
    if (SLAB && g.sg.flag[seg] != nullptr) {  // a band chunk: tell the neighbour once all its CTAs have stored
        __threadfence_system();
        __syncthreads();
        if (threadIdx.x == 0) seg_arrive(g.sg, seg);
    }
}
#undef LORA_CELL

template <int FORM, bool SLAB>
cudaError_t opt_in_3d_tb() {
    return cudaFuncSetAttribute(k_stencil3d_tb<FORM, SLAB>, cudaFuncAttributeMaxDynamicSharedMemorySize, kT3Smem);
}

}  // namespace

cudaError_t kernels_init_3d_tb() {
    cudaError_t e;
    if ((e = opt_in_3d_tb<LORA_FORM_STAR7, false>()) != cudaSuccess) return e;
    if ((e = opt_in_3d_tb<LORA_FORM_STAR7, true>()) != cudaSuccess) return e;
    if ((e = opt_in_3d_tb<LORA_FORM_SEP3, false>()) != cudaSuccess) return e;
    return opt_in_3d_tb<LORA_FORM_SEP3, true>();
}

cudaError_t launch_3d_tb(int form, const CUtensorMap &tmap, const Geom3DTB &g, const Weights3D &w, cudaStream_t s) {
    if (form != LORA_FORM_STAR7 && form != LORA_FORM_SEP3) return cudaErrorInvalidValue;
    const int chunks = (int)g.sg.first[g.sg.nseg];
    if (chunks <= 0) return cudaSuccess;
    dim3 grid(g.tiles_m * g.tiles_n, chunks);
    bool slab = false;  // any segment with a mirror or a flag
    for (int i = 0; i < g.sg.nseg; i++) slab = slab || g.sg.mirror[i] != 0 || g.sg.flag[i] != nullptr;
    static const bool force_slab = getenv("LORA_DEBUG_SLAB3D") != nullptr;  // measure what the slab code costs by itself
    slab = slab || force_slab;
    if (form == LORA_FORM_STAR7) {
        if (slab) k_stencil3d_tb<LORA_FORM_STAR7, true><<<grid, k3Threads, kT3Smem, s>>>(tmap, g, w);
        else k_stencil3d_tb<LORA_FORM_STAR7, false><<<grid, k3Threads, kT3Smem, s>>>(tmap, g, w);
    } else {
        if (slab) k_stencil3d_tb<LORA_FORM_SEP3, true><<<grid, k3Threads, kT3Smem, s>>>(tmap, g, w);
        else k_stencil3d_tb<LORA_FORM_SEP3, false><<<grid, k3Threads, kT3Smem, s>>>(tmap, g, w);
    }
    return cudaGetLastError();
}

}  // namespace lora
