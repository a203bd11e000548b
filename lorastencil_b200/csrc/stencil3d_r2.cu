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
//   SEP5     15 FMA: a[dh] * b[dr] * c[dc] (rank 1 along every axis): see k_stencil3d_r2_sep
// In-plane neighbours come through L1; no shared memory, no barriers, so rows / columns outside the grid simply retire.
// A chunk of planes costs 4 extra plane reads of warm-up.  Measured on B200 at 512^3 (profiles/r2_extensions_*.json):
// STAR13 225, SEP5 195, HSEP5 179, DIRECT125 57 GStencil/s -- first versions, short of memory-level parallelism (one
// dependent chain of planes per thread), see DESIGN.md section 2.5.
#include <cstdlib>

#include "common.cuh"
#include "kernels.h"
#include "../../include/lorastencil.h"

namespace lora {

namespace {

constexpr int kR2Cols = 128;  // threads along the columns
constexpr int kR2Rows = 2;    // rows per CTA

// One cell per thread.  UNROLL = 1: plain loads, one plane per trip (variant 0).  UNROLL = 4: read-only loads
// (LDG.CONSTANT), no aliasing between the two buffers, several planes per trip so that the loads of the next planes may
// be issued ahead of this plane's store (variant 1; the 125-tap form stays at one plane per trip: two need 249 registers).
template <int FORM, int UNROLL>
__global__ void __launch_bounds__(kR2Cols * kR2Rows)
k_stencil3d_r2(const __grid_constant__ Geom3DR2 g, const __grid_constant__ WeightsR2 w) {
    const int c = blockIdx.x * kR2Cols + threadIdx.x;
    const int r = blockIdx.y * kR2Rows + threadIdx.y;
    if (c >= g.n || r >= g.m) return;
    const long long q_lo = g.lo + (long long)blockIdx.z * g.planes_per_chunk;     // first output plane of this chunk
    const long long q_hi = min(q_lo + (long long)g.planes_per_chunk, g.hi);       // one past the last
    // cell (plane p, r, c) of the interior sits at padded (p + 2, r + 2, c + 4)
    const long long cell = (long long)(r + 2) * g.row_pitch + 4 + c;
    const double *__restrict__ in = g.in + cell;
    double *__restrict__ out = g.out + cell;
    auto ld = [](const double *q) { return UNROLL > 1 ? __ldg(q) : *q; };

    double acc[5] = {0.0, 0.0, 0.0, 0.0, 0.0};  // acc[k]: output plane p - 2 + k while plane p is being pushed
#pragma unroll(FORM == LORA_FORM_DIRECT125 ? 1 : UNROLL)
    for (long long p = q_lo - 2; p <= q_hi + 1; p++) {
        const double *pl = in + (p + 2) * g.plane_pitch;
        if (FORM == LORA_FORM_STAR13) {
            const double ctr = ld(pl);
            double t = w.w[62] * ctr;  // (0, 0, 0)
#pragma unroll
            for (int d = 1; d <= 2; d++) {
                t = fma(w.w[62 - d], ld(pl - d), t);                      // (0, 0, -d)
                t = fma(w.w[62 + d], ld(pl + d), t);                      // (0, 0, +d)
                t = fma(w.w[62 - 5 * d], ld(pl - d * g.row_pitch), t);    // (0, -d, 0)
                t = fma(w.w[62 + 5 * d], ld(pl + d * g.row_pitch), t);    // (0, +d, 0)
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
                for (int dc = -2; dc <= 2; dc++) v[(dr + 2) * 5 + dc + 2] = ld(pl + dr * g.row_pitch + dc);
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

// Variant 2: TWO adjacent cells per thread (even column counts, 16-byte aligned buffers).  A row of the neighbourhood is
// columns c-2 .. c+3 = three aligned 128-bit loads for two cells instead of five 64-bit loads per cell; the in-plane
// sums of the one-cell kernel are bound by exactly that -- L1 wavefronts (25 unaligned 256-byte warp loads per cell
// and plane) -- and every weight fetched feeds two FMAs.
template <int FORM>
__global__ void __launch_bounds__(kR2Cols * kR2Rows)
k_stencil3d_r2_pair(const __grid_constant__ Geom3DR2 g, const __grid_constant__ WeightsR2 w) {
    const int c = 2 * (blockIdx.x * kR2Cols + threadIdx.x);  // cells c and c + 1; n is even, so both or neither exist
    const int r = blockIdx.y * kR2Rows + threadIdx.y;
    if (c >= g.n || r >= g.m) return;
    const long long q_lo = g.lo + (long long)blockIdx.z * g.planes_per_chunk;
    const long long q_hi = min(q_lo + (long long)g.planes_per_chunk, g.hi);
    const long long cell = (long long)(r + 2) * g.row_pitch + 4 + c;  // padded column c + 4 is even: 16-byte aligned
    const double *__restrict__ in = g.in + cell;
    double *__restrict__ out = g.out + cell;
    auto ld2 = [](const double *q) { return __ldg(reinterpret_cast<const double2 *>(q)); };

    double acc0[5] = {0.0, 0.0, 0.0, 0.0, 0.0}, acc1[5] = {0.0, 0.0, 0.0, 0.0, 0.0};  // cell c / cell c + 1
    for (long long p = q_lo - 2; p <= q_hi + 1; p++) {
        const double *pl = in + (p + 2) * g.plane_pitch;
        if (FORM == LORA_FORM_STAR13) {
            const double2 a = ld2(pl - 2), b = ld2(pl), d = ld2(pl + 2);  // columns (c-2, c-1), (c, c+1), (c+2, c+3)
            double t0 = w.w[62] * b.x, t1 = w.w[62] * b.y;
            t0 = fma(w.w[61], a.y, t0);  // (0, 0, -1)
            t1 = fma(w.w[61], b.x, t1);
            t0 = fma(w.w[63], b.y, t0);  // (0, 0, +1)
            t1 = fma(w.w[63], d.x, t1);
            t0 = fma(w.w[60], a.x, t0);  // (0, 0, -2)
            t1 = fma(w.w[60], a.y, t1);
            t0 = fma(w.w[64], d.x, t0);  // (0, 0, +2)
            t1 = fma(w.w[64], d.y, t1);
#pragma unroll
            for (int k = 1; k <= 2; k++) {
                const double2 up = ld2(pl - k * g.row_pitch), dn = ld2(pl + k * g.row_pitch);
                t0 = fma(w.w[62 - 5 * k], up.x, t0);  // (0, -k, 0)
                t1 = fma(w.w[62 - 5 * k], up.y, t1);
                t0 = fma(w.w[62 + 5 * k], dn.x, t0);  // (0, +k, 0)
                t1 = fma(w.w[62 + 5 * k], dn.y, t1);
            }
            acc0[2] += t0;
            acc1[2] += t1;
            acc0[0] = fma(w.w[62 + 50], b.x, acc0[0]);  // tap (dh, 0, 0) into accumulator k = 2 - dh
            acc1[0] = fma(w.w[62 + 50], b.y, acc1[0]);
            acc0[1] = fma(w.w[62 + 25], b.x, acc0[1]);
            acc1[1] = fma(w.w[62 + 25], b.y, acc1[1]);
            acc0[3] = fma(w.w[62 - 25], b.x, acc0[3]);
            acc1[3] = fma(w.w[62 - 25], b.y, acc1[3]);
            acc0[4] = fma(w.w[62 - 50], b.x, acc0[4]);
            acc1[4] = fma(w.w[62 - 50], b.y, acc1[4]);
        } else {
            double v[5][6];  // v[dr + 2][j] = column c - 2 + j: cell c reads j = dc + 2, cell c + 1 reads j = dc + 3
#pragma unroll
            for (int dr = -2; dr <= 2; dr++) {
                const double *row = pl + dr * g.row_pitch;
                const double2 a = ld2(row - 2), b = ld2(row), d = ld2(row + 2);
                v[dr + 2][0] = a.x, v[dr + 2][1] = a.y, v[dr + 2][2] = b.x;
                v[dr + 2][3] = b.y, v[dr + 2][4] = d.x, v[dr + 2][5] = d.y;
            }
            if (FORM == LORA_FORM_HSEP5) {
                double t0 = 0.0, t1 = 0.0;
#pragma unroll
                for (int i = 0; i < 5; i++)
#pragma unroll
                    for (int j = 0; j < 5; j++) {
                        t0 = fma(w.q[i * 5 + j], v[i][j], t0);
                        t1 = fma(w.q[i * 5 + j], v[i][j + 1], t1);
                    }
#pragma unroll
                for (int k = 0; k < 5; k++) {
                    acc0[k] = fma(w.a[4 - k], t0, acc0[k]);
                    acc1[k] = fma(w.a[4 - k], t1, acc1[k]);
                }
            } else {
#pragma unroll
                for (int k = 0; k < 5; k++) {
                    const double *wk = w.w + (4 - k) * 25;
                    double t0 = acc0[k], t1 = acc1[k];
#pragma unroll
                    for (int i = 0; i < 5; i++)
#pragma unroll
                        for (int j = 0; j < 5; j++) {
                            t0 = fma(wk[i * 5 + j], v[i][j], t0);
                            t1 = fma(wk[i * 5 + j], v[i][j + 1], t1);
                        }
                    acc0[k] = t0;
                    acc1[k] = t1;
                }
            }
        }
        if (p - 2 >= q_lo) st_global_v2(out + p * g.plane_pitch, acc0[0], acc1[0]);
#pragma unroll
        for (int k = 0; k < 4; k++) acc0[k] = acc0[k + 1], acc1[k] = acc1[k + 1];
        acc0[4] = 0.0;
        acc1[4] = 0.0;
    }
}

// Fully separable tables, w[dh][dr][dc] = a[dh] b[dr] c[dc] (form SEP5: the default box3d2r table is one): 15 FMA and
// FIVE loads per cell and plane instead of 25.  A lane owns one column: it sums its column over the five rows with b
// (five aligned, fully coalesced loads), takes the row sums of the two columns either side from its neighbour lanes by
// shuffle and combines them with c, then pushes a[dh] * t into the five accumulators.  Lanes 0, 1, 30, 31 of a warp only
// feed their neighbours: a warp stores 28 columns and adjacent warps overlap by 4 (12.5 % redundant loads -- against
// five times fewer loads overall).
constexpr int kSepOut = 28;  // columns a warp stores

__global__ void __launch_bounds__(kR2Cols * kR2Rows)
k_stencil3d_r2_sep(const __grid_constant__ Geom3DR2 g, const __grid_constant__ WeightsR2 w) {
    const int lane = threadIdx.x & 31;
    const int wbase = (blockIdx.x * (kR2Cols / 32) + (threadIdx.x >> 5)) * kSepOut;  // first column this warp stores
    const int r = blockIdx.y * kR2Rows + threadIdx.y;
    if (wbase >= g.n || r >= g.m) return;  // warp-uniform: a warp is 32 consecutive x of one y
    const int c = wbase - 2 + lane;        // this lane's column; -2 .. n + 1 are cells of the padded row
    const int cl = min(c, g.n + 1);        // lanes beyond that re-read the row's last halo cell (feeds no stored column)
    const bool owner = lane >= 2 && lane < 2 + kSepOut && c < g.n;
    const long long q_lo = g.lo + (long long)blockIdx.z * g.planes_per_chunk;
    const long long q_hi = min(q_lo + (long long)g.planes_per_chunk, g.hi);
    const long long row0 = (long long)(r + 2) * g.row_pitch + 4;
    const double *__restrict__ in = g.in + row0 + cl;
    double *__restrict__ out = g.out + row0 + c;
    constexpr unsigned kFull = 0xffffffffu;

    double acc[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
    for (long long p = q_lo - 2; p <= q_hi + 1; p++) {
        const double *pl = in + (p + 2) * g.plane_pitch;
        double s = w.b[0] * __ldg(pl - 2 * g.row_pitch);
        s = fma(w.b[1], __ldg(pl - g.row_pitch), s);
        s = fma(w.b[2], __ldg(pl), s);
        s = fma(w.b[3], __ldg(pl + g.row_pitch), s);
        s = fma(w.b[4], __ldg(pl + 2 * g.row_pitch), s);
        double t = w.c[2] * s;
        t = fma(w.c[0], __shfl_up_sync(kFull, s, 2), t);    // column c - 2
        t = fma(w.c[1], __shfl_up_sync(kFull, s, 1), t);    // column c - 1
        t = fma(w.c[3], __shfl_down_sync(kFull, s, 1), t);  // column c + 1
        t = fma(w.c[4], __shfl_down_sync(kFull, s, 2), t);  // column c + 2
#pragma unroll
        for (int k = 0; k < 5; k++) acc[k] = fma(w.a[4 - k], t, acc[k]);  // dh = 2 - k
        if (owner && p - 2 >= q_lo) out[p * g.plane_pitch] = acc[0];
        acc[0] = acc[1];
        acc[1] = acc[2];
        acc[2] = acc[3];
        acc[3] = acc[4];
        acc[4] = 0.0;
    }
}

// which kernel for the other forms: 0 = one cell per thread, plain loads (the first version); 1 = read-only loads, four
// planes per trip; 2 = two cells per thread (needs an even column count and 16-byte aligned buffers, else 0).  Measured
// at 512^3 (profiles/r2_extensions_variants.json), GStencil/s for variants 0 / 1 / 2: 13-point 206 / 216 / 221, rank 1
// along the plane axis 126 / 86 / 178, 125 taps 29 / 29 / 55.  LORA_R2_VARIANT overrides the default (the parity tests
// run every variant).
constexpr int kDefaultVariant = 2;

template <int FORM>
cudaError_t launch_form(int variant, dim3 grid, dim3 block, const Geom3DR2 &g, const WeightsR2 &w, cudaStream_t s) {
    if (variant == 2)
        k_stencil3d_r2_pair<FORM><<<grid, block, 0, s>>>(g, w);
    else if (variant == 1)
        k_stencil3d_r2<FORM, 4><<<grid, block, 0, s>>>(g, w);
    else
        k_stencil3d_r2<FORM, 1><<<grid, block, 0, s>>>(g, w);
    return cudaGetLastError();
}

}  // namespace

// Planes per chunk: every chunk of planes re-reads 4 planes of warm-up, so chunks as long as possible -- but enough of
// them that the grid is many waves deep (about 32 CTAs of 256 threads per SM: 5 to 8 are resident, and with one chunk a
// 512^3 grid is 1.4 waves, a third of the GPU idle in the second), never more than ceil(planes / 16) of them, nor more than
// the z extent of a grid allows.  ctas_per_plane = CTAs one plane of one chunk takes.
long long r2_planes_per_chunk(long long planes, long long ctas_per_plane, int sm_count) {
    long long want = (32LL * sm_count + ctas_per_plane - 1) / ctas_per_plane;  // chunks wanted
    want = want < 1 ? 1 : want;
    if (want > (planes + 15) / 16) want = (planes + 15) / 16;
    long long L = (planes + want - 1) / want;
    if (L < (planes + 65534) / 65535) L = (planes + 65534) / 65535;
    return L > 0x7fffffffLL ? 0x7fffffffLL : L;
}

// columns one CTA stores: 128 (one cell per thread), 256 (two cells), 112 (SEP5: 4 warps of 28)
int r2_cols_per_cta(int form, int variant) {
    return form == LORA_FORM_SEP5 ? (kR2Cols / 32) * kSepOut : kR2Cols * (variant == 2 ? 2 : 1);
}
int r2_rows_per_cta() { return kR2Rows; }

cudaError_t launch_3d_r2(int form, Geom3DR2 g, const WeightsR2 &w, int sm_count, cudaStream_t s) {
    const long long planes = g.hi - g.lo;
    if (planes <= 0) return cudaSuccess;
    int variant = kDefaultVariant;
    if (const char *e = getenv("LORA_R2_VARIANT")) {
        const int v = atoi(e);
        if (v >= 0 && v <= 2) variant = v;
    }
    const bool pair_ok = g.n % 2 == 0 && reinterpret_cast<uintptr_t>(g.in) % 16 == 0 && reinterpret_cast<uintptr_t>(g.out) % 16 == 0;
    if (variant == 2 && !pair_ok) variant = 0;
    const int cols_per_cta = r2_cols_per_cta(form, variant);
    const long long bx = (g.n + cols_per_cta - 1) / cols_per_cta, by = (g.m + kR2Rows - 1) / kR2Rows;
    if (by > 65535) return cudaErrorInvalidConfiguration;
    g.planes_per_chunk = (int)r2_planes_per_chunk(planes, bx * by, sm_count);
    const long long chunks = (planes + g.planes_per_chunk - 1) / g.planes_per_chunk;
    const dim3 grid((unsigned)bx, (unsigned)by, (unsigned)chunks);
    const dim3 block(kR2Cols, kR2Rows);
    switch (form) {
        case LORA_FORM_SEP5:
            k_stencil3d_r2_sep<<<grid, block, 0, s>>>(g, w);
            return cudaGetLastError();
        case LORA_FORM_STAR13: return launch_form<LORA_FORM_STAR13>(variant, grid, block, g, w, s);
        case LORA_FORM_HSEP5: return launch_form<LORA_FORM_HSEP5>(variant, grid, block, g, w, s);
        case LORA_FORM_DIRECT125: return launch_form<LORA_FORM_DIRECT125>(variant, grid, block, g, w, s);
        default: return cudaErrorInvalidValue;
    }
}

}  // namespace lora
