// decompose.cpp -- see decompose.h.  Pure host C++ (no CUDA).
#include "decompose.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>

namespace lora {

int shape_dim(int s) {
    switch (s) {
        case LORA_1D1R: case LORA_1D2R: return 1;
        case LORA_STAR2D1R: case LORA_BOX2D1R: case LORA_STAR2D3R: case LORA_BOX2D3R: return 2;
        case LORA_BOX3D1R: case LORA_STAR3D1R: return 3;
        default: return 0;
    }
}

bool shape_is_r2(int s) { return s == LORA_BOX3D2R || s == LORA_STAR3D2R; }

int shape_nparams(int s) {
    if (shape_is_r2(s)) return 125;
    switch (shape_dim(s)) { case 1: return 9; case 2: return 49; case 3: return 27; default: return 0; }
}

const char *shape_cli_name(int s) {
    static const char *n[] = {"1d1r", "1d2r", "star2d1r", "box2d1r", "star2d3r", "box2d3r", "box3d1r", "star3d1r",
                              "box3d2r", "star3d2r"};
    return (s >= 0 && s < LORA_NUM_SHAPES_EXT) ? n[s] : "?";
}

// banners of the reference operators: src/1d/gpu_1r.cu:127, src/1d/gpu_2r.cu:129,
// src/2d/gpu.cu:415,474,549, src/3d/gpu_box.cu:216, src/3d/gpu_star.cu:185.
// box2d1r runs gpu_box_2d3r and therefore prints "box_2d3r" (src/2d/main.cu:276-279).
const char *shape_banner(int s) {
    static const char *n[] = {"1D 1d1r", "1D 1d2r", "2D star_2d1r", "2D box_2d3r", "2D star_2d3r", "2D box_2d3r",
                              "3D box_3d1r", "3D star_3d1r", "3D box_3d2r", "3D star_3d2r"};
    return (s >= 0 && s < LORA_NUM_SHAPES_EXT) ? n[s] : "?";
}

// src/1d/gpu_1r.cu:132 (x3), src/1d/gpu_2r.cu:134 (x2), src/2d/gpu.cu:419 (x3), :478 (x1), :553 (x3),
// src/3d/gpu_box.cu:221 (x1), src/3d/gpu_star.cu:190 (x1)
int shape_artifact_k(int s) {
    static const int k[] = {3, 2, 3, 3, 1, 3, 1, 1};
    return (s >= 0 && s < LORA_NUM_SHAPES) ? k[s] : 1;
}

void reference_table(int shape, double *out) {
    switch (shape) {
        case LORA_1D1R: {  // src/1d/main.cu:77
            const double w[9] = {0, 1, 2, 3, 4, 3, 2, 1, 0};
            std::memcpy(out, w, sizeof w);
            return;
        }
        case LORA_1D2R: {  // src/1d/main.cu:78
            const double w[9] = {1, 2, 3, 4, 5, 4, 3, 2, 1};
            std::memcpy(out, w, sizeof w);
            return;
        }
        case LORA_BOX2D1R:
        case LORA_BOX2D3R: {  // src/2d/main.cu:150-174: ring-wise numbering 1..10, 8-fold symmetric, centre 8
            std::fill(out, out + 49, 0.0);
            int num = 1;
            for (int i = -3; i <= 0; i++)
                for (int j = i; j <= 0; j++) {
                    const int a[2] = {i, -i}, b[2] = {j, -j};
                    for (int x = 0; x < 2; x++)
                        for (int y = 0; y < 2; y++) {
                            out[(a[x] + 3) * 7 + (b[y] + 3)] = num;
                            out[(b[y] + 3) * 7 + (a[x] + 3)] = num;
                        }
                    num++;
                }
            out[3 * 7 + 3] = 8;
            return;
        }
        case LORA_STAR2D3R: {  // src/2d/main.cu:176-184: arms 1,2,3 towards the centre 4
            std::fill(out, out + 49, 0.0);
            for (int d = 0; d <= 3; d++) {
                const double v = 4 - d;
                out[(3 - d) * 7 + 3] = out[(3 + d) * 7 + 3] = v;
                out[3 * 7 + 3 - d] = out[3 * 7 + 3 + d] = v;
            }
            return;
        }
        case LORA_STAR2D1R: {  // src/2d/main.cu:186-195: |dr|+|dc| <= 3 diamond, 16 / 2^(|dr|+|dc|), tips 1
            for (int dr = -3; dr <= 3; dr++)
                for (int dc = -3; dc <= 3; dc++) {
                    const int d = std::abs(dr) + std::abs(dc);
                    double v = 0.0;
                    if (d <= 2) v = 16 >> d;
                    else if (d == 3) v = (dr == 0 || dc == 0) ? 1.0 : 2.0;
                    out[(dr + 3) * 7 + dc + 3] = v;
                }
            return;
        }
        case LORA_BOX3D1R:  // src/3d/main.cu:112-119
            for (int i = 0; i < 27; i++) out[i] = (i % 3 == 1) ? 2.0 : 1.0;
            return;
        case LORA_STAR3D1R: {  // src/3d/main.cu:121-125
            std::fill(out, out + 27, 0.0);
            out[13] = 2;
            out[4] = out[22] = out[10] = out[16] = out[12] = out[14] = 1;
            return;
        }
        // radius-2 shapes: no reference table exists; small integers like the reference's, so that the first launches
        // are exact in FP64
        case LORA_BOX3D2R: {
            const double a[5] = {1, 2, 3, 2, 1};
            for (int h = 0; h < 5; h++)
                for (int r = 0; r < 5; r++)
                    for (int c = 0; c < 5; c++) out[h * 25 + r * 5 + c] = a[h] * a[r] * a[c];
            return;
        }
        case LORA_STAR3D2R: {
            std::fill(out, out + 125, 0.0);
            out[62] = 3;
            for (int d = 1; d <= 2; d++) {
                const double v = 3 - d;
                out[62 - d] = out[62 + d] = out[62 - 5 * d] = out[62 + 5 * d] = out[62 - 25 * d] = out[62 + 25 * d] = v;
            }
            return;
        }
        default:
            return;
    }
}

// ---------------------------------------------------------------------------------------------
// 1-D
// ---------------------------------------------------------------------------------------------
bool decompose_1d(int shape, int /*mode*/, const double *params, Decomp1D &d) {
    if (shape_dim(shape) != 1) return false;
    // both modes: the band matrix P[r+c][c] = params[r] is the 9 taps themselves (src/1d/gpu_1r.cu:95-99)
    for (int k = 0; k < 9; k++) d.w[k] = params[k];
    return true;
}

// ---------------------------------------------------------------------------------------------
// 2-D
// ---------------------------------------------------------------------------------------------
static inline double &at(double *m, int r, int c) { return m[r * 7 + c]; }
static inline double at(const double *m, int r, int c) { return m[r * 7 + c]; }

static double max_abs(const double *m, int n) {
    double s = 0;
    for (int i = 0; i < n; i++) s = std::max(s, std::fabs(m[i]));
    return s;
}

static void rebuild_effective(Decomp2D &d) {
    std::fill(d.effective, d.effective + 49, 0.0);
    if (d.form == LORA_FORM_DIRECT49) {
        std::memcpy(d.effective, d.direct, sizeof d.direct);
        return;
    }
    for (int t = 0; t < d.nterms; t++)
        for (int r = 0; r < 7; r++)
            for (int c = 0; c < 7; c++) at(d.effective, r, c) += d.vert[t][r] * d.horiz[t][c];
    at(d.effective, 3, 3) += d.centre;
    if (d.form == LORA_FORM_DIAMOND) {
        static const int pos[8][2] = {{0, -3}, {0, 3}, {-3, 0}, {3, 0}, {-2, -2}, {-2, 2}, {2, -2}, {2, 2}};
        for (int k = 0; k < 8; k++) at(d.effective, 3 + pos[k][0], 3 + pos[k][1]) += d.residual[k];
    }
}

static void finish(Decomp2D &d, const double *target) {
    rebuild_effective(d);
    d.recon_err = 0;
    if (target)
        for (int i = 0; i < 49; i++) d.recon_err = std::max(d.recon_err, std::fabs(d.effective[i] - target[i]));
}

// The reference's peel, restated (src/2d/gpu.cu:280-350): every level takes the first ROW of the
// current residual as u_t and (first column)/(pivot) as v_t, mirrors rows +-i, and the kernel then
// applies u_t along rows (vertically) and v_t along columns (src/2d/gpu.cu:358-369, :76-99).
// The 1x1 remainder is computed and never applied.
static void peel_reference(const double *P, Decomp2D &d) {
    double F[4][49] = {}, T[2][49] = {};
    for (int c = 0; c < 7; c++) { at(F[0], 0, c) = at(P, 0, c); at(F[0], 6, c) = at(P, 6, c); }
    for (int r = 1; r <= 3; r++) {
        const double prop = at(P, r, 0) / at(P, 0, 0);
        for (int c = 0; c < 7; c++) {
            at(F[0], r, c) = prop * at(P, 0, c);
            at(F[0], 6 - r, c) = at(F[0], r, c);
            at(T[0], r, c) = at(P, r, c) - at(F[0], r, c);
            at(T[0], 6 - r, c) = at(T[0], r, c);
        }
    }
    for (int c = 1; c <= 5; c++) { at(F[1], 1, c) = at(T[0], 1, c); at(F[1], 5, c) = at(F[1], 1, c); }
    for (int r = 2; r <= 3; r++) {
        const double prop = at(T[0], r, 1) / at(T[0], 1, 1);
        for (int c = 1; c <= 5; c++) {
            at(F[1], r, c) = prop * at(T[0], 1, c);
            at(F[1], 6 - r, c) = at(F[1], r, c);
            at(T[1], r, c) = at(T[0], r, c) - at(F[1], r, c);
            at(T[1], 6 - r, c) = at(T[1], r, c);
        }
    }
    for (int c = 2; c <= 4; c++) { at(F[2], 2, c) = at(T[1], 2, c); at(F[2], 4, c) = at(T[1], 2, c); }
    {
        const double prop = at(T[1], 3, 2) / at(T[1], 2, 2);
        for (int c = 2; c <= 4; c++) {
            at(F[2], 3, c) = prop * at(T[1], 2, c);
            at(F[3], 3, c) = at(T[1], 3, c) - at(F[2], 3, c);
        }
    }
    d.form = LORA_FORM_PYRAMID;
    d.nterms = 3;
    std::memset(d.vert, 0, sizeof d.vert);
    std::memset(d.horiz, 0, sizeof d.horiz);
    for (int t = 0; t < 3; t++)
        for (int i = t; i <= 6 - t; i++) {
            d.vert[t][i] = at(F[t], t, i);                     // u_t: a row, applied vertically
            d.horiz[t][i] = at(F[t], i, t) / at(F[t], t, t);   // v_t: column / pivot, applied horizontally
        }
    d.centre = 0.0;  // fact_param_matrix_h[3][3*7+3] is dropped (src/2d/gpu.cu:349-358)
    d.macs = 7 + 7 + 5 + 5 + 3 + 3;
    char buf[160];
    std::snprintf(buf, sizeof buf, "2d pyramid rank-3 (7/5/3), reference peel, dropped 1x1 remainder %.3g",
                  at(F[3], 3, 3));
    d.desc = buf;
}

// General pyramidal peel: level t fits the rank-1 term (column t / pivot) (x) (row t) to the outer
// ring of the residual's (7-2t)^2 core and requires the whole ring to vanish.
static bool peel_general(const double *P, Decomp2D &d, double tol) {
    double R[49];
    std::memcpy(R, P, sizeof R);
    std::memset(d.vert, 0, sizeof d.vert);
    std::memset(d.horiz, 0, sizeof d.horiz);
    for (int t = 0; t < 3; t++) {
        const int lo = t, hi = 6 - t;
        const double piv = at(R, lo, lo);
        bool ring_zero = true;
        for (int k = lo; k <= hi; k++)
            if (std::fabs(at(R, lo, k)) > tol || std::fabs(at(R, hi, k)) > tol || std::fabs(at(R, k, lo)) > tol ||
                std::fabs(at(R, k, hi)) > tol)
                ring_zero = false;
        if (ring_zero) continue;  // nothing to peel at this level
        if (std::fabs(piv) <= tol) return false;
        for (int k = lo; k <= hi; k++) {
            d.vert[t][k] = at(R, k, lo) / piv;
            d.horiz[t][k] = at(R, lo, k);
        }
        for (int r = lo; r <= hi; r++)
            for (int c = lo; c <= hi; c++) at(R, r, c) -= d.vert[t][r] * d.horiz[t][c];
        for (int k = lo; k <= hi; k++) {
            if (std::fabs(at(R, lo, k)) > tol || std::fabs(at(R, hi, k)) > tol || std::fabs(at(R, k, lo)) > tol ||
                std::fabs(at(R, k, hi)) > tol)
                return false;  // ring does not vanish: not pyramidal
            at(R, lo, k) = at(R, hi, k) = at(R, k, lo) = at(R, k, hi) = 0.0;
        }
    }
    d.form = LORA_FORM_PYRAMID;
    d.nterms = 3;
    d.centre = at(R, 3, 3);
    d.macs = 7 + 7 + 5 + 5 + 3 + 3 + 1;
    char buf[128];
    std::snprintf(buf, sizeof buf, "2d pyramid rank-3 (7/5/3) + centre %.6g", d.centre);
    d.desc = buf;
    return true;
}

// Rank-r fallback for tables the pyramidal peel rejects (the reference's peel, src/2d/gpu.cu:280-350, assumes rows
// +-i equal and pivots on the diagonal; it has no fallback at all): LU with full pivoting, i.e. cross approximation --
// pick the largest entry of the residual, subtract (its column / pivot) (x) (its row), at most three times.  A table of
// rank r <= 3 leaves a residual of rounding size; the terms have full support 7, so the form costs 14 r taps (28 / 42
// against 49 direct).  Rank 1 never gets here: the pyramidal peel takes it.
static bool lowrank_lu(const double *P, Decomp2D &d, double tol) {
    double R[49];
    std::memcpy(R, P, sizeof R);
    std::memset(d.vert, 0, sizeof d.vert);
    std::memset(d.horiz, 0, sizeof d.horiz);
    int nt = 0;
    for (; nt < 3; nt++) {
        int pi = 0, pj = 0;
        double best = 0;
        for (int r = 0; r < 7; r++)
            for (int c = 0; c < 7; c++)
                if (std::fabs(at(R, r, c)) > best) best = std::fabs(at(R, r, c)), pi = r, pj = c;
        if (best <= tol) break;
        const double piv = at(R, pi, pj);
        for (int k = 0; k < 7; k++) {
            d.vert[nt][k] = at(R, k, pj) / piv;
            d.horiz[nt][k] = at(R, pi, k);
        }
        for (int r = 0; r < 7; r++)
            for (int c = 0; c < 7; c++) at(R, r, c) -= d.vert[nt][r] * d.horiz[nt][c];
        for (int k = 0; k < 7; k++) at(R, k, pj) = at(R, pi, k) = 0.0;  // exact zeros by construction
    }
    if (nt < 2 || max_abs(R, 49) > tol) return false;
    d.form = nt == 2 ? LORA_FORM_RANK2 : LORA_FORM_RANK3;
    d.nterms = nt;
    d.centre = 0.0;
    d.macs = 14 * nt;
    char buf[128];
    std::snprintf(buf, sizeof buf, "2d rank-%d (LU with full pivoting, %d terms of support 7)", nt, nt);
    d.desc = buf;
    return true;
}

static bool is_cross(const double *P, double tol) {
    for (int r = 0; r < 7; r++)
        for (int c = 0; c < 7; c++)
            if (r != 3 && c != 3 && std::fabs(at(P, r, c)) > tol) return false;
    return true;
}

static void make_cross(const double *P, Decomp2D &d) {
    d.form = LORA_FORM_CROSS;
    d.nterms = 2;
    std::memset(d.vert, 0, sizeof d.vert);
    std::memset(d.horiz, 0, sizeof d.horiz);
    // term 0: column arm including the centre (src/2d/gpu.cu:433-437)
    for (int r = 0; r < 7; r++) d.vert[0][r] = at(P, r, 3);
    d.horiz[0][3] = 1.0;
    // term 1: row arm without the centre (src/2d/gpu.cu:438-444)
    d.vert[1][3] = 1.0;
    for (int c = 0; c < 7; c++) d.horiz[1][c] = (c == 3) ? 0.0 : at(P, 3, c);
    d.centre = 0.0;
    d.macs = 7 + 6;
    d.desc = "2d cross radius 3: column arm (7 taps) + row arm (6 taps)";
}

// one rank-1 term of support 5 fitted to the centre cross + a residual confined to the 8
// positions the reference's star2d1r kernel patches on the CUDA cores (src/2d/gpu.cu:249-264)
static bool try_diamond(const double *P, Decomp2D &d, double tol) {
    const double piv = at(P, 3, 3);
    if (std::fabs(piv) <= tol) return false;
    double a[7] = {}, b[7] = {};
    for (int k = 1; k <= 5; k++) { a[k] = at(P, k, 3) / piv; b[k] = at(P, 3, k); }
    static const int pos[8][2] = {{0, -3}, {0, 3}, {-3, 0}, {3, 0}, {-2, -2}, {-2, 2}, {2, -2}, {2, 2}};
    double res[8] = {};
    for (int r = 0; r < 7; r++)
        for (int c = 0; c < 7; c++) {
            const double e = at(P, r, c) - a[r] * b[c];
            int k = -1;
            for (int q = 0; q < 8; q++)
                if (r - 3 == pos[q][0] && c - 3 == pos[q][1]) k = q;
            if (k >= 0) res[k] = e;
            else if (std::fabs(e) > tol) return false;
        }
    d.form = LORA_FORM_DIAMOND;
    d.nterms = 1;
    std::memset(d.vert, 0, sizeof d.vert);
    std::memset(d.horiz, 0, sizeof d.horiz);
    std::memcpy(d.vert[0], a, sizeof a);
    std::memcpy(d.horiz[0], b, sizeof b);
    std::memcpy(d.residual, res, sizeof res);
    d.centre = 0.0;
    d.macs = 5 + 5 + 8;
    d.desc = "2d diamond: rank-1 (5x5) + 8 residual taps";
    return true;
}

// A pyramid whose middle (support-5) term vanishes at offsets +-1 in both profiles and that has no centre remainder
// runs the kernel variant compiled without those taps: 26 instead of 31 FP64 operations per cell.  The reference's
// box table peels into [1,2,3,4,3,2,1], [0,1,0,-1,0,1,0], [0,0,-1,-3,-1,0,0] -- exactly this pattern.
static void prune_pyramid(Decomp2D &d) {
    if (d.form != LORA_FORM_PYRAMID || d.centre != 0.0) return;
    if (d.vert[1][2] != 0.0 || d.vert[1][4] != 0.0 || d.horiz[1][2] != 0.0 || d.horiz[1][4] != 0.0) return;
    d.form = LORA_FORM_PYRAMID_PRUNED;
    d.macs = 7 + 7 + 3 + 3 + 3 + 3;
    d.desc += "; zero taps of the middle term pruned";
}

bool decompose_2d(int shape, int mode, const double *params, Decomp2D &d) {
    if (shape_dim(shape) != 2) return false;
    d = Decomp2D();
    if (mode == LORA_WEIGHTS_REFERENCE) {
        switch (shape) {
            case LORA_BOX2D1R:
            case LORA_BOX2D3R:
                peel_reference(params, d);
                finish(d, nullptr);
                prune_pyramid(d);
                return true;
            case LORA_STAR2D3R: {
                // only column 3 and row 3 of params are read (src/2d/gpu.cu:433-444)
                double cross[49] = {};
                for (int k = 0; k < 7; k++) { at(cross, k, 3) = at(params, k, 3); at(cross, 3, k) = at(params, 3, k); }
                make_cross(cross, d);
                finish(d, nullptr);
                return true;
            }
            case LORA_STAR2D1R: {
                // params ignored: u = v = {0,1,2,4,2,1,0} (src/2d/gpu.cu:486-487) plus
                // +1 at (0,+-3),(+-3,0) and -1 at (+-2,+-2) (src/2d/gpu.cu:254-262)
                d.form = LORA_FORM_DIAMOND;
                d.nterms = 1;
                const double uv[7] = {0, 1, 2, 4, 2, 1, 0};
                std::memcpy(d.vert[0], uv, sizeof uv);
                std::memcpy(d.horiz[0], uv, sizeof uv);
                for (int k = 0; k < 4; k++) d.residual[k] = 1.0;
                for (int k = 4; k < 8; k++) d.residual[k] = -1.0;
                d.macs = 5 + 5 + 8;
                d.desc = "2d diamond: fixed rank-1 {1,2,4,2,1}^2 + 8 residual taps (reference ignores params)";
                finish(d, nullptr);
                return true;
            }
            default:
                return false;
        }
    }
    // GENERAL: cheapest exact form first
    const double scale = std::max(max_abs(params, 49), 1e-300);
    const double tol = 64 * 2.220446049250313e-16 * scale;
    if (is_cross(params, 0.0)) {
        make_cross(params, d);
        finish(d, params);
        return true;
    }
    if (try_diamond(params, d, tol)) {
        finish(d, params);
        if (d.recon_err <= tol) return true;
    }
    d = Decomp2D();
    if (peel_general(params, d, tol)) {
        finish(d, params);
        if (d.recon_err <= tol) {
            prune_pyramid(d);
            return true;
        }
    }
    d = Decomp2D();
    if (lowrank_lu(params, d, tol)) {
        finish(d, params);
        if (d.recon_err <= tol) return true;
    }
    d = Decomp2D();
    d.form = LORA_FORM_DIRECT49;
    std::memcpy(d.direct, params, sizeof d.direct);
    d.macs = 49;
    d.desc = "2d direct 49 taps (table is not cross / diamond / pyramidal and has rank > 3)";
    finish(d, params);
    return true;
}

// ---------------------------------------------------------------------------------------------
// 3-D
// ---------------------------------------------------------------------------------------------
static void rebuild_effective3(Decomp3D &d) {
    std::fill(d.effective, d.effective + 27, 0.0);
    if (d.form == LORA_FORM_SEP3) {
        for (int h = 0; h < 3; h++)
            for (int r = 0; r < 3; r++)
                for (int c = 0; c < 3; c++) d.effective[h * 9 + r * 3 + c] = d.a[h] * d.b[r] * d.c[c];
    } else if (d.form == LORA_FORM_STAR7) {
        d.effective[13] = d.star[0];
        d.effective[12] = d.star[1];
        d.effective[14] = d.star[2];
        d.effective[10] = d.star[3];
        d.effective[16] = d.star[4];
        d.effective[4] = d.star[5];
        d.effective[22] = d.star[6];
    } else {
        std::memcpy(d.effective, d.direct, sizeof d.direct);
    }
}

static void finish3(Decomp3D &d, const double *target) {
    rebuild_effective3(d);
    d.recon_err = 0;
    if (target)
        for (int i = 0; i < 27; i++) d.recon_err = std::max(d.recon_err, std::fabs(d.effective[i] - target[i]));
}

bool decompose_3d(int shape, int mode, const double *params, Decomp3D &d) {
    if (shape_dim(shape) != 3) return false;
    d = Decomp3D();
    if (mode == LORA_WEIGHTS_REFERENCE) {
        if (shape == LORA_BOX3D1R) {
            // ones along h (three identical plane operators summed, src/3d/gpu_box.cu:126-139),
            // ones along m (all-ones band, :151-157), params[0..2] along n (:158-164)
            d.form = LORA_FORM_SEP3;
            for (int k = 0; k < 3; k++) { d.a[k] = 1.0; d.b[k] = 1.0; d.c[k] = params[k]; }
            d.macs = 9;
            d.desc = "3d separable ones(h) x ones(m) x params[0..2](n) (reference reads 3 of 27 weights)";
        } else {
            // params ignored; unit arms, in-plane row-sum + column-sum give the centre weight 2
            // (src/3d/gpu_star.cu:51, :66-80, :142-151)
            d.form = LORA_FORM_STAR7;
            d.star[0] = 2.0;
            for (int k = 1; k < 7; k++) d.star[k] = 1.0;
            d.macs = 7;
            d.desc = "3d 7-point star, unit arms, centre 2 (reference ignores params)";
        }
        finish3(d, nullptr);
        return true;
    }
    const double scale = std::max(max_abs(params, 27), 1e-300);
    const double tol = 64 * 2.220446049250313e-16 * scale;
    // 7-point support?
    bool star = true;
    for (int h = -1; h <= 1; h++)
        for (int r = -1; r <= 1; r++)
            for (int c = -1; c <= 1; c++)
                if (std::abs(h) + std::abs(r) + std::abs(c) > 1 && params[(h + 1) * 9 + (r + 1) * 3 + c + 1] != 0.0)
                    star = false;
    if (star) {
        d.form = LORA_FORM_STAR7;
        d.star[0] = params[13];
        d.star[1] = params[12];
        d.star[2] = params[14];
        d.star[3] = params[10];
        d.star[4] = params[16];
        d.star[5] = params[4];
        d.star[6] = params[22];
        d.macs = 7;
        d.desc = "3d 7-point star";
        finish3(d, params);
        return true;
    }
    // rank-1 a (x) b (x) c through the largest entry
    int best = 0;
    for (int i = 1; i < 27; i++)
        if (std::fabs(params[i]) > std::fabs(params[best])) best = i;
    const int h0 = best / 9, r0 = (best / 3) % 3, c0 = best % 3;
    const double piv = params[best];
    if (std::fabs(piv) > 0) {
        d.form = LORA_FORM_SEP3;
        for (int k = 0; k < 3; k++) {
            d.a[k] = params[k * 9 + r0 * 3 + c0] / piv;
            d.b[k] = params[h0 * 9 + k * 3 + c0] / piv;
            d.c[k] = params[h0 * 9 + r0 * 3 + k];
        }
        d.macs = 9;
        d.desc = "3d separable rank-1 a(h) x b(m) x c(n)";
        finish3(d, params);
        if (d.recon_err <= tol) return true;
    }
    d = Decomp3D();
    d.form = LORA_FORM_DIRECT27;
    std::memcpy(d.direct, params, sizeof d.direct);
    d.macs = 27;
    d.desc = "3d direct 27 taps (table is neither a star nor rank-1)";
    finish3(d, params);
    return true;
}

// ---------------------------------------------------------------------------------------------
// 3-D radius 2 (no reference counterpart: src/3d/3d_utils.h:39-42 stops at radius 1).  Every weight is honoured; the
// form is what the STRUCTURE of the table allows: 13-point star, rank 1 along the plane axis, or all 125 taps.
// ---------------------------------------------------------------------------------------------
// whether a fully separable table takes the SEP5 kernel (k_stencil3d_r2_sep) or stays with HSEP5: on -- 195 against 179
// GStencil/s for the default box table at 512^3 (profiles/r2_extensions_sep5.json); LORA_R2_SEP5=0|1 overrides
constexpr bool kDefaultSep5 = true;

bool decompose_3d_r2(int shape, const double *params, Decomp3DR2 &d) {
    if (!shape_is_r2(shape)) return false;
    d = Decomp3DR2();
    std::memcpy(d.w, params, sizeof d.w);
    bool star = true;
    for (int h = -2; h <= 2 && star; h++)
        for (int r = -2; r <= 2 && star; r++)
            for (int c = -2; c <= 2; c++) {
                const bool on_axis = (h == 0 && r == 0) || (h == 0 && c == 0) || (r == 0 && c == 0);
                if (!on_axis && params[(h + 2) * 25 + (r + 2) * 5 + c + 2] != 0.0) {
                    star = false;
                    break;
                }
            }
    if (star) {
        d.form = LORA_FORM_STAR13;
        d.macs = 13;
        d.desc = "3d radius-2 13-point star";
        return true;
    }
    // rank 1 along the plane axis: w[dh] = a[dh] * Q with Q = the plane through the largest entry, a = the column of
    // ratios through that entry -- accepted only when the product reproduces every weight to a few ulps; the kernel
    // then applies a[dh] * Q, and the effective taps reported are those products
    const double scale = std::max(max_abs(params, 125), 1e-300);
    const double tol = 64 * 2.220446049250313e-16 * scale;
    int best = 0;
    for (int i = 1; i < 125; i++)
        if (std::fabs(params[i]) > std::fabs(params[best])) best = i;
    const int h0 = best / 25, i0 = best % 25;
    if (std::fabs(params[best]) > 0) {
        // a = column / s, Q = plane * s / pivot for any s: try the smallest column entry first (integer tables such as
        // [1,2,3,2,1] (x) Q then factor into integers again and the kernel's sums stay exact), then the pivot, then 1;
        // the first scaling that reproduces every weight EXACTLY wins, otherwise the closest
        const double piv = params[best];
        double smin = std::fabs(piv);
        for (int h = 0; h < 5; h++) {
            const double v = std::fabs(params[h * 25 + i0]);
            if (v > 0 && v < smin) smin = v;
        }
        const double cand[3] = {smin, piv, 1.0};
        double best_err = -1, a[5], q[25];
        for (double sc : cand) {
            double ta[5], tq[25], err = 0;
            for (int h = 0; h < 5; h++) ta[h] = params[h * 25 + i0] / sc;
            for (int i = 0; i < 25; i++) tq[i] = params[h0 * 25 + i] * sc / piv;
            for (int h = 0; h < 5; h++)
                for (int i = 0; i < 25; i++) err = std::max(err, std::fabs(ta[h] * tq[i] - params[h * 25 + i]));
            if (best_err < 0 || err < best_err) {
                best_err = err;
                std::memcpy(a, ta, sizeof a);
                std::memcpy(q, tq, sizeof q);
            }
            if (err == 0) break;
        }
        if (best_err <= tol) {
            d.form = LORA_FORM_HSEP5;
            std::memcpy(d.a, a, sizeof a);
            std::memcpy(d.q, q, sizeof q);
            for (int h = 0; h < 5; h++)
                for (int i = 0; i < 25; i++) d.w[h * 25 + i] = d.a[h] * d.q[i];
            d.recon_err = best_err;
            d.macs = 30;
            d.desc = "3d radius-2 a(h) x Q(m,n): rank 1 along the plane axis, 25 + 5 taps";
            // ... and Q itself rank 1, b (x) c through its largest entry (same scaling rule)?  Then 5 + 5 + 5 taps.
            bool sep5 = kDefaultSep5;
            if (const char *e = getenv("LORA_R2_SEP5")) sep5 = atoi(e) != 0;
            int qb = 0;
            for (int i = 1; i < 25; i++)
                if (std::fabs(q[i]) > std::fabs(q[qb])) qb = i;
            if (sep5 && std::fabs(q[qb]) > 0) {
                const int r0 = qb / 5, c0 = qb % 5;
                double smin2 = std::fabs(q[qb]);
                for (int i = 0; i < 5; i++) {
                    const double v = std::fabs(q[i * 5 + c0]);
                    if (v > 0 && v < smin2) smin2 = v;
                }
                const double cand2[3] = {smin2, q[qb], 1.0};
                double err2 = -1, b[5], c[5];
                for (double sc : cand2) {
                    double tb[5], tc[5], err = 0;
                    for (int i = 0; i < 5; i++) tb[i] = q[i * 5 + c0] / sc;
                    for (int j = 0; j < 5; j++) tc[j] = q[r0 * 5 + j] * sc / q[qb];
                    for (int h = 0; h < 5; h++)
                        for (int i = 0; i < 5; i++)
                            for (int j = 0; j < 5; j++)
                                err = std::max(err, std::fabs(a[h] * (tb[i] * tc[j]) - params[h * 25 + i * 5 + j]));
                    if (err2 < 0 || err < err2) {
                        err2 = err;
                        std::memcpy(b, tb, sizeof b);
                        std::memcpy(c, tc, sizeof c);
                    }
                    if (err == 0) break;
                }
                if (err2 <= tol) {
                    d.form = LORA_FORM_SEP5;
                    std::memcpy(d.b, b, sizeof b);
                    std::memcpy(d.c, c, sizeof c);
                    for (int h = 0; h < 5; h++)
                        for (int i = 0; i < 5; i++)
                            for (int j = 0; j < 5; j++) d.w[h * 25 + i * 5 + j] = d.a[h] * (d.b[i] * d.c[j]);
                    d.recon_err = err2;
                    d.macs = 15;
                    d.desc = "3d radius-2 separable a(h) x b(m) x c(n): rank 1 along every axis, 5 + 5 + 5 taps";
                }
            }
            return true;
        }
    }
    d.form = LORA_FORM_DIRECT125;
    d.macs = 125;
    d.desc = "3d radius-2 direct 125 taps";
    return true;
}

}  // namespace lora
