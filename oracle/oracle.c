/*
 * oracle.c -- CPU restatement of the LoRAStencil hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library; the product (lorastencil_b200/) never links or calls it.
 *
 * What is restated (paths relative to /root/reference):
 *   - the direct-tap single-step CPU stencils `test_cpu`
 *       1-D  src/1d/main.cu:34-40      (9 taps, radius 4)
 *       2-D  src/2d/main.cu:38-93      (49 taps, 7x7 window)
 *       3-D  src/3d/main.cu:33-68      (27 taps, 3x3x3 window)
 *   - the buffer semantics of the GPU host operators (S1-S4 in SURVEY.md section 8a)
 *       ping-pong A<-input, B<-zeros, launch i reads buf[i%2], writes the INTERIOR of
 *       buf[(i+1)%2], result = whole padded buf[times%2]
 *       src/2d/gpu.cu:392-421, src/1d/gpu_1r.cu:103-134, src/3d/gpu_box.cu:190-223
 *       (1-D copies back only cols-1 doubles: src/1d/gpu_1r.cu:134)
 *   - the input fill: unseeded glibc rand() % 100 (2-D/3-D) or % 10000 (1-D) over the
 *       whole padded array  src/2d/main.cu:229-236, src/3d/main.cu:164-168,
 *       src/1d/main.cu:105-109
 *
 * Parity pinning: tests/test_oracle.py checks these functions against
 *   (a) the golden single-step table of SURVEY.md section 8(c) (tests/golden/), and
 *   (b) the reference's own `test_cpu`, compiled unmodified into oracle/_ref/ by
 *       oracle/Makefile (only where /root/reference exists or oracle/_ref was shipped).
 * Multi-step semantics are not pinned by any reference test; they are additionally
 * cross-checked on the GPU box against the reference GPU operators compiled for sm_100a
 * (oracle/_ref/libref_gpu_*.so, tests/test_parity_refgpu.py).
 *
 * Arithmetic: FP64, taps summed left to right in the reference's order, no FMA
 * contraction (-ffp-contract=off), so it is bit-identical to `test_cpu` built with
 * g++ on x86-64.
 */
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

typedef long long i64;

/* ---- fill: src/2d/main.cu:229-236 (mod 100), src/1d/main.cu:105-109 (mod 10000) ---- */
void oracle_fill_rand(double *buf, i64 count, int mod) {
    srand(1); /* an unseeded glibc rand() behaves as srand(1) */
    for (i64 i = 0; i < count; i++) buf[i] = (double)(rand() % mod);
}

int oracle_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

void oracle_set_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

/* ---- single steps: restatement of test_cpu ---- */

/* src/1d/main.cu:34-40 -- cols is the PADDED length, interior = [4, cols-4) */
void oracle_step_1d(const double *in, double *out, const double *param, i64 cols) {
#pragma omp parallel for schedule(static)
    for (i64 c = 4; c < cols - 4; c++) {
        double acc = param[0] * in[c - 4];
        for (int k = 1; k < 9; k++) acc = acc + param[k] * in[c - 4 + k];
        out[c] = acc;
    }
}

/* src/2d/main.cu:38-93 -- rows/cols PADDED, interior = [4,rows-4) x [4,cols-4),
 * always the full 7x7 window whatever the shape */
void oracle_step_2d(const double *in, double *out, const double *param, i64 rows, i64 cols) {
#pragma omp parallel for schedule(static)
    for (i64 r = 4; r < rows - 4; r++) {
        for (i64 c = 4; c < cols - 4; c++) {
            double acc = 0.0;
            int first = 1;
            for (int dr = -3; dr <= 3; dr++) {
                const double *row = in + (r + dr) * cols + c;
                const double *w = param + (dr + 3) * 7 + 3;
                for (int dc = -3; dc <= 3; dc++) {
                    double p = w[dc] * row[dc];
                    if (first) { acc = p; first = 0; } else acc = acc + p;
                }
            }
            out[r * cols + c] = acc;
        }
    }
}

/* src/3d/main.cu:33-68 -- heights/rows/cols PADDED with halos 1/2/4 */
void oracle_step_3d(const double *in, double *out, const double *param, i64 heights, i64 rows,
                    i64 cols) {
#pragma omp parallel for schedule(static)
    for (i64 h = 1; h < heights - 1; h++) {
        for (i64 r = 2; r < rows - 2; r++) {
            for (i64 c = 4; c < cols - 4; c++) {
                double acc = 0.0;
                int first = 1;
                for (int dh = -1; dh <= 1; dh++)
                    for (int dr = -1; dr <= 1; dr++)
                        for (int dc = -1; dc <= 1; dc++) {
                            double p = param[(dh + 1) * 9 + (dr + 1) * 3 + (dc + 1)] *
                                       in[((h + dh) * rows + (r + dr)) * cols + (c + dc)];
                            if (first) { acc = p; first = 0; } else acc = acc + p;
                        }
                out[(h * rows + r) * cols + c] = acc;
            }
        }
    }
}

/* ---- multi-step operators: S2/S3 semantics of gpu_* ----
 * `out` receives the whole padded buf[times%2] (halo = input halo for even times, zero
 * for odd times).  Returns 0, or -1 when the two work buffers cannot be allocated. */

static int run_generic(const double *in, double *out, i64 total, i64 copy_back, int times,
                       void (*step)(const double *, double *, const double *, const i64 *),
                       const double *param, const i64 *dims) {
    double *buf[2];
    buf[0] = (double *)malloc((size_t)total * sizeof(double));
    buf[1] = (double *)calloc((size_t)total, sizeof(double));
    if (!buf[0] || !buf[1]) { free(buf[0]); free(buf[1]); return -1; }
    memcpy(buf[0], in, (size_t)total * sizeof(double));
    int i = 0;
    for (; i < times; i++) step(buf[i % 2], buf[(i + 1) % 2], param, dims);
    memcpy(out, buf[i % 2], (size_t)copy_back * sizeof(double));
    free(buf[0]);
    free(buf[1]);
    return 0;
}

static void step1(const double *a, double *b, const double *p, const i64 *d) {
    oracle_step_1d(a, b, p, d[0]);
}
static void step2(const double *a, double *b, const double *p, const i64 *d) {
    oracle_step_2d(a, b, p, d[0], d[1]);
}
static void step3(const double *a, double *b, const double *p, const i64 *d) {
    oracle_step_3d(a, b, p, d[0], d[1], d[2]);
}

/* src/1d/gpu_1r.cu:103-134 -- note the D2H of cols-1 doubles: out[cols-1] is untouched */
int oracle_run_1d(const double *in, double *out, const double *param, int times, i64 n) {
    i64 dims[1] = {n + 8};
    return run_generic(in, out, dims[0], dims[0] - 1, times, step1, param, dims);
}

/* src/2d/gpu.cu:392-421 */
int oracle_run_2d(const double *in, double *out, const double *param, int times, i64 m, i64 n) {
    i64 dims[2] = {m + 8, n + 8};
    return run_generic(in, out, dims[0] * dims[1], dims[0] * dims[1], times, step2, param, dims);
}

/* src/3d/gpu_box.cu:190-223 */
int oracle_run_3d(const double *in, double *out, const double *param, int times, i64 h, i64 m,
                  i64 n) {
    i64 dims[3] = {h + 2, m + 4, n + 8};
    i64 total = dims[0] * dims[1] * dims[2];
    return run_generic(in, out, total, total, times, step3, param, dims);
}

/* ---- radius-2 3-D shapes (box3d2r / star3d2r) -- NOT in the reference (src/3d/3d_utils.h:39-42 stops at radius 1;
 * SURVEY.md section 8(f)-4).  The same direct-tap protocol as test_cpu (src/3d/main.cu:33-68) widened to a
 * 5 x 5 x 5 window, on the product's layout for these shapes: (h+4) x (m+4) x (n+8), interior origin [2][2][4];
 * 125 weights [(dh+2)*25 + (dr+2)*5 + dc+2], summed left to right.  Parity of this function is pinned by
 * tests/test_r2.py against scipy.ndimage.correlate, not by any reference vector ("unpinned by the reference"). */
void oracle_step_3d_r2(const double *in, double *out, const double *param, i64 heights, i64 rows, i64 cols) {
#pragma omp parallel for schedule(static)
    for (i64 h = 2; h < heights - 2; h++) {
        for (i64 r = 2; r < rows - 2; r++) {
            for (i64 c = 4; c < cols - 4; c++) {
                double acc = 0.0;
                int first = 1;
                for (int dh = -2; dh <= 2; dh++)
                    for (int dr = -2; dr <= 2; dr++)
                        for (int dc = -2; dc <= 2; dc++) {
                            double p = param[(dh + 2) * 25 + (dr + 2) * 5 + (dc + 2)] *
                                       in[((h + dh) * rows + (r + dr)) * cols + (c + dc)];
                            if (first) { acc = p; first = 0; } else acc = acc + p;
                        }
                out[(h * rows + r) * cols + c] = acc;
            }
        }
    }
}

static void step3r2(const double *a, double *b, const double *p, const i64 *d) {
    oracle_step_3d_r2(a, b, p, d[0], d[1], d[2]);
}

/* the S2 / S3 buffer semantics of the gpu_* operators (src/3d/gpu_box.cu:190-223) on that layout */
int oracle_run_3d_r2(const double *in, double *out, const double *param, int times, i64 h, i64 m, i64 n) {
    i64 dims[3] = {h + 4, m + 4, n + 8};
    i64 total = dims[0] * dims[1] * dims[2];
    return run_generic(in, out, total, total, times, step3r2, param, dims);
}
