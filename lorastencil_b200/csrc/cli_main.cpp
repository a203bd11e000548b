// cli_main.cpp -- the three command-line drivers lorastencil_1d / lorastencil_2d / lorastencil_3d
// (compile with -DLORA_CLI_DIM=1|2|3).  Same argv, same stdout/stderr lines and return codes as the
// reference drivers:
//   src/1d/main.cu:43-181   lorastencil_1d {1d1r|1d2r} n times
//   src/2d/main.cu:97-337   lorastencil_2d {star2d1r|box2d1r|star2d3r|box2d3r} m n times
//   src/3d/main.cu:71-253   lorastencil_3d {box3d1r|star3d1r} h m n times
// Extras the reference ignores: a trailing "--check" (or building with -DCHECK_ERROR) runs the
// reference's verification protocol -- one direct-tap CPU step against one GPU launch, every interior
// cell whose absolute difference exceeds 1e-7 is printed, then "Correct!" (src/2d/main.cu:282-328);
// the process additionally returns 2 when a mismatch was found.  A trailing "--gpus k" (or LORA_NGPU=k in the
// environment) cuts the grid into k slabs along its outermost axis, one per GPU, ghost zones exchanged over NVLink.
// A trailing "--weights FILE" replaces the reference's hard-coded table (src/2d/main.cu:139-195 and siblings) by the
// 9 / 49 / 27 whitespace-separated values of FILE (row-major, as the reference lays its tables out) and runs the
// operator in LORA_WEIGHTS_GENERAL mode: EVERY weight is honoured -- which is what --check then verifies -- where the
// reference operators ignore or drop some of theirs.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/lorastencil.h"

#ifndef LORA_CLI_DIM
#error "compile with -DLORA_CLI_DIM=1|2|3"
#endif

namespace {

struct ShapeName {
    const char *cli;   // argv[1]
    const char *info;  // ShapeStr[] of the reference
    int shape;
    bool r2 = false;   // radius-2 3-D extension: layout (h+4) x (m+4) x (n+8), 125 weights (include/lorastencil.h)
};

#if LORA_CLI_DIM == 1
constexpr int kDim = 1;
const ShapeName kShapes[] = {{"1d1r", "1d1r", LORA_1D1R}, {"1d2r", "1d2r", LORA_1D2R}};
const char *kHelp =
    "Program name: lorastencil_1d\n"
    "Usage: lorastencil_1d shape input_size time_size\n"
    "Shape: 1d1r or 1d2r\n";
const int kHalo[3] = {4, 0, 0};
const int kFillMod = 10000;  // src/1d/main.cu:108
const int kNParams = 9;
#elif LORA_CLI_DIM == 2
constexpr int kDim = 2;
const ShapeName kShapes[] = {{"box2d1r", "box_2d1r", LORA_BOX2D1R},
                             {"star2d1r", "star_2d1r", LORA_STAR2D1R},
                             {"star2d3r", "star_2d3r", LORA_STAR2D3R},
                             {"box2d3r", "box_2d3r", LORA_BOX2D3R}};
const char *kHelp =
    "Program name: lorastencil_2d\n"
    "Usage: lorastencil_2d shape input_size_of_first_dimension input_size_of_second_dimension time_size\n"
    "Shape: box2d1r or star2d1r or box2d3r or star2d3r\n";
const int kHalo[3] = {4, 4, 0};
const int kFillMod = 100;  // src/2d/main.cu:235
const int kNParams = 49;
#else
constexpr int kDim = 3;
const ShapeName kShapes[] = {{"box3d1r", "box_3d1r", LORA_BOX3D1R}, {"star3d1r", "star_3d1r", LORA_STAR3D1R},
                             {"box3d2r", "box_3d2r", LORA_BOX3D2R, true}, {"star3d2r", "star_3d2r", LORA_STAR3D2R, true}};
const char *kHelp =
    "Program name: lorastencil_3d\n"
    "Usage: lorastencil_3d shape input_size_of_first_dimension input_size_of_second_dimension "
    "input_size_of_third_dimension time_size\n"
    "Shape: box3d1r or star3d1r\n"
    "Extra shapes (not in the reference): box3d2r or star3d2r\n";
const int kHalo[3] = {1, 2, 4};
const int kFillMod = 100;  // src/3d/main.cu:167
const int kNParams = 27;
#endif

void print_help() { printf("%s\n", kHelp); }

// one direct-tap step over the interior (the protocol of test_cpu, written for any dimension)
void cpu_step(const std::vector<double> &in, std::vector<double> &out, const double *w, const long long *pd,
              const int *halo, int R) {
    (void)halo;
    (void)R;
#if LORA_CLI_DIM == 1
    for (long long c = 4; c < pd[0] - 4; c++) {
        double a = 0;
        for (int k = 0; k < 9; k++) a += w[k] * in[c - 4 + k];
        out[c] = a;
    }
#elif LORA_CLI_DIM == 2
    for (long long r = 4; r < pd[0] - 4; r++)
        for (long long c = 4; c < pd[1] - 4; c++) {
            double a = 0;
            for (int dr = -3; dr <= 3; dr++)
                for (int dc = -3; dc <= 3; dc++) a += w[(dr + 3) * 7 + dc + 3] * in[(r + dr) * pd[1] + c + dc];
            out[r * pd[1] + c] = a;
        }
#else
    const int W = 2 * R + 1;  // 3 (the reference's window, src/3d/main.cu:33-68) or 5
    for (long long h = halo[0]; h < pd[0] - halo[0]; h++)
        for (long long r = halo[1]; r < pd[1] - halo[1]; r++)
            for (long long c = halo[2]; c < pd[2] - halo[2]; c++) {
                double a = 0;
                for (int dh = -R; dh <= R; dh++)
                    for (int dr = -R; dr <= R; dr++)
                        for (int dc = -R; dc <= R; dc++)
                            a += w[((dh + R) * W + dr + R) * W + dc + R] *
                                 in[((h + dh) * pd[1] + r + dr) * pd[2] + c + dc];
                out[(h * pd[1] + r) * pd[2] + c] = a;
            }
#endif
}

}  // namespace

int main(int argc, char *argv[]) {
    if (argc < kDim + 3) {
        print_help();
        return 1;
    }
    const std::string arg1 = argv[1];
    const ShapeName *sn = nullptr;
    for (const auto &s : kShapes)
        if (arg1 == s.cli) sn = &s;
    if (!sn) {
        print_help();
        return 1;
    }

    long long dims[3] = {0, 0, 0};
    int times = 0;
    try {
        for (int i = 0; i < kDim; i++) dims[i] = std::stoi(argv[2 + i]);
        times = std::stoi(argv[2 + kDim]);
    } catch (const std::invalid_argument &) {
        std::cerr << "Invalid argument: cannot convert the parameter(s) to integer.\n";
        return 1;
    } catch (const std::out_of_range &) {
        std::cerr << "Argument out of range: the parameter(s) is(are) too large.\n";
        return 1;
    }
    bool check = false;
#if defined(CHECK_ERROR)
    check = true;
#endif
    const char *weights_path = nullptr;
    int halo[3] = {kHalo[0], kHalo[1], kHalo[2]}, nparams = kNParams, radius = 1;
    if (sn->r2) halo[0] = 2, nparams = 125, radius = 2;
    for (int i = kDim + 3; i < argc; i++) {
        if (std::strcmp(argv[i], "--check") == 0) check = true;
        // trailing "--gpus k": slab-decompose the grid over k GPUs of this box (same as LORA_NGPU=k)
        if (std::strcmp(argv[i], "--gpus") == 0 && i + 1 < argc) lora_set_gpus(atoi(argv[++i]));
        // trailing "--weights FILE": a caller's table instead of the hard-coded one, every weight honoured
        if (std::strcmp(argv[i], "--weights") == 0 && i + 1 < argc) weights_path = argv[++i];
    }
    double params[125];
    lora_reference_table(sn->shape, params);
    int mode = LORA_WEIGHTS_REFERENCE;
    if (weights_path) {
        std::ifstream f(weights_path);
        if (!f) {
            std::cerr << "Invalid argument: cannot open the weight file " << weights_path << ".\n";
            return 1;
        }
        std::vector<double> w;
        for (double v; f >> v;) w.push_back(v);
        if (!f.eof() || (int)w.size() != nparams) {
            std::cerr << "Invalid argument: the weight file must hold exactly " << nparams << " numbers (found " << w.size()
                      << (f.eof() ? "" : ", then something that is not a number") << ").\n";
            return 1;
        }
        for (int i = 0; i < nparams; i++) params[i] = w[i];
        mode = LORA_WEIGHTS_GENERAL;
    }

#if LORA_CLI_DIM == 1
    printf("INFO: shape = %s, n = %lld, times = %d\n", sn->info, dims[0], times);
#elif LORA_CLI_DIM == 2
    printf("INFO: shape = %s, m = %lld, n = %lld, times = %d\n", sn->info, dims[0], dims[1], times);
#else
    printf("INFO: shape = %s, h = %lld, m = %lld, n = %lld, times = %d\n", sn->info, dims[0], dims[1], dims[2],
           times);
#endif

    if (weights_path) printf("INFO: weights = %s (%d values, every one honoured)\n", weights_path, nparams);

    long long pd[3] = {1, 1, 1}, total = 1;
    for (int i = 0; i < kDim; i++) {
        if (dims[i] <= 0) {
            std::cerr << "Argument out of range: sizes must be positive.\n";
            return 1;
        }
        pd[i] = dims[i] + 2 * halo[i];
        total *= pd[i];
    }
    std::vector<double> matrix((size_t)total + 1), output((size_t)total + 1, 0.0);
    // FILL_RANDOM of the reference: unseeded rand() over the whole padded array, halo included
    for (long long i = 0; i < total; i++) matrix[i] = (double)(rand() % kFillMod);

    if (check) {
        std::cout << arg1 << std::endl;
#if LORA_CLI_DIM == 1
        for (int i = 0; i < 9; i++) std::cout << params[i] << std::endl;
#elif LORA_CLI_DIM == 2
        for (int i = 0; i < 7; i++) {
            for (int j = 0; j < 7; j++) std::cout << params[i * 7 + j] << " ";
            std::cout << std::endl;
        }
#else
        const int W = 2 * radius + 1;
        for (int h = 0; h < W; h++) {
            for (int r = 0; r < W; r++) {
                for (int c = 0; c < W; c++) std::cout << params[(h * W + r) * W + c] << " ";
                std::cout << std::endl;
            }
            std::cout << std::endl;
        }
#endif
    }

    lora_gpu_run_host(sn->shape, mode, matrix.data(), output.data(), params, times, dims);

    int rc = 0;
    if (check) {
        printf("\nChecking Correctness... \n");
        std::vector<double> naive((size_t)total + 1, 0.0), lora((size_t)total + 1, 0.0);
        cpu_step(matrix, naive, params, pd, halo, radius);
        lora_gpu_run_host(sn->shape, mode, matrix.data(), lora.data(), params, 1, dims);
        printf("Comparing naive and lora\n");
        long long bad = 0;
        const long long lo[3] = {halo[0], halo[1], halo[2]};
        for (long long a = lo[0]; a < pd[0] - lo[0]; a++)
            for (long long b = (kDim > 1 ? lo[1] : 0); b < (kDim > 1 ? pd[1] - lo[1] : 1); b++)
                for (long long c = (kDim > 2 ? lo[2] : 0); c < (kDim > 2 ? pd[2] - lo[2] : 1); c++) {
                    const long long idx = (kDim == 1) ? a : (kDim == 2) ? a * pd[1] + b : (a * pd[1] + b) * pd[2] + c;
                    if (std::fabs(naive[idx] - lora[idx]) > 1e-7) {
                        if (bad < 100)
                            printf("index = %lld, naive = %lf, lora = %lf\n", idx, naive[idx], lora[idx]);
                        bad++;
                    }
                }
        if (bad) {
            printf("%lld mismatching cells\n", bad);
            rc = 2;
        }
        printf("Correct!\n");
    }
    return rc;
}
