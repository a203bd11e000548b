// ref_stubs.cpp -- TEST INFRASTRUCTURE ONLY (see oracle/oracle.c header).
//
// The reference's main.cu files (src/{1d,2d,3d}/main.cu) compile unmodified with
// `g++ -x c++ -Dmain=ref_main` and leave the gpu_* host operators undefined.  To load the
// resulting object as a shared library (for its verbatim `test_cpu`) those symbols have to
// resolve; these stubs satisfy the linker and abort if anybody ever calls them.
// Signatures follow src/1d/1d_utils.h:45-47, src/2d/2d_utils.h:47-51, src/3d/3d_utils.h:44-48.
#include <cstdio>
#include <cstdlib>

[[noreturn]] static void never(const char *name) {
    std::fprintf(stderr, "oracle/_ref stub %s called: the CPU reference library has no GPU path\n", name);
    std::abort();
}

#if defined(REF_DIM) && REF_DIM == 1
void gpu_1d1r(const double *__restrict__, double *__restrict__, const double *__restrict__, const int, const int) { never("gpu_1d1r"); }
void gpu_1d2r(const double *__restrict__, double *__restrict__, const double *__restrict__, const int, const int) { never("gpu_1d2r"); }
#elif defined(REF_DIM) && REF_DIM == 2
void gpu_star_2d1r(const double *__restrict__, double *__restrict__, const double *__restrict__, const int, const int, const int) { never("gpu_star_2d1r"); }
void gpu_star_2d3r(const double *__restrict__, double *__restrict__, const double *__restrict__, const int, const int, const int) { never("gpu_star_2d3r"); }
void gpu_box_2d3r(const double *__restrict__, double *__restrict__, const double *__restrict__, const int, const int, const int) { never("gpu_box_2d3r"); }
#elif defined(REF_DIM) && REF_DIM == 3
void gpu_box_3d1r(const double *__restrict__, double *__restrict__, const double *__restrict__, const int, const int, const int, const int) { never("gpu_box_3d1r"); }
void gpu_star_3d1r(const double *__restrict__, double *__restrict__, const double *__restrict__, const int, const int, const int, const int) { never("gpu_star_3d1r"); }
#else
#error "compile with -DREF_DIM=1|2|3"
#endif
