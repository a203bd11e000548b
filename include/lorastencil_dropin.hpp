// lorastencil_dropin.hpp -- the reference's own C++ operator prototypes, exported by
// liblorastencil_b200.so with C++ linkage so that the reference's unmodified main.cu
// (src/{1d,2d,3d}/main.cu) links against this library instead of gpu_*.cu.
//
// Mangled names (checked by tests/test_abi.py):
//   _Z8gpu_1d1rPKdPdS0_ii            src/1d/1d_utils.h:45
//   _Z8gpu_1d2rPKdPdS0_ii            src/1d/1d_utils.h:47
//   _Z13gpu_star_2d1rPKdPdS0_iii     src/2d/2d_utils.h:47
//   _Z13gpu_star_2d3rPKdPdS0_iii     src/2d/2d_utils.h:49
//   _Z12gpu_box_2d3rPKdPdS0_iii      src/2d/2d_utils.h:51
//   _Z12gpu_box_3d1rPKdPdS0_iiii     src/3d/3d_utils.h:44
//   _Z13gpu_star_3d1rPKdPdS0_iiii    src/3d/3d_utils.h:47
// (test_gpu_star_2d1r, src/2d/2d_utils.h:53, has no definition or caller upstream: not exported.)
#pragma once

void gpu_1d1r(const double *__restrict__ in, double *__restrict__ out, const double *__restrict__ params,
              const int time, const int input_n);
void gpu_1d2r(const double *__restrict__ in, double *__restrict__ out, const double *__restrict__ params,
              const int time, const int input_n);

void gpu_star_2d1r(const double *__restrict__ in, double *__restrict__ out, const double *__restrict__ params,
                   const int times, const int input_m, const int input_n);
void gpu_star_2d3r(const double *__restrict__ in, double *__restrict__ out, const double *__restrict__ params,
                   const int times, const int input_m, const int input_n);
void gpu_box_2d3r(const double *__restrict__ in, double *__restrict__ out, const double *__restrict__ params,
                  const int times, const int input_m, const int input_n);

void gpu_box_3d1r(const double *__restrict__ in, double *__restrict__ out, const double *__restrict__ params,
                  const int times, const int input_h, const int input_m, const int input_n);
void gpu_star_3d1r(const double *__restrict__ in, double *__restrict__ out, const double *__restrict__ params,
                   const int times, const int input_h, const int input_m, const int input_n);
