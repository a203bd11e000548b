// decompose.h -- host-side low-rank adaptation of stencil weight tables (C++).
//
// Replaces the reference's hard-wired host factorisations
//   src/2d/gpu.cu:280-350   pyramidal rank-1 peel of the 7x7 box weights
//   src/2d/gpu.cu:433-444   cross (star2d3r): column arm + row arm
//   src/2d/gpu.cu:486-487   diamond (star2d1r): fixed rank-1 pair, residual at :249-264
//   src/3d/gpu_box.cu:151-164, src/3d/gpu_star.cu:142-151   3-D separable / 7-point
// with one general routine per dimension that (a) reproduces those forms on the reference's
// tables (mode REFERENCE, quirks included) and (b) finds the cheapest exact form for any table
// (mode GENERAL), falling back to direct taps when the table is not low-rank.
#pragma once
#include <string>

#include "../../include/lorastencil.h"

namespace lora {

struct Decomp1D {
    double w[9];  // tap k multiplies in[c - 4 + k]
};

struct Decomp2D {
    int form = LORA_FORM_DIRECT49;
    int nterms = 0;
    double vert[3][7] = {};    // [t][dr+3]
    double horiz[3][7] = {};   // [t][dc+3]
    double centre = 0.0;
    double residual[8] = {};   // (0,-3),(0,+3),(-3,0),(+3,0),(-2,-2),(-2,+2),(+2,-2),(+2,+2)
    double direct[49] = {};    // [ (dr+3)*7 + dc+3 ], form DIRECT49 only
    double effective[49] = {}; // what the chosen form applies, as direct taps
    double recon_err = 0.0;
    int macs = 49;
    std::string desc;
};

struct Decomp3D {
    int form = LORA_FORM_DIRECT27;
    double a[3] = {}, b[3] = {}, c[3] = {};  // SEP3: w[dh][dr][dc] = a[dh+1] b[dr+1] c[dc+1]
    double star[7] = {};                      // STAR7: centre, n-1, n+1, m-1, m+1, h-1, h+1
    double direct[27] = {};                   // [ (dh+1)*9 + (dr+1)*3 + dc+1 ]
    double effective[27] = {};
    double recon_err = 0.0;
    int macs = 27;
    std::string desc;
};

// radius-2 3-D shapes (box3d2r / star3d2r; stencil3d_r2.cu): their own layout and 125 weights
struct Decomp3DR2 {
    int form = LORA_FORM_DIRECT125;
    double w[125] = {};   // effective taps [(dh+2)*25 + (dr+2)*5 + dc+2] (== the table: every weight is honoured)
    double q[25] = {};    // HSEP5: w[dh][dr][dc] = a[dh+2] * q[(dr+2)*5 + dc+2]
    double a[5] = {};
    double b[5] = {}, c[5] = {};  // SEP5: q = b (x) c
    double recon_err = 0.0;
    int macs = 125;
    std::string desc;
};

int shape_dim(int shape);                 // 1, 2, 3 -- or 0 when invalid AND for the radius-2 shapes, which only the code
                                          // paths that ask shape_is_r2() know (everything else refuses them)
bool shape_is_r2(int shape);              // LORA_BOX3D2R / LORA_STAR3D2R
int shape_nparams(int shape);             // 9, 49, 27, 125
const char *shape_cli_name(int shape);    // "box2d1r" ...
const char *shape_banner(int shape);      // "2D box_2d3r" ... as printed by the reference operator
int shape_artifact_k(int shape);          // the K multiplier of the reference's GStencil/s printout

// the weight table the reference CLI passes for `shape` (src/*/main.cu)
void reference_table(int shape, double *out);

bool decompose_1d(int shape, int mode, const double *params, Decomp1D &d);
bool decompose_2d(int shape, int mode, const double *params, Decomp2D &d);
bool decompose_3d(int shape, int mode, const double *params, Decomp3D &d);
bool decompose_3d_r2(int shape, const double *params, Decomp3DR2 &d);  // LORA_R2_SEP5=0|1 overrides the SEP5 default

}  // namespace lora
