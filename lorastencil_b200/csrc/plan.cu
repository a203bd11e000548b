// plan.cu -- host side of liblorastencil_b200.so: plans, TMA descriptors, launch geometry, the
// drop-in host operators and the C ABI declared in include/lorastencil.h.
//
// Reference counterparts: the gpu_* host operators
//   src/1d/gpu_1r.cu:90-137, src/1d/gpu_2r.cu:91-137, src/2d/gpu.cu:276-557,
//   src/3d/gpu_box.cu:143-226, src/3d/gpu_star.cu:136-195
// (weight factorisation -> upload -> cudaMalloc x2 -> launch loop -> timing printout -> D2H).
#include <algorithm>
#include <chrono>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <tuple>
#include <string>
#include <vector>

#include "../../include/lorastencil.h"
#include "../../include/lorastencil_dropin.hpp"
#include "decompose.h"
#include "exchange.h"
#include "hostmove.h"
#include "kernels.h"

using namespace lora;

// ---------------------------------------------------------------------------------------------
// error plumbing
// ---------------------------------------------------------------------------------------------
static thread_local std::string g_err;

static int fail(int code, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}

#define CU_TRY(call)                                                                              \
    do {                                                                                          \
        cudaError_t e__ = (call);                                                                 \
        if (e__ != cudaSuccess)                                                                   \
            return fail(LORA_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
    } while (0)

int lora_fail(int code, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}

extern "C" const char *lora_last_error(void) { return g_err.c_str(); }

// ---------------------------------------------------------------------------------------------
// driver entry point for TMA descriptors (no link-time dependency on libcuda)
// ---------------------------------------------------------------------------------------------
typedef CUresult (*encode_tiled_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                    const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static encode_tiled_fn get_encode() {
    static encode_tiled_fn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<encode_tiled_fn>(p);
    });
    return fn;
}

// ---------------------------------------------------------------------------------------------
// plan
// ---------------------------------------------------------------------------------------------
struct lora_plan {
    int shape = 0, mode = 0, dim = 0;
    long long dims[3] = {0, 0, 0};    // interior
    long long padded[3] = {0, 0, 0};  // with halos
    int halo[3] = {0, 0, 0};          // storage halo per axis: S1 (4 | 4,4 | 1,2,4); radius-2 3-D shapes 2,2,4
    int radius0 = 0;                  // stencil radius along the outermost axis (4 | 3 | 1; radius-2 shapes 2)
    bool r2 = false;                  // LORA_BOX3D2R / LORA_STAR3D2R: own layout, own kernel (stencil3d_r2.cu)
    WeightsR2 wr2{};
    long long elems = 0;
    int form = 0;
    Weights1D w1{};
    Weights2D w2{};
    WeightsDirect49 wd{};
    Weights3D w3{};
    std::string desc;
    std::vector<std::pair<const void *, CUtensorMap>> maps;  // one TMA descriptor per source buffer
    struct Map1D {
        const void *ptr;
        long long off;
        CUtensorMap map;
    };
    std::vector<Map1D> maps1d;  // 1-D line viewed as rows of 16 doubles starting at cell `off`, 128B swizzle
    std::vector<std::pair<const void *, CUtensorMap>> maps3tb;  // 3-D, the fused kernel's box (kT3BoxRows rows)
    long long launches = 0;
    int sm_count = 148;
    int slots = 148 * 16;  // concurrently resident warp workers (1-D / 2-D) or CTAs (3-D)
    int device = 0;
    int max_tb = 1;        // deepest temporal block lora_plan_run may fuse (1 = one launch per time step)
    int boundary = LORA_BOUNDARY_REFERENCE;  // what the halo ring means from launch to launch (lora_plan_set_boundary)
    bool tb_auto = false;  // 2-D cross / diamond: max_tb is the form's default until lora_plan_run has measured both ways
    bool odd_cols = false; // 2-D / 3-D with an odd number of padded columns: no tensor map possible, direct-tap kernel
    WeightsDirect49 eff{}; // the effective direct taps (49, or 27 in 3-D) the chosen form equals
};

constexpr int kTb3 = 2;  // the 3-D temporal block (stencil3d_tb.cu)
static bool tb3_form(int form) { return form == LORA_FORM_STAR7 || form == LORA_FORM_SEP3; }
constexpr int kTb2 = 3;  // the 2-D temporal block (odd, so that time parity == buffer parity at every sweep)
static bool tb2_form(int form) {
    return form == LORA_FORM_CROSS || form == LORA_FORM_DIAMOND || form == LORA_FORM_PYRAMID || form == LORA_FORM_PYRAMID_PRUNED;
}
// ... and sweeps of TWO launches (no register spills where three spill): every form but the cross, which is cheap at three
static bool tb2_pair_form(int form) { return tb2_form(form) && form != LORA_FORM_CROSS; }
static int step_fused_3d(lora_plan *p, const double *src, double *dst, long long lo, long long hi, void *stream,
                         const lora_exchange *ex);
static int step_fused_2d(lora_plan *p, const double *src, double *dst, const double *halo_src, long long lo, long long hi,
                         int tb, int launches_before, int virt_lo, int virt_hi, const lora_exchange *ex,
                         const double *mirror_base, void *stream);

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) applies to the device that is current when it runs, and every
// kernel here needs more than the 48 KB default: the opt-in is done once per DEVICE, not once per process
constexpr int kMaxDevices = 64;
static std::mutex g_init_mutex;
static int g_init_state[kMaxDevices];  // 0 = not yet, 1 = done, 2 = failed
static cudaError_t g_init_err[kMaxDevices];

static int ensure_init(int *device_out = nullptr) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return fail(LORA_ERR_CUDA, "cudaGetDevice failed: %s", cudaGetErrorString(e));
    if (dev < 0 || dev >= kMaxDevices) return fail(LORA_ERR_UNSUPPORTED, "device ordinal %d out of range", dev);
    if (device_out) *device_out = dev;
    if (g_init_state[dev] != 1) {
        std::lock_guard<std::mutex> lk(g_init_mutex);
        if (g_init_state[dev] == 0) {
            g_init_err[dev] = kernels_init();
            g_init_state[dev] = g_init_err[dev] == cudaSuccess ? 1 : 2;
        }
        if (g_init_state[dev] != 1)
            return fail(LORA_ERR_CUDA, "kernel attribute setup failed on device %d: %s", dev, cudaGetErrorString(g_init_err[dev]));
    }
    return LORA_OK;
}

// a plan's tensor maps and task geometry belong to the device it was created on
static int check_device(const lora_plan *p) {
    int dev = -1;
    int rc = ensure_init(&dev);
    if (rc) return rc;
    if (dev != p->device)
        return fail(LORA_ERR_ARG, "plan was created on device %d but device %d is current (cudaSetDevice first)", p->device, dev);
    return LORA_OK;
}

extern "C" int lora_plan_create(lora_plan_t **out, int shape, int mode, const double *params, const long long *dims) {
    if (!out || !dims) return fail(LORA_ERR_ARG, "null argument");
    const bool r2 = shape_is_r2(shape);
    const int dim = r2 ? 3 : shape_dim(shape);
    if (dim == 0) return fail(LORA_ERR_ARG, "unknown shape %d", shape);
    if (mode != LORA_WEIGHTS_REFERENCE && mode != LORA_WEIGHTS_GENERAL) return fail(LORA_ERR_ARG, "bad mode %d", mode);
    for (int i = 0; i < dim; i++)
        if (dims[i] <= 0 || dims[i] > 0x7fffffffLL - 16) return fail(LORA_ERR_ARG, "bad size %lld", dims[i]);
    int rc = ensure_init();
    if (rc) return rc;

    lora_plan *p = new lora_plan;
    p->shape = shape;
    p->mode = mode;
    p->dim = dim;
    p->r2 = r2;
    static const int halo[5][3] = {{0, 0, 0}, {4, 0, 0}, {4, 4, 0}, {1, 2, 4}, {2, 2, 4}};
    static const int radius0[5] = {0, 4, 3, 1, 2};
    p->radius0 = radius0[r2 ? 4 : dim];
    p->elems = 1;
    for (int i = 0; i < dim; i++) {
        p->dims[i] = dims[i];
        p->halo[i] = halo[r2 ? 4 : dim][i];
        p->padded[i] = dims[i] + 2 * p->halo[i];
        p->elems *= p->padded[i];
    }
    double table[125];
    if (!params) {
        reference_table(shape, table);
        params = table;
    }
    if (r2) {
        Decomp3DR2 d;
        decompose_3d_r2(shape, params, d);
        std::memcpy(p->wr2.w, d.w, sizeof d.w);
        std::memcpy(p->wr2.q, d.q, sizeof d.q);
        std::memcpy(p->wr2.a, d.a, sizeof d.a);
        std::memcpy(p->wr2.b, d.b, sizeof d.b);
        std::memcpy(p->wr2.c, d.c, sizeof d.c);
        p->form = d.form;
        p->desc = d.desc;
    } else if (dim == 1) {
        Decomp1D d;
        decompose_1d(shape, mode, params, d);
        std::memcpy(p->w1.w, d.w, sizeof d.w);
        p->form = LORA_FORM_TAPS9;
        p->desc = (d.w[0] == 0.0 && d.w[8] == 0.0) ? "1d 9 direct taps, outer two are zero: 7 computed in fused sweeps"
                                                   : "1d 9 direct taps";
    } else if (dim == 2) {
        Decomp2D d;
        decompose_2d(shape, mode, params, d);
        std::memcpy(p->w2.vert, d.vert, sizeof d.vert);
        std::memcpy(p->w2.horiz, d.horiz, sizeof d.horiz);
        p->w2.centre = d.centre;
        std::memcpy(p->w2.residual, d.residual, sizeof d.residual);
        std::memcpy(p->wd.w, d.direct, sizeof d.direct);
        std::memcpy(p->eff.w, d.effective, sizeof d.effective);
        p->form = d.form;
        p->desc = d.desc;
    } else {
        Decomp3D d;
        decompose_3d(shape, mode, params, d);
        std::memcpy(p->w3.a, d.a, sizeof d.a);
        std::memcpy(p->w3.b, d.b, sizeof d.b);
        std::memcpy(p->w3.c, d.c, sizeof d.c);
        std::memcpy(p->w3.star, d.star, sizeof d.star);
        std::memcpy(p->w3.direct, d.direct, sizeof d.direct);
        std::memcpy(p->eff.w, d.effective, sizeof d.effective);
        p->form = d.form;
        p->desc = d.desc;
    }
    if (!r2 && dim >= 2 && p->padded[dim - 1] % 2) {
        // odd row length: rows are not 16-byte aligned, no TMA descriptor exists for this grid (stencil_direct.cu)
        p->odd_cols = true;
        p->desc += " [odd column count: direct taps without TMA]";
    }
    ensure_init(&p->device);
    int sms = 0;
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, p->device) == cudaSuccess && sms > 0) p->sm_count = sms;
    // task-planning granularity of the 1-D / 2-D kernels, in warp tasks per SM and wave.  12 warps are resident (3 CTAs:
    // the per-warp TMA ring, and the registers of the pyramid / direct forms); planning the cheap forms for 16 --
    // i.e. more, shorter tasks than resident warps -- measured 8 % faster (dynamic CTA dispatch evens out the tail)
    int warps_per_sm =
        (p->form == LORA_FORM_PYRAMID || p->form == LORA_FORM_PYRAMID_PRUNED || p->form == LORA_FORM_DIRECT49 ||
         p->form == LORA_FORM_RANK2 || p->form == LORA_FORM_RANK3) ? 12 : 16;
    if (const char *e = getenv("LORA_SLOTS_PER_SM")) {  // tuning knob
        const int v = atoi(e);
        if (v >= 1 && v <= 64) warps_per_sm = v;
    }
    p->slots = (dim == 3) ? p->sm_count : p->sm_count * warps_per_sm;
    if (r2) {  // one launch per time step, no TMA descriptor, no fused sweeps
        *out = p;
        return LORA_OK;
    }
    if (dim == 2 && tb2_form(p->form) && !p->odd_cols) {
        // 2-D fusion (stencil2d_tb.cu), 10240^2 on B200, GStencil/s: the cross form runs sweeps of three launches (627 vs
        // 369 unfused); the diamond and pyramid forms sweeps of two (538 vs 370, three: 515; 418 vs 372, three: 318)
        p->max_tb = p->form == LORA_FORM_CROSS ? kTb2 : 2;
        // ... and for the cross form a large grid settles it by measurement on first use (probe_tb2)
        p->tb_auto = p->max_tb == kTb2 && p->elems >= (1LL << 22) && p->dims[0] >= 512;
        if (const char *e = getenv("LORA_TB2")) {
            const int v = atoi(e);
            p->max_tb = v >= kTb2 ? kTb2 : ((v == 2 && tb2_pair_form(p->form)) ? 2 : 1);
            p->tb_auto = false;
        }
    }
    if (dim == 3 && tb3_form(p->form) && !p->odd_cols) {
        // 3-D fusion (stencil3d_tb.cu): two launches per sweep, on by default for both forms (512^3: 7-point 581 vs
        // 380 GStencil/s, separable 452 vs 389; 1024^3: 601 vs 388, 469 vs 389); LORA_TB3=1 turns it off
        p->max_tb = kTb3;
        if (const char *e = getenv("LORA_TB3")) p->max_tb = (atoi(e) >= kTb3) ? kTb3 : 1;
    }
    if (dim == 1) {
        p->max_tb = kDefaultTb1;
        if (const char *e = getenv("LORA_TB")) {
            const int v = atoi(e);
            if (v >= 1) p->max_tb = v < kMaxTb1 ? v : kMaxTb1;
        }
    }
    *out = p;
    return LORA_OK;
}

extern "C" void lora_plan_destroy(lora_plan_t *p) { delete p; }
extern "C" long long lora_plan_padded_elems(const lora_plan_t *p) { return p ? p->elems : 0; }
extern "C" long long lora_plan_launch_count(const lora_plan_t *p) { return p ? p->launches : 0; }
extern "C" const char *lora_plan_describe(const lora_plan_t *p) { return p ? p->desc.c_str() : ""; }

static int get_tmap(lora_plan *p, const double *src, const CUtensorMap **out) {
    for (auto &kv : p->maps)
        if (kv.first == src) {
            *out = &kv.second;
            return LORA_OK;
        }
    encode_tiled_fn enc = get_encode();
    if (!enc) return fail(LORA_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
    if (reinterpret_cast<uintptr_t>(src) % 16) return fail(LORA_ERR_ARG, "source buffer must be 16-byte aligned");
    CUtensorMap m;
    CUresult r;
    if (p->dim == 2) {
        if (p->padded[1] % 2) return fail(LORA_ERR_UNSUPPORTED, "2-D TMA path needs an even number of columns");
        cuuint64_t gdim[2] = {(cuuint64_t)p->padded[1], (cuuint64_t)p->padded[0]};
        cuuint64_t gstr[1] = {(cuuint64_t)p->padded[1] * 8};
        cuuint32_t box[2] = {(cuuint32_t)kBoxCols, (cuuint32_t)kRowsPerStage};
        cuuint32_t es[2] = {1, 1};
        r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<double *>(src), gdim, gstr, box, es,
                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    } else {
        if (p->padded[2] % 2) return fail(LORA_ERR_UNSUPPORTED, "3-D TMA path needs an even number of columns");
        cuuint64_t gdim[3] = {(cuuint64_t)p->padded[2], (cuuint64_t)p->padded[1], (cuuint64_t)p->padded[0]};
        cuuint64_t gstr[2] = {(cuuint64_t)p->padded[2] * 8, (cuuint64_t)p->padded[2] * p->padded[1] * 8};
        cuuint32_t box[3] = {(cuuint32_t)k3BoxCols, (cuuint32_t)k3BoxRows, 1};
        cuuint32_t es[3] = {1, 1, 1};
        r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, const_cast<double *>(src), gdim, gstr, box, es,
                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    }
    if (r != CUDA_SUCCESS) return fail(LORA_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
    if (p->maps.size() >= 8) p->maps.erase(p->maps.begin());
    p->maps.emplace_back(src, m);
    *out = &p->maps.back().second;
    return LORA_OK;
}

// 3-D fused sweeps: same tensor, boxes of k3BoxCols x kT3BoxRows x 1
static int get_tmap3tb(lora_plan *p, const double *src, const CUtensorMap **out) {
    for (auto &kv : p->maps3tb)
        if (kv.first == src) {
            *out = &kv.second;
            return LORA_OK;
        }
    encode_tiled_fn enc = get_encode();
    if (!enc) return fail(LORA_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
    if (reinterpret_cast<uintptr_t>(src) % 16) return fail(LORA_ERR_ARG, "source buffer must be 16-byte aligned");
    CUtensorMap m;
    cuuint64_t gdim[3] = {(cuuint64_t)p->padded[2], (cuuint64_t)p->padded[1], (cuuint64_t)p->padded[0]};
    cuuint64_t gstr[2] = {(cuuint64_t)p->padded[2] * 8, (cuuint64_t)p->padded[2] * p->padded[1] * 8};
    cuuint32_t box[3] = {(cuuint32_t)k3BoxCols, (cuuint32_t)kT3BoxRows, 1};
    cuuint32_t es[3] = {1, 1, 1};
    CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, const_cast<double *>(src), gdim, gstr, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(LORA_ERR_CUDA, "cuTensorMapEncodeTiled (fused 3-D box) failed with CUresult %d", (int)r);
    if (p->maps3tb.size() >= 8) p->maps3tb.erase(p->maps3tb.begin());
    p->maps3tb.emplace_back(src, m);
    *out = &p->maps3tb.back().second;
    return LORA_OK;
}

// 1-D temporal blocking: the padded line from cell `off` on as a (rows x 16) tensor, boxes of 32 rows = 512 cells
// = one warp row, 128-byte swizzle (stencil1d_tb.cu).  Used for loads (off 0) and stores (off = -4 tb mod 16).
static int get_tmap1d(lora_plan *p, const double *ptr, long long off, long long rows, CUtensorMap *out) {
    for (auto &m : p->maps1d)
        if (m.ptr == ptr && m.off == off) {
            *out = m.map;
            return LORA_OK;
        }
    encode_tiled_fn enc = get_encode();
    if (!enc) return fail(LORA_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
    lora_plan::Map1D m;
    m.ptr = ptr;
    m.off = off;
    cuuint64_t gdim[2] = {16, (cuuint64_t)rows};
    cuuint64_t gstr[1] = {128};
    cuuint32_t box[2] = {16, (cuuint32_t)(kTbRowCells / 16)};  // one warp row
    cuuint32_t es[2] = {1, 1};
    CUresult r = enc(&m.map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<double *>(ptr + off), gdim, gstr, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(LORA_ERR_CUDA, "cuTensorMapEncodeTiled (1-D line) failed with CUresult %d", (int)r);
    if (p->maps1d.size() >= 24) p->maps1d.erase(p->maps1d.begin());
    p->maps1d.push_back(m);
    *out = m.map;
    return LORA_OK;
}

// rows (2-D), 128-element rows (1-D) or planes (3-D) per task: the smallest whole number of
// "waves" k over `slots` concurrent workers whose tasks stay below `max_len`
static long long pick_len(long long total, long long lanes, long long slots, long long max_len, long long min_len) {
    if (total <= min_len) return total;
    for (long long k = 1; k < 64; k++) {
        long long q = (k * slots) / lanes;  // tasks along the swept axis
        if (q < 1) q = 1;
        long long len = (total + q - 1) / q;
        if (len <= max_len) return len < min_len ? min_len : len;
    }
    return max_len;
}

// 3-D plane chunks: a CTA streams its chunk of L planes through the tile (L + 2 plane loads, ~2 more plane times to
// fill its pipeline), and CTAs run in waves of `slots` (one per SM).  Pick the chunk count that minimises
// waves x (L + 4) among chunks of at most max_len planes: 512^3 -> 9 chunks of 57 (3.9 waves); a slab of 126 planes x
// 256 tiles (1024^3 on 8 GPUs) -> 4 chunks of 32 (6.9 waves) instead of 2 chunks of 63 (3.5 waves: a quarter of the
// last wave idle for 65 plane times).
static long long pick_chunk_3d(long long planes, long long tiles, long long slots, long long max_len, long long overhead = 4) {
    long long best_len = planes < max_len ? planes : max_len, best_cost = -1;
    for (long long c = 1; c <= planes; c++) {
        const long long L = (planes + c - 1) / c;
        if (L > max_len) continue;
        if (L < 8 && c > 1) break;
        const long long waves = (c * tiles + slots - 1) / slots;
        const long long cost = waves * (L + overhead);
        if (best_cost < 0 || cost < best_cost) best_cost = cost, best_len = L;
    }
    return best_len;
}

// ---------------------------------------------------------------------------------------------
// segments: how a launch's range is cut when the launch also serves neighbouring slabs (kernels.h: Segs)
// ---------------------------------------------------------------------------------------------
struct SegCut {
    int n = 0;
    long long lo[kMaxSegs], hi[kMaxSegs], mirror[kMaxSegs], mlo[kMaxSegs], mhi[kMaxSegs];
    unsigned long long *flag[kMaxSegs], *count[kMaxSegs], *arrived[kMaxSegs];
    bool band[kMaxSegs];
    int early[kMaxSegs];
    long long flag_tasks[kMaxSegs];  // >= 0: only this many tasks of the segment arrive on its flag (default: all)
    unsigned long long seq = 0;
    void add(long long a, long long b, long long mir, bool is_band, unsigned long long *f, unsigned long long *c,
             unsigned long long *arr) {
        lo[n] = a, hi[n] = b, mirror[n] = mir, band[n] = is_band, flag[n] = f, count[n] = c, arrived[n] = arr;
        mlo[n] = a, mhi[n] = b, early[n] = 0, flag_tasks[n] = -1;
        n++;
    }
    bool aligned4() const {
        for (int i = 0; i < n; i++)
            if (mirror[i] % 4) return false;
        return true;
    }
};

// [lo, hi) -> [lo band][hi band][middle] (bands first: their tasks are dispatched first).  mirror_whole: the
// older whole-launch form (every cell of the launch is mirrored, no flag).  Coordinates: whatever the caller uses.
static int cut_segments(long long lo, long long hi, const double *dst, const lora_exchange *ex, const double *mirror_whole,
                        SegCut &sc) {
    if (!ex || (ex->band_lo <= 0 && ex->band_hi <= 0)) {
        sc.add(lo, hi, mirror_whole ? (long long)(mirror_whole - dst) : 0, false, nullptr, nullptr, nullptr);
        return LORA_OK;
    }
    const long long bl = ex->band_lo > 0 ? ex->band_lo : 0, bh = ex->band_hi > 0 ? ex->band_hi : 0;
    if (bl + bh > hi - lo) return fail(LORA_ERR_ARG, "exchange bands (%lld + %lld) overlap in a range of %lld", bl, bh, hi - lo);
    if ((bl && !ex->mirror_lo) || (bh && !ex->mirror_hi)) return fail(LORA_ERR_ARG, "exchange band without a mirror address");
    sc.seq = ex->seq;
    if (bl) sc.add(lo, lo + bl, (long long)(ex->mirror_lo - dst), true, ex->flag_lo, ex->count_lo, ex->arrived_lo);
    if (bh) sc.add(hi - bh, hi, (long long)(ex->mirror_hi - dst), true, ex->flag_hi, ex->count_hi, ex->arrived_hi);
    if (lo + bl < hi - bh) sc.add(lo + bl, hi - bh, 0, false, nullptr, nullptr, nullptr);
    return LORA_OK;
}

// chunk[] and tasks-per-segment known: fill the kernel's table; band segments with a flag get their arrival target
static void fill_segs(Segs &sg, const SegCut &sc, const long long *chunk, const long long *tasks, long long arrivals_per_task) {
    std::memset(&sg, 0, sizeof sg);
    sg.nseg = sc.n;
    sg.seq = sc.seq;
    sg.first[0] = 0;
    for (int i = 0; i < sc.n; i++) {
        sg.lo[i] = sc.lo[i];
        sg.hi[i] = sc.hi[i];
        sg.chunk[i] = chunk[i];
        sg.first[i + 1] = sg.first[i] + tasks[i];
        sg.mirror[i] = sc.mirror[i];
        sg.mlo[i] = sc.mlo[i];
        sg.mhi[i] = sc.mhi[i];
        sg.early[i] = sc.early[i];
        if (sc.flag[i] && sc.count[i] && sc.arrived[i]) {
            *sc.arrived[i] += (unsigned long long)((sc.flag_tasks[i] >= 0 ? sc.flag_tasks[i] : tasks[i]) * arrivals_per_task);
            sg.flag[i] = sc.flag[i];
            sg.count[i] = sc.count[i];
            sg.target[i] = *sc.arrived[i];
        }
    }
}

static int step_unfused(lora_plan_t *p, const double *src, double *dst, long long lo, long long hi,
                        const lora_exchange *ex, const double *mirror_base, void *stream) {
    if (!p || !src || !dst) return fail(LORA_ERR_ARG, "null argument");
    if (lo < 0 || hi > p->dims[0] || lo > hi) return fail(LORA_ERR_ARG, "bad range [%lld, %lld)", lo, hi);
    if (lo == hi) return LORA_OK;
    if (int rc = check_device(p)) return rc;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    cudaError_t e;
    if (p->dim == 1) {
        if (ex && (ex->band_lo > 0 || ex->band_hi > 0))
            return fail(LORA_ERR_UNSUPPORTED, "1-D slab launches go through the temporally blocked kernel (tb >= 1)");
        if (lo % 2) return fail(LORA_ERR_ARG, "1-D range must start at an even index");
        if (reinterpret_cast<uintptr_t>(src) % 16 || reinterpret_cast<uintptr_t>(dst) % 16)
            return fail(LORA_ERR_ARG, "buffers must be 16-byte aligned");
        Geom1D g;
        g.in = src;
        g.out = dst;
        g.n = p->dims[0];
        g.lo = lo;
        g.hi = hi;
        const long long rows = (hi - lo + kWarpCols - 1) / kWarpCols;
        // short tasks again (see the 2-D planner): 16..32 rows of 128 cells per task measured 420-424 GStencil/s
        // against 388 at 256 rows and 319 at 1024 (105 % of the measured copy bandwidth)
        long long max_rows1 = 32;
        if (const char *e = getenv("LORA_MAX_ROWS_1D")) {  // tuning knob
            const long long v = atoll(e);
            if (v >= 4 && v <= 65536) max_rows1 = v;
        }
        g.rows_per_task = (int)pick_len(rows, 1, p->slots, max_rows1, max_rows1 < 16 ? max_rows1 : 16);
        g.ntasks = (rows + g.rows_per_task - 1) / g.rows_per_task;
        g.vec4 = (lo % 4 == 0) && (reinterpret_cast<uintptr_t>(dst) % 32 == 0);
        g.mirror = mirror_base ? (long long)(mirror_base - dst) : 0;
        e = launch_1d(g, p->w1, st);
    } else if (p->r2) {
        if (mirror_base || (ex && (ex->band_lo > 0 || ex->band_hi > 0)))
            return fail(LORA_ERR_UNSUPPORTED, "the radius-2 3-D shapes run on one GPU (no slab exchange)");
        Geom3DR2 g;
        g.in = src;
        g.out = dst;
        g.row_pitch = p->padded[2];
        g.plane_pitch = p->padded[1] * p->padded[2];
        g.m = (int)p->dims[1];
        g.n = (int)p->dims[2];
        g.lo = lo;
        g.hi = hi;
        g.planes_per_chunk = 0;  // chosen by the launcher
        e = launch_3d_r2(p->form, g, p->wr2, p->sm_count, st);
    } else if (p->odd_cols) {
        SegCut sc;
        if (int rc = cut_segments(lo, hi, dst, ex, mirror_base, sc)) return rc;
        GeomDirect g;
        g.in = src;
        g.out = dst;
        g.pitch = p->padded[p->dim - 1];
        g.plane_pitch = p->dim == 3 ? p->padded[1] * p->padded[2] : 0;
        g.m = (int)p->dims[p->dim - 2];
        g.n = (int)p->dims[p->dim - 1];
        g.col_blocks = (g.n + 255) / 256;
        long long chunk[kMaxSegs], tasks[kMaxSegs];
        for (int i = 0; i < sc.n; i++) chunk[i] = 1, tasks[i] = sc.hi[i] - sc.lo[i];
        fill_segs(g.sg, sc, chunk, tasks, (long long)g.col_blocks * (p->dim == 3 ? g.m : 1));
        e = launch_direct(p->dim, g, p->eff, st);
    } else if (p->dim == 2) {
        const CUtensorMap *tm;
        int rc = get_tmap(p, src, &tm);
        if (rc) return rc;
        SegCut sc;
        if ((rc = cut_segments(lo, hi, dst, ex, mirror_base, sc))) return rc;
        Geom2D g;
        g.out = dst;
        g.pitch = p->padded[1];
        g.m = (int)p->dims[0];
        g.n = (int)p->dims[1];
        g.nstrips = (g.n + kWarpCols - 1) / kWarpCols;
        long long chunk[kMaxSegs], tasks[kMaxSegs];
        for (int i = 0; i < sc.n; i++) {
            const long long rows = sc.hi[i] - sc.lo[i];
            if (sc.band[i]) {  // a band is a handful of rows: one chunk per strip
                chunk[i] = rows;
                tasks[i] = g.nstrips;
                continue;
            }
            // Chunk length: about 8 tasks per resident warp, between 64 and 256 rows (6 warm-up rows per chunk).  Many
            // short tasks beat whole waves of long ones here: CTAs are dispatched in task order, so the warps that are
            // resident at any moment work on a few neighbouring row bands (DRAM page and L2 locality) and the tail of
            // the launch is short.  Measured, box2d 10240^2: 337 (256 rows) / 367 (96) / 376 (64) / 357 (32) GStencil/s;
            // 40960^2: 397 at 96..256 rows (profiles/r1_chunk_rows_2d.log).
            long long want = rows * g.nstrips / (8 * (long long)p->slots);
            want = want < 64 ? 64 : (want > 256 ? 256 : want);
            if (const char *e = getenv("LORA_MAX_ROWS_2D")) {  // tuning knob
                const long long v = atoll(e);
                if (v >= 8 && v <= 4096) want = v;
            }
            const long long nch = (rows + want - 1) / want;
            chunk[i] = (rows + nch - 1) / nch;
            tasks[i] = ((rows + chunk[i] - 1) / chunk[i]) * g.nstrips;
        }
        fill_segs(g.sg, sc, chunk, tasks, 1);
        g.ntasks = (int)g.sg.first[sc.n];
        g.vec4 = (g.n % 4 == 0) && (reinterpret_cast<uintptr_t>(dst) % 32 == 0) && sc.aligned4();
        e = launch_2d(p->form, *tm, g, p->w2, p->wd, st);
    } else {
        const CUtensorMap *tm;
        int rc = get_tmap(p, src, &tm);
        if (rc) return rc;
        Geom3D g;
        g.out = dst;
        g.row_pitch = p->padded[2];
        g.plane_pitch = p->padded[1] * p->padded[2];
        g.h = (int)p->dims[0];
        g.m = (int)p->dims[1];
        g.n = (int)p->dims[2];
        g.tiles_m = (g.m + k3TileRows - 1) / k3TileRows;
        g.tiles_n = (g.n + k3TileCols - 1) / k3TileCols;
        const long long tiles = (long long)g.tiles_m * g.tiles_n;
        long long max_planes = 64;  // longer chunks lose: 1024^3 with 4 chunks of 256 planes ran at 364 GStencil/s, 16 x 64 at 389
        if (const char *e = getenv("LORA_MAX_PLANES_3D")) {  // tuning knob
            const long long v = atoll(e);
            if (v >= 2 && v <= 4096) max_planes = v;
        }
        SegCut sc;
        long long chunk[kMaxSegs], tasks[kMaxSegs];
        const long long bl = ex && ex->band_lo > 0 ? ex->band_lo : 0, bh = ex && ex->band_hi > 0 ? ex->band_hi : 0;
        const long long L = pick_chunk_3d(hi - lo, tiles, p->slots, max_planes);
        const long long nchunks = (hi - lo + L - 1) / L, last_lo = lo + (nchunks - 1) * L;
        if ((bl || bh) && nchunks >= 2 && bl <= L && bh <= hi - last_lo && !getenv("LORA_BAND_CHUNKS_3D")) {
            // Bands FOLDED into the ordinary plane chunks: no extra CTAs.  The last chunk goes first in dispatch order --
            // its last planes are the hi band, mirrored as they are stored, flag raised when its CTAs finish (one chunk
            // time into the launch).  The first chunk comes next -- its first planes are the lo band, flag raised
            // right after they are stored.  Then the chunks in between.
            if ((bl && !ex->mirror_lo) || (bh && !ex->mirror_hi)) return fail(LORA_ERR_ARG, "exchange band without a mirror address");
            sc.seq = ex->seq;
            sc.add(last_lo, hi, bh ? (long long)(ex->mirror_hi - dst) : 0, false, bh ? ex->flag_hi : nullptr,
                   bh ? ex->count_hi : nullptr, bh ? ex->arrived_hi : nullptr);
            sc.mlo[0] = hi - bh, sc.mhi[0] = hi;
            sc.add(lo, lo + L, bl ? (long long)(ex->mirror_lo - dst) : 0, false, bl ? ex->flag_lo : nullptr,
                   bl ? ex->count_lo : nullptr, bl ? ex->arrived_lo : nullptr);
            sc.mlo[1] = lo, sc.mhi[1] = lo + bl, sc.early[1] = 1;
            if (nchunks > 2) sc.add(lo + L, last_lo, 0, false, nullptr, nullptr, nullptr);
            for (int i = 0; i < sc.n; i++) {
                chunk[i] = L;
                tasks[i] = (sc.hi[i] - sc.lo[i] + L - 1) / L;
            }
        } else {
            if ((rc = cut_segments(lo, hi, dst, ex, mirror_base, sc))) return rc;
            for (int i = 0; i < sc.n; i++) {
                const long long planes = sc.hi[i] - sc.lo[i];
                chunk[i] = sc.band[i] ? planes : pick_chunk_3d(planes, tiles, p->slots, max_planes);
                tasks[i] = (planes + chunk[i] - 1) / chunk[i];  // plane chunks (blockIdx.y); every chunk is tiles_m x tiles_n CTAs
            }
        }
        fill_segs(g.sg, sc, chunk, tasks, (long long)g.tiles_m * g.tiles_n);
        g.vec4 = (g.n % 4 == 0) && (reinterpret_cast<uintptr_t>(dst) % 32 == 0) && sc.aligned4();
        e = launch_3d(p->form, *tm, g, p->w3, st);
    }
    if (e != cudaSuccess) return fail(LORA_ERR_CUDA, "kernel launch failed: %s", cudaGetErrorString(e));
    p->launches++;
    return LORA_OK;
}

static int plan_step_impl(lora_plan_t *p, const double *src, double *dst, long long lo, long long hi,
                          const double *mirror_base, void *stream) {
    return step_unfused(p, src, dst, lo, hi, nullptr, mirror_base, stream);
}

extern "C" int lora_plan_step(lora_plan_t *p, const double *src, double *dst, long long lo, long long hi,
                              void *stream) {
    return plan_step_impl(p, src, dst, lo, hi, nullptr, stream);
}

extern "C" int lora_plan_step_mirror(lora_plan_t *p, const double *src, double *dst, long long lo, long long hi,
                                     const double *mirror_base, void *stream) {
    return plan_step_impl(p, src, dst, lo, hi, mirror_base, stream);
}

extern "C" int lora_plan_set_temporal_block(lora_plan_t *p, int tb) {
    if (!p || tb < 1) return fail(LORA_ERR_ARG, "bad argument");
    p->tb_auto = false;  // an explicit choice is final
    if (p->dim == 1)
        p->max_tb = tb < kMaxTb1 ? tb : kMaxTb1;
    else if (p->dim == 2)
        p->max_tb = (tb >= kTb2 && tb2_form(p->form) && !p->odd_cols) ? kTb2
                    : (tb == 2 && tb2_pair_form(p->form) && !p->odd_cols) ? 2 : 1;  // 2-D fuses 3 or 2 launches, or none
    else
        p->max_tb = (tb >= kTb3 && tb3_form(p->form) && !p->odd_cols) ? kTb3 : 1;  // 3-D fuses exactly 2 launches or none
    return LORA_OK;
}

// Task geometry of a fused 2-D launch over `rows` rows (g.nstrips set): see decode_task_2dtb in kernels.h
static void plan_main_2dtb(Geom2DTB &g, long long rows, int sm_count, int tb);
static void plan_tasks_2dtb(Geom2DTB &g, long long rows, long long slots, long long min_waves = 1) {
    if (g.nstrips >= 3) {
        // Inner strips are cut into nchunks tasks each; the two edge strips patch every row (about twice the time
        // per row), so they are cut into tasks of half that length (<= kEdgeRows2Tb rows: their halo columns are
        // staged in shared memory).  Take the fewest waves k and, within it, the most chunks such that all tasks fit
        // k x the resident warps: equal-length tasks in whole waves, no straggler round.
        long long edge_cap = kEdgeRows2Tb;
        if (const char *e = getenv("LORA_TB2_EDGE_ROWS")) {  // tuning knob
            const long long v = atoll(e);
            if (v >= 8 && v <= kEdgeRows2Tb) edge_cap = v;
        }
        long long best_chunks = 0, k0 = min_waves;
        if (const char *e = getenv("LORA_TB2_MIN_WAVES")) {  // tuning knob
            const long long v = atoll(e);
            if (v >= 1 && v <= 16) k0 = v;
        }
        for (long long k = k0; k <= 64 && !best_chunks; k++) {
            for (long long nc = (rows + 95) / 96; nc >= 1; nc--) {  // most chunks first
                const long long R = (rows + nc - 1) / nc;
                if (R > 768) break;
                const long long el = R / 2 < edge_cap ? (R + 1) / 2 : edge_cap;
                const long long tasks = (g.nstrips - 2) * ((rows + R - 1) / R) + 2 * ((rows + el - 1) / el);
                if (tasks <= k * slots) {
                    best_chunks = nc;
                    break;
                }
            }
        }
        if (!best_chunks) best_chunks = (rows + 767) / 768;
        if (const char *e = getenv("LORA_TB2_ROWS")) {  // tuning knob: force the chunk length
            const long long v = atoll(e);
            if (v >= 16 && v <= 768) best_chunks = (rows + v - 1) / v;
        }
        g.rows_per_chunk = (int)((rows + best_chunks - 1) / best_chunks);
        g.nchunks = (int)((rows + g.rows_per_chunk - 1) / g.rows_per_chunk);
        g.edge_rows = (int)(g.rows_per_chunk / 2 < edge_cap ? (g.rows_per_chunk + 1) / 2 : edge_cap);
        g.nedge = (int)((rows + g.edge_rows - 1) / g.edge_rows);
        g.ntasks = 2 * g.nedge + (g.nstrips - 2) * g.nchunks;
    } else {
        g.rows_per_chunk = (int)pick_len(rows, g.nstrips, slots, kEdgeRows2Tb, 32);
        g.nchunks = (int)((rows + g.rows_per_chunk - 1) / g.rows_per_chunk);
        g.edge_rows = g.rows_per_chunk;
        g.nedge = 0;
        g.ntasks = g.nstrips * g.nchunks;
    }
}

// 2-D fused launch of kTb2 time steps over interior rows [lo, hi): see stencil2d_tb.cu
static int step_fused_2d(lora_plan *p, const double *src, double *dst, const double *halo_src, long long lo, long long hi,
                         int tb, int launches_before, int virt_lo, int virt_hi, const lora_exchange *ex,
                         const double *mirror_base, void *stream) {
    if (lo < 0 || hi > p->dims[0] || lo > hi) return fail(LORA_ERR_ARG, "bad range [%lld, %lld)", lo, hi);
    if (tb == 1) return step_unfused(p, src, dst, lo, hi, ex, mirror_base, stream);
    if (!((tb == kTb2 && tb2_form(p->form)) || (tb == 2 && tb2_pair_form(p->form))) || p->odd_cols)
        return fail(LORA_ERR_UNSUPPORTED, "2-D temporal blocking fuses %d launches (forms: cross, diamond, pyramid) or 2 (diamond, pyramid); even column counts", kTb2);
    if (!halo_src) return fail(LORA_ERR_ARG, "fused 2-D launches need halo_src (the buffer holding the caller's halo)");
    if (lo == hi) return LORA_OK;
    if (int rc = check_device(p)) return rc;
    const CUtensorMap *tm;
    int rc = get_tmap(p, src, &tm);
    if (rc) return rc;
    SegCut sc;
    if ((rc = cut_segments(lo, hi, dst, ex, mirror_base, sc))) return rc;
    Geom2DTB g{};
    g.out = dst;
    g.halo_src = halo_src;
    g.pitch = p->padded[1];
    g.m = (int)p->dims[0];
    g.n = (int)p->dims[1];
    const int wout = strip_out_cols_2d_tb(tb);
    g.nstrips = (g.n + wout - 1) / wout;
    long long chunk[kMaxSegs], tasks[kMaxSegs];
    for (int i = 0; i < sc.n; i++) {
        const long long rows = sc.hi[i] - sc.lo[i];
        if (i == sc.n - 1) {  // the main segment (the only one of a plain launch): the planner of plan_tasks_2dtb
            g.row_lo = (int)sc.lo[i];
            g.row_hi = (int)sc.hi[i];
            plan_main_2dtb(g, rows, p->sm_count, tb);
            chunk[i] = g.rows_per_chunk;
            tasks[i] = g.ntasks;
        } else {  // a band: short tasks, every strip
            chunk[i] = rows < kEdgeRows2Tb ? rows : kEdgeRows2Tb;
            tasks[i] = ((rows + chunk[i] - 1) / chunk[i]) * g.nstrips;
        }
    }
    fill_segs(g.sg, sc, chunk, tasks, 1);
    g.ntasks = (int)g.sg.first[sc.n];
    g.par0 = p->boundary == LORA_BOUNDARY_REFERENCE ? (launches_before & 1) : (p->boundary == LORA_BOUNDARY_ZERO ? 1 : 0);
    g.par_mask = p->boundary == LORA_BOUNDARY_REFERENCE ? 1 : 0;
    g.virt_top = virt_lo ? 1 : 0;
    g.virt_bot = virt_hi ? 1 : 0;
    g.vec4 = (g.n % 4 == 0) && (reinterpret_cast<uintptr_t>(dst) % 32 == 0) && sc.aligned4();
    cudaError_t e = launch_2d_tb(p->form, tb, *tm, g, p->w2, p->wd, static_cast<cudaStream_t>(stream));
    if (e != cudaSuccess) return fail(LORA_ERR_CUDA, "kernel launch failed: %s", cudaGetErrorString(e));
    p->launches++;
    return LORA_OK;
}

extern "C" int lora_plan_temporal_block(const lora_plan_t *p) { return p ? p->max_tb : 0; }

static int step_fused_impl(lora_plan_t *p, const double *src, double *dst, const double *halo_src, long long lo,
                           long long hi, int tb, int launches_before, int virt_lo, int virt_hi, const lora_exchange *ex,
                           const double *mirror_base, void *stream);

extern "C" int lora_plan_step_fused(lora_plan_t *p, const double *src, double *dst, const double *halo_src,
                                    long long lo, long long hi, int tb, int launches_before, int virt_lo, int virt_hi,
                                    void *stream) {
    return step_fused_impl(p, src, dst, halo_src, lo, hi, tb, launches_before, virt_lo, virt_hi, nullptr, nullptr, stream);
}

extern "C" int lora_plan_step_fused_mirror(lora_plan_t *p, const double *src, double *dst, const double *halo_src,
                                           long long lo, long long hi, int tb, int launches_before, int virt_lo,
                                           int virt_hi, const double *mirror_base, void *stream) {
    return step_fused_impl(p, src, dst, halo_src, lo, hi, tb, launches_before, virt_lo, virt_hi, nullptr, mirror_base, stream);
}

int lora_plan_step_exchange(lora_plan_t *p, const double *src, double *dst, const double *halo_src, long long lo,
                            long long hi, int tb, int launches_before, int virt_lo, int virt_hi, const lora_exchange *ex,
                            void *stream) {
    if (p && p->dim == 3) {
        if (tb == kTb3) return step_fused_3d(p, src, dst, lo, hi, stream, ex);
        if (tb != 1) return fail(LORA_ERR_UNSUPPORTED, "3-D launches advance one time step, fused sweeps two");
        return step_unfused(p, src, dst, lo, hi, ex, nullptr, stream);
    }
    return step_fused_impl(p, src, dst, halo_src, lo, hi, tb, launches_before, virt_lo, virt_hi, ex, nullptr, stream);
}

static int step_fused_impl(lora_plan_t *p, const double *src, double *dst, const double *halo_src, long long lo,
                           long long hi, int tb, int launches_before, int virt_lo, int virt_hi, const lora_exchange *ex,
                           const double *mirror_base, void *stream) {
    if (!p || !src || !dst) return fail(LORA_ERR_ARG, "null argument");
    if (p->boundary == LORA_BOUNDARY_PERIODIC && tb > 1)
        return fail(LORA_ERR_UNSUPPORTED, "a periodic boundary refreshes the halo ring before every launch: no fused sweeps");
    if (p->dim == 2)
        return step_fused_2d(p, src, dst, halo_src, lo, hi, tb, launches_before, virt_lo, virt_hi, ex, mirror_base, stream);
    if (p->dim != 1) return fail(LORA_ERR_UNSUPPORTED, "this entry point fuses 1-D and 2-D launches; 3-D sweeps of two are issued by lora_plan_run and the slab drivers");
    if (tb < 1 || tb > kMaxTb1) return fail(LORA_ERR_ARG, "temporal block must be 1..%d", kMaxTb1);
    if (lo < 0 || hi > p->dims[0] || lo > hi) return fail(LORA_ERR_ARG, "bad range [%lld, %lld)", lo, hi);
    if ((virt_lo || virt_hi) && !halo_src) return fail(LORA_ERR_ARG, "virtual halo needs halo_src");
    if (lo == hi) return LORA_OK;
    if (int rc = check_device(p)) return rc;
    if (reinterpret_cast<uintptr_t>(src) % 16 || reinterpret_cast<uintptr_t>(dst) % 16)
        return fail(LORA_ERR_ARG, "buffers must be 16-byte aligned");
    const long long P = p->dims[0] + 8;  // padded length
    SegCut sc;
    if (int rc = cut_segments(4 + lo, 4 + hi, dst, ex, mirror_base, sc)) return rc;  // padded coordinates X
    Geom1DTB g{};
    g.in = src;
    g.out = dst;
    g.halo_src = halo_src ? halo_src : src;
    g.n = p->dims[0];
    long long max_rows = 128;  // rows per task: whole waves of the longest tasks below this (1 warm-up row each)
    if (const char *e = getenv("LORA_TB1_MAX_ROWS")) {  // tuning knob
        const long long v = atoll(e);
        if (v >= 8 && v <= 4096) max_rows = v;
    }
    long long chunk[kMaxSegs], tasks[kMaxSegs];
    for (int i = 0; i < sc.n; i++) {
        // output row r (level tb) covers padded cells [512 r - 4 tb, 512 r - 4 tb + 512)
        const long long rho0 = (sc.lo[i] + 4 * tb) / kTbRowCells;
        const long long nrows = (sc.hi[i] - 1 + 4 * tb) / kTbRowCells - rho0 + 1;
        chunk[i] = sc.band[i] ? nrows  // a band is one or two rows: one task
                              : pick_len(nrows, 1, (long long)p->sm_count * kTbCtasPerSm * kWarpsPerCta, max_rows, 8);
        tasks[i] = (nrows + chunk[i] - 1) / chunk[i];
    }
    fill_segs(g.sg, sc, chunk, tasks, 1);
    g.ntasks = g.sg.first[sc.n];
    g.tb = tb;
    g.par0 = p->boundary == LORA_BOUNDARY_REFERENCE ? (launches_before & 1) : (p->boundary == LORA_BOUNDARY_ZERO ? 1 : 0);
    g.par_mask = p->boundary == LORA_BOUNDARY_REFERENCE ? 1 : 0;
    g.virt_left = virt_lo ? 1 : 0;
    g.virt_right = virt_hi ? 1 : 0;
    g.out_off = (16 - (4 * tb) % 16) % 16;
    const long long in_rows = P / 16, out_rows = (P - g.out_off) / 16;
    g.use_tma = (in_rows >= 1 && out_rows >= 1) ? 1 : 0;
    g.xcov = g.use_tma ? in_rows * 16 : 0;
    g.out_rows = out_rows;
    CUtensorMap imap, omap;
    std::memset(&imap, 0, sizeof imap);
    std::memset(&omap, 0, sizeof omap);
    if (g.use_tma) {
        int rc = get_tmap1d(p, src, 0, in_rows, &imap);
        if (rc) return rc;
        rc = get_tmap1d(p, dst, g.out_off, out_rows, &omap);
        if (rc) return rc;
    }
    cudaError_t e = launch_1d_tb(imap, omap, g, p->w1, static_cast<cudaStream_t>(stream));
    if (e != cudaSuccess) return fail(LORA_ERR_CUDA, "kernel launch failed: %s", cudaGetErrorString(e));
    p->launches++;
    return LORA_OK;
}

// Temporal blocks for `times` launches: as many blocks of max_tb as fit, the remainder, and -- because the
// result has to land in buf[times % 2] like the reference's ping-pong (S3) -- one block split in two when
// the number of fused launches would have the wrong parity.
static std::vector<int> temporal_schedule(int times, int max_tb) {
    std::vector<int> tbs;
    for (int left = times; left > 0;) {
        const int t = left < max_tb ? left : max_tb;
        tbs.push_back(t);
        left -= t;
    }
    if ((tbs.size() & 1) != (size_t)(times & 1)) {
        for (size_t i = tbs.size(); i-- > 0;)
            if (tbs[i] >= 2) {
                const int a = tbs[i] / 2, b = tbs[i] - a;
                tbs[i] = b;
                tbs.insert(tbs.begin() + i, a);
                break;
            }
    }
    return tbs;
}

// `times` launches of a 1-D plan as fused sweeps: launch k reads buf[k%2]; the halo every level sees on a virtual
// side is the caller's halo (it lives in buf0, whose halo nothing writes) at even times and zero at odd times
static int run_fused_1d(lora_plan *p, double *buf0, double *buf1, int times, int virt_lo, int virt_hi, void *stream) {
    double *buf[2] = {buf0, buf1};
    const std::vector<int> tbs = temporal_schedule(times, p->max_tb);
    int done = 0;
    for (size_t k = 0; k < tbs.size(); k++) {
        int rc = lora_plan_step_fused(p, buf[k % 2], buf[(k + 1) % 2], buf0, 0, p->dims[0], tbs[k], done, virt_lo, virt_hi,
                                      stream);
        if (rc) return rc;
        done += tbs[k];
    }
    return LORA_OK;
}

// ---- boundary modes (new).  The reference knows one behaviour: its kernels write the
// interior only, buffer A starts as the caller's padded array and buffer B as zeros, so the halo a launch sees
// alternates caller's / zero (S2).  DIRICHLET keeps the caller's halo values fixed for every launch, ZERO keeps a zero
// halo: both ping-pong buffers then carry the same ring, and the fused kernels' virtual halo stops alternating.
extern "C" int lora_plan_set_boundary(lora_plan_t *p, int mode) {
    if (!p || (mode != LORA_BOUNDARY_REFERENCE && mode != LORA_BOUNDARY_DIRICHLET && mode != LORA_BOUNDARY_ZERO &&
               mode != LORA_BOUNDARY_PERIODIC))
        return fail(LORA_ERR_ARG, "bad boundary mode");
    if (mode == LORA_BOUNDARY_PERIODIC) {  // the wrap reads `halo` interior cells behind each face
        for (int i = 0; i < p->dim; i++)
            if (p->dims[i] < p->halo[i])
                return fail(LORA_ERR_UNSUPPORTED, "a periodic boundary needs at least %d cells along axis %d (the storage halo), got %lld",
                            p->halo[i], i, p->dims[i]);
    }
    p->boundary = mode;
    return LORA_OK;
}
extern "C" int lora_plan_boundary(const lora_plan_t *p) { return p ? p->boundary : -1; }

// halo ring of dst <- halo ring of src (src == nullptr: zeros), everything outside the interior, nothing inside;
// restricted to padded indices [r0, r1) of the outermost axis (the bands of run_host_pipelined)
static int copy_ring_rows(const lora_plan *p, double *dst, const double *src, long long r0, long long r1, cudaStream_t st,
                          bool lead = true, bool trail = true) {
    const int dim = p->dim;
    const long long P0 = p->padded[0], rest = p->elems / P0, h0 = p->halo[0];
    r0 = std::max(r0, 0LL);
    r1 = std::min(r1, P0);
    if (r0 >= r1) return LORA_OK;
    auto flat = [&](long long off, long long cnt) -> cudaError_t {
        if (cnt <= 0) return cudaSuccess;
        return src ? cudaMemcpyAsync(dst + off, src + off, (size_t)cnt * 8, cudaMemcpyDeviceToDevice, st)
                   : cudaMemsetAsync(dst + off, 0, (size_t)cnt * 8, st);
    };
    auto strided = [&](long long off, long long width, long long height, long long pitch) -> cudaError_t {
        if (width <= 0 || height <= 0) return cudaSuccess;
        return src ? cudaMemcpy2DAsync(dst + off, (size_t)pitch * 8, src + off, (size_t)pitch * 8, (size_t)width * 8, (size_t)height,
                                       cudaMemcpyDeviceToDevice, st)
                   : cudaMemset2DAsync(dst + off, (size_t)pitch * 8, 0, (size_t)width * 8, (size_t)height, st);
    };
    // leading / trailing halo rows / planes (1-D: cells) inside the range (a slab that faces a neighbour keeps ghost
    // rows there instead: not part of the ring)
    if (lead) CU_TRY(flat(r0 * rest, (std::min(r1, h0) - r0) * rest));
    if (trail) CU_TRY(flat(std::max(r0, P0 - h0) * rest, (r1 - std::max(r0, P0 - h0)) * rest));
    const long long base = r0 * rest, cnt = r1 - r0;
    if (dim == 2) {
        CU_TRY(strided(base, 4, cnt, p->padded[1]));                      // left halo columns of every row
        CU_TRY(strided(base + p->padded[1] - 4, 4, cnt, p->padded[1]));   // right halo columns
    } else if (dim == 3) {
        const long long pitch = p->padded[2], plane = p->padded[1] * pitch;
        CU_TRY(strided(base, 2 * pitch, cnt, plane));                          // 2 leading halo rows of every plane
        CU_TRY(strided(base + plane - 2 * pitch, 2 * pitch, cnt, plane));      // 2 trailing halo rows
        CU_TRY(strided(base, 4, cnt * p->padded[1], pitch));                   // left halo columns of every row of every plane
        CU_TRY(strided(base + pitch - 4, 4, cnt * p->padded[1], pitch));       // right halo columns
    }
    return LORA_OK;
}
static int copy_ring(const lora_plan *p, double *dst, const double *src, cudaStream_t st) {
    return copy_ring_rows(p, dst, src, 0, p->padded[0], st);
}
// LORA_BOUNDARY_PERIODIC: halo ring of buf <- the periodic image of buf's interior (boundary.cu), innermost axis first
static int wrap_ring(const lora_plan *p, double *buf, cudaStream_t st) {
    long long inner = 1;
    for (int ax = p->dim - 1; ax >= 0; ax--) {
        const long long line = p->padded[ax] * inner;  // doubles in one padded line of this axis
        CU_TRY(launch_wrap_axis(buf, p->elems / line, p->dims[ax], p->halo[ax], inner, p->sm_count, st));
        inner = line;
    }
    return LORA_OK;
}
extern "C" int lora_plan_wrap_ring(lora_plan_t *p, double *buf, void *stream) {
    if (!p || !buf) return fail(LORA_ERR_ARG, "null argument");
    if (int rc = check_device(p)) return rc;
    for (int i = 0; i < p->dim; i++)
        if (p->dims[i] < p->halo[i]) return fail(LORA_ERR_UNSUPPORTED, "grid thinner than its storage halo along axis %d", i);
    return wrap_ring(p, buf, static_cast<cudaStream_t>(stream));
}

// the axis order and geometry of wrap_ring on a HOST array of the padded size (interior sizes dims[0..dim)): what the CPU
// tests compare with numpy's wrap padding -- no CUDA call
extern "C" int lora_debug_wrap_ring_host(int dim, const long long *dims, double *buf) {
    if (dim < 1 || dim > 3 || !dims || !buf) return fail(LORA_ERR_ARG, "bad argument");
    static const int halo[4][3] = {{0, 0, 0}, {4, 0, 0}, {4, 4, 0}, {1, 2, 4}};
    long long padded[3], elems = 1;
    for (int i = 0; i < dim; i++) {
        if (dims[i] < halo[dim][i]) return fail(LORA_ERR_UNSUPPORTED, "grid thinner than its storage halo along axis %d", i);
        padded[i] = dims[i] + 2 * halo[dim][i];
        elems *= padded[i];
    }
    long long inner = 1;
    for (int ax = dim - 1; ax >= 0; ax--) {
        const long long line = padded[ax] * inner;
        wrap_axis_host(buf, elems / line, dims[ax], halo[dim][ax], inner);
        inner = line;
    }
    return LORA_OK;
}

// the launch geometry of the radius-2 3-D kernels for planes [0, h) of an h x m x n grid (stencil3d_r2.cu), for the CPU
// tests: out4 = {grid.x, grid.y, plane chunks, planes per chunk}; no CUDA call
extern "C" int lora_debug_r2_grid(int form, int variant, long long h, int m, int n, int sm_count, long long *out4) {
    if (!out4 || h <= 0 || m <= 0 || n <= 0 || sm_count <= 0 || variant < 0 || variant > 2) return fail(LORA_ERR_ARG, "bad argument");
    if (form != LORA_FORM_STAR13 && form != LORA_FORM_HSEP5 && form != LORA_FORM_DIRECT125 && form != LORA_FORM_SEP5)
        return fail(LORA_ERR_ARG, "not a radius-2 form: %d", form);
    const int cols = r2_cols_per_cta(form, variant), rows = r2_rows_per_cta();
    out4[0] = (n + cols - 1) / cols;
    out4[1] = (m + rows - 1) / rows;
    out4[3] = r2_planes_per_chunk(h, out4[0] * out4[1], sm_count);
    out4[2] = (h + out4[3] - 1) / out4[3];
    return LORA_OK;
}

// for the slab driver (exchange.h): the ring of a slab's local array -- the side halo of every local row / plane, and the
// leading / trailing halo rows only where the slab ends the grid
int lora_plan_copy_ring(lora_plan_t *p, double *dst, const double *src, int lead, int trail, void *stream) {
    if (!p || !dst) return fail(LORA_ERR_ARG, "null argument");
    if (int rc = check_device(p)) return rc;
    return copy_ring_rows(p, dst, src, 0, p->padded[0], static_cast<cudaStream_t>(stream), lead != 0, trail != 0);
}

// 3-D fused launch of kTb3 = 2 time steps over interior planes [lo, hi): see stencil3d_tb.cu.  The source buffer's
// halo ring must hold what an EVEN time calls for (the caller's halo): lora_plan_run arranges that.
static int step_fused_3d(lora_plan *p, const double *src, double *dst, long long lo, long long hi, void *stream,
                         const lora_exchange *ex) {
    if (!p || !src || !dst) return fail(LORA_ERR_ARG, "null argument");
    if (p->dim != 3 || !tb3_form(p->form) || p->odd_cols)
        return fail(LORA_ERR_UNSUPPORTED, "3-D temporal blocking fuses %d launches of the 7-point and separable forms (even column counts)", kTb3);
    if (p->boundary == LORA_BOUNDARY_DIRICHLET)
        return fail(LORA_ERR_UNSUPPORTED, "fused 3-D sweeps keep a zero halo at the intermediate level: not with a Dirichlet boundary");
    if (lo < 0 || hi > p->dims[0] || lo > hi) return fail(LORA_ERR_ARG, "bad range [%lld, %lld)", lo, hi);
    if (lo == hi) return LORA_OK;
    if (int rc = check_device(p)) return rc;
    if (reinterpret_cast<uintptr_t>(dst) % 16) return fail(LORA_ERR_ARG, "destination buffer must be 16-byte aligned");
    const CUtensorMap *tm;
    if (int rc = get_tmap3tb(p, src, &tm)) return rc;
    Geom3DTB g{};
    g.out = dst;
    g.row_pitch = p->padded[2];
    g.plane_pitch = p->padded[1] * p->padded[2];
    g.h = (int)p->dims[0];
    g.m = (int)p->dims[1];
    g.n = (int)p->dims[2];
    g.tiles_m = (g.m + kT3OutRows - 1) / kT3OutRows;
    g.tiles_n = (g.n + kT3OutCols - 1) / kT3OutCols;
    long long max_planes = 96;
    if (const char *e = getenv("LORA_MAX_PLANES_3DTB")) {  // tuning knob
        const long long v = atoll(e);
        if (v >= 2 && v <= 4096) max_planes = v;
    }
    const long long tiles = (long long)g.tiles_m * g.tiles_n;
    const long long bl = ex && ex->band_lo > 0 ? ex->band_lo : 0, bh = ex && ex->band_hi > 0 ? ex->band_hi : 0;
    long long L = pick_chunk_3d(hi - lo, tiles, p->slots, max_planes, 6);
    long long nchunks = (hi - lo + L - 1) / L;
    if ((bl || bh) && nchunks <= 2) {  // a slab with bands needs a first and a last chunk
        L = (hi - lo + 1) / 2;
        nchunks = 2;
    }
    long long last_lo = lo + (nchunks - 1) * L;
    if ((bl || bh) && nchunks >= 3 && hi - last_lo < bh) {  // a short tail: the last chunk absorbs it (its own chunk length)
        nchunks--;
        last_lo = lo + (nchunks - 1) * L;
    }
    SegCut sc;
    long long chunk[kMaxSegs], tasks[kMaxSegs];
    if (bl || bh) {
        // A slab that faces neighbours: its bands (2 planes = radius x 2 launches) are FOLDED into the ordinary plane
        // chunks as in the unfused launch (step_unfused): the last chunk goes first in dispatch order -- its last
        // planes are the hi band, mirrored as they are stored, flag raised when its CTAs finish -- the first chunk next
        // -- its first planes are the lo band, flag raised when ITS CTAs finish -- then the chunks in between
        if ((bl && !ex->mirror_lo) || (bh && !ex->mirror_hi)) return fail(LORA_ERR_ARG, "exchange band without a mirror address");
        if (bl > L || bh > hi - last_lo)
            return fail(LORA_ERR_UNSUPPORTED, "fused 3-D slab of %lld planes is too thin for its bands: use fewer GPUs", hi - lo);
        sc.seq = ex->seq;
        sc.add(last_lo, hi, bh ? (long long)(ex->mirror_hi - dst) : 0, false, bh ? ex->flag_hi : nullptr,
               bh ? ex->count_hi : nullptr, bh ? ex->arrived_hi : nullptr);
        sc.mlo[0] = hi - bh, sc.mhi[0] = hi;
        sc.add(lo, lo + L, bl ? (long long)(ex->mirror_lo - dst) : 0, false, bl ? ex->flag_lo : nullptr,
               bl ? ex->count_lo : nullptr, bl ? ex->arrived_lo : nullptr);
        sc.mlo[1] = lo, sc.mhi[1] = lo + bl;  // (no early flag in the fused kernel: stencil3d_tb.cu)
        if (nchunks > 2) sc.add(lo + L, last_lo, 0, false, nullptr, nullptr, nullptr);
    } else {
        sc.add(lo, hi, 0, false, nullptr, nullptr, nullptr);
    }
    for (int i = 0; i < sc.n; i++) {
        chunk[i] = (bl || bh) && i == 0 ? hi - last_lo : L;  // the hi-band chunk is one task, whatever its length
        tasks[i] = (sc.hi[i] - sc.lo[i] + chunk[i] - 1) / chunk[i];
    }
    fill_segs(g.sg, sc, chunk, tasks, tiles);
    g.vec4 = (g.n % 4 == 0) && (reinterpret_cast<uintptr_t>(dst) % 32 == 0);
    cudaError_t e = launch_3d_tb(p->form, *tm, g, p->w3, static_cast<cudaStream_t>(stream));
    if (e != cudaSuccess) return fail(LORA_ERR_CUDA, "kernel launch failed: %s", cudaGetErrorString(e));
    p->launches++;
    return LORA_OK;
}

// Fused or not?  The fused 2-D kernel runs at the edge of the register file (255 registers, 8 warps per SM), and
// whether 3 launches per sweep beat 3 single launches has differed between boxes for the diamond form (387 vs 337
// GStencil/s on one, 322 vs 330 on another).  So the first lora_plan_run of a large cross / diamond plan measures both
// on a scratch grid of the same width (up to 3072 rows), keeps the winner and caches the verdict per (form, columns,
// rows, device) for the life of the process.  Results are bit-identical either way.
struct Tb2Key {
    int form, device;
    long long n, rows;
    bool operator<(const Tb2Key &o) const {
        return std::tie(form, device, n, rows) < std::tie(o.form, o.device, o.n, o.rows);
    }
};
static std::mutex g_tb2_mutex;
static std::map<Tb2Key, int> g_tb2_cache;
static double g_tb2_last_ms[2] = {0, 0};  // what the last probe measured: {3 unfused launches, 1 fused sweep}

extern "C" int lora_debug_tb2_probe(double *ms_unfused3, double *ms_fused) {
    if (ms_unfused3) *ms_unfused3 = g_tb2_last_ms[0];
    if (ms_fused) *ms_fused = g_tb2_last_ms[1];
    return (int)g_tb2_cache.size();
}

static void probe_tb2(lora_plan *p) {
    p->tb_auto = false;
    const long long rows = p->dims[0] < 3072 ? p->dims[0] : 3072;
    const Tb2Key key{p->form, p->device, p->dims[1], rows};
    std::lock_guard<std::mutex> lk(g_tb2_mutex);
    auto it = g_tb2_cache.find(key);
    if (it != g_tb2_cache.end()) {
        p->max_tb = it->second;
        return;
    }
    lora_plan q = *p;  // same weights and form, fewer rows, no cached tensor maps
    q.maps.clear();
    q.maps1d.clear();
    q.dims[0] = rows;
    q.padded[0] = rows + 8;
    q.elems = q.padded[0] * q.padded[1];
    q.launches = 0;
    double *b[2] = {nullptr, nullptr};
    cudaStream_t st = nullptr;
    cudaEvent_t ev[2] = {nullptr, nullptr};
    bool ok = cudaMalloc(&b[0], (size_t)q.elems * 8) == cudaSuccess && cudaMalloc(&b[1], (size_t)q.elems * 8) == cudaSuccess &&
              cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking) == cudaSuccess && cudaEventCreate(&ev[0]) == cudaSuccess &&
              cudaEventCreate(&ev[1]) == cudaSuccess && cudaMemsetAsync(b[0], 0, (size_t)q.elems * 8, st) == cudaSuccess &&
              cudaMemsetAsync(b[1], 0, (size_t)q.elems * 8, st) == cudaSuccess;
    float best[2] = {1e30f, 1e30f};
    for (int rep = 0; ok && rep < 4; rep++) {  // rep 0 warms up
        for (int fused = 0; ok && fused < 2; fused++) {
            ok = cudaEventRecord(ev[0], st) == cudaSuccess;
            if (fused) {
                ok = ok && step_fused_2d(&q, b[0], b[1], b[0], 0, rows, kTb2, 0, 1, 1, nullptr, nullptr, st) == LORA_OK;
            } else {
                for (int i = 0; ok && i < kTb2; i++)
                    ok = step_unfused(&q, b[i % 2], b[(i + 1) % 2], 0, rows, nullptr, nullptr, st) == LORA_OK;
            }
            ok = ok && cudaEventRecord(ev[1], st) == cudaSuccess && cudaEventSynchronize(ev[1]) == cudaSuccess;
            float ms = 0;
            ok = ok && cudaEventElapsedTime(&ms, ev[0], ev[1]) == cudaSuccess;
            if (ok && rep > 0 && ms < best[fused]) best[fused] = ms;
        }
    }
    if (ev[0]) cudaEventDestroy(ev[0]);
    if (ev[1]) cudaEventDestroy(ev[1]);
    if (st) cudaStreamDestroy(st);
    if (b[0]) cudaFree(b[0]);
    if (b[1]) cudaFree(b[1]);
    if (!ok) {  // out of memory for the scratch grid, ...: keep the form's default, do not cache
        cudaGetLastError();
        return;
    }
    g_tb2_last_ms[0] = best[0];
    g_tb2_last_ms[1] = best[1];
    p->max_tb = best[1] < 0.98f * best[0] ? kTb2 : 1;
    g_tb2_cache[key] = p->max_tb;
}

extern "C" int lora_plan_run(lora_plan_t *p, double *buf0, double *buf1, int times, void *stream) {
    if (!p || !buf0 || !buf1) return fail(LORA_ERR_ARG, "null argument");
    if (p->tb_auto && times >= kTb2 && p->boundary != LORA_BOUNDARY_PERIODIC) {
        if (int rc = check_device(p)) return rc;
        probe_tb2(p);
    }
    if (p->boundary == LORA_BOUNDARY_PERIODIC) {
        // one launch per time step, the source's ring refreshed from its interior before each; the result buffer's ring
        // is refreshed once more at the end, so that what comes back is a consistent periodic array
        if (int rc = check_device(p)) return rc;
        cudaStream_t st = static_cast<cudaStream_t>(stream);
        double *pb[2] = {buf0, buf1};
        for (int i = 0; i < times; i++) {
            if (int rc = wrap_ring(p, pb[i % 2], st)) return rc;
            if (int rc = lora_plan_step(p, pb[i % 2], pb[(i + 1) % 2], 0, p->dims[0], stream)) return rc;
        }
        return wrap_ring(p, pb[times % 2], st);
    }
    if (p->boundary != LORA_BOUNDARY_REFERENCE) {  // both buffers carry the same ring: the caller's, or zeros
        if (int rc = check_device(p)) return rc;
        cudaStream_t st = static_cast<cudaStream_t>(stream);
        if (p->boundary == LORA_BOUNDARY_ZERO)
            if (int rc = copy_ring(p, buf0, nullptr, st)) return rc;
        if (int rc = copy_ring(p, buf1, p->boundary == LORA_BOUNDARY_ZERO ? nullptr : buf0, st)) return rc;
    }
    if (p->dim == 1 && p->max_tb > 1 && times > 1) return run_fused_1d(p, buf0, buf1, times, 1, 1, stream);
    double *buf[2] = {buf0, buf1};
    if (p->dim == 2 && p->max_tb == kTb2 && times >= kTb2) {
        // sweeps of 3 launches, then the remainder one by one: every sweep advances an odd number of time steps,
        // so sweep k reads buf[k%2] at a time of parity k%2 and the result lands in buf[times%2] (S3)
        int k = 0;
        for (int left = times; left > 0; k++) {
            const int tb = left >= kTb2 ? kTb2 : 1;
            int rc = step_fused_2d(p, buf[k % 2], buf[(k + 1) % 2], buf0, 0, p->dims[0], tb, times - left, 1, 1, nullptr, nullptr, stream);
            if (rc) return rc;
            left -= tb;
        }
        return LORA_OK;
    }
    if (p->dim == 2 && p->max_tb == 2 && times >= 4) {
        // Sweeps of 2 launches, an EVEN number of them (so that the data is back in buffer 0 and the remaining 0..3
        // single launches see the rings they expect, S3).  Every sweep starts at an even time, so its level 0 needs
        // the caller's halo around its source: the ring of buffer 0 is copied into buffer 1 for the duration and
        // cleared again afterwards (as for the fused 3-D sweeps below); the intermediate level's halo is virtual.
        int a = times / 2;
        a -= a % 2;
        cudaStream_t st = static_cast<cudaStream_t>(stream);
        const bool ref_mode = p->boundary == LORA_BOUNDARY_REFERENCE;  // otherwise both rings are the same already
        if (ref_mode)
            if (int rc = copy_ring(p, buf1, buf0, st)) return rc;
        for (int k = 0; k < a; k++)
            if (int rc = step_fused_2d(p, buf[k % 2], buf[(k + 1) % 2], buf0, 0, p->dims[0], 2, 2 * k, 1, 1, nullptr, nullptr, stream)) return rc;
        if (ref_mode)
            if (int rc = copy_ring(p, buf1, nullptr, st)) return rc;
        for (int i = 2 * a; i < times; i++)
            if (int rc = lora_plan_step(p, buf[i % 2], buf[(i + 1) % 2], 0, p->dims[0], stream)) return rc;
        return LORA_OK;
    }
    if (p->dim == 3 && p->max_tb == kTb3 && times >= 2 * kTb3 && p->boundary != LORA_BOUNDARY_DIRICHLET) {
        // Sweeps of 2 launches, an EVEN number of them, then the remaining 1..3 launches one by one.  A fused sweep
        // starts at an even time and needs the caller's halo around its source: sweep k reads buf[k % 2], and buffer
        // 1's ring holds zeros (S2) -- so the ring of buffer 0 is copied into buffer 1 before the first odd sweep and
        // cleared again after the last one.  With an even number of sweeps the data is back in buffer 0, the single
        // launches then see the rings they expect, and the result lands in buf[times % 2] (S3).
        int a = times / kTb3;
        a -= a % 2;
        cudaStream_t st = static_cast<cudaStream_t>(stream);
        const bool zero_mode = p->boundary == LORA_BOUNDARY_ZERO;  // both rings are zero anyway
        if (!zero_mode)
            if (int rc = copy_ring(p, buf1, buf0, st)) return rc;
        for (int k = 0; k < a; k++)
            if (int rc = step_fused_3d(p, buf[k % 2], buf[(k + 1) % 2], 0, p->dims[0], stream, nullptr)) return rc;
        if (!zero_mode)
            if (int rc = copy_ring(p, buf1, nullptr, st)) return rc;
        for (int i = a * kTb3; i < times; i++)
            if (int rc = lora_plan_step(p, buf[i % 2], buf[(i + 1) % 2], 0, p->dims[0], stream)) return rc;
        return LORA_OK;
    }
    for (int i = 0; i < times; i++) {
        int rc = lora_plan_step(p, buf[i % 2], buf[(i + 1) % 2], 0, p->dims[0], stream);
        if (rc) return rc;
    }
    return LORA_OK;
}

// ---------------------------------------------------------------------------------------------
// host-side planning, exposed for the CPU tests (no GPU needed)
// ---------------------------------------------------------------------------------------------
// sweeps of two launches (2-D diamond / pyramid, 3-D): an even number of pairs, then the remaining 0..3 launches one by one
static std::vector<int> pair_schedule(int times) {
    std::vector<int> tbs;
    int a = times >= 4 ? times / 2 : 0;
    a -= a % 2;
    tbs.assign(a, 2);
    tbs.insert(tbs.end(), times - 2 * a, 1);
    return tbs;
}
extern "C" int lora_debug_pair_schedule(int times, int *out, int cap) {
    if (times < 0) return -1;
    const std::vector<int> tbs = pair_schedule(times);
    for (size_t i = 0; i < tbs.size() && (int)i < cap; i++) out[i] = tbs[i];
    return (int)tbs.size();
}

extern "C" int lora_debug_temporal_schedule(int times, int max_tb, int *out, int cap) {
    const std::vector<int> tbs = temporal_schedule(times, max_tb);
    for (size_t i = 0; i < tbs.size() && (int)i < cap; i++) out[i] = tbs[i];
    return (int)tbs.size();
}

// the main-segment task plan of a fused 2-D launch of `tb` (3 or 2) launches, as step_fused_2d makes it
static void plan_main_2dtb(Geom2DTB &g, long long rows, int sm_count, int tb) {
    // resident warps: 3 CTAs per SM for sweeps of two, 2 for sweeps of three (stencil2d_tb.cu); the two-launch sweeps
    // want two waves of shorter tasks (10240^2: pyramid 425 -> 458, diamond 564 -> 589 GStencil/s), the three-launch
    // sweeps one -- as long as the tasks stay long against their 12 warm-up rows (a band of the drop-in operators,
    // ~2000 rows, is better off with one wave of 100-row tasks than with two of 50)
    plan_tasks_2dtb(g, rows, (long long)sm_count * (tb == 2 ? 3 : 2) * kWarpsPerCta, tb == 2 ? 2 : 1);
    if (tb == 2 && g.rows_per_chunk < 192) plan_tasks_2dtb(g, rows, (long long)sm_count * 3 * kWarpsPerCta, 1);
}

static int debug_tasks_2dtb(int m, int n, int lo, int hi, int sm_count, int tb, int *out3, int cap) {
    if (m <= 0 || n <= 0 || lo < 0 || hi > m || lo >= hi || sm_count <= 0 || (tb != 2 && tb != kTb2)) return -1;
    Geom2DTB g{};
    g.m = m;
    g.n = n;
    g.row_lo = lo;
    g.row_hi = hi;
    const int wout = strip_out_cols_2d_tb(tb);
    g.nstrips = (n + wout - 1) / wout;
    plan_main_2dtb(g, hi - lo, sm_count, tb);
    g.sg.nseg = 1;
    g.sg.lo[0] = lo;
    g.sg.hi[0] = hi;
    g.sg.chunk[0] = g.rows_per_chunk;
    g.sg.first[0] = 0;
    g.sg.first[1] = g.ntasks;
    for (int t = 0; t < g.ntasks && t < cap; t++) {
        int strip = -1, r0 = 0, R = 0, seg = 0;
        if (!decode_task_2dtb(g, t, strip, r0, R, seg)) R = 0;
        out3[3 * t] = strip;
        out3[3 * t + 1] = r0;
        out3[3 * t + 2] = R;
    }
    return g.ntasks;
}
extern "C" int lora_debug_tasks_2dtb(int m, int n, int lo, int hi, int sm_count, int *out3, int cap) {
    return debug_tasks_2dtb(m, n, lo, hi, sm_count, kTb2, out3, cap);
}
extern "C" int lora_debug_tasks_2dtb_pairs(int m, int n, int lo, int hi, int sm_count, int *out3, int cap) {
    return debug_tasks_2dtb(m, n, lo, hi, sm_count, 2, out3, cap);
}

// ---------------------------------------------------------------------------------------------
// layer 3
// ---------------------------------------------------------------------------------------------
extern "C" int lora_decompose_2d(int shape, int mode, const double *params49, lora_decomp2d_t *out) {
    if (!params49 || !out) return fail(LORA_ERR_ARG, "null argument");
    Decomp2D d;
    if (!decompose_2d(shape, mode, params49, d)) return fail(LORA_ERR_ARG, "shape %d is not 2-D", shape);
    out->form = d.form;
    out->nterms = d.nterms;
    std::memcpy(out->vert, d.vert, sizeof d.vert);
    std::memcpy(out->horiz, d.horiz, sizeof d.horiz);
    out->centre = d.centre;
    std::memcpy(out->residual, d.residual, sizeof d.residual);
    out->recon_err = d.recon_err;
    out->macs_per_cell = d.macs;
    return LORA_OK;
}

extern "C" int lora_decompose_3d_r2(int shape, const double *params125, lora_decomp3d_r2_t *out) {
    if (!params125 || !out) return fail(LORA_ERR_ARG, "null argument");
    Decomp3DR2 d;
    if (!decompose_3d_r2(shape, params125, d)) return fail(LORA_ERR_ARG, "shape %d is not a radius-2 3-D shape", shape);
    out->form = d.form;
    std::memcpy(out->a, d.a, sizeof d.a);
    std::memcpy(out->b, d.b, sizeof d.b);
    std::memcpy(out->c, d.c, sizeof d.c);
    std::memcpy(out->q, d.q, sizeof d.q);
    out->recon_err = d.recon_err;
    out->macs_per_cell = d.macs;
    return LORA_OK;
}

extern "C" int lora_reference_table(int shape, double *table_out) {
    if (!table_out || (shape_dim(shape) == 0 && !shape_is_r2(shape))) return fail(LORA_ERR_ARG, "bad argument");
    reference_table(shape, table_out);
    return LORA_OK;
}

extern "C" int lora_effective_weights(int shape, int mode, const double *params, double *w) {
    if (!w) return fail(LORA_ERR_ARG, "null argument");
    double table[125];
    if (!params) {
        if (shape_dim(shape) == 0 && !shape_is_r2(shape)) return fail(LORA_ERR_ARG, "unknown shape %d", shape);
        reference_table(shape, table);
        params = table;
    }
    if (shape_is_r2(shape)) {
        Decomp3DR2 d;
        decompose_3d_r2(shape, params, d);
        std::memcpy(w, d.w, sizeof d.w);
        return LORA_OK;
    }
    switch (shape_dim(shape)) {
        case 1: {
            Decomp1D d;
            decompose_1d(shape, mode, params, d);
            std::memcpy(w, d.w, sizeof d.w);
            return LORA_OK;
        }
        case 2: {
            Decomp2D d;
            decompose_2d(shape, mode, params, d);
            std::memcpy(w, d.effective, sizeof d.effective);
            return LORA_OK;
        }
        case 3: {
            Decomp3D d;
            decompose_3d(shape, mode, params, d);
            std::memcpy(w, d.effective, sizeof d.effective);
            return LORA_OK;
        }
        default:
            return fail(LORA_ERR_ARG, "unknown shape %d", shape);
    }
}

// ---------------------------------------------------------------------------------------------
// layer 1: drop-in host operators
// ---------------------------------------------------------------------------------------------
static int g_verbose = -1;
// what the last drop-in call ON THIS THREAD measured / chose: concurrent callers (they serialise on the workspace mutex)
// each read their own figures back
static thread_local double g_loop_ms = 0, g_total_ms = 0;
static thread_local int g_chunks = 1;  // chunks the last drop-in call was cut into (1-D copy/compute overlap)
static thread_local int g_bands = 1;   // time-skewed bands of the last call that took the band pipeline (run_host_pipelined)
static std::mutex g_ws_mutex;
static std::vector<double *> g_ws;  // device workspace the drop-in operators cache between calls (equal-sized buffers)
static size_t g_ws_bytes = 0;
static int g_ws_device = -1;

extern "C" int lora_set_verbose(int on) {
    if (g_verbose < 0) {
        const char *q = getenv("LORA_QUIET");
        g_verbose = (q && q[0] && q[0] != '0') ? 0 : 1;
    }
    int prev = g_verbose;
    g_verbose = on ? 1 : 0;
    return prev;
}
static bool verbose() {
    if (g_verbose < 0) {
        const char *q = getenv("LORA_QUIET");
        g_verbose = (q && q[0] && q[0] != '0') ? 0 : 1;
    }
    return g_verbose != 0;
}

extern "C" double lora_last_loop_ms(void) { return g_loop_ms; }
extern "C" double lora_last_total_ms(void) { return g_total_ms; }
extern "C" int lora_last_chunks(void) { return g_chunks; }
extern "C" int lora_last_bands(void) { return g_bands; }

static void ws_free_locked() {
    for (double *b : g_ws)
        if (b) cudaFree(b);
    g_ws.clear();
    g_ws_bytes = 0;
    g_ws_device = -1;
}

extern "C" void lora_release_workspace(void) {
    std::lock_guard<std::mutex> lk(g_ws_mutex);
    ws_free_locked();
}

// the reference's CUDA_CHECK: report and exit(1) (src/2d/2d_utils.h:22-36)
[[noreturn]] static void die_cuda(cudaError_t e, const char *what, int line) {
    printf("CUDA Error:\n");
    printf("    File:       %s\n", __FILE__);
    printf("    Line:       %d\n", line);
    printf("    Error code: %d\n", (int)e);
    printf("    Error text: %s (%s)\n", cudaGetErrorString(e), what);
    fflush(stdout);
    exit(1);
}
#define CU_DIE(call)                                          \
    do {                                                      \
        cudaError_t e__ = (call);                             \
        if (e__ != cudaSuccess) die_cuda(e__, #call, __LINE__); \
    } while (0)

[[noreturn]] static void die_plan(const char *what) {
    printf("LoRAStencil error: %s: %s\n", what, lora_last_error());
    fflush(stdout);
    exit(1);
}

// at least `count` cached device buffers of at least `bytes` each on the current device (caller holds g_ws_mutex)
static void ws_reserve(size_t count, size_t bytes) {
    int dev = 0;
    CU_DIE(cudaGetDevice(&dev));
    // start over on another device, when the cached buffers are too small, and when they are far larger than this
    // call needs while more of them are wanted (a 13 GB 2-D workspace must not be multiplied by a chunked 1-D call)
    if (g_ws_device != dev || g_ws_bytes < bytes || (g_ws.size() < count && g_ws_bytes > 2 * bytes + (64u << 20)))
        ws_free_locked();
    if (g_ws.empty()) {
        g_ws_bytes = bytes + bytes / 8;  // slack: a slightly larger follow-up call does not reallocate everything
        g_ws_device = dev;
    }
    while (g_ws.size() < count) {
        double *b = nullptr;
        CU_DIE(cudaMalloc(&b, g_ws_bytes));
        g_ws.push_back(b);
    }
}

// pinned staging + worker threads for pageable caller buffers (hostmove.h); created on first use, never destroyed
// (its threads must not outlive-or-race the CUDA runtime's own teardown at exit)
static HostMover &mover() { return global_mover(); }

static void print_banner(int shape, int dim, const long long *dims, int times, double loop_us) {
    double cells = 1;
    for (int i = 0; i < dim; i++) cells *= (double)dims[i];
    printf("LoRAStencil(%s): \n", shape_banner(shape));
    printf("Time = %lld[ms]\n", (long long)(loop_us / 1e3));
    printf("GStencil/s = %f\n", cells * times * shape_artifact_k(shape) / (loop_us / 1e6) / 1e9);
    fflush(stdout);
}

// 1-D host operator, chunked and copy-overlapped.  A cell's value after `times` launches depends on 4 * times
// cells either side only, so a long line is cut into K chunks that each carry a ghost margin of G = 4 * times
// cells and run ALL their launches independently (bit-identical: the same operations on the same operands).
// Chunk c+2's H2D copy, the launches of chunks c and c+1 and chunk c-1's D2H copy then overlap instead of
// H2D -> launches -> D2H running back to back.  The ghost margins cost 2 G / chunk of redundant work (0.02 % for
// 2^28 points x 1000 launches).  Sides that are ends of the line keep the reference's halo semantics (S2) through
// the kernel's virtual halo.  Returns false when chunking does not pay (short line, wide cone): the caller then
// takes the plain path.  LORA_CHUNKS=0 disables, LORA_CHUNKS=k forces k chunks.
static bool run_host_1d_chunked(int shape, int mode, const double *in, double *out, const double *params, int times,
                                long long n) {
    const long long G = 4LL * times;
    long long K = n / (16LL << 20);  // chunks of >= 16 M cells keep the sweeps efficient; 8-16 chunks measured best
    if (K > 12) K = 12;
    bool forced = false;
    if (const char *e = getenv("LORA_CHUNKS")) {
        K = atoll(e);
        forced = true;
    }
    if (times < 1 || K < 2 || K > n) return false;
    long long C = ((n + K - 1) / K + 511) / 512 * 512;  // chunk length, whole kernel rows
    if (!forced && G > C / 16) return false;
    // chunk boundaries.  Only the first chunk's H2D copy and the last chunk's D2H copy are not hidden behind launches,
    // so (unless a chunk count is forced) those two chunks are quarter-length.
    std::vector<long long> cut{0};
    if (!forced && K >= 4) {
        C = ((long long)(n / (K - 1.5)) + 511) / 512 * 512;
        const long long small = (C / 4 + 511) / 512 * 512;
        cut.push_back(small < n ? small : n);
        while (cut.back() < n) {
            const long long left = n - cut.back();
            cut.push_back(left <= C + small ? (left > small + 512 ? n - small : n) : cut.back() + C);
        }
    } else {
        while (cut.back() < n) cut.push_back(cut.back() + C < n ? cut.back() + C : n);
    }
    K = (long long)cut.size() - 1;
    if (K < 2) return false;

    // Two chunks compute at a time (two launch streams): the tail of one chunk's sweep, where SMs run dry, is
    // back-filled by the other chunk's CTAs.  Buffer pairs in flight: two computing, one loading, one draining.
    constexpr int NCS = 2, NP = NCS + 2;
    std::lock_guard<std::mutex> lk(g_ws_mutex);
    ws_reserve(2 * NP, (size_t)(C + 2 * G + 8) * sizeof(double));
    cudaStream_t s_in, s_out, s_comp[NCS];
    CU_DIE(cudaStreamCreateWithFlags(&s_in, cudaStreamNonBlocking));
    CU_DIE(cudaStreamCreateWithFlags(&s_out, cudaStreamNonBlocking));
    for (auto &st : s_comp) CU_DIE(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    std::vector<cudaEvent_t> in_done(K), comp_done(K), out_done(K), t0(K);
    std::vector<lora_plan *> plans(K, nullptr);
    for (long long c = 0; c < K; c++) {
        CU_DIE(cudaEventCreateWithFlags(&in_done[c], cudaEventDisableTiming));
        CU_DIE(cudaEventCreate(&comp_done[c]));  // timed: end of the chunk's launch loop
        CU_DIE(cudaEventCreateWithFlags(&out_done[c], cudaEventDisableTiming));
        CU_DIE(cudaEventCreate(&t0[c]));         // timed: start of the chunk's launch loop
    }
    for (long long c = 0; c < K; c++) {
        const long long lo = cut[c], hi = cut[c + 1];
        const long long gl = G < lo ? G : lo, gr = G < n - hi ? G : n - hi;
        const long long nloc = (hi + gr) - (lo - gl);        // interior length of the chunk's own padded array
        const int virt_lo = (lo - gl == 0), virt_hi = (hi + gr == n);
        double *b0 = g_ws[2 * (c % NP)], *b1 = g_ws[2 * (c % NP) + 1];
        const size_t bytes = (size_t)(nloc + 8) * sizeof(double);
        cudaStream_t sc = s_comp[c % NCS];
        if (c >= NP) CU_DIE(cudaStreamWaitEvent(s_in, out_done[c - NP], 0));  // the pair's previous tenant has drained
        // S2 per chunk: buffer 0 <- the padded input segment, buffer 1 <- zeros
        CU_DIE(mover().h2d(b0, in + (lo - gl), bytes, s_in));  // pageable `in`: staged through pinned slots by worker threads
        CU_DIE(cudaMemsetAsync(b1, 0, bytes, s_in));
        CU_DIE(cudaEventRecord(in_done[c], s_in));

        const long long d[1] = {nloc};
        if (lora_plan_create(&plans[c], shape, mode, params, d) != LORA_OK) die_plan("plan");
        CU_DIE(cudaStreamWaitEvent(sc, in_done[c], 0));
        CU_DIE(cudaEventRecord(t0[c], sc));
        if (run_fused_1d(plans[c], b0, b1, times, virt_lo, virt_hi, sc) != LORA_OK) die_plan("launch");
        CU_DIE(cudaEventRecord(comp_done[c], sc));

        // S3: the chunk's own cells of buffer times%2; the first / last chunk also return the line's halo cells
        // (the reference copies back n + 7 doubles, src/1d/gpu_1r.cu:134)
        const double *res = (times % 2 == 0) ? b0 : b1;
        long long src_off = 4 + gl, dst_off = 4 + lo, cnt = hi - lo;
        if (c == 0) {
            src_off -= 4;
            dst_off -= 4;
            cnt += 4;
        }
        if (c == K - 1) cnt += 3;
        CU_DIE(cudaStreamWaitEvent(s_out, comp_done[c], 0));
        CU_DIE(mover().d2h(out + dst_off, res + src_off, (size_t)cnt * sizeof(double), s_out));
        CU_DIE(cudaEventRecord(out_done[c], s_out));
    }
    CU_DIE(cudaStreamSynchronize(s_out));
    CU_DIE(mover().finish());  // pageable `out`: the last staged pieces have been copied out
    for (auto &st : s_comp) CU_DIE(cudaStreamSynchronize(st));
    CU_DIE(cudaStreamSynchronize(s_in));
    // The reference's timed region is its launch loop (src/1d/gpu_1r.cu:118-126).  Here the chunks' launch loops
    // overlap each other and the copies, so the equivalent is the time during which AT LEAST ONE chunk's launch
    // loop was running: the union of the intervals [t0[c], comp_done[c]] (waiting for a chunk's H2D is not in it).
    std::vector<std::pair<float, float>> iv(K);
    for (long long c = 0; c < K; c++) {
        CU_DIE(cudaEventElapsedTime(&iv[c].first, t0[0], t0[c]));
        CU_DIE(cudaEventElapsedTime(&iv[c].second, t0[0], comp_done[c]));
    }
    std::sort(iv.begin(), iv.end());
    float ms = 0, cur_lo = iv[0].first, cur_hi = iv[0].second;
    for (long long c = 1; c < K; c++) {
        if (iv[c].first > cur_hi) {
            ms += cur_hi - cur_lo;
            cur_lo = iv[c].first;
            cur_hi = iv[c].second;
        } else if (iv[c].second > cur_hi) {
            cur_hi = iv[c].second;
        }
    }
    ms += cur_hi - cur_lo;
    for (long long c = 0; c < K; c++) {
        lora_plan_destroy(plans[c]);
        cudaEventDestroy(in_done[c]);
        cudaEventDestroy(comp_done[c]);
        cudaEventDestroy(out_done[c]);
        cudaEventDestroy(t0[c]);
    }
    cudaStreamDestroy(s_in);
    cudaStreamDestroy(s_out);
    for (auto &st : s_comp) cudaStreamDestroy(st);
    g_loop_ms = ms;
    g_chunks = (int)K;
    return true;
}

// sweeps of TWO launches (2-D diamond / pyramid on request, 3-D): every sweep starts at an even time, so the ring of
// BOTH ping-pong buffers has to hold the caller's halo while they run (lora_plan_run, run_host_pipelined)
static bool pair_sweeps(const lora_plan *p) {
    return p->boundary == LORA_BOUNDARY_REFERENCE && ((p->dim == 2 && p->max_tb == 2) || (p->dim == 3 && p->max_tb == kTb3));
}

// the sweeps lora_plan_run would issue for `times` launches (temporal blocks; their count has the parity of `times`)
static std::vector<int> plan_schedule(const lora_plan *p, int times) {
    if (p->dim == 1 && p->max_tb > 1 && times > 1) return temporal_schedule(times, p->max_tb);
    std::vector<int> tbs;
    if (p->dim == 2 && p->max_tb == kTb2 && times >= kTb2) {
        for (int left = times; left > 0;) {
            const int tb = left >= kTb2 ? kTb2 : 1;
            tbs.push_back(tb);
            left -= tb;
        }
        return tbs;
    }
    if (pair_sweeps(p) && times >= 4) return pair_schedule(times);  // as lora_plan_run
    tbs.assign(times, 1);
    return tbs;
}

// one sweep of `tb` launches over interior range [lo, hi) of the outermost axis (sweep index k, `done` launches before it)
static int sweep_range(lora_plan *p, double *buf0, double *buf1, int k, int tb, int done, long long lo, long long hi, void *stream) {
    double *buf[2] = {buf0, buf1};
    if (lo >= hi) return LORA_OK;
    if (p->dim == 1 && (tb > 1 || p->max_tb > 1))
        return lora_plan_step_fused(p, buf[k % 2], buf[(k + 1) % 2], buf0, lo, hi, tb, done, 1, 1, stream);
    if (p->dim == 2 && tb > 1) return step_fused_2d(p, buf[k % 2], buf[(k + 1) % 2], buf0, lo, hi, tb, done, 1, 1, nullptr, nullptr, stream);
    if (p->dim == 3 && tb > 1) return step_fused_3d(p, buf[k % 2], buf[(k + 1) % 2], lo, hi, stream, nullptr);
    return lora_plan_step(p, buf[k % 2], buf[(k + 1) % 2], lo, hi, stream);
}

// The drop-in operator on one GPU with its copies OVERLAPPED with its launch loop (the reference does
// cudaMemcpy -> launch loop -> cudaMemcpy back to back, src/2d/gpu.cu:396-421).  The grid is cut into K bands along
// the outermost axis and the bands are TIME-SKEWED: sweep s of band k covers rows [B_k - s r, B_k+1 - s r), r = the
// widest reach of one sweep (radius x temporal block).  Band k then depends only on itself and on bands < k (what it
// reads at sweep s was written at sweep s-1 by band k, or by band k-1 further up), and what it overwrites in the
// ping-pong buffer is exactly what band k+1 no longer needs -- so every band can run ALL its sweeps as soon as it has
// been uploaded, on one stream, band after band: same launches on the same operands as the plain sequence (results
// bit-identical), not one redundant cell.  Band k+1 uploads while band k computes, band k-1 downloads; only the first
// band's upload and the last band's download are exposed.  (A band's rows also stay L2-warm from sweep to sweep.)
// Pageable caller buffers are staged through pinned memory by worker threads (hostmove.h).
static void run_host_pipelined(lora_plan *p, const double *in, double *out, int times) {
    const size_t bytes = (size_t)p->elems * sizeof(double);
    const long long rows = p->padded[0], rest = p->elems / p->padded[0];
    const long long h0 = (p->padded[0] - p->dims[0]) / 2, n0 = p->dims[0];
    std::lock_guard<std::mutex> lk(g_ws_mutex);
    ws_reserve(2, bytes);
    double *b0 = g_ws[0], *b1 = g_ws[1];
    if (p->tb_auto && times >= kTb2) probe_tb2(p);
    const std::vector<int> tbs = plan_schedule(p, times);
    const int S = (int)tbs.size();
    int npairs = 0;  // leading sweeps of two launches (pair_sweeps): they need the caller's ring in buffer 1 too
    if (pair_sweeps(p))
        for (int tb : tbs) npairs += tb == 2;
    long long rmax = 0;
    for (int tb : tbs) rmax = std::max(rmax, (long long)p->radius0 * tb);
    // few, large bands: every band is swept launch by launch, and launches over a few hundred rows fill the GPU badly
    // (10240^2 x 100 launches: 12 bands -> 45 ms of launches, the whole grid at once 28 ms); about 200 MB per band, at
    // most 6, leaves a quarter to a sixth of one copy exposed at either end
    long long K = bytes < (128u << 20) ? 1 : (long long)((bytes + (100u << 20)) / (200u << 20));
    K = K < 1 ? 1 : (K > 6 ? 6 : K);
    if (const char *e = getenv("LORA_BANDS")) {  // tuning knob; 1 = copy, loop, copy back to back
        const long long v = atoll(e);
        if (v >= 1 && v <= 64) K = v;
    }
    // Only the first band's upload and the last band's download are exposed: with three bands or more those two are
    // HALF bands (one more band in total, the inner ones keep their size) -- 10240^2 x 100 launches: 4 uniform bands
    // expose 2 x 3.8 ms of a 28 ms call, 5 bands of 1/8, 1/4, 1/4, 1/4, 1/8 expose 2 x 1.9 ms
    const bool half_ends = K >= 3 && !getenv("LORA_UNIFORM_BANDS");
    if (half_ends && !getenv("LORA_BANDS")) K++;
    // the skew moves every band boundary S x rmax rows: keep it below half the grid, and bands wider than two reaches
    if (S * rmax > n0 / 2) K = 1;
    auto narrowest = [&](long long k) { return (n0 - S * rmax) / (half_ends && k >= 3 ? 2 * (k - 1) : k); };
    while (K > 1 && narrowest(K) < 2 * rmax + 1) K--;
    // final (download) partition F_k (uniform, or with half bands at the ends); the initial boundaries sit S x rmax further down
    std::vector<long long> B(K + 1, 0);
    for (long long k = 1; k < K; k++) {
        const long long num = half_ends && K >= 3 ? 2 * k - 1 : k, den = half_ends && K >= 3 ? 2 * (K - 1) : K;
        B[k] = std::min(n0, (n0 - S * rmax) * num / den + S * rmax);
    }
    B[K] = n0;
    auto lo_of = [&](long long k, int s) { return k == 0 ? 0LL : std::max(0LL, B[k] - s * rmax); };
    auto hi_of = [&](long long k, int s) { return k == K - 1 ? n0 : std::max(0LL, B[k + 1] - s * rmax); };
    cudaStream_t s_in, s_comp, s_out;
    CU_DIE(cudaStreamCreateWithFlags(&s_in, cudaStreamNonBlocking));
    CU_DIE(cudaStreamCreateWithFlags(&s_comp, cudaStreamNonBlocking));
    CU_DIE(cudaStreamCreateWithFlags(&s_out, cudaStreamNonBlocking));
    std::vector<cudaEvent_t> in_done(K), comp_done(K), t0(K);
    for (long long k = 0; k < K; k++) {
        CU_DIE(cudaEventCreateWithFlags(&in_done[k], cudaEventDisableTiming));
        CU_DIE(cudaEventCreate(&comp_done[k]));  // timed: end of the band's launch loop
        CU_DIE(cudaEventCreate(&t0[k]));         // timed: its start
    }
    // S2: buffer 1 <- zeros (its interior is overwritten by the first sweep; its halo ring must read as zero)
    CU_DIE(cudaMemsetAsync(b1, 0, bytes, s_comp));
    auto upload = [&](long long k) {  // S2: buffer 0 <- the padded input, halo included (src/2d/gpu.cu:396-400)
        const long long u0 = k == 0 ? 0 : std::min(rows, B[k] + h0), u1 = k == K - 1 ? rows : std::min(rows, B[k + 1] + h0);
        if (u1 > u0) CU_DIE(mover().h2d(b0 + u0 * rest, in + u0 * rest, (size_t)(u1 - u0) * rest * sizeof(double), s_in));
        CU_DIE(cudaEventRecord(in_done[k], s_in));
    };
    const double *res = (S % 2 == 0) ? b0 : b1;
    auto download = [&](long long k) {  // S3: padded buffer times%2; 1-D leaves the last double alone (src/1d/gpu_1r.cu:134)
        long long r_lo = lo_of(k, S) + h0, r_hi = hi_of(k, S) + h0;
        if (k == 0) r_lo = 0;           // + the halo rows at either end of the grid
        if (k == K - 1) r_hi = rows;
        if (r_hi <= r_lo) return;
        size_t cnt = (size_t)(r_hi - r_lo) * rest;
        if (p->dim == 1 && k == K - 1) cnt -= 1;
        CU_DIE(cudaStreamWaitEvent(s_out, comp_done[k], 0));
        CU_DIE(mover().d2h(out + r_lo * rest, res + r_lo * rest, cnt * sizeof(double), s_out));
    };
    upload(0);
    for (long long k = 0; k < K; k++) {
        if (k + 1 < K) upload(k + 1);  // in flight while band k computes (a pageable source blocks this thread only for
                                       // the staging memcpy; band k-1's launches are still queued on the GPU meanwhile)
        CU_DIE(cudaStreamWaitEvent(s_comp, in_done[k], 0));
        CU_DIE(cudaEventRecord(t0[k], s_comp));
        if (npairs > 0) {  // the rows just uploaded: their ring into buffer 1 as well
            const long long u0 = k == 0 ? 0 : std::min(rows, B[k] + h0), u1 = k == K - 1 ? rows : std::min(rows, B[k + 1] + h0);
            if (copy_ring_rows(p, b1, b0, u0, u1, s_comp) != LORA_OK) die_plan("ring copy");
        }
        int launched = 0;
        for (int s = 1; s <= S; s++) {
            if (npairs > 0 && s == npairs + 1) {
                // Buffer 1's ring back to zeros before a single launch reads it (and for the download, S3), in the rows
                // no later band's pair sweeps read from buffer 1 any more: band k+1's last pair sweep reads from interior
                // row B[k+1] - (npairs + 1) rmax on, this band's single launches stay strictly below that
                const long long c0 = k == 0 ? 0 : B[k] - (npairs + 1) * rmax + h0;
                const long long c1 = k == K - 1 ? rows : B[k + 1] - (npairs + 1) * rmax + h0;
                if (copy_ring_rows(p, b1, nullptr, c0, c1, s_comp) != LORA_OK) die_plan("ring reset");
            }
            if (sweep_range(p, b0, b1, s - 1, tbs[s - 1], launched, lo_of(k, s), hi_of(k, s), s_comp) != LORA_OK) die_plan("launch");
            launched += tbs[s - 1];
        }
        CU_DIE(cudaEventRecord(comp_done[k], s_comp));
        if (k >= 1) download(k - 1);  // after band k's launches were queued: a wait for staging slots delays no launch
    }
    download(K - 1);
    CU_DIE(cudaStreamSynchronize(s_out));
    CU_DIE(cudaStreamSynchronize(s_comp));
    CU_DIE(cudaStreamSynchronize(s_in));
    CU_DIE(mover().finish());
    // the reference's timed region is its launch loop + sync (src/2d/gpu.cu:408-414): here the bands' launch loops,
    // which run one after the other on one stream -- waiting for a band's upload is not in it
    float total = 0;
    for (long long k = 0; k < K; k++) {
        float ms = 0;
        CU_DIE(cudaEventElapsedTime(&ms, t0[k], comp_done[k]));
        total += ms;
        cudaEventDestroy(in_done[k]);
        cudaEventDestroy(comp_done[k]);
        cudaEventDestroy(t0[k]);
    }
    g_loop_ms = total;
    g_bands = (int)K;
    cudaStreamDestroy(s_in);
    cudaStreamDestroy(s_comp);
    cudaStreamDestroy(s_out);
}

// how many GPUs the drop-in operators spread one call over: LORA_NGPU=k (default 1), or lora_set_gpus(k)
static int g_ngpu = -1;
extern "C" int lora_set_gpus(int k) {
    const int prev = g_ngpu;
    g_ngpu = k;
    return prev < 0 ? 1 : prev;
}
// LORA_DEVICES="0,1,2,3": an explicit device list for the slabs (a device may appear more than once: several slabs
// then share it, which is how a single-GPU box exercises the whole exchange protocol); otherwise devices 0..k-1
static int wanted_gpus(std::vector<int> &devices) {
    devices.clear();
    int have = 0;
    if (cudaGetDeviceCount(&have) != cudaSuccess || have < 1) return 1;
    if (const char *e = getenv("LORA_DEVICES")) {
        if (g_ngpu < 0 || g_ngpu > 1) {
            for (const char *q = e; *q;) {
                char *end = nullptr;
                const long v = strtol(q, &end, 10);
                if (end == q) break;
                if (v >= 0 && v < have) devices.push_back((int)v);
                q = (*end == ',') ? end + 1 : end;
                if (*end != ',' ) break;
            }
            if (devices.size() > 1) return (int)devices.size();
            devices.clear();
        }
    }
    int k = g_ngpu;
    if (k < 0) {
        const char *e = getenv("LORA_NGPU");
        k = e ? atoi(e) : 1;
    }
    if (k <= 1) return 1;
    k = k < have ? k : have;
    for (int i = 0; i < k; i++) devices.push_back(i);
    return k;
}
static thread_local int g_last_gpus = 1;
extern "C" int lora_last_gpus(void) { return g_last_gpus; }

// The drop-in operator on k GPUs of this process (LORA_NGPU=k): the padded host grid is cut into k slabs along its
// outermost axis (slab.cu), every device takes its rows (+ ghost rows) H2D, the launch loop runs on all devices with
// the ghost zones exchanged inside the kernels over NVLink peer memory, and the slabs come back into `out`.  Same
// buffer semantics (S2 / S3), same banner, same timed region (launch loop + sync) as the single-GPU path; results
// are bit-identical to it.  Returns false when the grid is too thin to be cut that many ways (the caller then runs
// on one GPU).
static bool run_host_multi_gpu(const std::vector<int> &devices, int shape, int mode, const double *in, double *out,
                               const double *params, int times, const long long *dims) {
    using clk = std::chrono::steady_clock;
    const int k = (int)devices.size();
    lora_slabset_t *set = nullptr;
    if (lora_slabset_create(&set, shape, mode, params, dims, k, devices.data()) != LORA_OK) {
        if (verbose()) fprintf(stderr, "LoRAStencil: %d GPUs not usable for this call (%s); running on one\n", k, lora_last_error());
        return false;
    }
    if (lora_slabset_load(set, in) != LORA_OK) die_plan("multi-GPU H2D");
    const clk::time_point t0 = clk::now();
    if (lora_slabset_run(set, times) != LORA_OK) die_plan("multi-GPU launch");
    if (lora_slabset_sync(set) != LORA_OK) die_plan("multi-GPU sync");
    const long long us = std::chrono::duration_cast<std::chrono::microseconds>(clk::now() - t0).count();
    g_loop_ms = us / 1e3;
    if (verbose()) print_banner(shape, shape_dim(shape), dims, times, (double)us);
    if (lora_slabset_store(set, out) != LORA_OK) die_plan("multi-GPU D2H");
    lora_slabset_destroy(set);
    g_last_gpus = k;
    return true;
}

extern "C" void lora_gpu_run_host(int shape, int mode, const double *in, double *out, const double *params, int times,
                                  const long long *dims) {
    using clk = std::chrono::steady_clock;
    const clk::time_point t_begin = clk::now();
    g_chunks = 1;
    g_bands = 1;
    g_last_gpus = 1;
    std::vector<int> devices;
    const int k = wanted_gpus(devices);
    if (k > 1 && shape_is_r2(shape)) {
        fail(LORA_ERR_UNSUPPORTED, "the radius-2 3-D shapes run on one GPU (LORA_NGPU / lora_set_gpus asked for %d)", k);
        die_plan("slabs");
    }
    if (k > 1 && shape_dim(shape) != 0 && dims && in && out && times >= 0 &&
        run_host_multi_gpu(devices, shape, mode, in, out, params, times, dims)) {
        g_total_ms = std::chrono::duration_cast<std::chrono::microseconds>(clk::now() - t_begin).count() / 1e3;
        return;
    }
    if (shape_dim(shape) == 1 && dims && dims[0] > 0 && in && out &&
        run_host_1d_chunked(shape, mode, in, out, params, times, dims[0])) {
        if (verbose()) print_banner(shape, 1, dims, times, g_loop_ms * 1e3);
        g_total_ms = std::chrono::duration_cast<std::chrono::microseconds>(clk::now() - t_begin).count() / 1e3;
        return;
    }
    lora_plan *p = nullptr;
    if (lora_plan_create(&p, shape, mode, params, dims) != LORA_OK) die_plan("plan");
    run_host_pipelined(p, in, out, times);
    if (verbose()) print_banner(shape, p->dim, p->dims, times, g_loop_ms * 1e3);
    lora_plan_destroy(p);
    g_total_ms = std::chrono::duration_cast<std::chrono::microseconds>(clk::now() - t_begin).count() / 1e3;
}

extern "C" void lora_gpu_1d1r(const double *in, double *out, const double *params, int times, int n) {
    const long long d[1] = {n};
    lora_gpu_run_host(LORA_1D1R, LORA_WEIGHTS_REFERENCE, in, out, params, times, d);
}
extern "C" void lora_gpu_1d2r(const double *in, double *out, const double *params, int times, int n) {
    const long long d[1] = {n};
    lora_gpu_run_host(LORA_1D2R, LORA_WEIGHTS_REFERENCE, in, out, params, times, d);
}
extern "C" void lora_gpu_star_2d1r(const double *in, double *out, const double *params, int times, int m, int n) {
    const long long d[2] = {m, n};
    lora_gpu_run_host(LORA_STAR2D1R, LORA_WEIGHTS_REFERENCE, in, out, params, times, d);
}
extern "C" void lora_gpu_star_2d3r(const double *in, double *out, const double *params, int times, int m, int n) {
    const long long d[2] = {m, n};
    lora_gpu_run_host(LORA_STAR2D3R, LORA_WEIGHTS_REFERENCE, in, out, params, times, d);
}
extern "C" void lora_gpu_box_2d3r(const double *in, double *out, const double *params, int times, int m, int n) {
    const long long d[2] = {m, n};
    lora_gpu_run_host(LORA_BOX2D3R, LORA_WEIGHTS_REFERENCE, in, out, params, times, d);
}
extern "C" void lora_gpu_box_3d1r(const double *in, double *out, const double *params, int times, int h, int m, int n) {
    const long long d[3] = {h, m, n};
    lora_gpu_run_host(LORA_BOX3D1R, LORA_WEIGHTS_REFERENCE, in, out, params, times, d);
}
extern "C" void lora_gpu_box_3d2r(const double *in, double *out, const double *params, int times, int h, int m, int n) {
    const long long d[3] = {h, m, n};
    lora_gpu_run_host(LORA_BOX3D2R, LORA_WEIGHTS_GENERAL, in, out, params, times, d);
}
extern "C" void lora_gpu_star_3d2r(const double *in, double *out, const double *params, int times, int h, int m, int n) {
    const long long d[3] = {h, m, n};
    lora_gpu_run_host(LORA_STAR3D2R, LORA_WEIGHTS_GENERAL, in, out, params, times, d);
}
extern "C" void lora_gpu_star_3d1r(const double *in, double *out, const double *params, int times, int h, int m, int n) {
    const long long d[3] = {h, m, n};
    lora_gpu_run_host(LORA_STAR3D1R, LORA_WEIGHTS_REFERENCE, in, out, params, times, d);
}

// the reference's own C++ symbols (include/lorastencil_dropin.hpp)
void gpu_1d1r(const double *__restrict__ in, double *__restrict__ out, const double *__restrict__ params,
              const int time, const int input_n) {
    lora_gpu_1d1r(in, out, params, time, input_n);
}
void gpu_1d2r(const double *__restrict__ in, double *__restrict__ out, const double *__restrict__ params,
              const int time, const int input_n) {
    lora_gpu_1d2r(in, out, params, time, input_n);
}
void gpu_star_2d1r(const double *__restrict__ in, double *__restrict__ out, const double *__restrict__ params,
                   const int times, const int input_m, const int input_n) {
    lora_gpu_star_2d1r(in, out, params, times, input_m, input_n);
}
void gpu_star_2d3r(const double *__restrict__ in, double *__restrict__ out, const double *__restrict__ params,
                   const int times, const int input_m, const int input_n) {
    lora_gpu_star_2d3r(in, out, params, times, input_m, input_n);
}
void gpu_box_2d3r(const double *__restrict__ in, double *__restrict__ out, const double *__restrict__ params,
                  const int times, const int input_m, const int input_n) {
    lora_gpu_box_2d3r(in, out, params, times, input_m, input_n);
}
void gpu_box_3d1r(const double *__restrict__ in, double *__restrict__ out, const double *__restrict__ params,
                  const int times, const int input_h, const int input_m, const int input_n) {
    lora_gpu_box_3d1r(in, out, params, times, input_h, input_m, input_n);
}
void gpu_star_3d1r(const double *__restrict__ in, double *__restrict__ out, const double *__restrict__ params,
                   const int times, const int input_h, const int input_m, const int input_n) {
    lora_gpu_star_3d1r(in, out, params, times, input_h, input_m, input_n);
}
