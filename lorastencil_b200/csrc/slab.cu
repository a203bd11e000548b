// slab.cu -- multi-GPU domain decomposition: the grid is cut along its OUTERMOST axis into contiguous slabs, one
// per GPU; neighbours exchange ghost zones once per sweep (= once per launch, or once per temporal block where
// launches are fused) WITHOUT a communication library.  New functionality: the reference is single-GPU (no
// cudaSetDevice / NCCL / MPI anywhere under src/); the loop being sharded is its launch loop
// (src/2d/gpu.cu:408-414, src/3d/gpu_box.cu:206-214, src/1d/gpu_2r.cu:118-126).
//
// One sweep of a slab is ONE kernel launch (plan.cu: lora_plan_step_exchange, kernels.h: Segs):
//   * the tasks of the two BANDS (the cells a neighbour's next sweep reads: ghost-zone width) are dispatched first;
//     every cell they store goes to this GPU's buffer and, a second time, straight into the neighbour's ghost zone
//     (peer memory over NVLink: same-process peer access, or CUDA IPC between one-process-per-GPU ranks);
//   * the last task of a band raises a 64-bit flag in the neighbour's memory (fence + st.release.sys);
//   * the rest of the launch -- the interior -- reads this GPU's own cells only and overlaps all of that;
//   * the next sweep's launch is held back in stream order until both neighbours' flags have reached the sweep
//     number (cuStreamWaitValue64): their bands have arrived here, and they have stopped reading the ghost zones
//     this sweep's bands are about to overwrite there (band tasks are the only readers of a ghost zone).
// Outer faces of the global grid keep the reference's halo semantics (S2) untouched.  Results are bit-identical
// to a single-GPU run: every cell sees the same operands in the same order.
//
// Two front ends share this code: lora_slab_* = one slab (one-process-per-GPU ranks connect through IPC handles,
// lorastencil_b200/slab.py does the rendezvous with torch.distributed), lora_slabset_* = all slabs of one grid on
// several devices of ONE process (what the drop-in host operators use under LORA_NGPU=k).
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include <cuda.h>
#include <cuda_runtime.h>

#include "../../include/lorastencil.h"
#include "decompose.h"
#include "exchange.h"
#include "hostmove.h"
#include "kernels.h"

using namespace lora;

#define SL_TRY(call)                                                                                         \
    do {                                                                                                     \
        cudaError_t e__ = (call);                                                                            \
        if (e__ != cudaSuccess)                                                                              \
            return lora_fail(LORA_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
    } while (0)

namespace {

constexpr int kHalo0[4] = {0, 4, 4, 1};  // storage halo on the outermost axis (S1)
constexpr int kRadius0[4] = {0, 4, 3, 1};  // stencil radius on that axis

struct Geo {
    int dim = 0, world = 1, rank = 0;
    long long dims[3] = {0, 0, 0};
    long long lo = 0, hi = 0, slab = 0;
    long long halo = 0, wl = 0, wr = 0, off = 0;  // cells stored left / right of the slab; first slab cell in plan coordinates
    long long local_dims[3] = {0, 0, 0}, local_padded[3] = {0, 0, 0};
    long long rest = 1;  // doubles per outermost index of the padded local array
    bool has_prev = false, has_next = false;
};

// contiguous balanced split in units of `align` cells; `ghost` = cells kept beyond the slab on a side that faces a
// neighbour (>= halo).  Same rule as lorastencil_b200.slab.SlabGeometry.
int make_geo(Geo &g, int dim, const long long *dims, int world, int rank, long long ghost) {
    if (dim < 1 || dim > 3 || world < 1 || rank < 0 || rank >= world) return lora_fail(LORA_ERR_ARG, "bad slab arguments");
    static const int halo[4][3] = {{0, 0, 0}, {4, 0, 0}, {4, 4, 0}, {1, 2, 4}};
    g.dim = dim, g.world = world, g.rank = rank;
    for (int i = 0; i < dim; i++) g.dims[i] = dims[i];
    g.halo = kHalo0[dim];
    const long long align = dim == 1 ? 16 : 1, n0 = dims[0];
    const long long units = (n0 + align - 1) / align;
    auto bound = [&](int r) { return std::min(n0, align * (units * r / world)); };
    g.lo = bound(rank), g.hi = bound(rank + 1), g.slab = g.hi - g.lo;
    g.has_prev = rank > 0, g.has_next = rank < world - 1;
    if (ghost < g.halo) ghost = g.halo;
    g.wl = g.has_prev ? ghost : g.halo;
    g.wr = g.has_next ? ghost : g.halo;
    if (g.slab < std::max(g.wl, g.wr) || (world > 1 && g.slab < (g.has_prev ? ghost : 0) + (g.has_next ? ghost : 0)))
        return lora_fail(LORA_ERR_ARG, "slab of %lld is thinner than its ghost zones: use fewer GPUs", g.slab);
    g.off = g.wl - g.halo;
    g.local_dims[0] = g.slab + g.off + (g.wr - g.halo);
    g.rest = 1;
    for (int i = 1; i < dim; i++) g.local_dims[i] = dims[i];
    for (int i = 0; i < dim; i++) {
        g.local_padded[i] = g.local_dims[i] + 2 * halo[dim][i];
        if (i) g.rest *= g.local_padded[i];
    }
    return LORA_OK;
}

typedef CUresult (*wait64_fn)(CUstream, CUdeviceptr, cuuint64_t, unsigned int);
wait64_fn wait64() {
    static wait64_fn fn = [] {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuStreamWaitValue64", &p, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            p = nullptr;
        return reinterpret_cast<wait64_fn>(p);
    }();
    return fn;
}

std::vector<int> schedule_for(int dim, int times, int max_tb) {
    std::vector<int> tbs;
    if (dim == 1) {
        int n = lora_debug_temporal_schedule(times, max_tb, nullptr, 0);
        tbs.resize(n);
        lora_debug_temporal_schedule(times, max_tb, tbs.data(), n);
    } else if (dim == 2 && max_tb >= 3) {
        for (int i = 0; i < times / 3; i++) tbs.push_back(3);
        for (int i = 0; i < times % 3; i++) tbs.push_back(1);
    } else if ((dim == 2 || dim == 3) && max_tb == 2 && times >= 4) {
        // sweeps of two launches, an even number of them (the data is back in buffer 0), the rest one by one; the ring
        // of buffer 1 holds the caller's halo while they run (pair_ring below; same rule as lora_plan_run)
        int n = lora_debug_pair_schedule(times, nullptr, 0);
        tbs.resize(n);
        lora_debug_pair_schedule(times, tbs.data(), n);
    } else {
        tbs.assign(times, 1);
    }
    return tbs;
}

}  // namespace

struct lora_slab {
    Geo g;
    Geo gprev, gnext;
    lora_plan_t *plan = nullptr;
    int device = 0;
    int max_tb = 1;
    long long ghost = 0;
    long long elems = 0;
    double *buf[2] = {nullptr, nullptr};
    unsigned long long *sync = nullptr;  // device: [0] flag written by prev, [1] flag written by next, [2..3] band arrival counters
    // neighbours' memory as seen from this device: [side 0 = prev, 1 = next]
    double *peer_buf[2][2] = {{nullptr, nullptr}, {nullptr, nullptr}};
    unsigned long long *peer_sync[2] = {nullptr, nullptr};
    bool ipc_opened[2] = {false, false};
    bool connected[2] = {false, false};
    unsigned long long arrived[2] = {0, 0};  // host totals of the two arrival counters
    unsigned long long seq = 0;              // sweeps issued: the flags only ever grow
    long long launch = 0, time = 0;          // result sits in buf[launch % 2]
};

// Sweeps of two launches start at even times: level 0 must see the caller's halo around BOTH buffers while they run.
// before != 0: ring of buffer 1 <- ring of buffer 0; else: ring of buffer 1 <- zeros (S2 again for the single launches
// that follow and for the result).  Only ring cells are touched -- the neighbours' mirror stores write interior
// columns of the ghost rows only -- so this needs no ordering against the neighbours.
static int pair_ring(lora_slab *s, int before, void *stream) {
    return lora_plan_copy_ring(s->plan, s->buf[1], before ? s->buf[0] : nullptr, !s->g.has_prev, !s->g.has_next, stream);
}
static int count_pairs(const std::vector<int> &tbs, int dim, int max_tb) {
    int a = 0;
    if ((dim == 2 || dim == 3) && max_tb == 2)
        for (int tb : tbs) a += tb == 2;
    return a;
}

extern "C" int lora_slab_geometry(int dim, const long long *global_dims, int world, int rank, long long ghost,
                                  long long *out8) {
    if (!global_dims || !out8) return lora_fail(LORA_ERR_ARG, "null argument");
    Geo g;
    int rc = make_geo(g, dim, global_dims, world, rank, ghost);
    if (rc) return rc;
    out8[0] = g.lo, out8[1] = g.hi, out8[2] = g.wl, out8[3] = g.wr, out8[4] = g.off;
    out8[5] = g.local_dims[0], out8[6] = g.local_dims[1], out8[7] = g.local_dims[2];
    return LORA_OK;
}

extern "C" void lora_slab_destroy(lora_slab_t *s) {
    if (!s) return;
    int cur = -1;
    cudaGetDevice(&cur);
    cudaSetDevice(s->device);
    cudaDeviceSynchronize();
    for (int side = 0; side < 2; side++)
        if (s->ipc_opened[side]) {
            cudaIpcCloseMemHandle(s->peer_buf[side][0]);
            cudaIpcCloseMemHandle(s->peer_buf[side][1]);
            cudaIpcCloseMemHandle(s->peer_sync[side]);
        }
    if (s->buf[0]) cudaFree(s->buf[0]);
    if (s->buf[1]) cudaFree(s->buf[1]);
    if (s->sync) cudaFree(s->sync);
    if (s->plan) lora_plan_destroy(s->plan);
    if (cur >= 0) cudaSetDevice(cur);
    delete s;
}

extern "C" int lora_slab_create(lora_slab_t **out, int shape, int mode, const double *params, const long long *global_dims,
                                int world, int rank, int temporal_block) {
    if (!out || !global_dims) return lora_fail(LORA_ERR_ARG, "null argument");
    const int dim = shape_dim(shape);
    if (dim == 0) return lora_fail(LORA_ERR_ARG, "unknown shape %d", shape);
    lora_slab *s = new lora_slab;
    cudaGetDevice(&s->device);
    // the deepest temporal block decides the ghost width, so it is fixed before the geometry: ask a throw-away plan
    // what this shape's kernels fuse by default (1-D: 15, 2-D cross / diamond: 3, else 1)
    {
        lora_plan_t *probe = nullptr;
        long long d[3] = {64, 64, 64};
        for (int i = 1; i < dim; i++) d[i] = global_dims[i];  // an odd column count rules fusion out: same columns as the grid
        int rc = lora_plan_create(&probe, shape, mode, params, d);
        if (rc) {
            delete s;
            return rc;
        }
        if (temporal_block > 0) lora_plan_set_temporal_block(probe, temporal_block);
        s->max_tb = lora_plan_temporal_block(probe);  // 1-D: up to 15, 2-D: 3 / 2 / 1, 3-D: 2 / 1
        lora_plan_destroy(probe);
    }
    if (dim >= 2 && s->max_tb == 2 && world > 1) {
        // sweeps of two launches want room for both bands in every slab (3-D: a first and a last plane chunk that hold
        // one band each); thinner slabs advance one launch per sweep (same rule in lorastencil_b200/slab.py)
        const long long min_slab = global_dims[0] / world, need = 2 * (long long)kRadius0[dim] * 2;
        if (min_slab < need) s->max_tb = 1;
    }
    s->ghost = s->max_tb > 1 ? (long long)kRadius0[dim] * s->max_tb : kHalo0[dim];
    int rc = make_geo(s->g, dim, global_dims, world, rank, s->ghost);
    if (!rc && rank > 0) rc = make_geo(s->gprev, dim, global_dims, world, rank - 1, s->ghost);
    if (!rc && rank < world - 1) rc = make_geo(s->gnext, dim, global_dims, world, rank + 1, s->ghost);
    if (!rc) rc = lora_plan_create(&s->plan, shape, mode, params, s->g.local_dims);
    if (!rc) rc = lora_plan_set_temporal_block(s->plan, s->max_tb);
    if (rc) {
        lora_slab_destroy(s);
        return rc;
    }
    s->elems = lora_plan_padded_elems(s->plan);
    cudaError_t e = cudaMalloc(&s->buf[0], (size_t)s->elems * 8);
    if (e == cudaSuccess) e = cudaMalloc(&s->buf[1], (size_t)s->elems * 8);
    if (e == cudaSuccess) e = cudaMalloc(&s->sync, 64);
    if (e == cudaSuccess) e = cudaMemset(s->buf[0], 0, (size_t)s->elems * 8);
    if (e == cudaSuccess) e = cudaMemset(s->buf[1], 0, (size_t)s->elems * 8);
    if (e == cudaSuccess) e = cudaMemset(s->sync, 0, 64);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
        lora_slab_destroy(s);
        return lora_fail(LORA_ERR_CUDA, "slab buffers (%lld doubles x 2): %s", s->elems, cudaGetErrorString(e));
    }
    *out = s;
    return LORA_OK;
}

/* info10: lo, hi, wl, wr, off, local_padded[0..2], max_tb, ghost */
extern "C" int lora_slab_info(const lora_slab_t *s, long long *info10) {
    if (!s || !info10) return lora_fail(LORA_ERR_ARG, "null argument");
    info10[0] = s->g.lo, info10[1] = s->g.hi, info10[2] = s->g.wl, info10[3] = s->g.wr, info10[4] = s->g.off;
    for (int i = 0; i < 3; i++) info10[5 + i] = s->g.local_padded[i];
    info10[8] = s->max_tb, info10[9] = s->ghost;
    return LORA_OK;
}

extern "C" double *lora_slab_buffer(lora_slab_t *s, int which) { return (s && (which == 0 || which == 1)) ? s->buf[which] : nullptr; }
extern "C" int lora_slab_result_index(const lora_slab_t *s) { return s ? (int)(s->launch % 2) : 0; }
extern "C" long long lora_slab_launch_count(const lora_slab_t *s) { return s ? lora_plan_launch_count(s->plan) : 0; }
extern "C" lora_plan_t *lora_slab_plan(lora_slab_t *s) { return s ? s->plan : nullptr; }

extern "C" int lora_slab_export(lora_slab_t *s, void *handles192) {
    if (!s || !handles192) return lora_fail(LORA_ERR_ARG, "null argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    void *ptrs[3] = {s->buf[0], s->buf[1], s->sync};
    for (int i = 0; i < 3; i++) {
        cudaIpcMemHandle_t h;
        SL_TRY(cudaIpcGetMemHandle(&h, ptrs[i]));
        std::memcpy(static_cast<char *>(handles192) + 64 * i, &h, 64);
    }
    return LORA_OK;
}

extern "C" int lora_slab_connect_ipc(lora_slab_t *s, int side, const void *handles192) {
    if (!s || !handles192 || (side != 0 && side != 1)) return lora_fail(LORA_ERR_ARG, "bad argument");
    if ((side == 0 && !s->g.has_prev) || (side == 1 && !s->g.has_next)) return lora_fail(LORA_ERR_ARG, "no neighbour on that side");
    void *ptrs[3] = {nullptr, nullptr, nullptr};
    for (int i = 0; i < 3; i++) {
        cudaIpcMemHandle_t h;
        std::memcpy(&h, static_cast<const char *>(handles192) + 64 * i, 64);
        cudaError_t e = cudaIpcOpenMemHandle(&ptrs[i], h, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) {
            for (int j = 0; j < i; j++) cudaIpcCloseMemHandle(ptrs[j]);
            return lora_fail(LORA_ERR_CUDA, "cudaIpcOpenMemHandle: %s", cudaGetErrorString(e));
        }
    }
    s->peer_buf[side][0] = static_cast<double *>(ptrs[0]);
    s->peer_buf[side][1] = static_cast<double *>(ptrs[1]);
    s->peer_sync[side] = static_cast<unsigned long long *>(ptrs[2]);
    s->ipc_opened[side] = s->connected[side] = true;
    return LORA_OK;
}

extern "C" int lora_slab_connect_local(lora_slab_t *s, int side, lora_slab_t *nb) {
    if (!s || !nb || (side != 0 && side != 1)) return lora_fail(LORA_ERR_ARG, "bad argument");
    if (nb->device != s->device) {
        int can = 0;
        SL_TRY(cudaDeviceCanAccessPeer(&can, s->device, nb->device));
        if (!can) return lora_fail(LORA_ERR_UNSUPPORTED, "device %d cannot access device %d's memory", s->device, nb->device);
        int cur = -1;
        cudaGetDevice(&cur);
        SL_TRY(cudaSetDevice(s->device));
        cudaError_t e = cudaDeviceEnablePeerAccess(nb->device, 0);
        if (e == cudaErrorPeerAccessAlreadyEnabled) {
            cudaGetLastError();
            e = cudaSuccess;
        }
        if (cur >= 0) cudaSetDevice(cur);
        if (e != cudaSuccess) return lora_fail(LORA_ERR_CUDA, "cudaDeviceEnablePeerAccess: %s", cudaGetErrorString(e));
    }
    s->peer_buf[side][0] = nb->buf[0];
    s->peer_buf[side][1] = nb->buf[1];
    s->peer_sync[side] = nb->sync;
    s->connected[side] = true;
    return LORA_OK;
}

/* the buffers were (re)filled from outside: the next sweep is launch 0 at time 0 again (flags keep growing) */
extern "C" int lora_slab_reset(lora_slab_t *s) {
    if (!s) return lora_fail(LORA_ERR_ARG, "null argument");
    s->launch = s->time = 0;
    return LORA_OK;
}

extern "C" int lora_slab_sweep(lora_slab_t *s, int tb, void *stream) {
    if (!s || tb < 1) return lora_fail(LORA_ERR_ARG, "bad argument");
    const Geo &g = s->g;
    if ((g.has_prev && !s->connected[0]) || (g.has_next && !s->connected[1]))
        return lora_fail(LORA_ERR_ARG, "slab %d of %d is not connected to its neighbours", g.rank, g.world);
    if (tb > s->max_tb) return lora_fail(LORA_ERR_ARG, "temporal block %d exceeds the slab's ghost zones (max %d)", tb, s->max_tb);
    const int w = (int)((s->launch + 1) % 2);
    const double *src = s->buf[s->launch % 2];
    double *dst = s->buf[w];
    CUstream st = static_cast<CUstream>(stream);
    if (s->seq > 0 && (g.has_prev || g.has_next)) {
        wait64_fn fn = wait64();
        if (!fn) return lora_fail(LORA_ERR_UNSUPPORTED, "cuStreamWaitValue64 is not available");
        for (int side = 0; side < 2; side++)
            if (side == 0 ? g.has_prev : g.has_next)
                if (fn(st, reinterpret_cast<CUdeviceptr>(s->sync + side), s->seq, CU_STREAM_WAIT_VALUE_GEQ) != CUDA_SUCCESS)
                    return lora_fail(LORA_ERR_CUDA, "cuStreamWaitValue64 failed");
    }
    lora_exchange ex;
    ex.seq = s->seq + 1;
    if (g.has_prev) {  // my first slab cells -> prev's trailing ghost zone; I am prev's `next`: its flag [1]
        ex.band_lo = s->ghost;
        ex.mirror_lo = s->peer_buf[0][w] + ((s->gprev.wl + s->gprev.slab) - g.wl) * g.rest;
        ex.flag_lo = s->peer_sync[0] + 1;
        ex.count_lo = s->sync + 2;
        ex.arrived_lo = &s->arrived[0];
    }
    if (g.has_next) {  // my last slab cells -> next's leading ghost zone; I am next's `prev`: its flag [0]
        ex.band_hi = s->ghost;
        ex.mirror_hi = s->peer_buf[1][w] + (s->gnext.wl - g.wl - g.slab) * g.rest;
        ex.flag_hi = s->peer_sync[1] + 0;
        ex.count_hi = s->sync + 3;
        ex.arrived_hi = &s->arrived[1];
    }
    int rc = lora_plan_step_exchange(s->plan, src, dst, s->buf[0], g.off, g.off + g.slab, tb, (int)s->time, !g.has_prev,
                                     !g.has_next, &ex, stream);
    if (rc) return rc;
    s->seq++;
    s->launch++;
    s->time += tb;
    return LORA_OK;
}

extern "C" int lora_slab_schedule(const lora_slab_t *s, int times, int *tbs, int cap) {
    if (!s || times < 0) return -1;
    const std::vector<int> v = schedule_for(s->g.dim, times, s->max_tb);
    for (size_t i = 0; i < v.size() && (int)i < cap; i++) tbs[i] = v[i];
    return (int)v.size();
}

/* `times` launches of the reference operator on this slab; asynchronous on `stream`; the result is in
 * buffer lora_slab_result_index() -- buf[times % 2] from a fresh state, like the reference's ping-pong (S3) */
extern "C" int lora_slab_run(lora_slab_t *s, int times, void *stream) {
    if (!s || times < 0) return lora_fail(LORA_ERR_ARG, "bad argument");
    if (s->g.world == 1 && s->launch % 2 == 0 && s->time % 2 == 0) {  // whole grid on one device: the plan schedules it
        int rc = lora_plan_run(s->plan, s->buf[0], s->buf[1], times, stream);
        if (rc) return rc;
        s->launch += times;
        s->time += times;
        return LORA_OK;
    }
    if (s->max_tb > 1 && s->launch % 2 != s->time % 2)
        return lora_fail(LORA_ERR_ARG, "fused runs must start from a parity-consistent state");
    const std::vector<int> tbs = schedule_for(s->g.dim, times, s->max_tb);
    const int pairs = count_pairs(tbs, s->g.dim, s->max_tb);
    if (pairs > 0)
        if (int rc = pair_ring(s, 1, stream)) return rc;
    for (size_t k = 0; k < tbs.size(); k++) {
        if (pairs > 0 && (int)k == pairs)
            if (int rc = pair_ring(s, 0, stream)) return rc;
        int rc = lora_slab_sweep(s, tbs[k], stream);
        if (rc) return rc;
    }
    if (pairs > 0 && (int)tbs.size() == pairs)
        if (int rc = pair_ring(s, 0, stream)) return rc;
    return LORA_OK;
}

// ---------------------------------------------------------------------------------------------
// all slabs of one grid on several devices of one process
// ---------------------------------------------------------------------------------------------
struct lora_slabset {
    int dim = 0, ndev = 0;
    long long dims[3] = {0, 0, 0}, padded[3] = {0, 0, 0};
    std::vector<int> devices;
    std::vector<lora_slab *> slabs;
    std::vector<cudaStream_t> streams;
    int shape = 0;
};

extern "C" void lora_slabset_destroy(lora_slabset_t *set) {
    if (!set) return;
    int cur = -1;
    cudaGetDevice(&cur);
    for (size_t i = 0; i < set->slabs.size(); i++) {
        cudaSetDevice(set->devices[i]);
        cudaDeviceSynchronize();
    }
    for (size_t i = 0; i < set->slabs.size(); i++) {
        cudaSetDevice(set->devices[i]);
        if (i < set->streams.size() && set->streams[i]) cudaStreamDestroy(set->streams[i]);
        lora_slab_destroy(set->slabs[i]);
    }
    if (cur >= 0) cudaSetDevice(cur);
    delete set;
}

extern "C" int lora_slabset_create(lora_slabset_t **out, int shape, int mode, const double *params,
                                   const long long *global_dims, int ndev, const int *devices) {
    if (!out || !global_dims || ndev < 1) return lora_fail(LORA_ERR_ARG, "bad argument");
    const int dim = shape_dim(shape);  // 0 for the radius-2 3-D shapes too: one GPU only (stencil3d_r2.cu)
    if (dim == 0) return lora_fail(LORA_ERR_ARG, "unknown shape %d (slabs cut the reference's eight shapes)", shape);
    int have = 0;
    SL_TRY(cudaGetDeviceCount(&have));
    static const int halo[4][3] = {{0, 0, 0}, {4, 0, 0}, {4, 4, 0}, {1, 2, 4}};
    lora_slabset *set = new lora_slabset;
    set->dim = dim, set->ndev = ndev, set->shape = shape;
    for (int i = 0; i < dim; i++) set->dims[i] = global_dims[i], set->padded[i] = global_dims[i] + 2 * halo[dim][i];
    int cur = -1;
    cudaGetDevice(&cur);
    int rc = LORA_OK;
    for (int r = 0; r < ndev && !rc; r++) {
        const int dev = devices ? devices[r] : r;
        if (dev < 0 || dev >= have) {
            rc = lora_fail(LORA_ERR_ARG, "device %d requested, %d visible", dev, have);
            break;
        }
        set->devices.push_back(dev);
        cudaError_t e = cudaSetDevice(dev);
        if (e != cudaSuccess) {
            rc = lora_fail(LORA_ERR_CUDA, "cudaSetDevice(%d): %s", dev, cudaGetErrorString(e));
            break;
        }
        lora_slab *s = nullptr;
        rc = lora_slab_create(&s, shape, mode, params, global_dims, ndev, r, 0);
        if (rc) break;
        set->slabs.push_back(s);
        cudaStream_t st = nullptr;
        e = cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking);
        set->streams.push_back(st);
        if (e != cudaSuccess) rc = lora_fail(LORA_ERR_CUDA, "cudaStreamCreate: %s", cudaGetErrorString(e));
    }
    for (int r = 0; r < ndev && !rc; r++) {
        if (r > 0) rc = lora_slab_connect_local(set->slabs[r], 0, set->slabs[r - 1]);
        if (!rc && r < ndev - 1) rc = lora_slab_connect_local(set->slabs[r], 1, set->slabs[r + 1]);
    }
    if (cur >= 0) cudaSetDevice(cur);
    if (rc) {
        lora_slabset_destroy(set);
        return rc;
    }
    *out = set;
    return LORA_OK;
}

/* S2 per slab: buffer 0 <- the slab's rows of the padded host array (halo / ghost rows included), buffer 1 <- zeros */
extern "C" int lora_slabset_load(lora_slabset_t *set, const double *in) {
    if (!set || !in) return lora_fail(LORA_ERR_ARG, "null argument");
    int cur = -1;
    cudaGetDevice(&cur);
    for (int r = 0; r < set->ndev; r++) {
        lora_slab *s = set->slabs[r];
        SL_TRY(cudaSetDevice(set->devices[r]));
        const long long row0 = s->g.lo + s->g.halo - s->g.wl;  // first global padded row this slab mirrors
        SL_TRY(global_mover().h2d(s->buf[0], in + row0 * s->g.rest, (size_t)s->elems * 8, set->streams[r]));  // pageable: staged
        SL_TRY(cudaMemsetAsync(s->buf[1], 0, (size_t)s->elems * 8, set->streams[r]));
        lora_slab_reset(s);
    }
    // nobody may store into a neighbour's ghost rows before that neighbour has finished (re)filling its buffers
    for (int r = 0; r < set->ndev; r++) {
        SL_TRY(cudaSetDevice(set->devices[r]));
        SL_TRY(cudaStreamSynchronize(set->streams[r]));
    }
    if (cur >= 0) cudaSetDevice(cur);
    return LORA_OK;
}

/* sweep-major issue order: every device has sweep k queued before any device gets sweep k + 1, so a full launch
 * queue on one device can never starve the neighbour whose flag it is waiting for */
extern "C" int lora_slabset_run(lora_slabset_t *set, int times) {
    if (!set || times < 0) return lora_fail(LORA_ERR_ARG, "bad argument");
    int cur = -1;
    cudaGetDevice(&cur);
    int rc = LORA_OK;
    if (set->ndev == 1) {
        cudaSetDevice(set->devices[0]);
        rc = lora_slab_run(set->slabs[0], times, set->streams[0]);
    } else {
        lora_slab *s0 = set->slabs[0];
        if (s0->max_tb > 1 && s0->launch % 2 != s0->time % 2)
            rc = lora_fail(LORA_ERR_ARG, "fused runs must start from a parity-consistent state");
        const std::vector<int> tbs = schedule_for(set->dim, times, s0->max_tb);
        const int pairs = count_pairs(tbs, set->dim, s0->max_tb);
        auto ring_all = [&](int before) {
            for (int r = 0; r < set->ndev && !rc; r++) {
                cudaSetDevice(set->devices[r]);
                rc = pair_ring(set->slabs[r], before, set->streams[r]);
            }
        };
        if (pairs > 0) ring_all(1);
        for (size_t k = 0; k < tbs.size() && !rc; k++) {
            if (pairs > 0 && (int)k == pairs) ring_all(0);
            for (int r = 0; r < set->ndev && !rc; r++) {
                cudaSetDevice(set->devices[r]);
                rc = lora_slab_sweep(set->slabs[r], tbs[k], set->streams[r]);
            }
        }
        if (pairs > 0 && (int)tbs.size() == pairs) ring_all(0);
    }
    if (cur >= 0) cudaSetDevice(cur);
    return rc;
}

extern "C" int lora_slabset_sync(lora_slabset_t *set) {
    if (!set) return lora_fail(LORA_ERR_ARG, "null argument");
    int cur = -1;
    cudaGetDevice(&cur);
    for (int r = 0; r < set->ndev; r++) {
        SL_TRY(cudaSetDevice(set->devices[r]));
        SL_TRY(cudaStreamSynchronize(set->streams[r]));
    }
    if (cur >= 0) cudaSetDevice(cur);
    return LORA_OK;
}

/* S3: the whole padded result: interior rows from their owners, the outer halo rows from the end slabs.  1-D copies
 * back one double less, like the reference (src/1d/gpu_1r.cu:134) */
extern "C" int lora_slabset_store(lora_slabset_t *set, double *out) {
    if (!set || !out) return lora_fail(LORA_ERR_ARG, "null argument");
    int cur = -1;
    cudaGetDevice(&cur);
    for (int r = 0; r < set->ndev; r++) {
        lora_slab *s = set->slabs[r];
        const Geo &g = s->g;
        SL_TRY(cudaSetDevice(set->devices[r]));
        const double *res = s->buf[s->launch % 2];
        long long src_row = g.wl, dst_row = g.lo + g.halo, rows = g.slab;
        if (r == 0) src_row -= g.halo, dst_row -= g.halo, rows += g.halo;
        if (r == set->ndev - 1) rows += g.halo;
        long long cnt = rows * g.rest;
        if (set->dim == 1 && r == set->ndev - 1) cnt -= 1;
        SL_TRY(global_mover().d2h(out + dst_row * g.rest, res + src_row * g.rest, (size_t)cnt * 8, set->streams[r]));
    }
    for (int r = 0; r < set->ndev; r++) {
        SL_TRY(cudaSetDevice(set->devices[r]));
        SL_TRY(cudaStreamSynchronize(set->streams[r]));
    }
    SL_TRY(global_mover().finish());  // pageable `out`: the last staged pieces have been copied out
    if (cur >= 0) cudaSetDevice(cur);
    return LORA_OK;
}

extern "C" long long lora_slabset_launch_count(const lora_slabset_t *set) {
    long long n = 0;
    if (set)
        for (lora_slab *s : set->slabs) n += lora_slab_launch_count(s);
    return n;
}
extern "C" int lora_slabset_temporal_block(const lora_slabset_t *set) { return (set && !set->slabs.empty()) ? set->slabs[0]->max_tb : 0; }
