// kernels.h -- host-visible launch interface of the sm_100a stencil kernels.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

namespace lora {

// ---- tiling constants shared by host planning and device code ----
constexpr int kWarpCols = 128;     // one warp owns 32 lanes x 4 consecutive columns
constexpr int kWarpsPerCta = 4;     // 1-D / 2-D: warps are independent workers, the CTA is only a container
constexpr int kBoxCols = 136;      // 128 + 4 halo columns each side (stencil radius <= 4, 32-byte aligned)
// per-warp TMA ring of the 1-D / 2-D kernels: 2 stages x 8 rows measured best (profiles/r1_ring_variants.log:
// against 3 x 4, unfused diamond 334 -> 367, unfused 1-D 369 -> 388, pyramid 335 -> 343 GStencil/s, fused cross equal)
#ifndef LORA_ROWS_PER_STAGE
#define LORA_ROWS_PER_STAGE 8
#endif
#ifndef LORA_STAGES
#define LORA_STAGES 2
#endif
constexpr int kRowsPerStage = LORA_ROWS_PER_STAGE;   // rows of one TMA box
constexpr int kStages = LORA_STAGES;                 // per-warp ring depth
constexpr int kStageElems = kRowsPerStage * kBoxCols;            // 1088 doubles = 8704 B (68 x 128 B)
constexpr int kSmem12 = kWarpsPerCta * kStages * kStageElems * 8 + kWarpsPerCta * kStages * 8;
// ---- temporally blocked 1-D kernel (stencil1d_tb.cu) ----
// Rows of 512 cells (lane = 16 cells = 128 B).  Per-warp shared memory: kTbStages x 4 KB TMA load ring +
// kTbOutBufs x 4 KB output staging rows (both 128B-swizzled, 1024-byte aligned), per-level mailboxes
// (kMaxTb1 levels x 2 row parities x 64 B), mbarriers; rounded to a multiple of 1024.  The defaults put 4 CTAs x
// 4 warps on an SM (4 x (56 KB + 1 KB reserved) = 228 KB); the macros exist for tuning experiments only.
#ifndef LORA_TB_STAGES
#define LORA_TB_STAGES 2
#endif
#ifndef LORA_TB_OUTBUFS
#define LORA_TB_OUTBUFS 1
#endif
#ifndef LORA_TB_CTAS
#define LORA_TB_CTAS 4
#endif
#ifndef LORA_TB_MAX
#define LORA_TB_MAX 15
#endif
constexpr int kMaxTb1 = LORA_TB_MAX;         // deepest temporal block of the 1-D kernel
constexpr int kDefaultTb1 = 15;              // what lora_plan_run fuses unless told otherwise (LORA_TB / set_temporal_block)
#ifndef LORA_TB_CPL
#define LORA_TB_CPL 16
#endif
constexpr int kTbCellsPerLane = LORA_TB_CPL;  // 16 or 32 (a multiple of 16: whole 128-byte swizzle rows per lane)
constexpr int kTbLaneRows = kTbCellsPerLane / 16;  // rows of 16 doubles (= 128 B) a lane owns
static_assert(kTbCellsPerLane % 16 == 0, "a lane owns whole rows of the tensor map");
constexpr int kTbRowCells = 32 * kTbCellsPerLane;
constexpr int kTbStages = LORA_TB_STAGES;    // TMA load ring depth (rows in flight per warp)
constexpr int kTbOutBufs = LORA_TB_OUTBUFS;  // output staging rows per warp
constexpr int kTbCtasPerSm = LORA_TB_CTAS;   // resident CTAs (of kWarpsPerCta warps) per SM the kernel is built for
constexpr int kTbRing = kTbStages * kTbRowCells * 8, kTbOut = kTbOutBufs * kTbRowCells * 8, kTbMail = kMaxTb1 * 2 * 64,
              kTbBars = 64;
constexpr int kTbWarpSmem = ((kTbRing + kTbOut + kTbMail + kTbBars + 1023) / 1024) * 1024;
constexpr int kSmem1Tb = kWarpsPerCta * kTbWarpSmem;
static_assert(kTbStages * 8 <= kTbBars, "mbarrier area too small");

// 3-D: CTA tile of 32 rows x 128 columns per plane, 8 warps x (4 rows x 128 cols); warp 0 lane 0 also drives TMA
constexpr int k3TileRows = 32;
constexpr int k3TileCols = 128;
constexpr int k3BoxRows = k3TileRows + 2;
constexpr int k3BoxCols = k3TileCols + 4;  // 2 halo columns each side keeps the box origin 16-byte aligned
constexpr int k3Stages = 4;
constexpr int k3StageBytes = ((k3BoxRows * k3BoxCols * 8 + 127) / 128) * 128;
constexpr int k3GuardBytes = 128;  // one guard word per warp (stage release, see stencil3d.cu: plane_phase)
constexpr int k3Smem = k3Stages * k3StageBytes + 2 * k3Stages * 8 + k3GuardBytes;
constexpr int k3Warps = 8;
constexpr int k3Threads = 32 * k3Warps;

// ---- temporally blocked 3-D kernel (stencil3d_tb.cu): two launches per sweep ----
// thread tile 32 rows x 128 columns (8 warps x 4 rows, lane = 2 column pairs); a CTA writes 30 x 128 cells of it
constexpr int kT3Rm = 4;  // rows per lane: 3 -> 4 once the weights had moved to uniform registers (230-254 registers, no spills): +4..8 %
constexpr int kT3Rows = kT3Rm * k3Warps;
constexpr int kT3OutRows = kT3Rows - 2;
constexpr int kT3OutCols = k3TileCols;      // the level-1 columns just outside a tile are extra cells (stencil3d_tb.cu)
constexpr int kT3BoxRows = kT3Rows + 2;
constexpr int kT3StageBytes = ((kT3BoxRows * k3BoxCols * 8 + 127) / 128) * 128;
constexpr int kT3EdgePitch = k3TileCols + 4;  // columns -2 .. 129 of a level-1 edge row
constexpr int kT3EdgeBytes = 2 * k3Warps * 2 * kT3EdgePitch * 8;  // first / last level-1 row of every warp, double-buffered
constexpr int kT3Smem = k3Stages * kT3StageBytes + kT3EdgeBytes + 2 * k3Stages * 8 + k3GuardBytes;

struct Weights1D {
    double w[9];
};

// How one launch's range [lo, hi) of the swept (outermost) axis is cut into up to three SEGMENTS, and what a
// multi-GPU slab does with them (new: the reference is single-GPU).  A plain launch is one segment.  A slab launch
// that faces neighbours is [lo band][hi band][middle]: the two bands are the cells the neighbours' next sweep
// reads (ghost-zone width), their tasks are enumerated -- and therefore dispatched -- FIRST, every cell they store
// at out[x] is stored at out[x + mirror] too (the neighbour's ghost zone, peer memory over NVLink), and when the
// last task of a band has stored, it raises a 64-bit flag in the neighbour's memory to `seq`.  The middle segment
// reads this GPU's own cells only.  One kernel launch per sweep: halo transfer, its completion signal and the
// interior all overlap inside it.
constexpr int kMaxSegs = 3;
struct Segs {
    int nseg;                                // 1..3
    long long lo[kMaxSegs], hi[kMaxSegs];    // segment ranges on the swept axis (the kernel's own coordinate)
    long long chunk[kMaxSegs];               // chunk length of the segment's tasks along that axis
    long long first[kMaxSegs + 1];           // segment s owns tasks (3-D: plane chunks) [first[s], first[s + 1])
    long long mirror[kMaxSegs];              // != 0: cells of [mlo, mhi) stored at out[x] are also stored at out[x + mirror]
    long long mlo[kMaxSegs], mhi[kMaxSegs];  // the mirrored part of the segment on the swept axis (default: all of it)
    int early[kMaxSegs];                     // 3-D: arrive right after index mhi - 1 was stored, not at the end of the task
    unsigned long long *flag[kMaxSegs];      // nullptr, or the flag (peer memory) raised to `seq` when the segment is done
    unsigned long long *count[kMaxSegs];     // arrival counter of the segment (this GPU's memory, only ever grows)
    unsigned long long target[kMaxSegs];     // counter value that completes the segment in THIS launch
    unsigned long long seq;
};

struct Weights2D {
    double vert[3][7];
    double horiz[3][7];
    double centre;
    double residual[8];
};

struct WeightsDirect49 {
    double w[49];
};

struct Weights3D {
    double a[3], b[3], c[3];  // SEP3
    double star[7];           // STAR7: centre, n-1, n+1, m-1, m+1, h-1, h+1
    double direct[27];        // DIRECT27
};

#ifdef __CUDACC__
__device__ __forceinline__ int seg_of(const Segs &sg, long long task) {
    int s = 0;
    if (sg.nseg > 1 && task >= sg.first[1]) s = 1;
    if (sg.nseg > 2 && task >= sg.first[2]) s = 2;
    return s;
}
// One thread per task calls this AFTER the task's storing threads have executed __threadfence_system() and been
// joined (__syncwarp / __syncthreads): the last arrival of the segment publishes `seq` in the neighbour's flag.
__device__ __forceinline__ void seg_arrive(const Segs &sg, int s) {
    if (sg.flag[s] == nullptr) return;
    const unsigned long long old = atomicAdd(sg.count[s], 1ULL);
    if (old + 1 == sg.target[s]) {
        __threadfence_system();
        asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(sg.flag[s]), "l"(sg.seq) : "memory");
    }
}
#endif

struct Geom1D {
    const double *in;   // padded source
    double *out;        // padded destination
    long long n;        // interior length of the device array
    long long lo, hi;   // interior range of this launch (lo even)
    int rows_per_task;  // 128-element rows one warp sweeps
    long long ntasks;
    int vec4;           // 256-bit stores allowed (lo % 4 == 0 and 32-byte aligned base)
    long long mirror;   // != 0: every cell stored at out[x] is also stored at out[x + mirror] (a neighbour's halo, over NVLink)
};

// temporally blocked 1-D sweep (stencil1d_tb.cu): TB launches fused.  X = PADDED coordinates (interior cell i = X 4 + i)
struct Geom1DTB {
    const double *in;        // padded source (level 0)
    double *out;             // padded destination (level TB)
    const double *halo_src;  // padded buffer whose halo cells hold the caller's halo (buffer 0 of the ping-pong)
    long long n;             // interior length of the device array
    Segs sg;                 // segments in PADDED coordinates X: segment s writes cells [lo[s], hi[s]); its tasks sweep
                             // chunk[s] output rows (row r = cells [512 r - 4 TB, +512)) each
    int tb;                  // time steps fused by this launch (1..kMaxTb1)
    long long ntasks;
    int par0;                // parity of the launch count before level 0 (0: level 0 sees the caller's halo)
    int par_mask;            // 1: the halo alternates caller's / zero from level to level (the reference, S2); 0: every
                             // level sees what par0 says (fixed caller's halo = Dirichlet, or fixed zero)
    int virt_left, virt_right;  // this end of the array is an end of the global line: halo cells are virtual
    int use_tma;             // 0: array too small for the tensor maps, everything goes through plain accesses
    long long xcov;          // level-0 cells X >= xcov are not covered by the load map
    long long out_off;       // the store map starts at cell out_off (= -4 TB mod 16) ...
    long long out_rows;      // ... and covers out_rows rows of 16 cells
};

struct Geom2D {
    double *out;
    long long pitch;  // padded columns
    int m, n;
    Segs sg;          // segments = interior row ranges; a task = (strip, chunk of chunk[s] rows); strips vary fastest
    int nstrips;
    int ntasks;
    int vec4;  // 256-bit stores allowed (n % 4 == 0, 32-byte aligned base, mirrors a multiple of 4 doubles away)
};

// fused 2-D kernel: tasks of the first / last strip stage their rows' caller's-halo columns in shared memory
constexpr int kTb2Max = 3;
constexpr int kEdgeRows2Tb = 160;                          // longest edge-strip task (output rows)
constexpr int kHalRows2Tb = kEdgeRows2Tb + 6 * kTb2Max;    // its input rows
constexpr int kSmem2TbScratch = kSmem12 + kWarpsPerCta * kHalRows2Tb * 8 * 8;  // one scratch word per warp behind the staging area
constexpr int kSmem2Tb = kSmem2TbScratch + 64;

// temporally blocked 2-D sweep (stencil2d_tb.cu): TB (odd) launches fused; strips write 128 - 8 (TB - 1) columns
struct Geom2DTB {
    double *out;
    const double *halo_src;  // padded buffer whose halo ring holds the caller's halo (buffer 0 of the ping-pong)
    long long pitch;         // padded columns
    int m, n;
    Segs sg;                 // band segments (and every segment of a narrow grid): tasks = (strip, chunk of chunk[s] <=
                             // kEdgeRows2Tb rows), strips fastest.  The LAST segment of a grid with nstrips >= 3 is the
                             // main one: [row_lo, row_hi) below, cut as decode_task_2dtb describes
    int row_lo, row_hi;      // interior rows of the main segment
    int rows_per_chunk;
    int nstrips, nchunks;
    int edge_rows, nedge;    // nstrips >= 3: the two edge strips run as 2 * nedge tasks of edge_rows (<= kEdgeRows2Tb) rows
    int ntasks;              // all segments
    int par0;                // parity of the launch count before level 0 (== parity of the source buffer)
    int par_mask;            // 1: alternating halo (reference, S2); 0: every level sees what par0 says (Dirichlet / zero)
    int virt_top, virt_bot;  // rows beyond that end are the global halo ring (virtual halo), not neighbour-slab data
    int vec4;
};

// Which strip and rows a warp task of the fused 2-D kernel handles -- shared by the kernel and by the host-side
// coverage test (lora_debug_tasks_2dtb).  Edge strips (first / last: every row is patched) are cut into short tasks of
// g.edge_rows rows, whose halo columns fit the shared-memory staging area, and come first; then the inner strips chunk
// by chunk with the first and last chunk (whose first / last rows are patched) in front.  Returns false for an empty task.
#ifdef __CUDACC__
__host__ __device__
#endif
inline bool decode_task_2dtb(const Geom2DTB &g, int task, int &strip, int &r0, int &R, int &seg) {
    int s = 0;
    while (s + 1 < g.sg.nseg && task >= g.sg.first[s + 1]) s++;
    seg = s;
    const int tt = task - (int)g.sg.first[s];
    if (g.nstrips >= 3 && s == g.sg.nseg - 1) {  // the main segment
        const int nedge = 2 * g.nedge;
        if (tt < nedge) {
            strip = (tt & 1) ? g.nstrips - 1 : 0;
            r0 = g.row_lo + (tt >> 1) * g.edge_rows;
            R = g.row_hi - r0 < g.edge_rows ? g.row_hi - r0 : g.edge_rows;
        } else {
            const int t = tt - nedge, inner = g.nstrips - 2;
            strip = 1 + t % inner;
            int chunk = t / inner;
            chunk = chunk == 0 ? 0 : (chunk == 1 ? g.nchunks - 1 : chunk - 1);
            r0 = g.row_lo + chunk * g.rows_per_chunk;
            R = g.row_hi - r0 < g.rows_per_chunk ? g.row_hi - r0 : g.rows_per_chunk;
        }
    } else {  // band segment, or a narrow grid where every strip is an edge strip (the host keeps chunk <= kEdgeRows2Tb)
        const int ch = (int)g.sg.chunk[s], lo = (int)g.sg.lo[s], hi = (int)g.sg.hi[s];
        strip = tt % g.nstrips;
        r0 = lo + (tt / g.nstrips) * ch;
        R = hi - r0 < ch ? hi - r0 : ch;
    }
    return R > 0;
}

struct Geom3D {
    double *out;
    long long row_pitch;    // padded columns
    long long plane_pitch;  // padded rows * padded columns
    int h, m, n;
    Segs sg;  // segments = interior plane ranges; first[] counts plane chunks (blockIdx.y), a CTA = (tile, chunk)
    int tiles_m, tiles_n;
    int vec4;
};

// direct-tap kernels without TMA (stencil_direct.cu): grids with an odd number of padded columns
struct GeomDirect {
    const double *in;
    double *out;
    long long pitch, plane_pitch;  // padded columns; padded rows * padded columns (3-D)
    int m, n;                      // rows (3-D: per plane), columns
    int col_blocks;                // blocks of 256 columns
    Segs sg;  // segments = interior rows (2-D) / planes (3-D); first[] counts rows / planes; arrivals: one per CTA
};
cudaError_t launch_direct(int dim, const GeomDirect &g, const WeightsDirect49 &w, cudaStream_t s);

// radius-2 3-D shapes (stencil3d_r2.cu): layout (h + 4) x (m + 4) x (n + 8), interior planes [lo, hi) of the launch
struct Geom3DR2 {
    const double *in;
    double *out;
    long long row_pitch, plane_pitch;  // padded columns; padded rows * padded columns
    int m, n;                          // interior rows per plane, columns
    long long lo, hi;                  // output planes of this launch
    int planes_per_chunk;              // blockIdx.z walks chunks of this many output planes (+ 4 planes of warm-up)
};
struct WeightsR2 {
    double w[125];  // [(dh + 2) * 25 + (dr + 2) * 5 + dc + 2]: every form's effective taps (STAR13 / DIRECT125 read these)
    double q[25];   // HSEP5: in-plane table, w[dh][dr][dc] = a[dh + 2] * q[(dr + 2) * 5 + dc + 2]
    double a[5];
    double b[5], c[5];  // SEP5: q[(dr + 2) * 5 + dc + 2] = b[dr + 2] * c[dc + 2]
};
cudaError_t launch_3d_r2(int form, Geom3DR2 g, const WeightsR2 &w, int sm_count, cudaStream_t s);  // picks planes_per_chunk
long long r2_planes_per_chunk(long long planes, long long ctas_per_plane, int sm_count);
int r2_cols_per_cta(int form, int variant);
int r2_rows_per_cta();

// periodic halo ring (boundary.cu): one axis of an array seen as [outer][len + 2 halo][inner]
cudaError_t launch_wrap_axis(double *buf, long long outer, long long len, int halo, long long inner, int sm_count,
                             cudaStream_t s);
void wrap_axis_host(double *buf, long long outer, long long len, int halo, long long inner);

cudaError_t launch_1d(const Geom1D &g, const Weights1D &w, cudaStream_t s);
cudaError_t launch_1d_tb(const CUtensorMap &imap, const CUtensorMap &omap, const Geom1DTB &g, const Weights1D &w,
                         cudaStream_t s);
cudaError_t launch_2d(int form, const CUtensorMap &tmap, const Geom2D &g, const Weights2D &w,
                      const WeightsDirect49 &wd, cudaStream_t s);
cudaError_t launch_2d_tb(int form, int tb, const CUtensorMap &tmap, const Geom2DTB &g, const Weights2D &w,
                         const WeightsDirect49 &wd, cudaStream_t s);
int strip_out_cols_2d_tb(int tb);
cudaError_t launch_3d(int form, const CUtensorMap &tmap, const Geom3D &g, const Weights3D &w, cudaStream_t s);

// fused 3-D sweep of two launches (stencil3d_tb.cu); tmap: boxes of k3BoxCols x kT3BoxRows x 1
struct Geom3DTB {
    double *out;
    long long row_pitch, plane_pitch;
    int h, m, n;
    Segs sg;  // segments = interior plane ranges; first[] counts plane chunks (blockIdx.y), a CTA = (tile, chunk); a slab's
              // bands are folded into its last / first chunk like Geom3D's (planes [mlo, mhi) mirrored, early flag)
    int tiles_m, tiles_n;     // tiles of kT3OutRows x kT3OutCols
    int vec4;
};
cudaError_t launch_3d_tb(int form, const CUtensorMap &tmap, const Geom3DTB &g, const Weights3D &w, cudaStream_t s);
cudaError_t kernels_init_3d_tb();
cudaError_t kernels_init();     // opt in to large dynamic shared memory once per process/device
cudaError_t kernels_init_1d();
cudaError_t kernels_init_1d_tb();
cudaError_t kernels_init_2d();
cudaError_t kernels_init_2d_tb();
cudaError_t kernels_init_3d();

}  // namespace lora
