/*
 * lorastencil.h -- C ABI of liblorastencil_b200.so, the B200-native (sm_100a) replacement for
 * the LoRAStencil host operators.  Plain pointers and sizes only; no torch / C++ types.
 *
 * Citations are relative to the reference tree (zondie17/LoRAStencil).
 *
 * Layer 1 (drop-in): lora_gpu_*  == the reference's gpu_* host operators
 *     src/1d/1d_utils.h:45-47   gpu_1d1r, gpu_1d2r
 *     src/2d/2d_utils.h:47-51   gpu_star_2d1r, gpu_star_2d3r, gpu_box_2d3r
 *     src/3d/3d_utils.h:44-48   gpu_box_3d1r, gpu_star_3d1r
 *   Same argument order and meaning, same buffer semantics (halo-padded row-major FP64 host
 *   arrays, ping-pong over `times` launches, `out` = whole padded buffer `times % 2`), same
 *   stdout banner, same fatal-error behaviour (message + exit(1)).  The library additionally
 *   exports the reference's C++-mangled names (gpu_box_2d3r(...) etc., see
 *   include/lorastencil_dropin.hpp) so the reference's own main.cu links against it unmodified.
 *
 * Layer 2 (device-resident plan API): what the benchmark, the multi-GPU slab driver and a
 *   host application that keeps its grids in HBM use.  Returns error codes, never exits.
 *
 * Layer 3 (host-side low-rank decomposition): the C++ factorisation that replaces
 *   src/2d/gpu.cu:280-350 (pyramidal rank-1 peel) and its siblings, exposed for inspection.
 */
#ifndef LORASTENCIL_H
#define LORASTENCIL_H

#ifdef __cplusplus
extern "C" {
#endif

/* ---- shapes: the CLI names of README.md:39-48 ---- */
typedef enum {
    LORA_1D1R = 0,
    LORA_1D2R = 1,
    LORA_STAR2D1R = 2,
    LORA_BOX2D1R = 3,
    LORA_STAR2D3R = 4,
    LORA_BOX2D3R = 5,
    LORA_BOX3D1R = 6,
    LORA_STAR3D1R = 7,
    LORA_NUM_SHAPES = 8,   /* the reference's shapes (src/3d/3d_utils.h:39-42 and siblings) end here */
    /* radius-2 members of the family the reference does not have.  Their own layout:
     * (h+4) x (m+4) x (n+8) doubles (the reference's 3-D layout with the plane halo widened to the radius), 125 weights
     * [(dh+2)*25 + (dr+2)*5 + dc+2], every weight honoured in both modes (there is no reference operator whose quirks
     * REFERENCE mode could restate).  Available through lora_gpu_run_host / lora_gpu_{box,star}_3d2r and the plan API
     * on one GPU, one launch per time step; slabs, fused sweeps and the CLI do not know them. */
    LORA_BOX3D2R = 8,
    LORA_STAR3D2R = 9,
    LORA_NUM_SHAPES_EXT = 10
} lora_shape_t;

/* how `params` is interpreted when a plan is built */
typedef enum {
    /* exactly what the reference GPU operator does with `params`, quirks included:
     * star2d1r / star3d1r ignore it (src/2d/gpu.cu:486-487, src/3d/gpu_star.cu:142-151),
     * box3d1r reads params[0..2] only (src/3d/gpu_box.cu:161), box2d applies the three rank-1
     * terms of its peel and drops the 1x1 remainder (src/2d/gpu.cu:349-358). */
    LORA_WEIGHTS_REFERENCE = 0,
    /* every one of the 9 / 49 / 27 weights is honoured (== the reference's test_cpu); the host
     * decomposition picks the cheapest exact form (cross, pyramid, rank-1 + residual,
     * separable, direct taps). */
    LORA_WEIGHTS_GENERAL = 1
} lora_weight_mode_t;

/* error codes of layer 2/3 */
enum {
    LORA_OK = 0,
    LORA_ERR_ARG = 1,      /* bad shape / size / null pointer / misaligned buffer */
    LORA_ERR_CUDA = 2,     /* a CUDA runtime or driver call failed */
    LORA_ERR_UNSUPPORTED = 3
};

/* ------------------------------------------------------------------------------------------
 * Layer 1: drop-in host operators (host buffers in, host buffers out)
 * ------------------------------------------------------------------------------------------
 * in / out: PADDED arrays  1-D n+8 | 2-D (m+8)x(n+8) | 3-D (h+2)x(m+4)x(n+8)  doubles.
 * params:   9 | 49 | 27 doubles.   times: number of launches.   No size-multiple constraints
 * (the reference needs n%1024, m%32, n%64, m%8: src/1d/gpu_1r.cu:112, src/2d/gpu.cu:402-403,
 * src/3d/gpu_box.cu:201-202).  1-D copies back n+7 doubles like src/1d/gpu_1r.cu:134. */
void lora_gpu_1d1r(const double *in, double *out, const double *params, int times, int input_n);
void lora_gpu_1d2r(const double *in, double *out, const double *params, int times, int input_n);
void lora_gpu_star_2d1r(const double *in, double *out, const double *params, int times, int input_m, int input_n);
void lora_gpu_star_2d3r(const double *in, double *out, const double *params, int times, int input_m, int input_n);
void lora_gpu_box_2d3r(const double *in, double *out, const double *params, int times, int input_m, int input_n);
void lora_gpu_box_3d1r(const double *in, double *out, const double *params, int times, int input_h, int input_m, int input_n);
void lora_gpu_star_3d1r(const double *in, double *out, const double *params, int times, int input_h, int input_m, int input_n);
/* radius-2 shapes (new): in / out (h+4) x (m+4) x (n+8) doubles, params 125 doubles (NULL = lora_reference_table) */
void lora_gpu_box_3d2r(const double *in, double *out, const double *params, int times, int input_h, int input_m, int input_n);
void lora_gpu_star_3d2r(const double *in, double *out, const double *params, int times, int input_h, int input_m, int input_n);

/* generic form of the seven above: shape selects the operator (box2d1r and box2d3r both run
 * gpu_box_2d3r, src/2d/main.cu:276-279); dims = {n} | {m,n} | {h,m,n}.  Same fatal-error
 * behaviour.  mode as in lora_weight_mode_t. */
void lora_gpu_run_host(int shape, int mode, const double *in, double *out, const double *params, int times,
                       const long long *dims);

/* 1 (default): print the reference's banner "LoRAStencil(<dim> <shape>): / Time = N[ms] /
 * GStencil/s = x" from the drop-in operators (src/2d/gpu.cu:415-419); 0: stay silent.
 * Returns the previous value.  Also settable with the environment variable LORA_QUIET=1. */
int lora_set_verbose(int on);

/* milliseconds the last lora_gpu_* call of the CALLING THREAD spent in its launch loop (the reference's timed region:
 * launches + device sync, src/2d/gpu.cu:408-414), and in the whole call (alloc + H2D + D2H too) */
double lora_last_loop_ms(void);
double lora_last_total_ms(void);

/* The 1-D drop-in operators overlap their copies with their launches: a cell after `times` launches depends on
 * 4 * times cells either side only, so a long line is cut into chunks with ghost margins of that width; every
 * chunk runs all its launches on its own while the next chunk's H2D copy and the previous chunk's D2H copy are
 * in flight (bit-identical results; lora_last_loop_ms and the banner's Time are then the time during which at least
 * one chunk's launch loop was running).  Returns how
 * many chunks the last lora_gpu_* call used (1 = plain H2D -> launches -> D2H).  Environment: LORA_CHUNKS=0
 * disables the overlap, LORA_CHUNKS=k forces k chunks. */
int lora_last_chunks(void);

/* Every other drop-in call (2-D, 3-D, short 1-D lines) overlaps its copies with its launch loop too: the grid is cut
 * into time-skewed bands along the outermost axis (sweep s of band k covers rows [B_k - s r, B_k+1 - s r), r = the
 * reach of one sweep), so a band depends only on bands before it and runs all its sweeps as soon as it is uploaded
 * while the next band uploads and the previous one downloads -- the same launches on the same operands, bit-identical
 * results, no redundant work.  Pageable host buffers are staged through pinned memory by worker threads.  Returns how
 * many bands the last call used (1 = copy -> launch loop -> copy).  Environment: LORA_BANDS=k forces k bands. */
int lora_last_bands(void);

/* free the device workspace the drop-in operators cache between calls */
void lora_release_workspace(void);

/* Multi-GPU behind the same surface (new; the reference caller is one process calling one operator,
 * src/2d/main.cu:268-280): with the environment variable LORA_NGPU=k, or after lora_set_gpus(k), every lora_gpu_* /
 * gpu_* call cuts its grid into k slabs along the outermost axis, one per GPU of this process (cudaSetDevice + peer
 * access), and exchanges the ghost zones inside the kernels over NVLink (lora_slabset_* below).  Same buffer
 * semantics, banner and timed region; bit-identical results.  A grid too thin for k slabs runs on one GPU.
 * lora_set_gpus returns the previous setting; lora_last_gpus how many GPUs the last call actually used. */
int lora_set_gpus(int k);
int lora_last_gpus(void);

/* ------------------------------------------------------------------------------------------
 * Layer 2: device-resident plans
 * ------------------------------------------------------------------------------------------ */
typedef struct lora_plan lora_plan_t;

/* dims = interior sizes {n} | {m,n} | {h,m,n} of the grid THIS device holds (for a slab: the
 * slab's interior).  params may be NULL = the reference CLI's table for `shape`. */
int lora_plan_create(lora_plan_t **plan, int shape, int mode, const double *params, const long long *dims);
void lora_plan_destroy(lora_plan_t *plan);

/* number of doubles of one padded device buffer for this plan */
long long lora_plan_padded_elems(const lora_plan_t *plan);

/* One launch: read padded device buffer `src`, write the interior of `dst` for outermost-axis
 * interior indices [lo, hi) (0 <= lo <= hi <= dims[0]); asynchronous on `stream`
 * (a cudaStream_t passed as void*, NULL = default stream).  Halo cells of dst are not touched
 * (S2 of SURVEY.md section 8a). */
int lora_plan_step(lora_plan_t *plan, const double *src, double *dst, long long lo, long long hi, void *stream);

/* `times` launches ping-ponging buf0 -> buf1 -> buf0 ...; launch i reads buf[i%2].  The result
 * is in buf[times%2].  Asynchronous on `stream`. */
int lora_plan_run(lora_plan_t *plan, double *buf0, double *buf1, int times, void *stream);

/* Temporal blocking (new; the reference launches one kernel per time step).  lora_plan_run fuses up to
 * `tb` consecutive launches into one sweep that keeps the intermediate grids on chip; results are
 * bit-identical to unfused launches, halo semantics (S2) included.
 * 1-D: tb = 1..15, default 15 (environment variable LORA_TB).
 * 2-D: sweeps of 3 launches (tb >= 3; every low-rank form) or of 2 (tb == 2; every form but the cross), anything else
 * means one launch per step; default 3 for the cross form, 2 for the diamond and pyramid forms, 1 for the direct /
 * rank-2 / rank-3 forms and for odd column counts (environment variable LORA_TB2=3|2|1).  For large cross grids the
 * default is only provisional: the first lora_plan_run times 3 single launches against 1 fused sweep on a scratch grid
 * of the same width and keeps the winner (cached per form, size and device; an explicit lora_plan_set_temporal_block
 * or LORA_TB2 is final).
 * 3-D: sweeps of 2 launches (tb >= 2) for the 7-point and separable forms, default on (LORA_TB3=1 turns it off); the
 * 27-tap form, odd column counts and the radius-2 shapes run one launch per step.
 * lora_plan_temporal_block returns what lora_plan_run will fuse. */
int lora_plan_set_temporal_block(lora_plan_t *plan, int tb);
int lora_plan_temporal_block(const lora_plan_t *plan);

/* Boundary modes (new; the reference has exactly one).  LORA_BOUNDARY_REFERENCE: the reference's ping-pong -- no
 * launch writes a halo cell, buffer 0 holds the caller's halo, buffer 1 zeros, so even launches see the caller's halo
 * and odd launches a zero halo (S2: src/2d/gpu.cu:396-400, store offsets :106).  LORA_BOUNDARY_DIRICHLET: the caller's
 * halo values are the boundary condition of EVERY launch (lora_plan_run copies the halo ring of buf0 into buf1 first).
 * LORA_BOUNDARY_ZERO: zero halo for every launch (the ring of both buffers is cleared, buf0's included).
 * LORA_BOUNDARY_PERIODIC: the grid is a torus -- before every launch lora_plan_run rewrites the halo ring of the
 * source buffer with the interior cells it wraps around to (corners included), and once more on the result buffer, so
 * the caller's halo values are never read; one launch per time step (a fused sweep would need a ghost zone of radius x
 * depth cells, the storage halo of S1 holds one radius), every axis at least as long as its storage halo (4 / 4,4 /
 * 1,2,4).  Applies to lora_plan_run / lora_plan_step*; the drop-in operators and the slab drivers always use
 * REFERENCE. */
enum { LORA_BOUNDARY_REFERENCE = 0, LORA_BOUNDARY_DIRICHLET = 1, LORA_BOUNDARY_ZERO = 2, LORA_BOUNDARY_PERIODIC = 3 };
int lora_plan_set_boundary(lora_plan_t *plan, int mode);
int lora_plan_boundary(const lora_plan_t *plan);
/* The periodic refresh on its own: halo ring of `buf` <- periodic image of its interior (for callers that drive
 * lora_plan_step themselves).  Asynchronous on `stream`. */
int lora_plan_wrap_ring(lora_plan_t *plan, double *buf, void *stream);

/* One FUSED launch of `tb` time steps over interior range [lo, hi) of the outermost axis.
 * 2-D (tb = 1, 3, or 2 for the diamond / pyramid forms): rows [lo, hi); the ring of src must be the one its time parity
 * calls for (caller's halo at even `launches_before`, zeros at odd -- what the ping-pong gives when every sweep advances
 * an odd number of steps; sweeps of 2 all start at even times, which is why lora_plan_run lends buffer 1 the caller's
 * ring while they run), halo_src is required, virt_lo / virt_hi say whether the rows above 0 / below m are the global
 * halo ring.  3-D sweeps of 2 launches are issued by lora_plan_run and the slab drivers only (this entry point: 1-D, 2-D).
 * 1-D (tb = 1..15): reads src[lo - 4 tb, hi + 4 tb) clipped to the array, writes dst[lo, hi).  `launches_before` = time steps
 * already applied to src (its parity selects the halo each level sees).  virt_lo / virt_hi: that end of
 * the array is an end of the global line, whose halo cells are virtual -- caller's halo (read from the
 * padded buffer `halo_src`) at even times, zero at odd times; with 0 the side is an inter-slab boundary
 * whose halo / ghost cells hold real neighbour data in src. */
int lora_plan_step_fused(lora_plan_t *plan, const double *src, double *dst, const double *halo_src, long long lo,
                         long long hi, int tb, int launches_before, int virt_lo, int virt_hi, void *stream);

/* ---- multi-GPU slabs without a communication library (new; the reference is single-GPU) ----
 * One process per GPU on one box.  Each rank allocates its ping-pong buffers with lora_peer_alloc and hands the
 * 64-byte handle to its neighbours, which map the buffer with lora_peer_open (CUDA IPC over NVLink).  A rank then
 * launches its EDGE BANDS with lora_plan_step_mirror / lora_plan_step_fused_mirror: every cell the launch stores
 * at dst[x] is also stored at mirror_base[x], mirror_base being the address in the neighbour's buffer that
 * corresponds to dst[0] (neighbour buffer + a row / plane / cell shift), i.e. the band lands directly in the
 * neighbour's halo.  Ordering between ranks: 64-bit flags in peer memory, written (lora_stream_write_flag) and
 * awaited (lora_stream_wait_flag_geq: blocks the stream until *flag >= value) in stream order. */
int lora_peer_alloc(void **ptr, unsigned long long bytes, void *handle64_out);
int lora_peer_free(void *ptr);
int lora_peer_open(const void *handle64, void **ptr);
int lora_peer_close(void *ptr);
int lora_stream_write_flag(void *stream, void *flag, unsigned long long value);
int lora_stream_wait_flag_geq(void *stream, void *flag, unsigned long long value);
int lora_plan_step_mirror(lora_plan_t *plan, const double *src, double *dst, long long lo, long long hi,
                          const double *mirror_base, void *stream);
int lora_plan_step_fused_mirror(lora_plan_t *plan, const double *src, double *dst, const double *halo_src, long long lo,
                                long long hi, int tb, int launches_before, int virt_lo, int virt_hi,
                                const double *mirror_base, void *stream);

/* ---- slabs: the whole multi-GPU sweep in the library (new) ----
 * The grid is cut along its outermost axis into `world` contiguous slabs; each keeps the reference's two ping-pong
 * buffers for its rows plus ghost zones of (stencil radius x deepest temporal block) towards its neighbours and the
 * reference's storage halo towards the ends of the grid.  One sweep of a slab is ONE kernel launch: the tasks of the
 * two bands a neighbour needs run first, store their cells a second time straight into the neighbour's ghost zone
 * (peer memory) and raise a flag there when the band is complete; the interior overlaps that; the next sweep waits
 * for the neighbours' flags in stream order.  No communication library on the data path.
 *
 * lora_slab_*: ONE slab on the current device -- for one-process-per-GPU ranks.  Rank r creates its slab, exports
 * three 64-byte CUDA IPC handles (buffer 0, buffer 1, flags), receives its neighbours' (any rendezvous: the Python
 * layer uses torch.distributed) and connects: side 0 = rank r-1, side 1 = rank r+1.  lora_slab_connect_local does the
 * same for a neighbour slab living in this process (peer access instead of IPC).
 * info10: {lo, hi (the slab's interior range on the global outermost axis), wl, wr (cells stored before / after the
 * slab: ghost or halo), off (first slab cell in plan-interior coordinates), local padded sizes [3], deepest temporal
 * block, ghost width}.  The padded local buffer mirrors global padded rows [lo + halo - wl, hi + halo + wr).
 * temporal_block 0 = the shape's default.  lora_slab_run is asynchronous on `stream`; after it the result is in
 * buffer lora_slab_result_index().  After (re)filling the buffers from outside call lora_slab_reset -- and make sure
 * (device sync + rendezvous) no neighbour is still sweeping. */
typedef struct lora_slab lora_slab_t;
int lora_slab_create(lora_slab_t **slab, int shape, int mode, const double *params, const long long *global_dims,
                     int world, int rank, int temporal_block);
void lora_slab_destroy(lora_slab_t *slab);
int lora_slab_info(const lora_slab_t *slab, long long *info10);
double *lora_slab_buffer(lora_slab_t *slab, int which);
int lora_slab_export(lora_slab_t *slab, void *handles192);
int lora_slab_connect_ipc(lora_slab_t *slab, int side, const void *handles192);
int lora_slab_connect_local(lora_slab_t *slab, int side, lora_slab_t *neighbour);
int lora_slab_reset(lora_slab_t *slab);
int lora_slab_sweep(lora_slab_t *slab, int tb, void *stream);
int lora_slab_run(lora_slab_t *slab, int times, void *stream);
int lora_slab_schedule(const lora_slab_t *slab, int times, int *blocks_out, int cap);
int lora_slab_result_index(const lora_slab_t *slab);
long long lora_slab_launch_count(const lora_slab_t *slab);
lora_plan_t *lora_slab_plan(lora_slab_t *slab);
/* host-only: the partition rule (CPU tests).  out8: lo, hi, wl, wr, off, local interior sizes [3] */
int lora_slab_geometry(int dim, const long long *global_dims, int world, int rank, long long ghost, long long *out8);

/* lora_slabset_*: all slabs of one grid on `ndev` devices of THIS process (devices NULL = 0..ndev-1); what the
 * drop-in operators use under LORA_NGPU.  load: H2D scatter of a padded host grid (S2: buffer 1 <- zeros);
 * run: `times` launches on every device, issued sweep by sweep, asynchronous; sync; store: D2H gather of the padded
 * result (S3). */
typedef struct lora_slabset lora_slabset_t;
int lora_slabset_create(lora_slabset_t **set, int shape, int mode, const double *params, const long long *global_dims,
                        int ndev, const int *devices);
void lora_slabset_destroy(lora_slabset_t *set);
int lora_slabset_load(lora_slabset_t *set, const double *host_padded_in);
int lora_slabset_run(lora_slabset_t *set, int times);
int lora_slabset_sync(lora_slabset_t *set);
int lora_slabset_store(lora_slabset_t *set, double *host_padded_out);
long long lora_slabset_launch_count(const lora_slabset_t *set);
int lora_slabset_temporal_block(const lora_slabset_t *set);

/* how many kernel launches the plan has issued so far (bench.py's gpu_launches) */
long long lora_plan_launch_count(const lora_plan_t *plan);

/* textual description of the kernel form chosen by the host decomposition, e.g.
 * "2d pyramid rank-3 (7/5/3) + centre 0"; owned by the plan */
const char *lora_plan_describe(const lora_plan_t *plan);

/* last error message of layer 2/3 on this thread ("" if none) */
const char *lora_last_error(void);

/* ---- host-side planning, exposed for tests (no GPU needed) ----
 * lora_debug_temporal_schedule: the temporal blocks lora_plan_run uses for `times` launches of a 1-D plan whose deepest
 * block is max_tb (their count has the parity of `times`: S3); returns the number of blocks.
 * lora_debug_tasks_2dtb: the warp tasks (strip, first row, rows) of one fused 2-D launch over rows [lo, hi) of an
 * m x n grid on a GPU with sm_count SMs, in launch order; returns the number of tasks. */
int lora_debug_temporal_schedule(int times, int max_tb, int *blocks_out, int cap);
/* the sweeps of a plan that fuses TWO launches (2-D diamond / pyramid forms, 3-D): an even number of 2s, then the
 * remaining 0..3 launches one by one (no 2s below 4 launches); returns the number of sweeps */
int lora_debug_pair_schedule(int times, int *blocks_out, int cap);
/* what the last fused-or-not probe of a 2-D plan measured (milliseconds for 3 single launches / for 1 fused sweep of
 * 3 on the scratch grid); returns the number of cached verdicts */
int lora_debug_tb2_probe(double *ms_unfused3, double *ms_fused);
int lora_debug_tasks_2dtb(int m, int n, int lo, int hi, int sm_count, int *strip_row_rows_out, int cap);
/* the same for a sweep of TWO launches (diamond / pyramid forms: strips of 120 columns, 12 resident warps per SM) */
int lora_debug_tasks_2dtb_pairs(int m, int n, int lo, int hi, int sm_count, int *out3, int cap);
/* the periodic halo refresh (lora_plan_wrap_ring: same work items, same axis order) applied to a HOST array of the padded
 * size of a dim-D grid with interior sizes dims[0..dim) */
int lora_debug_wrap_ring_host(int dim, const long long *dims, double *buf);
/* launch geometry of the radius-2 3-D kernels for an h x m x n grid: out4 = {grid.x, grid.y, plane chunks, planes per
 * chunk}; form = LORA_FORM_STAR13 | HSEP5 | DIRECT125 | SEP5, variant = LORA_R2_VARIANT (0..2) */
int lora_debug_r2_grid(int form, int variant, long long h, int m, int n, int sm_count, long long *out4);

/* ------------------------------------------------------------------------------------------
 * Layer 3: host low-rank decomposition (inspection / tests)
 * ------------------------------------------------------------------------------------------ */
enum {
    LORA_FORM_TAPS9 = 0,      /* 1-D: 9 direct taps */
    LORA_FORM_CROSS = 1,      /* 2-D star: column arm (with centre) + row arm (without) */
    LORA_FORM_PYRAMID = 2,    /* 2-D box: sum of 3 rank-1 terms with support 7/5/3 + centre */
    LORA_FORM_DIAMOND = 3,    /* 2-D: one rank-1 term of support 5 + 8 residual taps */
    LORA_FORM_DIRECT49 = 4,   /* 2-D: all 49 taps */
    LORA_FORM_SEP3 = 5,       /* 3-D box: one rank-1 term a (x) b (x) c */
    LORA_FORM_STAR7 = 6,      /* 3-D star: 7 taps */
    LORA_FORM_DIRECT27 = 7,   /* 3-D: all 27 taps */
    LORA_FORM_PYRAMID_PRUNED = 8, /* PYRAMID whose middle term is zero at offsets +-1 and whose centre remainder is
                                    zero (true for the reference's box table): those taps are not computed at all */
    LORA_FORM_RANK2 = 9,      /* 2-D: sum of 2 rank-1 terms of full support 7 (LU / cross approximation with full
                                 pivoting) -- any rank-2 table that is not pyramidal: 28 instead of 49 taps */
    LORA_FORM_RANK3 = 10,     /* 2-D: sum of 3 such terms: 42 taps */
    LORA_FORM_STAR13 = 11,    /* 3-D radius 2: 13 taps */
    LORA_FORM_HSEP5 = 12,     /* 3-D radius 2: a(h) (x) Q(m, n), rank 1 along the plane axis, any in-plane 5 x 5 table: 30 taps */
    LORA_FORM_DIRECT125 = 13, /* 3-D radius 2: all 125 taps */
    LORA_FORM_SEP5 = 14       /* 3-D radius 2: a(h) (x) b(m) (x) c(n), rank 1 along every axis: 15 taps */
};

typedef struct {
    int form;            /* LORA_FORM_* */
    int nterms;          /* rank-1 terms in use (0..3) */
    double vert[3][7];   /* term t, vertical (row-offset) profile, index = dr+3 */
    double horiz[3][7];  /* term t, horizontal (col-offset) profile, index = dc+3 */
    double centre;       /* extra weight on the centre tap (pyramid remainder) */
    double residual[8];  /* DIAMOND: (0,-3),(0,+3),(-3,0),(+3,0),(-2,-2),(-2,+2),(+2,-2),(+2,+2) */
    double recon_err;    /* max |sum of terms - effective weights| */
    int macs_per_cell;   /* multiply-adds per output cell of the chosen form */
} lora_decomp2d_t;

/* factor a 7x7 weight table; shape picks the reference quirks when mode == REFERENCE */
int lora_decompose_2d(int shape, int mode, const double *params49, lora_decomp2d_t *out);

/* structure found in a 125-weight table of a radius-2 3-D shape (every weight is honoured whatever the form) */
typedef struct {
    int form;          /* LORA_FORM_STAR13 | LORA_FORM_SEP5 | LORA_FORM_HSEP5 | LORA_FORM_DIRECT125 */
    double a[5];       /* HSEP5 / SEP5: profile along the plane axis, index dh+2 */
    double b[5], c[5]; /* SEP5: profiles along rows / columns */
    double q[25];      /* HSEP5 / SEP5: the in-plane table (SEP5: == b (x) c) */
    double recon_err;  /* max |factors multiplied out - table| (0 for STAR13 / DIRECT125) */
    int macs_per_cell; /* 13 | 15 | 30 | 125 */
} lora_decomp3d_r2_t;
int lora_decompose_3d_r2(int shape, const double *params125, lora_decomp3d_r2_t *out);

/* the weight table the reference CLI passes for `shape` (src/1d/main.cu:77-78, src/2d/main.cu:139-195,
 * src/3d/main.cu:112-125): 9 / 49 / 27 doubles; for the radius-2 shapes our own default, 125 doubles (box3d2r:
 * [1,2,3,2,1] (x) [1,2,3,2,1] (x) [1,2,3,2,1]; star3d2r: centre 3, arms 2 then 1) */
int lora_reference_table(int shape, double *table_out);

/* effective direct-tap weights (9 / 49 / 27 / 125 doubles) a plan built from (shape, mode, params)
 * applies -- what the parity tests feed to the CPU oracle */
int lora_effective_weights(int shape, int mode, const double *params, double *weights_out);

#ifdef __cplusplus
}
#endif
#endif /* LORASTENCIL_H */
