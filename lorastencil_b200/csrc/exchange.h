// exchange.h -- internal interface between the plan layer (plan.cu) and the multi-GPU slab driver (slab.cu).
// New functionality: the reference is single-GPU (no cudaSetDevice / NCCL / MPI anywhere under src/).
#pragma once
#include "../../include/lorastencil.h"

// What ONE launch over [lo, hi) does for the neighbouring slabs.  The band_lo cells / rows / planes at the low end and
// the band_hi at the high end are what the neighbours' next sweep reads: the launch stores them a second time at
// mirror_lo[x] / mirror_hi[x] (the address in the neighbour's DESTINATION buffer that corresponds to dst[0]; peer
// memory), their tasks run first, and the last task of a band raises *flag_lo / *flag_hi (in the neighbour's memory)
// to `seq` (st.release.sys after the stores were fenced).  count_* are arrival counters in this device's memory that
// only ever grow; arrived_* the host's running totals for them (updated by the call).
struct lora_exchange {
    long long band_lo = 0, band_hi = 0;
    const double *mirror_lo = nullptr, *mirror_hi = nullptr;
    unsigned long long *flag_lo = nullptr, *flag_hi = nullptr;
    unsigned long long *count_lo = nullptr, *count_hi = nullptr;
    unsigned long long *arrived_lo = nullptr, *arrived_hi = nullptr;
    unsigned long long seq = 0;
};

// One launch of `tb` fused time steps (tb = 1: also the unfused kernels of 2-D / 3-D) with an exchange folded in.
// Arguments as lora_plan_step_fused (include/lorastencil.h); ex may be nullptr.
int lora_plan_step_exchange(lora_plan_t *p, const double *src, double *dst, const double *halo_src, long long lo,
                            long long hi, int tb, int launches_before, int virt_lo, int virt_hi, const lora_exchange *ex,
                            void *stream);

// error text for the calling thread (lora_last_error), settable from slab.cu / peer.cu
int lora_fail(int code, const char *fmt, ...);
// halo ring of dst <- halo ring of src (src == nullptr: zeros): side halo of every local row / plane, leading / trailing
// halo rows only if `lead` / `trail` (sweeps of two launches borrow buffer 1's ring for the caller's halo)
int lora_plan_copy_ring(lora_plan_t *p, double *dst, const double *src, int lead, int trail, void *stream);
