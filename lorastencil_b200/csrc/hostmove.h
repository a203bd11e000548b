// hostmove.h -- host <-> device copies for the drop-in operators that run at PCIe speed whatever memory the caller
// hands in.  The reference's caller mallocs pageable arrays (src/1d/main.cu:96-103, src/2d/main.cu:224-225) and the
// reference operator cudaMemcpy's them synchronously (src/2d/gpu.cu:396-400, :421).  A cudaMemcpyAsync from pageable
// memory is staged by the driver on ONE thread at 10-15 GB/s and blocks the caller; here pageable buffers go through
// a ring of pinned staging slots filled / drained by a small pool of worker threads (parallel memcpy), so the DMA
// engines see pinned memory and copies overlap the launch loop exactly as they do for pinned caller buffers.
#pragma once
#include <cstddef>

#include <cuda_runtime.h>

namespace lora {

class HostMover {
  public:
    // true when `p` is ordinary pageable host memory (not pinned / registered / managed)
    static bool pageable(const void *p);

    HostMover();
    ~HostMover();
    HostMover(const HostMover &) = delete;
    HostMover &operator=(const HostMover &) = delete;

    // dst_dev[0, bytes) <- src_host, ordered on `stream`.  Pinned source: one cudaMemcpyAsync.  Pageable source: the
    // calling thread returns once the last piece has been memcpy'd into a staging slot and its DMA queued.
    cudaError_t h2d(void *dst_dev, const void *src_host, size_t bytes, cudaStream_t stream);
    // dst_host <- src_dev[0, bytes), ordered on `stream`.  Pageable destination: the DMA lands in staging slots and a
    // drain thread copies every piece out as its DMA completes; the data is in dst_host after finish().
    cudaError_t d2h(void *dst_host, const void *src_dev, size_t bytes, cudaStream_t stream);
    // wait for every queued piece (and the streams' copies) to have reached its destination
    cudaError_t finish();

  private:
    struct Impl;
    Impl *impl_;
};

// the process-wide instance (created on first use, never destroyed: its threads must not race the CUDA runtime's own
// teardown at exit)
HostMover &global_mover();

}  // namespace lora
