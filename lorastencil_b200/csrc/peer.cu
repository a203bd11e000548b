// peer.cu -- the few primitives the multi-GPU slab driver needs to exchange halos WITHOUT a communication
// library: device buffers that neighbouring processes (one per GPU, same box) map into their own address space
// over NVLink (CUDA IPC), and monotonically increasing flags written / awaited in stream order
// (cuStreamWriteValue64 / cuStreamWaitValue64).  With these, a rank's edge-band kernel stores its rows straight
// into the neighbour's halo rows (the `mirror` argument of lora_plan_step_mirror) and then bumps the neighbour's
// flag; the neighbour's next edge-band launch waits for that flag in its own stream.  New functionality: the
// reference is single-GPU (no cudaSetDevice / NCCL / MPI anywhere under src/).
#include <cstdio>
#include <cstring>

#include <cuda.h>
#include <cuda_runtime.h>

#include "../../include/lorastencil.h"

namespace {

typedef CUresult (*write64_fn)(CUstream, CUdeviceptr, cuuint64_t, unsigned int);
typedef CUresult (*wait64_fn)(CUstream, CUdeviceptr, cuuint64_t, unsigned int);

template <typename F>
F entry(const char *name) {
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint(name, &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
        return nullptr;
    return reinterpret_cast<F>(p);
}

int cuda_rc(cudaError_t e) { return e == cudaSuccess ? LORA_OK : LORA_ERR_CUDA; }

}  // namespace

extern "C" int lora_peer_alloc(void **ptr, unsigned long long bytes, void *handle64) {
    if (!ptr || !handle64 || bytes == 0) return LORA_ERR_ARG;
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    void *p = nullptr;
    cudaError_t e = cudaMalloc(&p, bytes);
    if (e != cudaSuccess) return LORA_ERR_CUDA;
    if ((e = cudaMemset(p, 0, bytes)) != cudaSuccess) {
        cudaFree(p);
        return LORA_ERR_CUDA;
    }
    cudaIpcMemHandle_t h;
    if ((e = cudaIpcGetMemHandle(&h, p)) != cudaSuccess) {
        cudaFree(p);
        return LORA_ERR_CUDA;
    }
    std::memcpy(handle64, &h, 64);
    *ptr = p;
    return LORA_OK;
}

extern "C" int lora_peer_free(void *ptr) { return cuda_rc(cudaFree(ptr)); }

extern "C" int lora_peer_open(const void *handle64, void **ptr) {
    if (!ptr || !handle64) return LORA_ERR_ARG;
    cudaIpcMemHandle_t h;
    std::memcpy(&h, handle64, 64);
    return cuda_rc(cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess));
}

extern "C" int lora_peer_close(void *ptr) { return cuda_rc(cudaIpcCloseMemHandle(ptr)); }

extern "C" int lora_stream_write_flag(void *stream, void *flag, unsigned long long value) {
    static write64_fn fn = entry<write64_fn>("cuStreamWriteValue64");
    if (!fn) return LORA_ERR_UNSUPPORTED;
    return fn(static_cast<CUstream>(stream), reinterpret_cast<CUdeviceptr>(flag), value, CU_STREAM_WRITE_VALUE_DEFAULT) ==
                   CUDA_SUCCESS
               ? LORA_OK
               : LORA_ERR_CUDA;
}

extern "C" int lora_stream_wait_flag_geq(void *stream, void *flag, unsigned long long value) {
    static wait64_fn fn = entry<wait64_fn>("cuStreamWaitValue64");
    if (!fn) return LORA_ERR_UNSUPPORTED;
    return fn(static_cast<CUstream>(stream), reinterpret_cast<CUdeviceptr>(flag), value, CU_STREAM_WAIT_VALUE_GEQ) ==
                   CUDA_SUCCESS
               ? LORA_OK
               : LORA_ERR_CUDA;
}
