// fp64_pipes.cu -- measures, on the GPU it runs on, the sustained FP64 rates that decide the kernel
// design:
//   (1) DFMA only                 (FP64 FMA pipe)
//   (2) DMMA.8x8x4 only           (mma.sync.m8n8k4.f64, the only FP64 tensor shape on sm_100a)
//   (3) both interleaved          (do the two pipes overlap?)
// Prints TFLOP/s for each; `make -C profiles/microbench && profiles/microbench/fp64_pipes`.
#include <cstdio>
#include <cuda_runtime.h>

constexpr int ITER = 4096;

__global__ void k_dfma(double *out, double a, double b) {
    double x[8];
#pragma unroll
    for (int i = 0; i < 8; i++) x[i] = threadIdx.x + i;
    for (int it = 0; it < ITER; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) x[i] = fma(x[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__device__ __forceinline__ void dmma(double &c0, double &c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

__global__ void k_dmma(double *out, double a, double b) {
    double c[8][2];
#pragma unroll
    for (int i = 0; i < 8; i++) c[i][0] = c[i][1] = threadIdx.x + i;
    for (int it = 0; it < ITER; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) dmma(c[i][0], c[i][1], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_both(double *out, double a, double b) {
    double c[4][2], x[8];
#pragma unroll
    for (int i = 0; i < 4; i++) c[i][0] = c[i][1] = threadIdx.x + i;
#pragma unroll
    for (int i = 0; i < 8; i++) x[i] = threadIdx.x - i;
    for (int it = 0; it < ITER; it++) {
#pragma unroll
        for (int i = 0; i < 4; i++) {
            dmma(c[i][0], c[i][1], a, b);
            x[2 * i] = fma(x[2 * i], a, b);
            x[2 * i + 1] = fma(x[2 * i + 1], a, b);
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 4; i++) s += c[i][0] + c[i][1];
#pragma unroll
    for (int i = 0; i < 8; i++) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
float time_ms(F f) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    f();
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    for (int r = 0; r < 5; r++) f();
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    return ms / 5;
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    const int sms = p.multiProcessorCount, threads = 512, blocks = sms * 4;
    double *out;
    cudaMalloc(&out, sizeof(double) * blocks * threads);
    const double nthreads = (double)blocks * threads, nwarps = nthreads / 32;
    float t1 = time_ms([&] { k_dfma<<<blocks, threads>>>(out, 1.0000001, 1e-9); });
    float t2 = time_ms([&] { k_dmma<<<blocks, threads>>>(out, 1.0000001, 1e-9); });
    float t3 = time_ms([&] { k_both<<<blocks, threads>>>(out, 1.0000001, 1e-9); });
    const double f1 = nthreads * ITER * 8 * 2;              // 8 DFMA / iter / thread
    const double f2 = nwarps * ITER * 8 * (8 * 8 * 4 * 2);  // 8 DMMA / iter / warp, 512 flop each
    const double f3 = nwarps * ITER * 4 * 512 + nthreads * ITER * 8 * 2;
    printf("{\"gpu\": \"%s\", \"sms\": %d, \"dfma_tflops\": %.2f, \"dmma_tflops\": %.2f, \"mixed_tflops\": %.2f, "
           "\"mixed_dmma_share_tflops\": %.2f, \"mixed_dfma_share_tflops\": %.2f}\n",
           p.name, sms, f1 / t1 / 1e9, f2 / t2 / 1e9, f3 / t3 / 1e9, nwarps * ITER * 4 * 512 / t3 / 1e9,
           nthreads * ITER * 8 * 2 / t3 / 1e9);
    return 0;
}
