// dfma_operands.cu -- does the FP64 pipe care where the multiplier of a DFMA comes from?  acc[i] = fma(x[j], w[k], acc[i])
// with w[k] (a) in uniform registers (kernel parameter, `DFMA R, R, UR, R`) and (b) in vector registers (loaded from
// global memory, `DFMA R, R, R, R`: three 64-bit register operands per instruction), 16 accumulators x 8 weights per
// thread, at 4 .. 16 warps per SM.  The fused 2-D / 3-D kernels were compiled to form (b).
#include <cstdio>
#include <cuda_runtime.h>

constexpr int ITER = 2048, NA = 16, NW = 8;
struct W {
    double v[NW];
};

template <bool VEC>
__global__ void k(double *out, const double *in, const double *wg, const __grid_constant__ W wp) {
    double x[NA], acc[NA], w[NW];
#pragma unroll
    for (int i = 0; i < NA; i++) {
        x[i] = in[threadIdx.x + 32 * i];
        acc[i] = 0.0;
    }
#pragma unroll
    for (int k2 = 0; k2 < NW; k2++) w[k2] = VEC ? wg[k2 + (threadIdx.x & 1)] : wp.v[k2];
    for (int it = 0; it < ITER; it++) {
#pragma unroll
        for (int k2 = 0; k2 < NW; k2++)
#pragma unroll
            for (int i = 0; i < NA; i++) acc[i] = fma(x[(i + k2) % NA], w[k2], acc[i]);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NA; i++) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <bool VEC>
void run(double *out, const double *in, const double *wg, int sms, int warps) {
    W wp;
    for (int i = 0; i < NW; i++) wp.v[i] = 1e-3 * (i + 1);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    k<VEC><<<sms, 32 * warps>>>(out, in, wg, wp);
    cudaEventRecord(e0);
    k<VEC><<<sms, 32 * warps>>>(out, in, wg, wp);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double flop = 2.0 * NA * NW * ITER * 32.0 * warps * sms;
    printf("  multiplier in %s registers, warps/SM %2d: %6.2f TFLOP/s\n", VEC ? "vector " : "uniform", warps, flop / ms / 1e9);
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    double *out, *in, *wg;
    cudaMalloc(&out, sizeof(double) * p.multiProcessorCount * 1024);
    cudaMalloc(&in, sizeof(double) * 32 * NA);
    cudaMalloc(&wg, sizeof(double) * (NW + 1));
    cudaMemset(in, 0, sizeof(double) * 32 * NA);
    cudaMemset(wg, 0, sizeof(double) * (NW + 1));
    printf("%s, %d SMs\n", p.name, p.multiProcessorCount);
    for (int warps : {4, 8, 12, 16}) {
        run<false>(out, in, wg, p.multiProcessorCount, warps);
        run<true>(out, in, wg, p.multiProcessorCount, warps);
    }
    return 0;
}
