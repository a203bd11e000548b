// dfma_occupancy.cu -- how many resident warps per SM does the FP64 FMA pipe need?  Sustained DFMA rate as a function
// of warps per SM (one CTA per SM, 4 .. 32 warps) and of the number of independent FMA chains per thread (ILP 1 .. 16).
// The fused kernels run 8 warps per SM at 208-255 registers: this tells whether their 33-56 % of the pipe is an
// occupancy limit or a scheduling one.   nvcc -O3 -gencode arch=compute_100a,code=sm_100a dfma_occupancy.cu
#include <cstdio>
#include <cuda_runtime.h>

constexpr int ITER = 8192;

template <int ILP>
__global__ void k(double *out, double a, double b) {
    double x[ILP];
#pragma unroll
    for (int i = 0; i < ILP; i++) x[i] = threadIdx.x + i;
    for (int it = 0; it < ITER; it++) {
#pragma unroll
        for (int i = 0; i < ILP; i++) x[i] = fma(x[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; i++) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int ILP>
void run(double *out, int sms, int warps) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    k<ILP><<<sms, 32 * warps>>>(out, 1.0000001, 1e-9);
    cudaEventRecord(e0);
    k<ILP><<<sms, 32 * warps>>>(out, 1.0000001, 1e-9);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double flop = 2.0 * ILP * ITER * 32.0 * warps * sms;
    printf("  ilp %2d warps/SM %2d: %6.2f TFLOP/s\n", ILP, warps, flop / ms / 1e9);
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    double *out;
    cudaMalloc(&out, sizeof(double) * p.multiProcessorCount * 1024);
    printf("%s, %d SMs\n", p.name, p.multiProcessorCount);
    for (int warps : {4, 8, 12, 16, 32}) {
        run<1>(out, p.multiProcessorCount, warps);
        run<2>(out, p.multiProcessorCount, warps);
        run<4>(out, p.multiProcessorCount, warps);
        run<8>(out, p.multiProcessorCount, warps);
        run<16>(out, p.multiProcessorCount, warps);
    }
    return 0;
}
