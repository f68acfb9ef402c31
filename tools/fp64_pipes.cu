// fp64_pipes.cu -- micro-benchmark: FP64 FMA (DFMA) vs FP64 tensor MMA (DMMA m8n8k4) throughput on one GPU, alone
// and interleaved, to see whether they are separate pipes.  nvcc -O3 -gencode arch=compute_100a,code=sm_100a
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double &d0, double &d1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

template <int MODE>  // 0 = DFMA only, 1 = DMMA only, 2 = both interleaved
__global__ void __launch_bounds__(256) k(double *out, int iters, double a, double b)
{
    double f[16], d[16];
#pragma unroll
    for (int i = 0; i < 16; i++) { f[i] = threadIdx.x + i; d[i] = threadIdx.x - i; }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 16; i++) {
            if (MODE != 1) { f[i] = fma(f[i], a, b); }
        }
#pragma unroll
        for (int i = 0; i < 16; i += 2) {
            if (MODE != 0) dmma(d[i], d[i + 1], a, b);
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) s += f[i] + d[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE> void run(const char *name, double *out, int blocks)
{
    const int iters = 20000;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<blocks, 256>>>(out, 100, 1.0000001, 1e-9);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    k<MODE><<<blocks, 256>>>(out, iters, 1.0000001, 1e-9);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double thr = (double)blocks * 256;
    const double fma_flops = MODE != 1 ? thr * 16.0 * iters * 2 : 0;
    const double mma_flops = MODE != 0 ? (thr / 32) * 8.0 * iters * (8 * 8 * 4 * 2) : 0;
    printf("%-12s %8.3f ms  DFMA %.2f TFLOP/s  DMMA %.2f TFLOP/s  sum %.2f\n", name, ms, fma_flops / ms / 1e9,
           mma_flops / ms / 1e9, (fma_flops + mma_flops) / ms / 1e9);
}

int main()
{
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    const int blocks = p.multiProcessorCount * 4;
    double *out; cudaMalloc(&out, (size_t)blocks * 256 * 8);
    printf("%s, %d SMs\n", p.name, p.multiProcessorCount);
    run<0>("dfma", out, blocks);
    run<1>("dmma", out, blocks);
    run<2>("dfma+dmma", out, blocks);
    return 0;
}
