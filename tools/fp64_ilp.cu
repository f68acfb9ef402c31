// fp64_ilp.cu -- DFMA throughput as a function of independent chains per thread (ILP) and resident warps per SM:
// how many accumulators does a register block need in flight to keep the FP64 pipe busy at 12 warps per SM?
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP>
__global__ void k(double *out, int iters, double a, double b)
{
    double f[ILP];
#pragma unroll
    for (int i = 0; i < ILP; i++) f[i] = threadIdx.x + i;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int r = 0; r < 16 / ILP; r++)
#pragma unroll
            for (int i = 0; i < ILP; i++) f[i] = fma(f[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; i++) s += f[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int ILP> void run(double *out, int sms, int warps_per_sm)
{
    const int iters = 20000, threads = 128;
    const int blocks = sms * (warps_per_sm * 32 / threads);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<ILP><<<blocks, threads>>>(out, 100, 1.0000001, 1e-9);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    k<ILP><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("  ILP %2d: %6.2f TFLOP/s", ILP, (double)blocks * threads * 16.0 * iters * 2 / ms / 1e9);
}

int main()
{
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    double *out; cudaMalloc(&out, (size_t)p.multiProcessorCount * 2048 * 8);
    for (int w : {4, 8, 12, 16, 32}) {
        printf("%2d warps/SM:", w);
        run<1>(out, p.multiProcessorCount, w); run<2>(out, p.multiProcessorCount, w); run<4>(out, p.multiProcessorCount, w);
        run<8>(out, p.multiProcessorCount, w); run<16>(out, p.multiProcessorCount, w);
        printf("\n");
    }
    return 0;
}
