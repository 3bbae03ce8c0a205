// DMMA (mma.sync m8n8k4 fp64) throughput probe on B200: is the fp64 tensor path faster than DFMA?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/dmma_probe.bin tools/dmma_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int NACC>
__global__ void __launch_bounds__(256) dmma_kernel(double *out, int iters, double a0, double b0)
{
    double c[NACC][2];
#pragma unroll
    for (int i = 0; i < NACC; ++i) { c[i][0] = 0.0; c[i][1] = 0.0; }
    double a = a0 + threadIdx.x * 1e-9, b = b0 + threadIdx.x * 1e-9;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += c[i][0] + c[i][1];
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void __launch_bounds__(256) dfma_kernel(double *out, int iters, double a, double b)
{
    double x[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) x[k] = 1.0 + k * 1e-3 + threadIdx.x * 1e-9;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 8; ++r)
#pragma unroll
            for (int k = 0; k < 8; ++k) x[k] = fma(x[k], a, b);
    }
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < 8; ++k) s += x[k];
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <class F>
double time_ms(F f)
{
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    f();
    cudaEventRecord(e0);
    f();
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    return ms;
}

int main()
{
    double *d; cudaMalloc(&d, 148 * 8 * 256 * sizeof(double) * 4);
    const int iters = 20000;
    for (int bps = 1; bps <= 8; bps *= 2) {
        const int blocks = 148 * bps;
        double ms = time_ms([&] { dfma_kernel<<<blocks, 256>>>(d, iters, 0.999999, 1e-9); });
        printf("DFMA  %d blocks/SM: %.2f TFLOP/s\n", bps, 2.0 * 64 * iters * blocks * 256.0 / ms / 1e9);
        ms = time_ms([&] { dmma_kernel<4><<<blocks, 256>>>(d, iters, 0.5, 0.25); });
        printf("DMMA4 %d blocks/SM: %.2f TFLOP/s\n", bps, 2.0 * 256 * 4 * iters * blocks * 8.0 / ms / 1e9);
        ms = time_ms([&] { dmma_kernel<16><<<blocks, 256>>>(d, iters / 4, 0.5, 0.25); });
        printf("DMMA16 %d blocks/SM: %.2f TFLOP/s\n", bps, 2.0 * 256 * 16 * (iters / 4) * blocks * 8.0 / ms / 1e9);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
