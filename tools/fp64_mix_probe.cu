// fp64_mix_probe.cu -- FP64 pipe utilisation when DFMAs are interleaved with ALU / FMA-pipe
// instructions (what the prism kernels look like), and dependent-issue latency of DFMA.
#include <cstdio>
#include <cuda_runtime.h>

template <int NALU, int KIND>
__global__ void __launch_bounds__(128) mix(double *out, int iters, double a, double b, unsigned m)
{
    double x[8];
    unsigned v[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) { x[k] = threadIdx.x * 1e-3 + k; v[k] = threadIdx.x * 2654435761u + k; }
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                x[k] = fma(x[k], a, b);
#pragma unroll
                for (int j = 0; j < NALU; ++j) {
                    if (KIND == 0) v[(k + j) & 7] = (v[(k + j) & 7] ^ m) + (v[(k + j + 1) & 7] & 0x7fffffffu);   // LOP3 + IADD
                    if (KIND == 1) v[(k + j) & 7] = v[(k + j) & 7] * 3u + m;      // IMAD (FMA pipe)
                }
            }
        }
    }
    double s = 0; unsigned t = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) { s += x[k]; t ^= v[k]; }
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s + (double)t;
}

__global__ void latency(double *out, long long *cycles, int iters, double a, double b)
{
    double x = threadIdx.x;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 64; ++u) x = fma(x, a, b);
    }
    long long t1 = clock64();
    out[threadIdx.x] = x;
    if (threadIdx.x == 0) cycles[0] = t1 - t0;
}

template <int NALU, int KIND>
void run(const char *name, int bps)
{
    const int blocks = 148 * bps, threads = 128, iters = 1024;
    double *d; cudaMalloc(&d, (size_t)blocks * threads * 8);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    mix<NALU, KIND><<<blocks, threads>>>(d, iters, 0.999999, 1e-9, 12345u);
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 3; ++r) {
        cudaEventRecord(e0);
        mix<NALU, KIND><<<blocks, threads>>>(d, iters, 0.999999, 1e-9, 12345u);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    const double inst = (double)blocks * threads * iters * 64.0;
    printf("%-34s %2d warps/SM: FP64 pipe %.1f%%\n", name, bps * 4, 100.0 * inst / (best * 1e-3) / (148.0 * 64 * 1.965e9));
    cudaFree(d);
}

int main()
{
    double *d; long long *c; cudaMalloc(&d, 256 * 8); cudaMalloc(&c, 8);
    latency<<<1, 32>>>(d, c, 1000, 0.999999, 1e-9);
    long long h; cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
    printf("dependent DFMA latency: %.2f cycles\n", (double)h / (1000.0 * 64));
    for (int bps : {4, 6}) {
        run<0, 0>("DFMA only", bps);
        run<1, 0>("DFMA : (LOP3+IADD) = 1 : 2 instr", bps);
        run<1, 1>("DFMA : IMAD = 1 : 1", bps);
        run<2, 1>("DFMA : IMAD = 1 : 2", bps);
    }
    return 0;
}
