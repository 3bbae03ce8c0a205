// fp64_pipe_probe.cu -- what does the B200 FP64 pipe sustain for different operand mixes?
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a tools/fp64_pipe_probe.cu -o gpurun_out/fp64_probe
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(256) probe(double *out, int iters, double a, double b)
{
    double x[8], y[8], z[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) { x[k] = threadIdx.x * 1e-3 + k; y[k] = 1.0 + 1e-9 * (threadIdx.x + k); z[k] = 1e-9 * k; }
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 16; ++u) {
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                if (MODE == 0) x[k] = fma(x[k], a, b);                 // 1 register + 2 uniform operands
                if (MODE == 1) x[k] = fma(x[k], y[k], z[k]);           // 3 distinct register operands
                if (MODE == 2) x[k] = fma(x[k], y[k], b);              // 2 registers + 1 uniform
                if (MODE == 3) x[k] = __dmul_rn(x[k], y[k]);           // DMUL 2 registers
                if (MODE == 4) x[k] = __dadd_rn(x[k], y[k]);           // DADD 2 registers
                if (MODE == 5) x[k] = fma(x[k], x[k], z[k]);           // 2 registers (one repeated)
                if (MODE == 6) { x[k] = fma(x[k], y[k], z[k]); y[k] = fma(y[k], z[k], x[(k + 1) & 7]); }  // mixed deps
            }
        }
    }
    double s = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) s += x[k] + y[k];
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
void run(const char *name, int blocks, int threads, double per_iter_ops)
{
    double *d;
    cudaMalloc(&d, (size_t)blocks * threads * 8);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 1024;
    probe<MODE><<<blocks, threads>>>(d, iters, 0.999999, 1e-9);
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 5; ++r) {
        cudaEventRecord(e0);
        probe<MODE><<<blocks, threads>>>(d, iters, 0.999999, 1e-9);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    const double inst = (double)blocks * threads * iters * 16.0 * 8.0 * per_iter_ops;
    printf("%-44s blocks/SM %d x %d thr: %.3f Tinst-lanes/s (%.1f%% of 148*64*1.965e9)\n", name,
           blocks / 148, threads, inst / (best * 1e-3) / 1e12, 100.0 * inst / (best * 1e-3) / (148.0 * 64 * 1.965e9));
    cudaFree(d);
}

int main()
{
    for (int bps : {2, 4, 8}) {
        run<0>("DFMA reg,uniform,uniform", 148 * bps, 256, 1);
        run<1>("DFMA reg,reg,reg", 148 * bps, 256, 1);
        run<2>("DFMA reg,reg,uniform", 148 * bps, 256, 1);
        run<3>("DMUL reg,reg", 148 * bps, 256, 1);
        run<4>("DADD reg,reg", 148 * bps, 256, 1);
        run<5>("DFMA reg,same reg,reg", 148 * bps, 256, 1);
        run<6>("DFMA reg,reg,reg cross-dependent x2", 148 * bps, 256, 2);
    }
    return 0;
}
