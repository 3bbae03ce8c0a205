// harvest.cuh -- effect rows of the planting inversion's candidates (see harvest.cu).
#pragma once
#include "prism_common.cuh"
#include "prism_fastmath.cuh"

namespace fb200 {

constexpr int HV_THREADS = 128;
constexpr int HV_MAX_CELLS = 8;        // candidates per effect launch

struct EffectArgs {
    const double *xp, *yp, *zp;        // this data set's points
    int64_t n;
    double fvec[3];
    double scale;
    int ncells;
    double b[HV_MAX_CELLS][6];
    double p[HV_MAX_CELLS][3];         // density, or mx my mz
    int has[HV_MAX_CELLS];             // 0: the candidate lacks this data set's property
    double *out[HV_MAX_CELLS];         // effect row + data-set offset
};

template <uint32_t MASK, class Eval>
__global__ void __launch_bounds__(HV_THREADS) harvest_effect_kernel(const EffectArgs a)
{
    const int64_t l = (int64_t)blockIdx.x * HV_THREADS + threadIdx.x;
    const int c = blockIdx.y;
    if constexpr (Eval::USES_TABLES) {
        fm::load_tables();
        __syncthreads();
    }
    if (l >= a.n) return;
    double acc[1] = {0.0};
    if (a.has[c]) {
        const double f[3] = {a.fvec[0], a.fvec[1], a.fvec[2]};
        Eval::template pair<MASK>(acc, a.xp[l], a.yp[l], a.zp[l], a.b[c][0], a.b[c][1], a.b[c][2],
                                  a.b[c][3], a.b[c][4], a.b[c][5], a.p[c][0], a.p[c][1], a.p[c][2], f);
        acc[0] *= a.scale;
    }
    a.out[c][l] = acc[0];
}

template <class Eval>
cudaError_t launch_effect(int field, const EffectArgs &a, cudaStream_t s)
{
    dim3 grid((unsigned)((a.n + HV_THREADS - 1) / HV_THREADS), (unsigned)a.ncells);
    switch (field) {
#define CASE(F) case F: harvest_effect_kernel<(1u << F), Eval><<<grid, HV_THREADS, 0, s>>>(a); break;
        CASE(0) CASE(1) CASE(2) CASE(3) CASE(4) CASE(5) CASE(6) CASE(7) CASE(8) CASE(9)
        CASE(10) CASE(11) CASE(12) CASE(13)
#undef CASE
    default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}


cudaError_t launch_harvest_effect_exact(int field, const EffectArgs &a, cudaStream_t s);   // harvest_exact.cu
cudaError_t launch_harvest_effect_fast(int field, const EffectArgs &a, cudaStream_t s);    // harvest_fast.cu

}  // namespace fb200
