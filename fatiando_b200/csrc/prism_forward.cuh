// prism_forward.cuh -- the fused forward-field kernel (K1 gravity / K2 magnetic).
//
// Replaces the per-prism, per-field Python->C loop of
// fatiando/gravmag/prism.py:131-141 + _prism.pyx:265-275.  Mapping:
//   * one thread owns one observation point: its coordinates and one fp64
//     accumulator per requested component stay in registers for the whole
//     kernel, so the running sum sees prisms in list order, like res[l] does
//     in the reference;
//   * a block stages TILE prisms at a time in shared memory as SoA rows; every
//     lane then reads the same prism (shared-memory broadcast), so the prism
//     data costs ~9 doubles per TILE*~1000 FP64 instructions: the kernel is
//     FP64-pipe-bound, not memory-bound (DESIGN.md "Rooflines");
//   * blockIdx.y optionally splits the prism list into fixed chunks (small N,
//     huge M); partial sums go to scratch and reduce_partials adds them in
//     chunk order, so results are reproducible run to run.
#pragma once
#include "prism_common.cuh"

namespace fb200 {

// Eval::template pair<MASK>(acc, xp, yp, zp, x1, x2, y1, y2, z1, z2, p0, p1, p2, f)
// 4 resident blocks (16 warps) per SM = a 128-register budget.  Measured on B200:
// 3, 4, 5 or 6 blocks differ by < 2 % (the kernels are bound by instruction
// issue, not by latency hiding), 4 is marginally best for the 6-7 component sets.
#ifndef FB200_FWD_MIN_BLOCKS
#define FB200_FWD_MIN_BLOCKS 4
#endif

template <uint32_t MASK, class Eval>
__global__ void __launch_bounds__(FWD_THREADS, FB200_FWD_MIN_BLOCKS)
forward_kernel(const ForwardArgs a)
{
    constexpr int NC = popc(MASK);
    constexpr bool MAG = (MASK & MAGNETIC_MASK) != 0;
    constexpr int NP = MAG ? 3 : 1;
    __shared__ double sp[6 + NP][TILE];

    int64_t l = (int64_t)blockIdx.x * FWD_THREADS + threadIdx.x;
    const bool active = l < a.n;
    if (!active) l = a.n - 1;   // compute redundantly, never store
    const double xp = a.xp[l], yp = a.yp[l], zp = a.zp[l];
    const double f[3] = {a.fvec[0], a.fvec[1], a.fvec[2]};

    double acc[NC];
#pragma unroll
    for (int c = 0; c < NC; ++c) acc[c] = 0.0;
    if (a.init != nullptr) acc[0] = a.init[l];

    const int64_t q0 = (int64_t)blockIdx.y * a.chunk;
    const int64_t q1 = (q0 + a.chunk < a.model.m) ? q0 + a.chunk : a.model.m;
    for (int64_t t = q0; t < q1; t += TILE) {
        const int cnt = (q1 - t < TILE) ? (int)(q1 - t) : TILE;
        __syncthreads();   // previous tile fully consumed
        for (int idx = threadIdx.x; idx < cnt; idx += FWD_THREADS) {
#pragma unroll
            for (int r = 0; r < 6; ++r) sp[r][idx] = a.model.b[r][t + idx];
#pragma unroll
            for (int r = 0; r < NP; ++r)
                sp[6 + r][idx] = a.use_override ? a.override_p[r] : a.model.p[r][t + idx];
        }
        __syncthreads();
#pragma unroll 1
        for (int q = 0; q < cnt; ++q) {
            const double p1 = MAG ? sp[6 + (NP > 1 ? 1 : 0)][q] : 0.0;
            const double p2 = MAG ? sp[6 + (NP > 2 ? 2 : 0)][q] : 0.0;
            Eval::template pair<MASK>(acc, xp, yp, zp, sp[0][q], sp[1][q], sp[2][q],
                                      sp[3][q], sp[4][q], sp[5][q], sp[6][q], p1, p2, f);
        }
    }
    if (!active) return;
    if (a.nsplit == 1) {
#pragma unroll
        for (int c = 0; c < NC; ++c)
            a.out[(int64_t)a.out_row[c] * a.ld + l] = acc[c] * a.scale[c];
    } else {
        double *part = a.out + (int64_t)blockIdx.y * NC * a.ld;
#pragma unroll
        for (int c = 0; c < NC; ++c) part[(int64_t)c * a.ld + l] = acc[c];
    }
}

// Generic launcher body used by the per-group translation units.
template <uint32_t MASK, class Eval>
inline cudaError_t launch_forward_t(const ForwardArgs &a, cudaStream_t stream)
{
    dim3 grid((unsigned)((a.n + FWD_THREADS - 1) / FWD_THREADS), (unsigned)a.nsplit);
    forward_kernel<MASK, Eval><<<grid, FWD_THREADS, 0, stream>>>(a);
    return cudaGetLastError();
}

}  // namespace fb200
