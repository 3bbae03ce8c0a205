// prism_forward.cuh -- the fused forward-field kernel (K1 gravity / K2 magnetic).
//
// Replaces the per-prism, per-field Python->C loop of
// fatiando/gravmag/prism.py:131-141 + _prism.pyx:265-275.  Mapping:
//   * one thread owns one observation point: its coordinates and one fp64
//     accumulator per requested component stay in registers for the whole
//     kernel, so the running sum sees prisms in list order, like res[l] does
//     in the reference;
//   * a block stages TILE prisms at a time in shared memory as SoA rows, by TMA
//     bulk copies into a two-stage ring guarded by mbarriers; every lane then
//     reads the same prism (shared-memory broadcast), so the prism data costs
//     ~9 doubles per ~430 FP64 instructions: the kernel is FP64-pipe-bound, not
//     memory-bound (DESIGN.md "Rooflines");
//   * blockIdx.y optionally splits the prism list into fixed chunks (small N,
//     huge M); partial sums go to scratch and reduce_partials adds them in
//     chunk order, so results are reproducible run to run.
#pragma once
#include "prism_common.cuh"
#include "prism_fastmath.cuh"

namespace fb200 {

// Eval::template pair<MASK>(acc, xp, yp, zp, x1, x2, y1, y2, z1, z2, p0, p1, p2, f)
// 4 resident blocks (16 warps) per SM = a 128-register budget.  Measured on B200:
// 3, 4, 5 or 6 blocks differ by < 2 % (the kernels are bound by instruction
// issue, not by latency hiding), 4 is marginally best for the 6-7 component sets.
#ifndef FB200_FWD_MIN_BLOCKS
#define FB200_FWD_MIN_BLOCKS 4
#endif

// ---- TMA-staged prism tiles --------------------------------------------------
// Two shared-memory stages.  A stage is filled by 1-D bulk copies
// (cp.async.bulk.shared::cluster.global, one per SoA row, 2 KB each) that
// complete on the stage's mbarrier; every thread waits on that barrier's
// parity, never on the other warps.  When a warp is done with a stage it bumps
// a counter; the LAST warp out issues the refill with the tile two ahead, so
// no warp ever waits for a slower one ("last one out refills"): the block has
// no __syncthreads in its main loop.
constexpr int FWD_STAGES = 2;
constexpr int FWD_WARPS = FWD_THREADS / 32;

__device__ __forceinline__ unsigned smem_u32(const void *p)
{
    return (unsigned)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, unsigned parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "FB200_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra FB200_DONE;\n"
        "bra FB200_WAIT;\n"
        "FB200_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tma_load_1d(void *dst, const void *src, unsigned bytes, uint64_t *bar)
{
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_u32(dst)),
        "l"(src), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}

// Eval::template pair<MASK>(acc, xp, yp, zp, x1, x2, y1, y2, z1, z2, p0, p1, p2, f)
// Component sets whose evaluators keep more than ~60 doubles live (all ten gravity fields, the
// potential with its nine products, gx+gy+gz) get 168 registers (3 blocks per SM) instead of
// spilling at 128; measured neutral to +3 % for them, and the local-memory traffic is gone.
constexpr int fwd_min_blocks(uint32_t mask)
{
    return (mask == 0x3FFu || mask == 0x3FEu || mask == 0x001u || mask == 0x00Eu) ? 3 : FB200_FWD_MIN_BLOCKS;
}

template <uint32_t MASK, class Eval>
__global__ void __launch_bounds__(FWD_THREADS, fwd_min_blocks(MASK))
forward_kernel(const ForwardArgs a)
{
    constexpr int NC = popc(MASK);
    constexpr bool MAG = (MASK & MAGNETIC_MASK) != 0;
    constexpr int NP = MAG ? 3 : 1;
    constexpr int ROWS = 6 + NP;
    __shared__ __align__(128) double sp[FWD_STAGES][ROWS][TILE];
    __shared__ __align__(8) uint64_t full_bar[FWD_STAGES];
    __shared__ unsigned done_cnt[FWD_STAGES];

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
    const int ntiles = (int)((q1 - q0 + TILE - 1) / TILE);
    // property rows are not staged when a dens= / pmag= override replaces them
    const int rows = a.use_override ? 6 : ROWS;

    // the model arrays are padded to a multiple of TILE (prism_abi.cu: upload),
    // so a full tile can always be copied, 16-byte aligned, even at the tail
    auto issue = [&](int t, int s) {
        const int64_t off = q0 + (int64_t)t * TILE;
        mbar_expect_tx(&full_bar[s], (unsigned)(rows * TILE * sizeof(double)));
#pragma unroll
        for (int r = 0; r < 6; ++r)
            tma_load_1d(&sp[s][r][0], a.model.b[r] + off, TILE * sizeof(double), &full_bar[s]);
        if (!a.use_override) {
#pragma unroll
            for (int r = 0; r < NP; ++r)
                tma_load_1d(&sp[s][6 + r][0], a.model.p[r] + off, TILE * sizeof(double), &full_bar[s]);
        }
    };

    if (threadIdx.x == 0) {
#pragma unroll
        for (int s = 0; s < FWD_STAGES; ++s) {
            mbar_init(&full_bar[s], 1);
            done_cnt[s] = 0;
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if constexpr (Eval::USES_TABLES) fm::load_tables();
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int s = 0; s < FWD_STAGES && s < ntiles; ++s) issue(s, s);
    }

    for (int t = 0; t < ntiles; ++t) {
        const int s = t % FWD_STAGES;
        mbar_wait(&full_bar[s], (unsigned)((t / FWD_STAGES) & 1));
        const int64_t base = q0 + (int64_t)t * TILE;
        const int cnt = (q1 - base < TILE) ? (int)(q1 - base) : TILE;
#pragma unroll 1
        for (int q = 0; q < cnt; ++q) {
            double p0, p1 = 0.0, p2 = 0.0;
            if (a.use_override) {
                p0 = a.override_p[0];
                if (MAG) { p1 = a.override_p[1]; p2 = a.override_p[2]; }
            } else {
                p0 = sp[s][6][q];
                if (MAG) { p1 = sp[s][6 + (NP > 1 ? 1 : 0)][q]; p2 = sp[s][6 + (NP > 2 ? 2 : 0)][q]; }
            }
            if constexpr (Eval::NEEDS_FLAG)
                Eval::template pair_flag<MASK>(acc, xp, yp, zp, sp[s][0][q], sp[s][1][q], sp[s][2][q],
                                               sp[s][3][q], sp[s][4][q], sp[s][5][q], p0, p1, p2, f,
                                               a.flag + l);
            else
                Eval::template pair<MASK>(acc, xp, yp, zp, sp[s][0][q], sp[s][1][q], sp[s][2][q],
                                          sp[s][3][q], sp[s][4][q], sp[s][5][q], p0, p1, p2, f);
        }
        // release the stage; the last warp out refills it with the tile two ahead
        __syncwarp();
        if ((threadIdx.x & 31) == 0) {
            __threadfence_block();
            const unsigned before = atomicAdd(&done_cnt[s], 1u);
            if (before == FWD_WARPS - 1) {
                done_cnt[s] = 0;
                if (t + FWD_STAGES < ntiles) {
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    issue(t + FWD_STAGES, s);
                }
            }
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
// grid_y = 0: one block row per prism chunk (a.nsplit); otherwise the caller's own
// count (several launches sharing one set of partial sums, prism_abi.cu: run_set_surface)
template <uint32_t MASK, class Eval>
inline cudaError_t launch_forward_t(const ForwardArgs &a, cudaStream_t stream, int grid_y = 0)
{
    dim3 grid((unsigned)((a.n + FWD_THREADS - 1) / FWD_THREADS), (unsigned)(grid_y ? grid_y : a.nsplit));
    forward_kernel<MASK, Eval><<<grid, FWD_THREADS, 0, stream>>>(a);
    return cudaGetLastError();
}

}  // namespace fb200
