// prism_jacobian.cuh -- the sensitivity-matrix (Jacobian) tile writers (K3).
//
// Replaces the per-cell loop jac[:, q] = prism.<field>(x, y, z, [cell_q], dens=1)
// of the reference's inversion paths (eqlayer.py:108-111 layout;
// harvester.py:720-724, imaging.py:116-117 call pattern).  Every entry is an
// independent 8-corner sum started from 0.0 and multiplied by the unit scale,
// exactly what one reference call returns for one cell.
//
// Two mappings, chosen so that the fp64 stores of a warp are contiguous:
//   * C-order (ndata, nparams): lane <-> prism (bounds in registers), the block
//     walks a tile of points broadcast from shared memory; a warp store covers
//     32 consecutive doubles of one row (2 full 128-byte lines).
//   * transposed (nparams, ndata), the imaging.py layout: lane <-> point,
//     prisms broadcast from shared memory.
// Stores are streaming (st.global.cs): the matrix is written once and read by
// a later consumer, it must not evict the prism/point tiles from L2.
// Arithmetic per entry ~330-1000 FP64 instructions vs 8 bytes stored: the kernel
// is FP64-pipe-bound by >10x (DESIGN.md "Rooflines"); the DRAM-write rate is
// reported as a secondary counter.
#pragma once
#include "prism_common.cuh"
#include "prism_fastmath.cuh"

namespace fb200 {

constexpr int JAC_THREADS = 128;
constexpr int JAC_TILE = 128;    // points (or prisms) staged per block

struct JacobianArgs {
    ModelView model;
    const double *xp, *yp, *zp;
    int64_t n;
    double unit_p[3];    // unit density, or unit magnetization vector
    double fvec[3];
    double scale;
    double *out;
    int64_t ld;
};

// C-order: out[l * ld + q]
template <uint32_t MASK, class Eval>
__global__ void __launch_bounds__(JAC_THREADS)
jacobian_kernel(const JacobianArgs a)
{
    static_assert(popc(MASK) == 1, "one field per Jacobian");
    __shared__ double pts[3][JAC_TILE];
    int64_t q = (int64_t)blockIdx.x * JAC_THREADS + threadIdx.x;
    const bool active = q < a.model.m;
    if (!active) q = a.model.m - 1;
    const double x1 = a.model.b[0][q], x2 = a.model.b[1][q], y1 = a.model.b[2][q],
                 y2 = a.model.b[3][q], z1 = a.model.b[4][q], z2 = a.model.b[5][q];
    const double f[3] = {a.fvec[0], a.fvec[1], a.fvec[2]};
    const double p0 = a.unit_p[0], p1 = a.unit_p[1], p2 = a.unit_p[2];
    const int64_t l0 = (int64_t)blockIdx.y * JAC_TILE;
    const int cnt = (a.n - l0 < JAC_TILE) ? (int)(a.n - l0) : JAC_TILE;
    for (int idx = threadIdx.x; idx < cnt; idx += JAC_THREADS) {
        pts[0][idx] = a.xp[l0 + idx];
        pts[1][idx] = a.yp[l0 + idx];
        pts[2][idx] = a.zp[l0 + idx];
    }
    if constexpr (Eval::USES_TABLES) fm::load_tables();
    __syncthreads();
    double *row = a.out + l0 * a.ld + q;
#pragma unroll 1
    for (int r = 0; r < cnt; ++r) {
        double acc[1] = {0.0};
        Eval::template pair<MASK>(acc, pts[0][r], pts[1][r], pts[2][r], x1, x2, y1, y2, z1, z2,
                                  p0, p1, p2, f);
        if (active) __stcs(row + (int64_t)r * a.ld, acc[0] * a.scale);
    }
}

// transposed: out[q * ld + l]
template <uint32_t MASK, class Eval>
__global__ void __launch_bounds__(JAC_THREADS)
jacobian_t_kernel(const JacobianArgs a)
{
    static_assert(popc(MASK) == 1, "one field per Jacobian");
    __shared__ double sp[6][JAC_TILE];
    int64_t l = (int64_t)blockIdx.x * JAC_THREADS + threadIdx.x;
    const bool active = l < a.n;
    if (!active) l = a.n - 1;
    const double xp = a.xp[l], yp = a.yp[l], zp = a.zp[l];
    const double f[3] = {a.fvec[0], a.fvec[1], a.fvec[2]};
    const double p0 = a.unit_p[0], p1 = a.unit_p[1], p2 = a.unit_p[2];
    const int64_t q0 = (int64_t)blockIdx.y * JAC_TILE;
    const int cnt = (a.model.m - q0 < JAC_TILE) ? (int)(a.model.m - q0) : JAC_TILE;
    for (int idx = threadIdx.x; idx < cnt; idx += JAC_THREADS)
#pragma unroll
        for (int r = 0; r < 6; ++r) sp[r][idx] = a.model.b[r][q0 + idx];
    if constexpr (Eval::USES_TABLES) fm::load_tables();
    __syncthreads();
    double *col = a.out + q0 * a.ld + l;
#pragma unroll 1
    for (int q = 0; q < cnt; ++q) {
        double acc[1] = {0.0};
        Eval::template pair<MASK>(acc, xp, yp, zp, sp[0][q], sp[1][q], sp[2][q], sp[3][q],
                                  sp[4][q], sp[5][q], p0, p1, p2, f);
        if (active) __stcs(col + (int64_t)q * a.ld, acc[0] * a.scale);
    }
}

// ---- adjoint: out[q] = sum_l d[l] * field(point l, prism q, unit property) -----
// J^T d without ever storing J -- what imaging.py:116-118 computes per layer as
// dot(sensibility_T, gz).  Lane <-> prism as in jacobian_kernel; blockIdx.y
// walks a contiguous range of point tiles; partial sums per range go to scratch
// and are added in range order by reduce_partials_kernel.
struct AdjointArgs {
    JacobianArgs j;          // model, points, unit property, fvec (scale/out/ld unused)
    const double *weights;   // d[l]
    double *part;            // [nsplit][ld_p]
    int64_t ld_p;
    int64_t rows_per_split;  // multiple of JAC_TILE
};

template <uint32_t MASK, class Eval>
__global__ void __launch_bounds__(JAC_THREADS)
adjoint_kernel(const AdjointArgs a)
{
    static_assert(popc(MASK) == 1, "one field per adjoint");
    __shared__ double pts[4][JAC_TILE];
    const ModelView &mv = a.j.model;
    int64_t q = (int64_t)blockIdx.x * JAC_THREADS + threadIdx.x;
    const bool active = q < mv.m;
    if (!active) q = mv.m - 1;
    const double x1 = mv.b[0][q], x2 = mv.b[1][q], y1 = mv.b[2][q], y2 = mv.b[3][q],
                 z1 = mv.b[4][q], z2 = mv.b[5][q];
    const double f[3] = {a.j.fvec[0], a.j.fvec[1], a.j.fvec[2]};
    const double u0 = a.j.unit_p[0], u1 = a.j.unit_p[1], u2 = a.j.unit_p[2];
    const int64_t l_begin = (int64_t)blockIdx.y * a.rows_per_split;
    const int64_t l_end = (l_begin + a.rows_per_split < a.j.n) ? l_begin + a.rows_per_split : a.j.n;
    double acc[1] = {0.0};
    if constexpr (Eval::USES_TABLES) fm::load_tables();     // the loop opens with a block barrier
    for (int64_t l0 = l_begin; l0 < l_end; l0 += JAC_TILE) {
        const int cnt = (l_end - l0 < JAC_TILE) ? (int)(l_end - l0) : JAC_TILE;
        __syncthreads();
        for (int idx = threadIdx.x; idx < cnt; idx += JAC_THREADS) {
            pts[0][idx] = a.j.xp[l0 + idx];
            pts[1][idx] = a.j.yp[l0 + idx];
            pts[2][idx] = a.j.zp[l0 + idx];
            pts[3][idx] = a.weights[l0 + idx];
        }
        __syncthreads();
#pragma unroll 1
        for (int r = 0; r < cnt; ++r) {
            const double d = pts[3][r];
            // the field is linear in the property: weight it before the evaluation
            Eval::template pair<MASK>(acc, pts[0][r], pts[1][r], pts[2][r], x1, x2, y1, y2, z1, z2,
                                      u0 * d, u1 * d, u2 * d, f);
        }
    }
    if (active) a.part[(int64_t)blockIdx.y * a.ld_p + q] = acc[0];
}

template <uint32_t MASK, class Eval>
inline cudaError_t launch_adjoint_t(const AdjointArgs &a, int nsplit, cudaStream_t stream)
{
    dim3 grid((unsigned)((a.j.model.m + JAC_THREADS - 1) / JAC_THREADS), (unsigned)nsplit);
    adjoint_kernel<MASK, Eval><<<grid, JAC_THREADS, 0, stream>>>(a);
    return cudaGetLastError();
}

template <uint32_t MASK, class Eval>
inline cudaError_t launch_jacobian_t(const JacobianArgs &a, bool transposed, cudaStream_t stream)
{
    if (!transposed) {
        dim3 grid((unsigned)((a.model.m + JAC_THREADS - 1) / JAC_THREADS),
                  (unsigned)((a.n + JAC_TILE - 1) / JAC_TILE));
        jacobian_kernel<MASK, Eval><<<grid, JAC_THREADS, 0, stream>>>(a);
    } else {
        // grid.y is limited to 65535 blocks: walk the prisms in slabs of that many tiles
        const int64_t slab = (int64_t)65535 * JAC_TILE;
        for (int64_t q0 = 0; q0 < a.model.m; q0 += slab) {
            JacobianArgs s = a;
            s.model.m = (a.model.m - q0 < slab) ? a.model.m - q0 : slab;
            for (int r = 0; r < 6; ++r) s.model.b[r] = a.model.b[r] + q0;
            s.out = a.out + q0 * a.ld;
            dim3 grid((unsigned)((a.n + JAC_THREADS - 1) / JAC_THREADS),
                      (unsigned)((s.model.m + JAC_TILE - 1) / JAC_TILE));
            jacobian_t_kernel<MASK, Eval><<<grid, JAC_THREADS, 0, stream>>>(s);
        }
    }
    return cudaGetLastError();
}

}  // namespace fb200
