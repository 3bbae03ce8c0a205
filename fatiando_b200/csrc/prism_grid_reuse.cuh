// prism_grid_reuse.cuh -- sensitivity matrix of a regular mesh on a COMMENSURATE level grid.
//
// SURVEY 8(f) row 3 / mesher/mesh.py:626-634 x gridder/point_generation.py:89-96: when the
// observation points form a regular grid at one level whose spacing divides the cell size
// (cell = kx x ky grid steps) and all coordinates are such that node - point is the same
// double for every (node a, point i) with the same u = a*kx - i (the host verifies exactly
// that, fatiando_b200/prism.py: _grid_plan), the per-corner kernel K(node - point) takes only
// (nx*kx + ni) x (ny*ky + nj) x (nz + 1) distinct values instead of points x nodes:
//
//   1. node_table_kernel   T[c][v][u] = K(X(u), Y(v), Z(c)) * unit property     (~1.7e7 values at C4a)
//   2. cell_table_kernel   D[c][dv][du] = the signed sum of the eight corners of a cell whose
//                          lower corner is at (du, dv, c) -- the same additions, in the same
//                          order, as mesh_jacobian_kernel's cell loop
//   3. expand_rows_kernel  J[(i, j)][(ck, cj, ci)] = D[ck][cj*ky + (nj-1-j)][ci*kx + (ni-1-i)]:
//                          pure data movement.  D is stored with du = e*kx + r split into
//                          (r, e), so that the nx entries of a cell row are contiguous, and twice
//                          (the second copy shifted by one double) so that every source segment
//                          is 16-byte aligned: each segment is ONE TMA bulk load into shared
//                          memory, each ~32 KB of a row of J ONE TMA bulk store.
//
// Every entry is bit-identical to what mesh_jacobian_kernel writes (same node arithmetic, same
// order of the eight additions); the build is bound by the HBM write of J.
#pragma once
#include "prism_common.cuh"
#include "prism_forward.cuh"
#include "prism_mesh.cuh"

namespace fb200 {

struct GridPlan {
    int nx, ny, nz;          // cells
    int kx, ky;              // grid steps per cell
    int64_t ni, nj;          // grid points per axis (x slowest: row l = i*nj + j)
    int64_t nu, nv;          // distinct X, Y offsets: nx*kx + ni, ny*ky + nj
    int64_t ndu, ndv;        // lower-corner offsets of cells: (nx-1)*kx + ni, (ny-1)*ky + nj
    int64_t ne;              // entries per residue row of D, even, >= ceil(ndu / kx) + 2
};

#if defined(__CUDACC__)
// thread = (u, v) column, walking down the levels; whole warps (votes inside node_prims)
template <uint32_t MASK>
__global__ void __launch_bounds__(128) node_table_kernel(GridPlan g, const double *xu, const double *yv,
                                                         const double *zc, double u0, double u1, double u2,
                                                         double f0, double f1, double f2, double s0, double s1,
                                                         double s2, double *T)
{
    static_assert(popc(MASK) == 1, "one field");
    fm::load_tables();
    __syncthreads();
    const fm::tabref_t tb = fm::tab_ref();
    const int64_t total = g.nu * g.nv;
    int64_t id = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool active = id < total;
    if (!active) id = total - 1;
    const int64_t v = id / g.nu, u = id - v * g.nu;
    const double X = xu[u], Y = yv[v];
    const double f[3] = {f0, f1, f2}, sh[3] = {s0, s1, s2};
    const double YY = fm::mul(Y, Y);
    const ColumnBits cb = column_bits(X, Y);
    for (int c = 0; c <= g.nz; ++c) {
        const double Z = zc[c];
        double acc[1] = {0.0};
        const NodePrims np = node_prims<MASK>(X, Y, Z, fm::add(fm::mul(X, X), YY), bad_coordinates(cb, Z), sh, tb);
        node_add<MASK>(acc, np, X, Y, Z, u0, u1, u2, f);
        if (active) T[((int64_t)c * g.nv + v) * g.nu + u] = acc[0];
    }
}

// D copy p (p = 0, 1): Dp[((c*ndv + dv)*kx + r)*ne + e - p] = cell value at du = e*kx + r
static __global__ void cell_table_kernel(GridPlan g, const double *T, double *D0, double *D1)
{
    const int64_t du = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t dv = blockIdx.y;
    if (du >= g.ndu) return;
    const int64_t e = du / g.kx, r = du - e * g.kx;
    const int64_t lvl = g.nv * g.nu;
    const double *lo = T + dv * g.nu + du;                 // node (a, b) of the cell's lower corner
    const double *hi = lo + (int64_t)g.ky * g.nu;          // node row b + 1
    double below = (hi[g.kx] - hi[0]) - (lo[g.kx] - lo[0]);
    for (int c = 0; c < g.nz; ++c) {
        lo += lvl;
        hi += lvl;
        const double above = (hi[g.kx] - hi[0]) - (lo[g.kx] - lo[0]);
        const double val = above - below;
        const int64_t at = (((int64_t)c * g.ndv + dv) * g.kx + r) * g.ne + e;
        D0[at] = val;
        if (e >= 1) D1[at - 1] = val;
        below = above;
    }
}

// One warp per row of J drives the TMA engine: CH source segments of nx doubles per chunk ->
// shared memory (one bulk load each, issued by the 32 lanes in turn, all completing on the
// chunk's mbarrier), then ONE bulk store of the chunk into the row.  Two chunks in flight per
// warp, several warps per SM.
constexpr int GR_CHUNK_BYTES = 16384;
static __global__ void __launch_bounds__(32) expand_rows_kernel(GridPlan g, const double *D0, const double *D1,
                                                         int64_t row0, int64_t nrows, double *out, int64_t ld)
{
    extern __shared__ __align__(128) unsigned char gr_smem[];
    __shared__ __align__(8) uint64_t bar[2];
    const int lane = threadIdx.x;
    const int64_t lrow = blockIdx.x;
    const int64_t l = row0 + lrow;
    const int64_t i = l / g.nj, j = l - i * g.nj;
    const int ox = (int)(g.ni - 1 - i), oy = (int)(g.nj - 1 - j);
    const int si = ox / g.kx, ri = ox - si * g.kx;
    const int p = si & 1;
    const double *src0 = (p ? D1 : D0) + (int64_t)ri * g.ne + (si - p);  // + ((c*ndv + dv)*kx)*ne per segment
    const int64_t seg_stride = (int64_t)g.kx * g.ne;                    // per dv
    const unsigned seg_bytes = (unsigned)g.nx * 8u;
    const int nseg = g.nz * g.ny;
    const int ch = GR_CHUNK_BYTES / (int)seg_bytes;                     // segments per chunk (>= 1, host-checked)
    double *dst_row = out + lrow * ld;
    if (lane == 0) {
        mbar_init(&bar[0], 1);
        mbar_init(&bar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    const int nchunks = (nseg + ch - 1) / ch;
    auto load = [&](int k) {
        const int s0 = k * ch, cnt = (nseg - s0 < ch) ? nseg - s0 : ch;
        unsigned char *buf = gr_smem + (size_t)(k & 1) * GR_CHUNK_BYTES;
        if (lane == 0) mbar_expect_tx(&bar[k & 1], (unsigned)cnt * seg_bytes);
        __syncwarp();
        for (int s = lane; s < cnt; s += 32) {
            const int seg = s0 + s;
            const int ck = seg / g.ny, cj = seg - ck * g.ny;
            const int64_t dv = (int64_t)cj * g.ky + oy;
            tma_load_1d(buf + (size_t)s * seg_bytes, src0 + ((int64_t)ck * g.ndv + dv) * seg_stride, seg_bytes,
                        &bar[k & 1]);
        }
    };
    load(0);
    for (int k = 0; k < nchunks; ++k) {
        if (k + 1 < nchunks) {
            // the other buffer is free once the store issued from it (chunk k-1) has read it
            if (k >= 1 && lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            __syncwarp();
            load(k + 1);
        }
        mbar_wait(&bar[k & 1], (unsigned)((k >> 1) & 1));
        if (lane == 0) {
            const int s0 = k * ch, cnt = (nseg - s0 < ch) ? nseg - s0 : ch;
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_row + (int64_t)s0 * g.nx),
                         "r"(smem_u32(gr_smem + (size_t)(k & 1) * GR_CHUNK_BYTES)), "r"((unsigned)cnt * seg_bytes)
                         : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}
#endif  // __CUDACC__

}  // namespace fb200
