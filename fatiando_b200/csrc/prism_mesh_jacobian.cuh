// prism_mesh_jacobian.cuh -- sensitivity matrix of a REGULAR mesh by node sharing (K3n).
//
// Column q of the Jacobian is the field of cell q alone (eqlayer.py:108-111 layout;
// harvester.py:720-724, imaging.py:116-117 call pattern): the alternating sum of the
// per-corner kernel over the cell's eight corners (_prism.pyx:88-104).  In a regular
// mesh every interior node is a corner of eight cells, so the per-cell kernels
// (prism_jacobian.cuh) evaluate it eight times per point.  Here a block takes P
// points and walks the mesh one node ROW (fixed y) at a time: it evaluates the kernel
// once per (point, node) into shared memory -- one warp-strided slab of
// (nz+1) x (nx+1) values per point, current and previous row -- and then forms the
// nx x nz cells between the two rows as signed sums of eight shared-memory values,
// stored with consecutive lanes on consecutive cells (x fastest = the matrix's
// contiguous direction).  ~1.1 kernel evaluations per entry (~100 FP64 instructions)
// instead of the collapsed 8-corner evaluation (~320).
//
// Each entry is a plain fp64 sum of its eight corner terms, like the reference's own
// accumulation; the per-corner values come from node_eval (prism_mesh.cuh), i.e. with
// the reference's safe_log / safe_atan2 / shifted-r rules.
#pragma once
#include "prism_common.cuh"
#include "prism_mesh.cuh"

namespace fb200 {

constexpr int MJ_THREADS = 256;
constexpr int MJ_WARPS = MJ_THREADS / 32;

struct MeshJacArgs {
    const double *xn, *yn, *zn;     // node coordinates
    int nxp, nyp, nzp;              // nodes per axis (cells + 1)
    double sh[3];                   // 1e-5 * cell size
    const double *xp, *yp, *zp;
    int64_t n;
    double unit_p[3];
    double fvec[3];
    double scale;
    double *out;
    int64_t ld;
    int transposed;
    int points;                     // P: points per block, a divisor of 8
};

#if defined(__CUDACC__)
template <uint32_t MASK>
__global__ void __launch_bounds__(MJ_THREADS) mesh_jacobian_kernel(const MeshJacArgs a)
{
    static_assert(popc(MASK) == 1, "one field per Jacobian");
    extern __shared__ double sm[];
    const int P = a.points;
    const int nxp = a.nxp, nzp = a.nzp, nx = nxp - 1, nz = nzp - 1, ny = a.nyp - 1;
    const int slab = nzp * nxp;
    double *buf0 = sm, *buf1 = sm + (size_t)P * slab;
    double *xs = buf1 + (size_t)P * slab;          // node x coordinates
    double *zs = xs + nxp;                         // node z coordinates
    for (int i = threadIdx.x; i < nxp; i += MJ_THREADS) xs[i] = a.xn[i];
    for (int i = threadIdx.x; i < nzp; i += MJ_THREADS) zs[i] = a.zn[i];

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int p = warp % P;                        // this warp's point within the block
    const int sub = warp / P, nsub = MJ_WARPS / P; // its share of the slab
    const int64_t l_raw = (int64_t)blockIdx.x * P + p;
    const bool active = l_raw < a.n;
    const int64_t l = active ? l_raw : a.n - 1;    // compute redundantly, never store
    const double px = a.xp[l], py = a.yp[l], pz = a.zp[l];
    const double f[3] = {a.fvec[0], a.fvec[1], a.fvec[2]};
    const double sh[3] = {a.sh[0], a.sh[1], a.sh[2]};
    const double u0 = a.unit_p[0], u1 = a.unit_p[1], u2 = a.unit_p[2];
    __syncthreads();

    const int step = 32 * nsub;
    const int rounds = (slab + step - 1) / step;   // same trip count for every lane (warp votes inside)
    const int cells = nz * nx;
    for (int b = 0; b <= ny; ++b) {
        double *cur = ((b & 1) ? buf1 : buf0) + (size_t)p * slab;
        const double *prev = ((b & 1) ? buf0 : buf1) + (size_t)p * slab;
        const double Y = fm::sub(a.yn[b], py);
        const double YY = fm::mul(Y, Y);
        const bool ty = below_2m480(Y);
        // ---- one kernel evaluation per (point, node) of this row ----
        int e = sub * 32 + lane;
        int c = e / nxp, ia = e - c * nxp;
        for (int it = 0; it < rounds; ++it) {
            const bool live = e < slab;
            const int cc = live ? c : 0, aa = live ? ia : 0;
            const double X = fm::sub(xs[aa], px), Z = fm::sub(zs[cc], pz);
            const bool tx = below_2m480(X);
            double acc[1] = {0.0};
            node_eval<MASK>(acc, X, Y, Z, fm::add(fm::mul(X, X), YY), tx & ty, tx | ty, u0, u1, u2, f, sh);
            if (live) cur[e] = acc[0];
            e += step;
            ia += step;
            while (ia >= nxp) { ia -= nxp; ++c; }
        }
        __syncthreads();
        // ---- the cells between node rows b-1 and b: signed sum of eight node values ----
        if (b >= 1) {
            const int j = b - 1;
            int q = sub * 32 + lane;
            int k = q / nx, i = q - k * nx;
            for (; q < cells; q += step) {
                const int lo = k * nxp + i, hi = lo + nxp;            // z level k, k+1
                // corner (x2, y2, z2) carries +, every lower bound flips the sign (_prism.pyx:104)
                const double top = (cur[hi + 1] - cur[hi]) - (prev[hi + 1] - prev[hi]);
                const double bot = (cur[lo + 1] - cur[lo]) - (prev[lo + 1] - prev[lo]);
                const double v = (top - bot) * a.scale;
                const int64_t col = ((int64_t)k * ny + j) * nx + i;
                if (active) {
                    if (a.transposed) __stcs(a.out + col * a.ld + l, v);
                    else __stcs(a.out + l * a.ld + col, v);
                }
                i += step;
                while (i >= nx) { i -= nx; ++k; }
            }
        }
        __syncthreads();
    }
}

// shared memory the kernel needs for P points per block
inline size_t mesh_jacobian_smem(int nxp, int nzp, int P)
{
    return ((size_t)2 * P * nxp * nzp + nxp + nzp) * sizeof(double);
}

template <uint32_t MASK>
inline cudaError_t launch_mesh_jacobian_t(const MeshJacArgs &a, cudaStream_t stream)
{
    const size_t smem = mesh_jacobian_smem(a.nxp, a.nzp, a.points);
    cudaError_t e = cudaFuncSetAttribute(mesh_jacobian_kernel<MASK>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    const unsigned blocks = (unsigned)((a.n + a.points - 1) / a.points);
    mesh_jacobian_kernel<MASK><<<blocks, MJ_THREADS, smem, stream>>>(a);
    return cudaGetLastError();
}
#endif  // __CUDACC__

}  // namespace fb200
