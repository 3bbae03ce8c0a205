// prism_mesh_jacobian.cuh -- sensitivity matrix of a REGULAR mesh by node sharing (K3n).
//
// Column q of the Jacobian is the field of cell q alone (eqlayer.py:108-111 layout;
// harvester.py:720-724, imaging.py:116-117 call pattern): the alternating sum of the
// per-corner kernel over the cell's eight corners (_prism.pyx:88-104).  In a regular
// mesh every interior node is a corner of eight cells, so the per-cell kernels
// (prism_jacobian.cuh) evaluate it eight times per point.  Here a block takes P
// points (4 by default) and walks the mesh one node ROW (fixed y) at a time: it evaluates the kernel
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
#ifndef FB200_MJ_UNROLL
#define FB200_MJ_UNROLL 2
#endif
constexpr int MJ_UNROLL = FB200_MJ_UNROLL;
constexpr int MJ_WARPS = MJ_THREADS / 32;
constexpr int MJ_ZPAD = MJ_THREADS + 8;   // z levels the node loop may read past the last one

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
    int points;                     // P: points per block, a divisor of 8 (8 / P warps per point)
};

#ifndef FB200_MJ_MIN_BLOCKS
#define FB200_MJ_MIN_BLOCKS 3
#endif
#ifndef FB200_MJ_UNROLL
#define FB200_MJ_UNROLL 2
#endif
#if defined(__CUDACC__)
// 3 resident blocks of 256 threads per SM (85 registers): node and cell phases of different
// blocks overlap.  All shared-memory traffic of the two loops goes through explicit 32-bit
// shared-window addresses held in registers (fm::pin_u32), advanced by adds -- the index
// arithmetic the compiler generates for the (level, x) walk was a third of the instruction stream.
template <uint32_t MASK, bool TRANSPOSED>
__global__ void __launch_bounds__(MJ_THREADS, FB200_MJ_MIN_BLOCKS) mesh_jacobian_kernel(const MeshJacArgs a)
{
    static_assert(popc(MASK) == 1, "one field per Jacobian");
    extern __shared__ __align__(16) double sm[];
    const int P = a.points;
    const int nxp = a.nxp, nzp = a.nzp, nx = nxp - 1, nz = nzp - 1, ny = a.nyp - 1;
    const int stride = nxp + (nxp & 1);            // slab rows padded to an even length: 16-byte rows
    const int slab = nzp * stride;
    const int64_t ld = a.ld;
    const double *const yn = a.yn;
    double *buf0 = sm, *buf1 = sm + (size_t)P * slab;
    double *xs = buf1 + (size_t)P * slab;          // node x coordinates
    double *zs = xs + nxp;                         // node z coordinates
    for (int i = threadIdx.x; i < nxp; i += MJ_THREADS) xs[i] = a.xn[i];
    for (int i = threadIdx.x; i < nzp + MJ_ZPAD; i += MJ_THREADS) zs[i] = a.zn[i < nzp ? i : nzp - 1];
    fm::load_tables();

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int p = warp % P;                        // this warp's point within the block
    const int sub = warp / P, nsub = MJ_WARPS / P; // its share of the slab
    const int64_t l_raw = (int64_t)blockIdx.x * P + p;
    const bool active = l_raw < a.n;
    const int64_t l = active ? l_raw : a.n - 1;    // compute redundantly, never store
    const double px = a.xp[l], py = a.yp[l], pz = a.zp[l];
    const double f[3] = {a.fvec[0], a.fvec[1], a.fvec[2]};
    const double sh[3] = {a.sh[0], a.sh[1], a.sh[2]};
    // the unit conversion rides on the unit property: every node value is already scaled
    const double u0 = a.unit_p[0] * a.scale, u1 = a.unit_p[1] * a.scale, u2 = a.unit_p[2] * a.scale;
    // C-order: entry (l, col) at out[l*ld + col]; transposed: out[col*ld + l]
    double *const row_base = TRANSPOSED ? a.out + l : a.out + l * ld;
    __syncthreads();
    const fm::tabref_t tb = fm::tab_ref_pinned();
    const unsigned xs_a = fm::pin_u32(smem_u32(xs)), zs_a = fm::pin_u32(smem_u32(zs));
    // (byte strides pinned too: derived from kernel parameters they would be re-formed from
    // uniform registers at every use)
    const unsigned row_b = fm::pin_u32(8u * (unsigned)nxp), stride_b = fm::pin_u32(8u * (unsigned)stride);
    const unsigned xs_end = xs_a + row_b, zs_end = zs_a + 8u * (unsigned)nzp;
    const unsigned pad_b = stride_b - row_b;
    // two adjacent entries go out as one 16-byte store when the matrix allows it
    const bool vec_store = !TRANSPOSED && (nx % 2 == 0) && (ld % 2 == 0) &&
                           (reinterpret_cast<uintptr_t>(a.out) % 16 == 0);

    const int step = 32 * nsub;
    const int rounds = (nzp * nxp + step - 1) / step;   // same trip count for every lane (warp votes inside)
    const int dq = step / nxp, dr = step - dq * nxp;        // one step in (level, x index) form
    const unsigned dq_b = fm::pin_u32(8u * (unsigned)dq), dr_b = fm::pin_u32(8u * (unsigned)dr);
    const unsigned step_b = fm::pin_u32(8u * (unsigned)(dq * stride + dr));     // one step in the padded slab
    unsigned cur = fm::pin_u32(smem_u32(buf0 + (size_t)p * slab)), prev = fm::pin_u32(smem_u32(buf1 + (size_t)p * slab));
    const int e0 = sub * 32 + lane;
    const int c0 = e0 / nxp, a0 = e0 - c0 * nxp;
    for (int b = 0; b <= ny; ++b) {
        const double Y = fm::sub(yn[b], py);
        const double YY = fm::mul(Y, Y);
        const bool ty = below_2m480(Y);            // warp-uniform: the warp has ONE point
        // ---- one kernel evaluation per (point, node) of this row ----
        // (lanes past the end of the slab keep computing -- the warp votes inside node_prims
        // need them -- on the padded tail of zs, and do not store)
        unsigned ea = cur + 8u * (unsigned)(c0 * stride + a0), za = zs_a + 8u * (unsigned)c0,
                 xa = xs_a + 8u * (unsigned)a0;
#pragma unroll(MJ_UNROLL)
        for (int it = 0; it < rounds; ++it) {
            const double X = fm::sub(fm::lds_f64(xa), px), Z = fm::sub(fm::lds_f64(za), pz);
            // X and Z both below 2^-480 (with Y there too, rare and warp-uniform, the whole row
            // takes the careful variant: same bits, only slower)
            const int ex = fm::hi32(X) & 0x7ff00000, ez = fm::hi32(Z) & 0x7ff00000;
            const bool bad_coord = ty | ((ex > ez ? ex : ez) < 0x21f00000);
            double acc[1] = {0.0};
            const NodePrims np = node_prims<MASK>(X, Y, Z, fm::add(fm::mul(X, X), YY), bad_coord, sh, tb);
            node_add<MASK>(acc, np, X, Y, Z, u0, u1, u2, f);
            if (za < zs_end) fm::sts_f64(ea, acc[0]);
            ea += step_b;
            xa += dr_b;
            za += dq_b;
            if (xa >= xs_end) { xa -= row_b; za += 8u; ea += pad_b; }
        }
        __syncthreads();
        // ---- the cells between node rows b-1 and b ----
        // lane <-> cell column i, walking down the z levels: d(k) = the x- and y-difference of
        // the four node columns around the cell at level k, entry(k) = d(k+1) - d(k);
        // corner (x2, y2, z2) carries +, every lower bound flips the sign (_prism.pyx:104)
        if (b >= 1 && active) {
            // a lane takes TWO adjacent cell columns (i0, i0 + 1): three node values per row and
            // level (one 16-byte + one 8-byte load) instead of two per cell, one 16-byte store
            const int j = b - 1;
            const int64_t kstride = (int64_t)ny * nx * (TRANSPOSED ? ld : 1);
            const int npairs = (nx + 1) >> 1;
            for (int t = sub * 32 + lane; t < npairs; t += step) {
                const int i0 = 2 * t;
                const bool two = i0 + 1 < nx;
                unsigned ca = cur + 16u * (unsigned)t, pa = prev + 16u * (unsigned)t;
                double *dst = TRANSPOSED ? row_base + ((int64_t)j * nx + i0) * ld
                                         : row_base + ((int64_t)j * nx + i0);
                double2 c = fm::lds_f64x2_v(ca), p = fm::lds_f64x2_v(pa);
                double c2 = fm::lds_f64_v16(ca), p2 = fm::lds_f64_v16(pa);
                double below0 = (c.y - c.x) - (p.y - p.x), below1 = (c2 - c.y) - (p2 - p.y);
#pragma unroll 2
                for (int k = 0; k < nz; ++k) {
                    ca += stride_b;
                    pa += stride_b;
                    c = fm::lds_f64x2_v(ca); p = fm::lds_f64x2_v(pa);
                    c2 = fm::lds_f64_v16(ca); p2 = fm::lds_f64_v16(pa);
                    const double above0 = (c.y - c.x) - (p.y - p.x), above1 = (c2 - c.y) - (p2 - p.y);
                    const double e0v = above0 - below0, e1v = above1 - below1;
                    if (TRANSPOSED) {
                        __stcs(dst, e0v);
                        if (two) __stcs(dst + ld, e1v);
                    } else if (vec_store) {
                        __stcs(reinterpret_cast<double2 *>(dst), make_double2(e0v, e1v));
                    } else {
                        __stcs(dst, e0v);
                        if (two) __stcs(dst + 1, e1v);
                    }
                    dst += kstride;
                    below0 = above0;
                    below1 = above1;
                }
            }
        }
        __syncthreads();
        const unsigned t = cur; cur = prev; prev = t;
    }
}

// shared memory the kernel needs for P points per block
inline size_t mesh_jacobian_smem(int nxp, int nzp, int P)
{
    return ((size_t)2 * P * (nxp + (nxp & 1)) * nzp + nxp + nzp + MJ_ZPAD + 2) * sizeof(double);
}

template <uint32_t MASK>
inline cudaError_t launch_mesh_jacobian_t(const MeshJacArgs &a, cudaStream_t stream)
{
    const size_t smem = mesh_jacobian_smem(a.nxp, a.nzp, a.points);
    const unsigned blocks = (unsigned)((a.n + a.points - 1) / a.points);
    cudaError_t e;
    if (a.transposed) {
        e = cudaFuncSetAttribute(mesh_jacobian_kernel<MASK, true>,
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        mesh_jacobian_kernel<MASK, true><<<blocks, MJ_THREADS, smem, stream>>>(a);
    } else {
        e = cudaFuncSetAttribute(mesh_jacobian_kernel<MASK, false>,
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        mesh_jacobian_kernel<MASK, false><<<blocks, MJ_THREADS, smem, stream>>>(a);
    }
    return cudaGetLastError();
}
#endif  // __CUDACC__

}  // namespace fb200
