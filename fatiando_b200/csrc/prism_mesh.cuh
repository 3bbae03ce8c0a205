// prism_mesh.cuh -- node-sharing evaluation for REGULAR meshes (PrismMesh).
//
// In a regular mesh every interior node is a corner of eight cells.  The
// reference evaluates it eight times, once per cell and with alternating sign
// (fatiando/gravmag/_prism.pyx:88-104 inside the per-prism loop of
// prism.py:131-141); summed over the mesh that is
//
//     field(P) = sum_nodes K(node - P) * W_node,
//     W_node   = sum over the <= 8 adjacent cells of (+-) property
//
// with K the reference's per-corner kernel (_prism.pyx:36-68) and the sign that
// of the corner the node is in that cell.  W is a 2x2x2 difference stencil of
// the property field, built once on the host (fatiando_b200/flatten.py).  One
// kernel evaluation per NODE instead of eight per CELL: per cell-equivalent the
// six tensor components cost 191 FP64 instructions instead of 428 (ncu, profiles/),
// and nodes whose weight is zero (uniform regions, masked cells) are skipped outright.
//
// The arithmetic per node is the reference's own: r = sqrt((dx*dx + dy*dy) +
// dz*dz) with the same association and IEEE sqrt, safe_log / safe_atan2
// semantics (log and atan evaluated by the fast-path primitives of
// prism_fastmath.cuh), the 1e-5 shift of r for gxy/gxz/gyz.  What differs from
// the per-cell path is (i) a shared node uses ONE coordinate where the
// reference has x1 + dx of one cell and b0 + dx*(i+1) of the next (equal or one
// ulp apart) and (ii) the order of the additions -- both far inside the parity
// floor (tests: tests/test_hostsim.py, tests/test_gpu_parity.py).  A one-ulp
// coordinate change matters in one situation only: a point lying exactly IN a
// wall whose two roundings differ.  There the reference's two cells see a zero
// and a +-1-ulp atan2 denominator, their direction-dependent terms need not
// cancel, and its answer depends on that rounding (by multiples of pi * rho).
// The flattener lists such walls and calls with a point in one of them are
// evaluated by the per-cell kernels (fb200_model_set_hazard_planes).  Points
// inside the body on a line through mesh nodes sit on a genuine discontinuity
// of the field; there any two roundings disagree, the reference's own cells
// with each other included.
#pragma once
#include "prism_common.cuh"
#include "prism_fastmath.cuh"
#include "prism_exact.cuh"

namespace fb200 {

// log(a) for a >= 0 with safe_log(0) = 0 (_prism.pyx:28-34); denormal arguments (outside the
// table logarithm's range) go to the library.  Only on the careful paths.
FB_HD double node_log(double a, fm::tabref_t tb)
{
    if (fm::hi32(a) < 0x00100000) return (a == 0.0) ? 0.0 : log(a);
    return fm::log_tab(a, tb);
}

// |safe_atan2(y, x)| = atan(|y| / |x|), 0 for y == 0 (_prism.pyx:16-26); ay = |y|, ax = |x|
FB_HD double node_atan_abs(double ay, double ax, fm::tabref_t tb)
{
    // (0, 0) -> 0: give the division a non-zero denominator
    const double xs = (ax == 0.0 && ay == 0.0) ? 1.0 : ax;
    return fm::atan_tab(ay, xs, tb);
}

FB_HD double xor_sign(double a, int sg) { return fm::mk(fm::hi32(a) ^ sg, fm::lo32(a)); }

// the six per-corner primitives of one node, plus the logs gxy / gxz / gyz use (equal to
// Lz / Ly / Lx unless the reference shifts r, _prism.pyx:327-330, :359-362, :418-421).
// The three arctangents of a node, atan(ZY / Xr), atan(ZX / Yr), atan(XY / Zr), all have the
// sign of X*Y*Z: they are evaluated on absolute values and `sg` (0 or 0x80000000) is that sign.
struct NodePrims {
    double Lx, Ly, Lz, Axx, Ayy, Azz, Lxy, Lxz, Lyz;
    int sg;
};

// with every special case of the reference; out of line, rare
template <uint32_t MASK>
#if defined(__CUDACC__)
__device__ __noinline__
#else
static
#endif
NodePrims node_prims_careful(double X, double Y, double Z, double XX_YY, double s0, double s1, double s2,
                             fm::tabref_t tb)
{
    using namespace fm;
    NodePrims p = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    const double ZZ = mul(Z, Z);
    const double r2 = add(XX_YY, ZZ);
    double r = root(r2);
    if (hi32(r2) < 0x03500000) r = sqrt(r2);      // outside the range of the branch-free root
    if constexpr (MASK & NEED_LX) p.Lx = node_log(add(X, r), tb);
    if constexpr (MASK & NEED_LY) p.Ly = node_log(add(Y, r), tb);
    if constexpr (MASK & NEED_LZ) p.Lz = node_log(add(Z, r), tb);
    if constexpr (MASK & NEED_AXX) p.Axx = node_atan_abs(mul(fabs(Z), fabs(Y)), mul(fabs(X), r), tb);
    if constexpr (MASK & NEED_AYY) p.Ayy = node_atan_abs(mul(fabs(Z), fabs(X)), mul(fabs(Y), r), tb);
    if constexpr (MASK & NEED_AZZ) p.Azz = node_atan_abs(mul(fabs(X), fabs(Y)), mul(fabs(Z), r), tb);
    p.sg = (hi32(X) ^ hi32(Y) ^ hi32(Z)) & (int)0x80000000;
    p.Lxy = p.Lz; p.Lxz = p.Ly; p.Lyz = p.Lx;
    if constexpr (MASK & bit(F_GXY))
        if (X == 0.0 && Y == 0.0 && Z < 0.0)
            p.Lxy = safe_log_ref(add(Z, sqrt(add(add(mul(s0, s0), mul(s1, s1)), ZZ))));
    if constexpr (MASK & bit(F_GXZ))
        if (X == 0.0 && Z == 0.0 && Y < 0.0)
            p.Lxz = safe_log_ref(add(Y, sqrt(add(add(mul(s0, s0), mul(s2, s2)), mul(Y, Y)))));
    if constexpr (MASK & bit(F_GYZ))
        if (Y == 0.0 && Z == 0.0 && X < 0.0)
            p.Lyz = safe_log_ref(add(X, sqrt(add(add(mul(s1, s1), mul(s2, s2)), mul(X, X)))));
    return p;
}

// ---- hot / careful split ---------------------------------------------------------
// Every special case of the reference (safe_log(0), safe_atan2(0, 0), the shifted r,
// a root argument below the range of the branch-free root) needs coordinate
// differences that are zero (or below 2^-480) on TWO axes at once, or a log argument
// that rounds to zero.  node_prims tests exactly that with exponent compares on the
// ALU pipe; if ANY lane of the warp has such a node (warp vote: no divergence) the
// warp takes the out-of-line careful variant above, otherwise the straight-line code
// below, which carries no compares, selects or branches for them.  Both give the same
// bits for an ordinary node, so a point's result does not depend on its warp-mates.
FB_HD bool below_2m480(double d) { return (fm::hi32(d) & 0x7ff00000) < 0x21f00000; }

// The coordinate part of the test, "X and Y tiny, or Z and one of them", is carried by two
// integers per node COLUMN (column_bits): zmask = 0 if X and Y are both below 2^-480, else the
// exponent mask; zor = 0 if one of them is, else the exponent mask.  Then
// (hi32(Z) & zmask) | zor is below the exponent of 2^-480 exactly in that case: one LOP3 and
// one compare per node.
FB_HD int imin(int a, int b) { return a < b ? a : b; }
struct ColumnBits { int zmask, zor; };
FB_HD ColumnBits column_bits(double X, double Y)
{
    const bool tx = below_2m480(X), ty = below_2m480(Y);
    ColumnBits c;
    c.zmask = (tx & ty) ? 0 : 0x7ff00000;
    c.zor = (tx | ty) ? 0 : 0x7ff00000;
    return c;
}

FB_HD bool bad_coordinates(ColumnBits cb, double Z)
{
    return ((fm::hi32(Z) & cb.zmask) | cb.zor) < 0x21f00000;
}

// bad_coord: X and Y both below 2^-480, or Z and one of them (bad_coordinates, or the caller's own test)
template <uint32_t MASK>
FB_HD NodePrims node_prims(double X, double Y, double Z, double XX_YY, bool bad_coord,
                           const double (&sh)[3], fm::tabref_t tb)
{
    using namespace fm;
    const double ZZ = mul(Z, Z);
    const double r = root(add(XX_YY, ZZ));
    double ax = 1.0, ay = 1.0, az = 1.0;
    bool bad = bad_coord;
    // X + r >= +0: zero (or denormal) iff the high word is below the smallest normal's
    int lowest = 0x7fffffff;
    if constexpr (MASK & NEED_LX) { ax = add(X, r); lowest = imin(lowest, hi32(ax)); }
    if constexpr (MASK & NEED_LY) { ay = add(Y, r); lowest = imin(lowest, hi32(ay)); }
    if constexpr (MASK & NEED_LZ) { az = add(Z, r); lowest = imin(lowest, hi32(az)); }
    bad = bad | (lowest < 0x00100000);
#if defined(__CUDA_ARCH__)
    bad = __any_sync(0xffffffffu, bad);
#endif
    if (bad) return node_prims_careful<MASK>(X, Y, Z, XX_YY, sh[0], sh[1], sh[2], tb);
    NodePrims p = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    if constexpr (MASK & NEED_LX) p.Lx = log_tab(ax, tb);
    if constexpr (MASK & NEED_LY) p.Ly = log_tab(ay, tb);
    if constexpr (MASK & NEED_LZ) p.Lz = log_tab(az, tb);
    if constexpr (MASK & NEED_AXX) p.Axx = atan_tab(mul(fabs(Z), fabs(Y)), mul(fabs(X), r), tb);
    if constexpr (MASK & NEED_AYY) p.Ayy = atan_tab(mul(fabs(Z), fabs(X)), mul(fabs(Y), r), tb);
    if constexpr (MASK & NEED_AZZ) p.Azz = atan_tab(mul(fabs(X), fabs(Y)), mul(fabs(Z), r), tb);
    p.sg = (hi32(X) ^ hi32(Y) ^ hi32(Z)) & (int)0x80000000;
    p.Lxy = p.Lz; p.Lxz = p.Ly; p.Lyz = p.Lx;
    (void)ax; (void)ay; (void)az;
    return p;
}

// Adds one node's contribution to every requested component (kernel table of
// _prism.pyx:36-68; w0..w2 = node weight: density stencil, or the three magnetization
// stencils).  Explicit fma / separately rounded products: the same bits on every path.
template <uint32_t MASK>
FB_HD void node_add(double (&acc)[popc(MASK)], const NodePrims &p, double X, double Y, double Z,
                    double w0, double w1, double w2, const double (&f)[3])
{
    using namespace fm;
    constexpr uint32_t MIXED = bit(F_POT) | bit(F_GX) | bit(F_GY) | bit(F_GZ) | MAGNETIC_MASK;
    // signed arctangents where a field mixes them with other terms; the pure-arctangent
    // fields gxx / gyy / gzz flip the sign of the weight instead (one XOR for all three)
    double Axx = p.Axx, Ayy = p.Ayy, Azz = p.Azz;
    if constexpr (MASK & MIXED) {
        Axx = xor_sign(p.Axx, p.sg); Ayy = xor_sign(p.Ayy, p.sg); Azz = xor_sign(p.Azz, p.sg);
    }
    const double ws = xor_sign(w0, p.sg);
    int c = 0;
    if constexpr (MASK & bit(F_POT)) {   // _prism.pyx:36-39
        const double logs = fma(mul(X, Y), p.Lz, fma(mul(Y, Z), p.Lx, mul(mul(X, Z), p.Ly)));
        const double atans = fma(mul(X, X), Axx, fma(mul(Y, Y), Ayy, mul(mul(Z, Z), Azz)));
        acc[c] = fma(fma(-0.5, atans, logs), w0, acc[c]);
        ++c;
    }
    if constexpr (MASK & bit(F_GX)) {    // :43-44
        acc[c] = fma(fma(X, Axx, -fma(Y, p.Lz, mul(Z, p.Ly))), w0, acc[c]);
        ++c;
    }
    if constexpr (MASK & bit(F_GY)) {    // :46-47
        acc[c] = fma(fma(Y, Ayy, -fma(Z, p.Lx, mul(X, p.Lz))), w0, acc[c]);
        ++c;
    }
    if constexpr (MASK & bit(F_GZ)) {    // :49-50
        acc[c] = fma(fma(Z, Azz, -fma(X, p.Ly, mul(Y, p.Lx))), w0, acc[c]);
        ++c;
    }
    if constexpr (MASK & bit(F_GXX)) { acc[c] = fma(-p.Axx, ws, acc[c]); ++c; }   // :52-53
    if constexpr (MASK & bit(F_GXY)) { acc[c] = fma(p.Lxy, w0, acc[c]); ++c; }    // :55-56
    if constexpr (MASK & bit(F_GXZ)) { acc[c] = fma(p.Lxz, w0, acc[c]); ++c; }    // :58-59
    if constexpr (MASK & bit(F_GYY)) { acc[c] = fma(-p.Ayy, ws, acc[c]); ++c; }   // :61-62
    if constexpr (MASK & bit(F_GYZ)) { acc[c] = fma(p.Lyz, w0, acc[c]); ++c; }    // :64-65
    if constexpr (MASK & bit(F_GZZ)) { acc[c] = fma(-p.Azz, ws, acc[c]); ++c; }   // :67-68
    if constexpr (MASK & MAGNETIC_MASK) {
        // v1..v6 of _prism.pyx:94-99 at this node (unshifted); w0..w2 = stencils of mx, my, mz
        const double bx = fma(-Axx, w0, fma(p.Lz, w1, mul(p.Ly, w2)));
        const double by = fma(p.Lz, w0, fma(-Ayy, w1, mul(p.Lx, w2)));
        const double bz = fma(p.Ly, w0, fma(p.Lx, w1, mul(-Azz, w2)));
        if constexpr (MASK & bit(F_TF)) { acc[c] = add(acc[c], fma(f[0], bx, fma(f[1], by, mul(f[2], bz)))); ++c; }
        if constexpr (MASK & bit(F_BX)) { acc[c] = add(acc[c], bx); ++c; }
        if constexpr (MASK & bit(F_BY)) { acc[c] = add(acc[c], by); ++c; }
        if constexpr (MASK & bit(F_BZ)) { acc[c] = add(acc[c], bz); ++c; }
    }
    (void)ws; (void)Axx; (void)Ayy; (void)Azz;
}

// one node, hot path with the careful variant chosen by warp vote
template <uint32_t MASK>
FB_HD void node_eval(double (&acc)[popc(MASK)], double X, double Y, double Z, double XX_YY,
                     ColumnBits cb, double w0, double w1, double w2,
                     const double (&f)[3], const double (&sh)[3], fm::tabref_t tb)
{
    const NodePrims p = node_prims<MASK>(X, Y, Z, XX_YY, bad_coordinates(cb, Z), sh, tb);
    node_add<MASK>(acc, p, X, Y, Z, w0, w1, w2, f);
}

// one node with every special case of the reference applied inline (no warp vote): the
// careful paths of the relief surface form (prism_face.cuh) and its merged nodes
//   sh[3] : 1e-5 * cell size in x, y, z (the singularity shift, _prism.pyx:327-330)
template <uint32_t MASK>
FB_HD void node_accumulate(double (&acc)[popc(MASK)], double X, double Y, double Z,
                           double XX_YY, double w0, double w1, double w2,
                           const double (&f)[3], const double (&sh)[3])
{
    const NodePrims p = node_prims_careful<MASK>(X, Y, Z, XX_YY, sh[0], sh[1], sh[2], fm::tab_ref());
    node_add<MASK>(acc, p, X, Y, Z, w0, w1, w2, f);
}

#if defined(__CUDACC__)
// ---- the kernel ---------------------------------------------------------------
// thread = observation point, as in prism_forward.cuh.  Nodes are walked with z
// fastest, so X, Y and X*X + Y*Y are recomputed only once per node column; the
// node weights (1 or 3 rows) come through the same two-stage TMA/mbarrier ring as
// the prism tiles of the per-cell kernel; the node coordinate vectors (a few
// hundred doubles) are read through L1 (x, y: once per column) or shared memory (z).
struct MeshView {
    const double *xn, *yn, *zn;   // node coordinates, nxp, nyp, nzp entries
    int nxp, nyp, nzp;
    const double *w[3];           // node weights, index (b*nxp + a)*nzp + c
    int64_t nnodes;
    double sh[3];                 // 1e-5 * cell size (singularity shift)
};

struct MeshArgs {
    MeshView mesh;
    const double *xp, *yp, *zp;
    int64_t n;
    double fvec[3];
    int64_t chunk;                // nodes per blockIdx.y, multiple of TILE
    int nsplit;
    double *out;
    int64_t ld;
    double scale[F_COUNT];
    int out_row[F_COUNT];
};

constexpr int MESH_ZMAX = 2048;   // z node levels kept in shared memory

template <uint32_t MASK>
__global__ void __launch_bounds__(FWD_THREADS, FB200_FWD_MIN_BLOCKS)
mesh_forward_kernel(const MeshArgs a)
{
    constexpr int NC = popc(MASK);
    constexpr bool MAG = (MASK & MAGNETIC_MASK) != 0;
    constexpr int ROWS = MAG ? 3 : 1;
    __shared__ __align__(128) double sp[FWD_STAGES][ROWS][TILE];
    __shared__ __align__(8) uint64_t full_bar[FWD_STAGES];
    __shared__ unsigned done_cnt[FWD_STAGES];
    extern __shared__ double zn_s[];

    const MeshView &mv = a.mesh;
    int64_t l = (int64_t)blockIdx.x * FWD_THREADS + threadIdx.x;
    const bool active = l < a.n;
    if (!active) l = a.n - 1;
    const double xp = a.xp[l], yp = a.yp[l], zp = a.zp[l];
    const double f[3] = {a.fvec[0], a.fvec[1], a.fvec[2]};
    const double sh[3] = {mv.sh[0], mv.sh[1], mv.sh[2]};
    double acc[NC];
#pragma unroll
    for (int c = 0; c < NC; ++c) acc[c] = 0.0;

    const int64_t g0 = (int64_t)blockIdx.y * a.chunk;
    const int64_t g1 = (g0 + a.chunk < mv.nnodes) ? g0 + a.chunk : mv.nnodes;
    const int ntiles = (int)((g1 - g0 + TILE - 1) / TILE);

    auto issue = [&](int t, int s) {
        const int64_t off = g0 + (int64_t)t * TILE;
        mbar_expect_tx(&full_bar[s], (unsigned)(ROWS * TILE * sizeof(double)));
#pragma unroll
        for (int r = 0; r < ROWS; ++r)
            tma_load_1d(&sp[s][r][0], mv.w[r] + off, TILE * sizeof(double), &full_bar[s]);
    };

    if (threadIdx.x == 0) {
#pragma unroll
        for (int s = 0; s < FWD_STAGES; ++s) {
            mbar_init(&full_bar[s], 1);
            done_cnt[s] = 0;
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // the z levels live in shared memory (the host takes the per-cell kernels for meshes
    // with more than MESH_ZMAX of them)
    for (int i = threadIdx.x; i < mv.nzp; i += FWD_THREADS) zn_s[i] = mv.zn[i];
    fm::load_tables();
    __syncthreads();
    // shared-window addresses held in registers (see fm::tab_ref_pinned)
    const fm::tabref_t tb = fm::tab_ref_pinned();
    const unsigned zbase = fm::pin_u32(smem_u32(zn_s)), zend = zbase + 8u * (unsigned)mv.nzp;
    if (threadIdx.x == 0) {
        for (int s = 0; s < FWD_STAGES && s < ntiles; ++s) issue(s, s);
    }

    // position of the first node of this chunk: column (a_, b_) and level c_
    int64_t col = g0 / mv.nzp;
    int c_ = (int)(g0 - col * mv.nzp);
    int b_ = (int)(col / mv.nxp);
    int a_ = (int)(col - (int64_t)b_ * mv.nxp);
    double X = fm::sub(mv.xn[a_], xp), Y = fm::sub(mv.yn[b_ < mv.nyp ? b_ : mv.nyp - 1], yp);
    double sxy = fm::add(fm::mul(X, X), fm::mul(Y, Y));
    ColumnBits cb = column_bits(X, Y);
    unsigned zaddr = zbase + 8u * (unsigned)c_;

    for (int t = 0; t < ntiles; ++t) {
        const int s = t % FWD_STAGES;
        mbar_wait(&full_bar[s], (unsigned)((t / FWD_STAGES) & 1));
        const int64_t base = g0 + (int64_t)t * TILE;
        const int cnt = (g1 - base < TILE) ? (int)(g1 - base) : TILE;
#pragma unroll 1
        for (int q = 0; q < cnt; ++q) {
            const double w0 = sp[s][0][q];
            const double w1 = MAG ? sp[s][ROWS > 1 ? 1 : 0][q] : 0.0;
            const double w2 = MAG ? sp[s][ROWS > 2 ? 2 : 0][q] : 0.0;
            // zero weight (uniform region, masked cells): nothing to add -- uniform branch
            const bool live = MAG ? (w0 != 0.0 || w1 != 0.0 || w2 != 0.0) : (w0 != 0.0);
            if (live) {
                const double Z = fm::sub(fm::lds_f64(zaddr), zp);
                node_eval<MASK>(acc, X, Y, Z, sxy, cb, w0, w1, w2, f, sh, tb);
            }
            zaddr += 8u;
            if (zaddr == zend) {      // next node column: new X (and Y at the end of a row)
                zaddr = zbase;
                if (++a_ == mv.nxp) {
                    a_ = 0;
                    ++b_;
                    if (b_ < mv.nyp) Y = fm::sub(mv.yn[b_], yp);
                }
                X = fm::sub(mv.xn[a_], xp);
                sxy = fm::add(fm::mul(X, X), fm::mul(Y, Y));
                cb = column_bits(X, Y);
            }
        }
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

template <uint32_t MASK>
inline cudaError_t launch_mesh_t(const MeshArgs &a, cudaStream_t stream)
{
    dim3 grid((unsigned)((a.n + FWD_THREADS - 1) / FWD_THREADS), (unsigned)a.nsplit);
    const size_t zbytes = (size_t)a.mesh.nzp * sizeof(double);
    mesh_forward_kernel<MASK><<<grid, FWD_THREADS, zbytes, stream>>>(a);
    return cudaGetLastError();
}
#endif  // __CUDACC__

}  // namespace fb200
