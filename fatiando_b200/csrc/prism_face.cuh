// prism_face.cuh -- surface form of a relief: top faces + reference-plane nodes.
//
// A PrismRelief (mesher/mesh.py:391-492) is a carpet of prisms that all end on
// the same reference level: cell c spans [zc, ref] (or [ref, zc]) and, once the
// sign rule of addprop (:484-488) is undone, contributes
//
//     rho_c * ( G_c(ref) - G_c(zc) ),
//     G_c(z) = sum_{i,j} (-1)^(i+j) K(X_i, Y_j, z - zp)      (4 corners of one face)
//
// with K the reference's per-corner kernel (_prism.pyx:36-68; index 0 <-> the "2"
// bound, :82-92).  The reference evaluates all 8 corners of every cell.  Here
//   * the TOP faces are evaluated per cell with the collapsed transcendental sums
//     of the fast path (prism_fastmath.cuh) restricted to four corners: gz needs
//     4 sqrt + 4 log_ratio + 1 arctangent instead of 8 + 4 + 2;
//   * the REFERENCE-PLANE corners of neighbouring cells coincide when the relief
//     is a tight regular grid, so they are merged into nodes with weight = the
//     signed 2x2 sum of rho over the adjacent cells (flatten.relief_surface) --
//     for the usual uniform density every interior node cancels and only the rim
//     of the grid is left.
// Both are plain re-groupings of the reference's own sum (same summands, same
// argument roundings); the parity tests hold them to the same floor as the
// per-cell path.
//
// Both evaluators plug into forward_kernel (prism_forward.cuh) through the usual
// policy signature; what the nine staged rows mean:
//   EvalFace: x1, x2, y1, y2, z_face, z_other, w           (w = -rho_c)
//   EvalNode: xn, shx, yn, shy, zn, (unused), w            (sh = 1e-5 * cell size)
#pragma once
#include "prism_common.cuh"
#include "prism_fastmath.cuh"
#include "prism_exact.cuh"
#include "prism_fast.cuh"
#include "prism_mesh.cuh"

namespace fb200 {

enum FaceLog : int { FL_NONE = 0, FL_ALL, FL_BYI, FL_BYJ, FL_FULL };

constexpr FaceLog face_log(bool full, bool byi, bool byj, bool all)
{
    return full || (byi && byj) ? FL_FULL : byi ? FL_BYI : byj ? FL_BYJ : all ? FL_ALL : FL_NONE;
}

// corner index of a face: c = j*2 + i
//   FL_ALL : v[0] = sum_ij s_ij log a_ij
//   FL_BYI : v[i] = log a_i0 - log a_i1            (i = 0, 1)
//   FL_BYJ : v[j] = log a_0j - log a_1j            (j = 0, 1)
//   FL_FULL: v[c] = log a_c
template <FaceLog MODE>
struct FaceLogFam {
    static constexpr int NG = MODE == FL_FULL ? 4 : (MODE == FL_BYI || MODE == FL_BYJ) ? 2 : MODE == FL_ALL ? 1 : 0;
    double num[NG > 0 ? NG : 1], den[NG > 0 ? NG : 1], v[NG > 0 ? NG : 1];

    FB_HD_M void products(const double (&a)[4])
    {
        if constexpr (MODE == FL_ALL) {
            num[0] = a[0] * a[3];
            den[0] = a[1] * a[2];
        } else if constexpr (MODE == FL_BYI) {
            num[0] = a[0]; den[0] = a[2];
            num[1] = a[1]; den[1] = a[3];
        } else if constexpr (MODE == FL_BYJ) {
            num[0] = a[0]; den[0] = a[1];
            num[1] = a[2]; den[1] = a[3];
        } else if constexpr (MODE == FL_FULL) {
#pragma unroll
            for (int c = 0; c < 4; ++c) { num[c] = a[c]; den[c] = 1.0; }
        }
    }
    FB_HD_M bool any_zero() const
    {
        bool z = false;
#pragma unroll
        for (int g = 0; g < NG; ++g) z = z | (num[g] * den[g] == 0.0);
        return z;
    }
    FB_HD_M void finish()
    {
#pragma unroll
        for (int g = 0; g < NG; ++g) v[g] = fm::log_ratio(num[g], den[g]);
    }
    FB_HD_M double all() const
    {
        static_assert(MODE != FL_NONE, "family not evaluated");
        if constexpr (MODE == FL_ALL) return v[0];
        else if constexpr (MODE == FL_FULL) return (v[0] - v[1]) - (v[2] - v[3]);
        else return v[0] - v[1];
    }
    // sum_i (-1)^i w[i] * (L_i0 - L_i1)
    FB_HD_M double by_i(const double (&w)[2]) const
    {
        static_assert(MODE == FL_BYI || MODE == FL_FULL, "needs i groups");
        if constexpr (MODE == FL_BYI) return w[0] * v[0] - w[1] * v[1];
        else return w[0] * (v[0] - v[2]) - w[1] * (v[1] - v[3]);
    }
    // sum_j (-1)^j w[j] * (L_0j - L_1j)
    FB_HD_M double by_j(const double (&w)[2]) const
    {
        static_assert(MODE == FL_BYJ || MODE == FL_FULL, "needs j groups");
        if constexpr (MODE == FL_BYJ) return w[0] * v[0] - w[1] * v[1];
        else return w[0] * (v[0] - v[1]) - w[1] * (v[2] - v[3]);
    }
    // sum_ij s_ij w[j][i] * L_ij
    FB_HD_M double full(const double (&w)[2][2]) const
    {
        static_assert(MODE == FL_FULL, "needs the four logarithms");
        return (w[0][0] * v[0] - w[0][1] * v[1]) - (w[1][0] * v[2] - w[1][1] * v[3]);
    }
};

template <uint32_t MASK>
#if defined(__CUDACC__)
__device__ __noinline__
#else
static
#endif
Contribution<popc(MASK)> face_careful(double xp, double yp, double zp, double x1, double x2,
                                      double y1, double y2, double zf, double zo, double w);

// One top face.  Returns false -- having produced nothing -- when the face needs
// the reference's special cases (the caller then takes face_careful).
template <uint32_t MASK>
FB_HD bool face_core(double (&out)[popc(MASK)], double xp, double yp, double zp, double x1,
                     double x2, double y1, double y2, double zf, double w)
{
    using namespace fm;
    static_assert((MASK & MAGNETIC_MASK) == 0, "surface form: gravity fields only");
    constexpr bool has_pot = MASK & bit(F_POT), has_gx = MASK & bit(F_GX),
                   has_gy = MASK & bit(F_GY), has_gz = MASK & bit(F_GZ),
                   has_gxx = MASK & bit(F_GXX), has_gxy = MASK & bit(F_GXY),
                   has_gxz = MASK & bit(F_GXZ), has_gyy = MASK & bit(F_GYY),
                   has_gyz = MASK & bit(F_GYZ), has_gzz = MASK & bit(F_GZZ);
    constexpr FaceLog MLZ = face_log(has_pot, has_gy, has_gx, has_gxy);     // weights X_i | Y_j
    constexpr FaceLog MLY = face_log(false, has_pot || has_gz, false, has_gx || has_gxz);
    constexpr FaceLog MLX = face_log(false, false, has_pot || has_gz, has_gy || has_gyz);
    constexpr bool AX_BY = has_pot || has_gx, AX_ANY = AX_BY || has_gxx;
    constexpr bool AY_BY = has_pot || has_gy, AY_ANY = AY_BY || has_gyy;
    constexpr bool AZ_ANY = has_pot || has_gz || has_gzz;

    const double X[2] = {sub(x2, xp), sub(x1, xp)};
    const double Y[2] = {sub(y2, yp), sub(y1, yp)};
    const double Z = sub(zf, zp);
    const double XX[2] = {mul(X[0], X[0]), mul(X[1], X[1])};
    const double YY[2] = {mul(Y[0], Y[0]), mul(Y[1], Y[1])};
    const double ZZ = mul(Z, Z);
    const bool sx = tiny_or_zero(X[0]) | tiny_or_zero(X[1]);
    const bool sy = tiny_or_zero(Y[0]) | tiny_or_zero(Y[1]);
    const bool sz = tiny_or_zero(Z);
    bool bad = (sx & sy) | (sx & sz) | (sy & sz);

    double r[4], ax[4], ay[4], az[4], rex[4], rey[4], rez[4];
#pragma unroll
    for (int j = 0; j < 2; ++j)
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const int c = j * 2 + i;
            r[c] = root(add(add(XX[i], YY[j]), ZZ));      // (dx**2 + dy**2) + dz**2
            if constexpr (MLX != FL_NONE) ax[c] = add(X[i], r[c]);
            if constexpr (MLY != FL_NONE) ay[c] = add(Y[j], r[c]);
            if constexpr (MLZ != FL_NONE) az[c] = add(Z, r[c]);
            if constexpr (AX_ANY) rex[c] = mul(X[i], r[c]);
            if constexpr (AY_ANY) rey[c] = mul(Y[j], r[c]);
            if constexpr (AZ_ANY) rez[c] = mul(Z, r[c]);
        }
    FaceLogFam<MLX> Lx;
    FaceLogFam<MLY> Ly;
    FaceLogFam<MLZ> Lz;
    if constexpr (MLX != FL_NONE) { Lx.products(ax); bad = bad | Lx.any_zero(); }
    if constexpr (MLY != FL_NONE) { Ly.products(ay); bad = bad | Ly.any_zero(); }
    if constexpr (MLZ != FL_NONE) { Lz.products(az); bad = bad | Lz.any_zero(); }
#if defined(__CUDA_ARCH__)
    bad = __any_sync(0xffffffffu, bad);      // the warp never diverges
#endif
    if (bad) return false;

    if constexpr (MLX != FL_NONE) Lx.finish();
    if constexpr (MLY != FL_NONE) Ly.finish();
    if constexpr (MLZ != FL_NONE) Lz.finish();

    // arctangent sums: sum_j (-1)^j Axx_ij for each i, or over all four corners
    const int hx[2] = {hi32(X[0]), hi32(X[1])};
    const int hy[2] = {hi32(Y[0]), hi32(Y[1])};
    const int hz = hi32(Z);
    double ax_i[2] = {0.0, 0.0}, ax_all = 0.0, ay_j[2] = {0.0, 0.0}, ay_all = 0.0, az_all = 0.0;
    if constexpr (AX_ANY) {
        const double im[2] = {mul(Z, Y[0]), mul(Z, Y[1])};       // z*y (_prism.pyx:53)
        if constexpr (AX_BY) {
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                Wind wd = wind_start<false>(rex[i], im[0]);
                wind_mul<true>(wd, rex[2 + i], im[1], hx[i]);
                ax_i[i] = wind_angle(wd);
            }
            ax_all = ax_i[0] - ax_i[1];
        } else {
            Wind wd = wind_start<false>(rex[0], im[0]);
            wind_mul<true>(wd, rex[2], im[1], hx[0]);
            wind_mul<true>(wd, rex[1], im[0], hx[1]);
            wind_mul<false>(wd, rex[3], im[1], hx[1]);
            ax_all = wind_angle(wd);
        }
    }
    if constexpr (AY_ANY) {
        const double im[2] = {mul(Z, X[0]), mul(Z, X[1])};       // z*x (_prism.pyx:62)
        if constexpr (AY_BY) {
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                Wind wd = wind_start<false>(rey[2 * j], im[0]);
                wind_mul<true>(wd, rey[2 * j + 1], im[1], hy[j]);
                ay_j[j] = wind_angle(wd);
            }
            ay_all = ay_j[0] - ay_j[1];
        } else {
            Wind wd = wind_start<false>(rey[0], im[0]);
            wind_mul<true>(wd, rey[1], im[1], hy[0]);
            wind_mul<true>(wd, rey[2], im[0], hy[1]);
            wind_mul<false>(wd, rey[3], im[1], hy[1]);
            ay_all = wind_angle(wd);
        }
    }
    if constexpr (AZ_ANY) {
        Wind wd = wind_start<false>(rez[0], mul(X[0], Y[0]));    // x*y (_prism.pyx:68)
        wind_mul<true>(wd, rez[1], mul(X[1], Y[0]), hz);
        wind_mul<true>(wd, rez[2], mul(X[0], Y[1]), hz);
        wind_mul<false>(wd, rez[3], mul(X[1], Y[1]), hz);
        az_all = wind_angle(wd);
    }

    int c = 0;
    if constexpr (has_pot) {      // _prism.pyx:36-39
        const double wxy[2][2] = {{X[0] * Y[0], X[1] * Y[0]}, {X[0] * Y[1], X[1] * Y[1]}};   // [j][i]
        const double hxx[2] = {0.5 * XX[0], 0.5 * XX[1]};
        const double hyy[2] = {0.5 * YY[0], 0.5 * YY[1]};
        out[c++] = (((Lz.full(wxy) + Z * Lx.by_j(Y)) + Z * Ly.by_i(X))
                    - (((hxx[0] * ax_i[0] - hxx[1] * ax_i[1]) + (hyy[0] * ay_j[0] - hyy[1] * ay_j[1]))
                       + (0.5 * ZZ) * az_all)) * w;
    }
    if constexpr (has_gx)         // :43-44
        out[c++] = ((X[0] * ax_i[0] - X[1] * ax_i[1]) - (Lz.by_j(Y) + Z * Ly.all())) * w;
    if constexpr (has_gy)         // :46-47
        out[c++] = ((Y[0] * ay_j[0] - Y[1] * ay_j[1]) - (Z * Lx.all() + Lz.by_i(X))) * w;
    if constexpr (has_gz)         // :49-50
        out[c++] = (Z * az_all - (Ly.by_i(X) + Lx.by_j(Y))) * w;
    if constexpr (has_gxx) out[c++] = -ax_all * w;     // :52-53
    if constexpr (has_gxy) out[c++] = Lz.all() * w;    // :55-56
    if constexpr (has_gxz) out[c++] = Ly.all() * w;    // :58-59
    if constexpr (has_gyy) out[c++] = -ay_all * w;     // :61-62
    if constexpr (has_gyz) out[c++] = Lx.all() * w;    // :64-65
    if constexpr (has_gzz) out[c++] = -az_all * w;     // :67-68
    (void)ax_all; (void)ay_all; (void)az_all; (void)hx; (void)hy; (void)hz;
    return true;
}

// the four corners one by one with the reference's special cases (node_accumulate of
// prism_mesh.cuh: safe_log(0) = 0, safe_atan2(0, 0) = 0, the 1e-5 shift of gxy/gxz/gyz
// with THIS cell's sizes)
template <uint32_t MASK>
#if defined(__CUDACC__)
__device__ __noinline__
#else
static
#endif
Contribution<popc(MASK)> face_careful(double xp, double yp, double zp, double x1, double x2,
                                      double y1, double y2, double zf, double zo, double w)
{
    using namespace fm;
    Contribution<popc(MASK)> res;
#pragma unroll
    for (int c = 0; c < popc(MASK); ++c) res.v[c] = 0.0;
    const double X[2] = {add(sub(x2, xp), 0.0), add(sub(x1, xp), 0.0)};
    const double Y[2] = {add(sub(y2, yp), 0.0), add(sub(y1, yp), 0.0)};
    const double Z = add(sub(zf, zp), 0.0);
    const double f[3] = {0.0, 0.0, 0.0};
    const double sh[3] = {mul(0.00001, sub(x2, x1)), mul(0.00001, sub(y2, y1)),
                          mul(0.00001, sub(zo, zf))};
#pragma unroll
    for (int j = 0; j < 2; ++j)
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const double wc = ((i + j) & 1) ? -w : w;
            node_accumulate<MASK>(res.v, X[i], Y[j], Z, add(mul(X[i], X[i]), mul(Y[j], Y[j])),
                                  wc, 0.0, 0.0, f, sh);
        }
    return res;
}

#ifndef FB200_FORCE_CAREFUL
#define FB200_FORCE_CAREFUL 0
#endif
struct EvalFace {
    static constexpr bool NEEDS_FLAG = false;
    static constexpr bool USES_TABLES = true;
    template <uint32_t MASK>
    static FB_HD_M void pair(double (&acc)[popc(MASK)], double xp, double yp, double zp,
                           double x1, double x2, double y1, double y2, double zf, double zo,
                           double w, double, double, const double (&)[3])
    {
        constexpr int NC = popc(MASK);
        double out[NC];
        if (FB200_FORCE_CAREFUL || !face_core<MASK>(out, xp, yp, zp, x1, x2, y1, y2, zf, w)) {
            const Contribution<NC> res = face_careful<MASK>(xp, yp, zp, x1, x2, y1, y2, zf, zo, w);
#pragma unroll
            for (int c = 0; c < NC; ++c) out[c] = res.v[c];
        }
#pragma unroll
        for (int c = 0; c < NC; ++c) acc[c] += out[c];
    }
};

// A merged reference-plane node.  The 1e-5 shift of gxz / gyz involves the
// THICKNESS of the cell the corner belongs to (_prism.pyx:359-362, :418-421),
// which a merged node no longer has: when a point meets that case (it lies in
// the reference plane, on a grid line, beyond the node) the kernel raises
// *flag and the host re-evaluates the call cell by cell.
struct EvalNode {
    static constexpr bool NEEDS_FLAG = true;
    static constexpr bool USES_TABLES = true;
    template <uint32_t MASK>
    static FB_HD_M void pair_flag(double (&acc)[popc(MASK)], double xp, double yp, double zp,
                                double xn, double shx, double yn, double shy, double zn, double,
                                double w, double, double, const double (&f)[3], int *flag)
    {
        using namespace fm;
        const double X = add(sub(xn, xp), 0.0), Y = add(sub(yn, yp), 0.0), Z = add(sub(zn, zp), 0.0);
        if constexpr ((MASK & (bit(F_GXZ) | bit(F_GYZ))) != 0) {
            if (Z == 0.0 && ((X == 0.0 && Y < 0.0) || (Y == 0.0 && X < 0.0))) *flag = 1;
        }
        const double sh[3] = {shx, shy, 0.0};
        node_eval<MASK>(acc, X, Y, Z, add(mul(X, X), mul(Y, Y)), column_bits(X, Y), w, 0.0, 0.0, f, sh, fm::tab_ref());
    }
};

// The adjoint of a regular mesh by node sharing: the thread is a mesh NODE, the staged
// rows are observation points with weights w = data[l] * unit property,
//     V(node) = sum_l K(node - point_l) * w_l,
// and J^T d of a cell is the signed sum of V over its eight corners (prism_abi.cu).
// Rows: xpt, shx, ypt, shy, zpt, shz, w0 [, w1, w2]   (sh = 1e-5 * cell size).
struct EvalNodeRev {
    static constexpr bool NEEDS_FLAG = false;
    static constexpr bool USES_TABLES = true;
    template <uint32_t MASK>
    static FB_HD_M void pair(double (&acc)[popc(MASK)], double xn, double yn, double zn,
                           double xpt, double shx, double ypt, double shy, double zpt, double shz,
                           double w0, double w1, double w2, const double (&f)[3])
    {
        using namespace fm;
        const double X = add(sub(xn, xpt), 0.0), Y = add(sub(yn, ypt), 0.0), Z = add(sub(zn, zpt), 0.0);
        const double sh[3] = {shx, shy, shz};
        node_eval<MASK>(acc, X, Y, Z, add(mul(X, X), mul(Y, Y)), column_bits(X, Y), w0, w1, w2, f, sh, fm::tab_ref());
    }
};

}  // namespace fb200
