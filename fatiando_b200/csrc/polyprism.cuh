// polyprism.cuh -- the polygonal-prism (Plouff 1976) forward model as an evaluator policy.
//
// Sibling of the prism path: fatiando/gravmag/polyprism.py sums,
// per prism and per polygon EDGE, gz (:215-308), the second derivatives kernelxx ..
// kernelzz of the 1/r integral (:563-1076) and, from those, gxx..gzz (:311-560) and
// tf / bx / by / bz (:19-212).  An edge (vertex k -> vertex k+1, z1, z2) is one staged
// row: x[k], x[k+1], y[k], y[k+1], z1, z2, property -- the flattener expands every prism
// into its edges -- so the model plugs into forward_kernel like the prisms do, one thread
// per observation point, rows through the TMA ring.
//
// The arithmetic is the reference's, expression by expression, INCLUDING its
// `dummy = 1e-10` guards, which differ from kernel to kernel (added to deltax/deltay, p
// and d for xx, xy, xz, yy, yz; to dist only for zz; inside the logs for gz): parity with
// the reference near degenerate geometry depends on them.  sqrt and the divisions are IEEE;
// log / atan2 of ordinary arguments go through the fast-path primitives (pp_log, pp_atan2
// below), anything else through the library.  The translation unit is compiled with -fmad=false so that every product
// and sum rounds separately, like numpy.  All requested components share the edge's
// radii, distances and arctangent differences.
#pragma once
#include <cmath>
#include "prism_common.cuh"
#include "prism_fastmath.cuh"

namespace fb200 {

// log and atan2 of ordinary arguments by the table-driven primitives (prism_fastmath.cuh:
// log_tab without a division, atan_tab with one; 0.3 ulp rms), the library's for everything else (zeros, denormals,
// infinities, NaN, negative log arguments) so that the reference's behaviour in degenerate
// geometry -- which its `dummy` guards are there for -- is untouched.
FB_HD bool pp_ordinary(double v)          // finite, normal, non-zero
{
    const unsigned e = (unsigned)fm::hi32(v) & 0x7ff00000u;
    return e - 0x00100000u < 0x7fe00000u;
}

// Roots and quotients.  Both switches were MEASURED and are off (profiles/r02_siblings.txt):
// the range guards the reference's degenerate-geometry behaviour needs around fm::root /
// fm::div_fast cost more issue slots than the shorter sequences save -- with both on, gz ran
// 2 % and the tensor / magnetic sets 6 % SLOWER than with the library's sqrt() and IEEE
// divisions (whose own slow paths sit behind one predicated call each).
#ifndef FB200_PP_SQRT
#define FB200_PP_SQRT 0        // 1: fm::root on its range (same bits as sqrt()), sqrt() outside
#endif
#ifndef FB200_PP_DIV
#define FB200_PP_DIV 0         // 1: fm::div_fast for mid-range operands, IEEE division outside
#endif
FB_HD double pp_sqrt(double a)
{
#if FB200_PP_SQRT
    if ((unsigned)fm::hi32(a) - 0x03500000u < 0x7ca00000u) return fm::root(a);
#endif
    return sqrt(a);
}

// a / b: reciprocal seed + Newton step + residual correction (fm::div_fast, within an ulp of
// the IEEE quotient) when both exponents lie in [2^-511, 2^512) -- no overflow, underflow or
// special value can occur on that path; the IEEE division for everything else (zero
// numerators, infinities, NaN, the extremes)
FB_HD bool pp_midrange(double v)
{
    return ((unsigned)fm::hi32(v) & 0x7ff00000u) - 0x20000000u < 0x40000000u;
}
FB_HD double pp_div(double a, double b)
{
#if FB200_PP_DIV
    if (pp_midrange(a) && pp_midrange(b)) return fm::div_fast(a, b);
#endif
    return a / b;
}

FB_HD double pp_log(double x)
{
    if (x > 0.0 && pp_ordinary(x)) return fm::log_tab(x);        // division-free, 10 FP64 instructions
    return log(x);
}

FB_HD double pp_atan2(double y, double x)
{
    if (pp_ordinary(y) && pp_ordinary(x)) {
        const double r = fm::atan_ratio_tab(y, x);             // atan(y/x) in [-pi/2, pi/2], table-reduced
        if (x > 0.0) return r;
        const double s = (y < 0.0) ? -1.0 : 1.0;               // x < 0: +-pi on top
        return fma(s, fm::konst(4), fma(s, fm::konst(5), r));
    }
    return atan2(y, x);
}

struct EvalPolyEdge {
    static constexpr bool NEEDS_FLAG = false;
    static constexpr bool USES_TABLES = true;      // log / atan tables of prism_fastmath.cuh
    template <uint32_t MASK>
    static FB_HD_M void pair(double (&acc)[popc(MASK)], double xp, double yp, double zp,
                           double vx1, double vx2, double vy1, double vy2, double z1, double z2,
                           double p0, double p1, double p2, const double (&f)[3])
    {
        constexpr double d = 1e-10;                     // the reference's `dummy`
        constexpr bool has_gz = MASK & bit(F_GZ);
        constexpr bool mag = (MASK & MAGNETIC_MASK) != 0;
        constexpr bool has_tf = MASK & bit(F_TF), has_bx = MASK & bit(F_BX),
                       has_by = MASK & bit(F_BY), has_bz = MASK & bit(F_BZ);
        constexpr bool need_xx = (MASK & bit(F_GXX)) || has_tf || has_bx;
        constexpr bool need_xy = (MASK & bit(F_GXY)) || has_tf || has_bx || has_by;
        constexpr bool need_xz = (MASK & bit(F_GXZ)) || has_tf || has_bx || has_bz;
        constexpr bool need_yy = (MASK & bit(F_GYY)) || has_tf || has_by;
        constexpr bool need_yz = (MASK & bit(F_GYZ)) || has_tf || has_by || has_bz;
        constexpr bool need_zz = (MASK & bit(F_GZZ)) || has_tf || has_bz;
        constexpr bool need_a = need_xx || need_xy || need_yy;      // arctangent family
        constexpr bool need_l = need_xz || need_yz;                 // (R - d)/(R + d) log family

        const double X1 = vx1 - xp, Y1 = vy1 - yp, X2 = vx2 - xp, Y2 = vy2 - yp;
        const double Z1 = z1 - zp, Z2 = z2 - zp;
        const double Z1_sqr = Z1 * Z1, Z2_sqr = Z2 * Z2;
        int c = 0;
        if constexpr (MASK & bit(F_POT)) acc[c++] += 0.0;           // not offered by polyprism.py
        if constexpr (MASK & bit(F_GX)) acc[c++] += 0.0;
        if constexpr (MASK & bit(F_GY)) acc[c++] += 0.0;
        if constexpr (has_gz) {                                     // polyprism.py:270-305
            const double p = X1 * Y2 - X2 * Y1;
            const double p_sqr = p * p;
            const double Q1 = (Y2 - Y1) * Y1 + (X2 - X1) * X1;
            const double Q2 = (Y2 - Y1) * Y2 + (X2 - X1) * X2;
            const double A1s = X1 * X1 + Y1 * Y1, A2s = X2 * X2 + Y2 * Y2;
            const double R11 = pp_sqrt(A1s + Z1_sqr), R12 = pp_sqrt(A2s + Z1_sqr);
            const double R21 = pp_sqrt(A1s + Z2_sqr), R22 = pp_sqrt(A2s + Z2_sqr);
            const double A1 = pp_sqrt(A1s), A2 = pp_sqrt(A2s);
            const double B1 = pp_sqrt(Q1 * Q1 + p_sqr), B2 = pp_sqrt(Q2 * Q2 + p_sqr);
            const double E11 = R11 * B1, E12 = R12 * B2, E21 = R21 * B1, E22 = R22 * B2;
            double k = (Z2 - Z1) * (pp_atan2(Q2, p) - pp_atan2(Q1, p));
            k = k + Z2 * (pp_atan2(Z2 * Q1, R21 * p) - pp_atan2(Z2 * Q2, R22 * p));
            k = k + Z1 * (pp_atan2(Z1 * Q2, R12 * p) - pp_atan2(Z1 * Q1, R11 * p));
            const double C1 = Q1 * A1, C2 = Q2 * A2;
            k = k + pp_div(0.5 * p * A1, B1 + d) *
                        pp_log(pp_div((E11 - C1) * (E21 + C1), (E11 + C1) * (E21 - C1) + d) + d);
            k = k + 0.5 * p * pp_div(A2, B2 + d) *
                        pp_log(pp_div((E22 - C2) * (E12 + C2), (E22 + C2) * (E12 - C2) + d) + d);
            acc[c] = acc[c] + k * p0;
            ++c;
        }
        if constexpr (need_a || need_l || need_zz) {
            const double v1 = X1 * X1 + Y1 * Y1, v2 = X2 * X2 + Y2 * Y2;
            const double R11 = pp_sqrt(v1 + Z1_sqr), R12 = pp_sqrt(v1 + Z2_sqr);
            const double R21 = pp_sqrt(v2 + Z1_sqr), R22 = pp_sqrt(v2 + Z2_sqr);
            double exx = 0, exy = 0, exz = 0, eyy = 0, eyz = 0, ezz = 0;
            if constexpr (need_zz) {                                // polyprism.py:1054-1075
                const double dx = X2 - X1, dy = Y2 - Y1;
                const double dist = pp_sqrt(dx * dx + dy * dy) + d;
                const double p = pp_div(X1 * Y2 - X2 * Y1, dist);
                const double d1 = pp_div(dx * X1 + dy * Y1, dist), d2 = pp_div(dx * X2 + dy * Y2, dist);
                ezz = pp_atan2(Z2 * d2, p * R22) - pp_atan2(Z1 * d2, p * R21) - pp_atan2(Z2 * d1, p * R12) +
                      pp_atan2(Z1 * d1, p * R11);
            }
            if constexpr (need_a || need_l) {
                const double dx = X2 - X1 + d, dy = Y2 - Y1 + d;
                const double dist = pp_sqrt(dx * dx + dy * dy);
                const double d1 = pp_div(dx * X1 + dy * Y1, dist) + d, d2 = pp_div(dx * X2 + dy * Y2, dist) + d;
                const double n = pp_div(dx, dy), m = pp_div(dy, dx);
                if constexpr (need_l) {
                    const double log_r22 = pp_log(pp_div(R22 - d2, R22 + d2) + d);
                    const double log_r21 = pp_log(pp_div(R21 - d2, R21 + d2) + d);
                    const double log_r12 = pp_log(pp_div(R12 - d1, R12 + d1) + d);
                    const double log_r11 = pp_log(pp_div(R11 - d1, R11 + d1) + d);
                    if constexpr (need_xz) {                        // polyprism.py:791-824
                        const double n_sqr_p1 = n * n + 1;
                        const double g = X1 - Y1 * n, ng = n * g;
                        const double ld1 = pp_div(0.5, d1) * (log_r12 - log_r11);
                        const double ld2 = pp_div(0.5, d2) * (log_r22 - log_r21);
                        double t = (Y2 * n_sqr_p1 + ng) * ld2;
                        t = t - (Y1 * n_sqr_p1 + ng) * ld1;
                        exz = t * pp_div(-1.0, n_sqr_p1);
                    }
                    if constexpr (need_yz) {                        // polyprism.py:967-997
                        const double m_sqr_p1 = m * m + 1;
                        const double cc = Y1 - X1 * m, cm = cc * m;
                        double t = (X2 * m_sqr_p1 + cm) * pp_div(0.5, d2) * (log_r22 - log_r21);
                        t = t - (X1 * m_sqr_p1 + cm) * pp_div(0.5, d1) * (log_r12 - log_r11);
                        eyz = t * pp_div(1.0, m_sqr_p1);
                    }
                }
                if constexpr (need_a) {
                    const double p = pp_div(X1 * Y2 - X2 * Y1, dist) + d;
                    const double ad2 = pp_atan2(Z2 * d2, p * R22) - pp_atan2(Z1 * d2, p * R21);
                    const double ad1 = pp_atan2(Z2 * d1, p * R12) - pp_atan2(Z1 * d1, p * R11);
                    if constexpr (need_xx) {                        // polyprism.py:617-647
                        const double g = X1 - Y1 * n;
                        double t = pp_div(g * Y2 * ad2, p * d2) + pp_div(n * p * ad2, d2);
                        t = t - (pp_div(g * Y1 * ad1, p * d1) + pp_div(n * p * ad1, d1));
                        t = t + n * pp_log(pp_div((Z2 + R12) * (Z1 + R21), (Z1 + R11) * (Z2 + R22) + d) + d);
                        exx = t * pp_div(-1.0, 1 + n * n);
                    }
                    if constexpr (need_xy) {                        // polyprism.py:703-734
                        const double g = X1 - Y1 * n, g_sqr = g * g;
                        double t = pp_div((g_sqr + g * n * Y2) * ad2, p * d2) - pp_div(p * ad2, d2);
                        t = t - (pp_div((g_sqr + g * n * Y1) * ad1, p * d1) - pp_div(p * ad1, d1));
                        t = t + pp_log(pp_div((Z2 + R22) * (Z1 + R11), (Z1 + R21) * (Z2 + R12) + d) + d);
                        exy = t * pp_div(1.0, 1 + n * n);
                    }
                    if constexpr (need_yy) {                        // polyprism.py:880-910
                        const double cc = Y1 - X1 * m;
                        double t = pp_div(cc * X2 * ad2, p * d2) + pp_div(m * p * ad2, d2);
                        t = t - (pp_div(cc * X1 * ad1, p * d1) + pp_div(m * p * ad1, d1));
                        t = t + m * pp_log(pp_div((Z2 + R12) * (Z1 + R21), (Z2 + R22) * (Z1 + R11)) + d);
                        eyy = t * pp_div(1.0, 1 + m * m);
                    }
                }
            }
            if constexpr (MASK & bit(F_GXX)) { acc[c] = acc[c] + exx * p0; ++c; }
            if constexpr (MASK & bit(F_GXY)) { acc[c] = acc[c] + exy * p0; ++c; }
            if constexpr (MASK & bit(F_GXZ)) { acc[c] = acc[c] + exz * p0; ++c; }
            if constexpr (MASK & bit(F_GYY)) { acc[c] = acc[c] + eyy * p0; ++c; }
            if constexpr (MASK & bit(F_GYZ)) { acc[c] = acc[c] + eyz * p0; ++c; }
            if constexpr (MASK & bit(F_GZZ)) { acc[c] = acc[c] + ezz * p0; ++c; }
            if constexpr (mag) {                                    // polyprism.py:72-81, :118-124
                const double bx = exx * p0 + exy * p1 + exz * p2;
                const double by = exy * p0 + eyy * p1 + eyz * p2;
                const double bz = exz * p0 + eyz * p1 + ezz * p2;
                if constexpr (has_tf) { acc[c] = acc[c] + (f[0] * bx + f[1] * by + f[2] * bz); ++c; }
                if constexpr (has_bx) { acc[c] = acc[c] + bx; ++c; }
                if constexpr (has_by) { acc[c] = acc[c] + by; ++c; }
                if constexpr (has_bz) { acc[c] = acc[c] + bz; ++c; }
                (void)bx; (void)by; (void)bz;
            }
            (void)exx; (void)exy; (void)exz; (void)eyy; (void)eyz; (void)ezz;
        }
    }
};

}  // namespace fb200
