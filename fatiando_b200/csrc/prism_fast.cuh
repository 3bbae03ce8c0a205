// prism_fast.cuh -- fused fast evaluation of one prism x point pair.
//
// Same inputs, same log/atan2 ARGUMENT bits as the reference kernel
// (fatiando/gravmag/_prism.pyx:36-68,88-104), but the alternating 8-corner
// sums are collapsed before any transcendental is evaluated (see
// prism_fastmath.cuh).  Per pair this needs, e.g., for gz 4 log_ratio + 2
// arctangents instead of 16 log + 8 atan2, and for the six tensor components
// 3 + 3 instead of 24 + 24.  The special cases of the reference are kept:
//   * safe_log(0) = 0                        (_prism.pyx:28-34)
//   * safe_atan2(0, 0) = 0 and the folding of atan2 to (-pi/2, pi/2] (:16-26)
//   * the 1e-5*size shift of r for gxy/gxz/gyz when the point is aligned with
//     an edge on the far side (:327-330, :359-362, :418-421) -- and NOT for
//     tf/bx/by/bz, which use the unshifted kernels (:94-99).
// They only matter when coordinate differences are exactly zero on two axes
// at once, or a log argument rounds to exactly zero (point on, or within ~1e-8
// relative of, the line through an edge).  The hot path carries none of that: it detects those
// pairs with a handful of ALU-pipe tests and two compares per log group, and
// if ANY lane of the warp has one (warp vote, so the warp never diverges) the
// whole warp re-evaluates the pair with the out-of-line "careful" variant,
// which applies the special cases with selects.  Differences so small that
// their squares underflow (|d| < 2^-480, never seen in metres) fall back to
// the reference-order evaluator for that pair.
#pragma once
#include "prism_common.cuh"
#include "prism_fastmath.cuh"
#include "prism_exact.cuh"

namespace fb200 {

enum LogMode : int { LM_NONE = 0, LM_ALL, LM_BYP, LM_BYQ, LM_FULL };
enum AtanMode : int { AM_NONE = 0, AM_ALL, AM_BYOWN };

constexpr LogMode log_mode(bool full, bool byq, bool byp, bool all)
{
    return full || (byq && byp) ? LM_FULL : byq ? LM_BYQ : byp ? LM_BYP : all ? LM_ALL : LM_NONE;
}
constexpr AtanMode atan_mode(bool byown, bool all)
{
    return byown ? AM_BYOWN : all ? AM_ALL : AM_NONE;
}
constexpr int log_groups(LogMode m) { return m == LM_FULL ? 4 : (m == LM_BYQ || m == LM_BYP) ? 2 : m == LM_ALL ? 1 : 0; }

// corner index c = k*4 + j*2 + i.  A family has an "own" axis o (the coordinate
// that appears inside its log / as its atan denominator) and the two other
// axes p < q (by bit position).
template <int OWN>
FB_HD constexpr int cidx(int o, int p, int q)
{
    return OWN == 2 ? ((o << 2) | (q << 1) | p)
         : OWN == 1 ? ((q << 2) | (o << 1) | p)
                    : ((q << 2) | (p << 1) | o);
}

// sum over the own axis (and more) of s*log(a), by groups.  Two steps so that
// the caller can look at the products before any logarithm is taken:
//   products(a): num[g], den[g] of each group, factors multiplied in a fixed order
//   finish():    v[g] = log(num[g] / den[g])
template <int OWN, LogMode MODE>
struct LogFam {
    static constexpr int NG = log_groups(MODE);
    double num[NG > 0 ? NG : 1], den[NG > 0 ? NG : 1];
    double v[NG > 0 ? NG : 1];

    // CAREFUL: a factor that is exactly zero drops out -- the reference's
    // safe_log(0) = 0 (_prism.pyx:28-34)
    template <bool CAREFUL>
    static FB_HD_M double fac(double a)
    {
        if constexpr (CAREFUL) return (a == 0.0) ? 1.0 : a;
        else return a;
    }

    template <bool CAREFUL>
    FB_HD_M void products(const double (&a)[8])
    {
#define A_(o, p, q) fac<CAREFUL>(a[cidx<OWN>(o, p, q)])
        if constexpr (MODE == LM_FULL) {
#pragma unroll
            for (int q = 0; q < 2; ++q)
#pragma unroll
                for (int p = 0; p < 2; ++p) {
                    num[q * 2 + p] = A_(0, p, q);
                    den[q * 2 + p] = A_(1, p, q);
                }
        } else if constexpr (MODE == LM_BYQ) {
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                num[q] = A_(0, 0, q) * A_(1, 1, q);
                den[q] = A_(1, 0, q) * A_(0, 1, q);
            }
        } else if constexpr (MODE == LM_BYP) {
#pragma unroll
            for (int p = 0; p < 2; ++p) {
                num[p] = A_(0, p, 0) * A_(1, p, 1);
                den[p] = A_(1, p, 0) * A_(0, p, 1);
            }
        } else if constexpr (MODE == LM_ALL) {
            num[0] = (A_(0, 0, 0) * A_(1, 1, 0)) * (A_(1, 0, 1) * A_(0, 1, 1));
            den[0] = (A_(1, 0, 0) * A_(0, 1, 0)) * (A_(0, 0, 1) * A_(1, 1, 1));
        }
#undef A_
    }
    // true when some log argument of this family is exactly zero
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
    // sum_{p,q} (-1)^(p+q) w[q][p] * lambda[q][p]
    FB_HD_M double wsum_full(const double (&w)[2][2]) const
    {
        static_assert(MODE == LM_FULL, "needs LM_FULL");
        return (w[0][0] * v[0] - w[0][1] * v[1]) - (w[1][0] * v[2] - w[1][1] * v[3]);
    }
    // sum_q (-1)^q wq[q] * (lambda[q][0] - lambda[q][1])
    FB_HD_M double wsum_q(const double (&wq)[2]) const
    {
        static_assert(MODE == LM_FULL || MODE == LM_BYQ, "needs q groups");
        if constexpr (MODE == LM_FULL)
            return wq[0] * (v[0] - v[1]) - wq[1] * (v[2] - v[3]);
        else
            return wq[0] * v[0] - wq[1] * v[1];
    }
    // sum_p (-1)^p wp[p] * (lambda[0][p] - lambda[1][p])
    FB_HD_M double wsum_p(const double (&wp)[2]) const
    {
        static_assert(MODE == LM_FULL || MODE == LM_BYP, "needs p groups");
        if constexpr (MODE == LM_FULL)
            return wp[0] * (v[0] - v[2]) - wp[1] * (v[1] - v[3]);
        else
            return wp[0] * v[0] - wp[1] * v[1];
    }
    FB_HD_M double all() const
    {
        static_assert(MODE != LM_NONE, "family not evaluated");
        if constexpr (MODE == LM_FULL) return (v[0] - v[1]) - (v[2] - v[3]);
        else if constexpr (MODE == LM_ALL) return v[0];
        else return v[0] - v[1];
    }
};

// sum of s*safe_atan2(im, x*r) grouped by the own axis.  re[c] = x_o * r_c with
// the sign of x_o (the folding of atan2 to (-pi/2, pi/2] is done by the sign
// bookkeeping of Wind); im[q*2+p] = product of the two other coordinates;
// hown[o] = high word of the own coordinate x_o.
template <int OWN, AtanMode MODE>
struct AtanFam {
    double v[2];

    FB_HD_M void eval(const double (&re)[8], const double (&im)[4], const int (&hown)[2])
    {
        using namespace fm;
        if constexpr (MODE == AM_BYOWN) {
#pragma unroll
            for (int o = 0; o < 2; ++o) {
                Wind w = wind_start<false>(re[cidx<OWN>(o, 0, 0)], im[0]);
                wind_mul<true>(w, re[cidx<OWN>(o, 1, 0)], im[1], hown[o]);
                wind_mul<true>(w, re[cidx<OWN>(o, 0, 1)], im[2], hown[o]);
                wind_mul<false>(w, re[cidx<OWN>(o, 1, 1)], im[3], hown[o]);
                v[o] = wind_angle(w);
            }
        } else if constexpr (MODE == AM_ALL) {
            Wind w = wind_start<false>(re[cidx<OWN>(0, 0, 0)], im[0]);
            wind_mul<true>(w, re[cidx<OWN>(0, 1, 0)], im[1], hown[0]);
            wind_mul<true>(w, re[cidx<OWN>(0, 0, 1)], im[2], hown[0]);
            wind_mul<false>(w, re[cidx<OWN>(0, 1, 1)], im[3], hown[0]);
            wind_mul<true>(w, re[cidx<OWN>(1, 0, 0)], im[0], hown[1]);
            wind_mul<false>(w, re[cidx<OWN>(1, 1, 0)], im[1], hown[1]);
            wind_mul<false>(w, re[cidx<OWN>(1, 0, 1)], im[2], hown[1]);
            wind_mul<true>(w, re[cidx<OWN>(1, 1, 1)], im[3], hown[1]);
            v[0] = wind_angle(w);
        }
    }
    // sum_o (-1)^o wo[o] * theta[o]
    FB_HD_M double wsum_o(const double (&wo)[2]) const
    {
        static_assert(MODE == AM_BYOWN, "needs per-own-axis angles");
        return wo[0] * v[0] - wo[1] * v[1];
    }
    FB_HD_M double all() const
    {
        static_assert(MODE != AM_NONE, "family not evaluated");
        if constexpr (MODE == AM_BYOWN) return v[0] - v[1];
        else return v[0];
    }
};

// |d| < 2^-480 (zero, denormal, or so small that d*d leaves the range of the
// branch-free root): exponent test on the ALU pipe.
FB_HD bool tiny_or_zero(double d) { return (fm::hi32(d) & 0x7ff00000) < 0x21f00000; }

// contribution of one pair to every requested component (before it is added
// to the running sums): what the out-of-line careful variant hands back
template <int NC>
struct Contribution {
    double v[NC];
};

template <uint32_t MASK>
#if defined(__CUDACC__)
__device__ __noinline__
#else
static
#endif
Contribution<popc(MASK)> pair_careful(double xp, double yp, double zp, double x1, double x2,
                                      double y1, double y2, double z1, double z2, double p0,
                                      double p1, double p2, double f0, double f1, double f2);

// One pair.  CAREFUL = false: the hot path; returns false -- having added
// nothing -- when the pair needs the special cases.  CAREFUL = true: applies
// them, always returns true.  out[c] receives the pair's contribution.
template <uint32_t MASK, bool CAREFUL>
FB_HD bool pair_core(double (&out)[popc(MASK)], double xp, double yp, double zp,
                     double x1, double x2, double y1, double y2, double z1,
                     double z2, double p0, double p1, double p2,
                     const double (&f)[3])
{
    using namespace fm;
    constexpr bool MAG = (MASK & MAGNETIC_MASK) != 0;
    constexpr bool has_pot = MASK & bit(F_POT), has_gx = MASK & bit(F_GX),
                   has_gy = MASK & bit(F_GY), has_gz = MASK & bit(F_GZ),
                   has_gxx = MASK & bit(F_GXX), has_gxy = MASK & bit(F_GXY),
                   has_gxz = MASK & bit(F_GXZ), has_gyy = MASK & bit(F_GYY),
                   has_gyz = MASK & bit(F_GYZ), has_gzz = MASK & bit(F_GZZ);
    constexpr bool has_tf = MASK & bit(F_TF), has_bx = MASK & bit(F_BX),
                   has_by = MASK & bit(F_BY), has_bz = MASK & bit(F_BZ);
    // which groupings each family must deliver (see the header comment of
    // LogFam: Lz has p = i, q = j; Ly has p = i, q = k; Lx has p = j, q = k)
    constexpr LogMode MLZ = MAG ? ((has_tf || has_bx || has_by) ? LM_ALL : LM_NONE)
                                : log_mode(has_pot, has_gx, has_gy, has_gxy);
    constexpr LogMode MLY = MAG ? ((has_tf || has_bx || has_bz) ? LM_ALL : LM_NONE)
                                : log_mode(has_pot, has_gx, has_gz, has_gxz);
    constexpr LogMode MLX = MAG ? ((has_tf || has_by || has_bz) ? LM_ALL : LM_NONE)
                                : log_mode(has_pot, has_gy, has_gz, has_gyz);
    constexpr AtanMode MAX_ = MAG ? ((has_tf || has_bx) ? AM_ALL : AM_NONE)
                                  : atan_mode(has_pot || has_gx, has_gxx);
    constexpr AtanMode MAY_ = MAG ? ((has_tf || has_by) ? AM_ALL : AM_NONE)
                                  : atan_mode(has_pot || has_gy, has_gyy);
    constexpr AtanMode MAZ_ = MAG ? ((has_tf || has_bz) ? AM_ALL : AM_NONE)
                                  : atan_mode(has_pot || has_gz, has_gzz);

    // integration limits, index 0 <-> the "2" bound (_prism.pyx:82-84,88-92)
    double X[2] = {sub(x2, xp), sub(x1, xp)};
    double Y[2] = {sub(y2, yp), sub(y1, yp)};
    double Z[2] = {sub(z2, zp), sub(z1, zp)};
    if constexpr (CAREFUL) {
        // the sign bookkeeping below reads sign BITS; the reference's tests are
        // `< 0`, false for -0.0: make exact zeros +0.0
#pragma unroll
        for (int t = 0; t < 2; ++t) { X[t] = add(X[t], 0.0); Y[t] = add(Y[t], 0.0); Z[t] = add(Z[t], 0.0); }
    }
    const double XX[2] = {mul(X[0], X[0]), mul(X[1], X[1])};
    const double YY[2] = {mul(Y[0], Y[0]), mul(Y[1], Y[1])};
    const double ZZ[2] = {mul(Z[0], Z[0]), mul(Z[1], Z[1])};
    // Exact zeros (and differences whose squares would underflow), ALU pipe.  A
    // zero on ONE axis is harmless on the hot path (r > 0, every complex factor
    // is non-zero, signs are those of +0.0 because the model bounds are
    // canonicalised at upload); the special cases need zeros on two axes at once
    // (point on the line through an edge, or on a vertex).
    const bool sx = tiny_or_zero(X[0]) | tiny_or_zero(X[1]);
    const bool sy = tiny_or_zero(Y[0]) | tiny_or_zero(Y[1]);
    const bool sz = tiny_or_zero(Z[0]) | tiny_or_zero(Z[1]);
    const bool special = (sx & sy) | (sx & sz) | (sy & sz);
    bool zx[2] = {false, false}, zy[2] = {false, false}, zz[2] = {false, false};
    if constexpr (CAREFUL) {
        zx[0] = X[0] == 0.0; zx[1] = X[1] == 0.0;
        zy[0] = Y[0] == 0.0; zy[1] = Y[1] == 0.0;
        zz[0] = Z[0] == 0.0; zz[1] = Z[1] == 0.0;
        if ((tiny_or_zero(X[0]) & !zx[0]) | (tiny_or_zero(X[1]) & !zx[1]) |
            (tiny_or_zero(Y[0]) & !zy[0]) | (tiny_or_zero(Y[1]) & !zy[1]) |
            (tiny_or_zero(Z[0]) & !zz[0]) | (tiny_or_zero(Z[1]) & !zz[1])) {
            // squares underflow: evaluate this pair the reference's way
#pragma unroll
            for (int c = 0; c < popc(MASK); ++c) out[c] = 0.0;
            pair_exact<MASK>(out, xp, yp, zp, x1, x2, y1, y2, z1, z2, p0, p1, p2, f);
            return true;
        }
    }
    double r[8];
#pragma unroll
    for (int j = 0; j < 2; ++j)
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const double sxy = add(XX[i], YY[j]);        // (dx**2 + dy**2) + dz**2
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                const double rr = root(add(sxy, ZZ[k]));
                // on a vertex r2 == 0 and the branch-free root gives NaN
                if constexpr (CAREFUL) r[k * 4 + j * 2 + i] = (zx[i] & zy[j] & zz[k]) ? 0.0 : rr;
                else r[k * 4 + j * 2 + i] = rr;
            }
        }

    // ---- arguments, with the reference's roundings -------------------------
    double ax[8], ay[8], az[8];          // x + r, y + r, z + r
    double rex[8], rey[8], rez[8];       // x*r, y*r, z*r
    double imx[4], imy[4], imz[4];       // z*y [k][j], z*x [k][i], x*y [j][i]  (q*2+p)
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        const int i = c & 1, j = (c >> 1) & 1, k = c >> 2;
        if constexpr (MLX != LM_NONE) ax[c] = add(X[i], r[c]);
        if constexpr (MLY != LM_NONE) ay[c] = add(Y[j], r[c]);
        if constexpr (MLZ != LM_NONE) az[c] = add(Z[k], r[c]);
        if constexpr (MAX_ != AM_NONE) rex[c] = mul(X[i], r[c]);
        if constexpr (MAY_ != AM_NONE) rey[c] = mul(Y[j], r[c]);
        if constexpr (MAZ_ != AM_NONE) rez[c] = mul(Z[k], r[c]);
        if constexpr (CAREFUL) {
            // safe_atan2(0, 0) = 0 (_prism.pyx:16-26): unit factor
            if constexpr (MAX_ != AM_NONE) if (zx[i] & (zy[j] | zz[k])) rex[c] = 1.0;
            if constexpr (MAY_ != AM_NONE) if (zy[j] & (zx[i] | zz[k])) rey[c] = 1.0;
            if constexpr (MAZ_ != AM_NONE) if (zz[k] & (zx[i] | zy[j])) rez[c] = 1.0;
        }
    }
#pragma unroll
    for (int q = 0; q < 2; ++q)
#pragma unroll
        for (int p = 0; p < 2; ++p) {
            if constexpr (MAX_ != AM_NONE) imx[q * 2 + p] = mul(Z[q], Y[p]);
            if constexpr (MAY_ != AM_NONE) imy[q * 2 + p] = mul(Z[q], X[p]);
            if constexpr (MAZ_ != AM_NONE) imz[q * 2 + p] = mul(X[p], Y[q]);
        }

    LogFam<0, MLX> Lx;
    LogFam<1, MLY> Ly;
    LogFam<2, MLZ> Lz;
    if constexpr (MLX != LM_NONE) Lx.template products<CAREFUL>(ax);
    if constexpr (MLY != LM_NONE) Ly.template products<CAREFUL>(ay);
    if constexpr (MLZ != LM_NONE) Lz.template products<CAREFUL>(az);

    double ex_xy = 0.0, ex_xz = 0.0, ex_yz = 0.0;   // shifted-r terms of gxy, gxz, gyz
    if constexpr (!CAREFUL) {
        bool bad = special;
        if constexpr (MLX != LM_NONE) bad = bad | Lx.any_zero();
        if constexpr (MLY != LM_NONE) bad = bad | Ly.any_zero();
        if constexpr (MLZ != LM_NONE) bad = bad | Lz.any_zero();
#if defined(__CUDA_ARCH__)
        // warp vote: either every lane continues on the hot path or every lane
        // re-evaluates carefully -- the warp never executes both
        bad = __any_sync(0xffffffffu, bad);
#endif
        if (bad) return false;
    } else {
        // aligned with an edge on the far side: shifted r, gravity tensor only
        // (_prism.pyx:327-330, :359-362, :418-421)
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            const int i = c & 1, j = (c >> 1) & 1, k = c >> 2;
            const bool neg = (i + j + k) & 1;
            if constexpr (has_gxy) if (zx[i] & zy[j] & (Z[k] < 0.0)) {
                const double t1 = mul(0.00001, sub(x2, x1)), t2 = mul(0.00001, sub(y2, y1));
                const double L = safe_log_ref(add(Z[k], sqrt(add(add(mul(t1, t1), mul(t2, t2)), ZZ[k]))));
                ex_xy += neg ? -L : L;
            }
            if constexpr (has_gxz) if (zx[i] & zz[k] & (Y[j] < 0.0)) {
                const double t1 = mul(0.00001, sub(x2, x1)), t2 = mul(0.00001, sub(z2, z1));
                const double L = safe_log_ref(add(Y[j], sqrt(add(add(mul(t1, t1), mul(t2, t2)), YY[j]))));
                ex_xz += neg ? -L : L;
            }
            if constexpr (has_gyz) if (zy[j] & zz[k] & (X[i] < 0.0)) {
                const double t1 = mul(0.00001, sub(y2, y1)), t2 = mul(0.00001, sub(z2, z1));
                const double L = safe_log_ref(add(X[i], sqrt(add(add(mul(t1, t1), mul(t2, t2)), XX[i]))));
                ex_yz += neg ? -L : L;
            }
        }
    }

    // ---- collapsed transcendental evaluation: one straight-line block ---------
    AtanFam<0, MAX_> Ax;
    AtanFam<1, MAY_> Ay;
    AtanFam<2, MAZ_> Az;
    const int hx[2] = {hi32(X[0]), hi32(X[1])};
    const int hy[2] = {hi32(Y[0]), hi32(Y[1])};
    const int hz[2] = {hi32(Z[0]), hi32(Z[1])};
    if constexpr (MLX != LM_NONE) Lx.finish();
    if constexpr (MLY != LM_NONE) Ly.finish();
    if constexpr (MLZ != LM_NONE) Lz.finish();
    if constexpr (MAX_ != AM_NONE) Ax.eval(rex, imx, hx);
    if constexpr (MAY_ != AM_NONE) Ay.eval(rey, imy, hy);
    if constexpr (MAZ_ != AM_NONE) Az.eval(rez, imz, hz);

    int c = 0;
    if constexpr (!MAG) {
        if constexpr (has_pot) {      // _prism.pyx:36-39
            const double wxy[2][2] = {{X[0] * Y[0], X[1] * Y[0]}, {X[0] * Y[1], X[1] * Y[1]}};  // [j][i]
            const double wxz[2][2] = {{X[0] * Z[0], X[1] * Z[0]}, {X[0] * Z[1], X[1] * Z[1]}};  // [k][i]
            const double wyz[2][2] = {{Y[0] * Z[0], Y[1] * Z[0]}, {Y[0] * Z[1], Y[1] * Z[1]}};  // [k][j]
            const double hx[2] = {0.5 * XX[0], 0.5 * XX[1]};
            const double hy[2] = {0.5 * YY[0], 0.5 * YY[1]};
            const double hz[2] = {0.5 * ZZ[0], 0.5 * ZZ[1]};
            out[c++] = (((Lz.wsum_full(wxy) + Lx.wsum_full(wyz)) + Ly.wsum_full(wxz))
                        - ((Ax.wsum_o(hx) + Ay.wsum_o(hy)) + Az.wsum_o(hz))) * p0;
        }
        if constexpr (has_gx)         // :43-44
            out[c++] = (Ax.wsum_o(X) - (Lz.wsum_q(Y) + Ly.wsum_q(Z))) * p0;
        if constexpr (has_gy)         // :46-47
            out[c++] = (Ay.wsum_o(Y) - (Lx.wsum_q(Z) + Lz.wsum_p(X))) * p0;
        if constexpr (has_gz)         // :49-50
            out[c++] = (Az.wsum_o(Z) - (Ly.wsum_p(X) + Lx.wsum_p(Y))) * p0;
        if constexpr (has_gxx) out[c++] = -Ax.all() * p0;              // :52-53
        if constexpr (has_gxy) out[c++] = (Lz.all() + ex_xy) * p0;     // :55-56
        if constexpr (has_gxz) out[c++] = (Ly.all() + ex_xz) * p0;     // :58-59
        if constexpr (has_gyy) out[c++] = -Ay.all() * p0;              // :61-62
        if constexpr (has_gyz) out[c++] = (Lx.all() + ex_yz) * p0;     // :64-65
        if constexpr (has_gzz) out[c++] = -Az.all() * p0;              // :67-68
    } else {
        // second derivatives of the 1/r volume integral, summed over corners
        // (v1..v6 of _prism.pyx:94-99); unused ones are never evaluated
        double vxx = 0, vxy = 0, vxz = 0, vyy = 0, vyz = 0, vzz = 0;
        if constexpr (MAX_ != AM_NONE) vxx = -Ax.all();
        if constexpr (MLZ != LM_NONE) vxy = Lz.all();
        if constexpr (MLY != LM_NONE) vxz = Ly.all();
        if constexpr (MAY_ != AM_NONE) vyy = -Ay.all();
        if constexpr (MLX != LM_NONE) vyz = Lx.all();
        if constexpr (MAZ_ != AM_NONE) vzz = -Az.all();
        const double bx = fma(vxx, p0, fma(vxy, p1, vxz * p2));   // :100
        const double by = fma(vxy, p0, fma(vyy, p1, vyz * p2));   // :101
        const double bz = fma(vxz, p0, fma(vyz, p1, vzz * p2));   // :102
        if constexpr (has_tf) out[c++] = fma(f[0], bx, fma(f[1], by, f[2] * bz));   // :103
        if constexpr (has_bx) out[c++] = bx;
        if constexpr (has_by) out[c++] = by;
        if constexpr (has_bz) out[c++] = bz;
    }
    (void)ex_xy; (void)ex_xz; (void)ex_yz; (void)special;
    return true;
}

template <uint32_t MASK>
#if defined(__CUDACC__)
__device__ __noinline__
#else
static
#endif
Contribution<popc(MASK)> pair_careful(double xp, double yp, double zp, double x1, double x2,
                                      double y1, double y2, double z1, double z2, double p0,
                                      double p1, double p2, double f0, double f1, double f2)
{
    Contribution<popc(MASK)> res;
    const double f[3] = {f0, f1, f2};
    pair_core<MASK, true>(res.v, xp, yp, zp, x1, x2, y1, y2, z1, z2, p0, p1, p2, f);
    return res;
}

template <uint32_t MASK>
FB_HD void pair_fast(double (&acc)[popc(MASK)], double xp, double yp, double zp,
                     double x1, double x2, double y1, double y2, double z1,
                     double z2, double p0, double p1, double p2,
                     const double (&f)[3])
{
    constexpr int NC = popc(MASK);
    double out[NC];
    if (!pair_core<MASK, false>(out, xp, yp, zp, x1, x2, y1, y2, z1, z2, p0, p1, p2, f)) {
        const Contribution<NC> res =
            pair_careful<MASK>(xp, yp, zp, x1, x2, y1, y2, z1, z2, p0, p1, p2, f[0], f[1], f[2]);
#pragma unroll
        for (int c = 0; c < NC; ++c) out[c] = res.v[c];
    }
#pragma unroll
    for (int c = 0; c < NC; ++c) acc[c] += out[c];
}

struct EvalFast {
    static constexpr bool NEEDS_FLAG = false;
    static constexpr bool USES_TABLES = true;    // log / atan tables of prism_fastmath.cuh
    template <uint32_t MASK>
    static FB_HD_M void pair(double (&acc)[popc(MASK)], double xp, double yp, double zp,
                           double x1, double x2, double y1, double y2, double z1, double z2,
                           double p0, double p1, double p2, const double (&f)[3])
    {
        pair_fast<MASK>(acc, xp, yp, zp, x1, x2, y1, y2, z1, z2, p0, p1, p2, f);
    }
};

}  // namespace fb200
