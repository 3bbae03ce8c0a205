// prism_exact.cuh -- reference-order evaluation of one prism x point pair.
//
// This is the parity anchor of the CUDA path: the 8 corners are visited k
// (outer), j, i (inner) exactly like fatiando/gravmag/_prism.pyx:88-104 (and
// every other accumulator there), each corner's term goes straight into the
// running accumulator with the reference's association
//     res = res + ((sign * kernel) * density)            (_prism.pyx:275)
// and every product / sum rounds separately (the translation unit that
// includes this file is compiled with -fmad=false, matching the reference's
// x86-64 -O2 build which has no FMA contraction).  The only arithmetic that
// differs from the reference is inside libdevice log/atan2 (<= 1-2 ulp) versus
// glibc's; sqrt is IEEE-correct on both sides, so all log/atan2 ARGUMENTS are
// bit-identical.  What the fused kernel gains over the reference is that all
// requested fields share one set of r / log / atan2 per corner.
#pragma once
#include <cmath>
#include "prism_common.cuh"

namespace fb200 {

// fatiando/gravmag/_prism.pyx:16-26
FB_HD double safe_atan2_ref(double y, double x)
{
    // the reference's pi literal; rounds to the double nearest pi
    const double PI = 3.1415926535897931159979634685441851615906;
    if (y == 0.0) return 0.0;
    double a = atan2(y, x);
    if (x < 0.0) a = (y > 0.0) ? a - PI : a + PI;
    return a;
}

// fatiando/gravmag/_prism.pyx:28-34
FB_HD double safe_log_ref(double x)
{
    return (x == 0.0) ? 0.0 : log(x);
}

template <uint32_t MASK>
FB_HD void pair_exact(double (&acc)[popc(MASK)],
                                           double xp, double yp, double zp,
                                           double x1, double x2, double y1,
                                           double y2, double z1, double z2,
                                           double p0, double p1, double p2,
                                           const double (&f)[3])
{
    // x = [x2, x1] ... (_prism.pyx:82-84)
    const double X[2] = {x2 - xp, x1 - xp};
    const double Y[2] = {y2 - yp, y1 - yp};
    const double Z[2] = {z2 - zp, z1 - zp};
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        const double dz = Z[k];
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const double dy = Y[j];
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const double dx = X[i];
                const bool neg = (i + j + k) & 1;   // (-1.)**(i+j+k)
                const double r = sqrt(dx * dx + dy * dy + dz * dz);
                double Lx = 0, Ly = 0, Lz = 0, Axx = 0, Ayy = 0, Azz = 0;
                if constexpr (MASK & NEED_LX) Lx = safe_log_ref(dx + r);
                if constexpr (MASK & NEED_LY) Ly = safe_log_ref(dy + r);
                if constexpr (MASK & NEED_LZ) Lz = safe_log_ref(dz + r);
                if constexpr (MASK & NEED_AXX) Axx = safe_atan2_ref(dz * dy, dx * r);
                if constexpr (MASK & NEED_AYY) Ayy = safe_atan2_ref(dz * dx, dy * r);
                if constexpr (MASK & NEED_AZZ) Azz = safe_atan2_ref(dx * dy, dz * r);
                int c = 0;
                auto add = [&](double kernel, double prop) {
                    const double sk = neg ? -kernel : kernel;
                    acc[c] = acc[c] + sk * prop;
                    ++c;
                };
                auto addm = [&](double kernel) {
                    acc[c] = acc[c] + (neg ? -kernel : kernel);
                    ++c;
                };
                if constexpr (MASK & bit(F_POT))   // _prism.pyx:36-39
                    add(dx * dy * Lz + dy * dz * Lx + dx * dz * Ly
                            - 0.5 * (dx * dx) * Axx - 0.5 * (dy * dy) * Ayy
                            - 0.5 * (dz * dz) * Azz, p0);
                if constexpr (MASK & bit(F_GX))    // :43-44
                    add(-(dy * Lz + dz * Ly - dx * Axx), p0);
                if constexpr (MASK & bit(F_GY))    // :46-47
                    add(-(dz * Lx + dx * Lz - dy * Ayy), p0);
                if constexpr (MASK & bit(F_GZ))    // :49-50
                    add(-(dx * Ly + dy * Lx - dz * Azz), p0);
                if constexpr (MASK & bit(F_GXX))   // :52-53
                    add(-Axx, p0);
                if constexpr (MASK & bit(F_GXY)) { // :55-56 with the shift of :327-330
                    double L = Lz;
                    if (dx == 0.0 && dy == 0.0 && dz < 0.0) {
                        const double t1 = 0.00001 * (x2 - x1);
                        const double t2 = 0.00001 * (y2 - y1);
                        L = safe_log_ref(dz + sqrt(t1 * t1 + t2 * t2 + dz * dz));
                    }
                    add(L, p0);
                }
                if constexpr (MASK & bit(F_GXZ)) { // :58-59 with the shift of :359-362
                    double L = Ly;
                    if (dx == 0.0 && dz == 0.0 && dy < 0.0) {
                        const double t1 = 0.00001 * (x2 - x1);
                        const double t2 = 0.00001 * (z2 - z1);
                        L = safe_log_ref(dy + sqrt(t1 * t1 + t2 * t2 + dy * dy));
                    }
                    add(L, p0);
                }
                if constexpr (MASK & bit(F_GYY))   // :61-62
                    add(-Ayy, p0);
                if constexpr (MASK & bit(F_GYZ)) { // :64-65 with the shift of :418-421
                    double L = Lx;
                    if (dy == 0.0 && dz == 0.0 && dx < 0.0) {
                        const double t1 = 0.00001 * (y2 - y1);
                        const double t2 = 0.00001 * (z2 - z1);
                        L = safe_log_ref(dx + sqrt(t1 * t1 + t2 * t2 + dx * dx));
                    }
                    add(L, p0);
                }
                if constexpr (MASK & bit(F_GZZ))   // :67-68
                    add(-Azz, p0);
                // magnetic: v1..v6 = kernelxx, xy, xz, yy, yz, zz (unshifted), :94-99
                if constexpr (MASK & MAGNETIC_MASK) {
                    const double v1 = -Axx, v2 = Lz, v3 = Ly, v4 = -Ayy, v5 = Lx, v6 = -Azz;
                    if constexpr (MASK & bit(F_TF)) {  // :100-104
                        const double bx = (v1 * p0 + v2 * p1 + v3 * p2);
                        const double by = (v2 * p0 + v4 * p1 + v5 * p2);
                        const double bz = (v3 * p0 + v5 * p1 + v6 * p2);
                        addm(f[0] * bx + f[1] * by + f[2] * bz);
                    }
                    if constexpr (MASK & bit(F_BX)) addm(v1 * p0 + v2 * p1 + v3 * p2); // :133
                    if constexpr (MASK & bit(F_BY)) addm(v2 * p0 + v4 * p1 + v5 * p2); // :163
                    if constexpr (MASK & bit(F_BZ)) addm(v3 * p0 + v5 * p1 + v6 * p2); // :193
                }
            }
        }
    }
}

struct EvalExact {
    static constexpr bool NEEDS_FLAG = false;
    static constexpr bool USES_TABLES = false;
    template <uint32_t MASK>
    static FB_HD_M void pair(double (&acc)[popc(MASK)], double xp, double yp, double zp,
                           double x1, double x2, double y1, double y2, double z1, double z2,
                           double p0, double p1, double p2, const double (&f)[3])
    {
        pair_exact<MASK>(acc, xp, yp, zp, x1, x2, y1, y2, z1, z2, p0, p1, p2, f);
    }
};

}  // namespace fb200
