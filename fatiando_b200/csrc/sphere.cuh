// sphere.cuh -- the sphere (point-mass / dipole) forward model as an evaluator policy.
//
// Sibling of the prism path: fatiando/gravmag/sphere.py sums,
// per sphere and in numpy, gz (:316-373), the gradient tensor gxx..gzz (:376-746 with
// the second derivatives _v_xx.._v_zz of 1/r, :21-42) and tf / bx / by / bz (:45-313).
// Same outer structure as the prism kernels -- thread = observation point, spheres
// staged through the TMA ring of forward_kernel -- so it plugs in as a policy.
// Staged rows: xc, volume, yc, (unused), zc, (unused), p0 [, p1, p2]
//   volume = 4*pi*radius**3/3 evaluated on the host like sphere.py:348 does;
//   p = density, or the magnetization vector.
// Arithmetic follows the reference's order (x**2 + y**2 + z**2 left to right,
// r_5 = r*r*r*r*r, 3*dotprod*x - r_sqr*mx, ...) with one reciprocal of r_5 (or r^3)
// shared by all requested components instead of one division each.
#pragma once
#include "prism_common.cuh"
#include "prism_fastmath.cuh"

namespace fb200 {

// fields sphere.py offers: gz, the six tensor components, tf, bx, by, bz
constexpr uint32_t SPHERE_FIELDS = bit(F_GZ) | bit(F_GXX) | bit(F_GXY) | bit(F_GXZ) | bit(F_GYY) |
                                   bit(F_GYZ) | bit(F_GZZ) | MAGNETIC_MASK;

struct EvalSphere {
    static constexpr bool NEEDS_FLAG = false;
    static constexpr bool USES_TABLES = false;
    template <uint32_t MASK>
    static FB_HD_M void pair(double (&acc)[popc(MASK)], double xp, double yp, double zp,
                           double xc, double volume, double yc, double, double zc, double,
                           double p0, double p1, double p2, const double (&f)[3])
    {
        using namespace fm;
        const double x = sub(xc, xp), y = sub(yc, yp), z = sub(zc, zp);
        const double r_sqr = add(add(mul(x, x), mul(y, y)), mul(z, z));
        const double r = sqrt(r_sqr);
        int c = 0;
        if constexpr (MASK & bit(F_POT)) acc[c++] += 0.0;     // not offered by sphere.py
        if constexpr (MASK & bit(F_GX)) acc[c++] += 0.0;
        if constexpr (MASK & bit(F_GY)) acc[c++] += 0.0;
        if constexpr (MASK & bit(F_GZ)) {                     // sphere.py:366-371
            const double r_cb = mul(mul(r, r), r);
            acc[c] = add(acc[c], mul(mul(p0, volume), z) / r_cb);
            ++c;
        }
        if constexpr (MASK & (0x3F0u | MAGNETIC_MASK)) {
            const double r_5 = mul(mul(mul(mul(r, r), r), r), r);
            const double inv = 1.0 / r_5;
            if constexpr (MASK & 0x3F0u) {
                const double dv = mul(p0, volume);            // density*volume, sphere.py:432
                if constexpr (MASK & bit(F_GXX)) { acc[c] = fma(dv, mul(sub(mul(3.0, mul(x, x)), r_sqr), inv), acc[c]); ++c; }
                if constexpr (MASK & bit(F_GXY)) { acc[c] = fma(dv, mul(mul(mul(3.0, x), y), inv), acc[c]); ++c; }
                if constexpr (MASK & bit(F_GXZ)) { acc[c] = fma(dv, mul(mul(mul(3.0, x), z), inv), acc[c]); ++c; }
                if constexpr (MASK & bit(F_GYY)) { acc[c] = fma(dv, mul(sub(mul(3.0, mul(y, y)), r_sqr), inv), acc[c]); ++c; }
                if constexpr (MASK & bit(F_GYZ)) { acc[c] = fma(dv, mul(mul(mul(3.0, y), z), inv), acc[c]); ++c; }
                if constexpr (MASK & bit(F_GZZ)) { acc[c] = fma(dv, mul(sub(mul(3.0, mul(z, z)), r_sqr), inv), acc[c]); ++c; }
            }
            if constexpr (MASK & MAGNETIC_MASK) {
                // sphere.py:118-124: dotprod = mx*x + my*y + mz*z; b = (3*dotprod*x - r_sqr*mx)/r_5
                const double dot = add(add(mul(p0, x), mul(p1, y)), mul(p2, z));
                const double t3 = mul(3.0, dot);
                const double bx = mul(sub(mul(t3, x), mul(r_sqr, p0)), inv);
                const double by = mul(sub(mul(t3, y), mul(r_sqr, p1)), inv);
                const double bz = mul(sub(mul(t3, z), mul(r_sqr, p2)), inv);
                if constexpr (MASK & bit(F_TF)) {
                    acc[c] = fma(volume, add(add(mul(f[0], bx), mul(f[1], by)), mul(f[2], bz)), acc[c]);
                    ++c;
                }
                if constexpr (MASK & bit(F_BX)) { acc[c] = fma(volume, bx, acc[c]); ++c; }
                if constexpr (MASK & bit(F_BY)) { acc[c] = fma(volume, by, acc[c]); ++c; }
                if constexpr (MASK & bit(F_BZ)) { acc[c] = fma(volume, bz, acc[c]); ++c; }
            }
        }
    }
};

}  // namespace fb200
