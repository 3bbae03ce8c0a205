// prism_common.cuh -- shared definitions of the sm_100a prism kernels.
//
// Field ids and bit masks mirror include/fatiando_b200.h; the order is the
// order of the 14 accumulators of the reference kernel
// (fatiando/gravmag/_prism.pyx:72-479).
#pragma once
#include <cstdint>

#if defined(__CUDACC__)
#define FB_HD __host__ __device__ __forceinline__
#define FB_HD_M __host__ __device__ __forceinline__
#else
#define FB_HD static inline __attribute__((always_inline))
#define FB_HD_M inline __attribute__((always_inline))
#endif

namespace fb200 {

enum Field : int {
    F_POT = 0, F_GX, F_GY, F_GZ, F_GXX, F_GXY, F_GXZ, F_GYY, F_GYZ, F_GZZ,
    F_TF, F_BX, F_BY, F_BZ, F_COUNT
};

constexpr uint32_t bit(int f) { return 1u << f; }
constexpr uint32_t GRAVITY_MASK = 0x03FFu;
constexpr uint32_t MAGNETIC_MASK = 0x3C00u;

constexpr int popc(uint32_t m)
{
    int c = 0;
    for (; m; m &= m - 1) ++c;
    return c;
}

// Which of the six per-corner primitives a set of fields needs
// (kernel table of _prism.pyx:36-68):
//   Lx = safe_log(x + r)  -> potential, gy, gz, gyz | tf, by, bz
//   Ly = safe_log(y + r)  -> potential, gx, gz, gxz | tf, bx, bz
//   Lz = safe_log(z + r)  -> potential, gx, gy, gxy | tf, bx, by
//   Axx = safe_atan2(z*y, x*r) -> potential, gx, gxx | tf, bx
//   Ayy = safe_atan2(z*x, y*r) -> potential, gy, gyy | tf, by
//   Azz = safe_atan2(x*y, z*r) -> potential, gz, gzz | tf, bz
constexpr uint32_t NEED_LX = bit(F_POT) | bit(F_GY) | bit(F_GZ) | bit(F_GYZ) | bit(F_TF) | bit(F_BY) | bit(F_BZ);
constexpr uint32_t NEED_LY = bit(F_POT) | bit(F_GX) | bit(F_GZ) | bit(F_GXZ) | bit(F_TF) | bit(F_BX) | bit(F_BZ);
constexpr uint32_t NEED_LZ = bit(F_POT) | bit(F_GX) | bit(F_GY) | bit(F_GXY) | bit(F_TF) | bit(F_BX) | bit(F_BY);
constexpr uint32_t NEED_AXX = bit(F_POT) | bit(F_GX) | bit(F_GXX) | bit(F_TF) | bit(F_BX);
constexpr uint32_t NEED_AYY = bit(F_POT) | bit(F_GY) | bit(F_GYY) | bit(F_TF) | bit(F_BY);
constexpr uint32_t NEED_AZZ = bit(F_POT) | bit(F_GZ) | bit(F_GZZ) | bit(F_TF) | bit(F_BZ);

// Prisms staged per shared-memory tile and threads per block of the forward
// kernels.  9 SoA rows of TILE doubles = 18 KB per stage.
constexpr int TILE = 256;
constexpr int FWD_THREADS = 128;
constexpr int PRISM_ROWS = 9;  // x1 x2 y1 y2 z1 z2 p0 p1 p2

// Device-side view of a flattened model (SoA, allocation padded to TILE).
struct ModelView {
    const double *b[6];   // x1, x2, y1, y2, z1, z2
    const double *p[3];   // gravity: p[0] = density; magnetic: mx, my, mz
    int64_t m;
};

struct ForwardArgs {
    ModelView model;
    const double *xp, *yp, *zp;
    int64_t n;
    // property override (dens= / pmag= of prism.py): replaces p[] for every prism
    int use_override;
    double override_p[3];
    double fvec[3];          // direction cosines of the inducing field (tf)
    // prism split: block (bx, by) handles prisms [by*chunk, min(m,(by+1)*chunk))
    int64_t chunk;
    int nsplit;
    // nsplit == 1: out[out_row[c]*ld + l] = acc[c] * scale[c]
    // nsplit  > 1: out is scratch, out[(split*ncomp + c)*ld + l] = acc[c] (unscaled)
    double *out;
    int64_t ld;
    const double *init;      // optional starting value of the running sum (ncomp == 1)
    double scale[F_COUNT];   // per component of this launch, ascending field order
    int out_row[F_COUNT];    // destination row of each component
    int *flag;               // flag[l] raised for a point where an evaluator meets a case it cannot settle (EvalNode)
};

}  // namespace fb200
