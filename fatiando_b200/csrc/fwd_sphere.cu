// fwd_sphere.cu -- sphere forward and Jacobian kernels (sphere.cuh through the prism harness).
#include "prism_masks.h"
#include "sphere.cuh"
#include "prism_forward.cuh"
#include "prism_jacobian.cuh"
#include "prism_launch.h"

namespace fb200 {

cudaError_t launch_forward_sphere(uint32_t mask, const ForwardArgs &a, cudaStream_t s)
{
#define CASE(M) if (mask == (M)) return launch_forward_t<(M), EvalSphere>(a, s);
    CASE(0x008) CASE(0x010) CASE(0x020) CASE(0x040) CASE(0x080) CASE(0x100) CASE(0x200)
    CASE(0x3F0) CASE(0x3F8)
    FB200_MAGNETIC_SETS(CASE)
#undef CASE
    return cudaErrorInvalidValue;
}

cudaError_t launch_jacobian_sphere(int field, bool transposed, const JacobianArgs &a, cudaStream_t s)
{
    switch (field) {
#define CASE(F) case F: return launch_jacobian_t<(1u << F), EvalSphere>(a, transposed, s);
        CASE(3) CASE(4) CASE(5) CASE(6) CASE(7) CASE(8) CASE(9) CASE(10) CASE(11) CASE(12) CASE(13)
#undef CASE
    default: return cudaErrorInvalidValue;
    }
}

}  // namespace fb200
