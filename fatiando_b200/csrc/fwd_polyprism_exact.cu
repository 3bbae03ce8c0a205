// fwd_polyprism_exact.cu -- polygonal-prism forward kernels (polyprism.cuh); compiled with
// -fmad=false like the other *_exact.cu units: every product and sum rounds separately.
#include "prism_masks.h"
#include "polyprism.cuh"
#include "prism_forward.cuh"
#include "prism_launch.h"

namespace fb200 {

cudaError_t launch_forward_polyprism(uint32_t mask, const ForwardArgs &a, cudaStream_t s)
{
#define CASE(M) if (mask == (M)) return launch_forward_t<(M), EvalPolyEdge>(a, s);
    CASE(0x008) CASE(0x010) CASE(0x020) CASE(0x040) CASE(0x080) CASE(0x100) CASE(0x200)
    CASE(0x3F0) CASE(0x3F8)
    FB200_MAGNETIC_SETS(CASE)
#undef CASE
    return cudaErrorInvalidValue;
}

}  // namespace fb200
