// fwd_exact.cu -- reference-order forward kernels (compiled with -fmad=false).
#include "prism_masks.h"
#include "prism_exact.cuh"
#include "prism_forward.cuh"
#include "prism_launch.h"

namespace fb200 {

cudaError_t launch_forward_exact(uint32_t mask, const ForwardArgs &a, cudaStream_t s)
{
#define CASE(M) if (mask == (M)) return launch_forward_t<(M), EvalExact>(a, s);
    FB200_GRAVITY_SETS(CASE)
    FB200_MAGNETIC_SETS(CASE)
#undef CASE
    return cudaErrorInvalidValue;
}

}  // namespace fb200
