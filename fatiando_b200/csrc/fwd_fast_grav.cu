// fwd_fast_grav.cu -- fused fast-path forward kernels, gravity component sets (K1).
#include "prism_masks.h"
#include "prism_fast.cuh"
#include "prism_forward.cuh"
#include "prism_launch.h"

namespace fb200 {

cudaError_t launch_forward_fast_grav(uint32_t mask, const ForwardArgs &a, cudaStream_t s)
{
#define CASE(M) if (mask == (M)) return launch_forward_t<(M), EvalFast>(a, s);
    FB200_GRAVITY_SETS(CASE)
#undef CASE
    return cudaErrorInvalidValue;
}

}  // namespace fb200
