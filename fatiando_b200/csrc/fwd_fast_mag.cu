// fwd_fast_mag.cu -- fused fast-path forward kernels, magnetic component sets (K2).
#include "prism_masks.h"
#include "prism_fast.cuh"
#include "prism_forward.cuh"
#include "prism_launch.h"

namespace fb200 {

cudaError_t launch_forward_fast_mag(uint32_t mask, const ForwardArgs &a, cudaStream_t s)
{
#define CASE(M) if (mask == (M)) return launch_forward_t<(M), EvalFast>(a, s);
    FB200_MAGNETIC_SETS(CASE)
#undef CASE
    return cudaErrorInvalidValue;
}

}  // namespace fb200
