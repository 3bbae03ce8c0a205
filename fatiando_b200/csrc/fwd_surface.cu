// fwd_surface.cu -- relief surface form: top-face and reference-plane-node kernels
// (prism_face.cuh), gravity component sets.
#include "prism_masks.h"
#include "prism_forward.cuh"
#include "prism_face.cuh"
#include "prism_launch.h"

namespace fb200 {

cudaError_t launch_forward_face(uint32_t mask, const ForwardArgs &a, int grid_y, cudaStream_t s)
{
#define CASE(M) if (mask == (M)) return launch_forward_t<(M), EvalFace>(a, s, grid_y);
    FB200_GRAVITY_SETS(CASE)
#undef CASE
    return cudaErrorInvalidValue;
}

cudaError_t launch_forward_node(uint32_t mask, const ForwardArgs &a, int grid_y, cudaStream_t s)
{
#define CASE(M) if (mask == (M)) return launch_forward_t<(M), EvalNode>(a, s, grid_y);
    FB200_GRAVITY_SETS(CASE)
#undef CASE
    return cudaErrorInvalidValue;
}

}  // namespace fb200
