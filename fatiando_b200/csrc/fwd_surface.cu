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

cudaError_t launch_forward_node_rev(int field, const ForwardArgs &a, int grid_y, cudaStream_t s)
{
    switch (field) {
#define CASE(F) case F: return launch_forward_t<(1u << F), EvalNodeRev>(a, s, grid_y);
        CASE(0) CASE(1) CASE(2) CASE(3) CASE(4) CASE(5) CASE(6) CASE(7) CASE(8) CASE(9)
        CASE(10) CASE(11) CASE(12) CASE(13)
#undef CASE
    default: return cudaErrorInvalidValue;
    }
}

}  // namespace fb200
