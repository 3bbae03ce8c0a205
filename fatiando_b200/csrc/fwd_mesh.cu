// fwd_mesh.cu -- node-sharing forward kernels for regular meshes (prism_mesh.cuh).
#include "prism_masks.h"
#include "prism_forward.cuh"
#include "prism_mesh.cuh"
#include "prism_launch.h"

namespace fb200 {

cudaError_t launch_mesh_forward(uint32_t mask, const MeshArgs &a, cudaStream_t s)
{
#define CASE(M) if (mask == (M)) return launch_mesh_t<(M)>(a, s);
    FB200_GRAVITY_SETS(CASE)
    FB200_MAGNETIC_SETS(CASE)
#undef CASE
    return cudaErrorInvalidValue;
}

}  // namespace fb200
