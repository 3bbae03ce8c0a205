// fwd_mesh.cu -- node-sharing forward kernels for regular meshes (prism_mesh.cuh).
#include "prism_masks.h"
#include "prism_forward.cuh"
#include "prism_mesh.cuh"
#include "prism_mesh_jacobian.cuh"
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

cudaError_t launch_mesh_jacobian(int field, const MeshJacArgs &a, cudaStream_t s)
{
    switch (field) {
#define CASE(F) case F: return launch_mesh_jacobian_t<(1u << F)>(a, s);
        CASE(0) CASE(1) CASE(2) CASE(3) CASE(4) CASE(5) CASE(6) CASE(7) CASE(8) CASE(9)
        CASE(10) CASE(11) CASE(12) CASE(13)
#undef CASE
    default: return cudaErrorInvalidValue;
    }
}

}  // namespace fb200
