// prism_launch.h -- host-side entry points of the kernel translation units.
#pragma once
#include <cuda_runtime.h>
#include "prism_common.cuh"
#include "prism_jacobian.cuh"

namespace fb200 { struct MeshArgs; struct MeshJacArgs; struct GridPlan; }

namespace fb200 {

// Each returns cudaErrorInvalidValue when `mask` is not one of the instantiated
// component sets of prism_masks.h (callers decompose requests first).
cudaError_t launch_forward_exact(uint32_t mask, const ForwardArgs &a, cudaStream_t s);      // fwd_exact.cu
cudaError_t launch_forward_fast_grav(uint32_t mask, const ForwardArgs &a, cudaStream_t s);  // fwd_fast_grav.cu
cudaError_t launch_forward_fast_mag(uint32_t mask, const ForwardArgs &a, cudaStream_t s);   // fwd_fast_mag.cu
cudaError_t launch_jacobian_exact(int field, bool transposed, const JacobianArgs &a, cudaStream_t s);  // jac_exact.cu
cudaError_t launch_jacobian_fast(int field, bool transposed, const JacobianArgs &a, cudaStream_t s);   // jac_fast.cu

cudaError_t launch_adjoint_exact(int field, const AdjointArgs &a, int nsplit, cudaStream_t s);   // jac_exact.cu
cudaError_t launch_adjoint_fast(int field, const AdjointArgs &a, int nsplit, cudaStream_t s);    // jac_fast.cu

cudaError_t launch_forward_sphere(uint32_t mask, const ForwardArgs &a, cudaStream_t s);                  // fwd_sphere.cu
cudaError_t launch_jacobian_sphere(int field, bool transposed, const JacobianArgs &a, cudaStream_t s);  // fwd_sphere.cu
cudaError_t launch_forward_polyprism(uint32_t mask, const ForwardArgs &a, cudaStream_t s);               // fwd_polyprism_exact.cu
cudaError_t launch_forward_face(uint32_t mask, const ForwardArgs &a, int grid_y, cudaStream_t s);   // fwd_surface.cu
cudaError_t launch_forward_node(uint32_t mask, const ForwardArgs &a, int grid_y, cudaStream_t s);   // fwd_surface.cu
cudaError_t launch_forward_node_rev(int field, const ForwardArgs &a, int grid_y, cudaStream_t s);    // fwd_surface.cu
cudaError_t launch_mesh_forward(uint32_t mask, const MeshArgs &a, cudaStream_t s);               // fwd_mesh.cu
cudaError_t launch_mesh_jacobian(int field, const MeshJacArgs &a, cudaStream_t s);              // fwd_mesh.cu
// commensurate-grid Jacobian (prism_grid_reuse.cuh): node + cell tables, then rows [row0, row0 + nrows)
cudaError_t launch_grid_tables(int field, const GridPlan &g, const double *xu, const double *yv, const double *zc,
                               const double *unit_p, const double *fvec, const double *shift, double *T,
                               double *D0, double *D1, cudaStream_t s);                          // jac_grid.cu
cudaError_t launch_grid_expand(const GridPlan &g, const double *D0, const double *D1, int64_t row0, int64_t nrows,
                               double *out, int64_t ld, cudaStream_t s);                         // jac_grid.cu

}  // namespace fb200
