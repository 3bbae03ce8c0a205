// jac_grid.cu -- kernels of the commensurate-grid Jacobian (prism_grid_reuse.cuh).
#include "prism_grid_reuse.cuh"
#include "prism_launch.h"

namespace fb200 {

cudaError_t launch_grid_tables(int field, const GridPlan &g, const double *xu, const double *yv, const double *zc,
                               const double *unit_p, const double *fvec, const double *shift, double *T,
                               double *D0, double *D1, cudaStream_t s)
{
    const int64_t cols = g.nu * g.nv;
    const unsigned blocks = (unsigned)((cols + 127) / 128);
    switch (field) {
#define CASE(F)                                                                                         \
    case F:                                                                                             \
        node_table_kernel<(1u << F)><<<blocks, 128, 0, s>>>(g, xu, yv, zc, unit_p[0], unit_p[1], unit_p[2], \
                                                            fvec[0], fvec[1], fvec[2], shift[0], shift[1],  \
                                                            shift[2], T);                                  \
        break;
        CASE(0) CASE(1) CASE(2) CASE(3) CASE(4) CASE(5) CASE(6) CASE(7) CASE(8) CASE(9)
        CASE(10) CASE(11) CASE(12) CASE(13)
#undef CASE
    default: return cudaErrorInvalidValue;
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    cell_table_kernel<<<dim3((unsigned)((g.ndu + 127) / 128), (unsigned)g.ndv), 128, 0, s>>>(g, T, D0, D1);
    return cudaGetLastError();
}

cudaError_t launch_grid_expand(const GridPlan &g, const double *D0, const double *D1, int64_t row0, int64_t nrows,
                               double *out, int64_t ld, cudaStream_t s)
{
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(expand_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             2 * GR_CHUNK_BYTES);
        if (e != cudaSuccess) return e;
        configured = true;
    }
    expand_rows_kernel<<<(unsigned)nrows, 32, 2 * GR_CHUNK_BYTES, s>>>(g, D0, D1, row0, nrows, out, ld);
    return cudaGetLastError();
}

}  // namespace fb200
