// jac_exact.cu -- reference-order Jacobian tile writers (compiled with -fmad=false).
#include "prism_exact.cuh"
#include "prism_jacobian.cuh"
#include "prism_launch.h"

namespace fb200 {

cudaError_t launch_jacobian_exact(int field, bool transposed, const JacobianArgs &a, cudaStream_t s)
{
    switch (field) {
#define CASE(F) case F: return launch_jacobian_t<(1u << F), EvalExact>(a, transposed, s);
        CASE(0) CASE(1) CASE(2) CASE(3) CASE(4) CASE(5) CASE(6) CASE(7) CASE(8) CASE(9)
        CASE(10) CASE(11) CASE(12) CASE(13)
#undef CASE
    default: return cudaErrorInvalidValue;
    }
}

cudaError_t launch_adjoint_exact(int field, const AdjointArgs &a, int nsplit, cudaStream_t s)
{
    switch (field) {
#define CASE(F) case F: return launch_adjoint_t<(1u << F), EvalExact>(a, nsplit, s);
        CASE(0) CASE(1) CASE(2) CASE(3) CASE(4) CASE(5) CASE(6) CASE(7) CASE(8) CASE(9)
        CASE(10) CASE(11) CASE(12) CASE(13)
#undef CASE
    default: return cudaErrorInvalidValue;
    }
}

}  // namespace fb200
