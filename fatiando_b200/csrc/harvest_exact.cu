// harvest_exact.cu -- reference-order effect rows (compiled with -fmad=false).
#include "prism_exact.cuh"
#include "harvest.cuh"

namespace fb200 {

cudaError_t launch_harvest_effect_exact(int field, const EffectArgs &a, cudaStream_t s)
{
    return launch_effect<EvalExact>(field, a, s);
}

}  // namespace fb200
