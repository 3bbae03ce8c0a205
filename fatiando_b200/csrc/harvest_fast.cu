// harvest_fast.cu -- fused fast-path effect rows.
#include "prism_fast.cuh"
#include "harvest.cuh"

namespace fb200 {

cudaError_t launch_harvest_effect_fast(int field, const EffectArgs &a, cudaStream_t s)
{
    return launch_effect<EvalFast>(field, a, s);
}

}  // namespace fb200
