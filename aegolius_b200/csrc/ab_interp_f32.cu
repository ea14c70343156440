// ab_interp_f32.cu — one instantiation of the SDF interpreter (each variant sits in its own translation unit so
// that they compile in parallel): S = Pack<float, 4>, argument pool of float, tier 2 (full op set).
#define AB_INTERP_INSTANTIATE 1
#define AB_TIER_FULL 2
#include "ab_interp.cuh"

namespace ab {
template cudaError_t launch_interp<Pack<float, 4>, float, 2>(const KParams<float>&, const LaunchCfg&, cudaStream_t, int*);
}
