// ab_interp_f64.cu — one instantiation of the SDF interpreter (each variant sits in its own translation unit so
// that they compile in parallel): S = Pack<double, 2>, argument pool of double, tier 2 (full op set).
#define AB_INTERP_INSTANTIATE 1
#define AB_TIER_FULL 2
#include "ab_interp.cuh"

namespace ab {
template cudaError_t launch_interp<Pack<double, 2>, double, 2>(const KParams<double>&, const LaunchCfg&, cudaStream_t, int*);
}
