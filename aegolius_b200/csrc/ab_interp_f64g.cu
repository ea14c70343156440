// ab_interp_f64g.cu — one instantiation of the SDF interpreter (each variant sits in its own translation unit so
// that they compile in parallel): S = Dual<Pack<double, 1>, 3>, argument pool of double, tier 2 (full op set).
#define AB_INTERP_INSTANTIATE 1
#define AB_TIER_FULL 2
#include "ab_interp.cuh"

namespace ab {
template cudaError_t launch_interp<Dual<Pack<double, 1>, 3>, double, 2>(const KParams<double>&, const LaunchCfg&, cudaStream_t, int*);
}
