// ab_interp_f64_lite.cu — one instantiation of the SDF interpreter (each variant sits in its own translation unit so
// that they compile in parallel): S = Pack<double, 2>, argument pool of double, tier 0 (lite op set: no transcendentals, 40 registers).
#define AB_INTERP_INSTANTIATE 1
#define AB_TIER_FULL 0
#include "ab_interp.cuh"

namespace ab {
template cudaError_t launch_interp<Pack<double, 2>, double, 0>(const KParams<double>&, const LaunchCfg&, cudaStream_t, int*);
}
