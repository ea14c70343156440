// ab_interp_f32g_lite.cu — one instantiation of the SDF interpreter (each variant sits in its own translation unit so
// that they compile in parallel): S = Dual<Pack<float, 4>, 3>, argument pool of float, tier 0 (lite op set: no transcendentals, 40 registers).
#define AB_INTERP_INSTANTIATE 1
#define AB_TIER_FULL 0
#include "ab_interp.cuh"

namespace ab {
template cudaError_t launch_interp<Dual<Pack<float, 4>, 3>, float, 0>(const KParams<float>&, const LaunchCfg&, cudaStream_t, int*);
}
