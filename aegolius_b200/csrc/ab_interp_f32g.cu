// ab_interp_f32g.cu — one instantiation of the SDF interpreter (kept in its own translation unit so the four
// variants compile in parallel): S = Dual<Pack<float, 2>, 3>, argument pool of float.
#define AB_INTERP_INSTANTIATE 1
#include "ab_interp.cuh"

namespace ab {
template cudaError_t launch_interp<Dual<Pack<float, 2>, 3>, float>(const KParams<float>&, const LaunchCfg&, cudaStream_t, int*);
}
