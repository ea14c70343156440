// ab_interp_f32p.cu — parameter-tangent interpreter (AB_GRAD_PARAM): S = Dual<Pack<float, 2>, 1>, every argument read as a
// dual number carrying d arg / d theta; full op set.
#define AB_INTERP_INSTANTIATE 1
#define AB_TIER_FULL 2
#include "ab_interp.cuh"

namespace ab {
template cudaError_t launch_interp<Dual<Pack<float, 2>, 1>, float, 2, true>(const KParams<float>&, const LaunchCfg&, cudaStream_t, int*);
}
