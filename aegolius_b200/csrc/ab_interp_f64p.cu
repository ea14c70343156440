// ab_interp_f64p.cu — parameter-tangent interpreter (AB_GRAD_PARAM): S = Dual<Pack<double, 1>, 1>, every argument read as a
// dual number carrying d arg / d theta; full op set.
#define AB_INTERP_INSTANTIATE 1
#define AB_TIER_FULL 2
#include "ab_interp.cuh"

namespace ab {
template cudaError_t launch_interp<Dual<Pack<double, 1>, 1>, double, 2, true>(const KParams<double>&, const LaunchCfg&, cudaStream_t, int*);
}
