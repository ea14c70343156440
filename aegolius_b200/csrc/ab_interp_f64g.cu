// ab_interp_f64g.cu — one instantiation of the SDF interpreter (kept in its own translation unit so the four
// variants compile in parallel): S = Dual<Pack<double, 1>, 3>, argument pool of double.
#define AB_INTERP_INSTANTIATE 1
#include "ab_interp.cuh"

namespace ab {
template cudaError_t launch_interp<Dual<Pack<double, 1>, 3>, double>(const KParams<double>&, const LaunchCfg&, cudaStream_t, int*);
}
