// ab_interp_f64.cu — one instantiation of the SDF interpreter (kept in its own translation unit so the four
// variants compile in parallel): S = Pack<double, 2>, argument pool of double.
#define AB_INTERP_INSTANTIATE 1
#include "ab_interp.cuh"

namespace ab {
template cudaError_t launch_interp<Pack<double, 2>, double>(const KParams<double>&, const LaunchCfg&, cudaStream_t, int*);
}
