// ab_interp_f32.cu — one instantiation of the SDF interpreter (kept in its own translation unit so the four
// variants compile in parallel): S = Pack<float, 4>, argument pool of float.
#define AB_INTERP_INSTANTIATE 1
#include "ab_interp.cuh"

namespace ab {
template cudaError_t launch_interp<Pack<float, 4>, float>(const KParams<float>&, const LaunchCfg&, cudaStream_t, int*);
}
