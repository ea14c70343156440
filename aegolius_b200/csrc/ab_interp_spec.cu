// ab_interp_spec.cu — a PROGRAM-SPECIALISED build of the same interpreter: compiled on request (build.py:
// build_specialized) with -DAB_SPEC_<OP>=0 for every op the program does not use, so the kernel holds only the op bodies
// it executes (C3 tree: 36 KB of SASS instead of 94 KB, which is what fits the instruction cache; see
// profiles/r01_sweeps.md). The result is a separate shared object exporting one launcher, registered with the main
// library through ab_spec_register. Not part of the default build.
//   AB_SPEC_KIND: 0 fp32 values, 1 fp32 values + spatial gradient, 2 fp64 values, 3 fp64 values + spatial gradient,
//                 4 fp32 parameter tangent (AB_GRAD_PARAM, also the fused loss mode), 5 fp64 parameter tangent
#define AB_INTERP_INSTANTIATE 1
#define AB_TIER_FULL 2
#include "ab_interp.cuh"

#ifndef AB_SPEC_KIND
#error "AB_SPEC_KIND must be defined"
#endif

namespace ab {
#if AB_SPEC_KIND == 0
typedef float SpecT;
typedef Pack<float, 4> SpecS;
#elif AB_SPEC_KIND == 1
typedef float SpecT;
typedef Dual<Pack<float, 2>, 3> SpecS;
#elif AB_SPEC_KIND == 2
typedef double SpecT;
typedef Pack<double, 2> SpecS;
#elif AB_SPEC_KIND == 3
typedef double SpecT;
typedef Dual<Pack<double, 1>, 3> SpecS;
#elif AB_SPEC_KIND == 4
typedef float SpecT;
typedef Dual<Pack<float, 2>, 1> SpecS;
#else
typedef double SpecT;
typedef Dual<Pack<double, 1>, 1> SpecS;
#endif
constexpr bool kSpecParam = AB_SPEC_KIND >= 4;
template cudaError_t launch_interp<SpecS, SpecT, 2, kSpecParam>(const KParams<SpecT>&, const LaunchCfg&, cudaStream_t, int*);
}  // namespace ab

#define AB_SPEC_EXPORT extern "C" __attribute__((visibility("default")))

// returns the cudaError_t of the launch; *status as launch_interp sets it
AB_SPEC_EXPORT int ab_spec_launch(const void* kparams, int sms, unsigned long long smem_optin, void* stream, int* status) {
  ab::LaunchCfg cfg{sms, (size_t)smem_optin};
  return (int)ab::launch_interp<ab::SpecS, ab::SpecT, 2, ab::kSpecParam>(*reinterpret_cast<const ab::KParams<ab::SpecT>*>(kparams), cfg,
                                                        (cudaStream_t)stream, status);
}
// layout guard: the main library refuses a specialisation built against another KParams
AB_SPEC_EXPORT unsigned long long ab_spec_kparams_size(void) { return sizeof(ab::KParams<ab::SpecT>); }
AB_SPEC_EXPORT int ab_spec_kind(void) { return AB_SPEC_KIND; }
