// Pipe-throughput microbenchmark for sm_100a (B200): measures warp-instructions per clock per SM
// for the instruction classes the SDF interpreter is built from. Not part of the product path;
// its numbers are recorded in DESIGN.md and drive the Pack<T,W> design (scalar vs f32x2 packed).
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#define NCH 8
#define ITERS 4096

enum Kind { K_FFMA, K_FFMA2, K_FADD, K_FADD2, K_FMUL, K_FMNMX, K_MIX_FFMA_FMNMX, K_MIX_FFMA2_FMNMX,
            K_MUFU_RSQ, K_MUFU_SIN, K_MUFU_EX2, K_SQRT_RN, K_SQRT_APPROX, K_DFMA, K_DADD, K_IMAD, K_IADD3,
            K_LDS128, K_FSEL, K_MIX_FFMA_IADD, K_DIV_RN, K_DIV_APPROX, K_SINF, K_ATAN2F, K_NKINDS };
static const char* names[] = {"FFMA","FFMA2","FADD","FADD2","FMUL","FMNMX","FFMA+FMNMX","FFMA2+FMNMX",
  "MUFU.RSQ","MUFU.SIN","MUFU.EX2","sqrtf(rn)","sqrt.approx","DFMA","DADD","IMAD","IADD3","LDS.128 bcast","FSEL(cmp+sel)",
  "FFMA+IADD","div(rn)","div.approx","sinf","atan2f"};

template <int K>
__global__ void __launch_bounds__(256) bench(float* out, float seed, unsigned long long* cyc) {
  __shared__ float4 sm[64];
  if (threadIdx.x < 64) sm[threadIdx.x] = make_float4(seed, seed*2, seed*3, seed*4);
  __syncthreads();
  float a[NCH]; float2 p[NCH]; double d[NCH]; int q[NCH];
  #pragma unroll
  for (int i = 0; i < NCH; i++) { a[i] = seed + i + threadIdx.x * 1e-3f; p[i] = make_float2(a[i], a[i]+1); d[i] = a[i]; q[i] = i + threadIdx.x; }
  float b = seed * 0.999f, c = seed * 1e-3f; float2 b2 = make_float2(b, b), c2 = make_float2(c, c);
  double db = b, dc = c;
  unsigned long long t0 = clock64();
  #pragma unroll 1
  for (int it = 0; it < ITERS; it++) {
    #pragma unroll
    for (int i = 0; i < NCH; i++) {
      if (K == K_FFMA) a[i] = fmaf(a[i], b, c);
      if (K == K_FFMA2) p[i] = __ffma2_rn(p[i], b2, c2);
      if (K == K_FADD) a[i] = a[i] + b;
      if (K == K_FADD2) p[i] = __fadd2_rn(p[i], b2);
      if (K == K_FMUL) a[i] = a[i] * b;
      if (K == K_FMNMX) a[i] = fminf(a[i], b + i);
      if (K == K_MIX_FFMA_FMNMX) { a[i] = fmaf(a[i], b, c); p[i].x = fminf(p[i].x, a[(i+1)%NCH]); }
      if (K == K_MIX_FFMA2_FMNMX) { p[i] = __ffma2_rn(p[i], b2, c2); a[i] = fminf(a[i], p[(i+1)%NCH].x); }
      if (K == K_MUFU_RSQ) a[i] = rsqrtf(a[i]);   // may include fixup; see SASS
      if (K == K_MUFU_SIN) a[i] = __sinf(a[i]);
      if (K == K_MUFU_EX2) a[i] = exp2f(a[i]);
      if (K == K_SQRT_RN) a[i] = sqrtf(a[i]);
      if (K == K_SQRT_APPROX) { float r; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(a[i])); a[i] = r; }
      if (K == K_DFMA) d[i] = fma(d[i], db, dc);
      if (K == K_DADD) d[i] = d[i] + db;
      if (K == K_IMAD) q[i] = q[i] * 3 + it;
      if (K == K_IADD3) q[i] = q[i] + it + (q[(i+1)%NCH] & 1);
      if (K == K_LDS128) { float4 v = sm[(q[i] + it) & 63]; a[i] += v.x; q[i] += __float_as_int(v.y) & 1; }
      if (K == K_FSEL) a[i] = (a[i] > b) ? a[i] - 1.0f : c;
      if (K == K_MIX_FFMA_IADD) { a[i] = fmaf(a[i], b, c); q[i] = q[i] + it; }
      if (K == K_DIV_RN) a[i] = b / a[i];
      if (K == K_DIV_APPROX) a[i] = __fdividef(b, a[i]);
      if (K == K_SINF) a[i] = sinf(a[i]);
      if (K == K_ATAN2F) a[i] = atan2f(a[i], b);
    }
  }
  unsigned long long t1 = clock64();
  float s = 0; 
  #pragma unroll
  for (int i = 0; i < NCH; i++) s += a[i] + p[i].x + p[i].y + (float)d[i] + q[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int K> void run(float* out, unsigned long long* cyc, int sms) {
  int blocks_per_sm = 4, nb = sms * blocks_per_sm;   // 4 x 256 threads = 32 warps/SM = 8 per SMSP
  bench<K><<<nb, 256>>>(out, 1.0001f, cyc); cudaDeviceSynchronize();
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0); bench<K><<<nb, 256>>>(out, 1.0001f, cyc); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  unsigned long long h[4096]; cudaMemcpy(h, cyc, nb * 8, cudaMemcpyDeviceToHost);
  double mean = 0; for (int i = 0; i < nb; i++) mean += h[i]; mean /= nb;
  double src_ops = (double)ITERS * NCH;               // source-level ops per thread
  double warp_inst_per_sm = src_ops * (256 / 32) * blocks_per_sm;
  printf("%-16s cycles/block %.0f  src-op warp-inst/clk/SM %.3f  (per SMSP %.3f)  ms %.3f  clk_MHz~%.0f\n", names[K], mean,
         warp_inst_per_sm / mean, warp_inst_per_sm / mean / 4, ms, mean / (ms * 1e3));
}
template <int K> struct Runner { static void go(float* o, unsigned long long* c, int s) { run<K>(o, c, s); Runner<K+1>::go(o, c, s); } };
template <> struct Runner<K_NKINDS> { static void go(float*, unsigned long long*, int) {} };

int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  printf("device %s sms %d cc %d.%d smem/blk optin %zu L2 %d MB clock %d kHz\n", p.name, p.multiProcessorCount, p.major, p.minor,
         p.sharedMemPerBlockOptin, p.l2CacheSize >> 20, p.clockRate);
  float* out; unsigned long long* cyc; cudaMalloc(&out, 148 * 8 * 256 * 4 * 4); cudaMalloc(&cyc, 4096 * 8);
  Runner<0>::go(out, cyc, p.multiProcessorCount);
  printf("note: 'src-op' counts source-level ops; multi-instruction ops (sqrtf, div, sinf, atan2f) show ops/clk, not SASS inst/clk\n");
  return 0;
}
