// Measures issue throughput (cycles per warp instruction per SM sub-partition) of the instructions the attention
// kernel is built from, on the GPU it runs on.  Each kernel runs 8 independent dependency chains per thread so the
// pipe, not latency, is the limit; 16 warps per SM (4 per sub-partition), one CTA per SM.
#include <cstdio>
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s: %s\n", #x, cudaGetErrorString(e_)); return 1; } } while (0)

constexpr int kIters = 4096, kChains = 8;

#define DEF_KERNEL(NAME, BODY)                                                              \
  __global__ void NAME(float* out, long long* cyc) {                                        \
    float v[kChains];                                                                       \
    for (int i = 0; i < kChains; ++i) v[i] = 0.001f * (threadIdx.x + i) + 0.5f;             \
    __syncthreads();                                                                        \
    long long t0 = clock64();                                                               \
    for (int it = 0; it < kIters; ++it) {                                                   \
      _Pragma("unroll") for (int i = 0; i < kChains; ++i) { BODY }                          \
    }                                                                                       \
    long long t1 = clock64();                                                               \
    float s = 0.f;                                                                          \
    for (int i = 0; i < kChains; ++i) s += v[i];                                            \
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;                                         \
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;                                        \
  }

DEF_KERNEL(k_ex2, asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(v[i]));)
DEF_KERNEL(k_rcp, asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(v[i]));)
DEF_KERNEL(k_rsq, asm volatile("rsqrt.approx.ftz.f32 %0, %0;" : "+f"(v[i]));)
DEF_KERNEL(k_tanh, asm volatile("tanh.approx.f32 %0, %0;" : "+f"(v[i]));)
DEF_KERNEL(k_ex2h2, { unsigned u = __float_as_uint(v[i]); asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(u)); v[i] = __uint_as_float(u); })
DEF_KERNEL(k_tanhh2, { unsigned u = __float_as_uint(v[i]); asm volatile("tanh.approx.f16x2 %0, %0;" : "+r"(u)); v[i] = __uint_as_float(u); })
DEF_KERNEL(k_tanhbf2, { unsigned u = __float_as_uint(v[i]); asm volatile("tanh.approx.bf16x2 %0, %0;" : "+r"(u)); v[i] = __uint_as_float(u); })
DEF_KERNEL(k_ffma, asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(v[i]) : "f"(v[(i + 1) % kChains] ), "f"(1.0f));)
DEF_KERNEL(k_fmnmx, asm volatile("min.f32 %0, %0, %1;" : "+f"(v[i]) : "f"(30.0f));)
DEF_KERNEL(k_iadd, { unsigned u = __float_as_uint(v[i]); asm volatile("add.u32 %0, %0, %1;" : "+r"(u) : "r"(threadIdx.x)); v[i] = __uint_as_float(u); })

__global__ void k_ffma2(float* out, long long* cyc) {
  float2 v[kChains];
  for (int i = 0; i < kChains; ++i) v[i] = make_float2(0.001f * (threadIdx.x + i) + 0.5f, 0.25f);
  const float2 m = make_float2(0.999f, 1.001f), c = make_float2(1e-3f, 2e-3f);
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < kIters; ++it) {
#pragma unroll
    for (int i = 0; i < kChains; ++i) v[i] = __ffma2_rn(v[i], m, c);
  }
  long long t1 = clock64();
  float s = 0.f;
  for (int i = 0; i < kChains; ++i) s += v[i].x + v[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
// mixed: 4 FFMA2 per EX2 (does the XU pipe overlap with the FMA pipe?)
__global__ void k_mix(float* out, long long* cyc) {
  float2 v[kChains];
  float e[kChains];
  for (int i = 0; i < kChains; ++i) { v[i] = make_float2(0.001f * (threadIdx.x + i) + 0.5f, 0.25f); e[i] = 0.3f + i; }
  const float2 m = make_float2(0.999f, 1.001f), c = make_float2(1e-3f, 2e-3f);
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < kIters; ++it) {
#pragma unroll
    for (int i = 0; i < kChains; ++i) {
      asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(e[i]));
      v[i] = __ffma2_rn(v[i], m, c); v[i] = __ffma2_rn(v[i], m, c); v[i] = __ffma2_rn(v[i], m, c); v[i] = __ffma2_rn(v[i], m, c);
    }
  }
  long long t1 = clock64();
  float s = 0.f;
  for (int i = 0; i < kChains; ++i) s += v[i].x + v[i].y + e[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

int main() {
  int sms = 148;
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  float* out; long long* cyc;
  const int threads = 512;
  CK(cudaMalloc(&out, (size_t)sms * threads * 4)); CK(cudaMalloc(&cyc, sms * 8));
  struct { const char* name; void (*fn)(float*, long long*); double per_iter; } ks[] = {
    {"ex2.approx.ftz.f32", k_ex2, kChains}, {"rcp.approx.ftz.f32", k_rcp, kChains}, {"rsqrt.approx.ftz.f32", k_rsq, kChains},
    {"tanh.approx.f32", k_tanh, kChains}, {"ex2.approx.f16x2", k_ex2h2, kChains}, {"tanh.approx.f16x2", k_tanhh2, kChains},
    {"tanh.approx.bf16x2", k_tanhbf2, kChains}, {"fma.rn.f32 (3 reg)", k_ffma, kChains}, {"fma.rn.f32x2", k_ffma2, kChains},
    {"min.f32", k_fmnmx, kChains}, {"add.u32", k_iadd, kChains}, {"mix: 1 ex2 + 4 ffma2 (per group)", k_mix, kChains}};
  printf("%-36s %12s %s\n", "instruction", "cycles/CTA", "cycles per warp instruction per sub-partition (4 warps each)");
  for (auto& k : ks) {
    k.fn<<<sms, threads>>>(out, cyc);
    CK(cudaDeviceSynchronize());
    k.fn<<<sms, threads>>>(out, cyc);
    CK(cudaDeviceSynchronize());
    long long h[256];
    CK(cudaMemcpy(h, cyc, sms * 8, cudaMemcpyDeviceToHost));
    double avg = 0; for (int i = 0; i < sms; ++i) avg += (double)h[i]; avg /= sms;
    // per sub-partition: 4 warps each issue kIters * per_iter instructions
    printf("%-36s %12.0f %.2f\n", k.name, avg, avg / (4.0 * kIters * k.per_iter));
  }
  return 0;
}
