// pdl.cuh -- programmatic dependent launch helpers (sm_90+), shared by the GEMM and attention headers.
#pragma once
#include <cuda_runtime.h>

namespace comic {

// Programmatic dependent launch (sm_90+): a kernel launched with the programmatic-stream-serialization attribute may begin
// while its predecessor in the stream is still running; it must call pdl_wait() before it touches anything the
// predecessor produced (or still reads), and pdl_launch_dependents() lets ITS successor start early in turn.  Both are
// no-ops for kernels launched the ordinary way.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
// 1: the decode-step kernels are launched as programmatic dependents (process-wide; comic_set_option).  Off by default:
// measured 3.5 % SLOWER on the graph-replayed 512-image step (27.8 k vs 28.8 k captions/s, profiles/r09c_bench512_pdl=*.json)
// -- with every kernel releasing its successor at entry, the CTAs of up to four later kernels queue for the SMs behind
// the running one, and the replayed graph already launches each node within ~1 us of its predecessor's end.
inline int& pdl_mode() { static int v = 0; return v; }
// <<<>>> replacement that adds the attribute when pdl_mode() is on
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = pdl_mode() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

}  // namespace comic
