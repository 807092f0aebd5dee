// Shared helpers for libcffm_b200.so (sm_100a only).
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include "../../include/cffm_b200.h"

namespace cffm {

void set_error(const char* fmt, ...);

#define CFFM_REQUIRE(cond, code, ...)     \
  do {                                    \
    if (!(cond)) {                        \
      ::cffm::set_error(__VA_ARGS__);     \
      return (code);                      \
    }                                     \
  } while (0)

// Converts the launch status of the kernel just enqueued into the ABI return value.
inline int launch_status(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: %s", what, cudaGetErrorString(e));
    return -(int)e;
  }
  return CFFM_OK;
}

// ---- programmatic dependent launch (PDL).  Every kernel of this library is launched with the
// programmatic-stream-serialization attribute and starts with pdl_sync(): it lets the NEXT kernel's CTAs be
// scheduled (launch latency + prologue overlap this kernel's execution) and then blocks until the PREVIOUS
// kernel has completed and flushed, before the first global-memory access.  Works inside CUDA-graph capture.
// CFFM_PDL=0 in the environment falls back to plain stream-ordered launches.
__device__ __forceinline__ void pdl_sync() {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
}

bool pdl_enabled();

template <typename... KArgs, typename... Args>
inline void launch_k(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);   // status is read by launch_status()
}

// 2-D fp16 row-major [rows, K] tensor map, row stride ld (elements); box = [box_rows][64], 128-byte swizzle.
// `tm` is a CUtensorMap* (kept as void* here so that this header does not pull in <cuda.h>).
int make_tmap(struct CUtensorMap_st* tm, const void* base, int64_t rows, int64_t K, int64_t ld, int box_rows);
int num_sms();

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// Exact-erf GELU, 0.5 x (1 + erf(x / sqrt 2)) (nn.GELU() default), with
//   erf(z) = 1 - 2^(-z P5(z)),  z = min(|z|, 4):  weighted-minimax fit, max abs error 3.1e-7 in fp32
// (tools/fit_erf.py) -- one MUFU.EX2 and 7 FMAs instead of erff()'s two-branch evaluation.  The error
// in y is <= 0.5 |x| 3.1e-7, three orders of magnitude below the fp16 rounding of the result.
__device__ __forceinline__ float gelu_erf(float x) {
  const float a = fminf(fabsf(x) * 0.70710678118654752440f, 4.0f);
  float p = -1.42043592e-04f;
  p = fmaf(p, a, 3.66428169e-03f);
  p = fmaf(p, a, -3.08961913e-02f);
  p = fmaf(p, a, 1.49699434e-01f);
  p = fmaf(p, a, 9.18165470e-01f);
  p = fmaf(p, a, 1.62792507e+00f);
  const float t = 0.5f * x * exp2f(-p * a);            // 0.5 x erfc(|z|)
  return x >= 0.f ? x - t : t;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// 16-byte vector of 8 halves
struct __align__(16) half8 {
  __half2 h[4];
};

__device__ __forceinline__ void unpack8(const half8& v, float* f) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float2 t = __half22float2(v.h[i]);
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}

__device__ __forceinline__ half8 pack8(const float* f) {
  half8 v;
#pragma unroll
  for (int i = 0; i < 4; ++i) v.h[i] = __floats2half2_rn(f[2 * i], f[2 * i + 1]);
  return v;
}

__device__ __forceinline__ uint32_t pack_half2(float a, float b) {
  __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

}  // namespace cffm
