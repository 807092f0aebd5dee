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
int make_tmap_nd(struct CUtensorMap_st* tm, const void* base, int rank, const int64_t* dims, const int64_t* strides, const int* box);
int make_tmap_4d(struct CUtensorMap_st* tm, const void* base, const int64_t dims[4], const int64_t strides[3], const int box[4]);
int make_tmap_f32_4d(struct CUtensorMap_st* tm, const void* base, const int64_t dims[4], const int64_t strides[3], const int box[4]);
int num_sms();

// Raises the dynamic shared-memory cap of `Kernel` once per DEVICE (the attribute belongs to the kernel's per-device
// function state; a process that drives several GPUs must set it on each).  One flag word per kernel instantiation.
template <auto Kernel>
inline int set_dyn_smem(int bytes, const char* what) {
  static unsigned long long done = 0ull;                       // bit d: set on device d (benign race: the call is idempotent)
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev > 63) dev = 63;
  if ((done >> dev) & 1ull) return CFFM_OK;
  const cudaError_t e = cudaFuncSetAttribute(Kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e != cudaSuccess) {
    set_error("%s: cudaFuncSetAttribute(%d bytes): %s", what, bytes, cudaGetErrorString(e));
    return -(int)e;
  }
  if (dev < 63) done |= 1ull << dev;
  return CFFM_OK;
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// Exact-erf GELU, 0.5 x (1 + erf(x / sqrt 2)) (nn.GELU() default), with
//   erfc(z) = 2^(-z P5(z)),  z = min(|x| / sqrt 2, 4):  weighted-minimax fit, max abs error 3.1e-7 in fp32
// (tools/fit_erf.py) -- one bare MUFU.EX2 and 7 FMAs instead of erff()'s two-branch evaluation.  The 1/sqrt 2,
// the sign and the factor 0.5 are folded into the coefficients (Q(a) = -P5(a / sqrt 2) / sqrt 2, exponent
// a Q(a) - 1), and  y = max(x, 0) - |x| 2^(a Q(a) - 1)  needs no select: 10 instructions per value.
// Max abs error of y over [-8, 8]: 5.7e-7, three orders of magnitude below the fp16 rounding of the result.
__device__ __forceinline__ float gelu_erf(float x) {
  const float ax = fabsf(x);
  const float a = fminf(ax, 5.65685425f);
  float q = 1.775544843e-05f;
  q = fmaf(q, a, -6.477595889e-04f);
  q = fmaf(q, a, 7.724047638e-03f);
  q = fmaf(q, a, -5.292674154e-02f);
  q = fmaf(q, a, -4.590827227e-01f);
  q = fmaf(q, a, -1.151116848e+00f);
  float e;                                              // exponent in [-28, -1]: the bare MUFU.EX2 needs no range scaling
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(fmaf(q, a, -1.0f)));
  return fmaf(-ax, e, fmaxf(x, 0.f));
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// 16-byte vector of 8 halves.  The payload is ONE uint4 so that a copy through a pointer is a single 128-bit
// access (a struct of four __half2 is copied member by member: four 32-bit loads, 4x the memory instructions).
struct __align__(16) half8 {
  uint4 u;
};

__device__ __forceinline__ half8 zero_half8() {
  half8 v;
  v.u = make_uint4(0u, 0u, 0u, 0u);
  return v;
}

__device__ __forceinline__ void unpack8(const half8& v, float* f) {
  const uint32_t w[4] = {v.u.x, v.u.y, v.u.z, v.u.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 t = __half22float2(*reinterpret_cast<const __half2*>(&w[i]));
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}

__device__ __forceinline__ uint32_t pack_half2(float a, float b) {
  __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

__device__ __forceinline__ half8 pack8(const float* f) {
  half8 v;
  v.u = make_uint4(pack_half2(f[0], f[1]), pack_half2(f[2], f[3]), pack_half2(f[4], f[5]), pack_half2(f[6], f[7]));
  return v;
}

}  // namespace cffm
