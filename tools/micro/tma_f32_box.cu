// Bring-up probe: which fp32 non-swizzled 4-D TMA boxes does sm_100 accept?  usage: tma_f32_box W H box0 box1 box2 x y [c]
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <stdint.h>
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
__global__ void probe(const __grid_constant__ CUtensorMap tm, int bytes, int x, int y, int c, float* out) {
  extern __shared__ __align__(1024) uint8_t sm[];
  __shared__ uint64_t bar;
  uint32_t b = (uint32_t)__cvta_generic_to_shared(&bar), d = (uint32_t)__cvta_generic_to_shared(sm);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b));
    asm volatile("fence.mbarrier_init.release.cluster;");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes));
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                 ::"r"(d), "l"((uint64_t)&tm), "r"(b), "r"(x), "r"(y), "r"(c), "r"(0) : "memory");
    uint32_t ok = 0; int spins = 0;
    while (!ok && ++spins < (1 << 22))
      asm volatile("{.reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p;}" : "=r"(ok) : "r"(b));
    out[0] = ok ? reinterpret_cast<float*>(sm)[0] : -12345.f;
    out[1] = ok ? reinterpret_cast<float*>(sm)[bytes / 4 - 1] : -12345.f;
  }
}
int main(int argc, char** argv) {
  int W = atoi(argv[1]), H = atoi(argv[2]), b0 = atoi(argv[3]), b1 = atoi(argv[4]), b2 = atoi(argv[5]), x = atoi(argv[6]), y = atoi(argv[7]), c = argc > 8 ? atoi(argv[8]) : 0;
  float* d; cudaMalloc(&d, (size_t)2 * 3 * H * W * 4);
  float* h = (float*)malloc((size_t)2 * 3 * H * W * 4);
  for (size_t i = 0; i < (size_t)2 * 3 * H * W; ++i) h[i] = (float)i;
  cudaMemcpy(d, h, (size_t)2 * 3 * H * W * 4, cudaMemcpyHostToDevice);
  void* p = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
  CUtensorMap tm;
  cuuint64_t gdim[4] = {(cuuint64_t)W, (cuuint64_t)H, 3, 2}, gs[3] = {(cuuint64_t)W * 4, (cuuint64_t)H * W * 4, (cuuint64_t)3 * H * W * 4};
  cuuint32_t box[4] = {(cuuint32_t)b0, (cuuint32_t)b1, (cuuint32_t)b2, 1}, es[4] = {1, 1, 1, 1};
  CUresult r = ((EncodeTiledFn)p)(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, d, gdim, gs, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                  CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  int bytes = b0 * b1 * b2 * 4;
  float* out; cudaMalloc(&out, 8);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  probe<<<1, 32, bytes + 1024>>>(tm, bytes, x, y, c, out);
  cudaError_t e = cudaDeviceSynchronize();
  float res[2] = {0, 0}; cudaMemcpy(res, out, 8, cudaMemcpyDeviceToHost);
  printf("W=%d H=%d box=%dx%dx%d at (%d,%d,c=%d): encode=%d run=%s first=%.0f last=%.0f\n", W, H, b0, b1, b2, x, y, c, (int)r, cudaGetErrorString(e), res[0], res[1]);
  return 0;
}
