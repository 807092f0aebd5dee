// Microbenchmark: tcgen05.ld throughput / latency as a function of load width and warps per TMEM lane quarter.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_ld_bench tmem_ld_bench.cu ; ./tmem_ld_bench
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int X>
__device__ __forceinline__ uint32_t ld(uint32_t taddr) {
  uint32_t v[X];
  if constexpr (X == 16) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
                   "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]) : "r"(taddr) : "memory");
  } else if constexpr (X == 32) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
                   "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]),
                   "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]),
                   "=r"(v[30]), "=r"(v[31]) : "r"(taddr) : "memory");
  } else {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]) : "r"(taddr) : "memory");
  }
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  uint32_t a = 0;
#pragma unroll
  for (int i = 0; i < X; ++i) a ^= v[i];
  return a;
}

// two x32 loads in flight, one wait (throughput rather than latency)
__device__ __forceinline__ uint32_t ld2x32(uint32_t ta, uint32_t tb) {
  uint32_t v[32], w[32];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
                 "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]),
                 "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]),
                 "=r"(v[30]), "=r"(v[31]) : "r"(ta) : "memory");
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
               : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]), "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7]), "=r"(w[8]), "=r"(w[9]),
                 "=r"(w[10]), "=r"(w[11]), "=r"(w[12]), "=r"(w[13]), "=r"(w[14]), "=r"(w[15]), "=r"(w[16]), "=r"(w[17]), "=r"(w[18]), "=r"(w[19]),
                 "=r"(w[20]), "=r"(w[21]), "=r"(w[22]), "=r"(w[23]), "=r"(w[24]), "=r"(w[25]), "=r"(w[26]), "=r"(w[27]), "=r"(w[28]), "=r"(w[29]),
                 "=r"(w[30]), "=r"(w[31]) : "r"(tb) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  uint32_t a = 0;
#pragma unroll
  for (int i = 0; i < 32; ++i) a ^= v[i] ^ w[i];
  return a;
}

// WPQ warps per lane quarter, each issuing REPS loads of width X (one wait per load)
template <int X, int WPQ, bool PIPE = false>
__global__ void bench(long long* out, uint32_t* sink) {
  __shared__ uint32_t tbase;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tbase)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t base = tbase + ((uint32_t)((warp & 3) * 32) << 16);
  constexpr int REPS = 64;
  uint32_t acc = 0;
  __syncthreads();
  const long long t0 = clock64();
#pragma unroll 4
  for (int r = 0; r < REPS; ++r) {
    if constexpr (PIPE) acc ^= ld2x32(base + ((r * 64) & 255), base + ((r * 64 + 32) & 255));
    else acc ^= ld<X>(base + ((r * X + (warp >> 2) * 64) & 255));
  }
  const long long t1 = clock64();
  __syncthreads();
  if (threadIdx.x == 0 && blockIdx.x == 0) { out[0] = t1 - t0; }
  if (acc == 0x12345678u) sink[0] = acc;
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tbase) : "memory");
}

template <int X, int WPQ, bool PIPE = false>
void run(long long* d, uint32_t* sink) {
  bench<X, WPQ, PIPE><<<148, 128 * WPQ>>>(d, sink);
  cudaDeviceSynchronize();
  bench<X, WPQ, PIPE><<<148, 128 * WPQ>>>(d, sink);
  cudaDeviceSynchronize();
  long long h = 0;
  cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
  const double per_load = (double)h / 64;
  const double bytes = 32.0 * X * 4 * WPQ * (PIPE ? 2 : 1);                 // bytes per lane quarter per "round" of loads
  printf("%sx%-3d warps/quarter %d: %7.1f cycles per load (warp 0), %6.1f B/clk per lane quarter, %6.1f B/clk per SM   [%s]\n", PIPE ? "2 in flight " : "", X, WPQ,
         per_load, bytes / per_load, 4 * bytes / per_load, cudaGetErrorString(cudaGetLastError()));
}

int main() {
  long long* d; uint32_t* sink;
  cudaMalloc(&d, 64); cudaMalloc(&sink, 64);
  run<8, 1>(d, sink); run<16, 1>(d, sink); run<32, 1>(d, sink);
  run<8, 2>(d, sink); run<16, 2>(d, sink); run<32, 2>(d, sink);
  run<16, 4>(d, sink); run<32, 4>(d, sink);
  run<32, 1, true>(d, sink); run<32, 2, true>(d, sink);
  return 0;
}
