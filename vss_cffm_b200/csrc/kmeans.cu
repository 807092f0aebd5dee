// Lloyd iterations for the per-video prototype generation of CFFM++ (cffm_head.py:267-294 calls
// fast_pytorch_kmeans.KMeans(n_clusters, max_iter=10, mode='euclidean').fit_predict on the 1/8-scale decoder features).
// The two contractions of an iteration run on the tensor cores through cffm_gemm_f16 / cffm_gemm_f16_splitk:
//   scores  S = X C^T           X fp16 [Np, E] (the features are produced in fp16), C split into hi + lo fp16 halves
//   sums      = onehot(labels) X  one-hot fp16 [Kp, Np] times X^T-major [E, Np], fp32 accumulation (exact products)
// and the three kernels below do the rest: arg max + one-hot + counts, centroid update + error, and the transpose.
#include "common.cuh"

namespace cffm {
namespace {

// closest centroid of every point: arg max_j (2 S[p][j] - |c_j|^2)  (= arg max of the library's
// 2ab - |a|^2 - |b|^2: the |a|^2 term does not depend on j), first maximum wins.  One CTA = 256 points.
__global__ void __launch_bounds__(256)
kmeans_assign_kernel(const float* __restrict__ S, int64_t ldS, const float* __restrict__ cnorm, int Np, int Np_pad, int K,
                     int Kp, int64_t* __restrict__ labels, __half* __restrict__ mask, int* __restrict__ counts) {
  pdl_sync();
  __shared__ int lab[256];
  __shared__ int hist[256];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int p0 = blockIdx.x * 256;
  hist[threadIdx.x] = 0;
  for (int i = 0; i < 32; ++i) {
    const int p = p0 + warp * 32 + i;                          // warp-uniform
    float best = -INFINITY;
    int arg = 0x7fffffff;
    if (p < Np) {
      for (int j = lane; j < K; j += 32) {
        const float v = 2.f * S[static_cast<int64_t>(p) * ldS + j] - cnorm[j];
        if (v > best) { best = v; arg = j; }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, best, o);
        const int oa = __shfl_xor_sync(0xffffffffu, arg, o);
        if (ov > best || (ov == best && oa < arg)) { best = ov; arg = oa; }
      }
      if (arg == 0x7fffffff) arg = 0;                          // all scores NaN: the library's max() would also return 0
    }
    if (lane == 0) lab[warp * 32 + i] = p < Np ? arg : -1;
  }
  __syncthreads();
  const int p = p0 + threadIdx.x;
  const int mine = lab[threadIdx.x];
  if (mine >= 0) {
    labels[p] = mine;
    atomicAdd(&hist[mine], 1);
  }
  if (p < Np_pad) {
    for (int j = 0; j < Kp; ++j) mask[static_cast<int64_t>(j) * Np_pad + p] = __float2half_rn(j == mine ? 1.f : 0.f);
  }
  __syncthreads();
  if (threadIdx.x < K && hist[threadIdx.x] > 0) atomicAdd(&counts[threadIdx.x], hist[threadIdx.x]);
}

// c_new[j] = sum_s partials[s][j] / count[j] (0 for an empty cluster: the library zeroes the NaN), error = sum (c_new - c)^2,
// centroids <- c_new; also the fp16 hi / lo halves and |c|^2 for the next assignment.  One CTA per (padded) cluster.
__global__ void __launch_bounds__(256)
kmeans_update_kernel(const float* __restrict__ partials, int nsplit, int64_t split_stride, const int* __restrict__ counts,
                     float* __restrict__ centroids, __half* __restrict__ c_hi, __half* __restrict__ c_lo,
                     float* __restrict__ cnorm, float* __restrict__ err_per_cluster, float* __restrict__ error,
                     unsigned int* __restrict__ done, int K, int Kp, int E) {
  pdl_sync();
  __shared__ float red[2][8];
  __shared__ bool last;
  const int j = blockIdx.x;
  float e2 = 0.f, n2 = 0.f;
  for (int c = threadIdx.x; c < E; c += 256) {
    float cn = 0.f;
    if (j < K) {
      float s = 0.f;
      for (int sp = 0; sp < nsplit; ++sp) s += partials[sp * split_stride + static_cast<int64_t>(j) * E + c];
      const int n = counts[j];
      cn = n > 0 ? s / static_cast<float>(n) : 0.f;
      const float d = cn - centroids[static_cast<int64_t>(j) * E + c];
      e2 += d * d;
      n2 += cn * cn;
      centroids[static_cast<int64_t>(j) * E + c] = cn;
    }
    const __half hi = __float2half_rn(cn);
    c_hi[static_cast<int64_t>(j) * E + c] = hi;
    c_lo[static_cast<int64_t>(j) * E + c] = __float2half_rn(cn - __half2float(hi));
  }
  e2 = warp_sum(e2);
  n2 = warp_sum(n2);
  if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = e2; red[1][threadIdx.x >> 5] = n2; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float a = 0.f, b = 0.f;
    for (int w = 0; w < 8; ++w) { a += red[0][w]; b += red[1][w]; }
    err_per_cluster[j] = a;
    cnorm[j] = b;
    __threadfence();
    last = atomicAdd(done, 1u) == static_cast<unsigned int>(Kp) - 1u;
  }
  __syncthreads();
  if (last && threadIdx.x == 0) {                              // fixed summation order: the stop test is reproducible
    __threadfence();
    float tot = 0.f;
    for (int q = 0; q < K; ++q) tot += reinterpret_cast<volatile float*>(err_per_cluster)[q];
    *error = tot;
    *done = 0u;
  }
}

// centroids -> hi / lo halves and norms (initialisation from given centroids)
__global__ void __launch_bounds__(256)
kmeans_prepare_kernel(const float* __restrict__ centroids, __half* __restrict__ c_hi, __half* __restrict__ c_lo,
                      float* __restrict__ cnorm, int K, int E) {
  pdl_sync();
  __shared__ float red[8];
  const int j = blockIdx.x;
  float n2 = 0.f;
  for (int c = threadIdx.x; c < E; c += 256) {
    const float cn = j < K ? centroids[static_cast<int64_t>(j) * E + c] : 0.f;
    n2 += cn * cn;
    const __half hi = __float2half_rn(cn);
    c_hi[static_cast<int64_t>(j) * E + c] = hi;
    c_lo[static_cast<int64_t>(j) * E + c] = __float2half_rn(cn - __half2float(hi));
  }
  n2 = warp_sum(n2);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = n2;
  __syncthreads();
  if (threadIdx.x == 0) {
    float b = 0.f;
    for (int w = 0; w < 8; ++w) b += red[w];
    cnorm[j] = b;
  }
}

// X fp16 [Np, E] -> Xt fp16 [E, Np_pad] (columns >= Np zero): the K-major operand of the centroid-sum GEMM
__global__ void __launch_bounds__(256)
transpose_f16_kernel(const __half* __restrict__ X, int Np, int E, __half* __restrict__ Xt, int Np_pad) {
  pdl_sync();
  __shared__ __half tile[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;      // 32 x 8
  const int p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  for (int r = ty; r < 32; r += 8) {
    const int p = p0 + r, c = c0 + tx;
    tile[r][tx] = (p < Np && c < E) ? X[static_cast<int64_t>(p) * E + c] : __float2half_rn(0.f);
  }
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {
    const int c = c0 + r, p = p0 + tx;
    if (c < E && p < Np_pad) Xt[static_cast<int64_t>(c) * Np_pad + p] = tile[tx][r];
  }
}

}  // namespace
}  // namespace cffm

using namespace cffm;

extern "C" int cffm_kmeans_prepare(const float* centroids, void* c_hi, void* c_lo, float* cnorm, int K, int Kp, int E,
                                   void* stream) {
  CFFM_REQUIRE(centroids && c_hi && c_lo && cnorm, CFFM_E_BADARG, "kmeans_prepare: null pointer");
  CFFM_REQUIRE(K > 0 && Kp >= K && E > 0, CFFM_E_BADARG, "kmeans_prepare: bad size");
  launch_k(kmeans_prepare_kernel, Kp, 256, 0, static_cast<cudaStream_t>(stream), centroids, static_cast<__half*>(c_hi),
           static_cast<__half*>(c_lo), cnorm, K, E);
  return launch_status("kmeans_prepare_kernel");
}

extern "C" int cffm_kmeans_assign(const float* scores, int64_t lds, const float* cnorm, int Np, int Np_pad, int K, int Kp,
                                  int64_t* labels, void* onehot, int* counts, void* stream) {
  CFFM_REQUIRE(scores && cnorm && labels && onehot && counts, CFFM_E_BADARG, "kmeans_assign: null pointer");
  CFFM_REQUIRE(Np > 0 && Np_pad >= Np && K > 0 && K <= 256 && Kp >= K && lds >= K, CFFM_E_BADARG,
               "kmeans_assign: bad size (Np=%d Np_pad=%d K=%d Kp=%d); at most 256 clusters", Np, Np_pad, K, Kp);
  launch_k(kmeans_assign_kernel, (Np_pad + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream), scores, lds, cnorm, Np, Np_pad,
           K, Kp, labels, static_cast<__half*>(onehot), counts);
  return launch_status("kmeans_assign_kernel");
}

extern "C" int cffm_kmeans_update(const float* partials, int nsplit, const int* counts, float* centroids, void* c_hi,
                                  void* c_lo, float* cnorm, float* err_per_cluster, float* error, unsigned int* done,
                                  int K, int Kp, int E, void* stream) {
  CFFM_REQUIRE(partials && counts && centroids && c_hi && c_lo && cnorm && err_per_cluster && error && done, CFFM_E_BADARG,
               "kmeans_update: null pointer");
  CFFM_REQUIRE(nsplit >= 1 && K > 0 && Kp >= K && E > 0, CFFM_E_BADARG, "kmeans_update: bad size");
  launch_k(kmeans_update_kernel, Kp, 256, 0, static_cast<cudaStream_t>(stream), partials, nsplit, static_cast<int64_t>(Kp) * E,
           counts, centroids, static_cast<__half*>(c_hi), static_cast<__half*>(c_lo), cnorm, err_per_cluster, error, done, K, Kp, E);
  return launch_status("kmeans_update_kernel");
}

extern "C" int cffm_transpose_f16(const void* x, int rows, int cols, void* xt, int rows_pad, void* stream) {
  CFFM_REQUIRE(x && xt, CFFM_E_BADARG, "transpose: null pointer");
  CFFM_REQUIRE(rows > 0 && cols > 0 && rows_pad >= rows, CFFM_E_BADARG, "transpose: bad size");
  dim3 grid((rows_pad + 31) / 32, (cols + 31) / 32);
  launch_k(transpose_f16_kernel, grid, 256, 0, static_cast<cudaStream_t>(stream), static_cast<const __half*>(x), rows, cols,
           static_cast<__half*>(xt), rows_pad);
  return launch_status("transpose_f16_kernel");
}
