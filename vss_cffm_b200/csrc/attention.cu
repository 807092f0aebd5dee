// softmax(scale q k^T) v with the whole key set of one (batch, head) resident in shared memory, on warp-level
// mma.sync m16n8k16: the fallback of cffm_mha_f16 for the shapes the tcgen05 kernel (mha_sm100.cu) does not take.
// Flash-style online softmax over 64-key chunks with fp32 statistics and fp16 operands.
// (The CFM attention lives in cfm_sm100.cu.)
#include "common.cuh"
#include "ptx_sm100.cuh"

namespace cffm {
namespace {

constexpr float LOG2E = 1.4426950408889634f;

// ------------------------------------------------------------------------------------------------
// generic small-KV multi-head attention
template <int D>
__global__ void __launch_bounds__(256)
mha_small_kv_kernel(const __half* __restrict__ q, int64_t ldq, const __half* __restrict__ k,
                    const __half* __restrict__ v, int64_t ldkv, __half* __restrict__ out, int64_t ldo, int Nq, int Nkv,
                    int nkv_pad, float scale_log2e) {
  pdl_sync();
  constexpr int LD = D + 8;                                 // padded smem row (halves): conflict-free LDS/ldmatrix
  extern __shared__ __align__(16) uint8_t smem_raw[];
  __half* Ks = reinterpret_cast<__half*>(smem_raw);
  __half* Vs = Ks + static_cast<size_t>(nkv_pad) * LD;
  const int b = blockIdx.z, h = blockIdx.y;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t = lane & 3;

  // ---- stage K and V of this (batch, head) in shared memory
  constexpr int CH = D / 8;                                 // 16-byte chunks per row
  const __half* kb = k + (static_cast<int64_t>(b) * Nkv) * ldkv + h * D;
  const __half* vb = v + (static_cast<int64_t>(b) * Nkv) * ldkv + h * D;
  for (int i = tid; i < nkv_pad * CH; i += 256) {
    const int row = i / CH, ch = i % CH;
    __half* dk = Ks + row * LD + ch * 8;
    __half* dv = Vs + row * LD + ch * 8;
    if (row < Nkv) {
      ptx::cp_async16(dk, kb + static_cast<int64_t>(row) * ldkv + ch * 8);
      ptx::cp_async16(dv, vb + static_cast<int64_t>(row) * ldkv + ch * 8);
    } else {
      *reinterpret_cast<uint4*>(dk) = make_uint4(0, 0, 0, 0);
      *reinterpret_cast<uint4*>(dv) = make_uint4(0, 0, 0, 0);
    }
  }
  ptx::cp_async_commit();

  ptx::cp_async_wait_all();
  __syncthreads();
  const uint32_t vs_addr = ptx::smem_u32(Vs);
  // ---- loop over 128-query tiles: K/V of this (batch, head) stay resident in shared memory
  for (int qt = blockIdx.x; qt * 128 < Nq; qt += gridDim.x) {
  // Q fragments straight from global (each warp owns 16 query rows)
  const int r0 = qt * 128 + warp * 16;
  if (r0 >= Nq) continue;                                   // warp-uniform; no block syncs inside the loop
  const int rowA = r0 + g, rowB = r0 + g + 8;
  const __half* qA = q + (static_cast<int64_t>(b) * Nq + rowA) * ldq + h * D;
  const __half* qB = q + (static_cast<int64_t>(b) * Nq + rowB) * ldq + h * D;
  uint32_t a[D / 16][4];
#pragma unroll
  for (int kk = 0; kk < D / 16; ++kk) {
    a[kk][0] = rowA < Nq ? *reinterpret_cast<const uint32_t*>(qA + kk * 16 + 2 * t) : 0u;
    a[kk][1] = rowB < Nq ? *reinterpret_cast<const uint32_t*>(qB + kk * 16 + 2 * t) : 0u;
    a[kk][2] = rowA < Nq ? *reinterpret_cast<const uint32_t*>(qA + kk * 16 + 8 + 2 * t) : 0u;
    a[kk][3] = rowB < Nq ? *reinterpret_cast<const uint32_t*>(qB + kk * 16 + 8 + 2 * t) : 0u;
  }
  float o[D / 8][4];
#pragma unroll
  for (int i = 0; i < D / 8; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
  float mA = -INFINITY, mB = -INFINITY, lA = 0.f, lB = 0.f;

  for (int c = 0; c < nkv_pad; c += 64) {
    float s[8][4];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f;
#pragma unroll
    for (int kk = 0; kk < D / 16; ++kk) {
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        const __half* kr = Ks + (c + nt * 8 + g) * LD + kk * 16 + 2 * t;
        ptx::mma_m16n8k16(s[nt], a[kk], *reinterpret_cast<const uint32_t*>(kr),
                          *reinterpret_cast<const uint32_t*>(kr + 8));
      }
    }
    float cmA = -INFINITY, cmB = -INFINITY;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      const int key = c + nt * 8 + 2 * t;
      s[nt][0] = key < Nkv ? s[nt][0] * scale_log2e : -INFINITY;
      s[nt][1] = key + 1 < Nkv ? s[nt][1] * scale_log2e : -INFINITY;
      s[nt][2] = key < Nkv ? s[nt][2] * scale_log2e : -INFINITY;
      s[nt][3] = key + 1 < Nkv ? s[nt][3] * scale_log2e : -INFINITY;
      cmA = fmaxf(cmA, fmaxf(s[nt][0], s[nt][1]));
      cmB = fmaxf(cmB, fmaxf(s[nt][2], s[nt][3]));
    }
    cmA = fmaxf(cmA, __shfl_xor_sync(0xffffffffu, cmA, 1));
    cmA = fmaxf(cmA, __shfl_xor_sync(0xffffffffu, cmA, 2));
    cmB = fmaxf(cmB, __shfl_xor_sync(0xffffffffu, cmB, 1));
    cmB = fmaxf(cmB, __shfl_xor_sync(0xffffffffu, cmB, 2));
    const float nmA = fmaxf(mA, cmA), nmB = fmaxf(mB, cmB);
    const float alA = exp2f(mA - nmA), alB = exp2f(mB - nmB);
    mA = nmA; mB = nmB;
    float psA = 0.f, psB = 0.f;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      s[nt][0] = exp2f(s[nt][0] - mA); s[nt][1] = exp2f(s[nt][1] - mA);
      s[nt][2] = exp2f(s[nt][2] - mB); s[nt][3] = exp2f(s[nt][3] - mB);
      psA += s[nt][0] + s[nt][1];
      psB += s[nt][2] + s[nt][3];
    }
    lA = lA * alA + psA;
    lB = lB * alB + psB;
#pragma unroll
    for (int nd = 0; nd < D / 8; ++nd) { o[nd][0] *= alA; o[nd][1] *= alA; o[nd][2] *= alB; o[nd][3] *= alB; }
#pragma unroll
    for (int j = 0; j < 4; ++j) {                           // 16 keys per step
      uint32_t pa[4];
      pa[0] = pack_half2(s[2 * j][0], s[2 * j][1]);
      pa[1] = pack_half2(s[2 * j][2], s[2 * j][3]);
      pa[2] = pack_half2(s[2 * j + 1][0], s[2 * j + 1][1]);
      pa[3] = pack_half2(s[2 * j + 1][2], s[2 * j + 1][3]);
      const int krow = c + j * 16 + ((lane >> 3) & 1) * 8 + (lane & 7);
#pragma unroll
      for (int nd = 0; nd < D / 8; nd += 2) {
        uint32_t b0, b1, b2, b3;
        ptx::ldmatrix_x4_trans(b0, b1, b2, b3, vs_addr + (krow * LD + (nd + (lane >> 4)) * 8) * 2);
        ptx::mma_m16n8k16(o[nd], pa, b0, b1);
        ptx::mma_m16n8k16(o[nd + 1], pa, b2, b3);
      }
    }
  }
  lA += __shfl_xor_sync(0xffffffffu, lA, 1); lA += __shfl_xor_sync(0xffffffffu, lA, 2);
  lB += __shfl_xor_sync(0xffffffffu, lB, 1); lB += __shfl_xor_sync(0xffffffffu, lB, 2);
  const float iA = 1.f / lA, iB = 1.f / lB;
  __half* oA = out + (static_cast<int64_t>(b) * Nq + rowA) * ldo + h * D;
  __half* oB = out + (static_cast<int64_t>(b) * Nq + rowB) * ldo + h * D;
#pragma unroll
  for (int nd = 0; nd < D / 8; ++nd) {
    if (rowA < Nq) *reinterpret_cast<uint32_t*>(oA + nd * 8 + 2 * t) = pack_half2(o[nd][0] * iA, o[nd][1] * iA);
    if (rowB < Nq) *reinterpret_cast<uint32_t*>(oB + nd * 8 + 2 * t) = pack_half2(o[nd][2] * iB, o[nd][3] * iB);
  }
  }                                                         // query-tile loop
}

constexpr int SMEM_CAP = 200 * 1024;
template <auto Kernel>
int set_smem_attr() { return set_dyn_smem<Kernel>(SMEM_CAP, "mha"); }

}  // namespace
}  // namespace cffm

namespace cffm {
int mha_tcgen05_launch(const void* q, int64_t ldq, const void* k, const void* v, int64_t ldkv, void* out, int64_t ldo,
                       int batch, int Nq, int Nkv, int heads, int head_dim, float scale, cudaStream_t st, long long* prof = nullptr);
static bool mha_legacy() {
  static const bool on = [] { const char* e = getenv("CFFM_MHA_LEGACY"); return e && e[0] == '1'; }();
  return on;
}
}  // namespace cffm

extern "C" int cffm_mha_f16(const void* q, int64_t ldq, const void* k, const void* v, int64_t ldkv, void* out,
                            int64_t ldo, int batch, int Nq, int Nkv, int heads, int head_dim, float scale,
                            void* stream) {
  using namespace cffm;
  CFFM_REQUIRE(q && k && v && out, CFFM_E_BADARG, "mha: null pointer");
  CFFM_REQUIRE(batch > 0 && Nq > 0 && Nkv > 0 && heads > 0, CFFM_E_BADARG, "mha: non-positive size");
  CFFM_REQUIRE(head_dim == 32 || head_dim == 64, CFFM_E_UNSUPPORTED, "mha: head_dim %d not in {32,64}", head_dim);
  CFFM_REQUIRE(ldq % 8 == 0 && ldkv % 8 == 0 && ldo % 2 == 0 && aligned16(k) && aligned16(v) &&
                   (reinterpret_cast<uintptr_t>(q) & 3) == 0 && (reinterpret_cast<uintptr_t>(out) & 3) == 0,
               CFFM_E_BADARG, "mha: misaligned pointer or stride");
  CFFM_REQUIRE(batch <= 65535 && heads <= 65535, CFFM_E_UNSUPPORTED, "mha: batch/heads exceed grid limits");
  const int nkv_pad = (Nkv + 63) / 64 * 64;
  const int smem = 2 * nkv_pad * (head_dim + 8) * 2;
  CFFM_REQUIRE(smem <= 200 * 1024, CFFM_E_UNSUPPORTED, "mha: Nkv=%d does not fit in shared memory", Nkv);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (!mha_legacy()) {                                         // tcgen05 path (mha_sm100.cu); falls through when unsupported
    const int rc5 = mha_tcgen05_launch(q, ldq, k, v, ldkv, out, ldo, batch, Nq, Nkv, heads, head_dim, scale, st);
    if (rc5 != CFFM_E_UNSUPPORTED) return rc5;
  }
  // enough CTAs for ~2 per SM; each loops over query tiles so that the K/V staging is amortised
  const int qtiles = (Nq + 127) / 128;
  int gx = (2 * 148 + heads * batch - 1) / (heads * batch);
  if (gx > qtiles) gx = qtiles;
  if (gx < 1) gx = 1;
  dim3 grid(gx, heads, batch);
  const float sl = scale * 1.4426950408889634f;
  int rc;
  if (head_dim == 64) {
    if ((rc = set_smem_attr<mha_small_kv_kernel<64>>())) return rc;
    launch_k(mha_small_kv_kernel<64>, grid, 256, smem, st, static_cast<const __half*>(q), ldq, static_cast<const __half*>(k),
                                                     static_cast<const __half*>(v), ldkv, static_cast<__half*>(out),
                                                     ldo, Nq, Nkv, nkv_pad, sl);
  } else {
    if ((rc = set_smem_attr<mha_small_kv_kernel<32>>())) return rc;
    launch_k(mha_small_kv_kernel<32>, grid, 256, smem, st, static_cast<const __half*>(q), ldq, static_cast<const __half*>(k),
                                                     static_cast<const __half*>(v), ldkv, static_cast<__half*>(out),
                                                     ldo, Nq, Nkv, nkv_pad, sl);
  }
  return launch_status("mha_small_kv_kernel");
}

