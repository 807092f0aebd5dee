// softmax(scale q k^T) v on the 5th-generation tensor cores, for head_dim 64 and <= 256 keys:
// the MiT efficient attention (mix_transformer.py:96-117: N_kv = 225 at every stage for 480x480 input).
//
// Persistent CTAs loop over work items (frame b, head h, 128-query tile).  Per item:
//   TMA      Q tile [128 x 64], K [256 x 64], V [256 x 64] (fp16, 128-byte swizzle) -> shared memory
//   tcgen05  S[128 x 256] = Q K^T   (A = Q, B = K, both K-major; fp32 accumulator in TMEM columns 0..255)
//   softmax  8 warps: thread = (row, 128-column half); two passes over S straight out of TMEM (max, then
//            exp2 / sum / fp16 pack); P is written to shared memory in the K-major swizzled A-operand layout,
//            one 64-key chunk at a time, and handed to the MMA warp chunk by chunk
//   tcgen05  O[128 x 64] += P_chunk V_chunk  (A = P K-major, B = V MN-major: V is used as loaded, no transpose;
//            accumulator in TMEM columns 256..319)
//   epilogue O / rowsum -> fp16 -> global
// Keys >= N_kv (rows of the next frame, or TMA zero fill past the end) are masked to -inf before the softmax.
// Warp roles: 0..7 softmax / epilogue, 8 TMA producer, 9 TMEM allocator + MMA issuer.
#include <cuda.h>

#include "common.cuh"
#include "ptx_sm100.cuh"

namespace cffm {
namespace {

constexpr int QT = 128, KV_MAX = 256, D = 64;
constexpr int Q_BYTES = QT * D * 2, KV_BYTES = KV_MAX * D * 2, P_CHUNK_BYTES = QT * 64 * 2;
constexpr int SMEM_MHA = Q_BYTES + 2 * KV_BYTES + 4 * P_CHUNK_BYTES + 2 * 2 * QT * 4 /*row max / sum exchange*/ +
                         256 /*barriers*/ + 1024 /*align slack*/;
constexpr int MHA_THREADS = 320;
constexpr float LOG2E = 1.4426950408889634f;

// K-major / MN-major 128-byte-swizzled operand descriptor: 8-row (or 8-key) groups of 1024 bytes
__device__ __forceinline__ uint64_t desc_sw128(uint32_t smem_addr) { return ptx::make_smem_desc_sw128(smem_addr); }

__global__ void __launch_bounds__(MHA_THREADS, 1)
mha_tcgen05_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                   const __grid_constant__ CUtensorMap tmV, __half* __restrict__ out, int64_t ldo, int batch, int Nq,
                   int Nkv, int heads, float scale_log2e) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = ptx::smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + Q_BYTES;
  uint8_t* sV = sK + KV_BYTES;
  uint8_t* sP = sV + KV_BYTES;                                 // 4 chunks of [128 rows][64 keys]
  float* xmax = reinterpret_cast<float*>(sP + 4 * P_CHUNK_BYTES);   // [2 halves][128 rows]
  float* xsum = xmax + 2 * QT;
  uint64_t* bars = reinterpret_cast<uint64_t*>(xsum + 2 * QT);
  uint64_t* qk_full = bars + 0;    // TMA -> MMA: Q and K landed
  uint64_t* v_full = bars + 1;     // TMA -> MMA: V landed
  uint64_t* qk_empty = bars + 2;   // MMA -> TMA: Q, K consumed
  uint64_t* v_empty = bars + 3;    // MMA -> TMA: V consumed (all PV MMAs retired)
  uint64_t* s_full = bars + 4;     // MMA -> softmax: S complete
  uint64_t* s_empty = bars + 5;    // softmax -> MMA: S read out (8 warps)
  uint64_t* p_full = bars + 6;     // [4] softmax -> MMA: P chunk written (4 warps each)
  uint64_t* o_full = bars + 10;    // MMA -> epilogue: O complete
  uint64_t* o_empty = bars + 11;   // epilogue -> MMA: O read out (8 warps)
  uint32_t* tmem_base_smem = reinterpret_cast<uint32_t*>(bars + 12);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int qtiles = (Nq + QT - 1) / QT;
  const int n_items = batch * heads * qtiles;

  if (warp == 8 && lane == 0) {
    ptx::prefetch_tensormap(&tmQ);
    ptx::prefetch_tensormap(&tmK);
    ptx::prefetch_tensormap(&tmV);
    ptx::mbar_init(qk_full, 1); ptx::mbar_init(v_full, 1); ptx::mbar_init(qk_empty, 1); ptx::mbar_init(v_empty, 1);
    ptx::mbar_init(s_full, 1); ptx::mbar_init(s_empty, 8);
    for (int c = 0; c < 4; ++c) ptx::mbar_init(&p_full[c], 4);
    ptx::mbar_init(o_full, 1); ptx::mbar_init(o_empty, 8);
    ptx::fence_barrier_init();
  }
  if (warp == 9) {
    ptx::tmem_alloc(tmem_base_smem, 512);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_base_smem;
  pdl_sync();                                                  // prologue above overlaps the previous kernel

  if (warp == 8) {
    if (lane == 0) {
      // ===================== TMA producer =====================
      uint32_t it = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
        const int qt = item % qtiles, bh = item / qtiles, h = bh % heads, b = bh / heads;
        const uint32_t ph = it & 1u;
        ptx::mbar_wait(qk_empty, ph ^ 1u);
        ptx::mbar_arrive_expect_tx(qk_full, Q_BYTES + KV_BYTES);
        ptx::tma_load_2d(sQ, &tmQ, qk_full, h * D, b * Nq + qt * QT);
        ptx::tma_load_2d(sK, &tmK, qk_full, h * D, b * Nkv);
        ptx::mbar_wait(v_empty, ph ^ 1u);
        ptx::mbar_arrive_expect_tx(v_full, KV_BYTES);
        ptx::tma_load_2d(sV, &tmV, v_full, h * D, b * Nkv);
      }
    }
  } else if (warp == 9) {
    if (lane == 0) {
      // ===================== MMA issuer (one thread) =====================
      constexpr uint32_t idesc_qk = ptx::make_idesc_f16(QT, KV_MAX);                    // A, B K-major
      constexpr uint32_t idesc_pv = ptx::make_idesc_f16(QT, D) | (1u << 16);            // B (= V) MN-major
      const uint32_t tmem_s = tmem_base, tmem_o = tmem_base + KV_MAX;
      uint32_t it = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
        const uint32_t ph = it & 1u;
        ptx::mbar_wait(qk_full, ph);
        ptx::mbar_wait(s_empty, ph ^ 1u);                      // previous item's softmax has drained S
        ptx::tc_fence_after();
        const uint64_t dq = desc_sw128(ptx::smem_u32(sQ)), dk = desc_sw128(ptx::smem_u32(sK));
#pragma unroll
        for (int k = 0; k < D / 16; ++k) ptx::umma_f16(tmem_s, dq + 2u * k, dk + 2u * k, idesc_qk, k != 0 ? 1u : 0u);
        ptx::umma_commit(s_full);
        ptx::umma_commit(qk_empty);
        ptx::mbar_wait(v_full, ph);
        ptx::mbar_wait(o_empty, ph ^ 1u);                      // previous item's O has been read
        for (int c = 0; c < 4; ++c) {
          ptx::mbar_wait(&p_full[c], ph);
          ptx::tc_fence_after();
          const uint64_t dp = desc_sw128(ptx::smem_u32(sP + c * P_CHUNK_BYTES));
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            // 16 keys per MMA: A advances 32 bytes inside the swizzled row, B (MN-major V) advances 16 key rows
            const uint64_t dv = desc_sw128(ptx::smem_u32(sV + (c * 64 + k * 16) * (D * 2)));
            ptx::umma_f16(tmem_o, dp + 2u * k, dv, idesc_pv, (c | k) != 0 ? 1u : 0u);
          }
        }
        ptx::umma_commit(o_full);
        ptx::umma_commit(v_empty);
      }
    }
  } else {
    // ===================== softmax + epilogue: thread = (row, column half) =====================
    const int wq = warp & 3, hf = warp >> 2;                   // TMEM lane quarter, column half
    const int row = wq * 32 + lane;                            // row of the query tile = TMEM lane
    const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(wq * 32) << 16);
    const int bar_id = 1 + wq;                                 // pairs warp w with warp w + 4 (same rows)
    uint32_t it = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
      const int qt = item % qtiles, bh = item / qtiles, h = bh % heads, b = bh / heads;
      const uint32_t ph = it & 1u;
      ptx::mbar_wait(s_full, ph);
      ptx::tc_fence_after();
      // ---- pass 1: row maximum over this thread's 128 columns
      float m = -INFINITY;
#pragma unroll 1
      for (int sc = 0; sc < 4; ++sc) {
        const int c0 = hf * 128 + sc * 32;
        if (c0 >= Nkv) break;                                  // warp-uniform
        uint32_t v[32];
        ptx::tmem_ld_32x32b_x32(lane_addr + c0, v);
        ptx::tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (c0 + j < Nkv) m = fmaxf(m, __uint_as_float(v[j]));
      }
      xmax[hf * QT + row] = m;
      asm volatile("bar.sync %0, 64;" ::"r"(bar_id) : "memory");
      m = fmaxf(m, xmax[(hf ^ 1) * QT + row]) * scale_log2e;   // Nkv >= 1: finite
      // ---- pass 2: p = exp2(s * scale - m), row sum, fp16 P chunks in the swizzled A-operand layout
      float sum = 0.f;
#pragma unroll 1
      for (int j2 = 0; j2 < 2; ++j2) {
        const int c = hf * 2 + j2;                             // 64-key chunk
        uint8_t* prow = sP + c * P_CHUNK_BYTES + (row >> 3) * 1024 + (row & 7) * 128;
#pragma unroll 1
        for (int sc = 0; sc < 2; ++sc) {
          const int c0 = c * 64 + sc * 32;
          uint32_t v[32];
          if (c0 < Nkv) {                                      // warp-uniform
            ptx::tmem_ld_32x32b_x32(lane_addr + c0, v);
            ptx::tmem_ld_wait();
          }
#pragma unroll
          for (int g = 0; g < 4; ++g) {                        // 8 keys = one 16-byte piece
            float p[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              const int col = c0 + g * 8 + e;
              p[e] = col < Nkv ? exp2f(fmaf(__uint_as_float(v[g * 8 + e]), scale_log2e, -m)) : 0.f;
              sum += p[e];
            }
            const int piece = sc * 4 + g;
            *reinterpret_cast<uint4*>(prow + ((piece ^ (row & 7)) << 4)) =
                make_uint4(pack_half2(p[0], p[1]), pack_half2(p[2], p[3]), pack_half2(p[4], p[5]), pack_half2(p[6], p[7]));
          }
        }
        ptx::fence_proxy_async();                              // generic-proxy smem writes -> visible to the MMA (async proxy)
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&p_full[c]);
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(s_empty);                // S fully read by this warp
      xsum[hf * QT + row] = sum;
      asm volatile("bar.sync %0, 64;" ::"r"(bar_id) : "memory");
      const float inv = 1.f / (sum + xsum[(hf ^ 1) * QT + row]);
      // ---- epilogue: O[row, 32 hf .. +32) / rowsum -> fp16 -> global
      ptx::mbar_wait(o_full, ph);
      ptx::tc_fence_after();
      uint32_t o[32];
      ptx::tmem_ld_32x32b_x32(lane_addr + KV_MAX + hf * 32, o);
      ptx::tmem_ld_wait();
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(o_empty);
      const int qrow = qt * QT + row;
      if (qrow < Nq) {
        __half* dst = out + (static_cast<int64_t>(b) * Nq + qrow) * ldo + h * D + hf * 32;
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          uint4 w;
          w.x = pack_half2(__uint_as_float(o[g * 8]) * inv, __uint_as_float(o[g * 8 + 1]) * inv);
          w.y = pack_half2(__uint_as_float(o[g * 8 + 2]) * inv, __uint_as_float(o[g * 8 + 3]) * inv);
          w.z = pack_half2(__uint_as_float(o[g * 8 + 4]) * inv, __uint_as_float(o[g * 8 + 5]) * inv);
          w.w = pack_half2(__uint_as_float(o[g * 8 + 6]) * inv, __uint_as_float(o[g * 8 + 7]) * inv);
          *reinterpret_cast<uint4*>(dst + g * 8) = w;
        }
      }
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 9) ptx::tmem_dealloc(tmem_base, 512);
}

}  // namespace

// Returns CFFM_E_UNSUPPORTED (without setting an error) when the shape is outside this kernel: the caller falls
// back to the mma.sync kernel of attention.cu.
int mha_tcgen05_launch(const void* q, int64_t ldq, const void* k, const void* v, int64_t ldkv, void* out, int64_t ldo,
                       int batch, int Nq, int Nkv, int heads, float scale, cudaStream_t st) {
  if (Nkv > KV_MAX || ldq % 8 != 0 || ldkv % 8 != 0 || ldo % 8 != 0 || !aligned16(q) || !aligned16(k) || !aligned16(v) ||
      !aligned16(out))
    return CFFM_E_UNSUPPORTED;
  CUtensorMap tmQ, tmK, tmV;
  int rc = make_tmap(&tmQ, q, static_cast<int64_t>(batch) * Nq, static_cast<int64_t>(heads) * D, ldq, QT);
  if (rc) return rc;
  if ((rc = make_tmap(&tmK, k, static_cast<int64_t>(batch) * Nkv, static_cast<int64_t>(heads) * D, ldkv, KV_MAX))) return rc;
  if ((rc = make_tmap(&tmV, v, static_cast<int64_t>(batch) * Nkv, static_cast<int64_t>(heads) * D, ldkv, KV_MAX))) return rc;
  static cudaError_t attr_err =
      cudaFuncSetAttribute(mha_tcgen05_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_MHA);
  CFFM_REQUIRE(attr_err == cudaSuccess, -(int)attr_err, "cudaFuncSetAttribute: %s", cudaGetErrorString(attr_err));
  const int items = batch * heads * ((Nq + QT - 1) / QT);
  const int grid = items < num_sms() ? items : num_sms();
  launch_k(mha_tcgen05_kernel, grid, MHA_THREADS, SMEM_MHA, st, tmQ, tmK, tmV, static_cast<__half*>(out), ldo, batch, Nq,
           Nkv, heads, scale * LOG2E);
  return launch_status("mha_tcgen05_kernel");
}

}  // namespace cffm
