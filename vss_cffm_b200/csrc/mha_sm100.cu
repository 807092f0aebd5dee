// softmax(scale q k^T) v on the 5th-generation tensor cores, for head_dim 64 and <= 256 keys:
// the MiT efficient attention (mix_transformer.py:96-117: N_kv = 225 at every stage for 480x480 input).
//
// Persistent CTAs loop over work items (frame b, head h, 128-query tile).  Per item:
//   TMA      Q tile [128 x 64], K [256 x 64], V [256 x 64] (fp16, 128-byte swizzle) -> shared memory
//   tcgen05  S[128 x 256] = Q K^T   (A = Q, B = K, both K-major; fp32 accumulator in TMEM columns 0..255)
//   softmax  16 warps: thread = (row, 64-column quarter); two passes over S straight out of TMEM (max, then
//            exp2 / sum / fp16 pack); P is written to shared memory in the K-major swizzled A-operand layout,
//            one 64-key chunk at a time, and handed to the MMA warp chunk by chunk
//   tcgen05  O[128 x 64] += P_chunk V_chunk  (A = P K-major, B = V MN-major: V is used as loaded, no transpose;
//            accumulator aliased onto the first 64 columns of the item's own S buffer)
// S is double-buffered in TMEM (2 x 256 columns): Q K^T of item i+1 is issued before P V of item i.
//   epilogue O / rowsum -> fp16 -> global
// Keys >= N_kv (rows of the next frame, or TMA zero fill past the end) are masked to -inf before the softmax.
// Warp roles: 0..15 softmax / epilogue (4 per scheduler: the exp2 chains of one warp hide behind the others), 16 TMA
// producer, 17 TMEM allocator + MMA issuer.
#include <cuda.h>

#include "common.cuh"
#include "ptx_sm100.cuh"

namespace cffm {
namespace {

constexpr int QT = 128, KV_MAX = 256, D = 64;
constexpr int Q_BYTES = QT * D * 2, KV_BYTES = KV_MAX * D * 2, P_CHUNK_BYTES = QT * 64 * 2;
constexpr int SMEM_MHA = Q_BYTES + 3 * KV_BYTES + 4 * P_CHUNK_BYTES + 2 * 4 * QT * 4 /*row max / sum exchange*/ +
                         256 /*barriers*/ + 1024 /*align slack*/;
constexpr int SM_WARPS = 16;                      // softmax / epilogue warps: 4 TMEM lane quarters x 4 column quarters
constexpr int MHA_THREADS = (SM_WARPS + 2) * 32;
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
  uint8_t* sV = sK + KV_BYTES;                                 // 2 buffers
  uint8_t* sP = sV + 2 * KV_BYTES;                             // 4 chunks of [128 rows][64 keys]
  float* xmax = reinterpret_cast<float*>(sP + 4 * P_CHUNK_BYTES);   // [2 halves][128 rows]
  float* xsum = xmax + 4 * QT;
  uint64_t* bars = reinterpret_cast<uint64_t*>(xsum + 2 * QT);
  uint64_t* qk_full = bars + 0;    // TMA -> MMA: Q and K landed
  uint64_t* qk_empty = bars + 1;   // MMA -> TMA: Q, K consumed
  uint64_t* v_full = bars + 2;     // [2] TMA -> MMA: V landed
  uint64_t* v_empty = bars + 4;    // [2] MMA -> TMA: V consumed (all PV MMAs of the item retired)
  uint64_t* s_full = bars + 6;     // [2] MMA -> softmax: S complete (per TMEM buffer)
  uint64_t* p_full = bars + 8;     // [4] softmax -> MMA: P chunk written (4 warps each)
  uint64_t* o_full = bars + 12;    // [2] MMA -> epilogue: O complete
  uint64_t* o_empty = bars + 14;   // [2] epilogue -> MMA: O read out (8 warps): the TMEM buffer may be overwritten
  uint32_t* tmem_base_smem = reinterpret_cast<uint32_t*>(bars + 16);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int qtiles = (Nq + QT - 1) / QT;
  const int n_items = batch * heads * qtiles;

  if (warp == SM_WARPS && lane == 0) {
    ptx::prefetch_tensormap(&tmQ);
    ptx::prefetch_tensormap(&tmK);
    ptx::prefetch_tensormap(&tmV);
    ptx::mbar_init(qk_full, 1); ptx::mbar_init(qk_empty, 1);
    for (int b = 0; b < 2; ++b) {
      ptx::mbar_init(&v_full[b], 1); ptx::mbar_init(&v_empty[b], 1); ptx::mbar_init(&s_full[b], 1);
      ptx::mbar_init(&o_full[b], 1); ptx::mbar_init(&o_empty[b], SM_WARPS);
    }
    for (int c = 0; c < 4; ++c) ptx::mbar_init(&p_full[c], 4);
    ptx::fence_barrier_init();
  }
  if (warp == SM_WARPS + 1) {
    ptx::tmem_alloc(tmem_base_smem, 512);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_base_smem;
  pdl_sync();                                                  // prologue above overlaps the previous kernel

  // TMEM: two S buffers of 256 columns.  O of an item is accumulated into columns [0, 64) of ITS OWN S buffer: those
  // columns are dead once P chunk 0 has been produced, which is exactly when the first PV MMA is issued.  The MMA
  // warp issues Q K^T of item i+1 (other buffer) before P V of item i, so the tensor pipe works on the next tile
  // while the softmax warps are busy with the current one.
  // Producer and MMA roles: the whole warp walks the loop and waits on the barriers, one ELECTED lane issues the TMA /
  // tcgen05 instructions (with `if (lane == 0)` ptxas wraps every UTMALDG / UTCHMMA / UTCBAR in an ELECT + R2UR.BROADCAST
  // + BRA.U.ANY loop, ~15 dependent instructions each).
  if (warp == SM_WARPS) {
    // ===================== TMA producer =====================
    uint32_t it = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
      const int qt = item % qtiles, bh = item / qtiles, h = bh % heads, b = bh / heads;
      const uint32_t buf = it & 1u, bph = (it >> 1) & 1u;
      ptx::mbar_wait(qk_empty, (it & 1u) ^ 1u);
      if (ptx::elect_one()) {
        ptx::mbar_arrive_expect_tx(qk_full, Q_BYTES + KV_BYTES);
        ptx::tma_load_2d(sQ, &tmQ, qk_full, h * D, b * Nq + qt * QT);
        ptx::tma_load_2d(sK, &tmK, qk_full, h * D, b * Nkv);
      }
      __syncwarp();
      ptx::mbar_wait(&v_empty[buf], bph ^ 1u);
      if (ptx::elect_one()) {
        ptx::mbar_arrive_expect_tx(&v_full[buf], KV_BYTES);
        ptx::tma_load_2d(sV + buf * KV_BYTES, &tmV, &v_full[buf], h * D, b * Nkv);
      }
      __syncwarp();
    }
  } else if (warp == SM_WARPS + 1) {
    // ===================== MMA issuer =====================
    constexpr uint32_t idesc_qk = ptx::make_idesc_f16(QT, KV_MAX);                    // A, B K-major
    constexpr uint32_t idesc_pv = ptx::make_idesc_f16(QT, D) | (1u << 16);            // B (= V) MN-major
    const uint64_t dq = desc_sw128(ptx::smem_u32(sQ)), dk = desc_sw128(ptx::smem_u32(sK));
    auto issue_qk = [&](uint32_t j) {                          // S[j & 1] = Q K^T of the CTA's j-th item
      ptx::mbar_wait(qk_full, j & 1u);
      if (j >= 2) ptx::mbar_wait(&o_empty[j & 1u], ((j - 2) >> 1) & 1u);   // O of item j-2 (same buffer) was read
      ptx::tc_fence_after();
      const uint32_t tmem_s = tmem_base + (j & 1u) * KV_MAX;
      if (ptx::elect_one()) {
#pragma unroll
        for (int k = 0; k < D / 16; ++k) ptx::umma_f16(tmem_s, dq + 2u * k, dk + 2u * k, idesc_qk, k != 0 ? 1u : 0u);
        ptx::umma_commit(&s_full[j & 1u]);
        ptx::umma_commit(qk_empty);
      }
      __syncwarp();
    };
    const uint32_t my_items = blockIdx.x < static_cast<uint32_t>(n_items)
                                  ? (n_items - blockIdx.x + gridDim.x - 1) / gridDim.x : 0u;
    if (my_items > 0) issue_qk(0);
    for (uint32_t it = 0; it < my_items; ++it) {
      const uint32_t buf = it & 1u, bph = (it >> 1) & 1u;
      if (it + 1 < my_items) issue_qk(it + 1);                 // next tile's scores while this tile's softmax runs
      ptx::mbar_wait(&v_full[buf], bph);
      const uint32_t tmem_o = tmem_base + buf * KV_MAX;
      for (int c = 0; c < 4; ++c) {
        ptx::mbar_wait(&p_full[c], it & 1u);
        ptx::tc_fence_after();
        const uint64_t dp = desc_sw128(ptx::smem_u32(sP + c * P_CHUNK_BYTES));
        // 16 keys per MMA: A advances 32 bytes inside the swizzled row, B (MN-major V) advances 16 key rows
        const uint64_t dv = desc_sw128(ptx::smem_u32(sV + buf * KV_BYTES + c * 64 * (D * 2)));
        if (ptx::elect_one()) {
#pragma unroll
          for (int k = 0; k < 4; ++k)
            ptx::umma_f16(tmem_o, dp + 2u * k, dv + static_cast<uint64_t>(k * (16 * D * 2 / 16)), idesc_pv, (c | k) != 0 ? 1u : 0u);
          if (c == 3) {
            ptx::umma_commit(&o_full[buf]);
            ptx::umma_commit(&v_empty[buf]);
          }
        }
        __syncwarp();
      }
    }
  } else {
    // ===================== softmax + epilogue: thread = (row, column half) =====================
    const int wq = warp & 3, cf = warp >> 2;                   // TMEM lane quarter, 64-column quarter (= P chunk)
    const int row = wq * 32 + lane;                            // row of the query tile = TMEM lane
    const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(wq * 32) << 16);
    const int bar_id = 1 + wq;                                 // the four warps that share the same 32 rows

    // pass 1: scaled row maximum of the item whose scores sit in TMEM buffer `buf`
    auto row_max = [&](uint32_t buf) -> float {
      float m = -INFINITY;
#pragma unroll 1
      for (int sc = 0; sc < 2; ++sc) {
        const int c0 = cf * 64 + sc * 32;
        if (c0 >= Nkv) break;                                  // warp-uniform
        uint32_t v[32];
        ptx::tmem_ld_32x32b_x32(lane_addr + buf * KV_MAX + c0, v);
        ptx::tmem_ld_wait();
        if (c0 + 32 <= Nkv) {
          float m0 = __uint_as_float(v[0]), m1 = __uint_as_float(v[1]), m2 = __uint_as_float(v[2]), m3 = __uint_as_float(v[3]);
#pragma unroll
          for (int j = 4; j < 32; j += 4) {
            m0 = fmaxf(m0, __uint_as_float(v[j])); m1 = fmaxf(m1, __uint_as_float(v[j + 1]));
            m2 = fmaxf(m2, __uint_as_float(v[j + 2])); m3 = fmaxf(m3, __uint_as_float(v[j + 3]));
          }
          m = fmaxf(m, fmaxf(fmaxf(m0, m1), fmaxf(m2, m3)));
        } else {                                               // only the chunk that straddles N_kv pays predicates
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (c0 + j < Nkv) m = fmaxf(m, __uint_as_float(v[j]));
        }
      }
      xmax[cf * QT + row] = m;
      asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
      const float mo = fmaxf(fmaxf(xmax[row], xmax[QT + row]), fmaxf(xmax[2 * QT + row], xmax[3 * QT + row]));
      asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");     // xmax may be overwritten by the next call
      return mo * scale_log2e;                                 // positive scale; N_kv >= 1: finite
    };

    uint32_t it = 0;
    float m = 0.f;
    if (static_cast<int>(blockIdx.x) < n_items) {
      ptx::mbar_wait(&s_full[0], 0);
      ptx::tc_fence_after();
      m = row_max(0);
    }
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
      const int qt = item % qtiles, bh = item / qtiles, h = bh % heads, b = bh / heads;
      const uint32_t buf = it & 1u, bph = (it >> 1) & 1u;
      const uint32_t s_addr = lane_addr + buf * KV_MAX;
      // ---- pass 2: p = exp2(s * scale - m), row sum, fp16 P chunks in the swizzled A-operand layout
      float sum0 = 0.f, sum1 = 0.f;
      {
        const int c = cf;                                      // this warp's 64-key chunk
        uint8_t* prow = sP + c * P_CHUNK_BYTES + (row >> 3) * 1024 + (row & 7) * 128;
#pragma unroll 1
        for (int sc = 0; sc < 2; ++sc) {
          const int c0 = c * 64 + sc * 32;
          uint32_t hh[16];
          if (c0 >= Nkv) {                                     // warp-uniform: nothing but zeros
#pragma unroll
            for (int j = 0; j < 16; ++j) hh[j] = 0u;
          } else {
            uint32_t v[32];
            ptx::tmem_ld_32x32b_x32(s_addr + c0, v);
            ptx::tmem_ld_wait();
            if (c0 + 32 <= Nkv) {
#pragma unroll
              for (int j = 0; j < 16; ++j) {
                const float p0 = ptx::ex2_approx(fmaf(__uint_as_float(v[2 * j]), scale_log2e, -m));
                const float p1 = ptx::ex2_approx(fmaf(__uint_as_float(v[2 * j + 1]), scale_log2e, -m));
                sum0 += p0; sum1 += p1;
                hh[j] = pack_half2(p0, p1);
              }
            } else {
#pragma unroll
              for (int j = 0; j < 16; ++j) {
                const float p0 = c0 + 2 * j < Nkv ? ptx::ex2_approx(fmaf(__uint_as_float(v[2 * j]), scale_log2e, -m)) : 0.f;
                const float p1 = c0 + 2 * j + 1 < Nkv ? ptx::ex2_approx(fmaf(__uint_as_float(v[2 * j + 1]), scale_log2e, -m)) : 0.f;
                sum0 += p0; sum1 += p1;
                hh[j] = pack_half2(p0, p1);
              }
            }
          }
#pragma unroll
          for (int g = 0; g < 4; ++g)                          // 8 keys = one 16-byte piece
            *reinterpret_cast<uint4*>(prow + (((sc * 4 + g) ^ (row & 7)) << 4)) =
                make_uint4(hh[4 * g], hh[4 * g + 1], hh[4 * g + 2], hh[4 * g + 3]);
        }
        ptx::tc_fence_before();                                // TMEM reads of this chunk precede the MMA that overwrites it
        ptx::fence_proxy_async();                              // generic-proxy smem writes -> visible to the MMA (async proxy)
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&p_full[c]);
      }
      xsum[cf * QT + row] = sum0 + sum1;
      asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
      const float inv = 1.f / ((xsum[row] + xsum[QT + row]) + (xsum[2 * QT + row] + xsum[3 * QT + row]));   // same order in all 4 threads
      // ---- pass 1 of the NEXT item (its Q K^T was issued before this item's P V): hides the P V tail
      float m_next = 0.f;
      if (item + static_cast<int>(gridDim.x) < n_items) {
        ptx::mbar_wait(&s_full[buf ^ 1u], ((it + 1) >> 1) & 1u);
        ptx::tc_fence_after();
        m_next = row_max(buf ^ 1u);
      }
      // ---- epilogue: O[row, 16 cf .. +16) / rowsum -> fp16 -> global
      ptx::mbar_wait(&o_full[buf], bph);
      ptx::tc_fence_after();
      uint32_t o[16];
      ptx::tmem_ld_32x32b_x16(s_addr + cf * 16, o);
      ptx::tmem_ld_wait();
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&o_empty[buf]);
      const int qrow = qt * QT + row;
      if (qrow < Nq) {
        __half* dst = out + (static_cast<int64_t>(b) * Nq + qrow) * ldo + h * D + cf * 16;
#pragma unroll
        for (int g = 0; g < 2; ++g) {
          uint4 w;
          w.x = pack_half2(__uint_as_float(o[g * 8]) * inv, __uint_as_float(o[g * 8 + 1]) * inv);
          w.y = pack_half2(__uint_as_float(o[g * 8 + 2]) * inv, __uint_as_float(o[g * 8 + 3]) * inv);
          w.z = pack_half2(__uint_as_float(o[g * 8 + 4]) * inv, __uint_as_float(o[g * 8 + 5]) * inv);
          w.w = pack_half2(__uint_as_float(o[g * 8 + 6]) * inv, __uint_as_float(o[g * 8 + 7]) * inv);
          *reinterpret_cast<uint4*>(dst + g * 8) = w;
        }
      }
      m = m_next;
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == SM_WARPS + 1) ptx::tmem_dealloc(tmem_base, 512);
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
