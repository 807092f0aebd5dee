// Cross-frame Feature Mining attention (WindowAttention3d3.forward, cffm_transformer.py:364-601) on the 5th-generation
// tensor cores: TMA-assembled K/V, tcgen05 Q K^T and P V with TMEM accumulators, warp-specialised.
//
// Work item = (clip b, 7x7 window (wi, wj), head PAIR hp): 2 x 49 queries against the window's assembled key sequence.
//
// K/V assembling (cffm_transformer.py:378-522) is five TMA box loads per operand -- nothing is gathered by threads:
//   * the target-frame K/V map is kept with a 3-wide cyclic APRON ([B, Hp+6, Wp+6, 768], written by the norm1 kernel and
//     the QKV GEMM), so "own window + the four torch.roll'ed neighbour strips" (:389-418) is ONE 13 x 13 box: the halo
//     of the window.  The 12 ring keys the reference lists twice (valid_ind_rolled, :280-285) appear once; their two
//     bias entries are folded on the host into logaddexp(b1, b2), which is the same softmax (277 unique keys, not 289).
//   * the four pooled levels (target fc-pool, three reference frames) are nn.Unfold windows (:298-301, :339-343) of
//     small maps: 5x5 / 7x7 / 5x5 / 3x3 boxes at (i-2, j-2) / (i-3, j-3) / (2i-2, 2j-2) / (3i-1, 3j-1); the TMA
//     out-of-bounds zero fill IS Unfold's zero padding, and the -100 mask (:433-446, :481-492) follows from the same
//     coordinates (K = V = 0 there, so exp(-100) relative weight is replaced by an exact 0: difference < 4e-44).
// Every key row is 128 bytes = the 64 channels of the head pair, 128-byte swizzled, i.e. directly a tcgen05 operand.
//
// Per item:
//   S[128 x N] = Qbd K^T   M = 128 rows = (head 0: 49 queries | pad to 64 | head 1: 49 queries | pad), K = 64 channels with
//                          Q block-diagonal (head-0 rows carry zeros in head 1's channels and vice versa), so ONE
//                          M = 128 MMA computes both heads at the tensor-core cost of two M = 64 ones.
//   softmax                8 warps, thread = (row, half of each 64-key chunk); logits t = S + bias / scale straight out of
//                          TMEM (one FHFMA per element: fp16 bias operand, fp32 accumulator), exact two-pass row
//                          maximum, p = 2^((t - m) scale log2 e) -> fp16 P chunks in shared memory (swizzled A operand)
//   O[128 x 64] += P V     B = V chunk MN-major as loaded; only the diagonal 32-column blocks are read back
//   epilogue               O / rowsum -> fp16, window_reverse + crop (:812-821) fused into the store
// Warp roles: 0..7 softmax / epilogue, 8 producer (TMA + block-diagonal Q through cp.async), 9 TMEM allocator + MMA issuer.
// K, Q and V are single-buffered: K/Q of item i+1 are loaded while the softmax of item i runs, V after P V of item i.
#include <cuda.h>

#include "common.cuh"
#include "ptx_sm100.cuh"

namespace cffm {
namespace {

constexpr int WS = 7, RING = 3, HALO = WS + 2 * RING, NHALO = HALO * HALO;      // 13, 169
constexpr int CQKV = 768, CKV = 512, CPAIR = 64;                                // channels: qkv row, pooled K|V row, head pair
constexpr int ROWB = 128;                                                       // bytes per key / query row in shared memory
constexpr int SM_WARPS = 8, CFM_THREADS = (SM_WARPS + 2) * 32;
constexpr int P_RING = 2, P_CHUNK_BYTES = 128 * ROWB;
constexpr int Q_BYTES = 128 * ROWB;
constexpr float LOG2E = 1.4426950408889634f;

// Key-row layout of the assembled sequence.  TIGHT packs the five boxes back to back (TMA destinations are 128-byte
// aligned; the 128-byte swizzle is a function of the shared-memory address, so a box may start inside a 1024-byte
// swizzle atom).  The ALIGNED variant starts every box on a 1024-byte boundary (rows padded with masked zero keys).
template <bool ALIGNED>
struct Lay {
  static constexpr int R1 = ALIGNED ? 176 : NHALO;           // pooled target level, 5 x 5
  static constexpr int R2 = R1 + (ALIGNED ? 32 : 25);        // reference frame 0, 7 x 7
  static constexpr int R3 = R2 + (ALIGNED ? 56 : 49);        // reference frame 1, 5 x 5
  static constexpr int R4 = R3 + (ALIGNED ? 32 : 25);        // reference frame 2, 3 x 3
  static constexpr int REND = R4 + 9;
  static constexpr int NPAD = ALIGNED ? 320 : 288;           // MMA N (two halves, each a multiple of 16)
  static constexpr int BPITCH = ALIGNED ? 328 : 296;         // bias row pitch (halves): conflict-free 16-byte LDS per row
  static constexpr int NCH = (NPAD + 63) / 64;               // 64-key P chunks (the last one of TIGHT holds 32 keys)
  static constexpr int KV_BYTES = NPAD * ROWB;
  static constexpr int TX_BYTES = (NHALO + 25 + 49 + 25 + 9) * ROWB;
  static constexpr int BIAS_BYTES = 2 * 49 * BPITCH * 2;
  static constexpr int SMEM = 2 * KV_BYTES + Q_BYTES + P_RING * P_CHUNK_BYTES + BIAS_BYTES + 2 * NPAD * 4 /*mask*/ +
                              2 * 2 * 2 * 128 * 4 /*row max / sum exchange*/ + 256 /*barriers*/ + 1024 /*align slack*/;
  static constexpr int TMEM_O = NPAD;                        // two O buffers of 64 columns behind S
};

struct CfmParams {
  const __half* qkv_a;      // [B, Hp+6, Wp+6, 768] apron layout
  const __half* bias_tab;   // [8 heads, 49, BPITCH] fp16, bias / scale in the kernel's key order, -inf on unused columns
  __half* out;              // [B, H, W, 256]
  __half* dump;             // test hook: [items of head pair 0.., 2, NPAD, 64] assembled K and V tiles (or null)
  int B, H, W, nWh, nWw;
  float scale_log2e;
};

// Additive mask (in accumulator units) of key row n for window (wi, wj): 0, or -inf for an nn.Unfold zero-padding
// position of a pooled level.  Halo rows and unused rows are 0 (unused rows carry a -inf bias).
template <bool AL>
__device__ __forceinline__ float key_mask(int n, int wi, int wj, int nWh, int nWw) {
  using L = Lay<AL>;
  int m, kc, st, f;
  if (n >= L::R1 && n < L::R1 + 25) { m = n - L::R1; kc = 5; st = 1; f = 1; }
  else if (n >= L::R2 && n < L::R2 + 49) { m = n - L::R2; kc = 7; st = 1; f = 1; }
  else if (n >= L::R3 && n < L::R3 + 25) { m = n - L::R3; kc = 5; st = 2; f = 2; }
  else if (n >= L::R4 && n < L::R4 + 9) { m = n - L::R4; kc = 3; st = 3; f = 3; }
  else return 0.f;
  const int y = st * wi + m / kc - kc / 2, x = st * wj + m % kc - kc / 2;
  return (y >= 0 && y < f * nWh && x >= 0 && x < f * nWw) ? 0.f : -INFINITY;
}

// ---- softmax / epilogue role, G = which half of every 64-key chunk this warp owns (compile time: all column offsets
// are immediates)
template <bool AL, int G>
__device__ __forceinline__ void softmax_role(const CfmParams& p, uint8_t* sP, const __half* sBias, float* sMask, float* sX,
                                             uint64_t* s_full, uint64_t* s_empty, uint64_t* p_full, uint64_t* p_empty,
                                             uint64_t* o_full, uint64_t* o_empty, uint32_t tmem_base, int hp, int slot,
                                             int nslots) {
  using L = Lay<AL>;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int wq = warp & 3;                                   // TMEM lane quarter
  const int row = wq * 32 + lane;                            // row of S = TMEM lane
  const int hl = row >> 6, q = row & 63;                     // head of the pair, query index (>= 49: padding)
  const int qc = q < 49 ? q : 48;
  const __half* brow = sBias + (hl * 49 + qc) * L::BPITCH;
  const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(wq * 32) << 16);
  const int stid = threadIdx.x;                              // 0..255 among the softmax warps
  const int nW = p.nWh * p.nWw, n_items = p.B * nW;
  const float scl = p.scale_log2e;
  const int bar_pair = 1 + wq;                               // named barrier of the two warps that share these 32 rows

  uint32_t it = 0;
  for (int item = slot; item < n_items; item += nslots, ++it) {
    const int b = item / nW, w = item - b * nW, wi = w / p.nWw, wj = w - wi * p.nWw;
    float* mask = sMask + (it & 1u) * L::NPAD;
    for (int n = stid; n < L::NPAD; n += SM_WARPS * 32) mask[n] = key_mask<AL>(n, wi, wj, p.nWh, p.nWw);
    asm volatile("bar.sync 5, 256;" ::: "memory");

    ptx::mbar_wait(s_full, it & 1u);
    ptx::tc_fence_after();

    // t[j] = S[row, col0 + j] + bias / scale (+ mask): the logit divided by the (positive) attention scale
    auto logits = [&](int col0, int wcol, float* t) {
      uint32_t v[32];
      if (wcol == 32) ptx::tmem_ld_32x32b_x32(lane_addr + col0, v);
      else ptx::tmem_ld_32x32b_x16(lane_addr + col0, v);
      const uint4* bp = reinterpret_cast<const uint4*>(brow + col0);
      uint4 bw[4];
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (j * 8 < wcol) bw[j] = bp[j];
      ptx::tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (j * 8 < wcol) {
          const uint32_t ww[4] = {bw[j].x, bw[j].y, bw[j].z, bw[j].w};
#pragma unroll
          for (int e = 0; e < 4; ++e)
            ptx::fhfma2(ww[e], __uint_as_float(v[j * 8 + 2 * e]), __uint_as_float(v[j * 8 + 2 * e + 1]), t[j * 8 + 2 * e],
                        t[j * 8 + 2 * e + 1]);
        }
      }
      if (col0 + wcol > L::R1) {                             // pooled levels: nn.Unfold padding positions are masked
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          if (j * 4 < wcol) {
            const float4 mk = *reinterpret_cast<const float4*>(mask + col0 + j * 4);
            t[j * 4] += mk.x; t[j * 4 + 1] += mk.y; t[j * 4 + 2] += mk.z; t[j * 4 + 3] += mk.w;
          }
        }
      }
    };

    // ---- pass 1: exact row maximum
    float mx = -INFINITY;
#pragma unroll
    for (int c = 0; c < L::NCH; ++c) {
      const int wcol = (c * 64 + 64 <= L::NPAD) ? 32 : 16;
      const int col0 = c * 64 + G * wcol;
      float t[32];
      logits(col0, wcol, t);
      float m0 = t[0], m1 = t[1], m2 = t[2], m3 = t[3];
#pragma unroll
      for (int j = 4; j < 32; j += 4) {
        if (j < wcol) { m0 = fmaxf(m0, t[j]); m1 = fmaxf(m1, t[j + 1]); m2 = fmaxf(m2, t[j + 2]); m3 = fmaxf(m3, t[j + 3]); }
      }
      mx = fmaxf(mx, fmaxf(fmaxf(m0, m1), fmaxf(m2, m3)));
    }
    float* xm = sX + (it & 1u) * 512;                        // [max | sum][G][128 rows]
    xm[G * 128 + row] = mx;
    asm volatile("bar.sync %0, 64;" ::"r"(bar_pair) : "memory");
    mx = fmaxf(mx, xm[(G ^ 1) * 128 + row]);                 // the halo keys are never masked: finite
    const float msc = mx * scl;

    // ---- pass 2: p = 2^((t - m) scale log2 e), row sum, fp16 P chunks in the swizzled A-operand layout
    float sum = 0.f;
#pragma unroll
    for (int c = 0; c < L::NCH; ++c) {
      const int wcol = (c * 64 + 64 <= L::NPAD) ? 32 : 16;
      const int col0 = c * 64 + G * wcol;
      const uint32_t gc = it * L::NCH + c, ps = gc % P_RING;
      float t[32];
      logits(col0, wcol, t);
      if (c == L::NCH - 1) {                                 // S of this item has been read for the last time
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(s_empty);
      }
      uint32_t hh[16];
      float s0 = 0.f, s1 = 0.f;
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        if (2 * j < wcol) {
          const float p0 = ptx::ex2_approx(fmaf(t[2 * j], scl, -msc));
          const float p1 = ptx::ex2_approx(fmaf(t[2 * j + 1], scl, -msc));
          s0 += p0; s1 += p1;
          hh[j] = pack_half2(p0, p1);
        }
      }
      sum += s0 + s1;
      ptx::mbar_wait(&p_empty[ps], ((gc / P_RING) & 1u) ^ 1u);
      uint8_t* prow = sP + ps * P_CHUNK_BYTES + row * ROWB;
#pragma unroll
      for (int g = 0; g < 4; ++g) {                          // 8 keys = one 16-byte piece
        if (g * 8 < wcol) {
          const int piece = G * (wcol / 8) + g;
          *reinterpret_cast<uint4*>(prow + ((piece ^ (row & 7)) << 4)) =
              make_uint4(hh[4 * g], hh[4 * g + 1], hh[4 * g + 2], hh[4 * g + 3]);
        }
      }
      ptx::fence_proxy_async();                              // generic-proxy smem writes -> visible to the MMA (async proxy)
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&p_full[ps]);
    }
    xm[256 + G * 128 + row] = sum;
    asm volatile("bar.sync %0, 64;" ::"r"(bar_pair) : "memory");
    const float inv = 1.f / (xm[256 + row] + xm[256 + 128 + row]);    // same order in both threads of the row

    // ---- epilogue: O[row, 32 hl + 16 G .. +16) / rowsum -> fp16 -> out (window_reverse + crop)
    const uint32_t ob = it & 1u;
    ptx::mbar_wait(&o_full[ob], (it >> 1) & 1u);
    ptx::tc_fence_after();
    uint32_t o[16];
    ptx::tmem_ld_32x32b_x16(lane_addr + L::TMEM_O + ob * 64 + hl * 32 + G * 16, o);
    ptx::tmem_ld_wait();
    ptx::tc_fence_before();
    __syncwarp();
    if (lane == 0) ptx::mbar_arrive(&o_empty[ob]);
    const int y = WS * wi + q / WS, x = WS * wj + q % WS;
    if (q < 49 && y < p.H && x < p.W) {
      __half* dst = p.out + ((static_cast<int64_t>(b) * p.H + y) * p.W + x) * 256 + hp * CPAIR + hl * 32 + G * 16;
#pragma unroll
      for (int g = 0; g < 2; ++g) {
        uint4 wv;
        wv.x = pack_half2(__uint_as_float(o[g * 8]) * inv, __uint_as_float(o[g * 8 + 1]) * inv);
        wv.y = pack_half2(__uint_as_float(o[g * 8 + 2]) * inv, __uint_as_float(o[g * 8 + 3]) * inv);
        wv.z = pack_half2(__uint_as_float(o[g * 8 + 4]) * inv, __uint_as_float(o[g * 8 + 5]) * inv);
        wv.w = pack_half2(__uint_as_float(o[g * 8 + 6]) * inv, __uint_as_float(o[g * 8 + 7]) * inv);
        *reinterpret_cast<uint4*>(dst + g * 8) = wv;
      }
    }
  }
}

template <bool AL>
__global__ void __launch_bounds__(CFM_THREADS, 1)
cfm_attention_tc_kernel(const __grid_constant__ CUtensorMap tmT, const __grid_constant__ CUtensorMap tmL1,
                        const __grid_constant__ CUtensorMap tmL2, const __grid_constant__ CUtensorMap tmL3,
                        const __grid_constant__ CUtensorMap tmL4, const CfmParams p) {
  using L = Lay<AL>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = ptx::smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);
  uint8_t* sK = smem;
  uint8_t* sV = sK + L::KV_BYTES;
  uint8_t* sQ = sV + L::KV_BYTES;
  uint8_t* sP = sQ + Q_BYTES;
  __half* sBias = reinterpret_cast<__half*>(sP + P_RING * P_CHUNK_BYTES);
  float* sMask = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(sBias) + L::BIAS_BYTES);
  float* sX = sMask + 2 * L::NPAD;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sX + 2 * 512);
  uint64_t* k_full = bars + 0;     // TMA -> MMA: K tile landed (transaction bytes)
  uint64_t* q_full = bars + 1;     // producer warp -> MMA: block-diagonal Q written
  uint64_t* kq_empty = bars + 2;   // MMA -> producer: Q K^T retired, K and Q may be overwritten
  uint64_t* v_full = bars + 3;     // TMA -> MMA: V tile landed
  uint64_t* v_empty = bars + 4;    // MMA -> producer: P V retired
  uint64_t* s_full = bars + 5;     // MMA -> softmax: S complete
  uint64_t* s_empty = bars + 6;    // softmax (8 warps) -> MMA: S read for the last time
  uint64_t* p_full = bars + 7;     // [P_RING] softmax (8 warps) -> MMA: P chunk written
  uint64_t* p_empty = bars + 9;    // [P_RING] MMA -> softmax: P chunk consumed
  uint64_t* o_full = bars + 11;    // [2] MMA -> epilogue: O complete
  uint64_t* o_empty = bars + 13;   // [2] epilogue (8 warps) -> MMA: O read out
  uint32_t* tmem_base_smem = reinterpret_cast<uint32_t*>(bars + 16);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int hp = blockIdx.x & 3, slot = blockIdx.x >> 2, nslots = gridDim.x >> 2;
  const int nW = p.nWh * p.nWw, n_items = p.B * nW;

  if (warp == SM_WARPS && lane == 0) {
    ptx::prefetch_tensormap(&tmT); ptx::prefetch_tensormap(&tmL1); ptx::prefetch_tensormap(&tmL2);
    ptx::prefetch_tensormap(&tmL3); ptx::prefetch_tensormap(&tmL4);
    ptx::mbar_init(k_full, 1); ptx::mbar_init(q_full, 1); ptx::mbar_init(kq_empty, 1);
    ptx::mbar_init(v_full, 1); ptx::mbar_init(v_empty, 1);
    ptx::mbar_init(s_full, 1); ptx::mbar_init(s_empty, SM_WARPS);
    for (int i = 0; i < P_RING; ++i) { ptx::mbar_init(&p_full[i], SM_WARPS); ptx::mbar_init(&p_empty[i], 1); }
    for (int i = 0; i < 2; ++i) { ptx::mbar_init(&o_full[i], 1); ptx::mbar_init(&o_empty[i], SM_WARPS); }
    ptx::fence_barrier_init();
  }
  if (warp == SM_WARPS + 1) {
    ptx::tmem_alloc(tmem_base_smem, 512);
    ptx::tmem_relinquish();
  }
  // K, V, Q tiles start as zeros: rows the TMA boxes never touch (unused key rows, query rows >= 49, the other head's
  // channels of the block-diagonal Q) must stay finite -- they meet P = 0 or are never read back.
  for (int i = threadIdx.x; i < (2 * L::KV_BYTES + Q_BYTES) / 16; i += CFM_THREADS)
    reinterpret_cast<uint4*>(sK)[i] = make_uint4(0u, 0u, 0u, 0u);
  if (warp < SM_WARPS) {
    // the head pair's bias slice: a constant table (not produced by the previous kernel), so it is fetched before the
    // programmatic-dependent-launch wait and overlaps the tail of the QKV GEMM
    const uint4* src = reinterpret_cast<const uint4*>(p.bias_tab + static_cast<int64_t>(hp) * 2 * 49 * L::BPITCH);
    for (int i = threadIdx.x; i < L::BIAS_BYTES / 16; i += SM_WARPS * 32)
      ptx::cp_async16(reinterpret_cast<uint4*>(sBias) + i, src + i);
    ptx::cp_async_commit();
  }
  ptx::fence_proxy_async();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_base_smem;
  pdl_sync();

  if (warp == SM_WARPS) {
    // ===================== producer: TMA K / V boxes, block-diagonal Q =====================
    const int Wa = p.nWw * WS + 2 * RING, Ha = p.nWh * WS + 2 * RING;
    uint32_t it = 0;
    for (int item = slot; item < n_items; item += nslots, ++it) {
      const int b = item / nW, w = item - b * nW, wi = w / p.nWw, wj = w - wi * p.nWw;
      if (it > 0) ptx::mbar_wait(kq_empty, (it - 1) & 1u);
      auto boxes = [&](uint8_t* dst, uint64_t* bar, int ct, int cp) {      // ct / cp: channel of K (or V) in qkv / pooled rows
        ptx::mbar_arrive_expect_tx(bar, L::TX_BYTES);
        ptx::tma_load_4d(dst, &tmT, bar, ct, WS * wj, WS * wi, b);
        ptx::tma_load_4d(dst + L::R1 * ROWB, &tmL1, bar, cp, wj - 2, wi - 2, b);
        ptx::tma_load_4d(dst + L::R2 * ROWB, &tmL2, bar, cp, wj - 3, wi - 3, b);
        ptx::tma_load_4d(dst + L::R3 * ROWB, &tmL3, bar, cp, 2 * wj - 2, 2 * wi - 2, b);
        ptx::tma_load_4d(dst + L::R4 * ROWB, &tmL4, bar, cp, 3 * wj - 1, 3 * wi - 1, b);
      };
      if (lane == 0) boxes(sK, k_full, 256 + hp * CPAIR, hp * CPAIR);
      // Q: rows 0..48 = head 0 (bytes 0..63 of the row), rows 64..112 = head 1 (bytes 64..127); 16-byte pieces
      for (int i = lane; i < 2 * 49 * 4; i += 32) {
        const int pc = i & 3, rq = i >> 2, hl = rq >= 49 ? 1 : 0, q = rq - hl * 49;
        const int r = hl * 64 + q;
        const __half* src = p.qkv_a + ((static_cast<int64_t>(b) * Ha + WS * wi + RING + q / WS) * Wa + WS * wj + RING + q % WS) * CQKV +
                            hp * CPAIR + hl * 32 + pc * 8;
        ptx::cp_async16(sQ + r * ROWB + (((hl * 4 + pc) ^ (r & 7)) << 4), src);
      }
      ptx::cp_async_commit();
      ptx::cp_async_wait_all();
      ptx::fence_proxy_async();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(q_full);
      if (p.dump != nullptr) {                               // test hook: the assembled K tile, de-swizzled
        ptx::mbar_wait(k_full, it & 1u);
        __half* d = p.dump + (static_cast<int64_t>(hp) * n_items + item) * 2 * L::NPAD * CPAIR;
        for (int i = lane; i < L::NPAD * 8; i += 32) {
          const int n = i >> 3, pc = i & 7;
          *reinterpret_cast<uint4*>(d + n * CPAIR + pc * 8) = *reinterpret_cast<const uint4*>(sK + n * ROWB + ((pc ^ (n & 7)) << 4));
        }
      }
      if (it > 0) ptx::mbar_wait(v_empty, (it - 1) & 1u);
      if (lane == 0) boxes(sV, v_full, 512 + hp * CPAIR, 256 + hp * CPAIR);
      if (p.dump != nullptr) {
        ptx::mbar_wait(v_full, it & 1u);
        __half* d = p.dump + ((static_cast<int64_t>(hp) * n_items + item) * 2 + 1) * L::NPAD * CPAIR;
        for (int i = lane; i < L::NPAD * 8; i += 32) {
          const int n = i >> 3, pc = i & 7;
          *reinterpret_cast<uint4*>(d + n * CPAIR + pc * 8) = *reinterpret_cast<const uint4*>(sV + n * ROWB + ((pc ^ (n & 7)) << 4));
        }
      }
    }
  } else if (warp == SM_WARPS + 1) {
    if (lane == 0) {
      // ===================== MMA issuer (one thread) =====================
      constexpr int NH2 = L::NPAD / 2;
      constexpr uint32_t idesc_qk = ptx::make_idesc_f16(128, NH2);                       // A, B K-major
      constexpr uint32_t idesc_pv = ptx::make_idesc_f16(128, CPAIR) | (1u << 16);         // B (= V) MN-major
      const uint64_t dq = ptx::make_smem_desc_sw128(ptx::smem_u32(sQ));
      const uint64_t dk = ptx::make_smem_desc_sw128(ptx::smem_u32(sK));
      const uint32_t my_items = slot < n_items ? (n_items - slot + nslots - 1) / nslots : 0u;
      for (uint32_t it = 0; it < my_items; ++it) {
        ptx::mbar_wait(k_full, it & 1u);
        ptx::mbar_wait(q_full, it & 1u);
        if (it > 0) ptx::mbar_wait(s_empty, (it - 1) & 1u);
        ptx::tc_fence_after();
#pragma unroll
        for (int half = 0; half < 2; ++half)
#pragma unroll
          for (int k = 0; k < 4; ++k)
            ptx::umma_f16(tmem_base + half * NH2, dq + 2u * k, dk + static_cast<uint64_t>(half * NH2 * (ROWB / 16)) + 2u * k,
                          idesc_qk, k != 0 ? 1u : 0u);
        ptx::umma_commit(s_full);
        ptx::umma_commit(kq_empty);
        const uint32_t ob = it & 1u;
        const uint32_t tmem_o = tmem_base + L::TMEM_O + ob * 64;
#pragma unroll 1
        for (int c = 0; c < L::NCH; ++c) {
          const uint32_t gc = it * L::NCH + c, ps = gc % P_RING;
          ptx::mbar_wait(&p_full[ps], (gc / P_RING) & 1u);
          if (c == 0) {
            ptx::mbar_wait(v_full, it & 1u);
            if (it >= 2) ptx::mbar_wait(&o_empty[ob], ((it - 2) >> 1) & 1u);
          }
          ptx::tc_fence_after();
          const uint64_t dp = ptx::make_smem_desc_sw128(ptx::smem_u32(sP + ps * P_CHUNK_BYTES));
          const int ksteps = (c * 64 + 64 <= L::NPAD) ? 4 : 2;
          for (int k = 0; k < ksteps; ++k) {
            // 16 keys per MMA: A advances 32 bytes inside the swizzled row, B (MN-major V) advances 16 key rows
            const uint64_t dv = ptx::make_smem_desc_sw128(ptx::smem_u32(sV + (c * 64 + k * 16) * ROWB));
            ptx::umma_f16(tmem_o, dp + 2u * k, dv, idesc_pv, (c | k) != 0 ? 1u : 0u);
          }
          ptx::umma_commit(&p_empty[ps]);
        }
        ptx::umma_commit(&o_full[ob]);
        ptx::umma_commit(v_empty);
      }
    }
  } else {
    ptx::cp_async_wait_all();                                // bias slice (this thread's pieces)
    asm volatile("bar.sync 5, 256;" ::: "memory");           // ... and everybody else's
    if ((warp >> 2) == 0)
      softmax_role<AL, 0>(p, sP, sBias, sMask, sX, s_full, s_empty, p_full, p_empty, o_full, o_empty, tmem_base, hp, slot, nslots);
    else
      softmax_role<AL, 1>(p, sP, sBias, sMask, sX, s_full, s_empty, p_full, p_empty, o_full, o_empty, tmem_base, hp, slot, nslots);
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == SM_WARPS + 1) ptx::tmem_dealloc(tmem_base, 512);
}

bool cfm_aligned_layout() {
  static const bool on = [] { const char* e = getenv("CFFM_CFM_ALIGNED"); return e && e[0] == '1'; }();
  return on;
}

template <bool AL>
int launch_cfm(const CUtensorMap* tm, const CfmParams& p, cudaStream_t st) {
  using L = Lay<AL>;
  static cudaError_t attr_err =
      cudaFuncSetAttribute(cfm_attention_tc_kernel<AL>, cudaFuncAttributeMaxDynamicSharedMemorySize, L::SMEM);
  CFFM_REQUIRE(attr_err == cudaSuccess, -(int)attr_err, "cudaFuncSetAttribute(%d bytes): %s", L::SMEM, cudaGetErrorString(attr_err));
  // persistent CTAs, each bound to one head pair (its bias slice stays in shared memory) and walking over (clip, window) items
  const int n_items = p.B * p.nWh * p.nWw;
  int nslots = num_sms() / 4;
  if (nslots > n_items) nslots = n_items;
  if (nslots < 1) nslots = 1;
  launch_k(cfm_attention_tc_kernel<AL>, nslots * 4, CFM_THREADS, L::SMEM, st, tm[0], tm[1], tm[2], tm[3], tm[4], p);
  return launch_status("cfm_attention_tc_kernel");
}

int cfm_run(const void* qkv_a, const void* kv_pooled, const void* bias_tab, void* out, void* dump, int B, int H, int W, int C,
            int heads, float scale, void* stream) {
  CFFM_REQUIRE(qkv_a && kv_pooled && bias_tab && out, CFFM_E_BADARG, "cfm_attention: null pointer");
  CFFM_REQUIRE(B > 0 && H > 0 && W > 0 && scale > 0.f, CFFM_E_BADARG, "cfm_attention: non-positive size or scale");
  CFFM_REQUIRE(C == 256 && heads == 8, CFFM_E_UNSUPPORTED,
               "cfm_attention: built for C=256, heads=8 (cffm_head.py:74-95), got C=%d heads=%d", C, heads);
  CFFM_REQUIRE(aligned16(qkv_a) && aligned16(kv_pooled) && aligned16(bias_tab) && aligned16(out), CFFM_E_BADARG,
               "cfm_attention: pointers must be 16-byte aligned");
  const int nWh = (H + WS - 1) / WS, nWw = (W + WS - 1) / WS, nW = nWh * nWw;
  const int64_t Ha = nWh * WS + 2 * RING, Wa = nWw * WS + 2 * RING;
  CUtensorMap tm[5];
  {
    const int64_t dims[4] = {CQKV, Wa, Ha, B}, strides[3] = {CQKV, Wa * CQKV, Ha * Wa * CQKV};
    const int box[4] = {CPAIR, HALO, HALO, 1};
    int rc = make_tmap_4d(&tm[0], qkv_a, dims, strides, box);
    if (rc) return rc;
  }
  const int lev_f[4] = {1, 1, 2, 3}, lev_k[4] = {5, 7, 5, 3}, lev_base[4] = {0, 1, 2, 6};
  for (int l = 0; l < 4; ++l) {
    const int64_t gw = static_cast<int64_t>(lev_f[l]) * nWw, gh = static_cast<int64_t>(lev_f[l]) * nWh;
    const int64_t dims[4] = {CKV, gw, gh, B}, strides[3] = {CKV, gw * CKV, static_cast<int64_t>(15) * nW * CKV};
    const int box[4] = {CPAIR, lev_k[l], lev_k[l], 1};
    int rc = make_tmap_4d(&tm[1 + l], static_cast<const __half*>(kv_pooled) + static_cast<int64_t>(lev_base[l]) * nW * CKV, dims,
                          strides, box);
    if (rc) return rc;
  }
  CfmParams p;
  p.qkv_a = static_cast<const __half*>(qkv_a);
  p.bias_tab = static_cast<const __half*>(bias_tab);
  p.out = static_cast<__half*>(out);
  p.dump = static_cast<__half*>(dump);
  p.B = B; p.H = H; p.W = W; p.nWh = nWh; p.nWw = nWw;
  p.scale_log2e = scale * LOG2E;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  return cfm_aligned_layout() ? launch_cfm<true>(tm, p, st) : launch_cfm<false>(tm, p, st);
}

}  // namespace
}  // namespace cffm

extern "C" int cffm_cfm_attention(const void* qkv_a, const void* kv_pooled, const void* bias_tab, void* out, int B, int H,
                                  int W, int C, int heads, float scale, void* stream) {
  return cffm::cfm_run(qkv_a, kv_pooled, bias_tab, out, nullptr, B, H, W, C, heads, scale, stream);
}

extern "C" int cffm_cfm_attention_dump(const void* qkv_a, const void* kv_pooled, const void* bias_tab, void* out, void* dump,
                                       int B, int H, int W, int C, int heads, float scale, void* stream) {
  CFFM_REQUIRE(dump != nullptr, CFFM_E_BADARG, "cfm_attention_dump: null dump buffer");
  return cffm::cfm_run(qkv_a, kv_pooled, bias_tab, out, dump, B, H, W, C, heads, scale, stream);
}

extern "C" int cffm_cfm_layout(int32_t* out8) {
  using namespace cffm;
  CFFM_REQUIRE(out8 != nullptr, CFFM_E_BADARG, "cfm_layout: null pointer");
  const bool al = cfm_aligned_layout();
  out8[0] = 0;
  out8[1] = al ? Lay<true>::R1 : Lay<false>::R1;
  out8[2] = al ? Lay<true>::R2 : Lay<false>::R2;
  out8[3] = al ? Lay<true>::R3 : Lay<false>::R3;
  out8[4] = al ? Lay<true>::R4 : Lay<false>::R4;
  out8[5] = al ? Lay<true>::NPAD : Lay<false>::NPAD;
  out8[6] = al ? Lay<true>::BPITCH : Lay<false>::BPITCH;
  out8[7] = RING;
  return CFFM_OK;
}
