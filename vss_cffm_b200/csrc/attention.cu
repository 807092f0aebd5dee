// Attention kernels of the CFFM path.
//
//  * mha_small_kv_kernel  -- softmax(scale q k^T) v with the whole key set of one (batch, head)
//    resident in shared memory (MiT efficient attention: 225 keys at 480x480; CFFM++ prototypes).
//  * cfm_attention_kernel -- cross-frame feature mining: one CTA per (clip, 7x7 window, head).
//    The 289-key K/V sequence (own window | cyclic ring | pooled target | 3 pooled reference
//    frames) is never materialised in HBM: source coordinates are computed in-kernel and the
//    64-byte head slices are gathered with cp.async straight into shared memory.
//
// Both use flash-style online softmax over 64-key chunks with fp32 statistics, fp16 operands and
// warp-level mma.sync m16n8k16 (the problems are M=49, N=289, K=32: far below a tcgen05 tile).
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

// ------------------------------------------------------------------------------------------------
// CFM key assembling: source of key slot n of window (wi, wj).  Spec: SURVEY.md appendix, proven
// bit-exact against the reference's roll / window_partition / valid_ind_rolled / nn.Unfold / cat
// pipeline (cffm_transformer.py:378-522) by tests/test_oracle_golden.py + tests/test_gpu_index.py.
constexpr int WS = 7, EXPAND = 3, NKEYS = 289, NKEYS_PAD = 320;
struct KeySrc {
  int level;   // 0 target full-res map (cyclic); 1 pooled target; 2..4 pooled reference frame 0..2
  int y, x;    // -1 => nn.Unfold zero fill (K = V = 0, logit mask -100)
};

__device__ __forceinline__ KeySrc cfm_key_source(int n, int wi, int wj, int nWh, int nWw, int Hp, int Wp) {
  KeySrc r;
  if (n < 49) {                                             // own window, row-major
    r.level = 0; r.y = WS * wi + n / WS; r.x = WS * wj + n % WS;
    return r;
  }
  if (n < 181) {                                            // 4 x 33 ring keys, valid_ind_rolled order
    const int m = n - 49, quad = m / 33, idx = m % 33;
    const bool down = quad < 2;                             // tl, tr: rows shifted +3 ; bl, br: -3
    const bool right = (quad & 1) == 0;                     // tl, bl: cols shifted +3 ; tr, br: -3
    int rr, cc;
    if (down) {                                             // rows 0..3 contribute 3 cols, rows 4..6 all 7
      if (idx < 12) { rr = idx / 3; cc = idx % 3 + (right ? 4 : 0); }
      else { rr = 4 + (idx - 12) / 7; cc = (idx - 12) % 7; }
    } else {                                                // rows 0..2 all 7 cols, rows 3..6 contribute 3
      if (idx < 21) { rr = idx / 7; cc = idx % 7; }
      else { rr = 3 + (idx - 21) / 3; cc = (idx - 21) % 3 + (right ? 4 : 0); }
    }
    const int dy = rr + (down ? EXPAND : -EXPAND), dx = cc + (right ? EXPAND : -EXPAND);
    r.level = 0;
    r.y = (WS * wi + dy + Hp) % Hp;                         // torch.roll on the zero-padded map: cyclic
    r.x = (WS * wj + dx + Wp) % Wp;
    return r;
  }
  int m, kc, stride, gh, gw;
  if (n < 206) { m = n - 181; kc = 5; stride = 1; gh = nWh; gw = nWw; r.level = 1; }
  else if (n < 255) { m = n - 206; kc = 7; stride = 1; gh = nWh; gw = nWw; r.level = 2; }
  else if (n < 280) { m = n - 255; kc = 5; stride = 2; gh = 2 * nWh; gw = 2 * nWw; r.level = 3; }
  else { m = n - 280; kc = 3; stride = 3; gh = 3 * nWh; gw = 3 * nWw; r.level = 4; }
  const int y = stride * wi + m / kc - kc / 2, x = stride * wj + m % kc - kc / 2;
  const bool ok = y >= 0 && y < gh && x >= 0 && x < gw;
  r.y = ok ? y : -1;
  r.x = ok ? x : -1;
  return r;
}

__global__ void cfm_key_sources_kernel(int Hp, int Wp, int32_t* out) {
  pdl_sync();
  const int nWh = Hp / WS, nWw = Wp / WS;
  const int w = blockIdx.x, n = threadIdx.x;
  if (n >= NKEYS) return;
  const KeySrc s = cfm_key_source(n, w / nWw, w % nWw, nWh, nWw, Hp, Wp);
  int32_t* o = out + (static_cast<int64_t>(w) * NKEYS + n) * 3;
  o[0] = s.level; o[1] = s.y; o[2] = s.x;
}

// Persistent CTAs, each bound to ONE head: the head's additive bias slice [49 x 320] fp32 is staged in shared memory once
// (it used to be re-read from L2 by every (window, head, clip) CTA: 56 KB per CTA, more than the K/V gather itself), then
// the CTA walks over (clip, window) items.  Per item the 289-key K/V sequence is gathered with cp.async into [320][32] fp16
// tiles whose 16-byte pieces are XOR-swizzled with (key >> 1) & 3 (64-byte rows, no padding, conflict-free for the mma B
// fragments and for ldmatrix).  8 warps: warp = (key half kg, 16 query rows); each 64-key step is split between the two
// warp groups (32 keys each, own online softmax), and the two partial (max, sum, O) are merged through shared memory at
// the end -- twice the warps per item, half the dependent chain.  105 KB per CTA, two CTAs (16 warps) per SM.
constexpr int CFM_BLD = 328;                                   // bias row pitch (floats): 8 mod 32 banks -> conflict-free LDS.64
constexpr int CFM_SMEM = 49 * CFM_BLD * 4 + 2 * NKEYS_PAD * 32 * 2 + NKEYS_PAD * 4;
constexpr int CFM_THREADS = 256;

__device__ __forceinline__ int cfm_swz(int key, int chunk) { return key * 32 + ((chunk ^ ((key >> 1) & 3)) << 3); }   // halves

__global__ void __launch_bounds__(CFM_THREADS, 2)
cfm_attention_kernel(const __half* __restrict__ qkv_t, const __half* __restrict__ kv_pooled,
                     const float* __restrict__ bias, __half* __restrict__ out, int B, int H, int W, int Hp, int Wp, int P,
                     int heads, float scale) {
  pdl_sync();
  constexpr int D = 32, C = 256;
  extern __shared__ __align__(16) uint8_t smem_raw[];
  float* bias_s = reinterpret_cast<float*>(smem_raw);             // [49][CFM_BLD]
  __half* Ks = reinterpret_cast<__half*>(smem_raw + 49 * CFM_BLD * 4);
  __half* Vs = Ks + NKEYS_PAD * 32;
  float* madd = reinterpret_cast<float*>(Vs + NKEYS_PAD * 32);    // additive mask per key slot
  float* xch = reinterpret_cast<float*>(Ks);                      // [4 warps][20][32 lanes] partial results of key half 1 (after the loop)

  const int nWh = Hp / WS, nWw = Wp / WS, nW = nWh * nWw;
  const int h = blockIdx.x % heads, slot = blockIdx.x / heads, nslots = gridDim.x / heads;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t = lane & 3;
  const int wrow = warp & 3, kg = warp >> 2;                      // 16-row block of the window, key half of every 64-key step

  // ---- the head's bias slice (rows >= 49 of the [64][320] table are never used)
  {
    const float* src = bias + static_cast<int64_t>(h) * 64 * NKEYS_PAD;
    for (int i = tid; i < 49 * (NKEYS_PAD / 4); i += CFM_THREADS) {
      const int r = i / (NKEYS_PAD / 4), c4 = i % (NKEYS_PAD / 4);
      ptx::cp_async16(bias_s + r * CFM_BLD + 4 * c4, src + r * NKEYS_PAD + 4 * c4);
    }
  }
  const int qa = wrow * 16 + g, qb = qa + 8;
  const float* biasA = bias_s + min(qa, 48) * CFM_BLD;            // rows >= 49 compute garbage that is never stored
  const float* biasB = bias_s + min(qb, 48) * CFM_BLD;
  const float sc = scale * LOG2E;
  const uint32_t vs_addr = ptx::smem_u32(Vs);

  for (int item = slot; item < B * nW; item += nslots) {
    const int b = item / nW, w = item - b * nW;
    const int wi = w / nWw, wj = w % nWw;
    const __half* tq = qkv_t + static_cast<int64_t>(b) * Hp * Wp * (3 * C);
    const __half* pk = kv_pooled + static_cast<int64_t>(b) * P * (2 * C);

    // ---- gather the assembled K/V sequence of this window/head: one key (2 x 64 bytes) per thread per pass,
    // its source coordinate evaluated once
    for (int n = tid; n < NKEYS_PAD; n += CFM_THREADS) {
      bool valid = false;
      if (n < NKEYS) {
        const KeySrc s = cfm_key_source(n, wi, wj, nWh, nWw, Hp, Wp);
        if (s.y >= 0) {
          valid = true;
          const __half *ksrc, *vsrc;
          if (s.level == 0) {
            const __half* row = tq + static_cast<int64_t>(s.y * Wp + s.x) * (3 * C) + h * D;
            ksrc = row + C; vsrc = row + 2 * C;
          } else {
            const int lw = s.level <= 2 ? nWw : (s.level == 3 ? 2 * nWw : 3 * nWw);
            const int base = s.level == 1 ? 0 : (s.level == 2 ? nW : (s.level == 3 ? 2 * nW : 6 * nW));
            const __half* row = pk + static_cast<int64_t>(base + s.y * lw + s.x) * (2 * C) + h * D;
            ksrc = row; vsrc = row + C;
          }
#pragma unroll
          for (int ch = 0; ch < 4; ++ch) {
            ptx::cp_async16(Ks + cfm_swz(n, ch), ksrc + ch * 8);
            ptx::cp_async16(Vs + cfm_swz(n, ch), vsrc + ch * 8);
          }
        }
      }
      if (!valid) {
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) {
          *reinterpret_cast<uint4*>(Ks + n * 32 + ch * 8) = make_uint4(0, 0, 0, 0);
          *reinterpret_cast<uint4*>(Vs + n * 32 + ch * 8) = make_uint4(0, 0, 0, 0);
        }
      }
      madd[n] = n >= NKEYS ? -INFINITY : (valid ? 0.f : -100.f * LOG2E);   // -100, not -inf (:445,:490)
    }
    ptx::cp_async_commit();

    // ---- Q fragments (window tokens, row-major inside the window)
    uint32_t a[2][4];
    {
      const bool va = qa < 49, vb = qb < 49;
      const __half* ra = tq + static_cast<int64_t>((WS * wi + qa / WS) * Wp + WS * wj + qa % WS) * (3 * C) + h * D;
      const __half* rb = tq + static_cast<int64_t>((WS * wi + qb / WS) * Wp + WS * wj + qb % WS) * (3 * C) + h * D;
#pragma unroll
      for (int kk = 0; kk < 2; ++kk) {
        a[kk][0] = va ? *reinterpret_cast<const uint32_t*>(ra + kk * 16 + 2 * t) : 0u;
        a[kk][1] = vb ? *reinterpret_cast<const uint32_t*>(rb + kk * 16 + 2 * t) : 0u;
        a[kk][2] = va ? *reinterpret_cast<const uint32_t*>(ra + kk * 16 + 8 + 2 * t) : 0u;
        a[kk][3] = vb ? *reinterpret_cast<const uint32_t*>(rb + kk * 16 + 8 + 2 * t) : 0u;
      }
    }
    ptx::cp_async_wait_all();                                    // K/V of this item (and, the first time, the bias slice)
    __syncthreads();

    float o[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
    float mA = -INFINITY, mB = -INFINITY, lA = 0.f, lB = 0.f;

#pragma unroll 1
    for (int c = kg * 32; c < NKEYS_PAD; c += 64) {              // this warp group's 32 keys of every 64-key step
      float s[4][4];
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f;
#pragma unroll
      for (int kk = 0; kk < 2; ++kk) {
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
          const int key = c + nt * 8 + g;
          ptx::mma_m16n8k16(s[nt], a[kk], *reinterpret_cast<const uint32_t*>(Ks + cfm_swz(key, 2 * kk) + 2 * t),
                            *reinterpret_cast<const uint32_t*>(Ks + cfm_swz(key, 2 * kk + 1) + 2 * t));
        }
      }
      float cmA = -INFINITY, cmB = -INFINITY;
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        const int key = c + nt * 8 + 2 * t;
        const float2 ba = *reinterpret_cast<const float2*>(biasA + key);
        const float2 bb = *reinterpret_cast<const float2*>(biasB + key);
        const float m0 = madd[key], m1 = madd[key + 1];
        s[nt][0] = fmaf(s[nt][0], sc, fmaf(ba.x, LOG2E, m0));
        s[nt][1] = fmaf(s[nt][1], sc, fmaf(ba.y, LOG2E, m1));
        s[nt][2] = fmaf(s[nt][2], sc, fmaf(bb.x, LOG2E, m0));
        s[nt][3] = fmaf(s[nt][3], sc, fmaf(bb.y, LOG2E, m1));
        cmA = fmaxf(cmA, fmaxf(s[nt][0], s[nt][1]));
        cmB = fmaxf(cmB, fmaxf(s[nt][2], s[nt][3]));
      }
      cmA = fmaxf(cmA, __shfl_xor_sync(0xffffffffu, cmA, 1));
      cmA = fmaxf(cmA, __shfl_xor_sync(0xffffffffu, cmA, 2));
      cmB = fmaxf(cmB, __shfl_xor_sync(0xffffffffu, cmB, 1));
      cmB = fmaxf(cmB, __shfl_xor_sync(0xffffffffu, cmB, 2));
      const float nmA = fmaxf(mA, cmA), nmB = fmaxf(mB, cmB);    // finite: every 32-key group holds a key with a finite mask
      const float alA = ptx::ex2_approx(mA - nmA), alB = ptx::ex2_approx(mB - nmB);
      mA = nmA; mB = nmB;
      float psA = 0.f, psB = 0.f;
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        s[nt][0] = ptx::ex2_approx(s[nt][0] - mA); s[nt][1] = ptx::ex2_approx(s[nt][1] - mA);
        s[nt][2] = ptx::ex2_approx(s[nt][2] - mB); s[nt][3] = ptx::ex2_approx(s[nt][3] - mB);
        psA += s[nt][0] + s[nt][1];
        psB += s[nt][2] + s[nt][3];
      }
      lA = lA * alA + psA;
      lB = lB * alB + psB;
#pragma unroll
      for (int nd = 0; nd < 4; ++nd) { o[nd][0] *= alA; o[nd][1] *= alA; o[nd][2] *= alB; o[nd][3] *= alB; }
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        uint32_t pa[4];
        pa[0] = pack_half2(s[2 * j][0], s[2 * j][1]);
        pa[1] = pack_half2(s[2 * j][2], s[2 * j][3]);
        pa[2] = pack_half2(s[2 * j + 1][0], s[2 * j + 1][1]);
        pa[3] = pack_half2(s[2 * j + 1][2], s[2 * j + 1][3]);
        const int krow = c + j * 16 + ((lane >> 3) & 1) * 8 + (lane & 7);
#pragma unroll
        for (int nd = 0; nd < 4; nd += 2) {
          uint32_t b0, b1, b2, b3;
          ptx::ldmatrix_x4_trans(b0, b1, b2, b3, vs_addr + cfm_swz(krow, nd + (lane >> 4)) * 2);
          ptx::mma_m16n8k16(o[nd], pa, b0, b1);
          ptx::mma_m16n8k16(o[nd + 1], pa, b2, b3);
        }
      }
    }
    lA += __shfl_xor_sync(0xffffffffu, lA, 1); lA += __shfl_xor_sync(0xffffffffu, lA, 2);
    lB += __shfl_xor_sync(0xffffffffu, lB, 1); lB += __shfl_xor_sync(0xffffffffu, lB, 2);
    __syncthreads();                                             // every warp is done with K (xch aliases it)
    float* xw = xch + wrow * (20 * 32) + lane;
    if (kg == 1) {
#pragma unroll
      for (int nd = 0; nd < 4; ++nd)
#pragma unroll
        for (int e = 0; e < 4; ++e) xw[(nd * 4 + e) * 32] = o[nd][e];
      xw[16 * 32] = mA; xw[17 * 32] = mB; xw[18 * 32] = lA; xw[19 * 32] = lB;
    }
    __syncthreads();
    if (kg == 0) {
      // merge the two key halves: (m, l, O) = (max, sum l_i 2^(m_i - m), sum O_i 2^(m_i - m))
      const float m1A = xw[16 * 32], m1B = xw[17 * 32];
      const float mxA = fmaxf(mA, m1A), mxB = fmaxf(mB, m1B);
      const float a0A = ptx::ex2_approx(mA - mxA), a1A = ptx::ex2_approx(m1A - mxA);
      const float a0B = ptx::ex2_approx(mB - mxB), a1B = ptx::ex2_approx(m1B - mxB);
      const float iA = 1.f / (lA * a0A + xw[18 * 32] * a1A), iB = 1.f / (lB * a0B + xw[19 * 32] * a1B);
      // window_reverse + crop to (H, W) fused into the store (cffm_transformer.py:812-821)
      const int ya = WS * wi + qa / WS, xa = WS * wj + qa % WS;
      const int yb = WS * wi + qb / WS, xb = WS * wj + qb % WS;
      const bool sa = qa < 49 && ya < H && xa < W, sb = qb < 49 && yb < H && xb < W;
      __half* oA = out + (static_cast<int64_t>(b) * H * W + ya * W + xa) * C + h * D;
      __half* oB = out + (static_cast<int64_t>(b) * H * W + yb * W + xb) * C + h * D;
#pragma unroll
      for (int nd = 0; nd < 4; ++nd) {
        const float v0 = (o[nd][0] * a0A + xw[(nd * 4 + 0) * 32] * a1A) * iA, v1 = (o[nd][1] * a0A + xw[(nd * 4 + 1) * 32] * a1A) * iA;
        const float v2 = (o[nd][2] * a0B + xw[(nd * 4 + 2) * 32] * a1B) * iB, v3 = (o[nd][3] * a0B + xw[(nd * 4 + 3) * 32] * a1B) * iB;
        if (sa) *reinterpret_cast<uint32_t*>(oA + nd * 8 + 2 * t) = pack_half2(v0, v1);
        if (sb) *reinterpret_cast<uint32_t*>(oB + nd * 8 + 2 * t) = pack_half2(v2, v3);
      }
    }
    __syncthreads();                                             // K/V tiles (and xch) are overwritten by the next item's gather
  }
}

constexpr int SMEM_CAP = 200 * 1024;
// Raises the dynamic shared-memory cap of `kernel` once per process (never inside a later stream capture).
template <auto Kernel>                                          // one static per kernel, not per signature
int set_smem_attr() {
  static cudaError_t err = cudaFuncSetAttribute(Kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_CAP);
  if (err != cudaSuccess) {
    set_error("cudaFuncSetAttribute(%d bytes): %s", SMEM_CAP, cudaGetErrorString(err));
    return -(int)err;
  }
  return CFFM_OK;
}

}  // namespace
}  // namespace cffm

namespace cffm {
int mha_tcgen05_launch(const void* q, int64_t ldq, const void* k, const void* v, int64_t ldkv, void* out, int64_t ldo,
                       int batch, int Nq, int Nkv, int heads, float scale, cudaStream_t st);
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
  if (head_dim == 64 && !mha_legacy()) {                       // tcgen05 path (mha_sm100.cu); falls through when unsupported
    const int rc5 = mha_tcgen05_launch(q, ldq, k, v, ldkv, out, ldo, batch, Nq, Nkv, heads, scale, st);
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

extern "C" int cffm_cfm_attention(const void* qkv_t, const void* kv_pooled, const float* bias, void* out, int B, int H,
                                  int W, int C, int heads, float scale, void* stream) {
  using namespace cffm;
  CFFM_REQUIRE(qkv_t && kv_pooled && bias && out, CFFM_E_BADARG, "cfm_attention: null pointer");
  CFFM_REQUIRE(B > 0 && H > 0 && W > 0, CFFM_E_BADARG, "cfm_attention: non-positive size");
  CFFM_REQUIRE(C == 256 && heads == 8, CFFM_E_UNSUPPORTED,
               "cfm_attention: built for C=256, heads=8 (cffm_head.py:74-95), got C=%d heads=%d", C, heads);
  CFFM_REQUIRE(aligned16(qkv_t) && aligned16(kv_pooled) && aligned16(bias) && aligned16(out), CFFM_E_BADARG,
               "cfm_attention: pointers must be 16-byte aligned");
  const int Hp = (H + WS - 1) / WS * WS, Wp = (W + WS - 1) / WS * WS;
  const int nW = (Hp / WS) * (Wp / WS);
  CFFM_REQUIRE(B <= 65535, CFFM_E_UNSUPPORTED, "cfm_attention: B too large");
  int rc = set_smem_attr<cfm_attention_kernel>();
  if (rc) return rc;
  // persistent CTAs, each bound to one head: (2 CTAs per SM) / heads slots per head walk over the B * nW (clip, window) items
  int slots = 2 * num_sms() / heads;
  if (slots > B * nW) slots = B * nW;
  if (slots < 1) slots = 1;
  launch_k(cfm_attention_kernel, slots * heads, CFM_THREADS, CFM_SMEM, static_cast<cudaStream_t>(stream),
      static_cast<const __half*>(qkv_t), static_cast<const __half*>(kv_pooled), bias, static_cast<__half*>(out), B, H, W,
      Hp, Wp, 15 * nW, heads, scale);
  return launch_status("cfm_attention_kernel");
}

extern "C" int cffm_cfm_key_sources(int Hp, int Wp, int32_t* out, void* stream) {
  using namespace cffm;
  CFFM_REQUIRE(out && Hp > 0 && Wp > 0 && Hp % WS == 0 && Wp % WS == 0, CFFM_E_BADARG,
               "cfm_key_sources: Hp, Wp must be positive multiples of 7");
  launch_k(cfm_key_sources_kernel, (Hp / WS) * (Wp / WS), 320, 0, static_cast<cudaStream_t>(stream), Hp, Wp, out);
  return launch_status("cfm_key_sources_kernel");
}
