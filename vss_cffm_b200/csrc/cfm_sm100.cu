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
// Every key row is 128 bytes = the 64 channels of the head pair, 128-byte swizzled, i.e. directly a tcgen05 operand; the
// five boxes are packed back to back (a TMA destination only needs 128-byte alignment: the swizzle is a function of the
// shared-memory address, so a box may start inside a 1024-byte swizzle atom -- verified bit for bit on B200).
//
// Per item:
//   S[128 x 288] = Qbd K^T  M = 128 rows = (head 0: 49 queries | pad to 64 | head 1: 49 queries | pad), K = 64 channels with
//                           Q block-diagonal (head-0 rows carry zeros in head 1's channels and vice versa), so ONE
//                           M = 128 MMA computes both heads at the tensor-core cost of two M = 64 ones.  S is produced and
//                           released in two halves (columns [0,128) and [128,288)).
//   softmax                 16 warps, thread = (row, quarter of each 64-key chunk); logits t = S + bias / scale straight out
//                           of TMEM (one FHFMA per element: fp16 bias operand, fp32 accumulator).  Pass 1: upper bound of the
//                           row maximum (raw S + the largest bias of the thread's piece; softmax is shift-invariant, so the
//                           result is the exact softmax).  Pass 2: p = 2^((t - m) scale log2 e), one FFMA + one MUFU.EX2 per
//                           element, packed to fp16 and written back into TENSOR MEMORY (tcgen05.st): P never goes through
//                           shared memory, there is no proxy fence in the item loop.
//   O[128 x 64] += P V      TS-mode tcgen05.mma: A = P chunk from TMEM (three-chunk ring), B = V chunk MN-major as loaded;
//                           only the diagonal 32-column blocks are read back
//   epilogue                O / rowsum -> fp16 -> swizzled staging tile -> whole 64-byte row segments; window_reverse + crop
//                           (:812-821) fused into the store
// Warp roles: 0..15 softmax / epilogue, 16 producer (TMA boxes under elect.sync, block-diagonal Q through cp.async, the item's
// key mask), 17 TMEM allocator + P V issuer, 18 Q K^T issuer (two issue warps on two schedulers: as one warp the P V issue fell
// ~1300 cycles behind the softmax).  K and Q are single-buffered (those of item i+1 are loaded while the softmax of item i
// runs), V is double-buffered (loaded a whole item ahead), O is double-buffered in TMEM.
#include <cuda.h>

#include <type_traits>

#include "common.cuh"
#include "ptx_sm100.cuh"

namespace cffm {
namespace {

constexpr int WS = 7, RING = 3, HALO = WS + 2 * RING, NHALO = HALO * HALO;      // 13, 169
constexpr int CQKV = 768, CKV = 512, CPAIR = 64;                                // channels: qkv row, pooled K|V row, head pair
constexpr int ROWB = 128;                                                       // bytes per key / query row in shared memory
constexpr int SM_WARPS = 16, CFM_THREADS = (SM_WARPS + 3) * 32;   // + producer, P V issuer, Q K^T issuer
constexpr int P_RING = 3, P_COLS = 32;                                          // P chunks (64 keys = 32 packed cells) in TMEM
constexpr int Q_BYTES = 128 * ROWB;
constexpr float LOG2E = 1.4426950408889634f;
#define CFM_PROF(slot) do { if (p.prof != nullptr && it < 8) p.prof[(blockIdx.x * 8 + it) * 32 + (slot)] = clock64(); } while (0)

// Key-row layout of the assembled sequence: halo | pooled target 5x5 | reference 0 7x7 | reference 1 5x5 | reference 2 3x3
constexpr int R1 = NHALO, R2 = R1 + 25, R3 = R2 + 49, R4 = R3 + 25, REND = R4 + 9;   // 169, 194, 243, 268, 277
constexpr int NPAD = 288;                                    // MMA N (two halves of 144); rows >= 277 are zero keys with a -inf bias
constexpr int BPITCH = 296;                                  // bias row pitch (halves): conflict-free 16-byte LDS per row
constexpr int NCH = (NPAD + 63) / 64;                        // 64-key P chunks, the last one holds 32 keys
constexpr int NCH_A = 2, N_A = NCH_A * 64, N_B = NPAD - N_A; // S is produced and released in two halves: columns [0,128) and [128,288)
constexpr int X_FLOATS = 2 * 4 * 128;                        // [max | sum][column quarter][row]
constexpr int KV_BYTES = NPAD * ROWB;
constexpr int TX_BYTES = REND * ROWB;
constexpr int BIAS_BYTES = 2 * 49 * BPITCH * 2;
constexpr int O_STAGE_BYTES = 4 * 32 * 64;                   // output staging tile of the epilogue
constexpr int CFM_SMEM = 3 * KV_BYTES + 2 * Q_BYTES + BIAS_BYTES + 2 * (NPAD + 4) * 4 /*mask + item coordinates*/ +
                         2 * X_FLOATS * 4 /*row max / sum exchange, two item parities*/ + O_STAGE_BYTES + 256 /*barriers*/ + 1024 /*align slack*/;
constexpr int TMEM_O = NPAD;                                 // two O buffers of 64 columns behind S ...
constexpr int TMEM_P = TMEM_O + 2 * CPAIR;                   // ... and the P ring behind them: 288 + 128 + 96 = 512 columns
static_assert(TMEM_P + P_RING * P_COLS <= 512, "TMEM budget");
static_assert(CFM_SMEM <= 232448, "shared memory budget of one CTA per SM");

struct CfmParams {
  const __half* qkv_a;      // [B, Hp+6, Wp+6, 768] apron layout
  const __half* bias_tab;   // [8 heads, 49, BPITCH] fp16, bias / scale in the kernel's key order, -inf on unused columns
  __half* out;              // [B, H, W, 256]
  __half* dump;             // test hook: [4 head pairs, B nW items, 2, NPAD, 64] assembled K and V tiles (or null)
  long long* prof;          // bring-up hook: per CTA / item / event SM clock stamps [grid, 8, 32] (or null)
  const int32_t* ref_slot;  // [B, 3, 2] (slot, rank) = 4th and 5th coordinate of the level map of reference frame k of clip b, or null: (b, 0)
  int B, H, W, nWh, nWw;
  float scale_log2e;
};

struct CfmBars {
  uint64_t *s_full, *s_empty, *m_full, *p_full, *p_empty, *o_full, *o_empty;
};

// Issue (no wait) the TMEM load of this thread's piece of chunk C: S[row, 64 C + cq w .. + w), w = 16 (8 in the last chunk)
template <int C>
__device__ __forceinline__ void cfm_issue_load(uint32_t lane_addr, int cq, uint32_t* u) {
  if (C * 64 + 64 <= NPAD) ptx::tmem_ld_32x32b_x16(lane_addr + C * 64 + cq * 16, u);
  else ptx::tmem_ld_32x32b_x8(lane_addr + C * 64 + cq * 8, u);
}

// v[j] = u[j] (+ mask): raw accumulator, nn.Unfold padding positions of the pooled levels at -inf.  Chunks 0 and 1 hold halo
// keys only (never masked); the mask of a halo key in the later chunks is 0.
template <int C, int WCOL>
__device__ __forceinline__ void cfm_masked(const uint32_t* u, const float* mask_piece, float* v) {
#pragma unroll
  for (int j = 0; j < WCOL / 4; ++j) {
    if (C * 64 + 64 > R1) {
      const float4 mk = *reinterpret_cast<const float4*>(mask_piece + j * 4);
      v[j * 4] = __uint_as_float(u[j * 4]) + mk.x; v[j * 4 + 1] = __uint_as_float(u[j * 4 + 1]) + mk.y;
      v[j * 4 + 2] = __uint_as_float(u[j * 4 + 2]) + mk.z; v[j * 4 + 3] = __uint_as_float(u[j * 4 + 3]) + mk.w;
    } else {
      v[j * 4] = __uint_as_float(u[j * 4]); v[j * 4 + 1] = __uint_as_float(u[j * 4 + 1]);
      v[j * 4 + 2] = __uint_as_float(u[j * 4 + 2]); v[j * 4 + 3] = __uint_as_float(u[j * 4 + 3]);
    }
  }
}

// ---- softmax / epilogue role.  16 warps: thread = (row of S, quarter cq of every 64-key chunk).  All 16 warps run the SAME
// instruction stream (cq is a run-time, warp-uniform offset): with one template instantiation per quarter the four warps
// of a scheduler executed four different 13 KB code streams and the kernel was bound by instruction fetch (ncu: 7.6
// warps stalled on "no instruction" per issued instruction).  Machine facts this layout follows (measured on B200:
// tools/micro and the in-kernel timeline of tools/cfm_timeline.py):
//   * tcgen05.ld costs ~3 cycles per column per warp and overlaps across the warps of a TMEM lane quarter: every load is
//     issued one chunk ahead of its use (two register sets), so its latency hides behind the exp2 work of the chunk before;
//   * fence.proxy.async is a ~110-cycle MEMBAR that serialises across warps, and an mbarrier wait costs ~100 cycles even
//     when the phase is already complete: P therefore never goes through shared memory (tcgen05.st into TMEM, A operand of
//     a TS-mode MMA), and there is one mbarrier wait per chunk, placed ahead of the math;
//   * MUFU.EX2 (16 lanes / clock / SM: 512 cycles per 64-key chunk) is the floor of pass 2.
__device__ __forceinline__ void softmax_role(const CfmParams& p, const __half* sBias, const float* sMask, float* sX, uint8_t* sO,
                                          const CfmBars& bar, uint32_t tmem_base, int hp, int slot, int nslots) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int wq = warp & 3, cq = warp >> 2;                   // TMEM lane quarter, column quarter of every chunk
  const int row = wq * 32 + lane;                            // row of S = TMEM lane
  const int hl = row >> 6, q = row & 63;                     // head of the pair, query index (>= 49: padding)
  const int qc = q < 49 ? q : 48;
  const __half* brow = sBias + (hl * 49 + qc) * BPITCH;
  const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(wq * 32) << 16);
  const int stid = threadIdx.x;
  const int n_items = p.B * p.nWh * p.nWw;
  const float scl = p.scale_log2e;
  const int bar_rows = 1 + wq;                               // named barrier of the four warps that share these 32 rows

  // Largest bias of this thread's piece of every chunk (a constant of the head pair).  Pass 1 then needs no bias at all:
  // max_k (S_k + bias_k) <= max over pieces of (max_k S_k + max_k bias_k) =: m, an upper bound that is exact up to the
  // spread of the bias inside a 16-key piece.  Softmax is invariant to the shift; p = 2^((t - m) ...) <= 1 always.
  float bmax[NCH];
#pragma unroll
  for (int c = 0; c < NCH; ++c) {
    const int wcol = (c * 64 + 64 <= NPAD) ? 16 : 8, col0 = c * 64 + cq * wcol;
    float m = -INFINITY;
    for (int j = 0; j < wcol; ++j) m = fmaxf(m, __half2float(brow[col0 + j]));
    bmax[c] = m;                                             // -inf for a piece of unused key rows only
  }

  // epilogue of an item: O[row, 32 hl + 8 cq .. +8) / rowsum -> fp16 -> out (window_reverse + crop, :812-821).  (Running it
  // one item late, between the two passes of the next item, was measured slower: it then sits on the path to the first P chunk.)
  auto epilogue = [&](uint32_t e_it, int b, int wi, int wj, float inv) {
    const uint32_t ob = e_it & 1u;
    ptx::mbar_wait(&bar.o_full[ob], (e_it >> 1) & 1u);
    ptx::tc_fence_after();
    { const uint32_t it = e_it; if (stid == 0) CFM_PROF(17); }
    uint32_t o[8];
    ptx::tmem_ld_32x32b_x8(lane_addr + TMEM_O + ob * 64 + hl * 32 + cq * 8, o);
    ptx::tmem_ld_wait();
    ptx::tc_fence_before();
    __syncwarp();
    if (lane == 0) ptx::mbar_arrive(&bar.o_empty[ob]);
    { const uint32_t it = e_it; if (stid == 0) CFM_PROF(18); }
    uint4 wv;
    wv.x = pack_half2(__uint_as_float(o[0]) * inv, __uint_as_float(o[1]) * inv);
    wv.y = pack_half2(__uint_as_float(o[2]) * inv, __uint_as_float(o[3]) * inv);
    wv.z = pack_half2(__uint_as_float(o[4]) * inv, __uint_as_float(o[5]) * inv);
    wv.w = pack_half2(__uint_as_float(o[6]) * inv, __uint_as_float(o[7]) * inv);
    // A thread-per-row store sends 32 separate 16-byte pieces per instruction through the L1 (32 wavefronts).  The 32 rows x 64
    // bytes of this lane quarter go through a swizzled staging tile instead (piece j of row r at j ^ ((r >> 1) & 3): conflict-
    // free both ways) and leave as whole 64-byte row segments, eight rows per instruction.
    uint8_t* sq = sO + wq * (32 * 64);
    *reinterpret_cast<uint4*>(sq + lane * 64 + ((cq ^ ((lane >> 1) & 3)) << 4)) = wv;
    asm volatile("bar.sync %0, 128;" ::"r"(bar_rows) : "memory");
    {
      const int t = cq * 32 + lane, rl = t >> 2, j = t & 3;
      const int q2 = (wq * 32 + rl) & 63;                    // query of the staged row (same head: a quarter never straddles 64)
      const uint4 val = *reinterpret_cast<const uint4*>(sq + rl * 64 + ((j ^ ((rl >> 1) & 3)) << 4));
      const int y = WS * wi + q2 / WS, x = WS * wj + q2 % WS;
      if (q2 < 49 && y < p.H && x < p.W)
        *reinterpret_cast<uint4*>(p.out + ((static_cast<int64_t>(b) * p.H + y) * p.W + x) * 256 + hp * CPAIR + hl * 32 + j * 8) = val;
    }
    { const uint32_t it = e_it; if (stid == 0) CFM_PROF(19); }
  };

  uint32_t it = 0;
  float pinv = 0.f;
  for (int item = slot; item < n_items; item += nslots, ++it) {
    const float* mask = sMask + (it & 1u) * (NPAD + 4);
    float* xm = sX + (it & 1u) * X_FLOATS;
    uint32_t sb[2][16];                                      // two register sets: chunk C lives in sb[C & 1]
    if (stid == 0) CFM_PROF(8);
    ptx::mbar_wait(&bar.s_full[0], it & 1u);
    if (stid == 0) CFM_PROF(14);
    ptx::mbar_wait(&bar.m_full[it & 1u], (it >> 1) & 1u);
    if (stid == 0) CFM_PROF(15);
    // {clip, window row, window column} of this item: read NOW -- after the last read of S the producer may rewrite the buffer
    const int4 info = *reinterpret_cast<const int4*>(mask + NPAD);
    const int cb = info.x, cwi = info.y, cwj = info.z;
    ptx::tc_fence_after();
    if (stid == 0) CFM_PROF(9);
    cfm_issue_load<0>(lane_addr, cq, sb[0]);

    // ---- pass 1: upper bound of the row maximum
    float mx = -INFINITY;
    auto pass1 = [&](auto cc) {
      constexpr int C = decltype(cc)::value;
      constexpr int WCOL = (C * 64 + 64 <= NPAD) ? 16 : 8;
      ptx::tmem_ld_wait();                                   // chunk C has landed
      if (C + 1 == NCH_A) {                                  // the next chunk is the first of the second half of S
        ptx::mbar_wait(&bar.s_full[1], it & 1u);
        ptx::tc_fence_after();
      }
      if (C + 1 < NCH) cfm_issue_load<C + 1>(lane_addr, cq, sb[(C + 1) & 1]);
      else cfm_issue_load<0>(lane_addr, cq, sb[(C + 1) & 1]);      // chunk 0 again, for pass 2
      float v[WCOL];
      cfm_masked<C, WCOL>(sb[C & 1], mask + C * 64 + cq * WCOL, v);
      float m0 = fmaxf(v[0], v[1]), m1 = fmaxf(v[2], v[3]);
#pragma unroll
      for (int j = 4; j < WCOL; j += 4) { m0 = fmaxf(m0, fmaxf(v[j], v[j + 1])); m1 = fmaxf(m1, fmaxf(v[j + 2], v[j + 3])); }
      mx = fmaxf(mx, fmaxf(m0, m1) + bmax[C]);
    };
    pass1(std::integral_constant<int, 0>{}); pass1(std::integral_constant<int, 1>{}); pass1(std::integral_constant<int, 2>{});
    pass1(std::integral_constant<int, 3>{}); pass1(std::integral_constant<int, 4>{});
    xm[cq * 128 + row] = mx;
    asm volatile("bar.sync %0, 128;" ::"r"(bar_rows) : "memory");
    mx = fmaxf(fmaxf(xm[row], xm[128 + row]), fmaxf(xm[256 + row], xm[384 + row]));   // the halo keys are never masked: finite
    const float msc = mx * scl;
    if (stid == 0) CFM_PROF(10);

    // ---- pass 2: p = 2^((S + bias / scale - m) scale log2 e), row sum, packed fp16 P chunks into tensor memory.
    // Chunk C of pass 2 sits in sb[(NCH + C) & 1] (pass 1 left chunk 0 in flight in sb[NCH & 1]).
    float sum0 = 0.f, sum1 = 0.f;
    auto pass2 = [&](auto cc) {
      constexpr int C = decltype(cc)::value;
      constexpr int WCOL = (C * 64 + 64 <= NPAD) ? 16 : 8;
      const int col0 = C * 64 + cq * WCOL;
      const uint32_t gc = it * NCH + C, ps = gc % P_RING;
      ptx::tmem_ld_wait();                                   // chunk C has landed
      if (C + 1 < NCH) cfm_issue_load<C + 1>(lane_addr, cq, sb[(NCH + C + 1) & 1]);
      if (C == NCH_A - 1 || C == NCH - 1) {                  // this half of S has been read for the last time
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&bar.s_empty[C == NCH - 1 ? 1 : 0]);
      }
      // the ring slot is almost always free already; the probe is issued here and its ~100-cycle latency hides behind the math
      const bool slot_free = ptx::mbar_test_wait(&bar.p_empty[ps], ((gc / P_RING) & 1u) ^ 1u);
      uint4 bw[WCOL / 8];
#pragma unroll
      for (int j = 0; j < WCOL / 8; ++j) bw[j] = reinterpret_cast<const uint4*>(brow + col0)[j];
      float v[WCOL];
      cfm_masked<C, WCOL>(sb[(NCH + C) & 1], mask + col0, v);
      uint32_t hh[WCOL / 2];
#pragma unroll
      for (int j = 0; j < WCOL / 8; ++j) {
        const uint32_t ww[4] = {bw[j].x, bw[j].y, bw[j].z, bw[j].w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          float t0, t1;
          ptx::fhfma2(ww[e], v[j * 8 + 2 * e], v[j * 8 + 2 * e + 1], t0, t1);      // S + bias / scale: one FHFMA per element
          const float p0 = ptx::ex2_approx(fmaf(t0, scl, -msc));
          const float p1 = ptx::ex2_approx(fmaf(t1, scl, -msc));
          sum0 += p0; sum1 += p1;
          hh[j * 4 + e] = pack_half2(p0, p1);
        }
      }
      if (!slot_free) ptx::mbar_wait(&bar.p_empty[ps], ((gc / P_RING) & 1u) ^ 1u);
      ptx::tc_fence_after();
      // P stays in tensor memory (A operand of the P V MMA): packed fp16 pairs, 8 keys = 4 cells.  Published at once: P V of
      // the last chunks is on the path to the epilogue (publishing one chunk late was measured slower).
      const uint32_t pa = lane_addr + TMEM_P + ps * P_COLS + cq * (WCOL / 2);
      if (WCOL == 16) ptx::tmem_st_32x32b_x8(pa, hh);
      else ptx::tmem_st_32x32b_x4(pa, hh);
      ptx::tmem_st_wait();
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&bar.p_full[ps]);
    };
    pass2(std::integral_constant<int, 0>{}); pass2(std::integral_constant<int, 1>{}); pass2(std::integral_constant<int, 2>{});
    pass2(std::integral_constant<int, 3>{}); pass2(std::integral_constant<int, 4>{});
    if (stid == 0) CFM_PROF(11);
    xm[512 + cq * 128 + row] = sum0 + sum1;
    asm volatile("bar.sync %0, 128;" ::"r"(bar_rows) : "memory");
    pinv = 1.f / ((xm[512 + row] + xm[640 + row]) + (xm[768 + row] + xm[896 + row]));   // same order in all four threads
    if (stid == 0) CFM_PROF(12);
    epilogue(it, cb, cwi, cwj, pinv);
  }
}

__global__ void __launch_bounds__(CFM_THREADS, 1)
cfm_attention_tc_kernel(const __grid_constant__ CUtensorMap tmT, const __grid_constant__ CUtensorMap tmL1,
                        const __grid_constant__ CUtensorMap tmL2, const __grid_constant__ CUtensorMap tmL3,
                        const __grid_constant__ CUtensorMap tmL4, const CfmParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = ptx::smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);
  uint8_t* sK = smem;
  uint8_t* sV = sK + KV_BYTES;                                 // 2 buffers
  uint8_t* sQ = sV + 2 * KV_BYTES;                             // 2 buffers
  __half* sBias = reinterpret_cast<__half*>(sQ + 2 * Q_BYTES);
  float* sMask = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(sBias) + BIAS_BYTES);
  float* sX = sMask + 2 * (NPAD + 4);
  uint8_t* sO = reinterpret_cast<uint8_t*>(sX + 2 * X_FLOATS);   // output staging: 4 lane quarters x [32 rows][64 B]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sO + O_STAGE_BYTES);
  uint64_t* k_full = bars + 0;     // TMA -> Q K^T issuer: K tile landed (transaction bytes)
  uint64_t* kq_empty = bars + 1;   // Q K^T issuer -> producer: Q K^T of an item retired: K and that item's Q buffer may be overwritten
  uint64_t* q_full = bars + 2;     // [2] producer warp -> Q K^T issuer: block-diagonal Q written
  uint64_t* v_full = bars + 4;     // [2] TMA -> P V issuer: V tile landed
  uint64_t* v_empty = bars + 6;    // [2] P V issuer -> producer: P V retired
  CfmBars bar;
  bar.s_full = bars + 8;           // [2] Q K^T issuer -> softmax: half of S complete
  bar.s_empty = bars + 10;         // [2] softmax (16 warps) -> Q K^T issuer: half of S read for the last time
  // [2] producer warp -> softmax: the item's key mask written.  One barrier per mask buffer (item parity): the producer is
  // gated by the MMA side, not by the softmax warps, and with a single barrier it can complete the phases of items i and
  // i+1 before a slow softmax warp has waited for item i -- whose parity test would then never succeed.
  bar.m_full = bars + 12;
  bar.p_full = bars + 14;          // [P_RING] softmax (16 warps) -> P V issuer: P chunk written
  bar.p_empty = bars + 14 + P_RING;      // [P_RING] P V issuer -> softmax: P chunk consumed
  bar.o_full = bars + 14 + 2 * P_RING;   // [2] P V issuer -> epilogue: O complete
  bar.o_empty = bars + 16 + 2 * P_RING;  // [2] epilogue (16 warps) -> P V issuer: O read out
  uint32_t* tmem_base_smem = reinterpret_cast<uint32_t*>(bars + 18 + 2 * P_RING);
  static_assert(18 + 2 * P_RING < 32, "barrier area");

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int hp = blockIdx.x & 3, slot = blockIdx.x >> 2, nslots = gridDim.x >> 2;
  const int nW = p.nWh * p.nWw, n_items = p.B * nW;
  if (threadIdx.x == 0 && p.prof != nullptr) {
    long long gt;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
    p.prof[(blockIdx.x * 8 + 7) * 32 + 26] = gt;
    p.prof[(blockIdx.x * 8 + 7) * 32 + 28] = clock64();
  }

  if (warp == SM_WARPS && lane == 0) {
    ptx::prefetch_tensormap(&tmT); ptx::prefetch_tensormap(&tmL1); ptx::prefetch_tensormap(&tmL2);
    ptx::prefetch_tensormap(&tmL3); ptx::prefetch_tensormap(&tmL4);
    ptx::mbar_init(k_full, 1); ptx::mbar_init(kq_empty, 1);
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(&q_full[i], 1);
      ptx::mbar_init(&bar.m_full[i], 1);
      ptx::mbar_init(&bar.s_full[i], 1); ptx::mbar_init(&bar.s_empty[i], SM_WARPS);
      ptx::mbar_init(&v_full[i], 1); ptx::mbar_init(&v_empty[i], 1);
      ptx::mbar_init(&bar.o_full[i], 1); ptx::mbar_init(&bar.o_empty[i], SM_WARPS);
    }
    for (int i = 0; i < P_RING; ++i) { ptx::mbar_init(&bar.p_full[i], SM_WARPS); ptx::mbar_init(&bar.p_empty[i], 1); }
    ptx::fence_barrier_init();
  }
  if (warp == SM_WARPS + 1) {
    ptx::tmem_alloc(tmem_base_smem, 512);
    ptx::tmem_relinquish();
  }
  // Per-lane constants of the producer's item loop, computed once by ALL threads (one entry each: the divisions would cost
  // the producer warp ~3000 cycles on its own).  Q piece i (2 x 49 rows x four 16-byte pieces): source offset relative to the
  // window origin and destination in the block-diagonal tile (rows 0..48 = head 0 in bytes 0..63, rows 64..112 = head 1 in
  // bytes 64..127).  Key row n of a pooled level sits at (st wi + dy, st wj + dx) of an (f nWh) x (f nWw) map.
  __shared__ int2 q_tab[2 * 49 * 4];
  __shared__ int m_tab[NPAD];                                 // dy + 8 | (dx + 8) << 8 | st << 16 | f << 20, or 0: never masked
  {
    const int Wa_ = p.nWw * WS + 2 * RING;
    for (int i = threadIdx.x; i < 2 * 49 * 4; i += CFM_THREADS) {
      const int pc = i & 3, rq = i >> 2, hl = rq >= 49 ? 1 : 0, q = rq - hl * 49, r = hl * 64 + q;
      q_tab[i] = make_int2(((q / WS) * Wa_ + q % WS) * CQKV + hl * 32 + pc * 8, r * ROWB + (((hl * 4 + pc) ^ (r & 7)) << 4));
    }
    for (int n = threadIdx.x; n < NPAD; n += CFM_THREADS) {
      int m, kc, st, f;
      if (n >= R1 && n < R2) { m = n - R1; kc = 5; st = 1; f = 1; }
      else if (n >= R2 && n < R3) { m = n - R2; kc = 7; st = 1; f = 1; }
      else if (n >= R3 && n < R4) { m = n - R3; kc = 5; st = 2; f = 2; }
      else if (n >= R4 && n < REND) { m = n - R4; kc = 3; st = 3; f = 3; }
      else { m = 0; kc = 1; st = 0; f = 0; }
      m_tab[n] = f ? ((m / kc - kc / 2 + 8) | ((m % kc - kc / 2 + 8) << 8) | (st << 16) | (f << 20)) : 0;
    }
  }
  // K, V, Q tiles start as zeros: rows the TMA boxes never touch (unused key rows, query rows >= 49, the other head's
  // channels of the block-diagonal Q) must stay finite -- they meet P = 0 or are never read back.
  for (int i = threadIdx.x; i < 2 * Q_BYTES / 16; i += CFM_THREADS) reinterpret_cast<uint4*>(sQ)[i] = make_uint4(0u, 0u, 0u, 0u);
  for (int i = threadIdx.x; i < 3 * (NPAD - REND) * (ROWB / 16); i += CFM_THREADS) {      // key rows [277, 288) of K, V, V
    const int t = i / ((NPAD - REND) * (ROWB / 16)), r = i % ((NPAD - REND) * (ROWB / 16));
    reinterpret_cast<uint4*>(sK + t * KV_BYTES + REND * ROWB)[r] = make_uint4(0u, 0u, 0u, 0u);
  }
  if (warp < SM_WARPS) {
    // the head pair's bias slice: a constant table (not produced by the previous kernel), so it is fetched before the
    // programmatic-dependent-launch wait and overlaps the tail of the QKV GEMM
    const uint4* src = reinterpret_cast<const uint4*>(p.bias_tab + static_cast<int64_t>(hp) * 2 * 49 * BPITCH);
    for (int i = threadIdx.x; i < BIAS_BYTES / 16; i += SM_WARPS * 32)
      ptx::cp_async16(reinterpret_cast<uint4*>(sBias) + i, src + i);
    ptx::cp_async_commit();
  }
  ptx::tc_fence_before();
  __syncthreads();                                           // zero fill, barrier inits, TMEM address: visible to every warp
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_base_smem;
  if (threadIdx.x == 0 && p.prof != nullptr) p.prof[(blockIdx.x * 8 + 7) * 32 + 29] = clock64();
  pdl_sync();
  if (threadIdx.x == 0 && p.prof != nullptr) p.prof[(blockIdx.x * 8 + 7) * 32 + 30] = clock64();

  if (warp == SM_WARPS) {
    // ===================== producer: TMA K / V boxes, block-diagonal Q, key mask =====================
    const int Wa = p.nWw * WS + 2 * RING, Ha = p.nWh * WS + 2 * RING;
    struct Win { int b, wi, wj; };                             // clip, window row, window column of an item
    auto boxes = [&](uint8_t* dst, uint64_t* fb, const Win& c, int ct, int cp) {   // ct / cp: channel of K (or V) in qkv / pooled rows
      const int b = c.b, wi = c.wi, wj = c.wj;
      // where the three reference frames of this clip live (frame-sharded runs read the all-gathered buffer in place)
      int sl[3] = {b, b, b}, rk[3] = {0, 0, 0};
      if (p.ref_slot != nullptr) {
#pragma unroll
        for (int k = 0; k < 3; ++k) { sl[k] = p.ref_slot[6 * b + 2 * k]; rk[k] = p.ref_slot[6 * b + 2 * k + 1]; }
      }
      ptx::mbar_arrive_expect_tx(fb, TX_BYTES);
      ptx::tma_load_4d(dst, &tmT, fb, ct, WS * wj, WS * wi, b);
      ptx::tma_load_4d(dst + R1 * ROWB, &tmL1, fb, cp, wj - 2, wi - 2, b);
      ptx::tma_load_5d(dst + R2 * ROWB, &tmL2, fb, cp, wj - 3, wi - 3, sl[0], rk[0]);
      ptx::tma_load_5d(dst + R3 * ROWB, &tmL3, fb, cp, 2 * wj - 2, 2 * wi - 2, sl[1], rk[1]);
      ptx::tma_load_5d(dst + R4 * ROWB, &tmL4, fb, cp, 3 * wj - 1, 3 * wi - 1, sl[2], rk[2]);
    };
    auto dump_tile = [&](const uint8_t* tile, int item, int which) {           // test hook: an assembled tile, de-swizzled
      __half* d = p.dump + ((static_cast<int64_t>(hp) * n_items + item) * 2 + which) * NPAD * CPAIR;
      for (int i = lane; i < NPAD * 8; i += 32) {
        const int n = i >> 3, pc = i & 7;
        *reinterpret_cast<uint4*>(d + n * CPAIR + pc * 8) = *reinterpret_cast<const uint4*>(tile + n * ROWB + ((pc ^ (n & 7)) << 4));
      }
    };
    ptx::fence_proxy_async();                                // the zero fill (generic proxy) is ordered before the first TMA write
    // per-lane constants of the item loop (this warp shares its scheduler with four softmax warps: keep it lean)
    constexpr int QP = (2 * 49 * 4 + 31) / 32, MP = NPAD / 32;
    int q_src[QP], q_dst[QP], m_par[MP];
#pragma unroll
    for (int j = 0; j < QP; ++j) {
      const int i = lane + 32 * j;
      const int2 e = q_tab[i < 2 * 49 * 4 ? i : 0];
      q_src[j] = i < 2 * 49 * 4 ? e.x : -1;
      q_dst[j] = e.y;
    }
#pragma unroll
    for (int j = 0; j < MP; ++j) m_par[j] = m_tab[lane + 32 * j];
    // Q of an item: 13 cp.async per lane into the block-diagonal tile `buf`; published on q_full[buf] by finish_q
    auto issue_q = [&](const Win& c, uint32_t buf) {
      const int b = c.b, wi = c.wi, wj = c.wj;
      const __half* qbase = p.qkv_a + ((static_cast<int64_t>(b) * Ha + WS * wi + RING) * Wa + WS * wj + RING) * CQKV + hp * CPAIR;
#pragma unroll
      for (int j = 0; j < QP; ++j)
        if (q_src[j] >= 0) ptx::cp_async16(sQ + buf * Q_BYTES + q_dst[j], qbase + q_src[j]);
      ptx::cp_async_commit();
    };
    auto finish_q = [&](uint32_t buf) {
      ptx::cp_async_wait_all();
      ptx::fence_proxy_async();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&q_full[buf]);
    };
    // Items advance by nslots windows: the coordinates are stepped without divisions (this warp shares its scheduler with four
    // busy softmax warps and gets a fifth of the issue slots: sixteen integer divisions per item were most of its work).
    const int dq = nslots / p.nWw, dr = nslots - dq * p.nWw;
    auto step = [&](Win c) {
      c.wj += dr; c.wi += dq;
      if (c.wj >= p.nWw) { c.wj -= p.nWw; ++c.wi; }
      while (c.wi >= p.nWh) { c.wi -= p.nWh; ++c.b; }
      return c;
    };
    Win cur;
    cur.b = slot / nW;
    cur.wi = (slot - cur.b * nW) / p.nWw;
    cur.wj = slot - cur.b * nW - cur.wi * p.nWw;
    if (slot < n_items) {                                    // first item: K and Q first (Q K^T needs them), V behind
      if (ptx::elect_one()) boxes(sK, k_full, cur, 256 + hp * CPAIR, hp * CPAIR);
      __syncwarp();
      issue_q(cur, 0);
      if (ptx::elect_one()) boxes(sV, &v_full[0], cur, 512 + hp * CPAIR, 256 + hp * CPAIR);
      __syncwarp();
    }
    uint32_t it = 0;
    for (int item = slot; item < n_items; item += nslots, ++it) {
      const int next = item + nslots;
      const Win nxt = step(cur);
      // Q K^T of item it-1 has retired: K and the Q buffer of item it-1 (= that of item it+1) are free.  K goes out first: it
      // is single-buffered and its latency is the one on the path to the next S.
      if (it > 0) {
        ptx::mbar_wait(kq_empty, (it - 1) & 1u);
        if (lane == 0) CFM_PROF(0);
        if (ptx::elect_one()) boxes(sK, k_full, cur, 256 + hp * CPAIR, hp * CPAIR);
        __syncwarp();
      }
      finish_q(it & 1u);                                     // Q of THIS item: its copies were issued a whole item ago
      if (lane == 0) CFM_PROF(1);
      // The item's key mask; its buffer was last read for item it-2, whose reads precede kq_empty(it-1).  (The release arrive
      // below waits for this warp's cp.async copies in flight: with the next item's Q issued ahead of it, m_full came a whole
      // global-memory latency late and the 16 softmax warps waited ~600 cycles at the top of every item.)
      float* mask = sMask + (it & 1u) * (NPAD + 4);
      if (lane < 3) reinterpret_cast<int*>(mask + NPAD)[lane] = lane == 0 ? cur.b : (lane == 1 ? cur.wi : cur.wj);   // for the epilogue
#pragma unroll
      for (int j = 0; j < MP; ++j) {
        const int mp = m_par[j], st = (mp >> 16) & 15, f = mp >> 20;
        const int y = st * cur.wi + (mp & 255) - 8, x = st * cur.wj + ((mp >> 8) & 255) - 8;
        mask[lane + 32 * j] = (f == 0 || (y >= 0 && y < f * p.nWh && x >= 0 && x < f * p.nWw)) ? 0.f : -INFINITY;
      }
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&bar.m_full[it & 1u]);
      if (p.dump != nullptr) {
        ptx::mbar_wait(k_full, it & 1u);
        dump_tile(sK, item, 0);
        ptx::mbar_wait(&v_full[it & 1u], (it >> 1) & 1u);
        dump_tile(sV + (it & 1u) * KV_BYTES, item, 1);
        __syncwarp();
      }
      if (next < n_items) {
        // V of the NEXT item, a whole item ahead of its first use, then its Q copies (published at the top of the next round)
        const uint32_t nb = (it + 1) & 1u;
        if (it >= 1) ptx::mbar_wait(&v_empty[nb], ((it - 1) >> 1) & 1u);
        if (lane == 0) CFM_PROF(2);
        if (ptx::elect_one()) boxes(sV + nb * KV_BYTES, &v_full[nb], nxt, 512 + hp * CPAIR, 256 + hp * CPAIR);
        __syncwarp();
        issue_q(nxt, nb);
      }
      cur = nxt;
    }
  } else if (warp == SM_WARPS + 1) {
    // ===================== P V issuer =====================
    // The two MMA roles are separate warps on separate schedulers: each is a short, latency-bound instruction stream that
    // competes with four busy softmax warps for issue slots; as ONE warp the issue of P V fell ~1300 cycles behind the
    // softmax by the end of every item.  In both, the whole warp walks the loop and waits on the barriers and ONE elected
    // lane issues the tcgen05 instructions (entering the role with `if (lane == 0)` makes every descriptor a per-thread
    // value: each UTCHMMA is then wrapped in an ELECT / R2UR.BROADCAST / BRA.U.ANY loop of ~15 dependent instructions).
    constexpr uint32_t idesc_pv = ptx::make_idesc_f16(128, CPAIR) | (1u << 16);         // B (= V) MN-major
    const uint64_t dv0 = ptx::make_smem_desc_sw128(ptx::smem_u32(sV));
    const uint32_t my_items = slot < n_items ? (n_items - slot + nslots - 1) / nslots : 0u;
    if (lane == 0) ptx::fence_proxy_async();                 // the zero fill of the pad rows -> async proxy (tensor core reads)
    __syncwarp();
    for (uint32_t it = 0; it < my_items; ++it) {
      const uint32_t ob = it & 1u, vb = it & 1u;
      const uint32_t tmem_o = tmem_base + TMEM_O + ob * 64;
      uint64_t dv = dv0 + static_cast<uint64_t>(vb * (KV_BYTES / 16));
      uint32_t gc = it * NCH;
#pragma unroll
      for (int c = 0; c < NCH; ++c, ++gc) {
        const uint32_t ps = gc % P_RING;
        ptx::mbar_wait(&bar.p_full[ps], (gc / P_RING) & 1u);
        if (c == 0) {
          if (lane == 0) CFM_PROF(6);
          ptx::mbar_wait(&v_full[vb], (it >> 1) & 1u);
          if (it >= 2) ptx::mbar_wait(&bar.o_empty[ob], ((it - 2) >> 1) & 1u);
          if (lane == 0) CFM_PROF(7);
        }
        ptx::tc_fence_after();
        const uint32_t tp = tmem_base + TMEM_P + ps * P_COLS;
        if (ptx::elect_one()) {
#pragma unroll
          for (int k = 0; k < ((c * 64 + 64 <= NPAD) ? 4 : 2); ++k)
            // 16 keys per MMA: A (P in TMEM) advances 8 cells, B (MN-major V) advances 16 key rows
            ptx::umma_f16_ts(tmem_o, tp + 8u * k, dv + static_cast<uint64_t>(k * (16 * ROWB / 16)), idesc_pv, (c | k) != 0 ? 1u : 0u);
          ptx::umma_commit(&bar.p_empty[ps]);
          if (c == NCH - 1) {
            ptx::umma_commit(&bar.o_full[ob]);
            ptx::umma_commit(&v_empty[vb]);
          }
        }
        __syncwarp();
        dv += 4 * (16 * ROWB / 16);
      }
      if (lane == 0) CFM_PROF(13);
    }
  } else if (warp == SM_WARPS + 2) {
    // ===================== Q K^T issuer =====================
    // S is produced and released in two halves: columns [0,128) of item i+1 are written as soon as every softmax warp has
    // read them for the last time in its second pass over item i, columns [128,288) after the last read of S.  The softmax
    // warps therefore find the next item's scores ready when they get there.
    constexpr uint32_t idesc_qa = ptx::make_idesc_f16(128, N_A), idesc_qb = ptx::make_idesc_f16(128, N_B);   // A, B K-major
    const uint64_t dq0 = ptx::make_smem_desc_sw128(ptx::smem_u32(sQ));
    const uint64_t dk = ptx::make_smem_desc_sw128(ptx::smem_u32(sK));
    const uint32_t my_items = slot < n_items ? (n_items - slot + nslots - 1) / nslots : 0u;
    if (lane == 0) ptx::fence_proxy_async();                 // the zero fill of the pad rows -> async proxy (tensor core reads)
    __syncwarp();
    for (uint32_t it = 0; it < my_items; ++it) {
      const uint64_t dq = dq0 + static_cast<uint64_t>((it & 1u) * (Q_BYTES / 16));
      ptx::mbar_wait(k_full, it & 1u);
      if (lane == 0) CFM_PROF(3);
      ptx::mbar_wait(&q_full[it & 1u], (it >> 1) & 1u);
      if (lane == 0) CFM_PROF(4);
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        if (it > 0) ptx::mbar_wait(&bar.s_empty[half], (it - 1) & 1u);
        ptx::tc_fence_after();
        if (half == 1 && lane == 0) CFM_PROF(5);
        if (ptx::elect_one()) {
#pragma unroll
          for (int k = 0; k < 4; ++k)
            ptx::umma_f16(tmem_base + half * N_A, dq + 2u * k, dk + static_cast<uint64_t>(half * N_A * (ROWB / 16)) + 2u * k,
                          half ? idesc_qb : idesc_qa, k != 0 ? 1u : 0u);
          ptx::umma_commit(&bar.s_full[half]);
          if (half == 1) ptx::umma_commit(kq_empty);
        }
        __syncwarp();
      }
    }
  } else {
    ptx::cp_async_wait_all();                                // bias slice (this thread's pieces)
    asm volatile("bar.sync 5, 512;" ::: "memory");           // ... and everybody else's
    softmax_role(p, sBias, sMask, sX, sO, bar, tmem_base, hp, slot, nslots);
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0 && p.prof != nullptr) {
    long long gt;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
    p.prof[(blockIdx.x * 8 + 7) * 32 + 27] = gt;
    p.prof[(blockIdx.x * 8 + 7) * 32 + 31] = clock64();
  }
  if (warp == SM_WARPS + 1) ptx::tmem_dealloc(tmem_base, 512);
}

int launch_cfm(const CUtensorMap* tm, const CfmParams& p, cudaStream_t st) {
  if (const int rc = set_dyn_smem<cfm_attention_tc_kernel>(CFM_SMEM, "cfm_attention")) return rc;
  // persistent CTAs, each bound to one head pair (its bias slice stays in shared memory) and walking over (clip, window) items
  const int n_items = p.B * p.nWh * p.nWw;
  int nslots = num_sms() / 4;
  if (nslots > n_items) nslots = n_items;
  if (nslots < 1) nslots = 1;
  launch_k(cfm_attention_tc_kernel, nslots * 4, CFM_THREADS, CFM_SMEM, st, tm[0], tm[1], tm[2], tm[3], tm[4], p);
  return launch_status("cfm_attention_tc_kernel");
}

// Where the pooled K/V of a launch live.  kv_tgt: pooled target level, [B, nW, 512] with clip stride tgt_stride (elements).
// kv_ref[k]: base of the level maps of reference role k: the map of slot s of rank r starts s * slot_stride[k] +
// r * rank_stride elements further.  ref_slot: device table [B, 3, 2] naming (slot, rank) of every clip's reference frame k,
// or null ((clip index, 0)).
struct CfmKv {
  const void* kv_tgt; int64_t tgt_stride;
  const void* kv_ref[3]; int64_t slot_stride[3]; int64_t rank_stride;
  const int32_t* ref_slot; int n_slots[3]; int n_ranks;
};

int cfm_run(const void* qkv_a, const CfmKv& kv, const void* bias_tab, void* out, void* dump, int B, int H, int W, int C,
            int heads, float scale, void* stream, long long* prof = nullptr) {
  CFFM_REQUIRE(qkv_a && kv.kv_tgt && kv.kv_ref[0] && kv.kv_ref[1] && kv.kv_ref[2] && bias_tab && out, CFFM_E_BADARG,
               "cfm_attention: null pointer");
  CFFM_REQUIRE(B > 0 && H > 0 && W > 0 && scale > 0.f && kv.n_ranks > 0 && kv.n_slots[0] > 0 && kv.n_slots[1] > 0 && kv.n_slots[2] > 0,
               CFFM_E_BADARG, "cfm_attention: non-positive size or scale");
  CFFM_REQUIRE(C == 256 && heads == 8, CFFM_E_UNSUPPORTED,
               "cfm_attention: built for C=256, heads=8 (cffm_head.py:74-95), got C=%d heads=%d", C, heads);
  CFFM_REQUIRE(aligned16(qkv_a) && aligned16(kv.kv_tgt) && aligned16(kv.kv_ref[0]) && aligned16(kv.kv_ref[1]) && aligned16(kv.kv_ref[2]) &&
                   aligned16(bias_tab) && aligned16(out) && kv.tgt_stride % 8 == 0 && kv.rank_stride % 8 == 0 &&
                   kv.slot_stride[0] % 8 == 0 && kv.slot_stride[1] % 8 == 0 && kv.slot_stride[2] % 8 == 0,
               CFFM_E_BADARG, "cfm_attention: pointers and strides must be 16-byte aligned");
  const int nWh = (H + WS - 1) / WS, nWw = (W + WS - 1) / WS;
  const int64_t Ha = nWh * WS + 2 * RING, Wa = nWw * WS + 2 * RING;
  CUtensorMap tm[5];
  {
    const int64_t dims[4] = {CQKV, Wa, Ha, B}, strides[3] = {CQKV, Wa * CQKV, Ha * Wa * CQKV};
    const int box[4] = {CPAIR, HALO, HALO, 1};
    int rc = make_tmap_4d(&tm[0], qkv_a, dims, strides, box);
    if (rc) return rc;
  }
  const int lev_f[4] = {1, 1, 2, 3}, lev_k[4] = {5, 7, 5, 3};
  for (int l = 0; l < 4; ++l) {
    const int64_t gw = static_cast<int64_t>(lev_f[l]) * nWw, gh = static_cast<int64_t>(lev_f[l]) * nWh;
    int rc;
    if (l == 0) {
      const int64_t dims[4] = {CKV, gw, gh, B}, strides[3] = {CKV, gw * CKV, kv.tgt_stride};
      const int box[4] = {CPAIR, lev_k[l], lev_k[l], 1};
      rc = make_tmap_nd(&tm[1], kv.kv_tgt, 4, dims, strides, box);
    } else {
      // a stride of a dimension of extent 1 is never used, but the encoder still wants a 16-byte multiple > 0
      const int64_t rs = kv.rank_stride > 0 ? kv.rank_stride : kv.slot_stride[l - 1];
      const int64_t dims[5] = {CKV, gw, gh, kv.n_slots[l - 1], kv.n_ranks}, strides[4] = {CKV, gw * CKV, kv.slot_stride[l - 1], rs};
      const int box[5] = {CPAIR, lev_k[l], lev_k[l], 1, 1};
      rc = make_tmap_nd(&tm[1 + l], kv.kv_ref[l - 1], 5, dims, strides, box);
    }
    if (rc) return rc;
  }
  CfmParams p;
  p.qkv_a = static_cast<const __half*>(qkv_a);
  p.bias_tab = static_cast<const __half*>(bias_tab);
  p.out = static_cast<__half*>(out);
  p.dump = static_cast<__half*>(dump);
  p.prof = prof;
  p.ref_slot = kv.ref_slot;
  p.B = B; p.H = H; p.W = W; p.nWh = nWh; p.nWw = nWw;
  p.scale_log2e = scale * LOG2E;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  return launch_cfm(tm, p, st);
}

// kv_pooled fp16 [B, 15 nW, 512]: the four level maps of a clip back to back (target | ref 0 | ref 1 | ref 2)
CfmKv packed_kv(const void* kv_pooled, int B, int H, int W) {
  const int64_t nW = static_cast<int64_t>((H + WS - 1) / WS) * ((W + WS - 1) / WS);
  const __half* base = static_cast<const __half*>(kv_pooled);
  CfmKv kv;
  kv.kv_tgt = base; kv.tgt_stride = 15 * nW * CKV;
  kv.kv_ref[0] = base ? base + nW * CKV : nullptr; kv.kv_ref[1] = base ? base + 2 * nW * CKV : nullptr;
  kv.kv_ref[2] = base ? base + 6 * nW * CKV : nullptr;
  for (int k = 0; k < 3; ++k) { kv.slot_stride[k] = 15 * nW * CKV; kv.n_slots[k] = B; }
  kv.rank_stride = 0; kv.n_ranks = 1;
  kv.ref_slot = nullptr;
  return kv;
}

}  // namespace
}  // namespace cffm

extern "C" int cffm_cfm_attention(const void* qkv_a, const void* kv_pooled, const void* bias_tab, void* out, int B, int H,
                                  int W, int C, int heads, float scale, void* stream) {
  return cffm::cfm_run(qkv_a, cffm::packed_kv(kv_pooled, B, H, W), bias_tab, out, nullptr, B, H, W, C, heads, scale, stream);
}

extern "C" int cffm_cfm_attention_slots(const void* qkv_a, const void* kv_tgt, int64_t tgt_stride, const void* kv_ref0,
                                        const void* kv_ref1, const void* kv_ref2, int64_t slot_stride0, int64_t slot_stride1,
                                        int64_t slot_stride2, int n_slots0, int n_slots1, int n_slots2, int64_t rank_stride,
                                        int n_ranks, const int32_t* ref_slot, const void* bias_tab, void* out, int B, int H,
                                        int W, int C, int heads, float scale, void* stream) {
  cffm::CfmKv kv;
  kv.kv_tgt = kv_tgt; kv.tgt_stride = tgt_stride;
  kv.kv_ref[0] = kv_ref0; kv.kv_ref[1] = kv_ref1; kv.kv_ref[2] = kv_ref2;
  kv.slot_stride[0] = slot_stride0; kv.slot_stride[1] = slot_stride1; kv.slot_stride[2] = slot_stride2;
  kv.n_slots[0] = n_slots0; kv.n_slots[1] = n_slots1; kv.n_slots[2] = n_slots2;
  kv.rank_stride = rank_stride; kv.n_ranks = n_ranks;
  kv.ref_slot = ref_slot;
  return cffm::cfm_run(qkv_a, kv, bias_tab, out, nullptr, B, H, W, C, heads, scale, stream);
}

extern "C" int cffm_cfm_attention_dump(const void* qkv_a, const void* kv_pooled, const void* bias_tab, void* out, void* dump,
                                       int B, int H, int W, int C, int heads, float scale, void* stream) {
  CFFM_REQUIRE(dump != nullptr, CFFM_E_BADARG, "cfm_attention_dump: null dump buffer");
  return cffm::cfm_run(qkv_a, cffm::packed_kv(kv_pooled, B, H, W), bias_tab, out, dump, B, H, W, C, heads, scale, stream);
}

extern "C" int cffm_cfm_layout(int32_t* out8) {
  using namespace cffm;
  CFFM_REQUIRE(out8 != nullptr, CFFM_E_BADARG, "cfm_layout: null pointer");
  out8[0] = 0; out8[1] = R1; out8[2] = R2; out8[3] = R3; out8[4] = R4; out8[5] = NPAD; out8[6] = BPITCH;
  out8[7] = RING;
  return CFFM_OK;
}

/* Bring-up hook (not in the header): SM-clock stamps of the pipeline events of the first 8 items of every CTA. */
extern "C" int cffm_cfm_attention_prof(const void* qkv_a, const void* kv_pooled, const void* bias_tab, void* out, void* prof,
                                       int B, int H, int W, int C, int heads, float scale, void* stream) {
  return cffm::cfm_run(qkv_a, cffm::packed_kv(kv_pooled, B, H, W), bias_tab, out, nullptr, B, H, W, C, heads, scale, stream,
                       static_cast<long long*>(prof));
}
