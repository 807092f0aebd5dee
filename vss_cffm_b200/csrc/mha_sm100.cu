// softmax(scale q k^T) v on the 5th-generation tensor cores: the MiT efficient attention (mix_transformer.py:96-117) for
// head_dim 64 and 32 and up to 512 keys (N_kv = 225 at every stage for 480 x 480 input, 405 for 480 x 853, 512 for 512 x 1024).
// Same machine model as the CFM kernel (cfm_sm100.cu): P never touches shared memory, no proxy fences in the item loop.
//
// Work item = (frame b, head unit u, query tile qt); every persistent CTA owns a CONTIGUOUS range of items, so K and V of a
// (frame, unit) are loaded once and stay in shared memory while the CTA walks the query tiles (double-buffered across units).
//   TMA      K, V [<= 512 keys x 64 channels] and the Q tile [128 x 64], fp16, 128-byte swizzle
//   tcgen05  S[128 x 256] = Q K^T in TMEM, produced and released in two halves of 128 columns
//   softmax  16 warps, thread = (row, 16-column quarter of every 64-key chunk): exact row maximum (pass 1), then
//            p = 2^((s - m) scale log2 e), row sum, fp16 P chunks written back into TMEM with tcgen05.st (pass 2)
//   tcgen05  O[128 x 64] += P_chunk V_chunk, A = P from TMEM (TS-mode MMA), B = V MN-major as loaded (no transpose)
//   epilogue O / rowsum -> fp16 -> global
// More than 256 keys: S does not fit twice over, so an item takes FOUR rounds over the same S columns -- Q K^T of key block 0
// and of block 1 for the row maximum, then both again for the exponentials (the tensor core recomputes S, 4 MMAs per block,
// instead of a rescaling online softmax: the result is the exact two-pass softmax for every N_kv).
// head_dim 32: a unit is a PAIR of heads.  The 128 rows of the tile are 64 queries of head 2u (channels 0..31 of the pair,
// zeros in 32..63) and the same 64 queries of head 2u+1 (zeros | channels 32..63): one M = 128, K = 64 MMA against the
// pair's 64 key channels computes both heads (block-diagonal Q, written by the producer warp with cp.async); O is read back
// from the two diagonal 32-column blocks.  With a single head (MiT-B0 stage 1) the 64-channel boxes simply run past the
// tensor and the TMA's zero fill supplies channels 32..63.
// Keys >= N_kv inside a box (rows of the next frame, or zero fill past the end) are masked before both passes.
// Warp roles: 0..15 softmax / epilogue, 16 producer, 17 TMEM allocator + P V issuer, 18 Q K^T issuer.
#include <cuda.h>

#include <type_traits>

#include "common.cuh"
#include "ptx_sm100.cuh"

namespace cffm {
namespace {

constexpr int QT = 128, ROWB = 128, BLK = 256, KV_CAP = 2 * BLK;         // query rows, bytes per row, keys per S round, key limit
constexpr int Q_BYTES = QT * ROWB, BLK_BYTES = BLK * ROWB;
constexpr int SM_WARPS = 16, MHA_THREADS = (SM_WARPS + 3) * 32;
constexpr int P_RING = 4, P_COLS = 32;                                    // P chunks (64 keys = 32 packed cells) in TMEM
constexpr int TMEM_O = BLK, TMEM_P = TMEM_O + 2 * 64;                     // S 256 | O 2 x 64 | P ring 4 x 32 = 512 columns
constexpr int X_FLOATS = 2 * 4 * QT;                                      // [max | sum][column quarter][row]
constexpr int SMEM_MHA = 4 * BLK_BYTES + 3 * Q_BYTES /*2 Q tiles + the output staging tile*/ + 2 * X_FLOATS * 4 + 256 /*barriers*/ +
                         1024 /*align slack*/;
constexpr float LOG2E = 1.4426950408889634f;
#define MHA_PROF(slot) do { if (p.prof != nullptr && it < 8) p.prof[(blockIdx.x * 8 + it) * 32 + (slot)] = clock64(); } while (0)
static_assert(TMEM_P + P_RING * P_COLS <= 512, "TMEM budget");
static_assert(SMEM_MHA <= 232448, "shared memory budget of one CTA per SM");

struct MhaParams {
  const __half* q;          // pair mode: Q is gathered by the producer warp (cp.async)
  int64_t ldq;
  __half* out;
  int64_t ldo;
  int batch, Nq, Nkv, heads, dh;
  int units;                // heads, or head pairs (head_dim 32, more than one head)
  int qtiles;               // query tiles per (frame, unit): 128 queries, or 64 in pair mode
  int n_items, per_cta, extra;   // items, and their split over the grid
  int nb;                   // key blocks of 256: 1 or 2
  int kv_box_bytes;         // bytes of one K (or V) TMA box
  float scale_log2e;
  long long* prof;          // bring-up hook: per CTA / item / event SM clock stamps [grid, 8, 32] (or null)
};

struct MhaBars {
  uint64_t *s_full, *s_empty, *p_full, *p_empty, *o_full, *o_empty;
};

// Walks the CTA's contiguous item range without divisions (every role steps the same cursor)
struct ItemCursor {
  int b, u, qt;
  __device__ __forceinline__ ItemCursor(const MhaParams& p, int item) {
    const int bu = item / p.qtiles;
    qt = item - bu * p.qtiles;
    b = bu / p.units;
    u = bu - b * p.units;
  }
  __device__ __forceinline__ bool next(const MhaParams& p) {   // true when the (frame, unit) changes
    if (++qt < p.qtiles) return false;
    qt = 0;
    if (++u == p.units) { u = 0; ++b; }
    return true;
  }
};

// ---- softmax / epilogue role: one instruction stream for all 16 warps (cq is a run-time, warp-uniform offset)
// Pass 1 and pass 2 partition the columns of S differently.  The row maximum needs no particular order, so in pass 1 warp
// cq takes the 64-key chunk cq WHOLE, as two 32-column loads: few long TMEM reads in flight on all four warps of a lane
// quarter (a 16-column load per 64-key step left the pass latency-bound: 1100 cycles against ~350 of TMEM bandwidth).
// Pass 2 walks the chunks in P V order, every warp owning 16 columns of each, one load ahead of the exponentials.
template <int OC, bool PAIR>
__device__ __forceinline__ void mha_softmax_role(const MhaParams& p, float* sX, uint8_t* sO, const MhaBars& bar,
                                                 uint32_t tmem_base, int item0, int item1) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int wq = warp & 3, cq = warp >> 2;                   // TMEM lane quarter, column quarter of every chunk
  const int row = wq * 32 + lane;                            // row of S = TMEM lane
  const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(wq * 32) << 16);
  const float scl = p.scale_log2e;
  const int bar_rows = 1 + wq;                               // named barrier of the four warps that share these 32 rows
  const int rounds = p.nb == 1 ? 1 : 4;

  uint32_t it = 0, rc = 0, gc = 0;                           // items, S rounds, P chunks so far
  ItemCursor cur(p, item0);
  for (int item = item0; item < item1; ++item, ++it, cur.next(p)) {
    const int qt = cur.qt, b = cur.b, u = cur.u;
    float* xm = sX + (it & 1u) * X_FLOATS;
    uint32_t sb[32];                                         // pass 2: chunk C lives in sb[16 (C & 1) ..]; pass 1: one 32-column load
    float mx = -INFINITY, msc = 0.f, sum0 = 0.f, sum1 = 0.f;
    if (threadIdx.x == 0) MHA_PROF(8);
#pragma unroll 1
    for (int r = 0; r < rounds; ++r, ++rc) {
      const int blk = p.nb == 1 ? 0 : (r & 1);
      const int kb = min(BLK, p.Nkv - blk * BLK);            // keys of this block
      const int nch = (kb + 63) >> 6;                        // 64-key chunks: 1..4
      const bool do_max = p.nb == 1 || r < 2, do_exp = p.nb == 1 || r >= 2;
      const int rel0 = nch >= 2 ? 1 : 0;                     // last chunk that reads the first half of S
      bool h0 = false, h1 = false;                           // halves of S known to be complete
      auto need_half = [&](int h) {
        bool& have = h ? h1 : h0;
        if (!have) {
          ptx::mbar_wait(&bar.s_full[h], rc & 1u);
          ptx::tc_fence_after();
          have = true;
        }
      };

      if (do_max) {
        // ---- pass 1: row maximum of the raw scores (scale > 0)
        if (cq < nch) {
          need_half(cq >> 1);
          if (threadIdx.x == 0 && r == 0) MHA_PROF(9);
#pragma unroll
          for (int hf = 0; hf < 2; ++hf) {
            const int col0 = cq * 64 + hf * 32;
            if (col0 < kb) {
              ptx::tmem_ld_32x32b_x32(lane_addr + col0, sb);
              ptx::tmem_ld_wait();
              float m0, m1;
              if (col0 + 32 <= kb) {
                m0 = fmaxf(__uint_as_float(sb[0]), __uint_as_float(sb[1]));
                m1 = fmaxf(__uint_as_float(sb[2]), __uint_as_float(sb[3]));
#pragma unroll
                for (int j = 4; j < 32; j += 4) {
                  m0 = fmaxf(m0, fmaxf(__uint_as_float(sb[j]), __uint_as_float(sb[j + 1])));
                  m1 = fmaxf(m1, fmaxf(__uint_as_float(sb[j + 2]), __uint_as_float(sb[j + 3])));
                }
              } else {                                       // the piece that straddles N_kv
                m0 = m1 = -INFINITY;
#pragma unroll
                for (int j = 0; j < 32; ++j)
                  if (col0 + j < kb) m0 = fmaxf(m0, __uint_as_float(sb[j]));
              }
              mx = fmaxf(mx, fmaxf(m0, m1));
            }
          }
        }
        if (!do_exp) {                                       // maximum-only round: S goes straight back to the Q K^T issuer
          // (a warp may only arrive for a round whose S it has seen complete: otherwise a warp without columns in two
          // consecutive rounds would arrive twice in one phase and free S under the warps still reading it)
          need_half(0);
          need_half(1);
          ptx::tc_fence_before();
          __syncwarp();
          if (lane == 0) { ptx::mbar_arrive(&bar.s_empty[0]); ptx::mbar_arrive(&bar.s_empty[1]); }
        }
      }
      if (do_exp) {                                          // first chunk of pass 2: in flight across the exchange below
        need_half(0);
        ptx::tmem_ld_32x32b_x16(lane_addr + cq * 16, sb);
      }
      if (do_max && (p.nb == 1 || r == 1)) {
        xm[cq * QT + row] = mx;
        asm volatile("bar.sync %0, 128;" ::"r"(bar_rows) : "memory");
        mx = fmaxf(fmaxf(xm[row], xm[QT + row]), fmaxf(xm[2 * QT + row], xm[3 * QT + row]));   // column 0 is a key: finite
        msc = mx * scl;
        if (threadIdx.x == 0) MHA_PROF(10);
      }
      if (do_exp) {
        // ---- pass 2: p = 2^((s - m) scale log2 e), row sum, packed fp16 P chunks into tensor memory
        auto pass2 = [&](auto cc) {
          constexpr int C = decltype(cc)::value;
          if (C < nch) {
            const uint32_t ps = gc % P_RING, pph = ((gc / P_RING) & 1u) ^ 1u;
            ptx::tmem_ld_wait();                             // chunk C has landed
            if (C + 1 < nch) {
              if (C + 1 == 2) need_half(1);
              ptx::tmem_ld_32x32b_x16(lane_addr + (C + 1) * 64 + cq * 16, sb + 16 * ((C + 1) & 1));
            }
            if (C == rel0 || C == nch - 1) {                 // chunk C is in registers: hand finished halves of S back
              ptx::tc_fence_before();
              __syncwarp();
              if (lane == 0) {
                if (C == rel0) ptx::mbar_arrive(&bar.s_empty[0]);
                if (C == nch - 1) ptx::mbar_arrive(&bar.s_empty[1]);
              }
            }
            // the ring slot is almost always free already; the probe's latency hides behind the math
            const bool slot_free = ptx::mbar_test_wait(&bar.p_empty[ps], pph);
            const uint32_t* s = sb + 16 * (C & 1);
            const int col0 = C * 64 + cq * 16;
            uint32_t hh[8];
            if (col0 + 16 <= kb) {
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const float p0 = ptx::ex2_approx(fmaf(__uint_as_float(s[2 * j]), scl, -msc));
                const float p1 = ptx::ex2_approx(fmaf(__uint_as_float(s[2 * j + 1]), scl, -msc));
                sum0 += p0; sum1 += p1;
                hh[j] = pack_half2(p0, p1);
              }
            } else {                                         // the piece that straddles N_kv (or lies beyond it): key pairs past
#pragma unroll                                               // the end cost no exponentials (warp-uniform branches)
              for (int j = 0; j < 8; ++j) {
                hh[j] = 0u;
                if (col0 + 2 * j < kb) {
                  const float p0 = ptx::ex2_approx(fmaf(__uint_as_float(s[2 * j]), scl, -msc));
                  const float p1 = col0 + 2 * j + 1 < kb ? ptx::ex2_approx(fmaf(__uint_as_float(s[2 * j + 1]), scl, -msc)) : 0.f;
                  sum0 += p0; sum1 += p1;
                  hh[j] = pack_half2(p0, p1);
                }
              }
            }
            if (!slot_free) ptx::mbar_wait(&bar.p_empty[ps], pph);
            ptx::tc_fence_after();
            ptx::tmem_st_32x32b_x8(lane_addr + TMEM_P + ps * P_COLS + cq * 8, hh);
            ptx::tmem_st_wait();
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(&bar.p_full[ps]);
            ++gc;
          }
        };
        pass2(std::integral_constant<int, 0>{}); pass2(std::integral_constant<int, 1>{});
        pass2(std::integral_constant<int, 2>{}); pass2(std::integral_constant<int, 3>{});
      }
    }
    if (threadIdx.x == 0) MHA_PROF(11);
    xm[4 * QT + cq * QT + row] = sum0 + sum1;
    asm volatile("bar.sync %0, 128;" ::"r"(bar_rows) : "memory");
    const float inv = 1.f / ((xm[4 * QT + row] + xm[5 * QT + row]) + (xm[6 * QT + row] + xm[7 * QT + row]));   // same order in all four threads

    // ---- epilogue: this thread's OC columns of O / rowsum -> fp16 -> global
    const uint32_t ob = it & 1u;
    if (threadIdx.x == 0) MHA_PROF(12);
    ptx::mbar_wait(&bar.o_full[ob], (it >> 1) & 1u);
    ptx::tc_fence_after();
    if (threadIdx.x == 0) MHA_PROF(13);
    const int hl = PAIR ? row >> 6 : 0;
    uint32_t o[OC];
    const uint32_t oaddr = lane_addr + TMEM_O + ob * 64 + hl * 32 + cq * OC;
    if (OC == 16) ptx::tmem_ld_32x32b_x16(oaddr, o);
    else ptx::tmem_ld_32x32b_x8(oaddr, o);
    ptx::tmem_ld_wait();
    ptx::tc_fence_before();
    __syncwarp();
    if (lane == 0) ptx::mbar_arrive(&bar.o_empty[ob]);
    uint4 w[OC / 8];
#pragma unroll
    for (int g = 0; g < OC / 8; ++g) {
      w[g].x = pack_half2(__uint_as_float(o[g * 8]) * inv, __uint_as_float(o[g * 8 + 1]) * inv);
      w[g].y = pack_half2(__uint_as_float(o[g * 8 + 2]) * inv, __uint_as_float(o[g * 8 + 3]) * inv);
      w[g].z = pack_half2(__uint_as_float(o[g * 8 + 4]) * inv, __uint_as_float(o[g * 8 + 5]) * inv);
      w[g].w = pack_half2(__uint_as_float(o[g * 8 + 6]) * inv, __uint_as_float(o[g * 8 + 7]) * inv);
    }
    if constexpr (OC == 16) {
      // A thread-per-row store sends 32 separate 16-byte pieces per instruction through the L1 (one wavefront each: ~1000
      // cycles per tile).  The rows go through a swizzled staging tile instead (16-byte piece j of row r at j ^ (r & 7): the
      // eight lanes of a store phase hit eight distinct bank groups) and leave as whole 128-byte rows, four per instruction.
      uint8_t* srow = sO + row * ROWB;
#pragma unroll
      for (int g = 0; g < 2; ++g) *reinterpret_cast<uint4*>(srow + (((2 * cq + g) ^ (row & 7)) << 4)) = w[g];
      asm volatile("bar.sync %0, 128;" ::"r"(bar_rows) : "memory");
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int pi = cq * 32 + lane + 128 * i, rl = pi >> 3, j = pi & 7, r2 = wq * 32 + rl;
        const int qrow = qt * QT + r2;
        const uint4 val = *reinterpret_cast<const uint4*>(sO + r2 * ROWB + ((j ^ (rl & 7)) << 4));
        if (qrow < p.Nq)
          *reinterpret_cast<uint4*>(p.out + (static_cast<int64_t>(b) * p.Nq + qrow) * p.ldo + u * 64 + j * 8) = val;
      }
    } else {
      const int qrow = PAIR ? qt * 64 + (row & 63) : qt * QT + row;
      const int head = PAIR ? 2 * u + hl : u;
      if (qrow < p.Nq && head < p.heads)
        *reinterpret_cast<uint4*>(p.out + (static_cast<int64_t>(b) * p.Nq + qrow) * p.ldo + head * p.dh + cq * OC) = w[0];
    }
    if (threadIdx.x == 0) MHA_PROF(14);
  }
}

template <int OC, bool PAIR>
__global__ void __launch_bounds__(MHA_THREADS, 1)
mha_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
              const __grid_constant__ CUtensorMap tmV, const MhaParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = ptx::smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);
  uint8_t* sK = smem;                                          // 2 blocks: two generations of <= 256 keys, or one of <= 512
  uint8_t* sV = sK + 2 * BLK_BYTES;
  uint8_t* sQ = sV + 2 * BLK_BYTES;                            // 2 buffers
  uint8_t* sO = sQ + 2 * Q_BYTES;                              // output staging tile [128 rows][128 B]
  float* sX = reinterpret_cast<float*>(sO + Q_BYTES);          // row max / sum exchange, two item parities
  uint64_t* bars = reinterpret_cast<uint64_t*>(sX + 2 * X_FLOATS);
  uint64_t* kv_full = bars + 0;    // [2] TMA -> issuers: K and V of a (frame, unit) landed
  uint64_t* kv_empty = bars + 2;   // [2] P V issuer -> producer: last P V of the generation retired
  uint64_t* q_full = bars + 4;     // [2] producer -> Q K^T issuer: Q tile landed
  uint64_t* q_empty = bars + 6;    // [2] Q K^T issuer -> producer: last Q K^T of the item retired
  MhaBars bar;
  bar.s_full = bars + 8;           // [2] Q K^T issuer -> softmax: half of S complete
  bar.s_empty = bars + 10;         // [2] softmax (16 warps) -> Q K^T issuer: half of S read for the last time in this round
  bar.p_full = bars + 12;          // [P_RING] softmax (16 warps) -> P V issuer: P chunk written
  bar.p_empty = bars + 12 + P_RING;        // [P_RING] P V issuer -> softmax: P chunk consumed
  bar.o_full = bars + 12 + 2 * P_RING;     // [2] P V issuer -> epilogue: O complete
  bar.o_empty = bars + 14 + 2 * P_RING;    // [2] epilogue (16 warps) -> P V issuer: O read out
  uint32_t* tmem_base_smem = reinterpret_cast<uint32_t*>(bars + 16 + 2 * P_RING);
  static_assert(16 + 2 * P_RING < 32, "barrier area");

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // contiguous item range of this CTA: per_cta items each, the first `extra` CTAs one more
  const int bid = static_cast<int>(blockIdx.x);
  const int item0 = bid * p.per_cta + min(bid, p.extra), item1 = item0 + p.per_cta + (bid < p.extra ? 1 : 0);
  const int ngen = p.nb == 1 ? 2 : 1;                          // K/V generations resident at once
  const int rounds = p.nb == 1 ? 1 : 4;
  if (threadIdx.x == 0 && p.prof != nullptr) {
    long long gt;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
    p.prof[(blockIdx.x * 8 + 7) * 32 + 26] = gt;
    p.prof[(blockIdx.x * 8 + 7) * 32 + 28] = clock64();
  }

  if (warp == SM_WARPS && lane == 0) {
    ptx::prefetch_tensormap(&tmQ); ptx::prefetch_tensormap(&tmK); ptx::prefetch_tensormap(&tmV);
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(&kv_full[i], 1); ptx::mbar_init(&kv_empty[i], 1);
      ptx::mbar_init(&q_full[i], 1); ptx::mbar_init(&q_empty[i], 1);
      ptx::mbar_init(&bar.s_full[i], 1); ptx::mbar_init(&bar.s_empty[i], SM_WARPS);
      ptx::mbar_init(&bar.o_full[i], 1); ptx::mbar_init(&bar.o_empty[i], SM_WARPS);
    }
    for (int i = 0; i < P_RING; ++i) { ptx::mbar_init(&bar.p_full[i], SM_WARPS); ptx::mbar_init(&bar.p_empty[i], 1); }
    ptx::fence_barrier_init();
  }
  if (warp == SM_WARPS + 1) {
    ptx::tmem_alloc(tmem_base_smem, 512);
    ptx::tmem_relinquish();
  }
  if (PAIR) {
    // the block-diagonal Q tiles start as zeros: the other head's channels and the rows past N_q are never written
    for (int i = threadIdx.x; i < 2 * Q_BYTES / 16; i += MHA_THREADS) reinterpret_cast<uint4*>(sQ)[i] = make_uint4(0u, 0u, 0u, 0u);
    ptx::fence_proxy_async();                                  // generic-proxy writes -> visible to the tensor core (async proxy)
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_base_smem;
  if (threadIdx.x == 0 && p.prof != nullptr) p.prof[(blockIdx.x * 8 + 7) * 32 + 29] = clock64();
  pdl_sync();                                                  // prologue above overlaps the previous kernel
  if (threadIdx.x == 0 && p.prof != nullptr) p.prof[(blockIdx.x * 8 + 7) * 32 + 30] = clock64();

  // Producer and MMA roles: the whole warp walks the loop and waits on the barriers, one ELECTED lane issues the TMA /
  // tcgen05 instructions (see gemm_sm100.cu).
  if (warp == SM_WARPS) {
    // ===================== producer =====================
    int gen = -1;
    uint32_t it = 0;
    ItemCursor cur(p, item0);
    bool fresh = true;                                         // the item opens a new (frame, unit)
    for (int item = item0; item < item1; ++item, ++it, fresh = cur.next(p)) {
      const int qt = cur.qt, b = cur.b, u = cur.u;
      if (fresh) {                                             // K and V of the next (frame, unit)
        ++gen;
        const uint32_t gb = gen % ngen, gph = (gen / ngen) & 1u;
        ptx::mbar_wait(&kv_empty[gb], gph ^ 1u);
        if (lane == 0) MHA_PROF(0);
        if (ptx::elect_one()) {
          ptx::mbar_arrive_expect_tx(&kv_full[gb], 2 * p.nb * p.kv_box_bytes);
          for (int blk = 0; blk < p.nb; ++blk) {
            ptx::tma_load_2d(sK + (gb + blk) * BLK_BYTES, &tmK, &kv_full[gb], u * 64, b * p.Nkv + blk * BLK);
            ptx::tma_load_2d(sV + (gb + blk) * BLK_BYTES, &tmV, &kv_full[gb], u * 64, b * p.Nkv + blk * BLK);
          }
        }
        __syncwarp();
      }
      const uint32_t qb = it & 1u;
      ptx::mbar_wait(&q_empty[qb], ((it >> 1) & 1u) ^ 1u);
      if (lane == 0) MHA_PROF(1);
      if (!PAIR) {
        if (ptx::elect_one()) {
          ptx::mbar_arrive_expect_tx(&q_full[qb], Q_BYTES);
          ptx::tma_load_2d(sQ + qb * Q_BYTES, &tmQ, &q_full[qb], u * 64, b * p.Nq + qt * QT);
        }
        __syncwarp();
      } else {
        // 128 rows x four 16-byte pieces: rows 0..63 = head 2u in bytes 0..63, rows 64..127 = head 2u+1 in bytes 64..127
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const int i = lane + 32 * j, r = i >> 2, pc = i & 3, hl = r >> 6;
          const int qrow = qt * 64 + (r & 63), head = 2 * u + hl;
          if (qrow < p.Nq && head < p.heads)
            ptx::cp_async16(sQ + qb * Q_BYTES + r * ROWB + (((hl * 4 + pc) ^ (r & 7)) << 4),
                            p.q + (static_cast<int64_t>(b) * p.Nq + qrow) * p.ldq + head * 32 + pc * 8);
        }
        ptx::cp_async_commit();
        ptx::cp_async_wait_all();
        ptx::fence_proxy_async();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&q_full[qb]);
      }
    }
  } else if (warp == SM_WARPS + 1) {
    // ===================== P V issuer =====================
    constexpr uint32_t idesc_pv = ptx::make_idesc_f16(QT, 64) | (1u << 16);            // B (= V) MN-major
    const uint64_t dv0 = ptx::make_smem_desc_sw128(ptx::smem_u32(sV));
    int gen = -1;
    uint32_t it = 0, gc = 0, gb = 0;
    ItemCursor cur(p, item0);
    bool fresh = true;
    for (int item = item0; item < item1; ++item, ++it, fresh = cur.next(p)) {
      if (fresh) {
        ++gen;
        gb = gen % ngen;
        ptx::mbar_wait(&kv_full[gb], (gen / ngen) & 1u);
      }
      const bool last_of_gen = item + 1 == item1 || cur.qt + 1 == p.qtiles;
      const uint32_t ob = it & 1u;
      if (it >= 2) ptx::mbar_wait(&bar.o_empty[ob], ((it - 2) >> 1) & 1u);
      const uint32_t tmem_o = tmem_base + TMEM_O + ob * 64;
      uint32_t first = 1u;
      for (int blk = 0; blk < p.nb; ++blk) {
        const int nch = (min(BLK, p.Nkv - blk * BLK) + 63) >> 6;
        for (int c = 0; c < nch; ++c, ++gc) {
          const uint32_t ps = gc % P_RING;
          ptx::mbar_wait(&bar.p_full[ps], (gc / P_RING) & 1u);
          ptx::tc_fence_after();
          if (lane == 0 && first) MHA_PROF(6);
          const uint32_t tp = tmem_base + TMEM_P + ps * P_COLS;
          const uint64_t dv = dv0 + static_cast<uint64_t>(((gb + blk) * BLK_BYTES + c * 64 * ROWB) / 16);
          const bool last = blk == p.nb - 1 && c == nch - 1;
          if (ptx::elect_one()) {
#pragma unroll
            for (int k = 0; k < 4; ++k)
              // 16 keys per MMA: A (P in TMEM) advances 8 cells, B (MN-major V) advances 16 key rows
              ptx::umma_f16_ts(tmem_o, tp + 8u * k, dv + static_cast<uint64_t>(k * (16 * ROWB / 16)), idesc_pv, (k != 0 || !first) ? 1u : 0u);
            ptx::umma_commit(&bar.p_empty[ps]);
            if (last) {
              ptx::umma_commit(&bar.o_full[ob]);
              if (last_of_gen) ptx::umma_commit(&kv_empty[gb]);
            }
          }
          __syncwarp();
          first = 0u;
        }
      }
      if (lane == 0) MHA_PROF(7);
    }
  } else if (warp == SM_WARPS + 2) {
    // ===================== Q K^T issuer =====================
    constexpr uint32_t idesc_128 = ptx::make_idesc_f16(QT, 128), idesc_64 = ptx::make_idesc_f16(QT, 64);   // A, B K-major
    const uint64_t dq0 = ptx::make_smem_desc_sw128(ptx::smem_u32(sQ));
    const uint64_t dk0 = ptx::make_smem_desc_sw128(ptx::smem_u32(sK));
    int gen = -1;
    uint32_t it = 0, rc = 0, gb = 0;
    ItemCursor cur(p, item0);
    bool fresh = true;
    for (int item = item0; item < item1; ++item, ++it, fresh = cur.next(p)) {
      if (fresh) {
        ++gen;
        gb = gen % ngen;
        ptx::mbar_wait(&kv_full[gb], (gen / ngen) & 1u);
      }
      const uint32_t qb = it & 1u;
      ptx::mbar_wait(&q_full[qb], (it >> 1) & 1u);
      if (lane == 0) MHA_PROF(3);
      const uint64_t dq = dq0 + static_cast<uint64_t>(qb * (Q_BYTES / 16));
      for (int r = 0; r < rounds; ++r, ++rc) {
        const int blk = p.nb == 1 ? 0 : (r & 1);
        const int kpad = ((min(BLK, p.Nkv - blk * BLK) + 63) >> 6) << 6;
        const uint64_t dk = dk0 + static_cast<uint64_t>((gb + blk) * (BLK_BYTES / 16));
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          ptx::mbar_wait(&bar.s_empty[half], (rc & 1u) ^ 1u);
          ptx::tc_fence_after();
          if (lane == 0 && r == 0) MHA_PROF(4 + half);
          const int n = min(128, kpad - 128 * half);           // 128, 64 or <= 0 (nothing in this half)
          if (ptx::elect_one()) {
            if (n > 0) {
#pragma unroll
              for (int k = 0; k < 4; ++k)
                ptx::umma_f16(tmem_base + half * 128, dq + 2u * k, dk + static_cast<uint64_t>(half * 128 * (ROWB / 16)) + 2u * k,
                              n == 128 ? idesc_128 : idesc_64, k != 0 ? 1u : 0u);
            }
            ptx::umma_commit(&bar.s_full[half]);
            if (half == 1 && r == rounds - 1) ptx::umma_commit(&q_empty[qb]);
          }
          __syncwarp();
        }
      }
    }
  } else {
    mha_softmax_role<OC, PAIR>(p, sX, sO, bar, tmem_base, item0, item1);
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

template <int OC, bool PAIR>
int launch_mha(const CUtensorMap& tmQ, const CUtensorMap& tmK, const CUtensorMap& tmV, const MhaParams& p, cudaStream_t st) {
  if (const int rc = set_dyn_smem<mha_tc_kernel<OC, PAIR>>(SMEM_MHA, "mha")) return rc;
  const int grid = p.n_items < num_sms() ? p.n_items : num_sms();
  launch_k(mha_tc_kernel<OC, PAIR>, grid, MHA_THREADS, SMEM_MHA, st, tmQ, tmK, tmV, p);
  return launch_status("mha_tc_kernel");
}

}  // namespace

// Returns CFFM_E_UNSUPPORTED (without setting an error) when the shape is outside this kernel: the caller falls
// back to the mma.sync kernel of attention.cu.
int mha_tcgen05_launch(const void* q, int64_t ldq, const void* k, const void* v, int64_t ldkv, void* out, int64_t ldo,
                       int batch, int Nq, int Nkv, int heads, int head_dim, float scale, cudaStream_t st, long long* prof) {
  if (Nkv > KV_CAP || (head_dim != 64 && head_dim != 32) || ldq % 8 != 0 || ldkv % 8 != 0 || ldo % 8 != 0 || !aligned16(q) ||
      !aligned16(k) || !aligned16(v) || !aligned16(out))
    return CFFM_E_UNSUPPORTED;
  const bool pair = head_dim == 32 && heads > 1;
  MhaParams p;
  p.q = static_cast<const __half*>(q); p.ldq = ldq;
  p.out = static_cast<__half*>(out); p.ldo = ldo;
  p.batch = batch; p.Nq = Nq; p.Nkv = Nkv; p.heads = heads; p.dh = head_dim;
  p.units = pair ? (heads + 1) / 2 : heads;
  p.qtiles = pair ? (Nq + 63) / 64 : (Nq + QT - 1) / QT;
  const int64_t items = static_cast<int64_t>(batch) * p.units * p.qtiles;
  if (items >= (1ll << 31) || static_cast<int64_t>(batch) * (Nq > Nkv ? Nq : Nkv) >= (1ll << 31)) return CFFM_E_UNSUPPORTED;
  p.n_items = static_cast<int>(items);
  {
    const int grid = p.n_items < num_sms() ? p.n_items : num_sms();
    p.per_cta = p.n_items / grid; p.extra = p.n_items % grid;
  }
  p.nb = Nkv > BLK ? 2 : 1;
  const int kv_box = p.nb == 2 ? BLK : (Nkv + 63) / 64 * 64;
  p.kv_box_bytes = kv_box * ROWB;
  p.scale_log2e = scale * LOG2E;
  p.prof = prof;
  const int64_t chans = static_cast<int64_t>(heads) * head_dim;
  CUtensorMap tmQ, tmK, tmV;
  int rc = make_tmap(&tmK, k, static_cast<int64_t>(batch) * Nkv, chans, ldkv, kv_box);
  if (rc) return rc;
  if ((rc = make_tmap(&tmV, v, static_cast<int64_t>(batch) * Nkv, chans, ldkv, kv_box))) return rc;
  if (pair) tmQ = tmK;                                         // unused: Q goes through cp.async
  else if ((rc = make_tmap(&tmQ, q, static_cast<int64_t>(batch) * Nq, chans, ldq, QT))) return rc;
  if (pair) return launch_mha<8, true>(tmQ, tmK, tmV, p, st);
  if (head_dim == 32) return launch_mha<8, false>(tmQ, tmK, tmV, p, st);
  return launch_mha<16, false>(tmQ, tmK, tmV, p, st);
}

}  // namespace cffm

/* Bring-up hook (not in the header): SM-clock stamps of the pipeline events of the first 8 items of every CTA. */
extern "C" int cffm_mha_f16_prof(const void* q, int64_t ldq, const void* k, const void* v, int64_t ldkv, void* out, int64_t ldo,
                                 int batch, int Nq, int Nkv, int heads, int head_dim, float scale, void* prof, void* stream) {
  return cffm::mha_tcgen05_launch(q, ldq, k, v, ldkv, out, ldo, batch, Nq, Nkv, heads, head_dim, scale,
                                  static_cast<cudaStream_t>(stream), static_cast<long long*>(prof));
}
