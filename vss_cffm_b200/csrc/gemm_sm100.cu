// out = act(A . W^T + bias) (+ residual) on the 5th-generation tensor cores.
//
// Persistent, warp-specialised, one CTA per SM looping over 128 x BN output tiles:
//   warp 0      TMA producer   (cp.async.bulk.tensor, 128-byte swizzled [rows][64] fp16 stages)
//   warp 1      TMEM allocator + single-thread tcgen05.mma issuer (2 accumulator stages in TMEM)
//   warps 2..17 epilogue       (tcgen05.ld -> smem transpose -> bias / GELU / ReLU / fp32 residual / LayerNorm
//                               -> row-contiguous global stores), overlapped with the next tile's mainloop
// Barriers: full[s] (TMA -> MMA, transaction bytes) / empty[s] (tcgen05.commit -> TMA),
//           tfull[a] (tcgen05.commit -> epilogue) / tempty[a] (epilogue -> MMA).
// Almost every GEMM of this path has K <= 256, i.e. it is HBM-bound: the design goal is bytes in
// flight (4-6 TMA stages per SM) and fully coalesced epilogue traffic, not MMA issue rate.
// Tails in M, N and K are handled by TMA out-of-bounds zero fill plus predicated stores.
#include <cuda.h>

#include <mutex>
#include <stdlib.h>

#include "common.cuh"
#include "ptx_sm100.cuh"

namespace cffm {
namespace {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;                       // 64 halves = one 128-byte swizzle row
constexpr int UMMA_K = 16;
constexpr int A_STAGE_BYTES = BLOCK_M * BLOCK_K * 2;
constexpr int STG_LD = 36;                        // padded fp32 row of the per-warp 32x32 transpose buffer
constexpr int STG_BYTES = 5120;                   // per epilogue warp, 512-byte aligned: 32x36 fp32 transpose buffer, or a 32 x 64 B
                                                  // swizzled fp16 tile (TMA store source); the last 128 bytes hold a 32-float bias slice

struct Epilogue {
  const float* bias;
  const float* residual;
  int64_t ldr;
  __half* out16;
  int64_t ldo16;
  float* out32;
  int64_t ldo32;
  int act;
  // optional fused LayerNorm of the output row (only when one tile spans the whole row: N <= BN)
  const float* ln_gamma;
  const float* ln_beta;
  float ln_eps;
  __half* ln_out16;
  int64_t ldln;
  // optional SECOND LayerNorm chained on the first (patch-embed norm -> first block's norm1): out32 then receives
  // y = LN(x; ln_gamma, ln_beta) instead of x, and ln_out16 receives LN(y; ln2_gamma, ln2_beta)
  const float* ln2_gamma;
  const float* ln2_beta;
  float ln2_eps;
  // split-K: tile t covers k-blocks [split * kb_per, +kb_per) and writes its partial sums to out32 + split * split_stride
  int splits;
  int kb_per;
  int64_t split_stride;
  // Implicit-GEMM convolution (0 = plain GEMM): A is an NHWC fp16 image read through a 4-D tensor map whose boxes walk the
  // image with the convolution stride (cuTensorMap elementStrides), one box per (filter tap, 64-channel chunk) and M tile.
  // An M tile is conv_by full output rows (conv_by * Wo <= 128 rows of the GEMM) of ONE image; out-of-image taps are the TMA's
  // zero fill, i.e. the convolution's zero padding.  The rows of a tile are contiguous in M, but a tile starts at
  // image * Ho * Wo + tile_y * conv_by * Wo instead of a multiple of 128, and the last tile of an image is short.
  int conv_k;                // filter size (k x k), 0 = plain GEMM
  int conv_stride, conv_pad;
  int conv_cchunks;          // C / 64
  int conv_Wo, conv_Ho, conv_by, conv_tiles_y;
};

// First row and number of valid rows of M tile `mt`
__device__ __forceinline__ void tile_rows(const Epilogue& ep, int mt, int M, int& m0, int& mrows) {
  if (ep.conv_k == 0) {
    m0 = mt * 128;
    mrows = min(128, M - m0);
  } else {
    const int img = mt / ep.conv_tiles_y, ty = mt - img * ep.conv_tiles_y;
    const int per = ep.conv_by * ep.conv_Wo, grp = ep.conv_Ho * ep.conv_Wo;
    m0 = img * grp + ty * per;
    mrows = min(per, grp - ty * per);
  }
}

// EW = number of epilogue warps.  16: one CTA per SM, the throughput configuration (big GEMMs).  8 (BN = 64 only): a CTA
// of 10 warps and <= 112 KB of shared memory, so TWO CTAs share an SM -- the latency configuration for the GEMMs with few
// tiles (stages 3-4, the head): one CTA's TMA round trips hide behind the other's epilogue, and under programmatic
// dependent launch the next GEMM's prologue can start while this one still occupies half of the SM.
template <int BN, int EW, bool F16_ONLY>
struct Cfg {
  static_assert(EW == 16 || (EW == 8 && BN == 64), "8 epilogue warps are built for 64-wide tiles");
  static constexpr int STAGES = EW == 8 ? 3 : ((BN == 64) ? 5 : 4);
  static constexpr int TEAMS = EW / 4 / (BN / 32);  // BN=64, 16 warps: two 8-warp teams alternate tiles (accumulator stage = team)
  static constexpr int TEAM_WARPS = EW / TEAMS;
  static constexpr int THREADS = 64 + EW * 32;
  static constexpr int MIN_CTAS = EW == 8 ? 2 : 1;
  static constexpr int W_STAGE_BYTES = BN * BLOCK_K * 2;
  static constexpr int STAGE_BYTES = A_STAGE_BYTES + W_STAGE_BYTES;
  static constexpr int TMEM_COLS = 2 * BN;        // double-buffered accumulator
  // per-warp staging: 32x36 fp32 transpose buffer + bias slice, or (fp16-only) a 512-byte aligned 32 x 64 B swizzled tile
  static constexpr int STG = EW == 8 ? (F16_ONLY ? 2560 : 32 * STG_LD * 4 + 128) : STG_BYTES;
  static constexpr int LN_BYTES = EW == 8 ? 0 : 8192;   // LayerNorm row-statistics exchange (16-warp kernels only)
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + EW * STG + 1024 /*align slack*/ + 256 /*barriers*/ + LN_BYTES;
};

__device__ __forceinline__ float apply_act(float x, int act) {
  if (act == CFFM_ACT_GELU) return gelu_erf(x);
  if (act == CFFM_ACT_RELU) return fmaxf(x, 0.0f);
  return x;
}

// Persistent: grid = min(#tiles, #SMs); CTA c owns tiles c, c + grid, ... (n fastest, so the CTAs that
// share an A row-panel run at the same time and the panel is read from HBM once).
//   warp 0      TMA producer: smem ring of STAGES x (A 128x64 | W BNx64), runs ahead across tiles
//   warp 1      TMEM allocator + single-thread tcgen05.mma issuer, accumulator stage = tile parity
//   warps 2..17 epilogue: tcgen05.ld (thread = row) -> per-warp smem transpose -> row-contiguous
//               global accesses for residual / fp32 / fp16 (+ optional fused LayerNorm); overlaps the mainloop
template <int BN, bool F16_ONLY, int LN, int EW>   // LN: 0 none, 1 fused LayerNorm, 2 two chained LayerNorms
__global__ void __launch_bounds__(Cfg<BN, EW, F16_ONLY>::THREADS, Cfg<BN, EW, F16_ONLY>::MIN_CTAS)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW,
                    const __grid_constant__ CUtensorMap tmO, const Epilogue ep, const int M, const int N, const int K, const int n_tiles_n,
                    const int n_tiles) {
  using C = Cfg<BN, EW, F16_ONLY>;
  static_assert(LN == 0 || EW == 16, "the fused LayerNorm lives in the 16-warp kernel");
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = ptx::smem_u32(smem_raw);
  const uint32_t pad = ((raw_addr + 1023u) & ~1023u) - raw_addr;
  uint8_t* smem = smem_raw + pad;                              // 1024-byte aligned (SWIZZLE_128B atom)
  uint8_t* sA = smem;
  uint8_t* sW = smem + C::STAGES * A_STAGE_BYTES;
  float* stg_all = reinterpret_cast<float*>(smem + C::STAGES * C::STAGE_BYTES);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + C::STAGES * C::STAGE_BYTES + EW * C::STG);
  uint64_t* empty_bar = full_bar + C::STAGES;
  uint64_t* tfull_bar = empty_bar + C::STAGES;                 // [2] accumulator ready
  uint64_t* tempty_bar = tfull_bar + 2;                        // [2] accumulator drained
  uint32_t* tmem_base_smem = reinterpret_cast<uint32_t*>(tempty_bar + 2);
  float* lnbuf = reinterpret_cast<float*>(smem + C::STAGES * C::STAGE_BYTES + EW * C::STG + 256);   // [2][4][2][32]

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_kb = (K + BLOCK_K - 1) / BLOCK_K;
  const int act = ep.act;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(&tmA);
    ptx::prefetch_tensormap(&tmW);
    if (F16_ONLY) ptx::prefetch_tensormap(&tmO);
    for (int s = 0; s < C::STAGES; ++s) {
      ptx::mbar_init(&full_bar[s], 1);
      ptx::mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      ptx::mbar_init(&tfull_bar[a], 1);
      ptx::mbar_init(&tempty_bar[a], C::TEAM_WARPS);
    }
    ptx::fence_barrier_init();
  }
  if (warp == 1) {                                             // whole warp: .sync.aligned
    ptx::tmem_alloc(tmem_base_smem, C::TMEM_COLS);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_base_smem;
  pdl_sync();                                                  // prologue above overlaps the previous kernel

  // Producer and MMA roles: the WHOLE warp walks the loop and waits on the barriers, one ELECTED lane issues the TMA /
  // tcgen05 instructions.  (Entering a role with `if (lane == 0)` makes every descriptor a per-thread value, and ptxas then
  // wraps each UTMALDG / UTCHMMA / UTCBAR in an ELECT + R2UR.BROADCAST + BRA.U.ANY loop: ~15 dependent instructions per
  // MMA, which made the issuing thread slower than the tensor core for the K = 64 tiles.)
  if (warp == 0) {
    // ===================== TMA producer =====================
    uint32_t g = 0;                                            // k-blocks issued so far (all tiles)
    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
      const int split = t % ep.splits, tt = t / ep.splits;
      const int mt = tt / n_tiles_n, n0 = (tt % n_tiles_n) * BN;
      const int kb0 = split * ep.kb_per, kb1 = min(num_kb, kb0 + ep.kb_per);
      const int img = ep.conv_k ? mt / ep.conv_tiles_y : 0;
      const int y0 = ep.conv_k ? (mt - img * ep.conv_tiles_y) * ep.conv_by * ep.conv_stride - ep.conv_pad : 0;
      for (int kb = kb0; kb < kb1; ++kb, ++g) {
        const uint32_t s = g % C::STAGES, ph = (g / C::STAGES) & 1u;
        ptx::mbar_wait(&empty_bar[s], ph ^ 1u);                // slot free (passes on the first round)
        if (ptx::elect_one()) {
          if (ep.conv_k == 0) {
            ptx::mbar_arrive_expect_tx(&full_bar[s], C::STAGE_BYTES);
            ptx::tma_load_2d(sA + s * A_STAGE_BYTES, &tmA, &full_bar[s], kb * BLOCK_K, mt * BLOCK_M);
          } else {
            const int tap = kb / ep.conv_cchunks, c0 = (kb - tap * ep.conv_cchunks) * BLOCK_K;
            const int ky = tap / ep.conv_k, kx = tap - ky * ep.conv_k;
            ptx::mbar_arrive_expect_tx(&full_bar[s], ep.conv_by * ep.conv_Wo * BLOCK_K * 2 + C::W_STAGE_BYTES);
            ptx::tma_load_4d(sA + s * A_STAGE_BYTES, &tmA, &full_bar[s], c0, kx - ep.conv_pad, y0 + ky, img);
          }
          ptx::tma_load_2d(sW + s * C::W_STAGE_BYTES, &tmW, &full_bar[s], kb * BLOCK_K, n0);
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    constexpr uint32_t idesc = ptx::make_idesc_f16(BLOCK_M, BN);
    uint32_t g = 0, it = 0;
    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x, ++it) {
      const uint32_t acc = it & 1u, aph = (it >> 1) & 1u;
      ptx::mbar_wait(&tempty_bar[acc], aph ^ 1u);              // epilogue drained this accumulator stage
      ptx::tc_fence_after();
      const uint32_t tmem_d = tmem_base + acc * BN;
      const int kb0 = (t % ep.splits) * ep.kb_per, kb1 = min(num_kb, kb0 + ep.kb_per);
      for (int kb = kb0; kb < kb1; ++kb, ++g) {
        const uint32_t s = g % C::STAGES, ph = (g / C::STAGES) & 1u;
        ptx::mbar_wait(&full_bar[s], ph);                      // TMA bytes have landed
        ptx::tc_fence_after();
        const uint64_t da = ptx::make_smem_desc_sw128(ptx::smem_u32(sA + s * A_STAGE_BYTES));
        const uint64_t db = ptx::make_smem_desc_sw128(ptx::smem_u32(sW + s * C::W_STAGE_BYTES));
        if (ptx::elect_one()) {
#pragma unroll
          for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
            // advance 16 halves = 32 bytes inside the swizzle row: +2 in the (addr >> 4) field
            ptx::umma_f16(tmem_d, da + 2u * k, db + 2u * k, idesc, (kb != kb0 || k != 0) ? 1u : 0u);
          }
          ptx::umma_commit(&empty_bar[s]);                     // frees the smem slot when the MMAs retire
          if (kb == kb1 - 1) ptx::umma_commit(&tfull_bar[acc]);     // accumulator complete
        }
        __syncwarp();
      }
    }
  } else {
    // ===================== epilogue: TMEM -> regs -> smem transpose -> coalesced global =====================
    // 16 warps.  Warp (quarter wq, column group cg, team tm) owns the 32 x 32 chunk rows [32 wq, +32) x columns
    // [32 cg, +32) of every tile of its team.  BN = 128: one team, cg = 0..3.  BN = 64: two teams (cg = 0..1) that
    // alternate tiles, so two tiles are drained at once (accumulator stage = team).  Everything that does not
    // depend on the accumulator (bias, residual) is fetched BEFORE waiting for the MMA.
    const int ew = warp - 2;                                   // 0..15
    const int wq = warp & 3;                                   // TMEM lane quarter this warp may access
    const int cg = (ew >> 2) % (BN / 32);
    const int tm = (ew >> 2) / (BN / 32);                      // 0 for BN = 128
    uint8_t* stg8 = reinterpret_cast<uint8_t*>(stg_all) + ew * C::STG;
    float* stg = reinterpret_cast<float*>(stg8);
    float* sbias = reinterpret_cast<float*>(stg8 + C::STG - 128);      // [32] bias slice of this warp
    const int rsub = lane >> 3, csub = (lane & 7) * 4;         // after the fp32 transpose: 4 rows x 8 float4 per pass
    const int tstep = static_cast<int>(gridDim.x) * C::TEAMS;
    uint32_t it = tm;
    for (int t = blockIdx.x + tm * static_cast<int>(gridDim.x); t < n_tiles; t += tstep, it += C::TEAMS) {
      const int split = t % ep.splits, tt = t / ep.splits;
      const int n0 = (tt % n_tiles_n) * BN;
      int m0, mrows;
      tile_rows(ep, tt / n_tiles_n, M, m0, mrows);
      const int mend = m0 + mrows;                             // rows [m0, mend) of the output belong to this tile
      const int wcol0 = n0 + cg * 32;                          // first column of this warp
      const int wrow0 = m0 + wq * 32;
      const uint32_t acc = it & 1u, aph = (it >> 1) & 1u;
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(wq * 32) << 16) + acc * BN + cg * 32;
      if (F16_ONLY) {
        // ---- fp16-only output (q / kv / qkv / fc1 / folded decoder projections)
        {
          const int col = wcol0 + lane;
          sbias[lane] = (ep.bias != nullptr && col < N) ? __ldg(ep.bias + col) : 0.f;
        }
        __syncwarp();
        ptx::mbar_wait(&tfull_bar[acc], aph);
        ptx::tc_fence_after();
        uint32_t v[32];
        ptx::tmem_ld_32x32b_x32(taddr, v);
        ptx::tmem_ld_wait();
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&tempty_bar[acc]);     // registers hold the chunk: MMA may reuse the stage
        // bias + activation in the thread = row layout (bias is warp-uniform: broadcast LDS), pack to fp16 and stage the
        // warp's 32 x 32 chunk as [32 rows][64 B] in the 64-byte TMA swizzle (16-byte piece j of row r sits at
        // j ^ ((r >> 1) & 3): the 8 lanes of a store phase hit 8 distinct 16-byte bank groups).  One lane then hands the
        // tile to the TMA unit: a single bulk tensor store replaces the shared-memory read-back and the predicated
        // global stores (the L1/LSU pipe was the busiest unit of this kernel), and clips the M / N tails in hardware.
        if (lane == 0) ptx::bulk_wait_read0();                 // the previous tile's store has drained this buffer
        __syncwarp();
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float4 b0 = *reinterpret_cast<const float4*>(sbias + 8 * j);
          const float4 b1 = *reinterpret_cast<const float4*>(sbias + 8 * j + 4);
          float a[8] = {__uint_as_float(v[8 * j]) + b0.x, __uint_as_float(v[8 * j + 1]) + b0.y,
                        __uint_as_float(v[8 * j + 2]) + b0.z, __uint_as_float(v[8 * j + 3]) + b0.w,
                        __uint_as_float(v[8 * j + 4]) + b1.x, __uint_as_float(v[8 * j + 5]) + b1.y,
                        __uint_as_float(v[8 * j + 6]) + b1.z, __uint_as_float(v[8 * j + 7]) + b1.w};
          if (act != CFFM_ACT_NONE) {
#pragma unroll
            for (int e = 0; e < 8; ++e) a[e] = apply_act(a[e], act);
          }
          *reinterpret_cast<uint4*>(stg8 + lane * 64 + 16 * (j ^ ((lane >> 1) & 3))) =
              make_uint4(pack_half2(a[0], a[1]), pack_half2(a[2], a[3]), pack_half2(a[4], a[5]), pack_half2(a[6], a[7]));
        }
        ptx::fence_proxy_async();                              // generic-proxy writes -> visible to the TMA (async proxy)
        __syncwarp();
        if (lane == 0 && wcol0 < N && wrow0 < mend) {
          ptx::tma_store_2d(&tmO, stg8, wcol0, wrow0);
          ptx::bulk_commit();
        }
        continue;
      }
      // ---- general path: fp32 and/or fp16 output, optional fp32 residual (may alias out32), optional LayerNorm
      const int col = wcol0 + csub;                            // this lane's 4 columns after the transpose
      const bool cvalid = col < N;                             // N % 8 == 0: the float4 is all-valid or all-invalid
      float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f), r4[8];
      if (ep.bias != nullptr && cvalid) b4 = __ldg(reinterpret_cast<const float4*>(ep.bias + col));
      if (ep.residual != nullptr) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int row = wrow0 + rsub + 4 * i;
          r4[i] = (row < mend && cvalid) ? *reinterpret_cast<const float4*>(ep.residual + static_cast<int64_t>(row) * ep.ldr + col)
                                      : make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
      ptx::mbar_wait(&tfull_bar[acc], aph);
      ptx::tc_fence_after();
      uint32_t v[32];
      ptx::tmem_ld_32x32b_x32(taddr, v);
      ptx::tmem_ld_wait();
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&tempty_bar[acc]);
#pragma unroll
      for (int j = 0; j < 8; ++j)
        *reinterpret_cast<float4*>(stg + lane * STG_LD + 4 * j) =
            make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]), __uint_as_float(v[4 * j + 2]),
                        __uint_as_float(v[4 * j + 3]));
      __syncwarp();
      float4 av[8];                                            // finished output values (kept for the fused LayerNorm)
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int row = wrow0 + rsub + 4 * i;
        float4 a = *reinterpret_cast<const float4*>(stg + (rsub + 4 * i) * STG_LD + csub);
        a.x += b4.x; a.y += b4.y; a.z += b4.z; a.w += b4.w;
        if (act != CFFM_ACT_NONE) {
          a.x = apply_act(a.x, act); a.y = apply_act(a.y, act);
          a.z = apply_act(a.z, act); a.w = apply_act(a.w, act);
        }
        if (ep.residual != nullptr) { a.x += r4[i].x; a.y += r4[i].y; a.z += r4[i].z; a.w += r4[i].w; }
        av[i] = a;
        if (row < mend && cvalid) {
          if (ep.out32 != nullptr && LN != 2)
            *reinterpret_cast<float4*>(ep.out32 + split * ep.split_stride + static_cast<int64_t>(row) * ep.ldo32 + col) = a;
          if (ep.out16 != nullptr) {
            uint2 h;
            h.x = pack_half2(a.x, a.y);
            h.y = pack_half2(a.z, a.w);
            *reinterpret_cast<uint2*>(ep.out16 + static_cast<int64_t>(row) * ep.ldo16 + col) = h;
          }
        }
      }
      __syncwarp();                                            // transpose buffer is re-used by the next tile
      if constexpr (LN != 0) {
        // ---- fused LayerNorm over the N (<= BN) columns of each row.  A row lives in 8 lanes (same rsub) of each
        // of the BN/32 warps of this (team, quarter): two-pass statistics, partial sums exchanged through shared
        // memory with one named barrier per (team, quarter).
        constexpr int NG = BN / 32;                            // warps sharing a row
        const float invN = 1.f / static_cast<float>(N);
        float* bsum = lnbuf + ((tm * 4 + wq) * 2 + 0) * (4 * 32);      // [cg][32 rows]
        float* bsq = lnbuf + ((tm * 4 + wq) * 2 + 1) * (4 * 32);
        const int bar_id = 1 + tm * 4 + wq;
        float mean[8];
        auto row_stats = [&]() {                               // mean[] of the rows held in av[]; leaves the centred sums in bsq
          float st[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            float sacc = cvalid ? (av[i].x + av[i].y) + (av[i].z + av[i].w) : 0.f;
            sacc += __shfl_xor_sync(0xffffffffu, sacc, 1);
            sacc += __shfl_xor_sync(0xffffffffu, sacc, 2);
            sacc += __shfl_xor_sync(0xffffffffu, sacc, 4);
            st[i] = sacc;
          }
          if ((lane & 7) == 0) {
#pragma unroll
            for (int i = 0; i < 8; ++i) bsum[cg * 32 + rsub + 4 * i] = st[i];
          }
          asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "n"(NG * 32) : "memory");
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            float tot = 0.f;
#pragma unroll
            for (int gq = 0; gq < NG; ++gq) tot += bsum[gq * 32 + rsub + 4 * i];
            mean[i] = tot * invN;
          }
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float dx = av[i].x - mean[i], dy = av[i].y - mean[i], dz = av[i].z - mean[i], dw = av[i].w - mean[i];
            float sacc = cvalid ? (dx * dx + dy * dy) + (dz * dz + dw * dw) : 0.f;
            sacc += __shfl_xor_sync(0xffffffffu, sacc, 1);
            sacc += __shfl_xor_sync(0xffffffffu, sacc, 2);
            sacc += __shfl_xor_sync(0xffffffffu, sacc, 4);
            st[i] = sacc;
          }
          if ((lane & 7) == 0) {
#pragma unroll
            for (int i = 0; i < 8; ++i) bsq[cg * 32 + rsub + 4 * i] = st[i];
          }
          asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "n"(NG * 32) : "memory");
        };
        auto row_rstd = [&](int i, float eps) {
          float tot = 0.f;
#pragma unroll
          for (int gq = 0; gq < NG; ++gq) tot += bsq[gq * 32 + rsub + 4 * i];
          return rsqrtf(tot * invN + eps);
        };
        row_stats();
        constexpr bool chain = LN == 2;
        float4 g4 = make_float4(0.f, 0.f, 0.f, 0.f), be4 = g4;
        if (cvalid) {
          g4 = __ldg(reinterpret_cast<const float4*>(ep.ln_gamma + col));
          be4 = __ldg(reinterpret_cast<const float4*>(ep.ln_beta + col));
        }
        if constexpr (chain) {
          // y = LN1(x) is the fp32 output; the second LayerNorm runs on y without leaving the registers
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int row = wrow0 + rsub + 4 * i;
            const float rs = row_rstd(i, ep.ln_eps);
            av[i].x = (av[i].x - mean[i]) * rs * g4.x + be4.x; av[i].y = (av[i].y - mean[i]) * rs * g4.y + be4.y;
            av[i].z = (av[i].z - mean[i]) * rs * g4.z + be4.z; av[i].w = (av[i].w - mean[i]) * rs * g4.w + be4.w;
            if (row < mend && cvalid && ep.out32 != nullptr)
              *reinterpret_cast<float4*>(ep.out32 + static_cast<int64_t>(row) * ep.ldo32 + col) = av[i];
          }
          asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "n"(NG * 32) : "memory");   // everyone has read bsq of the first norm
          row_stats();
          if (cvalid) {
            g4 = __ldg(reinterpret_cast<const float4*>(ep.ln2_gamma + col));
            be4 = __ldg(reinterpret_cast<const float4*>(ep.ln2_beta + col));
          }
        }
        const float eps_out = chain ? ep.ln2_eps : ep.ln_eps;
        if (cvalid) {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int row = wrow0 + rsub + 4 * i;
            const float rs = row_rstd(i, eps_out);
            uint2 h;
            h.x = pack_half2((av[i].x - mean[i]) * rs * g4.x + be4.x, (av[i].y - mean[i]) * rs * g4.y + be4.y);
            h.y = pack_half2((av[i].z - mean[i]) * rs * g4.z + be4.z, (av[i].w - mean[i]) * rs * g4.w + be4.w);
            if (row < mend) *reinterpret_cast<uint2*>(ep.ln_out16 + static_cast<int64_t>(row) * ep.ldln + col) = h;
          }
        }
      }
    }
  }
  if (F16_ONLY && warp >= 2 && lane == 0) ptx::bulk_wait_read0();   // shared memory must outlive the bulk stores reading it
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) ptx::tmem_dealloc(tmem_base, C::TMEM_COLS);
}

// ----------------------------------------------------------------------------------------------
// Cross-check kernel (tests / bring-up only): plain CUDA-core tiled GEMM, same epilogue semantics.
constexpr int CK_T = 64;
__global__ void __launch_bounds__(256)
gemm_check_kernel(const __half* __restrict__ A, int64_t lda, const __half* __restrict__ W, int64_t ldw,
                  const Epilogue ep, int M, int N, int K) {
  pdl_sync();
  __shared__ float sA[16][CK_T + 1];
  __shared__ float sW[16][CK_T + 1];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int m0 = blockIdx.y * CK_T, n0 = blockIdx.x * CK_T;
  float acc[4][4] = {};
  for (int k0 = 0; k0 < K; k0 += 16) {
    for (int i = threadIdx.x; i < CK_T * 16; i += 256) {
      const int r = i >> 4, c = i & 15;
      const int gm = m0 + r, gn = n0 + r, gk = k0 + c;
      sA[c][r] = (gm < M && gk < K) ? __half2float(A[static_cast<int64_t>(gm) * lda + gk]) : 0.f;
      sW[c][r] = (gn < N && gk < K) ? __half2float(W[static_cast<int64_t>(gn) * ldw + gk]) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { a[i] = sA[k][ty * 4 + i]; b[i] = sW[k][tx * 4 + i]; }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) {
      const int m = m0 + ty * 4 + i, n = n0 + tx * 4 + j;
      if (m < M && n < N) {
        float f = acc[i][j] + (ep.bias ? ep.bias[n] : 0.f);
        f = apply_act(f, ep.act);
        if (ep.residual) f += ep.residual[static_cast<int64_t>(m) * ep.ldr + n];
        if (ep.out32) ep.out32[static_cast<int64_t>(m) * ep.ldo32 + n] = f;
        if (ep.out16) ep.out16[static_cast<int64_t>(m) * ep.ldo16 + n] = __float2half_rn(f);
      }
    }
}

// ----------------------------------------------------------------------------------------------
}  // namespace

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

int num_sms() {   // shared: common.cuh.  SM count of the CURRENT device (cached per device)
  static int cache[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev > 63) dev = 0;
  int n = cache[dev];
  if (n == 0) {
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cache[dev] = n;
  }
  return n;
}

// 2-D fp16 row-major [rows, K] with row stride ld (elements); box = [box_rows][64], 128-byte swizzle.
int make_tmap(CUtensorMap* tm, const void* base, int64_t rows, int64_t K, int64_t ld, int box_rows) {   // shared: common.cuh
  EncodeTiledFn fn = get_encode_fn();
  CFFM_REQUIRE(fn != nullptr, CFFM_E_DRIVER, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t gdim[2] = {static_cast<cuuint64_t>(K), static_cast<cuuint64_t>(rows)};
  cuuint64_t gstride[1] = {static_cast<cuuint64_t>(ld) * 2};
  cuuint32_t box[2] = {static_cast<cuuint32_t>(BLOCK_K), static_cast<cuuint32_t>(box_rows)};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  CFFM_REQUIRE(r == CUDA_SUCCESS, CFFM_E_DRIVER, "cuTensorMapEncodeTiled failed with CUresult %d (rows=%lld K=%lld ld=%lld)",
               static_cast<int>(r), (long long)rows, (long long)K, (long long)ld);
  return CFFM_OK;
}

// rank-D (3..5) fp16 tensor map, 128-byte swizzle: dims / box innermost first, strides (elements) of dims 1..rank-1.  The
// innermost box extent must be 64 elements (one 128-byte swizzled row per box position).  Shared: common.cuh.
int make_tmap_nd(CUtensorMap* tm, const void* base, int rank, const int64_t* dims, const int64_t* strides, const int* box) {
  EncodeTiledFn fn = get_encode_fn();
  CFFM_REQUIRE(fn != nullptr, CFFM_E_DRIVER, "cuTensorMapEncodeTiled entry point not available");
  CFFM_REQUIRE(rank >= 3 && rank <= 5, CFFM_E_BADARG, "make_tmap_nd: rank %d", rank);
  cuuint64_t gdim[5], gstride[4];
  cuuint32_t bx[5], estr[5] = {1, 1, 1, 1, 1};
  for (int i = 0; i < rank; ++i) { gdim[i] = static_cast<cuuint64_t>(dims[i]); bx[i] = static_cast<cuuint32_t>(box[i]); }
  for (int i = 0; i + 1 < rank; ++i) gstride[i] = static_cast<cuuint64_t>(strides[i]) * 2;
  CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, rank, const_cast<void*>(base), gdim, gstride, bx, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  CFFM_REQUIRE(r == CUDA_SUCCESS, CFFM_E_DRIVER, "cuTensorMapEncodeTiled (%d-D) failed with CUresult %d (dims %lld %lld %lld ..., box %d %d %d ...)",
               rank, static_cast<int>(r), (long long)dims[0], (long long)dims[1], (long long)dims[2], box[0], box[1], box[2]);
  return CFFM_OK;
}
int make_tmap_4d(CUtensorMap* tm, const void* base, const int64_t dims[4], const int64_t strides[3], const int box[4]) {
  return make_tmap_nd(tm, base, 4, dims, strides, box);
}

// Convolution operand: NHWC fp16 image [n, H, W, C] walked with the convolution stride.  Box = 64 channels x Wo output
// columns x by output rows of one image: boxDim = N * elementStride loads N elements (cuda.h, cuTensorMapEncodeTiled).
struct ConvDesc {
  const void* x;
  int n, H, W, C, k, stride, pad;
  int Ho, Wo, by, tiles_y;
};

static int conv_desc(ConvDesc* cv, const void* x, int n, int H, int W, int C, int k, int stride, int pad) {
  CFFM_REQUIRE(x && n > 0 && H > 0 && W > 0 && k >= 1 && stride >= 1 && stride <= 8 && pad >= 0 && pad < k, CFFM_E_BADARG,
               "conv_gemm: bad geometry n=%d H=%d W=%d k=%d stride=%d pad=%d", n, H, W, k, stride, pad);
  CFFM_REQUIRE(C % 64 == 0, CFFM_E_UNSUPPORTED, "conv_gemm: input channels must be a multiple of 64 (one k-block = one tap x 64 channels), got %d", C);
  CFFM_REQUIRE(aligned16(x), CFFM_E_BADARG, "conv_gemm: misaligned image");
  cv->x = x; cv->n = n; cv->H = H; cv->W = W; cv->C = C; cv->k = k; cv->stride = stride; cv->pad = pad;
  cv->Ho = (H + 2 * pad - k) / stride + 1;
  cv->Wo = (W + 2 * pad - k) / stride + 1;
  CFFM_REQUIRE(cv->Wo <= 128 && cv->Wo * stride <= 256, CFFM_E_UNSUPPORTED,
               "conv_gemm: output width %d (x stride %d) exceeds one TMA box; use cffm_im2col + cffm_gemm_f16 for this size", cv->Wo, stride);
  int by = 128 / cv->Wo;
  if (by * stride > 256) by = 256 / stride;
  if (by > cv->Ho) by = cv->Ho;
  cv->by = by;
  cv->tiles_y = (cv->Ho + by - 1) / by;
  return CFFM_OK;
}

static int make_tmap_conv(CUtensorMap* tm, const ConvDesc& cv) {
  EncodeTiledFn fn = get_encode_fn();
  CFFM_REQUIRE(fn != nullptr, CFFM_E_DRIVER, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t gdim[4] = {static_cast<cuuint64_t>(cv.C), static_cast<cuuint64_t>(cv.W), static_cast<cuuint64_t>(cv.H), static_cast<cuuint64_t>(cv.n)};
  cuuint64_t gstride[3] = {static_cast<cuuint64_t>(cv.C) * 2, static_cast<cuuint64_t>(cv.W) * cv.C * 2,
                           static_cast<cuuint64_t>(cv.H) * cv.W * cv.C * 2};
  cuuint32_t box[4] = {static_cast<cuuint32_t>(BLOCK_K), static_cast<cuuint32_t>(cv.Wo * cv.stride), static_cast<cuuint32_t>(cv.by * cv.stride), 1};
  cuuint32_t estr[4] = {1, static_cast<cuuint32_t>(cv.stride), static_cast<cuuint32_t>(cv.stride), 1};
  CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(cv.x), gdim, gstride, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  CFFM_REQUIRE(r == CUDA_SUCCESS, CFFM_E_DRIVER, "cuTensorMapEncodeTiled (conv) failed with CUresult %d (C=%d W=%d H=%d n=%d, box %u %u %u)",
               static_cast<int>(r), cv.C, cv.W, cv.H, cv.n, box[0], box[1], box[2]);
  return CFFM_OK;
}

// fp16 [rows, cols] output (row stride ld elements) as the target of per-warp 32 x 32 bulk stores (64-byte swizzle).
// 4-D fp32 tensor map without swizzle (dims / box innermost first, strides in ELEMENTS of dims 1..3): the NCHW input frames
// of the stage-1 patch embedding.  Shared: common.cuh.
int make_tmap_f32_4d(CUtensorMap* tm, const void* base, const int64_t dims[4], const int64_t strides[3], const int box[4]) {
  EncodeTiledFn fn = get_encode_fn();
  CFFM_REQUIRE(fn != nullptr, CFFM_E_DRIVER, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t gdim[4], gstride[3];
  cuuint32_t bx[4], estr[4] = {1, 1, 1, 1};
  for (int i = 0; i < 4; ++i) { gdim[i] = static_cast<cuuint64_t>(dims[i]); bx[i] = static_cast<cuuint32_t>(box[i]); }
  for (int i = 0; i < 3; ++i) gstride[i] = static_cast<cuuint64_t>(strides[i]) * 4;
  CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<void*>(base), gdim, gstride, bx, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  CFFM_REQUIRE(r == CUDA_SUCCESS, CFFM_E_DRIVER, "cuTensorMapEncodeTiled (fp32 4-D) failed with CUresult %d (dims %lld %lld %lld %lld, box %d %d %d)",
               static_cast<int>(r), (long long)dims[0], (long long)dims[1], (long long)dims[2], (long long)dims[3], box[0], box[1], box[2]);
  return CFFM_OK;
}

static int make_tmap_out(CUtensorMap* tm, const void* base, int64_t rows, int64_t cols, int64_t ld) {
  EncodeTiledFn fn = get_encode_fn();
  CFFM_REQUIRE(fn != nullptr, CFFM_E_DRIVER, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t gdim[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  cuuint64_t gstride[1] = {static_cast<cuuint64_t>(ld) * 2};
  cuuint32_t box[2] = {32, 32};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  CFFM_REQUIRE(r == CUDA_SUCCESS, CFFM_E_DRIVER, "cuTensorMapEncodeTiled (output) failed with CUresult %d (rows=%lld cols=%lld ld=%lld)",
               static_cast<int>(r), (long long)rows, (long long)cols, (long long)ld);
  return CFFM_OK;
}

namespace {

template <int BN, bool F16_ONLY, int LN = 0, int EW = 16>
int launch_tcgen05(const void* A, int64_t lda, const void* W, int64_t ldw, const Epilogue& ep_in, int M, int N, int K,
                   cudaStream_t st, const ConvDesc* cv = nullptr) {
  using C = Cfg<BN, EW, F16_ONLY>;
  CUtensorMap tmA, tmW;
  Epilogue ep = ep_in;
  int rc;
  if (cv != nullptr) {                                         // implicit-GEMM convolution: A = the image itself
    CFFM_REQUIRE(!F16_ONLY, CFFM_E_UNSUPPORTED, "conv_gemm: the bulk-store epilogue needs 128-row aligned tiles");
    ep.conv_k = cv->k; ep.conv_stride = cv->stride; ep.conv_pad = cv->pad; ep.conv_cchunks = cv->C / BLOCK_K;
    ep.conv_Wo = cv->Wo; ep.conv_Ho = cv->Ho; ep.conv_by = cv->by; ep.conv_tiles_y = cv->tiles_y;
    rc = make_tmap_conv(&tmA, *cv);
  } else {
    rc = make_tmap(&tmA, A, M, K, lda, BLOCK_M);
  }
  if (rc) return rc;
  rc = make_tmap(&tmW, W, N, K, ldw, BN);
  if (rc) return rc;
  CUtensorMap tmO = tmA;                                       // only read by the fp16-only epilogue
  if (F16_ONLY) {
    rc = make_tmap_out(&tmO, ep.out16, M, N, ep.ldo16);
    if (rc) return rc;
  }
  if ((rc = set_dyn_smem<gemm_tcgen05_kernel<BN, F16_ONLY, LN, EW>>(C::SMEM_BYTES, "gemm"))) return rc;
  const int tiles_n = (N + BN - 1) / BN, tiles_m = cv != nullptr ? cv->n * cv->tiles_y : (M + BLOCK_M - 1) / BLOCK_M;
  const int tiles = tiles_n * tiles_m * ep.splits;
  const int slots = num_sms() * C::MIN_CTAS;
  const int grid = tiles < slots ? tiles : slots;
  launch_k(gemm_tcgen05_kernel<BN, F16_ONLY, LN, EW>, grid, C::THREADS, C::SMEM_BYTES, st, tmA, tmW, tmO, ep, M, N, K, tiles_n, tiles);
  return launch_status("gemm_tcgen05_kernel");
}

// Few-tile GEMMs (latency-bound) run as 64-wide tiles in the two-CTAs-per-SM configuration.  The limit is in units of
// 128 x 64 tiles; it depends on the shape only (never on the data), and the per-element summation order over K is the
// same in every configuration, so the choice cannot change a result bit.
int small_gemm_max_tiles() {
  static const int v = [] {
    const char* e = getenv("CFFM_GEMM_SMALL_TILES");           // tuning override
    return e ? atoi(e) : 300;                                 // measured: beyond ~2 waves of the 296 CTA slots the 16-warp kernel wins
  }();
  return v;
}

}  // namespace
}  // namespace cffm

extern "C" int cffm_gemm_f16(const void* A, int64_t lda, const void* W, int64_t ldw, const float* bias,
                             const float* residual, int64_t ldr, void* out_f16, int64_t ldo16, float* out_f32,
                             int64_t ldo32, int M, int N, int K, int act, int impl, void* stream) {
  using namespace cffm;
  CFFM_REQUIRE(A && W && (out_f16 || out_f32), CFFM_E_BADARG, "gemm: null operand");
  CFFM_REQUIRE(M > 0 && N > 0 && K > 0, CFFM_E_BADARG, "gemm: non-positive size M=%d N=%d K=%d", M, N, K);
  CFFM_REQUIRE(K % 8 == 0 && N % 8 == 0 && lda % 8 == 0 && ldw % 8 == 0 && lda >= K && ldw >= K, CFFM_E_UNSUPPORTED,
               "gemm: need K,N,lda,ldw multiples of 8 (M=%d N=%d K=%d lda=%lld ldw=%lld)", M, N, K, (long long)lda,
               (long long)ldw);
  CFFM_REQUIRE(aligned16(A) && aligned16(W) && aligned16(bias) && aligned16(residual) && aligned16(out_f16) &&
                   aligned16(out_f32),
               CFFM_E_BADARG, "gemm: pointers must be 16-byte aligned");
  CFFM_REQUIRE((!out_f16 || (ldo16 % 8 == 0 && ldo16 >= N)) && (!out_f32 || (ldo32 % 4 == 0 && ldo32 >= N)) &&
                   (!residual || (ldr % 4 == 0 && ldr >= N)),
               CFFM_E_BADARG, "gemm: bad output/residual stride");
  CFFM_REQUIRE(act >= CFFM_ACT_NONE && act <= CFFM_ACT_RELU, CFFM_E_BADARG, "gemm: bad act %d", act);
  Epilogue ep{bias, residual, ldr, static_cast<__half*>(out_f16), ldo16, out_f32, ldo32, act, nullptr, nullptr, 0.f, nullptr, 0,
              nullptr, nullptr, 0.f, 1, (K + BLOCK_K - 1) / BLOCK_K, 0};
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (impl == CFFM_GEMM_CHECK) {
    dim3 grid((N + CK_T - 1) / CK_T, (M + CK_T - 1) / CK_T);
    launch_k(gemm_check_kernel, grid, 256, 0, st, static_cast<const __half*>(A), lda, static_cast<const __half*>(W), ldw, ep,
                                            M, N, K);
    return launch_status("gemm_check_kernel");
  }
  CFFM_REQUIRE(impl == CFFM_GEMM_TCGEN05, CFFM_E_BADARG, "gemm: bad impl %d", impl);
  const bool f16_only = out_f16 != nullptr && out_f32 == nullptr && residual == nullptr;
  const int tiles64 = ((M + BLOCK_M - 1) / BLOCK_M) * ((N + 63) / 64);
  if (tiles64 <= small_gemm_max_tiles()) {
    return f16_only ? launch_tcgen05<64, true, 0, 8>(A, lda, W, ldw, ep, M, N, K, st)
                    : launch_tcgen05<64, false, 0, 8>(A, lda, W, ldw, ep, M, N, K, st);
  }
  const bool wide = N % 128 == 0 || (N % 64 != 0 && N > 64);   // 128-wide tiles halve the A re-reads of the big GEMMs
  if (f16_only) {
    return wide ? launch_tcgen05<128, true>(A, lda, W, ldw, ep, M, N, K, st)
                : launch_tcgen05<64, true>(A, lda, W, ldw, ep, M, N, K, st);
  }
  return wide ? launch_tcgen05<128, false>(A, lda, W, ldw, ep, M, N, K, st)
              : launch_tcgen05<64, false>(A, lda, W, ldw, ep, M, N, K, st);
}

static int gemm_ln_impl(const void* A, int64_t lda, const void* W, int64_t ldw, const float* bias, const float* residual,
                       int64_t ldr, float* out_f32, int64_t ldo32, const float* ln_gamma, const float* ln_beta, float ln_eps,
                       const float* ln2_gamma, const float* ln2_beta, float ln2_eps, void* ln_out_f16, int64_t ldln, int M,
                       int N, int K, void* stream) {
  using namespace cffm;
  CFFM_REQUIRE(A && W && ln_gamma && ln_beta && ln_out_f16, CFFM_E_BADARG, "gemm_ln: null operand");
  CFFM_REQUIRE(M > 0 && N > 0 && K > 0, CFFM_E_BADARG, "gemm_ln: non-positive size M=%d N=%d K=%d", M, N, K);
  CFFM_REQUIRE(N <= 128, CFFM_E_UNSUPPORTED, "gemm_ln: the fused LayerNorm needs the whole row in one tile (N <= 128), got N=%d", N);
  CFFM_REQUIRE(K % 8 == 0 && N % 8 == 0 && lda % 8 == 0 && ldw % 8 == 0 && lda >= K && ldw >= K, CFFM_E_UNSUPPORTED,
               "gemm_ln: need K,N,lda,ldw multiples of 8 (M=%d N=%d K=%d)", M, N, K);
  CFFM_REQUIRE(aligned16(A) && aligned16(W) && aligned16(bias) && aligned16(residual) && aligned16(out_f32) &&
                   aligned16(ln_gamma) && aligned16(ln_beta) && (reinterpret_cast<uintptr_t>(ln_out_f16) & 7) == 0,
               CFFM_E_BADARG, "gemm_ln: misaligned pointer");
  CFFM_REQUIRE((!out_f32 || (ldo32 % 4 == 0 && ldo32 >= N)) && (!residual || (ldr % 4 == 0 && ldr >= N)) && ldln % 4 == 0 &&
                   ldln >= N,
               CFFM_E_BADARG, "gemm_ln: bad output/residual stride");
  Epilogue ep{bias, residual, ldr, nullptr, 0, out_f32, ldo32, CFFM_ACT_NONE, ln_gamma, ln_beta, ln_eps,
              static_cast<__half*>(ln_out_f16), ldln, ln2_gamma, ln2_beta, ln2_eps, 1, (K + BLOCK_K - 1) / BLOCK_K, 0};
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (ln2_gamma != nullptr)
    return N > 64 ? launch_tcgen05<128, false, 2>(A, lda, W, ldw, ep, M, N, K, st) : launch_tcgen05<64, false, 2>(A, lda, W, ldw, ep, M, N, K, st);
  return N > 64 ? launch_tcgen05<128, false, 1>(A, lda, W, ldw, ep, M, N, K, st) : launch_tcgen05<64, false, 1>(A, lda, W, ldw, ep, M, N, K, st);
}

extern "C" int cffm_gemm_f16_ln(const void* A, int64_t lda, const void* W, int64_t ldw, const float* bias,
                                const float* residual, int64_t ldr, float* out_f32, int64_t ldo32, const float* ln_gamma,
                                const float* ln_beta, float ln_eps, void* ln_out_f16, int64_t ldln, int M, int N, int K,
                                void* stream) {
  return gemm_ln_impl(A, lda, W, ldw, bias, residual, ldr, out_f32, ldo32, ln_gamma, ln_beta, ln_eps, nullptr, nullptr, 0.f,
                      ln_out_f16, ldln, M, N, K, stream);
}

extern "C" int cffm_gemm_f16_ln_chain(const void* A, int64_t lda, const void* W, int64_t ldw, const float* bias, float* out_f32,
                                      int64_t ldo32, const float* gamma1, const float* beta1, float eps1, const float* gamma2,
                                      const float* beta2, float eps2, void* ln_out_f16, int64_t ldln, int M, int N, int K,
                                      void* stream) {
  using namespace cffm;
  CFFM_REQUIRE(gamma2 && beta2 && out_f32, CFFM_E_BADARG, "gemm_ln_chain: null operand");
  CFFM_REQUIRE(aligned16(gamma2) && aligned16(beta2), CFFM_E_BADARG, "gemm_ln_chain: misaligned pointer");
  return gemm_ln_impl(A, lda, W, ldw, bias, nullptr, 0, out_f32, ldo32, gamma1, beta1, eps1, gamma2, beta2, eps2, ln_out_f16, ldln,
                      M, N, K, stream);
}

extern "C" int cffm_gemm_f16_splitk(const void* A, int64_t lda, const void* W, int64_t ldw, float* partials, int M, int N,
                                    int K, int splits, void* stream) {
  using namespace cffm;
  CFFM_REQUIRE(A && W && partials, CFFM_E_BADARG, "gemm_splitk: null operand");
  CFFM_REQUIRE(M > 0 && N > 0 && K > 0 && splits >= 1, CFFM_E_BADARG, "gemm_splitk: bad size M=%d N=%d K=%d S=%d", M, N, K, splits);
  CFFM_REQUIRE(K % 8 == 0 && N % 8 == 0 && lda % 8 == 0 && ldw % 8 == 0 && lda >= K && ldw >= K, CFFM_E_UNSUPPORTED,
               "gemm_splitk: need K,N,lda,ldw multiples of 8 (M=%d N=%d K=%d)", M, N, K);
  CFFM_REQUIRE(aligned16(A) && aligned16(W) && aligned16(partials), CFFM_E_BADARG, "gemm_splitk: misaligned pointer");
  const int num_kb = (K + BLOCK_K - 1) / BLOCK_K;
  CFFM_REQUIRE(splits <= num_kb, CFFM_E_BADARG, "gemm_splitk: %d splits for %d k-blocks", splits, num_kb);
  const int kb_per = (num_kb + splits - 1) / splits;
  CFFM_REQUIRE((splits - 1) * kb_per < num_kb, CFFM_E_BADARG, "gemm_splitk: empty split (use cffm_splitk_plan)");
  Epilogue ep{nullptr, nullptr, 0, nullptr, 0, partials, N, CFFM_ACT_NONE, nullptr, nullptr, 0.f, nullptr, 0, nullptr, nullptr, 0.f,
              splits, kb_per, static_cast<int64_t>(M) * N};
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int tiles64 = ((M + BLOCK_M - 1) / BLOCK_M) * ((N + 63) / 64) * splits;
  if (tiles64 <= small_gemm_max_tiles()) return launch_tcgen05<64, false, 0, 8>(A, lda, W, ldw, ep, M, N, K, st);
  const bool wide = N % 128 == 0 || (N % 64 != 0 && N > 64);
  return wide ? launch_tcgen05<128, false>(A, lda, W, ldw, ep, M, N, K, st) : launch_tcgen05<64, false>(A, lda, W, ldw, ep, M, N, K, st);
}

extern "C" int cffm_splitk_plan(int M, int N, int K) {
  // Split count for the long-K convolutions of the path.  It depends on K ONLY: the summation order of a row must not
  // change with the batch size, or a clip would no longer produce bit-identical results alone and inside a batch.
  // One split per 8 k-blocks (512 columns of K), at most 8: K = 4096 -> 8, 2880 -> 5, 2048 -> 4, 1152..1280 -> 2.
  using namespace cffm;
  (void)M; (void)N;
  const int num_kb = (K + BLOCK_K - 1) / BLOCK_K;
  int s = num_kb / 8;
  if (s > 8) s = 8;
  if (s < 1) s = 1;
  const int kb_per = (num_kb + s - 1) / s;
  return (num_kb + kb_per - 1) / kb_per;                       // no empty split
}

// ------------------------------------------------------------------------------------------------
// Convolutions as implicit GEMMs: no im2col matrix is ever written.
extern "C" int cffm_conv_gemm_f16_ln(const void* x, int n, int H, int W, int C, int ksize, int stride, int pad, const void* Wt,
                                     int64_t ldw, const float* bias, float* out_f32, int64_t ldo32, const float* gamma1,
                                     const float* beta1, float eps1, const float* gamma2, const float* beta2, float eps2,
                                     void* ln_out_f16, int64_t ldln, int Nout, void* stream) {
  using namespace cffm;
  ConvDesc cv;
  int rc = conv_desc(&cv, x, n, H, W, C, ksize, stride, pad);
  if (rc) return rc;
  const int M = n * cv.Ho * cv.Wo, K = ksize * ksize * C;
  CFFM_REQUIRE(Wt && gamma1 && beta1 && ln_out_f16, CFFM_E_BADARG, "conv_gemm_ln: null operand");
  CFFM_REQUIRE(Nout > 0 && Nout <= 128 && Nout % 8 == 0, CFFM_E_UNSUPPORTED,
               "conv_gemm_ln: the fused LayerNorm needs the whole row in one tile (N <= 128, multiple of 8), got N=%d", Nout);
  CFFM_REQUIRE(ldw % 8 == 0 && ldw >= K, CFFM_E_UNSUPPORTED, "conv_gemm_ln: bad weight stride %lld for K=%d", (long long)ldw, K);
  CFFM_REQUIRE(aligned16(Wt) && aligned16(bias) && aligned16(out_f32) && aligned16(gamma1) && aligned16(beta1) && aligned16(gamma2) &&
                   aligned16(beta2) && (reinterpret_cast<uintptr_t>(ln_out_f16) & 7) == 0,
               CFFM_E_BADARG, "conv_gemm_ln: misaligned pointer");
  CFFM_REQUIRE((!out_f32 || (ldo32 % 4 == 0 && ldo32 >= Nout)) && ldln % 4 == 0 && ldln >= Nout && (!gamma2 || (beta2 && out_f32)),
               CFFM_E_BADARG, "conv_gemm_ln: bad output stride / chained LayerNorm operands");
  Epilogue ep{bias, nullptr, 0, nullptr, 0, out_f32, ldo32, CFFM_ACT_NONE, gamma1, beta1, eps1,
              static_cast<__half*>(ln_out_f16), ldln, gamma2, beta2, eps2, 1, (K + BLOCK_K - 1) / BLOCK_K, 0};
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (gamma2 != nullptr)
    return Nout > 64 ? launch_tcgen05<128, false, 2>(nullptr, 0, Wt, ldw, ep, M, Nout, K, st, &cv)
                     : launch_tcgen05<64, false, 2>(nullptr, 0, Wt, ldw, ep, M, Nout, K, st, &cv);
  return Nout > 64 ? launch_tcgen05<128, false, 1>(nullptr, 0, Wt, ldw, ep, M, Nout, K, st, &cv)
                   : launch_tcgen05<64, false, 1>(nullptr, 0, Wt, ldw, ep, M, Nout, K, st, &cv);
}

extern "C" int cffm_conv_gemm_f16_splitk(const void* x, int n, int H, int W, int C, int ksize, int stride, int pad, const void* Wt,
                                         int64_t ldw, float* partials, int Nout, int splits, void* stream) {
  using namespace cffm;
  ConvDesc cv;
  int rc = conv_desc(&cv, x, n, H, W, C, ksize, stride, pad);
  if (rc) return rc;
  const int M = n * cv.Ho * cv.Wo, K = ksize * ksize * C;
  CFFM_REQUIRE(Wt && partials && Nout > 0 && Nout % 8 == 0 && splits >= 1, CFFM_E_BADARG, "conv_gemm_splitk: bad operand");
  CFFM_REQUIRE(ldw % 8 == 0 && ldw >= K && aligned16(Wt) && aligned16(partials), CFFM_E_BADARG, "conv_gemm_splitk: bad weight stride / alignment");
  const int num_kb = K / BLOCK_K;
  CFFM_REQUIRE(splits <= num_kb, CFFM_E_BADARG, "conv_gemm_splitk: %d splits for %d k-blocks", splits, num_kb);
  const int kb_per = (num_kb + splits - 1) / splits;
  CFFM_REQUIRE((splits - 1) * kb_per < num_kb, CFFM_E_BADARG, "conv_gemm_splitk: empty split (use cffm_splitk_plan)");
  Epilogue ep{nullptr, nullptr, 0, nullptr, 0, partials, Nout, CFFM_ACT_NONE, nullptr, nullptr, 0.f, nullptr, 0, nullptr, nullptr, 0.f,
              splits, kb_per, static_cast<int64_t>(M) * Nout};
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int tiles64 = cv.n * cv.tiles_y * ((Nout + 63) / 64) * splits;
  if (tiles64 <= small_gemm_max_tiles()) return launch_tcgen05<64, false, 0, 8>(nullptr, 0, Wt, ldw, ep, M, Nout, K, st, &cv);
  const bool wide = Nout % 128 == 0 || (Nout % 64 != 0 && Nout > 64);
  return wide ? launch_tcgen05<128, false>(nullptr, 0, Wt, ldw, ep, M, Nout, K, st, &cv)
              : launch_tcgen05<64, false>(nullptr, 0, Wt, ldw, ep, M, Nout, K, st, &cv);
}
