// Mix-FFN tail in one kernel: x += fc2(GELU(dwconv3x3(h) + b_dw)) + b_2, then the next LayerNorm of the row
// (Mlp.forward, mix_transformer.py:52-58 with DWConv :361-368; the residual add and the following norm are Block.forward
// :84-88).  h = fc1(x) is the only tensor of the FFN that still goes through HBM: the depthwise-convolved, GELU'd hidden map
// (59 MB per 8 frames at stage 1, written once and read once by the separate kernels) never leaves the SM.
//
// Persistent CTAs loop over 128-token tiles = TH x TW pixel patches of one frame (4 x 32 or 8 x 16).  Per 64-channel k-block:
//   TMA      the patch's halo box [(TH+2) x (TW+2) x 64 channels] of h (4-D tensor map over the NHWC map; the out-of-bounds
//            zero fill IS the convolution's zero padding) and the k-block of W2 [N x 64], 128-byte swizzle, 3-stage ring
//   compute  16 warps, thread = (token of the tile, two 8-channel pieces): 9 taps x 8 channels of mixed-precision FMA
//            (fp16 x fp16 + fp32: the products are exact), + bias, exact-erf GELU, packed to fp16 and written into TENSOR
//            MEMORY with tcgen05.st: the thread = token layout is exactly the A-operand layout of a TS-mode MMA, so the
//            activated tile needs no shared-memory round trip and no proxy fence (same pattern as P in the attention kernels)
//   tcgen05  acc[128 x N] += A_chunk[128 x 64] (TMEM) . W2_chunk^T (shared memory), one elected lane
// After the last k-block the same 16 warps drain the accumulator: tcgen05.ld -> per-warp shared-memory transpose -> bias +
// fp32 residual (in place) in whole 64/128-byte row segments -> fused LayerNorm (row statistics exchanged between the four
// warps that share a row) -> fp16.
// Warp roles: 0..15 compute / epilogue, 16 TMA producer, 17 TMEM allocator + MMA issuer.
#include <cuda.h>

#include "common.cuh"
#include "ptx_sm100.cuh"

namespace cffm {
namespace {

constexpr int TOK = 128, ROWB = 128;                           // tokens per tile, bytes per 64-channel row
constexpr int CW_WARPS = 16, FFN_THREADS = (CW_WARPS + 2) * 32;
constexpr int STAGES = 3, A_RING = 4, A_COLS = 32;             // smem ring (halo + W2 chunk), TMEM ring of A chunks (64 ch = 32 cells)
constexpr int HALO_ROWS_MAX = 6 * 34;                          // (4+2) x (32+2) >= (8+2) x (16+2)
constexpr int HALO_BYTES = (HALO_ROWS_MAX * ROWB + 1023) / 1024 * 1024;
constexpr int HD_MAX = 512;                                    // hidden width whose depthwise weights fit the budget below
constexpr int TMEM_A = 128;                                    // accumulator [0, N <= 128) | A ring [128, 256)

template <int N>
struct FfnCfg {
  static constexpr int CW = N / 4;                             // accumulator columns per epilogue warp: 16 or 32
  static constexpr int STG_LD = CW + 4;                        // padded fp32 row of the per-warp transpose buffer
  static constexpr int STG_BYTES = 32 * STG_LD * 4;
  static constexpr int W_BYTES = N * ROWB;
  static constexpr int STAGE_BYTES = HALO_BYTES + W_BYTES;
  static constexpr int DW_BYTES = 9 * HD_MAX * 2 + HD_MAX * 4; // depthwise weights [9][HD] fp16 + bias fp32
  static constexpr int LN_BYTES = 4 * 2 * 4 * 32 * 4;          // [quarter][sum | sq][column group][row]
  static constexpr int SMEM = STAGES * STAGE_BYTES + DW_BYTES + CW_WARPS * STG_BYTES + LN_BYTES + 256 /*barriers*/ + 1024 /*align*/;
  static_assert(SMEM <= 232448, "shared memory budget of one CTA per SM");
};

struct FfnParams {
  const __half* dw_w;       // [9, HD] fp16 (tap-major)
  const float* dw_b;        // [HD]
  const float* b2;          // [N]
  const float* residual;    // [M, N] fp32
  float* out32;             // [M, N] fp32 or null (may alias residual)
  const float* gamma;       // LayerNorm of the finished row -> ln_out (or null: no LayerNorm)
  const float* beta;
  float eps;
  __half* ln_out;           // [M, N] fp16
  int n, H, W, HD;
  int th, tw, tw_shift;     // tile: th x tw pixels, tw = 1 << tw_shift
  int tiles_x, tiles_y, n_tiles;
  int halo_bytes;           // bytes of one halo box
};

// acc += a * b for two packed fp16 pairs: fp16 x fp16 products are exact in fp32, one mixed-precision FMA each
__device__ __forceinline__ void fhfma2_mul(uint32_t a, uint32_t b, float& lo, float& hi) {
  asm("{\n\t.reg .b16 al, ah, bl, bh;\n\t"
      "mov.b32 {al, ah}, %2;\n\t"
      "mov.b32 {bl, bh}, %3;\n\t"
      "fma.rn.f32.f16 %0, al, bl, %0;\n\t"
      "fma.rn.f32.f16 %1, ah, bh, %1;\n\t}"
      : "+f"(lo), "+f"(hi)
      : "r"(a), "r"(b));
}

template <int N>
__global__ void __launch_bounds__(FFN_THREADS, 1)
mixffn_tail_kernel(const __grid_constant__ CUtensorMap tmH, const __grid_constant__ CUtensorMap tmW, const FfnParams p) {
  using C = FfnCfg<N>;
  constexpr int CW = C::CW, STG_LD = C::STG_LD;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = ptx::smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);
  uint8_t* sStage = smem;                                      // STAGES x (halo | W2 chunk), each part 1024-byte aligned
  __half* sDw = reinterpret_cast<__half*>(smem + STAGES * C::STAGE_BYTES);
  float* sDb = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(sDw) + 9 * HD_MAX * 2);
  float* sStg = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(sDw) + C::DW_BYTES);
  float* sLn = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(sStg) + CW_WARPS * C::STG_BYTES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(sLn) + C::LN_BYTES);
  uint64_t* full = bars + 0;        // [STAGES] TMA -> compute warps (halo) and MMA issuer (W2 chunk)
  uint64_t* empty = bars + 3;       // [STAGES] 16 compute warps (halo read) + MMA commit (W2 consumed) -> TMA
  uint64_t* a_full = bars + 6;      // [A_RING] compute warps -> MMA issuer: A chunk written to TMEM
  uint64_t* a_empty = bars + 10;    // [A_RING] MMA commit -> compute warps
  uint64_t* acc_full = bars + 14;   // MMA commit -> epilogue: accumulator of the tile complete
  uint64_t* acc_empty = bars + 15;  // epilogue (16 warps) -> MMA issuer: accumulator read out
  uint32_t* tmem_base_smem = reinterpret_cast<uint32_t*>(bars + 16);
  static_assert(STAGES == 3 && A_RING == 4, "barrier layout");

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nkb = p.HD >> 6;

  if (warp == CW_WARPS && lane == 0) {
    ptx::prefetch_tensormap(&tmH);
    ptx::prefetch_tensormap(&tmW);
    for (int i = 0; i < STAGES; ++i) { ptx::mbar_init(&full[i], 1); ptx::mbar_init(&empty[i], CW_WARPS + 1); }
    for (int i = 0; i < A_RING; ++i) { ptx::mbar_init(&a_full[i], CW_WARPS); ptx::mbar_init(&a_empty[i], 1); }
    ptx::mbar_init(acc_full, 1);
    ptx::mbar_init(acc_empty, CW_WARPS);
    ptx::fence_barrier_init();
  }
  if (warp == CW_WARPS + 1) {
    ptx::tmem_alloc(tmem_base_smem, 256);
    ptx::tmem_relinquish();
  }
  // depthwise weights and bias: constants of the layer, fetched before the programmatic-dependent-launch wait
  for (int i = threadIdx.x; i < 9 * p.HD / 8; i += FFN_THREADS) {
    const int tap = i / (p.HD / 8), c8 = i - tap * (p.HD / 8);
    reinterpret_cast<uint4*>(sDw)[tap * (HD_MAX / 8) + c8] = __ldg(reinterpret_cast<const uint4*>(p.dw_w) + i);
  }
  for (int i = threadIdx.x; i < p.HD / 4; i += FFN_THREADS) reinterpret_cast<float4*>(sDb)[i] = __ldg(reinterpret_cast<const float4*>(p.dw_b) + i);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_base_smem;
  pdl_sync();                                                  // prologue above overlaps the previous kernel

  if (warp == CW_WARPS) {
    // ===================== TMA producer =====================
    uint32_t g = 0;
    for (int t = blockIdx.x; t < p.n_tiles; t += gridDim.x) {
      const int img = t / (p.tiles_x * p.tiles_y), r = t - img * (p.tiles_x * p.tiles_y);
      const int ty = r / p.tiles_x, tx = r - ty * p.tiles_x;
      const int x0 = tx * p.tw - 1, y0 = ty * p.th - 1;
      for (int kb = 0; kb < nkb; ++kb, ++g) {
        const uint32_t s = g % STAGES, ph = (g / STAGES) & 1u;
        ptx::mbar_wait(&empty[s], ph ^ 1u);
        if (ptx::elect_one()) {
          uint8_t* st = sStage + s * C::STAGE_BYTES;
          ptx::mbar_arrive_expect_tx(&full[s], p.halo_bytes + C::W_BYTES);
          ptx::tma_load_4d(st, &tmH, &full[s], kb * 64, x0, y0, img);
          ptx::tma_load_2d(st + HALO_BYTES, &tmW, &full[s], kb * 64, 0);
        }
        __syncwarp();
      }
    }
  } else if (warp == CW_WARPS + 1) {
    // ===================== MMA issuer =====================
    constexpr uint32_t idesc = ptx::make_idesc_f16(TOK, N);   // A from TMEM (K-major cells), B = W2 chunk K-major
    uint32_t g = 0, ti = 0;
    for (int t = blockIdx.x; t < p.n_tiles; t += gridDim.x, ++ti) {
      ptx::mbar_wait(acc_empty, (ti & 1u) ^ 1u);               // the previous tile's accumulator has been read out
      ptx::tc_fence_after();
      for (int kb = 0; kb < nkb; ++kb, ++g) {
        const uint32_t s = g % STAGES, ph = (g / STAGES) & 1u, as = g % A_RING, aph = (g / A_RING) & 1u;
        ptx::mbar_wait(&full[s], ph);                          // W2 chunk landed
        ptx::mbar_wait(&a_full[as], aph);                      // A chunk written
        ptx::tc_fence_after();
        const uint64_t db = ptx::make_smem_desc_sw128(ptx::smem_u32(sStage + s * C::STAGE_BYTES + HALO_BYTES));
        const uint32_t ta = tmem_base + TMEM_A + as * A_COLS;
        if (ptx::elect_one()) {
#pragma unroll
          for (int k = 0; k < 4; ++k)                          // 16 channels per MMA: A advances 8 cells, B 32 bytes in the swizzled row
            ptx::umma_f16_ts(tmem_base, ta + 8u * k, db + 2u * k, idesc, (kb | k) != 0 ? 1u : 0u);
          ptx::umma_commit(&a_empty[as]);
          ptx::umma_commit(&empty[s]);
          if (kb == nkb - 1) ptx::umma_commit(acc_full);
        }
        __syncwarp();
      }
    }
  } else {
    // ===================== compute (depthwise conv + GELU -> A chunks) and epilogue =====================
    const int wq = warp & 3, cq = warp >> 2;                   // TMEM lane quarter; piece pair / accumulator column group
    const int row = wq * 32 + lane;                            // token of the tile = TMEM lane
    const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(wq * 32) << 16);
    const int py = row >> p.tw_shift, px = row & (p.tw - 1);   // pixel of the patch
    const int hw = p.tw + 2;                                   // halo row pitch (pixels)
    float* stg = sStg + warp * (C::STG_BYTES / 4);
    float* lnb = sLn + wq * (2 * 4 * 32);
    const int bar_rows = 1 + wq;
    constexpr int LPR = CW / 4, RPP = 32 / LPR, NP = 32 / RPP; // lanes per row after the transpose, rows per pass, passes
    const int rsub = lane / LPR, csub = (lane % LPR) * 4;
    const int col = cq * CW + csub;                            // this lane's 4 output columns after the transpose

    // byte offset of every (piece, tap) of this thread inside a halo tile: constants of the thread
    int toff[2][9];
#pragma unroll
    for (int pc = 0; pc < 2; ++pc)
#pragma unroll
      for (int tap = 0; tap < 9; ++tap) {
        const int hr = (py + tap / 3) * hw + px + tap % 3;     // halo row of this tap
        toff[pc][tap] = hr * ROWB + (((2 * cq + pc) ^ (hr & 7)) << 4);
      }
    const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.b2 + col));

    uint32_t g = 0, ti = 0;
    for (int t = blockIdx.x; t < p.n_tiles; t += gridDim.x, ++ti) {
      const int img = t / (p.tiles_x * p.tiles_y), r = t - img * (p.tiles_x * p.tiles_y);
      const int tyi = r / p.tiles_x, txi = r - tyi * p.tiles_x;
      const int x0 = txi * p.tw, y0 = tyi * p.th;
      for (int kb = 0; kb < nkb; ++kb, ++g) {
        const uint32_t s = g % STAGES, ph = (g / STAGES) & 1u, as = g % A_RING, aph = (g / A_RING) & 1u;
        const uint8_t* halo = sStage + s * C::STAGE_BYTES;
        ptx::mbar_wait(&full[s], ph);
        uint32_t hh[8];
#pragma unroll
        for (int pc = 0; pc < 2; ++pc) {
          const int ch = kb * 64 + (2 * cq + pc) * 8;          // first channel of this 8-channel piece
          float acc[8];
          {
            const float4 b0 = *reinterpret_cast<const float4*>(sDb + ch), b1 = *reinterpret_cast<const float4*>(sDb + ch + 4);
            acc[0] = b0.x; acc[1] = b0.y; acc[2] = b0.z; acc[3] = b0.w; acc[4] = b1.x; acc[5] = b1.y; acc[6] = b1.z; acc[7] = b1.w;
          }
#pragma unroll
          for (int ky = 0; ky < 3; ++ky) {
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) {
              const uint4 v = *reinterpret_cast<const uint4*>(halo + toff[pc][ky * 3 + kx]);
              const uint4 w = *reinterpret_cast<const uint4*>(sDw + (ky * 3 + kx) * HD_MAX + ch);
              fhfma2_mul(v.x, w.x, acc[0], acc[1]);
              fhfma2_mul(v.y, w.y, acc[2], acc[3]);
              fhfma2_mul(v.z, w.z, acc[4], acc[5]);
              fhfma2_mul(v.w, w.w, acc[6], acc[7]);
            }
          }
#pragma unroll
          for (int e = 0; e < 4; ++e) hh[pc * 4 + e] = pack_half2(gelu_erf(acc[2 * e]), gelu_erf(acc[2 * e + 1]));
        }
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&empty[s]);            // halo of this k-block has been read
        ptx::mbar_wait(&a_empty[as], aph ^ 1u);
        ptx::tc_fence_after();
        ptx::tmem_st_32x32b_x8(lane_addr + TMEM_A + as * A_COLS + cq * 8, hh);
        ptx::tmem_st_wait();
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&a_full[as]);
      }

      // ---- epilogue: residual rows of this lane (fetched before the accumulator is complete)
      float4 r4[NP];
      int64_t moff[NP];
#pragma unroll
      for (int i = 0; i < NP; ++i) {
        const int rr = wq * 32 + rsub + RPP * i;
        const int y = y0 + (rr >> p.tw_shift), x = x0 + (rr & (p.tw - 1));
        const bool ok = y < p.H && x < p.W;
        moff[i] = ok ? ((static_cast<int64_t>(img) * p.H + y) * p.W + x) * N + col : -1;
        r4[i] = ok ? *reinterpret_cast<const float4*>(p.residual + moff[i]) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
      ptx::mbar_wait(acc_full, ti & 1u);
      ptx::tc_fence_after();
      uint32_t v[CW];
      if (CW == 32) ptx::tmem_ld_32x32b_x32(lane_addr + cq * CW, v);
      else ptx::tmem_ld_32x32b_x16(lane_addr + cq * CW, v);
      ptx::tmem_ld_wait();
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(acc_empty);
#pragma unroll
      for (int j = 0; j < CW / 4; ++j)
        *reinterpret_cast<float4*>(stg + lane * STG_LD + 4 * j) =
            make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]), __uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3]));
      __syncwarp();
      float4 av[NP];
#pragma unroll
      for (int i = 0; i < NP; ++i) {
        float4 a = *reinterpret_cast<const float4*>(stg + (rsub + RPP * i) * STG_LD + csub);
        a.x += b4.x + r4[i].x; a.y += b4.y + r4[i].y; a.z += b4.z + r4[i].z; a.w += b4.w + r4[i].w;
        av[i] = a;
        if (moff[i] >= 0 && p.out32 != nullptr) *reinterpret_cast<float4*>(p.out32 + moff[i]) = a;
      }
      __syncwarp();                                            // transpose buffer is re-used by the next tile
      if (p.ln_out != nullptr) {
        // ---- LayerNorm over the N columns of each row: a row lives in LPR lanes of each of the four warps of this quarter
        const float invN = 1.f / static_cast<float>(N);
        float* bsum = lnb;                                     // [4 column groups][32 rows]
        float* bsq = lnb + 4 * 32;
        float mean[NP];
#pragma unroll
        for (int i = 0; i < NP; ++i) {
          float sacc = (av[i].x + av[i].y) + (av[i].z + av[i].w);
#pragma unroll
          for (int o = 1; o < LPR; o <<= 1) sacc += __shfl_xor_sync(0xffffffffu, sacc, o);
          if ((lane % LPR) == 0) bsum[cq * 32 + rsub + RPP * i] = sacc;
        }
        asm volatile("bar.sync %0, 128;" ::"r"(bar_rows) : "memory");
#pragma unroll
        for (int i = 0; i < NP; ++i) {
          const int rl = rsub + RPP * i;
          mean[i] = ((bsum[rl] + bsum[32 + rl]) + (bsum[64 + rl] + bsum[96 + rl])) * invN;
          const float dx = av[i].x - mean[i], dy = av[i].y - mean[i], dz = av[i].z - mean[i], dw = av[i].w - mean[i];
          float sacc = (dx * dx + dy * dy) + (dz * dz + dw * dw);
#pragma unroll
          for (int o = 1; o < LPR; o <<= 1) sacc += __shfl_xor_sync(0xffffffffu, sacc, o);
          if ((lane % LPR) == 0) bsq[cq * 32 + rl] = sacc;
        }
        asm volatile("bar.sync %0, 128;" ::"r"(bar_rows) : "memory");
        const float4 g4 = __ldg(reinterpret_cast<const float4*>(p.gamma + col)), be4 = __ldg(reinterpret_cast<const float4*>(p.beta + col));
#pragma unroll
        for (int i = 0; i < NP; ++i) {
          const int rl = rsub + RPP * i;
          const float rs = rsqrtf(((bsq[rl] + bsq[32 + rl]) + (bsq[64 + rl] + bsq[96 + rl])) * invN + p.eps);
          uint2 h;
          h.x = pack_half2((av[i].x - mean[i]) * rs * g4.x + be4.x, (av[i].y - mean[i]) * rs * g4.y + be4.y);
          h.y = pack_half2((av[i].z - mean[i]) * rs * g4.z + be4.z, (av[i].w - mean[i]) * rs * g4.w + be4.w);
          if (moff[i] >= 0) *reinterpret_cast<uint2*>(p.ln_out + moff[i]) = h;
        }
        asm volatile("bar.sync %0, 128;" ::"r"(bar_rows) : "memory");   // the exchange buffers are re-used by the next tile
      }
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == CW_WARPS + 1) ptx::tmem_dealloc(tmem_base, 256);
}

template <int N>
int launch_ffn(const CUtensorMap& tmH, const CUtensorMap& tmW, const FfnParams& p, cudaStream_t st) {
  if (const int rc = set_dyn_smem<mixffn_tail_kernel<N>>(FfnCfg<N>::SMEM, "mixffn_tail")) return rc;
  const int grid = p.n_tiles < num_sms() ? p.n_tiles : num_sms();
  launch_k(mixffn_tail_kernel<N>, grid, FFN_THREADS, FfnCfg<N>::SMEM, st, tmH, tmW, p);
  return launch_status("mixffn_tail_kernel");
}

}  // namespace
}  // namespace cffm

extern "C" int cffm_mixffn_tail_supported(int N, int HD) {
  return (N == 64 || N == 128) && HD % 64 == 0 && HD >= 64 && HD <= cffm::HD_MAX ? 1 : 0;
}

extern "C" int cffm_mixffn_tail(const void* hidden, int n, int H, int W, int HD, const void* dw_w, const float* dw_b, const void* W2,
                                int64_t ldw2, const float* b2, const float* residual, float* out32, const float* ln_gamma,
                                const float* ln_beta, float ln_eps, void* ln_out16, int N, void* stream) {
  using namespace cffm;
  CFFM_REQUIRE(hidden && dw_w && dw_b && W2 && b2 && residual, CFFM_E_BADARG, "mixffn_tail: null pointer");
  CFFM_REQUIRE(out32 || ln_out16, CFFM_E_BADARG, "mixffn_tail: no output requested");
  CFFM_REQUIRE(!ln_out16 || (ln_gamma && ln_beta), CFFM_E_BADARG, "mixffn_tail: LayerNorm output without gamma / beta");
  CFFM_REQUIRE(n > 0 && H > 0 && W > 0, CFFM_E_BADARG, "mixffn_tail: non-positive size");
  CFFM_REQUIRE(cffm_mixffn_tail_supported(N, HD), CFFM_E_UNSUPPORTED,
               "mixffn_tail: built for N in {64, 128} output channels and hidden %% 64 == 0, <= %d (got N=%d hidden=%d)", HD_MAX, N, HD);
  CFFM_REQUIRE(aligned16(hidden) && aligned16(dw_w) && aligned16(dw_b) && aligned16(W2) && aligned16(b2) && aligned16(residual) &&
                   aligned16(out32) && aligned16(ln_out16) && aligned16(ln_gamma) && aligned16(ln_beta) && ldw2 % 8 == 0,
               CFFM_E_BADARG, "mixffn_tail: misaligned pointer or stride");
  CFFM_REQUIRE(static_cast<int64_t>(n) * H * W < (1ll << 31) / 128, CFFM_E_UNSUPPORTED, "mixffn_tail: tensor too large for 32-bit tile indexing");
  FfnParams p;
  p.dw_w = static_cast<const __half*>(dw_w); p.dw_b = dw_b; p.b2 = b2; p.residual = residual; p.out32 = out32;
  p.gamma = ln_gamma; p.beta = ln_beta; p.eps = ln_eps; p.ln_out = static_cast<__half*>(ln_out16);
  p.n = n; p.H = H; p.W = W; p.HD = HD;
  // 4 x 32 patches unless the map is narrow (then 8 x 16): fewer empty columns in the last patch of a row
  const int waste32 = (W + 31) / 32 * 32 - W, waste16 = (W + 15) / 16 * 16 - W;
  if (waste16 * 2 < waste32 || W <= 16) { p.th = 8; p.tw = 16; p.tw_shift = 4; }
  else { p.th = 4; p.tw = 32; p.tw_shift = 5; }
  p.tiles_x = (W + p.tw - 1) / p.tw; p.tiles_y = (H + p.th - 1) / p.th;
  p.n_tiles = n * p.tiles_x * p.tiles_y;
  p.halo_bytes = (p.th + 2) * (p.tw + 2) * ROWB;
  CUtensorMap tmH, tmW;
  {
    const int64_t dims[4] = {HD, W, H, n}, strides[3] = {HD, static_cast<int64_t>(W) * HD, static_cast<int64_t>(H) * W * HD};
    const int box[4] = {64, p.tw + 2, p.th + 2, 1};
    int rc = make_tmap_4d(&tmH, hidden, dims, strides, box);
    if (rc) return rc;
  }
  int rc = make_tmap(&tmW, W2, N, HD, ldw2, N);
  if (rc) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  return N == 64 ? launch_ffn<64>(tmH, tmW, p, st) : launch_ffn<128>(tmH, tmW, p, st);
}
