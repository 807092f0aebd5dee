// Stage-1 OverlapPatchEmbed in one kernel: 7x7 stride-4 convolution of the fp32 NCHW frames + bias, LayerNorm (the patch
// embedding's own norm) -> fp32 residual stream, LayerNorm of the first block (norm1) -> fp16
// (OverlapPatchEmbed.forward, mix_transformer.py:173-200, followed by Block.norm1 :154).  No patch matrix is materialised.
//
// Persistent CTAs loop over 128-token tiles = 4 x 32 output pixels of one frame:
//   TMA      the tile's input window [3 channels x 19 rows x 136 columns] fp32 (4-D map over the NCHW frames, no swizzle; the
//            out-of-bounds zero fill IS the convolution's zero padding), 3-stage ring; the re-ordered weights W'[N x 192]
//            (K index = (channel * 7 + ky) * 8 + kx, kx = 7 and the rows past 21 are zero) once per CTA
//   convert  16 warps, thread = (token, two (channel, ky) rows of the k-block): three 16-byte shared-memory loads per row (the 8
//            input pixels under the filter row are contiguous), fp32 -> fp16, written into TENSOR MEMORY (tcgen05.st): the
//            thread = token layout is the A-operand layout of a TS-mode MMA
//   tcgen05  acc[128 x N] += A_chunk[128 x 64] (TMEM) . W'_chunk^T (shared memory), three chunks per tile
//   epilogue tcgen05.ld -> per-warp transpose -> + bias -> LayerNorm -> fp32 rows -> LayerNorm -> fp16 rows
// Warp roles: 0..15 convert / epilogue, 16 TMA producer, 17 TMEM allocator + MMA issuer.
#include <cuda.h>

#include "common.cuh"
#include "ptx_sm100.cuh"

namespace cffm {
namespace {

constexpr int ROWB = 128;
constexpr int PE_WARPS = 16, PE_THREADS = (PE_WARPS + 2) * 32;
constexpr int KS = 7, STRIDE = 4, PAD = 3, CIN = 3;            // the only OverlapPatchEmbed of stage 1 (mix_transformer.py:290)
constexpr int TH = 4, TW = 32;                                 // output pixels per tile
// A TMA box must start on a 16-byte boundary of the innermost dimension (measured on B200, tools/micro/tma_f32_box.cu: an fp32
// box at x = -3 is an illegal instruction, at x = -4 it loads): the window starts one pixel left of the first filter tap, at
// 4 x0 - 4, and is 4 * 31 + 1 + 8 = 133 -> 136 columns wide.
constexpr int IN_ROWS = STRIDE * (TH - 1) + KS, IN_COLS = 136, IN_X0 = PAD + 1;
constexpr int IN_BYTES = CIN * IN_ROWS * IN_COLS * 4;          // 31008
constexpr int IN_STAGE = (IN_BYTES + 127) / 128 * 128;
constexpr int STAGES = 3, A_RING = 4, A_COLS = 32, NKB = 3;    // K = 21 rows x 8 = 168 -> 192 = 3 chunks of 64
constexpr int TMEM_A = 64;                                     // accumulator [0, N <= 64) | A ring [64, 192)

template <int N>
struct PeCfg {
  static constexpr int CW = N / 4;                             // accumulator columns per epilogue warp: 8 or 16
  static constexpr int STG_LD = CW + 4;
  static constexpr int STG_BYTES = 32 * STG_LD * 4;
  static constexpr int W_BYTES = NKB * N * ROWB;               // three [N x 64] chunks, 128-byte swizzle
  static constexpr int LN_BYTES = 4 * 2 * 4 * 32 * 4;
  static constexpr int SMEM = W_BYTES + STAGES * IN_STAGE + PE_WARPS * STG_BYTES + LN_BYTES + 256 + 1024;
  static_assert(SMEM <= 232448, "shared memory budget");
};

struct PeParams {
  const float* bias;                          // [N]
  const float *g1, *b1, *g2, *b2;             // LayerNorm of the embedding, LayerNorm of the first block
  float eps1, eps2;
  float* out32;                               // [M, N] = LayerNorm_1(conv + bias)
  __half* ln_out;                             // [M, N] = LayerNorm_2(out32)
  int n, H, W, Ho, Wo;
  int tiles_x, tiles_y, n_tiles;
};

template <int N>
__global__ void __launch_bounds__(PE_THREADS, 1)
patch_embed_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW, const PeParams p) {
  using C = PeCfg<N>;
  constexpr int CW = C::CW, STG_LD = C::STG_LD;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = ptx::smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);
  uint8_t* sW = smem;                                          // 1024-byte aligned chunks
  uint8_t* sIn = sW + C::W_BYTES;
  float* sStg = reinterpret_cast<float*>(sIn + STAGES * IN_STAGE);
  float* sLn = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(sStg) + PE_WARPS * C::STG_BYTES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(sLn) + C::LN_BYTES);
  uint64_t* full = bars + 0;        // [STAGES] TMA -> convert warps: input window landed
  uint64_t* empty = bars + 3;       // [STAGES] 16 convert warps -> TMA
  uint64_t* a_full = bars + 6;      // [A_RING] convert warps -> MMA issuer
  uint64_t* a_empty = bars + 10;    // [A_RING] MMA commit -> convert warps
  uint64_t* acc_full = bars + 14;
  uint64_t* acc_empty = bars + 15;
  uint64_t* w_full = bars + 16;     // TMA -> MMA issuer: weights landed (once)
  uint32_t* tmem_base_smem = reinterpret_cast<uint32_t*>(bars + 17);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == PE_WARPS && lane == 0) {
    ptx::prefetch_tensormap(&tmX);
    ptx::prefetch_tensormap(&tmW);
    for (int i = 0; i < STAGES; ++i) { ptx::mbar_init(&full[i], 1); ptx::mbar_init(&empty[i], PE_WARPS); }
    for (int i = 0; i < A_RING; ++i) { ptx::mbar_init(&a_full[i], PE_WARPS); ptx::mbar_init(&a_empty[i], 1); }
    ptx::mbar_init(acc_full, 1);
    ptx::mbar_init(acc_empty, PE_WARPS);
    ptx::mbar_init(w_full, 1);
    ptx::fence_barrier_init();
  }
  if (warp == PE_WARPS + 1) {
    ptx::tmem_alloc(tmem_base_smem, 256);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_base_smem;
  pdl_sync();

  if (warp == PE_WARPS) {
    // ===================== TMA producer =====================
    if (ptx::elect_one()) {                                    // the weights: constants of the layer
      ptx::mbar_arrive_expect_tx(w_full, C::W_BYTES);
      for (int kb = 0; kb < NKB; ++kb) ptx::tma_load_2d(sW + kb * N * ROWB, &tmW, w_full, kb * 64, 0);
    }
    __syncwarp();
    uint32_t g = 0;
    for (int t = blockIdx.x; t < p.n_tiles; t += gridDim.x, ++g) {
      const int img = t / (p.tiles_x * p.tiles_y), r = t - img * (p.tiles_x * p.tiles_y);
      const int ty = r / p.tiles_x, tx = r - ty * p.tiles_x;
      const uint32_t s = g % STAGES, ph = (g / STAGES) & 1u;
      ptx::mbar_wait(&empty[s], ph ^ 1u);
      if (ptx::elect_one()) {
        ptx::mbar_arrive_expect_tx(&full[s], IN_BYTES);
        ptx::tma_load_4d(sIn + s * IN_STAGE, &tmX, &full[s], STRIDE * tx * TW - IN_X0, STRIDE * ty * TH - PAD, 0, img);
      }
      __syncwarp();
    }
  } else if (warp == PE_WARPS + 1) {
    // ===================== MMA issuer =====================
    constexpr uint32_t idesc = ptx::make_idesc_f16(128, N);
    ptx::mbar_wait(w_full, 0);
    uint32_t gc = 0, ti = 0;
    for (int t = blockIdx.x; t < p.n_tiles; t += gridDim.x, ++ti) {
      ptx::mbar_wait(acc_empty, (ti & 1u) ^ 1u);
      ptx::tc_fence_after();
      for (int kb = 0; kb < NKB; ++kb, ++gc) {
        const uint32_t as = gc % A_RING, aph = (gc / A_RING) & 1u;
        ptx::mbar_wait(&a_full[as], aph);
        ptx::tc_fence_after();
        const uint64_t db = ptx::make_smem_desc_sw128(ptx::smem_u32(sW + kb * N * ROWB));
        const uint32_t ta = tmem_base + TMEM_A + as * A_COLS;
        if (ptx::elect_one()) {
#pragma unroll
          for (int k = 0; k < 4; ++k) ptx::umma_f16_ts(tmem_base, ta + 8u * k, db + 2u * k, idesc, (kb | k) != 0 ? 1u : 0u);
          ptx::umma_commit(&a_empty[as]);
          if (kb == NKB - 1) ptx::umma_commit(acc_full);
        }
        __syncwarp();
      }
    }
  } else {
    // ===================== convert (fp32 window -> fp16 A chunks) and epilogue =====================
    const int wq = warp & 3, cq = warp >> 2;                   // TMEM lane quarter = tile row; (channel, ky) pair / column group
    const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(wq * 32) << 16);
    float* stg = sStg + warp * (C::STG_BYTES / 4);
    float* lnb = sLn + wq * (2 * 4 * 32);
    const int bar_rows = 1 + wq;
    constexpr int LPR = CW / 4, RPP = 32 / LPR, NP = 32 / RPP;
    const int rsub = lane / LPR, csub = (lane % LPR) * 4;
    const int col = cq * CW + csub;
    const float4 bias4 = __ldg(reinterpret_cast<const float4*>(p.bias + col));
    const float4 g1 = __ldg(reinterpret_cast<const float4*>(p.g1 + col)), e1 = __ldg(reinterpret_cast<const float4*>(p.b1 + col));
    const float4 g2 = __ldg(reinterpret_cast<const float4*>(p.g2 + col)), e2 = __ldg(reinterpret_cast<const float4*>(p.b2 + col));
    const float invN = 1.f / static_cast<float>(N);

    uint32_t g = 0, gc = 0, ti = 0;
    for (int t = blockIdx.x; t < p.n_tiles; t += gridDim.x, ++ti, ++g) {
      const int img = t / (p.tiles_x * p.tiles_y), r = t - img * (p.tiles_x * p.tiles_y);
      const int tyi = r / p.tiles_x, txi = r - tyi * p.tiles_x;
      const int x0 = txi * TW, y0 = tyi * TH;
      const uint32_t s = g % STAGES, ph = (g / STAGES) & 1u;
      const float* win = reinterpret_cast<const float*>(sIn + s * IN_STAGE);
      ptx::mbar_wait(&full[s], ph);
#pragma unroll
      for (int kb = 0; kb < NKB; ++kb, ++gc) {
        const uint32_t as = gc % A_RING, aph = (gc / A_RING) & 1u;
        uint32_t hh[8];
#pragma unroll
        for (int pc = 0; pc < 2; ++pc) {
          const int pr = kb * 8 + cq * 2 + pc;                 // (channel, ky) row of the filter: K columns [8 pr, 8 pr + 8)
          if (pr < CIN * KS) {                                 // warp-uniform
            const int c = pr / KS, ky = pr - c * KS;
            const float* src = win + (c * IN_ROWS + STRIDE * wq + ky) * IN_COLS + STRIDE * lane;
            // the 8 pixels under the filter row start at column 4 lane + 1 of the (16-byte aligned) window: three 16-byte loads
            const float4 v0 = *reinterpret_cast<const float4*>(src), v1 = *reinterpret_cast<const float4*>(src + 4),
                         v2 = *reinterpret_cast<const float4*>(src + 8);
            hh[pc * 4 + 0] = pack_half2(v0.y, v0.z); hh[pc * 4 + 1] = pack_half2(v0.w, v1.x);
            hh[pc * 4 + 2] = pack_half2(v1.y, v1.z); hh[pc * 4 + 3] = pack_half2(v1.w, v2.x);   // v2.x (kx = 7) meets a zero weight
          } else {
            hh[pc * 4 + 0] = hh[pc * 4 + 1] = hh[pc * 4 + 2] = hh[pc * 4 + 3] = 0u;
          }
        }
        if (kb == NKB - 1) {                                   // the window has been read for the last time
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive(&empty[s]);
        }
        ptx::mbar_wait(&a_empty[as], aph ^ 1u);
        ptx::tc_fence_after();
        ptx::tmem_st_32x32b_x8(lane_addr + TMEM_A + as * A_COLS + cq * 8, hh);
        ptx::tmem_st_wait();
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&a_full[as]);
      }

      // ---- epilogue
      int64_t moff[NP];
#pragma unroll
      for (int i = 0; i < NP; ++i) {
        const int rr = wq * 32 + rsub + RPP * i;
        const int y = y0 + (rr >> 5), x = x0 + (rr & 31);
        moff[i] = (y < p.Ho && x < p.Wo) ? ((static_cast<int64_t>(img) * p.Ho + y) * p.Wo + x) * N + col : -1;
      }
      ptx::mbar_wait(acc_full, ti & 1u);
      ptx::tc_fence_after();
      uint32_t v[CW];
      if (CW == 16) ptx::tmem_ld_32x32b_x16(lane_addr + cq * CW, v);
      else ptx::tmem_ld_32x32b_x8(lane_addr + cq * CW, v);
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
        a.x += bias4.x; a.y += bias4.y; a.z += bias4.z; a.w += bias4.w;
        av[i] = a;
      }
      __syncwarp();
      float* bsum = lnb;
      float* bsq = lnb + 4 * 32;
      float mean[NP];
      // row statistics of av[] over the N columns: LPR lanes of each of the four warps of this quarter hold a row
      auto row_stats = [&]() {
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
      };
      auto rstd = [&](int i, float eps) {
        const int rl = rsub + RPP * i;
        return rsqrtf(((bsq[rl] + bsq[32 + rl]) + (bsq[64 + rl] + bsq[96 + rl])) * invN + eps);
      };
      row_stats();
#pragma unroll
      for (int i = 0; i < NP; ++i) {                           // y = LayerNorm_1(x): the fp32 residual stream
        const float rs = rstd(i, p.eps1);
        av[i].x = (av[i].x - mean[i]) * rs * g1.x + e1.x; av[i].y = (av[i].y - mean[i]) * rs * g1.y + e1.y;
        av[i].z = (av[i].z - mean[i]) * rs * g1.z + e1.z; av[i].w = (av[i].w - mean[i]) * rs * g1.w + e1.w;
        if (moff[i] >= 0) *reinterpret_cast<float4*>(p.out32 + moff[i]) = av[i];
      }
      asm volatile("bar.sync %0, 128;" ::"r"(bar_rows) : "memory");   // everyone has read the statistics of the first norm
      row_stats();
#pragma unroll
      for (int i = 0; i < NP; ++i) {                           // LayerNorm_2(y): the first block's norm1, fp16
        const float rs = rstd(i, p.eps2);
        uint2 h;
        h.x = pack_half2((av[i].x - mean[i]) * rs * g2.x + e2.x, (av[i].y - mean[i]) * rs * g2.y + e2.y);
        h.y = pack_half2((av[i].z - mean[i]) * rs * g2.z + e2.z, (av[i].w - mean[i]) * rs * g2.w + e2.w);
        if (moff[i] >= 0) *reinterpret_cast<uint2*>(p.ln_out + moff[i]) = h;
      }
      asm volatile("bar.sync %0, 128;" ::"r"(bar_rows) : "memory");   // the exchange buffers are re-used by the next tile
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == PE_WARPS + 1) ptx::tmem_dealloc(tmem_base, 256);
}

template <int N>
int launch_pe(const CUtensorMap& tmX, const CUtensorMap& tmW, const PeParams& p, cudaStream_t st) {
  if (const int rc = set_dyn_smem<patch_embed_kernel<N>>(PeCfg<N>::SMEM, "patch_embed_s1")) return rc;
  const int grid = p.n_tiles < num_sms() ? p.n_tiles : num_sms();
  launch_k(patch_embed_kernel<N>, grid, PE_THREADS, PeCfg<N>::SMEM, st, tmX, tmW, p);
  return launch_status("patch_embed_kernel");
}

}  // namespace
}  // namespace cffm

extern "C" int cffm_patch_embed_s1_supported(int W, int Cin, int ksize, int stride, int pad, int Nout) {
  return (W % 4 == 0 && Cin == cffm::CIN && ksize == cffm::KS && stride == cffm::STRIDE && pad == cffm::PAD && (Nout == 32 || Nout == 64)) ? 1 : 0;
}

extern "C" int cffm_patch_embed_s1(const float* x, int n, int H, int W, const void* Wk, const float* bias, const float* gamma1,
                                   const float* beta1, float eps1, const float* gamma2, const float* beta2, float eps2,
                                   float* out_f32, void* ln_out_f16, int Nout, void* stream) {
  using namespace cffm;
  CFFM_REQUIRE(x && Wk && bias && gamma1 && beta1 && gamma2 && beta2 && out_f32 && ln_out_f16, CFFM_E_BADARG, "patch_embed_s1: null pointer");
  CFFM_REQUIRE(n > 0 && H >= KS && W >= KS, CFFM_E_BADARG, "patch_embed_s1: bad size");
  CFFM_REQUIRE(cffm_patch_embed_s1_supported(W, CIN, KS, STRIDE, PAD, Nout), CFFM_E_UNSUPPORTED,
               "patch_embed_s1: built for 3 input channels, k7 s4 p3, 32 or 64 output channels and W %% 4 == 0 (got W=%d N=%d)", W, Nout);
  CFFM_REQUIRE(aligned16(x) && aligned16(Wk) && aligned16(bias) && aligned16(gamma1) && aligned16(beta1) && aligned16(gamma2) &&
                   aligned16(beta2) && aligned16(out_f32) && aligned16(ln_out_f16), CFFM_E_BADARG, "patch_embed_s1: misaligned pointer");
  PeParams p;
  p.bias = bias; p.g1 = gamma1; p.b1 = beta1; p.g2 = gamma2; p.b2 = beta2; p.eps1 = eps1; p.eps2 = eps2;
  p.out32 = out_f32; p.ln_out = static_cast<__half*>(ln_out_f16);
  p.n = n; p.H = H; p.W = W;
  p.Ho = (H + 2 * PAD - KS) / STRIDE + 1; p.Wo = (W + 2 * PAD - KS) / STRIDE + 1;
  p.tiles_x = (p.Wo + TW - 1) / TW; p.tiles_y = (p.Ho + TH - 1) / TH;
  const int64_t tiles = static_cast<int64_t>(n) * p.tiles_x * p.tiles_y;
  CFFM_REQUIRE(tiles < (1ll << 31), CFFM_E_UNSUPPORTED, "patch_embed_s1: too many tiles");
  p.n_tiles = static_cast<int>(tiles);
  CUtensorMap tmX, tmW;
  {
    const int64_t dims[4] = {W, H, CIN, n}, strides[3] = {W, static_cast<int64_t>(H) * W, static_cast<int64_t>(CIN) * H * W};
    const int box[4] = {IN_COLS, IN_ROWS, CIN, 1};
    int rc = make_tmap_f32_4d(&tmX, x, dims, strides, box);
    if (rc) return rc;
  }
  int rc = make_tmap(&tmW, Wk, Nout, NKB * 64, NKB * 64, Nout);
  if (rc) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  return Nout == 64 ? launch_pe<64>(tmX, tmW, p, st) : launch_pe<32>(tmX, tmW, p, st);
}
