// HBM-bound kernels of the CFFM path: LayerNorm, patch extraction, depthwise conv + GELU, the
// folded MLP-decoder fuse, coarse-to-fine feature assembling (CFFA) and the bilinear tails.
// All use 128-bit (8 x fp16 / 4 x fp32) accesses along the contiguous channel dimension.
#include "common.cuh"

namespace cffm {
namespace {

// PyTorch's align_corners=False source index (aten UpSample.h area_pixel_compute_source_index).
struct Lerp {
  int i0, i1;
  float w0, w1;
};
__device__ __forceinline__ Lerp lerp_coord(int dst, float scale, int in_size) {
  float src = scale * (static_cast<float>(dst) + 0.5f) - 0.5f;
  if (src < 0.f) src = 0.f;
  Lerp l;
  l.i0 = min(static_cast<int>(src), in_size - 1);
  l.i1 = l.i0 + (l.i0 < in_size - 1 ? 1 : 0);
  l.w1 = src - static_cast<float>(l.i0);
  l.w0 = 1.f - l.w1;
  return l;
}

// ------------------------------------------------------------------------------------------------
// LayerNorm.  A row of C channels is C/4 four-element vectors handled by a group of L lanes (L = the
// power of two >= C/4, capped at 32), NV vectors per lane kept in registers; 32/L rows per warp, two-pass
// fp32 statistics with intra-group shuffles.  All global accesses are 16-byte (fp32) / 8-byte (fp16).
template <bool IN_F32, int L, int NV>
__global__ void __launch_bounds__(256)
layernorm_kernel(const void* __restrict__ xin, int64_t ldx, const float* __restrict__ gamma,
                 const float* __restrict__ beta, float eps, __half* __restrict__ o16, int64_t ldo16,
                 float* __restrict__ o32, int64_t ldo32, int M, int C, int nsum, int64_t sum_stride,
                 const float* __restrict__ sum_bias, const float* __restrict__ gamma2 = nullptr,
                 const float* __restrict__ beta2 = nullptr, float eps2 = 0.f) {
  // gamma2 != nullptr: two chained LayerNorms in one pass -- o32 receives y = LN(x), o16 receives LN2(y)
  // (patch-embed norm followed by the first block's norm1: the row never leaves the registers in between)
  pdl_sync();
  constexpr int RPW = 32 / L;                                  // rows per warp
  const int lane = threadIdx.x & 31;
  const int gl = lane % L;                                     // lane inside the row group
  const int nvec = C >> 2;
  const float invC = 1.f / static_cast<float>(C);
  float4 g[NV], b[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int v = gl + i * L;
    g[i] = v < nvec ? *reinterpret_cast<const float4*>(gamma + 4 * v) : make_float4(0.f, 0.f, 0.f, 0.f);
    b[i] = v < nvec ? *reinterpret_cast<const float4*>(beta + 4 * v) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  const int64_t warp_global = (blockIdx.x * 256ll + threadIdx.x) >> 5;
  const int64_t nwarps = (gridDim.x * 256ll) >> 5;
  for (int64_t rb = warp_global * RPW; rb < M; rb += nwarps * RPW) {      // warp-uniform trip count (shuffles inside)
    const int64_t row = rb + lane / L;
    const bool rv = row < M;
    float4 x[NV];
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int v = gl + i * L;
      x[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (rv && v < nvec) {
        if (IN_F32) {
          const float* px = static_cast<const float*>(xin) + row * ldx + 4 * v;
          x[i] = *reinterpret_cast<const float4*>(px);
          for (int sp = 1; sp < nsum; ++sp) {                   // split-K partial sums (fixed order: deterministic)
            const float4 t = *reinterpret_cast<const float4*>(px + sp * sum_stride);
            x[i].x += t.x; x[i].y += t.y; x[i].z += t.z; x[i].w += t.w;
          }
          if (sum_bias != nullptr) {
            const float4 t = *reinterpret_cast<const float4*>(sum_bias + 4 * v);
            x[i].x += t.x; x[i].y += t.y; x[i].z += t.z; x[i].w += t.w;
          }
        } else {
          const uint2 u = *reinterpret_cast<const uint2*>(static_cast<const __half*>(xin) + row * ldx + 4 * v);
          const float2 lo = __half22float2(*reinterpret_cast<const __half2*>(&u.x));
          const float2 hi = __half22float2(*reinterpret_cast<const __half2*>(&u.y));
          x[i] = make_float4(lo.x, lo.y, hi.x, hi.y);
        }
      }
      sum += (x[i].x + x[i].y) + (x[i].z + x[i].w);
    }
#pragma unroll
    for (int o = L / 2; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const float mean = sum * invC;
    float sq = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      if (gl + i * L < nvec) {
        x[i].x -= mean; x[i].y -= mean; x[i].z -= mean; x[i].w -= mean;
        sq += (x[i].x * x[i].x + x[i].y * x[i].y) + (x[i].z * x[i].z + x[i].w * x[i].w);
      }
    }
#pragma unroll
    for (int o = L / 2; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
    const float rstd = rsqrtf(sq * invC + eps);
    if (gamma2 == nullptr) {
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const int v = gl + i * L;
        if (rv && v < nvec) {
          float4 y;
          y.x = x[i].x * rstd * g[i].x + b[i].x; y.y = x[i].y * rstd * g[i].y + b[i].y;
          y.z = x[i].z * rstd * g[i].z + b[i].z; y.w = x[i].w * rstd * g[i].w + b[i].w;
          if (o32) *reinterpret_cast<float4*>(o32 + row * ldo32 + 4 * v) = y;
          if (o16) {
            uint2 h;
            h.x = pack_half2(y.x, y.y);
            h.y = pack_half2(y.z, y.w);
            *reinterpret_cast<uint2*>(o16 + row * ldo16 + 4 * v) = h;
          }
        }
      }
    } else {
      float sum2 = 0.f;
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const int v = gl + i * L;
        if (v < nvec) {
          x[i].x = x[i].x * rstd * g[i].x + b[i].x; x[i].y = x[i].y * rstd * g[i].y + b[i].y;
          x[i].z = x[i].z * rstd * g[i].z + b[i].z; x[i].w = x[i].w * rstd * g[i].w + b[i].w;
          if (rv && o32) *reinterpret_cast<float4*>(o32 + row * ldo32 + 4 * v) = x[i];
          sum2 += (x[i].x + x[i].y) + (x[i].z + x[i].w);
        }
      }
#pragma unroll
      for (int o = L / 2; o > 0; o >>= 1) sum2 += __shfl_xor_sync(0xffffffffu, sum2, o);
      const float mean2 = sum2 * invC;
      float sq2 = 0.f;
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        if (gl + i * L < nvec) {
          x[i].x -= mean2; x[i].y -= mean2; x[i].z -= mean2; x[i].w -= mean2;
          sq2 += (x[i].x * x[i].x + x[i].y * x[i].y) + (x[i].z * x[i].z + x[i].w * x[i].w);
        }
      }
#pragma unroll
      for (int o = L / 2; o > 0; o >>= 1) sq2 += __shfl_xor_sync(0xffffffffu, sq2, o);
      const float rstd2 = rsqrtf(sq2 * invC + eps2);
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const int v = gl + i * L;
        if (rv && v < nvec) {
          const float4 g2 = *reinterpret_cast<const float4*>(gamma2 + 4 * v), b2 = *reinterpret_cast<const float4*>(beta2 + 4 * v);
          uint2 h;
          h.x = pack_half2(x[i].x * rstd2 * g2.x + b2.x, x[i].y * rstd2 * g2.y + b2.y);
          h.y = pack_half2(x[i].z * rstd2 * g2.z + b2.z, x[i].w * rstd2 * g2.w + b2.w);
          *reinterpret_cast<uint2*>(o16 + row * ldo16 + 4 * v) = h;
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// im2col, K order (ky, kx, c)
__global__ void __launch_bounds__(256)
im2col_nhwc_kernel(const __half* __restrict__ x, int N, int H, int W, int C, int k, int stride, int pad, int Ho,
                   int Wo, __half* __restrict__ A, int Kpad) {
  pdl_sync();
  // 32-bit index arithmetic (the host checks that the chunk count fits): 64-bit divisions cost more than the copy
  const uint32_t chunks = Kpad / 8;
  const uint32_t total = static_cast<uint32_t>(N) * Ho * Wo * chunks;
  const uint32_t kdim = k * k * C;
  for (uint32_t i = blockIdx.x * 256u + threadIdx.x; i < total; i += 256u * gridDim.x) {
    const uint32_t m = i / chunks, kcol = (i - m * chunks) * 8;
    uint4 val = make_uint4(0, 0, 0, 0);
    if (kcol < kdim) {
      const uint32_t tap = kcol / C, c = kcol - tap * C;
      const uint32_t ky = tap / k, kx = tap - ky * k;
      const uint32_t t = m / Wo, ox = m - t * Wo, n = t / Ho, oy = t - n * Ho;
      const int iy = static_cast<int>(oy * stride + ky) - pad, ix = static_cast<int>(ox * stride + kx) - pad;
      if (iy >= 0 && iy < H && ix >= 0 && ix < W)
        val = *reinterpret_cast<const uint4*>(x + ((static_cast<int64_t>(n) * H + iy) * W + ix) * C + c);
    }
    *reinterpret_cast<uint4*>(A + static_cast<int64_t>(m) * Kpad + kcol) = val;
  }
}

// fp32 NCHW image -> fp16 patches.  One CTA per output row (n, oy): the k input rows of every channel are
// staged in shared memory (fp32 -> fp16 once; 16-byte loads, all of a thread's loads issued before the first
// conversion so that ~10 requests per thread are in flight), then each thread assembles 16-byte
// chunks of the patch matrix, so the (dominant) output traffic is written as full 16-byte stores.
// Shared rows are [C*k][Wp] with the image starting at column PADL (= pad rounded up to 4: 8-byte aligned stores).
__global__ void __launch_bounds__(256)
im2col_nchw_f32_kernel(const float* __restrict__ x, int N, int H, int W, int C, int k, int stride, int pad, int Ho,
                       int Wo, __half* __restrict__ A, int Kpad) {
  pdl_sync();
  extern __shared__ __align__(16) uint8_t im_smem[];
  const int PADL = (pad + 3) & ~3;
  const int Wp = ((W + PADL + pad + 3) & ~3);
  const int kdim = k * k * C, chunks = Kpad / 8;
  __half* rows = reinterpret_cast<__half*>(im_smem);           // [C*k][Wp] (zero borders)
  int* lut = reinterpret_cast<int*>(im_smem + ((static_cast<size_t>(C) * k * Wp * 2 + 15) & ~static_cast<size_t>(15)));   // [Kpad]
  const int oy = blockIdx.x % Ho, n = blockIdx.x / Ho;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int kc = threadIdx.x; kc < Kpad; kc += 256) {          // patch column -> offset inside `rows` (-1: zero pad)
    int off = -1;
    if (kc < kdim) {
      const int c = kc % C, tap = kc / C, kx = tap % k, ky = tap / k;
      off = (c * k + ky) * Wp + kx + (PADL - pad);
    }
    lut[kc] = off;
  }
  const int nrows = C * k;
  if ((W & 3) == 0) {
    const int W4 = W >> 2, items = nrows * W4;
    for (int r = threadIdx.x; r < nrows; r += 256) {            // zero borders
      for (int q = 0; q < PADL; ++q) rows[r * Wp + q] = __float2half_rn(0.f);
      for (int q = PADL + W; q < Wp; ++q) rows[r * Wp + q] = __float2half_rn(0.f);
    }
    constexpr int U = 10;                                       // loads in flight per thread
    for (int i0 = threadIdx.x; i0 < items; i0 += 256 * U) {
      float4 v[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int i = i0 + u * 256;
        v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (i < items) {
          const int r = i / W4, x4 = i - r * W4;
          const int ky = r % k, c = r / k;
          const int iy = oy * stride - pad + ky;
          if (iy >= 0 && iy < H)
            v[u] = *reinterpret_cast<const float4*>(x + ((static_cast<int64_t>(n) * C + c) * H + iy) * W + 4 * x4);
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int i = i0 + u * 256;
        if (i < items) {
          const int r = i / W4, x4 = i - r * W4;
          uint2 h;
          h.x = pack_half2(v[u].x, v[u].y);
          h.y = pack_half2(v[u].z, v[u].w);
          *reinterpret_cast<uint2*>(rows + r * Wp + PADL + 4 * x4) = h;
        }
      }
    }
  } else {
    for (int r = warp; r < nrows; r += 8) {                    // one (channel, ky) input row per warp pass
      const int ky = r % k, c = r / k;
      const int iy = oy * stride - pad + ky;
      const bool rv = iy >= 0 && iy < H;
      const float* src = x + ((static_cast<int64_t>(n) * C + c) * H + (rv ? iy : 0)) * W;
      for (int xp = lane; xp < Wp; xp += 32) {
        const int xx = xp - PADL;
        rows[r * Wp + xp] = __float2half_rn(rv && xx >= 0 && xx < W ? src[xx] : 0.f);
      }
    }
  }
  __syncthreads();
  __half* Arow = A + (static_cast<int64_t>(n) * Ho + oy) * Wo * Kpad;
  for (int ox = warp; ox < Wo; ox += 8) {                      // one output pixel per warp pass, lanes over 16-byte chunks
    const __half* rbase = rows + ox * stride;
    for (int ch = lane; ch < chunks; ch += 32) {
      __align__(16) __half v[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const int off = lut[ch * 8 + e];
        v[e] = off >= 0 ? rbase[off] : __float2half_rn(0.f);
      }
      *reinterpret_cast<uint4*>(Arow + static_cast<int64_t>(ox) * Kpad + ch * 8) = *reinterpret_cast<const uint4*>(v);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// depthwise 3x3 (pad 1) + bias + exact GELU, NHWC fp16 (Mix-FFN, mix_transformer.py:20-55).
// One thread = CPT channels x a PXW-pixel-wide column strip of RS output rows.  Every input
// row of the strip is loaded ONCE ((PXW + 2) vectors) and feeds the three output rows that see it through ky = 0, 1, 2
// (three rotating fp32 accumulator rows); weights and bias live in registers for the whole strip.  ~1.4 loads and
// conversions per output instead of 4.5, 32-bit index arithmetic, loops fully unrolled (static register rotation).
template <int CPT> struct HalfVec;
template <> struct HalfVec<2> { using T = uint32_t; };
template <> struct HalfVec<4> { using T = uint2; };
template <> struct HalfVec<8> { using T = uint4; };

template <int CPT>
__device__ __forceinline__ void load_halves(const __half* __restrict__ p, float* f) {
  const typename HalfVec<CPT>::T raw = *reinterpret_cast<const typename HalfVec<CPT>::T*>(p);
  const __half2* h = reinterpret_cast<const __half2*>(&raw);
#pragma unroll
  for (int i = 0; i < CPT / 2; ++i) {
    const float2 t = __half22float2(h[i]);
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}

template <int CPT>
__device__ __forceinline__ void store_halves(__half* __restrict__ p, const float* f) {
  typename HalfVec<CPT>::T raw;
  __half2* h = reinterpret_cast<__half2*>(&raw);
#pragma unroll
  for (int i = 0; i < CPT / 2; ++i) h[i] = __floats2half2_rn(f[2 * i], f[2 * i + 1]);
  *reinterpret_cast<typename HalfVec<CPT>::T*>(p) = raw;
}

// One column strip.  INTERIOR: all (PXW + 2) x (RS + 2) inputs and all PXW x RS outputs are inside the image, so the
// strip is straight-line code -- no row / column predicates and no branches, which lets the compiler issue the loads of
// the next input row under the FMAs of the current one; CT > 0: the channel count is a compile-time constant and every
// load / store of a row is `row pointer + immediate` (integer address arithmetic was ~35 % of the issued instructions).
template <int CPT, int PXW, int RS, int CT, bool INTERIOR>
__device__ __forceinline__ void dwconv_strip(const __half* __restrict__ xin_n, __half* __restrict__ out_n,
                                             const float (&wf)[9][CPT], const float (&bs)[CPT], int x0, int y0, int H,
                                             int W, int C_rt) {
  const int C = CT > 0 ? CT : C_rt;
  const int64_t row_stride = static_cast<int64_t>(W) * C;
  const __half* rp = xin_n + (static_cast<int64_t>(y0 - 1) * W + (x0 - 1)) * C;   // input row y0 - 1, column x0 - 1
  __half* op = out_n + (static_cast<int64_t>(y0) * W + x0) * C;
  float acc[3][PXW][CPT];
#pragma unroll
  for (int r = 0; r < RS + 2; ++r, rp += row_stride) {
    const int iy = y0 - 1 + r;
    if (INTERIOR || (iy >= 0 && iy < H)) {
      float xin[PXW + 2][CPT];
#pragma unroll
      for (int j = 0; j < PXW + 2; ++j) {
        if (INTERIOR || (x0 - 1 + j >= 0 && x0 - 1 + j < W)) load_halves<CPT>(rp + j * C, xin[j]);
        else {
#pragma unroll
          for (int e = 0; e < CPT; ++e) xin[j][e] = 0.f;
        }
      }
#pragma unroll
      for (int ky = 0; ky < 3; ++ky) {
        const int q = r - ky;                                   // output row (relative) that sees this input row through ky
        if (q >= 0 && q < RS) {
#pragma unroll
          for (int px = 0; px < PXW; ++px)
#pragma unroll
            for (int kx = 0; kx < 3; ++kx)
#pragma unroll
              for (int e = 0; e < CPT; ++e) {
                // the first tap of an output row starts from the bias (row q opens with input row r = q, ky = 0, kx = 0)
                const float prev = (ky == 0 && kx == 0) ? bs[e] : acc[q % 3][px][e];
                acc[q % 3][px][e] = fmaf(xin[px + kx][e], wf[ky * 3 + kx][e], prev);
              }
        }
      }
    } else if (r < RS) {                                        // zero-padding row above the image: row q = r opens with the bias
#pragma unroll
      for (int px = 0; px < PXW; ++px)
#pragma unroll
        for (int e = 0; e < CPT; ++e) acc[r % 3][px][e] = bs[e];
    }
    if (r >= 2) {                                               // output row q = r - 2 is complete
      if (INTERIOR || y0 + r - 2 < H) {
#pragma unroll
        for (int px = 0; px < PXW; ++px) {
          if (INTERIOR || x0 + px < W) {
            float o[CPT];
#pragma unroll
            for (int e = 0; e < CPT; ++e) o[e] = gelu_erf(acc[(r - 2) % 3][px][e]);
            store_halves<CPT>(op + px * C, o);
          }
        }
      }
      op += row_stride;
    }
  }
}

template <int CPT, int PXW, int RS, int CT>
__global__ void __launch_bounds__(256)
dwconv3x3_gelu_rows_kernel(const __half* __restrict__ x, const __half* __restrict__ w, const float* __restrict__ bias,
                           __half* __restrict__ out, int N, int H, int W, int C_rt, int strips_x, int strips_y) {
  pdl_sync();
  const int C = CT > 0 ? CT : C_rt;
  const uint32_t chunks = C / CPT;
  uint32_t idx = blockIdx.x * 256u + threadIdx.x;
  if (idx >= static_cast<uint32_t>(N) * strips_y * strips_x * chunks) return;
  const uint32_t cc = idx % chunks; idx /= chunks;
  const uint32_t xs = idx % strips_x; idx /= strips_x;
  const uint32_t ys = idx % strips_y, n = idx / strips_y;
  const int c = cc * CPT, x0 = xs * PXW, y0 = ys * RS;
  float wf[9][CPT], bs[CPT];
#pragma unroll
  for (int t = 0; t < 9; ++t) load_halves<CPT>(w + t * C + c, wf[t]);
#pragma unroll
  for (int e = 0; e < CPT; e += 2) {
    const float2 b2 = *reinterpret_cast<const float2*>(bias + c + e);
    bs[e] = b2.x; bs[e + 1] = b2.y;
  }
  const __half* xin_n = x + static_cast<int64_t>(n) * H * W * C + c;
  __half* out_n = out + static_cast<int64_t>(n) * H * W * C + c;
  if (x0 >= 1 && x0 + PXW + 1 <= W && y0 >= 1 && y0 + RS + 1 <= H)
    dwconv_strip<CPT, PXW, RS, CT, true>(xin_n, out_n, wf, bs, x0, y0, H, W, C);
  else dwconv_strip<CPT, PXW, RS, CT, false>(xin_n, out_n, wf, bs, x0, y0, H, W, C);
}

// ------------------------------------------------------------------------------------------------
// c = relu(p1 + up(p2) + up(p3) + up(p4) + shift); optional full-res fp16 copy; 2x2 mean outputs.
__device__ __forceinline__ void add_bilerp(float* acc, const __half* __restrict__ p, int n, int Hs, int Ws, int C,
                                           int c, const Lerp& ly, const Lerp& lx) {
  const __half* base = p + static_cast<int64_t>(n) * Hs * Ws * C + c;
  float t[8];
  unpack8(*reinterpret_cast<const half8*>(base + (static_cast<int64_t>(ly.i0) * Ws + lx.i0) * C), t);
#pragma unroll
  for (int e = 0; e < 8; ++e) acc[e] = fmaf(ly.w0 * lx.w0, t[e], acc[e]);
  unpack8(*reinterpret_cast<const half8*>(base + (static_cast<int64_t>(ly.i0) * Ws + lx.i1) * C), t);
#pragma unroll
  for (int e = 0; e < 8; ++e) acc[e] = fmaf(ly.w0 * lx.w1, t[e], acc[e]);
  unpack8(*reinterpret_cast<const half8*>(base + (static_cast<int64_t>(ly.i1) * Ws + lx.i0) * C), t);
#pragma unroll
  for (int e = 0; e < 8; ++e) acc[e] = fmaf(ly.w1 * lx.w0, t[e], acc[e]);
  unpack8(*reinterpret_cast<const half8*>(base + (static_cast<int64_t>(ly.i1) * Ws + lx.i1) * C), t);
#pragma unroll
  for (int e = 0; e < 8; ++e) acc[e] = fmaf(ly.w1 * lx.w1, t[e], acc[e]);
}

// One warp = one 2x2 block of full-resolution pixels x 64 channels: lane = q*8 + cc, q = pixel of the block,
// cc = 8-channel chunk.  Every lane evaluates ONE pixel (1 + 3x4 16-byte loads, ~50 registers -> high
// occupancy); the 2x2 mean (= resize(_c, 1/2)) is two xor-shuffles; the q = 0 lanes write the half-res outputs.
__global__ void __launch_bounds__(256)
head_fuse_kernel(const __half* __restrict__ p1, const __half* __restrict__ p2, const __half* __restrict__ p3,
                 const __half* __restrict__ p4, int N, int H1, int W1, int H2, int W2, int H3, int W3, int H4, int W4,
                 int C, int Tperm, const float* __restrict__ shift, __half* __restrict__ c_full,
                 float* __restrict__ h32, int64_t ldh32, __half* __restrict__ h16, int64_t ldh16) {
  pdl_sync();
  const uint32_t groups = C / 64, Hh = H1 / 2, Wh = W1 / 2;
  const uint32_t total_warps = static_cast<uint32_t>(N) * Hh * Wh * groups;   // host checks that it fits 32 bits
  const float sy2 = static_cast<float>(H2) / H1, sx2 = static_cast<float>(W2) / W1;
  const float sy3 = static_cast<float>(H3) / H1, sx3 = static_cast<float>(W3) / W1;
  const float sy4 = static_cast<float>(H4) / H1, sx4 = static_cast<float>(W4) / W1;
  const int lane = threadIdx.x & 31, q = lane >> 3, cc = lane & 7;
  for (uint32_t wi = (blockIdx.x * 256u + threadIdx.x) >> 5; wi < total_warps; wi += (gridDim.x * 256u) >> 5) {
    uint32_t p = wi / groups;
    const int g = static_cast<int>(wi - p * groups);
    const int xh = static_cast<int>(p % Wh); p /= Wh;
    const int yh = static_cast<int>(p % Hh), n = static_cast<int>(p / Hh);
    // input frame n = b*T + t (reference order) -> output slot t*B + b (frame-major, targets last)
    const int no = Tperm > 1 ? (n % Tperm) * (N / Tperm) + n / Tperm : n;
    const int c = g * 64 + cc * 8;
    const int y = 2 * yh + (q >> 1), x = 2 * xh + (q & 1);
    float acc[8];
    unpack8(*reinterpret_cast<const half8*>(p1 + ((static_cast<int64_t>(n) * H1 + y) * W1 + x) * C + c), acc);
    {
      const float4 s0 = *reinterpret_cast<const float4*>(shift + c), s1 = *reinterpret_cast<const float4*>(shift + c + 4);
      acc[0] += s0.x; acc[1] += s0.y; acc[2] += s0.z; acc[3] += s0.w;
      acc[4] += s1.x; acc[5] += s1.y; acc[6] += s1.z; acc[7] += s1.w;
    }
    add_bilerp(acc, p2, n, H2, W2, C, c, lerp_coord(y, sy2, H2), lerp_coord(x, sx2, W2));
    add_bilerp(acc, p3, n, H3, W3, C, c, lerp_coord(y, sy3, H3), lerp_coord(x, sx3, W3));
    add_bilerp(acc, p4, n, H4, W4, C, c, lerp_coord(y, sy4, H4), lerp_coord(x, sx4, W4));
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] = fmaxf(acc[e], 0.f);
    if (c_full) *reinterpret_cast<half8*>(c_full + ((static_cast<int64_t>(no) * H1 + y) * W1 + x) * C + c) = pack8(acc);
    // 2x2 mean in the order the scalar kernel used: ((p00 + p01) + p10) + p11 is not needed bit-for-bit;
    // pairwise (p00 + p01) + (p10 + p11) via xor-shuffles over the two q bits
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      acc[e] += __shfl_xor_sync(0xffffffffu, acc[e], 8);
      acc[e] += __shfl_xor_sync(0xffffffffu, acc[e], 16);
      acc[e] *= 0.25f;
    }
    if (q == 0) {
      const int64_t po = (static_cast<int64_t>(no) * Hh + yh) * Wh + xh;
      if (h32) {
        float* o = h32 + po * ldh32 + c;
        *reinterpret_cast<float4*>(o) = make_float4(acc[0], acc[1], acc[2], acc[3]);
        *reinterpret_cast<float4*>(o + 4) = make_float4(acc[4], acc[5], acc[6], acc[7]);
      }
      if (h16) *reinterpret_cast<half8*>(h16 + po * ldh16 + c) = pack8(acc);
    }
  }
}

// Same operation for an exact feature pyramid (H2 = H1/2, H3 = H1/4, H4 = H1/8, likewise W): one thread = one
// half-resolution pixel (2 x 2 full-resolution pixels) x 8 channels.  With exact power-of-two ratios the four pixels
// of a block interpolate inside ONE 3 x 3 neighbourhood of p2 and ONE 2 x 2 neighbourhood of p3 and of p4, so a
// block needs 4 + 9 + 4 + 4 = 21 sixteen-byte loads instead of 4 x 13, the horizontal interpolation of a tap row is
// shared by the two pixel rows, and the 2 x 2 mean stays in registers (no shuffles).  Interpolation weights still
// come from lerp_coord() (border clamps included); only the tap indices use the closed form.
__device__ __forceinline__ void ld8f(const __half* __restrict__ p, float* f) {
  unpack8(*reinterpret_cast<const half8*>(p), f);
}

__global__ void __launch_bounds__(256)
head_fuse_pyramid_kernel(const __half* __restrict__ p1, const __half* __restrict__ p2, const __half* __restrict__ p3,
                         const __half* __restrict__ p4, int N, int H1, int W1, int C, int Tperm,
                         const float* __restrict__ shift, __half* __restrict__ c_full, float* __restrict__ h32,
                         int64_t ldh32, __half* __restrict__ h16, int64_t ldh16) {
  pdl_sync();
  const uint32_t chunks = C / 8, Hh = H1 / 2, Wh = W1 / 2;
  uint32_t idx = blockIdx.x * 256u + threadIdx.x;
  if (idx >= static_cast<uint32_t>(N) * Hh * Wh * chunks) return;
  const int c = static_cast<int>(idx % chunks) * 8; idx /= chunks;
  const int xh = static_cast<int>(idx % Wh); idx /= Wh;
  const int yh = static_cast<int>(idx % Hh), n = static_cast<int>(idx / Hh);
  const int no = Tperm > 1 ? (n % Tperm) * (N / Tperm) + n / Tperm : n;   // clip-major input -> frame-major output
  float acc[2][2][8];
  {
    const float4 s0 = *reinterpret_cast<const float4*>(shift + c), s1 = *reinterpret_cast<const float4*>(shift + c + 4);
    const float sh[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
    const __half* b1 = p1 + ((static_cast<int64_t>(n) * H1 + 2 * yh) * W1 + 2 * xh) * C + c;
#pragma unroll
    for (int dy = 0; dy < 2; ++dy)
#pragma unroll
      for (int dx = 0; dx < 2; ++dx) {
        ld8f(b1 + (static_cast<int64_t>(dy) * W1 + dx) * C, acc[dy][dx]);
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[dy][dx][e] += sh[e];
      }
  }
  {   // ---- p2 (x2): rows / columns {h-1, h, h+1} clamped; pixel 0 uses taps (0,1), pixel 1 uses taps (1,2)
    const int H2 = Hh, W2 = Wh;
    const Lerp ly0 = lerp_coord(2 * yh, 0.5f, H2), ly1 = lerp_coord(2 * yh + 1, 0.5f, H2);
    const Lerp lx0 = lerp_coord(2 * xh, 0.5f, W2), lx1 = lerp_coord(2 * xh + 1, 0.5f, W2);
    const int rows[3] = {max(yh - 1, 0), yh, min(yh + 1, H2 - 1)};
    const int cols[3] = {max(xh - 1, 0), xh, min(xh + 1, W2 - 1)};
    const __half* base = p2 + static_cast<int64_t>(n) * H2 * W2 * C + c;
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      float t[3][8], hx[2][8];
#pragma unroll
      for (int k = 0; k < 3; ++k) ld8f(base + (static_cast<int64_t>(rows[r]) * W2 + cols[k]) * C, t[k]);
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        hx[0][e] = lx0.w0 * t[0][e] + lx0.w1 * t[1][e];
        hx[1][e] = lx1.w0 * t[1][e] + lx1.w1 * t[2][e];
      }
#pragma unroll
      for (int dx = 0; dx < 2; ++dx)
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          if (r == 0) acc[0][dx][e] = fmaf(ly0.w0, hx[dx][e], acc[0][dx][e]);
          if (r == 1) { acc[0][dx][e] = fmaf(ly0.w1, hx[dx][e], acc[0][dx][e]); acc[1][dx][e] = fmaf(ly1.w0, hx[dx][e], acc[1][dx][e]); }
          if (r == 2) acc[1][dx][e] = fmaf(ly1.w1, hx[dx][e], acc[1][dx][e]);
        }
    }
  }
#pragma unroll
  for (int lvl = 0; lvl < 2; ++lvl) {   // ---- p3 (x4), p4 (x8): the four pixels share one 2 x 2 tap neighbourhood
    const int Hs = lvl == 0 ? H1 / 4 : H1 / 8, Ws = lvl == 0 ? W1 / 4 : W1 / 8;
    const float sc = lvl == 0 ? 0.25f : 0.125f;
    const __half* base = (lvl == 0 ? p3 : p4) + static_cast<int64_t>(n) * Hs * Ws * C + c;
    const Lerp ly0 = lerp_coord(2 * yh, sc, Hs), ly1 = lerp_coord(2 * yh + 1, sc, Hs);
    const Lerp lx0 = lerp_coord(2 * xh, sc, Ws), lx1 = lerp_coord(2 * xh + 1, sc, Ws);
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      float t[2][8], hx[2][8];
      const int row = r == 0 ? ly0.i0 : ly0.i1;
      ld8f(base + (static_cast<int64_t>(row) * Ws + lx0.i0) * C, t[0]);
      ld8f(base + (static_cast<int64_t>(row) * Ws + lx0.i1) * C, t[1]);
      const float wy0 = r == 0 ? ly0.w0 : ly0.w1, wy1 = r == 0 ? ly1.w0 : ly1.w1;
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        hx[0][e] = lx0.w0 * t[0][e] + lx0.w1 * t[1][e];
        hx[1][e] = lx1.w0 * t[0][e] + lx1.w1 * t[1][e];
      }
#pragma unroll
      for (int dx = 0; dx < 2; ++dx)
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          acc[0][dx][e] = fmaf(wy0, hx[dx][e], acc[0][dx][e]);
          acc[1][dx][e] = fmaf(wy1, hx[dx][e], acc[1][dx][e]);
        }
    }
  }
#pragma unroll
  for (int dy = 0; dy < 2; ++dy)
#pragma unroll
    for (int dx = 0; dx < 2; ++dx) {
#pragma unroll
      for (int e = 0; e < 8; ++e) acc[dy][dx][e] = fmaxf(acc[dy][dx][e], 0.f);
      if (c_full)
        *reinterpret_cast<half8*>(c_full + ((static_cast<int64_t>(no) * H1 + 2 * yh + dy) * W1 + 2 * xh + dx) * C + c) = pack8(acc[dy][dx]);
    }
  float m[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) m[e] = ((acc[0][0][e] + acc[0][1][e]) + (acc[1][0][e] + acc[1][1][e])) * 0.25f;
  const int64_t po = (static_cast<int64_t>(no) * Hh + yh) * Wh + xh;
  if (h32) {
    float* o = h32 + po * ldh32 + c;
    *reinterpret_cast<float4*>(o) = make_float4(m[0], m[1], m[2], m[3]);
    *reinterpret_cast<float4*>(o + 4) = make_float4(m[4], m[5], m[6], m[7]);
  }
  if (h16) *reinterpret_cast<half8*>(h16 + po * ldh16 + c) = pack8(m);
}

// ------------------------------------------------------------------------------------------------
// CFFA norm: LN over C = 256 of every frame; one warp per OUTPUT row, 8 channels per lane.
// Reference frames: one row of xn per token.  Target frames: one row per position of the cyclic-apron map
// xt_apron [n_t, Hp+6, Wp+6, C] that the CFM attention reads its 13 x 13 key halos from: position (Y, X) holds the
// zero-padded LN map (pad AFTER norm, cffm_transformer.py:716-724) at ((Y-3) mod Hp, (X-3) mod Wp), i.e. the
// torch.roll wrap-around of :389-400 is materialised once here.  Pad positions are written (as zeros) on every call,
// so the buffer carries no state from an earlier geometry.  The un-wrapped copy of a real token also goes to xn.
__global__ void __launch_bounds__(256)
cffa_norm_kernel(const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                 float eps, __half* __restrict__ xn, __half* __restrict__ xt_apron, int n_frames, int first_target, int H,
                 int W, int Hp, int Wp) {
  pdl_sync();
  constexpr int C = 256, RING = 3;
  const int64_t row = blockIdx.x * 8ll + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  const int64_t ref_rows = static_cast<int64_t>(first_target) * H * W;
  const int Ha = Hp + 2 * RING, Wa = Wp + 2 * RING;
  int64_t tok;                                                  // source token, or -1: zero row
  __half* dst_apron = nullptr;
  bool to_xn = true;
  if (row < ref_rows) {
    tok = row;
  } else {
    const int64_t pos = row - ref_rows;
    if (pos >= static_cast<int64_t>(n_frames - first_target) * Ha * Wa) return;
    const int X = static_cast<int>(pos % Wa), Y = static_cast<int>((pos / Wa) % Ha);
    const int bi = static_cast<int>(pos / (static_cast<int64_t>(Wa) * Ha));
    const int yy = (Y - RING + Hp) % Hp, xx = (X - RING + Wp) % Wp;
    dst_apron = xt_apron + pos * C + lane * 8;
    to_xn = (Y - RING == yy) && (X - RING == xx);
    tok = (yy < H && xx < W) ? (static_cast<int64_t>(first_target + bi) * H + yy) * W + xx : -1;
    if (tok < 0) {
      *reinterpret_cast<half8*>(dst_apron) = zero_half8();
      return;
    }
  }
  const float* px = x + tok * C + lane * 8;
  const float4 a = *reinterpret_cast<const float4*>(px), b = *reinterpret_cast<const float4*>(px + 4);
  float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
  float s = 0.f;
#pragma unroll
  for (int e = 0; e < 8; ++e) s += v[e];
  const float mean = warp_sum(s) * (1.f / C);
  float sq = 0.f;
#pragma unroll
  for (int e = 0; e < 8; ++e) { v[e] -= mean; sq += v[e] * v[e]; }
  const float rstd = rsqrtf(warp_sum(sq) * (1.f / C) + eps);
  const float4 g0 = *reinterpret_cast<const float4*>(gamma + lane * 8), g1 = *reinterpret_cast<const float4*>(gamma + lane * 8 + 4);
  const float4 b0 = *reinterpret_cast<const float4*>(beta + lane * 8), b1 = *reinterpret_cast<const float4*>(beta + lane * 8 + 4);
  const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
  const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
  for (int e = 0; e < 8; ++e) v[e] = v[e] * rstd * gg[e] + bb[e];
  const half8 o = pack8(v);
  if (to_xn) *reinterpret_cast<half8*>(xn + tok * C + lane * 8) = o;
  if (dst_apron) *reinterpret_cast<half8*>(dst_apron) = o;
}

// CFFA pooling: one warp per pooled token. Levels: 0 target 7x7 | 1 ref0 7x7 | 2 ref1 resize+3x3 | 3 ref2 resize+2x2
__global__ void __launch_bounds__(256)
cffa_pool_kernel(const __half* __restrict__ xn, int B, int T, int H, int W, int Hp, int Wp,
                 const float* __restrict__ pool_w, const float* __restrict__ pool_b, __half* __restrict__ pooled,
                 int only_level) {
  pdl_sync();
  constexpr int C = 256, WS = 7;
  const int nWh = Hp / WS, nWw = Wp / WS, nW = nWh * nWw;
  // only_level < 0: the levels of B clips from the frame-major stack xn [T,B,H,W,C] (P = 15 nW tokens per clip):
  //   -1 all four; -2 the three reference levels only (xn = reference frames [3,B,...]); -3 the target level only
  //   (xn = target frames [B,...], T = 1);  the rows of the other levels are left untouched
  // only_level = l: level l of B independent frames xn [B,H,W,C] (P = {1,1,4,9}[l] nW tokens per frame)
  const int P = only_level < 0 ? 15 * nW : (only_level < 2 ? nW : (only_level == 2 ? 4 * nW : 9 * nW));
  const int64_t gw = blockIdx.x * 8ll + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (gw >= static_cast<int64_t>(B) * P) return;
  const int b = static_cast<int>(gw / P);
  int pos = static_cast<int>(gw % P);
  int level, frame, wg, gwid, woff;
  if (only_level >= 0) level = only_level;
  else if (pos < nW) level = 0;
  else if (pos < 2 * nW) { level = 1; pos -= nW; }
  else if (pos < 6 * nW) { level = 2; pos -= 2 * nW; }
  else { level = 3; pos -= 6 * nW; }
  if ((only_level == -2 && level == 0) || (only_level == -3 && level != 0)) return;
  if (level == 0) { frame = T - 1; wg = 7; gwid = nWw; woff = 0; }
  else if (level == 1) { frame = 0; wg = 7; gwid = nWw; woff = 49; }
  else if (level == 2) { frame = 1; wg = 3; gwid = 2 * nWw; woff = 98; }
  else { frame = 2; wg = 2; gwid = 3 * nWw; woff = 107; }
  const int py = pos / gwid, px = pos % gwid;
  const __half* src = xn + (only_level < 0 ? static_cast<int64_t>(frame) * B + b : static_cast<int64_t>(b)) * H * W * C + lane * 8;
  float acc[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) acc[e] = 0.f;
  if (level < 2) {                                            // 7x7 fc-pool on the zero-padded map itself
#pragma unroll
    for (int u = 0; u < WS; ++u) {
      const int y = WS * py + u;
      half8 row[WS];
#pragma unroll
      for (int v = 0; v < WS; ++v) {                              // 7 independent loads in flight
        const int xx = WS * px + v;
        if (y < H && xx < W) row[v] = *reinterpret_cast<const half8*>(src + (static_cast<int64_t>(y) * W + xx) * C);
        else row[v] = zero_half8();
      }
#pragma unroll
      for (int v = 0; v < WS; ++v) {
        const float wt = pool_w[woff + u * WS + v];
        float t[8];
        unpack8(row[v], t);
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[e] = fmaf(wt, t[e], acc[e]);
      }
    }
  } else {                                                    // bilinear (Hp,Wp)->(Hpool,Wpool) then wg x wg fc-pool
    const int Hpool = nWh * 6, Wpool = nWw * 6;                     // l * floor(7/l) = 2*3 = 3*2 = 6 rows per window
    const float sy = static_cast<float>(Hp) / Hpool, sx = static_cast<float>(Wp) / Wpool;
    for (int u = 0; u < wg; ++u) {
      const Lerp ly = lerp_coord(wg * py + u, sy, Hp);
      for (int v = 0; v < wg; ++v) {
        const Lerp lx = lerp_coord(wg * px + v, sx, Wp);
        const float wt = pool_w[woff + u * wg + v];
        const int ys[2] = {ly.i0, ly.i1}, xs[2] = {lx.i0, lx.i1};
        const float wy[2] = {ly.w0, ly.w1}, wx[2] = {lx.w0, lx.w1};
#pragma unroll
        for (int a = 0; a < 2; ++a)
#pragma unroll
          for (int bq = 0; bq < 2; ++bq) {
            if (ys[a] < H && xs[bq] < W) {                    // pad rows/cols of the LN'ed map are zero
              float t[8];
              unpack8(*reinterpret_cast<const half8*>(src + (static_cast<int64_t>(ys[a]) * W + xs[bq]) * C), t);
              const float ww = wt * wy[a] * wx[bq];
#pragma unroll
              for (int e = 0; e < 8; ++e) acc[e] = fmaf(ww, t[e], acc[e]);
            }
          }
      }
    }
  }
  const float pb = pool_b[level];
#pragma unroll
  for (int e = 0; e < 8; ++e) acc[e] += pb;
  *reinterpret_cast<half8*>(pooled + gw * C + lane * 8) = pack8(acc);
}

// ------------------------------------------------------------------------------------------------
template <bool IN_F32>
__global__ void __launch_bounds__(256)
resize_nhwc_to_nchw_kernel(const void* __restrict__ in, int64_t ldc, float* __restrict__ out, int B, int h, int w,
                           int ncls, int Ho, int Wo) {
  pdl_sync();
  const int64_t total = static_cast<int64_t>(B) * Ho * Wo;
  const float sy = static_cast<float>(h) / Ho, sx = static_cast<float>(w) / Wo;
  for (int64_t i = blockIdx.x * 256ll + threadIdx.x; i < total; i += 256ll * gridDim.x) {
    const int X = static_cast<int>(i % Wo), Y = static_cast<int>((i / Wo) % Ho), b = static_cast<int>(i / (static_cast<int64_t>(Wo) * Ho));
    const Lerp ly = lerp_coord(Y, sy, h), lx = lerp_coord(X, sx, w);
    const int64_t r00 = ((static_cast<int64_t>(b) * h + ly.i0) * w + lx.i0) * ldc, r01 = ((static_cast<int64_t>(b) * h + ly.i0) * w + lx.i1) * ldc;
    const int64_t r10 = ((static_cast<int64_t>(b) * h + ly.i1) * w + lx.i0) * ldc, r11 = ((static_cast<int64_t>(b) * h + ly.i1) * w + lx.i1) * ldc;
    for (int c = 0; c < ncls; ++c) {
      float v00, v01, v10, v11;
      if (IN_F32) {
        const float* p = static_cast<const float*>(in);
        v00 = p[r00 + c]; v01 = p[r01 + c]; v10 = p[r10 + c]; v11 = p[r11 + c];
      } else {
        const __half* p = static_cast<const __half*>(in);
        v00 = __half2float(p[r00 + c]); v01 = __half2float(p[r01 + c]);
        v10 = __half2float(p[r10 + c]); v11 = __half2float(p[r11 + c]);
      }
      // same association as aten: w0y*(w0x*a + w1x*b) + w1y*(w0x*c + w1x*d)
      const float v = ly.w0 * (lx.w0 * v00 + lx.w1 * v01) + ly.w1 * (lx.w0 * v10 + lx.w1 * v11);
      out[((static_cast<int64_t>(b) * ncls + c) * Ho + Y) * Wo + X] = v;
    }
  }
}

__global__ void __launch_bounds__(256)
resize_argmax_kernel(const float* __restrict__ logits, int64_t* __restrict__ labels, int B, int ncls, int h, int w,
                     int Ho, int Wo) {
  pdl_sync();
  const int64_t total = static_cast<int64_t>(B) * Ho * Wo;
  const float sy = static_cast<float>(h) / Ho, sx = static_cast<float>(w) / Wo;
  for (int64_t i = blockIdx.x * 256ll + threadIdx.x; i < total; i += 256ll * gridDim.x) {
    const int X = static_cast<int>(i % Wo), Y = static_cast<int>((i / Wo) % Ho), b = static_cast<int>(i / (static_cast<int64_t>(Wo) * Ho));
    const Lerp ly = lerp_coord(Y, sy, h), lx = lerp_coord(X, sx, w);
    const float* base = logits + static_cast<int64_t>(b) * ncls * h * w;
    const int o00 = ly.i0 * w + lx.i0, o01 = ly.i0 * w + lx.i1, o10 = ly.i1 * w + lx.i0, o11 = ly.i1 * w + lx.i1;
    float best = -INFINITY;
    int arg = 0;
    for (int c = 0; c < ncls; ++c) {
      const float* p = base + static_cast<int64_t>(c) * h * w;
      const float v = ly.w0 * (lx.w0 * p[o00] + lx.w1 * p[o01]) + ly.w1 * (lx.w0 * p[o10] + lx.w1 * p[o11]);
      if (v > best) { best = v; arg = c; }
    }
    labels[i] = arg;
  }
}

// plain fp32 NCHW -> NCHW bilinear resize (whole_inference rescale step)
__global__ void __launch_bounds__(256)
resize_nchw_kernel(const float* __restrict__ in, float* __restrict__ out, int64_t planes, int h, int w, int Ho, int Wo) {
  pdl_sync();
  const int64_t total = planes * Ho * Wo;
  const float sy = static_cast<float>(h) / Ho, sx = static_cast<float>(w) / Wo;
  for (int64_t i = blockIdx.x * 256ll + threadIdx.x; i < total; i += 256ll * gridDim.x) {
    const int X = static_cast<int>(i % Wo), Y = static_cast<int>((i / Wo) % Ho);
    const int64_t pl = i / (static_cast<int64_t>(Wo) * Ho);
    const Lerp ly = lerp_coord(Y, sy, h), lx = lerp_coord(X, sx, w);
    const float* p = in + pl * h * w;
    out[i] = ly.w0 * (lx.w0 * p[ly.i0 * w + lx.i0] + lx.w1 * p[ly.i0 * w + lx.i1]) +
             ly.w1 * (lx.w0 * p[ly.i1 * w + lx.i0] + lx.w1 * p[ly.i1 * w + lx.i1]);
  }
}

// softmax over the channel dimension of fp32 NCHW; one thread per pixel (coalesced along W)
__global__ void __launch_bounds__(256)
softmax_nchw_kernel(const float* __restrict__ in, float* __restrict__ out, int B, int C, int64_t HW) {
  pdl_sync();
  const int64_t total = static_cast<int64_t>(B) * HW;
  for (int64_t i = blockIdx.x * 256ll + threadIdx.x; i < total; i += 256ll * gridDim.x) {
    const int64_t b = i / HW, px = i % HW;
    const float* p = in + b * C * HW + px;
    float* o = out + b * C * HW + px;
    float m = -INFINITY;
    for (int c = 0; c < C; ++c) m = fmaxf(m, p[c * HW]);
    float sum = 0.f;
    for (int c = 0; c < C; ++c) sum += expf(p[c * HW] - m);
    const float inv = 1.f / sum;
    for (int c = 0; c < C; ++c) o[c * HW] = expf(p[c * HW] - m) * inv;
  }
}

// Two chained bilinear resizes + argmax in one pass: NHWC fp32 class scores [B,h,w,ldc] -> (Hm,Wm) ->
// (Ho,Wo) -> labels.  cffm_head.py:149 followed by encoder_decoder.py:373-377,542,564 (softmax is
// monotone).  One CTA = a 16x16 tile of output pixels: the (<= UP_MT x UP_MT) intermediate-resolution
// values it needs are computed once per class into shared memory from the low-resolution map
// (channel-contiguous = coalesced reads), then every thread scans the classes of its own pixel.
constexpr int UP_TILE = 16, UP_MT = 8;
template <typename LT>                                          // label type: int64_t (what argmax returns) or uint8_t (ncls <= 256)
__global__ void __launch_bounds__(256)
upsample2_argmax_kernel(const float* __restrict__ lg, int64_t ldc, LT* __restrict__ labels, int h, int w,
                        int ncls, int Hm, int Wm, int Ho, int Wo) {
  pdl_sync();
  extern __shared__ __align__(16) uint8_t up_smem[];
  float* mid = reinterpret_cast<float*>(up_smem);              // [UP_MT*UP_MT][CP], CP odd -> conflict-free
  const int CP = ncls | 1;
  const int b = blockIdx.z, Y0 = blockIdx.y * UP_TILE, X0 = blockIdx.x * UP_TILE;
  const float sy1 = static_cast<float>(h) / Hm, sx1 = static_cast<float>(w) / Wm;
  const float sy2 = static_cast<float>(Hm) / Ho, sx2 = static_cast<float>(Wm) / Wo;
  const int my0 = lerp_coord(Y0, sy2, Hm).i0, mx0 = lerp_coord(X0, sx2, Wm).i0;
  const int my1 = lerp_coord(min(Y0 + UP_TILE, Ho) - 1, sy2, Hm).i1, mx1 = lerp_coord(min(X0 + UP_TILE, Wo) - 1, sx2, Wm).i1;
  const int nmy = my1 - my0 + 1, nmx = mx1 - mx0 + 1;          // host guarantees <= UP_MT
  const float* base = lg + static_cast<int64_t>(b) * h * w * ldc;
  for (int i = threadIdx.x; i < nmy * nmx * ncls; i += 256) {
    const int c = i % ncls, pos = i / ncls, mx = pos % nmx, my = pos / nmx;
    const Lerp ly = lerp_coord(my0 + my, sy1, h), lx = lerp_coord(mx0 + mx, sx1, w);
    const float v00 = base[(static_cast<int64_t>(ly.i0) * w + lx.i0) * ldc + c], v01 = base[(static_cast<int64_t>(ly.i0) * w + lx.i1) * ldc + c];
    const float v10 = base[(static_cast<int64_t>(ly.i1) * w + lx.i0) * ldc + c], v11 = base[(static_cast<int64_t>(ly.i1) * w + lx.i1) * ldc + c];
    mid[(my * UP_MT + mx) * CP + c] = ly.w0 * (lx.w0 * v00 + lx.w1 * v01) + ly.w1 * (lx.w0 * v10 + lx.w1 * v11);
  }
  __syncthreads();
  const int Y = Y0 + threadIdx.x / UP_TILE, X = X0 + threadIdx.x % UP_TILE;
  if (Y >= Ho || X >= Wo) return;
  const Lerp ly = lerp_coord(Y, sy2, Hm), lx = lerp_coord(X, sx2, Wm);
  const float* p00 = mid + ((ly.i0 - my0) * UP_MT + (lx.i0 - mx0)) * CP;
  const float* p01 = mid + ((ly.i0 - my0) * UP_MT + (lx.i1 - mx0)) * CP;
  const float* p10 = mid + ((ly.i1 - my0) * UP_MT + (lx.i0 - mx0)) * CP;
  const float* p11 = mid + ((ly.i1 - my0) * UP_MT + (lx.i1 - mx0)) * CP;
  float best = -INFINITY;
  int arg = 0;
  for (int c = 0; c < ncls; ++c) {
    const float v = ly.w0 * (lx.w0 * p00[c] + lx.w1 * p01[c]) + ly.w1 * (lx.w0 * p10[c] + lx.w1 * p11[c]);
    if (v > best) { best = v; arg = c; }
  }
  labels[(static_cast<int64_t>(b) * Ho + Y) * Wo + X] = static_cast<LT>(arg);
}

// Fast path of the same operation for the production geometry: exact x2 (h -> 2h) followed by exact x4 (2h -> 8h).
// Output pixels Y = 4 cy + 2 + j (j = 0..3) all interpolate between the intermediate rows cy and cy + 1 (clamped), so a
// "cell" (cy, cx) owns a 4 x 4 block of output pixels whose four corner values are read ONCE per class; the horizontal
// interpolation is shared by the 4 rows.  One CTA = 8 x 8 cells x 4 class groups (thread = cell x group, classes
// c = 4 i + g): ~6.5 instructions per (pixel, class) instead of ~13 + 4 shared-memory reads.  The 9 x 9 intermediate
// values of the tile are built once per CTA from the low-resolution scores (16-byte loads along the class dimension)
// into shared memory [pos][132] (row pitch = 4 mod 32 banks: the 8 cells x 4 groups of a warp hit 32 distinct banks).
// Same interpolation formula and operand order as the generic kernels; the 4 partial winners of a pixel are merged
// with "greater, or equal and lower class index" so the result is the first maximum, like argmax.
constexpr int UQ_CELLS = 8, UQ_MID = UQ_CELLS + 1, UQ_PITCH = 132;
template <typename LT>
__global__ void __launch_bounds__(256)
upsample2x4_argmax_kernel(const float* __restrict__ lg, int64_t ldc, LT* __restrict__ labels, int h, int w, int ncls) {
  pdl_sync();
  __shared__ __align__(16) float mid[UQ_MID * UQ_MID * UQ_PITCH];
  const int Hm = 2 * h, Wm = 2 * w, Ho = 8 * h, Wo = 8 * w;
  const int b = blockIdx.z, cy0 = blockIdx.y * UQ_CELLS - 1, cx0 = blockIdx.x * UQ_CELLS - 1;   // first cell of the tile
  const float* base = lg + static_cast<int64_t>(b) * h * w * ldc;
  // ---- intermediate values (rows cy0 .. cy0 + 8, clamped = replicated at the borders), 4 classes per thread
  const int nvec = (ncls + 3) / 4;                              // <= 32; the padded columns of a row are readable (ldc >= 4 nvec)
  for (int i = threadIdx.x; i < UQ_MID * UQ_MID * 32; i += 256) {
    const int c4 = i & 31, pos = i >> 5;
    if (c4 >= nvec) continue;
    const int my = min(max(cy0 + pos / UQ_MID, 0), Hm - 1), mx = min(max(cx0 + pos % UQ_MID, 0), Wm - 1);
    const Lerp ly = lerp_coord(my, 0.5f, h), lx = lerp_coord(mx, 0.5f, w);
    const float4 v00 = *reinterpret_cast<const float4*>(base + (static_cast<int64_t>(ly.i0) * w + lx.i0) * ldc + 4 * c4);
    const float4 v01 = *reinterpret_cast<const float4*>(base + (static_cast<int64_t>(ly.i0) * w + lx.i1) * ldc + 4 * c4);
    const float4 v10 = *reinterpret_cast<const float4*>(base + (static_cast<int64_t>(ly.i1) * w + lx.i0) * ldc + 4 * c4);
    const float4 v11 = *reinterpret_cast<const float4*>(base + (static_cast<int64_t>(ly.i1) * w + lx.i1) * ldc + 4 * c4);
    float4 m;
    m.x = ly.w0 * (lx.w0 * v00.x + lx.w1 * v01.x) + ly.w1 * (lx.w0 * v10.x + lx.w1 * v11.x);
    m.y = ly.w0 * (lx.w0 * v00.y + lx.w1 * v01.y) + ly.w1 * (lx.w0 * v10.y + lx.w1 * v11.y);
    m.z = ly.w0 * (lx.w0 * v00.z + lx.w1 * v01.z) + ly.w1 * (lx.w0 * v10.z + lx.w1 * v11.z);
    m.w = ly.w0 * (lx.w0 * v00.w + lx.w1 * v01.w) + ly.w1 * (lx.w0 * v10.w + lx.w1 * v11.w);
    *reinterpret_cast<float4*>(mid + pos * UQ_PITCH + 4 * c4) = m;
  }
  __syncthreads();
  // ---- thread = (cell, class group)
  const int g = threadIdx.x & 3, cell = threadIdx.x >> 2, cxl = cell & (UQ_CELLS - 1), cyl = cell / UQ_CELLS;
  const int Yb = 4 * (cy0 + cyl) + 2, Xb = 4 * (cx0 + cxl) + 2;  // first output pixel of the cell (may be -2)
  float wy1[4], wx1[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {                                 // out-of-range pixels are never stored; clamp for the weights
    wy1[j] = lerp_coord(min(max(Yb + j, 0), Ho - 1), 0.25f, Hm).w1;
    wx1[j] = lerp_coord(min(max(Xb + j, 0), Wo - 1), 0.25f, Wm).w1;
  }
  const float* p00 = mid + (cyl * UQ_MID + cxl) * UQ_PITCH + g;
  const float* p01 = p00 + UQ_PITCH;
  const float* p10 = p00 + UQ_MID * UQ_PITCH;
  const float* p11 = p10 + UQ_PITCH;
  float best[16];
  int arg[16];
#pragma unroll
  for (int k = 0; k < 16; ++k) { best[k] = -INFINITY; arg[k] = 0; }
  // lerp(a, b; w1) evaluated as a + w1 (b - a): one FMA per interpolated value (the generic kernels use
  // w0 a + w1 b; the two differ by an ulp, which can only move an exact near-tie of the arg max)
#pragma unroll 2
  for (int c = g; c < ncls; c += 4) {
    const float a = p00[c - g], bq = p01[c - g], cq = p10[c - g], d = p11[c - g];
    const float dab = bq - a, dcd = d - cq;
    float top[4], dtb[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      top[j] = fmaf(wx1[j], dab, a);
      dtb[j] = fmaf(wx1[j], dcd, cq) - top[j];
    }
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float v = fmaf(wy1[r], dtb[j], top[j]);
        if (v > best[4 * r + j]) { best[4 * r + j] = v; arg[4 * r + j] = c; }
      }
  }
  // ---- merge the 4 class groups (adjacent lanes); every lane ends with the winner
#pragma unroll
  for (int k = 0; k < 16; ++k) {
#pragma unroll
    for (int o = 1; o <= 2; o <<= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, best[k], o);
      const int oa = __shfl_xor_sync(0xffffffffu, arg[k], o);
      if (ov > best[k] || (ov == best[k] && oa < arg[k])) { best[k] = ov; arg[k] = oa; }
    }
  }
  // lane g stores row g of the 4 x 4 block
  int rowv[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) rowv[j] = g == 0 ? arg[j] : (g == 1 ? arg[4 + j] : (g == 2 ? arg[8 + j] : arg[12 + j]));
  const int Y = Yb + g;
  if (Y >= 0 && Y < Ho) {
    LT* orow = labels + (static_cast<int64_t>(b) * Ho + Y) * Wo;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int X = Xb + j;
      if (X >= 0 && X < Wo) orow[X] = static_cast<LT>(rowv[j]);
    }
  }
}

inline int grid_for(int64_t work_items, int per_block = 256) {
  int64_t g = (work_items + per_block - 1) / per_block;
  const int64_t cap = 148 * 16;                               // grid-stride: a few waves over 148 SMs
  return static_cast<int>(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace
}  // namespace cffm

using namespace cffm;

template <bool F32, int L, int NV>
static void launch_ln(const void* x, int64_t ldx, const float* gamma, const float* beta, float eps, __half* o16,
                      int64_t ldo16, float* o32, int64_t ldo32, int M, int C, cudaStream_t st, int nsum = 1,
                      int64_t sum_stride = 0, const float* sum_bias = nullptr, const float* gamma2 = nullptr,
                      const float* beta2 = nullptr, float eps2 = 0.f) {
  constexpr int RPW = 32 / L;
  const int64_t warps = (static_cast<int64_t>(M) + RPW - 1) / RPW;
  int64_t grid = (warps + 7) / 8;
  const int64_t cap = 148 * 8 * 4;                             // 8 resident CTAs per SM, a few rows per warp
  if (grid > cap) grid = cap;
  launch_k(layernorm_kernel<F32, L, NV>, static_cast<int>(grid), 256, 0, st, x, ldx, gamma, beta, eps, o16, ldo16, o32, ldo32, M, C,
           nsum, sum_stride, sum_bias, gamma2, beta2, eps2);
}

extern "C" int cffm_layernorm(const void* x, int x_is_f32, int64_t ldx, const float* gamma, const float* beta,
                              float eps, void* out_f16, int64_t ldo16, float* out_f32, int64_t ldo32, int M, int C,
                              void* stream) {
  CFFM_REQUIRE(x && gamma && beta && (out_f16 || out_f32), CFFM_E_BADARG, "layernorm: null pointer");
  CFFM_REQUIRE(M > 0 && C > 0, CFFM_E_BADARG, "layernorm: non-positive size");
  CFFM_REQUIRE(C <= 512 && C % 4 == 0, CFFM_E_UNSUPPORTED, "layernorm: need C %% 4 == 0 and C <= 512, got %d", C);
  CFFM_REQUIRE(ldx % 4 == 0 && (!out_f16 || ldo16 % 4 == 0) && (!out_f32 || ldo32 % 4 == 0) && aligned16(gamma) &&
                   aligned16(beta) && (reinterpret_cast<uintptr_t>(x) & (x_is_f32 ? 15 : 7)) == 0 &&
                   (reinterpret_cast<uintptr_t>(out_f16) & 7) == 0 && aligned16(out_f32),
               CFFM_E_BADARG, "layernorm: misaligned pointer or stride");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  __half* o16 = static_cast<__half*>(out_f16);
  const int nvec = C / 4;
#define CFFM_LN(L, NV)                                                                                       \
  do {                                                                                                       \
    if (x_is_f32) launch_ln<true, L, NV>(x, ldx, gamma, beta, eps, o16, ldo16, out_f32, ldo32, M, C, st);    \
    else launch_ln<false, L, NV>(x, ldx, gamma, beta, eps, o16, ldo16, out_f32, ldo32, M, C, st);            \
  } while (0)
  if (nvec <= 8) CFFM_LN(8, 1);
  else if (nvec <= 16) CFFM_LN(16, 1);
  else if (nvec <= 32) CFFM_LN(32, 1);
  else if (nvec <= 64) CFFM_LN(32, 2);
  else if (nvec <= 96) CFFM_LN(32, 3);
  else CFFM_LN(32, 4);
#undef CFFM_LN
  return launch_status("layernorm_kernel");
}

extern "C" int cffm_im2col(const void* x, int layout, int N, int H, int W, int C, int k, int stride, int pad, void* A,
                           int Kpad, void* stream) {
  CFFM_REQUIRE(x && A, CFFM_E_BADARG, "im2col: null pointer");
  CFFM_REQUIRE(N > 0 && H > 0 && W > 0 && C > 0 && k > 0 && stride > 0 && pad >= 0, CFFM_E_BADARG, "im2col: bad size");
  CFFM_REQUIRE(Kpad >= k * k * C && Kpad % 8 == 0, CFFM_E_BADARG, "im2col: Kpad=%d must be >= k*k*C and %% 8 == 0", Kpad);
  const int Ho = (H + 2 * pad - k) / stride + 1, Wo = (W + 2 * pad - k) / stride + 1;
  CFFM_REQUIRE(Ho > 0 && Wo > 0, CFFM_E_BADARG, "im2col: empty output");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (layout == 1) {
    CFFM_REQUIRE(C % 8 == 0 && aligned16(x) && aligned16(A), CFFM_E_UNSUPPORTED, "im2col: NHWC path needs C %% 8 == 0");
    launch_k(im2col_nhwc_kernel, grid_for(static_cast<int64_t>(N) * Ho * Wo * (Kpad / 8)), 256, 0, st, 
        static_cast<const __half*>(x), N, H, W, C, k, stride, pad, Ho, Wo, static_cast<__half*>(A), Kpad);
  } else {
    CFFM_REQUIRE(layout == 0, CFFM_E_BADARG, "im2col: bad layout %d", layout);
    const int padl = (pad + 3) & ~3, wp = (W + padl + pad + 3) & ~3;
    const int smem = ((C * k * wp * 2 + 15) & ~15) + Kpad * 4;
    CFFM_REQUIRE(smem <= 200 * 1024 && aligned16(A), CFFM_E_UNSUPPORTED,
                 "im2col: NCHW path stages C*k*(W+2*pad) halves (%d bytes) in shared memory", smem);
    CFFM_REQUIRE(aligned16(x), CFFM_E_BADARG, "im2col: misaligned input");
    if (const int rc = set_dyn_smem<im2col_nchw_f32_kernel>(200 * 1024, "im2col")) return rc;
    launch_k(im2col_nchw_f32_kernel, N * Ho, 256, smem, st, static_cast<const float*>(x), N, H, W, C, k, stride, pad, Ho, Wo,
                                                       static_cast<__half*>(A), Kpad);
  }
  return launch_status("im2col_kernel");
}

extern "C" int cffm_dwconv3x3_gelu(const void* x, const void* w, const float* bias, void* out, int N, int H, int W,
                                   int C, void* stream) {
  CFFM_REQUIRE(x && w && bias && out, CFFM_E_BADARG, "dwconv: null pointer");
  CFFM_REQUIRE(N > 0 && H > 0 && W > 0 && C > 0 && C % 8 == 0, CFFM_E_UNSUPPORTED, "dwconv: need C %% 8 == 0");
  CFFM_REQUIRE(aligned16(x) && aligned16(w) && aligned16(bias) && aligned16(out), CFFM_E_BADARG, "dwconv: misaligned");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const __half* xp = static_cast<const __half*>(x);
  const __half* wp = static_cast<const __half*>(w);
  __half* op = static_cast<__half*>(out);
#define CFFM_DW_CT(CPT, PXW, RS, CT)                                                                                \
  launch_k(dwconv3x3_gelu_rows_kernel<CPT, PXW, RS, CT>, static_cast<int>((threads + 255) / 256), 256, 0, st, xp, wp,   \
           bias, op, N, H, W, C, sx, sy)
#define CFFM_DW(CPT, PXW, RS)                                                                                       \
  do {                                                                                                              \
    const int sx = (W + PXW - 1) / PXW, sy = (H + RS - 1) / RS;                                                     \
    const int64_t threads = static_cast<int64_t>(N) * sy * sx * (C / CPT);                                          \
    CFFM_REQUIRE(threads < (1ll << 31), CFFM_E_UNSUPPORTED, "dwconv: tensor too large for 32-bit indexing");       \
    switch (C) {                           /* hidden widths of MiT-B0 .. B5; anything else: run-time channel count */ \
      case 128: CFFM_DW_CT(CPT, PXW, RS, 128); break;                                                               \
      case 256: CFFM_DW_CT(CPT, PXW, RS, 256); break;                                                               \
      case 512: CFFM_DW_CT(CPT, PXW, RS, 512); break;                                                               \
      case 640: CFFM_DW_CT(CPT, PXW, RS, 640); break;                                                               \
      case 1024: CFFM_DW_CT(CPT, PXW, RS, 1024); break;                                                             \
      case 1280: CFFM_DW_CT(CPT, PXW, RS, 1280); break;                                                             \
      case 2048: CFFM_DW_CT(CPT, PXW, RS, 2048); break;                                                             \
      default: CFFM_DW_CT(CPT, PXW, RS, 0); break;                                                                  \
    }                                                                                                               \
  } while (0)
  CFFM_DW(4, 4, 4);   // 4 channels x 4 x 4 pixels per thread: best of the measured (CPT, PXW, RS) on all four stage shapes
#undef CFFM_DW
#undef CFFM_DW_CT
  return launch_status("dwconv3x3_gelu_rows_kernel");
}

extern "C" int cffm_head_fuse(const void* p1, const void* p2, const void* p3, const void* p4, int N, int H1, int W1,
                              int H2, int W2, int H3, int W3, int H4, int W4, int C, int T_perm, const float* shift,
                              void* c_full, float* c_half_f32, int64_t ldh32, void* c_half_f16, int64_t ldh16,
                              void* stream) {
  CFFM_REQUIRE(p1 && p2 && p3 && p4 && shift, CFFM_E_BADARG, "head_fuse: null pointer");
  CFFM_REQUIRE(c_full || c_half_f32 || c_half_f16, CFFM_E_BADARG, "head_fuse: no output requested");
  CFFM_REQUIRE(N > 0 && H1 > 0 && W1 > 0 && H2 > 0 && W2 > 0 && H3 > 0 && W3 > 0 && H4 > 0 && W4 > 0, CFFM_E_BADARG,
               "head_fuse: non-positive size");
  CFFM_REQUIRE(C % 64 == 0 && H1 % 2 == 0 && W1 % 2 == 0, CFFM_E_UNSUPPORTED, "head_fuse: need C %% 64 == 0 and even H1, W1");
  CFFM_REQUIRE((!c_half_f32 || ldh32 % 4 == 0) && (!c_half_f16 || ldh16 % 8 == 0), CFFM_E_BADARG, "head_fuse: bad stride");
  CFFM_REQUIRE(T_perm >= 0 && (T_perm <= 1 || N % T_perm == 0), CFFM_E_BADARG, "head_fuse: N=%d not a multiple of T_perm=%d", N, T_perm);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int64_t warps = static_cast<int64_t>(N) * (H1 / 2) * (W1 / 2) * (C / 64);
  CFFM_REQUIRE(warps * 32 < (1ll << 31), CFFM_E_UNSUPPORTED, "head_fuse: tensor too large for 32-bit indexing");
  const bool pyramid = H1 % 8 == 0 && W1 % 8 == 0 && H2 * 2 == H1 && W2 * 2 == W1 && H3 * 4 == H1 && W3 * 4 == W1 &&
                       H4 * 8 == H1 && W4 * 8 == W1;
  if (pyramid) {
    const int64_t threads = static_cast<int64_t>(N) * (H1 / 2) * (W1 / 2) * (C / 8);
    launch_k(head_fuse_pyramid_kernel, static_cast<int>((threads + 255) / 256), 256, 0, st, static_cast<const __half*>(p1),
             static_cast<const __half*>(p2), static_cast<const __half*>(p3), static_cast<const __half*>(p4), N, H1, W1, C, T_perm,
             shift, static_cast<__half*>(c_full), c_half_f32, ldh32, static_cast<__half*>(c_half_f16), ldh16);
    return launch_status("head_fuse_pyramid_kernel");
  }
  launch_k(head_fuse_kernel, grid_for(warps * 32), 256, 0, st,
      static_cast<const __half*>(p1), static_cast<const __half*>(p2), static_cast<const __half*>(p3),
      static_cast<const __half*>(p4), N, H1, W1, H2, W2, H3, W3, H4, W4, C, T_perm, shift, static_cast<__half*>(c_full),
      c_half_f32, ldh32, static_cast<__half*>(c_half_f16), ldh16);
  return launch_status("head_fuse_kernel");
}

extern "C" int cffm_cffa_norm(const float* x, const float* gamma, const float* beta, float eps, void* xn, void* xt_pad,
                              int B, int T, int H, int W, int Hp, int Wp, int C, void* stream) {
  CFFM_REQUIRE(x && gamma && beta && xn && xt_pad, CFFM_E_BADARG, "cffa_norm: null pointer");
  CFFM_REQUIRE(B > 0 && T > 0 && H > 0 && W > 0 && Hp >= H && Wp >= W, CFFM_E_BADARG, "cffa_norm: bad size");
  CFFM_REQUIRE(C == 256, CFFM_E_UNSUPPORTED, "cffa_norm: built for C=256, got %d", C);
  CFFM_REQUIRE(Hp % 7 == 0 && Wp % 7 == 0, CFFM_E_BADARG, "cffa_norm: Hp, Wp must be multiples of the window size 7");
  const int64_t rows = static_cast<int64_t>(B) * (T - 1) * H * W + static_cast<int64_t>(B) * (Hp + 6) * (Wp + 6);
  launch_k(cffa_norm_kernel, static_cast<int>((rows + 7) / 8), 256, 0, static_cast<cudaStream_t>(stream), 
      x, gamma, beta, eps, static_cast<__half*>(xn), static_cast<__half*>(xt_pad), B * T, (T - 1) * B, H, W, Hp, Wp);
  return launch_status("cffa_norm_kernel");
}

extern "C" int cffm_cffa_norm_frames(const float* x, const float* gamma, const float* beta, float eps, void* xn,
                                     void* xt_pad, int n_frames, int first_target, int H, int W, int Hp, int Wp, int C,
                                     void* stream) {
  CFFM_REQUIRE(x && gamma && beta && xn, CFFM_E_BADARG, "cffa_norm_frames: null pointer");
  CFFM_REQUIRE(n_frames > 0 && first_target >= 0 && first_target <= n_frames && H > 0 && W > 0 && Hp >= H && Wp >= W,
               CFFM_E_BADARG, "cffa_norm_frames: bad size");
  CFFM_REQUIRE(first_target == n_frames || xt_pad, CFFM_E_BADARG, "cffa_norm_frames: xt_pad required when targets exist");
  CFFM_REQUIRE(C == 256, CFFM_E_UNSUPPORTED, "cffa_norm_frames: built for C=256, got %d", C);
  CFFM_REQUIRE(Hp % 7 == 0 && Wp % 7 == 0, CFFM_E_BADARG, "cffa_norm_frames: Hp, Wp must be multiples of the window size 7");
  const int64_t rows = static_cast<int64_t>(first_target) * H * W + static_cast<int64_t>(n_frames - first_target) * (Hp + 6) * (Wp + 6);
  launch_k(cffa_norm_kernel, static_cast<int>((rows + 7) / 8), 256, 0, static_cast<cudaStream_t>(stream), x, gamma,
           beta, eps, static_cast<__half*>(xn), static_cast<__half*>(xt_pad), n_frames, first_target, H, W, Hp, Wp);
  return launch_status("cffa_norm_kernel");
}

extern "C" int cffm_cffa_pool(const void* xn, int B, int T, int H, int W, int C, const float* pool_w,
                              const float* pool_b, void* pooled, void* stream) {
  CFFM_REQUIRE(xn && pool_w && pool_b && pooled, CFFM_E_BADARG, "cffa_pool: null pointer");
  CFFM_REQUIRE(B > 0 && H > 0 && W > 0, CFFM_E_BADARG, "cffa_pool: bad size");
  CFFM_REQUIRE(C == 256 && T == 4, CFFM_E_UNSUPPORTED,
               "cffa_pool: built for C=256 and T=4 (3 reference frames, focal_l_clips=[1,2,3]); got C=%d T=%d", C, T);
  const int Hp = (H + 6) / 7 * 7, Wp = (W + 6) / 7 * 7;
  const int64_t warps = static_cast<int64_t>(B) * 15 * (Hp / 7) * (Wp / 7);
  launch_k(cffa_pool_kernel, static_cast<int>((warps + 7) / 8), 256, 0, static_cast<cudaStream_t>(stream), 
      static_cast<const __half*>(xn), B, T, H, W, Hp, Wp, pool_w, pool_b, static_cast<__half*>(pooled), -1);
  return launch_status("cffa_pool_kernel");
}

extern "C" int cffm_cffa_pool_part(const void* xn, int B, int part, int H, int W, int C, const float* pool_w,
                                   const float* pool_b, void* pooled, void* stream) {
  CFFM_REQUIRE(xn && pool_w && pool_b && pooled, CFFM_E_BADARG, "cffa_pool_part: null pointer");
  CFFM_REQUIRE(B > 0 && H > 0 && W > 0 && (part == 0 || part == 1), CFFM_E_BADARG, "cffa_pool_part: bad size/part");
  CFFM_REQUIRE(C == 256, CFFM_E_UNSUPPORTED, "cffa_pool_part: built for C=256, got %d", C);
  const int Hp = (H + 6) / 7 * 7, Wp = (W + 6) / 7 * 7;
  const int64_t warps = static_cast<int64_t>(B) * 15 * (Hp / 7) * (Wp / 7);
  launch_k(cffa_pool_kernel, static_cast<int>((warps + 7) / 8), 256, 0, static_cast<cudaStream_t>(stream),
           static_cast<const __half*>(xn), B, part == 0 ? 1 : 4, H, W, Hp, Wp, pool_w, pool_b, static_cast<__half*>(pooled),
           part == 0 ? -3 : -2);
  return launch_status("cffa_pool_kernel");
}

extern "C" int cffm_cffa_pool_level(const void* xn, int n_frames, int level, int H, int W, int C, const float* pool_w,
                                    const float* pool_b, void* pooled, void* stream) {
  CFFM_REQUIRE(xn && pool_w && pool_b && pooled, CFFM_E_BADARG, "cffa_pool_level: null pointer");
  CFFM_REQUIRE(n_frames > 0 && H > 0 && W > 0 && level >= 0 && level <= 3, CFFM_E_BADARG, "cffa_pool_level: bad size/level");
  CFFM_REQUIRE(C == 256, CFFM_E_UNSUPPORTED, "cffa_pool_level: built for C=256, got %d", C);
  const int Hp = (H + 6) / 7 * 7, Wp = (W + 6) / 7 * 7;
  const int per = level < 2 ? 1 : (level == 2 ? 4 : 9);
  const int64_t warps = static_cast<int64_t>(n_frames) * per * (Hp / 7) * (Wp / 7);
  launch_k(cffa_pool_kernel, static_cast<int>((warps + 7) / 8), 256, 0, static_cast<cudaStream_t>(stream),
           static_cast<const __half*>(xn), n_frames, 4, H, W, Hp, Wp, pool_w, pool_b, static_cast<__half*>(pooled), level);
  return launch_status("cffa_pool_kernel");
}

extern "C" int cffm_resize_nhwc_to_nchw(const void* in, int in_is_f32, int64_t ldc, float* out, int B, int h, int w,
                                        int ncls, int Ho, int Wo, void* stream) {
  CFFM_REQUIRE(in && out, CFFM_E_BADARG, "resize: null pointer");
  CFFM_REQUIRE(B > 0 && h > 0 && w > 0 && ncls > 0 && Ho > 0 && Wo > 0 && ldc >= ncls, CFFM_E_BADARG, "resize: bad size");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int grid = grid_for(static_cast<int64_t>(B) * Ho * Wo);
  if (in_is_f32) launch_k(resize_nhwc_to_nchw_kernel<true>, grid, 256, 0, st, in, ldc, out, B, h, w, ncls, Ho, Wo);
  else launch_k(resize_nhwc_to_nchw_kernel<false>, grid, 256, 0, st, in, ldc, out, B, h, w, ncls, Ho, Wo);
  return launch_status("resize_nhwc_to_nchw_kernel");
}

extern "C" int cffm_resize_argmax(const float* logits, int64_t* labels, int B, int ncls, int h, int w, int Ho, int Wo,
                                  void* stream) {
  CFFM_REQUIRE(logits && labels, CFFM_E_BADARG, "resize_argmax: null pointer");
  CFFM_REQUIRE(B > 0 && ncls > 0 && h > 0 && w > 0 && Ho > 0 && Wo > 0, CFFM_E_BADARG, "resize_argmax: bad size");
  launch_k(resize_argmax_kernel, grid_for(static_cast<int64_t>(B) * Ho * Wo), 256, 0, static_cast<cudaStream_t>(stream), 
      logits, labels, B, ncls, h, w, Ho, Wo);
  return launch_status("resize_argmax_kernel");
}

extern "C" int cffm_resize_nchw(const float* in, float* out, int B, int C, int h, int w, int Ho, int Wo, void* stream) {
  CFFM_REQUIRE(in && out, CFFM_E_BADARG, "resize_nchw: null pointer");
  CFFM_REQUIRE(B > 0 && C > 0 && h > 0 && w > 0 && Ho > 0 && Wo > 0, CFFM_E_BADARG, "resize_nchw: bad size");
  const int64_t planes = static_cast<int64_t>(B) * C;
  launch_k(resize_nchw_kernel, grid_for(planes * Ho * Wo), 256, 0, static_cast<cudaStream_t>(stream), in, out, planes, h, w, Ho, Wo);
  return launch_status("resize_nchw_kernel");
}

extern "C" int cffm_softmax_nchw(const float* in, float* out, int B, int C, int64_t HW, void* stream) {
  CFFM_REQUIRE(in && out, CFFM_E_BADARG, "softmax_nchw: null pointer");
  CFFM_REQUIRE(B > 0 && C > 0 && HW > 0, CFFM_E_BADARG, "softmax_nchw: bad size");
  launch_k(softmax_nchw_kernel, grid_for(static_cast<int64_t>(B) * HW), 256, 0, static_cast<cudaStream_t>(stream), in, out, B, C, HW);
  return launch_status("softmax_nchw_kernel");
}

namespace cffm {
namespace {
template <typename LT>
int upsample2_argmax_impl(const float* scores, int64_t ldc, LT* labels, int B, int h, int w, int ncls, int Hm, int Wm, int Ho,
                          int Wo, void* stream) {
  CFFM_REQUIRE(scores && labels, CFFM_E_BADARG, "upsample2_argmax: null pointer");
  CFFM_REQUIRE(B > 0 && h > 0 && w > 0 && ncls > 0 && Hm > 0 && Wm > 0 && Ho > 0 && Wo > 0 && ldc >= ncls && B <= 65535,
               CFFM_E_BADARG, "upsample2_argmax: bad size");
  CFFM_REQUIRE(sizeof(LT) > 1 || ncls <= 256, CFFM_E_UNSUPPORTED, "upsample2_argmax: %d classes do not fit 8-bit labels", ncls);
  if (Hm == 2 * h && Wm == 2 * w && Ho == 4 * Hm && Wo == 4 * Wm && ncls <= 128 && ldc % 4 == 0 && ldc >= (ncls + 3) / 4 * 4 &&
      aligned16(scores)) {
    // cells cy = -1 .. Hm - 1 (the first / last ones are half outside the image)
    dim3 grid((Wm + 1 + UQ_CELLS - 1) / UQ_CELLS, (Hm + 1 + UQ_CELLS - 1) / UQ_CELLS, B);
    launch_k(upsample2x4_argmax_kernel<LT>, grid, 256, 0, static_cast<cudaStream_t>(stream), scores, ldc, labels, h, w, ncls);
    return launch_status("upsample2x4_argmax_kernel");
  }
  // intermediate rows/cols touched by a 16-pixel output tile: 16 * Hm/Ho + 3 (two taps + rounding)
  const int need_y = (UP_TILE * Hm + Ho - 1) / Ho + 3, need_x = (UP_TILE * Wm + Wo - 1) / Wo + 3;
  CFFM_REQUIRE(need_y <= UP_MT && need_x <= UP_MT, CFFM_E_UNSUPPORTED,
               "upsample2_argmax: second stage must upsample by >= ~3.2x (tile needs %dx%d intermediate values, max %d)",
               need_y, need_x, UP_MT);
  const int smem = UP_MT * UP_MT * (ncls | 1) * 4;
  CFFM_REQUIRE(smem <= 200 * 1024, CFFM_E_UNSUPPORTED, "upsample2_argmax: too many classes (%d)", ncls);
  if (const int rc = set_dyn_smem<upsample2_argmax_kernel<LT>>(200 * 1024, "upsample2_argmax")) return rc;
  dim3 grid((Wo + UP_TILE - 1) / UP_TILE, (Ho + UP_TILE - 1) / UP_TILE, B);
  launch_k(upsample2_argmax_kernel<LT>, grid, 256, smem, static_cast<cudaStream_t>(stream), scores, ldc, labels, h, w, ncls, Hm, Wm, Ho, Wo);
  return launch_status("upsample2_argmax_kernel");
}
}  // namespace
}  // namespace cffm

extern "C" int cffm_upsample2_argmax(const float* scores, int64_t ldc, int64_t* labels, int B, int h, int w, int ncls,
                                     int Hm, int Wm, int Ho, int Wo, void* stream) {
  return cffm::upsample2_argmax_impl<int64_t>(scores, ldc, labels, B, h, w, ncls, Hm, Wm, Ho, Wo, stream);
}

extern "C" int cffm_upsample2_argmax_u8(const float* scores, int64_t ldc, uint8_t* labels, int B, int h, int w, int ncls,
                                        int Hm, int Wm, int Ho, int Wo, void* stream) {
  return cffm::upsample2_argmax_impl<uint8_t>(scores, ldc, labels, B, h, w, ncls, Hm, Wm, Ho, Wo, stream);
}

extern "C" int cffm_layernorm_sum(const float* partials, int nsum, const float* bias, const float* gamma, const float* beta,
                                  float eps, void* out_f16, int64_t ldo16, float* out_f32, int64_t ldo32, int M, int C,
                                  void* stream) {
  CFFM_REQUIRE(partials && gamma && beta && (out_f16 || out_f32), CFFM_E_BADARG, "layernorm_sum: null pointer");
  CFFM_REQUIRE(M > 0 && C > 0 && nsum >= 1, CFFM_E_BADARG, "layernorm_sum: bad size");
  CFFM_REQUIRE(C <= 512 && C % 4 == 0, CFFM_E_UNSUPPORTED, "layernorm_sum: need C %% 4 == 0 and C <= 512, got %d", C);
  CFFM_REQUIRE((!out_f16 || ldo16 % 4 == 0) && (!out_f32 || ldo32 % 4 == 0) && aligned16(gamma) && aligned16(beta) &&
                   aligned16(partials) && aligned16(bias) && (reinterpret_cast<uintptr_t>(out_f16) & 7) == 0 && aligned16(out_f32),
               CFFM_E_BADARG, "layernorm_sum: misaligned pointer or stride");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  __half* o16 = static_cast<__half*>(out_f16);
  const int nvec = C / 4;
  const int64_t stride = static_cast<int64_t>(M) * C;
#define CFFM_LNS(L, NV) launch_ln<true, L, NV>(partials, C, gamma, beta, eps, o16, ldo16, out_f32, ldo32, M, C, st, nsum, stride, bias)
  if (nvec <= 8) CFFM_LNS(8, 1);
  else if (nvec <= 16) CFFM_LNS(16, 1);
  else if (nvec <= 32) CFFM_LNS(32, 1);
  else if (nvec <= 64) CFFM_LNS(32, 2);
  else if (nvec <= 96) CFFM_LNS(32, 3);
  else CFFM_LNS(32, 4);
#undef CFFM_LNS
  return launch_status("layernorm_kernel");
}

extern "C" int cffm_layernorm_chain(const float* partials, int nsum, const float* bias, const float* gamma, const float* beta,
                                    float eps, float* out_f32, int64_t ldo32, const float* gamma2, const float* beta2,
                                    float eps2, void* out_f16, int64_t ldo16, int M, int C, void* stream) {
  CFFM_REQUIRE(partials && gamma && beta && gamma2 && beta2 && out_f32 && out_f16, CFFM_E_BADARG, "layernorm_chain: null pointer");
  CFFM_REQUIRE(M > 0 && C > 0 && nsum >= 1, CFFM_E_BADARG, "layernorm_chain: bad size");
  CFFM_REQUIRE(C <= 512 && C % 4 == 0, CFFM_E_UNSUPPORTED, "layernorm_chain: need C %% 4 == 0 and C <= 512, got %d", C);
  CFFM_REQUIRE(ldo16 % 4 == 0 && ldo32 % 4 == 0 && aligned16(gamma) && aligned16(beta) && aligned16(gamma2) && aligned16(beta2) &&
                   aligned16(partials) && aligned16(bias) && (reinterpret_cast<uintptr_t>(out_f16) & 7) == 0 && aligned16(out_f32),
               CFFM_E_BADARG, "layernorm_chain: misaligned pointer or stride");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  __half* o16 = static_cast<__half*>(out_f16);
  const int nvec = C / 4;
  const int64_t stride = static_cast<int64_t>(M) * C;
#define CFFM_LNC(L, NV) \
  launch_ln<true, L, NV>(partials, C, gamma, beta, eps, o16, ldo16, out_f32, ldo32, M, C, st, nsum, stride, bias, gamma2, beta2, eps2)
  if (nvec <= 8) CFFM_LNC(8, 1);
  else if (nvec <= 16) CFFM_LNC(16, 1);
  else if (nvec <= 32) CFFM_LNC(32, 1);
  else if (nvec <= 64) CFFM_LNC(32, 2);
  else if (nvec <= 96) CFFM_LNC(32, 3);
  else CFFM_LNC(32, 4);
#undef CFFM_LNC
  return launch_status("layernorm_kernel");
}
