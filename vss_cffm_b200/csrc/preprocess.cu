// Test-time clip preprocessing on the GPU: uint8 BGR HWC frames ->
// AlignedResize_clips (mmcv.imrescale + _align = two cv2.resize INTER_LINEAR, transforms.py:382-421) ->
// Normalize_clips (mmcv.imnormalize, transforms.py:1277-1297) -> CHW fp32, bit-exact with OpenCV / mmcv:
//   * cv2::resize INTER_LINEAR on 8-bit data is integer arithmetic: 11-bit fixed-point coefficients
//     (cvRound(c * 2048)), horizontal pass S = s0*a0 + s1*a1 in int32, vertical pass
//     (((b0 * (S0 >> 4)) >> 16) + ((b1 * (S1 >> 4)) >> 16) + 2) >> 2;
//   * imnormalize is float32(double(float32(x) - mean32) * (1 / double(std32))) after the BGR -> RGB swap.
// Byte / integer work, HBM-bound: one thread per output pixel, the three channels of a pixel together.
#include "common.cuh"

namespace cffm {
namespace {

struct Axis {
  int i0, i1, c0, c1;
};

// x rule of cv2 (source index clamped with the weight reset to 0 at the borders)
__device__ __forceinline__ Axis axis_x(int d, double scale, int src) {
  float f = static_cast<float>((d + 0.5) * scale - 0.5);
  int s = static_cast<int>(floorf(f));
  f -= static_cast<float>(s);
  if (s < 0) { f = 0.f; s = 0; }
  if (s >= src - 1) { f = 0.f; s = src - 1; }
  Axis a;
  a.i0 = s;
  a.i1 = min(s + 1, src - 1);
  a.c0 = __float2int_rn((1.f - f) * 2048.f);
  a.c1 = __float2int_rn(f * 2048.f);
  return a;
}

// y rule of cv2 (row indices clamped, weights kept)
__device__ __forceinline__ Axis axis_y(int d, double scale, int src) {
  float f = static_cast<float>((d + 0.5) * scale - 0.5);
  const int s = static_cast<int>(floorf(f));
  f -= static_cast<float>(s);
  Axis a;
  a.i0 = min(max(s, 0), src - 1);
  a.i1 = min(max(s + 1, 0), src - 1);
  a.c0 = __float2int_rn((1.f - f) * 2048.f);
  a.c1 = __float2int_rn(f * 2048.f);
  return a;
}

__device__ __forceinline__ void resize_pixel(const uint8_t* __restrict__ img, int w, const Axis& ay, const Axis& ax, int* v) {
  const uint8_t* r0 = img + static_cast<int64_t>(ay.i0) * w * 3;
  const uint8_t* r1 = img + static_cast<int64_t>(ay.i1) * w * 3;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const int s0 = r0[ax.i0 * 3 + c] * ax.c0 + r0[ax.i1 * 3 + c] * ax.c1;
    const int s1 = r1[ax.i0 * 3 + c] * ax.c0 + r1[ax.i1 * 3 + c] * ax.c1;
    const int o = (((ay.c0 * (s0 >> 4)) >> 16) + ((ay.c1 * (s1 >> 4)) >> 16) + 2) >> 2;
    v[c] = min(max(o, 0), 255);
  }
}

__global__ void __launch_bounds__(256)
resize_u8_kernel(const uint8_t* __restrict__ src, int N, int h, int w, uint8_t* __restrict__ dst, int H, int W,
                 double sy, double sx) {
  pdl_sync();
  const uint32_t total = static_cast<uint32_t>(N) * H * W;
  for (uint32_t i = blockIdx.x * 256u + threadIdx.x; i < total; i += 256u * gridDim.x) {
    const uint32_t X = i % W, t = i / W, Y = t % H, n = t / H;
    int v[3];
    resize_pixel(src + static_cast<int64_t>(n) * h * w * 3, w, axis_y(Y, sy, h), axis_x(X, sx, w), v);
    uint8_t* o = dst + static_cast<int64_t>(i) * 3;
    o[0] = static_cast<uint8_t>(v[0]); o[1] = static_cast<uint8_t>(v[1]); o[2] = static_cast<uint8_t>(v[2]);
  }
}

// resize (identity when the sizes are equal) + BGR->RGB + normalise -> planar fp32; frame n goes to out + n * out_stride
__global__ void __launch_bounds__(256)
resize_normalize_u8_kernel(const uint8_t* __restrict__ src, int N, int h, int w, float* __restrict__ out, int64_t out_stride,
                           int H, int W, double sy, double sx, float m0, float m1, float m2, double i0, double i1,
                           double i2, int to_rgb) {
  pdl_sync();
  const uint32_t total = static_cast<uint32_t>(N) * H * W;
  const float mean[3] = {m0, m1, m2};
  const double inv[3] = {i0, i1, i2};
  for (uint32_t i = blockIdx.x * 256u + threadIdx.x; i < total; i += 256u * gridDim.x) {
    const uint32_t X = i % W, t = i / W, Y = t % H, n = t / H;
    int v[3];
    resize_pixel(src + static_cast<int64_t>(n) * h * w * 3, w, axis_y(Y, sy, h), axis_x(X, sx, w), v);
    float* o = out + n * out_stride + static_cast<int64_t>(Y) * W + X;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float d = static_cast<float>(v[to_rgb ? 2 - c : c]) - mean[c];
      o[static_cast<int64_t>(c) * H * W] = static_cast<float>(static_cast<double>(d) * inv[c]);
    }
  }
}

inline int grid_px(int64_t px) {
  int64_t g = (px + 255) / 256;
  return static_cast<int>(g < 1 ? 1 : (g > 148 * 16 ? 148 * 16 : g));
}

}  // namespace
}  // namespace cffm

using namespace cffm;

extern "C" int cffm_resize_u8(const void* src, int N, int h, int w, void* dst, int H, int W, void* stream) {
  CFFM_REQUIRE(src && dst, CFFM_E_BADARG, "resize_u8: null pointer");
  CFFM_REQUIRE(N > 0 && h > 0 && w > 0 && H > 0 && W > 0, CFFM_E_BADARG, "resize_u8: non-positive size");
  CFFM_REQUIRE(static_cast<int64_t>(N) * H * W < (1ll << 31) && static_cast<int64_t>(h) * w * 3 < (1ll << 31), CFFM_E_UNSUPPORTED,
               "resize_u8: tensor too large for 32-bit indexing");
  launch_k(resize_u8_kernel, grid_px(static_cast<int64_t>(N) * H * W), 256, 0, static_cast<cudaStream_t>(stream),
           static_cast<const uint8_t*>(src), N, h, w, static_cast<uint8_t*>(dst), H, W, static_cast<double>(h) / H,
           static_cast<double>(w) / W);
  return launch_status("resize_u8_kernel");
}

extern "C" int cffm_resize_normalize_u8(const void* src, int N, int h, int w, float* out, int64_t out_stride, int H, int W,
                                        const float* mean3, const float* std3, int to_rgb, void* stream) {
  CFFM_REQUIRE(src && out && mean3 && std3, CFFM_E_BADARG, "resize_normalize_u8: null pointer");
  CFFM_REQUIRE(N > 0 && h > 0 && w > 0 && H > 0 && W > 0 && out_stride >= static_cast<int64_t>(3) * H * W, CFFM_E_BADARG,
               "resize_normalize_u8: bad size");
  CFFM_REQUIRE(static_cast<int64_t>(N) * H * W < (1ll << 31) && static_cast<int64_t>(h) * w * 3 < (1ll << 31), CFFM_E_UNSUPPORTED,
               "resize_normalize_u8: tensor too large for 32-bit indexing");
  // mean3 / std3 are HOST pointers (three floats each, like the reference's img_norm_cfg)
  launch_k(resize_normalize_u8_kernel, grid_px(static_cast<int64_t>(N) * H * W), 256, 0, static_cast<cudaStream_t>(stream),
           static_cast<const uint8_t*>(src), N, h, w, out, out_stride, H, W, static_cast<double>(h) / H,
           static_cast<double>(w) / W, mean3[0], mean3[1], mean3[2], 1.0 / static_cast<double>(std3[0]),
           1.0 / static_cast<double>(std3[1]), 1.0 / static_cast<double>(std3[2]), to_rgb);
  return launch_status("resize_normalize_u8_kernel");
}
