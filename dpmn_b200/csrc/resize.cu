// SURVEY.md 8(f) rank 2, second half: the recogniser-input resizes the reference runs around the hot path.
//   crnn_input      TextBase.parse_crnn_data (interfaces/base.py:419-425): F.interpolate(x, (32, 100), mode='bicubic')
//                   (ATen upsample_bicubic2d, align_corners=False, A = -0.75, taps clamped to the image) fused with the
//                   luma 0.299 R + 0.587 G + 0.114 B -> (B, 1, 32, 100).  fp32; one thread per output pixel, 48 taps.
//   visionlan_input TextBase.parse_visionlan_data (interfaces/base.py:473-478), batched: ToPILImage (x * 255 truncated to
//                   uint8) -> cv2.resize(img, (256, 64)) (INTER_LINEAR on uint8 = OpenCV's 11-bit fixed-point bilinear:
//                   horizontal pass in int32 at scale 2^11, vertical pass ((b * (row >> 4)) >> 16, + 2) >> 2) -> ToTensor
//                   (/ 255).  Integer work, bit-exact.  The reference does this per image through PIL + OpenCV on the host
//                   (D2H + H2D per image, super_resolution.py:177-178).
// Both are HBM-trivial (<= 2.4 MB per batch of 48): one coalesced pass, grid sized to the output.
#include "common.cuh"
#include "kernels.h"

namespace dpmn {

namespace {

__device__ __forceinline__ void cubic_coeffs(float t, float (&w)[4]) {
  const float A = -0.75f;
  const float x0 = t + 1.0f, x3 = 2.0f - t, x2 = 1.0f - t;
  w[0] = ((A * x0 - 5.0f * A) * x0 + 8.0f * A) * x0 - 4.0f * A;
  w[1] = ((A + 2.0f) * t - (A + 3.0f)) * t * t + 1.0f;
  w[2] = ((A + 2.0f) * x2 - (A + 3.0f)) * x2 * x2 + 1.0f;
  w[3] = ((A * x3 - 5.0f * A) * x3 + 8.0f * A) * x3 - 4.0f * A;
}

__global__ void __launch_bounds__(256) crnn_input_kernel(const float* __restrict__ img, long long img_bs, float* __restrict__ out,
                                                         int B, int H, int W, int OH, int OW, float scale_y, float scale_x) {
  const long long total = (long long)B * OH * OW;
  const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
  if (i >= total) return;
  const int ox = (int)(i % OW), oy = (int)((i / OW) % OH), b = (int)(i / ((long long)OW * OH));
  const float sy = scale_y * ((float)oy + 0.5f) - 0.5f, sx = scale_x * ((float)ox + 0.5f) - 0.5f;
  const float fy = floorf(sy), fx = floorf(sx);
  float wy[4], wx[4];
  cubic_coeffs(sy - fy, wy);
  cubic_coeffs(sx - fx, wx);
  int iy[4], ix[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    iy[k] = min(max((int)fy - 1 + k, 0), H - 1);
    ix[k] = min(max((int)fx - 1 + k, 0), W - 1);
  }
  const float* src = img + (long long)b * img_bs;
  float ch[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float* pl = src + (long long)c * H * W;
    float acc = 0.f;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const float* row = pl + (long long)iy[r] * W;
      float h = 0.f;
#pragma unroll
      for (int k = 0; k < 4; ++k) h += row[ix[k]] * wx[k];
      acc += h * wy[r];
    }
    ch[c] = acc;
  }
  out[i] = 0.299f * ch[0] + 0.587f * ch[1] + 0.114f * ch[2];
}

// OpenCV resize (INTER_LINEAR, 8U) coefficient of one destination coordinate: source index and the two 11-bit weights
__device__ __forceinline__ void cv_linear_coef(int d, double scale, int sn, bool collapse, int& s, int& a0, int& a1) {
  float f = (float)(((double)d + 0.5) * scale - 0.5);
  s = (int)floorf(f);
  f -= (float)s;
  if (collapse) {
    if (s < 0) { f = 0.f; s = 0; }
    if (s >= sn - 1) { f = 0.f; s = sn - 1; }
  }
  a0 = __float2int_rn((1.0f - f) * 2048.0f);
  a1 = __float2int_rn(f * 2048.0f);
}

__device__ __forceinline__ int to_u8(float v) { return (int)(unsigned char)(long long)(v * 255.0f); }

__global__ void __launch_bounds__(256) visionlan_input_kernel(const float* __restrict__ img, long long img_bs,
                                                              float* __restrict__ out, int B, int H, int W, int OH, int OW,
                                                              double scale_y, double scale_x) {
  const long long total = (long long)B * OH * OW;
  const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
  if (i >= total) return;
  const int dx = (int)(i % OW), dy = (int)((i / OW) % OH), b = (int)(i / ((long long)OW * OH));
  int sx, a0, a1, sy, b0, b1;
  cv_linear_coef(dx, scale_x, W, true, sx, a0, a1);
  cv_linear_coef(dy, scale_y, H, false, sy, b0, b1);
  const int x1 = min(sx + 1, W - 1);
  const int y0 = min(max(sy, 0), H - 1), y1 = min(max(sy + 1, 0), H - 1);
  const float* src = img + (long long)b * img_bs;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float* pl = src + (long long)c * H * W;
    const int r0 = to_u8(pl[y0 * W + sx]) * a0 + to_u8(pl[y0 * W + x1]) * a1;
    const int r1 = to_u8(pl[y1 * W + sx]) * a0 + to_u8(pl[y1 * W + x1]) * a1;
    const int v = (((b0 * (r0 >> 4)) >> 16) + ((b1 * (r1 >> 4)) >> 16) + 2) >> 2;
    out[(((long long)b * 3 + c) * OH + dy) * OW + dx] = (float)(v & 255) / 255.0f;
  }
}

}  // namespace

int launch_crnn_input(const float* img, long long img_bs, float* out, int B, int H, int W, int OH, int OW, cudaStream_t st) {
  const long long total = (long long)B * OH * OW;
  crnn_input_kernel<<<(int)((total + 255) / 256), 256, 0, st>>>(img, img_bs, out, B, H, W, OH, OW, (float)H / (float)OH,
                                                                (float)W / (float)OW);
  DPMN_LAUNCH_CHECK();
  return 0;
}

int launch_visionlan_input(const float* img, long long img_bs, float* out, int B, int H, int W, int OH, int OW,
                           cudaStream_t st) {
  const long long total = (long long)B * OH * OW;
  visionlan_input_kernel<<<(int)((total + 255) / 256), 256, 0, st>>>(img, img_bs, out, B, H, W, OH, OW,
                                                                     1.0 / ((double)OH / (double)H),
                                                                     1.0 / ((double)OW / (double)W));
  DPMN_LAUNCH_CHECK();
  return 0;
}

}  // namespace dpmn
