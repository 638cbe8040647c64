// SURVEY.md 8(f) "next" rows on either side of the hot path (sm_100a, HBM-bound elementwise work):
//   image_loss   ImageLoss(gradient=True) = w0 * MSE(out, target) + w1 * L1(gradient_map(out), gradient_map(target))
//                (loss/image_loss.py:15-43), value AND d loss / d out in one pass -- the reference evaluates it seven
//                times per training step (interfaces/super_resolution.py:212,239,267) as ~70 elementwise launches each
//   to_mask      toMask (utils/util.py:27-35): uint8 quantisation, ITU-R 601 luma in Pillow's fixed-point form, threshold
//                at the image's mean luma, inverted binary mask repeated on 3 channels -- the branch-2 prior of every
//                cascade step (super_resolution.py:220-226), a per-image PIL round trip through the host in the reference
#include "common.cuh"
#include "kernels.h"

namespace dpmn {

// gradient_map(x)[p] = sqrt(((x[right] - x[left]) / 2)^2 + ((x[top] - x[bottom]) / 2)^2 + 1e-6), zero padding
__device__ __forceinline__ void gmap_terms(const float* __restrict__ pl, int y, int x, int H, int W, float& a, float& b,
                                           float& g) {
  const float r = x + 1 < W ? pl[y * W + x + 1] : 0.f;
  const float l = x > 0 ? pl[y * W + x - 1] : 0.f;
  const float t = y > 0 ? pl[(y - 1) * W + x] : 0.f;
  const float bo = y + 1 < H ? pl[(y + 1) * W + x] : 0.f;
  a = (r - l) * 0.5f;
  b = (t - bo) * 0.5f;
  g = sqrtf(a * a + b * b + 1e-6f);
}

// s(p) * {a, b}(p) / g_out(p), s = sign(g_out - g_target): the factor every neighbour of p receives
__device__ __forceinline__ void gmap_back(const float* __restrict__ po, const float* __restrict__ pt, int y, int x, int H,
                                          int W, float& fa, float& fb) {
  if (y < 0 || y >= H || x < 0 || x >= W) { fa = 0.f; fb = 0.f; return; }
  float a, b, g, at, bt, gt;
  gmap_terms(po, y, x, H, W, a, b, g);
  gmap_terms(pt, y, x, H, W, at, bt, gt);
  const float d = g - gt;
  const float s = d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f);
  fa = s * a / g;
  fb = s * b / g;
}

__global__ void __launch_bounds__(256) image_loss_kernel(const float* __restrict__ out, long long out_bs,
                                                         const float* __restrict__ tgt, long long tgt_bs, int B, int C, int H,
                                                         int W, int gp_ch, float w_mse, float w_gp, float scale,
                                                         float* __restrict__ loss, float* __restrict__ d_out) {
  __shared__ float red[8];
  const long long plane = (long long)H * W;
  const long long total = (long long)B * C * plane;
  const float inv_n = 1.0f / (float)total;
  const float inv_ng = 1.0f / (float)((long long)B * gp_ch * plane);
  float acc = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(i % W);
    const int y = (int)((i / W) % H);
    const int c = (int)((i / plane) % C);
    const int b = (int)(i / (plane * C));
    const float* po = out + (long long)b * out_bs + c * plane;
    const float* pt = tgt + (long long)b * tgt_bs + c * plane;
    const float diff = po[y * W + x] - pt[y * W + x];
    float term = w_mse * diff * diff * inv_n;
    float grad = w_mse * 2.0f * diff * inv_n;
    if (c < gp_ch) {
      float a, bb, g, at, bt, gt;
      gmap_terms(po, y, x, H, W, a, bb, g);
      gmap_terms(pt, y, x, H, W, at, bt, gt);
      term += w_gp * fabsf(g - gt) * inv_ng;
      if (d_out != nullptr) {
        // x[q] enters a(q - ex) with +1/2, a(q + ex) with -1/2, b(q + ey) with +1/2, b(q - ey) with -1/2
        float fa, fb, gsum = 0.f;
        gmap_back(po, pt, y, x - 1, H, W, fa, fb); gsum += 0.5f * fa;
        gmap_back(po, pt, y, x + 1, H, W, fa, fb); gsum -= 0.5f * fa;
        gmap_back(po, pt, y + 1, x, H, W, fa, fb); gsum += 0.5f * fb;
        gmap_back(po, pt, y - 1, x, H, W, fa, fb); gsum -= 0.5f * fb;
        grad += w_gp * gsum * inv_ng;
      }
    }
    acc += term;
    if (d_out != nullptr) d_out[i] = grad * scale;
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += red[i];
    atomicAdd(loss, s * scale);
  }
}

int launch_image_loss(const float* out, long long out_bs, const float* tgt, long long tgt_bs, int B, int C, int H, int W,
                      float w_mse, float w_gp, float scale, float* loss, float* d_out, cudaStream_t st) {
  const long long total = (long long)B * C * H * W;
  int blocks = (int)((total + 255) / 256);
  if (blocks > 1184) blocks = 1184;
  image_loss_kernel<<<blocks, 256, 0, st>>>(out, out_bs, tgt, tgt_bs, B, C, H, W, C < 3 ? C : 3, w_mse, w_gp, scale, loss, d_out);
  DPMN_LAUNCH_CHECK();
  return 0;
}

// ---- toMask -------------------------------------------------------------------------------------------------------
// ToPILImage: byte = (uint8)(x * 255) (truncation; wraps modulo 256 outside [0, 1]); convert('L'): (R*19595 + G*38470 + B*7471 +
// 0x8000) >> 16; threshold = mean of L over the image; mask = L > thres ? 0 : 255; ToTensor: / 255 -> {0, 1}.
// One CTA per image: pass 1 luma + integer sum, pass 2 compare L * n > sum (exact integer form of L > mean).
__device__ __forceinline__ int luma_u8(const float* __restrict__ img, long long plane, int p) {
  // torch's float -> uint8 cast goes through int64 (c10 TypeCast): out-of-range values wrap modulo 256, as they do
  // in the reference when a cascade image leaves [0, 1]
  const int r = (int)(unsigned char)(long long)(img[p] * 255.0f);
  const int g = (int)(unsigned char)(long long)(img[plane + p] * 255.0f);
  const int b = (int)(unsigned char)(long long)(img[2 * plane + p] * 255.0f);
  return (r * 19595 + g * 38470 + b * 7471 + 0x8000) >> 16;
}

__global__ void __launch_bounds__(256) to_mask_kernel(const float* __restrict__ img, long long img_bs, float* __restrict__ mask,
                                                      int H, int W) {
  __shared__ unsigned long long red[8];
  __shared__ unsigned long long total;
  const int b = blockIdx.x;
  const long long plane = (long long)H * W;
  const float* src = img + (long long)b * img_bs;
  unsigned long long s = 0;
  for (int p = threadIdx.x; p < (int)plane; p += 256) s += (unsigned long long)luma_u8(src, plane, p);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned long long t = 0;
    for (int i = 0; i < 8; ++i) t += red[i];
    total = t;
  }
  __syncthreads();
  const unsigned long long sum = total;
  float* dst = mask + (long long)b * 3 * plane;
  for (int p = threadIdx.x; p < (int)plane; p += 256) {
    const unsigned long long l = (unsigned long long)luma_u8(src, plane, p);
    const float v = (l * (unsigned long long)plane > sum) ? 0.0f : 1.0f;      // L > mean ? 0 : 255, then / 255
    dst[p] = v; dst[plane + p] = v; dst[2 * plane + p] = v;
  }
}

int launch_to_mask(const float* img, long long img_bs, float* mask, int B, int H, int W, cudaStream_t st) {
  to_mask_kernel<<<B, 256, 0, st>>>(img, img_bs, mask, H, W);
  DPMN_LAUNCH_CHECK();
  return 0;
}

}  // namespace dpmn
