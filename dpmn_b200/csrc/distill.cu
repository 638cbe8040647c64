// DistillModule (model/distill_module.py:4-31), forward and backward, sm_100a SIMT.
//
//   u = conv3x3(cat[x_deep, x_shallow]) (6 -> 3)      a = ReLU(BN_1(u))        (distill_module.py:19-22)
//   v = conv3x3(x_shallow)              (3 -> 3)      s = ReLU(BN_2(v))        (:24-26)
//   loss = mean |a - s|   (nn.L1Loss, :28);  returns (loss, a)                 (:31)
//
// Four instances run per training step (interfaces/super_resolution.py:245-263), each ~12 torch launches forward and
// ~25 backward in the reference.  Here: forward = conv + per-channel sums (1 launch), statistics (1 tiny launch),
// BatchNorm + ReLU + L1 (1 launch); backward = 4 launches.  All tensors are (B, 3, 32, 128) fp32: the work is
// HBM/launch bound (per image 6 x 4096 x 4 B in, 3 x 4096 x 4 B out), every pass is one coalesced sweep over pixels.
// BatchNorm statistics are reduced deterministically (per-CTA fp32 partials, final sum in fp64).
#include "common.cuh"
#include "kernels.h"

namespace dpmn {

namespace {

constexpr int DS_T = 256;

__device__ __forceinline__ float ld_pad(const float* __restrict__ pl, int y, int x, int H, int W) {
  return (y >= 0 && y < H && x >= 0 && x < W) ? pl[y * W + x] : 0.f;
}

// 12 per-thread values -> one row of `partials` per CTA (fixed order: deterministic)
__device__ __forceinline__ void block_partials12(float (&v)[12], float* __restrict__ partials) {
  __shared__ float red[DS_T / 32][12];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < 12; ++i) {
    const float s = warp_sum(v[i]);
    if (lane == 0) red[warp][i] = s;
  }
  __syncthreads();
  if (threadIdx.x < 12) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < DS_T / 32; ++w) s += red[w][threadIdx.x];
    partials[(long long)blockIdx.x * 12 + threadIdx.x] = s;
  }
}

// u (3 ch) and v (3 ch) of every pixel into uv (B, 6, H, W); per-CTA sum / sum of squares of the 6 channels.
__global__ void __launch_bounds__(DS_T) distill_conv_kernel(const float* __restrict__ xd, long long d_bs,
                                                            const float* __restrict__ xs, long long s_bs, DistillParams w,
                                                            int B, int H, int W, float* __restrict__ uv,
                                                            float* __restrict__ partials) {
  __shared__ float s_wc[162], s_wf[81], s_b[6];
  for (int i = threadIdx.x; i < 162; i += DS_T) s_wc[i] = w.conv_cat_w[i];
  for (int i = threadIdx.x; i < 81; i += DS_T) s_wf[i] = w.conv_w[i];
  if (threadIdx.x < 3) { s_b[threadIdx.x] = w.conv_cat_b[threadIdx.x]; s_b[3 + threadIdx.x] = w.conv_b[threadIdx.x]; }
  __syncthreads();
  const int plane = H * W;
  const long long total = (long long)B * plane;
  const long long i = (long long)blockIdx.x * DS_T + threadIdx.x;
  float st[12];
#pragma unroll
  for (int k = 0; k < 12; ++k) st[k] = 0.f;
  if (i < total) {
    const int b = (int)(i / plane), p = (int)(i - (long long)b * plane);
    const int y = p / W, x = p - y * W;
    float acc[6];
#pragma unroll
    for (int c = 0; c < 6; ++c) acc[c] = s_b[c];
#pragma unroll
    for (int ci = 0; ci < 6; ++ci) {
      const float* pl = ci < 3 ? xd + (long long)b * d_bs + (long long)ci * plane
                               : xs + (long long)b * s_bs + (long long)(ci - 3) * plane;
#pragma unroll
      for (int ky = 0; ky < 3; ++ky)
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
          const float xv = ld_pad(pl, y + ky - 1, x + kx - 1, H, W);
          const int k = ky * 3 + kx;
#pragma unroll
          for (int co = 0; co < 3; ++co) acc[co] = fmaf(xv, s_wc[(co * 6 + ci) * 9 + k], acc[co]);
          if (ci >= 3) {
#pragma unroll
            for (int co = 0; co < 3; ++co) acc[3 + co] = fmaf(xv, s_wf[(co * 3 + ci - 3) * 9 + k], acc[3 + co]);
          }
        }
    }
    float* dst = uv + (long long)b * 6 * plane + p;
#pragma unroll
    for (int c = 0; c < 6; ++c) {
      dst[(long long)c * plane] = acc[c];
      st[c] = acc[c];
      st[6 + c] = acc[c] * acc[c];
    }
  }
  block_partials12(st, partials);
}

// 384 threads: warp q reduces quantity q of the partials in fp64.  Forward (mode 0): stats = {mean[6], rstd[6]} from batch
// statistics (training) or the running ones (eval) + the momentum update.  Backward (mode 1): the 12 sums are
// {sum dy[6], sum dy * xhat[6]}: BatchNorm affine gradients and the two projection coefficients of its backward.
__global__ void __launch_bounds__(384) distill_stats_kernel(const float* __restrict__ partials, int n_blocks, double count,
                                                            int mode, int training, int update_running, float eps,
                                                            float momentum, DistillParams w, float* __restrict__ stats,
                                                            float* g_bn1_w, float* g_bn1_b, float* g_bn2_w, float* g_bn2_b,
                                                            float* g_bc, float* g_bf) {
  __shared__ double tot[12];
  const int q = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double s = 0.0;
  if (mode == 1 || training)
    for (int i = lane; i < n_blocks; i += 32) s += (double)partials[(long long)i * 12 + q];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) tot[q] = s;
  __syncthreads();
  if (threadIdx.x >= 6) return;
  const int c = threadIdx.x;
  const float* rm = c < 3 ? w.bn1_mean : w.bn2_mean;
  const float* rv = c < 3 ? w.bn1_var : w.bn2_var;
  const int cc = c < 3 ? c : c - 3;
  if (mode == 0) {
    float mean, rstd;
    if (training) {
      const double m = tot[c] / count;
      double var = tot[6 + c] / count - m * m;
      if (var < 0.0) var = 0.0;
      mean = (float)m;
      rstd = (float)(1.0 / sqrt(var + (double)eps));
      if (update_running) {
        float* rmw = c < 3 ? w.bn1_mean : w.bn2_mean;
        float* rvw = c < 3 ? w.bn1_var : w.bn2_var;
        const double unbiased = count > 1.0 ? var * count / (count - 1.0) : var;
        rmw[cc] = (1.0f - momentum) * rmw[cc] + momentum * mean;
        rvw[cc] = (1.0f - momentum) * rvw[cc] + momentum * (float)unbiased;
      }
    } else {
      mean = rm[cc];
      rstd = 1.0f / sqrtf(rv[cc] + eps);
    }
    stats[c] = mean;
    stats[6 + c] = rstd;
  } else {
    const float dbeta = (float)tot[c], dgamma = (float)tot[6 + c];
    float* gw = c < 3 ? g_bn1_w : g_bn2_w;
    float* gb = c < 3 ? g_bn1_b : g_bn2_b;
    if (gw) gw[cc] += dgamma;
    if (gb) gb[cc] += dbeta;
    stats[12 + c] = training ? (float)(tot[c] / count) : 0.f;        // k1: mean of dy
    stats[18 + c] = training ? (float)(tot[6 + c] / count) : 0.f;    // k2: mean of dy * xhat
    // conv bias gradient = sum over pixels of d(conv out) = gamma * rstd * (sum dy - M k1 - k2 sum xhat); with batch
    // statistics sum xhat == 0 and the bracket vanishes identically
    const float gamma = (c < 3 ? w.bn1_w : w.bn2_w)[cc];
    float* gcb = c < 3 ? g_bc : g_bf;
    if (gcb && !training) gcb[cc] += gamma * stats[6 + c] * dbeta;
  }
}

// a = ReLU(BN_1(u)) -> feature; s = ReLU(BN_2(v)); loss += scale * sum |a - s|
__global__ void __launch_bounds__(DS_T) distill_bn_l1_kernel(const float* __restrict__ uv, const float* __restrict__ stats,
                                                             DistillParams w, int B, int H, int W, float inv_n,
                                                             float* __restrict__ feature, float* __restrict__ loss) {
  __shared__ float s_sc[6], s_sh[6];
  __shared__ float red[DS_T / 32];
  if (threadIdx.x < 6) {
    const int c = threadIdx.x, cc = c < 3 ? c : c - 3;
    const float g = (c < 3 ? w.bn1_w : w.bn2_w)[cc], be = (c < 3 ? w.bn1_b : w.bn2_b)[cc];
    s_sc[c] = stats[6 + c] * g;
    s_sh[c] = be - stats[c] * stats[6 + c] * g;
  }
  __syncthreads();
  const int plane = H * W;
  const long long total = (long long)B * plane;
  const long long i = (long long)blockIdx.x * DS_T + threadIdx.x;
  float acc = 0.f;
  if (i < total) {
    const int b = (int)(i / plane), p = (int)(i - (long long)b * plane);
    const float* src = uv + (long long)b * 6 * plane + p;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float a = fmaxf(fmaf(src[(long long)c * plane], s_sc[c], s_sh[c]), 0.f);
      const float s = fmaxf(fmaf(src[(long long)(3 + c) * plane], s_sc[3 + c], s_sh[3 + c]), 0.f);
      if (feature) feature[((long long)b * 3 + c) * plane + p] = a;
      acc += fabsf(a - s);
    }
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0 && loss) {
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < DS_T / 32; ++k) s += red[k];
    atomicAdd(loss, s * inv_n);
  }
}

// dy = gradient w.r.t. the BatchNorm outputs (through ReLU and L1) into dy (B, 6, H, W) + per-CTA {sum dy, sum dy*xhat}
__global__ void __launch_bounds__(DS_T) distill_bwd_dy_kernel(const float* __restrict__ uv, const float* __restrict__ stats,
                                                              DistillParams w, int B, int H, int W, float inv_n,
                                                              const float* __restrict__ d_loss,
                                                              const float* __restrict__ d_feature, float* __restrict__ dy,
                                                              float* __restrict__ partials) {
  __shared__ float s_g[6], s_b[6], s_m[6], s_r[6];
  if (threadIdx.x < 6) {
    const int c = threadIdx.x, cc = c < 3 ? c : c - 3;
    s_g[c] = (c < 3 ? w.bn1_w : w.bn2_w)[cc];
    s_b[c] = (c < 3 ? w.bn1_b : w.bn2_b)[cc];
    s_m[c] = stats[c];
    s_r[c] = stats[6 + c];
  }
  __syncthreads();
  const float gl = d_loss ? d_loss[0] * inv_n : 0.f;
  const int plane = H * W;
  const long long total = (long long)B * plane;
  const long long i = (long long)blockIdx.x * DS_T + threadIdx.x;
  float st[12];
#pragma unroll
  for (int k = 0; k < 12; ++k) st[k] = 0.f;
  if (i < total) {
    const int b = (int)(i / plane), p = (int)(i - (long long)b * plane);
    const float* src = uv + (long long)b * 6 * plane + p;
    float* dst = dy + (long long)b * 6 * plane + p;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float xh_a = (src[(long long)c * plane] - s_m[c]) * s_r[c];
      const float xh_s = (src[(long long)(3 + c) * plane] - s_m[3 + c]) * s_r[3 + c];
      const float a_pre = fmaf(xh_a, s_g[c], s_b[c]), s_pre = fmaf(xh_s, s_g[3 + c], s_b[3 + c]);
      const float diff = fmaxf(a_pre, 0.f) - fmaxf(s_pre, 0.f);
      const float sg = diff > 0.f ? gl : (diff < 0.f ? -gl : 0.f);
      float da = sg + (d_feature ? d_feature[((long long)b * 3 + c) * plane + p] : 0.f);
      da = a_pre > 0.f ? da : 0.f;
      const float ds = s_pre > 0.f ? -sg : 0.f;
      dst[(long long)c * plane] = da;
      dst[(long long)(3 + c) * plane] = ds;
      st[c] = da; st[3 + c] = ds;
      st[6 + c] = da * xh_a; st[9 + c] = ds * xh_s;
    }
  }
  block_partials12(st, partials);
}

// BatchNorm backward in place: dy -> d(conv out) = gamma * rstd * (dy - k1 - xhat * k2)   (k1 = k2 = 0 in eval)
__global__ void __launch_bounds__(DS_T) distill_bwd_du_kernel(const float* __restrict__ uv, const float* __restrict__ stats,
                                                              DistillParams w, long long total, int plane,
                                                              float* __restrict__ dy) {
  const long long i = (long long)blockIdx.x * DS_T + threadIdx.x;
  if (i >= total) return;
  const int c = (int)((i / plane) % 6), cc = c < 3 ? c : c - 3;
  const float g = (c < 3 ? w.bn1_w : w.bn2_w)[cc];
  const float xh = (uv[i] - stats[c]) * stats[6 + c];
  dy[i] = g * stats[6 + c] * (dy[i] - stats[12 + c] - xh * stats[18 + c]);
}

// data gradients (transposed 3x3 convs gathered per pixel) and the 243 weight gradients (per-warp shuffle sums ->
// shared accumulators -> one global atomic per weight per CTA)
__global__ void __launch_bounds__(DS_T) distill_bwd_conv_kernel(const float* __restrict__ xd, long long d_bs,
                                                                const float* __restrict__ xs, long long s_bs,
                                                                const float* __restrict__ du, DistillParams w, int B, int H,
                                                                int W, float* __restrict__ g_xd, float* __restrict__ g_xs,
                                                                float* __restrict__ g_wc, float* __restrict__ g_wf) {
  __shared__ float s_wc[162], s_wf[81], s_acc[243];
  for (int i = threadIdx.x; i < 162; i += DS_T) s_wc[i] = w.conv_cat_w[i];
  for (int i = threadIdx.x; i < 81; i += DS_T) s_wf[i] = w.conv_w[i];
  for (int i = threadIdx.x; i < 243; i += DS_T) s_acc[i] = 0.f;
  __syncthreads();
  const int plane = H * W;
  const long long total = (long long)B * plane;
  const long long i = (long long)blockIdx.x * DS_T + threadIdx.x;
  const bool live = i < total;
  const int b = live ? (int)(i / plane) : 0, p = live ? (int)(i - (long long)b * plane) : 0;
  const int y = p / W, x = p - y * W;
  const float* dub = du + (long long)b * 6 * plane;
  if (live && (g_xd || g_xs)) {
    float gd[3] = {0.f, 0.f, 0.f}, gs[3] = {0.f, 0.f, 0.f};
#pragma unroll
    for (int co = 0; co < 3; ++co)
#pragma unroll
      for (int ky = 0; ky < 3; ++ky)
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
          // out(q) reads in(q + off(k)): in(p) feeds out(p - off(k))
          const int yy = y - (ky - 1), xx = x - (kx - 1);
          const float dvu = ld_pad(dub + (long long)co * plane, yy, xx, H, W);
          const float dvv = ld_pad(dub + (long long)(3 + co) * plane, yy, xx, H, W);
          const int k = ky * 3 + kx;
#pragma unroll
          for (int ci = 0; ci < 3; ++ci) {
            gd[ci] = fmaf(dvu, s_wc[(co * 6 + ci) * 9 + k], gd[ci]);
            gs[ci] = fmaf(dvu, s_wc[(co * 6 + 3 + ci) * 9 + k], gs[ci]);
            gs[ci] = fmaf(dvv, s_wf[(co * 3 + ci) * 9 + k], gs[ci]);
          }
        }
#pragma unroll
    for (int ci = 0; ci < 3; ++ci) {
      if (g_xd) g_xd[((long long)b * 3 + ci) * plane + p] = gd[ci];
      if (g_xs) g_xs[((long long)b * 3 + ci) * plane + p] = gs[ci];
    }
  }
  if (g_wc || g_wf) {
    float d6[6];
#pragma unroll
    for (int c = 0; c < 6; ++c) d6[c] = live ? dub[(long long)c * plane + p] : 0.f;
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int ci = 0; ci < 6; ++ci) {
      const float* pl = ci < 3 ? xd + (long long)b * d_bs + (long long)ci * plane
                               : xs + (long long)b * s_bs + (long long)(ci - 3) * plane;
#pragma unroll
      for (int k = 0; k < 9; ++k) {
        const float xv = live ? ld_pad(pl, y + k / 3 - 1, x + k % 3 - 1, H, W) : 0.f;
#pragma unroll
        for (int co = 0; co < 3; ++co) {
          const float s = warp_sum(d6[co] * xv);
          if (lane == 0) atomicAdd(&s_acc[(co * 6 + ci) * 9 + k], s);
        }
        if (ci >= 3) {
#pragma unroll
          for (int co = 0; co < 3; ++co) {
            const float s = warp_sum(d6[3 + co] * xv);
            if (lane == 0) atomicAdd(&s_acc[162 + (co * 3 + ci - 3) * 9 + k], s);
          }
        }
      }
    }
    __syncthreads();
    for (int j = threadIdx.x; j < 243; j += DS_T) {
      if (j < 162) { if (g_wc) atomicAdd(&g_wc[j], s_acc[j]); }
      else if (g_wf) atomicAdd(&g_wf[j - 162], s_acc[j]);
    }
  }
}

}  // namespace

size_t distill_workspace_floats(int B, int H, int W) {
  const long long px = (long long)B * H * W;
  const long long blocks = (px + DS_T - 1) / DS_T;
  return (size_t)(12 * px + 12 * blocks + 64);
}

static void carve(float* ws, int B, int H, int W, float*& uv, float*& dy, float*& partials, float*& stats) {
  const long long px = (long long)B * H * W;
  const long long blocks = (px + DS_T - 1) / DS_T;
  uv = ws; dy = uv + 6 * px; partials = dy + 6 * px; stats = partials + 12 * blocks;
}

int launch_distill_forward(const float* xd, long long d_bs, const float* xs, long long s_bs, const DistillParams& w, int B,
                           int H, int W, int training, int update_running, float eps, float momentum, float* feature,
                           float* loss, float* ws, cudaStream_t st) {
  float *uv, *dy, *partials, *stats;
  carve(ws, B, H, W, uv, dy, partials, stats);
  const long long px = (long long)B * H * W;
  const int blocks = (int)((px + DS_T - 1) / DS_T);
  distill_conv_kernel<<<blocks, DS_T, 0, st>>>(xd, d_bs, xs, s_bs, w, B, H, W, uv, partials);
  DPMN_LAUNCH_CHECK();
  distill_stats_kernel<<<1, 384, 0, st>>>(partials, blocks, (double)px, 0, training, update_running, eps, momentum, w, stats,
                                          nullptr, nullptr, nullptr, nullptr, nullptr, nullptr);
  DPMN_LAUNCH_CHECK();
  if (feature || loss) {
    distill_bn_l1_kernel<<<blocks, DS_T, 0, st>>>(uv, stats, w, B, H, W, 1.0f / (float)(3 * px), feature, loss);
    DPMN_LAUNCH_CHECK();
  }
  return 0;
}

int launch_distill_backward(const float* xd, long long d_bs, const float* xs, long long s_bs, const DistillParams& w, int B,
                            int H, int W, int training, const float* d_loss, const float* d_feature, const DistillGrads& g,
                            float* ws, cudaStream_t st) {
  float *uv, *dy, *partials, *stats;
  carve(ws, B, H, W, uv, dy, partials, stats);
  const long long px = (long long)B * H * W;
  const int blocks = (int)((px + DS_T - 1) / DS_T);
  distill_bwd_dy_kernel<<<blocks, DS_T, 0, st>>>(uv, stats, w, B, H, W, 1.0f / (float)(3 * px), d_loss, d_feature, dy, partials);
  DPMN_LAUNCH_CHECK();
  distill_stats_kernel<<<1, 384, 0, st>>>(partials, blocks, (double)px, 1, training, 0, 0.f, 0.f, w, stats, g.bn1_w, g.bn1_b,
                                          g.bn2_w, g.bn2_b, g.conv_cat_b, g.conv_b);
  DPMN_LAUNCH_CHECK();
  const long long total = 6 * px;
  distill_bwd_du_kernel<<<(int)((total + DS_T - 1) / DS_T), DS_T, 0, st>>>(uv, stats, w, total, H * W, dy);
  DPMN_LAUNCH_CHECK();
  distill_bwd_conv_kernel<<<blocks, DS_T, 0, st>>>(xd, d_bs, xs, s_bs, dy, w, B, H, W, g.x_deep, g.x_shallow, g.conv_cat_w,
                                                   g.conv_w);
  DPMN_LAUNCH_CHECK();
  return 0;
}

}  // namespace dpmn
