// SIMT helper kernels of the tensor-core CMM path (cmm.py:80-161): weight staging, BatchNorm folding,
// the 3-channel stem conv, the SE gate on NHWC data and the 3-channel tail of de_1.
#include "common.cuh"
#include "kernels.h"

namespace dpmn {

// ---- weights: fp32 torch layout -> 16-bit [tap][rows][Cin] --------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) prep_weights_kernel(PrepBatch pb) {
  const PrepSeg sg = pb.seg[blockIdx.y];
  const long long n = (long long)sg.kk * sg.rows * sg.Cin;
  T* dst = reinterpret_cast<T*>(sg.dst);
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const int ci = (int)(i % sg.Cin);
    const int co = (int)((i / sg.Cin) % sg.rows);
    const int tap = (int)(i / ((long long)sg.Cin * sg.rows));
    float v = 0.f;
    if (co < sg.Cout)
      v = sg.transposed ? sg.src[((long long)ci * sg.Cout + co) * sg.kk + tap]
                        : sg.src[((long long)co * sg.Cin + ci) * sg.kk + tap];
    dst[i] = from_f32<T>(v);
  }
}

int launch_prep_weights(const PrepBatch& pb, DType t, cudaStream_t st) {
  if (pb.count < 1 || pb.count > 48) return -1;
  dim3 grid(296, pb.count);
  if (t == DT_F16) prep_weights_kernel<__half><<<grid, 256, 0, st>>>(pb);
  else if (t == DT_BF16) prep_weights_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(pb);
  else return -1;
  DPMN_LAUNCH_CHECK();
  return 0;
}

// ---- (conv bias, eval BatchNorm) -> (scale, shift):  y = acc * scale + shift -------------------------------
__global__ void __launch_bounds__(256) fold_bn_kernel(FoldBatch fb, float eps) {
  const FoldSeg s = fb.seg[blockIdx.y];
  for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < s.C; c += gridDim.x * blockDim.x) {
    const float bias = s.bias ? s.bias[c] : 0.f;
    if (s.bn_w == nullptr) {
      s.scale[c] = 1.f;
      s.shift[c] = bias;
    } else {
      const float sc = s.bn_w[c] / sqrtf(s.bn_rv[c] + eps);
      s.scale[c] = sc;
      s.shift[c] = (bias - s.bn_rm[c]) * sc + s.bn_b[c];
    }
  }
}

int launch_fold_bn(const FoldBatch& fb, float eps, cudaStream_t st) {
  if (fb.count < 1 || fb.count > 48) return -1;
  dim3 grid(4, fb.count);
  fold_bn_kernel<<<grid, 256, 0, st>>>(fb, eps);
  DPMN_LAUNCH_CHECK();
  return 0;
}

// ---- en_1: conv3x3 c_img -> cnum, fp32 NCHW in, two 16-bit NHWC outs -----------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) cmm_en1_kernel(const float* __restrict__ x1, const float* __restrict__ x2,
                                                      const float* __restrict__ w1, const float* __restrict__ b1,
                                                      const float* __restrict__ w2, const float* __restrict__ b2,
                                                      T* __restrict__ e1, T* __restrict__ cat1, int B, int H, int W,
                                                      int c_img, int cnum) {
  extern __shared__ __align__(16) float sw[];   // [c_img*9][cnum] + bias [cnum]
  const int g = blockIdx.y;
  const float* x = g == 0 ? x1 : x2;
  const float* w = g == 0 ? w1 : w2;
  const float* bsrc = g == 0 ? b1 : b2;
  const int K = c_img * 9;
  for (int i = threadIdx.x; i < K * cnum; i += blockDim.x) {
    const int co = i % cnum, k = i / cnum;
    sw[i] = w[co * K + k];
  }
  float* sb = sw + K * cnum;
  for (int i = threadIdx.x; i < cnum; i += blockDim.x) sb[i] = bsrc[i];
  __syncthreads();
  // one thread = two horizontally adjacent pixels x 8 output channels: each 3x3xc_img weight vector (two float4 from
  // shared memory) feeds 16 FMAs, and the 3 x 4 input patch is loaded once for both pixels
  const int groups = cnum / 8;
  const int Wp = W / 2;
  const long long total = (long long)B * H * Wp * groups;
  const long long HW = (long long)H * W;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int cg = (int)(idx % groups);
    const long long pp = idx / groups;
    const int xp = (int)(pp % Wp);
    const int yy = (int)((pp / Wp) % H);
    const int b = (int)(pp / ((long long)Wp * H));
    const int x0 = 2 * xp;
    float acc[2][8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { acc[0][j] = sb[cg * 8 + j]; acc[1][j] = acc[0][j]; }
    if (c_img == 3) {
      // all 36 input values first (independent predicated loads in flight together), then the 432 FMAs
      float v[3][3][4];
#pragma unroll
      for (int ci = 0; ci < 3; ++ci) {
        const float* xc = x + ((long long)b * 3 + ci) * HW;
#pragma unroll
        for (int ky = 0; ky < 3; ++ky) {
          const int y2 = yy + ky - 1;
          const bool row_ok = y2 >= 0 && y2 < H;
          const float* xr = xc + (long long)(row_ok ? y2 : yy) * W;
          const float2 mid = *reinterpret_cast<const float2*>(xr + x0);
          v[ci][ky][0] = (row_ok && x0 > 0) ? xr[x0 - 1] : 0.f;
          v[ci][ky][1] = row_ok ? mid.x : 0.f;
          v[ci][ky][2] = row_ok ? mid.y : 0.f;
          v[ci][ky][3] = (row_ok && x0 + 2 < W) ? xr[x0 + 2] : 0.f;
        }
      }
#pragma unroll
      for (int ci = 0; ci < 3; ++ci)
#pragma unroll
        for (int ky = 0; ky < 3; ++ky)
#pragma unroll
          for (int kx = 0; kx < 3; ++kx) {
            const float4* wr = reinterpret_cast<const float4*>(sw + ((ci * 3 + ky) * 3 + kx) * cnum + cg * 8);
            const float4 wa = wr[0], wb = wr[1];
            const float wv[8] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w};
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              acc[0][j] = fmaf(v[ci][ky][kx], wv[j], acc[0][j]);
              acc[1][j] = fmaf(v[ci][ky][kx + 1], wv[j], acc[1][j]);
            }
          }
    } else
    for (int ci = 0; ci < c_img; ++ci) {
      const float* xc = x + ((long long)b * c_img + ci) * HW;
#pragma unroll
      for (int ky = 0; ky < 3; ++ky) {
        const int y2 = yy + ky - 1;
        if (y2 < 0 || y2 >= H) continue;
        const float* xr = xc + (long long)y2 * W;
        float v[4];
        v[0] = x0 > 0 ? xr[x0 - 1] : 0.f;
        v[1] = xr[x0];
        v[2] = xr[x0 + 1];
        v[3] = x0 + 2 < W ? xr[x0 + 2] : 0.f;
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
          const float4* wr = reinterpret_cast<const float4*>(sw + ((ci * 3 + ky) * 3 + kx) * cnum + cg * 8);
          const float4 wa = wr[0], wb = wr[1];
          const float wv[8] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w};
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            acc[0][j] = fmaf(v[kx], wv[j], acc[0][j]);
            acc[1][j] = fmaf(v[kx + 1], wv[j], acc[1][j]);
          }
        }
      }
    }
#pragma unroll
    for (int px = 0; px < 2; ++px) {
      const long long pix = ((long long)b * H + yy) * W + x0 + px;
      union { uint4 u; T h[8]; } a, r;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        a.h[j] = from_f32<T>(acc[px][j] >= 0.f ? acc[px][j] : 0.2f * acc[px][j]);
        r.h[j] = from_f32<T>(fmaxf(acc[px][j], 0.f));
      }
      *reinterpret_cast<uint4*>(e1 + ((long long)g * B * HW + pix) * cnum + cg * 8) = a.u;
      *reinterpret_cast<uint4*>(cat1 + pix * (3 * cnum) + cnum * (1 + g) + cg * 8) = r.u;
    }
  }
}

// c_img = 3, W % 4 == 0: four horizontally adjacent pixels x 8 output channels per thread.  The two-pixel kernel above is
// bound by shared-memory bandwidth (ncu: MIO busy 86 %, one 2 x LDS.128 weight fetch per 16 FMAs); here the same fetch feeds
// 32 FMAs and the 3 x 3 x 6 input patch is loaded once for all four pixels.
template <typename T>
__global__ void __launch_bounds__(128, 4) cmm_en1_p4_kernel(const float* __restrict__ x1, const float* __restrict__ x2,
                                                         const float* __restrict__ w1, const float* __restrict__ b1,
                                                         const float* __restrict__ w2, const float* __restrict__ b2,
                                                         T* __restrict__ e1, T* __restrict__ cat1, int B, int H, int W,
                                                         int cnum) {
  extern __shared__ __align__(16) float sw[];   // [27][cnum] + bias [cnum]
  const int g = blockIdx.y;
  const float* x = g == 0 ? x1 : x2;
  const float* w = g == 0 ? w1 : w2;
  const float* bsrc = g == 0 ? b1 : b2;
  for (int i = threadIdx.x; i < 27 * cnum; i += blockDim.x) {
    const int co = i % cnum, k = i / cnum;
    sw[i] = w[co * 27 + k];
  }
  float* sb = sw + 27 * cnum;
  for (int i = threadIdx.x; i < cnum; i += blockDim.x) sb[i] = bsrc[i];
  __syncthreads();
  const int groups = cnum / 8;
  const int Wq = W / 4;
  const long long total = (long long)B * H * Wq * groups;
  const long long HW = (long long)H * W;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int cg = (int)(idx % groups);
    const long long pp = idx / groups;
    const int xq = (int)(pp % Wq);
    const int yy = (int)((pp / Wq) % H);
    const int b = (int)(pp / ((long long)Wq * H));
    const int x0 = 4 * xq;
    float v[3][3][6];
#pragma unroll
    for (int ci = 0; ci < 3; ++ci) {
      const float* xc = x + ((long long)b * 3 + ci) * HW;
#pragma unroll
      for (int ky = 0; ky < 3; ++ky) {
        const int y2 = yy + ky - 1;
        const bool row_ok = y2 >= 0 && y2 < H;
        const float* xr = xc + (long long)(row_ok ? y2 : yy) * W;
        const float4 mid = *reinterpret_cast<const float4*>(xr + x0);
        v[ci][ky][0] = (row_ok && x0 > 0) ? xr[x0 - 1] : 0.f;
        v[ci][ky][1] = row_ok ? mid.x : 0.f;
        v[ci][ky][2] = row_ok ? mid.y : 0.f;
        v[ci][ky][3] = row_ok ? mid.z : 0.f;
        v[ci][ky][4] = row_ok ? mid.w : 0.f;
        v[ci][ky][5] = (row_ok && x0 + 4 < W) ? xr[x0 + 4] : 0.f;
      }
    }
    float acc[4][8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float bj = sb[cg * 8 + j];
      acc[0][j] = bj; acc[1][j] = bj; acc[2][j] = bj; acc[3][j] = bj;
    }
#pragma unroll
    for (int ci = 0; ci < 3; ++ci)
#pragma unroll
      for (int ky = 0; ky < 3; ++ky)
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
          const float4* wr = reinterpret_cast<const float4*>(sw + ((ci * 3 + ky) * 3 + kx) * cnum + cg * 8);
          const float4 wa = wr[0], wb = wr[1];
          const float wv[8] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w};
#pragma unroll
          for (int px = 0; px < 4; ++px)
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[px][j] = fmaf(v[ci][ky][kx + px], wv[j], acc[px][j]);
        }
#pragma unroll
    for (int px = 0; px < 4; ++px) {
      const long long pix = ((long long)b * H + yy) * W + x0 + px;
      union { uint4 u; T h[8]; } a, r;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        a.h[j] = from_f32<T>(acc[px][j] >= 0.f ? acc[px][j] : 0.2f * acc[px][j]);
        r.h[j] = from_f32<T>(fmaxf(acc[px][j], 0.f));
      }
      *reinterpret_cast<uint4*>(e1 + ((long long)g * B * HW + pix) * cnum + cg * 8) = a.u;
      *reinterpret_cast<uint4*>(cat1 + pix * (3 * cnum) + cnum * (1 + g) + cg * 8) = r.u;
    }
  }
}

int launch_cmm_en1(const float* x1, const float* x2, const float* w1, const float* b1, const float* w2, const float* b2,
                   void* e1, void* cat1, DType t, int B, int H, int W, int c_img, int cnum, cudaStream_t st) {
  if (cnum % 8 || W % 2) return -2;
  const size_t smem = (size_t)(c_img * 9 * cnum + cnum) * sizeof(float);
  if (smem > 48 * 1024) return -2;
  dim3 grid(148 * 4, 2);
  if (c_img == 3 && W % 4 == 0 && ((reinterpret_cast<uintptr_t>(x1) | reinterpret_cast<uintptr_t>(x2)) & 15) == 0 &&
      (t == DT_F16 || t == DT_BF16)) {
    if (t == DT_F16)
      cmm_en1_p4_kernel<__half><<<dim3(148 * 8, 2), 128, smem, st>>>(x1, x2, w1, b1, w2, b2, (__half*)e1, (__half*)cat1, B, H, W, cnum);
    else
      cmm_en1_p4_kernel<__nv_bfloat16><<<dim3(148 * 8, 2), 128, smem, st>>>(x1, x2, w1, b1, w2, b2, (__nv_bfloat16*)e1,
                                                               (__nv_bfloat16*)cat1, B, H, W, cnum);
    DPMN_LAUNCH_CHECK();
    return 0;
  }
  if (t == DT_F16)
    cmm_en1_kernel<__half><<<grid, 256, smem, st>>>(x1, x2, w1, b1, w2, b2, (__half*)e1, (__half*)cat1, B, H, W, c_img, cnum);
  else if (t == DT_BF16)
    cmm_en1_kernel<__nv_bfloat16><<<grid, 256, smem, st>>>(x1, x2, w1, b1, w2, b2, (__nv_bfloat16*)e1,
                                                          (__nv_bfloat16*)cat1, B, H, W, c_img, cnum);
  else return -1;
  DPMN_LAUNCH_CHECK();
  return 0;
}

// ---- SE gate on NHWC fp32 halves -----------------------------------------------------------------------------
// Three small launches with one WARP per output so that the 2 x 1 MB of fc weights are read by thousands of
// warps in parallel instead of 48 latency-bound CTAs.
__global__ void __launch_bounds__(256) se_pool_kernel(const float* __restrict__ z6, float* __restrict__ pooled, int B,
                                                      int Cb, int hw) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;   // (b, c) with c in [0, 2*Cb)
  if (idx >= B * 2 * Cb) return;
  const int c = idx % (2 * Cb), b = idx / (2 * Cb);
  const int g = c / Cb, cl = c - g * Cb;
  const float* src = z6 + (((long long)g * B + b) * hw) * Cb + cl;
  float s = 0.f;
  for (int i = 0; i < hw; ++i) s += src[(long long)i * Cb];
  pooled[idx] = s / (float)hw;
}

__global__ void __launch_bounds__(256) se_fc1_kernel(const float* __restrict__ pooled, const float* __restrict__ w,
                                                     const float* __restrict__ bias, float* __restrict__ hid, int B,
                                                     int C2, int hidden) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= B * hidden) return;
  const int j = warp % hidden, b = warp / hidden;
  float s = 0.f;
  for (int c = lane; c < C2; c += 32) s = fmaf(w[(long long)j * C2 + c], pooled[(long long)b * C2 + c], s);
  s = warp_sum(s);
  if (lane == 0) hid[warp] = fmaxf(s + bias[j], 0.f);
}

template <typename T>
__global__ void __launch_bounds__(256) se_fc2_gate_kernel(const float* __restrict__ z6, const float* __restrict__ hid,
                                                          const float* __restrict__ w, const float* __restrict__ bias,
                                                          T* __restrict__ zg, int B, int Cb, int hw, int hidden) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const int C2 = 2 * Cb;
  if (warp >= B * C2) return;
  const int c = warp % C2, b = warp / C2;
  float s = 0.f;
  for (int j = lane; j < hidden; j += 32) s = fmaf(w[(long long)c * hidden + j], hid[(long long)b * hidden + j], s);
  s = warp_sum(s);
  const float gate = 1.0f / (1.0f + expf(-(s + bias[c])));
  const int g = c / Cb, cl = c - g * Cb;
  for (int i = lane; i < hw; i += 32) {
    const float v = z6[(((long long)g * B + b) * hw + i) * Cb + cl];
    zg[((long long)b * hw + i) * C2 + c] = from_f32<T>(fmaxf(fmaf(v, gate, v), 0.f));   // ReLU of de_6 folded in
  }
}

int launch_se_gate_nhwc(const float* z6, void* zg, DType t, const float* fc1_w, const float* fc1_b, const float* fc2_w,
                        const float* fc2_b, float* pooled, float* hid_buf, int B, int Cb, int hw, int hidden,
                        cudaStream_t st) {
  const int C2 = 2 * Cb;
  se_pool_kernel<<<(B * C2 + 255) / 256, 256, 0, st>>>(z6, pooled, B, Cb, hw);
  se_fc1_kernel<<<(B * hidden * 32 + 255) / 256, 256, 0, st>>>(pooled, fc1_w, fc1_b, hid_buf, B, C2, hidden);
  const int blocks = (B * C2 * 32 + 255) / 256;
  if (t == DT_F16)
    se_fc2_gate_kernel<__half><<<blocks, 256, 0, st>>>(z6, hid_buf, fc2_w, fc2_b, (__half*)zg, B, Cb, hw, hidden);
  else if (t == DT_BF16)
    se_fc2_gate_kernel<__nv_bfloat16><<<blocks, 256, 0, st>>>(z6, hid_buf, fc2_w, fc2_b, (__nv_bfloat16*)zg, B, Cb, hw,
                                                              hidden);
  else return -1;
  DPMN_LAUNCH_CHECK();
  return 0;
}

// ---- de_1 tail: gather the 9 shifted tap products, NCHW fp32 out ---------------------------------------------
__global__ void __launch_bounds__(256) de1_gather_kernel(const float* __restrict__ P, int ldp,
                                                         const float* __restrict__ bias, float* __restrict__ out, int B,
                                                         int H, int W, int c_img, const float* __restrict__ blend,
                                                         long long blend_bs, float alpha) {
  const long long total = (long long)B * H * W;
  const long long HW = (long long)H * W;
  for (long long pix = (long long)blockIdx.x * blockDim.x + threadIdx.x; pix < total;
       pix += (long long)gridDim.x * blockDim.x) {
    const int xx = (int)(pix % W);
    const int yy = (int)((pix / W) % H);
    const int b = (int)(pix / HW);
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
      const int y2 = yy + 1 - ky;
      if (y2 < 0 || y2 >= H) continue;
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        const int x2 = xx + 1 - kx;
        if (x2 < 0 || x2 >= W) continue;
        const float* src = P + ((long long)b * HW + (long long)y2 * W + x2) * ldp + (ky * 3 + kx) * c_img;
        for (int co = 0; co < c_img; ++co) acc[co] += src[co];
      }
    }
    for (int co = 0; co < c_img; ++co) {
      float v = acc[co] + bias[co];
      if (blend != nullptr)      // alpha * CMM + (1 - alpha) * PSN image, super_resolution.py:449,705
        v = alpha * v + (1.0f - alpha) * blend[(long long)b * blend_bs + (long long)co * HW + (long long)yy * W + xx];
      out[((long long)b * c_img + co) * HW + (long long)yy * W + xx] = v;
    }
  }
}

__global__ void __launch_bounds__(256) alpha_blend_kernel(float* __restrict__ out, const float* __restrict__ blend, long long blend_bs,
                                                          float alpha, long long per_image, long long total) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long b = i / per_image, r = i - b * per_image;
    out[i] = alpha * out[i] + (1.0f - alpha) * blend[b * blend_bs + r];
  }
}

int launch_alpha_blend(float* out, const float* blend, long long blend_bs, float alpha, int B, long long per_image, cudaStream_t st) {
  const long long total = (long long)B * per_image;
  const int blocks = (int)((total + 255) / 256 < 148 * 8 ? (total + 255) / 256 : 148 * 8);
  alpha_blend_kernel<<<blocks, 256, 0, st>>>(out, blend, blend_bs, alpha, per_image, total);
  DPMN_LAUNCH_CHECK();
  return 0;
}

int launch_de1_gather(const float* P, int ldp, const float* bias, float* out, int B, int H, int W, int c_img,
                      cudaStream_t st, const float* blend, long long blend_bs, float alpha) {
  if (c_img > 4 || 9 * c_img > ldp) return -2;
  de1_gather_kernel<<<148 * 8, 256, 0, st>>>(P, ldp, bias, out, B, H, W, c_img, blend, blend_bs, alpha);
  DPMN_LAUNCH_CHECK();
  return 0;
}

}  // namespace dpmn
