// Backward kernels of the Complementation Modulation Module (fp32 SIMT, sm_100a)       cmm.py:38-161
//
// Data gradients of every conv / transposed conv reuse conv_simt_kernel with the roles swapped (the dgrad of a
// conv is the transposed conv with the same weight tensor and vice versa).  This file adds what is left:
//   conv_wgrad_simt   dW[co, k] += sum_n dy[co, n] * gather(k, n)   -- the forward's implicit-GEMM operand, so the
//                     producer's BatchNorm affine, the consumer's activation and the channel concat are applied
//                     on load exactly as in the forward
//   bn_act_bwd        sum over consumers of du * act'(bn(y)), BatchNorm backward (batch or running statistics),
//                     conv-bias gradient; one CTA per channel
//   se_gate_bwd       the bottleneck gate (cmm.py:135-147)
// Parity target: the reference's autograd gradients (tests/golden/cmm_*_grad.npz).
#include "common.cuh"
#include "kernels.h"

namespace dpmn {

constexpr int WM = 64, WN = 64, WK = 16, WPAD = 68;

// MR: output-channel rows per thread (tile = 16*MR channels x 64 k-indices); MR = 1 for Cout <= 16 (de_1: Cout = 3)
template <bool TRANSPOSED, int MR>
__global__ void __launch_bounds__(256) conv_wgrad_simt_kernel(ConvArgs p, const float* __restrict__ dy,
                                                              float* __restrict__ dw, int nsplit) {
  __shared__ __align__(16) float As[WK][WPAD];   // dy       [pixel][co]
  __shared__ __align__(16) float Bs[WK][WPAD];   // gathered [pixel][k]
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  constexpr int TM = 16 * MR;
  const int m0 = blockIdx.y * TM, k0 = blockIdx.x * WN;
  const int kk = p.k * p.k;
  const int K = p.Cin * kk;
  const int HoWo = p.Ho * p.Wo;
  const long long Ntot = (long long)p.B * HoWo;
  const long long HW = (long long)p.H * p.W;
  const long long chunk = ((Ntot + nsplit - 1) / nsplit + WK - 1) / WK * WK;
  const long long nbeg = (long long)blockIdx.z * chunk;
  const long long nend = nbeg + chunk < Ntot ? nbeg + chunk : Ntot;

  // gather loader: fixed k index (tid % 64), pixel rows (tid / 64) + 4 i
  const int gk = k0 + (tid & 63), gp0 = tid >> 6;
  const bool k_ok = gk < K;
  int ci = 0, ky = 0, kx = 0, seg = 0, cl = 0;
  if (k_ok) {
    ci = gk / kk;
    const int tap = gk - ci * kk;
    ky = tap / p.k; kx = tap - ky * p.k;
    cl = ci;
    if (p.n_seg > 1 && cl >= p.seg_ch[0]) {
      cl -= p.seg_ch[0]; seg = 1;
      if (p.n_seg > 2 && cl >= p.seg_ch[1]) { cl -= p.seg_ch[1]; seg = 2; }
    }
  }
  const float* src = p.in[seg];
  const int segC = p.seg_ch[seg];
  const bool has_aff = p.in_scale[seg] != nullptr;
  const float sc = (k_ok && has_aff) ? p.in_scale[seg][cl] : 1.f;
  const float sh = (k_ok && has_aff) ? p.in_shift[seg][cl] : 0.f;
  // dy loader: pixel (tid % 16), co rows (tid / 16) + 16 i
  const int ap = tid & 15, am0 = tid >> 4;

  float acc[MR][4];
#pragma unroll
  for (int i = 0; i < MR; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (long long n0 = nbeg; n0 < nend; n0 += WK) {
    float av[4], bv[4];
    {
      const long long n = n0 + ap;
      int b = 0, r = 0;
      if (n < nend) { b = (int)(n / HoWo); r = (int)(n - (long long)b * HoWo); }
#pragma unroll
      for (int i = 0; i < MR; ++i) {
        const int co = m0 + am0 + 16 * i;
        av[i] = (n < nend && co < p.Cout) ? dy[((long long)b * p.Cout + co) * HoWo + r] : 0.f;
      }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const long long n = n0 + gp0 + 4 * i;
      float v = 0.f;
      if (k_ok && n < nend) {
        const int b = (int)(n / HoWo);
        const int r = (int)(n - (long long)b * HoWo);
        const int oy = r / p.Wo, ox = r - oy * p.Wo;
        int iy, ix;
        bool ok;
        if (TRANSPOSED) {
          const int ty2 = oy + p.pad - ky * p.dil, tx2 = ox + p.pad - kx * p.dil;
          ok = ty2 >= 0 && tx2 >= 0 && (ty2 % p.stride) == 0 && (tx2 % p.stride) == 0;
          iy = ty2 / p.stride; ix = tx2 / p.stride;
          ok = ok && iy < p.H && ix < p.W;
        } else {
          iy = oy * p.stride - p.pad + ky * p.dil;
          ix = ox * p.stride - p.pad + kx * p.dil;
          ok = iy >= 0 && iy < p.H && ix >= 0 && ix < p.W;
        }
        if (ok) {
          v = src[((long long)b * segC + cl) * HW + (long long)iy * p.W + ix];
          v = fmaf(v, sc, sh);
          if (p.in_act == 1) v = v >= 0.f ? v : 0.2f * v;
          else if (p.in_act == 2) v = fmaxf(v, 0.f);
        }
      }
      bv[i] = v;
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      if (i < MR) As[ap][am0 + 16 * i] = av[i];
      Bs[gp0 + 4 * i][tid & 63] = bv[i];
    }
    __syncthreads();
#pragma unroll
    for (int q = 0; q < WK; ++q) {
      float ar[MR];
      if constexpr (MR == 4) {
        const float4 a4 = *reinterpret_cast<const float4*>(&As[q][ty * 4]);
        ar[0] = a4.x; ar[1] = a4.y; ar[2] = a4.z; ar[3] = a4.w;
      } else {
#pragma unroll
        for (int i = 0; i < MR; ++i) ar[i] = As[q][ty * MR + i];
      }
      const float4 b4 = *reinterpret_cast<const float4*>(&Bs[q][tx * 4]);
      const float br[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
      for (int i = 0; i < MR; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(ar[i], br[j], acc[i][j]);
    }
  }
#pragma unroll
  for (int i = 0; i < MR; ++i) {
    const int co = m0 + ty * MR + i;
    if (co >= p.Cout) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int k = k0 + tx * 4 + j;
      if (k >= K) continue;
      long long o;
      if (TRANSPOSED) { const int c2 = k / kk; o = ((long long)c2 * p.Cout + co) * kk + (k - c2 * kk); }
      else o = (long long)co * K + k;
      atomicAdd(dw + o, acc[i][j]);
    }
  }
}

// Weight gradient of a stride-2 transposed conv, parity-class form (see conv_simt_t2_kernel): class (py, px) pairs the
// output pixels of that parity with the taps that are valid for it, so no product multiplies a structural zero.
// blockIdx.z = class * nsplit + pixel chunk.
__global__ void __launch_bounds__(256) conv_wgrad_simt_t2_kernel(ConvArgs p, const float* __restrict__ dy,
                                                                 float* __restrict__ dw, int nsplit) {
  __shared__ __align__(16) float As[WK][WPAD];
  __shared__ __align__(16) float Bs[WK][WPAD];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.y * WM, k0 = blockIdx.x * WN;
  const int cls = blockIdx.z / nsplit, chunk_id = blockIdx.z - cls * nsplit;
  const int py = cls >> 1, px = cls & 1;
  int kys[4], kxs[4], nky = 0, nkx = 0;
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    if (t < p.k && (((py + p.pad + t * p.dil) & 1) == 0)) kys[nky++] = t;
    if (t < p.k && (((px + p.pad + t * p.dil) & 1) == 0)) kxs[nkx++] = t;
  }
  const int nt = nky * nkx;
  const int kk = p.k * p.k;
  const int K = p.Cin * nt;
  if (k0 >= K) return;
  const int Hq = p.Ho >> 1, Wq = p.Wo >> 1, HqWq = Hq * Wq, HoWo = p.Ho * p.Wo;
  const long long Ntot = (long long)p.B * HqWq;
  const long long HW = (long long)p.H * p.W;
  const long long chunk = ((Ntot + nsplit - 1) / nsplit + WK - 1) / WK * WK;
  const long long nbeg = (long long)chunk_id * chunk;
  const long long nend = nbeg + chunk < Ntot ? nbeg + chunk : Ntot;

  const int gk = k0 + (tid & 63), gp0 = tid >> 6;
  const bool k_ok = gk < K;
  int ky = 0, kx = 0, seg = 0, cl = 0;
  if (k_ok) {
    const int ci = gk / nt, r = gk - ci * nt;
    const int a = r / nkx;
    ky = kys[a]; kx = kxs[r - a * nkx];
    cl = ci;
    if (p.n_seg > 1 && cl >= p.seg_ch[0]) {
      cl -= p.seg_ch[0]; seg = 1;
      if (p.n_seg > 2 && cl >= p.seg_ch[1]) { cl -= p.seg_ch[1]; seg = 2; }
    }
  }
  const float* src = p.in[seg];
  const int segC = p.seg_ch[seg];
  const bool has_aff = p.in_scale[seg] != nullptr;
  const float sc = (k_ok && has_aff) ? p.in_scale[seg][cl] : 1.f;
  const float sh = (k_ok && has_aff) ? p.in_shift[seg][cl] : 0.f;
  const int ap = tid & 15, am0 = tid >> 4;

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (long long n0 = nbeg; n0 < nend; n0 += WK) {
    float av[4], bv[4];
    {
      const long long n = n0 + ap;
      long long o = 0;
      if (n < nend) {
        const int b = (int)(n / HqWq);
        const int r = (int)(n - (long long)b * HqWq);
        const int qy = r / Wq, qx = r - qy * Wq;
        o = (long long)b * p.Cout * HoWo + (long long)(2 * qy + py) * p.Wo + 2 * qx + px;
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int co = m0 + am0 + 16 * i;
        av[i] = (n < nend && co < p.Cout) ? dy[o + (long long)co * HoWo] : 0.f;
      }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const long long n = n0 + gp0 + 4 * i;
      float v = 0.f;
      if (k_ok && n < nend) {
        const int b = (int)(n / HqWq);
        const int r = (int)(n - (long long)b * HqWq);
        const int qy = r / Wq, qx = r - qy * Wq;
        const int ty2 = 2 * qy + py + p.pad - ky * p.dil, tx2 = 2 * qx + px + p.pad - kx * p.dil;
        const int iy = ty2 >> 1, ix = tx2 >> 1;
        if (ty2 >= 0 && tx2 >= 0 && iy < p.H && ix < p.W) {
          v = src[((long long)b * segC + cl) * HW + (long long)iy * p.W + ix];
          v = fmaf(v, sc, sh);
          if (p.in_act == 1) v = v >= 0.f ? v : 0.2f * v;
          else if (p.in_act == 2) v = fmaxf(v, 0.f);
        }
      }
      bv[i] = v;
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      As[ap][am0 + 16 * i] = av[i];
      Bs[gp0 + 4 * i][tid & 63] = bv[i];
    }
    __syncthreads();
#pragma unroll
    for (int q = 0; q < WK; ++q) {
      const float4 a4 = *reinterpret_cast<const float4*>(&As[q][ty * 4]);
      const float4 b4 = *reinterpret_cast<const float4*>(&Bs[q][tx * 4]);
      const float ar[4] = {a4.x, a4.y, a4.z, a4.w};
      const float br[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(ar[i], br[j], acc[i][j]);
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int co = m0 + ty * 4 + i;
    if (co >= p.Cout) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int k = k0 + tx * 4 + j;
      if (k >= K) continue;
      const int c2 = k / nt, r = k - c2 * nt;
      const int a = r / nkx;
      atomicAdd(dw + ((long long)c2 * p.Cout + co) * kk + kys[a] * p.k + kxs[r - a * nkx], acc[i][j]);
    }
  }
}

// `a` describes the FORWARD conv (inputs / affines / activation / geometry); dy is d(raw conv output) (B, Cout, Ho, Wo).
int launch_conv_wgrad_simt(const ConvArgs& a, const float* dy, float* dw, cudaStream_t st) {
  if (a.n_seg < 1 || a.n_seg > 3) return -1;
  int cin = 0;
  for (int i = 0; i < a.n_seg; ++i) cin += a.seg_ch[i];
  if (cin != a.Cin) return -1;
  if (a.transposed && a.stride == 2 && a.k <= 4 && a.Ho % 2 == 0 && a.Wo % 2 == 0) {
    // parity-class form: per class at most Cin * ceil(k/2)^2 (dil 1) or Cin * k^2 (dil 2, one live class) taps
    const int ntmax = a.dil % 2 == 0 ? a.k * a.k : ((a.k + 1) / 2) * ((a.k + 1) / 2);
    const int Kc = a.Cin * ntmax;
    const long long Nc = (long long)a.B * (a.Ho / 2) * (a.Wo / 2);
    const int tiles_c = ((Kc + WN - 1) / WN) * ((a.Cout + WM - 1) / WM);
    long long nsp = 1184 / (4 * tiles_c);
    if (nsp > Nc / 64) nsp = Nc / 64;
    if (nsp < 1) nsp = 1;
    if (nsp > 128) nsp = 128;
    dim3 g2((Kc + WN - 1) / WN, (a.Cout + WM - 1) / WM, (unsigned)(4 * nsp));
    conv_wgrad_simt_t2_kernel<<<g2, 256, 0, st>>>(a, dy, dw, (int)nsp);
    DPMN_LAUNCH_CHECK();
    return 0;
  }
  const int K = a.Cin * a.k * a.k;
  const long long Ntot = (long long)a.B * a.Ho * a.Wo;
  const int tm = a.Cout <= 16 ? 16 : WM;
  const int tiles = ((K + WN - 1) / WN) * ((a.Cout + tm - 1) / tm);
  long long ns = 1184 / tiles;
  if (ns > Ntot / 64) ns = Ntot / 64;
  if (ns < 1) ns = 1;
  if (ns > 512) ns = 512;
  dim3 grid((K + WN - 1) / WN, (a.Cout + tm - 1) / tm, (unsigned)ns);
  if (a.Cout <= 16) {
    if (a.transposed) conv_wgrad_simt_kernel<true, 1><<<grid, 256, 0, st>>>(a, dy, dw, (int)ns);
    else conv_wgrad_simt_kernel<false, 1><<<grid, 256, 0, st>>>(a, dy, dw, (int)ns);
  } else if (a.transposed) conv_wgrad_simt_kernel<true, 4><<<grid, 256, 0, st>>>(a, dy, dw, (int)ns);
  else conv_wgrad_simt_kernel<false, 4><<<grid, 256, 0, st>>>(a, dy, dw, (int)ns);
  DPMN_LAUNCH_CHECK();
  return 0;
}

// -----------------------------------------------------------------------------------------------------
// d(raw conv output y) from the gradients of its consumers' inputs.   bn = sc*y + sh (or y when there is no
// BatchNorm);  dbn = sum_s du_s * act_s'(bn);  train: dy = sc * (dbn - mean(dbn) - xhat * mean(dbn*xhat));
// eval: dy = sc * dbn.  d gamma += sum dbn*xhat, d beta += sum dbn, d bias(conv) += sum dy.  One CTA / channel.
// -----------------------------------------------------------------------------------------------------
__device__ __forceinline__ float block_sum_256b(float v, float* red) {
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += red[i];
  return s;
}

__device__ __forceinline__ float act_grad(int act, float v) {
  return act == 1 ? (v >= 0.f ? 1.f : 0.2f) : act == 2 ? (v > 0.f ? 1.f : 0.f) : 1.f;
}

__global__ void __launch_bounds__(256) bn_act_bwd_kernel(BnBwdArgs p) {
  __shared__ float red[8];
  const int c = blockIdx.x;
  const long long n = (long long)p.B * p.HW;
  const bool has_bn = p.gamma != nullptr;
  float mean = 0.f, rstd = 1.f, sc = 1.f, sh = 0.f;
  if (has_bn) {
    float var;
    if (p.training) {
      float s = 0.f;
      for (long long i = threadIdx.x; i < n; i += 256) {
        const int b = (int)(i / p.HW); const int r = (int)(i - (long long)b * p.HW);
        s += p.y[((long long)b * p.C + c) * p.HW + r];
      }
      mean = block_sum_256b(s, red) / (float)n;
      float q = 0.f;
      for (long long i = threadIdx.x; i < n; i += 256) {
        const int b = (int)(i / p.HW); const int r = (int)(i - (long long)b * p.HW);
        const float d = p.y[((long long)b * p.C + c) * p.HW + r] - mean;
        q = fmaf(d, d, q);
      }
      var = block_sum_256b(q, red) / (float)n;
    } else {
      mean = p.run_mean[c]; var = p.run_var[c];
    }
    rstd = 1.0f / sqrtf(var + p.eps);
    sc = p.gamma[c] * rstd;
    sh = p.beta[c] - mean * sc;
  }
  auto dbn_at = [&](int b, int r, float yv) -> float {
    const float bn = fmaf(yv, sc, sh);
    float g = 0.f;
    for (int s = 0; s < p.n_src; ++s)
      g = fmaf(p.du[s][(long long)b * p.du_bs[s] + (long long)c * p.HW + r], act_grad(p.act[s], bn), g);
    return g;
  };
  float m1 = 0.f, m2 = 0.f;
  if (has_bn) {
    float s1 = 0.f, s2 = 0.f;
    for (long long i = threadIdx.x; i < n; i += 256) {
      const int b = (int)(i / p.HW); const int r = (int)(i - (long long)b * p.HW);
      const float yv = p.y[((long long)b * p.C + c) * p.HW + r];
      const float g = dbn_at(b, r, yv);
      s1 += g;
      s2 = fmaf(g, (yv - mean) * rstd, s2);
    }
    s1 = block_sum_256b(s1, red);
    s2 = block_sum_256b(s2, red);
    if (threadIdx.x == 0) { atomicAdd(p.dgamma + c, s2); atomicAdd(p.dbeta + c, s1); }
    if (p.training) { m1 = s1 / (float)n; m2 = s2 / (float)n; }
  }
  float sb = 0.f;
  for (long long i = threadIdx.x; i < n; i += 256) {
    const int b = (int)(i / p.HW); const int r = (int)(i - (long long)b * p.HW);
    const long long o = ((long long)b * p.C + c) * p.HW + r;
    const float yv = p.y[o];
    const float g = dbn_at(b, r, yv);
    const float d = has_bn ? sc * (g - m1 - (yv - mean) * rstd * m2) : g;
    p.dy[o] = d;
    sb += d;
  }
  if (p.dbias != nullptr) {
    sb = block_sum_256b(sb, red);
    if (threadIdx.x == 0) atomicAdd(p.dbias + c, sb);
  }
}

// ---- the same computation for layers with few channels and many pixels (C = 64..128 at 32x128 / 16x64: one CTA per
// channel leaves most SMs idle and each CTA streams 200 K elements three times).  Three fully parallel passes over
// (channel, pixel chunk) with per-channel atomics into a small scratch: [C][4] = sum(y - ref), sum((y - ref)^2),
// sum(dbn), sum(dbn * xhat); ref = the channel's first element (a shift that keeps the one-pass variance well
// conditioned).
__global__ void __launch_bounds__(256) bn_stats_part_kernel(BnBwdArgs p, float* __restrict__ scratch, int chunk) {
  __shared__ float red[8];
  const int c = blockIdx.x;
  const long long n = (long long)p.B * p.HW;
  const long long i0 = (long long)blockIdx.y * chunk, i1 = min(n, i0 + chunk);
  const float ref = p.y[(long long)c * p.HW];
  float s = 0.f, q = 0.f;
  for (long long i = i0 + threadIdx.x; i < i1; i += 256) {
    const int b = (int)(i / p.HW); const int r = (int)(i - (long long)b * p.HW);
    const float d = p.y[((long long)b * p.C + c) * p.HW + r] - ref;
    s += d; q = fmaf(d, d, q);
  }
  s = block_sum_256b(s, red);
  q = block_sum_256b(q, red);
  if (threadIdx.x == 0) { atomicAdd(scratch + c * 4 + 0, s); atomicAdd(scratch + c * 4 + 1, q); }
}

__device__ __forceinline__ void bn_channel_affine(const BnBwdArgs& p, const float* scratch, int c, float& mean, float& rstd,
                                                  float& sc, float& sh) {
  mean = 0.f; rstd = 1.f; sc = 1.f; sh = 0.f;
  if (p.gamma == nullptr) return;
  float var;
  if (p.training) {
    const float n = (float)((long long)p.B * p.HW);
    const float ref = p.y[(long long)c * p.HW];
    const float ms = scratch[c * 4 + 0] / n;
    mean = ref + ms;
    var = fmaxf(scratch[c * 4 + 1] / n - ms * ms, 0.f);
  } else {
    mean = p.run_mean[c]; var = p.run_var[c];
  }
  rstd = 1.0f / sqrtf(var + p.eps);
  sc = p.gamma[c] * rstd;
  sh = p.beta[c] - mean * sc;
}

template <bool APPLY>   // false: accumulate sum(dbn), sum(dbn*xhat); true: write dy (+ bias gradient)
__global__ void __launch_bounds__(256) bn_dbn_part_kernel(BnBwdArgs p, float* __restrict__ scratch, int chunk) {
  __shared__ float red[8];
  const int c = blockIdx.x;
  const long long n = (long long)p.B * p.HW;
  const long long i0 = (long long)blockIdx.y * chunk, i1 = min(n, i0 + chunk);
  float mean, rstd, sc, sh;
  bn_channel_affine(p, scratch, c, mean, rstd, sc, sh);
  const bool has_bn = p.gamma != nullptr;
  float m1 = 0.f, m2 = 0.f;
  if (APPLY && has_bn && p.training) { m1 = scratch[c * 4 + 2] / (float)n; m2 = scratch[c * 4 + 3] / (float)n; }
  float a0 = 0.f, a1 = 0.f;
  for (long long i = i0 + threadIdx.x; i < i1; i += 256) {
    const int b = (int)(i / p.HW); const int r = (int)(i - (long long)b * p.HW);
    const long long o = ((long long)b * p.C + c) * p.HW + r;
    const float yv = p.y[o];
    const float bn = fmaf(yv, sc, sh);
    float g = 0.f;
    for (int s = 0; s < p.n_src; ++s)
      g = fmaf(p.du[s][(long long)b * p.du_bs[s] + (long long)c * p.HW + r], act_grad(p.act[s], bn), g);
    const float xh = (yv - mean) * rstd;
    if (APPLY) {
      const float d = has_bn ? sc * (g - m1 - xh * m2) : g;
      p.dy[o] = d;
      a0 += d;
    } else {
      a0 += g;
      a1 = fmaf(g, xh, a1);
    }
  }
  a0 = block_sum_256b(a0, red);
  if (APPLY) {
    if (threadIdx.x == 0 && p.dbias != nullptr) atomicAdd(p.dbias + c, a0);
  } else {
    a1 = block_sum_256b(a1, red);
    if (threadIdx.x == 0) { atomicAdd(scratch + c * 4 + 2, a0); atomicAdd(scratch + c * 4 + 3, a1); }
  }
}

__global__ void bn_param_grad_kernel(BnBwdArgs p, const float* __restrict__ scratch) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < p.C) { atomicAdd(p.dgamma + c, scratch[c * 4 + 3]); atomicAdd(p.dbeta + c, scratch[c * 4 + 2]); }
}

int launch_bn_act_bwd(const BnBwdArgs& a, cudaStream_t st) {
  if (a.n_src < 1 || a.n_src > 2) return -1;
  const long long n = (long long)a.B * a.HW;
  if (a.scratch != nullptr && a.C <= 256 && n >= 8192) {
    const int chunk = 8192;
    const dim3 grid(a.C, (unsigned)((n + chunk - 1) / chunk));
    DPMN_CUDA_TRY(cudaMemsetAsync(a.scratch, 0, (size_t)a.C * 4 * sizeof(float), st));
    if (a.gamma != nullptr) {
      if (a.training) { bn_stats_part_kernel<<<grid, 256, 0, st>>>(a, a.scratch, chunk); DPMN_LAUNCH_CHECK(); }
      bn_dbn_part_kernel<false><<<grid, 256, 0, st>>>(a, a.scratch, chunk);
      DPMN_LAUNCH_CHECK();
      bn_param_grad_kernel<<<(a.C + 127) / 128, 128, 0, st>>>(a, a.scratch);
      DPMN_LAUNCH_CHECK();
    }
    bn_dbn_part_kernel<true><<<grid, 256, 0, st>>>(a, a.scratch, chunk);
    DPMN_LAUNCH_CHECK();
    return 0;
  }
  bn_act_bwd_kernel<<<a.C, 256, 0, st>>>(a);
  DPMN_LAUNCH_CHECK();
  return 0;
}

// out[c] += sum_{b, r} x[(b*C + c)*HW + r]          (bias gradient of a conv whose output gradient is given directly)
// grid (C, S): the S blocks of a channel split the (b, r) range (de_1 has 3 channels and 48 x 4096 values per channel: one
// block per channel took 440 us).
__global__ void __launch_bounds__(256) chan_sum_kernel(const float* __restrict__ x, int B, int C, int HW, float* __restrict__ out) {
  __shared__ float red[8];
  const int c = blockIdx.x;
  const long long n = (long long)B * HW;
  float s = 0.f;
  for (long long i = (long long)blockIdx.y * 256 + threadIdx.x; i < n; i += 256LL * gridDim.y) {
    const int b = (int)(i / HW); const int r = (int)(i - (long long)b * HW);
    s += x[((long long)b * C + c) * HW + r];
  }
  s = block_sum_256b(s, red);
  if (threadIdx.x == 0) atomicAdd(out + c, s);
}

int launch_chan_sum(const float* x, int B, int C, int HW, float* out, cudaStream_t st) {
  const long long n = (long long)B * HW;
  long long split = (n + 2047) / 2048;                   // >= 8 values per thread
  const long long cap = (148 * 8 + C - 1) / C;           // about 8 blocks per SM over all channels
  if (split > cap) split = cap;
  if (split < 1) split = 1;
  chan_sum_kernel<<<dim3(C, (unsigned)split), 256, 0, st>>>(x, B, C, HW, out);
  DPMN_LAUNCH_CHECK();
  return 0;
}

// -----------------------------------------------------------------------------------------------------
// SE gate backward (cmm.py:135-147): z = cat(z1, z2); g0 = mean_hw z; h = relu(fc1 g0 + b1); g = sigmoid(fc2 h + b2);
// out = z*g + z feeds de_6 through ReLU.  du = d relu(out) (B, 2Cb, hw).  One CTA per image.
// -----------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) se_gate_bwd_kernel(const float* __restrict__ z1, const float* __restrict__ z2,
                                                          const float* __restrict__ du, const float* __restrict__ fc1_w,
                                                          const float* __restrict__ fc1_b, const float* __restrict__ fc2_w,
                                                          const float* __restrict__ fc2_b, float* __restrict__ dz1,
                                                          float* __restrict__ dz2, float* __restrict__ d_fc1_w,
                                                          float* __restrict__ d_fc1_b, float* __restrict__ d_fc2_w,
                                                          float* __restrict__ d_fc2_b, int Cb, int hw, int hidden,
                                                          float* __restrict__ scratch) {
  // scratch != nullptr: the per-image factors of the two weight gradients (sds, sh, sdh, sg0) are written there and
  // se_gate_wgrad_kernel sums the outer products over the batch -- instead of 2 * C2 * hidden atomicAdds per image
  extern __shared__ float sm[];
  const int C2 = 2 * Cb;
  float* sg0 = sm;             // [C2] pooled
  float* sgate = sg0 + C2;     // [C2] sigmoid
  float* sds = sgate + C2;     // [C2] d (pre-sigmoid)
  float* sh = sds + C2;        // [hidden] relu(fc1)
  float* sdh = sh + hidden;    // [hidden] d (pre-relu)
  const int b = blockIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int c = threadIdx.x; c < C2; c += blockDim.x) {
    const float* src = c < Cb ? z1 + ((long long)b * Cb + c) * hw : z2 + ((long long)b * Cb + (c - Cb)) * hw;
    float s = 0.f;
    for (int i = 0; i < hw; ++i) s += src[i];
    sg0[c] = s / (float)hw;
  }
  __syncthreads();
  for (int j = warp; j < hidden; j += nw) {
    float s = 0.f;
    for (int c = lane; c < C2; c += 32) s = fmaf(fc1_w[(long long)j * C2 + c], sg0[c], s);
    s = warp_sum(s);
    if (lane == 0) sh[j] = fmaxf(s + fc1_b[j], 0.f);
  }
  __syncthreads();
  for (int c = warp; c < C2; c += nw) {
    float s = 0.f;
    for (int j = lane; j < hidden; j += 32) s = fmaf(fc2_w[(long long)c * hidden + j], sh[j], s);
    s = warp_sum(s);
    const float g = 1.0f / (1.0f + expf(-(s + fc2_b[c])));
    // d gate[c] = sum_hw dout * z, dout = du * [z*(1+g) > 0]
    const float* src = c < Cb ? z1 + ((long long)b * Cb + c) * hw : z2 + ((long long)b * Cb + (c - Cb)) * hw;
    float dg = 0.f;
    for (int i = lane; i < hw; i += 32) {
      const float v = src[i];
      const float o = fmaf(v, g, v);
      if (o > 0.f) dg = fmaf(du[((long long)b * C2 + c) * hw + i], v, dg);
    }
    dg = warp_sum(dg);
    if (lane == 0) {
      sgate[c] = g;
      const float ds = dg * g * (1.0f - g);
      sds[c] = ds;
      if (scratch == nullptr) atomicAdd(d_fc2_b + c, ds);
    }
  }
  __syncthreads();
  if (scratch == nullptr) {
    for (int i = threadIdx.x; i < C2 * hidden; i += blockDim.x) {
      const int c = i / hidden, j = i - c * hidden;
      atomicAdd(d_fc2_w + i, sds[c] * sh[j]);
    }
  }
  for (int j = warp; j < hidden; j += nw) {
    float s = 0.f;
    for (int c = lane; c < C2; c += 32) s = fmaf(sds[c], fc2_w[(long long)c * hidden + j], s);
    s = warp_sum(s);
    if (lane == 0) {
      const float d = sh[j] > 0.f ? s : 0.f;
      sdh[j] = d;
      if (scratch == nullptr) atomicAdd(d_fc1_b + j, d);
    }
  }
  __syncthreads();
  if (scratch == nullptr) {
    for (int i = threadIdx.x; i < hidden * C2; i += blockDim.x) {
      const int j = i / C2, c = i - j * C2;
      atomicAdd(d_fc1_w + i, sdh[j] * sg0[c]);
    }
  } else {
    float* row = scratch + (long long)b * (2 * C2 + 2 * hidden);      // [sds C2][sg0 C2][sh hidden][sdh hidden]
    for (int c = threadIdx.x; c < C2; c += blockDim.x) { row[c] = sds[c]; row[C2 + c] = sg0[c]; }
    for (int j = threadIdx.x; j < hidden; j += blockDim.x) { row[2 * C2 + j] = sh[j]; row[2 * C2 + hidden + j] = sdh[j]; }
  }
  for (int c = warp; c < C2; c += nw) {
    float s = 0.f;
    for (int j = lane; j < hidden; j += 32) s = fmaf(sdh[j], fc1_w[(long long)j * C2 + c], s);
    s = warp_sum(s) / (float)hw;                     // d pooled -> every pixel
    const float g = sgate[c];
    const float* src = c < Cb ? z1 + ((long long)b * Cb + c) * hw : z2 + ((long long)b * Cb + (c - Cb)) * hw;
    float* dst = c < Cb ? dz1 + ((long long)b * Cb + c) * hw : dz2 + ((long long)b * Cb + (c - Cb)) * hw;
    for (int i = lane; i < hw; i += 32) {
      const float v = src[i];
      const float o = fmaf(v, g, v);
      const float dout = o > 0.f ? du[((long long)b * C2 + c) * hw + i] : 0.f;
      dst[i] = fmaf(dout, 1.0f + g, s);
    }
  }
}

// d_fc2_w[c][j] += sum_b sds[b][c] sh[b][j];  d_fc1_w[j][c] += sum_b sdh[b][j] sg0[b][c];  biases likewise
__global__ void __launch_bounds__(256) se_gate_wgrad_kernel(const float* __restrict__ scratch, float* __restrict__ d_fc1_w,
                                                            float* __restrict__ d_fc1_b, float* __restrict__ d_fc2_w,
                                                            float* __restrict__ d_fc2_b, int B, int C2, int hidden) {
  const int stride = 2 * C2 + 2 * hidden;
  const long long n_w = (long long)C2 * hidden;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < 2 * n_w + C2 + hidden;
       i += (long long)gridDim.x * blockDim.x) {
    float a = 0.f;
    if (i < n_w) {                                    // fc2 weight (C2, hidden)
      const int c = (int)(i / hidden), j = (int)(i - (long long)c * hidden);
      for (int b = 0; b < B; ++b) a = fmaf(scratch[(long long)b * stride + c], scratch[(long long)b * stride + 2 * C2 + j], a);
      d_fc2_w[i] += a;
    } else if (i < 2 * n_w) {                         // fc1 weight (hidden, C2)
      const long long k = i - n_w;
      const int j = (int)(k / C2), c = (int)(k - (long long)j * C2);
      for (int b = 0; b < B; ++b)
        a = fmaf(scratch[(long long)b * stride + 2 * C2 + hidden + j], scratch[(long long)b * stride + C2 + c], a);
      d_fc1_w[k] += a;
    } else if (i < 2 * n_w + C2) {
      const int c = (int)(i - 2 * n_w);
      for (int b = 0; b < B; ++b) a += scratch[(long long)b * stride + c];
      d_fc2_b[c] += a;
    } else {
      const int j = (int)(i - 2 * n_w - C2);
      for (int b = 0; b < B; ++b) a += scratch[(long long)b * stride + 2 * C2 + hidden + j];
      d_fc1_b[j] += a;
    }
  }
}

// The backward over many CTAs (the one-CTA-per-image kernel above took 557 us at batch 48).  scratch (floats):
//   rows   [B][sds C2 | sg0 C2 | sh hidden | sdh hidden]   (what se_gate_wgrad_kernel reads)
//   sgate  [B][C2]
//   raw    [B][hidden]   unmasked d(pre-ReLU) sums, accumulated by atomics over 8 channel slices (zeroed first)
__global__ void __launch_bounds__(256) se_bwd_gate_kernel(const float* __restrict__ z1, const float* __restrict__ z2,
                                                          const float* __restrict__ du, const float* __restrict__ fc2_w,
                                                          const float* __restrict__ fc2_b, float* __restrict__ scratch,
                                                          int B, int Cb, int hw, int hidden) {
  extern __shared__ float sm[];
  const int C2 = 2 * Cb, b = blockIdx.x, stride = 2 * C2 + 2 * hidden;
  float* row = scratch + (long long)b * stride;
  float* sgate = scratch + (long long)B * stride + (long long)b * C2;
  for (int j = threadIdx.x; j < hidden; j += blockDim.x) sm[j] = row[2 * C2 + j];
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int cc = warp; cc < 64; cc += 8) {
    const int c = blockIdx.y * 64 + cc;
    if (c >= C2) break;
    float s = 0.f;
    for (int j = lane; j < hidden; j += 32) s = fmaf(fc2_w[(long long)c * hidden + j], sm[j], s);
    s = warp_sum(s);
    const float g = 1.0f / (1.0f + expf(-(s + fc2_b[c])));
    // d gate[c] = sum_hw dout * z, dout = du * [z*(1+g) > 0]
    const float* src = c < Cb ? z1 + ((long long)b * Cb + c) * hw : z2 + ((long long)b * Cb + (c - Cb)) * hw;
    float dg = 0.f;
    for (int i = lane; i < hw; i += 32) {
      const float v = src[i];
      const float o = fmaf(v, g, v);
      if (o > 0.f) dg = fmaf(du[((long long)b * C2 + c) * hw + i], v, dg);
    }
    dg = warp_sum(dg);
    if (lane == 0) { sgate[c] = g; row[c] = dg * g * (1.0f - g); }
  }
}

// raw[b][j] += sum over this block's channel slice of sds[b][c] fc2_w[c][j]        grid (B, 8)
__global__ void __launch_bounds__(256) se_bwd_dh_kernel(const float* __restrict__ fc2_w, float* __restrict__ scratch, int B,
                                                        int C2, int hidden) {
  extern __shared__ float sm[];
  const int b = blockIdx.x, stride = 2 * C2 + 2 * hidden;
  const float* row = scratch + (long long)b * stride;
  float* raw = scratch + (long long)B * (stride + C2) + (long long)b * hidden;
  const int per = (C2 + gridDim.y - 1) / gridDim.y, c0 = blockIdx.y * per, c1 = min(C2, c0 + per);
  for (int c = c0 + threadIdx.x; c < c1; c += blockDim.x) sm[c - c0] = row[c];
  __syncthreads();
  for (int j = threadIdx.x; j < hidden; j += blockDim.x) {
    float s = 0.f;
#pragma unroll 8
    for (int c = c0; c < c1; ++c) s = fmaf(sm[c - c0], fc2_w[(long long)c * hidden + j], s);
    atomicAdd(raw + j, s);
  }
}

// sdh = raw masked by the ReLU; d pooled[c] = sum_j sdh[j] fc1_w[j][c] / hw; dz = dout (1 + g) + d pooled      grid (B, C2 / 256)
__global__ void __launch_bounds__(256) se_bwd_dz_kernel(const float* __restrict__ z1, const float* __restrict__ z2,
                                                        const float* __restrict__ du, const float* __restrict__ fc1_w,
                                                        float* __restrict__ dz1, float* __restrict__ dz2,
                                                        float* __restrict__ scratch, int B, int Cb, int hw, int hidden) {
  extern __shared__ float sm[];
  const int C2 = 2 * Cb, b = blockIdx.x, stride = 2 * C2 + 2 * hidden;
  float* row = scratch + (long long)b * stride;
  const float* sgate = scratch + (long long)B * stride + (long long)b * C2;
  const float* raw = scratch + (long long)B * (stride + C2) + (long long)b * hidden;
  for (int j = threadIdx.x; j < hidden; j += blockDim.x) {
    const float d = row[2 * C2 + j] > 0.f ? raw[j] : 0.f;
    sm[j] = d;
    if (blockIdx.y == 0) row[2 * C2 + hidden + j] = d;
  }
  __syncthreads();
  const int c = blockIdx.y * blockDim.x + threadIdx.x;
  if (c >= C2) return;
  float s = 0.f;
#pragma unroll 8
  for (int j = 0; j < hidden; ++j) s = fmaf(sm[j], fc1_w[(long long)j * C2 + c], s);
  s /= (float)hw;                                    // d pooled -> every pixel
  const float g = sgate[c];
  const float* src = c < Cb ? z1 + ((long long)b * Cb + c) * hw : z2 + ((long long)b * Cb + (c - Cb)) * hw;
  float* dst = c < Cb ? dz1 + ((long long)b * Cb + c) * hw : dz2 + ((long long)b * Cb + (c - Cb)) * hw;
  for (int i = 0; i < hw; ++i) {
    const float v = src[i];
    const float o = fmaf(v, g, v);
    const float dout = o > 0.f ? du[((long long)b * C2 + c) * hw + i] : 0.f;
    dst[i] = fmaf(dout, 1.0f + g, s);
  }
}

// scratch: B * (3 * 2Cb + 3 * hidden) floats (layout above), or nullptr for the one-CTA-per-image kernel with atomics.
int launch_se_gate_bwd(const float* z1, const float* z2, const float* du, const float* fc1_w, const float* fc1_b,
                       const float* fc2_w, const float* fc2_b, float* dz1, float* dz2, float* d_fc1_w, float* d_fc1_b,
                       float* d_fc2_w, float* d_fc2_b, int B, int Cb, int hw, int hidden, cudaStream_t st, float* scratch) {
  const size_t smem = (size_t)(6 * Cb + 2 * hidden) * sizeof(float);
  if (smem > 48 * 1024) return -2;
  if (scratch != nullptr) {
    const int C2 = 2 * Cb, stride = 2 * C2 + 2 * hidden;
    float* raw = scratch + (long long)B * (stride + C2);
    DPMN_CUDA_TRY(cudaMemsetAsync(raw, 0, (size_t)B * hidden * sizeof(float), st));
    // pooled -> rows[.][C2 ..], h -> rows[.][2 C2 ..]
    const int rc = launch_se_pool_fc1(z1, z2, fc1_w, fc1_b, B, Cb, hw, hidden, scratch + 2 * C2, stride, scratch + C2, stride, st);
    if (rc) return rc;
    se_bwd_gate_kernel<<<dim3(B, (C2 + 63) / 64), 256, (size_t)hidden * sizeof(float), st>>>(z1, z2, du, fc2_w, fc2_b, scratch, B, Cb,
                                                                                           hw, hidden);
    DPMN_LAUNCH_CHECK();
    se_bwd_dh_kernel<<<dim3(B, 8), 256, (size_t)((C2 + 7) / 8) * sizeof(float), st>>>(fc2_w, scratch, B, C2, hidden);
    DPMN_LAUNCH_CHECK();
    se_bwd_dz_kernel<<<dim3(B, (C2 + 255) / 256), 256, (size_t)hidden * sizeof(float), st>>>(z1, z2, du, fc1_w, dz1, dz2, scratch, B,
                                                                                           Cb, hw, hidden);
    DPMN_LAUNCH_CHECK();
    se_gate_wgrad_kernel<<<148 * 4, 256, 0, st>>>(scratch, d_fc1_w, d_fc1_b, d_fc2_w, d_fc2_b, B, C2, hidden);
    DPMN_LAUNCH_CHECK();
    return 0;
  }
  se_gate_bwd_kernel<<<B, 256, smem, st>>>(z1, z2, du, fc1_w, fc1_b, fc2_w, fc2_b, dz1, dz2, d_fc1_w, d_fc1_b, d_fc2_w,
                                           d_fc2_b, Cb, hw, hidden, scratch);
  DPMN_LAUNCH_CHECK();
  return 0;
}

}  // namespace dpmn
