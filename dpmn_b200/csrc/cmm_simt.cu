// SIMT fp32 kernels of the Complementation Modulation Module (exact-arithmetic mode).
// Behaviour restated from /root/reference/model/cmm.py:10-161 (never copied).
//
// Every conv / transposed conv of the U-Net is one implicit GEMM:
//     out[co, n] = bias[co] + sum_k  Wmat[co, k] * gather(k, n)        M = Cout, N = B*Ho*Wo, K = Cin*k*k
// where gather() reads the channel-concatenated input, applies the producer's BatchNorm as a per-channel
// affine and then the consumer's leading activation -- so BN and activations never cost a pass of their own.
#include "common.cuh"
#include "kernels.h"

namespace dpmn {

constexpr int CM = 64, CN = 64, CK = 16, CPAD = 68;

// MR = output-channel rows per thread: the tile is (16*MR channels) x 64 pixels.  MR = 1 serves the layers with a handful
// of output channels (de_1 and the data gradient of en_1: Cout = 3), where a 64-row tile would be 95 % padding.
template <bool TRANSPOSED, int MR>
__global__ void __launch_bounds__(256) conv_simt_kernel(ConvArgs p) {
  constexpr int TM = 16 * MR;
  __shared__ __align__(16) float As[CK][CPAD];   // weights  [k][co]
  __shared__ __align__(16) float Bs[CK][CPAD];   // gathered input [k][pixel]

  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.y * TM, n0 = blockIdx.x * CN;
  const int kk = p.k * p.k;
  const int K = p.Cin * kk;
  const int HoWo = p.Ho * p.Wo;
  const long long Ntot = (long long)p.B * HoWo;
  const long long HW = (long long)p.H * p.W;

  // B loader: this thread always gathers for pixel (n0 + tid%64), k rows (tid/64) + 4*i
  const int bn = tid & 63, bk0 = tid >> 6;
  const long long npix = (long long)n0 + bn;
  const bool n_ok = npix < Ntot;
  int pb = 0, oy = 0, ox = 0;
  if (n_ok) {
    pb = (int)(npix / HoWo);
    const int r = (int)(npix - (long long)pb * HoWo);
    oy = r / p.Wo;
    ox = r - oy * p.Wo;
  }
  // A loader: weight row (tid/16) + 16*i, k column tid%16
  const int ak = tid & 15, am0 = tid >> 4;

  float acc[MR][4];
#pragma unroll
  for (int i = 0; i < MR; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  // split-K (deep layers: few output tiles, K = Cin*k*k up to 16384): blockIdx.z takes a K range, atomics combine
  const int kchunk = ((K + p.ksplit - 1) / p.ksplit + CK - 1) / CK * CK;
  const int kbeg = blockIdx.z * kchunk;
  const int kend = min(K, kbeg + kchunk);
  for (int k0 = kbeg; k0 < kend; k0 += CK) {
    float av[4], bv[4];
    {
      const int k = k0 + ak;
      int ci = 0, tap = 0;
      if (TRANSPOSED) { ci = k / kk; tap = k - ci * kk; }
#pragma unroll
      for (int i = 0; i < MR; ++i) {
        const int co = m0 + am0 + 16 * i;
        float v = 0.f;
        if (k < kend && co < p.Cout)
          v = TRANSPOSED ? p.w[((long long)ci * p.Cout + co) * kk + tap] : p.w[(long long)co * K + k];
        av[i] = v;
      }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int k = k0 + bk0 + 4 * i;
      float v = 0.f;
      if (n_ok && k < kend) {
        const int ci = k / kk;
        const int tap = k - ci * kk;
        const int ky = tap / p.k, kx = tap - ky * p.k;
        int iy, ix;
        bool ok;
        if (TRANSPOSED) {
          // out[oy] += in[iy] * w[ky]  with  oy = iy*stride - pad + ky*dil
          const int ty2 = oy + p.pad - ky * p.dil, tx2 = ox + p.pad - kx * p.dil;   // dil > 1: dgrad of a dilated conv
          ok = ty2 >= 0 && tx2 >= 0 && (ty2 % p.stride) == 0 && (tx2 % p.stride) == 0;
          iy = ty2 / p.stride;
          ix = tx2 / p.stride;
          ok = ok && iy < p.H && ix < p.W;
        } else {
          iy = oy * p.stride - p.pad + ky * p.dil;
          ix = ox * p.stride - p.pad + kx * p.dil;
          ok = iy >= 0 && iy < p.H && ix >= 0 && ix < p.W;
        }
        if (ok) {
          int seg = 0, cl = ci;
          if (p.n_seg > 1 && cl >= p.seg_ch[0]) {
            cl -= p.seg_ch[0]; seg = 1;
            if (p.n_seg > 2 && cl >= p.seg_ch[1]) { cl -= p.seg_ch[1]; seg = 2; }
          }
          v = p.in[seg][((long long)pb * p.seg_ch[seg] + cl) * HW + (long long)iy * p.W + ix];
          if (p.in_scale[seg] != nullptr) v = fmaf(v, p.in_scale[seg][cl], p.in_shift[seg][cl]);
          if (p.in_act == 1) v = v >= 0.f ? v : 0.2f * v;
          else if (p.in_act == 2) v = fmaxf(v, 0.f);
        }
      }
      bv[i] = v;
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      if (i < MR) As[ak][am0 + 16 * i] = av[i];
      Bs[bk0 + 4 * i][bn] = bv[i];
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < CK; ++k) {
      float ar[MR];
      if constexpr (MR == 4) {
        const float4 a4 = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
        ar[0] = a4.x; ar[1] = a4.y; ar[2] = a4.z; ar[3] = a4.w;
      } else {
#pragma unroll
        for (int i = 0; i < MR; ++i) ar[i] = As[k][ty * MR + i];
      }
      const float4 b4 = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
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
    const float bb = (p.bias && blockIdx.z == 0) ? p.bias[co] : 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const long long n = (long long)n0 + tx * 4 + j;
      if (n >= Ntot) continue;
      const int b = (int)(n / HoWo);
      const int r = (int)(n - (long long)b * HoWo);
      float* dst = p.out + ((long long)b * p.Cout + co) * HoWo + r;
      if (p.ksplit == 1) *dst = acc[i][j] + bb;
      else atomicAdd(dst, acc[i][j] + bb);       // `out` was zero-filled by the launcher
    }
  }
}

// Transposed conv with stride 2 (convT 4x4 s2 of the decoder, and the data gradients of the stride-2 encoder convs):
// in the gather form above 3 of 4 taps of every output pixel hit a non-integer input position and multiply zeros.  Here
// the output is split into its 4 parity classes (blockIdx.z); a class only enumerates the taps that are valid for it:
//   (oy + pad - ky*dil) even  <=>  ((py + pad + ky*dil) & 1) == 0          K_eff = Cin * nky * nkx  (a quarter for dil 1;
// for dil 2 one class carries all taps and the other three are bias-only), N = B * Ho/2 * Wo/2 pixels per class.
__global__ void __launch_bounds__(256) conv_simt_t2_kernel(ConvArgs p) {
  __shared__ __align__(16) float As[CK][CPAD];
  __shared__ __align__(16) float Bs[CK][CPAD];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.y * CM, n0 = blockIdx.x * CN;
  const int cls = blockIdx.z / p.ksplit, ks = blockIdx.z - cls * p.ksplit;
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
  const int kchunk = ((K + p.ksplit - 1) / p.ksplit + CK - 1) / CK * CK;
  const int kbeg = ks * kchunk;
  const int kend = min(K, kbeg + kchunk);
  const int Hq = p.Ho >> 1, Wq = p.Wo >> 1;
  const int HqWq = Hq * Wq;
  const int HoWo = p.Ho * p.Wo;
  const long long Ntot = (long long)p.B * HqWq;
  const long long HW = (long long)p.H * p.W;

  const int bn = tid & 63, bk0 = tid >> 6;
  const long long npix = (long long)n0 + bn;
  const bool n_ok = npix < Ntot;
  int pb = 0, oy = 0, ox = 0;
  if (n_ok) {
    pb = (int)(npix / HqWq);
    const int r = (int)(npix - (long long)pb * HqWq);
    const int qy = r / Wq;
    oy = 2 * qy + py;
    ox = 2 * (r - qy * Wq) + px;
  }
  const int ak = tid & 15, am0 = tid >> 4;

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int k0 = kbeg; k0 < kend; k0 += CK) {
    float av[4], bv[4];
    {
      const int k = k0 + ak;
      int wofs = 0;
      if (k < kend) {
        const int ci = k / nt, r = k - ci * nt;
        const int a = r / nkx, b2 = r - a * nkx;
        wofs = ci * p.Cout * kk + kys[a] * p.k + kxs[b2];
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int co = m0 + am0 + 16 * i;
        av[i] = (k < kend && co < p.Cout) ? p.w[(long long)wofs + (long long)co * kk] : 0.f;
      }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int k = k0 + bk0 + 4 * i;
      float v = 0.f;
      if (n_ok && k < kend) {
        const int ci = k / nt, r = k - ci * nt;
        const int a = r / nkx, b2 = r - a * nkx;
        const int ty2 = oy + p.pad - kys[a] * p.dil, tx2 = ox + p.pad - kxs[b2] * p.dil;   // even by construction
        const int iy = ty2 >> 1, ix = tx2 >> 1;
        if (ty2 >= 0 && tx2 >= 0 && iy < p.H && ix < p.W) {
          int seg = 0, cl = ci;
          if (p.n_seg > 1 && cl >= p.seg_ch[0]) {
            cl -= p.seg_ch[0]; seg = 1;
            if (p.n_seg > 2 && cl >= p.seg_ch[1]) { cl -= p.seg_ch[1]; seg = 2; }
          }
          v = p.in[seg][((long long)pb * p.seg_ch[seg] + cl) * HW + (long long)iy * p.W + ix];
          if (p.in_scale[seg] != nullptr) v = fmaf(v, p.in_scale[seg][cl], p.in_shift[seg][cl]);
          if (p.in_act == 1) v = v >= 0.f ? v : 0.2f * v;
          else if (p.in_act == 2) v = fmaxf(v, 0.f);
        }
      }
      bv[i] = v;
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      As[ak][am0 + 16 * i] = av[i];
      Bs[bk0 + 4 * i][bn] = bv[i];
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < CK; ++k) {
      const float4 a4 = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      const float4 b4 = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
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
    const float bb = (p.bias && ks == 0) ? p.bias[co] : 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const long long n = (long long)n0 + tx * 4 + j;
      if (n >= Ntot) continue;
      const int b = (int)(n / HqWq);
      const int r = (int)(n - (long long)b * HqWq);
      const int qy = r / Wq, qx = r - qy * Wq;
      float* dst = p.out + ((long long)b * p.Cout + co) * HoWo + (long long)(2 * qy + py) * p.Wo + 2 * qx + px;
      if (p.ksplit == 1) *dst = acc[i][j] + bb;
      else atomicAdd(dst, acc[i][j] + bb);
    }
  }
}

int launch_conv_simt(const ConvArgs& a_in, cudaStream_t st) {
  ConvArgs a = a_in;
  if (a.n_seg < 1 || a.n_seg > 3) return -1;
  int cin = 0;
  for (int i = 0; i < a.n_seg; ++i) cin += a.seg_ch[i];
  if (cin != a.Cin) return -1;
  const long long Ntot = (long long)a.B * a.Ho * a.Wo;
  const bool t2 = a.transposed && a.stride == 2 && a.k <= 4 && a.Ho % 2 == 0 && a.Wo % 2 == 0;
  const long long n_tiles = t2 ? (Ntot / 4 + CN - 1) / CN * 4 : (Ntot + CN - 1) / CN;
  const long long ctas = n_tiles * ((a.Cout + CM - 1) / CM);
  const int K = a.Cin * (t2 ? (a.dil % 2 == 0 ? a.k * a.k : ((a.k + 1) / 2) * ((a.k + 1) / 2)) : a.k * a.k);
  a.ksplit = 1;
  if (ctas < 296 && K >= 1024) {       // deep layers: fill the SMs by splitting K (atomics into a zero-filled output)
    long long ks = 592 / ctas;
    if (ks > K / 256) ks = K / 256;
    if (ks > 1) {
      a.ksplit = (int)ks;
      DPMN_CUDA_TRY(cudaMemsetAsync(a.out, 0, (size_t)a.B * a.Cout * a.Ho * a.Wo * sizeof(float), st));
    }
  }
  if (t2) {
    dim3 g2((unsigned)((Ntot / 4 + CN - 1) / CN), (a.Cout + CM - 1) / CM, 4 * a.ksplit);
    conv_simt_t2_kernel<<<g2, 256, 0, st>>>(a);
    DPMN_LAUNCH_CHECK();
    return 0;
  }
  if (a.Cout <= 16) {
    dim3 grid((unsigned)((Ntot + CN - 1) / CN), (a.Cout + 15) / 16, a.ksplit);
    if (a.transposed) conv_simt_kernel<true, 1><<<grid, 256, 0, st>>>(a);
    else conv_simt_kernel<false, 1><<<grid, 256, 0, st>>>(a);
    DPMN_LAUNCH_CHECK();
    return 0;
  }
  dim3 grid((unsigned)((Ntot + CN - 1) / CN), (a.Cout + CM - 1) / CM, a.ksplit);
  if (a.transposed) conv_simt_kernel<true, 4><<<grid, 256, 0, st>>>(a);
  else conv_simt_kernel<false, 4><<<grid, 256, 0, st>>>(a);
  DPMN_LAUNCH_CHECK();
  return 0;
}

// -----------------------------------------------------------------------------------------------------
// BatchNorm2d -> per-channel (scale, shift).  One CTA per channel; two-pass (mean, then centred second
// moment) so the batch variance has no cancellation.                                       cmm.py:12
// -----------------------------------------------------------------------------------------------------
__device__ __forceinline__ float block_sum_256(float v, float* red) {
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

__global__ void __launch_bounds__(256) bn_affine_kernel(const float* __restrict__ x, int B, int C, int HW,
                                                        const float* __restrict__ w, const float* __restrict__ b,
                                                        float* run_mean, float* run_var, int training,
                                                        int update_running, float eps, float* __restrict__ scale,
                                                        float* __restrict__ shift) {
  __shared__ float red[8];
  const int c = blockIdx.x;
  float mean, var;
  if (training) {
    const long long n = (long long)B * HW;
    float s = 0.f;
    for (long long i = threadIdx.x; i < n; i += 256) {
      const int bb = (int)(i / HW);
      const int r = (int)(i - (long long)bb * HW);
      s += x[((long long)bb * C + c) * HW + r];
    }
    mean = block_sum_256(s, red) / (float)n;
    float q = 0.f;
    for (long long i = threadIdx.x; i < n; i += 256) {
      const int bb = (int)(i / HW);
      const int r = (int)(i - (long long)bb * HW);
      const float d = x[((long long)bb * C + c) * HW + r] - mean;
      q = fmaf(d, d, q);
    }
    const float ss = block_sum_256(q, red);
    var = ss / (float)n;
    if (update_running && run_mean != nullptr && threadIdx.x == 0) {
      const float unbiased = n > 1 ? ss / (float)(n - 1) : var;
      run_mean[c] = 0.9f * run_mean[c] + 0.1f * mean;
      run_var[c] = 0.9f * run_var[c] + 0.1f * unbiased;
    }
  } else {
    mean = run_mean[c];
    var = run_var[c];
  }
  if (threadIdx.x == 0) {
    const float sc = w[c] / sqrtf(var + eps);
    scale[c] = sc;
    shift[c] = b[c] - mean * sc;
  }
}

// Batch statistics over many CTAs: grid (C, S) accumulates (sum, sum of squares) of a slice in fp64 (one pass; the fp64
// difference E[x^2] - E[x]^2 keeps the two-pass kernel's accuracy), then one thread per channel finishes.  The one-block-per-
// channel kernel above runs 64 blocks for the 64-channel layers (49 K - 196 K values each, two passes): 80 us per launch.
__global__ void __launch_bounds__(256) bn_stats_slices_kernel(const float* __restrict__ x, int B, int C, int HW, int hw_shift,
                                                              double* __restrict__ part) {
  __shared__ double red[2][8];
  const int c = blockIdx.x, S = gridDim.y;
  const long long n = (long long)B * HW;
  const long long per = (n + S - 1) / S, i0 = blockIdx.y * per, i1 = i0 + per < n ? i0 + per : n;
  double s = 0.0, q = 0.0;
  for (long long i = i0 + threadIdx.x; i < i1; i += 256) {
    const int bb = hw_shift >= 0 ? (int)(i >> hw_shift) : (int)(i / HW);
    const int r = (int)(i - (long long)bb * HW);
    const double v = (double)x[((long long)bb * C + c) * HW + r];
    s += v;
    q = fma(v, v, q);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { s += __shfl_xor_sync(0xffffffffu, s, o); q += __shfl_xor_sync(0xffffffffu, q, o); }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) { red[0][warp] = s; red[1][warp] = q; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double ts = 0.0, tq = 0.0;
#pragma unroll
    for (int k = 0; k < 8; ++k) { ts += red[0][k]; tq += red[1][k]; }
    part[((long long)c * S + blockIdx.y) * 2] = ts;
    part[((long long)c * S + blockIdx.y) * 2 + 1] = tq;
  }
}

__global__ void bn_affine_finish_kernel(const double* __restrict__ part, int S, long long n, int C, const float* __restrict__ w,
                                        const float* __restrict__ b, float* run_mean, float* run_var, int update_running,
                                        float eps, float* __restrict__ scale, float* __restrict__ shift) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  double s = 0.0, q = 0.0;
  for (int k = 0; k < S; ++k) { s += part[((long long)c * S + k) * 2]; q += part[((long long)c * S + k) * 2 + 1]; }
  const double mean_d = s / (double)n;
  double ss = q - s * mean_d;                            // sum of squared deviations
  if (ss < 0.0) ss = 0.0;
  const float mean = (float)mean_d, var = (float)(ss / (double)n);
  if (update_running && run_mean != nullptr) {
    const float unbiased = n > 1 ? (float)(ss / (double)(n - 1)) : var;
    run_mean[c] = 0.9f * run_mean[c] + 0.1f * mean;
    run_var[c] = 0.9f * run_var[c] + 0.1f * unbiased;
  }
  const float sc = w[c] / sqrtf(var + eps);
  scale[c] = sc;
  shift[c] = b[c] - mean * sc;
}

// stats_scratch: BN_STATS_SCRATCH_DOUBLES(C) doubles for the sliced training-statistics path, or nullptr (one block per channel)
int launch_bn_affine(const float* x, int B, int C, int HW, const float* w, const float* b, float* run_mean,
                     float* run_var, int training, int update_running, float eps, float* scale, float* shift,
                     cudaStream_t st, double* stats_scratch) {
  if (training && stats_scratch != nullptr) {
    const long long n = (long long)B * HW;
    int S = (148 * 4 + C - 1) / C;
    if (S > BN_STATS_MAX_SLICES) S = BN_STATS_MAX_SLICES;
    const long long cap = (n + 1023) / 1024;             // at least 4 values per thread
    if (S > cap) S = (int)cap;
    if (S < 1) S = 1;
    int sh = -1;
    for (int k = 0; k < 31; ++k) if ((1 << k) == HW) sh = k;
    bn_stats_slices_kernel<<<dim3(C, S), 256, 0, st>>>(x, B, C, HW, sh, stats_scratch);
    DPMN_LAUNCH_CHECK();
    bn_affine_finish_kernel<<<(C + 127) / 128, 128, 0, st>>>(stats_scratch, S, n, C, w, b, run_mean, run_var, update_running, eps,
                                                           scale, shift);
    DPMN_LAUNCH_CHECK();
    return 0;
  }
  bn_affine_kernel<<<C, 256, 0, st>>>(x, B, C, HW, w, b, run_mean, run_var, training, update_running, eps, scale,
                                      shift);
  DPMN_LAUNCH_CHECK();
  return 0;
}

// -----------------------------------------------------------------------------------------------------
// SE gate on the bottleneck                                                           cmm.py:135-147
// z = cat(z1, z2) (2*Cb channels);  g = sigmoid(fc2(relu(fc1(mean_hw z))));  out = z * g + z
// One CTA per image.
// -----------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) se_gate_kernel(const float* __restrict__ z1, const float* __restrict__ z2,
                                                      float* __restrict__ z, const float* __restrict__ fc1_w,
                                                      const float* __restrict__ fc1_b,
                                                      const float* __restrict__ fc2_w,
                                                      const float* __restrict__ fc2_b, int Cb, int hw, int hidden) {
  extern __shared__ float sm[];
  const int C2 = 2 * Cb;
  float* sg = sm;            // [C2] pooled
  float* sh = sg + C2;       // [hidden]
  const int b = blockIdx.x;
  for (int c = threadIdx.x; c < C2; c += blockDim.x) {
    const float* src = c < Cb ? z1 + ((long long)b * Cb + c) * hw : z2 + ((long long)b * Cb + (c - Cb)) * hw;
    float s = 0.f;
    for (int i = 0; i < hw; ++i) s += src[i];
    sg[c] = s / (float)hw;
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int j = warp; j < hidden; j += nw) {
    float s = 0.f;
    for (int c = lane; c < C2; c += 32) s = fmaf(fc1_w[(long long)j * C2 + c], sg[c], s);
    s = warp_sum(s);
    if (lane == 0) sh[j] = fmaxf(s + fc1_b[j], 0.f);
  }
  __syncthreads();
  for (int c = warp; c < C2; c += nw) {
    float s = 0.f;
    for (int j = lane; j < hidden; j += 32) s = fmaf(fc2_w[(long long)c * hidden + j], sh[j], s);
    s = warp_sum(s);
    const float g = 1.0f / (1.0f + expf(-(s + fc2_b[c])));
    const float* src = c < Cb ? z1 + ((long long)b * Cb + c) * hw : z2 + ((long long)b * Cb + (c - Cb)) * hw;
    for (int i = lane; i < hw; i += 32) {
      const float v = src[i];
      z[((long long)b * C2 + c) * hw + i] = fmaf(v, g, v);
    }
  }
}

// The same gate over many CTAs (one CTA per image runs 48 CTAs for 2 x 256 K multiply-adds each: 319 us at batch 48):
// (1) grid (B, hidden / 8): every CTA pools its image (C2 x hw loads) and its 8 warps take one fc1 row each -> h (B, hidden)
// (2) grid (B, C2 / 64):    h row in shared memory, a warp per channel: fc2 row, sigmoid, z * g + z
__global__ void __launch_bounds__(256) se_pool_fc1_kernel(const float* __restrict__ z1, const float* __restrict__ z2,
                                                          const float* __restrict__ fc1_w, const float* __restrict__ fc1_b,
                                                          int Cb, int hw, int hidden, float* __restrict__ h_out, int h_stride,
                                                          float* __restrict__ g0_out, int g0_stride) {
  extern __shared__ float sm[];
  const int C2 = 2 * Cb, b = blockIdx.x;
  for (int c = threadIdx.x; c < C2; c += blockDim.x) {
    const float* src = c < Cb ? z1 + ((long long)b * Cb + c) * hw : z2 + ((long long)b * Cb + (c - Cb)) * hw;
    float s = 0.f;
    for (int i = 0; i < hw; ++i) s += src[i];
    sm[c] = s / (float)hw;
    if (g0_out != nullptr && blockIdx.y == 0) g0_out[(long long)b * g0_stride + c] = sm[c];
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int j = blockIdx.y * 8 + warp;
  if (j >= hidden) return;
  float s = 0.f;
  for (int c = lane; c < C2; c += 32) s = fmaf(fc1_w[(long long)j * C2 + c], sm[c], s);
  s = warp_sum(s);
  if (lane == 0) h_out[(long long)b * h_stride + j] = fmaxf(s + fc1_b[j], 0.f);
}

__global__ void __launch_bounds__(256) se_fc2_gate_fwd_kernel(const float* __restrict__ z1, const float* __restrict__ z2,
                                                              float* __restrict__ z, const float* __restrict__ fc2_w,
                                                              const float* __restrict__ fc2_b, const float* __restrict__ h,
                                                              int h_stride, int Cb, int hw, int hidden) {
  extern __shared__ float sm[];
  const int C2 = 2 * Cb, b = blockIdx.x;
  for (int j = threadIdx.x; j < hidden; j += blockDim.x) sm[j] = h[(long long)b * h_stride + j];
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int cc = warp; cc < 64; cc += 8) {
    const int c = blockIdx.y * 64 + cc;
    if (c >= C2) break;
    float s = 0.f;
    for (int j = lane; j < hidden; j += 32) s = fmaf(fc2_w[(long long)c * hidden + j], sm[j], s);
    s = warp_sum(s);
    const float g = 1.0f / (1.0f + expf(-(s + fc2_b[c])));
    const float* src = c < Cb ? z1 + ((long long)b * Cb + c) * hw : z2 + ((long long)b * Cb + (c - Cb)) * hw;
    for (int i = lane; i < hw; i += 32) {
      const float v = src[i];
      z[((long long)b * C2 + c) * hw + i] = fmaf(v, g, v);
    }
  }
}

int launch_se_pool_fc1(const float* z1, const float* z2, const float* fc1_w, const float* fc1_b, int B, int Cb, int hw,
                       int hidden, float* h_out, int h_stride, float* g0_out, int g0_stride, cudaStream_t st) {
  if ((size_t)2 * Cb * sizeof(float) > 48 * 1024) return -2;
  se_pool_fc1_kernel<<<dim3(B, (hidden + 7) / 8), 256, (size_t)2 * Cb * sizeof(float), st>>>(z1, z2, fc1_w, fc1_b, Cb, hw, hidden,
                                                                                          h_out, h_stride, g0_out, g0_stride);
  DPMN_LAUNCH_CHECK();
  return 0;
}

// h_scratch: (B, hidden) floats, or nullptr for the one-CTA-per-image kernel.
int launch_se_gate(const float* z1, const float* z2, float* z, const float* fc1_w, const float* fc1_b,
                   const float* fc2_w, const float* fc2_b, int B, int Cb, int hw, int hidden, cudaStream_t st,
                   float* h_scratch) {
  const size_t smem = (size_t)(2 * Cb + hidden) * sizeof(float);
  if (smem > 48 * 1024) return -2;
  if (h_scratch != nullptr) {
    const int rc = launch_se_pool_fc1(z1, z2, fc1_w, fc1_b, B, Cb, hw, hidden, h_scratch, hidden, nullptr, 0, st);
    if (rc) return rc;
    se_fc2_gate_fwd_kernel<<<dim3(B, (2 * Cb + 63) / 64), 256, (size_t)hidden * sizeof(float), st>>>(z1, z2, z, fc2_w, fc2_b,
                                                                                                   h_scratch, hidden, Cb, hw, hidden);
    DPMN_LAUNCH_CHECK();
    return 0;
  }
  se_gate_kernel<<<B, 256, smem, st>>>(z1, z2, z, fc1_w, fc1_b, fc2_w, fc2_b, Cb, hw, hidden);
  DPMN_LAUNCH_CHECK();
  return 0;
}

}  // namespace dpmn
