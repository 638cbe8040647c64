// CMM convolutions of the TRAINING path on the tcgen05 GEMM (16-bit modes)                    cmm.py:38-161
//
// The training forward / backward of the CMM keeps fp32 NCHW tensors (raw conv outputs + BatchNorm affines, see
// api.cu cmm_forward_struct and api_bwd.inc), so its convs cannot use conv_tc.cu's NHWC TMA boxes.  Instead each conv
// -- forward, data gradient (= the opposite kind of conv on the same weights) and weight gradient -- is lowered to the
// NT GEMM of gemm_tc.cu through a 16-bit im2col matrix that one gather kernel writes:
//     forward / dgrad   out_b (Cout, HoWo) = W16 (Cout, Kp) * col_b (HoWo, Kp)^T           per image, fp32 NCHW output
//     wgrad             dW (Cout, Kp)     += dy16 (Cout, Ntot) * colT (Kp, Ntot)^T         split over pixel chunks
// The gather is the B-operand loader of conv_simt_kernel verbatim: channel concat, the producer's BatchNorm affine
// and the consumer's activation are applied on the way, transposed / strided / dilated geometry included (stride-2
// transposed convs in plain gather form: the structural zeros cost tensor-core FLOPs, which are not the bottleneck).
// K = Cin*k*k is padded to Kp (multiple of 16) with zeros on both operands.
#include "common.cuh"
#include "kernels.h"
#include <cstdlib>

namespace dpmn {

// KS = kernel size (3 or 4) as a template parameter: k -> (ci, ky, kx) becomes multiply-shift; strides are 1 or 2.
template <int KS>
__device__ __forceinline__ float conv_gather(const ConvArgs& p, int k, int K, int b, int oy, int ox, long long HW) {
  constexpr int kk = KS * KS;
  if (k >= K) return 0.f;
  const int ci = k / kk;
  const int tap = k - ci * kk;
  const int ky = tap / KS, kx = tap - ky * KS;
  const int sh = p.stride >> 1;                       // stride 1 -> 0, stride 2 -> 1
  int iy, ix;
  bool ok;
  if (p.transposed) {
    const int ty2 = oy + p.pad - ky * p.dil, tx2 = ox + p.pad - kx * p.dil;
    ok = ty2 >= 0 && tx2 >= 0 && ((ty2 | tx2) & sh) == 0;
    iy = ty2 >> sh; ix = tx2 >> sh;
    ok = ok && iy < p.H && ix < p.W;
  } else {
    iy = (oy << sh) - p.pad + ky * p.dil;
    ix = (ox << sh) - p.pad + kx * p.dil;
    ok = iy >= 0 && iy < p.H && ix >= 0 && ix < p.W;
  }
  if (!ok) return 0.f;
  int seg = 0, cl = ci;
  if (p.n_seg > 1 && cl >= p.seg_ch[0]) {
    cl -= p.seg_ch[0]; seg = 1;
    if (p.n_seg > 2 && cl >= p.seg_ch[1]) { cl -= p.seg_ch[1]; seg = 2; }
  }
  float v = p.in[seg][((long long)b * p.seg_ch[seg] + cl) * HW + (long long)iy * p.W + ix];
  if (p.in_scale[seg] != nullptr) v = fmaf(v, p.in_scale[seg][cl], p.in_shift[seg][cl]);
  if (p.in_act == 1) v = v >= 0.f ? v : 0.2f * v;
  else if (p.in_act == 2) v = fmaxf(v, 0.f);
  return v;
}

// COLT = false: col[n][k] (row length Kp); 32 (pixels) x 32 (k) tiles are gathered with threads along pixels (coalesced
// reads of the NCHW input) and written with threads along k (coalesced 16-bit rows).
// COLT = true:  colT[k][n] (row length Npad): threads along pixels for both.
// A CTA walks IM_KT consecutive k-tiles for its 32 pixels, so the pixel -> (image, y, x) decode is paid once.
constexpr int IM_KT = 8;

template <typename T, bool COLT, int KS>
__global__ void __launch_bounds__(256) im2col_kernel(ConvArgs p, T* __restrict__ col, int Kp, long long Npad) {
  __shared__ float tile[32][33];
  const int K = p.Cin * KS * KS;
  const int HoWo = p.Ho * p.Wo;
  const long long Ntot = (long long)p.B * HoWo;
  const long long HW = (long long)p.H * p.W;
  const long long n0 = (long long)blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const long long n = n0 + tx;
  int b = 0, oy = 0, ox = 0;
  const bool n_ok = n < Ntot;
  if (n_ok) {
    b = (int)(n / HoWo);
    const int r = (int)(n - (long long)b * HoWo);
    oy = r / p.Wo; ox = r - oy * p.Wo;
  }
  for (int kt = 0; kt < IM_KT; ++kt) {
    const int k0 = (blockIdx.y * IM_KT + kt) * 32;
    if (k0 >= Kp) break;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int k = k0 + ty + 8 * i;
      const float v = (n_ok && k < Kp) ? conv_gather<KS>(p, k, K, b, oy, ox, HW) : 0.f;
      if (COLT) {
        if (k < Kp && n < Npad) col[(long long)k * Npad + n] = from_f32<T>(v);
      } else {
        tile[ty + 8 * i][tx] = v;
      }
    }
    if (!COLT) {
      __syncthreads();
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const long long nn = n0 + ty + 8 * i;
        const int k = k0 + tx;
        if (nn < Ntot && k < Kp) col[nn * Kp + k] = from_f32<T>(tile[tx][ty + 8 * i]);
      }
      __syncthreads();
    }
  }
}

template <typename T, bool COLT>
static int launch_im2col(const ConvArgs& a, T* col, int Kp, long long Ntot, cudaStream_t st) {
  const dim3 grid((unsigned)((Ntot + 31) / 32), (Kp + 32 * IM_KT - 1) / (32 * IM_KT));
  if (a.k == 3) im2col_kernel<T, COLT, 3><<<grid, 256, 0, st>>>(a, col, Kp, Ntot);
  else if (a.k == 4) im2col_kernel<T, COLT, 4><<<grid, 256, 0, st>>>(a, col, Kp, Ntot);
  else return -2;
  DPMN_LAUNCH_CHECK();
  return 0;
}

// ---- fast gather path (Cin % 8 == 0): activate once, then copy ------------------------------------------------------------
// The element-wise gather above pays ~60 instructions per im2col element (index decode, segment lookup, affine, activation)
// and every input element is gathered k*k times.  Here the conv input -- channel concat, producer BatchNorm affine and
// consumer activation applied -- is written ONCE as NHWC 16-bit (act_nhwc_kernel, a tiled transpose of the fp32 NCHW
// tensors), and the im2col matrix, with K ordered (tap, ci), becomes a copy of whole channel vectors: 16-byte loads and
// stores, one index decode per (pixel, tap).  Weights are staged in the same (tap, ci) order.

// xa[b][y][x][c] (c over the concatenated channels) = act(affine(in)), 16-bit
template <typename T>
__global__ void __launch_bounds__(256) act_nhwc_kernel(ConvArgs p, T* __restrict__ xa) {
  __shared__ float tile[32][33];
  const long long HW = (long long)p.H * p.W;
  const int b = blockIdx.z;
  const long long hw0 = (long long)blockIdx.x * 32;
  const int c0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int c = c0 + ty + 8 * i;
    const long long hw = hw0 + tx;
    float v = 0.f;
    if (c < p.Cin && hw < HW) {
      int seg = 0, cl = c;
      if (p.n_seg > 1 && cl >= p.seg_ch[0]) {
        cl -= p.seg_ch[0]; seg = 1;
        if (p.n_seg > 2 && cl >= p.seg_ch[1]) { cl -= p.seg_ch[1]; seg = 2; }
      }
      v = p.in[seg][((long long)b * p.seg_ch[seg] + cl) * HW + hw];
      if (p.in_scale[seg] != nullptr) v = fmaf(v, p.in_scale[seg][cl], p.in_shift[seg][cl]);
      if (p.in_act == 1) v = v >= 0.f ? v : 0.2f * v;
      else if (p.in_act == 2) v = fmaxf(v, 0.f);
    }
    tile[ty + 8 * i][tx] = v;
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const long long hw = hw0 + ty + 8 * i;
    const int c = c0 + tx;
    if (hw < HW && c < p.Cin) xa[((long long)b * HW + hw) * p.Cin + c] = from_f32<T>(tile[tx][ty + 8 * i]);
  }
}

// input pixel of output pixel (oy, ox) under tap (ky, kx); false = structural zero / padding
__device__ __forceinline__ bool tap_source(const ConvArgs& p, int oy, int ox, int ky, int kx, int& iy, int& ix) {
  const int sh = p.stride >> 1;
  if (p.transposed) {
    const int ty2 = oy + p.pad - ky * p.dil, tx2 = ox + p.pad - kx * p.dil;
    if (ty2 < 0 || tx2 < 0 || ((ty2 | tx2) & sh) != 0) return false;
    iy = ty2 >> sh; ix = tx2 >> sh;
    return iy < p.H && ix < p.W;
  }
  iy = (oy << sh) - p.pad + ky * p.dil;
  ix = (ox << sh) - p.pad + kx * p.dil;
  return iy >= 0 && iy < p.H && ix >= 0 && ix < p.W;
}

// col[n][tap * Cin + ci]: one warp per output pixel at a time, lanes over (tap, 8-channel vector)
template <typename T, int KS>
__global__ void __launch_bounds__(256) im2col_nhwc_kernel(ConvArgs p, const T* __restrict__ xa, T* __restrict__ col, int Kp,
                                                          long long Ntot) {
  constexpr int kk = KS * KS;
  const int HoWo = p.Ho * p.Wo;
  const int vec = p.Cin >> 3;                         // 16-byte vectors per channel row
  const int items = kk * vec;
  const int lane = threadIdx.x & 31;
  const long long warp0 = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  const long long nwarps = (long long)gridDim.x * 8;
  const uint4 zero = make_uint4(0u, 0u, 0u, 0u);
  for (long long n = warp0; n < Ntot; n += nwarps) {
    const int b = (int)(n / HoWo);
    const int r = (int)(n - (long long)b * HoWo);
    const int oy = r / p.Wo, ox = r - oy * p.Wo;
    uint4* dst = reinterpret_cast<uint4*>(col + n * Kp);
    for (int it = lane; it < items; it += 32) {
      const int tap = it / vec, v = it - tap * vec;
      const int ky = tap / KS, kx = tap - ky * KS;
      int iy, ix;
      uint4 val = zero;
      if (tap_source(p, oy, ox, ky, kx, iy, ix))
        val = __ldg(reinterpret_cast<const uint4*>(xa + (((long long)b * p.H + iy) * p.W + ix) * p.Cin) + v);
      dst[it] = val;
    }
    // K is padded to a multiple of 16 elements = 2 vectors (only when kk * Cin is an odd multiple of 8)
    if (lane == 0 && items * 8 < Kp) dst[items] = zero;
  }
}

// colT[tap * Cin + ci][n] (row length Npad): 64 pixels x 64 channels per CTA per tap through a shared-memory transpose
template <typename T, int KS>
__global__ void __launch_bounds__(256) im2colT_nhwc_kernel(ConvArgs p, const T* __restrict__ xa, T* __restrict__ colT,
                                                           long long Npad, long long Ntot) {
  __shared__ uint32_t tile[64][33];                   // [pixel][channel pair]
  const int HoWo = p.Ho * p.Wo;
  const long long n0 = (long long)blockIdx.x * 64;
  const int c0 = blockIdx.y * 64;
  const int tap = blockIdx.z;
  const int ky = tap / KS, kx = tap - ky * KS;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // phase 1: warp w reads pixels n0 + w, w + 8, ...: 64 channels = 32 words per pixel (one 128-byte line)
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int pl = warp + 8 * i;
    const long long n = n0 + pl;
    uint32_t v = 0u;
    if (n < Ntot && c0 + 2 * lane < p.Cin) {
      const int b = (int)(n / HoWo);
      const int r = (int)(n - (long long)b * HoWo);
      const int oy = r / p.Wo, ox = r - oy * p.Wo;
      int iy, ix;
      if (tap_source(p, oy, ox, ky, kx, iy, ix))
        v = __ldg(reinterpret_cast<const uint32_t*>(xa + (((long long)b * p.H + iy) * p.W + ix) * p.Cin + c0) + lane);
    }
    tile[pl][lane] = v;
  }
  __syncthreads();
  // phase 2: warp w writes channels c0 + 8 w .. + 7; lane = pixel pair -> 4-byte stores, 128 contiguous bytes per row
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int cl = 8 * warp + j;                      // channel within the tile
    if (c0 + cl >= p.Cin) break;
    const uint32_t a = tile[2 * lane][cl >> 1], bq = tile[2 * lane + 1][cl >> 1];
    const uint32_t lo = (cl & 1) ? (a >> 16) : (a & 0xffffu);
    const uint32_t hi = (cl & 1) ? (bq >> 16) : (bq & 0xffffu);
    const long long n = n0 + 2 * lane;
    if (n < Npad) *reinterpret_cast<uint32_t*>(colT + ((long long)tap * p.Cin + c0 + cl) * Npad + n) = lo | (hi << 16);
  }
}

static bool nhwc_path_ok(const ConvArgs& a, const ConvTcScratch& s) {
  static const bool on = !(getenv("DPMN_IM2COL_NHWC") && atoi(getenv("DPMN_IM2COL_NHWC")) == 0);
  return on && a.Cin % 8 == 0 && s.act != nullptr && (size_t)a.B * a.H * a.W * a.Cin * 2 <= s.act_bytes;
}

template <typename T>
static int launch_act_nhwc(const ConvArgs& a, T* xa, cudaStream_t st) {
  const long long HW = (long long)a.H * a.W;
  const dim3 grid((unsigned)((HW + 31) / 32), (a.Cin + 31) / 32, a.B);
  act_nhwc_kernel<T><<<grid, 256, 0, st>>>(a, xa);
  DPMN_LAUNCH_CHECK();
  return 0;
}

// weights -> 16-bit (Cout, Kp) rows, k = ci*kk + tap (tap_major == 0) or tap*Cin + ci (tap_major == 1), zero padded
template <typename T>
__global__ void stage_conv_weight_kernel(const float* __restrict__ w, T* __restrict__ dst, int Cout, int Cin, int kk,
                                         int transposed, int Kp, int tap_major) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)Cout * Kp) return;
  const int co = (int)(i / Kp), k = (int)(i - (long long)co * Kp);
  float v = 0.f;
  if (k < Cin * kk) {
    const int ci = tap_major ? k % Cin : k / kk, tap = tap_major ? k / Cin : k - (k / kk) * kk;
    v = transposed ? w[((long long)ci * Cout + co) * kk + tap] : w[((long long)co * Cin + ci) * kk + tap];
  }
  dst[i] = from_f32<T>(v);
}

// dy (B, C, HW) fp32 -> (C, Npad) 16-bit, n = b*HW + r
template <typename T>
__global__ void nchw_to_cn_kernel(const float* __restrict__ src, T* __restrict__ dst, int B, int C, int HW, long long Npad) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)B * C * HW;
  if (i >= total) return;
  const int r = (int)(i % HW);
  const int c = (int)((i / HW) % C);
  const int b = (int)(i / ((long long)HW * C));
  dst[(long long)c * Npad + (long long)b * HW + r] = from_f32<T>(src[i]);
}

// dw (reference layout) += sum_s partial[s][co][k]
__global__ void reduce_conv_partials_kernel(const float* __restrict__ partial, float* __restrict__ dw, int S, int Cout,
                                            int Cin, int kk, int transposed, int Kp, int tap_major) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int K = Cin * kk;
  if (i >= (long long)Cout * K) return;
  const int co = (int)(i / K), k = (int)(i - (long long)co * K);
  float a = 0.f;
  for (int s = 0; s < S; ++s) a += partial[((long long)s * Cout + co) * Kp + k];
  const int ci = tap_major ? k % Cin : k / kk, tap = tap_major ? k / Cin : k - (k / kk) * kk;
  const long long o = transposed ? ((long long)ci * Cout + co) * kk + tap : ((long long)co * Cin + ci) * kk + tap;
  dw[o] += a;
}

// The same for the tap-major k order (k = tap * Cin + ci) through a shared-memory tile of 8 output channels x 32 input channels
// x all taps: the partials are read along ci (coalesced), dw is written in runs of 32 * kk (reference layout (co, ci, tap))
// or 8 * kk (transposed-conv layout (ci, co, tap)) consecutive floats.  The flat kernel above writes consecutive threads kk
// floats apart: every 4-byte read-modify-write of dw costs two 32-byte sectors (1.6 ms per training step over the CMM's
// 53.6 M weights).
__global__ void __launch_bounds__(256) reduce_conv_partials_tiled_kernel(const float* __restrict__ partial, float* __restrict__ dw,
                                                                         int S, int Cout, int Cin, int kk, int transposed, int Kp) {
  __shared__ float tile[8][16][33];
  const int ci0 = blockIdx.x * 32, co0 = blockIdx.y * 8;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int ci = ci0 + lane;
  for (int row = warp; row < 8 * kk; row += 8) {
    const int co_l = row / kk, tap = row - co_l * kk;
    const int co = co0 + co_l;
    float a = 0.f;
    if (co < Cout && ci < Cin) {
      const float* src = partial + (long long)co * Kp + (long long)tap * Cin + ci;
      const long long sstride = (long long)Cout * Kp;
      for (int sidx = 0; sidx < S; ++sidx) a += src[sidx * sstride];
    }
    tile[co_l][tap][lane] = a;
  }
  __syncthreads();
  if (!transposed) {
    const int n_ci = min(32, Cin - ci0);
    for (int co_l = 0; co_l < 8; ++co_l) {
      const int co = co0 + co_l;
      if (co >= Cout) break;
      float* dst = dw + ((long long)co * Cin + ci0) * kk;
      for (int j = threadIdx.x; j < n_ci * kk; j += 256) {
        const int ci_l = j / kk, tap = j - ci_l * kk;
        dst[j] += tile[co_l][tap][ci_l];
      }
    }
  } else {
    const int n_co = min(8, Cout - co0);
    for (int ci_l = warp; ci_l < 32; ci_l += 8) {
      if (ci0 + ci_l >= Cin) break;
      float* dst = dw + ((long long)(ci0 + ci_l) * Cout + co0) * kk;
      for (int j = lane; j < n_co * kk; j += 32) {
        const int co_l = j / kk, tap = j - co_l * kk;
        dst[j] += tile[co_l][tap][ci_l];
      }
    }
  }
}

template <typename T>
static int conv_tc_im2col_t(const ConvArgs& a, const ConvTcScratch& s, cudaStream_t st) {
  const int kk = a.k * a.k, K = a.Cin * kk, Kp = (K + 15) / 16 * 16;
  const int HoWo = a.Ho * a.Wo;
  const long long Ntot = (long long)a.B * HoWo;
  if ((size_t)Ntot * Kp * 2 > s.col_bytes || (size_t)a.Cout * Kp * 2 > s.w16_bytes) return -3;
  const bool fast = nhwc_path_ok(a, s);
  stage_conv_weight_kernel<T><<<(unsigned)(((long long)a.Cout * Kp + 255) / 256), 256, 0, st>>>(a.w, (T*)s.w16, a.Cout, a.Cin, kk,
                                                                                            a.transposed, Kp, fast ? 1 : 0);
  DPMN_LAUNCH_CHECK();
  if (fast) {
    int rc = launch_act_nhwc<T>(a, (T*)s.act, st);
    if (rc) return rc;
    const unsigned blocks = (unsigned)((Ntot + 7) / 8 < 148 * 16 ? (Ntot + 7) / 8 : 148 * 16);
    if (a.k == 3) im2col_nhwc_kernel<T, 3><<<blocks, 256, 0, st>>>(a, (const T*)s.act, (T*)s.col, Kp, Ntot);
    else im2col_nhwc_kernel<T, 4><<<blocks, 256, 0, st>>>(a, (const T*)s.act, (T*)s.col, Kp, Ntot);
    DPMN_LAUNCH_CHECK();
  } else {
    const int rc = launch_im2col<T, false>(a, (T*)s.col, Kp, Ntot, st);
    if (rc) return rc;
  }
  GemmTcArgs g;
  g.A = s.w16; g.a_bs = 0; g.lda = Kp; g.Bm = s.col; g.b_bs = (long long)HoWo * Kp; g.ldb = Kp; g.op_type = s.t;
  g.C = a.out; g.c_bs = (long long)a.Cout * HoWo; g.ldc = HoWo; g.out_type = DT_F32;
  g.M = a.Cout; g.N = HoWo; g.K = Kp; g.batch = a.B; g.bias = a.bias; g.bias_mode = a.bias ? 2 : 0;
  return launch_gemm_tc(g, st);
}

bool conv_tc_im2col_ok(const ConvArgs& a) {
  const int HoWo = a.Ho * a.Wo;
  return HoWo % 8 == 0 && HoWo >= 16 && (a.k == 3 || a.k == 4) && (a.stride == 1 || a.stride == 2);
}

int launch_conv_tc_im2col(const ConvArgs& a, const ConvTcScratch& s, cudaStream_t st) {
  if (s.t == DT_F16) return conv_tc_im2col_t<__half>(a, s, st);
  if (s.t == DT_BF16) return conv_tc_im2col_t<__nv_bfloat16>(a, s, st);
  return -1;
}

// pixel-chunk split of the weight gradient: S chunks of `chunk` pixels (multiple of 16), S * chunk == Ntot
static bool wgrad_split(long long Ntot, long long cout_kp, size_t part_bytes, int& S, long long& chunk) {
  if (Ntot % 16) return false;
  long long s = Ntot / 512;
  if (s > 64) s = 64;
  const long long cap = (long long)(part_bytes / 4) / cout_kp;
  if (s > cap) s = cap;
  if (s < 1) s = 1;
  while (s > 1 && (Ntot % (s * 16)) != 0) --s;
  if ((size_t)s * cout_kp * 4 > part_bytes) return false;
  S = (int)s; chunk = Ntot / s;
  return true;
}

bool conv_wgrad_tc_im2col_ok(const ConvArgs& a, const ConvTcScratch& s) {
  const int kk = a.k * a.k, Kp = (a.Cin * kk + 15) / 16 * 16;
  const long long Ntot = (long long)a.B * a.Ho * a.Wo;
  int S; long long chunk;
  return (a.k == 3 || a.k == 4) && (a.stride == 1 || a.stride == 2) && wgrad_split(Ntot, (long long)a.Cout * Kp, s.part_bytes, S, chunk) && (size_t)Ntot * Kp * 2 <= s.col_bytes &&
         (size_t)Ntot * a.Cout * 2 <= s.dy16_bytes;
}

template <typename T>
static int conv_wgrad_tc_im2col_t(const ConvArgs& a, const float* dy, float* dw, const ConvTcScratch& s, cudaStream_t st) {
  const int kk = a.k * a.k, K = a.Cin * kk, Kp = (K + 15) / 16 * 16;
  const int HoWo = a.Ho * a.Wo;
  const long long Ntot = (long long)a.B * HoWo;
  int S; long long chunk;
  if (!wgrad_split(Ntot, (long long)a.Cout * Kp, s.part_bytes, S, chunk)) return -2;
  const long long total = (long long)a.B * a.Cout * HoWo;
  nchw_to_cn_kernel<T><<<(unsigned)((total + 255) / 256), 256, 0, st>>>(dy, (T*)s.dy16, a.B, a.Cout, HoWo, Ntot);
  DPMN_LAUNCH_CHECK();
  const bool fast = nhwc_path_ok(a, s) && Ntot % 2 == 0;
  if (fast) {
    int rc = launch_act_nhwc<T>(a, (T*)s.act, st);
    if (rc) return rc;
    if (Kp > K) DPMN_CUDA_TRY(cudaMemsetAsync((T*)s.col + (long long)K * Ntot, 0, (size_t)(Kp - K) * Ntot * 2, st));
    const dim3 grid((unsigned)((Ntot + 63) / 64), (a.Cin + 63) / 64, kk);
    if (a.k == 3) im2colT_nhwc_kernel<T, 3><<<grid, 256, 0, st>>>(a, (const T*)s.act, (T*)s.col, Ntot, Ntot);
    else im2colT_nhwc_kernel<T, 4><<<grid, 256, 0, st>>>(a, (const T*)s.act, (T*)s.col, Ntot, Ntot);
    DPMN_LAUNCH_CHECK();
  } else {
    const int rc = launch_im2col<T, true>(a, (T*)s.col, Kp, Ntot, st);
    if (rc) return rc;
  }
  GemmTcArgs g;       // partial[s] (Cout, Kp) = dy16[:, chunk s] * colT[:, chunk s]^T
  g.A = s.dy16; g.a_bs = chunk; g.lda = (int)Ntot; g.Bm = s.col; g.b_bs = chunk; g.ldb = (int)Ntot; g.op_type = s.t;
  g.C = s.part; g.c_bs = (long long)a.Cout * Kp; g.ldc = Kp; g.out_type = DT_F32;
  g.M = a.Cout; g.N = Kp; g.K = (int)chunk; g.batch = S;
  int rc = launch_gemm_tc(g, st);
  if (rc) return rc;
  static const bool tiled = !(getenv("DPMN_REDUCE_TILED") && atoi(getenv("DPMN_REDUCE_TILED")) == 0);
  if (tiled && fast && kk <= 16)
    reduce_conv_partials_tiled_kernel<<<dim3((a.Cin + 31) / 32, (a.Cout + 7) / 8), 256, 0, st>>>(s.part, dw, S, a.Cout, a.Cin, kk,
                                                                                              a.transposed, Kp);
  else
    reduce_conv_partials_kernel<<<(unsigned)(((long long)a.Cout * K + 255) / 256), 256, 0, st>>>(s.part, dw, S, a.Cout, a.Cin, kk,
                                                                                            a.transposed, Kp, fast ? 1 : 0);
  DPMN_LAUNCH_CHECK();
  return 0;
}

int launch_conv_wgrad_tc_im2col(const ConvArgs& a, const float* dy, float* dw, const ConvTcScratch& s, cudaStream_t st) {
  if (s.t == DT_F16) return conv_wgrad_tc_im2col_t<__half>(a, dy, dw, s, st);
  if (s.t == DT_BF16) return conv_wgrad_tc_im2col_t<__nv_bfloat16>(a, dy, dw, s, st);
  return -1;
}

}  // namespace dpmn
