// Generic fp32 building blocks of the backward pass (sm_100a, SIMT): a strided GEMM with split-K atomic
// accumulation (every nn.Linear / 1x1-conv dgrad and wgrad of pgrm.py reduces to it), column / row sums (bias
// gradients), GELU' and LayerNorm backward.  The reference computes these through torch autograd; the parity
// target is the reference's .grad tensors (tests/golden/*_grad.npz).
#include "common.cuh"
#include "kernels.h"

namespace dpmn {

// =====================================================================================================
// C[z][m, n] (op)= sum_k A[z][m*sam + k*sak] * B[z][k*sbk + n*sbn]
// 64x64 tile, BK 16, 256 threads, 4x4 micro-tile.  grid.z = batch * ksplit; with mode 2 every (batch, k-chunk)
// atomically adds into C (c_bs may be 0: the batch is summed -- the weight gradient of a per-image GEMM).
// =====================================================================================================
constexpr int XM = 64, XN = 64, XK = 16, XPAD = 68;

__global__ void __launch_bounds__(256) gemm_gen_kernel(GemmGenArgs p) {
  __shared__ __align__(16) float As[XK][XPAD];
  __shared__ __align__(16) float Bs[XK][XPAD];
  const int z = blockIdx.z / p.ksplit, ks = blockIdx.z - z * p.ksplit;
  const int kchunk = (((p.K + p.ksplit - 1) / p.ksplit) + XK - 1) / XK * XK;
  const int kbeg = ks * kchunk;
  const int kend = min(p.K, kbeg + kchunk);
  const int m0 = blockIdx.y * XM, n0 = blockIdx.x * XN;
  const float* A = p.A + (long long)z * p.a_bs;
  const float* Bm = p.B + (long long)z * p.b_bs;
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const bool a_kfast = p.sak == 1;      // which index the loader walks fastest (the contiguous one)
  const bool b_nfast = p.sbn == 1;

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int k0 = kbeg; k0 < kend; k0 += XK) {
    float av[4], bv[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int e = tid + i * 256;
      const int ka = a_kfast ? (e & 15) : (e >> 6);
      const int ma = a_kfast ? (e >> 4) : (e & 63);
      av[i] = (m0 + ma < p.M && k0 + ka < kend) ? A[(long long)(m0 + ma) * p.sam + (long long)(k0 + ka) * p.sak] : 0.f;
      const int kb = b_nfast ? (e >> 6) : (e & 15);
      const int nb = b_nfast ? (e & 63) : (e >> 4);
      bv[i] = (n0 + nb < p.N && k0 + kb < kend) ? Bm[(long long)(k0 + kb) * p.sbk + (long long)(n0 + nb) * p.sbn] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int e = tid + i * 256;
      const int ka = a_kfast ? (e & 15) : (e >> 6);
      const int ma = a_kfast ? (e >> 4) : (e & 63);
      As[ka][ma] = av[i];
      const int kb = b_nfast ? (e >> 6) : (e & 15);
      const int nb = b_nfast ? (e & 63) : (e >> 4);
      Bs[kb][nb] = bv[i];
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < XK; ++k) {
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
  float* C = p.C + (long long)z * p.c_bs;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= p.M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= p.N) continue;
      float* dst = C + (long long)m * p.ldc + n;
      const float v = acc[i][j] * p.alpha;
      if (p.mode == 0) *dst = v;
      else if (p.mode == 1) *dst += v;
      else atomicAdd(dst, v);
    }
  }
}

int launch_gemm_gen(const GemmGenArgs& a, cudaStream_t st) {
  if (a.M < 1 || a.N < 1 || a.K < 1 || a.batch < 1 || a.ksplit < 1) return -1;
  if (a.mode != 2 && a.ksplit != 1) return -1;
  dim3 grid((a.N + XN - 1) / XN, (a.M + XM - 1) / XM, a.batch * a.ksplit);
  gemm_gen_kernel<<<grid, 256, 0, st>>>(a);
  DPMN_LAUNCH_CHECK();
  return 0;
}

// out[n] += sum_rows x[row*ld + n]        (bias gradient of a Linear)
__global__ void __launch_bounds__(256) colsum_kernel(const float* __restrict__ x, long long ld, int rows, int N,
                                                     int rows_per_block, float* __restrict__ out) {
  __shared__ float red[8][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int n = blockIdx.x * 32 + tx;
  const int r0 = blockIdx.y * rows_per_block;
  const int r1 = min(rows, r0 + rows_per_block);
  float s = 0.f;
  if (n < N)
    for (int r = r0 + ty; r < r1; r += 8) s += x[(long long)r * ld + n];
  red[ty][tx] = s;
  __syncthreads();
  if (ty == 0 && n < N) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += red[i][tx];
    atomicAdd(out + n, t);
  }
}

int launch_colsum(const float* x, long long ld, int rows, int N, float* out, cudaStream_t st) {
  const int rpb = 512;
  dim3 grid((N + 31) / 32, (rows + rpb - 1) / rpb);
  colsum_kernel<<<grid, 256, 0, st>>>(x, ld, rows, N, rpb, out);
  DPMN_LAUNCH_CHECK();
  return 0;
}

// out[row % mod] += sum_j x[row*len + j]   (bias gradient of the pointwise conv: rows are (image, channel) planes)
__global__ void __launch_bounds__(256) rowsum_kernel(const float* __restrict__ x, int rows, int len, int mod,
                                                     float* __restrict__ out) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= rows) return;
  const float* r = x + (long long)warp * len;
  float s = 0.f;
  for (int j = lane; j < len; j += 32) s += r[j];
  s = warp_sum(s);
  if (lane == 0) atomicAdd(out + warp % mod, s);
}

int launch_rowsum(const float* x, int rows, int len, int mod, float* out, cudaStream_t st) {
  rowsum_kernel<<<(rows * 32 + 255) / 256, 256, 0, st>>>(x, rows, len, mod, out);
  DPMN_LAUNCH_CHECK();
  return 0;
}

// g[i] *= gelu'(x[i])
__global__ void gelu_bwd_kernel(float* __restrict__ g, const float* __restrict__ x, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) g[i] *= gelu_grad(x[i]);
}

int launch_gelu_bwd(float* g, const float* x, long long n, cudaStream_t st) {
  gelu_bwd_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(g, x, n);
  DPMN_LAUNCH_CHECK();
  return 0;
}

// y[i] = a[i] + b[i]
__global__ void add_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ y, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) y[i] = a[i] + b[i];
}

int launch_add(const float* a, const float* b, float* y, long long n, cudaStream_t st) {
  add_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(a, b, y, n);
  DPMN_LAUNCH_CHECK();
  return 0;
}

__global__ void dropout_kernel(float* __restrict__ x, long long n, float p, unsigned long long seed, uint32_t site) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) x[i] *= drop_scale(p, seed, site, (unsigned long long)i);
}

int launch_dropout(float* x, long long n, float p, unsigned long long seed, uint32_t site, cudaStream_t st) {
  if (p <= 0.f) return 0;
  dropout_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(x, n, p, seed, site);
  DPMN_LAUNCH_CHECK();
  return 0;
}

__global__ void branch_combine_kernel(float* __restrict__ dst, const float* __restrict__ base, const float* __restrict__ y,
                                      long long n, long long per_image, float p_drop, uint32_t site_drop, float p_path,
                                      uint32_t site_path, unsigned long long seed) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float f = drop_scale(p_drop, seed, site_drop, (unsigned long long)i) *
                  drop_scale(p_path, seed, site_path, (unsigned long long)(i / per_image));
  const float v = y[i] * f;
  dst[i] = base != nullptr ? base[i] + v : v;
}

int launch_branch_combine(float* dst, const float* base, const float* y, long long n, long long per_image, float p_drop,
                          uint32_t site_drop, float p_path, uint32_t site_path, unsigned long long seed, cudaStream_t st) {
  branch_combine_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(dst, base, y, n, per_image, p_drop, site_drop, p_path,
                                                                    site_path, seed);
  DPMN_LAUNCH_CHECK();
  return 0;
}

// dst[z][c][r] = (T) src[z][r][c]: transposing fp32 -> 16-bit staging of a gradient / activation matrix, so that the
// weight-gradient contraction over rows becomes a K-contiguous tcgen05 GEMM (gemm_tc.cu).  32 x 32 smem tiles.
// copy (optional): the untransposed 16-bit copy of the same matrix, written from the same read
template <typename T>
__global__ void __launch_bounds__(256) transpose_convert_kernel(const float* __restrict__ src, T* __restrict__ dst, int R,
                                                                int Cc, long long src_bs, long long dst_bs,
                                                                T* __restrict__ copy = nullptr) {
  __shared__ float tile[32][33];
  const int z = blockIdx.z;
  const int r0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const float* s = src + (long long)z * src_bs;
  T* d = dst + (long long)z * dst_bs;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = r0 + ty + 8 * i, c = c0 + tx;
    const bool ok = r < R && c < Cc;
    const float v = ok ? s[(long long)r * Cc + c] : 0.f;
    tile[ty + 8 * i][tx] = v;
    if (ok && copy != nullptr) copy[(long long)z * src_bs + (long long)r * Cc + c] = from_f32<T>(v);
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int c = c0 + ty + 8 * i, r = r0 + tx;
    if (c < Cc && r < R) d[(long long)c * R + r] = from_f32<T>(tile[tx][ty + 8 * i]);
  }
}

int launch_transpose_convert(const float* src, void* dst, DType t, int batch, int R, int Cc, cudaStream_t st, void* copy16) {
  dim3 grid((Cc + 31) / 32, (R + 31) / 32, batch);
  const long long bs = (long long)R * Cc;
  if (t == DT_F16) transpose_convert_kernel<__half><<<grid, 256, 0, st>>>(src, (__half*)dst, R, Cc, bs, bs, (__half*)copy16);
  else if (t == DT_BF16)
    transpose_convert_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(src, (__nv_bfloat16*)dst, R, Cc, bs, bs, (__nv_bfloat16*)copy16);
  else return -1;
  DPMN_LAUNCH_CHECK();
  return 0;
}

// One pass over a gradient matrix src (R, Cc) fp32 for everything a Linear's backward wants from it (16-bit modes):
//   copy16 (R, Cc) for the data-gradient GEMM, trans16 (Cc, R) for the weight-gradient GEMM (contraction over rows),
//   colsum[c] += sum_r src[r][c] for the bias gradient  -- instead of convert + transpose_convert + colsum (three reads).
// Block = 32 columns x 128 rows (four 32 x 32 shared-memory tiles).
template <typename T>
__global__ void __launch_bounds__(256) stage_grad_kernel(const float* __restrict__ src, T* __restrict__ copy16,
                                                         T* __restrict__ trans16, float* __restrict__ colsum, int R, int Cc) {
  __shared__ float tile[32][33];
  __shared__ float red[8][33];
  const int c0 = blockIdx.x * 32, rb = blockIdx.y * 128;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  float cs = 0.f;
  for (int sub = 0; sub < 4; ++sub) {
    const int r0 = rb + sub * 32;
    if (r0 >= R) break;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int r = r0 + ty + 8 * i, c = c0 + tx;
      const bool ok = r < R && c < Cc;
      const float v = ok ? src[(long long)r * Cc + c] : 0.f;
      tile[ty + 8 * i][tx] = v;
      cs += v;
      if (ok && copy16 != nullptr) copy16[(long long)r * Cc + c] = from_f32<T>(v);
    }
    __syncthreads();
    if (trans16 != nullptr) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int c = c0 + ty + 8 * i, r = r0 + tx;
        if (c < Cc && r < R) trans16[(long long)c * R + r] = from_f32<T>(tile[tx][ty + 8 * i]);
      }
    }
    __syncthreads();
  }
  if (colsum != nullptr) {
    red[ty][tx] = cs;
    __syncthreads();
    if (ty == 0 && c0 + tx < Cc) {
      float t = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) t += red[i][tx];
      atomicAdd(colsum + c0 + tx, t);
    }
  }
}

int launch_stage_grad(const float* src, void* copy16, void* trans16, float* colsum, DType t, int R, int Cc, cudaStream_t st) {
  dim3 grid((Cc + 31) / 32, (R + 127) / 128);
  if (t == DT_F16) stage_grad_kernel<__half><<<grid, 256, 0, st>>>(src, (__half*)copy16, (__half*)trans16, colsum, R, Cc);
  else if (t == DT_BF16) stage_grad_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(src, (__nv_bfloat16*)copy16, (__nv_bfloat16*)trans16, colsum, R, Cc);
  else return -1;
  DPMN_LAUNCH_CHECK();
  return 0;
}

// Up to 16 small matrices (the weights of one PGRM) transposed + converted in ONE launch: segment = blockIdx.z, the grid's
// x / y extents cover the largest matrix.  dst[seg][c][r] = (T) src[seg][r][c].
template <typename T>
__global__ void __launch_bounds__(256) transpose_convert_batch_kernel(TransposeBatch tb) {
  __shared__ float tile[32][33];
  const int seg = blockIdx.z;
  const int R = tb.R[seg], Cc = tb.C[seg];
  const int r0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
  if (r0 >= R || c0 >= Cc) return;                      // uniform per block
  const float* __restrict__ s = tb.src[seg];
  T* __restrict__ d = reinterpret_cast<T*>(tb.dst[seg]);
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = r0 + ty + 8 * i, c = c0 + tx;
    tile[ty + 8 * i][tx] = (r < R && c < Cc) ? s[(long long)r * Cc + c] : 0.f;
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int c = c0 + ty + 8 * i, r = r0 + tx;
    if (c < Cc && r < R) d[(long long)c * R + r] = from_f32<T>(tile[tx][ty + 8 * i]);
  }
}

int launch_transpose_convert_batch(const TransposeBatch& tb, DType t, cudaStream_t st) {
  if (tb.count < 1 || tb.count > 16) return -1;
  int maxR = 0, maxC = 0;
  for (int i = 0; i < tb.count; ++i) { if (tb.R[i] > maxR) maxR = tb.R[i]; if (tb.C[i] > maxC) maxC = tb.C[i]; }
  dim3 grid((maxC + 31) / 32, (maxR + 31) / 32, tb.count);
  if (t == DT_F16) transpose_convert_batch_kernel<__half><<<grid, 256, 0, st>>>(tb);
  else if (t == DT_BF16) transpose_convert_batch_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(tb);
  else return -1;
  DPMN_LAUNCH_CHECK();
  return 0;
}

// dst[i] += sum_s partial[s*n + i]      (split-K partial products of the tensor-core weight gradients)
__global__ void reduce_partials_kernel(const float* __restrict__ partial, float* __restrict__ dst, int S, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float a = 0.f;
  for (int s = 0; s < S; ++s) a += partial[(long long)s * n + i];
  dst[i] += a;
}

int launch_reduce_partials(const float* partial, float* dst, int S, long long n, cudaStream_t st) {
  reduce_partials_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(partial, dst, S, n);
  DPMN_LAUNCH_CHECK();
  return 0;
}

// =====================================================================================================
// LayerNorm backward (eps 1e-5, nn.LayerNorm over the last dim).  One warp per row, C <= 256.
//   xhat = (x - mean) * rstd;  g = dy * w;  dx (+)= rstd * (g - mean(g) - xhat * mean(g * xhat))
//   dw += sum_rows dy * xhat;  db += sum_rows dy
// =====================================================================================================
constexpr int LNB_CPL = 8;

template <int CPL>       // channels per lane: C <= 32 * CPL (3 for the PGRM's C = 96: no predicated-off slots)
__global__ void __launch_bounds__(256) layernorm_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ x,
                                                            const float* __restrict__ w, float* __restrict__ dx,
                                                            int accumulate, float* __restrict__ dw,
                                                            float* __restrict__ db, int rows, int C) {
  __shared__ float red_w[8][256], red_b[8][256];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int gw = blockIdx.x * 8 + warp, nw = gridDim.x * 8;
  float aw[CPL], ab[CPL], wv[CPL];
#pragma unroll
  for (int i = 0; i < CPL; ++i) {
    aw[i] = 0.f; ab[i] = 0.f;
    const int c = lane + 32 * i;
    wv[i] = c < C ? w[c] : 0.f;
  }
  const float invC = 1.0f / (float)C;
  for (int r = gw; r < rows; r += nw) {
    float xv[CPL], gv[CPL];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < CPL; ++i) {
      const int c = lane + 32 * i;
      xv[i] = c < C ? x[(long long)r * C + c] : 0.f;
      gv[i] = c < C ? dy[(long long)r * C + c] : 0.f;
      s += xv[i];
    }
    const float mean = warp_sum(s) * invC;
    float v = 0.f;
#pragma unroll
    for (int i = 0; i < CPL; ++i) {
      const int c = lane + 32 * i;
      const float d = c < C ? xv[i] - mean : 0.f;
      xv[i] = d;
      v += d * d;
    }
    const float rstd = rsqrtf(warp_sum(v) * invC + 1e-5f);
    float m1 = 0.f, m2 = 0.f;
#pragma unroll
    for (int i = 0; i < CPL; ++i) {
      xv[i] *= rstd;                       // xhat
      aw[i] = fmaf(gv[i], xv[i], aw[i]);
      ab[i] += gv[i];
      gv[i] *= wv[i];                      // g
      m1 += gv[i];
      m2 = fmaf(gv[i], xv[i], m2);
    }
    m1 = warp_sum(m1) * invC;
    m2 = warp_sum(m2) * invC;
#pragma unroll
    for (int i = 0; i < CPL; ++i) {
      const int c = lane + 32 * i;
      if (c < C) {
        const float d = rstd * (gv[i] - m1 - xv[i] * m2);
        float* dst = dx + (long long)r * C + c;
        *dst = accumulate ? *dst + d : d;
      }
    }
  }
#pragma unroll
  for (int i = 0; i < CPL; ++i) { red_w[warp][lane + 32 * i] = aw[i]; red_b[warp][lane + 32 * i] = ab[i]; }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += 256) {
    float sw = 0.f, sb = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) { sw += red_w[i][c]; sb += red_b[i][c]; }
    atomicAdd(dw + c, sw);
    atomicAdd(db + c, sb);
  }
}

int launch_layernorm_bwd(const float* dy, const float* x, const float* w, float* dx, int accumulate, float* dw,
                         float* db, int rows, int C, cudaStream_t st) {
  if (C > 32 * LNB_CPL) return -2;
  int blocks = (rows + 63) / 64;
  if (blocks > 592) blocks = 592;
  if (C <= 96) layernorm_bwd_kernel<3><<<blocks, 256, 0, st>>>(dy, x, w, dx, accumulate, dw, db, rows, C);
  else layernorm_bwd_kernel<LNB_CPL><<<blocks, 256, 0, st>>>(dy, x, w, dx, accumulate, dw, db, rows, C);
  DPMN_LAUNCH_CHECK();
  return 0;
}

}  // namespace dpmn
