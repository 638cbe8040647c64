// Backward kernels of the PGRM (fp32 SIMT, sm_100a): the pieces of PGRM.forward (pgrm.py:546-565) that are not a
// plain Linear -- window attention, the SK gate, the raw-view depthwise conv, patch embed (+ prior fusion) and the
// conv / LeakyReLU / PixelShuffle / affine-mix head.  Parameter gradients ACCUMULATE (atomicAdd) into the caller's
// buffers; activation gradients are written.  Parity target: the reference's autograd .grad tensors
// (tests/golden/pgrm_*_grad.npz).
#include <cstdlib>
#include "common.cuh"
#include "kernels.h"

namespace dpmn {

// =====================================================================================================
// Window attention backward                                                              pgrm.py:197-268
// CTA = 64 consecutive window-major rows (64/N whole windows) of one (image, group), one head; 128 threads.
//   P = softmax(S),  dV = P^T dO,  dP = dO V^T,  dS = P o (dP - rowsum(dP o P)),
//   dQ = scale * dS K,  dK = scale * dS^T Q,  d table[idx(n,m), head] += dS[n,m]
// q / kv / dq / dkv are in TOKEN order (the roll + partition gather is undone on the way out); dO is the
// gradient of the window-major attention output (quirk 1: rows of `attn` are window-major).
// =====================================================================================================
constexpr int AB_ROWS = 64;

__global__ void __launch_bounds__(128) window_attn_bwd_kernel(
    const float* __restrict__ q, const float* __restrict__ kv, const float* __restrict__ d_attn,
    float* __restrict__ dq, float* __restrict__ dkv, const float* __restrict__ table, float* __restrict__ d_table,
    int C, int hpg, int ch0, int D, int H, int W, int ws, int shift, float scale, float p_drop, unsigned long long seed,
    uint32_t site, int g_index, int G, int rnd) {
  // rnd (1 fp16, 2 bf16): the forward ran on the tcgen05 kernel, which read q / k / v as 16-bit values and rounded the
  // unnormalised probabilities to 16 bits (attn2_tc.cu); P is recomputed from the same rounded values here
  auto r16 = [rnd](float x) { return rnd == 1 ? __half2float(__float2half_rn(x)) : (rnd == 2 ? __bfloat162float(__float2bfloat16_rn(x)) : x); };
  extern __shared__ float sm[];
  const int N = ws * ws, NS = N + 1, L = H * W, DS = D + 1, tw = 2 * ws - 1, TT = tw * tw;
  float* sQ = sm;                       // [64][DS]
  float* sK = sQ + AB_ROWS * DS;
  float* sV = sK + AB_ROWS * DS;
  float* sO = sV + AB_ROWS * DS;        // dO
  float* sP = sO + AB_ROWS * DS;        // [64][N + 1]
  float* sS = sP + AB_ROWS * NS;        // dP, then dS
  float* sTab = sS + AB_ROWS * NS;       // [TT] bias column of this head
  float* sDT = sTab + TT;               // [TT] table-gradient accumulators
  int* sTok = reinterpret_cast<int*>(sDT + TT);   // [64] original token of each row
  int* sLab = sTok + AB_ROWS;                     // [64] shift-mask label

  const int head = blockIdx.y;
  const int b = blockIdx.z / 1;
  const int row0 = blockIdx.x * AB_ROWS;          // window-major row within the image
  const int tid = threadIdx.x;
  const int ch = ch0 + head * D;

  for (int i = tid; i < TT; i += 128) { sTab[i] = table[i * hpg + head]; sDT[i] = 0.f; }
  if (tid < AB_ROWS) {
    const WinCoord wc = window_row_to_token(row0 + tid, H, W, ws, shift);
    sTok[tid] = wc.token;
    sLab[tid] = shift > 0 ? shift_region_label(wc.hp, wc.wp, H, W, ws, shift) : 0;
  }
  __syncthreads();
  for (int i = tid; i < AB_ROWS * D; i += 128) {
    const int r = i / D, e = i - r * D;
    const long long tok = (long long)b * L + sTok[r];
    sQ[r * DS + e] = r16(q[tok * C + ch + e]);
    sK[r * DS + e] = r16(kv[tok * 2 * C + ch + e]);
    sV[r * DS + e] = r16(kv[tok * 2 * C + C + ch + e]);
    sO[r * DS + e] = d_attn[((long long)b * L + row0 + r) * C + ch + e];
  }
  __syncthreads();
  // scores and dP
  for (int i = tid; i < AB_ROWS * N; i += 128) {
    const int r = i / N, m = i - r * N;
    const int kr = (r / N) * N + m;
    const int n = r % N;
    float s = 0.f, dp = 0.f;
    for (int e = 0; e < D; ++e) {
      s = fmaf(sQ[r * DS + e], sK[kr * DS + e], s);
      dp = fmaf(sO[r * DS + e], sV[kr * DS + e], dp);
    }
    s = s * scale + sTab[(n / ws - m / ws + ws - 1) * tw + (n % ws - m % ws + ws - 1)];
    if (sLab[r] != sLab[kr]) s += -100.0f;
    // forward with attn_drop: O = (P o M) V with M in {0, 1/(1-p)}  ->  dP = (dO V^T) o M
    if (p_drop > 0.f)
      dp *= drop_scale(p_drop, seed, site, ((((unsigned long long)b * G + g_index) * hpg + head) * L + row0 + r) * N + m);
    sP[r * NS + m] = s;
    sS[r * NS + m] = dp;
  }
  __syncthreads();
  if (tid < AB_ROWS) {
    float* pr = sP + tid * NS;
    float* dr = sS + tid * NS;
    float mx = -INFINITY;
    for (int m = 0; m < N; ++m) mx = fmaxf(mx, pr[m]);
    float den = 0.f;
    for (int m = 0; m < N; ++m) { const float e = expf(pr[m] - mx); pr[m] = r16(e); den += e; }
    const float inv = 1.0f / den;
    float dot = 0.f;
    for (int m = 0; m < N; ++m) { pr[m] *= inv; dot = fmaf(pr[m], dr[m], dot); }
    for (int m = 0; m < N; ++m) dr[m] = pr[m] * (dr[m] - dot);
  }
  __syncthreads();
  for (int i = tid; i < AB_ROWS * N; i += 128) {
    const int r = i / N, m = i - r * N;
    const int n = r % N;
    atomicAdd(&sDT[(n / ws - m / ws + ws - 1) * tw + (n % ws - m % ws + ws - 1)], sS[r * NS + m]);
  }
  for (int i = tid; i < AB_ROWS * D; i += 128) {
    const int r = i / D, e = i - r * D;
    const int w0 = (r / N) * N, nr = r - w0;        // first row of r's window; r's index within it
    float aq = 0.f, ak = 0.f, av = 0.f;
    for (int m = 0; m < N; ++m) {
      aq = fmaf(sS[r * NS + m], sK[(w0 + m) * DS + e], aq);            // dQ[r] = sum_m dS[r,m] K[m]
      ak = fmaf(sS[(w0 + m) * NS + nr], sQ[(w0 + m) * DS + e], ak);    // dK[r] = sum_n dS[n,r] Q[n]
      float pm = sP[(w0 + m) * NS + nr];                               // dV[r] = sum_n (P o M)[n,r] dO[n]
      if (p_drop > 0.f)
        pm *= drop_scale(p_drop, seed, site, ((((unsigned long long)b * G + g_index) * hpg + head) * L + row0 + w0 + m) * N + nr);
      av = fmaf(pm, sO[(w0 + m) * DS + e], av);
    }
    const long long tok = (long long)b * L + sTok[r];
    dq[tok * C + ch + e] = aq * scale;
    dkv[tok * 2 * C + ch + e] = ak * scale;
    dkv[tok * 2 * C + C + ch + e] = av;
  }
  __syncthreads();
  for (int i = tid; i < TT; i += 128) atomicAdd(d_table + i * hpg + head, sDT[i]);
}

int launch_window_attn_bwd(const AttnBwdArgs& a, cudaStream_t st) {
  const int L = a.H * a.W;
  const int cg = a.C / a.n_groups;
  const int D = cg / a.heads_per_group;
  for (int g = 0; g < a.n_groups; ++g) {
    const int ws = a.window[g], N = ws * ws;
    if (N > AB_ROWS || AB_ROWS % N || L % AB_ROWS || a.H % ws || a.W % ws) return -2;
    const int tw = 2 * ws - 1;
    const size_t smem = (size_t)(4 * AB_ROWS * (D + 1) + 2 * AB_ROWS * (N + 1) + 2 * tw * tw + 2 * AB_ROWS) * sizeof(float);
    if (smem > 200 * 1024) return -2;
    static size_t attr = 0;
    if (smem > 48 * 1024 && smem > attr) {
      DPMN_CUDA_TRY(cudaFuncSetAttribute(window_attn_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      attr = smem;
    }
    dim3 grid(L / AB_ROWS, a.heads_per_group, a.B);
    window_attn_bwd_kernel<<<grid, 128, smem, st>>>(a.q, a.kv, a.d_attn, a.dq, a.dkv, a.table[g], a.d_table[g], a.C,
                                                    a.heads_per_group, g * cg, D, a.H, a.W, ws, a.shift[g],
                                                    1.0f / sqrtf((float)D), a.p_drop, a.seed, a.site, g, a.n_groups, a.round16);
    DPMN_LAUNCH_CHECK();
  }
  return 0;
}

// =====================================================================================================
// SK gate (pgrm.py:79-96), training form: the pieces are kept apart because backward needs them.
// =====================================================================================================
// S[b, c] = mean_l GELU(F[b, l, c])
__global__ void __launch_bounds__(256) sk_pool_kernel(const float* __restrict__ F, float* __restrict__ S, int L, int C) {
  __shared__ float red[8][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + tx, b = blockIdx.y;
  float s = 0.f;
  if (c < C)
    for (int l = ty; l < L; l += 8) s += gelu_erf(F[((long long)b * L + l) * C + c]);
  red[ty][tx] = s;
  __syncthreads();
  if (ty == 0 && c < C) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += red[i][tx];
    S[b * C + c] = t / (float)L;
  }
}

// zpre = fc1 S + b1; z = GELU(zpre); u = fc2 z + b2; a = softmax over the G groups of u viewed (G, cg)
__global__ void __launch_bounds__(128) sk_mlp_fwd_kernel(const float* __restrict__ S, const float* __restrict__ w1,
                                                         const float* __restrict__ b1, const float* __restrict__ w2,
                                                         const float* __restrict__ b2, float* __restrict__ zpre_out,
                                                         float* __restrict__ a_out, int C, int G, int hidden) {
  __shared__ float sS[256], sZ[64], sU[256];
  const int b = blockIdx.x, cg = C / G;
  for (int c = threadIdx.x; c < C; c += 128) sS[c] = S[b * C + c];
  __syncthreads();
  for (int j = threadIdx.x; j < hidden; j += 128) {
    float s = b1[j];
    for (int c = 0; c < C; ++c) s = fmaf(w1[j * C + c], sS[c], s);
    zpre_out[b * hidden + j] = s;
    sZ[j] = gelu_erf(s);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < C; i += 128) {
    float s = b2[i];
    for (int j = 0; j < hidden; ++j) s = fmaf(w2[i * hidden + j], sZ[j], s);
    sU[i] = s;
  }
  __syncthreads();
  for (int c = threadIdx.x; c < cg; c += 128) {
    float mx = -INFINITY;
    for (int m = 0; m < G; ++m) mx = fmaxf(mx, sU[m * cg + c]);
    float den = 0.f;
    for (int m = 0; m < G; ++m) den += expf(sU[m * cg + c] - mx);
    for (int m = 0; m < G; ++m) a_out[b * C + m * cg + c] = expf(sU[m * cg + c] - mx) / den;
  }
}

// Xs[row, c] = sum_m a[b, m*cg + c] * A[row, m*cg + c]
__global__ void sk_mix_kernel(const float* __restrict__ A, const float* __restrict__ a, float* __restrict__ Xs,
                              long long total, int L, int C, int G) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int cg = C / G;
  const long long row = i / cg;
  const int c = (int)(i - row * cg);
  const int b = (int)(row / L);
  float s = 0.f;
  for (int m = 0; m < G; ++m) s = fmaf(a[b * C + m * cg + c], A[row * C + m * cg + c], s);
  Xs[i] = s;
}

int launch_sk_train_fwd(const float* F, const float* A, const float* w1, const float* b1, const float* w2,
                        const float* b2, float* S, float* zpre, float* a, float* Xs, int B, int L, int C, int G,
                        cudaStream_t st) {
  const int hidden = C / G / 2;
  if (C > 256 || hidden > 64 || hidden < 1) return -2;
  sk_pool_kernel<<<dim3((C + 31) / 32, B), 256, 0, st>>>(F, S, L, C);
  DPMN_LAUNCH_CHECK();
  sk_mlp_fwd_kernel<<<B, 128, 0, st>>>(S, w1, b1, w2, b2, zpre, a, C, G, hidden);
  DPMN_LAUNCH_CHECK();
  const long long total = (long long)B * L * (C / G);
  sk_mix_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(A, a, Xs, total, L, C, G);
  DPMN_LAUNCH_CHECK();
  return 0;
}

// da[b, ch] = sum_l dXs[b, l, ch % cg] * A[b, l, ch]
__global__ void __launch_bounds__(256) sk_da_kernel(const float* __restrict__ dXs, const float* __restrict__ A,
                                                    float* __restrict__ da, int L, int C, int cg) {
  __shared__ float red[8][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int ch = blockIdx.x * 32 + tx, b = blockIdx.y;
  float s = 0.f;
  if (ch < C) {
    const int c = ch % cg;
    for (int l = ty; l < L; l += 8) {
      const long long row = (long long)b * L + l;
      s = fmaf(dXs[row * cg + c], A[row * C + ch], s);
    }
  }
  red[ty][tx] = s;
  __syncthreads();
  if (ty == 0 && ch < C) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += red[i][tx];
    da[b * C + ch] = t;
  }
}

// softmax / fc2 / GELU / fc1 backward of the gate, one CTA per image -> dS (B, C); weight grads by atomics
__global__ void __launch_bounds__(128) sk_mlp_bwd_kernel(const float* __restrict__ da, const float* __restrict__ a,
                                                         const float* __restrict__ zpre, const float* __restrict__ S,
                                                         const float* __restrict__ w1, const float* __restrict__ w2,
                                                         float* __restrict__ dw1, float* __restrict__ db1,
                                                         float* __restrict__ dw2, float* __restrict__ db2,
                                                         float* __restrict__ dS, int C, int G, int hidden) {
  __shared__ float sU[256], sZ[64], sDZ[64], sS[256];
  const int b = blockIdx.x, cg = C / G;
  for (int c = threadIdx.x; c < C; c += 128) sS[c] = S[b * C + c];
  for (int j = threadIdx.x; j < hidden; j += 128) sZ[j] = gelu_erf(zpre[b * hidden + j]);
  for (int c = threadIdx.x; c < cg; c += 128) {
    float dot = 0.f;
    for (int m = 0; m < G; ++m) dot = fmaf(a[b * C + m * cg + c], da[b * C + m * cg + c], dot);
    for (int m = 0; m < G; ++m) sU[m * cg + c] = a[b * C + m * cg + c] * (da[b * C + m * cg + c] - dot);   // du
  }
  __syncthreads();
  for (int i = threadIdx.x; i < C; i += 128) {
    atomicAdd(db2 + i, sU[i]);
    for (int j = 0; j < hidden; ++j) atomicAdd(dw2 + i * hidden + j, sU[i] * sZ[j]);
  }
  for (int j = threadIdx.x; j < hidden; j += 128) {
    float s = 0.f;
    for (int i = 0; i < C; ++i) s = fmaf(sU[i], w2[i * hidden + j], s);
    s *= gelu_grad(zpre[b * hidden + j]);
    sDZ[j] = s;
    atomicAdd(db1 + j, s);
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += 128) {
    float s = 0.f;
    for (int j = 0; j < hidden; ++j) {
      s = fmaf(sDZ[j], w1[j * C + c], s);
      atomicAdd(dw1 + j * C + c, sDZ[j] * sS[c]);
    }
    dS[b * C + c] = s;
  }
}

// F[row, c] <- d_out[row, c] + dS[b, c] / L * gelu'(F[row, c])         (dF, in place on F)
// dA[row, ch] = dXs[row, ch % cg] * a[b, ch]
__global__ void sk_df_da_kernel(float* __restrict__ F, const float* __restrict__ d_out, const float* __restrict__ dS,
                                const float* __restrict__ dXs, const float* __restrict__ a, float* __restrict__ dA,
                                long long total, int L, int C, int cg) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const long long row = i / C;
  const int ch = (int)(i - row * C);
  const int b = (int)(row / L);
  F[i] = fmaf(dS[b * C + ch] / (float)L, gelu_grad(F[i]), d_out[i]);
  dA[i] = dXs[row * cg + ch % cg] * a[b * C + ch];
}

int launch_sk_bwd(const SkBwdArgs& s, cudaStream_t st) {
  const int cg = s.C / s.G, hidden = cg / 2;
  if (s.C > 256 || hidden > 64) return -2;
  sk_da_kernel<<<dim3((s.C + 31) / 32, s.B), 256, 0, st>>>(s.dXs, s.A, s.da, s.L, s.C, cg);
  DPMN_LAUNCH_CHECK();
  sk_mlp_bwd_kernel<<<s.B, 128, 0, st>>>(s.da, s.a, s.zpre, s.S, s.w1, s.w2, s.dw1, s.db1, s.dw2, s.db2, s.dS, s.C, s.G,
                                         hidden);
  DPMN_LAUNCH_CHECK();
  const long long total = (long long)s.B * s.L * s.C;
  sk_df_da_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(s.F, s.d_out, s.dS, s.dXs, s.a, s.dA, total, s.L, s.C, cg);
  DPMN_LAUNCH_CHECK();
  return 0;
}

// =====================================================================================================
// Mlp depthwise conv on the raw view (pgrm.py:33-36), training form and backward.
// h1pre (B, L, hid) is fc1's output BEFORE GELU; its raw view is (B, hid, side, side).  The forward writes the
// depthwise output before GELU (dtpre) and after (dt), both pixel-major (B, L, hid) so the pointwise conv is a
// K-contiguous GEMM.  CTA = (image, 32 channels, 8 rows).
// =====================================================================================================
constexpr int DWT_ROWS = 8;

__global__ void __launch_bounds__(256) dwconv_train_fwd_kernel(const float* __restrict__ h1pre, float* __restrict__ dtpre,
                                                               float* __restrict__ dt, const float* __restrict__ w,
                                                               const float* __restrict__ bias, int L, int hid, int side,
                                                               float p_drop, unsigned long long seed, uint32_t site) {
  extern __shared__ float sm[];
  const int b = blockIdx.z, c0 = blockIdx.y * 32, y0 = blockIdx.x * DWT_ROWS;
  const int rows_in = DWT_ROWS + 2;
  const int cstride = rows_in * side + 1;
  const float* hb = h1pre + (long long)b * L * hid;
  for (int i = threadIdx.x; i < 32 * rows_in * side; i += blockDim.x) {
    const int c = i / (rows_in * side);
    const int r = i - c * rows_in * side;
    const int yy = y0 - 1 + r / side, xx = r % side;
    float v = 0.f;
    if (yy >= 0 && yy < side) {
      const long long o = (long long)(c0 + c) * L + yy * side + xx;
      v = gelu_erf(hb[o]) * drop_scale(p_drop, seed, site, (unsigned long long)b * L * hid + o);   // pgrm.py:31-32
    }
    sm[c * cstride + r] = v;
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int c = c0 + lane;
  float wk[9];
#pragma unroll
  for (int i = 0; i < 9; ++i) wk[i] = w[c * 9 + i];
  const float bb = bias[c];
  const float* pl = sm + lane * cstride;
  for (int pi = warp; pi < DWT_ROWS * side; pi += 8) {
    const int yl = pi / side, xx = pi - yl * side;
    if (y0 + yl >= side) break;
    float acc = bb;
#pragma unroll
    for (int ky = 0; ky < 3; ++ky)
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        const int xc = xx + kx - 1;
        if (xc >= 0 && xc < side) acc = fmaf(pl[(yl + ky) * side + xc], wk[ky * 3 + kx], acc);
      }
    const long long o = ((long long)b * L + (y0 + yl) * side + xx) * hid + c;
    dtpre[o] = acc;
    dt[o] = gelu_erf(acc);
  }
}

// SIDE = 32 fast path of the two kernels around here: compile-time index arithmetic (the generic kernels spend a third of
// their issue slots on runtime div / mod), 16 channels per CTA so that three CTAs share an SM, the input halo tile loaded
// with 16-byte accesses, and -- in the backward -- erf evaluated ONCE per element of h1pre for both GELU(h) (weight
// gradient) and GELU'(h) (data gradient), with the dropout hash evaluated once as well.
constexpr int DWF_CH = 16;

__device__ __forceinline__ void gelu_and_grad(float x, float& g, float& dg) {
  const float cdf = 0.5f * (1.0f + erff(x * 0.70710678118654752440f));
  const float pdf = 0.39894228040143267794f * expf(-0.5f * x * x);
  g = x * cdf;
  dg = fmaf(x, pdf, cdf);
}

template <int SIDE>
__global__ void __launch_bounds__(256) dwconv_train_fwd_fast_kernel(const float* __restrict__ h1pre, float* __restrict__ dtpre,
                                                                    float* __restrict__ dt, const float* __restrict__ w,
                                                                    const float* __restrict__ bias, int hid, float p_drop,
                                                                    unsigned long long seed, uint32_t site,
                                                                    void* __restrict__ dt16, int type16) {
  // dt16 (optional): a 16-bit copy of dt, the operand of the pointwise-conv GEMM that follows (no convert pass of its own)
  constexpr int L = SIDE * SIDE, ROWS_IN = DWT_ROWS + 2, CSTRIDE = ROWS_IN * SIDE + 2;   // even stride: see the store below
  __shared__ float sm[DWF_CH * CSTRIDE];
  const int b = blockIdx.z, c0 = blockIdx.y * DWF_CH, y0 = blockIdx.x * DWT_ROWS;
  const float* hb = h1pre + (long long)b * L * hid;
  constexpr int VEC_ROW = SIDE / 4, N_VEC = DWF_CH * ROWS_IN * VEC_ROW;
  for (int i = threadIdx.x; i < N_VEC; i += 256) {
    const int xv = i % VEC_ROW, r = (i / VEC_ROW) % ROWS_IN, c = i / (VEC_ROW * ROWS_IN);
    const int yy = y0 - 1 + r;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (yy >= 0 && yy < SIDE) {
      const long long o = (long long)(c0 + c) * L + yy * SIDE + xv * 4;
      const float4 hv = *reinterpret_cast<const float4*>(hb + o);
      const unsigned long long e = (unsigned long long)b * L * hid + o;
      v.x = gelu_erf(hv.x) * drop_scale(p_drop, seed, site, e);                                   // pgrm.py:31-32
      v.y = gelu_erf(hv.y) * drop_scale(p_drop, seed, site, e + 1);
      v.z = gelu_erf(hv.z) * drop_scale(p_drop, seed, site, e + 2);
      v.w = gelu_erf(hv.w) * drop_scale(p_drop, seed, site, e + 3);
    }
    float* dst = sm + c * CSTRIDE + r * SIDE + xv * 4;
    *reinterpret_cast<float2*>(dst) = make_float2(v.x, v.y);
    *reinterpret_cast<float2*>(dst + 2) = make_float2(v.z, v.w);
  }
  __syncthreads();
  // lane -> (channel = lane & 15, pixel parity = lane >> 4): a half-warp writes 64 contiguous bytes of a pixel's channels
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int cl = lane & 15, c = c0 + cl;
  float wk[9];
#pragma unroll
  for (int i = 0; i < 9; ++i) wk[i] = w[c * 9 + i];
  const float bb = bias[c];
  const float* pl = sm + cl * CSTRIDE;
#pragma unroll 2
  for (int pi = warp * 2 + (lane >> 4); pi < DWT_ROWS * SIDE; pi += 16) {
    const int yl = pi / SIDE, xx = pi % SIDE;
    float acc = bb;
#pragma unroll
    for (int ky = 0; ky < 3; ++ky)
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        const int xc = xx + kx - 1;
        if (xc >= 0 && xc < SIDE) acc = fmaf(pl[(yl + ky) * SIDE + xc], wk[ky * 3 + kx], acc);
      }
    const long long o = ((long long)b * L + (y0 + yl) * SIDE + xx) * hid + c;
    dtpre[o] = acc;
    const float y = gelu_erf(acc);
    dt[o] = y;
    if (dt16 != nullptr) {
      if (type16 == DT_F16) reinterpret_cast<__half*>(dt16)[o] = __float2half_rn(y);
      else reinterpret_cast<__nv_bfloat16*>(dt16)[o] = __float2bfloat16_rn(y);
    }
  }
}

int launch_dwconv_train_fwd(const float* h1pre, float* dtpre, float* dt, const float* w, const float* b, int B, int L,
                            int hid, float p_drop, unsigned long long seed, uint32_t site, cudaStream_t st, void* dt16,
                            DType type16) {
  const int side = (int)(sqrtf((float)L) + 0.5f);
  if (side * side != L || hid % 32 || side % DWT_ROWS) return -2;
  if (dt16 != nullptr && type16 != DT_F16 && type16 != DT_BF16) return -2;
  if (side == 32 && hid % DWF_CH == 0 && (reinterpret_cast<uintptr_t>(h1pre) & 15) == 0) {
    dwconv_train_fwd_fast_kernel<32><<<dim3(32 / DWT_ROWS, hid / DWF_CH, B), 256, 0, st>>>(h1pre, dtpre, dt, w, b, hid, p_drop,
                                                                                          seed, site, dt16, (int)type16);
    DPMN_LAUNCH_CHECK();
    return 0;
  }
  const size_t smem = (size_t)32 * ((DWT_ROWS + 2) * side + 1) * sizeof(float);
  if (smem > 200 * 1024) return -2;
  if (smem > 48 * 1024)
    DPMN_CUDA_TRY(cudaFuncSetAttribute(dwconv_train_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dwconv_train_fwd_kernel<<<dim3(side / DWT_ROWS, hid / 32, B), 256, smem, st>>>(h1pre, dtpre, dt, w, b, L, hid, side,
                                                                                p_drop, seed, site);
  DPMN_LAUNCH_CHECK();
  if (dt16 != nullptr) return launch_convert(dt, dt16, type16, (long long)B * L * hid, st);
  return 0;
}

// Backward: g = d_dt * gelu'(dtpre) (pixel-major in);  d_h1pre[c][y][x] = gelu'(h1pre) * sum_taps w[c,tap] g[c][y-ky+1][x-kx+1]
// (raw layout out);  dw[c,tap] += sum g[c][y][x] * GELU(h1pre)[c][y+ky-1][x+kx-1];  db[c] += sum g.
__global__ void __launch_bounds__(256) dwconv_bwd_kernel(const float* __restrict__ d_dt, const float* __restrict__ dtpre,
                                                         const float* __restrict__ h1pre, const float* __restrict__ w,
                                                         float* __restrict__ d_h1pre, float* __restrict__ dw,
                                                         float* __restrict__ db, int L, int hid, int side, float p_drop,
                                                         unsigned long long seed, uint32_t site) {
  extern __shared__ float sm[];
  const int b = blockIdx.z, c0 = blockIdx.y * 32, y0 = blockIdx.x * DWT_ROWS;
  const int rows_in = DWT_ROWS + 2;
  const int cstride = rows_in * side + 1;
  float* sG = sm;                         // [32][rows_in*side (+1)]  g with one halo row each side
  float* sH = sm + 32 * cstride;          // same shape: GELU(h1pre)
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // g: pixel-major global -> lane = channel (coalesced), smem [c][r]
  for (int r = warp; r < rows_in * side; r += 8) {
    const int yy = y0 - 1 + r / side, xx = r % side;
    float v = 0.f;
    if (yy >= 0 && yy < side) {
      const long long o = ((long long)b * L + yy * side + xx) * hid + c0 + lane;
      v = d_dt[o] * gelu_grad(dtpre[o]);
    }
    sG[lane * cstride + r] = v;
  }
  const float* hb = h1pre + (long long)b * L * hid;
  for (int i = threadIdx.x; i < 32 * rows_in * side; i += blockDim.x) {
    const int c = i / (rows_in * side);
    const int r = i - c * rows_in * side;
    const int yy = y0 - 1 + r / side, xx = r % side;
    float v = 0.f;
    if (yy >= 0 && yy < side) {
      const long long o = (long long)(c0 + c) * L + yy * side + xx;
      v = gelu_erf(hb[o]) * drop_scale(p_drop, seed, site, (unsigned long long)b * L * hid + o);
    }
    sH[c * cstride + r] = v;
  }
  __syncthreads();
  // data gradient, raw (channel-plane) layout: consecutive threads -> consecutive x
  for (int i = threadIdx.x; i < 32 * DWT_ROWS * side; i += blockDim.x) {
    const int c = i / (DWT_ROWS * side);
    const int r = i - c * DWT_ROWS * side;
    const int yl = r / side, xx = r - yl * side;
    const float* g = sG + c * cstride;
    const float* wk = w + (c0 + c) * 9;
    float acc = 0.f;
#pragma unroll
    for (int ky = 0; ky < 3; ++ky)
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        const int xs = xx - kx + 1;                  // source pixel of g: (y - ky + 1, x - kx + 1)
        if (xs >= 0 && xs < side) acc = fmaf(g[(yl + 2 - ky) * side + xs], wk[ky * 3 + kx], acc);
      }
    const long long o = (long long)b * L * hid + (long long)(c0 + c) * L + (y0 + yl) * side + xx;
    d_h1pre[o] = acc * gelu_grad(h1pre[o]) * drop_scale(p_drop, seed, site, (unsigned long long)o);
  }
  // weight / bias gradient: warp w owns channels 4w .. 4w+3, lanes split the 8 x side positions
  for (int cc = 0; cc < 4; ++cc) {
    const int c = warp * 4 + cc;
    const float* g = sG + c * cstride + side;          // row y0
    const float* hh = sH + c * cstride;
    float acc[9], sb = 0.f;
#pragma unroll
    for (int t = 0; t < 9; ++t) acc[t] = 0.f;
    for (int r = lane; r < DWT_ROWS * side; r += 32) {
      const int yl = r / side, xx = r - yl * side;
      const float gv = g[r];
      sb += gv;
#pragma unroll
      for (int ky = 0; ky < 3; ++ky)
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
          const int xs = xx + kx - 1;
          if (xs >= 0 && xs < side) acc[ky * 3 + kx] = fmaf(gv, hh[(yl + ky) * side + xs], acc[ky * 3 + kx]);
        }
    }
#pragma unroll
    for (int t = 0; t < 9; ++t) {
      const float v = warp_sum(acc[t]);
      if (lane == 0) atomicAdd(dw + (c0 + c) * 9 + t, v);
    }
    sb = warp_sum(sb);
    if (lane == 0) atomicAdd(db + c0 + c, sb);
  }
}

template <int SIDE>
__global__ void __launch_bounds__(256) dwconv_bwd_fast_kernel(const float* __restrict__ d_dt, const float* __restrict__ dtpre,
                                                              const float* __restrict__ h1pre, const float* __restrict__ w,
                                                              float* __restrict__ d_h1pre, float* __restrict__ dw,
                                                              float* __restrict__ db, int hid, float p_drop,
                                                              unsigned long long seed, uint32_t site) {
  constexpr int L = SIDE * SIDE, ROWS_IN = DWT_ROWS + 2, CSTRIDE = ROWS_IN * SIDE + 2, DSTRIDE = DWT_ROWS * SIDE;
  extern __shared__ float sm[];
  float* sG = sm;                              // [16][CSTRIDE]  g = d_dt * GELU'(dtpre), one halo row each side
  float* sH = sG + DWF_CH * CSTRIDE;           // [16][CSTRIDE]  GELU(h1pre) * dropout scale, same rows
  float* sD = sH + DWF_CH * CSTRIDE;           // [16][DSTRIDE]  GELU'(h1pre) * dropout scale, interior rows
  const int b = blockIdx.z, c0 = blockIdx.y * DWF_CH, y0 = blockIdx.x * DWT_ROWS;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  {
    // g: pixel-major global; a half-warp reads the 16 channels of one pixel (64 contiguous bytes)
    const int cl = lane & 15;
    for (int r = warp * 2 + (lane >> 4); r < ROWS_IN * SIDE; r += 16) {
      const int yy = y0 - 1 + r / SIDE, xx = r % SIDE;
      float v = 0.f;
      if (yy >= 0 && yy < SIDE) {
        const long long o = ((long long)b * L + yy * SIDE + xx) * hid + c0 + cl;
        v = d_dt[o] * gelu_grad(dtpre[o]);
      }
      sG[cl * CSTRIDE + r] = v;                // even CSTRIDE: the two half-warps (r, r + 1) land on disjoint banks
    }
  }
  const float* hb = h1pre + (long long)b * L * hid;
  constexpr int VEC_ROW = SIDE / 4, N_VEC = DWF_CH * ROWS_IN * VEC_ROW;
  for (int i = threadIdx.x; i < N_VEC; i += 256) {
    const int xv = i % VEC_ROW, r = (i / VEC_ROW) % ROWS_IN, c = i / (VEC_ROW * ROWS_IN);
    const int yy = y0 - 1 + r;
    float gl[4] = {0.f, 0.f, 0.f, 0.f}, dg[4] = {0.f, 0.f, 0.f, 0.f};
    if (yy >= 0 && yy < SIDE) {
      const long long o = (long long)(c0 + c) * L + yy * SIDE + xv * 4;
      const float4 hv = *reinterpret_cast<const float4*>(hb + o);
      const float hx[4] = {hv.x, hv.y, hv.z, hv.w};
      const unsigned long long e = (unsigned long long)b * L * hid + o;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float sc = drop_scale(p_drop, seed, site, e + k);
        float g1, d1;
        gelu_and_grad(hx[k], g1, d1);
        gl[k] = g1 * sc;
        dg[k] = d1 * sc;
      }
    }
    float* dh = sH + c * CSTRIDE + r * SIDE + xv * 4;
    *reinterpret_cast<float2*>(dh) = make_float2(gl[0], gl[1]);
    *reinterpret_cast<float2*>(dh + 2) = make_float2(gl[2], gl[3]);
    if (r >= 1 && r <= DWT_ROWS)
      *reinterpret_cast<float4*>(sD + c * DSTRIDE + (r - 1) * SIDE + xv * 4) = make_float4(dg[0], dg[1], dg[2], dg[3]);
  }
  __syncthreads();
  // data gradient, raw (channel-plane) layout: consecutive threads -> consecutive x
#pragma unroll 2
  for (int i = threadIdx.x; i < DWF_CH * DSTRIDE; i += 256) {
    const int c = i / DSTRIDE, r = i % DSTRIDE;
    const int yl = r / SIDE, xx = r % SIDE;
    const float* g = sG + c * CSTRIDE;
    const float* wk = w + (c0 + c) * 9;
    float acc = 0.f;
#pragma unroll
    for (int ky = 0; ky < 3; ++ky)
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        const int xs = xx - kx + 1;                  // source pixel of g: (y - ky + 1, x - kx + 1)
        if (xs >= 0 && xs < SIDE) acc = fmaf(g[(yl + 2 - ky) * SIDE + xs], __ldg(wk + ky * 3 + kx), acc);
      }
    d_h1pre[(long long)b * L * hid + (long long)(c0 + c) * L + (y0 + yl) * SIDE + xx] = acc * sD[c * DSTRIDE + r];
  }
  // weight / bias gradient: warp w owns channels 2w, 2w+1; lanes split the 8 x SIDE positions
#pragma unroll
  for (int cc = 0; cc < 2; ++cc) {
    const int c = warp * 2 + cc;
    const float* g = sG + c * CSTRIDE + SIDE;          // row y0
    const float* hh = sH + c * CSTRIDE;
    float acc[9], sb = 0.f;
#pragma unroll
    for (int t = 0; t < 9; ++t) acc[t] = 0.f;
#pragma unroll 2
    for (int r = lane; r < DSTRIDE; r += 32) {
      const int yl = r / SIDE, xx = r % SIDE;
      const float gv = g[r];
      sb += gv;
#pragma unroll
      for (int ky = 0; ky < 3; ++ky)
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
          const int xs = xx + kx - 1;
          if (xs >= 0 && xs < SIDE) acc[ky * 3 + kx] = fmaf(gv, hh[(yl + ky) * SIDE + xs], acc[ky * 3 + kx]);
        }
    }
#pragma unroll
    for (int t = 0; t < 9; ++t) {
      const float v = warp_sum(acc[t]);
      if (lane == 0) atomicAdd(dw + (c0 + c) * 9 + t, v);
    }
    sb = warp_sum(sb);
    if (lane == 0) atomicAdd(db + c0 + c, sb);
  }
}

int launch_dwconv_bwd(const float* d_dt, const float* dtpre, const float* h1pre, const float* w, float* d_h1pre,
                      float* dw, float* db, int B, int L, int hid, float p_drop, unsigned long long seed, uint32_t site,
                      cudaStream_t st) {
  const int side = (int)(sqrtf((float)L) + 0.5f);
  if (side * side != L || hid % 32 || side % DWT_ROWS) return -2;
  if (side == 32 && hid % DWF_CH == 0 && (reinterpret_cast<uintptr_t>(h1pre) & 15) == 0) {
    constexpr int fast_smem = (2 * DWF_CH * ((DWT_ROWS + 2) * 32 + 2) + DWF_CH * DWT_ROWS * 32) * (int)sizeof(float);
    static PerDeviceOnce attr;
    DPMN_CUDA_TRY(attr.smem_attr(dwconv_bwd_fast_kernel<32>, fast_smem));
    dwconv_bwd_fast_kernel<32><<<dim3(32 / DWT_ROWS, hid / DWF_CH, B), 256, fast_smem, st>>>(d_dt, dtpre, h1pre, w, d_h1pre, dw,
                                                                                            db, hid, p_drop, seed, site);
    DPMN_LAUNCH_CHECK();
    return 0;
  }
  const size_t smem = (size_t)2 * 32 * ((DWT_ROWS + 2) * side + 1) * sizeof(float);
  if (smem > 200 * 1024) return -2;
  if (smem > 48 * 1024)
    DPMN_CUDA_TRY(cudaFuncSetAttribute(dwconv_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dwconv_bwd_kernel<<<dim3(side / DWT_ROWS, hid / 32, B), 256, smem, st>>>(d_dt, dtpre, h1pre, w, d_h1pre, dw, db, L, hid,
                                                                          side, p_drop, seed, site);
  DPMN_LAUNCH_CHECK();
  return 0;
}

// =====================================================================================================
// Patch embed backward (pgrm.py:419-426 [+ prior_fusion :547-548]): tokens = LN(conv_{k=s=2}(x)).
// One warp per token; the conv (12 MACs per channel) and, for the query stream, the 3x3 prior fusion are
// recomputed from the image.  Outputs: d pe_w/pe_b/ln_w/ln_b (atomics) and d x3 (the 3-channel conv input).
// =====================================================================================================
template <int PB_CPL>       // channels per lane: C <= 32 * PB_CPL
__global__ void __launch_bounds__(256) patch_embed_bwd_kernel(
    const float* __restrict__ x, long long x_bs, int in_ch, const float* __restrict__ fuse_w,
    const float* __restrict__ fuse_b, const float* __restrict__ pe_w, const float* __restrict__ pe_b,
    const float* __restrict__ ln_w, const float* __restrict__ d_tok, float* __restrict__ d_pe_w,
    float* __restrict__ d_pe_b, float* __restrict__ d_ln_w, float* __restrict__ d_ln_b, float* __restrict__ dx3,
    int B, int img_h, int img_w, int C) {
  extern __shared__ float red[];          // [8 warps][C * 15]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int gh = img_h / 2, gw = img_w / 2, L = gh * gw;
  const long long plane = (long long)img_h * img_w;
  float wv[PB_CPL][12], aw[PB_CPL][12], bv[PB_CPL], gv[PB_CPL], ab[PB_CPL], agw[PB_CPL], agb[PB_CPL];
#pragma unroll
  for (int i = 0; i < PB_CPL; ++i) {
    const int c = lane + 32 * i;
#pragma unroll
    for (int j = 0; j < 12; ++j) { wv[i][j] = c < C ? pe_w[c * 12 + j] : 0.f; aw[i][j] = 0.f; }
    bv[i] = c < C ? pe_b[c] : 0.f;
    gv[i] = c < C ? ln_w[c] : 0.f;
    ab[i] = 0.f; agw[i] = 0.f; agb[i] = 0.f;
  }
  const float invC = 1.0f / (float)C;
  const int total = B * L;
  for (int t = blockIdx.x * 8 + warp; t < total; t += gridDim.x * 8) {
    const int b = t / L, p = t - b * L;
    const int ty = p / gw, tx = p - ty * gw;
    // the 12 conv inputs (ci, dy, dx) of this token: lane j < 12 produces value j, then broadcast
    float mine = 0.f;
    if (lane < 12) {
      const int ci = lane >> 2, dy = (lane >> 1) & 1, dx = lane & 1;
      const int yy = 2 * ty + dy, xx = 2 * tx + dx;
      if (in_ch == 3) {
        mine = x[(long long)b * x_bs + ci * plane + (long long)yy * img_w + xx];
      } else {                                         // prior_fusion: conv3x3 (2 -> 3), pad 1
        float s = fuse_b[ci];
        for (int cj = 0; cj < 2; ++cj)
          for (int ky = 0; ky < 3; ++ky) {
            const int y2 = yy + ky - 1;
            if (y2 < 0 || y2 >= img_h) continue;
            for (int kx = 0; kx < 3; ++kx) {
              const int x2 = xx + kx - 1;
              if (x2 < 0 || x2 >= img_w) continue;
              s = fmaf(fuse_w[((ci * 2 + cj) * 3 + ky) * 3 + kx], x[(long long)b * x_bs + cj * plane + (long long)y2 * img_w + x2], s);
            }
          }
        mine = s;
      }
    }
    float xin[12];
#pragma unroll
    for (int j = 0; j < 12; ++j) xin[j] = __shfl_sync(0xffffffffu, mine, j);
    float pre[PB_CPL], s = 0.f;
#pragma unroll
    for (int i = 0; i < PB_CPL; ++i) {
      float a = bv[i];
#pragma unroll
      for (int j = 0; j < 12; ++j) a = fmaf(wv[i][j], xin[j], a);
      pre[i] = (lane + 32 * i) < C ? a : 0.f;
      s += pre[i];
    }
    const float mean = warp_sum(s) * invC;
    float v = 0.f;
#pragma unroll
    for (int i = 0; i < PB_CPL; ++i) {
      pre[i] = (lane + 32 * i) < C ? pre[i] - mean : 0.f;
      v = fmaf(pre[i], pre[i], v);
    }
    const float rstd = rsqrtf(warp_sum(v) * invC + 1e-5f);
    float dyv[PB_CPL], m1 = 0.f, m2 = 0.f;
#pragma unroll
    for (int i = 0; i < PB_CPL; ++i) {
      const int c = lane + 32 * i;
      pre[i] *= rstd;                                   // xhat
      const float d = c < C ? d_tok[(long long)t * C + c] : 0.f;
      agw[i] = fmaf(d, pre[i], agw[i]);
      agb[i] += d;
      dyv[i] = d * gv[i];
      m1 += dyv[i];
      m2 = fmaf(dyv[i], pre[i], m2);
    }
    m1 = warp_sum(m1) * invC;
    m2 = warp_sum(m2) * invC;
    float dxin[12];
#pragma unroll
    for (int j = 0; j < 12; ++j) dxin[j] = 0.f;
#pragma unroll
    for (int i = 0; i < PB_CPL; ++i) {
      const float dp = (lane + 32 * i) < C ? rstd * (dyv[i] - m1 - pre[i] * m2) : 0.f;   // d conv output
      ab[i] += dp;
#pragma unroll
      for (int j = 0; j < 12; ++j) {
        aw[i][j] = fmaf(dp, xin[j], aw[i][j]);
        dxin[j] = fmaf(dp, wv[i][j], dxin[j]);
      }
    }
    if (dx3 != nullptr) {
      float out = 0.f;
#pragma unroll
      for (int j = 0; j < 12; ++j) {
        const float r = warp_sum(dxin[j]);
        if (lane == j) out = r;
      }
      if (lane < 12) {
        const int ci = lane >> 2, dy = (lane >> 1) & 1, dx = lane & 1;
        dx3[(long long)b * 3 * plane + ci * plane + (long long)(2 * ty + dy) * img_w + 2 * tx + dx] = out;
      }
    }
  }
  // block reduction of the parameter accumulators, then one atomicAdd per parameter per CTA
  const int per = C * 15;
  float* mineR = red + warp * per;
#pragma unroll
  for (int i = 0; i < PB_CPL; ++i) {
    const int c = lane + 32 * i;
    if (c < C) {
#pragma unroll
      for (int j = 0; j < 12; ++j) mineR[c * 15 + j] = aw[i][j];
      mineR[c * 15 + 12] = ab[i];
      mineR[c * 15 + 13] = agw[i];
      mineR[c * 15 + 14] = agb[i];
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < per; i += 256) {
    float s = 0.f;
#pragma unroll
    for (int wq = 0; wq < 8; ++wq) s += red[wq * per + i];
    const int c = i / 15, j = i - c * 15;
    if (j < 12) atomicAdd(d_pe_w + c * 12 + j, s);
    else if (j == 12) atomicAdd(d_pe_b + c, s);
    else if (j == 13) atomicAdd(d_ln_w + c, s);
    else atomicAdd(d_ln_b + c, s);
  }
}

int launch_patch_embed_bwd(const float* x, long long x_bs, int in_ch, const float* fuse_w, const float* fuse_b,
                           const float* pe_w, const float* pe_b, const float* ln_w, const float* d_tok, float* d_pe_w,
                           float* d_pe_b, float* d_ln_w, float* d_ln_b, float* dx3, int B, int img_h, int img_w,
                           int patch, int C, cudaStream_t st) {
  if (patch != 2 || C > 256 || (in_ch != 2 && in_ch != 3)) return -2;
  const size_t smem = (size_t)8 * C * 15 * sizeof(float);
  if (smem > 200 * 1024) return -2;
  const int total = B * (img_h / 2) * (img_w / 2);
  int blocks = (total + 63) / 64;
  if (blocks > 296) blocks = 296;
  auto kern = C <= 96 ? patch_embed_bwd_kernel<3> : C <= 192 ? patch_embed_bwd_kernel<6> : patch_embed_bwd_kernel<8>;
  if (smem > 48 * 1024) DPMN_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<blocks, 256, smem, st>>>(x, x_bs, in_ch, fuse_w, fuse_b, pe_w, pe_b, ln_w, d_tok, d_pe_w, d_pe_b, d_ln_w, d_ln_b,
                                  dx3, B, img_h, img_w, C);
  DPMN_LAUNCH_CHECK();
  return 0;
}

// prior_fusion weight gradient: dW[co][ci][ky][kx] += sum dx3[b,co,y,x] * xq[b,ci,y+ky-1,x+kx-1]; db[co] += sum dx3
__global__ void __launch_bounds__(256) prior_fusion_wgrad_kernel(const float* __restrict__ dx3, const float* __restrict__ xq,
                                                                 long long xq_bs, float* __restrict__ dw,
                                                                 float* __restrict__ db, int B, int img_h, int img_w) {
  __shared__ float red[57];
  if (threadIdx.x < 57) red[threadIdx.x] = 0.f;
  __syncthreads();
  const long long plane = (long long)img_h * img_w;
  const long long total = (long long)B * plane;
  float acc[57];
#pragma unroll
  for (int i = 0; i < 57; ++i) acc[i] = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(i / plane);
    const int p = (int)(i - b * plane);
    const int yy = p / img_w, xx = p - yy * img_w;
    float g[3];
#pragma unroll
    for (int co = 0; co < 3; ++co) { g[co] = dx3[(long long)b * 3 * plane + co * plane + p]; acc[54 + co] += g[co]; }
#pragma unroll
    for (int ci = 0; ci < 2; ++ci)
#pragma unroll
      for (int ky = 0; ky < 3; ++ky)
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
          const int y2 = yy + ky - 1, x2 = xx + kx - 1;
          const float v = (y2 >= 0 && y2 < img_h && x2 >= 0 && x2 < img_w)
                              ? xq[(long long)b * xq_bs + ci * plane + (long long)y2 * img_w + x2] : 0.f;
#pragma unroll
          for (int co = 0; co < 3; ++co) acc[((co * 2 + ci) * 3 + ky) * 3 + kx] = fmaf(g[co], v, acc[((co * 2 + ci) * 3 + ky) * 3 + kx]);
        }
  }
#pragma unroll
  for (int i = 0; i < 57; ++i) {
    const float v = warp_sum(acc[i]);
    if ((threadIdx.x & 31) == 0) atomicAdd(&red[i], v);
  }
  __syncthreads();
  if (threadIdx.x < 54) atomicAdd(dw + threadIdx.x, red[threadIdx.x]);
  else if (threadIdx.x < 57) atomicAdd(db + threadIdx.x - 54, red[threadIdx.x]);
}

int launch_prior_fusion_wgrad(const float* dx3, const float* xq, long long xq_bs, float* dw, float* db, int B,
                              int img_h, int img_w, cudaStream_t st) {
  prior_fusion_wgrad_kernel<<<148, 256, 0, st>>>(dx3, xq, xq_bs, dw, db, B, img_h, img_w);
  DPMN_LAUNCH_CHECK();
  return 0;
}

// =====================================================================================================
// Head backward (pgrm.py:559-564): out = PixelShuffle(LeakyReLU(conv2(conv1(tokens)))) * w0 + sum_i res_i * w_i
// t1 (B, gh, gw, 16) is conv1's output (channels 12..15 zero).
// =====================================================================================================
constexpr int HB_PAD = 16;

// one thread per (image, low-res pixel, conv2 channel j): recompute conv2, LeakyReLU', shuffle; d w0 by atomics
__global__ void __launch_bounds__(128) head_mix_bwd_kernel(const float* __restrict__ t1, const float* __restrict__ w2,
                                                           const float* __restrict__ b2, const float* __restrict__ d_out,
                                                           const float* __restrict__ w0, float* __restrict__ d_w0,
                                                           float* __restrict__ dt2, int B, int gh, int gw, int hs, int r) {
  __shared__ float sw[9 * HB_PAD * HB_PAD];   // [tap][ci][co]
  const int hp = hs * r * r;
  for (int i = threadIdx.x; i < 9 * HB_PAD * HB_PAD; i += blockDim.x) {
    const int co = i % HB_PAD, ci = (i / HB_PAD) % HB_PAD, tap = i / (HB_PAD * HB_PAD);
    sw[i] = (co < hp && ci < hp) ? w2[(co * hp + ci) * 9 + tap] : 0.f;
  }
  __syncthreads();
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)B * gh * gw * HB_PAD;
  if (idx >= total) return;
  const int j = (int)(idx % HB_PAD);
  const long long pixb = idx / HB_PAD;
  const int xx = (int)(pixb % gw), yy = (int)((pixb / gw) % gh), b = (int)(pixb / ((long long)gw * gh));
  if (j >= hp) { dt2[idx] = 0.f; return; }
  float acc = b2[j];
  for (int ky = 0; ky < 3; ++ky) {
    const int y2 = yy + ky - 1;
    if (y2 < 0 || y2 >= gh) continue;
    for (int kx = 0; kx < 3; ++kx) {
      const int x2 = xx + kx - 1;
      if (x2 < 0 || x2 >= gw) continue;
      const float* src = t1 + ((long long)(b * gh + y2) * gw + x2) * HB_PAD;
      const float* wt = sw + (ky * 3 + kx) * HB_PAD * HB_PAD + j;
#pragma unroll
      for (int ci = 0; ci < HB_PAD; ++ci) acc = fmaf(src[ci], wt[ci * HB_PAD], acc);
    }
  }
  const int img_w = gw * r;
  const long long plane = (long long)gh * r * img_w;
  const int c = j / (r * r), rem = j - c * r * r, dy = rem / r, dx = rem - dy * r;
  const long long pix = (long long)c * plane + (long long)(yy * r + dy) * img_w + (xx * r + dx);
  const float g = d_out[(long long)b * hs * plane + pix];
  const float v = acc >= 0.f ? acc : 0.01f * acc;
  atomicAdd(d_w0 + pix, g * v);
  dt2[idx] = g * w0[pix] * (acc >= 0.f ? 1.0f : 0.01f);
}

// residual terms: d w_i[pix] += sum_b d_out * res_i ; d res_i = d_out * w_i
__global__ void mix_res_bwd_kernel(const float* __restrict__ d_out, MixBwdArgs m, int B, long long chw) {
  const long long pix = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (pix >= chw) return;
  for (int i = 1; i < m.n_mix; ++i) {
    const float wv = m.w[i][pix];
    float s = 0.f;
    for (int b = 0; b < B; ++b) {
      const float g = d_out[(long long)b * chw + pix];
      s = fmaf(g, m.in[i][(long long)b * m.in_bs[i] + pix], s);
      if (m.d_in[i] != nullptr) m.d_in[i][(long long)b * chw + pix] = g * wv;
    }
    if (m.d_w[i] != nullptr) atomicAdd(m.d_w[i] + pix, s);
  }
}

// dgrad of a 3x3 conv on NHWC with <= 16 output channels (padded rows of 16): dx[pix][ci] = sum dy[pix - tap][co] w[co][ci][tap]
// one input channel per thread: the general form (Cin % 4 != 0)
__global__ void __launch_bounds__(256) conv3x3_small_dgrad1_kernel(const float* __restrict__ dy, const float* __restrict__ w,
                                                                  float* __restrict__ dx, int B, int gh, int gw, int Cin,
                                                                  int Cout, int ldx) {
  extern __shared__ float sw[];           // [tap][co(16)][Cin]
  for (int i = threadIdx.x; i < 9 * HB_PAD * Cin; i += blockDim.x) {
    const int ci = i % Cin, co = (i / Cin) % HB_PAD, tap = i / (Cin * HB_PAD);
    sw[i] = co < Cout ? w[(co * Cin + ci) * 9 + tap] : 0.f;
  }
  __syncthreads();
  const long long total = (long long)B * gh * gw * Cin;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int ci = (int)(idx % Cin);
    const long long pixb = idx / Cin;
    const int xx = (int)(pixb % gw), yy = (int)((pixb / gw) % gh), b = (int)(pixb / ((long long)gw * gh));
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;      // four independent chains (one serial chain of 144 FMAs was latency-bound)
    for (int ky = 0; ky < 3; ++ky) {
      const int y2 = yy - ky + 1;
      if (y2 < 0 || y2 >= gh) continue;
      for (int kx = 0; kx < 3; ++kx) {
        const int x2 = xx - kx + 1;
        if (x2 < 0 || x2 >= gw) continue;
        const float4* src = reinterpret_cast<const float4*>(dy + ((long long)(b * gh + y2) * gw + x2) * HB_PAD);
        const float* wt = sw + (ky * 3 + kx) * HB_PAD * Cin + ci;
#pragma unroll
        for (int q = 0; q < HB_PAD / 4; ++q) {
          const float4 g = __ldg(src + q);
          a0 = fmaf(g.x, wt[(4 * q + 0) * Cin], a0);
          a1 = fmaf(g.y, wt[(4 * q + 1) * Cin], a1);
          a2 = fmaf(g.z, wt[(4 * q + 2) * Cin], a2);
          a3 = fmaf(g.w, wt[(4 * q + 3) * Cin], a3);
        }
      }
    }
    dx[pixb * ldx + ci] = (a0 + a1) + (a2 + a3);
  }
}

constexpr int HW_CO_PAD4 = 3;      // dy columns 0..11 carry the hp <= 12 head channels (columns 12..15 are zero padding)
__global__ void __launch_bounds__(256) conv3x3_small_dgrad_kernel(const float* __restrict__ dy, const float* __restrict__ w,
                                                                  float* __restrict__ dx, int B, int gh, int gw, int Cin,
                                                                  int Cout, int ldx) {
  extern __shared__ float sw[];           // [tap][co(16)][Cin]
  for (int i = threadIdx.x; i < 9 * HB_PAD * Cin; i += blockDim.x) {
    const int ci = i % Cin, co = (i / Cin) % HB_PAD, tap = i / (Cin * HB_PAD);
    sw[i] = co < Cout ? w[(co * Cin + ci) * 9 + tap] : 0.f;
  }
  __syncthreads();
  // four input channels per thread: one 16-byte weight read and one broadcast dy value feed four multiply-adds (the
  // one-channel form issued one shared-memory load per multiply-add: 113 us for conv1 at batch 48); Cin % 4 == 0
  const int Cq = Cin >> 2;
  const long long total = (long long)B * gh * gw * Cq;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int ci = (int)(idx % Cq) * 4;
    const long long pixb = idx / Cq;
    const int xx = (int)(pixb % gw), yy = (int)((pixb / gw) % gh), b = (int)(pixb / ((long long)gw * gh));
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f), acc2 = make_float4(0.f, 0.f, 0.f, 0.f);   // two chains per channel
    for (int ky = 0; ky < 3; ++ky) {
      const int y2 = yy - ky + 1;
      if (y2 < 0 || y2 >= gh) continue;
      for (int kx = 0; kx < 3; ++kx) {
        const int x2 = xx - kx + 1;
        if (x2 < 0 || x2 >= gw) continue;
        const float4* src = reinterpret_cast<const float4*>(dy + ((long long)(b * gh + y2) * gw + x2) * HB_PAD);
        const float* wt = sw + (ky * 3 + kx) * HB_PAD * Cin + ci;
#pragma unroll
        for (int q = 0; q < HW_CO_PAD4; ++q) {          // output channels 4 q .. 4 q + 3 (columns >= Cout of dy are zero)
          const float4 g = __ldg(src + q);
          const float4 w0 = *reinterpret_cast<const float4*>(wt + (4 * q + 0) * Cin);
          const float4 w1 = *reinterpret_cast<const float4*>(wt + (4 * q + 1) * Cin);
          const float4 w2 = *reinterpret_cast<const float4*>(wt + (4 * q + 2) * Cin);
          const float4 w3 = *reinterpret_cast<const float4*>(wt + (4 * q + 3) * Cin);
          acc.x = fmaf(g.x, w0.x, acc.x); acc.y = fmaf(g.x, w0.y, acc.y); acc.z = fmaf(g.x, w0.z, acc.z); acc.w = fmaf(g.x, w0.w, acc.w);
          acc2.x = fmaf(g.y, w1.x, acc2.x); acc2.y = fmaf(g.y, w1.y, acc2.y); acc2.z = fmaf(g.y, w1.z, acc2.z); acc2.w = fmaf(g.y, w1.w, acc2.w);
          acc.x = fmaf(g.z, w2.x, acc.x); acc.y = fmaf(g.z, w2.y, acc.y); acc.z = fmaf(g.z, w2.z, acc.z); acc.w = fmaf(g.z, w2.w, acc.w);
          acc2.x = fmaf(g.w, w3.x, acc2.x); acc2.y = fmaf(g.w, w3.y, acc2.y); acc2.z = fmaf(g.w, w3.z, acc2.z); acc2.w = fmaf(g.w, w3.w, acc2.w);
        }
      }
    }
    *reinterpret_cast<float4*>(dx + pixb * ldx + ci) = make_float4(acc.x + acc2.x, acc.y + acc2.y, acc.z + acc2.z, acc.w + acc2.w);
  }
}

// wgrad of the same conv: thread = input channel ci, registers hold dW[co < 12][tap]; a CTA covers `chunk` pixels.
// x (B, gh, gw, ldx) NHWC, dy (B, gh, gw, 16).  Also db[co] += sum dy.
constexpr int HW_CO = 12;
__global__ void __launch_bounds__(128) conv3x3_small_wgrad_kernel(const float* __restrict__ dy, const float* __restrict__ x,
                                                                  float* __restrict__ dw, float* __restrict__ db, int B,
                                                                  int gh, int gw, int Cin, int Cout, int ldx, int chunk) {
  const int ci = blockIdx.y * blockDim.x + threadIdx.x;
  const long long total = (long long)B * gh * gw;
  const long long p0 = (long long)blockIdx.x * chunk;
  const long long p1 = min(total, p0 + chunk);
  float acc[HW_CO][9];
#pragma unroll
  for (int co = 0; co < HW_CO; ++co)
#pragma unroll
    for (int t = 0; t < 9; ++t) acc[co][t] = 0.f;
  float sb = 0.f;
  if (ci < Cin) {
    for (long long p = p0; p < p1; ++p) {
      const int xx = (int)(p % gw), yy = (int)((p / gw) % gh);
      float g[HW_CO];
      const float* gp = dy + p * HB_PAD;
#pragma unroll
      for (int co = 0; co < HW_CO; ++co) g[co] = gp[co];
#pragma unroll
      for (int ky = 0; ky < 3; ++ky)
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
          const int y2 = yy + ky - 1, x2 = xx + kx - 1;
          if (y2 < 0 || y2 >= gh || x2 < 0 || x2 >= gw) continue;
          const float v = x[(p + (long long)(ky - 1) * gw + (kx - 1)) * ldx + ci];
#pragma unroll
          for (int co = 0; co < HW_CO; ++co) acc[co][ky * 3 + kx] = fmaf(g[co], v, acc[co][ky * 3 + kx]);
        }
      if (ci < Cout) sb += gp[ci];
    }
#pragma unroll
    for (int co = 0; co < HW_CO; ++co) {
      if (co >= Cout) break;
#pragma unroll
      for (int t = 0; t < 9; ++t) atomicAdd(dw + (co * Cin + ci) * 9 + t, acc[co][t]);
    }
    if (ci < Cout && blockIdx.y == 0) atomicAdd(db + ci, sb);
  }
}

static void launch_small_wgrad(const float* dy, const float* x, float* dw, float* db, int B, int gh, int gw, int Cin, int Cout,
                               int ldx, long long pix, cudaStream_t st) {
  const int chunk = 128;
  const unsigned blocks = (unsigned)((pix + chunk - 1) / chunk);
  if (Cin <= 32) conv3x3_small_wgrad_kernel<<<dim3(blocks, 1), 32, 0, st>>>(dy, x, dw, db, B, gh, gw, Cin, Cout, ldx, chunk);
  else conv3x3_small_wgrad_kernel<<<dim3(blocks, (Cin + 127) / 128), 128, 0, st>>>(dy, x, dw, db, B, gh, gw, Cin, Cout, ldx, chunk);
}

int launch_head_bwd(const HeadBwdArgs& h, cudaStream_t st) {
  const int hp = h.hs * h.patch * h.patch;
  if (hp > HW_CO || h.C % 4) return -2;
  const long long pix = (long long)h.B * h.gh * h.gw;
  const long long chw = (long long)h.hs * h.gh * h.patch * h.gw * h.patch;
  // (1) mix + shuffle + LeakyReLU backward -> dt2, d weight_list_0
  head_mix_bwd_kernel<<<(unsigned)((pix * HB_PAD + 127) / 128), 128, 0, st>>>(h.t1, h.w2, h.b2, h.d_out, h.mix.w[0], h.mix.d_w[0],
                                                                             h.dt2, h.B, h.gh, h.gw, h.hs, h.patch);
  DPMN_LAUNCH_CHECK();
  if (h.mix.n_mix > 1) {
    mix_res_bwd_kernel<<<(unsigned)((chw + 255) / 256), 256, 0, st>>>(h.d_out, h.mix, h.B, chw);
    DPMN_LAUNCH_CHECK();
  }
  // (2) conv2 (hp -> hp) backward: weights / bias, then data -> dt1 (B, gh, gw, 16)
  launch_small_wgrad(h.dt2, h.t1, h.d_w2, h.d_b2, h.B, h.gh, h.gw, hp, hp, HB_PAD, pix, st);
  DPMN_LAUNCH_CHECK();
  {
    const size_t smem = (size_t)9 * HB_PAD * hp * sizeof(float);   // dt1 columns hp..15 stay zero (memset by the caller)
    if (hp % 4 == 0) conv3x3_small_dgrad_kernel<<<592, 256, smem, st>>>(h.dt2, h.w2, h.dt1, h.B, h.gh, h.gw, hp, hp, HB_PAD);
    else conv3x3_small_dgrad1_kernel<<<592, 256, smem, st>>>(h.dt2, h.w2, h.dt1, h.B, h.gh, h.gw, hp, hp, HB_PAD);
    DPMN_LAUNCH_CHECK();
  }
  // (3) conv1 (C -> hp) backward: weights / bias, then data -> d tokens (B, L, C)
  launch_small_wgrad(h.dt1, h.tokens, h.d_w1, h.d_b1, h.B, h.gh, h.gw, h.C, hp, h.C, pix, st);
  DPMN_LAUNCH_CHECK();
  {
    const size_t smem = (size_t)9 * HB_PAD * h.C * sizeof(float);
    if (smem > 200 * 1024) return -2;
    static size_t attr = 0;
    if (smem > 48 * 1024 && smem > attr) {
      DPMN_CUDA_TRY(cudaFuncSetAttribute(conv3x3_small_dgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      attr = smem;
    }
    conv3x3_small_dgrad_kernel<<<592, 256, smem, st>>>(h.dt1, h.w1, h.d_tokens, h.B, h.gh, h.gw, h.C, hp, h.C);
    DPMN_LAUNCH_CHECK();
  }
  return 0;
}

// =====================================================================================================================
// Training forward, 16-bit modes: fp32 token-order q (rows, C) / kv (rows, 2C) -> the window-major 16-bit operands
// [G][rows][cg] of the tcgen05 attention kernel (roll + window_partition, pgrm.py:209-225).  One thread per 8 channels
// of one window-major row: 3 x 32-byte reads, 3 x 16-byte writes.
// =====================================================================================================================
struct ScatterGeom { int ws[4], shift[4]; };

template <typename T>
__global__ void __launch_bounds__(256) window_scatter16_kernel(const float* __restrict__ q, const float* __restrict__ kv,
                                                               T* __restrict__ qw, T* __restrict__ kw, T* __restrict__ vw,
                                                               long long rows, int L, int H, int W, int C, int G,
                                                               ScatterGeom geo) {
  const int cg = C / G, c8 = cg / 8;
  const long long total = (long long)G * rows * c8;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int ch = (int)(i % c8);
    const long long rp = i / c8;
    const long long p = rp % rows;
    const int g = (int)(rp / rows);
    const long long b = p / L;
    const int p_img = (int)(p - b * L);
    const int tok = window_row_to_token(p_img, H, W, geo.ws[g], geo.shift[g]).token;
    const long long src_row = b * L + tok;
    const int c = g * cg + ch * 8;
    const float4* sq = reinterpret_cast<const float4*>(q + src_row * C + c);
    const float4* sk = reinterpret_cast<const float4*>(kv + src_row * 2 * C + c);
    const float4* sv = reinterpret_cast<const float4*>(kv + src_row * 2 * C + C + c);
    const float4 a0 = sq[0], a1 = sq[1], k0 = sk[0], k1 = sk[1], v0 = sv[0], v1 = sv[1];
    auto pack = [](const float4& x, const float4& y) {
      union { uint4 u; T h[8]; } o;
      o.h[0] = from_f32<T>(x.x); o.h[1] = from_f32<T>(x.y); o.h[2] = from_f32<T>(x.z); o.h[3] = from_f32<T>(x.w);
      o.h[4] = from_f32<T>(y.x); o.h[5] = from_f32<T>(y.y); o.h[6] = from_f32<T>(y.z); o.h[7] = from_f32<T>(y.w);
      return o.u;
    };
    const long long dst = ((long long)g * rows + p) * cg + ch * 8;
    *reinterpret_cast<uint4*>(qw + dst) = pack(a0, a1);
    *reinterpret_cast<uint4*>(kw + dst) = pack(k0, k1);
    *reinterpret_cast<uint4*>(vw + dst) = pack(v0, v1);
  }
}

int launch_window_scatter16(const float* q, const float* kv, void* qw, void* kw, void* vw, DType t, int B, int H, int W, int C,
                            int G, const int* ws, const int* shift, cudaStream_t st) {
  if (t != DT_F16 && t != DT_BF16) return -2;
  if (G < 1 || G > 4 || C % G || (C / G) % 8) return -2;
  ScatterGeom geo;
  for (int g = 0; g < 4; ++g) { geo.ws[g] = g < G ? ws[g] : 1; geo.shift[g] = g < G ? shift[g] : 0; }
  const int L = H * W;
  const long long rows = (long long)B * L;
  const long long total = (long long)G * rows * (C / G / 8);
  const int blocks = (int)((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16);
  if (t == DT_F16)
    window_scatter16_kernel<__half><<<blocks, 256, 0, st>>>(q, kv, (__half*)qw, (__half*)kw, (__half*)vw, rows, L, H, W, C, G, geo);
  else
    window_scatter16_kernel<__nv_bfloat16><<<blocks, 256, 0, st>>>(q, kv, (__nv_bfloat16*)qw, (__nv_bfloat16*)kw,
                                                                   (__nv_bfloat16*)vw, rows, L, H, W, C, G, geo);
  DPMN_LAUNCH_CHECK();
  return 0;
}

}  // namespace dpmn
