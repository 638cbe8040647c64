// SIMT (CUDA-core, fp32-arithmetic) kernels of the PGRM hot path.  These are the exact-arithmetic mode
// (DPMN_PREC_F32: 1e-5 parity with the reference) and the non-GEMM stages of the tensor-core modes.
//
// Reference behaviour restated (never copied) from /root/reference/model/pgrm.py; line numbers per kernel.
#include "common.cuh"
#include "kernels.h"

namespace dpmn {

// =====================================================================================================
// K0  prior_fusion (optional) + PatchEmbed conv(k = s = patch) + LayerNorm        pgrm.py:419-426,547-550
// one warp per token; lane j < in3*p*p holds one (channel, dy, dx) input of the patch conv, every lane
// owns C/32 output channels.
// =====================================================================================================
template <int CPL>   // channels per lane = C / 32
__global__ void __launch_bounds__(256) patch_embed_kernel(
    const float* __restrict__ x, long long x_bs, int in_ch, const float* __restrict__ fuse_w,
    const float* __restrict__ fuse_b, const float* __restrict__ pe_w, const float* __restrict__ pe_b,
    const float* __restrict__ ln_w, const float* __restrict__ ln_b, float* __restrict__ tokens, int B, int img_h,
    int img_w, int patch, int n_tokens_total, PatchEmbedLn extra) {
  constexpr int C = CPL * 32;
  extern __shared__ float smem[];
  const int pk = 3 * patch * patch;              // inputs of the patch conv per token (<= 32)
  float* s_w = smem;                             // [pk][C]  transposed patch-embed weight
  float* s_fw = s_w + pk * C;                    // [3*2*9]  prior_fusion weight
  for (int i = threadIdx.x; i < pk * C; i += blockDim.x) {
    const int k = i / C, c = i - k * C;
    s_w[i] = pe_w[c * pk + k];
  }
  if (fuse_w != nullptr)
    for (int i = threadIdx.x; i < 54; i += blockDim.x) s_fw[i] = fuse_w[i];
  __syncthreads();

  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int n_warps = (gridDim.x * blockDim.x) >> 5;
  const int gw = img_w / patch, gh = img_h / patch;
  const int L = gh * gw;
  const long long plane = (long long)img_h * img_w;

  float bias_r[CPL], lw_r[CPL], lb_r[CPL];
#pragma unroll
  for (int i = 0; i < CPL; ++i) {
    bias_r[i] = pe_b[lane + 32 * i];
    lw_r[i] = ln_w[lane + 32 * i];
    lb_r[i] = ln_b[lane + 32 * i];
  }

  for (int tok = warp; tok < n_tokens_total; tok += n_warps) {
    const int b = tok / L, t = tok - b * L;
    const int ty = t / gw, tx = t - ty * gw;
    // lane j -> (ch, dy, dx) of the conv input
    float v = 0.f;
    if (lane < pk) {
      const int ch = lane / (patch * patch);
      const int r = lane - ch * patch * patch;
      const int dy = r / patch, dx = r - dy * patch;
      const int y = ty * patch + dy, xx = tx * patch + dx;
      const float* xb = x + (long long)b * x_bs;
      if (fuse_w == nullptr) {
        v = xb[ch * plane + (long long)y * img_w + xx];
      } else {
        // prior_fusion: conv3x3 pad 1, in_ch (2) -> 3, at full resolution (pgrm.py:471,547-548)
        float acc = fuse_b[ch];
        for (int ci = 0; ci < in_ch; ++ci)
#pragma unroll
          for (int ky = 0; ky < 3; ++ky) {
            const int yy = y + ky - 1;
            if (yy < 0 || yy >= img_h) continue;
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) {
              const int xc = xx + kx - 1;
              if (xc < 0 || xc >= img_w) continue;
              acc = fmaf(xb[ci * plane + (long long)yy * img_w + xc], s_fw[(ch * in_ch + ci) * 9 + ky * 3 + kx], acc);
            }
          }
        v = acc;
      }
    }
    float o[CPL];
#pragma unroll
    for (int i = 0; i < CPL; ++i) o[i] = bias_r[i];
    for (int k = 0; k < pk; ++k) {
      const float xv = __shfl_sync(0xffffffffu, v, k);
#pragma unroll
      for (int i = 0; i < CPL; ++i) o[i] = fmaf(xv, s_w[k * C + lane + 32 * i], o[i]);
    }
    // LayerNorm over C (biased variance, eps 1e-5)
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < CPL; ++i) s += o[i];
    const float mu = warp_sum(s) * (1.0f / C);
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < CPL; ++i) { const float d = o[i] - mu; q = fmaf(d, d, q); }
    const float rstd = rsqrtf(warp_sum(q) * (1.0f / C) + 1e-5f);
    float tv[CPL];
#pragma unroll
    for (int i = 0; i < CPL; ++i) tv[i] = (o[i] - mu) * rstd * lw_r[i] + lb_r[i];
    if (tokens != nullptr) {
      float* dst = tokens + (long long)tok * C;
#pragma unroll
      for (int i = 0; i < CPL; ++i) dst[lane + 32 * i] = tv[i];
    }
    if (extra.count > 0) {
      // the consumers' own LayerNorms (norm1_q of both blocks / norm1_kv of block 0, pgrm.py:322-323) of the token
      float s2 = 0.f;
#pragma unroll
      for (int i = 0; i < CPL; ++i) s2 += tv[i];
      const float mu2 = warp_sum(s2) * (1.0f / C);
      float q2 = 0.f;
#pragma unroll
      for (int i = 0; i < CPL; ++i) { const float d = tv[i] - mu2; q2 = fmaf(d, d, q2); }
      const float rstd2 = rsqrtf(warp_sum(q2) * (1.0f / C) + 1e-5f);
      for (int k = 0; k < extra.count; ++k) {
#pragma unroll
        for (int i = 0; i < CPL; ++i) {
          const int c = lane + 32 * i;
          const float y = (tv[i] - mu2) * rstd2 * extra.w[k][c] + extra.b[k][c];
          const long long idx = (long long)tok * C + c;
          if (extra.type == DT_F32) reinterpret_cast<float*>(extra.out[k])[idx] = y;
          else if (extra.type == DT_F16) reinterpret_cast<__half*>(extra.out[k])[idx] = __float2half_rn(y);
          else reinterpret_cast<__nv_bfloat16*>(extra.out[k])[idx] = __float2bfloat16_rn(y);
        }
      }
    }
  }
}

// Thread-per-token variant (C = 96, patch 2): all C accumulators of a token live in one thread, so both
// LayerNorm levels are register-local (no shuffles) and the weights are smem broadcasts.  ~10x fewer issue slots
// than the warp-per-token kernel above, which remains the generic fallback.
template <int C>
__global__ void __launch_bounds__(128) patch_embed_tok_kernel(
    const float* __restrict__ x, long long x_bs, int in_ch, const float* __restrict__ fuse_w,
    const float* __restrict__ fuse_b, const float* __restrict__ pe_w, const float* __restrict__ pe_b,
    const float* __restrict__ ln_w, const float* __restrict__ ln_b, float* __restrict__ tokens, int B, int img_h,
    int img_w, int n_tokens_total, PatchEmbedLn extra) {
  __shared__ __align__(16) float s_w[12 * C];      // [k][c]
  __shared__ __align__(16) float s_aff[6 * C];     // pe bias, ln w, ln b, + up to... (extra affines read from global)
  __shared__ float s_fw[54 + 3];
  for (int i = threadIdx.x; i < 12 * C; i += blockDim.x) {
    const int k = i / C, c = i - k * C;
    s_w[i] = pe_w[c * 12 + k];
  }
  for (int i = threadIdx.x; i < C; i += blockDim.x) { s_aff[i] = pe_b[i]; s_aff[C + i] = ln_w[i]; s_aff[2 * C + i] = ln_b[i]; }
  if (fuse_w != nullptr)
    for (int i = threadIdx.x; i < 57; i += blockDim.x) s_fw[i] = i < 54 ? fuse_w[i] : fuse_b[i - 54];
  __syncthreads();
  const int tok = blockIdx.x * blockDim.x + threadIdx.x;
  if (tok >= n_tokens_total) return;
  const int gw = img_w / 2, gh = img_h / 2, L = gh * gw;
  const int b = tok / L, t = tok - b * L;
  const int ty = t / gw, tx = t - ty * gw;
  const long long plane = (long long)img_h * img_w;
  const float* xb = x + (long long)b * x_bs;
  float in[12];   // (ch, dy, dx)
  if (fuse_w == nullptr) {
#pragma unroll
    for (int ch = 0; ch < 3; ++ch)
#pragma unroll
      for (int dy = 0; dy < 2; ++dy) {
        const float2 v = *reinterpret_cast<const float2*>(xb + ch * plane + (long long)(2 * ty + dy) * img_w + 2 * tx);
        in[ch * 4 + dy * 2] = v.x; in[ch * 4 + dy * 2 + 1] = v.y;
      }
  } else {
    // prior_fusion conv3x3 pad 1 (in_ch -> 3) at the 2x2 pixels of this token: a 4x4 input patch per channel
    float patch[2][4][4];
    for (int ci = 0; ci < 2; ++ci)
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const int yy = 2 * ty - 1 + r;
#pragma unroll
        for (int cidx = 0; cidx < 4; ++cidx) {
          const int xx = 2 * tx - 1 + cidx;
          patch[ci][r][cidx] = (ci < in_ch && yy >= 0 && yy < img_h && xx >= 0 && xx < img_w)
                                   ? xb[ci * plane + (long long)yy * img_w + xx] : 0.f;
        }
      }
#pragma unroll
    for (int ch = 0; ch < 3; ++ch)
#pragma unroll
      for (int dy = 0; dy < 2; ++dy)
#pragma unroll
        for (int dx = 0; dx < 2; ++dx) {
          float acc = s_fw[54 + ch];
          for (int ci = 0; ci < 2; ++ci)
#pragma unroll
            for (int ky = 0; ky < 3; ++ky)
#pragma unroll
              for (int kx = 0; kx < 3; ++kx)
                acc = fmaf(patch[ci][dy + ky][dx + kx], s_fw[(ch * in_ch + ci) * 9 + ky * 3 + kx], acc);
          in[ch * 4 + dy * 2 + dx] = acc;
        }
  }
  float o[C];
#pragma unroll
  for (int c = 0; c < C; ++c) o[c] = s_aff[c];
#pragma unroll
  for (int k = 0; k < 12; ++k) {
#pragma unroll
    for (int c = 0; c < C; c += 4) {
      const float4 w4 = *reinterpret_cast<const float4*>(&s_w[k * C + c]);
      o[c] = fmaf(in[k], w4.x, o[c]); o[c + 1] = fmaf(in[k], w4.y, o[c + 1]);
      o[c + 2] = fmaf(in[k], w4.z, o[c + 2]); o[c + 3] = fmaf(in[k], w4.w, o[c + 3]);
    }
  }
  auto layer_norm_inplace = [&](const float* w, const float* bsrc, bool from_smem) {
    float s1 = 0.f;
#pragma unroll
    for (int c = 0; c < C; ++c) s1 += o[c];
    const float mu = s1 * (1.0f / C);
    float s2 = 0.f;
#pragma unroll
    for (int c = 0; c < C; ++c) { const float dlt = o[c] - mu; s2 = fmaf(dlt, dlt, s2); }
    const float rstd = rsqrtf(s2 * (1.0f / C) + 1e-5f);
    (void)from_smem;
#pragma unroll
    for (int c = 0; c < C; ++c) o[c] = (o[c] - mu) * rstd * w[c] + bsrc[c];
  };
  layer_norm_inplace(s_aff + C, s_aff + 2 * C, true);          // patch_embed.norm
  if (tokens != nullptr) {
    float* dst = tokens + (long long)tok * C;
#pragma unroll
    for (int c = 0; c < C; c += 4) *reinterpret_cast<float4*>(dst + c) = make_float4(o[c], o[c + 1], o[c + 2], o[c + 3]);
  }
  if (extra.count > 0) {
    float s1 = 0.f;
#pragma unroll
    for (int c = 0; c < C; ++c) s1 += o[c];
    const float mu = s1 * (1.0f / C);
    float s2 = 0.f;
#pragma unroll
    for (int c = 0; c < C; ++c) { const float dlt = o[c] - mu; s2 = fmaf(dlt, dlt, s2); }
    const float rstd = rsqrtf(s2 * (1.0f / C) + 1e-5f);
    for (int k = 0; k < extra.count; ++k) {
      const float* w = extra.w[k];
      const float* bb = extra.b[k];
      const long long base = (long long)tok * C;
#pragma unroll
      for (int c = 0; c < C; c += 8) {
        float y[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) y[e] = (o[c + e] - mu) * rstd * __ldg(w + c + e) + __ldg(bb + c + e);
        if (extra.type == DT_F32) {
          float* dst = reinterpret_cast<float*>(extra.out[k]) + base + c;
          *reinterpret_cast<float4*>(dst) = make_float4(y[0], y[1], y[2], y[3]);
          *reinterpret_cast<float4*>(dst + 4) = make_float4(y[4], y[5], y[6], y[7]);
        } else if (extra.type == DT_F16) {
          union { uint4 u; __half h[8]; } pk;
#pragma unroll
          for (int e = 0; e < 8; ++e) pk.h[e] = __float2half_rn(y[e]);
          *reinterpret_cast<uint4*>(reinterpret_cast<__half*>(extra.out[k]) + base + c) = pk.u;
        } else {
          union { uint4 u; __nv_bfloat16 h[8]; } pk;
#pragma unroll
          for (int e = 0; e < 8; ++e) pk.h[e] = __float2bfloat16_rn(y[e]);
          *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(extra.out[k]) + base + c) = pk.u;
        }
      }
    }
  }
}

// Four lanes per token (C = 96, patch 2): lane `part` owns channels 32 j + 8 part + e (j < 3, e < 8), so every 16-bit
// output leaves as one 16-byte store per (lane, j) and the four lanes of a token cover a 128-byte line of the fp32 row.
// 4x the threads of the thread-per-token kernel at ~1/3 of its registers (occupancy was 16 % there); LayerNorm
// statistics cost two shuffles per sum; with prior fusion each lane computes one of the token's 2x2 fused pixels.
template <int C>
__global__ void __launch_bounds__(256) patch_embed_tok4_kernel(
    const float* __restrict__ x, long long x_bs, int in_ch, const float* __restrict__ fuse_w,
    const float* __restrict__ fuse_b, const float* __restrict__ pe_w, const float* __restrict__ pe_b,
    const float* __restrict__ ln_w, const float* __restrict__ ln_b, float* __restrict__ tokens, int B, int img_h,
    int img_w, int n_tokens_total, PatchEmbedLn extra) {
  static_assert(C == 96, "channel mapping assumes 3 groups of 32");
  __shared__ __align__(16) float s_w[12 * C];      // [k][c]
  __shared__ __align__(16) float s_aff[3 * C];     // pe bias, ln w, ln b
  __shared__ float s_fw[54 + 3];
  for (int i = threadIdx.x; i < 12 * C; i += blockDim.x) {
    const int k = i / C, c = i - k * C;
    s_w[i] = pe_w[c * 12 + k];
  }
  for (int i = threadIdx.x; i < C; i += blockDim.x) { s_aff[i] = pe_b[i]; s_aff[C + i] = ln_w[i]; s_aff[2 * C + i] = ln_b[i]; }
  if (fuse_w != nullptr)
    for (int i = threadIdx.x; i < 57; i += blockDim.x) s_fw[i] = i < 54 ? fuse_w[i] : fuse_b[i - 54];
  __syncthreads();
  const int gtid = blockIdx.x * blockDim.x + threadIdx.x;
  const int part = gtid & 3;
  const bool live = (gtid >> 2) < n_tokens_total;
  const int tok = live ? (gtid >> 2) : n_tokens_total - 1;      // dead lanes shadow the last token (shuffles stay full-warp)
  const int gw = img_w / 2, gh = img_h / 2, L = gh * gw;
  const int b = tok / L, t = tok - b * L;
  const int ty = t / gw, tx = t - ty * gw;
  const long long plane = (long long)img_h * img_w;
  const float* xb = x + (long long)b * x_bs;
  float in[12];   // (ch, dy, dx)
  if (fuse_w == nullptr) {
#pragma unroll
    for (int ch = 0; ch < 3; ++ch)
#pragma unroll
      for (int dy = 0; dy < 2; ++dy) {
        const float2 v = *reinterpret_cast<const float2*>(xb + ch * plane + (long long)(2 * ty + dy) * img_w + 2 * tx);
        in[ch * 4 + dy * 2] = v.x; in[ch * 4 + dy * 2 + 1] = v.y;
      }
  } else {
    // prior_fusion conv3x3 pad 1 (in_ch -> 3): this lane's pixel of the token is (2 ty + part / 2, 2 tx + part % 2)
    const int py = 2 * ty + (part >> 1), px = 2 * tx + (part & 1);
    float mine[3] = {s_fw[54], s_fw[55], s_fw[56]};
    for (int ci = 0; ci < 2; ++ci) {
      if (ci >= in_ch) break;
#pragma unroll
      for (int ky = 0; ky < 3; ++ky) {
        const int yy = py + ky - 1;
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
          const int xx = px + kx - 1;
          const float v = (yy >= 0 && yy < img_h && xx >= 0 && xx < img_w) ? xb[ci * plane + (long long)yy * img_w + xx] : 0.f;
#pragma unroll
          for (int ch = 0; ch < 3; ++ch) mine[ch] = fmaf(v, s_fw[(ch * in_ch + ci) * 9 + ky * 3 + kx], mine[ch]);
        }
      }
    }
    const int base_lane = (threadIdx.x & 31) & ~3;
#pragma unroll
    for (int ch = 0; ch < 3; ++ch)
#pragma unroll
      for (int q = 0; q < 4; ++q) in[ch * 4 + q] = __shfl_sync(0xffffffffu, mine[ch], base_lane + q);
  }
  float o[24];
#pragma unroll
  for (int j = 0; j < 3; ++j)
#pragma unroll
    for (int e = 0; e < 8; ++e) o[j * 8 + e] = s_aff[32 * j + 8 * part + e];
#pragma unroll
  for (int k = 0; k < 12; ++k) {
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const float4 wa = *reinterpret_cast<const float4*>(&s_w[k * C + 32 * j + 8 * part]);
      const float4 wb = *reinterpret_cast<const float4*>(&s_w[k * C + 32 * j + 8 * part + 4]);
      float* oo = o + j * 8;
      oo[0] = fmaf(in[k], wa.x, oo[0]); oo[1] = fmaf(in[k], wa.y, oo[1]);
      oo[2] = fmaf(in[k], wa.z, oo[2]); oo[3] = fmaf(in[k], wa.w, oo[3]);
      oo[4] = fmaf(in[k], wb.x, oo[4]); oo[5] = fmaf(in[k], wb.y, oo[5]);
      oo[6] = fmaf(in[k], wb.z, oo[6]); oo[7] = fmaf(in[k], wb.w, oo[7]);
    }
  }
  auto stats = [&](float& mu, float& rstd) {
    float s1 = 0.f;
#pragma unroll
    for (int c = 0; c < 24; ++c) s1 += o[c];
    s1 += __shfl_xor_sync(0xffffffffu, s1, 1);
    s1 += __shfl_xor_sync(0xffffffffu, s1, 2);
    mu = s1 * (1.0f / C);
    float s2 = 0.f;
#pragma unroll
    for (int c = 0; c < 24; ++c) { const float dlt = o[c] - mu; s2 = fmaf(dlt, dlt, s2); }
    s2 += __shfl_xor_sync(0xffffffffu, s2, 1);
    s2 += __shfl_xor_sync(0xffffffffu, s2, 2);
    rstd = rsqrtf(s2 * (1.0f / C) + 1e-5f);
  };
  float mu, rstd;
  stats(mu, rstd);                                              // patch_embed.norm
#pragma unroll
  for (int j = 0; j < 3; ++j)
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int c = 32 * j + 8 * part + e;
      o[j * 8 + e] = (o[j * 8 + e] - mu) * rstd * s_aff[C + c] + s_aff[2 * C + c];
    }
  if (tokens != nullptr && live) {
    float* dst = tokens + (long long)tok * C + 8 * part;
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      *reinterpret_cast<float4*>(dst + 32 * j) = make_float4(o[j * 8], o[j * 8 + 1], o[j * 8 + 2], o[j * 8 + 3]);
      *reinterpret_cast<float4*>(dst + 32 * j + 4) = make_float4(o[j * 8 + 4], o[j * 8 + 5], o[j * 8 + 6], o[j * 8 + 7]);
    }
  }
  if (extra.count > 0) {
    stats(mu, rstd);                                            // the consumers' own LayerNorms of the token
#pragma unroll
    for (int k = 0; k < 2; ++k) {                               // static indices: `extra` stays in the constant bank
      if (k >= extra.count) break;
      const float* w = extra.w[k];
      const float* bb = extra.b[k];
      const long long base = (long long)tok * C + 8 * part;
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        const float4 w0 = __ldg(reinterpret_cast<const float4*>(w + 32 * j + 8 * part));
        const float4 w1 = __ldg(reinterpret_cast<const float4*>(w + 32 * j + 8 * part + 4));
        const float4 b0 = __ldg(reinterpret_cast<const float4*>(bb + 32 * j + 8 * part));
        const float4 b1 = __ldg(reinterpret_cast<const float4*>(bb + 32 * j + 8 * part + 4));
        const float wv[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
        const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
        float y[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) y[e] = (o[j * 8 + e] - mu) * rstd * wv[e] + bv[e];
        if (!live) continue;
        if (extra.type == DT_F32) {
          float* dst = reinterpret_cast<float*>(extra.out[k]) + base + 32 * j;
          *reinterpret_cast<float4*>(dst) = make_float4(y[0], y[1], y[2], y[3]);
          *reinterpret_cast<float4*>(dst + 4) = make_float4(y[4], y[5], y[6], y[7]);
        } else if (extra.type == DT_F16) {
          union { uint4 u; __half h[8]; } pk;
#pragma unroll
          for (int e = 0; e < 8; ++e) pk.h[e] = __float2half_rn(y[e]);
          *reinterpret_cast<uint4*>(reinterpret_cast<__half*>(extra.out[k]) + base + 32 * j) = pk.u;
        } else {
          union { uint4 u; __nv_bfloat16 h[8]; } pk;
#pragma unroll
          for (int e = 0; e < 8; ++e) pk.h[e] = __float2bfloat16_rn(y[e]);
          *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(extra.out[k]) + base + 32 * j) = pk.u;
        }
      }
    }
  }
}

int launch_patch_embed(const float* x, long long x_bs, int in_ch, const float* fuse_w, const float* fuse_b,
                       const float* pe_w, const float* pe_b, const float* ln_w, const float* ln_b, float* tokens,
                       int B, int img_h, int img_w, int patch, int C, cudaStream_t st, const PatchEmbedLn* extra_in) {
  if (C % 32 != 0 || C > 256 || 3 * patch * patch > 32) return -2;
  PatchEmbedLn extra;
  if (extra_in) extra = *extra_in;
  if (C == 96 && patch == 2 && img_w % 2 == 0 && (x_bs % 2) == 0 && (reinterpret_cast<uintptr_t>(x) & 7) == 0 &&
      (in_ch == 3 || (in_ch == 2 && fuse_w != nullptr))) {
    const int total_tok = B * (img_h / 2) * (img_w / 2);
    const bool aligned16 = (reinterpret_cast<uintptr_t>(extra.w[0]) | reinterpret_cast<uintptr_t>(extra.b[0]) |
                            reinterpret_cast<uintptr_t>(extra.w[1]) | reinterpret_cast<uintptr_t>(extra.b[1])) % 16 == 0;
    if (aligned16)
      patch_embed_tok4_kernel<96><<<(total_tok * 4 + 255) / 256, 256, 0, st>>>(x, x_bs, in_ch, fuse_w, fuse_b, pe_w, pe_b, ln_w,
                                                                               ln_b, tokens, B, img_h, img_w, total_tok, extra);
    else
      patch_embed_tok_kernel<96><<<(total_tok + 127) / 128, 128, 0, st>>>(x, x_bs, in_ch, fuse_w, fuse_b, pe_w, pe_b, ln_w,
                                                                        ln_b, tokens, B, img_h, img_w, total_tok, extra);
    DPMN_LAUNCH_CHECK();
    return 0;
  }
  const int L = (img_h / patch) * (img_w / patch);
  const int total = B * L;
  const int threads = 256;
  int blocks = (total + 7) / 8;
  if (blocks > 148 * 8) blocks = 148 * 8;
  const size_t smem = (size_t)(3 * patch * patch * C + 64) * sizeof(float);
#define DPMN_PE(CPL_)                                                                                         \
  patch_embed_kernel<CPL_><<<blocks, threads, smem, st>>>(x, x_bs, in_ch, fuse_w, fuse_b, pe_w, pe_b, ln_w,  \
                                                          ln_b, tokens, B, img_h, img_w, patch, total, extra)
  switch (C / 32) {
    case 1: DPMN_PE(1); break;
    case 2: DPMN_PE(2); break;
    case 3: DPMN_PE(3); break;
    case 4: DPMN_PE(4); break;
    case 6: DPMN_PE(6); break;
    case 8: DPMN_PE(8); break;
    default: return -2;
  }
#undef DPMN_PE
  DPMN_LAUNCH_CHECK();
  return 0;
}

// =====================================================================================================
// LayerNorm over rows of (rows, C) fp32; one warp per row.                      pgrm.py:303-304,311,322-323
// =====================================================================================================
// out2 (optional): a second copy of the normalised row in 16 bits (type2: DT_F16 / DT_BF16) -- the training forward keeps the
// fp32 row for the backward and feeds the 16-bit one to the tcgen05 GEMM without a convert pass of its own.
template <typename OutT, int CPL>
__global__ void __launch_bounds__(256) layernorm_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                        const float* __restrict__ b, OutT* __restrict__ out,
                                                        int rows, void* __restrict__ out2 = nullptr, int type2 = 0) {
  constexpr int C = CPL * 32;
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int n_warps = (gridDim.x * blockDim.x) >> 5;
  float wr[CPL], br[CPL];
#pragma unroll
  for (int i = 0; i < CPL; ++i) { wr[i] = w[lane + 32 * i]; br[i] = b[lane + 32 * i]; }
  for (int r = warp; r < rows; r += n_warps) {
    const float* src = x + (long long)r * C;
    float v[CPL];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < CPL; ++i) { v[i] = src[lane + 32 * i]; s += v[i]; }
    const float mu = warp_sum(s) * (1.0f / C);
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < CPL; ++i) { const float d = v[i] - mu; q = fmaf(d, d, q); }
    const float rstd = rsqrtf(warp_sum(q) * (1.0f / C) + 1e-5f);
    OutT* dst = out + (long long)r * C;
#pragma unroll
    for (int i = 0; i < CPL; ++i) {
      const float y = (v[i] - mu) * rstd * wr[i] + br[i];
      dst[lane + 32 * i] = from_f32<OutT>(y);
      if (out2 != nullptr) {
        if (type2 == DT_F16) reinterpret_cast<__half*>(out2)[(long long)r * C + lane + 32 * i] = __float2half_rn(y);
        else reinterpret_cast<__nv_bfloat16*>(out2)[(long long)r * C + lane + 32 * i] = __float2bfloat16_rn(y);
      }
    }
  }
}

template <typename OutT>
static int launch_layernorm_t(const float* x, const float* w, const float* b, OutT* out, int rows, int C,
                              cudaStream_t st, void* out2 = nullptr, int type2 = 0) {
  int blocks = (rows + 7) / 8;
  if (blocks > 148 * 8) blocks = 148 * 8;
  switch (C / 32) {
    case 1: layernorm_kernel<OutT, 1><<<blocks, 256, 0, st>>>(x, w, b, out, rows, out2, type2); break;
    case 2: layernorm_kernel<OutT, 2><<<blocks, 256, 0, st>>>(x, w, b, out, rows, out2, type2); break;
    case 3: layernorm_kernel<OutT, 3><<<blocks, 256, 0, st>>>(x, w, b, out, rows, out2, type2); break;
    case 4: layernorm_kernel<OutT, 4><<<blocks, 256, 0, st>>>(x, w, b, out, rows, out2, type2); break;
    case 6: layernorm_kernel<OutT, 6><<<blocks, 256, 0, st>>>(x, w, b, out, rows, out2, type2); break;
    case 8: layernorm_kernel<OutT, 8><<<blocks, 256, 0, st>>>(x, w, b, out, rows, out2, type2); break;
    default: return -2;
  }
  DPMN_LAUNCH_CHECK();
  return 0;
}

// fp32 LayerNorm output + a 16-bit copy of it in one pass (training forward of the 16-bit modes)
int launch_layernorm_dual(const float* x, const float* w, const float* b, float* out, void* out16, DType type16, int rows, int C,
                          cudaStream_t st) {
  if (C % 32 != 0 || C > 256 || (type16 != DT_F16 && type16 != DT_BF16)) return -2;
  return launch_layernorm_t<float>(x, w, b, out, rows, C, st, out16, (int)type16);
}

int launch_layernorm(const float* x, const float* w, const float* b, void* out, DType out_type, int rows, int C,
                     cudaStream_t st) {
  if (C % 32 != 0 || C > 256) return -2;
  switch (out_type) {
    case DT_F32: return launch_layernorm_t<float>(x, w, b, (float*)out, rows, C, st);
    case DT_F16: return launch_layernorm_t<__half>(x, w, b, (__half*)out, rows, C, st);
    case DT_BF16: return launch_layernorm_t<__nv_bfloat16>(x, w, b, (__nv_bfloat16*)out, rows, C, st);
  }
  return -1;
}

// =====================================================================================================
// fp32 SIMT GEMM  C[z][m,n] = epi(sum_k A[z][m,k] * B[z][n,k])   (both operands K-contiguous)
// 64x64 tile, BK 16, 256 threads, 4x4 micro-tile.  Requires K % 4 == 0, lda/ldb % 4 == 0, N % 4 == 0,
// ldc % 4 == 0 and 16-byte aligned bases.
// =====================================================================================================
constexpr int GM = 64, GN = 64, GK = 16, GPAD = 68;

__global__ void __launch_bounds__(256) gemm_simt_kernel(GemmSimtArgs p) {
  __shared__ __align__(16) float As[GK][GPAD];
  __shared__ __align__(16) float Bs[GK][GPAD];
  __shared__ float red[16][GN];

  const int z = blockIdx.z;
  const int m0 = blockIdx.y * GM, n0 = blockIdx.x * GN;
  const float* A = p.A + (long long)z * p.a_bs;
  const float* Bm = p.Bm + (long long)z * p.b_bs;
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int lrow = tid >> 2, lk = (tid & 3) * 4;   // loader: row within tile, k offset

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  const bool a_ok = (m0 + lrow) < p.M, b_ok = (n0 + lrow) < p.N;
  const float* a_ptr = A + (long long)(m0 + lrow) * p.lda + lk;
  const float* b_ptr = Bm + (long long)(n0 + lrow) * p.ldb + lk;

  for (int k0 = 0; k0 < p.K; k0 += GK) {
    float4 av = make_float4(0.f, 0.f, 0.f, 0.f), bv = av;
    if (a_ok && (k0 + lk) < p.K) av = *reinterpret_cast<const float4*>(a_ptr + k0);
    if (b_ok && (k0 + lk) < p.K) bv = *reinterpret_cast<const float4*>(b_ptr + k0);
    __syncthreads();   // previous tile fully consumed
    As[lk + 0][lrow] = av.x; As[lk + 1][lrow] = av.y; As[lk + 2][lrow] = av.z; As[lk + 3][lrow] = av.w;
    Bs[lk + 0][lrow] = bv.x; Bs[lk + 1][lrow] = bv.y; Bs[lk + 2][lrow] = bv.z; Bs[lk + 3][lrow] = bv.w;
    __syncthreads();
#pragma unroll
    for (int k = 0; k < GK; ++k) {
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

  // epilogue
  const int n = n0 + tx * 4;
  const float* bias = p.bias ? p.bias + (long long)z * p.bias_bs : nullptr;
  float colpart[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= p.M || n >= p.N) continue;
    float v[4] = {acc[i][0], acc[i][1], acc[i][2], acc[i][3]};
    if (p.bias_mode == 1) {
      const float4 bb = *reinterpret_cast<const float4*>(bias + n);
      v[0] += bb.x; v[1] += bb.y; v[2] += bb.z; v[3] += bb.w;
    } else if (p.bias_mode == 2) {
      const float bb = bias[m];
      v[0] += bb; v[1] += bb; v[2] += bb; v[3] += bb;
    }
    if (p.act == 1) {
#pragma unroll
      for (int j = 0; j < 4; ++j) v[j] = gelu_erf(v[j]);
    }
    if (p.colsum != nullptr) {
#pragma unroll
      for (int j = 0; j < 4; ++j) colpart[j] += v[j];
      continue;
    }
    const long long off = (long long)z * p.c_bs + (long long)m * p.ldc + n;
    if (p.residual != nullptr) {
      const float4 rr = *reinterpret_cast<const float4*>(p.residual + off);
      v[0] += rr.x; v[1] += rr.y; v[2] += rr.z; v[3] += rr.w;
    }
    *reinterpret_cast<float4*>(p.C + off) = make_float4(v[0], v[1], v[2], v[3]);
  }
  if (p.colsum != nullptr) {
    // deterministic in-tile reduction over the 16 row-groups
#pragma unroll
    for (int j = 0; j < 4; ++j) red[ty][tx * 4 + j] = colpart[j];
    __syncthreads();
    if (tid < GN) {
      float s = 0.f;
#pragma unroll
      for (int r = 0; r < 16; ++r) s += red[r][tid];
      const int nn = n0 + tid;
      if (nn < p.N) p.colsum[((long long)z * gridDim.y + blockIdx.y) * p.N + nn] = s;
    }
  }
}

int launch_gemm_simt(const GemmSimtArgs& a, cudaStream_t st) {
  if (a.K % 4 || a.lda % 4 || a.ldb % 4 || a.N % 4 || (a.colsum == nullptr && a.ldc % 4)) return -2;
  dim3 grid((a.N + GN - 1) / GN, (a.M + GM - 1) / GM, a.batch);
  gemm_simt_kernel<<<grid, 256, 0, st>>>(a);
  DPMN_LAUNCH_CHECK();
  return 0;
}

// =====================================================================================================
// K2  windowed multi-head attention core, SIMT                                        pgrm.py:197-268
// One thread per (window-major row, head).  A CTA covers T = max(128, N) rows = T/N whole windows of
// one (image, group, head); K and V rows of those windows sit in shared memory (fp32).
//   S[n,m] = 0.25*<q_n,k_m> + table[idx(n,m), head] + (label_n != label_m ? -100 : 0);  P = softmax_m(S)
//   out[row p = w*N + n] = sum_m P[n,m] v_m       (window-major rows are kept: quirk 1)
// =====================================================================================================
template <typename T, int D>
__global__ void window_attn_simt_kernel(const T* __restrict__ q, const T* __restrict__ kv, T* __restrict__ out,
                                        int q_ld, int kv_ld, int v_off, int out_ld,
                                        const float* __restrict__ table, int hpg, int ch0, int H, int W, int ws,
                                        int shift, float scale, float p_drop, unsigned long long seed, uint32_t site,
                                        int g_index, int G) {
  extern __shared__ __align__(16) float sm[];
  const int N = ws * ws;
  const int L = H * W;
  const int Tn = blockDim.x;
  const int units = Tn / N;                  // windows per CTA
  const int ustride = N * D + 4;             // +4 floats: units of one warp land in different banks
  float* sK = sm;                            // [units][N][D] (+pad)
  float* sV = sK + units * ustride;
  float* sT = sV + units * ustride;          // [(2ws-1)^2] bias column of this head
  int* sLab = reinterpret_cast<int*>(sT + (2 * ws - 1) * (2 * ws - 1));   // [Tn] shift-mask labels

  const int head = blockIdx.y;
  const int b = blockIdx.z;
  const int row0 = blockIdx.x * Tn;          // first window-major row of this CTA
  const int t = threadIdx.x;
  const int p = row0 + t;
  const int u = t / N, n = t - u * N;
  const int ch = ch0 + head * D;

  for (int i = t; i < (2 * ws - 1) * (2 * ws - 1); i += Tn) sT[i] = table[i * hpg + head];

  const WinCoord wc = window_row_to_token(p, H, W, ws, shift);
  sLab[t] = shift > 0 ? shift_region_label(wc.hp, wc.wp, H, W, ws, shift) : 0;
  const long long tok = (long long)b * L + wc.token;
  float qr[D];
  {
    const T* qp = q + tok * q_ld + ch;
    const T* kp = kv + tok * kv_ld + ch;
    const T* vp = kp + v_off;
    float* dk = sK + u * ustride + n * D;
    float* dv = sV + u * ustride + n * D;
#pragma unroll
    for (int e = 0; e < D; ++e) {
      qr[e] = to_f32<T>(qp[e]) * scale;      // q * scale first (pgrm.py:230-231)
      dk[e] = to_f32<T>(kp[e]);
      dv[e] = to_f32<T>(vp[e]);
    }
  }
  __syncthreads();

  const int i_n = n / ws, j_n = n - i_n * ws;
  const int my_lab = sLab[t];
  const float* kbase = sK + u * ustride;
  const float* vbase = sV + u * ustride;
  const int* lab = sLab + u * N;
  const int tw = 2 * ws - 1;

  float mx = -INFINITY, den = 0.f;
  float o[D];
#pragma unroll
  for (int e = 0; e < D; ++e) o[e] = 0.f;
  int i_m = 0, j_m = 0;
  for (int m = 0; m < N; ++m) {
    const float4* k4 = reinterpret_cast<const float4*>(kbase + m * D);
    float s = 0.f;
#pragma unroll
    for (int e = 0; e < D / 4; ++e) {
      const float4 kk = k4[e];
      s = fmaf(qr[4 * e + 0], kk.x, s);
      s = fmaf(qr[4 * e + 1], kk.y, s);
      s = fmaf(qr[4 * e + 2], kk.z, s);
      s = fmaf(qr[4 * e + 3], kk.w, s);
    }
    s += sT[(i_n - i_m + ws - 1) * tw + (j_n - j_m + ws - 1)];
    if (lab[m] != my_lab) s += -100.0f;      // pgrm.py:173
    const float nmx = fmaxf(mx, s);
    const float corr = expf(mx - nmx);       // 0 on the first key (mx = -inf)
    float pexp = expf(s - nmx);
    den = den * corr + pexp;
    // attn_drop (pgrm.py:248, train mode): the normalised probability of key m is kept with 1/(1-p) or zeroed
    if (p_drop > 0.f)
      pexp *= drop_scale(p_drop, seed, site, ((((unsigned long long)b * G + g_index) * hpg + head) * L + p) * N + m);
    const float4* v4 = reinterpret_cast<const float4*>(vbase + m * D);
#pragma unroll
    for (int e = 0; e < D / 4; ++e) {
      const float4 vv = v4[e];
      o[4 * e + 0] = fmaf(pexp, vv.x, o[4 * e + 0] * corr);
      o[4 * e + 1] = fmaf(pexp, vv.y, o[4 * e + 1] * corr);
      o[4 * e + 2] = fmaf(pexp, vv.z, o[4 * e + 2] * corr);
      o[4 * e + 3] = fmaf(pexp, vv.w, o[4 * e + 3] * corr);
    }
    mx = nmx;
    if (++j_m == ws) { j_m = 0; ++i_m; }
  }
  const float inv = 1.0f / den;
  T* op = out + ((long long)b * L + p) * out_ld + ch;
#pragma unroll
  for (int e = 0; e < D; ++e) op[e] = from_f32<T>(o[e] * inv);
}

template <typename T, int D>
static int launch_attn_group(const AttnArgs& a, int g, cudaStream_t st) {
  const int ws = a.window[g], N = ws * ws, L = a.H * a.W;
  const int cg = a.C / a.n_groups;
  const int Tn = N > 128 ? N : 128;
  if (Tn > 1024 || L % Tn != 0 || Tn % N != 0) return -2;
  const int units = Tn / N;
  const size_t smem = (size_t)(2 * units * (N * D + 4) + (2 * ws - 1) * (2 * ws - 1)) * sizeof(float) +
                      (size_t)Tn * sizeof(int);
  auto kern = window_attn_simt_kernel<T, D>;
  if (smem > 48 * 1024) DPMN_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid(L / Tn, a.heads_per_group, a.B);
  kern<<<grid, Tn, smem, st>>>((const T*)a.q, (const T*)a.kv, (T*)a.out, a.q_ld, a.kv_ld, a.v_off, a.out_ld,
                               a.table[g], a.heads_per_group, g * cg, a.H, a.W, ws, a.shift[g],
                               1.0f / sqrtf((float)D), a.p_drop, a.seed, a.site, g, a.n_groups);
  DPMN_LAUNCH_CHECK();
  return 0;
}

template <typename T>
static int launch_attn_t(const AttnArgs& a, cudaStream_t st) {
  const int cg = a.C / a.n_groups;
  const int d = cg / a.heads_per_group;
  for (int g = 0; g < a.n_groups; ++g) {
    int rc;
    switch (d) {
      case 8: rc = launch_attn_group<T, 8>(a, g, st); break;
      case 16: rc = launch_attn_group<T, 16>(a, g, st); break;
      case 24: rc = launch_attn_group<T, 24>(a, g, st); break;
      case 32: rc = launch_attn_group<T, 32>(a, g, st); break;
      case 48: rc = launch_attn_group<T, 48>(a, g, st); break;
      case 64: rc = launch_attn_group<T, 64>(a, g, st); break;
      default: rc = -2;
    }
    if (rc != 0) return rc;
  }
  return 0;
}

int launch_window_attn_simt(const AttnArgs& a, cudaStream_t st) {
  if (a.n_groups < 1 || a.n_groups > 4 || a.C % a.n_groups || (a.C / a.n_groups) % a.heads_per_group) return -1;
  for (int g = 0; g < a.n_groups; ++g)
    if (a.window[g] < 1 || a.H % a.window[g] || a.W % a.window[g]) return -2;
  switch (a.io_type) {
    case DT_F32: return launch_attn_t<float>(a, st);
    case DT_F16: return launch_attn_t<__half>(a, st);
    case DT_BF16: return launch_attn_t<__nv_bfloat16>(a, st);
  }
  return -1;
}

// =====================================================================================================
// K3  SK gate folding                                                                  pgrm.py:84-95
// S = mean_L GELU(proj(x));  Z = GELU(fc1 S);  A = softmax_G(view(fc2 Z, (G, cg)));
// out = proj(x) + proj_head(sum_m A[m] * x_m)  ==  x * (Wp + Wh diag(A))^T + (bp + bh)
// One CTA per image builds the folded (C x C) weight and bias.
// =====================================================================================================
template <typename WT>
__global__ void __launch_bounds__(256) sk_gate_kernel(const float* __restrict__ colsum, int tiles, int L,
                                                      const float* __restrict__ wp, const float* __restrict__ bp,
                                                      const float* __restrict__ w1, const float* __restrict__ b1,
                                                      const float* __restrict__ w2, const float* __restrict__ b2,
                                                      const float* __restrict__ wh, const float* __restrict__ bh,
                                                      WT* __restrict__ wb_out, float* __restrict__ bias_out, int C,
                                                      int G) {
  extern __shared__ float sm[];
  const int cg = C / G, dz = cg / 2;
  float* sS = sm;          // [C]
  float* sZ = sS + C;      // [dz]
  float* sA = sZ + dz;     // [C]  attention vector, index m*cg + j
  const int b = blockIdx.x, t = threadIdx.x;
  for (int c = t; c < C; c += blockDim.x) {
    float s = 0.f;
#pragma unroll 8
    for (int i = 0; i < tiles; ++i) s += colsum[((long long)b * tiles + i) * C + c];
    sS[c] = s / (float)L;
  }
  __syncthreads();
  {
    const int lane = t & 31, warp = t >> 5, nw = blockDim.x >> 5;
    for (int j = warp; j < dz; j += nw) {               // one warp per output: coalesced, parallel loads
      float s = 0.f;
      for (int c = lane; c < C; c += 32) s = fmaf(w1[j * C + c], sS[c], s);
      s = warp_sum(s);
      if (lane == 0) sZ[j] = gelu_erf(s + b1[j]);
    }
  }
  __syncthreads();
  for (int c = t; c < C; c += blockDim.x) {
    float s = b2[c];
#pragma unroll 8
    for (int j = 0; j < dz; ++j) s = fmaf(w2[c * dz + j], sZ[j], s);
    sA[c] = s;
  }
  __syncthreads();
  for (int j = t; j < cg; j += blockDim.x) {   // softmax over the G groups for channel j
    float mx = -INFINITY;
    for (int m = 0; m < G; ++m) mx = fmaxf(mx, sA[m * cg + j]);
    float den = 0.f;
    for (int m = 0; m < G; ++m) den += expf(sA[m * cg + j] - mx);
    for (int m = 0; m < G; ++m) sA[m * cg + j] = expf(sA[m * cg + j] - mx) / den;
  }
  __syncthreads();
  // each of the gridDim.y CTAs of an image writes its slice of the folded weight (the tiny gate above is recomputed)
  const int per = (C * C + gridDim.y - 1) / gridDim.y;
  const int i_end = min(C * C, (int)(blockIdx.y + 1) * per);
#pragma unroll 4
  for (int i = blockIdx.y * per + t; i < i_end; i += blockDim.x) {
    const int o = i / C, c = i - o * C;
    const int j = c % cg;
    wb_out[(long long)b * C * C + i] = from_f32<WT>(fmaf(wh[o * cg + j], sA[c], wp[i]));
  }
  if (blockIdx.y == 0)
    for (int o = t; o < C; o += blockDim.x) bias_out[(long long)b * C + o] = bp[o] + bh[o];
}

// C = 96, G = 3 (cg 32, dz 16), 256 threads, 4 CTAs per image: the same computation with every parameter load issued
// before the first barrier.  The generic kernel above is a chain of five phases that each start with a global-memory
// round trip (column sums -> fc1 rows -> fc2 rows -> proj / proj_head slices): ~15 us of mostly exposed latency for a
// few kFLOP.  Here the phases only wait on shared memory.
template <typename WT>
__global__ void __launch_bounds__(256) sk_gate_c96_kernel(const float* __restrict__ colsum, int tiles, float inv_L,
                                                          const float* __restrict__ wp, const float* __restrict__ bp,
                                                          const float* __restrict__ w1, const float* __restrict__ b1,
                                                          const float* __restrict__ w2, const float* __restrict__ b2,
                                                          const float* __restrict__ wh, const float* __restrict__ bh,
                                                          WT* __restrict__ wb_out, float* __restrict__ bias_out) {
  constexpr int C = 96, CG = 32, DZ = 16, PER = C * C / 4, ITERS = PER / 256;   // 2304 folded weights per CTA, 9 per thread
  __shared__ float sS[C], sZ[DZ], sA[C];
  const int b = blockIdx.x, t = threadIdx.x, lane = t & 31, warp = t >> 5;
  asm volatile("griddepcontrol.wait;" ::: "memory");                 // PDL (tc_common.cuh): colsum is the predecessor's output
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  // ---- all global loads up front
  float cs = 0.f;
  if (t < C) {
#pragma unroll 8
    for (int i = 0; i < tiles; ++i) cs += colsum[((long long)b * tiles + i) * C + t];
  }
  float r1[2][3];
#pragma unroll
  for (int q = 0; q < 2; ++q)
#pragma unroll
    for (int k = 0; k < 3; ++k) r1[q][k] = w1[(warp + 8 * q) * C + lane + 32 * k];
  const float rb1 = b1[warp + 8 * (lane & 1)];
  float r2[DZ];
  float rb2 = 0.f;
  if (t < C) {
    const float4* src = reinterpret_cast<const float4*>(w2 + t * DZ);
#pragma unroll
    for (int q = 0; q < DZ / 4; ++q) {
      const float4 v = src[q];
      r2[4 * q] = v.x; r2[4 * q + 1] = v.y; r2[4 * q + 2] = v.z; r2[4 * q + 3] = v.w;
    }
    rb2 = b2[t];
  }
  float pw[ITERS], ph[ITERS];
#pragma unroll
  for (int k = 0; k < ITERS; ++k) {
    const int i = blockIdx.y * PER + t + 256 * k;
    const int o = i / C, c = i - o * C;
    pw[k] = wp[i];
    ph[k] = wh[o * CG + (c % CG)];
  }
  float bias_v = 0.f;
  if (blockIdx.y == 0 && t < C) bias_v = bp[t] + bh[t];
  // ---- S = mean_L GELU(proj(x))
  if (t < C) sS[t] = cs * inv_L;
  __syncthreads();
  // ---- Z = GELU(fc1 S): warp w owns outputs w and w + 8
#pragma unroll
  for (int q = 0; q < 2; ++q) {
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < 3; ++k) s = fmaf(r1[q][k], sS[lane + 32 * k], s);
    s = warp_sum(s);
    const float bq = __shfl_sync(0xffffffffu, rb1, q);
    if (lane == 0) sZ[warp + 8 * q] = gelu_erf(s + bq);
  }
  __syncthreads();
  // ---- fc2 and the softmax over the 3 groups
  if (t < C) {
    float s = rb2;
#pragma unroll
    for (int j = 0; j < DZ; ++j) s = fmaf(r2[j], sZ[j], s);
    sA[t] = s;
  }
  __syncthreads();
  if (t < CG) {
    const float a0 = sA[t], a1 = sA[CG + t], a2 = sA[2 * CG + t];
    const float mx = fmaxf(a0, fmaxf(a1, a2));
    const float e0 = expf(a0 - mx), e1 = expf(a1 - mx), e2 = expf(a2 - mx);
    const float den = e0 + e1 + e2;
    sA[t] = e0 / den; sA[CG + t] = e1 / den; sA[2 * CG + t] = e2 / den;
  }
  __syncthreads();
  // ---- folded weight slice of this CTA: Wp + Wh diag(A)
#pragma unroll
  for (int k = 0; k < ITERS; ++k) {
    const int i = blockIdx.y * PER + t + 256 * k;
    const int c = i % C;
    wb_out[(long long)b * C * C + i] = from_f32<WT>(fmaf(ph[k], sA[c], pw[k]));
  }
  if (blockIdx.y == 0 && t < C) bias_out[(long long)b * C + t] = bias_v;
}

template <typename WT>
static void launch_sk_gate_c96(const float* colsum, int tiles, int L, const float* wp, const float* bp, const float* w1,
                               const float* b1, const float* w2, const float* b2, const float* wh, const float* bh,
                               void* wb_out, float* bias_out, int B, cudaStream_t st) {
  launch_pdl(sk_gate_c96_kernel<WT>, dim3(B, 4), dim3(256), 0, st, colsum, tiles, 1.0f / (float)L, wp, bp, w1, b1, w2, b2, wh, bh,
             (WT*)wb_out, bias_out);
}

int launch_sk_gate(const float* colsum, int tiles_per_image, int L, const float* wp, const float* bp,
                   const float* w1, const float* b1, const float* w2, const float* b2, const float* wh,
                   const float* bh, void* wb_out, DType wb_type, float* bias_out, int B, int C, int G,
                   cudaStream_t st) {
  const int cg = C / G;
  if (C == 96 && G == 3 && (reinterpret_cast<uintptr_t>(w2) & 15) == 0) {
    switch (wb_type) {
      case DT_F32: launch_sk_gate_c96<float>(colsum, tiles_per_image, L, wp, bp, w1, b1, w2, b2, wh, bh, wb_out, bias_out, B, st); break;
      case DT_F16: launch_sk_gate_c96<__half>(colsum, tiles_per_image, L, wp, bp, w1, b1, w2, b2, wh, bh, wb_out, bias_out, B, st); break;
      case DT_BF16: launch_sk_gate_c96<__nv_bfloat16>(colsum, tiles_per_image, L, wp, bp, w1, b1, w2, b2, wh, bh, wb_out, bias_out, B, st); break;
    }
    DPMN_LAUNCH_CHECK();
    return 0;
  }
  const size_t smem = (size_t)(2 * C + cg / 2 + 8) * sizeof(float);
  switch (wb_type) {
    case DT_F32:
      sk_gate_kernel<float><<<dim3(B, 4), 256, smem, st>>>(colsum, tiles_per_image, L, wp, bp, w1, b1, w2, b2, wh, bh,
                                                  (float*)wb_out, bias_out, C, G);
      break;
    case DT_F16:
      sk_gate_kernel<__half><<<dim3(B, 4), 256, smem, st>>>(colsum, tiles_per_image, L, wp, bp, w1, b1, w2, b2, wh, bh,
                                                   (__half*)wb_out, bias_out, C, G);
      break;
    case DT_BF16:
      sk_gate_kernel<__nv_bfloat16><<<dim3(B, 4), 256, smem, st>>>(colsum, tiles_per_image, L, wp, bp, w1, b1, w2, b2, wh,
                                                          bh, (__nv_bfloat16*)wb_out, bias_out, C, G);
      break;
  }
  DPMN_LAUNCH_CHECK();
  return 0;
}

// =====================================================================================================
// K4a  depthwise 3x3 (+bias, GELU) on the RAW view of the hidden tensor                pgrm.py:33-36
// h is (B, L, hid) row-major; the reference reinterprets each image's L*hid floats as (hid, side, side)
// with no transpose (quirk 2).  The result is written transposed, dt[b][pixel][channel], so that the
// pointwise conv is a K-contiguous GEMM operand.
// CTA = (image, 32 channels, strip of 8 rows); lane = channel for conflict-free smem and coalesced stores.
// =====================================================================================================
constexpr int DW_ROWS = 8;

template <typename T>
__global__ void __launch_bounds__(256) dwconv_kernel(const T* __restrict__ h, T* __restrict__ dt,
                                                     const float* __restrict__ w, const float* __restrict__ bias,
                                                     int L, int hid, int side) {
  extern __shared__ float sm[];
  const int b = blockIdx.z;
  const int c0 = blockIdx.y * 32;
  const int y0 = blockIdx.x * DW_ROWS;
  const int rows_in = DW_ROWS + 2;
  const int cstride = rows_in * side + 1;
  const T* hb = h + (long long)b * L * hid;
  // stage rows y0-1 .. y0+DW_ROWS of 32 channel planes (zero outside the plane)
  for (int i = threadIdx.x; i < 32 * rows_in * side; i += blockDim.x) {
    const int c = i / (rows_in * side);
    const int r = i - c * rows_in * side;
    const int yy = y0 - 1 + r / side;
    const int xx = r % side;
    float v = 0.f;
    if (yy >= 0 && yy < side) v = to_f32<T>(hb[(long long)(c0 + c) * L + yy * side + xx]);
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
  const int n_pix = DW_ROWS * side;
  for (int pi = warp; pi < n_pix; pi += 8) {
    const int yl = pi / side, xx = pi - yl * side;
    if (y0 + yl >= side) break;
    float acc = bb;
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
      const float* row = pl + (yl + ky) * side;
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        const int xc = xx + kx - 1;
        if (xc >= 0 && xc < side) acc = fmaf(row[xc], wk[ky * 3 + kx], acc);
      }
    }
    const int pix = (y0 + yl) * side + xx;
    dt[((long long)b * L + pix) * hid + c] = from_f32<T>(gelu_erf(acc));
  }
}

// 16-bit variant: 64 channels x 8 rows per CTA, 16-byte global loads and stores, sliding 3x3 window in registers.
// Phase 1 stages the input rows (with halo) as [channel][row][x] in smem (channel stride 161 words: lanes that
// are consecutive channels hit distinct banks); phase 2 walks each (channel, row) along x; phase 3 writes the
// transposed tile [pixel][64 channels] with one 128-byte run per pixel.
template <typename T>
__global__ void __launch_bounds__(256) dwconv16_kernel(const T* __restrict__ h, T* __restrict__ dt,
                                                       const float* __restrict__ w, const float* __restrict__ bias,
                                                       int L, int hid, int side) {
  static_assert(sizeof(T) == 2, "16-bit storage");
  extern __shared__ __align__(16) unsigned char smraw[];
  const int rows_in = DW_ROWS + 2;
  const int cstride = rows_in * side + 2;            // halves; (cstride / 2) odd for side = 32
  T* s_in = reinterpret_cast<T*>(smraw);             // [64][cstride]
  T* s_out = s_in + 64 * cstride;                    // [DW_ROWS * side][64]
  const int b = blockIdx.z, c0 = blockIdx.y * 64, y0 = blockIdx.x * DW_ROWS;
  const T* hb = h + (long long)b * L * hid;
  const int vec_per_row = side / 8;
  const int n_vec = 64 * rows_in * vec_per_row;
  for (int i = threadIdx.x; i < n_vec; i += 256) {
    const int xv = i % vec_per_row;
    const int r = (i / vec_per_row) % rows_in;
    const int c = i / (vec_per_row * rows_in);
    const int yy = y0 - 1 + r;
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (yy >= 0 && yy < side) v = *reinterpret_cast<const uint4*>(hb + (long long)(c0 + c) * L + yy * side + xv * 8);
    uint32_t* dst = reinterpret_cast<uint32_t*>(s_in + c * cstride + r * side + xv * 8);
    dst[0] = v.x; dst[1] = v.y; dst[2] = v.z; dst[3] = v.w;
  }
  __syncthreads();
  {
    const int c = threadIdx.x & 63;
    float wk[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) wk[i] = w[(c0 + c) * 9 + i];
    const float bb = bias[c0 + c];
    for (int row = threadIdx.x >> 6; row < DW_ROWS; row += 4) {
      if (y0 + row >= side) break;
      const T* r0 = s_in + c * cstride + row * side;   // input rows row, row+1, row+2 (halo offset -1 applied)
      const T* r1 = r0 + side;
      const T* r2 = r1 + side;
      // sliding 3x3 window, two output columns per step (one 32-bit smem load per row per step)
      float a0 = 0.f, a1 = 0.f, a2 = 0.f;              // column x-1
      float2 c0 = to_f32x2<T>(r0), c1 = to_f32x2<T>(r1), c2 = to_f32x2<T>(r2);   // columns x, x+1
      for (int x = 0; x < side; x += 2) {
        float2 n0 = make_float2(0.f, 0.f), n1 = n0, n2 = n0;                    // columns x+2, x+3
        if (x + 2 < side) { n0 = to_f32x2<T>(r0 + x + 2); n1 = to_f32x2<T>(r1 + x + 2); n2 = to_f32x2<T>(r2 + x + 2); }
        float acc0 = bb, acc1 = bb;
        acc0 = fmaf(a0, wk[0], acc0); acc0 = fmaf(c0.x, wk[1], acc0); acc0 = fmaf(c0.y, wk[2], acc0);
        acc0 = fmaf(a1, wk[3], acc0); acc0 = fmaf(c1.x, wk[4], acc0); acc0 = fmaf(c1.y, wk[5], acc0);
        acc0 = fmaf(a2, wk[6], acc0); acc0 = fmaf(c2.x, wk[7], acc0); acc0 = fmaf(c2.y, wk[8], acc0);
        acc1 = fmaf(c0.x, wk[0], acc1); acc1 = fmaf(c0.y, wk[1], acc1); acc1 = fmaf(n0.x, wk[2], acc1);
        acc1 = fmaf(c1.x, wk[3], acc1); acc1 = fmaf(c1.y, wk[4], acc1); acc1 = fmaf(n1.x, wk[5], acc1);
        acc1 = fmaf(c2.x, wk[6], acc1); acc1 = fmaf(c2.y, wk[7], acc1); acc1 = fmaf(n2.x, wk[8], acc1);
        s_out[(row * side + x) * 64 + c] = from_f32<T>(gelu_fast(acc0));
        s_out[(row * side + x + 1) * 64 + c] = from_f32<T>(gelu_fast(acc1));
        a0 = c0.y; a1 = c1.y; a2 = c2.y;
        c0 = n0; c1 = n1; c2 = n2;
      }
    }
  }
  __syncthreads();
  const int n_out_vec = DW_ROWS * side * 8;            // 8 x 16 B per pixel
  for (int i = threadIdx.x; i < n_out_vec; i += 256) {
    const int pix_l = i >> 3, part = i & 7;
    const int yl = pix_l / side;
    if (y0 + yl >= side) continue;
    const int pix = y0 * side + pix_l;
    *reinterpret_cast<uint4*>(dt + ((long long)b * L + pix) * hid + c0 + part * 8) =
        *reinterpret_cast<const uint4*>(s_out + pix_l * 64 + part * 8);
  }
}

// ---- SIDE = 32 form of the 16-bit kernel (the BASELINE geometry: L = 1024 tokens -> 32 x 32 planes) -----------------
// dwconv16_kernel above is issue-bound (ncu: 73 % of issue slots, 56 instructions per output).  This form
//   * keeps TWO adjacent channels per thread, so every shared-memory access is one 32-bit word, the nine taps are nine
//     packed fma.rn.f32x2 (sm_100) and the polynomial part of the GELU is packed too: ~17 instructions per output;
//   * stages the input tile as [row][x + 1][channel pair] with zero-filled x borders (no bounds tests in the tap loop);
//     the transposition from the plane layout happens on the way in: lanes holding channels c and c + 1 swap halves with
//     one shuffle pair and write 32-bit {c, c + 1} words at a 33-word cell pitch (conflict-free);
//   * writes the result straight to global memory: a warp owns one pixel x 64 channels per store = one 128-byte line.
// Arithmetic (operation order, rounding, the ex2 / rcp GELU) is that of dwconv16_kernel: results are bit-identical.
__device__ __forceinline__ unsigned long long f2_pack(float x, float y) {
  unsigned long long r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(x), "f"(y));
  return r;
}
__device__ __forceinline__ void f2_unpack(unsigned long long v, float& x, float& y) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(x), "=f"(y) : "l"(v));
}
__device__ __forceinline__ unsigned long long f2_fma(unsigned long long a, unsigned long long b, unsigned long long c) {
  unsigned long long d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ unsigned long long f2_mul(unsigned long long a, unsigned long long b) {
  unsigned long long d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ unsigned long long f2_add(unsigned long long a, unsigned long long b) {
  unsigned long long d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
// gelu_fast on a packed pair (same operations as the scalar form in common.cuh)
__device__ __forceinline__ unsigned long long f2_gelu_fast(unsigned long long x) {
  float u0, u1;
  f2_unpack(f2_mul(x, x), u0, u1);
  const unsigned long long u = f2_pack(fminf(u0, 64.0f), fminf(u1, 64.0f));
  unsigned long long q = f2_fma(f2_pack(1.0142650e-3f, 1.0142650e-3f), u, f2_pack(-1.0677574e-1f, -1.0677574e-1f));
  q = f2_fma(q, u, f2_pack(-2.3011213f, -2.3011213f));
  float t0, t1, e0, e1, r0, r1;
  f2_unpack(f2_mul(x, q), t0, t1);
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(t0));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(t1));
  float d0, d1;
  f2_unpack(f2_add(f2_pack(1.0f, 1.0f), f2_pack(e0, e1)), d0, d1);
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r0) : "f"(d0));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r1) : "f"(d1));
  return f2_mul(x, f2_pack(r0, r1));
}
template <typename T> __device__ __forceinline__ unsigned long long word_to_f2(uint32_t w);
template <> __device__ __forceinline__ unsigned long long word_to_f2<__half>(uint32_t w) {
  const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w));
  return f2_pack(f.x, f.y);
}
template <> __device__ __forceinline__ unsigned long long word_to_f2<__nv_bfloat16>(uint32_t w) {
  return f2_pack(__uint_as_float(w << 16), __uint_as_float(w & 0xffff0000u));
}
template <typename T> __device__ __forceinline__ uint32_t f2_to_word(unsigned long long v);
template <> __device__ __forceinline__ uint32_t f2_to_word<__half>(unsigned long long v) {
  float a, b;
  f2_unpack(v, a, b);
  const __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&h);
}
template <> __device__ __forceinline__ uint32_t f2_to_word<__nv_bfloat16>(unsigned long long v) {
  float a, b;
  f2_unpack(v, a, b);
  const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&h);
}

constexpr int DW2_SIDE = 32, DW2_CELLS = DW2_SIDE + 2, DW2_PITCH = 33;     // words per cell: 32 channel pairs + 1 pad
constexpr int DW2_SMEM_WORDS = (DW_ROWS + 2) * DW2_CELLS * DW2_PITCH;

template <typename T>
__global__ void __launch_bounds__(256) dwconv16_v2_kernel(const T* __restrict__ h, T* __restrict__ dt,
                                                          const float* __restrict__ w, const float* __restrict__ bias,
                                                          int hid) {
  static_assert(sizeof(T) == 2, "16-bit storage");
  constexpr int SIDE = DW2_SIDE, L = SIDE * SIDE, ROWS_IN = DW_ROWS + 2;
  __shared__ uint32_t s_in[DW2_SMEM_WORDS];            // [row][x + 1][pair]
  const int b = blockIdx.z, c0 = blockIdx.y * 64, y0 = blockIdx.x * DW_ROWS;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // zero x borders (cells 0 and SIDE + 1 of every row)
  for (int i = threadIdx.x; i < ROWS_IN * 2 * 32; i += 256) {
    const int pr = i & 31, side = (i >> 5) & 1, r = i >> 6;
    s_in[(r * DW2_CELLS + (side ? SIDE + 1 : 0)) * DW2_PITCH + pr] = 0u;
  }
  // plane layout -> [row][x][pair]: lane = (c8 = channel within a group of 8, xv = 8-column vector of the row)
  {
    const T* hb = h + (long long)b * L * hid;
    const int c8 = lane & 7, xv = lane >> 3;
    const bool odd = (c8 & 1) != 0;
#pragma unroll 5                                     // five 16-byte loads in flight per thread (the stage-in is latency-bound)
    for (int it = 0; it < (8 * ROWS_IN) / 8; ++it) {
      const int combo = it * 8 + warp;
      const int cgrp = combo & 7, r = combo >> 3;
      const int yy = y0 - 1 + r;
      uint4 v = make_uint4(0u, 0u, 0u, 0u);
      if (yy >= 0 && yy < SIDE)
        v = *reinterpret_cast<const uint4*>(hb + (long long)(c0 + cgrp * 8 + c8) * L + yy * SIDE + xv * 8);
      const uint32_t send0 = odd ? v.x : v.z, send1 = odd ? v.y : v.w;
      const uint32_t recv0 = __shfl_xor_sync(0xffffffffu, send0, 1), recv1 = __shfl_xor_sync(0xffffffffu, send1, 1);
      const uint32_t mine0 = odd ? v.z : v.x, mine1 = odd ? v.w : v.y;
      const uint32_t a0 = odd ? recv0 : mine0, a1 = odd ? recv1 : mine1;       // even channel of the pair
      const uint32_t b0 = odd ? mine0 : recv0, b1 = odd ? mine1 : recv1;       // odd channel
      uint32_t* dst = s_in + (r * DW2_CELLS + 1 + xv * 8 + (odd ? 4 : 0)) * DW2_PITCH + cgrp * 4 + (c8 >> 1);
      dst[0 * DW2_PITCH] = __byte_perm(a0, b0, 0x5410);
      dst[1 * DW2_PITCH] = __byte_perm(a0, b0, 0x7632);
      dst[2 * DW2_PITCH] = __byte_perm(a1, b1, 0x5410);
      dst[3 * DW2_PITCH] = __byte_perm(a1, b1, 0x7632);
    }
  }
  __syncthreads();
  // thread = (channel pair = lane, output row = warp): sliding 3 x 3 window along x, two channels per packed operation
  const int c = c0 + 2 * lane;
  unsigned long long wk[9];
#pragma unroll
  for (int i = 0; i < 9; ++i) wk[i] = f2_pack(w[c * 9 + i], w[(c + 1) * 9 + i]);
  const unsigned long long bb = f2_pack(bias[c], bias[c + 1]);
  const uint32_t* r0 = s_in + ((warp + 0) * DW2_CELLS) * DW2_PITCH + lane;
  const uint32_t* r1 = r0 + DW2_CELLS * DW2_PITCH;
  const uint32_t* r2 = r1 + DW2_CELLS * DW2_PITCH;
  unsigned long long a0 = word_to_f2<T>(r0[0]), a1 = word_to_f2<T>(r1[0]), a2 = word_to_f2<T>(r2[0]);          // column x - 1
  unsigned long long m0 = word_to_f2<T>(r0[DW2_PITCH]), m1 = word_to_f2<T>(r1[DW2_PITCH]), m2 = word_to_f2<T>(r2[DW2_PITCH]);
  uint32_t* out = reinterpret_cast<uint32_t*>(dt + ((long long)b * L + (y0 + warp) * SIDE) * hid + c);
  const int out_step = hid / 2;                        // words per pixel
#pragma unroll 4
  for (int x = 0; x < SIDE; ++x) {
    const unsigned long long n0 = word_to_f2<T>(r0[(x + 2) * DW2_PITCH]), n1 = word_to_f2<T>(r1[(x + 2) * DW2_PITCH]),
                             n2 = word_to_f2<T>(r2[(x + 2) * DW2_PITCH]);                                           // column x + 1
    unsigned long long acc = bb;
    acc = f2_fma(a0, wk[0], acc); acc = f2_fma(m0, wk[1], acc); acc = f2_fma(n0, wk[2], acc);
    acc = f2_fma(a1, wk[3], acc); acc = f2_fma(m1, wk[4], acc); acc = f2_fma(n1, wk[5], acc);
    acc = f2_fma(a2, wk[6], acc); acc = f2_fma(m2, wk[7], acc); acc = f2_fma(n2, wk[8], acc);
    out[(long long)x * out_step] = f2_to_word<T>(f2_gelu_fast(acc));
    a0 = m0; a1 = m1; a2 = m2;
    m0 = n0; m1 = n1; m2 = n2;
  }
}

template <typename T>
static int launch_dwconv16(const void* h, void* dt, const float* w, const float* b, int B, int L, int hid, int side,
                           cudaStream_t st) {
  dim3 grid((side + DW_ROWS - 1) / DW_ROWS, hid / 64, B);
  if (side == DW2_SIDE && (reinterpret_cast<uintptr_t>(h) & 15) == 0 && (reinterpret_cast<uintptr_t>(dt) & 3) == 0) {
    dwconv16_v2_kernel<T><<<grid, 256, 0, st>>>((const T*)h, (T*)dt, w, b, hid);
    DPMN_LAUNCH_CHECK();
    return 0;
  }
  const size_t smem = (size_t)(64 * ((DW_ROWS + 2) * side + 2) + DW_ROWS * side * 64) * 2;
  auto k = dwconv16_kernel<T>;
  static PerDeviceOnce attr;      // per template instantiation, per device
  DPMN_CUDA_TRY(attr.smem_attr(k, (int)smem));
  k<<<grid, 256, smem, st>>>((const T*)h, (T*)dt, w, b, L, hid, side);
  DPMN_LAUNCH_CHECK();
  return 0;
}

int launch_dwconv(const void* h, void* dt, DType io_type, const float* w, const float* b, int B, int L, int hid,
                  cudaStream_t st) {
  int side = (int)(sqrtf((float)L) + 0.5f);
  if (side * side != L || hid % 32 != 0) return -2;   // the reference's view() needs a perfect square too
  if (io_type != DT_F32 && hid % 64 == 0 && side % 8 == 0 && side <= 64) {
    return io_type == DT_F16 ? launch_dwconv16<__half>(h, dt, w, b, B, L, hid, side, st)
                             : launch_dwconv16<__nv_bfloat16>(h, dt, w, b, B, L, hid, side, st);
  }
  dim3 grid((side + DW_ROWS - 1) / DW_ROWS, hid / 32, B);
  const size_t smem = (size_t)32 * ((DW_ROWS + 2) * side + 1) * sizeof(float);
  if (smem > 200 * 1024) return -2;
  switch (io_type) {
    case DT_F32: {
      auto k = dwconv_kernel<float>;
      if (smem > 48 * 1024) DPMN_CUDA_TRY(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      k<<<grid, 256, smem, st>>>((const float*)h, (float*)dt, w, b, L, hid, side);
      break;
    }
    case DT_F16: {
      auto k = dwconv_kernel<__half>;
      if (smem > 48 * 1024) DPMN_CUDA_TRY(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      k<<<grid, 256, smem, st>>>((const __half*)h, (__half*)dt, w, b, L, hid, side);
      break;
    }
    case DT_BF16: {
      auto k = dwconv_kernel<__nv_bfloat16>;
      if (smem > 48 * 1024) DPMN_CUDA_TRY(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      k<<<grid, 256, smem, st>>>((const __nv_bfloat16*)h, (__nv_bfloat16*)dt, w, b, L, hid, side);
      break;
    }
  }
  DPMN_LAUNCH_CHECK();
  return 0;
}

// =====================================================================================================
// K5  head: PatchUnEmbed + conv3x3 (C -> hp)                                           pgrm.py:559-560
// x is token-major (B, gh, gw, C) == NHWC, which is what PatchUnEmbed's transpose+view describes.
// One thread per output pixel, all hp (<= 16) outputs in registers; weights in smem as [tap][ci][hp].
// =====================================================================================================
constexpr int HP_MAX = 16;

__global__ void __launch_bounds__(128) head_conv1_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                         const float* __restrict__ bias, float* __restrict__ t1,
                                                         int total, int gh, int gw, int C, int hp) {
  extern __shared__ __align__(16) float sw[];   // [9][C][HP_MAX]
  for (int i = threadIdx.x; i < 9 * C * HP_MAX; i += blockDim.x) {
    const int co = i % HP_MAX;
    const int ci = (i / HP_MAX) % C;
    const int tap = i / (HP_MAX * C);
    sw[i] = co < hp ? w[(co * C + ci) * 9 + tap] : 0.f;
  }
  __syncthreads();
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int xx = idx % gw;
  const int yy = (idx / gw) % gh;
  const int b = idx / (gw * gh);
  float acc[HP_MAX];
#pragma unroll
  for (int j = 0; j < HP_MAX; ++j) acc[j] = j < hp ? bias[j] : 0.f;
  for (int ky = 0; ky < 3; ++ky) {
    const int y2 = yy + ky - 1;
    if (y2 < 0 || y2 >= gh) continue;
    for (int kx = 0; kx < 3; ++kx) {
      const int x2 = xx + kx - 1;
      if (x2 < 0 || x2 >= gw) continue;
      const float4* src = reinterpret_cast<const float4*>(x + ((long long)(b * gh + y2) * gw + x2) * C);
      const float* wt = sw + (ky * 3 + kx) * C * HP_MAX;
      for (int c4 = 0; c4 < C / 4; ++c4) {
        const float4 v = src[c4];
        const float vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const float4* w4 = reinterpret_cast<const float4*>(wt + (c4 * 4 + u) * HP_MAX);
#pragma unroll
          for (int j4 = 0; j4 < HP_MAX / 4; ++j4) {
            const float4 ww = w4[j4];
            acc[4 * j4 + 0] = fmaf(vv[u], ww.x, acc[4 * j4 + 0]);
            acc[4 * j4 + 1] = fmaf(vv[u], ww.y, acc[4 * j4 + 1]);
            acc[4 * j4 + 2] = fmaf(vv[u], ww.z, acc[4 * j4 + 2]);
            acc[4 * j4 + 3] = fmaf(vv[u], ww.w, acc[4 * j4 + 3]);
          }
        }
      }
    }
  }
  float* dst = t1 + (long long)idx * HP_MAX;
#pragma unroll
  for (int j4 = 0; j4 < HP_MAX / 4; ++j4)
    *reinterpret_cast<float4*>(dst + 4 * j4) = make_float4(acc[4 * j4], acc[4 * j4 + 1], acc[4 * j4 + 2], acc[4 * j4 + 3]);
}

int launch_head_conv1(const float* x, const float* w, const float* b, float* t1, int B, int gh, int gw, int C,
                      int hp, cudaStream_t st) {
  if (hp > HP_MAX || C % 4) return -2;
  const int total = B * gh * gw;
  const size_t smem = (size_t)9 * C * HP_MAX * sizeof(float);
  if (smem > 200 * 1024) return -2;
  if (smem > 48 * 1024)
    DPMN_CUDA_TRY(cudaFuncSetAttribute(head_conv1_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  head_conv1_kernel<<<(total + 127) / 128, 128, smem, st>>>(x, w, b, t1, total, gh, gw, C, hp);
  DPMN_LAUNCH_CHECK();
  return 0;
}

// conv3x3 (hp -> hp) + LeakyReLU(0.01) + PixelShuffle(r) + affine mix                  pgrm.py:560-564
// t1 is (B, gh, gw, HP_MAX) from head_conv1.  out[b, c, y*r+dy, x*r+dx] = lrelu(conv)[c*r*r + dy*r + dx]
// * weight_list_0 + sum_{i>=1} residual_i * weight_list_i   (residual_list[0] is never read: quirk 3).
__global__ void __launch_bounds__(128) head_conv2_mix_kernel(const float* __restrict__ t1,
                                                             const float* __restrict__ w,
                                                             const float* __restrict__ bias, float* __restrict__ out,
                                                             int total, int gh, int gw, int hs, int r, MixArgs mix) {
  __shared__ __align__(16) float sw[9 * HP_MAX * HP_MAX];   // [tap][ci][co]
  const int hp = hs * r * r;
  for (int i = threadIdx.x; i < 9 * HP_MAX * HP_MAX; i += blockDim.x) {
    const int co = i % HP_MAX;
    const int ci = (i / HP_MAX) % HP_MAX;
    const int tap = i / (HP_MAX * HP_MAX);
    sw[i] = (co < hp && ci < hp) ? w[(co * hp + ci) * 9 + tap] : 0.f;
  }
  __syncthreads();
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int xx = idx % gw;
  const int yy = (idx / gw) % gh;
  const int b = idx / (gw * gh);
  float acc[HP_MAX];
#pragma unroll
  for (int j = 0; j < HP_MAX; ++j) acc[j] = j < hp ? bias[j] : 0.f;
  for (int ky = 0; ky < 3; ++ky) {
    const int y2 = yy + ky - 1;
    if (y2 < 0 || y2 >= gh) continue;
    for (int kx = 0; kx < 3; ++kx) {
      const int x2 = xx + kx - 1;
      if (x2 < 0 || x2 >= gw) continue;
      const float4* src = reinterpret_cast<const float4*>(t1 + ((long long)(b * gh + y2) * gw + x2) * HP_MAX);
      const float* wt = sw + (ky * 3 + kx) * HP_MAX * HP_MAX;
#pragma unroll
      for (int c4 = 0; c4 < HP_MAX / 4; ++c4) {
        const float4 v = src[c4];
        const float vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const float4* w4 = reinterpret_cast<const float4*>(wt + (c4 * 4 + u) * HP_MAX);
#pragma unroll
          for (int j4 = 0; j4 < HP_MAX / 4; ++j4) {
            const float4 ww = w4[j4];
            acc[4 * j4 + 0] = fmaf(vv[u], ww.x, acc[4 * j4 + 0]);
            acc[4 * j4 + 1] = fmaf(vv[u], ww.y, acc[4 * j4 + 1]);
            acc[4 * j4 + 2] = fmaf(vv[u], ww.z, acc[4 * j4 + 2]);
            acc[4 * j4 + 3] = fmaf(vv[u], ww.w, acc[4 * j4 + 3]);
          }
        }
      }
    }
  }
  const int img_h = gh * r, img_w = gw * r;
  const long long plane = (long long)img_h * img_w;
#pragma unroll
  for (int j = 0; j < HP_MAX; ++j) {
    if (j >= hp) break;
    float v = acc[j];
    v = v >= 0.f ? v : 0.01f * v;                       // nn.LeakyReLU default slope (pgrm.py:520)
    const int c = j / (r * r);
    const int rem = j - c * r * r;
    const int dy = rem / r, dx = rem - dy * r;
    const long long pix = (long long)c * plane + (long long)(yy * r + dy) * img_w + (xx * r + dx);
    float res = v * mix.w[0][pix];
    for (int i = 1; i < mix.n_mix; ++i)
      res = fmaf(mix.in[i][(long long)b * mix.in_bs[i] + pix], mix.w[i][pix], res);
    out[(long long)b * hs * plane + pix] = res;
  }
}

// Four lanes per token (hp = hs r^2 = 12: hidden_size 3, PixelShuffle 2): lane q owns output channels 3 q .. 3 q + 2, so a
// thread runs 324 multiply-adds instead of 2 304 at 4x the threads (the one-thread-per-token kernel above has 2.6 warps per
// scheduler at batch 48 and is latency-bound at 21 us); weights sit in shared memory as [tap][ci][q] float4 (one LDS.128 per
// (tap, ci)), the four lanes of a token read the same 48 bytes of t1 per tap (one broadcast transaction).
__global__ void __launch_bounds__(256) head_conv2_mix4_kernel(const float* __restrict__ t1, const float* __restrict__ w,
                                                              const float* __restrict__ bias, float* __restrict__ out,
                                                              int total, int gh, int gw, int hs, int r, MixArgs mix) {
  constexpr int HP = 12;
  __shared__ __align__(16) float4 sw[9 * HP * 4];   // [tap][ci][q] -> (co = 3 q, 3 q + 1, 3 q + 2, -)
  for (int i = threadIdx.x; i < 9 * HP * 4; i += blockDim.x) {
    const int q = i & 3, ci = (i >> 2) % HP, tap = i / (4 * HP);
    sw[i] = make_float4(w[((3 * q) * HP + ci) * 9 + tap], w[((3 * q + 1) * HP + ci) * 9 + tap],
                        w[((3 * q + 2) * HP + ci) * 9 + tap], 0.f);
  }
  __syncthreads();
  const int gtid = blockIdx.x * blockDim.x + threadIdx.x;
  const int idx = gtid >> 2, q = gtid & 3;
  if (idx >= total) return;
  const int xx = idx % gw;
  const int yy = (idx / gw) % gh;
  const int b = idx / (gw * gh);
  float acc[3] = {bias[3 * q], bias[3 * q + 1], bias[3 * q + 2]};
#pragma unroll
  for (int ky = 0; ky < 3; ++ky) {
    const int y2 = yy + ky - 1;
    if (y2 < 0 || y2 >= gh) continue;
#pragma unroll
    for (int kx = 0; kx < 3; ++kx) {
      const int x2 = xx + kx - 1;
      if (x2 < 0 || x2 >= gw) continue;
      const float4* src = reinterpret_cast<const float4*>(t1 + ((long long)(b * gh + y2) * gw + x2) * HP_MAX);
      const float4* wt = sw + (ky * 3 + kx) * HP * 4 + q;
#pragma unroll
      for (int c4 = 0; c4 < HP / 4; ++c4) {
        const float4 v = __ldg(src + c4);
        const float vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const float4 ww = wt[(c4 * 4 + u) * 4];
          acc[0] = fmaf(vv[u], ww.x, acc[0]); acc[1] = fmaf(vv[u], ww.y, acc[1]); acc[2] = fmaf(vv[u], ww.z, acc[2]);
        }
      }
    }
  }
  const int img_h = gh * r, img_w = gw * r;
  const long long plane = (long long)img_h * img_w;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const int j = 3 * q + i;
    float v = acc[i];
    v = v >= 0.f ? v : 0.01f * v;                       // nn.LeakyReLU default slope (pgrm.py:520)
    const int c = j / (r * r);
    const int rem = j - c * r * r;
    const int dy = rem / r, dx = rem - dy * r;
    const long long pix = (long long)c * plane + (long long)(yy * r + dy) * img_w + (xx * r + dx);
    float res = v * mix.w[0][pix];
    for (int k = 1; k < mix.n_mix; ++k)
      res = fmaf(mix.in[k][(long long)b * mix.in_bs[k] + pix], mix.w[k][pix], res);
    out[(long long)b * hs * plane + pix] = res;
  }
}

int launch_head_conv2_mix(const float* t1, const float* w, const float* b, float* out, int B, int gh, int gw,
                          int hs, int patch, const MixArgs& mix, cudaStream_t st) {
  if (hs * patch * patch > HP_MAX || mix.n_mix < 1 || mix.n_mix > 8) return -2;
  const int total = B * gh * gw;
  static const int head_v = getenv("DPMN_HEAD_MIX") ? atoi(getenv("DPMN_HEAD_MIX")) : 4;   // 1: one thread per token (A/B)
  if (head_v == 4 && hs * patch * patch == 12)
    head_conv2_mix4_kernel<<<(total * 4 + 255) / 256, 256, 0, st>>>(t1, w, b, out, total, gh, gw, hs, patch, mix);
  else
    head_conv2_mix_kernel<<<(total + 127) / 128, 128, 0, st>>>(t1, w, b, out, total, gh, gw, hs, patch, mix);
  DPMN_LAUNCH_CHECK();
  return 0;
}

// =====================================================================================================
template <typename T>
__global__ void convert_kernel(const float* __restrict__ src, T* __restrict__ dst, long long n) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    dst[i] = from_f32<T>(src[i]);
}

int launch_convert(const float* src, void* dst, DType dst_type, long long n, cudaStream_t st) {
  long long blocks = (n + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  if (blocks < 1) blocks = 1;
  if (dst_type == DT_F16) convert_kernel<__half><<<(int)blocks, 256, 0, st>>>(src, (__half*)dst, n);
  else if (dst_type == DT_BF16) convert_kernel<__nv_bfloat16><<<(int)blocks, 256, 0, st>>>(src, (__nv_bfloat16*)dst, n);
  else return -1;
  DPMN_LAUNCH_CHECK();
  return 0;
}

template <typename T>
__global__ void convert_batch_kernel(ConvertBatch cb) {
  const int seg = blockIdx.y;
  const float* __restrict__ src = cb.src[seg];
  T* __restrict__ dst = reinterpret_cast<T*>(cb.dst[seg]);
  const long long n = cb.n[seg];
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) dst[i] = from_f32<T>(src[i]);
}

int launch_convert_batch(const ConvertBatch& cb, DType dst_type, cudaStream_t st) {
  if (cb.count < 1 || cb.count > 16) return -1;
  dim3 grid(64, cb.count);
  if (dst_type == DT_F16) convert_batch_kernel<__half><<<grid, 256, 0, st>>>(cb);
  else if (dst_type == DT_BF16) convert_batch_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(cb);
  else return -1;
  DPMN_LAUNCH_CHECK();
  return 0;
}

template <typename T>
__global__ void widen_kernel(const T* __restrict__ src, float* __restrict__ dst, long long n) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) dst[i] = to_f32<T>(src[i]);
}

int launch_widen(const void* src, DType src_type, float* dst, long long n, cudaStream_t st) {
  long long blocks = (n + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  if (src_type == DT_F16) widen_kernel<__half><<<(int)blocks, 256, 0, st>>>((const __half*)src, dst, n);
  else if (src_type == DT_BF16) widen_kernel<__nv_bfloat16><<<(int)blocks, 256, 0, st>>>((const __nv_bfloat16*)src, dst, n);
  else return -1;
  DPMN_LAUNCH_CHECK();
  return 0;
}

}  // namespace dpmn
