// tcgen05 implicit-GEMM convolution for the Complementation Modulation Module (cmm.py:38-118) on sm_100a.
//
//   out[pixel, co] = epi( sum_{tap} sum_{ci} X[pixel shifted by tap, ci] * Wt[tap][co][ci] )
//
// Activations are NHWC 16-bit.  M = 128 output-grid pixels (a TMA box {64 ch, bw, bh, bb} of the activation
// tensor per tap, OOB pixels zero-filled = the conv padding), N = BN output channels, K = taps x Cin in
// 64-channel blocks.  No im2col is ever materialised: every tap is a shifted TMA box load.
// Strided layers become stride-1 problems on sub-grids of the NHWC tensor (a tensor map with doubled strides):
//   * conv 4x4 stride 2 dilation 2 pad 3 (cmm.py:44) only reads odd rows/cols -> 4x4 stride-1 conv on that sub-grid
//   * conv 4x4 stride 2 pad 1 (cmm.py:92): each tap reads one of the 4 parity sub-grids
//   * convT 4x4 stride 2 pad 1 (cmm.py:67,109): 4 output-parity classes, each a 2x2 stride-1 conv
//   * convT 3x3 stride 1 pad 1 (cmm.py:62,115): conv with mirrored taps
// Same warp-specialised pipeline as gemm_tc.cu (TMA warp, MMA warp, 4 epilogue warps, TMEM double buffer).
// Epilogue: per-channel scale/shift (conv bias + eval BatchNorm folded), then up to two stores with their own
// activation (LeakyReLU 0.2 for the next encoder stage, ReLU for the decoder skip), 16-bit NHWC or fp32.
#include "common.cuh"
#include "kernels.h"
#include "tc_common.cuh"
#include <cstdlib>
#include <cstring>

namespace dpmn {

using namespace tc;

constexpr int CBM = 128;
constexpr int CBK = 64;
constexpr int CSTAGES = 3;   // <= 113 KB per CTA so that two CTAs share an SM
constexpr int CONV_THREADS = 192;

struct ConvTcParams {
  int Cin, Cout, B, G, P;            // P output-parity classes (1 or 4)
  int Hm, Wm;                        // grid the M tiles enumerate (per image)
  int bw, bh, bb;                    // box: bw*bh*bb == 128
  int wt, ht, bt, n_tiles;           // tile counts
  int n_taps;                        // taps per class
  ConvTap taps[4][16];
  int os, Ho, Wo;                    // output pixel = (y*os + py, x*os + px) in an (Ho, Wo) grid; class p -> (py, px) = (p>>1, p&1)
  const float* scale;                // [G][Cout] or nullptr (1)
  const float* shift;                // [G][Cout] or nullptr (0)
  ConvTcDest dst[2];
  int fmt;                           // 0 fp16, 1 bf16
  int ksplit, kb_per;                // split-K: slices and k-blocks per slice (ksplit == 1: none)
  float* part;                       // [ksplit][G][npix][Cout] fp32 partial sums
  long long npix;                    // B * Ho * Wo
};

template <int BN>
struct ConvSmem {
  static constexpr int A_BYTES = CBM * CBK * 2;
  static constexpr int B_BYTES = BN * CBK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int TOTAL = CSTAGES * STAGE_BYTES + 1024 + 256;
};

template <typename T>
__device__ __forceinline__ void store_chunk(const ConvTcDest& d, int g, long long pix, int n0, int ncols, const float (&v)[32]) {
  // 32 consecutive output channels of one pixel
  const long long off = (long long)g * d.g_stride + pix * d.ld + d.ch_off + g * d.ch_g_off + n0;
  float a[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    float x = v[j];
    if (d.act == 1) x = x >= 0.f ? x : 0.2f * x;
    else if (d.act == 2) x = fmaxf(x, 0.f);
    a[j] = x;
  }
  if constexpr (sizeof(T) == 4) {
    float* p = reinterpret_cast<float*>(d.ptr) + off;
#pragma unroll
    for (int j = 0; j < 32; j += 4)
      if (j < ncols) *reinterpret_cast<float4*>(p + j) = make_float4(a[j], a[j + 1], a[j + 2], a[j + 3]);
  } else {
    T* p = reinterpret_cast<T*>(d.ptr) + off;
#pragma unroll
    for (int j = 0; j < 32; j += 8) {
      if (j < ncols) {
        union { uint4 u; T h[8]; } pk;
#pragma unroll
        for (int e = 0; e < 8; ++e) pk.h[e] = from_f32<T>(a[j + e]);
        *reinterpret_cast<uint4*>(p + j) = pk.u;
      }
    }
  }
}

template <int BN>
__global__ void __launch_bounds__(CONV_THREADS, 2)
conv_tc_kernel(const __grid_constant__ CUtensorMap map_a0, const __grid_constant__ CUtensorMap map_a1,
               const __grid_constant__ CUtensorMap map_a2, const __grid_constant__ CUtensorMap map_a3,
               const __grid_constant__ CUtensorMap map_w, const __grid_constant__ ConvTcParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  using S = ConvSmem<BN>;
  uint8_t* tiles = smem;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + CSTAGES * S::STAGE_BYTES);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + CSTAGES;
  uint64_t* tmem_full = bars + 2 * CSTAGES;
  uint64_t* tmem_empty = bars + 2 * CSTAGES + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * CSTAGES + 4);
  constexpr uint32_t TMEM_COLS = (2 * BN <= 32) ? 32 : (2 * BN <= 64) ? 64 : (2 * BN <= 128) ? 128 : (2 * BN <= 256) ? 256 : 512;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int cblocks = (p.Cin + CBK - 1) / CBK;
  const int num_kb = p.n_taps * cblocks;
  const int pix_tiles = p.wt * p.ht * p.bt;
  const int total_tiles = p.G * p.P * pix_tiles * p.n_tiles * p.ksplit;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_a0);
    tma_prefetch_desc(&map_w);
    for (int i = 0; i < CSTAGES; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tmem_full[i], 1); mbar_init(&tmem_empty[i], 4); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();        // everything above overlapped the predecessor's tail; its outputs are visible from here on
  pdl_trigger();

  // tile index -> (n block, pixel-tile origin, parity class, group); n fastest so neighbours share the A boxes in L2
#define DPMN_DECODE_TILE(t)                                   \
  const int ks = (t) % p.ksplit;                              \
  const int kb0 = ks * p.kb_per;                              \
  const int kb1 = kb0 + p.kb_per < num_kb ? kb0 + p.kb_per : num_kb;   \
  const int n_blk = ((t) / p.ksplit) % p.n_tiles;             \
  int _r = (t) / (p.ksplit * p.n_tiles);                      \
  const int w0 = (_r % p.wt) * p.bw; _r /= p.wt;              \
  const int h0 = (_r % p.ht) * p.bh; _r /= p.ht;              \
  const int b0 = (_r % p.bt) * p.bb; _r /= p.bt;              \
  const int cls = _r % p.P;                                   \
  const int g = _r / p.P;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
        DPMN_DECODE_TILE(t)
        int tap = kb0 / cblocks, cb = kb0 - tap * cblocks;
        for (int kb = kb0; kb < kb1; ++kb) {
          const ConvTap tp = p.taps[cls][tap];
          const CUtensorMap* ma = tp.map == 0 ? &map_a0 : tp.map == 1 ? &map_a1 : tp.map == 2 ? &map_a2 : &map_a3;
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = tiles + stage * S::STAGE_BYTES;
          uint8_t* sb = sa + S::A_BYTES;
          mbar_arrive_expect_tx(&full_bar[stage], S::STAGE_BYTES);
          tma_load_5d(sa, ma, &full_bar[stage], cb * CBK, w0 + tp.dx, h0 + tp.dy, b0, g);
          tma_load_4d(sb, &map_w, &full_bar[stage], cb * CBK, n_blk * BN, tp.wslice, g);
          if (++stage == CSTAGES) { stage = 0; phase ^= 1; }
          if (++cb == cblocks) { cb = 0; ++tap; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = make_idesc_f16(p.fmt, CBM, BN);
      int stage = 0; uint32_t phase = 0;
      int acc = 0; uint32_t acc_phase = 0;
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
        const int kb0 = (t % p.ksplit) * p.kb_per;
        const int kb1 = kb0 + p.kb_per < num_kb ? kb0 + p.kb_per : num_kb;
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
        for (int kb = kb0; kb < kb1; ++kb) {
          const int cb = kb % cblocks;
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(tiles + stage * S::STAGE_BYTES);
          const uint64_t da = make_smem_desc_sw128(sa);
          const uint64_t db = make_smem_desc_sw128(sa + S::A_BYTES);
          const int c_left = p.Cin - cb * CBK;
          const int ksteps = c_left >= CBK ? CBK / 16 : (c_left + 15) / 16;
          for (int k = 0; k < ksteps; ++k)
            umma_f16(d_tmem, advance_desc_k(da, k), advance_desc_k(db, k), idesc, ((kb - kb0) | k) ? 1u : 0u);
          umma_commit(&empty_bar[stage]);
          if (++stage == CSTAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit(&tmem_full[acc]);
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    const int quarter = warp & 3;
    int acc = 0; uint32_t acc_phase = 0;
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
      DPMN_DECODE_TILE(t)
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      const int r = quarter * 32 + lane;              // tile row = box element (w fastest, then h, then image)
      const int wi = r % p.bw;
      const int hi = (r / p.bw) % p.bh;
      const int bi = r / (p.bw * p.bh);
      const int x = w0 + wi, y = h0 + hi, b = b0 + bi;
      const bool ok = x < p.Wm && y < p.Hm && b < p.B;
      const long long pix = ((long long)b * p.Ho + (y * p.os + (cls >> 1))) * p.Wo + (x * p.os + (cls & 1));
#pragma unroll 1
      for (int c0 = 0; c0 < BN; c0 += 32) {
        const int n0 = n_blk * BN + c0;
        // per-channel scale / shift first, unpredicated (clamped columns): in-order issue must not serialise them
        float4 sc4[8], sh4[8];
        if (p.scale != nullptr) {
#pragma unroll
          for (int j = 0; j < 8; ++j)
            sc4[j] = *reinterpret_cast<const float4*>(p.scale + (long long)g * p.Cout + min(n0 + 4 * j, p.Cout - 4));
        }
        if (p.shift != nullptr) {
#pragma unroll
          for (int j = 0; j < 8; ++j)
            sh4[j] = *reinterpret_cast<const float4*>(p.shift + (long long)g * p.Cout + min(n0 + 4 * j, p.Cout - 4));
        }
        uint32_t rr[32];
        tmem_ld_32x32(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * BN + c0), rr);
        tmem_ld_wait();
        if (!ok || n0 >= p.Cout) continue;
        const int ncols = p.Cout - n0 < 32 ? p.Cout - n0 : 32;
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(rr[j]);
        if (p.ksplit > 1) {                             // raw partial sums; conv_splitk_reduce_kernel applies the epilogue
          float* pp = p.part + (((long long)ks * p.G + g) * p.npix + pix) * p.Cout + n0;
#pragma unroll
          for (int j = 0; j < 32; j += 4)
            if (j < ncols) *reinterpret_cast<float4*>(pp + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
          continue;
        }
        if (p.scale != nullptr) {
#pragma unroll
          for (int j = 0; j < 8; ++j) { v[4 * j] *= sc4[j].x; v[4 * j + 1] *= sc4[j].y; v[4 * j + 2] *= sc4[j].z; v[4 * j + 3] *= sc4[j].w; }
        }
        if (p.shift != nullptr) {
#pragma unroll
          for (int j = 0; j < 8; ++j) { v[4 * j] += sh4[j].x; v[4 * j + 1] += sh4[j].y; v[4 * j + 2] += sh4[j].z; v[4 * j + 3] += sh4[j].w; }
        }
#pragma unroll
        for (int di = 0; di < 2; ++di) {
          const ConvTcDest& d = p.dst[di];
          if (d.ptr == nullptr) continue;
          if (d.type == DT_F32) store_chunk<float>(d, g, pix, n0, ncols, v);
          else if (d.type == DT_F16) store_chunk<__half>(d, g, pix, n0, ncols, v);
          else store_chunk<__nv_bfloat16>(d, g, pix, n0, ncols, v);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }
#undef DPMN_DECODE_TILE

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// Split-K second pass: out[g, pix, c] = epi( sum_ks part[ks, g, pix, c] ), slices added in order (deterministic); 8 channels
// per thread, the same scale / shift / dual-destination epilogue as the fused path.
struct SplitKReduceParams {
  const float* part; int ksplit, G, Cout; long long npix;
  const float* scale; const float* shift;
  ConvTcDest dst[2];
};
__global__ void __launch_bounds__(256) conv_splitk_reduce_kernel(const SplitKReduceParams p) {
  const int c8 = p.Cout / 8;
  const long long total = (long long)p.G * p.npix * c8;
  const long long slice = (long long)p.G * p.npix * p.Cout;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % c8) * 8;
    const long long gp = i / c8;
    const long long pix = gp % p.npix;
    const int g = (int)(gp / p.npix);
    const float* src = p.part + gp * p.Cout + c;
    float v[8];
    {
      const float4 a = *reinterpret_cast<const float4*>(src), b = *reinterpret_cast<const float4*>(src + 4);
      v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    }
    for (int ks = 1; ks < p.ksplit; ++ks) {
      const float4 a = *reinterpret_cast<const float4*>(src + ks * slice), b = *reinterpret_cast<const float4*>(src + ks * slice + 4);
      v[0] += a.x; v[1] += a.y; v[2] += a.z; v[3] += a.w; v[4] += b.x; v[5] += b.y; v[6] += b.z; v[7] += b.w;
    }
    if (p.scale != nullptr) {
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] *= p.scale[(long long)g * p.Cout + c + j];
    }
    if (p.shift != nullptr) {
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] += p.shift[(long long)g * p.Cout + c + j];
    }
#pragma unroll
    for (int di = 0; di < 2; ++di) {
      const ConvTcDest& d = p.dst[di];
      if (d.ptr == nullptr) continue;
      float a[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) a[j] = d.act == 1 ? (v[j] >= 0.f ? v[j] : 0.2f * v[j]) : (d.act == 2 ? fmaxf(v[j], 0.f) : v[j]);
      const long long off = (long long)g * d.g_stride + pix * d.ld + d.ch_off + g * d.ch_g_off + c;
      if (d.type == DT_F32) {
        float* q = reinterpret_cast<float*>(d.ptr) + off;
        *reinterpret_cast<float4*>(q) = make_float4(a[0], a[1], a[2], a[3]);
        *reinterpret_cast<float4*>(q + 4) = make_float4(a[4], a[5], a[6], a[7]);
      } else if (d.type == DT_F16) {
        union { uint4 u; __half h[8]; } pk;
#pragma unroll
        for (int j = 0; j < 8; ++j) pk.h[j] = __float2half_rn(a[j]);
        *reinterpret_cast<uint4*>(reinterpret_cast<__half*>(d.ptr) + off) = pk.u;
      } else {
        union { uint4 u; __nv_bfloat16 h[8]; } pk;
#pragma unroll
        for (int j = 0; j < 8; ++j) pk.h[j] = __float2bfloat16_rn(a[j]);
        *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(d.ptr) + off) = pk.u;
      }
    }
  }
}

// ---- host -------------------------------------------------------------------------------------------------
static int next_pow2(int v) { int p = 1; while (p < v) p <<= 1; return p; }

template <int BN>
static int launch_conv_bn(const ConvTcArgs& a, cudaStream_t st) {
  ConvTcParams p;
  memset(&p, 0, sizeof(p));
  p.Cin = a.Cin; p.Cout = a.Cout; p.B = a.B; p.G = a.G; p.P = a.P; p.Hm = a.Hm; p.Wm = a.Wm;
  p.bw = next_pow2(a.Wm) < CBM ? next_pow2(a.Wm) : CBM;
  p.bh = next_pow2(a.Hm) < CBM / p.bw ? next_pow2(a.Hm) : CBM / p.bw;
  p.bb = CBM / (p.bw * p.bh);
  p.wt = (a.Wm + p.bw - 1) / p.bw; p.ht = (a.Hm + p.bh - 1) / p.bh; p.bt = (a.B + p.bb - 1) / p.bb;
  p.n_tiles = (a.Cout + BN - 1) / BN;
  p.n_taps = a.n_taps;
  for (int c = 0; c < 4; ++c)
    for (int t = 0; t < 16; ++t) p.taps[c][t] = a.taps[c][t];
  p.os = a.os; p.Ho = a.Ho; p.Wo = a.Wo; p.scale = a.scale; p.shift = a.shift;
  p.dst[0] = a.dst[0]; p.dst[1] = a.dst[1];
  p.fmt = a.op_type == DT_BF16 ? 1 : 0;

  CUtensorMap maps[4];
  for (int m = 0; m < 4; ++m) {
    const ConvTcSrc& s = a.src[m < a.n_src ? m : 0];
    const uint64_t dims[5] = {(uint64_t)a.Cin, (uint64_t)a.Wm, (uint64_t)a.Hm, (uint64_t)a.B, (uint64_t)a.G};
    const uint64_t str[4] = {(uint64_t)s.sx * 2, (uint64_t)s.sy * 2, (uint64_t)s.sb * 2, (uint64_t)(a.G > 1 ? s.sg : s.sb * a.B) * 2};
    const uint32_t box[5] = {CBK, (uint32_t)p.bw, (uint32_t)p.bh, (uint32_t)p.bb, 1};
    int rc = make_tensor_map_16bit(&maps[m], s.base, 5, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
  }
  CUtensorMap map_w;
  {
    const uint64_t dims[4] = {(uint64_t)a.Cin, (uint64_t)a.Cout, (uint64_t)a.n_wslices, (uint64_t)a.G};
    const uint64_t str[3] = {(uint64_t)a.Cin * 2, (uint64_t)a.Cin * a.Cout * 2,
                             (uint64_t)a.Cin * a.Cout * a.n_wslices * 2};
    const uint32_t box[4] = {CBK, BN, 1, 1};
    int rc = make_tensor_map_16bit(&map_w, a.w, 4, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
  }
  int num_sms = 0;
  DPMN_CUDA_TRY(current_device_sms(&num_sms));
  const int tiles_mn = p.G * p.P * p.wt * p.ht * p.bt * p.n_tiles;
  // split-K: only when the (pixel, Cout) tiles cover less than half of the 2-CTA-per-SM grid and the reduction is long
  p.ksplit = 1; p.npix = (long long)a.B * a.Ho * a.Wo;
  {
    // Opt-in (DPMN_CONV_SPLITK=1).  Measured on B200 (profiles/r02_splitk_ab.md): the six deep CMM layers drop from 380 to
    // ~180 us in a one-stream forward (conv_tc class 1.00 -> 0.80 ms), but the pipelined step does not move (2.49 -> 2.50 ms:
    // those layers only occupy 32-96 SMs and the other streams' kernels already fill the rest), and a batch-dependent slice
    // count makes the result depend on the batch size in the last bit (tests pin batch-size invariance).
    static const bool splitk_on = getenv("DPMN_CONV_SPLITK") && atoi(getenv("DPMN_CONV_SPLITK")) != 0;
    const int num_kb = a.n_taps * ((a.Cin + CBK - 1) / CBK);
    if (splitk_on && a.splitk_ws != nullptr && tiles_mn * 2 <= 2 * num_sms && num_kb >= 16) {
      int ks = (2 * num_sms) / tiles_mn;
      if (ks > num_kb / 8) ks = num_kb / 8;                       // at least 8 k-blocks per slice
      const size_t per_slice = (size_t)a.G * p.npix * a.Cout;
      if ((size_t)ks * per_slice > a.splitk_floats) ks = (int)(a.splitk_floats / per_slice);
      if (ks >= 2) p.ksplit = ks;
    }
    p.kb_per = (num_kb + p.ksplit - 1) / p.ksplit;
    p.ksplit = (num_kb + p.kb_per - 1) / p.kb_per;                // no empty slice
    p.part = a.splitk_ws;
  }
  const int total = tiles_mn * p.ksplit;
  const int grid = total < 2 * num_sms ? total : 2 * num_sms;
  auto kern = conv_tc_kernel<BN>;
  constexpr int smem = ConvSmem<BN>::TOTAL;
  static PerDeviceOnce attr;      // per template instantiation, per device
  DPMN_CUDA_TRY(attr.smem_attr(kern, smem));
  DPMN_CUDA_TRY(launch_pdl(kern, dim3(grid), dim3(CONV_THREADS), smem, st, maps[0], maps[1], maps[2], maps[3], map_w, p));
  DPMN_LAUNCH_CHECK();
  if (p.ksplit > 1) {
    SplitKReduceParams r;
    r.part = p.part; r.ksplit = p.ksplit; r.G = a.G; r.Cout = a.Cout; r.npix = p.npix; r.scale = a.scale; r.shift = a.shift;
    r.dst[0] = a.dst[0]; r.dst[1] = a.dst[1];
    const long long items = (long long)a.G * p.npix * (a.Cout / 8);
    const int blocks = (int)((items + 255) / 256 < 4 * num_sms ? (items + 255) / 256 : 4 * num_sms);
    conv_splitk_reduce_kernel<<<blocks, 256, 0, st>>>(r);
    DPMN_LAUNCH_CHECK();
  }
  return 0;
}

int launch_conv_tc(const ConvTcArgs& a, cudaStream_t st) {
  if (a.op_type != DT_F16 && a.op_type != DT_BF16) return -1;
  if (a.Cin % 8 || a.Cout % 8 || a.n_src < 1 || a.n_src > 4 || a.n_taps < 1 || a.n_taps > 16) return -2;
  if (a.P != 1 && a.P != 4) return -2;
  for (int di = 0; di < 2; ++di)
    if (a.dst[di].ptr && ((a.dst[di].ld % 8) || (a.dst[di].ch_off % 8) || (a.dst[di].ch_g_off % 8))) return -2;
  // N tile: small enough that the layer still spreads over the SMs, large enough to amortise the A box
  const int bw = next_pow2(a.Wm) < CBM ? next_pow2(a.Wm) : CBM;
  const int bh = next_pow2(a.Hm) < CBM / bw ? next_pow2(a.Hm) : CBM / bw;
  const int bb = CBM / (bw * bh);
  const long long pix_tiles = (long long)a.G * a.P * ((a.Wm + bw - 1) / bw) * ((a.Hm + bh - 1) / bh) * ((a.B + bb - 1) / bb);
  int bn = 128;
  if (a.Cout <= 32) bn = 32;
  else if (a.Cout <= 64) bn = 64;
  else if (pix_tiles * ((a.Cout + 127) / 128) < 2 * 148) bn = 64;
  switch (bn) {
    case 32: return launch_conv_bn<32>(a, st);
    case 64: return launch_conv_bn<64>(a, st);
    default: return launch_conv_bn<128>(a, st);
  }
}

}  // namespace dpmn
