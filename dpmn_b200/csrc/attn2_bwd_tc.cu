// Backward of the window attention core on tcgen05 (sm_100a)                          pgrm.py:230-249
//
// Same unit structure as attn2_tc.cu (128 window-major rows of one group = two 64-row halves, HC heads, M = 64 MMAs whose
// halves interleave in TMEM lanes), five contractions per half and head instead of two:
//   S  = Q K^T,  dP = dO V^T                         (scores recomputed, never stored by the forward)
//   P  = softmax(scale S + bias + shift mask),  Pd = P o M   (M: the forward's attn_drop mask, 1/(1-p) or 0)
//   dS = P o (dP o M - rowsum(P o dP o M))
//   dQ = scale dS K,   dK = scale dS^T Q,   dV = Pd^T dO,   d table[idx(n, m), head] += dS[n, m]
// The softmax warps hold one row of S and dP in registers (tcgen05.ld 32x32b), and write three 64 x 64 16-bit operand
// tiles per half and head: dS row-major (A of dQ), dS^T and Pd^T (A of dK / dV; written element-wise -- a thread owns one
// COLUMN of a transposed tile, so no two threads ever touch the same element).  K, Q and dO are the MN-major B operands
// of dQ, dK, dV straight from their TMA tiles, exactly as V is in the forward.  The bias-table gradient accumulates in
// shared memory (one atomic per score) and leaves the CTA once.  Outputs dq (rows, C) and dkv (rows, 2C) are fp32 in TOKEN
// order (the roll + window_partition of the forward undone in the epilogue).  One CTA per SM.
#include "common.cuh"
#include "kernels.h"
#include "tc_common.cuh"
#include <cstring>

namespace dpmn {

using namespace tc;

namespace {

constexpr int B2_ROWS = 128;
constexpr int B2_TAB_STRIDE = 232;          // >= (2*8-1)^2
constexpr int B2_TAB_FLOATS = 6 * B2_TAB_STRIDE;
constexpr int B2_STAGES = 2;

struct AttnBwd2Params {
  int L, C, G, hpg, cg, H, W;
  int ws[4], shift[4];
  int upg, tiles, nhc, total_units;
  int tpi; uint32_t tpi_magic;
  int nww[4]; uint32_t nww_magic[4]; int lastrow_w0[4]; int cut[4];
  unsigned long long rows_lo[4], cols_lo[4], all_keys[4];
  const float* table[4];
  float* d_table[4];
  float* dq; float* dkv;
  int fmt;
  float scale;                      // head_dim^-0.5
  float p_drop, keep_inv;
  unsigned long long seed;
  uint32_t site;
};

template <int D>
__device__ __forceinline__ uint64_t b2_desc_rowD(uint32_t smem_addr) {
  constexpr uint64_t layout = D == 16 ? 6 : 4;           // SWIZZLE_32B : SWIZZLE_64B
  constexpr uint64_t sbo = D == 16 ? 256 : 512;          // 8 rows
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(sbo >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= layout << 61;
  return d;
}

template <typename T> __device__ __forceinline__ uint32_t b2_pack2(float a, float b);
template <> __device__ __forceinline__ uint32_t b2_pack2<__half>(float a, float b) {
  const __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&h);
}
template <> __device__ __forceinline__ uint32_t b2_pack2<__nv_bfloat16>(float a, float b) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&h);
}
template <typename T> __device__ __forceinline__ uint16_t b2_bits(float a);
template <> __device__ __forceinline__ uint16_t b2_bits<__half>(float a) { return __half_as_ushort(__float2half_rn(a)); }
template <> __device__ __forceinline__ uint16_t b2_bits<__nv_bfloat16>(float a) { return __bfloat16_as_ushort(__float2bfloat16_rn(a)); }

template <int D, int HC>
struct B2Smem {
  static constexpr int TILE = B2_ROWS * D * 2;                  // one Q / K / V / dO head tile
  static constexpr int STAGE = HC * 4 * TILE;
  static constexpr int OP_TILE = 64 * 128;                      // 64 x 64 16-bit operand tile
  static constexpr int OP_BYTES = HC * 2 * 3 * OP_TILE;         // [head][half][dS, dS^T, Pd^T]
  static constexpr int BAR_BYTES = 256;
  static constexpr int ACC_FLOATS = 64 * 64;                    // per head: bias-gradient accumulators of the current window size
  static constexpr int TOTAL = B2_STAGES * STAGE + OP_BYTES + BAR_BYTES + 2 * B2_TAB_FLOATS * 4 + HC * ACC_FLOATS * 4 + 1024;
};

// One row of one head: recompute P, form dS, write the three operand tiles, accumulate the bias-table gradient.
template <int WS, typename T, bool DROP>
__device__ __forceinline__ void b2_row(uint32_t s_addr, int quarter, int lane, const float* tab, float* acc, int g, int w_idx,
                                       uint8_t* tiles_half, uint64_t* s_empty_bar, uint64_t* ds_empty_bar, uint32_t ds_empty_parity,
                                       bool full, const AttnBwd2Params& p, unsigned long long drop_base) {
  constexpr int N = WS * WS;
  constexpr int TW = 2 * WS - 1;
  const int r16 = lane & 15;
  const int row_half = quarter * 16 + r16;
  float s[N], dp[N];
  {
    constexpr uint32_t DP_OFF = 64;
    if constexpr (WS == 8) {
      uint32_t a0[32], a1[32], b0[32], b1[32];
      tmem_ld_32x32(s_addr, a0);
      tmem_ld_32x32(s_addr + 32u, a1);
      tmem_ld_32x32(s_addr + DP_OFF, b0);
      tmem_ld_32x32(s_addr + DP_OFF + 32u, b1);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        s[j] = __uint_as_float(a0[j]); s[32 + j] = __uint_as_float(a1[j]);
        dp[j] = __uint_as_float(b0[j]); dp[32 + j] = __uint_as_float(b1[j]);
      }
    } else {
      uint32_t a[32], b[32];
      tmem_ld_32x16(s_addr + (uint32_t)(quarter * 16), a);
      tmem_ld_32x16(s_addr + DP_OFF + (uint32_t)(quarter * 16), b);
      tmem_ld_wait();
      if constexpr (WS == 4) {
#pragma unroll
        for (int j = 0; j < 16; ++j) { s[j] = __uint_as_float(a[j]); dp[j] = __uint_as_float(b[j]); }
      } else {
        const int sel = r16 >> 2;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          uint32_t v = a[j], w = b[j];
          v = sel == 1 ? a[4 + j] : v; w = sel == 1 ? b[4 + j] : w;
          v = sel == 2 ? a[8 + j] : v; w = sel == 2 ? b[8 + j] : w;
          v = sel == 3 ? a[12 + j] : v; w = sel == 3 ? b[12 + j] : w;
          s[j] = __uint_as_float(v); dp[j] = __uint_as_float(w);
        }
      }
    }
  }
  tc_fence_before();
  __syncwarp();
  if (lane == 0) mbar_arrive(s_empty_bar);

  const int n = WS == 8 ? row_half : (WS == 4 ? r16 : (r16 & 3));
  const int i_n = n / WS, j_n = n % WS;
  unsigned long long km = 0ull;                              // shift mask, closed form (see attn2_tc.cu)
  if (p.shift[g] > 0) {
    const int cut = p.cut[g];
    if (w_idx >= p.lastrow_w0[g]) km |= i_n >= cut ? p.rows_lo[g] : (p.all_keys[g] ^ p.rows_lo[g]);
    const int nww = p.nww[g];
    const int wq = (int)__umulhi((uint32_t)w_idx, p.nww_magic[g]);
    if (w_idx - wq * nww == nww - 1) km |= j_n >= cut ? p.cols_lo[g] : (p.all_keys[g] ^ p.cols_lo[g]);
  }
  const int tb0 = (i_n + WS - 1) * TW + (j_n + WS - 1);
  const float* tb = tab + tb0;
  const float sc = p.scale;
  float mx = -INFINITY;
#pragma unroll
  for (int m = 0; m < N; ++m) {
    float v = fmaf(s[m], sc, tb[-((m / WS) * TW + (m % WS))]);
    const bool hit = m < 32 ? (((uint32_t)km >> m) & 1u) : (((uint32_t)(km >> 32) >> (m - 32)) & 1u);
    v += hit ? -100.0f : 0.0f;                               // pgrm.py:173
    s[m] = v;
    mx = fmaxf(mx, v);
  }
  float den = 0.f;
#pragma unroll
  for (int m = 0; m < N; ++m) { s[m] = __expf(s[m] - mx); den += s[m]; }
  const float inv = 1.0f / den;
  // the forward's attn_drop mask of this row, one bit per key (same counter hash and element index as attn2_tc.cu)
  unsigned long long kept = ~0ull;
  if constexpr (DROP) {
    kept = 0ull;
#pragma unroll
    for (int m = 0; m < N; ++m) {
      const float u = (float)(dpmn_hash32(p.seed, p.site, drop_base + (unsigned long long)m) >> 8) * (1.0f / 16777216.0f);
      kept |= u >= p.p_drop ? (1ull << m) : 0ull;
    }
  }
  float delta = 0.f;
#pragma unroll
  for (int m = 0; m < N; ++m) {
    s[m] *= inv;                                             // P
    if constexpr (DROP) dp[m] *= ((kept >> m) & 1ull) ? p.keep_inv : 0.f;   // dP o M
    delta = fmaf(s[m], dp[m], delta);
  }
  // dS (kept in dp[]), the bias-table gradient, and Pd = P o M (kept in s[])
  // d table[idx(n, m)] += dS[n, m]: a shared-memory atomic per score would cost ~2 LSU cycles per lane (measured: 130 of the
  // kernel's 143 us).  Instead the lanes that hold the same in-window position n first add up through shuffles, and ONE of them
  // adds into a plain accumulator acc[.][m][n] that no other thread touches (per warp for windows 2 / 4, where every quarter
  // holds the same n); the accumulators are folded into the table bins once per run of equal window sizes (b2_flush).
#pragma unroll
  for (int m = 0; m < N; ++m) {
    const float ds = s[m] * (dp[m] - delta);
    if constexpr (DROP) s[m] *= ((kept >> m) & 1ull) ? p.keep_inv : 0.f;
    dp[m] = ds;
    float a = ds + __shfl_xor_sync(0xffffffffu, ds, 16);      // the other half's row with the same n
    if constexpr (WS == 2) {
      a += __shfl_xor_sync(0xffffffffu, a, 4);
      a += __shfl_xor_sync(0xffffffffu, a, 8);
      if (lane < 4) acc[quarter * 16 + m * 4 + n] += a;
    } else if constexpr (WS == 4) {
      if (lane < 16) acc[quarter * 256 + m * 16 + n] += a;
    } else {
      if (lane < 16) acc[m * 64 + n] += a;
    }
  }

  // ---- operand tiles of this half: [0] dS row-major, [1] dS^T, [2] Pd^T; 64 x 64, 128-byte rows, 128B swizzle
  const int swz = row_half & 7;
  uint8_t* ds_row = tiles_half + (row_half >> 3) * 1024 + swz * 128;
  uint8_t* dst_t = tiles_half + 64 * 128;                    // dS^T tile
  uint8_t* pt_t = tiles_half + 2 * 64 * 128;                 // Pd^T tile
  // element (key row k, column row_half) of a transposed tile
  const uint32_t col_chunk = (uint32_t)(row_half >> 3), col_in = (uint32_t)(row_half & 7) * 2u;
  auto t_off = [&](int k) { return (uint32_t)((k >> 3) * 1024 + (k & 7) * 128) + (((col_chunk ^ (uint32_t)(k & 7))) << 4) + col_in; };
  const int key0 = WS == 8 ? 0 : (WS == 4 ? quarter * 16 : quarter * 16 + 4 * (r16 >> 2));   // first key of the row's block in the half
  mbar_wait(ds_empty_bar, ds_empty_parity);                  // the MMAs that last read these tiles have retired
  if constexpr (WS == 8) {
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      uint4 v;
      v.x = b2_pack2<T>(dp[8 * c + 0], dp[8 * c + 1]); v.y = b2_pack2<T>(dp[8 * c + 2], dp[8 * c + 3]);
      v.z = b2_pack2<T>(dp[8 * c + 4], dp[8 * c + 5]); v.w = b2_pack2<T>(dp[8 * c + 6], dp[8 * c + 7]);
      *reinterpret_cast<uint4*>(ds_row + ((c ^ swz) << 4)) = v;
    }
#pragma unroll
    for (int m = 0; m < 64; ++m) {
      const uint32_t o = t_off(m);
      *reinterpret_cast<uint16_t*>(dst_t + o) = b2_bits<T>(dp[m]);
      *reinterpret_cast<uint16_t*>(pt_t + o) = b2_bits<T>(s[m]);
    }
  } else {
    if (full) {
      // first unit with this window size: everything outside the row's own block is (re)written as zero -- the row of
      // the row-major tile and the COLUMN of the transposed ones both belong to this thread alone
#pragma unroll
      for (int c = 0; c < 8; ++c) *reinterpret_cast<uint4*>(ds_row + ((c ^ swz) << 4)) = make_uint4(0u, 0u, 0u, 0u);
#pragma unroll 8
      for (int k = 0; k < 64; ++k) {
        const uint32_t o = t_off(k);
        *reinterpret_cast<uint16_t*>(dst_t + o) = 0;
        *reinterpret_cast<uint16_t*>(pt_t + o) = 0;
      }
    }
#pragma unroll
    for (int m = 0; m < N; ++m) {
      const int k = key0 + m;
      *reinterpret_cast<uint16_t*>(ds_row + ((((uint32_t)(k >> 3)) ^ (uint32_t)swz) << 4) + (uint32_t)(k & 7) * 2u) = b2_bits<T>(dp[m]);
      const uint32_t o = t_off(k);
      *reinterpret_cast<uint16_t*>(dst_t + o) = b2_bits<T>(dp[m]);
      *reinterpret_cast<uint16_t*>(pt_t + o) = b2_bits<T>(s[m]);
    }
  }
}

// Fold one head's accumulators (window size ws) into its table-gradient bins and clear them.  Called by the head's 128
// threads between two named-barrier syncs; t = thread index within the head.
__device__ __forceinline__ void b2_flush(float* acc, float* dtab, int ws, int t) {
  const int N = ws * ws, tw = 2 * ws - 1;
  const int total = ws == 8 ? 64 * 64 : 4 * N * N;           // [m][n] for window 8, [quarter][m][n] otherwise
  for (int e = t; e < total; e += 128) {
    const int r = ws == 8 ? e : e % (N * N);
    const int m = r / N, n = r - m * N;
    const float v = acc[e];
    acc[e] = 0.f;
    if (v != 0.f) atomicAdd(dtab + (n / ws - m / ws + ws - 1) * tw + (n % ws - m % ws + ws - 1), v);
  }
}

template <int D, int HC, typename T, bool DROP>
// (10 warps put three on one SM sub-partition: 16384 / 96 = 170 registers per thread is the hardware ceiling for HC = 2)
__global__ void __launch_bounds__(64 + HC * 128, 1)
attn2_bwd_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_k,
                 const __grid_constant__ CUtensorMap map_v, const __grid_constant__ CUtensorMap map_do,
                 const __grid_constant__ AttnBwd2Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  using S = B2Smem<D, HC>;
  constexpr int THREADS = 64 + HC * 128;
  constexpr int STAGES = B2_STAGES;
  uint8_t* stages = smem;
  uint8_t* op_tiles = smem + STAGES * S::STAGE;
  uint64_t* bars = reinterpret_cast<uint64_t*>(op_tiles + S::OP_BYTES);
  uint64_t* full_bar = bars;                       // [STAGES]
  uint64_t* empty_bar = bars + STAGES;             // [STAGES]
  uint64_t* s_full = bars + 2 * STAGES;
  uint64_t* s_empty = s_full + 1;
  uint64_t* ds_full = s_full + 2;
  uint64_t* ds_empty = s_full + 3;
  uint64_t* o_full = s_full + 4;
  uint64_t* o_empty = s_full + 5;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(s_full + 6);
  float* s_tab = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + S::BAR_BYTES);   // [G][hpg][TAB_STRIDE]
  float* s_dtab = s_tab + B2_TAB_FLOATS;
  float* s_acc = s_dtab + B2_TAB_FLOATS;           // [HC][64 * 64]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr uint32_t TMEM_COLS = HC == 2 ? 512 : 256;
  constexpr uint32_t O_COL0 = HC * 128;            // per head: S at h*128, dP at h*128 + 64; outputs dQ | dK | dV from O_COL0

  auto decode = [&](int u, int& g, int& tile, int& hc) {
    g = (u >= p.upg) + (u >= 2 * p.upg) + (u >= 3 * p.upg);
    const int r = u - g * p.upg;
    if (p.nhc == 1) { tile = r; hc = 0; } else { hc = r / p.tiles; tile = r - hc * p.tiles; }   // head-chunk major: long runs of one (group, head)
  };
  auto tma_unit = [&](int u, int it) {
    int g, tile, hc;
    decode(u, g, tile, hc);
    const int stage = it % STAGES;
    uint8_t* st = stages + stage * S::STAGE;
    mbar_arrive_expect_tx(&full_bar[stage], S::STAGE);
#pragma unroll
    for (int h = 0; h < HC; ++h) {
      const int ch = (hc * HC + h) * D;
      tma_load_3d(st + (h * 4 + 0) * S::TILE, &map_q, &full_bar[stage], ch, tile * B2_ROWS, g);
      tma_load_3d(st + (h * 4 + 1) * S::TILE, &map_k, &full_bar[stage], ch, tile * B2_ROWS, g);
      tma_load_3d(st + (h * 4 + 2) * S::TILE, &map_v, &full_bar[stage], ch, tile * B2_ROWS, g);
      tma_load_3d(st + (h * 4 + 3) * S::TILE, &map_do, &full_bar[stage], g * p.cg + ch, tile * B2_ROWS, 0);
    }
  };
  if (warp == 0) {
    if (lane == 0) {
      tma_prefetch_desc(&map_q); tma_prefetch_desc(&map_k); tma_prefetch_desc(&map_v); tma_prefetch_desc(&map_do);
      for (int i = 0; i < STAGES; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
      mbar_init(s_full, 1); mbar_init(s_empty, 4 * HC);
      mbar_init(ds_full, 4 * HC); mbar_init(ds_empty, 1);
      mbar_init(o_full, 1); mbar_init(o_empty, 4 * HC);
      fence_barrier_init();
      int it = 0;
      for (int u = blockIdx.x; u < p.total_units && it < STAGES; u += gridDim.x, ++it) tma_unit(u, it);
    }
  } else if (warp == 1) {
    tmem_alloc(tmem_slot, TMEM_COLS);
  } else {
    for (int i = threadIdx.x - 64; i < B2_TAB_FLOATS; i += THREADS - 64) {
      const int gh = i / B2_TAB_STRIDE, e = i - gh * B2_TAB_STRIDE;
      float v = 0.f;
      if (gh < p.G * p.hpg) {
        const int g = gh / p.hpg, h = gh - g * p.hpg;
        const int tw = 2 * p.ws[g] - 1;
        if (e < tw * tw) v = p.table[g][e * p.hpg + h];
      }
      s_tab[i] = v;
      s_dtab[i] = 0.f;
    }
    for (int i = threadIdx.x - 64; i < HC * S::ACC_FLOATS; i += THREADS - 64) s_acc[i] = 0.f;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      int it = STAGES;
      for (int u = blockIdx.x + STAGES * gridDim.x; u < p.total_units; u += gridDim.x, ++it) {
        mbar_wait(&empty_bar[it % STAGES], (uint32_t)(((it / STAGES) & 1) ^ 1));
        tma_unit(u, it);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc_s = make_idesc_f16(p.fmt, 64, 64);
      const uint32_t idesc_o = make_idesc_f16(p.fmt, 64, D) | (1u << 16);   // B operand is MN-major
      auto issue_grads = [&](int j) {
        const int stage = j % STAGES;
        mbar_wait(ds_full, (uint32_t)(j & 1));
        mbar_wait(o_empty, (uint32_t)((j & 1) ^ 1));
        tc_fence_after();
        const uint8_t* st = stages + stage * S::STAGE;
#pragma unroll
        for (int h = 0; h < HC; ++h) {
          const uint32_t qa = smem_u32(st + (h * 4 + 0) * S::TILE), ka = smem_u32(st + (h * 4 + 1) * S::TILE);
          const uint32_t oa = smem_u32(st + (h * 4 + 3) * S::TILE);
#pragma unroll
          for (int t = 0; t < 2; ++t) {
            const uint32_t d0 = tmem_base + O_COL0 + (uint32_t)(h * 3 * D) + ((uint32_t)(16 * t) << 16);
            const uint32_t tl = smem_u32(op_tiles + (h * 2 + t) * 3 * S::OP_TILE);
            const uint64_t a_ds = make_smem_desc_sw128(tl), a_dst = make_smem_desc_sw128(tl + S::OP_TILE);
            const uint64_t a_pt = make_smem_desc_sw128(tl + 2 * S::OP_TILE);
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
              const uint32_t roff = (uint32_t)((t * 64 + ks * 16) * D * 2);
              umma_f16(d0, advance_desc_k(a_ds, ks), b2_desc_rowD<D>(ka + roff), idesc_o, ks ? 1u : 0u);            // dQ = dS K
              umma_f16(d0 + D, advance_desc_k(a_dst, ks), b2_desc_rowD<D>(qa + roff), idesc_o, ks ? 1u : 0u);       // dK = dS^T Q
              umma_f16(d0 + 2 * D, advance_desc_k(a_pt, ks), b2_desc_rowD<D>(oa + roff), idesc_o, ks ? 1u : 0u);    // dV = Pd^T dO
            }
          }
        }
        umma_commit(&empty_bar[stage]);
        umma_commit(ds_empty);
        umma_commit(o_full);
      };
      int it = 0;
      for (int u = blockIdx.x; u < p.total_units; u += gridDim.x, ++it) {
        const int stage = it % STAGES;
        mbar_wait(&full_bar[stage], (uint32_t)((it / STAGES) & 1));
        mbar_wait(s_empty, (uint32_t)((it & 1) ^ 1));
        tc_fence_after();
        const uint8_t* st = stages + stage * S::STAGE;
#pragma unroll
        for (int h = 0; h < HC; ++h) {
#pragma unroll
          for (int t = 0; t < 2; ++t) {
            const uint32_t roff = (uint32_t)(t * 64 * D * 2);
            const uint64_t dq_ = b2_desc_rowD<D>(smem_u32(st + (h * 4 + 0) * S::TILE) + roff);
            const uint64_t dk_ = b2_desc_rowD<D>(smem_u32(st + (h * 4 + 1) * S::TILE) + roff);
            const uint64_t dv_ = b2_desc_rowD<D>(smem_u32(st + (h * 4 + 2) * S::TILE) + roff);
            const uint64_t do_ = b2_desc_rowD<D>(smem_u32(st + (h * 4 + 3) * S::TILE) + roff);
            const uint32_t d_s = tmem_base + (uint32_t)(h * 128) + ((uint32_t)(16 * t) << 16);
#pragma unroll
            for (int k = 0; k < D / 16; ++k) {
              umma_f16(d_s, advance_desc_k(dq_, k), advance_desc_k(dk_, k), idesc_s, k ? 1u : 0u);            // S = Q K^T
              umma_f16(d_s + 64u, advance_desc_k(do_, k), advance_desc_k(dv_, k), idesc_s, k ? 1u : 0u);      // dP = dO V^T
            }
          }
        }
        umma_commit(s_full);
        if (it > 0) issue_grads(it - 1);
      }
      if (it > 0) issue_grads(it - 1);
    }
  } else {
    const int quarter = warp & 3;
    const int h = (warp - 2) >> 2;
    const int half = lane >> 4, r16 = lane & 15;
    const int row = half * 64 + quarter * 16 + r16;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(quarter * 32) << 16);
    auto epilogue = [&](int j, long long tok, int ch) {
      mbar_wait(o_full, (uint32_t)(j & 1));
      tc_fence_after();
      uint32_t o[3][32];
      const uint32_t col = O_COL0 + (uint32_t)(h * 3 * D);
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        if constexpr (D == 16) tmem_ld_32x16(lane_addr + col + (uint32_t)(k * D), o[k]);
        else tmem_ld_32x32(lane_addr + col + (uint32_t)(k * D), o[k]);
      }
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(o_empty);
      float* pq = p.dq + tok * p.C + ch;
      float* pk = p.dkv + tok * 2 * p.C + ch;
      float* pv = pk + p.C;
      const float sc = p.scale;
#pragma unroll
      for (int c = 0; c < D; c += 4) {
        *reinterpret_cast<float4*>(pq + c) = make_float4(__uint_as_float(o[0][c]) * sc, __uint_as_float(o[0][c + 1]) * sc,
                                                         __uint_as_float(o[0][c + 2]) * sc, __uint_as_float(o[0][c + 3]) * sc);
        *reinterpret_cast<float4*>(pk + c) = make_float4(__uint_as_float(o[1][c]) * sc, __uint_as_float(o[1][c + 1]) * sc,
                                                         __uint_as_float(o[1][c + 2]) * sc, __uint_as_float(o[1][c + 3]) * sc);
        *reinterpret_cast<float4*>(pv + c) = make_float4(__uint_as_float(o[2][c]), __uint_as_float(o[2][c + 1]),
                                                         __uint_as_float(o[2][c + 2]), __uint_as_float(o[2][c + 3]));
      }
    };
    int it = 0, last_ws = -1, last_gh = -1;
    long long tok_prev = 0;
    int ch_prev = 0;
    for (int u = blockIdx.x; u < p.total_units; u += gridDim.x, ++it) {
      int g, tile, hc;
      decode(u, g, tile, hc);
      const int ws = p.ws[g];
      const int b_img = (int)__umulhi((uint32_t)tile, p.tpi_magic);
      const int p_img = (tile - b_img * p.tpi) * B2_ROWS + row;
      const int head = hc * HC + h;
      const long long tok = (long long)b_img * p.L + window_row_to_token(p_img, p.H, p.W, ws, p.shift[g]).token;
      const int ch = g * p.cg + head * D;
      unsigned long long drop_base = 0ull;
      if constexpr (DROP)
        drop_base = ((((unsigned long long)b_img * p.G + g) * p.hpg + head) * p.L + p_img) * (unsigned long long)(ws * ws);
      float* acc = s_acc + h * S::ACC_FLOATS;
      if (last_gh >= 0 && last_gh != g * p.hpg + head) {       // the (group, head) changed: fold its accumulators into its bins
        asm volatile("bar.sync %0, 128;" ::"r"(1 + h) : "memory");
        b2_flush(acc, s_dtab + last_gh * B2_TAB_STRIDE, last_ws, threadIdx.x - 64 - h * 128);
        asm volatile("bar.sync %0, 128;" ::"r"(1 + h) : "memory");
      }
      last_gh = g * p.hpg + head;
      mbar_wait(s_full, (uint32_t)(it & 1));
      tc_fence_after();
      {
        const float* tab = s_tab + (g * p.hpg + head) * B2_TAB_STRIDE;
        const uint32_t s_addr = lane_addr + (uint32_t)(h * 128);
        uint8_t* th = op_tiles + (h * 2 + half) * 3 * S::OP_TILE;
        const uint32_t par = (uint32_t)((it & 1) ^ 1);
        const bool full = last_ws != ws;
        last_ws = ws;
        if (ws == 8) b2_row<8, T, DROP>(s_addr, quarter, lane, tab, acc, g, p_img >> 6, th, s_empty, ds_empty, par, full, p, drop_base);
        else if (ws == 4) b2_row<4, T, DROP>(s_addr, quarter, lane, tab, acc, g, p_img >> 4, th, s_empty, ds_empty, par, full, p, drop_base);
        else b2_row<2, T, DROP>(s_addr, quarter, lane, tab, acc, g, p_img >> 2, th, s_empty, ds_empty, par, full, p, drop_base);
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(ds_full);
      if (it > 0) epilogue(it - 1, tok_prev, ch_prev);
      tok_prev = tok;
      ch_prev = ch;
    }
    if (it > 0) epilogue(it - 1, tok_prev, ch_prev);
    if (last_gh >= 0) {
      asm volatile("bar.sync %0, 128;" ::"r"(1 + h) : "memory");
      b2_flush(s_acc + h * S::ACC_FLOATS, s_dtab + last_gh * B2_TAB_STRIDE, last_ws, threadIdx.x - 64 - h * 128);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
  // bias-table gradient of this CTA -> global (pgrm.py:234-238: the table rows gathered by relative_position_index)
  for (int i = threadIdx.x; i < p.G * p.hpg * B2_TAB_STRIDE; i += THREADS) {
    const int gh = i / B2_TAB_STRIDE, e = i - gh * B2_TAB_STRIDE;
    const int g = gh / p.hpg, h = gh - g * p.hpg;
    const int tw = 2 * p.ws[g] - 1;
    const float v = s_dtab[i];
    if (e < tw * tw && v != 0.f) atomicAdd(p.d_table[g] + e * p.hpg + h, v);
  }
}

template <int D, int HC, typename T, bool DROP>
int launch_b2(const AttnBwdTcArgs& a, cudaStream_t st) {
  AttnBwd2Params p;
  memset(&p, 0, sizeof(p));
  p.L = a.H * a.W; p.H = a.H; p.W = a.W; p.C = a.C; p.G = a.n_groups; p.hpg = a.heads_per_group; p.cg = a.C / a.n_groups;
  for (int g = 0; g < a.n_groups; ++g) {
    const int ws = a.window[g], N = ws * ws, cut = ws - a.shift[g];
    p.ws[g] = ws; p.shift[g] = a.shift[g]; p.table[g] = a.table[g]; p.d_table[g] = a.d_table[g]; p.cut[g] = cut;
    p.nww[g] = a.W / ws; p.nww_magic[g] = (uint32_t)((0x100000000ull + p.nww[g] - 1) / p.nww[g]);
    p.lastrow_w0[g] = p.L / N - p.nww[g];
    const unsigned long long rep = ws == 8 ? 0x0101010101010101ull : (ws == 4 ? 0x1111ull : 0x5ull);
    p.all_keys[g] = N == 64 ? ~0ull : ((1ull << N) - 1ull);
    p.rows_lo[g] = cut * ws >= 64 ? ~0ull : (1ull << (cut * ws)) - 1ull;
    p.cols_lo[g] = ((1ull << cut) - 1ull) * rep & p.all_keys[g];
  }
  p.tiles = a.B * p.L / B2_ROWS; p.nhc = a.heads_per_group / HC; p.upg = p.tiles * p.nhc; p.total_units = p.upg * p.G;
  p.tpi = p.L / B2_ROWS; p.tpi_magic = (uint32_t)((0x100000000ull + p.tpi - 1) / p.tpi);
  p.dq = a.dq; p.dkv = a.dkv; p.fmt = a.io_type == DT_BF16 ? 1 : 0; p.scale = 1.0f / sqrtf((float)D);
  p.p_drop = a.p_drop; p.keep_inv = a.p_drop > 0.f ? 1.0f / (1.0f - a.p_drop) : 1.0f; p.seed = a.seed; p.site = a.site;
  CUtensorMap maps[4];
  const long long rows = (long long)a.B * p.L;
  const void* bases[3] = {a.qw, a.kw, a.vw};
  const CUtensorMapSwizzle sw = D == 16 ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_64B;
  for (int i = 0; i < 3; ++i) {
    const uint64_t dims[3] = {(uint64_t)p.cg, (uint64_t)rows, (uint64_t)p.G};
    const uint64_t str[2] = {(uint64_t)p.cg * 2, (uint64_t)rows * p.cg * 2};
    const uint32_t box[3] = {(uint32_t)D, B2_ROWS, 1};
    if (int rc = make_tensor_map_16bit(&maps[i], bases[i], 3, dims, str, box, sw)) return rc;
  }
  {
    const uint64_t dims[3] = {(uint64_t)p.C, (uint64_t)rows, 1};
    const uint64_t str[2] = {(uint64_t)p.C * 2, (uint64_t)rows * p.C * 2};
    const uint32_t box[3] = {(uint32_t)D, B2_ROWS, 1};
    if (int rc = make_tensor_map_16bit(&maps[3], a.d_out16, 3, dims, str, box, sw)) return rc;
  }
  int num_sms = 0;
  DPMN_CUDA_TRY(current_device_sms(&num_sms));
  const int grid = p.total_units < num_sms ? p.total_units : num_sms;
  auto kern = attn2_bwd_kernel<D, HC, T, DROP>;
  constexpr int smem = B2Smem<D, HC>::TOTAL;
  static_assert(smem <= 232448, "shared memory per CTA");
  static PerDeviceOnce attr;
  DPMN_CUDA_TRY(attr.smem_attr(kern, smem));
  kern<<<grid, 64 + HC * 128, smem, st>>>(maps[0], maps[1], maps[2], maps[3], p);
  DPMN_LAUNCH_CHECK();
  return 0;
}

template <typename T>
int launch_b2_dtype(const AttnBwdTcArgs& a, cudaStream_t st) {
  const int d = a.C / a.n_groups / a.heads_per_group;
  const bool drop = a.p_drop > 0.f;
  if (d == 16 && a.heads_per_group % 2 == 0) return drop ? launch_b2<16, 2, T, true>(a, st) : launch_b2<16, 2, T, false>(a, st);
  if (d == 16) return drop ? launch_b2<16, 1, T, true>(a, st) : launch_b2<16, 1, T, false>(a, st);
  return drop ? launch_b2<32, 1, T, true>(a, st) : launch_b2<32, 1, T, false>(a, st);
}

}  // namespace

bool attn_bwd_tc_supported(const AttnBwdTcArgs& a) {
  if (a.io_type != DT_F16 && a.io_type != DT_BF16) return false;
  if (a.n_groups < 1 || a.n_groups > 4 || a.C % a.n_groups) return false;
  const int cg = a.C / a.n_groups;
  if (a.heads_per_group < 1 || cg % a.heads_per_group) return false;
  const int d = cg / a.heads_per_group;
  if (d != 16 && d != 32) return false;
  if ((a.H * a.W) % B2_ROWS) return false;
  if (a.n_groups * a.heads_per_group * B2_TAB_STRIDE > B2_TAB_FLOATS) return false;
  if (a.p_drop < 0.f || a.p_drop >= 1.f) return false;
  for (int g = 0; g < a.n_groups; ++g) {
    const int ws = a.window[g];
    if (ws != 2 && ws != 4 && ws != 8) return false;
    if (a.H % ws || a.W % ws) return false;
    if (a.shift[g] < 0 || a.shift[g] >= ws) return false;
  }
  return true;
}

int launch_window_attn_bwd_tc(const AttnBwdTcArgs& a, cudaStream_t st) {
  if (!attn_bwd_tc_supported(a)) return -2;
  return a.io_type == DT_F16 ? launch_b2_dtype<__half>(a, st) : launch_b2_dtype<__nv_bfloat16>(a, st);
}

}  // namespace dpmn
