// Fused window attention on tcgen05 (sm_100a)                                         pgrm.py:197-268
//
// Inputs are the projected q / k / v of one block, already in WINDOW-MAJOR row order per window group (the
// roll + window_partition of pgrm.py:209-225 is folded into the epilogue of the q / kv projection GEMMs, see
// gemm_tc.cu "scatter"): Qw, Kw, Vw are [G][B*L][cg] 16-bit, cg = heads_per_group * D.
//
// One work unit = 128 consecutive window-major rows (= 128/N whole windows) of one group, two heads:
//   TMA      six boxes {D ch, 128 rows}: Q_h, K_h, V_h for both heads (32B/64B swizzle), 3-stage ring
//   MMA #1   S_h[128 x 128] = Q_h K_h^T          (tcgen05.mma M128 N128 K=D, fp32 in TMEM; only the N x N
//            diagonal blocks are used: the tensor pipe is idle anyway, the kernel is HBM-bound)
//   softmax  4 warps, one thread per row: tcgen05.ld of the row's own window block, * scale + relative
//            position bias (closed-form index, table in smem) + shift mask (closed-form {0,-100}), exp,
//            normalise, write P_h as a 128B-swizzled K-major operand tile (zeros outside the window block)
//   MMA #2   O_h[128 x D] = P_h V_h              (V_h is the MN-major B operand straight from the TMA box)
//   epilogue tcgen05.ld O, 16-bit, one 2*D*2-byte run per row into out (B*L, C) -- rows stay window-major
//            (quirk 1: the reference never applies window_reverse, pgrm.py:249,263)
#include "common.cuh"
#include "kernels.h"
#include "tc_common.cuh"
#include <cstdlib>
#include <cstring>

namespace dpmn {

using namespace tc;

constexpr int AT_ROWS = 128;
constexpr int AT_STAGES = 3;
constexpr int AT_HC = 2;          // heads per unit
constexpr int AT_THREADS = 64 + AT_HC * 128;   // TMA warp, MMA warp, 4 softmax/epilogue warps per head

struct AttnTcParams {
  int B, H, W, L, C, G, hpg, cg;
  int ws[4], shift[4];
  int tiles, nhc, total_units;
  const float* table[4];
  void* out;
  int fmt;
  float scale;
};

// descriptors for 32-byte / 64-byte swizzled tiles whose rows are D*2 bytes (D = 16 -> SW32, D = 32 -> SW64)
template <int D>
__device__ __forceinline__ uint64_t make_desc_rowD(uint32_t smem_addr) {
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

template <int D>
struct AttnSmem {
  static constexpr int TILE = AT_ROWS * D * 2;                  // one Q/K/V head tile
  static constexpr int STAGE = AT_HC * 3 * TILE;
  static constexpr int P_TILE = AT_ROWS * 128 * 2;              // 32 KB per head
  static constexpr int PBUF = D == 16 ? 2 : 1;                  // P double-buffered where shared memory allows
  static constexpr int TOTAL = AT_STAGES * STAGE + PBUF * AT_HC * P_TILE + 1024 + 256 + 4096 * 4;
};

template <int WS>
__device__ __forceinline__ void select_block(const uint32_t (&r)[64], int lane, float (&s)[WS * WS]) {
  constexpr int N = WS * WS;
  if constexpr (N == 64) {
#pragma unroll
    for (int j = 0; j < 64; ++j) s[j] = __uint_as_float(r[j]);
  } else if constexpr (N == 16) {
    const bool hi = (lane & 16) != 0;
#pragma unroll
    for (int j = 0; j < 16; ++j) s[j] = __uint_as_float(hi ? r[16 + j] : r[j]);
  } else {
    const int q = lane >> 2;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      uint32_t v = r[j];
#pragma unroll
      for (int k = 1; k < 8; ++k) v = (q == k) ? r[4 * k + j] : v;
      s[j] = __uint_as_float(v);
    }
  }
}

// softmax of one row of one head and the write of its P row (16 x 16-byte chunks, 128B swizzle)
template <int WS, typename T>
__device__ __forceinline__ void softmax_row(uint32_t tmem_s, int quarter, int lane, int row, int n, const float* tab,
                                            float scale, bool masked, uint32_t rh_bits, uint32_t rw_bits,
                                            uint8_t* p_tile, uint64_t* s_empty_bar, uint64_t* p_empty_bar,
                                            uint32_t p_empty_parity, bool full_row, float& inv_out) {
  constexpr int N = WS * WS;
  constexpr int TW = 2 * WS - 1;
  // the row's window block sits at columns [blk*N, blk*N + N); a warp's 32 rows share one 32/64-column span
  const int col0 = N == 64 ? (row / 64) * 64 : quarter * 32;
  uint32_t r[64];
  {
    uint32_t lo[32];
    tmem_ld_32x32(tmem_s + ((uint32_t)(quarter * 32) << 16) + (uint32_t)col0, lo);
#pragma unroll
    for (int j = 0; j < 32; ++j) r[j] = lo[j];
    if constexpr (N == 64) {
      uint32_t hi[32];
      tmem_ld_32x32(tmem_s + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(col0 + 32), hi);
#pragma unroll
      for (int j = 0; j < 32; ++j) r[32 + j] = hi[j];
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j) r[32 + j] = 0u;
    }
    tmem_ld_wait();
  }
  // the scores are in registers: hand the S accumulator back so the next unit's QK^T overlaps this softmax
  tc_fence_before();
  __syncwarp();
  if (lane == 0) mbar_arrive(s_empty_bar);
  float s[N];
  select_block<WS>(r, lane, s);
  const int i_n = n / WS, j_n = n % WS;
  const uint32_t my_h = (rh_bits >> (2 * i_n)) & 3u, my_w = (rw_bits >> (2 * j_n)) & 3u;
  // Everything is in the log2 domain (`scale` and the table carry log2(e)), so each key costs FMA, max, add, ex2, add.
  // The shift mask only exists in border windows: a window whose rows / columns all carry the same region label
  // (the common case) skips it.
  constexpr uint32_t FIELDS = (1u << (2 * WS)) - 1u;
  const bool need_mask = masked && ((rh_bits != ((0x55555555u & FIELDS) * my_h)) || (rw_bits != ((0x55555555u & FIELDS) * my_w)));
  float mx = -INFINITY;
  if (need_mask) {
#pragma unroll
    for (int m = 0; m < N; ++m) {
      const int i_m = m / WS, j_m = m % WS;
      float v = fmaf(s[m], scale, tab[(i_n - i_m + WS - 1) * TW + (j_n - j_m + WS - 1)]);
      const bool diff = (((rh_bits >> (2 * i_m)) & 3u) != my_h) || (((rw_bits >> (2 * j_m)) & 3u) != my_w);
      v += diff ? -144.26950408889634f : 0.0f;          // -100 (pgrm.py:173) * log2(e)
      s[m] = v;
      mx = fmaxf(mx, v);
    }
  } else {
#pragma unroll
    for (int m = 0; m < N; ++m) {
      const int i_m = m / WS, j_m = m % WS;
      const float v = fmaf(s[m], scale, tab[(i_n - i_m + WS - 1) * TW + (j_n - j_m + WS - 1)]);
      s[m] = v;
      mx = fmaxf(mx, v);
    }
  }
  float den = 0.f;
#pragma unroll
  for (int m = 0; m < N; ++m) { s[m] = exp2f(s[m] - mx); den += s[m]; }
  // P is stored UNNORMALISED (values in (0, 1]); the epilogue scales the D outputs of the row by 1/den instead of
  // the N probabilities here
  const float inv = 1.0f;
  inv_out = 1.0f / den;
  // P row: keys [key0, key0 + N) of the tile hold the probabilities, everything else is 0.  The row is
  // 16 x 16-byte chunks (8 keys each), two 64-key k-blocks of 16 KB, chunk index XOR (row & 7) = 128B swizzle.
  const int key0 = (row / N) * N;
  constexpr int OWN = N >= 8 ? N / 8 : 1;               // chunks that carry data
  uint4 own[OWN];
  if constexpr (N >= 8) {
#pragma unroll
    for (int q = 0; q < OWN; ++q) {
      union { uint4 u; T h[8]; } pk;
#pragma unroll
      for (int e = 0; e < 8; ++e) pk.h[e] = from_f32<T>(s[q * 8 + e] * inv);
      own[q] = pk.u;
    }
  } else {
    union { uint4 u; T h[8]; } pk;
    pk.u = make_uint4(0u, 0u, 0u, 0u);
    const int half = (key0 >> 2) & 1;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const T v = from_f32<T>(s[e] * inv);
      if (half) pk.h[4 + e] = v; else pk.h[e] = v;
    }
    own[0] = pk.u;
  }
  const int c_own0 = key0 >> 3;
  uint8_t* row_base = p_tile + (size_t)row * 128;
  mbar_wait(p_empty_bar, p_empty_parity);          // the P*V that last read this P buffer has retired
  if (full_row) {
    // first unit of a window group in this buffer: the zeros outside the row's own window block are (re)written
#pragma unroll
    for (int c = 0; c < 16; ++c) {
      uint4 v = make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
      for (int q = 0; q < OWN; ++q)
        if (c == c_own0 + q) v = own[q];
      const int kb = c >> 3, cc = c & 7;
      *reinterpret_cast<uint4*>(row_base + kb * (AT_ROWS * 128) + ((cc ^ (row & 7)) << 4)) = v;
    }
  } else {
    // same group as the previous unit in this buffer: the block position of every row is unchanged
#pragma unroll
    for (int q = 0; q < OWN; ++q) {
      const int c = c_own0 + q;
      const int kb = c >> 3, cc = c & 7;
      *reinterpret_cast<uint4*>(row_base + kb * (AT_ROWS * 128) + ((cc ^ (row & 7)) << 4)) = own[q];
    }
  }
}

template <int D, typename T>
__global__ void __launch_bounds__(AT_THREADS, 1)
attn_tc_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_k,
               const __grid_constant__ CUtensorMap map_v, const __grid_constant__ AttnTcParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  using S = AttnSmem<D>;
  uint8_t* stages = smem;
  uint8_t* p_tiles = smem + AT_STAGES * S::STAGE;
  uint64_t* bars = reinterpret_cast<uint64_t*>(p_tiles + S::PBUF * AT_HC * S::P_TILE);
  uint64_t* full_bar = bars;                       // [AT_STAGES]
  uint64_t* empty_bar = bars + AT_STAGES;          // [AT_STAGES]
  uint64_t* s_full = bars + 2 * AT_STAGES;
  uint64_t* s_empty = s_full + 1;
  uint64_t* p_full = s_full + 2;                   // [2]
  uint64_t* p_empty = s_full + 4;                  // [2]
  uint64_t* o_full = s_full + 6;                   // [2]
  uint64_t* o_empty = s_full + 8;                  // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(s_full + 10);
  constexpr int PBUF = S::PBUF;
  float* s_tab = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + 256);   // [G][hpg][tab_stride]
  constexpr int TAB_STRIDE = 232;                  // >= (2*8-1)^2 = 225

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr uint32_t TMEM_COLS = 512;
  constexpr uint32_t O_COL0 = 256;

  for (int i = threadIdx.x; i < p.G * p.hpg * TAB_STRIDE; i += AT_THREADS) {
    const int e = i % TAB_STRIDE, gh = i / TAB_STRIDE;
    const int g = gh / p.hpg, h = gh - g * p.hpg;
    const int tw = 2 * p.ws[g] - 1;
    s_tab[i] = e < tw * tw ? p.table[g][e * p.hpg + h] * 1.4426950408889634f : 0.f;   // log2 domain
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_q); tma_prefetch_desc(&map_k); tma_prefetch_desc(&map_v);
    for (int i = 0; i < AT_STAGES; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
    mbar_init(s_full, 1); mbar_init(s_empty, 4 * AT_HC);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&p_full[i], 4 * AT_HC); mbar_init(&p_empty[i], 1);
      mbar_init(&o_full[i], 1); mbar_init(&o_empty[i], 4 * AT_HC);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

// units are ordered group-major, so a CTA's consecutive units share the window size (the P zero pattern)
#define DPMN_UNIT(u)                                   \
  const int hc = (u) % p.nhc;                          \
  const int tile = ((u) / p.nhc) % p.tiles;            \
  const int g = (u) / (p.nhc * p.tiles);

  if (warp == 0) {
    if (lane == 0) {
      int it = 0;
      for (int u = blockIdx.x; u < p.total_units; u += gridDim.x, ++it) {
        DPMN_UNIT(u)
        const int stage = it % AT_STAGES;
        const uint32_t phase = (it / AT_STAGES) & 1;
        mbar_wait(&empty_bar[stage], phase ^ 1);
        uint8_t* st = stages + stage * S::STAGE;
        mbar_arrive_expect_tx(&full_bar[stage], S::STAGE);
#pragma unroll
        for (int h = 0; h < AT_HC; ++h) {
          const int ch = (hc * AT_HC + h) * D;
          tma_load_3d(st + (h * 3 + 0) * S::TILE, &map_q, &full_bar[stage], ch, tile * AT_ROWS, g);
          tma_load_3d(st + (h * 3 + 1) * S::TILE, &map_k, &full_bar[stage], ch, tile * AT_ROWS, g);
          tma_load_3d(st + (h * 3 + 2) * S::TILE, &map_v, &full_bar[stage], ch, tile * AT_ROWS, g);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc_s = make_idesc_f16(p.fmt, AT_ROWS, 128);
      const uint32_t idesc_o = make_idesc_f16(p.fmt, AT_ROWS, D) | (1u << 16);   // B operand (V) is MN-major
      auto issue_pv = [&](int j) {
        const int stage = j % AT_STAGES;
        const int ob = j & 1;
        const int pb = j % PBUF;
        mbar_wait(&p_full[pb], (uint32_t)((j / PBUF) & 1));
        mbar_wait(&o_empty[ob], (uint32_t)(((j >> 1) & 1) ^ 1));
        tc_fence_after();
        const uint8_t* st = stages + stage * S::STAGE;
#pragma unroll
        for (int h = 0; h < AT_HC; ++h) {
          const uint32_t d_o = tmem_base + O_COL0 + (uint32_t)(ob * AT_HC * D + h * D);
          const uint32_t pa = smem_u32(p_tiles + (pb * AT_HC + h) * S::P_TILE);
          const uint32_t va = smem_u32(st + (h * 3 + 2) * S::TILE);
#pragma unroll
          for (int ks = 0; ks < 8; ++ks) {   // 8 k-steps of 16 keys
            const uint64_t da = advance_desc_k(make_smem_desc_sw128(pa + (ks >> 2) * (AT_ROWS * 128)), ks & 3);
            const uint64_t db = make_desc_rowD<D>(va + ks * 16 * D * 2);
            umma_f16(d_o, da, db, idesc_o, ks ? 1u : 0u);
          }
        }
        umma_commit(&empty_bar[stage]);
        umma_commit(&p_empty[pb]);
        umma_commit(&o_full[ob]);
      };
      int it = 0;
      for (int u = blockIdx.x; u < p.total_units; u += gridDim.x, ++it) {
        const int stage = it % AT_STAGES;
        mbar_wait(&full_bar[stage], (uint32_t)((it / AT_STAGES) & 1));
        mbar_wait(s_empty, (uint32_t)((it & 1) ^ 1));
        tc_fence_after();
        const uint8_t* st = stages + stage * S::STAGE;
#pragma unroll
        for (int h = 0; h < AT_HC; ++h) {
          const uint64_t dq = make_desc_rowD<D>(smem_u32(st + (h * 3 + 0) * S::TILE));
          const uint64_t dk = make_desc_rowD<D>(smem_u32(st + (h * 3 + 1) * S::TILE));
#pragma unroll
          for (int k = 0; k < D / 16; ++k)
            umma_f16(tmem_base + (uint32_t)(h * 128), advance_desc_k(dq, k), advance_desc_k(dk, k), idesc_s, k ? 1u : 0u);
        }
        umma_commit(s_full);
        if (it > 0) issue_pv(it - 1);
      }
      if (it > 0) issue_pv(it - 1);
    }
  } else {
    const int quarter = warp & 3;                       // TMEM lane quarter
    const int h = (warp - 2) >> 2;                      // this warp set's head within the unit
    const int row = quarter * 32 + lane;
    T* out = reinterpret_cast<T*>(p.out);
    auto epilogue = [&](int j, int u_prev, float inv_row) {
      const int hc = u_prev % p.nhc;
      const int tile = (u_prev / p.nhc) % p.tiles;
      const int g = u_prev / (p.nhc * p.tiles);
      const int ob = j & 1;
      mbar_wait(&o_full[ob], (uint32_t)((j >> 1) & 1));
      tc_fence_after();
      uint32_t o[32];
      const uint32_t col = O_COL0 + (uint32_t)(ob * AT_HC * D + h * D);
      if constexpr (D == 16) tmem_ld_32x16(tmem_base + ((uint32_t)(quarter * 32) << 16) + col, o);
      else tmem_ld_32x32(tmem_base + ((uint32_t)(quarter * 32) << 16) + col, o);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&o_empty[ob]);
      T* dst = out + ((long long)tile * AT_ROWS + row) * p.C + g * p.cg + (hc * AT_HC + h) * D;
#pragma unroll
      for (int c = 0; c < D; c += 8) {
        union { uint4 u; T hh[8]; } pk;
#pragma unroll
        for (int e = 0; e < 8; ++e) pk.hh[e] = from_f32<T>(__uint_as_float(o[c + e]) * inv_row);
        *reinterpret_cast<uint4*>(dst + c) = pk.u;
      }
    };
    int it = 0, u_prev = -1;
    int last_g[2] = {-1, -1};                           // group whose zero pattern each P buffer currently holds
    float inv_prev = 1.f, inv_cur = 1.f;                // 1 / softmax denominator of this thread's row
    for (int u = blockIdx.x; u < p.total_units; u += gridDim.x, ++it) {
      DPMN_UNIT(u)
      const int ws = p.ws[g], N = ws * ws, shift = p.shift[g];
      const int p_img = (tile * AT_ROWS + row) % p.L;
      const int w_idx = p_img / N, n = p_img - w_idx * N;
      const int nWw = p.W / ws;
      const int base_h = (w_idx / nWw) * ws, base_w = (w_idx % nWw) * ws;
      uint32_t rh_bits = 0, rw_bits = 0;
      if (shift > 0) {
        for (int i = 0; i < ws; ++i) {
          const int hp = base_h + i, wp = base_w + i;
          rh_bits |= (uint32_t)((hp >= p.H - ws) + (hp >= p.H - shift)) << (2 * i);
          rw_bits |= (uint32_t)((wp >= p.W - ws) + (wp >= p.W - shift)) << (2 * i);
        }
      }
      mbar_wait(s_full, (uint32_t)(it & 1));
      tc_fence_after();
      const int pb = it % PBUF;
      {
        const float* tab = s_tab + (g * p.hpg + hc * AT_HC + h) * TAB_STRIDE;
        const uint32_t ts = tmem_base + (uint32_t)(h * 128);
        uint8_t* pt = p_tiles + (pb * AT_HC + h) * S::P_TILE;
        uint64_t* pe = &p_empty[pb];
        const uint32_t pe_par = (uint32_t)(((it / PBUF) & 1) ^ 1);
        const bool full_row = last_g[pb] != g;
        last_g[pb] = g;
        if (ws == 8) softmax_row<8, T>(ts, quarter, lane, row, n, tab, p.scale, shift > 0, rh_bits, rw_bits, pt, s_empty, pe, pe_par, full_row, inv_cur);
        else if (ws == 4) softmax_row<4, T>(ts, quarter, lane, row, n, tab, p.scale, shift > 0, rh_bits, rw_bits, pt, s_empty, pe, pe_par, full_row, inv_cur);
        else softmax_row<2, T>(ts, quarter, lane, row, n, tab, p.scale, shift > 0, rh_bits, rw_bits, pt, s_empty, pe, pe_par, full_row, inv_cur);
      }
      fence_proxy_async();          // P (generic-proxy stores) must be visible to the tensor core's async proxy
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[pb]);
      if (it > 0) epilogue(it - 1, u_prev, inv_prev);
      inv_prev = inv_cur;
      u_prev = u;
    }
    if (it > 0) epilogue(it - 1, u_prev, inv_prev);
  }
#undef DPMN_UNIT

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// ---- host ---------------------------------------------------------------------------------------------------
template <int D, typename T>
static int launch_attn_tc_t(const AttnTcArgs& a, cudaStream_t st) {
  AttnTcParams p;
  memset(&p, 0, sizeof(p));
  p.B = a.B; p.H = a.H; p.W = a.W; p.L = a.H * a.W; p.C = a.C; p.G = a.n_groups; p.hpg = a.heads_per_group;
  p.cg = a.C / a.n_groups;
  for (int g = 0; g < a.n_groups; ++g) { p.ws[g] = a.window[g]; p.shift[g] = a.shift[g]; p.table[g] = a.table[g]; }
  p.tiles = a.B * p.L / AT_ROWS; p.nhc = a.heads_per_group / AT_HC; p.total_units = p.tiles * p.nhc * p.G;
  p.out = a.out; p.fmt = a.io_type == DT_BF16 ? 1 : 0; p.scale = 1.4426950408889634f / sqrtf((float)D);   // d^-0.5 * log2(e)
  CUtensorMap maps[3];
  const void* bases[3] = {a.qw, a.kw, a.vw};
  const long long rows = (long long)a.B * p.L;
  for (int i = 0; i < 3; ++i) {
    const uint64_t dims[3] = {(uint64_t)p.cg, (uint64_t)rows, (uint64_t)p.G};
    const uint64_t str[2] = {(uint64_t)p.cg * 2, (uint64_t)rows * p.cg * 2};
    const uint32_t box[3] = {(uint32_t)D, AT_ROWS, 1};
    int rc = make_tensor_map_16bit(&maps[i], bases[i], 3, dims, str, box,
                                   D == 16 ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_64B);
    if (rc) return rc;
  }
  int num_sms = 0;
  DPMN_CUDA_TRY(current_device_sms(&num_sms));
  const int grid = p.total_units < num_sms ? p.total_units : num_sms;
  auto kern = attn_tc_kernel<D, T>;
  constexpr int smem = AttnSmem<D>::TOTAL;
  static PerDeviceOnce attr;      // per template instantiation, per device
  DPMN_CUDA_TRY(attr.smem_attr(kern, smem));
  kern<<<grid, AT_THREADS, smem, st>>>(maps[0], maps[1], maps[2], p);
  DPMN_LAUNCH_CHECK();
  return 0;
}

static bool attn_v1_forced() {
  static const bool v1 = getenv("DPMN_ATTN_V1") && atoi(getenv("DPMN_ATTN_V1")) != 0;
  return v1;
}

static bool attn_v1_supported(const AttnTcArgs& a) {
  if (a.p_drop > 0.f) return false;
  if (a.io_type != DT_F16 && a.io_type != DT_BF16) return false;
  if (a.n_groups < 1 || a.n_groups > 4 || a.C % a.n_groups) return false;
  const int cg = a.C / a.n_groups;
  if (a.heads_per_group % AT_HC || cg % a.heads_per_group) return false;
  const int d = cg / a.heads_per_group;
  if (d != 16 && d != 32) return false;
  if (cg % 32) return false;                       // the projection epilogue scatters 32-column chunks per group
  if ((a.H * a.W) % AT_ROWS) return false;
  if (a.n_groups * a.heads_per_group * 232 > 4096) return false;
  for (int g = 0; g < a.n_groups; ++g) {
    const int ws = a.window[g];
    if (ws != 2 && ws != 4 && ws != 8) return false;
    if (a.H % ws || a.W % ws) return false;
  }
  return true;
}

bool attn_tc_supported(const AttnTcArgs& a) {
  if (!attn_v1_forced() && attn2_tc_supported(a)) return true;
  return attn_v1_supported(a);
}

int launch_window_attn_tc(const AttnTcArgs& a, cudaStream_t st) {
  if (!attn_v1_forced() && attn2_tc_supported(a)) return launch_window_attn2_tc(a, st);
  if (!attn_v1_supported(a)) return -2;
  const int d = a.C / a.n_groups / a.heads_per_group;
  if (d == 16) {
    return a.io_type == DT_F16 ? launch_attn_tc_t<16, __half>(a, st) : launch_attn_tc_t<16, __nv_bfloat16>(a, st);
  }
  return a.io_type == DT_F16 ? launch_attn_tc_t<32, __half>(a, st) : launch_attn_tc_t<32, __nv_bfloat16>(a, st);
}

}  // namespace dpmn
