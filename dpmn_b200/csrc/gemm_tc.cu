// tcgen05 GEMM for sm_100a:  C[z][m, n] = epi( sum_k A[z][m, k] * B[z][n, k] )
//
// Both operands are K-contiguous 16-bit (fp16 or bf16) matrices in HBM.  Persistent, warp-specialised:
//   warp 0   TMA producer: 128 x 64 (A) and BN x 64 (B) boxes, 128B-swizzled, 4-stage mbarrier ring
//   warp 1   MMA issuer: one thread issues tcgen05.mma (M = 128, N = BN, K = 16) into a TMEM accumulator;
//            two accumulator stages so the epilogue of tile i overlaps the MMAs of tile i+1
//   warps 2-5 epilogue: tcgen05.ld (32 lanes x 32 columns), bias / GELU / residual / column-sum, 16-byte stores
// Every nn.Linear and the 1x1 conv of the PGRM (pgrm.py:30,37,39,82,188,194) run through this kernel in the
// fp16 / bf16 modes.
#include "common.cuh"
#include "kernels.h"
#include "tc_common.cuh"
#include <cstdlib>
#include <cstring>

namespace dpmn {

using namespace tc;

constexpr int TBM = 128;       // tile rows (UMMA M)
constexpr int TBK = 64;        // k-block: 64 x 16-bit = one 128-byte swizzle row
constexpr int TC_THREADS = 192;
// Two CTAs per SM: the epilogue (4 warps, latency-bound loads/stores) of one CTA overlaps the other CTA's work.
// Shared memory (<= 113 KB per CTA) and TMEM (<= 256 of 512 columns per CTA) are sized per N tile.
__device__ __forceinline__ uint32_t pack_f16x2(float a, float b) {
  const __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&h);
}
__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&h);
}

template <int BN> struct TcCfg {
  static constexpr int STAGES = BN <= 128 ? 3 : 2;
  static constexpr int ACC = BN <= 128 ? 2 : 1;        // TMEM accumulator stages
};

struct GemmTcParams {
  int M, N, K, batch;
  int m_tiles, n_tiles;
  int a_zmul, b_zmul;            // 0 when the operand is shared by every batch entry (weights)
  int fmt;                       // 0 fp16, 1 bf16 (operands)
  int out_type;                  // DType of C
  void* C; long long c_bs; int ldc;
  const float* bias; long long bias_bs; int bias_mode;   // 0 none, 1 per n, 2 per m
  int act;                       // 0 none, 1 GELU
  const float* residual;         // fp32, indexed like C (may alias C when C is fp32)
  float* colsum;                 // if set: no C store; colsum[z][m_tile*4 + quarter][n] = sum over 32 rows
  void* out16;                   // fp32 C only: second copy of C in the operand type (fmt), same indexing
  int ln_mode;                   // second output of the finished row: 0 none, 1 LayerNorm(ln_w, ln_b), 2 copy
  const float *ln_w, *ln_b;
  void* ln_out; int ln_type;     // (batch*M, N) rows of ln_type; requires N == BN <= 128
  int scatter, sc_C, sc_G, sc_H, sc_W, sc_cg;
  int sc_ws[4], sc_shift[4];
  int sc_pow2, sc_lW, sc_lL, sc_lws[4];   // W, H*W and every window size are powers of two (every DPMN geometry): shifts, no divisions
  void* sc_dst[2];
};

template <int BN>
struct TcSmem {
  static constexpr int A_BYTES = TBM * TBK * 2;
  static constexpr int B_BYTES = BN * TBK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int TOTAL = TcCfg<BN>::STAGES * STAGE_BYTES + 1024 /*alignment slack*/ + 256 /*barriers*/;
};

// EW = epilogue warps (4, or 8: two warps per TMEM lane quarter take alternate CW-column chunks -- the epilogue, not the
// MMA, bounds these K = 96..384 GEMMs, so more epilogue warps per SM hide more tcgen05.ld / store latency).
template <int BN, typename OutT, bool LN, int EW = 4>
__global__ void __launch_bounds__(64 + 32 * EW, LN ? 1 : 2)
gemm_tc_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
               const __grid_constant__ CUtensorMap map_a2, const __grid_constant__ CUtensorMap map_b2,
               const __grid_constant__ GemmTcParams p0, const __grid_constant__ GemmTcParams p1, const int tiles0, const int tiles1) {
  // Two independent problems may share one launch (tiles [0, tiles0) belong to p0 / map_a / map_b, the rest to p1 / map_a2 /
  // map_b2): the q and kv projections of a block read different inputs but are otherwise the same small GEMM, and one
  // persistent launch over both fills the machine better than two (pgrm.py:188,194).  tiles1 == 0: a single problem.
  constexpr int CW = EW == 8 ? 16 : 32;         // columns per epilogue chunk (8 warps: fewer live registers per thread)
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment for the 128B swizzle atoms
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  using S = TcSmem<BN>;
  constexpr int TSTAGES = TcCfg<BN>::STAGES;
  constexpr int ACC = TcCfg<BN>::ACC;
  uint8_t* tiles = smem;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + TSTAGES * S::STAGE_BYTES);
  uint64_t* full_bar = bars;                    // [TSTAGES]
  uint64_t* empty_bar = bars + TSTAGES;         // [TSTAGES]
  uint64_t* tmem_full = bars + 2 * TSTAGES;     // [2]
  uint64_t* tmem_empty = bars + 2 * TSTAGES + 2;// [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * TSTAGES + 4);

  constexpr uint32_t TMEM_COLS = (ACC * BN <= 32) ? 32 : (ACC * BN <= 64) ? 64 : (ACC * BN <= 128) ? 128 : 256;
  static_assert(ACC * BN <= 256, "two co-resident CTAs share the 512 TMEM columns");

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int total_tiles = tiles0 + tiles1;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_a);
    tma_prefetch_desc(&map_b);
    if (tiles1 > 0) { tma_prefetch_desc(&map_a2); tma_prefetch_desc(&map_b2); }
    for (int i = 0; i < TSTAGES; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tmem_full[i], 1); mbar_init(&tmem_empty[i], EW); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();        // everything above overlapped the predecessor's tail; its outputs are visible from here on
  pdl_trigger();

  if (warp == 0) {
    // ================= TMA producer =================
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
        const bool second = t >= tiles0;
        const GemmTcParams& p = second ? p1 : p0;
        const CUtensorMap* ma = second ? &map_a2 : &map_a;
        const CUtensorMap* mb = second ? &map_b2 : &map_b;
        const int tt = second ? t - tiles0 : t;
        const int n_blk = tt % p.n_tiles;
        const int m_blk = (tt / p.n_tiles) % p.m_tiles;
        const int z = tt / (p.n_tiles * p.m_tiles);
        const int num_kb = (p.K + TBK - 1) / TBK;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = tiles + stage * S::STAGE_BYTES;
          uint8_t* sb = sa + S::A_BYTES;
          mbar_arrive_expect_tx(&full_bar[stage], S::STAGE_BYTES);
          tma_load_3d(sa, ma, &full_bar[stage], kb * TBK, m_blk * TBM, z * p.a_zmul);
          tma_load_3d(sb, mb, &full_bar[stage], kb * TBK, n_blk * BN, z * p.b_zmul);
          if (++stage == TSTAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    if (lane == 0) {
      const uint32_t idesc = make_idesc_f16(p0.fmt, TBM, BN);
      int stage = 0; uint32_t phase = 0;
      int acc = 0; uint32_t acc_phase = 0;
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
        const GemmTcParams& p = t >= tiles0 ? p1 : p0;
        const int num_kb = (p.K + TBK - 1) / TBK;
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(tiles + stage * S::STAGE_BYTES);
          const uint64_t da = make_smem_desc_sw128(sa);
          const uint64_t db = make_smem_desc_sw128(sa + S::A_BYTES);
          const int k_left = p.K - kb * TBK;
          const int ksteps = k_left >= TBK ? TBK / 16 : (k_left + 15) / 16;
          for (int k = 0; k < ksteps; ++k)
            umma_f16(d_tmem, advance_desc_k(da, k), advance_desc_k(db, k), idesc, (kb | k) ? 1u : 0u);
          umma_commit(&empty_bar[stage]);            // frees the smem stage once these MMAs have read it
          if (++stage == TSTAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit(&tmem_full[acc]);                // accumulator complete -> epilogue
        if (++acc == ACC) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    // ================= epilogue (warps 2..2+EW-1) =================
    const int quarter = warp & 3;                    // TMEM lane quarter this warp may access
    const int chunk_first = EW == 8 ? ((warp - 2) >> 2) * CW : 0;
    constexpr int chunk_step = EW == 8 ? 2 * CW : CW;
    int acc = 0; uint32_t acc_phase = 0;
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
      const bool second = t >= tiles0;
      const GemmTcParams& p = second ? p1 : p0;
      const int tt = second ? t - tiles0 : t;
      int n_blk = 0, mz = tt;
      if (p.n_tiles > 1) { mz = tt / p.n_tiles; n_blk = tt - mz * p.n_tiles; }
      int z = 0, m_blk = mz;
      if (p.batch > 1) { z = mz / p.m_tiles; m_blk = mz - z * p.m_tiles; }
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      const int m = m_blk * TBM + quarter * 32 + lane;
      const bool row_ok = m < p.M;
      const float* bias = p.bias ? p.bias + (long long)z * p.bias_bs : nullptr;
      const float bias_m = (p.bias_mode == 2 && row_ok) ? bias[m] : 0.f;
      constexpr bool kCanLn = LN;                      // the whole output row passes through one thread
      float rowbuf[kCanLn ? BN : 1];
      const int m_ld = row_ok ? m : (p.M - 1);          // clamped row: loads are issued unconditionally
      // window scatter: everything that depends on the row only -- image, token, its window-major row in every group (roll +
      // window_partition, pgrm.py:209-225) -- once per tile instead of once per 16-column chunk (the per-chunk form was
      // ~1000 instructions per thread and tile, 10 M warp instructions per launch: profiles/r02_ncu_full_block_v4.txt id 1)
      long long sc_rowbase[4] = {0, 0, 0, 0};           // element offset of (group g, this row) in a scatter destination
      int sc_which0 = 0, sc_nn0 = 0;
      if (p.scatter) {
        const int nb = n_blk * BN;
        if (p.sc_pow2) {
          const int b = m_ld >> p.sc_lL, token = m_ld & ((1 << p.sc_lL) - 1);
          const int ho = token >> p.sc_lW, wo = token & (p.sc_W - 1);
          const int nimg = p.M >> p.sc_lL;
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            if (g < p.sc_G) {
              const int lws = p.sc_lws[g], sh = p.sc_shift[g];
              int hp = ho - sh; if (hp < 0) hp += p.sc_H;
              int wp = wo - sh; if (wp < 0) wp += p.sc_W;
              const int wm = (1 << lws) - 1;
              const int prow = ((((hp >> lws) << (p.sc_lW - lws)) + (wp >> lws)) << (2 * lws)) + ((hp & wm) << lws) + (wp & wm);
              sc_rowbase[g] = ((((long long)g * nimg + b) << p.sc_lL) + prow) * p.sc_cg;
            }
          }
          sc_which0 = nb >= p.sc_C ? (nb >= 2 * p.sc_C ? nb / p.sc_C : 1) : 0;
        } else {
          const int L = p.sc_H * p.sc_W;
          const int b = m_ld / L, token = m_ld - b * L;
          const int ho = token / p.sc_W, wo = token - ho * p.sc_W;
          const int nimg = p.M / L;
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            if (g < p.sc_G) {
              const int ws = p.sc_ws[g], sh = p.sc_shift[g];
              int hp = ho - sh; if (hp < 0) hp += p.sc_H;
              int wp = wo - sh; if (wp < 0) wp += p.sc_W;
              const int hq = hp / ws, wq = wp / ws;
              const int prow = (hq * (p.sc_W / ws) + wq) * ws * ws + (hp - hq * ws) * ws + (wp - wq * ws);
              sc_rowbase[g] = (((long long)g * nimg + b) * L + prow) * p.sc_cg;
            }
          }
          sc_which0 = nb / p.sc_C;
        }
        sc_nn0 = nb - sc_which0 * p.sc_C;
      }
      auto do_chunk = [&](const int c0) {
        const int n0 = n_blk * BN + c0;
        // (1) issue every global load of this chunk first, unpredicated (columns clamped into range), so that the
        //     in-order issue of the warp does not serialise one L2/DRAM round trip per float4
        float4 bb[CW / 4], rr[CW / 4];
        const bool full = n0 + CW <= p.N;                 // whole chunk inside the matrix: no clamps / predicates
        if (p.bias_mode == 1) {
          if (full) {
#pragma unroll
            for (int j = 0; j < CW / 4; ++j) bb[j] = *reinterpret_cast<const float4*>(bias + n0 + 4 * j);
          } else {
#pragma unroll
            for (int j = 0; j < CW / 4; ++j) {
              const int nc = min(n0 + 4 * j, p.N - 4);
              bb[j] = *reinterpret_cast<const float4*>(bias + nc);
            }
          }
        }
        const long long off = (long long)z * p.c_bs + (long long)m * p.ldc + n0;
        if (p.residual != nullptr) {
          const long long off_ld = (long long)z * p.c_bs + (long long)m_ld * p.ldc;
#pragma unroll
          for (int j = 0; j < CW / 4; ++j) {
            const int nc = min(n0 + 4 * j, p.N - 4);
            rr[j] = *reinterpret_cast<const float4*>(p.residual + off_ld + nc);
          }
        }
        // (2) accumulator chunk
        uint32_t r[32];
        if constexpr (CW == 32) tmem_ld_32x32(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * BN + c0), r);
        else tmem_ld_32x16(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * BN + c0), r);
        tmem_ld_wait();
        float v[32];
        if (p.bias_mode == 2) {
#pragma unroll
          for (int j = 0; j < CW; ++j) v[j] = __uint_as_float(r[j]) + bias_m;
        } else {
#pragma unroll
          for (int j = 0; j < CW; ++j) v[j] = __uint_as_float(r[j]);
        }
        if (p.bias_mode == 1) {
#pragma unroll
          for (int j = 0; j < CW / 4; ++j) { v[4 * j] += bb[j].x; v[4 * j + 1] += bb[j].y; v[4 * j + 2] += bb[j].z; v[4 * j + 3] += bb[j].w; }
        }
        if (p.act == 1) {
#pragma unroll
          for (int j = 0; j < CW; ++j) v[j] = gelu_fast(v[j]);
        }
        if constexpr (CW == 32) if (p.colsum != nullptr) {
          // butterfly transpose-reduce over the warp's 32 rows: lane l ends with the sum of column l
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = row_ok ? v[j] : 0.f;
#pragma unroll
          for (int s = 16; s >= 1; s >>= 1) {
#pragma unroll
            for (int i = 0; i < s; ++i) {
              const bool up = (lane & s) != 0;
              const float send = up ? v[i] : v[i + s];
              const float keep = up ? v[i + s] : v[i];
              v[i] = keep + __shfl_xor_sync(0xffffffffu, send, s);
            }
          }
          if (n0 + lane < p.N)
            p.colsum[((long long)z * (p.m_tiles * 4) + m_blk * 4 + quarter) * p.N + n0 + lane] = v[0];
          return;
        }
        if (p.scatter) {
          if constexpr (sizeof(OutT) == 2) {
            if (row_ok && n0 < p.N) {
              int which = sc_which0, nn = sc_nn0 + c0;
              while (nn >= p.sc_C) { nn -= p.sc_C; ++which; }
              const int g = (nn >= p.sc_cg) + (nn >= 2 * p.sc_cg) + (nn >= 3 * p.sc_cg);
              const int col = nn - g * p.sc_cg;
              const long long rb = g == 0 ? sc_rowbase[0] : (g == 1 ? sc_rowbase[1] : (g == 2 ? sc_rowbase[2] : sc_rowbase[3]));
              OutT* dst = reinterpret_cast<OutT*>(p.sc_dst[which]) + rb + col;
#pragma unroll
              for (int j = 0; j < CW; j += 8) {
                union { uint4 u; OutT h[8]; } pk;
#pragma unroll
                for (int e = 0; e < 8; ++e) pk.h[e] = from_f32<OutT>(v[j + e]);
                *reinterpret_cast<uint4*>(dst + j) = pk.u;
              }
            }
          }
          return;
        }
        if (p.residual != nullptr) {
#pragma unroll
          for (int j = 0; j < CW / 4; ++j) { v[4 * j] += rr[j].x; v[4 * j + 1] += rr[j].y; v[4 * j + 2] += rr[j].z; v[4 * j + 3] += rr[j].w; }
        }
        if constexpr (kCanLn) {
          if (p.ln_mode != 0) {
#pragma unroll
            for (int j = 0; j < CW; ++j) rowbuf[c0 + j] = v[j];   // static index: the chunk loop is fully unrolled
          }
        }
        if (!row_ok) return;
        if constexpr (sizeof(OutT) == 4) {
          float* dst = reinterpret_cast<float*>(p.C) + off;
#pragma unroll
          for (int j = 0; j < CW; j += 4)
            if (full || n0 + j < p.N) *reinterpret_cast<float4*>(dst + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
          if (p.out16 != nullptr) {                           // 16-bit copy of the same values (N % 8 == 0, ldc % 8 == 0: checked on the host)
#pragma unroll
            for (int j = 0; j < CW; j += 8) {
              if (full || n0 + j < p.N) {
                uint4 pk;
                if (p.fmt == 0) {
                  pk.x = pack_f16x2(v[j], v[j + 1]); pk.y = pack_f16x2(v[j + 2], v[j + 3]);
                  pk.z = pack_f16x2(v[j + 4], v[j + 5]); pk.w = pack_f16x2(v[j + 6], v[j + 7]);
                } else {
                  pk.x = pack_bf16x2(v[j], v[j + 1]); pk.y = pack_bf16x2(v[j + 2], v[j + 3]);
                  pk.z = pack_bf16x2(v[j + 4], v[j + 5]); pk.w = pack_bf16x2(v[j + 6], v[j + 7]);
                }
                *reinterpret_cast<uint4*>(reinterpret_cast<uint16_t*>(p.out16) + off + j) = pk;
              }
            }
          }
        } else {
          OutT* dst = reinterpret_cast<OutT*>(p.C) + off;
#pragma unroll
          for (int j = 0; j < CW; j += 8) {
            if (full || n0 + j < p.N) {
              union { uint4 u; OutT h[8]; } pk;
#pragma unroll
              for (int e = 0; e < 8; ++e) pk.h[e] = from_f32<OutT>(v[j + e]);
              *reinterpret_cast<uint4*>(dst + j) = pk.u;
            }
          }
        }
      };
      if constexpr (kCanLn) {
#pragma unroll
        for (int c0 = 0; c0 < BN; c0 += 32) do_chunk(c0);
      } else {
#pragma unroll 1
        for (int c0 = chunk_first; c0 < BN; c0 += chunk_step) do_chunk(c0);
      }
      // the accumulator stage is free as soon as its last chunk has been read
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[acc]);
      if constexpr (kCanLn) {
        // (optional) second output computed from the finished row: LayerNorm of it (the next consumer's norm,
        // pgrm.py:322-323,330) or a plain narrow copy -- saves a full read+write pass over the token stream
        if (p.ln_mode != 0 && row_ok) {
          float mu = 0.f, rstd = 1.f;
          if (p.ln_mode == 1) {
            float s1 = 0.f;
#pragma unroll
            for (int j = 0; j < BN; ++j) s1 += rowbuf[j];
            mu = s1 * (1.0f / BN);
            float s2 = 0.f;
#pragma unroll
            for (int j = 0; j < BN; ++j) { const float dlt = rowbuf[j] - mu; s2 = fmaf(dlt, dlt, s2); }
            rstd = rsqrtf(s2 * (1.0f / BN) + 1e-5f);
          }
          const long long roff = ((long long)z * p.M + m) * BN;
#pragma unroll
          for (int j = 0; j < BN; j += 8) {
            float y[8];
            if (p.ln_mode == 1) {
              const float4 w0 = *reinterpret_cast<const float4*>(p.ln_w + j), w1 = *reinterpret_cast<const float4*>(p.ln_w + j + 4);
              const float4 b0 = *reinterpret_cast<const float4*>(p.ln_b + j), b1 = *reinterpret_cast<const float4*>(p.ln_b + j + 4);
              const float ww[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
              const float bq[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
              for (int e = 0; e < 8; ++e) y[e] = (rowbuf[j + e] - mu) * rstd * ww[e] + bq[e];
            } else {
#pragma unroll
              for (int e = 0; e < 8; ++e) y[e] = rowbuf[j + e];
            }
            if (p.ln_type == DT_F32) {
              float* dst = reinterpret_cast<float*>(p.ln_out) + roff + j;
              *reinterpret_cast<float4*>(dst) = make_float4(y[0], y[1], y[2], y[3]);
              *reinterpret_cast<float4*>(dst + 4) = make_float4(y[4], y[5], y[6], y[7]);
            } else if (p.ln_type == DT_F16) {
              union { uint4 u; __half h[8]; } pk;
#pragma unroll
              for (int e = 0; e < 8; ++e) pk.h[e] = __float2half_rn(y[e]);
              *reinterpret_cast<uint4*>(reinterpret_cast<__half*>(p.ln_out) + roff + j) = pk.u;
            } else {
              union { uint4 u; __nv_bfloat16 h[8]; } pk;
#pragma unroll
              for (int e = 0; e < 8; ++e) pk.h[e] = __float2bfloat16_rn(y[e]);
              *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.ln_out) + roff + j) = pk.u;
            }
          }
        }
      }
      if (++acc == ACC) { acc = 0; acc_phase ^= 1; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// ---- host ------------------------------------------------------------------------------------------------
PFN_encodeTiled get_encode_tiled() {
  static PFN_encodeTiled fn = nullptr;
  if (fn) return fn;
  void* ptr = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) != cudaSuccess ||
      qres != cudaDriverEntryPointSuccess)
    return nullptr;
  fn = reinterpret_cast<PFN_encodeTiled>(ptr);
  return fn;
}

int make_tensor_map_16bit(CUtensorMap* map, const void* base, int rank, const uint64_t* dims,
                          const uint64_t* strides_bytes, const uint32_t* box, CUtensorMapSwizzle swizzle) {
  PFN_encodeTiled enc = get_encode_tiled();
  if (!enc) return -4;
  cuuint64_t gdims[5]; cuuint64_t gstr[4]; cuuint32_t gbox[5]; cuuint32_t estr[5];
  for (int i = 0; i < rank; ++i) { gdims[i] = dims[i]; gbox[i] = box[i]; estr[i] = 1; }
  for (int i = 0; i + 1 < rank; ++i) gstr[i] = strides_bytes[i];
  // fp16 and bf16 move identically through TMA; the element type only matters for OOB fill (zeros either way)
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), gdims, gstr, gbox,
                   estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : 1000 + (int)r;
}


static int tc_build(const GemmTcArgs& a, int BN, CUtensorMap* map_a, CUtensorMap* map_b, GemmTcParams* pp, bool LN) {
  {
    const bool batched = a.batch > 1 && a.a_bs != 0;
    const uint64_t dims[3] = {(uint64_t)a.K, (uint64_t)a.M, (uint64_t)(batched ? a.batch : 1)};
    const uint64_t str[2] = {(uint64_t)a.lda * 2, (uint64_t)(batched ? a.a_bs : (long long)a.M * a.lda) * 2};
    const uint32_t box[3] = {TBK, TBM, 1};
    int rc = make_tensor_map_16bit(map_a, a.A, 3, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
  }
  {
    const bool batched = a.batch > 1 && a.b_bs != 0;
    const uint64_t dims[3] = {(uint64_t)a.K, (uint64_t)a.N, (uint64_t)(batched ? a.batch : 1)};
    const uint64_t str[2] = {(uint64_t)a.ldb * 2, (uint64_t)(batched ? a.b_bs : (long long)a.N * a.ldb) * 2};
    const uint32_t box[3] = {TBK, (uint32_t)BN, 1};
    int rc = make_tensor_map_16bit(map_b, a.Bm, 3, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
  }
  GemmTcParams& p = *pp;
  p.M = a.M; p.N = a.N; p.K = a.K; p.batch = a.batch;
  p.m_tiles = (a.M + TBM - 1) / TBM; p.n_tiles = (a.N + BN - 1) / BN;
  p.a_zmul = (a.batch > 1 && a.a_bs != 0) ? 1 : 0; p.b_zmul = (a.batch > 1 && a.b_bs != 0) ? 1 : 0;
  p.fmt = a.op_type == DT_BF16 ? 1 : 0; p.out_type = a.out_type;
  p.C = a.C; p.c_bs = a.c_bs; p.ldc = a.ldc;
  p.bias = a.bias; p.bias_bs = a.bias_bs; p.bias_mode = a.bias ? a.bias_mode : 0;
  p.act = a.act; p.residual = a.residual; p.colsum = a.colsum;
  p.out16 = a.out16;
  if (a.out16 != nullptr && (a.out_type != DT_F32 || a.colsum || a.scatter || a.N % 8 || a.ldc % 8 || (a.batch > 1 && a.c_bs % 8) ||
                             (reinterpret_cast<uintptr_t>(a.out16) & 15)))
    return -2;
  p.ln_mode = a.ln_mode; p.ln_w = a.ln_w; p.ln_b = a.ln_b; p.ln_out = a.ln_out; p.ln_type = a.ln_type;
  if ((a.ln_mode != 0) != LN) return -2;
  if (a.ln_mode != 0 && (BN > 128 || a.N != BN || a.colsum || a.scatter || !a.ln_out)) return -2;
  p.scatter = a.scatter; p.sc_C = a.scatter_C; p.sc_G = a.scatter_G; p.sc_H = a.scatter_H; p.sc_W = a.scatter_W;
  p.sc_cg = a.scatter_G ? a.scatter_C / a.scatter_G : 0;
  for (int i = 0; i < 4; ++i) { p.sc_ws[i] = a.scatter_ws[i]; p.sc_shift[i] = a.scatter_shift[i]; }
  p.sc_dst[0] = a.scatter_dst[0]; p.sc_dst[1] = a.scatter_dst[1];
  p.sc_pow2 = 0; p.sc_lW = p.sc_lL = 0;
  for (int i = 0; i < 4; ++i) p.sc_lws[i] = 0;
  if (a.scatter) {
    auto lg = [](int v) { int l = 0; while ((1 << l) < v) ++l; return (v > 0 && (1 << l) == v) ? l : -1; };
    const int lW = lg(a.scatter_W), lL = lg(a.scatter_H * a.scatter_W);
    bool ok = lW >= 0 && lL >= 0;
    for (int i = 0; i < a.scatter_G && i < 4; ++i) { const int l = lg(a.scatter_ws[i]); ok = ok && l >= 0 && l <= lW; if (l >= 0) p.sc_lws[i] = l; }
    if (ok) { p.sc_pow2 = 1; p.sc_lW = lW; p.sc_lL = lL; }
  }
  return 0;
}

// `b` (optional): a second problem sharing the launch (same operand / output types and N tile; see the kernel).
template <int BN, typename OutT, bool LN, int EW = 4>
static int launch_tc_bn(const GemmTcArgs& a, cudaStream_t st, const GemmTcArgs* b = nullptr) {
  if constexpr (!LN && EW == 4 && BN % 32 == 0) {
    static const int env_ew = getenv("DPMN_TC_EPI_WARPS") ? atoi(getenv("DPMN_TC_EPI_WARPS")) : 8;
    if (env_ew == 8 && a.colsum == nullptr && (!b || b->colsum == nullptr)) return launch_tc_bn<BN, OutT, LN, 8>(a, st, b);
  }
  CUtensorMap map_a, map_b, map_a2, map_b2;
  GemmTcParams p, p2;
  memset(&p2, 0, sizeof(p2));
  int rc = tc_build(a, BN, &map_a, &map_b, &p, LN);
  if (rc) return rc;
  int tiles1 = 0;
  if (b) {
    rc = tc_build(*b, BN, &map_a2, &map_b2, &p2, LN);
    if (rc) return rc;
    if (p2.fmt != p.fmt) return -2;
    tiles1 = p2.batch * p2.m_tiles * p2.n_tiles;
  } else {
    map_a2 = map_a; map_b2 = map_b;
  }
  int g_num_sms = 0;
  DPMN_CUDA_TRY(current_device_sms(&g_num_sms));
  const int tiles0 = p.batch * p.m_tiles * p.n_tiles;
  const int total = tiles0 + tiles1;
  static const int env_per_sm = getenv("DPMN_TC_PER_SM") ? atoi(getenv("DPMN_TC_PER_SM")) : 2;
  const int per_sm = LN ? 1 : (env_per_sm >= 2 ? 2 : 1);
  const int grid = total < per_sm * g_num_sms ? total : per_sm * g_num_sms;
  auto kern = gemm_tc_kernel<BN, OutT, LN, EW>;
  constexpr int smem = TcSmem<BN>::TOTAL;
  static PerDeviceOnce attr;      // per template instantiation, per device
  DPMN_CUDA_TRY(attr.smem_attr(kern, smem));
  DPMN_CUDA_TRY(launch_pdl(kern, dim3(grid), dim3(64 + 32 * EW), smem, st, map_a, map_b, map_a2, map_b2, p, p2, tiles0, tiles1));
  DPMN_LAUNCH_CHECK();
  return 0;
}

template <typename OutT>
static int launch_tc_out(const GemmTcArgs& a, cudaStream_t st) {
  // N tile: the largest of {256, 192, 128, 96, 64, 32} that divides N (fewest wasted columns), else 128 with a tail
  const int N = a.N;
  if (a.ln_mode != 0) {
    if constexpr (sizeof(OutT) == 4) {     // the fused second output exists for the fp32 residual-stream GEMMs
      if (N == 128) return launch_tc_bn<128, OutT, true>(a, st);
      if (N == 96) return launch_tc_bn<96, OutT, true>(a, st);
      if (N == 64) return launch_tc_bn<64, OutT, true>(a, st);
      if (N == 32) return launch_tc_bn<32, OutT, true>(a, st);
    }
    return -2;
  }
  static const int bn_max = getenv("DPMN_TC_BN_MAX") ? atoi(getenv("DPMN_TC_BN_MAX")) : 256;   // A/B switch for the N-tile choice
  if (N % 256 == 0 && bn_max >= 256) return launch_tc_bn<256, OutT, false>(a, st);
  if (N % 192 == 0 && bn_max >= 192) return launch_tc_bn<192, OutT, false>(a, st);
  if (N % 128 == 0) return launch_tc_bn<128, OutT, false>(a, st);
  if (N % 96 == 0) return launch_tc_bn<96, OutT, false>(a, st);
  if (N % 64 == 0) return launch_tc_bn<64, OutT, false>(a, st);
  if (N % 32 == 0) return launch_tc_bn<32, OutT, false>(a, st);
  return launch_tc_bn<128, OutT, false>(a, st);
}

// Two scatter-epilogue projections (q and kv of one block) in one launch; both N must be multiples of 96 or of 32.
int launch_gemm_tc_dual(const GemmTcArgs& a, const GemmTcArgs& b, cudaStream_t st) {
  for (const GemmTcArgs* g : {&a, &b}) {
    if (g->op_type != DT_F16 && g->op_type != DT_BF16) return -1;
    if (g->ln_mode != 0 || g->colsum != nullptr || g->batch != 1 || !g->scatter) return -2;
    if (g->K % 16 || g->lda % 8 || g->ldb % 8) return -2;
    if ((reinterpret_cast<uintptr_t>(g->A) | reinterpret_cast<uintptr_t>(g->Bm)) & 15) return -2;
    if (g->out_type == DT_F32 || g->scatter_G < 1 || (g->scatter_C / g->scatter_G) % 32 || g->N % g->scatter_C ||
        g->M % (g->scatter_H * g->scatter_W))
      return -2;
  }
  if (a.op_type != b.op_type || a.out_type != b.out_type) return -2;
  const bool n96 = a.N % 96 == 0 && b.N % 96 == 0;
  if (a.out_type == DT_F16) return n96 ? launch_tc_bn<96, __half, false>(a, st, &b) : launch_tc_bn<32, __half, false>(a, st, &b);
  return n96 ? launch_tc_bn<96, __nv_bfloat16, false>(a, st, &b) : launch_tc_bn<32, __nv_bfloat16, false>(a, st, &b);
}

bool gemm_tc_can_fuse_row_output(int N) { return N == 32 || N == 64 || N == 96 || N == 128; }

int launch_gemm_tc(const GemmTcArgs& a, cudaStream_t st) {
  if (a.op_type != DT_F16 && a.op_type != DT_BF16) return -1;
  if (a.out16 != nullptr && a.ln_mode != 0) return -2;                                  // one second output at a time
  if (a.ln_mode != 0 && gemm_res_ln_supported(a)) return launch_gemm_res_ln(a, st);   // all-TMA residual-stream kernel
  if (a.K % 16 || a.lda % 8 || a.ldb % 8) return -2;                 // 16-byte TMA strides, whole UMMA k-steps
  if ((reinterpret_cast<uintptr_t>(a.A) | reinterpret_cast<uintptr_t>(a.Bm)) & 15) return -2;
  if (a.batch > 1 && ((a.a_bs % 8) || (a.b_bs % 8))) return -2;
  if (a.scatter) {
    if (a.out_type == DT_F32 || a.batch != 1 || a.scatter_G < 1 || (a.scatter_C / a.scatter_G) % 32 || a.N % a.scatter_C ||
        a.M % (a.scatter_H * a.scatter_W))
      return -2;
  } else if (a.colsum == nullptr) {
    if (a.N % 8 || a.ldc % 8 || (reinterpret_cast<uintptr_t>(a.C) & 15)) return -2;   // 16-byte stores
  }
  switch (a.out_type) {
    case DT_F32: return launch_tc_out<float>(a, st);
    case DT_F16: return launch_tc_out<__half>(a, st);
    case DT_BF16: return launch_tc_out<__nv_bfloat16>(a, st);
  }
  return -1;
}

}  // namespace dpmn
