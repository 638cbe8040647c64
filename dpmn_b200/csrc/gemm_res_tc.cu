// Residual-stream GEMM on tcgen05 with an all-TMA epilogue (sm_100a):
//
//     C[z][m, :] = R[z][m, :] + A[z][m, :] * B[z]^T + bias            fp32 C / R (in place allowed), N = BN <= 128
//     Y[z][m, :] = LayerNorm(C[z][m, :]; ln_w, ln_b)   or   = C[z][m, :]       written in the 16-bit operand type
//
// This is `x = x + f(x)` of SwinTransformerBlock.forward (pgrm.py:329-330) fused with the LayerNorm its consumer
// applies next (norm2, or norm1_kv of the next block, pgrm.py:322-323,330): the token stream is read and written
// exactly once per residual update.
//
// Why a kernel of its own: with one thread per accumulator row (the tcgen05.ld layout) direct global accesses
// touch 32 different 128-byte lines per warp instruction.  Here the residual tile arrives by TMA into 128B-swizzled
// 32-column chunk tiles, the epilogue works in shared memory only (conflict-free 16-byte accesses at swizzled
// positions), and both outputs leave by TMA bulk stores.
//   warp 0   TMA producer (A/B k-blocks + the residual tile of each output tile)
//   warp 1   MMA issuer (tcgen05.mma M128 x N=BN x K16, accumulators in TMEM, double buffered)
//   warps 2-5 epilogue; one elected thread issues the TMA stores
#include "common.cuh"
#include "kernels.h"
#include "tc_common.cuh"
#include <cstring>

namespace dpmn {

using namespace tc;

constexpr int RBM = 128, RBK = 64, RSTAGES = 2, RTHREADS = 64 + 256 + 32;   // TMA warp, MMA warp, 8 epilogue warps, store warp

struct GemmResParams {
  int M, N, K, batch, m_tiles;
  int a_zmul, b_zmul;
  int fmt;
  const float* bias; long long bias_bs;    // per n
  int ln_mode;                             // 1 LayerNorm, 2 copy
  const float *ln_w, *ln_b;
};

template <int BN>
struct ResSmem {
  static constexpr int A_BYTES = RBM * RBK * 2;
  static constexpr int B_BYTES = BN * RBK * 2;
  static constexpr int STAGE = A_BYTES + B_BYTES;
  static constexpr int CHUNKS = BN / 32;
  static constexpr int C_TILE = RBM * 128;             // 128 rows x 32 fp32, 128B swizzle
  static constexpr int Y_TILE = RBM * 64;              // 128 rows x 32 x 16-bit, 64B swizzle
  static constexpr int C_BUF = CHUNKS * C_TILE;        // one residual / result tile (in place)
  static constexpr int Y_BUF = CHUNKS * Y_TILE;
  static constexpr int STATS = RBM * 2 * 2 * 4;        // per row, per column half: (sum, sum of squares)
  static constexpr int LNP = 2 * BN * 4;               // LayerNorm weight and bias
  static constexpr int TOTAL = RSTAGES * STAGE + 2 * C_BUF + 2 * Y_BUF + STATS + LNP + 1024 + 256;
};

__device__ __forceinline__ void tma_store_3d(const CUtensorMap* map, const void* src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(map)),
               "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read1() { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }

// Pipeline (per CTA, persistent over output tiles; tile i uses residual / result buffer i & 1):
//   producer   residual tile of tile i -> buffer i & 1 as soon as the TMA store of tile i - 2 has read it out
//              (res_empty), then the A / B k-blocks through a 2-stage ring
//   MMA        accumulates tile i into TMEM stage i & 1
//   epilogue   8 warps = 2 per TMEM lane quarter; the two threads of a row take 48 of its 96 columns each (BN / 2 in
//              general), add accumulator + bias into the residual tile IN PLACE (swizzled shared memory), exchange their
//              partial LayerNorm sums through shared memory, write the normalised 16-bit row and arrive on out_ready
//   store warp one thread issues the TMA stores of both outputs of tile i, waits until they have read shared memory and
//              hands the buffer back to the producer (res_empty): the residual tile of tile i + 2 is in flight while
//              tile i + 1 is in its epilogue.
template <int BN, typename YT>
__global__ void __launch_bounds__(RTHREADS, 1)
gemm_res_ln_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                   const __grid_constant__ CUtensorMap map_r, const __grid_constant__ CUtensorMap map_c,
                   const __grid_constant__ CUtensorMap map_y, const __grid_constant__ GemmResParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  using S = ResSmem<BN>;
  static_assert(BN % 32 == 0 && (BN / 2) % 16 == 0, "column halves are whole 16-column tcgen05.ld chunks");
  uint8_t* tiles = smem;
  uint8_t* c_stage = smem + RSTAGES * S::STAGE;                 // [2][CHUNKS][128 x 128 B]
  uint8_t* y_stage = c_stage + 2 * S::C_BUF;                    // [2][CHUNKS][128 x 64 B]
  float* s_stat = reinterpret_cast<float*>(y_stage + 2 * S::Y_BUF);   // [128][2 halves][2]
  float* s_lnw = s_stat + RBM * 4;
  float* s_lnb = s_lnw + BN;
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_lnb + BN);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + RSTAGES;
  uint64_t* tmem_full = bars + 2 * RSTAGES;      // [2]
  uint64_t* tmem_empty = tmem_full + 2;          // [2]
  uint64_t* res_full = tmem_full + 4;            // [2]
  uint64_t* res_empty = tmem_full + 6;           // [2]
  uint64_t* out_ready = tmem_full + 8;           // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 10);
  constexpr uint32_t TMEM_COLS = (2 * BN <= 64) ? 64 : (2 * BN <= 128) ? 128 : 256;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_kb = (p.K + RBK - 1) / RBK;
  const int total_tiles = p.batch * p.m_tiles;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_a); tma_prefetch_desc(&map_b); tma_prefetch_desc(&map_r);
    tma_prefetch_desc(&map_c); tma_prefetch_desc(&map_y);
    for (int i = 0; i < RSTAGES; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1); mbar_init(&tmem_empty[i], 8);
      mbar_init(&res_full[i], 1); mbar_init(&res_empty[i], 1); mbar_init(&out_ready[i], 8);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, TMEM_COLS);
  if (p.ln_mode == 1)
    for (int i = threadIdx.x; i < BN; i += RTHREADS) { s_lnw[i] = __ldg(p.ln_w + i); s_lnb[i] = __ldg(p.ln_b + i); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();        // everything above overlapped the predecessor's tail; its outputs are visible from here on
  pdl_trigger();

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      int it = 0;
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++it) {
        const int m_blk = t % p.m_tiles, z = t / p.m_tiles;
        const int rb = it & 1;
        // residual tile first: it is what the epilogue waits for longest
        mbar_wait(&res_empty[rb], (uint32_t)(((it >> 1) & 1) ^ 1));
        mbar_arrive_expect_tx(&res_full[rb], S::C_BUF);
#pragma unroll
        for (int c = 0; c < S::CHUNKS; ++c)
          tma_load_3d(c_stage + rb * S::C_BUF + c * S::C_TILE, &map_r, &res_full[rb], c * 32, m_blk * RBM, z);
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = tiles + stage * S::STAGE;
          mbar_arrive_expect_tx(&full_bar[stage], S::STAGE);
          tma_load_3d(sa, &map_a, &full_bar[stage], kb * RBK, m_blk * RBM, z * p.a_zmul);
          tma_load_3d(sa + S::A_BYTES, &map_b, &full_bar[stage], kb * RBK, 0, z * p.b_zmul);
          if (++stage == RSTAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = make_idesc_f16(p.fmt, RBM, BN);
      int stage = 0; uint32_t phase = 0;
      int acc = 0; uint32_t acc_phase = 0;
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(tiles + stage * S::STAGE);
          const uint64_t da = make_smem_desc_sw128(sa);
          const uint64_t db = make_smem_desc_sw128(sa + S::A_BYTES);
          const int k_left = p.K - kb * RBK;
          const int ksteps = k_left >= RBK ? RBK / 16 : (k_left + 15) / 16;
          for (int k = 0; k < ksteps; ++k)
            umma_f16(d_tmem, advance_desc_k(da, k), advance_desc_k(db, k), idesc, (kb | k) ? 1u : 0u);
          umma_commit(&empty_bar[stage]);
          if (++stage == RSTAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit(&tmem_full[acc]);
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else if (warp == 10) {
    // ================= store warp =================
    if (lane == 0) {
      int it = 0;
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++it) {
        const int m_blk = t % p.m_tiles, z = t / p.m_tiles;
        const int rb = it & 1;
        mbar_wait(&out_ready[rb], (uint32_t)((it >> 1) & 1));
#pragma unroll
        for (int c = 0; c < S::CHUNKS; ++c) {
          tma_store_3d(&map_c, c_stage + rb * S::C_BUF + c * S::C_TILE, c * 32, m_blk * RBM, z);
          tma_store_3d(&map_y, y_stage + rb * S::Y_BUF + c * S::Y_TILE, c * 32, m_blk * RBM, z);
        }
        tma_store_commit();
        tma_store_wait_read0();                              // both buffers of this tile have been read out
        mbar_arrive(&res_empty[rb]);
      }
    }
  } else {
    constexpr int HALF = BN / 2;                         // columns per thread
    constexpr int NQ = HALF / 16;                        // 16-column tcgen05.ld chunks per thread
    const int quarter = warp & 3;
    const int half = (warp - 2) >> 2;                    // which column half of the row this thread owns
    const int row = quarter * 32 + lane;                 // tile row of this thread
    const int sw128 = (row & 7), sw64 = (row >> 1) & 3;
    const int col0 = half * HALF;
    int acc = 0; uint32_t acc_phase = 0;
    int it = 0;
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++it) {
      const int z = t / p.m_tiles;
      const int rb = it & 1;
      const float* bias = p.bias ? p.bias + (long long)z * p.bias_bs : nullptr;
      uint8_t* cbuf = c_stage + rb * S::C_BUF;
      uint8_t* ybuf = y_stage + rb * S::Y_BUF;
      // this thread's bias values: loaded before the waits, so their L2 latency hides behind the MMAs of the tile
      float4 bb_all[NQ][4];
#pragma unroll
      for (int q = 0; q < NQ; ++q)
#pragma unroll
        for (int j = 0; j < 4; ++j)
          bb_all[q][j] = bias ? __ldg(reinterpret_cast<const float4*>(bias + col0 + 16 * q) + j) : make_float4(0.f, 0.f, 0.f, 0.f);
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      mbar_wait(&res_full[rb], (uint32_t)((it >> 1) & 1));
      float rowbuf[HALF];
      float s1 = 0.f, s2 = 0.f;
#pragma unroll
      for (int q = 0; q < NQ; ++q) {
        const int n0 = col0 + 16 * q;                    // first column of this 16-column piece
        const float4 (&bb)[4] = bb_all[q];
        uint32_t r[32];
        tmem_ld_32x16(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * BN + n0), r);
        tmem_ld_wait();
        uint8_t* crow = cbuf + (n0 >> 5) * S::C_TILE + row * 128;
        const int u0 = (n0 & 31) >> 2;                   // first 16-byte unit of the piece inside the 128-byte row
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float4* ptr = reinterpret_cast<float4*>(crow + (((u0 + j) ^ sw128) << 4));
          float4 v = *ptr;                                   // residual
          v.x += __uint_as_float(r[4 * j]) + bb[j].x; v.y += __uint_as_float(r[4 * j + 1]) + bb[j].y;
          v.z += __uint_as_float(r[4 * j + 2]) + bb[j].z; v.w += __uint_as_float(r[4 * j + 3]) + bb[j].w;
          *ptr = v;                                          // result in place -> TMA store
          rowbuf[16 * q + 4 * j] = v.x; rowbuf[16 * q + 4 * j + 1] = v.y;
          rowbuf[16 * q + 4 * j + 2] = v.z; rowbuf[16 * q + 4 * j + 3] = v.w;
          s1 += (v.x + v.y) + (v.z + v.w);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      // second output: LayerNorm over the whole row (two-pass: mean first, then squared deviations; the two column halves
      // of a row exchange their partial sums through shared memory behind a 64-thread named barrier per lane quarter)
      float mu = 0.f, rstd = 1.f;
      if (p.ln_mode == 1) {
        s_stat[(row * 2 + half) * 2] = s1;
        asm volatile("bar.sync %0, 64;" ::"r"(2 + quarter) : "memory");
        mu = (s1 + s_stat[(row * 2 + (half ^ 1)) * 2]) * (1.0f / BN);
#pragma unroll
        for (int j = 0; j < HALF; ++j) { const float dlt = rowbuf[j] - mu; s2 = fmaf(dlt, dlt, s2); }
        s_stat[(row * 2 + half) * 2 + 1] = s2;
        asm volatile("bar.sync %0, 64;" ::"r"(2 + quarter) : "memory");
        rstd = rsqrtf((s2 + s_stat[(row * 2 + (half ^ 1)) * 2 + 1]) * (1.0f / BN) + 1e-5f);
      }
#pragma unroll
      for (int g8 = 0; g8 < HALF / 8; ++g8) {
        const int n0 = col0 + 8 * g8;
        union { uint4 u; YT h[8]; } pk;
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          float y = rowbuf[8 * g8 + e];
          if (p.ln_mode == 1) y = (y - mu) * rstd * s_lnw[n0 + e] + s_lnb[n0 + e];
          pk.h[e] = from_f32<YT>(y);
        }
        uint8_t* yrow = ybuf + (n0 >> 5) * S::Y_TILE + row * 64;
        *reinterpret_cast<uint4*>(yrow + ((((n0 & 31) >> 3) ^ sw64) << 4)) = pk.u;
      }
      fence_proxy_async();                                   // smem writes -> visible to the TMA (async proxy)
      __syncwarp();
      if (lane == 0) mbar_arrive(&out_ready[rb]);            // 8 warps -> the store warp issues this tile's TMA stores
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// ---- host -------------------------------------------------------------------------------------------------
int make_tensor_map_any(CUtensorMap* map, CUtensorMapDataType dt, int elem_bytes, const void* base, int rank,
                        const uint64_t* dims, const uint64_t* strides_bytes, const uint32_t* box, CUtensorMapSwizzle swizzle) {
  PFN_encodeTiled enc = get_encode_tiled();
  if (!enc) return -4;
  (void)elem_bytes;
  cuuint64_t gdims[5]; cuuint64_t gstr[4]; cuuint32_t gbox[5]; cuuint32_t estr[5];
  for (int i = 0; i < rank; ++i) { gdims[i] = dims[i]; gbox[i] = box[i]; estr[i] = 1; }
  for (int i = 0; i + 1 < rank; ++i) gstr[i] = strides_bytes[i];
  CUresult r = enc(map, dt, (cuuint32_t)rank, const_cast<void*>(base), gdims, gstr, gbox, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : 1000 + (int)r;
}

template <int BN, typename YT>
static int launch_res_bn(const GemmTcArgs& a, cudaStream_t st) {
  CUtensorMap map_a, map_b, map_r, map_c, map_y;
  {
    const bool batched = a.batch > 1 && a.a_bs != 0;
    const uint64_t dims[3] = {(uint64_t)a.K, (uint64_t)a.M, (uint64_t)(batched ? a.batch : 1)};
    const uint64_t str[2] = {(uint64_t)a.lda * 2, (uint64_t)(batched ? a.a_bs : (long long)a.M * a.lda) * 2};
    const uint32_t box[3] = {RBK, RBM, 1};
    int rc = make_tensor_map_16bit(&map_a, a.A, 3, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
  }
  {
    const bool batched = a.batch > 1 && a.b_bs != 0;
    const uint64_t dims[3] = {(uint64_t)a.K, (uint64_t)a.N, (uint64_t)(batched ? a.batch : 1)};
    const uint64_t str[2] = {(uint64_t)a.ldb * 2, (uint64_t)(batched ? a.b_bs : (long long)a.N * a.ldb) * 2};
    const uint32_t box[3] = {RBK, BN, 1};
    int rc = make_tensor_map_16bit(&map_b, a.Bm, 3, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
  }
  {
    const uint64_t dims[3] = {(uint64_t)a.N, (uint64_t)a.M, (uint64_t)a.batch};
    const uint64_t str[2] = {(uint64_t)a.ldc * 4, (uint64_t)(a.batch > 1 ? a.c_bs : (long long)a.M * a.ldc) * 4};
    const uint32_t box[3] = {32, RBM, 1};
    int rc = make_tensor_map_any(&map_r, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, a.residual, 3, dims, str, box,
                                 CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
    rc = make_tensor_map_any(&map_c, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, a.C, 3, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
  }
  {
    const uint64_t dims[3] = {(uint64_t)a.N, (uint64_t)a.M, (uint64_t)a.batch};
    const uint64_t str[2] = {(uint64_t)a.N * 2, (uint64_t)a.M * a.N * 2};
    const uint32_t box[3] = {32, RBM, 1};
    int rc = make_tensor_map_16bit(&map_y, a.ln_out, 3, dims, str, box, CU_TENSOR_MAP_SWIZZLE_64B);
    if (rc) return rc;
  }
  GemmResParams p;
  memset(&p, 0, sizeof(p));
  p.M = a.M; p.N = a.N; p.K = a.K; p.batch = a.batch; p.m_tiles = (a.M + RBM - 1) / RBM;
  p.a_zmul = (a.batch > 1 && a.a_bs != 0) ? 1 : 0; p.b_zmul = (a.batch > 1 && a.b_bs != 0) ? 1 : 0;
  p.fmt = a.op_type == DT_BF16 ? 1 : 0;
  p.bias = a.bias; p.bias_bs = a.bias_bs; p.ln_mode = a.ln_mode; p.ln_w = a.ln_w; p.ln_b = a.ln_b;
  int num_sms = 0;
  DPMN_CUDA_TRY(current_device_sms(&num_sms));
  const int total = p.batch * p.m_tiles;
  const int grid = total < num_sms ? total : num_sms;
  auto kern = gemm_res_ln_kernel<BN, YT>;
  constexpr int smem = ResSmem<BN>::TOTAL;
  static PerDeviceOnce attr;      // per template instantiation, per device
  DPMN_CUDA_TRY(attr.smem_attr(kern, smem));
  DPMN_CUDA_TRY(launch_pdl(kern, dim3(grid), dim3(RTHREADS), smem, st, map_a, map_b, map_r, map_c, map_y, p));
  DPMN_LAUNCH_CHECK();
  return 0;
}

bool gemm_res_ln_supported(const GemmTcArgs& a) {
  if (a.ln_mode == 0 || a.out_type != DT_F32 || a.residual == nullptr || a.ln_out == nullptr) return false;
  if (a.N != 32 && a.N != 64 && a.N != 96) return false;   // N = 128: two 64 KB fp32 tiles do not fit (generic kernel)
  if (a.ln_type != DT_F16 && a.ln_type != DT_BF16) return false;
  if (a.bias_mode > 1 || a.act != 0 || a.colsum || a.scatter) return false;
  if (a.K % 16 || a.lda % 8 || a.ldb % 8 || a.ldc % 4) return false;
  if ((reinterpret_cast<uintptr_t>(a.C) | reinterpret_cast<uintptr_t>(a.residual) | reinterpret_cast<uintptr_t>(a.ln_out)) & 15) return false;
  if (a.batch > 1 && ((a.c_bs % 4) || (a.a_bs % 8) || (a.b_bs % 8))) return false;
  return true;
}

template <typename YT>
static int launch_res_y(const GemmTcArgs& a, cudaStream_t st) {
  switch (a.N) {
    case 32: return launch_res_bn<32, YT>(a, st);
    case 64: return launch_res_bn<64, YT>(a, st);
    case 96: return launch_res_bn<96, YT>(a, st);
  }
  return -2;
}

int launch_gemm_res_ln(const GemmTcArgs& a, cudaStream_t st) {
  if (!gemm_res_ln_supported(a)) return -2;
  return a.ln_type == DT_F16 ? launch_res_y<__half>(a, st) : launch_res_y<__nv_bfloat16>(a, st);
}

}  // namespace dpmn
