// fc1 + GELU + depthwise 3x3 + GELU of Mlp.forward (pgrm.py:30-36) in ONE kernel ("kernel A" of DESIGN.md section 9).
// Called by pgrm_forward_impl (api.cu) in the 16-bit modes for the production geometry (C = 96, hidden 384, 32 x 32 raw
// view); verified on hardware against the oracle and under compute-sanitizer memcheck
// (profiles/r02_sanitizer_memcheck_mlp_fused_a.log).  dpmnx_mlp_fc1_dwconv at the bottom is the stand-alone test hook.
//
// Why: the PGRM GEMM class is HBM-bound and launch-granular; fc1 writes and the depthwise conv
// re-reads the 37.7 MB hidden tensor per block.  The raw view of quirk 2 makes the fusion local: a 128-token M-tile of
// fc1's output (128 x 384 values) IS 48 complete 32 x 32 planes of the (384, 32, 32) view (tests/test_mlp_tile_maps.py),
// so the conv needs nothing from other tiles.
//
//   warp 0     TMA: fc1 weight (384 x 96, two 64-column k-blocks, SW128) once per CTA; the A tile (128 x 96) per tile
//   warp 1     tcgen05.mma: 3 n-blocks of M128 x N128, K = 96 (4 + 2 k-steps) into TMEM columns 0..383
//   warps 2-9  epilogue: tcgen05.ld 32-column chunks (two warps per lane quarter take alternate chunks), + bias,
//              gelu_fast, 16-bit, into the plane buffer at flat 384 m + j (plane stride 1026 halves);
//              named barrier; depthwise 3x3 + GELU out of the plane buffer: warp = plane row, lane = channel pair
//              (24 of 32 lanes), two x per step, packed fma.rn.f32x2 across the pair, tap order of dwconv16_kernel;
//              one 96-byte run per pixel into dt (B, 1024, 384); named barrier (plane buffer free again).
//   The MMAs of tile i + 1 overlap the conv of tile i (TMEM is drained once the epilogue has arrived on tmem_empty).
// Arithmetic is meant to be bit-identical to gemm_tc (fc1 epilogue) followed by dwconv16_kernel.
#include "common.cuh"
#include "kernels.h"
#include "tc_common.cuh"

namespace dpmn {

using namespace tc;

namespace {

constexpr int MA_ROWS = 128, MA_C = 96, MA_HID = 384, MA_SIDE = 32, MA_L = MA_SIDE * MA_SIDE;
constexpr int MA_PLANES = MA_ROWS * MA_HID / MA_L;         // 48 planes per tile
constexpr int MA_PWORDS = (MA_L + 2) / 2;                  // 513 words per plane in shared memory (odd: conflict-free lanes)
constexpr int MA_KTILE = 128 * 128;                        // bytes of one 128-row x 64-column SW128 tile
constexpr int MA_A_BYTES = 2 * MA_KTILE;                   // K = 96 as two k-blocks (columns 96..127 are TMA zero fill)
constexpr int MA_W_BYTES = 2 * 3 * MA_KTILE;               // 384 rows x two k-blocks
constexpr int MA_P_BYTES = MA_PLANES * MA_PWORDS * 4;      // 98 496
constexpr int MA_THREADS = 64 + 256;
constexpr int MA_SMEM = MA_W_BYTES + MA_A_BYTES + MA_P_BYTES + 1024 /*alignment*/ + 256 /*barriers*/;   // 230 848 B
static_assert(MA_SMEM <= 232448, "227 KB of dynamic shared memory per CTA");
static_assert(MA_PLANES * MA_L == MA_ROWS * MA_HID, "a tile is a whole number of planes");

struct MlpAParams {
  int tiles;                  // B * 8
  int fmt;                    // 0 fp16, 1 bf16
  const float* fc1_b;         // (384)
  const float* dw_w;          // (384, 1, 3, 3)
  const float* dw_b;          // (384)
  void* dt;                   // (B, 1024, 384) 16-bit, pixel-major
};

__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

__device__ __forceinline__ unsigned long long p2(float x, float y) {
  unsigned long long r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(x), "f"(y));
  return r;
}
__device__ __forceinline__ void u2(unsigned long long v, float& x, float& y) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(x), "=f"(y) : "l"(v));
}
__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c) {
  unsigned long long d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
template <typename T> __device__ __forceinline__ float2 word_f2(uint32_t w);
template <> __device__ __forceinline__ float2 word_f2<__half>(uint32_t w) {
  return __half22float2(*reinterpret_cast<const __half2*>(&w));
}
template <> __device__ __forceinline__ float2 word_f2<__nv_bfloat16>(uint32_t w) {
  return make_float2(__uint_as_float(w << 16), __uint_as_float(w & 0xffff0000u));
}
template <typename T> __device__ __forceinline__ uint32_t f2_word(float a, float b);
template <> __device__ __forceinline__ uint32_t f2_word<__half>(float a, float b) {
  const __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&h);
}
template <> __device__ __forceinline__ uint32_t f2_word<__nv_bfloat16>(float a, float b) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&h);
}

template <typename T>
__global__ void __launch_bounds__(MA_THREADS, 1)
mlp_fc1_dw_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w, const MlpAParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* s_w = smem;                                     // [kb][384 rows x 128 B]
  uint8_t* s_a = s_w + MA_W_BYTES;                         // [kb][128 rows x 128 B]
  uint32_t* s_pl = reinterpret_cast<uint32_t*>(s_a + MA_A_BYTES);   // [48][513] words = planes of 32 x 32 halves (+2 pad)
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(s_pl) + MA_P_BYTES);
  uint64_t* w_full = bars;
  uint64_t* a_full = bars + 1;
  uint64_t* a_empty = bars + 2;
  uint64_t* tmem_full = bars + 3;
  uint64_t* tmem_empty = bars + 4;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 5);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_x);
    tma_prefetch_desc(&map_w);
    mbar_init(w_full, 1); mbar_init(a_full, 1); mbar_init(a_empty, 1);
    mbar_init(tmem_full, 1); mbar_init(tmem_empty, 8);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ================= TMA producer =================
    if (lane == 0) {
      mbar_arrive_expect_tx(w_full, MA_W_BYTES);
      for (int kb = 0; kb < 2; ++kb)
        for (int nb = 0; nb < 3; ++nb)
          tma_load_3d(s_w + (kb * 3 + nb) * MA_KTILE, &map_w, w_full, kb * 64, nb * 128, 0);
      int it = 0;
      for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x, ++it) {
        mbar_wait(a_empty, (uint32_t)((it & 1) ^ 1));
        mbar_arrive_expect_tx(a_full, MA_A_BYTES);
        for (int kb = 0; kb < 2; ++kb) tma_load_3d(s_a + kb * MA_KTILE, &map_x, a_full, kb * 64, tile * MA_ROWS, 0);
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    if (lane == 0) {
      const uint32_t idesc = make_idesc_f16(p.fmt, MA_ROWS, 128);
      mbar_wait(w_full, 0);
      int it = 0;
      for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x, ++it) {
        mbar_wait(a_full, (uint32_t)(it & 1));
        mbar_wait(tmem_empty, (uint32_t)((it & 1) ^ 1));
        tc_fence_after();
#pragma unroll
        for (int nb = 0; nb < 3; ++nb) {
#pragma unroll
          for (int kb = 0; kb < 2; ++kb) {
            const uint64_t da = make_smem_desc_sw128(smem_u32(s_a + kb * MA_KTILE));
            const uint64_t db = make_smem_desc_sw128(smem_u32(s_w + (kb * 3 + nb) * MA_KTILE));
            const int ksteps = kb == 0 ? 4 : 2;               // K = 96 = 64 + 32
            for (int k = 0; k < ksteps; ++k)
              umma_f16(tmem_base + (uint32_t)(nb * 128), advance_desc_k(da, k), advance_desc_k(db, k), idesc, (kb | k) ? 1u : 0u);
          }
        }
        umma_commit(a_empty);                                 // the A tile may be overwritten once these MMAs have read it
        umma_commit(tmem_full);
      }
    }
  } else {
    // ================= epilogue + depthwise conv (warps 2..9) =================
    const int quarter = warp & 3;                             // TMEM lane quarter this warp may access
    const int ew = warp - 2;                                  // 0..7
    const int half = ew >> 2;                                 // which of the two warps of the quarter
    const int m_local = quarter * 32 + lane;
    T* dt = reinterpret_cast<T*>(p.dt);
    int it = 0;
    for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x, ++it) {
      mbar_wait(tmem_full, (uint32_t)(it & 1));
      tc_fence_after();
#pragma unroll 1
      for (int c0 = half * 32; c0 < MA_HID; c0 += 64) {
        float4 bb[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) bb[j] = __ldg(reinterpret_cast<const float4*>(p.fc1_b + c0) + j);
        uint32_t r[32];
        tmem_ld_32x32(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)c0, r);
        tmem_ld_wait();
        float v[32];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          v[4 * j] = __uint_as_float(r[4 * j]) + bb[j].x;
          v[4 * j + 1] = __uint_as_float(r[4 * j + 1]) + bb[j].y;
          v[4 * j + 2] = __uint_as_float(r[4 * j + 2]) + bb[j].z;
          v[4 * j + 3] = __uint_as_float(r[4 * j + 3]) + bb[j].w;
        }
        // hidden value (token m, column j) sits at flat 384 m + j of the tile = plane flat / 1024, offset flat % 1024;
        // a 32-column chunk never leaves its plane row (tests/test_mlp_tile_maps.py)
        const int flat = MA_HID * m_local + c0;
        uint32_t* dst = s_pl + (flat >> 10) * MA_PWORDS + ((flat & (MA_L - 1)) >> 1);
#pragma unroll
        for (int e = 0; e < 16; ++e) dst[e] = f2_word<T>(gelu_fast(v[2 * e]), gelu_fast(v[2 * e + 1]));
      }
      // the accumulator is free as soon as its last chunk has been read: the MMAs of the next tile overlap the conv below
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tmem_empty);
      epi_bar_sync();                                         // all 48 planes of this tile are complete

      if (lane < MA_PLANES / 2) {
        const int b = tile >> 3, tt = tile & 7;
        const int c = MA_PLANES * tt + 2 * lane;              // global channel of the pair (c, c + 1)
        unsigned long long wk[9];
#pragma unroll
        for (int k = 0; k < 9; ++k) wk[k] = p2(__ldg(p.dw_w + c * 9 + k), __ldg(p.dw_w + (c + 1) * 9 + k));
        const unsigned long long bias = p2(__ldg(p.dw_b + c), __ldg(p.dw_b + c + 1));
        const uint32_t* pl0 = s_pl + (2 * lane) * MA_PWORDS;
        const uint32_t* pl1 = pl0 + MA_PWORDS;
#pragma unroll 1
        for (int y = ew; y < MA_SIDE; y += 8) {
          // rows y - 1, y, y + 1 of both planes; a row outside the plane contributes zeros
          const bool ok0 = y > 0, ok2 = y + 1 < MA_SIDE;
          const uint32_t* ra[3] = {pl0 + (ok0 ? y - 1 : y) * 16, pl0 + y * 16, pl0 + (ok2 ? y + 1 : y) * 16};
          const uint32_t* rb[3] = {pl1 + (ok0 ? y - 1 : y) * 16, pl1 + y * 16, pl1 + (ok2 ? y + 1 : y) * 16};
          const bool okr[3] = {ok0, true, ok2};
          unsigned long long pm1[3], p0[3], p1[3];
#pragma unroll
          for (int ky = 0; ky < 3; ++ky) {
            const float2 fa = okr[ky] ? word_f2<T>(ra[ky][0]) : make_float2(0.f, 0.f);
            const float2 fb = okr[ky] ? word_f2<T>(rb[ky][0]) : make_float2(0.f, 0.f);
            pm1[ky] = p2(0.f, 0.f);
            p0[ky] = p2(fa.x, fb.x);
            p1[ky] = p2(fa.y, fb.y);
          }
          uint32_t* out = reinterpret_cast<uint32_t*>(dt + ((long long)b * MA_L + y * MA_SIDE) * MA_HID + c);
#pragma unroll 2
          for (int xw = 0; xw < MA_SIDE / 2; ++xw) {
            unsigned long long n0[3], n1[3];
#pragma unroll
            for (int ky = 0; ky < 3; ++ky) {
              float2 fa = make_float2(0.f, 0.f), fb = fa;
              if (okr[ky] && xw + 1 < MA_SIDE / 2) { fa = word_f2<T>(ra[ky][xw + 1]); fb = word_f2<T>(rb[ky][xw + 1]); }
              n0[ky] = p2(fa.x, fb.x);
              n1[ky] = p2(fa.y, fb.y);
            }
            unsigned long long acc_a = bias, acc_b = bias;
#pragma unroll
            for (int ky = 0; ky < 3; ++ky) {
              acc_a = fma2(pm1[ky], wk[3 * ky], acc_a); acc_a = fma2(p0[ky], wk[3 * ky + 1], acc_a); acc_a = fma2(p1[ky], wk[3 * ky + 2], acc_a);
              acc_b = fma2(p0[ky], wk[3 * ky], acc_b); acc_b = fma2(p1[ky], wk[3 * ky + 1], acc_b); acc_b = fma2(n0[ky], wk[3 * ky + 2], acc_b);
            }
            float a0, a1, b0, b1;
            u2(acc_a, a0, a1);
            u2(acc_b, b0, b1);
            out[(long long)(2 * xw) * (MA_HID / 2)] = f2_word<T>(gelu_fast(a0), gelu_fast(a1));
            out[(long long)(2 * xw + 1) * (MA_HID / 2)] = f2_word<T>(gelu_fast(b0), gelu_fast(b1));
#pragma unroll
            for (int ky = 0; ky < 3; ++ky) { pm1[ky] = p1[ky]; p0[ky] = n0[ky]; p1[ky] = n1[ky]; }
          }
        }
      }
      epi_bar_sync();                                         // the plane buffer may be overwritten by the next tile
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

template <typename T>
int launch_mlp_a(const void* x16, const void* w16, const float* fc1_b, const float* dw_w, const float* dw_b, void* dt, int B,
                 int fmt, cudaStream_t st) {
  CUtensorMap map_x, map_w;
  const long long rows = (long long)B * MA_L;
  {
    const uint64_t dims[3] = {(uint64_t)MA_C, (uint64_t)rows, 1};
    const uint64_t str[2] = {(uint64_t)MA_C * 2, (uint64_t)rows * MA_C * 2};
    const uint32_t box[3] = {64, MA_ROWS, 1};
    if (int rc = make_tensor_map_16bit(&map_x, x16, 3, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B)) return rc;
  }
  {
    const uint64_t dims[3] = {(uint64_t)MA_C, (uint64_t)MA_HID, 1};
    const uint64_t str[2] = {(uint64_t)MA_C * 2, (uint64_t)MA_HID * MA_C * 2};
    const uint32_t box[3] = {64, 128, 1};
    if (int rc = make_tensor_map_16bit(&map_w, w16, 3, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B)) return rc;
  }
  MlpAParams p;
  p.tiles = B * (MA_L / MA_ROWS); p.fmt = fmt; p.fc1_b = fc1_b; p.dw_w = dw_w; p.dw_b = dw_b; p.dt = dt;
  int sms = 0;
  DPMN_CUDA_TRY(current_device_sms(&sms));
  auto kern = mlp_fc1_dw_kernel<T>;
  static PerDeviceOnce attr;      // per template instantiation, per device
  DPMN_CUDA_TRY(attr.smem_attr(kern, MA_SMEM));
  kern<<<p.tiles < sms ? p.tiles : sms, MA_THREADS, MA_SMEM, st>>>(map_x, map_w, p);
  DPMN_LAUNCH_CHECK();
  return 0;
}

}  // namespace

bool mlp_fc1_dw_supported(int C, int hid, int L) { return C == MA_C && hid == MA_HID && L == MA_L; }

int launch_mlp_fc1_dw(const void* x16, const void* w16, const float* fc1_b, const float* dw_w, const float* dw_b, void* dt, int B,
                      DType t, cudaStream_t st) {
  if (t == DT_F16) return launch_mlp_a<__half>(x16, w16, fc1_b, dw_w, dw_b, dt, B, 0, st);
  if (t == DT_BF16) return launch_mlp_a<__nv_bfloat16>(x16, w16, fc1_b, dw_w, dw_b, dt, B, 1, st);
  return -1;
}

}  // namespace dpmn

// ---- test hook (not part of include/dpmn_b200.h) -------------------------------------------------------------------------
// x (B * 1024, 96) fp32 = LayerNorm_2 output, fc1_w (384, 96) fp32, biases / depthwise weights fp32 -> dt_out (B, 1024, 384)
// fp32 (the 16-bit result widened).  precision: 1 fp16, 2 bf16.  workspace >= 2 * (B*1024*96 + 384*96 + B*1024*384) bytes.
extern "C" int dpmnx_mlp_fc1_dwconv(const float* x, const float* fc1_w, const float* fc1_b, const float* dw_w, const float* dw_b,
                                    float* dt_out, int B, int precision, void* workspace, size_t workspace_bytes, void* stream) {
  using namespace dpmn;
  if (!x || !fc1_w || !fc1_b || !dw_w || !dw_b || !dt_out || !workspace || B < 1) return -1;
  if (precision != 1 && precision != 2) return -2;
  const long long nx = (long long)B * 1024 * 96, nw = 384LL * 96, nd = (long long)B * 1024 * 384;
  if (workspace_bytes < (size_t)(2 * (nx + nw + nd) + 3 * 256)) return -3;
  cudaStream_t st = (cudaStream_t)stream;
  const DType t = precision == 1 ? DT_F16 : DT_BF16;
  char* base = (char*)workspace;
  auto up = [](size_t v) { return (v + 255) & ~(size_t)255; };
  void* x16 = base;
  void* w16 = base + up(2 * nx);
  void* d16 = base + up(2 * nx) + up(2 * nw);
  if (int rc = launch_convert(x, x16, t, nx, st)) return rc;
  if (int rc = launch_convert(fc1_w, w16, t, nw, st)) return rc;
  int rc = precision == 1 ? launch_mlp_a<__half>(x16, w16, fc1_b, dw_w, dw_b, d16, B, 0, st)
                          : launch_mlp_a<__nv_bfloat16>(x16, w16, fc1_b, dw_w, dw_b, d16, B, 1, st);
  if (rc) return rc;
  return launch_widen(d16, t, dt_out, nd, st);
}
