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
//   warp 0      TMA: fc1 weight (384 x 96, two 64-column k-blocks, SW128) once per CTA; the A tile (128 x 96) per tile
//   warp 1      tcgen05.mma: 3 n-blocks of M128 x N128, K = 96 (4 + 2 k-steps) into TMEM columns 0..383
//   warps 2-13  (12 worker warps, 3 per TMEM lane quarter)
//     phase 1   tcgen05.ld 32-column chunks (a chunk of a token row IS one 32-pixel row of one plane), + bias, GELU with the
//               polynomial part in packed fma.rn.f32x2, 16-bit, 4 x 8-byte shared stores into the plane buffer
//               [channel pair][y][plane of the pair][32 pixels] (pair stride 1026 words: conflict-free 8-byte loads);
//     phase 2   depthwise 3x3 + GELU: thread = (channel pair, two output rows) -- 24 x 16 = 384 items = one per worker
//               thread -- slides over x in 4-pixel chunks (8-byte loads of the 4 input rows of both planes, converted once,
//               reused by both output rows and all three kx taps), accumulates the pair in packed fma.rn.f32x2, and writes
//               one 4-byte (c, c+1) word per pixel into dt (B, 1024, 384): the 24 lanes of a row pair cover 96 contiguous bytes.
//   The MMAs of tile i + 1 overlap phase 2 of tile i (TMEM is drained once phase 1 has arrived on tmem_empty).
// Both phases are bound by the two MUFU ops of each GELU (ex2 + rcp; 4 per hidden element = 6.4 us per 128-token tile at
// 16 MUFU lanes per clock); everything else is sized to hide under that.  Arithmetic per element is that of gemm_tc's fc1
// epilogue followed by dwconv16_kernel (same GELU form, same tap order).
#include "common.cuh"
#include "kernels.h"
#include "tc_common.cuh"

namespace dpmn {

using namespace tc;

namespace {

constexpr int MA_ROWS = 128, MA_C = 96, MA_HID = 384, MA_SIDE = 32, MA_L = MA_SIDE * MA_SIDE;
constexpr int MA_PLANES = MA_ROWS * MA_HID / MA_L;         // 48 planes per tile
constexpr int MA_PAIRS = MA_PLANES / 2;                    // 24 channel pairs
constexpr int MA_PAIR_WORDS = 2 * MA_L / 2 + 2;            // 1026 words per pair: [y][plane of the pair][16 words] + 2 pad
constexpr int MA_KTILE = 128 * 128;                        // bytes of one 128-row x 64-column SW128 tile
constexpr int MA_A_BYTES = 2 * MA_KTILE;                   // K = 96 as two k-blocks (columns 96..127 are TMA zero fill)
constexpr int MA_W_BYTES = 2 * 3 * MA_KTILE;               // 384 rows x two k-blocks
constexpr int MA_P_BYTES = MA_PAIRS * MA_PAIR_WORDS * 4;   // 98 496
constexpr int MA_WORKERS = 12;                             // worker warps
constexpr int MA_WTHREADS = 32 * MA_WORKERS;               // 384 = 24 pairs x 16 row pairs
constexpr int MA_THREADS = 64 + MA_WTHREADS;
constexpr int MA_SMEM = MA_W_BYTES + MA_A_BYTES + MA_P_BYTES + 1024 /*alignment*/ + 256 /*barriers*/;   // 230 848 B
static_assert(MA_SMEM <= 232448, "227 KB of dynamic shared memory per CTA");
static_assert(MA_PLANES * MA_L == MA_ROWS * MA_HID, "a tile is a whole number of planes");
static_assert(MA_PAIRS * (MA_SIDE / 2) == MA_WTHREADS, "one (channel pair, row pair) item per worker thread");

struct MlpAParams {
  int tiles;                  // B * 8
  int fmt;                    // 0 fp16, 1 bf16
  const float* fc1_b;         // (384)
  const float* dw_w;          // (384, 1, 3, 3)
  const float* dw_b;          // (384)
  void* dt;                   // (B, 1024, 384) 16-bit, pixel-major
};

typedef unsigned long long u64;

__device__ __forceinline__ void worker_bar_sync() { asm volatile("bar.sync 1, 384;" ::: "memory"); }

__device__ __forceinline__ u64 p2(float x, float y) {
  u64 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(x), "f"(y));
  return r;
}
__device__ __forceinline__ void u2(u64 v, float& x, float& y) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(x), "=f"(y) : "l"(v));
}
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) {
  u64 d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ u64 mul2(u64 a, u64 b) {
  u64 d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ u64 add2(u64 a, u64 b) {
  u64 d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
// gelu_fast (common.cuh) on two values at once: the polynomial in packed f32x2, ex2 / rcp per lane.  Every operation is
// the .rn form of its scalar counterpart, so the results are bit-identical to gelu_fast.
__device__ __forceinline__ u64 gelu_fast2(u64 x) {
  float u0, u1;
  u2(mul2(x, x), u0, u1);
  const u64 u = p2(fminf(u0, 64.0f), fminf(u1, 64.0f));
  u64 q = fma2(p2(1.0142650e-3f, 1.0142650e-3f), u, p2(-1.0677574e-1f, -1.0677574e-1f));
  q = fma2(q, u, p2(-2.3011213f, -2.3011213f));
  float z0, z1, e0, e1, r0, r1;
  u2(mul2(x, q), z0, z1);
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(z0));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(z1));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r0) : "f"(1.0f + e0));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r1) : "f"(1.0f + e1));
  return mul2(x, p2(r0, r1));
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
template <typename T> __device__ __forceinline__ uint32_t u64_word(u64 v) {
  float a, b;
  u2(v, a, b);
  return f2_word<T>(a, b);
}

template <typename T>
__global__ void __launch_bounds__(MA_THREADS, 1)
mlp_fc1_dw_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w, const MlpAParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* s_w = smem;                                     // [kb][384 rows x 128 B]
  uint8_t* s_a = s_w + MA_W_BYTES;                         // [kb][128 rows x 128 B]
  uint32_t* s_pl = reinterpret_cast<uint32_t*>(s_a + MA_A_BYTES);   // [24 pairs][1026]: [y][plane of the pair][16 words]
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
    mbar_init(tmem_full, 1); mbar_init(tmem_empty, MA_WORKERS);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();        // everything above overlapped the predecessor's tail; its outputs are visible from here on
  pdl_trigger();

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
    // ================= 12 worker warps: fc1 epilogue, then the depthwise conv =================
    const int quarter = warp & 3;                             // TMEM lane quarter this warp may access
    const int third = (warp - 2) >> 2;                        // which of the quarter's three warps
    const int m_local = quarter * 32 + lane;
    const int wt = threadIdx.x - 64;                          // 0..383
    const int pair = wt % MA_PAIRS;                           // channel pair of the tile
    const int r0 = 2 * (wt / MA_PAIRS);                       // first of this thread's two output rows
    uint32_t* out_words = reinterpret_cast<uint32_t*>(p.dt);
    int it = 0;
    for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x, ++it) {
      // ---------------- phase 1: hidden = GELU(fc1) of this tile -> planes ----------------
      mbar_wait(tmem_full, (uint32_t)(it & 1));
      tc_fence_after();
#pragma unroll 1
      for (int c0 = third * 32; c0 < MA_HID; c0 += 96) {
        float4 bb[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) bb[j] = __ldg(reinterpret_cast<const float4*>(p.fc1_b + c0) + j);
        uint32_t r[32];
        tmem_ld_32x32(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)c0, r);
        tmem_ld_wait();
        // hidden value (token m, column j) sits at flat 384 m + j of the tile: plane flat / 1024, pixel flat % 1024; the 32
        // columns of a chunk are exactly one 32-pixel row of that plane (tests/test_mlp_tile_maps.py)
        const int flat = MA_HID * m_local + c0;
        const int pl = flat >> 10, y = (flat >> 5) & 31;
        uint2* dst = reinterpret_cast<uint2*>(s_pl + (pl >> 1) * MA_PAIR_WORDS + y * 32 + (pl & 1) * 16);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const u64 g0 = gelu_fast2(add2(p2(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1])), p2(bb[j].x, bb[j].y)));
          const u64 g1 = gelu_fast2(add2(p2(__uint_as_float(r[4 * j + 2]), __uint_as_float(r[4 * j + 3])), p2(bb[j].z, bb[j].w)));
          dst[j] = make_uint2(u64_word<T>(g0), u64_word<T>(g1));
        }
      }
      // the accumulator is free as soon as its last chunk has been read: the MMAs of the next tile overlap phase 2
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tmem_empty);
      worker_bar_sync();                                      // all 48 planes of this tile are complete

      // ---------------- phase 2: depthwise 3x3 + GELU, item = (pair, rows r0 and r0 + 1) ----------------
      {
        const int b = tile >> 3, tt = tile & 7;
        const int c = MA_PLANES * tt + 2 * pair;              // global channel of the pair (c, c + 1)
        u64 wk[9];
#pragma unroll
        for (int k = 0; k < 9; ++k) wk[k] = p2(__ldg(p.dw_w + c * 9 + k), __ldg(p.dw_w + (c + 1) * 9 + k));
        const u64 bias = p2(__ldg(p.dw_b + c), __ldg(p.dw_b + c + 1));
        const uint32_t* base = s_pl + pair * MA_PAIR_WORDS;
        // the four input rows r0 - 1 .. r0 + 2 (rows outside the plane contribute zeros)
        const bool ok[4] = {r0 > 0, true, true, r0 + 2 < MA_SIDE};
        const uint32_t* rowp[4] = {base + (r0 > 0 ? r0 - 1 : r0) * 32, base + r0 * 32, base + (r0 + 1) * 32,
                                   base + (r0 + 2 < MA_SIDE ? r0 + 2 : r0 + 1) * 32};
        u64 cur[4][4], left[4];
        uint2 na[4], nb[4];                                   // raw words of the next 4-pixel chunk, planes a / b
        auto load_raw = [&](int j) {
#pragma unroll
          for (int rr = 0; rr < 4; ++rr) {
            na[rr] = *reinterpret_cast<const uint2*>(rowp[rr] + 2 * j);
            nb[rr] = *reinterpret_cast<const uint2*>(rowp[rr] + 16 + 2 * j);
          }
        };
        auto convert = [&]() {
#pragma unroll
          for (int rr = 0; rr < 4; ++rr) {
            if (ok[rr]) {
              const float2 a0 = word_f2<T>(na[rr].x), a1 = word_f2<T>(na[rr].y);
              const float2 b0 = word_f2<T>(nb[rr].x), b1 = word_f2<T>(nb[rr].y);
              cur[rr][0] = p2(a0.x, b0.x); cur[rr][1] = p2(a0.y, b0.y); cur[rr][2] = p2(a1.x, b1.x); cur[rr][3] = p2(a1.y, b1.y);
            } else {
              cur[rr][0] = cur[rr][1] = cur[rr][2] = cur[rr][3] = 0ULL;
            }
          }
        };
        load_raw(0);
        convert();
#pragma unroll
        for (int rr = 0; rr < 4; ++rr) left[rr] = 0ULL;
        // word index of (pixel (r0, 0), channel pair c) in dt viewed as 32-bit words: ((b * 1024 + y * 32 + x) * 384 + c) / 2
        uint32_t* out0 = out_words + ((long long)b * MA_L + r0 * MA_SIDE) * (MA_HID / 2) + (c >> 1);
#pragma unroll 1
        for (int j = 0; j < MA_SIDE / 4; ++j) {
          u64 right[4];
          if (j + 1 < MA_SIDE / 4) {
            load_raw(j + 1);
#pragma unroll
            for (int rr = 0; rr < 4; ++rr)
              right[rr] = ok[rr] ? p2(word_f2<T>(na[rr].x).x, word_f2<T>(nb[rr].x).x) : 0ULL;
          } else {
#pragma unroll
            for (int rr = 0; rr < 4; ++rr) right[rr] = 0ULL;
          }
#pragma unroll
          for (int i = 0; i < 4; ++i) {
#pragma unroll
            for (int o = 0; o < 2; ++o) {
              u64 acc = bias;
#pragma unroll
              for (int ky = 0; ky < 3; ++ky) {               // tap order of dwconv16_kernel: ky outer, kx inner
                const int rr = o + ky;
                const u64 vl = i == 0 ? left[rr] : cur[rr][i == 0 ? 0 : i - 1];
                const u64 vr = i == 3 ? right[rr] : cur[rr][i == 3 ? 3 : i + 1];
                acc = fma2(vl, wk[3 * ky], acc);
                acc = fma2(cur[rr][i], wk[3 * ky + 1], acc);
                acc = fma2(vr, wk[3 * ky + 2], acc);
              }
              out0[((long long)o * MA_SIDE + 4 * j + i) * (MA_HID / 2)] = u64_word<T>(gelu_fast2(acc));
            }
          }
#pragma unroll
          for (int rr = 0; rr < 4; ++rr) left[rr] = cur[rr][3];
          if (j + 1 < MA_SIDE / 4) convert();
        }
      }
      worker_bar_sync();                                      // the plane buffer may be overwritten by the next tile
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
  DPMN_CUDA_TRY(launch_pdl(kern, dim3(p.tiles < sms ? p.tiles : sms), dim3(MA_THREADS), MA_SMEM, st, map_x, map_w, p));
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
