// Shared device helpers for the dpmn_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <stdint.h>

#define DPMN_CUDA_TRY(expr)                                   \
  do {                                                        \
    cudaError_t _e = (expr);                                  \
    if (_e != cudaSuccess) return (int)_e;                    \
  } while (0)

#define DPMN_LAUNCH_CHECK()                                   \
  do {                                                        \
    cudaError_t _e = cudaGetLastError();                      \
    if (_e != cudaSuccess) return (int)_e;                    \
  } while (0)

#include <atomic>
#include <cstdlib>

namespace dpmn {

// ---- per-device launch state (host) --------------------------------------------------------------------------------
// cudaFuncAttributeMaxDynamicSharedMemorySize and the SM count belong to a DEVICE, not to the process: a process that
// drives a second GPU must set the attribute there too and size its persistent grids with that GPU's SM count.  One
// bit per device ordinal (mod 64); the attribute is set BEFORE the bit is published, so a concurrent caller (ctypes
// releases the GIL) either sees the bit after the attribute exists or sets the attribute itself (idempotent).
struct PerDeviceOnce {
  std::atomic<unsigned long long> done{0};
  template <typename Kern>
  cudaError_t smem_attr(Kern kern, int smem_bytes) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    const unsigned long long bit = 1ULL << (dev & 63);
    if (done.load(std::memory_order_acquire) & bit) return cudaSuccess;
    e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
    if (e == cudaSuccess) done.fetch_or(bit, std::memory_order_release);
    return e;
  }
};
// SM count of the CURRENT device (cached per device ordinal).
inline cudaError_t current_device_sms(int* sms) {
  static std::atomic<int> cache[64];
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  int v = cache[dev & 63].load(std::memory_order_relaxed);
  if (v == 0) {
    e = cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
    if (e != cudaSuccess) return e;
    cache[dev & 63].store(v, std::memory_order_relaxed);
  }
  *sms = v;
  return cudaSuccess;
}

// Launch with the programmatic-stream-serialization attribute (see pdl_wait / pdl_trigger in tc_common.cuh).  Only kernels
// whose every predecessor-dependent access sits behind pdl_wait() may be launched this way.  Opt-in (DPMN_PDL=1): measured
// on B200 (profiles/r02_pdl_ab.md) it shortens back-to-back launches of one stream (stand-alone attention 7.2 -> 6.0 us at
// batch 8) but LENGTHENS the three-stream, two-slot graph of the pipelined forward (2.48 -> 2.59 ms per batch): the
// pre-launched CTAs hold shared memory that the other streams' kernels would have used.
inline bool pdl_enabled() {
  static const bool on = getenv("DPMN_PDL") && atoi(getenv("DPMN_PDL")) != 0;
  return on;
}
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
  cfg.attrs = at; cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

__device__ __forceinline__ float gelu_erf(float x) {
  // nn.GELU() default (exact erf form), pgrm.py:17
  return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f));
}

// GELU for the 16-bit modes: x * sigmoid(x * q(x^2)) with q fitted to logit(Phi(x)) / x (degree 2 in x^2, max
// |abs err| 2.5e-5 on the whole real line -- 20x below fp16 rounding at |y| ~ 1).  9 instructions (EX2 + RCP)
// instead of erff's ~30.  The coefficients carry the factor -log2(e) so that exp() is a bare ex2.
// The fp32 mode keeps the exact erff form above.
__device__ __forceinline__ float gelu_fast(float x) {
  const float u = fminf(x * x, 64.0f);
  float q = fmaf(1.0142650e-3f, u, -1.0677574e-1f);      // -log2e * (-0.00070303491, 0.07401130084)
  q = fmaf(q, u, -2.3011213f);                            // -log2e * 1.59501575816
  float e, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * q));     // bare MUFU.EX2 (exp2f adds a denormal-range fix-up)
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.0f + e));  // e = +inf -> r = 0 -> x * 0: the x -> -inf limit
  return x * r;
}

__device__ __forceinline__ float gelu_grad(float x) {
  // d/dx [ x * Phi(x) ] = Phi(x) + x * phi(x)
  const float cdf = 0.5f * (1.0f + erff(x * 0.70710678118654752440f));
  const float pdf = 0.39894228040143267794f * expf(-0.5f * x * x);
  return fmaf(x, pdf, cdf);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

template <typename T> __device__ __forceinline__ float to_f32(T v);
template <> __device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f32<__half>(__half v) { return __half2float(v); }
template <> __device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }

// two adjacent 16-bit values (4-byte aligned) -> float2
template <typename T> __device__ __forceinline__ float2 to_f32x2(const T* p);
template <> __device__ __forceinline__ float2 to_f32x2<__half>(const __half* p) {
  return __half22float2(*reinterpret_cast<const __half2*>(p));
}
template <> __device__ __forceinline__ float2 to_f32x2<__nv_bfloat16>(const __nv_bfloat16* p) {
  return __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(p));
}

template <typename T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ __half from_f32<__half>(float v) { return __float2half_rn(v); }
template <> __device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

// Counter-based Bernoulli masks of the train-mode Dropout / DropPath sites (pgrm.py:24,32,40,180,248,310,494,554-555).
// The reference draws them from torch's Philox stream, which cannot be reproduced outside torch ("parity
// unpinned"); here every mask element is a pure function of (seed, site, element index) -- splitmix64 finaliser -- so
// the backward regenerates exactly the masks of the forward and a test can rebuild them in numpy.
__host__ __device__ __forceinline__ uint32_t dpmn_hash32(unsigned long long seed, uint32_t site, unsigned long long idx) {
  unsigned long long z = seed + 0x9E3779B97F4A7C15ULL * (idx + 1ULL) + 0xD1B54A32D192ED03ULL * (unsigned long long)(site + 1u);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
  z ^= z >> 31;
  return (uint32_t)(z >> 32);
}
// multiplier of element idx: 1/(1-p) if kept, 0 if dropped (1 when p == 0)
__device__ __forceinline__ float drop_scale(float p, unsigned long long seed, uint32_t site, unsigned long long idx) {
  if (p <= 0.f) return 1.0f;
  const float u = (float)(dpmn_hash32(seed, site, idx) >> 8) * (1.0f / 16777216.0f);
  return u >= p ? 1.0f / (1.0f - p) : 0.0f;
}
enum DropSite : uint32_t { SITE_POS_Q = 1, SITE_POS_KV = 2, SITE_BLOCK = 16,   // + 16*blk:
                           SITE_ATTN = 0, SITE_MLP1 = 1, SITE_MLP2 = 2, SITE_PATH1 = 3, SITE_PATH2 = 4 };

// Window-major row p of group (ws, shift) -> original token index (pgrm.py:209-221 roll + partition).
// p = w*N + n, w = wr*(W/ws) + wc, n = i*ws + j; rolled coords (h', w') = (wr*ws+i, wc*ws+j) read
// original ((h'+shift) % H, (w'+shift) % W).
struct WinCoord {
  int hp, wp;   // rolled coordinates
  int token;    // original token index
  int win;      // window index within the image
  int n;        // index within the window
};

__device__ __forceinline__ WinCoord window_row_to_token(int p, int H, int W, int ws, int shift) {
  WinCoord c;
  const int N = ws * ws;
  const int nWw = W / ws;
  c.win = p / N;
  c.n = p - c.win * N;
  const int wr = c.win / nWw, wc = c.win - wr * nWw;
  const int i = c.n / ws, j = c.n - i * ws;
  c.hp = wr * ws + i;
  c.wp = wc * ws + j;
  int ho = c.hp + shift; if (ho >= H) ho -= H;
  int wo = c.wp + shift; if (wo >= W) wo -= W;
  c.token = ho * W + wo;
  return c;
}

// Inverse: original token -> window-major row of group (ws, shift).
__device__ __forceinline__ int token_to_window_row(int token, int H, int W, int ws, int shift) {
  int ho = token / W, wo = token - ho * W;
  int hp = ho - shift; if (hp < 0) hp += H;
  int wp = wo - shift; if (wp < 0) wp += W;
  const int nWw = W / ws;
  const int win = (hp / ws) * nWw + (wp / ws);
  const int n = (hp % ws) * ws + (wp % ws);
  return win * ws * ws + n;
}

// 3x3 slice label of the shift mask (pgrm.py:157-168) for rolled coordinate (hp, wp).
__device__ __forceinline__ int shift_region_label(int hp, int wp, int H, int W, int ws, int shift) {
  const int rh = (hp >= H - ws) + (hp >= H - shift);
  const int rw = (wp >= W - ws) + (wp >= W - shift);
  return 3 * rh + rw;
}

}  // namespace dpmn
