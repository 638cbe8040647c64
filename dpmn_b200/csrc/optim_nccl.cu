// The tail of one data-parallel training step (interfaces/super_resolution.py:268-278, interfaces/base.py:160-162,208-221):
//
//   dpmn_allreduce_bucket   the ONE collective of the step: ncclAllReduce(sum) over (a segment of) the flat fp32 gradient
//                           bucket, on the caller's stream.  The reference reduces gradients through nn.DataParallel
//                           (base.py:160-162: replicate + gather onto GPU 0 every iteration); here every rank owns a replica
//                           and NCCL moves the 228.7 MB bucket over NVLink 5 / NVSwitch.
//   dpmn_clip_adam_step     torch.nn.utils.clip_grad_norm_(module.parameters(), 0.25) for every module
//                           (super_resolution.py:270-275) + Adam(lr, betas (0.5, 0.999)) (base.py:208-221) over the flat
//                           parameter / gradient / moment buffers in two launches: per-module sums of squares, then one
//                           fused pass that applies the 1/world scale, the module's clip coefficient and the Adam update.
//
// libnccl is NOT a link-time dependency: the entry points are resolved with dlopen/dlsym at the first call, preferring the
// copy that is already loaded in the process (torch's bundled libnccl.so.2), so the library loads on a box without NCCL and
// only the collective itself fails there (DPMN_E_DEVICE).
#include <dlfcn.h>

#include <cmath>
#include <cstdlib>
#include <cstring>
#include <mutex>

#include "../../include/dpmn_b200.h"
#include "common.cuh"

namespace {

// ---- NCCL, resolved at run time ------------------------------------------------------------------------------------
struct NcclUniqueId { char internal[128]; };     // ncclUniqueId (NCCL_UNIQUE_ID_BYTES = 128)
typedef void* NcclComm;
typedef int (*PFN_ncclGetUniqueId)(NcclUniqueId*);
typedef int (*PFN_ncclCommInitRank)(NcclComm*, int, NcclUniqueId, int);
typedef int (*PFN_ncclCommDestroy)(NcclComm);
typedef int (*PFN_ncclAllReduce)(const void*, void*, size_t, int /*ncclDataType_t*/, int /*ncclRedOp_t*/, NcclComm, cudaStream_t);
typedef int (*PFN_ncclGetVersion)(int*);
constexpr int kNcclSum = 0, kNcclFloat16 = 6, kNcclFloat32 = 7, kNcclBfloat16 = 9;

struct NcclApi {
  void* handle = nullptr;
  PFN_ncclGetUniqueId get_unique_id = nullptr;
  PFN_ncclCommInitRank comm_init_rank = nullptr;
  PFN_ncclCommDestroy comm_destroy = nullptr;
  PFN_ncclAllReduce all_reduce = nullptr;
  PFN_ncclGetVersion get_version = nullptr;
  bool ok = false;
};

NcclApi& nccl_api() {
  static NcclApi api;
  static std::once_flag once;
  std::call_once(once, [] {
    // 1. the copy already mapped into this process (torch imports its bundled libnccl.so.2): same version as the
    //    process group that carried the unique id, no second NCCL in the address space
    //    (a different libnccl.so.2 mapped BEFORE torch is imported would shadow torch's own by soname and break its import,
    //    so dpmn_b200._lib.load() imports torch first and nothing here is RTLD_GLOBAL)
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD | RTLD_LOCAL);
    // 2. an explicit path, 3. the system copy
    if (!h && getenv("DPMN_NCCL_LIB")) h = dlopen(getenv("DPMN_NCCL_LIB"), RTLD_NOW | RTLD_LOCAL);
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_LOCAL);
    if (!h) return;
    api.handle = h;
    api.get_unique_id = (PFN_ncclGetUniqueId)dlsym(h, "ncclGetUniqueId");
    api.comm_init_rank = (PFN_ncclCommInitRank)dlsym(h, "ncclCommInitRank");
    api.comm_destroy = (PFN_ncclCommDestroy)dlsym(h, "ncclCommDestroy");
    api.all_reduce = (PFN_ncclAllReduce)dlsym(h, "ncclAllReduce");
    api.get_version = (PFN_ncclGetVersion)dlsym(h, "ncclGetVersion");
    api.ok = api.get_unique_id && api.comm_init_rank && api.comm_destroy && api.all_reduce;
  });
  return api;
}

// ---- clip + Adam ---------------------------------------------------------------------------------------------------
constexpr int kOptThreads = 256;
constexpr int kOptChunk = 8192;          // elements per block (32 per thread as 8 float4)

constexpr int kMaxSegments = 32;
// module boundaries inside the flat buffers: off[0] = 0 <= off[1] <= ... <= off[n] = total; passed by value (kernel argument)
struct SegOffsets { long long off[kMaxSegments + 1]; int n; };

__device__ __forceinline__ int find_segment(const SegOffsets& so, long long i) {
  int s = 0;
  while (s + 1 < so.n && i >= so.off[s + 1]) ++s;
  return s;
}

// sq[s] += sum over the segment's elements of (g * scale)^2, accumulated in double (one atomicAdd per block per segment)
__global__ void __launch_bounds__(kOptThreads)
grad_sqnorm_kernel(const float* __restrict__ g, const SegOffsets so, long long total, float scale, double* __restrict__ sq) {
  __shared__ double red[kOptThreads / 32];
  const long long base = (long long)blockIdx.x * kOptChunk;
  const long long end = base + kOptChunk < total ? base + kOptChunk : total;
  int s = find_segment(so, base);
  long long lo = base;
  while (lo < end) {                                       // a chunk rarely spans more than one segment
    const long long seg_end = so.off[s + 1] < end ? so.off[s + 1] : end;
    float acc = 0.f;
    // vector body when the sub-range is 16-byte aligned, scalar otherwise (segment boundaries are arbitrary)
    const long long a0 = (lo + 3) & ~3LL;
    for (long long i = lo + threadIdx.x; i < (a0 < seg_end ? a0 : seg_end); i += kOptThreads) { const float v = g[i] * scale; acc = fmaf(v, v, acc); }
    const long long nvec = a0 < seg_end ? (seg_end - a0) >> 2 : 0;
    const float4* g4 = reinterpret_cast<const float4*>(g + a0);
    for (long long i = threadIdx.x; i < nvec; i += kOptThreads) {
      const float4 v = __ldg(g4 + i);
      acc = fmaf(v.x * scale, v.x * scale, acc); acc = fmaf(v.y * scale, v.y * scale, acc);
      acc = fmaf(v.z * scale, v.z * scale, acc); acc = fmaf(v.w * scale, v.w * scale, acc);
    }
    for (long long i = a0 + 4 * nvec + threadIdx.x; i < seg_end && a0 < seg_end; i += kOptThreads) { const float v = g[i] * scale; acc = fmaf(v, v, acc); }
    double d = (double)dpmn::warp_sum(acc);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = d;
    __syncthreads();
    if (threadIdx.x == 0) {
      double t = 0.0;
      for (int w = 0; w < kOptThreads / 32; ++w) t += red[w];
      atomicAdd(&sq[s], t);
    }
    __syncthreads();
    lo = seg_end;
    ++s;
  }
}

struct AdamArgs {
  float* p; const float* g; float* m; float* v;
  SegOffsets so; long long total;
  const double* sq;               // per-segment sum of squares of the SCALED gradient
  float grad_scale;               // 1 / world (the all-reduce sums)
  float max_norm;                 // clip_grad_norm_ threshold (<= 0: no clipping)
  float lr, beta1, beta2, eps;
  float bias1, bias2_sqrt;        // 1 - beta1^t, sqrt(1 - beta2^t)
};

// torch.optim.Adam (no weight decay, no amsgrad), single-tensor formulation:
//   m = b1 m + (1 - b1) g;  v = b2 v + (1 - b2) g^2;  p -= (lr / bias1) * m / (sqrt(v) / sqrt(bias2) + eps)
// with g = grad * grad_scale * clip_coef(segment), clip_coef = min(1, max_norm / (norm + 1e-6))   (clip_grad_norm_)
__global__ void __launch_bounds__(kOptThreads) clip_adam_kernel(const AdamArgs a) {
  __shared__ float coef[kMaxSegments];
  if (threadIdx.x < a.so.n) {
    float c = a.grad_scale;
    if (a.max_norm > 0.f) {
      const float norm = (float)sqrt(a.sq[threadIdx.x]);
      const float k = a.max_norm / (norm + 1e-6f);
      c *= k < 1.0f ? k : 1.0f;
    }
    coef[threadIdx.x] = c;
  }
  __syncthreads();
  const long long base = (long long)blockIdx.x * kOptChunk;
  const long long end = base + kOptChunk < a.total ? base + kOptChunk : a.total;
  const float step = a.lr / a.bias1;
  const float inv_b2 = 1.0f / a.bias2_sqrt;
  auto upd = [&](float& p, float g, float& m, float& v, float c) {
    g *= c;
    m = fmaf(a.beta1, m, (1.0f - a.beta1) * g);
    v = fmaf(a.beta2, v, (1.0f - a.beta2) * g * g);
    p -= step * m / (sqrtf(v) * inv_b2 + a.eps);
  };
  int s = find_segment(a.so, base);
  // chunks are multiples of 4 elements and the buffers 16-byte aligned: float4 body, a segment boundary inside a
  // float4 is handled per lane
  for (long long i = base + 4LL * threadIdx.x; i < end; i += 4LL * kOptThreads) {
    while (s + 1 < a.so.n && i >= a.so.off[s + 1]) ++s;
    if (i + 4 <= end) {
      float4 p = *reinterpret_cast<float4*>(a.p + i);
      const float4 g = *reinterpret_cast<const float4*>(a.g + i);
      float4 m = *reinterpret_cast<float4*>(a.m + i);
      float4 v = *reinterpret_cast<float4*>(a.v + i);
      const long long nxt = a.so.off[s + 1];
      const float c0 = coef[s];
      if (i + 4 <= nxt) {
        upd(p.x, g.x, m.x, v.x, c0); upd(p.y, g.y, m.y, v.y, c0); upd(p.z, g.z, m.z, v.z, c0); upd(p.w, g.w, m.w, v.w, c0);
      } else {
        upd(p.x, g.x, m.x, v.x, coef[find_segment(a.so, i)]);
        upd(p.y, g.y, m.y, v.y, coef[find_segment(a.so, i + 1)]);
        upd(p.z, g.z, m.z, v.z, coef[find_segment(a.so, i + 2)]);
        upd(p.w, g.w, m.w, v.w, coef[find_segment(a.so, i + 3)]);
      }
      *reinterpret_cast<float4*>(a.p + i) = p;
      *reinterpret_cast<float4*>(a.m + i) = m;
      *reinterpret_cast<float4*>(a.v + i) = v;
    } else {
      for (long long j = i; j < end; ++j)
        upd(a.p[j], a.g[j], a.m[j], a.v[j], coef[find_segment(a.so, j)]);
    }
  }
}

}  // namespace

extern "C" {

int dpmn_nccl_available(void) { return nccl_api().ok ? 1 : 0; }

int dpmn_nccl_version(void) {
  NcclApi& n = nccl_api();
  int v = 0;
  if (!n.ok || !n.get_version || n.get_version(&v) != 0) return 0;
  return v;
}

int dpmn_nccl_unique_id(void* id128) {
  NcclApi& n = nccl_api();
  if (!n.ok) return DPMN_E_DEVICE;
  if (!id128) return DPMN_E_ARG;
  return n.get_unique_id(reinterpret_cast<NcclUniqueId*>(id128)) == 0 ? 0 : DPMN_E_DEVICE;
}

int dpmn_nccl_comm_init(const void* id128, int32_t world, int32_t rank, void** comm_out) {
  NcclApi& n = nccl_api();
  if (!n.ok) return DPMN_E_DEVICE;
  if (!id128 || !comm_out || world < 1 || rank < 0 || rank >= world) return DPMN_E_ARG;
  NcclUniqueId id;
  memcpy(&id, id128, sizeof(id));
  NcclComm c = nullptr;
  if (n.comm_init_rank(&c, world, id, rank) != 0) return DPMN_E_DEVICE;
  *comm_out = c;
  return 0;
}

int dpmn_nccl_comm_destroy(void* comm) {
  NcclApi& n = nccl_api();
  if (!n.ok) return DPMN_E_DEVICE;
  if (!comm) return DPMN_E_ARG;
  return n.comm_destroy((NcclComm)comm) == 0 ? 0 : DPMN_E_DEVICE;
}

int dpmn_allreduce_bucket(void* comm, void* bucket, size_t count, int32_t dtype, void* stream) {
  NcclApi& n = nccl_api();
  if (!n.ok) return DPMN_E_DEVICE;
  if (!comm || !bucket) return DPMN_E_ARG;
  if (count == 0) return 0;
  const int t = dtype == DPMN_PREC_F32 ? kNcclFloat32 : dtype == DPMN_PREC_F16 ? kNcclFloat16 : dtype == DPMN_PREC_BF16 ? kNcclBfloat16 : -1;
  if (t < 0) return DPMN_E_ARG;
  return n.all_reduce(bucket, bucket, count, t, kNcclSum, (NcclComm)comm, (cudaStream_t)stream) == 0 ? 0 : DPMN_E_DEVICE;
}

size_t dpmn_clip_adam_workspace_bytes(int32_t n_segments) { return n_segments > 0 ? (size_t)n_segments * sizeof(double) : 0; }

int dpmn_clip_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, const int64_t* segment_offsets,
                        int32_t n_segments, float grad_scale, float max_norm, float lr, float beta1, float beta2, float eps,
                        int64_t step, void* workspace, size_t workspace_bytes, void* stream) {
  if (!params || !grads || !exp_avg || !exp_avg_sq || !segment_offsets || !workspace) return DPMN_E_ARG;
  if (n_segments < 1 || n_segments > kMaxSegments || step < 1) return DPMN_E_ARG;
  if (workspace_bytes < dpmn_clip_adam_workspace_bytes(n_segments)) return DPMN_E_WORKSPACE;
  if ((reinterpret_cast<uintptr_t>(params) | reinterpret_cast<uintptr_t>(grads) | reinterpret_cast<uintptr_t>(exp_avg) |
       reinterpret_cast<uintptr_t>(exp_avg_sq)) & 15) return DPMN_E_ARG;
  SegOffsets so;                                   // HOST array -> kernel argument (<= 33 x 8 bytes)
  so.n = n_segments;
  for (int i = 0; i <= n_segments; ++i) {
    so.off[i] = segment_offsets[i];
    if (i > 0 && so.off[i] < so.off[i - 1]) return DPMN_E_ARG;
  }
  if (so.off[0] != 0 || so.off[n_segments] <= 0) return DPMN_E_ARG;
  const long long total = so.off[n_segments];
  cudaStream_t st = (cudaStream_t)stream;
  double* sq = reinterpret_cast<double*>(workspace);
  const unsigned blocks = (unsigned)((total + kOptChunk - 1) / kOptChunk);
  if (max_norm > 0.f) {
    DPMN_CUDA_TRY(cudaMemsetAsync(sq, 0, (size_t)n_segments * sizeof(double), st));
    grad_sqnorm_kernel<<<blocks, kOptThreads, 0, st>>>(grads, so, total, grad_scale, sq);
    DPMN_LAUNCH_CHECK();
  }
  AdamArgs a;
  a.p = params; a.g = grads; a.m = exp_avg; a.v = exp_avg_sq; a.so = so; a.total = total; a.sq = sq;
  a.grad_scale = grad_scale; a.max_norm = max_norm; a.lr = lr; a.beta1 = beta1; a.beta2 = beta2; a.eps = eps;
  a.bias1 = (float)(1.0 - pow((double)beta1, (double)step));
  a.bias2_sqrt = (float)sqrt(1.0 - pow((double)beta2, (double)step));
  clip_adam_kernel<<<blocks, kOptThreads, 0, st>>>(a);
  DPMN_LAUNCH_CHECK();
  return 0;
}

}  // extern "C"
