// Shared helpers of libmpa_b200.so (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <atomic>

#include "../../include/mpa_b200.h"

namespace mpa {

void set_error(const char* fmt, ...);
extern std::atomic<uint64_t> g_launches;
inline void count_launch(int n = 1) { g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }

// RAII CUDA-event pair around a kernel launch; a no-op unless mpa_profile_enable(1)
class ProfScope {
 public:
  ProfScope(const char* name, cudaStream_t stream);
  ~ProfScope();
 private:
  const char* name_;
  cudaStream_t stream_;
  cudaEvent_t a_ = nullptr, b_ = nullptr;
  bool active_ = false;
};

#define MPA_CHECK_ARG(cond, ...)            \
  do {                                      \
    if (!(cond)) {                          \
      ::mpa::set_error(__VA_ARGS__);        \
      return MPA_ERR_INVALID_ARG;           \
    }                                       \
  } while (0)

#define MPA_CUDA(expr)                                                              \
  do {                                                                              \
    cudaError_t _e = (expr);                                                        \
    if (_e != cudaSuccess) {                                                        \
      ::mpa::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e),      \
                       __FILE__, __LINE__);                                         \
      return MPA_ERR_CUDA;                                                          \
    }                                                                               \
  } while (0)

#define MPA_LAUNCH_CHECK()                                                          \
  do {                                                                              \
    ::mpa::count_launch();                                                          \
    MPA_CUDA(cudaGetLastError());                                                   \
  } while (0)

int launch_se3_backward(const float* quat, const float* pts, const float* grad_out,
                        const float* valids, int fill_invalid, int n_parts, int N, float* grad_pts,
                        float* grad_quat, float* grad_trans, cudaStream_t stream);

// csrc/linear.cu: batched [R x R] Gram matrices of the items' rows, fp32-accurate tensor-core mode
int launch_gram_batched(const __nv_bfloat16* x_planes, long long total_rows, int R, int K, int z0, int items,
                        const float* bias, const float* item_valid, float* out, const char* name,
                        cudaStream_t stream);

// ... and the slab-free variant: <= 64 candidate keys per row (1 = shape not supported, use the slab path)
int launch_gram_candidates(const __nv_bfloat16* x_planes, long long total_rows, int R, int K, int items,
                           const float* bias, const float* item_valid, const float* xx, const unsigned* xxmax,
                           float alpha, int k, unsigned long long* cand, int* count, cudaStream_t stream);

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// Function attributes (dynamic shared memory limits) and the SM count are PER DEVICE: a
// process that drives a second GPU must set / query them again.  `DeviceOnce` remembers, per
// call site, on which devices the one-time setup already ran (idempotent, so a race between
// two host threads only repeats it).
inline int current_device() {
  int d = 0;
  cudaGetDevice(&d);
  return d;
}
struct DeviceOnce {
  std::atomic<uint64_t> mask{0};
  bool pending() const { return (mask.load(std::memory_order_acquire) >> (current_device() & 63) & 1ull) == 0; }
  void done() { mask.fetch_or(1ull << (current_device() & 63), std::memory_order_release); }
};
inline int device_sms() {
  static std::atomic<int> cache[64];
  const int d = current_device() & 63;
  int s = cache[d].load(std::memory_order_relaxed);
  if (s == 0) {
    cudaDeviceGetAttribute(&s, cudaDevAttrMultiProcessorCount, d);
    if (s <= 0) s = 148;
    cache[d].store(s, std::memory_order_relaxed);
  }
  return s;
}

// scratch: caller-provided workspace or stream-ordered allocation
struct Scratch {
  void* base = nullptr;
  bool owned = false;
  cudaStream_t stream = nullptr;
  int acquire(void* ws, size_t ws_bytes, size_t need, cudaStream_t s) {
    stream = s;
    if (need == 0) return MPA_OK;
    if (ws != nullptr) {
      if (ws_bytes < need) {
        set_error("workspace too small: %zu < %zu bytes", ws_bytes, need);
        return MPA_ERR_WORKSPACE;
      }
      base = ws;
      return MPA_OK;
    }
    MPA_CUDA(cudaMallocAsync(&base, need, s));
    owned = true;
    return MPA_OK;
  }
  ~Scratch() {
    if (owned && base) cudaFreeAsync(base, stream);
  }
};

// ---- packed FP32x2 math (Blackwell FADD2 / FMUL2 / FFMA2) ----------------
__device__ __forceinline__ unsigned long long f2_as_u64(float2 v) {
  return (unsigned long long)__float_as_uint(v.x) |
         ((unsigned long long)__float_as_uint(v.y) << 32);
}
__device__ __forceinline__ float2 u64_as_f2(unsigned long long v) {
  return make_float2(__uint_as_float((unsigned)(v & 0xffffffffull)),
                     __uint_as_float((unsigned)(v >> 32)));
}
__device__ __forceinline__ unsigned long long add2(unsigned long long a, unsigned long long b) {
  unsigned long long r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ unsigned long long mul2(unsigned long long a, unsigned long long b) {
  unsigned long long r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b,
                                                   unsigned long long c) {
  unsigned long long r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}

// Squared distance in the exact rounding sequence nvcc gives the reference
// expression (chamfer_kernel.cu:80): t = dy*dy; t = fma(dx,dx,t); t = fma(dz,dz,t).
__device__ __forceinline__ float sqdist_ref(float x1, float y1, float z1, float x2, float y2,
                                            float z2) {
  float dx = __fsub_rn(x1, x2), dy = __fsub_rn(y1, y2), dz = __fsub_rn(z1, z2);
  float t = __fmul_rn(dy, dy);
  t = __fmaf_rn(dx, dx, t);
  t = __fmaf_rn(dz, dz, t);
  return t;
}

// Hamilton product in the evaluation order of pytorch3d quaternion_raw_multiply
// (each product and sum rounded separately, left to right; no contraction).
__device__ __forceinline__ void qmul_raw(const float a[4], const float b[4], float o[4]) {
  o[0] = __fsub_rn(__fsub_rn(__fsub_rn(__fmul_rn(a[0], b[0]), __fmul_rn(a[1], b[1])),
                             __fmul_rn(a[2], b[2])),
                   __fmul_rn(a[3], b[3]));
  o[1] = __fsub_rn(__fadd_rn(__fadd_rn(__fmul_rn(a[0], b[1]), __fmul_rn(a[1], b[0])),
                             __fmul_rn(a[2], b[3])),
                   __fmul_rn(a[3], b[2]));
  o[2] = __fadd_rn(__fadd_rn(__fsub_rn(__fmul_rn(a[0], b[2]), __fmul_rn(a[1], b[3])),
                             __fmul_rn(a[2], b[0])),
                   __fmul_rn(a[3], b[1]));
  o[3] = __fadd_rn(__fsub_rn(__fadd_rn(__fmul_rn(a[0], b[3]), __fmul_rn(a[1], b[2])),
                             __fmul_rn(a[2], b[1])),
                   __fmul_rn(a[3], b[0]));
}

// qrot/qtransform of one point (utils/transforms.py:75-109): q (0,v) conj(q), + t.
__device__ __forceinline__ float3 se3_apply(const float q[4], const float* t, float3 v) {
  const float qc[4] = {q[0], -q[1], -q[2], -q[3]};
  const float pv[4] = {0.0f, v.x, v.y, v.z};
  float a[4], b[4];
  qmul_raw(q, pv, a);
  qmul_raw(a, qc, b);
  float3 o = make_float3(b[1], b[2], b[3]);
  if (t != nullptr) {
    o.x = __fadd_rn(o.x, t[0]);
    o.y = __fadd_rn(o.y, t[1]);
    o.z = __fadd_rn(o.z, t[2]);
  }
  return o;
}

}  // namespace mpa
