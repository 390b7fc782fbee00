// PointNet++ set-abstraction primitives for sm_100a: furthest point sampling, ball query and the
// grouping gather that feeds the shared MLP.
//
// Replace the reference's pointnet2_ops CUDA extension
//   models/modules/encoder/pointnet2/pointnet2_ops_lib/pointnet2_ops/_ext-src/src/
//     sampling_gpu.cu:74-177   furthest_point_sampling_kernel
//     ball_query_gpu.cu:13-48  query_ball_point_kernel
//     group_points_gpu.cu:12-32 group_points_kernel (+ the xyz recentring / concatenation of
//                              pointnet2_utils.py:309-346 QueryAndGroup, :362-392 GroupAll)
// with the same results: identical arithmetic (the FMA contraction nvcc gives the reference's
// distance expression: t = dy*dy; fma(dx,dx,t); fma(dz,dz,t)), identical tie rules (see each
// kernel).  The shared MLPs run on the tcgen05 GEMM (csrc/linear.cu) with the BatchNorm / ReLU /
// max-pool passes of csrc/knn.cu.
#include "mpa_common.cuh"

namespace mpa {

// ---- furthest point sampling ------------------------------------------------------------
// One CTA per cloud; the cloud (x, y, z, running minimum distance) lives in shared memory.
// Semantics of sampling_gpu.cu:74-177: start from point 0; every round each point's distance
// to the last pick lowers its running minimum (initially 1e10, sampling.cpp:75) and the point
// with the LARGEST minimum is picked; points with |p|^2 <= 1e-3 are skipped (:100-101) --
// they are never updated nor picked.  Ties: the reference reduces per-thread strided maxima
// (first maximum of the stride, thread t = k mod block, block = largest power of two <= n
// capped at 512, cuda_utils.h:15-19) with a shared-memory tree that folds slot t + s onto
// slot t and keeps slot t on equality (__update, :64-71): two tied slots meet at the stage of
// their lowest differing bit and the one with a 0 there survives -- the smaller BIT-REVERSED
// thread index wins, then the smaller k.  That order is encoded in the low bits of the
// reduction key here, so the result does not depend on THIS kernel's thread count.
constexpr int FPS_THREADS = 256;
__device__ __forceinline__ unsigned long long fps_key(float d, int k, int ref_bits) {
  // d >= 0: float bits are order preserving; larger key wins: larger d, then smaller
  // (bit-reversed k mod block, k)
  const unsigned t = (unsigned)k & ((1u << ref_bits) - 1u);
  const unsigned tie = ((ref_bits ? (__brev(t) >> (32 - ref_bits)) : 0u) << 16) | (unsigned)k;
  return ((unsigned long long)__float_as_uint(d) << 32) | (unsigned long long)(~tie);
}
__global__ void __launch_bounds__(FPS_THREADS)
fps_kernel(const float* __restrict__ xyz, int n, int m, int ref_bits, int* __restrict__ idxs,
           float* __restrict__ new_xyz) {
  extern __shared__ float sm[];  // x[n] y[n] z[n] temp[n]
  float* sx = sm;
  float* sy = sx + n;
  float* sz = sy + n;
  float* st = sz + n;
  __shared__ unsigned long long s_best[FPS_THREADS / 32];
  __shared__ int s_old;
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float* p = xyz + (long long)b * n * 3;
  for (int k = tid; k < n; k += FPS_THREADS) {
    sx[k] = p[3 * k]; sy[k] = p[3 * k + 1]; sz[k] = p[3 * k + 2];
    st[k] = 1e10f;
  }
  if (tid == 0) {
    s_old = 0;
    idxs[(long long)b * m] = 0;
  }
  __syncthreads();
  for (int j = 1; j < m; ++j) {
    const int old = s_old;
    const float x1 = sx[old], y1 = sy[old], z1 = sz[old];
    unsigned long long best = 0ull;  // "none": the reference then keeps index 0 (besti = 0, :93)
    for (int k = tid; k < n; k += FPS_THREADS) {
      const float x2 = sx[k], y2 = sy[k], z2 = sz[k];
      const float mag = __fmaf_rn(z2, z2, __fmaf_rn(x2, x2, __fmul_rn(y2, y2)));
      if (mag <= 1e-3f) continue;
      const float d = sqdist_ref(x2, y2, z2, x1, y1, z1);
      const float d2 = fminf(d, st[k]);
      st[k] = d2;
      const unsigned long long key = fps_key(d2, k, ref_bits);
      best = key > best ? key : best;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const unsigned long long other = __shfl_xor_sync(0xffffffffu, best, o);
      best = other > best ? other : best;
    }
    if (lane == 0) s_best[warp] = best;
    __syncthreads();
    if (tid == 0) {
      unsigned long long bb = s_best[0];
#pragma unroll
      for (int w = 1; w < FPS_THREADS / 32; ++w) bb = s_best[w] > bb ? s_best[w] : bb;
      const int pick = bb == 0ull ? 0 : (int)((~(unsigned)(bb & 0xffffffffull)) & 0xffffu);
      s_old = pick;
      idxs[(long long)b * m + j] = pick;
    }
    __syncthreads();
  }
  // gather_operation (pointnet2_modules.py:53-61): the sampled centroids
  if (new_xyz != nullptr) {
    for (int j = tid; j < m; j += FPS_THREADS) {
      const int k = idxs[(long long)b * m + j];
      float* o = new_xyz + ((long long)b * m + j) * 3;
      o[0] = sx[k]; o[1] = sy[k]; o[2] = sz[k];
    }
  }
}

// ---- ball query -----------------------------------------------------------------------
// ball_query_gpu.cu:13-48: for centroid j the first `nsample` points (in index order) with
// d2 < radius^2; unused slots repeat the first hit; no hit at all leaves zeros.  One warp per
// centroid: 32 consecutive points per step, ballot + prefix keep the index order.
__global__ void ball_query_kernel(const float* __restrict__ xyz, const float* __restrict__ new_xyz, int B, int n,
                                  int m, float radius, int nsample, int* __restrict__ idx) {
  const int lane = threadIdx.x & 31;
  const long long w = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;  // (cloud, centroid)
  const int b = (int)(w / m), j = (int)(w % m);
  if (b >= B) return;  // warp-uniform
  const float* p = xyz + (long long)b * n * 3;
  const float* c = new_xyz + ((long long)b * m + j) * 3;
  int* out = idx + ((long long)b * m + j) * nsample;
  const float cx = c[0], cy = c[1], cz = c[2];
  const float r2 = __fmul_rn(radius, radius);
  int cnt = 0, first = 0;
  for (int k0 = 0; k0 < n && cnt < nsample; k0 += 32) {
    const int k = k0 + lane;
    bool hit = false;
    if (k < n) hit = sqdist_ref(cx, cy, cz, p[3 * k], p[3 * k + 1], p[3 * k + 2]) < r2;
    const unsigned mask = __ballot_sync(0xffffffffu, hit);
    if (mask != 0u) {
      if (cnt == 0) first = k0 + __ffs(mask) - 1;
      const int pos = cnt + __popc(mask & ((1u << lane) - 1u));
      if (hit && pos < nsample) out[pos] = k;
      cnt += __popc(mask);
    }
  }
  cnt = min(cnt, nsample);
  for (int s = cnt + lane; s < nsample; s += 32) out[s] = first;  // zeros when nothing was found
}

// ---- grouping: rows of the shared MLP's input matrix -------------------------------------
// QueryAndGroup (pointnet2_utils.py:309-346): row (b, j, s) = [ xyz[b, idx] - new_xyz[b, j] |
// features[b, idx, :] ]; GroupAll (:362-392, idx == nullptr): row (b, k) = [ xyz[b, k] |
// features[b, k, :] ].  Features are channels-last [B, n, C]; rows are padded with zeros to
// `ld` columns (the GEMM wants K % 8 == 0).
__global__ void group_rows_kernel(const float* __restrict__ xyz, const float* __restrict__ new_xyz,
                                  const float* __restrict__ feats, const int* __restrict__ idx, int n, int m,
                                  int nsample, int C, int ld, long long rows, float* __restrict__ out) {
  const long long total = rows * ld;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total;
       e += (long long)gridDim.x * blockDim.x) {
    const long long r = e / ld;
    const int c = (int)(e % ld);
    float v = 0.f;
    if (c < 3 + C) {
      long long b, k;
      if (idx != nullptr) {
        const long long g = r / nsample;  // (b, j)
        b = g / m;
        k = idx[r];
        if (c < 3) v = __fsub_rn(xyz[(b * n + k) * 3 + c], new_xyz[g * 3 + c]);
      } else {
        b = r / n;
        k = r % n;
        if (c < 3) v = xyz[(b * n + k) * 3 + c];
      }
      if (c >= 3) v = feats[(b * n + k) * C + (c - 3)];
    }
    out[e] = v;
  }
}

}  // namespace mpa

using namespace mpa;

extern "C" {

int mpa_furthest_point_sample(const float* xyz, int B, int n, int m, int32_t* idx, float* new_xyz,
                              void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  MPA_CHECK_ARG(B >= 0 && n > 0 && m > 0 && n <= 65535, "furthest_point_sample: bad sizes B=%d n=%d m=%d", B, n, m);
  if (B == 0) return MPA_OK;
  MPA_CHECK_ARG(xyz && idx, "furthest_point_sample: null pointer");
  int ref_bits = 0;
  while ((2 << ref_bits) <= n && ref_bits < 9) ++ref_bits;  // block = 1 << ref_bits = opt_n_threads(n), cuda_utils.h:15-19
  const size_t smem = sizeof(float) * 4 * (size_t)n;
  MPA_CHECK_ARG(smem <= 200 * 1024, "furthest_point_sample: clouds of at most %d points", 200 * 1024 / 16);
  static DeviceOnce attr;
  if (smem > 48 * 1024 && attr.pending()) {
    MPA_CUDA(cudaFuncSetAttribute(fps_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr.done();
  }
  {
    ProfScope ps("pointnet2_fps", stream);
    fps_kernel<<<B, FPS_THREADS, smem, stream>>>(xyz, n, m, ref_bits, idx, new_xyz);
  }
  MPA_LAUNCH_CHECK();
  return MPA_OK;
}

int mpa_ball_query(const float* xyz, const float* new_xyz, int B, int n, int m, float radius, int nsample,
                   int32_t* idx, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  MPA_CHECK_ARG(B >= 0 && n > 0 && m > 0 && nsample > 0, "ball_query: bad sizes");
  if (B == 0) return MPA_OK;
  MPA_CHECK_ARG(xyz && new_xyz && idx, "ball_query: null pointer");
  MPA_CUDA(cudaMemsetAsync(idx, 0, sizeof(int32_t) * (size_t)B * m * nsample, stream));  // ball_query.cpp: zeros
  const long long warps = (long long)B * m;
  const unsigned blocks = (unsigned)((warps + 7) / 8);
  {
    ProfScope ps("pointnet2_ball_query", stream);
    ball_query_kernel<<<blocks, 256, 0, stream>>>(xyz, new_xyz, B, n, m, radius, nsample, idx);
  }
  MPA_LAUNCH_CHECK();
  return MPA_OK;
}

int mpa_group_rows(const float* xyz, const float* new_xyz, const float* feats, const int32_t* idx, int B, int n,
                   int m, int nsample, int C, int ld, float* out, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  MPA_CHECK_ARG(B >= 0 && n > 0 && C >= 0 && ld >= 3 + C, "group_rows: bad sizes");
  MPA_CHECK_ARG(idx == nullptr || (m > 0 && nsample > 0 && new_xyz != nullptr), "group_rows: grouping needs centroids");
  if (B == 0) return MPA_OK;
  MPA_CHECK_ARG(xyz && out && (C == 0 || feats), "group_rows: null pointer");
  const long long rows = idx != nullptr ? (long long)B * m * nsample : (long long)B * n;
  const long long total = rows * ld;
  long long blocks = (total + 255) / 256;
  const long long cap = (long long)device_sms() * 16;
  if (blocks > cap) blocks = cap;
  {
    ProfScope ps("pointnet2_group", stream);
    group_rows_kernel<<<(unsigned)blocks, 256, 0, stream>>>(xyz, new_xyz, feats, idx, n, m, nsample, C, ld, rows, out);
  }
  MPA_LAUNCH_CHECK();
  return MPA_OK;
}

}  // extern "C"
