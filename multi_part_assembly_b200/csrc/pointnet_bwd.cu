// Training-step backward of the PointNet part encoder: the BatchNorm / ReLU /
// max-pool pieces between the (library) GEMMs.
//
// The reference differentiates PointNet.forward (models/modules/encoder/pointnet.py:
// 29-41) through cuDNN BatchNorm kernels on [n, C, N] tensors; on a B200 those
// BatchNorm forward/backward kernels alone take ~11 ms per 32-shape step.  Here the
// activations are kept point-major ([M = n*N, C] bf16, C contiguous), so that
//   * the five 1x1 convolutions and their two backward products are plain row-major
//     GEMMs (dW = dz^T a, da = dz W), and
//   * train-mode BatchNorm is two streaming passes per direction:
//       forward  : column sums of z, z^2        -> scale/shift  -> a = relu(z*scale+shift)
//       backward : column sums of dy, dy*zhat   -> dz = gamma*rstd*(dy - S1/M - zhat*S2/M)
//     with dy = da * [y > 0] (layers 1-4) or the max-pool scatter of the feature
//     gradient (layer 5: dy[m, c] = g[part, c] if m is the arg-max point of (part, c)).
// Every kernel reads/writes each element once with 16-byte packets: HBM-bound
// (2-6 B per element and pass).  Padded parts (valids == 0) are excluded from the
// statistics and get zero gradients, like the compaction of
// models/pn_transformer/network.py:59-68.
#include <cuda_bf16.h>

#include "mpa_common.cuh"

namespace mpa {

constexpr int PB_THREADS = 256;

struct Bf8 {
  float v[8];
};
__device__ __forceinline__ Bf8 load_bf8(const __nv_bfloat16* p) {
  const uint4 u = *reinterpret_cast<const uint4*>(p);
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
  Bf8 r;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 f = __bfloat1622float2(h[i]);
    r.v[2 * i] = f.x;
    r.v[2 * i + 1] = f.y;
  }
  return r;
}
constexpr int PB_UNROLL = 4;  // rows (16-byte packets) a thread keeps in flight
__device__ __forceinline__ uint4 ld_packet(const __nv_bfloat16* p) {
  return *reinterpret_cast<const uint4*>(p);
}
__device__ __forceinline__ Bf8 unpack_bf8(const uint4& u) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
  Bf8 r;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 f = __bfloat1622float2(h[i]);
    r.v[2 * i] = f.x;
    r.v[2 * i + 1] = f.y;
  }
  return r;
}
__device__ __forceinline__ void store_bf8(__nv_bfloat16* p, const Bf8& r) {
  uint4 u;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(r.v[2 * i], r.v[2 * i + 1]);
  *reinterpret_cast<uint4*>(p) = u;
}

// per-channel constants of one BatchNorm layer (all [C] fp32, device)
struct BnConst {
  const float* mean;
  const float* rstd;
  const float* gamma;
  const float* beta;
};

// Thread layout shared by all kernels: a thread owns 8 consecutive channels (one
// 16-byte packet); C/8 threads cover a row; a CTA walks rows with stride
// rows_per_iter * gridDim.x.  Column reductions: per-thread fp32 partials over at
// most a few dozen rows, combined per CTA in shared memory, then one fp64 atomic per
// channel and CTA.
template <int NACC>
__device__ __forceinline__ void cta_column_reduce(float (&acc)[NACC][8], int C, double* out /* [NACC][C] */) {
  __shared__ float red[PB_THREADS * 8];
  const int tpr = C >> 3;                 // threads per row
  const int col = threadIdx.x % tpr;      // packet index inside the row
  const int rows_per_iter = PB_THREADS / tpr;
#pragma unroll
  for (int a = 0; a < NACC; ++a) {
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 8; ++i) red[threadIdx.x * 8 + i] = acc[a][i];
    __syncthreads();
    if (threadIdx.x < tpr) {  // first row-group sums the others (fixed order)
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        float s = 0.f;
        for (int r = 0; r < rows_per_iter; ++r) s += red[(r * tpr + col) * 8 + i];
        atomicAdd(out + (size_t)a * C + col * 8 + i, (double)s);
      }
    }
  }
}

// sums[0][c] = sum_m z[m,c], sums[1][c] = sum_m z[m,c]^2 over the valid rows
__global__ void __launch_bounds__(PB_THREADS, 4)
bn_stats_kernel(const __nv_bfloat16* __restrict__ z, long long M, int C, int N,
                const float* __restrict__ valids, double* __restrict__ sums) {
  const int tpr = C >> 3, col = threadIdx.x % tpr, rows_per_iter = PB_THREADS / tpr;
  float acc[2][8] = {};
  const long long stride = (long long)gridDim.x * rows_per_iter;
  const long long rstride = stride * C;
  long long base = ((long long)blockIdx.x * rows_per_iter + threadIdx.x / tpr) * C + col * 8;
  for (long long m0 = (long long)blockIdx.x * rows_per_iter + threadIdx.x / tpr; m0 < M;
       m0 += PB_UNROLL * stride, base += PB_UNROLL * rstride) {
    uint4 raw[PB_UNROLL];
    bool ok[PB_UNROLL];
#pragma unroll
    for (int u = 0; u < PB_UNROLL; ++u) {  // all loads first: PB_UNROLL packets in flight per thread
      const long long m = m0 + u * stride;
      ok[u] = m < M && (valids == nullptr || valids[(unsigned)m / (unsigned)N] != 0.f);
      raw[u] = ok[u] ? ld_packet(z + (base + u * rstride)) : make_uint4(0u, 0u, 0u, 0u);
    }
#pragma unroll
    for (int u = 0; u < PB_UNROLL; ++u) {
      const Bf8 v = unpack_bf8(raw[u]);  // zeros for skipped rows
#pragma unroll
      for (int i = 0; i < 8; ++i) { acc[0][i] += v.v[i]; acc[1][i] = fmaf(v.v[i], v.v[i], acc[1][i]); }
    }
  }
  cta_column_reduce<2>(acc, C, sums);
}

// sums -> mean, rstd, scale = gamma*rstd, shift = beta - mean*scale; count = N * #valid parts
__global__ void bn_finalize_kernel(const double* __restrict__ sums, int C, int n_parts, int N,
                                   const float* __restrict__ valids, const float* __restrict__ gamma,
                                   const float* __restrict__ beta, float eps, float* __restrict__ mean,
                                   float* __restrict__ rstd, float* __restrict__ scale,
                                   float* __restrict__ shift, float* __restrict__ count_out) {
  __shared__ int s_valid;
  if (threadIdx.x == 0) s_valid = 0;
  __syncthreads();
  int nv = 0;
  for (int p = threadIdx.x; p < n_parts; p += blockDim.x) nv += (valids == nullptr || valids[p] != 0.f) ? 1 : 0;
  if (nv) atomicAdd(&s_valid, nv);  // integers: exact in any order
  __syncthreads();
  const double s_cnt = (double)s_valid * (double)N;
  if (count_out != nullptr && blockIdx.x == 0 && threadIdx.x == 0) *count_out = (float)s_cnt;
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const double n = s_cnt > 0.0 ? s_cnt : 1.0;
  const double mu = sums[c] / n;
  const double var = fmax(sums[C + c] / n - mu * mu, 0.0);
  const float r = (float)(1.0 / sqrt(var + (double)eps));
  mean[c] = (float)mu;
  rstd[c] = r;
  const float sc = gamma[c] * r;
  scale[c] = sc;
  shift[c] = beta[c] - (float)mu * sc;
}

// a = relu?(z * scale + shift), bf16 (rows of padded parts are written as zeros)
__global__ void __launch_bounds__(PB_THREADS, 4)
bn_act_kernel(const __nv_bfloat16* __restrict__ z, const float* __restrict__ scale,
              const float* __restrict__ shift, int relu, long long M, int C, int N,
              const float* __restrict__ valids, __nv_bfloat16* __restrict__ a) {
  const int tpr = C >> 3, col = threadIdx.x % tpr, rows_per_iter = PB_THREADS / tpr;
  float sc[8], sh[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { sc[i] = scale[col * 8 + i]; sh[i] = shift[col * 8 + i]; }
  const long long stride = (long long)gridDim.x * rows_per_iter;
  const long long rstride = stride * C;
  long long base = ((long long)blockIdx.x * rows_per_iter + threadIdx.x / tpr) * C + col * 8;
  for (long long m0 = (long long)blockIdx.x * rows_per_iter + threadIdx.x / tpr; m0 < M;
       m0 += PB_UNROLL * stride, base += PB_UNROLL * rstride) {
    uint4 raw[PB_UNROLL];
    bool ok[PB_UNROLL];
#pragma unroll
    for (int u = 0; u < PB_UNROLL; ++u) {
      const long long m = m0 + u * stride;
      ok[u] = m < M && (valids == nullptr || valids[(unsigned)m / (unsigned)N] != 0.f);
      raw[u] = ok[u] ? ld_packet(z + (base + u * rstride)) : make_uint4(0u, 0u, 0u, 0u);
    }
#pragma unroll
    for (int u = 0; u < PB_UNROLL; ++u) {
      const long long m = m0 + u * stride;
      if (m >= M) break;
      Bf8 v = unpack_bf8(raw[u]);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        float y = fmaf(v.v[i], sc[i], sh[i]);
        if (relu) y = fmaxf(y, 0.f);
        v.v[i] = ok[u] ? y : 0.f;
      }
      store_bf8(a + (base + u * rstride), v);
    }
  }
}

// incoming gradient of the BatchNorm output at (row m, channels col*8..):
//   da != nullptr : dy = da * [z*scale+shift > 0]   (ReLU layers)
//   else          : dy = g[part, c] where arg[part, c] == m - part*N   (max-pool, layer 5)
__device__ __forceinline__ Bf8 incoming_grad(bool has_da, const uint4& da_raw,
                                             const float* __restrict__ g, const int* __restrict__ arg,
                                             const Bf8& zv, const float (&sc)[8], const float (&sh)[8],
                                             long long m, int C, int N, int col) {
  Bf8 dy;
  if (has_da) {
    dy = unpack_bf8(da_raw);
#pragma unroll
    for (int i = 0; i < 8; ++i)
      if (!(fmaf(zv.v[i], sc[i], sh[i]) > 0.f)) dy.v[i] = 0.f;
  } else {
    const long long part = (unsigned)m / (unsigned)N;  // rows < 2^31 (checked on the host)
    const int local = (int)(m - part * N);
    const int4 a0 = *reinterpret_cast<const int4*>(arg + part * C + col * 8);
    const int4 a1 = *reinterpret_cast<const int4*>(arg + part * C + col * 8 + 4);
    const float4 g0 = *reinterpret_cast<const float4*>(g + part * C + col * 8);
    const float4 g1 = *reinterpret_cast<const float4*>(g + part * C + col * 8 + 4);
    dy.v[0] = a0.x == local ? g0.x : 0.f; dy.v[1] = a0.y == local ? g0.y : 0.f;
    dy.v[2] = a0.z == local ? g0.z : 0.f; dy.v[3] = a0.w == local ? g0.w : 0.f;
    dy.v[4] = a1.x == local ? g1.x : 0.f; dy.v[5] = a1.y == local ? g1.y : 0.f;
    dy.v[6] = a1.z == local ? g1.z : 0.f; dy.v[7] = a1.w == local ? g1.w : 0.f;
  }
  return dy;
}

// Pass 1 of the BatchNorm backward: sums[0][c] = sum dy, sums[1][c] = sum dy * z (raw z:
// the centring and 1/sigma are applied in fp64 by bn_bwd_fix_kernel, which keeps this
// kernel's per-thread state to the 16 mask constants and 16 accumulators).
constexpr int PB_BWD_UNROLL = 2;
__global__ void __launch_bounds__(PB_THREADS, 3)
bn_bwd_reduce_kernel(const __nv_bfloat16* __restrict__ da, const float* __restrict__ g,
                     const int* __restrict__ arg, const __nv_bfloat16* __restrict__ z, BnConst bn,
                     long long M, int C, int N, const float* __restrict__ valids,
                     double* __restrict__ sums) {
  const int tpr = C >> 3, col = threadIdx.x % tpr, rows_per_iter = PB_THREADS / tpr;
  float sc[8], sh[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int c = col * 8 + i;
    sc[i] = bn.gamma[c] * bn.rstd[c];
    sh[i] = bn.beta[c] - bn.mean[c] * sc[i];
  }
  float acc[2][8] = {};
  const long long stride = (long long)gridDim.x * rows_per_iter;
  const long long rstride = stride * C;
  long long base = ((long long)blockIdx.x * rows_per_iter + threadIdx.x / tpr) * C + col * 8;
  for (long long m0 = (long long)blockIdx.x * rows_per_iter + threadIdx.x / tpr; m0 < M;
       m0 += PB_BWD_UNROLL * stride, base += PB_BWD_UNROLL * rstride) {
    uint4 zr[PB_BWD_UNROLL], dr[PB_BWD_UNROLL];
    bool ok[PB_BWD_UNROLL];
#pragma unroll
    for (int u = 0; u < PB_BWD_UNROLL; ++u) {
      const long long m = m0 + u * stride;
      ok[u] = m < M && (valids == nullptr || valids[(unsigned)m / (unsigned)N] != 0.f);
      zr[u] = ok[u] ? ld_packet(z + (base + u * rstride)) : make_uint4(0u, 0u, 0u, 0u);
      dr[u] = (ok[u] && da != nullptr) ? ld_packet(da + (base + u * rstride)) : make_uint4(0u, 0u, 0u, 0u);
    }
#pragma unroll
    for (int u = 0; u < PB_BWD_UNROLL; ++u) {
      if (!ok[u]) continue;
      const long long m = m0 + u * stride;
      const Bf8 zv = unpack_bf8(zr[u]);
      const Bf8 dy = incoming_grad(da != nullptr, dr[u], g, arg, zv, sc, sh, m, C, N, col);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        acc[0][i] += dy.v[i];
        acc[1][i] = fmaf(dy.v[i], zv.v[i], acc[1][i]);
      }
    }
  }
  cta_column_reduce<2>(acc, C, sums);
}

// sums[1][c] <- rstd * (sum dy*z - mean * sum dy) = sum dy * zhat  (= d gamma), in fp64
__global__ void bn_bwd_fix_kernel(double* __restrict__ sums, int C, const float* __restrict__ mean,
                                  const float* __restrict__ rstd) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < C) sums[C + c] = (double)rstd[c] * (sums[C + c] - (double)mean[c] * sums[c]);
}

// Pass 2: dz = gamma*rstd*(dy - S1/n - zhat*S2/n) = A*dy + B + D*z with per-channel
// A = gamma*rstd, D = -A*rstd*S2/n, B = -A*S1/n - D*mean; bf16; zero rows for padded parts
__global__ void __launch_bounds__(PB_THREADS, 3)
bn_bwd_apply_kernel(const __nv_bfloat16* __restrict__ da, const float* __restrict__ g,
                    const int* __restrict__ arg, const __nv_bfloat16* __restrict__ z, BnConst bn,
                    const double* __restrict__ sums, const float* __restrict__ count, long long M,
                    int C, int N, const float* __restrict__ valids, __nv_bfloat16* __restrict__ dz) {
  const int tpr = C >> 3, col = threadIdx.x % tpr, rows_per_iter = PB_THREADS / tpr;
  float sc[8], sh[8], kb[8], kd[8];
  const float inv_n = 1.f / fmaxf(*count, 1.f);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int c = col * 8 + i;
    const float mu = bn.mean[c], rs = bn.rstd[c];
    sc[i] = bn.gamma[c] * rs;
    sh[i] = bn.beta[c] - mu * sc[i];
    kd[i] = -sc[i] * rs * ((float)sums[C + c] * inv_n);
    kb[i] = -sc[i] * ((float)sums[c] * inv_n) - kd[i] * mu;
  }
  const long long stride = (long long)gridDim.x * rows_per_iter;
  const long long rstride = stride * C;
  long long base = ((long long)blockIdx.x * rows_per_iter + threadIdx.x / tpr) * C + col * 8;
  for (long long m0 = (long long)blockIdx.x * rows_per_iter + threadIdx.x / tpr; m0 < M;
       m0 += PB_BWD_UNROLL * stride, base += PB_BWD_UNROLL * rstride) {
    uint4 zr[PB_BWD_UNROLL], dr[PB_BWD_UNROLL];
    bool ok[PB_BWD_UNROLL];
#pragma unroll
    for (int u = 0; u < PB_BWD_UNROLL; ++u) {
      const long long m = m0 + u * stride;
      ok[u] = m < M && (valids == nullptr || valids[(unsigned)m / (unsigned)N] != 0.f);
      zr[u] = ok[u] ? ld_packet(z + (base + u * rstride)) : make_uint4(0u, 0u, 0u, 0u);
      dr[u] = (ok[u] && da != nullptr) ? ld_packet(da + (base + u * rstride)) : make_uint4(0u, 0u, 0u, 0u);
    }
#pragma unroll
    for (int u = 0; u < PB_BWD_UNROLL; ++u) {
      const long long m = m0 + u * stride;
      if (m >= M) break;
      Bf8 out;
      if (!ok[u]) {
#pragma unroll
        for (int i = 0; i < 8; ++i) out.v[i] = 0.f;
      } else {
        const Bf8 zv = unpack_bf8(zr[u]);
        const Bf8 dy = incoming_grad(da != nullptr, dr[u], g, arg, zv, sc, sh, m, C, N, col);
#pragma unroll
        for (int i = 0; i < 8; ++i) out.v[i] = fmaf(sc[i], dy.v[i], fmaf(kd[i], zv.v[i], kb[i]));
      }
      store_bf8(dz + (base + u * rstride), out);
    }
  }
}

// arg[part, c] = index of the first point maximising scale[c] * z[part*N + i, c]
// (the sign of the BatchNorm scale decides between max and min of z); one CTA per part.
__global__ void __launch_bounds__(PB_THREADS)
pool_argmax_kernel(const __nv_bfloat16* __restrict__ z, const float* __restrict__ scale, int N, int C,
                   int* __restrict__ arg) {
  __shared__ float s_val[PB_THREADS * 8];
  __shared__ int s_idx[PB_THREADS * 8];
  const int part = blockIdx.x;
  const int tpr = C >> 3, col = threadIdx.x % tpr, rows_per_iter = PB_THREADS / tpr;
  float sgn[8], best[8];
  int bi[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    sgn[i] = scale[col * 8 + i] >= 0.f ? 1.f : -1.f;
    best[i] = -3.0e38f;
    bi[i] = 0;
  }
  for (int r = threadIdx.x / tpr; r < N; r += rows_per_iter) {
    const Bf8 v = load_bf8(z + ((long long)part * N + r) * C + col * 8);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float y = v.v[i] * sgn[i];
      if (y > best[i]) { best[i] = y; bi[i] = r; }
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) { s_val[threadIdx.x * 8 + i] = best[i]; s_idx[threadIdx.x * 8 + i] = bi[i]; }
  __syncthreads();
  if (threadIdx.x < tpr) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float b = -3.0e38f;
      int bidx = 0;
      for (int r = 0; r < rows_per_iter; ++r) {
        const float v = s_val[(r * tpr + col) * 8 + i];
        const int id = s_idx[(r * tpr + col) * 8 + i];
        if (v > b || (v == b && id < bidx)) { b = v; bidx = id; }
      }
      arg[(long long)part * C + col * 8 + i] = bidx;
    }
  }
}

static int pb_grid(long long M, int C, int ctas_per_sm) {
  const int rows_per_iter = PB_THREADS / (C >> 3);
  long long blocks = (M + rows_per_iter - 1) / rows_per_iter;
  const long long cap = 148LL * ctas_per_sm;  // one resident wave; bounds the fp64 atomics per channel
  return (int)(blocks < cap ? (blocks > 0 ? blocks : 1) : cap);
}

}  // namespace mpa

using namespace mpa;

#define PB_CHECK_SHAPE(what)                                                                        \
  MPA_CHECK_ARG(M >= 0 && M < (1ll << 31) && N > 0 && (C == 64 || C == 128 || C == 256),                       \
                what ": unsupported shape M=%lld C=%d N=%d (rows must fit 31 bits)", M, C, N)

extern "C" {

int mpa_bn_stats(const void* z, long long M, int C, int N, const float* valids, double* sums,
                 void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  PB_CHECK_SHAPE("bn_stats");
  MPA_CHECK_ARG(sums != nullptr && (M == 0 || z != nullptr), "bn_stats: null pointer");
  MPA_CUDA(cudaMemsetAsync(sums, 0, sizeof(double) * 2 * C, stream));
  if (M == 0) return MPA_OK;
  {
    ProfScope ps("bn_stats", stream);
    bn_stats_kernel<<<pb_grid(M, C, 4), PB_THREADS, 0, stream>>>((const __nv_bfloat16*)z, M, C, N, valids, sums);
  }
  MPA_LAUNCH_CHECK();
  return MPA_OK;
}

int mpa_bn_finalize(const double* sums, int C, int n_parts, int N, const float* valids,
                    const float* gamma, const float* beta, float eps, float* mean, float* rstd,
                    float* scale, float* shift, float* count, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  MPA_CHECK_ARG(sums && gamma && beta && mean && rstd && scale && shift && C > 0 && n_parts >= 0,
                "bn_finalize: bad arguments");
  {
    ProfScope ps("bn_finalize", stream);
    bn_finalize_kernel<<<(C + 127) / 128, 128, 0, stream>>>(sums, C, n_parts, N, valids, gamma, beta, eps,
                                                           mean, rstd, scale, shift, count);
  }
  MPA_LAUNCH_CHECK();
  return MPA_OK;
}

int mpa_bn_act(const void* z, const float* scale, const float* shift, int relu, long long M, int C,
               int N, const float* valids, void* a, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  PB_CHECK_SHAPE("bn_act");
  if (M == 0) return MPA_OK;
  MPA_CHECK_ARG(z && scale && shift && a, "bn_act: null pointer");
  {
    ProfScope ps("bn_act", stream);
    bn_act_kernel<<<pb_grid(M, C, 4), PB_THREADS, 0, stream>>>((const __nv_bfloat16*)z, scale, shift, relu, M, C,
                                                           N, valids, (__nv_bfloat16*)a);
  }
  MPA_LAUNCH_CHECK();
  return MPA_OK;
}

int mpa_bn_backward(const void* da, const float* g, const int32_t* arg, const void* z,
                    const float* mean, const float* rstd, const float* gamma, const float* beta,
                    const float* count, long long M, int C, int N, const float* valids, double* sums,
                    void* dz, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  PB_CHECK_SHAPE("bn_backward");
  MPA_CHECK_ARG(sums && mean && rstd && gamma && beta && count, "bn_backward: null pointer");
  MPA_CHECK_ARG((da != nullptr) != (g != nullptr && arg != nullptr),
                "bn_backward: pass either da (ReLU layer) or g + arg (max-pooled layer)");
  MPA_CUDA(cudaMemsetAsync(sums, 0, sizeof(double) * 2 * C, stream));
  if (M == 0) return MPA_OK;
  MPA_CHECK_ARG(z && dz, "bn_backward: null pointer");
  const BnConst bn{mean, rstd, gamma, beta};
  {
    ProfScope ps("bn_bwd_reduce", stream);
    bn_bwd_reduce_kernel<<<pb_grid(M, C, 3), PB_THREADS, 0, stream>>>(
        (const __nv_bfloat16*)da, g, arg, (const __nv_bfloat16*)z, bn, M, C, N, valids, sums);
  }
  MPA_LAUNCH_CHECK();
  bn_bwd_fix_kernel<<<(C + 127) / 128, 128, 0, stream>>>(sums, C, mean, rstd);
  MPA_LAUNCH_CHECK();
  {
    ProfScope ps("bn_bwd_apply", stream);
    bn_bwd_apply_kernel<<<pb_grid(M, C, 3), PB_THREADS, 0, stream>>>(
        (const __nv_bfloat16*)da, g, arg, (const __nv_bfloat16*)z, bn, sums, count, M, C, N, valids,
        (__nv_bfloat16*)dz);
  }
  MPA_LAUNCH_CHECK();
  return MPA_OK;
}

int mpa_pool_argmax(const void* z, const float* scale, int n_parts, int N, int C, int32_t* arg,
                    void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  MPA_CHECK_ARG(n_parts >= 0 && N > 0 && (C == 64 || C == 128 || C == 256), "pool_argmax: unsupported shape");
  if (n_parts == 0) return MPA_OK;
  MPA_CHECK_ARG(z && scale && arg, "pool_argmax: null pointer");
  {
    ProfScope ps("pool_argmax", stream);
    pool_argmax_kernel<<<n_parts, PB_THREADS, 0, stream>>>((const __nv_bfloat16*)z, scale, N, C, arg);
  }
  MPA_LAUNCH_CHECK();
  return MPA_OK;
}

}  // extern "C"
