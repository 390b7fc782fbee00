/*
 * mpa_oracle.c -- TEST INFRASTRUCTURE ONLY (never imported by the product path).
 *
 * Scalar CPU restatement of the native / third-party arithmetic on the
 * multi_part_assembly hot path.  Every function cites the reference lines it
 * follows (paths relative to /root/reference/multi_part_assembly).  Built by
 * oracle/Makefile into oracle/libmpa_oracle.so, loaded by oracle/cpu.py.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may call this.
 *
 * Compile with -ffp-contract=off: every fused multiply-add below is explicit
 * (fmaf) so the rounding sequence is the one written, on any host compiler.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* squared distance exactly as nvcc contracts the reference expression
 *   (x1-x2)*(x1-x2) + (y1-y2)*(y1-y2) + (z1-z2)*(z1-z2)
 * (utils/chamfer/cuda/chamfer_kernel.cu:80) with default -fmad=true:
 *   t = dy*dy ; t = fma(dx,dx,t) ; t = fma(dz,dz,t)
 * (checked with `nvcc -ptx` on that expression for sm_100a). */
static inline float sqdist_ref(float x1, float y1, float z1,
                               float x2, float y2, float z2) {
  float dx = x1 - x2, dy = y1 - y2, dz = z1 - z2;
  float t = dy * dy;
  t = fmaf(dx, dx, t);
  t = fmaf(dz, dz, t);
  return t;
}

/* same distance without contraction: the definition in
 * utils/chamfer/test_chamfer.py:8-18 (diff**2 summed over the last axis). */
static inline float sqdist_unfused(float x1, float y1, float z1,
                                   float x2, float y2, float z2) {
  float dx = x1 - x2, dy = y1 - y2, dz = z1 - z2;
  float a = dx * dx, b = dy * dy, c = dz * dz;
  return (a + b) + c;
}

/* One direction of ChamferForwardKernel (chamfer_kernel.cu:32-95):
 * for every point of xyz1[b] the squared distance to, and the index of, its
 * nearest point of xyz2[b]; init min_dist=1e32, min_idx=-1 (:60-61); strict
 * `d < min_dist` with j ascending so ties keep the lowest index (:81-85). */
static void nn_one_direction(const float* xyz1, const float* xyz2, int64_t B,
                             int64_t n1, int64_t n2, float* dist, int64_t* idx,
                             int fused) {
#pragma omp parallel for schedule(static)
  for (int64_t bi = 0; bi < B * n1; ++bi) {
    int64_t b = bi / n1;
    const float* p = xyz1 + bi * 3;
    const float* q = xyz2 + b * n2 * 3;
    float best = 1e32f;
    int64_t arg = -1;
    for (int64_t j = 0; j < n2; ++j) {
      float d = fused ? sqdist_ref(p[0], p[1], p[2], q[3 * j], q[3 * j + 1], q[3 * j + 2])
                      : sqdist_unfused(p[0], p[1], p[2], q[3 * j], q[3 * j + 1], q[3 * j + 2]);
      if (d < best) { best = d; arg = j; }
    }
    dist[bi] = best;
    idx[bi] = arg;
  }
}

/* ChamferForward (chamfer_kernel.cu:116-168): both directions. */
void oracle_chamfer_forward(const float* xyz1, const float* xyz2, int64_t B,
                            int64_t n1, int64_t n2, float* dist1, int64_t* idx1,
                            float* dist2, int64_t* idx2, int fused) {
  nn_one_direction(xyz1, xyz2, B, n1, n2, dist1, idx1, fused);
  nn_one_direction(xyz2, xyz1, B, n2, n1, dist2, idx2, fused);
}

/* ChamferBackwardKernel (chamfer_kernel.cu:175-210), one direction:
 * g = 2*grad_dist[i]; grad_a[i] += g*(a_i - b_idx); grad_b[idx] -= same.
 * The reference accumulates with float atomics (order unspecified); the
 * oracle accumulates in double in ascending i and rounds once, which is the
 * value every float summation order approximates. */
static void bwd_one_direction(const float* grad_dist, const int64_t* index,
                              const float* a, const float* b, double* ga,
                              double* gb, int64_t B, int64_t n1, int64_t n2) {
  for (int64_t i = 0; i < B * n1; ++i) {
    int64_t bt = i / n1;
    int64_t j = bt * n2 + index[i];
    float g = grad_dist[i] * 2.0f;
    for (int c = 0; c < 3; ++c) {
      float gc = g * (a[3 * i + c] - b[3 * j + c]);
      ga[3 * i + c] += (double)gc;
      gb[3 * j + c] -= (double)gc;
    }
  }
}

/* ChamferBackward (chamfer_kernel.cu:224-289). */
void oracle_chamfer_backward(const float* grad_dist1, const float* grad_dist2,
                             const float* xyz1, const float* xyz2,
                             const int64_t* idx1, const int64_t* idx2, int64_t B,
                             int64_t n1, int64_t n2, float* grad_xyz1,
                             float* grad_xyz2) {
  double* g1 = (double*)calloc((size_t)(B * n1 * 3), sizeof(double));
  double* g2 = (double*)calloc((size_t)(B * n2 * 3), sizeof(double));
  bwd_one_direction(grad_dist1, idx1, xyz1, xyz2, g1, g2, B, n1, n2);
  bwd_one_direction(grad_dist2, idx2, xyz2, xyz1, g2, g1, B, n2, n1);
  for (int64_t i = 0; i < B * n1 * 3; ++i) grad_xyz1[i] = (float)g1[i];
  for (int64_t i = 0; i < B * n2 * 3; ++i) grad_xyz2[i] = (float)g2[i];
  free(g1);
  free(g2);
}

/* Hamilton product, the published pytorch3d.transforms.quaternion_raw_multiply
 * (third-party, un-vendored and unpinned: setup.py:3-6 lists bare
 * 'pytorch3d'); real part first.  Each output is evaluated left to right as
 * the four-term expression of the published source, no contraction (torch
 * elementwise kernels do not fuse). */
static inline void qmul_raw(const float a[4], const float b[4], float o[4]) {
  o[0] = ((a[0] * b[0] - a[1] * b[1]) - a[2] * b[2]) - a[3] * b[3];
  o[1] = ((a[0] * b[1] + a[1] * b[0]) + a[2] * b[3]) - a[3] * b[2];
  o[2] = ((a[0] * b[2] - a[1] * b[3]) + a[2] * b[0]) + a[3] * b[1];
  o[3] = ((a[0] * b[3] + a[1] * b[2]) - a[2] * b[1]) + a[3] * b[0];
}

/* qrot / qtransform (utils/transforms.py:75-109) over a [n_parts, N, 3] cloud
 * with one quaternion (and optional translation) per part:
 *   v' = (q (x) (0,v) (x) conj(q))[1:] (+ t)
 * = pytorch3d quaternion_apply; conj only, no normalisation, so a non-unit q
 * scales by |q|^2 exactly as the reference does.  trans may be NULL (rot_pc). */
void oracle_se3_transform(const float* quat, const float* trans,
                          const float* pts, int64_t n_parts, int64_t N,
                          float* out) {
#pragma omp parallel for schedule(static)
  for (int64_t p = 0; p < n_parts; ++p) {
    const float* q = quat + 4 * p;
    float qc[4] = {q[0], -q[1], -q[2], -q[3]};
    for (int64_t i = 0; i < N; ++i) {
      const float* v = pts + (p * N + i) * 3;
      float pv[4] = {0.0f, v[0], v[1], v[2]};
      float t1[4], t2[4];
      qmul_raw(q, pv, t1);
      qmul_raw(t1, qc, t2);
      float* o = out + (p * N + i) * 3;
      for (int c = 0; c < 3; ++c)
        o[c] = trans ? t2[c + 1] + trans[3 * p + c] : t2[c + 1];
    }
  }
}

/* knn (models/modules/encoder/dgcnn.py:8-15) on x [n, C, N] (channel-major as
 * the reference passes it): score(i,j) = -|xi|^2 + 2 xi.xj - |xj|^2 in the
 * expanded form, top-k largest (self included).  The reference's matmul
 * accumulation order and topk tie order are unspecified, so the oracle fixes
 * them: sequential-c fp32 FMA dot product, selection by (score desc, index asc),
 * and the parity contract is the SORTED index set per row.
 * out_idx: [n, N, k] int64, ascending index order within a row. */
static int cmp_i64(const void* a, const void* b) {
  int64_t x = *(const int64_t*)a, y = *(const int64_t*)b;
  return (x > y) - (x < y);
}
void oracle_knn(const float* x, int64_t n, int64_t C, int64_t N, int64_t k,
                int64_t* out_idx) {
#pragma omp parallel
  {
    float* score = (float*)malloc((size_t)N * sizeof(float));
    float* xx = (float*)malloc((size_t)N * sizeof(float));
    char* taken = (char*)malloc((size_t)N);
#pragma omp for schedule(dynamic, 1)
    for (int64_t b = 0; b < n; ++b) {
      const float* xb = x + b * C * N;
      for (int64_t j = 0; j < N; ++j) {
        float s = 0.0f;
        for (int64_t c = 0; c < C; ++c) s = fmaf(xb[c * N + j], xb[c * N + j], s);
        xx[j] = s;
      }
      for (int64_t i = 0; i < N; ++i) {
        /* dgcnn.py:10-12: inner = -2 x^T x; xx is [B,1,N] so the FIRST term of
         * `-xx - inner - xx^T` is -|x_j|^2 and the transposed one is -|x_i|^2;
         * evaluated in the order written: (-xx_j - inner) - xx_i. */
        for (int64_t j = 0; j < N; ++j) {
          float dot = 0.0f;
          for (int64_t c = 0; c < C; ++c) dot = fmaf(xb[c * N + i], xb[c * N + j], dot);
          float inner = -2.0f * dot;
          score[j] = (-xx[j] - inner) - xx[i];
        }
        memset(taken, 0, (size_t)N);
        int64_t* row = out_idx + (b * N + i) * k;
        for (int64_t s = 0; s < k; ++s) {
          int64_t arg = -1;
          float best = -INFINITY;
          for (int64_t j = 0; j < N; ++j)
            if (!taken[j] && (arg < 0 || score[j] > best)) { best = score[j]; arg = j; }
          taken[arg] = 1;
          row[s] = arg;
        }
        qsort(row, (size_t)k, sizeof(int64_t), cmp_i64);
      }
    }
    free(score);
    free(xx);
    free(taken);
  }
}

/* ---- PointNet++ set abstraction (pointnet2_ops CUDA extension) ---------------------------- */
/* Squared distance as nvcc contracts the reference's expression (sampling_gpu.cu:103-104,
 * ball_query_gpu.cu:34-35): t = dy*dy; t = fma(dx,dx,t); t = fma(dz,dz,t). */
static float sqdist_fused(float x1, float y1, float z1, float x2, float y2, float z2) {
  const float dx = x1 - x2, dy = y1 - y2, dz = z1 - z2;
  float t = dy * dy;
  t = fmaf(dx, dx, t);
  t = fmaf(dz, dz, t);
  return t;
}

/* furthest_point_sampling_kernel (sampling_gpu.cu:74-177), restated thread by thread: `block`
 * threads (opt_n_threads(n), cuda_utils.h:15-19) each scan a strided slice keeping the first
 * strict maximum (:106-109), then the shared-memory tree keeps the LOWER slot on ties
 * (__update, :64-71).  temp starts at 1e10 (sampling.cpp:75); points with |p|^2 <= 1e-3 are
 * skipped (:100-101). */
void oracle_fps(const float* xyz, int64_t B, int64_t n, int64_t m, int64_t* idx) {
  int64_t block = 1;
  while (block * 2 <= n && block < 512) block *= 2;
#pragma omp parallel for schedule(dynamic)
  for (int64_t b = 0; b < B; ++b) {
    const float* p = xyz + b * n * 3;
    float* temp = (float*)malloc(sizeof(float) * (size_t)n);
    float* dists = (float*)malloc(sizeof(float) * (size_t)block);
    int64_t* dists_i = (int64_t*)malloc(sizeof(int64_t) * (size_t)block);
    for (int64_t k = 0; k < n; ++k) temp[k] = 1e10f;
    int64_t old = 0;
    idx[b * m] = 0;
    for (int64_t j = 1; j < m; ++j) {
      const float x1 = p[old * 3], y1 = p[old * 3 + 1], z1 = p[old * 3 + 2];
      for (int64_t tid = 0; tid < block; ++tid) {
        int64_t besti = 0;
        float best = -1.0f;
        for (int64_t k = tid; k < n; k += block) {
          const float x2 = p[k * 3], y2 = p[k * 3 + 1], z2 = p[k * 3 + 2];
          float mag = y2 * y2;
          mag = fmaf(x2, x2, mag);
          mag = fmaf(z2, z2, mag);
          if (mag <= 1e-3f) continue;
          const float d = sqdist_fused(x2, y2, z2, x1, y1, z1);
          const float d2 = d < temp[k] ? d : temp[k];
          temp[k] = d2;
          besti = d2 > best ? k : besti;
          best = d2 > best ? d2 : best;
        }
        dists[tid] = best;
        dists_i[tid] = besti;
      }
      for (int64_t s = block / 2; s >= 1; s /= 2)
        for (int64_t tid = 0; tid < s; ++tid) {
          const float v1 = dists[tid], v2 = dists[tid + s];
          const int64_t i1 = dists_i[tid], i2 = dists_i[tid + s];
          dists[tid] = v1 > v2 ? v1 : v2;
          dists_i[tid] = v2 > v1 ? i2 : i1;
        }
      old = dists_i[0];
      idx[b * m + j] = old;
    }
    free(temp);
    free(dists);
    free(dists_i);
  }
}

/* query_ball_point_kernel (ball_query_gpu.cu:13-48); idx is zero-initialised (ball_query.cpp). */
void oracle_ball_query(const float* xyz, const float* new_xyz, int64_t B, int64_t n, int64_t m, float radius,
                       int64_t nsample, int64_t* idx) {
  const float radius2 = radius * radius;
#pragma omp parallel for schedule(dynamic)
  for (int64_t b = 0; b < B; ++b) {
    const float* p = xyz + b * n * 3;
    for (int64_t j = 0; j < m; ++j) {
      const float* c = new_xyz + (b * m + j) * 3;
      int64_t* out = idx + (b * m + j) * nsample;
      for (int64_t l = 0; l < nsample; ++l) out[l] = 0;
      int64_t cnt = 0;
      for (int64_t k = 0; k < n && cnt < nsample; ++k) {
        const float d2 = sqdist_fused(c[0], c[1], c[2], p[k * 3], p[k * 3 + 1], p[k * 3 + 2]);
        if (d2 < radius2) {
          if (cnt == 0)
            for (int64_t l = 0; l < nsample; ++l) out[l] = k;
          out[cnt] = k;
          ++cnt;
        }
      }
    }
  }
}

int oracle_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
void oracle_set_num_threads(int n) {
#ifdef _OPENMP
  omp_set_num_threads(n);
#else
  (void)n;
#endif
}
