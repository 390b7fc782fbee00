// Fused reductions of the geometric loss terms (utils/loss.py:7-202 as assembled
// by BaseModel._calc_loss, models/modules/base_model.py:256-290): one launch for
// the per-part sums (incl. rot_points_l2_loss, which re-rotates the part twice
// in the reference) and one for the per-shape terms, instead of ~60 elementwise
// / reduction kernels.  Forward only (used when autograd is not recording).
#include "mpa_common.cuh"

namespace mpa {

// per part: sum of part-level dist1, dist2, shape-level dist1, dist2, and
// sum_i |R1 v_i - R2 v_i|^2.  out [B*P, 5].  One CTA per part.
__global__ void __launch_bounds__(256)
loss_part_sums_kernel(const float* __restrict__ pts, const float* __restrict__ q1,
                      const float* __restrict__ q2, const float* __restrict__ pd1,
                      const float* __restrict__ pd2, const float* __restrict__ sd1,
                      const float* __restrict__ sd2, int N, int want_l2, float* __restrict__ out) {
  __shared__ float red[5][8];
  const int part = blockIdx.x;
  float acc[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
  float a[4], b[4];
#pragma unroll
  for (int c = 0; c < 4; ++c) { a[c] = q1[4 * part + c]; b[c] = q2[4 * part + c]; }
  for (int i = threadIdx.x; i < N; i += blockDim.x) {
    const long long o = (long long)part * N + i;
    acc[0] += pd1[o]; acc[1] += pd2[o]; acc[2] += sd1[o]; acc[3] += sd2[o];
    if (want_l2) {
      const float3 v = make_float3(pts[3 * o], pts[3 * o + 1], pts[3 * o + 2]);
      const float3 r1 = se3_apply(a, nullptr, v), r2 = se3_apply(b, nullptr, v);
      const float dx = r1.x - r2.x, dy = r1.y - r2.y, dz = r1.z - r2.z;
      acc[4] += dx * dx + dy * dy + dz * dz;
    }
  }
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < 5; ++k) {
    float v = acc[k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    if (lane == 0) red[k][wid] = v;
  }
  __syncthreads();
  if (threadIdx.x < 5) {
    float v = 0.f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) v += red[threadIdx.x][w];
    out[(long long)part * 5 + threadIdx.x] = v;
  }
}

// per shape: the five loss terms and their weighted total.  One warp per shape.
// terms [6, B]: trans, rot_pt_cd, transform_pt_cd, rot (cosine), rot_pt_l2, total
__global__ void loss_shape_terms_kernel(const float* __restrict__ sums, const float* __restrict__ q1,
                                        const float* __restrict__ t1, const float* __restrict__ q2,
                                        const float* __restrict__ t2, const float* __restrict__ valids,
                                        int B, int P, int N, int training,
                                        const float* __restrict__ weights,  // device [5]
                                        float* __restrict__ terms) {
  const int b = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (b >= B) return;
  float nv = 0.f, tr = 0.f, cd = 0.f, cs = 0.f, l2 = 0.f, sh_train = 0.f, sh_eval = 0.f;
  for (int p = lane; p < P; p += 32) {
    const int g = b * P + p;
    const float v = valids[g];
    const float* s = sums + (long long)g * 5;
    float d = 0.f, dot = 0.f;
#pragma unroll
    for (int c = 0; c < 3; ++c) { const float e = t1[3 * g + c] - t2[3 * g + c]; d += e * e; }
#pragma unroll
    for (int c = 0; c < 4; ++c) dot += q1[4 * g + c] * q2[4 * g + c];
    nv += v;
    tr += d * v;                                   // trans_l2_loss (loss.py:22-35)
    cs += (1.f - fabsf(dot)) * v;                  // rot_cosine_loss (:59-86)
    cd += (s[0] / N + s[1] / N) * v;               // rot_points_cd_loss (:131-134)
    l2 += (s[4] / N) * v;                          // rot_points_l2_loss (:105-106)
    sh_train += s[2] + s[3];                       // shape_cd_loss training (:185-193): padded dist is 0
    sh_eval += ((s[2] + s[3]) / N) * v;            // shape_cd_loss eval (:195-198)
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    nv += __shfl_xor_sync(0xffffffffu, nv, o); tr += __shfl_xor_sync(0xffffffffu, tr, o);
    cs += __shfl_xor_sync(0xffffffffu, cs, o); cd += __shfl_xor_sync(0xffffffffu, cd, o);
    l2 += __shfl_xor_sync(0xffffffffu, l2, o); sh_train += __shfl_xor_sync(0xffffffffu, sh_train, o);
    sh_eval += __shfl_xor_sync(0xffffffffu, sh_eval, o);
  }
  if (lane == 0) {
    const float L_trans = tr / nv, L_cd = cd / nv, L_rot = cs / nv, L_l2 = l2 / nv;
    const float L_shape = training ? sh_train / (float)(P * N) : sh_eval / nv;
    terms[0 * B + b] = L_trans;
    terms[1 * B + b] = L_cd;
    terms[2 * B + b] = L_shape;
    terms[3 * B + b] = L_rot;
    terms[4 * B + b] = L_l2;
    terms[5 * B + b] = weights[0] * L_trans + weights[1] * L_cd + weights[2] * L_shape +
                       weights[3] * L_rot + weights[4] * L_l2;
  }
}

// rot = normalize(W_r f + b_r), trans = W_t f + b_t for every token; one warp per token
__global__ void pose_outputs_kernel(const float* __restrict__ f, int T, int K,
                                    const float* __restrict__ wr, const float* __restrict__ br,
                                    const float* __restrict__ wt, const float* __restrict__ bt,
                                    int normalize, float* __restrict__ rot, float* __restrict__ trans) {
  const int tok = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (tok >= T) return;
  float acc[7] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  for (int k = lane; k < K; k += 32) {
    const float x = f[(long long)tok * K + k];
#pragma unroll
    for (int o = 0; o < 4; ++o) acc[o] = fmaf(x, wr[o * K + k], acc[o]);
#pragma unroll
    for (int o = 0; o < 3; ++o) acc[4 + o] = fmaf(x, wt[o * K + k], acc[4 + o]);
  }
#pragma unroll
  for (int o = 0; o < 7; ++o)
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) acc[o] += __shfl_xor_sync(0xffffffffu, acc[o], s);
  if (lane == 0) {
    float q[4];
#pragma unroll
    for (int o = 0; o < 4; ++o) q[o] = acc[o] + br[o];
    if (normalize) {  // F.normalize(p=2, dim=-1, eps=1e-12)
      const float n = fmaxf(sqrtf(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]), 1e-12f);
#pragma unroll
      for (int o = 0; o < 4; ++o) q[o] /= n;
    }
#pragma unroll
    for (int o = 0; o < 4; ++o) rot[(long long)tok * 4 + o] = q[o];
#pragma unroll
    for (int o = 0; o < 3; ++o) trans[(long long)tok * 3 + o] = acc[4 + o] + bt[o];
  }
}


// ---- whole pose head in one launch ------------------------------------------
// PoseRegressor.forward (models/modules/regressor.py:58-68): fc(K0->H1) LeakyReLU
// fc(H1->H2) LeakyReLU, rot_head(H2->4) (+ L2 normalisation), trans_head(H2->3), fp32
// on the CUDA cores (0.1 MMAC per token: the cost is reading the weights, once per CTA
// of PH_TOK tokens, with coalesced row reads).  A warp owns an output channel: lanes
// split k, the per-token partial sums are folded with a transposing butterfly so that
// lane 4t ends up with token t's total.
constexpr int PH_TOK = 8;
constexpr int PH_THREADS = 384;
constexpr int PH_OUTS = 4;  // output channels per warp iteration: 32 weight loads in flight per lane (8: measured slower)

// per-lane partial sums of 8 tokens -> token (lane >> 2)'s total in every lane of its group of 4
__device__ __forceinline__ float fold8(const float (&a)[8], int lane) {
  float b[4], c[2];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const bool hi = lane & 16;
    const float recv = __shfl_xor_sync(0xffffffffu, hi ? a[i] : a[i + 4], 16);
    b[i] = (hi ? a[i + 4] : a[i]) + recv;
  }
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const bool hi = lane & 8;
    const float recv = __shfl_xor_sync(0xffffffffu, hi ? b[i] : b[i + 2], 8);
    c[i] = (hi ? b[i + 2] : b[i]) + recv;
  }
  const bool hi = lane & 4;
  float d = (hi ? c[1] : c[0]) + __shfl_xor_sync(0xffffffffu, hi ? c[0] : c[1], 4);
  d += __shfl_xor_sync(0xffffffffu, d, 2);
  d += __shfl_xor_sync(0xffffffffu, d, 1);
  return d;
}

// hout[t][j] = act(b[j] + sum_k W[j][k] xin[t][k]) for the CTA's PH_TOK tokens (all in shared memory)
__device__ void ph_layer(const float* __restrict__ xin, int K, const float* __restrict__ W,
                         const float* __restrict__ b, int J, int act, float* __restrict__ hout) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  // every CTA streams the same weight matrix: start each one at a different row so that
  // they do not all hit the same L2 slice at the same moment
  const int rot = (int)((blockIdx.x * 37u) % (unsigned)J);
  for (int kc = 0; kc < K; kc += 256) {
    float xr[PH_TOK][8];
#pragma unroll
    for (int t = 0; t < PH_TOK; ++t)
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int k = kc + lane + 32 * i;
        xr[t][i] = k < K ? xin[t * K + k] : 0.f;
      }
    for (int j0 = warp; j0 < J; j0 += PH_OUTS * nwarps) {
      float w[PH_OUTS][8];
#pragma unroll
      for (int o = 0; o < PH_OUTS; ++o) {
        const int jl = j0 + o * nwarps;
        const int j = jl + rot < J ? jl + rot : jl + rot - J;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int k = kc + lane + 32 * i;
          w[o][i] = (k < K && jl < J) ? __ldg(W + (long long)j * K + k) : 0.f;
        }
      }
      float d[PH_OUTS];
#pragma unroll
      for (int o = 0; o < PH_OUTS; ++o) {
        float a[PH_TOK];
#pragma unroll
        for (int t = 0; t < PH_TOK; ++t) {
          a[t] = 0.f;
#pragma unroll
          for (int i = 0; i < 8; ++i) a[t] = fmaf(w[o][i], xr[t][i], a[t]);
        }
        d[o] = fold8(a, lane);
      }
      if ((lane & 3) == 0) {
#pragma unroll
        for (int o = 0; o < PH_OUTS; ++o) {
          const int jl = j0 + o * nwarps;
          const int j = jl + rot < J ? jl + rot : jl + rot - J;
          if (jl < J) {
            float* dst = hout + (lane >> 2) * J + j;
            *dst = (kc == 0) ? d[o] + __ldg(b + j) : *dst + d[o];
          }
        }
      }
    }
  }
  __syncthreads();
  if (act == 2) {
    for (int i = threadIdx.x; i < PH_TOK * J; i += blockDim.x) {
      const float v = hout[i];
      hout[i] = v > 0.f ? v : 0.2f * v;
    }
    __syncthreads();
  }
}

__global__ void __launch_bounds__(PH_THREADS)
pose_head_kernel(const float* __restrict__ x, int T, int K0, const float* __restrict__ w0,
                 const float* __restrict__ b0, int H1, const float* __restrict__ w1,
                 const float* __restrict__ b1, int H2, const float* __restrict__ wr,
                 const float* __restrict__ br, const float* __restrict__ wt,
                 const float* __restrict__ bt, int normalize, float* __restrict__ rot,
                 float* __restrict__ trans) {
  extern __shared__ float ph_smem[];
  float* xs = ph_smem;                 // [PH_TOK][K0]
  float* h1 = xs + PH_TOK * K0;        // [PH_TOK][H1]
  float* h2 = h1 + PH_TOK * H1;        // [PH_TOK][H2]
  float* hq = h2 + PH_TOK * H2;        // [PH_TOK][4]
  float* ht = hq + PH_TOK * 4;         // [PH_TOK][3]
  const int tok0 = blockIdx.x * PH_TOK;
  for (int i = threadIdx.x; i < PH_TOK * K0; i += blockDim.x) {
    const int t = tok0 + i / K0;
    xs[i] = t < T ? x[(long long)t * K0 + i % K0] : 0.f;
  }
  __syncthreads();
  ph_layer(xs, K0, w0, b0, H1, 2, h1);
  ph_layer(h1, H1, w1, b1, H2, 2, h2);
  ph_layer(h2, H2, wr, br, 4, 0, hq);
  ph_layer(h2, H2, wt, bt, 3, 0, ht);
  if (threadIdx.x < PH_TOK && tok0 + threadIdx.x < T) {
    const int t = threadIdx.x;
    float q[4] = {hq[t * 4], hq[t * 4 + 1], hq[t * 4 + 2], hq[t * 4 + 3]};
    if (normalize) {  // F.normalize(p=2, dim=-1, eps=1e-12)
      const float n = fmaxf(sqrtf(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]), 1e-12f);
#pragma unroll
      for (int o = 0; o < 4; ++o) q[o] /= n;
    }
    const long long tok = tok0 + t;
#pragma unroll
    for (int o = 0; o < 4; ++o) rot[tok * 4 + o] = q[o];
#pragma unroll
    for (int o = 0; o < 3; ++o) trans[tok * 3 + o] = ht[t * 3 + o];
  }
}


// ---- batched linear sum assignment (Hungarian matching of equivalent parts) ----
// BaseModel._linear_sum_assignment (models/modules/base_model.py:150-179) hands the p x p
// Chamfer cost matrix of a group of geometrically equivalent parts to SciPy's
// linear_sum_assignment.  SciPy (scipy/optimize/rectangular_lsap, not part of the
// reference tree; the reference pins no version) implements the shortest augmenting path
// algorithm of D. F. Crouse, "On implementing 2D rectangular assignment algorithms", IEEE
// TAES 52(4), 2016; this kernel restates that algorithm step for step -- float64 duals,
// columns visited in descending index order, ties towards an unassigned column -- so the
// assignment is the same one, without the device-to-host copy per batch.  One thread per
// problem (p <= 64; the data sets have p <= 20).
constexpr int LSAP_MAX = 64;

__global__ void lsap_kernel(const float* __restrict__ costs, const int* __restrict__ cost_off,
                            const int* __restrict__ sizes, const int* __restrict__ out_off, int G,
                            int* __restrict__ col4row_out) {
  const int gidx = blockIdx.x * blockDim.x + threadIdx.x;
  if (gidx >= G) return;
  const int n = sizes[gidx];
  const float* __restrict__ cost = costs + cost_off[gidx];
  int* out = col4row_out + out_off[gidx];
  double u[LSAP_MAX], v[LSAP_MAX], spc[LSAP_MAX];
  int path[LSAP_MAX], col4row[LSAP_MAX], row4col[LSAP_MAX], remaining[LSAP_MAX];
  bool SR[LSAP_MAX], SC[LSAP_MAX];
  const double inf = __longlong_as_double(0x7ff0000000000000ll);
  for (int i = 0; i < n; ++i) { u[i] = 0.0; v[i] = 0.0; path[i] = -1; col4row[i] = -1; row4col[i] = -1; }
  for (int cur = 0; cur < n; ++cur) {
    // ---- shortest augmenting path from row `cur` ----
    double minVal = 0.0;
    int num_remaining = n;
    for (int it = 0; it < n; ++it) { remaining[it] = n - it - 1; SR[it] = false; SC[it] = false; spc[it] = inf; }
    int sink = -1, i = cur;
    while (sink == -1) {
      int index = -1;
      double lowest = inf;
      SR[i] = true;
      for (int it = 0; it < num_remaining; ++it) {
        const int j = remaining[it];
        const double r = minVal + (double)cost[i * n + j] - u[i] - v[j];
        if (r < spc[j]) { path[j] = i; spc[j] = r; }
        if (spc[j] < lowest || (spc[j] == lowest && row4col[j] == -1)) { lowest = spc[j]; index = it; }
      }
      minVal = lowest;
      if (index < 0 || minVal == inf) break;  // infeasible (non-finite costs): leave the identity
      const int j = remaining[index];
      if (row4col[j] == -1) sink = j; else i = row4col[j];
      SC[j] = true;
      remaining[index] = remaining[--num_remaining];
    }
    if (sink < 0) {
      for (int r = 0; r < n; ++r) out[r] = r;
      return;
    }
    // ---- dual update ----
    u[cur] += minVal;
    for (int r = 0; r < n; ++r)
      if (SR[r] && r != cur) u[r] += minVal - spc[col4row[r]];
    for (int j = 0; j < n; ++j)
      if (SC[j]) v[j] -= minVal - spc[j];
    // ---- augment ----
    int j = sink;
    while (true) {
      const int r = path[j];
      row4col[j] = r;
      const int t = col4row[r];
      col4row[r] = j;
      j = t;
      if (r == cur) break;
    }
  }
  for (int r = 0; r < n; ++r) out[r] = col4row[r];
}


// ---- part matching of the semantic models (reference base_model.py:150-238) ----------------
// Cost entry (i, j) of one group of geometrically equivalent parts: Chamfer distance between
// the group's i-th part under its PREDICTED pose and its j-th part under its GROUND-TRUTH pose,
// on the n points the group's random subsample picked (mean of the nearest squared distances
// of both directions, :170-174).  One CTA per entry: both clouds (n <= MATCH_MAX_N points) are
// transformed into shared memory, thread t owns query t of either direction.
constexpr int MATCH_MAX_N = 128;
constexpr int MATCH_MAXP = 32;   // parts per group
__device__ __forceinline__ float block_sum_128(float v, float* s_red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = v;
  __syncthreads();
  const float t = (s_red[0] + s_red[1]) + (s_red[2] + s_red[3]);
  __syncthreads();
  return t;
}
__global__ void __launch_bounds__(MATCH_MAX_N)
match_cost_kernel(const float* __restrict__ pts, const float* __restrict__ q1, const float* __restrict__ t1,
                  const float* __restrict__ q2, const float* __restrict__ t2, int P, int N, int n,
                  const int* __restrict__ g_shape, const int* __restrict__ g_size,
                  const int* __restrict__ g_cost_off, const int* __restrict__ g_members,
                  const int* __restrict__ g_sample, const int* __restrict__ pairs,
                  float* __restrict__ costs) {
  __shared__ float ax[MATCH_MAX_N], ay[MATCH_MAX_N], az[MATCH_MAX_N];
  __shared__ float bx[MATCH_MAX_N], by[MATCH_MAX_N], bz[MATCH_MAX_N];
  __shared__ float s_red[4];
  const int pr = pairs[blockIdx.x];
  const int g = pr >> 16, i = (pr >> 8) & 0xff, j = pr & 0xff;
  const int b = g_shape[g], p = g_size[g];
  const int part_i = g_members[g * MATCH_MAXP + i], part_j = g_members[g * MATCH_MAXP + j];
  const int t = threadIdx.x;
  if (t < n) {
    const int k = g_sample[g * n + t];
    const float* pi = pts + ((long long)(b * P + part_i) * N + k) * 3;
    const float* pj = pts + ((long long)(b * P + part_j) * N + k) * 3;
    const float* qa = q1 + (long long)(b * P + part_i) * 4;
    const float* qb = q2 + (long long)(b * P + part_j) * 4;
    const float qA[4] = {qa[0], qa[1], qa[2], qa[3]}, qB[4] = {qb[0], qb[1], qb[2], qb[3]};
    const float3 a = se3_apply(qA, t1 + (long long)(b * P + part_i) * 3, make_float3(pi[0], pi[1], pi[2]));
    const float3 c = se3_apply(qB, t2 + (long long)(b * P + part_j) * 3, make_float3(pj[0], pj[1], pj[2]));
    ax[t] = a.x; ay[t] = a.y; az[t] = a.z;
    bx[t] = c.x; by[t] = c.y; bz[t] = c.z;
  }
  __syncthreads();
  float d1 = 0.f, d2 = 0.f;
  if (t < n) {
    float m1 = 1e32f, m2 = 1e32f;  // chamfer_kernel.cu:60
    const float x1 = ax[t], y1 = ay[t], z1 = az[t], x2 = bx[t], y2 = by[t], z2 = bz[t];
    for (int k = 0; k < n; ++k) {
      m1 = fminf(m1, sqdist_ref(x1, y1, z1, bx[k], by[k], bz[k]));
      m2 = fminf(m2, sqdist_ref(x2, y2, z2, ax[k], ay[k], az[k]));
    }
    d1 = m1; d2 = m2;
  }
  const float s1 = block_sum_128(d1, s_red), s2 = block_sum_128(d2, s_red);
  if (t == 0) costs[g_cost_off[g] + i * p + j] = s1 / (float)n + s2 / (float)n;
}

// new_gt[b, members[r]] = gt[b, members[col[r]]] for every group (reference :229-233)
__global__ void match_permute_kernel(const float* __restrict__ gt_trans, const float* __restrict__ gt_quat, int P,
                                     const int* __restrict__ g_shape, const int* __restrict__ g_size,
                                     const int* __restrict__ g_out_off, const int* __restrict__ g_members,
                                     const int* __restrict__ col_of_row, int G, float* __restrict__ new_trans,
                                     float* __restrict__ new_quat) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  const int g = e / MATCH_MAXP, r = e % MATCH_MAXP;
  if (g >= G || r >= g_size[g]) return;
  const int b = g_shape[g];
  const int dst = b * P + g_members[g * MATCH_MAXP + r];
  const int src = b * P + g_members[g * MATCH_MAXP + col_of_row[g_out_off[g] + r]];
#pragma unroll
  for (int k = 0; k < 3; ++k) new_trans[dst * 3 + k] = gt_trans[src * 3 + k];
#pragma unroll
  for (int k = 0; k < 4; ++k) new_quat[dst * 4 + k] = gt_quat[src * 4 + k];
}

}  // namespace mpa

using namespace mpa;

extern "C" {

int mpa_geometric_losses(const float* pts, const float* quat1, const float* trans1, const float* quat2,
                         const float* trans2, const float* valids, const float* part_dist1,
                         const float* part_dist2, const float* shape_dist1, const float* shape_dist2,
                         int B, int P, int N, int training, int want_rot_l2, const float* weights,
                         float* terms, void* ws, size_t ws_bytes, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  MPA_CHECK_ARG(B >= 0 && P > 0 && N > 0, "geometric_losses: bad sizes");
  if (B == 0) return MPA_OK;
  MPA_CHECK_ARG(pts && quat1 && trans1 && quat2 && trans2 && valids && part_dist1 && part_dist2 &&
                    shape_dist1 && shape_dist2 && weights && terms,
                "geometric_losses: null pointer");
  Scratch scratch;
  int rc = scratch.acquire(ws, ws_bytes, sizeof(float) * 5 * (size_t)B * P, stream);
  if (rc != MPA_OK) return rc;
  float* sums = (float*)scratch.base;
  {
    ProfScope ps("loss_part_sums", stream);
    loss_part_sums_kernel<<<B * P, 256, 0, stream>>>(pts, quat1, quat2, part_dist1, part_dist2,
                                                     shape_dist1, shape_dist2, N, want_rot_l2, sums);
  }
  MPA_LAUNCH_CHECK();
  {
    ProfScope ps("loss_shape_terms", stream);
    loss_shape_terms_kernel<<<(B * 32 + 127) / 128, 128, 0, stream>>>(
        sums, quat1, trans1, quat2, trans2, valids, B, P, N, training, weights, terms);
  }
  MPA_LAUNCH_CHECK();
  return MPA_OK;
}

size_t mpa_geometric_losses_workspace_bytes(int B, int P) { return sizeof(float) * 5 * (size_t)B * P; }

int mpa_pose_outputs(const float* feats, int T, int K, const float* rot_w, const float* rot_b,
                     const float* trans_w, const float* trans_b, int normalize, float* rot,
                     float* trans, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  MPA_CHECK_ARG(T >= 0 && K > 0, "pose_outputs: bad sizes");
  if (T == 0) return MPA_OK;
  MPA_CHECK_ARG(feats && rot_w && rot_b && trans_w && trans_b && rot && trans, "pose_outputs: null pointer");
  {
    ProfScope ps("pose_outputs", stream);
    pose_outputs_kernel<<<(T * 32 + 255) / 256, 256, 0, stream>>>(feats, T, K, rot_w, rot_b, trans_w,
                                                                  trans_b, normalize, rot, trans);
  }
  MPA_LAUNCH_CHECK();
  return MPA_OK;
}

int mpa_pose_head_forward(const float* feats, int T, int K0, const float* fc0_w, const float* fc0_b,
                          int H1, const float* fc1_w, const float* fc1_b, int H2, const float* rot_w,
                          const float* rot_b, const float* trans_w, const float* trans_b, int normalize,
                          float* rot, float* trans, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  MPA_CHECK_ARG(T >= 0 && K0 > 0 && H1 > 0 && H2 > 0, "pose_head_forward: bad sizes");
  if (T == 0) return MPA_OK;
  MPA_CHECK_ARG(feats && fc0_w && fc0_b && fc1_w && fc1_b && rot_w && rot_b && trans_w && trans_b && rot &&
                    trans, "pose_head_forward: null pointer");
  const size_t smem = sizeof(float) * PH_TOK * ((size_t)K0 + H1 + H2 + 7);
  MPA_CHECK_ARG(smem <= 200 * 1024, "pose_head_forward: layer widths too large (%d, %d, %d)", K0, H1, H2);
  static DeviceOnce attr;  // the limit is per device: raise it to the checked maximum once on each
  if (smem > 48 * 1024 && attr.pending()) {
    MPA_CUDA(cudaFuncSetAttribute(pose_head_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr.done();
  }
  {
    ProfScope ps("pose_head", stream);
    pose_head_kernel<<<(T + PH_TOK - 1) / PH_TOK, PH_THREADS, smem, stream>>>(
        feats, T, K0, fc0_w, fc0_b, H1, fc1_w, fc1_b, H2, rot_w, rot_b, trans_w, trans_b, normalize, rot,
        trans);
  }
  MPA_LAUNCH_CHECK();
  return MPA_OK;
}

int mpa_lsap_batched(const float* costs, const int32_t* cost_offsets, const int32_t* sizes,
                     const int32_t* out_offsets, int n_problems, int max_size, int32_t* col_of_row,
                     void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  MPA_CHECK_ARG(n_problems >= 0, "lsap_batched: negative problem count");
  MPA_CHECK_ARG(max_size >= 0 && max_size <= LSAP_MAX, "lsap_batched: at most %d rows per problem (got %d)",
                LSAP_MAX, max_size);
  if (n_problems == 0) return MPA_OK;
  MPA_CHECK_ARG(costs && cost_offsets && sizes && out_offsets && col_of_row, "lsap_batched: null pointer");
  {
    ProfScope ps("lsap", stream);
    lsap_kernel<<<(n_problems + 31) / 32, 32, 0, stream>>>(costs, cost_offsets, sizes, out_offsets, n_problems,
                                                           col_of_row);
  }
  MPA_LAUNCH_CHECK();
  return MPA_OK;
}


/* Matching of all groups of equivalent parts of a batch in three launches: cost matrices,
 * assignments (lsap_kernel), permuted ground-truth poses.  `table` (device int32) holds, for G
 * groups: shape index [G], size p_g [G], cost offset [G], output offset [G], members
 * [G, 32] (part indices), subsample [G, n] (point indices) and the (g << 16 | i << 8 | j)
 * entry list [n_pairs]; new_trans / new_quat must already hold copies of the ground truth. */
size_t mpa_match_parts_workspace_bytes(int n_pairs, int total_rows) {
  return align_up(sizeof(float) * (size_t)(n_pairs > 0 ? n_pairs : 1), 256) +
         align_up(sizeof(int) * (size_t)(total_rows > 0 ? total_rows : 1), 256);
}

int mpa_match_parts(const float* pts, const float* pred_quat, const float* pred_trans, const float* gt_quat,
                    const float* gt_trans, int B, int P, int N, int n, const int32_t* table, int G, int n_pairs,
                    int total_rows, int max_size, float* new_trans, float* new_quat, float* costs_out,
                    int32_t* col_out, void* ws, size_t ws_bytes, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  MPA_CHECK_ARG(B >= 0 && P > 0 && N > 0 && n > 0 && n <= MATCH_MAX_N && n <= N, "match_parts: bad sizes");
  MPA_CHECK_ARG(G >= 0 && max_size <= MATCH_MAXP && max_size <= LSAP_MAX, "match_parts: groups of at most %d parts",
                MATCH_MAXP);
  if (G == 0) return MPA_OK;
  MPA_CHECK_ARG(pts && pred_quat && pred_trans && gt_quat && gt_trans && table && new_trans && new_quat,
                "match_parts: null pointer");
  Scratch scratch;
  int rc = scratch.acquire(ws, ws_bytes, mpa_match_parts_workspace_bytes(n_pairs, total_rows), stream);
  if (rc != MPA_OK) return rc;
  float* costs = costs_out != nullptr ? costs_out : (float*)scratch.base;
  int* col = col_out != nullptr ? col_out
                                : (int*)((char*)scratch.base + align_up(sizeof(float) * (size_t)n_pairs, 256));
  const int* g_shape = table;
  const int* g_size = table + G;
  const int* g_cost_off = table + 2 * G;
  const int* g_out_off = table + 3 * G;
  const int* g_members = table + 4 * G;
  const int* g_sample = g_members + (size_t)G * MATCH_MAXP;
  const int* pairs = g_sample + (size_t)G * n;
  {
    ProfScope ps("match_cost", stream);
    match_cost_kernel<<<n_pairs, MATCH_MAX_N, 0, stream>>>(pts, pred_quat, pred_trans, gt_quat, gt_trans, P, N, n,
                                                          g_shape, g_size, g_cost_off, g_members, g_sample, pairs,
                                                          costs);
  }
  MPA_LAUNCH_CHECK();
  {
    ProfScope ps("lsap", stream);
    lsap_kernel<<<(G + 31) / 32, 32, 0, stream>>>(costs, g_cost_off, g_size, g_out_off, G, col);
  }
  MPA_LAUNCH_CHECK();
  {
    ProfScope ps("match_permute", stream);
    match_permute_kernel<<<(G * MATCH_MAXP + 127) / 128, 128, 0, stream>>>(gt_trans, gt_quat, P, g_shape, g_size,
                                                                         g_out_off, g_members, col, G, new_trans,
                                                                         new_quat);
  }
  MPA_LAUNCH_CHECK();
  return MPA_OK;
}

}  // extern "C"
