// k-nearest-neighbour graph and EdgeConv aggregation of the DGCNN part encoder.
//
// knn: replaces `knn` (models/modules/encoder/dgcnn.py:8-15): for every point
// the k highest scores  -|x_i|^2 + 2 x_i.x_j - |x_j|^2  (self included), in
// the reference's expanded form.  The reference materialises the [n, N, N]
// score tensor with a batched cuBLAS SGEMM and runs a radix-select topk over
// it; here a CTA owns 32 (or 16) query rows of one part, streams the
// candidates through a register-tiled fp32 contraction (sequential-k FMA, the
// oracle's order -> bit-exact scores), keeps the score rows in shared memory
// and extracts the top k there.  Nothing of size N x N touches HBM.
//
// edge_aggregate: replaces get_graph_feature + 1x1 Conv2d + max over k
// (dgcnn.py:18-38, 81-95) after the algebraic split
//     W [x_j - x_i ; x_i] = W1 x_j + (W2 - W1) x_i = u_j + v_i,
// so the [n, 2C, N, k] edge tensor (10+ GB at n=512, C=128) never exists: per
// point it gathers the k neighbours' u rows (coalesced along channels), and
// keeps max, min, sum and sum of squares of u_j + v_i -- what BatchNorm2d
// (batch statistics over n*N*k edges) + LeakyReLU + max-over-k need.
#include <algorithm>

#include "mpa_common.cuh"

namespace mpa {

constexpr int KNN_THREADS = 256;
constexpr int KNN_BN = 128;   // candidates per block
constexpr int KNN_KC = 32;    // channels per chunk

__global__ void sqnorm_kernel(const float* __restrict__ x, long long rows, int C, float* __restrict__ xx) {
  for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r < rows;
       r += (long long)gridDim.x * blockDim.x) {
    const float* p = x + r * C;
    float s = 0.f;
    for (int c = 0; c < C; ++c) s = __fmaf_rn(p[c], p[c], s);
    xx[r] = s;
  }
}

// x [n, N, C] row-major, xx [n, N]; idx [n, N, k] int32 (best first; ties -> lower index)
template <int ROWS>
__global__ void __launch_bounds__(KNN_THREADS)
knn_kernel(const float* __restrict__ x, const float* __restrict__ xx, const float* __restrict__ valids,
           int N, int C, int k, int* __restrict__ idx) {
  extern __shared__ float sm[];
  const int Cp = ((C + 3) & ~3) + 4;           // padded query row stride
  float* As = sm;                               // [ROWS][Cp]
  float* Bs = As + ROWS * Cp;                   // [KNN_BN][KNN_KC + 4]
  float* S = Bs + KNN_BN * (KNN_KC + 4);        // [ROWS][Np]
  const int Np = (N + KNN_BN - 1) / KNN_BN * KNN_BN;
  const int tiles = (N + ROWS - 1) / ROWS;
  const int part = blockIdx.x / tiles, i0 = (blockIdx.x % tiles) * ROWS;
  if (valids != nullptr && valids[part] == 0.0f) return;  // padded part: no graph (CTA-uniform)
  const float* xp = x + (long long)part * N * C;
  const float* xxp = xx + (long long)part * N;
  const int tid = threadIdx.x, tx = tid & 31, ty = tid >> 5;
  constexpr int RPT = ROWS / 8;                 // rows per thread (8 warps)
  const float ninf = -__int_as_float(0x7f800000);

  for (int e = tid; e < ROWS * Cp; e += KNN_THREADS) {
    const int r = e / Cp, c = e % Cp;
    As[e] = (i0 + r < N && c < C) ? xp[(long long)(i0 + r) * C + c] : 0.f;
  }
  for (int j0 = 0; j0 < N; j0 += KNN_BN) {
    float acc[RPT][4];
#pragma unroll
    for (int r = 0; r < RPT; ++r)
#pragma unroll
      for (int c = 0; c < 4; ++c) acc[r][c] = 0.f;
    for (int kc = 0; kc < C; kc += KNN_KC) {
      __syncthreads();
      for (int e = tid; e < KNN_BN * KNN_KC; e += KNN_THREADS) {
        const int col = e / KNN_KC, kk = e % KNN_KC;
        Bs[col * (KNN_KC + 4) + kk] =
            (j0 + col < N && kc + kk < C) ? xp[(long long)(j0 + col) * C + kc + kk] : 0.f;
      }
      __syncthreads();
      const int klen = min(KNN_KC, ((C - kc) + 3) & ~3);
      for (int kk = 0; kk < klen; kk += 4) {
        float4 a[RPT], b[4];
#pragma unroll
        for (int r = 0; r < RPT; ++r)
          a[r] = *reinterpret_cast<const float4*>(&As[(ty * RPT + r) * Cp + kc + kk]);
#pragma unroll
        for (int c = 0; c < 4; ++c)
          b[c] = *reinterpret_cast<const float4*>(&Bs[(tx + 32 * c) * (KNN_KC + 4) + kk]);
#pragma unroll
        for (int r = 0; r < RPT; ++r)
#pragma unroll
          for (int c = 0; c < 4; ++c) {  // ascending k, one accumulator: the oracle's order
            float d = acc[r][c];
            d = __fmaf_rn(a[r].x, b[c].x, d);
            d = __fmaf_rn(a[r].y, b[c].y, d);
            d = __fmaf_rn(a[r].z, b[c].z, d);
            d = __fmaf_rn(a[r].w, b[c].w, d);
            acc[r][c] = d;
          }
      }
    }
#pragma unroll
    for (int r = 0; r < RPT; ++r) {
      const int row = ty * RPT + r;
      const float xi = (i0 + row < N) ? xxp[i0 + row] : 0.f;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int j = j0 + tx + 32 * c;
        float s = ninf;
        if (j < N) {
          const float inner = __fmul_rn(-2.0f, acc[r][c]);          // dgcnn.py:10
          s = __fsub_rn(__fsub_rn(-xxp[j], inner), xi);             // dgcnn.py:12
        }
        S[row * Np + j] = s;
      }
    }
  }
  __syncthreads();
  // ---- top-k per row: each lane caches the best of its strided slice ----
  for (int row = ty; row < ROWS; row += 8) {
    if (i0 + row >= N) break;
    float* Sr = S + row * Np;
    float best = ninf;
    int barg = 0x7fffffff;
    for (int j = tx; j < N; j += 32)
      if (Sr[j] > best) { best = Sr[j]; barg = j; }
    int* out = idx + ((long long)part * N + i0 + row) * k;
    for (int s = 0; s < k; ++s) {
      float v = best;
      int a = barg;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, v, o);
        const int oa = __shfl_xor_sync(0xffffffffu, a, o);
        if (ov > v || (ov == v && oa < a)) { v = ov; a = oa; }
      }
      if (tx == 0) out[s] = a;
      if ((a & 31) == tx && a < N) {  // owner lane: remove and rescan its slice
        Sr[a] = ninf;
        best = ninf; barg = 0x7fffffff;
        for (int j = tx; j < N; j += 32)
          if (Sr[j] > best) { best = Sr[j]; barg = j; }
      }
    }
  }
}

// ---- register-tiled variant (k <= 32) ------------------------------------------
// A CTA takes 128 query rows of one part at a time (persistent over the work list);
// candidates stream through in blocks of 128.  The 128 x 128 score block is an 8 x 8
// register tile per thread (64 FMAs per 4 16-byte shared-memory loads: FMA-issue bound
// instead of LDS bound), accumulated over the channels in ascending order with one
// accumulator per pair -- the oracle's order, so the scores are bit-identical to
// knn_kernel's.  The next operand chunk is prefetched into registers while the current
// one is multiplied.  Finished score rows go to a per-CTA scratch slab (128 x Np fp32,
// 512 KB: it lives in the 126 MB L2, never in HBM for long) and are selected from once
// the row is complete:
//   each lane takes the maximum of its 32-strided slice; the k-th largest of the 32
//   lane maxima (one 32-key bitonic sort by shuffles) is a lower bound T0 of the k-th
//   best score, typically passed by only ~k..2k candidates; those are compacted (ballot-
//   free: per-lane counts + warp scan) and one 64-key bitonic sort orders them.  Keys are
//   (score, index) packed into an order-preserving 64-bit integer, so ties resolve to the
//   lower index exactly as in the oracle.  More than 64 survivors (adversarial layouts)
//   fall back to k rounds of warp arg-max.
constexpr int KT = 128;        // query rows per work item = candidates per block
constexpr int KT_KC = 32;      // channels per chunk
constexpr int KT_THREADS = 256;
constexpr int KT_SMEM = sizeof(float) * (2 * KT_KC * KT) + sizeof(unsigned long long) * 8 * 64;

// (score, index) as one unsigned key: larger = better (higher score, then lower index)
__device__ __forceinline__ unsigned long long knn_key(float v, int j) {
  const unsigned b = __float_as_uint(v);
  const unsigned o = (b & 0x80000000u) ? ~b : (b | 0x80000000u);  // order-preserving float -> uint
  return ((unsigned long long)o << 32) | (unsigned long long)(~(unsigned)j);
}
__device__ __forceinline__ unsigned long long u64max(unsigned long long a, unsigned long long b) { return a > b ? a : b; }
__device__ __forceinline__ unsigned long long u64min(unsigned long long a, unsigned long long b) { return a > b ? b : a; }
// bitonic sort of 32 keys (one per lane) into descending order
__device__ __forceinline__ void knn_sort32(unsigned long long& a, int lane) {
#pragma unroll
  for (int kk = 2; kk <= 32; kk <<= 1) {
#pragma unroll
    for (int j = kk >> 1; j > 0; j >>= 1) {
      const unsigned long long pa = __shfl_xor_sync(0xffffffffu, a, j);
      const bool lower = (lane & j) == 0;
      const bool desc = (lane & kk) == 0 || kk == 32;
      a = (lower == desc) ? u64max(a, pa) : u64min(a, pa);
    }
  }
}
// bitonic sort of the 64 keys (a of lane l = element l, b = element 32 + l) into descending order
__device__ __forceinline__ void knn_sort64(unsigned long long& a, unsigned long long& b, int lane) {
#pragma unroll
  for (int kk = 2; kk <= 64; kk <<= 1) {
#pragma unroll
    for (int j = kk >> 1; j > 0; j >>= 1) {
      if (j == 32) {  // partner is the other register of the same lane; kk == 64: descending
        const unsigned long long hi = u64max(a, b), lo = u64min(a, b);
        a = hi; b = lo;
      } else {
        const unsigned long long pa = __shfl_xor_sync(0xffffffffu, a, j);
        const unsigned long long pb = __shfl_xor_sync(0xffffffffu, b, j);
        const bool lower = (lane & j) == 0;
        const bool desc_a = (lane & kk) == 0;          // element index = lane
        const bool desc_b = ((32 + lane) & kk) == 0;   // element index = 32 + lane
        a = (lower == desc_a) ? u64max(a, pa) : u64min(a, pa);
        b = (lower == desc_b) ? u64max(b, pb) : u64min(b, pb);
      }
    }
  }
}

// rare path: k rounds of warp arg-max straight from the score row (keys are unique, so
// round s takes the largest key below the previous winner)
__device__ __noinline__ void knn_select_slow(const float* __restrict__ Sr, int N, int k, int lane,
                                             int* __restrict__ out) {
  unsigned long long prev = ~0ull;
  for (int s_ = 0; s_ < k; ++s_) {
    unsigned long long best = 0ull;
    for (int j = lane; j < N; j += 32) {
      const unsigned long long key = knn_key(Sr[j], j);
      if (key < prev) best = u64max(best, key);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) best = u64max(best, __shfl_xor_sync(0xffffffffu, best, o));
    if (lane == 0) out[s_] = (int)(~(unsigned)(best & 0xffffffffull));
    prev = best;
  }
}

__global__ void __launch_bounds__(KT_THREADS, 1)
knn_tile_kernel(const float* __restrict__ x, const float* __restrict__ xx,
                const float* __restrict__ valids, int n_parts, int N, int C,
                int k, float* __restrict__ scratch, int* __restrict__ idx) {
  extern __shared__ float sm[];
  float* As = sm;                       // [KT_KC][KT]  queries, channel-major
  float* Bs = As + KT_KC * KT;          // [KT_KC][KT]  candidates
  unsigned long long* cbuf = reinterpret_cast<unsigned long long*>(Bs + KT_KC * KT);  // [8 warps][64]
  const int tiles = (N + KT - 1) / KT;
  const int Np = tiles * KT;
  float* S = scratch + (size_t)blockIdx.x * KT * Np;  // this CTA's score slab [KT][Np]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int tx = tid & 15, ty = tid >> 4;  // 16 x 16 threads; rows {4ty+r, 64+4ty+r}, cols {4tx+c, 64+4tx+c}
  const float ninf = -__int_as_float(0x7f800000);
  const int nchunks = (C + KT_KC - 1) / KT_KC;
  const int lr = tid & 127, lq = tid >> 7;  // loader role: tile row, channel quads lq, lq+2, lq+4, lq+6

  for (int work = blockIdx.x; work < n_parts * tiles; work += gridDim.x) {
    const int part = work / tiles, i0 = (work % tiles) * KT;
    if (valids != nullptr && valids[part] == 0.0f) continue;  // padded part (CTA-uniform)
    const float* xp = x + (long long)part * N * C;
    const float* xxp = xx + (long long)part * N;
    float4 pa[4], pb[4];
    auto fetch = [&](int j0, int kc) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int c = kc + 4 * (lq + 2 * i);
        const int ra = i0 + lr, rb = j0 + lr;
        pa[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        pb[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (c + 3 < C && (C & 3) == 0) {
          if (ra < N) pa[i] = *reinterpret_cast<const float4*>(xp + (long long)ra * C + c);
          if (rb < N) pb[i] = *reinterpret_cast<const float4*>(xp + (long long)rb * C + c);
        } else if (c < C) {  // unaligned rows / ragged channel tail
          float ta[4] = {0.f, 0.f, 0.f, 0.f}, tb[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            if (c + e < C) {
              if (ra < N) ta[e] = xp[(long long)ra * C + c + e];
              if (rb < N) tb[e] = xp[(long long)rb * C + c + e];
            }
          }
          pa[i] = make_float4(ta[0], ta[1], ta[2], ta[3]);
          pb[i] = make_float4(tb[0], tb[1], tb[2], tb[3]);
        }
      }
    };
    auto stash = [&]() {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int kk = 4 * (lq + 2 * i);
        As[(kk + 0) * KT + lr] = pa[i].x; As[(kk + 1) * KT + lr] = pa[i].y;
        As[(kk + 2) * KT + lr] = pa[i].z; As[(kk + 3) * KT + lr] = pa[i].w;
        Bs[(kk + 0) * KT + lr] = pb[i].x; Bs[(kk + 1) * KT + lr] = pb[i].y;
        Bs[(kk + 2) * KT + lr] = pb[i].z; Bs[(kk + 3) * KT + lr] = pb[i].w;
      }
    };

    fetch(0, 0);
    for (int jb = 0; jb < tiles; ++jb) {
      const int j0 = jb * KT;
      float acc[8][8];
#pragma unroll
      for (int r = 0; r < 8; ++r)
#pragma unroll
        for (int c = 0; c < 8; ++c) acc[r][c] = 0.f;
      for (int ch = 0; ch < nchunks; ++ch) {
        const int kc = ch * KT_KC;
        __syncthreads();  // the previous chunk's multiplies are done
        stash();
        __syncthreads();
        // prefetch the next chunk (next channel slice, or the first slice of the next block)
        if (ch + 1 < nchunks) fetch(j0, kc + KT_KC);
        else if (jb + 1 < tiles) fetch(j0 + KT, 0);
        const int klen = min(KT_KC, ((C - kc) + 3) & ~3);
#pragma unroll 4
        for (int kk = 0; kk < klen; ++kk) {
          const float4 a0 = *reinterpret_cast<const float4*>(&As[kk * KT + 4 * ty]);
          const float4 a1 = *reinterpret_cast<const float4*>(&As[kk * KT + 64 + 4 * ty]);
          const float4 b0 = *reinterpret_cast<const float4*>(&Bs[kk * KT + 4 * tx]);
          const float4 b1 = *reinterpret_cast<const float4*>(&Bs[kk * KT + 64 + 4 * tx]);
          const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
          const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
          for (int r = 0; r < 8; ++r)
#pragma unroll
            for (int c = 0; c < 8; ++c) acc[r][c] = __fmaf_rn(a[r], b[c], acc[r][c]);  // ascending channel
        }
      }
      // ---- scores of this block -> the CTA's slab (L2) ----
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        const int row = (r < 4 ? 0 : 60) + 4 * ty + r;  // 4ty+r or 64+4ty+(r-4)
        const float xi = (i0 + row < N) ? xxp[i0 + row] : 0.f;
        float sv[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const int j = j0 + (c < 4 ? 0 : 60) + 4 * tx + c;
          float sc = ninf;
          if (j < N) {
            const float inner = __fmul_rn(-2.0f, acc[r][c]);          // dgcnn.py:10
            sc = __fsub_rn(__fsub_rn(-xxp[j], inner), xi);            // dgcnn.py:12
          }
          sv[c] = sc;
        }
        float* dst = S + (size_t)row * Np + j0;
        *reinterpret_cast<float4*>(dst + 4 * tx) = make_float4(sv[0], sv[1], sv[2], sv[3]);
        *reinterpret_cast<float4*>(dst + 64 + 4 * tx) = make_float4(sv[4], sv[5], sv[6], sv[7]);
      }
    }
    __syncthreads();  // all score rows of this work item are written (same CTA: visible after the barrier)

    // ---- top-k per row: warp w owns rows 16w .. 16w+15 (two rows at a time measured slower:
    // 2.39 vs 1.93 ms at C = 3) ----
    unsigned long long* wb = cbuf + warp * 64;
    for (int rr = 0; rr < 16; ++rr) {
      const int row = warp * 16 + rr;
      if (i0 + row >= N) break;
      const float* Sr = S + (size_t)row * Np;
      float v[32];
#pragma unroll
      for (int c = 0; c < 32; ++c) {
        const int j = lane + 32 * c;
        v[c] = j < Np ? __ldcg(Sr + j) : ninf;  // columns >= N hold -inf already
      }
      // lane maximum of the strided slice (first maximum = lowest index)
      float mv = v[0];
      int mc = 0;
#pragma unroll
      for (int c = 1; c < 32; ++c)
        if (v[c] > mv) { mv = v[c]; mc = c; }
      unsigned long long t = (lane + 32 * mc < N) ? knn_key(mv, lane + 32 * mc) : 0ull;
      knn_sort32(t, lane);
      const unsigned long long T0 = __shfl_sync(0xffffffffu, t, k - 1);  // <= the k-th best key
      int cnt = 0;
#pragma unroll
      for (int c = 0; c < 32; ++c) {
        const int j = lane + 32 * c;
        cnt += (j < N && knn_key(v[c], j) >= T0) ? 1 : 0;
      }
      int incl = cnt;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int up = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += up;
      }
      const int total = __shfl_sync(0xffffffffu, incl, 31);
      int* out = idx + ((long long)part * N + i0 + row) * k;
      if (total <= 64) {
        int off = incl - cnt;
        __syncwarp();
#pragma unroll
        for (int c = 0; c < 32; ++c) {
          const int j = lane + 32 * c;
          const unsigned long long key = knn_key(v[c], j);
          if (j < N && key >= T0) wb[off++] = key;
        }
        __syncwarp();
        unsigned long long r0 = lane < total ? wb[lane] : 0ull;
        unsigned long long r1 = lane + 32 < total ? wb[lane + 32] : 0ull;
        knn_sort64(r0, r1, lane);
        if (lane < k) out[lane] = (int)(~(unsigned)(r0 & 0xffffffffull));
      } else {
        knn_select_slow(Sr, N, k, lane, out);  // > 64 survivors: adversarial layouts
      }
    }
    __syncthreads();  // selection done before the next work item overwrites the slab
  }
}

// ---- tensor-core variant: tcgen05 candidate filter + exact decision at the boundary ------
// The score matrix is a Gram matrix, so the contraction belongs on the tensor cores; what
// must not change is the RESULT, the oracle's top-k set under its own fp32 arithmetic
// (sequential-k FMA, dgcnn.py:8-15).  Per chunk of parts:
//   1. `launch_gram_batched` / `launch_gram_candidates` (csrc/linear.cu, two bf16 planes per operand):
//      t_ij = x_i . x_j - |x_j|^2 / 2, which orders the candidates of row i like the score does;
//   2. `knn_select_kernel`, one warp per row at full occupancy: lane maxima -> lower bound of
//      the k-th best -> candidates within `delta` of it (delta bounds twice the worst
//      difference between t and the oracle's score / 2: tensor-core accumulation, dropped
//      plane products, the oracle's own roundings, all <= alpha (|x_i|^2 + max_j |x_j|^2)) ->
//      one 64-key sort.  If the k-th and (k+1)-th approximate scores are more than delta apart
//      the approximate top k IS the oracle's set.  Otherwise the candidates within delta of the
//      k-th are re-scored exactly as the oracle scores them and the boundary is decided on
//      those values (ties -> lower index).  More than 64 candidates: the row is re-scored
//      exactly as a whole and selected by arg-max rounds.
// Only the set is contractual (EdgeConv takes a max over the k edges); the order written is
// approximate-score order with the exactly decided boundary members last.
// One pass over x: |x|^2 in the oracle's order (sequential-k FMA), -|x|^2 / 2 (the bias of the Gram
// products), the per-part maximum, and the two bf16 operand planes [2][rows][Kp] (hi, mid;
// channels >= C zero).  A warp takes 32 rows: coalesced loads into a padded shared-memory tile, the
// planes leave coalesced as well, then lane r walks row r.
constexpr int KP_THREADS = 256;
constexpr int KP_CC = 32;  // channels per tile pass
__global__ void __launch_bounds__(KP_THREADS)
knn_prepare_kernel(const float* __restrict__ x, long long rows, int N, int C, int Kp,
                   float* __restrict__ xx, float* __restrict__ hb, unsigned* __restrict__ xxmax,
                   __nv_bfloat16* __restrict__ planes) {
  __shared__ float tile[KP_THREADS / 32][32][KP_CC + 1];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long plane = rows * (long long)Kp;
  const long long n_blocks = (rows + 31) / 32;
  for (long long blk = (long long)blockIdx.x * (KP_THREADS / 32) + warp; blk < n_blocks;
       blk += (long long)gridDim.x * (KP_THREADS / 32)) {
    const long long r0 = blk * 32;
    float s = 0.f;
    for (int c0 = 0; c0 < Kp; c0 += KP_CC) {
      // rows r0 .. r0+31, channels c0 .. c0+31: 8 lanes x 4 channels per row, 4 rows per instruction
      const bool vec = (C & 3) == 0 && (Kp & 3) == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0;
      for (int it = 0; it < 8; ++it) {
        const int rr = it * 4 + (lane >> 3);
        const long long r = r0 + rr;
        const int c = c0 + 4 * (lane & 7);
        float v[4] = {0.f, 0.f, 0.f, 0.f};
        if (r < rows) {
          if (vec && c + 3 < C) {
            const float4 q4 = *reinterpret_cast<const float4*>(x + r * C + c);
            v[0] = q4.x; v[1] = q4.y; v[2] = q4.z; v[3] = q4.w;
          } else {
#pragma unroll
            for (int e = 0; e < 4; ++e)
              if (c + e < C) v[e] = x[r * C + c + e];
          }
        }
#pragma unroll
        for (int e = 0; e < 4; ++e) tile[warp][rr][4 * (lane & 7) + e] = v[e];
        if (r < rows && c < Kp) {  // Kp is a multiple of 8: the whole quad is inside
          const long long eo = r * Kp + c;
#pragma unroll
          for (int pl = 0; pl < 2; ++pl) {
            __nv_bfloat16 h[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              h[e] = __float2bfloat16_rn(v[e]);
              v[e] -= __bfloat162float(h[e]);
            }
            *reinterpret_cast<uint2*>(planes + pl * plane + eo) = *reinterpret_cast<const uint2*>(h);
          }
        }
      }
      __syncwarp();
      const int cn = min(KP_CC, C - c0);
      for (int c = 0; c < cn; ++c) s = __fmaf_rn(tile[warp][lane][c], tile[warp][lane][c], s);  // ascending channels
      __syncwarp();
    }
    const long long r = r0 + lane;
    if (r < rows) {
      xx[r] = s;
      hb[r] = -0.5f * s;
    }
    // per-part maximum: one atomic per warp and part (a 32-row block touches at most two parts when N >= 32)
    const int part = r < rows ? (int)(r / N) : -1;
    const int part0 = __shfl_sync(0xffffffffu, part, 0);
    const bool uniform = __all_sync(0xffffffffu, part == part0);
    if (uniform) {
      float m = s;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
      if (lane == 0 && part0 >= 0) atomicMax(xxmax + part0, __float_as_uint(m));
    } else if (part >= 0) {
      atomicMax(xxmax + part, __float_as_uint(s));  // s >= 0: unsigned order = float order
    }
  }
}

// the oracle's score of the pair (i, j): sequential-k FMA, then dgcnn.py:10,12
__device__ __forceinline__ float knn_exact_score(const float* __restrict__ xi, const float* __restrict__ xj,
                                                 int C, float xxi, float xxj) {
  float acc = 0.f;
  for (int c = 0; c < C; ++c) acc = __fmaf_rn(xi[c], xj[c], acc);
  const float inner = __fmul_rn(-2.0f, acc);
  return __fsub_rn(__fsub_rn(-xxj, inner), xxi);
}
__device__ __forceinline__ float knn_key_value(unsigned long long key) {
  const unsigned o = (unsigned)(key >> 32);
  return __uint_as_float((o & 0x80000000u) ? (o & 0x7fffffffu) : ~o);
}
__device__ __forceinline__ int knn_key_index(unsigned long long key) { return (int)(~(unsigned)(key & 0xffffffffull)); }

// <= 64 candidate keys of one row (r0 = candidate `lane`, r1 = candidate `lane + 32`, 0 = none):
// sort by approximate score; if the k-th and (k+1)-th are more than `delta` apart the first k are
// the oracle's set, otherwise the candidates within delta of the k-th are re-scored exactly and
// the boundary is decided on those values (see the comment above knn_prepare_kernel).
__device__ __forceinline__ void knn_finish_row(unsigned long long r0, unsigned long long r1, int total, int lane,
                                               unsigned long long* wb, float delta, int k,
                                               const float* __restrict__ xp, const float* __restrict__ xxp,
                                               float xxi, int i, int C, int* __restrict__ out) {
  if (total <= 32) knn_sort32(r0, lane);  // warp-uniform; the common case
  else knn_sort64(r0, r1, lane);          // approximate order, best first
  const float tk = knn_key_value(__shfl_sync(0xffffffffu, r0, k - 1));
  const unsigned long long next = k < 32 ? __shfl_sync(0xffffffffu, r0, k & 31) : __shfl_sync(0xffffffffu, r1, 0);
  const bool clear = total <= k || next == 0ull || tk - knn_key_value(next) > delta;
  if (clear) {  // the k-th and (k+1)-th are further apart than any rounding can move them
    if (lane < k) out[lane] = knn_key_index(r0);
    return;
  }
  // boundary decision on exact scores: `sure` = beats the k-th by more than delta (a prefix of
  // the sorted list, < k long), `band` = within delta of the k-th (the positions after it)
  const float hi = tk + delta, lo = tk - delta;
  const bool s0 = r0 != 0ull && knn_key_value(r0) > hi, s1 = r1 != 0ull && knn_key_value(r1) > hi;
  const bool b0 = r0 != 0ull && !s0 && knn_key_value(r0) >= lo, b1 = r1 != 0ull && !s1 && knn_key_value(r1) >= lo;
  const int m = __popc(__ballot_sync(0xffffffffu, s0)) + __popc(__ballot_sync(0xffffffffu, s1));
  const int nb = __popc(__ballot_sync(0xffffffffu, b0)) + __popc(__ballot_sync(0xffffffffu, b1));
  __syncwarp();
  wb[lane] = r0;
  wb[lane + 32] = r1;
  __syncwarp();
  unsigned long long e0 = 0ull, e1 = 0ull;
  if (lane < nb) {
    const int j = knn_key_index(wb[m + lane]);
    e0 = knn_key(knn_exact_score(xp + (long long)i * C, xp + (long long)j * C, C, xxi, xxp[j]), j);
  }
  if (lane + 32 < nb) {
    const int j = knn_key_index(wb[m + lane + 32]);
    e1 = knn_key(knn_exact_score(xp + (long long)i * C, xp + (long long)j * C, C, xxi, xxp[j]), j);
  }
  knn_sort64(e0, e1, lane);  // exact order of the band, best first
  const unsigned long long pick = __shfl_sync(0xffffffffu, e0, (lane - m) & 31);  // k - m <= 32
  if (lane < m) out[lane] = knn_key_index(r0);
  else if (lane < k) out[lane] = knn_key_index(pick);
}

constexpr int KS_THREADS = 256;
__global__ void __launch_bounds__(KS_THREADS)
knn_select_kernel(float* __restrict__ S, int z0, int items, const float* __restrict__ x,
                  const float* __restrict__ xx, const unsigned* __restrict__ xxmax,
                  const float* __restrict__ valids, int N, int C, int k, float alpha, int* __restrict__ idx) {
  __shared__ unsigned long long cbuf[KS_THREADS / 32][64];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long row_local = (long long)blockIdx.x * (KS_THREADS / 32) + warp;
  if (row_local >= (long long)items * N) return;
  const int part = z0 + (int)(row_local / N), i = (int)(row_local % N);
  if (valids != nullptr && valids[part] == 0.0f) return;
  float* Sr = S + row_local * N;
  const float ninf = -__int_as_float(0x7f800000);
  unsigned long long* wb = cbuf[warp];
  float v[32];
#pragma unroll
  for (int c = 0; c < 32; ++c) {
    const int j = lane + 32 * c;
    v[c] = j < N ? __ldcs(Sr + j) : ninf;
  }
  float mv = v[0];
  int mc = 0;
#pragma unroll
  for (int c = 1; c < 32; ++c)
    if (v[c] > mv) { mv = v[c]; mc = c; }
  unsigned long long t = (lane + 32 * mc < N) ? knn_key(mv, lane + 32 * mc) : 0ull;
  knn_sort32(t, lane);
  const unsigned long long T0 = __shfl_sync(0xffffffffu, t, k - 1);  // <= the k-th best approximate key
  const float xxi = xx[(long long)part * N + i];
  const float delta = alpha * (xxi + __uint_as_float(xxmax[part]));
  const float thr = T0 != 0ull ? knn_key_value(T0) - delta : ninf;  // fewer than k lane maxima: keep all
  int cnt = 0;
#pragma unroll
  for (int c = 0; c < 32; ++c) cnt += (lane + 32 * c < N && v[c] >= thr) ? 1 : 0;
  int incl = cnt;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int up = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += up;
  }
  const int total = __shfl_sync(0xffffffffu, incl, 31);
  int* out = idx + ((long long)part * N + i) * k;
  const float* xp = x + (long long)part * N * C;
  const float* xxp = xx + (long long)part * N;
  if (total <= 64) {
    int off = incl - cnt;
    __syncwarp();
#pragma unroll
    for (int c = 0; c < 32; ++c) {
      const int j = lane + 32 * c;
      if (j < N && v[c] >= thr) wb[off++] = knn_key(v[c], j);
    }
    __syncwarp();
    unsigned long long r0 = lane < total ? wb[lane] : 0ull;
    unsigned long long r1 = lane + 32 < total ? wb[lane + 32] : 0ull;
    knn_finish_row(r0, r1, total, lane, wb, delta, k, xp, xxp, xxi, i, C, out);
    return;
  }
  // more than 64 candidates (massive ties / adversarial layouts): exact scores for the row
  for (int j = lane; j < N; j += 32)
    Sr[j] = knn_exact_score(xp + (long long)i * C, xp + (long long)j * C, C, xxi, xxp[j]);
  __syncwarp();
  knn_select_slow(Sr, N, k, lane, out);
}

// slab-free path: the Gram kernel left <= 64 candidate keys per row (count > 64: overflow)
__global__ void __launch_bounds__(KS_THREADS)
knn_finish_kernel(const unsigned long long* __restrict__ cand, const int* __restrict__ count,
                  const float* __restrict__ x, const float* __restrict__ xx, const unsigned* __restrict__ xxmax,
                  const float* __restrict__ valids, long long rows, int N, int C, int k, float alpha,
                  int* __restrict__ idx) {
  __shared__ unsigned long long cbuf[KS_THREADS / 32][64];
  __shared__ float srow[KS_THREADS / 32][1024];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long row = (long long)blockIdx.x * (KS_THREADS / 32) + warp;
  if (row >= rows) return;
  const int part = (int)(row / N), i = (int)(row % N);
  if (valids != nullptr && valids[part] == 0.0f) return;
  const int total = count[row];
  const float xxi = xx[row];
  const float delta = alpha * (xxi + __uint_as_float(xxmax[part]));
  int* out = idx + row * k;
  const float* xp = x + (long long)part * N * C;
  const float* xxp = xx + (long long)part * N;
  if (total <= 64) {
    const unsigned long long* cl = cand + row * 64;
    const unsigned long long r0 = lane < total ? cl[lane] : 0ull;
    const unsigned long long r1 = lane + 32 < total ? cl[lane + 32] : 0ull;
    knn_finish_row(r0, r1, total, lane, cbuf[warp], delta, k, xp, xxp, xxi, i, C, out);
    return;
  }
  // overflow (massive ties / adversarial layouts / tiny clouds): exact scores for the whole row
  float* Sr = srow[warp];
  for (int j = lane; j < N; j += 32)
    Sr[j] = knn_exact_score(xp + (long long)i * C, xp + (long long)j * C, C, xxi, xxp[j]);
  __syncwarp();
  knn_select_slow(Sr, N, k, lane, out);
}

// uv [M, 2*Co] (u | v), idx [M, k] (indices local to the part), M = n*N.
// One CTA per PTS points, thread = channel.  ymax/ymin [M, Co]; partial [gridDim, Co, 2].
constexpr int EC_PTS = 8;
__global__ void edge_aggregate_kernel(const float* __restrict__ uv, const int* __restrict__ idx,
                                      const float* __restrict__ valids, long long M,
                                      int N, int Co, int k, float* __restrict__ ymax,
                                      float* __restrict__ ymin, float* __restrict__ partial) {
  // element offset of every neighbour's u row, worked out once per CTA (the address arithmetic of
  // the gather used to be half of the kernel's instructions); -1 = point of a padded part
  extern __shared__ long long soff[];  // [EC_PTS][k]
  const long long p0 = (long long)blockIdx.x * EC_PTS;
  const int two_co = 2 * Co;
  for (int e = threadIdx.x; e < EC_PTS * k; e += blockDim.x) {
    const int q = e / k, j = e - q * k;
    const long long p = p0 + q;
    long long off = -1;
    if (p < M) {
      const unsigned part = (unsigned)p / (unsigned)N;  // M < 2^31 (checked on the host)
      // padded parts have no graph (mpa_knn skipped them): never dereference their slots
      if (valids == nullptr || valids[part] != 0.0f)
        off = ((long long)part * N + idx[p * k + j]) * two_co;
    }
    soff[e] = off;
  }
  __syncthreads();
  const int c = threadIdx.x;
  float s1 = 0.f, s2 = 0.f;
  if (c < Co) {
    const float* __restrict__ uc = uv + c;
    for (int q = 0; q < EC_PTS; ++q) {
      const long long p = p0 + q;
      if (p >= M) break;
      const long long* so = soff + q * k;
      if (so[0] < 0) {  // point of a padded part: zero features, no BatchNorm contribution
        ymax[p * Co + c] = 0.f;
        ymin[p * Co + c] = 0.f;
        continue;
      }
      const float v = __ldg(uc + p * two_co + Co);
      float mx = -3.0e38f, mn = 3.0e38f;
#pragma unroll 4
      for (int e = 0; e < k; ++e) {
        const float y = __ldg(uc + so[e]) + v;
        mx = fmaxf(mx, y); mn = fminf(mn, y);
        s1 += y; s2 = fmaf(y, y, s2);
      }
      ymax[p * Co + c] = mx;
      ymin[p * Co + c] = mn;
    }
    partial[((long long)blockIdx.x * Co + c) * 2] = s1;
    partial[((long long)blockIdx.x * Co + c) * 2 + 1] = s2;
  }
}

// deterministic column sums of partial [rows, Co, 2] -> sums [Co, 2] (double accumulate):
// stage 1 reduces CS_SLICES row slices per channel (one warp each), stage 2 adds the
// slices in a fixed order.
constexpr int CS_SLICES = 64;
__global__ void column_sum_stage1_kernel(const float* __restrict__ partial, long long rows, int Co,
                                         double* __restrict__ slices) {
  const int c = blockIdx.x, sl = blockIdx.y, lane = threadIdx.x;
  const long long per = (rows + CS_SLICES - 1) / CS_SLICES;
  const long long r0 = sl * per, r1 = min(rows, r0 + per);
  double s1 = 0, s2 = 0;
  for (long long r = r0 + lane; r < r1; r += 32) {
    s1 += (double)partial[(r * Co + c) * 2];
    s2 += (double)partial[(r * Co + c) * 2 + 1];
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s1 += __shfl_xor_sync(0xffffffffu, s1, o);
    s2 += __shfl_xor_sync(0xffffffffu, s2, o);
  }
  if (lane == 0) { slices[((long long)sl * Co + c) * 2] = s1; slices[((long long)sl * Co + c) * 2 + 1] = s2; }
}
__global__ void column_sum_stage2_kernel(const double* __restrict__ slices, int Co, double* __restrict__ sums) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= Co) return;
  double s1 = 0, s2 = 0;
  for (int sl = 0; sl < CS_SLICES; ++sl) {
    s1 += slices[((long long)sl * Co + c) * 2];
    s2 += slices[((long long)sl * Co + c) * 2 + 1];
  }
  sums[2 * c] = s1;
  sums[2 * c + 1] = s2;
}


// ---- BatchNorm + LeakyReLU epilogues of the EdgeConv layers and of conv5 ----------------
// number of valid parts (valids == nullptr: all n); every thread of the CTA gets the result
__device__ __forceinline__ float block_valid_parts(const float* __restrict__ valids, int n, float* s_red) {
  if (valids == nullptr) return (float)n;
  float c = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) c += valids[i] != 0.0f ? 1.f : 0.f;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = c;
  __syncthreads();
  float t = 0.f;
  for (int w = 0; w < (int)((blockDim.x + 31) >> 5); ++w) t += s_red[w];
  __syncthreads();
  return t;
}

// scale / shift of channel c from batch sums (training) or running statistics (eval);
// CTA 0 also applies torch's running-statistics update (momentum, unbiased variance)
__device__ __forceinline__ void bn_channel_affine(int c, const double* __restrict__ sums, double count,
                                                  const float* __restrict__ w, const float* __restrict__ b,
                                                  float* running_mean, float* running_var, int training,
                                                  float momentum, float eps, bool update, float& scale,
                                                  float& shift) {
  double mean, var;
  if (training) {
    mean = sums[2 * c] / count;
    var = fmax(sums[2 * c + 1] / count - mean * mean, 0.0);
    if (update) {
      const double unbias = count / fmax(count - 1.0, 1.0);
      running_mean[c] = (float)((1.0 - momentum) * (double)running_mean[c] + momentum * mean);
      running_var[c] = (float)((1.0 - momentum) * (double)running_var[c] + momentum * var * unbias);
    }
  } else {
    mean = (double)running_mean[c];
    var = (double)running_var[c];
  }
  const double sc = (double)w[c] / sqrt(var + (double)eps);
  scale = (float)sc;
  shift = (float)((double)b[c] - mean * sc);
}

// h = LeakyReLU(BN(max over the k edges)) from the aggregate's ymax / ymin (the affine map is
// monotone per channel: the max over k of scale*y + shift is scale*ymax + shift for scale >= 0,
// scale*ymin + shift otherwise).  Writes the layer output contiguously (next layer's input)
// and into its column slice of the [M, ldc] concatenation that conv5 reads.
constexpr int EF_THREADS = 256;
__global__ void __launch_bounds__(EF_THREADS)
edgeconv_finish_kernel(const float* __restrict__ ymax, const float* __restrict__ ymin,
                       const double* __restrict__ sums, const float* __restrict__ valids, int n, int N,
                       int Co, int k_edges, const float* __restrict__ bn_w, const float* __restrict__ bn_b,
                       float* running_mean, float* running_var, int training, float momentum, float eps,
                       float slope, float* __restrict__ out, float* __restrict__ out_cat, int ldc, int c0) {
  extern __shared__ float s_aff[];  // [2][Co]
  __shared__ float s_red[EF_THREADS / 32];
  const float nv = block_valid_parts(valids, n, s_red);
  const double count = (double)nv * (double)N * (double)k_edges;
  for (int c = threadIdx.x; c < Co; c += EF_THREADS)
    bn_channel_affine(c, sums, count, bn_w, bn_b, running_mean, running_var, training, momentum, eps,
                      blockIdx.x == 0, s_aff[c], s_aff[Co + c]);
  __syncthreads();
  const long long M = (long long)n * N, total4 = M * Co / 4;  // Co % 4 == 0 (checked by the host)
  for (long long e = (long long)blockIdx.x * EF_THREADS + threadIdx.x; e < total4;
       e += (long long)gridDim.x * EF_THREADS) {
    const long long p = e * 4 / Co;
    const int c = (int)(e * 4 % Co);
    const float4 a = *reinterpret_cast<const float4*>(ymax + e * 4);
    const float4 b = *reinterpret_cast<const float4*>(ymin + e * 4);
    float r[4];
    const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float sc = s_aff[c + j], sh = s_aff[Co + c + j];
      const float v = fmaf(sc, sc >= 0.f ? av[j] : bv[j], sh);
      r[j] = v > 0.f ? v : slope * v;
    }
    const float4 o = make_float4(r[0], r[1], r[2], r[3]);
    if (out != nullptr) *reinterpret_cast<float4*>(out + e * 4) = o;
    if (out_cat != nullptr) *reinterpret_cast<float4*>(out_cat + p * ldc + c0 + c) = o;
  }
}

// conv5 epilogue, pass 1: per part and channel the sum, sum of squares, max and min of
// y [n*N, F] over the part's N points (padded parts: zeros, outside the statistics)
// 256 threads per part: thread = (row group, channel); a row group walks the rows rg, rg + RG, ...
// four at a time (independent loads in flight), the groups are then added in index order, so the
// result does not depend on scheduling.  (One thread per channel and a 1000-step dependent loop
// reached 1.4 TB/s.)
constexpr int PCS_THREADS = 256;
__global__ void __launch_bounds__(PCS_THREADS)
part_channel_stats_kernel(const float* __restrict__ y, const float* __restrict__ valids, int N,
                          int F, float* __restrict__ partial /* [n, F, 2] */,
                          float* __restrict__ mm /* [n, F, 2] max, min */) {
  __shared__ float red[PCS_THREADS][4];
  const int part = blockIdx.x;
  const bool live = valids == nullptr || valids[part] != 0.0f;
  const int fw = F < PCS_THREADS ? F : PCS_THREADS;  // channels handled side by side
  const int RG = PCS_THREADS / fw;                    // row groups (1 when F >= 256)
  const int rg = threadIdx.x / fw, cl = threadIdx.x - rg * fw;
  for (int c0 = 0; c0 < F; c0 += fw) {
    const int c = c0 + cl;
    float s1 = 0.f, s2 = 0.f, mx = -3.0e38f, mn = 3.0e38f;
    if (live && c < F && rg < RG) {
      const float* col = y + (long long)part * N * F + c;
      int i = rg;
      for (; i + 3 * RG < N; i += 4 * RG) {
        const float v0 = col[(long long)i * F], v1 = col[(long long)(i + RG) * F];
        const float v2 = col[(long long)(i + 2 * RG) * F], v3 = col[(long long)(i + 3 * RG) * F];
        s1 += v0; s2 = fmaf(v0, v0, s2); mx = fmaxf(mx, v0); mn = fminf(mn, v0);
        s1 += v1; s2 = fmaf(v1, v1, s2); mx = fmaxf(mx, v1); mn = fminf(mn, v1);
        s1 += v2; s2 = fmaf(v2, v2, s2); mx = fmaxf(mx, v2); mn = fminf(mn, v2);
        s1 += v3; s2 = fmaf(v3, v3, s2); mx = fmaxf(mx, v3); mn = fminf(mn, v3);
      }
      for (; i < N; i += RG) {
        const float v = col[(long long)i * F];
        s1 += v; s2 = fmaf(v, v, s2); mx = fmaxf(mx, v); mn = fminf(mn, v);
      }
    }
    red[threadIdx.x][0] = s1; red[threadIdx.x][1] = s2; red[threadIdx.x][2] = mx; red[threadIdx.x][3] = mn;
    __syncthreads();
    if (rg == 0 && c < F) {
      for (int g2 = 1; g2 < RG; ++g2) {  // fixed order
        const float* o = red[g2 * fw + cl];
        s1 += o[0]; s2 += o[1]; mx = fmaxf(mx, o[2]); mn = fminf(mn, o[3]);
      }
      if (!live) { s1 = 0.f; s2 = 0.f; mx = 0.f; mn = 0.f; }
      partial[((long long)part * F + c) * 2] = s1;
      partial[((long long)part * F + c) * 2 + 1] = s2;
      mm[((long long)part * F + c) * 2] = mx;
      mm[((long long)part * F + c) * 2 + 1] = mn;
    }
    __syncthreads();
  }
}

// pass 2: g[part] = [ max_i a(y_i) | mean_i a(y_i) ], a = LeakyReLU o BatchNorm (dgcnn.py:97-107)
__global__ void bn_pool_kernel(const float* __restrict__ y, const float* __restrict__ mm,
                               const double* __restrict__ sums, const float* __restrict__ valids, int n, int N,
                               int F, const float* __restrict__ bn_w, const float* __restrict__ bn_b,
                               float* running_mean, float* running_var, int training, float momentum,
                               float eps, float slope, float* __restrict__ g /* [n, 2F] */) {
  __shared__ float s_red[32];
  const int part = blockIdx.x;
  const float nv = block_valid_parts(valids, n, s_red);
  const double count = (double)nv * (double)N;
  const bool live = valids == nullptr || valids[part] != 0.0f;
  for (int c = threadIdx.x; c < F; c += blockDim.x) {
    float sc, sh;
    bn_channel_affine(c, sums, count, bn_w, bn_b, running_mean, running_var, training, momentum, eps,
                      part == 0, sc, sh);
    float mx = 0.f, mean = 0.f;
    if (live) {
      const float top = fmaf(sc, sc >= 0.f ? mm[((long long)part * F + c) * 2] : mm[((long long)part * F + c) * 2 + 1], sh);
      mx = top > 0.f ? top : slope * top;
      const float* col = y + (long long)part * N * F + c;
      float s = 0.f;
      for (int i = 0; i < N; ++i) {
        const float v = fmaf(sc, col[(long long)i * F], sh);
        s += v > 0.f ? v : slope * v;
      }
      mean = s / (float)N;
    }
    g[(long long)part * 2 * F + c] = mx;
    g[(long long)part * 2 * F + F + c] = mean;
  }
}

}  // namespace mpa

using namespace mpa;

extern "C" {

static int knn_tile_ctas() {  // one persistent CTA (and one 128 x Np score slab) per SM
  return device_sms();
}

size_t mpa_knn_workspace_bytes(int n, int N) {
  const size_t Np = (size_t)((N + KT - 1) / KT) * KT;
  return align_up(sizeof(float) * (size_t)n * N, 256) +       // squared norms
         sizeof(float) * (size_t)knn_tile_ctas() * KT * Np;   // per-CTA score slabs (L2-resident)
}

// workspace of the tensor-core path: xx, hb, per-part max, operand planes, score slab of one chunk
static int knn_kpad(int C) { return (C + 7) & ~7; }
static int knn_chunk_items(int n, int N) {
  const size_t per_item = sizeof(float) * (size_t)N * N;
  static const size_t mb = getenv("MPA_KNN_CHUNK_MB") ? (size_t)atoi(getenv("MPA_KNN_CHUNK_MB")) : 256;
  size_t items = (mb << 20) / per_item;  // scores in flight per Gram + select launch pair (fewer, longer launches measured faster than L2-sized chunks)
  if (items < 1) items = 1;
  if (items > (size_t)n) items = n;
  return (int)items;
}
size_t mpa_knn_workspace_bytes_c(int n, int N, int C) {
  const size_t rows = (size_t)n * N;
  return mpa_knn_workspace_bytes(n, N) + 2 * align_up(sizeof(float) * rows, 256) + align_up(sizeof(unsigned) * n, 256) +
         align_up((size_t)2 * rows * knn_kpad(C) * 2, 256) +
         // the larger of: score slab of one chunk (slab path) / 64 candidate keys + a count per row
         std::max(align_up(sizeof(float) * (size_t)knn_chunk_items(n, N) * N * N, 256),
                  align_up(rows * 64 * sizeof(unsigned long long), 256) + align_up(rows * sizeof(int), 256));
}

static int knn_tensor_core(const float* x, const float* valids, int n, int N, int C, int k, int32_t* idx, char* p,
                           cudaStream_t stream) {
  const long long rows = (long long)n * N;
  const int Kp = knn_kpad(C);
  float* xx = (float*)p; p += align_up(sizeof(float) * rows, 256);
  float* hb = (float*)p; p += align_up(sizeof(float) * rows, 256);
  unsigned* xxmax = (unsigned*)p; p += align_up(sizeof(unsigned) * n, 256);
  __nv_bfloat16* planes = (__nv_bfloat16*)p; p += align_up((size_t)2 * rows * Kp * 2, 256);
  float* S = (float*)p;
  MPA_CUDA(cudaMemsetAsync(xxmax, 0, sizeof(unsigned) * n, stream));
  {
    ProfScope ps("knn_prepare", stream);
    knn_prepare_kernel<<<2 * device_sms(), KP_THREADS, 0, stream>>>(x, rows, N, C, Kp, xx, hb, xxmax, planes);
  }
  MPA_LAUNCH_CHECK();
  // bound of (approximate - exact) / (|x_i|^2 + |x_j|^2), in units of the approximate score:
  // the plane products left out (hi.lo, mid.mid, lo.hi, ...: <= 3 * 2^-18 |x_i| |x_j| <= 2^-17 of
  // the norm sum), 3 products x Kp/16 accumulation steps of the tensor core (4 ulp each, generous),
  // the bias subtraction, and the oracle's own C + 8 roundings; x 2: both sides of a comparison
  const float alpha = 2.0f * (7.6293945e-6f + (3.0f * (float)(Kp / 16 + 1) * 4.0f + 2.0f) * 1.1920929e-7f +
                              (float)(C + 8) * 5.9604645e-8f);
  // slab-free: candidates straight from the Gram kernel's epilogue (K <= 128, N <= 1024)
  static const bool use_slab = getenv("MPA_KNN_SLAB") != nullptr;  // A/B: scores through the slab
  if (!use_slab) {
    unsigned long long* cand = (unsigned long long*)S;
    int* count = (int*)((char*)S + align_up((size_t)rows * 64 * sizeof(unsigned long long), 256));
    const int rc = launch_gram_candidates(planes, rows, N, Kp, n, hb, valids, xx, xxmax, alpha, k, cand, count, stream);
    if (rc < 0) return rc;
    if (rc == MPA_OK) {
      ProfScope ps("knn_select", stream);
      knn_finish_kernel<<<(unsigned)((rows + KS_THREADS / 32 - 1) / (KS_THREADS / 32)), KS_THREADS, 0, stream>>>(
          cand, count, x, xx, xxmax, valids, rows, N, C, k, alpha, idx);
      MPA_LAUNCH_CHECK();
      return MPA_OK;
    }
  }
  const int chunk = knn_chunk_items(n, N);
  for (int z0 = 0; z0 < n; z0 += chunk) {
    const int items = n - z0 < chunk ? n - z0 : chunk;
    int rc = launch_gram_batched(planes, rows, N, Kp, z0, items, hb, valids, S, "knn_gram", stream);
    if (rc != MPA_OK) return rc;
    {
      ProfScope ps("knn_select", stream);
      const long long nrows = (long long)items * N;
      knn_select_kernel<<<(unsigned)((nrows + KS_THREADS / 32 - 1) / (KS_THREADS / 32)), KS_THREADS, 0, stream>>>(
          S, z0, items, x, xx, xxmax, valids, N, C, k, alpha, idx);
    }
    MPA_LAUNCH_CHECK();
  }
  return MPA_OK;
}

int mpa_knn(const float* x, const float* valids, int n, int N, int C, int k, int32_t* idx, void* ws,
            size_t ws_bytes, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  MPA_CHECK_ARG(n >= 0 && N > 0 && C > 0 && k > 0, "knn: bad sizes n=%d N=%d C=%d k=%d", n, N, C, k);
  MPA_CHECK_ARG(k <= N, "knn: k=%d exceeds the %d points of a part", k, N);
  MPA_CHECK_ARG(N <= 2048 && C <= 512, "knn: supports N <= 2048 points and C <= 512 channels");
  if (n == 0) return MPA_OK;
  MPA_CHECK_ARG(x && idx, "knn: null pointer");
  Scratch scratch;
  int rc = scratch.acquire(ws, ws_bytes, mpa_knn_workspace_bytes(n, N), stream);
  if (rc != MPA_OK) return rc;
  static const int legacy = getenv("MPA_KNN_LEGACY") ? atoi(getenv("MPA_KNN_LEGACY")) : 0;
  // tensor-core scoring when the caller sized the workspace for it (mpa_knn_workspace_bytes_c)
  static const bool force_tile = getenv("MPA_KNN_TILE") != nullptr;  // A/B: the CUDA-core tile kernel
  if (k <= 32 && N <= 1024 && knn_kpad(C) <= 128 && !legacy && !force_tile && ws != nullptr &&
      ws_bytes >= mpa_knn_workspace_bytes_c(n, N, C))
    return knn_tensor_core(x, valids, n, N, C, k, idx, (char*)scratch.base + mpa_knn_workspace_bytes(n, N), stream);
  float* xx = (float*)scratch.base;
  {
    ProfScope ps("knn_sqnorm", stream);
    sqnorm_kernel<<<592, 256, 0, stream>>>(x, (long long)n * N, C, xx);
  }
  MPA_LAUNCH_CHECK();
  if (k <= 32 && N <= 1024 && !legacy) {
    static DeviceOnce attr_t;
    if (attr_t.pending()) {
      MPA_CUDA(cudaFuncSetAttribute(knn_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, KT_SMEM));
      attr_t.done();
    }
    const int tiles = (N + KT - 1) / KT;
    int ctas = knn_tile_ctas();
    if (ctas > n * tiles) ctas = n * tiles;
    float* slab = (float*)((char*)scratch.base + align_up(sizeof(float) * (size_t)n * N, 256));
    {
      ProfScope ps("knn", stream);
      knn_tile_kernel<<<ctas, KT_THREADS, KT_SMEM, stream>>>(x, xx, valids, n, N, C, k, slab, idx);
    }
    MPA_LAUNCH_CHECK();
    return MPA_OK;
  }
  static const int rows_env = getenv("MPA_KNN_ROWS") ? atoi(getenv("MPA_KNN_ROWS")) : 0;
  const int rows = rows_env ? rows_env : 16;  // 16 rows: 2 CTAs per SM (profiles/: 42.0 -> 33.5 ms at cfg D)
  const int Cp = ((C + 3) & ~3) + 4;
  const int Np = (N + KNN_BN - 1) / KNN_BN * KNN_BN;
  const size_t smem = sizeof(float) * ((size_t)rows * Cp + KNN_BN * (KNN_KC + 4) + (size_t)rows * Np);
  const int tiles = (N + rows - 1) / rows;
  static DeviceOnce attr;
  if (attr.pending()) {
    MPA_CUDA(cudaFuncSetAttribute(knn_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    MPA_CUDA(cudaFuncSetAttribute(knn_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr.done();
  }
  MPA_CHECK_ARG(smem <= 227 * 1024, "knn: shared memory need %zu exceeds 227 KB", smem);
  {
    ProfScope ps("knn", stream);
    if (rows == 32) knn_kernel<32><<<n * tiles, KNN_THREADS, smem, stream>>>(x, xx, valids, N, C, k, idx);
    else knn_kernel<16><<<n * tiles, KNN_THREADS, smem, stream>>>(x, xx, valids, N, C, k, idx);
  }
  MPA_LAUNCH_CHECK();
  return MPA_OK;
}

size_t mpa_edge_aggregate_workspace_bytes(long long M, int Co) {
  const long long blocks = (M + EC_PTS - 1) / EC_PTS;
  return align_up(sizeof(float) * 2 * (size_t)blocks * Co, 256) +
         align_up(sizeof(double) * 2 * (size_t)CS_SLICES * Co, 256);
}

int mpa_edge_aggregate(const float* uv, const int32_t* idx, const float* valids, int n, int N, int Co,
                       int k, float* ymax, float* ymin, double* sums, void* ws, size_t ws_bytes,
                       void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  MPA_CHECK_ARG(n >= 0 && N > 0 && Co > 0 && Co <= 1024 && k > 0, "edge_aggregate: bad sizes");
  if (n == 0) return MPA_OK;
  MPA_CHECK_ARG(uv && idx && ymax && ymin && sums, "edge_aggregate: null pointer");
  const long long M = (long long)n * N;
  const long long blocks = (M + EC_PTS - 1) / EC_PTS;
  MPA_CHECK_ARG(M < (1ll << 31), "edge_aggregate: n * N must fit 31 bits");
  Scratch scratch;
  int rc = scratch.acquire(ws, ws_bytes, mpa_edge_aggregate_workspace_bytes(M, Co), stream);
  if (rc != MPA_OK) return rc;
  float* partial = (float*)scratch.base;
  const int threads = (Co + 31) / 32 * 32;
  {
    ProfScope ps("edge_aggregate", stream);
    edge_aggregate_kernel<<<(unsigned)blocks, threads, sizeof(long long) * EC_PTS * k, stream>>>(
        uv, idx, valids, M, N, Co, k, ymax, ymin, partial);
  }
  MPA_LAUNCH_CHECK();
  double* slices = (double*)((char*)scratch.base + align_up(sizeof(float) * 2 * (size_t)blocks * Co, 256));
  {
    ProfScope ps("edge_column_sum", stream);
    column_sum_stage1_kernel<<<dim3(Co, CS_SLICES), 32, 0, stream>>>(partial, blocks, Co, slices);
    column_sum_stage2_kernel<<<(Co + 127) / 128, 128, 0, stream>>>(slices, Co, sums);
  }
  count_launch();
  MPA_LAUNCH_CHECK();
  return MPA_OK;
}


int mpa_edgeconv_finish(const float* ymax, const float* ymin, const double* sums, const float* valids, int n,
                        int N, int Co, int k_edges, const float* bn_w, const float* bn_b, float* running_mean,
                        float* running_var, int training, float momentum, float eps, float slope, float* out,
                        float* out_cat, int ldc, int c0, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  MPA_CHECK_ARG(n >= 0 && N > 0 && Co > 0 && Co % 4 == 0 && Co <= 4096 && k_edges > 0, "edgeconv_finish: bad sizes");
  if (n == 0) return MPA_OK;
  MPA_CHECK_ARG(ymax && ymin && bn_w && bn_b && running_mean && running_var && (out || out_cat) &&
                    (training == 0 || sums != nullptr),
                "edgeconv_finish: null pointer");
  MPA_CHECK_ARG(out_cat == nullptr || (ldc % 4 == 0 && c0 % 4 == 0 && c0 + Co <= ldc),
                "edgeconv_finish: bad concatenation slice");
  const long long total4 = (long long)n * N * Co / 4;
  long long blocks = (total4 + EF_THREADS - 1) / EF_THREADS;
  const long long cap = (long long)device_sms() * 8;
  if (blocks > cap) blocks = cap;
  {
    ProfScope ps("edgeconv_finish", stream);
    edgeconv_finish_kernel<<<(unsigned)blocks, EF_THREADS, sizeof(float) * 2 * Co, stream>>>(
        ymax, ymin, sums, valids, n, N, Co, k_edges, bn_w, bn_b, running_mean, running_var, training, momentum,
        eps, slope, out, out_cat, ldc, c0);
  }
  MPA_LAUNCH_CHECK();
  return MPA_OK;
}

size_t mpa_bn_pool_workspace_bytes(int n, int F) {
  return 2 * align_up(sizeof(float) * 2 * (size_t)n * F, 256) + align_up(sizeof(double) * 2 * (size_t)CS_SLICES * F, 256) +
         align_up(sizeof(double) * 2 * (size_t)F, 256);
}

int mpa_bn_pool(const float* y, const float* valids, int n, int N, int F, const float* bn_w, const float* bn_b,
                float* running_mean, float* running_var, int training, float momentum, float eps, float slope,
                float* g, void* ws, size_t ws_bytes, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  MPA_CHECK_ARG(n >= 0 && N > 0 && F > 0 && F <= 4096, "bn_pool: bad sizes");
  if (n == 0) return MPA_OK;
  MPA_CHECK_ARG(y && bn_w && bn_b && running_mean && running_var && g, "bn_pool: null pointer");
  Scratch scratch;
  int rc = scratch.acquire(ws, ws_bytes, mpa_bn_pool_workspace_bytes(n, F), stream);
  if (rc != MPA_OK) return rc;
  char* p = (char*)scratch.base;
  float* partial = (float*)p; p += align_up(sizeof(float) * 2 * (size_t)n * F, 256);
  float* mm = (float*)p; p += align_up(sizeof(float) * 2 * (size_t)n * F, 256);
  double* slices = (double*)p; p += align_up(sizeof(double) * 2 * (size_t)CS_SLICES * F, 256);
  double* sums = (double*)p;
  const int threads = F >= 256 ? 256 : ((F + 31) / 32 * 32);
  {
    ProfScope ps("bn_pool_stats", stream);
    part_channel_stats_kernel<<<n, PCS_THREADS, 0, stream>>>(y, valids, N, F, partial, mm);
    column_sum_stage1_kernel<<<dim3(F, CS_SLICES), 32, 0, stream>>>(partial, n, F, slices);
    column_sum_stage2_kernel<<<(F + 127) / 128, 128, 0, stream>>>(slices, F, sums);
  }
  count_launch(2);
  MPA_LAUNCH_CHECK();
  {
    ProfScope ps("bn_pool", stream);
    bn_pool_kernel<<<n, threads, 0, stream>>>(y, mm, sums, valids, n, N, F, bn_w, bn_b, running_mean, running_var,
                                             training, momentum, eps, slope, g);
  }
  MPA_LAUNCH_CHECK();
  return MPA_OK;
}


/* (sum, sum of squares) per column of y [n_blocks * R, F] in fp64, deterministic: the BatchNorm
 * batch statistics of a shared-MLP layer (PointNet++ set abstraction, fp32-mode PointNet);
 * valids [n_blocks] (nullable): blocks flagged 0 stay out of the sums. */
size_t mpa_column_stats_workspace_bytes(int n_blocks, int F) { return mpa_bn_pool_workspace_bytes(n_blocks, F); }

int mpa_column_stats(const float* y, const float* valids, int n_blocks, int R, int F, double* sums, void* ws,
                     size_t ws_bytes, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  MPA_CHECK_ARG(n_blocks >= 0 && R > 0 && F > 0 && F <= 4096, "column_stats: bad sizes");
  if (n_blocks == 0) return MPA_OK;
  MPA_CHECK_ARG(y && sums, "column_stats: null pointer");
  Scratch scratch;
  int rc = scratch.acquire(ws, ws_bytes, mpa_column_stats_workspace_bytes(n_blocks, F), stream);
  if (rc != MPA_OK) return rc;
  char* p = (char*)scratch.base;
  float* partial = (float*)p; p += align_up(sizeof(float) * 2 * (size_t)n_blocks * F, 256);
  float* mm = (float*)p; p += align_up(sizeof(float) * 2 * (size_t)n_blocks * F, 256);
  double* slices = (double*)p;
  const int threads = F >= 256 ? 256 : ((F + 31) / 32 * 32);
  {
    ProfScope ps("column_stats", stream);
    part_channel_stats_kernel<<<n_blocks, PCS_THREADS, 0, stream>>>(y, valids, R, F, partial, mm);
    column_sum_stage1_kernel<<<dim3(F, CS_SLICES), 32, 0, stream>>>(partial, n_blocks, F, slices);
    column_sum_stage2_kernel<<<(F + 127) / 128, 128, 0, stream>>>(slices, F, sums);
  }
  count_launch(2);
  MPA_LAUNCH_CHECK();
  return MPA_OK;
}

}  // extern "C"
