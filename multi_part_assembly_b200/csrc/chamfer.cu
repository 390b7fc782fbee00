// Bidirectional Chamfer nearest-neighbour search for sm_100a.
//
// Replaces ChamferForwardKernel / ChamferBackwardKernel of the reference
// (utils/chamfer/cuda/chamfer_kernel.cu:32-95, 175-210) and, in the fused
// "pose" entry, the qrot/qtransform + masked_fill + chamfer_distance chain of
// utils/loss.py:113-202.
//
// Two exact algorithms, bit-identical results (distance in the reference's
// rounding sequence, ties -> lowest index):
//   BRUTE : one thread per query, targets staged through shared memory as
//           negated SoA, distances on packed FP32x2 (FADD2/FMUL2/FFMA2).
//   GRID  : per cloud a counting sort into a uniform grid (one CTA per cloud,
//           histogram + scan in shared memory), then every query walks the
//           3x3x3 block around its cell and grows the block ring by ring until
//           the best distance is provably smaller than the distance to any
//           unexplored cell.  ~10^2 candidate pairs per query instead of N.
#include <math.h>
#include <stdlib.h>

#include "mpa_common.cuh"

namespace mpa {

// ======================================================================
// BRUTE FORCE
// ======================================================================
constexpr int BF_THREADS = 256;
constexpr int BF_TILE = 2048;  // targets per shared-memory tile (24 KB)

template <typename IdxT>
__global__ void __launch_bounds__(BF_THREADS)
chamfer_brute_kernel(const float* __restrict__ xyzA, const float* __restrict__ xyzB, int B,
                     int nA, int nB, float* __restrict__ distA, IdxT* __restrict__ idxA,
                     float* __restrict__ distB, IdxT* __restrict__ idxB) {
  __shared__ __align__(16) float sx[BF_TILE];
  __shared__ __align__(16) float sy[BF_TILE];
  __shared__ __align__(16) float sz[BF_TILE];
  const int blocksA = (nA + BF_THREADS - 1) / BF_THREADS;
  const int blocksB = (nB + BF_THREADS - 1) / BF_THREADS;
  const long long totalA = (long long)B * blocksA;
  const long long total = totalA + (long long)B * blocksB;
  const float qnan = __int_as_float(0x7fc00000);

  for (long long w = blockIdx.x; w < total; w += gridDim.x) {
    const bool dirB = w >= totalA;
    const long long ww = dirB ? w - totalA : w;
    const float* __restrict__ Q = dirB ? xyzB : xyzA;
    const float* __restrict__ T = dirB ? xyzA : xyzB;
    const int nq = dirB ? nB : nA;
    const int nt = dirB ? nA : nB;
    const int nblk = dirB ? blocksB : blocksA;
    float* dout = dirB ? distB : distA;
    IdxT* iout = dirB ? idxB : idxA;
    const int b = (int)(ww / nblk);
    const int qi = (int)(ww % nblk) * BF_THREADS + threadIdx.x;

    float qx = 0.f, qy = 0.f, qz = 0.f;
    if (qi < nq) {
      const float* p = Q + ((long long)b * nq + qi) * 3;
      qx = p[0]; qy = p[1]; qz = p[2];
    }
    const unsigned long long qx2 = f2_as_u64(make_float2(qx, qx));
    const unsigned long long qy2 = f2_as_u64(make_float2(qy, qy));
    const unsigned long long qz2 = f2_as_u64(make_float2(qz, qz));
    float best = 1e32f;  // chamfer_kernel.cu:60
    int bidx = -1;       // chamfer_kernel.cu:61

    for (int t0 = 0; t0 < nt; t0 += BF_TILE) {
      const int cnt = min(BF_TILE, nt - t0);
      const int cnt4 = (cnt + 3) & ~3;
      __syncthreads();
      for (int j = threadIdx.x; j < cnt4; j += BF_THREADS) {
        if (j < cnt) {
          const float* p = T + ((long long)b * nt + t0 + j) * 3;
          sx[j] = -p[0]; sy[j] = -p[1]; sz[j] = -p[2];
        } else {
          sx[j] = qnan; sy[j] = qnan; sz[j] = qnan;  // NaN never wins `d < best`
        }
      }
      __syncthreads();
#pragma unroll 2
      for (int j = 0; j < cnt4; j += 4) {
        const ulonglong2 X = *reinterpret_cast<const ulonglong2*>(&sx[j]);
        const ulonglong2 Y = *reinterpret_cast<const ulonglong2*>(&sy[j]);
        const ulonglong2 Z = *reinterpret_cast<const ulonglong2*>(&sz[j]);
        // q + (-t) == q - t exactly; then t = dy*dy; fma(dx,dx,t); fma(dz,dz,t)
        unsigned long long dx = add2(qx2, X.x), dy = add2(qy2, Y.x), dz = add2(qz2, Z.x);
        unsigned long long d01 = fma2(dz, dz, fma2(dx, dx, mul2(dy, dy)));
        dx = add2(qx2, X.y); dy = add2(qy2, Y.y); dz = add2(qz2, Z.y);
        unsigned long long d23 = fma2(dz, dz, fma2(dx, dx, mul2(dy, dy)));
        const float2 a = u64_as_f2(d01), c = u64_as_f2(d23);
        const int jj = t0 + j;
        if (a.x < best) { best = a.x; bidx = jj; }
        if (a.y < best) { best = a.y; bidx = jj + 1; }
        if (c.x < best) { best = c.x; bidx = jj + 2; }
        if (c.y < best) { best = c.y; bidx = jj + 3; }
      }
    }
    if (qi < nq) {
      const long long o = (long long)b * nq + qi;
      dout[o] = best;
      if (iout != nullptr) iout[o] = (IdxT)bidx;
    }
  }
}

// ======================================================================
// GRID
// ======================================================================
struct GridParams {
  float ox, oy, oz;  // origin (bbox min)
  float h, inv_h;    // cell size
  int dx, dy, dz;    // cells per axis
  int count;         // points in the grid (valid points of the segment)
  int nfar;          // padded parts kept as single far candidates (shape mode)
  int pad0, pad1;
};

// One cloud of a Chamfer call: S segments (clouds) of Nseg points each.
struct CloudDesc {
  const float* pts;     // [S, Nseg, 3] local-frame points
  const float* quat;    // [S*Nseg/ppp, 4] or nullptr (no pose)
  const float* trans;   // [S*Nseg/ppp, 3] or nullptr
  const float* valids;  // [S*Nseg/ppp] or nullptr (all valid)
  float* out_pts;       // [S, Nseg, 3] transformed points in input order, or nullptr
  int Nseg;             // points per segment
  int ppp;              // points per pose (= N of a part)
  int fill_invalid;     // shape mode: padded parts are the point (1e3,1e3,1e3) (loss.py:175)
  int dmax;             // max cells per axis
  float occ;            // target points per cell (in the effective volume)
  int use_eff;          // size cells from the effective (sigma) volume, not the bbox
};

constexpr int GRID_MAX_DIM = 38;  // 38^3+1 ints = 214 KB of shared memory in the build kernel
constexpr int MAX_FAR = 64;  // >= max parts per shape

// Occupancy pyramid over the cell grid (second stage of the search): level L >= 1 holds, per
// node of (2^L)^3 cells, the number of targets inside.  Level dimensions follow the host-side
// cell budget `dmax` (D[L] = ceil(dmax / 2^L) per axis), not the actual grid, so that every
// cloud of a launch shares one layout; nodes beyond the actual grid hold 0.
constexpr int PYR_MAX_LEVELS = 8;
struct PyrLayout {
  int top;                  // root level: D[top] == 1
  int D[PYR_MAX_LEVELS];    // nodes per axis at level L (D[0] = dmax, unused)
  int off[PYR_MAX_LEVELS];  // offset of level L in a cloud's pyramid (off[0] unused)
  int stride;               // ints per cloud
};
static PyrLayout make_pyr_layout(int dmax) {
  PyrLayout p;
  p.D[0] = dmax;
  p.off[0] = 0;
  int o = 0, L = 0;
  while (p.D[L] > 1 && L + 1 < PYR_MAX_LEVELS) {
    ++L;
    p.D[L] = (p.D[L - 1] + 1) / 2;
    p.off[L] = o;
    o += p.D[L] * p.D[L] * p.D[L];
  }
  if (L == 0) {  // a single cell: still give the search a root above it
    L = 1;
    p.D[1] = 1;
    p.off[1] = 0;
    o = 1;
  }
  p.top = L;
  for (int l = L + 1; l < PYR_MAX_LEVELS; ++l) { p.D[l] = 1; p.off[l] = 0; }
  p.stride = o;
  return p;
}

__device__ __forceinline__ int cell_coord(float x, float o, float inv_h, int dim) {
  int c = (int)floorf((x - o) * inv_h);
  return min(max(c, 0), dim - 1);
}

// load + (optionally) pose-transform point i of segment seg; reports validity
// (`cached`: the transformed point was already written to c.out_pts by pass 1 -- passes 2 and 3
// re-read 12 bytes instead of repeating the quaternion product)
__device__ __forceinline__ float3 load_point(const CloudDesc& c, int seg, int i, bool& valid,
                                             bool cached = false) {
  const long long g = (long long)seg * c.Nseg + i;
  // Nseg is a multiple of ppp (whole parts per segment): 32-bit division
  const long long part = (long long)seg * (c.Nseg / c.ppp) + (unsigned)i / (unsigned)c.ppp;
  valid = (c.valids == nullptr) || (c.valids[part] != 0.0f);
  float3 v;
  if (cached && c.out_pts != nullptr) {
    const float* p = c.out_pts + g * 3;
    return make_float3(p[0], p[1], p[2]);
  }
  if (!valid && c.fill_invalid) {
    v = make_float3(1e3f, 1e3f, 1e3f);
  } else {
    const float* p = c.pts + g * 3;
    v = make_float3(p[0], p[1], p[2]);
  }
  if (c.quat != nullptr) {
    const float* qp = c.quat + part * 4;
    const float q[4] = {qp[0], qp[1], qp[2], qp[3]};
    v = se3_apply(q, c.trans ? c.trans + part * 3 : nullptr, v);
  }
  return v;
}

__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// K1: one CTA per (cloud, segment): bbox -> grid params -> histogram -> scan -> scatter.
// dynamic smem: int cnt[dmax^3 + 1]
template <typename IdxT>
__global__ void grid_build_kernel(CloudDesc c0, CloudDesc c1, int S, float4* __restrict__ sorted0,
                                  float4* __restrict__ sorted1, int* __restrict__ cell_start,
                                  int cs_stride, GridParams* __restrict__ params,
                                  float4* __restrict__ far, int* pyr_base, PyrLayout pl,
                                  int* hard_count, float* dist0, IdxT* idx0, float* dist1,
                                  IdxT* idx1) {
  extern __shared__ int cnt[];
  __shared__ float red[6][32];
  __shared__ double redm[6][32];
  __shared__ int s_wsum[32];
  __shared__ GridParams gp;
  __shared__ int s_count, s_nfar;

  const int cloud = blockIdx.x / S;
  const int seg = blockIdx.x % S;
  const CloudDesc& c = cloud ? c1 : c0;
  float4* sorted = (cloud ? sorted1 : sorted0) + (long long)seg * c.Nseg;
  float* dist_out = cloud ? dist1 : dist0;
  IdxT* idx_out = cloud ? idx1 : idx0;
  int* cs = cell_start + (long long)(cloud * S + seg) * cs_stride;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, nwarps = blockDim.x >> 5;
  const int n = c.Nseg;
  const bool has_valid = c.valids != nullptr;

  if (tid == 0) { s_count = 0; s_nfar = 0; }
  if (blockIdx.x == 0 && tid == 0) { hard_count[0] = 0; hard_count[1] = 0; }  // append list + its read cursor
  __syncthreads();

  // ---- pass 1: bbox of the valid points, transformed output, zero-fill ----
  const float inf = __int_as_float(0x7f800000);
  float mnx = inf, mny = inf, mnz = inf, mxx = -inf, mxy = -inf, mxz = -inf;
  double m1x = 0, m1y = 0, m1z = 0, m2x = 0, m2y = 0, m2z = 0;  // coordinate moments
  int nvalid = 0;
  for (int i = tid; i < n; i += blockDim.x) {
    bool valid;
    const float3 v = load_point(c, seg, i, valid);
    if (c.out_pts != nullptr) {
      float* o = c.out_pts + ((long long)seg * n + i) * 3;
      o[0] = v.x; o[1] = v.y; o[2] = v.z;
    }
    if (valid) {
      mnx = fminf(mnx, v.x); mny = fminf(mny, v.y); mnz = fminf(mnz, v.z);
      mxx = fmaxf(mxx, v.x); mxy = fmaxf(mxy, v.y); mxz = fmaxf(mxz, v.z);
      m1x += v.x; m1y += v.y; m1z += v.z;
      m2x += (double)v.x * v.x; m2y += (double)v.y * v.y; m2z += (double)v.z * v.z;
      ++nvalid;
    } else if (has_valid) {
      const long long o = (long long)seg * n + i;
      if (dist_out != nullptr) dist_out[o] = 0.0f;
      if (idx_out != nullptr) idx_out[o] = (IdxT)-1;
    }
  }
  mnx = warp_min(mnx); mny = warp_min(mny); mnz = warp_min(mnz);
  mxx = warp_max(mxx); mxy = warp_max(mxy); mxz = warp_max(mxz);
  nvalid = __reduce_add_sync(0xffffffffu, nvalid);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    m1x += __shfl_xor_sync(0xffffffffu, m1x, o); m1y += __shfl_xor_sync(0xffffffffu, m1y, o);
    m1z += __shfl_xor_sync(0xffffffffu, m1z, o); m2x += __shfl_xor_sync(0xffffffffu, m2x, o);
    m2y += __shfl_xor_sync(0xffffffffu, m2y, o); m2z += __shfl_xor_sync(0xffffffffu, m2z, o);
  }
  if (lane == 0) {
    red[0][wid] = mnx; red[1][wid] = mny; red[2][wid] = mnz;
    red[3][wid] = mxx; red[4][wid] = mxy; red[5][wid] = mxz;
    redm[0][wid] = m1x; redm[1][wid] = m1y; redm[2][wid] = m1z;
    redm[3][wid] = m2x; redm[4][wid] = m2y; redm[5][wid] = m2z;
    atomicAdd(&s_count, nvalid);
  }
  // padded parts of a shape cloud: one far candidate each, index = first point
  if (c.fill_invalid && has_valid) {
    const int parts = n / c.ppp;
    for (int p = tid; p < parts; p += blockDim.x) {
      bool valid;
      const float3 v = load_point(c, seg, p * c.ppp, valid);
      if (!valid) {
        const int slot = atomicAdd(&s_nfar, 1);
        if (slot < MAX_FAR)
          far[(long long)(cloud * S + seg) * MAX_FAR + slot] =
              make_float4(v.x, v.y, v.z, __int_as_float(p * c.ppp));
      }
    }
  }
  __syncthreads();
  if (wid == 0) {
    float a = lane < nwarps ? red[0][lane] : inf, b = lane < nwarps ? red[1][lane] : inf,
          d = lane < nwarps ? red[2][lane] : inf, e = lane < nwarps ? red[3][lane] : -inf,
          f = lane < nwarps ? red[4][lane] : -inf, g = lane < nwarps ? red[5][lane] : -inf;
    a = warp_min(a); b = warp_min(b); d = warp_min(d);
    e = warp_max(e); f = warp_max(f); g = warp_max(g);
    if (lane == 0) {
      GridParams p;
      p.count = s_count;
      p.nfar = min(s_nfar, MAX_FAR);
      p.pad0 = p.pad1 = 0;
      if (p.count == 0) {  // empty (fully padded) segment
        p.ox = p.oy = p.oz = 0.f; p.h = p.inv_h = 1.f; p.dx = p.dy = p.dz = 1;
      } else {
        const float ex = e - a, ey = f - b, ez = g - d;
        const float emax = fmaxf(ex, fmaxf(ey, ez));
        const bool finite = (emax < inf) && (emax == emax);
        const float tiny = fmaxf(emax * 1e-3f, 1e-30f);
        // Cell size from the EFFECTIVE volume: per axis min(bbox extent, sqrt(12)*sigma)
        // (equal for a uniform box, much smaller for a dense core with a sparse halo),
        // aiming at ~1.5 points per cell where the points actually are.
        float eff[3] = {ex, ey, ez};
        for (int ax = 0; ax < 3; ++ax) {
          double s1 = 0, s2 = 0;
          for (int w = 0; w < nwarps; ++w) { s1 += redm[ax][w]; s2 += redm[3 + ax][w]; }
          const double mean = s1 / p.count;
          const double var = fmax(s2 / p.count - mean * mean, 0.0);
          const float se = (float)(3.4641016151377544 * sqrt(var));
          if (c.use_eff && se == se && se < eff[ax]) eff[ax] = se;
        }
        float h = cbrtf(fmaxf(eff[0], tiny) * fmaxf(eff[1], tiny) * fmaxf(eff[2], tiny) * c.occ /
                        (float)p.count);
        h = fmaxf(h, emax / (float)c.dmax * 1.0001f);
        if (!(h > 0.0f) || !(h < inf) || !finite) h = 1.0f;  // all points equal / non-finite cloud
        p.h = h;
        p.inv_h = 1.0f / h;
        p.ox = a; p.oy = b; p.oz = d;
        p.dx = finite ? max(1, min(c.dmax, (int)floorf(ex * p.inv_h) + 1)) : 1;
        p.dy = finite ? max(1, min(c.dmax, (int)floorf(ey * p.inv_h) + 1)) : 1;
        p.dz = finite ? max(1, min(c.dmax, (int)floorf(ez * p.inv_h) + 1)) : 1;
      }
      gp = p;
      params[cloud * S + seg] = p;
    }
  }
  __syncthreads();
  const GridParams g = gp;
  const int ncell = g.dx * g.dy * g.dz;
  for (int k = tid; k <= ncell; k += blockDim.x) cnt[k] = 0;
  __syncthreads();

  // ---- pass 2: histogram ----
  for (int i = tid; i < n; i += blockDim.x) {
    bool valid;
    const float3 v = load_point(c, seg, i, valid, true);
    if (valid) {
      const int cell = (cell_coord(v.z, g.oz, g.inv_h, g.dz) * g.dy +
                        cell_coord(v.y, g.oy, g.inv_h, g.dy)) * g.dx +
                       cell_coord(v.x, g.ox, g.inv_h, g.dx);
      atomicAdd(&cnt[cell], 1);
    }
  }
  __syncthreads();

  // ---- exclusive scan of cnt[0..ncell) in place; cnt[ncell] = total ----
  {
    const int per = (ncell + blockDim.x - 1) / blockDim.x;
    const int beg = min(tid * per, ncell), end = min(beg + per, ncell);
    int sum = 0;
    for (int k = beg; k < end; ++k) sum += cnt[k];
    int incl = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    if (lane == 31) s_wsum[wid] = incl;
    __syncthreads();
    if (wid == 0) {
      int v = lane < nwarps ? s_wsum[lane] : 0;
      int iv = v;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, iv, o);
        if (lane >= o) iv += t;
      }
      s_wsum[lane] = iv - v;  // exclusive warp offsets
    }
    __syncthreads();
    int off = incl - sum + s_wsum[wid];
    for (int k = beg; k < end; ++k) {
      const int v = cnt[k];
      cnt[k] = off;
      off += v;
    }
    if (tid == 0) cnt[ncell] = g.count;
  }
  __syncthreads();
  for (int k = tid; k <= ncell; k += blockDim.x) cs[k] = cnt[k];
  __syncthreads();

  // ---- pass 3: scatter (cnt[] doubles as the per-cell cursor) ----
  for (int i = tid; i < n; i += blockDim.x) {
    bool valid;
    const float3 v = load_point(c, seg, i, valid, true);
    if (valid) {
      const int cell = (cell_coord(v.z, g.oz, g.inv_h, g.dz) * g.dy +
                        cell_coord(v.y, g.oy, g.inv_h, g.dy)) * g.dx +
                       cell_coord(v.x, g.ox, g.inv_h, g.dx);
      const int pos = atomicAdd(&cnt[cell], 1);
      sorted[pos] = make_float4(v.x, v.y, v.z, __int_as_float(i));
    }
  }
  __syncthreads();

  // ---- occupancy pyramid of child masks (cnt[] now holds the cell ENDS): level 1 from the cells,
  // each further level from the one below (global memory written by this CTA, made visible to
  // it by __syncthreads) ----
  {
    int* pyr = pyr_base + (long long)(cloud * S + seg) * pl.stride;
    {
      // level 1: bit b = dx + 2 dy + 4 dz of a node's entry says that child cell (2x+dx, 2y+dy,
      // 2z+dz) holds points.  An entry is non-zero iff the node holds points.
      const int D = pl.D[1];
      int* lv = pyr + pl.off[1];
      for (int i = tid; i < D * D * D; i += blockDim.x) {
        const int x = i % D, y = (i / D) % D, z = i / (D * D);
        int mask = 0;
#pragma unroll
        for (int b = 0; b < 8; ++b) {
          const int xc = 2 * x + (b & 1), yc = 2 * y + ((b >> 1) & 1), zc = 2 * z + (b >> 2);
          if (xc < g.dx && yc < g.dy && zc < g.dz) {
            const int ci = (zc * g.dy + yc) * g.dx + xc;
            const int n_in = cnt[ci] - (ci == 0 ? 0 : cnt[ci - 1]);  // cnt[] holds the cell ends
            mask |= (n_in > 0) ? (1 << b) : 0;
          }
        }
        lv[i] = mask;
      }
    }
    for (int L = 2; L <= pl.top; ++L) {
      __syncthreads();
      const int D = pl.D[L], Dc = pl.D[L - 1];
      const int* lc = pyr + pl.off[L - 1];
      int* lv = pyr + pl.off[L];
      for (int i = tid; i < D * D * D; i += blockDim.x) {
        const int x = i % D, y = (i / D) % D, z = i / (D * D);
        int mask = 0;  // bit b: child node b of the level below is non-empty
#pragma unroll
        for (int b = 0; b < 8; ++b) {
          const int xc = 2 * x + (b & 1), yc = 2 * y + ((b >> 1) & 1), zc = 2 * z + (b >> 2);
          if (xc < Dc && yc < Dc && zc < Dc) mask |= (lc[(zc * Dc + yc) * Dc + xc] != 0) ? (1 << b) : 0;
        }
        lv[i] = mask;
      }
    }
  }
}

__device__ __forceinline__ void consider(const float4 t, float qx, float qy, float qz, float& best,
                                         int& bidx) {
  const float d = sqdist_ref(qx, qy, qz, t.x, t.y, t.z);
  if (d <= best) {  // rare after the first few candidates
    const int ti = __float_as_int(t.w);
    if (d < best || ti < bidx) { best = d; bidx = ti; }
  }
}

__device__ __forceinline__ void scan_range(const float4* __restrict__ T, int s, int e, float qx,
                                           float qy, float qz, float& best, int& bidx) {
  int p = s;
  for (; p + 3 < e; p += 4) {  // four independent loads in flight
    const float4 t0 = __ldg(&T[p]);
    const float4 t1 = __ldg(&T[p + 1]);
    const float4 t2 = __ldg(&T[p + 2]);
    const float4 t3 = __ldg(&T[p + 3]);
    consider(t0, qx, qy, qz, best, bidx);
    consider(t1, qx, qy, qz, best, bidx);
    consider(t2, qx, qy, qz, best, bidx);
    consider(t3, qx, qy, qz, best, bidx);
  }
  if (p + 1 < e) {
    const float4 t0 = __ldg(&T[p]);
    const float4 t1 = __ldg(&T[p + 1]);
    consider(t0, qx, qy, qz, best, bidx);
    consider(t1, qx, qy, qz, best, bidx);
    p += 2;
  }
  if (p < e) consider(__ldg(&T[p]), qx, qy, qz, best, bidx);
}

// K2: one lane per query (queries taken in their own cloud's cell order so a
// warp touches neighbouring target cells).
//   A  scan the query's own cell -> a first `best`;
//   B  (convergent, no scanning) for the other 26 cells of the 3x3x3 block decide
//      per row (fixed y,z; contiguous along x) which cells the sphere of radius
//      sqrt(best) can reach, from the squared distances to the own cell's faces,
//      and note the candidate ranges (<= 10 per lane) in shared memory;
//   C  one flat loop over the concatenated ranges: lanes stay converged on the
//      distance evaluation instead of idling while others walk different rows;
//   D  the block [c-1, c+1]^3 is proven complete when `best` is smaller than the
//      distance to the block's faces that have cells beyond them;
//   E  otherwise (sparse regions, queries outside the target cloud) a second grid
//      level of 4x4x4-cell blocks is searched: a converged loop over one representative
//      point per occupied block bounds `best`, the nearest block is opened first, then
//      every block (slab, row, cell range) that is not strictly farther than `best`.
// A cell is skipped only if it is STRICTLY farther than `best` (after a fp32
// slack), so neither a closer point nor a tie with a lower index can hide in it.
constexpr int NN_THREADS = 256;
constexpr int NN_RANGES = 10;

// everything a lane needs to search the target grid for one query
struct NNQuery {
  float4 q;                       // x, y, z, original index (bits)
  GridParams g;                   // target grid
  const float4* __restrict__ T;   // targets in cell order
  const int* __restrict__ cs;     // cell starts
  const int* __restrict__ pyr;    // occupancy pyramid of the target grid
  const float4* __restrict__ F;   // far candidates (padded parts)
  long long out;                  // output slot
  int cx, cy, cz;
  float slack;
  bool dir1, valid;
};

__device__ __forceinline__ NNQuery nn_setup(long long gid, const float4* __restrict__ sorted0,
                                            const float4* __restrict__ sorted1,
                                            const int* __restrict__ cell_start, int cs_stride,
                                            const GridParams* __restrict__ params,
                                            const float4* __restrict__ far,
                                            const int* __restrict__ pyr_base, int pyr_stride, int S,
                                            int N0, int N1) {
  NNQuery c;
  c.valid = false;
  const long long wg = gid >> 5;
  const int lane = (int)(gid & 31);
  const int per0 = (N0 + 31) >> 5, per1 = (N1 + 31) >> 5;
  const long long total0 = (long long)S * per0;
  const long long total = total0 + (long long)S * per1;
  if (wg >= total) return c;
  c.dir1 = wg >= total0;  // queries from cloud 1, targets cloud 0
  const long long ww = c.dir1 ? wg - total0 : wg;
  const int per = c.dir1 ? per1 : per0;
  const int seg = (int)(ww / per);
  const int i = (int)(ww % per) * 32 + lane;
  const int qc = c.dir1 ? 1 : 0, tc = c.dir1 ? 0 : 1;
  const int NQ = c.dir1 ? N1 : N0, NT = c.dir1 ? N0 : N1;
  if (i >= params[qc * S + seg].count) return c;
  c.valid = true;
  c.g = params[tc * S + seg];
  c.q = (c.dir1 ? sorted1 : sorted0)[(long long)seg * NQ + i];
  c.T = (c.dir1 ? sorted0 : sorted1) + (long long)seg * NT;
  c.cs = cell_start + (long long)(tc * S + seg) * cs_stride;
  c.pyr = pyr_base + (long long)(tc * S + seg) * pyr_stride;
  c.F = far + (long long)(tc * S + seg) * MAX_FAR;
  c.out = (long long)seg * NQ + __float_as_int(c.q.w);
  const GridParams& g = c.g;
  c.cx = cell_coord(c.q.x, g.ox, g.inv_h, g.dx);
  c.cy = cell_coord(c.q.y, g.oy, g.inv_h, g.dy);
  c.cz = cell_coord(c.q.z, g.oz, g.inv_h, g.dz);
  // positional uncertainty of cell planes / cell assignment in fp32
  c.slack = 1e-5f * (fabsf(c.q.x) + fabsf(c.q.y) + fabsf(c.q.z) + fabsf(g.ox) + fabsf(g.oy) +
                     fabsf(g.oz) + (float)(g.dx + g.dy + g.dz) * g.h);
  return c;
}

// phase E: best-first descent of the occupancy pyramid (see the kernel comment).  Depth-first
// with an explicit stack (local memory, <= 7 pending siblings per level); the children of a
// node are pushed farthest first in the XOR order of the child nearest to the query, so the
// nearest one is opened next.  A node is dropped only if the squared distance to its box
// (after the fp32 slack) is STRICTLY larger than `best`, or if it holds no target.
constexpr int PYR_STACK = 7 * (PYR_MAX_LEVELS - 1) + 2;

// squared gap between q and the slab [cc << L, min((cc + 1) << L, dim)) of cells along one axis
__device__ __forceinline__ float axis_gap2(float q, float o, float h, int cc, int L, int dim,
                                           float slack) {
  const int c0 = cc << L;
  if (c0 >= dim) return __int_as_float(0x7f800000);  // outside the grid: never opened
  const float lo = o + (float)c0 * h, hi = o + (float)min((cc + 1) << L, dim) * h;
  const float gpos = fmaxf(fmaxf(lo - q, q - hi) - slack, 0.f);
  return gpos * gpos;
}

// one stack entry: scan a cell or expand a node (pushes <= 8 children)
template <bool COUNT>
__device__ __forceinline__ void nn_pyramid_step(const NNQuery& c, const PyrLayout& pl, int* stack, int& sp,
                                                float& best, int& bidx, unsigned& ncand) {
  const GridParams& g = c.g;
  const float4 q = c.q;
  const int* __restrict__ cs = c.cs;
  const float slack = c.slack;
  const int e = stack[--sp];
  const int L = e >> 18, z = (e >> 12) & 63, y = (e >> 6) & 63, x = e & 63;
  if (L == 0) {  // a cell; those of the 3x3x3 block were handled (scanned or pruned) in A-C
    if (x >= c.cx - 1 && x <= c.cx + 1 && y >= c.cy - 1 && y <= c.cy + 1 && z >= c.cz - 1 && z <= c.cz + 1)
      return;
    const float lb = axis_gap2(q.x, g.ox, g.h, x, 0, g.dx, slack) + axis_gap2(q.y, g.oy, g.h, y, 0, g.dy, slack) +
                     axis_gap2(q.z, g.oz, g.h, z, 0, g.dz, slack);
    if (lb > best) return;  // `best` may have improved since the push
    const int ci = (z * g.dy + y) * g.dx + x;
    const int s0 = cs[ci], s1 = cs[ci + 1];
    if (COUNT) ncand += (unsigned)(s1 - s0);
    scan_range(c.T, s0, s1, q.x, q.y, q.z, best, bidx);
    return;
  }
  const int Lc = L - 1;
  // squared gaps to the two halves of the node along every axis
  const float gx0 = axis_gap2(q.x, g.ox, g.h, 2 * x, Lc, g.dx, slack);
  const float gx1 = axis_gap2(q.x, g.ox, g.h, 2 * x + 1, Lc, g.dx, slack);
  const float gy0 = axis_gap2(q.y, g.oy, g.h, 2 * y, Lc, g.dy, slack);
  const float gy1 = axis_gap2(q.y, g.oy, g.h, 2 * y + 1, Lc, g.dy, slack);
  const float gz0 = axis_gap2(q.z, g.oz, g.h, 2 * z, Lc, g.dz, slack);
  const float gz1 = axis_gap2(q.z, g.oz, g.h, 2 * z + 1, Lc, g.dz, slack);
  if (fminf(gx0, gx1) + fminf(gy0, gy1) + fminf(gz0, gz1) > best) return;  // whole node out of reach
  const int near = (gx1 < gx0 ? 1 : 0) | (gy1 < gy0 ? 2 : 0) | (gz1 < gz0 ? 4 : 0);
  // ONE load says which children hold points (it was requested when this node was pushed)
  const int D = pl.D[L];
  const unsigned kids = (unsigned)(c.pyr + pl.off[L])[(z * D + y) * D + x];
  const int Dc = pl.D[Lc];
  const int* __restrict__ lv = c.pyr + pl.off[Lc];
#pragma unroll
  for (int o = 7; o >= 0; --o) {  // farthest first: the nearest child ends on top of the stack
    const int b = o ^ near;
    if (((kids >> b) & 1u) == 0u) continue;  // empty (or outside the grid)
    const int xc = 2 * x + (b & 1), yc = 2 * y + ((b >> 1) & 1), zc = 2 * z + (b >> 2);
    const float lb = ((b & 1) ? gx1 : gx0) + ((b & 2) ? gy1 : gy0) + ((b & 4) ? gz1 : gz0);
    if (lb > best) continue;
    stack[sp++] = (Lc << 18) | (zc << 12) | (yc << 6) | xc;
    // what the pop of this entry will read first: its own child mask, or the cell's range
    const int* nextp = Lc == 0 ? cs + ((zc * g.dy + yc) * g.dx + xc) : lv + ((zc * Dc + yc) * Dc + xc);
    asm volatile("prefetch.global.L1 [%0];" ::"l"(nextp));
  }
}

template <typename IdxT>
__device__ __forceinline__ void nn_finish(const NNQuery& c, float best, int bidx, float* __restrict__ dist0,
                                          IdxT* __restrict__ idx0, float* __restrict__ dist1,
                                          IdxT* __restrict__ idx1) {
  for (int f = 0; f < c.g.nfar; ++f) consider(c.F[f], c.q.x, c.q.y, c.q.z, best, bidx);
  (c.dir1 ? dist1 : dist0)[c.out] = best;
  IdxT* io = c.dir1 ? idx1 : idx0;
  if (io != nullptr) io[c.out] = (IdxT)bidx;
}

// Kernel 1: every lane runs A-D for its own query.  Queries that need phase E are appended
// to a global list (one atomic per warp, no CTA barrier: the CTA leaves as soon as its
// slowest easy query is done) and kernel 2 (`grid_nn_hard_kernel`) deals them out densely,
// 32 per warp, to a persistent grid -- the long, divergent pyramid descent neither idles the
// lanes of easy queries nor holds their CTA slots.
#ifndef MPA_NN_MIN_CTAS
#define MPA_NN_MIN_CTAS 4
#endif
template <typename IdxT, bool COUNT>
__global__ void __launch_bounds__(NN_THREADS, MPA_NN_MIN_CTAS)
grid_nn_kernel(const float4* __restrict__ sorted0, const float4* __restrict__ sorted1,
               const int* __restrict__ cell_start, int cs_stride,
               const GridParams* __restrict__ params, const float4* __restrict__ far,
               const int* __restrict__ pyr_base, const PyrLayout pl, int S,
               int N0, int N1, float* __restrict__ dist0, IdxT* __restrict__ idx0,
               float* __restrict__ dist1, IdxT* __restrict__ idx1,
               int* __restrict__ hard_count, int4* __restrict__ hard_list,
               unsigned long long* __restrict__ pair_counter) {
  __shared__ int rs[NN_RANGES][NN_THREADS], re[NN_RANGES][NN_THREADS];
  __shared__ float rl[NN_RANGES][NN_THREADS];  // squared distance to each noted range's slab
  const int tid = threadIdx.x;
  const long long gid0 = (long long)blockIdx.x * NN_THREADS;
  unsigned ncand = 0;  // candidate pairs this lane evaluated (COUNT instantiation only)
  bool hard = false;
  float hbest = 0.f;
  int hbidx = -1;
  {
    const NNQuery c = nn_setup(gid0 + tid, sorted0, sorted1, cell_start, cs_stride, params, far, pyr_base,
                               pl.stride, S, N0, N1);
    if (c.valid) {
      const GridParams& g = c.g;
      const float4 q = c.q;
      const float4* __restrict__ T = c.T;
      const int* __restrict__ cs = c.cs;
      const int cx = c.cx, cy = c.cy, cz = c.cz;
      const float slack = c.slack;
      const float inf = __int_as_float(0x7f800000);
      float best = 1e32f;
      int bidx = -1;
      bool done = true;
      if (g.count > 0) {
        // squared distances from the query to the faces of its own cell (0 = lower side)
        float X0 = fmaxf(q.x - (g.ox + (float)cx * g.h) - slack, 0.f);
        float X1 = fmaxf((g.ox + (float)(cx + 1) * g.h) - q.x - slack, 0.f);
        float Y0 = fmaxf(q.y - (g.oy + (float)cy * g.h) - slack, 0.f);
        float Y1 = fmaxf((g.oy + (float)(cy + 1) * g.h) - q.y - slack, 0.f);
        float Z0 = fmaxf(q.z - (g.oz + (float)cz * g.h) - slack, 0.f);
        float Z1 = fmaxf((g.oz + (float)(cz + 1) * g.h) - q.z - slack, 0.f);
        X0 *= X0; X1 *= X1; Y0 *= Y0; Y1 *= Y1; Z0 *= Z0; Z1 *= Z1;
        const bool hasL = cx > 0, hasR = cx < g.dx - 1;

        // ---- A: own cell ----
        const int c0 = (cz * g.dy + cy) * g.dx + cx;
        if (COUNT) ncand += (unsigned)(cs[c0 + 1] - cs[c0]);
        scan_range(T, cs[c0], cs[c0 + 1], q.x, q.y, q.z, best, bidx);
        // ---- B: which other cells of the 3x3x3 block can still matter (nearest rows first) ----
        int nr = 0;
        if (hasL && X0 <= best) { rs[nr][tid] = cs[c0 - 1]; re[nr][tid] = cs[c0]; rl[nr][tid] = X0; ++nr; }
        if (hasR && X1 <= best) { rs[nr][tid] = cs[c0 + 1]; re[nr][tid] = cs[c0 + 2]; rl[nr][tid] = X1; ++nr; }
    #pragma unroll
        for (int o = 0; o < 8; ++o) {
          // face rows (one of dy, dz zero) before the four edge rows
          const int dy = o < 2 ? (o == 0 ? -1 : 1) : (o < 4 ? 0 : ((o & 1) ? 1 : -1));
          const int dz = o < 2 ? 0 : (o < 4 ? (o == 2 ? -1 : 1) : (o < 6 ? -1 : 1));
          const int zc = cz + dz, yc = cy + dy;
          if (zc < 0 || zc >= g.dz || yc < 0 || yc >= g.dy) continue;
          const float l2 = (dy < 0 ? Y0 : (dy > 0 ? Y1 : 0.f)) + (dz < 0 ? Z0 : (dz > 0 ? Z1 : 0.f));
          if (l2 > best) continue;
          const int xa = cx - ((hasL && X0 + l2 <= best) ? 1 : 0);
          const int xb = cx + ((hasR && X1 + l2 <= best) ? 1 : 0);
          const int row = (zc * g.dy + yc) * g.dx;
          rs[nr][tid] = cs[row + xa];
          re[nr][tid] = cs[row + xb + 1];
          rl[nr][tid] = l2;
          ++nr;
        }
        // ---- C: flat scan of the noted ranges (two candidates in flight); a range whose
        // slab has meanwhile become strictly farther than `best` is dropped unscanned ----
        {
          int k = -1, p = 0, e = 0;
          while (true) {
            if (p >= e) {
              if (++k >= nr) break;
              if (rl[k][tid] > best) continue;
              p = rs[k][tid];
              e = re[k][tid];
              if (COUNT) ncand += (unsigned)(e - p);
              continue;
            }
            const float4 t0 = __ldg(T + p);
            if (p + 3 < e) {
              const float4 t1 = __ldg(T + p + 1);
              const float4 t2 = __ldg(T + p + 2);
              const float4 t3 = __ldg(T + p + 3);
              consider(t0, q.x, q.y, q.z, best, bidx);
              consider(t1, q.x, q.y, q.z, best, bidx);
              consider(t2, q.x, q.y, q.z, best, bidx);
              consider(t3, q.x, q.y, q.z, best, bidx);
              p += 4;
            } else if (p + 1 < e) {
              const float4 t1 = __ldg(T + p + 1);
              consider(t0, q.x, q.y, q.z, best, bidx);
              consider(t1, q.x, q.y, q.z, best, bidx);
              p += 2;
            } else {
              consider(t0, q.x, q.y, q.z, best, bidx);
              ++p;
            }
          }
        }
        // ---- D: is the 3x3x3 block provably complete? ----
        done = (cx - 1 <= 0) && (cx + 1 >= g.dx - 1) && (cy - 1 <= 0) && (cy + 1 >= g.dy - 1) &&
                    (cz - 1 <= 0) && (cz + 1 >= g.dz - 1);
        if (!done) {
          float bound = inf;
          if (cx - 1 > 0) bound = fminf(bound, q.x - (g.ox + (float)(cx - 1) * g.h));
          if (cx + 1 < g.dx - 1) bound = fminf(bound, (g.ox + (float)(cx + 2) * g.h) - q.x);
          if (cy - 1 > 0) bound = fminf(bound, q.y - (g.oy + (float)(cy - 1) * g.h));
          if (cy + 1 < g.dy - 1) bound = fminf(bound, (g.oy + (float)(cy + 2) * g.h) - q.y);
          if (cz - 1 > 0) bound = fminf(bound, q.z - (g.oz + (float)(cz - 1) * g.h));
          if (cz + 1 < g.dz - 1) bound = fminf(bound, (g.oz + (float)(cz + 2) * g.h) - q.z);
          bound = (bound - slack) * 0.99999f;
          // every unexplored target is farther than `bound`: strict so that a tie with
          // a lower index cannot hide outside the block
          done = bound > 0.0f && best < bound * bound;
        }
      }
      if (done) {
        nn_finish<IdxT>(c, best, bidx, dist0, idx0, dist1, idx1);
      } else {
        hard = true;
        hbest = best;
        hbidx = bidx;
      }
    }
  }
  // warp-aggregated append of the unfinished queries (all lanes reconverge here)
  {
    const unsigned m = __ballot_sync(0xffffffffu, hard);
    if (m != 0u) {
      const int lane = tid & 31;
      int base = 0;
      if (lane == 0) base = atomicAdd(hard_count, __popc(m));
      base = __shfl_sync(0xffffffffu, base, 0);
      if (hard)
        hard_list[base + __popc(m & ((1u << lane) - 1u))] =
            make_int4((int)(gid0 + tid), __float_as_int(hbest), hbidx, 0);
    }
  }
  if (COUNT) {
    const unsigned long long w = (unsigned long long)__reduce_add_sync(0xffffffffu, ncand);
    if ((tid & 31) == 0 && w) atomicAdd(pair_counter, w);
  }
}

// Kernel 2: phase E for the queries kernel 1 could not finish.  Persistent lanes: the descent is
// one flat loop whose body handles ONE stack entry (a cell to scan or a node to expand), so the
// lanes of a warp stay converged on it whatever query each works on; a lane whose stack runs
// empty writes its result and waits until a quarter of the warp is idle (or nothing else is
// running), then the idle lanes draw new queries together (one atomic per refill) -- the
// descent lengths differ ~3x between neighbouring queries, which left 2/3 of the lanes idle
// when every warp simply took 32 queries.
constexpr int NNH_THREADS = 128;
constexpr int NNH_REFILL = 8;  // idle lanes that trigger a refill
template <typename IdxT, bool COUNT>
__global__ void __launch_bounds__(NNH_THREADS)
grid_nn_hard_kernel(const float4* __restrict__ sorted0, const float4* __restrict__ sorted1,
                    const int* __restrict__ cell_start, int cs_stride,
                    const GridParams* __restrict__ params, const float4* __restrict__ far,
                    const int* __restrict__ pyr_base, const PyrLayout pl, int S, int N0, int N1,
                    float* __restrict__ dist0, IdxT* __restrict__ idx0, float* __restrict__ dist1,
                    IdxT* __restrict__ idx1, const int* __restrict__ hard_count,
                    int* __restrict__ hard_cursor, const int4* __restrict__ hard_list,
                    unsigned long long* __restrict__ pair_counter) {
  const int n = *hard_count;
  const int lane = threadIdx.x & 31;
  unsigned ncand = 0;
  int stack[PYR_STACK];
  int sp = 0;
  bool have = false;      // this lane holds a query
  bool drained = false;   // the list is exhausted: no more refills
  NNQuery c;
  float best = 0.f;
  int bidx = -1;
  while (true) {
    const unsigned idle = __ballot_sync(0xffffffffu, sp == 0);
    if (idle == 0xffffffffu || (!drained && __popc(idle) >= NNH_REFILL)) {
      // ---- idle lanes retire their query and draw the next ones together ----
      if (sp == 0 && have) {
        nn_finish<IdxT>(c, best, bidx, dist0, idx0, dist1, idx1);
        have = false;
      }
      if (drained) break;  // every stack is empty and nothing is left
      int base = 0;
      if (lane == 0) base = atomicAdd(hard_cursor, __popc(idle));
      base = __shfl_sync(0xffffffffu, base, 0);
      if (base + __popc(idle) >= n) drained = true;  // warp-uniform
      if (sp == 0) {
        const int i = base + __popc(idle & ((1u << lane) - 1u));
        if (i < n) {
          const int4 h = hard_list[i];
          c = nn_setup((long long)h.x, sorted0, sorted1, cell_start, cs_stride, params, far, pyr_base,
                       pl.stride, S, N0, N1);
          best = __int_as_float(h.y);
          bidx = h.z;
          have = true;
          stack[0] = pl.top << 18;  // the root
          sp = 1;
        }
      }
      continue;
    }
    if (sp > 0) nn_pyramid_step<COUNT>(c, pl, stack, sp, best, bidx, ncand);
  }
  if (COUNT && ncand) atomicAdd(pair_counter, (unsigned long long)ncand);
}

// ======================================================================
// BACKWARD
// ======================================================================
// Both directions of ChamferBackwardKernel (chamfer_kernel.cu:175-210) in one
// launch.  grad buffers are zeroed by the host wrapper first.
template <typename IdxT>
__global__ void chamfer_backward_kernel(const float* __restrict__ g1, const float* __restrict__ g2,
                                        const float* __restrict__ xyz1, const float* __restrict__ xyz2,
                                        const IdxT* __restrict__ idx1, const IdxT* __restrict__ idx2,
                                        int B, int n1, int n2, float* gx1, float* gx2) {
  const long long t1 = (long long)B * n1, total = t1 + (long long)B * n2;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total;
       e += (long long)gridDim.x * blockDim.x) {
    const bool d2 = e >= t1;
    const long long i = d2 ? e - t1 : e;
    const int na = d2 ? n2 : n1, nb = d2 ? n1 : n2;
    const float* A = d2 ? xyz2 : xyz1;
    const float* Bp = d2 ? xyz1 : xyz2;
    float* GA = d2 ? gx2 : gx1;
    float* GB = d2 ? gx1 : gx2;
    const long long j0 = (long long)(d2 ? idx2[i] : idx1[i]);
    const float g = (d2 ? g2[i] : g1[i]) * 2.0f;
    if (j0 < 0 || g == 0.0f) continue;  // skipped (padded) queries of the fused path
    const long long b = i / na;
    const long long j = b * nb + j0;
    const float gx = g * (A[3 * i + 0] - Bp[3 * j + 0]);
    const float gy = g * (A[3 * i + 1] - Bp[3 * j + 1]);
    const float gz = g * (A[3 * i + 2] - Bp[3 * j + 2]);
    atomicAdd(GA + 3 * i + 0, gx); atomicAdd(GA + 3 * i + 1, gy); atomicAdd(GA + 3 * i + 2, gz);
    atomicAdd(GB + 3 * j + 0, -gx); atomicAdd(GB + 3 * j + 1, -gy); atomicAdd(GB + 3 * j + 2, -gz);
  }
}

// ======================================================================
// host side
// ======================================================================
static int num_sms() { return device_sms(); }

static int pick_dmax(int n) {
  // cells per axis the bbox of a uniform cloud needs at ~1 point per cell
  // (the device code aims at c.occ points per cell), capped by shared memory
  static double fine = -1.0;
  if (fine < 0.0) {  // tuning knob: cells-per-axis budget relative to cbrt(n)
    const char* e = getenv("MPA_GRID_FINE");
    fine = e ? atof(e) : 1.0;
  }
  int d = (int)ceil(cbrt((double)n) * fine);
  if (d < 1) d = 1;
  if (d > GRID_MAX_DIM) d = GRID_MAX_DIM;
  return d;
}

// ---- optional instrumentation (bench.py roofline_fp32): candidate pairs evaluated ----
// One device counter per launch kind (0 per-part pose search, 1 shape-level, 2 plain clouds),
// allocated on first use on the current device; enabled by mpa_chamfer_pair_count(1, ...).
static std::atomic<int> g_pair_count_on{0};
static unsigned long long* g_pair_counters[64] = {nullptr};
static bool pair_count_enabled() { return g_pair_count_on.load(std::memory_order_relaxed) != 0; }
static unsigned long long* pair_counter_device(int which) {
  const int d = current_device() & 63;
  if (g_pair_counters[d] == nullptr) {
    if (cudaMalloc((void**)&g_pair_counters[d], 3 * sizeof(unsigned long long)) != cudaSuccess) {
      set_error("pair counter: cudaMalloc failed");
      return nullptr;
    }
    cudaMemset(g_pair_counters[d], 0, 3 * sizeof(unsigned long long));
  }
  return g_pair_counters[d] + which;
}

struct GridLayout {
  size_t off_params, off_far, off_pyr, off_cs, off_sorted0, off_sorted1, off_hard, total;
  int cs_stride;
  PyrLayout pyr;
};

static GridLayout grid_layout(int S, int N0, int N1) {
  GridLayout L;
  const int d = pick_dmax(N0 > N1 ? N0 : N1);
  L.cs_stride = d * d * d + 1;
  size_t o = 0;
  L.off_params = o; o = align_up(o + sizeof(GridParams) * 2 * (size_t)S, 256);
  L.off_far = o; o = align_up(o + sizeof(float4) * 2 * (size_t)S * MAX_FAR, 256);
  L.pyr = make_pyr_layout(d);
  L.off_pyr = o; o = align_up(o + sizeof(int) * 2 * (size_t)S * L.pyr.stride, 256);
  L.off_cs = o; o = align_up(o + sizeof(int) * 2 * (size_t)S * L.cs_stride, 256);
  L.off_sorted0 = o; o = align_up(o + sizeof(float4) * (size_t)S * (N0 > 0 ? N0 : 1), 256);
  L.off_sorted1 = o; o = align_up(o + sizeof(float4) * (size_t)S * (N1 > 0 ? N1 : 1), 256);
  // counter (first 256 bytes) + one int4 per query of either direction (worst case: all unfinished)
  const size_t nq = (size_t)S * ((size_t)((N0 + 31) / 32) + (size_t)((N1 + 31) / 32)) * 32;
  L.off_hard = o; o = align_up(o + 256 + sizeof(int4) * nq, 256);
  L.total = o;
  return L;
}

template <typename IdxT>
static int run_grid(CloudDesc c0, CloudDesc c1, int S, float* dist0, IdxT* idx0, float* dist1,
                    IdxT* idx1, void* ws, size_t ws_bytes, cudaStream_t stream) {
  const GridLayout L = grid_layout(S, c0.Nseg, c1.Nseg);
  Scratch scratch;
  int rc = scratch.acquire(ws, ws_bytes, L.total, stream);
  if (rc != MPA_OK) return rc;
  char* base = (char*)scratch.base;
  GridParams* params = (GridParams*)(base + L.off_params);
  float4* far = (float4*)(base + L.off_far);
  int* pyr = (int*)(base + L.off_pyr);
  int* cs = (int*)(base + L.off_cs);
  float4* s0 = (float4*)(base + L.off_sorted0);
  float4* s1 = (float4*)(base + L.off_sorted1);
  int* hard_count = (int*)(base + L.off_hard);
  int4* hard_list = (int4*)(base + L.off_hard + 256);
  c0.dmax = pick_dmax(c0.Nseg);
  c1.dmax = pick_dmax(c1.Nseg);
  {
    static float occ = -1.f;
    static int use_eff = 1;
    if (occ < 0.f) {  // tuning knobs (defaults chosen from the sweep in profiles/)
      const char* e = getenv("MPA_GRID_OCC");
      occ = e ? (float)atof(e) : 3.0f;
      const char* u = getenv("MPA_GRID_EFF");
      use_eff = u ? atoi(u) : 0;
    }
    c0.occ = c1.occ = occ;
    c0.use_eff = c1.use_eff = use_eff;
  }

  const int nmax = c0.Nseg > c1.Nseg ? c0.Nseg : c1.Nseg;
  const int threads = nmax >= 8192 ? 1024 : (nmax >= 2048 ? 512 : 256);
  const size_t smem = sizeof(int) * (size_t)L.cs_stride;
  static DeviceOnce attr_set[2];
  const int which = sizeof(IdxT) == 8 ? 1 : 0;
  if (attr_set[which].pending()) {
    MPA_CUDA(cudaFuncSetAttribute(grid_build_kernel<IdxT>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)(sizeof(int) * (GRID_MAX_DIM * GRID_MAX_DIM * GRID_MAX_DIM + 1))));
    attr_set[which].done();
  }
  {
    ProfScope ps(c0.fill_invalid ? "chamfer_grid_build_shape" : (c0.quat ? "chamfer_grid_build_part" : "chamfer_grid_build"), stream);
    grid_build_kernel<IdxT><<<2 * S, threads, smem, stream>>>(c0, c1, S, s0, s1, cs, L.cs_stride,
                                                              params, far, pyr, L.pyr, hard_count, dist0,
                                                              idx0, dist1, idx1);
  }
  MPA_LAUNCH_CHECK();
  const long long warps = (long long)S * ((c0.Nseg + 31) / 32 + (c1.Nseg + 31) / 32);
  const long long blocks = (warps + 7) / 8;
  if (blocks > 0) {
    {
      const int which_counter = c0.fill_invalid ? 1 : (c0.quat ? 0 : 2);
      ProfScope ps(c0.fill_invalid ? "chamfer_grid_nn_shape" : (c0.quat ? "chamfer_grid_nn_part" : "chamfer_grid_nn"), stream);
      const int hard_ctas = num_sms() * 8;  // persistent: 8 CTAs of 128 threads per SM
      if (pair_count_enabled()) {
        unsigned long long* ctr = pair_counter_device(which_counter);
        if (ctr == nullptr) return MPA_ERR_CUDA;
        grid_nn_kernel<IdxT, true><<<(unsigned)blocks, 256, 0, stream>>>(
            s0, s1, cs, L.cs_stride, params, far, pyr, L.pyr, S, c0.Nseg, c1.Nseg, dist0, idx0, dist1, idx1,
            hard_count, hard_list, ctr);
        grid_nn_hard_kernel<IdxT, true><<<hard_ctas, NNH_THREADS, 0, stream>>>(
            s0, s1, cs, L.cs_stride, params, far, pyr, L.pyr, S, c0.Nseg, c1.Nseg, dist0, idx0, dist1, idx1,
            hard_count, hard_count + 1, hard_list, ctr);
      } else {
        grid_nn_kernel<IdxT, false><<<(unsigned)blocks, 256, 0, stream>>>(
            s0, s1, cs, L.cs_stride, params, far, pyr, L.pyr, S, c0.Nseg, c1.Nseg, dist0, idx0, dist1, idx1,
            hard_count, hard_list, nullptr);
        grid_nn_hard_kernel<IdxT, false><<<hard_ctas, NNH_THREADS, 0, stream>>>(
            s0, s1, cs, L.cs_stride, params, far, pyr, L.pyr, S, c0.Nseg, c1.Nseg, dist0, idx0, dist1, idx1,
            hard_count, hard_count + 1, hard_list, nullptr);
      }
      count_launch();
    }
    MPA_LAUNCH_CHECK();
  }
  return MPA_OK;
}

static bool use_grid(int algo, int N1, int N2) {
  if (algo == MPA_ALGO_GRID) return true;
  if (algo == MPA_ALGO_BRUTE) return false;
  return N1 >= 512 && N2 >= 512;
}

}  // namespace mpa

using namespace mpa;

extern "C" {

/* Instrumentation: on != 0 makes the grid searches count the candidate pairs they evaluate
 * (slower instantiation); out[3] (nullable) receives and clears the counters of the current
 * device: {per-part pose search, shape-level pose search, plain clouds}.  Synchronises. */
int mpa_chamfer_pair_count(int on, unsigned long long* out) {
  g_pair_count_on.store(on ? 1 : 0);
  if (out != nullptr) {
    unsigned long long* ctr = pair_counter_device(0);
    if (ctr == nullptr) return MPA_ERR_CUDA;
    MPA_CUDA(cudaDeviceSynchronize());
    MPA_CUDA(cudaMemcpy(out, ctr, 3 * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    MPA_CUDA(cudaMemset(ctr, 0, 3 * sizeof(unsigned long long)));
  }
  return MPA_OK;
}

size_t mpa_chamfer_forward_workspace_bytes(int B, int N1, int N2, int algo) {
  if (B <= 0 || N1 <= 0 || N2 <= 0) return 0;
  if (!use_grid(algo, N1, N2)) return 0;
  return grid_layout(B, N1, N2).total;
}

int mpa_chamfer_forward(const float* xyz1, const float* xyz2, int B, int N1, int N2, float* dist1,
                        int64_t* idx1, float* dist2, int64_t* idx2, int algo, void* ws,
                        size_t ws_bytes, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  MPA_CHECK_ARG(B >= 0 && N1 >= 0 && N2 >= 0, "chamfer_forward: negative size");
  MPA_CHECK_ARG(algo >= MPA_ALGO_AUTO && algo <= MPA_ALGO_GRID, "chamfer_forward: bad algo %d", algo);
  if (B == 0 || (N1 == 0 && N2 == 0)) return MPA_OK;
  MPA_CHECK_ARG((N1 == 0 || (xyz1 && dist1)) && (N2 == 0 || (xyz2 && dist2)),
                "chamfer_forward: null pointer");
  if (N1 > 0 && N2 > 0 && use_grid(algo, N1, N2)) {
    CloudDesc c0{xyz1, nullptr, nullptr, nullptr, nullptr, N1, N1, 0, 0, 0.f, 0};
    CloudDesc c1{xyz2, nullptr, nullptr, nullptr, nullptr, N2, N2, 0, 0, 0.f, 0};
    return run_grid<long long>(c0, c1, B, dist1, (long long*)idx1, dist2, (long long*)idx2, ws,
                               ws_bytes, stream);
  }
  const long long blocks = (long long)B * ((N1 + BF_THREADS - 1) / BF_THREADS +
                                           (N2 + BF_THREADS - 1) / BF_THREADS);
  const long long cap = (long long)num_sms() * 8;
  {
    ProfScope ps("chamfer_brute", stream);
    chamfer_brute_kernel<long long><<<(unsigned)(blocks < cap ? blocks : cap), BF_THREADS, 0, stream>>>(
        xyz1, xyz2, B, N1, N2, dist1, (long long*)idx1, dist2, (long long*)idx2);
  }
  MPA_LAUNCH_CHECK();
  return MPA_OK;
}

int mpa_chamfer_backward(const float* grad_dist1, const float* grad_dist2, const float* xyz1,
                         const float* xyz2, const int64_t* idx1, const int64_t* idx2, int B, int N1,
                         int N2, float* grad_xyz1, float* grad_xyz2, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  MPA_CHECK_ARG(B >= 0 && N1 >= 0 && N2 >= 0, "chamfer_backward: negative size");
  if (B == 0) return MPA_OK;
  if (N1 > 0) MPA_CUDA(cudaMemsetAsync(grad_xyz1, 0, sizeof(float) * 3 * (size_t)B * N1, stream));
  if (N2 > 0) MPA_CUDA(cudaMemsetAsync(grad_xyz2, 0, sizeof(float) * 3 * (size_t)B * N2, stream));
  if (N1 == 0 || N2 == 0) return MPA_OK;
  MPA_CHECK_ARG(grad_dist1 && grad_dist2 && xyz1 && xyz2 && idx1 && idx2 && grad_xyz1 && grad_xyz2,
                "chamfer_backward: null pointer");
  const long long total = (long long)B * (N1 + N2);
  const long long blocks = (total + 255) / 256;
  const long long cap = (long long)num_sms() * 16;
  {
    ProfScope ps("chamfer_backward", stream);
    chamfer_backward_kernel<long long><<<(unsigned)(blocks < cap ? blocks : cap), 256, 0, stream>>>(
        grad_dist1, grad_dist2, xyz1, xyz2, (const long long*)idx1, (const long long*)idx2, B, N1, N2,
        grad_xyz1, grad_xyz2);
  }
  MPA_LAUNCH_CHECK();
  return MPA_OK;
}

int mpa_chamfer_forward_host(const float* h_xyz1, const float* h_xyz2, int B, int N1, int N2,
                             float* h_dist1, int64_t* h_idx1, float* h_dist2, int64_t* h_idx2,
                             int algo, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  MPA_CHECK_ARG(B >= 0 && N1 >= 0 && N2 >= 0, "chamfer_forward_host: negative size");
  if (B == 0 || (N1 == 0 && N2 == 0)) return MPA_OK;
  const size_t n1 = (size_t)B * N1, n2 = (size_t)B * N2;
  const size_t bytes = align_up(n1 * 12, 256) + align_up(n2 * 12, 256) + align_up(n1 * 4, 256) +
                       align_up(n2 * 4, 256) + align_up(n1 * 8, 256) + align_up(n2 * 8, 256);
  char* base = nullptr;
  MPA_CUDA(cudaMallocAsync((void**)&base, bytes, stream));
  char* p = base;
  float* d_x1 = (float*)p; p += align_up(n1 * 12, 256);
  float* d_x2 = (float*)p; p += align_up(n2 * 12, 256);
  float* d_d1 = (float*)p; p += align_up(n1 * 4, 256);
  float* d_d2 = (float*)p; p += align_up(n2 * 4, 256);
  int64_t* d_i1 = (int64_t*)p; p += align_up(n1 * 8, 256);
  int64_t* d_i2 = (int64_t*)p;
  int rc = MPA_OK;
  cudaError_t e = cudaSuccess;
  if (n1) e = cudaMemcpyAsync(d_x1, h_xyz1, n1 * 12, cudaMemcpyHostToDevice, stream);
  if (e == cudaSuccess && n2) e = cudaMemcpyAsync(d_x2, h_xyz2, n2 * 12, cudaMemcpyHostToDevice, stream);
  if (e == cudaSuccess)
    rc = mpa_chamfer_forward(d_x1, d_x2, B, N1, N2, d_d1, d_i1, d_d2, d_i2, algo, nullptr, 0, stream);
  if (e == cudaSuccess && rc == MPA_OK) {
    if (n1 && h_dist1) e = cudaMemcpyAsync(h_dist1, d_d1, n1 * 4, cudaMemcpyDeviceToHost, stream);
    if (e == cudaSuccess && n2 && h_dist2) e = cudaMemcpyAsync(h_dist2, d_d2, n2 * 4, cudaMemcpyDeviceToHost, stream);
    if (e == cudaSuccess && n1 && h_idx1) e = cudaMemcpyAsync(h_idx1, d_i1, n1 * 8, cudaMemcpyDeviceToHost, stream);
    if (e == cudaSuccess && n2 && h_idx2) e = cudaMemcpyAsync(h_idx2, d_i2, n2 * 8, cudaMemcpyDeviceToHost, stream);
  }
  cudaFreeAsync(base, stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
  if (e != cudaSuccess) {
    set_error("chamfer_forward_host: %s", cudaGetErrorString(e));
    return MPA_ERR_CUDA;
  }
  return rc;
}

size_t mpa_pose_chamfer_workspace_bytes(int B, int P, int N, int mode) {
  if (B <= 0 || P <= 0 || N <= 0) return 0;
  if (mode == MPA_CD_PART) return grid_layout(B * P, N, N).total;
  return grid_layout(B, P * N, P * N).total;
}

int mpa_pose_chamfer(const float* pts, const float* quat1, const float* trans1, const float* quat2,
                     const float* trans2, const float* valids, int B, int P, int N, int mode,
                     float* dist1, int32_t* idx1, float* dist2, int32_t* idx2, float* pts1,
                     float* pts2, void* ws, size_t ws_bytes, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  MPA_CHECK_ARG(B >= 0 && P >= 0 && N >= 0, "pose_chamfer: negative size");
  MPA_CHECK_ARG(mode == MPA_CD_PART || mode == MPA_CD_SHAPE, "pose_chamfer: bad mode %d", mode);
  if (B == 0 || P == 0 || N == 0) return MPA_OK;
  MPA_CHECK_ARG(pts && quat1 && quat2 && dist1 && dist2, "pose_chamfer: null pointer");
  MPA_CHECK_ARG(P <= MAX_FAR, "pose_chamfer: at most %d parts per shape", MAX_FAR);
  const bool shape = mode == MPA_CD_SHAPE;
  const int S = shape ? B : B * P;
  const int Nseg = shape ? P * N : N;
  CloudDesc c0{pts, quat1, trans1, valids, pts1, Nseg, N, shape ? 1 : 0, 0, 0.f, 0};
  CloudDesc c1{pts, quat2, trans2, valids, pts2, Nseg, N, shape ? 1 : 0, 0, 0.f, 0};
  return run_grid<int>(c0, c1, S, dist1, idx1, dist2, idx2, ws, ws_bytes, stream);
}

size_t mpa_pose_chamfer_backward_workspace_bytes(int B, int P, int N) {
  if (B <= 0 || P <= 0 || N <= 0) return 0;
  return 2 * align_up(sizeof(float) * 3 * (size_t)B * P * N, 256);
}

int mpa_pose_chamfer_backward(const float* grad_dist1, const float* grad_dist2, const float* pts,
                              const float* quat1, const float* quat2, const float* valids,
                              const float* pts1, const float* pts2, const int32_t* idx1,
                              const int32_t* idx2, int B, int P, int N, int mode, float* grad_quat1,
                              float* grad_trans1, float* grad_quat2, float* grad_trans2, void* ws,
                              size_t ws_bytes, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  MPA_CHECK_ARG(B >= 0 && P >= 0 && N >= 0, "pose_chamfer_backward: negative size");
  MPA_CHECK_ARG(mode == MPA_CD_PART || mode == MPA_CD_SHAPE, "pose_chamfer_backward: bad mode %d", mode);
  if (B == 0 || P == 0 || N == 0) return MPA_OK;
  MPA_CHECK_ARG(grad_dist1 && grad_dist2 && pts && quat1 && quat2 && pts1 && pts2 && idx1 && idx2,
                "pose_chamfer_backward: null pointer");
  const bool shape = mode == MPA_CD_SHAPE;
  const int S = shape ? B : B * P;
  const int Nseg = shape ? P * N : N;
  const size_t one = align_up(sizeof(float) * 3 * (size_t)B * P * N, 256);
  Scratch scratch;
  int rc = scratch.acquire(ws, ws_bytes, 2 * one, stream);
  if (rc != MPA_OK) return rc;
  float* gp1 = (float*)scratch.base;
  float* gp2 = (float*)((char*)scratch.base + one);
  MPA_CUDA(cudaMemsetAsync(scratch.base, 0, 2 * one, stream));
  const long long total = (long long)S * Nseg * 2;
  const long long blocks = (total + 255) / 256;
  const long long cap = (long long)num_sms() * 16;
  {
    ProfScope ps("chamfer_backward", stream);
    chamfer_backward_kernel<int><<<(unsigned)(blocks < cap ? blocks : cap), 256, 0, stream>>>(
        grad_dist1, grad_dist2, pts1, pts2, idx1, idx2, S, Nseg, Nseg, gp1, gp2);
  }
  MPA_LAUNCH_CHECK();
  if (grad_quat1 != nullptr || grad_trans1 != nullptr) {
    rc = launch_se3_backward(quat1, pts, gp1, valids, shape ? 1 : 0, B * P, N, nullptr, grad_quat1,
                             grad_trans1, stream);
    if (rc != MPA_OK) return rc;
  }
  if (grad_quat2 != nullptr || grad_trans2 != nullptr) {
    rc = launch_se3_backward(quat2, pts, gp2, valids, shape ? 1 : 0, B * P, N, nullptr, grad_quat2,
                             grad_trans2, stream);
    if (rc != MPA_OK) return rc;
  }
  return MPA_OK;
}

}  // extern "C"
