// Fused PointNet part encoder on the 5th-generation tensor cores (tcgen05).
//
// Replaces PointNet.forward (models/modules/encoder/pointnet.py:29-41): five
// bias-free 1x1 convolutions 3->64->64->64->128->F with BatchNorm1d after each,
// ReLU after the first four, max over the N points of a part.  The reference
// runs ~15 cuDNN/elementwise kernels that write and re-read [n, C, N] fp32
// activations (~9 GB of HBM traffic at n=640, N=1000); here the activations of
// a 128-point tile never leave the SM:
//
//   orientation   D[channel, point] = W[channel, k] * X[point, k]^T, so a TMEM
//                 lane is a channel and a TMEM column is a point: BatchNorm
//                 scale/shift are per-thread constants, the batch statistics
//                 and the max-pool are per-thread reductions over registers.
//   operands      bf16, K-major, 128-byte swizzle; weights staged once per CTA
//                 (a pre-swizzled 128 KB image), activations written straight
//                 from the previous layer's epilogue into the B-operand tile.
//   accumulators  fp32 in TMEM, read back with tcgen05.ld (32x32b.x32).
//   training BN   needs the statistics of layer l before layer l+1 can run, so
//                 the forward is PHASE = 1..5 launches; launch l recomputes
//                 layers 1..l-1 with their final scale/shift (cheap: input is
//                 12 B/point) and reduces sum / sum-of-squares of layer l in
//                 registers.  Launch 5 also keeps per-part max and min, from
//                 which max_n(BN5(y)) follows for either sign of gamma.
//   eval BN       one launch (PHASE 5 with running statistics).
// PN_GROUPS independent 128-thread pipelines per CTA (each on its own 64-point
// tile and TMEM columns) share the weight image; while one waits for its MMA or
// its TMEM loads the others run their epilogues.
#include <stdlib.h>

#include "mpa_common.cuh"
#include "tc05.cuh"

namespace mpa {

#ifndef MPA_PN_TILE
#define MPA_PN_TILE 64
#endif
#ifndef MPA_PN_GROUPS
#define MPA_PN_GROUPS 4
#endif
constexpr int PN_TILE = MPA_PN_TILE;          // points per tile = UMMA N
constexpr int PN_GROUPS = MPA_PN_GROUPS;      // pipelines per CTA (GROUPS * 2 * TILE <= 512 TMEM columns)
constexpr int PN_THREADS = 128 * PN_GROUPS;
constexpr int PN_WTILE = 128 * 128;           // bytes of one [128 rows][64 k] bf16 tile
// weight image: layer1..4 one tile each (rows = out channels padded to 128),
// layer 5: [mblock 0..1][kblock 0..1] tiles
__host__ __device__ constexpr int pn_w_off(int layer0) { return layer0 * PN_WTILE; }
constexpr int PN_W_BYTES = 8 * PN_WTILE;      // 128 KB
constexpr int PN_ACT_KB = PN_TILE * 128;      // one K-block of the activation tile: [TILE points][64 ch] bf16
constexpr int PN_ACT_BYTES = 2 * PN_ACT_KB;   // per pipeline: two K-blocks
constexpr int PN_SMEM = PN_W_BYTES + PN_GROUPS * PN_ACT_BYTES + 1024;  // + alignment slack
constexpr int PN_MAXC = 256;
static_assert(PN_TILE == 64, "the MN-major activation tile is one 64-point swizzle atom wide");
static_assert(PN_GROUPS * 2 * PN_TILE <= 512, "TMEM has 512 columns");

struct PointNetArgs {
  const float* pts;         // [n_parts, N, 3]
  const float* valids;      // [n_parts] or nullptr
  const uint4* wimage;      // pre-swizzled bf16 weights (PN_W_BYTES)
  const float* scale;       // [5, 256] BN scale of the finished layers
  const float* shift;       // [5, 256]
  float* scale_out;         // same arrays, written by CTA 0 for the layer finished here
  float* shift_out;
  const float* partial_in;  // [ctas, 256, 2] sum / sumsq of layer PHASE-1 (previous launch)
  int partial_in_rows;      // rows of partial_in to add (0: gridDim.x)
  float* partial;           // [ctas, 256, 2] sum / sumsq of layer PHASE
  const float* gamma_prev;  // BatchNorm parameters of layer PHASE-1 (fused finalize) or nullptr
  const float* beta_prev;
  float* rmean_prev;        // running statistics of layer PHASE-1 (updated by CTA 0) or nullptr
  float* rvar_prev;
  float eps, momentum;
  unsigned* pmax;           // [n_parts, 256] ordered-uint max of layer-5 pre-activations
  unsigned* pmin;           // [n_parts, 256]
  int n_parts, N, F;        // F = channels of layer 5 (128 or 256)
  long long* dbg;           // optional cycle stamps (MPA_PN_DEBUG), nullptr in production
  // activation stash (training launches 2..5, optional): launch p leaves the 64-channel bf16
  // operand tile a_{p-1} it built for its last layer in global memory, as the exact shared
  // memory image (8 KB per 64-point tile), and launch p+1 starts from it instead of
  // recomputing layers 1..p-1 from the points.  Same operands, so the results are unchanged.
  float* stats_out;         // [5][4][256] batch mean, rstd, scale, shift per layer + [1] point count, or nullptr
  const uint4* stash_in;    // a_{PHASE-2} tiles or nullptr (start from the points)
  uint4* stash_out;         // a_{PHASE-1} tiles or nullptr
};

__device__ __forceinline__ int pn_cout(int layer, int F) {  // layer 1..5
  return layer <= 3 ? 64 : (layer == 4 ? 128 : F);
}

template <int PHASE>
__global__ void __launch_bounds__(PN_THREADS, 1) pointnet_phase_kernel(PointNetArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t mbar[PN_GROUPS];
  __shared__ uint64_t wbar;  // weight image arrival
  __shared__ uint32_t tmem_base_s;
  __shared__ double dred[PN_GROUPS][128][2];
  __shared__ float fred[PN_GROUPS][PN_MAXC][2];
  __shared__ double s_count;
  __shared__ float s_prev[2][128];

  const long long t_entry = clock64();
  const int tid = threadIdx.x;
  const int g = tid >> 7;            // pipeline
  const int t = tid & 127;           // thread in pipeline = channel (mod 128) = TMEM lane
  const int wq = t >> 5;             // warp quadrant -> TMEM lanes 32*wq..
  uint8_t* wsm = smem;
  uint8_t* act = smem + PN_W_BYTES + g * PN_ACT_BYTES;

  // ---- one-time setup: mbarriers, weights (bulk async copy), TMEM ----
  constexpr uint32_t W_BYTES = (PHASE < 5 ? PHASE : 8) * PN_WTILE;  // only the layers this phase runs
  if (tid == 0) {
    for (int i = 0; i < PN_GROUPS; ++i) tc::mbar_init(&mbar[i], 1);
    tc::mbar_init(&wbar, 1);
    tc::fence_barrier_init();
  }
  __syncthreads();
  // resuming from the stash (launches 4, 5): layers 1..PHASE-2 are not run; the first 32 KB of
  // the weight region (images of layers 1 and 2) hold the pipelines' prefetch tiles instead
  const bool resume = PHASE >= 4 && a.stash_in != nullptr;  // uniform
  const uint32_t w_first = resume ? (uint32_t)(PHASE - 2) * PN_WTILE : 0u;
  if (tid == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(tc::smem_u32(&wbar)),
                 "r"(W_BYTES - w_first)
                 : "memory");
    for (uint32_t off = w_first; off < W_BYTES; off += PN_WTILE)
      asm volatile(
          "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
              tc::smem_u32(wsm + off)),
          "l"(reinterpret_cast<const uint8_t*>(a.wimage) + off), "r"((uint32_t)PN_WTILE),
          "r"(tc::smem_u32(&wbar))
          : "memory");
  }
  if (tid < 32) tc::tmem_alloc<512>(&tmem_base_s);

  // ---- fused BatchNorm finalize of the previous layer (training) ----
  float sc_prev = 0.f, sh_prev = 0.f;
  if (PHASE >= 2 && a.gamma_prev != nullptr) {
    constexpr int CP = PHASE <= 4 ? 64 : 128;  // channels of layer PHASE-1
    // number of valid points
    if (tid == 0) s_count = 0.0;
    double cnt = 0.0;
    for (int p = tid; p < a.n_parts; p += PN_THREADS)
      cnt += (a.valids == nullptr || a.valids[p] != 0.0f) ? (double)a.N : 0.0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    __syncthreads();
    if ((tid & 31) == 0) atomicAdd(&s_count, cnt);  // integers: exact in any order
    // per-CTA partial sums of the previous launch: group g adds CTAs g, g+GROUPS, ...
    double s1 = 0.0, s2 = 0.0;
    if (t < CP) {
      const int rows_in = a.partial_in_rows > 0 ? a.partial_in_rows : (int)gridDim.x;
      for (int b = g; b < rows_in; b += PN_GROUPS) {
        s1 += (double)a.partial_in[((long long)b * PN_MAXC + t) * 2];
        s2 += (double)a.partial_in[((long long)b * PN_MAXC + t) * 2 + 1];
      }
      dred[g][t][0] = s1;
      dred[g][t][1] = s2;
    }
    __syncthreads();
    if (t < CP) {
      s1 = 0.0; s2 = 0.0;
#pragma unroll
      for (int k = 0; k < PN_GROUPS; ++k) { s1 += dred[k][t][0]; s2 += dred[k][t][1]; }
      const double n = s_count;
      const double mean = s1 / n;
      const double var = fmax(s2 / n - mean * mean, 0.0);
      sc_prev = a.gamma_prev[t] * (float)(1.0 / sqrt(var + (double)a.eps));
      sh_prev = a.beta_prev[t] - (float)mean * sc_prev;
      if (g == 0) { s_prev[0][t] = sc_prev; s_prev[1][t] = sh_prev; }
      if (blockIdx.x == 0 && g == 0) {
        a.scale_out[(PHASE - 2) * PN_MAXC + t] = sc_prev;
        a.shift_out[(PHASE - 2) * PN_MAXC + t] = sh_prev;
        if (a.stats_out != nullptr) {  // what the backward pass needs of this layer's BatchNorm
          float* so = a.stats_out + (PHASE - 2) * 4 * PN_MAXC;
          so[t] = (float)mean;
          so[PN_MAXC + t] = (float)(1.0 / sqrt(var + (double)a.eps));
          so[2 * PN_MAXC + t] = sc_prev;
          so[3 * PN_MAXC + t] = sh_prev;
          if (t == 0) a.stats_out[5 * 4 * PN_MAXC] = (float)n;
        }
        if (a.rmean_prev != nullptr) {
          a.rmean_prev[t] = (1.f - a.momentum) * a.rmean_prev[t] + a.momentum * (float)mean;
          const double unbiased = n > 1.0 ? var * n / (n - 1.0) : var;
          a.rvar_prev[t] = (1.f - a.momentum) * a.rvar_prev[t] + a.momentum * (float)unbiased;
        }
      }
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = tmem_base_s + (uint32_t)(g * 2 * PN_TILE);  // this pipeline's columns
  const uint32_t tmem_lane = tmem + ((uint32_t)(wq * 32) << 16);  // this warp's lane quadrant

  // per-thread BN constants of the finished layers (channel t; layer 4 has 128 channels)
  float sc[4], sh[4];
#pragma unroll
  for (int l = 0; l < 4; ++l) {
    // layers 1-3 have 64 channels whose accumulator rows are duplicated in TMEM lanes
    // 64..127 (duplicated weight rows), so lane t works on channel t & 63
    const int ch = l < 3 ? (t & 63) : t;
    const bool fused_here = (l + 2 == PHASE) && a.gamma_prev != nullptr;
    sc[l] = fused_here ? s_prev[0][ch] : ((l + 1 < PHASE) ? a.scale[l * PN_MAXC + ch] : 0.f);
    sh[l] = fused_here ? s_prev[1][ch] : ((l + 1 < PHASE) ? a.shift[l * PN_MAXC + ch] : 0.f);
  }
  const long long t_setup = clock64();
  tc::mbar_wait(&wbar, 0);  // weight image has landed
  const long long t_weights = clock64();
  constexpr uint32_t IDESC = tc::make_idesc_bf16_bmn(128, PN_TILE);  // activations are MN-major
  const int mblocks5 = a.F / 128;
  float ssum[2] = {0.f, 0.f}, ssq[2] = {0.f, 0.f};
  uint32_t parity = 0;

  const int tiles_per_part = (a.N + PN_TILE - 1) / PN_TILE;
  const int n_tiles = a.n_parts * tiles_per_part;  // 32-bit tile arithmetic (host checks the range)
  const int worker = blockIdx.x * PN_GROUPS + g;
  const int n_workers = gridDim.x * PN_GROUPS;

  // tiles of this pipeline, padded parts skipped (uniform per pipeline)
  auto next_tile = [&](int tl) {
    while (tl < n_tiles && a.valids != nullptr && a.valids[tl / tiles_per_part] == 0.0f) tl += n_workers;
    return tl;
  };
  // stash prefetch: the a_{PHASE-2} tile of the NEXT tile streams into this pipeline's 8 KB
  // buffer (cp.async, 4 x 16 B per thread) while the current tile is in its later layers
  uint8_t* pf = wsm + g * PN_ACT_KB;
  static_assert(PN_GROUPS * PN_ACT_KB <= 2 * PN_WTILE, "prefetch tiles fit the images of layers 1 and 2");
  auto prefetch = [&](int tl) {
    const uint4* src = a.stash_in + (long long)tl * (PN_ACT_KB / 16);
#pragma unroll
    for (int j = 0; j < PN_ACT_KB / 16 / 128; ++j)
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(tc::smem_u32(pf + (j * 128 + t) * 16)),
                   "l"(src + j * 128 + t)
                   : "memory");
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  int tile = next_tile(worker);
  if (resume && tile < n_tiles) prefetch(tile);

  for (; tile < n_tiles;) {
    const int tile_next = next_tile(tile + n_workers);
    const int part = tile / tiles_per_part;
    const int p0 = (tile - part * tiles_per_part) * PN_TILE;
    const int npts = min(PN_TILE, a.N - p0);

    const bool dbg_on = a.dbg != nullptr && blockIdx.x == 0 && tid == 0 && tile < 8 * n_workers;
    long long* dbg = dbg_on ? a.dbg + (tile / n_workers) * 16 : nullptr;
    if (dbg_on) dbg[0] = clock64();
    // the previous tile's stash copy read `act` with ordinary loads: all of them are done
    if (PHASE <= 4 && a.stash_out != nullptr) tc::group_sync(1 + g, 128);
    if (resume) {
      // ---- B operand of layer PHASE-1: the prefetched a_{PHASE-2} tile, used in place ----
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    } else {
    // ---- layer-1 B operand, MN-major: row k = coordinate k of the TILE points
    // (rows 3..15 zero).  Thread i fills one 16-byte chunk: row i>>3, points 8*(i&7).. ----
      const int row = t >> 3, chunk = t & 7;
      uint32_t w[4] = {0u, 0u, 0u, 0u};
      if (row < 3) {
#pragma unroll
        for (int h = 0; h < 4; ++h) {
          const int pa = chunk * 8 + 2 * h, pb = pa + 1;
          const float va = pa < npts ? a.pts[((long long)part * a.N + p0 + pa) * 3 + row] : 0.f;
          const float vb = pb < npts ? a.pts[((long long)part * a.N + p0 + pb) * 3 + row] : 0.f;
          const __nv_bfloat162 pk = __floats2bfloat162_rn(va, vb);
          w[h] = *reinterpret_cast<const uint32_t*>(&pk);
        }
      }
      if (row < 16 && chunk * 8 < PN_TILE)
        *reinterpret_cast<uint4*>(act + row * 128 + (((chunk ^ row) & 7) << 4)) =
            make_uint4(w[0], w[1], w[2], w[3]);
    }
    tc::fence_async_smem();
    tc::fence_before_sync();
    tc::group_sync(1 + g, 128);

#pragma unroll
    for (int layer = 1; layer <= PHASE; ++layer) {
      if (resume && layer < PHASE - 1) continue;  // their output came from the stash
      const uint8_t* bsrc = (resume && layer == PHASE - 1) ? pf : act;
      const int K = layer == 1 ? 16 : (layer == 5 ? 128 : 64);
      const int mblocks = layer == 5 ? mblocks5 : 1;
      // ---- MMA issue (one thread) ----
      if (t == 0) {
        tc::fence_after_sync();
        for (int mb = 0; mb < mblocks; ++mb) {
          for (int k = 0; k < K; k += 16) {
            const int kb = k >> 6, ks = k & 63;
            const uint32_t wa = tc::smem_u32(wsm) + pn_w_off(layer - 1) +
                                (layer == 5 ? (mb * 2 + kb) * PN_WTILE : 0) + ks * 2;
            const uint32_t ba = tc::smem_u32(bsrc) + k * 128;  // 16 channel rows per K step
            tc::mma_bf16(tmem + (uint32_t)(mb * PN_TILE), tc::make_desc_sw128(wa),
                         tc::make_desc_sw128_mn(ba, 128 * 128), IDESC, k > 0 ? 1u : 0u);
          }
        }
        tc::mma_commit(&mbar[g]);
      }
      if (dbg_on) dbg[3 * (layer - 1) + 1] = clock64();  // operand ready + MMAs issued
      tc::mbar_wait(&mbar[g], parity);
      parity ^= 1u;
      tc::fence_after_sync();
      if (dbg_on) dbg[3 * (layer - 1) + 2] = clock64();  // accumulator complete
      if (resume && layer == PHASE - 1 && tile_next < n_tiles) prefetch(tile_next);  // pf is free again

      // ---- epilogue ----
      // 64-channel layers: lanes 64..127 hold a copy of rows 0..63, so warps 2,3 take
      // the points 32..63 of the same channels and all four warps stay busy.
      const bool narrow = layer <= 3;
      if (layer < PHASE) {
        // BN + ReLU, bf16, into the B-operand tile of the next layer (this layer's
        // input tile is dead: its MMA has completed)
        const float s = sc[layer - 1], b = sh[layer - 1];
        const int ch = narrow ? (t & 63) : t;
        uint8_t* dst = act + ch * 128;  // channel row of the MN-major tile (TILE points = 128 B)
        const int jbeg = narrow ? (t >> 6) * 32 : 0;
        float v[64];
        tc::tmem_ld32(tmem_lane + (uint32_t)jbeg, v);
        if (!narrow) tc::tmem_ld32(tmem_lane + 32u, v + 32);
        tc::tmem_ld_wait();
#pragma unroll
        for (int jj = 0; jj < 64; jj += 8) {
          if (narrow && jj >= 32) break;
          uint32_t w[4];
#pragma unroll
          for (int h = 0; h < 4; ++h) {
            const float r0 = fmaxf(fmaf(v[jj + 2 * h], s, b), 0.f);
            const float r1 = fmaxf(fmaf(v[jj + 2 * h + 1], s, b), 0.f);
            const __nv_bfloat162 pk = __floats2bfloat162_rn(r0, r1);
            w[h] = *reinterpret_cast<const uint32_t*>(&pk);
          }
          const int chunk = (jbeg + jj) >> 3;
          *reinterpret_cast<uint4*>(dst + (((chunk ^ ch) & 7) << 4)) = make_uint4(w[0], w[1], w[2], w[3]);
        }
        tc::fence_async_smem();
        tc::fence_before_sync();
        tc::group_sync(1 + g, 128);
        if (PHASE <= 4 && layer == PHASE - 1 && a.stash_out != nullptr) {
          // a_{PHASE-1} (64 channels = K-block 0) -> global, coalesced, while the next MMA reads it
          uint4* dstg = a.stash_out + (long long)tile * (PN_ACT_KB / 16);
#pragma unroll
          for (int j = 0; j < PN_ACT_KB / 16 / 128; ++j)
            __stcs(dstg + j * 128 + t, *reinterpret_cast<const uint4*>(act + (j * 128 + t) * 16));
        }
      } else {
        // statistics of this layer's pre-activations (+ max/min for layer 5)
#pragma unroll
        for (int mb = 0; mb < 2; ++mb) {
          if (mb < mblocks) {
            const int jbeg = narrow ? (t >> 6) * 32 : 0;
            const int jn = narrow ? 32 : 64;
            float v[64];
            tc::tmem_ld32(tmem_lane + (uint32_t)(mb * PN_TILE + jbeg), v);
            if (!narrow) tc::tmem_ld32(tmem_lane + (uint32_t)(mb * PN_TILE + 32), v + 32);
            tc::tmem_ld_wait();
            float mx = -3.0e38f, mn = 3.0e38f, s1 = 0.f, s2 = 0.f;
            if (npts == PN_TILE) {  // full tile: no per-point predicate
#pragma unroll
              for (int j = 0; j < 64; ++j) {
                if (j < jn) {
                  s1 += v[j];
                  s2 = fmaf(v[j], v[j], s2);
                  if (PHASE == 5) { mx = fmaxf(mx, v[j]); mn = fminf(mn, v[j]); }
                }
              }
            } else {
#pragma unroll
              for (int j = 0; j < 64; ++j) {
                if (j < jn && jbeg + j < npts) {
                  s1 += v[j];
                  s2 = fmaf(v[j], v[j], s2);
                  if (PHASE == 5) { mx = fmaxf(mx, v[j]); mn = fminf(mn, v[j]); }
                }
              }
            }
            ssum[mb] += s1;
            ssq[mb] += s2;
            if (PHASE == 5) {
              const long long o = (long long)part * PN_MAXC + mb * 128 + t;
              atomicMax(a.pmax + o, tc::float_to_ordered(mx));
              atomicMin(a.pmin + o, tc::float_to_ordered(mn));
            }
          }
        }
        tc::fence_before_sync();  // TMEM reads done before the next tile's MMA overwrites
      }
      if (dbg_on) dbg[3 * (layer - 1) + 3] = clock64();  // epilogue done
    }
    tile = tile_next;
  }

  if (a.dbg != nullptr && blockIdx.x == 0 && tid == 0) {
    a.dbg[128] = t_setup - t_entry; a.dbg[129] = t_weights - t_entry; a.dbg[130] = clock64() - t_entry;
  }
  // ---- per-CTA partial statistics (groups combined in a fixed order) ----
#pragma unroll
  for (int mb = 0; mb < 2; ++mb) {
    fred[g][mb * 128 + t][0] = ssum[mb];
    fred[g][mb * 128 + t][1] = ssq[mb];
  }
  __syncthreads();
  if (tid < PN_MAXC) {
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int k = 0; k < PN_GROUPS; ++k) {
      s1 += fred[k][tid][0]; s2 += fred[k][tid][1];
      if (PHASE <= 3 && tid < 64) { s1 += fred[k][tid + 64][0]; s2 += fred[k][tid + 64][1]; }
    }
    float* o = a.partial + ((long long)blockIdx.x * PN_MAXC + tid) * 2;
    o[0] = s1;
    o[1] = s2;
  }
  tc::fence_before_sync();
  __syncthreads();
  if (tid < 32) tc::tmem_dealloc<512>(tmem_base_s);
}

// fp32 conv weights [Cout, Cin] -> pre-swizzled bf16 smem image (zero-filled beforehand)
__device__ __forceinline__ void pointnet_fill_weights(const float* w1, const float* w2, const float* w3,
                                                      const float* w4, const float* w5, int F,
                                                      uint8_t* image) {
  // one thread per (layer, out channel, in channel)
  const int cin[5] = {3, 64, 64, 64, 128};
  const int cout[5] = {64, 64, 64, 128, F};
  const float* w[5] = {w1, w2, w3, w4, w5};
  for (int layer = 0; layer < 5; ++layer) {
    const int n = cin[layer] * cout[layer];
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
      const int co = i / cin[layer], ci = i % cin[layer];
      const int mb = co >> 7, row = co & 127, kb = ci >> 6, col = ci & 63;
      uint8_t* tile = image + pn_w_off(layer) + (layer == 4 ? (mb * 2 + kb) * PN_WTILE : 0);
      const __nv_bfloat16 v = __float2bfloat16_rn(w[layer][i]);
      *reinterpret_cast<__nv_bfloat16*>(tile + tc::sw128_offset(row, col)) = v;
      if (layer < 3)  // 64-channel layers: rows 64..127 repeat rows 0..63 (see the epilogue)
        *reinterpret_cast<__nv_bfloat16*>(tile + tc::sw128_offset(row + 64, col)) = v;
    }
  }
}

// reduce the per-CTA partial sums of layer `layer` (0-based) in a fixed order,
// produce BN scale/shift, update the running statistics (train mode), and for the
// last layer turn per-part max/min into the pooled features.
__global__ void pointnet_finalize_kernel(const float* partial, int n_workers, int layer, int C,
                                         const float* valids, int n_parts, int N,
                                         const float* gamma, const float* beta, float eps,
                                         float momentum, float* running_mean, float* running_var,
                                         float* scale, float* shift, float* stats_out) {
  // one warp per channel; lanes stride over the workers (fixed order -> deterministic)
  const int c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (c >= C) return;
  double s1 = 0.0, s2 = 0.0, cnt = 0.0;
  for (int w = lane; w < n_workers; w += 32) {
    s1 += (double)partial[((long long)w * PN_MAXC + c) * 2];
    s2 += (double)partial[((long long)w * PN_MAXC + c) * 2 + 1];
  }
  for (int p = lane; p < n_parts; p += 32)
    cnt += (valids == nullptr || valids[p] != 0.0f) ? (double)N : 0.0;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s1 += __shfl_xor_sync(0xffffffffu, s1, o);
    s2 += __shfl_xor_sync(0xffffffffu, s2, o);
    cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  }
  if (lane != 0) return;
  const double mean = s1 / cnt;
  const double var = fmax(s2 / cnt - mean * mean, 0.0);
  const float sc = gamma[c] * (float)(1.0 / sqrt(var + (double)eps));
  scale[layer * PN_MAXC + c] = sc;
  shift[layer * PN_MAXC + c] = beta[c] - (float)mean * sc;
  if (stats_out != nullptr) {
    float* so = stats_out + layer * 4 * PN_MAXC;
    so[c] = (float)mean;
    so[PN_MAXC + c] = (float)(1.0 / sqrt(var + (double)eps));
    so[2 * PN_MAXC + c] = sc;
    so[3 * PN_MAXC + c] = beta[c] - (float)mean * sc;
  }
  if (running_mean != nullptr) {
    running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * (float)mean;
    const double unbiased = cnt > 1.0 ? var * cnt / (cnt - 1.0) : var;
    running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unbiased;
  }
}

__global__ void pointnet_eval_affine_kernel(const float* gamma, const float* beta,
                                            const float* running_mean, const float* running_var,
                                            float eps, int layer, int C, float* scale, float* shift) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const float sc = gamma[c] * rsqrtf(running_var[c] + eps);
  scale[layer * PN_MAXC + c] = sc;
  shift[layer * PN_MAXC + c] = beta[c] - running_mean[c] * sc;
}

__device__ __forceinline__ void pointnet_init_minmax(unsigned* pmax, unsigned* pmin, long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    pmax[i] = 0u;           // below every encoded float
    pmin[i] = 0xffffffffu;  // above every encoded float
  }
}

// Statistics of the first layer without running it.  z = W_1 x is linear in the 3 coordinates,
// so over all valid points  sum_c = w_c . S1  and  sumsq_c = w_c^T S2 w_c  with the first and
// second moments S1 [3], S2 [3x3] of the (bf16-rounded, as the MMA sees them) points.  One pass
// over 12 B/point replaces launch 1 (a K=16 MMA + epilogue per tile).  Every CTA writes its
// partial moments; the last one to finish (ticket) adds them in a fixed order in fp64 and emits
// the two rows (hi, lo floats of the fp64 sums) launch 2 reads as `partial_in`.
// The same launch packs the weight image and initialises the pooling extrema (the three
// preparations are independent; as separate launches they were 22 us in a row).
constexpr int PN_MOM_THREADS = 256;
struct PointNetPrepArgs {
  const float* w[5]; int F; uint8_t* image;        // weight image (zero-filled beforehand)
  unsigned* pmax; unsigned* pmin; long long n_mm;  // pooling extrema
  const float* pts; const float* valids; int n_parts, N;
  double* mom; unsigned* ticket; float* partial_out;  // moments: partial_out == nullptr skips them
};
__global__ void __launch_bounds__(PN_MOM_THREADS)
pointnet_prepare_kernel(PointNetPrepArgs pa) {
  pointnet_fill_weights(pa.w[0], pa.w[1], pa.w[2], pa.w[3], pa.w[4], pa.F, pa.image);
  pointnet_init_minmax(pa.pmax, pa.pmin, pa.n_mm);
  if (pa.partial_out == nullptr) return;
  const float* __restrict__ pts = pa.pts;
  const float* __restrict__ valids = pa.valids;
  const int n_parts = pa.n_parts, N = pa.N;
  const float* __restrict__ w1 = pa.w[0];
  double* __restrict__ mom = pa.mom;
  unsigned* __restrict__ ticket = pa.ticket;
  float* __restrict__ partial_out = pa.partial_out;
  __shared__ double red[PN_MOM_THREADS / 32][9];
  __shared__ double tot[9];
  __shared__ bool last;
  float m[9] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};  // x y z xx xy xz yy yz zz
  auto add_point = [&](float fx, float fy, float fz) {
    const float x = __bfloat162float(__float2bfloat16_rn(fx));
    const float y = __bfloat162float(__float2bfloat16_rn(fy));
    const float z = __bfloat162float(__float2bfloat16_rn(fz));
    m[0] += x; m[1] += y; m[2] += z;
    m[3] = fmaf(x, x, m[3]); m[4] = fmaf(x, y, m[4]); m[5] = fmaf(x, z, m[5]);
    m[6] = fmaf(y, y, m[6]); m[7] = fmaf(y, z, m[7]); m[8] = fmaf(z, z, m[8]);
  };
  // a CTA takes whole parts (one validity test per part); a thread takes 4 consecutive points =
  // three 16-byte loads when the part's rows are 16-byte aligned
  for (int part = blockIdx.x; part < n_parts; part += gridDim.x) {
    if (valids != nullptr && valids[part] == 0.0f) continue;
    const float* pp = pts + (long long)part * N * 3;
    const bool vec = ((N & 3) == 0) && ((reinterpret_cast<uintptr_t>(pp) & 15) == 0);
    if (vec) {
      for (int q = threadIdx.x; q < N / 4; q += blockDim.x) {
        const float4 a0 = __ldg(reinterpret_cast<const float4*>(pp) + 3 * q);
        const float4 a1 = __ldg(reinterpret_cast<const float4*>(pp) + 3 * q + 1);
        const float4 a2 = __ldg(reinterpret_cast<const float4*>(pp) + 3 * q + 2);
        add_point(a0.x, a0.y, a0.z);
        add_point(a0.w, a1.x, a1.y);
        add_point(a1.z, a1.w, a2.x);
        add_point(a2.y, a2.z, a2.w);
      }
    } else {
      for (int i = threadIdx.x; i < N; i += blockDim.x) add_point(pp[3 * i], pp[3 * i + 1], pp[3 * i + 2]);
    }
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < 9; ++k) {
    double v = (double)m[k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) red[warp][k] = v;
  }
  __syncthreads();
  if (threadIdx.x < 9) {
    double v = 0.0;
    for (int w = 0; w < PN_MOM_THREADS / 32; ++w) v += red[w][threadIdx.x];
    mom[(long long)blockIdx.x * 9 + threadIdx.x] = v;
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) last = atomicAdd(ticket, 1u) == gridDim.x - 1;
  __syncthreads();
  if (!last) return;
  __threadfence();
  // warp w adds moment w (warp 0 also moment 8) over the CTAs: lanes stride, then a fixed tree
  for (int k = warp; k < 9; k += PN_MOM_THREADS / 32) {
    double v = 0.0;
    for (unsigned b = lane; b < gridDim.x; b += 32) v += __ldcg(mom + (long long)b * 9 + k);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) tot[k] = v;
  }
  __syncthreads();
  if (threadIdx.x < 64) {  // channel of layer 1
    const int c = threadIdx.x;
    const double w0 = (double)__bfloat162float(__float2bfloat16_rn(w1[3 * c]));
    const double wa = (double)__bfloat162float(__float2bfloat16_rn(w1[3 * c + 1]));
    const double wb = (double)__bfloat162float(__float2bfloat16_rn(w1[3 * c + 2]));
    const double* S = tot;
    const double s1 = w0 * S[0] + wa * S[1] + wb * S[2];
    const double s2 = w0 * w0 * S[3] + wa * wa * S[6] + wb * wb * S[8] +
                      2.0 * (w0 * wa * S[4] + w0 * wb * S[5] + wa * wb * S[7]);
    const float h1 = (float)s1, h2 = (float)s2;
    partial_out[(0 * PN_MAXC + c) * 2] = h1;
    partial_out[(0 * PN_MAXC + c) * 2 + 1] = h2;
    partial_out[(1 * PN_MAXC + c) * 2] = (float)(s1 - (double)h1);
    partial_out[(1 * PN_MAXC + c) * 2 + 1] = (float)(s2 - (double)h2);
  }
  if (threadIdx.x == 0) *ticket = 0u;  // ready for the next forward (graph replays included)
}

// feats[part, c] = max_n BN5(y)[c] = scale*max + shift (scale >= 0) or scale*min + shift
__global__ void pointnet_pool_kernel(const unsigned* pmax, const unsigned* pmin, const float* scale,
                                     const float* shift, const float* valids, int n_parts, int F,
                                     float* feats) {
  const long long total = (long long)n_parts * F;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int part = (int)(i / F), c = (int)(i % F);
    float out = 0.f;
    if (valids == nullptr || valids[part] != 0.0f) {
      const float sc = scale[4 * PN_MAXC + c], sh = shift[4 * PN_MAXC + c];
      const float mx = tc::ordered_to_float(pmax[(long long)part * PN_MAXC + c]);
      const float mn = tc::ordered_to_float(pmin[(long long)part * PN_MAXC + c]);
      out = sc >= 0.f ? fmaf(sc, mx, sh) : fmaf(sc, mn, sh);
    }
    feats[i] = out;
  }
}

template <int PHASE>
static int launch_phase(const PointNetArgs& a, int grid, cudaStream_t stream) {
  static DeviceOnce attr;
  if (attr.pending()) {
    MPA_CUDA(cudaFuncSetAttribute(pointnet_phase_kernel<PHASE>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, PN_SMEM));
    attr.done();
  }
  {
    static const char* names[5] = {"pointnet_phase1", "pointnet_phase2", "pointnet_phase3",
                                   "pointnet_phase4", "pointnet_phase5"};
    ProfScope ps(names[PHASE - 1], stream);
    pointnet_phase_kernel<PHASE><<<grid, PN_THREADS, PN_SMEM, stream>>>(a);
  }
  MPA_LAUNCH_CHECK();
  return MPA_OK;
}

}  // namespace mpa

using namespace mpa;

extern "C" {

size_t mpa_pointnet_workspace_bytes(int n_parts);
int mpa_pointnet_forward_ex(const float* pts, const float* valids, int n_parts, int N, int F,
                            const float* const* conv_w, const float* const* bn_gamma,
                            const float* const* bn_beta, float* const* bn_running_mean,
                            float* const* bn_running_var, int training, float eps, float momentum,
                            float* feats, float* bn_batch_stats, void* ws, size_t ws_bytes, void* stream_);
static size_t pointnet_stash_bytes(int n_parts, int N) {  // one ping-pong buffer of a_k tiles
  const size_t tiles = (size_t)n_parts * ((N + PN_TILE - 1) / PN_TILE);
  return align_up(tiles * PN_ACT_KB, 256);
}

// with room for the activation stash of the training launches (2 buffers of 128 B per point)
size_t mpa_pointnet_workspace_bytes_n(int n_parts, int N) {
  if (n_parts <= 0) return 0;
  return mpa_pointnet_workspace_bytes(n_parts) + 2 * pointnet_stash_bytes(n_parts, N);
}

size_t mpa_pointnet_workspace_bytes(int n_parts) {
  if (n_parts <= 0) return 0;
  size_t o = 0;
  o += align_up(PN_W_BYTES, 256) + 256;                             // weight image, ticket of the moments kernel
  o += align_up(sizeof(float) * 2 * 5 * PN_MAXC, 256);              // scale, shift
  o += align_up(sizeof(float) * 2 * PN_MAXC * 2 * 160 * PN_GROUPS, 256);  // partial (<=160 CTAs)
  o += 2 * align_up(sizeof(unsigned) * (size_t)n_parts * PN_MAXC, 256);   // pmax, pmin
  return o;
}

int mpa_pointnet_forward(const float* pts, const float* valids, int n_parts, int N, int F,
                         const float* const* conv_w, const float* const* bn_gamma,
                         const float* const* bn_beta, float* const* bn_running_mean,
                         float* const* bn_running_var, int training, float eps, float momentum,
                         float* feats, void* ws, size_t ws_bytes, void* stream_) {
  return mpa_pointnet_forward_ex(pts, valids, n_parts, N, F, conv_w, bn_gamma, bn_beta, bn_running_mean,
                                 bn_running_var, training, eps, momentum, feats, nullptr, ws, ws_bytes, stream_);
}

int mpa_pointnet_forward_ex(const float* pts, const float* valids, int n_parts, int N, int F,
                            const float* const* conv_w, const float* const* bn_gamma,
                            const float* const* bn_beta, float* const* bn_running_mean,
                            float* const* bn_running_var, int training, float eps, float momentum,
                            float* feats, float* bn_batch_stats, void* ws, size_t ws_bytes, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  MPA_CHECK_ARG(n_parts >= 0 && N >= 0, "pointnet_forward: negative size");
  MPA_CHECK_ARG(F == 128 || F == 256, "pointnet_forward: feat_dim must be 128 or 256 (got %d)", F);
  if (n_parts == 0) return MPA_OK;
  MPA_CHECK_ARG(N > 0, "pointnet_forward: parts need at least one point");
  MPA_CHECK_ARG(pts && conv_w && bn_gamma && bn_beta && bn_running_mean && bn_running_var && feats,
                "pointnet_forward: null pointer");
  int sms = 0, dev = 0;
  MPA_CUDA(cudaGetDevice(&dev));
  MPA_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int tiles_per_part = (N + PN_TILE - 1) / PN_TILE;
  const long long n_tiles = (long long)n_parts * tiles_per_part;
  MPA_CHECK_ARG(n_tiles + 4096 < (1ll << 31), "pointnet_forward: n_parts * ceil(N / 64) must fit 31 bits");
  int grid = (int)((n_tiles + PN_GROUPS - 1) / PN_GROUPS);
  if (grid > sms) grid = sms;
  if (grid > 160) grid = 160;
  Scratch scratch;
  const size_t need = mpa_pointnet_workspace_bytes(n_parts);
  // a caller that sized the workspace with mpa_pointnet_workspace_bytes_n gets the stash
  static const bool no_stash = getenv("MPA_PN_NO_STASH") != nullptr;  // A/B switch
  const bool stash = training && ws != nullptr && !no_stash &&
                     ws_bytes >= mpa_pointnet_workspace_bytes_n(n_parts, N);
  int rc = scratch.acquire(ws, ws_bytes, need, stream);
  if (rc != MPA_OK) return rc;
  char* p = (char*)scratch.base;
  uint8_t* image = (uint8_t*)p; p += align_up(PN_W_BYTES, 256);
  unsigned* ticket = (unsigned*)p; p += 256;  // zeroed together with the image
  float* scale = (float*)p;
  float* shift = scale + 5 * PN_MAXC; p += align_up(sizeof(float) * 2 * 5 * PN_MAXC, 256);
  float* partial = (float*)p; p += align_up(sizeof(float) * 2 * PN_MAXC * 2 * 160 * PN_GROUPS, 256);
  unsigned* pmax = (unsigned*)p; p += align_up(sizeof(unsigned) * (size_t)n_parts * PN_MAXC, 256);
  unsigned* pmin = (unsigned*)p; p += align_up(sizeof(unsigned) * (size_t)n_parts * PN_MAXC, 256);
  uint4* stash_buf[2] = {nullptr, nullptr};
  if (stash) {
    stash_buf[0] = (uint4*)p;
    stash_buf[1] = (uint4*)(p + pointnet_stash_bytes(n_parts, N));
  }

  MPA_CUDA(cudaMemsetAsync(image, 0, PN_W_BYTES + 256, stream));
  static const bool run_phase1 = getenv("MPA_PN_PHASE1") != nullptr;  // A/B: layer-1 statistics by the MMA launch
  PointNetPrepArgs pa{};
  for (int i = 0; i < 5; ++i) pa.w[i] = conv_w[i];
  pa.F = F; pa.image = image; pa.pmax = pmax; pa.pmin = pmin; pa.n_mm = (long long)n_parts * PN_MAXC;
  pa.pts = pts; pa.valids = valids; pa.n_parts = n_parts; pa.N = N;
  pa.mom = (double*)(partial + (size_t)2 * 2 * PN_MAXC * 160);  // slack of the partial region: [2 sms][9]
  pa.ticket = ticket;
  pa.partial_out = (training && !run_phase1) ? partial : nullptr;  // = partial_buf[0], what launch 1 would write
  {
    ProfScope ps("pointnet_prepare", stream);
    pointnet_prepare_kernel<<<2 * sms, PN_MOM_THREADS, 0, stream>>>(pa);
  }
  MPA_LAUNCH_CHECK();

  long long* dbg = nullptr;
  static const bool want_dbg = getenv("MPA_PN_DEBUG") != nullptr;
  if (want_dbg) {
    MPA_CUDA(cudaMalloc((void**)&dbg, sizeof(long long) * 16 * 9));
    MPA_CUDA(cudaMemset(dbg, 0, sizeof(long long) * 16 * 9));
  }
  // per-CTA partial sums, double buffered: launch l reads what launch l-1 wrote
  float* partial_buf[2] = {partial, partial + (size_t)2 * PN_MAXC * 160};

  PointNetArgs a{};
  a.pts = pts; a.valids = valids; a.wimage = (const uint4*)image;
  a.scale = scale; a.shift = shift; a.scale_out = scale; a.shift_out = shift;
  a.pmax = pmax; a.pmin = pmin; a.n_parts = n_parts; a.N = N; a.F = F; a.dbg = dbg;
  a.eps = eps; a.momentum = momentum;
  a.stats_out = training ? bn_batch_stats : nullptr;
  const int C[5] = {64, 64, 64, 128, F};
  if (training) {
    for (int layer = 0; layer < 5; ++layer) {
      a.partial = partial_buf[layer & 1];
      a.partial_in = partial_buf[(layer + 1) & 1];
      // launch `layer` finalizes the BatchNorm of layer-1 in its prologue
      a.gamma_prev = layer > 0 ? bn_gamma[layer - 1] : nullptr;
      a.beta_prev = layer > 0 ? bn_beta[layer - 1] : nullptr;
      a.rmean_prev = layer > 0 ? bn_running_mean[layer - 1] : nullptr;
      a.rvar_prev = layer > 0 ? bn_running_var[layer - 1] : nullptr;
      // launch 3 stashes a_2 in buffer 0; launch 4 starts from it and stashes a_3 in buffer 1;
      // launch 5 starts from a_3.  (Launch 3 itself runs from the points: skipping one layer
      // does not pay for a 128 B/point round trip, measured.)
      a.stash_out = (stash && (layer == 2 || layer == 3)) ? stash_buf[layer & 1] : nullptr;
      a.stash_in = (stash && layer >= 3) ? stash_buf[(layer - 1) & 1] : nullptr;
      a.partial_in_rows = 0;
      if (layer == 0 && !run_phase1) continue;  // layer-1 statistics came from the moments of the points
      if (layer == 1 && !run_phase1) a.partial_in_rows = 2;
      switch (layer) {
        case 0: rc = launch_phase<1>(a, grid, stream); break;
        case 1: rc = launch_phase<2>(a, grid, stream); break;
        case 2: rc = launch_phase<3>(a, grid, stream); break;
        case 3: rc = launch_phase<4>(a, grid, stream); break;
        default: rc = launch_phase<5>(a, grid, stream); break;
      }
      if (rc != MPA_OK) return rc;
      if (want_dbg) {  // debug: cycle stamps of CTA 0 / pipeline 0, first 8 tiles
        long long h[16 * 9];
        cudaStreamSynchronize(stream);
        cudaMemcpy(h, dbg, sizeof(h), cudaMemcpyDeviceToHost);
        for (int t = 0; t < 8; ++t) {
          if (t == 0) fprintf(stderr, "[pn phase %d] setup %lld weights %lld loop-end %lld cycles\n", layer + 1, h[128], h[129], h[130]);
          fprintf(stderr, "[pn phase %d tile %d] (issued, mma done, epilogue done) per layer:", layer + 1, t);
          for (int k = 1; k < 16; ++k) fprintf(stderr, "%s%lld", (k % 3 == 1) ? " | " : " ", h[t * 16 + k] ? h[t * 16 + k] - h[t * 16] : -1);
          fprintf(stderr, "  (next tile +%lld)\n", t < 7 ? h[(t + 1) * 16] - h[t * 16] : 0);
        }
        cudaMemset(dbg, 0, sizeof(h));
      }
    }
    {  // the last layer's statistics feed the pooling kernel
      ProfScope ps("pointnet_bn_finalize", stream);
      pointnet_finalize_kernel<<<(C[4] * 32 + 255) / 256, 256, 0, stream>>>(
          partial_buf[0], grid, 4, C[4], valids, n_parts, N, bn_gamma[4], bn_beta[4], eps, momentum,
          bn_running_mean[4], bn_running_var[4], scale, shift, bn_batch_stats);
    }
    MPA_LAUNCH_CHECK();
  } else {
    for (int layer = 0; layer < 5; ++layer) {
      pointnet_eval_affine_kernel<<<(C[layer] + 127) / 128, 128, 0, stream>>>(
          bn_gamma[layer], bn_beta[layer], bn_running_mean[layer], bn_running_var[layer], eps, layer,
          C[layer], scale, shift);
      MPA_LAUNCH_CHECK();
    }
    a.partial = partial_buf[0];
    a.partial_in = partial_buf[1];
    a.stash_in = nullptr; a.stash_out = nullptr;
    rc = launch_phase<5>(a, grid, stream);
    if (rc != MPA_OK) return rc;
  }
  {
    ProfScope ps("pointnet_pool", stream);
    pointnet_pool_kernel<<<(int)(((long long)n_parts * F + 255) / 256), 256, 0, stream>>>(
        pmax, pmin, scale, shift, valids, n_parts, F, feats);
  }
  MPA_LAUNCH_CHECK();
  return MPA_OK;
}

}  // extern "C"
