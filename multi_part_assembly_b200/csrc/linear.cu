// Dense token-level contractions of the inter-part transformer (and the pose
// head) on tcgen05 tensor cores, fed by TMA.
//
//   Y[M, N] = epilogue( X[M, K] (bf16) * W[N, K]^T (bf16) + bias[N] )
//
// X is a row-major activation matrix (tokens x features), W is an nn.Linear
// weight as torch stores it ([out, in], K-major) -- both operands are K-major,
// so TMA boxes of {64 k, 128 rows} with the 128-byte swizzle land in shared
// memory exactly in the canonical UMMA layout.  One CTA computes one 128 x BN
// output tile: warp 0 is the TMA producer, warp 1 issues tcgen05.mma into TMEM,
// warps 2-5 run the epilogue (tcgen05.ld -> bias / activation / residual ->
// global).  A 4-stage mbarrier ring overlaps the K loop.
//
// Replaces the cuBLAS GEMMs inside nn.TransformerEncoderLayer
// (models/pn_transformer/transformer.py:23-34) and PoseRegressor
// (models/modules/regressor.py:45-68).
#include <cuda.h>
#include <stdlib.h>

#include "mpa_common.cuh"
#include "tc05.cuh"

namespace mpa {

constexpr int LN_BM = 128;      // rows per tile (UMMA M)
constexpr int LN_BK = 64;       // k per stage (one 128-byte swizzle span of bf16)
constexpr int LN_EPI_WARPS = 8;  // two warps per TMEM lane quadrant, each takes half of the columns
constexpr int LN_THREADS = 64 + 32 * LN_EPI_WARPS;  // warp 0: TMA, warp 1: MMA, warps 2-9: epilogue
// SPLIT = 1: bf16 operands.  SPLIT = 3: fp32-accurate mode -- every fp32 operand is carried as
// three bf16 planes (hi, mid, lo: 3 x 8 mantissa bits) and a k-step issues the six products
// whose weight is >= 2^-16 (hi.hi, hi.mid, mid.hi, hi.lo, lo.hi, mid.mid) into the same fp32
// accumulator: the result agrees with an fp32 GEMM to ~1e-6 relative, on the tensor cores.
__host__ __device__ constexpr int ln_stage_bytes(int bn, int split = 1) { return (LN_BM + bn) * LN_BK * 2 * split; }
__host__ __device__ constexpr int ln_stages(int bn, int split) { return split == 1 ? 4 : (bn > 128 ? 1 : 2); }
__host__ __device__ constexpr int ln_smem_bytes(int bn, int split = 1) {
  return ln_stages(bn, split) * ln_stage_bytes(bn, split) + 1024;
}

enum { ACT_NONE = 0, ACT_RELU = 1, ACT_LEAKY02 = 2, ACT_SIGMOID = 3 };

// Inverted dropout (torch semantics: keep with probability 1 - p, scale kept values by
// 1 / (1 - p)) drawn from a counter-based generator: Philox4x32-10 keyed by `rng[0]` (seed) with
// the counter (element quad, row, site, rng[1] = number of forward calls so far).  The keep mask
// is written out (one byte per element) so that the backward pass applies the same mask.
struct DropoutSpec {
  const unsigned long long* rng = nullptr;  // device: two key words of this forward; nullptr = no dropout
  float p = 0.f;
  unsigned site = 0;                        // distinguishes the dropout sites of a forward
  unsigned char* mask = nullptr;            // [rows, cols] keep mask out (nullable)
};

__device__ __forceinline__ uint4 philox4x32_10(uint4 ctr, uint2 key) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const unsigned lo0 = 0xD2511F53u * ctr.x, hi0 = __umulhi(0xD2511F53u, ctr.x);
    const unsigned lo1 = 0xCD9E8D57u * ctr.z, hi1 = __umulhi(0xCD9E8D57u, ctr.z);
    ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
    key.x += 0x9E3779B9u;
    key.y += 0xBB67AE85u;
  }
  return ctr;
}
// keep decisions of the 4 elements (row, 4*quad .. 4*quad+3) of dropout site `site`
__device__ __forceinline__ uint4 dropout_quad(const unsigned long long* rng, unsigned site, unsigned row,
                                              unsigned quad, float p) {
  const unsigned long long seed = rng[0], calls = rng[1];
  const uint4 r = philox4x32_10(make_uint4(quad, row, site, (unsigned)calls),
                                make_uint2((unsigned)seed, (unsigned)(seed >> 32) ^ (unsigned)(calls >> 32)));
  const unsigned thr = (unsigned)(p * 16777216.0f);  // keep iff the top 24 bits >= p * 2^24
  return make_uint4((r.x >> 8) >= thr, (r.y >> 8) >= thr, (r.z >> 8) >= thr, (r.w >> 8) >= thr);
}
// dropout of the 32 consecutive columns col0.. of `row` held in v[]
__device__ __forceinline__ void dropout32(float* v, const DropoutSpec& d, int row, int col0, int ncols) {
  const float scale = 1.0f / (1.0f - d.p);
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const uint4 k = dropout_quad(d.rng, d.site, (unsigned)row, (unsigned)((col0 >> 2) + j), d.p);
    v[4 * j + 0] = k.x ? v[4 * j + 0] * scale : 0.f;
    v[4 * j + 1] = k.y ? v[4 * j + 1] * scale : 0.f;
    v[4 * j + 2] = k.z ? v[4 * j + 2] * scale : 0.f;
    v[4 * j + 3] = k.w ? v[4 * j + 3] * scale : 0.f;
    if (d.mask != nullptr && col0 + 4 * j + 3 < ncols)
      *reinterpret_cast<uchar4*>(d.mask + (long long)row * ncols + col0 + 4 * j) =
          make_uchar4((unsigned char)k.x, (unsigned char)k.y, (unsigned char)k.z, (unsigned char)k.w);
  }
}

struct LinearEpilogue {
  const float* bias;       // [N] or nullptr
  const float* residual;   // [M, N] fp32 added after the activation, or nullptr
  float* out_f32;          // [M, N] or nullptr
  __nv_bfloat16* out_bf16; // [M, N] or nullptr
  int act;
  // fused LayerNorm of the finished row (only when one tile spans the row: N == tile width)
  const float* ln_gamma = nullptr;
  const float* ln_beta = nullptr;
  float ln_eps = 0.f;
  // batched block-diagonal products (blockIdx.z = item): item z uses rows z * batch_rows .. of
  // BOTH operands (and of `bias`), M = N = rows of one item, output z * out_batch_stride on
  int batch_rows = 0;
  int batch_row0 = 0;  // operand row of item 0 (bias / out are passed already offset)
  long long out_batch_stride = 0;
  __nv_bfloat16* ln_out_bf16 = nullptr;  // LayerNorm(out) as the next GEMM's operand
  float* ln_out_f32 = nullptr;           // ... and/or in fp32 (final encoder norm)
  int vec = 0;                           // rows are 16-byte aligned: packed loads / stores
  DropoutSpec drop;                      // applied to act(acc + bias), before the residual add
  long long plane = 0;                   // SPLIT = 3: elements between the hi / mid / lo planes of
                                         // out_bf16 and ln_out_bf16 (0 = plain bf16 output)
};

__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, int c0, int c1,
                                            uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
      ::"r"(tc::smem_u32(smem_dst)), "l"(map), "r"(c0), "r"(c1), "r"(tc::smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(tc::smem_u32(bar)),
               "r"(bytes)
               : "memory");
}

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tc::smem_u32(bar)) : "memory");
}

__device__ __forceinline__ float apply_act(float x, int act) {
  if (act == ACT_RELU) return fmaxf(x, 0.f);
  if (act == ACT_LEAKY02) return x > 0.f ? x : 0.2f * x;
  if (act == ACT_SIGMOID) return 1.0f / (1.0f + expf(-x));
  return x;
}

// ---- epilogue building blocks -------------------------------------------------
// A TMEM lane is an output row, so after tcgen05.ld a thread holds 32 consecutive
// columns of ITS row, while a coalesced global access wants a warp instruction to
// cover whole rows.  Each epilogue warp therefore owns a [32 rows][32 cols] fp32
// staging tile in the (idle) operand ring, row pitch 36 floats:
//   put_row : thread r writes its row as 8 x 16 B       (conflict-free: pitch 144 B)
//   quad i  : lane l holds 4 consecutive columns 4*(l&7).. of row (l>>3)+4i, i.e. one
//             instruction moves 4 rows x 128 contiguous bytes to / from global memory
constexpr int EP_PITCH = 36;
constexpr int EP_TILE_FLOATS = 32 * EP_PITCH;

__device__ __forceinline__ void tile_put_row(float* tile, int lane, const float* v) {
  float4* d = reinterpret_cast<float4*>(tile + lane * EP_PITCH);
#pragma unroll
  for (int j = 0; j < 8; ++j) d[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
}
__device__ __forceinline__ void tile_get_row(const float* tile, int lane, float* v) {
  const float4* d = reinterpret_cast<const float4*>(tile + lane * EP_PITCH);
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float4 t = d[j];
    v[4 * j] = t.x; v[4 * j + 1] = t.y; v[4 * j + 2] = t.z; v[4 * j + 3] = t.w;
  }
}
__device__ __forceinline__ float4* tile_quad(float* tile, int lane, int i) {
  return reinterpret_cast<float4*>(tile + ((lane >> 3) + 4 * i) * EP_PITCH + 4 * (lane & 7));
}
// `bias` points at the CTA's shared-memory copy of the tile's bias slice (or is nullptr)
__device__ __forceinline__ void add_bias_act(float* v, const float* bias, int col0, int act) {
  if (bias != nullptr) {
    const float4* b4 = reinterpret_cast<const float4*>(bias + col0);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float4 b = b4[j];  // same address in every lane: a broadcast
      v[4 * j + 0] += b.x; v[4 * j + 1] += b.y; v[4 * j + 2] += b.z; v[4 * j + 3] += b.w;
    }
  }
  if (act != ACT_NONE) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = apply_act(v[j], act);
  }
}
__device__ __forceinline__ uint2 pack_bf16x4(float4 q) {
  const __nv_bfloat162 a = __floats2bfloat162_rn(q.x, q.y), b = __floats2bfloat162_rn(q.z, q.w);
  uint2 u;
  u.x = *reinterpret_cast<const uint32_t*>(&a);
  u.y = *reinterpret_cast<const uint32_t*>(&b);
  return u;
}

// fp32 quad -> bf16 operand: one plane, or the hi / mid / lo planes of the fp32-accurate mode
__device__ __forceinline__ void store_operand4(__nv_bfloat16* base, long long o, long long plane, float4 x) {
  *reinterpret_cast<uint2*>(base + o) = pack_bf16x4(x);
  if (plane != 0) {
    float r[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
    for (int s = 1; s < 3; ++s) {
#pragma unroll
      for (int j = 0; j < 4; ++j) r[j] -= __bfloat162float(__float2bfloat16_rn(r[j]));
      *reinterpret_cast<uint2*>(base + (long long)s * plane + o) = pack_bf16x4(make_float4(r[0], r[1], r[2], r[3]));
    }
  }
}
__device__ __forceinline__ void store_operand1(__nv_bfloat16* base, long long o, long long plane, float x) {
  __nv_bfloat16 h = __float2bfloat16_rn(x);
  base[o] = h;
  if (plane != 0) {
    x -= __bfloat162float(h);
    h = __float2bfloat16_rn(x);
    base[plane + o] = h;
    x -= __bfloat162float(h);
    base[2 * plane + o] = __float2bfloat16_rn(x);
  }
}

// One CTA = one 128 x BN output tile; 8 epilogue warps = 4 TMEM lane quadrants x 2
// column halves.  FUSE_LN (BN == N): the finished row (residual stream) is parked back
// in TMEM and LayerNorm-ed from there in two more passes (mean, then centred variance),
// the two half-row warps meeting in shared memory.
template <int BN, bool FUSE_LN, int SPLIT>
__global__ void __launch_bounds__(LN_THREADS, 1)
linear_bf16_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w,
                   int M, int N, int K, int x_plane_rows, int w_plane_rows, LinearEpilogue ep_in) {
  LinearEpilogue ep = ep_in;
  const int zoff = ep.batch_row0 + (int)blockIdx.z * ep.batch_rows;
  if (ep.batch_rows > 0) {
    if (ep.out_f32 != nullptr) ep.out_f32 += (long long)blockIdx.z * ep.out_batch_stride;
    if (ep.bias != nullptr) ep.bias += (int)blockIdx.z * ep.batch_rows;
  }
  constexpr int LN_STAGES = ln_stages(BN, SPLIT);
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t full_bar[LN_STAGES], empty_bar[LN_STAGES], done_bar;
  __shared__ uint32_t tmem_base_s;
  __shared__ float ln_part[2][2][LN_BM];
  __shared__ __align__(16) float s_vec[3][BN];  // bias, LayerNorm gamma, beta of this tile's columns
  constexpr int STAGE = ln_stage_bytes(BN, SPLIT);
  constexpr int A_TILE = LN_BM * LN_BK * 2, B_TILE = BN * LN_BK * 2;  // one plane of a stage
  static_assert(LN_EPI_WARPS * EP_TILE_FLOATS * 4 <= LN_STAGES * ln_stage_bytes(BN, SPLIT), "staging tiles live in the ring");

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * LN_BM, n0 = blockIdx.y * BN;
  const int num_k = (K + LN_BK - 1) / LN_BK;

  if (threadIdx.x == 0) {
    for (int i = 0; i < LN_STAGES; ++i) { tc::mbar_init(&full_bar[i], 1); tc::mbar_init(&empty_bar[i], 1); }
    tc::mbar_init(&done_bar, 1);
    tc::fence_barrier_init();
  }
  if (warp == 1) tc::tmem_alloc<BN>(&tmem_base_s);
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = tmem_base_s;

  if (warp == 0) {
    if (lane == 0) {  // ===== TMA producer =====
      for (int kb = 0; kb < num_k; ++kb) {
        const int s = kb % LN_STAGES;
        if (kb >= LN_STAGES) tc::mbar_wait(&empty_bar[s], ((kb / LN_STAGES) - 1) & 1);
        uint8_t* a_dst = smem + s * STAGE;            // stage = SPLIT A planes, then SPLIT B planes
        uint8_t* b_dst = a_dst + SPLIT * A_TILE;
        mbar_expect_tx(&full_bar[s], STAGE);
#pragma unroll
        for (int pl = 0; pl < SPLIT; ++pl) {  // plane pl of an operand starts pl * plane_rows rows down
          tma_load_2d(a_dst + pl * A_TILE, &map_x, kb * LN_BK, pl * x_plane_rows + zoff + m0, &full_bar[s]);
          tma_load_2d(b_dst + pl * B_TILE, &map_w, kb * LN_BK, pl * w_plane_rows + zoff + n0, &full_bar[s]);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {  // ===== MMA issuer =====
      constexpr uint32_t IDESC = tc::make_idesc_bf16(LN_BM, BN);
      for (int kb = 0; kb < num_k; ++kb) {
        const int s = kb % LN_STAGES;
        tc::mbar_wait(&full_bar[s], (kb / LN_STAGES) & 1);
        tc::fence_after_sync();
        const uint32_t a_addr = tc::smem_u32(smem + s * STAGE);
        const uint32_t b_addr = a_addr + SPLIT * A_TILE;
#pragma unroll
        for (int k = 0; k < LN_BK; k += 16) {
          if (SPLIT == 1) {
            tc::mma_bf16(tmem, tc::make_desc_sw128(a_addr + k * 2), tc::make_desc_sw128(b_addr + k * 2),
                         IDESC, (kb > 0 || k > 0) ? 1u : 0u);
          } else {
            // smallest terms first; (plane of A, plane of B)
            constexpr int PA[6] = {1, 2, 0, 1, 0, 0}, PB[6] = {1, 0, 2, 0, 1, 0};
#pragma unroll
            for (int t = 0; t < 6; ++t)
              tc::mma_bf16(tmem, tc::make_desc_sw128(a_addr + PA[t] * A_TILE + k * 2),
                           tc::make_desc_sw128(b_addr + PB[t] * B_TILE + k * 2), IDESC,
                           (kb > 0 || k > 0 || t > 0) ? 1u : 0u);
          }
        }
        tc::mma_commit(&empty_bar[s]);  // frees this stage when the MMAs above retire
      }
      tc::mma_commit(&done_bar);
    }
  } else {  // ===== epilogue: warps 2..9, TMEM lane quadrant = warp % 4 =====
    const int q = warp & 3, half = (warp - 2) >> 2;
    const int row_local = q * 32 + lane;
    const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16);
    const int c_begin = half * (BN / 2), c_end = c_begin + BN / 2;
    float* tile = reinterpret_cast<float*>(smem) + (warp - 2) * EP_TILE_FLOATS;
    const int qrow0 = m0 + q * 32 + (lane >> 3);  // global row of this lane's quad i = 0
    const int qcol = 4 * (lane & 7);
    // while the K loop runs: per-column vectors into shared memory, first residual slab into registers
    for (int c = threadIdx.x - 64; c < BN; c += LN_EPI_WARPS * 32) {
      const bool in = n0 + c < N;
      s_vec[0][c] = (ep.bias != nullptr && in) ? __ldg(ep.bias + n0 + c) : 0.f;
      if (FUSE_LN) {
        s_vec[1][c] = in ? __ldg(ep.ln_gamma + n0 + c) : 0.f;
        s_vec[2][c] = in ? __ldg(ep.ln_beta + n0 + c) : 0.f;
      }
    }
    const float* sbias = ep.bias != nullptr ? s_vec[0] : nullptr;
    float4 res[8];
    if (FUSE_LN) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int row = qrow0 + 4 * i;
        res[i] = (ep.residual != nullptr && row < M)
                     ? *reinterpret_cast<const float4*>(ep.residual + (long long)row * N + c_begin + qcol)
                     : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
    tc::group_sync(1, LN_EPI_WARPS * 32);  // s_vec complete
    tc::mbar_wait(&done_bar, 0);  // every MMA has retired: accumulator final, operand ring idle
    tc::fence_after_sync();
    if (!FUSE_LN) {
#pragma unroll 1
      for (int j0 = c_begin; j0 < c_end; j0 += 32) {
        const int col0 = n0 + j0;
        if (col0 >= N) break;  // warp-uniform
        float v[32];
        tc::tmem_ld32(taddr + (uint32_t)j0, v);
        tc::tmem_ld_wait();
        if (ep.vec && col0 + 32 <= N) {
          // the residual may alias the output (x += f(x)): fetch the whole slab before the
          // first store, or every load would be ordered behind the previous store
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int row = qrow0 + 4 * i;
            res[i] = (ep.residual != nullptr && row < M)
                         ? *reinterpret_cast<const float4*>(ep.residual + (long long)row * N + col0 + qcol)
                         : make_float4(0.f, 0.f, 0.f, 0.f);
          }
          add_bias_act(v, sbias, j0, ep.act);
          if (ep.drop.rng != nullptr && m0 + row_local < M) dropout32(v, ep.drop, m0 + row_local, col0, N);
          tile_put_row(tile, lane, v);
          __syncwarp();
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int row = qrow0 + 4 * i;
            if (row < M) {
              float4 x = *tile_quad(tile, lane, i);
              const long long o = (long long)row * N + col0 + qcol;
              x.x += res[i].x; x.y += res[i].y; x.z += res[i].z; x.w += res[i].w;
              if (ep.out_f32) *reinterpret_cast<float4*>(ep.out_f32 + o) = x;
              if (ep.out_bf16) store_operand4(ep.out_bf16, o, ep.plane, x);
            }
          }
          __syncwarp();
        } else if (m0 + row_local < M) {
          const long long rowoff = (long long)(m0 + row_local) * N;
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const int col = col0 + j;
            if (col >= N) continue;
            float x = v[j] + (ep.bias ? __ldg(ep.bias + col) : 0.f);
            x = apply_act(x, ep.act);
            if (ep.residual) x += ep.residual[rowoff + col];
            if (ep.out_f32) ep.out_f32[rowoff + col] = x;
            if (ep.out_bf16) store_operand1(ep.out_bf16, rowoff + col, ep.plane, x);
          }
        }
      }
    } else {
      // pass 1: x = act(acc + bias) + residual -> fp32 output and back into TMEM; row sum
      float sum = 0.f;
#pragma unroll 1
      for (int j0 = c_begin; j0 < c_end; j0 += 32) {
        // `res` holds this chunk's residual slab (it aliases the output, so it has to be
        // in registers before the first store); the next chunk's is fetched below, ahead
        // of this chunk's stores, and lands while this chunk is processed
        float v[32];
        tc::tmem_ld32(taddr + (uint32_t)j0, v);
        tc::tmem_ld_wait();
        add_bias_act(v, sbias, j0, ep.act);
        if (ep.drop.rng != nullptr && m0 + row_local < M) dropout32(v, ep.drop, m0 + row_local, j0, N);
        tile_put_row(tile, lane, v);
        __syncwarp();
        float4 xq[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float4 a = *tile_quad(tile, lane, i);
          xq[i] = make_float4(a.x + res[i].x, a.y + res[i].y, a.z + res[i].z, a.w + res[i].w);
        }
        if (j0 + 32 < c_end) {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int row = qrow0 + 4 * i;
            res[i] = (ep.residual != nullptr && row < M)
                         ? *reinterpret_cast<const float4*>(ep.residual + (long long)row * N + j0 + 32 + qcol)
                         : make_float4(0.f, 0.f, 0.f, 0.f);
          }
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int row = qrow0 + 4 * i;
          float4* tq = tile_quad(tile, lane, i);
          if (row < M) {
            const float4 x = xq[i];
            const long long o = (long long)row * N + j0 + qcol;
            if (ep.out_f32) *reinterpret_cast<float4*>(ep.out_f32 + o) = x;
            *tq = x;
          }
        }
        __syncwarp();
        tile_get_row(tile, lane, v);
        __syncwarp();
#pragma unroll
        for (int j = 0; j < 32; ++j) sum += v[j];
        tc::tmem_st32(taddr + (uint32_t)j0, v);
      }
      tc::tmem_st_wait();
      ln_part[0][half][row_local] = sum;
      tc::group_sync(1, LN_EPI_WARPS * 32);
      const float mean = (ln_part[0][0][row_local] + ln_part[0][1][row_local]) / (float)N;
      // pass 2: centred sum of squares
      float sq = 0.f;
#pragma unroll 1
      for (int j0 = c_begin; j0 < c_end; j0 += 32) {
        float v[32];
        tc::tmem_ld32(taddr + (uint32_t)j0, v);
        tc::tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) { const float d = v[j] - mean; sq = fmaf(d, d, sq); }
      }
      ln_part[1][half][row_local] = sq;
      tc::group_sync(1, LN_EPI_WARPS * 32);
      const float rstd = rsqrtf((ln_part[1][0][row_local] + ln_part[1][1][row_local]) / (float)N + ep.ln_eps);
      // pass 3: normalise, scale, shift
#pragma unroll 1
      for (int j0 = c_begin; j0 < c_end; j0 += 32) {
        float v[32];
        tc::tmem_ld32(taddr + (uint32_t)j0, v);
        tc::tmem_ld_wait();
        const float4* g4 = reinterpret_cast<const float4*>(s_vec[1] + j0);
        const float4* b4 = reinterpret_cast<const float4*>(s_vec[2] + j0);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 g = g4[j], b = b4[j];
          v[4 * j + 0] = (v[4 * j + 0] - mean) * rstd * g.x + b.x;
          v[4 * j + 1] = (v[4 * j + 1] - mean) * rstd * g.y + b.y;
          v[4 * j + 2] = (v[4 * j + 2] - mean) * rstd * g.z + b.z;
          v[4 * j + 3] = (v[4 * j + 3] - mean) * rstd * g.w + b.w;
        }
        tile_put_row(tile, lane, v);
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int row = qrow0 + 4 * i;
          if (row < M) {
            const float4 y = *tile_quad(tile, lane, i);
            const long long o = (long long)row * N + j0 + qcol;
            if (ep.ln_out_bf16) store_operand4(ep.ln_out_bf16, o, ep.plane, y);
            if (ep.ln_out_f32) *reinterpret_cast<float4*>(ep.ln_out_f32 + o) = y;
          }
        }
        __syncwarp();
      }
    }
    tc::fence_before_sync();
  }
  __syncthreads();
  if (warp == 1) tc::tmem_dealloc<BN>(tmem);
}

// ---- driver-API entry point for tensor maps, resolved at run time (no -lcuda) ----
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);
static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn == nullptr) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  }
  return fn;
}

// bf16 row-major [rows, cols] matrix, box = {64 cols, box_rows rows}, 128-byte swizzle, zero OOB fill
static int make_map(CUtensorMap* map, const void* base, int rows, int cols, int box_rows) {
  EncodeTiledFn fn = get_encode_fn();
  if (fn == nullptr) {
    set_error("cuTensorMapEncodeTiled is not available from the driver");
    return MPA_ERR_CUDA;
  }
  const cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  const cuuint64_t gstride[1] = {(cuuint64_t)cols * 2};
  const cuuint32_t box[2] = {(cuuint32_t)LN_BK, (cuuint32_t)box_rows};
  const cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstride, box,
                  estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (%d) for a [%d, %d] bf16 matrix", (int)r, rows, cols);
    return MPA_ERR_CUDA;
  }
  return MPA_OK;
}

static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

constexpr int LN_FUSED_WIDTH = 256;  // row width the fused-LayerNorm tile supports (d_model of the reference)

constexpr int GR_KTILE_BYTES = LN_BM * LN_BK * 2;  // one [128 x 64] bf16 operand k-block

// ---- persistent variant of the plain GEMM (no fused LayerNorm) for many output tiles ------
// The one-tile-per-CTA kernel above pays barrier init, TMEM allocation, the first TMA round trip
// and a cold epilogue for every 128 x 128 tile; with M = 512 k rows (the DGCNN / PointNet++ /
// fp32-mode GEMMs) that start-up is most of a CTA's life.  Here a CTA walks the tiles
// blockIdx.x, +gridDim.x, ...: the operand ring keeps running across tiles, the accumulator is
// double-buffered in TMEM, and the epilogue of tile t (staging tiles outside the ring) overlaps
// the TMA + MMA of tile t+1.
template <int SPLIT>
__host__ __device__ constexpr int lp_stages() { return SPLIT == 1 ? 4 : 2; }
template <int SPLIT>
__host__ __device__ constexpr int lp_pitch() { return SPLIT == 1 ? EP_PITCH : 33; }
template <int SPLIT>
__host__ __device__ constexpr int lp_smem_bytes() {
  return lp_stages<SPLIT>() * SPLIT * 2 * GR_KTILE_BYTES + LN_EPI_WARPS * 32 * lp_pitch<SPLIT>() * 4 + 1024;
}

template <int SPLIT>
__global__ void __launch_bounds__(LN_THREADS, 1)
linear_persistent_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w,
                         int M, int N, int K, int x_plane_rows, int w_plane_rows, LinearEpilogue ep) {
  constexpr int NST = lp_stages<SPLIT>();
  constexpr int A_TILE = GR_KTILE_BYTES, STAGE = SPLIT * 2 * A_TILE;  // [SPLIT A planes | SPLIT B planes]
  constexpr int PITCH = lp_pitch<SPLIT>();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  float* stage_tiles = reinterpret_cast<float*>(smem + NST * STAGE);
  __shared__ uint64_t full_bar[NST], empty_bar[NST], acc_full[2], acc_free[2];
  __shared__ uint32_t tmem_base_s;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_k = (K + LN_BK - 1) / LN_BK;
  const int tiles_n = (N + 127) / 128, tiles_m = (M + LN_BM - 1) / LN_BM;
  const int n_tiles = tiles_m * tiles_n;

  if (threadIdx.x == 0) {
    for (int i = 0; i < NST; ++i) { tc::mbar_init(&full_bar[i], 1); tc::mbar_init(&empty_bar[i], 1); }
    for (int i = 0; i < 2; ++i) { tc::mbar_init(&acc_full[i], 1); tc::mbar_init(&acc_free[i], LN_EPI_WARPS * 32); }
    tc::fence_barrier_init();
  }
  if (warp == 1) tc::tmem_alloc<256>(&tmem_base_s);
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = tmem_base_s;

  if (warp == 0) {
    if (lane == 0) {  // ===== TMA producer =====
      int it = 0;
      for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        const int m0 = (t / tiles_n) * LN_BM, n0 = (t % tiles_n) * 128;
        for (int kb = 0; kb < num_k; ++kb, ++it) {
          const int s = it % NST;
          if (it >= NST) tc::mbar_wait(&empty_bar[s], ((it / NST) - 1) & 1);
          uint8_t* a_dst = smem + s * STAGE;
          uint8_t* b_dst = a_dst + SPLIT * A_TILE;
          mbar_expect_tx(&full_bar[s], STAGE);
#pragma unroll
          for (int pl = 0; pl < SPLIT; ++pl) {
            tma_load_2d(a_dst + pl * A_TILE, &map_x, kb * LN_BK, pl * x_plane_rows + m0, &full_bar[s]);
            tma_load_2d(b_dst + pl * A_TILE, &map_w, kb * LN_BK, pl * w_plane_rows + n0, &full_bar[s]);
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {  // ===== MMA issuer =====
      constexpr uint32_t IDESC = tc::make_idesc_bf16(LN_BM, 128);
      int it = 0, tl = 0;
      for (int t = blockIdx.x; t < n_tiles; t += gridDim.x, ++tl) {
        const int buf = tl & 1;
        if (tl >= 2) tc::mbar_wait(&acc_free[buf], ((tl >> 1) - 1) & 1);
        const uint32_t acc = tmem + (uint32_t)(buf * 128);
        for (int kb = 0; kb < num_k; ++kb, ++it) {
          const int s = it % NST;
          tc::mbar_wait(&full_bar[s], (it / NST) & 1);
          tc::fence_after_sync();
          const uint32_t a_addr = tc::smem_u32(smem + s * STAGE);
          const uint32_t b_addr = a_addr + SPLIT * A_TILE;
#pragma unroll
          for (int k = 0; k < LN_BK; k += 16) {
            if (SPLIT == 1) {
              tc::mma_bf16(acc, tc::make_desc_sw128(a_addr + k * 2), tc::make_desc_sw128(b_addr + k * 2), IDESC,
                           (kb > 0 || k > 0) ? 1u : 0u);
            } else {
              constexpr int PA[6] = {1, 2, 0, 1, 0, 0}, PB[6] = {1, 0, 2, 0, 1, 0};  // smallest terms first
#pragma unroll
              for (int p6 = 0; p6 < 6; ++p6)
                tc::mma_bf16(acc, tc::make_desc_sw128(a_addr + PA[p6] * A_TILE + k * 2),
                             tc::make_desc_sw128(b_addr + PB[p6] * A_TILE + k * 2), IDESC,
                             (kb > 0 || k > 0 || p6 > 0) ? 1u : 0u);
            }
          }
          tc::mma_commit(&empty_bar[s]);
        }
        tc::mma_commit(&acc_full[buf]);
      }
    }
  } else {  // ===== epilogue: warps 2..9 =====
    const int q = warp & 3, half = (warp - 2) >> 2;
    const int row_local = q * 32 + lane;
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
    float* tile = stage_tiles + (warp - 2) * 32 * PITCH;
    const int qcol = 4 * (lane & 7);
    int tl = 0;
    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x, ++tl) {
      const int buf = tl & 1;
      const int m0 = (t / tiles_n) * LN_BM, n0 = (t % tiles_n) * 128;
      const int qrow0 = m0 + q * 32 + (lane >> 3);
      tc::mbar_wait(&acc_full[buf], (tl >> 1) & 1);
      tc::fence_after_sync();
      const uint32_t acc = tmem + (uint32_t)(buf * 128) + lane_off;
#pragma unroll 1
      for (int j0 = half * 64; j0 < half * 64 + 64; j0 += 32) {
        float v[32];
        tc::tmem_ld32(acc + (uint32_t)j0, v);
        tc::tmem_ld_wait();
        const int col0 = n0 + j0;
        if (col0 >= N) continue;  // warp-uniform
        if (ep.vec && col0 + 32 <= N) {
          float4 res[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int row = qrow0 + 4 * i;
            res[i] = (ep.residual != nullptr && row < M)
                         ? *reinterpret_cast<const float4*>(ep.residual + (long long)row * N + col0 + qcol)
                         : make_float4(0.f, 0.f, 0.f, 0.f);
          }
          if (ep.bias != nullptr) {  // the warp's 32 columns: broadcast 16-byte loads
#pragma unroll
            for (int j4 = 0; j4 < 8; ++j4) {
              const float4 b4 = __ldg(reinterpret_cast<const float4*>(ep.bias + col0) + j4);
              v[4 * j4] += b4.x; v[4 * j4 + 1] += b4.y; v[4 * j4 + 2] += b4.z; v[4 * j4 + 3] += b4.w;
            }
          }
          if (ep.act != ACT_NONE) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = apply_act(v[j], ep.act);
          }
          if (ep.drop.rng != nullptr && m0 + row_local < M) dropout32(v, ep.drop, m0 + row_local, col0, N);
          if (SPLIT == 1) {
            tile_put_row(tile, lane, v);
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) tile[lane * PITCH + j] = v[j];
          }
          __syncwarp();
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int row = qrow0 + 4 * i;
            float4 x;
            if (SPLIT == 1) {
              x = *tile_quad(tile, lane, i);
            } else {
              const float* src = tile + ((lane >> 3) + 4 * i) * PITCH + qcol;
              x = make_float4(src[0], src[1], src[2], src[3]);
            }
            if (row < M) {
              const long long o = (long long)row * N + col0 + qcol;
              x.x += res[i].x; x.y += res[i].y; x.z += res[i].z; x.w += res[i].w;
              if (ep.out_f32) *reinterpret_cast<float4*>(ep.out_f32 + o) = x;
              if (ep.out_bf16) store_operand4(ep.out_bf16, o, ep.plane, x);
            }
          }
          __syncwarp();
        } else if (m0 + row_local < M) {
          const long long rowoff = (long long)(m0 + row_local) * N;
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const int col = col0 + j;
            if (col >= N) continue;
            float x = v[j] + (ep.bias ? __ldg(ep.bias + col) : 0.f);
            x = apply_act(x, ep.act);
            if (ep.residual) x += ep.residual[rowoff + col];
            if (ep.out_f32) ep.out_f32[rowoff + col] = x;
            if (ep.out_bf16) store_operand1(ep.out_bf16, rowoff + col, ep.plane, x);
          }
        }
      }
      tc::fence_before_sync();
      mbar_arrive(&acc_free[buf]);
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 1) tc::tmem_dealloc<256>(tmem);
}

// x: [split, M, K] bf16 planes, w: [split, N, K] (split = 1: plain bf16 operands)
int launch_linear(const __nv_bfloat16* x, const __nv_bfloat16* w, int M, int N, int K,
                  LinearEpilogue ep, const char* name, cudaStream_t stream, int split = 1) {
  MPA_CHECK_ARG(K % 8 == 0, "linear: K must be a multiple of 8 (got %d)", K);
  MPA_CHECK_ARG(split == 1 || split == 3, "linear: split must be 1 or 3");
  const bool fuse_ln = ep.ln_gamma != nullptr;
  ep.vec = (N % 8 == 0) && aligned16(ep.bias) && aligned16(ep.residual) && aligned16(ep.out_f32) &&
           aligned16(ep.out_bf16) && aligned16(ep.ln_gamma) && aligned16(ep.ln_beta) &&
           aligned16(ep.ln_out_bf16) && aligned16(ep.ln_out_f32) && (ep.plane % 8 == 0);
  if (fuse_ln)
    MPA_CHECK_ARG(N == LN_FUSED_WIDTH && ep.vec && ep.ln_beta != nullptr,
                  "linear: fused LayerNorm needs N == %d and 16-byte aligned rows", LN_FUSED_WIDTH);
  const int bn = fuse_ln ? LN_FUSED_WIDTH : 128;
  CUtensorMap mx, mw;
  int rc = make_map(&mx, x, split * M, K, LN_BM);
  if (rc != MPA_OK) return rc;
  rc = make_map(&mw, w, split * N, K, bn);
  if (rc != MPA_OK) return rc;
  static DeviceOnce attr;
  if (attr.pending()) {
    MPA_CUDA(cudaFuncSetAttribute(linear_bf16_kernel<128, false, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  ln_smem_bytes(128, 1)));
    MPA_CUDA(cudaFuncSetAttribute(linear_bf16_kernel<LN_FUSED_WIDTH, true, 1>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, ln_smem_bytes(LN_FUSED_WIDTH, 1)));
    MPA_CUDA(cudaFuncSetAttribute(linear_bf16_kernel<128, false, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  ln_smem_bytes(128, 3)));
    MPA_CUDA(cudaFuncSetAttribute(linear_bf16_kernel<LN_FUSED_WIDTH, true, 3>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, ln_smem_bytes(LN_FUSED_WIDTH, 3)));
    attr.done();
  }
  dim3 grid((M + LN_BM - 1) / LN_BM, (N + bn - 1) / bn);
  // many tiles and no fused LayerNorm: the persistent kernel (every CTA walks several tiles)
  static const bool no_persist = getenv("MPA_LINEAR_ONE_TILE") != nullptr;  // A/B switch
  if (!fuse_ln && !no_persist && (long long)grid.x * grid.y >= 2ll * device_sms() && (!ep.vec || aligned16(ep.bias))) {
    static DeviceOnce attr_p;
    if (attr_p.pending()) {
      MPA_CUDA(cudaFuncSetAttribute(linear_persistent_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    lp_smem_bytes<1>()));
      MPA_CUDA(cudaFuncSetAttribute(linear_persistent_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    lp_smem_bytes<3>()));
      attr_p.done();
    }
    {
      ProfScope ps(name, stream);
      if (split == 1)
        linear_persistent_kernel<1><<<device_sms(), LN_THREADS, lp_smem_bytes<1>(), stream>>>(mx, mw, M, N, K, M, N, ep);
      else
        linear_persistent_kernel<3><<<device_sms(), LN_THREADS, lp_smem_bytes<3>(), stream>>>(mx, mw, M, N, K, M, N, ep);
    }
    MPA_LAUNCH_CHECK();
    return MPA_OK;
  }
  {
    ProfScope ps(name, stream);
    if (split == 1) {
      if (fuse_ln)
        linear_bf16_kernel<LN_FUSED_WIDTH, true, 1><<<grid, LN_THREADS, ln_smem_bytes(LN_FUSED_WIDTH, 1), stream>>>(
            mx, mw, M, N, K, M, N, ep);
      else
        linear_bf16_kernel<128, false, 1><<<grid, LN_THREADS, ln_smem_bytes(128, 1), stream>>>(mx, mw, M, N, K, M,
                                                                                             N, ep);
    } else {
      if (fuse_ln)
        linear_bf16_kernel<LN_FUSED_WIDTH, true, 3><<<grid, LN_THREADS, ln_smem_bytes(LN_FUSED_WIDTH, 3), stream>>>(
            mx, mw, M, N, K, M, N, ep);
      else
        linear_bf16_kernel<128, false, 3><<<grid, LN_THREADS, ln_smem_bytes(128, 3), stream>>>(mx, mw, M, N, K, M,
                                                                                             N, ep);
    }
  }
  MPA_LAUNCH_CHECK();
  return MPA_OK;
}

// ---- persistent Gram kernel (k-NN scoring) ---------------------------------------------
// out[z][i][j] ~ x_{z,i} . x_{z,j} + bias[z R + j] for the R rows of every item z: two bf16 planes
// per operand (hi, mid), three plane products, smallest first -- a filter with a known error bound.  K <= 128, so a whole
// [128 x K] operand tile is resident: a CTA walks a contiguous range of [128 x 128] output
// tiles, keeps the row-tile operand while the column tile changes, and double-buffers the
// accumulator in TMEM so that the TMA load + MMAs of the next tile run under the epilogue of
// the current one -- the kernel is bound by the 64 KB fp32 write of each tile, not by per-CTA
// start-up as the one-tile-per-CTA GEMM is.
constexpr int GR_KTILE = LN_BM * LN_BK * 2;  // [128 x 64] bf16 = 16 KB
constexpr int GR_NB = 3;                     // column-tile k-blocks in flight
constexpr int GR_NPL = 2;                    // operand planes (hi, mid): products hi.hi + hi.mid + mid.hi, the
                                             // dropped ones are <= 3 * 2^-18 |x_i| |x_j| (the caller's delta covers them)
// staging tile of an epilogue warp: [32][36] floats (16-byte accesses) when it fits, [32][33]
// (scalar accesses, conflict-free) for K = 128, where the operands leave 33 KB
template <int KB>
__host__ __device__ constexpr int gr_pitch() { return KB == 1 ? EP_PITCH : 33; }
template <int KB>
__host__ __device__ constexpr int gr_smem_bytes() {
  // staging / tables: the slab mode needs 8 warp tiles, the candidate mode 128 x 33 group maxima +
  // thresholds + counters + 1024 biases + 256 x 33 value strips = 55.8 KB
  return GR_NPL * KB * GR_KTILE + GR_NB * GR_NPL * GR_KTILE + 56 * 1024 + 1024;
}
static_assert(gr_smem_bytes<2>() <= 227 * 1024, "K = 128 operands + staging fit one SM");

// CAND mode (k-NN without the score slab): every row tile is swept twice over its column tiles.
// Sweep 0 keeps, per row, the maxima of the 32-column groups; the k-th largest of those is a lower
// bound of the row's k-th best score (k distinct elements reach it).  Sweep 1 repeats the products
// (the tensor pipe has the time: the kernel was bound by writing 64 KB per tile) and appends every
// score within `delta` of that bound to the row's candidate list (<= 64 keys per row in global
// memory; the count says when a row overflowed).  `knn_finish_kernel` (csrc/knn.cu) then sorts 64
// keys per row instead of reading 4 KB of scores.
struct GramCand {
  const float* xx;          // [items * R] |x|^2 of the rows (for delta)
  const unsigned* xxmax;    // [items] max |x|^2 of the item, as ordered uint bits
  float alpha;              // delta = alpha * (|x_i|^2 + max_j |x_j|^2)
  int k;
  unsigned long long* cand; // [items * R][64] keys (score, ~column)
  int* count;               // [items * R]
};
__device__ __forceinline__ unsigned long long gram_key(float v, int j) {  // = knn_key of csrc/knn.cu
  const unsigned b = __float_as_uint(v);
  const unsigned o = (b & 0x80000000u) ? ~b : (b | 0x80000000u);
  return ((unsigned long long)o << 32) | (unsigned long long)(~(unsigned)j);
}
constexpr int GC_PITCH = 33;  // floats per row of the group-maxima table (conflict-free per-row walks)

template <int KB, bool CAND>
__global__ void __launch_bounds__(LN_THREADS, 1)
gram_scores_kernel(const __grid_constant__ CUtensorMap map_x, int plane_rows, int R, int row0, int items,
                   const float* __restrict__ bias, const float* __restrict__ item_valid, float* __restrict__ out,
                   GramCand gc) {
  constexpr int NSWEEP = CAND ? 2 : 1;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* As = smem;                                   // [planes][KB][128 x 64]: row tile, all of K
  uint8_t* Bs = smem + GR_NPL * KB * GR_KTILE;          // ring of GR_NB x [planes][128 x 64]: column-tile k-blocks
  float* stage = reinterpret_cast<float*>(Bs + GR_NB * GR_NPL * GR_KTILE);
  constexpr int PITCH = gr_pitch<KB>();
  __shared__ uint64_t a_full, b_full[GR_NB], b_empty[GR_NB], acc_full[2], acc_free[2];
  __shared__ uint32_t tmem_base_s;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int T = (R + LN_BM - 1) / LN_BM;
  const int n_rt = items * T;  // row tiles; a CTA takes every gridDim.x-th one with all its column tiles

  if (threadIdx.x == 0) {
    tc::mbar_init(&a_full, 1);
    for (int i = 0; i < GR_NB; ++i) { tc::mbar_init(&b_full[i], 1); tc::mbar_init(&b_empty[i], 1); }
    for (int i = 0; i < 2; ++i) { tc::mbar_init(&acc_full[i], 1); tc::mbar_init(&acc_free[i], LN_EPI_WARPS * 32); }
    tc::fence_barrier_init();
  }
  if (warp == 1) tc::tmem_alloc<256>(&tmem_base_s);
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = tmem_base_s;

  if (warp == 0) {
    if (lane == 0) {  // ===== TMA producer =====
      int prev_rt = -1, u = 0;
      for (int rt = blockIdx.x; rt < n_rt; rt += gridDim.x) for (int sw = 0; sw < NSWEEP; ++sw) for (int nj = 0; nj < T; ++nj) {
        const int z = rt / T, mi = rt - z * T;          // rt = z * T + mi
        if (item_valid != nullptr && item_valid[z] == 0.0f) continue;  // padded item: every role skips its tiles
        const int arow = row0 + z * R + mi * LN_BM, brow = row0 + z * R + nj * LN_BM;
        for (int kb = 0; kb < KB; ++kb, ++u) {
          const int s = u % GR_NB;
          if (u >= GR_NB) tc::mbar_wait(&b_empty[s], ((u / GR_NB) - 1) & 1);
          if (kb == 0 && rt != prev_rt) {
            // the MMAs that read the old row tile are those of the k-block uses < u: in order,
            // so the last one's commit covers them all
            if (u > 0) tc::mbar_wait(&b_empty[(u - 1) % GR_NB], ((u - 1) / GR_NB) & 1);
            mbar_expect_tx(&a_full, GR_NPL * KB * GR_KTILE);
            for (int pl = 0; pl < GR_NPL; ++pl)
              for (int k2 = 0; k2 < KB; ++k2)
                tma_load_2d(As + (pl * KB + k2) * GR_KTILE, &map_x, k2 * LN_BK, pl * plane_rows + arow, &a_full);
            prev_rt = rt;
          }
          mbar_expect_tx(&b_full[s], GR_NPL * GR_KTILE);
          for (int pl = 0; pl < GR_NPL; ++pl)
            tma_load_2d(Bs + (s * GR_NPL + pl) * GR_KTILE, &map_x, kb * LN_BK, pl * plane_rows + brow, &b_full[s]);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {  // ===== MMA issuer =====
      constexpr uint32_t IDESC = tc::make_idesc_bf16(LN_BM, 128);
      constexpr int NPROD = 3, PA[NPROD] = {1, 0, 0}, PB[NPROD] = {0, 1, 0};  // (plane of A, plane of B), smallest first
      int prev_rt = -1, u = 0, a_cnt = 0, tl = -1;
      for (int rt = blockIdx.x; rt < n_rt; rt += gridDim.x) for (int sw = 0; sw < NSWEEP; ++sw) for (int nj = 0; nj < T; ++nj) {
        if (item_valid != nullptr && item_valid[rt / T] == 0.0f) continue;
        ++tl;  // tiles actually processed: accumulator buffer and barrier phases follow this count
        const int buf = tl & 1;
        if (rt != prev_rt) {
          tc::mbar_wait(&a_full, a_cnt & 1);
          ++a_cnt;
          prev_rt = rt;
        }
        if (tl >= 2) tc::mbar_wait(&acc_free[buf], ((tl >> 1) - 1) & 1);
        const uint32_t acc = tmem + (uint32_t)(buf * 128);
        for (int kb = 0; kb < KB; ++kb, ++u) {
          const int s = u % GR_NB;
          tc::mbar_wait(&b_full[s], (u / GR_NB) & 1);
          tc::fence_after_sync();
          const uint32_t a_addr = tc::smem_u32(As), b_addr = tc::smem_u32(Bs + s * GR_NPL * GR_KTILE);
#pragma unroll
          for (int k = 0; k < LN_BK; k += 16) {
#pragma unroll
            for (int p6 = 0; p6 < NPROD; ++p6)
              tc::mma_bf16(acc, tc::make_desc_sw128(a_addr + (PA[p6] * KB + kb) * GR_KTILE + k * 2),
                           tc::make_desc_sw128(b_addr + PB[p6] * GR_KTILE + k * 2), IDESC,
                           (kb > 0 || k > 0 || p6 > 0) ? 1u : 0u);
          }
          tc::mma_commit(&b_empty[s]);
        }
        tc::mma_commit(&acc_full[buf]);
      }
    }
  } else {  // ===== epilogue: warps 2..9 =====
    const int q = warp & 3, half = (warp - 2) >> 2;
    const int row_local = q * 32 + lane;
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
    float* tile = stage + (warp - 2) * 32 * PITCH;
    const bool vec = (R & 3) == 0;
    int tl = -1;
    // CAND: per-row tables in the (unused) staging area
    float* s_mx = stage;                                           // [128][GC_PITCH] group maxima
    float* s_thr = stage + LN_BM * GC_PITCH;                       // [128]
    int* s_cnt = reinterpret_cast<int*>(s_thr + LN_BM);            // [128]
    float* s_bias = reinterpret_cast<float*>(s_cnt + LN_BM);       // [8 * 128] the item's column bias
    float* s_vst = s_bias + 8 * LN_BM;                             // [256][GC_PITCH] per-thread strips
    const float ninf = -__int_as_float(0x7f800000);
    for (int rt = blockIdx.x; rt < n_rt; rt += gridDim.x) for (int sw = 0; sw < NSWEEP; ++sw) for (int nj = 0; nj < T; ++nj) {
      const int z = rt / T, mi = rt - z * T;
      if (item_valid != nullptr && item_valid[z] == 0.0f) continue;
      if (CAND && sw == 0 && nj == 0) {
        // new row tile: clear this thread's half of the row's group maxima, stage the item's bias
#pragma unroll
        for (int gidx = 0; gidx < 16; ++gidx) s_mx[row_local * GC_PITCH + half * 16 + gidx] = ninf;
        for (int c = threadIdx.x - 64; c < T * LN_BM; c += LN_EPI_WARPS * 32)
          s_bias[c] = (bias != nullptr && c < R) ? __ldg(bias + (long long)z * R + c) : 0.f;
        tc::group_sync(1, LN_EPI_WARPS * 32);
      }
      ++tl;
      const int buf = tl & 1;
      const int m0 = mi * LN_BM, n0 = nj * LN_BM;
      tc::mbar_wait(&acc_full[buf], (tl >> 1) & 1);
      tc::fence_after_sync();
      const uint32_t acc = tmem + (uint32_t)(buf * 128) + lane_off;
      float* obase = out + (long long)z * R * R;
      if (CAND) {
        // this thread's 64 columns of its row: both 32-column groups at once
        float v[64];
        tc::tmem_ld32(acc + (uint32_t)(half * 64), v);
        tc::tmem_ld32(acc + (uint32_t)(half * 64 + 32), v + 32);
        tc::tmem_ld_wait();
        const int colb = n0 + half * 64;
        const float* sb = s_bias + colb;
        if (sw == 0) {
#pragma unroll
          for (int gq = 0; gq < 2; ++gq) {
            float m8[8] = {ninf, ninf, ninf, ninf, ninf, ninf, ninf, ninf};  // independent chains
#pragma unroll
            for (int j4 = 0; j4 < 8; ++j4) {
              const float4 b4 = *reinterpret_cast<const float4*>(sb + 32 * gq + 4 * j4);
              const float bq[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
              for (int e4 = 0; e4 < 4; ++e4) {
                const int j = 4 * j4 + e4;
                const float t = v[32 * gq + j] + bq[e4];
                m8[j & 7] = (colb + 32 * gq + j < R) ? fmaxf(m8[j & 7], t) : m8[j & 7];
              }
            }
            const float m = fmaxf(fmaxf(fmaxf(m8[0], m8[1]), fmaxf(m8[2], m8[3])),
                                  fmaxf(fmaxf(m8[4], m8[5]), fmaxf(m8[6], m8[7])));
            const int gidx = 2 * nj + gq;  // group index inside this half: <= 15 for T <= 8
            s_mx[row_local * GC_PITCH + half * 16 + gidx] = m;
          }
        } else if (m0 + row_local < R) {
          // ~1 of 32 scores passes: one shared-memory atomic per thread and group reserves the
          // slots, the stores are predicated (a branch per element would serialise the warp)
          const float thr = s_thr[row_local];
          unsigned long long* cl = gc.cand + ((long long)z * R + m0 + row_local) * 64;
          float* vst = s_vst + (threadIdx.x - 64) * GC_PITCH;
#pragma unroll
          for (int gq = 0; gq < 2; ++gq) {
            unsigned pass = 0u;
#pragma unroll
            for (int j4 = 0; j4 < 8; ++j4) {
              const float4 b4 = *reinterpret_cast<const float4*>(sb + 32 * gq + 4 * j4);
              const float bq[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
              for (int e4 = 0; e4 < 4; ++e4) {
                const int j = 4 * j4 + e4;
                v[32 * gq + j] += bq[e4];
                pass |= (colb + 32 * gq + j < R && v[32 * gq + j] >= thr) ? (1u << j) : 0u;
              }
            }
            if (pass != 0u) {
              // the group goes to this thread's private strip of shared memory, so that the few
              // set bits can be walked with a dynamic index (32 tests + branches per group were
              // a third of the kernel's stall samples)
#pragma unroll
              for (int j = 0; j < 32; ++j) vst[j] = v[32 * gq + j];
              int slot = atomicAdd(&s_cnt[row_local], __popc(pass));
              while (pass != 0u) {
                const int j = __ffs(pass) - 1;
                pass &= pass - 1u;
                if (slot < 64) cl[slot] = gram_key(vst[j], colb + 32 * gq + j);
                ++slot;
              }
            }
          }
        }
      }
#pragma unroll 1
      for (int j0 = half * 64; !CAND && j0 < half * 64 + 64; j0 += 32) {
        float v[32];
        tc::tmem_ld32(acc + (uint32_t)j0, v);
        tc::tmem_ld_wait();
        const int col0 = n0 + j0;
        if (col0 >= R) continue;  // warp-uniform
        if (bias != nullptr) {  // the warp's 32 columns: the same address in every lane (broadcast loads)
          const float* bz = bias + (long long)z * R + col0;
          if (vec && col0 + 32 <= R && (reinterpret_cast<uintptr_t>(bz) & 15) == 0) {
#pragma unroll
            for (int j4 = 0; j4 < 8; ++j4) {
              const float4 b4 = __ldg(reinterpret_cast<const float4*>(bz) + j4);
              v[4 * j4] += b4.x; v[4 * j4 + 1] += b4.y; v[4 * j4 + 2] += b4.z; v[4 * j4 + 3] += b4.w;
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] += (col0 + j < R) ? __ldg(bz + j) : 0.f;
          }
        }
        if (vec && col0 + 32 <= R) {
          // through the warp's staging tile: one instruction then moves 4 rows x 128 contiguous bytes
          const int qrow0 = m0 + q * 32 + (lane >> 3), qcol = 4 * (lane & 7);
          if (KB == 1) {
            tile_put_row(tile, lane, v);
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) tile[lane * PITCH + j] = v[j];
          }
          __syncwarp();
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int row = qrow0 + 4 * i;
            float4 o4;
            if (KB == 1) {
              o4 = *tile_quad(tile, lane, i);
            } else {
              const float* src = tile + ((lane >> 3) + 4 * i) * PITCH + qcol;
              o4 = make_float4(src[0], src[1], src[2], src[3]);
            }
            if (row < R) *reinterpret_cast<float4*>(obase + (long long)row * R + col0 + qcol) = o4;
          }
          __syncwarp();
        } else if (m0 + row_local < R) {
          float* orow = obase + (long long)(m0 + row_local) * R;
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (col0 + j < R) orow[col0 + j] = v[j];
        }
      }
      tc::fence_before_sync();
      mbar_arrive(&acc_free[buf]);
      if (CAND && nj == T - 1) {
        tc::group_sync(1, LN_EPI_WARPS * 32);  // the sweep's tables are complete
        if (sw == 0) {
          // k-th largest of every row's 32 group maxima: warp w sorts the rows 16 w .. 16 w + 15,
          // one value per lane, bitonic network by shuffles (descending)
          {
            const int w8 = warp - 2;
#pragma unroll 4
            for (int rr = 0; rr < 16; ++rr) {
              const int rloc = w8 * 16 + rr;
              float a = s_mx[rloc * GC_PITCH + lane];
#pragma unroll
              for (int kk = 2; kk <= 32; kk <<= 1) {
#pragma unroll
                for (int jj = kk >> 1; jj > 0; jj >>= 1) {
                  const float pa = __shfl_xor_sync(0xffffffffu, a, jj);
                  const bool lower = (lane & jj) == 0;
                  const bool desc = (lane & kk) == 0 || kk == 32;
                  a = (lower == desc) ? fmaxf(a, pa) : fminf(a, pa);
                }
              }
              if (lane == gc.k - 1) {
                const long long grow = (long long)z * R + m0 + rloc;
                float thr = ninf;
                if (m0 + rloc < R && a > ninf) thr = a - gc.alpha * (gc.xx[grow] + __uint_as_float(gc.xxmax[z]));
                s_thr[rloc] = thr;
                s_cnt[rloc] = 0;
              }
            }
          }
        } else if (half == 0 && m0 + row_local < R) {
          gc.count[(long long)z * R + m0 + row_local] = s_cnt[row_local];
        }
        tc::group_sync(1, LN_EPI_WARPS * 32);
      }
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 1) tc::tmem_dealloc<256>(tmem);
}

// Batched Gram products in the fp32-accurate mode: for every item z of [z0, z0 + items) the
// [R x R] matrix  out[z - z0][i][j] = x_{z,i} . x_{z,j} + bias[z * R + j]  of the item's R rows
// of x ([3 planes][total_rows][K] bf16).  Used by the k-NN scoring (csrc/knn.cu).
int launch_gram_batched(const __nv_bfloat16* x_planes, long long total_rows, int R, int K, int z0, int items,
                        const float* bias, const float* item_valid, float* out, const char* name,
                        cudaStream_t stream) {
  MPA_CHECK_ARG(K % 8 == 0 && K <= 2 * LN_BK && total_rows * GR_NPL < (1ll << 31),
                "gram: K %% 8 == 0, K <= 128 and planes * rows < 2^31");
  CUtensorMap mx;
  int rc = make_map(&mx, x_planes, (int)(GR_NPL * total_rows), K, LN_BM);
  if (rc != MPA_OK) return rc;
  static DeviceOnce attr2;
  if (attr2.pending()) {
    MPA_CUDA(cudaFuncSetAttribute(gram_scores_kernel<1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, gr_smem_bytes<1>()));
    MPA_CUDA(cudaFuncSetAttribute(gram_scores_kernel<2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, gr_smem_bytes<2>()));
    attr2.done();
  }
  const int T = (R + LN_BM - 1) / LN_BM;
  int ctas = device_sms();
  if (ctas > items * T) ctas = items * T;
  const float* b = bias != nullptr ? bias + (long long)z0 * R : nullptr;
  const float* iv = item_valid != nullptr ? item_valid + z0 : nullptr;  // [items]: 0 = skip the item
  {
    ProfScope ps(name, stream);
    if (K <= LN_BK)
      gram_scores_kernel<1, false><<<ctas, LN_THREADS, gr_smem_bytes<1>(), stream>>>(mx, (int)total_rows, R, z0 * R,
                                                                                     items, b, iv, out, GramCand{});
    else
      gram_scores_kernel<2, false><<<ctas, LN_THREADS, gr_smem_bytes<2>(), stream>>>(mx, (int)total_rows, R, z0 * R,
                                                                                     items, b, iv, out, GramCand{});
  }
  MPA_LAUNCH_CHECK();
  return MPA_OK;
}

// k-NN candidate extraction for ALL items in one launch (no score slab): see GramCand.
// Returns MPA_ERR_ARG-free "not supported" as 1 when K > 128 (the caller falls back to the slab path).
int launch_gram_candidates(const __nv_bfloat16* x_planes, long long total_rows, int R, int K, int items,
                           const float* bias, const float* item_valid, const float* xx, const unsigned* xxmax,
                           float alpha, int k, unsigned long long* cand, int* count, cudaStream_t stream) {
  if (K > 2 * LN_BK || (R + LN_BM - 1) / LN_BM > 8) return 1;
  MPA_CHECK_ARG(K % 8 == 0 && total_rows * GR_NPL < (1ll << 31), "gram: K %% 8 == 0 and planes * rows < 2^31");
  CUtensorMap mx;
  int rc = make_map(&mx, x_planes, (int)(GR_NPL * total_rows), K, LN_BM);
  if (rc != MPA_OK) return rc;
  static DeviceOnce attr;
  if (attr.pending()) {
    MPA_CUDA(cudaFuncSetAttribute(gram_scores_kernel<1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, gr_smem_bytes<1>()));
    MPA_CUDA(cudaFuncSetAttribute(gram_scores_kernel<2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, gr_smem_bytes<2>()));
    attr.done();
  }
  const int T = (R + LN_BM - 1) / LN_BM;
  int ctas = device_sms();
  if (ctas > items * T) ctas = items * T;
  GramCand gc{xx, xxmax, alpha, k, cand, count};
  {
    ProfScope ps("knn_gram", stream);
    if (K <= LN_BK)
      gram_scores_kernel<1, true><<<ctas, LN_THREADS, gr_smem_bytes<1>(), stream>>>(mx, (int)total_rows, R, 0, items, bias,
                                                                                    item_valid, nullptr, gc);
    else
      gram_scores_kernel<2, true><<<ctas, LN_THREADS, gr_smem_bytes<2>(), stream>>>(mx, (int)total_rows, R, 0, items, bias,
                                                                                    item_valid, nullptr, gc);
  }
  MPA_LAUNCH_CHECK();
  return MPA_OK;
}

// ======================================================================================
// Fused second half of a pre-LN encoder layer for d_model = 256 (bf16 operands):
//     x  <- x + dropout1(att W_o^T + b_o)            (out_proj)
//     h  <- LayerNorm2(x)
//     x  <- x + dropout2(dropout(relu(h W_1^T + b_1)) W_2^T + b_2)      (FFN)
//     xn <- LayerNorm_next(x)        (norm1 of the next layer, or the encoder's final norm)
// One CTA owns 128 token rows for the whole chain, so the residual stream, LayerNorm2's
// output and the FFN hidden activations never leave the SM: the att tile and LN2(x) are A
// operands in shared memory (K-major, 128-byte swizzle, written by the epilogue warps), the
// hidden layer is produced 128 columns at a time into a double-buffered A operand while the
// previous chunk is already being contracted with W_2, and the accumulators live in TMEM
// (256 columns for the two 256-wide results, 2 x 128 for the hidden chunks).  Weights stream
// through a 2-stage TMA ring of 32 KB tiles.  Replaces three GEMM launches per layer
// (out_proj, linear1, linear2 of nn.TransformerEncoderLayer, transformer.py:23-34).
constexpr int FB_D = 256;           // d_model
constexpr int FB_CH = 128;          // hidden columns per chunk
constexpr int FB_STAGE = 32 * 1024; // one ring stage: a [256 x 64] weight tile or two [128 x 64] tiles
constexpr int FB_NST = 2;
constexpr int FB_KTILE = LN_BM * LN_BK * 2;  // [128 x 64] bf16 A-operand k-block = 16 KB
constexpr int FB_SMEM = 2 * 4 * FB_KTILE + FB_NST * FB_STAGE + 1024;
constexpr int FB_EPI = LN_EPI_WARPS * 32;    // 256 epilogue threads
constexpr int FB_MAX_FF = 1024;              // hidden width whose bias fits the static shared memory

struct FfnBlockArgs {
  const float* b_o; const float* b1; const float* b2;
  const float* ln2_g; const float* ln2_b; const float* lnn_g; const float* lnn_b;  // lnn_* may be null
  const float* x_in;        // [M, 256] residual stream in
  float* x_out;             // [M, 256] residual stream out (may alias x_in); null: not stored
  __nv_bfloat16* xn_out;    // [M, 256] LayerNorm_next(x) as the next GEMM's operand (nullable)
  float* out_f32;           // [M, 256] fp32 result: LayerNorm_next(x) if lnn_g else x (nullable)
  int M, FF;
  float eps;
  DropoutSpec drop1, drop_h, drop2;
  long long* dbg = nullptr;  // MPA_FFN_DEBUG: cycle stamps of CTA 0 (see tools/profile_transformer.py)
  // hidden-dimension split over a thread-block cluster (few token tiles: 640 tokens are 5 tiles
  // on 148 SMs): `cl` CTAs share a tile, each contracts FF / cl hidden columns and leaves a
  // partial [128 x 256] result in `part`; after the cluster barrier every CTA sums, finishes
  // and normalises its own 128 / cl rows.  cl == 1: one CTA does the whole tile.
  int cl = 1;
  float* part = nullptr;     // [tiles, cl, 128, 256] fp32 partial FFN outputs
  float* x1s = nullptr;      // [tiles, 128, 256] residual stream after attention (owner rows)
};
#define FB_STAMP(i) do { if (a.dbg != nullptr && blockIdx.x == 0) a.dbg[i] = clock64(); } while (0)


// LayerNorm of the 128 rows parked in TMEM columns [0, 256).  The caller accumulated the row
// sums and sums of squares of its half while parking (one-pass statistics: the residual stream
// is O(1) with |mean| << std, and the bf16 operand this feeds has 8 mantissa bits); the two
// half-row warps of a row meet in shared memory.  `emit(j0, v)` receives 32 normalised columns
// j0.. of this thread's row.
template <typename Emit>
__device__ __forceinline__ void fb_layernorm_rows(uint32_t taddr, int c_begin, int c_end, int half, int row_local,
                                                  float sum, float sumsq, float (*ln_part)[2][LN_BM],
                                                  const float* g, const float* b, float eps, Emit emit) {
  ln_part[0][half][row_local] = sum;
  ln_part[1][half][row_local] = sumsq;
  tc::group_sync(1, FB_EPI);
  const float mean = (ln_part[0][0][row_local] + ln_part[0][1][row_local]) * (1.0f / FB_D);
  const float ex2 = (ln_part[1][0][row_local] + ln_part[1][1][row_local]) * (1.0f / FB_D);
  const float rstd = rsqrtf(fmaxf(ex2 - mean * mean, 0.f) + eps);
#pragma unroll 1
  for (int j0 = c_begin; j0 < c_end; j0 += 32) {
    float v[32];
    tc::tmem_ld32(taddr + (uint32_t)j0, v);
    tc::tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = (v[j] - mean) * rstd * g[j0 + j] + b[j0 + j];
    emit(j0, v);
  }
  tc::group_sync(1, FB_EPI);  // ln_part may be reused
}

// 32 fp32 values of one row -> bf16 into a K-major SWIZZLE_128B A operand made of [128 x 64]
// k-block tiles (`tile0` = k-block holding column col0's first element)
__device__ __forceinline__ void fb_store_operand_row(uint8_t* tiles, int row, int col0, const float* v) {
#pragma unroll
  for (int j = 0; j < 4; ++j) {  // four 16-byte chunks of 8 bf16
    const int col = col0 + 8 * j;
    uint8_t* dst = tiles + (col >> 6) * FB_KTILE + tc::sw128_offset(row, col & 63);
    const uint2 lo = pack_bf16x4(make_float4(v[8 * j], v[8 * j + 1], v[8 * j + 2], v[8 * j + 3]));
    const uint2 hi = pack_bf16x4(make_float4(v[8 * j + 4], v[8 * j + 5], v[8 * j + 6], v[8 * j + 7]));
    *reinterpret_cast<uint4*>(dst) = make_uint4(lo.x, lo.y, hi.x, hi.y);
  }
}

__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// Cluster mode, after the barrier: rows [rank * 128 / CL, +128 / CL) of the tile belong to this
// CTA.  2 CL threads per row (consecutive lanes), each 64-float stripe of the row is covered by
// one coalesced float4 access of the row's threads:  x = sum of the CL partials + b_2 (dropout2)
// + x_after_attention, then LayerNorm_next with shuffle reductions inside the row's lanes.
template <int CL>
__device__ __forceinline__ void fb_reduce_owned_rows(const FfnBlockArgs& a, int tile_i, int rank,
                                                     const float (*s_vec)[FB_D]) {
  constexpr int RO = LN_BM / CL, TPR = FB_EPI / RO, NQ = FB_D / 4 / TPR;
  const int t = threadIdx.x - 64;
  const int rl = rank * RO + t / TPR, cg = t % TPR;
  const int row = tile_i * LN_BM + rl;
  const bool row_ok = row < a.M;
  float4 acc[NQ];
#pragma unroll
  for (int i = 0; i < NQ; ++i) acc[i] = *reinterpret_cast<const float4*>(&s_vec[3][4 * (i * TPR + cg)]);
#pragma unroll
  for (int p = 0; p < CL; ++p) {
    const float* pp = a.part + ((long long)(tile_i * CL + p) * LN_BM + rl) * FB_D;
#pragma unroll
    for (int i = 0; i < NQ; ++i) {
      const float4 u = __ldcg(reinterpret_cast<const float4*>(pp + 4 * (i * TPR + cg)));
      acc[i].x += u.x; acc[i].y += u.y; acc[i].z += u.z; acc[i].w += u.w;
    }
  }
  if (a.drop2.rng != nullptr && row_ok) {
    const float scale = 1.0f / (1.0f - a.drop2.p);
#pragma unroll
    for (int i = 0; i < NQ; ++i) {
      const int col = 4 * (i * TPR + cg);
      const uint4 k = dropout_quad(a.drop2.rng, a.drop2.site, (unsigned)row, (unsigned)(col >> 2), a.drop2.p);
      acc[i].x = k.x ? acc[i].x * scale : 0.f; acc[i].y = k.y ? acc[i].y * scale : 0.f;
      acc[i].z = k.z ? acc[i].z * scale : 0.f; acc[i].w = k.w ? acc[i].w * scale : 0.f;
      if (a.drop2.mask != nullptr)
        *reinterpret_cast<uchar4*>(a.drop2.mask + (long long)row * FB_D + col) =
            make_uchar4((unsigned char)k.x, (unsigned char)k.y, (unsigned char)k.z, (unsigned char)k.w);
    }
  }
  const float* xs = a.x1s + ((long long)tile_i * LN_BM + rl) * FB_D;
  const bool norm = a.lnn_g != nullptr;
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < NQ; ++i) {
    const int col = 4 * (i * TPR + cg);
    const float4 x = __ldcg(reinterpret_cast<const float4*>(xs + col));
    acc[i].x += x.x; acc[i].y += x.y; acc[i].z += x.z; acc[i].w += x.w;
    sum += (acc[i].x + acc[i].y) + (acc[i].z + acc[i].w);
    if (row_ok) {
      const long long o = (long long)row * FB_D + col;
      if (a.x_out != nullptr) *reinterpret_cast<float4*>(a.x_out + o) = acc[i];
      if (!norm && a.out_f32 != nullptr) *reinterpret_cast<float4*>(a.out_f32 + o) = acc[i];
    }
  }
  if (!norm) return;  // uniform
#pragma unroll
  for (int o = TPR / 2; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  const float mean = sum * (1.0f / FB_D);
  float sq = 0.f;
#pragma unroll
  for (int i = 0; i < NQ; ++i) {
    const float dx = acc[i].x - mean, dy = acc[i].y - mean, dz = acc[i].z - mean, dw = acc[i].w - mean;
    sq = fmaf(dx, dx, sq); sq = fmaf(dy, dy, sq); sq = fmaf(dz, dz, sq); sq = fmaf(dw, dw, sq);
  }
#pragma unroll
  for (int o = TPR / 2; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
  const float rstd = rsqrtf(sq * (1.0f / FB_D) + a.eps);
  if (!row_ok) return;
#pragma unroll
  for (int i = 0; i < NQ; ++i) {
    const int col = 4 * (i * TPR + cg);
    const float4 g = *reinterpret_cast<const float4*>(&s_vec[4][col]);
    const float4 b = *reinterpret_cast<const float4*>(&s_vec[5][col]);
    const float4 y = make_float4((acc[i].x - mean) * rstd * g.x + b.x, (acc[i].y - mean) * rstd * g.y + b.y,
                                 (acc[i].z - mean) * rstd * g.z + b.z, (acc[i].w - mean) * rstd * g.w + b.w);
    const long long o = (long long)row * FB_D + col;
    if (a.xn_out != nullptr) *reinterpret_cast<uint2*>(a.xn_out + o) = pack_bf16x4(y);
    if (a.out_f32 != nullptr) *reinterpret_cast<float4*>(a.out_f32 + o) = y;
  }
}

__global__ void __launch_bounds__(LN_THREADS, 1)
encoder_ffn_block_kernel(const __grid_constant__ CUtensorMap map_att, const __grid_constant__ CUtensorMap map_wo,
                         const __grid_constant__ CUtensorMap map_w1, const __grid_constant__ CUtensorMap map_w2,
                         FfnBlockArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* A0 = smem;                      // att tile (4 k-blocks), later hid[2] (2 k-blocks each)
  uint8_t* A1 = smem + 4 * FB_KTILE;       // LayerNorm2(x) (4 k-blocks)
  uint8_t* ring = smem + 8 * FB_KTILE;
  __shared__ uint64_t full_bar[FB_NST], empty_bar[FB_NST];
  __shared__ uint64_t att_full, acc0_o_done, a1_ready, acc0_f2_done;
  __shared__ uint64_t f1_done[2], acc1_free[2], hid_ready[2], hid_free[2];
  __shared__ uint32_t tmem_base_s;
  __shared__ float ln_part[2][2][LN_BM];
  __shared__ __align__(16) float s_vec[6][FB_D];  // b_o, ln2 g, ln2 b, b2, lnn g, lnn b
  __shared__ __align__(16) float s_b1[FB_MAX_FF];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (a.dbg != nullptr && blockIdx.x == 0 && threadIdx.x == 32) {  // kernel entry, wall clock (ns) and cycles
    unsigned long long gt;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
    a.dbg[60] = (long long)gt;
    a.dbg[61] = clock64();
  }
  const int cl = a.cl;
  const int tile_i = blockIdx.x / cl, rank = blockIdx.x - tile_i * cl;  // cluster = cl consecutive CTAs
  const int m0 = tile_i * LN_BM;
  const int C = a.FF / FB_CH / cl;  // hidden chunks of this CTA
  const int c_lo = rank * C;        // ... starting at this chunk of the hidden layer

  if (threadIdx.x == 0) {
    for (int i = 0; i < FB_NST; ++i) { tc::mbar_init(&full_bar[i], 1); tc::mbar_init(&empty_bar[i], 1); }
    tc::mbar_init(&att_full, 1);
    tc::mbar_init(&acc0_o_done, 1);
    tc::mbar_init(&a1_ready, FB_EPI);
    tc::mbar_init(&acc0_f2_done, 1);
    for (int i = 0; i < 2; ++i) {
      tc::mbar_init(&f1_done[i], 1);
      tc::mbar_init(&acc1_free[i], FB_EPI);
      tc::mbar_init(&hid_ready[i], FB_EPI);
      tc::mbar_init(&hid_free[i], 1);
    }
    tc::fence_barrier_init();
  }
  if (warp == 1) tc::tmem_alloc<512>(&tmem_base_s);
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = tmem_base_s;
  const uint32_t acc0 = tmem, acc1[2] = {tmem + 256u, tmem + 384u};

  if (warp == 0) {
    if (lane == 0) {  // ===== TMA producer: tiles in exactly the order the MMA warp consumes them =====
      mbar_expect_tx(&att_full, 4 * FB_KTILE);
      for (int kb = 0; kb < 4; ++kb) tma_load_2d(A0 + kb * FB_KTILE, &map_att, kb * LN_BK, m0, &att_full);
      int it = 0;
      auto stage = [&](auto&& issue) {
        const int s = it % FB_NST;
        if (it >= FB_NST) tc::mbar_wait(&empty_bar[s], ((it / FB_NST) - 1) & 1);
        mbar_expect_tx(&full_bar[s], FB_STAGE);
        issue(ring + s * FB_STAGE, &full_bar[s]);
        ++it;
      };
      for (int kb = 0; kb < 4; ++kb)  // W_o [256 x 64] k-blocks
        stage([&](uint8_t* dst, uint64_t* bar) { tma_load_2d(dst, &map_wo, kb * LN_BK, 0, bar); });
      auto f1 = [&](int c) {  // W_1 rows [128 c, +128): two stages of two [128 x 64] k-blocks
        for (int st = 0; st < 2; ++st)
          stage([&](uint8_t* dst, uint64_t* bar) {
            tma_load_2d(dst, &map_w1, (2 * st) * LN_BK, (c_lo + c) * FB_CH, bar);
            tma_load_2d(dst + FB_STAGE / 2, &map_w1, (2 * st + 1) * LN_BK, (c_lo + c) * FB_CH, bar);
          });
      };
      auto f2 = [&](int c) {  // W_2 [256 x 64] k-blocks of hidden columns [128 c, +128)
        for (int st = 0; st < 2; ++st)
          stage([&](uint8_t* dst, uint64_t* bar) {
            tma_load_2d(dst, &map_w2, (c_lo + c) * FB_CH + st * LN_BK, 0, bar);
          });
      };
      f1(0);
      if (C > 1) f1(1);
      for (int c = 0; c < C; ++c) {
        f2(c);
        if (c + 2 < C) f1(c + 2);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {  // ===== MMA issuer =====
      constexpr uint32_t IDESC_256 = tc::make_idesc_bf16(LN_BM, 256);
      constexpr uint32_t IDESC_128 = tc::make_idesc_bf16(LN_BM, FB_CH);
      int it = 0;
      auto wait_stage = [&]() -> uint32_t {
        const int s = it % FB_NST;
        tc::mbar_wait(&full_bar[s], (it / FB_NST) & 1);
        tc::fence_after_sync();
        return tc::smem_u32(ring + s * FB_STAGE);
      };
      auto release_stage = [&]() { tc::mma_commit(&empty_bar[it % FB_NST]); ++it; };
      // ---- out_proj: acc0 = att W_o^T ----
      FB_STAMP(0);
      tc::mbar_wait(&att_full, 0);
      tc::fence_after_sync();
      FB_STAMP(1);
      for (int kb = 0; kb < 4; ++kb) {
        const uint32_t b_addr = wait_stage();
        const uint32_t a_addr = tc::smem_u32(A0 + kb * FB_KTILE);
#pragma unroll
        for (int k = 0; k < LN_BK; k += 16)
          tc::mma_bf16(acc0, tc::make_desc_sw128(a_addr + k * 2), tc::make_desc_sw128(b_addr + k * 2), IDESC_256,
                       (kb > 0 || k > 0) ? 1u : 0u);
        release_stage();
      }
      tc::mma_commit(&acc0_o_done);
      FB_STAMP(2);
      // ---- FFN ----
      tc::mbar_wait(&a1_ready, 0);  // LayerNorm2(x) is in A1; acc0 and A0 are free again
      tc::fence_after_sync();
      FB_STAMP(3);
      auto f1 = [&](int c) {  // acc1[c & 1] = A1 W_1[chunk c]^T
        for (int st = 0; st < 2; ++st) {
          const uint32_t b_addr = wait_stage();
#pragma unroll
          for (int kk = 0; kk < 2; ++kk) {
            const uint32_t a_addr = tc::smem_u32(A1 + (2 * st + kk) * FB_KTILE);
#pragma unroll
            for (int k = 0; k < LN_BK; k += 16)
              tc::mma_bf16(acc1[c & 1], tc::make_desc_sw128(a_addr + k * 2),
                           tc::make_desc_sw128(b_addr + kk * (FB_STAGE / 2) + k * 2), IDESC_128,
                           (st > 0 || kk > 0 || k > 0) ? 1u : 0u);
          }
          release_stage();
        }
        tc::mma_commit(&f1_done[c & 1]);
      };
      f1(0);
      if (C > 1) f1(1);
      for (int c = 0; c < C; ++c) {
        const int b = c & 1;
        tc::mbar_wait(&hid_ready[b], (c >> 1) & 1);  // relu(hidden chunk c) is in hid[b]
        tc::fence_after_sync();
        FB_STAMP(8 + c);
        for (int st = 0; st < 2; ++st) {           // acc0 += hid[b] W_2[:, chunk c]^T
          const uint32_t b_addr = wait_stage();
          const uint32_t a_addr = tc::smem_u32(A0 + (2 * b + st) * FB_KTILE);
#pragma unroll
          for (int k = 0; k < LN_BK; k += 16)
            tc::mma_bf16(acc0, tc::make_desc_sw128(a_addr + k * 2), tc::make_desc_sw128(b_addr + k * 2), IDESC_256,
                         (c > 0 || st > 0 || k > 0) ? 1u : 0u);
          release_stage();
        }
        tc::mma_commit(&hid_free[b]);
        if (c + 2 < C) {
          tc::mbar_wait(&acc1_free[b], (c >> 1) & 1);  // the epilogue has read acc1[b]
          tc::fence_after_sync();
          f1(c + 2);
        }
      }
      tc::mma_commit(&acc0_f2_done);
      FB_STAMP(4);
    }
  } else {  // ===== epilogue warps 2..9: TMEM lane quadrant = warp % 4, column half = (warp - 2) / 4 =====
    const int q = warp & 3, half = (warp - 2) >> 2;
    const int row_local = q * 32 + lane, row = m0 + row_local;
    const bool row_ok = row < a.M;
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
    for (int c = threadIdx.x - 64; c < FB_D; c += FB_EPI) {
      s_vec[0][c] = __ldg(a.b_o + c);
      s_vec[1][c] = __ldg(a.ln2_g + c);
      s_vec[2][c] = __ldg(a.ln2_b + c);
      s_vec[3][c] = __ldg(a.b2 + c);
      s_vec[4][c] = a.lnn_g != nullptr ? __ldg(a.lnn_g + c) : 1.f;
      s_vec[5][c] = a.lnn_g != nullptr ? __ldg(a.lnn_b + c) : 0.f;
    }
    for (int c = threadIdx.x - 64; c < a.FF; c += FB_EPI) s_b1[c] = __ldg(a.b1 + c);
    tc::group_sync(1, FB_EPI);
    const int c_begin = half * (FB_D / 2), c_end = c_begin + FB_D / 2;

    // ---- E1: x = acc0 + b_o (dropout1) + x_in ; LayerNorm2 -> A1 ----
    // Rows meet global memory through a per-warp [32 x 32] fp32 staging tile (in the A0 region,
    // idle here): a TMEM lane is a row, a coalesced access wants 4 rows x 128 contiguous bytes.
    float* tile = reinterpret_cast<float*>(A0) + (warp - 2) * EP_TILE_FLOATS;
    const int qrow0 = m0 + q * 32 + (lane >> 3), qcol = 4 * (lane & 7);
    float4 res[8];
    auto fetch_rows = [&](const float* src, int j0) {  // rows qrow0 + 4 i, columns j0 + qcol .. +3
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int r = qrow0 + 4 * i;
        res[i] = r < a.M ? *reinterpret_cast<const float4*>(src + (long long)r * FB_D + j0 + qcol)
                         : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    };
    const int own_n = LN_BM / cl, own0 = rank * own_n;  // tile rows this CTA finishes in cluster mode
    fetch_rows(a.x_in, c_begin);  // issued before the wait: its latency hides behind out_proj
    tc::mbar_wait(&acc0_o_done, 0);
    tc::fence_after_sync();
    if (threadIdx.x == 64) FB_STAMP(32);
    float sum = 0.f, sumsq = 0.f;
#pragma unroll 1
    for (int j0 = c_begin; j0 < c_end; j0 += 32) {
      float v[32];
      tc::tmem_ld32(acc0 + lane_off + (uint32_t)j0, v);
      tc::tmem_ld_wait();
      add_bias_act(v, s_vec[0], j0, ACT_NONE);
      if (a.drop1.rng != nullptr && row_ok) dropout32(v, a.drop1, row, j0, FB_D);
      tile_put_row(tile, lane, v);
      __syncwarp();
      float4 xq[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float4 t = *tile_quad(tile, lane, i);
        xq[i] = make_float4(t.x + res[i].x, t.y + res[i].y, t.z + res[i].z, t.w + res[i].w);
      }
      if (j0 + 32 < c_end) fetch_rows(a.x_in, j0 + 32);  // next slab: before this one's stores (x may alias)
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int r = qrow0 + 4 * i;
        if (cl == 1) {
          if (r < a.M) *reinterpret_cast<float4*>(a.x_out + (long long)r * FB_D + j0 + qcol) = xq[i];
        } else {
          // every CTA of the cluster computes the whole tile (x_out may alias x_in, which the
          // siblings are still reading): keep only the rows this CTA finishes, in scratch
          const int rl = r - m0;
          if (rl >= own0 && rl < own0 + own_n)
            *reinterpret_cast<float4*>(a.x1s + ((long long)tile_i * LN_BM + rl) * FB_D + j0 + qcol) = xq[i];
        }
        *tile_quad(tile, lane, i) = xq[i];
      }
      __syncwarp();
      tile_get_row(tile, lane, v);
      __syncwarp();
#pragma unroll
      for (int j = 0; j < 32; ++j) { sum += v[j]; sumsq = fmaf(v[j], v[j], sumsq); }
      tc::tmem_st32(acc0 + lane_off + (uint32_t)j0, v);
    }
    tc::tmem_st_wait();
    if (threadIdx.x == 64) FB_STAMP(37);
    fb_layernorm_rows(acc0 + lane_off, c_begin, c_end, half, row_local, sum, sumsq, ln_part, s_vec[1], s_vec[2], a.eps,
                      [&](int j0, const float* v) { fb_store_operand_row(A1, row_local, j0, v); });
    tc::fence_async_smem();        // A1 written through the generic proxy -> visible to the tensor core
    tc::fence_before_sync();       // ... and the TMEM reads of acc0 are complete
    mbar_arrive(&a1_ready);
    if (threadIdx.x == 64) FB_STAMP(33);

    // ---- E2: hidden chunks: relu(acc1 + b_1) (dropout) -> hid[b] ----
    for (int c = 0; c < C; ++c) {
      const int b = c & 1;
      tc::mbar_wait(&f1_done[b], (c >> 1) & 1);
      tc::fence_after_sync();
      if (threadIdx.x == 64) FB_STAMP(40 + 2 * c);
      float v[2][32];
      const int h0 = half * (FB_CH / 2);  // this warp's 64 of the chunk's 128 columns
      tc::tmem_ld32(acc1[b] + lane_off + (uint32_t)h0, v[0]);
      tc::tmem_ld32(acc1[b] + lane_off + (uint32_t)(h0 + 32), v[1]);
      tc::tmem_ld_wait();
      tc::fence_before_sync();
      mbar_arrive(&acc1_free[b]);  // acc1[b] may be overwritten by chunk c + 2
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        const int col = (c_lo + c) * FB_CH + h0 + 32 * t;  // hidden column of v[t][0]
#pragma unroll
        for (int j = 0; j < 32; ++j) v[t][j] = fmaxf(v[t][j] + s_b1[col + j], 0.f);
        if (a.drop_h.rng != nullptr && row_ok) dropout32(v[t], a.drop_h, row, col, a.FF);
      }
      if (c >= 2) {  // chunk c - 2's contraction with W_2 must have consumed hid[b]
        tc::mbar_wait(&hid_free[b], ((c >> 1) - 1) & 1);
      }
      fb_store_operand_row(A0 + 2 * b * FB_KTILE, row_local, h0, v[0]);
      fb_store_operand_row(A0 + 2 * b * FB_KTILE, row_local, h0 + 32, v[1]);
      tc::fence_async_smem();
      mbar_arrive(&hid_ready[b]);
      if (threadIdx.x == 64) FB_STAMP(41 + 2 * c);
    }

    // ---- E3: x = acc0 + b_2 (dropout2) + x ; LayerNorm_next ----
    tc::mbar_wait(&acc0_f2_done, 0);
    tc::fence_after_sync();
    if (threadIdx.x == 64) FB_STAMP(34);
    if (cl > 1) {
      // cluster mode: this CTA's partial result -> scratch; finished after the cluster barrier
      float* pbase = a.part + (long long)(tile_i * cl + rank) * LN_BM * FB_D;
#pragma unroll 1
      for (int j0 = c_begin; j0 < c_end; j0 += 32) {
        float v[32];
        tc::tmem_ld32(acc0 + lane_off + (uint32_t)j0, v);
        tc::tmem_ld_wait();
        tile_put_row(tile, lane, v);
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int rl = q * 32 + (lane >> 3) + 4 * i;
          *reinterpret_cast<float4*>(pbase + (long long)rl * FB_D + j0 + qcol) = *tile_quad(tile, lane, i);
        }
        __syncwarp();
      }
    } else {
    sum = 0.f; sumsq = 0.f;
    const bool norm = a.lnn_g != nullptr;
    fetch_rows(a.x_out, c_begin);  // x after attention (written in E1, re-read coalesced)
#pragma unroll 1
    for (int j0 = c_begin; j0 < c_end; j0 += 32) {
      float v[32];
      tc::tmem_ld32(acc0 + lane_off + (uint32_t)j0, v);
      tc::tmem_ld_wait();
      add_bias_act(v, s_vec[3], j0, ACT_NONE);
      if (a.drop2.rng != nullptr && row_ok) dropout32(v, a.drop2, row, j0, FB_D);
      tile_put_row(tile, lane, v);
      __syncwarp();
      float4 xq[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float4 t = *tile_quad(tile, lane, i);
        xq[i] = make_float4(t.x + res[i].x, t.y + res[i].y, t.z + res[i].z, t.w + res[i].w);
      }
      if (j0 + 32 < c_end) fetch_rows(a.x_out, j0 + 32);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int r = qrow0 + 4 * i;
        if (r < a.M) {
          const long long o = (long long)r * FB_D + j0 + qcol;
          *reinterpret_cast<float4*>(a.x_out + o) = xq[i];
          if (!norm && a.out_f32 != nullptr) *reinterpret_cast<float4*>(a.out_f32 + o) = xq[i];
        }
        *tile_quad(tile, lane, i) = xq[i];
      }
      __syncwarp();
      if (norm) {
        tile_get_row(tile, lane, v);
#pragma unroll
        for (int j = 0; j < 32; ++j) { sum += v[j]; sumsq = fmaf(v[j], v[j], sumsq); }
        tc::tmem_st32(acc0 + lane_off + (uint32_t)j0, v);
      }
      __syncwarp();
    }
    if (norm) {
      tc::tmem_st_wait();
      fb_layernorm_rows(acc0 + lane_off, c_begin, c_end, half, row_local, sum, sumsq, ln_part, s_vec[4], s_vec[5], a.eps,
                        [&](int j0, const float* v) {
                          tile_put_row(tile, lane, v);
                          __syncwarp();
#pragma unroll
                          for (int i = 0; i < 8; ++i) {
                            const int r = qrow0 + 4 * i;
                            if (r < a.M) {
                              const float4 y = *tile_quad(tile, lane, i);
                              const long long o = (long long)r * FB_D + j0 + qcol;
                              if (a.xn_out != nullptr) *reinterpret_cast<uint2*>(a.xn_out + o) = pack_bf16x4(y);
                              if (a.out_f32 != nullptr) *reinterpret_cast<float4*>(a.out_f32 + o) = y;
                            }
                          }
                          __syncwarp();
                        });
    }
    }
    if (threadIdx.x == 64) FB_STAMP(35);
    tc::fence_before_sync();
  }
  __syncthreads();
  if (warp == 1) tc::tmem_dealloc<512>(tmem);
  if (cl > 1) {
    cluster_sync_all();  // the partial results of all CTAs of the cluster are in global memory
    if (warp >= 2) {
      if (cl == 8) fb_reduce_owned_rows<8>(a, tile_i, rank, s_vec);
      else if (cl == 4) fb_reduce_owned_rows<4>(a, tile_i, rank, s_vec);
      else fb_reduce_owned_rows<2>(a, tile_i, rank, s_vec);
      if (threadIdx.x == 64) {
        FB_STAMP(36);
        if (a.dbg != nullptr && blockIdx.x == 0) {
          unsigned long long gt;
          asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
          a.dbg[62] = (long long)gt;
        }
      }
    }
  }
}

// host side: tensor maps with a caller-chosen box height
static int make_map_box(CUtensorMap* map, const void* base, int rows, int cols, int box_rows) {
  return make_map(map, base, rows, cols, box_rows);
}

int launch_ffn_block(const __nv_bfloat16* att, const __nv_bfloat16* wo, const __nv_bfloat16* w1,
                     const __nv_bfloat16* w2, FfnBlockArgs a, cudaStream_t stream) {
  MPA_CHECK_ARG(a.FF % FB_CH == 0 && a.FF >= FB_CH && a.FF <= FB_MAX_FF,
                "ffn_block: FF must be a multiple of %d, at most %d", FB_CH, FB_MAX_FF);
  CUtensorMap m_att, m_wo, m_w1, m_w2;
  int rc = make_map_box(&m_att, att, a.M, FB_D, LN_BM);
  if (rc == MPA_OK) rc = make_map_box(&m_wo, wo, FB_D, FB_D, 256);
  if (rc == MPA_OK) rc = make_map_box(&m_w1, w1, a.FF, FB_D, FB_CH);
  if (rc == MPA_OK) rc = make_map_box(&m_w2, w2, FB_D, a.FF, 256);
  if (rc != MPA_OK) return rc;
  static DeviceOnce attr;
  if (attr.pending()) {
    MPA_CUDA(cudaFuncSetAttribute(encoder_ffn_block_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FB_SMEM));
    attr.done();
  }
  if (const char* e = getenv("MPA_FFN_DEBUG")) a.dbg = (long long*)strtoull(e, nullptr, 0);  // device pointer
  const int tiles = (a.M + LN_BM - 1) / LN_BM;
  if (a.part == nullptr || a.x1s == nullptr) a.cl = 1;
  {
    ProfScope ps("encoder_ffn_block", stream);
    if (a.cl == 1) {
      encoder_ffn_block_kernel<<<tiles, LN_THREADS, FB_SMEM, stream>>>(m_att, m_wo, m_w1, m_w2, a);
    } else {
      cudaLaunchConfig_t cfg{};
      cfg.gridDim = dim3(tiles * a.cl);
      cfg.blockDim = dim3(LN_THREADS);
      cfg.dynamicSmemBytes = FB_SMEM;
      cfg.stream = stream;
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeClusterDimension;
      at[0].val.clusterDim.x = a.cl; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
      cfg.attrs = at;
      cfg.numAttrs = 1;
      MPA_CUDA(cudaLaunchKernelEx(&cfg, encoder_ffn_block_kernel, m_att, m_wo, m_w1, m_w2, a));
    }
  }
  MPA_LAUNCH_CHECK();
  return MPA_OK;
}

// cluster width of the FFN block: split the hidden layer over as many CTAs as keep tiles x cl
// within one wave of SMs (8 = portable cluster limit); MPA_FFN_CLUSTER overrides (A/B runs)
static int ffn_cluster_size(int M, int FF) {
  static const int forced = getenv("MPA_FFN_CLUSTER") ? atoi(getenv("MPA_FFN_CLUSTER")) : 0;
  const int tiles = (M + LN_BM - 1) / LN_BM, C = FF / FB_CH;
  if (forced == 1 || forced == 2 || forced == 4 || forced == 8) return (C % forced == 0) ? forced : 1;
  for (int cl = 8; cl >= 2; cl >>= 1)
    if (C % cl == 0 && tiles * cl <= device_sms()) return cl;
  return 1;
}

// ======================================================================================
// Fused first half of a pre-LN encoder layer for d_model = 256, head dim 32 (bf16 operands):
//     qkv <- LayerNorm1(x) W_in^T + b_in ;  att <- softmax(q k^T / sqrt(32) + key mask) v
// One CTA owns the tokens of SPT = 128 / P whole shapes (attention never leaves a shape) and
// AT_HPC of the 8 heads.  Per head the three 32-row slices of W_in (q, k, v rows of that head)
// are stacked by TMA into one [96 x 256] B operand, so ONE tcgen05 GEMM (M=128, N=96) produces
// the head's q | k | v in TMEM; the MMA of head h+1 overlaps the attention of head h (two TMEM
// slots, two weight stages).  The 8 epilogue warps then run the attention on CUDA cores (S=20
// tokens: a 20x20x32 problem per (shape, head)): k and v rows go to shared
// memory, each thread of a row pair scores half of the keys, the full score row is exchanged
// through shared memory, softmax in registers, and each thread accumulates half of the head's
// 32 output channels.  Replaces the QKV GEMM launch + the attention launch of every layer
// (nn.MultiheadAttention inside nn.TransformerEncoderLayer, transformer.py:23-34).
constexpr int AT_HD = 32;                     // head dim
constexpr int AT_NB = 3 * AT_HD;              // 96 accumulator columns per head: q | k | v
constexpr int AT_WSTAGE = 4 * AT_NB * 128;    // one head's stacked weight slices: 4 k-blocks of [96 x 64] bf16 = 48 KB
constexpr int AT_SMEM = 4 * FB_KTILE + 2 * AT_WSTAGE + 2 * LN_BM * AT_HD * 4 + 1024;

struct AttnArgs {
  const float* b_in;            // [768]
  const unsigned char* valid;   // [B*P] key-padding mask (1 = valid) or nullptr
  __nv_bfloat16* att;           // [T, 256] out
  int B, P, SPT, heads_per_cta;
  DropoutSpec drop;             // on the attention probabilities; mask [B, 8, P, P]
};

template <int PT>  // compile-time bound of tokens per shape (P <= PT): score rows live in registers
__global__ void __launch_bounds__(LN_THREADS, 1)
encoder_attn_kernel(const __grid_constant__ CUtensorMap map_xn, const __grid_constant__ CUtensorMap map_win,
                    AttnArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* A = smem;                        // LayerNorm1(x) tile, 4 k-blocks
  uint8_t* wring = smem + 4 * FB_KTILE;     // 2 weight stages
  __shared__ uint64_t a_full, w_full[2], w_empty[2], acc_full[2], acc_free[2];
  __shared__ uint32_t tmem_base_s;
  float (*ks)[AT_HD] = reinterpret_cast<float (*)[AT_HD]>(wring + 2 * AT_WSTAGE);  // k rows of the head, 16 KB
  float (*vs)[AT_HD] = ks + LN_BM;                                                  // v rows, 16 KB
  __shared__ float sc[LN_BM][PT + 1];                // score exchange
  __shared__ float s_bias[3 * 256];
  __shared__ unsigned char s_valid[LN_BM];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int P = a.P, H = 8;
  const int shape0 = blockIdx.x * a.SPT;
  const int nshapes = min(a.SPT, a.B - shape0);
  const int row0 = shape0 * P;              // first global token row of this tile
  const int nrows = nshapes * P;
  const int h_begin = blockIdx.y * a.heads_per_cta;
  const int nh = a.heads_per_cta;

  if (threadIdx.x == 0) {
    tc::mbar_init(&a_full, 1);
    for (int i = 0; i < 2; ++i) {
      tc::mbar_init(&w_full[i], 1);
      tc::mbar_init(&w_empty[i], 1);
      tc::mbar_init(&acc_full[i], 1);
      tc::mbar_init(&acc_free[i], FB_EPI);
    }
    tc::fence_barrier_init();
  }
  if (warp == 1) tc::tmem_alloc<256>(&tmem_base_s);
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = tmem_base_s;

  if (warp == 0) {
    if (lane == 0) {  // ===== TMA producer =====
      mbar_expect_tx(&a_full, 4 * FB_KTILE);
      for (int kb = 0; kb < 4; ++kb) tma_load_2d(A + kb * FB_KTILE, &map_xn, kb * LN_BK, row0, &a_full);
      for (int hh = 0; hh < nh; ++hh) {
        const int s = hh & 1, h = h_begin + hh;
        if (hh >= 2) tc::mbar_wait(&w_empty[s], ((hh >> 1) - 1) & 1);
        uint8_t* dst = wring + s * AT_WSTAGE;
        mbar_expect_tx(&w_full[s], AT_WSTAGE);
        for (int kb = 0; kb < 4; ++kb)
          for (int part = 0; part < 3; ++part)  // q, k, v rows of head h -> rows 32 part .. of the stacked tile
            tma_load_2d(dst + kb * (AT_NB * 128) + part * (AT_HD * 128), &map_win, kb * LN_BK,
                        part * 256 + h * AT_HD, &w_full[s]);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {  // ===== MMA issuer =====
      constexpr uint32_t IDESC = tc::make_idesc_bf16(LN_BM, AT_NB);
      tc::mbar_wait(&a_full, 0);
      tc::fence_after_sync();
      for (int hh = 0; hh < nh; ++hh) {
        const int s = hh & 1;
        tc::mbar_wait(&w_full[s], (hh >> 1) & 1);
        if (hh >= 2) tc::mbar_wait(&acc_free[s], ((hh >> 1) - 1) & 1);
        tc::fence_after_sync();
        const uint32_t acc = tmem + (uint32_t)(s * 128);
        const uint32_t b_base = tc::smem_u32(wring + s * AT_WSTAGE);
        for (int kb = 0; kb < 4; ++kb) {
          const uint32_t a_addr = tc::smem_u32(A + kb * FB_KTILE), b_addr = b_base + kb * (AT_NB * 128);
#pragma unroll
          for (int k = 0; k < LN_BK; k += 16)
            tc::mma_bf16(acc, tc::make_desc_sw128(a_addr + k * 2), tc::make_desc_sw128(b_addr + k * 2), IDESC,
                         (kb > 0 || k > 0) ? 1u : 0u);
        }
        tc::mma_commit(&w_empty[s]);
        tc::mma_commit(&acc_full[s]);
      }
    }
  } else {  // ===== epilogue + attention: warps 2..9 =====
    const int q4 = warp & 3, half = (warp - 2) >> 2;
    const int r = q4 * 32 + lane;            // row in the tile = TMEM lane
    const uint32_t lane_off = (uint32_t)(q4 * 32) << 16;
    for (int c = threadIdx.x - 64; c < 3 * 256; c += FB_EPI) s_bias[c] = __ldg(a.b_in + c);
    for (int c = threadIdx.x - 64; c < LN_BM; c += FB_EPI)
      s_valid[c] = (c < nrows && (a.valid == nullptr || a.valid[row0 + c] != 0)) ? 1 : 0;
    tc::group_sync(1, FB_EPI);
    const bool live = r < nrows;
    const int sh = live ? r / P : 0, tok = r - sh * P;  // local shape, token
    const int kbase = sh * P;                             // first key row of this row's shape
    const float scale = rsqrtf((float)AT_HD);
    const int j_lo = half * (PT / 2), j_hi = half ? PT : PT / 2;  // this thread's share of the keys

    for (int hh = 0; hh < nh; ++hh) {
      const int s = hh & 1, h = h_begin + hh;
      tc::mbar_wait(&acc_full[s], (hh >> 1) & 1);
      tc::fence_after_sync();
      const uint32_t acc = tmem + (uint32_t)(s * 128) + lane_off;
      float q[AT_HD], kv[AT_HD];
      tc::tmem_ld32(acc, q);
      tc::tmem_ld32(acc + (uint32_t)(half ? 64 : 32), kv);  // half 0: k row, half 1: v row
      tc::tmem_ld_wait();
      tc::fence_before_sync();
      mbar_arrive(&acc_free[s]);
      {
        const float* bq = s_bias + h * AT_HD;
        const float* bkv = s_bias + (half ? 512 : 256) + h * AT_HD;
        float* dst = half ? vs[r] : ks[r];
#pragma unroll
        for (int c = 0; c < AT_HD; ++c) { q[c] += bq[c]; kv[c] += bkv[c]; }
#pragma unroll
        for (int c = 0; c < AT_HD; c += 4)
          *reinterpret_cast<float4*>(dst + c) = make_float4(kv[c], kv[c + 1], kv[c + 2], kv[c + 3]);
      }
      tc::group_sync(1, FB_EPI);  // k, v of every row are in shared memory
      // scores of this thread's keys
#pragma unroll
      for (int j = j_lo; j < j_hi; ++j) {
        float d = -3.0e38f;
        if (j < P && live && s_valid[kbase + j]) {
          const float4* kr = reinterpret_cast<const float4*>(ks[kbase + j]);
          float d0 = 0.f, d1 = 0.f;
#pragma unroll
          for (int c4 = 0; c4 < AT_HD / 4; ++c4) {
            const float4 kk = kr[c4];
            d0 = fmaf(q[4 * c4], kk.x, d0); d1 = fmaf(q[4 * c4 + 1], kk.y, d1);
            d0 = fmaf(q[4 * c4 + 2], kk.z, d0); d1 = fmaf(q[4 * c4 + 3], kk.w, d1);
          }
          d = (d0 + d1) * scale;
        }
        sc[r][j] = d;
      }
      tc::group_sync(1, FB_EPI);  // the full score row is in shared memory
      float p[PT];
      float mx = -3.0e38f;
#pragma unroll
      for (int j = 0; j < PT; ++j) { p[j] = sc[r][j]; mx = fmaxf(mx, p[j]); }
      float den = 0.f;
#pragma unroll
      for (int j = 0; j < PT; ++j) { p[j] = p[j] > -1.0e38f ? __expf(p[j] - mx) : 0.f; den += p[j]; }
      const float inv = den > 0.f ? 1.0f / den : 0.f;  // a shape without a valid key cannot occur; guard anyway
      if (a.drop.rng != nullptr && live) {  // dropout on the attention probabilities
        const unsigned rowid = (unsigned)(((shape0 + sh) * H + h) * P + tok);
        const float keep_scale = 1.0f / (1.0f - a.drop.p);
#pragma unroll
        for (int jq = 0; jq < PT; jq += 4) {
          const uint4 k4 = dropout_quad(a.drop.rng, a.drop.site, rowid, (unsigned)(jq >> 2), a.drop.p);
          const unsigned kk[4] = {k4.x, k4.y, k4.z, k4.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            if (jq + e < PT) {
              p[jq + e] = kk[e] ? p[jq + e] * keep_scale : 0.f;
              if (a.drop.mask != nullptr && half == 0 && jq + e < P)
                a.drop.mask[(long long)rowid * P + jq + e] = (unsigned char)kk[e];
            }
          }
        }
      }
      // this thread's 16 output channels over all keys
      float o[AT_HD / 2];
#pragma unroll
      for (int c = 0; c < AT_HD / 2; ++c) o[c] = 0.f;
      const int c0 = half * (AT_HD / 2);
#pragma unroll
      for (int j = 0; j < PT; ++j) {
        if (j < P) {
          const float pj = p[j] * inv;
          const float4* vr = reinterpret_cast<const float4*>(vs[kbase + j] + c0);
#pragma unroll
          for (int c4 = 0; c4 < AT_HD / 8; ++c4) {
            const float4 vv = vr[c4];
            o[4 * c4] = fmaf(pj, vv.x, o[4 * c4]); o[4 * c4 + 1] = fmaf(pj, vv.y, o[4 * c4 + 1]);
            o[4 * c4 + 2] = fmaf(pj, vv.z, o[4 * c4 + 2]); o[4 * c4 + 3] = fmaf(pj, vv.w, o[4 * c4 + 3]);
          }
        }
      }
      if (live) {
        __nv_bfloat16* op = a.att + (long long)(row0 + r) * 256 + h * AT_HD + c0;
#pragma unroll
        for (int c = 0; c < AT_HD / 2; c += 4)
          *reinterpret_cast<uint2*>(op + c) = pack_bf16x4(make_float4(o[c], o[c + 1], o[c + 2], o[c + 3]));
      }
      tc::group_sync(1, FB_EPI);  // ks / vs / sc are reused by the next head
    }
    tc::fence_before_sync();
  }
  __syncthreads();
  if (warp == 1) tc::tmem_dealloc<256>(tmem);
}

// tensor map with an arbitrary box height (the weight slices are 32 rows)
static int make_map_rows(CUtensorMap* map, const void* base, int rows, int cols, int box_rows) {
  return make_map(map, base, rows, cols, box_rows);
}

int launch_encoder_attn(const __nv_bfloat16* xn, const __nv_bfloat16* w_in, AttnArgs a, int T, cudaStream_t stream) {
  MPA_CHECK_ARG(a.P >= 1 && a.P <= 32, "encoder_attn: 1 <= P <= 32");
  a.SPT = LN_BM / a.P;
  const int tiles = (a.B + a.SPT - 1) / a.SPT;
  // few token tiles: split the 8 heads over more CTAs (each re-reads the 64 KB activation tile)
  int hs = 1;
  while (hs < 8 && tiles * hs * 2 <= device_sms()) hs *= 2;
  a.heads_per_cta = 8 / hs;
  CUtensorMap m_xn, m_w;
  int rc = make_map_rows(&m_xn, xn, T, 256, LN_BM);
  if (rc == MPA_OK) rc = make_map_rows(&m_w, w_in, 768, 256, AT_HD);
  if (rc != MPA_OK) return rc;
  static DeviceOnce attr;
  if (attr.pending()) {
    MPA_CUDA(cudaFuncSetAttribute(encoder_attn_kernel<20>, cudaFuncAttributeMaxDynamicSharedMemorySize, AT_SMEM));
    MPA_CUDA(cudaFuncSetAttribute(encoder_attn_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, AT_SMEM));
    attr.done();
  }
  {
    ProfScope ps("encoder_attn", stream);
    if (a.P <= 20)
      encoder_attn_kernel<20><<<dim3(tiles, hs), LN_THREADS, AT_SMEM, stream>>>(m_xn, m_w, a);
    else
      encoder_attn_kernel<32><<<dim3(tiles, hs), LN_THREADS, AT_SMEM, stream>>>(m_xn, m_w, a);
  }
  MPA_LAUNCH_CHECK();
  return MPA_OK;
}

// ---- small token-level kernels ---------------------------------------------
// LayerNorm over the last dim (D <= 1024, multiple of 32): one warp per row,
// fp32 statistics (two-pass in registers), output bf16 (GEMM operand) and/or fp32.
__global__ void layernorm_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                 const float* __restrict__ beta, int rows, int D, float eps,
                                 __nv_bfloat16* __restrict__ out_bf16, float* __restrict__ out_f32,
                                 long long plane) {
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float* xr = x + (long long)row * D;
  float v[32];
  const int per = D / 32;
  float s = 0.f;
  for (int i = 0; i < per; ++i) { v[i] = xr[lane + 32 * i]; s += v[i]; }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const float mean = s / (float)D;
  float q = 0.f;
  for (int i = 0; i < per; ++i) { const float d = v[i] - mean; q = fmaf(d, d, q); }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
  const float rstd = rsqrtf(q / (float)D + eps);
  for (int i = 0; i < per; ++i) {
    const int c = lane + 32 * i;
    const float r = (v[i] - mean) * rstd * gamma[c] + beta[c];
    if (out_bf16) store_operand1(out_bf16, (long long)row * D + c, plane, r);
    if (out_f32) out_f32[(long long)row * D + c] = r;
  }
}

// Multi-head self-attention over the P part tokens of one shape with a
// key-padding mask (nn.MultiheadAttention semantics: softmax over valid keys
// only; scale 1/sqrt(hd)).  qkv [B*P, 3*D] fp32 (q | k | v), out [B*P, D] bf16.
// One warp per (shape, head): K (row-padded) and V staged in shared memory;
// per query the 32 lanes first hold one key score each (softmax by shuffles),
// then one (or two) output channels each.  P <= 32, hd <= 64.
constexpr int ATT_MAX_WARPS = 32;  // one warp per query row (P <= 32): the rows run concurrently
__global__ void __launch_bounds__(ATT_MAX_WARPS * 32, 1)
attention_kernel(const float* __restrict__ qkv, const unsigned char* __restrict__ valid, int B, int P,
                 int H, int hd, __nv_bfloat16* __restrict__ out, long long plane, DropoutSpec drop) {
  extern __shared__ float sm[];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int gw = blockIdx.x;  // (shape, head)
  const int ldk = hd + 1;
  float* qs = sm;
  float* ks = qs + P * hd;
  float* vs = ks + P * ldk;
  const int b = gw / H, h = gw % H, D = H * hd;
  for (int i = threadIdx.x; i < P * hd; i += blockDim.x) {
    const int p = i / hd, c = i % hd;
    const float* base = qkv + (long long)(b * P + p) * 3 * D + h * hd + c;
    qs[i] = base[0];
    ks[p * ldk + c] = base[D];
    vs[i] = base[2 * D];
  }
  __syncthreads();
  const bool key_ok = lane < P && (valid == nullptr || valid[b * P + lane] != 0);
  const float scale = rsqrtf((float)hd);
  for (int i = w; i < P; i += (int)(blockDim.x >> 5)) {
    float s = -3.0e38f;
    if (key_ok) {
      float d0 = 0.f, d1 = 0.f;
      for (int c = 0; c + 1 < hd; c += 2) {
        d0 = fmaf(qs[i * hd + c], ks[lane * ldk + c], d0);
        d1 = fmaf(qs[i * hd + c + 1], ks[lane * ldk + c + 1], d1);
      }
      if (hd & 1) d0 = fmaf(qs[i * hd + hd - 1], ks[lane * ldk + hd - 1], d0);
      s = (d0 + d1) * scale;
    }
    float mx = s;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    float e = key_ok ? __expf(s - mx) : 0.f;
    float den = e;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) den += __shfl_xor_sync(0xffffffffu, den, o);
    // a shape without any valid key cannot occur (>= 1 valid part); guard anyway
    e = den > 0.f ? e / den : 0.f;
    if (drop.rng != nullptr) {  // dropout on the attention probabilities (nn.MultiheadAttention)
      const unsigned el = (unsigned)((gw * P + i) * P + lane);  // [B, H, P, P] element
      const bool keep = lane < P ? dropout_quad(drop.rng, drop.site, 0u, el, drop.p).x != 0u : false;
      e = keep ? e / (1.0f - drop.p) : 0.f;
      if (drop.mask != nullptr && lane < P) drop.mask[el] = keep ? 1 : 0;
    }
    float o0 = 0.f, o1 = 0.f;
    for (int j = 0; j < P; ++j) {
      const float pj = __shfl_sync(0xffffffffu, e, j);
      if (lane < hd) o0 = fmaf(pj, vs[j * hd + lane], o0);
      if (lane + 32 < hd) o1 = fmaf(pj, vs[j * hd + lane + 32], o1);
    }
    const long long ob = (long long)(b * P + i) * D + h * hd;
    if (lane < hd) store_operand1(out, ob + lane, plane, o0);
    if (lane + 32 < hd) store_operand1(out, ob + lane + 32, plane, o1);
  }
}

// several fp32 -> bf16 conversions in one launch (blockIdx.y selects the segment)
struct CvtBatch {
  const float* src[64];
  __nv_bfloat16* dst[64];
  long long n[64];
  long long plane[64];  // 0: one bf16 plane; else hi / mid / lo planes this many elements apart
};
__global__ void f32_to_bf16_batch_kernel(CvtBatch cb) {
  const float* __restrict__ in = cb.src[blockIdx.y];
  __nv_bfloat16* __restrict__ o = cb.dst[blockIdx.y];
  const long long n = cb.n[blockIdx.y], plane = cb.plane[blockIdx.y];
  for (long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4; i < n;
       i += (long long)gridDim.x * blockDim.x * 4) {
    if (i + 3 < n && (plane & 3) == 0) {
      store_operand4(o, i, plane, *reinterpret_cast<const float4*>(in + i));
    } else {
      for (long long k = i; k < n && k < i + 4; ++k) store_operand1(o, k, plane, in[k]);
    }
  }
}

}  // namespace mpa

using namespace mpa;

extern "C" {

/* Y = act(X W^T + b) (+ residual); X [M,K] fp32, W [N,K] fp32 (converted to bf16
 * operands), out fp32.  Generic entry (pose head, tests). */
size_t mpa_linear_workspace_bytes_ex(int M, int N, int K, int precision) {
  const size_t planes = precision == MPA_PRECISION_FP32 ? 3 : 1;
  return align_up(planes * (size_t)M * K * 2, 256) + align_up(planes * (size_t)N * K * 2, 256);
}
size_t mpa_linear_workspace_bytes(int M, int N, int K) { return mpa_linear_workspace_bytes_ex(M, N, K, MPA_PRECISION_BF16); }

int mpa_linear_forward_ex(const float* x, const float* w, const float* bias, const float* residual, int M,
                          int N, int K, int act, int precision, float* out, void* ws, size_t ws_bytes,
                          void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  MPA_CHECK_ARG(M >= 0 && N > 0 && K > 0, "linear_forward: bad sizes %d %d %d", M, N, K);
  MPA_CHECK_ARG(act >= ACT_NONE && act <= ACT_SIGMOID, "linear_forward: bad activation %d", act);
  MPA_CHECK_ARG(precision == MPA_PRECISION_BF16 || precision == MPA_PRECISION_FP32,
                "linear_forward: bad precision %d", precision);
  if (M == 0) return MPA_OK;
  MPA_CHECK_ARG(x && w && out, "linear_forward: null pointer");
  const int split = precision == MPA_PRECISION_FP32 ? 3 : 1;
  Scratch scratch;
  int rc = scratch.acquire(ws, ws_bytes, mpa_linear_workspace_bytes_ex(M, N, K, precision), stream);
  if (rc != MPA_OK) return rc;
  __nv_bfloat16* xb = (__nv_bfloat16*)scratch.base;
  __nv_bfloat16* wb = (__nv_bfloat16*)((char*)scratch.base + align_up((size_t)split * M * K * 2, 256));
  {
    ProfScope ps("linear_operands_to_bf16", stream);
    CvtBatch cb;
    cb.src[0] = x; cb.dst[0] = xb; cb.n[0] = (long long)M * K; cb.plane[0] = split == 3 ? (long long)M * K : 0;
    cb.src[1] = w; cb.dst[1] = wb; cb.n[1] = (long long)N * K; cb.plane[1] = split == 3 ? (long long)N * K : 0;
    // enough CTAs to stream a [512 k, 512] activation matrix (DGCNN) at HBM speed
    const long long quads = ((long long)M * K + 1023) / 1024;
    const int gx = (int)(quads < 64 ? 64 : (quads > 148 * 8 ? 148 * 8 : quads));
    f32_to_bf16_batch_kernel<<<dim3(gx, 2), 256, 0, stream>>>(cb);
  }
  MPA_LAUNCH_CHECK();
  LinearEpilogue ep{bias, residual, out, nullptr, act};
  return launch_linear(xb, wb, M, N, K, ep, split == 3 ? "linear_fp32x3" : "linear_bf16", stream, split);
}

int mpa_linear_forward(const float* x, const float* w, const float* bias, const float* residual, int M,
                       int N, int K, int act, float* out, void* ws, size_t ws_bytes, void* stream_) {
  return mpa_linear_forward_ex(x, w, bias, residual, M, N, K, act, MPA_PRECISION_BF16, out, ws, ws_bytes,
                               stream_);
}

/* Pre-LN transformer encoder (nn.TransformerEncoder with norm_first=True,
 * batch_first, ReLU FFN, final LayerNorm), key-padding mask from `valid`.
 * tokens/out [B*P, D] fp32.  Per-layer parameter arrays hold `layers` device pointers each.
 * Training-mode dropout (the reference trains with p = 0.1, transformer.py:10,47): with
 * dropout_p > 0, `rng_state` (device, two 64-bit words the caller draws per forward) keys an
 * in-kernel Philox4x32-10 at the four sites of nn.TransformerEncoderLayer -- attention
 * probabilities, after out_proj, FFN hidden, after linear2 -- and `masks` receives the keep
 * masks per layer (mpa_transformer_mask_bytes) for the backward pass. */
size_t mpa_transformer_mask_bytes(int B, int P, int D, int H, int FF, int layers) {
  const size_t T = (size_t)B * P;
  return (size_t)layers * ((size_t)B * H * P * P + 2 * T * D + T * FF);
}


size_t mpa_transformer_workspace_bytes(int B, int P, int D, int FF, int layers) {
  const size_t T = (size_t)B * P;
  const size_t planes = 3;  // sized for the fp32-accurate mode (hi / mid / lo operand planes)
  size_t o = 0;
  o += align_up(planes * (size_t)layers * ((size_t)3 * D * D + (size_t)D * D + 2 * (size_t)FF * D) * 2, 256);  // weights
  o += align_up(T * D * 4, 256);                // residual stream x
  o += align_up(planes * T * D * 2, 256);       // LN output (operand)
  o += align_up(T * 3 * D * 4, 256);            // qkv fp32
  o += align_up(planes * T * D * 2, 256);       // attention output (operand)
  o += align_up(planes * T * FF * 2, 256);      // FFN hidden (operand)
  // cluster mode of the fused block: tiles x cl <= SMs partial results + the owner rows of x
  const size_t tiles = (T + LN_BM - 1) / LN_BM;
  o += align_up(((size_t)device_sms() + tiles) * LN_BM * FB_D * 4, 256);
  return o;
}

int mpa_transformer_forward(const float* tokens, const unsigned char* valid, int B, int P, int D, int H,
                            int FF, int layers, const float* const* in_proj_w,
                            const float* const* in_proj_b, const float* const* out_proj_w,
                            const float* const* out_proj_b, const float* const* lin1_w,
                            const float* const* lin1_b, const float* const* lin2_w,
                            const float* const* lin2_b, const float* const* norm1_w,
                            const float* const* norm1_b, const float* const* norm2_w,
                            const float* const* norm2_b, const float* final_norm_w,
                            const float* final_norm_b, float eps, float dropout_p,
                            unsigned long long* rng_state, unsigned char* masks, int precision,
                            float* out, void* ws, size_t ws_bytes, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  MPA_CHECK_ARG(precision == MPA_PRECISION_BF16 || precision == MPA_PRECISION_FP32,
                "transformer_forward: bad precision %d", precision);
  const int split = precision == MPA_PRECISION_FP32 ? 3 : 1;
  MPA_CHECK_ARG(B >= 0 && P > 0 && P <= 32, "transformer_forward: 1 <= P <= 32 parts (got %d)", P);
  MPA_CHECK_ARG(dropout_p >= 0.f && dropout_p < 1.f, "transformer_forward: dropout %f", dropout_p);
  const bool drop = dropout_p > 0.f;
  MPA_CHECK_ARG(!drop || (rng_state != nullptr && D % 32 == 0 && FF % 32 == 0),
                "transformer_forward: dropout needs rng_state and D, FF multiples of 32");
  MPA_CHECK_ARG(D % 32 == 0 && D <= 1024 && H > 0 && D % H == 0 && D / H <= 64 && FF % 8 == 0,
                "transformer_forward: unsupported dims D=%d H=%d FF=%d", D, H, FF);
  if (B == 0) return MPA_OK;
  const int T = B * P, hd = D / H;
  Scratch scratch;
  int rc = scratch.acquire(ws, ws_bytes, mpa_transformer_workspace_bytes(B, P, D, FF, layers), stream);
  if (rc != MPA_OK) return rc;
  char* p = (char*)scratch.base;
  // every weight matrix is stored as `split` consecutive planes [split][rows][cols]
  const size_t per_layer = ((size_t)3 * D * D + (size_t)D * D + 2 * (size_t)FF * D) * split;
  __nv_bfloat16* wts = (__nv_bfloat16*)p; p += align_up(3 * (size_t)layers * (per_layer / split) * 2, 256);
  float* x = (float*)p; p += align_up((size_t)T * D * 4, 256);
  __nv_bfloat16* xn = (__nv_bfloat16*)p; p += align_up(3 * (size_t)T * D * 2, 256);
  float* qkv = (float*)p; p += align_up((size_t)T * 3 * D * 4, 256);
  __nv_bfloat16* att = (__nv_bfloat16*)p; p += align_up(3 * (size_t)T * D * 2, 256);
  __nv_bfloat16* hid = (__nv_bfloat16*)p; p += align_up(3 * (size_t)T * FF * 2, 256);
  float* ffn_part = (float*)p;                                     // [<= SMs][128][256]
  float* ffn_x1 = ffn_part + (size_t)device_sms() * LN_BM * FB_D;  // [tiles][128][256]
  const long long pl_TD = split == 3 ? (long long)T * D : 0, pl_TF = split == 3 ? (long long)T * FF : 0;
  const size_t o_out = (size_t)3 * D * D * split, o_l1 = o_out + (size_t)D * D * split,
               o_l2 = o_l1 + (size_t)FF * D * split;  // offsets of out_proj / linear1 / linear2 in a layer

  MPA_CHECK_ARG(layers * 4 <= 64, "transformer_forward: at most 16 layers");
  {
    ProfScope ps("transformer_weights_to_bf16", stream);
    CvtBatch cb;
    for (int l = 0; l < layers; ++l) {
      __nv_bfloat16* wl = wts + (size_t)l * per_layer;
      cb.src[4 * l + 0] = in_proj_w[l];  cb.dst[4 * l + 0] = wl;         cb.n[4 * l + 0] = (long long)3 * D * D;
      cb.src[4 * l + 1] = out_proj_w[l]; cb.dst[4 * l + 1] = wl + o_out; cb.n[4 * l + 1] = (long long)D * D;
      cb.src[4 * l + 2] = lin1_w[l];     cb.dst[4 * l + 2] = wl + o_l1;  cb.n[4 * l + 2] = (long long)FF * D;
      cb.src[4 * l + 3] = lin2_w[l];     cb.dst[4 * l + 3] = wl + o_l2;  cb.n[4 * l + 3] = (long long)FF * D;
      for (int k = 0; k < 4; ++k) cb.plane[4 * l + k] = split == 3 ? cb.n[4 * l + k] : 0;
    }
    f32_to_bf16_batch_kernel<<<dim3(32, layers * 4), 256, 0, stream>>>(cb);
  }
  MPA_LAUNCH_CHECK();
  const int ln_blocks = (T * 32 + 255) / 256;
  const size_t att_smem = (size_t)(2 * P * hd + P * (hd + 1)) * sizeof(float);
  // D == 256 (the reference's d_model): the two GEMMs that finish a residual-stream row
  // (out_proj, FFN2) apply the NEXT LayerNorm in their epilogue -> 5 launches per layer
  const bool fused = D == LN_FUSED_WIDTH;
  if (!fused) MPA_CUDA(cudaMemcpyAsync(x, tokens, (size_t)T * D * 4, cudaMemcpyDeviceToDevice, stream));
  if (fused && layers > 0) {
    ProfScope ps("layernorm", stream);
    layernorm_kernel<<<ln_blocks, 256, 0, stream>>>(tokens, norm1_w[0], norm1_b[0], T, D, eps, xn, nullptr, pl_TD);
    MPA_LAUNCH_CHECK();
  }
  const size_t mask_layer = (size_t)B * H * P * P + 2 * (size_t)T * D + (size_t)T * FF;
  auto site = [&](int l, int k) {  // k: 0 attention, 1 after out_proj, 2 FFN hidden, 3 after linear2
    DropoutSpec d;
    if (!drop) return d;
    d.rng = rng_state;
    d.p = dropout_p;
    d.site = (unsigned)(4 * l + k);
    if (masks != nullptr) {
      unsigned char* m = masks + (size_t)l * mask_layer;
      const size_t offs[4] = {0, (size_t)B * H * P * P, (size_t)B * H * P * P + (size_t)T * D,
                              (size_t)B * H * P * P + (size_t)T * D + (size_t)T * FF};
      d.mask = m + offs[k];
    }
    return d;
  };
  for (int l = 0; l < layers; ++l) {
    const __nv_bfloat16* wl = wts + (size_t)l * per_layer;
    if (!fused) {
      { ProfScope ps("layernorm", stream);
        layernorm_kernel<<<ln_blocks, 256, 0, stream>>>(x, norm1_w[l], norm1_b[l], T, D, eps, xn, nullptr, pl_TD); }
      MPA_LAUNCH_CHECK();
    }
    static const bool no_attn = getenv("MPA_NO_FUSED_ATTN") != nullptr;  // A/B switch (tools/ab_step.sh)
    const bool fused_attn = fused && split == 1 && hd == AT_HD && H == 8 && !no_attn;
    if (fused_attn) {
      // QKV projection + attention in ONE kernel (whole shapes per CTA, heads split over CTAs)
      AttnArgs aa{};
      aa.b_in = in_proj_b[l]; aa.valid = valid; aa.att = att; aa.B = B; aa.P = P;
      aa.drop = site(l, 0);
      rc = launch_encoder_attn(xn, wl, aa, T, stream);
      if (rc != MPA_OK) return rc;
    } else {
    LinearEpilogue e_qkv{in_proj_b[l], nullptr, qkv, nullptr, ACT_NONE};
    rc = launch_linear(xn, wl, T, 3 * D, D, e_qkv, "linear_qkv", stream, split);
    if (rc != MPA_OK) return rc;
    { ProfScope ps("attention", stream);
      attention_kernel<<<B * H, 32 * (P < ATT_MAX_WARPS ? P : ATT_MAX_WARPS), att_smem, stream>>>(
          qkv, valid, B, P, H, hd, att, pl_TD, site(l, 0)); }
    MPA_LAUNCH_CHECK();
    }
    const bool last = l + 1 == layers;
    if (fused && split == 1 && FF % FB_CH == 0 && FF <= FB_MAX_FF) {
      // out_proj + LayerNorm2 + FFN + the next LayerNorm in ONE kernel (128 token rows per CTA)
      FfnBlockArgs fa{};
      fa.b_o = out_proj_b[l]; fa.b1 = lin1_b[l]; fa.b2 = lin2_b[l];
      fa.ln2_g = norm2_w[l]; fa.ln2_b = norm2_b[l];
      fa.lnn_g = last ? final_norm_w : norm1_w[l + 1];
      fa.lnn_b = last ? final_norm_b : norm1_b[l + 1];
      fa.x_in = l == 0 ? tokens : x;
      fa.x_out = x;
      fa.xn_out = last ? nullptr : xn;
      fa.out_f32 = last ? out : nullptr;
      fa.M = T; fa.FF = FF; fa.eps = eps;
      fa.drop1 = site(l, 1); fa.drop_h = site(l, 2); fa.drop2 = site(l, 3);
      fa.cl = ffn_cluster_size(T, FF); fa.part = ffn_part; fa.x1s = ffn_x1;
      rc = launch_ffn_block(att, wl + o_out, wl + o_l1, wl + o_l2, fa, stream);
      if (rc != MPA_OK) return rc;
      continue;
    }
    // x <- x + dropout1(out_proj(att))  [+ xn <- LayerNorm2(x)]
    LinearEpilogue e_o{out_proj_b[l], (fused && l == 0) ? tokens : x, x, nullptr, ACT_NONE};
    e_o.drop = site(l, 1);
    if (fused) { e_o.ln_gamma = norm2_w[l]; e_o.ln_beta = norm2_b[l]; e_o.ln_eps = eps; e_o.ln_out_bf16 = xn; }
    e_o.plane = pl_TD;
    rc = launch_linear(att, wl + o_out, T, D, D, e_o, "linear_out_proj", stream, split);
    if (rc != MPA_OK) return rc;
    if (!fused) {
      { ProfScope ps("layernorm", stream);
        layernorm_kernel<<<ln_blocks, 256, 0, stream>>>(x, norm2_w[l], norm2_b[l], T, D, eps, xn, nullptr, pl_TD); }
      MPA_LAUNCH_CHECK();
    }
    LinearEpilogue e_f1{lin1_b[l], nullptr, nullptr, hid, ACT_RELU};
    e_f1.drop = site(l, 2);
    e_f1.plane = pl_TF;
    rc = launch_linear(xn, wl + o_l1, T, FF, D, e_f1, "linear_ffn1", stream, split);
    if (rc != MPA_OK) return rc;
    // x <- x + FFN2(hid)  [+ LayerNorm1 of the next layer, or the final encoder norm]
    LinearEpilogue e_f2{lin2_b[l], x, (fused && last) ? (final_norm_w == nullptr ? out : nullptr) : x, nullptr, ACT_NONE};
    e_f2.drop = site(l, 3);
    if (fused && !last) {
      e_f2.ln_gamma = norm1_w[l + 1]; e_f2.ln_beta = norm1_b[l + 1]; e_f2.ln_eps = eps; e_f2.ln_out_bf16 = xn;
    } else if (fused && final_norm_w != nullptr) {
      e_f2.ln_gamma = final_norm_w; e_f2.ln_beta = final_norm_b; e_f2.ln_eps = eps; e_f2.ln_out_f32 = out;
    }
    e_f2.plane = pl_TD;
    rc = launch_linear(hid, wl + o_l2, T, D, FF, e_f2, "linear_ffn2", stream, split);
    if (rc != MPA_OK) return rc;
  }
  if (fused && layers > 0) return MPA_OK;
  if (final_norm_w != nullptr) {
    { ProfScope ps("layernorm", stream);
      layernorm_kernel<<<ln_blocks, 256, 0, stream>>>(layers > 0 ? x : tokens, final_norm_w, final_norm_b, T, D,
                                                      eps, nullptr, out, 0); }
    MPA_LAUNCH_CHECK();
  } else {
    MPA_CUDA(cudaMemcpyAsync(out, layers > 0 ? x : tokens, (size_t)T * D * 4, cudaMemcpyDeviceToDevice, stream));
  }
  return MPA_OK;
}

}  // extern "C"
