/*
 * mpa_b200.h -- C ABI of libmpa_b200.so, the B200 (sm_100a) native hot path of
 * multi_part_assembly.  Plain pointers and sizes only; no torch types.
 *
 * Conventions
 *   - every pointer named d_* / without a "_host" suffix in the function name is a
 *     DEVICE pointer valid on the current CUDA device; functions ending in
 *     `_host` take HOST pointers and do their own H2D/D2H copies.
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream,
 *     which is what the reference kernels use: chamfer_kernel.cu:147).
 *   - return value: 0 on success, negative MPA_ERR_* otherwise; the message is
 *     available from mpa_last_error() (thread-local).  Nothing is printed and
 *     the process is never exited (contrast pointnet2_ops cuda_utils.h:30-39).
 *   - workspaces: `ws` may be NULL, in which case the library allocates and frees
 *     scratch with cudaMallocAsync/cudaFreeAsync on `stream`; otherwise `ws` must
 *     hold at least the size the matching *_workspace_bytes() returns.
 *   - all arithmetic is fp32 (the reference force-casts Chamfer and rotations to
 *     float32: utils/chamfer/chamfer.py:14, utils/rotation.py:141).
 *
 * Reference paths are relative to /root/reference/multi_part_assembly.
 */
#ifndef MPA_B200_H_
#define MPA_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MPA_OK 0
#define MPA_ERR_INVALID_ARG (-1)
#define MPA_ERR_CUDA (-2)
#define MPA_ERR_WORKSPACE (-3)
#define MPA_ERR_UNSUPPORTED (-4)

/* Chamfer nearest-neighbour search algorithm selector (results are identical
 * bit for bit; this only picks the kernel). */
#define MPA_ALGO_AUTO 0
#define MPA_ALGO_BRUTE 1 /* one thread per query, all pairs, packed FP32x2 math */
#define MPA_ALGO_GRID 2  /* exact uniform-grid search (counting sort + ring walk) */

const char* mpa_last_error(void);
int mpa_version(void);
/* number of kernels this library has launched in this process (bench.py's
 * `gpu_launches` claim is the difference across the timed region). */
uint64_t mpa_launch_count(void);

/* Optional per-kernel CUDA-event timing (used by bench.py for the roofline of the
 * dominant kernel).  mpa_profile_report waits for the recorded events, writes
 * "kernel launches total_ms" lines into buf and clears the records. */
void mpa_profile_enable(int on);
size_t mpa_profile_report(char* buf, size_t cap);

/* ---- Chamfer distance -------------------------------------------------- */
/* Instrumentation for bench.py's roofline_fp32 (SURVEY.md 8d asks for pair evaluations per
 * second next to the GB/s figure): on != 0 switches the exact grid searches to an instantiation
 * that counts the candidate pairs it evaluates; out[3] (nullable) receives and clears the
 * counters of the current device {per-part pose search, shape-level pose search, plain clouds}.
 * No reference counterpart. */
int mpa_chamfer_pair_count(int on, unsigned long long* out);

/* Replaces chamfer_cuda.chamfer_forward (utils/chamfer/cuda/chamfer.cpp:8-11,21;
 * ChamferForward, chamfer_kernel.cu:116-168).
 *   xyz1 [B,N1,3], xyz2 [B,N2,3] contiguous fp32
 *   dist1 [B,N1], dist2 [B,N2] fp32 : squared distance to the nearest neighbour
 *   idx1 [B,N1], idx2 [B,N2] int64  : its index (lowest index on ties,
 *                                     chamfer_kernel.cu:82); may be NULL
 * Both directions run in one launch sequence on `stream`. */
size_t mpa_chamfer_forward_workspace_bytes(int B, int N1, int N2, int algo);
int mpa_chamfer_forward(const float* xyz1, const float* xyz2, int B, int N1, int N2,
                        float* dist1, int64_t* idx1, float* dist2, int64_t* idx2,
                        int algo, void* ws, size_t ws_bytes, void* stream);

/* Replaces chamfer_cuda.chamfer_backward (chamfer.cpp:13-19,22; ChamferBackward,
 * chamfer_kernel.cu:224-289).  grad_xyz1 [B,N1,3], grad_xyz2 [B,N2,3] are
 * OVERWRITTEN (the reference returns freshly zeroed tensors, :252-253).
 * Accumulates with fp32 atomics like the reference (:203-208), so the
 * summation order of coincident neighbours is unspecified. */
int mpa_chamfer_backward(const float* grad_dist1, const float* grad_dist2,
                         const float* xyz1, const float* xyz2, const int64_t* idx1,
                         const int64_t* idx2, int B, int N1, int N2, float* grad_xyz1,
                         float* grad_xyz2, void* stream);

/* Same as mpa_chamfer_forward with HOST buffers: copies inputs H2D, runs, copies
 * the four outputs D2H, synchronises `stream`.  This is the end-to-end entry the
 * `e2e` leg of bench.py times. */
int mpa_chamfer_forward_host(const float* h_xyz1, const float* h_xyz2, int B, int N1,
                             int N2, float* h_dist1, int64_t* h_idx1, float* h_dist2,
                             int64_t* h_idx2, int algo, void* stream);

/* ---- SE(3) on part point clouds ---------------------------------------- */
/* Replaces qrot / qtransform (utils/transforms.py:75-109), i.e.
 * pytorch3d.transforms.quaternion_apply broadcast over the N points of a part:
 *   out[p,i,:] = (q_p (x) (0,pts[p,i,:]) (x) conj(q_p))[1:] (+ trans[p,:])
 * quat [n_parts,4] real part first, trans [n_parts,3] or NULL (rot_pc),
 * pts/out [n_parts,N,3]. */
int mpa_se3_transform(const float* quat, const float* trans, const float* pts,
                      int n_parts, int N, float* out, void* stream);
/* Backward of the above: grad_pts (always), grad_quat/grad_trans (may be NULL). */
int mpa_se3_transform_backward(const float* quat, const float* pts, const float* grad_out,
                               int n_parts, int N, float* grad_pts, float* grad_quat,
                               float* grad_trans, void* stream);

/* Rotation3D's constructor rule for quaternions (utils/rotation.py:121-128): rows whose
 * norm is <= 0.5 (the all-zero quaternions of padded parts) become the identity (1,0,0,0).
 * quat / out [n,4] fp32, 16-byte aligned; out may alias quat. */
int mpa_quat_fix_zero(const float* quat, long long n, float* out, void* stream);

/* ---- fused pose losses -------------------------------------------------- */
/* Fused SE(3) + bidirectional Chamfer for the two Chamfer losses of
 * utils/loss.py.  pts [B,P,N,3]; quat1/quat2 [B,P,4]; trans1/trans2 [B,P,3] or
 * NULL; valids [B,P] (1.0 valid / 0.0 padded).
 *
 * mode MPA_CD_PART  : rot_points_cd_loss (loss.py:113-138): every part is its own
 *                     cloud.  Padded parts are skipped (their dist is 0).
 * mode MPA_CD_SHAPE : shape_cd_loss (loss.py:141-202): padded parts are filled
 *                     with 1e3 (:173-175), all parts of a shape form one cloud
 *                     of P*N points.  Padded query points get dist 0; padded
 *                     target parts stay candidates exactly as in the reference.
 * Outputs (device): dist1/dist2 [B,P,N] fp32, idx1/idx2 [B,P,N] int32 index into
 * the cloud (part-local for PART, shape-level p*N+i for SHAPE; -1 for skipped),
 * pts1/pts2 [B,P,N,3] transformed clouds or NULL. */
#define MPA_CD_PART 0
#define MPA_CD_SHAPE 1
size_t mpa_pose_chamfer_workspace_bytes(int B, int P, int N, int mode);
int mpa_pose_chamfer(const float* pts, const float* quat1, const float* trans1,
                     const float* quat2, const float* trans2, const float* valids,
                     int B, int P, int N, int mode, float* dist1, int32_t* idx1,
                     float* dist2, int32_t* idx2, float* pts1, float* pts2, void* ws,
                     size_t ws_bytes, void* stream);

/* Backward of mpa_pose_chamfer w.r.t. the two poses (the points carry no
 * gradient: loss.py:172 detaches them, and part_pcs is data).  Chains
 * ChamferBackward (chamfer_kernel.cu:175-210) into the SE(3) backward.
 * grad_quat{1,2} [B,P,4] and grad_trans{1,2} [B,P,3] may be NULL (e.g. the ground-truth pose). */
size_t mpa_pose_chamfer_backward_workspace_bytes(int B, int P, int N);
int mpa_pose_chamfer_backward(const float* grad_dist1, const float* grad_dist2, const float* pts,
                              const float* quat1, const float* quat2, const float* valids,
                              const float* pts1, const float* pts2, const int32_t* idx1,
                              const int32_t* idx2, int B, int P, int N, int mode, float* grad_quat1,
                              float* grad_trans1, float* grad_quat2, float* grad_trans2, void* ws,
                              size_t ws_bytes, void* stream);

/* Fused reductions of the geometric loss terms as BaseModel._calc_loss assembles
 * them (models/modules/base_model.py:256-290; utils/loss.py:22-35, 59-86, 89-110,
 * 131-134, 185-198), forward only.  part_dist{1,2} and shape_dist{1,2} [B,P,N] are the
 * outputs of mpa_pose_chamfer in PART / SHAPE mode (0 on padded parts).
 * weights (DEVICE pointer, 5 floats) = {trans_loss_w, rot_pt_cd_loss_w, transform_pt_cd_loss_w, rot_loss_w,
 * rot_pt_l2_loss_w} (0 for a disabled term).  terms [6,B] = trans_loss,
 * rot_pt_cd_loss, transform_pt_cd_loss, rot_loss, rot_pt_l2_loss, weighted total.
 * training: shape_cd_loss divides by P*N (:185-193), else per-part means (:195-198). */
size_t mpa_geometric_losses_workspace_bytes(int B, int P);
int mpa_geometric_losses(const float* pts, const float* quat1, const float* trans1,
                         const float* quat2, const float* trans2, const float* valids,
                         const float* part_dist1, const float* part_dist2,
                         const float* shape_dist1, const float* shape_dist2, int B, int P, int N,
                         int training, int want_rot_l2, const float* weights, float* terms,
                         void* ws, size_t ws_bytes, void* stream);

/* ---- Hungarian matching of equivalent parts -------------------------------- */
/* Replaces the SciPy call of BaseModel._linear_sum_assignment (models/modules/
 * base_model.py:175-176): `n_problems` square min-cost assignment problems in one launch.
 * costs: the row-major fp32 cost matrices back to back (problem g starts at
 * cost_offsets[g], is sizes[g] x sizes[g]); col_of_row[out_offsets[g] + r] = column
 * assigned to row r (SciPy's col_ind; row_ind is 0..p-1).  All pointers are device
 * pointers.  Same algorithm as SciPy's rectangular_lsap (Crouse 2016), float64 duals, so
 * the assignment is identical; sizes[g] <= max_size <= 64.  Matrices with non-finite
 * entries that make a problem infeasible yield the identity. */
int mpa_lsap_batched(const float* costs, const int32_t* cost_offsets, const int32_t* sizes,
                     const int32_t* out_offsets, int n_problems, int max_size, int32_t* col_of_row,
                     void* stream);

/* ---- PointNet part encoder --------------------------------------------- */
/* Replaces PointNet.forward with global_feat=True (models/modules/encoder/
 * pointnet.py:29-41) fused with the valid-part selection of _extract_part_feats
 * (models/pn_transformer/network.py:59-68): pts [n_parts,N,3]; valids [n_parts]
 * or NULL (padded parts are skipped, excluded from the BatchNorm statistics and
 * get zero features); conv_w[5] = conv{1..5}.weight ([Cout,Cin] fp32, no bias);
 * bn_* [5] = bn{1..5}.{weight,bias,running_mean,running_var}; feats [n_parts,F],
 * F in {128,256}.  training != 0: batch statistics (biased variance) and the
 * running statistics are updated in place with `momentum`; 0: running statistics.
 * GEMMs run on tcgen05 tensor cores with bf16 operands and fp32 accumulation
 * (the reference's --fp16 autocast analogue); BatchNorm math is fp32. */
size_t mpa_pointnet_workspace_bytes(int n_parts);
/* Workspace with room for the activation stash (128 B per point, two buffers): the
 * training launches then start from the previous launch's last operand tile instead of
 * recomputing the earlier layers from the points.  Results are identical either way. */
size_t mpa_pointnet_workspace_bytes_n(int n_parts, int N);
int mpa_pointnet_forward(const float* pts, const float* valids, int n_parts, int N, int F,
                         const float* const* conv_w, const float* const* bn_gamma,
                         const float* const* bn_beta, float* const* bn_running_mean,
                         float* const* bn_running_var, int training, float eps, float momentum,
                         float* feats, void* ws, size_t ws_bytes, void* stream);
/* Same, and (training, bn_batch_stats != NULL) also returns what a backward pass needs of the
 * five BatchNorm layers: [5][4][256] floats = batch mean, 1/sqrt(var + eps), scale, shift per
 * channel, followed by one float = the number of points the statistics were taken over. */
int mpa_pointnet_forward_ex(const float* pts, const float* valids, int n_parts, int N, int F,
                            const float* const* conv_w, const float* const* bn_gamma,
                            const float* const* bn_beta, float* const* bn_running_mean,
                            float* const* bn_running_var, int training, float eps, float momentum,
                            float* feats, float* bn_batch_stats, void* ws, size_t ws_bytes, void* stream);

/* ---- PointNet backward: BatchNorm / ReLU / max-pool between the GEMMs ---------- */
/* Training-step backward of PointNet.forward (models/modules/encoder/pointnet.py:29-41,
 * i.e. what autograd derives from nn.Conv1d/BatchNorm1d/ReLU/torch.max there) on
 * point-major activations: z, a, da, dz are [M = n_parts*N, C] bf16 with C contiguous
 * (C in {64,128,256}), so the convolutions and their two backward products are plain
 * GEMMs done by the caller.  valids [n_parts] or NULL: rows of padded parts are left
 * out of every statistic and get zero outputs.
 *   mpa_bn_stats     sums[0..C) = sum z, sums[C..2C) = sum z^2   (fp64, device)
 *   mpa_bn_finalize  -> mean, rstd, scale = gamma*rstd, shift = beta - mean*scale, count
 *   mpa_bn_act       a = relu?(z*scale + shift)
 *   mpa_bn_backward  dz = gamma*rstd*(dy - S1/count - zhat*S2/count); sums receives
 *                    S1 = sum dy (= d beta), S2 = sum dy*zhat (= d gamma).  dy is
 *                    da * [z*scale+shift > 0] when `da` is given (ReLU layers), else the
 *                    max-pool scatter g[part,c] * [arg[part,c] == point] (last layer).
 *   mpa_pool_argmax  arg[part,c] = first point maximising BN(z)[.,c] over the part */
int mpa_bn_stats(const void* z, long long M, int C, int N, const float* valids, double* sums,
                 void* stream);
int mpa_bn_finalize(const double* sums, int C, int n_parts, int N, const float* valids,
                    const float* gamma, const float* beta, float eps, float* mean, float* rstd,
                    float* scale, float* shift, float* count, void* stream);
int mpa_bn_act(const void* z, const float* scale, const float* shift, int relu, long long M, int C,
               int N, const float* valids, void* a, void* stream);
int mpa_bn_backward(const void* da, const float* g, const int32_t* arg, const void* z,
                    const float* mean, const float* rstd, const float* gamma, const float* beta,
                    const float* count, long long M, int C, int N, const float* valids, double* sums,
                    void* dz, void* stream);
int mpa_pool_argmax(const void* z, const float* scale, int n_parts, int N, int C, int32_t* arg,
                    void* stream);

/* ---- token-level dense layers on tcgen05 + TMA ---------------------------- */
/* Y = act(X W^T + bias) (+ residual): X [M,K], W [N,K] (an nn.Linear weight),
 * bias [N] or NULL, residual [M,N] or NULL, out [M,N]; fp32 in memory, bf16
 * tensor-core operands with fp32 accumulation.  act: 0 none, 1 ReLU,
 * 2 LeakyReLU(0.2), 3 sigmoid.  K must be a multiple of 8.  Replaces the cuBLAS GEMMs of
 * PoseRegressor (models/modules/regressor.py:45-68). */
#define MPA_PRECISION_BF16 0 /* bf16 operands, fp32 accumulation */
#define MPA_PRECISION_FP32 1 /* fp32-accurate: 3 bf16 planes per operand, 6 tensor-core products per k-step */
size_t mpa_linear_workspace_bytes(int M, int N, int K);
/* Same with a precision selector: MPA_PRECISION_FP32 replaces the fp32 cuBLAS/cuDNN GEMMs
 * of the reference's default (non --fp16) mode, e.g. the 1x1 convolutions of DGCNN
 * (models/modules/encoder/dgcnn.py:51-68). */
size_t mpa_linear_workspace_bytes_ex(int M, int N, int K, int precision);
int mpa_linear_forward_ex(const float* x, const float* w, const float* bias, const float* residual,
                          int M, int N, int K, int act, int precision, float* out, void* ws,
                          size_t ws_bytes, void* stream);
int mpa_linear_forward(const float* x, const float* w, const float* bias, const float* residual,
                       int M, int N, int K, int act, float* out, void* ws, size_t ws_bytes,
                       void* stream);

/* Output rows of PoseRegressor (models/modules/regressor.py:61-67): rot [T,4] =
 * rot_head(feats) (L2-normalised when `normalize`), trans [T,3] = trans_head(feats);
 * feats [T,K] fp32, weights as nn.Linear stores them.  Forward only. */
int mpa_pose_outputs(const float* feats, int T, int K, const float* rot_w, const float* rot_b,
                     const float* trans_w, const float* trans_b, int normalize, float* rot,
                     float* trans, void* stream);

/* The whole PoseRegressor.forward (models/modules/regressor.py:58-68) in one launch,
 * fp32: feats [T,K0] -> LeakyReLU(fc0: K0->H1) -> LeakyReLU(fc1: H1->H2) -> rot [T,4]
 * (rot_head, L2-normalised when `normalize`) and trans [T,3] (trans_head).  Weights as
 * nn.Linear stores them ([out,in]).  Forward only. */
int mpa_pose_head_forward(const float* feats, int T, int K0, const float* fc0_w, const float* fc0_b,
                          int H1, const float* fc1_w, const float* fc1_b, int H2, const float* rot_w,
                          const float* rot_b, const float* trans_w, const float* trans_b, int normalize,
                          float* rot, float* trans, void* stream);

/* Replaces TransformerEncoder.forward (models/pn_transformer/transformer.py:63-79),
 * i.e. nn.TransformerEncoder built at :23-34: `layers` pre-LN encoder layers
 * (MHA with H heads, ReLU FFN of width FF) + final LayerNorm (final_norm_w may be
 * NULL), key-padding mask `valid` [B*P] bytes (1 = valid; NULL = all valid).
 * tokens/out [B*P, D] fp32.  Every *_w/*_b argument is an array of `layers`
 * device pointers laid out as torch stores them (in_proj_weight [3D,D] = [Wq;Wk;Wv]).
 * Training-mode dropout (the reference trains with p = 0.1, transformer.py:10,47): with
 * dropout_p > 0, rng_state (device, two 64-bit key words the caller draws for THIS forward)
 * keys an in-kernel Philox4x32-10 at the four dropout sites of nn.TransformerEncoderLayer
 * (attention probabilities, after out_proj, FFN hidden, after linear2); `masks` (nullable,
 * mpa_transformer_mask_bytes, per layer [B,H,P,P | T,D | T,FF | T,D] keep bytes) receives the
 * masks for the backward pass.  dropout_p = 0: eval.  precision: MPA_PRECISION_BF16 (bf16
 * operands, the analogue of the reference's --fp16 autocast) or MPA_PRECISION_FP32 (the
 * reference default, scripts/train.py:88: every operand is carried as three bf16 planes and
 * six tensor-core products per k-step reproduce an fp32 GEMM to ~1e-6). */
size_t mpa_transformer_workspace_bytes(int B, int P, int D, int FF, int layers);
size_t mpa_transformer_mask_bytes(int B, int P, int D, int H, int FF, int layers);
int mpa_transformer_forward(const float* tokens, const unsigned char* valid, int B, int P, int D,
                            int H, int FF, int layers, const float* const* in_proj_w,
                            const float* const* in_proj_b, const float* const* out_proj_w,
                            const float* const* out_proj_b, const float* const* lin1_w,
                            const float* const* lin1_b, const float* const* lin2_w,
                            const float* const* lin2_b, const float* const* norm1_w,
                            const float* const* norm1_b, const float* const* norm2_w,
                            const float* const* norm2_b, const float* final_norm_w,
                            const float* final_norm_b, float eps, float dropout_p,
                            unsigned long long* rng_state, unsigned char* masks, int precision,
                            float* out, void* ws, size_t ws_bytes, void* stream);

/* ---- DGCNN: k-NN graph and EdgeConv aggregation -------------------------- */
/* Replaces knn (models/modules/encoder/dgcnn.py:8-15).  x [n,N,C] fp32 with the
 * points of a part as ROWS (the reference passes [n,C,N]; transpose first);
 * idx [n,N,k] int32, indices local to the part, best score first, ties to the
 * lower index.  Scores are the reference's expanded form evaluated in fp32 with
 * a sequential-channel FMA dot product (the oracle's order).  N <= 2048.
 * valids [n] (nullable): parts flagged 0 (padding) are skipped on the device --
 * their idx rows are left untouched -- so the caller never compacts on the host
 * (the reference's boolean-mask gather, models/dgl/network.py:90-99). */
size_t mpa_knn_workspace_bytes(int n, int N);
/* Workspace that also holds the operand planes and the score slab of the tensor-core
 * scoring path (tcgen05 candidate filter + exact decision at the k-th boundary, same index
 * sets); with the smaller size above mpa_knn runs the CUDA-core tile kernel. */
size_t mpa_knn_workspace_bytes_c(int n, int N, int C);
int mpa_knn(const float* x, const float* valids, int n, int N, int C, int k, int32_t* idx, void* ws,
            size_t ws_bytes, void* stream);

/* EdgeConv after the split W [xj-xi ; xi] = W1 xj + (W2-W1) xi (dgcnn.py:18-38,
 * 81-95): uv [n*N, 2*Co] holds u = W1 x in the first Co columns and
 * v = (W2-W1) x in the last Co; idx [n*N,k] from mpa_knn.  Writes the max and
 * min over the k edges of u_j + v_i (ymax/ymin [n*N,Co]) and sums [Co,2] =
 * (sum, sum of squares) over all n*N*k edges in fp64 -- the inputs of
 * BatchNorm2d batch statistics + LeakyReLU + max over k.  valids [n] (nullable):
 * points of parts flagged 0 get ymax = ymin = 0 and stay out of the sums. */
size_t mpa_edge_aggregate_workspace_bytes(long long M, int Co);
int mpa_edge_aggregate(const float* uv, const int32_t* idx, const float* valids, int n, int N,
                       int Co, int k, float* ymax, float* ymin, double* sums, void* ws,
                       size_t ws_bytes, void* stream);

/* Second pass of an EdgeConv layer: BatchNorm2d (batch statistics from the aggregate's sums
 * over n_valid*N*k_edges edges, or running statistics; torch's running update applied) +
 * LeakyReLU(slope) + the max over the k edges (dgcnn.py:81-95), from ymax / ymin.  Writes the
 * layer output [n*N,Co] (`out`, nullable) and its slice [.., c0:c0+Co] of the [n*N,ldc]
 * concatenation conv5 reads (`out_cat`, nullable): no torch.cat. */
int mpa_edgeconv_finish(const float* ymax, const float* ymin, const double* sums, const float* valids,
                        int n, int N, int Co, int k_edges, const float* bn_w, const float* bn_b,
                        float* running_mean, float* running_var, int training, float momentum,
                        float eps, float slope, float* out, float* out_cat, int ldc, int c0,
                        void* stream);

/* conv5 epilogue (dgcnn.py:97-107): BatchNorm1d (batch statistics over the valid parts' n_valid*N
 * points, or running statistics) + LeakyReLU(slope) on y [n*N,F], then max and mean over the N
 * points of every part: g [n, 2F] = [max | mean] (zeros for padded parts). */
size_t mpa_bn_pool_workspace_bytes(int n, int F);
int mpa_bn_pool(const float* y, const float* valids, int n, int N, int F, const float* bn_w,
                const float* bn_b, float* running_mean, float* running_var, int training,
                float momentum, float eps, float slope, float* g, void* ws, size_t ws_bytes,
                void* stream);

/* BaseModel._match_parts (models/modules/base_model.py:181-238) for a whole batch: for every
 * group of geometrically equivalent parts the p x p cost matrix of :162-174 (Chamfer distance
 * between part i under its predicted pose and part j under its ground-truth pose on the group's
 * n-point random subsample), the min-cost assignment (SciPy's linear_sum_assignment algorithm,
 * :175) and the permuted ground-truth poses (:229-233), in three launches.  The group table is
 * built on the host, which also draws the subsamples with the reference's RNG calls
 * (torch.randperm per group, :165).  table (device int32): shape [G] | size [G] | cost offset
 * [G] | row offset [G] | members [G,32] | subsample [G,n] | entries (g<<16 | i<<8 | j)
 * [n_pairs].  new_trans / new_quat [B,P,3/4] must hold copies of the ground truth on entry.
 * costs_out / col_out (nullable): the cost matrices and assignments for inspection. */
size_t mpa_match_parts_workspace_bytes(int n_pairs, int total_rows);
int mpa_match_parts(const float* pts, const float* pred_quat, const float* pred_trans,
                    const float* gt_quat, const float* gt_trans, int B, int P, int N, int n,
                    const int32_t* table, int G, int n_pairs, int total_rows, int max_size,
                    float* new_trans, float* new_quat, float* costs_out, int32_t* col_out, void* ws,
                    size_t ws_bytes, void* stream);

/* ---- PointNet++ set abstraction (encoder/pointnet2; pointnet2_ops CUDA extension) ---- */
/* furthest_point_sampling (_ext-src/src/sampling_gpu.cu:74-177) + gather_operation of the picked
 * centroids (pointnet2_modules.py:53-61): xyz [B,n,3] -> idx [B,m] int32 (first pick = point 0;
 * points with |p|^2 <= 1e-3 are skipped; ties follow the reference's strided tree reduction for
 * its block size opt_n_threads(n)), new_xyz [B,m,3] (nullable). */
int mpa_furthest_point_sample(const float* xyz, int B, int n, int m, int32_t* idx, float* new_xyz,
                              void* stream);
/* ball_query (_ext-src/src/ball_query_gpu.cu:13-48): idx [B,m,nsample] = the first nsample
 * points (index order) within `radius` of each centroid, padded with the first hit (zeros when
 * there is none). */
int mpa_ball_query(const float* xyz, const float* new_xyz, int B, int n, int m, float radius,
                   int nsample, int32_t* idx, void* stream);
/* group_points + QueryAndGroup / GroupAll (group_points_gpu.cu:12-32, pointnet2_utils.py:309-392)
 * as the row matrix of the shared MLP: out [B*m*nsample, ld] rows = [xyz[idx] - centroid |
 * feats[idx]] (idx == NULL: GroupAll, rows [B*n, ld] = [xyz | feats]); feats channels-last
 * [B,n,C] (nullable when C == 0); columns >= 3 + C are zero padding. */
int mpa_group_rows(const float* xyz, const float* new_xyz, const float* feats, const int32_t* idx,
                   int B, int n, int m, int nsample, int C, int ld, float* out, void* stream);
/* fp64 (sum, sum of squares) per column of y [n_blocks*R, F]: BatchNorm batch statistics of a
 * shared-MLP layer (pointnet2_modules.py:10-22; encoder/pointnet.py:29-41 in fp32 mode);
 * valids [n_blocks] (nullable): blocks (parts) flagged 0 stay out of the sums. */
size_t mpa_column_stats_workspace_bytes(int n_blocks, int F);
int mpa_column_stats(const float* y, const float* valids, int n_blocks, int R, int F, double* sums,
                     void* ws, size_t ws_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MPA_B200_H_ */
