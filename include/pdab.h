/*
 * pdab.h — C ABI of libpdab.so: B200-native (sm_100a) kernels for the PDA-SSD
 * point-backbone hot path of Geo3DSmart/PDANet.
 *
 * This is the drop-in boundary.  Each entry point replaces one function of the
 * reference's two pybind11 extension modules; the citation after "replaces:"
 * is the reference interface (paths relative to the reference root,
 * PB = pcdet/ops/pointnet2/pointnet2_batch, IOU = pcdet/ops/iou3d_nms).
 *
 * Conventions (same as the reference unless stated):
 *   - all pointers are DEVICE pointers to contiguous row-major arrays, except
 *     where a parameter is named *_host;
 *   - the CALLER allocates every output (PB/pointnet2_utils.py:25-26,83,200,246);
 *   - fp32 data, int32 indices, int64 NMS keep lists;
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream);
 *     the reference always launches on the legacy default stream;
 *   - nothing allocates, synchronises or touches process-wide state unless stated.  The three launch-policy setters
 *     (pdab_set_persistent_ctas, pdab_set_fps_max_cluster, pdab_set_cta_pairs) are THREAD-LOCAL: they change how the calling
 *     host thread's later launches are shaped (never the results), two pipelines driven from two threads do not see each
 *     other's settings, and a CUDA graph keeps the shape it was captured with;
 *   - return value: 0 on success, a positive cudaError_t value on a CUDA error,
 *     a negative PDAB_E* code on a bad argument.  The library never aborts the
 *     process (the reference calls exit(-1) on launch failure,
 *     PB/src/sampling_gpu.cu:39-43).
 */
#ifndef PDAB_H_
#define PDAB_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PDAB_EINVAL (-1)      /* bad size / null pointer */
#define PDAB_EUNSUPPORTED (-2) /* size outside what the kernels cover */

typedef void *pdab_stream_t;

/* Library / device info. */
const char *pdab_version(void);
/* Human-readable text for a return code of any function below. */
const char *pdab_error_string(int code);

/* ---- pointnet2_batch_cuda -------------------------------------------------- */

/* Distance farthest-point sampling.
 * replaces: farthest_point_sampling_wrapper, PB/src/pointnet2_api.cpp:22,
 *           PB/src/sampling.cpp:34-43, kernel PB/src/sampling_gpu.cu:93-253.
 * xyz (B,N,3); temp (B,N) in/out: running min-distances, caller pre-fills 1e10
 * (values must be >= 0), holds the final min-distances on return as in the
 * reference; idx (B,m) out.  idx[.,0] = 0.  Ties at the maximum resolve exactly
 * as the reference's 1024-thread tree does: argmin over tied k of
 * (bitrev_L(k mod BS), k), BS = largest power of two <= min(N,1024), L = log2 BS.
 * N <= 262144. */
int pdab_fps(int b, int n, int m, const float *xyz, float *temp, int *idx, pdab_stream_t stream);

/* Farthest-point sampling on a precomputed (B,N,N) distance matrix (F-FPS).
 * replaces: furthest_point_sampling_with_dist_wrapper, PB/src/pointnet2_api.cpp:23,
 *           PB/src/sampling.cpp:46-56, kernel PB/src/sampling_gpu.cu:256-416.
 * N <= 16384. */
int pdab_fps_with_dist(int b, int n, int m, const float *dist, float *temp, int *idx, pdab_stream_t stream);

/* out[b,c,j] = points[b,c,idx[b,j]].
 * replaces: gather_points_wrapper, PB/src/pointnet2_api.cpp:19, PB/src/sampling_gpu.cu:8-44. */
int pdab_gather_points(int b, int c, int n, int npoints, const float *points, const int *idx, float *out,
                       pdab_stream_t stream);
/* grad_points[b,c,idx[b,j]] += grad_out[b,c,j]   (grad_points pre-zeroed by the caller).
 * replaces: gather_points_grad_wrapper, PB/src/pointnet2_api.cpp:20, PB/src/sampling_gpu.cu:46-83. */
int pdab_gather_points_grad(int b, int c, int n, int npoints, const float *grad_out, const int *idx,
                            float *grad_points, pdab_stream_t stream);

/* Ball query: first `nsample` points (in index order) with |p - c|^2 < radius^2;
 * unfilled slots repeat the first hit; a ball with no hit leaves its row untouched
 * (the caller pre-zeroes idx, PB/pointnet2_utils.py:246).
 * replaces: ball_query_wrapper, PB/src/pointnet2_api.cpp:13, PB/src/ball_query_gpu.cu:9-67.
 * new_xyz (B,M,3), xyz (B,N,3), idx (B,M,nsample).  nsample <= 256. */
int pdab_ball_query(int b, int n, int m, float radius, int nsample, const float *new_xyz, const float *xyz,
                    int *idx, pdab_stream_t stream);
/* replaces: ball_query_dilated_wrapper, PB/src/pointnet2_api.cpp:14, PB/src/ball_query_gpu.cu:70-139. */
int pdab_ball_query_dilated(int b, int n, int m, float max_radius, float min_radius, int nsample,
                            const float *new_xyz, const float *xyz, int *idx, pdab_stream_t stream);

/* pdab_ball_query through a hashed cell list: the call first buckets every scene's points by cell (edge = radius + 0.1 %), the
 * centres then test the 27 cells around them and keep the nsample smallest indices among the hits — the same idx, bit for bit,
 * at ~N / 30 of the distance tests when the balls are small against the cloud.  Centres whose 27 cells hold many more points
 * than slots (dense balls) are handed to the in-order scan kernel, which stops at nsample hits.
 * workspace: pdab_ball_query_grid_workspace_bytes(b, n, m) bytes on the device, 16-byte aligned.  nsample <= 256. */
size_t pdab_ball_query_grid_workspace_bytes(int b, int n, int m);
int pdab_ball_query_grid(int b, int n, int m, float radius, int nsample, const float *new_xyz, const float *xyz, int *idx,
                         void *workspace, pdab_stream_t stream);

/* out[b,c,j,s] = points[b,c,idx[b,j,s]].
 * replaces: group_points_wrapper, PB/src/pointnet2_api.cpp:16, PB/src/group_points_gpu.cu:53-92. */
int pdab_group_points(int b, int c, int n, int npoints, int nsample, const float *points, const int *idx,
                      float *out, pdab_stream_t stream);
/* replaces: group_points_grad_wrapper, PB/src/pointnet2_api.cpp:17, PB/src/group_points_gpu.cu:14-50. */
int pdab_group_points_grad(int b, int c, int n, int npoints, int nsample, const float *grad_out, const int *idx,
                           float *grad_points, pdab_stream_t stream);

/* Three nearest neighbours of every `unknown` point among the `known` points of its scene.
 * replaces: three_nn_wrapper, PB/src/pointnet2_api.cpp:27, PB/src/interpolate_gpu.cu:16-78.
 * unknown (B,N,3), known (B,M,3) -> dist2 (B,N,3) squared distances ascending, idx (B,N,3); ties keep the lowest index;
 * with fewer than three candidates the missing entries are (+inf, 0) as in the reference. */
int pdab_three_nn(int b, int n, int m, const float *unknown, const float *known, float *dist2, int *idx,
                  pdab_stream_t stream);
/* out[b,c,i] = sum_k weight[b,i,k] * points[b,c,idx[b,i,k]].
 * replaces: three_interpolate_wrapper, PB/src/pointnet2_api.cpp:28, PB/src/interpolate_gpu.cu:84-120.
 * points (B,C,M), idx (B,N,3), weight (B,N,3) -> out (B,C,N). */
int pdab_three_interpolate(int b, int c, int m, int n, const float *points, const int *idx, const float *weight,
                           float *out, pdab_stream_t stream);
/* grad_points[b,c,idx[b,i,k]] += grad_out[b,c,i] * weight[b,i,k]   (grad_points pre-zeroed by the caller).
 * replaces: three_interpolate_grad_wrapper, PB/src/pointnet2_api.cpp:29, PB/src/interpolate_gpu.cu:127-166. */
int pdab_three_interpolate_grad(int b, int c, int n, int m, const float *grad_out, const int *idx, const float *weight,
                                float *grad_points, pdab_stream_t stream);

/* ---- roiaware_pool3d_cuda (training-side target assignment of PDA-SSD) ----------- */

/* box_idx_of_points[b,i] = lowest k with point i of scene b inside box k (rotated box, 1e-5 margin in x/y), untouched
 * otherwise (the caller pre-fills -1, pcdet/ops/roiaware_pool3d/roiaware_pool3d_utils.py:38).
 * replaces: points_in_boxes_gpu, pcdet/ops/roiaware_pool3d/src/roiaware_pool3d.cpp (pybind name),
 *           kernel roiaware_pool3d_kernel.cu:313-359.
 * boxes (B,T,7) [x,y,z,dx,dy,dz,heading], pts (B,M,3). */
int pdab_points_in_boxes(int batch_size, int boxes_num, int pts_num, const float *boxes, const float *pts,
                         int *box_idx_of_points, pdab_stream_t stream);

/* ---- fused ops (additive; no reference native unit, they replace Python glue) ---- */

/* Class-aware ("ctr_aware") sampling: idx[b,:] = indices of the npoint largest
 * max_c cls[b,k,c], ordered by (value desc, index asc).
 * replaces: cls.max(-1) -> sigmoid -> torch.topk -> .int(), PB/pointnet2_modules.py:761-770.
 * cls (B,N,C) fp32, idx (B,npoint).  N <= 65536. */
int pdab_topk_ctr(int b, int n, int c, int npoint, const float *cls, int *idx, pdab_stream_t stream);

/* PDA grouper: ball query + grouping + Gaussian density + direction encoding in one pass.
 * replaces: QueryAndGroup_alone_grouped_density_directional.forward, PB/pointnet2_utils.py:567-614.
 * xyz (B,N,3), new_xyz (B,M,3), features (B,C,N), out (B,7+C,M,nsample) with channels
 * [xyz(3, not centred), density, direction(3), features(C)];
 * idx_out (B,M,nsample) optional (may be NULL), written in full. */
int pdab_pda_group(int b, int c, int n, int m, float radius, int nsample, const float *xyz, const float *new_xyz,
                   const float *features, float *out, int *idx_out, pdab_stream_t stream);

/* Token-major PDA grouper: same search and encoding as pdab_pda_group, laid out for the per-token consumers of the
 * PDA block.  features_t (B,N,C) is the POINT-major copy of the features; out (B,M,nsample,pitch) holds one row per
 * (centre, neighbour): [0..2] xyz, [3] density, [4..6] direction, [7] 0, [8..8+C) features.
 * pitch >= 8+C, pitch % 4 == 0, C % 4 == 0.  idx_out optional.
 * replaces: the same Python grouper + the permute/contiguous copies that feed the transformer,
 *           PB/pointnet2_utils.py:567-614, PB/pointnet2_modules.py:879-927. */
int pdab_pda_group_tokens(int b, int c, int n, int m, float radius, int nsample, int pitch, const float *xyz,
                          const float *new_xyz, const float *features_t, float *out, int *idx_out,
                          pdab_stream_t stream);

/* Fused PDA token encoder: ordered ball query -> gather -> density + direction -> relative position encoding ->
 * position MLP (12 -> c/2 -> c, folded BN, ReLU) ; density / neighbourhood max -> DensityNet (1 -> 16 -> 8 -> 1, ReLU);
 * out row = LayerNorm(cat[pos, feat * scale, feat, glob[centre]]) — the (b*m*nsample, 4c) input of the transformer.
 * xyz (B,N,3), new_xyz (B,M,3), features_t (B,N,c) point-major, glob (B*M, c), out (B*M*nsample, 4c); c in {64, 128},
 * nsample in {16, 32}.  params = W1 [c/2][12] | b1 [c/2] | W2^T [c/2][c] | b2 [c] | DensityNet w1[16] b1[16] w2[8][16]
 * b2[8] w3[8] b3[1] pad[3] | gamma [4c] | beta [4c]  (pdab_pda_encode_param_floats(c) floats, 0 = unsupported c).
 * Same results as pdab_pda_group_tokens + the two position-MLP layers + DensityNet + pdab_pda_assemble_ln_split, with
 * all products in fp32 FMA.
 * replaces: PB/pointnet2_utils.py:567-614 (grouper), PB/pointnet2_modules.py:893-927 (encoding, position MLP, density
 *           re-weighting, token cat), :958-1006 (DensityNet), PB/PointFormer.py:29 (norm1). */
size_t pdab_pda_encode_param_floats(int c);
int pdab_pda_encode_ln(int b, int c, int n, int m, float radius, int nsample, const float *xyz, const float *new_xyz,
                       const float *features_t, const float *glob, const float *params, float eps, float *out,
                       pdab_stream_t stream);
/* Same, written as a (hi, lo) pair of fp16 planes (b*m*nsample, 4c) each: row = hi + lo to ~2^-22 — hi is the TMA operand
 * of pdab_tc_linear_h, hi + lo the residual stream. */
int pdab_pda_encode_ln_h(int b, int c, int n, int m, float radius, int nsample, const float *xyz, const float *new_xyz,
                         const float *features_t, const float *glob, const float *params, float eps, void *out_hi,
                         void *out_lo, pdab_stream_t stream);

/* Row-wise kernels of the PDA block, each fused with the hi/lo split (hi = top 19 bits, exactly representable in
 * TF32; lo = value - hi) that feeds error-compensated 3xTF32 tensor-core projections.  `tokens` rows of e = 4c
 * channels (e in {256, 512}); all tensors contiguous fp32.
 * replaces: torch.cat / mul / copy / nn.LayerNorm / residual add / ReLU / max-pool launches of the PDA block,
 *           PB/pointnet2_modules.py:893-931, PB/PointFormer.py:28-38. */
/* row = LayerNorm(cat[pos (c), feat*scale (c), feat (c), glob[row / nsample] (c)]); feat = x[row, 8 : 8+c] with row
 * pitch xpitch (the pdab_pda_group_tokens layout); pos (tokens,c), scale (tokens), glob (tokens/nsample, c).
 * lo == NULL: the un-split row is written to `hi` (the tcgen05 GEMMs of pdab_tc_linear split their operand themselves). */
int pdab_pda_assemble_ln_split(long long tokens, int nsample, int c, int xpitch, const float *pos, const float *x,
                               const float *scale, const float *glob, const float *gamma, const float *beta,
                               float eps, float *hi, float *lo, pdab_stream_t stream);
/* Self-attention inside each neighbourhood, all heads: ctx = softmax(q k^T / sqrt(head_dim)) v per (group, head).
 * replaces: the attention core of nn.MultiheadAttention in TransformerEncoderLayerPreNorm (q/k/v permutes, 2 bmm,
 *           softmax, permute back), PB/PointFormer.py:30, PB/pointnet2_modules.py:929.
 * qkv (groups*nsample, 3*heads*head_dim) = [q | k | v] rows as in_proj leaves them; the nsample rows of a group are
 * consecutive; ctx (groups*nsample, heads*head_dim).  nsample in {16, 32}, head_dim in {64, 128}.
 * npass: product class of the two contractions on mma.sync — 3 (or 1): error-compensated 3xTF32 (fp32-level); 2: split-bf16
 * m16n8k16 (hi / lo bf16 pairs, ~2^-17 per product, the class of the split-bf16 GEMMs; half the instructions). */
int pdab_group_attention(long long groups, int nsample, int heads, int head_dim, int npass, const float *qkv, float *ctx,
                         pdab_stream_t stream);

/* fp16 form (the fp16 single-pass mode of pdab_tc_linear_h): qkv (groups*nsample, 3E) and ctx (groups*nsample, E) are fp16;
 * both contractions are fp16 m16n8k16 MMAs with fp32 accumulation, the softmax is fp32.  replaces: the same lines. */
int pdab_group_attention_h(long long groups, int nsample, int heads, int head_dim, const void *qkv, void *ctx,
                           pdab_stream_t stream);

/* Fused plain set-abstraction scale: ball query -> group (xyz centred) -> shared MLP
 * (1x1 conv with eval-mode BatchNorm folded in, ReLU) x nlayers -> max over nsample.
 * The grouped tensor never reaches HBM.
 * replaces: QueryAndGroup.forward + mlps[i] + max_pool2d, PB/pointnet2_utils.py:681-704,
 *           PB/pointnet2_modules.py:1655-1672.
 * features (B,C,N) or NULL (C=0); dims[0] = 3+C, dims[l+1] = cout of layer l;
 * weights[l] device (dims[l+1], dims[l]) row-major, biases[l] device (dims[l+1]);
 * `weights` / `biases` / `dims` themselves are HOST arrays of length nlayers (+1).
 * out (B, dims[nlayers], M).  Shapes covered in this round: 3 layers, dims[0] <= 8,
 * (dims[1..3]) in {(16,16,32), (32,32,64)}, nsample <= 32; anything else returns PDAB_EUNSUPPORTED. */
int pdab_sa_fused(int b, int c, int n, int m, float radius, int nsample, const float *xyz, const float *new_xyz,
                  const float *features, int nlayers, const int *dims_host, const float *const *weights_host,
                  const float *const *biases_host, float *out, pdab_stream_t stream);

/* Two scales of one plain SA layer (same centres, same cloud) in ONE kernel: a single scan of the cloud serves both radii,
 * then both MLP + max-pool phases run; out (B, cout_a + cout_b, M) is the channel-concatenation of the two scales
 * (PB/pointnet2_modules.py:1655-1674).  weights_host / biases_host: HOST arrays of 6 device pointers, scale a's three
 * layers then scale b's.  Shapes covered: dims[0] <= 8, a = (16,16,32), b = (32,32,64), nsample <= 32; else
 * PDAB_EUNSUPPORTED (the caller runs pdab_sa_fused per scale). */
int pdab_sa_fused_pair(int b, int c, int n, int m, float radius_a, int nsample_a, float radius_b, int nsample_b,
                       const float *xyz, const float *new_xyz, const float *features, const int *dims_a_host,
                       const int *dims_b_host, const float *const *weights_host, const float *const *biases_host,
                       float *out, void *workspace, pdab_stream_t stream);
/* The same layer with the MLP contractions as fp16 x fp16 single-pass products (fp32 accumulation, bias / ReLU / max in fp32):
 * the product class of the `_h` tensor-core entry points and of the TF32 cuDNN convolutions the reference runs these layers
 * as (PB/pointnet2_modules.py:1655-1672); the neighbour lists are the same bit for bit, features differ by <= 1e-3 relative.
 * One mma.sync per product instead of three. */
int pdab_sa_fused_pair_h(int b, int c, int n, int m, float radius_a, int nsample_a, float radius_b, int nsample_b,
                         const float *xyz, const float *new_xyz, const float *features, const int *dims_a_host,
                         const int *dims_b_host, const float *const *weights_host, const float *const *biases_host,
                         float *out, void *workspace, pdab_stream_t stream);
/* workspace: NULL = every centre scans the whole cloud (M * N distance tests per scene); otherwise a device buffer of
 * pdab_sa_grid_workspace_bytes(b, n) bytes (16-byte aligned): the call first buckets every scene's points into a hashed
 * cell list (cell edge = the larger radius, + 0.1 %) and the centres then test only the 27 cells around them, keeping the
 * nsample smallest indices among the hits — the same neighbour lists, bit-identical outputs, ~N / 30 of the tests. */
size_t pdab_sa_grid_workspace_bytes(int b, int n);

/* Deterministic gradient of gather / group / three_interpolate: a sorted segmented sum instead of the
 * reference's float atomicAdd scatter (PB/src/group_points_gpu.cu:14-50, sampling_gpu.cu:46-83, interpolate_gpu.cu:120-160),
 * bit-identical from run to run.  The forward op read point idx[b, p] at slot p (p in [0, e): e = npoints, npoints *
 * nsample, or n * 3 for the interpolation); the caller passes `order` (B, e) int32 = the slots sorted by the point they
 * read (stable) and `seg_start` (B, n + 1) int32 = where each point's run begins in that order.  Then
 *   grad_points[b, c, i] = sum_{k in [seg_start[b,i], seg_start[b,i+1])} grad_out[b, c, order[b,k] / div] * weight[b, order[b,k]]
 * added up in order (weight == NULL: 1; div = 3 for three_interpolate, whose grad_out is (B, c, e / 3), else 1).
 * grad_points (B, c, n) is fully written (points nobody read get 0). */
int pdab_segment_sum_grad(int b, int c, int n, int e, int div, const float *grad_out, const float *weight, const int *order,
                          const int *seg_start, float *grad_points, pdab_stream_t stream);

/* ---- tensor-core (tcgen05 / TMEM) contractions ---------------------------------- */

/* Epilogues of pdab_tc_linear. */
#define PDAB_EPI_STORE 0        /* out = acc + bias */
#define PDAB_EPI_RELU 1         /* out = relu(acc + bias) */
#define PDAB_EPI_ADD_LN 2       /* out = LayerNorm(acc + bias + residual) over the full row (nout in {256, 512}) */
#define PDAB_EPI_ADD_MAXPOOL 3  /* out[g] = max over the nsample rows of group g of (acc + bias + residual) */
#define PDAB_EPI_RELU_MAXPOOL 4 /* out[g] = max over the nsample rows of group g of relu(acc + bias) */
#define PDAB_EPI_ATTN 5         /* in_proj + neighbourhood self-attention (head_dim 64): W rows grouped head by head,
                                 * [q_h | k_h | v_h] = one bn = 192 chunk per head (nout = 3E); out (rows, E) =
                                 * softmax(q k^T / 8) v inside each group of nsample consecutive rows, heads
                                 * concatenated — pdab_tc_linear(STORE) + pdab_group_attention in one kernel, the
                                 * (rows, 3E) qkv matrix is never stored.  replaces: nn.MultiheadAttention in_proj +
                                 * attention core, PB/PointFormer.py:30 */

/* out = epilogue(a (rows,k) . W (nout,k)^T + bias) on the 5th-generation tensor cores (tcgen05.mma kind::tf32,
 * fp32 accumulation in TMEM).  npass = 1: operands rounded to TF32 (the precision class of the reference's cuDNN
 * 1x1 convolutions); npass = 3: error-compensated 3xTF32 (hi/lo operand split, fp32-level results, what the
 * reference's nn.Linear / nn.MultiheadAttention projections compute).
 * replaces: the Conv2d(1x1)+BN+ReLU layers and max_pool2d of a plain SA scale, PB/pointnet2_modules.py:1478-1491,
 *           1657-1672 (BN folded into W / bias by the caller), and the in_proj / out_proj / linear1 / linear2 GEMMs
 *           of TransformerEncoderLayerPreNorm with the LayerNorm / residual / ReLU / max-pool between them,
 *           PB/PointFormer.py:28-38, PB/pointnet2_modules.py:929-931.
 * a (rows, lda) fp32; w_packed: W pre-packed by pdab_tc_pack_weights (bn = 128 or 256 columns per accumulator
 * chunk; 192 for PDAB_EPI_ATTN); bias (nout) or NULL; residual (rows, ldr) for the ADD_* epilogues; gamma/beta/eps for ADD_LN;
 * nsample in {16, 32} for the *_MAXPOOL epilogues (rows % nsample == 0; out has rows / nsample rows).
 * k, lda, ldo, ldr, nout multiples of 4; all pointers 16-byte aligned. */
int pdab_tc_linear(long long rows, int k, int nout, int npass, int bn, int epilogue, const float *a, int lda,
                   const float *w_packed, const float *bias, const float *residual, int ldr, const float *gamma,
                   const float *beta, float eps, int nsample, float *out, int ldo, pdab_stream_t stream);

/* Launch policy of the CALLING THREAD (thread-local; results never change, only how launches are shaped).
 * Largest thread-block cluster (CTAs per scene) the n > 16384 FPS path may use beyond the minimum that holds the scene
 * (default 16: shortest chain).  A caller that pipelines batches lowers it so one batch's FPS chain occupies few SMs and runs
 * beside the other batches' kernels.  1 <= n <= 16. */
int pdab_set_fps_max_cluster(int n);

/* Grid size of the persistent tensor-core kernels launched by the calling thread: 0 (default) = one CTA per SM of the
 * current device (cudaDevAttrMultiProcessorCount).  A caller that pipelines batches on several streams lowers it (e.g.
 * SMs - scenes per batch) so the persistent grid never queues behind the one-CTA-per-scene FPS kernels of another batch. */
int pdab_set_persistent_ctas(int n);

/* 1 (default): the calling thread's tensor-core kernels run as cta_group::2 CTA pairs (thread-block clusters of 2, M = 256)
 * when the problem is large enough; 0: single CTAs only.  Same results either way (a schedule, not arithmetic). */
int pdab_set_cta_pairs(int on);

/* fp16 single-pass form of pdab_tc_linear (tcgen05.mma kind::f16, fp16 x fp16 products, fp32 accumulation in TMEM; 11-bit
 * significands = the TF32 precision class at a third of the MMAs of the split modes), with 16-bit activations BETWEEN
 * kernels:
 *   a        a_fp16 = 1: fp16 (rows, lda) row-major, 16-byte aligned, lda % 8 == 0 — loaded by the TMA engine
 *            (cp.async.bulk.tensor through a SWIZZLE_128B tensor map: no producer warps, no register round trip);
 *            a_fp16 = 0: fp32 (rows, lda), converted to fp16 by the kernel's producer warps.
 *   w_packed pdab_tc_pack_weights(npass = 4): one fp16 plane.
 *   residual res_lo == NULL: fp32 (rows, ldr) in res_hi; otherwise a (hi, lo) pair of fp16 planes, residual = hi + lo
 *            (22 significand bits: the residual stream of the transformer keeps fp32-level precision — it, not the
 *            products, sets the output error of the block).
 *   out      out_fmt 0: fp32 (.., ldo); 1: fp16; 2 (ADD_LN only): (hi, lo) fp16 planes in out / out_lo — the next
 *            GEMM's TMA operand is the hi plane, the next residual is hi + lo.  *_MAXPOOL outputs are fp32.
 *   nsample  additionally 64 for PDAB_EPI_RELU_MAXPOOL (out is zeroed first: a memset node on `stream`).
 *   k % 8 == 0.  Everything else as pdab_tc_linear.
 * replaces: the same reference lines as pdab_tc_linear. */
int pdab_tc_linear_h(long long rows, int k, int nout, int bn, int epilogue, const void *a, int lda, int a_fp16,
                     const float *w_packed, const float *bias, const void *res_hi, const void *res_lo, int ldr,
                     const float *gamma, const float *beta, float eps, int nsample, void *out, void *out_lo, int ldo,
                     int out_fmt, pdab_stream_t stream);

/* Fused second half of the PDA transformer block, d_model e = 256, ONE launch (csrc/tc_ffn.cu):
 *   z = LayerNorm(y + ctx . Wo^T + bo);  h = relu(z . W1^T + b1);  out[g] = max over the nsample rows of group g of (z + h . W2^T + b2)
 * = pdab_tc_linear_h(ADD_LN) -> pdab_tc_linear_h(RELU) -> pdab_tc_linear_h(ADD_MAXPOOL) with z (rows, e) and h (rows, e/2)
 * kept in shared memory / TMEM as the next MMA's operand: 6 e bytes of HBM traffic per row instead of 18 e.
 * ctx fp16 (rows, ldc); y_hi / y_lo fp16 planes (rows, ldy), y = hi + lo; wo / w1 / w2 packed with
 * pdab_tc_pack_weights(npass = 4; bn = 256, 128, 256); out fp32 (rows / nsample, ldo); nsample in {16, 32},
 * rows % nsample == 0.  e != 256 returns PDAB_EUNSUPPORTED (the caller runs the three-launch chain).
 * replaces: out_proj + residual + norm2 + linear1 + ReLU + linear2 + residual of TransformerEncoderLayerPreNorm and the
 *           max over the neighbourhood, PB/PointFormer.py:31-37, PB/pointnet2_modules.py:929-931. */
int pdab_tc_ffn_h(long long rows, int e, int nsample, const void *ctx, int ldc, const void *y_hi, const void *y_lo, int ldy,
                  const float *wo_packed, const float *bo, const float *gamma, const float *beta, float eps,
                  const float *w1_packed, const float *b1, const float *w2_packed, const float *b2, float *out, int ldo,
                  pdab_stream_t stream);

/* Number of floats pdab_tc_pack_weights writes for a (nout, k) weight matrix. */
size_t pdab_tc_packed_floats(int nout, int k, int npass, int bn);
/* Packs W (nout, k) row-major (device) into the shared-memory image the tensor-core kernels stream:
 * [column chunk of bn][k-atom of 32][hi | lo][bn rows x 128 B, K-major, 128-byte swizzle], zero padded.
 * npass = 3 writes the (hi, lo) TF32 split, npass = 1 the round-to-nearest TF32 value, npass = 2 the (hi, lo) bf16 split
 * (k-atoms of 64), npass = 4 one fp16 plane (k-atoms of 64; pdab_tc_linear_h).  If xyz_last > 0 the first
 * xyz_last input columns of W are moved behind the others (the gather prologue of pdab_tc_sa_gather_linear feeds
 * [features, centred xyz] while the reference's Conv2d expects [xyz, features], PB/pointnet2_utils.py:692-699). */
int pdab_tc_pack_weights(int nout, int k, int npass, int bn, int xyz_last, const float *w, float *packed,
                         pdab_stream_t stream);

/* First layer of a plain SA scale with the grouping fused into the GEMM prologue:
 * out[(b,j,s), :] = relu(W . [features_t[b, idx[b,j,s], :], xyz[b, idx[b,j,s]] - new_xyz[b,j]] + bias).
 * The grouped (B, 3+C, M, nsample) tensor of the reference never exists.
 * replaces: grouping_operation x2 + centre subtraction + cat + first Conv2d/BN/ReLU,
 *           PB/pointnet2_utils.py:689-704, PB/pointnet2_modules.py:1657-1658.
 * idx (B,M,nsample) from pdab_ball_query; features_t (B,N,C) POINT-major; w_packed packed with bn = 256 and
 * xyz_last = 3; out (B*M*nsample, ldo). */
int pdab_tc_sa_gather_linear(int b, int c, int n, int m, int nsample, int nout, int npass, const int *idx,
                             const float *features_t, const float *xyz, const float *new_xyz, const float *w_packed,
                             const float *bias, float *out, int ldo, pdab_stream_t stream);

/* fp16 single-pass form of pdab_tc_sa_gather_linear (w_packed: npass = 4); out fp32 or, out_fp16 != 0, fp16 (the TMA
 * operand of the next pdab_tc_linear_h layer).  replaces: the same reference lines. */
int pdab_tc_sa_gather_linear_h(int b, int c, int n, int m, int nsample, int nout, const int *idx,
                               const float *features_t, const float *xyz, const float *new_xyz, const float *w_packed,
                               const float *bias, void *out, int ldo, int out_fp16, pdab_stream_t stream);

/* ---- iou3d_nms_cuda -------------------------------------------------------- */

/* Bytes of device workspace pdab_nms_device needs for n boxes. */
size_t pdab_nms_workspace_bytes(int n);

/* Rotated-BEV NMS, entirely on the device: boxes (n,7) [x,y,z,dx,dy,dz,heading]
 * already sorted by descending score; keep (n) int64 receives positions into that
 * order, num_keep (1) int32 the count.  Same bitmask + greedy result as the reference.
 * replaces: nms_gpu, IOU/src/iou3d_nms_api.cpp:14, IOU/src/iou3d_nms.cpp:90-136,
 *           kernel IOU/src/iou3d_nms_kernel.cu:267-311. */
int pdab_nms_device(const float *boxes, int n, float thresh, int64_t *keep, int *num_keep, void *workspace,
                    pdab_stream_t stream);

/* Batched form: nscenes box lists stored back to back with row stride `stride`
 * boxes each; counts[s] boxes valid in scene s (device int32); outputs
 * keep (nscenes,stride) and num_keep (nscenes).  workspace:
 * nscenes * pdab_nms_workspace_bytes(stride) bytes.  One launch for the batch. */
int pdab_nms_batched(const float *boxes, const int *counts, int nscenes, int stride, float thresh, int64_t *keep,
                     int *num_keep, void *workspace, pdab_stream_t stream);

/* Drop-in for the reference pybind signature: device boxes, HOST keep list,
 * returns num_to_keep (>= 0) or a negative error.  Device scratch and a pinned
 * staging buffer are cached per host thread (grow-only; freed at thread exit);
 * the keep list returns in ONE device-to-host copy followed by one
 * synchronisation of `stream` (the reference does cudaMalloc + blocking
 * cudaMemcpy + cudaFree + a host loop here).
 * `normal` != 0 selects the axis-aligned variant (nms_normal_gpu,
 * IOU/src/iou3d_nms.cpp:139-185, kernel IOU/src/iou3d_nms_kernel.cu:328-372). */
int pdab_nms_host(const float *boxes, int n, float thresh, int64_t *keep_host, int normal, pdab_stream_t stream);

/* Batched post-processing around the NMS, no host synchronisation, fixed shapes.
 * pdab_post_front: per scene (b scenes of m centres each): p_c = sigmoid(cls[., c]) (or cls itself if `normalized`), score =
 * max_c p_c (first maximum), label = argmax + 1, valid = score >= score_thresh; the centres are ordered by (valid ? score :
 * -inf) descending, ties by ascending index (`sort(descending, stable)`).  Writes order (b, m) int32, sorted_boxes (b, m, 7)
 * = boxes gathered into that order (the NMS input), counts (b) = min(#valid, pre_max), scores / raw_max (b, m) (raw_max = the
 * maximum LOGIT, for OUTPUT_RAW_SCORE) and labels (b, m) int64.  cls (b*m, ldc) with nc classes, boxes (b*m, ldb) with ldb >= 7;
 * m <= 4096.
 * pdab_post_select: after pdab_nms_batched: out row j < min(num_keep, p) = centre order[keep[j]] of the scene (boxes with nb
 * columns, score, label); rows behind it are zero; out_num (b) = min(num_keep, p).
 * replaces: the per-scene loop of Detector3DTemplate.post_processing (pcdet/models/detectors/detector3d_template.py:196-285)
 *           and class_agnostic_nms (pcdet/models/model_utils/model_nms_utils.py:6-25) around nms_gpu. */
int pdab_post_front(int b, int m, int nc, int ldc, int ldb, int normalized, float score_thresh, int pre_max, const float *cls,
                    const float *boxes, int *order, float *sorted_boxes, int *counts, float *scores, float *raw_max,
                    int64_t *labels, pdab_stream_t stream);
int pdab_post_select(int b, int m, int p, int nb, int ldb, const int64_t *keep, const int *num_keep, const int *order,
                     const float *boxes, const float *scores, const int64_t *labels, float *out_boxes, float *out_scores,
                     int64_t *out_labels, int *out_num, pdab_stream_t stream);

/* Pairwise rotated BEV overlap area / IoU: out (na, nb).
 * replaces: boxes_overlap_bev_gpu / boxes_iou_bev_gpu, IOU/src/iou3d_nms_api.cpp:12-13,
 *           IOU/src/iou3d_nms.cpp:46-88, kernels IOU/src/iou3d_nms_kernel.cu:236-265. */
int pdab_boxes_overlap_bev(int na, const float *boxes_a, int nb, const float *boxes_b, float *out,
                           pdab_stream_t stream);
int pdab_boxes_iou_bev(int na, const float *boxes_a, int nb, const float *boxes_b, float *out, pdab_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* PDAB_H_ */
