/*
 * butd_b200.h — C ABI of libbutd_b200.so: the B200 (sm_100a) forward hot path of BUTD-DETR.
 *
 * This is the drop-in boundary.  Part A replaces, one for one, the nine functions of the
 * reference's native plugin `pointnet2._ext` (pybind11 module,
 * /root/reference/pointnet2/_ext_src/src/bindings.cpp:11-24); Part B are the fused forward
 * entry points the B200 model engine is built from (they replace the ATen / cuDNN / cuBLAS
 * call sequences of models/bdetr.py:193-319).  INTEGRATION.md shows the reference-side
 * binding (a ctypes `pointnet2/_ext.py`).
 *
 * Conventions (all functions):
 *   - plain pointers and sizes only; every pointer is DEVICE memory owned by the caller
 *     (the reference's ops allocate their outputs with torch::zeros; here the host shim does);
 *   - asynchronous: work is enqueued on `stream` (a cudaStream_t passed as void*; NULL = the
 *     legacy default stream); no host synchronisation, no allocation, graph-capturable;
 *   - returns BD_OK or an error code; never exits or throws (the reference's
 *     CUDA_CHECK_ERRORS calls exit(-1), cuda_utils.h:35-44).  bd_last_error() returns a
 *     thread-local message for the last failing call of the calling thread;
 *   - re-entrant and thread-safe; the device is the one current on the calling thread;
 *   - float = IEEE fp32, indices = int32, layouts row-major contiguous unless a leading
 *     dimension (`ld*`, in elements) is given.
 */
#ifndef BUTD_B200_H_
#define BUTD_B200_H_

#ifdef __cplusplus
extern "C" {
#endif

typedef void *bd_stream_t;

enum {
  BD_OK = 0,
  BD_ERR_INVALID_ARG = 1, /* null pointer, non-positive size, unsupported shape */
  BD_ERR_CUDA = 2,        /* launch / runtime error; see bd_last_error() */
  BD_ERR_UNSUPPORTED = 3
};

int bd_version(void);
const char *bd_last_error(void);
/* Compiled-for architecture string, e.g. "sm_100a". */
const char *bd_arch(void);

/* ===================================================================================== */
/* Part A — the nine ops of pointnet2._ext (argument meaning as in the reference headers) */
/* ===================================================================================== */

/* furthest_point_sampling(points (B,N,3), nsamples) -> idx (B,m) int32
 * replaces _ext_src/src/sampling.cpp:70-91 + sampling_gpu.cu:74-234.
 * `ld` = floats per point row (3 for the reference layout; 6 to sample straight from the
 * (B,N,3+C) point cloud).  `tmp` = scratch of B*N floats, only touched when N exceeds the
 * register-resident capacity reported by bd_fps_resident_capacity(); may be NULL below it.
 * Tie-breaking is bit-identical to the reference kernel launched with
 * opt_n_threads(N) threads (cuda_utils.h:20-24). */
int bd_fps(const float *xyz, int ld, int B, int N, int m, float *tmp, int *idx, bd_stream_t stream);
int bd_fps_resident_capacity(void);
/* Throughput variant of bd_fps over the cell list of the same clouds (bd_grid_build(xyz, ld, B, N,
 * radius, grid_workspace); any radius, <= 0 picks a cell size): one CTA per scene walks the
 * cell-ordered points in buckets of 32 and skips every bucket whose bounding box is farther from
 * the new sample than the bucket's largest running distance (~2 % of the cloud is touched per
 * round), so a batch of up to 148 scenes runs in one wave; beyond one scene per SM an 8-warp variant puts two
 * scenes on each SM (296 scenes: 6.7 ms against two waves of 4.1 ms).  Same indices as bd_fps, bit for bit.
 * N <= bd_fps_grid_capacity(); `scratch`: bd_fps_grid_scratch_bytes(B, N) bytes. */
int bd_fps_grid(const float *xyz, int ld, int B, int N, int m, void *grid_workspace, float *scratch,
                int *idx, bd_stream_t stream);
int bd_fps_grid_capacity(void);
long long bd_fps_grid_scratch_bytes(int B, int N);
/* Tuning / test hooks: force the cluster size (4, 8 or 16 CTAs) bd_fps uses for clouds that need
 * more than one CTA (-1 restores the automatic choice: 16 for B <= 4 scenes, 8 for B <= 8, else 4);
 * warps per CTA of the bd_fps_grid kernel (8 = two CTAs per SM, 16 or 32; 0 restores the choice by the number
 * of scenes). */
int bd_fps_set_cluster(int cluster);
int bd_fps_grid_set_warps(int warps);
/* Tools: switch the counting variant of the bd_fps_grid kernel on / off and read (and clear) its
 * counters: out16 = {bucket visits, visit batches, warp-rounds with a visit, warp-rounds, -, -, -, -, 7 per-phase cycle sums of warp 0 of CTA 0, -}. */
int bd_fps_grid_stats(int enable, unsigned long long *out16);

/* gather_points(points (B,C,N), idx (B,m)) -> out (B,C,m)      sampling.cpp:20-43 */
int bd_gather_points(const float *points, const int *idx, int B, int C, int N, int m, float *out,
                     bd_stream_t stream);
/* gather_points_grad(grad_out (B,C,m), idx (B,m), N) -> grad_points (B,C,N), zero-filled here
 * sampling.cpp:45-69 */
int bd_gather_points_grad(const float *grad_out, const int *idx, int B, int C, int N, int m,
                          float *grad_points, bd_stream_t stream);

/* ball_query(new_xyz (B,m,3), xyz (B,n,3), radius, nsample) -> idx (B,m,nsample)
 * ball_query.cpp:13-37 + ball_query_gpu.cu:14-59: first `nsample` in-ball points in ascending
 * index order, remaining slots = first hit, all zeros when the ball is empty.
 * `ld_xyz` = floats per row of `xyz` (3 in the reference layout). */
int bd_ball_query(const float *new_xyz, const float *xyz, int ld_xyz, int B, int n, int m,
                  float radius, int nsample, int *idx, bd_stream_t stream);

/* Same result as bd_ball_query, bit for bit, through a per-scene uniform grid (cell list): points
 * are binned into cells of edge >= radius, each centre visits its 27 neighbour cells and keeps the
 * `nsample` smallest in-ball indices (sorted register file + index threshold, no hit cap), which
 * restores the reference's index order.  Replaces the 102 M brute-force distance tests per scene
 * of ball_query_gpu.cu:14-49 at SA1 by ~0.5 M.  nsample > 64 falls through to bd_ball_query.
 * `workspace`: bd_ball_query_grid_workspace_bytes(B, n) bytes of device scratch. */
long long bd_ball_query_grid_workspace_bytes(int B, int n);
int bd_ball_query_grid(const float *new_xyz, const float *xyz, int ld_xyz, int B, int n, int m,
                       float radius, int nsample, int *idx, void *workspace, bd_stream_t stream);
/* The two halves of bd_ball_query_grid: bd_grid_build bins the clouds once (cells of edge >=
 * radius; radius <= 0: 1/32 of the largest extent); bd_ball_query_grid_query answers queries
 * against that cell list (same xyz; cells >= radius are the efficient case, smaller cells are
 * visited further out); bd_grid_order returns the device array (B,n) of point indices grouped by
 * cell.  The same cell list drives bd_fps_grid. */
int bd_grid_build(const float *xyz, int ld_xyz, int B, int n, float radius, void *workspace,
                  bd_stream_t stream);
const int *bd_grid_order(void *workspace, int B, int n);
int bd_ball_query_grid_query(const float *new_xyz, const float *xyz, int ld_xyz, int B, int n, int m,
                             float radius, int nsample, int *idx, void *workspace,
                             bd_stream_t stream);

/* group_points(points (B,C,n), idx (B,m,ns)) -> out (B,C,m,ns)   group_points.cpp:17-40 */
int bd_group_points(const float *points, const int *idx, int B, int C, int n, int m, int ns,
                    float *out, bd_stream_t stream);
/* group_points_grad(grad_out (B,C,m,ns), idx, n) -> grad_points (B,C,n)  group_points.cpp:42-65 */
int bd_group_points_grad(const float *grad_out, const int *idx, int B, int C, int n, int m, int ns,
                         float *grad_points, bd_stream_t stream);

/* three_nn(unknown (B,n,3), known (B,m,3)) -> dist2 (B,n,3) SQUARED, idx (B,n,3)
 * interpolate.cpp:19-45 + interpolate_gpu.cu:14-73 */
int bd_three_nn(const float *unknown, const float *known, int B, int n, int m, float *dist2,
                int *idx, bd_stream_t stream);
/* three_interpolate(points (B,C,m), idx (B,n,3), weight (B,n,3)) -> out (B,C,n)
 * interpolate.cpp:47-75 */
int bd_three_interpolate(const float *points, const int *idx, const float *weight, int B, int C,
                         int m, int n, float *out, bd_stream_t stream);
/* three_interpolate_grad(grad_out (B,C,n), idx, weight, m) -> grad_points (B,C,m)
 * interpolate.cpp:76-104 */
int bd_three_interpolate_grad(const float *grad_out, const int *idx, const float *weight, int B,
                              int C, int n, int m, float *grad_points, bd_stream_t stream);

/* ===================================================================================== */
/* Part B — fused forward entry points (token-major activations: one row per point/token) */
/* ===================================================================================== */

/* out[b,j,0:w] = src[b, idx[b,j], 0:w]   (rows of `ld_src` floats; out rows of `ld_out`).
 * Token-major gather: new_xyz from FPS indices, query features/xyz from top-k indices
 * (pointnet2_modules.py:238, models/modules.py:80-84). */
int bd_gather_rows(const float *src, int ld_src, const int *idx, int B, int n_src, int m, int w,
                   float *out, int ld_out, bd_stream_t stream);

/* QueryAndGroup.forward (pointnet2_utils.py:334-359, use_xyz, normalize_xyz) on an existing
 * ball-query result, token-major:  out[(b*m+j)*ns+s, :] =
 *   [ (xyz[b,idx]-new_xyz[b,j]) * (1/radius)  (3) | feats[b,idx,0:C] (C) | 0 ... ]
 * rows of ld_out >= 3+C floats, zero padded (lets the next GEMM use a K that is a multiple of 4) */
int bd_group_rows(const float *xyz, int ld_xyz, const float *feats, int ld_feats, int C,
                  const float *new_xyz, const int *idx, int B, int n, int m, int ns, float radius,
                  float *out, int ld_out, bd_stream_t stream);

/* F.max_pool2d over nsample (pointnet2_modules.py:255): out[r, c] = max_s in[r*ns+s, c] */
int bd_maxpool_rows(const float *in, int rows_out, int ns, int C, float *out, bd_stream_t stream);

/* PointnetFPModule.forward up to the concat (pointnet2_modules.py:393-408), token-major:
 * w = 1/(sqrt(dist2)+1e-8) normalised over the 3 neighbours;
 * out[b,j,:] = [ sum_t w_t * known_feats[b,idx_t,0:C2] | unknown_feats[b,j,0:C1] ] */
int bd_fp_interp_concat(const float *dist2, const int *idx, const float *known_feats, int C2,
                        const float *unknown_feats, int C1, int B, int n, int m, float *out,
                        bd_stream_t stream);
/* Same with the row format chosen by the caller: out_half != 0 writes fp16 rows (B*n, C2+C1) — the operand of the
 * linear layer that follows (C1 % 4 == 0, C2 % 4 == 0, 16-byte aligned tensors); out_half = 0 is
 * bd_fp_interp_concat. */
int bd_fp_interp_concat_h(const float *dist2, const int *idx, const float *known_feats, int C2,
                          const float *unknown_feats, int C1, int B, int n, int m, void *out,
                          int out_half, bd_stream_t stream);

/* Y = act( (A [+ A2]) · Wᵀ + bias ) : A (M,K) lda, A2 optional same shape (lda2), W (N,K)
 * row-major (torch Linear / 1x1-conv weight, BatchNorm folded by the host), bias (N) or NULL,
 * Y (M,N) ldy.  relu = 1 applies max(.,0), relu = 2 the exact (erf) GELU of RoBERTa's feed-forward block (the
 * same codes hold for bd_linear_tc / bd_linear_tc_h).  fp32 SIMT path (1e-3 parity gate). */
int bd_linear_f32(const float *A, int lda, const float *A2, int lda2, const float *W,
                  const float *bias, float *Y, int ldy, int M, int N, int K, int relu,
                  bd_stream_t stream);

/* Same contract as bd_linear_f32, computed on the 5th-gen tensor cores (tcgen05.mma, 16-bit
 * operands, fp32 accumulation in TMEM).  A (and A2) arrive through the TMA unit as raw fp32
 * chunks and are converted while staged; `Wp` is the weight pre-packed by the host
 * (butd_detr_b200.engine.pack_weight_tc) into the kernel's shared-memory layout (128-byte swizzle,
 * K-major):  Wp[n_group][k_chunk][part][sub][BN rows][64 k] 16-bit, zero padded, where one CTA
 * computes n_sub consecutive BN-wide column tiles (n_group = n / (n_sub*BN), sub = (n / BN) %
 * n_sub) and k_chunk = k / 64.  KC = 64; BN: multiple of 16, <= 256; n_sub * BN <= 512 (TMEM
 * columns).
 * split = 1: FP16 operands (part = {fp16(W)}), one MMA per product — outputs of the whole forward
 * within the 1e-2 gate.  split = 3 ("bf16x3"): every fp32 operand is carried as bf16 hi + bf16 lo
 * and D += Ahi*Whi + Alo*Whi + Ahi*Wlo (part = {hi, lo}), which restores fp32-grade products on
 * the bf16 tensor pipe (1e-3 gate).  Two pipeline stages must fit 217 KB of shared memory. */
int bd_linear_tc(const float *A, int lda, const float *A2, int lda2, const void *Wp,
                 const float *bias, float *Y, int ldy, int M, int N, int K, int KC, int n_chunks,
                 int BN, int n_sub, int relu, int split, bd_stream_t stream);

/* First SharedMLP layer of a set-abstraction level with QueryAndGroup FUSED into the operand
 * staging (pointnet2_utils.py:334-359 + pytorch_utils.py:25-36): row (b, j, s) of the implicit A is
 * [ feats[b, idx[b,j,s], 0:C] | (xyz[b, idx] - new_xyz[b, j]) / radius | 0 ], K = round_up(C + 3, 8);
 * Y (B*m*ns, N) = relu(A Wᵀ + bias).  The packed weight's K columns must be in that order
 * (features first).  The grouped tensor never exists in HBM. */
int bd_sa_group_linear_tc(const int *idx, const float *feats, int ld_feats, int C, const float *xyz,
                          int ld_xyz, const float *new_xyz, int B, int n, int m, int ns,
                          float radius, const void *Wp, const float *bias, float *Y, int ldy, int N,
                          int KC, int n_chunks, int BN, int n_sub, int split, bd_stream_t stream);

/* Last SharedMLP layer with the max-pool over nsample fused into the epilogue
 * (pointnet2_modules.py:251-257): Y (M / pool, N) = max over groups of `pool` consecutive rows of
 * relu(A Wᵀ + bias); pool divides 128 and M. */
int bd_linear_pool_tc(const float *A, int lda, const void *Wp, const float *bias, float *Y, int ldy,
                      int M, int N, int K, int KC, int n_chunks, int BN, int n_sub, int pool,
                      int split, bd_stream_t stream);

/* One whole set-abstraction level in one kernel: QueryAndGroup (as bd_sa_group_linear_tc) ->
 * SharedMLP of three [1x1 conv + folded BN + ReLU] layers -> max-pool over nsample
 * (pointnet2_modules.py:243-257).  The hidden activations stay in shared / tensor memory.
 * Wp0/1/2: packed weights (pack_weight_tc with full_rows tiling; layer 0's K columns ordered
 * [features | xyz | 0], K = round_up(C + 3, 8)); N0, N1 in {64,128}, N2 in {64,128,256}; nsample
 * divides 128.  Y (B*m, N2) pooled rows. */
int bd_sa_mlp_tc(const int *idx, const float *feats, int ld_feats, int C, const float *xyz, int ld_xyz,
                 const float *new_xyz, int B, int n, int m, int ns, float radius, const void *Wp0,
                 const float *b0, int N0, const void *Wp1, const float *b1, int N1, const void *Wp2,
                 const float *b2, int N2, float *Y, int ldy, int split, bd_stream_t stream);
/* bd_sa_mlp_tc with 16-bit feature rows between the levels (fp16 mode): feats_half != 0 -> `feats` are fp16 rows
 * (ld_feats in halfs; C % 8 == 0, ld_feats % 8 == 0, 16-byte aligned) that are copied into the operand tiles as
 * they are; Y16 != NULL -> the pooled rows are ALSO written as fp16 (ldy16 halfs per row), the gather source of the
 * next level.  Same values as bd_sa_mlp_tc in the fp16 mode (it rounds the gathered features to fp16 itself). */
int bd_sa_mlp_tc_h(const int *idx, const void *feats, int ld_feats, int C, int feats_half, const float *xyz,
                   int ld_xyz, const float *new_xyz, int B, int n, int m, int ns, float radius,
                   const void *Wp0, const float *b0, int N0, const void *Wp1, const float *b1, int N1,
                   const void *Wp2, const float *b2, int N2, float *Y, int ldy, void *Y16, int ldy16,
                   int split, bd_stream_t stream);

/* 16-bit activation variants of bd_linear_tc / bd_linear_ln_tc (fp16 operand mode; an fp32 A2 may be
 * added to an fp32 A as in bd_linear_tc): a_half = A
 * is an fp16 (M, K) matrix (lda in halfs, % 8) that arrives by tensor copy directly in the swizzled
 * operand layout — no conversion pass, half the bytes; y_half = Y is written as fp16 rows (ldy in
 * halfs, % 8).  bd_linear_ln_tc_h always writes its fp32 rows (the residual stream) and, when Y16 is
 * not NULL, an fp16 copy for the projections that consume the rows next.  Values equal the fp32
 * entry points' followed by the fp16 rounding their consumers apply anyway. */
int bd_linear_tc_h(const void *A, int lda, int a_half, const float *A2, int lda2, const void *Wp,
                   const float *bias, void *Y, int ldy, int y_half, int M, int N, int K, int kc,
                   int n_chunks, int BN, int n_sub, int relu, bd_stream_t stream);
int bd_linear_ln_tc_h(const void *A, int lda, int a_half, const void *Wp, const float *bias,
                      const float *R, int ldr, const float *gamma, const float *beta, float eps,
                      float *Y, int ldy, void *Y16, int ldy16, int M, int N, int K, int kc,
                      int n_chunks, int BN, int n_sub, bd_stream_t stream);
/* Occupancy policy of bd_linear_tc / bd_linear_pool_tc: 1 (default) = two CTAs per SM where the
 * accumulators fit 256 TMEM columns and the grid exceeds one wave; 0 = one CTA per SM, deepest ring. */
int bd_linear_tc_set_occupancy(int two_per_sm);

/* Tuning aid: device buffer (>= 64 x int64) that receives clock64() stamps of the phases of CTA
 * (0,0) of every following bd_linear*_tc launch; NULL disables. */
int bd_linear_tc_set_debug(long long *buf);

/* Fused  Y = LayerNorm(R + (A [+ A2]) · Wᵀ + bias) * gamma + beta  on the same kernel: the CTA
 * owns complete rows (n_sub * BN >= N, N <= 320), so the residual add and the row-wise
 * LayerNorm run in the epilogue (post-LN blocks of encoder_decoder_layers.py: attention
 * out-projection + norm, FFN second layer + norm). */
int bd_linear_ln_tc(const float *A, int lda, const float *A2, int lda2, const void *Wp,
                    const float *bias, const float *R, int ldr, const float *gamma,
                    const float *beta, float eps, float *Y, int ldy, int M, int N, int K, int KC,
                    int n_chunks, int BN, int n_sub, int split, bd_stream_t stream);

/* Y[r,:] = LayerNorm(X[r,:] + R[r,:]) * gamma + beta   (R may be NULL), rows of D floats,
 * biased variance, eps inside the sqrt (torch.nn.LayerNorm). */
int bd_add_layernorm_f32(const float *X, const float *R, const float *gamma, const float *beta,
                         float *Y, int M, int D, float eps, bd_stream_t stream);

/* Multi-head attention core of nn.MultiheadAttention (eval):
 * O[b,i,h*hd:(h+1)*hd] = softmax_j( scale * <Q[b,i,h,:], K[b,j,h,:]> + mask ) · V[b,j,h,:]
 * Q rows: Q + b*sq_b + i*ldq (+ h*hd), likewise K, V, O.  key_padding_mask (B,Lk) bytes,
 * nonzero = ignore (-inf), may be NULL.  A fully masked row yields NaN like the reference. */
int bd_attention_f32(const float *Q, int ldq, long long sq_b, const float *K, int ldk,
                     long long sk_b, const float *V, int ldv, long long sv_b,
                     const unsigned char *key_padding_mask, float *O, int ldo, long long so_b,
                     int B, int H, int Lq, int Lk, int hd, float scale, bd_stream_t stream);

/* Same contract as bd_attention_f32 on the tensor cores (tcgen05.mma, accumulators in TMEM, exact
 * online softmax in fp32).  head_dim 36 only; Q / K rows 16-byte aligned.  split = 1: fp16
 * operands; split = 3: bf16 hi/lo split operands for Q·Kᵀ and P·V (fp32-grade).  `workspace`:
 * bd_attention_tc_workspace_bytes(...) bytes of device scratch (16-byte aligned) that receives the
 * packed 16-bit K / Vᵀ operand tiles (a pack kernel runs first, then the tensor-core kernel,
 * which converts its own Q rows). */
long long bd_attention_tc_workspace_bytes(int B, int H, int Lq, int Lk, int split);
int bd_attention_tc(const float *Q, int ldq, long long sq_b, const float *K, int ldk,
                    long long sk_b, const float *V, int ldv, long long sv_b,
                    const unsigned char *key_padding_mask, float *O, int ldo, long long so_b,
                    int B, int H, int Lq, int Lk, int hd, float scale, int split, void *workspace,
                    bd_stream_t stream);
/* Programmatic dependent launch of the tensor-core kernels (each starts its prologue and weight
 * copies under the tail of its predecessor in the stream and waits for it before touching
 * activations): 1 = on (default), 0 = plain stream order.  Process-wide. */
int bd_set_pdl(int enabled);

/* Kernel generation used by bd_attention_tc: 1 = warp-specialised (loader / MMA issuer / two
 * softmax warpgroups, probabilities kept in TMEM; default), 0 = first-generation kernel (kept for
 * A/B measurements).  Process-wide; not meant to be flipped while launches are in flight. */
int bd_attention_tc_select(int impl);
/* The two halves of bd_attention_tc, for keys / values that exist long before their queries (the
 * decoder's memory K / V): bd_attention_tc_pack_kv writes the K / V^T operand tiles into `workspace`
 * (any stream, any time after K / V are complete), bd_attention_tc_packed attends over them.  Same
 * B, H, Lq, Lk, split and workspace in both calls. */
int bd_attention_tc_pack_kv(const void *K, int ldk, long long sk_b, const void *V, int ldv,
                            long long sv_b, int kv_half, int B, int H, int Lq, int Lk, int hd,
                            int split, void *workspace, bd_stream_t stream);
int bd_attention_tc_packed(const void *Q, int ldq, long long sq_b,
                           const unsigned char *key_padding_mask, void *O, int ldo, long long so_b,
                           int io_half, int B, int H, int Lq, int Lk, int hd, float scale, int split,
                           void *workspace, bd_stream_t stream);
/* bd_attention_tc with 16-bit tensors in HBM: io_half bit 0 = Q, bit 1 = K, bit 2 = V, bit 3 = O are
 * fp16 (leading dimensions / batch strides in halfs, rows 8-byte aligned; kv_half of
 * bd_attention_tc_pack_kv = both K and V; bd_attention_tc_packed reads bits 0 and 3). */
int bd_attention_tc_h(const void *Q, int ldq, long long sq_b, const void *K, int ldk, long long sk_b,
                      const void *V, int ldv, long long sv_b,
                      const unsigned char *key_padding_mask, void *O, int ldo, long long so_b,
                      int io_half, int B, int H, int Lq, int Lk, int hd, float scale, int split,
                      void *workspace, bd_stream_t stream);
/* Tuning hook: key sequences of at most `nk` tiles of 128 keys run as one query tile per CTA with
 * two CTAs per SM (256 TMEM columns each); longer ones as two ping-ponged query tiles per CTA.
 * Default: every length in the fp16 mode (measured faster at all of the model's shapes), none in the
 * bf16x3 mode; 0 = always the two-tile kernel; any other value applies to both modes. */
int bd_attention_tc_set_small_nk(int nk);
/* fp16 K AND V rows in HBM (io_half bits 1 and 2), split 1, dense batches (sk_b = Lk * ldk, sv_b = Lk * ldv), ldk / ldv
 * multiples of 8 halfs, K / V 16-byte aligned: bd_attention_tc_h runs WITHOUT the pack kernel — the attention
 * kernel's loader fetches each head's K and V tile from the projection output with one tensor copy each (128-byte
 * swizzle; V stays [key][dim]: MN-major B operand; softmax denominators from an N = 16 MMA against a ones tile) —
 * and `workspace` may be NULL.  on = 0 switches this off (A/B reference: the pack kernel).  Default on. */
int bd_attention_tc_set_direct(int on);

/* Hungarian matcher on the device (reference: models/losses.py:256-331, which builds the cost matrix in torch,
 * copies it to the host and calls scipy.optimize.linear_sum_assignment per scene).  Targets of all scenes are
 * concatenated; scene b owns targets tgt_offset[b] .. tgt_offset[b+1]-1 (tgt_offset: B+1 ints, device).
 * bd_matcher_cost: cost[t][q] (target-major, Q floats per target) = w_bbox * L1(boxes[b,q], tgt_boxes[t])
 *   + w_class * -(softmax(logits[b,q]) . positive_map[t])  (labels != NULL: -softmax(...)[labels[t]] instead)
 *   + w_giou * -GIoU3D(corners(boxes[b,q]), corners(tgt_boxes[t])).  logits (B,Q,C) C <= 512, boxes (B,Q,6) and
 *   tgt_boxes (T,6) as cx cy cz w h d, positive_map (T, ld_pm >= C).
 * bd_hungarian: optimal assignment of every scene's targets to distinct queries (targets per scene <= max_targets
 *   <= Q <= 1024): match_q / match_t [tgt_offset[b] + k], k < T_b, sorted by query index — the (row_ind, col_ind)
 *   linear_sum_assignment returns for the (Q x T_b) matrix.  A scene whose costs leave no finite augmenting path
 *   (inf / NaN) gets -1 entries and sets *status (optional, device int) to 1. */
int bd_matcher_cost(const float *logits, const float *boxes, const float *tgt_boxes,
                    const float *positive_map, int ld_pm, const long long *labels, const int *tgt_offset,
                    int B, int Q, int C, float w_class, float w_bbox, float w_giou, float *cost,
                    bd_stream_t stream);
int bd_hungarian(const float *cost, const int *tgt_offset, int B, int Q, int max_targets,
                 long long *match_q, long long *match_t, int *status, bd_stream_t stream);

/* bd_linear_tc_h with fp16 rows in AND out, no second operand and more row tiles than SMs runs as a PERSISTENT
 * kernel (csrc/gemm_stream.cu: operand copies of tile i + 1 under the MMAs, the TMEM drain and the store of tile
 * i; bit-identical results).  on = 0 switches it off (A/B reference).  Default on. */
int bd_linear_stream_set(int on);

/* Narrow-input layers, K <= 8 (the first layer of the learned position embeddings: models/modules.py
 * PositionEmbeddingLearned on xyz / centre + size / detected boxes): Y (M,N) = act(A (M,K) W^T + bias), W (N,K) in
 * torch's layout, fp32 rows or (y_half != 0) fp16 rows — the operand of the layer that follows; ldy even, Y 8-byte
 * aligned, relu 0 / 1.  Same operation order as bd_linear_f32. */
int bd_linear_smallk(const float *A, int lda, const float *W, const float *bias, void *Y, int ldy,
                     int y_half, int M, int N, int K, int relu, bd_stream_t stream);

/* Key sequences of at most 136 keys (one key tile of 128 + up to 8 keys handled on the FMA pipe: the 132 detected
 * boxes) with fp16 K / V rows and head_dim 36 run a dedicated kernel: P in shared memory, O over the score columns
 * — 128 TMEM columns, four CTAs per SM (these launches are bound by the per-CTA latency chain, not by arithmetic).
 * on = 0: the general kernel (A/B reference).  Default on. */
int bd_attention_tc_set_short(int on);

/* RoBERTa input embeddings (text side, reference call site models/bdetr.py:168 -> transformers
 * RobertaEmbeddings.forward): Y (B*L, D) = LayerNorm(word[ids] + position[pid] + token_type[0]) with
 * pid = pad_idx + running count of non-pad tokens (pad tokens: pad_idx).  ids (B,L) int64; word (vocab,D),
 * position (n_pos,D), type (>=1,D) fp32; D <= 1024.  Out-of-table ids are clamped. */
int bd_roberta_embed(const long long *ids, const float *word, int vocab, const float *pos, int n_pos,
                     const float *type, const float *gamma, const float *beta, float *Y, int B, int L,
                     int D, int pad_idx, float eps, bd_stream_t stream);

/* torch.topk(sigmoid(logits), k)[1].int() (models/bdetr.py:181-184): per batch row of n
 * logits, indices of the k largest sigmoid values, descending, ties -> lower index. n <= 4096 */
int bd_topk_sigmoid(const float *logits, int B, int n, int k, int *idx, bd_stream_t stream);

/* F.normalize(x, p=2, dim=-1, eps=1e-12) on rows of D floats (in place allowed). */
int bd_l2_normalize_rows(const float *X, float *Y, int M, int D, bd_stream_t stream);

/* out[r, 0:w] = table[ids[r], 0:w]  (nn.Embedding lookup, int64 ids). */
int bd_embedding_rows(const float *table, int w, const long long *ids, int M, float *out,
                      int ld_out, bd_stream_t stream);

/* center = base_xyz + residual ; generic Y = X1 + X2 on (M,w) with leading dims. */
int bd_add_rows(const float *X1, int ld1, const float *X2, int ld2, float *Y, int ldy, int M, int w,
                bd_stream_t stream);

/* Y[r,:] = [ X1[r,0:w1] | X2[r,0:w2] ]  — query_pos = cat(base_xyz, base_size) (models/bdetr.py:287) */
int bd_concat_rows(const float *X1, int ld1, int w1, const float *X2, int ld2, int w2, float *Y,
                   int ldy, int M, bd_stream_t stream);

/* out[b, c, j] = in[b, j, c] — token-major (B,n,C) -> channel-major (B,C,n) for the
 * end_points entries the reference returns channel-major (backbone_module.py:118-139). */
int bd_transpose_rows(const float *in, int B, int n, int C, float *out, bd_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* BUTD_B200_H_ */
