/*
 * comb200.h — C-ABI of libcomb200.so: the B200 (sm_100a) voxel-detector hot path of COM / OpenPCDet.
 *
 * Plain pointers + sizes + a CUDA stream in, an int status out.  No torch types.  Every entry point
 * names the reference interface it replaces (paths relative to the reference tree, `file:line`).
 *
 * Conventions
 *  - All buffer pointers are DEVICE pointers unless the parameter name ends in `_host`.
 *  - `stream` is a `cudaStream_t` passed as `void*` (0 = legacy default stream).  Nothing here
 *    synchronises the stream except the functions documented as "blocking".
 *  - Return value: 0 on success; <0 on error (COMB_E*).  `comb_last_error()` returns a thread-local
 *    human readable message.  (Reference convention is `fprintf(stderr)+exit(-1)`,
 *    pcdet/ops/iou3d_nms/src/iou3d_nms.cpp:14-38; status codes are strictly friendlier.)
 *  - Row counts that are only known on the device are passed as `const int* n_dev` (nullable);
 *    `n_max` then bounds the launch and threads beyond `*n_dev` exit.  This keeps the whole
 *    voxelize -> backbone -> BEV chain free of host synchronisation.
 *  - Coordinates are int32 rows (b, z, y, x) exactly like `voxel_coords` after
 *    pcdet/datasets/dataset.py:254-259.
 */
#ifndef COMB200_H_
#define COMB200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define COMB_OK 0
#define COMB_EINVAL (-1)  /* bad argument */
#define COMB_ECUDA (-2)   /* CUDA runtime error (message in comb_last_error) */
#define COMB_ERANGE (-3)  /* a size exceeds what the 32-bit key space supports */
#define COMB_ENODEV (-4)  /* no sm_100 device */

#define COMB_DT_F32 0
#define COMB_DT_BF16 1

/* Epilogue flags of the sparse convolution (comb_spconv_fwd_*). */
#define COMB_EPI_BIAS 1      /* out += bias[co]                      */
#define COMB_EPI_AFFINE 2    /* out = out*scale[co] + shift[co]  (eval-mode BatchNorm1d folded) */
#define COMB_EPI_RESIDUAL 4  /* out += residual[o, co]  (SparseBasicBlock identity add)       */
#define COMB_EPI_RELU 8      /* out = max(out, 0)                    */

/* ---- library ------------------------------------------------------------------------------- */
int comb_version(void);
const char* comb_last_error(void);
/* Number of SMs of the current device (148 on B200); <0 on error. */
int comb_sm_count(void);
/* Number of kernels this library has launched in the calling process so far (monotonic). */
long long comb_launch_count(void);

/* ---- a1/a2/a4: voxelization (+ fused MeanVFE) ------------------------------------------------
 * Replaces spconv.utils.Point2VoxelCPU3d.point_to_voxel / VoxelGenerator.generate as called by
 * VoxelGeneratorWrapper.generate (pcdet/datasets/processor/data_processor.py:44-60) and, when
 * `mean_out` is given, MeanVFE.forward (pcdet/models/backbones_3d/vfe/mean_vfe.py:14-31).
 *
 * Semantics (bit-exact with the sequential generator): for each frame independently, scanning
 * points in order: c_j = floor((p_j - range_min_j) / vsize_j) in fp32 (IEEE subtract + divide);
 * drop the point unless 0 <= c_j < grid_j on all axes; voxel ids are handed out in order of first
 * appearance, at most `max_voxels` per frame (later new voxels are dropped, their points too);
 * each voxel keeps its first `max_points` points in point order.
 *
 * points       [n_total, C] fp32, frames concatenated.
 * frame_offsets_host [batch+1] host ints, frame b = rows [off[b], off[b+1]);  OR
 * frame_offsets_dev  [batch+1] device ints with the same meaning (host pointer ignored): the launch sequence
 *              then depends only on n_cap (an upper bound of off[batch], also used to size the workspace), so
 *              the call can be captured in a CUDA graph and replayed with different frames.
 * voxels       [batch*max_voxels, max_points, C] fp32 or NULL (rows [0,total) written, zero padded)
 * coords       [batch*max_voxels, 4] int32 (b, z, y, x); frames are packed back to back
 * num_points   [batch*max_voxels] int32
 * mean_out     NULL or [batch*max_voxels, mean_ld] (dtype mean_dtype): mean over the kept points of
 *              channels [mean_c0, C); columns beyond C-mean_c0 up to mean_ld are zero filled
 * counts       [batch+1] int32 device: voxels per frame, then the total.
 * workspace    comb_voxelize_workspace_bytes(...) bytes, 256-byte aligned.
 */
size_t comb_voxelize_workspace_bytes(int n_total, int batch, int max_voxels, int max_points);
int comb_voxelize(const float* points, const int* frame_offsets_host, const int* frame_offsets_dev, int n_cap,
                  int batch, int C,
                  const float* vsize_xyz_host, const float* range_xyz_host,
                  int max_points, int max_voxels,
                  float* voxels, int* coords, int* num_points,
                  void* mean_out, int mean_dtype, int mean_c0, int mean_ld,
                  int* counts, void* workspace, size_t workspace_bytes, void* stream);

/* a4 stand-alone: MeanVFE over an existing (M, T, C) voxel tensor (mean_vfe.py:26-29).
 * num_points may be int32 (num_is_float=0) or fp32 (=1, as after load_data_to_gpu). */
int comb_mean_vfe(const float* voxels, const void* num_points, int num_is_float, int M, int T, int C,
                  float* out, void* stream);

/* ---- a6: coordinate hash table + rulebook ----------------------------------------------------
 * Replaces the indice-pair generation inside spconv's SubMConv3d / SparseConv3d forward
 * (call sites pcdet/models/backbones_3d/spconv_backbone.py:191-232).
 *
 * The hash table maps linear key ((b*D+z)*H+y)*W+x -> row.  `slots` must be a power of two
 * >= 2*n_max (comb_hash_slots).  table = slots * 8 bytes (uint32 key, int32 row interleaved).
 */
int comb_hash_slots(int n_max);
int comb_hash_build(const int* coords, int n_max, const int* n_dev, int batch, int D, int H, int W,
                    void* table, int slots, void* stream);

/* Output coordinate set of a strided SparseConv3d, in canonical order (ascending linear key of
 * the OUTPUT grid).  out = floor((in + 2p - d(k-1) - 1)/s) + 1 per axis is computed by the caller
 * and passed as oD,oH,oW.  bitmap workspace: comb_outcoords_workspace_bytes(batch,oD,oH,oW).
 * out_coords [out_cap,4]; out_count (device int) receives the number of rows (clamped to out_cap).
 */
size_t comb_outcoords_workspace_bytes(int batch, int oD, int oH, int oW);
int comb_conv_out_coords(const int* in_coords, int n_max, const int* n_dev, int batch,
                         int oD, int oH, int oW,
                         const int* ksize, const int* stride, const int* pad, const int* dil,
                         int* out_coords, int out_cap, int* out_count,
                         void* workspace, size_t workspace_bytes, void* stream);

/* Gather-form rulebook: nbr[k*ld + o] = input row feeding output row o through kernel offset
 * k = (kz*KH + ky)*KW + kx, i.e. the row whose coordinate is o*s - p + k*d, or -1.
 * For SubMConv3d pass out_coords == in_coords, stride 1, pad = k/2 (spconv ignores the user's
 * padding for SubM).  The classic pair lists are P_k = {(nbr[k][o], o) : nbr[k][o] >= 0}. */
int comb_nbrmap_build(const int* out_coords, int no_max, const int* no_dev,
                      const void* in_table, int in_slots, int batch, int iD, int iH, int iW,
                      const int* ksize, const int* stride, const int* pad, const int* dil,
                      int* nbr, int ld, void* stream);

/* Scatter-form (transposed) rulebook used by dgrad and SparseInverseConv3d:
 * nbr_t[k*ld_t + i] = output row o with nbr[k][o] == i, or -1.  */
int comb_nbrmap_transpose(const int* nbr, int K, int no_max, const int* no_dev, int ld,
                          int* nbr_t, int ni_max, int ld_t, void* stream);

/* Compact pair lists in spconv's `indice_pairs` layout: pairs[(0*K+k)*ld + j] = in row,
 * pairs[(1*K+k)*ld + j] = out row for j < pair_num[k], ordered by ascending out row; rest -1. */
int comb_nbrmap_to_pairs(const int* nbr, int K, int no_max, const int* no_dev, int ld,
                         int* pairs, int* pair_num, void* stream);

/* ---- a6 (fused pipeline): bitmap-rank grid index ----------------------------------------------
 * Rows of every level are kept in canonical order (ascending linear key), so the coordinate index of a
 * level is a bitmap over the grid (1 bit per cell) plus the number of set bits before every 32-byte
 * block: row(key) = prefix[key>>8] + popcount of the block's bits below key.  No hash table, no sort.
 *
 * comb_index_build: ksize == NULL -> index of `coords` themselves (D,H,W = their grid);
 *                   ksize != NULL -> index of the OUTPUT set of the strided conv (ksize,stride,pad,dil)
 *                                    applied to `coords` (D,H,W = the output grid).
 *   bitmap [comb_index_bitmap_bytes], prefix [comb_index_prefix_bytes]; out_coords (nullable)
 *   [out_cap,4] receives the rows in key order, out_count (device int) their number (clamped to out_cap).
 * comb_index_rank: rows[i] = row of coords[i] in the index, -1 if absent (voxel order -> key order).
 * comb_index_rank_scatter: comb_index_rank that ALSO writes sorted_coords[rows[i]] = coords[i] — for UNIQUE coords
 *   (a voxel list) this is the row list in key order without enumerating the bitmap (build the index with
 *   out_coords == NULL, which then only finalises the block prefixes).
 * comb_nbrmap_build_indexed: same contract as comb_nbrmap_build with the index of the INPUT level. */
size_t comb_index_bitmap_bytes(int batch, int D, int H, int W);
size_t comb_index_prefix_bytes(int batch, int D, int H, int W);
int comb_index_build(const int* coords, int n_max, const int* n_dev, int batch, int D, int H, int W,
                     const int* ksize, const int* stride, const int* pad, const int* dil,
                     void* bitmap, void* prefix, int* out_coords, int out_cap, int* out_count, void* stream);
int comb_index_rank(const int* coords, int n_max, const int* n_dev, int batch, int D, int H, int W,
                    const void* bitmap, const void* prefix, int* rows, void* stream);
int comb_index_rank_scatter(const int* coords, int n_max, const int* n_dev, int batch, int D, int H, int W,
                            const void* bitmap, const void* prefix, int* rows, int* sorted_coords,
                            int sorted_cap, void* stream);
int comb_nbrmap_build_indexed(const int* out_coords, int no_max, const int* no_dev,
                              const void* bitmap, const void* prefix, int batch, int iD, int iH, int iW,
                              const int* ksize, const int* stride, const int* pad, const int* dil,
                              int* nbr, int ld, void* stream);

/* ---- a7/a8/a9: sparse convolution ------------------------------------------------------------
 * Replaces the gather-GEMM-scatter of spconv's SparseConvolution.forward/backward.
 * Weights are in spconv-2.x layout [Cout, K, Cin] (= (Cout,kz,ky,kx,Cin), one of the layouts
 * pcdet/models/detectors/detector3d_template.py:337-348 adapts checkpoints to).
 *
 *   out[o,:] = epi( sum_k  in[nbr[k][o], :] . W[:,k,:]^T )
 *
 * fp32 check mode (CUDA cores, fixed summation order k then ci). */
int comb_spconv_fwd_f32(const float* in_feats, int Cin, const float* weight, int K, int Cout,
                        const int* nbr, int ld, int no_max, const int* no_dev,
                        int epi_flags, const float* bias, const float* scale, const float* shift,
                        const float* residual, float* out, void* stream);
/* dgrad: din[i,:] = sum_k dout[nbr_t[k][i], :] . W[:,k,:]   (same kernel family, W not transposed) */
int comb_spconv_dgrad_f32(const float* dout, int Cout, const float* weight, int K, int Cin,
                          const int* nbr_t, int ld_t, int ni_max, const int* ni_dev,
                          float* din, void* stream);
/* wgrad: dW[co,k,ci] = sum_o dout[o,co] * in[nbr[k][o], ci];  dW is overwritten. */
int comb_spconv_wgrad_f32(const float* in_feats, int Cin, const float* dout, int Cout, int K,
                          const int* nbr, int ld, int no_max, const int* no_dev,
                          float* dweight, void* stream);

/* bf16 tensor-core path: tcgen05.mma (M=128 rows x N=Cout, K chunks of 64 = packed kernel
 * offsets), fp32 accumulators in TMEM, gathered A rows staged with cp.async into 128B-swizzled
 * shared memory, packed weights streamed with cp.async.bulk.
 *  in_feats  [ni, Cin_p] bf16 (Cin_p = Cin rounded up to 16, zero padded)
 *  wpacked   image produced by comb_spconv_pack_weight_bf16 (bytes: comb_spconv_packed_bytes)
 *  out       [no, Cout] bf16 (out_dtype=COMB_DT_BF16) or fp32
 *  residual  [no, Cout] bf16 or NULL
 * Supported: Cin_p, Cout in {16, 32, 64, 128}. */
size_t comb_spconv_packed_bytes(int Cin_p, int K, int Cout);
int comb_spconv_pack_weight_bf16(const float* weight, int Cout, int K, int Cin, int Cin_p,
                                 void* wpacked, void* stream);
int comb_spconv_fwd_bf16(const void* in_feats, int Cin_p, const void* wpacked, int K, int Cout,
                         const int* nbr, int ld, int no_max, const int* no_dev,
                         int epi_flags, const float* bias, const float* scale, const float* shift,
                         const void* residual, void* out, int out_dtype, void* stream);

/* wgrad on the tensor cores (a8): dW[co,k,ci] = sum_o dout[o,co] * in[nbr[k][o], ci] with bf16 operands and fp32
 * accumulation in TMEM (tcgen05.mma, both operands MN-major: the gathered rows are used as they are gathered).
 *  in_feats [ni, Cin_p] bf16 (zero padded beyond Cin), dout [no, Cout] bf16, dweight [Cout, K, Cin] fp32 (overwritten).
 *  workspace: comb_spconv_wgrad_bf16_workspace_bytes(...) bytes of device memory (per-row-chunk partial sums that a
 *  second kernel adds in a fixed order: the result is deterministic).  Supported: Cin_p, Cout in {16, 32, 64, 128}.
 * Replaces the wgrad GEMMs of spconv's SparseConvolution backward (spconv_backbone.py:12-15,38-45,191-232). */
size_t comb_spconv_wgrad_bf16_workspace_bytes(int Cin_p, int Cin, int Cout, int K, int no_max);
int comb_spconv_wgrad_bf16(const void* in_feats, int Cin_p, int Cin, const void* dout, int Cout, int K,
                           const int* nbr, int ld, int no_max, const int* no_dev, float* dweight,
                           void* workspace, size_t workspace_bytes, void* stream);

/* Debug hook: CTA 0 of every following comb_spconv_fwd_bf16 launch records clock64 stamps of its pipeline
 * events (first 512 chunks, 8 int64 slots each) into `buf` (device, 32 KB); NULL switches tracing off. */
int comb_debug_conv_trace(void* buf);

/* Elementwise helpers used between convolutions (a9: BatchNorm1d(eval) + ReLU + residual). */
int comb_affine_relu(const void* x, int dtype, int n_max, const int* n_dev, int C,
                     const float* scale, const float* shift, const void* residual, int relu,
                     void* out, void* stream);
int comb_cast_pad(const float* x, int n_max, const int* n_dev, int C, void* out_bf16, int ld,
                  void* stream);
/* Row permutation: out[row_map[r], :] = in[r, :] (scatter=1) or out[r, :] = in[row_map[r], :] (scatter=0) for
 * r < n; rows whose map entry is negative are skipped (scatter) / zero filled (gather).  `row_bytes` is
 * a multiple of 4.  Used to move features between voxel order and key order. */
int comb_permute_rows(const void* in, const int* row_map, int n_max, const int* n_dev, int row_bytes,
                      int scatter, void* out, void* stream);

/* ---- a9, training form: BatchNorm1d(train) + ReLU + residual over the active rows ------------------------
 * Replaces the eager chain between two sparse convolutions in train() mode — SparseSequential(conv,
 * nn.BatchNorm1d(eps=1e-3, momentum=0.01), nn.ReLU()) (pcdet/models/backbones_3d/spconv_backbone.py:21-25) and
 * SparseBasicBlock.forward's bn -> relu -> conv -> bn -> (+identity) -> relu (:50-66) — for the fused training step.
 * x (the convolution output) is fp32 or bf16 (x_dtype = COMB_DT_*); residual, out, dy, act, dx, g_out are bf16;
 * all [n_max, C] row-major, C in {16,32,64,128}; n_dev (device int, may
 * be NULL) is the live row count.  Statistics are reduced deterministically (per-block partial sums in fp64).
 *   fwd: batch mean / biased variance over the rows; running_mean/var (may be NULL) updated like torch (momentum,
 *        unbiased variance); out = relu?(x*gamma*invstd + beta - mean*gamma*invstd (+ residual)); save_mean /
 *        save_invstd [C] are kept for the backward pass.
 *   bwd: g = dy * (act > 0) if relu else dy (act = the forward's out); dgamma = sum g*xhat, dbeta = sum g,
 *        dx = gamma*invstd*(g - dbeta/n - xhat*dgamma/n); g_out (may be NULL) receives g, the gradient of the
 *        residual branch.
 * comb_col_sum: sum[c] = sum over rows of x[:, c] (bias gradient of a convolution).
 * workspace: comb_bn_workspace_bytes(C). */
size_t comb_bn_workspace_bytes(int C);
int comb_bn_train_fwd(const void* x, int x_dtype, int n_max, const int* n_dev, int C, const float* gamma, const float* beta,
                      float eps, float momentum, float* running_mean, float* running_var, const void* residual,
                      int relu, void* out, float* save_mean, float* save_invstd, void* workspace,
                      size_t workspace_bytes, void* stream);
int comb_bn_train_bwd(const void* dy, const void* act, const void* x, int x_dtype, int n_max, const int* n_dev, int C,
                      const float* gamma, const float* save_mean, const float* save_invstd, int relu, void* dx,
                      void* g_out, float* dgamma, float* dbeta, void* workspace, size_t workspace_bytes, void* stream);
int comb_col_sum(const void* x, int n_max, const int* n_dev, int C, float* sum, void* workspace,
                 size_t workspace_bytes, void* stream);

/* ---- a10: HeightCompression / SparseConvTensor.dense() ---------------------------------------
 * Replaces encoded_spconv_tensor.dense() (pcdet/models/backbones_2d/map_to_bev/
 * height_compression.py:21): out[b, c, z, y, x] = feats[row, c], zero elsewhere; out is fp32
 * NCDHW so that .view(N, C*D, H, W) is free.  The whole tensor is written (no prior memset). */
int comb_dense(const void* feats, int dtype, const int* coords, int n_max, const int* n_dev,
               int batch, int C, int D, int H, int W, float* out, void* workspace,
               size_t workspace_bytes, void* stream);
size_t comb_dense_workspace_bytes(int batch, int D, int H, int W);
/* Scatter form of the same operation: `out` must already be zero (clear it early, off the critical path — the
 * zero-fill of the 36 MB/frame BEV tensor does not depend on the features); only the active cells are written. */
int comb_dense_scatter(const void* feats, int dtype, const int* coords, int n_max, const int* n_dev,
                       int batch, int C, int D, int H, int W, float* out, void* stream);

/* Adjoint of dense() for training through HeightCompression (the reference relies on spconv's autograd through
 * SparseConvTensor.dense(), height_compression.py:21): grad_feats[row, c] = grad_dense[b, c, z, y, x]. */
int comb_dense_gather(const float* dense, const int* coords, int n_max, const int* n_dev, int batch, int C,
                      int D, int H, int W, void* grad_feats, int dtype, void* stream);

/* ---- a11/a12: points in boxes -----------------------------------------------------------------
 * comb_points_in_boxes_mask replaces points_in_boxes_cpu (pcdet/ops/roiaware_pool3d/src/
 * roiaware_pool3d.cpp:143-168, MARGIN 1e-2): mask[b*P + p] = 1 iff point p lies in box b.
 * To be bit-identical with the reference's glibc cosf/sinf the per-box rotation is supplied by
 * the caller: box_trig [nb,2] = (cosf(-rz), sinf(-rz)) computed with the host libm
 * (comb_box_trig_host does exactly that).  Arithmetic is IEEE fp32 without FMA contraction and
 * fp64 thresholds, like the reference compiled by g++ for x86-64.
 *
 * comb_points_in_boxes_index replaces points_in_boxes_gpu (roiaware_pool3d_kernel.cu:313-336,
 * MARGIN 1e-5): idx[b*P+p] = first box of frame b containing the point, or -1; trigonometry and
 * contraction follow the reference's device code. */
void comb_box_trig_host(const float* boxes_host, int nb, float* trig_host);
int comb_points_in_boxes_mask(const float* points, int P, int point_stride, const float* boxes,
                              const float* box_trig, int nb, int* mask, void* stream);
int comb_points_in_boxes_index(const float* points, const float* boxes, int batch, int P, int T,
                               int* idx, void* stream);
/* comb_points_in_any_box: any[p] = 1 iff point p lies in at least one box (P bytes) — the only thing
 * remove_points_in_boxes3d (pcdet/utils/box_utils.py:117-131) and COMAug's point removal
 * (pcdet/datasets/augmentor/database_sampler_v2.py:535-539) use of the (Nb,P) mask: `mask.sum(0) != 0`.
 * Same arithmetic and box_trig contract as comb_points_in_boxes_mask. */
int comb_points_in_any_box(const float* points, int P, int point_stride, const float* boxes,
                           const float* box_trig, int nb, unsigned char* any, void* stream);

/* ---- a13/a14: rotated BEV IoU -----------------------------------------------------------------
 * flavour 0 ("cpu"): replaces boxes_iou_bev_cpu (pcdet/ops/iou3d_nms/src/iou3d_cpu.cpp:232-252);
 *   trig_a/trig_b [n,4] = (cosf(rz), sinf(rz), cosf(-rz), sinf(-rz)) from the host libm
 *   (comb_box_trig4_host), no FMA contraction.
 * flavour 1 ("gpu"): replaces boxes_iou_bev_gpu / boxes_overlap_bev_gpu (iou3d_nms.cpp:49-88,
 *   kernels iou3d_nms_kernel.cu:236-265); trig pointers may be NULL (computed on device).
 * what: 0 = IoU, 1 = overlap area. out [na, nb] fp32. */
void comb_box_trig4_host(const float* boxes_host, int n, float* trig_host);
int comb_boxes_bev(const float* boxes_a, const float* trig_a, int na,
                   const float* boxes_b, const float* trig_b, int nb,
                   int flavour, int what, float* out, void* stream);

/* ---- a15/a16: NMS ------------------------------------------------------------------------------
 * Replaces nms_gpu / nms_normal_gpu (iou3d_nms.cpp:90-188 + kernels :267-372): boxes must already
 * be sorted by descending score (the Python wrapper does that, iou3d_nms_utils.py:91-95).
 * The suppression bit-matrix and the greedy sweep both run on the device; `keep` [n] int64 and
 * `num_keep` (1 int) are device buffers.  rotated=1 -> rotated BEV IoU, 0 -> axis aligned.
 * flavour as in comb_boxes_bev (trig may be NULL for flavour 1).
 * workspace: comb_nms_workspace_bytes(n). */
size_t comb_nms_workspace_bytes(int n);
int comb_nms(const float* boxes, const float* trig, int n, float thresh, int rotated, int flavour,
             long long* keep, int* num_keep, void* workspace, size_t workspace_bytes, void* stream);

/* Same with the number of boxes on the device: n_dev (may be NULL) holds the live count, n_max sizes the launch and
 * the workspace — for callers whose box list is produced by a previous kernel (comb_centerhead_decode_nms). */
int comb_nms_dev(const float* boxes, const float* trig, int n_max, const int* n_dev, float thresh, int rotated,
                 int flavour, long long* keep, int* num_keep, void* workspace, size_t workspace_bytes, void* stream);

/* ---- f2: CenterHead post-processing ------------------------------------------------------------------------------
 * Replaces CenterHead.generate_predicted_boxes (pcdet/models/dense_heads/center_head.py:266-317) =
 * centernet_utils.decode_bbox_from_heatmap (pcdet/models/model_utils/centernet_utils.py:199-279: _topk, gathers,
 * atan2 / exp / grid->metric decode, range + score mask) followed by model_nms_utils.class_agnostic_nms
 * (model_nms_utils.py:6-25, NMS_TYPE nms_gpu) for one separate head, with no host round trip:
 *   hm [B,C,H,W] RAW logits (sigmoid is applied here), center [B,2,H,W], center_z [B,1,H,W], dim [B,3,H,W] RAW (exp is
 *   applied here), rot [B,2,H,W] = (cos, sin); all fp32 NCHW, contiguous.
 *   K = MAX_OBJ_PER_SAMPLE (<= 1024); limit_range = POST_CENTER_LIMIT_RANGE (6 host floats); label_map (device, C
 *   ints, may be NULL) = class_id_mapping_each_head of the head; labels come out 1-based like the reference's.
 * Outputs are capacity-sized: out_boxes [B,K,7], out_scores [B,K], out_labels [B,K] int32, out_counts [B] int32 (the
 * number of detections per frame, on the device).  workspace: comb_centerhead_workspace_bytes(B, K). */
size_t comb_centerhead_workspace_bytes(int B, int K);
int comb_centerhead_decode_nms(const float* hm, const float* center, const float* center_z, const float* dim,
                               const float* rot, int B, int C, int H, int W, int K, float stride, float vx, float vy,
                               float rx, float ry, const float* limit_range, float score_thresh, const int* label_map,
                               float nms_thresh, int nms_pre_max, int nms_post_max, float* out_boxes,
                               float* out_scores, int* out_labels, int* out_counts, void* workspace,
                               size_t workspace_bytes, void* stream);

/* Heads with a velocity branch ('vel' in SEPARATE_HEAD_CFG.HEAD_ORDER: tools/cfgs/nuscenes_models/cbgs_*_centerpoint.yaml,
 * tools/cfgs/waymo_models/centerpoint_4frames.yaml; centernet_utils.py:241-245): vel [B,2,H,W] is gathered at the top-K
 * cells and rides through mask / NMS / gather as columns 7..8 — out_boxes is then [B,K,9]; the NMS itself sees the first
 * seven columns (model_nms_utils.py:13 boxes_for_nms[:, 0:7]).  vel == NULL is comb_centerhead_decode_nms. */
int comb_centerhead_decode_nms_vel(const float* hm, const float* center, const float* center_z, const float* dim,
                                   const float* rot, const float* vel, int B, int C, int H, int W, int K, float stride,
                                   float vx, float vy, float rx, float ry, const float* limit_range, float score_thresh,
                                   const int* label_map, float nms_thresh, int nms_pre_max, int nms_post_max,
                                   float* out_boxes, float* out_scores, int* out_labels, int* out_counts,
                                   void* workspace, size_t workspace_bytes, void* stream);

/* ---- f4: the BEV tensor in channels-last bf16 -------------------------------------------------------------------------
 * HeightCompression (pcdet/models/backbones_2d/map_to_bev/height_compression.py:21-24: dense() + view(N, C*D, H, W))
 * for a 2D backbone (pcdet/models/backbones_2d/base_bev_backbone.py:81-112) that runs in bf16 NHWC: the rows are
 * scattered into out[b][y][x][c*D + z] (bf16, [batch,H,W,C*D], ZERO-filled by the caller); feats are fp32 or bf16
 * [n,C], coords int32 (b,z,y,x).  comb_dense_gather_nhwc is the adjoint (grad fp32 or bf16 in the same layout). */
int comb_dense_scatter_nhwc_bf16(const void* feats, int dtype, const int* coords, int n_max, const int* n_dev, int batch,
                                 int C, int D, int H, int W, void* out, void* stream);
int comb_dense_gather_nhwc(const void* grad, int grad_dtype, const int* coords, int n_max, const int* n_dev, int batch,
                           int C, int D, int H, int W, void* out, int dtype, void* stream);

/* ---- f3: COMAug placement test ------------------------------------------------------------------------------------------
 * The collision test of the COMAug database sampler (pcdet/datasets/augmentor/database_sampler_v2.py:600-604):
 *   valid[i] = (max_j iou1[i][j] + max_j iou2[i][j]) == 0, iou2's diagonal taken as 0, iou1 replaced by iou2 when E == 0,
 * from the two BEV IoU matrices (comb_boxes_bev, flavour 0: sampled x existing [S,E], sampled x sampled [S,S]) where they
 * lie on the device; valid is S bytes. */
int comb_comaug_valid_mask(const float* iou1, const float* iou2, int S, int E, unsigned char* valid, void* stream);

/* ---- f1: CenterHead target assignment and the COM loss re-weighting -------------------------------------------------
 * Replaces the per-object Python loops of CurriculumCenterHead (pcdet/models/dense_heads/curriculum_center_head.py):
 * cluster (:431-473), assign_targets / assign_target_of_single_head (:120-296; CenterHead.assign_targets,
 * center_head.py:119-236, is the same without the point-count filter and the group column) with
 * centernet_utils.gaussian_radius / draw_gaussian_to_heatmap (pcdet/models/model_utils/centernet_utils.py:48-108), and
 * of FocalLossCenterCurriculum (pcdet/utils/loss_utils.py): confidence_of_all_groups (:1131-1178) and the object loop
 * of neg_loss (:1222-1287) with centernet_utils.draw_mask_to_heatmap (:110-131).  All pointers are device pointers.
 *
 * comb_centerhead_assign_targets: ONE separate head, all frames.  gt_boxes [B,M,C] fp32 (C >= 8, class id 1-based in
 *   the last column, 0 = padding), npgt [B,M] fp32, group [B,M] int64 or NULL (R = 5 needs it), cls_map [n_cls+1]
 *   int32: global class id -> index inside this head or -1.  x0,y0 = POINT_CLOUD_RANGE[0:2], vx,vy = VOXEL_SIZE[0:2],
 *   stride = FEATURE_MAP_STRIDE, (W,H) = feature map, overlap = GAUSSIAN_OVERLAP, filter_points = (epoch <=
 *   EPOCH_THRED), gtab / gtab_off / rmax: Gaussian windows of radius 0..rmax built on the host with the reference's
 *   numpy formula (gtab_off[r] = offset of the (2r+1)^2 window).  Outputs must be ZERO-filled by the caller:
 *   heatmap [B,Ch,H,W], ret_boxes [B,max_objs,C], inds [B,max_objs] int64, mask [B,max_objs] fp32, radius_map
 *   [B,max_objs,R] int64 (class, x, y, radius[, group]).  relabel_in_place = 1 reproduces the reference's side effect
 *   (curriculum_center_head.py:252-254): the class column of this head's boxes in gt_boxes is overwritten with the
 *   head-local 1-based id, which is what the NEXT head's call then reads (multi-head configurations). */
int comb_centerhead_assign_targets(const float* gt_boxes, const float* npgt, const long long* group,
                                   const int* cls_map, int n_cls, int B, int M, int C, float x0, float y0, float vx,
                                   float vy, float stride, int W, int H, int max_objs, double overlap, int min_radius,
                                   int filter_points, float min_points, const float* gtab, const int* gtab_off,
                                   int rmax, int Ch, float* heatmap, float* ret_boxes, long long* inds, float* mask,
                                   long long* radius_map, int R, int relabel_in_place, void* stream);
/* cluster(): curriculum group of each of n ground-truth boxes (gt_boxes [n,C]); the three attribute arrays are fp32. */
int comb_centerhead_cluster_groups(const float* gt_boxes, int n, int C, const float* true_object,
                                   const float* occupancy_ratio, const float* facade_type, long long* group,
                                   void* stream);
/* confidence_of_all_groups(): conf / num [n_class, n_group] fp32 from pred [B,Ch,H,W] and radius_map [B,nobj,R]. */
int comb_comloss_group_confidence(const float* pred, int B, int Ch, int H, int W, const long long* radius_map, int nobj,
                                  int R, int n_class, int n_group, float* conf, float* num, void* stream);
/* The object loop of neg_loss(): per object weight from the predicted confidence at its centre (mode 0: height / (1 +
 * exp(elongation (c - threshold))) + 1 - height / 2; 1: K (c - threshold) + 1; 2: 1), written to box_mask [B,nobj]
 * and drawn (assignment, object order) into mask [B,Ch,H,W]; active = (START <= epoch <= END). */
int comb_comloss_reweight(const float* pred, int B, int Ch, int H, int W, const long long* radius_map, int nobj, int R,
                          double threshold, double elongation, double height, double K, int mode, int fixed_radius,
                          int add_radius, int only_center, int active, float* box_mask, float* mask, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* COMB200_H_ */
