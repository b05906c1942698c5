/* pcab200 -- C ABI of the B200 (sm_100a) hot path of prs-eth/PCAccumulation.
 *
 * One entry point per fused stage of MotionNet.forward (reference models/motionnet.py:137-262), the
 * voxeliser (libs/voxel_generator.py) and the Chamfer extension (chamfer_distance/).  Conventions:
 *   - every pointer is a DEVICE pointer owned by the caller (torch-owned memory in the Python host),
 *     except the small `const float* range6 / voxel_size3` geometry arrays, which are HOST pointers;
 *   - the library allocates nothing and keeps no state besides the thread-local error string and per-device caches of
 *     immutable facts (SM count, "kernel attribute already set" marks);
 *   - calls are asynchronous on `stream`, never synchronise the device, return 0 on success or a negative
 *     PCAB_ERR_* code (pcab_last_error() gives the message); nothing is printed, nothing exits;
 *   - scratch memory is passed in; each pcab_*_workspace() returns the bytes the matching call needs;
 *   - BEV activations are NHWC, in one of two element formats selected by an `fmt` argument: 0 = float32, 1 = P16
 *     (csrc/pair16.cuh: every 32-channel group of a pixel is 32 fp16 h | 32 fp16 l with x = h + l, the operand format of the
 *     tcgen05 convolutions; same 4 bytes per element); integer outputs are bit-exact with the reference.
 *
 * The reference interface each entry replaces is cited as file:line (relative to the reference root).
 */
#ifndef PCAB200_H_
#define PCAB200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct CUstream_st* pcab_stream_t; /* == cudaStream_t */

const char* pcab_last_error(void);
int pcab_version(void);

/* ---- voxeliser: libs/voxel_generator.py:4-61,117-154 + libs/dataloader.py:33-38 ------------------------ */
size_t pcab_voxelize_workspace(int n_points, long long n_cells);
int pcab_voxelize(const float* points4 /* [N,4] x,y,z,t */, const int* point_batch /* [N] or NULL */, int n_points,
                  int batch_size, const float* range6 /* host */, const float* voxel_size3 /* host */, int n_sweeps,
                  int* coords_zyxt /* [>=M,4] */, int* pillar_batch /* [>=M] */, int* p2v /* [N], -1 = rejected */,
                  int* num_voxels /* [B] */, int* total_voxels /* [1] */, void* workspace, size_t workspace_bytes,
                  pcab_stream_t stream);

/* stable sort of the points by pillar + segment table (replaces the torch_scatter index plumbing) */
size_t pcab_pillar_index_workspace(int n_points);
int pcab_pillar_index(const int* p2v, int n_points, int n_pillars, int* order /* [N] */, int* pstart /* [M+1] */,
                      void* workspace, size_t workspace_bytes, pcab_stream_t stream);

/* models/motionnet.py:159-160: scatter(mean) of xyz, scatter(max) of the GT foreground label */
int pcab_pillar_stats(const float* xyz, const long long* fb_labels /* or NULL */, const int* order, const int* pstart,
                      int n_pillars, float* pillar_mean /* [M,3] */, int* fb_sub /* [M] or NULL */,
                      pcab_stream_t stream);

/* canvas cell of each pillar, models/pillar_encoder.py:158; cell_to_pillar must be pre-filled with -1 */
int pcab_pillar_cells(const int* coords_zyxt, const int* pillar_batch, int n_pillars, int n_sweeps, int ny, int nx,
                      int* pillar_cell /* [M] */, int* pillar_frame /* [M] = b*T+t, or NULL */,
                      int* cell_to_pillar /* [B*T*ny*nx] or NULL */, pcab_stream_t stream);

/* models/pillar_encoder.py:97-122 (PillarFeatureNet) + :125-174 (scatter to the [B*T,ny,nx,32] canvas) */
int pcab_pfn_pack_size(void);
size_t pcab_pillar_encode_workspace(int n_points, int n_pillars);
int pcab_pillar_encode(const float* xyz, const int* point_time, const int* order, const int* p2v, const int* pstart,
                       const int* coords_zyxt, const int* pillar_cell, const float* pillar_mean,
                       const float* weight_pack, int n_points, int n_pillars, const float* range6 /* host */,
                       const float* voxel_size3 /* host */, int n_sweeps, float* pillar_feats /* [M,32] */,
                       float* canvas_nhwc /* zero-filled by the caller */, int canvas_fmt, void* workspace,
                       size_t workspace_bytes, pcab_stream_t stream);

/* the same encoder on the tensor cores (tcgen05, fp16-pair operands; csrc/pillar.cu:pfn_tc): w_tc = 3 stage blobs of fp16
 * [fc_0 64x64 | [shortcut|fc_1] 2x64x64 | fc_pos 128x64 or fc_c 64x64 (+pad)] with rows = [h rows; l rows] (tc_pack.pack_pfn_tc),
 * bias_tc = per stage b0[32] b1[32] bx[64] (device), scales_inv9 = per stage 1/scale of the three matrices (HOST) */
int pcab_pillar_encode_tc(const float* xyz, const int* point_time, const int* order, const int* p2v, const int* coords_zyxt,
                          const int* pillar_cell, const float* pillar_mean, const void* w_tc, const float* bias_tc,
                          const float* scales_inv9 /* host */, int n_points, int n_pillars, const float* range6 /* host */,
                          const float* voxel_size3 /* host */, int n_sweeps, float* pillar_feats /* [M,32] */,
                          float* canvas_nhwc /* zero-filled by the caller */, int canvas_fmt, void* workspace,
                          size_t workspace_bytes, pcab_stream_t stream);

/* ---- convolutions: models/unet.py:11-113, models/stpn.py:13-22,80 ---------------------------------------- */
/* FP32 CUDA-core path; sources accumulate into one output (concat / temporal 3x3x3) */
int pcab_conv3x3_f32(const float* src0, int c0, const float* src1, int c1, const float* src2, int c2, int temporal_T,
                     const float* weight_packed /* per source [9][C_s][Cout] */, const float* bias,
                     const float* bn_scale, const float* bn_shift, int relu, float* out, int n_images, int H, int W,
                     int Cout, int out_cstride, int out_coff, pcab_stream_t stream);
int pcab_convT2x2_f32(const float* in, const float* weight_packed /* [4][Cin][Cout] */, const float* bias, float* out,
                      int n_images, int H, int W, int Cin, int Cout, int out_cstride, int out_coff,
                      pcab_stream_t stream);
int pcab_maxpool2x2(const float* in, float* out, int n_images, int H, int W, int C, int fmt, pcab_stream_t stream);
int pcab_temporal_max(const float* in, float* out, int B, int T, int H, int W, int C, int fmt, pcab_stream_t stream);

/* tcgen05 tensor-core path (3xTF32 split, FP32 accumulate in TMEM); same semantics as pcab_conv3x3_f32 */
int pcab_conv3x3_tc_supported(int n_sources, int c0, int c1, int c2, int Cout, int H, int W);
int pcab_conv3x3_tc_plan(int n_images, int H, int W, int Cout, int* out10 /* host: c, mt, strip, mtx, R, Wt, tiles_x, tiles_y, cout tiles, work items */);
size_t pcab_conv3x3_tc_pack_floats(int cin_total, int Cout);
int pcab_conv3x3_tc(const float* src0, int c0, const float* src1, int c1, const float* src2, int c2, int temporal_T,
                    const float* weight_tc_packed, const float* bias, const float* bn_scale, const float* bn_shift,
                    int relu, float* out, int n_images, int H, int W, int Cout, int out_cstride, int out_coff,
                    pcab_stream_t stream);
/* the same with fp16-pair operands (kind::f16): weights as fp16 [2][Cout][9*cin/32][64] (32 used per group), pre-multiplied by
 * 1/weight_scale_inv (a power of two) */
int pcab_conv3x3_tc_f16(const float* src0, int c0, const float* src1, int c1, const float* src2, int c2, int temporal_T,
                        const void* weight_f16_packed, float weight_scale_inv, const float* bias, const float* bn_scale,
                        const float* bn_shift, int relu, float* out, int n_images, int H, int W, int Cout, int out_cstride,
                        int out_coff, pcab_stream_t stream);

/* tcgen05 path over P16 activations (csrc/conv_p16.cu): sources and output are P16 tensors, the epilogue writes the (h, l)
 * pairs and TMA-stores them; weights as for pcab_conv3x3_tc_f16.  src0_cstride: 0, or the channel count of the tensor src0
 * points INTO (src0 = first byte of the first used 32-channel group) -- reads a channel slice of a wider tensor.
 * sat_counter (device, may be NULL) is incremented when an output beyond +-65504 was clamped.
 * Weights: fp16 [2 (h, l)][columns][Kpad], K dense in the kernel's consumption order -- per source, per 32-channel chunk,
 * per tap, 32 channels -- zero-padded to a multiple of 64 (tc_pack.pack_conv_p16 / pack_convT_p16), pre-multiplied by
 * 1/weight_scale_inv (a power of two).
 * pcab_convT2x2_p16: ConvTranspose2d(kernel 2, stride 2) (models/unet.py:24-33) as a 1-tap GEMM with 4*Cout columns on the
 * same pipeline; column = (dy*2+dx)*Cout + co. */
int pcab_conv3x3_p16_supported(int n_sources, int c0, int c1, int c2, int Cout, int H, int W);
int pcab_conv_p16_plan(int n_images, int H, int W, int Cout, int cin_total, int ntaps, int* out13 /* host */);
int pcab_conv3x3_p16(const void* src0, int c0, int src0_cstride, const void* src1, int c1, const void* src2, int c2,
                     int temporal_T, const void* weight_f16_packed, float weight_scale_inv, const float* bias,
                     const float* bn_scale, const float* bn_shift, int relu, void* out, int n_images, int H, int W, int Cout,
                     unsigned int* sat_counter, pcab_stream_t stream);
/* Conv3d 3x3x3 (32 -> 32 channels; models/stpn.py:13-22) with the temporal taps fused into the MMA N dimension: every input
 * frame is staged once and feeds output frames f-1, f, f+1; weights fp16 [2][96 = (kt 2,1,0) x 32][320] (tc_pack.pack_conv3d_fused_p16) */
int pcab_conv3d_p16(const void* src, int T, const void* weight_f16_packed, float weight_scale_inv, const float* bias, int relu,
                    void* out, int n_images /* B*T */, int H, int W, unsigned int* sat_counter, pcab_stream_t stream);
int pcab_convT2x2_p16(const void* in, int Cin, const void* weight_f16_packed, float weight_scale_inv, const float* bias,
                      void* out /* [n,2H,2W,Cout] P16 */, int n_images, int H, int W, int Cout, unsigned int* sat_counter,
                      pcab_stream_t stream);

/* ---- heads / BEV ops: models/motionnet.py:45-135,167-170,188-194 ------------------------------------------ */
int pcab_head2_conv(const float* in_nhwc, int cin, int in_cstride /* channels per pixel of `in` (>= cin) */, int fmt,
                    const float* weight_packed /* [9][cin][2] */, const float* bias, int n_images, int H, int W,
                    float* logits_nchw /* [n,2,H,W] */, int* argmax_map /* [n,H,W] */, pcab_stream_t stream);
int pcab_fb_per_point(const int* fb_map, const int* pillar_cell, const int* p2v, int n_points,
                      long long* fb_per_point, pcab_stream_t stream);
int pcab_canvases(const int* pillar_cell, const int* fb_sub, const float* pillar_mean, int n_pillars, int H, int W,
                  float* occ_map, long long* fb_map, float* mean_map /* [B*T,3,H,W] */, pcab_stream_t stream);
int pcab_warp_bev(const float* bev_nhwc, const float* pose /* [B*T,4,4] */, int B, int T, int H, int W, int C, float vx,
                  float vy, float x_min, float y_min, float* out_nhwc, int fmt, pcab_stream_t stream);
int pcab_transform_points(const float* xyz, const int* point_frame, const float* pose, int n_points, float* out,
                          pcab_stream_t stream);

/* ---- ego motion: models/egomotion.py:100-469, toolbox/register_utils.py:19-56,184-197,247-318 ------------- */
size_t pcab_bg_compact_workspace(long long n_cells);
int pcab_bg_compact(const int* cell_to_pillar, const int* fb_est, int n_frames, int H, int W, int* bg_cells,
                    int* frame_off /* [n_frames+1] */, void* workspace, size_t workspace_bytes, pcab_stream_t stream);
/* models/egomotion.py:451-455: out2 = {ego_rot_error, ego_trans_error} of [B,T,4,4] poses (recomputed after the ICP refinement) */
int pcab_ego_pose_errors(const float* ego_est, const float* ego_gt, int B, int T, float* out2, pcab_stream_t stream);
size_t pcab_ego_pairs_workspace(int npairs);
int pcab_ego_pairs(const float* geo_nhwc /* [B*T,H,W,64] */, int geo_fmt, const int* cell_to_pillar, const float* pillar_mean,
                   const int* pillar_frame, int n_pillars, const int* bg_cells, const int* frame_off,
                   const int* pair_frames /* [P,2] source,target frame */, const int* choice /* [P,2,1024] */,
                   const float* thr2 /* [P] */, int npairs, const float* alpha, const float* beta, int sinkhorn_iters,
                   const float* ego_gt /* [B*T,4,4] */, const int* chain_pair /* [B*T] */, int B, int T, int chain_mode,
                   float* perm_out /* [P,1024,1024] */, float* pose_pairs /* [P,4,4] */, float* ego_est /* [B*T,4,4] */,
                   float* ego_gt_out /* [B*T,4,4] */, float* scalars /* l1,l2,rot_err,trans_err */, void* workspace,
                   size_t workspace_bytes, pcab_stream_t stream);

/* ---- per-point stages: models/pillar_encoder.py:206-267, models/stpn.py:91-103 ---------------------------- */
size_t pcab_select_workspace(int n);
int pcab_select_indices(const int* flags, const long long* values, long long value, int n, int* idx, int* count,
                        void* workspace, size_t workspace_bytes, pcab_stream_t stream);
int pcab_ungrid(const float* feats_nhwc, int C, int fmt, int H, int W, const float* xyz, const int* frame_of_point,
                const int* idx, int k, float x_abs, float y_abs, float* out /* [k,C] */, pcab_stream_t stream);
int pcab_stpn_head_pack_size(void);
int pcab_init_point_outputs(int n_points, float* mos, float* offset, pcab_stream_t stream);
int pcab_stpn_head(const float* mos_feats_nhwc /* [B,H,W,64] */, int H, int W, const float* transformed_points,
                   const int* point_batch, const int* fg_idx, int n_fg, const float* weight_pack, float x_abs,
                   float y_abs, float* mos_out /* [N,2] */, float* offset_out /* [N,2] */, pcab_stream_t stream);
/* the same head on the tensor cores (tcgen05, 3xTF32; csrc/mlp_tc.cu).  which = 0: floats of w1_tc = positional_encoding[2]
 * as [hi 64 rows; lo 64 rows][32]; which = 1: floats of w_tc = [final_proj; mos_seg[0]; offset_head[0]] x [hi 128 rows; lo 128
 * rows][128] (rows = output channels, K-major, hi = weight rounded to tf32, lo = weight - hi). */
size_t pcab_stpn_head_tc_pack_floats(int which);
int pcab_stpn_head_tc(const float* mos_feats_nhwc, int feats_fmt, int H, int W, const float* transformed_points, const int* point_batch,
                      const int* fg_idx, int n_fg, const float* weight_pack_host /* HOST copy of the pcab_stpn_head pack */,
                      const float* w1_tc, const float* w_tc,
                      float x_abs, float y_abs, float* mos_out, float* offset_out, pcab_stream_t stream);

/* ---- clustering: models/cluster.py:9-110 (sparse_quantize + sklearn DBSCAN + canonicalise) ---------------- */
int pcab_dynamic_flags(const float* mos, int n0, int n, int* flags, pcab_stream_t stream);
size_t pcab_cluster_workspace(int n_selected);
int pcab_cluster_scene(const float* transformed_points, const float* offset, const int* sel, int n0, int n_selected,
                       float dedupe_voxel, double eps, int min_samples, int min_p_cluster, long long* inst_out,
                       int* n_instances_out, void* workspace, size_t workspace_bytes, pcab_stream_t stream);

/* ---- TubeNet: models/tpointnet.py:211-305, toolbox/register_utils.py:72-93 -------------------------------- */
/* models/alignnet.py:115-163,201-225: relabel non-empty instances, pad instances without anchor-frame rows, order rows */
int pcab_tpn_relabel(const long long* inst, const long long* tidx, int n, int K0, int T, int* frame_count /* [K0*T] */,
                     int* mapping /* [K0] */, int* pad_frame /* [K0] */, int* totals /* {K, P} */, pcab_stream_t stream);
size_t pcab_tpn_rows_workspace(int n_rows);
int pcab_tpn_rows(const long long* inst, const long long* tidx, int n, int n_pad_rows, int T, const int* mapping,
                  const int* pad_frame, const float* pts_rec, int* seg_rows /* [n] */, long long* inst_new /* [n] */,
                  int* row_src, int* row_inst, int* row_time, int* row_seg, float* row_pts /* [n+P,...] */,
                  void* workspace, size_t workspace_bytes, pcab_stream_t stream);
int pcab_tpn_static_embed(const float* mos_feat /* [n_src,64] */, const float* geo_feat /* [n_src,32] */,
                          const int* src_idx /* [n] row of each (padded) point */, const int* inst, int n, int K,
                          const float* pack_motion, const float* pack_geo, float* mos_emb /* [K,128] */,
                          float* geo_emb /* [K,128] */, pcab_stream_t stream);
size_t pcab_tpn_iteration_workspace(int K, int T);
int pcab_tpn_iteration(const float* points, const int* inst, const int* tidx, int n, int K, int T, const float* mos_emb,
                       const float* geo_emb, const float* pack_pos, const float* pack_regressor,
                       float* pose_out /* [K*T,4,4] */, float* pose_centered_out /* or NULL */,
                       float* rep_out /* [K*T,7] or NULL */, void* workspace, size_t workspace_bytes,
                       const float* pos_w0_tc /* NULL = FP32 positional embedding; else the tensor-core packs of its layers */,
                       const float* pos_w1_tc, const float* pos_bias_host /* HOST: b1[64] b2[128] */,
                       void* pos_scratch /* n * (32 floats + 1 int) */, pcab_stream_t stream);
/* TubeNet embeddings on the tensor cores (csrc/mlp_tc.cu): MLP + max over the rows of each segment.  which: 0 = motion_embed
 * (64-64-128-128), 1 = geo_embed (32-32-64-128), 2 = layers 1-2 of pos_embed (32-64-128) over the rows of pcab_tpn_pos_l0.
 * w*_tc: per layer [hi rows; lo rows] K-major; bias_host: the layers' biases back to back (HOST memory); seg ascending. */
int pcab_embed_segmax_tc(int which, const float* feat, const int* src_idx /* or NULL */, const int* seg, int n, int n_seg,
                         const float* w0_tc, const float* w1_tc, const float* w2_tc, const float* bias_host,
                         float* out /* [n_seg,128] */, pcab_stream_t stream);
int pcab_tpn_pos_l0(const float* points, const int* inst, const int* tidx, int n, int T, const double* frame_sums,
                    const float* pack_pos, float* rows /* [n,32] */, int* seg /* [n] */, pcab_stream_t stream);
int pcab_apply_seg_pose(const float* points, const int* seg, const float* pose, int n, float* out,
                        pcab_stream_t stream);
int pcab_scatter_rows3(const float* src, const int* idx, int k, float* dst, pcab_stream_t stream);
/* models/motionnet.py:246-258: the TubeNet inputs of the selected points (labels, batch / time index, coordinates, motion labels) */
int pcab_tpn_gather(const int* idx, int k, const long long* inst, const int* point_batch, const int* point_time,
                    const float* points, const long long* sd_labels, long long* inst_out, long long* batch_out,
                    long long* time_out, float* points_out, long long* mos_out, pcab_stream_t stream);
/* models/alignnet.py:9-38 (test mode): out[t] = pose_gt[t] @ inv(pose_est[t]) */
int pcab_pose_error(const float* pose_gt, const float* pose_est, int T, float* out /* [T,4,4] */, pcab_stream_t stream);
/* models/alignnet.py:271-279: out2 = {inst_l2_error, dynamic_inst_l2_error}; scratch4 = four doubles */
int pcab_inst_errors(const float* rec_est, const float* rec_gt, const long long* time_idx, const long long* mos_labels, int n,
                     double* scratch4, float* out2, pcab_stream_t stream);

/* ---- data front-end: libs/dataset.py:163-207 steps 2-4 (crop, ground removal) ------------- */
size_t pcab_prep_points_workspace(int n_points);
int pcab_prep_points(const float* raw_points /* [N,3] */, const long long* time_idx, const long long* sd_labels,
                     const long long* fb_labels, const long long* inst_labels, int n_points, float crop_xy, float crop_z_min,
                     float crop_z_max, int remove_ground, float ground_height /* incl. slack */, float* points4_out /* [N,4] */,
                     int* time_out, long long* sd_out, long long* fb_out, long long* inst_out, int* count_out /* device */,
                     void* workspace, size_t workspace_bytes, pcab_stream_t stream);
/* the same with step 1 (training-time augmentation, libs/dataset.py:90-113,167-171) applied first, in float64 like numpy:
 * p' = ((R p + t) + (u - 0.5) * noise_amp) * scale.  tsfm16: HOST [4,4] float64; noise: DEVICE [n,3] float64 uniforms drawn by
 * the host in the reference's order, or NULL for the device generator seeded with `seed`. */
int pcab_prep_points_augmented(const float* raw_points, const long long* time_idx, const long long* sd_labels,
                               const long long* fb_labels, const long long* inst_labels, int n, const double* tsfm16,
                               const double* noise, unsigned long long seed, double noise_amp, double scale, double crop_xy,
                               double crop_z_min, double crop_z_max, int remove_ground, double ground_height, float* points4_out,
                               int* time_out, long long* sd_out, long long* fb_out, long long* inst_out, int* count_out,
                               void* workspace, size_t workspace_bytes, pcab_stream_t stream);

/* ---- evaluation tail: libs/tester.py:58-88, toolbox/register_utils.py:59-93, toolbox/sf_eval_utils.py:46-52,71-100,
 *      libs/loss.py:17-48,139-149 ------------------------------------------------------------ */
int pcab_flow_eval(const float* input_points, const int* time_idx, const float* rec_est, const float* ego_motion_gt /* [T,4,4] */,
                   const long long* inst_labels, const float* inst_motion_gt /* [K,T,4,4] */, int n_instances,
                   const long long* fb_labels, const long long* sd_labels, const float* mos_est /* [N,2] */,
                   const long long* fb_est_per_point, int n_points, int n_frames, float* epe_out /* [N] */,
                   float* rel_out /* [N] */, double* sf_counters /* [3][6], accumulated */,
                   long long* mos_counters /* [8], accumulated */, pcab_stream_t stream);
/* instance-segmentation scores of toolbox/cluster_eval.py:71-152 for one scene; counters[28] doubles, accumulated:
 * class c: [4c] sum of mean coverage, [4c+1] sum of weighted coverage, [4c+2] scenes with gt instances, [4c+3] gt instances;
 * threshold k (0.5 .. 0.9), class c: tp at [8 + 2(2k+c)], fp at [8 + 2(2k+c) + 1] */
size_t pcab_cluster_eval_workspace(int max_est, int max_gt);
int pcab_cluster_eval(const long long* inst_est, const long long* inst_gt, const long long* mos_label, int n_points, int max_est,
                      int max_gt, double* counters, void* workspace, size_t workspace_bytes, pcab_stream_t stream);

/* ---- Chamfer distance: chamfer_distance/chamfer_distance.cpp:27-56 (forward_cuda / backward_cuda) --------- */
size_t pcab_chamfer_workspace(int B, int n, int m);
int pcab_chamfer_forward(const float* xyz1, const float* xyz2, int B, int n, int m, float* dist1, float* dist2,
                         int* idx1, int* idx2, void* workspace, size_t workspace_bytes, pcab_stream_t stream);
int pcab_chamfer_backward(const float* xyz1, const float* xyz2, int B, int n, int m, const float* grad_dist1,
                          const float* grad_dist2, const int* idx1, const int* idx2, float* grad_xyz1,
                          float* grad_xyz2, pcab_stream_t stream);

/* same results with every pair evaluated (the formulation of chamfer_distance.cu:6-136); workspace as pcab_chamfer_workspace */
int pcab_chamfer_forward_brute(const float* xyz1, const float* xyz2, int B, int n, int m, float* dist1, float* dist2,
                               int* idx1, int* idx2, void* workspace, size_t workspace_bytes, pcab_stream_t stream);
/* One direction of that search (chamfer_distance.cpp:59-84, nnsearch) through the exact uniform-grid search: dist / idx of the
 * nearest target of every query, bit-identical to the brute force (lowest index wins ties).  max_dist > 0 bounds the search
 * (the KD-tree hybrid search of Open3D's registration, models/egomotion.py:21): dist = NaN, idx = -1 where no target lies
 * strictly within max_dist; max_dist <= 0 = unbounded. */
size_t pcab_nn_workspace(int n_queries, int n_targets);
int pcab_nn_search(const float* queries, int n_queries, const float* targets, int n_targets, float max_dist, float* dist,
                   int* idx, void* workspace, size_t workspace_bytes, pcab_stream_t stream);

/* ---- ICP refinement: models/egomotion.py:9-28,360-384 (model.ego_icp) and
 *      models/alignnet.py:54-112 (model.tpointnet_icp), i.e. open3d registration_icp(src, tgt, max_dist, init,
 *      TransformationEstimationPointToPoint(), ICPConvergenceCriteria(max_iteration)) ------------------------------------
 * n_problems independent registrations in one call.  src_problem[i] (NULL = all 0; < 0 = ignored) assigns source point i to
 * a problem; tgt_group[j] / problem_group[p] (both NULL = one group) restrict problem p to the targets of its group.
 * init_pose [P,4,4] (NULL = identity) is the starting transformation; pose_out [P,4,4] = total transformation (update_k ...
 * update_1 init); stats [P,3] = fitness, inlier rmse, updates applied (may be NULL).  rel_fitness / rel_rmse = Open3D's
 * ICPConvergenceCriteria defaults 1e-6. */
size_t pcab_icp_workspace(int n_targets, int n_problems);
int pcab_icp_point_to_point(const float* src, const int* src_problem, int n_src, const float* tgt, const int* tgt_group, int n_tgt,
                            const int* problem_group, int n_problems, const float* init_pose, float max_dist, int max_iter,
                            float rel_fitness, float rel_rmse, float* pose_out, float* stats, void* workspace,
                            size_t workspace_bytes, pcab_stream_t stream);

/* ---- FuseLoss forward + gradients w.r.t. the network outputs: libs/loss.py:52-320,
 *      libs/lovasz_softmax.py:56-94, libs/outlier_loss.py:15-29 ------------------------------------------------------------ */
/* get_seg_loss (libs/loss.py:113-136) for the FG/BG map (get_fb_loss :165-186) or the motion logits (get_mos_loss :139-163).
 * logits: [.., 2, hw] float (hw = Ny*Nx for fb_seg_est [B,T,2,Ny,Nx]; hw = 1 for mos_est [N,2]); gt: int64 per item; an item
 * takes part when sel_float[i] == 1 or sel_a[i] == 1 or sel_b[i] == 1 (NULL arrays are skipped).
 * out13 = {weighted CE, Lovasz-softmax, intersection[2], union[2], pred_positives[2], gt_positives[2] (raw counts: compute_iou
 * :17-48 divides by 1e3), n_selected, class weights[2]}.  The workspace keeps what pcab_seg_loss_grad needs. */
size_t pcab_seg_loss_workspace(long long n_items);
int pcab_seg_loss(const float* logits, int hw, const long long* gt, const float* sel_float, const long long* sel_a,
                  const long long* sel_b, long long n_items, float* out13, void* workspace, size_t workspace_bytes,
                  pcab_stream_t stream);
/* grad (layout of logits) = d(w_ce * CE + w_lovasz * Lovasz) / d logits; same arguments and workspace as the forward call */
int pcab_seg_loss_grad(const float* logits, int hw, const long long* gt, const float* sel_float, const long long* sel_a,
                       const long long* sel_b, long long n_items, const float* out13, float w_ce, float w_lovasz, float* grad,
                       void* workspace, size_t workspace_bytes, pcab_stream_t stream);
/* get_offset_loss (libs/loss.py:189-245).  inst_motion_gt = the per-scene [K_b,T,4,4] lists concatenated, inst_offset[b] = first
 * row of scene b.  out4 = {offset_norm_loss, offset_dir_loss, offset_l2_error, n_foreground}; gt_offset [N,2] may be NULL. */
size_t pcab_offset_loss_workspace(int n_instances_total);
int pcab_offset_loss(const float* points, const int* point_batch, const int* point_time, const long long* inst_labels,
                     const long long* fb_labels, const float* ego_motion_gt, const float* inst_motion_gt, const int* inst_offset,
                     int n_instances_total, int T, const float* transformed_points, const float* offset_est, long long n_points,
                     float* gt_offset, float* out4, void* workspace, size_t workspace_bytes, pcab_stream_t stream);
int pcab_offset_loss_grad(const long long* fb_labels, const float* gt_offset, const float* offset_est, long long n_points,
                          const float* out4, float w_norm, float w_dir, float* grad /* [N,2] */, pcab_stream_t stream);
/* OutlierLoss (libs/outlier_loss.py:15-29) over n_mats contiguous [m,m] soft-assignment matrices; scratch1 = one double */
int pcab_perm_loss(const float* perm, int n_mats, int m, double* scratch1, float* out1, pcab_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* PCAB200_H_ */
