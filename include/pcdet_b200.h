/*
 * pcdet_b200.h -- C ABI of libpcdet_b200.so, the B200 (sm_100a) implementation of PCDet's voxel hot path.
 *
 * Every entry point takes plain device pointers, sizes and a CUDA stream (passed as void*, i.e. a
 * cudaStream_t), allocates nothing, launches on the given stream only, never synchronises the host,
 * and returns a status code (0 = ok, see PCDB_* below) instead of calling exit() like the
 * reference's gpuAssert (pcdet/ops/iou3d_nms/src/iou3d_nms.cpp:19-27).  Scratch memory comes from
 * the caller through (workspace, workspace_bytes); the matching *_workspace_bytes() call sizes it.
 *
 * Sizes that are only known on the device (number of voxels, number of active output sites) are
 * written to int32 device scalars; every kernel that consumes them accepts either the host value
 * or a device pointer (`*_dev`, may be NULL) and bounds its grid by the capacity argument, so a
 * whole forward pass can run -- and be captured in a CUDA graph -- without a device->host copy.
 *
 * Each function names the reference interface it replaces (paths relative to the PCDet tree;
 * "spconv" = traveller59/spconv v1.0 @ 8da6f96, the external dependency the reference binds).
 *
 * Size limits (every one is checked and reported as PCDB_KEY_OVERFLOW / PCDB_UNSUPPORTED, never silently wrapped):
 *   - hash keys are 32-bit linear cell indices: batch * Z * Y * X of the voxel grid (pcdb_voxelize) and of the input /
 *     output level of a rulebook (pcdb_rulebook_*) must stay below 2^32 - 1.  KITTI (41 x 1600 x 1408 = 92.4 M cells per
 *     sample) therefore allows a batch of 46 per call, nuScenes (41 x 1024 x 1024) of 99; split larger batches.  (The
 *     reference's dense grid of int32 indices has the limit batch * volume < 2^31, half of this.)
 *   - a strided rulebook build packs (input row, kernel offset) into 32 bits: n_in * K < 2^32 - 1, K <= 32; its table
 *     holds at most 2^26 slots (n_out_cap <= 2^25);
 *   - pcdb_rulebook_chain: levels >= 1 are limited to 2^31 cells (one 64-bit occupancy word per 32 cells);
 *   - the tcgen05 convolution takes c_in in {16, 32, 64}, c_out in {16, 32, 64, 128}, K <= 27; other shapes run on the
 *     FMA-pipe kernel.
 */
#ifndef PCDET_B200_H_
#define PCDET_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PCDB_OK 0
#define PCDB_INVALID_ARGUMENT 1
#define PCDB_WORKSPACE_TOO_SMALL 2
#define PCDB_KEY_OVERFLOW 3
#define PCDB_CUDA_ERROR 4
#define PCDB_UNSUPPORTED 5

/* feature / weight storage types */
#define PCDB_F32 0
#define PCDB_BF16 1

/* flags for pcdb_sparse_conv_fwd */
#define PCDB_EPI_RELU 1
#define PCDB_WEIGHT_PACKED 2
/* weight was produced by pcdb_pack_conv_weights (tensor-core operand image) */
/* Programmatic dependent launch of the tcgen05 kernel: its set-up (barriers, TMEM, rulebook slice) may overlap the
 * tail of the kernel launched just before it on the same stream.  Only valid when `features` (and whoever still
 * reads `out`) is ALL that kernel has to do with this call: nbr, n_out_dev, weight, scale, shift and bias must come
 * from work that completed earlier (another stream joined by an event, or an earlier kernel). */
#define PCDB_CONV_PDL 4
#define PCDB_CONV_SHALLOW_RING 8   /* flag of pcdb_sparse_conv_fwd: two-stage shared-memory ring (small footprint), for
                                    * deployments that keep several steps in flight on one GPU */
#define PCDB_CONV_ROWS_HINT(rows) ((int)(rows) << 8)   /* or'ed into `algo`, see pcdb_sparse_conv_fwd */

int pcdb_abi_version(void);
/* Message describing the last non-zero status returned on this thread. */
const char *pcdb_last_error(void);

/* ---------------------------------------------------------------------------------------------
 * Voxelisation + mean VFE.
 * Replaces spconv.utils.VoxelGenerator.generate -> points_to_voxel_3d_np (called from
 * pcdet/datasets/dataset.py:163), the per-frame collate of pcdet/datasets/dataset.py:266-299 and
 * MeanVoxelFeatureExtractor.forward (pcdet/models/vfe/vfe_utils.py:26-34).
 *
 * points:        (n_points, n_feat) f32, frames concatenated; columns 0..2 are x,y,z
 * frame_offsets: (batch+1) i32 device array, frame b owns points [off[b], off[b+1])
 * grid_xyz:      round((range[3:]-range[:3]) / voxel_size), x,y,z
 * overflow_break: 1 = a frame stops at the first point that would open voxel #max_voxels (spconv
 *                v1.0), 0 = that point is skipped and later points still join existing voxels (v1.1+)
 * Outputs (capacity rows = min(n_points, batch*max_voxels)):
 *   voxels      (cap, max_points, n_feat) f32 zero padded          -- may be NULL
 *   coords      (cap, 4) i32 [b, z, y, x]
 *   num_points  (cap) i32
 *   mean        (cap, mean_stride) in mean_dtype: sum over points / num_points in the first n_feat
 *               columns, zeros in the rest (mean_stride >= n_feat)  -- may be NULL
 *   point_idx   (cap, max_points) i32 index into `points`, -1 padded -- may be NULL
 *   voxel_offsets (batch+1) i32: frame b owns voxel rows [vo[b], vo[b+1]); vo[batch] = total
 * Voxel order inside a frame is first appearance in point order and each voxel keeps the
 * max_points points with the smallest indices, exactly like the serial reference loop.
 * ------------------------------------------------------------------------------------------- */
size_t pcdb_voxelize_workspace_bytes(int n_points, int batch, int max_points, int max_voxels);
int pcdb_voxelize(const float *points, int n_points, int n_feat, const int32_t *frame_offsets, int batch,
                  const float *voxel_size_xyz, const float *range_xyzxyz, const int32_t *grid_xyz,
                  int max_points, int max_voxels, int overflow_break,
                  float *voxels, int32_t *coords, int32_t *num_points, void *mean, int mean_dtype,
                  int mean_stride, int32_t *point_idx, int32_t *voxel_offsets,
                  void *workspace, size_t workspace_bytes, void *stream);

/* The same call in two halves for callers that overlap them (pcdet_b200/pipeline.py), identical arguments:
 * `_sites` = hash, voxel ranks, `coords` and `voxel_offsets` -- all the rulebook builds need;
 * `_points` = point assignment, `voxels`, `num_points`, `mean`.  pcdb_voxelize == _sites then _points. */
int pcdb_voxelize_sites(const float *points, int n_points, int n_feat, const int32_t *frame_offsets, int batch,
                  const float *voxel_size_xyz, const float *range_xyzxyz, const int32_t *grid_xyz,
                  int max_points, int max_voxels, int overflow_break,
                  float *voxels, int32_t *coords, int32_t *num_points, void *mean, int mean_dtype,
                  int mean_stride, int32_t *point_idx, int32_t *voxel_offsets,
                  void *workspace, size_t workspace_bytes, void *stream);
int pcdb_voxelize_points(const float *points, int n_points, int n_feat, const int32_t *frame_offsets, int batch,
                  const float *voxel_size_xyz, const float *range_xyzxyz, const int32_t *grid_xyz,
                  int max_points, int max_voxels, int overflow_break,
                  float *voxels, int32_t *coords, int32_t *num_points, void *mean, int mean_dtype,
                  int mean_stride, int32_t *point_idx, int32_t *voxel_offsets,
                  void *workspace, size_t workspace_bytes, void *stream);

/* MeanVoxelFeatureExtractor.forward on an already voxelised batch (vfe_utils.py:26-34). */
int pcdb_vfe_mean(const float *voxels, const int32_t *num_points, int n_voxels, int max_points,
                  int n_feat, void *mean, int mean_dtype, int mean_stride, void *stream);

/* Ingest filters in front of the voxelizer, for raw clouds that are already on the device: KittiDataset.__getitem__
 * FOV_POINTS_ONLY (pcdet/datasets/kitti/kitti_dataset.py:714-717: Calibration.lidar_to_rect + rect_to_img,
 * pcdet/utils/calibration.py:66-85, and get_fov_flag, kitti_dataset.py:236-253) and mask_points_by_range
 * (pcdet/utils/common_utils.py:47-51, called at pcdet/datasets/dataset.py:184).
 *   points (n, c) f32, frames concatenated; frame_offsets (batch+1) i32 on the device; batch <= 64;
 *   calib: NULL or batch records of 26 f32 = [v2r (4,3) row-major with [x y z 1].v2r = rectified camera coordinates
 *          (= V2C^T . R0^T) | P2 (3,4) row-major | img_h, img_w]: keep the points that project inside the image with
 *          depth >= 0;
 *   range_xy: NULL or [x_min, y_min, x_max, y_max] (device): keep x_min <= x <= x_max and y_min <= y <= y_max.
 * Surviving points are written in their original order to out_points (capacity n rows), the new frame boundaries to
 * out_offsets (batch+1), and, if out_index != NULL, their original row numbers to out_index.  No host round trip:
 * out_points / out_offsets go straight into pcdb_voxelize. */
size_t pcdb_filter_points_workspace_bytes(int n_points);
int pcdb_filter_points(const float *points, int n, int c, const int32_t *frame_offsets, int batch,
                       const float *calib, const float *range_xy, float *out_points, int32_t *out_offsets,
                       int32_t *out_index, void *workspace, size_t workspace_bytes, void *stream);

/* PointPillars: PillarFeatureNetOld2.forward (vfe_utils.py:168-215; one PFNLayer, vfe_utils.py:61-116, BatchNorm in
 * eval mode folded into scale/shift) fused with PointPillarsScatter.forward (pcdet/models/rpn/pillar_scatter.py:23-55).
 * voxels (n, max_points, n_feat) f32 zero padded, num_points (n), coords (n,4) [b,z,y,x]; center_offset_xyz =
 * voxel_size / 2 + range_min evaluated in double (vfe_utils.py:162-164); weight (n_filters,
 * n_feat + 6 [+1 with_distance]) = PFNLayer.linear.weight; out_features (n, n_filters) and/or canvas
 * (batch, n_filters * nz, ny, nx) f32 (either may be NULL; the canvas is cleared here). */
int pcdb_pillar_vfe(const float *voxels, const int32_t *num_points, const int32_t *coords, int n,
                    const int32_t *n_dev, int max_points, int n_feat, const float *voxel_size_xyz,
                    const float *center_offset_xyz, int with_distance, const float *weight, int n_filters,
                    const float *scale, const float *shift, float *out_features, float *canvas, int batch,
                    const int32_t *canvas_shape_zyx, void *stream);

/* ---------------------------------------------------------------------------------------------
 * Rulebook construction.  Replaces spconv.ops.get_indice_pairs -> getIndicePair<3> (called from
 * spconv.conv.SparseConvolution.forward for every new indice_key; 8 builds per BackBone8x forward,
 * pcdet/models/rpn/rpn_backbone.py:12-51).
 *
 * The rulebook is emitted as a neighbour map: nbr[k * ld + o] = input row feeding output row o
 * through kernel offset k (row-major over kz,ky,kx), or -1.  The reference's
 * (indicePairs, indiceNum) form is recoverable from it (pcdet_b200/spconv/ops.py) and both list the
 * same (input, output) pairs per offset.
 *
 * indices (n,4) i32 [b,z,y,x]; n_dev (optional) overrides n with a device-side count <= n.
 * ------------------------------------------------------------------------------------------- */
size_t pcdb_rulebook_workspace_bytes(int n_in_cap, int kernel_volume, int n_out_cap);

/* Submanifold convolution (stride 1, padding k/2 forced as in spconv; outputs == inputs). */
int pcdb_rulebook_subm(const int32_t *indices, int n, const int32_t *n_dev, int batch,
                       const int32_t *spatial_shape_zyx, const int32_t *ksize_zyx,
                       const int32_t *dilation_zyx, int32_t *nbr, int ld,
                       void *workspace, size_t workspace_bytes, void *stream);

/* Same rulebook, for a level whose sites are the OUTPUT sites of an earlier pcdb_rulebook_conv call (every
 * SubM block of BackBone8x after the first: rpn_backbone.py:22-45): the hash table that call left in its
 * workspace (site -> output row) is looked up directly, nothing is inserted.  conv_workspace and the three
 * conv_* sizes are the workspace pointer and the (n_in_cap, kernel_volume, n_out_cap) of that call, whose
 * out_shape must equal spatial_shape_zyx and whose out_indices must be `indices`.  spconv v1.0 has no such
 * entry: it re-inserts the sites into a fresh dense grid for every rulebook. */
int pcdb_rulebook_subm_reuse(const int32_t *indices, int n, const int32_t *n_dev, int batch,
                             const int32_t *spatial_shape_zyx, const int32_t *ksize_zyx,
                             const int32_t *dilation_zyx, int32_t *nbr, int ld,
                             const void *conv_workspace, int conv_n_in_cap, int conv_kernel_volume,
                             int conv_n_out_cap, int flags, void *stream);

/* flags of pcdb_rulebook_subm_reuse / _conv_sites / _conv_pairs: the memsets these calls start with (hash table,
 * owner masks, nbr = -1) have already been done by the caller -- pcdb_rulebook_conv_clear for a strided build,
 * a fill of nbr with -1 for _subm_reuse -- at a time when nothing was waiting for them (they depend on nothing,
 * but inside the build they sit in front of every kernel of a chain that the convolutions wait for). */
#define PCDB_RB_CLEARED 1
/* pcdb_rulebook_chain together with PCDB_RB_CLEARED: only the neighbour maps were cleared by the caller; the workspace is as
 * phase 8 of the previous build (or pcdb_rulebook_chain_clear) left it, and phase 1 resets the level-0 table itself. */
#define PCDB_RB_UNDONE 2
int pcdb_rulebook_conv_clear(void *workspace, size_t workspace_bytes, int n_in_cap, int kernel_volume,
                             int n_out_cap, int32_t *nbr_fwd, int ld_out, void *stream);
/* dst[k*ld + r] = value for k < n_maps and r < min(*rows_dev, rows_cap): clears the rows of a neighbour map that the
 * previous build wrote, instead of the whole capacity.  A map that starts all -1 and whose extent (the row count of
 * the build that filled it, still in its device counter) is cleared before every build stays -1 behind the extent. */
int pcdb_fill_rows_i32(int32_t *dst, int ld, int n_maps, const int32_t *rows_dev, int rows_cap, int value, void *stream);

/* The input-stationary reading of an output-stationary map: nbr_inv[k*ld_in + i] = o for every pair nbr[k*ld_out + o] = i,
 * -1 elsewhere in rows [0, n_in).  It is what the input gradient of a strided convolution gathers by (spconv's
 * indiceConvBackward walks the same pairs); pcdb_rulebook_conv can emit it directly, this call derives it
 * for maps that pcdb_rulebook_chain built.  n_out_dev / n_in_dev (optional) override the row counts on the device. */
int pcdb_rulebook_invert(const int32_t *nbr, int ld_out, int kernel_volume, int n_out, const int32_t *n_out_dev,
                         int32_t *nbr_inv, int ld_in, int n_in, const int32_t *n_in_dev, void *stream);

/* Regular (strided) sparse convolution.  out_indices (n_out_cap,4) i32 in first-touch order of the
 * serial reference loop (input row ascending, then kernel offset ascending); n_out_dev receives the
 * count (clamped to n_out_cap; overflow sets status flag word n_out_dev[1] = 1).
 * nbr_fwd (K, ld_out): output-stationary map; nbr_inv (K, ld_in), optional: for input row i the
 * output row it feeds through offset k (used by SparseInverseConv3d and the backward pass). */
int pcdb_rulebook_conv(const int32_t *indices, int n, const int32_t *n_dev, int batch,
                       const int32_t *spatial_shape_zyx, const int32_t *out_shape_zyx,
                       const int32_t *ksize_zyx, const int32_t *stride_zyx, const int32_t *padding_zyx,
                       const int32_t *dilation_zyx, int32_t *out_indices, int n_out_cap,
                       int32_t *n_out_dev, int32_t *nbr_fwd, int ld_out, int32_t *nbr_inv, int ld_in,
                       void *workspace, size_t workspace_bytes, void *stream);

/* The same build in two halves, so that a caller can overlap them (pcdet_b200/pipeline.py): `_sites` numbers the
 * output sites (out_indices, n_out_dev, the site table in the workspace) -- all the next level needs -- and
 * `_pairs` then emits nbr_fwd / nbr_inv from the workspace `_sites` filled (same n, n_dev, kernel size, stride,
 * dilation and n_out_cap).  pcdb_rulebook_conv == _sites followed by _pairs on one stream. */
int pcdb_rulebook_conv_sites(const int32_t *indices, int n, const int32_t *n_dev, int batch,
                             const int32_t *spatial_shape_zyx, const int32_t *out_shape_zyx,
                             const int32_t *ksize_zyx, const int32_t *stride_zyx, const int32_t *padding_zyx,
                             const int32_t *dilation_zyx, int32_t *out_indices, int n_out_cap,
                             int32_t *n_out_dev, void *workspace, size_t workspace_bytes, int flags, void *stream);
int pcdb_rulebook_conv_pairs(int n, const int32_t *n_dev, const int32_t *ksize_zyx, const int32_t *stride_zyx,
                             const int32_t *dilation_zyx, int n_out_cap, int32_t *nbr_fwd, int ld_out,
                             int32_t *nbr_inv, int ld_in, const void *workspace, int flags, void *stream);

/* ---------------------------------------------------------------------------------------------
 * Every rulebook of a strided backbone at once (csrc/rulebook_chain.cu): level 0 = the caller's active sites, level l >= 1 =
 * output sites of strided conv l applied to level l-1 (dilation 1, kernel >= stride), plus an optional centred SubM map
 * per level.  Replaces the eight ops.get_indice_pairs calls of one BackBone8x forward (pcdet/models/rpn/rpn_backbone.py:
 * 54-77; spconv getIndicePair<3>) by four launches with no level-to-level chain.  Same site SETS and pair
 * SETS as pcdb_rulebook_conv / pcdb_rulebook_subm; the rows of levels >= 1 come in ascending (b, z, y, x) order -- the order
 * of the reference's CUDA rulebook (sorted linear index) -- instead of the first-touch order of its CPU loop.
 * Levels >= 1 are limited to batch * volume <= 2^31 cells (8 bytes of workspace per 32 cells).
 *
 *   coords0 (caps[0], 4) [b,z,y,x], n0_dev its device-side row count; shapes_zyx [n_levels][3];
 *   ksize / stride / padding [n_levels-1][3] of conv l = 1 .. n_levels-1; caps [n_levels] row capacities;
 *   coords[l] (caps[l], 4), counts[l] [count, overflow] and nbr_conv[l] (K_l, caps[l]) for l >= 1 (entry 0 unused);
 *   subm_ksize_zyx [n_levels][3] (all zero = no SubM map at that level), nbr_subm[l] (K, caps[l]).
 *   The pointer arrays are HOST arrays of device pointers.  With PCDB_RB_CLEARED the caller has run
 *   pcdb_rulebook_chain_clear on the workspace and filled the maps with -1 (pcdb_fill_rows_i32) beforehand.
 *   rows_hint [n_levels] or NULL: row counts the caller expects (0 = unknown); they size the grids only.
 *   phase (a mask; the parts of one build must be issued in this order, possibly from separate calls so that the caller can
 *          put events between them): 1 = occupancy of every level (needs the cleared workspace), 4 = the SubM map of level 0
 *          (all that the first convolutions wait for; needs that map's cleared rows), 2 = numbering, coordinates, counts
 *          and every other map, 8 = undo: zero the occupancy words and counters this build set, so that the next build
 *          needs no workspace clear (PCDB_RB_UNDONE); 7 = a complete build.  2 | 32 = numbering and only the maps of level 1
 *          (the strided conv into it and its SubM map), 16 = the maps of the levels >= 2 afterwards: a caller whose second
 *          group of convolutions only needs level 1 can let it start before the deeper maps exist.
 * ------------------------------------------------------------------------------------------- */
size_t pcdb_rulebook_chain_workspace_bytes(int batch, int n_levels, const int32_t *shapes_zyx, const int32_t *caps);
/* Everything a build with PCDB_RB_CLEARED expects: the workspace (level-0 table, occupancy words, counters) and -- when the
 * map arguments are given -- the rows of every neighbour map that its previous build wrote (extent = the level's device
 * counter; maps start all -1), in one launch instead of one pcdb_fill_rows_i32 per map.  n0_dev .. nbr_subm as in
 * pcdb_rulebook_chain; pass NULL for them to clear the workspace only, or a NULL workspace to clear the maps only. */
int pcdb_rulebook_chain_clear(void *workspace, size_t workspace_bytes, int batch, int n_levels, const int32_t *shapes_zyx,
                              const int32_t *caps, const int32_t *ksize_zyx, const int32_t *subm_ksize_zyx,
                              const int32_t *n0_dev, int32_t *const *counts, int32_t *const *nbr_conv,
                              int32_t *const *nbr_subm, void *stream);
int pcdb_rulebook_chain(const int32_t *coords0, const int32_t *n0_dev, int batch, int n_levels,
                        const int32_t *shapes_zyx, const int32_t *ksize_zyx, const int32_t *stride_zyx,
                        const int32_t *padding_zyx, const int32_t *caps, int32_t *const *coords,
                        int32_t *const *counts, int32_t *const *nbr_conv, const int32_t *subm_ksize_zyx,
                        int32_t *const *nbr_subm, const int32_t *rows_hint, void *workspace, size_t workspace_bytes,
                        int flags, int phase, void *stream);

/* ---------------------------------------------------------------------------------------------
 * Sparse convolution forward.  Replaces spconv.ops.indice_conv / indice_subm_conv /
 * indice_inverse_conv -> indiceConv<T> (gather -> cuBLAS GEMM -> scatter-add per kernel offset), plus,
 * when scale/shift/flags are given, the eval-mode BatchNorm1d + ReLU that SparseSequential applies
 * next (pcdet/models/rpn/rpn_backbone.py:79-103).
 *
 *   out[o, :] = epilogue( sum_k  features[nbr[k*ld + o], :] @ weight[k] )
 *   epilogue(y) = relu?( (y + bias) * scale + shift )         (scale/shift/bias optional, f32, c_out; the conv bias
 *                                                              precedes the folded BatchNorm, as in spconv + BatchNorm1d)
 *
 * features (n_in, c_in) and out (n_out, c_out) in `dtype` (PCDB_F32 or PCDB_BF16), row-major,
 * contiguous; n_in = rows of the features buffer (capacity is fine).  weight (K, c_in, c_out) in `dtype`,
 * or, with PCDB_WEIGHT_PACKED, the buffer written by pcdb_pack_conv_weights.  Accumulation is fp32.
 * PCDB_F32 runs on the fp32 FMA pipe (<=1e-4 of an fp32 reference).  PCDB_BF16 with packed weights,
 * c_in in {16,32,64} and c_out in {16,32,64,128} runs on the tcgen05 tensor cores with the accumulator
 * in TMEM and TMA gather4 staging; other bf16 shapes use the FMA pipe (unpacked weights).
 * n_out_dev (optional) overrides n_out with a device-side count.
 * algo: low byte 0 = auto (tcgen05 + cp.async gather where eligible), 1 = FMA-pipe kernel, 2 = tcgen05 + TMA
 *       gather4, 3 = tcgen05 + cp.async gather.  Bits 8 and up (PCDB_CONV_ROWS_HINT): the number of output rows
 *       the caller expects when n_out is only a capacity and the count lives in n_out_dev (0 = n_out); it only
 *       tunes the kernel's shared-memory ring depth / CTAs per SM, never the result.
 * ------------------------------------------------------------------------------------------- */
int pcdb_sparse_conv_fwd(const void *features, int n_in, const void *weight, const int32_t *nbr, int ld,
                         int kernel_volume, int n_out, const int32_t *n_out_dev, int c_in, int c_out,
                         int dtype, const float *scale, const float *shift, const float *bias,
                         int flags, void *out, int algo, void *stream);
/* The same with a residual row per output row (tcgen05 path only: bf16, packed weights): residual (n_out, c_out) bf16 is
 * added to the accumulator BEFORE the epilogue's scale / shift -- a convolution over more input channels than one launch
 * takes runs as two launches over channel halves, the first without epilogue, the second with the first's output as
 * residual (UNetV2's 128 -> 64 merge convolutions, rpn_unet.py:414-422) -- or, with PCDB_EPI_RESIDUAL_POST, AFTER it and
 * before the ReLU (the shortcut of SparseBasicBlock, resnet_utils.py:17-48). */
#define PCDB_EPI_RESIDUAL_POST 16
int pcdb_sparse_conv_fwd_ex(const void *features, int n_in, const void *weight, const int32_t *nbr, int ld,
                            int kernel_volume, int n_out, const int32_t *n_out_dev, int c_in, int c_out,
                            int dtype, const float *scale, const float *shift, const float *bias,
                            const void *residual, int flags, void *out, int algo, void *stream);


/* Tensor-core operand image of a bf16 (K, c_in, c_out) weight: per kernel offset the (c_out x c_in)
 * K-major tile in the shared-memory swizzle the MMA reads, so the kernel stages it with one bulk copy.
 * pcdb_conv_packed_weight_bytes returns 0 for shapes the tcgen05 kernels do not take. */
size_t pcdb_conv_packed_weight_bytes(int kernel_volume, int c_in, int c_out);
int pcdb_pack_conv_weights(const void *weight, int kernel_volume, int c_in, int c_out, void *packed, void *stream);

/* The same image from fp32 or bf16 weights (`dtype`), optionally for the INPUT-GRADIENT convolution of a layer:
 * c_in / c_out are those of the convolution the image is for; with PCDB_PACK_TRANSPOSE `weight` is the forward layer's
 * (K, c_out, c_in) parameter and the image holds W[k]^T; PCDB_PACK_FLIP reverses the offsets (k -> K-1-k), which is how
 * the rulebook of a centred submanifold convolution reads the other way round. */
#define PCDB_PACK_TRANSPOSE 1
#define PCDB_PACK_FLIP 2
int pcdb_pack_conv_weights_ex(const void *weight, int dtype, int kernel_volume, int c_in, int c_out, int flags,
                              void *packed, void *stream);

/* Backward of the above without epilogue (spconv indiceConvBackward), fp32 only:
 *   grad_features[i,:] += sum over (k,o) with nbr[k*ld+o]==i of grad_out[o,:] @ weight[k]^T
 *   grad_weight[k]     += sum over o of features[nbr[k*ld+o],:]^T (x) grad_out[o,:]
 * grad_features (n_in,c_in) and grad_weight (K,c_in,c_out) must be zeroed by the caller. */
int pcdb_sparse_conv_bwd(const float *features, const float *weight, const float *grad_out,
                         const int32_t *nbr, int ld, int kernel_volume, int n_in, int n_out,
                         int c_in, int c_out, float *grad_features, float *grad_weight, void *stream);

/* ---------------------------------------------------------------------------------------------
 * Mixed-precision training path on the tensor cores (tools/train.py:119-122 runs the backbone of
 * pcdet/models/rpn/rpn_backbone.py:79-103 in train mode).  bf16 activations and gradients, fp32 accumulation,
 * fp32 master weights and weight gradients.
 *
 *   forward          pcdb_sparse_conv_fwd (bf16, weights packed from the fp32 parameter by pcdb_pack_conv_weights_ex)
 *   input gradient   pcdb_sparse_conv_fwd again: grad_out as features, the PCDB_PACK_TRANSPOSE image, the rulebook read
 *                    the other way round (nbr_inv of a strided conv; nbr with PCDB_PACK_FLIP for a centred SubM conv)
 *   weight gradient  pcdb_sparse_conv_wgrad (replaces the per-offset cuBLAS GEMM of spconv indiceConvBackward):
 *                      grad_weight[k][ci][co] (+)= sum_o features[nbr[k*ld + o]][ci] * grad_out[o][co]
 *                    one tcgen05 contraction per layer over the output rows, both operands MN-major in shared memory,
 *                    per-CTA partial sums in `workspace` added in index order (no atomics, bit-reproducible).
 *                    features (n_in, c_in), grad_out (n_out, c_out) bf16; grad_weight (K, c_in, c_out) fp32, overwritten
 *                    unless `accumulate`; c_in in {16,32,64}, c_out in {16,32,64,128}, K <= 27.
 * ------------------------------------------------------------------------------------------- */
size_t pcdb_sparse_conv_wgrad_workspace_bytes(int kernel_volume, int n_out, int c_in, int c_out);
int pcdb_sparse_conv_wgrad(const void *features, int n_in, const void *grad_out, const int32_t *nbr, int ld,
                           int kernel_volume, int n_out, const int32_t *n_out_dev, int c_in, int c_out,
                           float *grad_weight, int accumulate, void *workspace, size_t workspace_bytes, void *stream);

/* Train-mode BatchNorm1d (+ ReLU when flags has PCDB_EPI_RELU) over the n rows of a sparse tensor, torch.nn.BatchNorm1d
 * semantics (batch statistics, running_mean / running_var updated with `momentum`, unbiased variance in the running
 * estimate).  y, out, grad_out, grad_y: (n, c) in `dtype`, c a multiple of 8 up to 128; gamma / beta / running_* /
 * grad_gamma / grad_beta: fp32 (c), optional.  stats (4, c) fp32 is written by the forward call (mean, 1/std, folded
 * scale, folded shift) and read by the backward call.  out == NULL computes the statistics only.  conv_partials
 * (optional): n_conv_partials blocks of (2, c) per-tile channel sums / sums of squares written by a convolution epilogue
 * instead of a pass over y.  n_dev (optional) overrides n with a device-side count.
 *   backward: dz = grad_out * [out > 0];  grad_beta = sum dz;  grad_gamma = sum dz * xhat;
 *             grad_y = gamma / std * (dz - mean(dz) - xhat * mean(dz * xhat)). */
size_t pcdb_bn_train_workspace_bytes(void);
int pcdb_bn_train_fwd(const void *y, int n, const int32_t *n_dev, int c, int dtype, const float *gamma, const float *beta,
                      float eps, float momentum, float *running_mean, float *running_var, int flags, void *out,
                      float *stats, const float *conv_partials, int n_conv_partials, void *workspace,
                      size_t workspace_bytes, void *stream);
int pcdb_bn_train_bwd(const void *grad_out, const void *out, const void *y, int n, const int32_t *n_dev, int c, int dtype,
                      const float *gamma, const float *stats, int flags, void *grad_y, float *grad_gamma,
                      float *grad_beta, int accumulate, void *workspace, size_t workspace_bytes, void *stream);

/* SyncBatchNorm (tools/train.py:94-95, `--sync_bn`, used by every multi-GPU script of the reference): the two calls above in
 * halves, so that the caller can all-reduce (SUM) the per-rank sums in between -- torch.nn.SyncBatchNorm's protocol.
 *   forward   pcdb_bn_train_sums -> sums (2c + 1 doubles: channel sums, sums of squares, row count) -> all-reduce ->
 *             pcdb_bn_train_fwd_from_sums (statistics of ALL ranks' rows; running statistics; out)
 *   backward  pcdb_bn_train_bwd_sums -> sums2 (2c doubles: sum dz, sum dz * xhat) -> all-reduce a copy ->
 *             pcdb_bn_train_bwd_from_sums: grad_gamma / grad_beta from the LOCAL sums (the gradient all-reduce averages them
 *             later), grad_y from the global sums and the global row count (fwd_sums[2c]). */
int pcdb_bn_train_sums(const void *y, int n, const int32_t *n_dev, int c, int dtype, double *sums, void *workspace,
                       size_t workspace_bytes, void *stream);
int pcdb_bn_train_fwd_from_sums(const void *y, int n, const int32_t *n_dev, int c, int dtype, const double *sums,
                                const float *gamma, const float *beta, float eps, float momentum, float *running_mean,
                                float *running_var, int flags, void *out, float *stats, void *stream);
int pcdb_bn_train_bwd_sums(const void *grad_out, const void *out, const void *y, int n, const int32_t *n_dev, int c, int dtype,
                           const float *stats, int flags, double *sums2, void *workspace, size_t workspace_bytes, void *stream);
int pcdb_bn_train_bwd_from_sums(const void *grad_out, const void *out, const void *y, int n, const int32_t *n_dev, int c,
                                int dtype, const float *gamma, const float *stats, const double *local_sums2,
                                const double *global_sums2, const double *fwd_sums, int flags, void *grad_y,
                                float *grad_gamma, float *grad_beta, int accumulate, void *workspace,
                                size_t workspace_bytes, void *stream);

/* spconv.ops.indice_maxpool (SparseMaxPool3d forward; pcdet/models/rcnn/partA2_rcnn_net.py:165):
 *   out[o, c] = max(0, max_k features[nbr[k*ld + o], c])   (the reference's output starts from zeros). */
int pcdb_sparse_maxpool_fwd(const void *features, const int32_t *nbr, int ld, int kernel_volume, int n_out,
                            const int32_t *n_out_dev, int c, int dtype, void *out, void *stream);
/* spconv.ops.indice_maxpool_backward: grad_features[i, c] += grad_out[o, c] for every pair (i, o) of the rulebook with
 * features[i, c] == out[o, c] (fp32; grad_features (n_in, c) zeroed by the caller). */
int pcdb_sparse_maxpool_bwd(const float *features, const float *out, const float *grad_out, const int32_t *nbr, int ld,
                            int kernel_volume, int n_out, int c, float *grad_features, void *stream);

/* RoI-aware point pooling (pcdet/ops/roiaware_pool3d; bound by roiaware_pool3d.cpp:174-179, called from
 * pcdet/models/rcnn/partA2_rcnn_net.py:256-295 and pcdet/ops/roiaware_pool3d/roiaware_pool3d_utils.py:19-66).
 * rois (n_rois,7) [x,y,z bottom centre,w,l,h,rz]; pts (n_pts,3); pts_feature (n_pts,channels);
 * pts_idx_of_voxels (n_rois,out_x,out_y,out_z,max_pts_each_voxel) i32, slot 0 = count, ZEROED by the caller as in
 * the reference (roiaware_pool3d_utils.py:47); pooled_features (n_rois,out_x,out_y,out_z,channels) zeroed by the
 * caller; argmax same shape i32 (max pooling).  pool_method 0 = max, 1 = avg. */
int pcdb_roiaware_pool3d_fwd(const float *rois, int n_rois, const float *pts, int n_pts, const float *pts_feature,
                             int channels, int out_x, int out_y, int out_z, int max_pts_each_voxel,
                             int pool_method, int32_t *argmax, int32_t *pts_idx_of_voxels,
                             float *pooled_features, void *stream);
/* The same with two additions for batches whose per-frame point ranges live on the device (pcdet_b200/parta2.py):
 * pts_range_dev (optional) -> only the points [pts_range_dev[0], pts_range_dev[1]) are candidates (the reference slices
 * the frame on the host after a boolean-mask sync, partA2_rcnn_net.py:272-276); the lists and argmax hold indices into the
 * whole pts / pts_feature arrays.  PCDB_ROI_REUSE_LISTS: pts_idx_of_voxels was filled by an earlier call with the same rois
 * and points (the reference collects the same lists twice, once for the avg and once for the max pooling). */
#define PCDB_ROI_REUSE_LISTS 1
int pcdb_roiaware_pool3d_fwd_ex(const float *rois, int n_rois, const float *pts, int n_pts, const int32_t *pts_range_dev,
                                const float *pts_feature, int channels, int out_x, int out_y, int out_z,
                                int max_pts_each_voxel, int pool_method, int flags, int32_t *argmax,
                                int32_t *pts_idx_of_voxels, float *pooled_features, void *stream);

/* grad_in (n_pts, channels) zeroed by the caller (roiaware_pool3d.cpp:71-98) */
int pcdb_roiaware_pool3d_bwd(const int32_t *pts_idx_of_voxels, const int32_t *argmax, const float *grad_out,
                             int n_rois, int out_x, int out_y, int out_z, int channels, int max_pts_each_voxel,
                             int pool_method, float *grad_in, void *stream);
/* points_in_boxes_gpu (roiaware_pool3d.cpp:100-121): boxes (batch,n_boxes,7), pts (batch,n_pts,3) ->
 * box_idx_of_points (batch,n_pts) = first box containing the point; pre-filled with -1 by the caller. */
int pcdb_points_in_boxes(const float *boxes, int batch, int n_boxes, const float *pts, int n_pts,
                         int32_t *box_idx_of_points, void *stream);

/* SparseConvTensor.dense() (spconv; used at pcdet/models/rpn/rpn_backbone.py:70-74):
 * scatters rows into a zeroed (batch, c, D, H, W) tensor (channels first), dtype in -> dtype out.
 * dense_dtype | PCDB_DENSE_CLEARED: `dense` is already all zeros (the caller cleared it earlier, off its
 * critical path -- the 72 MB memset of the KITTI batch-4 BEV tensor is 13 us); otherwise it is cleared here. */
#define PCDB_DENSE_CLEARED 0x100
int pcdb_to_dense(const void *features, const int32_t *indices, int n, const int32_t *n_dev, int c,
                  int dtype, int batch, const int32_t *spatial_shape_zyx, void *dense, int dense_dtype,
                  void *stream);
/* Backward of pcdb_to_dense (the gradient of SparseConvTensor.dense() with respect to the features): a gather of the dense
 * gradient (batch, c, D, H, W) at the n active sites into features (n, c); either side fp32 or bf16. */
int pcdb_from_dense(const void *dense, int dense_dtype, const int32_t *indices, int n, const int32_t *n_dev, int c,
                    int batch, const int32_t *spatial_shape_zyx, void *features, int dtype, void *stream);

/* Undo of pcdb_to_dense: zeroes exactly the cells of the rows in `indices` (the coordinates a previous pcdb_to_dense
 * scattered), so that a tensor which is reused step after step never needs the full memset again: keep a copy of the
 * coordinates and the count of the last scatter, clear those rows, scatter with PCDB_DENSE_CLEARED. */
int pcdb_dense_clear_rows(const int32_t *indices, int n, const int32_t *n_dev, int c, int batch,
                          const int32_t *spatial_shape_zyx, void *dense, int dense_dtype, void *stream);

/* ---------------------------------------------------------------------------------------------
 * Rotated BEV IoU / NMS.  Replaces the pybind module iou3d_nms_cuda
 * (pcdet/ops/iou3d_nms/src/iou3d_nms.cpp:180-185).  Boxes are (n,5) f32 [x1,y1,x2,y2,ry].
 * ------------------------------------------------------------------------------------------- */
/* boxes_overlap_bev_gpu (iou3d_nms.cpp:36-55): ans (na, nb) f32 intersection areas. */
int pcdb_boxes_overlap_bev(const float *boxes_a, int na, const float *boxes_b, int nb, float *ans,
                           void *stream);
/* boxes_iou_bev_gpu (iou3d_nms.cpp:57-76). */
int pcdb_boxes_iou_bev(const float *boxes_a, int na, const float *boxes_b, int nb, float *ans,
                       void *stream);
/* iou3d_nms_utils.boxes_iou3d_gpu (pcdet/ops/iou3d_nms/iou3d_nms_utils.py:27-59): boxes (n, 7) f32 [x, y, z, w, l, h, ry] in LiDAR
 * coordinates, z = bottom face -> ans (na, nb) 3-D IoU = BEV overlap x height overlap / union of the volumes, in one launch. */
int pcdb_boxes_iou3d(const float *boxes_a, int na, const float *boxes_b, int nb, float *ans, void *stream);

/* nms_gpu / nms_normal_gpu (iou3d_nms.cpp:79-177), batched and fully on the device.
 * boxes: n_sets problems back to back; set s owns rows [set_offsets[s], set_offsets[s+1]) (host
 * array), each already sorted by descending score.  For every set the kept positions (relative
 * to the set start, ascending = score order) go to keep + s*keep_stride (int64, at most keep_stride
 * of them; unused tail = -1) and the count to num_keep[s].  normal != 0 selects axis-aligned IoU.
 * max_boxes_per_set bounds the workspace; the bitmask never leaves the GPU. */
size_t pcdb_nms_workspace_bytes(int n_sets, int max_boxes_per_set);
int pcdb_nms(const float *boxes, const int32_t *set_offsets_host, int n_sets, float thresh, int normal,
             int64_t *keep, int keep_stride, int32_t *num_keep, void *workspace, size_t workspace_bytes,
             void *stream);

/* The same with the number of boxes of every set read on the device: set s owns rows [set_offsets[s], set_offsets[s] +
 * min(set_counts[s], capacity)), capacity = set_offsets[s+1] - set_offsets[s].  For box sets whose size is itself a
 * device result (pcdb_decode_select's count): no host round trip, and the cost follows the count, not the capacity. */
int pcdb_nms_counts(const float *boxes, const int32_t *set_offsets_host, const int32_t *set_counts, int n_sets,
                    float thresh, int normal, int64_t *keep, int keep_stride, int32_t *num_keep, void *workspace,
                    size_t workspace_bytes, void *stream);

/* boxes3d_to_bevboxes_lidar_torch (pcdet/utils/box_utils.py:237-250): (n,7)->(n,5). */
int pcdb_boxes3d_to_bev(const float *boxes3d, int n, float *boxes_bev, void *stream);

/* ---------------------------------------------------------------------------------------------
 * Front of the post-processing (class-agnostic path): Detector3D.predict_boxes + post_processing +
 * class_agnostic_nms up to the NMS call (pcdet/models/detectors/detector3d.py:112-128, 166-215, 278-290),
 * ResidualCoder.decode_with_head_direction_torch (pcdet/utils/box_coder_utils.py:89-144) and the BEV
 * conversion (box_utils.py:237-250), for a whole batch, without host round trips.
 *   cls_preds (batch, n_anchors, cls_stride) f32 logits, the first n_classes of every anchor are used
 *             (encode_background_as_zeros False: pass cls_preds + 1 and n_classes = cls_stride - 1);
 *   box_preds (batch, n_anchors, 7) f32 residuals [xt,yt,zt,wt,lt,ht,rt]; anchors (n_anchors, 7) f32;
 *   dir_cls_preds (batch, n_anchors, num_dir_bins) f32 or NULL; flags & PCDB_DIR_BINARY selects
 *             use_binary_dir_classifier.
 * Per frame: rank score = max over classes, label = first argmax + 1, candidates = sigmoid(score) >=
 * score_thresh, the pre_max best candidates (score desc, ties: lower anchor index first) are decoded into
 *   boxes3d (batch, pre_max, 7), boxes_bev (batch, pre_max, 5), scores (batch, pre_max) raw rank scores,
 *   labels / anchor_index (batch, pre_max) i32, count (batch) i32 = min(pre_max, candidates).
 * Rows >= count hold zero-area BEV boxes far outside any scene (boxes3d 0, label 0, anchor_index -1), so
 * boxes_bev goes to pcdb_nms as `batch` sets of pre_max rows and kept positions >= count[b] are dropped.
 * pre_max <= 16384. */
#define PCDB_DIR_BINARY 1
size_t pcdb_decode_select_workspace_bytes(int batch, int n_anchors, int pre_max);
int pcdb_decode_select(const float *cls_preds, int cls_stride, const float *box_preds, const float *dir_cls_preds,
                       const float *anchors, int batch, int n_anchors, int n_classes, int num_dir_bins,
                       float dir_offset, float dir_limit_offset, float score_thresh, int pre_max, int flags,
                       float *boxes3d, float *boxes_bev, float *scores, int32_t *labels, int32_t *anchor_index,
                       int32_t *count, void *workspace, size_t workspace_bytes, void *stream);

/* Tail of class_agnostic_nms / post_processing (detector3d.py:290-299, 211-219): the first post_max kept positions of
 * pcdb_nms (keep (batch, keep_stride) i64, run on boxes_bev of pcdb_decode_select) -> out_boxes (batch, post_max, 7),
 * out_scores (raw, or sigmoid when sigmoid_scores != 0 = USE_RAW_SCORE False), out_labels / out_selected (anchor index)
 * i64, out_num (batch) i32 = kept real detections; rows >= out_num[b]: boxes 0, score pad_score, label pad_label,
 * selected -1.  With pad_score = -100000 and pad_label = 1 the outputs are the rois / roi_raw_scores / roi_labels of
 * proposal_layer (pcdet/models/model_utils/proposal_layer.py:14-23, 57-60). */
int pcdb_gather_kept(const int64_t *keep, int keep_stride, const int32_t *count, int batch, int pre_max,
                     const float *boxes3d, const float *scores, const int32_t *labels, const int32_t *anchor_index,
                     int post_max, int sigmoid_scores, float pad_score, int pad_label, float *out_boxes,
                     float *out_scores, int64_t *out_labels, int64_t *out_selected, int32_t *out_num, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* PCDET_B200_H_ */
