// Point-cloud ingest on the device: camera-FOV filter + x/y range filter in front of the voxelizer.
//
// Replaces, for raw KITTI clouds already copied to the GPU:
//   KittiDataset.__getitem__            pcdet/datasets/kitti/kitti_dataset.py:714-717  (FOV_POINTS_ONLY)
//   Calibration.lidar_to_rect           pcdet/utils/calibration.py:66-74
//   Calibration.rect_to_img             pcdet/utils/calibration.py:76-85
//   KittiDataset.get_fov_flag           kitti_dataset.py:236-253
//   common_utils.mask_points_by_range   pcdet/utils/common_utils.py:47-51 (called at pcdet/datasets/dataset.py:184)
// The reference does this with numpy in a DataLoader worker and ships the filtered cloud through pickling and a
// pageable H2D copy of the VOXELS; here the raw .bin payload goes to the device once and the filtered cloud feeds
// pcdb_voxelize without leaving it.  The compaction keeps the point order (first-come voxel ids depend on it):
// per-block counts, block prefix, ballot ranks -- no atomics, deterministic.  HBM-bound: 16 B read per point,
// 16 B written per surviving point.
#include "common.cuh"
#include "../../include/pcdet_b200.h"

namespace pcdb {

constexpr int kIngBlock = 256;

struct FrameCalib {
    float v2r[12];      // (4,3) row-major: [x y z 1] . v2r = rectified camera coordinates (V2C^T . R0^T)
    float p2[12];       // (3,4) row-major camera projection P2
    float img_h, img_w;
};

// true when the point survives; frame calibration in shared or constant memory
__device__ __forceinline__ bool point_valid(const float *__restrict__ p, const FrameCalib *c, const float *range_xy)
{
    const float x = p[0], y = p[1], z = p[2];
    if (range_xy && !(x >= range_xy[0] && x <= range_xy[2] && y >= range_xy[1] && y <= range_xy[3])) return false;
    if (!c) return true;
    // calibration.py:72: np.dot(pts_lidar_hom, M) -- row times column, accumulated left to right
    float r[3];
#pragma unroll
    for (int j = 0; j < 3; ++j)
        r[j] = fmaf(z, c->v2r[6 + j], fmaf(y, c->v2r[3 + j], x * c->v2r[j])) + c->v2r[9 + j];
    // calibration.py:82-84: pts_2d_hom = [r 1] . P2^T ; image = hom[:2] / r.z ; depth = hom[2] - P2[2][3]
    float h[3];
#pragma unroll
    for (int j = 0; j < 3; ++j)
        h[j] = fmaf(r[2], c->p2[j * 4 + 2], fmaf(r[1], c->p2[j * 4 + 1], r[0] * c->p2[j * 4])) + c->p2[j * 4 + 3];
    const float u = __fdiv_rn(h[0], r[2]), v = __fdiv_rn(h[1], r[2]);
    const float depth = h[2] - c->p2[11];
    return u >= 0.f && u < c->img_w && v >= 0.f && v < c->img_h && depth >= 0.f;      // kitti_dataset.py:248-251
}

__device__ __forceinline__ int frame_of(const int *__restrict__ offs, int batch, int i)
{
    int b = 0;
    while (b + 1 < batch && i >= offs[b + 1]) ++b;
    return b;
}

// pass 1: validity flags (one byte per point) and per-block counts.  Index n is a sentinel (never valid) so that
// frame boundaries equal to n get an output offset in pass 2.
__global__ void __launch_bounds__(kIngBlock)
ing_flag_count(const float *__restrict__ points, int n, int c, const int *__restrict__ frame_offsets, int batch,
               const FrameCalib *__restrict__ calib, const float *__restrict__ range_xy, unsigned char *__restrict__ flags,
               int *__restrict__ blk_counts)
{
    __shared__ int s_offs[65];
    __shared__ float s_range[4];
    __shared__ int s_w[kIngBlock / 32];
    for (int j = threadIdx.x; j <= batch; j += kIngBlock) s_offs[j] = __ldg(frame_offsets + j);
    if (range_xy && threadIdx.x < 4) s_range[threadIdx.x] = __ldg(range_xy + threadIdx.x);
    __syncthreads();
    const int i = blockIdx.x * kIngBlock + threadIdx.x;
    bool ok = false;
    if (i < n && i >= s_offs[0] && i < s_offs[batch])
        ok = point_valid(points + (size_t)i * c, calib ? calib + frame_of(s_offs, batch, i) : nullptr, range_xy ? s_range : nullptr);
    if (i < n) flags[i] = ok;
    const uint32_t m = __ballot_sync(0xffffffffu, ok);
    if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = __popc(m);
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0;
        for (int w = 0; w < kIngBlock / 32; ++w) t += s_w[w];
        blk_counts[blockIdx.x] = t;
    }
}

// pass 2: exclusive prefix of the block counts (warp 0), ordered scatter, output frame offsets
__global__ void __launch_bounds__(kIngBlock)
ing_scatter(const float *__restrict__ points, int n, int c, const int *__restrict__ frame_offsets, int batch,
            const unsigned char *__restrict__ flags, const int *__restrict__ blk_counts, float *__restrict__ out_points,
            int *__restrict__ out_offsets, int *__restrict__ out_index)
{
    __shared__ int s_offs[65];
    __shared__ int s_base;
    __shared__ int s_w[kIngBlock / 32];
    for (int j = threadIdx.x; j <= batch; j += kIngBlock) s_offs[j] = __ldg(frame_offsets + j);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (warp == 0) {
        int t = 0;
        for (int j = lane; j < (int)blockIdx.x; j += 32) t += __ldg(blk_counts + j);
#pragma unroll
        for (int d = 16; d; d >>= 1) t += __shfl_xor_sync(0xffffffffu, t, d);
        if (lane == 0) s_base = t;
    }
    const int i = blockIdx.x * kIngBlock + threadIdx.x;
    const bool ok = i < n && flags[i];
    const uint32_t m = __ballot_sync(0xffffffffu, ok);
    if (lane == 0) s_w[warp] = __popc(m);
    __syncthreads();
    int pos = s_base + __popc(m & ((1u << lane) - 1u));
    for (int w = 0; w < warp; ++w) pos += s_w[w];
    if (ok) {
        const float *src = points + (size_t)i * c;
        float *dst = out_points + (size_t)pos * c;
        if (c == 4) {
            *reinterpret_cast<float4 *>(dst) = __ldg(reinterpret_cast<const float4 *>(src));
        } else {
            for (int j = 0; j < c; ++j) dst[j] = __ldg(src + j);
        }
        if (out_index) out_index[pos] = i;
    }
    // pos = surviving points before input index i: the output offset of every frame that starts at i
    if (i <= n) {
        for (int b = 0; b <= batch; ++b) {
            const int o = min(max(s_offs[b], s_offs[0]), n);
            if (o == i) out_offsets[b] = pos;
        }
    }
}

}  // namespace pcdb

using namespace pcdb;

extern "C" size_t pcdb_filter_points_workspace_bytes(int n_points)
{
    if (n_points < 0) return 0;
    const size_t blocks = (size_t)n_points / kIngBlock + 1;
    return align_up((size_t)n_points + 1, 256) + align_up(sizeof(int) * blocks, 256);
}

extern "C" int pcdb_filter_points(const float *points, int n, int c, const int32_t *frame_offsets, int batch,
                                  const float *calib, const float *range_xy, float *out_points, int32_t *out_offsets,
                                  int32_t *out_index, void *workspace, size_t workspace_bytes, void *stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    if (n < 0 || c < 3 || batch < 1 || batch > 64 || !frame_offsets || !out_points || !out_offsets || (n > 0 && !points)) {
        set_last_error("pcdb_filter_points: invalid argument (n=%d c=%d batch=%d, at most 64 frames per call)", n, c, batch);
        return kInvalidArgument;
    }
    if (c == 4 && ((((uintptr_t)points) | ((uintptr_t)out_points)) & 15)) {
        set_last_error("pcdb_filter_points: (n,4) point arrays must be 16-byte aligned");
        return kInvalidArgument;
    }
    const size_t need = pcdb_filter_points_workspace_bytes(n);
    if (!workspace || workspace_bytes < need) {
        set_last_error("pcdb_filter_points: workspace %zu < required %zu bytes", workspace_bytes, need);
        return kWorkspaceTooSmall;
    }
    unsigned char *flags = (unsigned char *)workspace;
    int *blk_counts = (int *)((char *)workspace + align_up((size_t)n + 1, 256));
    const int blocks = n / kIngBlock + 1;                         // covers the sentinel index n
    ing_flag_count<<<blocks, kIngBlock, 0, stream>>>(points, n, c, frame_offsets, batch, (const FrameCalib *)calib, range_xy, flags,
                                                     blk_counts);
    ing_scatter<<<blocks, kIngBlock, 0, stream>>>(points, n, c, frame_offsets, batch, flags, blk_counts, out_points, out_offsets,
                                                  out_index);
    return check_launch("pcdb_filter_points");
}
