// RoI-aware point pooling (Part-A^2 stage 2) for sm_100a.  SURVEY §8(f) rank 3.
//
// Replaces pcdet/ops/roiaware_pool3d/src/roiaware_pool3d_kernel.cu (generate_pts_mask_for_box3d :43,
// collect_inside_pts_for_box3d :84, roiaware_maxpool3d :116 / avgpool3d :170, the two backward kernels :242 / :264,
// points_in_boxes_kernel :312) and the host code of roiaware_pool3d.cpp:27-98.  The reference materialises an
// (N boxes x M points) int mask in a per-call cudaMalloc and then walks it with ONE THREAD PER BOX (:84-114,
// M serial iterations each).  Here one warp per box streams the points in order, 32 at a time: lanes inside the box
// that fall into the same pooling voxel find each other with __match_any_sync, the lowest lane reserves their
// slots in a shared-memory counter array and every lane stores its point index -- the same lists in the same
// (point index) order, no mask array, no allocation.
#include "common.cuh"
#include "../../include/pcdet_b200.h"

namespace pcdb {

// check_pt_in_box3d (roiaware_pool3d_kernel.cu:25-40): box = (cx, cy, cz bottom centre, w, l, h, rz)
__device__ __forceinline__ bool pt_in_box3d(float x, float y, float z, const float *__restrict__ box, float *local_x, float *local_y)
{
    const float cx = box[0], cy = box[1], w = box[3], l = box[4], h = box[5], rz = box[6];
    const float cz = (float)((double)box[2] + (double)h / 2.0);
    if ((double)fabsf(z - cz) > (double)h / 2.0) return false;
    const float rot_angle = (float)((double)rz + 3.14159265358979323846 / 2);
    const float cosa = cosf(rot_angle), sina = sinf(rot_angle);
    const float sx = x - cx, sy = y - cy;
    *local_x = sx * cosa + sy * (-sina);
    *local_y = sx * sina + sy * cosa;
    return ((double)*local_x > -(double)l / 2.0) & ((double)*local_x < (double)l / 2.0) & ((double)*local_y > -(double)w / 2.0) &
           ((double)*local_y < (double)w / 2.0);
}

// grid: n_rois CTAs of 32 threads; dynamic shared memory: one int counter per pooling voxel (or none -> the counter
// slots of pts_idx_of_voxels themselves are used)
__global__ void __launch_bounds__(32)
roi_collect_kernel(const float *__restrict__ rois, const float *__restrict__ pts, int n_pts, const int *__restrict__ range_dev,
                   int out_x, int out_y, int out_z, int max_pts_each_voxel, int *__restrict__ pts_idx_of_voxels, int use_smem)
{
    // range_dev (optional): only the points [range_dev[0], range_dev[1]) are candidates -- one frame of a batch whose
    // per-frame offsets live on the device; the lists keep the indices into the whole array
    int p_begin = 0;
    if (range_dev) {
        p_begin = max(0, __ldg(range_dev));
        n_pts = min(n_pts, __ldg(range_dev + 1));
    }
    extern __shared__ int s_cnt[];
    const int lane = threadIdx.x, n_vox = out_x * out_y * out_z;
    const float *box = rois + (size_t)blockIdx.x * 7;
    int *lists = pts_idx_of_voxels + (size_t)blockIdx.x * n_vox * max_pts_each_voxel;
    if (use_smem) {
        for (int v = lane; v < n_vox; v += 32) s_cnt[v] = 0;
        __syncwarp();
    }
    const int max_num = max_pts_each_voxel - 1;                 // slot 0 of every list is the counter
    const float w = box[3], l = box[4], h = box[5];
    const float x_res = l / out_x, y_res = w / out_y, z_res = h / out_z;
    // Conservative pre-test: a point inside the box lies inside the circle around its centre through the corners.  The
    // radius is inflated well past any fp32 rounding of the exact test, so the pre-test never rejects a point the
    // reference's arithmetic (pt_in_box3d, evaluated unchanged on the survivors) accepts; it spares the other ~99 % of the
    // points the rotation and the double-precision comparisons.
    const float r2 = (0.25f * (l * l + w * w)) * 1.001f + 1e-4f;
    // The walk is in point order (the lists keep the FIRST points of a voxel), one warp per box: what it waits for is the
    // latency of the coordinate loads (measured 870 cycles per 32 points with one chunk in flight), so eight chunks are
    // loaded before the first is examined.
    constexpr int kAhead = 8;
    for (int q0 = p_begin; q0 < n_pts; q0 += 32 * kAhead) {
      float px[kAhead], py[kAhead], pz[kAhead];
#pragma unroll
      for (int j = 0; j < kAhead; ++j) {
          const int p = q0 + 32 * j + lane;
          const bool ok = p < n_pts;
          px[j] = ok ? __ldg(pts + (size_t)p * 3) : 0.f;
          py[j] = ok ? __ldg(pts + (size_t)p * 3 + 1) : 0.f;
          pz[j] = ok ? __ldg(pts + (size_t)p * 3 + 2) : 0.f;
      }
#pragma unroll
      for (int j = 0; j < kAhead; ++j) {
        const int p = q0 + 32 * j + lane;
        int vox = -1;
        if (p < n_pts) {
            const float x = px[j], y = py[j], z = pz[j];
            const float dx = x - box[0], dy = y - box[1];
            float lx, ly;
            if (dx * dx + dy * dy <= r2 && pt_in_box3d(x, y, z, box, &lx, &ly)) {
                // generate_pts_mask_for_box3d (:61-75); the unsigned conversions clamp negatives to the last voxel
                const float lz = z - box[2];
                unsigned int xi = (unsigned int)(int)((lx + l / 2) / x_res);
                unsigned int yi = (unsigned int)(int)((ly + w / 2) / y_res);
                unsigned int zi = (unsigned int)(int)(lz / z_res);
                xi = min(max(xi, 0u), (unsigned int)(out_x - 1));
                yi = min(max(yi, 0u), (unsigned int)(out_y - 1));
                zi = min(max(zi, 0u), (unsigned int)(out_z - 1));
                // the reference packs the indices into 8-bit fields (:71, :99-101)
                vox = (int)(((xi & 0xFF) * out_y + (yi & 0xFF)) * out_z + (zi & 0xFF));
            }
        }
        const unsigned active = __ballot_sync(0xffffffffu, vox >= 0);
        if (active == 0u) continue;
        if (vox >= 0) {
            const unsigned peers = __match_any_sync(active, vox);
            const int rank = __popc(peers & ((1u << lane) - 1u));
            const int leader = __ffs(peers) - 1;
            int base = 0;
            if (lane == leader) {
                int *c = use_smem ? &s_cnt[vox] : &lists[(size_t)vox * max_pts_each_voxel];
                base = *c;
                *c = min(base + __popc(peers), max_num);
            }
            base = __shfl_sync(peers, base, leader);
            if (base + rank < max_num) lists[(size_t)vox * max_pts_each_voxel + base + rank + 1] = p;
        }
        __syncwarp();
      }
    }
    if (use_smem)
        for (int v = lane; v < n_vox; v += 32) lists[(size_t)v * max_pts_each_voxel] = s_cnt[v];
}

// one thread per (roi, voxel, channel); pool_method 0 = max (first maximum in list order, strict >), 1 = avg
__global__ void __launch_bounds__(256)
roi_pool_kernel(int n_rois, int channels, int max_pts_each_voxel, int n_vox, const float *__restrict__ pts_feature,
                const int *__restrict__ pts_idx_of_voxels, float *__restrict__ pooled, int *__restrict__ argmax, int pool_method)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)n_rois * n_vox * channels) return;
    const int c = (int)(t % channels);
    const long long rv = t / channels;                       // roi * n_vox + voxel
    const int *list = pts_idx_of_voxels + rv * max_pts_each_voxel;
    const int total = list[0];
    if (pool_method == 0) {
        int arg = -1;
        float best = -1e50f;                                 // = -inf in fp32, as in the reference (:143)
        for (int k = 1; k <= total; ++k) {
            const float v = __ldg(pts_feature + (size_t)list[k] * channels + c);
            if (v > best) { best = v; arg = list[k]; }
        }
        if (arg != -1) pooled[t] = best;
        argmax[t] = arg;
    } else {
        float sum = 0.f;
        for (int k = 1; k <= total; ++k) sum += __ldg(pts_feature + (size_t)list[k] * channels + c);
        if (total > 0) pooled[t] = sum / total;
    }
}

__global__ void __launch_bounds__(256)
roi_pool_bwd_kernel(int n_rois, int channels, int max_pts_each_voxel, int n_vox, const int *__restrict__ pts_idx_of_voxels,
                    const int *__restrict__ argmax, const float *__restrict__ grad_out, float *__restrict__ grad_in, int pool_method)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)n_rois * n_vox * channels) return;
    const int c = (int)(t % channels);
    if (pool_method == 0) {
        const int a = argmax[t];
        if (a != -1) atomicAdd(grad_in + (size_t)a * channels + c, grad_out[t]);
    } else {
        const int *list = pts_idx_of_voxels + (t / channels) * max_pts_each_voxel;
        const int total = list[0];
        const float g = grad_out[t] * (1 / fmaxf((float)total, 1.0f));
        for (int k = 1; k <= total; ++k) atomicAdd(grad_in + (size_t)list[k] * channels + c, g);
    }
}

// points_in_boxes_kernel (:312-333): index of the FIRST box containing the point, output pre-filled with -1 by the caller
__global__ void __launch_bounds__(256)
points_in_boxes_kernel(int n_boxes, int n_pts, const float *__restrict__ boxes, const float *__restrict__ pts,
                       int *__restrict__ box_idx_of_points)
{
    const int b = blockIdx.y, p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n_pts) return;
    const float *q = pts + ((size_t)b * n_pts + p) * 3;
    const float x = q[0], y = q[1], z = q[2];
    for (int k = 0; k < n_boxes; ++k) {
        float lx, ly;
        if (pt_in_box3d(x, y, z, boxes + ((size_t)b * n_boxes + k) * 7, &lx, &ly)) {
            box_idx_of_points[(size_t)b * n_pts + p] = k;
            break;
        }
    }
}

}  // namespace pcdb

using namespace pcdb;

extern "C" int pcdb_roiaware_pool3d_fwd_ex(const float *rois, int n_rois, const float *pts, int n_pts, const int32_t *pts_range_dev,
                                           const float *pts_feature, int channels, int out_x, int out_y, int out_z,
                                           int max_pts_each_voxel, int pool_method, int flags, int32_t *argmax,
                                           int32_t *pts_idx_of_voxels, float *pooled_features, void *stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    if (n_rois < 0 || n_pts < 0 || channels < 1 || out_x < 1 || out_y < 1 || out_z < 1 || out_x > 256 || out_y > 256 ||
        out_z > 256 || max_pts_each_voxel < 2 || (pool_method != 0 && pool_method != 1) || !pts_idx_of_voxels ||
        !pooled_features || (pool_method == 0 && !argmax)) {
        set_last_error("pcdb_roiaware_pool3d_fwd: invalid argument (n_rois=%d n_pts=%d channels=%d out=%dx%dx%d max_pts=%d)",
                       n_rois, n_pts, channels, out_x, out_y, out_z, max_pts_each_voxel);
        return kInvalidArgument;
    }
    if (n_rois == 0) return kOk;
    const int n_vox = out_x * out_y * out_z;
    if (!(flags & PCDB_ROI_REUSE_LISTS)) {
        const size_t smem = (size_t)n_vox * 4;
        const int use_smem = smem <= 48 * 1024;
        roi_collect_kernel<<<n_rois, 32, use_smem ? smem : 0, stream>>>(rois, pts, n_pts, pts_range_dev, out_x, out_y, out_z,
                                                                      max_pts_each_voxel, pts_idx_of_voxels, use_smem);
    }
    const long long total = (long long)n_rois * n_vox * channels;
    roi_pool_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(n_rois, channels, max_pts_each_voxel, n_vox, pts_feature,
                                                                         pts_idx_of_voxels, pooled_features, argmax, pool_method);
    return check_launch("pcdb_roiaware_pool3d_fwd");
}

extern "C" int pcdb_roiaware_pool3d_fwd(const float *rois, int n_rois, const float *pts, int n_pts, const float *pts_feature,
                                        int channels, int out_x, int out_y, int out_z, int max_pts_each_voxel,
                                        int pool_method, int32_t *argmax, int32_t *pts_idx_of_voxels,
                                        float *pooled_features, void *stream_)
{
    return pcdb_roiaware_pool3d_fwd_ex(rois, n_rois, pts, n_pts, nullptr, pts_feature, channels, out_x, out_y, out_z,
                                       max_pts_each_voxel, pool_method, 0, argmax, pts_idx_of_voxels, pooled_features, stream_);
}

extern "C" int pcdb_roiaware_pool3d_bwd(const int32_t *pts_idx_of_voxels, const int32_t *argmax, const float *grad_out,
                                        int n_rois, int out_x, int out_y, int out_z, int channels, int max_pts_each_voxel,
                                        int pool_method, float *grad_in, void *stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    if (n_rois < 0 || channels < 1 || !grad_out || !grad_in || (pool_method == 0 ? !argmax : !pts_idx_of_voxels)) {
        set_last_error("pcdb_roiaware_pool3d_bwd: invalid argument");
        return kInvalidArgument;
    }
    if (n_rois == 0) return kOk;
    const int n_vox = out_x * out_y * out_z;
    const long long total = (long long)n_rois * n_vox * channels;
    roi_pool_bwd_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(n_rois, channels, max_pts_each_voxel, n_vox,
                                                                             pts_idx_of_voxels, argmax, grad_out, grad_in, pool_method);
    return check_launch("pcdb_roiaware_pool3d_bwd");
}

extern "C" int pcdb_points_in_boxes(const float *boxes, int batch, int n_boxes, const float *pts, int n_pts,
                                    int32_t *box_idx_of_points, void *stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    if (batch < 0 || n_boxes < 0 || n_pts < 0 || !box_idx_of_points) {
        set_last_error("pcdb_points_in_boxes: invalid argument");
        return kInvalidArgument;
    }
    if (batch == 0 || n_pts == 0) return kOk;
    points_in_boxes_kernel<<<dim3((n_pts + 255) / 256, batch), 256, 0, stream>>>(n_boxes, n_pts, boxes, pts, box_idx_of_points);
    return check_launch("pcdb_points_in_boxes");
}
