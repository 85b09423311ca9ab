// Host shim around the REFERENCE kernels of pcdet/ops/roiaware_pool3d/src/roiaware_pool3d_kernel.cu (linked from
// /root/reference, not copied).  TEST INFRASTRUCTURE ONLY: lets the GPU tests and tests/golden/make_golden_gpu.py
// drive the reference launchers through ctypes over raw device pointers (what roiaware_pool3d.cpp:27-121 does
// with torch tensors).
#include <cuda_runtime.h>

void roiaware_pool3d_launcher(int boxes_num, int pts_num, int channels, int max_pts_each_voxel, int out_x, int out_y, int out_z,
    const float *rois, const float *pts, const float *pts_feature, int *argmax, int *pts_idx_of_voxels, float *pooled_features,
    int pool_method);
void points_in_boxes_launcher(int batch_size, int boxes_num, int pts_num, const float *boxes, const float *pts, int *box_idx_of_points);

extern "C" {

int ref_roiaware_pool3d(const float *rois, int n_rois, const float *pts, int n_pts, const float *feat, int channels, int out_x,
                        int out_y, int out_z, int max_pts, int pool_method, int *argmax, int *pts_idx_of_voxels, float *pooled)
{
    roiaware_pool3d_launcher(n_rois, n_pts, channels, max_pts, out_x, out_y, out_z, rois, pts, feat, argmax, pts_idx_of_voxels,
                             pooled, pool_method);
    return (int)cudaDeviceSynchronize();
}

int ref_points_in_boxes(const float *boxes, int batch, int n_boxes, const float *pts, int n_pts, int *box_idx)
{
    points_in_boxes_launcher(batch, n_boxes, n_pts, boxes, pts, box_idx);
    return (int)cudaDeviceSynchronize();
}

}  // extern "C"
