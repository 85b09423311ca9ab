// Host shim around the REFERENCE kernels (pcdet/ops/iou3d_nms/src/iou3d_nms_kernel.cu, linked from
// /root/reference, not copied).  TEST INFRASTRUCTURE ONLY.  Mirrors the control flow of
// iou3d_nms.cpp:36-126 (cudaMalloc mask -> kernel -> D2H -> serial host sweep) over raw pointers so
// it can be driven through ctypes without torch headers.
#include <cuda_runtime.h>
#include <cstdint>
#include <vector>

void boxesoverlapLauncher(const int num_a, const float *boxes_a, const int num_b, const float *boxes_b, float *ans_overlap);
void boxesioubevLauncher(const int num_a, const float *boxes_a, const int num_b, const float *boxes_b, float *ans_iou);
void nmsLauncher(const float *boxes, unsigned long long *mask, int boxes_num, float nms_overlap_thresh);
void nmsNormalLauncher(const float *boxes, unsigned long long *mask, int boxes_num, float nms_overlap_thresh);

extern "C" {

int ref_boxes_overlap_bev(const float *a_dev, int na, const float *b_dev, int nb, float *out_dev)
{
    boxesoverlapLauncher(na, a_dev, nb, b_dev, out_dev);
    return (int)cudaDeviceSynchronize();
}

int ref_boxes_iou_bev(const float *a_dev, int na, const float *b_dev, int nb, float *out_dev)
{
    boxesioubevLauncher(na, a_dev, nb, b_dev, out_dev);
    return (int)cudaDeviceSynchronize();
}

// boxes_dev: (n,5) score-sorted; keep_host: (n) int64; mask_host (optional): n*ceil(n/64) words.
// Returns the keep count (or a negative CUDA error).
int ref_nms(const float *boxes_dev, int n, float thresh, int normal, int64_t *keep_host,
            unsigned long long *mask_host)
{
    const int col_blocks = (n + 63) / 64;
    unsigned long long *mask_dev = nullptr;
    if (cudaMalloc(&mask_dev, sizeof(unsigned long long) * (size_t)n * col_blocks) != cudaSuccess) return -1;
    if (normal) nmsNormalLauncher(boxes_dev, mask_dev, n, thresh);
    else nmsLauncher(boxes_dev, mask_dev, n, thresh);
    std::vector<unsigned long long> mask((size_t)n * col_blocks);
    cudaError_t e = cudaMemcpy(mask.data(), mask_dev, sizeof(unsigned long long) * mask.size(), cudaMemcpyDeviceToHost);
    cudaFree(mask_dev);
    if (e != cudaSuccess) return -2;
    if (mask_host) for (size_t i = 0; i < mask.size(); ++i) mask_host[i] = mask[i];
    std::vector<unsigned long long> remv(col_blocks, 0ULL);
    int nk = 0;
    for (int i = 0; i < n; ++i) {
        int nb = i / 64, ib = i % 64;
        if (!(remv[nb] & (1ULL << ib))) {
            keep_host[nk++] = i;
            const unsigned long long *p = mask.data() + (size_t)i * col_blocks;
            for (int j = nb; j < col_blocks; ++j) remv[j] |= p[j];
        }
    }
    return nk;
}

}  // extern "C"
