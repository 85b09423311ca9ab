// SparseConvTensor.dense() for sm_100a: scatter active rows into a zeroed channels-first tensor.
// Replaces spconv's scatter_nd + permute(0,4,1,2,3).contiguous(), used at
// pcdet/models/rpn/rpn_backbone.py:70-74: three passes over the dense tensor there, one memset plus
// one pass over the ACTIVE rows here (the dense tensor is ~97 % zeros after BackBone8x).
#include "common.cuh"
#include "../../include/pcdet_b200.h"

namespace pcdb {

// One thread per (row, group of 8 channels): one coordinate load and one 16/32-byte feature load feed eight
// scattered stores (channels are D*H*W elements apart in the channels-first tensor).
template <typename TIn, typename TOut>
__global__ void __launch_bounds__(256)
to_dense_kernel(const TIn *__restrict__ feat, const int4 *__restrict__ indices, int n, const int *__restrict__ n_dev,
                int c, int D, int H, int W, TOut *__restrict__ dense)
{
    if (n_dev) { const int m = __ldg(n_dev); n = m < n ? m : n; }
    // Consecutive THREADS take consecutive ROWS of the same channel group, so that a warp's store instruction writes 32 cells
    // of one channel plane; the rows of a level come in ascending (b, z, y, x) order from the rulebook chain, i.e. mostly
    // x-adjacent cells = a few 32-byte sectors per instruction instead of 32 (the loads pay for it: 32 rows x 16 B).
    const int groups = (c + 7) >> 3;
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)((n + 31) / 32) * 32 * groups) return;
    const long long tile = t / (32 * groups);
    const int within = (int)(t % (32 * groups));
    const int row = (int)(tile * 32) + (within & 31), ch0 = (within >> 5) * 8;
    if (row >= n) return;
    const int4 p = __ldg(indices + row);
    const size_t vol = (size_t)D * H * W;
    TOut *dst = dense + ((size_t)p.x * c + ch0) * vol + ((size_t)p.y * H + p.z) * W + p.w;
    const TIn *src = feat + (size_t)row * c + ch0;
    TIn v[8];
    if (ch0 + 8 <= c && (c & 7) == 0) {
        if constexpr (sizeof(TIn) == 2) {
            *reinterpret_cast<uint4 *>(v) = __ldg(reinterpret_cast<const uint4 *>(src));
        } else {
            *reinterpret_cast<uint4 *>(v) = __ldg(reinterpret_cast<const uint4 *>(src));
            *reinterpret_cast<uint4 *>(v + 4) = __ldg(reinterpret_cast<const uint4 *>(src) + 1);
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) dst[(size_t)j * vol] = from_float<TOut>(to_float(v[j]));
    } else {
        for (int j = 0; j < 8 && ch0 + j < c; ++j) dst[(size_t)j * vol] = from_float<TOut>(to_float(src[j]));
    }
}

// The inverse: zero exactly the cells a previous pcdb_to_dense wrote.  A BEV map of BackBone8x is ~97 % zeros, so
// undoing the last scatter (2.4 MB for a KITTI batch of 4) replaces the 72 MB memset of the whole tensor -- which, even
// on its own stream, took 5 % of the step away from the kernels it ran next to.
template <typename TOut>
__global__ void __launch_bounds__(256)
dense_clear_rows_kernel(const int4 *__restrict__ indices, int n, const int *__restrict__ n_dev, int c, int D, int H, int W,
                        TOut *__restrict__ dense)
{
    if (n_dev) { const int m = __ldg(n_dev); n = m < n ? m : n; }
    // Consecutive THREADS take consecutive ROWS of the same channel group, so that a warp's store instruction writes 32 cells
    // of one channel plane; the rows of a level come in ascending (b, z, y, x) order from the rulebook chain, i.e. mostly
    // x-adjacent cells = a few 32-byte sectors per instruction instead of 32 (the loads pay for it: 32 rows x 16 B).
    const int groups = (c + 7) >> 3;
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)((n + 31) / 32) * 32 * groups) return;
    const long long tile = t / (32 * groups);
    const int within = (int)(t % (32 * groups));
    const int row = (int)(tile * 32) + (within & 31), ch0 = (within >> 5) * 8;
    if (row >= n) return;
    const int4 p = __ldg(indices + row);
    const size_t vol = (size_t)D * H * W;
    TOut *dst = dense + ((size_t)p.x * c + ch0) * vol + ((size_t)p.y * H + p.z) * W + p.w;
    for (int j = 0; j < 8 && ch0 + j < c; ++j) dst[(size_t)j * vol] = from_float<TOut>(0.f);
}

// Backward of to_dense: grad_features[row, ch] = grad_dense[b, ch, z, y, x] at the row's site (a gather; rows beyond the
// count are left alone).  One thread per (row, group of 8 channels).
template <typename TIn, typename TOut>
__global__ void __launch_bounds__(256)
from_dense_kernel(const TIn *__restrict__ dense, const int4 *__restrict__ indices, int n, const int *__restrict__ n_dev,
                  int c, int D, int H, int W, TOut *__restrict__ feat)
{
    if (n_dev) { const int m = __ldg(n_dev); n = m < n ? m : n; }
    // Consecutive THREADS take consecutive ROWS of the same channel group, so that a warp's store instruction writes 32 cells
    // of one channel plane; the rows of a level come in ascending (b, z, y, x) order from the rulebook chain, i.e. mostly
    // x-adjacent cells = a few 32-byte sectors per instruction instead of 32 (the loads pay for it: 32 rows x 16 B).
    const int groups = (c + 7) >> 3;
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)((n + 31) / 32) * 32 * groups) return;
    const long long tile = t / (32 * groups);
    const int within = (int)(t % (32 * groups));
    const int row = (int)(tile * 32) + (within & 31), ch0 = (within >> 5) * 8;
    if (row >= n) return;
    const int4 p = __ldg(indices + row);
    const size_t vol = (size_t)D * H * W;
    const TIn *src = dense + ((size_t)p.x * c + ch0) * vol + ((size_t)p.y * H + p.z) * W + p.w;
    TOut *dst = feat + (size_t)row * c + ch0;
    for (int j = 0; j < 8 && ch0 + j < c; ++j) dst[j] = from_float<TOut>(to_float(__ldg(src + (size_t)j * vol)));
}

}  // namespace pcdb

using namespace pcdb;

extern "C" int pcdb_from_dense(const void *dense, int dense_dtype, const int32_t *indices, int n, const int32_t *n_dev, int c,
                               int batch, const int32_t *spatial_shape_zyx, void *features, int dtype, void *stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    if (n < 0 || c < 1 || batch < 1 || !dense || !spatial_shape_zyx || !features || (n > 0 && !indices)) {
        set_last_error("pcdb_from_dense: invalid argument");
        return kInvalidArgument;
    }
    if (n == 0) return kOk;
    const int D = spatial_shape_zyx[0], H = spatial_shape_zyx[1], W = spatial_shape_zyx[2];
    const long long total = (long long)((n + 31) / 32) * 32 * ((c + 7) / 8);
    const int nb = (int)((total + 255) / 256);
    const int4 *idx = (const int4 *)indices;
    if (dense_dtype == PCDB_BF16 && dtype == PCDB_BF16)
        from_dense_kernel<<<nb, 256, 0, stream>>>((const __nv_bfloat16 *)dense, idx, n, n_dev, c, D, H, W, (__nv_bfloat16 *)features);
    else if (dense_dtype == PCDB_BF16)
        from_dense_kernel<<<nb, 256, 0, stream>>>((const __nv_bfloat16 *)dense, idx, n, n_dev, c, D, H, W, (float *)features);
    else if (dtype == PCDB_BF16)
        from_dense_kernel<<<nb, 256, 0, stream>>>((const float *)dense, idx, n, n_dev, c, D, H, W, (__nv_bfloat16 *)features);
    else
        from_dense_kernel<<<nb, 256, 0, stream>>>((const float *)dense, idx, n, n_dev, c, D, H, W, (float *)features);
    return check_launch("pcdb_from_dense");
}

extern "C" int pcdb_dense_clear_rows(const int32_t *indices, int n, const int32_t *n_dev, int c, int batch,
                                     const int32_t *spatial_shape_zyx, void *dense, int dense_dtype, void *stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    if (n < 0 || c < 1 || batch < 1 || !dense || !spatial_shape_zyx || (n > 0 && !indices)) {
        set_last_error("pcdb_dense_clear_rows: invalid argument");
        return kInvalidArgument;
    }
    if (n == 0) return kOk;
    const int D = spatial_shape_zyx[0], H = spatial_shape_zyx[1], W = spatial_shape_zyx[2];
    const long long total = (long long)((n + 31) / 32) * 32 * ((c + 7) / 8);
    const int nb = (int)((total + 255) / 256);
    if (dense_dtype == PCDB_BF16)
        dense_clear_rows_kernel<<<nb, 256, 0, stream>>>((const int4 *)indices, n, n_dev, c, D, H, W, (__nv_bfloat16 *)dense);
    else
        dense_clear_rows_kernel<<<nb, 256, 0, stream>>>((const int4 *)indices, n, n_dev, c, D, H, W, (float *)dense);
    return check_launch("pcdb_dense_clear_rows");
}

extern "C" int pcdb_to_dense(const void *features, const int32_t *indices, int n, const int32_t *n_dev, int c,
                             int dtype, int batch, const int32_t *spatial_shape_zyx, void *dense, int dense_dtype,
                             void *stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    if (n < 0 || c < 1 || batch < 1 || !dense || !spatial_shape_zyx) {
        set_last_error("pcdb_to_dense: invalid argument");
        return kInvalidArgument;
    }
    const int D = spatial_shape_zyx[0], H = spatial_shape_zyx[1], W = spatial_shape_zyx[2];
    const size_t elems = (size_t)batch * c * D * H * W;
    const bool cleared = (dense_dtype & PCDB_DENSE_CLEARED) != 0;     // the caller zeroed it (off its critical path)
    dense_dtype &= ~PCDB_DENSE_CLEARED;
    if (!cleared) cudaMemsetAsync(dense, 0, elems * (dense_dtype == PCDB_BF16 ? 2 : 4), stream);
    if (n > 0) {
        const long long total = (long long)((n + 31) / 32) * 32 * ((c + 7) / 8);
        const int nb = (int)((total + 255) / 256);
        const int4 *idx = (const int4 *)indices;
        if (dtype == PCDB_BF16 && dense_dtype == PCDB_BF16)
            to_dense_kernel<<<nb, 256, 0, stream>>>((const __nv_bfloat16 *)features, idx, n, n_dev, c, D, H, W, (__nv_bfloat16 *)dense);
        else if (dtype == PCDB_BF16)
            to_dense_kernel<<<nb, 256, 0, stream>>>((const __nv_bfloat16 *)features, idx, n, n_dev, c, D, H, W, (float *)dense);
        else if (dense_dtype == PCDB_BF16)
            to_dense_kernel<<<nb, 256, 0, stream>>>((const float *)features, idx, n, n_dev, c, D, H, W, (__nv_bfloat16 *)dense);
        else
            to_dense_kernel<<<nb, 256, 0, stream>>>((const float *)features, idx, n, n_dev, c, D, H, W, (float *)dense);
    }
    return check_launch("pcdb_to_dense");
}
