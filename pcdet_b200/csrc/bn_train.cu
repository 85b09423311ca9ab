// Train-mode BatchNorm1d (+ ReLU) over the rows of a sparse tensor, forward and backward.
//
// Every block of the reference backbone is conv -> BatchNorm1d -> ReLU (pcdet/models/rpn/rpn_backbone.py:79-103) and the
// training loop runs it in train mode (tools/train.py:119-122): batch statistics over the active rows, running
// statistics updated with `momentum`, unbiased variance in the running estimate -- torch.nn.BatchNorm1d's contract.
// HBM-bound row streaming, 16-byte accesses, fp32 arithmetic on bf16 or fp32 storage:
//
//   forward   bn_stats  (per-CTA channel sums and sums of squares; or the convolution's epilogue writes the same
//                        partials, PCDB_EPI_STATS) -> bn_finalize (partials added in index order in fp64, mean /
//                        1/std, running statistics, folded scale / shift) -> bn_apply (x = relu(y * scale + shift))
//   backward  dz = grad_out * [x > 0];  bn_bwd_reduce (sum dz, sum dz * xhat) -> bn_bwd_finalize (grad_gamma,
//             grad_beta, coefficients) -> bn_bwd_apply (grad_y = gamma / std * (dz - mean(dz) - xhat * mean(dz * xhat)))
//
// No atomics: partials are reduced in a fixed order, so a step is bit-reproducible.
#include "common.cuh"
#include "../../include/pcdet_b200.h"

namespace pcdb {
namespace bn {

constexpr int kBlock = 256;
constexpr int kMaxBlocks = kNumSMs;
constexpr int kMaxC = 128;

template <typename T> struct Vec8;
template <> struct Vec8<__nv_bfloat16> {
    static __device__ __forceinline__ void load(const __nv_bfloat16 *p, float *v)
    {
        const uint4 raw = *reinterpret_cast<const uint4 *>(p);
        const __nv_bfloat162 *h = reinterpret_cast<const __nv_bfloat162 *>(&raw);
#pragma unroll
        for (int j = 0; j < 4; ++j) { const float2 f = __bfloat1622float2(h[j]); v[2 * j] = f.x; v[2 * j + 1] = f.y; }
    }
    static __device__ __forceinline__ void store(__nv_bfloat16 *p, const float *v)
    {
        uint4 raw;
        __nv_bfloat162 *h = reinterpret_cast<__nv_bfloat162 *>(&raw);
#pragma unroll
        for (int j = 0; j < 4; ++j) h[j] = __floats2bfloat162_rn(v[2 * j], v[2 * j + 1]);
        *reinterpret_cast<uint4 *>(p) = raw;
    }
};
template <> struct Vec8<float> {
    static __device__ __forceinline__ void load(const float *p, float *v)
    {
        const float4 a = reinterpret_cast<const float4 *>(p)[0], b = reinterpret_cast<const float4 *>(p)[1];
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    }
    static __device__ __forceinline__ void store(float *p, const float *v)
    {
        reinterpret_cast<float4 *>(p)[0] = make_float4(v[0], v[1], v[2], v[3]);
        reinterpret_cast<float4 *>(p)[1] = make_float4(v[4], v[5], v[6], v[7]);
    }
};

__device__ __forceinline__ int rows_of(int n, const int *n_dev)
{
    if (n_dev) { const int m = __ldg(n_dev); return m < n ? m : n; }
    return n;
}

// Sum over the row lanes of a block: thread t owns the 8 channels of chunk t % cpr in row lane t / cpr.  a, b: the
// thread's 8 + 8 partial sums.  Leaves the block's totals in partial[blockIdx.x][0..1][c].
__device__ __forceinline__ void block_channel_sums(const float *a, const float *b, int c, float *partial)
{
    __shared__ float red[2][8][kBlock];
    const int cpr = c / 8, lanes = kBlock / cpr;
#pragma unroll
    for (int j = 0; j < 8; ++j) { red[0][j][threadIdx.x] = a[j]; red[1][j][threadIdx.x] = b[j]; }
    __syncthreads();
    for (int o = threadIdx.x; o < 2 * c; o += kBlock) {
        const int which = o / c, ch = o % c, chunk = ch / 8, j = ch % 8;
        float s = 0.f;
        for (int l = 0; l < lanes; ++l) s += red[which][j][l * cpr + chunk];
        partial[((size_t)blockIdx.x * 2 + which) * c + ch] = s;
    }
}

// Sum of partial[p][which][ch] over p for both `which` by one WARP per channel: lane l adds the blocks p = l, l + 32, ... (at
// most kMaxBlocks / 32 = 5, all loads in flight at once) in fp64, then a butterfly over the lanes -- a fixed order, so the
// result does not depend on scheduling.  blockDim = (32, kFinChannels), grid = ceil(c / kFinChannels); returns false for the
// threads that have nothing more to do (lanes other than 0, channels past c).
constexpr int kFinLanes = 32;
constexpr int kFinChannels = 16;
__device__ __forceinline__ bool reduce_partials(const float *__restrict__ partial, int n_partials, int c, int &ch, double &s, double &q)
{
    const int lane = threadIdx.x;
    ch = blockIdx.x * kFinChannels + threadIdx.y;
    const int chc = ch < c ? ch : c - 1;
    constexpr int kPer = (kMaxBlocks + kFinLanes - 1) / kFinLanes;       // 5: one round covers the kernels' own partials
    double a = 0.0, b = 0.0;
    for (int base = 0; base < n_partials; base += kPer * kFinLanes) {     // (more rounds only for a convolution's per-tile partials)
        float va[kPer], vb[kPer];
#pragma unroll
        for (int j = 0; j < kPer; ++j) {
            const int p = base + lane + j * kFinLanes;
            va[j] = p < n_partials ? __ldg(partial + ((size_t)p * 2 + 0) * c + chc) : 0.f;
            vb[j] = p < n_partials ? __ldg(partial + ((size_t)p * 2 + 1) * c + chc) : 0.f;
        }
#pragma unroll
        for (int j = 0; j < kPer; ++j) { a += (double)va[j]; b += (double)vb[j]; }
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        a += __shfl_xor_sync(0xffffffffu, a, d);
        b += __shfl_xor_sync(0xffffffffu, b, d);
    }
    s = a; q = b;
    return lane == 0 && ch < c;
}

template <typename T>
__global__ void __launch_bounds__(kBlock) bn_stats_kernel(const T *__restrict__ y, int n, const int *__restrict__ n_dev, int c,
                                                          float *__restrict__ partial)
{
    n = rows_of(n, n_dev);
    const int cpr = c / 8, lanes = kBlock / cpr;
    const int chunk = threadIdx.x % cpr, lane = threadIdx.x / cpr;
    float s[8] = {0, 0, 0, 0, 0, 0, 0, 0}, q[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (lane < lanes)
        for (int r = blockIdx.x * lanes + lane; r < n; r += gridDim.x * lanes) {
            float v[8];
            Vec8<T>::load(y + (size_t)r * c + chunk * 8, v);
#pragma unroll
            for (int j = 0; j < 8; ++j) { s[j] += v[j]; q[j] = fmaf(v[j], v[j], q[j]); }
        }
    block_channel_sums(s, q, c, partial);
}

// One warp per channel (see reduce_partials).  stats: [0] mean, [1] 1/std, [2] scale = gamma/std, [3] shift = beta - mean*scale.
// sums (optional, SyncBatchNorm): channel sums, sums of squares and the row count of ALL ranks ((2c + 1) doubles, all-reduced by
// the caller) replace the local partials.
__global__ void bn_finalize_kernel(const float *__restrict__ partial, int n_partials, int n, const int *__restrict__ n_dev, int c,
                                   const double *__restrict__ sums, const float *__restrict__ gamma, const float *__restrict__ beta,
                                   float eps, float momentum, float *__restrict__ running_mean, float *__restrict__ running_var,
                                   float *__restrict__ stats)
{
    n = rows_of(n, n_dev);
    int ch;
    double s, q;
    if (sums) {
        ch = blockIdx.x * kFinChannels + threadIdx.y;
        if (threadIdx.x != 0 || ch >= c) return;
        s = sums[ch]; q = sums[c + ch];
        n = (int)sums[2 * c];
    } else if (!reduce_partials(partial, n_partials, c, ch, s, q)) {
        return;
    }
    const double cnt = n > 0 ? (double)n : 1.0;
    const double mean = s / cnt;
    double var = q / cnt - mean * mean;
    var = var < 0.0 ? 0.0 : var;
    const float invstd = (float)(1.0 / sqrt(var + (double)eps));
    const float g = gamma ? gamma[ch] : 1.f, b = beta ? beta[ch] : 0.f;
    stats[ch] = (float)mean;
    stats[c + ch] = invstd;
    stats[2 * c + ch] = g * invstd;
    stats[3 * c + ch] = b - (float)mean * g * invstd;
    if (running_mean && n > 0) running_mean[ch] = (1.f - momentum) * running_mean[ch] + momentum * (float)mean;
    if (running_var && n > 0) {
        const double unbiased = n > 1 ? var * cnt / (cnt - 1.0) : var;
        running_var[ch] = (1.f - momentum) * running_var[ch] + momentum * (float)unbiased;
    }
}

template <typename T>
__global__ void __launch_bounds__(kBlock) bn_apply_kernel(const T *__restrict__ y, int n, const int *__restrict__ n_dev, int c,
                                                          const float *__restrict__ stats, int relu, T *__restrict__ out)
{
    n = rows_of(n, n_dev);
    const int cpr = c / 8;
    const size_t total = (size_t)n * cpr;
    for (size_t i = (size_t)blockIdx.x * kBlock + threadIdx.x; i < total; i += (size_t)gridDim.x * kBlock) {
        const int chunk = (int)(i % cpr);
        float v[8];
        Vec8<T>::load(y + i * 8, v);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float x = fmaf(v[j], __ldg(stats + 2 * c + chunk * 8 + j), __ldg(stats + 3 * c + chunk * 8 + j));
            v[j] = relu ? fmaxf(x, 0.f) : x;
        }
        Vec8<T>::store(out + i * 8, v);
    }
}

// partial[b][0][ch] = sum dz, partial[b][1][ch] = sum dz * xhat over the block's rows
template <typename T>
__global__ void __launch_bounds__(kBlock) bn_bwd_reduce_kernel(const T *__restrict__ grad_out, const T *__restrict__ out,
                                                               const T *__restrict__ y, int n, const int *__restrict__ n_dev, int c,
                                                               const float *__restrict__ stats, int relu, float *__restrict__ partial)
{
    n = rows_of(n, n_dev);
    const int cpr = c / 8, lanes = kBlock / cpr;
    const int chunk = threadIdx.x % cpr, lane = threadIdx.x / cpr;
    float s[8] = {0, 0, 0, 0, 0, 0, 0, 0}, q[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    float mean[8], invstd[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { mean[j] = stats[chunk * 8 + j]; invstd[j] = stats[c + chunk * 8 + j]; }
    if (lane < lanes)
        for (int r = blockIdx.x * lanes + lane; r < n; r += gridDim.x * lanes) {
            float g[8], o[8], v[8];
            const size_t at = (size_t)r * c + chunk * 8;
            Vec8<T>::load(grad_out + at, g);
            Vec8<T>::load(y + at, v);
            if (relu) Vec8<T>::load(out + at, o);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float dz = (!relu || o[j] > 0.f) ? g[j] : 0.f;
                s[j] += dz;
                q[j] = fmaf(dz, (v[j] - mean[j]) * invstd[j], q[j]);
            }
        }
    block_channel_sums(s, q, c, partial);
}

// Local channel sums of a rank as doubles: forward (sum y, sum y^2, row count) or backward (sum dz, sum dz * xhat)
__global__ void bn_sums_kernel(const float *__restrict__ partial, int n_partials, int n, const int *__restrict__ n_dev, int c,
                               int with_count, double *__restrict__ sums)
{
    n = rows_of(n, n_dev);
    int ch;
    double s, q;
    if (!reduce_partials(partial, n_partials, c, ch, s, q)) return;
    sums[ch] = s;
    sums[c + ch] = q;
    if (with_count && ch == 0) sums[2 * c] = (double)n;
}

// coef: [0] a = gamma/std, [1] b = mean(dz), [2] cc = mean(dz * xhat).  SyncBatchNorm: local2 = this rank's sums (they make
// grad_gamma / grad_beta, which the gradient all-reduce averages later), global2 = the all-reduced sums and fwd_sums[2c] the
// global row count (they make the input gradient) -- torch.nn.SyncBatchNorm's backward.
__global__ void bn_bwd_finalize_kernel(const float *__restrict__ partial, int n_partials, int n, const int *__restrict__ n_dev, int c,
                                       const double *__restrict__ local2, const double *__restrict__ global2,
                                       const double *__restrict__ fwd_sums, const float *__restrict__ gamma,
                                       const float *__restrict__ stats, int accumulate, float *__restrict__ grad_gamma,
                                       float *__restrict__ grad_beta, float *__restrict__ coef)
{
    n = rows_of(n, n_dev);
    int ch;
    double s, q, sg, qg;
    if (local2) {
        ch = blockIdx.x * kFinChannels + threadIdx.y;
        if (threadIdx.x != 0 || ch >= c) return;
        s = local2[ch]; q = local2[c + ch];
        sg = global2[ch]; qg = global2[c + ch];
        n = (int)fwd_sums[2 * c];
    } else {
        if (!reduce_partials(partial, n_partials, c, ch, s, q)) return;
        sg = s; qg = q;
    }
    const double cnt = n > 0 ? (double)n : 1.0;
    coef[ch] = (gamma ? gamma[ch] : 1.f) * stats[c + ch];
    coef[c + ch] = (float)(sg / cnt);
    coef[2 * c + ch] = (float)(qg / cnt);
    if (grad_beta) grad_beta[ch] = (accumulate ? grad_beta[ch] : 0.f) + (float)s;
    if (grad_gamma) grad_gamma[ch] = (accumulate ? grad_gamma[ch] : 0.f) + (float)q;
}

template <typename T>
__global__ void __launch_bounds__(kBlock) bn_bwd_apply_kernel(const T *__restrict__ grad_out, const T *__restrict__ out,
                                                              const T *__restrict__ y, int n, const int *__restrict__ n_dev, int c,
                                                              const float *__restrict__ stats, const float *__restrict__ coef, int relu,
                                                              T *__restrict__ grad_y)
{
    n = rows_of(n, n_dev);
    const int cpr = c / 8;
    const size_t total = (size_t)n * cpr;
    for (size_t i = (size_t)blockIdx.x * kBlock + threadIdx.x; i < total; i += (size_t)gridDim.x * kBlock) {
        const int chunk = (int)(i % cpr);
        float g[8], o[8], v[8];
        Vec8<T>::load(grad_out + i * 8, g);
        Vec8<T>::load(y + i * 8, v);
        if (relu) Vec8<T>::load(out + i * 8, o);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int ch = chunk * 8 + j;
            const float dz = (!relu || o[j] > 0.f) ? g[j] : 0.f;
            const float xhat = (v[j] - __ldg(stats + ch)) * __ldg(stats + c + ch);
            g[j] = __ldg(coef + ch) * (dz - __ldg(coef + c + ch) - xhat * __ldg(coef + 2 * c + ch));
        }
        Vec8<T>::store(grad_y + i * 8, g);
    }
}

inline int stat_blocks(int n, int c)
{
    const int lanes = kBlock / (c / 8);
    int b = (n + lanes * 8 - 1) / (lanes * 8);          // >= 8 rows per row lane
    return b < 1 ? 1 : (b > kMaxBlocks ? kMaxBlocks : b);
}
inline int apply_blocks(int n, int c)
{
    const size_t total = (size_t)n * (c / 8);
    size_t b = (total + kBlock * 4 - 1) / (kBlock * 4);
    return b < 1 ? 1 : (b > (size_t)8 * kNumSMs ? 8 * kNumSMs : (int)b);
}

template <typename T>
int fwd(const void *y, int n, const int *n_dev, int c, const float *gamma, const float *beta, float eps, float momentum,
        float *running_mean, float *running_var, int flags, void *out, float *stats, const float *conv_partials, int n_conv_partials,
        float *ws, cudaStream_t stream)
{
    const float *partial = conv_partials;
    int n_partials = n_conv_partials;
    if (!partial) {
        n_partials = stat_blocks(n, c);
        bn_stats_kernel<T><<<n_partials, kBlock, 0, stream>>>((const T *)y, n, n_dev, c, ws);
        partial = ws;
    }
    bn_finalize_kernel<<<(c + kFinChannels - 1) / kFinChannels, dim3(kFinLanes, kFinChannels), 0, stream>>>(partial, n_partials, n, n_dev, c, nullptr, gamma, beta, eps, momentum, running_mean, running_var, stats);
    if (out && n > 0)
        bn_apply_kernel<T><<<apply_blocks(n, c), kBlock, 0, stream>>>((const T *)y, n, n_dev, c, stats, flags & PCDB_EPI_RELU, (T *)out);
    return check_launch("pcdb_bn_train_fwd");
}

template <typename T>
int bwd(const void *grad_out, const void *out, const void *y, int n, const int *n_dev, int c, const float *gamma, const float *stats,
        int flags, void *grad_y, float *grad_gamma, float *grad_beta, int accumulate, float *ws, cudaStream_t stream)
{
    const int relu = flags & PCDB_EPI_RELU;
    const int n_partials = stat_blocks(n, c);
    float *coef = ws + (size_t)2 * kMaxC * kMaxBlocks;
    bn_bwd_reduce_kernel<T><<<n_partials, kBlock, 0, stream>>>((const T *)grad_out, (const T *)out, (const T *)y, n, n_dev, c, stats, relu, ws);
    bn_bwd_finalize_kernel<<<(c + kFinChannels - 1) / kFinChannels, dim3(kFinLanes, kFinChannels), 0, stream>>>(ws, n_partials, n, n_dev, c, nullptr, nullptr, nullptr, gamma, stats, accumulate, grad_gamma, grad_beta, coef);
    if (n > 0)
        bn_bwd_apply_kernel<T><<<apply_blocks(n, c), kBlock, 0, stream>>>((const T *)grad_out, (const T *)out, (const T *)y, n, n_dev, c, stats,
                                                                         coef, relu, (T *)grad_y);
    return check_launch("pcdb_bn_train_bwd");
}

// ---- SyncBatchNorm (tools/train.py:94-95): the same kernels in two halves, the caller all-reduces the sums in between ----
template <typename T>
int fwd_sums(const void *y, int n, const int *n_dev, int c, double *sums, float *ws, cudaStream_t stream)
{
    const int n_partials = stat_blocks(n, c);
    bn_stats_kernel<T><<<n_partials, kBlock, 0, stream>>>((const T *)y, n, n_dev, c, ws);
    bn_sums_kernel<<<(c + kFinChannels - 1) / kFinChannels, dim3(kFinLanes, kFinChannels), 0, stream>>>(ws, n_partials, n, n_dev, c, 1, sums);
    return check_launch("pcdb_bn_train_sums");
}

template <typename T>
int fwd_from_sums(const void *y, int n, const int *n_dev, int c, const double *sums, const float *gamma, const float *beta, float eps,
                  float momentum, float *running_mean, float *running_var, int flags, void *out, float *stats, cudaStream_t stream)
{
    bn_finalize_kernel<<<(c + kFinChannels - 1) / kFinChannels, dim3(kFinLanes, kFinChannels), 0, stream>>>(nullptr, 0, n, n_dev, c, sums, gamma, beta, eps, momentum, running_mean, running_var, stats);
    if (out && n > 0)
        bn_apply_kernel<T><<<apply_blocks(n, c), kBlock, 0, stream>>>((const T *)y, n, n_dev, c, stats, flags & PCDB_EPI_RELU, (T *)out);
    return check_launch("pcdb_bn_train_fwd_from_sums");
}

template <typename T>
int bwd_sums(const void *grad_out, const void *out, const void *y, int n, const int *n_dev, int c, const float *stats, int flags,
             double *sums2, float *ws, cudaStream_t stream)
{
    const int n_partials = stat_blocks(n, c);
    bn_bwd_reduce_kernel<T><<<n_partials, kBlock, 0, stream>>>((const T *)grad_out, (const T *)out, (const T *)y, n, n_dev, c, stats,
                                                              flags & PCDB_EPI_RELU, ws);
    bn_sums_kernel<<<(c + kFinChannels - 1) / kFinChannels, dim3(kFinLanes, kFinChannels), 0, stream>>>(ws, n_partials, n, n_dev, c, 0, sums2);
    return check_launch("pcdb_bn_train_bwd_sums");
}

template <typename T>
int bwd_from_sums(const void *grad_out, const void *out, const void *y, int n, const int *n_dev, int c, const float *gamma,
                  const float *stats, const double *local2, const double *global2, const double *fwd, int flags, void *grad_y,
                  float *grad_gamma, float *grad_beta, int accumulate, float *ws, cudaStream_t stream)
{
    float *coef = ws + (size_t)2 * kMaxC * kMaxBlocks;
    bn_bwd_finalize_kernel<<<(c + kFinChannels - 1) / kFinChannels, dim3(kFinLanes, kFinChannels), 0, stream>>>(nullptr, 0, n, n_dev, c, local2, global2, fwd, gamma, stats, accumulate,
                                                                 grad_gamma, grad_beta, coef);
    if (n > 0)
        bn_bwd_apply_kernel<T><<<apply_blocks(n, c), kBlock, 0, stream>>>((const T *)grad_out, (const T *)out, (const T *)y, n, n_dev, c, stats,
                                                                         coef, flags & PCDB_EPI_RELU, (T *)grad_y);
    return check_launch("pcdb_bn_train_bwd_from_sums");
}

}  // namespace bn
}  // namespace pcdb

using namespace pcdb;

extern "C" size_t pcdb_bn_train_workspace_bytes(void)
{
    return ((size_t)2 * bn::kMaxC * bn::kMaxBlocks + 3 * bn::kMaxC) * sizeof(float);
}

static bool bn_args_ok(const char *what, int n, int c, int dtype, const void *ws, size_t ws_bytes)
{
    if (n < 0 || c < 8 || c > bn::kMaxC || c % 8 != 0 || (dtype != PCDB_F32 && dtype != PCDB_BF16) || !ws ||
        ws_bytes < pcdb_bn_train_workspace_bytes()) {
        set_last_error("%s: invalid argument (n=%d, c=%d must be a multiple of 8 in [8,128], dtype=%d, workspace %zu of %zu bytes)",
                       what, n, c, dtype, ws_bytes, pcdb_bn_train_workspace_bytes());
        return false;
    }
    return true;
}

extern "C" int pcdb_bn_train_fwd(const void *y, int n, const int32_t *n_dev, int c, int dtype, const float *gamma, const float *beta,
                                 float eps, float momentum, float *running_mean, float *running_var, int flags, void *out,
                                 float *stats, const float *conv_partials, int n_conv_partials, void *workspace,
                                 size_t workspace_bytes, void *stream)
{
    if (!y || !stats || (conv_partials && n_conv_partials < 1)) {
        set_last_error("pcdb_bn_train_fwd: null argument");
        return kInvalidArgument;
    }
    if (!bn_args_ok("pcdb_bn_train_fwd", n, c, dtype, workspace, workspace_bytes)) return kInvalidArgument;
    if (dtype == PCDB_BF16)
        return bn::fwd<__nv_bfloat16>(y, n, n_dev, c, gamma, beta, eps, momentum, running_mean, running_var, flags, out, stats,
                                      conv_partials, n_conv_partials, (float *)workspace, (cudaStream_t)stream);
    return bn::fwd<float>(y, n, n_dev, c, gamma, beta, eps, momentum, running_mean, running_var, flags, out, stats, conv_partials,
                          n_conv_partials, (float *)workspace, (cudaStream_t)stream);
}

extern "C" int pcdb_bn_train_bwd(const void *grad_out, const void *out, const void *y, int n, const int32_t *n_dev, int c, int dtype,
                                 const float *gamma, const float *stats, int flags, void *grad_y, float *grad_gamma,
                                 float *grad_beta, int accumulate, void *workspace, size_t workspace_bytes, void *stream)
{
    if (!bn_args_ok("pcdb_bn_train_bwd", n, c, dtype, workspace, workspace_bytes)) return kInvalidArgument;
    if (!grad_out || !y || !stats || !grad_y || ((flags & PCDB_EPI_RELU) && !out)) {
        set_last_error("pcdb_bn_train_bwd: null argument");
        return kInvalidArgument;
    }
    if (dtype == PCDB_BF16)
        return bn::bwd<__nv_bfloat16>(grad_out, out, y, n, n_dev, c, gamma, stats, flags, grad_y, grad_gamma, grad_beta, accumulate,
                                      (float *)workspace, (cudaStream_t)stream);
    return bn::bwd<float>(grad_out, out, y, n, n_dev, c, gamma, stats, flags, grad_y, grad_gamma, grad_beta, accumulate,
                          (float *)workspace, (cudaStream_t)stream);
}

// ---- SyncBatchNorm halves ---------------------------------------------------------------------------------------------------
#define PCDB_BN_DISPATCH(call_bf16, call_f32) (dtype == PCDB_BF16 ? (call_bf16) : (call_f32))

extern "C" int pcdb_bn_train_sums(const void *y, int n, const int32_t *n_dev, int c, int dtype, double *sums, void *workspace,
                                  size_t workspace_bytes, void *stream)
{
    if (!y || !sums) { set_last_error("pcdb_bn_train_sums: null argument"); return kInvalidArgument; }
    if (!bn_args_ok("pcdb_bn_train_sums", n, c, dtype, workspace, workspace_bytes)) return kInvalidArgument;
    return PCDB_BN_DISPATCH(bn::fwd_sums<__nv_bfloat16>(y, n, n_dev, c, sums, (float *)workspace, (cudaStream_t)stream),
                            bn::fwd_sums<float>(y, n, n_dev, c, sums, (float *)workspace, (cudaStream_t)stream));
}

extern "C" int pcdb_bn_train_fwd_from_sums(const void *y, int n, const int32_t *n_dev, int c, int dtype, const double *sums,
                                           const float *gamma, const float *beta, float eps, float momentum, float *running_mean,
                                           float *running_var, int flags, void *out, float *stats, void *stream)
{
    if (!y || !sums || !stats || n < 0 || c < 8 || c > bn::kMaxC || c % 8 != 0 || (dtype != PCDB_F32 && dtype != PCDB_BF16)) {
        set_last_error("pcdb_bn_train_fwd_from_sums: invalid argument (n=%d c=%d dtype=%d)", n, c, dtype);
        return kInvalidArgument;
    }
    return PCDB_BN_DISPATCH(bn::fwd_from_sums<__nv_bfloat16>(y, n, n_dev, c, sums, gamma, beta, eps, momentum, running_mean, running_var, flags, out, stats, (cudaStream_t)stream),
                            bn::fwd_from_sums<float>(y, n, n_dev, c, sums, gamma, beta, eps, momentum, running_mean, running_var, flags, out, stats, (cudaStream_t)stream));
}

extern "C" int pcdb_bn_train_bwd_sums(const void *grad_out, const void *out, const void *y, int n, const int32_t *n_dev, int c,
                                      int dtype, const float *stats, int flags, double *sums2, void *workspace,
                                      size_t workspace_bytes, void *stream)
{
    if (!grad_out || !y || !stats || !sums2 || ((flags & PCDB_EPI_RELU) && !out)) { set_last_error("pcdb_bn_train_bwd_sums: null argument"); return kInvalidArgument; }
    if (!bn_args_ok("pcdb_bn_train_bwd_sums", n, c, dtype, workspace, workspace_bytes)) return kInvalidArgument;
    return PCDB_BN_DISPATCH(bn::bwd_sums<__nv_bfloat16>(grad_out, out, y, n, n_dev, c, stats, flags, sums2, (float *)workspace, (cudaStream_t)stream),
                            bn::bwd_sums<float>(grad_out, out, y, n, n_dev, c, stats, flags, sums2, (float *)workspace, (cudaStream_t)stream));
}

extern "C" int pcdb_bn_train_bwd_from_sums(const void *grad_out, const void *out, const void *y, int n, const int32_t *n_dev, int c,
                                           int dtype, const float *gamma, const float *stats, const double *local_sums2,
                                           const double *global_sums2, const double *fwd_sums, int flags, void *grad_y,
                                           float *grad_gamma, float *grad_beta, int accumulate, void *workspace,
                                           size_t workspace_bytes, void *stream)
{
    if (!grad_out || !y || !stats || !grad_y || !local_sums2 || !global_sums2 || !fwd_sums || ((flags & PCDB_EPI_RELU) && !out)) {
        set_last_error("pcdb_bn_train_bwd_from_sums: null argument");
        return kInvalidArgument;
    }
    if (!bn_args_ok("pcdb_bn_train_bwd_from_sums", n, c, dtype, workspace, workspace_bytes)) return kInvalidArgument;
    return PCDB_BN_DISPATCH(bn::bwd_from_sums<__nv_bfloat16>(grad_out, out, y, n, n_dev, c, gamma, stats, local_sums2, global_sums2, fwd_sums, flags, grad_y, grad_gamma, grad_beta, accumulate, (float *)workspace, (cudaStream_t)stream),
                            bn::bwd_from_sums<float>(grad_out, out, y, n, n_dev, c, gamma, stats, local_sums2, global_sums2, fwd_sums, flags, grad_y, grad_gamma, grad_beta, accumulate, (float *)workspace, (cudaStream_t)stream));
}
