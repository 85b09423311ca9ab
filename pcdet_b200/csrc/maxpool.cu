// Sparse max pooling forward (spconv v1.0 indice_maxpool / maxPoolFwd*Kernel; used by
// pcdet/models/rcnn/partA2_rcnn_net.py:165).  Same output-stationary rulebook as the convolutions:
//   out[o, c] = max(0, max over offsets k with nbr[k][o] >= 0 of features[nbr[k][o], c])
// (the reference initialises the output with zeros and keeps a running maximum, hence the max with 0).
#include "common.cuh"
#include "../../include/pcdet_b200.h"

namespace pcdb {

template <typename T>
__global__ void __launch_bounds__(256)
maxpool_fwd_kernel(const T *__restrict__ feat, const int *__restrict__ nbr, int ld, int K, int n_out,
                   const int *__restrict__ n_out_dev, int c, T *__restrict__ out)
{
    if (n_out_dev) { const int m = __ldg(n_out_dev); n_out = m < n_out ? m : n_out; }
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)n_out * c) return;
    const int o = (int)(t / c), ch = (int)(t % c);
    float best = 0.f;
    for (int k = 0; k < K; ++k) {
        const int i = __ldg(nbr + (size_t)k * ld + o);
        if (i >= 0) best = fmaxf(best, to_float(feat[(size_t)i * c + ch]));
    }
    out[t] = from_float<T>(best);
}

// spconv v1.0 maxPoolBwd: every input that equals the pooled maximum receives the output's gradient (ties: all of them,
// once per offset through which they reach the output).  An output whose inputs are all negative pooled to 0 and passes
// nothing back.  grad_in is zeroed by the caller; one thread per (output row, channel), atomics on the inputs.
__global__ void __launch_bounds__(256)
maxpool_bwd_kernel(const float *__restrict__ feat, const float *__restrict__ out, const float *__restrict__ grad_out,
                   const int *__restrict__ nbr, int ld, int K, int n_out, int c, float *__restrict__ grad_in)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)n_out * c) return;
    const int o = (int)(t / c), ch = (int)(t % c);
    const float y = out[t], g = grad_out[t];
    for (int k = 0; k < K; ++k) {
        const int i = __ldg(nbr + (size_t)k * ld + o);
        if (i >= 0 && feat[(size_t)i * c + ch] == y) atomicAdd(grad_in + (size_t)i * c + ch, g);
    }
}

}  // namespace pcdb

using namespace pcdb;

extern "C" int pcdb_sparse_maxpool_bwd(const float *features, const float *out, const float *grad_out, const int32_t *nbr, int ld,
                                       int kernel_volume, int n_out, int c, float *grad_features, void *stream_)
{
    if (!features || !out || !grad_out || !nbr || !grad_features || n_out < 0 || c < 1 || kernel_volume < 1 || ld < n_out) {
        set_last_error("pcdb_sparse_maxpool_bwd: invalid argument (n_out=%d c=%d K=%d ld=%d)", n_out, c, kernel_volume, ld);
        return kInvalidArgument;
    }
    if (n_out == 0) return kOk;
    const long long total = (long long)n_out * c;
    maxpool_bwd_kernel<<<(int)((total + 255) / 256), 256, 0, (cudaStream_t)stream_>>>(features, out, grad_out, nbr, ld, kernel_volume,
                                                                                      n_out, c, grad_features);
    return check_launch("pcdb_sparse_maxpool_bwd");
}

extern "C" int pcdb_sparse_maxpool_fwd(const void *features, const int32_t *nbr, int ld, int kernel_volume, int n_out,
                                       const int32_t *n_out_dev, int c, int dtype, void *out, void *stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    if (!features || !nbr || !out || n_out < 0 || c < 1 || kernel_volume < 1 || ld < n_out ||
        (dtype != PCDB_F32 && dtype != PCDB_BF16)) {
        set_last_error("pcdb_sparse_maxpool_fwd: invalid argument (n_out=%d c=%d K=%d ld=%d dtype=%d)", n_out, c,
                       kernel_volume, ld, dtype);
        return kInvalidArgument;
    }
    if (n_out == 0) return kOk;
    const long long total = (long long)n_out * c;
    const int nb = (int)((total + 255) / 256);
    if (dtype == PCDB_BF16)
        maxpool_fwd_kernel<<<nb, 256, 0, stream>>>((const __nv_bfloat16 *)features, nbr, ld, kernel_volume, n_out, n_out_dev, c,
                                                   (__nv_bfloat16 *)out);
    else
        maxpool_fwd_kernel<<<nb, 256, 0, stream>>>((const float *)features, nbr, ld, kernel_volume, n_out, n_out_dev, c,
                                                   (float *)out);
    return check_launch("pcdb_sparse_maxpool_fwd");
}
