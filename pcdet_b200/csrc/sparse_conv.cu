// Dispatcher for the sparse convolution forward pass (C-ABI entry point).
#include "common.cuh"
#include "../../include/pcdet_b200.h"

namespace pcdb {
int launch_conv_fwd_simt(const void *features, const void *weight, const int32_t *nbr, int ld, int K, int n_out,
                         const int32_t *n_out_dev, int c_in, int c_out, int dtype, const float *scale,
                         const float *shift, const float *bias, int flags, void *out, cudaStream_t stream);
// returns kUnsupported when the shape is not covered by the tensor-core kernel
int launch_conv_fwd_tc(const void *features, const void *weight, const int32_t *nbr, int ld, int K, int n_out,
                       const int32_t *n_out_dev, int c_in, int c_out, const float *scale, const float *shift,
                       const float *bias, int flags, void *out, cudaStream_t stream);
bool conv_tc_supported(int c_in, int c_out, int K);
}  // namespace pcdb

using namespace pcdb;

extern "C" int pcdb_sparse_conv_fwd(const void *features, const void *weight, const int32_t *nbr, int ld,
                                    int kernel_volume, int n_out, const int32_t *n_out_dev, int c_in, int c_out,
                                    int dtype, const float *scale, const float *shift, const float *bias,
                                    int flags, void *out, int algo, void *stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    if (!features || !weight || !nbr || !out || n_out < 0 || c_in < 1 || c_out < 1 || kernel_volume < 1 || ld < n_out ||
        (dtype != PCDB_F32 && dtype != PCDB_BF16)) {
        set_last_error("pcdb_sparse_conv_fwd: invalid argument (n_out=%d c_in=%d c_out=%d K=%d ld=%d dtype=%d)",
                       n_out, c_in, c_out, kernel_volume, ld, dtype);
        return kInvalidArgument;
    }
    if (n_out == 0) return kOk;
    const bool transposed = (flags & PCDB_WEIGHT_TRANSPOSED) != 0;
    const bool tc_ok = dtype == PCDB_BF16 && transposed && conv_tc_supported(c_in, c_out, kernel_volume);
    if (algo == 2 && !tc_ok) {
        set_last_error("pcdb_sparse_conv_fwd: tcgen05 kernel does not cover c_in=%d c_out=%d dtype=%d", c_in, c_out, dtype);
        return kUnsupported;
    }
    if (transposed && (!tc_ok || algo == 1)) {
        set_last_error("pcdb_sparse_conv_fwd: (K,c_out,c_in) weights are only consumed by the tcgen05 kernel "
                       "(bf16, c_in in {16,32,64}, c_out in {16,32,64,128}); got c_in=%d c_out=%d dtype=%d algo=%d",
                       c_in, c_out, dtype, algo);
        return kUnsupported;
    }
    if (tc_ok && algo != 1)
        return launch_conv_fwd_tc(features, weight, nbr, ld, kernel_volume, n_out, n_out_dev, c_in, c_out, scale, shift,
                                  bias, flags, out, stream);
    return launch_conv_fwd_simt(features, weight, nbr, ld, kernel_volume, n_out, n_out_dev, c_in, c_out, dtype, scale,
                                shift, bias, flags, out, stream);
}
