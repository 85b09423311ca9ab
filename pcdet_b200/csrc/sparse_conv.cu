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
    const bool tc_ok = dtype == PCDB_BF16 && conv_tc_supported(c_in, c_out, kernel_volume);
    if (algo == 2 && !tc_ok) {
        set_last_error("pcdb_sparse_conv_fwd: tcgen05 kernel does not cover c_in=%d c_out=%d dtype=%d", c_in, c_out, dtype);
        return kUnsupported;
    }
    if (tc_ok && algo != 1)
        return launch_conv_fwd_tc(features, weight, nbr, ld, kernel_volume, n_out, n_out_dev, c_in, c_out, scale, shift,
                                  bias, flags, out, stream);
    return launch_conv_fwd_simt(features, weight, nbr, ld, kernel_volume, n_out, n_out_dev, c_in, c_out, dtype, scale,
                                shift, bias, flags, out, stream);
}
