// Dispatcher for the sparse convolution forward pass (C-ABI entry points).
#include "common.cuh"
#include "../../include/pcdet_b200.h"

namespace pcdb {
int launch_conv_fwd_simt(const void *features, const void *weight, const int32_t *nbr, int ld, int K, int n_out,
                         const int32_t *n_out_dev, int c_in, int c_out, int dtype, const float *scale,
                         const float *shift, const float *bias, int flags, void *out, cudaStream_t stream);
int launch_conv_fwd_tc(const void *features, int n_in, const void *w_packed, const int32_t *nbr, int ld, int K, int n_out,
                       const int32_t *n_out_dev, int c_in, int c_out, const float *scale, const float *shift,
                       const float *bias, const void *residual, int flags, void *out, bool use_tma, int rows_hint, cudaStream_t stream);
bool conv_tc_supported(int c_in, int c_out, int K);
size_t conv_tc_packed_bytes(int c_in, int c_out, int K);
int conv_tc_pack_weights(const void *weight, int dtype, int K, int c_in, int c_out, int flags, void *packed, cudaStream_t stream);
}  // namespace pcdb

using namespace pcdb;

extern "C" size_t pcdb_conv_packed_weight_bytes(int kernel_volume, int c_in, int c_out)
{
    if (!conv_tc_supported(c_in, c_out, kernel_volume)) return 0;
    return conv_tc_packed_bytes(c_in, c_out, kernel_volume);
}

extern "C" int pcdb_pack_conv_weights(const void *weight, int kernel_volume, int c_in, int c_out, void *packed, void *stream)
{
    if (!weight || !packed || !conv_tc_supported(c_in, c_out, kernel_volume)) {
        set_last_error("pcdb_pack_conv_weights: unsupported shape K=%d c_in=%d c_out=%d", kernel_volume, c_in, c_out);
        return kUnsupported;
    }
    return conv_tc_pack_weights(weight, PCDB_BF16, kernel_volume, c_in, c_out, 0, packed, (cudaStream_t)stream);
}

extern "C" int pcdb_pack_conv_weights_ex(const void *weight, int dtype, int kernel_volume, int c_in, int c_out, int flags,
                                         void *packed, void *stream)
{
    if (!weight || !packed || !conv_tc_supported(c_in, c_out, kernel_volume) || (dtype != PCDB_F32 && dtype != PCDB_BF16)) {
        set_last_error("pcdb_pack_conv_weights_ex: unsupported shape K=%d c_in=%d c_out=%d dtype=%d", kernel_volume, c_in, c_out, dtype);
        return kUnsupported;
    }
    return conv_tc_pack_weights(weight, dtype, kernel_volume, c_in, c_out, flags, packed, (cudaStream_t)stream);
}

extern "C" int pcdb_sparse_conv_fwd(const void *features, int n_in, const void *weight, const int32_t *nbr, int ld,
                                    int kernel_volume, int n_out, const int32_t *n_out_dev, int c_in, int c_out,
                                    int dtype, const float *scale, const float *shift, const float *bias,
                                    int flags, void *out, int algo, void *stream_)
{
    return pcdb_sparse_conv_fwd_ex(features, n_in, weight, nbr, ld, kernel_volume, n_out, n_out_dev, c_in, c_out, dtype, scale, shift,
                                   bias, nullptr, flags, out, algo, stream_);
}

extern "C" int pcdb_sparse_conv_fwd_ex(const void *features, int n_in, const void *weight, const int32_t *nbr, int ld,
                                       int kernel_volume, int n_out, const int32_t *n_out_dev, int c_in, int c_out,
                                       int dtype, const float *scale, const float *shift, const float *bias,
                                       const void *residual, int flags, void *out, int algo, void *stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    if (!features || !weight || !nbr || !out || n_out < 0 || n_in < 0 || c_in < 1 || c_out < 1 || kernel_volume < 1 ||
        ld < n_out || (dtype != PCDB_F32 && dtype != PCDB_BF16)) {
        set_last_error("pcdb_sparse_conv_fwd: invalid argument (n_in=%d n_out=%d c_in=%d c_out=%d K=%d ld=%d dtype=%d)",
                       n_in, n_out, c_in, c_out, kernel_volume, ld, dtype);
        return kInvalidArgument;
    }
    if (n_out == 0) return kOk;
    const bool packed = (flags & PCDB_WEIGHT_PACKED) != 0;
    const bool tc_ok = dtype == PCDB_BF16 && packed && conv_tc_supported(c_in, c_out, kernel_volume);
    const int rows_hint = algo >> 8;       // bits 8..: expected output rows when n_out is a capacity (0 = n_out)
    algo &= 0xff;
    if ((algo >= 2 && !tc_ok) || (packed && (!tc_ok || algo == 1))) {
        set_last_error("pcdb_sparse_conv_fwd: the tcgen05 kernels take bf16, packed weights (pcdb_pack_conv_weights), c_in in "
                       "{16,32,64}, c_out in {16,32,64,128}; got c_in=%d c_out=%d dtype=%d packed=%d algo=%d",
                       c_in, c_out, dtype, (int)packed, algo);
        return kUnsupported;
    }
    if (residual && !tc_ok) {
        set_last_error("pcdb_sparse_conv_fwd_ex: the residual epilogue exists in the tcgen05 kernels only (bf16, packed weights)");
        return kUnsupported;
    }
    if (tc_ok)
        return launch_conv_fwd_tc(features, n_in, weight, nbr, ld, kernel_volume, n_out, n_out_dev, c_in, c_out, scale, shift,
                                  bias, residual, flags, out, /*use_tma=*/algo == 2, rows_hint, stream);
    return launch_conv_fwd_simt(features, weight, nbr, ld, kernel_volume, n_out, n_out_dev, c_in, c_out, dtype, scale,
                                shift, bias, flags, out, stream);
}
