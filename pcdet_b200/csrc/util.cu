// Error reporting and tiny utility kernels shared by the C-ABI entry points.
#include "common.cuh"
#include "../../include/pcdet_b200.h"
#include <cstdarg>
#include <cstdio>

namespace pcdb {

static thread_local char g_last_error[512] = "";

void set_last_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_last_error, sizeof(g_last_error), fmt, ap);
    va_end(ap);
}

int check_launch(const char *what)
{
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) return kOk;
    set_last_error("%s: CUDA error %d (%s)", what, (int)e, cudaGetErrorString(e));
    return kCudaError;
}

__global__ void fill_i32_kernel(int *dst, int value, size_t count)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < count; i += stride) dst[i] = value;
}

void fill_i32(int *dst, int value, size_t count, cudaStream_t stream)
{
    if (count == 0) return;
    size_t blocks = (count + 255) / 256;
    if (blocks > (size_t)kNumSMs * 16) blocks = (size_t)kNumSMs * 16;
    fill_i32_kernel<<<(int)blocks, 256, 0, stream>>>(dst, value, count);
}

// dst[k*ld + r] = value for k < n_maps, r < min(*rows_dev, rows_cap); grid (row chunks of 1024, n_maps)
__global__ void __launch_bounds__(256)
fill_rows_i32_kernel(int *__restrict__ dst, int ld, const int *__restrict__ rows_dev, int rows_cap, int value)
{
    int n = __ldg(rows_dev);
    n = n < 0 ? 0 : (n < rows_cap ? n : rows_cap);
    const int r0 = (blockIdx.x * 256 + threadIdx.x) * 4;
    if (r0 >= n) return;
    int *p = dst + (size_t)blockIdx.y * ld + r0;
    if (r0 + 4 <= n && ((((size_t)blockIdx.y * ld) & 3) == 0)) {
        *reinterpret_cast<int4 *>(p) = make_int4(value, value, value, value);
    } else {
        for (int j = 0; j < 4 && r0 + j < n; ++j) p[j] = value;
    }
}

// nbr_inv[k][nbr[k][o]] = o for every pair of an output-stationary map: the same rulebook read from the input side
// (for a fixed offset an input row reaches at most one output row, so the writes never collide)
__global__ void __launch_bounds__(256)
invert_map_kernel(const int *__restrict__ nbr, int ld_out, int n_out, const int *__restrict__ n_out_dev, int *__restrict__ nbr_inv,
                  int ld_in, int n_in_cap)
{
    if (n_out_dev) { const int m = __ldg(n_out_dev); n_out = m < n_out ? m : n_out; }
    const int o = blockIdx.x * 256 + threadIdx.x;
    if (o >= n_out) return;
    const int i = __ldg(nbr + (size_t)blockIdx.y * ld_out + o);
    if (i >= 0 && i < n_in_cap) nbr_inv[(size_t)blockIdx.y * ld_in + i] = o;
}

}  // namespace pcdb

extern "C" int pcdb_rulebook_invert(const int32_t *nbr, int ld_out, int kernel_volume, int n_out, const int32_t *n_out_dev,
                                    int32_t *nbr_inv, int ld_in, int n_in, const int32_t *n_in_dev, void *stream)
{
    using namespace pcdb;
    if (!nbr || !nbr_inv || ld_out < n_out || ld_in < n_in || kernel_volume < 1 || kernel_volume > 65535 || n_out < 0 || n_in < 0 ||
        (((uintptr_t)nbr_inv) & 15)) {
        set_last_error("pcdb_rulebook_invert: invalid argument (K=%d n_out=%d ld_out=%d n_in=%d ld_in=%d)", kernel_volume, n_out, ld_out, n_in, ld_in);
        return kInvalidArgument;
    }
    if (n_in == 0) return kOk;
    if (n_in_dev)
        fill_rows_i32_kernel<<<dim3((n_in + 1023) / 1024, kernel_volume), 256, 0, (cudaStream_t)stream>>>(nbr_inv, ld_in, n_in_dev, n_in, -1);
    else
        for (int k = 0; k < kernel_volume; ++k) fill_i32(nbr_inv + (size_t)k * ld_in, -1, (size_t)n_in, (cudaStream_t)stream);
    if (n_out > 0)
        invert_map_kernel<<<dim3((n_out + 255) / 256, kernel_volume), 256, 0, (cudaStream_t)stream>>>(nbr, ld_out, n_out, n_out_dev, nbr_inv, ld_in, n_in);
    return check_launch("pcdb_rulebook_invert");
}

extern "C" int pcdb_fill_rows_i32(int32_t *dst, int ld, int n_maps, const int32_t *rows_dev, int rows_cap, int value,
                                  void *stream)
{
    using namespace pcdb;
    if (!dst || !rows_dev || ld < rows_cap || n_maps < 1 || n_maps > 65535 || rows_cap < 0 || (((uintptr_t)dst) & 15)) {
        set_last_error("pcdb_fill_rows_i32: invalid argument (ld=%d n_maps=%d rows_cap=%d; dst must be 16-byte aligned)", ld, n_maps, rows_cap);
        return kInvalidArgument;
    }
    if (rows_cap == 0) return kOk;
    fill_rows_i32_kernel<<<dim3((rows_cap + 1023) / 1024, n_maps), 256, 0, (cudaStream_t)stream>>>(dst, ld, rows_dev, rows_cap, value);
    return check_launch("pcdb_fill_rows_i32");
}

extern "C" int pcdb_abi_version(void) { return 1; }
extern "C" const char *pcdb_last_error(void) { return pcdb::g_last_error; }
