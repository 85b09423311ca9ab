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

}  // namespace pcdb

extern "C" int pcdb_abi_version(void) { return 1; }
extern "C" const char *pcdb_last_error(void) { return pcdb::g_last_error; }
