// Microbenchmark 2: issue-side limits of the row gather.  One CTA per SM, W warps; every warp owns a private
// 8 KB shared-memory window and loops: 8 independent 4-row (512 B) copies of random 128-byte rows, then waits for them
// (cp.async.wait_group / register loads), so nothing but the LSU path is measured.
//   mode 0: cp.async.cg 16 B (LDGSTS), indices in registers (no shuffles, no dependent chain)
//   mode 1: ld.global.nc.v4 -> st.shared.v4 (software pipelined: next batch's loads in flight while storing)
//   mode 2: LDGSTS, 2 rows per instruction only (16 lanes active) -- does the cost scale with lines or with instructions?
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/_bin/mb_gather2 tools/mb_gather2.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cstdint>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t swz(uint32_t r, uint32_t c) { const uint32_t o = r * 128 + c * 16; return o ^ (((o >> 7) & 7u) << 4); }

template <int MODE>
__global__ void __launch_bounds__(1024) k(const uint8_t *__restrict__ feat, int n_rows, const int *__restrict__ idx, int iters, long long *cyc, unsigned *sink)
{
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, piece = lane & 7, jw = lane >> 3;
    const uint32_t win = base + warp * 4096;            // 32 rows x 128 B per warp
    const int *my = idx + (size_t)(blockIdx.x * 32 + warp) * 64;
    int rows[8];
#pragma unroll
    for (int p = 0; p < 8; ++p) rows[p] = my[(p * 4 + jw) & 63];
    __syncthreads();
    const long long t0 = clock64();
    if (MODE == 0 || MODE == 2) {
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int p = 0; p < 8; ++p) {
                int r = rows[p] + it * 131; r = r % n_rows;
                const uint32_t dst = win + swz(p * 4 + jw, piece);
                if (MODE == 0 || lane < 16)
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(feat + (size_t)r * 128 + piece * 16) : "memory");
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
            asm volatile("cp.async.wait_group 2;" ::: "memory");
        }
        asm volatile("cp.async.wait_group 0;" ::: "memory");
    } else {
        uint4 buf[8];
#pragma unroll
        for (int p = 0; p < 8; ++p) { int r = rows[p] % n_rows; buf[p] = __ldg(reinterpret_cast<const uint4 *>(feat + (size_t)r * 128 + piece * 16)); }
        for (int it = 1; it <= iters; ++it) {
            uint4 nxt[8];
#pragma unroll
            for (int p = 0; p < 8; ++p) { int r = (rows[p] + it * 131) % n_rows; nxt[p] = __ldg(reinterpret_cast<const uint4 *>(feat + (size_t)r * 128 + piece * 16)); }
#pragma unroll
            for (int p = 0; p < 8; ++p)
                asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(win + swz(p * 4 + jw, piece)), "r"(buf[p].x), "r"(buf[p].y), "r"(buf[p].z), "r"(buf[p].w) : "memory");
#pragma unroll
            for (int p = 0; p < 8; ++p) buf[p] = nxt[p];
        }
    }
    const long long t1 = clock64();
    __syncthreads();
    if (threadIdx.x == 0) { cyc[blockIdx.x] = t1 - t0; sink[blockIdx.x] = *reinterpret_cast<volatile unsigned *>(smem_raw + 2048); }
}

template <int MODE>
void run(const char *name, const uint8_t *feat, int n_rows, const int *idx, int iters, int warps)
{
    long long *cyc; unsigned *sink;
    CK(cudaMalloc(&cyc, 148 * 8)); CK(cudaMalloc(&sink, 148 * 4));
    const int smem = warps * 4096 + 2048;
    CK(cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 3; ++rep) {
        CK(cudaEventRecord(e0));
        k<MODE><<<148, warps * 32, smem>>>(feat, n_rows, idx, iters, cyc, sink);
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (rep > 0 && ms < best) best = ms;
    }
    std::vector<long long> h(148);
    CK(cudaMemcpy(h.data(), cyc, 148 * 8, cudaMemcpyDeviceToHost));
    double avg = 0; for (auto c : h) avg += (double)c / 148;
    const double instr_sm = (double)iters * 8 * warps;
    printf("%-22s warps %2d  cycles/instr per warp %6.1f  per SM %5.1f  -> %5.1f B/cyc/SM, chip %6.0f GB/s\n", name, warps, avg / (iters * 8.0),
           avg / instr_sm, instr_sm * (MODE == 2 ? 256 : 512) / avg, instr_sm * (MODE == 2 ? 256 : 512) * 148 / (best * 1e-3) / 1e9);
    cudaFree(cyc); cudaFree(sink);
}

int main()
{
    setvbuf(stdout, nullptr, _IONBF, 0);
    const int n_rows = 45312, iters = 400;
    uint8_t *feat; int *idx;
    CK(cudaMalloc(&feat, (size_t)n_rows * 128)); CK(cudaMemset(feat, 1, (size_t)n_rows * 128));
    std::vector<int> h(148 * 32 * 64);
    srand(1);
    for (auto &v : h) v = (int)(((unsigned)rand() * 32768u + (unsigned)rand()) % (unsigned)n_rows);
    CK(cudaMalloc(&idx, h.size() * 4)); CK(cudaMemcpy(idx, h.data(), h.size() * 4, cudaMemcpyHostToDevice));
    for (int w : {1, 2, 4, 8, 12, 16, 24, 32}) run<0>("LDGSTS 4 rows/instr", feat, n_rows, idx, iters, w);
    for (int w : {4, 8, 16, 32}) run<2>("LDGSTS 2 rows/instr", feat, n_rows, idx, iters, w);
    for (int w : {1, 2, 4, 8, 12, 16, 24}) run<1>("LDG.128 + STS.128", feat, n_rows, idx, iters, w);
    return 0;
}
