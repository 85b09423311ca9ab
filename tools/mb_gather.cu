// Microbenchmark: how fast can one SM (and the whole chip) pull scattered 128-byte feature rows out of L2 into
// the 128B-swizzled K-major shared-memory image tcgen05.mma reads?  Decides the gather engine of sparse_conv_tc.cu.
//
//   variant 0  LDGSTS   128 threads, 8 lanes per row (16 B each), completion by cp.async.mbarrier.arrive.noinc
//   variant 1  LDGSTS   256 threads
//   variant 2  TMA      cp.async.bulk.tensor.2d ... tile::gather4, 32 lanes of ONE warp issue 4 rows each
//   variant 3  TMA      gather4 issued by 4 warps (8 lanes each ... every lane still 4 rows)
//   variant 4  bulk     cp.async.bulk 16 KB contiguous per stage (upper bound of L2 -> SM streaming)
//   variant 5  TMA      gather4, 42 % of the rows replaced by one fixed (L2-hot) dummy row
//   variant 6  LDGSTS   128 threads, only 58 % of the rows fetched, compacted (4 rows per warp instruction)
//
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o tools/_bin/mb_gather tools/mb_gather.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cstdint>
#include <vector>

constexpr int kRows = 128, kRowBytes = 128, kStageBytes = kRows * kRowBytes, kStages = 6;

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count)); }
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    for (uint32_t spin = 0; spin < (1u << 22); ++spin) if (mbar_try_wait(bar, parity)) return;
    __trap();
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void *src) { asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory"); }
__device__ __forceinline__ void cp_async_arrive(uint32_t bar) { asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory"); }
__device__ __forceinline__ void tma_gather4(uint32_t dst, const CUtensorMap *tmap, uint32_t bar, int col, int r0, int r1, int r2, int r3)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
                 ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(bar), "r"(col), "r"(r0), "r"(r1), "r"(r2), "r"(r3) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ uint32_t swz(uint32_t r, uint32_t c) { const uint32_t o = r * 128 + c * 16; return o ^ (((o >> 7) & 7u) << 4); }

// idx: [iters][128] row indices (shared by all CTAs, offset by blockIdx so CTAs touch different rows)
template <int VARIANT>
__global__ void __launch_bounds__(256) gather_kernel(const __grid_constant__ CUtensorMap tmap, const uint8_t *__restrict__ feat, int n_rows,
                                                     const int *__restrict__ idx, int iters, long long *cycles, unsigned *sink)
{
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    __shared__ uint64_t bars[kStages];
    __shared__ uint8_t s_rank[8][32];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    constexpr int kThreads = (VARIANT == 1) ? 256 : 128;
    constexpr bool kLdgsts = VARIANT == 0 || VARIANT == 1 || VARIANT == 6;
    if (tid == 0) {
        for (int s = 0; s < kStages; ++s) mbar_init(smem_u32(&bars[s]), kLdgsts ? kThreads : 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int shift = (blockIdx.x * 977) % n_rows;
    const long long t0 = clock64();
    if (kLdgsts) {
        if (tid < kThreads) {
            // rows per warp per stage: 128 / (kThreads / 32); lanes: 8 per row -> 4 rows per instruction
            constexpr int kWarps = kThreads / 32, kRowsPerWarp = kRows / kWarps, kPasses = kRowsPerWarp / 4;
            const int piece = lane & 7, jw = lane >> 3;
            int nxt = __ldg(idx + (size_t)0 * kRows + (tid % kRows));
            for (int it = 0; it < iters; ++it) {
                const int s = it % kStages;
                const uint32_t bar = smem_u32(&bars[s]);
                if (it >= kStages) mbar_wait(bar, ((it / kStages) - 1) & 1);
                int mine = nxt + shift; mine = mine >= n_rows ? mine - n_rows : mine;
                if (it + 1 < iters) nxt = __ldg(idx + (size_t)(it + 1) * kRows + (tid % kRows));
                const uint32_t sbase = base + s * kStageBytes;
                if (VARIANT == 6) {
                    // compaction: only rows with (hash & 127) < 74 are "present"; lane group j fetches the
                    // (4p + j)-th present row of this warp's 32 rows
                    const bool have = ((unsigned)(mine * 2654435761u) >> 25) < 74u;
                    const unsigned m = __ballot_sync(0xffffffffu, have);
                    const int cnt = __popc(m);
                    if (have) s_rank[warp][__popc(m & ((1u << lane) - 1u))] = (uint8_t)lane;
                    __syncwarp();
                    for (int p = 0; p * 4 < cnt; ++p) {
                        const int want = p * 4 + jw;
                        const int srcl = s_rank[warp][want & 31];
                        const int src = __shfl_sync(0xffffffffu, mine, srcl);
                        if (want < cnt) cp_async16(sbase + swz(warp * 32 + srcl, piece), feat + (size_t)src * kRowBytes + piece * 16);
                    }
                    __syncwarp();
                } else {
#pragma unroll
                    for (int p = 0; p < kPasses; ++p) {
                        const int half = (VARIANT == 1) ? (warp >> 2) : 0;
                        const int r = half * 16 + jw * kPasses + p;      // lane (of this warp) that holds the row index
                        const int src = __shfl_sync(0xffffffffu, mine, r);
                        const int row = (warp & 3) * 32 + r;
                        cp_async16(sbase + swz(row, piece), feat + (size_t)src * kRowBytes + piece * 16);
                    }
                }
                cp_async_arrive(bar);
            }
            // drain
            for (int s = 0; s < kStages && s < iters; ++s) {
                const int last = ((iters - 1 - s) / kStages) * kStages + s;        // last iteration that used stage s
                mbar_wait(smem_u32(&bars[s]), (last / kStages) & 1);
            }
        }
    } else if (VARIANT == 4) {
        if (tid == 0) {
            for (int it = 0; it < iters; ++it) {
                const int s = it % kStages;
                const uint32_t bar = smem_u32(&bars[s]);
                if (it >= kStages) mbar_wait(bar, ((it / kStages) - 1) & 1);
                mbar_expect_tx(bar, kStageBytes);
                const size_t chunk = ((size_t)(it * 131 + blockIdx.x * 17) % (size_t)(n_rows / kRows)) * kStageBytes;
                bulk_g2s(base + s * kStageBytes, feat + chunk, kStageBytes, bar);
            }
            for (int s = 0; s < kStages && s < iters; ++s) {
                const int last = ((iters - 1 - s) / kStages) * kStages + s;
                mbar_wait(smem_u32(&bars[s]), (last / kStages) & 1);
            }
        }
    } else {
        // TMA gather4: variant 2/5 one warp, variant 3 four warps (warp w: rows 32w .. 32w+31, lanes 0-7 issue)
        const bool issuer = (VARIANT == 3) ? (warp < 4 && lane < 8) : (warp == 0);
        const int row4 = (VARIANT == 3) ? (warp * 8 + lane) : lane;       // which group of 4 rows
        if (issuer) {
            for (int it = 0; it < iters; ++it) {
                const int s = it % kStages;
                const uint32_t bar = smem_u32(&bars[s]);
                if (it >= kStages) mbar_wait(bar, ((it / kStages) - 1) & 1);
                int4 r = *reinterpret_cast<const int4 *>(idx + (size_t)it * kRows + row4 * 4);
                int v[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    int m = v[j] + shift; m = m >= n_rows ? m - n_rows : m;
                    if (VARIANT == 5 && (((unsigned)(m * 2654435761u) >> 25) >= 74u)) m = 0;
                    v[j] = m;
                }
                if (row4 == 0) mbar_expect_tx(bar, kStageBytes);
                __syncwarp(VARIANT == 3 ? 0xffu : 0xffffffffu);
                tma_gather4(base + s * kStageBytes + row4 * 4 * kRowBytes, &tmap, bar, 0, v[0], v[1], v[2], v[3]);
            }
            for (int s = 0; s < kStages && s < iters; ++s) {
                const int last = ((iters - 1 - s) / kStages) * kStages + s;
                mbar_wait(smem_u32(&bars[s]), (last / kStages) & 1);
            }
        }
    }
    __syncthreads();
    const long long t1 = clock64();
    if (tid == 0) {
        cycles[blockIdx.x] = t1 - t0;
        sink[blockIdx.x] = *reinterpret_cast<volatile unsigned *>(smem_raw + 1024 + (blockIdx.x & 255) * 4);
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

template <int V>
void run(const char *name, const CUtensorMap &tmap, const uint8_t *feat, int n_rows, const int *idx, int iters, int grid, double frac_rows)
{
    long long *cyc; unsigned *sink;
    CK(cudaMalloc(&cyc, grid * sizeof(long long))); CK(cudaMalloc(&sink, grid * sizeof(unsigned)));
    const int smem = kStages * kStageBytes + 2048;
    CK(cudaFuncSetAttribute(gather_kernel<V>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
        CK(cudaEventRecord(e0));
        gather_kernel<V><<<grid, 256, smem>>>(tmap, feat, n_rows, idx, iters, cyc, sink);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (rep > 0 && ms < best) best = ms;
    }
    std::vector<long long> h(grid);
    CK(cudaMemcpy(h.data(), cyc, grid * sizeof(long long), cudaMemcpyDeviceToHost));
    long long mx = 0; double avg = 0;
    for (auto c : h) { mx = c > mx ? c : mx; avg += (double)c / grid; }
    const double bytes_cta = (double)iters * kStageBytes * frac_rows;
    printf("%-34s grid %4d  %8.1f us  per-SM %6.1f B/cyc (avg cyc/stage %7.1f, max %7.1f)  chip %7.1f GB/s\n", name, grid, best * 1e3,
           bytes_cta / avg, avg / iters, (double)mx / iters, bytes_cta * grid / (best * 1e-3) / 1e9);
    cudaFree(cyc); cudaFree(sink);
}

int main(int argc, char **argv)
{
    setvbuf(stdout, nullptr, _IONBF, 0);
    const int n_rows = argc > 1 ? atoi(argv[1]) : 45312;          // 5.8 MB of 128-byte rows (L2 resident)
    const int iters = argc > 2 ? atoi(argv[2]) : 2000;
    uint8_t *feat; int *idx;
    CK(cudaMalloc(&feat, (size_t)n_rows * kRowBytes));
    CK(cudaMemset(feat, 1, (size_t)n_rows * kRowBytes));
    std::vector<int> h((size_t)iters * kRows);
    srand(1);
    for (auto &v : h) v = (int)(((unsigned)rand() * 32768u + (unsigned)rand()) % (unsigned)n_rows);
    CK(cudaMalloc(&idx, h.size() * 4));
    CK(cudaMemcpy(idx, h.data(), h.size() * 4, cudaMemcpyHostToDevice));
    CUtensorMap tmap; memset(&tmap, 0, sizeof(tmap));
    void *p = nullptr; cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
    const cuuint64_t gdim[2] = {64, (cuuint64_t)n_rows}; const cuuint64_t gstride[1] = {128};
    const cuuint32_t box[2] = {64, 1}; const cuuint32_t estr[2] = {1, 1};
    CUresult r = ((EncodeTiledFn)p)(&tmap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, feat, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 1; }
    printf("rows %d (%.1f MB), %d stages of 128 rows x 128 B per CTA\n", n_rows, n_rows * 128.0 / 1e6, iters);
    for (int grid : {1, 148, 296}) {
        run<0>("LDGSTS 128 thr", tmap, feat, n_rows, idx, iters, grid, 1.0);
        run<1>("LDGSTS 256 thr", tmap, feat, n_rows, idx, iters, grid, 1.0);
        run<6>("LDGSTS 128 thr compact 58%", tmap, feat, n_rows, idx, iters, grid, 74.0 / 128.0);
        run<2>("TMA gather4 1 warp", tmap, feat, n_rows, idx, iters, grid, 1.0);
        run<3>("TMA gather4 4 warps x 8 lanes", tmap, feat, n_rows, idx, iters, grid, 1.0);
        run<5>("TMA gather4 42% dummy row", tmap, feat, n_rows, idx, iters, grid, 1.0);
        run<4>("bulk 16 KB contiguous", tmap, feat, n_rows, idx, iters, grid, 1.0);
    }
    return 0;
}
