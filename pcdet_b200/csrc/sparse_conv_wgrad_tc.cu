// Weight gradient of the sparse convolution on the 5th-generation tensor cores (tcgen05 / TMEM), sm_100a.
//
// spconv v1.0 indiceConvBackward computes, per kernel offset, gather(features) ^T x gather(grad_out) with one cuBLAS GEMM
// per offset.  Here the whole layer is ONE output-stationary contraction whose reduction dimension is
// the OUTPUT ROW:
//
//     dW[k][ci][co] = sum over output rows o of  X[nbr[k][o]][ci] * dY[o][co]          (X[-1] = 0)
//
//   D (TMEM, fp32)   M = 128 = (2 * kGroup kernel offsets) x Cin      N = Cout      one accumulator per group of offsets
//   A (shared)       the SAME image the forward kernel gathers: 128 output rows x 128 bytes, kGroup offsets side by side
//                    in a row; two such tiles per stage.  Read as an MN-major operand (the 64 channels of a row are the
//                    M index, the row is the K index), SWIZZLE_128B, LBO = one tile, SBO = 8 rows.
//   B (shared)       128 rows of grad_out, NOT gathered and shared by every offset of the layer: MN-major as well
//                    (row = K index), swizzle by row width (128 / 64 / 32 bytes for Cout >= 64 / 32 / 16).
//   tcgen05.mma      M = 128, N = Cout, K = 16 output rows per instruction, 8 instructions per (tile, group).
//
// A CTA walks row tiles blockIdx.x, blockIdx.x + gridDim.x, ... and keeps accumulating into the same TMEM columns, so
// there is one epilogue per CTA: its partial (K, Cin, Cout) fp32 block goes to `partial[blockIdx.x]` and
// wgrad_reduce_kernel adds the blocks in index order -- no atomics, the same bits every run.  When the accumulators of
// all offsets do not fit the 512 TMEM columns (Cin = 64, Cout >= 64, K = 27) blockIdx.y splits the offsets.
//
// Rows without a neighbour must contribute zero, so unlike the forward kernel (which masks accumulator lanes) every
// (row, offset) piece is copied, zero-filled by cp.async when there is no neighbour.  Twelve producer warps in three
// groups fill alternate stages (a warp issues one LDGSTS every ~80 cycles, profiles/r02_mb_gather.md); two more warps
// stream grad_out; one elected lane issues the MMAs.
#include "tc_common.cuh"
#include "../../include/pcdet_b200.h"

namespace pcdb {
namespace wg {

using namespace tc;

constexpr int kRows = 128;                 // output rows per tile = 8 MMA K-steps
constexpr int kMaxK = 27;
constexpr int kProdGroups = 3;
constexpr int kGroupThreads = 128;
constexpr int kProdThreads = kProdGroups * kGroupThreads;
constexpr int kBThreads = 64;
constexpr int kThreads = kProdThreads + kBThreads + 32;      // 15 warps: 12 gather, 2 grad_out, 1 MMA / TMEM
constexpr int kMaxStages = 6;

template <int CIN, int COUT>
struct Cfg {
    static_assert(CIN == 16 || CIN == 32 || CIN == 64, "CIN must be 16, 32 or 64");
    static_assert(COUT == 16 || COUT == 32 || COUT == 64 || COUT == 128, "COUT must be 16, 32, 64 or 128");
    static constexpr int kGroup = 64 / CIN;                  // offsets per 128-byte A row
    static constexpr int kOffPerMma = 2 * kGroup;            // offsets per accumulator (M = 128)
    static constexpr int kCpo = CIN * 2 / 16;                // 16-byte pieces per (row, offset)
    static constexpr int kATile = kRows * 128;
    static constexpr int kAStage = 2 * kATile;
    static constexpr int kBRowBytes = COUT * 2 > 128 ? 128 : COUT * 2;
    static constexpr int kBHalves = COUT * 2 / kBRowBytes;   // 2 for Cout = 128
    static constexpr int kBHalfBytes = kRows * kBRowBytes;
    static constexpr int kBTile = kBHalves * kBHalfBytes;
    static constexpr int kBSwBits = kBRowBytes == 128 ? 3 : (kBRowBytes == 64 ? 2 : 1);
    static constexpr uint64_t kBLayout = kBRowBytes == 128 ? 2 : (kBRowBytes == 64 ? 4 : 6);   // SWIZZLE_128B / 64B / 32B
    static constexpr int kBChunksPerRow = COUT * 2 / 16;
    static constexpr int kMaxGroups = (kMaxK + kOffPerMma - 1) / kOffPerMma;                   // 14 / 7 / 4
    static constexpr int kGroupsPerCta = kMaxGroups < 512 / COUT ? kMaxGroups : 512 / COUT;
    static constexpr int cols(int g) { int c = 32; while (c < g * COUT) c <<= 1; return c; }
    // instruction descriptor: D = f32, A = B = bf16, BOTH MN-major (bits 15, 16), N >> 3 at bit 17, M >> 4 at bit 24
    static constexpr uint32_t kIdesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) |
                                       ((uint32_t)(COUT >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
};

// shared-memory matrix descriptor, MN-major: LBO = distance between 64-element blocks along M/N, SBO = distance
// between groups of 8 rows (the K index)
__device__ __forceinline__ uint64_t make_desc_mn(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint64_t layout)
{
    return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo_bytes >> 4) << 16) | ((uint64_t)(sbo_bytes >> 4) << 32) |
           (1ull << 46) | (layout << 61);
}

__device__ __forceinline__ void umma_bf16_acc(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}

// One 16-byte piece of input row `src` into shared memory, zeros when src < 0.
template <int ROW_BYTES>
__device__ __forceinline__ void gather_piece_zfill(uint32_t dst, const uint8_t *feat_piece, int src)
{
    const int s = src < 0 ? 0 : src;
    cp_async16(dst, feat_piece + (size_t)s * ROW_BYTES, src < 0 ? 0u : 16u);
}

template <int OFF_BYTES, int CPO, int P>
__device__ __forceinline__ void gather_passes_zfill(const uint32_t *dst, uint32_t tile_base, const uint8_t *feat_piece, int src_own)
{
    if constexpr (P < CPO) {
        gather_piece_zfill<OFF_BYTES>(dst[P] + tile_base, feat_piece, __shfl_sync(0xffffffffu, src_own, P, CPO));
        gather_passes_zfill<OFF_BYTES, CPO, P + 1>(dst, tile_base, feat_piece, src_own);
    }
}

// features (n_in, CIN) bf16; gout (n_out, COUT) bf16; nbr (K, ld); partial (gridDim.x, K, CIN, COUT) fp32.
template <int CIN, int COUT>
__global__ void __launch_bounds__(kThreads, 1)
conv_wgrad_tc(const __nv_bfloat16 *__restrict__ feat, const __nv_bfloat16 *__restrict__ gout, const int *__restrict__ nbr,
              int ld, int K, int n_out, const int *__restrict__ n_out_dev, float *__restrict__ partial, int n_stages)
{
    using C = Cfg<CIN, COUT>;
    extern __shared__ uint8_t smem_raw[];
    if (n_out_dev) { const int m = __ldg(n_out_dev); n_out = m < n_out ? m : n_out; }
    const int tiles = (n_out + kRows - 1) / kRows;
    // this CTA's accumulator groups [g0, g0 + ng) and kernel offsets [k0, ...)
    const int groups_total = (K + C::kOffPerMma - 1) / C::kOffPerMma;
    const int gpc = (groups_total + (int)gridDim.y - 1) / (int)gridDim.y;      // <= kGroupsPerCta (see launch)
    const int g0 = blockIdx.y * gpc;
    const int ng = max(0, min(gpc, groups_total - g0));
    const int k0 = g0 * C::kOffPerMma;
    const int my_tiles = (int)blockIdx.x < tiles ? (tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;

    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t *aligned = smem_raw + (base - smem_u32(smem_raw));
    const uint32_t a_base = base, b_base = base + n_stages * C::kAStage;
    uint64_t *bars = reinterpret_cast<uint64_t *>(aligned + n_stages * C::kAStage + 2 * C::kBTile);
    // full_a[6] empty_a[6] full_b[2] empty_b[2] acc
    const uint32_t bar_full_a = smem_u32(bars), bar_empty_a = smem_u32(bars + kMaxStages);
    const uint32_t bar_full_b = smem_u32(bars + 2 * kMaxStages), bar_empty_b = smem_u32(bars + 2 * kMaxStages + 2);
    const uint32_t bar_acc = smem_u32(bars + 2 * kMaxStages + 4);
    uint32_t *s_tmem = reinterpret_cast<uint32_t *>(bars + 2 * kMaxStages + 5);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int tmem_cols = C::cols(C::kGroupsPerCta);
    if (tid == 0) {
        for (int s = 0; s < n_stages; ++s) {
            mbar_init(bar_full_a + 8 * s, kGroupThreads);      // the async arrivals of the producer group that fills it
            mbar_init(bar_empty_a + 8 * s, 1);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(bar_full_b + 8 * s, kBThreads);
            mbar_init(bar_empty_b + 8 * s, 1);
        }
        mbar_init(bar_acc, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == kThreads / 32 - 1) tmem_alloc(smem_u32(s_tmem), tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = uniform(*s_tmem);

    if (warp < kProdThreads / 32) {
        // ===== gather producers: group pg fills the stages of accumulator groups gi with gi % 3 == pg =============
        const int pg = warp >> 2, r = tid & (kGroupThreads - 1), wq = warp & 3;
        constexpr int kCpo = C::kCpo;
        const int piece = lane % kCpo, jw = lane / kCpo;
        const uint8_t *feat_b = reinterpret_cast<const uint8_t *>(feat) + piece * 16;
        uint32_t dst[C::kGroup * kCpo];         // offset of this lane's piece inside an A tile, per (sub-offset, pass)
#pragma unroll
        for (int sub = 0; sub < C::kGroup; ++sub)
#pragma unroll
            for (int p = 0; p < kCpo; ++p)
                dst[sub * kCpo + p] = swizzled_offset<128, 3>(32 * wq + jw * kCpo + p, sub * kCpo + piece);
        int it = 0;
        for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x, ++it) {
            const int row = tile * kRows + r;
#pragma unroll
            for (int gi = 0; gi < C::kGroupsPerCta; ++gi) {
                if (gi >= ng || gi % kProdGroups != pg) continue;            // uniform over the warp
                // this row's neighbours at the group's offsets: independent loads, issued before the wait
                int src[C::kOffPerMma];
#pragma unroll
                for (int j = 0; j < C::kOffPerMma; ++j) {
                    const int k = k0 + gi * C::kOffPerMma + j;
                    src[j] = (k < K && row < n_out) ? __ldg(nbr + (size_t)k * ld + row) : -1;
                }
                const int i = it * ng + gi, s = i % n_stages, wrap = i / n_stages;
                if (wrap > 0) mbar_wait(bar_empty_a + 8 * s, (uint32_t)(wrap - 1) & 1u);
                const uint32_t st = a_base + s * C::kAStage;
#pragma unroll
                for (int t = 0; t < 2; ++t)
#pragma unroll
                    for (int sub = 0; sub < C::kGroup; ++sub)
                        if (k0 + gi * C::kOffPerMma + t * C::kGroup + sub < K)      // offsets past K: lanes nobody reads
                            gather_passes_zfill<CIN * 2, kCpo, 0>(dst + sub * kCpo, st + t * C::kATile, feat_b, src[t * C::kGroup + sub]);
                cp_async_arrive(bar_full_a + 8 * s);
            }
        }
        if (warp < 4) {
            // ===== epilogue (producer group 0): TMEM lane m = accumulator row (offset, ci) -> partial[blockIdx.x] ====
            if (my_tiles > 0 && ng > 0) {
                mbar_wait(bar_acc, 0);
                tc_fence_after();
            }
            const int m = 32 * warp + lane, t = m >> 6, sub = (m & 63) / CIN, ci = (m & 63) % CIN;
            float *mine = partial + (size_t)blockIdx.x * K * CIN * COUT;
#pragma unroll 1
            for (int gi = 0; gi < ng; ++gi) {
                const int k = k0 + gi * C::kOffPerMma + t * C::kGroup + sub;
#pragma unroll 1
                for (int c0 = 0; c0 < COUT; c0 += 16) {
                    uint32_t v[16];
                    tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)(gi * COUT + c0), v);
                    tmem_ld_wait();
                    if (k < K) {
                        float4 *o = reinterpret_cast<float4 *>(mine + ((size_t)k * CIN + ci) * COUT + c0);
#pragma unroll
                        for (int q = 0; q < 4; ++q)
                            o[q] = my_tiles > 0 ? make_float4(__uint_as_float(v[4 * q]), __uint_as_float(v[4 * q + 1]),
                                                              __uint_as_float(v[4 * q + 2]), __uint_as_float(v[4 * q + 3]))
                                                : make_float4(0.f, 0.f, 0.f, 0.f);
                    }
                }
            }
        }
    } else if (warp < (kProdThreads + kBThreads) / 32) {
        // ===== grad_out rows of the tile, shared by all offsets: double-buffered =====================================
        const int tb = tid - kProdThreads;
        const uint8_t *g8 = reinterpret_cast<const uint8_t *>(gout);
        int it = 0;
        for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x, ++it) {
            const int buf = it & 1;
            if (it >= 2) mbar_wait(bar_empty_b + 8 * buf, (uint32_t)((it >> 1) - 1) & 1u);
            const uint32_t bb = b_base + buf * C::kBTile;
#pragma unroll 4
            for (int c = tb; c < kRows * C::kBChunksPerRow; c += kBThreads) {
                const int rr = c / C::kBChunksPerRow, ch = c % C::kBChunksPerRow;
                const int half = ch / (C::kBRowBytes / 16), cc = ch % (C::kBRowBytes / 16);
                const int row = tile * kRows + rr;
                const bool ok = row < n_out;
                cp_async16(bb + half * C::kBHalfBytes + swizzled_offset<C::kBRowBytes, C::kBSwBits>(rr, cc),
                           g8 + ((size_t)(ok ? row : 0) * COUT * 2 + ch * 16), ok ? 16u : 0u);
            }
            cp_async_arrive(bar_full_b + 8 * buf);
        }
    } else {
        // ===== MMA issuer ================================================================================================
        // (the broadcast tells ptxas the trip count is warp-uniform, so the descriptors stay in uniform registers)
        const int tiles_u = (int)uniform((uint32_t)tiles);
        int it = 0;
        for (int tile = blockIdx.x; tile < tiles_u; tile += gridDim.x, ++it) {
            const int buf = it & 1;
            for (int gi = 0; gi < ng; ++gi) {
                const int i = it * ng + gi, s = i % n_stages, wrap = i / n_stages;
                mbar_wait(bar_full_a + 8 * s, (uint32_t)wrap & 1u);
                if (gi == 0) mbar_wait(bar_full_b + 8 * buf, (uint32_t)(it >> 1) & 1u);
                tc_fence_after();
                const uint32_t a0 = a_base + s * C::kAStage, b0 = b_base + buf * C::kBTile;
                if (elect_one()) {
#pragma unroll
                    for (int j = 0; j < kRows / 16; ++j) {
                        const uint64_t da = make_desc_mn(a0 + j * 16 * 128, C::kATile, 8 * 128, 2);
                        const uint64_t db = make_desc_mn(b0 + j * 16 * C::kBRowBytes, C::kBHalfBytes, 8 * C::kBRowBytes, C::kBLayout);
                        umma_bf16_acc(tmem + (uint32_t)(gi * COUT), da, db, C::kIdesc, (it > 0 || j > 0) ? 1u : 0u);
                    }
                }
                __syncwarp();
                if (elect_one()) {
                    umma_commit(bar_empty_a + 8 * s);
                    if (gi == ng - 1) umma_commit(bar_empty_b + 8 * buf);
                }
                __syncwarp();
            }
        }
        if (my_tiles > 0 && ng > 0 && elect_one()) umma_commit(bar_acc);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == kThreads / 32 - 1) {
        __syncwarp();
        tmem_dealloc(tmem, tmem_cols);
    }
}

// grad_weight[e] (+)= sum over c of partial[c][e], in index order (deterministic).  Eight independent loads in flight per
// thread: the pass is a latency-bound walk over up to 148 partial blocks otherwise.
__global__ void __launch_bounds__(128)
wgrad_reduce_kernel(const float *__restrict__ partial, int n_partials, int count4, int accumulate, float *__restrict__ grad_w)
{
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= count4) return;
    const float4 *p = reinterpret_cast<const float4 *>(partial) + e;
    float4 acc = accumulate ? reinterpret_cast<float4 *>(grad_w)[e] : make_float4(0.f, 0.f, 0.f, 0.f);
    int c = 0;
    for (; c + 8 <= n_partials; c += 8) {
        float4 v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = __ldg(p + (size_t)(c + j) * count4);
#pragma unroll
        for (int j = 0; j < 8; ++j) { acc.x += v[j].x; acc.y += v[j].y; acc.z += v[j].z; acc.w += v[j].w; }
    }
    for (; c < n_partials; ++c) {
        const float4 v = __ldg(p + (size_t)c * count4);
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    reinterpret_cast<float4 *>(grad_w)[e] = acc;
}

template <int CIN, int COUT>
int grid_x(int K, int n_out)
{
    using C = Cfg<CIN, COUT>;
    const int groups_total = (K + C::kOffPerMma - 1) / C::kOffPerMma;
    const int ysplit = (groups_total + C::kGroupsPerCta - 1) / C::kGroupsPerCta;
    const int tiles = (n_out + kRows - 1) / kRows;
    int gx = kNumSMs / ysplit;
    gx = gx < 1 ? 1 : gx;
    return tiles < gx ? (tiles < 1 ? 1 : tiles) : gx;
}

template <int CIN, int COUT>
int launch(const void *features, const void *grad_out, const int32_t *nbr, int ld, int K, int n_out, const int32_t *n_out_dev,
           float *grad_weight, int accumulate, void *workspace, size_t workspace_bytes, cudaStream_t stream)
{
    using C = Cfg<CIN, COUT>;
    const int groups_total = (K + C::kOffPerMma - 1) / C::kOffPerMma;
    const int ysplit = (groups_total + C::kGroupsPerCta - 1) / C::kGroupsPerCta;
    const int gx = grid_x<CIN, COUT>(K, n_out);
    const size_t need = (size_t)gx * K * CIN * COUT * sizeof(float);
    if (workspace_bytes < need) {
        set_last_error("pcdb_sparse_conv_wgrad: workspace too small (%zu < %zu bytes)", workspace_bytes, need);
        return kWorkspaceTooSmall;
    }
    const int fixed = 1024 + 2 * C::kBTile + (2 * kMaxStages + 6) * 8 + 64;
    int n_stages = (225 * 1024 - fixed) / C::kAStage;
    n_stages = n_stages > kMaxStages ? kMaxStages : n_stages;
    // The three producer groups run ahead of each other: a group's wait for the w-th release of a stage is only safe (no
    // parity aliasing on the mbarrier) if its previous wait already implied the (w-1)-th, i.e. if the ring is at least as deep
    // as the largest step between two consecutive (tile, group) items of one producer group, which is 4 (7 or 4 groups per
    // CTA, three producer groups).  Every shape gets 4-6 stages; 2 or 3 deadlock or race (measured).
    if (n_stages < 4) { set_last_error("pcdb_sparse_conv_wgrad: internal: ring of %d stages", n_stages); return kUnsupported; }
    const int smem = fixed + n_stages * C::kAStage;
    cudaError_t err = cudaFuncSetAttribute(conv_wgrad_tc<CIN, COUT>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (err != cudaSuccess) {
        set_last_error("pcdb_sparse_conv_wgrad: cudaFuncSetAttribute(%d bytes) failed: %s", smem, cudaGetErrorString(err));
        return kCudaError;
    }
    conv_wgrad_tc<CIN, COUT><<<dim3(gx, ysplit), kThreads, smem, stream>>>(
        (const __nv_bfloat16 *)features, (const __nv_bfloat16 *)grad_out, nbr, ld, K, n_out, n_out_dev, (float *)workspace, n_stages);
    int rc = check_launch("pcdb_sparse_conv_wgrad(tcgen05)");
    if (rc != kOk) return rc;
    const int count4 = K * CIN * COUT / 4;
    wgrad_reduce_kernel<<<(count4 + 127) / 128, 128, 0, stream>>>((const float *)workspace, gx, count4, accumulate, grad_weight);
    return check_launch("pcdb_sparse_conv_wgrad(reduce)");
}

}  // namespace wg

#define PCDB_WG_SHAPES(X) \
    X(16, 16) X(16, 32) X(16, 64) X(16, 128) X(32, 16) X(32, 32) X(32, 64) X(32, 128) X(64, 16) X(64, 32) X(64, 64) X(64, 128)

}  // namespace pcdb

using namespace pcdb;

extern "C" size_t pcdb_sparse_conv_wgrad_workspace_bytes(int kernel_volume, int n_out, int c_in, int c_out)
{
    if (kernel_volume < 1 || kernel_volume > wg::kMaxK || n_out < 0) return 0;
#define PCDB_WG_CASE(CI, CO) \
    if (c_in == CI && c_out == CO) return (size_t)wg::grid_x<CI, CO>(kernel_volume, n_out) * kernel_volume * CI * CO * sizeof(float);
    PCDB_WG_SHAPES(PCDB_WG_CASE)
#undef PCDB_WG_CASE
    return 0;
}

extern "C" int pcdb_sparse_conv_wgrad(const void *features, int n_in, const void *grad_out, const int32_t *nbr, int ld,
                                      int kernel_volume, int n_out, const int32_t *n_out_dev, int c_in, int c_out,
                                      float *grad_weight, int accumulate, void *workspace, size_t workspace_bytes, void *stream)
{
    if (!features || !grad_out || !nbr || !grad_weight || !workspace || n_in < 0 || n_out < 0 || ld < n_out ||
        kernel_volume < 1 || kernel_volume > wg::kMaxK) {
        set_last_error("pcdb_sparse_conv_wgrad: invalid argument (n_in=%d n_out=%d K=%d ld=%d)", n_in, n_out, kernel_volume, ld);
        return kInvalidArgument;
    }
#define PCDB_WG_CASE(CI, CO) \
    if (c_in == CI && c_out == CO) \
        return wg::launch<CI, CO>(features, grad_out, nbr, ld, kernel_volume, n_out, n_out_dev, grad_weight, accumulate, workspace, \
                                  workspace_bytes, (cudaStream_t)stream);
    PCDB_WG_SHAPES(PCDB_WG_CASE)
#undef PCDB_WG_CASE
    set_last_error("pcdb_sparse_conv_wgrad: the tcgen05 kernel takes c_in in {16,32,64}, c_out in {16,32,64,128}; got %d -> %d", c_in, c_out);
    return kUnsupported;
}
