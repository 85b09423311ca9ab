// Sparse convolution forward on the 5th-generation tensor cores (tcgen05 / TMEM), sm_100a.
//
// The only dense contraction of the hot path (spconv v1.0 indiceConv: per offset gather -> cuBLAS
// GEMM -> scatter-add) as ONE output-stationary kernel per layer:
//
//   CTA = 128 output rows.  A pipeline stage carries 64 input channels of the tile: ONE kernel offset for
//   Cin = 64, TWO for Cin = 32, FOUR for Cin = 16, side by side along K (see Cfg).  For every stage that has at
//   least one neighbour in the tile:
//     the gather engine brings the contributing input rows nbr[k][row] and the stage's (Cout x 64) weight tile
//       into a 2-8 stage shared-memory ring, in the 128B-swizzled K-major image the tensor core reads:
//         * cp.async (default, algo 3): 128 threads, 16-byte LDGSTS, 8/4/2 consecutive lanes fetch one row;
//           rows WITHOUT a neighbour are not fetched at all -- the MMA of that offset carries a
//           disable-output-lane mask, so their shared-memory rows are never read into the accumulator;
//           completion through cp.async.mbarrier.arrive; the weight tile arrives by one cp.async.bulk (UBLKCP);
//         * TMA (algo 2, Cin = 64): one warp, each lane issues ONE cp.async.bulk.tensor ...tile::gather4 for
//           4 rows (missing neighbours are out-of-bounds row indices, which TMA zero-fills).  Measured 2x slower
//           than LDGSTS (155 vs 79 us, 1184 tiles of 64->64).  Round 2 measured the engine alone with every row in
//           bounds (tools/mb_gather.cu, profiles/r02_mb_gather.md): 32-72 cycles per gather4 = 7-16 B/cycle/SM against
//           27-59 for LDGSTS, and slower again when missing rows point at one dummy row -- so it stays the fallback;
//     one elected lane of a converged warp issues tcgen05.mma (M=128, N=Cout, K=16 per instruction)
//       accumulating ALL offsets into the same fp32 accumulator in TMEM (zeroed once), and tcgen05.commit
//       releases the stage;
//   epilogue: tcgen05.ld the accumulator, apply the folded BatchNorm scale/shift (+bias), ReLU,
//     convert to bf16, stage through shared memory and store the tile with one coalesced copy.
//   Ring depth / CTAs per SM are chosen per launch from the expected tile count; consecutive layers overlap
//   set-up and tail through programmatic dependent launch (griddepcontrol).  DESIGN.md §5 lists what was
//   measured on the way (and what was tried and dropped).
//
// No scatter, no atomics, every output row written exactly once, fixed summation order.
// Algorithmic traffic per layer: N_in*Cin*2 + N_out*Cout*2 + K*Cin*Cout*2 + 4*K*N_out bytes.
#include "tc_common.cuh"   // CUtensorMap types only; the encoder is fetched through cudaGetDriverEntryPoint
#include "../../include/pcdet_b200.h"
#include <cstdlib>

namespace pcdb {

namespace tc {

#ifdef PCDB_TC_TRACE
__device__ long long g_trace[8][32];
#define TRACE(slot, it) do { if (blockIdx.x == PCDB_TC_TRACE) g_trace[slot][it] = clock64(); } while (0)
#else
#define TRACE(slot, it) do { } while (0)
#endif

constexpr int kTileM = 128;
constexpr int kMaxK = 27;          // kernel offsets (3x3x3)
// Every CTA of a layer streams the same K weight tiles.  Round 1 replicated the packed weights 16 times (CTAs read
// replica blockIdx.x % kWReplicas) against an L2 hot spot measured with the first kernel; with the present ring the
// step is the same with 1, 2, 4 or 16 replicas (17 197 / 17 099 / 17 209 / 17 104 frames/s, serial step 0.331 /
// 0.336 / 0.330 / 0.336 ms, gpurun_out/s2_bench_rep*.json), so there is ONE copy: 3.5 MB less cold DRAM traffic per
// 64 -> 64 layer.  -DPCDB_W_REPLICAS=n rebuilds the replicated variant.
#ifndef PCDB_W_REPLICAS
#define PCDB_W_REPLICAS 1
#endif
constexpr int kWReplicas = PCDB_W_REPLICAS;
constexpr int kProducerThreads = 128;
constexpr int kThreads = 192;      // 4 epilogue (and cp.async producer) warps + TMA warp + MMA/TMEM warp

// A pipeline stage always carries 64 input channels per tile row (128 B, SWIZZLE_128B, 4 MMA K-steps): ONE kernel
// offset for Cin = 64, TWO for Cin = 32, FOUR for Cin = 16, side by side along K -- the per-stage cost (barrier
// round trip, producer loop overhead, MMA-thread wake-up, ~0.3 us) is what these small layers were paying 27
// times per tile.  Each K-step still carries the disable-output-lane mask of ITS offset.
template <int CIN, int COUT>
struct Cfg {
    static_assert(CIN == 16 || CIN == 32 || CIN == 64, "CIN must be 16, 32 or 64");
    static_assert(COUT % 16 == 0 && COUT >= 16 && COUT <= 256, "COUT must be a multiple of 16 in [16, 256]");
    static constexpr int kGroup = 64 / CIN;                           // kernel offsets per stage
    static constexpr int kNumGroups = (kMaxK + kGroup - 1) / kGroup;  // 27 / 14 / 7
    static constexpr int kOffBytes = CIN * 2;                         // one offset's slice of a stage row: 128 / 64 / 32 B
    static constexpr int kCpo = kOffBytes / 16;                       // 16-byte pieces per (row, offset): 8 / 4 / 2
    static constexpr int kKStepsPerOff = CIN / 16;                    // tcgen05.mma K = 16 for bf16
    static constexpr int kRowBytes = 128;                             // one K-major operand row of a stage
    static constexpr int kChunks = kRowBytes / 16;
    static constexpr int kSwizzleBits = 3;
    static constexpr uint64_t kLayoutType = 2;                        // SWIZZLE_128B
    static constexpr int kABytes = kTileM * kRowBytes;
    static constexpr int kBBytes = (COUT * kRowBytes + 1023) / 1024 * 1024;
    static constexpr int kStageBytes = kABytes + kBBytes;
    static constexpr int kMaxStages = 8;      // barrier slots reserved; the ring depth is chosen per launch
    static constexpr int kTmemCols = COUT <= 32 ? 32 : (COUT <= 64 ? 64 : (COUT <= 128 ? 128 : 256));
    static constexpr int kNbrBytes = (2 * kMaxStages + 2) * 8 + 16 + 28 * 16;     // barriers; tmem base, mask; row masks
    static constexpr int kSrcBytes = kMaxK * kTileM * 4;                // s_src (TMA variant only)
    // instruction descriptor: D=f32, A=B=bf16, both K-major, N>>3 at bit 17, M>>4 at bit 24
    static constexpr uint32_t kIdesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(COUT >> 3) << 17) |
                                       ((uint32_t)(kTileM >> 4) << 24);
    static constexpr int groups(int K) { return (K + kGroup - 1) / kGroup; }
};

// The CPO passes of one (stage, offset): in pass P a group of CPO lanes fetches the OFF_BYTES of the row owned by
// lane P of the group (segmented shuffle with an immediate source lane); dst[P] = where that row's piece lands.
template <int OFF_BYTES, int CPO, int P>
__device__ __forceinline__ void gather_passes(const uint32_t *dst, uint32_t stage_off, const uint8_t *feat_piece, int src_own)
{
    if constexpr (P < CPO) {
        gather_piece<OFF_BYTES>(dst[P] + stage_off, feat_piece, __shfl_sync(0xffffffffu, src_own, P, CPO));
        gather_passes<OFF_BYTES, CPO, P + 1>(dst, stage_off, feat_piece, src_own);
    }
}

template <int CIN, int COUT>
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr)
{
    using C = Cfg<CIN, COUT>;
    return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | (1ull << 16) | ((uint64_t)((8 * C::kRowBytes) >> 4) << 32) |
           (1ull << 46) | (C::kLayoutType << 61);
}

// features (n_in, CIN) bf16 (also described by tmap_feat when TMA); w_packed: per offset the swizzled
// shared-memory image of the (COUT x CIN) weight tile (pcdb_pack_conv_weights); nbr (K, ld);
// out (n_out, COUT) bf16.
//
// Warp roles: 0-3 gather producers (cp.async engine) and epilogue; 4 weight-tile bulk copies (and the
// gather4 issue when TMA); 5 TMEM allocation + single-thread MMA issue.
template <int CIN, int COUT, bool TMA>
__global__ void __launch_bounds__(kThreads)
conv_fwd_tc(const __grid_constant__ CUtensorMap tmap_feat, const __nv_bfloat16 *__restrict__ feat, int n_in,
            const uint8_t *w_packed, const int *__restrict__ nbr, int ld, int K, int n_out,
            const int *__restrict__ n_out_dev, const float *__restrict__ scale, const float *__restrict__ shift,
            const float *__restrict__ bias, const __nv_bfloat16 *__restrict__ residual, int flags, __nv_bfloat16 *__restrict__ out,
            int n_stages)
{
    using C = Cfg<CIN, COUT>;
    extern __shared__ uint8_t smem_raw[];
    w_packed += (size_t)(blockIdx.x % kWReplicas) * (size_t)C::groups(K) * C::kBBytes;     // this CTA's weight replica
    if (n_out_dev) { const int m = __ldg(n_out_dev); n_out = m < n_out ? m : n_out; }
    // Programmatic dependent launch: let the next kernel of the stream (the next layer) be scheduled as soon as
    // every CTA of this one has started, so that its set-up overlaps this kernel's tail ...
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    const int row0 = blockIdx.x * kTileM;
    if (row0 >= n_out) return;
    if (threadIdx.x == 0) TRACE(6, 0);

    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t *aligned = smem_raw + (base - smem_u32(smem_raw));
    // TMA variant only: [kMaxK][128] input row per (offset, tile row); the cp.async producers keep their
    // slice of the rulebook in registers and exchange it with warp shuffles (no shared memory -> 4 CTAs/SM)
    int *s_src = reinterpret_cast<int *>(aligned + n_stages * C::kStageBytes);
    uint64_t *bars = reinterpret_cast<uint64_t *>(s_src + (TMA ? kMaxK * kTileM : 0));
    // bars[0..8) full, bars[8..16) empty, bars[16] accumulator ready; then tmem base and tile mask
    uint32_t *s_tmem = reinterpret_cast<uint32_t *>(bars + 2 * C::kMaxStages + 2);
    uint32_t *s_mask = s_tmem + 1;
    uint32_t *s_off = s_tmem + 4;          // [kMaxK][4] disable-output-lane words per offset (16-byte aligned)
    const uint32_t bar_full = smem_u32(bars), bar_empty = smem_u32(bars + C::kMaxStages), bar_acc = smem_u32(bars + 2 * C::kMaxStages);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        for (int s = 0; s < n_stages; ++s) {
            // full: every producer thread's async arrival (cp.async engine) + the weight copy's expect_tx arrival
            mbar_init(bar_full + 8 * s, TMA ? 1 : kProducerThreads + 1);
            mbar_init(bar_empty + 8 * s, 1);
        }
        mbar_init(bar_acc, 1);
        *s_mask = 0u;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // ---- this tile's slice of the rulebook: 27 independent loads per thread (thread t <-> tile row t <-> TMEM
    //      lane t), issued before the first barrier so that they overlap the mbarrier / TMEM set-up ----------
    int src_reg[kMaxK];
    if (warp < 4) {
        const int row = row0 + tid;
#pragma unroll
        for (int k = 0; k < kMaxK; ++k) src_reg[k] = (k < K && row < n_out) ? __ldg(nbr + (size_t)k * ld + row) : -1;
    }
    if (warp == 5) tmem_alloc(smem_u32(s_tmem), C::kTmemCols);
    tc_fence_before();
    __syncthreads();        // barriers initialised, *s_mask cleared, TMEM base published
    tc_fence_after();
    const uint32_t tmem = uniform(*s_tmem);
    if (warp < 4) {
        // Per offset: which of this warp's 32 rows have a neighbour.  The complement goes to the MMA issuer as
        // the disable-output-lane vector, so rows without a neighbour are neither fetched nor zero-filled.
        uint32_t mine = 0;
#pragma unroll
        for (int k = 0; k < kMaxK; ++k) {
            if (TMA) s_src[k * kTileM + tid] = src_reg[k];
            const uint32_t have = __ballot_sync(0xffffffffu, src_reg[k] >= 0);
            if (lane == 0) s_off[k * 4 + warp] = TMA ? 0u : ~have;      // TMA zero-fills instead (out-of-bounds rows)
            mine |= (have ? 1u : 0u) << k;
        }
        if (lane == 0 && mine) atomicOr(s_mask, mine);     // which offsets the tile touches at all
        // the accumulator starts from zero and every MMA accumulates (a masked row is never written)
#pragma unroll
        for (int c0 = 0; c0 < COUT; c0 += 16) tmem_zero16(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0);
        tmem_st_wait();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    // ... and wait here, after the set-up (barriers, TMEM, rulebook slice, row masks), for the previous kernel of
    // the stream to have completed: its output is this layer's `feat`, and `out` may be a buffer it still reads.
    asm volatile("griddepcontrol.wait;" ::: "memory");
    const uint32_t mask = uniform(*s_mask);
    const int n_iter = __popc(mask);

    if (warp < 4) {
        if (!TMA) {
            // ===== gather producers (cp.async engine) ================================================
            // kChunks consecutive lanes fetch the 16-byte pieces of ONE input row, so a warp-wide cp.async
            // touches 32/kChunks cache lines instead of 32; in pass p lane group j of warp w fetches tile row
            // 32w + j*kChunks + p, whose rulebook entry sits in lane j*kChunks + p of the same warp.
            // Measured on B200: the kernel's pace is set by the LDGSTS pipe (~14 cycles per warp-wide 16-byte
            // copy per SM, whether or not it zero-fills) and, with few CTAs per SM, by the producers' dependent
            // instruction chain.  Hence (1) rows without a neighbour issue NO copy (predicated off; the MMA masks
            // the row), and (2) the offset loop is fully unrolled with a uniform skip -- src_reg[k] is a fixed
            // register -- so that a piece costs SHFL (immediate lane, segment width), ISETP, LEA, LEA.HI.X, IADD,
            // LDGSTS.
            constexpr int kCpo = C::kCpo;
            const int piece = lane % kCpo, jw = lane / kCpo;
            const uint8_t *feat_b = reinterpret_cast<const uint8_t *>(feat) + piece * 16;
            // dst[sub * kCpo + p]: where pass p of the stage's sub-th offset puts this lane's piece
            uint32_t dst[C::kGroup * kCpo];
#pragma unroll
            for (int sub = 0; sub < C::kGroup; ++sub)
#pragma unroll
                for (int p = 0; p < kCpo; ++p)
                    dst[sub * kCpo + p] = base + swizzled_offset<C::kRowBytes, C::kSwizzleBits>(32 * warp + jw * kCpo + p,
                                                                                                sub * kCpo + piece);
            uint32_t stage_off = 0, bf = bar_full, be = bar_empty;
            int s = 0;
            uint32_t empty_parity = 0;         // parity of the (w-1)-th completion at ring wrap w
            bool first_pass = true;            // first pass over the ring: nothing to wait for
#pragma unroll
            for (int g = 0; g < C::kNumGroups; ++g) {
                if (!((mask >> (g * C::kGroup)) & ((1u << C::kGroup) - 1u))) continue;      // uniform over the CTA
                if (!first_pass) mbar_wait(be, empty_parity);
#pragma unroll
                for (int sub = 0; sub < C::kGroup; ++sub) {
                    constexpr int kDummy = 0; (void)kDummy;
                    const int k = g * C::kGroup + sub;
                    if (k < kMaxK && ((mask >> k) & 1u))          // an untouched offset's columns are never read
                        gather_passes<C::kOffBytes, kCpo, 0>(dst + sub * kCpo, stage_off, feat_b, src_reg[k < kMaxK ? k : 0]);
                }
                cp_async_arrive(bf);
                stage_off += C::kStageBytes; bf += 8; be += 8;
                if (++s == n_stages) {
                    s = 0;
                    stage_off = 0; bf = bar_full; be = bar_empty;
                    if (first_pass) first_pass = false; else empty_parity ^= 1u;
                }
            }
        }
        // ===== epilogue: warp w owns TMEM lanes [32w, 32w+32) = tile rows ===========================
        // accumulator -> registers -> scale/shift/ReLU -> bf16 -> shared memory (swizzled) -> one linear,
        // fully coalesced copy of the tile to global memory (output rows are contiguous).
        if (n_iter > 0) {
            mbar_wait(bar_acc, 0);
            tc_fence_after();
        }
        if (tid == 0) TRACE(6, 1);
        const bool relu = flags & PCDB_EPI_RELU;
        constexpr int kOutPitch = COUT * 2;
        const uint32_t o_base = base;       // the operand stages are free once the accumulator is complete
#pragma unroll 1
        for (int c0 = 0; c0 < COUT; c0 += 16) {
            uint32_t r[16];
            tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, r);
            tmem_ld_wait();
            // optional residual row (n_out, COUT) bf16: added to the accumulator (a partial sum over more input channels than
            // one launch takes) or, with PCDB_EPI_RESIDUAL_POST, behind the BatchNorm (the shortcut of a residual block)
            uint32_t res[8] = {0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};
            if (residual && row0 + tid < n_out) {
                const uint4 *rp = reinterpret_cast<const uint4 *>(residual + (size_t)(row0 + tid) * COUT + c0);
                const uint4 a = __ldg(rp), b = __ldg(rp + 1);
                res[0] = a.x; res[1] = a.y; res[2] = a.z; res[3] = a.w; res[4] = b.x; res[5] = b.y; res[6] = b.z; res[7] = b.w;
            }
            const bool res_post = flags & PCDB_EPI_RESIDUAL_POST;
            uint32_t packed[8];
#pragma unroll
            for (int j = 0; j < 16; j += 2) {
                float y0 = __uint_as_float(r[j]), y1 = __uint_as_float(r[j + 1]);
                const float2 rr = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162 *>(&res[j >> 1]));
                if (!res_post) { y0 += rr.x; y1 += rr.y; }
                const float s0 = scale ? __ldg(scale + c0 + j) : 1.f, s1 = scale ? __ldg(scale + c0 + j + 1) : 1.f;
                float h0 = shift ? __ldg(shift + c0 + j) : 0.f, h1 = shift ? __ldg(shift + c0 + j + 1) : 0.f;
                if (bias) { h0 = fmaf(__ldg(bias + c0 + j), s0, h0); h1 = fmaf(__ldg(bias + c0 + j + 1), s1, h1); }     // (y + bias) * scale + shift
                y0 = fmaf(y0, s0, h0); y1 = fmaf(y1, s1, h1);
                if (res_post) { y0 += rr.x; y1 += rr.y; }
                if (relu) { y0 = fmaxf(y0, 0.f); y1 = fmaxf(y1, 0.f); }
                __nv_bfloat162 pk = __floats2bfloat162_rn(y0, y1);
                packed[j >> 1] = *reinterpret_cast<uint32_t *>(&pk);
            }
            const uint32_t o0 = (uint32_t)tid * kOutPitch + (uint32_t)c0 * 2;
            st_shared_v4(o_base + swizzle_out(o0), packed[0], packed[1], packed[2], packed[3]);
            st_shared_v4(o_base + swizzle_out(o0 + 16), packed[4], packed[5], packed[6], packed[7]);
        }
        asm volatile("bar.sync 1, %0;" ::"n"(kProducerThreads) : "memory");     // the 4 epilogue warps only
        const int rows_here = min(kTileM, n_out - row0);
        const uint32_t valid_bytes = (uint32_t)rows_here * kOutPitch;
        uint8_t *gdst = reinterpret_cast<uint8_t *>(out + (size_t)row0 * COUT);
#pragma unroll
        for (uint32_t o = (uint32_t)tid * 16; o < (uint32_t)kTileM * kOutPitch; o += kProducerThreads * 16) {
            if (o < valid_bytes) {
                uint32_t v0, v1, v2, v3;
                ld_shared_v4(o_base + swizzle_out(o), v0, v1, v2, v3);
                *reinterpret_cast<uint4 *>(gdst + o) = make_uint4(v0, v1, v2, v3);
            }
        }
    } else if (warp == 4) {
        // ===== weight tiles by bulk copy; with TMA (Cin = 64 only) also the row gather (lane l: rows 4l..4l+3) ====
        int s = 0;
        uint32_t empty_parity = 0;
        bool first_pass = true;
        for (int g = 0; g < C::kNumGroups; ++g) {
            if (!((mask >> (g * C::kGroup)) & ((1u << C::kGroup) - 1u))) continue;
            if (!first_pass) mbar_wait(bar_empty + 8 * s, empty_parity);
            const uint32_t a_base = base + s * C::kStageBytes, b_base = a_base + C::kABytes;
            if (elect_one()) {
                mbar_arrive_expect_tx(bar_full + 8 * s, (TMA ? C::kABytes : 0) + COUT * C::kRowBytes);
                bulk_copy_g2s(b_base, w_packed + (size_t)g * C::kBBytes, COUT * C::kRowBytes, bar_full + 8 * s);
            }
            if (TMA) {
                __syncwarp();
                const int4 idx = *reinterpret_cast<const int4 *>(s_src + g * kTileM + 4 * lane);
                // a missing neighbour becomes row n_in, which is outside the tensor map: TMA writes zeros
                tma_gather4(a_base + lane * 4 * C::kRowBytes, &tmap_feat, bar_full + 8 * s, 0, idx.x >= 0 ? idx.x : n_in,
                            idx.y >= 0 ? idx.y : n_in, idx.z >= 0 ? idx.z : n_in, idx.w >= 0 ? idx.w : n_in);
            }
            if (++s == n_stages) {
                s = 0;
                if (first_pass) first_pass = false; else empty_parity ^= 1u;
            }
        }
    } else {
        // ===== MMA issuer: warp 5 stays converged, one elected lane issues ==============================
        // descriptors differ between stages only in the 14-bit start-address field: build them once
        const uint64_t desc_a0 = make_desc<CIN, COUT>(base), desc_b0 = make_desc<CIN, COUT>(base + C::kABytes);
        int s = 0, it = 0;
        uint32_t full_parity = 0;
        for (int g = 0; g < C::kNumGroups; ++g) {
            const uint32_t gm = (mask >> (g * C::kGroup)) & ((1u << C::kGroup) - 1u);
            if (!gm) continue;
            // the disable-output-lane words of the stage's offsets: lane 4*sub + i reads word i of offset sub,
            // broadcasts make them warp-uniform
            const uint32_t word = lane < 4 * C::kGroup ? s_off[4 * (g * C::kGroup) + lane] : 0u;
            mbar_wait(bar_full + 8 * s, full_parity);
            if (lane == 0) TRACE(2, g);
            tc_fence_after();
            const uint64_t step = (uint64_t)((s * C::kStageBytes) >> 4);
#pragma unroll
            for (int sub = 0; sub < C::kGroup; ++sub) {
                uint4 off;
                off.x = __shfl_sync(0xffffffffu, word, 4 * sub + 0); off.y = __shfl_sync(0xffffffffu, word, 4 * sub + 1);
                off.z = __shfl_sync(0xffffffffu, word, 4 * sub + 2); off.w = __shfl_sync(0xffffffffu, word, 4 * sub + 3);
                if (((gm >> sub) & 1u) && elect_one()) {
#pragma unroll
                    for (int js = 0; js < C::kKStepsPerOff; ++js) {
                        const int j = sub * C::kKStepsPerOff + js;
                        umma_bf16(tmem, desc_a0 + step + 2 * j, desc_b0 + step + 2 * j, C::kIdesc, off);
                    }
                }
            }
            if (elect_one()) umma_commit(bar_empty + 8 * s);      // stage reusable once these MMAs have read it
            if (lane == 0) TRACE(3, g);
            if (++s == n_stages) { s = 0; full_parity ^= 1u; }
            ++it;
        }
        if (it > 0 && elect_one()) umma_commit(bar_acc);        // accumulator complete
    }
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x == 0) TRACE(6, 2);
    if (warp == 5) {
        __syncwarp();
        tmem_dealloc(tmem, C::kTmemCols);
    }
}

// ---- weight packing -----------------------------------------------------------------------------
// (K, CIN, COUT) row-major bf16 -> per GROUP of kGroup offsets the K-major, 128B-swizzled (COUT rows x 64) image the
// stage's B operand is: chunk c of row n holds 8 input channels of offset g*kGroup + c / kCpo; kBBytes apart.
// SRC: bf16 or fp32 master weights (converted here).  transpose: `w` is the (K, COUT, CIN) weight of the FORWARD layer and
// the image is that of its input-gradient convolution (W[k]^T); flip: offsets reversed (k -> K-1-k), which is how a
// centred submanifold rulebook reads the other way round (SparseConvFunction).
template <int CIN, int COUT, typename SRC>
__global__ void pack_weights_kernel(const SRC *__restrict__ w, int K, int transpose, int flip, uint8_t *__restrict__ packed)
{
    using C = Cfg<CIN, COUT>;
    const int G = C::groups(K);
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= G * COUT * C::kChunks) return;
    const int c = t % C::kChunks, n = (t / C::kChunks) % COUT, g = t / (C::kChunks * COUT);
    const int k = g * C::kGroup + c / C::kCpo, piece = c % C::kCpo;
    const int ks = flip ? K - 1 - k : k;
    __nv_bfloat16 v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int ci = piece * 8 + j;
        const size_t at = transpose ? ((size_t)ks * COUT + n) * CIN + ci : ((size_t)ks * CIN + ci) * COUT + n;
        v[j] = k < K ? from_float<__nv_bfloat16>(to_float(w[at])) : __float2bfloat16(0.f);
    }
    for (int rep = 0; rep < kWReplicas; ++rep)
        *reinterpret_cast<uint4 *>(packed + ((size_t)rep * G + g) * C::kBBytes + swizzled_offset<C::kRowBytes, C::kSwizzleBits>(n, c)) =
            *reinterpret_cast<const uint4 *>(v);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn()
{
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    }
    return fn;
}

template <int CIN, int COUT>
int launch(const void *features, int n_in, const void *w_packed, const int32_t *nbr, int ld, int K, int n_out,
           const int32_t *n_out_dev, const float *scale, const float *shift, const float *bias, const void *residual, int flags,
           void *out, bool use_tma, int rows_hint, cudaStream_t stream)
{
    using C = Cfg<CIN, COUT>;
    if (CIN != 64) use_tma = false;       // gather4 writes whole rows: only a one-offset stage (Cin = 64) has that layout
    CUtensorMap tmap;
    memset(&tmap, 0, sizeof(tmap));
    if (use_tma) {
        EncodeTiledFn enc = encode_fn();
        if (!enc) { set_last_error("tcgen05 sparse conv: cuTensorMapEncodeTiled is unavailable"); return kCudaError; }
        const cuuint64_t gdim[2] = {(cuuint64_t)CIN, (cuuint64_t)(n_in > 0 ? n_in : 1)};
        const cuuint64_t gstride[1] = {(cuuint64_t)CIN * 2};
        const cuuint32_t box[2] = {(cuuint32_t)CIN, 1u};       // gather4 fetches four such one-row boxes
        const cuuint32_t estr[2] = {1u, 1u};
        const CUtensorMapSwizzle sw = CIN == 64 ? CU_TENSOR_MAP_SWIZZLE_128B
                                                : (CIN == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
        const CUresult r = enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void *>(features), gdim, gstride, box,
                               estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { set_last_error("tcgen05 sparse conv: cuTensorMapEncodeTiled failed (%d)", (int)r); return kCudaError; }
    }
    const int tiles = (n_out + kTileM - 1) / kTileM;
    // Ring depth.  What is in flight per SM is bounded by its shared memory (stages x CTAs), and a stage's round trip
    // (fill, land, MMA, release: ~0.8 us) is what has to be covered: few tiles per SM -> one CTA with a deep ring,
    // many tiles -> more CTAs with shallow rings (measured on B200).  rows_hint is the number of
    // output rows the caller expects when n_out is only a capacity.
    const int rows = rows_hint > 0 && rows_hint < n_out ? rows_hint : n_out;
    const int tiles_exp = (rows + kTileM - 1) / kTileM;
    int ctas = (tiles_exp + kNumSMs - 1) / kNumSMs;
    ctas = ctas < 1 ? 1 : (ctas > 5 ? 5 : ctas);
    const int fixed = 1024 + C::kNbrBytes + 256 + (use_tma ? C::kSrcBytes : 0);
    int n_stages = (233472 / ctas - 1024 - fixed) / C::kStageBytes;
    // PCDB_CONV_SHALLOW_RING: two stages.  Costs ~4 % when the kernel has the GPU to itself (KITTI step 0.382 -> 0.396 ms)
    // and wins ~5 % when several steps are in flight (12 680 -> 13 370 frames/s at 4 in flight): the smaller footprint
    // lets the CTAs of the other steps' kernels become resident next to the convolution's.
    const int cap = (flags & PCDB_CONV_SHALLOW_RING) ? 2 : (ctas == 1 ? C::kMaxStages : 4);
    n_stages = n_stages > cap ? cap : (n_stages < 2 ? 2 : n_stages);
    while (n_stages * C::kStageBytes < kTileM * COUT * 2) ++n_stages;          // the epilogue stages the tile in the ring
    const int smem = fixed + n_stages * C::kStageBytes;
    // per device and context, cheap: set on every call (a process-wide "largest so far" cache breaks on a second GPU)
    cudaError_t attr_err = use_tma ? cudaSuccess
                                   : cudaFuncSetAttribute(conv_fwd_tc<CIN, COUT, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if constexpr (CIN == 64)
        if (use_tma) attr_err = cudaFuncSetAttribute(conv_fwd_tc<CIN, COUT, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (attr_err != cudaSuccess) {
        set_last_error("pcdb_sparse_conv_fwd(tcgen05): cudaFuncSetAttribute(%d bytes) failed: %s", smem, cudaGetErrorString(attr_err));
        return kCudaError;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(tiles);
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = (flags & PCDB_CONV_PDL) ? 1 : 0;
    const __nv_bfloat16 *f = (const __nv_bfloat16 *)features;
    const uint8_t *wp = (const uint8_t *)w_packed;
    __nv_bfloat16 *o = (__nv_bfloat16 *)out;
    const __nv_bfloat16 *rs = (const __nv_bfloat16 *)residual;
    const int *nb = nbr, *nd = n_out_dev;
    cudaError_t err = cudaSuccess;
    if constexpr (CIN == 64) {
        if (use_tma)
            err = cudaLaunchKernelEx(&cfg, conv_fwd_tc<CIN, COUT, true>, tmap, f, n_in, wp, nb, ld, K, n_out, nd, scale, shift,
                                     bias, rs, flags, o, n_stages);
    }
    if (!use_tma)
        err = cudaLaunchKernelEx(&cfg, conv_fwd_tc<CIN, COUT, false>, tmap, f, n_in, wp, nb, ld, K, n_out, nd, scale, shift, bias,
                                 rs, flags, o, n_stages);
    if (err != cudaSuccess) {
        set_last_error("pcdb_sparse_conv_fwd(tcgen05): launch failed: %s", cudaGetErrorString(err));
        return kCudaError;
    }
    return check_launch("pcdb_sparse_conv_fwd(tcgen05)");
}

}  // namespace tc

bool conv_tc_supported(int c_in, int c_out, int K)
{
    if (K > tc::kMaxK) return false;
    const bool cin_ok = c_in == 16 || c_in == 32 || c_in == 64;
    const bool cout_ok = c_out == 16 || c_out == 32 || c_out == 64 || c_out == 128;
    return cin_ok && cout_ok;
}

#define PCDB_TC_SHAPES(X) \
    X(16, 16) X(16, 32) X(16, 64) X(16, 128) X(32, 16) X(32, 32) X(32, 64) X(32, 128) X(64, 16) X(64, 32) X(64, 64) X(64, 128)

size_t conv_tc_packed_bytes(int c_in, int c_out, int K)
{
#define PCDB_TC_CASE(CI, CO) if (c_in == CI && c_out == CO) return (size_t)tc::kWReplicas * tc::Cfg<CI, CO>::groups(K) * tc::Cfg<CI, CO>::kBBytes;
    PCDB_TC_SHAPES(PCDB_TC_CASE)
#undef PCDB_TC_CASE
    return 0;
}

// Every 16-byte unit of the image is written (offsets past K as zeros), so the buffer needs no clear.
// c_in / c_out are those of the convolution the image is FOR (with PCDB_PACK_TRANSPOSE the source weight is (K, c_out, c_in))
int conv_tc_pack_weights(const void *weight, int dtype, int K, int c_in, int c_out, int flags, void *packed, cudaStream_t stream)
{
    const int tr = (flags & PCDB_PACK_TRANSPOSE) ? 1 : 0, fl = (flags & PCDB_PACK_FLIP) ? 1 : 0;
#define PCDB_TC_CASE(CI, CO) \
    if (c_in == CI && c_out == CO) { \
        const int total = tc::Cfg<CI, CO>::groups(K) * CO * tc::Cfg<CI, CO>::kChunks; \
        if (dtype == PCDB_BF16) \
            tc::pack_weights_kernel<CI, CO, __nv_bfloat16><<<(total + 255) / 256, 256, 0, stream>>>((const __nv_bfloat16 *)weight, K, tr, fl, (uint8_t *)packed); \
        else \
            tc::pack_weights_kernel<CI, CO, float><<<(total + 255) / 256, 256, 0, stream>>>((const float *)weight, K, tr, fl, (uint8_t *)packed); \
        return check_launch("pcdb_pack_conv_weights"); \
    }
    PCDB_TC_SHAPES(PCDB_TC_CASE)
#undef PCDB_TC_CASE
    set_last_error("pcdb_pack_conv_weights: unsupported channels c_in=%d c_out=%d", c_in, c_out);
    return kUnsupported;
}

int launch_conv_fwd_tc(const void *features, int n_in, const void *w_packed, const int32_t *nbr, int ld, int K, int n_out,
                       const int32_t *n_out_dev, int c_in, int c_out, const float *scale, const float *shift,
                       const float *bias, const void *residual, int flags, void *out, bool use_tma, int rows_hint, cudaStream_t stream)
{
#define PCDB_TC_CASE(CI, CO) \
    if (c_in == CI && c_out == CO) \
        return tc::launch<CI, CO>(features, n_in, w_packed, nbr, ld, K, n_out, n_out_dev, scale, shift, bias, residual, flags, out, \
                                  use_tma, rows_hint, stream);
    PCDB_TC_SHAPES(PCDB_TC_CASE)
#undef PCDB_TC_CASE
    set_last_error("tcgen05 sparse conv: unsupported channels c_in=%d c_out=%d", c_in, c_out);
    return kUnsupported;
}

#ifdef PCDB_TC_TRACE
extern "C" int pcdb_debug_trace(long long *host) { return (int)cudaMemcpyFromSymbol(host, tc::g_trace, sizeof(tc::g_trace)); }
#endif

}  // namespace pcdb
