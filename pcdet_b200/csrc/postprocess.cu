// Front of the detector's post-processing on sm_100a: class max + sigmoid threshold + top-k + box decoding.
//
// Replaces, for the class-agnostic path of SECOND / PointPillars (MULTI_CLASSES_NMS False):
//   Detector3D.predict_boxes          pcdet/models/detectors/detector3d.py:112-128   (decode ALL anchors)
//   Detector3D.post_processing        detector3d.py:166-215                           (sigmoid, class max)
//   Detector3D.class_agnostic_nms     detector3d.py:278-290                           (mask, topk, BEV boxes)
//   ResidualCoder.decode_torch / decode_with_head_direction_torch   pcdet/utils/box_coder_utils.py:89-144
//   boxes3d_to_bevboxes_lidar_torch   pcdet/utils/box_utils.py:237-250
// The reference decodes all 211 200 anchors of a frame (~20 elementwise launches), builds a boolean mask, compacts
// with it (nonzero: a device->host sync per frame), runs torch.topk and gathers.  Here only the class scores of all
// anchors are read (12 B per anchor); the k-th largest score is found by a 3-pass radix select (12/10/10 bits of an
// order-preserving 32-bit key) whose histograms are reduced by the last block of each pass; the selected anchors
// are compacted in anchor order (deterministic, no atomics: ties at the k-th score go to the lower anchor index),
// ranked by (score desc, anchor asc) by counting in shared memory, and only those <= k anchors are decoded.  Everything stays on
// the device; counts are device scalars; the outputs are capacity-sized and feed pcdb_nms directly.
#include "common.cuh"
#include "../../include/pcdet_b200.h"

namespace pcdb {

constexpr int kPpBlock = 256;
constexpr int kBins0 = 4096, kBins1 = 1024, kBins2 = 1024;
constexpr int kHistWords = kBins0 + kBins1 + kBins2;

struct PpState {
    int bin0, rem0;          // level 0: bucket of the k-th key, how many of the k are still to be found inside it
    int bin1, rem1;          // level 1
    uint32_t thresh_key;     // key of the k-th candidate (0: every candidate is selected)
    int need_eq;             // how many candidates with key == thresh_key are selected (lowest anchor indices)
    int count;               // selected candidates = min(k, candidates)
    int all;                 // fewer than k candidates: all of them are selected
};

struct PpWorkspace {
    uint32_t *hist;          // [batch][kHistWords]   zeroed per call
    unsigned int *done;      // [batch][4]            zeroed per call (tickets of the three histogram passes)
    PpState *state;          // [batch]
    int2 *blk_counts;        // [batch][nblk]         (keys > T, keys == T) per block
    uint32_t *keys;          // [batch][n_anchors]
    unsigned long long *sel; // [batch][pre_max]      key << 32 | ~anchor
    size_t zero_bytes, bytes;
};

// Anchors per block, a multiple of the block size: `target` when that needs at most `max_blocks` blocks per frame.
static int span_for(int n_anchors, int target, int max_blocks)
{
    int span = target;
    if ((n_anchors + span - 1) / span > max_blocks) span = ((n_anchors + max_blocks - 1) / max_blocks + kPpBlock - 1) / kPpBlock * kPpBlock;
    return span;
}
static int blocks_for(int n_anchors, int span) { return (n_anchors + span - 1) / span; }
// score pass: 4096-bin shared histogram per block -> long spans; refinement: 1024 bins; compaction: short spans, but the
// block-count prefix of pp_scatter is linear in the number of blocks
constexpr int kStageFloats = 6144;          // 24 KB of logits per block next to the 16 KB histogram
static bool score_staged(int cls_stride) { return cls_stride * kPpBlock <= kStageFloats; }
static int score_span(int n_anchors, int cls_stride)
{
    int target = 2048;
    if (score_staged(cls_stride) && target * cls_stride > kStageFloats) target = kStageFloats / cls_stride / kPpBlock * kPpBlock;
    return span_for(n_anchors, target, 512);
}
static int refine_span(int n_anchors) { return span_for(n_anchors, 1024, 1024); }
static int compact_span(int n_anchors) { return span_for(n_anchors, 512, 1024); }

static PpWorkspace carve_pp(void *base, int batch, int n_anchors, int pre_max)
{
    PpWorkspace w;
    char *p = (char *)base;
    size_t off = 0;
    auto take = [&](size_t bytes) { char *q = p + off; off += align_up(bytes, 256); return q; };
    w.hist = (uint32_t *)take(sizeof(uint32_t) * (size_t)batch * kHistWords);
    w.done = (unsigned int *)take(sizeof(unsigned int) * (size_t)batch * 4);
    w.zero_bytes = off;
    w.state = (PpState *)take(sizeof(PpState) * (size_t)batch);
    w.blk_counts = (int2 *)take(sizeof(int2) * (size_t)batch * blocks_for(n_anchors, compact_span(n_anchors)));
    w.keys = (uint32_t *)take(sizeof(uint32_t) * (size_t)batch * n_anchors);
    w.sel = (unsigned long long *)take(sizeof(unsigned long long) * (size_t)batch * pre_max);
    w.bytes = off;
    return w;
}

// order-preserving map float -> uint32 (larger score <=> larger key); no finite score maps to 0; -0.0 and +0.0 share a key
__device__ __forceinline__ uint32_t score_key(float x)
{
    const uint32_t u = x == 0.f ? 0u : __float_as_uint(x);
    return u ^ ((u >> 31) ? 0xFFFFFFFFu : 0x80000000u);
}
__device__ __forceinline__ float key_score(uint32_t k)
{
    return __uint_as_float((k & 0x80000000u) ? k ^ 0x80000000u : ~k);
}

// max over the classes and its FIRST index (torch.max(dim=-1), detector3d.py:195)
__device__ __forceinline__ float class_max(const float *__restrict__ p, int n_classes, int *label)
{
    float best = __ldg(p);
    int arg = 0;
    for (int c = 1; c < n_classes; ++c) {
        const float v = __ldg(p + c);
        if (v > best) { best = v; arg = c; }
    }
    *label = arg;
    return best;
}

// true in exactly one block per (frame, pass): the last one to get here; its reads see every other block's atomics
__device__ __forceinline__ bool last_block_of_frame(unsigned int *done, int nblk)
{
    __shared__ bool s_last;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = atomicAdd(done, 1u) == (unsigned int)(nblk - 1);
    __syncthreads();
    if (s_last) __threadfence();
    return s_last;
}

// Walks a histogram from its top bin down to the bin holding the k-th largest key: *bin (or -1 when the histogram
// holds fewer than k keys), *rem = how many of the k lie inside that bin, *total = keys in the histogram.
template <int BINS>
__device__ __forceinline__ void pick_level(const uint32_t *hist, int k, int *bin, int *rem, int *total)
{
    constexpr int PER = BINS / kPpBlock;
    __shared__ int s_bin, s_rem;
    if (threadIdx.x == 0) { s_bin = -1; s_rem = 0; }
    int local[PER];
    int sum = 0;
    const int hi = BINS - 1 - (int)threadIdx.x * PER;
#pragma unroll
    for (int j = 0; j < PER; ++j) { local[j] = (int)__ldcg(hist + hi - j); sum += local[j]; }
    const int excl = block_exclusive_scan<kPpBlock>(sum, total);
    if (excl < k && k <= excl + sum) {
        int cum = excl;
#pragma unroll
        for (int j = 0; j < PER; ++j) {
            if (k > cum && k <= cum + local[j]) { s_bin = hi - j; s_rem = k - cum; }
            cum += local[j];
        }
    }
    __syncthreads();
    *bin = s_bin;
    *rem = s_rem;
}

// class_max with the class count known at compile time (N_CLASSES 1..4; 0 = run-time count): the loads of the
// unrolled anchors below are then all issued before the first one is used
template <int N_CLASSES>
__device__ __forceinline__ float class_max_fixed(const float *__restrict__ p, int n_classes)
{
    if (N_CLASSES == 0) { int label; return class_max(p, n_classes, &label); }
    float v[N_CLASSES > 0 ? N_CLASSES : 1];
#pragma unroll
    for (int c = 0; c < N_CLASSES; ++c) v[c] = __ldg(p + c);
    float best = v[0];
#pragma unroll
    for (int c = 1; c < N_CLASSES; ++c) best = v[c] > best ? v[c] : best;
    return best;
}

// grid: (nblk, batch).  Block blk of a frame owns the anchors [blk*span, (blk+1)*span).
// STAGED: the block's span * cls_stride logits are first copied to shared memory with coalesced, 8-deep unrolled loads
// (a thread reading its own anchors' 12-byte records keeps ~100 B in flight; the copy keeps ~25 KB per block in
// flight, which is what an HBM-bound pass over 10 MB needs), then every thread scores its anchors out of it.
template <int N_CLASSES, bool STAGED>
__global__ void __launch_bounds__(kPpBlock)
pp_score_hist(const float *__restrict__ cls, int n_anchors, int n_classes, int cls_stride, float score_thresh, int span,
              int k, uint32_t *__restrict__ keys, uint32_t *hist, unsigned int *done, PpState *state)
{
    extern __shared__ float s_cls[];
    __shared__ uint32_t s_hist[kBins0];
    const int b = blockIdx.y;
    const int begin = blockIdx.x * span, end = min(n_anchors, begin + span);
    const float *src = cls + ((size_t)b * n_anchors + begin) * cls_stride;
    if (STAGED) {
        const int total = (end - begin) * cls_stride;
#pragma unroll 8
        for (int t = threadIdx.x; t < total; t += kPpBlock) s_cls[t] = __ldg(src + t);
    }
    for (int j = threadIdx.x; j < kBins0; j += kPpBlock) s_hist[j] = 0u;
    __syncthreads();
    constexpr int kUnroll = 4;
    for (int i0 = begin + threadIdx.x; i0 < end; i0 += kUnroll * kPpBlock) {
        float best[kUnroll];
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
            const int i = i0 + u * kPpBlock;
            const float *p = STAGED ? s_cls + (size_t)(i - begin) * cls_stride : src + (size_t)(i - begin) * cls_stride;
            best[u] = 0.f;
            if (i < end) {
                if (STAGED) {
                    float m = p[0];
                    if (N_CLASSES > 0) {
#pragma unroll
                        for (int c = 1; c < N_CLASSES; ++c) m = p[c] > m ? p[c] : m;
                    } else {
                        for (int c = 1; c < n_classes; ++c) m = p[c] > m ? p[c] : m;
                    }
                    best[u] = m;
                } else {
                    best[u] = class_max_fixed<N_CLASSES>(p, n_classes);
                }
            }
        }
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
            const int i = i0 + u * kPpBlock;
            if (i >= end) break;
            const float prob = 1.f / (1.f + expf(-best[u]));                                // torch.sigmoid
            const uint32_t key = prob >= score_thresh ? score_key(best[u]) : 0u;            // detector3d.py:279
            keys[(size_t)b * n_anchors + i] = key;
            if (key) atomicAdd(&s_hist[key >> 20], 1u);
        }
    }
    __syncthreads();
    uint32_t *h0 = hist + (size_t)b * kHistWords;
    for (int j = threadIdx.x; j < kBins0; j += kPpBlock)
        if (s_hist[j]) atomicAdd(h0 + j, s_hist[j]);
    if (!last_block_of_frame(done + b * 4, gridDim.x)) return;
    int bin, rem, total;
    pick_level<kBins0>(h0, k, &bin, &rem, &total);
    if (threadIdx.x == 0) {
        PpState st;
        st.bin0 = bin; st.rem0 = rem; st.bin1 = 0; st.rem1 = 0;
        st.all = bin < 0;
        st.thresh_key = 0u; st.need_eq = 0;
        st.count = bin < 0 ? total : k;
        state[b] = st;
    }
}

// LEVEL 1: histogram of key bits 19..10 among the keys of bucket bin0; LEVEL 2: bits 9..0 among bucket (bin0,bin1)
template <int LEVEL>
__global__ void __launch_bounds__(kPpBlock)
pp_refine(const uint32_t *__restrict__ keys, int n_anchors, int span, uint32_t *hist, unsigned int *done, PpState *state)
{
    __shared__ uint32_t s_hist[kBins1];
    const int b = blockIdx.y;
    for (int j = threadIdx.x; j < kBins1; j += kPpBlock) s_hist[j] = 0u;
    const PpState st = state[b];
    if (st.all) return;
    __syncthreads();
    const uint32_t prefix = LEVEL == 1 ? (uint32_t)st.bin0 : ((uint32_t)st.bin0 << 10) | (uint32_t)st.bin1;
    constexpr int kShift = LEVEL == 1 ? 20 : 10;
    const int begin = blockIdx.x * span, end = min(n_anchors, begin + span);
    for (int i = begin + threadIdx.x; i < end; i += kPpBlock) {
        const uint32_t key = __ldg(keys + (size_t)b * n_anchors + i);
        if ((key >> kShift) == prefix) atomicAdd(&s_hist[(key >> (kShift - 10)) & 1023u], 1u);
    }
    __syncthreads();
    uint32_t *h = hist + (size_t)b * kHistWords + (LEVEL == 1 ? kBins0 : kBins0 + kBins1);
    for (int j = threadIdx.x; j < kBins1; j += kPpBlock)
        if (s_hist[j]) atomicAdd(h + j, s_hist[j]);
    if (!last_block_of_frame(done + b * 4 + LEVEL, gridDim.x)) return;
    int bin, rem, total;
    pick_level<kBins1>(h, LEVEL == 1 ? st.rem0 : st.rem1, &bin, &rem, &total);
    if (threadIdx.x == 0) {
        if (LEVEL == 1) {
            state[b].bin1 = bin;
            state[b].rem1 = rem;
        } else {
            state[b].thresh_key = (prefix << 10) | (uint32_t)bin;
            state[b].need_eq = rem;
        }
    }
}

__global__ void __launch_bounds__(kPpBlock)
pp_count(const uint32_t *__restrict__ keys, int n_anchors, int span, const PpState *__restrict__ state, int2 *__restrict__ blk_counts)
{
    __shared__ int s_gt[kPpBlock / 32], s_eq[kPpBlock / 32];
    const int b = blockIdx.y;
    const PpState st = state[b];
    const uint32_t T = st.thresh_key;
    const int begin = blockIdx.x * span, end = min(n_anchors, begin + span);
    int gt = 0, eq = 0;
    for (int i = begin + threadIdx.x; i < end; i += kPpBlock) {
        const uint32_t key = __ldg(keys + (size_t)b * n_anchors + i);
        gt += key > T;
        eq += !st.all && key == T;
    }
#pragma unroll
    for (int d = 16; d; d >>= 1) {
        gt += __shfl_xor_sync(0xffffffffu, gt, d);
        eq += __shfl_xor_sync(0xffffffffu, eq, d);
    }
    if ((threadIdx.x & 31) == 0) { s_gt[threadIdx.x >> 5] = gt; s_eq[threadIdx.x >> 5] = eq; }
    __syncthreads();
    if (threadIdx.x == 0) {
        int g = 0, e = 0;
        for (int w = 0; w < kPpBlock / 32; ++w) { g += s_gt[w]; e += s_eq[w]; }
        blk_counts[b * gridDim.x + blockIdx.x] = make_int2(g, e);
    }
}

// Ordered compaction: the candidates above the k-th key in anchor order, then the first need_eq candidates AT the
// k-th key in anchor order.
__global__ void __launch_bounds__(kPpBlock)
pp_scatter(const uint32_t *__restrict__ keys, int n_anchors, int span, const PpState *__restrict__ state,
           const int2 *__restrict__ blk_counts, unsigned long long *__restrict__ sel, int sel_stride)
{
    __shared__ int s_base[2];
    __shared__ int s_w[kPpBlock / 32][2];
    const int b = blockIdx.y;
    const PpState st = state[b];
    const uint32_t T = st.thresh_key;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (warp == 0) {
        int g = 0, e = 0;
        for (int j = lane; j < (int)blockIdx.x; j += 32) {
            const int2 c = __ldg(blk_counts + b * gridDim.x + j);
            g += c.x; e += c.y;
        }
#pragma unroll
        for (int d = 16; d; d >>= 1) {
            g += __shfl_xor_sync(0xffffffffu, g, d);
            e += __shfl_xor_sync(0xffffffffu, e, d);
        }
        if (lane == 0) { s_base[0] = g; s_base[1] = e; }
    }
    __syncthreads();
    int base_gt = s_base[0], base_eq = s_base[1];
    const int count_gt = st.count - st.need_eq;
    unsigned long long *out = sel + (size_t)b * sel_stride;
    const int begin = blockIdx.x * span, end = min(n_anchors, begin + span);
    for (int i0 = begin; i0 < end; i0 += kPpBlock) {
        const int i = i0 + threadIdx.x;
        const uint32_t key = i < end ? __ldg(keys + (size_t)b * n_anchors + i) : 0u;
        const bool is_gt = key > T, is_eq = !st.all && key == T;
        const uint32_t m_gt = __ballot_sync(0xffffffffu, is_gt), m_eq = __ballot_sync(0xffffffffu, is_eq);
        if (lane == 0) { s_w[warp][0] = __popc(m_gt); s_w[warp][1] = __popc(m_eq); }
        __syncthreads();
        int off_gt = 0, off_eq = 0, tot_gt = 0, tot_eq = 0;
#pragma unroll
        for (int w = 0; w < kPpBlock / 32; ++w) {
            const int g = s_w[w][0], e = s_w[w][1];
            if (w < warp) { off_gt += g; off_eq += e; }
            tot_gt += g; tot_eq += e;
        }
        const uint32_t lt = (1u << lane) - 1u;
        const unsigned long long word = ((unsigned long long)key << 32) | (0xFFFFFFFFu - (uint32_t)i);
        if (is_gt) out[base_gt + off_gt + __popc(m_gt & lt)] = word;
        if (is_eq) {
            const int r = base_eq + off_eq + __popc(m_eq & lt);
            if (r < st.need_eq) out[count_gt + r] = word;
        }
        base_gt += tot_gt;
        base_eq += tot_eq;
        __syncthreads();
    }
}

struct DecodeArgs {
    const float *cls, *box, *dir, *anchors;
    int n_anchors, n_classes, cls_stride, num_dir_bins, binary_dir, pre_max;
    float dir_offset, dir_limit_offset, period;
    float *boxes3d, *boxes_bev, *scores;
    int *labels, *anchor_index, *count;
};

constexpr int kRankTile = 64;                  // candidates per block
constexpr int kRankParts = 4;                  // each candidate's comparisons are split over this many threads
constexpr int kRankThreads = kRankTile * kRankParts;

// grid: (ceil(pre_max / 64), batch).  Sorting by counting: the position of a selected candidate in (score desc,
// anchor asc) order is the number of selected candidates that precede it.  The compacted list is in anchor order (the
// candidates above the k-th key, then the ties at the k-th key, which are smaller than all of them), so candidate i is
// preceded by the j with key_j > key_i and by the j < i with key_j == key_i: 32-bit keys suffice -- ">=" for the list
// positions in front of the block's 64 candidates, ">" for those behind, the exact rule inside.  Every block holds the
// frame's keys in shared memory (LDS.128 broadcasts, four keys per load), ranks 64 candidates with their comparisons
// split four ways and decodes those 64 straight into their output rows: 256 independent blocks per SECOND batch instead
// of a four-CTA sorting network, and the gathers of the residuals / anchors walk the anchors in ascending order.
__global__ void __launch_bounds__(kRankThreads)
pp_rank_decode(const PpState *__restrict__ state, const unsigned long long *__restrict__ sel, const __grid_constant__ DecodeArgs a)
{
    extern __shared__ __align__(16) uint32_t s_key[];
    __shared__ int s_rank[kRankParts][kRankTile];
    const int b = blockIdx.y;
    const int count = state[b].count;
    const int e0 = blockIdx.x * kRankTile;
    if (blockIdx.x == 0 && threadIdx.x == 0) a.count[b] = count;
    // unused tail: zero-area boxes far apart from everything (they neither suppress nor get suppressed)
    for (int i = max(e0, count) + (int)threadIdx.x; i < min(e0 + kRankTile, a.pre_max); i += kRankThreads) {
        const size_t o = (size_t)b * a.pre_max + i;
        float *b3 = a.boxes3d + o * 7, *bev = a.boxes_bev + o * 5;
#pragma unroll
        for (int c = 0; c < 7; ++c) b3[c] = 0.f;
        const float px = 1.0e6f + 16.f * (float)i;
        bev[0] = px; bev[1] = 1.0e6f; bev[2] = px; bev[3] = 1.0e6f; bev[4] = 0.f;
        a.scores[o] = 0.f; a.labels[o] = 0; a.anchor_index[o] = -1;
    }
    if (e0 >= count) return;
    const unsigned long long *src = sel + (size_t)b * a.pre_max;
    const int padded = (count + 15) & ~15;                           // zero keys behind the list: never counted
    for (int i = threadIdx.x; i < padded; i += kRankThreads) s_key[i] = i < count ? (uint32_t)(__ldg(src + i) >> 32) : 0u;
    __syncthreads();
    const int e = threadIdx.x & (kRankTile - 1), part = threadIdx.x / kRankTile;
    const int p = e0 + e;
    const uint32_t mine = p < count ? s_key[p] : 0xFFFFFFFFu;
    const int per = padded / kRankParts;                             // a multiple of 4
    const int lo = part * per, hi = lo + per;
    int ahead = 0;
    for (int j = lo; j < min(hi, e0); j += 4) {                      // in front of the tile
        const uint4 q = *reinterpret_cast<const uint4 *>(s_key + j);
        ahead += (q.x >= mine) + (q.y >= mine) + (q.z >= mine) + (q.w >= mine);
    }
    for (int j = max(lo, e0 + kRankTile); j < hi; j += 4) {          // behind the tile
        const uint4 q = *reinterpret_cast<const uint4 *>(s_key + j);
        ahead += (q.x > mine) + (q.y > mine) + (q.z > mine) + (q.w > mine);
    }
    constexpr int kOwn = kRankTile / kRankParts;                     // the tile itself, a quarter per part
    for (int j = e0 + part * kOwn; j < min(e0 + (part + 1) * kOwn, padded); ++j) {
        const uint32_t q = s_key[j];
        ahead += q > mine || (q == mine && j < p);
    }
    s_rank[part][e] = ahead;
    __syncthreads();
    if (threadIdx.x >= kRankTile || p >= count) return;
    int rank = 0;
#pragma unroll
    for (int q = 0; q < kRankParts; ++q) rank += s_rank[q][e];
    const unsigned long long word = __ldg(src + p);
    const size_t o = (size_t)b * a.pre_max + rank;
    float *b3 = a.boxes3d + o * 7, *bev = a.boxes_bev + o * 5;
    const int idx = (int)(0xFFFFFFFFu - (uint32_t)word);
    const float *t = a.box + ((size_t)b * a.n_anchors + idx) * 7, *an = a.anchors + (size_t)idx * 7;
    float tv[7], av[7];
#pragma unroll
    for (int c = 0; c < 7; ++c) { tv[c] = __ldg(t + c); av[c] = __ldg(an + c); }
    int label;
    class_max(a.cls + ((size_t)b * a.n_anchors + idx) * a.cls_stride, a.n_classes, &label);
    const float xa = av[0], ya = av[1], wa = av[3], la = av[4], ha = av[5], ra = av[6];
    const float za = __fadd_rn(av[2], ha * 0.5f);                                               // box_coder_utils.py:99
    const float diagonal = sqrtf(__fadd_rn(__fmul_rn(la, la), __fmul_rn(wa, wa)));              // :101
    const float xg = __fadd_rn(__fmul_rn(tv[0], diagonal), xa);                                 // :102-104
    const float yg = __fadd_rn(__fmul_rn(tv[1], diagonal), ya);
    float zg = __fadd_rn(__fmul_rn(tv[2], ha), za);
    const float wg = __fmul_rn(expf(tv[3]), wa);                                                // :106-108 (wt, lt, ht)
    const float lg = __fmul_rn(expf(tv[4]), la);
    const float hg = __fmul_rn(expf(tv[5]), ha);
    float rg = __fadd_rn(tv[6], ra);                                                            // :109
    zg = __fsub_rn(zg, hg * 0.5f);                                                              // :111
    if (a.dir) {
        const float *d = a.dir + ((size_t)b * a.n_anchors + idx) * a.num_dir_bins;
        int dir_label;
        class_max(d, a.num_dir_bins, &dir_label);                                               // :127 / :135
        if (a.binary_dir) {
            if ((rg > 0.f) != (dir_label != 0)) rg = __fadd_rn(rg, 3.14159265358979323846f);    // :128-133
        } else {
            const float val = __fsub_rn(rg, a.dir_offset);                                      // :138-141, common_utils.py:95-96
            const float turns = floorf(__fadd_rn(__fdiv_rn(val, a.period), a.dir_limit_offset));
            const float dir_rot = __fsub_rn(val, __fmul_rn(turns, a.period));
            rg = __fadd_rn(__fadd_rn(dir_rot, a.dir_offset), __fmul_rn(a.period, (float)dir_label));
        }
    }
    b3[0] = xg; b3[1] = yg; b3[2] = zg; b3[3] = wg; b3[4] = lg; b3[5] = hg; b3[6] = rg;
    const float hw = wg * 0.5f, hl = lg * 0.5f;                                                 // box_utils.py:244-249
    bev[0] = __fsub_rn(xg, hw); bev[1] = __fsub_rn(yg, hl); bev[2] = __fadd_rn(xg, hw); bev[3] = __fadd_rn(yg, hl); bev[4] = rg;
    a.scores[o] = key_score(mine);
    a.labels[o] = label + 1;                                                                    // detector3d.py:197
    a.anchor_index[o] = idx;
}

// grid: (batch).  The tail of class_agnostic_nms / post_processing (detector3d.py:290-299, 211-219): kept positions ->
// boxes, scores, labels, anchor indices of the kept detections; positions >= count[b] are the padding rows.
__global__ void __launch_bounds__(256)
pp_gather_kept(const long long *__restrict__ keep, int keep_stride, const int *__restrict__ count, int pre_max,
               const float *__restrict__ boxes3d, const float *__restrict__ scores, const int *__restrict__ labels,
               const int *__restrict__ anchor_index, int post_max, int sigmoid_scores, float pad_score, int pad_label,
               float *__restrict__ out_boxes,
               float *__restrict__ out_scores, long long *__restrict__ out_labels, long long *__restrict__ out_selected,
               int *__restrict__ out_num)
{
    __shared__ int s_num;
    const int b = blockIdx.x;
    const int n = __ldg(count + b);
    if (threadIdx.x == 0) s_num = 0;
    __syncthreads();
    int mine = 0;
    for (int i = threadIdx.x; i < post_max; i += blockDim.x) {
        const long long pos = i < keep_stride ? keep[(size_t)b * keep_stride + i] : -1;
        const bool valid = pos >= 0 && pos < n;
        const size_t o = (size_t)b * post_max + i, src = (size_t)b * pre_max + (valid ? pos : 0);
#pragma unroll
        for (int c = 0; c < 7; ++c) out_boxes[o * 7 + c] = valid ? boxes3d[src * 7 + c] : 0.f;
        const float sc = valid ? scores[src] : 0.f;
        out_scores[o] = valid ? (sigmoid_scores ? 1.f / (1.f + expf(-sc)) : sc) : pad_score;
        out_labels[o] = valid ? labels[src] : pad_label;
        out_selected[o] = valid ? anchor_index[src] : -1;
        mine += valid;
    }
    if (mine) atomicAdd(&s_num, mine);
    __syncthreads();
    if (threadIdx.x == 0) out_num[b] = s_num;
}

}  // namespace pcdb

using namespace pcdb;

extern "C" size_t pcdb_decode_select_workspace_bytes(int batch, int n_anchors, int pre_max)
{
    if (batch < 1 || n_anchors < 1 || pre_max < 1) return 0;
    return carve_pp(nullptr, batch, n_anchors, pre_max).bytes;
}

extern "C" int pcdb_decode_select(const float *cls_preds, int cls_stride, const float *box_preds, const float *dir_cls_preds,
                                  const float *anchors, int batch, int n_anchors, int n_classes, int num_dir_bins,
                                  float dir_offset, float dir_limit_offset, float score_thresh, int pre_max, int flags,
                                  float *boxes3d, float *boxes_bev, float *scores, int32_t *labels, int32_t *anchor_index,
                                  int32_t *count, void *workspace, size_t workspace_bytes, void *stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    if (!cls_preds || !box_preds || !anchors || batch < 1 || batch > 65535 || n_anchors < 1 || n_classes < 1 ||
        cls_stride < n_classes || pre_max < 1 || !boxes3d || !boxes_bev || !scores || !labels || !anchor_index || !count ||
        (dir_cls_preds && num_dir_bins < 1)) {
        set_last_error("pcdb_decode_select: invalid argument (batch=%d n_anchors=%d n_classes=%d pre_max=%d)", batch, n_anchors,
                       n_classes, pre_max);
        return kInvalidArgument;
    }
    if (pre_max > 16384) { set_last_error("pcdb_decode_select: pre_max %d > 16384 (shared-memory sort)", pre_max); return kUnsupported; }
    const PpWorkspace w = carve_pp(workspace, batch, n_anchors, pre_max);
    if (!workspace || workspace_bytes < w.bytes) {
        set_last_error("pcdb_decode_select: workspace %zu < required %zu bytes", workspace_bytes, w.bytes);
        return kWorkspaceTooSmall;
    }
    const int span0 = score_span(n_anchors, cls_stride), span1 = refine_span(n_anchors), span2 = compact_span(n_anchors);
    const dim3 grid0(blocks_for(n_anchors, span0), batch), grid1(blocks_for(n_anchors, span1), batch), grid2(blocks_for(n_anchors, span2), batch);
    cudaMemsetAsync(w.hist, 0, w.zero_bytes, stream);
    // staging needs span * cls_stride floats of shared memory; span_for may have widened the span for very long frames
    const bool staged = score_staged(cls_stride) && (size_t)span0 * cls_stride <= kStageFloats;
    auto score_hist = staged ? (n_classes == 1 ? pp_score_hist<1, true> : n_classes == 2 ? pp_score_hist<2, true>
                                : n_classes == 3 ? pp_score_hist<3, true> : n_classes == 4 ? pp_score_hist<4, true> : pp_score_hist<0, true>)
                             : (n_classes == 1 ? pp_score_hist<1, false> : n_classes == 2 ? pp_score_hist<2, false>
                                : n_classes == 3 ? pp_score_hist<3, false> : n_classes == 4 ? pp_score_hist<4, false> : pp_score_hist<0, false>);
    score_hist<<<grid0, kPpBlock, staged ? sizeof(float) * (size_t)span0 * cls_stride : 0, stream>>>(
        cls_preds, n_anchors, n_classes, cls_stride, score_thresh, span0, pre_max, w.keys, w.hist, w.done, w.state);
    pp_refine<1><<<grid1, kPpBlock, 0, stream>>>(w.keys, n_anchors, span1, w.hist, w.done, w.state);
    pp_refine<2><<<grid1, kPpBlock, 0, stream>>>(w.keys, n_anchors, span1, w.hist, w.done, w.state);
    pp_count<<<grid2, kPpBlock, 0, stream>>>(w.keys, n_anchors, span2, w.state, w.blk_counts);
    pp_scatter<<<grid2, kPpBlock, 0, stream>>>(w.keys, n_anchors, span2, w.state, w.blk_counts, w.sel, pre_max);
    DecodeArgs a;
    a.cls = cls_preds; a.box = box_preds; a.dir = dir_cls_preds; a.anchors = anchors;
    a.n_anchors = n_anchors; a.n_classes = n_classes; a.cls_stride = cls_stride; a.num_dir_bins = num_dir_bins;
    a.binary_dir = (flags & PCDB_DIR_BINARY) ? 1 : 0; a.pre_max = pre_max;
    a.dir_offset = dir_offset; a.dir_limit_offset = dir_limit_offset;
    a.period = num_dir_bins > 0 ? (float)(2.0 * 3.14159265358979323846 / (double)num_dir_bins) : 0.f;
    a.boxes3d = boxes3d; a.boxes_bev = boxes_bev; a.scores = scores; a.labels = labels; a.anchor_index = anchor_index; a.count = count;
    const size_t smem = sizeof(uint32_t) * (size_t)(pre_max + 16);
    // per device and context, cheap: set on every call (a process-wide "done once" flag breaks on a second GPU)
    if (cudaFuncSetAttribute(pp_rank_decode, cudaFuncAttributeMaxDynamicSharedMemorySize, (16384 + 16) * 4) != cudaSuccess)
        return check_launch("pcdb_decode_select(cudaFuncSetAttribute)");
    pp_rank_decode<<<dim3((pre_max + kRankTile - 1) / kRankTile, batch), kRankThreads, smem, stream>>>(w.state, w.sel, a);
    return check_launch("pcdb_decode_select");
}

extern "C" int pcdb_gather_kept(const int64_t *keep, int keep_stride, const int32_t *count, int batch, int pre_max,
                                const float *boxes3d, const float *scores, const int32_t *labels, const int32_t *anchor_index,
                                int post_max, int sigmoid_scores, float pad_score, int pad_label, float *out_boxes,
                                float *out_scores, int64_t *out_labels, int64_t *out_selected, int32_t *out_num, void *stream_)
{
    if (!keep || !count || !boxes3d || !scores || !labels || !anchor_index || !out_boxes || !out_scores || !out_labels ||
        !out_selected || !out_num || batch < 1 || pre_max < 1 || post_max < 1 || keep_stride < 1) {
        set_last_error("pcdb_gather_kept: invalid argument (batch=%d pre_max=%d post_max=%d)", batch, pre_max, post_max);
        return kInvalidArgument;
    }
    pp_gather_kept<<<batch, 256, 0, (cudaStream_t)stream_>>>((const long long *)keep, keep_stride, count, pre_max, boxes3d, scores, labels,
                                                             anchor_index, post_max, sigmoid_scores, pad_score, pad_label, out_boxes, out_scores,
                                                             (long long *)out_labels, (long long *)out_selected, out_num);
    return check_launch("pcdb_gather_kept");
}
