// Rulebook (indice pair) construction for sparse 3-D convolutions on sm_100a.
//
// Replaces spconv v1.0 getIndicePair<3>: the reference fills a dense int32 grid of
// batch*Z*Y*X cells (370 MB per KITTI sample at level 1) for every build, appends pairs with atomics
// and sorts the touched cells with torch::_unique.  Here the active sites live in an open-addressing
// hash table in HBM (8 B per slot, load factor <= 0.5, a few hundred KB), and the rulebook is emitted
// output-stationary: nbr[k*ld + o] = input row feeding output row o through kernel offset k, or -1.
// That is the layout the convolution kernel consumes (one coalesced index load per tile and offset,
// no scatter, no atomics, deterministic summation order).
//
// Output rows of a strided convolution are numbered in the first-touch order of the reference's
// serial CPU loop (input row ascending, kernel offset ascending), recovered in parallel exactly like
// voxel ids: the table keeps the minimum (row*K + k) per output site, the holder of that minimum is
// the site's owner, and an exclusive scan of owner counts over input rows gives the id.
#include "rulebook.cuh"
#include "../../include/pcdet_b200.h"

namespace pcdb {

constexpr int kRbScanBlock = 256;

// ---- submanifold -----------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
rb_insert_rows(const int4 *__restrict__ indices, int n, const int *__restrict__ n_dev, ConvGeom g,
               unsigned long long *slots, uint32_t mask)
{
    n = row_count(n, n_dev);
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    const int4 c = __ldg(indices + r);
    table_insert_min(slots, mask, lin_index(c.x, c.y, c.z, c.w, g.in_shape), (uint32_t)r);
}

// grid: (ceil(n/256), SYMMETRIC ? K/2 + 1 : K) -- one thread per (site, offset): the table probes are random
// 8-byte reads and only massive thread parallelism hides their latency (a thread-per-site loop over the offsets
// was measured 2x slower).  Site o receives from in = o - pad + k*dil (stride 1, pad = k/2 forced).
// SYMMETRIC (centred offsets: odd kernel sizes, dilation 1): offset K-1-k is the negation of offset k, so nbr[k][o] = i implies
// nbr[K-1-k][i] = o -- only the first half of the offsets is probed (13 instead of 27 random table reads per
// site), hits are written in both directions, the centre (blockIdx.y == K/2) is the site itself, and the
// caller pre-fills nbr with -1.  Otherwise every offset is probed and written (including the misses).
template <bool SYMMETRIC>
__global__ void __launch_bounds__(256)
rb_subm_neighbours(const int4 *__restrict__ indices, int n, const int *__restrict__ n_dev, ConvGeom g,
                   const unsigned long long *__restrict__ slots, uint32_t mask, const int *__restrict__ slot_oid,
                   int *__restrict__ nbr, int ld)
{
    n = row_count(n, n_dev);
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    const int k = blockIdx.y;
    if (SYMMETRIC && k == (g.K >> 1)) { nbr[(size_t)k * ld + r] = r; return; }
    const int4 c = __ldg(indices + r);
    const int hit = site_row(g, slots, mask, slot_oid, n, c.x, c.y - g.pad[0] + g.dk[k][0], c.z - g.pad[1] + g.dk[k][1],
                             c.w - g.pad[2] + g.dk[k][2]);
    if (!SYMMETRIC) {
        nbr[(size_t)k * ld + r] = hit;
    } else if (hit >= 0) {
        nbr[(size_t)k * ld + r] = hit;
        nbr[(size_t)(g.K - 1 - k) * ld + hit] = r;
    }
}

// grid: (ceil(n/256), combos).  Keeps the smallest row*K+k per output site and records slot and offset of every
// (row, candidate) -- pair_slot[cand*ld + row] = slot << 5 | k, or -1 -- so that the later passes read them back
// coalesced instead of probing the table again.
__global__ void __launch_bounds__(256)
rb_conv_insert(const int4 *__restrict__ indices, int n, const int *__restrict__ n_dev, ConvGeom g,
               unsigned long long *slots, uint32_t mask, int *__restrict__ pair_slot, int ld_in)
{
    n = row_count(n, n_dev);
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    const int cand = blockIdx.y;
    const int4 c = __ldg(indices + r);
    int k, oz, oy, ox;
    int v = -1;
    if (conv_candidate(g, c, cand, &k, &oz, &oy, &ox))
        v = (int)(table_insert_min(slots, mask, lin_index(c.x, oz, oy, ox, g.out_shape),
                                   (uint32_t)r * (uint32_t)g.K + (uint32_t)k) << 5) | k;
    pair_slot[(size_t)cand * ld_in + r] = v;
}

// One thread per TABLE SLOT: the payload row*K+k that survived in a slot names the first toucher of that output
// site, so bit k of own_mask[row] is set from the slot itself (table_cap threads instead of n*K candidate
// pairs).  own_mask is zeroed by the caller.
__global__ void __launch_bounds__(256)
rb_conv_mark(const unsigned long long *__restrict__ slots, uint32_t table_cap, int K, uint32_t *__restrict__ own_mask)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= table_cap) return;
    const unsigned long long w = __ldg(slots + i);
    if (w == kEmptySlot) return;
    const uint32_t payload = (uint32_t)w, r = payload / (uint32_t)K;
    atomicOr(own_mask + r, 1u << (payload - r * (uint32_t)K));
}

// Single-pass numbering of the output sites (decoupled look-back scan over blocks of input rows).
// Row r OWNS the sites of which it is the first toucher (own_mask, from rb_conv_mark); the exclusive
// prefix of the owner counts over rows -- then over the set bits of a row -- is the reference's first-touch
// output id.  Block ids come from a ticket, so a block's predecessors are always already running.
// scan_state[b]: flag (high word: 0xFFFFFFFF empty / 1 block total / 2 inclusive prefix) | value (low word);
// the ticket and the states live in the region the caller fills with 0xFF.
__global__ void __launch_bounds__(kRbScanBlock)
rb_conv_number(const int4 *__restrict__ indices, int n, const int *__restrict__ n_dev, ConvGeom g,
               const int *__restrict__ pair_slot, int ld_in, const uint32_t *__restrict__ own_mask,
               unsigned long long *scan_state, unsigned int *ticket, int *__restrict__ slot_oid,
               int4 *__restrict__ out_indices, int n_out_cap, int *n_out_dev)
{
    __shared__ int s_bid, s_prefix;
    n = row_count(n, n_dev);
    if (threadIdx.x == 0) s_bid = (int)(atomicAdd(ticket, 1u) + 1u);
    __syncthreads();
    const int bid = s_bid;
    const int r = bid * kRbScanBlock + threadIdx.x;
    uint32_t own = r < n ? __ldg(own_mask + r) : 0u;
    int total;
    int oid = block_exclusive_scan<kRbScanBlock>(__popc(own), &total);
    if (threadIdx.x < 32) {
        // warp-wide look-back: 32 predecessors per step, stop at the nearest one that already has its prefix
        volatile unsigned long long *st = scan_state;
        const int lane = threadIdx.x;
        if (lane == 0 && bid > 0) st[bid] = (1ull << 32) | (uint32_t)total;
        int prefix = 0;
        for (int base = bid - 1; base >= 0; base -= 32) {
            const int p = base - lane;
            unsigned long long w = 2ull << 32;                      // before block 0: an (empty) prefix
            if (p >= 0) do { w = st[p]; } while ((uint32_t)(w >> 32) == 0xFFFFFFFFu);
            const uint32_t has_prefix = __ballot_sync(0xffffffffu, (uint32_t)(w >> 32) == 2u);
            const int upto = has_prefix ? __ffs(has_prefix) - 1 : 31;     // lanes 0..upto contribute
            int v = lane <= upto ? (int)(uint32_t)w : 0;
#pragma unroll
            for (int d = 16; d; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
            prefix += v;
            if (has_prefix) break;
        }
        if (lane == 0) {
            st[bid] = (2ull << 32) | (uint32_t)(prefix + total);
            s_prefix = prefix;
            if (bid == (int)gridDim.x - 1) {
                const int all = prefix + total;
                n_out_dev[0] = all < n_out_cap ? all : n_out_cap;
                n_out_dev[1] = all > n_out_cap ? 1 : 0;
            }
        }
    }
    __syncthreads();
    oid += s_prefix;
    if (own) {
        const int4 c = __ldg(indices + r);
        for (; own; own &= own - 1, ++oid) {
            const int k = __ffs(own) - 1;
            int oz, oy, ox;
            out_site(g, c, k, &oz, &oy, &ox);
            slot_oid[__ldg(pair_slot + (size_t)candidate_of(g, c, k) * ld_in + r) >> 5] = oid;
            if (oid < n_out_cap) out_indices[oid] = make_int4(c.x, oz, oy, ox);
        }
    }
}

// grid: (ceil(n/256), combos).  nbr_inv, when wanted, was filled with -1 by the caller.
__global__ void __launch_bounds__(256)
rb_conv_fill(int n, const int *__restrict__ n_dev, const int *__restrict__ pair_slot, int ld_ps,
             const int *__restrict__ slot_oid, int n_out_cap, int *__restrict__ nbr_fwd, int ld_out,
             int *__restrict__ nbr_inv, int ld_in)
{
    n = row_count(n, n_dev);
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    const int v = __ldg(pair_slot + (size_t)blockIdx.y * ld_ps + r);
    if (v < 0) return;
    const int k = v & 31, oid = __ldg(slot_oid + (v >> 5));
    if (oid >= n_out_cap) return;
    nbr_fwd[(size_t)k * ld_out + oid] = r;
    if (nbr_inv) nbr_inv[(size_t)k * ld_in + r] = oid;
}

// Workspace of one build.  [slots | ticket | scan_state] is filled with 0xFF before every build; slots and
// slot_oid stay valid afterwards: pcdb_rulebook_subm_reuse reads them as the site table of the output level.
struct RbWorkspace {
    unsigned long long *slots;
    unsigned int *ticket;
    unsigned long long *scan_state;
    int *slot_oid, *pair_slot;
    uint32_t *own_mask;
    uint32_t table_cap;
    int nblocks;
    size_t fill_bytes, bytes;
};

static RbWorkspace carve_rb(void *base, int n_in_cap, int n_sites_cap, int K = 0)
{
    RbWorkspace w{};
    const size_t sites = (size_t)(n_sites_cap > 0 ? n_sites_cap : 1);
    w.table_cap = next_pow2(sites * 2 < 1024 ? 1024 : sites * 2);
    w.nblocks = (int)(((size_t)(n_in_cap > 0 ? n_in_cap : 1) + kRbScanBlock - 1) / kRbScanBlock);
    size_t off = 0;
    char *b = (char *)base;
    auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes, 256); return b ? (void *)(b + o) : (void *)nullptr; };
    w.slots = (unsigned long long *)take((size_t)w.table_cap * 8);
    w.ticket = (unsigned int *)take(4);
    w.scan_state = (unsigned long long *)take((size_t)w.nblocks * 8);
    w.fill_bytes = off;
    w.slot_oid = (int *)take((size_t)w.table_cap * 4);
    w.pair_slot = (int *)take((size_t)K * (size_t)(n_in_cap > 0 ? n_in_cap : 1) * 4);
    w.own_mask = (uint32_t *)take((size_t)(n_in_cap > 0 ? n_in_cap : 1) * 4);
    w.bytes = off;
    return w;
}

}  // namespace pcdb

using namespace pcdb;

extern "C" size_t pcdb_rulebook_workspace_bytes(int n_in_cap, int kernel_volume, int n_out_cap)
{
    const int sites = n_out_cap > n_in_cap ? n_out_cap : n_in_cap;
    return carve_rb(nullptr, n_in_cap, sites, kernel_volume).bytes;
}

static void launch_subm_neighbours(const int32_t *indices, int n, const int32_t *n_dev, const ConvGeom &g,
                                   const unsigned long long *slots, uint32_t table_cap, const int *slot_oid,
                                   int32_t *nbr, int ld, cudaStream_t stream, bool nbr_cleared = false)
{
    const int nb = (n + 255) / 256;
    if (symmetric_offsets(g)) {
        if (!nbr_cleared) cudaMemsetAsync(nbr, 0xFF, sizeof(int32_t) * (size_t)g.K * ld, stream);
        rb_subm_neighbours<true><<<dim3(nb, g.K / 2 + 1), 256, 0, stream>>>((const int4 *)indices, n, n_dev, g, slots, table_cap - 1, slot_oid, nbr, ld);
    } else {
        rb_subm_neighbours<false><<<dim3(nb, g.K), 256, 0, stream>>>((const int4 *)indices, n, n_dev, g, slots, table_cap - 1, slot_oid, nbr, ld);
    }
}

static int check_subm_args(const char *who, ConvGeom &g, const int32_t *indices, int n, int batch, const int32_t *shape,
                           const int32_t *ksize, const int32_t *dil, const int32_t *nbr, int ld)
{
    if (n < 0 || batch < 1 || !indices || !nbr || ld < n || !fill_geom(g, shape, nullptr, ksize, nullptr, nullptr, dil)) {
        set_last_error("%s: invalid argument (n=%d batch=%d ld=%d)", who, n, batch, ld);
        return kInvalidArgument;
    }
    // padding = k/2 is forced by fill_geom (pad == nullptr), as spconv does for SubM
    const uint64_t cells = (uint64_t)batch * g.in_shape[0] * (uint64_t)g.in_shape[1] * g.in_shape[2];
    if (cells >= 0xFFFFFFFFull) {
        set_last_error("%s: batch*volume = %llu exceeds the 32-bit hash key", who, (unsigned long long)cells);
        return kKeyOverflow;
    }
    return kOk;
}

extern "C" int pcdb_rulebook_subm(const int32_t *indices, int n, const int32_t *n_dev, int batch,
                                  const int32_t *spatial_shape_zyx, const int32_t *ksize_zyx,
                                  const int32_t *dilation_zyx, int32_t *nbr, int ld,
                                  void *workspace, size_t workspace_bytes, void *stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    ConvGeom g;
    const int st = check_subm_args("pcdb_rulebook_subm", g, indices, n, batch, spatial_shape_zyx, ksize_zyx, dilation_zyx, nbr, ld);
    if (st != kOk) return st;
    if (n == 0) return kOk;
    RbWorkspace w = carve_rb(workspace, n, n);
    if (!workspace || workspace_bytes < w.bytes) {
        set_last_error("pcdb_rulebook_subm: workspace %zu < required %zu bytes", workspace_bytes, w.bytes);
        return kWorkspaceTooSmall;
    }
    cudaMemsetAsync(w.slots, 0xFF, (size_t)w.table_cap * 8, stream);
    rb_insert_rows<<<(n + 255) / 256, 256, 0, stream>>>((const int4 *)indices, n, n_dev, g, w.slots, w.table_cap - 1);
    launch_subm_neighbours(indices, n, n_dev, g, w.slots, w.table_cap, nullptr, nbr, ld, stream);
    return check_launch("pcdb_rulebook_subm");
}

extern "C" int pcdb_rulebook_subm_reuse(const int32_t *indices, int n, const int32_t *n_dev, int batch,
                                        const int32_t *spatial_shape_zyx, const int32_t *ksize_zyx,
                                        const int32_t *dilation_zyx, int32_t *nbr, int ld,
                                        const void *conv_workspace, int conv_n_in_cap, int conv_kernel_volume,
                                        int conv_n_out_cap, int flags, void *stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    ConvGeom g;
    const int st = check_subm_args("pcdb_rulebook_subm_reuse", g, indices, n, batch, spatial_shape_zyx, ksize_zyx, dilation_zyx, nbr, ld);
    if (st != kOk) return st;
    if (!conv_workspace || conv_n_in_cap < 1 || conv_n_out_cap < 1 || n > conv_n_out_cap) {
        set_last_error("pcdb_rulebook_subm_reuse: invalid site table (n=%d conv_n_out_cap=%d)", n, conv_n_out_cap);
        return kInvalidArgument;
    }
    if (n == 0) return kOk;
    const RbWorkspace w = carve_rb(const_cast<void *>(conv_workspace), conv_n_in_cap, conv_n_out_cap, conv_kernel_volume);
    launch_subm_neighbours(indices, n, n_dev, g, w.slots, w.table_cap, w.slot_oid, nbr, ld, stream,
                           (flags & PCDB_RB_CLEARED) != 0);
    return check_launch("pcdb_rulebook_subm_reuse");
}

extern "C" int pcdb_rulebook_conv_sites(const int32_t *indices, int n, const int32_t *n_dev, int batch,
                                        const int32_t *spatial_shape_zyx, const int32_t *out_shape_zyx,
                                        const int32_t *ksize_zyx, const int32_t *stride_zyx, const int32_t *padding_zyx,
                                        const int32_t *dilation_zyx, int32_t *out_indices, int n_out_cap,
                                        int32_t *n_out_dev, void *workspace, size_t workspace_bytes, int flags,
                                        void *stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    ConvGeom g;
    if (n < 0 || batch < 1 || !indices || !out_indices || !n_out_dev || n_out_cap < 1 ||
        !fill_geom(g, spatial_shape_zyx, out_shape_zyx, ksize_zyx, stride_zyx, padding_zyx, dilation_zyx)) {
        set_last_error("pcdb_rulebook_conv: invalid argument (n=%d batch=%d n_out_cap=%d)", n, batch, n_out_cap);
        return kInvalidArgument;
    }
    const uint64_t cells = (uint64_t)batch * g.out_shape[0] * (uint64_t)g.out_shape[1] * g.out_shape[2];
    if (cells >= 0xFFFFFFFFull || (uint64_t)n * g.K >= 0xFFFFFFFFull) {
        set_last_error("pcdb_rulebook_conv: key/payload exceed 32 bits (cells=%llu n*K=%llu)",
                       (unsigned long long)cells, (unsigned long long)n * g.K);
        return kKeyOverflow;
    }
    if (n == 0) {
        cudaMemsetAsync(n_out_dev, 0, 8, stream);
        return check_launch("pcdb_rulebook_conv(memset)");
    }
    RbWorkspace w = carve_rb(workspace, n, n_out_cap, g.K);
    if (!workspace || workspace_bytes < w.bytes) {
        set_last_error("pcdb_rulebook_conv: workspace %zu < required %zu bytes", workspace_bytes, w.bytes);
        return kWorkspaceTooSmall;
    }
    if (w.table_cap > (1u << 26)) {       // slot << 5 | k must fit an int32
        set_last_error("pcdb_rulebook_conv: n_out_cap %d needs a hash table beyond 2^26 slots", n_out_cap);
        return kKeyOverflow;
    }
    if (!(flags & PCDB_RB_CLEARED)) {
        cudaMemsetAsync(w.slots, 0xFF, w.fill_bytes, stream);
        cudaMemsetAsync(w.own_mask, 0, sizeof(uint32_t) * (size_t)n, stream);
    }
    const int nb = (n + 255) / 256;
    rb_conv_insert<<<dim3(nb, g.combos), 256, 0, stream>>>((const int4 *)indices, n, n_dev, g, w.slots, w.table_cap - 1, w.pair_slot, n);
    rb_conv_mark<<<(w.table_cap + 255) / 256, 256, 0, stream>>>(w.slots, w.table_cap, g.K, w.own_mask);
    rb_conv_number<<<w.nblocks, kRbScanBlock, 0, stream>>>((const int4 *)indices, n, n_dev, g, w.pair_slot, n, w.own_mask,
                                                           w.scan_state, w.ticket, w.slot_oid, (int4 *)out_indices,
                                                           n_out_cap, n_out_dev);
    return check_launch("pcdb_rulebook_conv_sites");
}

extern "C" int pcdb_rulebook_conv_clear(void *workspace, size_t workspace_bytes, int n_in_cap, int kernel_volume,
                                        int n_out_cap, int32_t *nbr_fwd, int ld_out, void *stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    if (!workspace || n_in_cap < 1 || n_out_cap < 1 || kernel_volume < 1 || kernel_volume > 32 || (nbr_fwd && ld_out < n_out_cap)) {
        set_last_error("pcdb_rulebook_conv_clear: invalid argument");
        return kInvalidArgument;
    }
    RbWorkspace w = carve_rb(workspace, n_in_cap, n_out_cap, kernel_volume);
    if (workspace_bytes < w.bytes) {
        set_last_error("pcdb_rulebook_conv_clear: workspace %zu < required %zu bytes", workspace_bytes, w.bytes);
        return kWorkspaceTooSmall;
    }
    cudaMemsetAsync(w.slots, 0xFF, w.fill_bytes, stream);
    cudaMemsetAsync(w.own_mask, 0, sizeof(uint32_t) * (size_t)n_in_cap, stream);
    if (nbr_fwd) cudaMemsetAsync(nbr_fwd, 0xFF, sizeof(int32_t) * (size_t)kernel_volume * ld_out, stream);
    return check_launch("pcdb_rulebook_conv_clear");
}

extern "C" int pcdb_rulebook_conv_pairs(int n, const int32_t *n_dev, const int32_t *ksize_zyx, const int32_t *stride_zyx,
                                        const int32_t *dilation_zyx, int n_out_cap, int32_t *nbr_fwd, int ld_out,
                                        int32_t *nbr_inv, int ld_in, const void *workspace, int flags, void *stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    ConvGeom g;
    const int32_t one[3] = {1, 1, 1};
    if (n < 0 || !ksize_zyx || !nbr_fwd || n_out_cap < 1 || ld_out < n_out_cap || (nbr_inv && ld_in < n) || (n > 0 && !workspace) ||
        !fill_geom(g, one, one, ksize_zyx, stride_zyx, nullptr, dilation_zyx)) {      // only K and the candidate count matter
        set_last_error("pcdb_rulebook_conv_pairs: invalid argument (n=%d n_out_cap=%d)", n, n_out_cap);
        return kInvalidArgument;
    }
    if (n == 0) return kOk;
    const RbWorkspace w = carve_rb(const_cast<void *>(workspace), n, n_out_cap, g.K);
    if (!(flags & PCDB_RB_CLEARED)) cudaMemsetAsync(nbr_fwd, 0xFF, sizeof(int32_t) * (size_t)g.K * ld_out, stream);
    if (nbr_inv) cudaMemsetAsync(nbr_inv, 0xFF, sizeof(int32_t) * (size_t)g.K * ld_in, stream);
    rb_conv_fill<<<dim3((n + 255) / 256, g.combos), 256, 0, stream>>>(n, n_dev, w.pair_slot, n, w.slot_oid, n_out_cap,
                                                                    nbr_fwd, ld_out, nbr_inv, ld_in);
    return check_launch("pcdb_rulebook_conv_pairs");
}

extern "C" int pcdb_rulebook_conv(const int32_t *indices, int n, const int32_t *n_dev, int batch,
                                  const int32_t *spatial_shape_zyx, const int32_t *out_shape_zyx,
                                  const int32_t *ksize_zyx, const int32_t *stride_zyx, const int32_t *padding_zyx,
                                  const int32_t *dilation_zyx, int32_t *out_indices, int n_out_cap,
                                  int32_t *n_out_dev, int32_t *nbr_fwd, int ld_out, int32_t *nbr_inv, int ld_in,
                                  void *workspace, size_t workspace_bytes, void *stream_)
{
    if (!nbr_fwd || ld_out < n_out_cap || (nbr_inv && ld_in < n)) {
        set_last_error("pcdb_rulebook_conv: invalid argument (n=%d ld_out=%d n_out_cap=%d)", n, ld_out, n_out_cap);
        return kInvalidArgument;
    }
    const int st = pcdb_rulebook_conv_sites(indices, n, n_dev, batch, spatial_shape_zyx, out_shape_zyx, ksize_zyx, stride_zyx,
                                            padding_zyx, dilation_zyx, out_indices, n_out_cap, n_out_dev, workspace,
                                            workspace_bytes, 0, stream_);
    if (st != kOk || n == 0) return st;
    return pcdb_rulebook_conv_pairs(n, n_dev, ksize_zyx, stride_zyx, dilation_zyx, n_out_cap, nbr_fwd, ld_out, nbr_inv, ld_in,
                                    workspace, 0, stream_);
}
