// Every rulebook of a strided sparse-conv backbone in five launches (sm_100a).
//
// pcdb_rulebook_conv_sites / _pairs / pcdb_rulebook_subm_reuse (rulebook.cu) reproduce the row ORDER of the reference's
// serial CPU loop (spconv v1.0 getIndicePair<3>), which forces a chain: the sites of level l+1 are
// numbered by first touch in level-l row order, so each of BackBone8x's four strided builds waits for the previous
// one -- 24 launches of 3-12 us, as long as the twelve convolutions they feed (round 1: 172 us of a 373 us step).
// The reference's own CUDA path numbers the same sites differently again: ascending linear index (torch::_unique over
// the touched cells of its dense grid).  The order is an implementation detail, the SETS are the contract, and the
// ascending-index order needs no chain:
//
//   1. rbc_insert   The active set of level l is the union over the level-0 voxels of the sites each voxel reaches
//                   through convs 1..l, and what one voxel reaches is a BOX (an interval per dimension, propagated
//                   conv by conv).  One launch marks every level's boxes, straight from the level-0 coordinates, in a
//                   per-level OCCUPANCY BITMAP (one 64-bit word per 32 cells: the reference's dense grid, 370 MB per
//                   sample at level 1, shrinks to 12 MB for a batch of 4 at level 2 and to KBs above), and inserts the
//                   level-0 voxels into a hash table (that grid is too large: 369 M cells).
//   2. rbc_maps     level 0's SubM map (it needs the level-0 table only, so the first convolutions can start)
//   3. rbc_count    set bits per 1024-word block, exclusive prefix by the last block
//      rbc_assign   prefix of the set bits over the words of each level -> the high half of every word; a cell's row id
//                   is prefix + popcount(bits below it), i.e. rows in ascending (b, z, y, x) order -- the order of the
//                   reference's CUDA rulebook; coordinates decoded from the cell index; counts and overflow flags
//   4. rbc_maps     all other neighbour maps at once.  A lookup is ONE 8-byte read, and the cells x-1, x, x+1 of a
//                   SubM row of offsets share it.
//
// Pair sets per offset are identical to the reference's (tests/test_gpu_rulebook_chain.py compares them through the
// coordinates).
#include "rulebook.cuh"
#include "../../include/pcdet_b200.h"

#ifndef PCDB_RBC_CHECK
#define PCDB_RBC_CHECK 1
#endif

namespace pcdb {

constexpr int kChainMaxLevels = 6;        // level 0 + five strided convolutions
constexpr int kChainScanBlock = 1024;

struct ChainLevel {
    ConvGeom conv;                  // l >= 1: the strided conv (l-1) -> l (in_shape = level l-1, out_shape = level l)
    ConvGeom subm;                  // SubM geometry of this level; K = 0: no SubM map wanted
    unsigned long long *slots;      // level 0: hash table, key = linear cell (high half), payload = row (low half)
    uint32_t mask;                  // level >= 1: occupancy words, bits of cells 32w .. 32w+31 (low half), rows before (high)
    uint32_t n_words;               // level >= 1
    int4 *coords;                   // (cap, 4) rows [b, z, y, x]
    const int *count;               // device row count (level 0: the caller's; l >= 1: count_out)
    int *count_out;                 // l >= 1: [count, overflow]
    int cap;
    int *nbr_conv;                  // l >= 1: (conv.K, cap) map of the strided conv, rows = this level's
    int *nbr_subm;                  // (subm.K, cap)
    int *block_sums;                // l >= 1: set bits per 1024-word block (rbc_count)
    int *block_prefix;              // l >= 1: their exclusive prefix (+ total at [nblocks]), by rbc_count's last block
    int nblocks;                    // 1024-word blocks
    int scan_block0;                // first block of this level in rbc_count (kChunksPerBlock chunks per block)
    int assign_block0;              // first block of this level in rbc_assign (one chunk per block)
};

struct ChainMap { int kind, level, blocks_x, rows_y, block0; };     // kind 0: strided conv into `level`, 1: SubM at `level`

struct ChainParams {
    int n_levels, batch, n_maps;
    int *overflow;                  // set when the level-0 table filled up
    unsigned int *ticket;           // rbc_count: blocks that have finished
    ChainLevel lv[kChainMaxLevels];
    ChainMap maps[2 * kChainMaxLevels];
};

__device__ __forceinline__ int floor_div(int a, int b) { const int q = a / b; return (a % b != 0 && (a < 0) != (b < 0)) ? q - 1 : q; }
// floor(a / stride[d]) and ceil(a / stride[d]) for any sign of a: an arithmetic shift for the power-of-two strides (a runtime
// division is ~40 instructions, and rbc_insert does up to 24 of them per thread -- they were most of its 19 M warp instructions)
__device__ __forceinline__ int floor_div_stride(const ConvGeom &g, int d, int a) { return g.sshift[d] >= 0 ? a >> g.sshift[d] : floor_div(a, g.stride[d]); }
__device__ __forceinline__ int ceil_div_stride(const ConvGeom &g, int d, int a) { return -floor_div_stride(g, d, -a); }

// row id of cell `cell` of a level >= 1 (after rbc_assign), or -1
__device__ __forceinline__ int cell_row(const unsigned long long *__restrict__ words, uint32_t cell)
{
    const unsigned long long w = __ldg(words + (cell >> 5));
    const uint32_t bits = (uint32_t)w, b = cell & 31u;
    if (!((bits >> b) & 1u)) return -1;
    return (int)(uint32_t)(w >> 32) + __popc(bits & ((1u << b) - 1u));
}

// grid: (blocks over the level-0 rows (grid-stride), 1 + 3 * (n_levels - 1)).  y = 0: level 0 itself into its hash table;
// y = 1 + 3*(l-1) + zi: the z-slice zi of the box voxel r reaches at level l (zi = 2 also takes whatever lies beyond a
// 3-deep box).  Neighbouring voxels reach the same sites: a bit already set costs one read and no atomic.
__global__ void __launch_bounds__(256) rbc_insert(const ChainParams P)
{
    const ChainLevel &l0 = P.lv[0];
    const int n = row_count(l0.cap, l0.count);
    const int lvl = blockIdx.y == 0 ? 0 : 1 + (blockIdx.y - 1) / 3, zi = blockIdx.y == 0 ? 0 : (blockIdx.y - 1) % 3;
    const ChainLevel &L = P.lv[lvl];
    unsigned int *bits32 = reinterpret_cast<unsigned int *>(L.slots);         // low half of word w = bits32[2w]
    for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < n; r += gridDim.x * blockDim.x) {
        const int4 c = __ldg(l0.coords + r);
        if (lvl == 0) {
            if (table_insert_min(l0.slots, l0.mask, lin_index(c.x, c.y, c.z, c.w, l0.subm.in_shape), (uint32_t)r) == 0xFFFFFFFFu)
                *P.overflow = 1;
            continue;
        }
        int lo[3] = {c.y, c.z, c.w}, hi[3] = {c.y, c.z, c.w};
        bool empty = false;
        for (int j = 1; j <= lvl; ++j) {
            const ConvGeom &g = P.lv[j].conv;
#pragma unroll
            for (int d = 0; d < 3; ++d) {
                // outputs o whose window [o*s - p, o*s - p + k - 1] meets [lo, hi]; every one of them is active because
                // k >= s (checked on the host) and every point of the input box is
                lo[d] = max(ceil_div_stride(g, d, lo[d] + g.pad[d] - (g.ksize[d] - 1)), 0);
                hi[d] = min(floor_div_stride(g, d, hi[d] + g.pad[d]), g.out_shape[d] - 1);
            }
            empty |= lo[0] > hi[0] || lo[1] > hi[1] || lo[2] > hi[2];
        }
        if (empty) continue;
        const int z0 = lo[0] + zi, z1 = zi == 2 ? hi[0] : min(z0, hi[0]);
        for (int z = z0; z <= z1; ++z)
            for (int y = lo[1]; y <= hi[1]; ++y) {
                // the cells [lo_x, hi_x] of this row of the box: one or two words.  A reduction without a return value
                // (RED): the thread does not wait for it -- with the old value returned (to count the fresh bits here)
                // every box row was a dependent L2 round trip and the kernel took twice as long.
                const uint32_t first = lin_index(c.x, z, y, lo[2], L.conv.out_shape), last = first + (uint32_t)(hi[2] - lo[2]);
                for (uint32_t w = first >> 5; w <= (last >> 5); ++w) {
                    const uint32_t a = w == (first >> 5) ? (first & 31u) : 0u, b = w == (last >> 5) ? (last & 31u) : 31u;
                    const uint32_t want = (b == 31u ? 0xFFFFFFFFu : ((1u << (b + 1)) - 1u)) & ~((1u << a) - 1u);
                    // Neighbouring voxels reach the same sites, deeper levels all the more (a level-3 cell is reached from
                    // up to 15^3 voxels): a plain load first -- through L1, possibly stale, which only means an atomic that
                    // was not needed -- keeps the already-set words away from the L2 atomic units (nuScenes batch of 4:
                    // 27 M reductions).  PCDB_RBC_CHECK=0 at build time restores the unconditional reduction.
#if PCDB_RBC_CHECK
                    if ((__ldg(bits32 + 2 * (size_t)w) & want) == want) continue;
#endif
                    atomicOr(bits32 + 2 * (size_t)w, want);
                }
            }
    }
}

__device__ __forceinline__ int scan_level(const ChainParams &P, int block)
{
    int l = 1;
    while (l + 1 < P.n_levels && block >= P.lv[l + 1].scan_block0) ++l;
    return l;
}

// Set bits per 1024-word chunk of every level, kChunksPerBlock chunks per thread block (one fence + ticket per block: with a
// block per chunk the 1664 fences and tickets of a KITTI batch cost 20 us); the last block to finish turns the counters of
// all levels into exclusive prefixes (+ total), so that rbc_assign starts from them.  256 threads x 4 consecutive words:
// light blocks that find room next to the convolutions' CTAs (1024-thread blocks waited for whole SMs).
constexpr int kChunksPerBlock = 8;
constexpr int kScanThreads = kChainScanBlock / 4;

__device__ __forceinline__ void load_words4(const ChainLevel &L, uint32_t w0, uint32_t bits[4])
{
    if (w0 + 4 <= L.n_words) {
        const uint4 a = *reinterpret_cast<const uint4 *>(L.slots + w0), b = *reinterpret_cast<const uint4 *>(L.slots + w0 + 2);
        bits[0] = a.x; bits[1] = a.z; bits[2] = b.x; bits[3] = b.z;
    } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) bits[j] = w0 + j < L.n_words ? (uint32_t)L.slots[w0 + j] : 0u;
    }
}

__global__ void __launch_bounds__(kScanThreads) rbc_count(const ChainParams P, int total_blocks)
{
    __shared__ int s_part[kChunksPerBlock][kScanThreads / 32];
    __shared__ bool s_last;
    const int l = scan_level(P, blockIdx.x);
    const ChainLevel &L = P.lv[l];
    const int chunk0 = (blockIdx.x - L.scan_block0) * kChunksPerBlock;
    int v[kChunksPerBlock];
#pragma unroll
    for (int c = 0; c < kChunksPerBlock; ++c) {          // all loads in flight together
        uint32_t bits[4];
        load_words4(L, (uint32_t)(chunk0 + c) * kChainScanBlock + threadIdx.x * 4, bits);
        v[c] = __popc(bits[0]) + __popc(bits[1]) + __popc(bits[2]) + __popc(bits[3]);
    }
#pragma unroll
    for (int c = 0; c < kChunksPerBlock; ++c) {
#pragma unroll
        for (int d = 16; d; d >>= 1) v[c] += __shfl_xor_sync(0xffffffffu, v[c], d);
        if ((threadIdx.x & 31) == 0) s_part[c][threadIdx.x >> 5] = v[c];
    }
    __syncthreads();
    if (threadIdx.x < kChunksPerBlock && chunk0 + (int)threadIdx.x < L.nblocks) {
        int t = 0;
        for (int i = 0; i < kScanThreads / 32; ++i) t += s_part[threadIdx.x][i];
        L.block_sums[chunk0 + threadIdx.x] = t;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        s_last = atomicAdd(P.ticket, 1u) + 1u == (unsigned int)total_blocks;
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    // One scan over the counters of ALL levels laid end to end (a block scan per 256 counters and level was nine dependent
    // rounds of global loads and barriers, longer than the counting itself); a level's prefixes are the global ones minus
    // the global prefix at its first chunk.
    __shared__ int s_level_base[kChainMaxLevels + 1];
    int first[kChainMaxLevels + 1];                      // index of each level's first counter in the concatenation
    first[1] = 0;
    for (int lv = 1; lv < P.n_levels; ++lv) first[lv + 1] = first[lv] + P.lv[lv].nblocks;
    const int total = first[P.n_levels], per = (total + kScanThreads - 1) / kScanThreads;
    const int begin = min((int)threadIdx.x * per, total), end = min(begin + per, total);
    auto counter = [&](int i, int *lv_out) -> volatile int * {
        int lv = 1;
        while (lv + 1 < P.n_levels && i >= first[lv + 1]) ++lv;
        *lv_out = lv;
        return (volatile int *)P.lv[lv].block_sums + (i - first[lv]);
    };
    int sum = 0, lv;
    for (int i = begin; i < end; ++i) sum += *counter(i, &lv);
    const int ex = block_exclusive_scan<kScanThreads>(sum, nullptr);
    if (threadIdx.x == 0) s_level_base[P.n_levels] = 0;      // (no counters at all)
    __syncthreads();
    int run = ex;
    for (int i = begin; i < end; ++i) {                   // the global prefix at every level's first chunk, and the grand total
        const int v = *counter(i, &lv);
        if (i == first[lv]) s_level_base[lv] = run;
        run += v;
    }
    if (begin < total && end == total) s_level_base[P.n_levels] = run;
    __syncthreads();
    run = ex;
    for (int i = begin; i < end; ++i) {
        const int v = *counter(i, &lv);
        P.lv[lv].block_prefix[i - first[lv]] = run - s_level_base[lv];
        run += v;
    }
    for (int l2 = 1 + (int)threadIdx.x; l2 < P.n_levels; l2 += kScanThreads)
        P.lv[l2].block_prefix[P.lv[l2].nblocks] = s_level_base[l2 + 1] - s_level_base[l2];
}

// Every thread numbers 4 consecutive occupancy words (one chunk of 1024 words per block, empty chunks leave at once):
// rows before a word (chunk prefix from rbc_count + scan inside the chunk) into its high half, coordinates of its set bits
// into the row table.
__global__ void __launch_bounds__(kScanThreads) rbc_assign(const ChainParams P)
{
    int l = 1;
    while (l + 1 < P.n_levels && (int)blockIdx.x >= P.lv[l + 1].assign_block0) ++l;
    const ChainLevel &L = P.lv[l];
    const int chunk0 = blockIdx.x - L.assign_block0;      // one chunk per block: a dense region must not queue up in one block
    const int s_total = L.block_prefix[L.nblocks];
    const int *shape = L.conv.out_shape;
    if (chunk0 == 0 && threadIdx.x == 0) {
        L.count_out[0] = s_total < L.cap ? s_total : L.cap;
        L.count_out[1] = (s_total > L.cap || *P.overflow) ? 1 : 0;
    }
    {
        const int blk = chunk0;
        const int s_before = L.block_prefix[blk];
        if (L.block_prefix[blk + 1] == s_before) return;            // (uniform over the block) nothing to number here
        const uint32_t w0 = (uint32_t)blk * kChainScanBlock + threadIdx.x * 4;
        uint32_t bits4[4];
        load_words4(L, w0, bits4);
        int id = block_exclusive_scan<kScanThreads>(__popc(bits4[0]) + __popc(bits4[1]) + __popc(bits4[2]) + __popc(bits4[3]), nullptr) + s_before;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            uint32_t bits = bits4[j];
            if (!bits) continue;
            const uint32_t w = w0 + j;
            L.slots[w] = ((unsigned long long)(uint32_t)id << 32) | bits;
            uint32_t cell = w << 5;
            const int x0 = (int)(cell % (uint32_t)shape[2]); cell /= (uint32_t)shape[2];
            const int y0 = (int)(cell % (uint32_t)shape[1]); cell /= (uint32_t)shape[1];
            const int z0 = (int)(cell % (uint32_t)shape[0]); cell /= (uint32_t)shape[0];
            for (; bits; bits &= bits - 1, ++id) {
                if (id >= L.cap) continue;
                // cell 32w + t: carry x over the row / plane / sample boundaries a word may straddle
                int x = x0 + __ffs(bits) - 1, y = y0, z = z0, bb = (int)cell;
                while (x >= shape[2]) { x -= shape[2]; if (++y == shape[1]) { y = 0; if (++z == shape[0]) { z = 0; ++bb; } } }
                L.coords[id] = make_int4(bb, z, y, x);
            }
        }
    }
}

// Flat grid over the maps [first_map, ...) (grid-stride over the rows of each).
//   strided conv into level l, input-driven: thread (input row r, (mz, my)) walks the candidates along x -- their output
//     cells are neighbours in the occupancy words;
//   SubM at level 0 (hash table), centred offsets: thread (row r, row of offsets (dz, dy)) probes dx ascending for the
//     offsets k < K/2 and writes both directions; the thread of the last such row also writes the centre (the site itself);
//   SubM at a level >= 1 (occupancy words): thread (row r, row of offsets (dz, dy)) looks up all its offsets.
__global__ void __launch_bounds__(256) rbc_maps(const ChainParams P, int first_map, int block_base)
{
    int m = first_map;
    const int gb = blockIdx.x + block_base;
    while (m + 1 < P.n_maps && gb >= P.maps[m + 1].block0) ++m;
    const ChainMap mp = P.maps[m];
    const int local = gb - mp.block0, bx = local % mp.blocks_x, by = local / mp.blocks_x;
    const ChainLevel &L = P.lv[mp.level];
    if (mp.kind == 0) {
        const ChainLevel &I = P.lv[mp.level - 1];
        const ConvGeom &g = L.conv;
        const int n = row_count(I.cap, I.count);
        for (int r = bx * 256 + threadIdx.x; r < n; r += mp.blocks_x * 256) {
            const int4 c = __ldg(I.coords + r);
            for (int cx = 0; cx < g.comb[2]; ++cx) {
                int k, oz, oy, ox;
                if (!conv_candidate(g, c, by * g.comb[2] + cx, &k, &oz, &oy, &ox)) continue;
                const int oid = cell_row(L.slots, lin_index(c.x, oz, oy, ox, g.out_shape));
                if (oid >= 0 && oid < L.cap) L.nbr_conv[(size_t)k * L.cap + oid] = r;
            }
        }
    } else {
        const ConvGeom &g = L.subm;
        const int n = row_count(L.cap, L.count), half = g.K >> 1, kx = g.ksize[2];
        if (mp.level > 0) {
            // levels >= 1: EVERY offset of the row (dz, dy) is looked up -- the cells x-1, x, x+1 share an occupancy word, so
            // the second direction is cheaper to probe than to scatter -- and every entry is written, hit or not: these
            // maps need no clearing between builds
            for (int r = bx * 256 + threadIdx.x; r < n; r += mp.blocks_x * 256) {
                const int4 c = __ldg(L.coords + r);
                const int z = c.y - g.pad[0] + g.dk[by * kx][0], y = c.z - g.pad[1] + g.dk[by * kx][1];
                const bool row_ok = z >= 0 && z < g.in_shape[0] && y >= 0 && y < g.in_shape[1];
                for (int j = 0; j < kx; ++j) {
                    const int k = by * kx + j, x = c.w - g.pad[2] + g.dk[k][2];
                    int hit = -1;
                    if (k == half) hit = r;
                    else if (row_ok && x >= 0 && x < g.in_shape[2]) hit = cell_row(L.slots, lin_index(c.x, z, y, x, g.in_shape));
                    L.nbr_subm[(size_t)k * L.cap + r] = hit < n ? hit : -1;
                }
            }
            return;
        }
        for (int r = bx * 256 + threadIdx.x; r < n; r += mp.blocks_x * 256) {
            const int4 c = __ldg(L.coords + r);
            for (int j = 0; j < kx; ++j) {
                const int k = by * kx + j;
                if (k > half) break;
                if (k == half) { L.nbr_subm[(size_t)k * L.cap + r] = r; break; }
                const int z = c.y - g.pad[0] + g.dk[k][0], y = c.z - g.pad[1] + g.dk[k][1], x = c.w - g.pad[2] + g.dk[k][2];
                if (z < 0 || z >= g.in_shape[0] || y < 0 || y >= g.in_shape[1] || x < 0 || x >= g.in_shape[2]) continue;
                uint32_t payload;
                const int hit = table_find(L.slots, L.mask, lin_index(c.x, z, y, x, g.in_shape), &payload) == 0xFFFFFFFFu ? -1 : (int)payload;
                if (hit >= 0 && hit < n) {
                    L.nbr_subm[(size_t)k * L.cap + r] = hit;
                    L.nbr_subm[(size_t)(g.K - 1 - k) * L.cap + hit] = r;
                }
            }
        }
    }
}

// Leaves the occupancy words and counters of the levels >= 1 zero again by undoing what this build set -- one scattered
// store per row -- instead of a 13 MB memset in front of the next build (KITTI batch of 4: 226 k rows).  A level that
// overflowed its row capacity has set bits without a row: its words are cleared wholesale.  grid: (blocks, n_levels - 1).
__global__ void __launch_bounds__(256) rbc_undo(const ChainParams P)
{
    const ChainLevel &L = P.lv[1 + blockIdx.y];
    const int n = L.count_out[0];
    if (L.count_out[1]) {
        for (uint32_t w = blockIdx.x * blockDim.x + threadIdx.x; w < L.n_words; w += gridDim.x * blockDim.x) L.slots[w] = 0ull;
    } else {
        for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < n; r += gridDim.x * blockDim.x) {
            const int4 c = __ldg(L.coords + r);
            const uint32_t w = lin_index(c.x, c.y, c.z, c.w, L.conv.out_shape) >> 5;
            L.slots[w] = 0ull;
        }
    }
    if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) { *P.overflow = 0; *P.ticket = 0u; }
}

// dst[k*ld + r] = -1 for k < K, r < min(*rows, cap), for every map of the list in one launch: what the previous build of
// each neighbour map wrote (its extent is still in the level's device counter), instead of K x capacity memsets.
struct ChainFill { int *dst; const int *rows; int ld, K, cap, block0; };
struct ChainFillList { int n; ChainFill f[2 * kChainMaxLevels]; };

__global__ void __launch_bounds__(256) rbc_fill_maps(const ChainFillList F)
{
    int m = 0;
    while (m + 1 < F.n && (int)blockIdx.x >= F.f[m + 1].block0) ++m;
    const ChainFill f = F.f[m];
    const int chunks = (f.cap + 1023) / 1024, local = blockIdx.x - f.block0, k = local / chunks;
    int n = __ldg(f.rows);
    n = n < 0 ? 0 : (n < f.cap ? n : f.cap);
    const int r0 = ((local % chunks) * 256 + threadIdx.x) * 4;
    if (r0 >= n) return;
    int *p = f.dst + (size_t)k * f.ld + r0;
    if (r0 + 4 <= n && ((((size_t)k * f.ld) & 3) == 0) && ((reinterpret_cast<uintptr_t>(f.dst) & 15) == 0)) *reinterpret_cast<int4 *>(p) = make_int4(-1, -1, -1, -1);
    else for (int j = 0; j < 4 && r0 + j < n; ++j) p[j] = -1;
}

constexpr uint64_t kChainMaxCells = 1ull << 31;      // occupancy words of a level >= 1: 8 B per 32 cells, <= 512 MB

struct ChainWorkspace {
    unsigned long long *slots[kChainMaxLevels];
    uint32_t table_cap;                         // level-0 hash table
    uint32_t n_words[kChainMaxLevels];          // levels >= 1
    int *overflow;
    int *block_sums[kChainMaxLevels], *block_prefix[kChainMaxLevels];
    unsigned int *ticket;
    int nblocks[kChainMaxLevels];
    size_t fill_bytes, zero_off, zero_bytes, bytes;
};

// cells[l] = batch * volume of level l
static ChainWorkspace carve_chain(void *base, int n_levels, const int32_t *caps, const uint64_t *cells)
{
    ChainWorkspace w{};
    size_t off = 0;
    char *b = (char *)base;
    auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes, 256); return b ? (void *)(b + o) : (void *)nullptr; };
    const size_t sites = (size_t)(caps[0] > 0 ? caps[0] : 1);
    w.table_cap = next_pow2(sites * 2 < 1024 ? 1024 : sites * 2);
    w.slots[0] = (unsigned long long *)take((size_t)w.table_cap * 8);
    w.fill_bytes = off;                          // the level-0 table: 0xFF before every build
    w.zero_off = off;                            // occupancy words, block counters, overflow flag: 0 before every build
    for (int l = 1; l < n_levels; ++l) {
        w.n_words[l] = (uint32_t)((cells[l] + 31) / 32);
        w.nblocks[l] = (int)((w.n_words[l] + kChainScanBlock - 1) / kChainScanBlock);
        w.slots[l] = (unsigned long long *)take((size_t)w.n_words[l] * 8);
        w.block_sums[l] = (int *)take((size_t)w.nblocks[l] * 4);
    }
    w.overflow = (int *)take(4);
    w.ticket = (unsigned int *)take(4);
    w.zero_bytes = off - w.zero_off;
    for (int l = 1; l < n_levels; ++l) w.block_prefix[l] = (int *)take(((size_t)w.nblocks[l] + 1) * 4);
    w.bytes = off;
    return w;
}

static bool chain_cells(int batch, int n_levels, const int32_t *shapes_zyx, uint64_t *cells)
{
    for (int l = 0; l < n_levels; ++l) {
        const int32_t *s = shapes_zyx + 3 * l;
        if (s[0] < 1 || s[1] < 1 || s[2] < 1) return false;
        cells[l] = (uint64_t)batch * s[0] * (uint64_t)s[1] * s[2];
        if (cells[l] >= 0xFFFFFFFFull || (l > 0 && cells[l] > kChainMaxCells)) return false;
    }
    return true;
}

}  // namespace pcdb

using namespace pcdb;

extern "C" size_t pcdb_rulebook_chain_workspace_bytes(int batch, int n_levels, const int32_t *shapes_zyx, const int32_t *caps)
{
    uint64_t cells[kChainMaxLevels];
    if (batch < 1 || n_levels < 1 || n_levels > kChainMaxLevels || !caps || !shapes_zyx || !chain_cells(batch, n_levels, shapes_zyx, cells))
        return 0;
    return carve_chain(nullptr, n_levels, caps, cells).bytes;
}

extern "C" int pcdb_rulebook_chain_clear(void *workspace, size_t workspace_bytes, int batch, int n_levels, const int32_t *shapes_zyx,
                                         const int32_t *caps, const int32_t *ksize_zyx, const int32_t *subm_ksize_zyx,
                                         const int32_t *n0_dev, int32_t *const *counts, int32_t *const *nbr_conv,
                                         int32_t *const *nbr_subm, void *stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    uint64_t cells[kChainMaxLevels];
    if (batch < 1 || n_levels < 1 || n_levels > kChainMaxLevels || !caps || !shapes_zyx || !chain_cells(batch, n_levels, shapes_zyx, cells)) {
        set_last_error("pcdb_rulebook_chain_clear: invalid argument");
        return kInvalidArgument;
    }
    if (workspace) {
        const ChainWorkspace w = carve_chain(workspace, n_levels, caps, cells);
        if (workspace_bytes < w.bytes) {
            set_last_error("pcdb_rulebook_chain_clear: workspace %zu < required %zu bytes", workspace_bytes, w.bytes);
            return kWorkspaceTooSmall;
        }
        cudaMemsetAsync(workspace, 0xFF, w.fill_bytes, stream);
        cudaMemsetAsync((char *)workspace + w.zero_off, 0, w.zero_bytes, stream);
    }
    if (nbr_conv && nbr_subm && counts && n0_dev && ksize_zyx && subm_ksize_zyx) {
        // the extents of the maps: what their previous build wrote
        ChainFillList F{};
        int blocks = 0;
        auto add = [&](int32_t *dst, int l, int K) {
            if (!dst || K < 1) return;
            ChainFill &f = F.f[F.n++];
            f.dst = dst; f.rows = l == 0 ? n0_dev : counts[l]; f.ld = caps[l]; f.K = K; f.cap = caps[l]; f.block0 = blocks;
            blocks += ((caps[l] + 1023) / 1024) * K;
        };
        for (int l = 0; l < n_levels; ++l) {
            if (l > 0) { const int32_t *k = ksize_zyx + 3 * (l - 1); add(nbr_conv[l], l, k[0] * k[1] * k[2]); }
            const int32_t *sk = subm_ksize_zyx + 3 * l;
            if (sk[0] > 0 && l == 0) add(nbr_subm[l], l, sk[0] * sk[1] * sk[2]);      // (levels >= 1: every entry is rewritten)
        }
        if (blocks > 0) rbc_fill_maps<<<blocks, 256, 0, stream>>>(F);
    }
    return check_launch("pcdb_rulebook_chain_clear");
}

extern "C" int pcdb_rulebook_chain(const int32_t *coords0, const int32_t *n0_dev, int batch, int n_levels,
                                   const int32_t *shapes_zyx, const int32_t *ksize_zyx, const int32_t *stride_zyx,
                                   const int32_t *padding_zyx, const int32_t *caps, int32_t *const *coords,
                                   int32_t *const *counts, int32_t *const *nbr_conv, const int32_t *subm_ksize_zyx,
                                   int32_t *const *nbr_subm, const int32_t *rows_hint, void *workspace, size_t workspace_bytes,
                                   int flags, int phase, void *stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    if (phase < 1 || phase > 63 || ((phase & 32) && !(phase & 2))) {
        set_last_error("pcdb_rulebook_chain: phase is a mask of 1, 2, 4, 8, 16 and 32 (32 only together with 2)");
        return kInvalidArgument;
    }
    if (!coords0 || !n0_dev || batch < 1 || n_levels < 1 || n_levels > kChainMaxLevels || !shapes_zyx || !caps || !coords || !counts ||
        !nbr_conv || !subm_ksize_zyx || !nbr_subm || (n_levels > 1 && (!ksize_zyx || !stride_zyx || !padding_zyx))) {
        set_last_error("pcdb_rulebook_chain: invalid argument (n_levels=%d batch=%d)", n_levels, batch);
        return kInvalidArgument;
    }
    uint64_t cells[kChainMaxLevels];
    if (!chain_cells(batch, n_levels, shapes_zyx, cells)) {
        set_last_error("pcdb_rulebook_chain: batch*volume exceeds the 32-bit cell index (level 0) or 2^31 cells (levels >= 1)");
        return kKeyOverflow;
    }
    const ChainWorkspace w = carve_chain(workspace, n_levels, caps, cells);
    if (!workspace || workspace_bytes < w.bytes) {
        set_last_error("pcdb_rulebook_chain: workspace %zu < required %zu bytes", workspace_bytes, w.bytes);
        return kWorkspaceTooSmall;
    }
    ChainParams P{};
    P.n_levels = n_levels; P.batch = batch; P.overflow = w.overflow; P.ticket = w.ticket;
    const int32_t one[3] = {1, 1, 1};
    int scan_blocks = 0, assign_blocks = 0, map_blocks = 0;
    for (int l = 0; l < n_levels; ++l) {
        ChainLevel &L = P.lv[l];
        const int32_t *shape = shapes_zyx + 3 * l;
        if (caps[l] < 1 || (l > 0 && (!coords[l] || !counts[l]))) {
            set_last_error("pcdb_rulebook_chain: level %d: capacity / buffers missing", l);
            return kInvalidArgument;
        }
        if (l > 0) {
            const int c = 3 * (l - 1);
            if (!fill_geom(L.conv, shapes_zyx + 3 * (l - 1), shape, ksize_zyx + c, stride_zyx + c, padding_zyx + c, one)) {
                set_last_error("pcdb_rulebook_chain: conv %d: unsupported geometry", l);
                return kInvalidArgument;
            }
            for (int d = 0; d < 3; ++d)
                if (L.conv.ksize[d] < L.conv.stride[d]) {      // the box argument of rbc_insert needs overlapping / abutting windows
                    set_last_error("pcdb_rulebook_chain: conv %d: kernel size < stride is not supported", l);
                    return kUnsupported;
                }
            if (!nbr_conv[l]) { set_last_error("pcdb_rulebook_chain: conv %d: neighbour map missing", l); return kInvalidArgument; }
        } else {
            if (!fill_geom(L.conv, shape, shape, one, one, nullptr, one)) return kInvalidArgument;      // only out_shape is read
        }
        const int32_t *sk = subm_ksize_zyx + 3 * l;
        L.subm.K = 0;
        if (sk[0] > 0) {
            if (!fill_geom(L.subm, shape, nullptr, sk, nullptr, nullptr, one) || !symmetric_offsets(L.subm) || !nbr_subm[l]) {
                set_last_error("pcdb_rulebook_chain: level %d: SubM kernel must be odd (centred offsets) and have a map buffer", l);
                return kUnsupported;
            }
        } else {
            if (!fill_geom(L.subm, shape, nullptr, one, nullptr, nullptr, one)) return kInvalidArgument;   // in_shape for level 0's key
            L.subm.K = 0;
        }
        L.slots = w.slots[l]; L.mask = w.table_cap - 1; L.n_words = w.n_words[l];
        L.coords = (int4 *)(l == 0 ? const_cast<int32_t *>(coords0) : coords[l]);
        L.count = l == 0 ? n0_dev : counts[l];
        L.count_out = l == 0 ? nullptr : counts[l];
        L.cap = caps[l];
        L.nbr_conv = nbr_conv[l]; L.nbr_subm = nbr_subm[l];
        L.block_sums = w.block_sums[l]; L.block_prefix = w.block_prefix[l]; L.nblocks = w.nblocks[l];
        L.scan_block0 = scan_blocks;
        L.assign_block0 = assign_blocks;
        if (l > 0) { scan_blocks += (w.nblocks[l] + kChunksPerBlock - 1) / kChunksPerBlock; assign_blocks += w.nblocks[l]; }
    }
    P.lv[0].scan_block0 = 0;
    // grid sizes: capacities, or -- when the caller knows what to expect (a captured pipeline after its warm-up step) --
    // the hinted row counts plus a margin; every kernel strides over the rows, so a hint never changes the result
    auto rows_for_grid = [&](int l) {
        if (!rows_hint || rows_hint[l] <= 0) return caps[l];
        const long long want = (long long)rows_hint[l] + rows_hint[l] / 4 + 256;
        return (int)(want < caps[l] ? want : caps[l]);
    };
    // the SubM map of level 0 first: it only needs the level-0 table, i.e. phase 1
    auto add_map = [&](int kind, int l) {
        ChainMap &m = P.maps[P.n_maps++];
        m.kind = kind; m.level = l; m.block0 = map_blocks;
        if (kind == 0) { m.blocks_x = (rows_for_grid(l - 1) + 255) / 256; m.rows_y = P.lv[l].conv.comb[0] * P.lv[l].conv.comb[1]; }
        else {
            m.blocks_x = (rows_for_grid(l) + 255) / 256;
            m.rows_y = l == 0 ? (P.lv[l].subm.K / 2 + P.lv[l].subm.ksize[2]) / P.lv[l].subm.ksize[2] : P.lv[l].subm.ksize[0] * P.lv[l].subm.ksize[1];
        }
        map_blocks += m.blocks_x * m.rows_y;
    };
    const bool map0 = P.lv[0].subm.K > 0;
    if (map0) add_map(1, 0);
    for (int l = 1; l < n_levels; ++l) {
        add_map(0, l);
        if (P.lv[l].subm.K > 0) add_map(1, l);
    }
    const int blocks_phase1 = map0 ? P.maps[0].blocks_x * P.maps[0].rows_y : 0;
    // first map that belongs to a level >= 2 (maps are laid out level by level)
    int first_map_level2 = P.n_maps, blocks_level1_end = map_blocks;
    for (int m = 0; m < P.n_maps; ++m)
        if (P.maps[m].level >= 2) { first_map_level2 = m; blocks_level1_end = P.maps[m].block0; break; }
    if ((phase & 1) && !(flags & PCDB_RB_CLEARED)) {
        cudaMemsetAsync(workspace, 0xFF, w.fill_bytes, stream);
        cudaMemsetAsync((char *)workspace + w.zero_off, 0, w.zero_bytes, stream);
        for (int l = 0; l < n_levels; ++l) {
            if (l > 0) cudaMemsetAsync(nbr_conv[l], 0xFF, sizeof(int32_t) * (size_t)P.lv[l].conv.K * caps[l], stream);
            if (P.lv[l].subm.K > 0 && l == 0) cudaMemsetAsync(nbr_subm[l], 0xFF, sizeof(int32_t) * (size_t)P.lv[l].subm.K * caps[l], stream);
        }
    }
    if ((phase & 1) && (flags & PCDB_RB_CLEARED) && (flags & PCDB_RB_UNDONE)) cudaMemsetAsync(workspace, 0xFF, w.fill_bytes, stream);
    if (phase & 1) rbc_insert<<<dim3((rows_for_grid(0) + 255) / 256, 1 + 3 * (n_levels - 1)), 256, 0, stream>>>(P);
    if ((phase & 4) && blocks_phase1 > 0) rbc_maps<<<blocks_phase1, 256, 0, stream>>>(P, 0, 0);
    if (phase & 2) {
        if (n_levels > 1) {
            rbc_count<<<scan_blocks, kScanThreads, 0, stream>>>(P, scan_blocks);
            rbc_assign<<<assign_blocks, kScanThreads, 0, stream>>>(P);
        }
        // with bit 32 only the maps of level 1 (its strided conv and its SubM map: what the second group of convolutions
        // waits for); the maps of the deeper levels then come from a later call with bit 16
        const int end = (phase & 32) ? blocks_level1_end : map_blocks;
        if (end > blocks_phase1) rbc_maps<<<end - blocks_phase1, 256, 0, stream>>>(P, map0 ? 1 : 0, blocks_phase1);
    }
    if ((phase & 16) && map_blocks > blocks_level1_end)
        rbc_maps<<<map_blocks - blocks_level1_end, 256, 0, stream>>>(P, first_map_level2, blocks_level1_end);
    if ((phase & 8) && n_levels > 1) {
        int rows = 1;
        for (int l = 1; l < n_levels; ++l) rows = rows_for_grid(l) > rows ? rows_for_grid(l) : rows;
        rbc_undo<<<dim3((rows + 255) / 256, n_levels - 1), 256, 0, stream>>>(P);
    }
    return check_launch("pcdb_rulebook_chain");
}
