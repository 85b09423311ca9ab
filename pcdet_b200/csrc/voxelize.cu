// Point -> voxel hashing and the mean VFE for sm_100a.
//
// Replaces spconv v1.0 points_to_voxel_3d_np (serial CPU loop over a dense 369 MB lookup grid; SURVEY
// App. A.1, called from pcdet/datasets/dataset.py:163), the collate of dataset.py:266-299 and
// MeanVoxelFeatureExtractor.forward (pcdet/models/vfe/vfe_utils.py:26-34).
//
// The serial loop defines voxel ids by first appearance and keeps the first P points of a voxel.
// Both are functions of the ORIGINAL POINT INDEX only, so they can be recovered in parallel:
//   1. insert every in-range point into an open-addressing hash table keyed by its linear cell,
//      keeping the minimum point index per cell (one 64-bit atomicMin per distinct cell per warp);
//   2. a point is its cell's owner iff it holds that minimum; an exclusive scan of the owner flags
//      in point order is exactly the reference's first-appearance voxel id;
//   3. every point is pushed through a cascade of atomicMin over its voxel's P slots, which leaves
//      the P smallest point indices of the voxel in ascending order, whatever the interleaving;
//   4. one thread per voxel gathers those points, writes the padded voxel tensor / coordinates /
//      count and accumulates the mean in index order (fp32 add, IEEE divide).
// Traffic per frame: 16 B/point read twice + 8 B hash slot, V*(P*C*4 + 20) B written.
#include "common.cuh"
#include "../../include/pcdet_b200.h"

namespace pcdb {

constexpr int kMaxBatch = 256;
constexpr int kScanBlock = 1024;
constexpr unsigned int kEmptyIdx = 0xFFFFFFFFu;

struct VoxParams {
    float lo[3];
    float vs[3];
    int grid[3];  // x, y, z
    int n_points, n_feat, batch, max_points, max_voxels, overflow_break;
};

// upper_bound over frame_offsets[1..batch]: frame of point i
__device__ __forceinline__ int frame_of(const int *offs, int batch, int i)
{
    int lo = 0, hi = batch;  // answer in [0, batch)
    while (hi - lo > 1) {
        int mid = (lo + hi) >> 1;
        if (offs[mid] <= i) lo = mid; else hi = mid;
    }
    return lo;
}

// fp32 subtract, IEEE divide, floor -- the reference expression, never a reciprocal multiply.
__device__ __forceinline__ bool cell_of(const VoxParams &p, float x, float y, float z, int *cx, int *cy, int *cz)
{
    const float fx = floorf(__fdiv_rn(__fsub_rn(x, p.lo[0]), p.vs[0]));
    const float fy = floorf(__fdiv_rn(__fsub_rn(y, p.lo[1]), p.vs[1]));
    const float fz = floorf(__fdiv_rn(__fsub_rn(z, p.lo[2]), p.vs[2]));
    // comparisons in float so that NaN / huge values are rejected before the int conversion
    const bool ok = fx >= 0.f && fx < (float)p.grid[0] && fy >= 0.f && fy < (float)p.grid[1] &&
                    fz >= 0.f && fz < (float)p.grid[2];
    *cx = (int)fx; *cy = (int)fy; *cz = (int)fz;
    return ok;
}

__global__ void __launch_bounds__(256)
vox_hash_insert(const float *__restrict__ points, const int *__restrict__ frame_offsets, VoxParams p,
                unsigned long long *slots, uint32_t mask, int *__restrict__ pt_slot)
{
    __shared__ int s_off[kMaxBatch + 1];
    for (int t = threadIdx.x; t <= p.batch; t += blockDim.x) s_off[t] = frame_offsets[t];
    __syncthreads();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    bool valid = false;
    uint32_t key = 0xFFFFFFFFu;
    if (i < p.n_points && i >= s_off[0] && i < s_off[p.batch]) {
        float x, y, z;
        if (p.n_feat == 4) {
            const float4 v = __ldg(reinterpret_cast<const float4 *>(points) + i);
            x = v.x; y = v.y; z = v.z;
        } else {
            const float *q = points + (size_t)i * p.n_feat;
            x = __ldg(q); y = __ldg(q + 1); z = __ldg(q + 2);
        }
        int cx, cy, cz;
        if (cell_of(p, x, y, z, &cx, &cy, &cz)) {
            const int b = frame_of(s_off, p.batch, i);
            key = (uint32_t)(((b * p.grid[2] + cz) * p.grid[1] + cy)) * (uint32_t)p.grid[0] + (uint32_t)cx;
            valid = true;
        }
    }
    // warp-cooperative insert: lanes that fall into the same cell elect the lowest lane (= smallest
    // point index) to touch the table; the others reuse its slot.
    const unsigned peers = __match_any_sync(0xffffffffu, key);
    const int leader = __ffs(peers) - 1;
    uint32_t slot = 0xFFFFFFFFu;
    if (valid && (int)(threadIdx.x & 31) == leader) slot = table_insert_min(slots, mask, key, (uint32_t)i);
    slot = __shfl_sync(0xffffffffu, slot, leader);
    if (i < p.n_points) pt_slot[i] = valid ? (int)slot : -1;
}

__device__ __forceinline__ bool is_owner(const unsigned long long *slots, const int *pt_slot, int i, int n, int *slot_out)
{
    if (i >= n) return false;
    const int s = pt_slot[i];
    *slot_out = s;
    if (s < 0) return false;
    return (uint32_t)slots[s] == (uint32_t)i;
}

__global__ void __launch_bounds__(kScanBlock)
vox_count_owners(const unsigned long long *__restrict__ slots, const int *__restrict__ pt_slot, int n,
                 int *block_sums, unsigned int *ticket)
{
    const int i = blockIdx.x * kScanBlock + threadIdx.x;
    int s;
    const int cnt = __syncthreads_count(is_owner(slots, pt_slot, i, n, &s));
    if (threadIdx.x == 0) block_sums[blockIdx.x] = cnt;
    last_block_scan<kScanBlock>(block_sums, gridDim.x, ticket);
}

__global__ void __launch_bounds__(kScanBlock)
vox_rank_owners(const unsigned long long *__restrict__ slots, const int *__restrict__ pt_slot,
                const int *__restrict__ frame_offsets, int batch, int n, const int *__restrict__ block_sums,
                int nblocks, int *slot_rank, int *owner_of_rank, int *frame_start)
{
    __shared__ int s_off[kMaxBatch + 1];
    for (int t = threadIdx.x; t <= batch; t += blockDim.x) s_off[t] = frame_offsets[t];
    __syncthreads();
    const int i = blockIdx.x * kScanBlock + threadIdx.x;
    int s = -1;
    const bool own = is_owner(slots, pt_slot, i, n, &s);
    const int rank = block_exclusive_scan<kScanBlock>(own ? 1 : 0, nullptr) + block_sums[blockIdx.x];
    if (own) {
        slot_rank[s] = rank;
        owner_of_rank[rank] = i;
    }
    if (i < n) {
        // voxel rank at every frame boundary (0..batch, the last one being the end sentinel) that
        // sits on point i; n may exceed the real point count when the caller passes a capacity
        int b = frame_of(s_off, batch + 1, i);
        while (b >= 0 && s_off[b] == i) { frame_start[b] = rank; --b; }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        const int total = block_sums[nblocks];
        for (int b = batch; b >= 0 && s_off[b] >= n; --b) frame_start[b] = total;
    }
}

// per-frame bookkeeping shared by the two kernels below (batch <= kMaxBatch, done by thread 0)
struct FrameInfo {
    int start[kMaxBatch + 1];     // first global voxel rank of the frame
    int out_base[kMaxBatch + 1];  // first OUTPUT row of the frame (after the max_voxels clamp)
    int i_break[kMaxBatch];       // first point index dropped by the v1.0 `break`
};

__device__ __forceinline__ void load_frame_info(FrameInfo &f, const int *frame_start, const int *owner_of_rank,
                                                int batch, int max_voxels, int overflow_break)
{
    if (threadIdx.x == 0) {
        int run = 0;
        for (int b = 0; b <= batch; ++b) f.start[b] = frame_start[b];
        for (int b = 0; b < batch; ++b) {
            const int cnt = f.start[b + 1] - f.start[b];
            f.out_base[b] = run;
            run += min(cnt, max_voxels);
            f.i_break[b] = (overflow_break && cnt > max_voxels) ? owner_of_rank[f.start[b] + max_voxels] : 0x7FFFFFFF;
        }
        f.out_base[batch] = run;
    }
    __syncthreads();
}

__global__ void __launch_bounds__(256)
vox_assign_points(const int *__restrict__ pt_slot, const int *__restrict__ slot_rank,
                  const int *__restrict__ frame_offsets, const int *__restrict__ frame_start,
                  const int *__restrict__ owner_of_rank, VoxParams p, unsigned int *vox_pts)
{
    __shared__ FrameInfo f;
    __shared__ int s_off[kMaxBatch + 1];
    for (int t = threadIdx.x; t <= p.batch; t += blockDim.x) s_off[t] = frame_offsets[t];
    load_frame_info(f, frame_start, owner_of_rank, p.batch, p.max_voxels, p.overflow_break);
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= p.n_points) return;
    const int s = pt_slot[i];
    if (s < 0) return;
    const int b = frame_of(s_off, p.batch, i);
    if (i >= f.i_break[b]) return;
    const int local = slot_rank[s] - f.start[b];
    if (local >= p.max_voxels) return;
    unsigned int *list = vox_pts + (size_t)(f.out_base[b] + local) * p.max_points;
    // already P smaller indices in place -> this point is one of the silently dropped extras
    if (((volatile unsigned int *)list)[p.max_points - 1] < (unsigned int)i) return;
    unsigned int v = (unsigned int)i;
    for (int k = 0; k < p.max_points; ++k) {
        const unsigned int old = atomicMin(list + k, v);
        if (old == kEmptyIdx) break;        // took an empty slot
        v = max(old, v);                    // carry the larger of the two onwards
    }
}

constexpr int kGatherMaxP = 10;        // voxel point slots the prefetching gather keeps in registers (KITTI 5, nuScenes 10)

template <typename TMean>
__global__ void __launch_bounds__(128)
vox_gather(const float *__restrict__ points, const unsigned int *__restrict__ vox_pts,
           const int *__restrict__ frame_start, const int *__restrict__ owner_of_rank, VoxParams p,
           float *__restrict__ voxels, int *__restrict__ coords, int *__restrict__ num_points,
           TMean *__restrict__ mean, int mean_stride, int *__restrict__ point_idx, int *__restrict__ voxel_offsets)
{
    __shared__ FrameInfo f;
    load_frame_info(f, frame_start, owner_of_rank, p.batch, p.max_voxels, p.overflow_break);
    if (blockIdx.x == 0)
        for (int t = threadIdx.x; t <= p.batch; t += blockDim.x) voxel_offsets[t] = f.out_base[t];
    const int total = f.out_base[p.batch];
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= total) return;
    int b = frame_of(f.out_base, p.batch, v);
    // empty frames share an out_base with their successor: move to the frame that really owns v
    while (b + 1 < p.batch && f.out_base[b + 1] <= v) ++b;
    const unsigned int *list = vox_pts + (size_t)v * p.max_points;
    const int C = p.n_feat, P = p.max_points;
    int cnt = 0;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    if (C == 4 && P <= kGatherMaxP) {
        // All slot indices first, then all points: two dependent memory round trips per voxel instead of 2 P (the slot-by-slot
        // loop was a chain of 10 / 20 L2 latencies per thread on KITTI / nuScenes voxels).  Sums in slot order, as before.
        unsigned int idx[kGatherMaxP];
        float4 q[kGatherMaxP];
#pragma unroll
        for (int s = 0; s < kGatherMaxP; ++s) idx[s] = s < P ? __ldg(list + s) : kEmptyIdx;
#pragma unroll
        for (int s = 0; s < kGatherMaxP; ++s)
            q[s] = (s < P && idx[s] != kEmptyIdx) ? __ldg(reinterpret_cast<const float4 *>(points) + idx[s]) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int s = 0; s < kGatherMaxP; ++s) {
            if (s >= P) break;
            const bool has = idx[s] != kEmptyIdx;
            if (point_idx) point_idx[(size_t)v * P + s] = has ? (int)idx[s] : -1;
            if (voxels) reinterpret_cast<float4 *>(voxels)[(size_t)v * P + s] = q[s];
            if (has) { acc[0] += q[s].x; acc[1] += q[s].y; acc[2] += q[s].z; acc[3] += q[s].w; }
            cnt += has ? 1 : 0;
        }
        int cx, cy, cz;
        cell_of(p, q[0].x, q[0].y, q[0].z, &cx, &cy, &cz);
        reinterpret_cast<int4 *>(coords)[v] = make_int4(b, cz, cy, cx);
    } else
    for (int s = 0; s < P; ++s) {
        const unsigned int idx = list[s];
        const bool has = idx != kEmptyIdx;
        if (point_idx) point_idx[(size_t)v * P + s] = has ? (int)idx : -1;
        if (C == 4) {
            float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
            if (has) q = __ldg(reinterpret_cast<const float4 *>(points) + idx);
            if (voxels) reinterpret_cast<float4 *>(voxels)[(size_t)v * P + s] = q;
            if (has) { acc[0] += q.x; acc[1] += q.y; acc[2] += q.z; acc[3] += q.w; }
            if (s == 0) {
                int cx, cy, cz;
                cell_of(p, q.x, q.y, q.z, &cx, &cy, &cz);
                reinterpret_cast<int4 *>(coords)[v] = make_int4(b, cz, cy, cx);
            }
        } else {
            const float *q = points + (size_t)idx * C;
            for (int c = 0; c < C; ++c) {
                const float val = has ? __ldg(q + c) : 0.f;
                if (voxels) voxels[((size_t)v * P + s) * C + c] = val;
            }
            if (s == 0) {
                int cx, cy, cz;
                cell_of(p, __ldg(q), __ldg(q + 1), __ldg(q + 2), &cx, &cy, &cz);
                reinterpret_cast<int4 *>(coords)[v] = make_int4(b, cz, cy, cx);
            }
        }
        cnt += has ? 1 : 0;
    }
    num_points[v] = cnt;
    if (mean) {
        TMean *m = mean + (size_t)v * mean_stride;
        const float denom = (float)cnt;
        if (C == 4) {
#pragma unroll
            for (int c = 0; c < 4; ++c) m[c] = from_float<TMean>(__fdiv_rn(acc[c], denom));
        } else {
            for (int c = 0; c < C; ++c) {
                float s_ = 0.f;
                for (int s = 0; s < cnt; ++s) s_ += __ldg(points + (size_t)list[s] * C + c);
                m[c] = from_float<TMean>(__fdiv_rn(s_, denom));
            }
        }
        for (int c = C; c < mean_stride; ++c) m[c] = from_float<TMean>(0.f);
    }
}

// Coordinates and per-frame offsets of the output rows as soon as the voxel ranks exist: they are all the rulebook
// builds need, so those can start while the points are still being assigned and gathered.  Row v of frame b is the
// voxel of global rank start[b] + (v - out_base[b]); its cell is the cell of its owner (smallest) point.
__global__ void __launch_bounds__(128)
vox_write_coords(const float *__restrict__ points, const int *__restrict__ frame_start, const int *__restrict__ owner_of_rank,
                 VoxParams p, int *__restrict__ coords, int *__restrict__ voxel_offsets)
{
    __shared__ FrameInfo f;
    load_frame_info(f, frame_start, owner_of_rank, p.batch, p.max_voxels, p.overflow_break);
    if (blockIdx.x == 0)
        for (int t = threadIdx.x; t <= p.batch; t += blockDim.x) voxel_offsets[t] = f.out_base[t];
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= f.out_base[p.batch]) return;
    int b = frame_of(f.out_base, p.batch, v);
    while (b + 1 < p.batch && f.out_base[b + 1] <= v) ++b;
    const int i = owner_of_rank[f.start[b] + (v - f.out_base[b])];
    const float *q = points + (size_t)i * p.n_feat;
    int cx, cy, cz;
    cell_of(p, __ldg(q), __ldg(q + 1), __ldg(q + 2), &cx, &cy, &cz);
    reinterpret_cast<int4 *>(coords)[v] = make_int4(b, cz, cy, cx);
}

template <typename TMean>
__global__ void __launch_bounds__(256)
vfe_mean_kernel(const float *__restrict__ voxels, const int *__restrict__ num_points, int n, int P, int C,
                TMean *__restrict__ mean, int mean_stride)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n * mean_stride) return;
    const int v = t / mean_stride, c = t % mean_stride;
    float s = 0.f;
    if (c < C) {
        for (int k = 0; k < P; ++k) s += __ldg(voxels + ((size_t)v * P + k) * C + c);
        s = __fdiv_rn(s, (float)num_points[v]);
    }
    mean[t] = from_float<TMean>(s);
}

struct VoxWorkspace {
    unsigned long long *slots;
    int *pt_slot, *slot_rank, *owner_of_rank, *block_sums, *frame_start;
    unsigned int *vox_pts, *ticket;
    size_t fill_bytes;  // [slots | vox_pts | ticket] is one contiguous region initialised to 0xFF
    uint32_t table_cap;
    int nblocks;
    size_t bytes;
};

static VoxWorkspace carve_vox(void *base, int n_points, int batch, int max_points, int max_voxels)
{
    VoxWorkspace w{};
    const size_t n = (size_t)(n_points > 0 ? n_points : 1);
    w.table_cap = next_pow2(n * 2 < 1024 ? 1024 : n * 2);
    w.nblocks = (int)((n + kScanBlock - 1) / kScanBlock);
    size_t cap_rows = (size_t)batch * (size_t)max_voxels;
    if (cap_rows > n) cap_rows = n;
    size_t off = 0;
    char *b = (char *)base;
    auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes, 256); return b ? (void *)(b + o) : (void *)nullptr; };
    w.slots = (unsigned long long *)take((size_t)w.table_cap * 8);
    w.vox_pts = (unsigned int *)take(cap_rows * (size_t)max_points * 4);
    w.ticket = (unsigned int *)take(4);
    w.fill_bytes = off;
    w.pt_slot = (int *)take(n * 4);
    w.slot_rank = (int *)take((size_t)w.table_cap * 4);
    w.owner_of_rank = (int *)take(n * 4);
    w.block_sums = (int *)take(((size_t)w.nblocks + 1) * 4);
    w.frame_start = (int *)take(((size_t)batch + 1) * 4);
    w.bytes = off;
    return w;
}

}  // namespace pcdb

using namespace pcdb;

extern "C" size_t pcdb_voxelize_workspace_bytes(int n_points, int batch, int max_points, int max_voxels)
{
    return carve_vox(nullptr, n_points, batch, max_points, max_voxels).bytes;
}

// phase: 1 = sites (hash, ranks, coordinates, offsets), 2 = points (assignment, gather, mean), 3 = both
static int voxelize_impl(int phase, const float *points, int n_points, int n_feat, const int32_t *frame_offsets, int batch,
                         const float *voxel_size_xyz, const float *range_xyzxyz, const int32_t *grid_xyz,
                         int max_points, int max_voxels, int overflow_break,
                         float *voxels, int32_t *coords, int32_t *num_points, void *mean, int mean_dtype,
                         int mean_stride, int32_t *point_idx, int32_t *voxel_offsets,
                         void *workspace, size_t workspace_bytes, void *stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    if (n_points < 0 || n_feat < 3 || batch < 1 || batch > kMaxBatch || max_points < 1 || max_voxels < 1 ||
        !frame_offsets || !coords || !num_points || !voxel_offsets || (mean && mean_stride < n_feat)) {
        set_last_error("pcdb_voxelize: invalid argument (n_points=%d n_feat=%d batch=%d max_points=%d max_voxels=%d)",
                       n_points, n_feat, batch, max_points, max_voxels);
        return kInvalidArgument;
    }
    const uint64_t cells = (uint64_t)batch * (uint64_t)grid_xyz[0] * (uint64_t)grid_xyz[1] * (uint64_t)grid_xyz[2];
    if (cells >= 0xFFFFFFFFull) {
        set_last_error("pcdb_voxelize: batch*grid = %llu cells exceeds the 32-bit hash key; split the batch",
                       (unsigned long long)cells);
        return kKeyOverflow;
    }
    if (n_points == 0) {
        if (phase & 1) cudaMemsetAsync(voxel_offsets, 0, sizeof(int32_t) * (batch + 1), stream);
        return check_launch("pcdb_voxelize(memset)");
    }
    VoxWorkspace w = carve_vox(workspace, n_points, batch, max_points, max_voxels);
    if (!workspace || workspace_bytes < w.bytes) {
        set_last_error("pcdb_voxelize: workspace %zu < required %zu bytes", workspace_bytes, w.bytes);
        return kWorkspaceTooSmall;
    }
    VoxParams p;
    for (int d = 0; d < 3; ++d) { p.lo[d] = range_xyzxyz[d]; p.vs[d] = voxel_size_xyz[d]; p.grid[d] = grid_xyz[d]; }
    p.n_points = n_points; p.n_feat = n_feat; p.batch = batch; p.max_points = max_points;
    p.max_voxels = max_voxels; p.overflow_break = overflow_break;

    size_t cap_rows = (size_t)batch * (size_t)max_voxels;
    if (cap_rows > (size_t)n_points) cap_rows = (size_t)n_points;
    const int nb256 = (n_points + 255) / 256;
    const int nbv = (int)((cap_rows + 127) / 128);
    if (phase & 1) {
        // empty hash slots, empty point lists and the idle scan ticket are all 0xFF bytes: one memset
        cudaMemsetAsync(w.slots, 0xFF, w.fill_bytes, stream);
        vox_hash_insert<<<nb256, 256, 0, stream>>>(points, frame_offsets, p, w.slots, w.table_cap - 1, w.pt_slot);
        vox_count_owners<<<w.nblocks, kScanBlock, 0, stream>>>(w.slots, w.pt_slot, n_points, w.block_sums, w.ticket);
        vox_rank_owners<<<w.nblocks, kScanBlock, 0, stream>>>(w.slots, w.pt_slot, frame_offsets, batch, n_points,
                                                              w.block_sums, w.nblocks, w.slot_rank, w.owner_of_rank,
                                                              w.frame_start);
        if (phase == 1)      // (the gather of phase 2 writes the same coordinates and offsets again)
            vox_write_coords<<<nbv, 128, 0, stream>>>(points, w.frame_start, w.owner_of_rank, p, coords, voxel_offsets);
    }
    if (!(phase & 2)) return check_launch("pcdb_voxelize_sites");
    vox_assign_points<<<nb256, 256, 0, stream>>>(w.pt_slot, w.slot_rank, frame_offsets, w.frame_start,
                                                  w.owner_of_rank, p, w.vox_pts);
    if (mean && mean_dtype == PCDB_BF16)
        vox_gather<__nv_bfloat16><<<nbv, 128, 0, stream>>>(points, w.vox_pts, w.frame_start, w.owner_of_rank, p, voxels,
                                                            coords, num_points, (__nv_bfloat16 *)mean, mean_stride,
                                                            point_idx, voxel_offsets);
    else
        vox_gather<float><<<nbv, 128, 0, stream>>>(points, w.vox_pts, w.frame_start, w.owner_of_rank, p, voxels, coords,
                                                   num_points, (float *)mean, mean_stride, point_idx, voxel_offsets);
    return check_launch("pcdb_voxelize");
}

#define PCDB_VOXELIZE_ARGS \
    const float *points, int n_points, int n_feat, const int32_t *frame_offsets, int batch, const float *voxel_size_xyz, \
    const float *range_xyzxyz, const int32_t *grid_xyz, int max_points, int max_voxels, int overflow_break, float *voxels, \
    int32_t *coords, int32_t *num_points, void *mean, int mean_dtype, int mean_stride, int32_t *point_idx, \
    int32_t *voxel_offsets, void *workspace, size_t workspace_bytes, void *stream
#define PCDB_VOXELIZE_PASS \
    points, n_points, n_feat, frame_offsets, batch, voxel_size_xyz, range_xyzxyz, grid_xyz, max_points, max_voxels, \
    overflow_break, voxels, coords, num_points, mean, mean_dtype, mean_stride, point_idx, voxel_offsets, workspace, \
    workspace_bytes, stream

extern "C" int pcdb_voxelize(PCDB_VOXELIZE_ARGS) { return voxelize_impl(3, PCDB_VOXELIZE_PASS); }
extern "C" int pcdb_voxelize_sites(PCDB_VOXELIZE_ARGS) { return voxelize_impl(1, PCDB_VOXELIZE_PASS); }
extern "C" int pcdb_voxelize_points(PCDB_VOXELIZE_ARGS) { return voxelize_impl(2, PCDB_VOXELIZE_PASS); }

extern "C" int pcdb_vfe_mean(const float *voxels, const int32_t *num_points, int n_voxels, int max_points,
                             int n_feat, void *mean, int mean_dtype, int mean_stride, void *stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    if (n_voxels < 0 || max_points < 1 || n_feat < 1 || mean_stride < n_feat || !mean) {
        set_last_error("pcdb_vfe_mean: invalid argument");
        return kInvalidArgument;
    }
    if (n_voxels == 0) return kOk;
    const long long total = (long long)n_voxels * mean_stride;
    const int nb = (int)((total + 255) / 256);
    if (mean_dtype == PCDB_BF16)
        vfe_mean_kernel<__nv_bfloat16><<<nb, 256, 0, stream>>>(voxels, num_points, n_voxels, max_points, n_feat,
                                                               (__nv_bfloat16 *)mean, mean_stride);
    else
        vfe_mean_kernel<float><<<nb, 256, 0, stream>>>(voxels, num_points, n_voxels, max_points, n_feat,
                                                       (float *)mean, mean_stride);
    return check_launch("pcdb_vfe_mean");
}
