// Rotated-BEV overlap / IoU matrices and bitmask NMS for sm_100a.
//
// Replaces pcdet/ops/iou3d_nms/src/iou3d_nms_kernel.cu (boxes_overlap_kernel :224, boxes_iou_bev_kernel
// :237, nms_kernel :251, nms_normal_kernel :307) and the host side of iou3d_nms.cpp:79-177
// (cudaMalloc -> kernel -> 2 MB synchronous D2H -> serial CPU sweep -> H2D of the keep list).
//
// Geometry.  The reference intersects all 16 edge pairs, collects corners with a 1e-5 margin and
// bubble-sorts the polygon by atan2 (95 registers + 208 B of stack, ~60 atan2 per pair).  Here box
// A is expressed in box B's frame, where B is an axis-aligned rectangle; the 4 edges of each box are clipped
// against the other box (Liang-Barsky) and Green's theorem over the 8 clipped segments gives the area.  Per-box
// trigonometry is hoisted into a 32-byte record computed once per box instead of once per pair.
// Both compute the area of the same convex polygon; results agree to fp32 rounding (~1e-6 in IoU),
// so keep/suppress decisions match except for pairs whose IoU sits within rounding of the threshold.
//
// NMS.  Kernel 1 builds the 64x64-tiled suppression bitmask for the upper triangle only (the sweep
// never reads the lower one; the reference computes it anyway), with a separating-axis pre-test
// so that only overlapping pairs pay for clipping.  Kernel 2 is the greedy sweep, one CTA per box
// set, entirely on the device: the mask never crosses PCIe and the keep list is produced where the
// detector needs it.
#include "common.cuh"
#include "../../include/pcdet_b200.h"

namespace pcdb {

struct BoxRec {          // 32 bytes
    float cx, cy, hx, hy;    // centre, half extents
    float c, s, area, rad;   // cos(ry), sin(ry), (x2-x1)*(y2-y1), circumradius (slightly inflated)
};

__device__ __forceinline__ BoxRec make_rec(const float *b)
{
    BoxRec r;
    r.cx = (b[0] + b[2]) * 0.5f;
    r.cy = (b[1] + b[3]) * 0.5f;
    r.hx = (b[2] - b[0]) * 0.5f;
    r.hy = (b[3] - b[1]) * 0.5f;
    sincosf(b[4], &r.s, &r.c);
    r.area = (b[2] - b[0]) * (b[3] - b[1]);
    r.rad = sqrtf(r.hx * r.hx + r.hy * r.hy) * 1.0001f + 1e-6f;
    return r;
}

// Relative pose of box A in box B's frame.  Corner convention of the reference
// (rotate_around_center, iou3d_nms_kernel.cu:100-104): global = (lx*c + ly*s + cx, -lx*s + ly*c + cy).
struct RelPose { float ox, oy, cr, sr; };

__device__ __forceinline__ RelPose rel_pose(const BoxRec &a, const BoxRec &b)
{
    RelPose p;
    const float dx = a.cx - b.cx, dy = a.cy - b.cy;
    p.ox = dx * b.c - dy * b.s;
    p.oy = dx * b.s + dy * b.c;
    p.cr = a.c * b.c + a.s * b.s;
    p.sr = a.s * b.c - a.c * b.s;
    return p;
}

// Separating-axis test on B's axes and on A's axes (exact: disjoint projections => area 0).
__device__ __forceinline__ bool sat_overlap(const BoxRec &a, const BoxRec &b, const RelPose &p)
{
    const float acr = fabsf(p.cr), asr = fabsf(p.sr);
    const float ex = acr * a.hx + asr * a.hy;   // A's half extent along B.x
    const float ey = asr * a.hx + acr * a.hy;   // ... along B.y
    if (fabsf(p.ox) >= ex + b.hx || fabsf(p.oy) >= ey + b.hy) return false;
    const float pax = p.ox * p.cr - p.oy * p.sr;   // centre offset seen from A's frame
    const float pay = p.ox * p.sr + p.oy * p.cr;
    const float fx = acr * b.hx + asr * b.hy;
    const float fy = asr * b.hx + acr * b.hy;
    return !(fabsf(pax) >= fx + a.hx || fabsf(pay) >= fy + a.hy);
}

// Area of the intersection of A and B, in B's frame.  The boundary of the intersection of two convex polygons is
// made of the parts of A's edges inside B and the parts of B's edges inside A, so its area is half the sum of the
// Green's-theorem terms (x0*y1 - x1*y0) of those <= 8 clipped segments: no vertex lists, no data-dependent loops,
// everything in registers (the Sutherland-Hodgman version kept four 8-entry arrays in local memory and took
// ~15k cycles per pair).
// Every edge of A is clipped against B's two slabs (Liang-Barsky).  The point X where A's edge line i crosses B's
// edge line j is ALSO where B's edge j enters or leaves the half-plane of A's edge i, and the clip interval of
// B's edge j is built from those very points: when two edges are nearly parallel X is ill-conditioned ALONG
// them, which moves the end of one segment and the start of the other by the same amount and changes nothing --
// clipping B's edges independently in A's frame double-counts or drops up to eps/angle of a shared edge.
// Exactly parallel edges: all or nothing, inclusive for A's edge and strict for B's, so that a shared edge is
// counted once.  Callers run sat_overlap first (touching boxes never get here).
// Corner convention of the reference (rotate_around_center, iou3d_nms_kernel.cu:100-104).
struct ClipState { float smin[4], smax[4], acc; };

// one edge of A: start (px, py), direction (dx, dy), all in B's frame; B = |x| <= hx, |y| <= hy
__device__ __forceinline__ void clip_edge(ClipState &c, float px, float py, float dx, float dy, float hx, float hy)
{
    float tmin = 0.f, tmax = 1.f;
    bool alive = true;
    // B's edges: 0 bottom (y = -hy, towards +x), 1 right (x = +hx, towards +y), 2 top (towards -x), 3 left (towards -y)
    if (fabsf(dx) < 1e-12f) {
        alive = fabsf(px) <= hx;
        if (!(-dy * (hx - px) > 0.f)) c.smax[1] = -1.f;
        if (!(-dy * (-hx - px) > 0.f)) c.smax[3] = -1.f;
    } else {
        const float inv = 1.0f / dx;
        const float tl = (-hx - px) * inv, tr = (hx - px) * inv;
        tmin = fmaxf(tmin, fminf(tl, tr));
        tmax = fminf(tmax, fmaxf(tl, tr));
        const float ih = 0.5f / hy;
        const float s1 = (fmaf(tr, dy, py) + hy) * ih, s3 = (hy - fmaf(tl, dy, py)) * ih;
        if (dx > 0.f) { c.smin[1] = fmaxf(c.smin[1], s1); c.smax[3] = fminf(c.smax[3], s3); }
        else          { c.smax[1] = fminf(c.smax[1], s1); c.smin[3] = fmaxf(c.smin[3], s3); }
    }
    if (fabsf(dy) < 1e-12f) {
        alive = alive && fabsf(py) <= hy;
        if (!(dx * (-hy - py) > 0.f)) c.smax[0] = -1.f;
        if (!(dx * (hy - py) > 0.f)) c.smax[2] = -1.f;
    } else {
        const float inv = 1.0f / dy;
        const float tb = (-hy - py) * inv, tt = (hy - py) * inv;
        tmin = fmaxf(tmin, fminf(tb, tt));
        tmax = fminf(tmax, fmaxf(tb, tt));
        const float ih = 0.5f / hx;
        const float s0 = (fmaf(tb, dx, px) + hx) * ih, s2 = (hx - fmaf(tt, dx, px)) * ih;
        if (dy < 0.f) { c.smin[0] = fmaxf(c.smin[0], s0); c.smax[2] = fminf(c.smax[2], s2); }
        else          { c.smax[0] = fminf(c.smax[0], s0); c.smin[2] = fmaxf(c.smin[2], s2); }
    }
    if (alive && tmin < tmax) {
        const float x0 = fmaf(tmin, dx, px), y0 = fmaf(tmin, dy, py), x1 = fmaf(tmax, dx, px), y1 = fmaf(tmax, dy, py);
        c.acc += x0 * y1 - x1 * y0;
    }
}

__device__ float clip_area(const BoxRec &a, const BoxRec &b, const RelPose &p)
{
    // A's local axes in B's frame: x -> (cr, -sr), y -> (sr, cr); A's centre at (ox, oy)
    const float ux = p.cr * a.hx, uy = -p.sr * a.hx;        // half edge vectors of A
    const float vx = p.sr * a.hy, vy = p.cr * a.hy;
    ClipState c;
#pragma unroll
    for (int j = 0; j < 4; ++j) { c.smin[j] = 0.f; c.smax[j] = 1.f; }
    c.acc = 0.f;
    // A's edges, counter-clockwise: -u-v -> +u-v -> +u+v -> -u+v
    clip_edge(c, p.ox - ux - vx, p.oy - uy - vy, 2.f * ux, 2.f * uy, b.hx, b.hy);
    clip_edge(c, p.ox + ux - vx, p.oy + uy - vy, 2.f * vx, 2.f * vy, b.hx, b.hy);
    clip_edge(c, p.ox + ux + vx, p.oy + uy + vy, -2.f * ux, -2.f * uy, b.hx, b.hy);
    clip_edge(c, p.ox - ux + vx, p.oy - uy + vy, -2.f * vx, -2.f * vy, b.hx, b.hy);
    // what is left of B's edges (axis aligned in this frame)
    float acc = c.acc;
    if (c.smin[0] < c.smax[0]) acc += 2.f * b.hx * b.hy * (c.smax[0] - c.smin[0]);      // bottom: y = -hy, x0*y1 - x1*y0 = hy*(x1 - x0)
    if (c.smin[1] < c.smax[1]) acc += 2.f * b.hx * b.hy * (c.smax[1] - c.smin[1]);      // right:  x = +hx, hx*(y1 - y0)
    if (c.smin[2] < c.smax[2]) acc += 2.f * b.hx * b.hy * (c.smax[2] - c.smin[2]);      // top:    y = +hy, -hy*(x1 - x0), x decreasing
    if (c.smin[3] < c.smax[3]) acc += 2.f * b.hx * b.hy * (c.smax[3] - c.smin[3]);      // left:   x = -hx, hx*(y0 - y1), y decreasing
    return fmaxf(0.5f * acc, 0.f);
}

__device__ __forceinline__ float rect_overlap(const BoxRec &a, const BoxRec &b)
{
    const RelPose p = rel_pose(a, b);
    if (!sat_overlap(a, b, p)) return 0.f;
    return clip_area(a, b, p);
}

__device__ __forceinline__ float iou_rot(const BoxRec &a, const BoxRec &b)
{
    const float so = rect_overlap(a, b);
    return so / fmaxf(a.area + b.area - so, 1e-8f);
}

// axis-aligned IoU of nms_normal (iou3d_nms_kernel.cu:296-305) on the raw (x1,y1,x2,y2)
__device__ __forceinline__ float iou_axis(const BoxRec &a, const BoxRec &b)
{
    const float left = fmaxf(a.cx - a.hx, b.cx - b.hx), right = fminf(a.cx + a.hx, b.cx + b.hx);
    const float top = fmaxf(a.cy - a.hy, b.cy - b.hy), bottom = fminf(a.cy + a.hy, b.cy + b.hy);
    const float inter = fmaxf(right - left, 0.f) * fmaxf(bottom - top, 0.f);
    return inter / fmaxf(a.area + b.area - inter, 1e-8f);
}

// ---- N x M matrices --------------------------------------------------------------------------
template <bool IOU>
__global__ void __launch_bounds__(256)
pair_matrix_kernel(const float *__restrict__ boxes_a, int na, const float *__restrict__ boxes_b, int nb,
                   float *__restrict__ ans)
{
    __shared__ BoxRec sa[16], sb[16];
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const int a0 = blockIdx.y * 16, b0 = blockIdx.x * 16;
    if (threadIdx.x < 16) {
        if (a0 + threadIdx.x < na) sa[threadIdx.x] = make_rec(boxes_a + (size_t)(a0 + threadIdx.x) * 5);
    } else if (threadIdx.x < 32) {
        const int t = threadIdx.x - 16;
        if (b0 + t < nb) sb[t] = make_rec(boxes_b + (size_t)(b0 + t) * 5);
    }
    __syncthreads();
    const int ia = a0 + ty, ib = b0 + tx;
    if (ia >= na || ib >= nb) return;
    const float so = rect_overlap(sa[ty], sb[tx]);
    ans[(size_t)ia * nb + ib] = IOU ? so / fmaxf(sa[ty].area + sb[tx].area - so, 1e-8f) : so;
}

// 3-D IoU of LiDAR boxes [x, y, z, w, l, h, ry] (z = bottom face): rotated BEV overlap x overlap of the height intervals over
// the union of the volumes -- iou3d_nms_utils.boxes_iou3d_gpu (pcdet/ops/iou3d_nms/iou3d_nms_utils.py:27-59: two BEV
// conversions, the overlap kernel and ten elementwise torch kernels) in one launch, same operation order.
__device__ __forceinline__ void lidar_to_bev5(const float *b, float *v)
{
    const float hw = b[3] / 2.f, hl = b[4] / 2.f;          // box_utils.py:237-250: x -+ w/2, y -+ l/2
    v[0] = b[0] - hw; v[1] = b[1] - hl; v[2] = b[0] + hw; v[3] = b[1] + hl; v[4] = b[6];
}

__global__ void __launch_bounds__(256)
iou3d_matrix_kernel(const float *__restrict__ boxes_a, int na, const float *__restrict__ boxes_b, int nb, float *__restrict__ ans)
{
    __shared__ BoxRec sa[16], sb[16];
    __shared__ float za[16][3], zb[16][3];      // bottom, top, volume
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const int a0 = blockIdx.y * 16, b0 = blockIdx.x * 16;
    if (threadIdx.x < 32) {
        const bool is_a = threadIdx.x < 16;
        const int t = threadIdx.x & 15, i = (is_a ? a0 : b0) + t;
        if (i < (is_a ? na : nb)) {
            const float *b = (is_a ? boxes_a : boxes_b) + (size_t)i * 7;
            float v[5];
            lidar_to_bev5(b, v);
            (is_a ? sa : sb)[t] = make_rec(v);
            float *z = is_a ? za[t] : zb[t];
            z[0] = b[2]; z[1] = b[2] + b[5]; z[2] = b[3] * b[4] * b[5];
        }
    }
    __syncthreads();
    const int ia = a0 + ty, ib = b0 + tx;
    if (ia >= na || ib >= nb) return;
    const float bev = rect_overlap(sa[ty], sb[tx]);
    const float h = fmaxf(fminf(za[ty][1], zb[tx][1]) - fmaxf(za[ty][0], zb[tx][0]), 0.f);
    const float inter = bev * h;
    ans[(size_t)ia * nb + ib] = inter / fmaxf(za[ty][2] + zb[tx][2] - inter, 1e-6f);
}

// ---- NMS ---------------------------------------------------------------------------------------
struct NmsSet {
    int box_begin;      // first row in `boxes`
    int n;              // boxes in the set
    int col_blocks;     // ceil(n/64)
    int tile_begin;     // first row block (64 boxes) of the set among the row blocks of all sets of the launch
    long long mask_off; // first word of the set's mask
    long long diag_off; // first word of the set's transposed diagonal tiles (col_blocks * 64 words)
};
constexpr int kMaxSetsPerLaunch = 64;
// passed BY VALUE as a kernel parameter (1.5 KB): no staging copy, safe under CUDA-graph capture
struct NmsSetTable {
    int n_sets;
    int pad;
    const int *counts;      // optional device array: boxes actually present in every set (<= its n), see pcdb_nms_counts
    NmsSet s[kMaxSetsPerLaunch];
};

// boxes of set s: its capacity n, or the device-side count when the caller gave one (mask rows keep the capacity stride)
__device__ __forceinline__ int set_size(const NmsSetTable &sets, int s)
{
    const int n = sets.s[s].n;
    if (!sets.counts) return n;
    const int c = __ldg(sets.counts + s);
    return c < 0 ? 0 : (c < n ? c : n);
}

__global__ void __launch_bounds__(256)
nms_prepare(const float *__restrict__ boxes, int total, BoxRec *__restrict__ recs)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < total) recs[i] = make_rec(boxes + (size_t)i * 5);
}

// Suppression mask of the rotated NMS in three kernels:
//   1. nms_mask_kernel: one CTA (256 threads) per STRIP of kStripTiles upper-triangular 64x64 tiles of one row
//      block runs the circumcircle rejection test on ALL pairs -- branch free, 8 instructions per pair -- and
//      appends the survivors (~0.6 % on detector-like boxes) to a global candidate list;
//   2. nms_resolve_kernel: one thread per candidate: separating-axis test, exact area bounds, polygon clipping,
//      threshold, atomicOr of the result bit into the (zeroed) mask;
//   3. nms_diag_kernel: the transpose of every diagonal tile (per column: which earlier rows of the chunk
//      suppress it), which the sweep uses to resolve a 64-box chunk in a few warp-wide rounds.
// Measured on B200 (4 x 4096 boxes): 72 us as ONE kernel that resolved its candidates inside the CTA -- a few
// threads clipping polygons while the other 200 waited at the barrier (ncu: 12 barrier-stall cycles per issued
// instruction, 37 % issue utilisation).  If the candidate list is full (degenerate inputs: everything overlaps
// everything) a CTA resolves its own candidates the old way, so the result never depends on the list's capacity.
// nms_normal (axis-aligned IoU) is decided inside the first kernel.
#ifndef PCDB_NMS_STRIP
#define PCDB_NMS_STRIP 8
#endif
#ifndef PCDB_NMS_MINB
#define PCDB_NMS_MINB 8
#endif
constexpr int kStripTiles = PCDB_NMS_STRIP;
constexpr int kListCap = 5120;          // shared-memory candidate entries per strip; more -> slow path

struct MaskSmem {
    BoxRec row[64];
    float4 colc[kStripTiles * 64];              // (cx, cy, circumradius, -) of the column boxes; far away beyond n
    unsigned long long bits[kStripTiles][64];   // nms_normal only
    unsigned short list[kListCap];              // pairs passing the circle test: row << 10 | tile << 6 | col (<= 16 tiles)
    int count, base;
};

struct NmsStrip { int set, rt, ct0, n_ct; };      // a strip whose candidates did not fit into the global list

__device__ __forceinline__ unsigned long long cand_encode(int set, int row, int col)
{
    return ((unsigned long long)set << 48) | ((unsigned long long)row << 24) | (unsigned long long)col;
}
constexpr unsigned long long kCandInvalid = ~0ull;

// IoU(a, b) > thresh, decided by the cheapest sufficient test
__device__ __forceinline__ bool pair_suppresses(const BoxRec &a, const BoxRec &b, float thresh)
{
    const RelPose p = rel_pose(a, b);
    if (!sat_overlap(a, b, p)) return false;
    // exact-math bounds on the intersection area decide most pairs without clipping:
    //   upper: the intersection lies inside B and inside A's bounding box in B's frame;
    //   lower: a disc contained in the discs inscribed in A and in B.
    // A margin of 1e-4 in IoU keeps these shortcuts away from pairs that rounding could flip.
    const float acr = fabsf(p.cr), asr = fabsf(p.sr);
    const float ex = acr * a.hx + asr * a.hy, ey = asr * a.hx + acr * a.hy;
    const float wx = fminf(p.ox + ex, b.hx) - fmaxf(p.ox - ex, -b.hx);
    const float wy = fminf(p.oy + ey, b.hy) - fmaxf(p.oy - ey, -b.hy);
    const float sum = a.area + b.area;
    const float ub = fminf(fmaxf(wx, 0.f) * fmaxf(wy, 0.f), fminf(a.area, b.area));
    if (ub < (thresh - 1e-4f) * (sum - ub)) return false;          // IoU certainly below the threshold
    // the disc of radius min(ra, rb) - d/2 around the midpoint of the two centres lies inside both inscribed
    // discs, hence inside both boxes
    const float ra = fminf(a.hx, a.hy), rb = fminf(b.hx, b.hy);
    const float rho = fminf(ra, rb) - 0.5f * sqrtf(p.ox * p.ox + p.oy * p.oy);
    if (rho > 0.f) {
        const float lb = 3.14159f * rho * rho;
        if (lb > (thresh + 1e-4f) * (sum - lb)) return true;       // IoU certainly above the threshold
    }
    const float so = clip_area(a, b, p);
    return so / fmaxf(sum - so, 1e-8f) > thresh;
}

// Hands the CTA's pending candidates to the global list.  Returns false when the list is full: the caller then
// files the whole strip for nms_resolve_kernel's slow path (the part it had reserved is marked invalid).
__device__ __forceinline__ bool mask_emit(MaskSmem &sm, int set, int row0, int col0, unsigned long long *__restrict__ cand,
                                          unsigned int *cand_count, unsigned int cand_cap)
{
    __syncthreads();                    // the list and its count are complete
    const int total = sm.count;
    if (threadIdx.x == 0) sm.base = total ? (int)min(atomicAdd(cand_count, (unsigned int)total), 0x7FFFFFFFu) : 0;
    __syncthreads();
    const unsigned int base = (unsigned int)sm.base;
    const bool fits = base + (unsigned int)total <= cand_cap;
    for (int i = threadIdx.x; i < total; i += 256) {
        const int e = sm.list[i];
        if (fits) cand[base + i] = cand_encode(set, row0 + (e >> 10), col0 + (e & 1023));
        else if (base + (unsigned int)i < cand_cap) cand[base + i] = kCandInvalid;
    }
    __syncthreads();
    if (threadIdx.x == 0) sm.count = 0;
    // the next append is separated from this reset by the caller's barrier (or nothing follows)
    return fits;
}

template <bool NORMAL>
__global__ void __launch_bounds__(256, NORMAL ? 4 : PCDB_NMS_MINB)
nms_mask_kernel(const BoxRec *__restrict__ recs, const __grid_constant__ NmsSetTable sets, float thresh,
                unsigned long long *__restrict__ mask, unsigned long long *__restrict__ cand, unsigned int *cand_count,
                unsigned int cand_cap, NmsStrip *__restrict__ ovf_strips, unsigned int *ovf_count)
{
    __shared__ MaskSmem sm;
    // grid: (strips of the longest row block, row blocks of all sets); tile_begin = the set's first row block
    int s = 0;
    while (s + 1 < sets.n_sets && sets.s[s + 1].tile_begin <= (int)blockIdx.y) ++s;
    const NmsSet st = sets.s[s];
    const int n = set_size(sets, s);
    const int cb_ = (n + 63) >> 6;
    const int rt = (int)blockIdx.y - st.tile_begin;
    const int ct0 = rt + (int)blockIdx.x * kStripTiles;
    if (ct0 >= cb_) return;
    const int n_ct = min(kStripTiles, cb_ - ct0);
    const BoxRec *cols = recs + st.box_begin + ct0 * 64;
    for (int t = threadIdx.x; t < kStripTiles * 64; t += 256) {
        // columns beyond the set sit infinitely far away: the circle test rejects them without a bounds check
        float4 cc = make_float4(1e30f, 1e30f, 0.f, 0.f);
        if (ct0 * 64 + t < n) {
            const float4 lo = __ldg(reinterpret_cast<const float4 *>(cols + t));             // cx, cy, hx, hy
            const float4 hi = __ldg(reinterpret_cast<const float4 *>(cols + t) + 1);         // c, s, area, rad
            cc = make_float4(lo.x, lo.y, hi.w, 0.f);
        }
        sm.colc[t] = cc;
        sm.bits[t >> 6][t & 63] = 0ull;
    }
    if (threadIdx.x < 64) {
        const int r = rt * 64 + threadIdx.x;
        if (r < n) sm.row[threadIdx.x] = recs[st.box_begin + r];
    }
    if (threadIdx.x == 0) sm.count = 0;
    __syncthreads();
    bool fits = true;
    const int row = threadIdx.x & 63, quarter = threadIdx.x >> 6;
    const int r = rt * 64 + row;
    const int c_lo = quarter * 16;
    // on the diagonal tile only columns right of the row count: bits j of this thread's 16 columns with c_lo + j > row
    const int below = row - c_lo;                   // columns c_lo .. c_lo + below are at or left of the diagonal
    const uint32_t diag_keep = below < 0 ? 0xFFFFu : (below >= 15 ? 0u : (0xFFFFu & ~((2u << below) - 1u)));
    float ax = 0.f, ay = 0.f, ar = 0.f;
    if (r < n) { ax = sm.row[row].cx; ay = sm.row[row].cy; ar = sm.row[row].rad; }
    for (int t = 0; t < n_ct; ++t) {
        const int ct = ct0 + t;
        uint32_t cand_bits = 0;
        if (NORMAL) {
            if (r < n) {
                const BoxRec a = sm.row[row];
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const int c = ct * 64 + c_lo + j;
                    if (c < n && iou_axis(a, cols[t * 64 + c_lo + j]) > thresh) cand_bits |= 1u << j;   // decided right here
                }
            }
        } else {
            // d^2 - (ra + rb)^2 is negative for a candidate: its sign bit is shifted straight into the bit set
            // (8 instructions per pair: LDS.128, 3 FADD, FMUL, 2 FFMA, SHF)
            const float4 *cc = sm.colc + t * 64 + c_lo;
#pragma unroll
            for (int j = 15; j >= 0; --j) {
                const float4 b = cc[j];
                const float dx = ax - b.x, dy = ay - b.y, rr = ar + b.z;
                const float v = fmaf(-rr, rr, fmaf(dy, dy, dx * dx));
                cand_bits = __funnelshift_l(__float_as_uint(v), cand_bits, 1);
            }
        }
        if (r >= n) cand_bits = 0;
        if (rt == ct) cand_bits &= diag_keep;
        if (NORMAL) {
            if (cand_bits) atomicOr(&sm.bits[t][row], (unsigned long long)cand_bits << (quarter * 16));
        } else {
            // candidates are rare (~3 per warp and tile): every lane that has some reserves its own slots; a lane
            // whose candidates do not fit any more drops them, and the count (> kListCap) then tells the CTA to
            // file the whole strip for the slow path -- no barrier between the tiles
            if (cand_bits) {
                int base = atomicAdd(&sm.count, __popc(cand_bits));
                if (base + __popc(cand_bits) <= kListCap)
                    for (uint32_t m = cand_bits; m; m &= m - 1)
                        sm.list[base++] = (unsigned short)((row << 10) | (t << 6) | (c_lo + __ffs(m) - 1));
            }
        }
    }
    if (NORMAL) {
        __syncthreads();
        for (int t = 0; t < n_ct; ++t)
            if (threadIdx.x < 64 && r < n) mask[st.mask_off + (long long)r * st.col_blocks + ct0 + t] = sm.bits[t][threadIdx.x];
    } else {
        __syncthreads();
        if (sm.count > kListCap) {
            fits = false;               // shared-memory list overflowed (nothing was emitted)
        } else {
            fits = mask_emit(sm, s, rt * 64, ct0 * 64, cand, cand_count, cand_cap);
        }
        if (!fits && threadIdx.x == 0) {
            // candidate list full: the whole strip is redone by nms_resolve_kernel (bits are OR-ed, doing some twice is fine)
            const unsigned int o = atomicAdd(ovf_count, 1u);
            ovf_strips[o] = NmsStrip{s, rt, ct0, n_ct};
        }
    }
}

// One thread per candidate pair (grid-stride).  Then the slow path: strips whose candidates did not fit into the
// list are redone pair by pair (circle test and, right away, the exact test) -- only degenerate inputs get here.
__global__ void __launch_bounds__(128)
nms_resolve_kernel(const BoxRec *__restrict__ recs, const __grid_constant__ NmsSetTable sets, float thresh,
                   const unsigned long long *__restrict__ cand, const unsigned int *__restrict__ cand_count,
                   unsigned int cand_cap, const NmsStrip *__restrict__ ovf_strips, const unsigned int *__restrict__ ovf_count,
                   unsigned long long *__restrict__ mask)
{
    const unsigned int total = min(*cand_count, cand_cap);
    for (unsigned int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const unsigned long long e = cand[i];
        if (e == kCandInvalid) continue;
        const NmsSet &st = sets.s[(int)(e >> 48)];
        const int row = (int)((e >> 24) & 0xFFFFFFu), col = (int)(e & 0xFFFFFFu);
        if (pair_suppresses(recs[st.box_begin + row], recs[st.box_begin + col], thresh))
            atomicOr(&mask[st.mask_off + (long long)row * st.col_blocks + (col >> 6)], 1ull << (col & 63));
    }
    const unsigned int n_ovf = *ovf_count;
    for (unsigned int o = blockIdx.x; o < n_ovf; o += gridDim.x) {
        const NmsStrip sp = ovf_strips[o];
        const NmsSet &st = sets.s[sp.set];
        const int n = set_size(sets, sp.set);
        for (int p = threadIdx.x; p < sp.n_ct * 4096; p += blockDim.x) {
            const int row = sp.rt * 64 + ((p >> 6) & 63), col = (sp.ct0 + (p >> 12)) * 64 + (p & 63);
            if (row >= n || col >= n || col <= row) continue;
            const BoxRec a = recs[st.box_begin + row], b = recs[st.box_begin + col];
            const float dx = a.cx - b.cx, dy = a.cy - b.cy, rr = a.rad + b.rad;
            if (fmaf(dy, dy, dx * dx) < rr * rr && pair_suppresses(a, b, thresh))
                atomicOr(&mask[st.mask_off + (long long)row * st.col_blocks + (col >> 6)], 1ull << (col & 63));
        }
    }
}

// grid: one CTA of 64 threads per 64-box chunk of every set (tile_begin is not used: chunks are found by walking
// the sets); transposes the diagonal tile of the finished mask.
__global__ void __launch_bounds__(64)
nms_diag_kernel(const unsigned long long *__restrict__ mask, const __grid_constant__ NmsSetTable sets,
                unsigned long long *__restrict__ diag_t)
{
    __shared__ unsigned long long s_bits[64];
    int s = 0, chunk = blockIdx.x;
    while (s < sets.n_sets && chunk >= sets.s[s].col_blocks) { chunk -= sets.s[s].col_blocks; ++s; }
    if (s >= sets.n_sets) return;
    const NmsSet &st = sets.s[s];
    const int n = set_size(sets, s);
    if (chunk * 64 >= n) return;
    const int r = chunk * 64 + threadIdx.x;
    s_bits[threadIdx.x] = r < n ? mask[st.mask_off + (long long)r * st.col_blocks + chunk] : 0ull;
    __syncthreads();
    const int c = threadIdx.x;
    unsigned long long col = 0ull;
#pragma unroll 8
    for (int rr = 0; rr < 64; ++rr) col |= ((s_bits[rr] >> c) & 1ull) << rr;
    diag_t[st.diag_off + (long long)chunk * 64 + c] = col;
}

// ---- greedy sweep -------------------------------------------------------------------------------
// One CTA per box set, warp-specialised so that the serial part of greedy NMS never waits for L2:
//   warp 0 (critical path): per 64-box chunk c it (1) ORs the suppression bits the last kNear chunks
//       contribute to column c, read from shared-memory tiles that were prefetched ahead of time,
//       (2) resolves the chunk in a few warp-wide rounds using the diagonal tile and its transpose,
//       (3) publishes the kept mask and appends the kept positions to the output;
//   warps 1..4 (prefetch): stream the diagonal tile, its transpose and the kNear tiles to the right
//       of every chunk from L2 into a shared-memory ring, running ahead of warp 0;
//   warps 5..20 (far columns): once chunk c is resolved they OR the mask rows of its kept boxes
//       into the "removed" words of columns > c + kNear, which warp 0 only needs kNear chunks later.
constexpr int kSweepNear = 6;
constexpr int kSweepRing = 16;
constexpr int kSweepPrefetchWarps = 4;
constexpr int kSweepFarWarps = 16;
constexpr int kSweepThreads = 32 * (1 + kSweepPrefetchWarps + kSweepFarWarps);

__device__ __forceinline__ int ld_volatile(const int *p) { return *((const volatile int *)p); }
__device__ __forceinline__ void cp_async8(void *smem_dst, const void *src, uint32_t src_bytes)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(src),
                 "r"(src_bytes) : "memory");
}

__global__ void __launch_bounds__(kSweepThreads)
nms_sweep_kernel(const unsigned long long *__restrict__ mask, const unsigned long long *__restrict__ diag_t,
                 const __grid_constant__ NmsSetTable sets, long long *__restrict__ keep, int keep_stride,
                 int *__restrict__ num_keep)
{
    extern __shared__ unsigned long long s_dyn[];
    const NmsSet st = sets.s[blockIdx.x];
    const int n = set_size(sets, blockIdx.x), cb = (n + 63) >> 6;   // boxes present / their 64-box chunks
    const int cbs = st.col_blocks;                                  // words per mask row (capacity)
    // dynamic shared memory carve-up
    unsigned long long *s_far = s_dyn;                              // [cbs]  far-column removed bits
    unsigned long long *s_kept = s_far + cbs;                       // [cbs]  kept mask per chunk
    unsigned long long *s_ring = s_kept + cbs;                      // [ring][near+1][64] row words
    unsigned long long *s_ringt = s_ring + kSweepRing * (kSweepNear + 1) * 64;   // [ring][64] transposed diagonal
    int *s_ready = reinterpret_cast<int *>(s_ringt + kSweepRing * 64);           // [cbs] tiles of chunk c are in the ring
    int *s_fardone = s_ready + cbs;                                 // [cbs] far contributions of chunk c are in s_far
    __shared__ int s_resolved, s_exit, s_count;
    __shared__ unsigned char s_rows[kSweepFarWarps * 64];

    const unsigned long long *m = mask + st.mask_off;
    const unsigned long long *dt = diag_t + st.diag_off;
    long long *kp = keep + (long long)blockIdx.x * keep_stride;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int j = threadIdx.x; j < cb; j += blockDim.x) { s_far[j] = 0ull; s_kept[j] = 0ull; s_ready[j] = 0; s_fardone[j] = 0; }
    if (threadIdx.x == 0) { s_resolved = 0; s_exit = 0; s_count = 0; }
    __syncthreads();

    if (warp == 0) {
        // =================================== critical path ==========================================
        int count = 0;
        for (int c = 0; c < cb; ++c) {
            while (ld_volatile(&s_ready[c]) == 0) {}
            if (c - kSweepNear - 1 >= 0)
                while (ld_volatile(&s_fardone[c - kSweepNear - 1]) == 0) {}
            __threadfence_block();
            // (1) removed bits of column c: far part + the near tiles of the previous kNear chunks
            unsigned long long acc = 0ull;
#pragma unroll
            for (int d = 1; d <= kSweepNear; ++d) {
                const int cp = c - d;
                if (cp < 0) break;
                const unsigned long long kept_p = s_kept[cp];
                const unsigned long long *tile = s_ring + ((cp % kSweepRing) * (kSweepNear + 1) + d) * 64;
                if ((kept_p >> lane) & 1ull) acc |= tile[lane];
                if ((kept_p >> (lane + 32)) & 1ull) acc |= tile[lane + 32];
            }
            uint32_t lo = __reduce_or_sync(0xffffffffu, (uint32_t)acc);
            uint32_t hi = __reduce_or_sync(0xffffffffu, (uint32_t)(acc >> 32));
            const unsigned long long remv = (((unsigned long long)hi << 32) | lo) | s_far[c];
            // (2) resolve the chunk: lane owns columns lane and lane+32
            const int rows = min(64, n - c * 64);
            unsigned long long und = ~remv;
            if (rows < 64) und &= (1ull << rows) - 1ull;
            unsigned long long kept = 0ull;
            const unsigned long long *tt = s_ringt + (c % kSweepRing) * 64;
            const unsigned long long cm0 = tt[lane], cm1 = tt[lane + 32];
            while (und) {
                // a column is dead if a kept row suppresses it, kept if no undecided row can still do so
                const bool u0 = (und >> lane) & 1ull, u1 = (und >> (lane + 32)) & 1ull;
                const bool dead0 = u0 && (cm0 & kept), dead1 = u1 && (cm1 & kept);
                const bool keep0 = u0 && !dead0 && !(cm0 & und), keep1 = u1 && !dead1 && !(cm1 & und);
                const unsigned long long nk = (unsigned long long)__ballot_sync(0xffffffffu, keep0) |
                                              ((unsigned long long)__ballot_sync(0xffffffffu, keep1) << 32);
                const unsigned long long nd = (unsigned long long)__ballot_sync(0xffffffffu, dead0) |
                                              ((unsigned long long)__ballot_sync(0xffffffffu, dead1) << 32);
                kept |= nk;
                und &= ~(nk | nd);
            }
            // (3) publish and emit
            if (lane == 0) s_kept[c] = kept;
            __threadfence_block();
            __syncwarp();
            if (lane == 0) *((volatile int *)&s_resolved) = c + 1;
            if ((kept >> lane) & 1ull) {
                const int pos = count + __popcll(kept & ((1ull << lane) - 1ull));
                if (pos < keep_stride) kp[pos] = (long long)c * 64 + lane;
            }
            if ((kept >> (lane + 32)) & 1ull) {
                const int pos = count + __popcll(kept & ((1ull << (lane + 32)) - 1ull));
                if (pos < keep_stride) kp[pos] = (long long)c * 64 + lane + 32;
            }
            count += __popcll(kept);
            if (count >= keep_stride) break;       // the caller only wants the first keep_stride boxes
        }
        if (lane == 0) { s_count = count; *((volatile int *)&s_exit) = 1; }
    } else if (warp <= kSweepPrefetchWarps) {
        // =================================== prefetch =============================================
        for (int c = warp - 1; c < cb; c += kSweepPrefetchWarps) {
            // slot c % ring is free once chunk c - ring is no longer a "near" neighbour of anything unresolved
            while (ld_volatile(&s_resolved) < c - kSweepRing + kSweepNear + 1 && !ld_volatile(&s_exit)) __nanosleep(64);
            if (ld_volatile(&s_exit)) break;
            unsigned long long *slot = s_ring + (c % kSweepRing) * (kSweepNear + 1) * 64;
            // 8-byte cp.async straight into the ring: all 16 copies of a lane are in flight at once
#pragma unroll
            for (int d = 0; d <= kSweepNear; ++d) {
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int r = c * 64 + lane + 32 * h;
                    const bool ok = r < n && c + d < cb;
                    cp_async8(slot + d * 64 + lane + 32 * h, ok ? m + (long long)r * cbs + c + d : m, ok ? 8u : 0u);
                }
            }
            cp_async8(s_ringt + (c % kSweepRing) * 64 + lane, dt + (long long)c * 64 + lane, 8u);
            cp_async8(s_ringt + (c % kSweepRing) * 64 + lane + 32, dt + (long long)c * 64 + lane + 32, 8u);
            asm volatile("cp.async.wait_all;" ::: "memory");
            __threadfence_block();
            __syncwarp();
            if (lane == 0) *((volatile int *)&s_ready[c]) = 1;
        }
    } else {
        // =================================== far columns ==========================================
        const int w = warp - 1 - kSweepPrefetchWarps;
        for (int c = w; c < cb; c += kSweepFarWarps) {
            while (ld_volatile(&s_resolved) <= c && !ld_volatile(&s_exit)) __nanosleep(32);
            if (ld_volatile(&s_resolved) <= c) break;     // early exit before this chunk was resolved
            __threadfence_block();
            const unsigned long long kept = s_kept[c];
            if (kept && c + kSweepNear + 1 < cb) {
                // compact the kept rows of the chunk, then issue the row loads eight at a time
                unsigned char *rows = s_rows + w * 64;
                __syncwarp();
                if ((kept >> lane) & 1ull) rows[__popcll(kept & ((1ull << lane) - 1ull))] = (unsigned char)lane;
                if ((kept >> (lane + 32)) & 1ull) rows[__popcll(kept & ((1ull << (lane + 32)) - 1ull))] = (unsigned char)(lane + 32);
                __syncwarp();
                const int nk = __popcll(kept);
                for (int col = c + kSweepNear + 1 + lane; col < cb; col += 32) {
                    unsigned long long acc = 0ull;
                    for (int i = 0; i < nk; i += 8) {
                        unsigned long long v[8];
#pragma unroll
                        for (int j = 0; j < 8; ++j)
                            v[j] = i + j < nk ? __ldg(m + (long long)(c * 64 + rows[i + j]) * cbs + col) : 0ull;
#pragma unroll
                        for (int j = 0; j < 8; ++j) acc |= v[j];
                    }
                    if (acc) atomicOr(&s_far[col], acc);
                }
            }
            __threadfence_block();
            __syncwarp();
            if (lane == 0) *((volatile int *)&s_fardone[c]) = 1;
        }
    }
    __syncthreads();
    const int total = s_count;
    if (threadIdx.x == 0) num_keep[blockIdx.x] = total < keep_stride ? total : keep_stride;
    for (int j = total + threadIdx.x; j < keep_stride; j += blockDim.x) kp[j] = -1;
}

__global__ void __launch_bounds__(256)
boxes3d_to_bev_kernel(const float *__restrict__ b3, int n, float *__restrict__ bev)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float *b = b3 + (size_t)i * 7;
    const float hw = b[3] / 2.f, hl = b[4] / 2.f;
    float *o = bev + (size_t)i * 5;
    o[0] = b[0] - hw; o[1] = b[1] - hl; o[2] = b[0] + hw; o[3] = b[1] + hl; o[4] = b[6];
}

struct NmsWorkspace {
    BoxRec *recs;
    unsigned long long *mask, *diag_t, *cand;
    unsigned int *cand_count;       // [cand_count, ovf_count] directly behind the mask: one memset clears all three
    unsigned int cand_cap;
    NmsStrip *ovf_strips;
    size_t mask_bytes, bytes;
};

static NmsWorkspace carve_nms(void *base, int n_sets, int max_boxes)
{
    NmsWorkspace w{};
    size_t off = 0;
    char *b = (char *)base;
    auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes, 256); return b ? (void *)(b + o) : (void *)nullptr; };
    const size_t cb = ((size_t)max_boxes + 63) / 64;
    w.recs = (BoxRec *)take(sizeof(BoxRec) * (size_t)n_sets * max_boxes);
    w.mask = (unsigned long long *)take(8 * (size_t)n_sets * max_boxes * cb);
    w.cand_count = (unsigned int *)take(8);
    w.mask_bytes = (size_t)((char *)w.cand_count - (char *)w.mask) + 8;
    w.diag_t = (unsigned long long *)take(8 * (size_t)n_sets * cb * 64);
    // candidate pairs of the rotated NMS: 64 per box is ~10x what detector-like boxes produce; a full list only
    // costs speed (the CTAs then resolve their candidates themselves)
    const size_t cap = (size_t)n_sets * max_boxes * 64;
    w.cand_cap = (unsigned int)(cap < 65536 ? 65536 : (cap > (1u << 28) ? (1u << 28) : cap));
    w.cand = (unsigned long long *)take(8 * (size_t)w.cand_cap);
    size_t strips = 0;
    for (size_t rt = 0; rt < cb; ++rt) strips += (cb - rt + kStripTiles - 1) / kStripTiles;
    w.ovf_strips = (NmsStrip *)take(sizeof(NmsStrip) * (size_t)n_sets * strips);
    w.bytes = off;
    return w;
}

}  // namespace pcdb

using namespace pcdb;

extern "C" int pcdb_boxes_overlap_bev(const float *boxes_a, int na, const float *boxes_b, int nb, float *ans, void *stream)
{
    if (na < 0 || nb < 0 || !ans) { set_last_error("pcdb_boxes_overlap_bev: invalid argument"); return kInvalidArgument; }
    if (na == 0 || nb == 0) return kOk;
    pair_matrix_kernel<false><<<dim3((nb + 15) / 16, (na + 15) / 16), 256, 0, (cudaStream_t)stream>>>(boxes_a, na, boxes_b, nb, ans);
    return check_launch("pcdb_boxes_overlap_bev");
}

extern "C" int pcdb_boxes_iou_bev(const float *boxes_a, int na, const float *boxes_b, int nb, float *ans, void *stream)
{
    if (na < 0 || nb < 0 || !ans) { set_last_error("pcdb_boxes_iou_bev: invalid argument"); return kInvalidArgument; }
    if (na == 0 || nb == 0) return kOk;
    pair_matrix_kernel<true><<<dim3((nb + 15) / 16, (na + 15) / 16), 256, 0, (cudaStream_t)stream>>>(boxes_a, na, boxes_b, nb, ans);
    return check_launch("pcdb_boxes_iou_bev");
}

extern "C" int pcdb_boxes_iou3d(const float *boxes_a, int na, const float *boxes_b, int nb, float *ans, void *stream)
{
    if (na < 0 || nb < 0 || !ans) { set_last_error("pcdb_boxes_iou3d: invalid argument"); return kInvalidArgument; }
    if (na == 0 || nb == 0) return kOk;
    iou3d_matrix_kernel<<<dim3((nb + 15) / 16, (na + 15) / 16), 256, 0, (cudaStream_t)stream>>>(boxes_a, na, boxes_b, nb, ans);
    return check_launch("pcdb_boxes_iou3d");
}

extern "C" size_t pcdb_nms_workspace_bytes(int n_sets, int max_boxes_per_set)
{
    return carve_nms(nullptr, n_sets > 0 ? n_sets : 1, max_boxes_per_set > 0 ? max_boxes_per_set : 1).bytes;
}

static int nms_impl(const float *boxes, const int32_t *set_offsets_host, const int32_t *set_counts, int n_sets, float thresh,
                    int normal, int64_t *keep, int keep_stride, int32_t *num_keep, void *workspace, size_t workspace_bytes,
                    void *stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    if (n_sets < 1 || !set_offsets_host || !keep || !num_keep || keep_stride < 1) {
        set_last_error("pcdb_nms: invalid argument (n_sets=%d keep_stride=%d)", n_sets, keep_stride);
        return kInvalidArgument;
    }
    int max_boxes = 0;
    for (int s = 0; s < n_sets; ++s) {
        const int n = set_offsets_host[s + 1] - set_offsets_host[s];
        if (n < 0) { set_last_error("pcdb_nms: set offsets must be non-decreasing"); return kInvalidArgument; }
        if (n > max_boxes) max_boxes = n;
    }
    if (max_boxes > 65536) { set_last_error("pcdb_nms: at most 65536 boxes per set (got %d)", max_boxes); return kUnsupported; }
    NmsWorkspace w = carve_nms(workspace, n_sets, max_boxes > 0 ? max_boxes : 1);
    if (!workspace || workspace_bytes < w.bytes) {
        set_last_error("pcdb_nms: workspace %zu < required %zu bytes", workspace_bytes, w.bytes);
        return kWorkspaceTooSmall;
    }
    const float *b0 = boxes + (size_t)set_offsets_host[0] * 5;
    const size_t cbmax = (size_t)((max_boxes + 63) / 64 + 1);
    const size_t smem = 8 * (2 * cbmax + (size_t)kSweepRing * (kSweepNear + 2) * 64) + 4 * 2 * cbmax + 64;
    // per device and context, cheap: set on every call (a process-wide "done once" flag breaks on a second GPU)
    if (cudaFuncSetAttribute(nms_sweep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) != cudaSuccess)
        return check_launch("pcdb_nms(cudaFuncSetAttribute)");
    if (smem > 200 * 1024) { set_last_error("pcdb_nms: %d boxes per set need too much shared memory", max_boxes); return kUnsupported; }
    long long mask_off = 0, diag_off = 0;
    for (int s0 = 0; s0 < n_sets; s0 += kMaxSetsPerLaunch) {
        NmsSetTable tab;
        tab.n_sets = n_sets - s0 < kMaxSetsPerLaunch ? n_sets - s0 : kMaxSetsPerLaunch;
        tab.pad = 0;
        tab.counts = set_counts ? set_counts + s0 : nullptr;
        int tiles = 0, max_cb = 0;
        for (int s = 0; s < tab.n_sets; ++s) {
            NmsSet &st = tab.s[s];
            st.box_begin = set_offsets_host[s0 + s] - set_offsets_host[0];
            st.n = set_offsets_host[s0 + s + 1] - set_offsets_host[s0 + s];
            st.col_blocks = (st.n + 63) / 64;
            st.mask_off = mask_off;
            st.diag_off = diag_off;
            st.tile_begin = tiles;
            mask_off += (long long)st.n * st.col_blocks;
            diag_off += (long long)st.col_blocks * 64;
            tiles += st.col_blocks;
            if (st.col_blocks > max_cb) max_cb = st.col_blocks;
        }
        const int first = tab.s[0].box_begin;
        const int count = set_offsets_host[s0 + tab.n_sets] - set_offsets_host[s0];
        if (tiles > 65535) { set_last_error("pcdb_nms: too many 64-box chunks in one launch group (%d)", tiles); return kUnsupported; }
        const dim3 mask_grid((max_cb + kStripTiles - 1) / kStripTiles > 0 ? (max_cb + kStripTiles - 1) / kStripTiles : 1, tiles > 0 ? tiles : 1);
        if (count > 0) {
            nms_prepare<<<(count + 255) / 256, 256, 0, stream>>>(b0 + (size_t)first * 5, count, w.recs + first);
            if (normal) {
                nms_mask_kernel<true><<<mask_grid, 256, 0, stream>>>(w.recs, tab, thresh, w.mask, w.cand, w.cand_count, w.cand_cap,
                                                                 w.ovf_strips, w.cand_count + 1);
            } else {
                if (s0 == 0) cudaMemsetAsync(w.mask, 0, w.mask_bytes, stream);         // bits are OR-ed in; counter = 0
                else cudaMemsetAsync(w.cand_count, 0, 8, stream);
                nms_mask_kernel<false><<<mask_grid, 256, 0, stream>>>(w.recs, tab, thresh, w.mask, w.cand, w.cand_count, w.cand_cap,
                                                                  w.ovf_strips, w.cand_count + 1);
                nms_resolve_kernel<<<kNumSMs * 8, 128, 0, stream>>>(w.recs, tab, thresh, w.cand, w.cand_count, w.cand_cap,
                                                                    w.ovf_strips, w.cand_count + 1, w.mask);
            }
            if (tiles > 0) nms_diag_kernel<<<tiles, 64, 0, stream>>>(w.mask, tab, w.diag_t);
        }
        nms_sweep_kernel<<<tab.n_sets, kSweepThreads, smem, stream>>>(w.mask, w.diag_t, tab,
                                                                      (long long *)keep + (size_t)s0 * keep_stride,
                                                                      keep_stride, num_keep + s0);
    }
    return check_launch("pcdb_nms");
}

extern "C" int pcdb_nms(const float *boxes, const int32_t *set_offsets_host, int n_sets, float thresh, int normal,
                        int64_t *keep, int keep_stride, int32_t *num_keep, void *workspace, size_t workspace_bytes,
                        void *stream)
{
    return nms_impl(boxes, set_offsets_host, nullptr, n_sets, thresh, normal, keep, keep_stride, num_keep, workspace,
                    workspace_bytes, stream);
}

extern "C" int pcdb_nms_counts(const float *boxes, const int32_t *set_offsets_host, const int32_t *set_counts, int n_sets,
                               float thresh, int normal, int64_t *keep, int keep_stride, int32_t *num_keep, void *workspace,
                               size_t workspace_bytes, void *stream)
{
    if (!set_counts) { set_last_error("pcdb_nms_counts: set_counts is NULL"); return kInvalidArgument; }
    return nms_impl(boxes, set_offsets_host, set_counts, n_sets, thresh, normal, keep, keep_stride, num_keep, workspace,
                    workspace_bytes, stream);
}

extern "C" int pcdb_boxes3d_to_bev(const float *boxes3d, int n, float *boxes_bev, void *stream)
{
    if (n < 0 || !boxes_bev) { set_last_error("pcdb_boxes3d_to_bev: invalid argument"); return kInvalidArgument; }
    if (n == 0) return kOk;
    boxes3d_to_bev_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(boxes3d, n, boxes_bev);
    return check_launch("pcdb_boxes3d_to_bev");
}
