// Rotated-BEV overlap / IoU matrices and bitmask NMS for sm_100a.
//
// Replaces pcdet/ops/iou3d_nms/src/iou3d_nms_kernel.cu (boxes_overlap_kernel :224, boxes_iou_bev_kernel
// :237, nms_kernel :251, nms_normal_kernel :307) and the host side of iou3d_nms.cpp:79-177
// (cudaMalloc -> kernel -> 2 MB synchronous D2H -> serial CPU sweep -> H2D of the keep list).
//
// Geometry.  The reference intersects all 16 edge pairs, collects corners with a 1e-5 margin and
// bubble-sorts the polygon by atan2 (95 registers + 208 B of stack, ~60 atan2 per pair).  Here box
// A is expressed in box B's frame, where B is an axis-aligned rectangle, and clipped against B's
// four sides (Sutherland-Hodgman on at most 8 vertices); the shoelace formula gives the area.  Per-box
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

// Clip polygon (px,py,n) against the half-plane  sgn*coord <= lim  along axis X (or Y).
template <bool AXIS_X>
__device__ __forceinline__ int clip_axis(const float *px, const float *py, int n, float sgn, float lim,
                                         float *qx, float *qy)
{
    int m = 0;
    for (int i = 0; i < n; ++i) {
        const int j = (i + 1 == n) ? 0 : i + 1;
        const float ci = sgn * (AXIS_X ? px[i] : py[i]);
        const float cj = sgn * (AXIS_X ? px[j] : py[j]);
        const bool in_i = ci <= lim, in_j = cj <= lim;
        if (in_i) { qx[m] = px[i]; qy[m] = py[i]; ++m; }
        if (in_i != in_j) {
            const float t = (lim - ci) / (cj - ci);
            float nx = fmaf(t, px[j] - px[i], px[i]);
            float ny = fmaf(t, py[j] - py[i], py[i]);
            if (AXIS_X) nx = sgn * lim; else ny = sgn * lim;   // land exactly on the clip line
            qx[m] = nx; qy[m] = ny; ++m;
        }
    }
    return m;
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

// Area of A clipped by B's four sides (Sutherland-Hodgman, at most 8 vertices) in B's frame.
__device__ float clip_area(const BoxRec &a, const BoxRec &b, const RelPose &p)
{
    float px[8], py[8], qx[8], qy[8];
    const float lx[4] = {-a.hx, a.hx, a.hx, -a.hx}, ly[4] = {-a.hy, -a.hy, a.hy, a.hy};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        px[k] = lx[k] * p.cr + ly[k] * p.sr + p.ox;
        py[k] = -lx[k] * p.sr + ly[k] * p.cr + p.oy;
    }
    int n = 4;
    n = clip_axis<true>(px, py, n, 1.f, b.hx, qx, qy);
    if (n < 3) return 0.f;
    n = clip_axis<true>(qx, qy, n, -1.f, b.hx, px, py);
    if (n < 3) return 0.f;
    n = clip_axis<false>(px, py, n, 1.f, b.hy, qx, qy);
    if (n < 3) return 0.f;
    n = clip_axis<false>(qx, qy, n, -1.f, b.hy, px, py);
    if (n < 3) return 0.f;
    // shoelace about vertex 0 (keeps magnitudes small)
    float area = 0.f;
    for (int i = 1; i + 1 < n; ++i) {
        const float ux = px[i] - px[0], uy = py[i] - py[0];
        const float vx = px[i + 1] - px[0], vy = py[i + 1] - py[0];
        area += ux * vy - uy * vx;
    }
    return fabsf(area) * 0.5f;
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

// ---- NMS ---------------------------------------------------------------------------------------
struct NmsSet {
    int box_begin;      // first row in `boxes`
    int n;              // boxes in the set
    int col_blocks;     // ceil(n/64)
    int tile_begin;     // first CTA of the set in the mask grid
    long long mask_off; // first word of the set's mask
    long long diag_off; // first word of the set's transposed diagonal tiles (col_blocks * 64 words)
};
constexpr int kMaxSetsPerLaunch = 64;
// passed BY VALUE as a kernel parameter (1.5 KB): no staging copy, safe under CUDA-graph capture
struct NmsSetTable {
    int n_sets;
    int pad;
    NmsSet s[kMaxSetsPerLaunch];
};

__global__ void __launch_bounds__(256)
nms_prepare(const float *__restrict__ boxes, int total, BoxRec *__restrict__ recs)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < total) recs[i] = make_rec(boxes + (size_t)i * 5);
}

// One CTA (256 threads) per upper-triangular 64x64 tile, two phases so that the expensive polygon
// clipping runs on a DENSE list of candidate pairs instead of inside a divergent 4096-pair sweep:
//   phase 1: every thread runs the circumcircle rejection test on 16 pairs and appends the
//            survivors (typically < 1 %) to a shared-memory list (warp-aggregated append);
//   phase 2: the list is split evenly over the 256 threads: separating-axis test, clipping, IoU,
//            threshold, atomicOr of the result bit into the tile's 64 row words.
// Diagonal tiles also emit their transpose (per column: which earlier rows suppress it), which the
// sweep uses to resolve a 64-box chunk in a few warp-wide rounds instead of a 64-step serial loop.
template <bool NORMAL>
__global__ void __launch_bounds__(256)
nms_mask_kernel(const BoxRec *__restrict__ recs, const __grid_constant__ NmsSetTable sets, float thresh,
                unsigned long long *__restrict__ mask, unsigned long long *__restrict__ diag_t)
{
    __shared__ BoxRec s_col[64];
    __shared__ BoxRec s_row[64];
    __shared__ unsigned long long s_bits[64];
    __shared__ float4 s_colc[64];               // (cx, cy, circumradius, -) of the column boxes
    __shared__ unsigned short s_list[4096];     // pairs passing the circle test
    __shared__ unsigned short s_list2[4096];    // ... and the separating-axis test
    __shared__ int s_count, s_count2;
    int s = 0;
    while (s + 1 < sets.n_sets && sets.s[s + 1].tile_begin <= (int)blockIdx.x) ++s;
    const NmsSet st = sets.s[s];
    // decode the upper-triangular tile index: tiles of row rt are (rt, rt..cb-1); rows before rt hold
    // rt*cb - rt*(rt-1)/2 tiles.  Closed form + one correction step instead of a 64-iteration walk.
    const int t_lin = blockIdx.x - st.tile_begin, cb_ = st.col_blocks;
    const float disc = (2.f * cb_ + 1.f) * (2.f * cb_ + 1.f) - 8.f * (float)t_lin;
    int rt = (int)(((2.f * cb_ + 1.f) - sqrtf(fmaxf(disc, 0.f))) * 0.5f);
    rt = max(0, min(rt, cb_ - 1));
    while (rt > 0 && rt * cb_ - rt * (rt - 1) / 2 > t_lin) --rt;
    while ((rt + 1) * cb_ - (rt + 1) * rt / 2 <= t_lin) ++rt;
    const int ct = rt + (t_lin - (rt * cb_ - rt * (rt - 1) / 2));
    const int n = st.n;
    if (threadIdx.x < 64) {
        s_bits[threadIdx.x] = 0ull;
        const int c = ct * 64 + threadIdx.x;
        if (c < n) {
            const BoxRec b = recs[st.box_begin + c];
            s_col[threadIdx.x] = b;
            s_colc[threadIdx.x] = make_float4(b.cx, b.cy, b.rad, 0.f);
        }
    } else if (threadIdx.x < 128) {
        const int r = rt * 64 + threadIdx.x - 64;
        if (r < n) s_row[threadIdx.x - 64] = recs[st.box_begin + r];
    }
    if (threadIdx.x == 0) { s_count = 0; s_count2 = 0; }
    __syncthreads();
    const int row = threadIdx.x & 63, quarter = threadIdx.x >> 6;
    const int r = rt * 64 + row;
    // ---- phase 1: cheap rejection ------------------------------------------------------------------
    uint32_t cand = 0;
    if (r < n) {
        const BoxRec a = s_row[row];
        const int c_lo = quarter * 16;
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            const int cl = c_lo + j, c = ct * 64 + cl;
            if (c >= n || (rt == ct && cl <= row)) continue;
            bool keep_pair;
            if (NORMAL) {
                keep_pair = iou_axis(a, s_col[cl]) > thresh;      // axis-aligned IoU is cheap: decide right here
            } else {
                const float4 b = s_colc[cl];
                const float dx = a.cx - b.x, dy = a.cy - b.y, rr = a.rad + b.z;
                keep_pair = dx * dx + dy * dy < rr * rr;
            }
            if (keep_pair) cand |= 1u << j;
        }
    }
    if (NORMAL) {
        if (cand) atomicOr(&s_bits[row], (unsigned long long)cand << (quarter * 16));
    } else {
        // warp-aggregated append of (row, col) pairs
        const int lane = threadIdx.x & 31;
        const int cnt = __popc(cand);
        int incl = cnt;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += v;
        }
        int base = 0;
        if (lane == 31) base = atomicAdd(&s_count, incl);
        base = __shfl_sync(0xffffffffu, base, 31) + incl - cnt;
        for (uint32_t m = cand; m; m &= m - 1)
            s_list[base++] = (unsigned short)((row << 6) | (quarter * 16 + __ffs(m) - 1));
    }
    __syncthreads();
    // ---- phase 2: separating-axis test on the dense candidate list -> second, shorter list ----------
    // ---- phase 3: polygon clipping + threshold on what is left ---------------------------------------
    if (!NORMAL) {
        const int total = s_count;
        const int lane = threadIdx.x & 31;
        for (int i0 = 0; i0 < total; i0 += 256) {
            const int i = i0 + threadIdx.x;
            bool pass = false;
            int e = 0;
            if (i < total) {
                e = s_list[i];
                const BoxRec &a = s_row[e >> 6], &b = s_col[e & 63];
                const RelPose p = rel_pose(a, b);
                pass = sat_overlap(a, b, p);
                if (pass) {
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
                    if (ub < (thresh - 1e-4f) * (sum - ub)) {
                        pass = false;                                   // IoU certainly below the threshold
                    } else {
                        // the disc of radius min(ra, rb) - d/2 around the midpoint of the two centres lies
                        // inside both inscribed discs, hence inside both boxes
                        const float ra = fminf(a.hx, a.hy), rb = fminf(b.hx, b.hy);
                        const float rho = fminf(ra, rb) - 0.5f * sqrtf(p.ox * p.ox + p.oy * p.oy);
                        if (rho > 0.f) {
                            const float lb = 3.14159f * rho * rho;
                            if (lb > (thresh + 1e-4f) * (sum - lb)) {   // IoU certainly above the threshold
                                atomicOr(&s_bits[e >> 6], 1ull << (e & 63));
                                pass = false;
                            }
                        }
                    }
                }
            }
            const unsigned bal = __ballot_sync(0xffffffffu, pass);
            int base = 0;
            if (lane == 0 && bal) base = atomicAdd(&s_count2, __popc(bal));
            base = __shfl_sync(0xffffffffu, base, 0);
            if (pass) s_list2[base + __popc(bal & ((1u << lane) - 1u))] = (unsigned short)e;
        }
        __syncthreads();
        const int total2 = s_count2;
        for (int i = threadIdx.x; i < total2; i += 256) {
            const int e = s_list2[i], er = e >> 6, ec = e & 63;
            const BoxRec &a = s_row[er], &b = s_col[ec];
            const float so = clip_area(a, b, rel_pose(a, b));
            if (so / fmaxf(a.area + b.area - so, 1e-8f) > thresh) atomicOr(&s_bits[er], 1ull << ec);
        }
        __syncthreads();
    }
    if (threadIdx.x < 64 && r < n) mask[st.mask_off + (long long)r * st.col_blocks + ct] = s_bits[threadIdx.x];
    if (rt == ct && threadIdx.x >= 64 && threadIdx.x < 128) {
        // transpose: for column c, the set of rows of this chunk that suppress it
        const int c = threadIdx.x - 64;
        unsigned long long col = 0ull;
#pragma unroll 8
        for (int rr = 0; rr < 64; ++rr) col |= ((s_bits[rr] >> c) & 1ull) << rr;
        diag_t[st.diag_off + (long long)rt * 64 + c] = col;
    }
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
    const int n = st.n, cb = st.col_blocks;
    // dynamic shared memory carve-up
    unsigned long long *s_far = s_dyn;                              // [cb]   far-column removed bits
    unsigned long long *s_kept = s_far + cb;                        // [cb]   kept mask per chunk
    unsigned long long *s_ring = s_kept + cb;                       // [ring][near+1][64] row words
    unsigned long long *s_ringt = s_ring + kSweepRing * (kSweepNear + 1) * 64;   // [ring][64] transposed diagonal
    int *s_ready = reinterpret_cast<int *>(s_ringt + kSweepRing * 64);           // [cb] tiles of chunk c are in the ring
    int *s_fardone = s_ready + cb;                                  // [cb] far contributions of chunk c are in s_far
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
                    cp_async8(slot + d * 64 + lane + 32 * h, ok ? m + (long long)r * cb + c + d : m, ok ? 8u : 0u);
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
                            v[j] = i + j < nk ? __ldg(m + (long long)(c * 64 + rows[i + j]) * cb + col) : 0ull;
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
    unsigned long long *mask, *diag_t;
    size_t bytes;
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
    w.diag_t = (unsigned long long *)take(8 * (size_t)n_sets * cb * 64);
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

extern "C" size_t pcdb_nms_workspace_bytes(int n_sets, int max_boxes_per_set)
{
    return carve_nms(nullptr, n_sets > 0 ? n_sets : 1, max_boxes_per_set > 0 ? max_boxes_per_set : 1).bytes;
}

extern "C" int pcdb_nms(const float *boxes, const int32_t *set_offsets_host, int n_sets, float thresh, int normal,
                        int64_t *keep, int keep_stride, int32_t *num_keep, void *workspace, size_t workspace_bytes,
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
    static bool configured = false;
    if (!configured) {
        cudaFuncSetAttribute(nms_sweep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        configured = true;
    }
    if (smem > 200 * 1024) { set_last_error("pcdb_nms: %d boxes per set need too much shared memory", max_boxes); return kUnsupported; }
    long long mask_off = 0, diag_off = 0;
    for (int s0 = 0; s0 < n_sets; s0 += kMaxSetsPerLaunch) {
        NmsSetTable tab;
        tab.n_sets = n_sets - s0 < kMaxSetsPerLaunch ? n_sets - s0 : kMaxSetsPerLaunch;
        tab.pad = 0;
        int tiles = 0;
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
            tiles += st.col_blocks * (st.col_blocks + 1) / 2;
        }
        const int first = tab.s[0].box_begin;
        const int count = set_offsets_host[s0 + tab.n_sets] - set_offsets_host[s0];
        if (count > 0) {
            nms_prepare<<<(count + 255) / 256, 256, 0, stream>>>(b0 + (size_t)first * 5, count, w.recs + first);
            if (normal) nms_mask_kernel<true><<<tiles, 256, 0, stream>>>(w.recs, tab, thresh, w.mask, w.diag_t);
            else nms_mask_kernel<false><<<tiles, 256, 0, stream>>>(w.recs, tab, thresh, w.mask, w.diag_t);
        }
        nms_sweep_kernel<<<tab.n_sets, kSweepThreads, smem, stream>>>(w.mask, w.diag_t, tab,
                                                                      (long long *)keep + (size_t)s0 * keep_stride,
                                                                      keep_stride, num_keep + s0);
    }
    return check_launch("pcdb_nms");
}

extern "C" int pcdb_boxes3d_to_bev(const float *boxes3d, int n, float *boxes_bev, void *stream)
{
    if (n < 0 || !boxes_bev) { set_last_error("pcdb_boxes3d_to_bev: invalid argument"); return kInvalidArgument; }
    if (n == 0) return kOk;
    boxes3d_to_bev_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(boxes3d, n, boxes_bev);
    return check_launch("pcdb_boxes3d_to_bev");
}
