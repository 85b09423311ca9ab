// Geometry helpers shared by the rulebook builders (rulebook.cu: one map per call in the reference's row order;
// rulebook_chain.cu: every level of a backbone in four launches).
#pragma once
#include "common.cuh"

namespace pcdb {

struct ConvGeom {
    int in_shape[3], out_shape[3], ksize[3], stride[3], pad[3], dil[3];
    int K;
    int sshift[3];            // log2(stride) when the stride is a power of two, else -1 (generic division)
    int comb[3], combos;      // candidate offsets per dimension / per input row, see conv_candidate
    signed char dk[32][4];    // per kernel offset: (kz, ky, kx) * dilation -- no divisions in the kernels
    signed char cm[32][4];    // per candidate: its (mz, my, mx), see conv_candidate
};

// x mod stride / x div stride of dimension d (x >= 0): mask and shift for the power-of-two strides
__device__ __forceinline__ int mod_stride(const ConvGeom &g, int d, int x)
{
    return g.sshift[d] >= 0 ? x & (g.stride[d] - 1) : x % g.stride[d];
}
__device__ __forceinline__ int div_stride(const ConvGeom &g, int d, int x)
{
    return g.sshift[d] >= 0 ? x >> g.sshift[d] : x / g.stride[d];
}

__device__ __forceinline__ uint32_t lin_index(int b, int z, int y, int x, const int *shape)
{
    return (uint32_t)((b * shape[0] + z) * shape[1] + y) * (uint32_t)shape[2] + (uint32_t)x;
}

__device__ __forceinline__ int row_count(int n, const int *n_dev)
{
    if (!n_dev) return n;
    const int m = __ldg(n_dev);
    return m < n ? m : n;
}

// Row id of the active site (b, z, y, x), or -1.  With slot_oid the table is the one a strided build left
// behind (payload = first-touch key, row id in slot_oid[slot]); otherwise the payload is the row id.
__device__ __forceinline__ int site_row(const ConvGeom &g, const unsigned long long *__restrict__ slots, uint32_t mask,
                                        const int *__restrict__ slot_oid, int n, int b, int z, int y, int x)
{
    if (z < 0 || z >= g.in_shape[0] || y < 0 || y >= g.in_shape[1] || x < 0 || x >= g.in_shape[2]) return -1;
    uint32_t payload;
    const uint32_t s = table_find(slots, mask, lin_index(b, z, y, x, g.in_shape), &payload);
    if (s == 0xFFFFFFFFu) return -1;
    const int row = slot_oid ? __ldg(slot_oid + s) : (int)payload;
    return row < n ? row : -1;       // rows beyond the capacity of a strided build do not exist
}

// ---- strided convolution ---------------------------------------------------------------------
// Output site reached from input c through offset (kz,ky,kx): out = (in + pad - k*dil) / stride when
// divisible and in bounds.  Returns false otherwise.
__device__ __forceinline__ bool out_site(const ConvGeom &g, const int4 &c, int k, int *oz, int *oy, int *ox)
{
    const int tz = c.y + g.pad[0] - g.dk[k][0];
    const int ty = c.z + g.pad[1] - g.dk[k][1];
    const int tx = c.w + g.pad[2] - g.dk[k][2];
    if ((tz | ty | tx) < 0) return false;
    if (g.sshift[0] >= 0 && g.sshift[1] >= 0 && g.sshift[2] >= 0) {      // strides 1 / 2 / 4: shifts and masks
        if ((tz & (g.stride[0] - 1)) | (ty & (g.stride[1] - 1)) | (tx & (g.stride[2] - 1))) return false;
        *oz = tz >> g.sshift[0]; *oy = ty >> g.sshift[1]; *ox = tx >> g.sshift[2];
    } else {
        if (tz % g.stride[0] || ty % g.stride[1] || tx % g.stride[2]) return false;
        *oz = tz / g.stride[0]; *oy = ty / g.stride[1]; *ox = tx / g.stride[2];
    }
    return *oz < g.out_shape[0] && *oy < g.out_shape[1] && *ox < g.out_shape[2];
}

// An input coordinate reaches an output site only through the kernel offsets k with (i + pad - k) divisible by the
// stride (dilation 1): ceil(ksize/stride) candidates per dimension instead of ksize -- 8 instead of 27 per input
// row for the 3x3x3 / stride 2 convolutions of BackBone8x.  Candidate `cand` of a row -> (k, output site), or
// false.  With a dilation every offset stays a candidate (comb == ksize).
__device__ __forceinline__ bool conv_candidate(const ConvGeom &g, const int4 &c, int cand, int *k, int *oz, int *oy, int *ox)
{
    if (g.combos == g.K) { *k = cand; return out_site(g, c, cand, oz, oy, ox); }
    const int tz = c.y + g.pad[0], ty = c.z + g.pad[1], tx = c.w + g.pad[2];
    const int kz = mod_stride(g, 0, tz) + g.cm[cand][0] * g.stride[0], ky = mod_stride(g, 1, ty) + g.cm[cand][1] * g.stride[1],
              kx = mod_stride(g, 2, tx) + g.cm[cand][2] * g.stride[2];
    if (kz >= g.ksize[0] || ky >= g.ksize[1] || kx >= g.ksize[2] || kz > tz || ky > ty || kx > tx) return false;
    *oz = div_stride(g, 0, tz - kz); *oy = div_stride(g, 1, ty - ky); *ox = div_stride(g, 2, tx - kx);
    *k = (kz * g.ksize[1] + ky) * g.ksize[2] + kx;
    return *oz < g.out_shape[0] && *oy < g.out_shape[1] && *ox < g.out_shape[2];
}

// the candidate index under which conv_candidate reports offset k for input c (k must be one of its candidates)
__device__ __forceinline__ int candidate_of(const ConvGeom &g, const int4 &c, int k)
{
    if (g.combos == g.K) return k;
    const int mz = div_stride(g, 0, g.dk[k][0] - mod_stride(g, 0, c.y + g.pad[0])),      // dilation 1 here: dk = (kz, ky, kx)
              my = div_stride(g, 1, g.dk[k][1] - mod_stride(g, 1, c.z + g.pad[1])),
              mx = div_stride(g, 2, g.dk[k][2] - mod_stride(g, 2, c.w + g.pad[2]));
    return (mz * g.comb[1] + my) * g.comb[2] + mx;
}

inline bool fill_geom(ConvGeom &g, const int32_t *in_shape, const int32_t *out_shape, const int32_t *ksize,
                      const int32_t *stride, const int32_t *pad, const int32_t *dil)
{
    for (int d = 0; d < 3; ++d) {
        g.in_shape[d] = in_shape[d];
        g.out_shape[d] = out_shape ? out_shape[d] : in_shape[d];
        g.ksize[d] = ksize[d];
        g.stride[d] = stride ? stride[d] : 1;
        g.pad[d] = pad ? pad[d] : ksize[d] / 2;
        g.dil[d] = dil ? dil[d] : 1;
        if (g.ksize[d] < 1 || g.stride[d] < 1 || g.dil[d] < 1 || g.in_shape[d] < 1 || g.out_shape[d] < 1) return false;
    }
    g.K = g.ksize[0] * g.ksize[1] * g.ksize[2];
    if (g.K > 32) return false;   // owner bitmasks are 32 bits wide
    for (int d = 0; d < 3; ++d) {
        g.sshift[d] = -1;
        for (int sh = 0; sh < 8; ++sh) if (g.stride[d] == (1 << sh)) g.sshift[d] = sh;
    }
    const bool undilated = g.dil[0] == 1 && g.dil[1] == 1 && g.dil[2] == 1;
    for (int d = 0; d < 3; ++d) g.comb[d] = undilated ? (g.ksize[d] + g.stride[d] - 1) / g.stride[d] : g.ksize[d];
    g.combos = undilated ? g.comb[0] * g.comb[1] * g.comb[2] : g.K;
    for (int k = 0; k < g.K; ++k) {
        const int kx = k % g.ksize[2], ky = (k / g.ksize[2]) % g.ksize[1], kz = k / (g.ksize[2] * g.ksize[1]);
        const int v[3] = {kz * g.dil[0], ky * g.dil[1], kx * g.dil[2]};
        for (int d = 0; d < 3; ++d) {
            if (v[d] > 127) return false;
            g.dk[k][d] = (signed char)v[d];
        }
        g.dk[k][3] = 0;
    }
    for (int cand = 0; cand < 32; ++cand) {
        g.cm[cand][0] = (signed char)(cand / (g.comb[2] * g.comb[1]));
        g.cm[cand][1] = (signed char)((cand / g.comb[2]) % g.comb[1]);
        g.cm[cand][2] = (signed char)(cand % g.comb[2]);
        g.cm[cand][3] = 0;
    }
    return true;
}

// offset K-1-k is the negation of offset k iff the offsets are centred: (ksize-1)*dil == 2*pad in every dimension
// (odd kernel, dilation 1 under spconv's forced SubM padding k/2)
inline bool symmetric_offsets(const ConvGeom &g)
{
    for (int d = 0; d < 3; ++d)
        if ((g.ksize[d] - 1) * g.dil[d] != 2 * g.pad[d]) return false;
    return true;
}


}  // namespace pcdb
