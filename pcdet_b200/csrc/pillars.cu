// PointPillars pillar feature net + BEV scatter for sm_100a (BASELINE config 2).
//
// Replaces PillarFeatureNetOld2.forward (pcdet/models/vfe/vfe_utils.py:168-215: three decorations, a padding
// mask, one PFNLayer = Linear(C+6 -> F, no bias) + BatchNorm1d(eval) + ReLU + max over the pillar's point slots,
// vfe_utils.py:61-116) and PointPillarsScatter.forward (pcdet/models/rpn/pillar_scatter.py:23-55: a zero
// canvas per sample, boolean-mask indexing per sample, transposes) -- ~25 torch kernels and a (N, P, C+6)
// intermediate -- by ONE kernel: a pillar's <= P points are decorated in shared memory, every output channel is
// one thread, and the max goes straight to the (B, F, ny, nx) canvas and/or the (N, F) feature matrix.
//
// Reference semantics kept: the mean divides the sum over ALL P slots (zero padded) by num_points; padded slots
// are zeroed AFTER decoration, so they contribute relu(BN(0)) = relu(shift) to the max whenever a pillar has
// fewer than P points.
#include "common.cuh"
#include "../../include/pcdet_b200.h"

namespace pcdb {

constexpr int kPfnMaxIn = 16;        // C + 6 (+1 with_distance)
constexpr int kPfnMaxPts = 128;

struct PfnParams {
    float vx, vy, vz, x_off, y_off, z_off;
    int P, C, F, with_distance, n_in;
    int nz, ny, nx, batch;
};

// grid: ceil(n / kPillarsPerCta); block: kPillarsPerCta * F threads (F = 64: 4 pillars, 256 threads)
template <int F>
__global__ void __launch_bounds__(256)
pillar_vfe_kernel(const float *__restrict__ voxels, const int *__restrict__ num_points, const int4 *__restrict__ coords,
                  int n, const int *__restrict__ n_dev, PfnParams p, const float *__restrict__ weight,
                  const float *__restrict__ scale, const float *__restrict__ shift, float *__restrict__ out_features,
                  float *__restrict__ canvas)
{
    constexpr int kPillars = 256 / F;
    __shared__ float s_feat[kPillars][kPfnMaxPts][kPfnMaxIn + 1];
    __shared__ float s_mean[kPillars][3];
    if (n_dev) { const int m = __ldg(n_dev); n = m < n ? m : n; }
    const int sub = threadIdx.x / F, c = threadIdx.x % F;
    const int v = blockIdx.x * kPillars + sub;
    const bool active = v < n;
    const int cnt = active ? min(__ldg(num_points + v), p.P) : 0;
    int4 co = make_int4(0, 0, 0, 0);
    if (active) co = __ldg(coords + v);
    // raw point features of ALL P slots (the reference sums the zero padding too)
    const float *src = voxels + (size_t)(active ? v : 0) * p.P * p.C;
    for (int t = c; t < p.P * p.C; t += F) s_feat[sub][t / p.C][t % p.C] = active ? __ldg(src + t) : 0.f;
    __syncthreads();
    if (c < 3) {
        float sum = 0.f;
        for (int i = 0; i < p.P; ++i) sum += s_feat[sub][i][c];
        s_mean[sub][c] = sum / (float)max(cnt, 1);
    }
    __syncthreads();
    // decorations (vfe_utils.py:179-196) and the padding mask (:198-203)
    const float cxv = (float)co.w * p.vx + p.x_off, cyv = (float)co.z * p.vy + p.y_off, czv = (float)co.y * p.vz + p.z_off;
    for (int i = c; i < p.P; i += F) {
        float *f = s_feat[sub][i];
        const float x = f[0], y = f[1], z = f[2];
        const bool real = i < cnt;
        int o = p.C;
        f[o++] = real ? x - s_mean[sub][0] : 0.f;
        f[o++] = real ? y - s_mean[sub][1] : 0.f;
        f[o++] = real ? z - s_mean[sub][2] : 0.f;
        f[o++] = real ? x - cxv : 0.f;
        f[o++] = real ? y - cyv : 0.f;
        f[o++] = real ? z - czv : 0.f;
        if (p.with_distance) f[o++] = real ? sqrtf(x * x + y * y + z * z) : 0.f;
        if (!real) for (int j = 0; j < p.C; ++j) f[j] = 0.f;
    }
    __syncthreads();
    if (!active) return;
    // PFNLayer: y = relu(scale * (W f) + shift), max over the P slots (vfe_utils.py:102-116)
    float w[kPfnMaxIn];
#pragma unroll
    for (int j = 0; j < kPfnMaxIn; ++j) w[j] = j < p.n_in ? __ldg(weight + (size_t)c * p.n_in + j) : 0.f;
    const float sc = scale ? __ldg(scale + c) : 1.f, sh = shift ? __ldg(shift + c) : 0.f;
    float best = cnt < p.P ? fmaxf(sh, 0.f) : 0.f;      // a padded slot is relu(BN(0)); relu output is >= 0 anyway
    for (int i = 0; i < cnt; ++i) {
        const float *f = s_feat[sub][i];
        float acc = 0.f;
#pragma unroll
        for (int j = 0; j < kPfnMaxIn; ++j) acc = fmaf(w[j], j < p.n_in ? f[j] : 0.f, acc);
        best = fmaxf(best, fmaxf(fmaf(acc, sc, sh), 0.f));
    }
    if (out_features) out_features[(size_t)v * F + c] = best;
    if (canvas && co.x >= 0 && co.x < p.batch) {
        // pillar_scatter.py:41: index = z * nz + y * nx + x into a (F, nz*ny*nx) canvas per sample
        const long long cell = (long long)co.y * p.nz + (long long)co.z * p.nx + co.w;
        if (cell >= 0 && cell < (long long)p.nz * p.ny * p.nx)
            canvas[((size_t)co.x * F + c) * ((size_t)p.nz * p.ny * p.nx) + cell] = best;
    }
}

}  // namespace pcdb

using namespace pcdb;

extern "C" int pcdb_pillar_vfe(const float *voxels, const int32_t *num_points, const int32_t *coords, int n,
                               const int32_t *n_dev, int max_points, int n_feat, const float *voxel_size_xyz,
                               const float *center_offset_xyz, int with_distance, const float *weight, int n_filters,
                               const float *scale, const float *shift, float *out_features, float *canvas, int batch,
                               const int32_t *canvas_shape_zyx, void *stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    const int n_in = n_feat + 6 + (with_distance ? 1 : 0);
    if (n < 0 || !voxels || !num_points || !coords || !weight || !voxel_size_xyz || !center_offset_xyz || n_feat < 3 ||
        max_points < 1 || max_points > kPfnMaxPts || n_in > kPfnMaxIn || (!out_features && !canvas) ||
        (canvas && (!canvas_shape_zyx || batch < 1))) {
        set_last_error("pcdb_pillar_vfe: invalid argument (n=%d max_points=%d n_feat=%d)", n, max_points, n_feat);
        return kInvalidArgument;
    }
    if (n_filters != 64) {
        set_last_error("pcdb_pillar_vfe: one PFN layer of 64 filters is built (got %d)", n_filters);
        return kUnsupported;
    }
    PfnParams p;
    p.vx = voxel_size_xyz[0]; p.vy = voxel_size_xyz[1]; p.vz = voxel_size_xyz[2];
    // vfe_utils.py:162-164: offset = voxel / 2 + range_min is evaluated by the caller in double (as Python does)
    // and enters the fp32 arithmetic as a scalar
    p.x_off = center_offset_xyz[0]; p.y_off = center_offset_xyz[1]; p.z_off = center_offset_xyz[2];
    p.P = max_points; p.C = n_feat; p.F = n_filters; p.with_distance = with_distance ? 1 : 0; p.n_in = n_in;
    p.nz = canvas ? canvas_shape_zyx[0] : 1; p.ny = canvas ? canvas_shape_zyx[1] : 1; p.nx = canvas ? canvas_shape_zyx[2] : 1;
    p.batch = batch;
    if (canvas) cudaMemsetAsync(canvas, 0, sizeof(float) * (size_t)batch * n_filters * p.nz * p.ny * p.nx, stream);
    if (n > 0)
        pillar_vfe_kernel<64><<<(n + 3) / 4, 256, 0, stream>>>(voxels, num_points, (const int4 *)coords, n, n_dev, p, weight,
                                                               scale, shift, out_features, canvas);
    return check_launch("pcdb_pillar_vfe");
}
