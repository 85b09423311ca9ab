// Sparse convolution forward/backward on the fp32 FMA pipe (sm_100a).
//
// Replaces spconv v1.0 indiceConv<T> / indiceConvBackward<T>: the reference runs,
// per kernel offset, a gather kernel, a cuBLAS GEMM and a scatter-add kernel (~80 launches and one
// device->host sync per layer).  Here one launch covers the layer: every CTA owns a tile of 64
// OUTPUT rows, walks the kernel offsets, gathers the contributing input rows through the
// output-stationary neighbour map into shared memory and accumulates in registers, so there is no
// scatter, no atomic and a fixed summation order (offset ascending, channel ascending).
//
// This file is the exact path (PCDB_F32, <= 1e-4 of an fp32 reference) and the fallback for channel
// counts the tensor-core kernel does not take; with PCDB_BF16 storage it reproduces the tensor-core
// kernel's numerics (bf16 operands, fp32 accumulate) on CUDA cores, which the tests use to
// cross-check sparse_conv_tc.cu.
#include "common.cuh"
#include "../../include/pcdet_b200.h"

namespace pcdb {

constexpr int kTM = 64;   // output rows per CTA
constexpr int kTN = 64;   // output channels per CTA
constexpr int kKC = 16;   // input channels per shared-memory stage
constexpr int kLdA = kTM + 4;

template <typename T> struct Vec4;
template <> struct Vec4<float> {
    static __device__ __forceinline__ void load(const float *p, float *v)
    {
        const float4 q = __ldg(reinterpret_cast<const float4 *>(p));
        v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
    }
    static __device__ __forceinline__ void store(float *p, const float *v)
    {
        *reinterpret_cast<float4 *>(p) = make_float4(v[0], v[1], v[2], v[3]);
    }
};
template <> struct Vec4<__nv_bfloat16> {
    static __device__ __forceinline__ void load(const __nv_bfloat16 *p, float *v)
    {
        const uint2 q = __ldg(reinterpret_cast<const uint2 *>(p));
        const __nv_bfloat162 a = *reinterpret_cast<const __nv_bfloat162 *>(&q.x);
        const __nv_bfloat162 b = *reinterpret_cast<const __nv_bfloat162 *>(&q.y);
        v[0] = __low2float(a); v[1] = __high2float(a); v[2] = __low2float(b); v[3] = __high2float(b);
    }
    static __device__ __forceinline__ void store(__nv_bfloat16 *p, const float *v)
    {
        __nv_bfloat162 a = __floats2bfloat162_rn(v[0], v[1]);
        __nv_bfloat162 b = __floats2bfloat162_rn(v[2], v[3]);
        uint2 q;
        q.x = *reinterpret_cast<uint32_t *>(&a);
        q.y = *reinterpret_cast<uint32_t *>(&b);
        *reinterpret_cast<uint2 *>(p) = q;
    }
};

// grid: (ceil(n_out/64), ceil(c_out/64)); block 256 = 16 (tx: 4 channels each) x 16 (ty: 4 rows each)
template <typename T>
__global__ void __launch_bounds__(256)
conv_fwd_simt(const T *__restrict__ feat, const T *__restrict__ weight, const int *__restrict__ nbr, int ld,
              int K, int n_out, const int *__restrict__ n_out_dev, int c_in, int c_out,
              const float *__restrict__ scale, const float *__restrict__ shift, const float *__restrict__ bias,
              int flags, T *__restrict__ out)
{
    __shared__ __align__(16) float As[kKC][kLdA];   // [channel][row]
    __shared__ __align__(16) float Ws[kKC][kTN];    // [channel][out channel]
    if (n_out_dev) { const int m = __ldg(n_out_dev); n_out = m < n_out ? m : n_out; }
    const int row0 = blockIdx.x * kTM, col0 = blockIdx.y * kTN;
    if (row0 >= n_out) return;
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int g_row = tid >> 2, g_chunk = tid & 3;          // gather role: row in tile, 4-channel chunk
    const int w_row = tid >> 4, w_chunk = tid & 15;         // weight role: channel in stage, 4-col chunk
    const bool vec_ok = (c_in % 4 == 0), wvec_ok = (c_out % 4 == 0);

    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    for (int k = 0; k < K; ++k) {
        int src = -1;
        if (row0 + g_row < n_out) src = __ldg(nbr + (size_t)k * ld + row0 + g_row);
        if (!__syncthreads_or(src >= 0)) continue;   // nobody in this tile has a neighbour at offset k
        const T *wk = weight + (size_t)k * c_in * c_out;
        for (int c0 = 0; c0 < c_in; c0 += kKC) {
            // gather 64 rows x 16 channels, transposed into As
            float v[4] = {0.f, 0.f, 0.f, 0.f};
            const int ch = c0 + g_chunk * 4;
            if (src >= 0 && ch < c_in) {
                const T *p = feat + (size_t)src * c_in + ch;
                if (vec_ok) Vec4<T>::load(p, v);
                else
                    for (int j = 0; j < 4; ++j) if (ch + j < c_in) v[j] = to_float(p[j]);
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) As[g_chunk * 4 + j][g_row] = v[j];
            // weights: 16 channels x 64 out channels
            float w[4] = {0.f, 0.f, 0.f, 0.f};
            const int wc = c0 + w_row, wcol = col0 + w_chunk * 4;
            if (wc < c_in && wcol < c_out) {
                const T *p = wk + (size_t)wc * c_out + wcol;
                if (wvec_ok) Vec4<T>::load(p, w);
                else
                    for (int j = 0; j < 4; ++j) if (wcol + j < c_out) w[j] = to_float(p[j]);
            }
            *reinterpret_cast<float4 *>(&Ws[w_row][w_chunk * 4]) = make_float4(w[0], w[1], w[2], w[3]);
            __syncthreads();
#pragma unroll
            for (int kk = 0; kk < kKC; ++kk) {
                const float4 a = *reinterpret_cast<const float4 *>(&As[kk][ty * 4]);
                const float4 b = *reinterpret_cast<const float4 *>(&Ws[kk][tx * 4]);
                const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
            }
            __syncthreads();
        }
    }
    // epilogue: (y + bias)*scale + shift, relu -- the conv bias comes BEFORE the folded BatchNorm, as in the reference
    const int col = col0 + tx * 4;
    if (col >= c_out) return;
    float sc[4] = {1.f, 1.f, 1.f, 1.f}, sh[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int j = 0; j < 4; ++j)
        if (col + j < c_out) {
            if (scale) sc[j] = __ldg(scale + col + j);
            if (shift) sh[j] = __ldg(shift + col + j);
            if (bias) sh[j] = fmaf(__ldg(bias + col + j), sc[j], sh[j]);
        }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int row = row0 + ty * 4 + i;
        if (row >= n_out) continue;
        float y[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            y[j] = fmaf(acc[i][j], sc[j], sh[j]);
            if (flags & PCDB_EPI_RELU) y[j] = fmaxf(y[j], 0.f);
        }
        T *p = out + (size_t)row * c_out + col;
        if (wvec_ok) Vec4<T>::store(p, y);
        else
            for (int j = 0; j < 4; ++j) if (col + j < c_out) p[j] = from_float<T>(y[j]);
    }
}

// ---- backward (fp32) -------------------------------------------------------------------------
// grad_features[i] += grad_out[o] @ W_k^T for every pair (i = nbr[k][o]); one thread per
// (output row, input channel), atomics because several offsets/outputs hit the same input row.
__global__ void __launch_bounds__(256)
conv_bwd_input(const float *__restrict__ weight, const float *__restrict__ grad_out, const int *__restrict__ nbr,
               int ld, int K, int n_out, int c_in, int c_out, float *__restrict__ grad_feat)
{
    extern __shared__ float s_go[];  // (rows_per_block, c_out)
    const int rows_per_block = blockDim.x / c_in;
    const int lr = threadIdx.x / c_in, ci = threadIdx.x % c_in;
    const int row0 = blockIdx.x * rows_per_block;
    for (int t = threadIdx.x; t < rows_per_block * c_out; t += blockDim.x) {
        const int r = row0 + t / c_out;
        s_go[t] = r < n_out ? grad_out[(size_t)r * c_out + t % c_out] : 0.f;
    }
    __syncthreads();
    const int o = row0 + lr;
    if (lr >= rows_per_block || o >= n_out) return;
    const float *go = s_go + lr * c_out;
    for (int k = 0; k < K; ++k) {
        const int i = __ldg(nbr + (size_t)k * ld + o);
        if (i < 0) continue;
        const float *w = weight + ((size_t)k * c_in + ci) * c_out;
        float s = 0.f;
        for (int co = 0; co < c_out; ++co) s = fmaf(go[co], __ldg(w + co), s);
        atomicAdd(grad_feat + (size_t)i * c_in + ci, s);
    }
}

// grad_weight[k][ci][co] += sum_o feat[nbr[k][o]][ci] * grad_out[o][co].
// grid: (K, row chunks); block: 256 threads tile the (c_in, c_out) matrix; one atomicAdd per
// element per chunk.
__global__ void __launch_bounds__(256)
conv_bwd_weight(const float *__restrict__ feat, const float *__restrict__ grad_out, const int *__restrict__ nbr,
                int ld, int n_out, int c_in, int c_out, int rows_per_chunk, float *__restrict__ grad_weight)
{
    extern __shared__ float sm[];  // 32 rows of (c_in + c_out)
    float *s_f = sm, *s_g = sm + 32 * c_in;
    const int k = blockIdx.x;
    const int r_begin = blockIdx.y * rows_per_chunk;
    const int r_end = min(n_out, r_begin + rows_per_chunk);
    const int elems = c_in * c_out;
    constexpr int kMaxPerThread = 64;  // supports c_in*c_out <= 16384
    float acc[kMaxPerThread];
#pragma unroll
    for (int e = 0; e < kMaxPerThread; ++e) acc[e] = 0.f;
    for (int r0 = r_begin; r0 < r_end; r0 += 32) {
        __syncthreads();
        for (int t = threadIdx.x; t < 32 * c_in; t += blockDim.x) {
            const int r = r0 + t / c_in;
            int i = -1;
            if (r < r_end) i = __ldg(nbr + (size_t)k * ld + r);
            s_f[t] = i >= 0 ? __ldg(feat + (size_t)i * c_in + t % c_in) : 0.f;
        }
        for (int t = threadIdx.x; t < 32 * c_out; t += blockDim.x) {
            const int r = r0 + t / c_out;
            s_g[t] = r < r_end ? __ldg(grad_out + (size_t)r * c_out + t % c_out) : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int e = 0; e < kMaxPerThread; ++e) {
            const int idx = threadIdx.x + e * 256;
            if (idx < elems) {
                const int ci = idx / c_out, co = idx % c_out;
                float s = acc[e];
                for (int r = 0; r < 32; ++r) s = fmaf(s_f[r * c_in + ci], s_g[r * c_out + co], s);
                acc[e] = s;
            }
        }
    }
#pragma unroll
    for (int e = 0; e < kMaxPerThread; ++e) {
        const int idx = threadIdx.x + e * 256;
        if (idx < elems && acc[e] != 0.f) atomicAdd(grad_weight + (size_t)k * elems + idx, acc[e]);
    }
}

// The same reduction, register-tiled: the (c_in, c_out) matrix is cut into 4x4 tiles, a thread owns one tile (up to
// four when c_in*c_out > 4096) and, when there are fewer tiles than threads, the rows of a stage are dealt round-robin to
// 256/tiles thread groups whose partial sums meet in shared memory.  Per row and tile: two LDS.128 feed 16 FMAs (the
// kernel above needs two LDS per FMA).  The rows of a chunk are first compacted to the ones that HAVE a neighbour at
// offset k (ballot + prefix), so the empty two thirds of a submanifold map cost nothing.  c_in, c_out multiples of 4.
constexpr int kWgStage = 32;
constexpr int kWgMaxTiles = 4;
__global__ void __launch_bounds__(256)
conv_bwd_weight_tiled(const float *__restrict__ feat, const float *__restrict__ grad_out, const int *__restrict__ nbr,
                      int ld, int n_out, int c_in, int c_out, int rows_per_chunk, float *__restrict__ grad_weight)
{
    extern __shared__ __align__(16) float sm[];
    __shared__ int2 s_pairs[256];
    __shared__ int s_wcnt[8];
    float *s_f = sm, *s_g = sm + kWgStage * c_in;          // [32][c_in], [32][c_out]; later the cross-group reduction
    const int k = blockIdx.x;
    const int r_begin = blockIdx.y * rows_per_chunk, r_end = min(n_out, r_begin + rows_per_chunk);
    const int tci_n = c_in >> 2, tco_n = c_out >> 2, tiles = tci_n * tco_n;
    const int groups = tiles >= 256 ? 1 : 256 / tiles;
    const int per_thread = tiles >= 256 ? tiles / 256 : 1;
    const int grp = tiles >= 256 ? 0 : threadIdx.x / tiles;
    const int tile0 = tiles >= 256 ? threadIdx.x : threadIdx.x % tiles;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float acc[kWgMaxTiles][4][4];
#pragma unroll
    for (int t = 0; t < kWgMaxTiles; ++t)
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b) acc[t][a][b] = 0.f;

    for (int r0 = r_begin; r0 < r_end; r0 += 256) {
        const int r = r0 + threadIdx.x;
        const int i = r < r_end ? __ldg(nbr + (size_t)k * ld + r) : -1;
        const uint32_t m = __ballot_sync(0xffffffffu, i >= 0);
        if (lane == 0) s_wcnt[warp] = __popc(m);
        __syncthreads();
        int pos = __popc(m & ((1u << lane) - 1u)), total = 0;
#pragma unroll
        for (int w = 0; w < 8; ++w) {
            if (w < warp) pos += s_wcnt[w];
            total += s_wcnt[w];
        }
        if (i >= 0) s_pairs[pos] = make_int2(i, r);
        __syncthreads();
        for (int t0 = 0; t0 < total; t0 += kWgStage) {
            const int rows = min(kWgStage, total - t0);
            for (int t = threadIdx.x; t < rows * tci_n; t += 256) {
                const int row = t / tci_n, c4 = t - row * tci_n;
                *reinterpret_cast<float4 *>(s_f + row * c_in + c4 * 4) =
                    __ldg(reinterpret_cast<const float4 *>(feat + (size_t)s_pairs[t0 + row].x * c_in) + c4);
            }
            for (int t = threadIdx.x; t < rows * tco_n; t += 256) {
                const int row = t / tco_n, c4 = t - row * tco_n;
                *reinterpret_cast<float4 *>(s_g + row * c_out + c4 * 4) =
                    __ldg(reinterpret_cast<const float4 *>(grad_out + (size_t)s_pairs[t0 + row].y * c_out) + c4);
            }
            __syncthreads();
#pragma unroll
            for (int t = 0; t < kWgMaxTiles; ++t) {
                if (t >= per_thread) break;
                const int tile = tile0 + t * 256;
                const int ci0 = (tile / tco_n) * 4, co0 = (tile % tco_n) * 4;
                for (int rr = grp; rr < rows; rr += groups) {
                    const float4 a = *reinterpret_cast<const float4 *>(s_f + rr * c_in + ci0);
                    const float4 b = *reinterpret_cast<const float4 *>(s_g + rr * c_out + co0);
                    const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
                    for (int x = 0; x < 4; ++x)
#pragma unroll
                        for (int y = 0; y < 4; ++y) acc[t][x][y] = fmaf(av[x], bv[y], acc[t][x][y]);
                }
            }
            __syncthreads();
        }
    }
    float *gw = grad_weight + (size_t)k * c_in * c_out;
    if (groups == 1) {
#pragma unroll
        for (int t = 0; t < kWgMaxTiles; ++t) {
            if (t >= per_thread) break;
            const int tile = tile0 + t * 256;
            const int ci0 = (tile / tco_n) * 4, co0 = (tile % tco_n) * 4;
#pragma unroll
            for (int x = 0; x < 4; ++x)
#pragma unroll
                for (int y = 0; y < 4; ++y)
                    if (acc[t][x][y] != 0.f) atomicAdd(gw + (size_t)(ci0 + x) * c_out + co0 + y, acc[t][x][y]);
        }
        return;
    }
    // partial sums of the thread groups -> shared memory [group][tile][16] -> one atomicAdd per matrix element
    float *s_red = sm;
#pragma unroll
    for (int x = 0; x < 4; ++x)
#pragma unroll
        for (int y = 0; y < 4; ++y) s_red[(grp * tiles + tile0) * 16 + x * 4 + y] = acc[0][x][y];
    __syncthreads();
    for (int e = threadIdx.x; e < tiles * 16; e += 256) {
        float v = 0.f;
        for (int g = 0; g < groups; ++g) v += s_red[g * tiles * 16 + e];
        if (v != 0.f) {
            const int tile = e >> 4, x = (e >> 2) & 3, y = e & 3;
            atomicAdd(gw + (size_t)((tile / tco_n) * 4 + x) * c_out + (tile % tco_n) * 4 + y, v);
        }
    }
}

int launch_conv_fwd_simt(const void *features, const void *weight, const int32_t *nbr, int ld, int K, int n_out,
                         const int32_t *n_out_dev, int c_in, int c_out, int dtype, const float *scale,
                         const float *shift, const float *bias, int flags, void *out, cudaStream_t stream)
{
    dim3 grid((n_out + kTM - 1) / kTM, (c_out + kTN - 1) / kTN);
    if (dtype == PCDB_BF16)
        conv_fwd_simt<__nv_bfloat16><<<grid, 256, 0, stream>>>((const __nv_bfloat16 *)features, (const __nv_bfloat16 *)weight,
                                                               nbr, ld, K, n_out, n_out_dev, c_in, c_out, scale, shift,
                                                               bias, flags, (__nv_bfloat16 *)out);
    else
        conv_fwd_simt<float><<<grid, 256, 0, stream>>>((const float *)features, (const float *)weight, nbr, ld, K, n_out,
                                                       n_out_dev, c_in, c_out, scale, shift, bias, flags, (float *)out);
    return check_launch("pcdb_sparse_conv_fwd(simt)");
}

}  // namespace pcdb

using namespace pcdb;

extern "C" int pcdb_sparse_conv_bwd(const float *features, const float *weight, const float *grad_out,
                                    const int32_t *nbr, int ld, int kernel_volume, int n_in, int n_out,
                                    int c_in, int c_out, float *grad_features, float *grad_weight, void *stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    if (!features || !weight || !grad_out || !nbr || n_out < 0 || n_in < 0 || c_in < 1 || c_out < 1 ||
        c_in > 256 || c_out > 1024 || (size_t)c_in * c_out > 16384) {
        set_last_error("pcdb_sparse_conv_bwd: invalid or unsupported shape (c_in=%d c_out=%d)", c_in, c_out);
        return kInvalidArgument;
    }
    if (n_out == 0) return kOk;
    if (grad_features) {
        const int rows_per_block = 256 / c_in > 0 ? 256 / c_in : 1;
        const int threads = rows_per_block * c_in;
        const int blocks = (n_out + rows_per_block - 1) / rows_per_block;
        conv_bwd_input<<<blocks, threads, sizeof(float) * rows_per_block * c_out, stream>>>(
            weight, grad_out, nbr, ld, kernel_volume, n_out, c_in, c_out, grad_features);
    }
    if (grad_weight) {
        int chunks = (kNumSMs * 4 + kernel_volume - 1) / kernel_volume;
        int rows_per_chunk = (n_out + chunks - 1) / chunks;
        const int tiles = (c_in / 4) * (c_out / 4);
        const bool tiled = c_in % 4 == 0 && c_out % 4 == 0 && tiles <= 256 * kWgMaxTiles && (tiles >= 256 ? tiles % 256 == 0 : 256 % tiles == 0);
        if (tiled) {
            rows_per_chunk = (rows_per_chunk + 255) / 256 * 256;
            chunks = (n_out + rows_per_chunk - 1) / rows_per_chunk;
            size_t smem = sizeof(float) * kWgStage * (c_in + c_out);
            if (tiles < 256 && smem < sizeof(float) * 256 * 16) smem = sizeof(float) * 256 * 16;     // cross-group reduction
            // per device and context, cheap: set on every call (a process-wide "done once" flag breaks on a second GPU)
            if (cudaFuncSetAttribute(conv_bwd_weight_tiled, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024) != cudaSuccess)
                return check_launch("pcdb_sparse_conv_bwd(cudaFuncSetAttribute)");
            conv_bwd_weight_tiled<<<dim3(kernel_volume, chunks), 256, smem, stream>>>(features, grad_out, nbr, ld, n_out, c_in, c_out,
                                                                                    rows_per_chunk, grad_weight);
        } else {
            rows_per_chunk = (rows_per_chunk + 31) / 32 * 32;
            chunks = (n_out + rows_per_chunk - 1) / rows_per_chunk;
            conv_bwd_weight<<<dim3(kernel_volume, chunks), 256, sizeof(float) * 32 * (c_in + c_out), stream>>>(
                features, grad_out, nbr, ld, n_out, c_in, c_out, rows_per_chunk, grad_weight);
        }
    }
    return check_launch("pcdb_sparse_conv_bwd");
}
