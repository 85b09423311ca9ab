// Shared device helpers for the pcdet_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>

namespace pcdb {

constexpr int kNumSMs = 148;  // B200: 2 dies x 74 SMs

// status codes returned by every C-ABI entry point (include/pcdet_b200.h)
enum Status : int {
    kOk = 0,
    kInvalidArgument = 1,
    kWorkspaceTooSmall = 2,
    kKeyOverflow = 3,      // batch * grid volume does not fit the 32-bit hash key
    kCudaError = 4,
    kUnsupported = 5,
};

void set_last_error(const char *fmt, ...);
int check_launch(const char *what);
void fill_i32(int *dst, int value, size_t count, cudaStream_t stream);

// ---------------------------------------------------------------------------------------------
// Open-addressing hash table in HBM.  One 64-bit word per slot: key in the high half, payload in the
// low half, so a single atomicMin on the word keeps the SMALLEST payload per key (first point of a
// voxel, first (input,offset) touching an output site) without a second array or a lock.
// ---------------------------------------------------------------------------------------------
constexpr unsigned long long kEmptySlot = 0xFFFFFFFFFFFFFFFFull;

__host__ __device__ __forceinline__ uint32_t hash_u32(uint32_t k)
{
    k ^= k >> 16; k *= 0x7feb352dU; k ^= k >> 15; k *= 0x846ca68bU; k ^= k >> 16;
    return k;
}

// Insert (key, payload) keeping the minimum payload; returns the slot index.
__device__ __forceinline__ uint32_t table_insert_min(unsigned long long *slots, uint32_t mask,
                                                     uint32_t key, uint32_t payload)
{
    const unsigned long long word = ((unsigned long long)key << 32) | payload;
    uint32_t s = hash_u32(key) & mask;
    // bounded probe: a table the caller sized too small reports 0xFFFFFFFF instead of spinning
    for (uint32_t probe = 0; probe <= mask; ++probe) {
        unsigned long long cur = *((volatile unsigned long long *)(slots + s));
        if (cur == kEmptySlot) {
            cur = atomicCAS(slots + s, kEmptySlot, word);
            if (cur == kEmptySlot) return s;
        }
        if ((uint32_t)(cur >> 32) == key) {
            if ((uint32_t)cur > payload) atomicMin(slots + s, word);
            return s;
        }
        s = (s + 1) & mask;
    }
    return 0xFFFFFFFFu;
}

// Ordered insertion (Amble & Knuth): every probe keeps the SMALLER word in the slot and carries the larger one on, so
// the final layout is a function of the key SET only -- whatever the order or interleaving of the inserts.  Numbering
// the occupied slots in slot order therefore gives deterministic row ids without any first-toucher bookkeeping.
// A word already present ends the walk.  Lookups are table_find as usual (a key never sits before its home slot and
// no empty slot separates them).  Returns false when the table is full.
__device__ __forceinline__ bool table_insert_ordered(unsigned long long *slots, uint32_t mask, unsigned long long word)
{
    uint32_t s = hash_u32((uint32_t)(word >> 32)) & mask;
    for (uint32_t probe = 0; probe <= mask; ++probe) {
        const unsigned long long cur = *((volatile unsigned long long *)(slots + s));
        if (cur == word) return true;
        if (cur > word) {                               // empty (all ones) or a larger word: take the slot
            const unsigned long long old = atomicMin(slots + s, word);
            if (old == kEmptySlot || old == word) return true;
            if (old > word) word = old;                 // displaced: carry it on; else somebody smaller got in first
        }
        s = (s + 1) & mask;
    }
    return false;
}

// Returns the slot holding key, or 0xFFFFFFFF when absent.  payload_out receives the low half.
__device__ __forceinline__ uint32_t table_find(const unsigned long long *__restrict__ slots,
                                               uint32_t mask, uint32_t key, uint32_t *payload_out)
{
    uint32_t s = hash_u32(key) & mask;
    for (uint32_t probe = 0; probe <= mask; ++probe) {
        const unsigned long long cur = __ldg(slots + s);
        if (cur == kEmptySlot) return 0xFFFFFFFFu;
        if ((uint32_t)(cur >> 32) == key) { *payload_out = (uint32_t)cur; return s; }
        s = (s + 1) & mask;
    }
    return 0xFFFFFFFFu;
}

// ---------------------------------------------------------------------------------------------
// Block-wide exclusive scan of one int per thread (BLOCK multiple of 32, <= 1024).
// ---------------------------------------------------------------------------------------------
template <int BLOCK>
__device__ __forceinline__ int block_exclusive_scan(int v, int *total)
{
    __shared__ int warp_sums[BLOCK / 32];
    __shared__ int block_total;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= d) inc += t;
    }
    if (lane == 31) warp_sums[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        int w = lane < BLOCK / 32 ? warp_sums[lane] : 0;
        int winc = w;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            int t = __shfl_up_sync(0xffffffffu, winc, d);
            if (lane >= d) winc += t;
        }
        if (lane < BLOCK / 32) warp_sums[lane] = winc - w;
        if (lane == 31) block_total = winc;
    }
    __syncthreads();
    const int res = inc - v + warp_sums[warp];
    if (total) *total = block_total;
    __syncthreads();  // warp_sums / block_total may be reused by the next call
    return res;
}

// The last block to finish (ticket == gridDim.x-1) turns block_sums[0..nblocks) into exclusive
// offsets in place and writes the grand total to block_sums[nblocks].  Call from ALL threads of
// every block after block_sums[blockIdx.x] has been written by thread 0.
template <int BLOCK>
__device__ __forceinline__ void last_block_scan(int *block_sums, int nblocks, unsigned int *ticket)
{
    // idle state of *ticket is 0xFFFFFFFF (so it can live in a 0xFF-memset region): the tickets
    // handed out are 0xFFFFFFFF, 0, 1, ... and the last arrival sees nblocks-2 (mod 2^32).
    __shared__ bool is_last;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned int t = atomicAdd(ticket, 1u);
        is_last = (t + 1u == (unsigned int)nblocks - 1u);
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    int carry = 0;
    for (int base = 0; base < nblocks; base += BLOCK) {
        const int i = base + threadIdx.x;
        const int v = i < nblocks ? ((volatile int *)block_sums)[i] : 0;
        int tot;
        const int ex = block_exclusive_scan<BLOCK>(v, &tot);
        if (i < nblocks) block_sums[i] = carry + ex;
        carry += tot;
    }
    if (threadIdx.x == 0) {
        block_sums[nblocks] = carry;
        *ticket = 0xFFFFFFFFu;  // re-arm for the next launch
    }
}

__device__ __forceinline__ float to_float(float v) { return v; }
__device__ __forceinline__ float to_float(__nv_bfloat16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_float(float v);
template <> __device__ __forceinline__ float from_float<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_float<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

inline uint32_t next_pow2(uint64_t v)
{
    uint64_t p = 1;
    while (p < v) p <<= 1;
    return (uint32_t)p;
}

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

}  // namespace pcdb
