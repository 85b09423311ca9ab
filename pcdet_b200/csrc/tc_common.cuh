// tcgen05 / TMEM / mbarrier / cp.async helpers shared by the tensor-core kernels (sparse_conv_tc.cu, sparse_conv_wgrad_tc.cu).
#pragma once
#include "common.cuh"
#include <cuda.h>   // CUtensorMap (types only)

namespace pcdb {
namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// One lane of a converged warp.  Together with values made provably warp-uniform (uniform(), below) this lets
// ptxas keep descriptors / barrier addresses in uniform registers: issued from a divergent `lane == 0` branch
// every tcgen05.mma is wrapped in an ELECT + 8x R2UR "waterfall" loop (measured: 590 cycles per offset in
// the MMA thread instead of ~150).
__device__ __forceinline__ bool elect_one()
{
    uint32_t pred;
    asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
    return pred != 0;
}
// x is the same in all lanes (e.g. read from shared memory); the broadcast tells the compiler so
__device__ __forceinline__ uint32_t uniform(uint32_t x) { return __shfl_sync(0xffffffffu, x, 0); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ bool mbar_test_wait(uint32_t bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_spin(uint32_t bar, uint32_t parity)
{
#pragma unroll 1
    for (uint32_t spin = 0; spin < (1u << 28); ++spin)
        if (mbar_test_wait(bar, parity)) return;
    __trap();
}
// Bounded wait: a protocol bug traps (launch failure) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
#pragma unroll 1
    for (uint32_t spin = 0; spin < (1u << 24); ++spin)
        if (mbar_try_wait(bar, parity)) return;
    __trap();
}

__device__ __forceinline__ void cp_async16(uint32_t dst, const void *src, uint32_t src_bytes)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
// One 16-byte piece of input row `src` (ROW_BYTES apart from `feat_piece`) into shared memory; nothing at all
// happens for src < 0 (no neighbour): that tile row is masked out of the MMA instead of being zero-filled.
// Four instructions: ISETP, LEA, LEA.HI.X, @p LDGSTS.
template <int ROW_BYTES>
__device__ __forceinline__ void gather_piece(uint32_t dst, const uint8_t *feat_piece, int src)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 a;\n\t"
        "setp.ge.s32 p, %2, 0;\n\t"
        "mad.wide.s32 a, %2, %3, %1;\n\t"
        "@p cp.async.cg.shared.global [%0], [a], 16;\n\t}"
        ::"r"(dst), "l"(feat_piece), "r"(src), "n"(ROW_BYTES) : "memory");
}
// The mbarrier receives one arrival from this thread once ALL its earlier cp.async copies have landed
// (.noinc: the arrival counts against the barrier's expected count), so a producer never waits for
// its own loads -- it only waits for a free stage.
__device__ __forceinline__ void cp_async_arrive(uint32_t bar)
{
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// 4 rows (box = {row width, 1}) of a 2-D tensor map into 4 consecutive swizzled rows of shared memory
__device__ __forceinline__ void tma_gather4(uint32_t dst, const CUtensorMap *tmap, uint32_t bar, int col, int r0, int r1,
                                            int r2, int r3)
{
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile::gather4.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
        ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(bar), "r"(col), "r"(r0), "r"(r1), "r"(r2), "r"(r3) : "memory");
}
__device__ __forceinline__ void bulk_copy_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
// make generic-proxy shared-memory stores visible to the async proxy (tcgen05.mma operand reads)
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols)
{
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols)
{
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

// D[tmem] += A[smem desc] * B[smem desc]; single-thread issue.  Bit r of the 128-bit `off` vector keeps
// accumulator row (TMEM lane) r untouched: tile rows without a neighbour at this offset need no operand data.
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, const uint4 &off)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.eq.b32 p, 0, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, {%4, %5, %6, %7}, p;\n\t}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(off.x), "r"(off.y), "r"(off.z), "r"(off.w) : "memory");
}
// arrive on an mbarrier once every previously issued MMA of this thread has completed
__device__ __forceinline__ void umma_commit(uint32_t bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// 32 lanes x 16 consecutive fp32 columns: thread t of the warp receives row (lane_base + t)
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t *r)
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
}
// zero 32 lanes x 16 fp32 columns (the accumulator is always accumulated into, see umma_bf16)
__device__ __forceinline__ void tmem_zero16(uint32_t taddr)
{
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1};"
        ::"r"(taddr), "r"(0u) : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// Byte offset of 16-byte chunk c of row r inside a swizzled K-major operand tile (Swizzle<B,4,3>).
template <int ROW_BYTES, int SW_BITS>
__device__ __forceinline__ uint32_t swizzled_offset(uint32_t r, uint32_t c)
{
    const uint32_t o = r * ROW_BYTES + c * 16;
    return o ^ (((o >> 7) & ((1u << SW_BITS) - 1u)) << 4);
}

// Epilogue staging: any fixed permutation of 16-byte units inside 1 KB blocks works (write and read use the
// same one); Swizzle<3,4,3> keeps both the row-per-thread writes and the linear reads conflict free.
__device__ __forceinline__ uint32_t swizzle_out(uint32_t o) { return o ^ (((o >> 7) & 7u) << 4); }
__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d)
{
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void ld_shared_v4(uint32_t addr, uint32_t &a, uint32_t &b, uint32_t &c, uint32_t &d)
{
    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "r"(addr) : "memory");
}

}  // namespace tc
}  // namespace pcdb
