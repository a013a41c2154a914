// umma.cuh — thin inline-PTX layer over the sm_100a tensor-core path used by the DCN GEMMs:
// tcgen05.mma (kind::tf32, operands in shared memory, accumulator in TMEM), TMEM
// alloc/ld/dealloc, mbarriers, tcgen05.commit and the 1-D bulk (TMA) copy.
//
// Shared-memory operand layout used everywhere in this repo: the canonical NO-SWIZZLE
// ("interleaved") layout made of 128-byte core matrices = 8 rows x 16 bytes (4 fp32).
//   K-major operand  (rows = M or N index, 16-byte chunks run along K):
//       addr(row, kchunk) = base + (row / 8) * SBO + kchunk * LBO + (row % 8) * 16
//   MN-major operand (16-byte chunks run along M/N, the 8 rows of a core matrix along K):
//       addr(mnchunk, k)  = base + mnchunk * SBO + (k / 8) * LBO + (k % 8) * 16
// One kind::tf32 MMA consumes K = 8 elements: two K-chunks of a K-major operand, one
// 8-row group of an MN-major operand.
//
// fp32-accurate products on the TF32 pipe ("3xTF32"): x = hi + lo with hi = x with the low
// 13 mantissa bits cleared (exactly a TF32 number) and lo = x - hi (exact in fp32);
// a*b ~= lo_a*hi_b + hi_a*lo_b + hi_a*hi_b, accumulated in fp32 by the tensor core. The
// dropped lo*lo term and the truncation of lo to TF32 are both ~2^-22 relative.
#pragma once
#include <cstdint>
#include <cstdio>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

namespace umma {

__device__ __forceinline__ uint32_t smem_u32(const void *p)
{
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// ---- 3xTF32 split -----------------------------------------------------------------
__device__ __forceinline__ float tf32_hi(float x) { return __uint_as_float(__float_as_uint(x) & 0xFFFFE000u); }
__device__ __forceinline__ void split_tf32(float x, float &hi, float &lo)
{
    hi = tf32_hi(x);
    lo = x - hi;
}

// ---- bf16 pair split ("bf16x3") -----------------------------------------------------
// x ~= hi + lo with hi = bf16(x), lo = bf16(x - hi): 16 significand bits, products
// lo_a*hi_b + hi_a*lo_b + hi_a*hi_b carry ~2^-16 relative error — used where the gate is 1e-4
// (gradients) because the operands take half the shared memory of the TF32 pair.
__device__ __forceinline__ void split_bf16(float x, unsigned short &hi, unsigned short &lo)
{
    const __nv_bfloat16 h = __float2bfloat16_rn(x);
    const __nv_bfloat16 l = __float2bfloat16_rn(x - __bfloat162float(h));
    hi = __bfloat16_as_ushort(h);
    lo = __bfloat16_as_ushort(l);
}

// ---- descriptors ------------------------------------------------------------------
// 64-bit shared-memory matrix descriptor, SWIZZLE_NONE, Blackwell version field = 1
// (bit layout: cute/arch/mma_sm100_desc.hpp, union SmemDescriptor).
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes)
{
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFFu);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}

// 32-bit instruction descriptor for kind::tf32 with fp32 accumulation
// (union InstrDescriptor): c_format=F32 [4,6), a/b_format=TF32 [7,10)/[10,13),
// a_major [15], b_major [16] (0 = K-major, 1 = MN-major), N>>3 [17,23), M>>4 [24,29).
__host__ __device__ constexpr uint32_t instr_desc_tf32(int M, int N, int a_mn_major, int b_mn_major)
{
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) |
           ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// kind::f16 with BF16 operands, fp32 accumulation: a/b_format = 1.
__host__ __device__ constexpr uint32_t instr_desc_bf16(int M, int N)
{
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// ---- MMA, commit, fences ------------------------------------------------------------
// D[tmem] (+)= A[smem] * B[smem]; issued by ONE thread on behalf of the CTA.
__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                         bool accumulate)
{
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"((uint32_t)accumulate)
        : "memory");
}

// Same for 16-bit operands (K = 16 elements = two 16-byte chunks per instruction).
__device__ __forceinline__ void mma_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                        bool accumulate)
{
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"((uint32_t)accumulate)
        : "memory");
}

// Arrive on `bar` once every MMA issued so far by this thread has completed
// (implies tcgen05.fence::before_thread_sync).
__device__ __forceinline__ void commit(uint64_t *bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}

__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// Make this thread's generic-proxy shared-memory writes visible to the async proxy
// (tensor core / bulk copy) — required between st.shared of an operand and the MMA.
__device__ __forceinline__ void fence_smem_to_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// One lane of a converged warp. Pattern for the MMA-issuing warp: ALL lanes run the (warp-uniform) issue loop so
// that descriptor arithmetic stays on the uniform datapath, and only the tcgen05 / bulk-copy / commit instructions
// are predicated on the elected lane — `if (lane == 0) { whole loop }` costs ~40 instructions per MMA instead of ~4.
__device__ __forceinline__ bool elect_one()
{
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}

// ---- mbarrier -----------------------------------------------------------------------
// -DEBFI_DEBUG_HANG: every wait loop reports (block, thread, shared address, parity) and traps after ~16 M polls, so a
// protocol deadlock shows up as an error with a location instead of a hung GPU.
#ifdef EBFI_DEBUG_HANG
#define EBFI_HANG_CHECK(n, bar, parity)                                                                     \
    if (++(n) > (1u << 24)) {                                                                               \
        printf("mbarrier wait stuck: block %d thread %d bar 0x%x parity %u line %d\n", (int)blockIdx.x,     \
               (int)threadIdx.x, smem_u32(bar), parity, __LINE__);                                          \
        __trap();                                                                                           \
    }
#else
#define EBFI_HANG_CHECK(n, bar, parity)
#endif
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// Warp-level wait for warps that are NOT on the critical path (epilogue, producer): one lane polls, with a
// hardware suspend-time hint, and the warp reconverges. 384 threads spinning on try_wait take shared-memory cycles
// away from the tensor core's operand reads.
__device__ __forceinline__ void mbar_wait_warp(uint64_t *bar, uint32_t parity)
{
    if ((threadIdx.x & 31) == 0) {
        uint32_t done, polls = 0;
        (void)polls;
        do {
            asm volatile(
                "{\n\t.reg .pred p;\n\t"
                "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
                "selp.u32 %0, 1, 0, p;\n\t}\n"
                : "=r"(done)
                : "r"(smem_u32(bar)), "r"(parity), "r"(20000u)
                : "memory");
            EBFI_HANG_CHECK(polls, bar, parity)
        } while (!done);
    }
    __syncwarp();
}

// Single-thread wait with a hardware suspend-time hint: for loader / controller lanes whose spinning would otherwise
// take issue slots from the warps that do the work.
__device__ __forceinline__ void mbar_wait_sleep(uint64_t *bar, uint32_t parity)
{
    uint32_t done, polls = 0;
    (void)polls;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}\n"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity), "r"(20000u)
            : "memory");
        EBFI_HANG_CHECK(polls, bar, parity)
    } while (!done);
}

__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    uint32_t done, polls = 0;
    (void)polls;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}\n"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
        EBFI_HANG_CHECK(polls, bar, parity)
    } while (!done);
}

// ---- 1-D bulk copy global -> shared (TMA engine), completes on an mbarrier ---------
// dst/src 16-byte aligned, bytes a multiple of 16.
__device__ __forceinline__ void bulk_g2s(void *smem_dst, const void *gmem_src, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// ---- TMEM ---------------------------------------------------------------------------
// Executed by ONE full warp. ncols: power of two in [32, 512]. The base address lands in *slot.
template <int NCOLS> __device__ __forceinline__ void tmem_alloc(uint32_t *slot)
{
    static_assert(NCOLS >= 32 && NCOLS <= 512 && (NCOLS & (NCOLS - 1)) == 0, "TMEM columns: power of 2 in [32,512]");
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "n"(NCOLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <int NCOLS> __device__ __forceinline__ void tmem_dealloc(uint32_t taddr)
{
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(NCOLS) : "memory");
}

// 32 lanes x 8 consecutive columns -> 8 registers per thread (thread i of the warp = TMEM lane
// base_lane + i; a warp may only touch lanes 32*(warp_id % 4) .. +31).
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float (&v)[8])
{
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr)
                 : "memory");
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ uint32_t tmem_addr(uint32_t base, int lane, int col)
{
    return base + ((uint32_t)lane << 16) + (uint32_t)col;
}

}  // namespace umma
