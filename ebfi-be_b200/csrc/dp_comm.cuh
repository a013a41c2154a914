// dp_comm.cuh — one-shot all-reduce over NVLink peer memory, usable INSIDE a producing kernel.
//
// Every rank owns one symmetric allocation (same layout on all ranks, mapped into every peer's address space by the
// host: torch symmetric memory / cudaIpc — plumbing, include/ebfi_b200.h `ebfi_dp_comm`):
//     [0, 256)                     : epoch counter + block ticket (used by the owner only)
//     [256, 256 + 8*1024*4)        : flags[src rank][block] (uint32), written by the peers
//     [.., + 2 * capacity * 4)     : data[parity][capacity] (fp32), written by the owner, read by the peers
// Protocol per kernel launch (all ranks launch the same grid): epoch = counter + 1; a block writes its slice of the
// local values into data[epoch & 1], fences, stores `epoch` into flags[my rank][block] of every peer (st.release.sys
// over NVLink), spins until its own flags[q][block] reached `epoch` for every peer q (ld.acquire.sys, local memory),
// then reads the same slice from every rank's data (ld.relaxed.sys — never from a stale L1 line) and adds them in rank
// order: the sum is identical on all ranks and run to run. The last block to finish bumps the counter (device side, so
// the launch can sit in a CUDA graph). Double-buffered data: a rank can run at most one launch ahead of a peer, because
// launch e+1 needs every peer's flags of e+1, which a peer only sets after its launch e has finished reading.
// A block only waits for blocks that signal BEFORE they wait, and blocks are dispatched in index order on every GPU:
// the lowest unfinished block index always completes, so the exchange cannot deadlock; a dead peer trips the
// watchdog (~4 s) into a trap instead of hanging the GPU.
#pragma once
#include <cstdint>
#include <cstdio>

#include "common.cuh"

namespace ebfi_dp {

constexpr int MAX_WORLD = 8, MAX_BLOCKS = 1024;
constexpr size_t CTR_BYTES = 256, FLAG_BYTES = (size_t)MAX_WORLD * MAX_BLOCKS * 4, HDR_BYTES = CTR_BYTES + FLAG_BYTES;

struct View {
    int world, rank;
    unsigned char *base[MAX_WORLD];
    size_t cap;             // floats per parity
};

inline size_t bytes_for(size_t n_floats) { return HDR_BYTES + 2 * ebfi::round_up(n_floats, (size_t)64) * sizeof(float); }

// host: validate and convert the C struct; world == 1 is allowed (no peers: the exchange degenerates to a copy)
int make_view(const ebfi_dp_comm *c, size_t n_floats, View &v);
// host: a[0, na) | b[0, nb) <- sum over ranks, in place (stand-alone exchange kernel, dp_comm.cu)
int allreduce_sum(cudaStream_t st, const View &v, float *a, size_t na, float *b, size_t nb);

#ifdef __CUDACC__
__device__ __forceinline__ unsigned *ctr(const View &v) { return reinterpret_cast<unsigned *>(v.base[v.rank]); }
__device__ __forceinline__ unsigned *flags(const View &v, int q) { return reinterpret_cast<unsigned *>(v.base[q] + CTR_BYTES); }
__device__ __forceinline__ float *data(const View &v, int q, unsigned parity)
{
    return reinterpret_cast<float *>(v.base[q] + HDR_BYTES) + (size_t)parity * v.cap;
}
__device__ __forceinline__ unsigned epoch_of_launch(const View &v) { return *reinterpret_cast<volatile unsigned *>(ctr(v)) + 1u; }

__device__ __forceinline__ void st_release_sys(unsigned *p, unsigned x)
{
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(x) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned *p)
{
    unsigned x;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(x) : "l"(p) : "memory");
    return x;
}
__device__ __forceinline__ float ld_relaxed_sys(const float *p)
{
    float x;
    asm volatile("ld.relaxed.sys.global.f32 %0, [%1];" : "=f"(x) : "l"(p) : "memory");
    return x;
}

// Called by ALL threads of the block after their stores into data(v, rank, epoch & 1): publish the block's slice to
// the peers and wait for theirs. blockDim.x >= world.
__device__ __forceinline__ void publish_and_wait(const View &v, unsigned epoch, int blk)
{
    __syncthreads();
    const int q = (int)threadIdx.x;
    if (q < v.world && q != v.rank) {
        __threadfence_system();                                   // the block's data stores, system-wide
        st_release_sys(flags(v, q) + v.rank * MAX_BLOCKS + blk, epoch);
        const unsigned *mine = flags(v, v.rank) + q * MAX_BLOCKS + blk;
        const long long t0 = clock64();
        while ((int)(ld_acquire_sys(mine) - epoch) < 0) {
            if (clock64() - t0 > (8LL << 30)) {                   // ~4 s at 2 GHz: a peer never arrived
                printf("ebfi_dp: rank %d block %d waited for rank %d epoch %u\n", v.rank, blk, q, epoch);
                __trap();
            }
        }
    }
    __syncthreads();
}

// sum over the ranks, in rank order, of element e of data[epoch & 1]
__device__ __forceinline__ float gather_sum(const View &v, unsigned epoch, size_t e)
{
    float a = 0.f;
    for (int q = 0; q < v.world; ++q) a += ld_relaxed_sys(data(v, q, epoch & 1u) + e);
    return a;
}

// Called by all threads of the block at the very end: the last block of the grid advances the epoch counter.
__device__ __forceinline__ void finish_launch(const View &v, unsigned epoch)
{
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned *c = ctr(v);
        const unsigned nblk = gridDim.x * gridDim.y * gridDim.z;
        if (atomicAdd(c + 1, 1u) == nblk - 1) {
            c[1] = 0u;
            __threadfence();
            *reinterpret_cast<volatile unsigned *>(c) = epoch;
        }
    }
}
#endif

}  // namespace ebfi_dp
