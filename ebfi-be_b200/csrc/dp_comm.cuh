// dp_comm.cuh — all-reduce over NVLink peer memory whose first half runs INSIDE the producing kernel.
//
// Every rank owns one symmetric allocation (same layout on all ranks, mapped into every peer's address space by the
// host: torch symmetric memory / cudaIpc — plumbing, include/ebfi_b200.h `ebfi_dp_comm`):
//     [0, 256)        : epoch counter + block ticket (used by the owner only)
//     [256, 512)      : flags[src rank] (uint32), written by the peers
//     [512, ...)      : data[parity][src rank][capacity] (fp32): slot [.][q] is WRITTEN BY RANK q over NVLink, read by the owner
// publish (any kernel that produces the values; all ranks issue the same sequence): epoch = counter + 1; every block
// stores its share of the local values into slot [epoch & 1][my rank] of EVERY rank's buffer (P2P stores: push, not
// pull — a peer that is saturating its HBM would serve remote loads tens of microseconds late, measured at N = 8),
// fences (system scope) and takes a ticket; the LAST block of the grid
// issues one system-scope fence (cumulative over the other blocks' stores, which it observed through the ticket), stores
// `epoch` into flags[my rank] of every peer (st.release.sys over NVLink) and advances the counter — device side, so
// the launch can sit in a CUDA graph. It never waits.
// complete (a tiny kernel at the point of use): waits until flags[q] reached the current epoch for every peer q
// (ld.acquire.sys on local memory), then reads the world slots of its OWN buffer (ld.relaxed.sys — never from a stale L1
// line; no traffic to the peers at all) and adds them in rank order: the sums are identical on all ranks and run to
// run. Between the two halves the NVLink latency and the skew between the ranks hide behind whatever the stream runs (at most one publish outstanding per communicator).
// Double-buffered data: a rank can be at most one publish ahead of a peer, because its complete(e+1) needs the peer's
// flag e+1, which the peer only sets after its own complete(e) has been issued. Nothing that waits holds up a
// publish, so the exchange cannot deadlock; a dead peer trips the watchdog (~4 s) into a trap instead of a hang.
#pragma once
#include <cstdint>
#include <cstdio>

#include "common.cuh"

namespace ebfi_dp {

constexpr int MAX_WORLD = 8;
constexpr size_t CTR_BYTES = 256, FLAG_BYTES = 256, HDR_BYTES = CTR_BYTES + FLAG_BYTES;

struct View {
    int world, rank;
    unsigned char *base[MAX_WORLD];
    size_t cap;             // floats per (parity, source rank) slot
};

inline size_t bytes_for(size_t n_floats) { return HDR_BYTES + 2 * MAX_WORLD * ebfi::round_up(n_floats, (size_t)64) * sizeof(float); }

// host: validate and convert the C struct; world == 1 is allowed (no peers: the exchange degenerates to a copy)
int make_view(const ebfi_dp_comm *c, size_t n_floats, View &v);
// host: a[0, na) | b[0, nb) <- sum over ranks, in place. mode 0: publish kernel + complete kernel; mode 1: publish only
// (the values stay local); mode 2: complete the last publish (dp_comm.cu)
int allreduce_sum(cudaStream_t st, const View &v, float *a, size_t na, float *b, size_t nb, int mode = 0);

#ifdef __CUDACC__
__device__ __forceinline__ unsigned *ctr(const View &v) { return reinterpret_cast<unsigned *>(v.base[v.rank]); }
__device__ __forceinline__ unsigned *flags(const View &v, int q) { return reinterpret_cast<unsigned *>(v.base[q] + CTR_BYTES); }
// slot of source rank `src` inside rank `owner`'s buffer
__device__ __forceinline__ float *data(const View &v, int owner, unsigned parity, int src)
{
    return reinterpret_cast<float *>(v.base[owner] + HDR_BYTES) + ((size_t)parity * MAX_WORLD + src) * v.cap;
}
// hand value x of element e to every rank (the local copy included)
__device__ __forceinline__ void push(const View &v, unsigned epoch, size_t e, float x)
{
    for (int t = 0; t < v.world; ++t) data(v, t, epoch & 1u, v.rank)[e] = x;
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

// Called by ALL threads of EVERY block of the publishing grid after their push() calls: the last
// block to arrive hands the whole epoch to the peers and advances the counter. blockDim.x >= world.
__device__ __forceinline__ void publish(const View &v, unsigned epoch)
{
    __shared__ int last;
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence_system();                                   // this block's (peer) data stores before its ticket
        unsigned *c = ctr(v);
        const unsigned nblk = gridDim.x * gridDim.y * gridDim.z;
        last = atomicAdd(c + 1, 1u) == nblk - 1;
        if (last) c[1] = 0u;
    }
    __syncthreads();
    if (last) {
        const int q = (int)threadIdx.x;
        if (q < v.world && q != v.rank) {
            __threadfence_system();                               // every block's data stores, system-wide
            st_release_sys(flags(v, q) + v.rank, epoch);
        }
        __syncthreads();
        if (threadIdx.x == 0) *reinterpret_cast<volatile unsigned *>(ctr(v)) = epoch;
    }
}

// Called by ALL threads of the block: wait until every peer has published `epoch`.
__device__ __forceinline__ void wait_peers(const View &v, unsigned epoch)
{
    const int q = (int)threadIdx.x;
    if (q < v.world && q != v.rank) {
        const unsigned *mine = flags(v, v.rank) + q;
        const long long t0 = clock64();
        while ((int)(ld_acquire_sys(mine) - epoch) < 0) {
            if (clock64() - t0 > (8LL << 30)) {                   // ~4 s at 2 GHz: a peer never arrived
                printf("ebfi_dp: rank %d waited for rank %d epoch %u\n", v.rank, q, epoch);
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
    for (int q = 0; q < v.world; ++q) a += ld_relaxed_sys(data(v, v.rank, epoch & 1u, q) + e);
    return a;
}

#endif

}  // namespace ebfi_dp
