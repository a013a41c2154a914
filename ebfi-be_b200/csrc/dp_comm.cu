// dp_comm.cu — host side of the peer-memory all-reduce (dp_comm.cuh) and its stand-alone kernel: the data-parallel
// weight-gradient sum of SURVEY §8e for the paths that do not fuse it into their own reduction kernel.
#include "dp_comm.cuh"

#include <algorithm>

namespace ebfi_dp {

int make_view(const ebfi_dp_comm *c, size_t n_floats, View &v)
{
    EBFI_REQUIRE(c != nullptr, "dp: null communicator");
    EBFI_REQUIRE(c->world >= 1 && c->world <= MAX_WORLD && c->rank >= 0 && c->rank < c->world,
                 "dp: bad world / rank (%d / %d; at most %d ranks)", c->world, c->rank, MAX_WORLD);
    EBFI_REQUIRE(c->bytes >= bytes_for(n_floats), "dp: symmetric allocation of %zu bytes < %zu needed for %zu floats",
                 c->bytes, bytes_for(n_floats), n_floats);
    v.world = c->world; v.rank = c->rank;
    for (int q = 0; q < MAX_WORLD; ++q) {
        v.base[q] = static_cast<unsigned char *>(q < c->world ? c->peer_base[q] : nullptr);
        EBFI_REQUIRE(q >= c->world || (v.base[q] && (reinterpret_cast<uintptr_t>(v.base[q]) & 255u) == 0),
                     "dp: peer_base[%d] is null or not 256-byte aligned", q);
    }
    v.cap = (c->bytes - HDR_BYTES) / (2 * MAX_WORLD * sizeof(float));
    return EBFI_OK;
}

namespace {

// a | b (logically concatenated) -> slot [rank] of every rank's symmetric buffer
__global__ void __launch_bounds__(256) dp_publish_kernel(View v, const float *a, size_t na, const float *b, size_t nb)
{
    const unsigned epoch = epoch_of_launch(v);
    for (size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x; e < na + nb; e += (size_t)gridDim.x * blockDim.x)
        push(v, epoch, e, e < na ? a[e] : b[e - na]);
    publish(v, epoch);
}

// a | b <- rank-ordered sums of the epoch published last
__global__ void __launch_bounds__(256) dp_complete_kernel(View v, float *a, size_t na, float *b, size_t nb)
{
    const unsigned epoch = epoch_of_launch(v) - 1u;
    wait_peers(v, epoch);
    for (size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x; e < na + nb; e += (size_t)gridDim.x * blockDim.x) {
        const float s = gather_sum(v, epoch, e);
        if (e < na) a[e] = s; else b[e - na] = s;
    }
}

}  // namespace

int allreduce_sum(cudaStream_t st, const View &v, float *a, size_t na, float *b, size_t nb, int mode)
{
    const size_t n = na + nb;
    if (n == 0) return EBFI_OK;
    const unsigned grid = (unsigned)std::min<size_t>(ebfi::ceil_div(n, (size_t)256), (size_t)ebfi::sm_count() * 4);
    if (mode != 2) {
        dp_publish_kernel<<<grid, 256, 0, st>>>(v, a, na, b, nb);
        EBFI_LAUNCH_OK("dp_publish_kernel");
    }
    if (mode != 1) {
        dp_complete_kernel<<<grid, 256, 0, st>>>(v, a, na, b, nb);
        EBFI_LAUNCH_OK("dp_complete_kernel");
    }
    return EBFI_OK;
}

}  // namespace ebfi_dp

extern "C" {

size_t ebfi_dp_comm_bytes(size_t n_floats) { return ebfi_dp::bytes_for(n_floats); }

static int dp_call(void *stream, const ebfi_dp_comm *comm, float *a, size_t na, float *b, size_t nb, int mode)
{
    EBFI_REQUIRE((a || na == 0) && (b || nb == 0), "dp_allreduce: null pointer");
    ebfi_dp::View v{};
    if (int rc = ebfi_dp::make_view(comm, na + nb, v)) return rc;
    return ebfi_dp::allreduce_sum(ebfi::as_stream(stream), v, a, na, b, nb, mode);
}

int ebfi_dp_allreduce_sum(void *stream, const ebfi_dp_comm *comm, float *a, size_t na, float *b, size_t nb)
{
    return dp_call(stream, comm, a, na, b, nb, 0);
}

int ebfi_dp_publish(void *stream, const ebfi_dp_comm *comm, const float *a, size_t na, const float *b, size_t nb)
{
    return dp_call(stream, comm, const_cast<float *>(a), na, const_cast<float *>(b), nb, 1);
}

int ebfi_dp_complete(void *stream, const ebfi_dp_comm *comm, float *a, size_t na, float *b, size_t nb)
{
    return dp_call(stream, comm, a, na, b, nb, 2);
}

}  // extern "C"
