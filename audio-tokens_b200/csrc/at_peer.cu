// libat_b200: the one exchange step of the path -- the per-iteration sum of the k-means accumulators of all ranks
// ("one all-reduce of the K x 64 sums and K counts per Lloyd iteration") -- as ONE kernel over peer
// memory instead of a library collective.
//
// Each rank owns a window of device memory [flags | buffer 0 | buffer 1] that every other rank of the node maps through
// CUDA IPC (NVLink / NVSwitch peer access).  at_kmeans_accumulate writes the rank's exact int64 sums / counts / sum of
// squares straight into buffer (seq & 1) of its own window; k_peer_reduce then
//   1. publishes the iteration number in the `flags[rank]` slot of every peer's window (system-scope release),
//   2. waits until the flags of all ranks in its own window have reached the iteration (acquire),
//   3. adds the W windows, read over NVLink in rank order, into the local total that at_kmeans_finalize consumes.
// The sums are integers, so the total is bit-identical on every rank and for every rank count; W x 532 KB cross the links
// per rank and iteration (K = 1024), a few microseconds, with no host involvement and no second launch for the collective.
// Double buffering makes the window safe to overwrite: a rank can only reach iteration i + 2 after every peer has
// published i + 1, i.e. after every peer has finished reading iteration i.
//
// A rank that never arrives would leave the others spinning: the wait gives up after ~4 s of SM clock and raises an error
// word that at_peer_status reports (the kernel then still terminates, with a wrong total, instead of hanging the GPU).
#include "at_common.cuh"

#include <new>
#include <string.h>

struct at_peer {
    int rank = 0, world = 1;
    int64_t words = 0;
    unsigned char *window = nullptr;            // local: flags (AT_PEER_FLAG_BYTES) | buffer 0 | buffer 1
    unsigned char *peer_window[AT_PEER_MAX] = {};   // mapped windows (peer_window[rank] = window)
    unsigned long long **d_flag_slots = nullptr;    // device array [world]: &flags_of_peer_r[rank]
    const long long **d_bufs = nullptr;             // device array [2 * world]: buffer b of rank r at [b * world + r]
    long long *total = nullptr;                     // local sum of all ranks (words)
    unsigned int *d_error = nullptr;
    unsigned long long seq = 0;                     // iterations completed
    bool connected = false;
};

namespace at {

constexpr size_t PEER_FLAG_BYTES = 1024;

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// grid: any; block 256.  flags: this rank's own flag array (world slots).
__global__ void __launch_bounds__(256) k_peer_reduce(unsigned long long *const *__restrict__ flag_slots,
                                                     const unsigned long long *__restrict__ flags,
                                                     const long long *const *__restrict__ bufs, int world, int rank,
                                                     unsigned long long seq, int64_t words, long long *__restrict__ total,
                                                     unsigned int *__restrict__ error) {
    if (threadIdx.x < world) {
        // block 0 publishes; every block waits for the peers on its own (no block of this grid waits for another one)
        if (blockIdx.x == 0) {
            __threadfence_system();   // the accumulate kernels' writes to this rank's buffer are visible before the flag
            st_release_sys(flag_slots[threadIdx.x], seq);
        }
        const long long t0 = clock64();
        bool ok = true;
        while (ld_acquire_sys(flags + threadIdx.x) < seq) {
            if (clock64() - t0 > (8LL << 30)) {   // ~4 s of SM clock: a peer is gone
                ok = false;
                break;
            }
            __nanosleep(64);
        }
        if (!ok) atomicExch(error, 1u + (unsigned int)threadIdx.x);
    }
    __syncthreads();
    const long long *const *src = bufs + (size_t)(seq & 1ULL) * world;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < words; i += (int64_t)gridDim.x * blockDim.x) {
        long long s = 0;
        for (int r = 0; r < world; r++) s += __ldcv(src[r] + i);   // peer loads bypass L2; keep them out of a stale L1 too
        total[i] = s;
    }
}

}  // namespace at

using namespace at;

extern "C" {

int at_peer_create(int rank, int world, int64_t words, at_peer **out) {
    AT_REQUIRE(out && world >= 1 && world <= AT_PEER_MAX && rank >= 0 && rank < world && words > 0,
               "at_peer_create: bad arguments");
    at_peer *p = new (std::nothrow) at_peer();
    if (!p) return AT_ERR_NOMEM;
    p->rank = rank, p->world = world, p->words = words;
    const size_t bytes = PEER_FLAG_BYTES + 2 * sizeof(long long) * (size_t)words;
    cudaError_t e = cudaMalloc(&p->window, bytes);
    if (e == cudaSuccess) e = cudaMemset(p->window, 0, bytes);
    if (e == cudaSuccess) e = cudaMalloc(&p->total, sizeof(long long) * (size_t)words);
    if (e == cudaSuccess) e = cudaMalloc(&p->d_flag_slots, sizeof(void *) * (size_t)world);
    if (e == cudaSuccess) e = cudaMalloc(&p->d_bufs, sizeof(void *) * 2 * (size_t)world);
    if (e == cudaSuccess) e = cudaMalloc(&p->d_error, sizeof(unsigned int));
    if (e == cudaSuccess) e = cudaMemset(p->d_error, 0, sizeof(unsigned int));
    if (e != cudaSuccess) {
        set_error("at_peer_create: %s", cudaGetErrorString(e));
        at_peer_destroy(p);
        return AT_ERR_CUDA;
    }
    p->peer_window[rank] = p->window;
    *out = p;
    return AT_OK;
}

int at_peer_export(const at_peer *p, void *handle64) {
    AT_REQUIRE(p && handle64, "at_peer_export: bad arguments");
    static_assert(sizeof(cudaIpcMemHandle_t) == AT_PEER_HANDLE_BYTES, "handle size");
    cudaIpcMemHandle_t h;
    AT_CUDA_OK(cudaIpcGetMemHandle(&h, p->window));
    memcpy(handle64, &h, sizeof(h));
    return AT_OK;
}

int at_peer_import(at_peer *p, int peer_rank, const void *handle64) {
    AT_REQUIRE(p && handle64 && peer_rank >= 0 && peer_rank < p->world, "at_peer_import: bad arguments");
    if (peer_rank == p->rank) return AT_OK;
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, sizeof(h));
    void *ptr = nullptr;
    AT_CUDA_OK(cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess));
    p->peer_window[peer_rank] = (unsigned char *)ptr;
    return AT_OK;
}

// Call after every peer has been imported (and after a host-side barrier: every rank's window exists and is zeroed).
int at_peer_connect(at_peer *p) {
    AT_REQUIRE(p, "at_peer_connect: bad arguments");
    unsigned long long *slots[AT_PEER_MAX];
    const long long *bufs[2 * AT_PEER_MAX];
    for (int r = 0; r < p->world; r++) {
        AT_REQUIRE(p->peer_window[r], "at_peer_connect: rank %d has not been imported", r);
        slots[r] = reinterpret_cast<unsigned long long *>(p->peer_window[r]) + p->rank;
        for (int b = 0; b < 2; b++)
            bufs[b * p->world + r] = reinterpret_cast<const long long *>(p->peer_window[r] + PEER_FLAG_BYTES) + (size_t)b * p->words;
    }
    AT_CUDA_OK(cudaMemcpy(p->d_flag_slots, slots, sizeof(void *) * (size_t)p->world, cudaMemcpyHostToDevice));
    AT_CUDA_OK(cudaMemcpy(p->d_bufs, bufs, sizeof(void *) * 2 * (size_t)p->world, cudaMemcpyHostToDevice));
    p->connected = true;
    return AT_OK;
}

int at_peer_destroy(at_peer *p) {
    if (!p) return AT_OK;
    cudaDeviceSynchronize();
    for (int r = 0; r < p->world; r++)
        if (r != p->rank && p->peer_window[r]) cudaIpcCloseMemHandle(p->peer_window[r]);
    cudaFree(p->window), cudaFree(p->total), cudaFree(p->d_flag_slots), cudaFree(p->d_bufs), cudaFree(p->d_error);
    delete p;
    return AT_OK;
}

int64_t *at_peer_local_buffer(at_peer *p) {
    if (!p) return nullptr;
    // buffer of the NEXT iteration: (seq + 1) & 1
    return reinterpret_cast<int64_t *>(p->window + PEER_FLAG_BYTES) + (size_t)((p->seq + 1) & 1ULL) * p->words;
}

const int64_t *at_peer_total(const at_peer *p) { return p ? reinterpret_cast<const int64_t *>(p->total) : nullptr; }

int at_peer_reduce(at_peer *p, void *stream) {
    AT_REQUIRE(p && p->connected, "at_peer_reduce: connect the peers first");
    p->seq++;
    int blocks = (int)ceil_div(p->words, 256 * 4);
    if (blocks > sm_count()) blocks = sm_count();
    if (blocks < 1) blocks = 1;
    k_peer_reduce<<<blocks, 256, 0, (cudaStream_t)stream>>>(
        p->d_flag_slots, reinterpret_cast<const unsigned long long *>(p->window), p->d_bufs, p->world, p->rank, p->seq,
        p->words, p->total, p->d_error);
    AT_LAUNCH_OK();
    return AT_OK;
}

int at_peer_status(at_peer *p, void *stream) {
    AT_REQUIRE(p, "at_peer_status: bad arguments");
    unsigned int err = 0;
    AT_CUDA_OK(cudaStreamSynchronize((cudaStream_t)stream));
    AT_CUDA_OK(cudaMemcpy(&err, p->d_error, sizeof(err), cudaMemcpyDeviceToHost));
    if (err) {
        set_error("at_peer: rank %d gave up waiting for rank %u (iteration %llu)", p->rank, err - 1, p->seq);
        return AT_ERR_CUDA;
    }
    return AT_OK;
}

}  // extern "C"
