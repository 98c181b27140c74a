// Peer-memory halo exchange for the row-partitioned multi-GPU levels.
//
// One process per GPU.  Every rank owns a "window" (one cudaMalloc block exported with CUDA IPC and
// mapped by its peers over NVLink 5 / NVSwitch).  A *channel* is one exchange step of the cycle (e.g.
// "halo of x before the level-0 residual"); it has a receive region per parity inside the window and
// one 64-bit flag per source rank.
//
//   push  : the sender packs src[send_idx[i]] straight into the receivers' regions with peer stores
//           (no staging copy, no NCCL), fences system-wide, and the last CTA publishes the channel's
//           sequence number into the receivers' flags;
//   wait  : the receiver spins on its local flags until they reach the sequence number, then unpacks
//           the region into the halo part of the vector.  Interior rows are launched between push and
//           wait, so the NVLink transfer and the neighbours' skew hide behind them.
//
// All state (sequence numbers, CTA counters) lives in device memory, so a cycle is a fixed list of
// kernel launches: it replays from a CUDA graph with no host involvement and no collective call.
// Regions are double-buffered by sequence parity; with one all-to-all dependency per cycle (the
// coarse-level gather) a sender can never be two uses ahead of a receiver on the same channel.
// The spin has a wall-clock timeout (sets the channel's error word instead of hanging the GPU).
#include <string.h>
#include "common.cuh"

namespace mlamg {

constexpr int MAX_PEERS = 16;
constexpr int PEER_THREADS = 256;
constexpr unsigned long long SPIN_TIMEOUT_NS = 20ull * 1000ull * 1000ull * 1000ull;

// channel state words (device, local): [0] seq  [1] push CTA counter  [2] wait CTA counter  [3] error
struct ChannelDev {
    int n_send_peers, n_recv_peers;
    int n_send, n_recv;
    int send_start[MAX_PEERS + 1];
    int recv_start[MAX_PEERS + 1];
    void *send_dst[2][MAX_PEERS];                  // remote region (already offset to my segment), per parity
    unsigned long long *send_flag[MAX_PEERS];      // remote flag slot [channel][me]
    const void *recv_region[2];                    // local region per parity, segments in halo order
    const unsigned long long *recv_flag[MAX_PEERS];   // local flag slot [channel][source]
    unsigned long long *state;
};

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

template <typename T>
__global__ void __launch_bounds__(PEER_THREADS)
channel_push_kernel(const ChannelDev ch, const int *__restrict__ send_idx, const T *__restrict__ src) {
    const unsigned long long seq = *(volatile unsigned long long *)ch.state + 1ull;
    const int par = (int)(seq & 1ull);
    for (long long i = (long long)blockIdx.x * PEER_THREADS + threadIdx.x; i < ch.n_send;
         i += (long long)gridDim.x * PEER_THREADS) {
        int p = 0;
        while (p + 1 < ch.n_send_peers && i >= ch.send_start[p + 1]) p++;
        const T v = src[send_idx ? send_idx[i] : (int)i];
        reinterpret_cast<T *>(ch.send_dst[par][p])[i - ch.send_start[p]] = v;
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned *done = reinterpret_cast<unsigned *>(ch.state + 1);
        const unsigned prev = atomicAdd(done, 1u);
        if (prev == gridDim.x - 1) {          // last CTA: every CTA's peer stores are fenced before its atomicAdd
            *done = 0u;
            __threadfence_system();
            for (int p = 0; p < ch.n_send_peers; p++) st_release_sys(ch.send_flag[p], seq);
        }
    }
}

template <typename T>
__global__ void __launch_bounds__(PEER_THREADS)
channel_wait_kernel(const ChannelDev ch, T *__restrict__ dst) {
    const unsigned long long seq = *(volatile unsigned long long *)ch.state + 1ull;
    const int par = (int)(seq & 1ull);
    const long long per_block = ((long long)ch.n_recv + gridDim.x - 1) / gridDim.x;
    const long long lo = (long long)blockIdx.x * per_block;
    const long long hi = min(lo + per_block, (long long)ch.n_recv);
    if (threadIdx.x < ch.n_recv_peers) {      // one thread per source whose segment meets this CTA's range
        const int p = threadIdx.x;
        const bool empty = ch.recv_start[p + 1] == ch.recv_start[p];   // flag-only segment: CTA 0 waits for it
        if ((ch.recv_start[p] < hi && ch.recv_start[p + 1] > lo) || (empty && blockIdx.x == 0)) {
            const unsigned long long t0 = global_ns();
            while (ld_acquire_sys(ch.recv_flag[p]) < seq) {
                if (global_ns() - t0 > SPIN_TIMEOUT_NS) {
                    atomicExch(ch.state + 3, 1ull + (unsigned long long)p);
                    break;
                }
                __nanosleep(64);
            }
        }
    }
    __syncthreads();
    const T *region = reinterpret_cast<const T *>(ch.recv_region[par]);
    for (long long i = lo + threadIdx.x; i < hi; i += PEER_THREADS) dst[i] = __ldcg(region + i);
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned *done = reinterpret_cast<unsigned *>(ch.state + 2);
        const unsigned prev = atomicAdd(done, 1u);
        if (prev == gridDim.x - 1) {          // every CTA has read `seq` before the last one advances it
            *done = 0u;
            *(volatile unsigned long long *)ch.state = seq;
        }
    }
}

}  // namespace mlamg

struct mlamg_channel {
    mlamg::ChannelDev dev;
};

using namespace mlamg;

extern "C" {

int mlamg_peer_alloc(long long bytes, void **ptr, void *ipc_handle_host) {
    if (bytes <= 0 || !ptr || !ipc_handle_host) return set_error(MLAMG_EINVAL, "peer_alloc: bad arguments");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    void *p = nullptr;
    MLAMG_CUDA(cudaMalloc(&p, (size_t)bytes));
    MLAMG_CUDA(cudaMemset(p, 0, (size_t)bytes));
    MLAMG_CUDA(cudaDeviceSynchronize());
    cudaIpcMemHandle_t h;
    cudaError_t e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) { cudaFree(p); return set_cuda_error(e, __FILE__, __LINE__); }
    memcpy(ipc_handle_host, &h, sizeof(h));
    *ptr = p;
    return MLAMG_OK;
}

int mlamg_peer_open(const void *ipc_handle_host, void **ptr) {
    if (!ipc_handle_host || !ptr) return set_error(MLAMG_EINVAL, "peer_open: bad arguments");
    cudaIpcMemHandle_t h;
    memcpy(&h, ipc_handle_host, sizeof(h));
    MLAMG_CUDA(cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return MLAMG_OK;
}

int mlamg_peer_close(void *ptr) {
    if (ptr) MLAMG_CUDA(cudaIpcCloseMemHandle(ptr));
    return MLAMG_OK;
}

int mlamg_peer_free(void *ptr) {
    if (ptr) MLAMG_CUDA(cudaFree(ptr));
    return MLAMG_OK;
}

int mlamg_channel_create(int n_send_peers, const int *send_counts_host, void *const *send_dst0_host,
                         void *const *send_dst1_host, void *const *send_flag_host, int n_recv_peers,
                         const int *recv_counts_host, const void *recv_region0, const void *recv_region1,
                         void *const *recv_flag_host, void *state, mlamg_channel_t *out) {
    if (!out || !state || n_send_peers < 0 || n_recv_peers < 0 || n_send_peers > MAX_PEERS || n_recv_peers > MAX_PEERS)
        return set_error(MLAMG_EINVAL, "channel_create: bad arguments (at most %d peers)", MAX_PEERS);
    mlamg_channel *c = new mlamg_channel();
    ChannelDev &d = c->dev;
    memset(&d, 0, sizeof(d));
    d.n_send_peers = n_send_peers;
    d.n_recv_peers = n_recv_peers;
    long long tot = 0;
    for (int p = 0; p < n_send_peers; p++) {
        if (send_counts_host[p] < 0 || !send_flag_host[p] ||
            (send_counts_host[p] > 0 && (!send_dst0_host[p] || !send_dst1_host[p]))) {
            delete c;
            return set_error(MLAMG_EINVAL, "channel_create: bad send segment %d", p);
        }
        d.send_start[p] = (int)tot;
        tot += send_counts_host[p];
        d.send_dst[0][p] = send_dst0_host[p];
        d.send_dst[1][p] = send_dst1_host[p];
        d.send_flag[p] = reinterpret_cast<unsigned long long *>(send_flag_host[p]);
    }
    for (int p = n_send_peers; p <= MAX_PEERS; p++) d.send_start[p] = (int)tot;
    d.n_send = (int)tot;
    tot = 0;
    for (int p = 0; p < n_recv_peers; p++) {
        if (recv_counts_host[p] < 0 || !recv_flag_host[p]) {
            delete c;
            return set_error(MLAMG_EINVAL, "channel_create: bad receive segment %d", p);
        }
        d.recv_start[p] = (int)tot;
        tot += recv_counts_host[p];
        d.recv_flag[p] = reinterpret_cast<const unsigned long long *>(recv_flag_host[p]);
    }
    for (int p = n_recv_peers; p <= MAX_PEERS; p++) d.recv_start[p] = (int)tot;
    d.n_recv = (int)tot;
    if (tot > 0 && (!recv_region0 || !recv_region1)) {
        delete c;
        return set_error(MLAMG_EINVAL, "channel_create: missing receive region");
    }
    d.recv_region[0] = recv_region0;
    d.recv_region[1] = recv_region1;
    d.state = reinterpret_cast<unsigned long long *>(state);
    *out = c;
    return MLAMG_OK;
}

int mlamg_channel_destroy(mlamg_channel_t ch) {
    delete ch;
    return MLAMG_OK;
}

int mlamg_channel_push(mlamg_channel_t ch, int dtype, const int *send_idx, const void *src, mlamg_stream_t stream) {
    if (!ch) return set_error(MLAMG_EINVAL, "channel_push: null channel");
    const ChannelDev &d = ch->dev;
    if (d.n_send_peers == 0) return MLAMG_OK;
    cudaStream_t s = as_stream(stream);
    unsigned blocks = d.n_send > 0 ? cdiv(d.n_send, PEER_THREADS) : 1u;   // flag-only segments still publish
    if (blocks > 148u * 4u) blocks = 148u * 4u;
    MLAMG_DISPATCH(dtype, (channel_push_kernel<T><<<blocks, PEER_THREADS, 0, s>>>(d, send_idx, (const T *)src)));
    MLAMG_LAUNCHED();
    return MLAMG_OK;
}

int mlamg_channel_wait(mlamg_channel_t ch, int dtype, void *dst, mlamg_stream_t stream) {
    if (!ch) return set_error(MLAMG_EINVAL, "channel_wait: null channel");
    const ChannelDev &d = ch->dev;
    if (d.n_send_peers == 0 && d.n_recv_peers == 0) return MLAMG_OK;
    cudaStream_t s = as_stream(stream);
    // always launched when the channel is live: it advances the sequence number even with nothing to receive
    unsigned blocks = d.n_recv > 0 ? cdiv(d.n_recv, 4 * PEER_THREADS) : 1u;
    if (blocks > 148u * 2u) blocks = 148u * 2u;
    MLAMG_DISPATCH(dtype, (channel_wait_kernel<T><<<blocks, PEER_THREADS, 0, s>>>(d, (T *)dst)));
    MLAMG_LAUNCHED();
    return MLAMG_OK;
}

}  // extern "C"
