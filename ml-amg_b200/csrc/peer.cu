// Peer-memory halo exchange for the row-partitioned multi-GPU levels.
//
// One process per GPU.  Every rank owns a "window" (one cudaMalloc block exported with CUDA IPC and
// mapped by its peers over NVLink 5 / NVSwitch).  A *channel* is one exchange step of the cycle (e.g.
// "halo of x before the level-0 residual") with a receive region per sequence parity inside the window.
//
//   push   : the sender packs src[send_idx[i]] straight into the receivers' regions with peer stores, each
//            value carrying the channel's sequence tag in the same atomic unit (peer.cuh) — no staging
//            copy, no fence, no flag, no NCCL;
//   consume: either the row-op kernel itself gathers halo columns from the region, spinning on the tag of
//            the few values that have not landed yet (apply.cu, mlamg_channel_rowop), or
//            mlamg_channel_unpack copies the region into a plain vector.
//
// Interior rows are launched between push and consume, so the NVLink latency and the neighbours' skew hide
// behind them.  All state (sequence number, CTA counter) lives in device memory, so a cycle is a fixed list
// of kernel launches: it replays from a CUDA graph with no host involvement and no collective call.
// Regions are double-buffered by sequence parity; with one all-to-all dependency per cycle (the
// coarse-level gather) a sender can never be two uses ahead of a receiver on the same channel.
// Every spin has a wall-clock timeout (sets the channel's error word instead of hanging the GPU).
#include <string.h>
#include "peer.cuh"

namespace mlamg {

constexpr int PEER_THREADS = 256;

template <typename T>
__global__ void __launch_bounds__(PEER_THREADS)
channel_push_kernel(const ChannelDev ch, const int *__restrict__ send_idx, const T *__restrict__ src,
                    const T *__restrict__ scale) {
    const unsigned long long seq = *(volatile unsigned long long *)ch.state + 1ull;
    const int par = (int)(seq & 1ull);
    const unsigned tag = ll_tag(seq);
    for (long long i = (long long)blockIdx.x * PEER_THREADS + threadIdx.x; i < ch.n_send;
         i += (long long)gridDim.x * PEER_THREADS) {
        int p = 0;
        while (p + 1 < ch.n_send_peers && i >= ch.send_start[p + 1]) p++;
        const int j = send_idx ? send_idx[i] : (int)i;
        T v = j >= 0 ? src[j] : (T)0;                   // negative index: padding slot (keeps two ranks in step)
        if (scale && j >= 0) v = scale[j] * v;          // x = dw .* b sent without materialising x first
        ll_store(reinterpret_cast<T *>(ch.send_dst[par][p]), i - ch.send_start[p], v, tag);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned *done = reinterpret_cast<unsigned *>(ch.state + 1);
        const unsigned prev = atomicAdd(done, 1u);
        if (prev == gridDim.x - 1) {          // every CTA has read the sequence number before the last one advances it
            *done = 0u;
            *(volatile unsigned long long *)ch.state = seq;
        }
    }
}

// dst[dst_idx[i]] = region[i] (dst_idx == NULL: dst[i]; negative: padding, waited for but not stored)
template <typename T>
__global__ void __launch_bounds__(PEER_THREADS)
channel_unpack_kernel(const ChannelDev ch, const int *__restrict__ dst_idx, T *__restrict__ dst) {
    const unsigned long long seq = *(volatile unsigned long long *)ch.state;   // already advanced by this use's push
    const T *region = reinterpret_cast<const T *>(ch.recv_region[(int)(seq & 1ull)]);
    const unsigned tag = ll_tag(seq);
    for (long long i = (long long)blockIdx.x * PEER_THREADS + threadIdx.x; i < ch.n_recv;
         i += (long long)gridDim.x * PEER_THREADS) {
        const T v = ll_load(region, i, tag, ch.state);
        const int j = dst_idx ? dst_idx[i] : (int)i;
        if (j >= 0) dst[j] = v;
    }
}

}  // namespace mlamg

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

int mlamg_channel_slot_bytes(int dtype) { return dtype == MLAMG_F32 ? 8 : 16; }

int mlamg_channel_create(int n_send_peers, const int *send_counts_host, void *const *send_dst0_host,
                         void *const *send_dst1_host, int n_recv, const void *recv_region0,
                         const void *recv_region1, void *state, mlamg_channel_t *out) {
    if (!out || !state || n_send_peers < 0 || n_send_peers > MAX_PEERS || n_recv < 0)
        return set_error(MLAMG_EINVAL, "channel_create: bad arguments (at most %d peers)", MAX_PEERS);
    mlamg_channel *c = new mlamg_channel();
    ChannelDev &d = c->dev;
    memset(&d, 0, sizeof(d));
    d.n_send_peers = n_send_peers;
    d.n_recv_peers = n_recv > 0 ? 1 : 0;
    long long tot = 0;
    for (int p = 0; p < n_send_peers; p++) {
        if (send_counts_host[p] <= 0 || !send_dst0_host[p] || !send_dst1_host[p]) {
            delete c;
            return set_error(MLAMG_EINVAL, "channel_create: bad send segment %d", p);
        }
        d.send_start[p] = (int)tot;
        tot += send_counts_host[p];
        d.send_dst[0][p] = send_dst0_host[p];
        d.send_dst[1][p] = send_dst1_host[p];
    }
    for (int p = n_send_peers; p <= MAX_PEERS; p++) d.send_start[p] = (int)tot;
    d.n_send = (int)tot;
    d.n_recv = n_recv;
    if (n_recv > 0 && (!recv_region0 || !recv_region1)) {
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

int mlamg_channel_push(mlamg_channel_t ch, int dtype, const int *send_idx, const void *src, const void *scale,
                       mlamg_stream_t stream) {
    if (!ch) return set_error(MLAMG_EINVAL, "channel_push: null channel");
    const ChannelDev &d = ch->dev;
    if (d.n_send == 0 && d.n_recv == 0) return MLAMG_OK;
    cudaStream_t s = as_stream(stream);
    // always launched on a live channel: it advances the sequence number even with nothing to send
    unsigned blocks = d.n_send > 0 ? cdiv(d.n_send, 2 * PEER_THREADS) : 1u;
    if (blocks > 148u) blocks = 148u;
    MLAMG_DISPATCH(dtype, (channel_push_kernel<T><<<blocks, PEER_THREADS, 0, s>>>(d, send_idx, (const T *)src, (const T *)scale)));
    MLAMG_LAUNCHED();
    return MLAMG_OK;
}

int mlamg_channel_unpack(mlamg_channel_t ch, int dtype, const int *dst_idx, void *dst, mlamg_stream_t stream) {
    if (!ch) return set_error(MLAMG_EINVAL, "channel_unpack: null channel");
    const ChannelDev &d = ch->dev;
    if (d.n_recv == 0) return MLAMG_OK;
    cudaStream_t s = as_stream(stream);
    unsigned blocks = cdiv(d.n_recv, 2 * PEER_THREADS);
    if (blocks > 148u) blocks = 148u;
    MLAMG_DISPATCH(dtype, (channel_unpack_kernel<T><<<blocks, PEER_THREADS, 0, s>>>(d, dst_idx, (T *)dst)));
    MLAMG_LAUNCHED();
    return MLAMG_OK;
}

}  // extern "C"
