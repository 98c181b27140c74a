// Peer-memory exchange channel: device-side description shared by peer.cu (push / unpack kernels) and
// apply.cu (row-op kernels that gather halo values straight from the receive region).
//
// Wire format ("LL", value + tag in one store, as NCCL's low-latency protocol): every element travels with
// the channel's sequence tag in the same naturally-atomic 8-byte unit, so there is no fence, no flag and no
// second round trip — a value is valid exactly when its tag equals the receiver's current sequence number.
//   f64: 16 B  { lo32, tag, hi32, tag }      f32: 8 B  { bits, tag }
#pragma once
#include "common.cuh"

namespace mlamg {

constexpr int MAX_PEERS = 16;

// channel state words (device, local): [0] sequence number  [1] push CTA counter  [2] spare  [3] error
struct ChannelDev {
    int n_send_peers, n_recv_peers;
    int n_send, n_recv;
    int send_start[MAX_PEERS + 1];
    void *send_dst[2][MAX_PEERS];      // remote region start of my segment, per sequence parity
    const void *recv_region[2];        // local regions, segments back to back in halo order
    unsigned long long *state;
};

// what a consumer kernel needs to read halo values in place
struct HaloLL {
    int n_own;                          // columns >= n_own are halo slots
    const void *region[2];
    unsigned long long *state;
};

__device__ __forceinline__ unsigned ll_tag(unsigned long long seq) { return (unsigned)(seq % 0xFFFFFFFFull) + 1u; }

__device__ __forceinline__ unsigned long long ll_global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

constexpr unsigned long long LL_TIMEOUT_NS = 20ull * 1000ull * 1000ull * 1000ull;

__device__ __forceinline__ void ll_store(double *slot_base, long long i, double v, unsigned tag) {
    const unsigned lo = (unsigned)__double2loint(v), hi = (unsigned)__double2hiint(v);
    uint4 *p = reinterpret_cast<uint4 *>(slot_base) + i;
    asm volatile("st.volatile.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(lo), "r"(tag), "r"(hi), "r"(tag) : "memory");
}
__device__ __forceinline__ void ll_store(float *slot_base, long long i, float v, unsigned tag) {
    uint2 *p = reinterpret_cast<uint2 *>(slot_base) + i;
    asm volatile("st.volatile.global.v2.u32 [%0], {%1, %2};" ::"l"(p), "r"(__float_as_uint(v)), "r"(tag) : "memory");
}

// Element i of the region is valid once it carries `tag`.  Fast path (inlined into the consuming row-op kernel): one
// volatile load and a compare — in the overlapped cycle the value has almost always landed.  Slow path (out of line, so
// that the spin state costs the row-op kernels no registers): spin with a wall-clock timeout.  A timeout sets the
// channel's error word [3] and returns what is there; once the word is set every later wait on the channel gives up at
// once (one dead peer costs one timeout per kernel, not one per halo element), and the host side refuses to use the
// result (DistHierarchy checks the word after every solve / at every synchronisation point).
static __device__ __noinline__ uint4 ll_spin16(const uint4 *p, unsigned tag, unsigned long long *state) {
    uint4 v;
    unsigned long long t0 = 0;
    for (unsigned spins = 0;; spins++) {
        asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
        if (v.y == tag && v.w == tag) break;
#ifndef MLAMG_LL_TIGHT_SPIN
        // back off: a warp that polls in a tight loop takes issue slots, LSU bandwidth and power from the interior-row
        // kernel running beside it (the boundary rows are launched before their halo can have arrived)
        if (spins >= 4u) __nanosleep(spins < 64u ? 100u : 400u);
#endif
        if ((spins & 1023u) == 0u) {
            if (*(volatile unsigned long long *)(state + 3) != 0ull) break;      // the channel already timed out
            if (spins == 0u) t0 = ll_global_ns();
            else if (ll_global_ns() - t0 > LL_TIMEOUT_NS) {
                atomicExch(state + 3, 1ull);
                break;
            }
        }
    }
    return v;
}
static __device__ __noinline__ uint2 ll_spin8(const uint2 *p, unsigned tag, unsigned long long *state) {
    uint2 v;
    unsigned long long t0 = 0;
    for (unsigned spins = 0;; spins++) {
        asm volatile("ld.volatile.global.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p) : "memory");
        if (v.y == tag) break;
#ifndef MLAMG_LL_TIGHT_SPIN
        if (spins >= 4u) __nanosleep(spins < 64u ? 100u : 400u);
#endif
        if ((spins & 1023u) == 0u) {
            if (*(volatile unsigned long long *)(state + 3) != 0ull) break;
            if (spins == 0u) t0 = ll_global_ns();
            else if (ll_global_ns() - t0 > LL_TIMEOUT_NS) {
                atomicExch(state + 3, 1ull);
                break;
            }
        }
    }
    return v;
}

__device__ __forceinline__ double ll_load(const double *slot_base, long long i, unsigned tag, unsigned long long *state) {
    const uint4 *p = reinterpret_cast<const uint4 *>(slot_base) + i;
    uint4 v;
    asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    if (!(v.y == tag && v.w == tag)) v = ll_spin16(p, tag, state);
    return __hiloint2double((int)v.z, (int)v.x);
}
__device__ __forceinline__ float ll_load(const float *slot_base, long long i, unsigned tag, unsigned long long *state) {
    const uint2 *p = reinterpret_cast<const uint2 *>(slot_base) + i;
    uint2 v;
    asm volatile("ld.volatile.global.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p) : "memory");
    if (v.y != tag) v = ll_spin8(p, tag, state);
    return __uint_as_float(v.x);
}

}  // namespace mlamg

struct mlamg_channel {
    mlamg::ChannelDev dev;
};
