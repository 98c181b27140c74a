// Host-side: the seeds of ns/lib/graph.py:229-231, `RandomState(rand).permutation(N)[:num_seeds]`, without numpy.
//
// The reference draws its Lloyd seeds from numpy's LEGACY generator: MT19937 seeded with init_genrand(seed), then a
// Fisher-Yates shuffle of arange(N) running from the END of the array (`for i in reversed(range(1, n)): j =
// random_interval(i); swap(x[i], x[j])`), `random_interval` being masked rejection sampling on 32-bit draws.  The head
// of the permutation therefore depends on all N-1 draws: the sequence is inherently serial and stays on the host.
// numpy's own loop moves 8-byte items with three memcpy calls per step (0.34-0.43 s at N = 16.7 M); this one shuffles
// int32 ids, draws the indices in batches and prefetches the swap targets (the random access to x[j] is what the loop
// waits for).  Bit-identical to numpy (tests/test_host_logic.py compares against numpy for many (seed, N)).
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
#include "common.cuh"

namespace mlamg {

struct Mt19937 {
    uint32_t key[624];
    int pos;
    explicit Mt19937(uint32_t seed) {
        for (int i = 0; i < 624; i++) {
            key[i] = seed;
            seed = 1812433253u * (seed ^ (seed >> 30)) + (uint32_t)i + 1u;
        }
        pos = 624;
    }
    void refill() {
        const uint32_t UPPER = 0x80000000u, LOWER = 0x7fffffffu, MAG = 0x9908b0dfu;
        int i;
        uint32_t y;
        for (i = 0; i < 624 - 397; i++) {
            y = (key[i] & UPPER) | (key[i + 1] & LOWER);
            key[i] = key[i + 397] ^ (y >> 1) ^ ((y & 1u) ? MAG : 0u);
        }
        for (; i < 623; i++) {
            y = (key[i] & UPPER) | (key[i + 1] & LOWER);
            key[i] = key[i + (397 - 624)] ^ (y >> 1) ^ ((y & 1u) ? MAG : 0u);
        }
        y = (key[623] & UPPER) | (key[0] & LOWER);
        key[623] = key[396] ^ (y >> 1) ^ ((y & 1u) ? MAG : 0u);
        pos = 0;
    }
    inline uint32_t next() {
        if (pos == 624) refill();
        uint32_t y = key[pos++];
        y ^= (y >> 11);
        y ^= (y << 7) & 0x9d2c5680u;
        y ^= (y << 15) & 0xefc60000u;
        y ^= (y >> 18);
        return y;
    }
};

}  // namespace mlamg

using namespace mlamg;

extern "C" {

int mlamg_legacy_permutation_head(unsigned seed, long long n, long long k, int *out_host) {
    if (n < 0 || k < 0 || k > n || n > 0x7fffffffLL) return set_error(MLAMG_EINVAL, "legacy_permutation_head: bad n/k");
    if (n == 0 || k == 0) return MLAMG_OK;
    std::vector<int> x((size_t)n);
    for (long long i = 0; i < n; i++) x[(size_t)i] = (int)i;
    Mt19937 rng(seed);
    constexpr int B = 32;
    uint32_t jb[B];
    long long i = n - 1;
    while (i >= 1) {
        const int cnt = (int)(i < B ? i : B);
        for (int t = 0; t < cnt; t++) {                // j for i, i-1, ..., in draw order
            const uint32_t mx = (uint32_t)(i - t);
            uint32_t mask = mx;
            mask |= mask >> 1; mask |= mask >> 2; mask |= mask >> 4; mask |= mask >> 8; mask |= mask >> 16;
            uint32_t v;
            while ((v = (rng.next() & mask)) > mx) {}
            jb[t] = v;
            __builtin_prefetch(&x[v], 1, 0);
        }
        for (int t = 0; t < cnt; t++) {
            const size_t a = (size_t)(i - t), b = jb[t];
            const int tmp = x[b];
            x[b] = x[a];
            x[a] = tmp;
        }
        i -= cnt;
    }
    memcpy(out_host, x.data(), (size_t)k * sizeof(int));
    return MLAMG_OK;
}

}  // extern "C"
