// rt_rng.cuh -- the reference generator (random.h:4-61) in a 28-byte per-path form.
//
// RandomState is 16 x u64 + index (136 B): too big to stream through a wavefront. For the first 15
// draws the generator only ever touches two words, so a path carries
//     cur = the word written by the previous draw (state[n]),
//     x   = the xorshift64* seeding chain value after n+1 rounds (state[n+1] = xs(x) * C is derived on demand),
// and the draw count. From the 16th draw on the ring wraps and older words are needed; that (rare: > 15
// draws per (pixel, sample)) case replays the stream from the seed, which is exact by construction.
#pragma once
#include "rt_common.cuh"

struct PathRng {
    uint64_t cur;
    uint64_t x;
    uint64_t seed;
    uint32_t n;
};

RT_DEVICE uint64_t rng_xs(uint64_t x) {       // random.h:21-23 (>>12, >>25, >>27)
    x ^= x >> 12;
    x ^= x >> 25;
    x ^= x >> 27;
    return x;
}

RT_DEVICE void rng_seed(PathRng &r, uint64_t seed) {   // random.h:9-27
    if (seed == 0) seed = 0x5555555555555555ULL;
    r.seed = seed;
    r.x = rng_xs(seed);
    r.cur = r.x * 2685821657736338717ULL;               // state[0]
    r.n = 0;
}

RT_DEVICE uint64_t rng_step(uint64_t s0, uint64_t s1) { // random.h:35-39 (note &=)
    s1 ^= s1 << 31;
    s1 ^= s1 >> 11;
    s0 &= s0 >> 30;
    return s0 ^ s1;
}

static __device__ __noinline__ uint64_t rng_replay(uint64_t seed, uint32_t n) {   // value of draw index n (0-based)
    uint64_t st[16];
    uint64_t x = seed;
    for (int i = 0; i < 16; ++i) { x = rng_xs(x); st[i] = x * 2685821657736338717ULL; }
    int p = 0;
    uint64_t out = 0;
    for (uint32_t k = 0; k <= n; ++k) {
        uint64_t s0 = st[p];
        p = (p + 1) & 15;
        st[p] = rng_step(s0, st[p]);
        out = st[p];
    }
    return out * 1181783497276652981ULL;
}

RT_DEVICE uint64_t rng_next(PathRng &r) {              // random.h:29-42
    if (r.n < 15u) {
        r.x = rng_xs(r.x);
        uint64_t s1 = r.x * 2685821657736338717ULL;     // untouched state[n + 1]
        r.cur = rng_step(r.cur, s1);
        r.n++;
        return r.cur * 1181783497276652981ULL;
    }
    uint64_t v = rng_replay(r.seed, r.n);
    r.n++;
    return v;
}

RT_DEVICE float rng_float01(PathRng &r) {              // random.h:49-56
    // (float)u64 is round-to-nearest; (float)0xFFFF...F == 2^64, and dividing by 2^64 is exact scaling.
    float f = __ull2float_rn(rng_next(r)) * 5.42101086242752217e-20f;   // 2^-64
    return clampf(f, 0.0f, 1.0f);
}

RT_DEVICE float rng_float11(PathRng &r) { return (rng_float01(r) * 2.0f) - 1.0f; }   // random.h:58-61

RT_DEVICE uint64_t sample_seed(uint64_t base, uint32_t pixel, uint32_t sample) {
    return base ^ ((uint64_t)pixel * RT_SEED_MULT + (uint64_t)sample);
}
