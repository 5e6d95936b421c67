// Philox4x32-10 (Salmon et al., SC'11) — device copy of the build's draw specification.
// key = (seed, global_env_id), counter = (tick, stream, index, sub); see DESIGN.md §RNG.
#pragma once
#include <stdint.h>

namespace maexp {

enum Stream : uint32_t {
    kStreamOrder = 0,     // abstract_trainer.py:66 move order, Fisher-Yates draws
    kStreamAction = 1,    // abstract_agent.py:31-34 random policy
    kStreamRespawn = 2,   // default_environment.py:137-138 during step
    kStreamInitPerm = 3,  // default_environment.py:97-98
    kStreamDest = 4,      // three_rooms_map.py:52-55
    kStreamObjects = 5,   // default_environment.py:123-148 during reset
    kStreamNoise = 6      // noises/*.py: gaussian, one call per 4 elements; uniform (flip), 16-bit draws, one call per 8
};

__device__ __forceinline__ uint4 philox4x32_10(uint32_t k0, uint32_t k1, uint32_t c0, uint32_t c1, uint32_t c2,
                                               uint32_t c3) {
    constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t h0 = __umulhi(M0, c0), l0 = M0 * c0;
        const uint32_t h1 = __umulhi(M1, c2), l1 = M1 * c2;
        c0 = h1 ^ c1 ^ k0;
        c1 = l1;
        c2 = h0 ^ c3 ^ k1;
        c3 = l0;
        k0 += W0;
        k1 += W1;
    }
    return make_uint4(c0, c1, c2, c3);
}

__device__ __forceinline__ uint32_t pick(const uint4 &w, uint32_t i) {
    return i == 0 ? w.x : (i == 1 ? w.y : (i == 2 ? w.z : w.w));
}

// floor(u / 2^32 * n) == python int(random() * n) for random() = u / 2**32
__device__ __forceinline__ uint32_t scale(uint32_t u, uint32_t n) { return __umulhi(u, n); }

__device__ __forceinline__ float unit_float(uint32_t u) { return (float)(u >> 8) * (1.0f / 16777216.0f); }

}  // namespace maexp
