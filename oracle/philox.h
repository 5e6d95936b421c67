/* Philox4x32-10 counter-based generator — shared RNG *specification* of the build.
 *
 * TEST INFRASTRUCTURE (oracle side).  The CUDA product path has its own device copy in
 * maexp_b200/csrc/philox.cuh; both follow the published algorithm (Salmon et al.,
 * "Parallel random numbers: as easy as 1, 2, 3", SC'11) and are cross-checked in
 * tests/test_philox.py against the paper's known-answer vectors.
 *
 * The reference (Yoshi-0921/MAEXP) draws every random number from three global,
 * interleaved generators (core/utils/seed.py:14-17); bit-exact parity is only defined
 * with injected draws.  The build therefore keys every draw as
 *     key     = (seed, global_env_id)
 *     counter = (tick, stream, index, sub)
 * so that the oracle harness can hand the very same numbers to the live reference
 * through tape objects (oracle/ref_runner.py) and the kernels can generate them on chip.
 */
#ifndef MAEXP_ORACLE_PHILOX_H
#define MAEXP_ORACLE_PHILOX_H
#include <stdint.h>

enum {
    MAEXP_STREAM_ORDER = 0,       /* per-step move order: Fisher-Yates draws           */
    MAEXP_STREAM_ACTION = 1,      /* random policy: one draw per agent                 */
    MAEXP_STREAM_RESPAWN = 2,     /* default_environment.py:137-138, attempts in a step */
    MAEXP_STREAM_INIT_PERM = 3,   /* default_environment.py:97-98                      */
    MAEXP_STREAM_DEST = 4,        /* three_rooms_map.py:52-55                          */
    MAEXP_STREAM_OBJECTS = 5,     /* default_environment.py:123-148 at reset           */
    MAEXP_STREAM_NOISE = 6        /* noises/*.py: gaussian, one call per 4 elements; uniform (flip), 16-bit draws, one call per 8 */
};

static inline void philox4x32_10(uint32_t k0, uint32_t k1, uint32_t c0, uint32_t c1,
                                 uint32_t c2, uint32_t c3, uint32_t out[4]) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)M0 * c0, p1 = (uint64_t)M1 * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += W0; k1 += W1;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

/* draw number `d` of a stream: Philox call d>>2, word d&3 */
static inline uint32_t maexp_draw(uint32_t seed, uint32_t env, uint32_t tick, uint32_t stream,
                                  uint32_t d, uint32_t sub) {
    uint32_t o[4];
    philox4x32_10(seed, env, tick, stream, d >> 2, sub, o);
    return o[d & 3];
}

/* floor(u/2^32 * n): identical to python `int(random() * n)` for random() = u / 2**32 */
static inline uint32_t maexp_scale(uint32_t u, uint32_t n) {
    return (uint32_t)(((uint64_t)u * n) >> 32);
}
#endif
