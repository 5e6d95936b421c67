// Kernel template of the batched gridworld step + observation builder (see env.cu for the
// C-ABI and the design notes).  Instantiated per visible_range in env_v5.cu / env_v7.cu / env_v9.cu.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/maexp_b200.h"
#include "philox.cuh"

namespace maexp {

constexpr int kWarpsPerCta = 4;
constexpr int kRespawnCap = 1 << 16;  // the reference loop is unbounded

struct EnvParams {
    // shape
    int R;  // V / 2: static bit rows are stored pre-shifted by R
    int E, N, T, X, Y, P, C, K, slot_floats, qps, wpa, ws, gs, as, area_bit0, M;
    int q_str, q_tail, sub_str, sub_tail;  // phase-B quad classes (area channel offset, tail padding)
    uint32_t qps_magic;
    int agent_view, object_view, merged, noise, reward_mode, shared_max_is_mean;
    int fast3;  // 3 channels (total agents, all objects, area) and V <= 7: compile-time bit assembly
    int L, num_objects, keep, shuffle_init, auto_reset, random_dest, env_kind;
    int n_init;  // len(config.init_xys) >= N: the start-cell order has one entry per init_xys row
    int stagger_ns, stagger_div;  // experiment knob (MAEXP_STAGGER_NS): CTAs of every other resident slot start late
    uint32_t seed, gid0, tick;
    int force_reset;
    const uint8_t *env_mask;
    // static tables, bit rows pre-shifted left by R = V/2 (see the padded shared-memory layout); read through the
    // read-only L1 path: 6.6 KB shared by every CTA on an SM, no per-CTA staging
    const uint32_t *wall, *noise_area, *area, *dest;
    const uint64_t *los;
    // state
    uint16_t *pos;
    uint32_t *obj;
    int32_t *step;
    uint8_t *pending;
    int32_t *counters;
    uint8_t *init_perm;
    uint8_t *dest_sel;
    uint16_t *past;
    unsigned long long *stats;  // [0] env-steps, [1] resets
    uint32_t tx_magic;          // ceil(2^32 / (T*X))
    int32_t *error_flag;
    uint32_t *queue;  // env_step_lanes_kernel: [0] work items handed out past the first static round, [1] warps that left
    // dynamic_simple_map.py: world steps taken since creation (the obstacle phase; survives resets), or nullptr
    uint32_t *wsteps;
    // io
    const int8_t *actions;
    const uint8_t *order;
    const float *noise_tape;
    maexp_obs_out out;
    // small per-agent config
    uint16_t init_xy[MAEXP_MAX_AGENTS];
    int8_t task_order[MAEXP_MAX_AGENTS][MAEXP_MAX_TYPES];
    int8_t no_task[MAEXP_MAX_AGENTS];
    float flip_range[MAEXP_MAX_RINGS], std_dist[MAEXP_MAX_RINGS], noise_thr[MAEXP_MAX_RINGS];
};

#ifndef MAEXP_ST
#define MAEXP_ST 0
#endif
__device__ __forceinline__ void store_quad(float4 *ptr, const float4 &v) {
#if MAEXP_ST == 0
    __stcs(ptr, v);   // st.global.cs: streaming, evict-first
#elif MAEXP_ST == 1
    *ptr = v;         // default write-back policy
#elif MAEXP_ST == 3
    if (__float_as_uint(v.x) == 0x12345678u) __stcs(ptr, v);  // experiment: everything but the store (never taken)
#else
    __stcg(ptr, v);   // st.global.cg: cache in L2 only
#endif
}
#define MAEXP_STORE 1

template <int V>
struct Win {  // V*V-bit window mask, bit a*V+b  (a: x offset, b: y offset)
    using T = uint64_t;
};
template <>
struct Win<9> {
    using T = unsigned __int128;
};

__device__ __forceinline__ uint32_t nib(uint64_t v, int i) { return (uint32_t)(v >> (4 * i)) & 0xFu; }
__device__ __forceinline__ uint64_t nib_swap(uint64_t v, int i, int j) {
    const uint64_t d = ((v >> (4 * i)) ^ (v >> (4 * j))) & 0xFull;
    return v ^ (d << (4 * i)) ^ (d << (4 * j));
}

// Fisher-Yates over `n` nibbles with stream draws d0, d0+1, ... (oracle/philox.h tape_shuffle)
__device__ __forceinline__ uint64_t fisher_yates(uint64_t v, int n, uint32_t seed, uint32_t gid, uint32_t tick,
                                                 uint32_t stream, uint32_t &d) {
    uint4 w = make_uint4(0, 0, 0, 0);
    uint32_t have = 0xFFFFFFFFu;
    for (int k = n - 1; k >= 1; --k) {
        if ((d >> 2) != have) {
            have = d >> 2;
            w = philox4x32_10(seed, gid, tick, stream, have, 0);
        }
        const uint32_t j = scale(pick(w, d & 3), (uint32_t)(k + 1));
        ++d;
        v = nib_swap(v, k, (int)j);
    }
    return v;
}

// Appends V*V-bit channel masks to a bit stream in output order.  Pending bits live in a
// register accumulator; completed 32-bit words are written once (no shared-memory RMW chain).
// Word k of an env lives at out[k * STRIDE] (STRIDE = envs per warp): the lanes of a warp (one env each) write
// consecutive banks, and the expansion phase reads four envs' words with one 128-bit load.
template <int V, int STRIDE>
struct BitStream {
    unsigned __int128 acc = 0;
    int nbits = 0;
    uint32_t *out;
    __device__ __forceinline__ explicit BitStream(uint32_t *o) : out(o) {}
    __device__ __forceinline__ void put(typename Win<V>::T m) {
        acc |= (unsigned __int128)m << nbits;  // nbits < 32, V*V <= 81
        nbits += V * V;
        while (nbits >= 32) {
            *out = (uint32_t)acc;
            out += STRIDE;
            acc >>= 32;
            nbits -= 32;
        }
    }
    __device__ __forceinline__ void finish() {  // flush the tail and one zero word of slack
        out[0] = (uint32_t)acc;
        out[STRIDE] = 0u;
        out[2 * STRIDE] = 0u;
    }
};

// Phase B hot path (no noise, every env of the group active, one output slot): expands a full group's bit words of one
// agent into fp32 and streams them out.  A lane owns quad q (4 consecutive floats) of every env it visits, so the shift
// and the quad's class are per-lane constants; the envs' words come four at a time (128-bit shared loads) and with SLOT
// (floats per agent slot) known at compile time every store address is base + immediate.
// The expansion is arithmetic (no table lookup: the shared-memory / L1 pipe is this kernel's busiest unit and already
// carries the stores): element j of the quad is bit_j - sub_j, sub_j = 1 for elements of the area channel
// (observation_handler.py:107-108), i.e. ((bit_j ^ sub_j) ? (sub_j ? -1.0f : 1.0f) : 0.0f); with t = (x ^ sub) & 2^j in
// {0, 2^j} the float's bit pattern is t * (pattern >> j) — one LOP3 and one IMAD per element (1.0f and -1.0f have 23 zero
// low bits, so the shift loses nothing).
template <int SLOT, int EPW>
__device__ __forceinline__ void expand_group(const uint32_t *s_w, float *dst, int lane, int slot_rt, int q_str, int q_tail,
                                             int sub_str, int sub_tail) {
    const int slot = SLOT ? SLOT : slot_rt, qps = slot >> 2;
    const int nfull = qps >> 5, rem = qps & 31;
    struct QuadConst { uint32_t sub, h0, h1, h2, h3; int sh; };
    auto consts = [&](int q) {
        QuadConst c;
        const int cls = q > q_tail ? 0 : ((q >= q_str) + (q > q_str) + (q == q_tail));
        c.sub = cls == 0 ? 0u : (cls == 1 ? (uint32_t)sub_str : (cls == 2 ? 15u : (uint32_t)sub_tail));
        c.h0 = (0x3f800000u | ((c.sub & 1u) << 31));
        c.h1 = (0x3f800000u | (((c.sub >> 1) & 1u) << 31)) >> 1;
        c.h2 = (0x3f800000u | (((c.sub >> 2) & 1u) << 31)) >> 2;
        c.h3 = (0x3f800000u | (((c.sub >> 3) & 1u) << 31)) >> 3;
        c.sh = (q & 7) * 4;
        return c;
    };
    auto expand = [](const QuadConst &c, uint32_t w) {
        const uint32_t x = w >> c.sh;
        return make_float4(__uint_as_float(((x ^ c.sub) & 1u) * c.h0), __uint_as_float(((x ^ c.sub) & 2u) * c.h1),
                           __uint_as_float(((x ^ c.sub) & 4u) * c.h2), __uint_as_float(((x ^ c.sub) & 8u) * c.h3));
    };
#pragma unroll
    for (int c = 0; c < nfull; ++c) {
        const int q = c * 32 + lane;
        const QuadConst k = consts(q);
        const uint32_t *wq = s_w + (q >> 3) * EPW;
        float *dp = dst + q * 4;
#pragma unroll
        for (int e = 0; e < EPW; e += 4) {
            const uint4 w = *reinterpret_cast<const uint4 *>(wq + e);
            store_quad(reinterpret_cast<float4 *>(dp + (e + 0) * slot), expand(k, w.x));
            store_quad(reinterpret_cast<float4 *>(dp + (e + 1) * slot), expand(k, w.y));
            store_quad(reinterpret_cast<float4 *>(dp + (e + 2) * slot), expand(k, w.z));
            store_quad(reinterpret_cast<float4 *>(dp + (e + 3) * slot), expand(k, w.w));
        }
    }
    if (rem) {  // trailing quads: lanes split as (env block, quad) with a power-of-two quad pitch
        int pitch = 1;
        while (pitch < rem) pitch <<= 1;
        const int r = lane & (pitch - 1), blocks = 32 / pitch, per = EPW / blocks;  // per: envs a lane visits (contiguous)
        if (per >= 4) {
            const int e0 = (lane / pitch) * per;
            if (r < rem) {
                const int q = nfull * 32 + r;
                const QuadConst k = consts(q);
                const uint32_t *wq = s_w + (q >> 3) * EPW + e0;
                float *dp = dst + (size_t)e0 * slot + q * 4;
#pragma unroll
                for (int e = 0; e < per; e += 4) {
                    const uint4 w = *reinterpret_cast<const uint4 *>(wq + e);
                    store_quad(reinterpret_cast<float4 *>(dp + (e + 0) * slot), expand(k, w.x));
                    store_quad(reinterpret_cast<float4 *>(dp + (e + 1) * slot), expand(k, w.y));
                    store_quad(reinterpret_cast<float4 *>(dp + (e + 2) * slot), expand(k, w.z));
                    store_quad(reinterpret_cast<float4 *>(dp + (e + 3) * slot), expand(k, w.w));
                }
            }
        } else {  // more env blocks than a lane can fill with a vector load: one env at a time
            const int esub = lane / pitch;
            if (r < rem) {
                const int q = nfull * 32 + r;
                const QuadConst k = consts(q);
                const uint32_t *wq = s_w + (q >> 3) * EPW;
                for (int e = esub; e < EPW; e += blocks)
                    store_quad(reinterpret_cast<float4 *>(dst + (size_t)e * slot + q * 4), expand(k, wq[e]));
            }
        }
    }
}

#ifndef MAEXP_NOISE_UNROLL
#define MAEXP_NOISE_UNROLL 2
#endif
template <int V, int EPW, int NOISE, int NSPEC>
#ifndef MAEXP_MINB
#define MAEXP_MINB 7
#endif
__global__ void __launch_bounds__(kWarpsPerCta * 32, (V == 9 || NOISE != 0) ? 4 : MAEXP_MINB) env_step_kernel(const EnvParams p) {
    constexpr int R = V / 2, VV = V * V;
    constexpr int LW = (V == 9) ? 2 : 1;
    using WinT = typename Win<V>::T;
    extern __shared__ __align__(1024) uint32_t smem[];

    // NSPEC = 8 / 4 / 2: agents and object types are compile-time constants (N = NSPEC, T = 1, no transition frames), so
    // the per-agent loops unroll; NSPEC = 0: everything from the parameter block
    const int N = NSPEC ? NSPEC : p.N, T = NSPEC ? 1 : p.T, X = p.X, Y = p.Y, P = NSPEC ? 0 : p.P, TX = T * X;
    // Shared-memory bit rows are padded: R zero rows on both sides of x and every row shifted left by
    // R bits, so window crops need no bounds checks and no signed shifts.
    const int XP = X + 2 * R;
    // ---- shared memory carve-up (must match env_smem_bytes) ----
    float4 *s_lut = reinterpret_cast<float4 *>(smem);  // [4 classes][16 nibbles] → 4 fp32 values
    uint32_t *s_warp0 = smem + 256;
    const uint32_t *const s_wall = p.wall, *const s_noise = p.noise_area, *const s_area = p.area, *const s_dest = p.dest;
    const uint64_t *const s_los = p.los;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int per_warp_words = (EPW * p.gs + EPW * p.as + EPW * p.ws + (N * EPW + P * N * EPW + 1) / 2 + (EPW + 3) / 4 + 3) & ~3;
    uint32_t *s_w = s_warp0 + (size_t)warp * per_warp_words;  // observation bit words, word k of env e at s_w[k * EPW + e]
    uint32_t *s_grid = s_w + EPW * p.ws;
    uint32_t *s_agrid = s_grid + EPW * p.gs;  // agents occupancy, one bit row per map x
    uint16_t *s_pos = reinterpret_cast<uint16_t *>(s_agrid + EPW * p.as);
    uint16_t *s_past = s_pos + N * EPW;
    uint8_t *s_active = reinterpret_cast<uint8_t *>(s_agrid + EPW * p.as + (N * EPW + P * N * EPW + 1) / 2);

    if (threadIdx.x < 64) {  // nibble → float4 expansion table, classes: 0 plain, 1 straddle, 2 area, 3 tail
        const int cls = threadIdx.x >> 4, n = threadIdx.x & 15;
        const int sub = cls == 0 ? 0 : (cls == 1 ? p.sub_str : (cls == 2 ? 15 : p.sub_tail));
        float v[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) v[j] = (float)((n >> j) & 1) - (float)((sub >> j) & 1);
        s_lut[threadIdx.x] = make_float4(v[0], v[1], v[2], v[3]);
    }
    {  // zero the pad rows (x < 0, x >= X) of this warp's grids once; nothing writes them afterwards
        const int planes = T + 1, per_env = planes * 2 * R;
        for (int i = lane; i < EPW * per_env; i += 32) {
            const int e = i / per_env, j = i - e * per_env, pl = j / (2 * R), r = j - pl * 2 * R;
            const int x = r < R ? r : X + r;  // padded row index in [0, R) or [X+R, X+2R)
            if (pl < T) s_grid[e * p.gs + pl * XP + x] = 0;
            else s_agrid[e * p.as + x] = 0;
        }
    }
    __syncthreads();

    // One resident wave of CTAs; groups of EPW environments are dealt round-robin (all groups cost about the
    // same, and at the benchmark size every warp owns exactly one group).
    const int num_groups = (p.E + EPW - 1) / EPW;
    if (p.stagger_ns && ((blockIdx.x / p.stagger_div) & 1)) {
        // the sequential phase issues no stores and the expansion phase little else: two populations of warps half an
        // agent-iteration apart keep the memory pipe and the issue slots busy at the same time
        unsigned long long t0, t1;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
        do {
            __nanosleep(200);
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
        } while (t1 - t0 < (unsigned long long)p.stagger_ns);
    }
    for (int group = blockIdx.x * kWarpsPerCta + warp; group < num_groups; group += gridDim.x * kWarpsPerCta) {
    const int env0 = group * EPW;
    const int nenv = min(EPW, p.E - env0);
    {  // coalesced stage-in of this warp's grids (flat index, loads batched by the unroll)
        const uint32_t *src = p.obj + (size_t)env0 * TX;
#pragma unroll 4
        for (int i = lane; i < nenv * TX; i += 32) {
            const int e = (int)__umulhi((uint32_t)i, p.tx_magic), r = i - e * TX;
            int t = NSPEC ? 0 : (r >= X) + (r >= 2 * X) + (r >= 3 * X);
            if (!NSPEC && T > 4) t += (r >= 4 * X) + (r >= 5 * X) + (r >= 6 * X) + (r >= 7 * X);
            s_grid[e * p.gs + R + r + t * 2 * R] = src[i] << R;
        }
    }
    __syncwarp();

    // ---------------------------------------------------------------- phase A state (lane = env)
    const bool mine = lane < nenv;
    const int env = env0 + lane;
    const uint32_t gid = p.gid0 + (uint32_t)env;
    uint32_t *grid = s_grid + lane * p.gs + R;    // row (t, x) at grid[t * XP + x], x in [-R, X + R)
    uint32_t *agrid = s_agrid + lane * p.as + R;  // row x at agrid[x]
    uint32_t *wv = s_w + lane;  // word k at wv[k * EPW]
    uint16_t *mypos = s_pos + lane;  // agent k at mypos[k*EPW]
    int stepc = 0, nobj = 0;
    int c_gen = 0, c_done = 0, c_ca = 0, c_cw = 0;
    uint32_t ca = 0, cw = 0, rew_pos = 0, rew_neg = 0, done_bits = 0;
    uint64_t dsel = 0, acts_taken = 0;
    bool is_reset = false, active = false;
    uint32_t attempt = 0;
    uint4 rw = make_uint4(0, 0, 0, 0);

    // Moving obstacles of dynamic_simple (dynamic_simple_map.py:29-63), packed x | y << 8; -1 on every other map.
    // Obstacle 0 shuttles along y in [1, Y-2] on column X/2 starting at y = 1, obstacle 1 along x in [1, X-2] on
    // row Y/2 starting at x = X/2, one cell per world step, turning at the ends: a triangle wave of the number of
    // world steps k since creation (map.reset() does not touch them).
    int ob0 = -1, ob1 = -1;
    auto place_obstacles = [&](uint32_t k) {
        const uint32_t L0 = (uint32_t)(Y - 3), L1 = (uint32_t)(X - 3);
        const uint32_t m0 = k % (2 * L0), m1 = (k + (uint32_t)(X / 2 - 1)) % (2 * L1);
        ob0 = (X / 2) | ((1 + (int)(m0 <= L0 ? m0 : 2 * L0 - m0)) << 8);
        ob1 = (1 + (int)(m1 <= L1 ? m1 : 2 * L1 - m1)) | ((Y / 2) << 8);
    };
    auto wall_at = [&](int x, int y) {
        const int cell = x | (y << 8);
        return (__ldg(&s_wall[x]) & (1u << (y + R))) != 0 || cell == ob0 || cell == ob1;
    };
    auto agent_at = [&](int x, int y) { return (agrid[x] >> (y + R)) & 1u; };
    // default_environment.py:134-148, one accepted object; two uniforms per attempt
    int32_t *const hm = p.out.heatmaps;  // optional visit statistics, see maexp_obs_out
    auto spawn = [&](int t, uint32_t stream) {
        for (;;) {
            if (attempt >= (uint32_t)kRespawnCap) {
                atomicOr(p.error_flag, 1);
                return;
            }
            if ((attempt & 1u) == 0) rw = philox4x32_10(p.seed, gid, p.tick, stream, attempt >> 1, 0);
            const uint32_t u1 = (attempt & 1u) ? rw.z : rw.x, u2 = (attempt & 1u) ? rw.w : rw.y;
            ++attempt;
            const int x = 1 + (int)scale(u1, (uint32_t)(X - 1)), y = 1 + (int)scale(u2, (uint32_t)(Y - 1));
            const uint32_t bit = 1u << (y + R);
            bool ok = !wall_at(x, y) && (__ldg(&s_area[t * X + x]) & bit);
            for (int t2 = 0; t2 < T; ++t2) ok = ok && !(grid[t2 * XP + x] & bit);
            ok = ok && !agent_at(x, y);
            if (ok) {
                grid[t * XP + x] |= bit;
                ++c_gen;
                ++nobj;
                if (hm) atomicAdd(hm + ((2 * N + t) * X + x) * Y + y, 1);  // heatmap_objects (default_environment.py:145)
                return;
            }
        }
    };

    if (mine) {
        const bool pend = p.pending[env] != 0;
        const bool masked = p.env_mask ? (p.env_mask[env] != 0) : true;
        is_reset = p.force_reset ? masked : (p.auto_reset && pend);
        active = p.force_reset ? masked : true;
    }
    if (mine) s_active[lane] = active ? 1 : 0;

    if (mine && active) {
        stepc = p.step[env];
        const int4 cn = *reinterpret_cast<const int4 *>(p.counters + (size_t)env * 4);
        c_gen = cn.x; c_done = cn.y; c_ca = cn.z; c_cw = cn.w;
        for (int k = 0; k < N; ++k) mypos[k * EPW] = p.pos[(size_t)env * N + k];
        if (p.random_dest)
            for (int k = 0; k < N; ++k) dsel |= (uint64_t)p.dest_sel[(size_t)env * N + k] << (4 * k);
        for (int q = 0; q < P * N; ++q) s_past[q * EPW + lane] = p.past[(size_t)env * P * N + q];
        if (p.wsteps) {  // abstract_world.py:40: the map steps first; a reset leaves the obstacles where they are
            uint32_t k = p.wsteps[env];
            if (!is_reset) p.wsteps[env] = ++k;
            place_obstacles(k);
        }

        if (is_reset) {
            // ------------------------------------------------ default_environment.py:64-116
            c_gen = c_done = c_ca = c_cw = 0;
            stepc = 0;
            for (int t = 0; t < T; ++t)
                for (int x = 0; x < X; ++x) grid[t * XP + x] = 0;
            if (p.random_dest) {  // three_rooms_map.py:42-55
                uint64_t perm = 0;
                for (int k = 0; k < p.M; ++k) perm |= (uint64_t)k << (4 * k);
                uint32_t d = 0;
                perm = fisher_yates(perm, p.M, p.seed, gid, p.tick, kStreamDest, d);
                dsel = 0;
                for (int k = 0; k < N; ++k) {
                    const uint4 w = philox4x32_10(p.seed, gid, p.tick, kStreamDest, d >> 2, 0);
                    const uint32_t c = scale(pick(w, d & 3), (uint32_t)p.M);
                    ++d;
                    dsel |= (uint64_t)nib(perm, (int)c) << (4 * k);
                }
                for (int k = 0; k < N; ++k) p.dest_sel[(size_t)env * N + k] = (uint8_t)nib(dsel, k);
            }
            uint64_t perm = 0;  // default_environment.py:33,97-98 (persists across episodes)
            const int NI = p.n_init;  // all rows are shuffled, the first N of the order are used
            for (int k = 0; k < NI; ++k) perm |= (uint64_t)p.init_perm[(size_t)env * NI + k] << (4 * k);
            if (p.shuffle_init) {
                uint32_t d = 0;
                perm = fisher_yates(perm, NI, p.seed, gid, p.tick, kStreamInitPerm, d);
                for (int k = 0; k < NI; ++k) p.init_perm[(size_t)env * NI + k] = (uint8_t)nib(perm, k);
            }
            for (int k = 0; k < N; ++k) mypos[k * EPW] = p.init_xy[nib(perm, k)];
            for (int x = 0; x < X; ++x) agrid[x] = 0;
            for (int k = 0; k < N; ++k) agrid[mypos[k * EPW] & 0xFF] |= 1u << ((mypos[k * EPW] >> 8) + R);
            attempt = 0;
            nobj = 0;
            // sequential_environment.py:21-23 / junctional_environment.py:21-23 only generate type 0
            const int t_gen = (NSPEC || p.env_kind == MAEXP_ENV_DEFAULT) ? T : 1;
            for (int t = 0; t < t_gen; ++t)
                for (int n = 0; n < p.num_objects; ++n) spawn(t, kStreamObjects);
            for (int q = 0; q < P; ++q)  // transition_agent_observation.py:76-79
                for (int k = 0; k < N; ++k) s_past[(q * N + k) * EPW + lane] = mypos[k * EPW];
            attempt = 0;
        } else {
            // ------------------------------------------------ abstract_world.py:39-78
            for (int t = 0; t < T; ++t)
                for (int x = 0; x < X; ++x) nobj += __popc(grid[t * XP + x]);
            for (int x = 0; x < X; ++x) agrid[x] = 0;
            for (int k = 0; k < N; ++k) agrid[mypos[k * EPW] & 0xFF] |= 1u << ((mypos[k * EPW] >> 8) + R);
            uint64_t acts = 0;  // 3 bits per agent
            if (p.actions) {
                for (int k = 0; k < N; ++k) acts |= (uint64_t)(p.actions[(size_t)env * N + k] & 7) << (3 * k);
            } else {
                for (int c = 0; c < (N + 3) / 4; ++c) {
                    const uint4 w = philox4x32_10(p.seed, gid, p.tick, kStreamAction, (uint32_t)c, 0);
                    for (int j = 0; j < 4; ++j)
                        if (c * 4 + j < N) acts |= (uint64_t)(pick(w, j) >> 30) << (3 * (c * 4 + j));
                }
            }
            acts_taken = acts;
            uint64_t ord = 0;
            if (p.order) {
                for (int k = 0; k < N; ++k) ord |= (uint64_t)(p.order[(size_t)env * N + k] & 15) << (4 * k);
            } else {
                for (int k = 0; k < N; ++k) ord |= (uint64_t)k << (4 * k);
                uint32_t d = 0;
                ord = fisher_yates(ord, N, p.seed, gid, p.tick, kStreamOrder, d);
            }
            for (int j = 0; j < N; ++j) {
                const int id = (int)nib(ord, j);
                const int a = (int)((acts >> (3 * id)) & 7);
                // action.py:4-9 in index space (abstract_map.py:49-50 flips y)
                const int dx = (a == 0) - (a == 2), dy = (a == 3) - (a == 1);
                const uint16_t cur = mypos[id * EPW];
                const int nx = (cur & 0xFF) + dx, ny = (cur >> 8) + dy;
                bool move = true;
                // another agent on the target cell (abstract_world.py:56-64; own cell excluded: only STAY hits it)
                if ((dx | dy) != 0 && ((agrid[nx] >> (ny + R)) & 1u)) {
                    ca |= 1u << id;
                    move = false;
                }
                if (wall_at(nx, ny)) {
                    cw |= 1u << id;
                    move = false;
                }
                if (move) {
                    agrid[cur & 0xFF] &= ~(1u << ((cur >> 8) + R));
                    agrid[nx] |= 1u << (ny + R);
                    mypos[id * EPW] = (uint16_t)(nx | (ny << 8));
                }
            }
        }
    }
    __syncwarp();

    const bool all_on = __all_sync(0xFFFFFFFFu, !mine || active);
    // ------------------------------------------------------------ per-agent interleave
    for (int i = 0; i < N; ++i) {
        bool gate = false;
        if (mine && active) {
            const uint16_t me = mypos[i * EPW];
            const int px = me & 0xFF, py = me >> 8;
            if (!is_reset) {
                // ---------------------------------------- default_environment.py:196-231
                if (hm) atomicAdd(hm + (i * X + px) * Y + py, 1);  // heatmap_agents
                if (!p.no_task[i]) {
                    const int sel = p.random_dest ? (int)nib(dsel, i) : (i < p.M ? i : p.M - 1);
                    const uint32_t bit = 1u << (py + R);
                    for (int q = 0; q < MAEXP_MAX_TYPES && p.task_order[i][q] >= 0; ++q) {
                        const int t = p.task_order[i][q];
                        if ((grid[t * XP + px] & bit) && (__ldg(&s_dest[sel * X + px]) & bit)) {
                            rew_pos |= 1u << i;
                            grid[t * XP + px] &= ~bit;
                            --nobj;
                            ++c_done;
                            if (hm) atomicAdd(hm + ((N + i) * X + px) * Y + py, 1);  // heatmap_complete
                            if (NSPEC || p.env_kind == MAEXP_ENV_DEFAULT) {
                                if (p.keep) spawn(t, kStreamRespawn);
                            } else {
                                // sequential_environment.py:44-48, junctional_environment.py:44-56: the next object of
                                // the chain appears on this cell (generate_object_at), the last type respawns a type 0
                                int next = -1;
                                bool respawn = false;
                                if (p.env_kind == MAEXP_ENV_JUNCTIONAL && t == 1) {
                                    if (i == 2 || i == 3) next = 3;
                                    if (i == 4 || i == 5) next = 2;
                                } else if (p.env_kind == MAEXP_ENV_JUNCTIONAL && t == 2) next = 4;
                                else if (t < T - 1) next = t + 1;
                                else respawn = p.keep != 0;
                                if (next >= 0) {
                                    if (!(grid[next * XP + px] & bit)) ++nobj;
                                    grid[next * XP + px] |= bit;
                                    ++c_gen;
                                    if (hm) atomicAdd(hm + ((2 * N + next) * X + px) * Y + py, 1);
                                }
                                if (respawn) spawn(0, kStreamRespawn);
                            }
                        }
                    }
                    if (ca & (1u << i)) {
                        rew_neg |= 1u << i; ++c_ca;
                        if (hm) atomicAdd(hm + ((2 * N + 2 * T + 1) * X + px) * Y + py, 1);  // heatmap_agents_collision
                    }
                    if (cw & (1u << i)) {
                        rew_neg |= 1u << i; ++c_cw;
                        if (hm) atomicAdd(hm + ((2 * N + 2 * T) * X + px) * Y + py, 1);  // heatmap_wall_collision
                    }
                }
                // ---------------------------------------- default_environment.py:233-237
                if (nobj == 0 || stepc == p.L - 1) done_bits |= 1u << i;
            }
            // -------------------------------------------- observation_handler.py:67-105, bits in output order
            WinT los = __ldg(&s_los[(px * Y + py) * LW]);
            if constexpr (V == 9) los |= (WinT)__ldg(&s_los[(px * Y + py) * LW + 1]) << 64;
            auto win_bit = [&](uint16_t cell) -> WinT {  // cell → window bit or 0
                const int a = (int)(cell & 0xFF) - px + R, b = (int)(cell >> 8) - py + R;
                return (a >= 0 && a < V && b >= 0 && b < V) ? ((WinT)1 << (a * V + b)) : (WinT)0;
            };
            auto row_window = [&](const uint32_t *rows) -> WinT {  // V x V crop of a padded bit-row grid around (px, py)
                WinT m = 0;
                const uint32_t *r0 = rows + (px - R);
#pragma unroll
                for (int a = 0; a < V; ++a) m |= (WinT)((r0[a] >> py) & ((1u << V) - 1u)) << (a * V);
                return m;
            };
            bool done_fast = false;
            if constexpr (V <= 7) {
                if (p.fast3) {  // channels: total agents, all objects, area — assembled with compile-time shifts
                    uint64_t A = row_window(agrid), O = 0;
                    for (int t = 0; t < T; ++t) O |= row_window(grid + t * XP);
                    if (p.agent_view == MAEXP_AV_DEFAULT) A &= los;
                    if (p.object_view == MAEXP_OV_DEFAULT) O &= los;
                    const unsigned __int128 lo = (unsigned __int128)A | ((unsigned __int128)O << VV) |
                                                 ((unsigned __int128)los << (2 * VV));
                    wv[0] = (uint32_t)lo;
                    wv[EPW] = (uint32_t)(lo >> 32);
                    wv[2 * EPW] = (uint32_t)(lo >> 64);
                    wv[3 * EPW] = (uint32_t)(lo >> 96);
                    constexpr int kHi = 128 - 2 * VV;
                    wv[4 * EPW] = kHi < 64 ? (uint32_t)(los >> (kHi < 64 ? kHi : 0)) : 0u;
                    wv[5 * EPW] = 0u;
                    done_fast = true;
                }
            }
            if (!done_fast) {
            BitStream<V, EPW> bs(wv);
            const int av = p.agent_view;
            if (av == MAEXP_AV_INDIVIDUAL || av == MAEXP_AV_INDIVIDUAL_TRANSITION)
                for (int k = 0; k < N; ++k) bs.put(win_bit(mypos[k * EPW]) & los);
            if (av != MAEXP_AV_INDIVIDUAL) {
                WinT m = row_window(agrid);
                if (av != MAEXP_AV_SIMPLE) m &= los;
                bs.put(m);
            }
            if (av == MAEXP_AV_TRANSITION || av == MAEXP_AV_INDIVIDUAL_TRANSITION)
                for (int q = 0; q < P; ++q) {  // transition_agent_observation.py:47-57,81-94
                    const uint16_t mine_then = s_past[(q * N + i) * EPW + lane];
                    const int qx = mine_then & 0xFF, qy = mine_then >> 8;
                    WinT m = 0;
                    for (int k = 0; k < N; ++k) {
                        const uint16_t c = s_past[(q * N + k) * EPW + lane];
                        const int ux = c & 0xFF, uy = c >> 8;
                        if (abs(ux - qx) <= R && abs(uy - qy) <= R) m |= win_bit(c);
                    }
                    bs.put(m);
                }
            auto obj_window = [&](int t) -> WinT { return row_window(grid + t * XP); };
            if (p.object_view == MAEXP_OV_TYPES) {
                for (int t = 0; t < T; ++t) bs.put(obj_window(t) & los);
            } else {
                WinT m = 0;
                for (int t = 0; t < T; ++t) m |= obj_window(t);
                if (p.object_view == MAEXP_OV_DEFAULT) m &= los;
                bs.put(m);
            }
            bs.put(los);  // observation_handler.py:107-108 (value = bit - 1)
            bs.finish();
            }
            gate = (__ldg(&s_noise[px]) >> (py + R)) & 1u;

            // -------------------------------------------- optional reference-format extras
            if (p.out.coordinates) {  // local_area_mask.py:18-47
                int32_t *c = p.out.coordinates + ((size_t)env * N + i) * 12;
                c[0] = c[4] = abs(min(0, px - R));
                c[1] = c[5] = V - max(0, px + R - (X - 1));
                c[2] = c[6] = abs(min(0, py - R));
                c[3] = c[7] = V - max(0, py + R - (Y - 1));
                c[8] = max(0, px - R);
                c[9] = min(X - 1, px + R) + 1;
                c[10] = max(0, py - R);
                c[11] = min(Y - 1, py + R) + 1;
            }
            if (p.out.objects_bits) {  // the same snapshot as bit rows, env-minor: a warp's lanes store contiguous words
                uint32_t *o = p.out.objects_bits + (size_t)i * TX * p.E + env;
                for (int r = 0; r < TX; ++r) o[(size_t)r * p.E] = grid[(r / X) * XP + (r % X)] >> R;
            }
            if (p.out.objects) {
                int8_t *o = p.out.objects + ((size_t)env * N + i) * TX * Y;
                for (int t = 0; t < T; ++t)
                    for (int x = 0; x < X; ++x)
                        for (int y = 0; y < Y; ++y)
                            o[(t * X + x) * Y + y] = (int8_t)((grid[t * XP + x] >> (y + R)) & 1u);
            }
        }
        if constexpr (NOISE != MAEXP_NOISE_NONE) {
            if (mine) s_active[lane] = (uint8_t)((active ? 1 : 0) | (gate ? 2 : 0));
        }
        __syncwarp();

        // ---------------------------------------------------------- phase B: expand + stream out
        if (p.out.obs) {
            const int qps = p.qps;
            // ring slot layout is agent-major [N][E][slot_floats]: agent i's slots of this group are one contiguous run
            const int env_stride = p.slot_floats;
            const size_t base = ((size_t)i * p.E + env0) * p.slot_floats;
            float *const dst = p.out.obs + base;
            float *const dst_rel = (p.merged && p.out.obs_rel) ? p.out.obs_rel + base : nullptr;
            const bool check = !all_on;  // some env of the group is masked out of this call
            // envs interleaved per lane in the expansion loops: independent Philox chains of the noise variants overlap
            constexpr int kUnrollEnvs = NOISE != MAEXP_NOISE_NONE ? MAEXP_NOISE_UNROLL : 4;

            // per-element noise constants of quad q (the same for every env a lane visits): scale of the draw, threshold,
            // and the +1 / -1 the flip noise wraps around channel 2 (flip_observation_noise.py:31,44)
            struct QuadNoise { float scale[4], thr[4], pre[4]; };
            auto quad_noise = [&](int q) {
                QuadNoise n;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int k = q * 4 + j;
                    const int c = k / VV, rem = k - c * VV, a = rem / V, b = rem - a * V;
                    const int ring = min(min(min(a, b), min(V - 1 - a, V - 1 - b)), V / 2 - 1);
                    const bool in = k < p.K;
                    n.pre[j] = 0.0f;
                    n.thr[j] = 0.0f;
                    if constexpr (NOISE == MAEXP_NOISE_FLIP) {
                        n.scale[j] = in ? p.flip_range[ring] : 0.0f;
                        n.pre[j] = (in && c == 2) ? 1.0f : 0.0f;
                    } else if constexpr (NOISE == MAEXP_NOISE_FLAT) {
                        n.scale[j] = (in && c < 3) ? 0.03f : 0.0f;  // flat_observation_noise.py:28-34 (3 hard-coded channels)
                    } else {
                        n.scale[j] = in ? p.std_dist[ring] : 0.0f;
                        n.thr[j] = p.noise_thr[ring];
                    }
                }
                return n;
            };
            // one quad (4 consecutive fp32 of the slot) of env e: table lookup, optional noise, 128-bit stores
            auto emit = [&](int e, int q, uint32_t word, int sh, int lut_off, const QuadNoise &qn) {
                const float4 o = s_lut[lut_off + ((word >> sh) & 0xFu)];
                const int off = e * env_stride + q * 4;
                if constexpr (NOISE == MAEXP_NOISE_NONE) {
                    store_quad(reinterpret_cast<float4 *>(dst + off), o);
                    if (dst_rel) store_quad(reinterpret_cast<float4 *>(dst_rel + off), o);
                } else {
                    if (dst_rel) store_quad(reinterpret_cast<float4 *>(dst_rel + off), o);  // merged: relative stays clean
                    float v[4] = {o.x, o.y, o.z, o.w};
                    const int k0 = q * 4;
                    if (s_active[e] & 2) {
                        float dr[4];
                        if (p.noise_tape) {
                            const float *t = p.noise_tape + ((size_t)(env0 + e) * N + i) * p.K + k0;
#pragma unroll
                            for (int j = 0; j < 4; ++j) dr[j] = (k0 + j < p.K) ? t[j] : 0.0f;
                        } else {
                            const uint4 w = philox4x32_10(p.seed, p.gid0 + (uint32_t)(env0 + e), p.tick, kStreamNoise,
                                                          (uint32_t)q, (uint32_t)i);
                            if constexpr (NOISE == MAEXP_NOISE_FLIP) {
                                dr[0] = unit_float(w.x); dr[1] = unit_float(w.y);
                                dr[2] = unit_float(w.z); dr[3] = unit_float(w.w);
                            } else {  // Box-Muller pairs; bit-exact parity is only claimed for taped draws
                                const float r0 = sqrtf(-2.0f * logf(((float)(w.x >> 8) + 0.5f) * (1.0f / 16777216.0f)));
                                const float r1 = sqrtf(-2.0f * logf(((float)(w.z >> 8) + 0.5f) * (1.0f / 16777216.0f)));
                                const float a0 = 6.28318530717958647692f * unit_float(w.y);
                                const float a1 = 6.28318530717958647692f * unit_float(w.w);
                                dr[0] = r0 * cosf(a0); dr[1] = r0 * sinf(a0);
                                dr[2] = r1 * cosf(a1); dr[3] = r1 * sinf(a1);
                            }
                        }
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            float x = v[j];
                            if constexpr (NOISE == MAEXP_NOISE_FLIP) {  // flip_observation_noise.py:28-46
                                x = __fadd_rn(x, qn.pre[j]);
                                x = __fadd_rn(x, __fmul_rn(dr[j], qn.scale[j]));
                                x = fabsf(rintf(x));
                                x = __fadd_rn(x, -qn.pre[j]);
                            } else if constexpr (NOISE == MAEXP_NOISE_THRESHOLD_DIST) {  // dist_observation_noise.py:53-75
                                const float g = __fmul_rn(dr[j], qn.scale[j]), th = qn.thr[j];
                                float n = 0.0f;
                                if (g < -th) n = g;
                                if (th < g) n = g;
                                x = __fadd_rn(x, n);
                            } else {  // dist_observation_noise.py:29-41, flat_observation_noise.py:28-34
                                x = __fadd_rn(x, __fmul_rn(dr[j], qn.scale[j]));
                            }
                            v[j] = x;
                        }
                    }
                    store_quad(reinterpret_cast<float4 *>(dst + off), make_float4(v[0], v[1], v[2], v[3]));
                }
            };
            auto quad_class = [&](int q) { return q > p.q_tail ? 0 : ((q >= p.q_str) + (q > p.q_str) + (q == p.q_tail)) * 16; };

            if (NOISE == MAEXP_NOISE_NONE && !check && !dst_rel && nenv == EPW) {
                // hot path: the whole group, compile-time slot geometry for the common 3-channel shapes
                switch (p.slot_floats) {
                case 152: expand_group<152, EPW>(s_w, dst, lane, 152, p.q_str, p.q_tail, p.sub_str, p.sub_tail); break;  // V = 7, C = 3
                case 80: expand_group<80, EPW>(s_w, dst, lane, 80, p.q_str, p.q_tail, p.sub_str, p.sub_tail); break;     // V = 5, C = 3
                default: expand_group<0, EPW>(s_w, dst, lane, p.slot_floats, p.q_str, p.q_tail, p.sub_str, p.sub_tail);
                }
            } else {
            bool pairs_done = false;
            if constexpr (NOISE == MAEXP_NOISE_FLIP) {
                if (!p.noise_tape) {
                    // Flip noise with on-chip draws: 16-bit uniforms, EIGHT per Philox4x32-10 call (element k: call k >> 3, word
                    // (k & 7) >> 1, low half for even k; DESIGN.md RNG specification, oracle noise_draw).  A lane owns a PAIR of
                    // quads (8 consecutive elements = one call = 32 contiguous output bytes), so the draws cost half the calls of
                    // the one-word-per-element form (86 -> 43 per agent-step at C = 7, V = 7).
                    auto u16f = [](uint32_t h) { return __fadd_rn(__uint_as_float(0x3f800000u | (h << 7)), -1.0f); };  // h / 65536, exact
                    auto flip4 = [&](float (&v)[4], uint32_t w0, uint32_t w1, const QuadNoise &qn) {
                        const float dr[4] = {u16f(w0 & 0xFFFFu), u16f(w0 >> 16), u16f(w1 & 0xFFFFu), u16f(w1 >> 16)};
#pragma unroll
                        for (int j = 0; j < 4; ++j) {  // flip_observation_noise.py:28-46
                            float x = __fadd_rn(v[j], qn.pre[j]);
                            x = __fadd_rn(x, __fmul_rn(dr[j], qn.scale[j]));
                            x = fabsf(rintf(x));
                            v[j] = __fadd_rn(x, -qn.pre[j]);
                        }
                    };
                    auto emit_pair = [&](int e, int qp, uint32_t word, int lut_a, int lut_b, const QuadNoise &qa, const QuadNoise &qb,
                                         bool has_b) {
                        const int q0 = 2 * qp, sh = (q0 & 7) * 4, off = e * env_stride + q0 * 4;
                        const float4 oa = s_lut[lut_a + ((word >> sh) & 0xFu)], ob = s_lut[lut_b + ((word >> (sh + 4)) & 0xFu)];
                        if (dst_rel) {  // merged: the relative key stays clean
                            store_quad(reinterpret_cast<float4 *>(dst_rel + off), oa);
                            if (has_b) store_quad(reinterpret_cast<float4 *>(dst_rel + off + 4), ob);
                        }
                        float va[4] = {oa.x, oa.y, oa.z, oa.w}, vb[4] = {ob.x, ob.y, ob.z, ob.w};
                        if (s_active[e] & 2) {
                            const uint4 w = philox4x32_10(p.seed, p.gid0 + (uint32_t)(env0 + e), p.tick, kStreamNoise, (uint32_t)qp, (uint32_t)i);
                            flip4(va, w.x, w.y, qa);
                            flip4(vb, w.z, w.w, qb);
                        }
                        store_quad(reinterpret_cast<float4 *>(dst + off), make_float4(va[0], va[1], va[2], va[3]));
                        if (has_b) store_quad(reinterpret_cast<float4 *>(dst + off + 4), make_float4(vb[0], vb[1], vb[2], vb[3]));
                    };
                    const int npairs = (qps + 1) >> 1, nfullp = npairs >> 5, remp = npairs & 31;
                    for (int c = 0; c < nfullp; ++c) {
                        const int qp = c * 32 + lane, q0 = 2 * qp;
                        const bool has_b = q0 + 1 < qps;
                        const int lut_a = quad_class(q0), lut_b = quad_class(q0 + 1);
                        const uint32_t *wq = s_w + (q0 >> 3) * EPW;
                        const QuadNoise qa = quad_noise(q0), qb = quad_noise(q0 + 1);
                        if (!check) {
#pragma unroll kUnrollEnvs
                            for (int e = 0; e < nenv; ++e) emit_pair(e, qp, wq[e], lut_a, lut_b, qa, qb, has_b);
                        } else {
                            for (int e = 0; e < nenv; ++e)
                                if (s_active[e] & 1) emit_pair(e, qp, wq[e], lut_a, lut_b, qa, qb, has_b);
                        }
                    }
                    if (remp) {
                        int pitch = 1;
                        while (pitch < remp) pitch <<= 1;
                        const int r = lane & (pitch - 1), esub = lane / pitch, estep = 32 / pitch;
                        if (r < remp) {
                            const int qp = nfullp * 32 + r, q0 = 2 * qp;
                            const bool has_b = q0 + 1 < qps;
                            const int lut_a = quad_class(q0), lut_b = quad_class(q0 + 1);
                            const uint32_t *wq = s_w + (q0 >> 3) * EPW;
                            const QuadNoise qa = quad_noise(q0), qb = quad_noise(q0 + 1);
                            if (!check) {
#pragma unroll kUnrollEnvs
                                for (int e = esub; e < nenv; e += estep) emit_pair(e, qp, wq[e], lut_a, lut_b, qa, qb, has_b);
                            } else {
                                for (int e = esub; e < nenv; e += estep)
                                    if (s_active[e] & 1) emit_pair(e, qp, wq[e], lut_a, lut_b, qa, qb, has_b);
                            }
                        }
                    }
                    pairs_done = true;
                }
            }
            if (!pairs_done) {
            // (1) full chunks: lane owns quad q = c*32 + lane of every env → per-lane constants hoisted,
            //     each store instruction covers 512 contiguous bytes of one agent slot
            const int nfull = qps >> 5;
            for (int c = 0; c < nfull; ++c) {
                const int q = c * 32 + lane;
                const int sh = (q & 7) * 4, lut_off = quad_class(q);
                const uint32_t *wq = s_w + (q >> 3) * EPW;
                const QuadNoise qn = quad_noise(q);
                if (!check) {
#pragma unroll kUnrollEnvs
                    for (int e = 0; e < nenv; ++e) emit(e, q, wq[e], sh, lut_off, qn);
                } else {
                    for (int e = 0; e < nenv; ++e)
                        if (s_active[e] & 1) emit(e, q, wq[e], sh, lut_off, qn);
                }
            }
            // (2) the qps % 32 trailing quads: lanes split as (env sub-index, quad) with a power-of-two quad
            //     pitch, so a lane again owns a fixed quad and only the env advances
            const int rem = qps & 31;
            if (rem) {
                int pitch = 1;
                while (pitch < rem) pitch <<= 1;
                const int r = lane & (pitch - 1), esub = lane / pitch, estep = 32 / pitch;
                if (r < rem) {
                    const int q = nfull * 32 + r;
                    const int sh = (q & 7) * 4, lut_off = quad_class(q);
                    const uint32_t *wq = s_w + (q >> 3) * EPW;
                    const QuadNoise qn = quad_noise(q);
                    if (!check) {
#pragma unroll kUnrollEnvs
                        for (int e = esub; e < nenv; e += estep) emit(e, q, wq[e], sh, lut_off, qn);
                    } else {
                        for (int e = esub; e < nenv; e += estep)
                            if (s_active[e] & 1) emit(e, q, wq[e], sh, lut_off, qn);
                    }
                }
            }
            }
            }
        }
        __syncwarp();
    }

    // -------------------------------------------------------------------- epilogue (lane = env)
    if (mine && active) {
        if (!is_reset) {
            // observation_handler.py:146-148 → transition_agent_observation.py:71-74
            for (int q = P - 1; q >= 1; --q)
                for (int k = 0; k < N; ++k) s_past[(q * N + k) * EPW + lane] = s_past[((q - 1) * N + k) * EPW + lane];
            if (P > 0)
                for (int k = 0; k < N; ++k) s_past[k * EPW + lane] = mypos[k * EPW];
            ++stepc;  // default_environment.py:172
            if (hm)  // :173 heatmap_objects_left += objects_matrix
                for (int t = 0; t < T; ++t)
                    for (int x = 0; x < X; ++x)
                        for (uint32_t row = grid[t * XP + x] >> R; row; row &= row - 1)
                            atomicAdd(hm + ((2 * N + T + t) * X + x) * Y + (__ffs(row) - 1), 1);
        }
        float shared = 0.0f;
        if (p.reward_mode != MAEXP_REWARD_DEFAULT) {  // shared_reward_environment.py:15-28
            const int sum = __popc(rew_pos & ~rew_neg) - __popc(rew_neg);
            const bool mean = p.reward_mode == MAEXP_REWARD_SHARED_MEAN || p.shared_max_is_mean;
            if (mean) shared = (float)((double)sum / (double)N);
            else shared = (rew_pos & ~rew_neg) ? 1.0f : ((rew_neg == ((1u << N) - 1u)) ? -1.0f : 0.0f);
        }
        auto reward_of = [&](int k) -> float {
            float r = (rew_neg >> k) & 1u ? -1.0f : ((rew_pos >> k) & 1u ? 1.0f : 0.0f);
            if (p.reward_mode != MAEXP_REWARD_DEFAULT) r = shared;
            return is_reset ? 0.0f : r;
        };
        if ((N & 7) == 0) {
            // eight agents at a time: every per-agent output row of this env is an aligned 8 / 16 / 32-byte run
            for (int k0 = 0; k0 < N; k0 += 8) {
                const size_t o = (size_t)env * N + k0;
                uint32_t pw[4];
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    pw[j] = (uint32_t)mypos[(k0 + 2 * j) * EPW] | ((uint32_t)mypos[(k0 + 2 * j + 1) * EPW] << 16);
                *reinterpret_cast<uint4 *>(p.pos + o) = make_uint4(pw[0], pw[1], pw[2], pw[3]);
                if (p.out.rewards) {
                    float4 *dst = reinterpret_cast<float4 *>(p.out.rewards + o);
                    dst[0] = make_float4(reward_of(k0), reward_of(k0 + 1), reward_of(k0 + 2), reward_of(k0 + 3));
                    dst[1] = make_float4(reward_of(k0 + 4), reward_of(k0 + 5), reward_of(k0 + 6), reward_of(k0 + 7));
                }
                if (p.out.rewards_i8) {  // -1 / 0 / +1 as bytes (default reward mode)
                    uint32_t r8[2] = {0u, 0u};
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const int k = k0 + j;
                        const uint32_t v = is_reset ? 0u : ((rew_neg >> k) & 1u ? 0xFFu : ((rew_pos >> k) & 1u));
                        r8[j >> 2] |= v << (8 * (j & 3));
                    }
                    *reinterpret_cast<uint2 *>(p.out.rewards_i8 + o) = make_uint2(r8[0], r8[1]);
                }
                if (p.out.dones) {
                    const uint32_t d = done_bits >> k0;  // bit k -> byte k
                    const uint32_t lo = (d & 1u) | ((d & 2u) << 7) | ((d & 4u) << 14) | ((d & 8u) << 21);
                    const uint32_t hi = ((d >> 4) & 1u) | (((d >> 4) & 2u) << 7) | (((d >> 4) & 4u) << 14) | (((d >> 4) & 8u) << 21);
                    *reinterpret_cast<uint2 *>(p.out.dones + o) = make_uint2(lo, hi);
                }
                if (p.out.positions) {  // int8 (x, y) pairs: the u16 cells already are x | y << 8
                    *reinterpret_cast<uint4 *>(p.out.positions + o * 2) = make_uint4(pw[0], pw[1], pw[2], pw[3]);
                }
                if (p.out.actions_taken) {
                    uint32_t a[2] = {0u, 0u};
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const uint32_t v = is_reset ? 0xFFu : (uint32_t)((acts_taken >> (3 * (k0 + j))) & 7);
                        a[j >> 2] |= v << (8 * (j & 3));
                    }
                    *reinterpret_cast<uint2 *>(p.out.actions_taken + o) = make_uint2(a[0], a[1]);
                }
                if (p.out.destination)
                    for (int j = 0; j < 8; ++j) {
                        const int k = k0 + j;
                        p.out.destination[o + j] = (uint8_t)(p.random_dest ? nib(dsel, k) : (k < p.M ? k : p.M - 1));
                    }
            }
        } else {
        for (int k = 0; k < N; ++k) {
            const float r = reward_of(k);
            if (p.out.rewards) p.out.rewards[(size_t)env * N + k] = r;
            if (p.out.rewards_i8) p.out.rewards_i8[(size_t)env * N + k] = (int8_t)r;
            if (p.out.dones) p.out.dones[(size_t)env * N + k] = (uint8_t)((done_bits >> k) & 1u);
            if (p.out.positions) {
                p.out.positions[((size_t)env * N + k) * 2 + 0] = (int8_t)(mypos[k * EPW] & 0xFF);
                p.out.positions[((size_t)env * N + k) * 2 + 1] = (int8_t)(mypos[k * EPW] >> 8);
            }
            if (p.out.actions_taken)
                p.out.actions_taken[(size_t)env * N + k] = is_reset ? (int8_t)-1 : (int8_t)((acts_taken >> (3 * k)) & 7);
            if (p.out.destination)
                p.out.destination[(size_t)env * N + k] =
                    (uint8_t)(p.random_dest ? nib(dsel, k) : (k < p.M ? k : p.M - 1));
            p.pos[(size_t)env * N + k] = mypos[k * EPW];
        }
        }
        if (p.out.was_reset) p.out.was_reset[env] = is_reset ? 1 : 0;
        if (p.out.done_env) p.out.done_env[env] = (uint8_t)(done_bits & 1u);
        p.step[env] = stepc;
        p.pending[env] = (p.auto_reset && !is_reset && (done_bits & 1u)) ? 1 : 0;  // default_trainer.py:85
        *reinterpret_cast<int4 *>(p.counters + (size_t)env * 4) = make_int4(c_gen, c_done, c_ca, c_cw);
        for (int q = 0; q < P * N; ++q) p.past[(size_t)env * P * N + q] = s_past[q * EPW + lane];
    }
    // warp-aggregated statistics
    const unsigned stepped = __ballot_sync(0xFFFFFFFFu, mine && active && !is_reset);
    const unsigned resets = __ballot_sync(0xFFFFFFFFu, mine && active && is_reset);
    if (lane == 0) {
        if (stepped) atomicAdd(p.stats + 0, (unsigned long long)__popc(stepped));
        if (resets) atomicAdd(p.stats + 1, (unsigned long long)__popc(resets));
    }
    __syncwarp();
    {  // coalesced write-back of the grids
        uint32_t *dst = p.obj + (size_t)env0 * TX;
#pragma unroll 4
        for (int i = lane; i < nenv * TX; i += 32) {
            const int e = (int)__umulhi((uint32_t)i, p.tx_magic), r = i - e * TX;
            int t = NSPEC ? 0 : (r >= X) + (r >= 2 * X) + (r >= 3 * X);
            if (!NSPEC && T > 4) t += (r >= 4 * X) + (r >= 5 * X) + (r >= 6 * X) + (r >= 7 * X);
            if (s_active[e] & 1) dst[i] = s_grid[e * p.gs + R + r + t * 2 * R] >> R;
        }
    }
    __syncwarp();
    }  // persistent group loop
}


using KernelFn = void (*)(const EnvParams);

template <int V, int EPW>
KernelFn pick_noise(int noise, int nspec) {
    switch (noise) {
    case MAEXP_NOISE_NONE:
        if (nspec == 8) return env_step_kernel<V, EPW, MAEXP_NOISE_NONE, 8>;
        if (nspec == 4) return env_step_kernel<V, EPW, MAEXP_NOISE_NONE, 4>;
        if (nspec == 2) return env_step_kernel<V, EPW, MAEXP_NOISE_NONE, 2>;
        return env_step_kernel<V, EPW, MAEXP_NOISE_NONE, 0>;
    case MAEXP_NOISE_FLIP: return env_step_kernel<V, EPW, MAEXP_NOISE_FLIP, 0>;
    case MAEXP_NOISE_DIST: return env_step_kernel<V, EPW, MAEXP_NOISE_DIST, 0>;
    case MAEXP_NOISE_THRESHOLD_DIST: return env_step_kernel<V, EPW, MAEXP_NOISE_THRESHOLD_DIST, 0>;
    case MAEXP_NOISE_FLAT: return env_step_kernel<V, EPW, MAEXP_NOISE_FLAT, 0>;
    }
    return nullptr;
}
template <int V>
KernelFn pick_epw(int epw, int noise, int nspec) {
    switch (epw) {
    case 32: return pick_noise<V, 32>(noise, nspec);
    case 16: return pick_noise<V, 16>(noise, nspec);
    case 8: return pick_noise<V, 8>(noise, nspec);
    }
    return nullptr;
}
KernelFn pick_v5(int epw, int noise, int nspec);
KernelFn pick_v7(int epw, int noise, int nspec);
KernelFn pick_v9(int epw, int noise, int nspec);
// env_lanes.cu: one lane per (environment, agent); nullptr for shapes it does not take
KernelFn pick_lanes(int V, int N);
size_t lanes_warp_words(int N, int V, int gs, int as);

}  // namespace maexp
