// Batched multi-agent gridworld step + egocentric observation builder for sm_100a.
//
// Replaces, for E environments at once, the reference's per-step Python path
//   DefaultEnvironment.step            core/environments/default_environment.py:153-175
//   AbstractWorld.step                 core/worlds/abstract_world.py:39-78
//   reward_ind / done_ind              default_environment.py:196-237
//   ObservationHandler.observation_ind core/handlers/observations/observation_handler.py:67-105
//   agent / object views, noises       core/handlers/observations/{agents,objects,noises}/*.py
//   DefaultEnvironment.reset           default_environment.py:64-116
//
// Design: the path is HBM-bound integer work, no tensor cores.
//  * State lives in HBM bit-packed: one uint32 row per map x (bit y) per object type, agents as
//    (x,y) bytes.  A CTA stages its environments' grids in shared memory with coalesced loads.
//  * Phase A, one LANE per environment: the order-dependent logic the reference runs
//    sequentially (moves in `order`, pickups in agent-index order with immediate respawn by
//    rejection sampling, per-agent done) on the shared-memory grid.
//  * Phase B, whole warp, once per agent index i (right after agent i's pickup, which is the
//    interleave of default_environment.py:165-168): every lane has packed agent i's C*V*V
//    observation bits in OUTPUT order; the warp expands them to fp32 and streams them with
//    128-bit `st.global.cs` stores into the replay-ring slot (16-byte aligned agent slots).
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <string>
#include <vector>

#include "../../include/maexp_b200.h"
#include "philox.cuh"

namespace maexp {

thread_local std::string g_last_error;
int fail(const std::string &msg) {
    g_last_error = msg;
    return -1;
}
#define MAEXP_CUDA(call)                                                                          \
    do {                                                                                          \
        cudaError_t err__ = (call);                                                               \
        if (err__ != cudaSuccess)                                                                 \
            return fail(std::string(#call) + ": " + cudaGetErrorString(err__));                   \
    } while (0)

constexpr int kWarpsPerCta = 4;
constexpr int kRespawnCap = 1 << 16;  // the reference loop is unbounded

struct EnvParams {
    // shape
    int E, N, T, X, Y, P, C, K, slot_floats, qps, wpa, ws, gs, area_bit0, M;
    uint32_t qps_magic;
    int agent_view, object_view, merged, noise, reward_mode, shared_max_is_mean;
    int L, num_objects, keep, shuffle_init, auto_reset, random_dest;
    uint32_t seed, gid0, tick;
    int force_reset;
    const uint8_t *env_mask;
    // static tables
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
    int32_t *error_flag;
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

template <int V>
__device__ __forceinline__ void put_bits(uint32_t *w, int bit0, typename Win<V>::T m) {
    const int i = bit0 >> 5, s = bit0 & 31;
    if constexpr (V == 9) {
        const unsigned __int128 v = m << s;  // 81 + 31 bits
        w[i] |= (uint32_t)v;
        w[i + 1] |= (uint32_t)(v >> 32);
        w[i + 2] |= (uint32_t)(v >> 64);
        w[i + 3] |= (uint32_t)(v >> 96);
    } else {
        w[i] |= (uint32_t)(m << s);
        const uint64_t r = (m >> (31 - s)) >> 1;  // m >> (32 - s) without a 32-bit-wide shift
        w[i + 1] |= (uint32_t)r;
        w[i + 2] |= (uint32_t)(r >> 32);
    }
}

template <int V, int EPW, int NOISE>
__global__ void __launch_bounds__(kWarpsPerCta * 32) env_step_kernel(const EnvParams p) {
    constexpr int R = V / 2, VV = V * V;
    constexpr int LW = (V == 9) ? 2 : 1;
    using WinT = typename Win<V>::T;
    extern __shared__ __align__(16) uint32_t smem[];

    const int N = p.N, T = p.T, X = p.X, Y = p.Y, P = p.P, TX = T * X;
    // ---- shared memory carve-up (must match env_smem_bytes) ----
    uint64_t *s_los = reinterpret_cast<uint64_t *>(smem);
    uint32_t *s_wall = reinterpret_cast<uint32_t *>(s_los + (size_t)X * Y * LW);
    uint32_t *s_noise = s_wall + X;
    uint32_t *s_area = s_noise + X;
    uint32_t *s_dest = s_area + TX;
    uint32_t *s_warp0 = s_dest + p.M * X;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int per_warp_words = EPW * p.gs + EPW * p.ws + (N * EPW + P * N * EPW + 1) / 2 + (EPW + 3) / 4;
    uint32_t *s_grid = s_warp0 + (size_t)warp * per_warp_words;
    uint32_t *s_w = s_grid + EPW * p.gs;
    uint16_t *s_pos = reinterpret_cast<uint16_t *>(s_w + EPW * p.ws);
    uint16_t *s_past = s_pos + N * EPW;
    uint8_t *s_active = reinterpret_cast<uint8_t *>(s_w + EPW * p.ws + (N * EPW + P * N * EPW + 1) / 2);

    for (int i = threadIdx.x; i < X * Y * LW; i += blockDim.x) s_los[i] = p.los[i];
    for (int i = threadIdx.x; i < X; i += blockDim.x) {
        s_wall[i] = p.wall[i];
        s_noise[i] = p.noise_area[i];
    }
    for (int i = threadIdx.x; i < TX; i += blockDim.x) s_area[i] = p.area[i];
    for (int i = threadIdx.x; i < p.M * X; i += blockDim.x) s_dest[i] = p.dest[i];

    const int env0 = (blockIdx.x * kWarpsPerCta + warp) * EPW;
    const int nenv = min(EPW, p.E - env0);
    if (nenv > 0) {  // coalesced stage-in of this warp's grids
        const uint32_t *src = p.obj + (size_t)env0 * TX;
        for (int i = lane; i < nenv * TX; i += 32) {
            const int e = i / TX, r = i - e * TX;
            s_grid[e * p.gs + r] = src[i];
        }
    }
    __syncthreads();
    if (nenv <= 0) return;

    // ---------------------------------------------------------------- phase A state (lane = env)
    const bool mine = lane < nenv;
    const int env = env0 + lane;
    const uint32_t gid = p.gid0 + (uint32_t)env;
    uint32_t *grid = s_grid + lane * p.gs;
    uint32_t *wv = s_w + lane * p.ws;
    uint16_t *mypos = s_pos + lane;  // agent k at mypos[k*EPW]
    int stepc = 0, nobj = 0;
    int c_gen = 0, c_done = 0, c_ca = 0, c_cw = 0;
    uint32_t ca = 0, cw = 0, rew_pos = 0, rew_neg = 0, done_bits = 0;
    uint64_t dsel = 0;
    bool is_reset = false, active = false;
    uint32_t attempt = 0;
    uint4 rw = make_uint4(0, 0, 0, 0);

    auto agent_at = [&](int x, int y) {
        const uint16_t key = (uint16_t)(x | (y << 8));
        bool hit = false;
        for (int k = 0; k < N; ++k) hit |= (mypos[k * EPW] == key);
        return hit;
    };
    // default_environment.py:134-148, one accepted object; two uniforms per attempt
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
            const uint32_t bit = 1u << y;
            bool ok = !(s_wall[x] & bit) && (s_area[t * X + x] & bit);
            for (int t2 = 0; t2 < T; ++t2) ok = ok && !(grid[t2 * X + x] & bit);
            ok = ok && !agent_at(x, y);
            if (ok) {
                grid[t * X + x] |= bit;
                ++c_gen;
                ++nobj;
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

        if (is_reset) {
            // ------------------------------------------------ default_environment.py:64-116
            c_gen = c_done = c_ca = c_cw = 0;
            stepc = 0;
            for (int r = 0; r < TX; ++r) grid[r] = 0;
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
            for (int k = 0; k < N; ++k) perm |= (uint64_t)p.init_perm[(size_t)env * N + k] << (4 * k);
            if (p.shuffle_init) {
                uint32_t d = 0;
                perm = fisher_yates(perm, N, p.seed, gid, p.tick, kStreamInitPerm, d);
                for (int k = 0; k < N; ++k) p.init_perm[(size_t)env * N + k] = (uint8_t)nib(perm, k);
            }
            for (int k = 0; k < N; ++k) mypos[k * EPW] = p.init_xy[nib(perm, k)];
            attempt = 0;
            nobj = 0;
            for (int t = 0; t < T; ++t)
                for (int n = 0; n < p.num_objects; ++n) spawn(t, kStreamObjects);
            for (int q = 0; q < P; ++q)  // transition_agent_observation.py:76-79
                for (int k = 0; k < N; ++k) s_past[(q * N + k) * EPW + lane] = mypos[k * EPW];
            attempt = 0;
        } else {
            // ------------------------------------------------ abstract_world.py:39-78
            for (int r = 0; r < TX; ++r) nobj += __popc(grid[r]);
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
                const uint16_t key = (uint16_t)(nx | (ny << 8));
                bool move = true;
                for (int k = 0; k < N; ++k)
                    if (k != id && mypos[k * EPW] == key) {
                        ca |= 1u << id;
                        move = false;
                    }
                if (s_wall[nx] & (1u << ny)) {
                    cw |= 1u << id;
                    move = false;
                }
                if (move) mypos[id * EPW] = key;
            }
        }
    }
    __syncwarp();

    // ------------------------------------------------------------ per-agent interleave
    for (int i = 0; i < N; ++i) {
        bool gate = false;
        if (mine && active) {
            const uint16_t me = mypos[i * EPW];
            const int px = me & 0xFF, py = me >> 8;
            if (!is_reset) {
                // ---------------------------------------- default_environment.py:196-231
                if (!p.no_task[i]) {
                    const int sel = p.random_dest ? (int)nib(dsel, i) : (i < p.M ? i : p.M - 1);
                    const uint32_t bit = 1u << py;
                    for (int q = 0; q < MAEXP_MAX_TYPES && p.task_order[i][q] >= 0; ++q) {
                        const int t = p.task_order[i][q];
                        if ((grid[t * X + px] & bit) && (s_dest[sel * X + px] & bit)) {
                            rew_pos |= 1u << i;
                            grid[t * X + px] &= ~bit;
                            --nobj;
                            ++c_done;
                            if (p.keep) spawn(t, kStreamRespawn);
                        }
                    }
                    if (ca & (1u << i)) { rew_neg |= 1u << i; ++c_ca; }
                    if (cw & (1u << i)) { rew_neg |= 1u << i; ++c_cw; }
                }
                // ---------------------------------------- default_environment.py:233-237
                if (nobj == 0 || stepc == p.L - 1) done_bits |= 1u << i;
            }
            // -------------------------------------------- observation_handler.py:67-105, bits in output order
            for (int w = 0; w < p.wpa + 3; ++w) wv[w] = 0;
            WinT los = s_los[(px * Y + py) * LW];
            if constexpr (V == 9) los |= (WinT)s_los[(px * Y + py) * LW + 1] << 64;
            int ch = 0;
            auto win_bit = [&](uint16_t cell) -> WinT {  // cell → window bit or 0
                const int a = (int)(cell & 0xFF) - px + R, b = (int)(cell >> 8) - py + R;
                return (a >= 0 && a < V && b >= 0 && b < V) ? ((WinT)1 << (a * V + b)) : (WinT)0;
            };
            const int av = p.agent_view;
            if (av == MAEXP_AV_INDIVIDUAL || av == MAEXP_AV_INDIVIDUAL_TRANSITION)
                for (int k = 0; k < N; ++k) put_bits<V>(wv, (ch++) * VV, win_bit(mypos[k * EPW]) & los);
            if (av != MAEXP_AV_INDIVIDUAL) {
                WinT m = 0;
                for (int k = 0; k < N; ++k) m |= win_bit(mypos[k * EPW]);
                if (av != MAEXP_AV_SIMPLE) m &= los;
                put_bits<V>(wv, (ch++) * VV, m);
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
                    put_bits<V>(wv, (ch++) * VV, m);
                }
            auto obj_window = [&](int t) -> WinT {
                WinT m = 0;
#pragma unroll
                for (int a = 0; a < V; ++a) {
                    const int gx = px - R + a;
                    uint32_t row = (gx >= 0 && gx < X) ? grid[t * X + gx] : 0u;
                    const int sh = py - R;
                    row = sh >= 0 ? (row >> sh) : (row << (-sh));
                    m |= (WinT)(row & ((1u << V) - 1u)) << (a * V);
                }
                return m;
            };
            if (p.object_view == MAEXP_OV_TYPES) {
                for (int t = 0; t < T; ++t) put_bits<V>(wv, (ch++) * VV, obj_window(t) & los);
            } else {
                WinT m = 0;
                for (int t = 0; t < T; ++t) m |= obj_window(t);
                if (p.object_view == MAEXP_OV_DEFAULT) m &= los;
                put_bits<V>(wv, (ch++) * VV, m);
            }
            put_bits<V>(wv, ch * VV, los);  // observation_handler.py:107-108 (value = bit - 1)
            gate = (s_noise[px] >> py) & 1u;

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
            if (p.out.objects) {
                int8_t *o = p.out.objects + ((size_t)env * N + i) * TX * Y;
                for (int r = 0; r < TX; ++r)
                    for (int y = 0; y < Y; ++y) o[r * Y + y] = (int8_t)((grid[r] >> y) & 1u);
            }
        }
        if constexpr (NOISE != MAEXP_NOISE_NONE) {
            if (mine) s_active[lane] = (uint8_t)((active ? 1 : 0) | (gate ? 2 : 0));
        }
        __syncwarp();

        // ---------------------------------------------------------- phase B: expand + stream out
        const int total = nenv * p.qps;
        for (int f = lane; f < total; f += 32) {
            const int e = (int)__umulhi((uint32_t)f, p.qps_magic);
            const int q = f - e * p.qps;
            const uint8_t flags = s_active[e];
            if (!(flags & 1)) continue;
            const uint32_t bits = (s_w[e * p.ws + (q >> 3)] >> ((q & 7) * 4)) & 0xFu;
            const int k0 = q * 4;
            float v[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int k = k0 + j;
                v[j] = (float)((bits >> j) & 1u) - ((k >= p.area_bit0 && k < p.K) ? 1.0f : 0.0f);
            }
            const size_t slot = ((size_t)(env0 + e) * N + i) * p.slot_floats + k0;
            if constexpr (NOISE == MAEXP_NOISE_NONE) {
                const float4 o = make_float4(v[0], v[1], v[2], v[3]);
                if (p.out.obs) __stcs(reinterpret_cast<float4 *>(p.out.obs + slot), o);
                if (p.merged && p.out.obs_rel) __stcs(reinterpret_cast<float4 *>(p.out.obs_rel + slot), o);
            } else {
                if (p.merged && p.out.obs_rel)  // merged_observtion_handler.py:27: relative stays clean
                    __stcs(reinterpret_cast<float4 *>(p.out.obs_rel + slot), make_float4(v[0], v[1], v[2], v[3]));
                if (flags & 2) {
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
                        const int k = k0 + j;
                        if (k >= p.K) continue;
                        const int c = k / VV, rem = k - c * VV, a = rem / V, b = rem - a * V;
                        const int ring = min(min(min(a, b), min(V - 1 - a, V - 1 - b)), V / 2 - 1);
                        float x = v[j];
                        if constexpr (NOISE == MAEXP_NOISE_FLIP) {  // flip_observation_noise.py:28-46
                            if (c == 2) x = __fadd_rn(x, 1.0f);
                            x = __fadd_rn(x, __fmul_rn(dr[j], p.flip_range[ring]));
                            x = fabsf(rintf(x));
                            if (c == 2) x = __fadd_rn(x, -1.0f);
                        } else if constexpr (NOISE == MAEXP_NOISE_DIST) {  // dist_observation_noise.py:29-41
                            x = __fadd_rn(x, __fmul_rn(dr[j], p.std_dist[ring]));
                        } else if constexpr (NOISE == MAEXP_NOISE_THRESHOLD_DIST) {  // :53-75
                            const float g = __fmul_rn(dr[j], p.std_dist[ring]), th = p.noise_thr[ring];
                            float n = 0.0f;
                            if (g < -th) n = g;
                            if (th < g) n = g;
                            x = __fadd_rn(x, n);
                        } else {  // flat_observation_noise.py:28-34 (3 hard-coded channels)
                            if (c < 3) x = __fadd_rn(x, __fmul_rn(dr[j], 0.03f));
                        }
                        v[j] = x;
                    }
                }
                if (p.out.obs)
                    __stcs(reinterpret_cast<float4 *>(p.out.obs + slot), make_float4(v[0], v[1], v[2], v[3]));
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
        }
        float shared = 0.0f;
        if (p.reward_mode != MAEXP_REWARD_DEFAULT) {  // shared_reward_environment.py:15-28
            const int sum = __popc(rew_pos & ~rew_neg) - __popc(rew_neg);
            const bool mean = p.reward_mode == MAEXP_REWARD_SHARED_MEAN || p.shared_max_is_mean;
            if (mean) shared = (float)((double)sum / (double)N);
            else shared = (rew_pos & ~rew_neg) ? 1.0f : ((rew_neg == ((1u << N) - 1u)) ? -1.0f : 0.0f);
        }
        for (int k = 0; k < N; ++k) {
            float r = (rew_neg >> k) & 1u ? -1.0f : ((rew_pos >> k) & 1u ? 1.0f : 0.0f);
            if (p.reward_mode != MAEXP_REWARD_DEFAULT) r = shared;
            if (is_reset) r = 0.0f;
            if (p.out.rewards) p.out.rewards[(size_t)env * N + k] = r;
            if (p.out.dones) p.out.dones[(size_t)env * N + k] = (uint8_t)((done_bits >> k) & 1u);
            if (p.out.positions) {
                p.out.positions[((size_t)env * N + k) * 2 + 0] = (int8_t)(mypos[k * EPW] & 0xFF);
                p.out.positions[((size_t)env * N + k) * 2 + 1] = (int8_t)(mypos[k * EPW] >> 8);
            }
            if (p.out.destination)
                p.out.destination[(size_t)env * N + k] =
                    (uint8_t)(p.random_dest ? nib(dsel, k) : (k < p.M ? k : p.M - 1));
            p.pos[(size_t)env * N + k] = mypos[k * EPW];
        }
        if (p.out.was_reset) p.out.was_reset[env] = is_reset ? 1 : 0;
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
        for (int i = lane; i < nenv * TX; i += 32) {
            const int e = i / TX, r = i - e * TX;
            if (s_active[e] & 1) dst[i] = s_grid[e * p.gs + r];
        }
    }
}

// ------------------------------------------------------------------------ state taps
__global__ void get_state_kernel(const EnvParams p, int8_t *pos, int8_t *objects, int32_t *counters, int32_t *steps,
                                 int8_t *destination) {
    const int env = blockIdx.x * blockDim.x + threadIdx.x;
    if (env >= p.E) return;
    const int N = p.N, TX = p.T * p.X, Y = p.Y, X = p.X;
    if (pos)
        for (int k = 0; k < N; ++k) {
            const uint16_t c = p.pos[(size_t)env * N + k];
            pos[((size_t)env * N + k) * 2 + 0] = (int8_t)(c & 0xFF);
            pos[((size_t)env * N + k) * 2 + 1] = (int8_t)(c >> 8);
        }
    if (objects)
        for (int r = 0; r < TX; ++r) {
            const uint32_t row = p.obj[(size_t)env * TX + r];
            for (int y = 0; y < Y; ++y) objects[((size_t)env * TX + r) * Y + y] = (int8_t)((row >> y) & 1u);
        }
    if (counters)
        for (int j = 0; j < 4; ++j) counters[(size_t)env * 4 + j] = p.counters[(size_t)env * 4 + j];
    if (steps) steps[env] = p.step[env];
    if (destination)
        for (int k = 0; k < N; ++k) {
            const int sel = p.random_dest ? p.dest_sel[(size_t)env * N + k] : (k < p.M ? k : p.M - 1);
            for (int x = 0; x < X; ++x)
                for (int y = 0; y < Y; ++y)
                    destination[(((size_t)env * N + k) * X + x) * Y + y] = (int8_t)((p.dest[sel * X + x] >> y) & 1u);
        }
}

__global__ void set_state_kernel(const EnvParams p, const int8_t *pos, const int8_t *objects) {
    const int env = blockIdx.x * blockDim.x + threadIdx.x;
    if (env >= p.E) return;
    const int N = p.N, TX = p.T * p.X, Y = p.Y;
    if (pos)
        for (int k = 0; k < N; ++k)
            p.pos[(size_t)env * N + k] = (uint16_t)((uint8_t)pos[((size_t)env * N + k) * 2] |
                                                    ((uint8_t)pos[((size_t)env * N + k) * 2 + 1] << 8));
    if (objects)
        for (int r = 0; r < TX; ++r) {
            uint32_t row = 0;
            for (int y = 0; y < Y; ++y) row |= (uint32_t)(objects[((size_t)env * TX + r) * Y + y] != 0) << y;
            p.obj[(size_t)env * TX + r] = row;
        }
}

__global__ void init_state_kernel(const EnvParams p) {
    const int env = blockIdx.x * blockDim.x + threadIdx.x;
    if (env >= p.E) return;
    for (int k = 0; k < p.N; ++k) {
        p.init_perm[(size_t)env * p.N + k] = (uint8_t)k;  // default_environment.py:33
        p.dest_sel[(size_t)env * p.N + k] = 0;
        p.pos[(size_t)env * p.N + k] = p.init_xy[k];
    }
}

}  // namespace maexp

using namespace maexp;

struct maexp_env {
    maexp_env_cfg cfg;
    EnvParams p;
    int V, epw;
    size_t smem_bytes;
    std::vector<void *> allocs;
    // pinned-free staging for the host-buffer entry point
    int8_t *d_actions = nullptr;
    float *d_rewards = nullptr;
    uint8_t *d_dones = nullptr, *d_was_reset = nullptr;
};

namespace {

template <typename Tp>
int dev_alloc(maexp_env *h, Tp **ptr, size_t count, bool zero = true) {
    void *d = nullptr;
    const size_t bytes = (count ? count : 1) * sizeof(Tp);
    MAEXP_CUDA(cudaMalloc(&d, bytes));
    if (zero) MAEXP_CUDA(cudaMemset(d, 0, bytes));
    h->allocs.push_back(d);
    *ptr = static_cast<Tp *>(d);
    return 0;
}

template <typename Tp>
int upload(maexp_env *h, const Tp **ptr, const std::vector<Tp> &host) {
    Tp *d = nullptr;
    if (dev_alloc(h, &d, host.size(), false)) return -1;
    MAEXP_CUDA(cudaMemcpy(d, host.data(), host.size() * sizeof(Tp), cudaMemcpyHostToDevice));
    *ptr = d;
    return 0;
}

size_t env_smem_bytes(const EnvParams &p, int V, int epw) {
    const int LW = (V == 9) ? 2 : 1;
    size_t words = (size_t)p.X * p.Y * LW * 2 + p.X + p.X + (size_t)p.T * p.X + (size_t)p.M * p.X;
    const size_t per_warp = (size_t)epw * p.gs + (size_t)epw * p.ws + ((size_t)p.N * epw + (size_t)p.P * p.N * epw + 1) / 2 +
                            (epw + 3) / 4;
    words += per_warp * kWarpsPerCta;
    return words * 4;
}

using KernelFn = void (*)(const EnvParams);

template <int V, int EPW>
KernelFn pick_noise(int noise) {
    switch (noise) {
    case MAEXP_NOISE_NONE: return env_step_kernel<V, EPW, MAEXP_NOISE_NONE>;
    case MAEXP_NOISE_FLIP: return env_step_kernel<V, EPW, MAEXP_NOISE_FLIP>;
    case MAEXP_NOISE_DIST: return env_step_kernel<V, EPW, MAEXP_NOISE_DIST>;
    case MAEXP_NOISE_THRESHOLD_DIST: return env_step_kernel<V, EPW, MAEXP_NOISE_THRESHOLD_DIST>;
    case MAEXP_NOISE_FLAT: return env_step_kernel<V, EPW, MAEXP_NOISE_FLAT>;
    }
    return nullptr;
}
template <int V>
KernelFn pick_epw(int epw, int noise) {
    switch (epw) {
    case 32: return pick_noise<V, 32>(noise);
    case 16: return pick_noise<V, 16>(noise);
    case 8: return pick_noise<V, 8>(noise);
    }
    return nullptr;
}
KernelFn pick_kernel(int V, int epw, int noise) {
    switch (V) {
    case 5: return pick_epw<5>(epw, noise);
    case 7: return pick_epw<7>(epw, noise);
    case 9: return pick_epw<9>(epw, noise);
    }
    return nullptr;
}

int launch(maexp_env *h, uint32_t tick, int force_reset, const uint8_t *env_mask, const int8_t *actions,
           const uint8_t *order, const float *noise_tape, const maexp_obs_out *out, cudaStream_t stream) {
    EnvParams p = h->p;
    p.tick = tick;
    p.force_reset = force_reset;
    p.env_mask = env_mask;
    p.actions = actions;
    p.order = order;
    p.noise_tape = noise_tape;
    if (out) p.out = *out;
    else memset(&p.out, 0, sizeof(p.out));
    KernelFn fn = pick_kernel(h->V, h->epw, p.noise);
    if (!fn) return fail("maexp: no kernel for this visible_range / noise");
    const int envs_per_cta = h->epw * kWarpsPerCta;
    const int grid = (p.E + envs_per_cta - 1) / envs_per_cta;
    MAEXP_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem_bytes));
    fn<<<grid, kWarpsPerCta * 32, h->smem_bytes, stream>>>(p);
    MAEXP_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace

extern "C" {

const char *maexp_last_error(void) { return g_last_error.c_str(); }
int maexp_version(void) { return 1; }

int maexp_env_create(const maexp_env_cfg *cfg, const maexp_tables *tb, maexp_env **out) {
    if (!cfg || !tb || !out) return fail("maexp_env_create: null argument");
    const int E = cfg->num_envs, N = cfg->num_agents, T = cfg->type_objects, X = cfg->size_x, Y = cfg->size_y;
    const int V = cfg->visible_range, P = cfg->past_step, M = cfg->num_destinations;
    if (E < 1) return fail("maexp_env_create: num_envs < 1");
    if (N < 1 || N > MAEXP_MAX_AGENTS) return fail("maexp_env_create: num_agents must be in 1..16");
    if (T < 1 || T > MAEXP_MAX_TYPES) return fail("maexp_env_create: type_objects must be in 1..4");
    if (X < 3 || X > 255 || Y < 3 || Y > 32) return fail("maexp_env_create: map must satisfy 3<=X<=255, 3<=Y<=32");
    if (V != 5 && V != 7 && V != 9) return fail("maexp_env_create: visible_range must be 5, 7 or 9");
    if (M < 1 || M > 15) return fail("maexp_env_create: num_destinations must be in 1..15");
    if (!tb->wall || !tb->objects_area || !tb->noise_area || !tb->destination || !tb->los)
        return fail("maexp_env_create: missing table");
    const bool transition = cfg->agent_view == MAEXP_AV_TRANSITION || cfg->agent_view == MAEXP_AV_INDIVIDUAL_TRANSITION;
    if (transition && (P < 1 || P > 8)) return fail("maexp_env_create: transition views need 1 <= past_step <= 8");
    if (cfg->noise == MAEXP_NOISE_FLAT) {
        /* flat_observation_noise.py:30 hard-codes a 3-channel draw */
    }
    maexp_env *h = new maexp_env();
    h->cfg = *cfg;
    h->V = V;
    EnvParams &p = h->p;
    memset(&p, 0, sizeof(p));
    p.E = E; p.N = N; p.T = T; p.X = X; p.Y = Y; p.P = transition ? P : 0; p.M = M;
    int agent_ch = 1;
    switch (cfg->agent_view) {
    case MAEXP_AV_DEFAULT: case MAEXP_AV_SIMPLE: agent_ch = 1; break;
    case MAEXP_AV_INDIVIDUAL: agent_ch = N; break;
    case MAEXP_AV_TRANSITION: agent_ch = 1 + P; break;
    case MAEXP_AV_INDIVIDUAL_TRANSITION: agent_ch = N + 1 + P; break;
    default: delete h; return fail("maexp_env_create: unknown agent_view");
    }
    if (cfg->object_view < 0 || cfg->object_view > 2) { delete h; return fail("maexp_env_create: unknown object_view"); }
    const int object_ch = cfg->object_view == MAEXP_OV_TYPES ? T : 1;
    p.C = agent_ch + object_ch + 1;  // observation_handler.py:56-61
    p.K = p.C * V * V;
    if (cfg->noise == MAEXP_NOISE_FLAT && p.C != 3) { delete h; return fail("maexp_env_create: flat noise needs C == 3 (flat_observation_noise.py:30)"); }
    if (cfg->noise != MAEXP_NOISE_NONE && cfg->area_mask == MAEXP_MASK_RELATIVE) {
        delete h;
        return fail("maexp_env_create: observation noise is undefined for observation_area_mask=relative (SURVEY.md appendix C)");
    }
    p.slot_floats = (p.K + 3) & ~3;
    p.qps = p.slot_floats / 4;
    p.qps_magic = (uint32_t)((0x100000000ull + p.qps - 1) / p.qps);
    p.wpa = (p.K + 31) / 32;
    p.ws = (p.wpa + 3) | 1;
    p.gs = (T * X) | 1;
    p.area_bit0 = (p.C - 1) * V * V;
    p.agent_view = cfg->agent_view; p.object_view = cfg->object_view;
    p.merged = cfg->area_mask == MAEXP_MASK_MERGED;
    p.noise = cfg->noise; p.reward_mode = cfg->reward_mode; p.shared_max_is_mean = cfg->shared_max_is_mean;
    p.L = cfg->max_episode_length; p.num_objects = cfg->num_objects; p.keep = cfg->keep_objects_num;
    p.shuffle_init = cfg->shuffle_init_xys; p.auto_reset = cfg->auto_reset; p.random_dest = cfg->random_destinations;
    p.seed = cfg->seed; p.gid0 = cfg->first_global_env_id;
    for (int i = 0; i < N; ++i) {
        p.init_xy[i] = (uint16_t)(cfg->init_xys[i][0] | (cfg->init_xys[i][1] << 8));
        p.no_task[i] = cfg->task_mask[i] < 0;
        for (int q = 0; q < MAEXP_MAX_TYPES; ++q) p.task_order[i][q] = (int8_t)cfg->task_order[i][q];
    }
    for (int i = 0; i < MAEXP_MAX_RINGS; ++i) {
        p.flip_range[i] = cfg->flip_noise_range[i];
        p.std_dist[i] = cfg->std_distribution[i];
        p.noise_thr[i] = cfg->noise_threshold[i];
    }
    // magic-division self check over the range phase B uses
    for (int f = 0; f < 32 * p.qps; ++f)
        if ((int)(((uint64_t)f * p.qps_magic) >> 32) != f / p.qps) { delete h; return fail("maexp_env_create: internal magic division"); }

    // ---- pack static tables to bit rows ----
    auto rows = [&](const int8_t *src, int planes) {
        std::vector<uint32_t> r((size_t)planes * X, 0u);
        for (int pl = 0; pl < planes; ++pl)
            for (int x = 0; x < X; ++x)
                for (int y = 0; y < Y; ++y)
                    if (src[((size_t)pl * X + x) * Y + y]) r[(size_t)pl * X + x] |= 1u << y;
        return r;
    };
    const int LW = (V == 9) ? 2 : 1;
    std::vector<uint64_t> los((size_t)X * Y * LW, 0ull);
    for (int x = 0; x < X; ++x)
        for (int y = 0; y < Y; ++y)
            for (int a = 0; a < V; ++a)
                for (int b = 0; b < V; ++b)
                    if (tb->los[(((size_t)x * Y + y) * V + a) * V + b]) {
                        const int bit = a * V + b;
                        los[((size_t)x * Y + y) * LW + (bit >> 6)] |= 1ull << (bit & 63);
                    }
    int rc = 0;
    rc |= upload(h, &p.wall, rows(tb->wall, 1));
    rc |= upload(h, &p.noise_area, rows(tb->noise_area, 1));
    rc |= upload(h, &p.area, rows(tb->objects_area, T));
    rc |= upload(h, &p.dest, rows(tb->destination, M));
    rc |= upload(h, &p.los, los);
    // ---- dynamic state ----
    rc |= dev_alloc(h, &p.pos, (size_t)E * N);
    rc |= dev_alloc(h, &p.obj, (size_t)E * T * X);
    rc |= dev_alloc(h, &p.step, (size_t)E);
    rc |= dev_alloc(h, &p.pending, (size_t)E);
    rc |= dev_alloc(h, &p.counters, (size_t)E * 4);
    rc |= dev_alloc(h, &p.init_perm, (size_t)E * N);
    rc |= dev_alloc(h, &p.dest_sel, (size_t)E * N);
    rc |= dev_alloc(h, &p.past, (size_t)E * (p.P ? p.P : 1) * N);
    rc |= dev_alloc(h, &p.stats, 2);
    rc |= dev_alloc(h, &p.error_flag, 1);
    rc |= dev_alloc(h, &h->d_actions, (size_t)E * N);
    rc |= dev_alloc(h, &h->d_rewards, (size_t)E * N);
    rc |= dev_alloc(h, &h->d_dones, (size_t)E * N);
    rc |= dev_alloc(h, &h->d_was_reset, (size_t)E);
    if (rc) { maexp_env_destroy(h); return -1; }

    // envs per warp: keep >= ~24 warps per SM in flight when E allows it
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    h->epw = 32;
    while (h->epw > 8 && (E + h->epw - 1) / h->epw < sms * 24) h->epw >>= 1;
    h->smem_bytes = env_smem_bytes(p, V, h->epw);
    while (h->smem_bytes > 200 * 1024 && h->epw > 8) {
        h->epw >>= 1;
        h->smem_bytes = env_smem_bytes(p, V, h->epw);
    }
    if (h->smem_bytes > 227 * 1024) { maexp_env_destroy(h); return fail("maexp_env_create: configuration needs more than 227 KB shared memory per CTA"); }
    init_state_kernel<<<(E + 127) / 128, 128>>>(p);
    if (cudaDeviceSynchronize() != cudaSuccess) { maexp_env_destroy(h); return fail("maexp_env_create: init kernel failed"); }
    *out = h;
    return 0;
}

void maexp_env_destroy(maexp_env *h) {
    if (!h) return;
    for (void *d : h->allocs) cudaFree(d);
    delete h;
}

int maexp_env_channels(const maexp_env *h) { return h ? h->p.C : -1; }
int maexp_env_slot_floats(const maexp_env *h) { return h ? h->p.slot_floats : -1; }

int maexp_env_reset(maexp_env *h, uint32_t tick, const uint8_t *env_mask, const float *noise_tape,
                    const maexp_obs_out *out, void *stream) {
    if (!h) return fail("maexp_env_reset: null handle");
    return launch(h, tick, 1, env_mask, nullptr, nullptr, noise_tape, out, static_cast<cudaStream_t>(stream));
}

int maexp_env_step(maexp_env *h, uint32_t tick, const int8_t *actions, const uint8_t *order, const float *noise_tape,
                   const maexp_obs_out *out, void *stream) {
    if (!h) return fail("maexp_env_step: null handle");
    return launch(h, tick, 0, nullptr, actions, order, noise_tape, out, static_cast<cudaStream_t>(stream));
}

int maexp_env_step_host(maexp_env *h, uint32_t tick, const int8_t *actions_host, float *obs_ring_slot,
                        float *rewards_host, uint8_t *dones_host, uint8_t *was_reset_host, void *stream_) {
    if (!h) return fail("maexp_env_step_host: null handle");
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    const size_t EN = (size_t)h->p.E * h->p.N;
    if (actions_host) MAEXP_CUDA(cudaMemcpyAsync(h->d_actions, actions_host, EN, cudaMemcpyHostToDevice, stream));
    maexp_obs_out out;
    memset(&out, 0, sizeof(out));
    out.obs = obs_ring_slot;
    out.rewards = h->d_rewards;
    out.dones = h->d_dones;
    out.was_reset = h->d_was_reset;
    if (launch(h, tick, 0, nullptr, actions_host ? h->d_actions : nullptr, nullptr, nullptr, &out, stream)) return -1;
    if (rewards_host) MAEXP_CUDA(cudaMemcpyAsync(rewards_host, h->d_rewards, EN * sizeof(float), cudaMemcpyDeviceToHost, stream));
    if (dones_host) MAEXP_CUDA(cudaMemcpyAsync(dones_host, h->d_dones, EN, cudaMemcpyDeviceToHost, stream));
    if (was_reset_host) MAEXP_CUDA(cudaMemcpyAsync(was_reset_host, h->d_was_reset, h->p.E, cudaMemcpyDeviceToHost, stream));
    MAEXP_CUDA(cudaStreamSynchronize(stream));
    return 0;
}

int maexp_env_get_state(maexp_env *h, int8_t *pos, int8_t *objects, int32_t *counters, int32_t *steps,
                        int8_t *destination, void *stream) {
    if (!h) return fail("maexp_env_get_state: null handle");
    get_state_kernel<<<(h->p.E + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream)>>>(h->p, pos, objects, counters,
                                                                                         steps, destination);
    MAEXP_CUDA(cudaGetLastError());
    return 0;
}

int maexp_env_set_state(maexp_env *h, const int8_t *pos, const int8_t *objects, void *stream) {
    if (!h) return fail("maexp_env_set_state: null handle");
    set_state_kernel<<<(h->p.E + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream)>>>(h->p, pos, objects);
    MAEXP_CUDA(cudaGetLastError());
    return 0;
}

int maexp_env_stats(maexp_env *h, uint64_t *env_steps, uint64_t *resets, int32_t *error_flag, void *stream) {
    if (!h) return fail("maexp_env_stats: null handle");
    MAEXP_CUDA(cudaStreamSynchronize(static_cast<cudaStream_t>(stream)));
    unsigned long long s[2];
    int32_t e = 0;
    MAEXP_CUDA(cudaMemcpy(s, h->p.stats, sizeof(s), cudaMemcpyDeviceToHost));
    MAEXP_CUDA(cudaMemcpy(&e, h->p.error_flag, sizeof(e), cudaMemcpyDeviceToHost));
    if (env_steps) *env_steps = s[0];
    if (resets) *resets = s[1];
    if (error_flag) *error_flag = e;
    return 0;
}

}  // extern "C"
