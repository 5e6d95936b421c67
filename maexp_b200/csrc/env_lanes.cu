// Agent-parallel form of the batched gridworld step + observation builder (round 2): one LANE per (environment, agent).
//
// env_step_kernel (env_kernel.cuh) gives every environment one lane for the whole sequential phase, so half of the warp
// idles there (16 environments per warp) and every agent's window crop sits on one long per-lane chain.  The reference's
// order dependence is much narrower than that (default_environment.py:153-175, abstract_world.py:39-78):
//   * moves are sequential in `order`  -> every lane derives its agent's rank, target cell and wall test in parallel, then N short
//     rounds (one mover per environment and round) update the agents bit grid;
//   * a pickup happens on the agent's own cell, agents never share a cell and a respawned object never lands on an agent
//     (default_environment.py:134-148), so all N pickup tests are independent -> one test per lane, a ballot collects them;
//   * respawns draw from one per-environment stream and must not collide with each other -> the leader replays the (rare)
//     pickups in agent order and logs where each respawn landed;
//   * agent i observes the objects after pickups / respawns 0..i -> every lane crops the FINAL grid and undoes the logged
//     events of the agents behind it (pickup cells and respawn cells are disjoint sets, so the order of the undo is free);
//   * per-agent rewards, done flags, positions and the 147 observation bits are element-wise -> one lane each, coalesced.
// A warp owns 32 / N environments per work item (4 at N = 8; 4x more, shorter items than env_step_kernel's groups), items are
// dealt round-robin (an atomic queue -- one same-address fetch per item -- cost 10 us per step and stays behind
// MAEXP_LANES_QUEUE=1), the next item's state is requested before the current item's stores start, and the expansion phase is
// the same arithmetic 4-bit -> float4 expansion with 128-bit streaming stores.
// Shapes: N in {2, 4, 8}, one object type, the three default channels (agents, objects, area), visible_range 5 / 7, no
// observation noise; everything else (and calls that ask for heat-maps, coordinates or object snapshots) runs
// env_step_kernel on the same state.  Results are bit-identical to it (tests/test_env_gpu.py runs both).
#include "env_kernel.cuh"

namespace maexp {

namespace {

constexpr unsigned kFull = 0xFFFFFFFFu;

template <int N> struct LaneShape {
    static constexpr int EPW = 32 / N;                      // environments per warp
    static constexpr int NA = (N + 3) / 4, NO = (N + 2) / 4;  // Philox calls per env: actions, Fisher-Yates draws
    static constexpr int NC = NA + NO;
};

template <int V> struct SlotShape {
    static constexpr int K = 3 * V * V, SLOT = (K + 7) / 8 * 8, QPS = SLOT / 4, NW = (QPS + 7) / 8;
};

// 4 bits -> float4 of one quad: element j is bit_j - sub_j (env_kernel.cuh expand_group)
struct QuadConst { uint32_t sub, h0, h1, h2, h3; int sh; };
__device__ __forceinline__ QuadConst quad_consts(int q, int q_str, int q_tail, int sub_str, int sub_tail) {
    QuadConst c;
    const int cls = q > q_tail ? 0 : ((q >= q_str) + (q > q_str) + (q == q_tail));
    c.sub = cls == 0 ? 0u : (cls == 1 ? (uint32_t)sub_str : (cls == 2 ? 15u : (uint32_t)sub_tail));
    c.h0 = (0x3f800000u | ((c.sub & 1u) << 31));
    c.h1 = (0x3f800000u | (((c.sub >> 1) & 1u) << 31)) >> 1;
    c.h2 = (0x3f800000u | (((c.sub >> 2) & 1u) << 31)) >> 2;
    c.h3 = (0x3f800000u | (((c.sub >> 3) & 1u) << 31)) >> 3;
    c.sh = (q & 7) * 4;
    return c;
}
__device__ __forceinline__ float4 expand_quad(const QuadConst &c, uint32_t w) {
    const uint32_t x = w >> c.sh;
    return make_float4(__uint_as_float(((x ^ c.sub) & 1u) * c.h0), __uint_as_float(((x ^ c.sub) & 2u) * c.h1),
                       __uint_as_float(((x ^ c.sub) & 4u) * c.h2), __uint_as_float(((x ^ c.sub) & 8u) * c.h3));
}

}  // namespace

size_t lanes_warp_words(int N, int V, int gs, int as) {
    const int epw = 32 / N, nc = (N + 3) / 4 + (N + 2) / 4;
    const int qps = (3 * V * V + 7) / 8 * 8 / 4, nw = (qps + 7) / 8;
    // bit words | Philox words | object rows | agent rows | cells u16[32] | respawn cells u16[32] | order u8[32] | dest u8[32] | 2 words per env
    size_t words = (size_t)nw * 32 + (size_t)epw * nc * 4 + (size_t)epw * gs + (size_t)epw * as + 16 + 16 + 8 + 8 + 2 * epw;
    return (words + 3) & ~(size_t)3;
}

// Concatenation of bit fields into 32-bit words with compile-time offsets: the part of `v` (placed at bit `off` of the
// stream) that falls into word k.
template <int OFF, int K>
__device__ __forceinline__ uint32_t stream_piece(uint64_t v) {
    constexpr int s = OFF - 32 * K;
    if constexpr (s >= 32 || s <= -64) return 0u;
    else if constexpr (s >= 0) return (uint32_t)v << s;
    else return (uint32_t)(v >> (-s));
}

// V x V crop of a padded bit-row grid (rows shifted left by R) around column py, rows r0[0..V): bit a * V + b
template <int V>
__device__ __forceinline__ uint64_t crop_rows(const uint32_t *r0, int py) {
    constexpr int RPH = 32 / V;  // rows per 32-bit half
    constexpr uint32_t M = (1u << V) - 1u;
    uint32_t lo = 0, hi = 0;
#pragma unroll
    for (int k = (RPH < V ? RPH : V) - 1; k >= 0; --k) lo = (lo << V) | ((r0[k] >> py) & M);
    if constexpr (V > RPH) {
#pragma unroll
        for (int k = V - 1; k >= RPH; --k) hi = (hi << V) | ((r0[k] >> py) & M);
    }
    return (uint64_t)lo | ((uint64_t)hi << (RPH * V));
}

struct ItemState {  // global state of one work item, loaded one item ahead
    uint32_t rows[4];  // object rows, when the item's EPW * X words fit four coalesced rounds
    int step;
    // raw loads, untouched until the item starts: anything computed from them here would wait for the loads at request time
    uint16_t cell;
    uint8_t pend, masked, act, ord, dsel;
};

#ifndef MAEXP_LANES_UNROLL_A
#define MAEXP_LANES_UNROLL_A 8
#endif
// Resident CTAs per SM the register budget is sized for.  Measured on B200 at 65 536 envs x 8 agents: the step time does not
// depend on the occupancy between 4 and 8 CTAs per SM (63.9 / 64.1 / 64.7 / 65.7 / 65.6 us at 4 / 5 / 6 / 7 / 8 with 64
// registers) -- with more warps the logic phase speeds up (no-store build: 85 / 58 / 49 us at 2 / 4 / 8) but the store phase
// slows down by as much -- so N = 8 runs 4 CTAs per SM with a 5-CTA register budget (95 registers, 62.4 us); the
// short-chain shapes (N = 2, 4) keep 8 CTAs per SM (cfg1: 16.1 us at 8, 18.8 at 4).
#ifndef MAEXP_LANES_MINB
#define MAEXP_LANES_MINB (N == 8 ? 5 : 8)
#endif
constexpr int kUnrollA = MAEXP_LANES_UNROLL_A;  // agents per unrolled pass of the expansion loops
template <int V, int N>
__global__ void __launch_bounds__(kWarpsPerCta * 32, MAEXP_LANES_MINB) env_step_lanes_kernel(const EnvParams p) {
    using LS = LaneShape<N>;
    using SS = SlotShape<V>;
    constexpr int R = V / 2, VV = V * V, EPW = LS::EPW, NA = LS::NA, NC = LS::NC, SLOT = SS::SLOT, NW = SS::NW;
    constexpr uint32_t MASKN = (1u << N) - 1u;
    // phase-B quad classes of the three-channel slot (env.cu computes the same for the generic kernel)
    constexpr int Q_STR = (2 * VV) >> 2, Q_TAIL = (SS::K - 1) >> 2;
    constexpr int SUB_STR = 0xF & ~((1 << ((2 * VV) & 3)) - 1), SUB_TAIL = (1 << (SS::K - 4 * Q_TAIL)) - 1;
    extern __shared__ __align__(1024) uint32_t smem[];
    const int X = p.X, Y = p.Y;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int e = lane / N, a = lane % N, eshift = e * N;
    const bool lead = a == 0;

    const size_t per_warp = ((size_t)NW * 32 + EPW * NC * 4 + (size_t)EPW * p.gs + (size_t)EPW * p.as + 16 + 16 + 8 + 8 + 2 * EPW + 3) & ~(size_t)3;
    uint32_t *const s_w = smem + warp * per_warp;  // word k of pair (agent a, env e) at s_w[k * 32 + a * EPW + e]
    uint32_t *const s_rng = s_w + NW * 32;         // uint4 per (env, call)
    uint32_t *const s_grid = s_rng + EPW * NC * 4;
    uint32_t *const s_agrid = s_grid + EPW * p.gs;
    uint16_t *const s_pos = reinterpret_cast<uint16_t *>(s_agrid + EPW * p.as);  // cell of (env e, agent k) at s_pos[e * N + k]
    uint16_t *const s_ev = s_pos + 32;                                            // where agent k's respawn landed
    uint8_t *const s_ord = reinterpret_cast<uint8_t *>(s_ev + 32);
    uint8_t *const s_dsel = s_ord + 32;
    uint32_t *const s_flags = reinterpret_cast<uint32_t *>(s_dsel + 32);  // [EPW] pickup events of an env

    // pad rows (x < 0, x >= X) of this warp's grids: zero once, nothing writes them afterwards
    for (int i = lane; i < EPW * 2 * R; i += 32) {
        const int e2 = i / (2 * R), r = i % (2 * R);
        const int x = r < R ? r : X + r;
        s_grid[e2 * p.gs + x] = 0;
        s_agrid[e2 * p.as + x] = 0;
    }
    // per-lane constants: the lane's agent never changes
    bool task0 = false;
    const bool has_task = !p.no_task[a];
    for (int q = 0; q < MAEXP_MAX_TYPES && p.task_order[a][q] >= 0; ++q) task0 = task0 || p.task_order[a][q] == 0;
    task0 = task0 && has_task;
    const uint32_t task_mask = (__ballot_sync(kFull, has_task) >> eshift) & MASKN;
    const int fixed_sel = a < p.M ? a : p.M - 1;
    uint32_t *const grid = s_grid + e * p.gs + R;    // row x at grid[x], x in [-R, X + R), rows shifted left by R
    uint32_t *const agrid = s_agrid + e * p.as + R;
    // this lane's words of the coalesced object-row transfers: an item has EPW * X of them; up to 128 (four rounds) travel
    // through registers one item ahead, larger items are staged in place.  Env of word lane + 32 k in bits 4k..4k+3,
    // 15 = past the end
    const bool rows_ahead = EPW <= 8 && EPW * X <= 128;
    uint32_t row_env = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int i = lane + 32 * k;
        row_env |= (uint32_t)(i < EPW * X ? (i / X) & 15 : 15) << (4 * k);
    }
    const int row_skip = p.gs - X;  // row x of env e2 sits at word i + R + e2 * row_skip of the staged grids (i = e2 * X + x)
    __syncwarp();

    const int num_groups = (p.E + EPW - 1) / EPW, nwarps = gridDim.x * kWarpsPerCta;
    auto load_item = [&](int g) {
        ItemState st;
        const int env = g * EPW + e;
        const uint32_t *src = p.obj + (size_t)g * EPW * X;
        const int nrows = min(EPW, p.E - g * EPW) * X;
#pragma unroll
        for (int k = 0; k < 4; ++k) st.rows[k] = (rows_ahead && lane + 32 * k < nrows) ? src[lane + 32 * k] : 0u;
        st.step = 0; st.cell = 0; st.pend = 0; st.masked = 1; st.act = 0; st.ord = 0; st.dsel = 0;
        if (env < p.E) {
            const size_t o = (size_t)env * N + a;
            st.step = p.step[env];
            st.cell = p.pos[o];
            st.pend = p.pending[env];
            if (p.env_mask) st.masked = p.env_mask[env];
            if (p.actions) st.act = (uint8_t)p.actions[o];
            if (p.order) st.ord = p.order[o];
            if (p.random_dest) st.dsel = p.dest_sel[o];
        }
        return st;
    };
    int group = blockIdx.x * kWarpsPerCta + warp;
    ItemState cur;
    if (group < num_groups) cur = load_item(group);
    int n_stepped = 0, n_resets = 0;  // statistics of this warp's items, added once at the end
    while (group < num_groups) {
        int next_group = group + nwarps;  // static round-robin, or (p.queue) the next item of the atomic queue
        if (p.queue && lane == 0) next_group = nwarps + (int)atomicAdd(p.queue, 1u);  // needed after the logic phase only
        const int env0 = group * EPW, env = env0 + e;
        const int nenv = min(EPW, p.E - env0);
        const bool mine = env < p.E;
        const uint32_t gid = p.gid0 + (uint32_t)env;
        if (rows_ahead) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int e2 = (int)((row_env >> (4 * k)) & 15u);
                if (e2 != 15) {  // T = 1: the object and the agent grids have the same pitch
                    const int w = lane + 32 * k + R + e2 * row_skip;
                    s_grid[w] = cur.rows[k] << R;
                    s_agrid[w] = 0u;
                }
            }
        } else {
            const uint32_t *src = p.obj + (size_t)env0 * X;
            for (int i = lane; i < EPW * X; i += 32) {
                const int e2 = (int)__umulhi((uint32_t)i, p.tx_magic), w = i + R + e2 * row_skip;
                s_grid[w] = i < nenv * X ? src[i] << R : 0u;
                s_agrid[w] = 0u;
            }
        }
        bool is_reset = false, active = false;
        if (mine) {
            const bool masked = cur.masked != 0, pend = cur.pend != 0;
            is_reset = p.force_reset ? masked : (p.auto_reset && pend);
            active = p.force_reset ? masked : true;
        }
        const bool live = mine && active, stepping = live && !is_reset;
        const unsigned live_b = __ballot_sync(kFull, live);
        const int stepc = cur.step;
        uint32_t act = cur.act & 7u;
        uint16_t me = cur.cell;
        int sel = p.random_dest ? (int)(cur.dsel & 15u) : fixed_sel;
        if (p.order) s_ord[eshift + (cur.ord & (N - 1))] = (uint8_t)a;  // inverse permutation: rank of agent ord[a] is a
        if (!p.actions || !p.order) {  // one Philox call per lane: lanes 0..NA-1 the action words, the next NO the order draws
            const bool is_act = a < NA;
            const uint4 w = philox4x32_10(p.seed, gid, p.tick, is_act ? kStreamAction : kStreamOrder,
                                          (uint32_t)(is_act ? a : a - NA), 0);
            if (a < NC) reinterpret_cast<uint4 *>(s_rng)[e * NC + a] = w;
        }
        __syncwarp();
        // -------------------------------------------------------------- moves (abstract_world.py:39-78)
        // rank of this agent in the move order: Fisher-Yates over the identity (abstract_trainer.py:66) followed per
        // element -- swapping positions k and j moves the element at k to j and the one at j to k
        int rank = a;
        if (p.order) rank = s_ord[lane];
        else {
            const uint32_t *draws = s_rng + (e * NC + NA) * 4;
#pragma unroll
            for (int k = N - 1; k >= 1; --k) {
                const int j = (int)scale(draws[N - 1 - k], (uint32_t)(k + 1));
                rank = rank == k ? j : (rank == j ? k : rank);
            }
        }
        if (!p.actions) act = s_rng[(e * NC + (a >> 2)) * 4 + (a & 3)] >> 30;  // abstract_agent.py:31-34
        bool hit_agent = false, hit_wall = false;
        {
            // action.py:4-9 in index space (abstract_map.py:49-50 flips y): (dx + 1) | (dy + 1) << 2 per action
            const uint32_t t = (0x55559416u >> (act * 4)) & 15u;
            const int cx = me & 0xFF, cy = me >> 8;
            const int nx = cx + (int)(t & 3u) - 1, ny = cy + (int)(t >> 2) - 1;
            const uint32_t cbit = 1u << (cy + R), nbit = 1u << (ny + R);
            const bool moving = stepping && t != 5u;
            if (stepping) {
                hit_wall = (__ldg(&p.wall[nx]) & nbit) != 0;
                atomicOr(&agrid[cx], cbit);
            }
            __syncwarp();
            // one round per place in the order; a round has one mover per environment, so plain read-modify-writes do
#pragma unroll
            for (int j = 0; j < N; ++j) {
                if (moving && rank == j) {
                    if (agrid[nx] & nbit) hit_agent = true;  // abstract_world.py:56-64 (own cell excluded: only STAY hits it)
                    else if (!hit_wall) {
                        agrid[cx] &= ~cbit;
                        agrid[nx] |= nbit;
                        me = (uint16_t)(nx | (ny << 8));
                    }
                }
                __syncwarp();
            }
            if (stepping) s_pos[lane] = me;
        }
        const unsigned ca_b = __ballot_sync(kFull, hit_agent), cw_b = __ballot_sync(kFull, hit_wall);
        __syncwarp();

        // -------------------------------------------------------------- leader: reset (default_environment.py:64-116)
        uint32_t attempt = 0;
        uint4 rw = make_uint4(0, 0, 0, 0);
        int c_gen = 0, c_done = 0;  // objects generated / delivered by this call
        auto spawn = [&](uint32_t stream) -> uint32_t {  // default_environment.py:134-148; returns the cell or 0xFFFF
            for (;;) {
                if (attempt >= (uint32_t)kRespawnCap) {
                    atomicOr(p.error_flag, 1);
                    return 0xFFFFu;
                }
                if ((attempt & 1u) == 0) rw = philox4x32_10(p.seed, gid, p.tick, stream, attempt >> 1, 0);
                const uint32_t u1 = (attempt & 1u) ? rw.z : rw.x, u2 = (attempt & 1u) ? rw.w : rw.y;
                ++attempt;
                const int x = 1 + (int)scale(u1, (uint32_t)(X - 1)), y = 1 + (int)scale(u2, (uint32_t)(Y - 1));
                const uint32_t bit = 1u << (y + R);
                if (!(__ldg(&p.wall[x]) & bit) && (__ldg(&p.area[x]) & bit) && !(grid[x] & bit) && !(agrid[x] & bit)) {
                    grid[x] |= bit;
                    ++c_gen;
                    return (uint32_t)(x | (y << 8));
                }
            }
        };
        if (live_b & ~__ballot_sync(kFull, stepping)) {  // some environment of this item resets (rare)
            if (lead && live && is_reset) {
                for (int x = 0; x < X; ++x) grid[x] = 0;
                if (p.random_dest) {  // three_rooms_map.py:42-55
                    uint64_t perm = 0;
                    for (int k = 0; k < p.M; ++k) perm |= (uint64_t)k << (4 * k);
                    uint32_t d = 0;
                    perm = fisher_yates(perm, p.M, p.seed, gid, p.tick, kStreamDest, d);
                    for (int k = 0; k < N; ++k) {
                        const uint4 w = philox4x32_10(p.seed, gid, p.tick, kStreamDest, d >> 2, 0);
                        const uint32_t c = scale(pick(w, d & 3), (uint32_t)p.M);
                        ++d;
                        const uint8_t s = (uint8_t)nib(perm, (int)c);
                        s_dsel[eshift + k] = s;
                        p.dest_sel[(size_t)env * N + k] = s;
                    }
                }
                uint64_t perm = 0;  // default_environment.py:33,97-98 (persists across episodes)
                const int NI = p.n_init;
                for (int k = 0; k < NI; ++k) perm |= (uint64_t)p.init_perm[(size_t)env * NI + k] << (4 * k);
                if (p.shuffle_init) {
                    uint32_t d = 0;
                    perm = fisher_yates(perm, NI, p.seed, gid, p.tick, kStreamInitPerm, d);
                    for (int k = 0; k < NI; ++k) p.init_perm[(size_t)env * NI + k] = (uint8_t)nib(perm, k);
                }
                for (int k = 0; k < N; ++k) {
                    const uint16_t c = p.init_xy[nib(perm, k)];
                    s_pos[eshift + k] = c;
                    agrid[c & 0xFF] |= 1u << ((c >> 8) + R);
                }
                for (int n = 0; n < p.num_objects; ++n) spawn(kStreamObjects);
                attempt = 0;
            }
            __syncwarp();
            if (live && is_reset) {
                me = s_pos[lane];
                if (p.random_dest) sel = s_dsel[lane];
            }
        }

        // -------------------------------------------------------------- one lane per agent
        const int px = me & 0xFF, py = me >> 8;
        // pickup test on the agent's own cell (default_environment.py:196-231)
        const uint32_t mybit = 1u << (py + R);
        const bool want = stepping && task0 && (grid[px] & mybit) && (__ldg(&p.dest[sel * X + px]) & mybit);
        const unsigned want_b = __ballot_sync(kFull, want);
        int nobj = 0;
        if (stepping)
            for (int x = a; x < X; x += N) nobj += __popc(grid[x]);
#pragma unroll
        for (int s = 1; s < N; s <<= 1) nobj += __shfl_xor_sync(kFull, nobj, s);
        uint32_t pm = (want_b >> eshift) & MASKN, sp = 0;
        if (want_b) {  // rare: the leader replays this environment's pickups in agent order
            if (lead && pm) {
                uint32_t took = 0;
                for (uint32_t m = pm; m; m &= m - 1) {
                    const int j = __ffs(m) - 1;
                    const uint16_t c = s_pos[eshift + j];
                    const uint32_t bit = 1u << ((c >> 8) + R);
                    if (!(grid[c & 0xFF] & bit)) continue;  // two agents forced onto one cell (maexp_env_set_state): the first took it
                    grid[c & 0xFF] &= ~bit;
                    took |= 1u << j;
                    ++c_done;
                    uint32_t cell = 0xFFFFu;
                    if (p.keep) cell = spawn(kStreamRespawn);
                    s_ev[eshift + j] = (uint16_t)cell;
                    if (cell != 0xFFFFu) sp |= 1u << j;
                }
                s_flags[e] = took | (sp << 16);
            }
            __syncwarp();
            if (pm) {
                const uint32_t ev = s_flags[e];
                pm = ev & 0xFFFFu;
                sp = ev >> 16;
            }
        }
        const uint32_t upto = (2u << a) - 1u;  // agents 0..a
        const bool rew_pos = (pm >> a) & 1u;
        const bool rew_neg = has_task && (hit_agent || hit_wall);
        const int nobj_a = nobj - __popc(pm & upto) + __popc(sp & upto);
        const bool done = stepping && (nobj_a == 0 || stepc == p.L - 1);  // default_environment.py:233-237
        const unsigned pos_b = __ballot_sync(kFull, rew_pos), neg_b = __ballot_sync(kFull, rew_neg),
                       done_b = __ballot_sync(kFull, done);

        if (live) {
            // ------------------------------------------------ observation_handler.py:67-105, bits in output order
            const uint64_t los = __ldg(&p.los[px * Y + py]);
            auto win_bit = [&](uint16_t cell) -> uint64_t {
                const int wa = (int)(cell & 0xFF) - px + R, wb = (int)(cell >> 8) - py + R;
                return (wa >= 0 && wa < V && wb >= 0 && wb < V) ? (1ull << (wa * V + wb)) : 0ull;
            };
            uint64_t A = crop_rows<V>(agrid + (px - R), py), O = crop_rows<V>(grid + (px - R), py);
            for (uint32_t m = pm & ~upto; m; m &= m - 1) {  // events of the agents behind this one have not happened yet
                const int j = __ffs(m) - 1;
                O |= win_bit(s_pos[eshift + j]);
                if ((sp >> j) & 1u) O &= ~win_bit(s_ev[eshift + j]);
            }
            if (p.agent_view == MAEXP_AV_DEFAULT) A &= los;
            if (p.object_view == MAEXP_OV_DEFAULT) O &= los;
            uint32_t *wv = s_w + a * EPW + e;
            wv[0] = stream_piece<0, 0>(A) | stream_piece<VV, 0>(O) | stream_piece<2 * VV, 0>(los);
            wv[32] = stream_piece<0, 1>(A) | stream_piece<VV, 1>(O) | stream_piece<2 * VV, 1>(los);
            wv[64] = stream_piece<0, 2>(A) | stream_piece<VV, 2>(O) | stream_piece<2 * VV, 2>(los);
            if constexpr (NW > 3) {
                wv[96] = stream_piece<0, 3>(A) | stream_piece<VV, 3>(O) | stream_piece<2 * VV, 3>(los);
                wv[128] = stream_piece<0, 4>(A) | stream_piece<VV, 4>(O) | stream_piece<2 * VV, 4>(los);
            }

            // ------------------------------------------------ per-agent outputs, [E, N] arrays: one element per lane
            const size_t o = (size_t)env * N + a;
            p.pos[o] = me;
            if (p.out.rewards || p.out.rewards_i8) {
                float r = rew_neg ? -1.0f : (rew_pos ? 1.0f : 0.0f);
                if (p.reward_mode != MAEXP_REWARD_DEFAULT) {  // shared_reward_environment.py:15-28
                    const uint32_t rp = (pos_b >> eshift) & MASKN, rn = (neg_b >> eshift) & MASKN;
                    const int sum = __popc(rp & ~rn) - __popc(rn);
                    const bool mean = p.reward_mode == MAEXP_REWARD_SHARED_MEAN || p.shared_max_is_mean;
                    if (mean) r = (float)((double)sum / (double)N);
                    else r = (rp & ~rn) ? 1.0f : ((rn == MASKN) ? -1.0f : 0.0f);
                }
                if (is_reset) r = 0.0f;
                if (p.out.rewards) p.out.rewards[o] = r;
                if (p.out.rewards_i8) p.out.rewards_i8[o] = is_reset ? (int8_t)0 : (int8_t)(rew_neg ? -1 : (rew_pos ? 1 : 0));
            }
            if (p.out.dones) p.out.dones[o] = done ? 1 : 0;
            if (p.out.positions) reinterpret_cast<uint16_t *>(p.out.positions)[o] = me;  // int8 (x, y) = the u16 cell
            if (p.out.actions_taken) p.out.actions_taken[o] = is_reset ? (int8_t)-1 : (int8_t)act;
            if (p.out.destination) p.out.destination[o] = (uint8_t)sel;
            if (lead) {
                const bool done0 = (done_b >> eshift) & 1u;
                if (p.out.was_reset) p.out.was_reset[env] = is_reset ? 1 : 0;
                if (p.out.done_env) p.out.done_env[env] = done0 ? 1 : 0;
                p.step[env] = is_reset ? 0 : stepc + 1;  // default_environment.py:172
                p.pending[env] = (p.auto_reset && !is_reset && done0) ? 1 : 0;  // default_trainer.py:85
                // generated, completed, agents_collided, walls_collided (default_environment.py:65-69,145,207,217,226)
                int32_t *cnt = p.counters + (size_t)env * 4;
                if (is_reset) *reinterpret_cast<int4 *>(cnt) = make_int4(c_gen, 0, 0, 0);
                else {  // increments without reading the old values back
                    const int c_ca = __popc((ca_b >> eshift) & task_mask), c_cw = __popc((cw_b >> eshift) & task_mask);
                    if (c_gen) atomicAdd(cnt + 0, c_gen);
                    if (c_done) atomicAdd(cnt + 1, c_done);
                    if (c_ca) atomicAdd(cnt + 2, c_ca);
                    if (c_cw) atomicAdd(cnt + 3, c_cw);
                }
            }
        }
        n_stepped += __popc(__ballot_sync(kFull, lead && stepping));
        n_resets += __popc(__ballot_sync(kFull, lead && live && is_reset));
        __syncwarp();
        // the object rows go back (coalesced) and the next item's state is requested before the stores start
        if (rows_ahead) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int e2 = (int)((row_env >> (4 * k)) & 15u);
                if (e2 != 15 && lane + 32 * k < nenv * X && ((live_b >> (e2 * N)) & 1u))
                    p.obj[(size_t)env0 * X + lane + 32 * k] = s_grid[lane + 32 * k + R + e2 * row_skip] >> R;
            }
        } else {
            for (int i = lane; i < nenv * X; i += 32) {
                const int e2 = (int)__umulhi((uint32_t)i, p.tx_magic);
                if ((live_b >> (e2 * N)) & 1u) p.obj[(size_t)env0 * X + i] = s_grid[i + R + e2 * row_skip] >> R;
            }
        }
        const int following = __shfl_sync(kFull, next_group, 0);
        if (following < num_groups) cur = load_item(following);

        // -------------------------------------------------------------- expand + stream out (whole warp)
        if (p.out.obs) {
            constexpr int QPS = SS::QPS, NFULL = QPS >> 5, REM = QPS & 31;
            constexpr int PITCH = REM <= 1 ? 1 : (REM <= 2 ? 2 : (REM <= 4 ? 4 : (REM <= 8 ? 8 : (REM <= 16 ? 16 : 32))));
            constexpr int BLOCKS = 32 / PITCH;  // env blocks the lanes split into for the trailing quads
            const bool whole = live_b == kFull;
#ifdef MAEXP_LANES_LINEAR  // experiment: env-major slot order [E][N][SLOT] (timing only, the readers expect agent-major)
            float *const dst0 = p.out.obs + (size_t)env0 * N * SLOT;
            const size_t agent_stride = SLOT;
#define MAEXP_ENV_PITCH (N * SLOT)
#else
            float *const dst0 = p.out.obs + (size_t)env0 * SLOT;
            const size_t agent_stride = (size_t)p.E * SLOT;
#define MAEXP_ENV_PITCH SLOT
#endif
            if constexpr (NFULL > 0) {
#pragma unroll
                for (int c = 0; c < NFULL; ++c) {
                    const int q = c * 32 + lane;
                    const QuadConst k = quad_consts(q, Q_STR, Q_TAIL, SUB_STR, SUB_TAIL);
                    const uint32_t *wq = s_w + (q >> 3) * 32;
#pragma unroll kUnrollA
                    for (int a2 = 0; a2 < N; ++a2) {
                        float *dp = dst0 + a2 * agent_stride + q * 4;
#pragma unroll
                        for (int e2 = 0; e2 < EPW; e2 += 4) {
                            const uint4 w = *reinterpret_cast<const uint4 *>(wq + a2 * EPW + e2);
                            if (whole || ((live_b >> ((e2 + 0) * N)) & 1u)) store_quad(reinterpret_cast<float4 *>(dp + (e2 + 0) * MAEXP_ENV_PITCH), expand_quad(k, w.x));
                            if (whole || ((live_b >> ((e2 + 1) * N)) & 1u)) store_quad(reinterpret_cast<float4 *>(dp + (e2 + 1) * MAEXP_ENV_PITCH), expand_quad(k, w.y));
                            if (whole || ((live_b >> ((e2 + 2) * N)) & 1u)) store_quad(reinterpret_cast<float4 *>(dp + (e2 + 2) * MAEXP_ENV_PITCH), expand_quad(k, w.z));
                            if (whole || ((live_b >> ((e2 + 3) * N)) & 1u)) store_quad(reinterpret_cast<float4 *>(dp + (e2 + 3) * MAEXP_ENV_PITCH), expand_quad(k, w.w));
                        }
                    }
                }
            }
            if constexpr (REM > 0) {  // trailing quads: lanes split as (env block, quad) with a power-of-two quad pitch
                const int r = lane & (PITCH - 1), blk = lane / PITCH;
                if (r < REM) {
                    const int q = NFULL * 32 + r;
                    const QuadConst k = quad_consts(q, Q_STR, Q_TAIL, SUB_STR, SUB_TAIL);
                    const uint32_t *wq = s_w + (q >> 3) * 32;
                    if constexpr (EPW / BLOCKS >= 4) {  // a lane visits EPW / BLOCKS consecutive envs of every agent
                        constexpr int PER = EPW / BLOCKS;
                        const int eb = blk * PER;
#pragma unroll
                        for (int a2 = 0; a2 < N; ++a2) {
                            float *dp = dst0 + a2 * agent_stride + (size_t)eb * MAEXP_ENV_PITCH + q * 4;
#pragma unroll
                            for (int e2 = 0; e2 < PER; e2 += 4) {
                                const uint4 w = *reinterpret_cast<const uint4 *>(wq + a2 * EPW + eb + e2);
                                if (whole || ((live_b >> ((eb + e2 + 0) * N)) & 1u)) store_quad(reinterpret_cast<float4 *>(dp + (e2 + 0) * MAEXP_ENV_PITCH), expand_quad(k, w.x));
                                if (whole || ((live_b >> ((eb + e2 + 1) * N)) & 1u)) store_quad(reinterpret_cast<float4 *>(dp + (e2 + 1) * MAEXP_ENV_PITCH), expand_quad(k, w.y));
                                if (whole || ((live_b >> ((eb + e2 + 2) * N)) & 1u)) store_quad(reinterpret_cast<float4 *>(dp + (e2 + 2) * MAEXP_ENV_PITCH), expand_quad(k, w.z));
                                if (whole || ((live_b >> ((eb + e2 + 3) * N)) & 1u)) store_quad(reinterpret_cast<float4 *>(dp + (e2 + 3) * MAEXP_ENV_PITCH), expand_quad(k, w.w));
                            }
                        }
                    } else {  // (agent, env) pairs dealt to the blocks: pair index = a2 * EPW + e2, exactly the word index
#pragma unroll kUnrollA
                        for (int pr = 0; pr < 32; pr += BLOCKS) {
                            const int pair = pr + blk, a2 = pair / EPW, e2 = pair % EPW;
                            if (whole || ((live_b >> (e2 * N)) & 1u))
                                store_quad(reinterpret_cast<float4 *>(dst0 + a2 * agent_stride + (size_t)e2 * MAEXP_ENV_PITCH + q * 4),
                                           expand_quad(k, wq[pair]));
                        }
                    }
                }
            }
        }
        group = following;
        __syncwarp();
    }
    // the last warp out re-arms the queue for the next launch (works under CUDA-graph replay: no host-side reset).  Every
    // warp has received the value of its last fetch before it gets here, so no fetch can land after the re-arm.
    if (lane == 0) {
        if (n_stepped) atomicAdd(p.stats + 0, (unsigned long long)n_stepped);
        if (n_resets) atomicAdd(p.stats + 1, (unsigned long long)n_resets);
    }
    if (p.queue && lane == 0) {
        const unsigned done = atomicAdd(p.queue + 1, 1u);
        if (done == (unsigned)nwarps - 1u) {
            p.queue[0] = 0u;
            p.queue[1] = 0u;
        }
    }
}

KernelFn pick_lanes(int V, int N) {
    if (V == 7) {
        if (N == 8) return env_step_lanes_kernel<7, 8>;
        if (N == 4) return env_step_lanes_kernel<7, 4>;
        if (N == 2) return env_step_lanes_kernel<7, 2>;
    }
    if (V == 5) {
        if (N == 8) return env_step_lanes_kernel<5, 8>;
        if (N == 4) return env_step_lanes_kernel<5, 4>;
        if (N == 2) return env_step_lanes_kernel<5, 2>;
    }
    return nullptr;
}

}  // namespace maexp
