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
#include <stdlib.h>
#include <string.h>

#include <string>
#include <vector>

#include "../../include/maexp_b200.h"
#include "env_kernel.cuh"

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
                    destination[(((size_t)env * N + k) * X + x) * Y + y] = (int8_t)((p.dest[sel * X + x] >> (y + p.R)) & 1u);
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
    for (int k = 0; k < p.n_init; ++k) p.init_perm[(size_t)env * p.n_init + k] = (uint8_t)k;  // default_environment.py:33
    for (int k = 0; k < p.N; ++k) {
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
    int grid_ctas = 0;
    KernelFn fn = nullptr;
    // agent-parallel kernel (env_lanes.cu) for the shapes it takes; chosen per call when the outputs asked for allow it
    KernelFn fn_lanes = nullptr;
    size_t lanes_smem = 0;
    int lanes_ctas = 0;
    bool lanes_queue = false;  // MAEXP_LANES_QUEUE=1: work items from an atomic queue instead of the static round-robin
    std::vector<void *> allocs;
    // pinned-free staging for the host-buffer entry point
    int8_t *d_actions = nullptr;
    float *d_rewards = nullptr;
    uint8_t *d_dones = nullptr, *d_was_reset = nullptr;
    // pipelined host entry point: kHostDepth steps in flight, copies on their own streams
    struct HostSlot {
        int8_t *actions = nullptr, *rewards_i8 = nullptr;
        float *rewards = nullptr;
        uint8_t *dones = nullptr, *was_reset = nullptr, *done_env = nullptr;
        cudaEvent_t in = nullptr, stepped = nullptr, out = nullptr;
    };
    static constexpr int kHostDepth = 3;
    HostSlot slots[kHostDepth];
    cudaStream_t h2d = nullptr, d2h = nullptr;
    uint64_t issued = 0;
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
    (void)V;
    size_t words = 256;  // nibble -> float4 table; static map tables stay in global memory (read-only L1 path)
    const size_t per_warp = ((size_t)epw * p.gs + (size_t)epw * p.as + (size_t)epw * p.ws +
                             ((size_t)p.N * epw + (size_t)p.P * p.N * epw + 1) / 2 + (epw + 3) / 4 + 3) & ~(size_t)3;
    words += per_warp * kWarpsPerCta;
    return words * 4;
}

KernelFn pick_kernel(int V, int epw, int noise, int nspec) {
    switch (V) {
    case 5: return pick_v5(epw, noise, nspec);
    case 7: return pick_v7(epw, noise, nspec);
    case 9: return pick_v9(epw, noise, nspec);
    }
    return nullptr;
}

inline bool misaligned_obs(const void *ptr) { return ptr && (reinterpret_cast<uintptr_t>(ptr) & 15); }

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
    if ((p.N & 7) == 0) {  // the kernel writes eight agents per vector store
        auto misaligned = [](const void *ptr, size_t a) { return ptr && (reinterpret_cast<uintptr_t>(ptr) & (a - 1)); };
        if (misaligned(p.out.rewards, 16) || misaligned(p.out.dones, 8) || misaligned(p.out.positions, 16) ||
            misaligned(p.out.actions_taken, 8) || misaligned(p.out.rewards_i8, 8))
            return fail("maexp_env_step: rewards / positions must be 16-byte aligned, dones / actions_taken 8-byte aligned");
    }
    if (misaligned_obs(p.out.obs) || misaligned_obs(p.out.obs_rel)) return fail("maexp_env_step: obs slots must be 16-byte aligned");
    if (!h->lanes_queue) p.queue = nullptr;
    const bool lanes = h->fn_lanes && !p.out.obs_rel && !p.out.coordinates && !p.out.objects && !p.out.objects_bits &&
                       !p.out.heatmaps && !(reinterpret_cast<uintptr_t>(p.out.positions) & 1);
    if (lanes) h->fn_lanes<<<h->lanes_ctas, kWarpsPerCta * 32, h->lanes_smem, stream>>>(p);
    else h->fn<<<h->grid_ctas, kWarpsPerCta * 32, h->smem_bytes, stream>>>(p);
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
    if (T < 1 || T > MAEXP_MAX_TYPES) return fail("maexp_env_create: type_objects must be in 1..8");
    if (cfg->env_kind < MAEXP_ENV_DEFAULT || cfg->env_kind > MAEXP_ENV_JUNCTIONAL) return fail("maexp_env_create: unknown env_kind");
    if (cfg->env_kind == MAEXP_ENV_JUNCTIONAL && T < 5)
        return fail("maexp_env_create: the junctional environment places object types 2, 3 and 4 (needs type_objects >= 5)");
    if (X < 3 || X > 255 || Y < 3 || Y + cfg->visible_range / 2 > 32)
        return fail("maexp_env_create: map must satisfy 3<=X<=255, 3<=Y and Y + visible_range/2 <= 32 (32-bit grid rows)");
    if (V != 5 && V != 7 && V != 9) return fail("maexp_env_create: visible_range must be 5, 7 or 9");
    if (cfg->num_init_xys > MAEXP_MAX_AGENTS) return fail("maexp_env_create: at most 16 init_xys rows");
    if (cfg->dynamic_obstacles && (X < 5 || Y < 5)) return fail("maexp_env_create: dynamic_obstacles needs a map of at least 5 x 5");
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
    p.R = V / 2;
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
    int slot_align = 8;  // floats: agent slots start on 32-byte DRAM sectors (measured: 104 -> 90 us per step vs 16-byte alignment)
    if (const char *a = getenv("MAEXP_SLOT_ALIGN")) { const int v = atoi(a); if (v == 4 || v == 8 || v == 16 || v == 32) slot_align = v; }
    p.slot_floats = (p.K + slot_align - 1) / slot_align * slot_align;
    p.qps = p.slot_floats / 4;
    p.qps_magic = (uint32_t)((0x100000000ull + p.qps - 1) / p.qps);
    p.wpa = (p.K + 31) / 32;
    p.ws = p.wpa + 3;  // words per env in the (word-major) observation bit buffer, incl. the zero slack BitStream::finish writes
    p.gs = (T * (X + 2 * (V / 2))) | 1;
    p.tx_magic = (uint32_t)((0x100000000ull + (uint64_t)(T * X) - 1) / (uint64_t)(T * X));
    p.as = (X + 2 * (V / 2)) | 1;
    p.fast3 = (V <= 7) && (cfg->agent_view == MAEXP_AV_DEFAULT || cfg->agent_view == MAEXP_AV_SIMPLE) &&
              (cfg->object_view == MAEXP_OV_DEFAULT || cfg->object_view == MAEXP_OV_SIMPLE);
    p.area_bit0 = (p.C - 1) * V * V;
    p.q_str = p.area_bit0 >> 2;
    p.q_tail = (p.K - 1) >> 2;
    p.sub_str = 0xF & ~((1 << (p.area_bit0 & 3)) - 1);
    p.sub_tail = (1 << (p.K - 4 * p.q_tail)) - 1;
    p.agent_view = cfg->agent_view; p.object_view = cfg->object_view;
    p.merged = cfg->area_mask == MAEXP_MASK_MERGED;
    p.noise = cfg->noise; p.reward_mode = cfg->reward_mode; p.shared_max_is_mean = cfg->shared_max_is_mean;
    p.L = cfg->max_episode_length; p.num_objects = cfg->num_objects; p.keep = cfg->keep_objects_num;
    p.shuffle_init = cfg->shuffle_init_xys; p.auto_reset = cfg->auto_reset; p.random_dest = cfg->random_destinations;
    p.seed = cfg->seed; p.gid0 = cfg->first_global_env_id;
    p.env_kind = cfg->env_kind;
    p.n_init = cfg->num_init_xys > N ? cfg->num_init_xys : N;
    for (int i = 0; i < p.n_init; ++i) p.init_xy[i] = (uint16_t)(cfg->init_xys[i][0] | (cfg->init_xys[i][1] << 8));
    for (int i = 0; i < N; ++i) {
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
                    if (src[((size_t)pl * X + x) * Y + y]) r[(size_t)pl * X + x] |= 1u << (y + V / 2);  // pre-shifted by R
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
    std::vector<uint32_t> wall_rows = rows(tb->wall, 1);
    if (cfg->dynamic_obstacles) {  // the kernel adds the two moving cells itself; the table keeps the fixed walls
        wall_rows[X / 2] &= ~(1u << (1 + V / 2));
        wall_rows[X / 2] &= ~(1u << (Y / 2 + V / 2));
        rc |= dev_alloc(h, &p.wsteps, (size_t)E);
    }
    rc |= upload(h, &p.wall, wall_rows);
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
    rc |= dev_alloc(h, &p.init_perm, (size_t)E * p.n_init);
    rc |= dev_alloc(h, &p.dest_sel, (size_t)E * N);
    rc |= dev_alloc(h, &p.past, (size_t)E * (p.P ? p.P : 1) * N);
    rc |= dev_alloc(h, &p.stats, 2);
    rc |= dev_alloc(h, &p.error_flag, 1);
    rc |= dev_alloc(h, &p.queue, 2);
    rc |= dev_alloc(h, &h->d_actions, (size_t)E * N);
    rc |= dev_alloc(h, &h->d_rewards, (size_t)E * N);
    rc |= dev_alloc(h, &h->d_dones, (size_t)E * N);
    rc |= dev_alloc(h, &h->d_was_reset, (size_t)E);
    if (rc) { maexp_env_destroy(h); return -1; }

    // envs per warp: keep >= ~24 warps per SM in flight when E allows it
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    // envs per warp (one lane per env in the sequential phase).  Measured on B200 at E = 65536:
    // 8 -> 119 us, 16 -> 89 us, 32 -> 93 us per step; small batches prefer more, shorter groups.
    // With observation noise the expansion phase dominates (one Philox call per quad), and it runs at full lane
    // efficiency only when a warp's group has 32 envs: flip-noise cfg2 256 -> 223 us per step.
    p.stagger_ns = 0;
    p.stagger_div = sms;
    if (const char *sg = getenv("MAEXP_STAGGER_NS")) p.stagger_ns = atoi(sg);
    h->epw = (E >= 16 * sms * 8) ? (p.noise != MAEXP_NOISE_NONE && E >= 32 * sms * 4 ? 32 : 16) : 8;
    if (const char *force = getenv("MAEXP_EPW")) {
        const int v = atoi(force);
        if (v == 8 || v == 16 || v == 32) h->epw = v;
    }
    h->smem_bytes = env_smem_bytes(p, V, h->epw);
    while (h->smem_bytes > 200 * 1024 && h->epw > 8) {
        h->epw >>= 1;
        h->smem_bytes = env_smem_bytes(p, V, h->epw);
    }
    if (h->smem_bytes > 227 * 1024) { maexp_env_destroy(h); return fail("maexp_env_create: configuration needs more than 227 KB shared memory per CTA"); }
    for (int i = 0; i < 32 * T * X; ++i)
        if ((int)(((uint64_t)i * p.tx_magic) >> 32) != i / (T * X)) { maexp_env_destroy(h); return fail("maexp_env_create: internal magic division (TX)"); }
    int nspec = ((N == 8 || N == 4 || N == 2) && T == 1 && !transition) ? N : 0;
    if (getenv("MAEXP_NOSPEC")) nspec = 0;
    h->fn = pick_kernel(V, h->epw, p.noise, nspec);
    if (!h->fn) { maexp_env_destroy(h); return fail("maexp_env_create: no kernel for this visible_range / noise"); }
    if (cudaFuncSetAttribute(h->fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem_bytes) != cudaSuccess) {
        maexp_env_destroy(h);
        return fail("maexp_env_create: cannot reserve shared memory");
    }
    int per_sm = 1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, h->fn, kWarpsPerCta * 32, h->smem_bytes);
    if (per_sm < 1) per_sm = 1;
    const int groups = (E + h->epw - 1) / h->epw;
    h->grid_ctas = per_sm * sms;  // persistent: exactly one resident wave
    if (h->grid_ctas > (groups + kWarpsPerCta - 1) / kWarpsPerCta) h->grid_ctas = (groups + kWarpsPerCta - 1) / kWarpsPerCta;
    // one lane per (env, agent) where the shape allows it (env_lanes.cu); MAEXP_LANES=0 keeps the lane-per-env kernel
    // Default: N = 8 (62.4 against 67.3 us per step at 65 536 envs; with two agents per environment the lane-per-env kernel
    // already fills half a warp and measures the same, cfg1 19.7 against 20.4 us).  MAEXP_LANES=1 / 0 forces the choice.
    const char *lanes_env = getenv("MAEXP_LANES");
    const bool lanes_wanted = lanes_env ? atoi(lanes_env) != 0 : N == 8;
    if (nspec && p.fast3 && p.noise == MAEXP_NOISE_NONE && !p.merged && !cfg->dynamic_obstacles &&
        p.slot_floats == (p.K + 7) / 8 * 8 && lanes_wanted) {
        h->fn_lanes = pick_lanes(V, N);
        if (h->fn_lanes) {
            if (const char *q = getenv("MAEXP_LANES_QUEUE")) h->lanes_queue = atoi(q) != 0;
            h->lanes_smem = lanes_warp_words(N, V, p.gs, p.as) * 4 * kWarpsPerCta;
            int lanes_per_sm = 0;
            if (cudaFuncSetAttribute(h->fn_lanes, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->lanes_smem) != cudaSuccess ||
                cudaOccupancyMaxActiveBlocksPerMultiprocessor(&lanes_per_sm, h->fn_lanes, kWarpsPerCta * 32, h->lanes_smem) != cudaSuccess ||
                lanes_per_sm < 1) {
                cudaGetLastError();
                h->fn_lanes = nullptr;
            } else {
                const int items = (E + 32 / N - 1) / (32 / N);
                if (N == 8 && lanes_per_sm > 4) lanes_per_sm = 4;  // see MAEXP_LANES_MINB in env_lanes.cu
                if (const char *c = getenv("MAEXP_LANES_CTAS")) { const int v = atoi(c); if (v >= 1 && v <= 8) lanes_per_sm = v; }
                h->lanes_ctas = lanes_per_sm * sms;
                if (h->lanes_ctas > (items + kWarpsPerCta - 1) / kWarpsPerCta) h->lanes_ctas = (items + kWarpsPerCta - 1) / kWarpsPerCta;
            }
        }
    }
    init_state_kernel<<<(E + 127) / 128, 128>>>(p);
    if (cudaDeviceSynchronize() != cudaSuccess) { maexp_env_destroy(h); return fail("maexp_env_create: init kernel failed"); }
    *out = h;
    return 0;
}

void maexp_env_destroy(maexp_env *h) {
    if (!h) return;
    if (h->h2d) {
        cudaStreamSynchronize(h->h2d);
        cudaStreamSynchronize(h->d2h);
        for (auto &sl : h->slots) {
            if (sl.in) cudaEventDestroy(sl.in);
            if (sl.stepped) cudaEventDestroy(sl.stepped);
            if (sl.out) cudaEventDestroy(sl.out);
        }
        cudaStreamDestroy(h->h2d);
        cudaStreamDestroy(h->d2h);
    }
    for (void *d : h->allocs) cudaFree(d);
    delete h;
}

int maexp_env_channels(const maexp_env *h) { return h ? h->p.C : -1; }
int maexp_env_agent_parallel(const maexp_env *h) { return (h && h->fn_lanes) ? 1 : 0; }
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

int64_t maexp_env_heatmap_ints(const maexp_env *h) {
    if (!h) return -1;
    return (int64_t)(2 * h->p.N + 2 * h->p.T + 2) * h->p.X * h->p.Y;
}

}  // extern "C"

namespace {
// compact: rewards_host is int8 [E,N], dones_host is uint8 [E]
int step_host_pipelined(maexp_env *h, uint32_t tick, const int8_t *actions_host, float *obs_ring_slot, void *rewards_host,
                        uint8_t *dones_host, uint8_t *was_reset_host, void *stream_, uint64_t *ticket, bool compact) {
    if (!h) return fail("maexp_env_step_host_async: null handle");
    if (compact && h->p.reward_mode != MAEXP_REWARD_DEFAULT)
        return fail("maexp_env_step_host_compact_async: shared rewards are fractions, use maexp_env_step_host_async");
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    const size_t E = (size_t)h->p.E, EN = E * h->p.N;
    if (!h->h2d) {
        MAEXP_CUDA(cudaStreamCreateWithFlags(&h->h2d, cudaStreamNonBlocking));
        MAEXP_CUDA(cudaStreamCreateWithFlags(&h->d2h, cudaStreamNonBlocking));
        for (auto &sl : h->slots) {
            int rc = dev_alloc(h, &sl.actions, EN) | dev_alloc(h, &sl.rewards, EN) | dev_alloc(h, &sl.dones, EN) |
                     dev_alloc(h, &sl.was_reset, E) | dev_alloc(h, &sl.rewards_i8, EN) | dev_alloc(h, &sl.done_env, E);
            if (rc) return -1;
            MAEXP_CUDA(cudaEventCreateWithFlags(&sl.in, cudaEventDisableTiming));
            MAEXP_CUDA(cudaEventCreateWithFlags(&sl.stepped, cudaEventDisableTiming));
            MAEXP_CUDA(cudaEventCreateWithFlags(&sl.out, cudaEventDisableTiming));
        }
    }
    maexp_env::HostSlot &sl = h->slots[h->issued % maexp_env::kHostDepth];
    // the step issued kHostDepth calls ago used this slot: wait until its results reached the host
    if (h->issued >= (uint64_t)maexp_env::kHostDepth) MAEXP_CUDA(cudaEventSynchronize(sl.out));
    if (actions_host) {
        MAEXP_CUDA(cudaMemcpyAsync(sl.actions, actions_host, EN, cudaMemcpyHostToDevice, h->h2d));
        MAEXP_CUDA(cudaEventRecord(sl.in, h->h2d));
        MAEXP_CUDA(cudaStreamWaitEvent(stream, sl.in, 0));
    }
    maexp_obs_out out;
    memset(&out, 0, sizeof(out));
    out.obs = obs_ring_slot;
    if (compact) { out.rewards_i8 = sl.rewards_i8; out.done_env = sl.done_env; }
    else { out.rewards = sl.rewards; out.dones = sl.dones; }
    out.was_reset = sl.was_reset;
    if (launch(h, tick, 0, nullptr, actions_host ? sl.actions : nullptr, nullptr, nullptr, &out, stream)) return -1;
    MAEXP_CUDA(cudaEventRecord(sl.stepped, stream));
    MAEXP_CUDA(cudaStreamWaitEvent(h->d2h, sl.stepped, 0));
    if (compact) {
        if (rewards_host) MAEXP_CUDA(cudaMemcpyAsync(rewards_host, sl.rewards_i8, EN, cudaMemcpyDeviceToHost, h->d2h));
        if (dones_host) MAEXP_CUDA(cudaMemcpyAsync(dones_host, sl.done_env, E, cudaMemcpyDeviceToHost, h->d2h));
    } else {
        if (rewards_host) MAEXP_CUDA(cudaMemcpyAsync(rewards_host, sl.rewards, EN * sizeof(float), cudaMemcpyDeviceToHost, h->d2h));
        if (dones_host) MAEXP_CUDA(cudaMemcpyAsync(dones_host, sl.dones, EN, cudaMemcpyDeviceToHost, h->d2h));
    }
    if (was_reset_host) MAEXP_CUDA(cudaMemcpyAsync(was_reset_host, sl.was_reset, E, cudaMemcpyDeviceToHost, h->d2h));
    MAEXP_CUDA(cudaEventRecord(sl.out, h->d2h));
    h->issued += 1;
    if (ticket) *ticket = h->issued;
    return 0;
}
}  // namespace

extern "C" {

int maexp_env_step_host_async(maexp_env *h, uint32_t tick, const int8_t *actions_host, float *obs_ring_slot,
                              float *rewards_host, uint8_t *dones_host, uint8_t *was_reset_host, void *stream,
                              uint64_t *ticket) {
    return step_host_pipelined(h, tick, actions_host, obs_ring_slot, rewards_host, dones_host, was_reset_host, stream, ticket,
                               false);
}

int maexp_env_step_host_compact_async(maexp_env *h, uint32_t tick, const int8_t *actions_host, float *obs_ring_slot,
                                      int8_t *rewards_i8_host, uint8_t *done_env_host, uint8_t *was_reset_host, void *stream,
                                      uint64_t *ticket) {
    return step_host_pipelined(h, tick, actions_host, obs_ring_slot, rewards_i8_host, done_env_host, was_reset_host, stream,
                               ticket, true);
}

int maexp_env_host_wait(maexp_env *h, uint64_t ticket) {
    if (!h) return fail("maexp_env_host_wait: null handle");
    if (!h->h2d || h->issued == 0) return 0;
    if (ticket > h->issued) return fail("maexp_env_host_wait: unknown ticket");
    const uint64_t depth = maexp_env::kHostDepth;
    const uint64_t first = ticket ? ticket : (h->issued > depth ? h->issued - depth + 1 : 1);
    const uint64_t last = ticket ? ticket : h->issued;
    for (uint64_t t = first; t <= last; ++t) {
        if (t + depth <= h->issued) continue;  // its slot was reused, so it was already waited for
        MAEXP_CUDA(cudaEventSynchronize(h->slots[(t - 1) % depth].out));
    }
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
