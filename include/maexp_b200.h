/* maexp_b200 — C-ABI of the B200-native MAEXP hot path (libmaexp_b200.so).
 *
 * The reference (Yoshi-0921/MAEXP) is pure Python and has NO FFI of its own
 * (SURVEY.md §2.1, §8b).  This boundary is therefore introduced *underneath* the reference's
 * Python plugin surface; each entry point names the reference interface it replaces
 * (paths relative to the reference root).  INTEGRATION.md shows the ctypes binding a
 * maintainer of the reference would add.
 *
 * Conventions: plain pointers and sizes only (no torch types).  Unless a name ends in
 * `_host`, every data pointer is a DEVICE pointer owned by the caller (normally a torch
 * tensor); the library owns only its opaque handles.  Calls are asynchronous on the given
 * CUDA stream (passed as void* = cudaStream_t).  Return value: 0 ok, <0 error, text via
 * maexp_last_error().  One host thread per handle; one process per GPU.
 */
#ifndef MAEXP_B200_H
#define MAEXP_B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MAEXP_MAX_AGENTS 16
#define MAEXP_MAX_TYPES 8
#define MAEXP_MAX_RINGS 4

/* agent_view_method (core/handlers/observations/agents/__init__.py:22-46) */
enum { MAEXP_AV_DEFAULT = 0, MAEXP_AV_SIMPLE = 1, MAEXP_AV_INDIVIDUAL = 2, MAEXP_AV_TRANSITION = 3, MAEXP_AV_INDIVIDUAL_TRANSITION = 4 };
/* object_view_method (core/handlers/observations/objects/__init__.py:18-35) */
enum { MAEXP_OV_DEFAULT = 0, MAEXP_OV_SIMPLE = 1, MAEXP_OV_TYPES = 2 };
/* observation_area_mask (core/handlers/observations/__init__.py:21-28, masks/__init__.py:26-43) */
enum { MAEXP_MASK_LOCAL = 0, MAEXP_MASK_RELATIVE = 1, MAEXP_MASK_MERGED = 2 };
/* observation_noise (core/handlers/observations/noises/__init__.py:34-69) */
enum { MAEXP_NOISE_NONE = 0, MAEXP_NOISE_FLIP = 1, MAEXP_NOISE_DIST = 2, MAEXP_NOISE_THRESHOLD_DIST = 3, MAEXP_NOISE_FLAT = 4 };
/* environment (core/environments/__init__.py:37-71; shared_reward_environment.py:15-28) */
enum { MAEXP_REWARD_DEFAULT = 0, MAEXP_REWARD_SHARED_MEAN = 1, MAEXP_REWARD_SHARED_MAX = 2 };
/* which reward_ind / generate_objects override runs: DefaultEnvironment (default_environment.py:123-132,196-231),
 * SequentialEnvironment (sequential_environment.py:20-64: picking type t puts type t+1 on the same cell, the last
 * type respawns one type-0 object; only type 0 is generated at reset), JunctionalEnvironment
 * (junctional_environment.py:20-71: type 1 -> 3 for agents 2,3 / -> 2 for agents 4,5, type 2 -> 4) */
enum { MAEXP_ENV_DEFAULT = 0, MAEXP_ENV_SEQUENTIAL = 1, MAEXP_ENV_JUNCTIONAL = 2 };

typedef struct {
    int32_t num_envs;            /* E environments resident on this GPU */
    int32_t num_agents;          /* config.num_agents, N <= 16 */
    int32_t type_objects;        /* config.type_objects, T <= 8 */
    int32_t size_x, size_y;      /* map.SIZE_X / SIZE_Y (Y + V/2 <= 32: grid rows are 32-bit) */
    int32_t visible_range;       /* 5, 7 or 9 */
    int32_t past_step;           /* transition views only */
    int32_t agent_view, object_view, area_mask, noise, reward_mode;
    int32_t max_episode_length;
    int32_t num_objects;
    int32_t keep_objects_num;
    int32_t shuffle_init_xys;
    int32_t shared_max_is_mean;  /* 1: reproduce the reference's "max" class, which takes the mean */
    int32_t auto_reset;          /* 1: an env whose done[0] fired is reset by the NEXT step call */
    int32_t num_destinations;    /* M canonical destination masks in maexp_tables.destination */
    int32_t random_destinations; /* 1: three_rooms_map.py:42-55 (shuffle the M masks, choice per agent) */
    uint32_t seed;
    uint32_t first_global_env_id;/* envs are keyed globally so that shards reproduce a single-GPU run */
    int32_t init_xys[MAEXP_MAX_AGENTS][2];   /* INDEX-space cells (abstract_map.py:44-53 applied by the host) */
    int32_t task_mask[MAEXP_MAX_AGENTS];     /* -1 for agent_tasks "-1" (default_environment.py:201-202) */
    int32_t task_order[MAEXP_MAX_AGENTS][MAEXP_MAX_TYPES]; /* digits of agent_tasks[i] in order, -1 terminated */
    float flip_noise_range[MAEXP_MAX_RINGS]; /* fp32 0.5/(p-1), flip_observation_noise.py:28 */
    float std_distribution[MAEXP_MAX_RINGS];
    float noise_threshold[MAEXP_MAX_RINGS];  /* fp32(std*ratio), dist_observation_noise.py:60,66 */
    int32_t env_kind;            /* MAEXP_ENV_* */
    int32_t dynamic_obstacles;   /* 1: the dynamic_simple map (core/maps/dynamic_simple_map.py:12-63): the wall cells
                                  * (X/2, 1) and (X/2, Y/2) of maexp_tables.wall are obstacles that world.step moves
                                  * one cell per step (abstract_world.py:40) -- the first up and down column X/2
                                  * between y = 1 and Y-2, the second along row Y/2 between x = 1 and X-2 -- and that
                                  * keep their place across resets; moves and object respawns see the moved cells,
                                  * the line-of-sight table stays the one of the initial layout, as in the reference */
    int32_t num_init_xys;        /* len(config.init_xys), num_agents..16 (0: num_agents).  default_environment.py:33,97-98
                                  * shuffles the order of ALL rows at every reset (the order persists from episode to
                                  * episode) and agent i starts on row order[i]; rows past num_agents are spare cells */
} maexp_env_cfg;

/* Static per-map tables in the REFERENCE's layout, HOST pointers; packed to bit rows and
 * uploaded once by maexp_env_create.  Replaces the numpy attributes of AbstractMap
 * (core/maps/abstract_map.py:20-24,78) and generate_local_area_mask's return value
 * (core/handlers/observations/masks/local_area_mask.py:50-97). */
typedef struct {
    const int8_t *wall;          /* [X,Y] */
    const int8_t *objects_area;  /* [T,X,Y] */
    const int8_t *noise_area;    /* [X,Y] */
    const int8_t *destination;   /* [M,X,Y] */
    const int8_t *los;           /* [X,Y,V,V] 1 visible, 0 hidden */
} maexp_tables;

/* Per-call outputs (device pointers, any may be NULL).  obs / obs_rel point at ONE slot of the
 * device-resident replay ring, AGENT-MAJOR: fp32 [N, E, slot_floats], slot_floats = maexp_env_slot_floats()
 * (C*V*V rounded up to a multiple of 8 so that every agent slot starts on a 32-byte sector). */
typedef struct {
    float *obs;            /* state[view] of observation_handler.py:95-105 (noisy one for merged) */
    float *obs_rel;        /* merged only: the "relative" key (merged_observtion_handler.py:41-47) */
    float *rewards;        /* [E,N] fp32 (shared modes: fp32 of the float64 mean) */
    uint8_t *dones;        /* [E,N] default_environment.py:233-237, evaluated per agent */
    uint8_t *was_reset;    /* [E] 1: this call reset the env instead of stepping it */
    int32_t *coordinates;  /* [E,N,12] local_area_mask.py:18-47 key order */
    int8_t *positions;     /* [E,N,2] index-space positions after the call ("agents" key) */
    int8_t *objects;       /* [E,N,T,X,Y] objects_matrix snapshot at agent i's observation time */
    uint8_t *destination;  /* [E,N] selected destination mask per agent ("destination" key = table row) */
    int8_t *actions_taken; /* [E,N] the actions executed by this call (given or Philox-drawn); -1 on reset */
    int8_t *rewards_i8;    /* [E,N] compact rewards for host-driven rollouts: exactly -1 / 0 / +1, so only written for the
                              default reward mode (shared rewards are fractions) */
    uint8_t *done_env;     /* [E] done_n[0] (done_ind is the same for every agent, default_environment.py:233-237) */
    int32_t *heatmaps;     /* optional visit statistics of default_environment.py:37-61,118-148,196-231,172, summed over ALL
                              environments with atomics (the caller zeroes them, as reset() does in the reference):
                              [agents N·X·Y | complete N·X·Y | objects T·X·Y | objects_left T·X·Y | wall_collision X·Y |
                              agents_collision X·Y], maexp_env_heatmap_ints() values in total */
    uint32_t *objects_bits;/* [N,T,X,E] the same objects_matrix snapshot as bit rows (bit y of word x), ENV-MINOR so that the
                              step kernel's stores coalesce: what the replay ring keeps for the `objects_channel` canvas input
                              of the relative / DA6 models (state["objects"], observation_handler.py:99-101) */
} maexp_obs_out;

typedef struct maexp_env maexp_env;

/* generate_map + generate_world + generate_environment + generate_observation_handler
 * (core/maps/__init__.py:25, core/worlds/__init__.py:19, core/environments/__init__.py:37,
 * core/handlers/observations/__init__.py:21) for E batched environments. */
int maexp_env_create(const maexp_env_cfg *cfg, const maexp_tables *tables, maexp_env **out);
void maexp_env_destroy(maexp_env *env);
int maexp_env_channels(const maexp_env *env);     /* ObservationHandler.get_channel, observation_handler.py:55-61 */
int maexp_env_slot_floats(const maexp_env *env);
/* 1 when plain calls (no coordinates / objects / objects_bits / heatmaps / obs_rel output) run the agent-parallel step
 * kernel (csrc/env_lanes.cu: one lane per (environment, agent); N in {2,4,8}, one object type, the three default channels,
 * visible_range 5 / 7, no observation noise), 0 when every call runs the lane-per-environment kernel.  Both produce the
 * same bits.  The library takes it by default for N = 8; the environment variable MAEXP_LANES=1 / 0 at creation time forces
 * it on for every shape it supports / off. */
int maexp_env_agent_parallel(const maexp_env *env);

/* DefaultEnvironment.reset (default_environment.py:64-116) for every env with env_mask[e] != 0
 * (NULL: all).  noise_tape: optional fp32 [E,N,C*V*V] injected draws (NULL: on-chip Philox). */
int maexp_env_reset(maexp_env *env, uint32_t tick, const uint8_t *env_mask, const float *noise_tape,
                    const maexp_obs_out *out, void *stream);

/* DefaultEnvironment.step(action_n, order) (default_environment.py:153-175) →
 * AbstractWorld.step (core/worlds/abstract_world.py:39-78) → reward_ind / done_ind /
 * observation_ind interleaved per agent → ObservationHandler.step.  Shared-reward subclasses:
 * shared_reward_environment.py:15-28.
 * actions int8 [E,N] in {0..4} (NULL: uniform random policy {0..3}, abstract_agent.py:31-34);
 * order uint8 [E,N] permutation (NULL: Philox Fisher-Yates, abstract_trainer.py:66). */
int maexp_env_step(maexp_env *env, uint32_t tick, const int8_t *actions, const uint8_t *order,
                   const float *noise_tape, const maexp_obs_out *out, void *stream);

/* number of int32 values behind maexp_obs_out.heatmaps: (2N + 2T + 2)·X·Y */
int64_t maexp_env_heatmap_ints(const maexp_env *env);

/* Same call with HOST buffers (pinned or pageable): copies actions in, steps, copies rewards /
 * dones / was_reset out and synchronises.  Observations stay in the device ring slot `obs`. */
int maexp_env_step_host(maexp_env *env, uint32_t tick, const int8_t *actions_host, float *obs_ring_slot,
                        float *rewards_host, uint8_t *dones_host, uint8_t *was_reset_host, void *stream);

/* Pipelined form of the same call for host-driven rollouts (the loop of abstract_trainer.py:63-89 with a
 * scripted or random host policy): returns once the step is enqueued.  The action copy, the step kernel and
 * the result copies of consecutive steps overlap on separate streams; up to 3 steps are in flight and a call
 * blocks only until the step issued 3 calls earlier has delivered its results.  Every step still copies its
 * actions in and its rewards / dones / was_reset out; the host buffers of a step (use pinned memory and a
 * distinct set per in-flight step) are valid after maexp_env_host_wait(env, ticket); ticket 0 waits for all. */
int maexp_env_step_host_async(maexp_env *env, uint32_t tick, const int8_t *actions_host, float *obs_ring_slot,
                              float *rewards_host, uint8_t *dones_host, uint8_t *was_reset_host, void *stream,
                              uint64_t *ticket);
int maexp_env_host_wait(maexp_env *env, uint64_t ticket);

/* The same pipelined call with the compact result encoding (6.5 bytes per env instead of 41 for N = 8, which keeps
 * eight ranks inside the host's PCIe bandwidth): rewards int8 [E,N] in {-1,0,+1}, one done flag per env.
 * Only for the default reward mode (returns an error for shared rewards, whose values are fractions). */
int maexp_env_step_host_compact_async(maexp_env *env, uint32_t tick, const int8_t *actions_host, float *obs_ring_slot,
                                      int8_t *rewards_i8_host, uint8_t *done_env_host, uint8_t *was_reset_host,
                                      void *stream, uint64_t *ticket);

/* State taps: map.agents_matrix / objects_matrix / destination_area_matrix and the env counters
 * (default_environment.py:65-69) in the reference's dtypes.  Device pointers, may be NULL. */
int maexp_env_get_state(maexp_env *env, int8_t *pos /*[E,N,2]*/, int8_t *objects /*[E,T,X,Y]*/,
                        int32_t *counters /*[E,4] generated,completed,agents_collided,walls_collided*/,
                        int32_t *steps /*[E]*/, int8_t *destination /*[E,N,X,Y]*/, void *stream);
int maexp_env_set_state(maexp_env *env, const int8_t *pos /*[E,N,2]*/, const int8_t *objects /*[E,T,X,Y]*/,
                        void *stream);
/* number of env-steps (not resets) executed since creation, summed over envs; error flag of the
 * bounded rejection sampler.  Synchronises the stream. */
int maexp_env_stats(maexp_env *env, uint64_t *env_steps, uint64_t *resets, int32_t *error_flag, void *stream);

/* ------------------------------------------------------------------------------------------
 * Device-resident replay ring (replaces ReplayBuffer / RLDataset / DataLoader collate:
 * core/utils/buffer.py:36-63, core/utils/dataset.py:24-29, abstract_trainer.py:91-97).
 * The ring is caller-owned device memory; slot s of every array is what one maexp_env_step /
 * maexp_env_reset call wrote (obs, rewards, dones, was_reset) plus the actions that led to it.
 * The transition (s -> s+1) of env e is valid iff was_reset[s+1][e] == 0. */
typedef struct {
    const float *obs;          /* [S,N,E,slot_floats] (agent-major slots, as maexp_env_step writes them) */
    const int8_t *actions;     /* [S,E,N] action taken to reach slot s */
    const float *rewards;      /* [S,E,N] */
    const uint8_t *dones;      /* [S,E,N] */
    const uint8_t *was_reset;  /* [S,E] */
    const int8_t *positions;   /* [S,E,N,2] index-space agent positions after the call (may be NULL) */
    int32_t num_slots;         /* S */
    int32_t num_envs, num_agents, slot_floats;
    int32_t head;              /* newest written slot */
    int32_t count;             /* number of valid consecutive slots ending at head (>= 2) */
    /* inputs of the extra canvas channels (objects_channel / agents_channel / destination_channel), any may be NULL */
    const uint32_t *objects_bits; /* [S,N,T,X,E] maexp_obs_out.objects_bits of every slot */
    const uint8_t *destination;   /* [S,E,N] maexp_obs_out.destination (index of the agent's destination mask) */
    const uint32_t *dest_table;   /* [M,X] the destination masks as bit rows (bit y of word x) */
    int32_t type_objects, size_x; /* T, X of the two arrays above */
} maexp_replay_view;

typedef struct {
    float *obs;        /* [N,B,slot_floats] state      */
    float *next_obs;   /* [N,B,slot_floats] new_state  */
    int32_t *actions;  /* [N,B] */
    float *rewards;    /* [N,B] */
    uint8_t *dones;    /* [N,B] */
    int32_t *picked;   /* [B,2] (slot of the state, env) — may be NULL */
    int8_t *pos;       /* [N,B,2] observing agent's position in the state (relative / merged models) — may be NULL */
    int8_t *next_pos;  /* [N,B,2] ... in the new state */
    int32_t *error_flag; /* one int32, OR-ed with MAEXP_REPLAY_ERR_* by the sampler (never cleared by it) — may be NULL */
    /* extra canvas-channel inputs of the state and the new state (what maexp_learner_bind_canvas_inputs takes), may be NULL:
     * state["objects"] as bit rows, state["agents"], state["destination"] as bit rows (observation_handler.py:95-105) */
    uint32_t *obj_rows, *next_obj_rows;   /* [N,B,T,X] */
    int8_t *agents, *next_agents;         /* [N,B,num_agents,2] index-space cells of ALL agents (needs ring positions) */
    uint32_t *dest_rows, *next_dest_rows; /* [N,B,X] */
} maexp_batch;

/* maexp_batch.error_flag bits.  NO_VALID: 64 redraws and a 4096-pair walk found only transitions into reset slots (a ring of
 * 2-3 slots whose environments all reset on the same tick), the sampled row is not a transition; INDEX_RANGE: an injected
 * (slot, env) lay outside the ring and was clamped; INDEX_INVALID: an injected pair names a transition into a reset slot;
 * NOT_UNIQUE: the ring holds fewer distinct transitions than the batch asks for, a pick repeats.
 * A flagged row carries action 0 instead of the reset slot's -1, so the loss kernels never index out of bounds. */
enum { MAEXP_REPLAY_ERR_NO_VALID = 1, MAEXP_REPLAY_ERR_INDEX_RANGE = 2, MAEXP_REPLAY_ERR_INDEX_INVALID = 4,
       MAEXP_REPLAY_ERR_NOT_UNIQUE = 8 /* fewer distinct transitions than samples: the batch repeats one (np.random.choice raises) */ };

/* Uniform sampling of `batch` DISTINCT whole-team transitions (np.random.choice(len, batch, replace=False), buffer.py:61:
 * sequential rejection sampling, sample b draws again while it equals a pick of a lower index; batch > 1024 samples with
 * replacement) fused with the gather into per-agent batches.
 * indices: optional int32 [B,2] (slot, env) injected picks (NULL: Philox keyed by seed/tick). */
int maexp_replay_sample(const maexp_replay_view *ring, int32_t batch, uint32_t seed, uint32_t tick,
                        const int32_t *indices, const maexp_batch *out, void *stream);

/* ------------------------------------------------------------------------------------------
 * Learner: G = num_agents independent value networks updated in grouped launches.
 * generate_brain / generate_network / generate_optimizer / brain.learn
 * (core/agents/brains/__init__.py:42-85, models/__init__.py:33-105, optimizers/__init__.py:17-42). */
enum {
    MAEXP_LEARNER_DQN_CONV_MLP = 0, /* brain dqn / da3_baseline + model conv_mlp: da3_brain.py:58-89, conv_mlp.py:21-66 */
    MAEXP_LEARNER_C51 = 1,          /* categorical_dqn_brain.py:42-89, customs/categorical_dqn.py:28-79 */
    MAEXP_LEARNER_QR = 2,           /* qr_dqn_brain.py:81-119, customs/qr_dqn.py:20-65 */
    MAEXP_LEARNER_DQN_MLP = 3,      /* model mlp on the flattened observation (mlp.py:14-34) */
    MAEXP_LEARNER_IQN = 4,          /* iqn_brain.py:60-115, customs/iqn.py:21-162 (num_atoms = model.num_quantiles) */
    MAEXP_LEARNER_DA3 = 5,          /* brain da3 + model da3, local view: da3_brain.py:58-89, customs/da3.py:22-96, vit.py */
    MAEXP_LEARNER_DA6 = 6           /* brain da6 + model da6 (merged view): da6_brain.py:44-75, customs/da6.py:22-126 — a relative
                                       transformer on the area canvas (patch_size = model.relative_patch_size) feeds token 0 of the
                                       local transformer (local_patch_size) */
};
enum { MAEXP_BUF_ONLINE = 0, MAEXP_BUF_TARGET = 1, MAEXP_BUF_GRADS = 2, MAEXP_BUF_SQUARE_AVG = 3, MAEXP_BUF_SUPPORT = 4,
       MAEXP_BUF_EXP_AVG = 5,     /* Adam only (NULL for RMSprop); MAEXP_BUF_SQUARE_AVG is Adam's exp_avg_sq */
       MAEXP_BUF_ADAM_STEPS = 6   /* Adam only: int32 [G] step counters behind the float pointer */ };

typedef struct {
    int32_t kind;
    int32_t num_agents;      /* G */
    int32_t batch_size;      /* B */
    int32_t channels, visible_range, num_actions;
    int32_t hidden;          /* model.hidden_layer_sizes[0] */
    int32_t conv1_channels;  /* model.hidden_channels[0] */
    int32_t conv2_channels;  /* model.output_channel */
    int32_t num_atoms;       /* C51: model.num_atoms; QR: model.num_quantiles */
    int32_t obs_stride;      /* floats between consecutive observations of a batch (>= C*V*V) */
    float gamma;
    float v_min, v_max, delta_z;  /* C51 support bounds; delta_z = fp32((V_max - V_min)/(num_atoms - 1)) */
    float learning_rate, alpha, one_minus_alpha, eps;  /* optim.RMSprop, configs/opt/rmsprop.yaml */
    float max_norm;          /* clip_grad_norm_: 0.1 (dqn, qr, iqn), 10 (c51) */
    int32_t num_quantiles_dash, num_cosines;  /* IQN: model.num_quantiles_dash, model.num_cosines */
    int32_t embed_dim, num_heads, mlp_hidden, block_loop;  /* DA3: model.embed_dim, num_heads, int(embed_dim*mlp_ratio), block_loop */
    int32_t area_mask;        /* MAEXP_MASK_*: relative (conv_mlp on the decoded [C,X,Y] canvas, conv_mlp.py:56-66) and merged
                                 (MergedConvMLP conv_mlp.py:69-108, MergedDA3 da3.py:99-225) need the observing agents' positions */
    int32_t size_x, size_y;   /* map size for the relative canvas */
    int32_t patch_size;       /* model.patch_size: only sizes the parent's unused parameters of MergedDA3 / MergedDA3_IQN / DA6_IQN */
    int32_t local_patch_size; /* DA6: model.local_patch_size; DA6_IQN: 1 (da6_iqn.py:55-58) */
    int32_t relative_patch_size; /* DA6: model.relative_patch_size; DA6_IQN: 5 (da6_iqn.py:50-53) */
    int32_t iqn_trunk;        /* MAEXP_LEARNER_IQN only: which network produces the state embedding, MAEXP_IQN_TRUNK_* */
    int32_t optimizer;        /* generate_optimizer (optimizers/__init__.py:17-42): 0 = optim.RMSprop, 1 = optim.Adam */
    float beta1, beta2;       /* Adam betas (configs/opt/adam.yaml); learning_rate and eps are shared */
    int32_t attention_hard;   /* 1: model.attention == "hard" -- the transformer blocks are HardShrinkBlock
                                 (hard_shrink_attention.py:13-105; da3.py:46,144, da6.py:67, da3_iqn.py:118-122,245, da6_iqn.py:76-80):
                                 softmax weights <= af_lambd are zeroed before they multiply V and in the exported maps */
    float af_lambd;           /* model.af_lambd (nn.Hardshrink lambd), read when attention_hard */
    /* config.objects_channel / agents_channel / destination_channel: extra [T,X,Y] / [num_agents,X,Y] / [1,X,Y] channels
     * concatenated behind the decoded relative canvas (conv_mlp.py:25-27,61; da3.py:31-33,104; iqn.py:55-60,153-159;
     * da3_iqn.py:96-104,195-211; da6.py:32-37,170-177; da6_iqn.py:42-47,167-174).  Each family reads the flags its
     * reference constructor reads; subclasses that never use the parent's first layer still size it with them. */
    int32_t objects_channel, agents_channel, destination_channel;
    int32_t type_objects;     /* config.type_objects (T) */
    int32_t team_agents;      /* config.num_agents (the agents channel has one plane per agent of the team) */
} maexp_learner_cfg;

/* state-embedding producer under the IQN head (brain iqn / da3_iqn, iqn_brain.py:60-115, da3_iqn_brain.py:14-54) */
enum {
    MAEXP_IQN_TRUNK_CONV = 0,        /* model iqn: conv + state_embedder (customs/iqn.py:36-94) */
    MAEXP_IQN_TRUNK_DA3 = 1,         /* model da3_iqn, local view (customs/da3_iqn.py:87-190) */
    MAEXP_IQN_TRUNK_MERGED_DA3 = 2,  /* model da3_iqn, merged view (customs/da3_iqn.py:213-330) */
    MAEXP_IQN_TRUNK_DA6 = 3          /* model da6_iqn (customs/da6_iqn.py:24-176) */
};

typedef struct maexp_learner maexp_learner;

int maexp_learner_create(const maexp_learner_cfg *cfg, maexp_learner **out);
void maexp_learner_destroy(maexp_learner *learner);
/* parameters per agent, and the table (state_dict key, offset, count) that maps the flat fp32
 * buffers to the reference's `agent{i}.pth` tensors (abstract_loop_handler.py:149-180) */
int64_t maexp_learner_num_params(const maexp_learner *learner);
int maexp_learner_num_tensors(const maexp_learner *learner);
int maexp_learner_tensor(const maexp_learner *learner, int index, char *name, int name_len, int64_t *offset,
                         int64_t *count);
/* device buffers [G, num_params] (MAEXP_BUF_SUPPORT: [128] atoms z_j or tau_hat_j, written by the host) */
float *maexp_learner_buffer(maexp_learner *learner, int which);
/* Attention weights computed by the LAST maexp_learner_q_values / maexp_learner_q_values_iqn call — what forward_attn of the
 * attention models returns next to the Q-values (da3.py:64-88, da6.py:128-165, vit.py Block.forward_attn) and
 * AttentionAgent.get_action_attns stacks (attention_agent.py:14-22).  stack 0 = local transformer, 1 = relative
 * transformer (DA6 / DA6_IQN only); out fp32 [G, batch, heads, T, T], T = maexp_learner_attention_tokens(stack). */
int maexp_learner_attention_tokens(const maexp_learner *learner, int stack);
int maexp_learner_attention(maexp_learner *learner, int stack, int layer, int batch, float *out, void *stream);

/* Tensor-core GEMM of the large acting / evaluation batches (csrc/tc_gemm.cu: tcgen05.mma kind::tf32, accumulators in
 * tensor memory): C[M,N] = epi(A[M,K] * W[N,K]^T + bias) with fp32 device buffers (column tiles of 256); the layer products of
 * models/mlp.py:28-34 at inference batch sizes.  split = 3 evaluates every product as three TF32 terms (fp32-level
 * accuracy), split = 1 is plain TF32. */
int maexp_tc_gemm(const float *A, const float *W, const float *bias, float *C, int32_t M, int32_t N, int32_t K, int64_t lda,
                  int64_t ldw, int64_t ldc, int32_t relu, int32_t split, void *stream);

/* Tensor-core weight-gradient product of an update (csrc/tc_gemm.cu, tc_gemm_tn_kernel): C[M,N] = sum_k A[k][m] * B[k][n] with
 * both operands stored k-outermost (A = dY [K rows][lda], B = X [K rows][ldb]) -- grad_weight = dY^T X of nn.Linear
 * (models/mlp.py:28-34) and of the im2col form of nn.Conv2d (models/convolution.py:44-55); three TF32 terms per product.
 * K >= 1024 with a partials buffer: split-K slices write dense tiles partials[slice][M][N] (*ksplit_out slices, to be added
 * in slice order by the caller) instead of C; otherwise *ksplit_out = 1 and C (row stride ldc) holds the result.  The
 * learner uses it behind MAEXP_TC_WGRAD=1 (opt-in: measured slower than the CUDA-core split-K GEMMs inside an update). */
int maexp_tc_gemm_tn(const float *A, const float *B, float *C, float *partials, int64_t partials_floats, int32_t M, int32_t N,
                     int32_t K, int64_t lda, int64_t ldb, int64_t ldc, int32_t *ksplit_out, void *stream);

/* relative / merged models: index-space positions of the observing agent, int8 [G,B,2] for the state batch and for
 * the next-state batch (maexp_batch.pos / next_pos); used by the following grads / q_values calls */
int maexp_learner_bind_positions(maexp_learner *learner, const int8_t *pos, const int8_t *next_pos);
/* Inputs of the extra canvas channels for the following grads / q_values calls (device pointers, NULL where the flag is
 * off): state["objects"] as bit rows uint32 [G,B,T,X] (bit y of word x; maexp_batch.obj_rows), state["agents"] int8
 * [G,B,num_agents,2] (every agent's index-space cell) and state["destination"] as bit rows uint32 [G,B,X]
 * (observation_handler.py:95-105); the first of each pair belongs to the state batch, the second to the new state. */
int maexp_learner_bind_canvas_inputs(maexp_learner *learner, const uint32_t *obj_rows, const uint32_t *next_obj_rows,
                                     const int8_t *agents, const int8_t *next_agents, const uint32_t *dest_rows,
                                     const uint32_t *next_dest_rows);
/* Merged models with observation noise: the reference feeds them the NOISE-FREE "relative" key
 * (merged_observtion_handler.py:27,41-47), of which they read only the area channel (conv_mlp.py:100-108, da3.py:217-225,
 * da6.py:161-179) — a pure function of the observing agent's cell.  los_host: int8 [X,Y,V,V] HOST table (1 visible, 0 hidden;
 * generate_local_area_mask's return value, the same array maexp_tables.los takes); once set, the area canvas is built from
 * it instead of from the stored (noisy) crop.  NULL unsets. */
int maexp_learner_set_los(maexp_learner *learner, const int8_t *los_host);
/* hard_update (core/utils/updates.py:25-34) */
int maexp_learner_sync_target(maexp_learner *learner, void *stream);
/* brain.get_action's values: Q [G,batch,A] (C51: sum_j p_j z_j, QR: mean of quantiles) */
int maexp_learner_q_values(maexp_learner *learner, const float *obs /*[G,batch,obs_stride]*/, int batch,
                           int use_target, float *q_out, void *stream);
/* forward(s), forward(s'), TD target / projection, loss [G] and gradients into MAEXP_BUF_GRADS */
int maexp_learner_grads(maexp_learner *learner, const float *obs, const float *next_obs, const int32_t *actions,
                        const float *rewards, const uint8_t *dones, float *loss_out, void *stream);
/* IQN variants: taus fp32 [G,B,Nq] and tau_dashes [G,B,Nq'] are the per-sample quantile fractions
 * (torch.rand in iqn_brain.py:74,87), tau_sel [G,Nq] the shared fractions of the action-selecting forward
 * (customs/iqn.py:105-106).  Terminal states are not masked, as in the reference. */
int maexp_learner_q_values_iqn(maexp_learner *learner, const float *obs, int batch, int use_target, const float *tau_sel,
                               float *q_out, void *stream);
int maexp_learner_grads_iqn(maexp_learner *learner, const float *obs, const float *next_obs, const int32_t *actions,
                            const float *rewards, const uint8_t *dones, const float *taus, const float *tau_dashes,
                            const float *tau_sel, float *loss_out, void *stream);
/* grads *= grad_scale (1/world_size after an all-reduce), clip_grad_norm_, RMSprop step.
 * active: optional uint8 [G], 0 skips the agent (agent_tasks "-1", abstract_trainer.py:49-51) */
int maexp_learner_apply(maexp_learner *learner, float grad_scale, const uint8_t *active, float *grad_norm_out,
                        void *stream);
/* Data-parallel learner without a separate collective: ONE cooperative kernel does the gradient exchange
 * over NVLink peer memory and the optimizer step,  sum over ranks -> * 1/world -> clip_grad_norm_ -> RMSprop,  replacing
 * ncclAllReduce + maexp_learner_apply.  Every rank's gradient buffer must live in peer-mapped ("symmetric") memory:
 * maexp_learner_bind_grads points the learner's gradient output at such a buffer ([G, num_params] floats, 16-byte aligned,
 * zero-initialised; NULL re-binds the library's own buffer), maexp_dp_peers lists all ranks' mappings of it, of a
 * [world, G] float scratch and of a zero-initialised signal pad (>= 8 uint32 per rank).  All ranks must call it once per
 * update, on devices that can access each other's memory (one process per GPU); the kernel spins on the peers' signals.
 * Replicas end bit-identical: every reduced value and partial norm is computed by exactly one rank and read by all. */
typedef struct {
    int32_t world_size, rank;
    float *grads[8];
    float *partials[8];
    uint32_t *signals[8];
    float *multicast_grads;   /* optional NVLS multicast mapping of the same gradient buffers (NVSwitch): the reduce becomes one
                                 multimem.ld_reduce and the broadcast one multimem.st per 16 bytes; NULL: peer loads / stores */
} maexp_dp_peers;
int maexp_learner_bind_grads(maexp_learner *learner, float *external_grads);
int maexp_learner_apply_data_parallel(maexp_learner *learner, const maexp_dp_peers *peers, const uint8_t *active,
                                      float *grad_norm_out, void *stream);
/* grads + apply on one GPU */
int maexp_learner_update(maexp_learner *learner, const float *obs, const float *next_obs, const int32_t *actions,
                         const float *rewards, const uint8_t *dones, float *loss_out, void *stream);

/* ------------------------------------------------------------------------------------------
 * Stand-alone loss / optimizer kernels on CALLER-OWNED network outputs (fp32 device tensors, e.g. of the reference's own
 * torch modules): the TD target, loss [G] and the gradient with respect to the online network's output on the state batch,
 * which the caller back-propagates (out.backward(d_out)).  G independent networks per call (G = 1 for one brain).  They run the
 * same kernels as maexp_learner_grads / maexp_learner_apply; scratch memory is stream-ordered (cudaMallocAsync).
 *
 * maexp_loss_dqn: DA3Brain.learn / DQN TD target + MSE (da3_brain.py:69-87): q = net(s) [G,B,A], q_next_target = target(s'),
 *   dones uint8 [G,B] mask the bootstrap; dq_out [G,B,A].
 * maexp_loss_c51: CategoricalDQNBrain.learn (categorical_dqn_brain.py:51-87): RAW logits [G,B,A,Z] (ApplySoftmax.NONE) of
 *   net(s), net(s') (double-DQN action choice) and target(s'); support [Z] = linspace(V_min, V_max, Z), delta_z the fp32 atom
 *   spacing; projection with the reference's floor / ceil fix-ups, cross-entropy, no terminal mask (as the reference).
 * maexp_loss_qr: QRDQNBrain.learn (qr_dqn_brain.py:15-51,91-117): quantiles [G,B,A,Z]; tau_hats [Z] = (2j+1)/(2Z);
 *   element-wise (j <-> j) quantile Huber, kappa = 1.
 * maexp_loss_iqn: IQNBrain.learn (iqn_brain.py:15-50,73-113): get_quantiles outputs [G,B,Nq,A] of net(s, taus), net(s', tau_sel)
 *   (its mean over quantiles picks the action) and [G,B,Nd,A] of target(s', tau_dashes); taus [G,B,Nq]; pairwise quantile Huber.
 * maexp_clip_rmsprop_step: clip_grad_norm_(max_norm) over each network's flat gradient [G,count] (after grads *= grad_scale),
 *   then optim.RMSprop (optimizers/__init__.py:28-33; not centered, no momentum) on params / square_avg [G,count];
 *   active uint8 [G] or NULL; grad_norm_out [G] or NULL (the norm before clipping, clip_grad_norm_'s return value). */
int maexp_loss_dqn(const float *q, const float *q_next_target, const int32_t *actions, const float *rewards, const uint8_t *dones,
                   float gamma, int32_t G, int32_t B, int32_t A, float *loss_out, float *dq_out, void *stream);
int maexp_loss_c51(const float *logits, const float *logits_next_online, const float *logits_next_target, const int32_t *actions,
                   const float *rewards, const float *support, float gamma, float v_min, float v_max, float delta_z, int32_t G,
                   int32_t B, int32_t A, int32_t Z, float *loss_out, float *d_logits_out, void *stream);
int maexp_loss_qr(const float *quantiles, const float *quantiles_next_online, const float *quantiles_next_target,
                  const int32_t *actions, const float *rewards, const float *tau_hats, float gamma, int32_t G, int32_t B, int32_t A,
                  int32_t Z, float *loss_out, float *d_quantiles_out, void *stream);
int maexp_loss_iqn(const float *quantiles, const float *quantiles_next_online, const float *quantiles_next_target,
                   const float *taus, const int32_t *actions, const float *rewards, float gamma, int32_t G, int32_t B, int32_t A,
                   int32_t Nq, int32_t Nd, float *loss_out, float *d_quantiles_out, void *stream);
int maexp_clip_rmsprop_step(float *params, const float *grads, float *square_avg, int64_t count, int32_t G, float grad_scale,
                            float max_norm, float learning_rate, float alpha, float eps, const uint8_t *active,
                            float *grad_norm_out, void *stream);

/* ObservationHandler.decode_relative_state (observation_handler.py:150-216): fp32 obs rows
 * [rows, obs_stride] (C*V*V valid) + int8 index-space positions [rows,2] -> canvas [rows,C,X,Y]
 * (0 everywhere, last channel -1, window∩map pasted).  only_area_plus_one = 1 writes
 * canvas[:, -1:] + 1 -> [rows,1,X,Y], which is all the merged models read (conv_mlp.py:100-108). */
int maexp_decode_relative(const float *obs, const int8_t *positions, float *canvas, int64_t rows, int channels,
                          int visible_range, int size_x, int size_y, int obs_stride, int only_area_plus_one, void *stream);

/* heatmap_observation of DefaultEnvironment (default_environment.py:93,238-247), summed over the environments of one
 * step: heat[n][k] += sum_e obs[e*env_stride + n*agent_stride + k] for k < elems, skipping environments with skip[e] != 0
 * (pass maexp_obs_out.was_reset: a reset observation is not counted by the reference).  obs is a slot the step wrote
 * (local view: elems = C*V*V; ring slots are AGENT-major [N, E, slot_floats], so env_stride = slot_floats and
 * agent_stride = E*slot_floats) or the canvas maexp_decode_relative
 * made of it (relative / merged views: elems = C*X*Y).  heat: fp32 [N, elems], accumulated, zeroed by the caller. */
int maexp_heatmap_observation_add(const float *obs, int64_t num_envs, int num_agents, int elems, int64_t env_stride,
                                  int64_t agent_stride, const uint8_t *skip, float *heat, void *stream);

/* DefaultAgent.get_action for a chunk of environments (default_agent.py:12-19): with probability epsilon a
 * uniform random action int(random()*A) (abstract_agent.py:31-34), else argmax_a Q.  q_values [G,batch,A];
 * actions int8 [E,G] written at rows env_offset .. env_offset+batch; no_task uint8 [G]: agent_tasks "-1" agents
 * always act randomly (abstract_trainer.py:70-72).  Draws: Philox keyed (seed, global env id, tick). */
int maexp_select_actions(const float *q_values, int8_t *actions, int num_agents, int batch, int num_actions, int env_offset,
                         float epsilon, uint32_t seed, uint32_t first_global_env_id, uint32_t tick, const uint8_t *no_task,
                         void *stream);

/* Quantile fractions of the IQN brains: count floats uniform in [0,1) (24-bit mantissa draws, as torch.rand), Philox keyed
 * (seed, tick).  Replaces the torch.rand calls of iqn_brain.py:74,87 and customs/iqn.py:105-106 so that a graph-replayed
 * update reads fresh fractions from fixed buffers. */
int maexp_fill_uniform(float *out, int64_t count, uint32_t seed, uint32_t tick, void *stream);

const char *maexp_last_error(void);
int maexp_version(void);

#ifdef __cplusplus
}
#endif
#endif
