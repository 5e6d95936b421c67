/* CPU oracle for the MAEXP hot path (gridworld step + egocentric observation).
 *
 * TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and the cpu_baseline /
 * `--impl reference` legs of bench.py may load this library; the product path
 * (maexp_b200/) never does and fails loudly when its CUDA extension is missing.
 *
 * Parity status: the reference ships NO tests, golden vectors or known-answer vectors
 *  — parity is unpinned by the reference's own tests.  This oracle is
 * instead pinned against the LIVE reference imported in the build container
 * (oracle/ref_runner.py → tests/golden/*.npz, tests/test_oracle_vs_golden.py).
 */
#ifndef MAEXP_ORACLE_H
#define MAEXP_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { ORC_MAP_SIMPLE = 0, ORC_MAP_FOUR_ROOMS = 1, ORC_MAP_THREE_ROOMS = 2, ORC_MAP_FOUR_RECTANGLE = 3,
       ORC_MAP_CENTRAL_ROOM = 4, ORC_MAP_CENTRAL_ROOM_DESTINATION = 5, ORC_MAP_CENTRAL_ROOM_LARGE = 6,
       ORC_MAP_CENTRAL_ROOM_LARGE_DESTINATION = 7, ORC_MAP_TWO_CENTRAL_ROOMS = 8,
       ORC_MAP_DYNAMIC_SIMPLE = 9 };
/* core/environments/__init__.py:40-71: which reward_ind / generate_objects override runs */
enum { ORC_ENV_DEFAULT = 0, ORC_ENV_SEQUENTIAL = 1, ORC_ENV_JUNCTIONAL = 2 };
enum { ORC_AV_DEFAULT = 0, ORC_AV_SIMPLE = 1, ORC_AV_INDIVIDUAL = 2, ORC_AV_TRANSITION = 3, ORC_AV_INDIVIDUAL_TRANSITION = 4 };
enum { ORC_OV_DEFAULT = 0, ORC_OV_SIMPLE = 1, ORC_OV_TYPES = 2 };
enum { ORC_MASK_LOCAL = 0, ORC_MASK_RELATIVE = 1, ORC_MASK_MERGED = 2 };
enum { ORC_NOISE_NONE = 0, ORC_NOISE_FLIP = 1, ORC_NOISE_DIST = 2, ORC_NOISE_THRESHOLD_DIST = 3, ORC_NOISE_FLAT = 4 };
enum { ORC_REWARD_DEFAULT = 0, ORC_REWARD_SHARED_MEAN = 1, ORC_REWARD_SHARED_MAX = 2 };

#define ORC_MAX_AGENTS 16
#define ORC_MAX_TYPES 8
#define ORC_MAX_RINGS 4

typedef struct {
    int32_t num_envs;              /* E */
    int32_t num_agents;            /* N */
    int32_t type_objects;          /* T */
    int32_t map_id;                /* ORC_MAP_* */
    int32_t size_x, size_y;        /* only read for ORC_MAP_SIMPLE / ORC_MAP_DYNAMIC_SIMPLE (core/maps/__init__.py:26-39) */
    int32_t visible_range;         /* V in {5,7,9} */
    int32_t past_step;             /* transition views */
    int32_t agent_view, object_view, area_mask, noise, reward_mode;
    int32_t transparent_observation;
    int32_t max_episode_length;
    int32_t num_objects;
    int32_t keep_objects_num;
    int32_t shuffle_init_xys;
    int32_t shared_max_is_mean;    /* reference "max" class computes the mean */
    int32_t auto_reset;            /* batched mode: an env whose done[0] fired resets on the next call */
    uint32_t seed;
    uint32_t first_global_env_id;
    int32_t init_xys[ORC_MAX_AGENTS][2];   /* centred coords, config.init_xys */
    int32_t task_mask[ORC_MAX_AGENTS];     /* bit t set: digit t in config.agent_tasks[i]; -1 for "-1" (no task) */
    int32_t task_order[ORC_MAX_AGENTS][ORC_MAX_TYPES]; /* digits in string order, -1 terminated */
    float flip_noise_range[ORC_MAX_RINGS]; /* fp32 0.5/(p-1), flip_observation_noise.py:28 */
    float std_distribution[ORC_MAX_RINGS];
    float noise_threshold[ORC_MAX_RINGS];  /* fp32(std*ratio), dist_observation_noise.py:60-66 */
    int32_t env_kind;              /* ORC_ENV_* */
    int32_t num_init_xys;          /* len(config.init_xys) >= num_agents (0: num_agents): default_environment.py:33,97-98
                                    * shuffles ALL entries and hands the first num_agents of the order to the agents */
} orc_cfg;

typedef struct orc_env orc_env;

orc_env *orc_create(const orc_cfg *cfg);
void orc_destroy(orc_env *);
int orc_channels(const orc_env *);      /* C */
int orc_obs_elems(const orc_env *);     /* C*V*V */
int orc_size_x(const orc_env *);
int orc_size_y(const orc_env *);

/* static tables in the reference's layout (for golden-hash checks) */
void orc_get_tables(const orc_env *, int8_t *wall /*[X,Y]*/, int8_t *objects_area /*[T,X,Y]*/,
                    int8_t *los /*[X,Y,V,V], 1 visible*/);

/* Outputs may be NULL.  obs/obs_rel: fp32 [E,N,C,V,V]; obs_rel only written for merged.
 * noise_tape: optional fp32 [E,N,C,V,V] draws (uniform for flip, N(0,1) for the others);
 *             NULL => Philox stream MAEXP_STREAM_NOISE (uniform) / Box-Muller on it. */
typedef struct {
    float *obs;
    float *obs_rel;
    float *rewards;        /* [E,N] */
    uint8_t *dones;        /* [E,N] */
    uint8_t *was_reset;    /* [E]: this call reset the env instead of stepping it */
    int32_t *coordinates;  /* [E,N,12] obs_x_min,obs_x_max,obs_y_min,obs_y_max,area_*(4),map_x_min,map_x_max,map_y_min,map_y_max */
    int8_t *agents;        /* [E,N,N,2] index-space positions as seen by agent i */
    int8_t *objects;       /* [E,N,T,X,Y] snapshot at agent i's observation time */
    int8_t *destination;   /* [E,N,X,Y] */
    const float *noise_tape;
    int32_t *heatmaps;     /* optional, summed over all envs, never zeroed here (default_environment.py:70-96 zeroes per reset):
                              [heatmap_agents N,X,Y | heatmap_complete N,X,Y | heatmap_objects T,X,Y |
                               heatmap_objects_left T,X,Y | heatmap_wall_collision X,Y | heatmap_agents_collision X,Y] */
} orc_out;

/* reset every env with env_mask[e]!=0 (NULL: all). tick keys the RNG draws. */
void orc_reset(orc_env *, uint32_t tick, const uint8_t *env_mask, const orc_out *out);
/* actions int8 [E,N] (NULL: Philox random policy), order uint8 [E,N] (NULL: Philox Fisher-Yates) */
void orc_step(orc_env *, uint32_t tick, const int8_t *actions, const uint8_t *order, const orc_out *out,
              int num_threads);

/* state taps */
void orc_get_state(const orc_env *, int8_t *pos /*[E,N,2] index space*/, int8_t *objects /*[E,T,X,Y]*/,
                   int32_t *counters /*[E,4] generated,completed,agents_collided,walls_collided*/,
                   int32_t *steps /*[E]*/, int8_t *destination /*[E,N,X,Y]*/);
/* overwrite state (hand-built scenarios, e.g. core/environments/test_environment.py:63-72) */
void orc_set_state(orc_env *, const int8_t *pos, const int8_t *objects);

/* RNG spec taps for the python tape objects */
void orc_philox(uint32_t k0, uint32_t k1, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t *out4);

#ifdef __cplusplus
}
#endif
#endif
