/*
 * deepcomp_b200 -- C ABI of the B200-native batched DeepCoMP environment step.
 *
 * This is the drop-in boundary for ONE path of CN-UPB/DeepCoMP: the mobile-cellular gym env step
 * (deepcomp/env/single_ue/base.py:413-466 `MobileEnv.step`, with the observation / reward variants
 * deepcomp/env/multi_ue/central.py:143-152 `CentralRelNormEnv` and deepcomp/env/multi_ue/multi_agent.py:6-107
 * `MultiAgentMobileEnv`).  The reference is pure Python and has no FFI; the entry points below are what a
 * ctypes binding inside the reference's env classes would call (INTEGRATION.md shows that binding).
 *
 * Conventions
 *   - plain C types only; every function returns 0 on success or a negative dcb_status, never throws;
 *     `dcb_last_error()` returns a thread-local message for the last failure.
 *   - one handle == one CUDA device == K independent env instances; a handle is not thread-safe, distinct
 *     handles are.
 *   - the caller owns every buffer it passes (device pointers unless the name says `host`); the library owns only
 *     its handle and the state slabs allocated in dcb_create.
 *   - all device work is enqueued on the `stream` argument (a cudaStream_t passed as void*; NULL = legacy default
 *     stream) and is asynchronous unless stated otherwise.
 *
 * Batched data layout (K envs, N UEs per env, M base stations, row-major, innermost last):
 *   actions     int32 [T][K][N]      0 = no-op, b+1 = toggle the link to BS b   (base.py:247-282, user.py:190-229)
 *   obs central float [T][K][2NM+N]  connected[N][M] | dr[N][M] | utility[N]    (central.py:31-57,147-152)
 *   obs multi   float [T][K][N][4M+1] connected[M] | dr[M] | ues_at_bs[M] | util_at_bs[M] | utility[1]
 *                                    (variants.py:271-303; alphabetical key order = RLlib's Dict flattening)
 *   reward      float central [T][K] (central.py:65-73) / multi [T][K][N] (multi_agent.py:39-95)
 *   lost_conn   uint8 [T][K][N]      links dropped by movement this step (user.py:175-188; returned by
 *                                    base.py:337-348 and discarded by the reference's step, base.py:447)
 */
#ifndef DEEPCOMP_B200_H
#define DEEPCOMP_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DCB_ABI_VERSION 1

typedef struct dcb_env dcb_env;

typedef enum dcb_status {
    DCB_OK = 0,
    DCB_ERR_INVALID_ARG = -1,
    DCB_ERR_UNSUPPORTED = -2,   /* shape outside what the kernels handle (see dcb_create) */
    DCB_ERR_CUDA = -3,
    DCB_ERR_ACTION_RANGE = -4,  /* an action outside [0, M] was seen on the device (base.py:238, central.py:61) */
    DCB_ERR_TABLE_EXHAUSTED = -5 /* waypoint table ran dry: call dcb_reset or dcb_extend_waypoints in time */
} dcb_status;

/* deepcomp/util/env_setup.py:23-37: 'central' -> CentralRelNormEnv, 'multi' -> MultiAgentMobileEnv */
typedef enum dcb_kind { DCB_KIND_CENTRAL = 0, DCB_KIND_MULTI = 1 } dcb_kind;
/* env_config['reward'] (central.py:19, multi_agent.py:19) */
typedef enum dcb_reward { DCB_REWARD_AVG = 0, DCB_REWARD_SUM = 1, DCB_REWARD_MIN = 2 } dcb_reward;
/* Basestation.sharing_model (station.py:152-202) */
typedef enum dcb_sharing {
    DCB_SHARE_RESOURCE_FAIR = 0, DCB_SHARE_RATE_FAIR = 1, DCB_SHARE_PROPORTIONAL_FAIR = 2, DCB_SHARE_MAX_CAP = 3
} dcb_sharing;

/*
 * Scripted baseline policies evaluated on the device (closed rollout loop, no host round trip per step):
 * deepcomp/agent/heuristics.py:13-187 (Heuristic3GPP, FullCoMP, DynamicSelection, StaticClustering) and
 * deepcomp/agent/dummy.py:6-50 (RandomAgent, FixedAgent).
 */
typedef enum dcb_policy_kind {
    DCB_POLICY_NONE = 0, DCB_POLICY_3GPP = 1, DCB_POLICY_FULLCOMP = 2, DCB_POLICY_DYNAMIC = 3, DCB_POLICY_STATIC = 4,
    DCB_POLICY_FIXED = 5, DCB_POLICY_RANDOM = 6
} dcb_policy_kind;

typedef struct dcb_policy {
    int32_t kind;                       /* dcb_policy_kind */
    int32_t noop_interval;              /* FIXED: no-op steps between repetitions (dummy.py:29-45) */
    double epsilon;                     /* DYNAMIC: scaling factor in [0, 1] (heuristics.py:79-91) */
    const uint64_t *host_cluster_masks; /* STATIC: [M] bitmask of the cluster each BS belongs to (heuristics.py:127-167) */
    const int32_t *host_fixed_action;   /* FIXED: [N] action per UE */
    uint64_t seed;                      /* RANDOM */
    int64_t calls_before;               /* compute_action calls THIS agent has made before the rollout (FIXED: phase of the
                                           no-op interval, dummy.py:37-45; RANDOM: position in its stream); < 0 = use the
                                           handle's running count of rollout steps (one agent per handle) */
} dcb_policy;

/* Velocity spec of a UE's RandomWaypoint (movement.py:112-117): a number >= 0 is used as is */
#define DCB_VELOCITY_SLOW (-1.0) /* rng.randint(1, 3)  */
#define DCB_VELOCITY_FAST (-2.0) /* rng.randint(5, 10) */

typedef struct dcb_config {
    int32_t abi_version;     /* DCB_ABI_VERSION */
    int32_t device;          /* CUDA device ordinal */
    int32_t kind;            /* dcb_kind */
    int32_t reward;          /* dcb_reward */
    int32_t num_envs;        /* K */
    int32_t n_ue;            /* N (fixed population: max_ues == num_ue, base.py:80-84) */
    int32_t n_bs;            /* M, 1..64 */
    int32_t map_width;       /* Map.width / Map.height are ints (map.py:20-21), < 16384 */
    int32_t map_height;
    int32_t episode_length;  /* env_config['episode_length']; sizes the waypoint table */
    int32_t rand_episodes;   /* env_config['rand_episodes'] (base.py:171-173): 0 = re-seed on every reset */
    int32_t auto_reset;      /* 1: an env whose time reached episode_length resets at the top of its next step (a benchmark
                                convenience: the observation returned AT the boundary is the old episode's last one, the
                                reset state itself is never observed -- learners should call dcb_reset + dcb_observe) */
    int32_t pause_duration;  /* RandomWaypoint(pause_duration=2) movement.py:87 */
    int32_t border_buffer;   /* RandomWaypoint(border_buffer=10) movement.py:87 */
    const double *host_bs_xy;     /* [M][2] BS positions (Basestation.pos) */
    const int32_t *host_sharing;  /* [M] dcb_sharing per BS (env_setup.py:40-49) */
    const double *host_velocity;  /* [N] DCB_VELOCITY_SLOW / _FAST / fixed number (env_setup.py:145-161) */
    const double *host_init_xy;   /* [N][2] initial position, NaN = 'random' (user.py:98-109) */
    const int64_t *host_seeds;    /* [K] env seeds; UE i (1-based) draws from random.Random(seed + 100*i) for both
                                     its position RNG and its movement RNG (base.py:132-143, user.py:94-96) */
} dcb_config;

/*
 * Output buffers of one dcb_step / dcb_step_many / dcb_observe call.  Every pointer is a device pointer and may
 * be NULL (= not wanted).  For dcb_step_many the `*_stride` fields give the distance in ELEMENTS between
 * consecutive steps (0 = every step overwrites the same buffer, i.e. only the last step survives).
 */
typedef struct dcb_outputs {
    float *obs;
    float *reward;
    uint8_t *lost_conn;
    float *curr_dr;      /* [K][N] info()['vector_metrics']['dr']      (base.py:407) */
    float *utility;      /* [K][N] info()['vector_metrics']['utility'] (base.py:408) */
    float *sum_utility;  /* [K]    info()['scalar_metrics']['sum_utility'] (base.py:402) */
    int64_t obs_stride, reward_stride, lost_conn_stride, curr_dr_stride, utility_stride, sum_utility_stride;
    /* fp64 taps for parity tests (same layouts as above, double precision, last step only) */
    double *dbg_obs;
    double *dbg_reward;
    double *dbg_snr;         /* [K][N][M] SNR at the post-move positions (station.py:122-127) */
    double *dbg_link_rate;   /* [K][N][M] cached ue.bs_dr values, 0 where not connected (user.py:143-146) */
    double *dbg_curr_dr;     /* [K][N] */
    double *dbg_utility;     /* [K][N] */
    double *dbg_sum_utility; /* [K] */
} dcb_outputs;

/* Host-side snapshot of the env state, for tests and checkpointing.  NULL members are skipped. */
typedef struct dcb_state_host {
    double *pos;        /* [K][N][2] */
    uint64_t *mask;     /* [K][N] bit b = connected to BS b */
    double *ewma;       /* [K][N] User.ewma_dr (user.py:148-157) */
    double *movement;   /* [K][N][5] velocity, waypoint x, waypoint y, pausing, curr_pause (movement.py:96-104) */
    int32_t *time;      /* [K] MobileEnv.time */
} dcb_state_host;

int dcb_abi_version(void);
const char *dcb_last_error(void);

/*
 * Allocate the state slabs for K envs on cfg->device and generate the per-UE RNG tables.  Supported shapes:
 * 1 <= M <= 64, N <= 1024.  Envs of up to 512 UEs whose working set (obs tile 4*(4M+1) + link values 8*(M|1) +
 * ~100 bytes per UE) fits the 227 KB of shared memory of one CTA run on the fused, pipelined kernel (several envs per
 * CTA); larger envs (BASELINE config 4: 1000 UE x 50 BS) run on the wide kernel, one CTA per env (dcb_kernel_name
 * tells which; the environment variable DCB_FORCE_WIDE=1 selects the wide kernel for any shape).  Synchronous.
 */
int dcb_create(const dcb_config *cfg, dcb_env **out);
void dcb_destroy(dcb_env *env);

/*
 * MobileEnv.reset (base.py:169-189) for all envs (env_ids == NULL) or for the n envs listed in the HOST array
 * env_ids.  Re-seeds when rand_episodes == 0.  Follow with dcb_observe to get the first observation.
 */
int dcb_reset(dcb_env *env, const int32_t *host_env_ids, int32_t n, void *stream);

/*
 * Continuous stepping past episode_length without a reset (the reference's --cont-train / soft_horizon; `done` is never
 * set, base.py:371-381).  The pre-drawn waypoint table of a UE covers episode_length steps from the last dcb_reset /
 * dcb_extend_waypoints; this call moves every UE's table row on to the draws from its cursor onwards (the per-UE MT19937
 * streams simply continue, as random.Random does in the reference).  Call it at least every episode_length steps when not
 * resetting (BatchedMobileEnv does).  Not with auto_reset or a variable UE population.  Asynchronous on `stream`.
 */
int dcb_extend_waypoints(dcb_env *env, void *stream);

/*
 * Variable UE population (base.py:80-84 `max_ues`, central.py:46-55 zero padding): of the n_ue = max_ues slots of every
 * env only slots [0, n_active) hold UEs.  Padding slots ignore their actions, keep their state, and read as zeros in
 * every output (observation rows / entries, reward, lost_conn, curr_dr, utility); per-env quantities (`ues_at_bs` =
 * |C_b| / num_ue variants.py:296, the central 'avg' reward central.py:65-73, sum_utility) count the UEs present.
 * Takes effect with the next launch; all envs of the handle share the count (they step in lockstep).  Default: n_ue.
 * This is the ORIGINAL population (base.py:52 original_ue_list): dcb_reset goes back to it after arrivals / departures.
 */
int dcb_set_active_ues(dcb_env *env, int32_t n_active);
int32_t dcb_get_active_ues(const dcb_env *env);

/*
 * Arrival / departure of UEs between the application of a step's actions and its rate update (base.py:429-443):
 * n_remove times `remove_ue` (base.py:610-617: a uniformly random UE, drawn from the global `random` module that
 * MobileEnv.seed seeded with the env seed; later UEs move up one slot), then n_add times `add_new_ue` (base.py:592-608:
 * id = last id + 1, position = Map.rand_border_point map.py:52-65, 'slow' RandomWaypoint seeded with env_seed + 100 id).
 * Call it right BEFORE the dcb_step of the step the event belongs to, with that step's device action buffer
 * (int32 [K][n_ue], edited in place: actions follow their UEs, arriving UEs get a no-op; may be NULL).  Every env of
 * the handle sees the same event (lockstep batch); needs rand_episodes = 0.  dcb_reset restores the original population
 * the way the reference does (MobileEnv.seed walks the list as it stands first, base.py:132-143, 169-189).
 */
int dcb_population_event(dcb_env *env, int32_t n_add, int32_t n_remove, int32_t *d_actions, void *stream);
/* UE ids per slot, host int32 [K][n_ue] (User.id as an integer; slots >= dcb_get_active_ues are stale).  Synchronous. */
int dcb_get_ue_ids(dcb_env *env, int32_t *host_ids);

/*
 * User.util_func (user.py:81-92; CLI --util): DCB_UTILITY_LOG (default, utility.py:36-54) or DCB_UTILITY_STEP
 * (utility.py:23-33: MAX_UTILITY when the UE's rate reaches dr_req, else MIN_UTILITY; User.dr_req defaults to 1).  One
 * setting per handle; takes effect with the next launch.  'linear' is not offered: the reference's own assert
 * (utility.py:18) rules it out for MIN/MAX_UTILITY = -20/20.
 */
typedef enum dcb_utility { DCB_UTILITY_LOG = 0, DCB_UTILITY_STEP = 1 } dcb_utility;
int dcb_set_utility(dcb_env *env, int32_t kind, double dr_req);

/*
 * Normalisation of the observation entry 'dr' (one setting per handle; takes effect with the next launch):
 * DCB_OBS_RELNORM (default): snr_b / max_b snr_b, RelNormEnv.get_ue_obs (single_ue/variants.py:276-284);
 * DCB_OBS_MAXNORM: (min(snr_b, 7e-6) - 2e-8) / (7e-6 - 2e-8), MaxNormEnv.get_ue_obs (single_ue/variants.py:308-332;
 * CentralMaxNormEnv multi_ue/central.py:155-164, the alternative named in util/env_setup.py:35) -- range [-0.0029, 1],
 * negative where the BS is out of range.  Everything else in the observation is unchanged.  The scripted device
 * policies (dcb_rollout) read the RelNorm observation and refuse a MaxNorm handle.
 */
typedef enum dcb_obs_norm { DCB_OBS_RELNORM = 0, DCB_OBS_MAXNORM = 1 } dcb_obs_norm;
int dcb_set_obs_norm(dcb_env *env, int32_t kind);

/*
 * The data-rate observation classes (one setting per handle, central kind only -- the reference has CentralNormDrEnv and
 * CentralDrEnv, multi_ue/central.py:75-140): every UE observes the SHARED rate it gets, or would get if it connected, from
 * every BS (Basestation.data_rate, station.py:204-220: 0 out of range; a UE that is not connected is counted in
 * temporarily) instead of the normalised SNR.
 *   DCB_OBSVAR_NORMDR   NormDrMobileEnv.get_ue_obs (single_ue/variants.py:198-250): dr = min(rate, 100) / 100,
 *                       dr_total = min(curr_dr, 100) / 100.
 *   DCB_OBSVAR_DATARATE DatarateMobileEnv.get_ue_obs (single_ue/variants.py:127-170) with its env_config options:
 *                       dr_mode AUTO min(rate - req, req) / req | SUB_REQ min(rate - req, dr_cutoff) | PLAIN
 *                       min(rate, dr_cutoff); curr_dr_obs adds dr_total = min(curr_dr - req, req) / req; ues_at_bs_obs the
 *                       number of UEs linked to each BS (per UE, not normalised); dist_obs the UE-BS distances / map
 *                       diagonal; next_dist_obs the same after the UE's next step towards its waypoint (movement.py:132-156).
 *                       req = dcb_set_utility's dr_req.
 * Layout: the keys present, in alphabetical order (connected, dist, dr, dr_total, next_dist, ues_at_bs), each as one
 * segment [N][M] (dr_total: [N]) -- central.py:31-57.  dcb_obs_size changes accordingly; the launch geometry is chosen
 * again (the fused kernel keeps a step's per-BS aggregates per step parity for its observers).  Call before the first
 * dcb_observe / dcb_step.
 */
typedef enum dcb_obs_variant_kind { DCB_OBSVAR_NONE = 0, DCB_OBSVAR_NORMDR = 1, DCB_OBSVAR_DATARATE = 2 } dcb_obs_variant_kind;
typedef enum dcb_dr_mode { DCB_DR_AUTO = 0, DCB_DR_SUB_REQ = 1, DCB_DR_PLAIN = 2 } dcb_dr_mode;
typedef struct dcb_obs_variant {
    int32_t kind;          /* dcb_obs_variant_kind */
    int32_t dr_mode;       /* dcb_dr_mode (DATARATE) */
    double dr_cutoff;      /* DATARATE, dr_mode != AUTO */
    int32_t curr_dr_obs, ues_at_bs_obs, dist_obs, next_dist_obs;   /* DATARATE options (variants.py:56-79) */
} dcb_obs_variant;
int dcb_set_obs_variant(dcb_env *env, const dcb_obs_variant *variant);

/*
 * Interference extension (NOT in the reference, which is SNR only: station.py:122-127, docs/model.md:15-19; named by
 * BASELINE.json's north star and config 4): with interference on, the signal quality of link (i, b) is
 *   SINR(i, b) = P(i, b) / (noise + sum_{b' != b} P(i, b')),  P = received power (station.py:116-120),
 * i.e. snr_b / (1 + sum_{b' != b} snr_b') in units of the noise floor, and it replaces the SNR everywhere the SNR is
 * used: can_connect (SINR > 2e-8), the unshared rate bw log2(1 + SINR), the observation entry 'dr' (SINR / max SINR).
 * The per-UE sum over the base stations is a warp-shuffle reduction over BS lanes.  Runs on the one-CTA-per-env kernel.
 * Not parity-graded against the reference; its only oracle is the restatement in oracle/dcb_oracle.c.  Default off.
 */
int dcb_set_interference(dcb_env *env, int32_t on);

/* get_obs() of the current state (central.py:31-57 / multi_agent.py:32-37) without stepping; reward is not written */
int dcb_observe(dcb_env *env, const dcb_outputs *out, void *stream);

/* MobileEnv.step (base.py:413-466) for all K envs: d_actions int32 [K][N] on the device */
int dcb_step(dcb_env *env, const int32_t *d_actions, const dcb_outputs *out, void *stream);

/*
 * A step WITHOUT the movement half: actions are applied, rates / rewards / observation are updated, but no UE moves, no
 * link is dropped, the EWMA rates and the env time stay.  SeqMultiAgentMobileEnv (multi_ue/multi_agent.py:110-179) calls
 * this for every UE of a round but the last (whose call is a plain dcb_step with that UE's action alone).
 */
int dcb_step_no_move(dcb_env *env, const int32_t *d_actions, const dcb_outputs *out, void *stream);

/*
 * UniformMovement UEs (util/movement.py:26-80): UE i moves by (move_x, move_y) every step and bounces off the map border
 * (both components flip when the next point would not be strictly inside the map).  host_kind int32 [N][2] per UE and
 * component: 0 = this UE keeps its RandomWaypoint (both components 0), 1 = the number in host_value [N][2], 2 = 'slow'
 * (randint(1, 5) from the UE's movement generator at every reset), 3 = 'fast' (randint(10, 20)).  Regenerates the RNG
 * tables and resets the batch (synchronous): call it right after dcb_create.  Not with a variable UE population.
 */
int dcb_set_uniform_movement(dcb_env *env, const int32_t *host_kind, const double *host_value);

/* T consecutive steps in ONE launch (state stays on chip between steps): d_actions int32 [T][K][N] */
int dcb_step_many(dcb_env *env, const int32_t *d_actions, int32_t T, const dcb_outputs *out, void *stream);

/*
 * T consecutive steps driven by a scripted policy: action(t+1) = policy(state after step t), the first action of a
 * launch from the current state (e.g. right after dcb_reset).  d_actions_out (int32 [T][K][N], may be NULL) receives
 * the actions taken.  The policy's call counter (FixedAgent interval, RandomAgent stream) persists in the handle.
 */
int dcb_rollout(dcb_env *env, const dcb_policy *policy, int32_t T, int32_t *d_actions_out, const dcb_outputs *out,
                void *stream);

/*
 * Convenience for host callers (the e2e path of bench.py and the K=1 gym facades): copies `h_actions` [K][N] to the
 * device, steps, copies obs / reward / lost_conn back into the host buffers and synchronises the stream.  Host
 * buffers should be pinned for full PCIe bandwidth; NULL outputs are skipped.
 */
int dcb_step_host(dcb_env *env, const int32_t *h_actions, float *h_obs, float *h_reward, uint8_t *h_lost_conn,
                  void *stream);

/*
 * T consecutive steps through host memory, pipelined: h_actions int32 [T][K][N] go to the device in one copy; the steps
 * run in chunks (a few steps per launch, dcb_step_many) into two device staging buffers, and while chunk c+1 computes,
 * chunk c's obs / reward / lost_conn travel to h_obs [T][...] / h_reward [T][...] / h_lost_conn [T][K][N] on a second
 * (library-owned) copy stream.  ONE stream synchronise at the end: the call returns when every output is in host memory.
 * PCIe-bound for every shape (the device produces 8.4 MB of observations per 5.6 us step at the headline shape), so
 * pinned host buffers are required for the overlap; NULL outputs are skipped.  chunk_steps <= 0 picks ~64 MB chunks (at
 * least three per call).
 */
int dcb_step_many_host(dcb_env *env, const int32_t *h_actions, int32_t T, float *h_obs, float *h_reward,
                       uint8_t *h_lost_conn, int32_t chunk_steps, void *stream);

/* Sticky device-side error flags (action range, table exhaustion) since the last call; synchronises the stream. */
int dcb_check_errors(dcb_env *env, void *stream);

/*
 * Brute-force search support: the central step reward (central.py:65-73, the handle's `reward` aggregation) of the joint
 * actions [first, first + count) of env `env_index`, each tried on the env's current state as MobileEnv.test_ue_actions
 * does (base.py:284-313: apply, rates, EWMA update, rates and rewards; the state itself is left untouched).  Candidate c
 * is the action vector whose entries are the digits of c in base M + 1, first UE = most significant digit
 * (agent/brute_force.py:26-62); BruteForceAgent takes the first maximum (brute_force.py:90-92).  d_rewards: device
 * double [count].  At most 16 UEs; dcb_num_joint_actions = (M + 1)^num_ue or -1 if it does not fit 63 bits.
 */
int64_t dcb_num_joint_actions(const dcb_env *env);
int dcb_test_actions(dcb_env *env, int32_t env_index, int64_t first, int64_t count, double *d_rewards, void *stream);

/* State snapshot / injection (synchronous). */
int dcb_get_state(dcb_env *env, dcb_state_host *state);
int dcb_set_state(dcb_env *env, const dcb_state_host *state);

/* Shape helpers */
int64_t dcb_obs_size(const dcb_env *env);     /* floats per env: 2NM+N (central) or N*(4M+1) (multi) */
int64_t dcb_reward_size(const dcb_env *env);  /* floats per env: 1 (central) or N (multi) */
/* Algorithmic HBM bytes of one env-step as defined in SURVEY.md section 8(d) (the roofline numerator) */
int64_t dcb_algorithmic_bytes_per_env_step(const dcb_env *env);
/* Kernel launches issued through this handle so far (bench.py's gpu_launches) */
int64_t dcb_launch_count(const dcb_env *env);
/* Name of the CUDA kernel that steps this handle: "dcb_step_kernel" (fused) or "dcb_wide_kernel" (one CTA per env) */
const char *dcb_kernel_name(const dcb_env *env);
/* Launch geometry chosen for the step kernel: envs per CTA, threads per CTA, dynamic smem bytes, grid size */
int dcb_launch_geometry(const dcb_env *env, int32_t *envs_per_cta, int32_t *threads, int32_t *smem_bytes,
                        int32_t *grid);

#ifdef __cplusplus
}
#endif
#endif /* DEEPCOMP_B200_H */
