// C ABI of libdeepcomp_b200.so (include/deepcomp_b200.h): handle management, launch geometry, host-side glue.
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <new>
#include <vector>

#include "dcb_internal.h"

namespace {

thread_local char g_err[512] = "";

int fail(int code, const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

#define CU(call)                                                                                        \
    do {                                                                                                \
        cudaError_t e_ = (call);                                                                        \
        if (e_ != cudaSuccess)                                                                          \
            return fail(DCB_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, \
                        __LINE__);                                                                      \
    } while (0)

struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int dev) {
        cudaGetDevice(&prev);
        if (prev != dev) cudaSetDevice(dev);
        else prev = -1;
    }
    ~DeviceGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

// station.py:110-127 evaluated on the host with the host libm (the one the reference's NumPy scalars use)
double host_snr_of_d2(double c1, double c2, double d2) {
    const double d = sqrt(d2);
    const double pl = c1 + c2 * log10(d + DCB_EPSILON);
    const double signal = pow(10.0, (DCB_TX_POWER - pl) / 10.0);
    return signal / DCB_NOISE;
}

// Largest squared distance d2 with snr(sqrt(d2)) > SNR_THRESHOLD (station.py:224): the range decision becomes one
// fp64 compare on the device and agrees with the reference's own evaluation on this side of the boundary.
double threshold_d2(double c1, double c2) {
    double lo = 60.0 * 60.0, hi = 80.0 * 80.0;
    for (;;) {
        const double mid = 0.5 * (lo + hi);
        if (mid == lo || mid == hi) break;
        if (host_snr_of_d2(c1, c2, mid) > DCB_SNR_THRESHOLD) lo = mid;
        else hi = mid;
    }
    return lo;
}

// Largest squared distance d2 with fl(sqrt(d2)) <= vel (movement.py:142-145 as one compare; see snap_threshold in
// dcb_device.cuh -- sqrt is correctly rounded on both sides, so host and device agree bit for bit)
double host_snap_threshold(double vel) {
    if (!(vel > 0.0)) return 0.0;
    double c = vel * vel;
    while (sqrt(c) > vel) c = nextafter(c, 0.0);
    while (sqrt(nextafter(c, INFINITY)) <= vel) c = nextafter(c, INFINITY);
    return c;
}

// MathTables of dcb_math.cuh (inv, l2c, ex2, pwm, pwe: 5 x 16 doubles) followed by the snap thresholds of the drawn
// velocities 0..15, built with the host libm once per handle
void host_math_tables(double h, double c0, double *out) {
    for (int j = 0; j < 16; j++) {
        const double c = 1.0 + ((double)j + 0.5) / 16.0;
        const double inv = 1.0 / c;
        out[j] = inv;
        out[16 + j] = -log2(inv);
        out[32 + j] = exp2((double)j / 16.0);
        out[48 + j] = pow(inv, h);
        out[64 + j] = exp2(c0 - h * (double)j);
        out[80 + j] = host_snap_threshold((double)j);
    }
}

}  // namespace

struct dcb_env {
    dcb_config cfg;   // host pointers inside are NOT retained
    int device = 0;
    DevParams p;
    int threads = 0, grid = 0;
    size_t smem = 0;
    bool wide = false;   // one CTA per env (dcb_wide.cu) instead of the fused kernel (dcb_step.cu)
    int64_t launches = 0;
    // device allocations
    double *d_bs_xy = nullptr, *d_vel = nullptr, *d_init_xy = nullptr, *d_tabs = nullptr;
    uint16_t *d_pair_order = nullptr;
    int *d_sharing = nullptr;
    long long *d_seeds = nullptr;
    double2 *d_pos = nullptr, *d_init_pos = nullptr;
    uint2 *d_mv = nullptr;
    unsigned long long *d_mask = nullptr;
    double *d_ewma = nullptr;
    int *d_time = nullptr, *d_err = nullptr;
    uint32_t *d_table = nullptr, *d_pos_skip = nullptr, *d_mv_skip = nullptr;
    // variable UE population (dcb_population_event)
    int32_t *d_uid = nullptr;
    uint32_t *d_map_draws = nullptr, *d_glob_draws = nullptr;
    int na_reset = 0;        // UEs present after a reset (the original ue_list, base.py:176-182)
    bool table_shifted = false;   // dcb_extend_waypoints moved some table rows past the start of their streams
    bool pop_used = false;   // the population has changed at least once: resets go through the re-seeding path from then on
    long long *d_ue_seed = nullptr;                      // per original UE: seed / draws consumed since that seeding
    uint32_t *d_ue_pos_used = nullptr, *d_ue_mv_used = nullptr;
    double *d_vel_u = nullptr;                           // velocity spec per env and slot (slots change owners)
    int32_t *d_env_ids = nullptr;
    int env_ids_cap = 0;
    int32_t *d_uni_kind = nullptr;                       // UniformMovement (dcb_set_uniform_movement)
    double *d_uni_val = nullptr;
    std::vector<int> h_sharing;                          // sharing model per BS (host copy)
    std::vector<int32_t> h_uni_kind;
    std::vector<double> h_uni_val;
    // scripted policies (dcb_rollout)
    unsigned long long *d_cluster = nullptr;
    int32_t *d_fixed = nullptr;
    long long policy_calls = 0;
    // dcb_step_host staging
    int32_t *d_h_actions = nullptr;
    float *d_h_obs = nullptr, *d_h_reward = nullptr;
    uint8_t *d_h_lost = nullptr;
    size_t h_obs_cap = 0;    // floats d_h_obs holds (the observation size can grow: dcb_set_obs_variant)
    // dcb_step_many_host: actions of the fragment, two chunk-sized staging sets, copy stream, events
    int32_t *d_hm_actions = nullptr;
    size_t hm_actions_cap = 0;
    float *d_hm_obs[2] = {nullptr, nullptr}, *d_hm_reward[2] = {nullptr, nullptr};
    uint8_t *d_hm_lost[2] = {nullptr, nullptr};
    int hm_chunk_cap = 0;
    size_t hm_obs_cap = 0;   // floats per step the two staging sets were sized for
    cudaStream_t hm_copy_stream = nullptr;
    cudaEvent_t hm_done[2] = {nullptr, nullptr}, hm_copied[2] = {nullptr, nullptr};
};

namespace {

// Envs per CTA.  The fused kernel wants every CTA resident at once (one wave), few idle lanes in the last warp of
// each group, an even spread over the SMs, and -- when the batch is large enough to need several waves -- as many
// resident warps as registers and shared memory allow.
int choose_envs_per_cta(int K, int N, int M, int kind, int num_sms, size_t smem_cap, int var = 0) {
    const char *ov = getenv("DCB_ENVS_PER_CTA");
    if (ov && atoi(ov) > 0) return atoi(ov);
    int best_e = 1;
    double best_score = -1.0;
    const int max_group = N <= 256 ? 384 : 512;     // threads per warp group; the CTA has two groups
    for (int E = 1; E * N <= max_group && E <= K; E++) {
        const size_t sm = dcb_step_smem_bytes(kind, N, M, E, var);
        if (sm > smem_cap) break;
        const int group = (E * N + 31) / 32 * 32;
        const int threads = 2 * group;
        const int regs = dcb_step_regs_per_thread(threads, M);
        int per_sm = (int)(smem_cap / sm);
        if (2048 / threads < per_sm) per_sm = 2048 / threads;
        if (65536 / (regs * threads) < per_sm) per_sm = 65536 / (regs * threads);
        if (per_sm < 1) continue;
        const int grid = (K + E - 1) / E;
        const long slots = (long)num_sms * per_sm;
        const long waves = (grid + slots - 1) / slots;
        // SMs that get work at all, and -- over several waves -- how full the last wave is (free CTA slots of a
        // single wave cost nothing)
        const double sm_cover = (double)(grid < num_sms ? grid : num_sms) / num_sms;
        const double tail = waves > 1 ? (double)grid / (double)(waves * slots) : 1.0;
        // CTAs per SM are whole numbers: the busiest SM sets the time of a single wave
        const double balance = waves > 1 ? 1.0 : ((double)grid / num_sms) / (double)((grid + num_sms - 1) / num_sms);
        const double lane_util = (double)(E * N) / group;
        const double env_util = (double)K / ((double)grid * E);
        double resident = (double)grid / num_sms;
        if (resident > per_sm) resident = per_sm;
        const double warps = resident * threads / 32.0;
        const double occ = warps >= 32.0 ? 1.0 : 0.5 + 0.5 * warps / 32.0;       // latency hiding saturates
        const double score = lane_util * env_util * sm_cover * tail * (grid >= num_sms ? balance : 1.0) * occ;
        if (score > best_score) { best_score = score; best_e = E; }
    }
    return best_e;
}

// Launch geometry of the fused kernel for this handle: envs per CTA, threads, grid, shared memory, reducer lanes, and the
// reducer's pair order.  var: the handle observes a data-rate class (larger shared-memory layout).  p.sharing holds the
// sharing models on the device; host_sharing the same on the host.
int configure_fused(dcb_env *env, const int *host_sharing, int var) {
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, env->device));
    const size_t smem_cap = prop.sharedMemPerBlockOptin;
    DevParams &p = env->p;
    const int K = p.K, N = p.N, M = p.M;
    const int E = choose_envs_per_cta(K, N, M, p.kind, prop.multiProcessorCount, smem_cap, var);
    if (E * N > 512 || dcb_step_smem_bytes(p.kind, N, M, E, var) > smem_cap)
        return fail(DCB_ERR_INVALID_ARG, "%d envs per CTA do not fit the fused kernel (%zu B of shared memory)", E,
                    dcb_step_smem_bytes(p.kind, N, M, E, var));
    const int group = (E * N + 31) / 32 * 32;      // threads per warp group (physics / observers)
    env->threads = 2 * group;
    env->grid = (K + E - 1) / E;
    env->smem = dcb_step_smem_bytes(p.kind, N, M, E, var);
    // reducer lanes per (env, BS) pair: one per 32-UE bitset word, power of two, while the pairs still fit the CTA
    // (more lanes than words: the words are cut into 16- or 8-bit chunks)
    int S = 1, CS = 0;
    const int NW = (N + 31) / 32;
    while (S < 4 * NW && S * 2 <= 32 && E * M * S * 2 <= group) S *= 2;
    if (const char *ov = getenv("DCB_REDUCE_LANES")) {       // experiments: 1, 2, 4, 8, ...
        const int v = atoi(ov);
        if (v >= 1 && v <= 32 && (v & (v - 1)) == 0 && v <= 4 * NW) S = v;
    }
    while ((NW << CS) < S && CS < 2) CS++;
    p.E = E; p.S = S; p.CS = CS;
    // reducer order of a CTA's (env, BS) pairs: resource-fair base stations first (their factor is a bit count, no walk
    // over the link values), then the others, each group env-major
    std::vector<uint16_t> order;
    for (int pass = 0; pass < 2; pass++)
        for (int le = 0; le < E; le++)
            for (int b = 0; b < M; b++)
                if ((host_sharing[b] == DCB_SHARE_RESOURCE_FAIR) == (pass == 0)) order.push_back((uint16_t)(le * M + b));
    CU(cudaMemcpy(env->d_pair_order, order.data(), sizeof(uint16_t) * order.size(), cudaMemcpyHostToDevice));
    CU(dcb_step_set_smem_limit(env->threads, M, smem_cap));
    return DCB_OK;
}

// Move a handle to the one-CTA-per-env kernel (observation variants and the interference extension live there only)
int force_wide_kernel(dcb_env *env) {
    if (env->wide) return DCB_OK;
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, env->device));
    const DevParams &p = env->p;
    const WideLayout W = dcb_wide_layout(p.N, p.M, p.LC);
    if ((size_t)W.total > prop.sharedMemPerBlockOptin)
        return fail(DCB_ERR_UNSUPPORTED, "one env of %d UEs x %d BS needs %d B of shared memory on the wide kernel", p.N, p.M,
                    W.total);
    int threads = (p.N + 31) / 32 * 32;
    if (threads < 128) threads = 128;
    CU(dcb_wide_set_smem_limit(prop.sharedMemPerBlockOptin));
    env->wide = true;
    env->threads = threads;
    env->grid = p.K;
    env->smem = (size_t)W.total;
    env->p.E = 1;
    return DCB_OK;
}

int launch_step(dcb_env *env, const int32_t *d_actions, int T, const dcb_outputs *out, cudaStream_t s,
                const PolicyParams *pol = nullptr, int32_t *d_actions_out = nullptr, int flags = 0) {
    StepArgs a;
    a.flags = flags;
    memset(&a.pol, 0, sizeof(a.pol));
    if (pol) a.pol = *pol;
    a.actions_out = d_actions_out;
    a.p = env->p;
    a.L = dcb_smem_layout(env->p.kind, env->p.N, env->p.M, env->p.E, env->p.obs_var != 0 && !env->wide);
    a.W = dcb_wide_layout(env->p.N, env->p.M, env->p.LC);
    a.actions = d_actions;
    a.T = T;
    a.threads = env->threads;
    if (out) a.out = *out;
    else memset(&a.out, 0, sizeof(a.out));
    if (a.out.dbg_link_rate)
        CU(cudaMemsetAsync(a.out.dbg_link_rate, 0, sizeof(double) * (size_t)env->p.K * env->p.N * env->p.M, s));
    if (env->wide) CU(dcb_launch_wide(a, env->threads, env->grid, env->smem, s));
    else CU(dcb_launch_step(a, env->threads, env->grid, env->smem, s));
    env->launches++;
    return DCB_OK;
}

}  // namespace

extern "C" {

int dcb_abi_version(void) { return DCB_ABI_VERSION; }

const char *dcb_last_error(void) { return g_err; }

int64_t dcb_obs_size(const dcb_env *env) {
    const int64_t N = env->p.N, M = env->p.M;
    if (env->p.obs_var) return env->p.var_obs_size;
    return env->p.kind == DCB_KIND_CENTRAL ? 2 * N * M + N : N * (4 * M + 1);
}

int64_t dcb_reward_size(const dcb_env *env) { return env->p.kind == DCB_KIND_CENTRAL ? 1 : env->p.N; }

int64_t dcb_algorithmic_bytes_per_env_step(const dcb_env *env) {
    // state R+W per UE 56 + 8W (W = ceil(M/32) mask words), actions 4, obs f32, reward f32
    // = N (60 + 8W) + 4 obs_size + 4 reward_size, which also covers the observation variants
    const int64_t N = env->p.N, M = env->p.M, W = (M + 31) / 32;
    return N * (60 + 8 * W) + 4 * dcb_obs_size(env) + 4 * dcb_reward_size(env);
}

int64_t dcb_launch_count(const dcb_env *env) { return env->launches; }

const char *dcb_kernel_name(const dcb_env *env) { return env && env->wide ? "dcb_wide_kernel" : "dcb_step_kernel"; }

int dcb_launch_geometry(const dcb_env *env, int32_t *envs_per_cta, int32_t *threads, int32_t *smem_bytes,
                        int32_t *grid) {
    if (!env) return fail(DCB_ERR_INVALID_ARG, "null handle");
    if (envs_per_cta) *envs_per_cta = env->p.E;
    if (threads) *threads = env->threads;
    if (smem_bytes) *smem_bytes = (int32_t)env->smem;
    if (grid) *grid = env->grid;
    return DCB_OK;
}

void dcb_destroy(dcb_env *env) {
    if (!env) return;
    DeviceGuard g(env->device);
    cudaFree(env->d_bs_xy); cudaFree(env->d_vel); cudaFree(env->d_init_xy); cudaFree(env->d_sharing); cudaFree(env->d_tabs); cudaFree(env->d_pair_order);
    cudaFree(env->d_seeds); cudaFree(env->d_pos); cudaFree(env->d_init_pos); cudaFree(env->d_mv);
    cudaFree(env->d_mask); cudaFree(env->d_ewma); cudaFree(env->d_time); cudaFree(env->d_err);
    cudaFree(env->d_table); cudaFree(env->d_pos_skip); cudaFree(env->d_mv_skip); cudaFree(env->d_env_ids);
    cudaFree(env->d_cluster); cudaFree(env->d_fixed);
    cudaFree(env->d_uid); cudaFree(env->d_map_draws); cudaFree(env->d_glob_draws);
    cudaFree(env->d_ue_seed); cudaFree(env->d_ue_pos_used); cudaFree(env->d_ue_mv_used); cudaFree(env->d_vel_u);
    cudaFree(env->d_h_actions); cudaFree(env->d_h_obs); cudaFree(env->d_h_reward); cudaFree(env->d_h_lost);
    cudaFree(env->d_uni_kind); cudaFree(env->d_uni_val);
    cudaFree(env->d_hm_actions);
    for (int j = 0; j < 2; j++) {
        cudaFree(env->d_hm_obs[j]); cudaFree(env->d_hm_reward[j]); cudaFree(env->d_hm_lost[j]);
        if (env->hm_done[j]) cudaEventDestroy(env->hm_done[j]);
        if (env->hm_copied[j]) cudaEventDestroy(env->hm_copied[j]);
    }
    if (env->hm_copy_stream) cudaStreamDestroy(env->hm_copy_stream);
    delete env;
}

int dcb_create(const dcb_config *cfg, dcb_env **out) {
    if (!cfg || !out) return fail(DCB_ERR_INVALID_ARG, "null argument");
    *out = nullptr;
    if (cfg->abi_version != DCB_ABI_VERSION)
        return fail(DCB_ERR_INVALID_ARG, "abi_version %d != %d", cfg->abi_version, DCB_ABI_VERSION);
    const int K = cfg->num_envs, N = cfg->n_ue, M = cfg->n_bs;
    if (K < 1 || N < 1 || M < 1) return fail(DCB_ERR_INVALID_ARG, "num_envs, n_ue, n_bs must be >= 1");
    if (M > 64) return fail(DCB_ERR_UNSUPPORTED, "n_bs = %d > 64 (connection mask is one 64-bit word per UE)", M);
    if (N > 1024)
        return fail(DCB_ERR_UNSUPPORTED, "n_ue = %d > 1024 (one thread per UE, one CTA per env)", N);
    if (cfg->kind != DCB_KIND_CENTRAL && cfg->kind != DCB_KIND_MULTI) return fail(DCB_ERR_INVALID_ARG, "bad kind");
    if (cfg->reward < DCB_REWARD_AVG || cfg->reward > DCB_REWARD_MIN)
        return fail(DCB_ERR_INVALID_ARG, "bad reward aggregation %d", cfg->reward);   // central.py:73, multi_agent.py:92
    if (cfg->map_width < 1 || cfg->map_height < 1 || cfg->map_width >= 16384 || cfg->map_height >= 16384)
        return fail(DCB_ERR_UNSUPPORTED, "map %dx%d outside [1, 16383]", cfg->map_width, cfg->map_height);
    if (cfg->border_buffer <= 0)   // movement.py:103
        return fail(DCB_ERR_INVALID_ARG, "border_buffer must be > 0");
    if (cfg->map_width - cfg->border_buffer < cfg->border_buffer ||
        cfg->map_height - cfg->border_buffer < cfg->border_buffer)
        return fail(DCB_ERR_INVALID_ARG, "map smaller than twice the border buffer");
    if (cfg->pause_duration < 0 || cfg->pause_duration > 126)
        return fail(DCB_ERR_UNSUPPORTED, "pause_duration outside [0, 126]");
    if (cfg->episode_length < 1) return fail(DCB_ERR_INVALID_ARG, "episode_length must be >= 1");
    if (cfg->auto_reset && cfg->rand_episodes)
        return fail(DCB_ERR_UNSUPPORTED, "auto_reset replays the seeded episode; it cannot be combined with rand_episodes");
    if (!cfg->host_bs_xy || !cfg->host_sharing || !cfg->host_velocity || !cfg->host_init_xy || !cfg->host_seeds)
        return fail(DCB_ERR_INVALID_ARG, "null host array in config");
    bool has_maxcap = false, has_pf = false;
    for (int b = 0; b < M; b++) {
        const int s = cfg->host_sharing[b];
        if (s < 0 || s > 3) return fail(DCB_ERR_INVALID_ARG, "sharing[%d] = %d not supported", b, s);   // station.py:21
        has_maxcap |= s == DCB_SHARE_MAX_CAP;
        has_pf |= s == DCB_SHARE_PROPORTIONAL_FAIR;
    }
    for (int i = 0; i < N; i++) {
        const double v = cfg->host_velocity[i];
        if (!(v >= 0.0 || v == DCB_VELOCITY_SLOW || v == DCB_VELOCITY_FAST))
            return fail(DCB_ERR_INVALID_ARG, "velocity[%d] = %g", i, v);
    }

    int ndev = 0;
    CU(cudaGetDeviceCount(&ndev));
    if (cfg->device < 0 || cfg->device >= ndev) return fail(DCB_ERR_INVALID_ARG, "device %d of %d", cfg->device, ndev);
    DeviceGuard guard(cfg->device);
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, cfg->device));

    dcb_env *env = new (std::nothrow) dcb_env();
    if (!env) return fail(DCB_ERR_INVALID_ARG, "out of host memory");
    env->cfg = *cfg;
    env->cfg.host_bs_xy = nullptr; env->cfg.host_sharing = nullptr; env->cfg.host_velocity = nullptr;
    env->cfg.host_init_xy = nullptr; env->cfg.host_seeds = nullptr;
    env->device = cfg->device;

    const size_t smem_cap = prop.sharedMemPerBlockOptin;
    // Kernel choice: the fused, pipelined kernel (several envs per CTA, two threads per UE) when one env fits it, else
    // the wide kernel (one CTA per env, one thread per UE).  DCB_FORCE_WIDE=1 selects the wide kernel for any shape.
    const char *fw = getenv("DCB_FORCE_WIDE");
    env->wide = (fw && atoi(fw) > 0) || N > 512 || dcb_step_smem_bytes(cfg->kind, N, M, 1) > smem_cap;
    // link slots per UE of the wide kernel: every BS a UE is linked to is within range r of the UE, so those BS are
    // within 2r of each other: max_b #{b' : |b - b'| <= 2r} bounds the links any reachable state can hold
    // station.py:112-114 with the host libm, exactly as the reference evaluates them
    const double ch = 0.8 + (1.1 * log10(2500.0) - 0.7) * 1.5 - 1.56 * log10(2500.0);
    const double c1 = 69.55 + 26.16 * log10(2500.0) - 13.82 * log10(50.0) - ch;
    const double c2 = 44.9 - 6.55 * log10(50.0);
    const double thr_d2 = threshold_d2(c1, c2);
    int LC = 1;
    for (int b = 0; b < M; b++) {
        int c = 0;
        for (int b2 = 0; b2 < M; b2++) {
            const double dx = cfg->host_bs_xy[2 * b] - cfg->host_bs_xy[2 * b2];
            const double dy = cfg->host_bs_xy[2 * b + 1] - cfg->host_bs_xy[2 * b2 + 1];
            if (dx * dx + dy * dy <= 4.0 * thr_d2 * (1.0 + 1e-9)) c++;
        }
        if (c > LC) LC = c;
    }
    if (env->wide) {
        const WideLayout W = dcb_wide_layout(N, M, LC);
        if ((size_t)W.total > smem_cap) {
            delete env;
            return fail(DCB_ERR_UNSUPPORTED, "one env of %d UEs x %d BS (%d link slots) needs %d B of shared memory (> %zu)",
                        N, M, LC, W.total, smem_cap);
        }
        int threads = (N + 31) / 32 * 32;
        if (threads < 128) threads = 128;
        env->threads = threads;
        env->grid = K;
        env->smem = (size_t)W.total;
    }

    const size_t KN = (size_t)K * N;
    // pause_duration + 1 steps is the shortest possible redraw cycle (movement.py:168-181); +2 = entry 0 and slack
    const int D = cfg->episode_length / (cfg->pause_duration + 1) + 3;

#define ALLOC(ptr, count)                                                           \
    do {                                                                            \
        cudaError_t e_ = cudaMalloc((void **)&(ptr), sizeof(*(ptr)) * (count));     \
        if (e_ != cudaSuccess) {                                                    \
            dcb_destroy(env);                                                       \
            return fail(DCB_ERR_CUDA, "cudaMalloc(%s): %s", #ptr, cudaGetErrorString(e_)); \
        }                                                                           \
    } while (0)
    ALLOC(env->d_bs_xy, 2 * M); ALLOC(env->d_sharing, M); ALLOC(env->d_vel, N); ALLOC(env->d_init_xy, 2 * N);
    ALLOC(env->d_seeds, K); ALLOC(env->d_pos, KN); ALLOC(env->d_init_pos, KN); ALLOC(env->d_mv, KN);
    ALLOC(env->d_mask, KN); ALLOC(env->d_ewma, KN); ALLOC(env->d_time, K); ALLOC(env->d_err, 1);
    ALLOC(env->d_table, KN * D);
    ALLOC(env->d_tabs, 96);
    ALLOC(env->d_pair_order, (size_t)(512 / N + 1) * M);
    ALLOC(env->d_uid, KN); ALLOC(env->d_map_draws, K); ALLOC(env->d_glob_draws, K);
    if (cfg->rand_episodes) { ALLOC(env->d_pos_skip, K); ALLOC(env->d_mv_skip, KN); }
#undef ALLOC
    CU(cudaMemcpy(env->d_bs_xy, cfg->host_bs_xy, sizeof(double) * 2 * M, cudaMemcpyHostToDevice));
    CU(cudaMemcpy(env->d_sharing, cfg->host_sharing, sizeof(int) * M, cudaMemcpyHostToDevice));
    CU(cudaMemcpy(env->d_vel, cfg->host_velocity, sizeof(double) * N, cudaMemcpyHostToDevice));
    CU(cudaMemcpy(env->d_init_xy, cfg->host_init_xy, sizeof(double) * 2 * N, cudaMemcpyHostToDevice));
    CU(cudaMemcpy(env->d_seeds, cfg->host_seeds, sizeof(long long) * K, cudaMemcpyHostToDevice));
    CU(cudaMemset(env->d_err, 0, sizeof(int)));
    CU(cudaMemset(env->d_map_draws, 0, sizeof(uint32_t) * K));
    CU(cudaMemset(env->d_glob_draws, 0, sizeof(uint32_t) * K));
    CU(dcb_launch_iota_uid(env->d_uid, K, N, 0));
    env->na_reset = N;
    if (cfg->rand_episodes) {
        CU(cudaMemset(env->d_pos_skip, 0, sizeof(uint32_t) * K));
        CU(cudaMemset(env->d_mv_skip, 0, sizeof(uint32_t) * KN));
    }

    DevParams &p = env->p;
    memset(&p, 0, sizeof(p));
    p.K = K; p.N = N; p.NA = N; p.M = M; p.kind = cfg->kind; p.reward = cfg->reward;
    p.map_w = (double)cfg->map_width; p.map_h = (double)cfg->map_height;
    p.episode_length = cfg->episode_length; p.auto_reset = cfg->auto_reset; p.pause_duration = cfg->pause_duration;
    p.D = D; p.E = 1; p.S = 1; p.CS = 0; p.LC = LC;
    p.has_maxcap = has_maxcap; p.has_propfair = has_pf;
    p.util_step = 0; p.dr_req = 1.0;
    p.obs_maxnorm = 0;
    p.c1 = c1;
    p.c2 = c2;
    p.thr_d2 = thr_d2;
    p.snr_c0 = log2(10.0) * (DCB_TX_POWER - p.c1) / 10.0 - log2(DCB_NOISE);
    p.snr_h = p.c2 / 20.0;
    p.snr_hr = (float)(p.snr_h - 1.5);
    // (1 + r)^(-h) = sum_k binom(-h, k) r^k
    p.pw[0] = 1.0;
    for (int k = 1; k < 10; k++) p.pw[k] = p.pw[k - 1] * (-p.snr_h - (double)(k - 1)) / (double)k;
    {
        double tabs[96];
        host_math_tables(p.snr_h, p.snr_c0, tabs);
        CU(cudaMemcpy(env->d_tabs, tabs, sizeof(tabs), cudaMemcpyHostToDevice));
    }
    p.tabs = env->d_tabs;
    p.pair_order = env->d_pair_order;
    p.bs_xy = env->d_bs_xy; p.sharing = env->d_sharing; p.vel_spec = env->d_vel;
    p.pos = env->d_pos; p.mv = env->d_mv; p.mask = env->d_mask; p.ewma = env->d_ewma; p.time = env->d_time;
    p.init_pos = env->d_init_pos; p.table = env->d_table; p.err = env->d_err;

    // the attribute is per kernel, not per handle: always raise it to the device limit so that a later, smaller handle
    // of the same kernel class cannot lower it under an earlier one
    env->h_sharing.assign(cfg->host_sharing, cfg->host_sharing + M);
    cudaError_t e = cudaSuccess;
    if (env->wide) {
        e = dcb_wide_set_smem_limit(smem_cap);
        if (e != cudaSuccess) {
            dcb_destroy(env);
            return fail(DCB_ERR_CUDA, "cudaFuncSetAttribute(smem=%zu): %s", env->smem, cudaGetErrorString(e));
        }
    } else {
        const int rc = configure_fused(env, env->h_sharing.data(), 0);
        if (rc != DCB_OK) {
            dcb_destroy(env);
            return rc;
        }
    }

    // initial tables + state (as after the first reset)
    GenArgs g;
    g.K = K; g.N = N; g.D = D; g.W = cfg->map_width; g.H = cfg->map_height; g.border_buffer = cfg->border_buffer;
    g.seeds = env->d_seeds; g.vel_spec = env->d_vel; g.init_xy = env->d_init_xy; g.uni_kind = env->d_uni_kind;
    g.pos_skip = nullptr; g.mv_skip = nullptr; g.env_ids = nullptr; g.n_ids = 0;
    g.ue_seed = nullptr; g.ue_pos_skip = nullptr;
    g.init_pos = env->d_init_pos; g.table = env->d_table;
    ResetArgs r;
    r.K = K; r.N = N; r.D = D; r.env_ids = nullptr; r.n_ids = 0; r.init_pos = env->d_init_pos; r.table = env->d_table;
    r.pos = env->d_pos; r.mv = env->d_mv; r.mask = env->d_mask; r.ewma = env->d_ewma; r.time = env->d_time;
    r.pos_skip = nullptr;
    e = dcb_launch_generate(g, 0);
    if (e == cudaSuccess) e = dcb_launch_reset(r, 0);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
        dcb_destroy(env);
        return fail(DCB_ERR_CUDA, "table generation: %s", cudaGetErrorString(e));
    }
    env->launches += 2;
    *out = env;
    return DCB_OK;
}

int dcb_reset(dcb_env *env, const int32_t *host_env_ids, int32_t n, void *stream) {
    if (!env) return fail(DCB_ERR_INVALID_ARG, "null handle");
    DeviceGuard guard(env->device);
    cudaStream_t s = (cudaStream_t)stream;
    const DevParams &p = env->p;
    const int32_t *d_ids = nullptr;
    if (host_env_ids) {
        if (n < 0) return fail(DCB_ERR_INVALID_ARG, "negative env count");
        for (int j = 0; j < n; j++)
            if (host_env_ids[j] < 0 || host_env_ids[j] >= p.K)
                return fail(DCB_ERR_INVALID_ARG, "env id %d outside [0, %d)", host_env_ids[j], p.K);
        if (n == 0) return DCB_OK;
        if (n > env->env_ids_cap) {
            cudaFree(env->d_env_ids);
            env->d_env_ids = nullptr;
            CU(cudaMalloc((void **)&env->d_env_ids, sizeof(int32_t) * n));
            env->env_ids_cap = n;
        }
        CU(cudaMemcpyAsync(env->d_env_ids, host_env_ids, sizeof(int32_t) * n, cudaMemcpyHostToDevice, s));
        d_ids = env->d_env_ids;
    }
    if (env->pop_used) {
        // Reset of a batch whose population has changed (base.py:169-189).  MobileEnv.seed runs FIRST and walks the
        // current list -- the UE at list position p gets seed + 100 (p + 1), whoever it is; original UEs that left the
        // list keep their generators and continue their streams -- then the original list comes back and every UE of
        // it draws its position and first waypoint.  map.rng / the global `random` module restart (base.py:134-136).
        if (host_env_ids) return fail(DCB_ERR_UNSUPPORTED, "partial reset of a batch whose UE population changed");
        const size_t KN = (size_t)p.K * p.N;
        ReseedArgs r;
        r.K = p.K; r.N = p.N; r.NA = p.NA; r.n_orig = env->na_reset; r.seeds = env->d_seeds; r.uid = env->d_uid;
        r.ue_seed = env->d_ue_seed; r.ue_pos_used = env->d_ue_pos_used; r.ue_mv_used = env->d_ue_mv_used;
        CU(dcb_launch_pop_reseed(r, s));
        CU(dcb_launch_broadcast_vel(env->d_vel_u, env->d_vel, p.K, p.N, s));    // every original UE is back in its slot
        GenArgs g;
        g.K = p.K; g.N = p.N; g.D = p.D; g.W = env->cfg.map_width; g.H = env->cfg.map_height;
        g.border_buffer = env->cfg.border_buffer;
        g.seeds = env->d_seeds; g.vel_spec = env->d_vel; g.init_xy = env->d_init_xy; g.uni_kind = env->d_uni_kind;
        g.pos_skip = nullptr; g.mv_skip = env->d_ue_mv_used; g.env_ids = nullptr; g.n_ids = 0;
        g.ue_seed = env->d_ue_seed; g.ue_pos_skip = env->d_ue_pos_used;
        g.init_pos = env->d_init_pos; g.table = env->d_table;
        CU(dcb_launch_generate(g, s));
        CU(dcb_launch_add_u32(env->d_ue_pos_used, (long long)KN, 1u, s));      // this reset's reset_pos() draw
        CU(dcb_launch_iota_uid(env->d_uid, p.K, p.N, s));
        CU(cudaMemsetAsync(env->d_map_draws, 0, sizeof(uint32_t) * p.K, s));
        CU(cudaMemsetAsync(env->d_glob_draws, 0, sizeof(uint32_t) * p.K, s));
        env->launches += 4;
    }
    env->p.NA = env->na_reset;
    if (env->table_shifted && !env->cfg.rand_episodes && !env->pop_used) {
        // the episode restarts the seeded streams (base.py:171-173): bring the table rows of the envs that reset back to
        // the first draws
        CU(dcb_launch_table_cursor(p.K, p.N, d_ids, n, env->d_mv, env->d_mv_skip, 1, s));
        GenArgs g;
        g.K = p.K; g.N = p.N; g.D = p.D; g.W = env->cfg.map_width; g.H = env->cfg.map_height;
        g.border_buffer = env->cfg.border_buffer;
        g.seeds = env->d_seeds; g.vel_spec = env->d_vel; g.init_xy = env->d_init_xy; g.uni_kind = env->d_uni_kind;
        g.pos_skip = nullptr; g.mv_skip = nullptr; g.env_ids = d_ids; g.n_ids = n;
        g.ue_seed = nullptr; g.ue_pos_skip = nullptr;
        g.init_pos = env->d_init_pos; g.table = env->d_table;
        CU(dcb_launch_generate(g, s));
        env->launches += 2;
        if (!host_env_ids) env->table_shifted = false;
    }
    if (env->cfg.rand_episodes) {
        // base.py:171-173: no re-seed -> continue every UE's stream where the last episode left it
        CU(dcb_launch_advance_skip(p.K, p.N, d_ids, n, env->d_mv, env->d_mv_skip, env->d_pos_skip, s));
        GenArgs g;
        g.K = p.K; g.N = p.N; g.D = p.D; g.W = env->cfg.map_width; g.H = env->cfg.map_height;
        g.border_buffer = env->cfg.border_buffer;
        g.seeds = env->d_seeds; g.vel_spec = env->d_vel; g.init_xy = env->d_init_xy; g.uni_kind = env->d_uni_kind;
        g.pos_skip = env->d_pos_skip; g.mv_skip = env->d_mv_skip; g.env_ids = d_ids; g.n_ids = n;
        g.ue_seed = nullptr; g.ue_pos_skip = nullptr;
        g.init_pos = env->d_init_pos; g.table = env->d_table;
        CU(dcb_launch_generate(g, s));
        env->launches += 2;
    }
    ResetArgs r;
    r.K = p.K; r.N = p.N; r.D = p.D; r.env_ids = d_ids; r.n_ids = n; r.init_pos = env->d_init_pos;
    r.table = env->d_table; r.pos = env->d_pos; r.mv = env->d_mv; r.mask = env->d_mask; r.ewma = env->d_ewma;
    r.time = env->d_time; r.pos_skip = env->cfg.rand_episodes ? env->d_pos_skip : nullptr;
    CU(dcb_launch_reset(r, s));
    env->launches++;
    if (host_env_ids) CU(cudaStreamSynchronize(s));   // the id list is reused by the next partial reset
    return DCB_OK;
}

int dcb_extend_waypoints(dcb_env *env, void *stream) {
    if (!env) return fail(DCB_ERR_INVALID_ARG, "null handle");
    if (env->cfg.auto_reset)
        return fail(DCB_ERR_UNSUPPORTED, "auto_reset replays the seeded episode from the start of the table; it cannot be "
                                         "combined with continuous stepping");
    if (env->pop_used || env->p.NA < env->p.N)
        return fail(DCB_ERR_UNSUPPORTED, "continuous stepping past episode_length with a variable UE population");
    DeviceGuard guard(env->device);
    cudaStream_t s = (cudaStream_t)stream;
    const DevParams &p = env->p;
    if (!env->d_mv_skip) {
        CU(cudaMalloc((void **)&env->d_mv_skip, sizeof(uint32_t) * (size_t)p.K * p.N));
        CU(cudaMemsetAsync(env->d_mv_skip, 0, sizeof(uint32_t) * (size_t)p.K * p.N, s));
    }
    CU(dcb_launch_table_cursor(p.K, p.N, nullptr, 0, env->d_mv, env->d_mv_skip, 0, s));
    GenArgs g;
    g.K = p.K; g.N = p.N; g.D = p.D; g.W = env->cfg.map_width; g.H = env->cfg.map_height;
    g.border_buffer = env->cfg.border_buffer;
    g.seeds = env->d_seeds; g.vel_spec = env->d_vel; g.init_xy = env->d_init_xy; g.uni_kind = env->d_uni_kind;
    // (rand_episodes: the initial positions are re-derived for the episodes completed so far and overwritten by the next
    // reset anyway)
    g.pos_skip = env->cfg.rand_episodes ? env->d_pos_skip : nullptr; g.mv_skip = env->d_mv_skip;
    g.env_ids = nullptr; g.n_ids = 0; g.ue_seed = nullptr; g.ue_pos_skip = nullptr;
    g.init_pos = env->d_init_pos; g.table = env->d_table;
    CU(dcb_launch_generate(g, s));
    env->launches += 2;
    env->table_shifted = true;
    return DCB_OK;
}

int dcb_set_active_ues(dcb_env *env, int32_t n_active) {
    if (!env) return fail(DCB_ERR_INVALID_ARG, "null handle");
    if (n_active < 1 || n_active > env->p.N)
        return fail(DCB_ERR_INVALID_ARG, "n_active = %d outside [1, n_ue = %d]", n_active, env->p.N);
    env->p.NA = n_active;
    env->na_reset = n_active;      // the original ue_list: what dcb_reset goes back to after arrivals / departures
    return DCB_OK;
}

int32_t dcb_get_active_ues(const dcb_env *env) { return env ? env->p.NA : 0; }

int dcb_set_utility(dcb_env *env, int32_t kind, double dr_req) {
    if (!env) return fail(DCB_ERR_INVALID_ARG, "null handle");
    if (kind != DCB_UTILITY_LOG && kind != DCB_UTILITY_STEP)
        return fail(DCB_ERR_UNSUPPORTED, "Utility function %d not implemented!", kind);      // user.py:92
    env->p.util_step = kind == DCB_UTILITY_STEP;
    env->p.dr_req = dr_req;
    return DCB_OK;
}

int dcb_set_obs_norm(dcb_env *env, int32_t kind) {
    if (!env) return fail(DCB_ERR_INVALID_ARG, "null handle");
    if (kind != DCB_OBS_RELNORM && kind != DCB_OBS_MAXNORM)
        return fail(DCB_ERR_INVALID_ARG, "unknown observation normalisation %d", kind);
    if (kind == DCB_OBS_MAXNORM && env->p.obs_var)
        return fail(DCB_ERR_INVALID_ARG, "MaxNorm and a data-rate observation are different classes");
    env->p.obs_maxnorm = kind == DCB_OBS_MAXNORM;
    return DCB_OK;
}

int dcb_set_obs_variant(dcb_env *env, const dcb_obs_variant *v) {
    if (!env || !v) return fail(DCB_ERR_INVALID_ARG, "null argument");
    DevParams &p = env->p;
    if (v->kind == DCB_OBSVAR_NONE) {
        p.obs_var = 0;
        return DCB_OK;
    }
    if (v->kind != DCB_OBSVAR_NORMDR && v->kind != DCB_OBSVAR_DATARATE)
        return fail(DCB_ERR_INVALID_ARG, "unknown observation variant %d", v->kind);
    if (p.kind != DCB_KIND_CENTRAL)    // the reference has CentralNormDrEnv / CentralDrEnv (central.py:75-140) only
        return fail(DCB_ERR_UNSUPPORTED, "the data-rate observation classes exist for the central env only");
    if (p.obs_maxnorm) return fail(DCB_ERR_INVALID_ARG, "MaxNorm and a data-rate observation are different classes");
    bool tot = true, ues = false, dist = false, next = false;
    if (v->kind == DCB_OBSVAR_DATARATE) {
        if (v->dr_mode < DCB_DR_AUTO || v->dr_mode > DCB_DR_PLAIN) return fail(DCB_ERR_INVALID_ARG, "bad dr_mode");
        // variants.py:75-79
        if (v->curr_dr_obs && v->dr_mode != DCB_DR_AUTO)
            return fail(DCB_ERR_INVALID_ARG, "Enable all processing to add extra obs (curr_dr_obs needs dr_cutoff 'auto')");
        if (v->next_dist_obs && !v->dist_obs)
            return fail(DCB_ERR_INVALID_ARG, "Also enable 'dist_obs' when using 'next_dist_obs'");
        if (v->dr_mode != DCB_DR_AUTO && !(p.dr_req < v->dr_cutoff))      // variants.py:87-88
            return fail(DCB_ERR_INVALID_ARG, "dr_cutoff should be higher than max required dr. by UEs");
        if (v->next_dist_obs && p.uni_kind)
            return fail(DCB_ERR_UNSUPPORTED, "next_dist_obs needs RandomWaypoint UEs (step_towards_waypoint)");
        tot = v->curr_dr_obs != 0; ues = v->ues_at_bs_obs != 0; dist = v->dist_obs != 0; next = v->next_dist_obs != 0;
    }
    DeviceGuard guard(env->device);
    if (!env->wide) {
        // the fused kernel keeps this step's per-(env, BS) aggregates for its observers: a larger shared-memory layout,
        // possibly fewer envs per CTA; envs that do not fit go to the one-CTA-per-env kernel
        CU(cudaDeviceSynchronize());
        if (dcb_step_smem_bytes(p.kind, p.N, p.M, 1, 1) > 227u * 1024u || configure_fused(env, env->h_sharing.data(), 1) != DCB_OK) {
            const int rc = force_wide_kernel(env);
            if (rc != DCB_OK) return rc;
        }
    }
    // alphabetical key order (gym.spaces.Dict sorts; central.py:33-44): connected, dist, dr, dr_total, next_dist, ues_at_bs
    const int NM = p.N * p.M;
    int o = 0;
    p.vo_conn = o; o += NM;
    p.vo_dist = dist ? o : -1; o += dist ? NM : 0;
    p.vo_dr = o; o += NM;
    p.vo_tot = tot ? o : -1; o += tot ? p.N : 0;
    p.vo_next = next ? o : -1; o += next ? NM : 0;
    p.vo_ues = ues ? o : -1; o += ues ? NM : 0;
    p.var_obs_size = o;
    p.obs_var = v->kind;
    p.dr_mode = v->dr_mode;
    p.dr_cutoff = v->dr_cutoff;
    p.map_diag = sqrt((double)env->cfg.map_width * env->cfg.map_width + (double)env->cfg.map_height * env->cfg.map_height);
    return DCB_OK;
}

int dcb_set_interference(dcb_env *env, int32_t on) {
    if (!env) return fail(DCB_ERR_INVALID_ARG, "null handle");
    if (!on) {
        env->p.interference = 0;
        return DCB_OK;
    }
    DeviceGuard guard(env->device);
    const int rc = force_wide_kernel(env);
    if (rc != DCB_OK) return rc;
    CU(dcb_wide_upload_interference_constants(env->p.pw, env->p.snr_c0, env->p.snr_h));
    env->p.interference = 1;
    return DCB_OK;
}

int64_t dcb_num_joint_actions(const dcb_env *env) {
    if (!env) return 0;
    double n = pow((double)(env->p.M + 1), (double)env->p.NA);
    return n < 9.0e18 ? (int64_t)llround(n) : -1;
}

int dcb_test_actions(dcb_env *env, int32_t env_index, int64_t first, int64_t count, double *d_rewards, void *stream) {
    if (!env || !d_rewards) return fail(DCB_ERR_INVALID_ARG, "null argument");
    const DevParams &p = env->p;
    if (env_index < 0 || env_index >= p.K) return fail(DCB_ERR_INVALID_ARG, "env %d outside [0, %d)", env_index, p.K);
    if (p.NA > 16) return fail(DCB_ERR_UNSUPPORTED, "brute force over %d UEs (at most 16)", p.NA);
    if (p.interference) return fail(DCB_ERR_UNSUPPORTED, "brute force with the interference extension");
    const int64_t total = dcb_num_joint_actions(env);
    if (total < 0) return fail(DCB_ERR_UNSUPPORTED, "(%d + 1)^%d joint actions do not fit 63 bits", p.M, p.NA);
    if (first < 0 || count < 0 || first + count > total)
        return fail(DCB_ERR_INVALID_ARG, "candidates [%lld, %lld) outside [0, %lld)", (long long)first,
                    (long long)(first + count), (long long)total);
    if (count == 0) return DCB_OK;
    DeviceGuard guard(env->device);
    BruteArgs a;
    a.p = p; a.env = env_index; a.first = first; a.count = count; a.rewards = d_rewards;
    CU(dcb_launch_brute(a, (cudaStream_t)stream));
    env->launches++;
    return DCB_OK;
}

int dcb_get_ue_ids(dcb_env *env, int32_t *host_ids) {
    if (!env || !host_ids) return fail(DCB_ERR_INVALID_ARG, "null argument");
    DeviceGuard guard(env->device);
    CU(cudaDeviceSynchronize());
    const size_t n = (size_t)env->p.K * env->p.N;
    CU(cudaMemcpy(host_ids, env->d_uid, sizeof(int32_t) * n, cudaMemcpyDeviceToHost));
    for (size_t j = 0; j < n; j++) host_ids[j] &= ~DCB_UID_ARRIVED;
    return DCB_OK;
}

int dcb_population_event(dcb_env *env, int32_t n_add, int32_t n_remove, int32_t *d_actions, void *stream) {
    if (!env) return fail(DCB_ERR_INVALID_ARG, "null handle");
    if (n_add < 0 || n_remove < 0) return fail(DCB_ERR_INVALID_ARG, "negative UE count");
    if (n_add == 0 && n_remove == 0) return DCB_OK;
    const DevParams &p = env->p;
    if (p.NA - n_remove < 1) return fail(DCB_ERR_INVALID_ARG, "cannot remove %d of %d UEs", n_remove, p.NA);
    if (p.NA - n_remove + n_add > p.N)
        return fail(DCB_ERR_INVALID_ARG, "%d UEs would exceed max_ues = %d (base.py:84)", p.NA - n_remove + n_add, p.N);
    if (env->cfg.rand_episodes)
        return fail(DCB_ERR_UNSUPPORTED, "variable UE population needs rand_episodes = 0 (streams restart at reset)");
    if (env->p.uni_kind) return fail(DCB_ERR_UNSUPPORTED, "variable UE population with UniformMovement UEs");
    DeviceGuard guard(env->device);
    cudaStream_t s = (cudaStream_t)stream;
    if (!env->pop_used) {
        // per original UE: the seed of its generators and how far its two streams got (all as after a plain reset:
        // seed + 100 (i + 1), one reset_pos() draw, table row starting at the stream's first entry)
        const size_t KN = (size_t)p.K * p.N;
        CU(cudaMalloc((void **)&env->d_ue_seed, sizeof(long long) * KN));
        CU(cudaMalloc((void **)&env->d_ue_pos_used, sizeof(uint32_t) * KN));
        CU(cudaMalloc((void **)&env->d_ue_mv_used, sizeof(uint32_t) * KN));
        CU(cudaMalloc((void **)&env->d_vel_u, sizeof(double) * KN));
        CU(dcb_launch_broadcast_vel(env->d_vel_u, env->d_vel, p.K, p.N, s));
        env->p.vel_u = env->d_vel_u;
        CU(dcb_launch_pop_seed_init(env->d_ue_seed, env->d_ue_pos_used, env->d_ue_mv_used, env->d_seeds, p.K, p.N, s));
        CU(dcb_launch_add_u32(env->d_ue_pos_used, (long long)KN, 1u, s));
        env->pop_used = true;
    }
    PopArgs a;
    a.K = p.K; a.N = p.N; a.D = p.D; a.W = env->cfg.map_width; a.H = env->cfg.map_height;
    a.border_buffer = env->cfg.border_buffer;
    a.NA = p.NA; a.n_add = n_add; a.n_rem = n_remove;
    a.seeds = env->d_seeds; a.map_draws = env->d_map_draws; a.glob_draws = env->d_glob_draws; a.uid = env->d_uid;
    a.pos = env->d_pos; a.mv = env->d_mv; a.mask = env->d_mask; a.ewma = env->d_ewma; a.table = env->d_table;
    a.actions = d_actions;
    a.n_orig = env->na_reset; a.ue_mv_used = env->d_ue_mv_used; a.vel_u = env->d_vel_u;
    CU(dcb_launch_population(a, s));
    env->launches++;
    env->p.NA = p.NA - n_remove + n_add;
    return DCB_OK;
}

int dcb_observe(dcb_env *env, const dcb_outputs *out, void *stream) {
    if (!env) return fail(DCB_ERR_INVALID_ARG, "null handle");
    DeviceGuard guard(env->device);
    return launch_step(env, nullptr, 0, out, (cudaStream_t)stream);
}

int dcb_step(dcb_env *env, const int32_t *d_actions, const dcb_outputs *out, void *stream) {
    if (!env || !d_actions) return fail(DCB_ERR_INVALID_ARG, "null argument");
    DeviceGuard guard(env->device);
    return launch_step(env, d_actions, 1, out, (cudaStream_t)stream);
}

int dcb_step_no_move(dcb_env *env, const int32_t *d_actions, const dcb_outputs *out, void *stream) {
    if (!env || !d_actions) return fail(DCB_ERR_INVALID_ARG, "null argument");
    DeviceGuard guard(env->device);
    return launch_step(env, d_actions, 1, out, (cudaStream_t)stream, nullptr, nullptr, DCB_STEPF_NO_MOVE);
}

int dcb_set_uniform_movement(dcb_env *env, const int32_t *host_kind, const double *host_value) {
    if (!env || !host_kind || !host_value) return fail(DCB_ERR_INVALID_ARG, "null argument");
    const DevParams &p = env->p;
    if (env->pop_used || p.NA < p.N)
        return fail(DCB_ERR_UNSUPPORTED, "UniformMovement UEs with a variable UE population");
    bool any = false;
    for (int i = 0; i < p.N; i++) {
        const int kx = host_kind[2 * i], ky = host_kind[2 * i + 1];
        if (kx < 0 || kx > 3 || ky < 0 || ky > 3) return fail(DCB_ERR_INVALID_ARG, "uniform movement kind of UE %d", i);
        if ((kx == 0) != (ky == 0))
            return fail(DCB_ERR_INVALID_ARG, "UE %d: both components or none (UniformMovement has move_x and move_y)", i);
        if ((kx == 1 && !isfinite(host_value[2 * i])) || (ky == 1 && !isfinite(host_value[2 * i + 1])))
            return fail(DCB_ERR_INVALID_ARG, "UE %d: move_x / move_y must be finite numbers", i);
        any |= kx != 0;
    }
    DeviceGuard guard(env->device);
    CU(cudaDeviceSynchronize());
    if (!env->d_uni_kind) {
        CU(cudaMalloc((void **)&env->d_uni_kind, sizeof(int32_t) * 2 * p.N));
        CU(cudaMalloc((void **)&env->d_uni_val, sizeof(double) * 2 * p.N));
    }
    CU(cudaMemcpy(env->d_uni_kind, host_kind, sizeof(int32_t) * 2 * p.N, cudaMemcpyHostToDevice));
    CU(cudaMemcpy(env->d_uni_val, host_value, sizeof(double) * 2 * p.N, cudaMemcpyHostToDevice));
    env->h_uni_kind.assign(host_kind, host_kind + 2 * p.N);
    env->h_uni_val.assign(host_value, host_value + 2 * p.N);
    env->p.uni_kind = any ? env->d_uni_kind : nullptr;
    env->p.uni_val = any ? env->d_uni_val : nullptr;
    // the draw sequences of those UEs' movement generators change: regenerate the tables and start over (as dcb_create)
    if (env->cfg.rand_episodes) {
        CU(cudaMemset(env->d_pos_skip, 0, sizeof(uint32_t) * p.K));
        CU(cudaMemset(env->d_mv_skip, 0, sizeof(uint32_t) * (size_t)p.K * p.N));
    }
    GenArgs g;
    g.K = p.K; g.N = p.N; g.D = p.D; g.W = env->cfg.map_width; g.H = env->cfg.map_height;
    g.border_buffer = env->cfg.border_buffer;
    g.seeds = env->d_seeds; g.vel_spec = env->d_vel; g.init_xy = env->d_init_xy; g.uni_kind = env->d_uni_kind;
    g.pos_skip = nullptr; g.mv_skip = nullptr; g.env_ids = nullptr; g.n_ids = 0;
    g.ue_seed = nullptr; g.ue_pos_skip = nullptr;
    g.init_pos = env->d_init_pos; g.table = env->d_table;
    ResetArgs r;
    r.K = p.K; r.N = p.N; r.D = p.D; r.env_ids = nullptr; r.n_ids = 0; r.init_pos = env->d_init_pos;
    r.table = env->d_table; r.pos = env->d_pos; r.mv = env->d_mv; r.mask = env->d_mask; r.ewma = env->d_ewma;
    r.time = env->d_time; r.pos_skip = nullptr;
    CU(dcb_launch_generate(g, 0));
    CU(dcb_launch_reset(r, 0));
    CU(cudaDeviceSynchronize());
    env->launches += 2;
    return DCB_OK;
}

int dcb_step_many(dcb_env *env, const int32_t *d_actions, int32_t T, const dcb_outputs *out, void *stream) {
    if (!env || !d_actions) return fail(DCB_ERR_INVALID_ARG, "null argument");
    if (T < 1) return fail(DCB_ERR_INVALID_ARG, "T must be >= 1");
    DeviceGuard guard(env->device);
    return launch_step(env, d_actions, T, out, (cudaStream_t)stream);
}

int dcb_rollout(dcb_env *env, const dcb_policy *policy, int32_t T, int32_t *d_actions_out, const dcb_outputs *out,
                void *stream) {
    if (!env || !policy) return fail(DCB_ERR_INVALID_ARG, "null argument");
    if (T < 1) return fail(DCB_ERR_INVALID_ARG, "T must be >= 1");
    if (policy->kind < DCB_POLICY_3GPP || policy->kind > DCB_POLICY_RANDOM)
        return fail(DCB_ERR_INVALID_ARG, "unknown policy kind %d", policy->kind);
    if (env->p.obs_var || env->p.interference)
        return fail(DCB_ERR_UNSUPPORTED, "the scripted device policies read the RelNorm observation of the SNR model");
    if (env->p.obs_maxnorm)   // the agents pick by 'dr' (heuristics.py:19-38,44-65,86-108); MaxNorm caps it at 7e-6
        return fail(DCB_ERR_UNSUPPORTED, "the scripted device policies read the RelNorm observation; this handle is MaxNorm");
    DeviceGuard guard(env->device);
    cudaStream_t s = (cudaStream_t)stream;
    const DevParams &p = env->p;
    PolicyParams q;
    memset(&q, 0, sizeof(q));
    q.kind = policy->kind;
    q.call0 = policy->calls_before >= 0 ? policy->calls_before : env->policy_calls;
    q.seed = policy->seed;
    if (policy->kind == DCB_POLICY_DYNAMIC) {
        // heuristics.py:86-91: selected = {b: snr_b >= epsilon * best_snr}; snr ~ (d^2)^-h  =>  d2_b <= d2min * epsilon^(-1/h)
        if (!(policy->epsilon >= 0.0 && policy->epsilon <= 1.0))   // cli.py:100
            return fail(DCB_ERR_INVALID_ARG, "epsilon must be within [0, 1] but is %g", policy->epsilon);
        q.gain = policy->epsilon > 0.0 ? pow(policy->epsilon, -1.0 / p.snr_h) : INFINITY;
    }
    if (policy->kind == DCB_POLICY_STATIC) {
        if (!policy->host_cluster_masks) return fail(DCB_ERR_INVALID_ARG, "static clustering needs host_cluster_masks");
        if (!env->d_cluster) CU(cudaMalloc((void **)&env->d_cluster, sizeof(unsigned long long) * p.M));
        CU(cudaMemcpyAsync(env->d_cluster, policy->host_cluster_masks, sizeof(unsigned long long) * p.M,
                           cudaMemcpyHostToDevice, s));
        q.cluster = env->d_cluster;
    }
    if (policy->kind == DCB_POLICY_FIXED) {
        if (!policy->host_fixed_action) return fail(DCB_ERR_INVALID_ARG, "fixed agent needs host_fixed_action");
        if (policy->noop_interval < 0) return fail(DCB_ERR_INVALID_ARG, "noop_interval must be >= 0");
        for (int i = 0; i < p.N; i++)
            if (policy->host_fixed_action[i] < 0 || policy->host_fixed_action[i] > p.M)
                return fail(DCB_ERR_ACTION_RANGE, "fixed action %d of UE %d outside [0, %d]", policy->host_fixed_action[i], i, p.M);
        if (!env->d_fixed) CU(cudaMalloc((void **)&env->d_fixed, sizeof(int32_t) * p.N));
        CU(cudaMemcpyAsync(env->d_fixed, policy->host_fixed_action, sizeof(int32_t) * p.N, cudaMemcpyHostToDevice, s));
        q.fixed = env->d_fixed;
        q.noop_interval = policy->noop_interval;
    }
    const int rc = launch_step(env, nullptr, T, out, s, &q, d_actions_out);
    if (rc != DCB_OK) return rc;
    env->policy_calls += T;
    if (policy->kind == DCB_POLICY_STATIC || policy->kind == DCB_POLICY_FIXED)
        CU(cudaStreamSynchronize(s));   // the host arrays of the policy may be reused by the caller
    return DCB_OK;
}

int dcb_step_host(dcb_env *env, const int32_t *h_actions, float *h_obs, float *h_reward, uint8_t *h_lost_conn,
                  void *stream) {
    if (!env || !h_actions) return fail(DCB_ERR_INVALID_ARG, "null argument");
    DeviceGuard guard(env->device);
    cudaStream_t s = (cudaStream_t)stream;
    const DevParams &p = env->p;
    const size_t KN = (size_t)p.K * p.N;
    const size_t n_obs = (size_t)p.K * dcb_obs_size(env), n_rew = (size_t)p.K * dcb_reward_size(env);
    if (!env->d_h_actions) {
        CU(cudaMalloc((void **)&env->d_h_actions, sizeof(int32_t) * KN));
        CU(cudaMalloc((void **)&env->d_h_reward, sizeof(float) * n_rew));
        CU(cudaMalloc((void **)&env->d_h_lost, KN));
    }
    if (n_obs > env->h_obs_cap) {      // first call, or the observation grew since (dcb_set_obs_variant)
        CU(cudaStreamSynchronize(s));
        cudaFree(env->d_h_obs);
        env->d_h_obs = nullptr;
        env->h_obs_cap = 0;
        CU(cudaMalloc((void **)&env->d_h_obs, sizeof(float) * n_obs));
        env->h_obs_cap = n_obs;
    }
    CU(cudaMemcpyAsync(env->d_h_actions, h_actions, sizeof(int32_t) * KN, cudaMemcpyHostToDevice, s));
    dcb_outputs o;
    memset(&o, 0, sizeof(o));
    o.obs = h_obs ? env->d_h_obs : nullptr;
    o.reward = h_reward ? env->d_h_reward : nullptr;
    o.lost_conn = h_lost_conn ? env->d_h_lost : nullptr;
    const int rc = launch_step(env, env->d_h_actions, 1, &o, s);
    if (rc != DCB_OK) return rc;
    if (h_obs) CU(cudaMemcpyAsync(h_obs, env->d_h_obs, sizeof(float) * n_obs, cudaMemcpyDeviceToHost, s));
    if (h_reward) CU(cudaMemcpyAsync(h_reward, env->d_h_reward, sizeof(float) * n_rew, cudaMemcpyDeviceToHost, s));
    if (h_lost_conn) CU(cudaMemcpyAsync(h_lost_conn, env->d_h_lost, KN, cudaMemcpyDeviceToHost, s));
    CU(cudaStreamSynchronize(s));
    return DCB_OK;
}

int dcb_step_many_host(dcb_env *env, const int32_t *h_actions, int32_t T, float *h_obs, float *h_reward,
                       uint8_t *h_lost_conn, int32_t chunk_steps, void *stream) {
    if (!env || !h_actions) return fail(DCB_ERR_INVALID_ARG, "null argument");
    if (T < 1) return fail(DCB_ERR_INVALID_ARG, "T must be >= 1");
    DeviceGuard guard(env->device);
    cudaStream_t s = (cudaStream_t)stream;
    const DevParams &p = env->p;
    const size_t KN = (size_t)p.K * p.N;
    const size_t n_obs = (size_t)p.K * dcb_obs_size(env), n_rew = (size_t)p.K * dcb_reward_size(env);
    const size_t step_bytes = n_obs * 4 + n_rew * 4 + KN;
    // default: ~64 MB chunks (the two small copies per chunk -- reward, lost_conn -- then cost a few per cent of the chunk's
    // transfer time), but at least three chunks per call so that only the first chunk's kernel is exposed
    int C = chunk_steps > 0 ? chunk_steps : (int)((64u << 20) / step_bytes);
    if (chunk_steps <= 0 && C > (T + 2) / 3) C = (T + 2) / 3;
    if (C < 1) C = 1;
    if (C > T) C = T;
    if (!env->hm_copy_stream) {
        CU(cudaStreamCreateWithFlags(&env->hm_copy_stream, cudaStreamNonBlocking));
        for (int j = 0; j < 2; j++) {
            CU(cudaEventCreateWithFlags(&env->hm_done[j], cudaEventDisableTiming));
            CU(cudaEventCreateWithFlags(&env->hm_copied[j], cudaEventDisableTiming));
        }
    }
    if ((size_t)T * KN > env->hm_actions_cap) {
        CU(cudaStreamSynchronize(s));
        cudaFree(env->d_hm_actions);
        env->d_hm_actions = nullptr;
        env->hm_actions_cap = 0;
        CU(cudaMalloc((void **)&env->d_hm_actions, sizeof(int32_t) * (size_t)T * KN));
        env->hm_actions_cap = (size_t)T * KN;
    }
    if (C > env->hm_chunk_cap || n_obs > env->hm_obs_cap) {
        CU(cudaStreamSynchronize(s));
        CU(cudaStreamSynchronize(env->hm_copy_stream));
        for (int j = 0; j < 2; j++) {
            cudaFree(env->d_hm_obs[j]); cudaFree(env->d_hm_reward[j]); cudaFree(env->d_hm_lost[j]);
            env->d_hm_obs[j] = nullptr; env->d_hm_reward[j] = nullptr; env->d_hm_lost[j] = nullptr;
        }
        const int cap = C > env->hm_chunk_cap ? C : env->hm_chunk_cap;
        env->hm_chunk_cap = 0;
        env->hm_obs_cap = 0;
        for (int j = 0; j < 2; j++) {
            CU(cudaMalloc((void **)&env->d_hm_obs[j], sizeof(float) * n_obs * cap));
            CU(cudaMalloc((void **)&env->d_hm_reward[j], sizeof(float) * n_rew * cap));
            CU(cudaMalloc((void **)&env->d_hm_lost[j], KN * cap));
        }
        env->hm_chunk_cap = cap;
        env->hm_obs_cap = n_obs;
    }
    CU(cudaMemcpyAsync(env->d_hm_actions, h_actions, sizeof(int32_t) * (size_t)T * KN, cudaMemcpyHostToDevice, s));
    int chunk = 0;
    for (int t0 = 0; t0 < T; t0 += C, chunk++) {
        const int n = T - t0 < C ? T - t0 : C;
        const int j = chunk & 1;
        // staging set j is free once the copies of chunk - 2 have left it
        if (chunk >= 2) CU(cudaStreamWaitEvent(s, env->hm_copied[j], 0));
        dcb_outputs o;
        memset(&o, 0, sizeof(o));
        if (h_obs) { o.obs = env->d_hm_obs[j]; o.obs_stride = (int64_t)n_obs; }
        if (h_reward) { o.reward = env->d_hm_reward[j]; o.reward_stride = (int64_t)n_rew; }
        if (h_lost_conn) { o.lost_conn = env->d_hm_lost[j]; o.lost_conn_stride = (int64_t)KN; }
        const int rc = launch_step(env, env->d_hm_actions + (size_t)t0 * KN, n, &o, s);
        if (rc != DCB_OK) return rc;
        CU(cudaEventRecord(env->hm_done[j], s));
        CU(cudaStreamWaitEvent(env->hm_copy_stream, env->hm_done[j], 0));
        if (h_obs)
            CU(cudaMemcpyAsync(h_obs + (size_t)t0 * n_obs, env->d_hm_obs[j], sizeof(float) * n_obs * n,
                               cudaMemcpyDeviceToHost, env->hm_copy_stream));
        if (h_reward)
            CU(cudaMemcpyAsync(h_reward + (size_t)t0 * n_rew, env->d_hm_reward[j], sizeof(float) * n_rew * n,
                               cudaMemcpyDeviceToHost, env->hm_copy_stream));
        if (h_lost_conn)
            CU(cudaMemcpyAsync(h_lost_conn + (size_t)t0 * KN, env->d_hm_lost[j], KN * n, cudaMemcpyDeviceToHost,
                               env->hm_copy_stream));
        CU(cudaEventRecord(env->hm_copied[j], env->hm_copy_stream));
    }
    // the caller's stream is ordered behind the copies too (a later launch must not overwrite the staging sets early)
    CU(cudaStreamWaitEvent(s, env->hm_copied[(chunk - 1) & 1], 0));
    if (chunk >= 2) CU(cudaStreamWaitEvent(s, env->hm_copied[chunk & 1], 0));
    CU(cudaStreamSynchronize(s));
    return DCB_OK;
}

int dcb_check_errors(dcb_env *env, void *stream) {
    if (!env) return fail(DCB_ERR_INVALID_ARG, "null handle");
    DeviceGuard guard(env->device);
    cudaStream_t s = (cudaStream_t)stream;
    int flags = 0;
    CU(cudaMemcpyAsync(&flags, env->d_err, sizeof(int), cudaMemcpyDeviceToHost, s));
    CU(cudaStreamSynchronize(s));
    if (flags) {
        CU(cudaMemsetAsync(env->d_err, 0, sizeof(int), s));
        if (flags & DCB_ERRBIT_LINKS)
            return fail(DCB_ERR_UNSUPPORTED, "a UE held more than %d links (an unreachable state was injected); the excess "
                                             "links were dropped", env->p.LC);
        if (flags & DCB_ERRBIT_ACTION)
            return fail(DCB_ERR_ACTION_RANGE, "an action outside [0, %d] was passed to step (treated as no-op)", env->p.M);
        return fail(DCB_ERR_TABLE_EXHAUSTED, "a UE ran out of pre-drawn waypoints (stepped past episode_length "
                                             "without reset)");
    }
    return DCB_OK;
}

int dcb_get_state(dcb_env *env, dcb_state_host *st) {
    if (!env || !st) return fail(DCB_ERR_INVALID_ARG, "null argument");
    DeviceGuard guard(env->device);
    const DevParams &p = env->p;
    const size_t KN = (size_t)p.K * p.N;
    CU(cudaDeviceSynchronize());
    if (st->pos) CU(cudaMemcpy(st->pos, env->d_pos, sizeof(double2) * KN, cudaMemcpyDeviceToHost));
    if (st->mask) CU(cudaMemcpy(st->mask, env->d_mask, sizeof(uint64_t) * KN, cudaMemcpyDeviceToHost));
    if (st->ewma) CU(cudaMemcpy(st->ewma, env->d_ewma, sizeof(double) * KN, cudaMemcpyDeviceToHost));
    if (st->time) CU(cudaMemcpy(st->time, env->d_time, sizeof(int) * p.K, cudaMemcpyDeviceToHost));
    if (st->movement) {
        std::vector<uint2> mv(KN);
        std::vector<double> vel(env->d_vel_u ? KN : (size_t)p.N);
        CU(cudaMemcpy(mv.data(), env->d_mv, sizeof(uint2) * KN, cudaMemcpyDeviceToHost));
        CU(cudaMemcpy(vel.data(), env->d_vel_u ? env->d_vel_u : env->d_vel, sizeof(double) * vel.size(), cudaMemcpyDeviceToHost));
        for (size_t u = 0; u < KN; u++) {
            const double vf = env->d_vel_u ? vel[u] : vel[u % p.N];
            double *m = st->movement + 5 * u;
            if (env->p.uni_kind && env->h_uni_kind[2 * (u % p.N)]) {
                // UniformMovement: move_x, move_y (current sign), -1, 0, 0
                const size_t i = u % p.N;
                double mx = env->h_uni_kind[2 * i] == 1 ? env->h_uni_val[2 * i] : (double)(mv[u].x & 0xffffu);
                double my = env->h_uni_kind[2 * i + 1] == 1 ? env->h_uni_val[2 * i + 1] : (double)(mv[u].x >> 16);
                if (mv[u].y & 0x8000u) { mx = -mx; my = -my; }
                m[0] = mx; m[1] = my; m[2] = -1.0; m[3] = 0.0; m[4] = 0.0;
                continue;
            }
            m[0] = vf >= 0.0 ? vf : (double)(mv[u].y & 0xffu);
            m[1] = (double)(mv[u].x & 0xffffu);
            m[2] = (double)(mv[u].x >> 16);
            m[3] = (double)((mv[u].y >> 15) & 1u);
            m[4] = (double)((mv[u].y >> 8) & 0x7fu);
        }
    }
    return DCB_OK;
}

int dcb_set_state(dcb_env *env, const dcb_state_host *st) {
    if (!env || !st) return fail(DCB_ERR_INVALID_ARG, "null argument");
    DeviceGuard guard(env->device);
    const DevParams &p = env->p;
    const size_t KN = (size_t)p.K * p.N;
    CU(cudaDeviceSynchronize());
    if (st->pos) CU(cudaMemcpy(env->d_pos, st->pos, sizeof(double2) * KN, cudaMemcpyHostToDevice));
    if (st->mask) CU(cudaMemcpy(env->d_mask, st->mask, sizeof(uint64_t) * KN, cudaMemcpyHostToDevice));
    if (st->ewma) CU(cudaMemcpy(env->d_ewma, st->ewma, sizeof(double) * KN, cudaMemcpyHostToDevice));
    if (st->time) CU(cudaMemcpy(env->d_time, st->time, sizeof(int) * p.K, cudaMemcpyHostToDevice));
    if (st->movement) {
        // velocity / waypoint / pause state are injected; the table cursor (tidx) is kept
        std::vector<uint2> mv(KN);
        CU(cudaMemcpy(mv.data(), env->d_mv, sizeof(uint2) * KN, cudaMemcpyDeviceToHost));
        for (size_t u = 0; u < KN; u++) {
            const double *m = st->movement + 5 * u;
            if (env->p.uni_kind && env->h_uni_kind[2 * (u % p.N)]) continue;      // UniformMovement UEs keep their word
            const unsigned wx = (unsigned)m[1], wy = (unsigned)m[2];
            if (wx >= 16384u || wy >= 16384u) return fail(DCB_ERR_INVALID_ARG, "waypoint outside the map");
            const unsigned v = (unsigned)m[0] & 0xffu;
            const unsigned pause = ((m[3] != 0.0) ? 0x80u : 0u) | ((unsigned)m[4] & 0x7fu);
            mv[u].x = wx | (wy << 16);
            mv[u].y = v | (pause << 8) | (mv[u].y & 0xffff0000u);
        }
        CU(cudaMemcpy(env->d_mv, mv.data(), sizeof(uint2) * KN, cudaMemcpyHostToDevice));
    }
    return DCB_OK;
}

}  // extern "C"
