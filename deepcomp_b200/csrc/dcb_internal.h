// Internal declarations shared by the CUDA translation units of libdeepcomp_b200.so (not part of the C ABI).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/deepcomp_b200.h"

// Radio model constants: deepcomp/env/entities/station.py:10,26-30; deepcomp/util/constants.py:28,40-41
#define DCB_EPSILON 1e-16
#define DCB_BW 9e6
#define DCB_NOISE 1e-9
#define DCB_TX_POWER 30.0
#define DCB_SNR_THRESHOLD 2e-8
#define DCB_MIN_UTILITY (-20.0)
#define DCB_MAX_UTILITY 20.0
#define DCB_MAX_SNR_THRESHOLD 7e-6   // MaxNormEnv.MAX_SNR_THRESHOLD (single_ue/variants.py:311)

// Sticky device-side error bits (DevParams::err)
#define DCB_ERRBIT_ACTION 1
#define DCB_ERRBIT_TABLE 2
#define DCB_ERRBIT_LINKS 4    // a UE held more links than the wide kernel's slot capacity (unreachable state injected)

// Packed per-UE movement state (8 bytes):  x = wx | wy << 16,  y = vel | pause << 8 | tidx << 16
//   wx, wy : current waypoint (integers, movement.py:126-127)
//   vel    : drawn velocity 1..10 for 'slow'/'fast' UEs (movement.py:112-115); unused for fixed-velocity UEs
//   pause  : bit 7 = RandomWaypoint.pausing, bits 0..6 = curr_pause (movement.py:101-102)
//   tidx   : next unread entry of this UE's waypoint table
// UniformMovement UEs reuse the word: wx, wy = the drawn |move_x|, |move_y| ('slow' / 'fast' components), bit 7 of pause
// = "both components have flipped sign" (movement.py:76-78), tidx stays 1
// Packed waypoint-table entry (4 bytes): wx | wy << 14 | vel << 28

struct DevParams {
    int K, N, M, kind, reward;
    int NA;              // active UEs per env: slots [0, NA) of the N = max_ues slots exist (base.py:80-84, central.py:46-55)
    int episode_length, auto_reset, pause_duration;
    int D;               // waypoint-table depth per UE
    int E;               // envs per CTA
    int S;               // reducer lanes per (env, BS) pair, power of two <= 32
    int CS;              // log2 of the bitset chunks per 32-UE word the reducer lanes deal out (0: whole words)
    int has_maxcap, has_propfair;
    int util_step;       // 0: log utility (utility.py:36-54); 1: step utility (utility.py:23-33) at dr_req
    double dr_req;       // User.dr_req (user.py:17-30)
    int obs_maxnorm;     // observation 'dr': 0 = snr / max snr (variants.py:276-284); 1 = MaxNormEnv (variants.py:308-332)
    // data-rate observation classes (dcb_set_obs_variant; wide kernel, central layout): segment offsets in floats within an
    // env's observation, -1 = key absent; order = alphabetical keys (central.py:31-57)
    int obs_var;         // dcb_obs_variant_kind
    int dr_mode;         // dcb_dr_mode
    double dr_cutoff;
    double map_diag;     // Map.diagonal (map.py:26)
    int vo_conn, vo_dist, vo_dr, vo_tot, vo_next, vo_ues;
    int var_obs_size;    // floats per env of the variant observation
    int interference;    // extension: SINR instead of SNR (dcb_set_interference)
    int LC;              // wide kernel: link slots per UE (bound on the base stations any point can be in range of)
    float snr_hr;        // (float)(snr_h - 1.5): exponent left over by norm_snr_f32 (dcb_device.cuh), host-computed
    double thr_d2;       // largest squared distance that is still in range (snr > 2e-8, station.py:224)
    double c1, c2;       // Okumura-Hata constants (station.py:112-114)
    double snr_c0, snr_h; // snr(d) = 2^(snr_c0 - snr_h * log2(d^2)): the same model folded for the fast path
    double pw[10];       // binomial series of (1 + r)^(-snr_h): coefficients of r^0 .. r^9 (dcb_snr_inrange)
    const uint16_t *pair_order;  // [E*M] the (env, BS) pairs of a CTA ordered by sharing model (resource-fair pairs first:
                                 // their reduction is a bit count), so that a reducer warp works on one model
    const double *tabs;  // [80 + 16] MathTables (dcb_math.cuh) + snap thresholds of the drawn velocities 0..15, host-built
    const double *bs_xy; // [M][2]
    const int *sharing;  // [M]
    // UniformMovement UEs (util/movement.py:26-80): per slot and component 0 = none (RandomWaypoint), 1 = fixed number
    // (uni_val), 2 = 'slow' randint(1, 5), 3 = 'fast' randint(10, 20) drawn per reset (table entry 0); NULL = no such UE
    const int32_t *uni_kind; // [N][2]
    const double *uni_val;   // [N][2]
    double map_w, map_h;     // Map.width / height (map.py:20-21): UniformMovement bounces off the border
    const double *vel_spec;  // [N] velocity spec per slot (the same in every env) ...
    const double *vel_u;     // ... or [K*N] per env and slot once UEs have changed slots (variable population), else NULL
    // state slabs, flat UE index u = k*N + i
    double2 *pos;        // [K*N]
    uint2 *mv;           // [K*N]
    unsigned long long *mask;  // [K*N]
    double *ewma;        // [K*N]
    int *time;           // [K]
    const double2 *init_pos;   // [K*N] position drawn by reset_pos (user.py:98-109)
    const uint32_t *table;     // [K*N][D] successive movement.reset() draws (movement.py:110-130)
    int *err;
};

// ---- shared-memory layout of the step kernel (byte offsets), computed once on the host
struct SmemLayout {
    int off_tab, off_stage, off_x, off_fac_pre, off_fac_post, off_hx, off_hy,
        off_hmask, off_hutil, off_hrb, off_hdr, off_hlost, off_bsx, off_bsy, off_vel,
        off_arg_pre, off_arg_post, off_bits, off_share, off_links, off_vthr, off_wagg, off_snext, off_porder,
        off_vfac, off_varg, off_rcnt, off_rsum, off_rbest, off_hewma, off_hmv;   // data-rate observation classes (var)
    int nbits;   // words per bitset
    int links_per_warp;   // capacity (entries) of one physics warp's link list
    int wagg_pairs;       // (env, BS) pairs one observer warp aggregates for itself: the envs its 32 rows touch x M
    int wagg_stride;      // bytes of one observer warp's aggregate block
    int total;
};

__host__ __device__ inline int align16(int x) { return (x + 15) & ~15; }

__host__ __device__ inline int obs_width(int kind, int M) { return kind == DCB_KIND_CENTRAL ? 2 * M + 1 : 4 * M + 1; }

// Row stride (in doubles) of the [E*N][M] link matrix: odd, so that the 16 lanes of one 64-bit shared-memory access
// phase (consecutive UEs, same BS) hit 16 different bank pairs.
__host__ __device__ inline int row_stride(int M) { return M | 1; }

// var: the handle observes a data-rate class (dcb_set_obs_variant): the observers -- one step behind the physics warps -- then
// need that step's per-(env, BS) aggregates and a little more per-UE state, double-buffered by step parity
__host__ __device__ inline SmemLayout dcb_smem_layout(int kind, int N, int M, int E, int var = 0) {
    SmemLayout L;
    const int EN = E * N, EM = E * M;
    int o = 0;
    L.off_tab = o;      o += 5 * 16 * 8;                             // 5 x 128 B: one bank row per table
    L.off_stage = o;    o += align16(EN * obs_width(kind, M) * 4) + 16;   // float obs tile of the CTA (+ alignment shift)
    L.off_x = o;        o += align16(EN * row_stride(M) * 8);        // link values of connected links
    L.off_fac_pre = o;  o += align16(EM * 8);                        // per-(env, BS) sharing factor, next step's masks
    L.off_fac_post = o; o += align16(EM * 8);                        // ... current masks
    // per-(env, BS) utility aggregates and per-env sums the physics warps hand to the observers, two parities
    // physics -> observer hand-off, two parities: position, mask, utility, pre-move reward, rate, lost links
    L.off_hx = o;       o += align16(2 * EN * 8);
    L.off_hy = o;       o += align16(2 * EN * 8);
    L.off_hmask = o;    o += align16(2 * EN * 8);
    L.off_hutil = o;    o += align16(2 * EN * 8);
    L.off_hrb = o;      o += align16(2 * EN * 8);
    L.off_hdr = o;      o += align16(2 * EN * 8);
    L.off_hlost = o;    o += align16(2 * EN * 4);
    L.off_bsx = o;      o += align16(M * 8);
    L.off_bsy = o;      o += align16(M * 8);
    L.off_vel = o;      o += align16(N * 8);
    L.off_arg_pre = o;  o += align16(EM * 4);
    L.off_arg_post = o; o += align16(EM * 4);
    L.nbits = EM * ((N + 31) / 32);
    L.off_bits = o;     o += align16(6 * L.nbits * 4);               // UE bitsets per (env, BS): post[3], pre[2], fresh
    L.off_share = o;    o += align16(M * 4);
    // per physics warp: compacted list of the warp's links (owner lane, BS, membership flags), 2 bytes per entry;
    // worst case every UE of the warp is linked to every BS
    L.links_per_warp = 32 * M;
    L.off_links = o;    o += align16(((EN + 31) / 32) * L.links_per_warp * 2);
    L.off_vthr = o;     o += align16(16 * 8);                        // snap thresholds for drawn velocities 0..15
    L.off_snext = o;    o += align16(EN * 4);                        // per UE: prefetched waypoint-table entry
    L.off_porder = o;   o += align16(EM * 2);                        // (env, BS) pairs in the reducer's order (by sharing model)
    // per observer warp: utility aggregates of the (env, BS) pairs its rows need -- usum, umin (double), cnt (int),
    // f_ues, f_util (float) -- computed by the warp itself so that observer warps never synchronise with each other
    L.wagg_pairs = ((31 / N + 2) * M + 1) & ~1;
    L.wagg_stride = align16(L.wagg_pairs * 28);
    L.off_wagg = o;     o += ((EN + 31) / 32) * L.wagg_stride;
    L.off_vfac = L.off_varg = L.off_rcnt = L.off_rsum = L.off_rbest = L.off_hewma = L.off_hmv = 0;
    if (var) {
        L.off_vfac = o;  o += align16(2 * EM * 8);     // sharing factors / arg-max of the current masks, per parity
        L.off_rsum = o;  o += align16(2 * EM * 8);     // raw sums of the link values
        L.off_rbest = o; o += align16(2 * EM * 8);     // largest unshared rate (max-cap)
        L.off_varg = o;  o += align16(2 * EM * 4);
        L.off_rcnt = o;  o += align16(2 * EM * 4);     // linked UEs
        L.off_hewma = o; o += align16(2 * EN * 8);     // hand-off: EWMA rate, packed movement word
        L.off_hmv = o;   o += align16(2 * EN * 8);
    }
    L.total = o;
    return L;
}

// ---- shared-memory layout of the wide-env kernel (dcb_wide.cu: one CTA per env)
struct WideLayout {
    int off_tab, off_bsxy, off_share, off_vthr, off_xs, off_rec, off_sew, off_smv,
        off_bits, off_fac,
        off_arg, off_cnt, off_usum, off_umin, off_fues, off_futil, off_env,
        off_sdr, off_ssum, off_smax, off_sbmax, off_lsum, off_lbest, off_lcnt;   // general instance: curr_dr / interference sum per UE, raw link aggregates per BS
    int total;
};

__host__ __device__ inline WideLayout dcb_wide_layout(int N, int M, int LC) {
    WideLayout L;
    const int NW = (N + 31) / 32;
    int o = 0;
    L.off_tab = o;   o += 5 * 16 * 8;
    L.off_bsxy = o;  o += align16(M * 16);
    L.off_vthr = o;  o += align16(16 * 8);
    L.off_xs = o;    o += align16(N * LC * 8);
    L.off_rec = o;   o += align16(N * 48);                   // per UE: position, utility, mask, in-range set, reward (UeRec)
    L.off_sew = o;   o += align16(N * 8);
    L.off_smv = o;   o += align16(N * 8);
    L.off_fac = o;   o += align16(M * 8);
    L.off_usum = o;  o += align16(M * 8);
    L.off_umin = o;  o += align16(M * 8);
    L.off_env = o;   o += 16;
    L.off_bits = o;  o += align16(M * NW * 4);
    L.off_share = o; o += align16(M * 4);
    L.off_arg = o;   o += align16(M * 4);
    L.off_cnt = o;   o += align16(M * 4);
    L.off_fues = o;  o += align16(M * 4);
    L.off_futil = o; o += align16(M * 4);
    L.off_sdr = o;   o += align16(N * 8);
    L.off_ssum = o;  o += align16(N * 8);
    L.off_smax = o;  o += align16(N * 8);
    L.off_sbmax = o; o += align16(N * 4);
    L.off_lsum = o;  o += align16(M * 8);
    L.off_lbest = o; o += align16(M * 8);
    L.off_lcnt = o;  o += align16(M * 4);
    L.total = o;
    return L;
}

// Scripted per-UE policy evaluated on the device (reference deepcomp/agent/heuristics.py, dummy.py); kind 0 = none:
// actions come from StepArgs::actions.
struct PolicyParams {
    int kind;                          // dcb_policy_kind
    int noop_interval;                 // FixedAgent(noop_interval)
    double gain;                       // DynamicSelection: epsilon^(-20/c2) -- "snr >= eps * best" as a d^2 ratio
    const unsigned long long *cluster; // StaticClustering: [M] bitmask of the cluster of each BS
    const int32_t *fixed;              // FixedAgent: [N] action per UE
    long long call0;                   // compute_action calls made before this launch (FixedAgent interval, RNG stream)
    unsigned long long seed;           // RandomAgent
};

// StepArgs::flags
#define DCB_STEPF_NO_MOVE 1   // apply actions, rates, rewards, observation -- no movement, link drop, EWMA update or time
                              // increment (SeqMultiAgentMobileEnv between the UEs of one round, multi_agent.py:149-179)

struct StepArgs {
    DevParams p;
    SmemLayout L;
    WideLayout W;            // used by the wide kernel instead of L
    const int32_t *actions;  // [T][K][N], or NULL when a policy drives the envs
    int T;                   // 0 = observe only
    int threads;             // CTA size of the launch
    PolicyParams pol;
    int flags;               // DCB_STEPF_*
    int32_t *actions_out;    // [T][K][N] actions the policy took, or NULL
    dcb_outputs out;
};

struct GenArgs {
    int K, N, D, W, H, border_buffer;
    const long long *seeds;     // [K]
    const double *vel_spec;     // [N]
    const int32_t *uni_kind;    // [N][2] UniformMovement components (DevParams::uni_kind) or NULL
    const double *init_xy;      // [N][2]
    const uint32_t *pos_skip;   // [K] reset_pos() calls already consumed per env (rand_episodes) or NULL
    const uint32_t *mv_skip;    // [K*N] movement.reset() calls already consumed per UE or NULL
    const long long *ue_seed;   // [K*N] seed per UE instead of seeds[k] + 100 (i + 1), or NULL (variable population)
    const uint32_t *ue_pos_skip; // [K*N] reset_pos() calls already consumed per UE (instead of pos_skip[k]) or NULL
    const int32_t *env_ids;     // [n_ids] or NULL = all envs
    int n_ids;
    double2 *init_pos;          // [K*N]
    uint32_t *table;            // [K*N][D]
};

struct ResetArgs {
    int K, N, D;
    const int32_t *env_ids;
    int n_ids;
    const double2 *init_pos;
    const uint32_t *table;
    double2 *pos;
    uint2 *mv;
    unsigned long long *mask;
    double *ewma;
    int *time;
    uint32_t *pos_skip;   // rand_episodes: bumped once per reset env, else NULL
};

// uid slab: User.id as an integer; arrivals carry this flag (an arrival's id is the last id + 1 and can repeat the id of
// an original UE that has left -- they are different UEs with different generators)
#define DCB_UID_ARRIVED 0x40000000

// Arrival / departure of UEs (single_ue/base.py:433-443, 592-617), one thread per env
struct PopArgs {
    int K, N, D, W, H, border_buffer;
    int NA;                     // UEs present before the event (the same in every env)
    int n_add, n_rem;           // removals first, then arrivals
    const long long *seeds;     // [K] env seeds: map.rng and the global `random` module are both seeded with them
    uint32_t *map_draws;        // [K] 32-bit outputs consumed from map.rng since the last seeding
    uint32_t *glob_draws;       // [K] ... from the global `random` module
    int32_t *uid;               // [K*N] UE id per slot (ids of arriving UEs continue after the last UE's id)
    double2 *pos;
    uint2 *mv;
    unsigned long long *mask;
    double *ewma;
    uint32_t *table;            // [K*N][D]
    int32_t *actions;           // [K][N] this step's actions (follow their UEs when slots shift) or NULL
    int n_orig;                 // UEs of the original list (ids 1..n_orig)
    uint32_t *ue_mv_used;       // [K*N] movement.reset() draws an ORIGINAL UE has consumed since its last seeding
    double *vel_u;              // [K*N] velocity spec per env and slot (moves with its UE; arrivals are 'slow')
};

// reset() of a batch whose population changed (single_ue/base.py:169-189): MobileEnv.seed first re-seeds the UEs of the
// CURRENT list by list position, then the original list comes back
struct ReseedArgs {
    int K, N, NA, n_orig;
    const long long *seeds;     // [K]
    const int32_t *uid;         // [K*N] current list (ids per slot)
    long long *ue_seed;         // [K*N] per ORIGINAL UE: seed of its generators
    uint32_t *ue_pos_used;      // [K*N] ... reset_pos() draws since that seeding
    uint32_t *ue_mv_used;       // [K*N] ... movement.reset() draws since that seeding
};
// The general (PAD) template instances carry everything that is not the measured fixed-population RelNorm path: padding
// slots, observation variants, UniformMovement, ...
__host__ inline bool dcb_step_needs_general(const DevParams &p, int flags) {
    return p.NA < p.N || p.obs_maxnorm || p.obs_var || p.uni_kind || (flags & DCB_STEPF_NO_MOVE);
}

cudaError_t dcb_launch_pop_reseed(const ReseedArgs &a, cudaStream_t s);
cudaError_t dcb_launch_pop_seed_init(long long *ue_seed, uint32_t *pos_used, uint32_t *mv_used, const long long *seeds,
                                     int K, int N, cudaStream_t s);
cudaError_t dcb_launch_add_u32(uint32_t *a, long long n, uint32_t v, cudaStream_t s);
cudaError_t dcb_launch_broadcast_vel(double *vel_u, const double *vel_spec, int K, int N, cudaStream_t s);

// Brute-force candidate evaluation (dcb_brute.cu)
struct BruteArgs {
    DevParams p;
    int env;                 // which env of the batch
    long long first, count;  // candidates [first, first + count) in the reference's enumeration (brute_force.py:59-62)
    double *rewards;         // [count] central step reward of each candidate
};
cudaError_t dcb_launch_brute(const BruteArgs &a, cudaStream_t s);

cudaError_t dcb_launch_generate(const GenArgs &a, cudaStream_t s);
cudaError_t dcb_launch_population(const PopArgs &a, cudaStream_t s);
cudaError_t dcb_launch_iota_uid(int32_t *uid, int K, int N, cudaStream_t s);
cudaError_t dcb_launch_reset(const ResetArgs &a, cudaStream_t s);
cudaError_t dcb_launch_advance_skip(int K, int N, const int32_t *env_ids, int n_ids, const uint2 *mv,
                                    uint32_t *mv_skip, const uint32_t *pos_skip, cudaStream_t s);
cudaError_t dcb_launch_table_cursor(int K, int N, const int32_t *env_ids, int n_ids, uint2 *mv, uint32_t *mv_skip, int mode,
                                    cudaStream_t s);
cudaError_t dcb_launch_step(const StepArgs &a, int threads, int grid, size_t smem, cudaStream_t s);
cudaError_t dcb_step_set_smem_limit(int threads, int n_bs, size_t smem);
cudaError_t dcb_launch_wide(const StepArgs &a, int threads, int grid, size_t smem, cudaStream_t s);
cudaError_t dcb_wide_set_smem_limit(size_t smem);
// radio constants of the interference pass -> constant memory of the current device (dcb_wide.cu)
cudaError_t dcb_wide_upload_interference_constants(const double *pw10, double snr_c0, double snr_h);
size_t dcb_step_smem_bytes(int kind, int N, int M, int E, int var = 0);
int dcb_step_regs_per_thread(int threads, int n_bs);
