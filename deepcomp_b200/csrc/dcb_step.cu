// Fused env.step kernel: T consecutive steps of K independent env instances in one launch.
//
// Restates deepcomp/env/single_ue/base.py:413-466 (MobileEnv.step) with the observation / reward variants of
// deepcomp/env/multi_ue/central.py:143-152 and deepcomp/env/multi_ue/multi_agent.py:6-107.  All file:line
// citations are relative to /root/reference/deepcomp/.
//
// Mapping.  A CTA owns E consecutive envs; thread t owns UE slot t of the CTA's E*N UEs, which are contiguous in
// every [K][N] state slab, so state loads / reward stores are perfectly coalesced.  The per-UE state (position,
// waypoint, pause counter, connection bitmask, EWMA rate) stays in registers for all T steps.  Each thread walks
// its M UE x BS pairs serially; the per-BS reductions over UEs (connected count, sum of inverse rates, sum of
// priorities, arg-max rate, sum of utilities) go through a dense [E*N][M|1] fp64 matrix in shared memory that
// S lanes per (env, BS) pair column-sum and combine with warp shuffles, in a fixed order (deterministic results).
// The observation tile of the CTA is staged in shared memory and written out with coalesced 16-byte stores.
//
// Arithmetic.  Positions and every range decision are fp64 with the reference's operation order (no FMA
// contraction: the library is built with --fmad=false; the one FMA the reference has, inside np.linalg.norm, is
// explicit), so trajectories, connection masks and lost-connection counts are bit-exact.  SNR / rate / utility are
// fp64 through the table-driven log2 / exp2 of dcb_math.cuh (<= 1e-13 relative to the reference's libm chain); a
// UE closer than 1 m to a BS -- where the reference's `distance + EPSILON` matters -- takes the libm path.
#include <math_constants.h>

#include "dcb_internal.h"
#include "dcb_math.cuh"

namespace {

struct SmemLayout {
    int off_tab, off_a, off_b, off_sum_pre, off_sum_post, off_usum, off_umin, off_fues, off_futil, off_su, off_srb,
        off_smask, off_env_rew, off_env_sumu, off_bsx, off_bsy, off_vel, off_cnt_pre, off_arg_pre, off_cnt_post,
        off_arg_post, off_share;
    int total;
};

__host__ __device__ inline int align16(int x) { return (x + 15) & ~15; }

__host__ __device__ inline int obs_width(int kind, int M) { return kind == DCB_KIND_CENTRAL ? 2 * M + 1 : 4 * M + 1; }

// Row stride (in doubles) of the [E*N][M] matrices: odd, so that the 16 lanes of one 64-bit shared-memory access
// phase (consecutive UEs, same BS) hit 16 different bank pairs.
__host__ __device__ inline int row_stride(int M) { return M | 1; }

__host__ __device__ inline SmemLayout smem_layout(int kind, int N, int M, int E) {
    SmemLayout L;
    const int EN = E * N, EM = E * M;
    int o = 0;
    L.off_tab = o;      o += (int)sizeof(MathTables);                // 3 x 128 B: one bank row per table
    const int stage = EN * obs_width(kind, M) * 4;
    const int amat = EN * row_stride(M) * 8;
    L.off_a = o;        o += align16(stage > amat ? stage : amat);   // matrix A, later the obs staging tile
    L.off_b = o;        o += align16(amat);
    L.off_sum_pre = o;  o += align16(EM * 8);
    L.off_sum_post = o; o += align16(EM * 8);
    L.off_usum = o;     o += align16(EM * 8);
    L.off_umin = o;     o += align16(EM * 8);
    L.off_fues = o;     o += align16(EM * 8);
    L.off_futil = o;    o += align16(EM * 8);
    L.off_su = o;       o += align16(EN * 8);
    L.off_srb = o;      o += align16(EN * 8);
    L.off_smask = o;    o += align16(EN * 8);
    L.off_env_rew = o;  o += align16(E * 8);
    L.off_env_sumu = o; o += align16(E * 8);
    L.off_bsx = o;      o += align16(M * 8);
    L.off_bsy = o;      o += align16(M * 8);
    L.off_vel = o;      o += align16(N * 8);
    L.off_cnt_pre = o;  o += align16(EM * 4);
    L.off_arg_pre = o;  o += align16(EM * 4);
    L.off_cnt_post = o; o += align16(EM * 4);
    L.off_arg_post = o; o += align16(EM * 4);
    L.off_share = o;    o += align16(M * 4);
    L.total = o;
    return L;
}

// ------------------------------------------------------------------------------------------------ radio model
__device__ __forceinline__ double dist2(double ax, double ay, double bx, double by) {
    // shapely/GEOS Point.distance = sqrt(dx*dx + dy*dy) (station.py:124); the square is compared / rooted later
    const double dx = ax - bx, dy = ay - by;
    return dx * dx + dy * dy;
}

// station.py:110-127 verbatim with libm: used when the UE is within 1 m of the BS (distance + EPSILON matters)
__device__ __noinline__ double snr_of_d2_libm(double c1, double c2, double d2) {
    const double d = sqrt(d2);
    const double pl = c1 + c2 * log10(d + DCB_EPSILON);
    const double signal = pow(10.0, (DCB_TX_POWER - pl) / 10.0);
    return signal / DCB_NOISE;
}

// SNR = 10^((30 - c1 - c2 log10(d)) / 10) / 1e-9 = 2^(c0 - h log2(d^2)),  h = c2 / 20
__device__ __forceinline__ double snr_of_d2(const DevParams &p, const MathTables *tab, double d2) {
    if (d2 < 1.0) return snr_of_d2_libm(p.c1, p.c2, d2);
    return dcb_exp2(tab, fma(-p.snr_h, dcb_log2(tab, d2), p.snr_c0));
}

__device__ __forceinline__ double rate_unshared(const MathTables *tab, double snr) {
    return DCB_BW * dcb_log2_1p(tab, snr);   // station.py:129-138
}

__device__ __forceinline__ double log_utility(const MathTables *tab, double dr) {
    // env/util/utility.py:36-54: clip(10 log10(dr), -20, 20); dr <= 0.01 / >= 100 clip without evaluating the log
    if (dr <= 0.01) return DCB_MIN_UTILITY;
    if (dr >= 100.0) return DCB_MAX_UTILITY;
    const double u = 3.0102999566398119521 * dcb_log2(tab, dr);   // 10 log10(2) log2(dr)
    return fmin(fmax(u, DCB_MIN_UTILITY), DCB_MAX_UTILITY);
}

// value a connected link contributes to its BS's reduction, by sharing model (station.py:170-195)
__device__ __forceinline__ double link_value(int model, double r0, double ewma) {
    if (model == DCB_SHARE_RATE_FAIR) return 1.0 / r0;                               // :178
    if (model == DCB_SHARE_PROPORTIONAL_FAIR) return r0 / (ewma + DCB_EPSILON);      // :150 (alpha = beta = 1)
    return r0;                                                                       // resource-fair / max-cap
}

// ------------------------------------------------------------------------------------------------ reductions
// Column reduction of the dense [E*N][MS] matrix A: for every (env, BS) pair count the non-zero entries, sum them
// and (max-cap only) find the first arg-max.  S lanes per pair, fixed combination order.
__device__ __forceinline__ void reduce_links(const double *A, int N, int M, int MS, int n_env, int S, bool want_arg,
                                             int *cnt, double *sum, int *arg) {
    const int R = n_env * M;
    const int ppp = blockDim.x / S;
    const int seg = threadIdx.x & (S - 1);
    const int chunk = (N + S - 1) / S;
    for (int base = 0; base < R; base += ppp) {
        const int pair = base + threadIdx.x / S;
        const bool ok = pair < R;
        int c = 0, bi = 0x7fffffff;
        double s = 0.0, best = 0.0;
        if (ok) {
            const int le = pair / M, b = pair - le * M;
            const int i0 = seg * chunk;
            const int i1 = min(N, i0 + chunk);
            const double *col = A + (size_t)(le * N) * MS + b;
#pragma unroll 4
            for (int i = i0; i < i1; i++) {
                const double v = col[(size_t)i * MS];
                c += (v != 0.0);
                s += v;
                if (want_arg && v > best) { best = v; bi = i; }
            }
        }
        for (int off = S >> 1; off > 0; off >>= 1) {
            c += __shfl_xor_sync(0xffffffffu, c, off);
            s += __shfl_xor_sync(0xffffffffu, s, off);
            if (want_arg) {
                const double ob = __shfl_xor_sync(0xffffffffu, best, off);
                const int oi = __shfl_xor_sync(0xffffffffu, bi, off);
                if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
            }
        }
        if (ok && seg == 0) { cnt[pair] = c; sum[pair] = s; arg[pair] = bi; }
    }
}

// Per-BS total utility (station.py:63-69) -> usum, the two per-BS observation entries (variants.py:296-299,
// station.py:71-76) -> f_ues, f_util; optional masked min (station.py:78-83).
__device__ __forceinline__ void reduce_utility(const double *A, const unsigned long long *smask, const double *su,
                                               const int *cnt, int N, int M, int MS, int n_env, int S, bool want_min,
                                               double *usum, double *umin, double *f_ues, double *f_util) {
    const int R = n_env * M;
    const int ppp = blockDim.x / S;
    const int seg = threadIdx.x & (S - 1);
    const int chunk = (N + S - 1) / S;
    for (int base = 0; base < R; base += ppp) {
        const int pair = base + threadIdx.x / S;
        const bool ok = pair < R;
        double s = 0.0, mn = DCB_MAX_UTILITY;
        if (ok) {
            const int le = pair / M, b = pair - le * M;
            const int i0 = seg * chunk;
            const int i1 = min(N, i0 + chunk);
            const double *col = A + (size_t)(le * N) * MS + b;
#pragma unroll 4
            for (int i = i0; i < i1; i++) s += col[(size_t)i * MS];
            if (want_min)
                for (int i = i0; i < i1; i++)
                    if ((smask[le * N + i] >> b) & 1ull) mn = fmin(mn, su[le * N + i]);
        }
        for (int off = S >> 1; off > 0; off >>= 1) {
            s += __shfl_xor_sync(0xffffffffu, s, off);
            if (want_min) mn = fmin(mn, __shfl_xor_sync(0xffffffffu, mn, off));
        }
        if (ok && seg == 0) {
            const int c = cnt[pair];
            usum[pair] = s;
            umin[pair] = mn;
            f_ues[pair] = (double)c / (double)N;
            f_util[pair] = (c > 0 ? s / (double)c : 0.0) / DCB_MAX_UTILITY;
        }
    }
}

// Per-env reduction of a per-UE vector: mode 0 = sum, 2 = min (one warp per env)
__device__ __forceinline__ void reduce_env(const double *v, int N, int n_env, int mode, double *out) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const int chunk = (N + 31) / 32;
    for (int le = warp; le < n_env; le += nwarps) {
        const int i0 = lane * chunk, i1 = min(N, i0 + chunk);
        double s = mode == 2 ? CUDART_INF : 0.0;
        for (int i = i0; i < i1; i++) s = mode == 2 ? fmin(s, v[le * N + i]) : s + v[le * N + i];
        for (int off = 16; off > 0; off >>= 1) {
            const double o = __shfl_xor_sync(0xffffffffu, s, off);
            s = mode == 2 ? fmin(s, o) : s + o;
        }
        if (lane == 0) out[le] = s;
    }
}

// ------------------------------------------------------------------------------------------------ the kernel
template <int MAXT>
__global__ void __launch_bounds__(MAXT) dcb_step_kernel(const StepArgs a) {
    extern __shared__ __align__(128) unsigned char smem[];
    const DevParams &p = a.p;
    const int N = p.N, M = p.M, E = p.E, S = p.S;
    const int MS = row_stride(M);
    const SmemLayout L = smem_layout(p.kind, N, M, E);
    MathTables *tab = reinterpret_cast<MathTables *>(smem + L.off_tab);
    double *A = reinterpret_cast<double *>(smem + L.off_a);
    float *stage = reinterpret_cast<float *>(smem + L.off_a);
    double *B = reinterpret_cast<double *>(smem + L.off_b);
    int *cnt_pre = reinterpret_cast<int *>(smem + L.off_cnt_pre);
    double *sum_pre = reinterpret_cast<double *>(smem + L.off_sum_pre);
    int *arg_pre = reinterpret_cast<int *>(smem + L.off_arg_pre);
    int *cnt_post = reinterpret_cast<int *>(smem + L.off_cnt_post);
    double *sum_post = reinterpret_cast<double *>(smem + L.off_sum_post);
    int *arg_post = reinterpret_cast<int *>(smem + L.off_arg_post);
    double *usum = reinterpret_cast<double *>(smem + L.off_usum);
    double *umin = reinterpret_cast<double *>(smem + L.off_umin);
    double *f_ues = reinterpret_cast<double *>(smem + L.off_fues);
    double *f_util = reinterpret_cast<double *>(smem + L.off_futil);
    double *su = reinterpret_cast<double *>(smem + L.off_su);
    double *srb = reinterpret_cast<double *>(smem + L.off_srb);
    unsigned long long *smask = reinterpret_cast<unsigned long long *>(smem + L.off_smask);
    double *env_rew = reinterpret_cast<double *>(smem + L.off_env_rew);
    double *env_sumu = reinterpret_cast<double *>(smem + L.off_env_sumu);
    double *bsx = reinterpret_cast<double *>(smem + L.off_bsx);
    double *bsy = reinterpret_cast<double *>(smem + L.off_bsy);
    int *share = reinterpret_cast<int *>(smem + L.off_share);
    double *velspec = reinterpret_cast<double *>(smem + L.off_vel);

    const int t = threadIdx.x;
    const int env0 = blockIdx.x * E;
    const int n_env = min(E, p.K - env0);
    const bool valid = t < n_env * N;
    const int le = valid ? t / N : 0;
    const int i = valid ? t - le * N : 0;
    const int k = env0 + le;
    const long long u = (long long)k * N + i;
    const bool central = p.kind == DCB_KIND_CENTRAL;
    const int OW = obs_width(p.kind, M);

    dcb_math_init(tab, t);
    for (int b = t; b < M; b += blockDim.x) {
        bsx[b] = p.bs_xy[2 * b];
        bsy[b] = p.bs_xy[2 * b + 1];
        share[b] = p.sharing[b];
    }
    for (int j = t; j < N; j += blockDim.x) velspec[j] = p.vel_spec[j];

    // ---- per-UE state -> registers
    double x = 0, y = 0, ewma = 0;
    unsigned long long mask = 0;
    unsigned wxy = 0, vpt = 0;
    int tk = 0;
    if (valid) {
        const double2 ps = p.pos[u];
        x = ps.x; y = ps.y;
        const uint2 mv = p.mv[u];
        wxy = mv.x; vpt = mv.y;
        mask = p.mask[u];
        ewma = p.ewma[u];
        tk = p.time[k];
    }
    __syncthreads();
    const double vfix = valid ? velspec[i] : 0.0;
    double *Arow = A + (size_t)t * MS;
    double *Brow = B + (size_t)t * MS;
    const int T = a.T;
    const int n_iter = T > 0 ? T : 1;

    for (int step = 0; step < n_iter; step++) {
        const bool last = step == n_iter - 1;
        double rb = 0.0;      // reward before the move (base.py:446)
        int lost = 0;
        if (T > 0) {
            // ---- episode boundary: MobileEnv.reset before the next step (base.py:169-189)
            if (valid && p.auto_reset && tk >= p.episode_length) {
                const double2 ps = p.init_pos[u];
                x = ps.x; y = ps.y;
                const uint32_t e = p.table[u * p.D];
                wxy = (e & 0x3fffu) | (((e >> 14) & 0x3fffu) << 16);
                vpt = (e >> 28) | (1u << 16);
                mask = 0ull; ewma = 0.0; tk = 0;
            }
            // ---- apply_ue_actions (base.py:247-282) -> User.connect_to_bs(disconnect=True) (user.py:190-229)
            if (valid) {
                const int act = a.actions[(size_t)step * p.K * N + u];
                if (act < 0 || act > M) {
                    atomicOr(p.err, DCB_ERRBIT_ACTION);
                } else if (act > 0) {
                    const int b = act - 1;
                    const unsigned long long bit = 1ull << b;
                    if (mask & bit) mask &= ~bit;
                    else if (dist2(bsx[b], bsy[b], x, y) <= p.thr_d2) mask |= bit;     // can_connect, station.py:222-226
                }
                // ---- link values for update_ue_drs_rewards (base.py:315-335) at the pre-move position
                for (int b = 0; b < M; b++) Arow[b] = 0.0;
                for (unsigned long long m = mask; m; m &= m - 1) {
                    const int b = __ffsll((long long)m) - 1;
                    const double r0 = rate_unshared(tab, snr_of_d2(p, tab, dist2(bsx[b], bsy[b], x, y)));
                    Arow[b] = link_value(share[b], r0, ewma);
                    Brow[b] = r0;
                }
            }
            __syncthreads();
            reduce_links(A, N, M, MS, n_env, S, p.has_maxcap, cnt_pre, sum_pre, arg_pre);
            __syncthreads();
            if (valid) {
                // ---- Basestation.data_rate_shared (station.py:152-202) per connected link; ue.bs_dr cache in Brow
                double dr = 0.0;
                for (unsigned long long m = mask; m; m &= m - 1) {
                    const int b = __ffsll((long long)m) - 1;
                    const int pr = le * M + b;
                    const int model = share[b];
                    const double r0 = Brow[b];
                    double r;
                    if (model == DCB_SHARE_RESOURCE_FAIR) r = r0 / (double)cnt_pre[pr];
                    else if (model == DCB_SHARE_RATE_FAIR) r = 1.0 / sum_pre[pr];
                    else if (model == DCB_SHARE_MAX_CAP) r = (arg_pre[pr] == i) ? r0 : 0.0;
                    else r = Arow[b] / (sum_pre[pr] + DCB_EPSILON) * r0;
                    Brow[b] = r;
                    dr += r;                                                           // user.py:64-69
                }
                // ---- calc_reward (base.py:158-167), penalties are identically 0 (base.py:257)
                rb = log_utility(tab, dr) / DCB_MAX_UTILITY;
                // ---- User.move (user.py:159-173) -> RandomWaypoint.step (movement.py:158-181)
                double wx = (double)(wxy & 0xffffu), wy = (double)(wxy >> 16);
                unsigned pause = (vpt >> 8) & 0xffu;
                bool moving = true;
                if (x == wx && y == wy) pause |= 0x80u;                                // movement.py:169-170
                if (pause & 0x80u) {
                    if ((int)(pause & 0x7fu) < p.pause_duration) {                     // movement.py:174-176
                        pause++;
                        moving = false;
                    } else {                                                           // movement.py:177 -> reset()
                        unsigned tidx = vpt >> 16;
                        if ((int)tidx >= p.D) {
                            atomicOr(p.err, DCB_ERRBIT_TABLE);
                            tidx = p.D - 1;
                        }
                        const uint32_t e = p.table[u * p.D + tidx];
                        wxy = (e & 0x3fffu) | (((e >> 14) & 0x3fffu) << 16);
                        vpt = (e >> 28) | ((tidx + 1) << 16);
                        wx = (double)(wxy & 0xffffu); wy = (double)(wxy >> 16);
                        pause = 0;
                    }
                }
                vpt = (vpt & 0xffff00ffu) | (pause << 8);
                if (moving) {
                    // movement.py:132-156
                    const double vel = vfix >= 0.0 ? vfix : (double)(vpt & 0xffu);
                    if (sqrt(dist2(x, y, wx, wy)) <= vel) {
                        x = wx; y = wy;
                    } else {
                        const double vx = wx - x, vy = wy - y;
                        const double norm = sqrt(fma(vy, vy, vx * vx));   // np.linalg.norm -> FMA-accumulating ddot
                        x = x + vel * (vx / norm);
                        y = y + vel * (vy / norm);
                    }
                }
                // ---- check_bs_connection (user.py:175-188) + update_ewma_dr (user.py:148-157)
                double keep = 0.0;
                for (unsigned long long m = mask; m; m &= m - 1) {
                    const int b = __ffsll((long long)m) - 1;
                    if (dist2(bsx[b], bsy[b], x, y) <= p.thr_d2) keep += Brow[b];
                    else { mask &= ~(1ull << b); lost++; }
                }
                ewma = 0.9 * keep + (1 - 0.9) * ewma;
                tk += 1;                                                               // base.py:454
            }
        }
        // =========================== observe the (post-move) state ===========================
        double mx = 0.0;
        unsigned long long inrange = 0ull;
        if (valid) {
            // SNR of every pair at the current position (variants.py:278) -> Brow; in-range set (multi_agent.py:60);
            // link values at the new position for update_ue_drs_rewards(update_only=True) (base.py:451) -> Arow
#pragma unroll 2
            for (int b = 0; b < M; b++) {
                const double d2 = dist2(bsx[b], bsy[b], x, y);
                const double s = snr_of_d2(p, tab, d2);
                Brow[b] = s;
                mx = fmax(mx, s);
                if (d2 <= p.thr_d2) inrange |= 1ull << b;
                double v = 0.0;
                if ((mask >> b) & 1ull) v = link_value(share[b], rate_unshared(tab, s), ewma);
                Arow[b] = v;
            }
        }
        __syncthreads();
        reduce_links(A, N, M, MS, n_env, S, p.has_maxcap, cnt_post, sum_post, arg_post);
        __syncthreads();
        double dr = 0.0, util = 0.0;
        if (valid) {
            for (unsigned long long m = mask; m; m &= m - 1) {
                const int b = __ffsll((long long)m) - 1;
                const int pr = le * M + b;
                const int model = share[b];
                const double v = Arow[b];
                double r;
                if (model == DCB_SHARE_RESOURCE_FAIR) r = v / (double)cnt_post[pr];
                else if (model == DCB_SHARE_RATE_FAIR) r = 1.0 / sum_post[pr];
                else if (model == DCB_SHARE_MAX_CAP) r = (arg_post[pr] == i) ? v : 0.0;
                else r = v / (sum_post[pr] + DCB_EPSILON) * rate_unshared(tab, Brow[b]);
                if (last && a.out.dbg_link_rate) a.out.dbg_link_rate[u * M + b] = r;
                dr += r;
            }
            util = log_utility(tab, dr);                                               // user.py:76-92
            su[t] = util;
            srb[t] = rb;
            smask[t] = mask;
            if (!central)
                for (int b = 0; b < M; b++) Arow[b] = ((mask >> b) & 1ull) ? util : 0.0;
        }
        __syncthreads();
        if (!central)
            reduce_utility(A, smask, su, cnt_post, N, M, MS, n_env, S, p.reward == DCB_REWARD_MIN, usum, umin, f_ues,
                           f_util);
        if (central && T > 0) reduce_env(srb, N, n_env, p.reward == DCB_REWARD_MIN ? 2 : 0, env_rew);
        if (a.out.sum_utility || a.out.dbg_sum_utility) reduce_env(su, N, n_env, 0, env_sumu);
        __syncthreads();
        // ---- observation row -> staging tile (A is dead now), rewards and per-UE outputs -> global
        if (valid) {
            double *dobs = (last && a.out.dbg_obs) ? a.out.dbg_obs : nullptr;
            const double un = util / DCB_MAX_UTILITY;                                  // variants.py:287
            const double inv_mx = mx == 0.0 ? 0.0 : 1.0 / mx;                          // variants.py:279-284
            if (central) {
                // central.py:31-57: [connected(N*M) | dr(N*M) | utility(N)] per env
                float *row = stage + (size_t)le * (2 * N * M + N);
                double *drow = dobs ? dobs + (size_t)k * (2 * N * M + N) : nullptr;
                for (int b = 0; b < M; b++) {
                    const double c = (double)((mask >> b) & 1ull);
                    const double r = Brow[b] * inv_mx;
                    row[i * M + b] = (float)c;
                    row[N * M + i * M + b] = (float)r;
                    if (drow) { drow[i * M + b] = c; drow[N * M + i * M + b] = r; }
                }
                row[2 * N * M + i] = (float)un;
                if (drow) drow[2 * N * M + i] = un;
            } else {
                // variants.py:271-303: [connected(M) | dr(M) | ues_at_bs(M) | util_at_bs(M) | utility(1)] per UE
                float *row = stage + (size_t)t * OW;
                double *drow = dobs ? dobs + (size_t)u * OW : nullptr;
                for (int b = 0; b < M; b++) {
                    const int pr = le * M + b;
                    const double c = (double)((mask >> b) & 1ull);
                    const double r = Brow[b] * inv_mx;
                    const double ab = f_ues[pr], ub = f_util[pr];
                    row[b] = (float)c; row[M + b] = (float)r; row[2 * M + b] = (float)ab; row[3 * M + b] = (float)ub;
                    if (drow) { drow[b] = c; drow[M + b] = r; drow[2 * M + b] = ab; drow[3 * M + b] = ub; }
                }
                row[4 * M] = (float)un;
                if (drow) drow[4 * M] = un;
            }
            if (last && a.out.dbg_snr)
                for (int b = 0; b < M; b++) a.out.dbg_snr[u * M + b] = Brow[b];
            if (a.out.curr_dr) a.out.curr_dr[(size_t)step * a.out.curr_dr_stride + u] = (float)dr;
            if (a.out.utility) a.out.utility[(size_t)step * a.out.utility_stride + u] = (float)util;
            if (last && a.out.dbg_curr_dr) a.out.dbg_curr_dr[u] = dr;
            if (last && a.out.dbg_utility) a.out.dbg_utility[u] = util;
            if (i == 0) {
                if (a.out.sum_utility) a.out.sum_utility[(size_t)step * a.out.sum_utility_stride + k] = (float)env_sumu[le];
                if (last && a.out.dbg_sum_utility) a.out.dbg_sum_utility[k] = env_sumu[le];
            }
            if (T > 0) {
                if (a.out.lost_conn) a.out.lost_conn[(size_t)step * a.out.lost_conn_stride + u] = (uint8_t)lost;
                if (central) {
                    if (i == 0) {
                        // central.py:65-73 over the PRE-move rewards
                        double r = env_rew[le];
                        if (p.reward == DCB_REWARD_AVG) r = r / (double)N;
                        if (a.out.reward) a.out.reward[(size_t)step * a.out.reward_stride + k] = (float)r;
                        if (last && a.out.dbg_reward) a.out.dbg_reward[k] = r;
                    }
                } else {
                    // multi_agent.py:39-95 on the POST-move state
                    double agg = util;
                    if (inrange) {
                        if (p.reward == DCB_REWARD_AVG) {
                            int nn = 0;
                            double tot = 0.0;
                            for (unsigned long long m = inrange; m; m &= m - 1) {
                                const int b = __ffsll((long long)m) - 1;
                                nn += cnt_post[le * M + b];
                                tot += usum[le * M + b];
                            }
                            if (nn > 0) agg = mask == 0ull ? (tot + util) / (double)(nn + 1) : tot / (double)nn;
                        } else if (p.reward == DCB_REWARD_SUM) {
                            // user.py:238-244: UEs sharing any BS with this UE; their PRE-move rewards
                            agg = 0.0;
                            for (int j = 0; j < N; j++)
                                if (smask[le * N + j] & mask) agg += srb[le * N + j];
                        } else {
                            for (unsigned long long m = inrange; m; m &= m - 1) {
                                const int b = __ffsll((long long)m) - 1;
                                agg = fmin(agg, umin[le * M + b]);
                            }
                        }
                    }
                    if (a.out.reward) a.out.reward[(size_t)step * a.out.reward_stride + u] = (float)agg;
                    if (last && a.out.dbg_reward) a.out.dbg_reward[u] = agg;
                }
            }
        }
        __syncthreads();
        // ---- staging tile -> global observation buffer (contiguous span of this CTA, coalesced)
        if (a.out.obs) {
            const size_t per_env = central ? (size_t)(2 * N * M + N) : (size_t)N * OW;
            float *dst = a.out.obs + (size_t)step * a.out.obs_stride + (size_t)env0 * per_env;
            const int n = (int)(per_env * n_env);
            if ((((size_t)dst) & 15) == 0) {
                const int n4 = n >> 2;
                const float4 *s4 = reinterpret_cast<const float4 *>(stage);
                float4 *d4 = reinterpret_cast<float4 *>(dst);
                for (int j = t; j < n4; j += blockDim.x) __stcs(d4 + j, s4[j]);
                for (int j = (n4 << 2) + t; j < n; j += blockDim.x) dst[j] = stage[j];
            } else {
                for (int j = t; j < n; j += blockDim.x) dst[j] = stage[j];
            }
        }
        __syncthreads();
    }

    // ---- registers -> state slabs
    if (valid && T > 0) {
        p.pos[u] = make_double2(x, y);
        p.mv[u] = make_uint2(wxy, vpt);
        p.mask[u] = mask;
        p.ewma[u] = ewma;
        if (i == 0) p.time[k] = tk;
    }
}

}  // namespace

size_t dcb_step_smem_bytes(int kind, int N, int M, int E) { return (size_t)smem_layout(kind, N, M, E).total; }

cudaError_t dcb_step_set_smem_limit(int threads, size_t smem) {
    if (threads <= 256)
        return cudaFuncSetAttribute(dcb_step_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (threads <= 512)
        return cudaFuncSetAttribute(dcb_step_kernel<512>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    return cudaFuncSetAttribute(dcb_step_kernel<1024>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
}

cudaError_t dcb_launch_step(const StepArgs &a, int threads, int grid, size_t smem, cudaStream_t s) {
    if (threads <= 256) dcb_step_kernel<256><<<grid, threads, smem, s>>>(a);
    else if (threads <= 512) dcb_step_kernel<512><<<grid, threads, smem, s>>>(a);
    else dcb_step_kernel<1024><<<grid, threads, smem, s>>>(a);
    return cudaGetLastError();
}
