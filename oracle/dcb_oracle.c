/*
 * TEST INFRASTRUCTURE ONLY -- fast C restatement of the reference env.step hot path.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load this library; the product
 * (deepcomp_b200/) never links or calls it.
 *
 * Parity status: pinned.  tests/test_oracle_c.py checks this restatement against (i) the Python restatement
 * oracle/deepcomp_oracle.py (itself bit-identical to the live reference, tests/test_oracle_vs_reference.py) and
 * (ii) the golden traces of the reference itself in tests/golden/.  Integer state (masks, lost-connection
 * counts, RNG draws, pause counters) and positions are bit-exact; floating-point aggregates use a canonical
 * summation order (UE index within a BS, BS index within a UE) instead of the reference's connection order, which
 * moves results by O(1 ulp) (SURVEY.md section 7 hard part d).
 *
 * All file:line citations are relative to /root/reference/deepcomp/.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* util/constants.py:28,40-41; env/entities/station.py:10,26-30 */
#define EPSILON 1e-16
#define MIN_UTILITY (-20.0)
#define MAX_UTILITY 20.0
#define SNR_THRESHOLD 2e-8
#define MAX_SNR_THRESHOLD 7e-6   /* MaxNormEnv.MAX_SNR_THRESHOLD, single_ue/variants.py:311 */
#define BW 9e6
#define NOISE 1e-9
#define TX_POWER 30.0

enum { SH_RESOURCE_FAIR = 0, SH_RATE_FAIR = 1, SH_PROP_FAIR = 2, SH_MAX_CAP = 3 };
enum { KIND_CENTRAL = 0, KIND_MULTI = 1 };
enum { RW_AVG = 0, RW_SUM = 1, RW_MIN = 2 };

/* ------------------------------------------------------------------------------------------------ MT19937
 * CPython's random.Random: _randommodule.c (init_genrand / init_by_array / genrand_uint32) and Lib/random.py
 * (randint -> randrange -> _randbelow_with_getrandbits).  The reference draws through random.Random in
 * entities/user.py:38,103-107 and util/movement.py:14,113-127.
 */
typedef struct { uint32_t mt[624]; int idx; } mt_t;

static void mt_init_genrand(mt_t *g, uint32_t s) {
    g->mt[0] = s;
    for (int i = 1; i < 624; i++) g->mt[i] = 1812433253u * (g->mt[i - 1] ^ (g->mt[i - 1] >> 30)) + (uint32_t)i;
    g->idx = 624;
}

static void mt_init_by_array(mt_t *g, const uint32_t *key, int len) {
    mt_init_genrand(g, 19650218u);
    int i = 1, j = 0;
    int k = 624 > len ? 624 : len;
    for (; k; k--) {
        g->mt[i] = (g->mt[i] ^ ((g->mt[i - 1] ^ (g->mt[i - 1] >> 30)) * 1664525u)) + key[j] + (uint32_t)j;
        i++; j++;
        if (i >= 624) { g->mt[0] = g->mt[623]; i = 1; }
        if (j >= len) j = 0;
    }
    for (k = 623; k; k--) {
        g->mt[i] = (g->mt[i] ^ ((g->mt[i - 1] ^ (g->mt[i - 1] >> 30)) * 1566083941u)) - (uint32_t)i;
        i++;
        if (i >= 624) { g->mt[0] = g->mt[623]; i = 1; }
    }
    g->mt[0] = 0x80000000u;
}

/* random.seed(int): key = 32-bit little-endian words of abs(seed), at least one word */
static void mt_seed_int(mt_t *g, long long seed) {
    unsigned long long a = seed < 0 ? (unsigned long long)(-(seed + 1)) + 1ull : (unsigned long long)seed;
    uint32_t key[2] = { (uint32_t)(a & 0xffffffffu), (uint32_t)(a >> 32) };
    mt_init_by_array(g, key, key[1] ? 2 : 1);
}

static uint32_t mt_genrand(mt_t *g) {
    if (g->idx >= 624) {
        uint32_t *mt = g->mt;
        int kk;
        for (kk = 0; kk < 624 - 397; kk++) {
            uint32_t y = (mt[kk] & 0x80000000u) | (mt[kk + 1] & 0x7fffffffu);
            mt[kk] = mt[kk + 397] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
        }
        for (; kk < 623; kk++) {
            uint32_t y = (mt[kk] & 0x80000000u) | (mt[kk + 1] & 0x7fffffffu);
            mt[kk] = mt[kk + (397 - 624)] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
        }
        uint32_t y = (mt[623] & 0x80000000u) | (mt[0] & 0x7fffffffu);
        mt[623] = mt[396] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
        g->idx = 0;
    }
    uint32_t y = g->mt[g->idx++];
    y ^= (y >> 11);
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= (y >> 18);
    return y;
}

/* Lib/random.py _randbelow_with_getrandbits, n < 2**32 */
static uint32_t mt_randbelow(mt_t *g, uint32_t n) {
    int k = 32 - __builtin_clz(n);
    uint32_t r = mt_genrand(g) >> (32 - k);
    while (r >= n) r = mt_genrand(g) >> (32 - k);
    return r;
}

static int mt_randint(mt_t *g, int a, int b) { return a + (int)mt_randbelow(g, (uint32_t)(b - a + 1)); }

/* exported for the RNG known-answer tests */
void orc_rng_draws(long long seed, int n_raw, uint32_t *raw, int n_int, const int *lo, const int *hi, int *out) {
    mt_t g;
    mt_seed_int(&g, seed);
    for (int i = 0; i < n_raw; i++) raw[i] = mt_genrand(&g);
    mt_seed_int(&g, seed);
    for (int i = 0; i < n_int; i++) out[i] = mt_randint(&g, lo[i], hi[i]);
}

/* ------------------------------------------------------------------------------------------------ env */
typedef struct orc_env {
    int kind, n_ue, n_bs, width, height, reward_agg, rand_episodes, pause_duration, border_buffer, has_seed;
    int interference;     /* EXTENSION (not in the reference, which is SNR only: station.py:122-127, docs/model.md:15-19):
                             every use of the SNR sees SINR_b = snr_b / (1 + sum_{b' != b} snr_b') instead */
    int util_step;        /* User.util_func: 0 = 'log' (utility.py:36-54), 1 = 'step' (utility.py:23-33) */
    double dr_req;        /* User.dr_req (user.py:17-30) */
    int obs_maxnorm;      /* observation 'dr': 0 = RelNormEnv (variants.py:276-284), 1 = MaxNormEnv (variants.py:308-332) */
    long long seed;
    int time;
    double total_utility;
    double const1, const2;
    double *bs_xy;        /* [M][2] */
    int *sharing;         /* [M] */
    double *vel_spec;     /* [N]: -1 slow, -2 fast, >= 0 fixed number */
    double *init_xy;      /* [N][2], NaN = 'random' */
    mt_t *rng, *mrng;     /* [N] each: User.rng (user.py:38) and movement.rng (movement.py:14) */
    /* state */
    double *pos;          /* [N][2] */
    double *wp;           /* [N][2] */
    double *vel;          /* [N] */
    int *pausing, *curr_pause;
    uint8_t *mask;        /* [N][M] */
    double *ewma;         /* [N] */
    /* per-step results */
    double *link_rate;    /* [N][M] cached rate of connected links (ue.bs_dr values) */
    double *snr;          /* [N][M] at current positions */
    double *curr_dr, *utility, *reward_before;   /* [N] */
    int *lost_conn;       /* [N] */
    double *obs;          /* central: 2NM+N ; multi: N*(4M+1) */
    double *reward;       /* central: 1 ; multi: N */
    double sum_utility;
    /* scratch */
    double *bs_cnt, *bs_inv, *bs_prio, *bs_util_sum, *bs_util_min;
    int *bs_argmax;
} orc_env;

static double snr_of_distance(const orc_env *e, double d) {
    /* station.py:110-127 */
    double pl = e->const1 + e->const2 * log10(d + EPSILON);
    double signal = pow(10.0, (TX_POWER - pl) / 10.0);
    return signal / NOISE;
}

static double dist(double ax, double ay, double bx, double by) {
    /* shapely/GEOS point distance: sqrt(dx*dx + dy*dy); built with -ffp-contract=off so no FMA sneaks in */
    double dx = ax - bx, dy = ay - by;
    return sqrt(dx * dx + dy * dy);
}

static double log_utility(double dr) {
    /* env/util/utility.py:36-54 */
    if (dr == 0) return MIN_UTILITY;
    double u = 10.0 * log10(dr);
    return u < MIN_UTILITY ? MIN_UTILITY : (u > MAX_UTILITY ? MAX_UTILITY : u);
}

int orc_obs_size(const orc_env *e) {
    return e->kind == KIND_CENTRAL ? 2 * e->n_ue * e->n_bs + e->n_ue : e->n_ue * (4 * e->n_bs + 1);
}
int orc_reward_size(const orc_env *e) { return e->kind == KIND_CENTRAL ? 1 : e->n_ue; }

/* signal quality of UE i at its current position towards every BS: SNR (station.py:122-127) or, with the interference
 * extension, SINR = P_b / (noise + sum of the other BS' received powers) = snr_b / (1 + sum_{b' != b} snr_b') */
static void ue_quality(const orc_env *e, int i, double *out) {
    const int M = e->n_bs;
    for (int b = 0; b < M; b++)
        out[b] = snr_of_distance(e, dist(e->bs_xy[2 * b], e->bs_xy[2 * b + 1], e->pos[2 * i], e->pos[2 * i + 1]));
    if (e->interference) {
        double snr[64];
        memcpy(snr, out, sizeof(double) * M);
        for (int b = 0; b < M; b++) {
            double others = 0;
            for (int c = 0; c < M; c++)
                if (c != b) others += snr[c];
            out[b] = snr[b] / (1.0 + others);
        }
    }
}

static void compute_snr(orc_env *e) {
    for (int i = 0; i < e->n_ue; i++) ue_quality(e, i, e->snr + (size_t)i * e->n_bs);
}

/* station.py:129-220 over all connected links, user.py:143-146, 64-92; needs e->snr at the current positions */
static void update_rates(orc_env *e) {
    const int N = e->n_ue, M = e->n_bs;
    for (int b = 0; b < M; b++) {
        double cnt = 0, inv = 0, prio = 0, best = -1;
        int arg = -1;
        for (int i = 0; i < N; i++) {
            if (!e->mask[i * M + b]) continue;
            double r0 = BW * log2(1.0 + e->snr[i * M + b]);
            cnt += 1;
            inv += 1.0 / r0;                                       /* station.py:178 */
            prio += r0 / (e->ewma[i] + EPSILON);                   /* station.py:150,194 */
            if (r0 > best) { best = r0; arg = i; }                 /* station.py:184: first argmax */
        }
        e->bs_cnt[b] = cnt; e->bs_inv[b] = inv; e->bs_prio[b] = prio; e->bs_argmax[b] = arg;
    }
    for (int i = 0; i < N; i++) {
        double total = 0;
        for (int b = 0; b < M; b++) {
            double r = 0;
            if (e->mask[i * M + b]) {
                double s = e->snr[i * M + b];
                if (s > SNR_THRESHOLD) {                           /* station.py:212 */
                    double r0 = BW * log2(1.0 + s);
                    switch (e->sharing[b]) {
                    case SH_RESOURCE_FAIR: r = r0 / e->bs_cnt[b]; break;                       /* :173 */
                    case SH_RATE_FAIR: r = 1.0 / e->bs_inv[b]; break;                          /* :180 */
                    case SH_MAX_CAP: r = (e->bs_argmax[b] == i) ? r0 : 0.0; break;             /* :184-187 */
                    case SH_PROP_FAIR:                                                         /* :194-195 */
                        r = (r0 / (e->ewma[i] + EPSILON)) / (e->bs_prio[b] + EPSILON) * r0; break;
                    }
                }
                total += r;
            }
            e->link_rate[i * M + b] = r;
        }
        e->curr_dr[i] = total;
        /* user.py:81-92: 'log' or 'step' (MAX_UTILITY at or above the required rate, else MIN_UTILITY) */
        e->utility[i] = e->util_step ? (total >= e->dr_req ? MAX_UTILITY : MIN_UTILITY) : log_utility(total);
    }
}

static void movement_reset(orc_env *e, int i) {
    /* util/movement.py:110-130 */
    double vs = e->vel_spec[i];
    if (vs == -1.0) e->vel[i] = mt_randint(&e->mrng[i], 1, 3);
    else if (vs == -2.0) e->vel[i] = mt_randint(&e->mrng[i], 5, 10);
    else e->vel[i] = vs;
    e->wp[2 * i] = mt_randint(&e->mrng[i], e->border_buffer, e->width - e->border_buffer);
    e->wp[2 * i + 1] = mt_randint(&e->mrng[i], e->border_buffer, e->height - e->border_buffer);
    e->pausing[i] = 0;
    e->curr_pause[i] = 0;
}

static void movement_step(orc_env *e, int i) {
    /* util/movement.py:132-181 */
    double x = e->pos[2 * i], y = e->pos[2 * i + 1];
    if (x == e->wp[2 * i] && y == e->wp[2 * i + 1]) e->pausing[i] = 1;
    if (e->pausing[i]) {
        if (e->curr_pause[i] < e->pause_duration) { e->curr_pause[i]++; return; }
        movement_reset(e, i);
    }
    double wx = e->wp[2 * i], wy = e->wp[2 * i + 1];
    if (dist(x, y, wx, wy) <= e->vel[i]) { e->pos[2 * i] = wx; e->pos[2 * i + 1] = wy; return; }
    double vx = wx - x, vy = wy - y;
    /* np.linalg.norm -> OpenBLAS ddot accumulates with FMA: sqrt(fma(vy, vy, vx*vx)) (oracle/deepcomp_oracle.py:_norm2) */
    double norm = sqrt(fma(vy, vy, vx * vx));
    e->pos[2 * i] = x + e->vel[i] * (vx / norm);
    e->pos[2 * i + 1] = y + e->vel[i] * (vy / norm);
}

static void build_obs_reward(orc_env *e) {
    const int N = e->n_ue, M = e->n_bs;
    /* per-BS aggregates of the post-move state: station.py:63-83 */
    for (int b = 0; b < M; b++) {
        double cnt = 0, sum = 0, mn = MAX_UTILITY;
        int any = 0;
        for (int i = 0; i < N; i++)
            if (e->mask[i * M + b]) {
                cnt += 1; sum += e->utility[i];
                if (!any || e->utility[i] < mn) mn = e->utility[i];
                any = 1;
            }
        e->bs_cnt[b] = cnt; e->bs_util_sum[b] = sum; e->bs_util_min[b] = mn;
    }
    for (int i = 0; i < N; i++) {
        /* single_ue/variants.py:271-303 */
        double mx = 0;
        for (int b = 0; b < M; b++) if (e->snr[i * M + b] > mx) mx = e->snr[i * M + b];
        for (int b = 0; b < M; b++) {
            double conn = e->mask[i * M + b];
            double dr = mx == 0 ? 0 : e->snr[i * M + b] / mx;
            if (e->obs_maxnorm) {
                /* MaxNormEnv.get_ue_obs (variants.py:322-330): cap, subtract the required SNR, normalise */
                double sn = e->snr[i * M + b] < MAX_SNR_THRESHOLD ? e->snr[i * M + b] : MAX_SNR_THRESHOLD;
                dr = (sn - SNR_THRESHOLD) / (MAX_SNR_THRESHOLD - SNR_THRESHOLD);
            }
            if (e->kind == KIND_CENTRAL) {
                e->obs[i * M + b] = conn;
                e->obs[N * M + i * M + b] = dr;
            } else {
                double *o = e->obs + (size_t)i * (4 * M + 1);
                o[b] = conn;
                o[M + b] = dr;
                o[2 * M + b] = e->bs_cnt[b] / N;
                o[3 * M + b] = (e->bs_cnt[b] > 0 ? e->bs_util_sum[b] / e->bs_cnt[b] : 0.0) / MAX_UTILITY;
            }
        }
        if (e->kind == KIND_CENTRAL) e->obs[2 * N * M + i] = e->utility[i] / MAX_UTILITY;
        else e->obs[(size_t)i * (4 * M + 1) + 4 * M] = e->utility[i] / MAX_UTILITY;
    }
    if (e->kind == KIND_CENTRAL) {
        /* multi_ue/central.py:65-73 over the PRE-move rewards */
        double r;
        if (e->reward_agg == RW_AVG) { r = 0; for (int i = 0; i < N; i++) r += e->reward_before[i]; r /= N; }
        else if (e->reward_agg == RW_SUM) { r = 0; for (int i = 0; i < N; i++) r += e->reward_before[i]; }
        else { r = e->reward_before[0]; for (int i = 1; i < N; i++) if (e->reward_before[i] < r) r = e->reward_before[i]; }
        e->reward[0] = r;
        return;
    }
    /* multi_ue/multi_agent.py:39-95 on the POST-move state */
    for (int i = 0; i < N; i++) {
        double agg = e->utility[i];
        int n_in_range = 0, n_conn = 0;
        double num_neigh = 0, total = 0, mn = e->utility[i];
        for (int b = 0; b < M; b++) {
            if (e->mask[i * M + b]) n_conn++;
            if (e->snr[i * M + b] > SNR_THRESHOLD) {
                n_in_range++;
                num_neigh += e->bs_cnt[b];
                total += e->bs_util_sum[b];
                if (e->bs_util_min[b] < mn) mn = e->bs_util_min[b];
            }
        }
        if (n_in_range > 0) {
            if (e->reward_agg == RW_AVG) {
                if (num_neigh > 0) agg = n_conn == 0 ? (total + e->utility[i]) / (num_neigh + 1) : total / num_neigh;
            } else if (e->reward_agg == RW_SUM) {
                /* user.py:238-244: union of UEs at any BS this UE is connected to; sum of their PRE-move rewards */
                agg = 0;
                for (int j = 0; j < N; j++) {
                    int shared = 0;
                    for (int b = 0; b < M && !shared; b++) shared = e->mask[i * M + b] && e->mask[j * M + b];
                    if (shared) agg += e->reward_before[j];
                }
            } else {
                agg = mn;
            }
        }
        e->reward[i] = agg;
    }
}

orc_env *orc_create(int kind, int n_ue, int n_bs, const double *bs_xy, int width, int height, const int *sharing,
                    const double *vel_spec, const double *init_xy, long long seed, int has_seed, int reward_agg,
                    int rand_episodes, int pause_duration, int border_buffer) {
    orc_env *e = (orc_env *)calloc(1, sizeof(orc_env));
    const int N = n_ue, M = n_bs;
    e->kind = kind; e->n_ue = N; e->n_bs = M; e->width = width; e->height = height; e->reward_agg = reward_agg;
    e->rand_episodes = rand_episodes; e->pause_duration = pause_duration; e->border_buffer = border_buffer;
    e->seed = seed; e->has_seed = has_seed;
    /* station.py:112-114 */
    double ch = 0.8 + (1.1 * log10(2500.0) - 0.7) * 1.5 - 1.56 * log10(2500.0);
    e->const1 = 69.55 + 26.16 * log10(2500.0) - 13.82 * log10(50.0) - ch;
    e->const2 = 44.9 - 6.55 * log10(50.0);
#define ALLOC(p, n) p = calloc((size_t)(n), sizeof(*(p)))
    ALLOC(e->bs_xy, 2 * M); memcpy(e->bs_xy, bs_xy, sizeof(double) * 2 * M);
    ALLOC(e->sharing, M); memcpy(e->sharing, sharing, sizeof(int) * M);
    ALLOC(e->vel_spec, N); memcpy(e->vel_spec, vel_spec, sizeof(double) * N);
    ALLOC(e->init_xy, 2 * N); memcpy(e->init_xy, init_xy, sizeof(double) * 2 * N);
    ALLOC(e->rng, N); ALLOC(e->mrng, N);
    ALLOC(e->pos, 2 * N); ALLOC(e->wp, 2 * N); ALLOC(e->vel, N); ALLOC(e->pausing, N); ALLOC(e->curr_pause, N);
    ALLOC(e->mask, N * M); ALLOC(e->ewma, N); ALLOC(e->link_rate, N * M); ALLOC(e->snr, N * M);
    ALLOC(e->curr_dr, N); ALLOC(e->utility, N); ALLOC(e->reward_before, N); ALLOC(e->lost_conn, N);
    ALLOC(e->obs, orc_obs_size(e)); ALLOC(e->reward, orc_reward_size(e));
    ALLOC(e->bs_cnt, M); ALLOC(e->bs_inv, M); ALLOC(e->bs_prio, M); ALLOC(e->bs_util_sum, M);
    ALLOC(e->bs_util_min, M); ALLOC(e->bs_argmax, M);
    /* single_ue/base.py:132-143: seeded once at construction (base.py:60) */
    for (int i = 0; i < N; i++) {
        long long s = has_seed ? seed + 100ll * (i + 1) : 12345ll + i;
        mt_seed_int(&e->rng[i], s);
        mt_seed_int(&e->mrng[i], s);
    }
    return e;
}

/* per-env switches that leave the step itself untouched: User.util_func / dr_req and the observation normalisation */
void orc_set_variants(orc_env *e, int util_step, double dr_req, int obs_maxnorm) {
    e->util_step = util_step; e->dr_req = dr_req; e->obs_maxnorm = obs_maxnorm;
}

/* interference extension on / off (takes effect with the next reset / step) */
void orc_set_interference(orc_env *e, int on) { e->interference = on; }

void orc_destroy(orc_env *e) {
    if (!e) return;
    free(e->bs_xy); free(e->sharing); free(e->vel_spec); free(e->init_xy); free(e->rng); free(e->mrng);
    free(e->pos); free(e->wp); free(e->vel); free(e->pausing); free(e->curr_pause); free(e->mask); free(e->ewma);
    free(e->link_rate); free(e->snr); free(e->curr_dr); free(e->utility); free(e->reward_before);
    free(e->lost_conn); free(e->obs); free(e->reward); free(e->bs_cnt); free(e->bs_inv); free(e->bs_prio);
    free(e->bs_util_sum); free(e->bs_util_min); free(e->bs_argmax);
    free(e);
}

void orc_reset(orc_env *e) {
    /* single_ue/base.py:169-189 */
    const int N = e->n_ue, M = e->n_bs;
    if (!e->rand_episodes && e->has_seed)
        for (int i = 0; i < N; i++) {
            long long s = e->seed + 100ll * (i + 1);
            mt_seed_int(&e->rng[i], s);
            mt_seed_int(&e->mrng[i], s);
        }
    e->time = 0;
    for (int i = 0; i < N; i++) {
        /* user.py:98-116 */
        double px = e->init_xy[2 * i], py = e->init_xy[2 * i + 1];
        if (isnan(px)) px = mt_randint(&e->rng[i], 0, e->width);
        if (isnan(py)) py = mt_randint(&e->rng[i], 0, e->height);
        e->pos[2 * i] = px; e->pos[2 * i + 1] = py;
        movement_reset(e, i);
        e->ewma[i] = 0;
        e->lost_conn[i] = 0;
        e->reward_before[i] = 0;
    }
    memset(e->mask, 0, (size_t)N * M);
    compute_snr(e);
    update_rates(e);
    build_obs_reward(e);
    e->sum_utility = 0;
    for (int i = 0; i < N; i++) e->sum_utility += e->utility[i];
}

void orc_step(orc_env *e, const int *actions) {
    /* single_ue/base.py:413-466 */
    const int N = e->n_ue, M = e->n_bs;
    /* base.py:247-282 / user.py:190-229: toggle; e->snr holds the SNR at the current (pre-move) positions */
    for (int i = 0; i < N; i++) {
        int a = actions[i];
        if (a <= 0 || a > M) continue;
        int b = a - 1;
        if (e->mask[i * M + b]) e->mask[i * M + b] = 0;
        else if (e->snr[i * M + b] > SNR_THRESHOLD) e->mask[i * M + b] = 1;
    }
    /* base.py:446: rates + rewards before moving */
    update_rates(e);
    for (int i = 0; i < N; i++) {
        double u = e->utility[i];                                  /* base.py:158-167 */
        u = u < MIN_UTILITY ? MIN_UTILITY : (u > MAX_UTILITY ? MAX_UTILITY : u);
        e->reward_before[i] = u / MAX_UTILITY;
    }
    /* base.py:447 / user.py:159-188,148-157 */
    for (int i = 0; i < N; i++) {
        movement_step(e, i);
        int lost = 0;
        double dr = 0;
        double q[64];
        ue_quality(e, i, q);
        for (int b = 0; b < M; b++) {
            if (!e->mask[i * M + b]) continue;
            double s = q[b];
            if (!(s > SNR_THRESHOLD)) { e->mask[i * M + b] = 0; lost++; }
            else dr += e->link_rate[i * M + b];
        }
        e->lost_conn[i] = lost;
        e->ewma[i] = 0.9 * dr + (1 - 0.9) * e->ewma[i];
    }
    /* base.py:451 */
    compute_snr(e);
    update_rates(e);
    e->time += 1;
    e->sum_utility = 0;
    for (int i = 0; i < N; i++) e->sum_utility += e->utility[i];
    e->total_utility += e->sum_utility;
    build_obs_reward(e);
}

/* copy-out accessors (NULL pointers are skipped) */
void orc_get(const orc_env *e, double *pos, uint8_t *mask, double *link_rates, double *snr, double *curr_dr,
             double *ewma, double *utility, double *movement, double *obs, double *reward, int *lost_conn,
             double *sum_utility, int *time) {
    const int N = e->n_ue, M = e->n_bs;
    if (pos) memcpy(pos, e->pos, sizeof(double) * 2 * N);
    if (mask) memcpy(mask, e->mask, (size_t)N * M);
    if (link_rates) memcpy(link_rates, e->link_rate, sizeof(double) * N * M);
    if (snr) memcpy(snr, e->snr, sizeof(double) * N * M);
    if (curr_dr) memcpy(curr_dr, e->curr_dr, sizeof(double) * N);
    if (ewma) memcpy(ewma, e->ewma, sizeof(double) * N);
    if (utility) memcpy(utility, e->utility, sizeof(double) * N);
    if (movement)
        for (int i = 0; i < N; i++) {
            movement[5 * i] = e->vel[i]; movement[5 * i + 1] = e->wp[2 * i]; movement[5 * i + 2] = e->wp[2 * i + 1];
            movement[5 * i + 3] = e->pausing[i]; movement[5 * i + 4] = e->curr_pause[i];
        }
    if (obs) memcpy(obs, e->obs, sizeof(double) * orc_obs_size(e));
    if (reward) memcpy(reward, e->reward, sizeof(double) * orc_reward_size(e));
    if (lost_conn) memcpy(lost_conn, e->lost_conn, sizeof(int) * N);
    if (sum_utility) *sum_utility = e->sum_utility;
    if (time) *time = e->time;
}

/* Step K independent envs T times (actions [T][K][N]); OpenMP over envs.  Used as the native CPU baseline. */
void orc_batch_run(orc_env **envs, int K, const int *actions, int T, int nthreads) {
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
    (void)nthreads;
#pragma omp parallel for schedule(static)
    for (int k = 0; k < K; k++) {
        const int N = envs[k]->n_ue;
        for (int t = 0; t < T; t++) orc_step(envs[k], actions + ((size_t)t * K + k) * N);
    }
}

/* As orc_batch_run, recording what the soak tests compare bit for bit: the connection mask of every UE after every step
 * (bit b = linked to BS b, [T][K][N]), the links each UE lost through movement ([T][K][N]), the step rewards
 * ([T][K][R]) and the positions after every `pos_every`-th step ([T / pos_every][K][N][2]).  NULL outputs are skipped. */
void orc_batch_trace(orc_env **envs, int K, const int *actions, int T, int nthreads, uint64_t *mask_out,
                     uint8_t *lost_out, double *reward_out, double *pos_out, int pos_every) {
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
    (void)nthreads;
#pragma omp parallel for schedule(static)
    for (int k = 0; k < K; k++) {
        orc_env *e = envs[k];
        const int N = e->n_ue, M = e->n_bs, R = orc_reward_size(e);
        for (int t = 0; t < T; t++) {
            orc_step(e, actions + ((size_t)t * K + k) * N);
            const size_t row = ((size_t)t * K + k) * N;
            for (int i = 0; i < N; i++) {
                if (mask_out) {
                    uint64_t m = 0;
                    for (int b = 0; b < M; b++) m |= (uint64_t)(e->mask[i * M + b] != 0) << b;
                    mask_out[row + i] = m;
                }
                if (lost_out) lost_out[row + i] = (uint8_t)e->lost_conn[i];
            }
            if (reward_out) memcpy(reward_out + ((size_t)t * K + k) * R, e->reward, sizeof(double) * R);
            if (pos_out && pos_every > 0 && (t + 1) % pos_every == 0)
                memcpy(pos_out + (((size_t)((t + 1) / pos_every - 1) * K + k) * N) * 2, e->pos, sizeof(double) * 2 * N);
        }
    }
}

int orc_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
