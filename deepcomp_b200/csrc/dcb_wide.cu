// Wide-env step kernel: ONE CTA PER ENV, for envs too large for the fused kernel of dcb_step.cu (n_ue > 512, or an
// observation tile + link matrix that do not fit a CTA's shared memory; BASELINE config 4: 1000 UE x 50 BS).
//
// Same path, same state slabs, same outputs as dcb_step.cu (MobileEnv.step, deepcomp/env/single_ue/base.py:413-466); all
// file:line citations are relative to /root/reference/deepcomp/.  Mapping:
//   * per-UE phases  -- thread u < N owns UE u and carries its state in registers across the T steps of a launch:
//     action, rates before the move, reward, move, link drop, EWMA, rates after the move, utility;
//   * per-BS phases  -- one warp per base station walks the bitset of the UEs linked to it, one 32-UE word per lane, and
//     folds count / sum / arg-max (sharing models, station.py:152-202) or count / utility sum / min (station.py:63-83)
//     with warp shuffles;
//   * per-row phase  -- one warp per UE row of the observation, lanes over the base stations: squared distances, the
//     closest BS by a shuffle min-reduction, the in-range set by ballot, normalised SNR, and the row leaves as coalesced
//     128-byte segments straight to the observation buffer (no staging tile: one env's observation is up to 804 KB).
//   * interference extension (general instance only) -- one more row-parallel phase right after the move: the SNR of every
//     pair of the env, per-UE interference sum, and what of the observation depends on the positions alone
//     (wide_interference_pass).
// A UE's link values live in a compact per-UE slot list (slot = rank of the BS in the UE's mask) instead of the dense
// [N][M] matrix of the fused kernel; the slot capacity LC is the largest number of base stations any point of the map
// can be in range of (host-computed bound, dcb_api.cu).
// Unlike the fused kernel there is no pipelining between steps and no inheritance of aggregates: every step evaluates the
// link rates twice (before / after the move) exactly as the reference does.
#include "dcb_device.cuh"

namespace {

typedef unsigned long long u64;

// What the row and reward phases need of a UE, in one 48-byte record (16-byte loads / one store at ONE address per row
// instead of six arrays with an address each)
struct __align__(16) UeRec {
    double x, y;       // position after the move
    double util;       // utility after the move
    u64 mask;          // linked base stations
    u64 inr;           // base stations in range (bit b), left by the row phase for the reward phase
    double rb;         // reward before the move
};

struct WideSmem {
    MathTables *tab;
    double2 *bsxy;
    int *share;
    double *vthr;
    double *Xs;        // [N][LC] link values / cached shared rates
    UeRec *rec;        // [N] position / utility after the move, link mask, in-range set, reward before the move
    double *sew;       // [N] EWMA rate
    uint2 *smv;        // [N] packed movement state
    unsigned *bits;    // [M][NW] UEs linked to each BS
    double *fac;       // [M]
    int *arg;          // [M]
    int *cnt_obs;      // [M]
    double *usum, *umin;   // [M]
    float *f_ues, *f_util; // [M]
    double *env_red;   // [2] per-env reward / utility sum
    // general instance (EXT): data-rate observation classes and the interference extension
    double *sdr;       // [N] curr_dr after the move (user.py:64-69)
    // interference extension, per UE at its current position: the largest SNR over the base stations, which BS that is, and
    // the sum of all the OTHER base stations' SNR (kept apart: a UE on top of a BS has snr ~ 1e52 there, which would absorb
    // the rest of the sum)
    double *ssum;      // [N] sum of the SNR over all BS but the strongest
    double *smax;      // [N] SNR of the strongest BS
    int *sbmax;        // [N] index of the strongest BS
    double *lsum, *lbest;  // [M] raw link aggregates of the current masks: sum of link values, largest unshared rate
    int *lcnt;         // [M] ... number of linked UEs
};

// Signal quality of one UE-BS pair: SNR (station.py:122-127) or, with the interference extension, SINR =
// snr_b / (1 + sum_{b' != b} snr_b'), from the per-UE triple (sum over all BS but the strongest, the strongest, which one).
// The reciprocal is dcb_rcp (<= 1 ulp, a quarter of the instructions of an IEEE division); every place that judges or
// uses a link goes through this one function, so the in-range decisions of all phases agree bit for bit.
struct UeInterference { double rest, smax; int bmax; };
__device__ __forceinline__ double pair_sinr(double snr, int b, const UeInterference &u) {
    const double others = b == u.bmax ? u.rest : (u.rest - snr) + u.smax;
    return snr * dcb_rcp(1.0 + others);
}

// SNR of ANY pair of the map for the interference pass, where most pairs are far beyond the connection range: the table
// form of dcb_snr_inrange takes the binary exponent of d^2 modulo 16, so exponents 16..31 (256 m <= d < 65 km) are the
// entries of 0..15 times k16 = 2^(-16 h) -- no log2 / exp2 round trip and no divergent call for a far base station.
__device__ __forceinline__ double snr_of_d2_anywhere(const double *pw, double c0, double h, const MathTables *tab, double k16,
                                                     double d2) {
    if (d2 >= 1.0 && d2 < 4294967296.0) {
        const double s = dcb_snr_inrange(tab, pw, d2);
        return d2 < DCB_FAR_D2 ? s : s * k16;
    }
    return snr_of_d2_general(c0, h, tab, d2);
}
__device__ __forceinline__ double snr_of_d2_anywhere(const DevParams &p, const MathTables *tab, double k16, double d2) {
    return snr_of_d2_anywhere(p.pw, p.snr_c0, p.snr_h, tab, k16, d2);
}

__device__ __forceinline__ int rank_of(u64 mask, int b) { return __popcll(mask & (((u64)1 << b) - 1)); }

// per-BS reduction over the linked UEs: count, sum of link values, first arg-max -> sharing factor (one warp per BS)
template <bool EXT>
__device__ __forceinline__ void wide_reduce_links(const WideSmem &S, int N, int M, int NW, int LC, bool want_arg,
                                                  int warp, int lane, int nwarps) {
    for (int b = warp; b < M; b += nwarps) {
        int c = 0, a0 = 0x7fffffff;
        double s = 0.0, best = 0.0;
        for (int w = lane; w < NW; w += 32) {
            unsigned wa = S.bits[b * NW + w];
            c += __popc(wa);
            while (wa) {
                const int j = __ffs(wa) - 1;
                wa &= wa - 1;
                const int i = (w << 5) + j;
                const double v = S.Xs[(size_t)i * LC + rank_of(S.rec[i].mask, b)];
                s += v;
                if (want_arg && v > best) { best = v; a0 = i; }      // station.py:184: first arg-max
            }
        }
        for (int off = 16; off > 0; off >>= 1) {
            c += __shfl_xor_sync(0xffffffffu, c, off);
            s += __shfl_xor_sync(0xffffffffu, s, off);
            if (want_arg) {
                const double ob = __shfl_xor_sync(0xffffffffu, best, off);
                const int oi = __shfl_xor_sync(0xffffffffu, a0, off);
                if (ob > best || (ob == best && oi < a0)) { best = ob; a0 = oi; }
            }
        }
        if (lane == 0) {
            S.fac[b] = share_factor(S.share[b], c, s);
            S.arg[b] = a0;
            if (EXT) { S.lcnt[b] = c; S.lsum[b] = s; S.lbest[b] = best; }
        }
    }
}

// per-BS utility aggregates for the observation / multi-agent reward (one warp per BS)
__device__ __forceinline__ void wide_reduce_utility(const WideSmem &S, int NA, int M, int NW, bool want_min, int warp,
                                                    int lane, int nwarps) {
    const double inv_n = 1.0 / (double)NA;      // self.num_ue = UEs present (variants.py:296)
    for (int b = warp; b < M; b += nwarps) {
        int c = 0;
        double s = 0.0, mn = DCB_MAX_UTILITY;
        for (int w = lane; w < NW; w += 32) {
            unsigned wa = S.bits[b * NW + w];
            c += __popc(wa);
            while (wa) {
                const int j = __ffs(wa) - 1;
                wa &= wa - 1;
                const double uu = S.rec[(w << 5) + j].util;
                s += uu;
                if (want_min) mn = uu < mn ? uu : mn;
            }
        }
        for (int off = 16; off > 0; off >>= 1) {
            c += __shfl_xor_sync(0xffffffffu, c, off);
            s += __shfl_xor_sync(0xffffffffu, s, off);
            if (want_min) {
                const double o = __shfl_xor_sync(0xffffffffu, mn, off);
                mn = o < mn ? o : mn;
            }
        }
        if (lane == 0) {
            S.cnt_obs[b] = c;
            S.usum[b] = s;
            S.umin[b] = mn;
            S.f_ues[b] = (float)((double)c * inv_n);                                               // variants.py:296
            S.f_util[b] = c > 0 ? (float)(s * dcb_rcp((double)c) * (1.0 / DCB_MAX_UTILITY)) : 0.0f; // station.py:71-76
        }
    }
}

// max over the warp of NON-NEGATIVE doubles: their bit patterns order like unsigned integers, so two redux.sync (high
// words, then the low words of the lanes that hold the largest high word) replace a five-level shuffle ladder
__device__ __forceinline__ double warp_max_nonneg(double v) {
    const unsigned hi = (unsigned)__double2hiint(v), lo = (unsigned)__double2loint(v);
    const unsigned mh = __reduce_max_sync(0xffffffffu, hi);
    const unsigned ml = __reduce_max_sync(0xffffffffu, hi == mh ? lo : 0u);
    return __hiloint2double((int)mh, (int)ml);
}
__device__ __forceinline__ double warp_sum(double v) {
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    return v;
}
__device__ __forceinline__ double warp_min(double v) {
    for (int off = 16; off > 0; off >>= 1) {
        const double o = __shfl_xor_sync(0xffffffffu, v, off);
        v = o < v ? o : v;
    }
    return v;
}

// ------------------------------------------------------------------------------------------------ interference pass
// Interference extension, the pass over ALL pairs of the env at the UEs' current positions (UeRec::x / y): one warp per UE
// row, lanes over the base stations (two passes cover M <= 64).  The SNR of every BS, the strongest BS (first index on
// ties) and the sum of the OTHERS -> ssum / smax / sbmax, which every later judgement of a link of that UE reads.  With
// `emit` (the positions are those of this step's observation) the same pass finishes everything that depends on the
// positions alone, so that no pair is evaluated twice per step: the SINR of every pair, the in-range set for the reward
// phase (UeRec::inr), and the observation entry 'dr' = SINR_b / max SINR (variants.py:276-284; MaxNorm: variants.py:322-330)
// -- written straight to the row's 'dr' segment of the observation buffer; the row phase adds the other segments.  The
// data-rate observation classes need the per-BS aggregates of the step as well and keep their own row phase.
//
// A function of its own, NOT inlined: inside the 1024-thread kernel (64 registers) the row loop shared its registers with
// everything else that is live across a step, and 15 % of its instructions re-read %tid and re-derived shared-memory
// addresses (ncu source view, profiles/r02_wide_interf_hotlines.txt).  The radio constants it needs come from constant
// memory, as instruction operands -- like DevParams in the kernel: pw[0..9] (binomial series of dcb_snr_inrange), snr_c0,
// snr_h; the same for every handle (dcb_create derives them from the reference's constants), uploaded by
// dcb_wide_upload_interference_constants.
__constant__ double c_interf[12];

struct InterfPass {
    int off_tab, off_bsxy, off_rec, off_ssum, off_smax, off_sbmax;   // shared-memory offsets (WideLayout)
    int NA, M, nwarps;
    int dr_row, dr_off;          // the 'dr' entry of (row r, BS b) is element r * dr_row + dr_off + b of the env's observation
    int emit, want_inr, maxnorm;
    double k16;                  // 2^(-16 h), snr_of_d2_anywhere
    float *obs_env;              // this step's observation of the env, or NULL
    double *dbg_env, *dbg_snr;   // fp64 taps of the launch's last step, or NULL
};

__device__ __noinline__ void wide_interference_pass(const InterfPass q) {
    extern __shared__ __align__(128) unsigned char smem[];
    const MathTables *tab = reinterpret_cast<const MathTables *>(smem + q.off_tab);
    const double2 *bsxy = reinterpret_cast<const double2 *>(smem + q.off_bsxy);
    UeRec *rec = reinterpret_cast<UeRec *>(smem + q.off_rec);
    double *ssum = reinterpret_cast<double *>(smem + q.off_ssum);
    double *smax = reinterpret_cast<double *>(smem + q.off_smax);
    int *sbmax = reinterpret_cast<int *>(smem + q.off_sbmax);
    // (volatile: read once and kept, where ptxas would re-read %tid and mask it at every use in the row loop)
    unsigned lane_u;
    asm volatile("mov.u32 %0, %%laneid;" : "=r"(lane_u));
    const int lane = (int)lane_u, warp = threadIdx.x >> 5;
    const int M = q.M;
    const int b0 = lane, b1 = lane + 32;
    const bool ok0 = b0 < M, ok1 = b1 < M;
    const double2 bs0 = ok0 ? bsxy[b0] : make_double2(0.0, 0.0), bs1 = ok1 ? bsxy[b1] : make_double2(0.0, 0.0);
    const double k16 = q.k16;
    float *const obs_env = q.obs_env;
    double *const dbg_env = q.dbg_env, *const dbg_snr = q.dbg_snr;
    for (int r = warp; r < q.NA; r += q.nwarps) {
        const double rx = rec[r].x, ry = rec[r].y;
        const double v0 = ok0 ? snr_of_d2_anywhere(c_interf, c_interf[10], c_interf[11], tab, k16, dist2(bs0, rx, ry)) : 0.0;
        const double v1 = ok1 ? snr_of_d2_anywhere(c_interf, c_interf[10], c_interf[11], tab, k16, dist2(bs1, rx, ry)) : 0.0;
        // the strongest BS, first index on ties: the SNR is positive (0 on the lanes past M), so the maximum is two
        // redux.sync on the bit patterns and its owner the first set bit of two ballots
        const double mx = warp_max_nonneg(v1 > v0 ? v1 : v0);
        const unsigned eq0 = __ballot_sync(0xffffffffu, v0 == mx), eq1 = __ballot_sync(0xffffffffu, v1 == mx);
        const int bm = eq0 ? __ffs(eq0) - 1 : 31 + __ffs(eq1);
        const double v = warp_sum((b0 == bm ? 0.0 : v0) + (b1 == bm ? 0.0 : v1));
        if (lane == 0) { ssum[r] = v; smax[r] = mx; sbmax[r] = bm; }
        if (!q.emit) continue;
        UeInterference tot;
        tot.rest = v; tot.smax = mx; tot.bmax = bm;
        const double q0 = ok0 ? pair_sinr(v0, b0, tot) : 0.0, q1 = ok1 ? pair_sinr(v1, b1, tot) : 0.0;
        if (q.want_inr) {
            const unsigned in0 = __ballot_sync(0xffffffffu, q0 > DCB_SNR_THRESHOLD);
            const unsigned in1 = __ballot_sync(0xffffffffu, q1 > DCB_SNR_THRESHOLD);
            if (lane == 0) rec[r].inr = ((u64)in1 << 32) | in0;
        }
        float dr0, dr1;
        if (q.maxnorm) {
            dr0 = max_norm_snr(q0); dr1 = max_norm_snr(q1);
        } else {
            const double inv_max = dcb_rcp(warp_max_nonneg(q0 > q1 ? q0 : q1));   // (some SINR is > 0: M >= 1)
            dr0 = (float)(q0 * inv_max); dr1 = (float)(q1 * inv_max);
        }
        const size_t e0 = (size_t)r * q.dr_row + q.dr_off + b0;
        if (obs_env) {
            if (ok0) obs_env[e0] = dr0;
            if (ok1) obs_env[e0 + 32] = dr1;
        }
        if (dbg_env) {
            if (ok0) dbg_env[e0] = (double)dr0;
            if (ok1) dbg_env[e0 + 32] = (double)dr1;
        }
        if (dbg_snr) {
            if (ok0) dbg_snr[(size_t)r * M + b0] = q0;
            if (ok1) dbg_snr[(size_t)r * M + b1] = q1;
        }
    }
}

// PAD: the envs have padding slots (NA < N, variable UE population)
// EXT: the general instance -- data-rate observation classes (dcb_set_obs_variant), the interference extension
// (dcb_set_interference), UniformMovement UEs, the no-move launch mode; the plain instances compile without them
// CENTRAL: observation / reward layout of the central agent instead of the per-UE multi-agent one (compile-time, as in
// the fused kernel: it selects store patterns and the reward code of every row)
template <bool PAD, bool EXT, bool CENTRAL>
__global__ void __launch_bounds__(1024, 1) dcb_wide_kernel(const __grid_constant__ StepArgs a) {
    extern __shared__ __align__(128) unsigned char smem[];
    const DevParams &p = a.p;
    const WideLayout &L = a.W;
    const int N = p.N, M = p.M, LC = p.LC;
    const int NW = (N + 31) >> 5;
    WideSmem S;
    S.tab = reinterpret_cast<MathTables *>(smem + L.off_tab);
    S.bsxy = reinterpret_cast<double2 *>(smem + L.off_bsxy);
    S.share = reinterpret_cast<int *>(smem + L.off_share);
    S.vthr = reinterpret_cast<double *>(smem + L.off_vthr);
    S.Xs = reinterpret_cast<double *>(smem + L.off_xs);
    S.rec = reinterpret_cast<UeRec *>(smem + L.off_rec);
    S.sew = reinterpret_cast<double *>(smem + L.off_sew);
    S.smv = reinterpret_cast<uint2 *>(smem + L.off_smv);
    S.bits = reinterpret_cast<unsigned *>(smem + L.off_bits);
    S.fac = reinterpret_cast<double *>(smem + L.off_fac);
    S.arg = reinterpret_cast<int *>(smem + L.off_arg);
    S.cnt_obs = reinterpret_cast<int *>(smem + L.off_cnt);
    S.usum = reinterpret_cast<double *>(smem + L.off_usum);
    S.umin = reinterpret_cast<double *>(smem + L.off_umin);
    S.f_ues = reinterpret_cast<float *>(smem + L.off_fues);
    S.f_util = reinterpret_cast<float *>(smem + L.off_futil);
    S.env_red = reinterpret_cast<double *>(smem + L.off_env);
    S.sdr = reinterpret_cast<double *>(smem + L.off_sdr);
    S.ssum = reinterpret_cast<double *>(smem + L.off_ssum);
    S.smax = reinterpret_cast<double *>(smem + L.off_smax);
    S.sbmax = reinterpret_cast<int *>(smem + L.off_sbmax);
    S.lsum = reinterpret_cast<double *>(smem + L.off_lsum);
    S.lbest = reinterpret_cast<double *>(smem + L.off_lbest);
    S.lcnt = reinterpret_cast<int *>(smem + L.off_lcnt);
    const MathTables *tab = S.tab;
    const bool interf = EXT && p.interference;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = a.threads >> 5;   // a.threads == blockDim.x
    const int k = blockIdx.x;
    const int NA = PAD ? p.NA : N;         // slots [0, NA) hold UEs, the rest is padding (max_ues > num_ue)
    const bool valid = tid < NA;
    const int i = valid ? tid : 0;
    const long long u = (long long)k * N + i;
    constexpr bool central = CENTRAL;
    const int OW = obs_width(central ? DCB_KIND_CENTRAL : DCB_KIND_MULTI, M);
    const size_t per_env = (EXT && p.obs_var) ? (size_t)p.var_obs_size
                                              : (central ? (size_t)(2 * N * M + N) : (size_t)N * OW);
    const int T = a.T;
    const int n_iter = T > 0 ? T : 1;
    const float hr = p.snr_hr;

    dcb_math_init(S.tab, S.vthr, tid, blockDim.x, p.tabs);
    for (int b = tid; b < M; b += blockDim.x) {
        S.bsxy[b] = make_double2(p.bs_xy[2 * b], p.bs_xy[2 * b + 1]);
        S.share[b] = p.sharing[b];
    }
    for (int j = tid; j < M * NW; j += blockDim.x) S.bits[j] = 0u;

    // ---- per-UE state: global slabs -> shared memory for the launch.  A UE's thread pulls it into registers for the
    // per-UE phases of a step only, so the row-parallel phase (most of the instructions) has the registers to itself
    int tk = p.time[k];
    if (valid) {
        const double2 ps = p.pos[u];
        S.rec[i].x = ps.x; S.rec[i].y = ps.y;
        S.smv[i] = p.mv[u];
        S.rec[i].mask = p.mask[u];
        S.sew[i] = p.ewma[u];
    }
    const double vfix = valid ? (p.vel_u ? p.vel_u[u] : p.vel_spec[i]) : 0.0;
    const double vfix_thr = vfix >= 0.0 ? snap_threshold(vfix) : 0.0;
    // UniformMovement UEs (movement.py:26-80) and the no-move launch mode (DCB_STEPF_NO_MOVE)
    int ukx = 0, uky = 0;
    double uvx = 0.0, uvy = 0.0;
    if (EXT && p.uni_kind && valid) {
        ukx = p.uni_kind[2 * i]; uky = p.uni_kind[2 * i + 1];
        uvx = p.uni_val[2 * i]; uvy = p.uni_val[2 * i + 1];
    }
    const bool no_move = EXT && (a.flags & DCB_STEPF_NO_MOVE) != 0;
    double *Xrow = S.Xs + (size_t)i * LC;
    __syncthreads();
    // interference extension: the pass over all pairs of the env (wide_interference_pass above)
    const double k16 = interf ? dcb_exp2(tab, -16.0 * p.snr_h) : 1.0;
    auto interference_pass = [&](bool emit, int step, bool last) {
        InterfPass q;
        q.off_tab = L.off_tab; q.off_bsxy = L.off_bsxy; q.off_rec = L.off_rec;
        q.off_ssum = L.off_ssum; q.off_smax = L.off_smax; q.off_sbmax = L.off_sbmax;
        q.NA = NA; q.M = M; q.nwarps = nwarps;
        q.dr_row = central ? M : OW; q.dr_off = central ? N * M : M;
        emit = emit && !p.obs_var;
        q.emit = emit; q.want_inr = emit && !central && T > 0; q.maxnorm = p.obs_maxnorm;
        q.k16 = k16;
        q.obs_env = (emit && a.out.obs) ? a.out.obs + (size_t)step * a.out.obs_stride + (size_t)k * per_env : nullptr;
        q.dbg_env = (emit && last && a.out.dbg_obs) ? a.out.dbg_obs + (size_t)k * per_env : nullptr;
        q.dbg_snr = (emit && last && a.out.dbg_snr) ? a.out.dbg_snr + (size_t)k * N * M : nullptr;
        wide_interference_pass(q);
    };
    if (interf) {
        // the sums at the positions the launch starts from (an observe-only launch: that is its observation); every step
        // refreshes them after its move
        interference_pass(T == 0, 0, true);
        __syncthreads();
    }
    // in range (can_connect, station.py:222-226) and unshared rate (station.py:129-138) of one pair, SNR or SINR based
    auto ue_interf = [&](int r) -> UeInterference {
        UeInterference q;
        q.rest = 0.0; q.smax = 0.0; q.bmax = -1;
        if (interf) { q.rest = S.ssum[r]; q.smax = S.smax[r]; q.bmax = S.sbmax[r]; }
        return q;
    };
    auto in_range = [&](double d2, int b, const UeInterference &q) -> bool {
        if (interf) return pair_sinr(snr_of_d2_anywhere(p, tab, k16, d2), b, q) > DCB_SNR_THRESHOLD;
        return d2 <= p.thr_d2;
    };
    auto unshared_rate = [&](double d2, int b, const UeInterference &q) -> double {
        if (interf) return DCB_BW * dcb_log2_1p(tab, pair_sinr(snr_of_d2_anywhere(p, tab, k16, d2), b, q));
        return rate_of_d2(p, tab, d2);
    };

    for (int step = 0; step < n_iter; step++) {
        const bool last = step == n_iter - 1;
        double rb = 0.0, dr = 0.0, util = DCB_MIN_UTILITY;
        int lost = 0;
        double x = 0, y = 0, ewma = 0;
        u64 mask = 0;
        unsigned wxy = 0, vpt = 0;
        if (valid) {
            x = S.rec[i].x; y = S.rec[i].y; ewma = S.sew[i]; mask = S.rec[i].mask;
            const uint2 mv = S.smv[i];
            wxy = mv.x; vpt = mv.y;
        }
        if (T > 0) {
            // ---- MobileEnv.reset before the step of an env that reached its episode length (base.py:169-189)
            if (p.auto_reset && tk >= p.episode_length) {
                if (valid) {
                    ue_reset(p, u, x, y, wxy, vpt);
                    mask = 0; ewma = 0.0;
                }
                tk = 0;
                if (interf) {               // the SNR sums follow the UEs to their initial positions (CTA-uniform branch)
                    if (valid) { S.rec[i].x = x; S.rec[i].y = y; }
                    __syncthreads();
                    interference_pass(false, step, last);
                    __syncthreads();
                }
            }
            if (valid) {
                // ---- apply_ue_actions (base.py:247-282) -> User.connect_to_bs(disconnect=True) (user.py:190-229)
                int act;
                if (a.pol.kind) {
                    act = policy_action<u64>(a.pol, mask, x, y, S.bsxy, M, i, a.pol.call0 + step, u);
                    if (a.actions_out) a.actions_out[(size_t)step * p.K * N + u] = act;
                } else {
                    act = a.actions[(size_t)step * p.K * N + u];
                }
                if (act < 0 || act > M) {
                    atomicOr(p.err, DCB_ERRBIT_ACTION);
                } else if (act > 0) {
                    const int b = act - 1;
                    const u64 bit = (u64)1 << b;
                    if (mask & bit) mask &= ~bit;
                    else if (in_range(dist2(S.bsxy[b], x, y), b, ue_interf(i))) mask |= bit;   // can_connect, station.py:222-226
                }
                if (__popcll(mask) > LC) {          // cannot happen on a reachable state (LC bounds the BS in range)
                    atomicOr(p.err, DCB_ERRBIT_LINKS);
                    while (__popcll(mask) > LC) mask &= mask - 1;
                }
                // ---- link values at the pre-move position (station.py:129-150)
                const double iee = dcb_rcp(ewma + DCB_EPSILON);
                int slot = 0;
                const UeInterference tot0 = ue_interf(i);
                for (u64 m = mask; m; m &= m - 1, slot++) {
                    const int b = __ffsll((long long)m) - 1;
                    Xrow[slot] = link_value(S.share[b], unshared_rate(dist2(S.bsxy[b], x, y), b, tot0), iee);
                    atomicOr(&S.bits[b * NW + (i >> 5)], 1u << (i & 31));
                }
                S.rec[i].mask = mask;
            }
            __syncthreads();
            wide_reduce_links<EXT>(S, N, M, NW, LC, p.has_maxcap, warp, lane, nwarps);
            __syncthreads();
            for (int j = tid; j < M * NW; j += blockDim.x) S.bits[j] = 0u;
            // check_bs_connection (user.py:175-188) + update_ewma_dr (user.py:148-157) at the new position
            auto drop_links = [&]() {
                const UeInterference tot1 = ue_interf(i);
                double keep = 0.0;
                int slot = 0;
                for (u64 m = mask; m; m &= m - 1, slot++) {
                    const int b = __ffsll((long long)m) - 1;
                    if (in_range(dist2(S.bsxy[b], x, y), b, tot1)) keep += Xrow[slot];
                    else { mask &= ~((u64)1 << b); lost++; }
                }
                ewma = 0.9 * keep + (1 - 0.9) * ewma;
            };
            if (valid) {
                // ---- update_ue_drs_rewards (base.py:315-335): shared rate of every connected link -> ue.bs_dr cache
                // (back into the slots), calc_reward (base.py:158-167; penalties are identically 0, base.py:257)
                const double ee = ewma + DCB_EPSILON;
                double dr0 = 0.0;
                int slot = 0;
                for (u64 m = mask; m; m &= m - 1, slot++) {
                    const int b = __ffsll((long long)m) - 1;
                    const double r = shared_rate(S.share[b], Xrow[slot], S.fac[b], S.arg[b], i, ee);
                    Xrow[slot] = r;
                    dr0 += r;                                                          // user.py:64-69
                }
                rb = ue_utility(p, tab, dr0) * (1.0 / DCB_MAX_UTILITY);
                // ---- User.move (user.py:159-173), check_bs_connection (user.py:175-188), update_ewma_dr (user.py:148-157)
                if (!no_move) {
                    if (ukx) ue_move_uniform(p, ukx, uky, uvx, uvy, x, y, wxy, vpt);
                    else ue_move<false>(p, u, vfix, vfix_thr, S.vthr, x, y, wxy, vpt, nullptr);
                    if (!interf) drop_links();          // (same basic block as the move in the plain instances)
                }
            }
            if (interf) {
                // the SINR at the new positions needs the SNR sums there before any link can be judged; these are the
                // positions of the step's observation, so the pass emits what depends on them alone (no-move mode: the
                // positions stay, the pass only emits)
                if (valid && !no_move) { S.rec[i].x = x; S.rec[i].y = y; }
                __syncthreads();
                interference_pass(true, step, last);
                __syncthreads();
                if (valid && !no_move) drop_links();
            }
            if (!no_move) tk += 1;                                                     // base.py:454
            __syncthreads();      // bits cleared, slots of the pre-move pass consumed
        }
        // ---- link values at the new position for update_ue_drs_rewards(update_only=True) (base.py:451)
        if (valid) {
            const double iee = dcb_rcp(ewma + DCB_EPSILON);
            const UeInterference tot1 = ue_interf(i);
            int slot = 0;
            for (u64 m = mask; m; m &= m - 1, slot++) {
                const int b = __ffsll((long long)m) - 1;
                if (slot < LC) {
                    Xrow[slot] = link_value(S.share[b], unshared_rate(dist2(S.bsxy[b], x, y), b, tot1), iee);
                    atomicOr(&S.bits[b * NW + (i >> 5)], 1u << (i & 31));
                } else {
                    atomicOr(p.err, DCB_ERRBIT_LINKS);     // observe-only launch on an injected state with too many links
                    mask &= ~((u64)1 << b);
                }
            }
            S.rec[i].mask = mask;
            S.rec[i].x = x; S.rec[i].y = y;
            S.sew[i] = ewma;
            S.smv[i] = make_uint2(wxy, vpt);
        }
        __syncthreads();
        wide_reduce_links<EXT>(S, N, M, NW, LC, p.has_maxcap, warp, lane, nwarps);
        __syncthreads();
        if (valid) {
            // ---- post-move rates -> utility (user.py:76-92)
            const double ee = ewma + DCB_EPSILON;
            int slot = 0;
            for (u64 m = mask; m; m &= m - 1, slot++) {
                const int b = __ffsll((long long)m) - 1;
                const double r = shared_rate(S.share[b], Xrow[slot], S.fac[b], S.arg[b], i, ee);
                if (last && a.out.dbg_link_rate) a.out.dbg_link_rate[u * M + b] = r;
                dr += r;
            }
            util = ue_utility(p, tab, dr);
            S.rec[i].util = util;
            S.rec[i].rb = rb;
            if (EXT) S.sdr[i] = dr;
            // ---- per-UE info outputs (base.py:383-411)
            if (a.out.curr_dr) a.out.curr_dr[(size_t)step * a.out.curr_dr_stride + u] = (float)dr;
            if (a.out.utility) a.out.utility[(size_t)step * a.out.utility_stride + u] = (float)util;
            if (last && a.out.dbg_curr_dr) a.out.dbg_curr_dr[u] = dr;
            if (last && a.out.dbg_utility) a.out.dbg_utility[u] = util;
            if (T > 0 && a.out.lost_conn) a.out.lost_conn[(size_t)step * a.out.lost_conn_stride + u] = (uint8_t)lost;
        }
        __syncthreads();
        if (!central) wide_reduce_utility(S, NA, M, NW, p.reward == DCB_REWARD_MIN, warp, lane, nwarps);
        if (warp == nwarps - 1) {
            // per-env sums: utility (base.py:402) and the central reward over the PRE-move rewards (central.py:65-73)
            double s_u = 0.0, s_r = p.reward == DCB_REWARD_MIN ? CUDART_INF : 0.0;
            for (int j = lane; j < NA; j += 32) {
                s_u += S.rec[j].util;
                const double r = S.rec[j].rb;
                s_r = p.reward == DCB_REWARD_MIN ? (r < s_r ? r : s_r) : s_r + r;
            }
            s_u = warp_sum(s_u);
            s_r = p.reward == DCB_REWARD_MIN ? warp_min(s_r) : warp_sum(s_r);
            if (lane == 0) { S.env_red[0] = s_u; S.env_red[1] = s_r; }
        }
        __syncthreads();
        // ---- clear the bitsets for the next step (all reducers are past them)
        for (int j = tid; j < M * NW; j += blockDim.x) S.bits[j] = 0u;
        if (tid == 0) {
            if (a.out.sum_utility) a.out.sum_utility[(size_t)step * a.out.sum_utility_stride + k] = (float)S.env_red[0];
            if (last && a.out.dbg_sum_utility) a.out.dbg_sum_utility[k] = S.env_red[0];
            if (central && T > 0) {
                double r = S.env_red[1];
                if (p.reward == DCB_REWARD_AVG) r = r / (double)NA;
                if (a.out.reward) a.out.reward[(size_t)step * a.out.reward_stride + k] = (float)r;
                if (last && a.out.dbg_reward) a.out.dbg_reward[k] = r;
            }
        }
        // ---- observation rows and per-UE rewards: one warp per row, lanes over the base stations (two passes cover
        // M <= 64).  variants.py:271-303, central.py:31-57, multi_agent.py:32-95
        float *obs_env = a.out.obs ? a.out.obs + (size_t)step * a.out.obs_stride + (size_t)k * per_env : nullptr;
        const bool want_dbg = last && (a.out.dbg_obs || a.out.dbg_snr);
        // this lane's two base stations: everything that does not depend on the row stays in registers
        const int b0 = lane, b1 = lane + 32;
        const bool ok0 = b0 < M, ok1 = b1 < M;
        const double2 bs0 = ok0 ? S.bsxy[b0] : make_double2(0.0, 0.0), bs1 = ok1 ? S.bsxy[b1] : make_double2(0.0, 0.0);
        float fu0 = 0.0f, fu1 = 0.0f, fa0 = 0.0f, fa1 = 0.0f;
        if (!central) {
            if (ok0) { fu0 = S.f_ues[b0]; fa0 = S.f_util[b0]; }
            if (ok1) { fu1 = S.f_ues[b1]; fa1 = S.f_util[b1]; }
        }
        const unsigned lanebit = 1u << lane;
        // this lane's output cursor: column b0 of the row the warp is at, advanced by nwarps rows per trip; the four
        // (multi) / two (central) segments of a row are fixed offsets from it, the second pass (b1 = b0 + 32) is +128
        // bytes on the same addresses
        const int seg1 = central ? N * M : M, seg2 = 2 * M, seg3 = 3 * M;      // in floats
        const int row_stride_f = central ? M : OW;
        float *o = obs_env + b0 + warp * row_stride_f;                // (never dereferenced when obs_env is NULL)
        const int o_step = nwarps * row_stride_f;
        const long long kN = (long long)k * N;
        // the per-UE rewards of the multi-agent env are NOT computed here: the row phase leaves the in-range set of every
        // UE in its record and a thread per UE folds the per-BS aggregates over it afterwards (one warp instruction serves 32
        // UEs there, against 15 shuffles per row here)
        const bool want_reward = !central && T > 0;
        float *reward_env = (a.out.reward && want_reward) ? a.out.reward + (size_t)step * a.out.reward_stride + kN : nullptr;
        double *dbg_reward_env = (last && a.out.dbg_reward && want_reward) ? a.out.dbg_reward + kN : nullptr;
        for (int r = warp; r < N; r += nwarps, o += o_step) {
            if (PAD && r >= NA) {
                // ---- padding slot (no UE there: max_ues > num_ue): zeros, as central.py:46-55 pads the observation
                const long long ru = kN + r;
                if (EXT && p.obs_var) {
                    const int offs[5] = {p.vo_conn, p.vo_dist, p.vo_dr, p.vo_next, p.vo_ues};
                    double *dbg_env = (last && a.out.dbg_obs) ? a.out.dbg_obs + (size_t)k * per_env : nullptr;
                    for (int sgm = 0; sgm < 5; sgm++) {
                        if (offs[sgm] < 0) continue;
                        for (int b = lane; b < M; b += 32) {
                            if (obs_env) obs_env[(size_t)offs[sgm] + (size_t)r * M + b] = 0.0f;
                            if (dbg_env) dbg_env[(size_t)offs[sgm] + (size_t)r * M + b] = 0.0;
                        }
                    }
                    if (lane == 0 && p.vo_tot >= 0) {
                        if (obs_env) obs_env[(size_t)p.vo_tot + r] = 0.0f;
                        if (dbg_env) dbg_env[(size_t)p.vo_tot + r] = 0.0;
                    }
                } else
                if (obs_env) {
                    if (central) {
                        if (ok0) { o[0] = 0.0f; o[seg1] = 0.0f; }
                        if (ok1) { o[32] = 0.0f; o[seg1 + 32] = 0.0f; }
                        if (lane == 0) obs_env[(size_t)2 * N * M + r] = 0.0f;
                    } else {
                        if (ok0) { o[0] = 0.0f; o[seg1] = 0.0f; o[seg2] = 0.0f; o[seg3] = 0.0f; }
                        if (ok1) { o[32] = 0.0f; o[seg1 + 32] = 0.0f; o[seg2 + 32] = 0.0f; o[seg3 + 32] = 0.0f; }
                        if (lane == 0) o[4 * M] = 0.0f;
                    }
                }
                if (lane == 0) {
                    if (a.out.curr_dr) a.out.curr_dr[(size_t)step * a.out.curr_dr_stride + ru] = 0.0f;
                    if (a.out.utility) a.out.utility[(size_t)step * a.out.utility_stride + ru] = 0.0f;
                    if (T > 0 && a.out.lost_conn) a.out.lost_conn[(size_t)step * a.out.lost_conn_stride + ru] = 0;
                    if (reward_env) reward_env[r] = 0.0f;
                    if (dbg_reward_env) dbg_reward_env[r] = 0.0;
                    if (last && a.out.dbg_curr_dr) a.out.dbg_curr_dr[ru] = 0.0;
                    if (last && a.out.dbg_utility) a.out.dbg_utility[ru] = 0.0;
                }
                if (last && a.out.dbg_obs && !(EXT && p.obs_var)) {
                    if (central) {
                        double *drow = a.out.dbg_obs + (size_t)k * (2 * N * M + N);
                        if (ok0) { drow[r * M + b0] = 0.0; drow[N * M + r * M + b0] = 0.0; }
                        if (ok1) { drow[r * M + b1] = 0.0; drow[N * M + r * M + b1] = 0.0; }
                        if (lane == 0) drow[2 * N * M + r] = 0.0;
                    } else {
                        double *drow = a.out.dbg_obs + (size_t)ru * OW;
                        for (int c = lane; c < OW; c += 32) drow[c] = 0.0;
                    }
                }
                if (last && a.out.dbg_snr) {
                    if (ok0) a.out.dbg_snr[ru * M + b0] = 0.0;
                    if (ok1) a.out.dbg_snr[ru * M + b1] = 0.0;
                }
                continue;
            }
            const UeRec *rp = S.rec + r;
            const double rx = rp->x, ry = rp->y, rutil = rp->util;
            const u64 rmask = rp->mask;
            if (EXT && interf && !p.obs_var) {
                // ---- interference extension with the RelNorm / MaxNorm observation: the row's 'dr' entries, its in-range set
                // and the SINR taps were written by this step's interference pass (they depend on the positions alone);
                // what is left are the segments that depend on the links
                const float c0 = ((unsigned)rmask & lanebit) ? 1.0f : 0.0f;
                const float c1 = ((unsigned)(rmask >> 32) & lanebit) ? 1.0f : 0.0f;
                const double un = rutil * (1.0 / DCB_MAX_UTILITY);
                if (obs_env) {
                    if (central) {
                        if (ok0) o[0] = c0;
                        if (ok1) o[32] = c1;
                        if (lane == 0) obs_env[(size_t)2 * N * M + r] = (float)un;
                    } else {
                        if (ok0) { o[0] = c0; o[seg2] = fu0; o[seg3] = fa0; }
                        if (ok1) { o[32] = c1; o[seg2 + 32] = fu1; o[seg3 + 32] = fa1; }
                        if (lane == 0) o[4 * M] = (float)un;
                    }
                }
                if (last && a.out.dbg_obs) {
                    double *dbg_env = a.out.dbg_obs + (size_t)k * per_env;
                    if (central) {
                        if (ok0) dbg_env[r * M + b0] = (double)c0;
                        if (ok1) dbg_env[r * M + b1] = (double)c1;
                        if (lane == 0) dbg_env[2 * N * M + r] = un;
                    } else {
                        double *drow = dbg_env + (size_t)r * OW;
#pragma unroll
                        for (int q = 0; q < 2; q++) {
                            const int b = q ? b1 : b0;
                            if (b < M) {
                                const int c = S.cnt_obs[b];
                                drow[b] = (double)(q ? c1 : c0);
                                drow[2 * M + b] = (double)c / (double)NA;
                                drow[3 * M + b] = (c > 0 ? S.usum[b] / (double)c : 0.0) / DCB_MAX_UTILITY;
                            }
                        }
                        if (lane == 0) drow[4 * M] = un;
                    }
                }
                continue;
            }
            const double d20 = ok0 ? dist2(bs0, rx, ry) : CUDART_INF;
            const double d21 = ok1 ? dist2(bs1, rx, ry) : CUDART_INF;
            if (EXT && p.obs_var) {
                // ---- general instance: the data-rate observation classes (NormDrMobileEnv / DatarateMobileEnv.get_ue_obs,
                // variants.py:127-250; central layout), on the SNR or -- interference extension -- the SINR
                const UeInterference tot = ue_interf(r);
                const double q0 = ok0 ? (interf ? pair_sinr(snr_of_d2_anywhere(p, tab, k16, d20), b0, tot) : snr_of_d2(p, tab, d20)) : 0.0;
                const double q1 = ok1 ? (interf ? pair_sinr(snr_of_d2_anywhere(p, tab, k16, d21), b1, tot) : snr_of_d2(p, tab, d21)) : 0.0;
                const bool r0in = ok0 && (interf ? q0 > DCB_SNR_THRESHOLD : d20 <= p.thr_d2);
                const bool r1in = ok1 && (interf ? q1 > DCB_SNR_THRESHOLD : d21 <= p.thr_d2);
                const unsigned in0 = __ballot_sync(0xffffffffu, r0in), in1 = __ballot_sync(0xffffffffu, r1in);
                const bool cb0 = ((unsigned)rmask >> lane) & 1u, cb1 = ((unsigned)(rmask >> 32) >> lane) & 1u;
                const long long ru = kN + r;
                double *dbg_env = (last && a.out.dbg_obs) ? a.out.dbg_obs + (size_t)k * per_env : nullptr;
                const double ew = S.sew[r], ee = ew + DCB_EPSILON, iee = dcb_rcp(ee);
                double nx = rx, ny = ry;
                if (p.vo_next >= 0) {
                    const uint2 mv = S.smv[r];
                    const double vf = p.vel_u ? p.vel_u[ru] : p.vel_spec[r];
                    step_towards_waypoint(rx, ry, (double)(mv.x & 0xffffu), (double)(mv.x >> 16),
                                          vf >= 0.0 ? vf : (double)(mv.y & 0xffu), nx, ny);
                }
#pragma unroll
                for (int q = 0; q < 2; q++) {
                    const int b = q ? b1 : b0;
                    if (b >= M) continue;
                    const double d2 = q ? d21 : d20, qual = q ? q1 : q0;
                    const bool inr = q ? r1in : r0in, conn = q ? cb1 : cb0;
                    // Basestation.data_rate (station.py:204-220): the shared rate this UE gets, or would get if it were
                    // counted in (data_rate_shared adds it temporarily, station.py:164-168, 197-201)
                    double rate = 0.0;
                    if (inr) {
                        const int model = S.share[b];
                        const double r0 = DCB_BW * dcb_log2_1p(tab, qual);
                        if (conn) {
                            rate = shared_rate(model, link_value(model, r0, iee), S.fac[b], S.arg[b], r, ee);
                        } else {
                            rate = rate_if_added(model, r0, ee, S.lcnt[b], S.lsum[b], S.lbest[b]);
                        }
                    }
                    const double dro = obs_dr_entry(p, rate);                                       // variants.py:131-141, 213-221
                    const size_t e = (size_t)r * M + b;
                    const double vals[5] = {conn ? 1.0 : 0.0, sqrt(d2) / p.map_diag, dro,
                                            sqrt(dist2(q ? bs1 : bs0, nx, ny)) / p.map_diag, (double)S.lcnt[b]};
                    const int offs[5] = {p.vo_conn, p.vo_dist, p.vo_dr, p.vo_next, p.vo_ues};
#pragma unroll
                    for (int sgm = 0; sgm < 5; sgm++) {
                        if (offs[sgm] < 0) continue;
                        if (obs_env) obs_env[(size_t)offs[sgm] + e] = (float)vals[sgm];
                        if (dbg_env) dbg_env[(size_t)offs[sgm] + e] = vals[sgm];
                    }
                }
                if (lane == 0 && p.vo_tot >= 0) {
                    const double cd = S.sdr[r];
                    const double tot_o = obs_dr_total(p, cd);                                       // variants.py:147-152, 222
                    if (obs_env) obs_env[(size_t)p.vo_tot + r] = (float)tot_o;
                    if (dbg_env) dbg_env[(size_t)p.vo_tot + r] = tot_o;
                }
                if (last && a.out.dbg_snr) {
                    if (ok0) a.out.dbg_snr[ru * M + b0] = q0;
                    if (ok1) a.out.dbg_snr[ru * M + b1] = q1;
                }
                if (want_reward) S.rec[r].inr = ((u64)in1 << 32) | in0;     // SINR-based in-range set for the reward phase
                continue;
            }
            const float f0 = (float)d20, f1 = (float)d21;
            // closest BS: squared distances are >= +0 (or +inf on the lanes past M), so their fp32 bit patterns order like
            // unsigned integers and ONE redux.sync replaces a five-level shuffle ladder
            const float d2minf = __uint_as_float(__reduce_min_sync(0xffffffffu, __float_as_uint(fminf(f0, f1))));
            // in-range set (multi_agent.py:60 / station.py:222-226): exact fp64 decision, gathered by ballot
            const unsigned in0 = __ballot_sync(0xffffffffu, ok0 && d20 <= p.thr_d2);
            const unsigned in1 = __ballot_sync(0xffffffffu, ok1 && d21 <= p.thr_d2);
            if (want_reward) S.rec[r].inr = ((u64)in1 << 32) | in0;   // (every lane stores the same word)
            float dr0, dr1;
            if (EXT && p.obs_maxnorm) {  // MaxNormEnv (variants.py:308-332): per-handle variant, general instance only
                dr0 = ok0 ? max_norm_snr(snr_of_d2_general(p.snr_c0, p.snr_h, tab, d20)) : 0.0f;
                dr1 = ok1 ? max_norm_snr(snr_of_d2_general(p.snr_c0, p.snr_h, tab, d21)) : 0.0f;
            } else if (d2minf >= 1e-6f) {       // 'dr' = snr_b / max_b snr_b (variants.py:276-284) = (d2min / d2_b)^h in fp32
                dr0 = norm_snr_f32(f0, d2minf, hr);
                dr1 = norm_snr_f32(f1, d2minf, hr);
            } else {                     // a UE sitting on a BS: d + EPSILON matters
                double d2min = d20 < d21 ? d20 : d21;
                d2min = warp_min(d2min);
                const double inv_max = dcb_rcp(snr_of_d2(p, tab, d2min));
                dr0 = ok0 ? (float)(snr_of_d2(p, tab, d20) * inv_max) : 0.0f;
                dr1 = ok1 ? (float)(snr_of_d2(p, tab, d21) * inv_max) : 0.0f;
            }
            // 'connected' (variants.py:272): bit b0 of the low word / bit b0 of the high word (b1 = b0 + 32)
            const float c0 = ((unsigned)rmask & lanebit) ? 1.0f : 0.0f;
            const float c1 = ((unsigned)(rmask >> 32) & lanebit) ? 1.0f : 0.0f;
            const double un = rutil * (1.0 / DCB_MAX_UTILITY);                              // variants.py:287
            const long long ru = kN + r;
            if (obs_env) {
                if (central) {
                    float *o1 = o + seg1;
                    if (ok0) { o[0] = c0; o1[0] = dr0; }
                    if (ok1) { o[32] = c1; o1[32] = dr1; }
                    if (lane == 0) obs_env[(size_t)2 * N * M + r] = (float)un;
                } else {
                    float *o1 = o + seg1, *o2 = o + seg2, *o3 = o + seg3;
                    if (ok0) { o[0] = c0; o1[0] = dr0; o2[0] = fu0; o3[0] = fa0; }
                    if (ok1) { o[32] = c1; o1[32] = dr1; o2[32] = fu1; o3[32] = fa1; }
                    if (lane == 0) o[4 * M] = (float)un;
                }
            }
            if (want_dbg) {
                // test taps: fp64 copy of the observation (the 'dr' entries are the fp32 values) and the fp64 SNR
                if (a.out.dbg_obs) {
                    if (central) {
                        double *drow = a.out.dbg_obs + (size_t)k * (2 * N * M + N);
                        if (ok0) { drow[r * M + b0] = (double)c0; drow[N * M + r * M + b0] = (double)dr0; }
                        if (ok1) { drow[r * M + b1] = (double)c1; drow[N * M + r * M + b1] = (double)dr1; }
                        if (lane == 0) drow[2 * N * M + r] = un;
                    } else {
                        double *drow = a.out.dbg_obs + (size_t)ru * OW;
#pragma unroll
                        for (int q = 0; q < 2; q++) {
                            const int b = q ? b1 : b0;
                            if (b < M) {
                                const int c = S.cnt_obs[b];
                                drow[b] = (double)(q ? c1 : c0);
                                drow[M + b] = (double)(q ? dr1 : dr0);
                                drow[2 * M + b] = (double)c / (double)NA;
                                drow[3 * M + b] = (c > 0 ? S.usum[b] / (double)c : 0.0) / DCB_MAX_UTILITY;
                            }
                        }
                        if (lane == 0) drow[4 * M] = un;
                    }
                }
                if (a.out.dbg_snr) {
                    if (ok0) a.out.dbg_snr[ru * M + b0] = snr_of_d2(p, tab, d20);
                    if (ok1) a.out.dbg_snr[ru * M + b1] = snr_of_d2(p, tab, d21);
                }
            }
        }
        __syncthreads();      // rows done: records / aggregates may be overwritten by the next step
        if (want_reward) {
            // ---- per-UE reward of the multi-agent env (multi_agent.py:39-95 -> user.py:231-260) on the POST-move state: a
            // thread per UE folds the per-BS aggregates over the base stations in range of it (UeRec::inr, left by the row phase)
            if (valid) {
                const UeRec me = S.rec[i];
                const u64 inm = me.inr;
                double agg = me.util;
                if (inm) {
                    if (p.reward == DCB_REWARD_AVG) {
                        int nn = 0;
                        double tot = 0.0;
                        for (u64 m = inm; m; m &= m - 1) {
                            const int b = __ffsll((long long)m) - 1;
                            nn += S.cnt_obs[b];
                            tot += S.usum[b];
                        }
                        if (nn > 0) agg = (me.mask == 0 ? tot + me.util : tot) * dcb_rcp((double)(me.mask == 0 ? nn + 1 : nn));
                    } else if (p.reward == DCB_REWARD_SUM) {
                        // user.py:238-244: UEs sharing any BS with this UE; their PRE-move rewards
                        double s = 0.0;
                        for (int j = 0; j < NA; j++)
                            if (S.rec[j].mask & me.mask) s += S.rec[j].rb;
                        agg = s;
                    } else {
                        double mn = CUDART_INF;
                        for (u64 m = inm; m; m &= m - 1) {
                            const double um = S.umin[__ffsll((long long)m) - 1];
                            mn = um < mn ? um : mn;
                        }
                        agg = mn < agg ? mn : agg;
                    }
                }
                if (reward_env) reward_env[i] = (float)agg;
                if (dbg_reward_env) dbg_reward_env[i] = agg;
            }
            // the 'sum' reward reads the other UEs' masks, which their threads rewrite at the top of the next step
            if (p.reward == DCB_REWARD_SUM) __syncthreads();
        }
    }

    // ---- shared memory -> state slabs
    if (T > 0) {
        if (valid) {
            p.pos[u] = make_double2(S.rec[i].x, S.rec[i].y);
            p.mv[u] = S.smv[i];
            p.mask[u] = S.rec[i].mask;
            p.ewma[u] = S.sew[i];
        }
        if (tid == 0) p.time[k] = tk;
    }
}

}  // namespace

#define DCB_WIDE_DISPATCH(pad, ext, central, EXPR)                                                  \
    do {                                                                                            \
        if (central) {                                                                              \
            if (ext) { if (pad) { auto kern = dcb_wide_kernel<true, true, true>; EXPR; } else { auto kern = dcb_wide_kernel<false, true, true>; EXPR; } }     \
            else { if (pad) { auto kern = dcb_wide_kernel<true, false, true>; EXPR; } else { auto kern = dcb_wide_kernel<false, false, true>; EXPR; } }      \
        } else {                                                                                    \
            if (ext) { if (pad) { auto kern = dcb_wide_kernel<true, true, false>; EXPR; } else { auto kern = dcb_wide_kernel<false, true, false>; EXPR; } }   \
            else { if (pad) { auto kern = dcb_wide_kernel<true, false, false>; EXPR; } else { auto kern = dcb_wide_kernel<false, false, false>; EXPR; } }    \
        }                                                                                           \
    } while (0)

cudaError_t dcb_wide_upload_interference_constants(const double *pw10, double snr_c0, double snr_h) {
    double v[12];
    for (int j = 0; j < 10; j++) v[j] = pw10[j];
    v[10] = snr_c0;
    v[11] = snr_h;
    return cudaMemcpyToSymbol(c_interf, v, sizeof(v));
}

cudaError_t dcb_wide_set_smem_limit(size_t smem) {
    cudaError_t e = cudaSuccess;
    for (int v = 0; v < 8 && e == cudaSuccess; v++)
        DCB_WIDE_DISPATCH(v & 1, (v >> 1) & 1, v >> 2,
                          e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    return e;
}

cudaError_t dcb_launch_wide(const StepArgs &a, int threads, int grid, size_t smem, cudaStream_t s) {
    // the general instance (EXT) carries the data-rate observation classes, the interference extension, UniformMovement
    // UEs and the no-move mode; the plain instances -- the measured path -- compile without them
    const bool ext = a.p.obs_var || a.p.interference || a.p.uni_kind || a.p.obs_maxnorm || (a.flags & DCB_STEPF_NO_MOVE);
    const bool pad = a.p.NA < a.p.N;
    DCB_WIDE_DISPATCH(pad, ext, a.p.kind == DCB_KIND_CENTRAL, (kern<<<grid, threads, smem, s>>>(a)));
    return cudaGetLastError();
}
