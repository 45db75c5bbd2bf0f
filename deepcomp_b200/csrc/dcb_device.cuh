// Device-side building blocks shared by the step kernels (dcb_step.cu: fused, pipelined kernel for envs that fit a CTA's
// registers and shared memory; dcb_wide.cu: one CTA per env for large envs): radio model, utility, resource sharing,
// movement, scripted policies.  Citations `file:line` are relative to /root/reference/deepcomp/.
#pragma once

#include <math_constants.h>

#include "dcb_internal.h"
#include "dcb_math.cuh"

namespace {

// [region:helpers.radio]
// ------------------------------------------------------------------------------------------------ radio model
__device__ __forceinline__ double dist2(double2 a, double bx, double by) {   // BS (one 16-byte load) to UE
    const double dx = a.x - bx, dy = a.y - by;
    return dx * dx + dy * dy;
}

// SNR = 10^((30 - c1 - c2 log10(d + EPSILON)) / 10) / 1e-9 = 2^(c0 - 2 h log2(d + EPSILON)),  h = c2 / 20  (station.py:110-127).
// EPSILON = 1e-16 changes d by less than half an ulp for d >= 1 m, so from 1 m on the SNR is a power law in d^2.
#define DCB_NEAR_D2 1.1        // below this squared distance snr can exceed 1/32: general log2(1 + snr)
#define DCB_FAR_D2 65536.0     // dcb_snr_inrange covers binary exponents 0..15 of d^2

// any distance, including d = 0 (a UE that snapped onto a waypoint at a BS position): ~3x the cost of the table form
__device__ __noinline__ double snr_of_d2_general(double c0, double h, const MathTables *tab, double d2) {
    const double d = sqrt(d2) + DCB_EPSILON;
    return dcb_exp2(tab, fma(-2.0 * h, dcb_log2(tab, d), c0));
}
__device__ __forceinline__ double snr_of_d2(const DevParams &p, const MathTables *tab, double d2) {
    if (d2 >= 1.0 && d2 < DCB_FAR_D2) return dcb_snr_inrange(tab, p.pw, d2);
    return snr_of_d2_general(p.snr_c0, p.snr_h, tab, d2);
}

// Unshared rate bw * log2(1 + snr) (station.py:129-138).  In range and not within ~1 m of the BS (every link but a
// handful): table form of the power law, snr < 1/32 -> series in fl(1 + snr) - 1
__device__ __forceinline__ double rate_of_d2_inrange(const DevParams &p, const MathTables *tab, double d2) {
    const double s = dcb_snr_inrange(tab, p.pw, d2);
    return DCB_BW * dcb_log2_1p_small((1.0 + s) - 1.0);
}
__device__ __forceinline__ double rate_of_d2(const DevParams &p, const MathTables *tab, double d2) {
    if (d2 >= DCB_NEAR_D2 && d2 < DCB_FAR_D2) return rate_of_d2_inrange(p, tab, d2);
    return DCB_BW * dcb_log2_1p(tab, snr_of_d2(p, tab, d2));
}

__device__ __forceinline__ double log_utility(const MathTables *tab, double dr) {
    // env/util/utility.py:36-54: clip(10 log10(dr), -20, 20); dr <= 0.01 / >= 100 clip without evaluating the log.
    // Branch-free (selects): two utilities of one UE are evaluated back to back and should share a basic block.
    const double u = 3.0102999566398119521 * dcb_log2(tab, dr);   // 10 log10(2) log2(dr); unused (finite garbage) for dr <= 0.01
    const double c = u > DCB_MAX_UTILITY ? DCB_MAX_UTILITY : (u < DCB_MIN_UTILITY ? DCB_MIN_UTILITY : u);
    return dr <= 0.01 ? DCB_MIN_UTILITY : (dr >= 100.0 ? DCB_MAX_UTILITY : c);
}

// User.dr_to_utility (user.py:81-92): 'log' or 'step' (+-20 around the required rate, utility.py:23-33); 'linear' cannot be
// used with the reference's MIN/MAX_UTILITY = -20/20 (its assert, utility.py:18)
__device__ __forceinline__ double ue_utility(const DevParams &p, const MathTables *tab, double dr) {
    if (p.util_step) return dr >= p.dr_req ? DCB_MAX_UTILITY : DCB_MIN_UTILITY;
    return log_utility(tab, dr);
}

// Value a connected link contributes to its BS's reduction, by sharing model (station.py:170-195):
// resource-fair / max-cap: r0; rate-fair: 1/r0 (:178); proportional-fair: priority r0/(ewma + eps) (:150)
__device__ __forceinline__ double link_value(int model, double r0, double inv_ewma_eps) {
    if (model == DCB_SHARE_RATE_FAIR) return dcb_rcp(r0);
    if (model == DCB_SHARE_PROPORTIONAL_FAIR) return r0 * inv_ewma_eps;
    return r0;
}
// branch-free form for the balanced link loop (both candidates are cheap; selects keep two links in one basic block)
__device__ __forceinline__ double link_value_sel(int model, double r0, double inv_ewma_eps) {
    const double a = dcb_rcp(r0), b = r0 * inv_ewma_eps;
    return model == DCB_SHARE_RATE_FAIR ? a : (model == DCB_SHARE_PROPORTIONAL_FAIR ? b : r0);
}

// Per-(env, BS) factor the reducer leaves behind so that a link's shared rate is a couple of multiplies:
// resource-fair 1/|C_b| (station.py:173), rate-fair 1/sum(1/r0) (:180), proportional-fair 1/(sum(priority) + eps) (:194)
__device__ __forceinline__ double share_factor(int model, int cnt, double sum) {
    const double d = model == DCB_SHARE_RESOURCE_FAIR ? (double)cnt
                                                      : (model == DCB_SHARE_PROPORTIONAL_FAIR ? sum + DCB_EPSILON : sum);
    return dcb_rcp(d);
}

// Shared rate of one link from its value and the BS factor (station.py:152-202)
__device__ __forceinline__ double shared_rate(int model, double v, double fac, int arg, int i, double ewma_eps) {
    if (model == DCB_SHARE_RESOURCE_FAIR) return v * fac;                                       // :173
    if (model == DCB_SHARE_RATE_FAIR) return fac;                                               // :180
    if (model == DCB_SHARE_MAX_CAP) return arg == i ? v : 0.0;                                  // :184-187
    return v * fac * (v * ewma_eps);                                                            // :194-195, r0 = v (ewma + eps)
}

// Basestation.data_rate (station.py:204-220) for a UE that is in range of the BS but NOT connected to it: data_rate_shared
// counts it in temporarily (station.py:164-168, 197-201).  r0 = its unshared rate, ee = its ewma + EPSILON; cnt / sum / best
// = the raw aggregates of the UEs that are connected (count, sum of the link values of link_value(), largest unshared rate)
__device__ __forceinline__ double rate_if_added(int model, double r0, double ee, int cnt, double sum, double best) {
    if (model == DCB_SHARE_RESOURCE_FAIR) return r0 / (double)(cnt + 1);                        // :173
    if (model == DCB_SHARE_RATE_FAIR) return 1.0 / (sum + 1.0 / r0);                            // :178-180
    if (model == DCB_SHARE_MAX_CAP) return (cnt == 0 || r0 > best) ? r0 : 0.0;                  // :184-187 (first arg-max)
    const double pr = r0 / ee;                                                                  // :150, 194-195
    return pr / (sum + pr + DCB_EPSILON) * r0;
}

// The 'dr' / 'dr_total' entries of the data-rate observation classes from a rate (variants.py:131-152, 213-222)
__device__ __forceinline__ double obs_dr_entry(const DevParams &p, double rate) {
    if (p.obs_var == DCB_OBSVAR_NORMDR) return fmin(rate, 100.0) / 100.0;
    if (p.dr_mode == DCB_DR_AUTO) return fmin(rate - p.dr_req, p.dr_req) / p.dr_req;
    if (p.dr_mode == DCB_DR_SUB_REQ) return fmin(rate - p.dr_req, p.dr_cutoff);
    return fmin(rate, p.dr_cutoff);
}
__device__ __forceinline__ double obs_dr_total(const DevParams &p, double curr_dr) {
    if (p.obs_var == DCB_OBSVAR_NORMDR) return fmin(curr_dr, 100.0) / 100.0;
    return fmin(curr_dr - p.dr_req, p.dr_req) / p.dr_req;
}

// The position one step closer to the waypoint (RandomWaypoint.step_towards_waypoint, movement.py:132-156), the UE itself
// is not moved: for the 'next_dist' observation (variants.py:166-169)
__device__ __forceinline__ void step_towards_waypoint(double x, double y, double wx, double wy, double vel, double &nx,
                                                      double &ny) {
    const double vx = wx - x, vy = wy - y;
    if (sqrt(vx * vx + vy * vy) <= vel) { nx = wx; ny = wy; return; }
    const double norm = sqrt(fma(vy, vy, vx * vx));
    nx = x + vel * (vx / norm);
    ny = y + vel * (vy / norm);
}

// Largest squared distance d2 with fl(sqrt(d2)) <= vel: `curr_pos.distance(waypoint) <= velocity` (movement.py:142-145)
// becomes one compare.  fl(sqrt) is monotone, so the set is a down-set and the boundary sits within a few ulps of vel^2.
__device__ __noinline__ double snap_threshold(double vel) {
    if (!(vel > 0.0)) return 0.0;            // sqrt(d2) <= 0  <=>  d2 == 0
    double c = vel * vel;
    for (int it = 0; it < 64 && sqrt(c) > vel; it++) c = __longlong_as_double(__double_as_longlong(c) - 1);
    for (int it = 0; it < 64; it++) {
        const double n = __longlong_as_double(__double_as_longlong(c) + 1);
        if (!(sqrt(n) <= vel)) break;
        c = n;
    }
    return c;
}

// fp32 normalised SNR (variants.py:276-284): (d2min / d2)^h with h = c2/20 = 1.5 + hr
__device__ __forceinline__ float norm_snr_f32(float d2, float d2min, float hr) {
    float rc, lg, sq, ex;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rc) : "f"(d2));
    const float q = d2min * rc;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(sq) : "f"(q));
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(lg) : "f"(q));
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(ex) : "f"(hr * lg));
    return d2 == d2min ? 1.0f : fminf(q * sq * ex, 1.0f);   // the closest BS is exactly 1 (variants.py:284)
}

// MaxNormEnv.get_ue_obs (variants.py:322-330): SNR capped at MAX_SNR_THRESHOLD, minus the connection threshold, scaled so
// that the cap maps to 1 (out-of-range base stations come out slightly negative).  fp64 throughout, rounded once.
__device__ __forceinline__ float max_norm_snr(double snr) {
    const double s = snr < DCB_MAX_SNR_THRESHOLD ? snr : DCB_MAX_SNR_THRESHOLD;
    return (float)((s - DCB_SNR_THRESHOLD) / (DCB_MAX_SNR_THRESHOLD - DCB_SNR_THRESHOLD));
}

template <bool M32> struct MaskType { typedef unsigned long long type; };
template <> struct MaskType<true> { typedef unsigned type; };
__device__ __forceinline__ int mask_ffs(unsigned m) { return __ffs((int)m); }
__device__ __forceinline__ int mask_ffs(unsigned long long m) { return __ffsll((long long)m); }

// ------------------------------------------------------------------------------------------------ scripted policies
// The reference's baseline agents (deepcomp/agent/heuristics.py:13-187, dummy.py:6-50) act per UE on obs['connected'] and
// obs['dr'] = snr_b / max snr.  SNR is a decreasing function of the distance, so "highest dr" is "smallest squared
// distance" (first index on ties, as np.argmax / the agents' loops do) and "dr_b >= eps" is "d2_b <= d2min * gain"
// with gain = eps^(-1/h): the physics warps evaluate the policies exactly, from the state they already hold, one
// step ahead of the observation that the host would have needed.
template <typename mask_t>
__device__ __forceinline__ int policy_action(const PolicyParams &q, mask_t mask, double x, double y, const double2 *bsxy,
                                             int M, int i, long long call_idx, long long u) {
    if (q.kind == DCB_POLICY_FIXED) {                    // dummy.py:25-50
        const long long period = (long long)q.noop_interval + 1;
        return (call_idx % period) == 0 ? q.fixed[i] : 0;
    }
    if (q.kind == DCB_POLICY_RANDOM) {                   // dummy.py:6-22 (uniform over Discrete(M + 1); own counter-based RNG)
        unsigned long long z = q.seed + 0x9E3779B97F4A7C15ull * (unsigned long long)(u + 1) +
                               0xD1B54A32D192ED03ull * (unsigned long long)(call_idx + 1);
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        z ^= z >> 31;
        return (int)(((z >> 32) * (unsigned long long)(M + 1)) >> 32);
    }
    // closest BS overall and closest BS this UE is not linked to (first index on ties)
    double d2min = CUDART_INF, d2free = CUDART_INF;
    int best = 0, best_free = -1;
    for (int b = 0; b < M; b++) {
        const double d2 = dist2(bsxy[b], x, y);
        if (d2 < d2min) { d2min = d2; best = b; }
        if (!((mask >> b) & 1) && d2 < d2free) { d2free = d2; best_free = b; }
    }
    if (q.kind == DCB_POLICY_3GPP) {                     // heuristics.py:19-38
        if ((mask >> best) & 1) return 0;
        if (mask) return mask_ffs(mask);                 // disconnect from the (first) other BS first
        return best + 1;
    }
    if (q.kind == DCB_POLICY_FULLCOMP)                   // heuristics.py:44-65
        return best_free + 1;                            // -1 + 1 = 0 = noop when linked to every BS
    mask_t selected = 0;
    if (q.kind == DCB_POLICY_DYNAMIC) {                  // heuristics.py:86-108: strongest BS and all within eps of it
        // epsilon = 0 selects every BS (heuristics.py:86-91: threshold 0); gain = inf there, and inf * 0 would be NaN for a
        // UE sitting exactly on a BS
        const double thr = isinf(q.gain) ? CUDART_INF : d2min * q.gain;
        for (int b = 0; b < M; b++)
            if (dist2(bsxy[b], x, y) <= thr) selected |= (mask_t)1 << b;
    } else {                                             // heuristics.py:169-187: the static cluster of the strongest BS
        selected = (mask_t)q.cluster[best];
    }
    const mask_t drop = mask & ~selected;
    if (drop) return mask_ffs(drop);                     // leave BS outside the set, lowest index first
    const mask_t want = selected & ~mask;
    if (!want) return 0;
    double d2w = CUDART_INF;                             // join the set, strongest first
    int bw = 0;
    for (mask_t m = want; m; m &= m - 1) {
        const int b = mask_ffs(m) - 1;
        const double d2 = dist2(bsxy[b], x, y);
        if (d2 < d2w) { d2w = d2; bw = b; }
    }
    return bw + 1;
}

// ------------------------------------------------------------------------------------------------ movement
// User.move (user.py:159-173) -> RandomWaypoint.step (movement.py:158-181) for one UE.  State: position (x, y), packed
// waypoint wxy and velocity / pause / table cursor vpt (dcb_internal.h).  vfix >= 0: fixed velocity with snap threshold
// vfix_thr, else the drawn velocity in vpt with its threshold from vthr[] (snap_threshold).
// PREFETCH: `*next_slot` (shared memory) receives the table entry under the cursor by an asynchronous copy (cp.async,
// SASS LDGSTS) issued when the previous entry was consumed -- a redraw happens about once per 50 steps and UE, and a warp
// that waits ~1 us for the table in global memory holds up every warp of its group at the next barrier.  A load into a
// register would do the same on paper, but its scoreboard is shared with later shared-memory loads and stalls them.
__device__ __forceinline__ void prefetch_table_entry(uint32_t *slot, const uint32_t *src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((unsigned)__cvta_generic_to_shared(slot)), "l"(src)
                 : "memory");
}
__device__ __forceinline__ void prefetch_wait() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// UniformMovement.step (movement.py:66-80) for one UE: constant (move_x, move_y) per step; when the next point would not
// be strictly inside the map (Point.within: a point on the border is outside) BOTH components flip sign and the step is
// taken with the flipped vector, without a second check.  kx, ky: component kinds (DevParams::uni_kind), vx, vy: the fixed
// numbers; drawn magnitudes and the flip bit live in the packed movement word.
__device__ __forceinline__ void ue_move_uniform(const DevParams &p, int kx, int ky, double vx, double vy, double &x,
                                                double &y, unsigned wxy, unsigned &vpt) {
    double mx = kx == 1 ? vx : (double)(wxy & 0xffffu);
    double my = ky == 1 ? vy : (double)(wxy >> 16);
    if (vpt & 0x8000u) { mx = -mx; my = -my; }
    double nx = x + mx, ny = y + my;
    if (!(0.0 < nx && nx < p.map_w && 0.0 < ny && ny < p.map_h)) {
        vpt ^= 0x8000u;
        mx = -mx; my = -my;
        nx = x + mx; ny = y + my;
    }
    x = nx; y = ny;
}

template <bool PREFETCH>
__device__ __forceinline__ void ue_move(const DevParams &p, long long u, double vfix, double vfix_thr, const double *vthr,
                                        double &x, double &y, unsigned &wxy, unsigned &vpt, uint32_t *next_slot) {
    double wx = (double)(wxy & 0xffffu), wy = (double)(wxy >> 16);
    unsigned pause = (vpt >> 8) & 0xffu;
    bool moving = true;
    if (x == wx && y == wy) pause |= 0x80u;                            // movement.py:169-170
    if (pause & 0x80u) {
        if ((int)(pause & 0x7fu) < p.pause_duration) {                 // movement.py:174-176
            pause++;
            moving = false;
        } else {                                                       // movement.py:177 -> reset()
            unsigned tidx = vpt >> 16;
            uint32_t e;
            if ((int)tidx >= p.D) {
                atomicOr(p.err, DCB_ERRBIT_TABLE);
                tidx = p.D - 1;
                e = p.table[u * p.D + tidx];
            } else if (PREFETCH) {
                prefetch_wait();
                e = *next_slot;
            } else {
                e = p.table[u * p.D + tidx];
            }
            if (PREFETCH && (int)(tidx + 1) < p.D) prefetch_table_entry(next_slot, p.table + u * p.D + tidx + 1);
            wxy = (e & 0x3fffu) | (((e >> 14) & 0x3fffu) << 16);
            vpt = (e >> 28) | ((tidx + 1) << 16);
            wx = (double)(wxy & 0xffffu); wy = (double)(wxy >> 16);
            pause = 0;
        }
    }
    vpt = (vpt & 0xffff00ffu) | (pause << 8);
    if (moving) {
        // movement.py:132-156; `distance <= velocity` as a compare of the squared distance (snap_threshold)
        const bool drawn = vfix < 0.0;
        const double vel = drawn ? (double)(vpt & 0xffu) : vfix;
        const double snap = drawn ? vthr[vpt & 0xfu] : vfix_thr;
        const double vx = wx - x, vy = wy - y;
        if (vx * vx + vy * vy <= snap) {
            x = wx; y = wy;
        } else {
            const double norm = sqrt(fma(vy, vy, vx * vx));   // np.linalg.norm -> FMA-accumulating ddot
            x = x + vel * (vx / norm);
            y = y + vel * (vy / norm);
        }
    }
}

// MobileEnv.reset of one UE's state (base.py:169-189) from its pre-drawn initial position / first table entry
__device__ __forceinline__ void ue_reset(const DevParams &p, long long u, double &x, double &y, unsigned &wxy,
                                         unsigned &vpt) {
    const double2 ps = p.init_pos[u];
    x = ps.x; y = ps.y;
    const uint32_t e = p.table[u * p.D];
    wxy = (e & 0x3fffu) | (((e >> 14) & 0x3fffu) << 16);
    vpt = (e >> 28) | (1u << 16);
}

}  // namespace
