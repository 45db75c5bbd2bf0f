// Brute-force search support: the reward of EVERY joint action of one env, tried on its
// current state -- deepcomp/agent/brute_force.py:59-94 (BruteForceAgent.compute_action) over
// deepcomp/env/single_ue/base.py:284-313 (MobileEnv.test_ue_actions).  The reference tries the (M+1)^N candidates one
// after the other on the live env (apply, update, revert); here the candidates are the parallel axis: one thread per
// candidate, the env state (unshared rates of all N x M pairs, in-range sets, masks, EWMA rates) staged once per CTA in
// shared memory.  All file:line citations are relative to /root/reference/deepcomp/.
#include "dcb_device.cuh"

namespace {

#define DCB_BRUTE_MAX_UE 16

__global__ void __launch_bounds__(256) dcb_brute_kernel(BruteArgs a) {
    extern __shared__ __align__(16) unsigned char smem[];
    const DevParams &p = a.p;
    const int N = p.NA, M = p.M;
    MathTables *tab = reinterpret_cast<MathTables *>(smem);
    double *r0 = reinterpret_cast<double *>(smem + sizeof(MathTables));          // [N][M] unshared rates (station.py:129-138)
    double *ewma0 = r0 + N * M;                                                   // [N]
    unsigned long long *mask0 = reinterpret_cast<unsigned long long *>(ewma0 + N);   // [N] current links
    unsigned long long *inr = mask0 + N;                                          // [N] base stations in range
    int *share = reinterpret_cast<int *>(inr + N);                                // [M]
    dcb_math_init(tab, nullptr, threadIdx.x, blockDim.x, p.tabs);
    __syncthreads();
    const long long base = (long long)a.env * p.N;
    for (int i = threadIdx.x; i < N; i += blockDim.x) {
        const double2 ps = p.pos[base + i];
        unsigned long long rng = 0ull;
        for (int b = 0; b < M; b++) {
            const double d2 = dist2(make_double2(p.bs_xy[2 * b], p.bs_xy[2 * b + 1]), ps.x, ps.y);
            r0[i * M + b] = rate_of_d2(p, tab, d2);
            if (d2 <= p.thr_d2) rng |= 1ull << b;                                  // can_connect, station.py:222-226
        }
        inr[i] = rng;
        mask0[i] = p.mask[base + i];
        ewma0[i] = p.ewma[base + i];
    }
    for (int b = threadIdx.x; b < M; b += blockDim.x) share[b] = p.sharing[b];
    __syncthreads();

    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= a.count) return;
    // ---- the candidate: digits of its number in base M + 1, most significant digit = first UE (brute_force.py:26-62);
    // apply_ue_actions (base.py:247-282): toggle the link, connecting only in range (user.py:190-229)
    unsigned long long m[DCB_BRUTE_MAX_UE];
    double ew[DCB_BRUTE_MAX_UE], dr[DCB_BRUTE_MAX_UE];
    long long c = a.first + idx;
    for (int i = N - 1; i >= 0; i--) {
        const int act = (int)(c % (M + 1));
        c /= (M + 1);
        unsigned long long mm = mask0[i];
        if (act > 0) {
            const unsigned long long bit = 1ull << (act - 1);
            if (mm & bit) mm &= ~bit;
            else if (inr[i] & bit) mm |= bit;
        }
        m[i] = mm;
        ew[i] = ewma0[i];
    }
    // ---- base.py:294-299: rates with the old EWMA -> update_ewma_dr (user.py:148-157) -> rates and rewards again
    for (int pass = 0; pass < 2; pass++) {
        for (int i = 0; i < N; i++) dr[i] = 0.0;
        for (int b = 0; b < M; b++) {
            const int model = share[b];
            int cnt = 0, arg = 0x7fffffff;
            double sum = 0.0, best = 0.0;
            for (int i = 0; i < N; i++) {
                if ((m[i] >> b) & 1ull) {
                    const double v = link_value(model, r0[i * M + b], dcb_rcp(ew[i] + DCB_EPSILON));
                    cnt++;
                    sum += v;
                    if (v > best) { best = v; arg = i; }                           // station.py:184: first arg-max
                }
            }
            if (cnt == 0) continue;
            const double fac = share_factor(model, cnt, sum);
            for (int i = 0; i < N; i++) {
                if ((m[i] >> b) & 1ull) {
                    const double ee = ew[i] + DCB_EPSILON;
                    dr[i] += shared_rate(model, link_value(model, r0[i * M + b], dcb_rcp(ee)), fac, arg, i, ee);
                }
            }
        }
        if (pass == 0)
            for (int i = 0; i < N; i++) ew[i] = 0.9 * dr[i] + (1 - 0.9) * ew[i];
    }
    // ---- calc_reward (base.py:158-167) per UE, central step_reward (central.py:65-73)
    double agg = p.reward == DCB_REWARD_MIN ? CUDART_INF : 0.0;
    for (int i = 0; i < N; i++) {
        const double r = ue_utility(p, tab, dr[i]) / DCB_MAX_UTILITY;
        agg = p.reward == DCB_REWARD_MIN ? (r < agg ? r : agg) : agg + r;
    }
    if (p.reward == DCB_REWARD_AVG) agg = agg / (double)N;
    a.rewards[idx] = agg;
}

}  // namespace

cudaError_t dcb_launch_brute(const BruteArgs &a, cudaStream_t s) {
    const int N = a.p.NA, M = a.p.M;
    const size_t smem = sizeof(MathTables) + sizeof(double) * (N * M + N) + sizeof(unsigned long long) * 2 * N + sizeof(int) * M;
    const long long grid = (a.count + 255) / 256;
    if (grid > 0x7fffffffLL) return cudaErrorInvalidValue;
    dcb_brute_kernel<<<(unsigned)grid, 256, smem, s>>>(a);
    return cudaGetLastError();
}
