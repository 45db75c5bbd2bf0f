// Fused env.step kernel: T consecutive steps of K independent env instances in one launch.
//
// Restates deepcomp/env/single_ue/base.py:413-466 (MobileEnv.step) with the observation / reward variants of
// deepcomp/env/multi_ue/central.py:143-152 and deepcomp/env/multi_ue/multi_agent.py:6-107.  All file:line
// citations are relative to /root/reference/deepcomp/.
//
// Mapping.  A CTA owns E consecutive envs and runs TWO warp groups of G = ceil32(E*N) threads each; thread t of
// either group owns UE slot t of the CTA's E*N UEs, which are contiguous in every [K][N] state slab (coalesced
// state loads / reward stores):
//   * physics warps  -- carry the per-UE state (position, waypoint, pause counter, connection bitmask, EWMA rate) in
//     registers across all T steps and run the state-changing chain of a step;
//   * observer warps -- turn the state a step left behind into that step's observation tile, rewards and info.
// The groups are pipelined one step apart through a double-buffered hand-off in shared memory and named barriers
// (FULL / EMPTY per buffer parity): while the observers emit step t, the physics warps already run step t+1.  A
// 1024-env batch gives a B200 only ~11 UE-warps per SM, so the step is latency-bound; the split halves the
// per-warp instruction stream and doubles the warps in flight without redundant work.
//
// Physics, one step (after the first):
//   1. pre-move rates from the aggregates computed during the PREVIOUS step (the UE positions of "after move t"
//      and "before move t+1" are the same; only the action's one toggled link differs, so the reducer produces both
//      sets of per-BS aggregates in one pass)                                  base.py:446, station.py:152-220
//   2. move, drop out-of-range links, EWMA                                     user.py:148-188, movement.py:132-181
//   3. sparse loop over the (few) connected links: fp64 SNR -> unshared rate -> link value -> matrix X and
//      per-(env, BS) UE bitsets                                               station.py:129-150
//   4. reduce per (env, BS) over the CONNECTED UEs only (bitset walk): count / sum of link values / arg-max, for
//      the current mask and for the next step's mask; fixed order (deterministic)
//   5. post-move rates -> utility -> hand-off                                  base.py:451, user.py:76-92
// Observers, one step: per-BS utility sums (bitset walk), dense pair loop over the M base stations (squared
// distance, in-range bit, normalised SNR in fp32), observation tile, rewards, info; the tile leaves through ONE
// TMA bulk store (cp.async.bulk shared -> global) per CTA and step.           variants.py:271-303, multi_agent.py:39-95
// The first step of a launch (and a step that starts with an episode reset) has no aggregates to inherit and
// computes them stand-alone.
//
// Arithmetic.  Positions and every range decision are fp64 with the reference's operation order (no FMA
// contraction: the library is built with --fmad=false; the one FMA the reference has, inside np.linalg.norm, is
// explicit), so trajectories, connection masks and lost-connection counts are bit-exact.  Everything that feeds
// rates, utilities and rewards is fp64 through the table-driven log2 / exp2 of dcb_math.cuh (<= 1e-13 relative to
// the reference's libm chain); a UE closer than ~1 m to a BS -- where the reference's `distance + EPSILON` matters
// -- takes the libm path.  The observation entry 'dr' = snr_b / max_b snr_b (variants.py:276-284) is a float32
// output that feeds nothing else; it is evaluated as (d2min/d2_b)^h in fp32 (q * sqrt(q) * 2^((h - 1.5) log2 q),
// MUFU lg2 / ex2 / rcp / sqrt), <= 1e-6 relative against the reference (north star: 1e-5).
#include "dcb_device.cuh"

static_assert(sizeof(MathTables) == 5 * 16 * 8, "SmemLayout reserves 5 x 128 B for the math tables");

#if defined(DCB_TRACE) && DCB_STEP_CLASS == 704
#define DCB_TRACE_ON 1
// Phase timeline of CTA 0 (A/B builds only, -DDCB_TRACE; the 704-thread class = the headline shape): clock64 of lane 0 of every warp at the phase boundaries of
// steps 40..47 of a launch: [role][warp][step - 40][point]
static __device__ long long dcb_trace_buf[2 * 16 * 8 * 8];
static __device__ long long dcb_trace_cta[3 * 4096];     // per CTA: globaltimer (ns) at entry / exit, SM id
__device__ __forceinline__ long long dcb_globaltimer() {
    long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ unsigned dcb_smid() {
    unsigned r;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(r));
    return r;
}
#define DCB_TRACE_PT(role, pt)                                                                              \
    do {                                                                                                    \
        if (blockIdx.x == 0 && (t & 31) == 0 && step >= 40 && step < 48)                                    \
            dcb_trace_buf[(((role) * 16 + (t >> 5)) * 8 + (step - 40)) * 8 + (pt)] = clock64();             \
    } while (0)
#else
#define DCB_TRACE_PT(role, pt) do { } while (0)
#endif

#ifndef DCB_OBS_UNROLL
#define DCB_OBS_UNROLL 2   // unroll factor of the observers' pair loops
#endif
#define DCB_PRAGMA(x) _Pragma(#x)
#define DCB_UNROLL(n) DCB_PRAGMA(unroll n)

namespace {

// [region:helpers.barriers]
// ------------------------------------------------------------------------------------------------ named barriers
// Barrier 0 is __syncthreads (set-up only).  Each warp group has a private barrier; FULL[parity] / EMPTY[parity]
// hand a buffer from the physics warps to the observer warps and back (arrive on one side, sync on the other).
enum { BAR_PHYS = 1, BAR_OBS = 2, BAR_FULL = 3, BAR_EMPTY = 5 };

__device__ __forceinline__ void bar_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void bar_arrive(int id, int n) {
    asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(n) : "memory");
}
__device__ __forceinline__ bool bar_or(int id, int n, bool pred) {
    int r;
    asm volatile(
        "{\n\t.reg .pred p, q;\n\tsetp.ne.b32 q, %3, 0;\n\tbar.red.or.pred p, %1, %2, q;\n\tselp.b32 %0, 1, 0, p;\n\t}"
        : "=r"(r) : "r"(id), "r"(n), "r"((int)pred) : "memory");
    return r != 0;
}

// ------------------------------------------------------------------------------------------------ reductions
// (tid, gsize: thread index within / size of the calling warp group)
// [region:reduce_links]
// For every (env, BS) pair walk the bitset of connected UEs: count, sum of link values X[ue][bs], first arg-max
// (max-cap only), folded into the pair's sharing factor (share_factor).  Done for two bitsets (current masks -> *_a,
// next step's masks -> *_b).  S lanes per pair take the pair's chunks round-robin -- a chunk is a 32-UE bitset word
// or, when there are more lanes than words, a 16- / 8-bit piece of one (CS = log2 chunks per word); fixed
// combination order -> deterministic.
__device__ __forceinline__ void reduce_links(int tid, int gsize, const double *X, const unsigned *bits_a,
                                             const unsigned *bits_b, int N, int M, int MS, int n_env, int n_pairs,
                                             const unsigned short *porder, int S, int CS, bool want_arg, const int *share,
                                             double *fac_a, int *arg_a, double *fac_b, int *arg_b, int *raw_cnt = nullptr,
                                             double *raw_sum = nullptr, double *raw_best = nullptr) {
    const int R = n_env * M;
    const int NW = (N + 31) >> 5;
    const int ls = 31 - __clz(S);                 // S is a power of two
    const int ppp = gsize >> ls;
    const int seg = tid & (S - 1);
    const float inv_m = 1.0f / (float)M;
    for (int base = 0; base < n_pairs; base += ppp) {
        // pairs are taken in the order of `porder`: resource-fair pairs first, so that (almost) every warp works on one
        // sharing model -- and the warps of the resource-fair pairs only count bits
        const int slot = base + (tid >> ls);
        const int pair = slot < n_pairs ? (int)porder[slot] : R;
        const bool ok = pair < R;                 // (a CTA at the end of the batch holds fewer than E envs)
        int c0 = 0, c1 = 0, a0 = 0x7fffffff, a1 = 0x7fffffff, model = 0;
        double s0 = 0.0, s1 = 0.0, b0 = 0.0, b1 = 0.0;
        bool walk = false;
        if (ok) {
            const int le = __float2int_rz(((float)pair + 0.5f) * inv_m), b = pair - le * M;   // exact: pair < 2^16
            model = share[b];
            walk = want_arg || model != DCB_SHARE_RESOURCE_FAIR;
            const double *col = X + (size_t)(le * N) * MS + b;
            const int cb = 32 >> CS;                                   // bits per chunk
            const unsigned cmask = 0xffffffffu >> (32 - cb);
            for (int c = seg; c < (NW << CS); c += S) {
                const int w = c >> CS, sh = (c & ((1 << CS) - 1)) * cb;
                const unsigned wa = (bits_a[pair * NW + w] >> sh) & cmask;
                const unsigned wb = (bits_b[pair * NW + w] >> sh) & cmask;
                c0 += __popc(wa);
                c1 += __popc(wb);
                unsigned both = walk ? (wa | wb) : 0u;
                const int ibase = (w << 5) + sh;
                while (both) {
                    // two UEs per trip: both loads are in flight before the (ordered) accumulation
                    const int j = __ffs(both) - 1;
                    both &= both - 1;
                    const bool two = both != 0u;
                    const int j2 = two ? __ffs(both) - 1 : j;
                    both &= both - 1;
                    const double v = col[(ibase + j) * MS];
                    const double v2 = col[(ibase + j2) * MS];
                    const bool ina = (wa >> j) & 1u, inb = (wb >> j) & 1u;
                    const bool ina2 = two && ((wa >> j2) & 1u), inb2 = two && ((wb >> j2) & 1u);
                    if (ina) s0 += v;
                    if (inb) s1 += v;
                    if (ina2) s0 += v2;
                    if (inb2) s1 += v2;
                    if (want_arg) {               // max-cap only (station.py:184): first arg-max
                        if (ina && v > b0) { b0 = v; a0 = ibase + j; }
                        if (inb && v > b1) { b1 = v; a1 = ibase + j; }
                        if (ina2 && v2 > b0) { b0 = v2; a0 = ibase + j2; }
                        if (inb2 && v2 > b1) { b1 = v2; a1 = ibase + j2; }
                    }
                }
            }
        }
        const bool any_walk = __any_sync(0xffffffffu, walk);           // warp-uniform: sums only where some lane has one
        for (int off = S >> 1; off > 0; off >>= 1) {
            c0 += __shfl_xor_sync(0xffffffffu, c0, off);
            c1 += __shfl_xor_sync(0xffffffffu, c1, off);
            if (any_walk) {
                s0 += __shfl_xor_sync(0xffffffffu, s0, off);
                s1 += __shfl_xor_sync(0xffffffffu, s1, off);
            }
            if (want_arg) {
                double ob = __shfl_xor_sync(0xffffffffu, b0, off);
                int oi = __shfl_xor_sync(0xffffffffu, a0, off);
                if (ob > b0 || (ob == b0 && oi < a0)) { b0 = ob; a0 = oi; }
                ob = __shfl_xor_sync(0xffffffffu, b1, off);
                oi = __shfl_xor_sync(0xffffffffu, a1, off);
                if (ob > b1 || (ob == b1 && oi < a1)) { b1 = ob; a1 = oi; }
            }
        }
        if (ok && seg == 0) {
            fac_a[pair] = share_factor(model, c0, s0);
            fac_b[pair] = share_factor(model, c1, s1);
            if (want_arg) { arg_a[pair] = a0; arg_b[pair] = a1; }
            if (raw_cnt) { raw_cnt[pair] = c0; raw_sum[pair] = s0; raw_best[pair] = b0; }   // data-rate observation classes
        }
    }
}

// [region:reduce_utility]
// Observer-side aggregates, computed by every observer warp for itself (no barrier between observer warps): for the
// n_le envs [le0, le0 + n_le) that the warp's 32 rows belong to and every BS -- connected count -> cnt, total utility
// (station.py:63-69) -> usum, min (station.py:78-83) -> umin, and the two per-BS observation entries (variants.py:296-299,
// station.py:71-76) -> f_ues, f_util.  Lane q walks the UE bitset of pair q in UE order; warps that share an env
// compute bit-identical values.
__device__ __forceinline__ void warp_reduce_utility(int lane, const unsigned *bits, const double *su, int N, int NA,
                                                    int M, int le0, int n_le, bool want_min, int *cnt, double *usum,
                                                    double *umin, float *f_ues, float *f_util) {
    const int NW = (N + 31) >> 5;
    const int PW = n_le * M;
    const double inv_n = 1.0 / (double)NA;       // self.num_ue = UEs present (variants.py:296)
    const float inv_m = 1.0f / (float)M;
    for (int q = lane; q < PW; q += 32) {
        const int ll = __float2int_rz(((float)q + 0.5f) * inv_m), b = q - ll * M;   // exact: q < 2^16
        const int le = le0 + ll;
        const unsigned *pb = bits + (le * M + b) * NW;
        const double *sue = su + le * N;
        double s = 0.0, mn = DCB_MAX_UTILITY;
        int c = 0;
        if (want_min) {
            for (int w = 0; w < NW; w++) {
                unsigned wa = pb[w];
                c += __popc(wa);
                while (wa) {
                    const int j = __ffs(wa) - 1;
                    wa &= wa - 1;
                    const double uu = sue[(w << 5) + j];
                    s += uu;
                    mn = uu < mn ? uu : mn;
                }
            }
        } else {
            for (int w = 0; w < NW; w++) {
                unsigned wa = pb[w];
                c += __popc(wa);
                while (wa) {
                    // two UEs per trip: both loads in flight before the (ordered) additions
                    const int j = __ffs(wa) - 1;
                    wa &= wa - 1;
                    const bool two = wa != 0u;
                    const int j2 = two ? __ffs(wa) - 1 : j;
                    wa &= wa - 1;
                    const double uu = sue[(w << 5) + j], uu2 = sue[(w << 5) + j2];
                    s += uu;
                    if (two) s += uu2;
                }
            }
        }
        cnt[q] = c;
        usum[q] = s;
        umin[q] = mn;
        f_ues[q] = (float)((double)c * inv_n);                                             // |C_b| / N (variants.py:296)
        f_util[q] = c > 0 ? (float)(s * dcb_rcp((double)c) * (1.0 / DCB_MAX_UTILITY)) : 0.0f;
    }
}


// [region:reduce_env]
// Reduction of one env's per-UE vector by one warp: mode 0 = sum, 2 = min; every lane gets the result
__device__ __forceinline__ double warp_reduce_env(int lane, const double *v, int N, int mode) {
    const int chunk = (N + 31) / 32;
    const int i0 = lane * chunk, i1 = min(N, i0 + chunk);
    double s = mode == 2 ? CUDART_INF : 0.0;
    for (int i = i0; i < i1; i++) s = mode == 2 ? fmin(s, v[i]) : s + v[i];
    for (int off = 16; off > 0; off >>= 1) {
        const double o = __shfl_xor_sync(0xffffffffu, s, off);
        s = mode == 2 ? fmin(s, o) : s + o;
    }
    return s;
}

// [region:kernel.setup]
// ------------------------------------------------------------------------------------------------ the kernel
// M32: all connection / in-range masks fit 32 bits (n_bs <= 32) -- halves the integer work on the mask paths
// PAD: the general instance -- the envs may have padding slots (NA < N, variable UE population) and the observation may be
// a per-handle variant (MaxNorm); the common fixed-population RelNorm case compiles without the extra compares, the
// padding branch of the observers and the variant branch
// CENTRAL: observation / reward layout of the central agent (central.py:31-73) instead of the per-UE multi-agent one
// (multi_agent.py:32-95) -- a compile-time choice: it selects offsets, store paths and reward code all over the observers
template <int MAXT, bool M32, bool PAD, bool CENTRAL>
__device__ __forceinline__ void dcb_step_body(const StepArgs &a) {
    using mask_t = typename MaskType<M32>::type;
    extern __shared__ __align__(128) unsigned char smem[];
    const DevParams &p = a.p;
    const int N = p.N, M = p.M, E = p.E, S = p.S;
    const int MS = row_stride(M);
    const int NW = (N + 31) >> 5;
    const SmemLayout &L = a.L;      // computed on the host: offsets come straight from the constant bank
    MathTables *tab = reinterpret_cast<MathTables *>(smem + L.off_tab);
    float *stage = reinterpret_cast<float *>(smem + L.off_stage);
    double *X = reinterpret_cast<double *>(smem + L.off_x);
    double *fac_pre = reinterpret_cast<double *>(smem + L.off_fac_pre);
    int *arg_pre = reinterpret_cast<int *>(smem + L.off_arg_pre);
    double *fac_post = reinterpret_cast<double *>(smem + L.off_fac_post);
    int *arg_post = reinterpret_cast<int *>(smem + L.off_arg_post);
    double *hx = reinterpret_cast<double *>(smem + L.off_hx);
    double *hy = reinterpret_cast<double *>(smem + L.off_hy);
    unsigned long long *hmask = reinterpret_cast<unsigned long long *>(smem + L.off_hmask);
    double *hutil = reinterpret_cast<double *>(smem + L.off_hutil);
    double *hrb = reinterpret_cast<double *>(smem + L.off_hrb);
    double *hdr = reinterpret_cast<double *>(smem + L.off_hdr);
    int *hlost = reinterpret_cast<int *>(smem + L.off_hlost);
    double2 *bsxy = reinterpret_cast<double2 *>(smem + L.off_bsx);   // [M] interleaved (off_bsx, off_bsy are adjacent)
    int *share = reinterpret_cast<int *>(smem + L.off_share);
    double *velspec = reinterpret_cast<double *>(smem + L.off_vel);
    unsigned *bits_post3 = reinterpret_cast<unsigned *>(smem + L.off_bits);   // [3][nbits]
    unsigned *bits_pre2 = bits_post3 + 3 * L.nbits;                            // [2][nbits]
    unsigned *bits_fresh = bits_pre2 + 2 * L.nbits;                            // [nbits]
    unsigned short *links = reinterpret_cast<unsigned short *>(smem + L.off_links);
    double *vthr = reinterpret_cast<double *>(smem + L.off_vthr);
    uint32_t *snext = reinterpret_cast<uint32_t *>(smem + L.off_snext);
    unsigned short *porder = reinterpret_cast<unsigned short *>(smem + L.off_porder);
    // general instance, data-rate observation classes (dcb_set_obs_variant): per step parity, for the observers
    const bool obs_var = PAD && p.obs_var != 0;
    const int EMv = E * M;
    double *vfac = reinterpret_cast<double *>(smem + L.off_vfac);
    double *rsum = reinterpret_cast<double *>(smem + L.off_rsum);
    double *rbest = reinterpret_cast<double *>(smem + L.off_rbest);
    int *varg = reinterpret_cast<int *>(smem + L.off_varg);
    int *rcnt = reinterpret_cast<int *>(smem + L.off_rcnt);
    double *hewma = reinterpret_cast<double *>(smem + L.off_hewma);
    uint2 *hmv = reinterpret_cast<uint2 *>(smem + L.off_hmv);

    const int G = blockDim.x >> 1;                  // threads per warp group
    const bool is_obs = (int)threadIdx.x >= G;
    const int t = (int)threadIdx.x - (is_obs ? G : 0);
    const int env0 = blockIdx.x * E;
    const int n_env = min(E, p.K - env0);
    const int EN = E * N;
    const bool in_cta = t < n_env * N;              // a UE slot of one of this CTA's envs
    const int le = in_cta ? t / N : 0;
    const int i = in_cta ? t - le * N : 0;
    const int NA = PAD ? p.NA : N;                  // slots [0, NA) hold UEs, the rest is padding (max_ues > num_ue)
    const bool valid = PAD ? (in_cta && i < NA) : in_cta;
    const int k = env0 + le;
    const long long u = (long long)k * N + i;
    constexpr bool central = CENTRAL;
    const int OW = central ? 2 * M + 1 : 4 * M + 1;
    const int T = a.T;
    const int n_iter = T > 0 ? T : 1;

    // the physics threads' state and first action: issued ahead of the table / parameter loads of the prologue, so that a
    // launch pays ONE global-memory latency before its first step instead of three in a row (tables, state, action)
    double2 ld_ps = make_double2(0.0, 0.0);
    uint2 ld_mv = make_uint2(0u, 0u);
    unsigned long long ld_mask = 0;
    double ld_ewma = 0.0;
    int ld_tk = 0, ld_act = 0;
    if (!is_obs && valid) {
        ld_ps = p.pos[u];
        ld_mv = p.mv[u];
        ld_mask = p.mask[u];
        ld_ewma = p.ewma[u];
        ld_tk = p.time[k];
        if (T > 0 && a.actions) ld_act = a.actions[u];
    }
    dcb_math_init(tab, vthr, threadIdx.x, blockDim.x, p.tabs);
    for (int b = threadIdx.x; b < M; b += blockDim.x) {
        bsxy[b] = make_double2(p.bs_xy[2 * b], p.bs_xy[2 * b + 1]);
        share[b] = p.sharing[b];
    }
    for (int j = threadIdx.x; j < N; j += blockDim.x) velspec[j] = p.vel_spec[j];
    for (int j = threadIdx.x; j < E * M; j += blockDim.x) porder[j] = p.pair_order[j];
    for (int j = threadIdx.x; j < 6 * L.nbits; j += blockDim.x) bits_post3[j] = 0u;
    __syncthreads();

#ifdef DCB_TRACE_ON
    if (threadIdx.x == 0 && blockIdx.x < 4096) dcb_trace_cta[3 * blockIdx.x] = dcb_globaltimer();
#endif
    if (!is_obs) {
// [region:P.load]
        // ===================================================================== physics warps
        double x = 0, y = 0, ewma = 0;
        mask_t mask = 0;
        unsigned wxy = 0, vpt = 0;
        int tk = 0;
        x = ld_ps.x; y = ld_ps.y;
        wxy = ld_mv.x; vpt = ld_mv.y;
        mask = (mask_t)ld_mask;
        ewma = ld_ewma;
        tk = ld_tk;
        // the waypoint-table entry under the cursor, fetched ahead of its use (ue_move)
        uint32_t *next_slot = snext + t;
        if (valid && (int)(vpt >> 16) < p.D) prefetch_table_entry(next_slot, p.table + u * p.D + (vpt >> 16));
        const double vfix = valid ? (p.vel_u ? p.vel_u[u] : velspec[i]) : 0.0;
        const double vfix_thr = vfix >= 0.0 ? snap_threshold(vfix) : 0.0;
        // general instance only: UniformMovement UEs (movement.py:26-80) and the no-move launch mode (DCB_STEPF_NO_MOVE)
        int ukx = 0, uky = 0;
        double uvx = 0.0, uvy = 0.0;
        if (PAD && p.uni_kind && valid) {
            ukx = p.uni_kind[2 * i]; uky = p.uni_kind[2 * i + 1];
            uvx = p.uni_val[2 * i]; uvy = p.uni_val[2 * i + 1];
        }
        const bool no_move = PAD && (a.flags & DCB_STEPF_NO_MOVE);
        double *Xrow = X + (size_t)t * MS;
        // this UE's bit in the per-(env, BS) UE bitsets: word index (low 24 bits) and bit number (high 8 bits)
        const int bit_word = le * M * NW + (i >> 5);
        const int bit_info = bit_word | ((i & 31) << 24);
        const int lane = t & 31;
        unsigned short *wl = links + (t >> 5) * L.links_per_warp;   // this warp's link list
        double *Xwarp = X + (size_t)(t & ~31) * MS;
        // carried from the end of one step to the next (only meaningful when the next step is not fresh):
        mask_t mask_next = 0;     // mask after the next step's action
        double rb_next = 0.0;     // the next step's reward before the move (base.py:446)
        // ... which only the central reward (central.py:65-73) and the multi-agent 'sum' reward (multi_agent.py:79-86) read
        const bool need_rb = central || p.reward == DCB_REWARD_SUM;
        // this UE's action of the current step (running pointer: no 64-bit index arithmetic per step)
        const size_t act_stride = (size_t)p.K * N;
        const int32_t *act_row = a.actions ? a.actions + u : nullptr;
        bool any_fresh = true;    // some env of this CTA starts the step without inherited aggregates (CTA-uniform)
        int rot = 0;              // step % 3: the post-move UE bitsets rotate over three buffers (see the clear below)

        for (int step = 0; step < n_iter; step++) {
            const bool last = step == n_iter - 1;
            const int par = step & 1;
            unsigned *bits_post = bits_post3 + rot * L.nbits;
            unsigned *bits_pre = bits_pre2 + par * L.nbits;
            rot = rot == 2 ? 0 : rot + 1;              // now (step + 1) % 3
            double rb = rb_next;
            int lost = 0;
            // next step's action: issued now so that the global-load latency hides behind this step's work
            int act_next = 0;
            if (valid && T > 0 && !last && !a.pol.kind) act_next = act_row[act_stride];
            act_row += act_stride;                     // -> actions of step + 1
// [region:P.top+fresh]
            // the observers must be done with this parity's hand-off buffers (step - 2)
            if (step >= 2) bar_sync(BAR_EMPTY + par, 2 * G);
            DCB_TRACE_PT(0, 0);
            if (T > 0) {
                if (any_fresh) {
                    // ---- stand-alone pre phase: first step of the launch or a step that starts with an episode
                    // reset in some env of this CTA (every env of the CTA recomputes; same arithmetic, same values)
                    bool fresh = step == 0;
                    if (valid && p.auto_reset && tk >= p.episode_length) {
                        // MobileEnv.reset before the next step (base.py:169-189)
                        ue_reset(p, u, x, y, wxy, vpt);
                        prefetch_wait();                    // (an older copy into the slot must land first)
                        prefetch_table_entry(next_slot, p.table + u * p.D + 1);   // cursor is 1 after a reset; D >= 3
                        mask = 0; ewma = 0.0; tk = 0;
                        fresh = true;
                    }
                    if (valid) {
                        if (fresh) {
                            // apply_ue_actions (base.py:247-282) -> User.connect_to_bs(disconnect=True) (user.py:190-229)
                            int act;
                            if (a.pol.kind) {
                                act = policy_action<mask_t>(a.pol, mask, x, y, bsxy, M, i, a.pol.call0 + step, u);
                                if (a.actions_out) a.actions_out[(size_t)step * p.K * N + u] = act;
                            } else {
                                // (the launch's first action came in with the state; a later fresh step -- an on-device
                                // episode reset -- reads its own)
                                act = step == 0 ? ld_act : *(act_row - act_stride);
                            }
                            if (act < 0 || act > M) {
                                atomicOr(p.err, DCB_ERRBIT_ACTION);
                            } else if (act > 0) {
                                const int b = act - 1;
                                const mask_t bit = (mask_t)1 << b;
                                if (mask & bit) mask &= ~bit;
                                else if (dist2(bsxy[b], x, y) <= p.thr_d2) mask |= bit;   // can_connect, station.py:222-226
                            }
                        } else {
                            mask = mask_next;
                        }
                        const double iee = dcb_rcp(ewma + DCB_EPSILON);
                        for (mask_t m = mask; m; m &= m - 1) {
                            const int b = mask_ffs(m) - 1;
                            Xrow[b] = link_value(share[b], rate_of_d2(p, tab, dist2(bsxy[b], x, y)), iee);
                            atomicOr(&bits_fresh[bit_word + b * NW], 1u << (i & 31));
                        }
                    }
                    bar_sync(BAR_PHYS, G);
                    reduce_links(t, G, X, bits_fresh, bits_fresh, N, M, MS, n_env, E * M, porder, S, p.CS, p.has_maxcap, share, fac_post,
                                 arg_post, fac_pre, arg_pre);
                    bar_sync(BAR_PHYS, G);
                    for (int j = t; j < L.nbits; j += G) bits_fresh[j] = 0u;
                    if (valid) {
                        // ---- update_ue_drs_rewards (base.py:315-335): Basestation.data_rate_shared (station.py:152-202)
                        // per connected link; the ue.bs_dr cache goes back into Xrow; calc_reward (base.py:158-167),
                        // penalties are identically 0 (base.py:257)
                        const double ee = ewma + DCB_EPSILON;
                        double dr = 0.0;
                        for (mask_t m = mask; m; m &= m - 1) {
                            const int b = mask_ffs(m) - 1;
                            const int pr = le * M + b;
                            const double r = shared_rate(share[b], Xrow[b], fac_pre[pr], arg_pre[pr], i, ee);
                            Xrow[b] = r;
                            dr += r;                                                   // user.py:64-69
                        }
                        if (need_rb) rb = ue_utility(p, tab, dr) * (1.0 / DCB_MAX_UTILITY);
                    }
                } else {
                    mask = mask_next;
                }
// [region:P.move]
                if (valid && !(PAD && no_move)) {
                    DCB_TRACE_PT(0, 1);
                    if (PAD && ukx) ue_move_uniform(p, ukx, uky, uvx, uvy, x, y, wxy, vpt);
                    else ue_move<true>(p, u, vfix, vfix_thr, vthr, x, y, wxy, vpt, next_slot);
// [region:P.drop+ewma]
                    // ---- check_bs_connection (user.py:175-188) + update_ewma_dr (user.py:148-157); Xrow holds the
                    // pre-move shared rates (ue.bs_dr)
                    double keep = 0.0;
                    for (mask_t m = mask; m; m &= m - 1) {
                        const int b = mask_ffs(m) - 1;
                        if (dist2(bsxy[b], x, y) <= p.thr_d2) keep += Xrow[b];
                        else { mask &= ~((mask_t)1 << b); lost++; }
                    }
                    ewma = 0.9 * keep + (1 - 0.9) * ewma;
                    tk += 1;                                                           // base.py:454
                }
            }
// [region:P.next_action]
            // ---- the next step's action toggles one link (user.py:190-229): known now, so the link values and the
            // reduction below serve update_ue_drs_rewards(update_only=True) of this step (base.py:451) AND the
            // pre-move update of the next step (base.py:446)
            mask_next = mask;
            if (valid && T > 0 && !last) {
                int act = act_next;
                if (a.pol.kind) {
                    act = policy_action<mask_t>(a.pol, mask, x, y, bsxy, M, i, a.pol.call0 + step + 1, u);
                    if (a.actions_out) a.actions_out[(size_t)(step + 1) * p.K * N + u] = act;
                }
                if (act < 0 || act > M) {
                    atomicOr(p.err, DCB_ERRBIT_ACTION);
                } else if (act > 0) {
                    const int b = act - 1;
                    const mask_t bit = (mask_t)1 << b;
                    if ((mask & bit) || dist2(bsxy[b], x, y) <= p.thr_d2) mask_next = mask ^ bit;
                }
            }
            DCB_TRACE_PT(0, 2);
// [region:P.links]
            // ---- link values at the new position, balanced over the warp: the lanes' links (1.6 on average, up to M)
            // are compacted into a per-warp list and dealt out round-robin, LW per lane and trip in one basic block
            {
                const mask_t un = valid ? (mask | mask_next) : (mask_t)0;
                const int n = M32 ? __popc((unsigned)un) : __popcll((unsigned long long)un);
                int incl = n;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const int o = __shfl_up_sync(0xffffffffu, incl, d);
                    if (lane >= d) incl += o;
                }
                const int total = __shfl_sync(0xffffffffu, incl, 31);
                int pos = incl - n;
                for (mask_t m = un; m; m &= m - 1) {
                    const int b = mask_ffs(m) - 1;
                    const unsigned fl = ((unsigned)(mask >> b) & 1u) | (((unsigned)(mask_next >> b) & 1u) << 1);
                    wl[pos++] = (unsigned short)(lane | (b << 5) | (fl << 11));
                }
                __syncwarp();
                const double iee = dcb_rcp(ewma + DCB_EPSILON);
#ifndef DCB_LW
#define DCB_LW 1     // links per lane and trip of the balanced link loop (2 and 3 measured slower: 1.66e8 / 1.53e8 vs 1.81e8)
#endif
                constexpr int LW = DCB_LW;
                for (int base = 0; base < total; base += 32 * LW) {
                    unsigned e[LW];
                    double d2[LW], v[LW], oi[LW];
                    int ow[LW], bi[LW];
#pragma unroll
                    for (int q = 0; q < LW; q++) {
                        const int idx = base + q * 32 + lane;
                        e[q] = idx < total ? (unsigned)wl[idx] | 0x8000u : (unsigned)lane;
                        ow[q] = e[q] & 31;
                        const double ox = __shfl_sync(0xffffffffu, x, ow[q]);
                        const double oy = __shfl_sync(0xffffffffu, y, ow[q]);
                        oi[q] = __shfl_sync(0xffffffffu, iee, ow[q]);
                        bi[q] = __shfl_sync(0xffffffffu, bit_info, ow[q]);
                        const int b = (e[q] >> 5) & 63;
                        d2[q] = dist2(bsxy[b], ox, oy);
                        v[q] = link_value_sel(share[b], rate_of_d2_inrange(p, tab, d2[q]), oi[q]);
                    }
#pragma unroll
                    for (int q = 0; q < LW; q++) {
                        if (e[q] & 0x8000u) {
                            const int b = (e[q] >> 5) & 63;
                            double val = v[q];
                            if (d2[q] < DCB_NEAR_D2 || d2[q] >= DCB_FAR_D2)   // within ~1 m of the BS (far: never in range)
                                val = link_value(share[b], rate_of_d2(p, tab, d2[q]), oi[q]);
                            Xwarp[(size_t)ow[q] * MS + b] = val;
                            const int bw = (bi[q] & 0xffffff) + b * NW;
                            const unsigned bv = 1u << ((unsigned)bi[q] >> 24);
                            if (e[q] & (1u << 11)) atomicOr(&bits_post[bw], bv);
                            if (e[q] & (1u << 12)) atomicOr(&bits_pre[bw], bv);
                        }
                    }
                }
            }
// [region:P.reduce_phase]
            // barrier + "does any env of this CTA reset before the next step?" in one bar.red
            DCB_TRACE_PT(0, 3);
            any_fresh = bar_or(BAR_PHYS, G, valid && T > 0 && p.auto_reset && tk >= p.episode_length);
            DCB_TRACE_PT(0, 4);
            reduce_links(t, G, X, bits_post, bits_pre, N, M, MS, n_env, E * M, porder, S, p.CS, p.has_maxcap, share, fac_post, arg_post,
                         fac_pre, arg_pre, obs_var ? rcnt + par * EMv : nullptr, rsum + par * EMv, rbest + par * EMv);
            bar_sync(BAR_PHYS, G);
            if (obs_var)       // the observers read this step's factors one step later: a copy per parity
                for (int j = t; j < EMv; j += G) { vfac[par * EMv + j] = fac_post[j]; varg[par * EMv + j] = arg_post[j]; }
            DCB_TRACE_PT(0, 5);
            // bits_pre of this parity is consumed; its next use is two steps (>= 2 group barriers) away.  The post
            // bitsets of step - 2 were read by the observers, who are done with that step (EMPTY wait above); that
            // buffer, (step + 1) % 3, is the one the next step fills
            for (int j = t; j < L.nbits; j += G) {
                bits_pre[j] = 0u;
                if (step >= 2) bits_post3[rot * L.nbits + j] = 0u;
            }
            if (valid) {
// [region:P.rates+handoff]
                // ---- post-move rates of this step and pre-move rates of the next one in ONE pass over the links;
                // the next step's ue.bs_dr cache goes back into Xrow.  utility (user.py:76-92), reward (base.py:158-167)
                const double ee = ewma + DCB_EPSILON;
                double dr = 0.0, dr_pre = 0.0;
                for (mask_t m = mask | mask_next; m; m &= m - 1) {
                    const int b = mask_ffs(m) - 1;
                    const int pr = le * M + b;
                    const int model = share[b];
                    const double v = Xrow[b];
                    const double r_post = shared_rate(model, v, fac_post[pr], arg_post[pr], i, ee);
                    const double r_pre = shared_rate(model, v, fac_pre[pr], arg_pre[pr], i, ee);
                    if ((mask >> b) & 1) {
                        if (last && a.out.dbg_link_rate) a.out.dbg_link_rate[u * M + b] = r_post;
                        dr += r_post;
                    }
                    if ((mask_next >> b) & 1) dr_pre += r_pre;                         // user.py:64-69
                    Xrow[b] = r_pre;
                }
                const double util = ue_utility(p, tab, dr);
                if (need_rb) rb_next = ue_utility(p, tab, dr_pre) * (1.0 / DCB_MAX_UTILITY);
                const int h = par * EN + t;
                hx[h] = x; hy[h] = y; hmask[h] = mask;
                hutil[h] = util;
                hrb[h] = rb; hdr[h] = dr; hlost[h] = lost;
                if (obs_var) { hewma[h] = ewma; hmv[h] = make_uint2(wxy, vpt); }
            }
            bar_arrive(BAR_FULL + par, 2 * G);
            DCB_TRACE_PT(0, 7);
        }
// [region:P.drain+store]
        // ---- registers -> state slabs, while the observers are still busy with the last step (they never read the slabs)
        if (valid && T > 0) {
            p.pos[u] = make_double2(x, y);
            p.mv[u] = make_uint2(wxy, vpt);
            p.mask[u] = (unsigned long long)mask;
            p.ewma[u] = ewma;
            if (i == 0) p.time[k] = tk;
        }
        // drain: the observers' last (up to two) EMPTY arrivals
        if (n_iter >= 2) bar_sync(BAR_EMPTY + (n_iter & 1), 2 * G);
        bar_sync(BAR_EMPTY + ((n_iter - 1) & 1), 2 * G);

#ifdef DCB_TRACE_ON
        if (threadIdx.x == 0 && blockIdx.x < 4096) {
            dcb_trace_cta[3 * blockIdx.x + 1] = dcb_globaltimer();
            dcb_trace_cta[3 * blockIdx.x + 2] = dcb_smid();
        }
#endif
    } else {
// [region:O.setup]
        // ===================================================================== observer warps
        // obs tile row of this UE: multi [connected(M) | dr(M) | ues_at_bs(M) | util_at_bs(M) | utility(1)] per UE
        // (variants.py:271-303); central [connected(N*M) | dr(N*M) | utility(N)] per env (central.py:31-57)
        const int row_off = central ? le * (2 * N * M + N) + i * M : t * OW;
        const float hr = (float)(p.snr_h - 1.5);
        const bool dbg_any = a.out.dbg_obs || a.out.dbg_snr;
        // obs tile -> global: ONE TMA bulk store (cp.async.bulk shared -> global) per CTA and step.  TMA wants 16-byte
        // aligned addresses and sizes; the CTA's span of the observation buffer starts at an arbitrary multiple of 4
        // bytes, so the tile is built in shared memory at the same offset modulo 16 and the (< 16 byte) head and tail
        // are written with scalar stores.
        const size_t per_env = obs_var ? (size_t)p.var_obs_size : (central ? (size_t)(2 * N * M + N) : (size_t)N * OW);
        const unsigned tile_bytes = (unsigned)(per_env * n_env * 4);
        // multi: a warp's 32 rows are one contiguous span of 128 * OW bytes (a multiple of 16 from the tile start), so
        // every warp stores its own span and the observer warps never wait for each other
        const bool warp_store = !central;
        const int lane = t & 31;
        const int w0 = t & ~31;
        const int wrows = max(0, min(32, n_env * N - w0));
        // envs this warp's rows belong to, and its private aggregate block (warp_reduce_utility)
        const int le0 = w0 / N;
        const int n_le = wrows > 0 ? (w0 + wrows - 1) / N - le0 + 1 : 0;
        const int le_s0 = (w0 + N - 1) / N;          // first env whose UE 0 is one of this warp's rows
        unsigned char *wagg = smem + L.off_wagg + (t >> 5) * L.wagg_stride;
        double *usum_o = reinterpret_cast<double *>(wagg);
        double *umin_o = usum_o + L.wagg_pairs;
        int *cnt_o = reinterpret_cast<int *>(umin_o + L.wagg_pairs);
        float *f_ues = reinterpret_cast<float *>(cnt_o + L.wagg_pairs);
        float *f_util = f_ues + L.wagg_pairs;
        const int lq = (le - le0) * M;               // this UE's env within the block
        const bool want_env_rew = central && T > 0;
        const bool want_env_sumu = a.out.sum_utility || a.out.dbg_sum_utility;
        int orot = 0;                                 // step % 3 (the physics warps' rotating post bitsets)
        // production outputs of the current step as running pointers (no 64-bit index arithmetic per step)
        float *obs_step = a.out.obs ? a.out.obs + (size_t)env0 * per_env : nullptr;
        float *rew_step = a.out.reward ? a.out.reward + (central ? (long long)k : u) : nullptr;
        uint8_t *lost_step = a.out.lost_conn ? a.out.lost_conn + u : nullptr;

        for (int step = 0; step < n_iter; step++) {
            const bool last = step == n_iter - 1;
            const int par = step & 1;
            const int hbase = par * EN;
            float *dst = obs_step;
            float *rew_dst = rew_step;
            uint8_t *lost_dst = lost_step;
            if (obs_step) obs_step += a.out.obs_stride;
            if (rew_step) rew_step += a.out.reward_stride;
            if (lost_step) lost_step += a.out.lost_conn_stride;
            const unsigned mis = (unsigned)((size_t)dst & 15);
            float *tile = stage + (mis >> 2);
            float *row_conn = tile + row_off;
            float *row_dr = central ? row_conn + N * M : row_conn + M;
// [region:O.full_wait]
            bar_sync(BAR_FULL + par, 2 * G);
            DCB_TRACE_PT(1, 0);
            // the previous step's TMA store(s) must have finished reading the tile before anyone rewrites it
            if (warp_store) {
                if (lane == 0 && step > 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            } else {
                if (t == 0 && step > 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                bar_sync(BAR_OBS, G);
            }
// [region:O.util_reduce]
            const unsigned *bits_post = bits_post3 + orot * L.nbits;
            orot = orot == 2 ? 0 : orot + 1;
            if (!central)
                warp_reduce_utility(lane, bits_post, hutil + hbase, N, NA, M, le0, n_le, p.reward == DCB_REWARD_MIN, cnt_o,
                                    usum_o, umin_o, f_ues, f_util);
            double env_rew_v = 0.0, env_sumu_v = 0.0;          // of the env whose UE 0 this thread is
            if (want_env_rew || want_env_sumu) {
                for (int es = le_s0; es < n_env && es * N < w0 + 32; es++) {
                    double r1 = 0.0, r2 = 0.0;
                    if (want_env_rew) r1 = warp_reduce_env(lane, hrb + hbase + es * N, NA, p.reward == DCB_REWARD_MIN ? 2 : 0);
                    if (want_env_sumu) r2 = warp_reduce_env(lane, hutil + hbase + es * N, NA, 0);
                    if (valid && le == es && i == 0) { env_rew_v = r1; env_sumu_v = r2; }
                }
            }
            __syncwarp();
            DCB_TRACE_PT(1, 1);
            if (valid) {
                const int h = hbase + t;
                const double x = hx[h], y = hy[h], util = hutil[h], dr = hdr[h];
                const mask_t mask = (mask_t)hmask[h];
// [region:O.dense]
                mask_t inrange = 0;
                if (obs_var) {
                    // ---- data-rate observation classes (general instance, central layout; NormDrMobileEnv /
                    // DatarateMobileEnv.get_ue_obs, variants.py:127-250): per BS the shared rate this UE gets or would get
                    // (station.py:204-220), from this step's per-(env, BS) aggregates (kept per parity by the physics
                    // warps); written straight to the observation buffer, one segment per key (central.py:31-57)
                    const double ee = hewma[h] + DCB_EPSILON, iee = dcb_rcp(ee);
                    float *oenv = dst ? dst + (size_t)le * per_env : nullptr;
                    double *denv = (last && a.out.dbg_obs) ? a.out.dbg_obs + (size_t)k * per_env : nullptr;
                    double nx = x, ny = y;
                    if (p.vo_next >= 0) {
                        const uint2 mv = hmv[h];
                        const double vf = p.vel_u ? p.vel_u[u] : velspec[i];
                        step_towards_waypoint(x, y, (double)(mv.x & 0xffffu), (double)(mv.x >> 16),
                                              vf >= 0.0 ? vf : (double)(mv.y & 0xffu), nx, ny);
                    }
                    const int offs[5] = {p.vo_conn, p.vo_dist, p.vo_dr, p.vo_next, p.vo_ues};
                    for (int b = 0; b < M; b++) {
                        const double d2 = dist2(bsxy[b], x, y);
                        const bool conn = (mask >> b) & 1;
                        const int pr = par * EMv + le * M + b;
                        double rate = 0.0;
                        if (d2 <= p.thr_d2) {
                            const int model = share[b];
                            const double r0 = rate_of_d2(p, tab, d2);
                            rate = conn ? shared_rate(model, link_value(model, r0, iee), vfac[pr], varg[pr], i, ee)
                                        : rate_if_added(model, r0, ee, rcnt[pr], rsum[pr], rbest[pr]);
                        }
                        const double vals[5] = {conn ? 1.0 : 0.0, sqrt(d2) / p.map_diag, obs_dr_entry(p, rate),
                                                sqrt(dist2(bsxy[b], nx, ny)) / p.map_diag, (double)rcnt[pr]};
                        const size_t e = (size_t)i * M + b;
                        for (int sgm = 0; sgm < 5; sgm++) {
                            if (offs[sgm] < 0) continue;
                            if (oenv) oenv[(size_t)offs[sgm] + e] = (float)vals[sgm];
                            if (denv) denv[(size_t)offs[sgm] + e] = vals[sgm];
                        }
                        if (last && a.out.dbg_snr) a.out.dbg_snr[u * M + b] = snr_of_d2(p, tab, d2);
                    }
                    if (p.vo_tot >= 0) {
                        const double tot_o = obs_dr_total(p, dr);
                        if (oenv) oenv[(size_t)p.vo_tot + i] = (float)tot_o;
                        if (denv) denv[(size_t)p.vo_tot + i] = tot_o;
                    }
                } else {
                // ---- dense pass A: squared distances (fp64, exact range decision multi_agent.py:60 /
                // station.py:222-226), parked in the tile as float for pass B
                float d2minf = CUDART_INF_F;
                DCB_UNROLL(DCB_OBS_UNROLL)
                for (int b = 0; b < M; b++) {
                    const double d2 = dist2(bsxy[b], x, y);
                    const float d2f = (float)d2;
                    d2minf = fminf(d2minf, d2f);            // float conversion is monotone: the min commutes with it
                    if (d2 <= p.thr_d2) inrange |= (mask_t)1 << b;
                    row_dr[b] = d2f;
                }
                // ---- dense pass B: 'dr' = snr_b / max_b snr_b (variants.py:276-284) = (d2min / d2_b)^h
                if (PAD && p.obs_maxnorm) {   // MaxNormEnv (variants.py:308-332): per-handle variant (general instance only), out-of-line fp64 SNR
                    for (int b = 0; b < M; b++)
                        row_dr[b] = max_norm_snr(snr_of_d2_general(p.snr_c0, p.snr_h, tab, dist2(bsxy[b], x, y)));
                } else if (d2minf >= 1e-6f) {        // below: d + EPSILON matters (in practice d = 0 exactly)
                DCB_UNROLL(DCB_OBS_UNROLL)
                    for (int b = 0; b < M; b++) row_dr[b] = norm_snr_f32(row_dr[b], d2minf, hr);
                } else {
                    double d2min = CUDART_INF;
                    for (int b = 0; b < M; b++) {
                        const double d2 = dist2(bsxy[b], x, y);
                        d2min = d2 < d2min ? d2 : d2min;
                    }
                    const double inv_max = dcb_rcp(snr_of_d2(p, tab, d2min));
                    for (int b = 0; b < M; b++)
                        row_dr[b] = (float)(snr_of_d2(p, tab, dist2(bsxy[b], x, y)) * inv_max);
                }
                DCB_TRACE_PT(1, 2);
// [region:O.staging]
                // ---- rest of the observation row
                const double un = util * (1.0 / DCB_MAX_UTILITY);                      // variants.py:287
                if (central) {
                    for (int b = 0; b < M; b++) row_conn[b] = (float)((unsigned)(mask >> b) & 1u);
                    tile[(size_t)le * (2 * N * M + N) + 2 * N * M + i] = (float)un;
                } else {
                    const float *fu = f_ues + lq, *fa = f_util + lq;
                    for (int b = 0; b < M; b++) {
                        row_conn[b] = (float)((unsigned)(mask >> b) & 1u);
                        row_conn[2 * M + b] = fu[b];
                        row_conn[3 * M + b] = fa[b];
                    }
                    row_conn[4 * M] = (float)un;
                }
                if (last && dbg_any) {
                    // test taps: fp64 copy of the observation (the 'dr' entries are the fp32 values) and the fp64
                    // SNR of every pair (station.py:122-127)
                    if (a.out.dbg_obs) {
                        if (central) {
                            double *drow = a.out.dbg_obs + (size_t)k * (2 * N * M + N);
                            for (int b = 0; b < M; b++) {
                                drow[i * M + b] = (double)((unsigned)(mask >> b) & 1u);
                                drow[N * M + i * M + b] = (double)row_dr[b];
                            }
                            drow[2 * N * M + i] = un;
                        } else {
                            double *drow = a.out.dbg_obs + (size_t)u * OW;
                            for (int b = 0; b < M; b++) {
                                const int c = cnt_o[lq + b];
                                drow[b] = (double)((unsigned)(mask >> b) & 1u);
                                drow[M + b] = (double)row_dr[b];
                                drow[2 * M + b] = (double)c / (double)NA;
                                drow[3 * M + b] = (c > 0 ? usum_o[lq + b] / (double)c : 0.0) / DCB_MAX_UTILITY;
                            }
                            drow[4 * M] = un;
                        }
                    }
                    if (a.out.dbg_snr)
                        for (int b = 0; b < M; b++)
                            a.out.dbg_snr[u * M + b] = snr_of_d2(p, tab, dist2(bsxy[b], x, y));
                }
                }
                DCB_TRACE_PT(1, 3);
// [region:O.outputs+reward]
                // ---- per-UE outputs and rewards -> global
                if (a.out.curr_dr) a.out.curr_dr[(size_t)step * a.out.curr_dr_stride + u] = (float)dr;
                if (a.out.utility) a.out.utility[(size_t)step * a.out.utility_stride + u] = (float)util;
                if (last && a.out.dbg_curr_dr) a.out.dbg_curr_dr[u] = dr;
                if (last && a.out.dbg_utility) a.out.dbg_utility[u] = util;
                if (i == 0) {
                    if (a.out.sum_utility)
                        a.out.sum_utility[(size_t)step * a.out.sum_utility_stride + k] = (float)env_sumu_v;
                    if (last && a.out.dbg_sum_utility) a.out.dbg_sum_utility[k] = env_sumu_v;
                }
                if (T > 0) {
                    if (lost_dst) *lost_dst = (uint8_t)hlost[h];
                    if (central) {
                        if (i == 0) {
                            // central.py:65-73 over the PRE-move rewards
                            double r = env_rew_v;
                            if (p.reward == DCB_REWARD_AVG) r = r / (double)NA;
                            if (rew_dst) *rew_dst = (float)r;
                            if (last && a.out.dbg_reward) a.out.dbg_reward[k] = r;
                        }
                    } else {
                        // multi_agent.py:39-95 on the POST-move state
                        double agg = util;
                        if (inrange) {
                            if (p.reward == DCB_REWARD_AVG) {
                                int nn = 0;
                                double tot = 0.0;
                                for (mask_t m = inrange; m; m &= m - 1) {
                                    const int b = mask_ffs(m) - 1;
                                    nn += cnt_o[lq + b];
                                    tot += usum_o[lq + b];
                                }
                                if (nn > 0) agg = (mask == 0 ? tot + util : tot) * dcb_rcp((double)(mask == 0 ? nn + 1 : nn));
                            } else if (p.reward == DCB_REWARD_SUM) {
                                // user.py:238-244: UEs sharing any BS with this UE; their PRE-move rewards
                                agg = 0.0;
                                for (int j = 0; j < NA; j++)
                                    if ((mask_t)hmask[hbase + le * N + j] & mask) agg += hrb[hbase + le * N + j];
                            } else {
                                for (mask_t m = inrange; m; m &= m - 1) {
                                    const int b = mask_ffs(m) - 1;
                                    agg = fmin(agg, umin_o[lq + b]);
                                }
                            }
                        }
                        if (rew_dst) *rew_dst = (float)agg;
                        if (last && a.out.dbg_reward) a.out.dbg_reward[u] = agg;
                    }
                }
            } else if (PAD && in_cta && obs_var) {
                // ---- padding slot under a data-rate observation class: zeros in every segment (central.py:46-55)
                float *oenv = dst ? dst + (size_t)le * per_env : nullptr;
                double *denv = (last && a.out.dbg_obs) ? a.out.dbg_obs + (size_t)k * per_env : nullptr;
                const int offs[5] = {p.vo_conn, p.vo_dist, p.vo_dr, p.vo_next, p.vo_ues};
                for (int sgm = 0; sgm < 5; sgm++) {
                    if (offs[sgm] < 0) continue;
                    for (int b = 0; b < M; b++) {
                        if (oenv) oenv[(size_t)offs[sgm] + (size_t)i * M + b] = 0.0f;
                        if (denv) denv[(size_t)offs[sgm] + (size_t)i * M + b] = 0.0;
                    }
                }
                if (p.vo_tot >= 0) {
                    if (oenv) oenv[(size_t)p.vo_tot + i] = 0.0f;
                    if (denv) denv[(size_t)p.vo_tot + i] = 0.0;
                }
                if (a.out.curr_dr) a.out.curr_dr[(size_t)step * a.out.curr_dr_stride + u] = 0.0f;
                if (a.out.utility) a.out.utility[(size_t)step * a.out.utility_stride + u] = 0.0f;
                if (T > 0 && lost_dst) *lost_dst = 0;
                if (last) {
                    if (a.out.dbg_curr_dr) a.out.dbg_curr_dr[u] = 0.0;
                    if (a.out.dbg_utility) a.out.dbg_utility[u] = 0.0;
                    if (a.out.dbg_snr)
                        for (int b = 0; b < M; b++) a.out.dbg_snr[u * M + b] = 0.0;
                }
            } else if (PAD && in_cta) {
                // ---- padding slot (no UE there: max_ues > num_ue): zeros, as central.py:46-55 pads the observation
                for (int b = 0; b < M; b++) {
                    row_conn[b] = 0.0f;
                    row_dr[b] = 0.0f;
                    if (!central) { row_conn[2 * M + b] = 0.0f; row_conn[3 * M + b] = 0.0f; }
                }
                if (central) tile[(size_t)le * (2 * N * M + N) + 2 * N * M + i] = 0.0f;
                else row_conn[4 * M] = 0.0f;
                if (a.out.curr_dr) a.out.curr_dr[(size_t)step * a.out.curr_dr_stride + u] = 0.0f;
                if (a.out.utility) a.out.utility[(size_t)step * a.out.utility_stride + u] = 0.0f;
                if (T > 0) {
                    if (lost_dst) *lost_dst = 0;
                    if (!central && rew_dst) *rew_dst = 0.0f;
                }
                if (last) {
                    if (a.out.dbg_curr_dr) a.out.dbg_curr_dr[u] = 0.0;
                    if (a.out.dbg_utility) a.out.dbg_utility[u] = 0.0;
                    if (!central && T > 0 && a.out.dbg_reward) a.out.dbg_reward[u] = 0.0;
                    if (a.out.dbg_obs) {
                        if (central) {
                            double *drow = a.out.dbg_obs + (size_t)k * (2 * N * M + N);
                            for (int b = 0; b < M; b++) { drow[i * M + b] = 0.0; drow[N * M + i * M + b] = 0.0; }
                            drow[2 * N * M + i] = 0.0;
                        } else {
                            double *drow = a.out.dbg_obs + (size_t)u * OW;
                            for (int b = 0; b <= 4 * M; b++) drow[b] = 0.0;
                        }
                    }
                    if (a.out.dbg_snr)
                        for (int b = 0; b < M; b++) a.out.dbg_snr[u * M + b] = 0.0;
                }
            }
            DCB_TRACE_PT(1, 4);
// [region:O.tile_out]
            // ---- obs tile -> global observation buffer (contiguous span of this CTA): generic-proxy writes of the
            // tile -> visible to the async proxy; then an elected thread issues the bulk copy (TMA, UBLKCP); its read
            // completion is awaited before the tile is rewritten next step
            if (obs_var) {
                // (data-rate observation classes: the rows went straight to the observation buffer)
            } else if (dst && warp_store) {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncwarp();
                if (wrows > 0) {
                    const unsigned wbytes = (unsigned)(wrows * OW * 4);
                    const float *sp = tile + (size_t)w0 * OW;
                    float *gp = dst + (size_t)w0 * OW;
                    const unsigned head = (16u - mis) & 15u;                   // bytes up to the first 16-byte boundary
                    const unsigned bulk = wbytes > head ? (wbytes - head) & ~15u : 0u;
                    if (lane == 0 && bulk) {
                        const unsigned src = (unsigned)__cvta_generic_to_shared(sp) + head;
                        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                                     :: "l"(reinterpret_cast<char *>(gp) + head), "r"(src), "r"(bulk) : "memory");
                        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                    }
                    const int nf = (int)(wbytes >> 2);
                    if (bulk) {
                        const int hf = (int)(head >> 2), tail0 = (int)((head + bulk) >> 2);   // <= 3 floats on either side
                        if (lane >= 1 && lane <= 3 && lane - 1 < hf) gp[lane - 1] = sp[lane - 1];
                        if (lane >= 4 && lane <= 6 && tail0 + lane - 4 < nf) gp[tail0 + lane - 4] = sp[tail0 + lane - 4];
                    } else {
                        for (int j = lane; j < nf; j += 32) gp[j] = sp[j];
                    }
                }
            } else if (dst) {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                bar_sync(BAR_OBS, G);
                const unsigned head = (16u - mis) & 15u;
                const unsigned bulk = tile_bytes > head ? (tile_bytes - head) & ~15u : 0u;
                if (t == 0 && bulk) {
                    const unsigned src = (unsigned)__cvta_generic_to_shared(tile) + head;
                    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                                 :: "l"(reinterpret_cast<char *>(dst) + head), "r"(src), "r"(bulk) : "memory");
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                }
                const int nf = (int)(tile_bytes >> 2);
                if (bulk) {
                    const int hf = (int)(head >> 2), tail0 = (int)((head + bulk) >> 2);
                    if (t >= 1 && t <= 3 && t - 1 < hf) dst[t - 1] = tile[t - 1];
                    if (t >= 4 && t <= 6 && tail0 + t - 4 < nf) dst[tail0 + t - 4] = tile[tail0 + t - 4];
                } else {
                    for (int j = t; j < nf; j += G) dst[j] = tile[j];
                }
            }
            // hand the parity's buffers back to the physics warps
            bar_arrive(BAR_EMPTY + par, 2 * G);
            DCB_TRACE_PT(1, 5);
        }
        if (warp_store ? lane == 0 : t == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    }
}

// [region:wrappers]
// One translation unit per CTA-size class (dcb_step_k<threads>.cu defines DCB_STEP_CLASS and DCB_STEP_REGS and includes
// this file; the classes compile in parallel).  Register budget per class so that one CTA of that size (three of the
// smallest) is always resident: 65536 registers / (warps rounded up to a multiple of 4 x 32), in the allocation granule
// of 8 registers per thread (88 registers x 704 threads does not launch: warps are allocated in fours).
template <bool M32, bool PAD, bool CENTRAL>
__global__ void __maxnreg__(DCB_STEP_REGS) DCB_STEP_KERNEL_NAME(const __grid_constant__ StepArgs a) {
    dcb_step_body<DCB_STEP_CLASS, M32, PAD, CENTRAL>(a);
}

}  // namespace

#ifdef DCB_TRACE_ON
extern "C" int dcb_trace_read(long long *out) {
    return (int)cudaMemcpyFromSymbol(out, dcb_trace_buf, sizeof(dcb_trace_buf));
}
extern "C" int dcb_trace_read_cta(long long *out) {
    return (int)cudaMemcpyFromSymbol(out, dcb_trace_cta, sizeof(dcb_trace_cta));
}
#endif

#define DCB_DISPATCH3(M32V, PADV, central, EXPR)                                                    \
    do {                                                                                            \
        if (central) { auto kern = DCB_STEP_KERNEL_NAME<M32V, PADV, true>; EXPR; }                  \
        else { auto kern = DCB_STEP_KERNEL_NAME<M32V, PADV, false>; EXPR; }                         \
    } while (0)
#define DCB_DISPATCH(m32, pad, central, EXPR)                                                       \
    do {                                                                                            \
        if (m32) {                                                                                  \
            if (pad) DCB_DISPATCH3(true, true, central, EXPR);                                      \
            else DCB_DISPATCH3(true, false, central, EXPR);                                         \
        } else {                                                                                    \
            if (pad) DCB_DISPATCH3(false, true, central, EXPR);                                     \
            else DCB_DISPATCH3(false, false, central, EXPR);                                        \
        }                                                                                           \
    } while (0)

cudaError_t DCB_STEP_CLASS_FN(set_smem)(int n_bs, size_t smem) {
    cudaError_t e = cudaSuccess;
    for (int v = 0; v < 4 && e == cudaSuccess; v++)
        DCB_DISPATCH(n_bs <= 32, v & 1, v >> 1,
                     e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    return e;
}

int DCB_STEP_CLASS_FN(regs)(int n_bs) {
    cudaFuncAttributes at;
    cudaError_t e = cudaSuccess;
    DCB_DISPATCH(n_bs <= 32, false, false, e = cudaFuncGetAttributes(&at, kern));
    return e == cudaSuccess ? at.numRegs : 128;
}

cudaError_t DCB_STEP_CLASS_FN(launch)(const StepArgs &a, int threads, int grid, size_t smem, cudaStream_t s) {
    // the general instance (PAD) also carries the per-handle variants; the fixed-population RelNorm case -- the measured
    // path -- compiles without them
    DCB_DISPATCH(a.p.M <= 32, dcb_step_needs_general(a.p, a.flags), a.p.kind == DCB_KIND_CENTRAL,
                 (kern<<<grid, threads, smem, s>>>(a)));
    return cudaGetLastError();
}
