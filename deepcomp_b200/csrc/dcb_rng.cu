// Per-UE random draws of the reference, generated on the device.
//
// The reference gives every UE two `random.Random` generators seeded with the SAME value seed + 100*i
// (deepcomp/env/single_ue/base.py:132-143, deepcomp/env/entities/user.py:94-96):
//   User.rng       -> initial position: randint(0, W), randint(0, H)            (user.py:98-109)
//   movement.rng   -> per movement.reset(): [velocity randint] , waypoint x, y   (util/movement.py:110-130)
// Draw TIMES are data dependent (a redraw happens when a pause ends, movement.py:172-177) but the draw SEQUENCE
// per UE is not, so the whole sequence is materialised once per reset as a table the step kernel consumes by
// index.  CPython's generator is MT19937 seeded through init_by_array, randint is rejection sampling on the top
// bits of one 32-bit output (Lib/random.py _randbelow_with_getrandbits) -- restated here.
//
// One thread per UE; the 624-word MT state lives in local memory (2.5 KB per thread, L1/L2 backed) -- this
// kernel runs once per episode, it is not on the per-step path.
#include "dcb_internal.h"

namespace {

struct MT {
    uint32_t mt[624];
    int idx;
    uint32_t drawn;     // 32-bit outputs produced since mt_seed
};

__device__ void mt_seed(MT &g, long long seed) {
    // random.seed(int): key = little-endian 32-bit words of abs(seed), at least one
    unsigned long long a = seed < 0 ? (unsigned long long)(-(seed + 1)) + 1ull : (unsigned long long)seed;
    uint32_t key[2] = {(uint32_t)(a & 0xffffffffu), (uint32_t)(a >> 32)};
    const int len = key[1] ? 2 : 1;
    g.mt[0] = 19650218u;
    for (int i = 1; i < 624; i++) g.mt[i] = 1812433253u * (g.mt[i - 1] ^ (g.mt[i - 1] >> 30)) + (uint32_t)i;
    int i = 1, j = 0;
    for (int k = 624; k; k--) {
        g.mt[i] = (g.mt[i] ^ ((g.mt[i - 1] ^ (g.mt[i - 1] >> 30)) * 1664525u)) + key[j] + (uint32_t)j;
        i++; j++;
        if (i >= 624) { g.mt[0] = g.mt[623]; i = 1; }
        if (j >= len) j = 0;
    }
    for (int k = 623; k; k--) {
        g.mt[i] = (g.mt[i] ^ ((g.mt[i - 1] ^ (g.mt[i - 1] >> 30)) * 1566083941u)) - (uint32_t)i;
        i++;
        if (i >= 624) { g.mt[0] = g.mt[623]; i = 1; }
    }
    g.mt[0] = 0x80000000u;
    g.idx = 624;
    g.drawn = 0u;
}

__device__ uint32_t mt_next(MT &g) {
    if (g.idx >= 624) {
        uint32_t *mt = g.mt;
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
        g.idx = 0;
    }
    uint32_t y = g.mt[g.idx++];
    g.drawn++;
    y ^= (y >> 11);
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= (y >> 18);
    return y;
}

__device__ int mt_randint(MT &g, int a, int b) {
    const uint32_t n = (uint32_t)(b - a + 1);
    const int k = 32 - __clz(n);
    uint32_t r = mt_next(g) >> (32 - k);
    while (r >= n) r = mt_next(g) >> (32 - k);
    return a + (int)r;
}

__global__ void __launch_bounds__(128) dcb_generate_kernel(GenArgs a) {
    const long long n_env = a.env_ids ? a.n_ids : a.K;
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_env * a.N) return;
    const int slot = (int)(t / a.N);
    const int i = (int)(t % a.N);
    const int k = a.env_ids ? a.env_ids[slot] : slot;
    const long long u = (long long)k * a.N + i;
    const long long seed = a.ue_seed ? a.ue_seed[u] : a.seeds[k] + 100ll * (i + 1);

    MT g;
    // ---- User.rng: initial position (user.py:98-109); earlier episodes' draws are skipped (rand_episodes)
    const double ix = a.init_xy[2 * i], iy = a.init_xy[2 * i + 1];
    const bool rx = isnan(ix), ry = isnan(iy);
    double px = ix, py = iy;
    if (rx || ry) {
        mt_seed(g, seed);
        const uint32_t skip = a.ue_pos_skip ? a.ue_pos_skip[u] : (a.pos_skip ? a.pos_skip[k] : 0u);
        for (uint32_t e = 0; e <= skip; e++) {
            if (rx) px = (double)mt_randint(g, 0, a.W);
            if (ry) py = (double)mt_randint(g, 0, a.H);
        }
    }
    a.init_pos[u] = make_double2(px, py);

    // ---- movement.rng: successive movement.reset() draws (movement.py:110-130)
    mt_seed(g, seed);
    const double vs = a.vel_spec[i];
    const uint32_t skip = a.mv_skip ? a.mv_skip[u] : 0u;
    uint32_t *row = a.table + u * a.D;
    const int ukx = a.uni_kind ? a.uni_kind[2 * i] : 0, uky = a.uni_kind ? a.uni_kind[2 * i + 1] : 0;
    for (uint32_t e = 0; e < skip + (uint32_t)a.D; e++) {
        if (ukx) {
            // UniformMovement.reset (movement.py:47-64): move_x, then move_y; 'slow' = randint(1, 5), 'fast' = randint(10, 20)
            const int mx = ukx == 2 ? mt_randint(g, 1, 5) : (ukx == 3 ? mt_randint(g, 10, 20) : 0);
            const int my = uky == 2 ? mt_randint(g, 1, 5) : (uky == 3 ? mt_randint(g, 10, 20) : 0);
            if (e >= skip) row[e - skip] = (uint32_t)mx | ((uint32_t)my << 14);
            continue;
        }
        int v = 0;
        if (vs == DCB_VELOCITY_SLOW) v = mt_randint(g, 1, 3);
        else if (vs == DCB_VELOCITY_FAST) v = mt_randint(g, 5, 10);
        const int wx = mt_randint(g, a.border_buffer, a.W - a.border_buffer);
        const int wy = mt_randint(g, a.border_buffer, a.H - a.border_buffer);
        if (e >= skip) row[e - skip] = (uint32_t)wx | ((uint32_t)wy << 14) | ((uint32_t)v << 28);
    }
}

// MobileEnv.reset (base.py:169-189) -> User.reset (user.py:111-116) + Basestation.reset (station.py:106-108)
__global__ void __launch_bounds__(256) dcb_reset_kernel(ResetArgs a) {
    const long long n_env = a.env_ids ? a.n_ids : a.K;
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_env * a.N) return;
    const int slot = (int)(t / a.N);
    const int i = (int)(t % a.N);
    const int k = a.env_ids ? a.env_ids[slot] : slot;
    const long long u = (long long)k * a.N + i;
    a.pos[u] = a.init_pos[u];
    const uint32_t e = a.table[u * a.D];
    const uint32_t wx = e & 0x3fffu, wy = (e >> 14) & 0x3fffu, v = e >> 28;
    a.mv[u] = make_uint2(wx | (wy << 16), v | (1u << 16));   // pausing = 0, curr_pause = 0, tidx = 1
    a.mask[u] = 0ull;
    a.ewma[u] = 0.0;
    if (i == 0) {
        a.time[k] = 0;
        if (a.pos_skip) a.pos_skip[k] += 1u;
    }
}

// rand_episodes (base.py:171-173: no re-seed on reset): before regenerating, remember how many movement.reset()
// draws each UE has consumed so far (table base + tidx) so that the next generate call continues the stream.
// pos_skip[k] counts the resets env k has completed (= reset_pos() draws consumed); it is bumped by the reset
// kernel AFTER generation, so the reset that follows dcb_create replays the first draws.
__global__ void __launch_bounds__(256) dcb_advance_skip_kernel(int K, int N, const int32_t *env_ids, int n_ids,
                                                               const uint2 *mv, uint32_t *mv_skip,
                                                               const uint32_t *pos_skip) {
    const long long n_env = env_ids ? n_ids : K;
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_env * N) return;
    const int slot = (int)(t / N);
    const int i = (int)(t % N);
    const int k = env_ids ? env_ids[slot] : slot;
    const long long u = (long long)k * N + i;
    if (pos_skip[k] >= 1u) mv_skip[u] += mv[u].y >> 16;
}

// ---------------------------------------------------------------------------------------------- variable population
// add_new_ue / remove_ue (single_ue/base.py:592-617) for every env of a lockstep batch.  The two generators involved --
// Map.rng (entities/map.py:30,49-65: rand_border_point) and the global `random` module (base.py:134,612) -- are both
// seeded with the env seed at reset; their position is kept as "outputs consumed so far" and the stream is re-derived
// from the seed for each event (events are rare: a handful per episode).
__global__ void __launch_bounds__(64) dcb_population_kernel(PopArgs a) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= a.K) return;
    const long long base = (long long)k * a.N;
    int na = a.NA;
    MT g;
    // ---- remove_ue (base.py:610-617): idx = random.randint(0, num_ue - 1); ue_list.pop(idx): later UEs move up one
    // list position -- and so does everything indexed by list position, including this step's actions (they were
    // assigned by position BEFORE the removal, base.py:426-427 / central.py:59-63)
    for (int r = 0; r < a.n_rem; r++) {
        mt_seed(g, a.seeds[k]);
        for (uint32_t s = a.glob_draws[k]; s; s--) mt_next(g);
        const int idx = mt_randint(g, 0, na - 1);
        a.glob_draws[k] = g.drawn;
        // an original UE that leaves keeps its generators (it is not in the list MobileEnv.seed walks at the next
        // reset, base.py:138-143): remember how far its movement stream got
        if (!(a.uid[base + idx] & DCB_UID_ARRIVED)) a.ue_mv_used[base + a.uid[base + idx] - 1] += a.mv[base + idx].y >> 16;
        for (int j = idx; j < na - 1; j++) {
            const long long d = base + j, s = d + 1;
            a.pos[d] = a.pos[s]; a.mv[d] = a.mv[s]; a.mask[d] = a.mask[s]; a.ewma[d] = a.ewma[s]; a.uid[d] = a.uid[s];
            a.vel_u[d] = a.vel_u[s];
            for (int e = 0; e < a.D; e++) a.table[d * a.D + e] = a.table[s * a.D + e];
            if (a.actions) a.actions[d] = a.actions[s];
        }
        na--;
        a.mask[base + na] = 0ull;
        if (a.actions) a.actions[base + na] = 0;
    }
    // ---- add_new_ue (base.py:592-608): id = last id + 1, position = map.rand_border_point(), 'slow' RandomWaypoint,
    // both of the UE's generators seeded with env_seed + 100 * id, then User.reset() (fixed position -> no draw from
    // User.rng; movement.reset() -> velocity, waypoint).  The UE is not in the action dict of this step.
    for (int r = 0; r < a.n_add; r++) {
        mt_seed(g, a.seeds[k]);
        for (uint32_t s = a.map_draws[k]; s; s--) mt_next(g);
        const int x = mt_randint(g, 0, a.W);                    // map.py:54-55 (min_x = min_y = 0)
        const int y = mt_randint(g, 0, a.H);
        const int side = mt_randint(g, 0, 3);                   // rng.choice(['left', 'right', 'top', 'bottom'])
        a.map_draws[k] = g.drawn;
        double px, py;
        if (side == 0) { px = 0.0; py = (double)y; }
        else if (side == 1) { px = (double)(a.W - 1); py = (double)y; }
        else if (side == 2) { px = (double)x; py = (double)(a.H - 1); }
        else { px = (double)x; py = 0.0; }
        const long long d = base + na;
        const int new_id = (a.uid[d - 1] & ~DCB_UID_ARRIVED) + 1;     // may repeat the id of an original UE that left
        mt_seed(g, a.seeds[k] + 100ll * new_id);
        uint32_t *row = a.table + d * a.D;
        for (int e = 0; e < a.D; e++) {
            const int v = mt_randint(g, 1, 3);                  // add_new_ue(velocity='slow'), movement.py:112-113
            const int wx = mt_randint(g, a.border_buffer, a.W - a.border_buffer);
            const int wy = mt_randint(g, a.border_buffer, a.H - a.border_buffer);
            row[e] = (uint32_t)wx | ((uint32_t)wy << 14) | ((uint32_t)v << 28);
        }
        const uint32_t e0 = row[0];
        a.pos[d] = make_double2(px, py);
        a.mv[d] = make_uint2((e0 & 0x3fffu) | (((e0 >> 14) & 0x3fffu) << 16), (e0 >> 28) | (1u << 16));
        a.mask[d] = 0ull;
        a.ewma[d] = 0.0;
        a.uid[d] = new_id | DCB_UID_ARRIVED;
        a.vel_u[d] = DCB_VELOCITY_SLOW;
        if (a.actions) a.actions[d] = 0;
        na++;
    }
}

// MobileEnv.seed at reset (base.py:132-143, 171-173) walks the CURRENT list: the UE at list position p gets seed + 100 (p + 1)
// for both of its generators.  Original UEs that left the list are not touched and continue their streams.
__global__ void dcb_pop_reseed_kernel(ReseedArgs a) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= a.K) return;
    const long long base = (long long)k * a.N;
    for (int p = 0; p < a.NA; p++) {
        const int id = a.uid[base + p];
        if (!(id & DCB_UID_ARRIVED)) {       // an original UE (ids of arrivals can repeat those of originals that left)
            a.ue_seed[base + id - 1] = a.seeds[k] + 100ll * (p + 1);
            a.ue_pos_used[base + id - 1] = 0u;
            a.ue_mv_used[base + id - 1] = 0u;
        }
    }
}

__global__ void dcb_pop_seed_init_kernel(long long *ue_seed, uint32_t *pos_used, uint32_t *mv_used, const long long *seeds,
                                         long long n, int N) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    ue_seed[t] = seeds[t / N] + 100ll * ((t % N) + 1);
    pos_used[t] = 0u;
    mv_used[t] = 0u;
}

__global__ void dcb_broadcast_vel_kernel(double *vel_u, const double *vel_spec, long long n, int N) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n) vel_u[t] = vel_spec[t % N];
}

__global__ void dcb_add_u32_kernel(uint32_t *a, long long n, uint32_t v) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n) a[t] += v;
}

__global__ void dcb_iota_uid_kernel(int32_t *uid, long long n, int N) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n) uid[t] = (int32_t)(t % N) + 1;                   // ids "1".."N" (env_setup.py:148-160)
}

}  // namespace

cudaError_t dcb_launch_population(const PopArgs &a, cudaStream_t s) {
    dcb_population_kernel<<<(a.K + 63) / 64, 64, 0, s>>>(a);
    return cudaGetLastError();
}

cudaError_t dcb_launch_pop_reseed(const ReseedArgs &a, cudaStream_t s) {
    dcb_pop_reseed_kernel<<<(a.K + 63) / 64, 64, 0, s>>>(a);
    return cudaGetLastError();
}

cudaError_t dcb_launch_pop_seed_init(long long *ue_seed, uint32_t *pos_used, uint32_t *mv_used, const long long *seeds,
                                     int K, int N, cudaStream_t s) {
    const long long n = (long long)K * N;
    dcb_pop_seed_init_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(ue_seed, pos_used, mv_used, seeds, n, N);
    return cudaGetLastError();
}

cudaError_t dcb_launch_broadcast_vel(double *vel_u, const double *vel_spec, int K, int N, cudaStream_t s) {
    const long long n = (long long)K * N;
    dcb_broadcast_vel_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(vel_u, vel_spec, n, N);
    return cudaGetLastError();
}

cudaError_t dcb_launch_add_u32(uint32_t *a, long long n, uint32_t v, cudaStream_t s) {
    dcb_add_u32_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(a, n, v);
    return cudaGetLastError();
}

cudaError_t dcb_launch_iota_uid(int32_t *uid, int K, int N, cudaStream_t s) {
    const long long n = (long long)K * N;
    dcb_iota_uid_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(uid, n, N);
    return cudaGetLastError();
}

cudaError_t dcb_launch_advance_skip(int K, int N, const int32_t *env_ids, int n_ids, const uint2 *mv,
                                    uint32_t *mv_skip, const uint32_t *pos_skip, cudaStream_t s) {
    const long long n = (long long)(env_ids ? n_ids : K) * N;
    if (n == 0) return cudaSuccess;
    dcb_advance_skip_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(K, N, env_ids, n_ids, mv, mv_skip, pos_skip);
    return cudaGetLastError();
}

// Continuous stepping past episode_length (the reference's --cont-train / soft_horizon: `done` is never set, base.py:371-381):
// the table row of UE u moves on to the entries from its cursor onwards.  mode 0: mv_skip[u] += cursor, cursor = 0 (then
// regenerate with mv_skip); mode 1: mv_skip[u] = 0 for the listed envs (a reset goes back to the start of the stream).
__global__ void __launch_bounds__(256) dcb_table_cursor_kernel(int K, int N, const int32_t *env_ids, int n_ids, uint2 *mv,
                                                               uint32_t *mv_skip, int mode) {
    const long long n_env = env_ids ? n_ids : K;
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_env * N) return;
    const int k = env_ids ? env_ids[t / N] : (int)(t / N);
    const long long u = (long long)k * N + (t % N);
    if (mode == 0) {
        mv_skip[u] += mv[u].y >> 16;
        mv[u].y &= 0xffffu;
    } else {
        mv_skip[u] = 0u;
    }
}

cudaError_t dcb_launch_table_cursor(int K, int N, const int32_t *env_ids, int n_ids, uint2 *mv, uint32_t *mv_skip, int mode,
                                    cudaStream_t s) {
    const long long n = (long long)(env_ids ? n_ids : K) * N;
    if (n == 0) return cudaSuccess;
    dcb_table_cursor_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(K, N, env_ids, n_ids, mv, mv_skip, mode);
    return cudaGetLastError();
}

cudaError_t dcb_launch_generate(const GenArgs &a, cudaStream_t s) {
    const long long n = (long long)(a.env_ids ? a.n_ids : a.K) * a.N;
    if (n == 0) return cudaSuccess;
    dcb_generate_kernel<<<(unsigned)((n + 127) / 128), 128, 0, s>>>(a);
    return cudaGetLastError();
}

cudaError_t dcb_launch_reset(const ResetArgs &a, cudaStream_t s) {
    const long long n = (long long)(a.env_ids ? a.n_ids : a.K) * a.N;
    if (n == 0) return cudaSuccess;
    dcb_reset_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(a);
    return cudaGetLastError();
}
