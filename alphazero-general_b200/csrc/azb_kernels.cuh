// azb_kernels.cuh -- the hot kernels of the batched self-play MCTS engine,
// templated on the game rules (G = Connect4 | Brandubh).
//
// Work decomposition: G::LANES consecutive threads of a warp (a "group") own
// one game slot; child k of the node being scanned lives in lane k % LANES, so
// the N/Q/P loads of a sibling block are coalesced and the argmax is a
// shuffle reduction.  There is exactly one simulation in flight per game
// (the reference has no virtual loss, SelfPlayAgent.pyx:108-110), so the
// parallelism is across the B games.
//
//   select_game         MCTS.find_leaf             MCTS.pyx:208-228, :86-104, :76-79
//   expand_backup_game  MCTS.process_results       MCTS.pyx:230-289, :197-206, :291-295
//   play_move_game      SelfPlayAgent.playMoves    SelfPlayAgent.pyx:153-176, MCTS.pyx:185-195, :297-327
//   k_finalize / k_emit SelfPlayAgent.playMoves    SelfPlayAgent.pyx:176-202 (quota, samples, reset)
#pragma once
#include "azb_common.cuh"

namespace azb {

constexpr int CTA_THREADS = 128;

template <class G>
struct GroupSmem {
    float vec[G::A];        // action-indexed scratch (masked priors / counts)
    float vec2[G::A];       // action-indexed scratch (probabilities)
    uint32_t key[G::MAXC];  // Philox sort keys
    short act[G::MAXC];     // valid actions, ascending
    short order[G::MAXC];   // order[k] = index into act[] of the child at position k
};

// ------------------------------------------------------------------------------
// numpy float32 add.reduce (pairwise_sum) over an action-indexed vector in
// shared memory; every lane of the group returns the same value.
// ------------------------------------------------------------------------------
__device__ __forceinline__ float np_leaf_sum(const float *a, int n)
{
    // n <= 128: one lane's serial restatement (used for small n)
    if (n < 8) {
        float r = 0.0f;
        for (int i = 0; i < n; i++) r = f_add(r, a[i]);
        return r;
    }
    float r[8];
#pragma unroll
    for (int j = 0; j < 8; j++) r[j] = a[j];
    int i;
    for (i = 8; i < n - (n % 8); i += 8) {
#pragma unroll
        for (int j = 0; j < 8; j++) r[j] = f_add(r[j], a[i + j]);
    }
    float res = f_add(f_add(f_add(r[0], r[1]), f_add(r[2], r[3])), f_add(f_add(r[4], r[5]), f_add(r[6], r[7])));
    for (; i < n; i++) res = f_add(res, a[i]);
    return res;
}

__device__ inline float np_sum_serial(const float *a, int n)
{
    if (n <= 128) return np_leaf_sum(a, n);
    int n2 = n / 2;
    n2 -= n2 % 8;
    return f_add(np_sum_serial(a, n2), np_sum_serial(a + n2, n - n2));
}

template <class G>
__device__ __forceinline__ float np_sum_group(const float *vec, int lane, unsigned gmask)
{
    __syncwarp(gmask);
    float r;
    if (G::A < 8) {
        r = np_leaf_sum(vec, G::A);          // every lane, broadcast reads
    } else {
        r = 0.0f;
        if (lane == 0) r = np_sum_serial(vec, G::A);
        r = __shfl_sync(gmask, r, 0, G::LANES);
    }
    return r;
}

// ------------------------------------------------------------------------------
// RNG draws of a slot (lane 0 draws, the group gets the value)
// ------------------------------------------------------------------------------
template <class G>
__device__ __forceinline__ void rng_two_words(const DevView &d, int g, int lane, unsigned gmask, uint32_t &a, uint32_t &b)
{
    a = b = 0;
    if (lane == 0) {
        if (d.rng_mode == 0) {
            uint32_t *st = d.mt + (size_t)g * 625;
            a = mt_next(st);
            b = mt_next(st);
        } else {
            unsigned long long c = d.ctr[g];
            a = philox_word(d.seed, (unsigned long long)(d.gid_base + g), c);
            b = philox_word(d.seed, (unsigned long long)(d.gid_base + g), c + 1);
            d.ctr[g] = c + 2;
        }
    }
    a = __shfl_sync(gmask, a, 0, G::LANES);
    b = __shfl_sync(gmask, b, 0, G::LANES);
}

// Child order of a freshly expanded node with C children: fills sm.order
// (order[k] = index into sm.act of the child at list position k).
//   MT mode     : numpy legacy list shuffle (reversed Fisher-Yates, masked
//                 rejection), drawn serially by lane 0.
//   Philox mode : child j takes key word ctr+j; children are ordered by
//                 (key, j) -- computed in parallel by rank counting.
// With store == false only the RNG consumption is replayed (terminal leaves:
// the reference shuffles their never-used children too, MCTS.pyx:223-226).
template <class G>
__device__ __forceinline__ void child_order(const DevView &d, int g, int C, bool store, int lane, unsigned gmask, GroupSmem<G> &sm)
{
    constexpr int L = G::LANES;
    if (d.rng_mode == 0) {
        if (lane == 0) {
            uint32_t *st = d.mt + (size_t)g * 625;
            if (store) for (int k = 0; k < C; k++) sm.order[k] = (short)k;
            for (int i = C - 1; i >= 1; i--) {
                uint32_t j = mt_interval(st, (uint32_t)i);
                if (store) { short t = sm.order[i]; sm.order[i] = sm.order[j]; sm.order[j] = t; }
            }
        }
    } else {
        unsigned long long c0 = d.ctr[g];
        __syncwarp(gmask);
        if (lane == 0) d.ctr[g] = c0 + (unsigned long long)C;
        if (store) {
            for (int j = lane; j < C; j += L)
                sm.key[j] = philox_word(d.seed, (unsigned long long)(d.gid_base + g), c0 + (unsigned long long)j);
            __syncwarp(gmask);
            for (int j = lane; j < C; j += L) {
                uint32_t kj = sm.key[j];
                int rank = 0;
                for (int i = 0; i < C; i++) {
                    uint32_t ki = sm.key[i];
                    rank += (ki < kj) || (ki == kj && i < j);
                }
                sm.order[rank] = (short)j;
            }
        }
    }
    __syncwarp(gmask);
}

// ------------------------------------------------------------------------------
// seen_policy = float(sum(p of visited children)) as CPython's sum() computes it
// (double accumulation in child order, Neumaier-compensated; MCTS.pyx:91).
// Fast path: when every visited prior is 0 or >= 2^-28 the double partial sums
// are exact, so an order-free integer reduction gives the identical result.
// ------------------------------------------------------------------------------
template <class G, int IT>
__device__ __forceinline__ float seen_policy(const int (&kn)[IT], const float (&kp)[IT], int C, int lane, unsigned gmask)
{
    constexpr int L = G::LANES;
    bool ok = true;
    unsigned long long acc = 0ULL;
#pragma unroll
    for (int i = 0; i < IT; i++) {
        int k = lane + i * L;
        if (k < C && kn[i] > 0) {
            float pv = kp[i];
            ok = ok && ((pv == 0.0f) || (pv >= 3.7252902984619140625e-09f && pv < 2.0f));
            acc += (unsigned long long)(pv * 2251799813685248.0f);   // p * 2^51, exact
        }
    }
    if (__all_sync(gmask, ok)) {
#pragma unroll
        for (int off = L / 2; off >= 1; off >>= 1) acc += __shfl_xor_sync(gmask, acc, off, L);
        return (float)((double)acc * 4.44089209850062616169452667236328125e-16);   // * 2^-51
    }
    // general path: serial compensated sum in child order
    double f = 0.0, comp = 0.0;
    for (int k = 0; k < C; k++) {
        int i = k / L, src = k - i * L;
        int nn = 0; float pv = 0.0f;
#pragma unroll
        for (int ii = 0; ii < IT; ii++) if (ii == i) { nn = kn[ii]; pv = kp[ii]; }
        nn = __shfl_sync(gmask, nn, src, L);
        pv = __shfl_sync(gmask, pv, src, L);
        if (nn > 0) {
            double x = (double)pv, t = __dadd_rn(f, x);
            if (fabs(f) >= fabs(x)) comp = __dadd_rn(comp, __dadd_rn(__dsub_rn(f, t), x));
            else comp = __dadd_rn(comp, __dadd_rn(__dsub_rn(x, t), f));
            f = t;
        }
    }
    if (comp != 0.0 && isfinite(comp)) f = __dadd_rn(f, comp);
    return (float)f;
}

// ------------------------------------------------------------------------------
// MCTS.find_leaf
// ------------------------------------------------------------------------------
template <class G, bool WRITE_OBS>
__device__ __forceinline__ void select_game(const DevView &d, int g, int lane, unsigned gmask, GroupSmem<G> &sm)
{
    constexpr int L = G::LANES;
    constexpr int IT = (G::MAXC + L - 1) / L;
    if (d.finished[g] != 0) return;                 // dead slot (finished beyond the quota)
    GState st = d.state[g];
    const size_t nb = (size_t)g * (size_t)d.npg;
    int *path = d.path + (size_t)g * G::MAXD;
    int cur = d.root[g];
    int cn = d.n[nb + cur];
    uint32_t cmeta = d.meta[nb + cur];
    int cch = d.child0[nb + cur];
    float cv = d.v[nb + cur];
    int depth = 0, sumc = 0;

    while (cn > 0 && meta_e(cmeta) == 0) {
        const int C = meta_nc(cmeta);
        if (C == 0 || depth >= G::MAXD - 1) break;  // only after a pool-exhaustion error
        if (lane == 0) path[depth] = cur;
        const size_t cb = nb + (size_t)cch;
        int kn[IT], kc[IT]; float kq[IT], kp[IT], kv[IT]; uint32_t km[IT];
#pragma unroll
        for (int i = 0; i < IT; i++) {
            int k = lane + i * L;
            bool in = k < C;
            kn[i] = in ? d.n[cb + k] : 0;
            kq[i] = in ? d.q[cb + k] : 0.0f;
            kp[i] = in ? d.p[cb + k] : 0.0f;
            kv[i] = in ? d.v[cb + k] : 0.0f;
            kc[i] = in ? d.child0[cb + k] : -1;
            km[i] = in ? d.meta[cb + k] : 0u;
        }
        // Node.best_child: fpu value in double, uct in float32, first strict maximum
        float seen = seen_policy<G, IT>(kn, kp, C, lane, gmask);
        float fpu = (float)__dsub_rn((double)cv, __dmul_rn((double)d.fpu_reduction, __dsqrt_rn((double)seen)));
        float sqrt_n = (float)__dsqrt_rn((double)cn);
        float bu = -CUDART_INF_F; int bk = 0x7fffffff;
#pragma unroll
        for (int i = 0; i < IT; i++) {
            int k = lane + i * L;
            if (k < C) {
                float t = f_div(f_mul(f_mul(d.cpuct, kp[i]), sqrt_n), (float)(1 + kn[i]));
                float u = f_add(kn[i] == 0 ? fpu : kq[i], t);
                if (u > bu) { bu = u; bk = k; }
            }
        }
#pragma unroll
        for (int off = L / 2; off >= 1; off >>= 1) {
            float ou = __shfl_xor_sync(gmask, bu, off, L);
            int ok = __shfl_xor_sync(gmask, bk, off, L);
            if (ou > bu || (ou == bu && ok < bk)) { bu = ou; bk = ok; }
        }
        if (bk == 0x7fffffff) {                      // every uct was NaN
            if (lane == 0) atomicOr(d.err, ERRB_FP);
            bk = 0;
        }
        const int bi = bk / L, src = bk - bi * L;
        int sn = 0, sc = -1; float sv = 0.0f; uint32_t smeta = 0u;
#pragma unroll
        for (int i = 0; i < IT; i++) if (i == bi) { sn = kn[i]; sc = kc[i]; sv = kv[i]; smeta = km[i]; }
        cn = __shfl_sync(gmask, sn, src, L);
        cch = __shfl_sync(gmask, sc, src, L);
        cv = __shfl_sync(gmask, sv, src, L);
        cmeta = __shfl_sync(gmask, smeta, src, L);
        cur = (int)(cb - nb) + bk;
        G::play(st, meta_action(cmeta));
        sumc += C;
        depth++;
    }

    if (cn == 0) {
        // first visit: record player and win state, materialise the children
        const int player = G::player(st);
        const int e = G::win_code(st);
        const int C = G::list_valid(st, sm.act, lane, gmask);
        int nc = 0, base = -1;
        if (e != 0) {
            child_order<G>(d, g, C, false, lane, gmask, sm);
        } else {
            base = d.alloc[g];
            __syncwarp(gmask);
            if (base + C > d.npg) {
                if (lane == 0) atomicOr(d.err, ERRB_POOL);
                child_order<G>(d, g, C, false, lane, gmask, sm);
                base = -1;
            } else {
                child_order<G>(d, g, C, true, lane, gmask, sm);
                nc = C;
                for (int k = lane; k < C; k += L) {
                    size_t idx = nb + (size_t)(base + k);
                    d.n[idx] = 0; d.q[idx] = 0.0f; d.p[idx] = 0.0f; d.v[idx] = 0.0f;
                    d.child0[idx] = -1;
                    d.meta[idx] = meta_pack((uint32_t)sm.act[sm.order[k]], 0u, 0u, 0u);
                }
                if (lane == 0) {
                    d.alloc[g] = base + C;
                    SlotStats &ss = d.stats[g];
                    ss.nodes_created += (unsigned long long)C;
                    if (base + C > ss.peak_nodes) ss.peak_nodes = base + C;
                }
            }
        }
        cmeta = meta_pack((uint32_t)meta_action(cmeta), (uint32_t)nc, (uint32_t)e, (uint32_t)player);
        if (lane == 0) { d.meta[nb + cur] = cmeta; d.child0[nb + cur] = base; }
    }
    if (WRITE_OBS) {
        float *o = d.obs + (size_t)g * G::OBS;
        for (int i = lane; i < G::OBS; i += L) o[i] = G::obs_value(st, i);
    }
    if (lane == 0) {
        d.leaf[g] = cur;
        d.path_len[g] = depth;
        SlotStats &ss = d.stats[g];
        ss.sum_depth += (unsigned long long)depth;
        ss.sum_children += (unsigned long long)sumc;
        if (meta_e(cmeta) != 0) ss.terminal_leaves += 1ULL;
    }
    __syncwarp(gmask);
}

// ------------------------------------------------------------------------------
// Device Dirichlet(alpha, ..., alpha) for the root noise when no table is fed
// (MCTS._add_root_noise, MCTS.pyx:197-206: alpha = 10.83 / C).  Child k draws a
// Gamma(alpha) variate (Marsaglia-Tsang, boosted for alpha < 1) from a private
// 64-word window of the slot's Philox stream; the variates are normalised by
// their sum.  Statistically equivalent to numpy's sampler, not bit-equal to it:
// parity runs feed the vectors instead (azb_set_root_noise).
// ------------------------------------------------------------------------------
__device__ inline float gamma_variate(unsigned long long seed, unsigned long long gid, unsigned long long w0, double alpha)
{
    unsigned long long w = w0;
    auto uni = [&]() {   // (0,1)
        uint32_t a = philox_word(seed, gid, w), b = philox_word(seed, gid, w + 1);
        w += 2;
        return (((double)(a >> 5) * 67108864.0 + (double)(b >> 6)) + 0.5) / 9007199254740992.0;
    };
    double boost = 1.0, a = alpha;
    if (a < 1.0) { boost = pow(uni(), 1.0 / a); a += 1.0; }
    const double dd = a - 1.0 / 3.0, cc = 1.0 / sqrt(9.0 * dd);
    double out = dd;
    for (int it = 0; it < 12; it++) {
        double u1 = uni(), u2 = uni();
        double x = sqrt(-2.0 * log(u1)) * cospi(2.0 * u2);      // Box-Muller normal
        double vv = 1.0 + cc * x;
        if (vv <= 0.0) continue;
        vv = vv * vv * vv;
        double u = uni();
        out = dd * vv;
        if (log(u) < 0.5 * x * x + dd - dd * vv + dd * log(vv)) break;
    }
    return (float)(out * boost);
}

// ------------------------------------------------------------------------------
// MCTS.process_results
// ------------------------------------------------------------------------------
template <class G>
__device__ __forceinline__ void expand_backup_game(const DevView &d, int g, int lane, unsigned gmask, GroupSmem<G> &sm,
                                                   const float *pol_row, const float *val_row)
{
    constexpr int L = G::LANES;
    if (d.finished[g] != 0) return;
    const size_t nb = (size_t)g * (size_t)d.npg;
    const int leaf = d.leaf[g], depth = d.path_len[g], root = d.root[g];
    const uint32_t lmeta = d.meta[nb + leaf];
    const int e = meta_e(lmeta);
    float val0, val1, val2;
    if (e != 0) {
        // value = np.array(self._curnode.e, dtype=np.float32)
        val0 = e == 1 ? 1.0f : 0.0f; val1 = e == 2 ? 1.0f : 0.0f; val2 = e == 3 ? 1.0f : 0.0f;
    } else {
        val0 = val_row[0]; val1 = val_row[1]; val2 = val_row[2];
        const int C = meta_nc(lmeta);
        const size_t cb = nb + (size_t)d.child0[nb + leaf];
        // pi *= valids (valids rebuilt from the children); pi /= np.sum(pi)
        for (int a = lane; a < G::A; a += L) sm.vec[a] = 0.0f;
        __syncwarp(gmask);
        for (int k = lane; k < C; k += L) {
            int a = meta_action(d.meta[cb + k]);
            sm.vec[a] = pol_row[a];
        }
        float sum = np_sum_group<G>(sm.vec, lane, gmask);
        if (!(sum > 0.0f) && lane == 0 && C > 0) atomicOr(d.err, ERRB_FP);
        for (int a = lane; a < G::A; a += L) sm.vec[a] = f_div(sm.vec[a], sum);
        if (leaf == root) {
            if (d.add_temp) {
                __syncwarp(gmask);
                for (int a = lane; a < G::A; a += L) sm.vec[a] = pow_det(sm.vec[a], d.root_temp_exp);
                float sum2 = np_sum_group<G>(sm.vec, lane, gmask);
                for (int a = lane; a < G::A; a += L) sm.vec[a] = f_div(sm.vec[a], sum2);
            }
            __syncwarp(gmask);
            const float *nz = nullptr;
            bool gen = false;
            if (d.add_noise) {
                int ev = d.noise_event[g];
                if (d.noise != nullptr) {
                    if (ev < d.noise_events && C <= d.noise_stride)
                        nz = d.noise + ((size_t)g * d.noise_events + ev) * d.noise_stride;
                    else if (lane == 0) atomicOr(d.err, ERRB_NOISE);
                } else gen = true;
            }
            const float keep = __fsub_rn(1.0f, d.noise_frac);
            if (gen) {
                // sample Dirichlet(10.83 / C) on the device; sm.vec2[k] = gamma variate of child k
                const unsigned long long c0 = d.ctr[g];
                __syncwarp(gmask);
                if (lane == 0) d.ctr[g] = c0 + 64ULL * (unsigned long long)C;
                float part = 0.0f;
                for (int k = lane; k < C; k += L) {
                    float gv = gamma_variate(d.seed, (unsigned long long)(d.gid_base + g), c0 + 64ULL * k, 10.83 / (double)C);
                    sm.key[k] = __float_as_uint(gv);
                    part += gv;
                }
#pragma unroll
                for (int off = L / 2; off >= 1; off >>= 1) part += __shfl_xor_sync(gmask, part, off, L);
                __syncwarp(gmask);
                for (int k = lane; k < C; k += L) {
                    float pk = sm.vec[meta_action(d.meta[cb + k])];
                    float nk = __uint_as_float(sm.key[k]) / part;
                    d.p[cb + k] = f_add(f_mul(pk, keep), f_mul(d.noise_frac, nk));
                }
            } else {
                for (int k = lane; k < C; k += L) {
                    float pk = sm.vec[meta_action(d.meta[cb + k])];
                    if (nz != nullptr) pk = f_add(f_mul(pk, keep), f_mul(d.noise_frac, nz[k]));
                    d.p[cb + k] = pk;
                }
            }
            __syncwarp(gmask);
            if (lane == 0) d.noise_event[g] += 1;
        } else {
            __syncwarp(gmask);
            for (int k = lane; k < C; k += L) d.p[cb + k] = sm.vec[meta_action(d.meta[cb + k])];
        }
    }
    // backup along the stored path; level i updates the node entered at step i
    const int *path = d.path + (size_t)g * G::MAXD;
    const float share = f_div(val2, 2.0f);            // value[num_players] / num_players
    for (int i = lane; i < depth; i += L) {
        const int node = (i == depth - 1) ? leaf : path[i + 1];
        const int parent = path[i];
        const int pp = meta_player(d.meta[nb + parent]);
        const float v = f_add(pp == 0 ? val0 : val1, share);
        const size_t idx = nb + (size_t)node;
        const int nn = d.n[idx];
        const float qq = d.q[idx];
        d.q[idx] = f_div(f_add(f_mul(qq, (float)nn), f_mul(v, 1.0f)), (float)(nn + 1));
        if (nn == 0) {
            const int np_ = (i == depth - 1) ? meta_player(lmeta) : meta_player(d.meta[idx]);
            d.v[idx] = f_add(np_ == 0 ? val0 : val1, share);
        }
        d.n[idx] = nn + 1;
    }
    if (lane == 0) {
        d.n[nb + root] += 1;
        d.stats[g].sims += 1ULL;
    }
    __syncwarp(gmask);
}

// ------------------------------------------------------------------------------
// MCTS.probs on an action-indexed count vector in shared memory
// ------------------------------------------------------------------------------
template <class G>
__device__ __forceinline__ void probs_group(const DevView &d, const float *counts, float temp, float *out, int lane, unsigned gmask)
{
    constexpr int L = G::LANES;
    __syncwarp(gmask);
    if (temp == 0.0f) {
        int best = 0;
        for (int a = 1; a < G::A; a++) if (counts[a] > counts[best]) best = a;   // np.argmax: first maximum
        for (int a = lane; a < G::A; a += L) out[a] = a == best ? 1.0f : 0.0f;
        __syncwarp(gmask);
        return;
    }
    float sum = np_sum_group<G>(counts, lane, gmask);
    if (!(sum > 0.0f) && lane == 0) atomicOr(d.err, ERRB_FP);
    const float e = (float)(1.0 / (double)temp);
    for (int a = lane; a < G::A; a += L) out[a] = pow_det(f_div(counts[a], sum), e);
    float sum2 = np_sum_group<G>(out, lane, gmask);
    for (int a = lane; a < G::A; a += L) out[a] = f_div(out[a], sum2);
    __syncwarp(gmask);
}

__device__ __forceinline__ void tree_reset(const DevView &d, int g)
{
    const size_t nb = (size_t)g * (size_t)d.npg;
    d.n[nb] = 0; d.q[nb] = 0.0f; d.p[nb] = 0.0f; d.v[nb] = 0.0f; d.child0[nb] = -1;
    d.meta[nb] = meta_pack(META_ACTION_NONE, 0u, 0u, 0u);
    d.root[g] = 0;
    d.alloc[g] = 1;
    d.path_len[g] = 0;
    d.leaf[g] = 0;
}

// ------------------------------------------------------------------------------
// SelfPlayAgent.playMoves, the per-game part up to the terminal test
// ------------------------------------------------------------------------------
template <class G>
__device__ __forceinline__ void play_move_game(const DevView &d, int g, int fast, int lane, unsigned gmask, GroupSmem<G> &sm)
{
    constexpr int L = G::LANES;
    if (d.finished[g] != 0) return;
    GState st = d.state[g];
    const size_t nb = (size_t)g * (size_t)d.npg;
    const int root = d.root[g];
    const uint32_t rmeta = d.meta[nb + root];
    const int C = meta_nc(rmeta);
    const size_t cb = nb + (size_t)d.child0[nb + root];
    // MCTS.counts
    for (int a = lane; a < G::A; a += L) sm.vec[a] = 0.0f;
    __syncwarp(gmask);
    for (int k = lane; k < C; k += L) sm.vec[meta_action(d.meta[cb + k])] = (float)d.n[cb + k];
    __syncwarp(gmask);
    int t = st.turns < d.temp_len ? st.turns : d.temp_len - 1;
    const float temp = d.temp_table[t];
    probs_group<G>(d, sm.vec, temp, sm.vec2, lane, gmask);
    // np.random.choice(A, p=policy): cdf in double, one 53-bit uniform, searchsorted 'right'
    uint32_t wa, wb;
    rng_two_words<G>(d, g, lane, gmask, wa, wb);
    int action = 0;
    if (lane == 0) {
        const double u = u53(wa, wb);
        double last = 0.0;
        for (int a = 0; a < G::A; a++) last = __dadd_rn(last, (double)sm.vec2[a]);
        double acc = 0.0;
        action = G::A;
        for (int a = 0; a < G::A; a++) {
            acc = __dadd_rn(acc, (double)sm.vec2[a]);
            if (u < __ddiv_rn(acc, last)) { action = a; break; }
        }
    }
    action = __shfl_sync(gmask, action, 0, L);
    if (!fast) {
        // histories[i].append((game.clone(), mcts.probs(game)))  -- temp = 1
        const int hl = d.hist_len[g];
        if (hl < d.hist_cap) {
            float *hp = d.hist_pi + ((size_t)g * d.hist_cap + hl) * G::A;
            probs_group<G>(d, sm.vec, 1.0f, sm.vec2, lane, gmask);
            for (int a = lane; a < G::A; a += L) hp[a] = sm.vec2[a];
            if (lane == 0) { d.hist_state[(size_t)g * d.hist_cap + hl] = st; d.hist_len[g] = hl + 1; }
        } else if (lane == 0) atomicOr(d.err, ERRB_SAMPLES);
    }
    // MCTS.update_root
    int found = -1;
    for (int k = lane; k < C; k += L) if (meta_action(d.meta[cb + k]) == action) found = k;
#pragma unroll
    for (int off = L / 2; off >= 1; off >>= 1) found = max(found, __shfl_xor_sync(gmask, found, off, L));
    if (found < 0) {
        if (lane == 0) { atomicOr(d.err, ERRB_ACTION); d.finished[g] = 2; }
        __syncwarp(gmask);
        return;
    }
    G::play(st, action);
    const int e = G::win_code(st);
    if (lane == 0) {
        d.root[g] = (int)(cb - nb) + found;
        d.state[g] = st;
        d.last_action[g] = action;
        d.stats[g].moves += 1ULL;
        if (d.reset_threshold && st.turns >= d.next_reset[g]) {
            tree_reset(d, g);
            d.next_reset[g] = st.turns + d.reset_threshold;
        }
        if (e != 0) { d.finished[g] = 1; d.fin_code[g] = e; }
    }
    __syncwarp(gmask);
}

// ------------------------------------------------------------------------------
// kernels
// ------------------------------------------------------------------------------
template <class G>
__device__ __forceinline__ bool group_setup(int first, int count, int &g, int &lane, unsigned &gmask, int &gi)
{
    constexpr int L = G::LANES;
    const int tid = blockIdx.x * blockDim.x + threadIdx.x;
    const int grp = tid / L;
    lane = threadIdx.x % L;
    gi = threadIdx.x / L;
    const int sub = (threadIdx.x % 32) / L;
    gmask = (L == 32) ? 0xffffffffu : (((1u << L) - 1u) << (sub * L));
    g = first + grp;
    return grp < count;
}

template <class G>
__global__ void __launch_bounds__(CTA_THREADS) k_select(DevView d, int first, int count)
{
    __shared__ GroupSmem<G> sm[CTA_THREADS / G::LANES];
    int g, lane, gi; unsigned gmask;
    if (!group_setup<G>(first, count, g, lane, gmask, gi)) return;
    select_game<G, true>(d, g, lane, gmask, sm[gi]);
}

template <class G>
__global__ void __launch_bounds__(CTA_THREADS) k_expand_backup(DevView d, int first, int count, const float *policy, const float *value)
{
    __shared__ GroupSmem<G> sm[CTA_THREADS / G::LANES];
    int g, lane, gi; unsigned gmask;
    if (!group_setup<G>(first, count, g, lane, gmask, gi)) return;
    expand_backup_game<G>(d, g, lane, gmask, sm[gi], policy + (size_t)g * G::A, value + (size_t)g * 3);
}

// `sims` simulations per game with constant NN outputs, no NN round trip
template <class G>
__global__ void __launch_bounds__(CTA_THREADS) k_warmup_sims(DevView d, int sims)
{
    __shared__ GroupSmem<G> sm[CTA_THREADS / G::LANES];
    int g, lane, gi; unsigned gmask;
    if (!group_setup<G>(0, d.B, g, lane, gmask, gi)) return;
    for (int s = 0; s < sims; s++) {
        select_game<G, false>(d, g, lane, gmask, sm[gi]);
        expand_backup_game<G>(d, g, lane, gmask, sm[gi], d.warm_policy, d.warm_value);
    }
}

template <class G>
__global__ void __launch_bounds__(CTA_THREADS) k_play_moves(DevView d, int fast)
{
    __shared__ GroupSmem<G> sm[CTA_THREADS / G::LANES];
    int g, lane, gi; unsigned gmask;
    if (!group_setup<G>(0, d.B, g, lane, gmask, gi)) return;
    play_move_game<G>(d, g, fast, lane, gmask, sm[gi]);
}

// Terminal handling in slot order (one CTA): result_queue.put for every
// finished game; the games_played quota decides, in slot order as the
// reference's worker loop does, which of them emit samples and restart.
template <class G>
__global__ void __launch_bounds__(1024) k_finalize(DevView d)
{
    __shared__ int s_scan[1024];
    __shared__ long long s_scan2[1024];
    __shared__ long long s_base[3];     // results, games_played, sample offset
    const int tid = threadIdx.x;
    if (tid == 0) {
        s_base[0] = d.counters->result_count;
        s_base[1] = d.counters->games_played;
        s_base[2] = d.counters->sample_count;
    }
    __syncthreads();
    const int per = d.symmetric ? G::NSYM : 1;
    for (int g0 = 0; g0 < d.B; g0 += 1024) {
        const int g = g0 + tid;
        const int fin = (g < d.B && d.finished[g] == 1) ? 1 : 0;
        // inclusive scan of fin
        s_scan[tid] = fin;
        __syncthreads();
        for (int off = 1; off < 1024; off <<= 1) {
            int v = tid >= off ? s_scan[tid - off] : 0;
            __syncthreads();
            s_scan[tid] += v;
            __syncthreads();
        }
        const int rank = s_scan[tid] - fin;          // finished games before this one in the chunk
        const int total_fin = s_scan[1023];
        const long long rbase = s_base[0], gbase = s_base[1];
        const bool accepted = fin && (gbase + rank < d.quota);
        long long nsamp = accepted ? (long long)d.hist_len[g] * per : 0;
        s_scan2[tid] = nsamp;
        __syncthreads();
        for (int off = 1; off < 1024; off <<= 1) {
            long long v = tid >= off ? s_scan2[tid - off] : 0;
            __syncthreads();
            s_scan2[tid] += v;
            __syncthreads();
        }
        const long long soff = s_base[2] + s_scan2[tid] - nsamp;
        const long long total_samp = s_scan2[1023];
        if (fin) {
            const long long ri = rbase + rank;
            if (ri < d.r_cap) {
                const int code = d.fin_code[g];
                d.r_slot[ri] = g;
                d.r_turns[ri] = d.state[g].turns;
                d.r_win[ri * 3 + 0] = code == 1; d.r_win[ri * 3 + 1] = code == 2; d.r_win[ri * 3 + 2] = code == 3;
            } else atomicOr(d.err, ERRB_SAMPLES);
            if (accepted) {
                if (soff + nsamp <= d.s_cap) d.emit_off[g] = soff;
                else { d.emit_off[g] = -2; atomicOr(d.err, ERRB_SAMPLES); }
            } else d.emit_off[g] = -1;
        }
        __syncthreads();
        if (tid == 0) {
            long long acc_games = d.quota - gbase;
            if (acc_games < 0) acc_games = 0;
            if (acc_games > total_fin) acc_games = total_fin;
            s_base[0] = rbase + total_fin;
            s_base[1] = gbase + acc_games;
            s_base[2] += total_samp;
        }
        __syncthreads();
    }
    if (tid == 0) {
        long long rc = s_base[0] < d.r_cap ? s_base[0] : d.r_cap;
        d.counters->results += s_base[0] - d.counters->result_count;
        d.counters->result_count = rc;
        d.counters->games_played = s_base[1];
        long long sc = s_base[2] < d.s_cap ? s_base[2] : d.s_cap;
        d.counters->samples_total += s_base[2] - d.counters->sample_count;
        d.counters->sample_count = sc;
    }
}

// Sample emission with symmetries and game/tree restart for the finished games.
template <class G>
__global__ void __launch_bounds__(CTA_THREADS) k_emit(DevView d)
{
    constexpr int L = G::LANES;
    int g, lane, gi; unsigned gmask;
    if (!group_setup<G>(0, d.B, g, lane, gmask, gi)) return;
    if (d.finished[g] != 1) return;
    const long long off = d.emit_off[g];
    if (off == -1) {                    // beyond the quota: the game stays finished
        if (lane == 0) d.finished[g] = 2;
        return;
    }
    const int code = d.fin_code[g];
    const int hl = d.hist_len[g];
    const int per = d.symmetric ? G::NSYM : 1;
    if (off >= 0) {
        for (int h = 0; h < hl; h++) {
            const GState hs = d.hist_state[(size_t)g * d.hist_cap + h];
            const float *hp = d.hist_pi + ((size_t)g * d.hist_cap + h) * G::A;
            for (int k = 0; k < per; k++) {
                const long long si = off + (long long)h * per + k;
                const GState ss = G::symmetry(hs, k);
                float *o = d.s_obs + (size_t)si * G::OBS;
                for (int i = lane; i < G::OBS; i += L) o[i] = G::obs_value(ss, i);
                float *pp = d.s_pi + (size_t)si * G::A;
                for (int a = lane; a < G::A; a += L) pp[G::sym_action(k, a)] = hp[a];
                if (lane == 0) {
                    d.s_z[si * 3 + 0] = code == 1 ? 1.0f : 0.0f;
                    d.s_z[si * 3 + 1] = code == 2 ? 1.0f : 0.0f;
                    d.s_z[si * 3 + 2] = code == 3 ? 1.0f : 0.0f;
                    d.s_slot[si] = g;
                }
            }
        }
    }
    __syncwarp(gmask);
    if (lane == 0) {
        GState st; G::init(st);
        d.state[g] = st;
        d.hist_len[g] = 0;
        tree_reset(d, g);
        d.finished[g] = 0;
        d.fin_code[g] = 0;
    }
}

// MCTS.counts of every root (host introspection)
template <class G>
__global__ void k_root_counts(DevView d, int *out)
{
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= d.B) return;
    const size_t nb = (size_t)g * (size_t)d.npg;
    const int root = d.root[g];
    const int C = meta_nc(d.meta[nb + root]);
    const size_t cb = nb + (size_t)d.child0[nb + root];
    for (int a = 0; a < G::A; a++) out[(size_t)g * G::A + a] = 0;
    for (int k = 0; k < C; k++) out[(size_t)g * G::A + meta_action(d.meta[cb + k])] = d.n[cb + k];
}

template <class G>
__global__ void k_boards(DevView d, int8_t *out)
{
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= d.B) return;
    const GState st = d.state[g];
    for (int i = 0; i < G::CELLS; i++) out[(size_t)g * G::CELLS + i] = (int8_t)G::cell_code(st, i);
}

template <class G>
__global__ void k_init_slots(DevView d, const uint32_t *mt_seeds)
{
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= d.B) return;
    GState st; G::init(st);
    d.state[g] = st;
    tree_reset(d, g);
    d.hist_len[g] = 0; d.next_reset[g] = 0; d.noise_event[g] = 0; d.last_action[g] = -1;
    d.finished[g] = 0; d.fin_code[g] = 0; d.emit_off[g] = -1;
    d.ctr[g] = 0ULL;
    SlotStats z = {};
    z.peak_nodes = 1;
    d.stats[g] = z;
    if (d.mt != nullptr) {
        uint32_t s = mt_seeds ? mt_seeds[g] : (uint32_t)(d.seed + (unsigned long long)d.gid_base + (unsigned long long)g);
        mt_seed(s, d.mt + (size_t)g * 625);
    }
}

}  // namespace azb
