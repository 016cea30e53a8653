// azb_kernels.cuh -- the hot kernels of the batched self-play MCTS engine,
// templated on the game rules (G = Connect4T<L> | Brandubh).
//
// Work decomposition: G::LANES consecutive threads of a warp (a "group") own
// one game slot; child k of the node being scanned lives in lane k % LANES, so
// the sibling block (16 B hot records) is one coalesced read and the argmax is
// a shuffle reduction.  There is exactly one simulation in flight per game
// (the reference has no virtual loss, SelfPlayAgent.pyx:108-110), so the
// parallelism is across the B games.
//
// Control flow is warp-uniform: the 32/LANES games of a warp step through the
// same code under per-group predicates, so every collective uses the full
// mask (segmented by `width`) instead of sub-warp masks.
//
//   select_game         MCTS.find_leaf             MCTS.pyx:208-228, :86-104, :76-79
//   expand_backup_game  MCTS.process_results       MCTS.pyx:230-289, :197-206, :291-295
//   play_move_game      SelfPlayAgent.playMoves    SelfPlayAgent.pyx:153-176, MCTS.pyx:185-195, :297-327
//   k_finalize / k_emit SelfPlayAgent.playMoves    SelfPlayAgent.pyx:176-202 (quota, samples, reset)
#pragma once
#include "azb_common.cuh"

namespace azb {

constexpr int CTA_THREADS = 128;       // default threads per CTA (G::CTA: games with large per-group scratch use fewer)
constexpr unsigned FULL = 0xffffffffu;

template <class G>
struct GroupSmem {
    float vec[G::A];        // action-indexed scratch (masked priors / counts)
    float vec2[G::A];       // action-indexed scratch (probabilities)
    uint32_t key[G::MAXC];  // Philox sort keys / gamma variates
    short act[G::MAXC];     // valid actions, ascending
    short order[G::MAXC];   // order[k] = index into act[] of the child at position k
};

// group-wide helpers (all 32 lanes of the warp must call them)
template <int L>
__device__ __forceinline__ unsigned group_ballot(bool p, int sub)
{
    const unsigned b = __ballot_sync(FULL, p);
    if constexpr (L == 32) return b;
    else return (b >> (sub * L)) & ((1u << L) - 1u);
}
template <int L, class T>
__device__ __forceinline__ T group_bcast(T v, int src) { return __shfl_sync(FULL, v, src, L); }

template <class G> using HeadOf = SlotHeadT<typename G::State>;
template <class G>
__device__ __forceinline__ HeadOf<G> *head_at(const DevView &d, int g) { return reinterpret_cast<HeadOf<G> *>(d.head) + g; }
template <class G>
__device__ __forceinline__ typename G::State *hist_at(const DevView &d, size_t i) { return reinterpret_cast<typename G::State *>(d.hist_state) + i; }

template <class H>
__device__ __forceinline__ H load_head(const H *p)
{
    static_assert(sizeof(H) % 16 == 0, "slot header is copied in 16-byte pieces");
    H h;
    const int4 *s = reinterpret_cast<const int4 *>(p);
    int4 *dst = reinterpret_cast<int4 *>(&h);
#pragma unroll
    for (int i = 0; i < (int)(sizeof(H) / 16); i++) dst[i] = s[i];
    return h;
}
// the 32 bytes of root fields behind the game state
template <class H>
__device__ __forceinline__ void store_head_tail(H *p, const H &h)
{
    constexpr int N = (int)(sizeof(H) / 16);
    int4 *dst = reinterpret_cast<int4 *>(p);
    const int4 *s = reinterpret_cast<const int4 *>(&h);
    dst[N - 2] = s[N - 2]; dst[N - 1] = s[N - 1];
}
template <class H>
__device__ __forceinline__ void store_head(H *p, const H &h)
{
    int4 *dst = reinterpret_cast<int4 *>(p);
    const int4 *s = reinterpret_cast<const int4 *>(&h);
#pragma unroll
    for (int i = 0; i < (int)(sizeof(H) / 16); i++) dst[i] = s[i];
}
__device__ __forceinline__ NodeHot load_hot(const NodeHot *p)
{
    const int4 v = *reinterpret_cast<const int4 *>(p);
    NodeHot h; h.n = v.x; h.q = __int_as_float(v.y); h.p = __int_as_float(v.z); h.child0 = v.w;
    return h;
}
__device__ __forceinline__ NodeCold load_cold(const NodeCold *p)
{
    const int2 v = *reinterpret_cast<const int2 *>(p);
    NodeCold c; c.v = __int_as_float(v.x); c.meta = (uint32_t)v.y;
    return c;
}

// ------------------------------------------------------------------------------
// numpy float32 add.reduce (pairwise_sum) over an action-indexed vector in
// shared memory; every lane of the group returns the same value.
// ------------------------------------------------------------------------------
__device__ __forceinline__ float np_leaf_sum(const float *a, int n)
{
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

// n = 588 (brandubh): the recursion splits into eight leaf blocks
// [72 x7, 84]; block b, accumulator j is an independent serial chain, so the 64
// chains run two per lane and are combined in the recursion's order:
// ((B0+B1)+(B2+B3)) + ((B4+B5)+(B6+B7)).
__device__ __forceinline__ float np_sum_588_warp(const float *a, int lane)
{
    float blk[2];
#pragma unroll
    for (int h = 0; h < 2; h++) {
        const int b = (lane >> 3) + 4 * h, j = lane & 7;
        const int off = 72 * b, rows = (b == 7) ? 10 : 9;
        float r = a[off + j];
        for (int i = 1; i < rows; i++) r = f_add(r, a[off + 8 * i + j]);
        // ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)) within the 8 lanes of the block
        float t = f_add(r, __shfl_xor_sync(FULL, r, 1));      // lanes j even hold r_j + r_{j+1}
        float u = f_add(t, __shfl_xor_sync(FULL, t, 2));      // j % 4 == 0: (r_j+r_j+1)+(r_j+2+r_j+3)
        float w = f_add(u, __shfl_xor_sync(FULL, u, 4));      // j == 0: full block sum
        if (b == 7) {                                         // tail of the 84-block: 4 serial adds
            for (int i = 80; i < 84; i++) w = f_add(w, a[off + i]);
        }
        blk[h] = w;                                           // valid in lanes with j == 0
    }
    // lanes 0, 8, 16, 24 hold B0..B3 (blk[0]) and B4..B7 (blk[1])
    float s0 = f_add(blk[0], __shfl_xor_sync(FULL, blk[0], 8));    // lane 0: B0+B1, lane 16: B2+B3
    float s1 = f_add(blk[1], __shfl_xor_sync(FULL, blk[1], 8));    // lane 0: B4+B5, lane 16: B6+B7
    float t0 = f_add(s0, __shfl_xor_sync(FULL, s0, 16));           // lane 0: (B0+B1)+(B2+B3)
    float t1 = f_add(s1, __shfl_xor_sync(FULL, s1, 16));           // lane 0: (B4+B5)+(B6+B7)
    return __shfl_sync(FULL, f_add(t0, t1), 0);
}

// n = 2420 (hnefatafl): the recursion is a perfect binary tree of depth 5 whose 32 leaves have 72, 80 or 84 elements
// (azb_hnefatafl.cuh np2420_leaf): lane b sums leaf b with the leaf loop, the tree is five xor-shuffle steps
// (level 0 joins neighbouring leaves first; a + b is commutative, only the tree shape matters).
__device__ __forceinline__ float np_sum_2420_warp(const float *a, int lane)
{
    int off = 0, n = 2420;
#pragma unroll
    for (int level = 4; level >= 0; level--) {
        int n2 = n / 2;
        n2 -= n2 % 8;
        if ((lane >> level) & 1) { off += n2; n -= n2; }
        else n = n2;
    }
    float r = np_leaf_sum(a + off, n);
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) r = f_add(r, __shfl_xor_sync(FULL, r, o));
    return r;
}

template <class G>
__device__ __forceinline__ float np_sum_group(const float *vec, int lane)
{
    __syncwarp();
    float r;
    if (G::A < 8) {
        r = np_leaf_sum(vec, G::A);          // every lane, broadcast reads
    } else if (G::A == 588 && G::LANES == 32) {
        r = np_sum_588_warp(vec, lane);
    } else if (G::A == 2420 && G::LANES == 32) {
        r = np_sum_2420_warp(vec, lane);
    } else {
        r = 0.0f;
        if (lane == 0) r = np_sum_serial(vec, G::A);
        r = group_bcast<G::LANES>(r, 0);
    }
    return r;
}

// ------------------------------------------------------------------------------
// Child order of a freshly expanded node with C children (Node.add_children,
// MCTS.pyx:76-79).  Lane slot i (child j = lane + i*L of the ascending valid
// list) receives the list position cpos[i] and the action cact[i] it writes.
//   MT mode     : numpy legacy list shuffle (reversed Fisher-Yates, masked
//                 rejection) drawn serially by lane 0; lane takes position j.
//   Philox mode : child j takes key word ctr+j; its position is the rank of
//                 (key, j) -- computed in parallel.
// `consume` replays the RNG consumption (terminal leaves shuffle their unused
// children too, MCTS.pyx:223-226); `store` also produces the order.
// ------------------------------------------------------------------------------
template <class G, int IT>
__device__ __forceinline__ void child_order(const DevView &d, int g, int C, bool consume, bool store, int lane,
                                            GroupSmem<G> &sm, uint32_t valid_mask, int (&cpos)[IT], int (&cact)[IT],
                                            bool (&cwr)[IT])
{
    constexpr int L = G::LANES;
#pragma unroll
    for (int i = 0; i < IT; i++) { cpos[i] = lane + i * L; cact[i] = 0; cwr[i] = store && (lane + i * L < C); }
    if (d.rng_mode == 0) {
        if (consume && lane == 0) {
            uint32_t *st = d.mt + (size_t)rng_slot(d, g) * 625;
            if (store) for (int k = 0; k < C; k++) sm.order[k] = (short)k;
            for (int i = C - 1; i >= 1; i--) {
                uint32_t j = mt_interval(st, (uint32_t)i);
                if (store) { short t = sm.order[i]; sm.order[i] = sm.order[j]; sm.order[j] = t; }
            }
        }
        __syncwarp();
        if (store) {
#pragma unroll
            for (int i = 0; i < IT; i++) {
                const int k = lane + i * L;
                if (k < C) cact[i] = G::nth_valid(sm.act, valid_mask, sm.order[k]);
            }
        }
    } else {
        unsigned long long c0 = 0ULL;
        if (consume) c0 = d.ctr[rng_slot(d, g)];
        __syncwarp();
        if (consume && lane == 0) d.ctr[rng_slot(d, g)] = c0 + (unsigned long long)C;
        uint32_t key[IT];
        if constexpr (G::LANE_IS_ACTION) {
            // small action space: lane a owns the child of action a; its ascending index is the
            // number of valid actions below it
            const bool valid = (valid_mask >> lane) & 1u;
            const int j = __popc(valid_mask & ((1u << lane) - 1u));
            key[0] = (store && valid) ? philox_word(d.seed, rng_gid(d, g), c0 + (unsigned long long)j) : 0u;
            int rank = 0;
#pragma unroll
            for (int i = 0; i < G::A; i++) {
                const uint32_t ki = group_bcast<L>(key[0], i);
                rank += ((valid_mask >> i) & 1u) && ((ki < key[0]) || (ki == key[0] && i < lane));
            }
            cpos[0] = rank; cact[0] = lane; cwr[0] = store && valid;
            return;
        }
#pragma unroll
        for (int i = 0; i < IT; i++) {
            const int j = lane + i * L;
            key[i] = (store && j < C) ? philox_word(d.seed, rng_gid(d, g), c0 + (unsigned long long)j) : 0u;
        }
        if (IT == 1) {
            int rank = 0;
#pragma unroll
            for (int i = 0; i < G::MAXC; i++) {
                const uint32_t ki = group_bcast<L>(key[0], i);
                rank += (i < C) && ((ki < key[0]) || (ki == key[0] && i < lane));
            }
            cpos[0] = rank;
            if (store && lane < C) cact[0] = G::nth_valid(sm.act, valid_mask, lane);
        } else {
            if (store) {
#pragma unroll
                for (int i = 0; i < IT; i++) { const int j = lane + i * L; if (j < C) sm.key[j] = key[i]; }
            }
            __syncwarp();
            if (store) {
#pragma unroll
                for (int i = 0; i < IT; i++) {
                    const int j = lane + i * L;
                    if (j < C) {
                        int rank = 0;
                        for (int t = 0; t < C; t++) {
                            const uint32_t kt = sm.key[t];
                            rank += (kt < key[i]) || (kt == key[i] && t < j);
                        }
                        cpos[i] = rank;
                        cact[i] = G::nth_valid(sm.act, valid_mask, j);
                    }
                }
            }
            __syncwarp();
        }
    }
}

// ------------------------------------------------------------------------------
// seen_policy = float(sum(p of visited children)) as CPython's sum() computes it
// (double accumulation in child order, Neumaier-compensated; MCTS.pyx:91).
// Fast path: when every visited prior is 0 or >= 2^-28 the double partial sums
// are exact, so an order-free integer reduction gives the identical result.
// Lanes whose slot is out of range pass n = 0.
// ------------------------------------------------------------------------------
template <class G, int IT>
__device__ __forceinline__ float seen_policy(const NodeHot (&kh)[IT], const bool (&kin)[IT], int C, int lane, int sub)
{
    constexpr int L = G::LANES;
    bool ok = true;
    unsigned long long acc = 0ULL;
#pragma unroll
    for (int i = 0; i < IT; i++) {
        if (kin[i] && kh[i].n > 0) {
            const float pv = kh[i].p;
            ok = ok && ((pv == 0.0f) || (pv >= 3.7252902984619140625e-09f && pv < 2.0f));
            acc += (unsigned long long)(pv * 2251799813685248.0f);   // p * 2^51, exact
        }
    }
    const unsigned bad = group_ballot<L>(!ok, sub);
#pragma unroll
    for (int off = L / 2; off >= 1; off >>= 1) acc += __shfl_xor_sync(FULL, acc, off, L);
    float fast = (float)((double)acc * 4.44089209850062616169452667236328125e-16);   // * 2^-51
    if (__any_sync(FULL, bad != 0u)) {
        // general path (some group holds a tiny or out-of-range prior): serial compensated sum
        double f = 0.0, comp = 0.0;
        for (int k = 0; k < G::MAXC; k++) {
            const int i = k / L, src = k - i * L;
            int nn = 0; float pv = 0.0f;
#pragma unroll
            for (int ii = 0; ii < IT; ii++) if (ii == i) { nn = kin[ii] ? kh[ii].n : 0; pv = kh[ii].p; }
            nn = __shfl_sync(FULL, nn, src, L);
            pv = __shfl_sync(FULL, pv, src, L);
            if (k < C && nn > 0) {
                const double x = (double)pv, t = __dadd_rn(f, x);
                if (fabs(f) >= fabs(x)) comp = __dadd_rn(comp, __dadd_rn(__dsub_rn(f, t), x));
                else comp = __dadd_rn(comp, __dadd_rn(__dsub_rn(x, t), f));
                f = t;
            }
        }
        if (comp != 0.0 && isfinite(comp)) f = __dadd_rn(f, comp);
        if (bad != 0u) fast = (float)f;
    }
    return fast;
}

// ------------------------------------------------------------------------------
// Leaf de-duplication (DevView::dd_table): one thread of a game whose leaf needs the network publishes the leaf's state
// and returns the slot whose policy / value rows answer for it -- its own (it is the first game of this select launch with
// that state: the representative, the only one listed for the evaluator) or the representative's.  The answers are
// bit-identical either way: the evaluator's arithmetic for a row does not depend on where the row sits in the batch
// (tests: compact == dense, CTA pairs == single CTAs, persistent == one-round plans).
//   representative: state -> dd_state[g]; __threadfence; CAS (entry of another epoch -> epoch | tag | g)
//   duplicate:      sees an entry of this epoch with its tag; __threadfence; reads dd_state[rep] from L2; equal keys
// Open addressing, linear probing, load factor <= 1/2.  A full table (cannot happen) degrades to "evaluate it".
// ------------------------------------------------------------------------------
template <class S>
__device__ __noinline__ int dedup_leaf(unsigned long long *table, unsigned mask, unsigned epoch32, S *states, int g, const S st)
{
    using K = LeafKey<S>;
    static_assert(sizeof(S) % 16 == 0, "state copied in 16-byte pieces");
    unsigned long long key[K::N];
    K::get(st, key);
    unsigned long long h = 0x9E3779B97F4A7C15ULL;
#pragma unroll
    for (int i = 0; i < K::N; i++) {               // splitmix-style mixing of the key words
        h ^= key[i] + 0x9E3779B97F4A7C15ULL + (h << 6) + (h >> 2);
        h *= 0xBF58476D1CE4E5B9ULL;
        h ^= h >> 29;
    }
    S *mine = states + g;
    {
        const uint4 *src = reinterpret_cast<const uint4 *>(&st);
        uint4 *dst = reinterpret_cast<uint4 *>(mine);
#pragma unroll
        for (int i = 0; i < (int)(sizeof(S) / 16); i++) __stcg(dst + i, src[i]);
    }
    __threadfence();
    const unsigned long long epoch = (unsigned long long)epoch32 << 32;
    const unsigned long long tag = ((h >> 44) & 0xFFFULL) << 20;
    const unsigned long long entry = epoch | tag | (unsigned long long)g;
    unsigned slot = (unsigned)h & mask;
    for (unsigned probe = 0; probe <= 2u * mask + 1u; probe++) {
        unsigned long long cur = *reinterpret_cast<volatile unsigned long long *>(table + slot);
        if ((cur & 0xFFFFFFFF00000000ULL) != epoch) {                     // free in this epoch
            const unsigned long long old = atomicCAS(table + slot, cur, entry);
            if (old == cur) return g;                                      // representative
            cur = old;
            if ((cur & 0xFFFFFFFF00000000ULL) != epoch) continue;          // (only entries of this epoch are ever written)
        }
        if ((cur & (0xFFFULL << 20)) == tag) {
            const int r = (int)(cur & 0xFFFFFULL);
            __threadfence();
            const uint4 *rp = reinterpret_cast<const uint4 *>(states + r);
            S other;
            uint4 *op = reinterpret_cast<uint4 *>(&other);
#pragma unroll
            for (int i = 0; i < (int)(sizeof(S) / 16); i++) op[i] = __ldcg(rp + i);
            unsigned long long ok[K::N];
            K::get(other, ok);
            bool same = true;
#pragma unroll
            for (int i = 0; i < K::N; i++) same = same && ok[i] == key[i];
            if (same) return r;
        }
        slot = (slot + 1u) & mask;
    }
    return g;
}

// ------------------------------------------------------------------------------
// MCTS.find_leaf
// ------------------------------------------------------------------------------
template <class G, bool WRITE_OBS>
__device__ __forceinline__ void select_game(const DevView &d, int g, bool active, int lane, int sub, GroupSmem<G> &sm)
{
    constexpr int L = G::LANES;
    constexpr int IT = (G::MAXC + L - 1) / L;
    constexpr bool SPEC = (G::MAXC <= L);       // small fan-out: prefetch the next sibling block speculatively
    const size_t nb = (size_t)g * (size_t)d.npg;
    const bool in_range = active;               // out-of-range groups alias slot `first`: they must not write
    HeadOf<G> H = load_head(head_at<G>(d, g));
    if (H.st.flags & (GF_FINISHED | GF_DEAD)) active = false;
    if (d.arena && ((g ^ H.st.turns) & 1)) active = false;      // arena: only the tree of the player to move searches
    typename G::State st = H.st;
    int *path = d.path + (size_t)g * G::MAXD;
    // this game's statistics row is updated at the very end: fetch it now, under the descent
    uint4 stat_row = make_uint4(0u, 0u, 0u, 0u);
    if (lane == 0 && in_range) stat_row = *reinterpret_cast<const uint4 *>(d.stats + g);
    int cur = H.root, cn = H.root_n, cch = H.root_child0;
    float cv = H.root_v;
    uint32_t cmeta = H.root_meta;
    int depth = 0, sumc = 0;
    bool desc = active && cn > 0 && meta_e(cmeta) == 0 && meta_nc(cmeta) > 0;

    NodeHot kh[IT];
    bool kin[IT];
#pragma unroll
    for (int i = 0; i < IT; i++) {
        const int k = lane + i * L;
        kin[i] = desc && k < meta_nc(cmeta);
        if (kin[i]) kh[i] = load_hot(d.hot + nb + cch + k);
        else { kh[i].n = 0; kh[i].q = 0.0f; kh[i].p = 0.0f; kh[i].child0 = -1; }
    }

    while (__any_sync(FULL, desc)) {
        const int C = meta_nc(cmeta);
        if (desc && lane == 0) path[depth] = cur;
        // Node.best_child: fpu value in double, uct in float32, first strict maximum
        const float seen = seen_policy<G, IT>(kh, kin, C, lane, sub);
        const float fpu = (float)__dsub_rn((double)cv, __dmul_rn((double)d.fpu_reduction, __dsqrt_rn((double)seen)));
        const float sqrt_n = __fsqrt_rn((float)cn);      // == (float)sqrt((double)n): double rounding is innocuous for sqrt
        float bu = -CUDART_INF_F; int bk = 0x7fffffff;
#pragma unroll
        for (int i = 0; i < IT; i++) {
            if (kin[i]) {
                const float t = f_div(f_mul(f_mul(d.cpuct, kh[i].p), sqrt_n), (float)(1 + kh[i].n));
                const float u = f_add(kh[i].n == 0 ? fpu : kh[i].q, t);
                if (u > bu) { bu = u; bk = lane + i * L; }
            }
        }
#pragma unroll
        for (int off = L / 2; off >= 1; off >>= 1) {
            const float ou = __shfl_xor_sync(FULL, bu, off, L);
            const int ok = __shfl_xor_sync(FULL, bk, off, L);
            if (ou > bu || (ou == bu && ok < bk)) { bu = ou; bk = ok; }
        }
        if (desc && bk == 0x7fffffff) {               // every uct was NaN
            if (lane == 0) atomicOr(d.err, ERRB_FP);
        }
        if (bk == 0x7fffffff) bk = 0;
        const int bi = bk / L, src = bk - bi * L;
        int sn = 0, sc = -1;
#pragma unroll
        for (int i = 0; i < IT; i++) if (i == bi) { sn = kh[i].n; sc = kh[i].child0; }
        const int nn = __shfl_sync(FULL, sn, src, L);
        const int nch = __shfl_sync(FULL, sc, src, L);
        const int idx = cch + bk;
        NodeCold cc; cc.v = 0.0f; cc.meta = 0u;
        if (desc) cc = load_cold(d.cold + nb + idx);                    // the chosen child's cold half ...
        NodeHot nh[IT];
        if constexpr (SPEC) {                                           // ... and, in flight with it, its children
            const bool pre = desc && nn > 0 && nch >= 0 && lane < G::MAXC;
            if (pre) nh[0] = load_hot(d.hot + nb + nch + lane);
            else { nh[0].n = 0; nh[0].q = 0.0f; nh[0].p = 0.0f; nh[0].child0 = -1; }
        }
        if (desc) {
            G::play(st, meta_action(cc.meta));
            sumc += C;
            depth++;
            cur = idx; cn = nn; cch = nch; cv = cc.v; cmeta = cc.meta;
        }
        desc = desc && cn > 0 && meta_e(cmeta) == 0 && meta_nc(cmeta) > 0 && depth < G::MAXD - 1;
#pragma unroll
        for (int i = 0; i < IT; i++) {
            const int k = lane + i * L;
            kin[i] = desc && k < meta_nc(cmeta);
            if constexpr (SPEC) {
                kh[i] = nh[i];
            } else {
                if (kin[i]) kh[i] = load_hot(d.hot + nb + cch + k);
                else { kh[i].n = 0; kh[i].q = 0.0f; kh[i].p = 0.0f; kh[i].child0 = -1; }
            }
        }
    }

    // first visit: record player and win state, materialise the children
    const bool expd = active && cn == 0;
    bool head_dirty = false;
    int nodes_new = 0;
    if (__any_sync(FULL, expd)) {
        const int player = G::player(st);
        const int e = G::win_code(st);
        uint32_t vmask = 0u;
        const int C = G::list_valid(st, sm.act, vmask, expd, lane);
        const bool term = e != 0;
        int base = -1, nc = 0;
        bool store = false;
        if (expd && !term) {
            base = H.alloc;
            if (base + C > (H.root >= d.half ? d.npg : d.half)) { if (lane == 0) atomicOr(d.err, ERRB_POOL); base = -1; }
            else { store = true; nc = C; }
        }
        int cpos[IT], cact[IT];
        bool cwr[IT];
        child_order<G, IT>(d, g, C, expd, store, lane, sm, vmask, cpos, cact, cwr);
        if (store) {
#pragma unroll
            for (int i = 0; i < IT; i++) {
                if (cwr[i]) {
                    const size_t idx = nb + (size_t)(base + cpos[i]);
                    *reinterpret_cast<int4 *>(d.hot + idx) = make_int4(0, 0, 0, -1);
                    *reinterpret_cast<int2 *>(d.cold + idx) = make_int2(0, (int)meta_pack((uint32_t)cact[i], 0u, 0u, 0u));
                }
            }
            H.alloc = base + C;
            nodes_new = C;
        }
        if (expd) {
            cmeta = meta_pack((uint32_t)meta_action(cmeta), (uint32_t)nc, (uint32_t)e, (uint32_t)player);
            cch = base;
            if (depth == 0) { H.root_meta = cmeta; H.root_child0 = base; }
            else if (lane == 0) { d.cold[nb + cur].meta = cmeta; d.hot[nb + cur].child0 = base; }
            head_dirty = true;
        }
    }
    if (WRITE_OBS) {
        if (active) G::write_obs(st, d.obs + (size_t)g * G::OBS, lane);
        // the leaves the network has to evaluate: a terminal leaf's value is its win state (MCTS.pyx:234-235).  One
        // atomic per warp and model, not per game: 8192 adds to one counter were a quarter of this kernel's stall samples
        // arena: one list per model (the model of env player p is p ^ arena_swap), each at most B / 2 long
        bool want = active && in_range && lane == 0 && meta_e(cmeta) == 0;
        if (d.dd_table != nullptr) {                  // leaf de-duplication: only the first game with this state is listed
            int src = g;
            if (want) {
                src = dedup_leaf<typename G::State>(d.dd_table, d.dd_mask, d.dd_epoch, reinterpret_cast<typename G::State *>(d.dd_state), g, st);
                d.dd_src[g] = src;
            }
            const bool dup = want && src != g;
            const unsigned db = __ballot_sync(FULL, dup);
            if (db != 0u && (threadIdx.x & 31u) == (unsigned)__ffs((int)db) - 1u) atomicAdd(d.dd_dups, (unsigned long long)__popc(db));
            if (dup) want = false;
        }
        const int m = d.arena ? ((g & 1) ^ d.arena_swap) : 0;
        const unsigned cap = d.arena ? (unsigned)(d.B / 2) : (unsigned)d.B;
        const unsigned wl = threadIdx.x & 31u;
        for (int mm = 0; mm < (d.arena ? 2 : 1); mm++) {
            const unsigned bal = __ballot_sync(FULL, want && m == mm);
            if (bal == 0u) continue;
            const unsigned leader = (unsigned)__ffs((int)bal) - 1u;
            unsigned base = 0u;
            if (wl == leader) base = (unsigned)atomicAdd(d.nn_count + 2 * d.nn_par + mm, (int)__popc(bal));
            base = __shfl_sync(FULL, base, (int)leader);
            if (want && m == mm) {
                const unsigned idx = base + (unsigned)__popc(bal & ((1u << wl) - 1u));
                if (idx < cap) d.nn_rows[(size_t)mm * cap + idx] = g;
            }
        }
    }
    if (lane == 0 && in_range) {
        LeafInfo li;
        li.leaf = active ? cur : -1; li.depth = depth; li.child0 = cch; li.meta = cmeta;
        *reinterpret_cast<int4 *>(d.leafinfo + g) = *reinterpret_cast<const int4 *>(&li);
        if (active) {
            if (head_dirty) store_head_tail(head_at<G>(d, g), H);
            uint4 s = stat_row;
            s.x += (unsigned)depth; s.y += (unsigned)sumc; s.z += (unsigned)nodes_new; s.w += (meta_e(cmeta) != 0) ? 1u : 0u;
            *reinterpret_cast<uint4 *>(d.stats + g) = s;
        }
    }
    __syncwarp();
}

// ------------------------------------------------------------------------------
// Device Dirichlet(alpha, ..., alpha) for the root noise when no table is fed
// (MCTS._add_root_noise, MCTS.pyx:197-206: alpha = 10.83 / C).  Child k draws a
// Gamma(alpha) variate (Marsaglia-Tsang, boosted for alpha < 1) from a private
// 64-word window of the slot's Philox stream; the variates are normalised by
// their sum.  Statistically equivalent to numpy's sampler, not bit-equal to it:
// parity runs feed the vectors instead (azb_set_root_noise).
// ------------------------------------------------------------------------------
__device__ __noinline__ float gamma_variate(unsigned long long seed, unsigned long long gid, unsigned long long w0, double alpha)
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

// root prior post-processing for child position k with prior pk (MCTS.pyx:247-256)
template <class G>
__device__ __forceinline__ float root_noise_mix(const DevView &d, int g, int C, int k, float pk, bool on, float gen_sum,
                                                float gen_val, const float *nz)
{
    const float keep = __fsub_rn(1.0f, d.noise_frac);
    if (!on) return pk;
    if (nz != nullptr) return f_add(f_mul(pk, keep), f_mul(d.noise_frac, nz[k]));
    if (gen_sum > 0.0f) return f_add(f_mul(pk, keep), f_mul(d.noise_frac, gen_val / gen_sum));
    return pk;
}

// ------------------------------------------------------------------------------
// MCTS.process_results
// ------------------------------------------------------------------------------
template <class G>
__device__ __forceinline__ void expand_backup_game(const DevView &d, int g, bool active, int lane, int sub,
                                                   GroupSmem<G> &sm, const float *pol_row, const float *val_row)
{
    constexpr int L = G::LANES;
    constexpr int IT = (G::MAXC + L - 1) / L;
    const size_t nb = (size_t)g * (size_t)d.npg;
    const int4 li4 = *reinterpret_cast<const int4 *>(d.leafinfo + g);
    const int4 hr = reinterpret_cast<const int4 *>(head_at<G>(d, g))[sizeof(typename G::State) / 16];   // root, root_n, root_v, root_child0
    const int leaf = li4.x, depth = li4.y, base = li4.z;
    const uint32_t lmeta = (uint32_t)li4.w;
    const bool on = active && leaf >= 0;
    const int e = meta_e(lmeta), C = meta_nc(lmeta);
    const bool term = e != 0;
    const bool is_root = leaf == hr.x;
    float val0, val1, val2;
    if (term) {
        // value = np.array(self._curnode.e, dtype=np.float32)
        val0 = e == 1 ? 1.0f : 0.0f; val1 = e == 2 ? 1.0f : 0.0f; val2 = e == 3 ? 1.0f : 0.0f;
    } else {
        val0 = val_row[0]; val1 = val_row[1]; val2 = val_row[2];
    }
    const bool ex = on && !term;                       // priors are written
    const bool rootx = ex && is_root;
    const size_t cb = nb + (size_t)(base < 0 ? 0 : base);
    // children's actions (cold halves of the sibling block)
    int ak[IT];
    bool kin[IT];
#pragma unroll
    for (int i = 0; i < IT; i++) {
        const int k = lane + i * L;
        kin[i] = ex && k < C;
        ak[i] = kin[i] ? meta_action(d.cold[cb + k].meta) : 0;
    }
    // fed noise row of this root expansion
    const float *nz = nullptr;
    bool gen = false;
    if (rootx && d.add_noise) {
        const int ev = d.noise_event[g];
        if (d.noise != nullptr) {
            if (ev < d.noise_events && C <= d.noise_stride) nz = d.noise + ((size_t)g * d.noise_events + ev) * d.noise_stride;
            else if (lane == 0) atomicOr(d.err, ERRB_NOISE);
        } else gen = true;
    }
    float gsum = 0.0f, gval[IT];
#pragma unroll
    for (int i = 0; i < IT; i++) gval[i] = 0.0f;
    if (__any_sync(FULL, gen)) {
        unsigned long long c0 = 0ULL;
        if (gen) c0 = d.ctr[g];
        __syncwarp();
        if (gen && lane == 0) d.ctr[g] = c0 + 64ULL * (unsigned long long)C;
#pragma unroll
        for (int i = 0; i < IT; i++) {
            const int k = lane + i * L;
            if (gen && k < C) gval[i] = gamma_variate(d.seed, (unsigned long long)(d.gid_base + g), c0 + 64ULL * k, 10.83 / (double)C);
            gsum += gval[i];
        }
#pragma unroll
        for (int off = L / 2; off >= 1; off >>= 1) gsum += __shfl_xor_sync(FULL, gsum, off, L);
    }

    if (G::A <= L) {
        // ---- small action space: the action-indexed vector lives in lanes 0..A-1 ----
        float pa = (ex && lane < G::A) ? pol_row[lane] : 0.0f;
        unsigned vm = kin[0] ? (1u << ak[0]) : 0u;
#pragma unroll
        for (int off = L / 2; off >= 1; off >>= 1) vm |= __shfl_xor_sync(FULL, vm, off, L);
        pa = ((vm >> lane) & 1u) ? pa : 0.0f;                    // pi *= valids
        float s = 0.0f;                                          // np.sum, n < 8: serial from 0
#pragma unroll
        for (int a = 0; a < G::A; a++) s = f_add(s, group_bcast<L>(pa, a));
        if (ex && C > 0 && !(s > 0.0f) && lane == 0) atomicOr(d.err, ERRB_FP);
        float pn = f_div(pa, s);                                 // pi /= np.sum(pi)
        if (__any_sync(FULL, rootx && d.add_temp)) {
            float pt = pow_det(pn, d.root_temp_exp);             // pi ** (1 / root_temp)
            float s2 = 0.0f;
#pragma unroll
            for (int a = 0; a < G::A; a++) s2 = f_add(s2, group_bcast<L>(pt, a));
            pt = f_div(pt, s2);
            if (rootx && d.add_temp) pn = pt;
        }
        float pk = group_bcast<L>(pn, ak[0]);                    // Node.update_policy: c.p = pi[c.a]
        pk = root_noise_mix<G>(d, g, C, lane, pk, kin[0] && rootx && d.add_noise, gsum, gval[0], nz);
        if (kin[0]) d.hot[cb + lane].p = pk;
    } else {
        // ---- large action space: action-indexed vector in shared memory ----
        // sm.vec is all zero outside the children's entries (zero-filled once per launch by the
        // kernel, restored below), so only the C child entries are ever touched; NumPy's
        // pairwise sum still reads the whole vector (adding +0.0 is exact).
        if (__any_sync(FULL, ex)) {
#pragma unroll
            for (int i = 0; i < IT; i++) if (kin[i]) sm.vec[ak[i]] = pol_row[ak[i]];
            float sum = np_sum_group<G>(sm.vec, lane);
            if (ex && C > 0 && !(sum > 0.0f) && lane == 0) atomicOr(d.err, ERRB_FP);
            float pk[IT];
#pragma unroll
            for (int i = 0; i < IT; i++) pk[i] = kin[i] ? f_div(sm.vec[ak[i]], sum) : 0.0f;     // pi /= np.sum(pi)
            if (__any_sync(FULL, rootx && d.add_temp)) {
                __syncwarp();
#pragma unroll
                for (int i = 0; i < IT; i++) if (kin[i]) sm.vec[ak[i]] = pow_det(pk[i], d.root_temp_exp);
                float sum2 = np_sum_group<G>(sm.vec, lane);
#pragma unroll
                for (int i = 0; i < IT; i++) if (kin[i]) pk[i] = f_div(sm.vec[ak[i]], sum2);
            }
            __syncwarp();
#pragma unroll
            for (int i = 0; i < IT; i++) {
                const int k = lane + i * L;
                if (kin[i]) {
                    sm.vec[ak[i]] = 0.0f;                                   // restore the all-zero invariant
                    d.hot[cb + k].p = root_noise_mix<G>(d, g, C, k, pk[i], rootx && d.add_noise, gsum, gval[i], nz);
                }
            }
            __syncwarp();
        }
    }
    if (rootx && lane == 0) d.noise_event[g] += 1;

    // backup along the stored path; level i updates the node entered at step i.
    // player(node at depth t) = (leaf player - depth + t) & 1: turns alternate.
    const int *path = d.path + (size_t)g * G::MAXD;
    const float share = f_mul(val2, 0.5f);              // value[num_players] / num_players (halving is exact)
    const int root_player = (meta_player(lmeta) - depth) & 1;
    if (on) {
        for (int i = lane; i < depth; i += L) {
            const int node = (i == depth - 1) ? leaf : path[i + 1];
            const int pp = (root_player + i) & 1;
            const float v = f_add(pp == 0 ? val0 : val1, share);
            NodeHot *hp = d.hot + nb + node;
            const int2 nq = *reinterpret_cast<const int2 *>(hp);
            const int nn = nq.x;
            const float qq = __int_as_float(nq.y);
            const float qn = f_div(f_add(f_mul(qq, (float)nn), f_mul(v, 1.0f)), (float)(nn + 1));
            if (nn == 0) d.cold[nb + node].v = f_add(pp == 0 ? val1 : val0, share);   // the node's own player = pp ^ 1
            *reinterpret_cast<int2 *>(hp) = make_int2(nn + 1, __float_as_int(qn));
        }
        if (lane == 0) {
            head_at<G>(d, g)->root_n = hr.y + 1;
            reinterpret_cast<unsigned *>(d.stats + g)[4] += 1u;     // sims
        }
    }
    __syncwarp();
}

// ------------------------------------------------------------------------------
// MCTS.probs on an action-indexed count vector in shared memory
// ------------------------------------------------------------------------------
template <class G>
__device__ __forceinline__ void probs_group(const DevView &d, const float *counts, float temp, float *out, bool on, int lane)
{
    constexpr int L = G::LANES;
    __syncwarp();
    // temp == 0: one-hot of np.argmax (first maximum)
    int best = 0;
    if (temp == 0.0f) for (int a = 1; a < G::A; a++) if (counts[a] > counts[best]) best = a;
    const float sum = np_sum_group<G>(counts, lane);
    if (on && temp != 0.0f && !(sum > 0.0f) && lane == 0) atomicOr(d.err, ERRB_FP);
    const float e = temp != 0.0f ? (float)(1.0 / (double)temp) : 1.0f;
    for (int a = lane; a < G::A; a += L) out[a] = temp == 0.0f ? (a == best ? 1.0f : 0.0f) : pow_det(f_div(counts[a], sum), e);
    const float sum2 = np_sum_group<G>(out, lane);
    if (temp != 0.0f) for (int a = lane; a < G::A; a += L) out[a] = f_div(out[a], sum2);
    __syncwarp();
}

template <class H_>
__device__ __forceinline__ void tree_reset(H_ &H)
{
    H.root = 0; H.root_n = 0; H.root_v = 0.0f; H.root_child0 = -1;
    H.root_meta = meta_pack(META_ACTION_NONE, 0u, 0u, 0u);
    H.alloc = 1; H.root_rec = 0;
}

// ------------------------------------------------------------------------------
// Re-root compaction.  A slot's arena is two halves of d.half entries; after MCTS.update_root (MCTS.pyx:185-195) the
// kept subtree is copied breadth-first into the other half -- the C children of a node stay one contiguous block in
// the reference's list order, so the scan is unchanged -- and the discarded siblings' subtrees (which the reference frees
// by refcount) are left behind: the live tree is always one dense prefix [base, base + live) of a half.
// Every lane of the warp calls this; `on` = this group compacts.  The new root is node `src_root` of the current half;
// its record is copied to entry 0 of the other half.  A batch of up to L copied nodes is expanded per step: lane j reads
// node head+j (its child0 still names the OLD block), an exclusive scan of the child counts places the new blocks, then
// the group copies block after block (16 B + 8 B per child, coalesced).  Returns the new live size; H.root /
// H.root_child0 / H.alloc are updated on every lane of the group.
// ------------------------------------------------------------------------------
template <class G>
__device__ __forceinline__ int compact_subtree(const DevView &d, size_t nb, HeadOf<G> &H, bool on, int lane)
{
    constexpr int L = G::LANES;
    const int dst = H.root >= d.half ? 0 : d.half;          // the half the tree moves to
    NodeHot *hot = d.hot + nb;
    NodeCold *cold = d.cold + nb;
    if (on && lane == 0) {                                   // the root's own record (q, p, action of the played child)
        NodeHot h = load_hot(hot + H.root);
        h.child0 = H.root_child0;
        *reinterpret_cast<int4 *>(hot + dst) = make_int4(h.n, __float_as_int(h.q), __float_as_int(h.p), h.child0);
        NodeCold c = load_cold(cold + H.root);
        c.meta = H.root_meta;
        *reinterpret_cast<int2 *>(cold + dst) = make_int2(__float_as_int(c.v), (int)c.meta);
    }
    __syncwarp();
    int head = 0, tail = on ? 1 : 0;
    while (__any_sync(FULL, head < tail)) {
        const int i = head + lane;
        int c0 = -1, C = 0;
        if (i < tail) {
            c0 = hot[dst + i].child0;
            C = c0 >= 0 ? meta_nc(cold[dst + i].meta) : 0;
        }
        // exclusive scan of C over the group's lanes
        int off = C;
#pragma unroll
        for (int o = 1; o < L; o <<= 1) {
            const int v = __shfl_up_sync(FULL, off, o, L);
            if (lane >= o) off += v;
        }
        const int total = __shfl_sync(FULL, off, L - 1, L);
        off -= C;
        if (i < tail && C > 0) hot[dst + i].child0 = dst + tail + off;
        const int nbatch = min(L, tail - head);
        for (int j = 0; j < L; j++) {
            if (!__any_sync(FULL, j < nbatch)) break;
            const int sc0 = __shfl_sync(FULL, c0, j, L), sC = __shfl_sync(FULL, C, j, L), so = __shfl_sync(FULL, off, j, L);
            if (j < nbatch) {
                for (int k = lane; k < sC; k += L) {
                    *reinterpret_cast<int4 *>(hot + dst + tail + so + k) = *reinterpret_cast<const int4 *>(hot + sc0 + k);
                    *reinterpret_cast<int2 *>(cold + dst + tail + so + k) = *reinterpret_cast<const int2 *>(cold + sc0 + k);
                }
            }
        }
        head += nbatch > 0 ? nbatch : 0;
        tail += total;
        __syncwarp();
    }
    if (on) {
        H.root_child0 = hot[dst].child0;
        H.root = dst;
        H.alloc = dst + tail;
    }
    return tail;
}

// is `action` one of the C valid moves listed by G::list_valid (every lane answers)
template <class G>
__device__ __forceinline__ bool vmask_has(const short *act, uint32_t vmask, int C, int action)
{
    if constexpr (G::LANE_IS_ACTION) return (vmask >> action) & 1u;
    for (int k = 0; k < C; k++) if (G::nth_valid(act, vmask, k) == action) return true;
    return false;
}

// ------------------------------------------------------------------------------
// SelfPlayAgent.playMoves, the per-game part up to the terminal test
// ------------------------------------------------------------------------------
// `forced` >= 0 (arena, the idle tree of the pair): no sampling -- the root follows the action the searching tree
// played (MCTS.update_root, MCTS.pyx:185-195; an unexpanded root first gets its children, i.e. their shuffle is
// drawn, and the chosen child is a fresh node).  Returns the action played (G::A if none).
template <class G>
__device__ __forceinline__ int play_move_game(const DevView &d, int g, bool active, int fast, int lane, int sub, GroupSmem<G> &sm,
                                              int forced = -1)
{
    constexpr int L = G::LANES;
    constexpr int IT = (G::MAXC + L - 1) / L;
    const size_t nb = (size_t)g * (size_t)d.npg;
    HeadOf<G> H = load_head(head_at<G>(d, g));
    bool on = active && !(H.st.flags & (GF_FINISHED | GF_DEAD));
    typename G::State st = H.st;
    const int C = on ? meta_nc(H.root_meta) : 0;
    const size_t cb = nb + (size_t)(H.root_child0 < 0 ? 0 : H.root_child0);
    // MCTS.counts
    for (int a = lane; a < G::A; a += L) sm.vec[a] = 0.0f;
    __syncwarp();
    int ka[IT], kn[IT];
#pragma unroll
    for (int i = 0; i < IT; i++) {
        const int k = lane + i * L;
        ka[i] = -1; kn[i] = 0;
        if (k < C) {
            ka[i] = meta_action(d.cold[cb + k].meta);
            kn[i] = d.hot[cb + k].n;
            sm.vec[ka[i]] = (float)kn[i];
        }
    }
    __syncwarp();
    const int t = st.turns < d.temp_len ? st.turns : d.temp_len - 1;
    const float temp = d.temp_table[t < 0 ? 0 : t];
    const bool sample = on && forced < 0;
    probs_group<G>(d, sm.vec, temp, sm.vec2, sample, lane);
    // np.random.choice(A, p=policy): cdf in double, one 53-bit uniform, searchsorted 'right'
    uint32_t wa = 0, wb = 0;
    if (sample && lane == 0) {
        const int rs = rng_slot(d, g);
        if (d.rng_mode == 0) {
            uint32_t *ms = d.mt + (size_t)rs * 625;
            wa = mt_next(ms); wb = mt_next(ms);
        } else {
            const unsigned long long c = d.ctr[rs];
            wa = philox_word(d.seed, rng_gid(d, g), c);
            wb = philox_word(d.seed, rng_gid(d, g), c + 1);
            d.ctr[rs] = c + 2;
        }
    }
    int action = G::A;
    if (on && forced >= 0) action = forced;
    if (sample && lane == 0) {
        const double u = u53(wa, wb);
        double last = 0.0;
        for (int a = 0; a < G::A; a++) last = __dadd_rn(last, (double)sm.vec2[a]);
        double acc = 0.0;
        for (int a = 0; a < G::A; a++) {
            acc = __dadd_rn(acc, (double)sm.vec2[a]);
            if (u < __ddiv_rn(acc, last)) { action = a; break; }
        }
    }
    action = group_bcast<L>(action, 0);
    // arena, idle tree whose root was never expanded: add_children draws the shuffle of the root's children, the
    // child `action` becomes the new (fresh) root
    const bool fresh = on && forced >= 0 && C == 0;
    if (__any_sync(FULL, fresh)) {
        uint32_t vmask = 0u;
        const int Cv = G::list_valid(st, sm.act, vmask, fresh, lane);
        int cpos[IT], cact[IT];
        bool cwr[IT];
        child_order<G, IT>(d, g, Cv, fresh, false, lane, sm, vmask, cpos, cact, cwr);
        if (fresh && !((vmask_has<G>(sm.act, vmask, Cv, action)))) {
            if (lane == 0) atomicOr(d.err, ERRB_ACTION);
        }
    }
    if (!fast) {
        // histories[i].append((game.clone(), mcts.probs(game)))  -- temp = 1
        const int hl = on ? d.hist_len[g] : 0;
        probs_group<G>(d, sm.vec, 1.0f, sm.vec2, on, lane);
        if (on) {
            if (hl < d.hist_cap) {
                float *hp = d.hist_pi + ((size_t)g * d.hist_cap + hl) * G::A;
                for (int a = lane; a < G::A; a += L) hp[a] = sm.vec2[a];
                if (lane == 0) { *hist_at<G>(d, (size_t)g * d.hist_cap + hl) = st; d.hist_len[g] = hl + 1; }
            } else if (lane == 0) atomicOr(d.err, ERRB_SAMPLES);
        }
    }
    // MCTS.update_root
    int found = -1;
#pragma unroll
    for (int i = 0; i < IT; i++) if (ka[i] == action && lane + i * L < C) found = lane + i * L;
#pragma unroll
    for (int off = L / 2; off >= 1; off >>= 1) found = max(found, __shfl_xor_sync(FULL, found, off, L));
    if (fresh) found = 0;
    if (on && found < 0) {
        if (lane == 0) atomicOr(d.err, ERRB_ACTION);
        H.st.flags |= GF_DEAD;
        if (lane == 0) store_head(head_at<G>(d, g), H);
        on = false;
    }
    bool compact = false;
    if (on && lane == 0) {
        if (fresh) {
            tree_reset(H);
            H.root_meta = meta_pack((uint32_t)action, 0u, 0u, 0u);
        } else {
            const int nr = H.root_child0 + found;
            const NodeHot h = load_hot(d.hot + nb + nr);
            const NodeCold c = load_cold(d.cold + nb + nr);
            H.root = nr; H.root_n = h.n; H.root_v = c.v; H.root_child0 = h.child0; H.root_meta = c.meta; H.root_rec = 1;
            compact = true;
        }
        G::play(st, action);
        const int e = G::win_code(st);
        st.flags &= 0xff;
        if (e != 0) { st.flags |= GF_FINISHED; d.fin_code[g] = e; compact = false; }      // the tree is dropped with the game
        H.st = st;
        d.last_action[g] = action;
        if (forced < 0) reinterpret_cast<unsigned *>(d.stats + g)[5] += 1u;     // moves
        if (d.reset_threshold && st.turns >= d.next_reset[g]) {
            tree_reset(H);
            d.next_reset[g] = st.turns + d.reset_threshold;
            compact = false;
        }
    }
    // peak of the live tree (entries in use in the current half), sampled before the tree shrinks
    if (on && lane == 0) {
        const int used = H.alloc - (H.root >= d.half ? d.half : 0);
        int *pk = reinterpret_cast<int *>(d.stats + g) + 6;
        if (!fresh && used > *pk) *pk = used;
    }
    compact = group_bcast<L>((int)compact, 0) != 0;
    if (__any_sync(FULL, compact)) {
        H.root = group_bcast<L>(H.root, 0);
        H.root_child0 = group_bcast<L>(H.root_child0, 0);
        H.root_meta = group_bcast<L>(H.root_meta, 0);
        compact_subtree<G>(d, nb, H, compact, lane);
    }
    if (on && lane == 0) store_head(head_at<G>(d, g), H);
    __syncwarp();
    return on ? action : G::A;
}

// ------------------------------------------------------------------------------
// kernels
// ------------------------------------------------------------------------------
template <class G>
__device__ __forceinline__ bool group_setup(int first, int count, int &g, bool &active, int &lane, int &sub, int &gi)
{
    constexpr int L = G::LANES;
    const int tid = blockIdx.x * blockDim.x + threadIdx.x;
    const int grp = tid / L;
    lane = threadIdx.x % L;
    gi = threadIdx.x / L;
    sub = (threadIdx.x % 32) / L;
    active = grp < count;
    g = first + (active ? grp : 0);
    return __any_sync(FULL, active);          // false: the whole warp is out of range
}

template <class G>
__global__ void __launch_bounds__(G::CTA) k_select(DevView d, int first, int count)
{
    __shared__ GroupSmem<G> sm[G::CTA / G::LANES];
    int g, lane, sub, gi; bool active;
    if (blockIdx.x == 0 && threadIdx.x == 0) { d.nn_count[2 * (d.nn_par ^ 1)] = 0; d.nn_count[2 * (d.nn_par ^ 1) + 1] = 0; }   // consumed
    if (!group_setup<G>(first, count, g, active, lane, sub, gi)) return;
    select_game<G, true>(d, g, active, lane, sub, sm[gi]);
}

template <class G>
__global__ void __launch_bounds__(G::CTA) k_expand_backup(DevView d, int first, int count, const float *policy, const float *value)
{
    __shared__ GroupSmem<G> sm[G::CTA / G::LANES];
    int g, lane, sub, gi; bool active;
    if (!group_setup<G>(first, count, g, active, lane, sub, gi)) return;
    if (G::A > G::LANES) { for (int a = lane; a < G::A; a += G::LANES) sm[gi].vec[a] = 0.0f; __syncwarp(); }
    const int src = d.dd_table != nullptr ? d.dd_src[g] : g;       // leaf de-duplication: the representative's rows
    expand_backup_game<G>(d, g, active, lane, sub, sm[gi], policy + (size_t)src * G::A, value + (size_t)src * 3);
}

// processBatch of simulation s fused with generateBatch of simulation s+1 for the same slots: one launch instead of
// two between two network evaluations, and the slot's header / leaf record stay in cache
template <class G>
__global__ void __launch_bounds__(G::CTA) k_expand_select(DevView d, int first, int count, const float *policy, const float *value)
{
    __shared__ GroupSmem<G> sm[G::CTA / G::LANES];
    int g, lane, sub, gi; bool active;
    if (blockIdx.x == 0 && threadIdx.x == 0) { d.nn_count[2 * (d.nn_par ^ 1)] = 0; d.nn_count[2 * (d.nn_par ^ 1) + 1] = 0; }   // consumed
    if (!group_setup<G>(first, count, g, active, lane, sub, gi)) return;
    if (G::A > G::LANES) { for (int a = lane; a < G::A; a += G::LANES) sm[gi].vec[a] = 0.0f; __syncwarp(); }
    const int src = d.dd_table != nullptr ? d.dd_src[g] : g;       // leaf de-duplication: the representative's rows
    expand_backup_game<G>(d, g, active, lane, sub, sm[gi], policy + (size_t)src * G::A, value + (size_t)src * 3);
    __syncwarp();
    select_game<G, true>(d, g, active, lane, sub, sm[gi]);
}

// `sims` simulations per game with constant NN outputs, no NN round trip
template <class G>
__global__ void __launch_bounds__(G::CTA) k_warmup_sims(DevView d, int sims)
{
    __shared__ GroupSmem<G> sm[G::CTA / G::LANES];
    int g, lane, sub, gi; bool active;
    if (!group_setup<G>(0, d.B, g, active, lane, sub, gi)) return;
    if (G::A > G::LANES) { for (int a = lane; a < G::A; a += G::LANES) sm[gi].vec[a] = 0.0f; __syncwarp(); }
    for (int s = 0; s < sims; s++) {
        select_game<G, false>(d, g, active, lane, sub, sm[gi]);
        expand_backup_game<G>(d, g, active, lane, sub, sm[gi], d.warm_policy, d.warm_value);
    }
}

template <class G>
__global__ void __launch_bounds__(G::CTA) k_play_moves(DevView d, int fast)
{
    __shared__ GroupSmem<G> sm[G::CTA / G::LANES];
    int g, lane, sub, gi; bool active;
    if (!group_setup<G>(0, d.B, g, active, lane, sub, gi)) return;
    play_move_game<G>(d, g, active, fast, lane, sub, sm[gi]);
}

// Arena: one group per game; the tree of the player to move samples the move, then the other tree follows it
// ([mcts.update_root(game, action) for mcts in self.mcts[i]], SelfPlayAgent.pyx:167-168 -- the searching tree's
// update never draws, so the order of the two updates does not matter for the game's RNG stream).
template <class G>
__global__ void __launch_bounds__(G::CTA) k_play_moves_arena(DevView d)
{
    __shared__ GroupSmem<G> sm[G::CTA / G::LANES];
    int g, lane, sub, gi; bool active;
    if (!group_setup<G>(0, d.B / 2, g, active, lane, sub, gi)) return;
    const int mover = head_at<G>(d, 2 * g)->st.turns & 1;
    const int action = play_move_game<G>(d, 2 * g + mover, active, 1, lane, sub, sm[gi]);
    play_move_game<G>(d, 2 * g + (mover ^ 1), active && action < G::A, 1, lane, sub, sm[gi], action < G::A ? action : 0);
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
        int flags = 0, turns = 0;
        if (g < d.B) { flags = head_at<G>(d, g)->st.flags; turns = head_at<G>(d, g)->st.turns; }
        const int fin = ((flags & GF_FINISHED) && !(d.arena && (g & 1))) ? 1 : 0;     // arena: a game is two slots
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
                d.r_slot[ri] = d.arena ? (g >> 1) : g;
                d.r_turns[ri] = turns;
                d.r_win[ri * 3 + 0] = code == 1; d.r_win[ri * 3 + 1] = code == 2; d.r_win[ri * 3 + 2] = code == 3;
            } else atomicOr(d.err, ERRB_SAMPLES);
            if (accepted) {
                if (soff + nsamp <= d.s_cap) d.emit_off[g] = soff;
                else { d.emit_off[g] = -2; atomicOr(d.err, ERRB_SAMPLES); }
            } else d.emit_off[g] = -1;
            if (d.arena) d.emit_off[g + 1] = d.emit_off[g];       // the second tree of the game restarts (or dies) with it
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
        d.nn_count[0] = d.nn_count[1] = d.nn_count[2] = d.nn_count[3] = 0;   // every move-round starts with parity 0 and empty lists
        long long rc = s_base[0] < d.r_cap ? s_base[0] : d.r_cap;
        d.counters->results += s_base[0] - d.counters->result_count;
        d.counters->result_count = rc;
        d.counters->games_played = s_base[1];
        long long sc = s_base[2] < d.s_cap ? s_base[2] : d.s_cap;
        d.counters->samples_total += s_base[2] - d.counters->sample_count;
        d.counters->sample_count = sc;
    }
}

// Sample emission with symmetries for the finished games: blockIdx.y = chunk of the game's history (a finished tafl game
// is up to 100 positions x 8 symmetries x (980 B observation + 588-entry policy scatter): one group per game took 3.4 ms per
// move-round of 4096 brandubh games, 3.7 % of the step).  Read-only on the game state; k_emit_reset restarts the games.
template <class G>
__global__ void __launch_bounds__(G::CTA) k_emit(DevView d)
{
    constexpr int L = G::LANES;
    const int tid = blockIdx.x * blockDim.x + threadIdx.x;
    const int g = tid / L, lane = threadIdx.x % L;
    if (g >= d.B) return;
    if (!(head_at<G>(d, g)->st.flags & GF_FINISHED)) return;
    const long long off = d.emit_off[g];
    if (off < 0) return;                // beyond the quota (-1) or no samples wanted
    const int code = d.fin_code[g];
    const int hl = d.hist_len[g];
    const int per = d.symmetric ? G::NSYM : 1;
    for (int h = (int)blockIdx.y; h < hl; h += (int)gridDim.y) {
        const typename G::State hs = *hist_at<G>(d, (size_t)g * d.hist_cap + h);
        const float *hp = d.hist_pi + ((size_t)g * d.hist_cap + h) * G::A;
        for (int k = 0; k < per; k++) {
            const long long si = off + (long long)h * per + k;
            const typename G::State ss = d.symmetric ? G::symmetry(hs, k) : hs;
            G::write_obs(ss, d.s_obs + (size_t)si * G::OBS, lane);
            float *pp = d.s_pi + (size_t)si * G::A;
            for (int a = lane; a < G::A; a += L) pp[d.symmetric ? G::sym_action(k, a) : a] = hp[a];
            if (lane == 0) {
                d.s_z[si * 3 + 0] = code == 1 ? 1.0f : 0.0f;
                d.s_z[si * 3 + 1] = code == 2 ? 1.0f : 0.0f;
                d.s_z[si * 3 + 2] = code == 3 ? 1.0f : 0.0f;
                d.s_slot[si] = g;
            }
        }
    }
}

// game / tree restart of the finished games (after k_emit has read their histories)
template <class G>
__global__ void __launch_bounds__(128) k_emit_reset(DevView d)
{
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= d.B) return;
    HeadOf<G> H = load_head(head_at<G>(d, g));
    if (!(H.st.flags & GF_FINISHED)) return;
    if (d.emit_off[g] == -1) {          // beyond the quota: the game stays finished
        H.st.flags = (H.st.flags & ~GF_FINISHED) | GF_DEAD;
        store_head(head_at<G>(d, g), H);
        return;
    }
    int *pk = reinterpret_cast<int *>(d.stats + g) + 6;
    const int used = H.alloc - (H.root >= d.half ? d.half : 0);
    if (used > *pk) *pk = used;
    G::init(H.st);
    tree_reset(H);
    store_head(head_at<G>(d, g), H);
    d.hist_len[g] = 0;
    d.fin_code[g] = 0;
}

// MCTS.counts of every root (host introspection)
template <class G>
__global__ void k_root_counts(DevView d, int *out)
{
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= d.B) return;
    const size_t nb = (size_t)g * (size_t)d.npg;
    const HeadOf<G> H = load_head(head_at<G>(d, g));
    const int C = meta_nc(H.root_meta);
    const size_t cb = nb + (size_t)(H.root_child0 < 0 ? 0 : H.root_child0);
    for (int a = 0; a < G::A; a++) out[(size_t)g * G::A + a] = 0;
    for (int k = 0; k < C; k++) out[(size_t)g * G::A + meta_action(d.cold[cb + k].meta)] = d.hot[cb + k].n;
}

// Single-tree API (MCTS.search on an arbitrary position): put a slot on a given position with an empty tree ...
template <class G>
__global__ void k_set_state(DevView d, int g, const signed char *cells, int turns)
{
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    HeadOf<G> H = load_head(head_at<G>(d, g));
    G::from_cells(H.st, cells, turns);
    tree_reset(H);
    store_head(head_at<G>(d, g), H);
    LeafInfo li; li.leaf = -1; li.depth = 0; li.child0 = -1; li.meta = 0u;
    d.leafinfo[g] = li;
    d.hist_len[g] = 0; d.next_reset[g] = 0; d.noise_event[g] = 0; d.last_action[g] = -1; d.fin_code[g] = 0; d.emit_off[g] = -1;
}

// ... and MCTS.update_root(gs, action) + gs.play_action(action) for one slot (MCTS.pyx:185-195): the root follows a
// move decided by the caller; an unexpanded root draws the shuffle of its children first.  No terminal handling:
// a finished game just stops being searched until the slot is set to a new position.
template <class G>
__global__ void __launch_bounds__(G::CTA) k_force_move(DevView d, int g0, int action)
{
    __shared__ GroupSmem<G> sm[G::CTA / G::LANES];
    int g, lane, sub, gi; bool active;
    if (!group_setup<G>(g0, 1, g, active, lane, sub, gi)) return;
    play_move_game<G>(d, g, active, 1, lane, sub, sm[gi], action);
}

// arena: which env player's tree searches in each slot this round (-1: idle tree / finished game)
template <class G>
__global__ void k_arena_players(DevView d, int *out)
{
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= d.B) return;
    const typename G::State st = head_at<G>(d, g)->st;
    const bool live = !(st.flags & (GF_FINISHED | GF_DEAD));
    out[g] = (live && !((g ^ st.turns) & 1)) ? (g & 1) : -1;
}

template <class G>
__global__ void k_boards(DevView d, int8_t *out)
{
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= d.B) return;
    const typename G::State st = head_at<G>(d, g)->st;
    for (int i = 0; i < G::CELLS; i++) out[(size_t)g * G::CELLS + i] = (int8_t)G::cell_code(st, i);
}

template <class G>
__global__ void k_init_slots(DevView d, const uint32_t *mt_seeds)
{
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= d.B) return;
    HeadOf<G> H;
    G::init(H.st);
    tree_reset(H);
    H.root_rec = H.pad1 = 0;
    store_head(head_at<G>(d, g), H);
    LeafInfo li; li.leaf = -1; li.depth = 0; li.child0 = -1; li.meta = 0u;
    d.leafinfo[g] = li;
    d.hist_len[g] = 0; d.next_reset[g] = 0; d.noise_event[g] = 0; d.last_action[g] = -1;
    d.fin_code[g] = 0; d.emit_off[g] = -1;
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
