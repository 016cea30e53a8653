/*
 * azb_oracle.c -- CPU restatement (plain scalar C) of the reference's batched
 * self-play MCTS path.  TEST INFRASTRUCTURE ONLY -- see azb_oracle.h.
 *
 * Every function cites the reference file:line it follows (paths relative to
 * the reference tree).  Data structures are deliberately the reference's
 * (heap nodes with child lists, integer cell boards): the CUDA engine uses a
 * node pool and bitboards, so agreement between the two is a real check.
 */
#include "azb_oracle.h"
#include "orc_game.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* ======================================================================= */
/* numpy float32 helpers                                                    */
/* ======================================================================= */

/* numpy add.reduce over a contiguous float32 vector (pairwise_sum in
 * numpy/_core/src/umath/loops_utils.h.src, called from MCTS.pyx:245,252,320,321
 * through np.sum).  Verified bit-exact against NumPy 2.3.5. */
static float np_pairwise_f32(const float *a, int n)
{
    if (n < 8) {
        float r = 0.0f;
        for (int i = 0; i < n; i++) r += a[i];
        return r;
    } else if (n <= 128) {
        float r[8];
        int i;
        for (int j = 0; j < 8; j++) r[j] = a[j];
        for (i = 8; i < n - (n % 8); i += 8)
            for (int j = 0; j < 8; j++) r[j] += a[i + j];
        float res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
        for (; i < n; i++) res += a[i];
        return res;
    } else {
        int n2 = n / 2;
        n2 -= n2 % 8;
        return np_pairwise_f32(a, n2) + np_pairwise_f32(a + n2, n - n2);
    }
}

float orc_np_sum_f32(const float *a, int n) { return np_pairwise_f32(a, n); }

/* float32 power as this framework defines it: the correctly rounded result,
 * obtained through a double pow (MCTS.pyx:250 root temperature, :320 probs).
 * NumPy's own float32 power is SIMD-dispatch dependent (SVML on AVX-512, libm
 * otherwise; 1-ulp differences on ~20% of inputs, measured), so the reference
 * is patched at exactly this operation when it is compared (tests/_refdriver). */
float orc_pow_f32(float x, float e)
{
    if (e == 1.0f) return x;
    return (float)pow((double)x, (double)e);
}

/* ======================================================================= */
/* RNG: numpy legacy RandomState (MT19937) and the Philox4x32-10 stream     */
/* ======================================================================= */

void orc_mt_seed(uint32_t seed, uint32_t *st)
{
    /* np.random.seed(int) -> init_genrand (numpy/random/src/mt19937/mt19937.c) */
    st[0] = seed;
    for (int i = 1; i < 624; i++)
        st[i] = 1812433253u * (st[i - 1] ^ (st[i - 1] >> 30)) + (uint32_t)i;
    st[624] = 624;
}

uint32_t orc_mt_next(uint32_t *st)
{
    if (st[624] >= 624) {
        int k;
        uint32_t y;
        for (k = 0; k < 624 - 397; k++) {
            y = (st[k] & 0x80000000u) | (st[k + 1] & 0x7fffffffu);
            st[k] = st[k + 397] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
        }
        for (; k < 623; k++) {
            y = (st[k] & 0x80000000u) | (st[k + 1] & 0x7fffffffu);
            st[k] = st[k + (397 - 624)] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
        }
        y = (st[623] & 0x80000000u) | (st[0] & 0x7fffffffu);
        st[623] = st[396] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
        st[624] = 0;
    }
    uint32_t y = st[st[624]++];
    y ^= (y >> 11);
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= (y >> 18);
    return y;
}

static void philox4x32_10(const uint32_t ctr_in[4], const uint32_t key_in[2], uint32_t out[4])
{
    uint32_t c0 = ctr_in[0], c1 = ctr_in[1], c2 = ctr_in[2], c3 = ctr_in[3];
    uint32_t k0 = key_in[0], k1 = key_in[1];
    for (int r = 0; r < 10; r++) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

/* word w of the stream of (seed, game_id): block w/4, lane w%4; counter =
 * (block_lo, block_hi, game_lo, game_hi), key = (seed_lo, seed_hi). */
static uint32_t philox_word(uint64_t seed, uint64_t gid, uint64_t w)
{
    uint32_t ctr[4] = { (uint32_t)(w >> 2), (uint32_t)(w >> 34), (uint32_t)gid, (uint32_t)(gid >> 32) };
    uint32_t key[2] = { (uint32_t)seed, (uint32_t)(seed >> 32) };
    uint32_t out[4];
    philox4x32_10(ctr, key, out);
    return out[w & 3];
}

void orc_philox_words(uint64_t seed, uint64_t gid, uint64_t first, int n, uint32_t *out)
{
    for (int i = 0; i < n; i++) out[i] = philox_word(seed, gid, first + (uint64_t)i);
}

typedef struct rng_t {
    int mode;
    uint32_t mt[625];
    uint64_t seed, gid, ctr;     /* philox */
} rng_t;

static uint32_t rng_u32(rng_t *r)
{
    if (r->mode == ORC_RNG_MT19937) return orc_mt_next(r->mt);
    return philox_word(r->seed, r->gid, r->ctr++);
}

/* legacy_double / random_sample: 53-bit uniform from two 32-bit draws
 * (numpy/random/src/mt19937/mt19937.h mt19937_next_double) */
static double rng_double(rng_t *r)
{
    uint32_t a = rng_u32(r) >> 5, b = rng_u32(r) >> 6;
    return (a * 67108864.0 + b) / 9007199254740992.0;
}

double orc_mt_double(uint32_t *st)
{
    uint32_t a = orc_mt_next(st) >> 5, b = orc_mt_next(st) >> 6;
    return (a * 67108864.0 + b) / 9007199254740992.0;
}

/* random_interval (numpy/random/src/distributions/distributions.c): masked
 * rejection on 32-bit draws, used by RandomState.shuffle for Python lists. */
static uint32_t mt_interval(uint32_t *st, uint32_t max)
{
    if (max == 0) return 0;
    uint32_t mask = max, v;
    mask |= mask >> 1; mask |= mask >> 2; mask |= mask >> 4; mask |= mask >> 8; mask |= mask >> 16;
    while ((v = (orc_mt_next(st) & mask)) > max) { }
    return v;
}

void orc_mt_shuffle(uint32_t *st, int n, int32_t *perm)
{
    for (int i = 0; i < n; i++) perm[i] = i;
    for (int i = n - 1; i >= 1; i--) {
        uint32_t j = mt_interval(st, (uint32_t)i);
        int32_t t = perm[i]; perm[i] = perm[j]; perm[j] = t;
    }
}

/* Child-order permutation for a freshly expanded node with c children
 * (Node.add_children, MCTS.pyx:76-79: np.random.shuffle of the new list).
 * order[k] = index (in ascending-action order) of the child at position k.
 *  MT mode:     numpy legacy list shuffle, reversed Fisher-Yates.
 *  Philox mode: child j draws key word j; children are stably sorted by key
 *               (a uniformly random permutation that needs no serial loop on
 *               the GPU); c words are consumed. */
static void rng_child_order(rng_t *r, int c, int32_t *order)
{
    if (r->mode == ORC_RNG_MT19937) {
        orc_mt_shuffle(r->mt, c, order);
        return;
    }
    uint32_t key[ORC_MAX_CHILDREN];
    for (int j = 0; j < c; j++) key[j] = rng_u32(r);
    for (int j = 0; j < c; j++) {
        int rank = 0;
        for (int i = 0; i < c; i++)
            rank += (key[i] < key[j]) || (key[i] == key[j] && i < j);
        order[rank] = j;
    }
}

/* ======================================================================= */
/* MCTS tree (MCTS.pyx:49-104, 119-344)                                     */
/* ======================================================================= */

typedef struct node_t {
    struct node_t *children;   /* _children, contiguous, in list order */
    int32_t nchildren;
    int32_t a;                 /* action that leads here */
    uint8_t e[3];              /* win state recorded at first visit */
    int32_t n;
    float q, v, p;
    int32_t player;
} node_t;

static void node_init(node_t *nd, int a)
{
    memset(nd, 0, sizeof(*nd));
    nd->a = a;
}

static void node_free_children(node_t *nd)
{
    for (int i = 0; i < nd->nchildren; i++) node_free_children(&nd->children[i]);
    free(nd->children);
    nd->children = NULL;
    nd->nchildren = 0;
}

static int node_terminal(const node_t *nd) { return nd->e[0] | nd->e[1] | nd->e[2]; }

typedef struct mcts_t {
    node_t root;               /* MCTS._root (owned copy) */
    node_t *cur;               /* MCTS._curnode */
    node_t *path[ORC_MAX_PATH];/* MCTS._path */
    int path_len;
} mcts_t;

typedef struct slot_t {
    orc_game game;             /* live game (SelfPlayAgent.games[i]) */
    orc_game leaf;             /* state returned by find_leaf */
    mcts_t m[2];               /* SelfPlayAgent.mcts[i]: one tree, or in arena mode one per player
                                * (SelfPlayAgent._get_mcts, SelfPlayAgent.pyx:62-66) */
    rng_t rng;
    /* SelfPlayAgent per-game state */
    int hist_len;
    orc_game *hist_state;      /* histories[i][k][0] */
    float *hist_pi;            /* histories[i][k][1], [k][A] */
    int hist_cap;
    int next_reset;            /* SelfPlayAgent.next_reset[i] */
    int noise_event;           /* root expansions seen (index into fed noise) */
    int last_action;
} slot_t;

struct orc_agent {
    orc_args args;
    double *temp_table;
    const orc_game_ops *ops;
    int A, obs;
    slot_t *slots;
    const float *noise; int noise_events, noise_stride;
    float *noise_own;
    orc_stats st;
    /* queues */
    float *s_obs, *s_pi, *s_z; int32_t *s_slot; int64_t s_n, s_cap;
    int32_t *r_slot, *r_turns; uint8_t *r_win; int64_t r_n, r_cap;
};

/* SelfPlayAgent._mcts (SelfPlayAgent.pyx:68-73): the tree of the player to move in arena mode */
static mcts_t *slot_mcts(const orc_agent *ag, slot_t *s)
{
    return &s->m[ag->args.arena ? s->game.player : 0];
}

/* Node.add_children (MCTS.pyx:76-79) */
static void add_children(orc_agent *ag, slot_t *s, node_t *nd, const uint8_t *valid)
{
    int acts[ORC_MAX_CHILDREN], c = 0;
    for (int a = 0; a < ag->A; a++)
        if (valid[a]) {
            if (c >= ORC_MAX_CHILDREN) { fprintf(stderr, "orc: too many children\n"); abort(); }
            acts[c++] = a;
        }
    int32_t order[ORC_MAX_CHILDREN];
    rng_child_order(&s->rng, c, order);
    nd->children = c ? (node_t *)malloc(sizeof(node_t) * (size_t)c) : NULL;
    nd->nchildren = c;
    for (int k = 0; k < c; k++) node_init(&nd->children[k], acts[order[k]]);
    ag->st.nodes_created += c;
}

/* Node.best_child + Node.uct (MCTS.pyx:86-104).  Arithmetic as in the C that
 * Cython generates: seen_policy is Python's sum() over doubles (Neumaier
 * compensated since CPython 3.12, Python/bltinmodule.c) cast to float; the fpu
 * value is evaluated in double and cast to float; uct is float32. */
static node_t *best_child(const orc_agent *ag, node_t *nd)
{
    double f = 0.0, comp = 0.0;
    for (int i = 0; i < nd->nchildren; i++) {
        const node_t *c = &nd->children[i];
        if (c->n > 0) {
            double x = (double)c->p, t = f + x;
            if (fabs(f) >= fabs(x)) comp += (f - t) + x; else comp += (x - t) + f;
            f = t;
        }
    }
    if (comp != 0.0 && isfinite(comp)) f += comp;
    float seen_policy = (float)f;
    float fpu_value = (float)((double)nd->v - (double)ag->args.fpu_reduction * sqrt((double)seen_policy));
    float cur_best = -INFINITY;
    float sqrt_n = (float)sqrt((double)nd->n);
    node_t *best = NULL;
    for (int i = 0; i < nd->nchildren; i++) {
        node_t *c = &nd->children[i];
        float t1 = ag->args.cpuct * c->p;
        float t2 = t1 * sqrt_n;
        float t3 = t2 / (float)(1 + c->n);
        float u = (c->n == 0 ? fpu_value : c->q) + t3;
        if (u > cur_best) { cur_best = u; best = c; }
    }
    return best;
}

/* MCTS.find_leaf (MCTS.pyx:208-228) */
static void find_leaf(orc_agent *ag, slot_t *s)
{
    int depth = 0;
    mcts_t *m = slot_mcts(ag, s);
    m->cur = &m->root;
    s->leaf = s->game;                       /* gs.clone() */
    while (m->cur->n > 0 && !node_terminal(m->cur)) {
        if (m->path_len >= ORC_MAX_PATH) { fprintf(stderr, "orc: path overflow\n"); abort(); }
        m->path[m->path_len++] = m->cur;
        ag->st.sum_children += m->cur->nchildren;
        m->cur = best_child(ag, m->cur);
        ag->ops->play(&s->leaf, m->cur->a);
        depth++;
    }
    ag->st.sum_depth += depth;
    if (m->cur->n == 0) {
        uint8_t valid[ORC_MAX_ACTIONS];
        m->cur->player = s->leaf.player;
        ag->ops->win_state(&s->leaf, m->cur->e);
        ag->ops->valid_moves(&s->leaf, valid);
        add_children(ag, s, m->cur, valid);
        /* children of a terminal leaf are never used; the engine does not
         * materialise them, so they are left out of the statistic */
        if (node_terminal(m->cur)) ag->st.nodes_created -= m->cur->nchildren;
    }
    if (node_terminal(m->cur)) ag->st.terminal_leaves++;
}

/* MCTS._get_value (MCTS.pyx:291-295) with value.size == num_players + 1 */
static float get_value(const float *value, int player)
{
    float share = value[2] / (float)2;
    return value[player] + share;
}

/* MCTS.process_results (MCTS.pyx:230-289) incl. _add_root_noise (:197-206) */
static void process_results(orc_agent *ag, slot_t *s, const float *value_in, const float *pi_in)
{
    float value[3];
    mcts_t *m = slot_mcts(ag, s);
    node_t *cur = m->cur;
    if (node_terminal(cur)) {
        for (int i = 0; i < 3; i++) value[i] = (float)cur->e[i];
    } else {
        float pi[ORC_MAX_ACTIONS];
        memcpy(value, value_in, sizeof(value));
        float valids[ORC_MAX_ACTIONS];
        for (int a = 0; a < ag->A; a++) valids[a] = 0.0f;
        for (int i = 0; i < cur->nchildren; i++) valids[cur->children[i].a] = 1.0f;
        for (int a = 0; a < ag->A; a++) pi[a] = pi_in[a] * valids[a];   /* pi *= valids */
        float sum = np_pairwise_f32(pi, ag->A);
        for (int a = 0; a < ag->A; a++) pi[a] = pi[a] / sum;
        if (cur == &m->root) {
            if (ag->args.add_root_temp && !ag->args.arena) {
                /* 1.0 / self.root_temp is a Python float; NumPy casts the weak
                 * scalar exponent to float32 */
                float e = (float)(1.0 / (double)ag->args.root_policy_temp);
                for (int a = 0; a < ag->A; a++) pi[a] = orc_pow_f32(pi[a], e);
                sum = np_pairwise_f32(pi, ag->A);
                for (int a = 0; a < ag->A; a++) pi[a] = pi[a] / sum;
            }
            for (int i = 0; i < cur->nchildren; i++) cur->children[i].p = pi[cur->children[i].a];
            if (ag->args.add_root_noise && !ag->args.arena) {
                if (!ag->noise || s->noise_event >= ag->noise_events || cur->nchildren > ag->noise_stride) {
                    fprintf(stderr, "orc: root noise requested but not fed (event %d)\n", s->noise_event);
                    abort();
                }
                const float *nz = ag->noise + ((size_t)(s - ag->slots) * ag->noise_events + s->noise_event) * ag->noise_stride;
                float keep = 1 - ag->args.root_noise_frac;
                for (int i = 0; i < cur->nchildren; i++) {
                    float x = cur->children[i].p * keep;
                    float y = ag->args.root_noise_frac * nz[i];
                    cur->children[i].p = x + y;
                }
            }
            s->noise_event++;
        } else {
            for (int i = 0; i < cur->nchildren; i++) cur->children[i].p = pi[cur->children[i].a];
        }
    }
    /* backup.  min_discount ** (i / _discount_max_depth) is a C integer
     * division with cdivision=True and i < _discount_max_depth always, so the
     * discount is exactly 1 whatever min_discount is (MCTS.pyx:270-277). */
    while (m->path_len) {
        node_t *parent = m->path[--m->path_len];
        float v = get_value(value, parent->player);
        float qn = cur->q * (float)cur->n;
        float vd = v * 1.0f;
        float num = qn + vd;
        cur->q = num / (float)(cur->n + 1);
        if (cur->n == 0) cur->v = get_value(value, cur->player);
        cur->n += 1;
        cur = parent;
    }
    m->cur = cur;
    m->root.n += 1;
}

/* MCTS.update_root (MCTS.pyx:185-195); returns -1 on ValueError */
static int update_root(orc_agent *ag, slot_t *s, mcts_t *m, int a)
{
    if (m->root.nchildren == 0) {
        uint8_t valid[ORC_MAX_ACTIONS];
        ag->ops->valid_moves(&s->game, valid);
        add_children(ag, s, &m->root, valid);
        /* only the arena's idle tree gets here with numMCTSSims >= 1; its new children are dropped at once by the
         * re-root below and the engine does not materialise them, so they are left out of the statistic */
        if (ag->args.arena) ag->st.nodes_created -= m->root.nchildren;
    }
    for (int i = 0; i < m->root.nchildren; i++) {
        if (m->root.children[i].a == a) {
            node_t keep = m->root.children[i];
            m->root.children[i].children = NULL;
            m->root.children[i].nchildren = 0;
            node_free_children(&m->root);
            m->root = keep;
            return 0;
        }
    }
    return -1;
}

static void mcts_reset(slot_t *s)
{
    for (int t = 0; t < 2; t++) {
        mcts_t *m = &s->m[t];
        node_free_children(&m->root);
        node_init(&m->root, -1);
        m->cur = &m->root;
        m->path_len = 0;
    }
}

/* MCTS.probs (MCTS.pyx:308-327) */
static void mcts_probs(const orc_agent *ag, const mcts_t *s, float temp, float *probs)
{
    int A = ag->A;
    float counts[ORC_MAX_ACTIONS];
    for (int a = 0; a < A; a++) counts[a] = 0.0f;
    for (int i = 0; i < s->root.nchildren; i++) counts[s->root.children[i].a] = (float)s->root.children[i].n;
    if (temp == 0) {
        int best = 0;
        for (int a = 1; a < A; a++) if (counts[a] > counts[best]) best = a;
        for (int a = 0; a < A; a++) probs[a] = 0.0f;
        probs[best] = 1.0f;
        return;
    }
    float sum = np_pairwise_f32(counts, A);
    float e = (float)(1.0 / (double)temp);
    for (int a = 0; a < A; a++) probs[a] = orc_pow_f32(counts[a] / sum, e);
    sum = np_pairwise_f32(probs, A);
    for (int a = 0; a < A; a++) probs[a] = probs[a] / sum;
}

/* np.random.choice(A, p=policy) of the legacy RandomState (numpy/random/
 * mtrand.pyx choice): cdf = cumsum(p as f64) / last; one 53-bit uniform;
 * searchsorted side='right' */
static int rng_choice(rng_t *r, const float *p, int A)
{
    double cdf[ORC_MAX_ACTIONS], acc = 0.0;
    for (int a = 0; a < A; a++) { acc += (double)p[a]; cdf[a] = acc; }
    double last = cdf[A - 1];
    for (int a = 0; a < A; a++) cdf[a] /= last;
    double u = rng_double(r);
    int lo = 0, hi = A;
    while (lo < hi) { int mid = (lo + hi) >> 1; if (u < cdf[mid]) hi = mid; else lo = mid + 1; }
    return lo;
}

/* ======================================================================= */
/* agent                                                                    */
/* ======================================================================= */

static void push_sample(orc_agent *ag, const orc_game *g, const float *pi, const uint8_t *win, int slot)
{
    if (ag->s_n == ag->s_cap) {
        ag->s_cap = ag->s_cap ? ag->s_cap * 2 : 1024;
        ag->s_obs = (float *)realloc(ag->s_obs, sizeof(float) * (size_t)ag->s_cap * ag->obs);
        ag->s_pi = (float *)realloc(ag->s_pi, sizeof(float) * (size_t)ag->s_cap * ag->A);
        ag->s_z = (float *)realloc(ag->s_z, sizeof(float) * (size_t)ag->s_cap * 3);
        ag->s_slot = (int32_t *)realloc(ag->s_slot, sizeof(int32_t) * (size_t)ag->s_cap);
    }
    ag->ops->observation(g, ag->s_obs + (size_t)ag->s_n * ag->obs);
    memcpy(ag->s_pi + (size_t)ag->s_n * ag->A, pi, sizeof(float) * (size_t)ag->A);
    for (int i = 0; i < 3; i++) ag->s_z[ag->s_n * 3 + i] = (float)win[i];
    ag->s_slot[ag->s_n] = slot;
    ag->s_n++;
    ag->st.samples++;
}

static void push_result(orc_agent *ag, int slot, int turns, const uint8_t *win)
{
    if (ag->r_n == ag->r_cap) {
        ag->r_cap = ag->r_cap ? ag->r_cap * 2 : 256;
        ag->r_slot = (int32_t *)realloc(ag->r_slot, sizeof(int32_t) * (size_t)ag->r_cap);
        ag->r_turns = (int32_t *)realloc(ag->r_turns, sizeof(int32_t) * (size_t)ag->r_cap);
        ag->r_win = (uint8_t *)realloc(ag->r_win, (size_t)ag->r_cap * 3);
    }
    ag->r_slot[ag->r_n] = slot;
    ag->r_turns[ag->r_n] = turns;
    memcpy(ag->r_win + ag->r_n * 3, win, 3);
    ag->r_n++;
    ag->st.results++;
}

orc_agent *orc_create(const orc_args *args)
{
    const orc_game_ops *ops = orc_get_game_ops(args->game);
    if (!ops || args->num_slots <= 0) return NULL;
    orc_agent *ag = (orc_agent *)calloc(1, sizeof(*ag));
    ag->args = *args;
    ag->ops = ops;
    ag->A = ops->action_size;
    ag->obs = ops->obs_size;
    int tl = args->temp_table_len > 0 ? args->temp_table_len : 1;
    ag->temp_table = (double *)malloc(sizeof(double) * (size_t)tl);
    if (args->temp_table_len > 0) memcpy(ag->temp_table, args->temp_table, sizeof(double) * (size_t)tl);
    else ag->temp_table[0] = 1.0;
    ag->args.temp_table = ag->temp_table;
    ag->args.temp_table_len = tl;
    ag->slots = (slot_t *)calloc((size_t)args->num_slots, sizeof(slot_t));
    for (int i = 0; i < args->num_slots; i++) {
        slot_t *s = &ag->slots[i];
        ops->init(&s->game);
        for (int t = 0; t < 2; t++) { node_init(&s->m[t].root, -1); s->m[t].cur = &s->m[t].root; }
        s->rng.mode = args->rng_mode;
        s->rng.seed = args->seed;
        s->rng.gid = (uint64_t)(args->game_id_base + i);
        s->rng.ctr = 0;
        uint32_t ms = args->mt_seeds ? args->mt_seeds[i] : (uint32_t)(args->seed + (uint64_t)args->game_id_base + (uint64_t)i);
        orc_mt_seed(ms, s->rng.mt);
        s->last_action = -1;
    }
    ag->args.mt_seeds = NULL;
    return ag;
}

void orc_destroy(orc_agent *ag)
{
    if (!ag) return;
    for (int i = 0; i < ag->args.num_slots; i++) {
        node_free_children(&ag->slots[i].m[0].root);
        node_free_children(&ag->slots[i].m[1].root);
        free(ag->slots[i].hist_state);
        free(ag->slots[i].hist_pi);
    }
    free(ag->slots); free(ag->temp_table); free(ag->noise_own);
    free(ag->s_obs); free(ag->s_pi); free(ag->s_z); free(ag->s_slot);
    free(ag->r_slot); free(ag->r_turns); free(ag->r_win);
    free(ag);
}

int orc_action_size(const orc_agent *ag) { return ag->A; }
int orc_obs_size(const orc_agent *ag) { return ag->obs; }

void orc_set_root_noise(orc_agent *ag, const float *noise, int events, int stride)
{
    size_t n = (size_t)ag->args.num_slots * (size_t)events * (size_t)stride;
    free(ag->noise_own);
    ag->noise_own = (float *)malloc(sizeof(float) * n);
    memcpy(ag->noise_own, noise, sizeof(float) * n);
    ag->noise = ag->noise_own;
    ag->noise_events = events;
    ag->noise_stride = stride;
    for (int i = 0; i < ag->args.num_slots; i++) ag->slots[i].noise_event = 0;
}

/* arena: env player to move in every slot (its tree searches; player_to_index maps it to a model) */
void orc_players(const orc_agent *ag, int32_t *players)
{
    for (int i = 0; i < ag->args.num_slots; i++) players[i] = ag->slots[i].game.player;
}

/* SelfPlayAgent.generateBatch (SelfPlayAgent.pyx:103-135; the per-model regrouping of arena mode is the caller's) */
void orc_generate_batch(orc_agent *ag, float *obs_out)
{
    for (int i = 0; i < ag->args.num_slots; i++) {
        slot_t *s = &ag->slots[i];
        find_leaf(ag, s);
        if (obs_out) ag->ops->observation(&s->leaf, obs_out + (size_t)i * ag->obs);
    }
}

/* SelfPlayAgent.processBatch (SelfPlayAgent.pyx:137-151) */
void orc_process_batch(orc_agent *ag, const float *policy, const float *value)
{
    for (int i = 0; i < ag->args.num_slots; i++) {
        process_results(ag, &ag->slots[i], value + (size_t)i * 3, policy + (size_t)i * ag->A);
        ag->st.sims++;
    }
}

/* SelfPlayAgent.playMoves (SelfPlayAgent.pyx:153-202) */
void orc_play_moves(orc_agent *ag, int fast)
{
    int A = ag->A;
    const int arena = ag->args.arena;
    float policy[ORC_MAX_ACTIONS];
    for (int i = 0; i < ag->args.num_slots; i++) {
        slot_t *s = &ag->slots[i];
        mcts_t *m = slot_mcts(ag, s);
        int t = s->game.turns;
        if (t >= ag->args.temp_table_len) t = ag->args.temp_table_len - 1;
        /* probs(gs, float temp); in arena mode the table holds args.arenaTemp (:157-158) */
        float temp = (float)ag->temp_table[t];
        mcts_probs(ag, m, temp, policy);
        int action = rng_choice(&s->rng, policy, A);
        if (!fast && !arena) {
            if (s->hist_len == s->hist_cap) {
                s->hist_cap = s->hist_cap ? s->hist_cap * 2 : 64;
                s->hist_state = (orc_game *)realloc(s->hist_state, sizeof(orc_game) * (size_t)s->hist_cap);
                s->hist_pi = (float *)realloc(s->hist_pi, sizeof(float) * (size_t)s->hist_cap * A);
            }
            s->hist_state[s->hist_len] = s->game;
            mcts_probs(ag, m, 1.0f, s->hist_pi + (size_t)s->hist_len * A);
            s->hist_len++;
        }
        /* arena: [mcts.update_root(game, action) for mcts in self.mcts[i]] -- both trees, in tuple order (:167-168) */
        for (int tr = 0; tr < (arena ? 2 : 1); tr++) {
            if (update_root(ag, s, arena ? &s->m[tr] : m, action) != 0) {
                fprintf(stderr, "orc: invalid action %d while updating root (slot %d)\n", action, i);
                abort();
            }
        }
        ag->ops->play(&s->game, action);
        s->last_action = action;
        ag->st.moves++;
        if (ag->args.mcts_reset_threshold && s->game.turns >= s->next_reset) {
            mcts_reset(s);
            s->next_reset = s->game.turns + ag->args.mcts_reset_threshold;
        }
        uint8_t win[3];
        ag->ops->win_state(&s->game, win);
        if (win[0] | win[1] | win[2]) {
            push_result(ag, i, s->game.turns, win);
            if (ag->st.games_played < ag->args.games_per_iteration) {
                ag->st.games_played++;
                for (int h = 0; h < s->hist_len; h++) {
                    const float *pi = s->hist_pi + (size_t)h * A;
                    if (ag->args.symmetric_samples) {
                        int ns = ag->ops->num_symmetries;
                        for (int k = 0; k < ns; k++) {
                            orc_game g2; float pi2[ORC_MAX_ACTIONS];
                            ag->ops->symmetry(&s->hist_state[h], pi, k, &g2, pi2);
                            push_sample(ag, &g2, pi2, win, i);
                        }
                    } else {
                        push_sample(ag, &s->hist_state[h], pi, win, i);
                    }
                }
                ag->ops->init(&s->game);
                s->hist_len = 0;
                mcts_reset(s);
            }
        }
    }
}

void orc_root_counts(const orc_agent *ag, int32_t *counts)
{
    memset(counts, 0, sizeof(int32_t) * (size_t)ag->args.num_slots * ag->A);
    for (int i = 0; i < ag->args.num_slots; i++) {
        const node_t *r = &slot_mcts(ag, &ag->slots[i])->root;
        for (int k = 0; k < r->nchildren; k++) counts[(size_t)i * ag->A + r->children[k].a] = r->children[k].n;
    }
}

void orc_last_actions(const orc_agent *ag, int32_t *actions)
{
    for (int i = 0; i < ag->args.num_slots; i++) actions[i] = ag->slots[i].last_action;
}

void orc_turns(const orc_agent *ag, int32_t *turns)
{
    for (int i = 0; i < ag->args.num_slots; i++) turns[i] = ag->slots[i].game.turns;
}

void orc_boards(const orc_agent *ag, int8_t *cells)
{
    int n = ag->ops->num_cells;
    for (int i = 0; i < ag->args.num_slots; i++) ag->ops->cells(&ag->slots[i].game, cells + (size_t)i * n);
}

void orc_get_stats(const orc_agent *ag, orc_stats *st) { *st = ag->st; }

int64_t orc_num_samples(const orc_agent *ag) { return ag->s_n; }

void orc_get_samples(const orc_agent *ag, float *obs, float *pi, float *z, int32_t *slot)
{
    if (obs) memcpy(obs, ag->s_obs, sizeof(float) * (size_t)ag->s_n * ag->obs);
    if (pi) memcpy(pi, ag->s_pi, sizeof(float) * (size_t)ag->s_n * ag->A);
    if (z) memcpy(z, ag->s_z, sizeof(float) * (size_t)ag->s_n * 3);
    if (slot) memcpy(slot, ag->s_slot, sizeof(int32_t) * (size_t)ag->s_n);
}

void orc_clear_samples(orc_agent *ag) { ag->s_n = 0; }

int64_t orc_num_results(const orc_agent *ag) { return ag->r_n; }

void orc_get_results(const orc_agent *ag, int32_t *slot, int32_t *turns, uint8_t *win)
{
    if (slot) memcpy(slot, ag->r_slot, sizeof(int32_t) * (size_t)ag->r_n);
    if (turns) memcpy(turns, ag->r_turns, sizeof(int32_t) * (size_t)ag->r_n);
    if (win) memcpy(win, ag->r_win, (size_t)ag->r_n * 3);
}

int orc_rules_play_from(int game, const int8_t *cells, int turns, const int32_t *actions, int n, int8_t *cells_out,
                        uint8_t *valid_out, uint8_t *win_out, float *obs_out)
{
    const orc_game_ops *ops = orc_get_game_ops(game);
    if (!ops) return -2;
    orc_game g;
    ops->init(&g);
    if (cells) {                      /* a constructed position: the game's cell codes, `turns` plies played */
        memcpy(g.cells, cells, (size_t)ops->num_cells);
        g.turns = turns;
        g.player = turns % 2;
    }
    for (int i = 0; i < n; i++)
        if (ops->play(&g, actions[i]) != 0) return -1;
    if (cells_out) ops->cells(&g, cells_out);
    if (valid_out) ops->valid_moves(&g, valid_out);
    if (win_out) ops->win_state(&g, win_out);
    if (obs_out) ops->observation(&g, obs_out);
    return 0;
}

int orc_rules_symmetry(int game, const int8_t *cells, int turns, const float *pi, int k, int8_t *cells_out, float *pi_out)
{
    const orc_game_ops *ops = orc_get_game_ops(game);
    if (!ops || k < 0 || k >= ops->num_symmetries) return -2;
    orc_game g, g2;
    ops->init(&g);
    memcpy(g.cells, cells, (size_t)ops->num_cells);
    g.turns = turns;
    g.player = turns % 2;
    ops->symmetry(&g, pi, k, &g2, pi_out);
    ops->cells(&g2, cells_out);
    return 0;
}

int orc_rules_play(int game, const int32_t *actions, int n, int8_t *cells_out,
                   uint8_t *valid_out, uint8_t *win_out, float *obs_out)
{
    return orc_rules_play_from(game, NULL, 0, actions, n, cells_out, valid_out, win_out, obs_out);
}
