// azb_common.cuh -- device-side data layout, RNG streams and the numeric
// helpers whose rounding must match the reference bit for bit.
//
// Layout in HBM (one engine = B game slots, NPG node entries per slot):
//   node pool, two parallel arrays, entry index = slot * NPG + local id; a slot's NPG entries are two halves and the
//     live tree is a dense prefix of one of them: re-rooting (playMoves) copies the kept subtree breadth-first into the
//     other half, so the discarded siblings never accumulate and the working set stays compact
//     hot  (16 B)  n int32, q float, p float, child0 int32     -- what the PUCT
//                  scan reads for every sibling (Node.n/.q/.p, MCTS.pyx:53-56)
//                  plus the link to the node's own children
//     cold ( 8 B)  v float, meta uint32 (action:10 | nchild:8 | e:2 | player:1)
//                  -- read only for the chosen child (Node.v/.a/.e/.player)
//     The C children of a node are contiguous, in the reference's shuffled
//     list order (Node._children, MCTS.pyx:50,76-79): one level of the scan is
//     one coalesced 16*C-byte read.
//   slot header (64 B, one cache-line half): packed game bitboards + the root's
//     own fields + allocator, so a simulation starts with a single load.
//   per slot besides: leaf record, path buffer, RNG stream, move history.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math_constants.h>

namespace azb {

// ---- node meta word --------------------------------------------------------
// action:12 (hnefatafl has 2420 actions) | nchild:8 | e:2 | player:1
constexpr uint32_t META_ACTION_MASK = 4095u;
constexpr uint32_t META_ACTION_NONE = 4095u;
__host__ __device__ __forceinline__ uint32_t meta_pack(uint32_t a, uint32_t nc, uint32_t e, uint32_t pl)
{
    return (a & META_ACTION_MASK) | ((nc & 255u) << 12) | ((e & 3u) << 20) | ((pl & 1u) << 22);
}
__host__ __device__ __forceinline__ int meta_action(uint32_t m) { return (int)(m & META_ACTION_MASK); }
__host__ __device__ __forceinline__ int meta_nc(uint32_t m) { return (int)((m >> 12) & 255u); }
__host__ __device__ __forceinline__ int meta_e(uint32_t m) { return (int)((m >> 20) & 3u); }
__host__ __device__ __forceinline__ int meta_player(uint32_t m) { return (int)((m >> 22) & 1u); }

// ---- packed game state (32 B) ------------------------------------------------
struct __align__(16) GState {
    unsigned long long b0, b1, b2;   // game-specific bitboards
    int turns;                       // GameState._turns (player = turns & 1 for both games)
    int flags;                       // game-specific
};

// game life-cycle bits kept in GState::flags (game flags use the low byte)
constexpr int GF_FINISHED = 0x100;   // terminal, waiting for k_finalize / k_emit
constexpr int GF_DEAD = 0x200;       // finished beyond the games_played quota: no longer stepped

// The words of a game state its observation is a function of (every Game.observation here reads the bitboards and the
// turn counter, never the flags): the key of leaf de-duplication.  Equal keys <=> bit-equal observation rows is what the
// engine needs in one direction only (equal keys => equal rows); states that differ in the turn counter but not in the
// planes derived from it are simply not merged.
template <class S> struct LeafKey;
template <> struct LeafKey<GState> {
    static constexpr int N = 4;
    __host__ __device__ __forceinline__ static void get(const GState &s, unsigned long long (&k)[N])
    {
        k[0] = s.b0; k[1] = s.b1; k[2] = s.b2; k[3] = (unsigned long long)(unsigned)s.turns;
    }
};

struct __align__(16) NodeHot { int n; float q; float p; int child0; };
struct __align__(8) NodeCold { float v; uint32_t meta; };

// Slot header: the game's packed state (G::State: 32 B for Connect4 / brandubh -> a 64-byte header; 64 B for the
// 121-cell hnefatafl boards -> 96 B) followed by 32 B of root fields.  While a node is the root its n / v / child0 /
// meta live here (the pool record of a re-rooted child is copied in by play_moves).
template <class S>
struct __align__(16) SlotHeadT {
    S st;
    int root; int root_n; float root_v; int root_child0;
    uint32_t root_meta; int alloc; int root_rec /* 1: the root has a pool record (it was a child once) */; int pad1;
};
using SlotHead = SlotHeadT<GState>;

// what select leaves for expand/backup (MCTS._curnode / len(_path))
struct __align__(16) LeafInfo { int leaf; int depth; int child0; uint32_t meta; };

// device error bits (sticky word in DevView::err)
enum : uint32_t {
    ERRB_POOL = 1u, ERRB_ACTION = 2u, ERRB_FP = 4u, ERRB_SAMPLES = 8u, ERRB_NOISE = 16u
};

// per-slot statistics (one row per slot, reduced on the host)
struct __align__(16) SlotStats {
    unsigned sum_depth, sum_children, nodes_created, terminal_leaves;   // select: one 16-byte RMW
    unsigned sims, moves; int peak_nodes; int pad;
};

struct Counters {
    long long games_played;
    long long results;
    long long samples_total;   // emitted since reset
    long long sample_count;    // currently in the ring
    long long result_count;    // currently in the ring
};

struct DevView {
    int B, npg, half;   // npg = 2 * half entries per slot: two halves, the live tree occupies a prefix of one (re-root compaction)
    // node pool
    NodeHot *hot; NodeCold *cold;
    // per slot
    void *head;                        // SlotHeadT<G::State>[B]
    LeafInfo *leafinfo; int *path;
    uint32_t *mt; unsigned long long *ctr;
    void *hist_state;                  // G::State[B][hist_cap]
    float *hist_pi; int *hist_len; int hist_cap;
    int *next_reset; int *noise_event; int *last_action; int *fin_code;
    long long *emit_off;
    SlotStats *stats;
    // NN I/O
    float *obs; float *policy; float *value;
    int *nn_rows; int *nn_count;       // slots whose leaf needs the network (non-terminal), filled by select;
    int nn_par;                        // nn_count[2][2]: select adds to [nn_par][model] and clears [nn_par ^ 1][*] for the next one
    int arena_swap;                    // arena: model = env player ^ arena_swap (SelfPlayAgent.player_to_index); rows of model m
                                       // are listed at nn_rows + m * (B / 2)
    // leaf de-duplication (azb_set_leaf_dedup; off: dd_table == nullptr): games whose leaves have the same observation
    // share ONE network evaluation.  select publishes a leaf in dd_table (open addressing, entry = epoch:32 | tag:12 |
    // slot:20, epoch = the select launch, so the table never needs clearing between simulations), the first game with a
    // given state is the representative and the only one listed in nn_rows; dd_src[g] names the row expand/backup reads
    unsigned long long *dd_table; unsigned dd_mask, dd_epoch;
    void *dd_state;                    // G::State[B]: the leaf state of slot g (what its observation is a function of)
    int *dd_src;                       // [B]
    unsigned long long *dd_dups;       // leaves that were served by another game's evaluation
    const float *warm_policy; const float *warm_value;
    // fed root noise
    const float *noise; int noise_events, noise_stride;
    // parameters
    float cpuct, fpu_reduction, noise_frac, root_temp_exp;
    int add_noise, add_temp, rng_mode, symmetric, reset_threshold;
    int arena;       // SelfPlayAgent(_is_arena=True): slots 2i / 2i+1 are the trees of player 0 / 1 of game i
    unsigned long long seed; long long gid_base;
    const float *temp_table; int temp_len;
    long long quota;
    // queues
    float *s_obs; float *s_pi; float *s_z; int *s_slot; long long s_cap;
    int *r_slot; int *r_turns; uint8_t *r_win; long long r_cap;
    Counters *counters;
    uint32_t *err;
};

// Arena mode: the two trees of a game share the game's RNG stream (the reference draws shuffles and the move from
// one np.random stream, SelfPlayAgent.pyx:160 / MCTS.pyx:79); it lives in the even slot of the pair.
__device__ __forceinline__ int rng_slot(const DevView &d, int g) { return d.arena ? (g & ~1) : g; }
__device__ __forceinline__ unsigned long long rng_gid(const DevView &d, int g)
{
    return (unsigned long long)(d.gid_base + (long long)(d.arena ? (g >> 1) : g));
}

// ---- float32 arithmetic with the reference's rounding -------------------------
// The reference's Cython compiles to scalar C without FMA contraction; every
// float op rounds once.  (The library is also built with -fmad=false.)
__device__ __forceinline__ float f_add(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float f_mul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float f_div(float a, float b) { return __fdiv_rn(a, b); }

// float32 power as the framework defines it: correctly rounded via a double pow
// (MCTS.pyx:250 root temperature, :320 probs temperature).
__device__ __forceinline__ float pow_det(float x, float e)
{
    if (e == 1.0f) return x;
    if (x == 0.0f) return 0.0f;
    return (float)pow((double)x, (double)e);
}

// ---- RNG ------------------------------------------------------------------------
// numpy legacy MT19937, regenerated one word at a time (identical sequence to
// the batched twist of numpy/random/src/mt19937/mt19937.c).  st[624] = index.
__device__ inline uint32_t mt_next(uint32_t *st)
{
    uint32_t k = st[624];
    if (k >= 624u) k = 0u;
    uint32_t k1 = (k + 1u == 624u) ? 0u : k + 1u;
    uint32_t km = (k + 397u >= 624u) ? k + 397u - 624u : k + 397u;
    uint32_t y = (st[k] & 0x80000000u) | (st[k1] & 0x7fffffffu);
    y = st[km] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
    st[k] = y;
    st[624] = k + 1u;
    y ^= (y >> 11);
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= (y >> 18);
    return y;
}

__host__ __device__ inline void mt_seed(uint32_t seed, uint32_t *st)
{
    st[0] = seed;
    for (int i = 1; i < 624; i++) st[i] = 1812433253u * (st[i - 1] ^ (st[i - 1] >> 30)) + (uint32_t)i;
    st[624] = 624u;
}

// random_interval: masked rejection on 32-bit draws (RandomState.shuffle of a list)
__device__ inline uint32_t mt_interval(uint32_t *st, uint32_t max)
{
    if (max == 0u) return 0u;
    uint32_t mask = max, v;
    mask |= mask >> 1; mask |= mask >> 2; mask |= mask >> 4; mask |= mask >> 8; mask |= mask >> 16;
    do { v = mt_next(st) & mask; } while (v > max);
    return v;
}

// Philox4x32-10; word w of stream (seed, gid) = lane w&3 of block w>>2 with
// counter (block_lo, block_hi, gid_lo, gid_hi) and key (seed_lo, seed_hi).
__device__ __forceinline__ uint32_t philox_word(unsigned long long seed, unsigned long long gid, unsigned long long w)
{
    uint32_t c0 = (uint32_t)(w >> 2), c1 = (uint32_t)(w >> 34), c2 = (uint32_t)gid, c3 = (uint32_t)(gid >> 32);
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
    for (int r = 0; r < 10; r++) {
        uint32_t h0 = __umulhi(0xD2511F53u, c0), l0 = 0xD2511F53u * c0;
        uint32_t h1 = __umulhi(0xCD9E8D57u, c2), l1 = 0xCD9E8D57u * c2;
        uint32_t n0 = h1 ^ c1 ^ k0, n2 = h0 ^ c3 ^ k1;
        c0 = n0; c1 = l1; c2 = n2; c3 = l0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    uint32_t sel = (uint32_t)w & 3u;
    return sel == 0u ? c0 : sel == 1u ? c1 : sel == 2u ? c2 : c3;
}

// 53-bit uniform of numpy's legacy random_sample from two 32-bit words
__device__ __forceinline__ double u53(uint32_t a, uint32_t b)
{
    return ((double)(a >> 5) * 67108864.0 + (double)(b >> 6)) / 9007199254740992.0;
}

}  // namespace azb
