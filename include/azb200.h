/*
 * azb200.h -- C ABI of libazb200.so, the B200 (sm_100a) batched self-play MCTS
 * engine.
 *
 * The reference (kevaday/alphazero-general) has no FFI for this path: its
 * boundary is the Python object protocol of alphazero/SelfPlayAgent.pyx.  Each
 * entry point below names the reference interface it replaces (file:line
 * relative to the reference tree); INTEGRATION.md shows the ctypes binding a
 * reference maintainer would add.  All pointers are plain C pointers; no
 * torch / Python types cross this boundary.
 *
 * Conventions
 *   - every call returns 0 on success or a negative azb_status; the message of
 *     the last failure on the calling thread is azb_last_error().
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).
 *     Kernels are enqueued asynchronously on it; calls that return data to
 *     host memory synchronise that stream before returning.
 *   - one host thread drives one engine; engines are independent (one per GPU
 *     rank, or several per GPU).
 *   - a "slot" is one of the num_games concurrent games (the reference's
 *     SelfPlayAgent.games[i]).  Slot i owns RNG stream game_id_base + i.
 */
#ifndef AZB200_H
#define AZB200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AZB_ABI_VERSION 1

typedef enum azb_status {
    AZB_OK = 0,
    AZB_ERR_BAD_CONFIG = -1,      /* invalid azb_config field                              */
    AZB_ERR_CUDA = -2,            /* CUDA runtime failure (message has the cudaError)     */
    AZB_ERR_POOL_EXHAUSTED = -3,  /* a slot ran out of node-pool entries                  */
    AZB_ERR_INVALID_ACTION = -4,  /* re-root at an action that is no child:               */
                                  /* ValueError of MCTS.update_root, MCTS.pyx:195         */
    AZB_ERR_FLOATING_POINT = -5,  /* zero/NaN prior or visit sum: FloatingPointError from */
                                  /* np.seterr(all='raise'), MCTS.pyx:23                  */
    AZB_ERR_SAMPLE_OVERFLOW = -6, /* sample / result buffer full, entries were dropped    */
    AZB_ERR_BAD_ARGUMENT = -7,
    AZB_ERR_NOISE_UNDERRUN = -8   /* fed root-noise table exhausted                       */
} azb_status;

enum { AZB_GAME_CONNECT4 = 0, AZB_GAME_BRANDUBH = 1, AZB_GAME_HNEFATAFL = 2 /* 11x11, alphazero/envs/hnefatafl/fastafl.pyx */ };
/* AZB_RNG_MT19937: per-slot numpy-legacy MT19937 streams with numpy's shuffle /
 * choice consumption -> bit-equal to an unmodified reference agent seeded with
 * np.random.seed(mt_seeds[i]).  AZB_RNG_PHILOX: counter-based Philox4x32-10
 * stream per slot (parallel key-sort shuffle); the production mode. */
enum { AZB_RNG_MT19937 = 0, AZB_RNG_PHILOX = 1 };

/* Mirrors the args keys SelfPlayAgent / MCTS read (Coach.py:25-117,
 * MCTS.pyx:133-139, SelfPlayAgent.pyx:58,84-86,149-150,156-158,172-174,181,187). */
typedef struct azb_config {
    int32_t abi_version;          /* AZB_ABI_VERSION                                       */
    int32_t game;                 /* AZB_GAME_*  (game_cls)                                */
    int32_t num_games;            /* process_batch_size: concurrent games on this engine   */
    int32_t device;               /* CUDA device ordinal                                   */
    int32_t rng_mode;             /* AZB_RNG_*                                             */
    int32_t add_root_noise;       /* args.add_root_noise                                   */
    int32_t add_root_temp;        /* args.add_root_temp                                    */
    int32_t symmetric_samples;    /* args.symmetricSamples                                 */
    int32_t mcts_reset_threshold; /* args.mctsResetThreshold (0 = None)                    */
    int32_t max_sims_per_move;    /* largest numMCTSSims/numFastSims/numWarmupSims used;   */
                                  /* sizes the per-slot node pool when max_nodes_per_game=0*/
    int32_t max_nodes_per_game;   /* capacity of one game's LIVE tree in nodes (0 = derive: 8 moves' worth of      */
                                  /* simulations); the slot's arena is twice that -- re-rooting copies the kept      */
                                  /* subtree into the other half, discarded siblings never accumulate                */
    int32_t temp_table_len;       /* entries of temp_table (0 = constant 1.0)              */
    int32_t lanes_per_game;       /* threads cooperating on one game: 0 = default, Connect4 */
                                  /* accepts 8, 16, 32 (tuning knob, results are identical) */
    int32_t arena;                /* 1 = SelfPlayAgent(_is_arena=True), SelfPlayAgent.pyx:23,44-47,62-73,      */
                                  /* 104-132,144-150,167-168: slots 2i / 2i+1 are the trees of env player 0 / 1 */
                                  /* of game i (num_games = 2 x games, even); only the tree of the player to    */
                                  /* move searches (the other slot's obs / policy / value rows are idle), both   */
                                  /* follow every move; no root noise / temperature, no samples; temp_table =    */
                                  /* [args.arenaTemp]; results, games_played and the quota count games, result   */
                                  /* slots are game indices; the pair shares the RNG stream of its even slot     */
    int64_t games_per_iteration;  /* args.gamesPerIteration (quota of counted games)       */
    int64_t sample_capacity;      /* sample ring entries (0 = derive from the quota)       */
    int64_t game_id_base;         /* global id of slot 0 (multi-GPU: rank * num_games)     */
    uint64_t seed;                /* Philox key; MT seeds default to seed+game_id_base+i   */
    float cpuct;                  /* args.cpuct                                            */
    float fpu_reduction;          /* args.fpu_reduction                                    */
    float root_noise_frac;        /* args.root_noise_frac                                  */
    float root_policy_temp;       /* args.root_policy_temp                                 */
    const double *temp_table;     /* host: temperature of the move made at turn t, i.e.    */
                                  /* args.temp_scaling_fn iterated from args.startTemp     */
                                  /* (SelfPlayAgent.pyx:156-158); last entry repeats       */
    const uint32_t *mt_seeds;     /* host, num_games entries or NULL (MT mode)             */
} azb_config;

typedef struct azb_engine azb_engine;

typedef struct azb_stats {
    int64_t sims;             /* find_leaf + process_results pairs completed              */
    int64_t sum_depth;        /* selection levels walked (sum of D over sims)             */
    int64_t sum_children;     /* children scanned by the PUCT selection (sum of C)        */
    int64_t nodes_created;    /* child nodes materialised                                 */
    int64_t terminal_leaves;  /* sims whose leaf was terminal                             */
    int64_t games_played;     /* counted games (<= games_per_iteration)                   */
    int64_t results;          /* finished games reported (result_queue puts)              */
    int64_t samples;          /* samples emitted since creation / reset                   */
    int64_t moves;            /* moves played by azb_play_moves                           */
    int64_t peak_nodes;       /* largest per-slot pool use seen                           */
    int64_t pool_bytes;       /* device bytes of the node pool                            */
    int64_t device_bytes;     /* all device bytes owned by the engine                     */
} azb_stats;

/* ---- lifetime ---------------------------------------------------------- */

/* replaces SelfPlayAgent.__init__ (SelfPlayAgent.pyx:14-60): B fresh games,
 * one empty tree per game, temperatures at startTemp. */
int azb_create(const azb_config *cfg, azb_engine **out);
int azb_destroy(azb_engine *e);
/* start a new iteration: fresh games / trees / RNG streams (seed as in
 * azb_config.seed), counters and queues cleared -- what Coach does by building
 * new agents every iteration (Coach.py:291-323). */
int azb_reset_games(azb_engine *e, uint64_t seed, const uint32_t *mt_seeds);
/* change args.gamesPerIteration of a live engine */
int azb_set_quota(azb_engine *e, int64_t games_per_iteration);

/* ---- geometry ----------------------------------------------------------- */
int azb_action_size(const azb_engine *e);                 /* Game.action_size()       */
int azb_observation_size(const azb_engine *e, int32_t chw[3]); /* Game.observation_size() */
int azb_num_games(const azb_engine *e);

/* ---- NN I/O buffers (device memory owned by the engine) ----------------- */
/* batch_tensor / policy_tensor / value_tensor of SelfPlayAgent.__init__
 * (SelfPlayAgent.pyx:22,27,28; Coach.py:297-309): float32 [B,C,H,W], [B,A],
 * [B,3], row i = slot i.  The caller reads obs and writes policy/value between
 * azb_select and azb_expand_backup. */
float *azb_obs_ptr(azb_engine *e);
float *azb_policy_ptr(azb_engine *e);
float *azb_value_ptr(azb_engine *e);
/* Written by azb_select (and azb_expand_backup_select) for the slot range of that call: the slots whose leaf needs the
 * network -- live games with a non-terminal leaf -- in no particular order (device int32 [B]) and their number
 * (device int32; two counters alternate between consecutive select calls so that none needs a reset launch:
 * azb_nn_count_ptr returns the one the LAST select call wrote, query it after every select).  A terminal leaf's value is its win state and the reference discards the network's answer for it
 * (MCTS.pyx:234-235; SelfPlayAgent.generateBatch still evaluates it, SelfPlayAgent.pyx:116-123), so a caller may
 * evaluate just these rows of azb_obs_ptr (azb_nn_forward_tc_rows does). */
int32_t *azb_nn_rows_ptr(azb_engine *e);
int32_t *azb_nn_count_ptr(azb_engine *e);

/* ---- the hot path -------------------------------------------------------- */
/* SelfPlayAgent.generateBatch (SelfPlayAgent.pyx:103-135) = MCTS.find_leaf
 * (MCTS.pyx:208-228) for slots [first, first+count): PUCT descent with the game
 * replayed on bitboards, first-visit expansion (children in shuffled order),
 * leaf observation written to obs rows. count = 0 means "to the last slot". */
int azb_select(azb_engine *e, int32_t first, int32_t count, void *stream);
/* SelfPlayAgent.processBatch (SelfPlayAgent.pyx:137-151) = MCTS.process_results
 * (MCTS.pyx:230-289): prior masking/renormalisation, root temperature and
 * Dirichlet mix, prior scatter, value backup along the stored path.
 * policy/value: device float32 [B,A] / [B,3] indexed by slot, NULL = the
 * engine-owned buffers. */
int azb_expand_backup(azb_engine *e, int32_t first, int32_t count,
                      const float *policy, const float *value, void *stream);
/* azb_expand_backup of the current simulation followed by azb_select of the next one for the same slots, as one
 * launch (the loop of SelfPlayAgent.run, SelfPlayAgent.pyx:88-92, calls them back to back).  The observation rows
 * are overwritten, so the caller must have consumed them; results are identical to the two separate calls. */
int azb_expand_backup_select(azb_engine *e, int32_t first, int32_t count,
                             const float *policy, const float *value, void *stream);
/* SelfPlayAgent.playMoves (SelfPlayAgent.pyx:153-202): temperature schedule,
 * MCTS.probs, np.random.choice, history append (unless fast), MCTS.update_root,
 * play_action, terminal handling (result, quota, sample emission with
 * symmetries, game/tree reset). Acts on all slots. */
int azb_play_moves(azb_engine *e, int32_t fast, void *stream);
/* ---- single-tree API: MCTS.search / update_root on a caller's position (MCTS.pyx:154-195) -------------------------
 * azb_set_state: slot `slot` restarts from the position given as the reference's cell codes (the layout azb_boards
 * returns: Connect4 Board.pieces +1/-1/0; brandubh Board._state codes), host int8 [H*W], with `turns` moves played,
 * and an empty tree (MCTS.reset).  Synchronous.
 * azb_force_move: MCTS.update_root(gs, action) followed by gs.play_action(action) for one slot -- the root follows a
 * move the caller decided (an unexpanded root first draws the shuffle of its children, as update_root does); an
 * action that is not legal raises AZB_ERR_INVALID_ACTION at the next azb_check_errors (ValueError, MCTS.pyx:195).
 * azb_set_root_flags: the add_root_noise / add_root_temp arguments of MCTS.search for the following simulations. */
int azb_set_state(azb_engine *e, int32_t slot, const int8_t *cells, int32_t turns, void *stream);
int azb_force_move(azb_engine *e, int32_t slot, int32_t action, void *stream);
int azb_set_root_flags(azb_engine *e, int32_t add_root_noise, int32_t add_root_temp);
/* Leaf de-duplication (off by default; no counterpart in the reference, whose NN server evaluates every row of
 * batch_tensor, Coach.py:337-342).  While it is on, games whose leaves have the same observation (same bitboards and
 * turn counter) share ONE network evaluation: azb_select lists only the first such game of its launch in
 * azb_nn_rows_ptr / azb_nn_count_ptr, and azb_expand_backup / azb_expand_backup_select read the other games' policy /
 * value from that game's rows.  Every observation row is still written.  Results are bit-identical to the engine without
 * it as long as the evaluator's answer for a row does not depend on the other rows of the batch (true of the evaluators
 * in azb200_nn.h).  Only with the engine-owned policy / value rows and the compact row list; not in arena mode.
 * Synchronous (allocates 2 x num_games table entries + one state per game on first use).
 * azb_duplicate_leaves: leaves served by another game's evaluation since creation / reset (synchronous). */
int azb_set_leaf_dedup(azb_engine *e, int32_t on);
int azb_duplicate_leaves(azb_engine *e, int64_t *out);
/* arena mode: players[slot] = env player whose tree this slot holds if it searches in the current simulation round
 * (the player to move of a live game), -1 for the idle tree of the pair and for finished games.  Device int32 [B],
 * asynchronous on `stream`.  The caller maps players to models (SelfPlayAgent.player_to_index) and evaluates the
 * active rows of azb_obs_ptr with them (Arena.pyx:262-275). */
int azb_arena_players(azb_engine *e, int32_t *players_device, void *stream);
/* arena mode, device-resident evaluation: SelfPlayAgent.player_to_index is [m, 1 - m] with m = model_of_player0
 * (default 0); azb_select then lists the slots whose leaf needs model k (non-terminal leaves of the games where the
 * player of model k is to move) at azb_arena_rows_ptr(e, k) (device int32 [num_games / 2]) with their number at
 * azb_arena_count_ptr(e, k) (the counter of the LAST select call, as azb_nn_count_ptr) -- the batches
 * SelfPlayAgent.generateBatch builds per player (SelfPlayAgent.pyx:113-131), without leaving the device. */
int azb_arena_set_player_to_index(azb_engine *e, int32_t model_of_player0);
int32_t *azb_arena_rows_ptr(azb_engine *e, int32_t model);
int32_t *azb_arena_count_ptr(azb_engine *e, int32_t model);
/* `sims` x (generateBatch + processBatch) with the warmup constants policy =
 * 1/A, value = 1/3 (SelfPlayAgent.pyx:48-52,111-114) in ONE kernel launch:
 * the NN-free tree-only mode (numWarmupSims). */
int azb_warmup_sims(azb_engine *e, int32_t sims, void *stream);

/* ---- root noise ----------------------------------------------------------- */
/* Parity mode: feed the Dirichlet vectors MCTS._add_root_noise (MCTS.pyx:197-206)
 * would draw. noise: host float32 [num_games][events][stride]; the k-th root
 * expansion of a slot uses row k, entry j for the child at position j.
 * Without a fed table the engine samples Dirichlet(10.83/C) on the device from
 * the slot's Philox stream. */
int azb_set_root_noise(azb_engine *e, const float *noise, int32_t events, int32_t stride);

/* ---- queues --------------------------------------------------------------- */
/* output_queue (SelfPlayAgent.pyx:194-196): copies up to `capacity` samples in
 * emission order to HOST buffers obs [n,C,H,W], pi [n,A], z [n,3], slot [n]
 * (any may be NULL), removes them from the ring, stores n in *count. */
int azb_drain_samples(azb_engine *e, float *obs, float *pi, float *z, int32_t *slot,
                      int64_t capacity, int64_t *count, void *stream);
/* same, but the destination pointers are DEVICE memory (no host copy) */
int azb_drain_samples_device(azb_engine *e, float *obs, float *pi, float *z, int32_t *slot,
                             int64_t capacity, int64_t *count, void *stream);
int azb_sample_count(azb_engine *e, int64_t *count, void *stream);
/* result_queue (SelfPlayAgent.pyx:178): per finished game its slot, final
 * turns and winstate uint8[3]; host buffers. */
int azb_drain_results(azb_engine *e, int32_t *slot, int32_t *turns, uint8_t *winstate,
                      int64_t capacity, int64_t *count, void *stream);
/* games_played.value (SelfPlayAgent.pyx:181-182) */
int azb_games_played(azb_engine *e, int64_t *count, void *stream);

/* ---- introspection (MCTS / Game public attributes used by callers) -------- */
/* MCTS.counts (MCTS.pyx:297-303) of every slot's current root: host int32 [B,A] */
int azb_root_counts(azb_engine *e, int32_t *counts, void *stream);
/* Game.last_action / Game.turns of every slot: host int32 [B] (NULL to skip) */
int azb_game_info(azb_engine *e, int32_t *last_action, int32_t *turns, void *stream);
/* live boards as the reference's cell codes, host int8 [B, H*W] row-major
 * (Connect4: Board.pieces +1/-1/0, Connect4Logic.pyx:23;
 *  brandubh: Board._state codes, fastafl/cengine.pyx:24-32) */
int azb_boards(azb_engine *e, int8_t *cells, void *stream);
/* pre-order dump of one slot's tree, rows of 10 doubles
 * (depth, a, n, q, v, p, player, e0, e1, e2) as Node's public attributes
 * (MCTS.pyx:50-57); *rows_out = rows written (<= max_rows). */
int azb_tree_dump(azb_engine *e, int32_t slot, double *rows, int64_t max_rows,
                  int64_t *rows_out, void *stream);
int azb_stats_get(azb_engine *e, azb_stats *out, void *stream);
/* sticky device error word -> azb_status (0 if clean); clears it */
int azb_check_errors(azb_engine *e, void *stream);

const char *azb_last_error(void);
int azb_abi_version(void);

#ifdef __cplusplus
}
#endif
#endif
