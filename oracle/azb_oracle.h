/*
 * azb_oracle.h -- CPU restatement of the reference self-play MCTS path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product package
 * (alphazero-general_b200/) links, imports or executes this.  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load liborc.so.
 *
 * It restates, in plain scalar C, the algorithm of (paths relative to the
 * reference tree):
 *   alphazero/MCTS.pyx            Node / MCTS (find_leaf, process_results,
 *                                 update_root, counts, probs)
 *   alphazero/SelfPlayAgent.pyx   generateBatch / processBatch / playMoves
 *   alphazero/envs/connect4/      Connect4 rules, observation, symmetries
 *   fastafl/cengine.pyx + boardgame/board.pyx + envs/brandubh/fastafl.pyx
 *                                 brandubh tafl rules, observation, symmetries
 *   numpy legacy RandomState      MT19937, shuffle, random_sample, choice
 *   numpy float32 pairwise sum
 *
 * Parity status: pinned against the compiled reference itself (oracle/_ref,
 * built by oracle/build_ref.py) in tests/test_oracle_vs_ref.py, against the
 * golden traces in tests/golden/ (generated from oracle/_ref by
 * tests/golden/make_golden.py) and against the Connect4 known-answer boards
 * of alphazero/envs/connect4/test_connect4.py:97-156.
 *
 * One deliberate, documented deviation: every game slot owns its own RNG
 * stream (== a reference SelfPlayAgent with batch size 1 seeded with that
 * slot's seed); the reference interleaves one process-wide stream over the
 * games of a worker, which cannot be reproduced by a parallel engine.
 */
#ifndef AZB_ORACLE_H
#define AZB_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { ORC_GAME_CONNECT4 = 0, ORC_GAME_BRANDUBH = 1, ORC_GAME_HNEFATAFL = 2 };
enum { ORC_RNG_MT19937 = 0, ORC_RNG_PHILOX = 1 };

typedef struct orc_args {
    int32_t game;              /* ORC_GAME_*                                    */
    int32_t num_slots;         /* B: games played in lock-step                   */
    int32_t rng_mode;          /* ORC_RNG_*                                      */
    int32_t add_root_noise;    /* args.add_root_noise                            */
    int32_t add_root_temp;     /* args.add_root_temp                             */
    int32_t symmetric_samples; /* args.symmetricSamples                          */
    int32_t mcts_reset_threshold; /* args.mctsResetThreshold, 0 == None          */
    int32_t temp_table_len;    /* entries in temp_table                          */
    int64_t games_per_iteration; /* args.gamesPerIteration                       */
    int64_t game_id_base;      /* global id of slot 0 (Philox stream key)        */
    uint64_t seed;             /* Philox key / base for MT seeds                 */
    float cpuct;               /* args.cpuct                                     */
    float fpu_reduction;       /* args.fpu_reduction                             */
    float root_noise_frac;     /* args.root_noise_frac                           */
    float root_policy_temp;    /* args.root_policy_temp                          */
    const double *temp_table;  /* temperature used at a move made at turn t      */
                               /* (temp_scaling_fn iterated from startTemp);     */
                               /* t >= len uses the last entry                   */
    const uint32_t *mt_seeds;  /* per-slot np.random.seed() values (MT mode);    */
                               /* NULL: seed + game_id_base + slot               */
    int32_t arena;             /* SelfPlayAgent(_is_arena=True): one tree per     */
                               /* player, no noise / root temperature / samples,  */
                               /* temp_table = [args.arenaTemp]                   */
} orc_args;

typedef struct orc_agent orc_agent;

typedef struct orc_stats {
    int64_t sims;          /* find_leaf/process_results pairs                    */
    int64_t sum_depth;     /* sum over sims of selection levels                  */
    int64_t sum_children;  /* sum over sims of children scanned                  */
    int64_t nodes_created; /* children materialised                              */
    int64_t terminal_leaves;
    int64_t games_played;  /* counted games (<= games_per_iteration)             */
    int64_t results;       /* result_queue entries                               */
    int64_t samples;       /* output_queue entries                               */
    int64_t moves;         /* play_action calls made by play_moves               */
} orc_stats;

orc_agent *orc_create(const orc_args *args);
void orc_destroy(orc_agent *ag);

int orc_action_size(const orc_agent *ag);
int orc_obs_size(const orc_agent *ag);       /* C*H*W floats */

/* SelfPlayAgent.generateBatch: find_leaf for every slot; obs_out[B][obs] */
void orc_generate_batch(orc_agent *ag, float *obs_out);
/* SelfPlayAgent.processBatch: policy[B][A], value[B][3] (not modified) */
void orc_process_batch(orc_agent *ag, const float *policy, const float *value);
/* SelfPlayAgent.playMoves (fast != 0: no history/sample for this move) */
void orc_play_moves(orc_agent *ag, int fast);

/* host-fed Dirichlet noise: noise[slot][event][stride]; the e-th root
 * expansion of a slot mixes noise[slot][e][0..C) by child position. */
void orc_set_root_noise(orc_agent *ag, const float *noise, int events, int stride);

/* MCTS.counts for the current root of every slot: counts[B][A] */
void orc_root_counts(const orc_agent *ag, int32_t *counts);
/* env player to move per slot (arena: whose tree searches) */
void orc_players(const orc_agent *ag, int32_t *players);
/* last action played by play_moves per slot (-1 none) */
void orc_last_actions(const orc_agent *ag, int32_t *actions);
/* per-slot turns of the live game */
void orc_turns(const orc_agent *ag, int32_t *turns);
/* live board of every slot as int8 cell codes, row-major [B][H*W] */
void orc_boards(const orc_agent *ag, int8_t *cells);

void orc_get_stats(const orc_agent *ag, orc_stats *st);

/* output_queue: samples in emission order */
int64_t orc_num_samples(const orc_agent *ag);
void orc_get_samples(const orc_agent *ag, float *obs, float *pi, float *z, int32_t *slot);
void orc_clear_samples(orc_agent *ag);
/* result_queue: (slot, turns, winstate[3]) per finished game */
int64_t orc_num_results(const orc_agent *ag);
void orc_get_results(const orc_agent *ag, int32_t *slot, int32_t *turns, uint8_t *winstate);

/* ---- stand-alone rule / RNG probes used by the unit tests ---------------- */
/* play a move list from the start position; returns 0 ok, -1 illegal drop.
 * cells_out[H*W] int8, valid_out[A] u8, win_out[3] u8 */
int orc_rules_play(int game, const int32_t *actions, int n, int8_t *cells_out,
                   uint8_t *valid_out, uint8_t *win_out, float *obs_out);
/* the same from a constructed tafl position (cell codes, plies played so far) instead of the start position */
int orc_rules_play_from(int game, const int8_t *cells, int turns, const int32_t *actions, int n, int8_t *cells_out,
                        uint8_t *valid_out, uint8_t *win_out, float *obs_out);
/* Game.symmetries entry k of (position, pi): transformed cells and permuted policy */
int orc_rules_symmetry(int game, const int8_t *cells, int turns, const float *pi, int k, int8_t *cells_out, float *pi_out);
/* connect4 win scan on an arbitrary h x w board (cells +1/-1/0), win length k:
 * returns 0 none, 1 player +1, -1 player -1, 2 draw */
int orc_c4_win_state(const int32_t *cells, int h, int w, int k);
float orc_np_sum_f32(const float *a, int n);
float orc_pow_f32(float x, float e);
void orc_mt_seed(uint32_t seed, uint32_t *state625);
uint32_t orc_mt_next(uint32_t *state625);
void orc_philox_words(uint64_t seed, uint64_t game_id, uint64_t first, int n, uint32_t *out);
/* numpy legacy list shuffle of [0..n) with the MT state; perm_out[n] */
void orc_mt_shuffle(uint32_t *state625, int n, int32_t *perm_out);
double orc_mt_double(uint32_t *state625);

#ifdef __cplusplus
}
#endif
#endif
