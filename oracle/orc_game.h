/*
 * orc_game.h -- game plug-in table of the CPU oracle (the reference's
 * GameState contract, alphazero/Game.py:7-113, as a C vtable).
 * TEST INFRASTRUCTURE ONLY -- see azb_oracle.h.
 */
#ifndef ORC_GAME_H
#define ORC_GAME_H
#include <stdint.h>

#define ORC_MAX_ACTIONS 2420     /* hnefatafl: 11 * 11 * 20 */
#define ORC_MAX_CHILDREN 512
#define ORC_MAX_PATH 1024
#define ORC_MAX_CELLS 121

typedef struct orc_game {
    int8_t cells[ORC_MAX_CELLS]; /* row-major board; meaning is per game       */
    int32_t player;              /* GameState._player                            */
    int32_t turns;               /* GameState._turns                             */
    int32_t flags;               /* game specific (tafl: king captured / escaped)*/
    int32_t variant;             /* tafl: 0 brandubh, 1 hnefatafl (set by init)  */
} orc_game;

typedef struct orc_game_ops {
    int action_size, obs_size, num_cells, num_symmetries;
    void (*init)(orc_game *g);
    int (*play)(orc_game *g, int action);                 /* 0 ok, -1 illegal   */
    void (*valid_moves)(const orc_game *g, uint8_t *valid);
    void (*win_state)(const orc_game *g, uint8_t win[3]);
    void (*observation)(const orc_game *g, float *obs);
    void (*symmetry)(const orc_game *g, const float *pi, int k, orc_game *g2, float *pi2);
    void (*cells)(const orc_game *g, int8_t *out);
} orc_game_ops;

const orc_game_ops *orc_get_game_ops(int game);
extern const orc_game_ops orc_connect4_ops;
extern const orc_game_ops orc_brandubh_ops;
extern const orc_game_ops orc_hnefatafl_ops;

#endif
