/*
 * orc_connect4.c -- Connect4 rules of the reference restated on an integer
 * cell board (alphazero/envs/connect4/Connect4Logic.pyx:40-110 and
 * alphazero/envs/connect4/connect4.pyx:34-99).  TEST INFRASTRUCTURE ONLY.
 */
#include "azb_oracle.h"
#include "orc_game.h"
#include <string.h>

#define H 6
#define W 7
#define K 4
#define MAX_TURNS 42

static void c4_init(orc_game *g)
{
    memset(g, 0, sizeof(*g));
}

/* Board.add_stone (Connect4Logic.pyx:40-48) + Game.play_action
 * (connect4.pyx:65-68): stone +1 for player 0, -1 for player 1, lowest empty
 * row of the column (row H-1 is the bottom). */
static int c4_play(orc_game *g, int col)
{
    int stone = g->player == 0 ? 1 : -1;
    for (int r = H - 1; r >= 0; r--) {
        if (g->cells[r * W + col] == 0) {
            g->cells[r * W + col] = (int8_t)stone;
            g->player = (g->player + 1) % 2;
            g->turns += 1;
            return 0;
        }
    }
    return -1;                                   /* reference raises ValueError */
}

/* Board.get_valid_moves (Connect4Logic.pyx:50-58) */
static void c4_valid(const orc_game *g, uint8_t *valid)
{
    for (int c = 0; c < W; c++) valid[c] = g->cells[c] == 0;
}

/* Board.get_win_state (Connect4Logic.pyx:60-110) on a general h x w board:
 * player +1 is scanned before -1; rows, columns, then the two diagonals; a
 * full top row with no winner is a draw. */
int orc_c4_win_state(const int32_t *cells, int h, int w, int k)
{
    static const int players[2] = { 1, -1 };
    for (int pi = 0; pi < 2; pi++) {
        int p = players[pi];
        for (int r = 0; r < h; r++) {
            int run = 0;
            for (int c = 0; c < w; c++) {
                run = cells[r * w + c] == p ? run + 1 : 0;
                if (run == k) return p;
            }
        }
        for (int c = 0; c < w; c++) {
            int run = 0;
            for (int r = 0; r < h; r++) {
                run = cells[r * w + c] == p ? run + 1 : 0;
                if (run == k) return p;
            }
        }
        for (int r = 0; r + k <= h; r++) {
            for (int c = 0; c + k <= w; c++) {
                int good = 1;
                for (int x = 0; x < k && good; x++) good = cells[(r + x) * w + c + x] == p;
                if (good) return p;
            }
            for (int c = k - 1; c < w; c++) {
                int good = 1;
                for (int x = 0; x < k && good; x++) good = cells[(r + x) * w + c - x] == p;
                if (good) return p;
            }
        }
    }
    for (int c = 0; c < w; c++)
        if (cells[c] == 0) return 0;
    return 2;
}

/* Game.win_state (connect4.pyx:70-82): [p0 won, p1 won, draw] */
static void c4_win(const orc_game *g, uint8_t win[3])
{
    int32_t cells[H * W];
    for (int i = 0; i < H * W; i++) cells[i] = g->cells[i];
    int r = orc_c4_win_state(cells, H, W, K);
    win[0] = r == 1; win[1] = r == -1; win[2] = r == 2;
}

/* Game.observation (connect4.pyx:84-91): planes [cells==1, cells==-1,
 * full(player), full(float32(turns / 42))] */
static void c4_obs(const orc_game *g, float *obs)
{
    float turn = (float)((double)g->turns / (double)MAX_TURNS);
    for (int i = 0; i < H * W; i++) {
        obs[i] = g->cells[i] == 1 ? 1.0f : 0.0f;
        obs[H * W + i] = g->cells[i] == -1 ? 1.0f : 0.0f;
        obs[2 * H * W + i] = (float)g->player;
        obs[3 * H * W + i] = turn;
    }
}

/* Game.symmetries (connect4.pyx:96-99): identity, then the column mirror with
 * pi reversed */
static void c4_sym(const orc_game *g, const float *pi, int k, orc_game *g2, float *pi2)
{
    *g2 = *g;
    if (k == 0) {
        memcpy(pi2, pi, sizeof(float) * W);
        return;
    }
    for (int r = 0; r < H; r++)
        for (int c = 0; c < W; c++) g2->cells[r * W + c] = g->cells[r * W + (W - 1 - c)];
    for (int c = 0; c < W; c++) pi2[c] = pi[W - 1 - c];
}

static void c4_cells(const orc_game *g, int8_t *out) { memcpy(out, g->cells, H * W); }

const orc_game_ops orc_connect4_ops = {
    W, 4 * H * W, H * W, 2, c4_init, c4_play, c4_valid, c4_win, c4_obs, c4_sym, c4_cells
};

const orc_game_ops *orc_get_game_ops(int game)
{
    if (game == ORC_GAME_CONNECT4) return &orc_connect4_ops;
#ifdef ORC_HAVE_BRANDUBH
    if (game == ORC_GAME_BRANDUBH) return &orc_brandubh_ops;
    if (game == ORC_GAME_HNEFATAFL) return &orc_hnefatafl_ops;
#endif
    return 0;
}
