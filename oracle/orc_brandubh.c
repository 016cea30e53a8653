/*
 * orc_brandubh.c -- the tafl rules of the reference restated on a cell-code
 * board: brandubh (7x7) and hnefatafl (11x11).  TEST INFRASTRUCTURE ONLY -- see
 * azb_oracle.h.
 *
 * Follows fastafl/cengine.pyx (Board: legal_moves :109-132, _has_legals_check
 * :134-141, get_winner :146-169, _check_capture :174-199, _check_surround
 * :201-247, move :249-272, add_piece/remove_piece :294-330, to_play :332-333),
 * boardgame/board.pyx (has_legal_moves :197-221, _surrounding_squares :269-279)
 * and alphazero/envs/brandubh/fastafl.pyx (action codec :48-81, observation
 * :84-99, valid_moves :176-183, play_action :185-189, win_state :191-203,
 * symmetries :213-256) with the brandubh variant flags of fastafl/variants.py:22
 * (king_two_sided_capture=True, move_over_throne=True, king_can_enter_throne=False).
 * hnefatafl = alphazero/envs/hnefatafl/fastafl.pyx (the same env code with
 * variants.hnefatafl_args, fastafl/variants.py:1-11,21: 11x11 board,
 * king_two_sided_capture=False, DRAW_MOVE_COUNT = 512 :43): the king is never
 * taken by a sandwich; Board.king_captured (cengine.pyx:153-161) instead tests,
 * when the winner is asked for, whether every in-bounds neighbour of the king
 * is in KING_CAPTURE = (side 2, throne, escape) (cengine.pyx:42).
 *
 * Cell codes (cengine.pyx:24-32): 0 empty, 1 king's side ("attacker" in the
 * reference's naming), 2 edge side ("defender", moves first = env player 0),
 * 3 king, 7 king on throne, 8 king on escape, 4 empty throne, 5 empty escape.
 * cells[y * N + x]; Square(x, y).  orc_game.variant selects the variant (set by init).
 */
#include "azb_oracle.h"
#include "orc_game.h"
#include <string.h>

#define MAXN 11
#define FLAG_KING_CAPTURED 1

typedef struct { int n, draw_moves, two_sided; const char *start; } tafl_variant;
static const tafl_variant VAR[2] = {
    { 7, 100, 1, "5002005" "0002000" "0001000" "2217122" "0001000" "0002000" "5002005" },
    { 11, 512, 0, "50022222005" "00000200000" "00000000000" "20000100002" "20001110002" "22011711022"
                  "20001110002" "20000100002" "00000000000" "00000200000" "50022222005" },
};
#define N (VAR[g->variant].n)
#define A_SIZE (N * N * (2 * N - 2))
#define DRAW_MOVE_COUNT (VAR[g->variant].draw_moves)

static const int DX[4] = { 0, 1, 0, -1 };   /* DIRECTIONS (cengine.pyx:46) as (dx, dy) */
static const int DY[4] = { 1, 0, -1, 0 };

#define inb(x, y) ((x) >= 0 && (x) < N && (y) >= 0 && (y) < N)
static int is_king_val(int v) { return v == 3 || v == 7 || v == 8; }
static int in_attackers(int v) { return v == 1 || is_king_val(v); }   /* ATTACKERS */

static void tafl_init_variant(orc_game *g, int variant)
{
    memset(g, 0, sizeof(*g));
    g->variant = variant;
    for (int i = 0; i < N * N; i++) g->cells[i] = (int8_t)(VAR[variant].start[i] - '0');
}
static void tafl_init(orc_game *g) { tafl_init_variant(g, 0); }
static void hnefatafl_init(orc_game *g) { tafl_init_variant(g, 1); }

/* Board.to_play: 2 - num_turns % 2  (side 2 moves first) */
static int to_play(const orc_game *g) { return 2 - (g->turns % 2); }

/* fastafl.pyx get_action */
static int encode_action(const orc_game *g, int x, int y, int nx, int ny)
{
    int mt;
    if (x == nx) mt = ny < y ? ny : ny - 1;
    else { mt = N + nx - 1; if (nx >= x) mt -= 1; }
    return (2 * N - 2) * (x + y * N) + mt;
}

/* fastafl.pyx get_move */
static void decode_action(const orc_game *g, int a, int *x, int *y, int *nx, int *ny)
{
    int size = 2 * N - 2, mt = a % size, sq = a / size;
    *x = sq % N; *y = sq / N;
    if (mt < N - 1) { *nx = *x; *ny = mt; if (mt >= *y) *ny += 1; }
    else { *nx = mt - N + 1; if (*nx >= *x) *nx += 1; *ny = *y; }
}

/* Board._is_valid with king_can_enter_throne == False */
static int sq_valid(const orc_game *g, int x, int y, int is_king)
{
    if (!inb(x, y)) return 0;
    int v = g->cells[y * N + x];
    if (v == 0) return 1;
    if (v == 5) return is_king;
    return 0;
}

/* Game.valid_moves: legal_moves(piece_type=to_play) -> action mask */
static void tafl_valid(const orc_game *g, uint8_t *valid)
{
    memset(valid, 0, A_SIZE);
    int side = to_play(g);
    for (int y = 0; y < N; y++)
        for (int x = 0; x < N; x++) {
            int v = g->cells[y * N + x];
            int mine = side == 1 ? in_attackers(v) : v == 2;
            if (!mine) continue;
            int king = is_king_val(v);
            for (int d = 0; d < 4; d++) {
                int cx = x + DX[d], cy = y + DY[d];
                int throne = inb(cx, cy) && g->cells[cy * N + cx] == 4;   /* move_over_throne */
                while (throne || sq_valid(g, cx, cy, king)) {
                    if (!throne) valid[encode_action(g, x, y, cx, cy)] = 1;
                    cx += DX[d]; cy += DY[d];
                    throne = inb(cx, cy) && g->cells[cy * N + cx] == 4;
                }
            }
        }
}

/* Board._check_capture */
static void check_capture(orc_game *g, int mx, int my)
{
    int pv = g->cells[my * N + mx];
    int friendly_att = in_attackers(pv);
    int enemy = pv != 3 ? 3 - pv : 2;
    for (int d = 0; d < 4; d++) {
        int ex = mx + DX[d], ey = my + DY[d];
        if (!inb(ex, ey)) continue;
        int v = g->cells[ey * N + ex];
        int do_capture = VAR[g->variant].two_sided && v == 3;   /* king_two_sided_capture and value == piece_king */
        if (v == enemy || do_capture) {
            int fx = ex + DX[d], fy = ey + DY[d];
            if (!inb(fx, fy)) continue;
            int w = g->cells[fy * N + fx];
            int is_friend = friendly_att ? in_attackers(w) : (w == pv);
            if (is_friend || w == 4 || w == 5) {
                if (do_capture) g->flags |= FLAG_KING_CAPTURED;
                else g->cells[ey * N + ex] = 0;
            }
        }
    }
}

/* Board._check_surround / __recurse_check: every enemy group touching the moved
 * piece is captured when no member has an in-bounds neighbour equal to
 * tile_normal.  The reference's depth-first search with its early exit decides
 * exactly this AND over the connected group. */
static void check_surround(orc_game *g, int mx, int my)
{
    int pv = g->cells[my * N + mx];
    int enemy_is_att = pv == 2;                          /* _get_team(piece, enemy=True) */
    uint8_t seen[MAXN * MAXN];
    memset(seen, 0, sizeof(seen));
    for (int d = 0; d < 4; d++) {
        int sx = mx + DX[d], sy = my + DY[d];
        if (!inb(sx, sy)) continue;
        int v = g->cells[sy * N + sx];
        if (!(enemy_is_att ? in_attackers(v) : v == 2)) continue;
        if (seen[sy * N + sx]) continue;
        /* flood the group */
        int stack[MAXN * MAXN], sp = 0, group[MAXN * MAXN], gn = 0, free_nb = 0;
        stack[sp++] = sy * N + sx;
        seen[sy * N + sx] = 1;
        while (sp) {
            int c = stack[--sp];
            group[gn++] = c;
            int cx = c % N, cy = c / N;
            for (int e = 0; e < 4; e++) {
                int nx = cx + DX[e], ny = cy + DY[e];
                if (!inb(nx, ny)) continue;
                int w = g->cells[ny * N + nx];
                if (w == 0) free_nb = 1;
                if ((enemy_is_att ? in_attackers(w) : w == 2) && !seen[ny * N + nx]) {
                    seen[ny * N + nx] = 1;
                    stack[sp++] = ny * N + nx;
                }
            }
        }
        if (!free_nb) {
            for (int i = 0; i < gn; i++) {
                int w = g->cells[group[i]];
                if (is_king_val(w)) g->flags |= FLAG_KING_CAPTURED;
                else g->cells[group[i]] = 0;             /* remove_piece of a plain piece */
            }
        }
    }
}

/* Game.play_action -> Board.move(check_turn=False, _check_valid=False, _check_win=False) */
static int tafl_play(orc_game *g, int action)
{
    int x, y, nx, ny;
    decode_action(g, action, &x, &y, &nx, &ny);
    int sv = g->cells[y * N + x];
    int piece = sv, left = 0;
    if (sv == 7) { left = 4; piece = 3; } else if (sv == 8) { left = 5; piece = 3; }   /* remove_piece */
    g->cells[y * N + x] = (int8_t)left;
    int dv = g->cells[ny * N + nx];
    if (dv == 4 || dv == 5) {                            /* add_piece: only a king may land here */
        if (piece != 3) return -1;
        g->cells[ny * N + nx] = (int8_t)(3 + dv);
    } else {
        g->cells[ny * N + nx] = (int8_t)piece;
    }
    check_capture(g, nx, ny);
    check_surround(g, nx, ny);
    g->turns += 1;
    g->player = (g->player + 1) % 2;
    return 0;
}

/* BaseBoard.has_legal_moves(piece_type=side) with Board._has_legals_check: one
 * step only, no throne hopping */
static int has_legal(const orc_game *g, int side)
{
    for (int y = 0; y < N; y++)
        for (int x = 0; x < N; x++) {
            int v = g->cells[y * N + x];
            int mine = side == 1 ? in_attackers(v) : v == 2;
            if (!mine) continue;
            int king = is_king_val(v);
            for (int d = 0; d < 4; d++)
                if (sq_valid(g, x + DX[d], y + DY[d], king)) return 1;
        }
    return 0;
}

/* Board.get_winner */
static int get_winner(const orc_game *g)
{
    int escaped = 0;
    for (int i = 0; i < N * N; i++) escaped |= g->cells[i] == 8;
    if (escaped || !has_legal(g, 2)) return 1;
    int captured = (g->flags & FLAG_KING_CAPTURED) != 0;
    if (!captured && !VAR[g->variant].two_sided) {
        /* Board.king_captured: all(in-bounds neighbours of a king in KING_CAPTURE = (2, 4, 5)) */
        for (int i = 0; i < N * N && !captured; i++) {
            if (!is_king_val(g->cells[i])) continue;
            int all_in = 1;
            for (int d = 0; d < 4; d++) {
                int x = i % N + DX[d], y = i / N + DY[d];
                if (!inb(x, y)) continue;
                int w = g->cells[y * N + x];
                if (!(w == 2 || w == 4 || w == 5)) all_in = 0;
            }
            captured = all_in;
        }
    }
    if (captured || !has_legal(g, 1)) return 2;
    return 0;
}

/* Game.win_state: draw at 100 turns is tested first; result[2 - winner] */
static void tafl_win(const orc_game *g, uint8_t win[3])
{
    win[0] = win[1] = win[2] = 0;
    if (g->turns >= DRAW_MOVE_COUNT) { win[2] = 1; return; }
    int w = get_winner(g);
    if (w) win[2 - w] = 1;
}

/* _add_obs: [state==2, state==1, king mask, full(2 - to_play), full(num_turns / 100)]
 * -- the last quotient is a C integer division in the compiled reference */
static void tafl_obs(const orc_game *g, float *obs)
{
    float colour = (float)(2 - to_play(g));
    float turn = (float)(g->turns / DRAW_MOVE_COUNT);
    for (int i = 0; i < N * N; i++) {
        int v = g->cells[i];
        obs[i] = v == 2 ? 1.0f : 0.0f;
        obs[N * N + i] = v == 1 ? 1.0f : 0.0f;
        obs[2 * N * N + i] = is_king_val(v) ? 1.0f : 0.0f;
        obs[3 * N * N + i] = colour;
        obs[4 * N * N + i] = turn;
    }
}

/* Game.symmetries, entry k = (i - 1) * 2 + flip for i = 1..4 quarter turns:
 * the board array is np.rot90'ed i times (counter-clockwise) and optionally
 * fliplr'ed, while every move's coordinates are turned with
 * (x, y) -> (W-1-y, x) i times -- the opposite sense for odd i; kept as is. */
static void tafl_sym(const orc_game *g, const float *pi, int k, orc_game *g2, float *pi2)
{
    int rot = k / 2 + 1, flip = k & 1;
    *g2 = *g;
    int8_t cur[MAXN * MAXN], nxt[MAXN * MAXN];
    memcpy(cur, g->cells, N * N);
    for (int r = 0; r < rot; r++) {
        /* np.rot90: out[i][j] = in[j][N-1-i] */
        for (int i = 0; i < N; i++)
            for (int j = 0; j < N; j++) nxt[i * N + j] = cur[j * N + (N - 1 - i)];
        memcpy(cur, nxt, N * N);
    }
    if (flip) {
        for (int i = 0; i < N; i++)
            for (int j = 0; j < N; j++) nxt[i * N + j] = cur[i * N + (N - 1 - j)];
        memcpy(cur, nxt, N * N);
    }
    memcpy(g2->cells, cur, N * N);
    const int asz = A_SIZE;
    for (int a = 0; a < asz; a++) pi2[a] = 0.0f;
    for (int a = 0; a < asz; a++) {
        int x, y, nx, ny;
        decode_action(g, a, &x, &y, &nx, &ny);
        for (int r = 0; r < rot; r++) {
            int tx = x, tnx = nx;
            x = N - 1 - y; nx = N - 1 - ny;
            y = tx; ny = tnx;
        }
        if (flip) { x = N - 1 - x; nx = N - 1 - nx; }
        pi2[encode_action(g, x, y, nx, ny)] = pi[a];
    }
}

static void tafl_cells(const orc_game *g, int8_t *out) { memcpy(out, g->cells, N * N); }

const orc_game_ops orc_brandubh_ops = {
    588, 5 * 49, 49, 8, tafl_init, tafl_play, tafl_valid, tafl_win, tafl_obs, tafl_sym, tafl_cells
};
const orc_game_ops orc_hnefatafl_ops = {
    2420, 5 * 121, 121, 8, hnefatafl_init, tafl_play, tafl_valid, tafl_win, tafl_obs, tafl_sym, tafl_cells
};
