// azb_brandubh.cuh -- brandubh (7x7 tafl) rules on packed bitboards.
//
// Restates fastafl/cengine.pyx (legal_moves :109-132, _has_legals_check :134-141,
// get_winner :146-169, _check_capture :174-199, _check_surround :201-247,
// move :249-272, add/remove_piece :294-330), boardgame/board.pyx
// (has_legal_moves :197-221) and alphazero/envs/brandubh/fastafl.pyx (action
// codec :48-81, observation :84-99, valid_moves :176-183, play_action :185-189,
// win_state :191-203, symmetries :213-256) for the brandubh variant
// (fastafl/variants.py:22: king_two_sided_capture, move_over_throne, the king
// may not re-enter the throne).
//
// Board: bit index = y * 7 + x.  b0 = plain pieces of side 1 (cell code 1, the
// king's side), b1 = pieces of side 2 (code 2, moves first = env player 0),
// b2 = the king (code 3; 7 on the throne; 8 on a corner).  flags bit 0 =
// Board._king_captured.  Empty throne = code 4, empty corner = code 5.
#pragma once
#include "azb_common.cuh"

namespace azb {

struct Brandubh {
    static constexpr int N = 7;
    static constexpr int A = 588;          // 7 * 7 * (7 + 7 - 2)
    static constexpr int H = 7, W = 7;
    static constexpr int OBS_C = 5;
    static constexpr int OBS = OBS_C * H * W;
    static constexpr int CELLS = H * W;
    static constexpr int MAXC = 96;        // 8 pieces x 12 destinations
    static constexpr int MAX_TURNS = 100;  // DRAW_MOVE_COUNT
    static constexpr int MAXD = 104;
    static constexpr int NSYM = 8;
    using State = GState;
    static constexpr int CTA = 128;        // threads per CTA of the tree kernels
    static constexpr int TYPC = 40;        // typical children per expansion (measured mean ~33)
    static constexpr int LANES = 32;
    static constexpr bool LANE_IS_ACTION = false;
    static constexpr unsigned long long BOARD = (1ULL << 49) - 1ULL;
    static constexpr unsigned long long THRONE = 1ULL << 24;
    static constexpr unsigned long long CORNERS = (1ULL << 0) | (1ULL << 6) | (1ULL << 42) | (1ULL << 48);
    static constexpr unsigned long long COL0 = 0x0000040810204081ULL;     // bits 0, 7, ..., 42
    static constexpr unsigned long long NOT_COL0 = BOARD & ~COL0;
    static constexpr unsigned long long NOT_COL6 = BOARD & ~(COL0 << 6);

    // variants.brandubh: rows "5002005" "0002000" "0001000" "2217122" "0001000" "0002000" "5002005"
    __device__ __forceinline__ static void init(GState &s)
    {
        s.b0 = (1ULL << 17) | (1ULL << 23) | (1ULL << 25) | (1ULL << 31);
        s.b1 = (1ULL << 3) | (1ULL << 10) | (1ULL << 21) | (1ULL << 22) | (1ULL << 26) | (1ULL << 27) | (1ULL << 38) | (1ULL << 45);
        s.b2 = THRONE;
        s.turns = 0;
        s.flags = 0;
    }
    __device__ __forceinline__ static int player(const GState &s) { return s.turns & 1; }

    __device__ __forceinline__ static int cell_code(const GState &s, int i)
    {
        const unsigned long long bit = 1ULL << i;
        if (s.b2 & bit) return (bit & THRONE) ? 7 : (bit & CORNERS) ? 8 : 3;
        if (s.b0 & bit) return 1;
        if (s.b1 & bit) return 2;
        if (bit & THRONE) return 4;
        if (bit & CORNERS) return 5;
        return 0;
    }

    // inverse of cell_code: Board._state codes -> bitboards (a captured king ends the game, so a state that can be
    // searched has the flag clear)
    __device__ __forceinline__ static void from_cells(GState &s, const signed char *cells, int turns)
    {
        s.b0 = s.b1 = s.b2 = 0ULL;
        for (int i = 0; i < CELLS; i++) {
            const unsigned long long bit = 1ULL << i;
            const int v = cells[i];
            if (v == 1) s.b0 |= bit; else if (v == 2) s.b1 |= bit; else if (v == 3 || v == 7 || v == 8) s.b2 |= bit;
        }
        s.turns = turns; s.flags = 0;
    }

    // fastafl.pyx get_move
    __device__ __forceinline__ static void decode(int a, int &x, int &y, int &nx, int &ny)
    {
        const int sq = a / 12, mt = a - sq * 12;
        y = sq / 7; x = sq - y * 7;
        if (mt < 6) { nx = x; ny = mt + (mt >= y ? 1 : 0); }
        else { nx = mt - 6; nx += (nx >= x ? 1 : 0); ny = y; }
    }
    // fastafl.pyx get_action
    __device__ __forceinline__ static int encode(int x, int y, int nx, int ny)
    {
        int mt;
        if (x == nx) mt = ny < y ? ny : ny - 1;
        else mt = nx < x ? 6 + nx : 5 + nx;
        return 12 * (x + 7 * y) + mt;
    }

    __device__ __forceinline__ static unsigned long long neighbours(unsigned long long m)
    {
        return ((m << 7) | (m >> 7) | ((m & NOT_COL6) << 1) | ((m & NOT_COL0) >> 1)) & BOARD;
    }
    // squares whose cell code is tile_normal (0)
    __device__ __forceinline__ static unsigned long long free_squares(const GState &s)
    {
        return BOARD & ~(s.b0 | s.b1 | s.b2 | THRONE | CORNERS);
    }

    // Board._check_capture for the piece that just arrived at (mx, my)
    __device__ __forceinline__ static void check_capture(GState &s, int mx, int my)
    {
        const int pv = cell_code(s, my * 7 + mx);
        const bool friend_att = (pv == 1 || pv == 3 || pv == 7 || pv == 8);
        const int enemy = pv != 3 ? 3 - pv : 2;
        const int dx[4] = {0, 1, 0, -1}, dy[4] = {1, 0, -1, 0};
#pragma unroll
        for (int d = 0; d < 4; d++) {
            const int ex = mx + dx[d], ey = my + dy[d], fx = ex + dx[d], fy = ey + dy[d];
            if (ex < 0 || ex > 6 || ey < 0 || ey > 6 || fx < 0 || fx > 6 || fy < 0 || fy > 6) continue;
            const int v = cell_code(s, ey * 7 + ex);
            const bool do_capture = v == 3;
            if (v == enemy || do_capture) {
                const int w = cell_code(s, fy * 7 + fx);
                const bool is_friend = friend_att ? (w == 1 || w == 3 || w == 7 || w == 8) : (w == pv);
                if (is_friend || w == 4 || w == 5) {
                    if (do_capture) s.flags |= 1;
                    else { const unsigned long long m = ~(1ULL << (ey * 7 + ex)); s.b0 &= m; s.b1 &= m; }
                }
            }
        }
    }

    // Board._check_surround: enemy groups touching the moved piece with no
    // tile_normal neighbour are captured (the king only raises the flag)
    __device__ __forceinline__ static void check_surround(GState &s, int mx, int my)
    {
        const unsigned long long moved = 1ULL << (my * 7 + mx);
        const bool mover_is_2 = (s.b1 & moved) != 0ULL;
        unsigned long long starts = neighbours(moved) & (mover_is_2 ? (s.b0 | s.b2) : s.b1);
        while (starts) {
            const unsigned long long enemy = mover_is_2 ? (s.b0 | s.b2) : s.b1;
            unsigned long long comp = starts & (~starts + 1ULL);      // lowest start square
            for (;;) {
                unsigned long long grown = (comp | neighbours(comp)) & enemy;
                if (grown == comp) break;
                comp = grown;
            }
            starts &= ~comp;
            if ((neighbours(comp) & free_squares(s)) == 0ULL) {
                if (comp & s.b2) s.flags |= 1;
                s.b0 &= ~comp;
                s.b1 &= ~comp;
            }
        }
    }

    // Game.play_action -> Board.move(_check_valid=False, _check_win=False)
    __device__ __forceinline__ static void play(GState &s, int action)
    {
        int x, y, nx, ny;
        decode(action, x, y, nx, ny);
        const unsigned long long src = 1ULL << (y * 7 + x), dst = 1ULL << (ny * 7 + nx);
        if (s.b2 & src) s.b2 = dst;
        else if (s.b0 & src) s.b0 = (s.b0 & ~src) | dst;
        else s.b1 = (s.b1 & ~src) | dst;
        check_capture(s, nx, ny);
        check_surround(s, nx, ny);
        s.turns += 1;
    }

    __device__ __forceinline__ static bool can_step(unsigned long long pieces, unsigned long long targets)
    {
        return (neighbours(pieces) & targets) != 0ULL;
    }

    // Game.win_state as a code: 0 none, 1 player 0 (side 2) won, 2 player 1 (side 1) won, 3 draw
    __device__ __forceinline__ static int win_code(const GState &s)
    {
        if (s.turns >= MAX_TURNS) return 3;
        const unsigned long long fr = free_squares(s);
        // winner 1: king on an escape square, or side 2 cannot step anywhere
        if ((s.b2 & CORNERS) || !can_step(s.b1, fr)) return 2;
        // winner 2: king captured, or side 1 (incl. king; the king may step on a corner) cannot step
        if ((s.flags & 1) || !(can_step(s.b0, fr) || can_step(s.b2, fr | CORNERS))) return 1;
        return 0;
    }

    __device__ __forceinline__ static int nth_set_bit(unsigned long long m, int r)
    {
        const unsigned lo = (unsigned)m, hi = (unsigned)(m >> 32);
        const int nlo = __popc(lo);
        return r < nlo ? (int)__fns(lo, 0, r + 1) : 32 + (int)__fns(hi, 0, r - nlo + 1);
    }

    // Game.valid_moves: legal actions of the side to move, ascending, into act[]
    __device__ __forceinline__ static int list_valid(const GState &s, short *act, uint32_t &vmask, bool on, int lane)
    {
        vmask = 0u;
        if (!on) return 0;                      // one game per warp: uniform
        const unsigned long long occ = s.b0 | s.b1 | s.b2;
        const unsigned long long mine = (s.turns & 1) ? (s.b0 | s.b2) : s.b1;
        const int total = __popcll(mine) * 12;
        int count = 0;
        for (int base = 0; base < total; base += 32) {
            const int c = base + lane;
            bool ok = false;
            int action = 0;
            if (c < total) {
                const int r = c / 12, mt = c - r * 12;
                const int sq = nth_set_bit(mine, r);
                const int y = sq / 7, x = sq - y * 7;
                int nx, ny;
                if (mt < 6) { nx = x; ny = mt + (mt >= y ? 1 : 0); }
                else { nx = mt - 6; nx += (nx >= x ? 1 : 0); ny = y; }
                const int dsq = ny * 7 + nx;
                const int lo = sq < dsq ? sq : dsq, hi = sq < dsq ? dsq : sq;
                unsigned long long between = ((1ULL << hi) - 1ULL) & ~((1ULL << (lo + 1)) - 1ULL);
                if (nx == x) between &= (COL0 << x);
                const unsigned long long dbit = 1ULL << dsq;
                const bool king = (s.b2 >> sq) & 1ULL;
                const bool dest_ok = !(occ & dbit) && !(dbit & THRONE) && (!(dbit & CORNERS) || king);
                ok = dest_ok && !(between & occ);
                action = 12 * sq + mt;
            }
            const unsigned bal = __ballot_sync(0xffffffffu, ok);
            if (ok) act[count + __popc(bal & ((1u << lane) - 1u))] = (short)action;
            count += __popc(bal);
        }
        __syncwarp();
        return count;
    }
    __device__ __forceinline__ static int nth_valid(const short *act, uint32_t vmask, int j) { return (int)act[j]; }

    // _add_obs: [code 2, code 1, king, full(player), full(num_turns / 100 as C int division)]
    __device__ __forceinline__ static void write_obs(const GState &s, float *out, int lane)
    {
        const float pl = (float)(s.turns & 1);
        const float tn = (float)(s.turns / MAX_TURNS);
        for (int c = lane; c < CELLS; c += LANES) {
            out[c] = (float)((s.b1 >> c) & 1ULL);
            out[CELLS + c] = (float)((s.b0 >> c) & 1ULL);
            out[2 * CELLS + c] = (float)((s.b2 >> c) & 1ULL);
            out[3 * CELLS + c] = pl;
            out[4 * CELLS + c] = tn;
        }
    }

    // np.rot90 (counter-clockwise) `rot` times then optional fliplr: output cell
    // (i, j) takes input cell src_cell(...)
    __device__ __forceinline__ static unsigned long long transform(unsigned long long b, int rot, int flip)
    {
        unsigned long long out = 0ULL;
        for (int i = 0; i < 7; i++)
            for (int j = 0; j < 7; j++) {
                int r = i, c = flip ? 6 - j : j;           // undo fliplr
                for (int t = 0; t < rot; t++) { const int rr = c, cc = 6 - r; r = rr; c = cc; }   // out[i][j] = in[j][6-i]
                out |= ((b >> (r * 7 + c)) & 1ULL) << (i * 7 + j);
            }
        return out;
    }

    // Game.symmetries entry k = (rot - 1) * 2 + flip, rot = 1..4
    __device__ __forceinline__ static GState symmetry(const GState &s, int k)
    {
        const int rot = (k >> 1) + 1, flip = k & 1;
        GState o = s;
        o.b0 = transform(s.b0, rot, flip);
        o.b1 = transform(s.b1, rot, flip);
        o.b2 = transform(s.b2, rot, flip);
        return o;
    }
    // move coordinates are turned with (x, y) -> (6 - y, x) per quarter turn (the
    // opposite sense of np.rot90 for odd counts -- as the reference does)
    __device__ __forceinline__ static int sym_action(int k, int a)
    {
        const int rot = (k >> 1) + 1, flip = k & 1;
        int x, y, nx, ny;
        decode(a, x, y, nx, ny);
        for (int t = 0; t < rot; t++) {
            const int tx = x, tnx = nx;
            x = 6 - y; nx = 6 - ny; y = tx; ny = tnx;
        }
        if (flip) { x = 6 - x; nx = 6 - nx; }
        return encode(x, y, nx, ny);
    }
};

}  // namespace azb
