// azb_hnefatafl.cuh -- hnefatafl (11x11 tafl) rules on packed 128-bit bitboards.
//
// NOT YET PART OF libazb200.so: the engine's slot header and move history hold a 3 x u64 state today (azb_common.cuh,
// GState); serving 121-cell boards needs GState templated per game.  This header is the rules half of that work, written
// so that every function is plain scalar code (`AZB_HD` = __host__ __device__): tests/test_hnefatafl_bitboards.py compiles
// it for the HOST (oracle/t128_host.cpp) and checks it bit for bit against the C oracle -- which is pinned to the
// compiled reference -- on random playouts, constructed captures and all 8 symmetries, so the arithmetic that is easy
// to get wrong (shifts across the 64-bit seam, column masks, flood fill) is verified before a kernel ever runs it.
//
// Restates fastafl/cengine.pyx (legal_moves :109-132, _has_legals_check :134-141, king_captured :153-161, get_winner
// :163-169, _check_capture :174-199, _check_surround :201-247, move :249-272) and alphazero/envs/hnefatafl/fastafl.pyx
// (action codec :46-79, observation :82-97, valid_moves, play_action, win_state, symmetries) for variants.hnefatafl_args
// (fastafl/variants.py:1-11,21): king_two_sided_capture = False, move_over_throne, the king may not re-enter the
// throne, DRAW_MOVE_COUNT = 512.
//
// Board: bit index = y * 11 + x (121 bits: lo = bits 0..63, hi = bits 64..120).  b0 = plain pieces of side 1 (code 1,
// the king's side), b1 = side 2 (code 2, moves first = env player 0), b2 = the king (3; 7 on the throne; 8 on a
// corner).  flags bit 0 = Board._king_captured.  Empty throne = 4, empty corner = 5.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define AZB_HD __host__ __device__ __forceinline__
#else
#define AZB_HD inline
#endif

namespace azb {

struct B128 {
    unsigned long long lo, hi;
};

AZB_HD B128 b128(unsigned long long lo, unsigned long long hi) { B128 r; r.lo = lo; r.hi = hi; return r; }
AZB_HD B128 operator&(B128 a, B128 b) { return b128(a.lo & b.lo, a.hi & b.hi); }
AZB_HD B128 operator|(B128 a, B128 b) { return b128(a.lo | b.lo, a.hi | b.hi); }
AZB_HD B128 operator~(B128 a) { return b128(~a.lo, ~a.hi); }
AZB_HD bool operator==(B128 a, B128 b) { return a.lo == b.lo && a.hi == b.hi; }
AZB_HD bool any(B128 a) { return (a.lo | a.hi) != 0ULL; }
AZB_HD B128 bit128(int i) { return i < 64 ? b128(1ULL << i, 0ULL) : b128(0ULL, 1ULL << (i - 64)); }
AZB_HD bool test128(B128 a, int i) { return ((i < 64 ? a.lo >> i : a.hi >> (i - 64)) & 1ULL) != 0ULL; }
// shifts by 0 < n < 64 (the rules use 1 and 11)
AZB_HD B128 shl128(B128 a, int n) { return b128(a.lo << n, (a.hi << n) | (a.lo >> (64 - n))); }
AZB_HD B128 shr128(B128 a, int n) { return b128((a.lo >> n) | (a.hi << (64 - n)), a.hi >> n); }
// bits [0, i)
AZB_HD B128 below128(int i)
{
    if (i <= 0) return b128(0ULL, 0ULL);
    if (i < 64) return b128((1ULL << i) - 1ULL, 0ULL);
    if (i == 64) return b128(~0ULL, 0ULL);
    return b128(~0ULL, i >= 128 ? ~0ULL : (1ULL << (i - 64)) - 1ULL);
}
AZB_HD int popc64(unsigned long long v)
{
#if defined(__CUDA_ARCH__)
    return __popcll(v);
#else
    return __builtin_popcountll(v);
#endif
}
AZB_HD int popc128(B128 a) { return popc64(a.lo) + popc64(a.hi); }
AZB_HD int ctz64(unsigned long long v)
{
#if defined(__CUDA_ARCH__)
    return __ffsll((long long)v) - 1;
#else
    return __builtin_ctzll(v);
#endif
}
AZB_HD B128 lowest128(B128 a) { return a.lo ? b128(a.lo & (~a.lo + 1ULL), 0ULL) : b128(0ULL, a.hi & (~a.hi + 1ULL)); }
// index of the r-th set bit (r = 0 is the lowest); r < popc128(a)
AZB_HD int nth_set128(B128 a, int r)
{
    unsigned long long w = a.lo;
    int base = 0;
    const int nlo = popc64(a.lo);
    if (r >= nlo) { w = a.hi; base = 64; r -= nlo; }
    for (int i = 0; i < r; i++) w &= w - 1ULL;
    return base + ctz64(w);
}

struct TState128 {
    B128 b0, b1, b2;
    int turns, flags;
};

struct Hnefatafl {
    static constexpr int N = 11;
    static constexpr int MT = 2 * N - 2;           // 20 move types per square
    static constexpr int A = N * N * MT;            // 2420
    static constexpr int H = N, W = N;
    static constexpr int OBS_C = 5;
    static constexpr int CELLS = N * N;
    static constexpr int OBS = OBS_C * CELLS;
    static constexpr int MAX_TURNS = 512;           // DRAW_MOVE_COUNT
    static constexpr int NSYM = 8;

    AZB_HD static B128 board() { return b128(~0ULL, (1ULL << 57) - 1ULL); }                    // 121 bits
    AZB_HD static B128 throne() { return bit128(5 * N + 5); }
    AZB_HD static B128 corners() { return bit128(0) | bit128(N - 1) | bit128(N * (N - 1)) | bit128(N * N - 1); }
    AZB_HD static B128 col0()
    {
        B128 m = b128(0ULL, 0ULL);
        for (int y = 0; y < N; y++) m = m | bit128(y * N);
        return m;
    }

    // variants.hnefatafl (fastafl/variants.py:1-11), row y = line y of the string
    AZB_HD static void init(TState128 &s)
    {
        const char *rows[N] = {"50022222005", "00000200000", "00000000000", "20000100002", "20001110002", "22011711022",
                               "20001110002", "20000100002", "00000000000", "00000200000", "50022222005"};
        s.b0 = s.b1 = s.b2 = b128(0ULL, 0ULL);
        for (int y = 0; y < N; y++)
            for (int x = 0; x < N; x++) {
                const char c = rows[y][x];
                if (c == '1') s.b0 = s.b0 | bit128(y * N + x);
                else if (c == '2') s.b1 = s.b1 | bit128(y * N + x);
                else if (c == '3' || c == '7' || c == '8') s.b2 = s.b2 | bit128(y * N + x);
            }
        s.turns = 0;
        s.flags = 0;
    }
    AZB_HD static int player(const TState128 &s) { return s.turns & 1; }

    AZB_HD static int cell_code(const TState128 &s, int i)
    {
        const bool th = i == 5 * N + 5, co = i == 0 || i == N - 1 || i == N * (N - 1) || i == N * N - 1;
        if (test128(s.b2, i)) return th ? 7 : co ? 8 : 3;
        if (test128(s.b0, i)) return 1;
        if (test128(s.b1, i)) return 2;
        return th ? 4 : co ? 5 : 0;
    }
    AZB_HD static void from_cells(TState128 &s, const signed char *cells, int turns)
    {
        s.b0 = s.b1 = s.b2 = b128(0ULL, 0ULL);
        for (int i = 0; i < CELLS; i++) {
            const int v = cells[i];
            if (v == 1) s.b0 = s.b0 | bit128(i);
            else if (v == 2) s.b1 = s.b1 | bit128(i);
            else if (v == 3 || v == 7 || v == 8) s.b2 = s.b2 | bit128(i);
        }
        s.turns = turns;
        s.flags = 0;
    }

    // fastafl.pyx get_move / get_action
    AZB_HD static void decode(int a, int &x, int &y, int &nx, int &ny)
    {
        const int sq = a / MT, mt = a - sq * MT;
        y = sq / N; x = sq - y * N;
        if (mt < N - 1) { nx = x; ny = mt + (mt >= y ? 1 : 0); }
        else { nx = mt - (N - 1); nx += (nx >= x ? 1 : 0); ny = y; }
    }
    AZB_HD static int encode(int x, int y, int nx, int ny)
    {
        int mt;
        if (x == nx) mt = ny < y ? ny : ny - 1;
        else mt = nx < x ? (N - 1) + nx : (N - 2) + nx;
        return MT * (x + N * y) + mt;
    }

    AZB_HD static B128 neighbours(B128 m)
    {
        const B128 not_c0 = ~col0(), not_cl = ~shl128(col0(), N - 1);
        return (shl128(m, N) | shr128(m, N) | shl128(m & not_cl, 1) | shr128(m & not_c0, 1)) & board();
    }
    // squares whose cell code is tile_normal (0)
    AZB_HD static B128 free_squares(const TState128 &s) { return board() & ~(s.b0 | s.b1 | s.b2 | throne() | corners()); }

    // Board._check_capture for the piece that just arrived at (mx, my); king_two_sided_capture = False, so the king is
    // never taken here (do_capture is always False) and only plain pieces of the other side are removed
    AZB_HD static void check_capture(TState128 &s, int mx, int my)
    {
        const int pv = cell_code(s, my * N + mx);
        const bool friend_att = (pv == 1 || pv == 3 || pv == 7 || pv == 8);
        const int enemy = pv != 3 ? 3 - pv : 2;
        const int dx[4] = {0, 1, 0, -1}, dy[4] = {1, 0, -1, 0};
        for (int d = 0; d < 4; d++) {
            const int ex = mx + dx[d], ey = my + dy[d], fx = ex + dx[d], fy = ey + dy[d];
            if (ex < 0 || ex >= N || ey < 0 || ey >= N || fx < 0 || fx >= N || fy < 0 || fy >= N) continue;
            if (cell_code(s, ey * N + ex) != enemy) continue;
            const int w = cell_code(s, fy * N + fx);
            const bool is_friend = friend_att ? (w == 1 || w == 3 || w == 7 || w == 8) : (w == pv);
            if (is_friend || w == 4 || w == 5) {
                const B128 m = ~bit128(ey * N + ex);
                s.b0 = s.b0 & m; s.b1 = s.b1 & m;
            }
        }
    }

    // Board._check_surround: enemy groups touching the moved piece with no tile_normal neighbour are captured (a
    // king in the group only raises the flag and stays on the board)
    AZB_HD static void check_surround(TState128 &s, int mx, int my)
    {
        const B128 moved = bit128(my * N + mx);
        const bool mover_is_2 = any(s.b1 & moved);
        B128 starts = neighbours(moved) & (mover_is_2 ? (s.b0 | s.b2) : s.b1);
        while (any(starts)) {
            const B128 enemy = mover_is_2 ? (s.b0 | s.b2) : s.b1;
            B128 comp = lowest128(starts);
            for (;;) {
                const B128 grown = (comp | neighbours(comp)) & enemy;
                if (grown == comp) break;
                comp = grown;
            }
            starts = starts & ~comp;
            if (!any(neighbours(comp) & free_squares(s))) {
                if (any(comp & s.b2)) s.flags |= 1;
                s.b0 = s.b0 & ~comp;
                s.b1 = s.b1 & ~comp;
            }
        }
    }

    // Game.play_action -> Board.move(_check_valid=False, _check_win=False)
    AZB_HD static void play(TState128 &s, int action)
    {
        int x, y, nx, ny;
        decode(action, x, y, nx, ny);
        const B128 src = bit128(y * N + x), dst = bit128(ny * N + nx);
        if (any(s.b2 & src)) s.b2 = dst;
        else if (any(s.b0 & src)) s.b0 = (s.b0 & ~src) | dst;
        else s.b1 = (s.b1 & ~src) | dst;
        check_capture(s, nx, ny);
        check_surround(s, nx, ny);
        s.turns += 1;
    }

    AZB_HD static bool can_step(B128 pieces, B128 targets) { return any(neighbours(pieces) & targets); }

    // Game.win_state as a code: 0 none, 1 player 0 (side 2) won, 2 player 1 (side 1) won, 3 draw
    AZB_HD static int win_code(const TState128 &s)
    {
        if (s.turns >= MAX_TURNS) return 3;
        const B128 fr = free_squares(s);
        // winner 1: king on an escape square, or side 2 cannot step anywhere (_has_legals_check: one step, no throne hop)
        if (any(s.b2 & corners()) || !can_step(s.b1, fr)) return 2;
        // winner 2: Board.king_captured -- the flag, or every in-bounds neighbour of the king in KING_CAPTURE = (side 2,
        // throne, escape) -- or side 1 (incl. the king, which may step on a corner) cannot step
        const bool captured = (s.flags & 1) || (any(s.b2) && !any(neighbours(s.b2) & ~(s.b1 | throne() | corners())));
        if (captured || !(can_step(s.b0, fr) || can_step(s.b2, fr | corners()))) return 1;
        return 0;
    }

    // candidate c of Game.valid_moves: piece r = c / MT of the side to move (ascending square), move type c % MT.
    // Returns whether the move is legal and its action id (MT * square + move type): candidates in ascending c are the
    // legal actions in ascending order.
    AZB_HD static int num_candidates(const TState128 &s) { return popc128((s.turns & 1) ? (s.b0 | s.b2) : s.b1) * MT; }
    AZB_HD static bool candidate(const TState128 &s, int c, int &action)
    {
        const B128 occ = s.b0 | s.b1 | s.b2;
        const B128 mine = (s.turns & 1) ? (s.b0 | s.b2) : s.b1;
        const int r = c / MT, mt = c - r * MT;
        const int sq = nth_set128(mine, r);
        const int y = sq / N, x = sq - y * N;
        int nx, ny;
        if (mt < N - 1) { nx = x; ny = mt + (mt >= y ? 1 : 0); }
        else { nx = mt - (N - 1); nx += (nx >= x ? 1 : 0); ny = y; }
        const int dsq = ny * N + nx;
        const int lo = sq < dsq ? sq : dsq, hi = sq < dsq ? dsq : sq;
        B128 between = below128(hi) & ~below128(lo + 1);
        if (nx == x) between = between & shl128_or_same(col0(), x);
        const B128 dbit = bit128(dsq);
        const bool king = test128(s.b2, sq);
        const bool dest_ok = !any(occ & dbit) && !any(dbit & throne()) && (!any(dbit & corners()) || king);
        action = MT * sq + mt;
        return dest_ok && !any(between & occ);           // an EMPTY throne may be crossed (move_over_throne), not entered
    }
    AZB_HD static B128 shl128_or_same(B128 a, int n) { return n == 0 ? a : shl128(a, n); }

    // _add_obs plane values: [code 2, code 1, king, full(2 - to_play), full(num_turns / 512 as C int division)]
    AZB_HD static float obs_value(const TState128 &s, int plane, int cell)
    {
        switch (plane) {
        case 0: return test128(s.b1, cell) ? 1.0f : 0.0f;
        case 1: return test128(s.b0, cell) ? 1.0f : 0.0f;
        case 2: return test128(s.b2, cell) ? 1.0f : 0.0f;
        case 3: return (float)(s.turns & 1);
        default: return (float)(s.turns / MAX_TURNS);
        }
    }

    // np.rot90 (counter-clockwise) `rot` times then optional fliplr: output cell (i, j) takes input cell (r, c)
    AZB_HD static B128 transform(B128 b, int rot, int flip)
    {
        B128 out = b128(0ULL, 0ULL);
        for (int i = 0; i < N; i++)
            for (int j = 0; j < N; j++) {
                int r = i, c = flip ? N - 1 - j : j;
                for (int t = 0; t < rot; t++) { const int rr = c, cc = N - 1 - r; r = rr; c = cc; }
                if (test128(b, r * N + c)) out = out | bit128(i * N + j);
            }
        return out;
    }
    // Game.symmetries entry k = (rot - 1) * 2 + flip, rot = 1..4
    AZB_HD static TState128 symmetry(const TState128 &s, int k)
    {
        const int rot = (k >> 1) + 1, flip = k & 1;
        TState128 o = s;
        o.b0 = transform(s.b0, rot, flip);
        o.b1 = transform(s.b1, rot, flip);
        o.b2 = transform(s.b2, rot, flip);
        return o;
    }
    // move coordinates are turned with (x, y) -> (N-1-y, x) per quarter turn (the opposite sense of np.rot90 for odd
    // counts -- as the reference does)
    AZB_HD static int sym_action(int k, int a)
    {
        const int rot = (k >> 1) + 1, flip = k & 1;
        int x, y, nx, ny;
        decode(a, x, y, nx, ny);
        for (int t = 0; t < rot; t++) {
            const int tx = x, tnx = nx;
            x = N - 1 - y; nx = N - 1 - ny; y = tx; ny = tnx;
        }
        if (flip) { x = N - 1 - x; nx = N - 1 - nx; }
        return encode(x, y, nx, ny);
    }
};

// ---- NumPy float32 add.reduce over the 2420-entry policy vector ---------------------------------------------------
// np.sum's pairwise_sum (numpy/core/src/umath/loops_utils.h.src) halves n (rounding the left half down to a multiple
// of 8) until a block has <= 128 elements; for n = 2420 that recursion is a PERFECT binary tree of depth 5 whose 32
// leaves have 72, 80 or 84 elements -- so lane b of a warp sums leaf b (eight strided accumulators, then the serial
// tail, exactly as the leaf loop does) and the tree is five xor-shuffle steps (a + b is commutative in IEEE
// arithmetic, only the tree shape matters).  np2420_leaf gives leaf b's [offset, offset + len).
AZB_HD void np2420_leaf(int b, int &off, int &len)
{
    int o = 0, n = Hnefatafl::A;
    for (int level = 4; level >= 0; level--) {
        int n2 = n / 2;
        n2 -= n2 % 8;
        if ((b >> level) & 1) { o += n2; n -= n2; }
        else n = n2;
    }
    off = o;
    len = n;
}
// the <= 128-element leaf loop of pairwise_sum (n >= 8 here); every add rounds once
AZB_HD float np_leaf_sum_f32(const float *a, int n)
{
    float r[8];
    for (int j = 0; j < 8; j++) r[j] = a[j];
    int i;
    for (i = 8; i < n - (n % 8); i += 8)
        for (int j = 0; j < 8; j++) r[j] = r[j] + a[i + j];
    float res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
    for (; i < n; i++) res = res + a[i];
    return res;
}

}  // namespace azb
