// azb_connect4.cuh -- Connect4 rules on packed bitboards.
//
// Restates alphazero/envs/connect4/Connect4Logic.pyx:40-110 (add_stone,
// get_valid_moves, get_win_state) and alphazero/envs/connect4/connect4.pyx:62-99
// (valid_moves, play_action, win_state, observation, symmetries).
//
// Board: b0 = stones of player 0 (reference cell value +1), b1 = stones of
// player 1 (cell value -1).  Bit index = col * 7 + height, height 0 = bottom
// row (reference row 5), bit col*7+6 is an always-empty guard so that shifted
// run tests never wrap between columns.  player to move = turns & 1.
#pragma once
#include "azb_common.cuh"

namespace azb {

template <int LANES_>
struct Connect4T {
    static constexpr int A = 7;            // Game.action_size()
    static constexpr int H = 6, W = 7;
    static constexpr int OBS_C = 4;
    static constexpr int OBS = OBS_C * H * W;   // Game.observation_size() = (4, 6, 7)
    static constexpr int CELLS = H * W;
    static constexpr int MAXC = 7;         // max children of a node
    static constexpr int MAX_TURNS = 42;   // Game.max_turns()
    static constexpr int MAXD = 44;        // path buffer entries per slot
    static constexpr int NSYM = 2;         // symmetries(): identity, mirror
    using State = GState;
    static constexpr int CTA = 128;        // threads per CTA of the tree kernels
    static constexpr int TYPC = 7;         // typical children per expansion (sizes the default live-tree capacity)
    static constexpr int LANES = LANES_;   // threads cooperating on one game (8, 16 or 32)
    static constexpr bool LANE_IS_ACTION = true;   // A <= LANES: lane a can own the child of action a
    static constexpr unsigned long long TOP = 0x0810204081020ULL;  // bits col*7+5

    __device__ __forceinline__ static void init(GState &s) { s.b0 = s.b1 = s.b2 = 0ULL; s.turns = 0; s.flags = 0; }
    __device__ __forceinline__ static int player(const GState &s) { return s.turns & 1; }

    // Board.add_stone + Game.play_action: lowest empty cell of the column
    __device__ __forceinline__ static void play(GState &s, int col)
    {
        unsigned long long occ = s.b0 | s.b1;
        int h = __popcll((occ >> (col * 7)) & 0x3FULL);
        unsigned long long bit = 1ULL << (col * 7 + h);
        if (s.turns & 1) s.b1 |= bit; else s.b0 |= bit;
        s.turns += 1;
    }

    __device__ __forceinline__ static bool four(unsigned long long b)
    {
        unsigned long long m;
        m = b & (b >> 1); if (m & (m >> 2)) return true;     // vertical
        m = b & (b >> 7); if (m & (m >> 14)) return true;    // horizontal
        m = b & (b >> 6); if (m & (m >> 12)) return true;    // diagonal
        m = b & (b >> 8); if (m & (m >> 16)) return true;    // anti-diagonal
        return false;
    }

    // Game.win_state as a code: 0 none, 1 player 0 won, 2 player 1 won, 3 draw.
    // The reference scans player +1 before -1, then tests for a full top row.
    __device__ __forceinline__ static int win_code(const GState &s)
    {
        if (four(s.b0)) return 1;
        if (four(s.b1)) return 2;
        if (((s.b0 | s.b1) & TOP) == TOP) return 3;
        return 0;
    }

    // Board.get_valid_moves: bit c set when column c's top cell is empty
    __device__ __forceinline__ static uint32_t valid_mask(const GState &s)
    {
        unsigned long long t = ~(s.b0 | s.b1) & TOP;
        uint32_t m = 0;
#pragma unroll
        for (int c = 0; c < W; c++) m |= (uint32_t)((t >> (c * 7 + 5)) & 1ULL) << c;
        return m;
    }

    // Game.valid_moves: number of valid columns; vmask gets one bit per valid action
    __device__ __forceinline__ static int list_valid(const GState &s, short *act, uint32_t &vmask, bool on, int lane)
    {
        vmask = valid_mask(s);
        return __popc(vmask);
    }
    // j-th valid action in ascending order
    __device__ __forceinline__ static int nth_valid(const short *act, uint32_t vmask, int j)
    {
        return (int)__fns(vmask, 0, j + 1);
    }

    // Game.observation: [cells==+1, cells==-1, full(player), full(float32(turns/42))]
    // in the reference's row order (row 0 = top); written by the LANES threads of the group.
    __device__ __forceinline__ static void write_obs(const GState &s, float *out, int lane)
    {
        const float pl = (float)(s.turns & 1);
        const float tn = (float)((double)s.turns / 42.0);
        for (int c = lane; c < CELLS; c += LANES) {
            const int r = c / W, col = c - r * W;
            const int bit = col * 7 + (H - 1 - r);
            out[c] = (float)((s.b0 >> bit) & 1ULL);
            out[CELLS + c] = (float)((s.b1 >> bit) & 1ULL);
            out[2 * CELLS + c] = pl;
            out[3 * CELLS + c] = tn;
        }
    }

    // reference cell code (Board.pieces) of cell i (row-major, row 0 = top)
    __device__ __forceinline__ static int cell_code(const GState &s, int i)
    {
        int r = i / W, c = i - r * W;
        int bit = c * 7 + (H - 1 - r);
        return (int)((s.b0 >> bit) & 1ULL) - (int)((s.b1 >> bit) & 1ULL);
    }

    // inverse of cell_code: Board.pieces (row-major, row 0 = top, +1 / -1 / 0) -> bitboards
    __device__ __forceinline__ static void from_cells(GState &s, const signed char *cells, int turns)
    {
        s.b0 = s.b1 = s.b2 = 0ULL;
        for (int i = 0; i < CELLS; i++) {
            const int r = i / W, c = i - r * W;
            const unsigned long long bit = 1ULL << (c * 7 + (H - 1 - r));
            if (cells[i] > 0) s.b0 |= bit; else if (cells[i] < 0) s.b1 |= bit;
        }
        s.turns = turns; s.flags = 0;
    }

    __device__ __forceinline__ static unsigned long long mirror(unsigned long long b)
    {
        unsigned long long r = 0;
#pragma unroll
        for (int c = 0; c < W; c++) r |= ((b >> (c * 7)) & 0x7FULL) << ((W - 1 - c) * 7);
        return r;
    }

    // Game.symmetries: k = 0 identity, k = 1 column mirror (pi reversed)
    __device__ __forceinline__ static GState symmetry(const GState &s, int k)
    {
        GState o = s;
        if (k == 1) { o.b0 = mirror(s.b0); o.b1 = mirror(s.b1); }
        return o;
    }
    __device__ __forceinline__ static int sym_action(int k, int a) { return k == 1 ? (A - 1 - a) : a; }
};
using Connect4 = Connect4T<8>;

}  // namespace azb
