// azb_hnefatafl_game.cuh -- hnefatafl (11x11 tafl, alphazero/envs/hnefatafl/fastafl.pyx) as a game of the engine: the
// traits the tree kernels are templated on (azb_connect4.cuh / azb_brandubh.cuh show the contract), on top of the
// 128-bit bitboard rules of azb_hnefatafl.cuh (host-tested bit for bit against the oracle,
// tests/test_hnefatafl_bitboards.py).  One warp per game; the per-group scratch (two 2420-entry action vectors) is
// why a CTA holds two games instead of four.
#pragma once
#include "azb_common.cuh"
#include "azb_hnefatafl.cuh"

namespace azb {

// the rules' state padded to 64 bytes so that the slot header stays a whole number of 16-byte pieces
struct __align__(16) HnefState : TState128 {
    int pad0, pad1;
};
static_assert(sizeof(HnefState) == 64, "hnefatafl slot state");
template <> struct LeafKey<HnefState> {
    static constexpr int N = 7;
    __host__ __device__ __forceinline__ static void get(const HnefState &s, unsigned long long (&k)[N])
    {
        k[0] = s.b0.lo; k[1] = s.b0.hi; k[2] = s.b1.lo; k[3] = s.b1.hi; k[4] = s.b2.lo; k[5] = s.b2.hi;
        k[6] = (unsigned long long)(unsigned)s.turns;
    }
};

struct HnefataflG {
    using State = HnefState;
    using R = Hnefatafl;
    static constexpr int A = R::A;                 // 2420
    static constexpr int H = R::H, W = R::W;
    static constexpr int OBS_C = R::OBS_C;
    static constexpr int OBS = R::OBS;
    static constexpr int CELLS = R::CELLS;
    static constexpr int MAXC = 256;               // legal moves of one side: the node record counts children in 8 bits
    static constexpr int MAX_TURNS = R::MAX_TURNS; // 512
    static constexpr int MAXD = R::MAX_TURNS + 4;
    static constexpr int NSYM = R::NSYM;
    static constexpr int LANES = 32;
    static constexpr int CTA = 64;
    static constexpr int TYPC = 128;               // typical children per expansion (116 at the start position)
    static constexpr bool LANE_IS_ACTION = false;

    __device__ __forceinline__ static void init(State &s) { R::init(s); s.pad0 = s.pad1 = 0; }
    __device__ __forceinline__ static int player(const State &s) { return R::player(s); }
    __device__ __forceinline__ static int cell_code(const State &s, int i) { return R::cell_code(s, i); }
    __device__ __forceinline__ static void from_cells(State &s, const signed char *cells, int turns)
    {
        R::from_cells(s, cells, turns);
        s.pad0 = s.pad1 = 0;
    }
    __device__ __forceinline__ static void play(State &s, int action) { R::play(s, action); }
    __device__ __forceinline__ static int win_code(const State &s) { return R::win_code(s); }

    // Game.valid_moves: legal actions of the side to move, ascending, into act[] (candidates in ascending order are the
    // actions in ascending order; a ballot compacts the legal ones)
    __device__ __forceinline__ static int list_valid(const State &s, short *act, uint32_t &vmask, bool on, int lane)
    {
        vmask = 0u;
        if (!on) return 0;                      // one game per warp: uniform
        const int total = R::num_candidates(s);
        int count = 0;
        for (int base = 0; base < total; base += 32) {
            const int c = base + lane;
            int action = 0;
            const bool ok = c < total && R::candidate(s, c, action);
            const unsigned bal = __ballot_sync(0xffffffffu, ok);
            const int at = count + __popc(bal & ((1u << lane) - 1u));
            if (ok && at < MAXC) act[at] = (short)action;
            count += __popc(bal);
        }
        __syncwarp();
        return count < MAXC ? count : MAXC - 1;            // 255 children fit the node record; never reached in play
    }
    __device__ __forceinline__ static int nth_valid(const short *act, uint32_t vmask, int j) { return (int)act[j]; }

    // _add_obs: [code 2, code 1, king, full(player), full(num_turns / 512 as C int division)]
    __device__ __forceinline__ static void write_obs(const State &s, float *out, int lane)
    {
        for (int i = lane; i < OBS; i += LANES) {
            const int plane = i / CELLS, cell = i - plane * CELLS;
            out[i] = R::obs_value(s, plane, cell);
        }
    }
    __device__ __forceinline__ static State symmetry(const State &s, int k)
    {
        const TState128 t = R::symmetry(s, k);
        State o;
        o.b0 = t.b0; o.b1 = t.b1; o.b2 = t.b2; o.turns = t.turns; o.flags = t.flags; o.pad0 = o.pad1 = 0;
        return o;
    }
    __device__ __forceinline__ static int sym_action(int k, int a) { return R::sym_action(k, a); }
};

}  // namespace azb
