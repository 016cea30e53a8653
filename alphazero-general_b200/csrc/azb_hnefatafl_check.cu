// Device-code compile check of azb_hnefatafl.cuh (the header is validated on the host, tests/test_hnefatafl_bitboards.py;
// this translation unit only proves that the same code builds for sm_100a).  Not linked into libazb200.so.
#include "azb_hnefatafl.cuh"

using namespace azb;

__global__ void k_hnefatafl_probe(const int *actions, int n, signed char *cells, unsigned char *valid, int *win)
{
    TState128 s;
    Hnefatafl::init(s);
    for (int i = 0; i < n; i++) Hnefatafl::play(s, actions[i]);
    for (int i = threadIdx.x; i < Hnefatafl::CELLS; i += blockDim.x) cells[i] = (signed char)Hnefatafl::cell_code(s, i);
    for (int c = threadIdx.x; c < Hnefatafl::num_candidates(s); c += blockDim.x) {
        int a;
        if (Hnefatafl::candidate(s, c, a)) valid[a] = 1;
    }
    if (threadIdx.x == 0) *win = Hnefatafl::win_code(s) + Hnefatafl::sym_action(3, actions[0]) + (int)Hnefatafl::symmetry(s, 2).b0.lo;
}
