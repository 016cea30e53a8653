/*
 * t128_host.cpp -- HOST build of the engine's 128-bit hnefatafl rules
 * (alphazero-general_b200/csrc/azb_hnefatafl.cuh) behind the same probe interface as the C oracle's
 * orc_rules_play_from, so tests/test_hnefatafl_bitboards.py can compare the two bit for bit on the CPU.
 * TEST INFRASTRUCTURE ONLY -- see azb_oracle.h.  Built by oracle/Makefile into libt128.so with g++ (no CUDA).
 */
#include "../alphazero-general_b200/csrc/azb_hnefatafl.cuh"
#include <string.h>

using namespace azb;
typedef Hnefatafl G;

static void outputs(const TState128 &s, int8_t *cells_out, uint8_t *valid_out, uint8_t *win_out, float *obs_out)
{
    if (cells_out)
        for (int i = 0; i < G::CELLS; i++) cells_out[i] = (int8_t)G::cell_code(s, i);
    if (valid_out) {
        memset(valid_out, 0, G::A);
        int prev = -1;
        const int n = G::num_candidates(s);
        for (int c = 0; c < n; c++) {
            int a;
            if (G::candidate(s, c, a)) {
                if (a <= prev) valid_out[0] = 255;          /* ascending order is part of the contract */
                prev = a;
                valid_out[a] = 1;
            }
        }
    }
    if (win_out) {
        win_out[0] = win_out[1] = win_out[2] = 0;
        const int w = G::win_code(s);
        if (w) win_out[w - 1] = 1;
    }
    if (obs_out)
        for (int p = 0; p < G::OBS_C; p++)
            for (int i = 0; i < G::CELLS; i++) obs_out[p * G::CELLS + i] = G::obs_value(s, p, i);
}

extern "C" int t128_rules_play_from(const int8_t *cells, int turns, const int32_t *actions, int n, int8_t *cells_out,
                                    uint8_t *valid_out, uint8_t *win_out, float *obs_out, int32_t *flags_out)
{
    TState128 s;
    if (cells) G::from_cells(s, (const signed char *)cells, turns);
    else G::init(s);
    for (int i = 0; i < n; i++) G::play(s, actions[i]);
    outputs(s, cells_out, valid_out, win_out, obs_out);
    if (flags_out) *flags_out = s.flags;
    return 0;
}

/* Game.symmetries entry k of (state, pi): transformed cells and the permuted policy */
extern "C" void t128_symmetry(const int8_t *cells, int turns, const float *pi, int k, int8_t *cells_out, float *pi_out)
{
    TState128 s;
    G::from_cells(s, (const signed char *)cells, turns);
    const TState128 o = G::symmetry(s, k);
    for (int i = 0; i < G::CELLS; i++) cells_out[i] = (int8_t)G::cell_code(o, i);
    for (int a = 0; a < G::A; a++) pi_out[a] = 0.0f;
    for (int a = 0; a < G::A; a++) pi_out[G::sym_action(k, a)] = pi[a];
}

extern "C" int t128_codec_roundtrip(void)
{
    for (int a = 0; a < G::A; a++) {
        int x, y, nx, ny;
        G::decode(a, x, y, nx, ny);
        if (x < 0 || x >= G::N || y < 0 || y >= G::N || nx < 0 || nx >= G::N || ny < 0 || ny >= G::N) return a + 1;
        if ((x == nx) == (y == ny)) return a + 1;           /* a rook move changes exactly one coordinate */
        if (G::encode(x, y, nx, ny) != a) return a + 1;
    }
    return 0;
}

/* np.sum(float32[2420]) evaluated the way a warp will: lane b sums leaf b, then five xor-shuffle steps */
extern "C" float t128_np_sum_2420(const float *a, int32_t *leaf_off, int32_t *leaf_len)
{
    float lane[32];
    for (int b = 0; b < 32; b++) {
        int off, len;
        np2420_leaf(b, off, len);
        if (leaf_off) leaf_off[b] = off;
        if (leaf_len) leaf_len[b] = len;
        lane[b] = np_leaf_sum_f32(a + off, len);
    }
    for (int step = 1; step < 32; step <<= 1) {
        float nxt[32];
        for (int b = 0; b < 32; b++) nxt[b] = lane[b] + lane[b ^ step];      /* __shfl_xor_sync(.., step) */
        for (int b = 0; b < 32; b++) lane[b] = nxt[b];
    }
    return lane[0];
}
