// azb_engine.cu -- host side of libazb200.so: the C ABI of include/azb200.h.
// Owns the device memory of one engine, launches the kernels of
// azb_kernels.cuh on the caller's stream and moves queue contents to the host.
#include "../../include/azb200.h"
#include "azb_common.cuh"
#include "azb_connect4.cuh"
#ifdef AZB_HAVE_BRANDUBH
#include "azb_brandubh.cuh"
#include "azb_hnefatafl_game.cuh"
#endif
#include "azb_kernels.cuh"

#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

using namespace azb;

static thread_local std::string g_err;

static int fail(int code, const char *fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}

#define CK(call)                                                                              \
    do {                                                                                      \
        cudaError_t _e = (call);                                                              \
        if (_e != cudaSuccess)                                                                \
            return fail(AZB_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(_e), \
                        __FILE__, __LINE__);                                                  \
    } while (0)

struct GameDims {
    int A, obs_c, obs_h, obs_w, cells, maxc, max_turns, maxd, nsym, lanes;
    int state_bytes, head_bytes, turns_off;      // layout of SlotHeadT<G::State> for the host-side readers
    int typc;                                    // typical children per expansion
};

template <class G>
static GameDims dims_of()
{
    using S = typename G::State;
    S probe;
    const int turns_off = (int)(reinterpret_cast<const char *>(&probe.turns) - reinterpret_cast<const char *>(&probe));
    return GameDims{G::A, G::OBS_C, G::H, G::W, G::CELLS, G::MAXC, G::MAX_TURNS, G::MAXD, G::NSYM, G::LANES,
                    (int)sizeof(S), (int)sizeof(SlotHeadT<S>), turns_off, G::TYPC};
}

// the game-independent part of a slot header, as the host reads it
struct HeadView {
    int turns, flags, root, root_n; float root_v; int root_child0; uint32_t root_meta; int alloc, root_rec;
};
static HeadView head_view(const GameDims &gd, const unsigned char *raw)
{
    HeadView h;
    memcpy(&h.turns, raw + gd.turns_off, 4); memcpy(&h.flags, raw + gd.turns_off + 4, 4);
    const unsigned char *r = raw + gd.state_bytes;
    memcpy(&h.root, r, 4); memcpy(&h.root_n, r + 4, 4); memcpy(&h.root_v, r + 8, 4); memcpy(&h.root_child0, r + 12, 4);
    memcpy(&h.root_meta, r + 16, 4); memcpy(&h.alloc, r + 20, 4); memcpy(&h.root_rec, r + 24, 4);
    return h;
}

struct azb_engine {
    azb_config cfg;
    GameDims gd;
    DevView d;
    std::vector<void *> allocs;
    std::vector<float> temp_host;
    float *noise_dev = nullptr;
    long long device_bytes = 0, pool_bytes = 0;
    int lanes = 8;                  // threads per game (Connect4: 8 / 16 / 32)
    int *scratch_i32 = nullptr;     // B * A ints for introspection
    int8_t *scratch_i8 = nullptr;
    // leaf de-duplication (azb_set_leaf_dedup): buffers allocated on first use, d.dd_table != nullptr while it is on
    unsigned long long *dd_table = nullptr;
    unsigned dd_slots = 0, dd_epoch = 0;
};

// leaf de-duplication: the next select launch gets a fresh epoch (entries of older epochs read as free)
static inline void dd_next_epoch(azb_engine *e) { if (e->d.dd_table) e->d.dd_epoch = ++e->dd_epoch; }
// ... and the table is cleared whenever the epochs start over (every move-round: a captured round graph replays the
// same epochs, so the clear is part of azb_play_moves)
static inline int dd_restart(azb_engine *e, cudaStream_t s)
{
    if (!e->dd_table) return 0;
    e->dd_epoch = 0;
    return cudaMemsetAsync(e->dd_table, 0, (size_t)e->dd_slots * sizeof(unsigned long long), s) == cudaSuccess ? 0 : 1;
}

template <class T>
static int dev_alloc(azb_engine *e, T **p, size_t count, bool zero = true)
{
    void *q = nullptr;
    size_t bytes = count * sizeof(T);
    if (bytes == 0) bytes = sizeof(T);
    CK(cudaMalloc(&q, bytes));
    if (zero) CK(cudaMemset(q, 0, bytes));
    e->allocs.push_back(q);
    e->device_bytes += (long long)bytes;
    *p = (T *)q;
    return 0;
}

#define TRY(x) do { int _r = (x); if (_r != 0) return _r; } while (0)


template <class G>
static inline int grid_for(int games) { return (games * G::LANES + G::CTA - 1) / G::CTA; }

// launchers (one per kernel) so that the game dispatch is a plain if/else
template <class G> static void l_init(azb_engine *e, const uint32_t *seeds_dev, cudaStream_t s)
{
    k_init_slots<G><<<(e->d.B + 127) / 128, 128, 0, s>>>(e->d, seeds_dev);
}
template <class G> static void l_select(azb_engine *e, int first, int count, cudaStream_t s)
{
    k_select<G><<<grid_for<G>(count), G::CTA, 0, s>>>(e->d, first, count);
}
template <class G> static void l_expand(azb_engine *e, int first, int count, const float *pol, const float *val, cudaStream_t s)
{
    k_expand_backup<G><<<grid_for<G>(count), G::CTA, 0, s>>>(e->d, first, count, pol, val);
}
template <class G> static void l_expand_select(azb_engine *e, int first, int count, const float *pol, const float *val, cudaStream_t s)
{
    k_expand_select<G><<<grid_for<G>(count), G::CTA, 0, s>>>(e->d, first, count, pol, val);
}
template <class G> static void l_play(azb_engine *e, int fast, cudaStream_t s)
{
    if (e->d.arena) k_play_moves_arena<G><<<grid_for<G>(e->d.B / 2), G::CTA, 0, s>>>(e->d);
    else k_play_moves<G><<<grid_for<G>(e->d.B), G::CTA, 0, s>>>(e->d, fast);
    k_finalize<G><<<1, 1024, 0, s>>>(e->d);
    k_emit<G><<<dim3(grid_for<G>(e->d.B), G::A >= 256 ? 16 : 2), G::CTA, 0, s>>>(e->d);
    k_emit_reset<G><<<(e->d.B + 127) / 128, 128, 0, s>>>(e->d);
}
template <class G> static void l_arena_players(azb_engine *e, int *out, cudaStream_t s)
{
    k_arena_players<G><<<(e->d.B + 127) / 128, 128, 0, s>>>(e->d, out);
}
template <class G> static void l_set_state(azb_engine *e, int slot, const signed char *cells, int turns, cudaStream_t s)
{
    k_set_state<G><<<1, 32, 0, s>>>(e->d, slot, cells, turns);
}
template <class G> static void l_force_move(azb_engine *e, int slot, int action, cudaStream_t s)
{
    k_force_move<G><<<1, G::CTA, 0, s>>>(e->d, slot, action);
}
template <class G> static void l_warmup(azb_engine *e, int sims, cudaStream_t s)
{
    k_warmup_sims<G><<<grid_for<G>(e->d.B), G::CTA, 0, s>>>(e->d, sims);
}
template <class G> static void l_counts(azb_engine *e, cudaStream_t s)
{
    k_root_counts<G><<<(e->d.B + 127) / 128, 128, 0, s>>>(e->d, e->scratch_i32);
}
template <class G> static void l_boards(azb_engine *e, cudaStream_t s)
{
    k_boards<G><<<(e->d.B + 127) / 128, 128, 0, s>>>(e->d, e->scratch_i8);
}

#define DISPATCH_C4(e, fn, ...) do { \
        if ((e)->lanes == 32) fn<Connect4T<32>>(__VA_ARGS__); \
        else if ((e)->lanes == 16) fn<Connect4T<16>>(__VA_ARGS__); \
        else fn<Connect4T<8>>(__VA_ARGS__); } while (0)
#ifdef AZB_HAVE_BRANDUBH
#define DISPATCH(e, fn, ...) do { if ((e)->cfg.game == AZB_GAME_CONNECT4) DISPATCH_C4(e, fn, __VA_ARGS__); \
        else if ((e)->cfg.game == AZB_GAME_HNEFATAFL) fn<HnefataflG>(__VA_ARGS__); else fn<Brandubh>(__VA_ARGS__); } while (0)
#else
#define DISPATCH(e, fn, ...) DISPATCH_C4(e, fn, __VA_ARGS__)
#endif

static int init_slots(azb_engine *e, const uint32_t *mt_seeds_host)
{
    uint32_t *seeds_dev = nullptr;
    if (mt_seeds_host && e->d.mt) {
        CK(cudaMalloc(&seeds_dev, sizeof(uint32_t) * (size_t)e->d.B));
        CK(cudaMemcpy(seeds_dev, mt_seeds_host, sizeof(uint32_t) * (size_t)e->d.B, cudaMemcpyHostToDevice));
    }
    CK(cudaMemset(e->d.counters, 0, sizeof(Counters)));
    CK(cudaMemset(e->d.err, 0, sizeof(uint32_t)));
    if (e->dd_table) {
        CK(cudaMemset(e->dd_table, 0, (size_t)e->dd_slots * sizeof(unsigned long long)));
        CK(cudaMemset(e->d.dd_dups, 0, sizeof(unsigned long long)));
        e->dd_epoch = 0;
    }
    DISPATCH(e, l_init, e, seeds_dev, (cudaStream_t)0);
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    if (seeds_dev) CK(cudaFree(seeds_dev));
    return 0;
}

extern "C" int azb_abi_version(void) { return AZB_ABI_VERSION; }
extern "C" const char *azb_last_error(void) { return g_err.c_str(); }

extern "C" int azb_create(const azb_config *cfg, azb_engine **out)
{
    if (!cfg || !out) return fail(AZB_ERR_BAD_ARGUMENT, "azb_create: null argument");
    *out = nullptr;
    if (cfg->abi_version != AZB_ABI_VERSION)
        return fail(AZB_ERR_BAD_CONFIG, "abi_version %d, library is %d", cfg->abi_version, AZB_ABI_VERSION);
    if (cfg->num_games <= 0) return fail(AZB_ERR_BAD_CONFIG, "num_games must be positive");
    if (cfg->arena && (cfg->num_games & 1)) return fail(AZB_ERR_BAD_CONFIG, "arena mode: num_games counts slots, two per game");
    if (cfg->arena && (cfg->add_root_noise || cfg->add_root_temp))
        return fail(AZB_ERR_BAD_CONFIG, "arena mode runs without root noise / temperature (SelfPlayAgent.pyx:147-150)");
    if (cfg->rng_mode != AZB_RNG_MT19937 && cfg->rng_mode != AZB_RNG_PHILOX)
        return fail(AZB_ERR_BAD_CONFIG, "unknown rng_mode %d", cfg->rng_mode);
    GameDims gd;
    if (cfg->game == AZB_GAME_CONNECT4) gd = dims_of<Connect4>();
#ifdef AZB_HAVE_BRANDUBH
    else if (cfg->game == AZB_GAME_BRANDUBH) gd = dims_of<Brandubh>();
    else if (cfg->game == AZB_GAME_HNEFATAFL) gd = dims_of<HnefataflG>();
#endif
    else return fail(AZB_ERR_BAD_CONFIG, "unknown game %d", cfg->game);
    if (cfg->temp_table_len < 0 || (cfg->temp_table_len > 0 && !cfg->temp_table))
        return fail(AZB_ERR_BAD_CONFIG, "temp_table missing");
    if (!(cfg->root_policy_temp > 0.0f)) return fail(AZB_ERR_BAD_CONFIG, "root_policy_temp must be > 0");
    int ndev = 0;
    CK(cudaGetDeviceCount(&ndev));
    if (cfg->device < 0 || cfg->device >= ndev) return fail(AZB_ERR_BAD_CONFIG, "device %d of %d", cfg->device, ndev);
    CK(cudaSetDevice(cfg->device));

    azb_engine *e = new azb_engine();
    e->cfg = *cfg;
    e->cfg.temp_table = nullptr;
    e->cfg.mt_seeds = nullptr;
    e->gd = gd;
    if (cfg->game == AZB_GAME_CONNECT4) {
        const char *env = getenv("AZB_C4_LANES");
        int l = cfg->lanes_per_game > 0 ? cfg->lanes_per_game : (env ? atoi(env) : 8);
        if (l != 8 && l != 16 && l != 32) { delete e; return fail(AZB_ERR_BAD_CONFIG, "lanes_per_game must be 8, 16 or 32"); }
        e->lanes = l;
    } else e->lanes = gd.lanes;
    DevView &d = e->d;
    memset(&d, 0, sizeof(d));
    const int B = cfg->num_games;
    d.B = B;
    int sims = cfg->max_sims_per_move > 0 ? cfg->max_sims_per_move : 100;
    // Capacity of the LIVE tree of one game (entries of one half of the slot's arena; re-rooting compacts the kept
    // subtree into the other half, azb_kernels.cuh compact_subtree).  The live tree holds at most one block of children
    // per simulation that went through the current root, i.e. <= root.n * maxc entries; root.n grows by `sims` per move
    // and shrinks to the played child's share at every re-root.  Default: room for 8 moves' worth of simulations at the
    // game's typical fan-out (G::TYPC; brandubh ~35 of at most 96), never more than the whole-game bound.  Exhaustion is reported
    // (AZB_ERR_POOL_EXHAUSTED), never overrun.
    const long long whole_game = 1 + (long long)gd.max_turns * sims * gd.maxc;
    long long live = cfg->max_nodes_per_game;
    if (live <= 0) {
        live = 1 + 8LL * sims * gd.typc + gd.maxc;
        if (live > whole_game) live = whole_game;
    }
    if (live < 2 + gd.maxc || 2 * live > 0x7fffffffLL) { delete e; return fail(AZB_ERR_BAD_CONFIG, "max_nodes_per_game %lld", live); }
    const long long npg = 2 * live;
    d.npg = (int)npg;
    d.half = (int)live;
    const size_t N = (size_t)B * (size_t)npg;
    int rc = 0;
#define A_(x) do { if (!rc) rc = (x); } while (0)
    // +128 entries: the speculative sibling-block prefetch may read past a slot's last block
    A_(dev_alloc(e, &d.hot, N + 128, false)); A_(dev_alloc(e, &d.cold, N + 128, false));
    e->pool_bytes = e->device_bytes;
    { unsigned char *hp = nullptr; A_(dev_alloc(e, &hp, (size_t)B * gd.head_bytes)); d.head = hp; }
    A_(dev_alloc(e, &d.leafinfo, B));
    A_(dev_alloc(e, &d.path, (size_t)B * gd.maxd));
    if (cfg->rng_mode == AZB_RNG_MT19937) A_(dev_alloc(e, &d.mt, (size_t)B * 625));
    A_(dev_alloc(e, &d.ctr, B));
    d.hist_cap = gd.max_turns;
    { unsigned char *hs = nullptr; A_(dev_alloc(e, &hs, (size_t)B * d.hist_cap * gd.state_bytes)); d.hist_state = hs; }
    A_(dev_alloc(e, &d.hist_pi, (size_t)B * d.hist_cap * gd.A));
    A_(dev_alloc(e, &d.hist_len, B)); A_(dev_alloc(e, &d.next_reset, B)); A_(dev_alloc(e, &d.noise_event, B));
    A_(dev_alloc(e, &d.last_action, B)); A_(dev_alloc(e, &d.fin_code, B));
    A_(dev_alloc(e, &d.emit_off, B)); A_(dev_alloc(e, &d.stats, B));
    const int obs = gd.obs_c * gd.obs_h * gd.obs_w;
    A_(dev_alloc(e, &d.obs, (size_t)B * obs)); A_(dev_alloc(e, &d.policy, (size_t)B * gd.A));
    A_(dev_alloc(e, &d.value, (size_t)B * 3));
    A_(dev_alloc(e, &d.nn_rows, (size_t)B)); A_(dev_alloc(e, &d.nn_count, 4));
    float *wp = nullptr, *wv = nullptr;
    A_(dev_alloc(e, &wp, gd.A)); A_(dev_alloc(e, &wv, 3));
    // queues
    long long quota = cfg->games_per_iteration > 0 ? cfg->games_per_iteration : (1LL << 62);
    d.quota = quota;
    const int per = cfg->symmetric_samples ? gd.nsym : 1;
    long long scap = cfg->sample_capacity;
    if (scap <= 0) {
        long long games = quota < (long long)B * 4 ? quota : (long long)B * 4;
        if (games < B) games = B;
        scap = games * gd.max_turns * per;
        long long bytes_per = (long long)(obs + gd.A + 4) * 4;
        long long cap_bytes = 8LL << 30;
        if (scap * bytes_per > cap_bytes) scap = cap_bytes / bytes_per;
    }
    d.s_cap = scap;
    d.r_cap = quota < (1LL << 40) ? quota + B : (long long)B * 64;
    if (d.r_cap < (long long)B * 2) d.r_cap = (long long)B * 2;
    A_(dev_alloc(e, &d.s_obs, (size_t)scap * obs, false)); A_(dev_alloc(e, &d.s_pi, (size_t)scap * gd.A, false));
    A_(dev_alloc(e, &d.s_z, (size_t)scap * 3, false)); A_(dev_alloc(e, &d.s_slot, (size_t)scap, false));
    A_(dev_alloc(e, &d.r_slot, (size_t)d.r_cap)); A_(dev_alloc(e, &d.r_turns, (size_t)d.r_cap));
    A_(dev_alloc(e, &d.r_win, (size_t)d.r_cap * 3));
    A_(dev_alloc(e, &d.counters, 1)); A_(dev_alloc(e, &d.err, 1));
    A_(dev_alloc(e, &e->scratch_i32, (size_t)B * gd.A)); A_(dev_alloc(e, &e->scratch_i8, (size_t)B * gd.cells));
    // temperature table (float32 of the doubles, as probs(gs, float temp) casts)
    int tl = cfg->temp_table_len > 0 ? cfg->temp_table_len : 1;
    e->temp_host.resize(tl);
    for (int i = 0; i < tl; i++) e->temp_host[i] = cfg->temp_table_len > 0 ? (float)cfg->temp_table[i] : 1.0f;
    float *tt = nullptr;
    A_(dev_alloc(e, &tt, tl));
#undef A_
    if (rc) { azb_destroy(e); return rc; }
    cudaError_t ce = cudaMemcpy(tt, e->temp_host.data(), sizeof(float) * tl, cudaMemcpyHostToDevice);
    std::vector<float> wph(gd.A, (float)(1.0 / gd.A));
    float wvh[3] = {(float)(1.0 / 3), (float)(1.0 / 3), (float)(1.0 / 3)};
    if (ce == cudaSuccess) ce = cudaMemcpy(wp, wph.data(), sizeof(float) * gd.A, cudaMemcpyHostToDevice);
    if (ce == cudaSuccess) ce = cudaMemcpy(wv, wvh, sizeof(wvh), cudaMemcpyHostToDevice);
    if (ce != cudaSuccess) { azb_destroy(e); return fail(AZB_ERR_CUDA, "upload failed: %s", cudaGetErrorString(ce)); }
    d.temp_table = tt; d.temp_len = tl;
    d.warm_policy = wp; d.warm_value = wv;
    d.cpuct = cfg->cpuct; d.fpu_reduction = cfg->fpu_reduction; d.noise_frac = cfg->root_noise_frac;
    d.root_temp_exp = (float)(1.0 / (double)cfg->root_policy_temp);
    d.add_noise = cfg->add_root_noise; d.add_temp = cfg->add_root_temp; d.rng_mode = cfg->rng_mode;
    d.symmetric = cfg->symmetric_samples; d.reset_threshold = cfg->mcts_reset_threshold;
    d.arena = cfg->arena ? 1 : 0;
    d.seed = cfg->seed; d.gid_base = cfg->game_id_base;
    rc = init_slots(e, cfg->mt_seeds);
    if (rc) { azb_destroy(e); return rc; }
    *out = e;
    return AZB_OK;
}

extern "C" int azb_destroy(azb_engine *e)
{
    if (!e) return AZB_OK;
    cudaSetDevice(e->cfg.device);
    cudaDeviceSynchronize();
    for (void *p : e->allocs) cudaFree(p);
    if (e->noise_dev) cudaFree(e->noise_dev);
    delete e;
    return AZB_OK;
}

extern "C" int azb_reset_games(azb_engine *e, uint64_t seed, const uint32_t *mt_seeds)
{
    if (!e) return fail(AZB_ERR_BAD_ARGUMENT, "null engine");
    CK(cudaSetDevice(e->cfg.device));
    CK(cudaDeviceSynchronize());
    e->d.seed = seed;
    e->cfg.seed = seed;
    return init_slots(e, mt_seeds);
}

extern "C" int azb_set_quota(azb_engine *e, int64_t q)
{
    if (!e) return fail(AZB_ERR_BAD_ARGUMENT, "null engine");
    e->d.quota = q > 0 ? q : (1LL << 62);
    return AZB_OK;
}

extern "C" int azb_action_size(const azb_engine *e) { return e ? e->gd.A : AZB_ERR_BAD_ARGUMENT; }
extern "C" int azb_num_games(const azb_engine *e) { return e ? e->d.B : AZB_ERR_BAD_ARGUMENT; }
extern "C" int azb_observation_size(const azb_engine *e, int32_t chw[3])
{
    if (!e || !chw) return AZB_ERR_BAD_ARGUMENT;
    chw[0] = e->gd.obs_c; chw[1] = e->gd.obs_h; chw[2] = e->gd.obs_w;
    return AZB_OK;
}
extern "C" float *azb_obs_ptr(azb_engine *e) { return e ? e->d.obs : nullptr; }
extern "C" float *azb_policy_ptr(azb_engine *e) { return e ? e->d.policy : nullptr; }
extern "C" float *azb_value_ptr(azb_engine *e) { return e ? e->d.value : nullptr; }
extern "C" int32_t *azb_nn_rows_ptr(azb_engine *e) { return e ? e->d.nn_rows : nullptr; }
extern "C" int32_t *azb_nn_count_ptr(azb_engine *e) { return e ? e->d.nn_count + 2 * e->d.nn_par : nullptr; }
extern "C" int32_t *azb_arena_rows_ptr(azb_engine *e, int32_t model)
{
    return (e && e->d.arena && (model == 0 || model == 1)) ? e->d.nn_rows + (size_t)model * (size_t)(e->d.B / 2) : nullptr;
}
extern "C" int32_t *azb_arena_count_ptr(azb_engine *e, int32_t model)
{
    return (e && e->d.arena && (model == 0 || model == 1)) ? e->d.nn_count + 2 * e->d.nn_par + model : nullptr;
}
extern "C" int azb_arena_set_player_to_index(azb_engine *e, int32_t model_of_player0)
{
    if (!e || !e->d.arena || (model_of_player0 != 0 && model_of_player0 != 1)) return fail(AZB_ERR_BAD_ARGUMENT, "arena engine, model 0 or 1");
    e->d.arena_swap = model_of_player0;
    return AZB_OK;
}

static int range_ok(azb_engine *e, int32_t first, int32_t &count)
{
    if (!e) return fail(AZB_ERR_BAD_ARGUMENT, "null engine");
    if (count == 0) count = e->d.B - first;
    if (first < 0 || count <= 0 || first + count > e->d.B)
        return fail(AZB_ERR_BAD_ARGUMENT, "slot range [%d, %d) outside [0, %d)", first, first + count, e->d.B);
    return 0;
}

extern "C" int azb_select(azb_engine *e, int32_t first, int32_t count, void *stream)
{
    TRY(range_ok(e, first, count));
    cudaStream_t s = (cudaStream_t)stream;
    e->d.nn_par ^= 1;
    dd_next_epoch(e);
    DISPATCH(e, l_select, e, first, count, s);
    CK(cudaGetLastError());
    return AZB_OK;
}

extern "C" int azb_expand_backup(azb_engine *e, int32_t first, int32_t count, const float *policy,
                                 const float *value, void *stream)
{
    TRY(range_ok(e, first, count));
    cudaStream_t s = (cudaStream_t)stream;
    const float *pol = policy ? policy : e->d.policy;
    const float *val = value ? value : e->d.value;
    DISPATCH(e, l_expand, e, first, count, pol, val, s);
    CK(cudaGetLastError());
    return AZB_OK;
}

extern "C" int azb_expand_backup_select(azb_engine *e, int32_t first, int32_t count, const float *policy,
                                        const float *value, void *stream)
{
    TRY(range_ok(e, first, count));
    cudaStream_t s = (cudaStream_t)stream;
    const float *pol = policy ? policy : e->d.policy;
    const float *val = value ? value : e->d.value;
    e->d.nn_par ^= 1;
    dd_next_epoch(e);
    DISPATCH(e, l_expand_select, e, first, count, pol, val, s);
    CK(cudaGetLastError());
    return AZB_OK;
}

extern "C" int azb_play_moves(azb_engine *e, int32_t fast, void *stream)
{
    if (!e) return fail(AZB_ERR_BAD_ARGUMENT, "null engine");
    cudaStream_t s = (cudaStream_t)stream;
    DISPATCH(e, l_play, e, fast, s);
    e->d.nn_par = 1;                    // the next select uses counter 0 (k_finalize cleared both)
    CK(cudaGetLastError());
    if (dd_restart(e, s) != 0) return fail(AZB_ERR_CUDA, "leaf de-duplication table clear failed");
    return AZB_OK;
}

// Leaf de-duplication: games whose leaves have bit-equal observations share one network evaluation -- only the first
// such game of a select launch is listed in azb_nn_rows_ptr, the others read its policy / value rows in expand/backup.
extern "C" int azb_set_leaf_dedup(azb_engine *e, int32_t on)
{
    if (!e) return fail(AZB_ERR_BAD_ARGUMENT, "null engine");
    if (!on) { e->d.dd_table = nullptr; return AZB_OK; }
    if (e->d.arena) return fail(AZB_ERR_BAD_ARGUMENT, "leaf de-duplication: not in arena mode (one row list per model)");
    if (e->d.B > (1 << 20)) return fail(AZB_ERR_BAD_ARGUMENT, "leaf de-duplication: at most 2^20 games per engine");
    CK(cudaSetDevice(e->cfg.device));
    CK(cudaDeviceSynchronize());
    if (!e->dd_table) {
        unsigned slots = 1024;
        while (slots < 2u * (unsigned)e->d.B) slots <<= 1;
        e->dd_slots = slots;
        TRY(dev_alloc(e, &e->dd_table, (size_t)slots));
        unsigned char *st = nullptr;
        TRY(dev_alloc(e, &st, (size_t)e->d.B * (size_t)e->gd.state_bytes));
        e->d.dd_state = st;
        TRY(dev_alloc(e, &e->d.dd_src, (size_t)e->d.B));
        TRY(dev_alloc(e, &e->d.dd_dups, 1));
        e->d.dd_mask = slots - 1u;
    }
    std::vector<int> ident((size_t)e->d.B);
    for (int i = 0; i < e->d.B; i++) ident[(size_t)i] = i;
    CK(cudaMemcpy(e->d.dd_src, ident.data(), ident.size() * sizeof(int), cudaMemcpyHostToDevice));
    CK(cudaMemset(e->dd_table, 0, (size_t)e->dd_slots * sizeof(unsigned long long)));
    e->dd_epoch = 0;
    e->d.dd_table = e->dd_table;
    return AZB_OK;
}

extern "C" int azb_duplicate_leaves(azb_engine *e, int64_t *out)
{
    if (!e || !out) return fail(AZB_ERR_BAD_ARGUMENT, "null argument");
    *out = 0;
    if (!e->dd_table) return AZB_OK;
    CK(cudaSetDevice(e->cfg.device));
    unsigned long long v = 0;
    CK(cudaMemcpy(&v, e->d.dd_dups, sizeof(v), cudaMemcpyDeviceToHost));
    *out = (int64_t)v;
    return AZB_OK;
}

extern "C" int azb_set_state(azb_engine *e, int32_t slot, const int8_t *cells, int32_t turns, void *stream)
{
    if (!e || !cells) return fail(AZB_ERR_BAD_ARGUMENT, "null argument");
    if (slot < 0 || slot >= e->d.B || turns < 0) return fail(AZB_ERR_BAD_ARGUMENT, "slot %d / turns %d", slot, turns);
    cudaStream_t s = (cudaStream_t)stream;
    CK(cudaMemcpyAsync(e->scratch_i8, cells, (size_t)e->gd.cells, cudaMemcpyHostToDevice, s));
    DISPATCH(e, l_set_state, e, slot, (const signed char *)e->scratch_i8, turns, s);
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(s));
    return AZB_OK;
}

extern "C" int azb_force_move(azb_engine *e, int32_t slot, int32_t action, void *stream)
{
    if (!e) return fail(AZB_ERR_BAD_ARGUMENT, "null engine");
    if (slot < 0 || slot >= e->d.B || action < 0 || action >= e->gd.A)
        return fail(AZB_ERR_BAD_ARGUMENT, "slot %d / action %d", slot, action);
    cudaStream_t s = (cudaStream_t)stream;
    DISPATCH(e, l_force_move, e, slot, action, s);
    CK(cudaGetLastError());
    return AZB_OK;
}

extern "C" int azb_set_root_flags(azb_engine *e, int32_t add_root_noise, int32_t add_root_temp)
{
    if (!e) return fail(AZB_ERR_BAD_ARGUMENT, "null engine");
    if (e->d.arena && (add_root_noise || add_root_temp)) return fail(AZB_ERR_BAD_ARGUMENT, "arena mode has no root noise / temperature");
    e->d.add_noise = add_root_noise ? 1 : 0;
    e->d.add_temp = add_root_temp ? 1 : 0;
    return AZB_OK;
}

extern "C" int azb_arena_players(azb_engine *e, int32_t *players_device, void *stream)
{
    if (!e || !players_device) return fail(AZB_ERR_BAD_ARGUMENT, "null argument");
    if (!e->d.arena) return fail(AZB_ERR_BAD_ARGUMENT, "engine is not in arena mode");
    cudaStream_t s = (cudaStream_t)stream;
    DISPATCH(e, l_arena_players, e, players_device, s);
    CK(cudaGetLastError());
    return AZB_OK;
}

extern "C" int azb_warmup_sims(azb_engine *e, int32_t sims, void *stream)
{
    if (!e) return fail(AZB_ERR_BAD_ARGUMENT, "null engine");
    if (sims <= 0) return fail(AZB_ERR_BAD_ARGUMENT, "sims must be positive");
    cudaStream_t s = (cudaStream_t)stream;
    DISPATCH(e, l_warmup, e, sims, s);
    CK(cudaGetLastError());
    return AZB_OK;
}

extern "C" int azb_set_root_noise(azb_engine *e, const float *noise, int32_t events, int32_t stride)
{
    if (!e) return fail(AZB_ERR_BAD_ARGUMENT, "null engine");
    CK(cudaSetDevice(e->cfg.device));
    CK(cudaDeviceSynchronize());
    if (e->noise_dev) { CK(cudaFree(e->noise_dev)); e->noise_dev = nullptr; }
    e->d.noise = nullptr; e->d.noise_events = 0; e->d.noise_stride = 0;
    if (!noise || events <= 0 || stride <= 0) return AZB_OK;
    size_t n = (size_t)e->d.B * (size_t)events * (size_t)stride;
    CK(cudaMalloc(&e->noise_dev, n * sizeof(float)));
    CK(cudaMemcpy(e->noise_dev, noise, n * sizeof(float), cudaMemcpyHostToDevice));
    CK(cudaMemset(e->d.noise_event, 0, sizeof(int) * (size_t)e->d.B));
    e->d.noise = e->noise_dev; e->d.noise_events = events; e->d.noise_stride = stride;
    return AZB_OK;
}

static int read_counters(azb_engine *e, Counters *c, cudaStream_t s)
{
    CK(cudaMemcpyAsync(c, e->d.counters, sizeof(Counters), cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    return 0;
}

static int drain_samples_impl(azb_engine *e, float *obs, float *pi, float *z, int32_t *slot, int64_t capacity,
                              int64_t *count, cudaStream_t s, cudaMemcpyKind kind)
{
    if (!e || !count) return fail(AZB_ERR_BAD_ARGUMENT, "null argument");
    Counters c;
    TRY(read_counters(e, &c, s));
    long long n = c.sample_count;
    if (n > capacity) return fail(AZB_ERR_BAD_ARGUMENT, "sample buffers hold %lld, %lld queued", (long long)capacity, n);
    const size_t obsn = (size_t)e->gd.obs_c * e->gd.obs_h * e->gd.obs_w;
    if (n > 0) {
        if (obs) CK(cudaMemcpyAsync(obs, e->d.s_obs, sizeof(float) * obsn * (size_t)n, kind, s));
        if (pi) CK(cudaMemcpyAsync(pi, e->d.s_pi, sizeof(float) * (size_t)e->gd.A * (size_t)n, kind, s));
        if (z) CK(cudaMemcpyAsync(z, e->d.s_z, sizeof(float) * 3 * (size_t)n, kind, s));
        if (slot) CK(cudaMemcpyAsync(slot, e->d.s_slot, sizeof(int32_t) * (size_t)n, kind, s));
    }
    CK(cudaMemsetAsync(&e->d.counters->sample_count, 0, sizeof(long long), s));
    CK(cudaStreamSynchronize(s));
    *count = n;
    return AZB_OK;
}

extern "C" int azb_drain_samples(azb_engine *e, float *obs, float *pi, float *z, int32_t *slot, int64_t capacity,
                                 int64_t *count, void *stream)
{
    return drain_samples_impl(e, obs, pi, z, slot, capacity, count, (cudaStream_t)stream, cudaMemcpyDeviceToHost);
}

extern "C" int azb_drain_samples_device(azb_engine *e, float *obs, float *pi, float *z, int32_t *slot,
                                        int64_t capacity, int64_t *count, void *stream)
{
    return drain_samples_impl(e, obs, pi, z, slot, capacity, count, (cudaStream_t)stream, cudaMemcpyDeviceToDevice);
}

extern "C" int azb_sample_count(azb_engine *e, int64_t *count, void *stream)
{
    if (!e || !count) return fail(AZB_ERR_BAD_ARGUMENT, "null argument");
    Counters c;
    TRY(read_counters(e, &c, (cudaStream_t)stream));
    *count = c.sample_count;
    return AZB_OK;
}

extern "C" int azb_games_played(azb_engine *e, int64_t *count, void *stream)
{
    if (!e || !count) return fail(AZB_ERR_BAD_ARGUMENT, "null argument");
    Counters c;
    TRY(read_counters(e, &c, (cudaStream_t)stream));
    *count = c.games_played;
    return AZB_OK;
}

extern "C" int azb_drain_results(azb_engine *e, int32_t *slot, int32_t *turns, uint8_t *win, int64_t capacity,
                                 int64_t *count, void *stream)
{
    if (!e || !count) return fail(AZB_ERR_BAD_ARGUMENT, "null argument");
    cudaStream_t s = (cudaStream_t)stream;
    Counters c;
    TRY(read_counters(e, &c, s));
    long long n = c.result_count;
    if (n > capacity) return fail(AZB_ERR_BAD_ARGUMENT, "result buffers hold %lld, %lld queued", (long long)capacity, n);
    if (n > 0) {
        if (slot) CK(cudaMemcpyAsync(slot, e->d.r_slot, sizeof(int32_t) * (size_t)n, cudaMemcpyDeviceToHost, s));
        if (turns) CK(cudaMemcpyAsync(turns, e->d.r_turns, sizeof(int32_t) * (size_t)n, cudaMemcpyDeviceToHost, s));
        if (win) CK(cudaMemcpyAsync(win, e->d.r_win, 3 * (size_t)n, cudaMemcpyDeviceToHost, s));
    }
    CK(cudaMemsetAsync(&e->d.counters->result_count, 0, sizeof(long long), s));
    CK(cudaStreamSynchronize(s));
    *count = n;
    return AZB_OK;
}

extern "C" int azb_root_counts(azb_engine *e, int32_t *counts, void *stream)
{
    if (!e || !counts) return fail(AZB_ERR_BAD_ARGUMENT, "null argument");
    cudaStream_t s = (cudaStream_t)stream;
    const int B = e->d.B;
    DISPATCH(e, l_counts, e, s);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(counts, e->scratch_i32, sizeof(int32_t) * (size_t)B * e->gd.A, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    return AZB_OK;
}

extern "C" int azb_game_info(azb_engine *e, int32_t *last_action, int32_t *turns, void *stream)
{
    if (!e) return fail(AZB_ERR_BAD_ARGUMENT, "null engine");
    cudaStream_t s = (cudaStream_t)stream;
    const int B = e->d.B;
    if (last_action) CK(cudaMemcpyAsync(last_action, e->d.last_action, sizeof(int32_t) * (size_t)B, cudaMemcpyDeviceToHost, s));
    std::vector<unsigned char> hd;
    const size_t hb = (size_t)e->gd.head_bytes;
    if (turns) {
        hd.resize(hb * (size_t)B);
        CK(cudaMemcpyAsync(hd.data(), e->d.head, hb * (size_t)B, cudaMemcpyDeviceToHost, s));
    }
    CK(cudaStreamSynchronize(s));
    if (turns) for (int i = 0; i < B; i++) turns[i] = head_view(e->gd, hd.data() + hb * (size_t)i).turns;
    return AZB_OK;
}

extern "C" int azb_boards(azb_engine *e, int8_t *cells, void *stream)
{
    if (!e || !cells) return fail(AZB_ERR_BAD_ARGUMENT, "null argument");
    cudaStream_t s = (cudaStream_t)stream;
    const int B = e->d.B;
    DISPATCH(e, l_boards, e, s);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(cells, e->scratch_i8, (size_t)B * e->gd.cells, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    return AZB_OK;
}

extern "C" int azb_tree_dump(azb_engine *e, int32_t slot, double *rows, int64_t max_rows, int64_t *rows_out, void *stream)
{
    if (!e || !rows || !rows_out) return fail(AZB_ERR_BAD_ARGUMENT, "null argument");
    if (slot < 0 || slot >= e->d.B) return fail(AZB_ERR_BAD_ARGUMENT, "slot %d", slot);
    cudaStream_t s = (cudaStream_t)stream;
    CK(cudaStreamSynchronize(s));
    std::vector<unsigned char> hraw((size_t)e->gd.head_bytes);
    CK(cudaMemcpy(hraw.data(), (const unsigned char *)e->d.head + (size_t)slot * e->gd.head_bytes, hraw.size(), cudaMemcpyDeviceToHost));
    const HeadView hd = head_view(e->gd, hraw.data());
    const int used = hd.alloc, root = hd.root;      // entries [0, alloc) cover the live half (and, in the upper half, stale ones below it)
    const size_t nb = (size_t)slot * (size_t)e->d.npg;
    std::vector<NodeHot> hot(used); std::vector<NodeCold> cold(used);
    CK(cudaMemcpy(hot.data(), e->d.hot + nb, sizeof(NodeHot) * used, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(cold.data(), e->d.cold + nb, sizeof(NodeCold) * used, cudaMemcpyDeviceToHost));
    // while a node is the root its live fields are in the slot header
    hot[root].n = hd.root_n; hot[root].child0 = hd.root_child0; cold[root].v = hd.root_v;
    cold[root].meta = (cold[root].meta & META_ACTION_MASK) | (hd.root_meta & ~META_ACTION_MASK);
    if (!hd.root_rec) {
        cold[root].meta = hd.root_meta; hot[root].q = 0.0f; hot[root].p = 0.0f;                       // fresh root: no pool record
    }
    std::vector<int> n(used), c0(used); std::vector<float> q(used), p(used), v(used); std::vector<uint32_t> m(used);
    for (int i = 0; i < used; i++) {
        n[i] = hot[i].n; c0[i] = hot[i].child0; q[i] = hot[i].q; p[i] = hot[i].p; v[i] = cold[i].v; m[i] = cold[i].meta;
    }
    int64_t w = 0;
    std::vector<std::pair<int, int>> stack;   // (node, depth)
    stack.push_back({root, 0});
    while (!stack.empty() && w < max_rows) {
        auto [nd, dep] = stack.back();
        stack.pop_back();
        double *r = rows + w * 10;
        int a = meta_action(m[nd]);
        int ec = meta_e(m[nd]);
        r[0] = dep; r[1] = (a == (int)META_ACTION_NONE) ? -1 : a; r[2] = n[nd]; r[3] = q[nd]; r[4] = v[nd];
        r[5] = p[nd]; r[6] = meta_player(m[nd]); r[7] = ec == 1; r[8] = ec == 2; r[9] = ec == 3;
        w++;
        int nc = meta_nc(m[nd]);
        for (int k = nc - 1; k >= 0; k--) stack.push_back({c0[nd] + k, dep + 1});
    }
    *rows_out = w;
    return AZB_OK;
}

extern "C" int azb_stats_get(azb_engine *e, azb_stats *out, void *stream)
{
    if (!e || !out) return fail(AZB_ERR_BAD_ARGUMENT, "null argument");
    cudaStream_t s = (cudaStream_t)stream;
    const int B = e->d.B;
    std::vector<SlotStats> ss(B);
    Counters c;
    CK(cudaMemcpyAsync(ss.data(), e->d.stats, sizeof(SlotStats) * (size_t)B, cudaMemcpyDeviceToHost, s));
    TRY(read_counters(e, &c, s));
    memset(out, 0, sizeof(*out));
    const size_t hb = (size_t)e->gd.head_bytes;
    std::vector<unsigned char> hraw(hb * (size_t)B);
    CK(cudaMemcpyAsync(hraw.data(), e->d.head, hb * (size_t)B, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    for (int i = 0; i < B; i++) {
        out->sims += (int64_t)ss[i].sims; out->sum_depth += (int64_t)ss[i].sum_depth;
        out->sum_children += (int64_t)ss[i].sum_children; out->nodes_created += (int64_t)ss[i].nodes_created;
        out->terminal_leaves += (int64_t)ss[i].terminal_leaves; out->moves += (int64_t)ss[i].moves;
        const HeadView hv = head_view(e->gd, hraw.data() + hb * (size_t)i);
        const int used = hv.alloc - (hv.root >= e->d.half ? e->d.half : 0);
        const int pk = ss[i].peak_nodes > used ? ss[i].peak_nodes : used;
        if (pk > out->peak_nodes) out->peak_nodes = pk;
    }
    out->games_played = c.games_played; out->results = c.results; out->samples = c.samples_total;
    out->pool_bytes = e->pool_bytes; out->device_bytes = e->device_bytes;
    return AZB_OK;
}

extern "C" int azb_check_errors(azb_engine *e, void *stream)
{
    if (!e) return fail(AZB_ERR_BAD_ARGUMENT, "null engine");
    cudaStream_t s = (cudaStream_t)stream;
    uint32_t w = 0;
    CK(cudaMemcpyAsync(&w, e->d.err, sizeof(w), cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    if (w == 0) return AZB_OK;
    CK(cudaMemsetAsync(e->d.err, 0, sizeof(w), s));
    if (w & ERRB_POOL) return fail(AZB_ERR_POOL_EXHAUSTED, "node pool exhausted (%d entries per game); raise max_nodes_per_game", e->d.npg);
    if (w & ERRB_ACTION) return fail(AZB_ERR_INVALID_ACTION, "Invalid action encountered while updating root");
    if (w & ERRB_FP) return fail(AZB_ERR_FLOATING_POINT, "zero or NaN prior / visit-count sum");
    if (w & ERRB_NOISE) return fail(AZB_ERR_NOISE_UNDERRUN, "root noise requested but the fed table is exhausted or absent");
    return fail(AZB_ERR_SAMPLE_OVERFLOW, "sample / result / history buffer full; entries were dropped");
}
