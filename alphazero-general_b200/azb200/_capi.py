"""ctypes declarations of libazb200.so (include/azb200.h).  The library is the
product: importing this module fails loudly when it has not been built --
there is no CPU or PyTorch fallback for the hot path."""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libazb200.so")

ABI_VERSION = 1
GAME_CONNECT4, GAME_BRANDUBH, GAME_HNEFATAFL = 0, 1, 2
RNG_MT19937, RNG_PHILOX = 0, 1

STATUS_NAMES = {
    0: "AZB_OK", -1: "AZB_ERR_BAD_CONFIG", -2: "AZB_ERR_CUDA", -3: "AZB_ERR_POOL_EXHAUSTED",
    -4: "AZB_ERR_INVALID_ACTION", -5: "AZB_ERR_FLOATING_POINT", -6: "AZB_ERR_SAMPLE_OVERFLOW",
    -7: "AZB_ERR_BAD_ARGUMENT", -8: "AZB_ERR_NOISE_UNDERRUN",
}


class AzbConfig(C.Structure):
    _fields_ = [
        ("abi_version", C.c_int32), ("game", C.c_int32), ("num_games", C.c_int32), ("device", C.c_int32),
        ("rng_mode", C.c_int32), ("add_root_noise", C.c_int32), ("add_root_temp", C.c_int32),
        ("symmetric_samples", C.c_int32), ("mcts_reset_threshold", C.c_int32),
        ("max_sims_per_move", C.c_int32), ("max_nodes_per_game", C.c_int32), ("temp_table_len", C.c_int32),
        ("lanes_per_game", C.c_int32), ("arena", C.c_int32),
        ("games_per_iteration", C.c_int64), ("sample_capacity", C.c_int64), ("game_id_base", C.c_int64),
        ("seed", C.c_uint64),
        ("cpuct", C.c_float), ("fpu_reduction", C.c_float), ("root_noise_frac", C.c_float),
        ("root_policy_temp", C.c_float),
        ("temp_table", C.POINTER(C.c_double)), ("mt_seeds", C.POINTER(C.c_uint32)),
    ]


class AzbStats(C.Structure):
    _fields_ = [(n, C.c_int64) for n in (
        "sims", "sum_depth", "sum_children", "nodes_created", "terminal_leaves", "games_played",
        "results", "samples", "moves", "peak_nodes", "pool_bytes", "device_bytes")]


# every symbol include/azb200.h declares: name -> (restype, argtypes)
_vp, _i32, _i64 = C.c_void_p, C.c_int32, C.c_int64
SYMBOLS = {
    "azb_create": (C.c_int, [C.POINTER(AzbConfig), C.POINTER(_vp)]),
    "azb_destroy": (C.c_int, [_vp]),
    "azb_reset_games": (C.c_int, [_vp, C.c_uint64, _vp]),
    "azb_set_quota": (C.c_int, [_vp, _i64]),
    "azb_action_size": (C.c_int, [_vp]),
    "azb_observation_size": (C.c_int, [_vp, C.POINTER(_i32 * 3)]),
    "azb_num_games": (C.c_int, [_vp]),
    "azb_obs_ptr": (_vp, [_vp]),
    "azb_policy_ptr": (_vp, [_vp]),
    "azb_value_ptr": (_vp, [_vp]),
    "azb_nn_rows_ptr": (_vp, [_vp]),
    "azb_nn_count_ptr": (_vp, [_vp]),
    "azb_select": (C.c_int, [_vp, _i32, _i32, _vp]),
    "azb_expand_backup": (C.c_int, [_vp, _i32, _i32, _vp, _vp, _vp]),
    "azb_expand_backup_select": (C.c_int, [_vp, _i32, _i32, _vp, _vp, _vp]),
    "azb_play_moves": (C.c_int, [_vp, _i32, _vp]),
    "azb_warmup_sims": (C.c_int, [_vp, _i32, _vp]),
    "azb_arena_players": (C.c_int, [_vp, _vp, _vp]),
    "azb_arena_set_player_to_index": (C.c_int, [_vp, _i32]),
    "azb_arena_rows_ptr": (_vp, [_vp, _i32]),
    "azb_arena_count_ptr": (_vp, [_vp, _i32]),
    "azb_set_state": (C.c_int, [_vp, _i32, _vp, _i32, _vp]),
    "azb_force_move": (C.c_int, [_vp, _i32, _i32, _vp]),
    "azb_set_root_flags": (C.c_int, [_vp, _i32, _i32]),
    "azb_set_leaf_dedup": (C.c_int, [_vp, _i32]),
    "azb_duplicate_leaves": (C.c_int, [_vp, _vp]),
    "azb_set_root_noise": (C.c_int, [_vp, _vp, _i32, _i32]),
    "azb_drain_samples": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _i64, C.POINTER(_i64), _vp]),
    "azb_drain_samples_device": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _i64, C.POINTER(_i64), _vp]),
    "azb_sample_count": (C.c_int, [_vp, C.POINTER(_i64), _vp]),
    "azb_drain_results": (C.c_int, [_vp, _vp, _vp, _vp, _i64, C.POINTER(_i64), _vp]),
    "azb_games_played": (C.c_int, [_vp, C.POINTER(_i64), _vp]),
    "azb_root_counts": (C.c_int, [_vp, _vp, _vp]),
    "azb_game_info": (C.c_int, [_vp, _vp, _vp, _vp]),
    "azb_boards": (C.c_int, [_vp, _vp, _vp]),
    "azb_tree_dump": (C.c_int, [_vp, _i32, _vp, _i64, C.POINTER(_i64), _vp]),
    "azb_stats_get": (C.c_int, [_vp, C.POINTER(AzbStats), _vp]),
    "azb_check_errors": (C.c_int, [_vp, _vp]),
    "azb_last_error": (C.c_char_p, []),
    "azb_abi_version": (C.c_int, []),
    # include/azb200_nn.h (weight structs are passed by address: fused_nn._NNWeights / nn_tc._NNGNet)
    "azb_nn_forward": (C.c_int, [_vp, _vp, _vp, _vp, _i32, _vp]),
    "azb_nn_forward_tc": (C.c_int, [_vp, _vp, _vp, _vp, _i32, _vp]),
    "azb_nn_forward_tc_rows": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _i32, _vp]),
    "azb_nn_forward_tc_debug": (C.c_int, [_vp, _vp, _vp, _vp, _i32, _vp, _vp, _i32]),
    "azb_nn_tc_layer_bytes": (C.c_int, []),
    "azb_nn_tc_head_row_stride": (C.c_int, []),
    "azb_nn_tc_boards_per_cta": (C.c_int, []),
    "azb_nn_tc_frame_rows_per_board": (C.c_int, []),
    "azb_nn_weight_row_stride": (C.c_int, []),
    "azb_nn_head_row_stride": (C.c_int, []),
    "azb_nn_boards_per_cta": (C.c_int, []),
    "azb_upload_pinned": (C.c_int, [_vp, _vp, _i64, _vp]),
    "azb_nng_layout": (C.c_int, [_i32, _i32, _vp]),
    "azb_nng_tile_plan": (C.c_int, [_i32, _i32, _i32, _i32, _i32, _i32, _vp]),
    "azb_nng_forward": (C.c_int, [_vp, _vp, _vp, _vp, _i32, _vp, _vp, _vp]),
    "azb_nng_forward_debug": (C.c_int, [_vp, _vp, _vp, _vp, _i32, _vp, _vp, _i32]),
    "azb_nng_trace": (C.c_int, [_vp, _i32]),
    "azb_nng_cta_trace": (C.c_int, [_vp, _i32]),
}

_lib = None


class AzbError(RuntimeError):
    def __init__(self, status, message):
        super().__init__(f"{STATUS_NAMES.get(status, status)}: {message}")
        self.status = status


def load():
    """dlopen libazb200.so and bind every exported symbol; raises if missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `make -C alphazero-general_b200/csrc` "
            "(or __graft_entry__.build()). The engine has no fallback path.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)           # AttributeError if the library lacks a declared symbol
        fn.restype = res
        fn.argtypes = args
    if lib.azb_abi_version() != ABI_VERSION:
        raise ImportError(f"libazb200.so ABI {lib.azb_abi_version()} != binding {ABI_VERSION}")
    _lib = lib
    return lib


def check(status):
    if status != 0:
        msg = load().azb_last_error()
        raise AzbError(status, msg.decode() if msg else "")
