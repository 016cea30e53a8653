"""azb200 -- B200 (sm_100a) batched self-play MCTS engine behind the
Game / NNetWrapper / Coach surface of kevaday/alphazero-general.

The hot path lives in libazb200.so (CUDA, built from ../csrc); this package is
the host-side mirror of the reference's SelfPlayAgent / Coach self-play
interface on top of its C ABI (include/azb200.h)."""
from ._capi import AzbError, load as load_library  # noqa: F401
from .engine import SelfPlayEngine, default_temp_scaling, temp_table  # noqa: F401

__all__ = ["SelfPlayEngine", "AzbError", "load_library", "default_temp_scaling", "temp_table"]
