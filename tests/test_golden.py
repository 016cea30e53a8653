"""Golden traces recorded from the compiled reference (tests/golden/*.npz):
the CPU oracle must reproduce them (pins the oracle, runs everywhere) and so
must the CUDA engine through the C ABI (GPU)."""
import pytest

import _golden
import _orc


@pytest.mark.parametrize("name", _golden.cases())
def test_oracle_reproduces_reference_golden(name):
    g = _golden.load(name)
    game, kw = _golden.agent_kwargs(g)
    ag = _orc.OracleAgent(_golden.GAME_IDS[game], len(g["seeds"]), rng_mode=_orc.RNG_MT19937,
                          mt_seeds=g["seeds"].tolist(), **kw)
    _golden.check(ag, g, _golden.replay(ag, g))


@pytest.mark.gpu
@pytest.mark.parametrize("name", _golden.cases(engine=True))
def test_engine_reproduces_reference_golden(name):
    from _engine_agent import EngineAgent
    g = _golden.load(name)
    game, kw = _golden.agent_kwargs(g)
    ag = EngineAgent(game, len(g["seeds"]), rng="mt19937", mt_seeds=g["seeds"].tolist(),
                     max_sims_per_move=int(g["sims"]), **kw)
    _golden.check(ag, g, _golden.replay(ag, g))


@pytest.mark.parametrize("name", _golden.cases(arena=True))
def test_oracle_reproduces_reference_arena_golden(name):
    """SelfPlayAgent(_is_arena=True) traces recorded from the compiled reference (two networks, arenaTemp)."""
    g = _golden.load(name)
    game, kw = _golden.agent_kwargs(g)
    ag = _orc.OracleAgent(_golden.GAME_IDS[game], len(g["seeds"]), rng_mode=_orc.RNG_MT19937,
                          mt_seeds=g["seeds"].tolist(), arena=True, **kw)
    _golden.check(ag, g, _golden.replay(ag, g))


@pytest.mark.gpu
@pytest.mark.parametrize("name", _golden.cases(arena=True, engine=True))
def test_engine_reproduces_reference_arena_golden(name):
    from _engine_agent import ArenaEngineAgent
    g = _golden.load(name)
    game, kw = _golden.agent_kwargs(g)
    kw["mcts_reset_threshold"] = kw["mcts_reset_threshold"] or None
    ag = ArenaEngineAgent(game, len(g["seeds"]), rng="mt19937", mt_seeds=g["seeds"].tolist(),
                          max_sims_per_move=int(g["sims"]), **kw)
    _golden.check(ag, g, _golden.replay(ag, g))
