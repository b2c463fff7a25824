"""Row b of SURVEY.md section 8 (the drop-in boundary), proven the direct way: the reference's OWN Python agent layer
-- agents/mcts.py (MCTSAgent, RandomMCTSAgent, RAVEAgent, TraditionalAgent), agents/utils.py (dual_play, eval_agents),
agents/agent.py -- runs UNMODIFIED with `core` -> `gomokuai_b200.core`.

Where it comes from: /root/reference when it is there (the build container), else oracle/_ref/pyref.zip -- the same
files compiled to sourceless byte code by `make -C oracle pyref` (a built artefact that travels to the GPU box like
oracle/_ref/libgomoku_ref.so; the sources are never copied).  GK_PYREF_ONLY=1 forces the byte-code build.  No stand-in is written here: without either, the tests
skip.  The CPU tests fill the Policy slots from Python (how agents/alphazero.py plugs a network in, core/py_ext/src/
mcts_ext.hpp:43-61), the GPU tests use RandomPolicy / PoolRAVEPolicy / TraditionalPolicy, whose simulate slots run on
the B200."""
import importlib
import os
import sys

import numpy as np
import pytest

from search_util import injected_eval_state

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CANDIDATES = ([] if os.environ.get("GK_PYREF_ONLY") else ["/root/reference"]) + [os.path.join(ROOT, "oracle", "_ref", "pyref.zip")]


@pytest.fixture(scope="module")
def ref_agents(gk):
    from gomokuai_b200 import build
    build.build_pyext()
    from gomokuai_b200 import core
    base = next((p for p in CANDIDATES if os.path.isdir(os.path.join(p, "agents")) or os.path.isfile(p)), None)
    if base is None:
        pytest.skip("neither /root/reference nor oracle/_ref/pyref.zip is present")
    saved = {k: sys.modules.get(k) for k in ("core", "agents", "config")}
    for k in list(sys.modules):
        if k == "agents" or k.startswith("agents."):
            del sys.modules[k]
    sys.modules["core"] = core                      # the drop-in: `from core import MCTS, RandomPolicy, PoolRAVEPolicy, ...`
    sys.modules.pop("config", None)
    sys.path.insert(0, base)
    try:
        agents = importlib.import_module("agents")  # agents/__init__.py:1-6 imports every agent module
        utils = importlib.import_module("agents.utils")
        origin = getattr(agents, "__file__", "") or ""
        assert origin.startswith(base), origin
        yield core, agents, utils, base
    finally:
        sys.path.remove(base)
        for k in list(sys.modules):
            if k == "agents" or k.startswith("agents."):
                del sys.modules[k]
        for k, v in saved.items():
            if v is not None:
                sys.modules[k] = v
            else:
                sys.modules.pop(k, None)


def test_reference_agent_layer_imports_against_the_drop_in(ref_agents):
    core, agents, utils, base = ref_agents
    # agents/mcts.py:1 needs all four names; agents/__init__.py:9-18 builds the map of agent factories
    assert set(agents.AGENT_MAP) == {"random", "console", "botzone", "mcts", "random_mcts", "rave_mcts", "traditional_mcts", "alphazero"}
    # the factories construct without a GPU (policy constructors do not touch the device)
    a = agents.RandomMCTSAgent(5.0, 5, c_iterations=10)
    b = agents.RAVEAgent(2.0, 0.1, c_iterations=10)
    c = agents.TraditionalAgent(5.0, c_iterations=10)
    d = agents.TraditionalAgent(5.0, 0.1, True, c_iterations=10)      # use_rave=True (policy_ext.hpp:39-45)
    e = agents.MCTSAgent(c_duration=__import__("datetime").timedelta(milliseconds=5))
    assert [repr(x) for x in (a, b, c, d, e)] == ["MCTS Agent with RandomPolicy", "MCTS Agent with PoolRAVEPolicy",
                                                  "MCTS Agent with TraditionalPolicy", "MCTS Agent with TraditionalPolicy",
                                                  "MCTS Agent with RandomPolicy"]
    r = agents.RandomAgent()                                           # agents/agent.py: Board.random_move
    board = core.Board()
    assert board.check_move(r.get_action(board))


def test_reference_dual_play_runs_on_the_mirror(ref_agents):
    """agents/utils.py:5-63 dual_play, both return formats, and eval_agents (:66-101), with Python-filled Policy slots"""
    core, agents, utils, base = ref_agents

    def agent(n):
        return agents.MCTSAgent(policy=core.Policy(eval_state=injected_eval_state, c_puct=5.0), c_iterations=n)

    players = {core.Player.black: agent(40), core.Player.white: agent(25)}
    board = core.Board()
    for c in (112, 113, 97, 98, 127, 128, 82):      # black has four in column 7 -> a short game
        board.apply_move(c)
    data = utils.dual_play(players, board, verbose=True)
    assert len(data) >= 1 and board.status["is_end"]
    for state, z, probs in data:
        assert state.shape == (6, 15, 15) and state.dtype == np.uint8 and probs.shape == (225,)
        assert float(z) in (-1.0, 0.0, 1.0) and abs(float(probs.sum()) - 1.0) < 1e-3
        assert (probs[state[2].reshape(-1) == 0] == 0).all()          # no probability on occupied cells
    [a.reset() for a in players.values()]
    winner = utils.dual_play(players, board)                           # an ended board is reset (agents/utils.py:18-21)
    assert winner in (core.Player.black, core.Player.white, core.Player.none) and board.status["is_end"]
    rates = utils.eval_agents([agent(12), agent(12)], num_games=2)
    assert abs(sum(rates) - 1.0) < 1e-9


@pytest.mark.gpu
def test_reference_agents_play_on_the_gpu(ref_agents, gpu):
    """dual_play between the reference's agent factories with the GPU-backed policies of the drop-in"""
    core, agents, utils, base = ref_agents
    core.seed(11)
    pairs = [(agents.RandomMCTSAgent(5.0, 5, c_iterations=120), agents.TraditionalAgent(5.0, c_iterations=60)),
             (agents.RAVEAgent(2.0, 0.1, c_iterations=150), agents.TraditionalAgent(5.0, 0.1, True, c_iterations=60))]
    for black, white in pairs:
        board = core.Board()
        for c in (112, 113, 97, 98):
            board.apply_move(c)
        data = utils.dual_play({core.Player.black: black, core.Player.white: white}, board, verbose=True)
        assert board.status["is_end"] and len(data) == len(board.move_record) - 4
        assert all(s.shape == (6, 15, 15) and abs(float(p.sum()) - 1.0) < 1e-3 for s, _, p in data)
        # the value targets are the final result seen from each sample's side to move (agents/utils.py:47-53)
        w = float(board.status["winner"])
        assert [float(z) for _, z, _ in data[:2]] == [w * 1.0, w * -1.0]
        black.reset(); white.reset()
    print("reference agents from", base)
