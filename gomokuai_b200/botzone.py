"""Botzone wire format (row f4 of SURVEY.md section 8): `{"requests": [{x, y}, ...], "responses": [{x, y}, ...]}` in,
`{"response": {x, y}, "debug": ...}` out.

Mirrors BotzoneInterface / KeepAliveBotzoneInterface (reference core/interface/src/Interface.h:9-63) and the
encoder BotzoneAgent._parse_state (agents/botzone.py:27-35): requests are the opponent's moves (one more than the
responses; `{-1, -1}` as the first request means "you are black, move first"), responses are our earlier moves.
Moves are replayed without victory checks, an out-of-board move is a no-op (Board::applyMove, Game.cpp:37-47).
Only host logic lives here; choosing the move is the agent's job (e.g. MCTS with a GPU policy)."""
import json

WIDTH = HEIGHT = 15
KEEP_RUNNING = ">>>BOTZONE_REQUEST_KEEP_RUNNING<<<"


def _cell(move):
    x, y = int(move["x"]), int(move["y"])
    return y * WIDTH + x if 0 <= x < WIDTH and 0 <= y < HEIGHT else -1


def _xy(cell):
    return {"x": int(cell) % WIDTH, "y": int(cell) // WIDTH} if cell >= 0 else {"x": -1, "y": -1}


def moves_from_input(data):
    """Interface.h:16-22: interleave requests[i], responses[i] for i < len(responses), then the newest request.
    Returns the list of cell ids actually played (invalid / {-1,-1} moves dropped, as applyMove ignores them)."""
    if isinstance(data, str):
        data = json.loads(data)
    requests, responses = data["requests"], data.get("responses", [])
    if len(requests) != len(responses) + 1:
        raise ValueError("botzone input needs exactly one more request than responses")
    order = []
    for i, resp in enumerate(responses):
        order += [requests[i], resp]
    order.append(requests[len(responses)])
    played, seen = [], set()
    for mv in order:
        c = _cell(mv)
        if c >= 0 and c not in seen:                           # Board::checkMove: on the board and empty
            played.append(c)
            seen.add(c)
    return played


def input_from_moves(move_record):
    """BotzoneAgent._parse_state (agents/botzone.py:27-35): the JSON a bot receives for a game so far."""
    rec = [int(m) for m in move_record]
    offset = len(rec) % 2                                      # 0: black (first player) is to move
    padding = [{"x": -1, "y": -1}] * (1 - offset)
    return json.dumps({"requests": padding + [_xy(c) for c in rec[1 - offset::2]],
                       "responses": [_xy(c) for c in rec[offset::2]]})


def respond(text, choose_move, debug=""):
    """One BotzoneInterface turn (Interface.h:9-31): `choose_move(moves) -> cell`; returns the output JSON line."""
    moves = moves_from_input(text)
    return json.dumps({"response": _xy(choose_move(moves)), "debug": debug})


def mcts_chooser(policy=None, iterations=2000):
    """A chooser backed by this repo's CorePyExt mirror (host tree, GPU simulate slot)."""
    from . import core

    def choose(moves):
        board = core.Board()
        for c in moves:
            board.apply_move(c, False)
        agent = core.MCTS(c_iterations=iterations, policy=policy or core.TraditionalPolicy())
        agent.sync_with_board(board)
        return int(agent.get_action(board))
    return choose


def main():
    import sys
    print(respond(sys.stdin.read(), mcts_chooser()))


if __name__ == "__main__":
    main()
