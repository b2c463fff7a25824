"""Row f4: Botzone JSON wire format (reference core/interface/src/Interface.h:9-31, agents/botzone.py:27-41)."""
import json

from gomokuai_b200 import botzone


def test_first_turn_as_black_and_as_white():
    assert botzone.moves_from_input({"requests": [{"x": -1, "y": -1}], "responses": []}) == []          # we are black
    assert botzone.moves_from_input('{"requests": [{"x": 7, "y": 7}], "responses": []}') == [112]        # we are white


def test_interleaving_follows_the_reference_loop():
    data = {"requests": [{"x": -1, "y": -1}, {"x": 8, "y": 7}, {"x": 9, "y": 9}],
            "responses": [{"x": 7, "y": 7}, {"x": 7, "y": 8}]}
    # Interface.h:17-21: req0 (ignored), resp0, req1, resp1, req2
    assert botzone.moves_from_input(data) == [112, 113, 127, 144]


def test_round_trip_with_the_reference_encoder_layout():
    for rec in ([], [112], [112, 113], [112, 113, 97], [0, 224, 14, 210, 7]):
        text = botzone.input_from_moves(rec)
        data = json.loads(text)
        assert len(data["requests"]) == len(data["responses"]) + 1
        assert botzone.moves_from_input(text) == rec
    # agents/botzone.py:29-34 for a 3-move record: white to move -> requests = black's moves, responses = white's
    data = json.loads(botzone.input_from_moves([112, 113, 97]))
    assert data == {"requests": [{"x": 7, "y": 7}, {"x": 7, "y": 6}], "responses": [{"x": 8, "y": 7}]}


def test_respond_formats_position_as_the_reference_json():
    out = json.loads(botzone.respond(botzone.input_from_moves([112]), lambda moves: 113, debug="d"))
    assert out == {"response": {"x": 8, "y": 7}, "debug": "d"}       # Position -> {x, y} (Game.h to_json)
