"""Per-game facts mirrored from csrc/game_db.h (what ALE's getMinimalActionSet reports): used where the
reference asks for `num_actions` before any emulator exists (environment_creator.py:27-29)."""

MINIMAL_ACTIONS = {
    "pong": [0, 1, 3, 4, 11, 12],
    "breakout": [0, 1, 3, 4],
    "seaquest": list(range(18)),
    "space_invaders": [0, 1, 3, 4, 11, 12],
    "ms_pacman": [0, 2, 3, 4, 5, 6, 7, 8, 9],
    "asterix": [0, 2, 3, 4, 5, 6, 7, 8, 9],
    "asteroids": [0, 1, 2, 3, 4, 5, 6, 7, 10, 11, 12, 13, 14, 15],
    "enduro": [0, 1, 3, 4, 5, 8, 9, 11, 12],
    "gopher": [0, 1, 2, 3, 4, 10, 11, 12],
    "gravitar": list(range(18)),
    "montezuma_revenge": list(range(18)),
    "yars_revenge": list(range(18)),
}


def minimal_action_count(game):
    return len(MINIMAL_ACTIONS.get(game, list(range(18))))
