"""The Game-API cases the reference's own env tests pin (envs/test_equivalence_game.py:39-383,
envs/test_equivalence_board.py:25-167), run through the CUDA-backed OthelloGameNew and checked
against the CPU oracle."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def games():
    import oracle as O
    from alphazero_othello_b200.envs.othello import OthelloGameNew
    return OthelloGameNew(8), O.OracleGame()


def _same(g, o, state, player):
    m = g.get_valid_moves(state, player)
    assert np.array_equal(m, o.get_valid_moves(state, player))
    assert g.get_value_and_terminated(state, None, player) == o.get_value_and_terminated(state, None, player)
    for a in np.nonzero(m)[0]:
        assert np.array_equal(g.get_next_state(state, int(a), player), o.get_next_state(state, int(a), player))
    return m


def test_square_piece_and_sizes(games):
    g, _ = games
    assert [g.get_square_piece(p) for p in (-1, 0, 1)] == ["X", "-", "O"]
    assert g.action_size == 65 and g.state_size == 64 and g.get_opponent(1) == -1 and g.get_opponent(-1) == 1


def test_all_first_moves_and_openings(games):
    g, o = games
    s = g.get_initial_state()
    assert list(np.nonzero(_same(g, o, s, 1))[0]) == [19, 26, 37, 44]
    assert list(np.nonzero(_same(g, o, s, -1))[0]) == [20, 29, 34, 43]
    for a in (19, 26, 37, 44):
        _same(g, o, g.get_next_state(s, a, 1), -1)


def test_pass_only_position_and_pass_returns_copy(games):
    g, o = games
    s = np.ones((8, 8), np.int8)
    s[4:, 4:] = 0           # three quadrants of one colour: nobody can flip anything
    for pl in (1, -1):
        m = _same(g, o, s, pl)
        assert m[64] == 1 and m[:64].sum() == 0
        nxt = g.get_next_state(s, 64, pl)
        assert nxt is not s and np.array_equal(nxt, s)
    assert g.get_value_and_terminated(s, 64, 1) == (1, True) and g.get_value_and_terminated(s, 64, -1) == (-1, True)


def test_illegal_moves_raise_valueerror(games):
    g, _ = games
    s = g.get_initial_state()
    for a in (0, 27, 28, 63, 18):   # empty without flips, occupied squares, empty diagonal neighbour
        with pytest.raises(ValueError, match=f"Illegal move: {a}"):
            g.get_next_state(s, a, 1)
    assert np.array_equal(s, g.get_initial_state())  # input never mutated


@pytest.mark.parametrize("corner,step", [((0, 7), (0, 1)), ((7, 7), (1, 0)), ((7, 0), (0, -1)), ((0, 0), (-1, 0)),
                                         ((3, 7), (1, 1)), ((4, 0), (-1, -1)), ((0, 3), (-1, 1)), ((7, 4), (1, -1))])
def test_no_wraparound_across_edges(games, corner, step):
    """A run that would continue past the edge must not re-enter on the other side (E/S/W/N and diagonals)."""
    g, o = games
    s = np.zeros((8, 8), np.int8)
    r, c = corner
    s[r, c] = -1
    pr, pc = r - step[0], c - step[1]
    if 0 <= pr < 8 and 0 <= pc < 8:
        s[pr, pc] = 1
    # the square "behind" the edge in flattened order would be legal if shifts wrapped
    for pl in (1, -1):
        _same(g, o, s, pl)


def test_double_pass_and_terminal_value_cases(games):
    g, o = games
    empty = np.zeros((8, 8), np.int8)
    for pl in (1, -1):
        assert g.get_value_and_terminated(empty, None, pl) == (0, True)  # no discs, nobody moves: draw
        assert g.get_valid_moves(empty, pl)[64] == 1
    full = np.ones((8, 8), np.int8)
    full[:4] = -1
    assert g.get_value_and_terminated(full, None, 1) == (0, True)
    full[3, 0] = 1
    assert g.get_value_and_terminated(full, None, 1) == (1, True) and g.get_value_and_terminated(full, None, -1) == (-1, True)
    s = np.zeros((8, 8), np.int8)  # empty squares are not awarded to anyone (envs/othello.py:447-454)
    s[0, 0] = 1
    assert g.get_value_and_terminated(s, None, 1) == (1, True) and g.get_score(s, 1) == 1 and g.get_score(s, -1) == -1


def test_score_consistency_random_endgames(games):
    g, o = games
    rs = np.random.RandomState(8)
    boards = rs.choice([-1, 1], size=(64, 8, 8)).astype(np.int8)
    for b in boards[:16]:
        assert g.get_score(b, 1) == o.get_score(b, 1) == -g.get_score(b, -1)
    v1, t1 = g.value_and_terminated_batch(boards, np.ones(64, np.int8))
    assert t1.all() and np.array_equal(v1, np.sign(boards.reshape(64, -1).sum(1)))


def test_full_lowest_index_game_matches_survey_known_answer(games):
    """always the smallest legal action; 64 actions, 4 passes, final score -26."""
    g, _ = games
    s, pl, acts = g.get_initial_state(), 1, []
    while True:
        a = int(np.nonzero(g.get_valid_moves(s, pl))[0][0])
        acts.append(a)
        s = g.get_next_state(s, a, pl)
        if g.get_value_and_terminated(s, a, pl)[1]:
            break
        pl = -pl
    assert len(acts) == 64 and acts.count(64) == 4 and acts[:8] == [19, 18, 17, 9, 1, 0, 26, 2] and g.get_score(s, 1) == -26
