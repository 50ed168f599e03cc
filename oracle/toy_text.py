"""Oracle: CliffWalking-v1 and Taxi-v4 (default dry / non-fickle variant) tabular dynamics, bit-exact integer state.

Oracle only (see oracle/__init__.py).  Restates
  * ``CliffWalkingEnv.__init__/_calculate_transition_prob``  gymnasium/envs/toy_text/cliffwalking.py:102-177
    (cliff cells send the agent back to the start with -100, not terminated; slippery variant: three outcomes of
    probability 1/3 for ``[(move-1)%4, move, (move+1)%4]``)
  * ``TaxiEnv.__init__/_build_dry_transitions/_pickup/_dropoff/encode/action_mask``
    gymnasium/envs/toy_text/taxi.py:172-235, :299-371, :373-430 (``is_rainy=False, fickle_passenger=False``)
  * step / reset of both: one ``categorical_sample`` draw per call (cliffwalking.py:179-203, taxi.py:432-470)
Pinned by tests/golden/cliffwalking_*.npz and taxi_*.npz (live reference).
"""
from __future__ import annotations

import numpy as np

from .frozenlake import categorical_sample
from .vector import OracleVectorEnv

UP, RIGHT, DOWN, LEFT = 0, 1, 2, 3
POSITION_MAPPING = {UP: [-1, 0], RIGHT: [0, 1], DOWN: [1, 0], LEFT: [0, -1]}  # cliffwalking.py:15-20


def build_cliff(is_slippery=False):
    shape = (4, 12)
    nS, nA = 48, 4
    start = 36
    cliff = np.zeros(shape, dtype=bool)
    cliff[3, 1:-1] = True
    P = {}
    for s in range(nS):
        pos = np.unravel_index(s, shape)
        P[s] = {}
        for move in (UP, RIGHT, DOWN, LEFT):
            deltas = [POSITION_MAPPING[move]] if not is_slippery else [
                POSITION_MAPPING[a] for a in [(move - 1) % 4, move, (move + 1) % 4]]
            out = []
            for delta in deltas:
                new = [min(max(pos[0] + delta[0], 0), shape[0] - 1), min(max(pos[1] + delta[1], 0), shape[1] - 1)]
                ns = new[0] * shape[1] + new[1]
                if cliff[tuple(new)]:
                    out.append((1 / len(deltas), start, -100, False))
                else:
                    out.append((1 / len(deltas), ns, -1, tuple(new) == (shape[0] - 1, shape[1] - 1)))
            P[s][move] = out
    isd = np.zeros(nS)
    isd[start] = 1.0
    return P, isd


TAXI_MAP = ["+---------+", "|R: | : :G|", "| : | : : |", "| : : : : |", "| | : | : |", "|Y| : |B: |", "+---------+"]
TAXI_LOCS = [(0, 0), (0, 4), (4, 0), (4, 3)]


def taxi_encode(row, col, pass_loc, dest):
    return ((row * 5 + col) * 5 + pass_loc) * 4 + dest


def taxi_decode(i):
    dest = i % 4; i //= 4
    p = i % 5; i //= 5
    col = i % 5; i //= 5
    return i, col, p, dest


_RAINY = {0: ((0, 1), (0, -1)), 1: ((0, -1), (0, 1)), 2: ((-1, 0), (1, 0)), 3: ((1, 0), (-1, 0))}  # (left, right) drift per heading


def _shift(desc, row, col, dr, dc):
    """``TaxiEnv._calc_new_position`` (taxi.py:227-244)."""
    nr, nc = max(0, min(row + dr, 4)), max(0, min(col + dc, 4))
    if dc == 1 and desc[1 + nr, 2 * nc] != b":":
        return row, col
    if dc == -1 and desc[1 + nr, 2 * nc + 2] != b":":
        return row, col
    return nr, nc


def build_taxi(is_rainy=False, rainy_probability=0.8):
    """``_build_dry_transitions`` (taxi.py:207-225) or ``_build_rainy_transitions`` (taxi.py:246-307)."""
    lateral = (1.0 - rainy_probability) / 2.0
    desc = np.asarray(TAXI_MAP, dtype="c")
    locs = TAXI_LOCS
    P = {s: {a: [] for a in range(6)} for s in range(500)}
    isd = np.zeros(500)
    for row in range(5):
        for col in range(5):
            for pass_idx in range(5):
                for dest_idx in range(4):
                    state = taxi_encode(row, col, pass_idx, dest_idx)
                    if pass_idx < 4 and pass_idx != dest_idx:
                        isd[state] += 1
                    for action in range(6):
                        taxi_loc = (row, col)
                        nr, nc, npass = row, col, pass_idx
                        reward, term = -1, False
                        if action == 0:
                            nr = min(row + 1, 4)
                        elif action == 1:
                            nr = max(row - 1, 0)
                        if action == 2 and desc[1 + row, 2 * col + 2] == b":":
                            nc = min(col + 1, 4)
                        elif action == 3 and desc[1 + row, 2 * col] == b":":
                            nc = max(col - 1, 0)
                        elif action == 4:
                            if pass_idx < 4 and taxi_loc == locs[pass_idx]:
                                npass = 4
                            else:
                                reward = -10
                        elif action == 5:
                            if taxi_loc == locs[dest_idx] and pass_idx == 4:
                                npass, term, reward = dest_idx, True, 20
                            elif taxi_loc in locs and pass_idx == 4:
                                npass = locs.index(taxi_loc)
                            else:
                                reward = -10
                        if is_rainy and action <= 3:
                            ok = (row < 4, row > 0, desc[1 + row, 2 * col + 2] == b":", desc[1 + row, 2 * col] == b":")[action]
                            left, right = [(_shift(desc, row, col, *mv) if ok else taxi_loc) for mv in _RAINY[action]]
                            P[state][action].append((rainy_probability, taxi_encode(nr, nc, pass_idx, dest_idx), -1, False))
                            P[state][action].append((lateral, taxi_encode(left[0], left[1], pass_idx, dest_idx), -1, False))
                            P[state][action].append((lateral, taxi_encode(right[0], right[1], pass_idx, dest_idx), -1, False))
                        else:
                            P[state][action].append((1.0, taxi_encode(nr, nc, npass, dest_idx), reward, term))
    isd /= isd.sum()
    return P, isd


def taxi_action_mask(state):
    desc = np.asarray(TAXI_MAP, dtype="c")
    row, col, p, dest = taxi_decode(state)
    m = np.zeros(6, dtype=np.int8)
    if row < 4: m[0] = 1
    if row > 0: m[1] = 1
    if col < 4 and desc[row + 1, 2 * col + 2] == b":": m[2] = 1
    if col > 0 and desc[row + 1, 2 * col] == b":": m[3] = 1
    if p < 4 and (row, col) == TAXI_LOCS[p]: m[4] = 1
    if p == 4 and ((row, col) == TAXI_LOCS[dest] or (row, col) in TAXI_LOCS): m[5] = 1
    return m


class OracleTabular(OracleVectorEnv):
    def __init__(self, num_envs, P, isd, max_episode_steps=None, autoreset_mode="NextStep"):
        super().__init__(num_envs, max_episode_steps, autoreset_mode)
        self.P, self.isd = P, isd
        self.s = np.zeros(num_envs, dtype=np.int64)

    def _reset_env(self, i, options):
        self.s[i] = categorical_sample(self.isd, self._rng(i))

    def _reset_info(self, lanes):
        return {"prob": np.ones(self.num_envs, dtype=np.float64)}

    def _step_lanes(self, lanes, actions):
        reward = np.zeros(len(lanes)); term = np.zeros(len(lanes), dtype=bool); prob = np.zeros(len(lanes))
        for k, (i, a) in enumerate(zip(lanes, actions)):
            tr = self.P[int(self.s[i])][int(a)]
            j = categorical_sample([t[0] for t in tr], self._rng(int(i)))
            p, s2, r, d = tr[j]
            self.s[i] = s2
            reward[k], term[k], prob[k] = r, d, p
        return reward, term, {"prob": prob}

    def _obs(self):
        return self.s.copy()


class OracleCliffWalking(OracleTabular):
    def __init__(self, num_envs, is_slippery=False, max_episode_steps=None, autoreset_mode="NextStep"):
        P, isd = build_cliff(is_slippery)
        super().__init__(num_envs, P, isd, max_episode_steps, autoreset_mode)


class OracleTaxi(OracleTabular):
    """``TaxiEnv`` (taxi.py:309-487).  ``fickle_passenger`` (taxi.py:436-452, :466-468): a passenger who rides changes
    his mind once per episode with probability ``fickle_probability`` -- decided by one extra ``random()`` draw at reset
    -- on the first step in which the taxi moves with him on board; the new destination is
    ``np_random.choice`` of the three other locations."""

    def __init__(self, num_envs, max_episode_steps=200, autoreset_mode="NextStep", is_rainy=False, fickle_passenger=False,
                 fickle_probability=0.3):
        P, isd = build_taxi(is_rainy)
        super().__init__(num_envs, P, isd, max_episode_steps, autoreset_mode)
        self.fickle_passenger, self.fickle_probability = bool(fickle_passenger), fickle_probability
        self.fickle_step = np.zeros(num_envs, dtype=bool)

    def _reset_env(self, i, options):
        super()._reset_env(i, options)
        self.fickle_step[i] = self.fickle_passenger and self._rng(i).random() < self.fickle_probability

    def _step_lanes(self, lanes, actions):
        before = self.s.copy()
        out = super()._step_lanes(lanes, actions)
        if self.fickle_passenger:
            for i in lanes:
                i = int(i)
                r0, c0, p0, d0 = taxi_decode(int(before[i]))
                r1, c1, p1, _ = taxi_decode(int(self.s[i]))
                if self.fickle_step[i] and p0 == 4 and (r1 != r0 or c1 != c0):
                    self.fickle_step[i] = False
                    dest = int(self._rng(i).choice([k for k in range(4) if k != d0]))
                    self.s[i] = taxi_encode(r1, c1, p1, dest)
        return out

    def action_mask(self):
        return np.stack([taxi_action_mask(int(s)) for s in self.s])
