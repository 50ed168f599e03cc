"""Oracle: FrozenLake-v1 tabular dynamics, bit-exact integer state.

Oracle only (see oracle/__init__.py).  Restates
  * map ``MAPS``                 gymnasium/envs/toy_text/frozen_lake.py:14-31 (re-typed here as data)
  * ``FrozenLakeEnv.__init__``   frozen_lake.py:232-305 -- P[s][a] = [(p, s', r, done)...]; terminal tiles
    self-loop :290-291; slippery outcomes ``[(a-1)%4, a, (a+1)%4]`` with ``[fail, success, fail]`` :293-300
  * ``FrozenLakeEnv.step``       frozen_lake.py:324-334 + ``categorical_sample`` toy_text/utils.py:4-8
    (``argmax(cumsum(p) > random())``, one draw every step)
  * ``FrozenLakeEnv.reset``      frozen_lake.py:336-348 (one draw; info prob = 1)
Deviation kept on purpose: ``prob`` is always float64 (1.0 on reset calls).
Pinned by tests/golden/frozenlake_*.npz (live reference).
"""
from __future__ import annotations

import numpy as np

from .vector import OracleVectorEnv

LEFT, DOWN, RIGHT, UP = 0, 1, 2, 3
MAPS = {
    "4x4": ["SFFF", "FHFH", "FFFH", "HFFG"],
    "8x8": ["SFFFFFFF", "FFFFFFFF", "FFFHFFFF", "FFFFFHFF", "FFFHFFFF", "FHHFFFHF", "FHFFHFHF", "FFFHFFFG"],
}


def build_table(desc, is_slippery=True, success_rate=1.0 / 3.0, reward_schedule=(1, 0, 0)):
    """Returns (P, initial_state_distrib) with P[s][a] = list of (p, s', r, done)."""
    desc = [list(r) for r in desc]
    nrow, ncol = len(desc), len(desc[0])
    fail_rate = (1.0 - success_rate) / 2.0
    isd = np.array([[c == "S" for c in r] for r in desc], dtype=np.float64).ravel()
    isd /= isd.sum()

    def inc(row, col, a):
        if a == LEFT:
            col = max(col - 1, 0)
        elif a == DOWN:
            row = min(row + 1, nrow - 1)
        elif a == RIGHT:
            col = min(col + 1, ncol - 1)
        elif a == UP:
            row = max(row - 1, 0)
        return row, col

    def upd(row, col, a):
        nr, nc = inc(row, col, a)
        letter = desc[nr][nc]
        done = letter in "GH"
        reward = reward_schedule["GHF".index(letter if letter in "GHF" else "F")]
        return nr * ncol + nc, reward, done

    P = {}
    for row in range(nrow):
        for col in range(ncol):
            s = row * ncol + col
            P[s] = {}
            for a in range(4):
                li = []
                if desc[row][col] in "GH":
                    li.append((1.0, s, 0, True))
                elif is_slippery:
                    for b in [(a - 1) % 4, a, (a + 1) % 4]:
                        li.append((success_rate if b == a else fail_rate, *upd(row, col, b)))
                else:
                    li.append((1.0, *upd(row, col, a)))
                P[s][a] = li
    return P, isd


def categorical_sample(prob_n, rng):
    csprob_n = np.cumsum(np.asarray(prob_n))
    return int(np.argmax(csprob_n > rng.random()))


class OracleFrozenLake(OracleVectorEnv):
    def __init__(self, num_envs, map_name="8x8", desc=None, is_slippery=True, success_rate=1.0 / 3.0,
                 reward_schedule=(1, 0, 0), max_episode_steps=100, autoreset_mode="NextStep"):
        super().__init__(num_envs, max_episode_steps, autoreset_mode)
        self.desc = MAPS[map_name] if desc is None else desc
        self.P, self.isd = build_table(self.desc, is_slippery, success_rate, reward_schedule)
        self.s = np.zeros(num_envs, dtype=np.int64)

    def _reset_env(self, i, options):
        self.s[i] = categorical_sample(self.isd, self._rng(i))

    def _reset_info(self, lanes):
        return {"prob": np.ones(self.num_envs, dtype=np.float64)}

    def _step_lanes(self, lanes, actions):
        reward = np.zeros(len(lanes), dtype=np.float64)
        term = np.zeros(len(lanes), dtype=bool)
        prob = np.zeros(len(lanes), dtype=np.float64)
        for k, (i, a) in enumerate(zip(lanes, actions)):
            tr = self.P[int(self.s[i])][int(a)]
            j = categorical_sample([t[0] for t in tr], self._rng(int(i)))
            p, s2, r, d = tr[j]
            self.s[i] = s2
            reward[k], term[k], prob[k] = r, d, p
        return reward, term, {"prob": prob}

    def _obs(self):
        return self.s.copy()
