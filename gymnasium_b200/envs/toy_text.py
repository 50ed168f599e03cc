"""CliffWalking-v1 / CliffWalkingSlippery-v1 and Taxi-v4 on the B200 engine (the generic tabular kernel).

Host-side table builders mirror ``CliffWalkingEnv.__init__/_calculate_transition_prob``
(gymnasium/envs/toy_text/cliffwalking.py:102-177) and ``TaxiEnv.__init__/_build_dry_transitions/_pickup/_dropoff``
(gymnasium/envs/toy_text/taxi.py:172-235, :299-371) and pack ``P[s][a]`` for ``csrc/frozenlake.cu`` (entry layout in
include/b200env.h).  Taxi: dry (the registered default), ``is_rainy=True`` and ``fickle_passenger=True`` (csrc/taxi.cu).
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from .. import _lib
from .._api import AutoresetMode
from ..vector_env import ptr
from .frozen_lake import TabularVectorEnv

# cliffwalking.py:11-20
UP, RIGHT, DOWN, LEFT = 0, 1, 2, 3
_DELTA = {UP: (-1, 0), RIGHT: (0, 1), DOWN: (1, 0), LEFT: (0, -1)}


def _entry(next_state: int, done: bool, reward_class: int, n_out: int) -> int:
    return next_state | (int(done) << 16) | (reward_class << 17) | (n_out << 20)


def pack_cliffwalking(is_slippery: bool):
    rows, cols = 4, 12
    nS, nA, start, goal = rows * cols, 4, 36, (3, 11)
    table = np.zeros((nS, nA, 3), dtype=np.uint32)
    for s in range(nS):
        r, c = divmod(s, cols)
        for move in range(nA):
            moves = [(move - 1) % 4, move, (move + 1) % 4] if is_slippery else [move]  # cliffwalking.py:159-164
            for k, a in enumerate(moves):
                nr = min(max(r + _DELTA[a][0], 0), rows - 1)
                nc = min(max(c + _DELTA[a][1], 0), cols - 1)
                if nr == 3 and 1 <= nc <= cols - 2:  # the cliff: back to the start, -100, episode continues (:170-171)
                    e = _entry(start, False, 1, len(moves))
                else:
                    e = _entry(nr * cols + nc, (nr, nc) == goal, 0, len(moves))
                table[s, move, k] = e
            if not is_slippery:
                table[s, move, 1:] = table[s, move, 0]
    p3 = np.array([1 / 3, 1 / 3, 1 / 3], dtype=np.float64)  # 1 / len(deltas)
    isd = np.zeros(nS)
    isd[start] = 1.0
    return table.reshape(-1), np.cumsum(p3), p3, np.cumsum(isd), nS, nA, (-1.0, -100.0, 0.0)


_TAXI_MAP = ["+---------+", "|R: | : :G|", "| : | : : |", "| : : : : |", "| | : | : |", "|Y| : |B: |", "+---------+"]  # taxi.py:15-23
_TAXI_LOCS = [(0, 0), (0, 4), (4, 0), (4, 3)]


def _taxi_encode(row, col, pass_loc, dest):  # taxi.py:373-382
    return ((row * 5 + col) * 5 + pass_loc) * 4 + dest


# rainy Taxi (taxi.py:246-307): a move that is possible goes as intended with p = rainy_probability and drifts to the
# heading's left / right cell with (1 - p) / 2 each; (forward, left, right) per action 0..3 = south, north, east, west
_RAINY_MOVES = {0: ((1, 0), (0, 1), (0, -1)), 1: ((-1, 0), (0, -1), (0, 1)), 2: ((0, 1), (-1, 0), (1, 0)),
                3: ((0, -1), (1, 0), (-1, 0))}


def _taxi_shift(desc, row, col, move):  # taxi.py:227-244: clamp to the grid, east/west moves stop at walls
    dr, dc = move
    nr, nc = max(0, min(row + dr, 4)), max(0, min(col + dc, 4))
    if dc == 1 and desc[1 + nr, 2 * nc] != b":":
        return row, col
    if dc == -1 and desc[1 + nr, 2 * nc + 2] != b":":
        return row, col
    return nr, nc


def pack_taxi(is_rainy: bool = False, rainy_probability: float = 0.8):
    """Non-fickle Taxi-v4: 500 states x 6 actions; reward classes (-1, -10, +20).  Dry: one outcome per (s, a).  Rainy:
    three outcomes (intended, left drift, right drift) for the four moves, one for pickup / dropoff."""
    desc = np.asarray(_TAXI_MAP, dtype="c")
    nS, nA = 500, 6
    table = np.zeros((nS, nA, 3), dtype=np.uint32)
    isd = np.zeros(nS)
    mask = np.zeros((nS, nA), dtype=np.int8)
    for row in range(5):
        for col in range(5):
            for p in range(5):
                for dest in range(4):
                    s = _taxi_encode(row, col, p, dest)
                    if p < 4 and p != dest:  # taxi.py:345-346
                        isd[s] += 1
                    here = (row, col)
                    east_open = desc[1 + row, 2 * col + 2] == b":"
                    west_open = desc[1 + row, 2 * col] == b":"
                    for a in range(nA):
                        nr, nc, np_, rc, term = row, col, p, 0, False
                        if a == 0:
                            nr = min(row + 1, 4)
                        elif a == 1:
                            nr = max(row - 1, 0)
                        if a == 2 and east_open:
                            nc = min(col + 1, 4)
                        elif a == 3 and west_open:
                            nc = max(col - 1, 0)
                        elif a == 4:  # pickup (taxi.py:179-188)
                            if p < 4 and here == _TAXI_LOCS[p]:
                                np_ = 4
                            else:
                                rc = 1
                        elif a == 5:  # dropoff (taxi.py:190-205)
                            if here == _TAXI_LOCS[dest] and p == 4:
                                np_, term, rc = dest, True, 2
                            elif here in _TAXI_LOCS and p == 4:
                                np_ = _TAXI_LOCS.index(here)
                            else:
                                rc = 1
                        if is_rainy and a <= 3:
                            can_move = (row < 4, row > 0, east_open, west_open)[a]
                            cells = [(nr, nc)] + ([_taxi_shift(desc, row, col, mv) for mv in _RAINY_MOVES[a][1:]]
                                                  if can_move else [here, here])
                            for k, (r2, c2) in enumerate(cells):
                                table[s, a, k] = _entry(_taxi_encode(r2, c2, p, dest), False, 0, 3)
                        else:
                            table[s, a, :] = _entry(_taxi_encode(nr, nc, np_, dest), term, rc, 1)
                    # action mask (taxi.py:398-419)
                    mask[s, 0] = row < 4
                    mask[s, 1] = row > 0
                    mask[s, 2] = col < 4 and east_open
                    mask[s, 3] = col > 0 and west_open
                    mask[s, 4] = p < 4 and here == _TAXI_LOCS[p]
                    mask[s, 5] = p == 4 and (here == _TAXI_LOCS[dest] or here in _TAXI_LOCS)
    isd /= isd.sum()
    lateral = (1.0 - rainy_probability) / 2.0
    p3 = np.array([rainy_probability, lateral, lateral]) if is_rainy else np.array([1.0, 1.0, 1.0])
    return table.reshape(-1), np.cumsum(p3), p3, np.cumsum(isd), nS, nA, (-1.0, -10.0, 20.0), mask


class CliffWalkingVectorEnv(TabularVectorEnv):
    """N CliffWalking-v1 envs (``is_slippery=True`` = CliffWalkingSlippery-v1).  No time limit by default, as registered
    (gymnasium/envs/__init__.py:156-166)."""

    def __init__(self, num_envs: int = 1, max_episode_steps: int | None = None, is_slippery: bool = False,
                 render_mode: str | None = None, **engine_kwargs):
        table, cum3, p3, isd_cum, nS, nA, rewards = pack_cliffwalking(bool(is_slippery))
        super().__init__(num_envs, nS, nA, table, cum3, p3, isd_cum, rewards, max_episode_steps=max_episode_steps,
                         render_mode=render_mode, **engine_kwargs)
        self.is_slippery = bool(is_slippery)
        self.shape = (4, 12)


class TaxiVectorEnv(TabularVectorEnv):
    """N Taxi-v4 envs; ``info`` carries ``prob`` and the per-state ``action_mask`` ``(N, 6) int8`` like the reference
    (taxi.py:457, :470)."""

    def __init__(self, num_envs: int = 1, max_episode_steps: int | None = 200, is_rainy: bool = False,
                 fickle_passenger: bool = False, rainy_probability: float = 0.8, fickle_probability: float = 0.3,
                 render_mode: str | None = None, **engine_kwargs):
        table, cum3, p3, isd_cum, nS, nA, rewards, mask = pack_taxi(bool(is_rainy), float(rainy_probability))
        self.is_rainy, self.fickle_passenger = bool(is_rainy), bool(fickle_passenger)
        self.fickle_probability = float(fickle_probability)
        super().__init__(num_envs, nS, nA, table, cum3, p3, isd_cum, rewards, max_episode_steps=max_episode_steps,
                         render_mode=render_mode, **engine_kwargs)
        self._mask_host = mask
        self._mask_dev = torch.from_numpy(mask).to(self.device)
        if self.fickle_passenger:  # taxi.py:436-452, :466-468 -> csrc/taxi.cu fix-up kernels around the tabular step
            if self.rng_mode != "numpy" or self.autoreset_mode == AutoresetMode.SAME_STEP:
                raise NotImplementedError("fickle_passenger=True needs rng='numpy' and NEXT_STEP or DISABLED autoreset")
            n, dev = self.num_envs, self.device
            self._fickle = torch.zeros(n, dtype=torch.uint8, device=dev)
            self._u32buf = torch.zeros(n, dtype=torch.int64, device=dev)
            self._prev_state = torch.zeros(n, dtype=torch.int32, device=dev)

    def _on_streams_seeded(self, lanes):
        if self.fickle_passenger:  # a freshly seeded numpy Generator starts with an empty 32-bit word buffer
            if lanes is None:
                self._u32buf.zero_()
            else:
                self._u32buf.masked_fill_(lanes, 0)

    def _reset_kernel(self, mask, options, out):
        super()._reset_kernel(mask, options, out)
        if self.fickle_passenger:
            _lib.check(
                self._lib.b2e_taxi_fickle_reset(C.byref(self._batch), self.fickle_probability,
                                                ptr(None if mask is None else mask.view(torch.uint8)), ptr(self._rng),
                                                ptr(self._fickle), self._stream),
                "b2e_taxi_fickle_reset",
            )

    @property
    def replayable_step(self) -> bool:
        """False when the step is more than C-ABI calls (distributed.HostBatchPipeline then keeps to the Python path)."""
        return not self.fickle_passenger

    def _step_kernel(self, actions, out):
        if self.fickle_passenger:
            self._prev_state.copy_(self._pstate)
        super()._step_kernel(actions, out)
        if self.fickle_passenger:
            _lib.check(
                self._lib.b2e_taxi_fickle_step(C.byref(self._batch), self.fickle_probability, ptr(self._prev_state),
                                               ptr(self._pstate), ptr(self._ctrl), ptr(self._rng), ptr(self._u32buf),
                                               ptr(self._fickle), ptr(out["obs"]), self._stream),
                "b2e_taxi_fickle_step",
            )

    def _with_mask(self, info, obs, valid):
        if isinstance(obs, np.ndarray):
            info["action_mask"] = self._mask_host[obs]
        else:
            info["action_mask"] = self._mask_dev[obs]
        info["_action_mask"] = valid
        return info

    def _reset_info(self, out, mask):
        info = super()._reset_info(out, mask)
        return self._with_mask(info, out["obs"], info["_prob"])

    def _step_info(self, out):
        info = super()._step_info(out)
        return self._with_mask(info, out["obs"], info["_prob"])
