"""Oracle: Blackjack-v1 (gymnasium/envs/toy_text/blackjack.py:10-238) behind SyncVectorEnv semantics.

Oracle only (see oracle/__init__.py).  Restates
  * ``draw_card`` / ``draw_hand``            blackjack.py:15-23  -- ``int(np_random.choice(deck))``, 13-card deck
  * ``usable_ace / sum_hand / is_bust / score / is_natural``  blackjack.py:26-45
  * ``BlackjackEnv.step``                    blackjack.py:178-208 (hit: bust -> -1; stick: dealer draws to >= 17, ``cmp`` of
    the scores, ``sab`` auto-win / ``natural`` 1.5 pay-out)
  * ``BlackjackEnv.reset``                   blackjack.py:215-238 (dealer hand, player hand, then the two draws that only
    pick the rendered suit / face of the dealer's top card -- they advance the RNG stream and are kept)
  * ``_get_obs``                             blackjack.py:210-213 -> (player sum, dealer's first card, usable ace)
The card draws go through numpy's own ``Generator.choice`` here (numpy is the reference's RNG dependency);
oracle/np_rng.py restates the algorithm behind it (buffered 32-bit Lemire) for the CUDA side.
Pinned by tests/golden/blackjack_*.npz (live reference).  Observations are returned stacked as ``(n, 3)`` int64.
"""
from __future__ import annotations

import numpy as np

from .vector import OracleVectorEnv

DECK = [1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 10, 10, 10]


def _sum_hand(hand):
    s = sum(hand)
    return s + 10 if (1 in hand and s + 10 <= 21) else s


def _score(hand):
    s = _sum_hand(hand)
    return 0 if s > 21 else s


def _natural(hand):
    return sorted(hand) == [1, 10]


class OracleBlackjack(OracleVectorEnv):
    def __init__(self, num_envs, natural=False, sab=True, max_episode_steps=None, autoreset_mode="NextStep"):
        super().__init__(num_envs, max_episode_steps, autoreset_mode)
        self.natural, self.sab = bool(natural), bool(sab)
        self.player = [[] for _ in range(num_envs)]
        self.dealer = [[] for _ in range(num_envs)]

    def _draw(self, i):
        return int(self._rng(i).choice(DECK))

    def _reset_env(self, i, options):
        self.dealer[i] = [self._draw(i), self._draw(i)]
        self.player[i] = [self._draw(i), self._draw(i)]
        self._rng(i).choice(["C", "D", "H", "S"])          # dealer_top_card_suit (render only, but it is a draw)
        if self.dealer[i][0] == 10:
            self._rng(i).choice(["J", "Q", "K"])           # dealer_top_card_value_str

    def _step_lanes(self, lanes, actions):
        reward = np.zeros(len(lanes), dtype=np.float64)
        term = np.zeros(len(lanes), dtype=bool)
        for k, (i, a) in enumerate(zip(lanes, actions)):
            i = int(i)
            if a:
                self.player[i].append(self._draw(i))
                if _sum_hand(self.player[i]) > 21:
                    term[k], reward[k] = True, -1.0
            else:
                term[k] = True
                while _sum_hand(self.dealer[i]) < 17:
                    self.dealer[i].append(self._draw(i))
                ps, ds = _score(self.player[i]), _score(self.dealer[i])
                r = float(ps > ds) - float(ps < ds)
                if self.sab and _natural(self.player[i]) and not _natural(self.dealer[i]):
                    r = 1.0
                elif not self.sab and self.natural and _natural(self.player[i]) and r == 1.0:
                    r = 1.5
                reward[k] = r
        return reward, term, {}

    def _obs(self):
        out = np.zeros((self.num_envs, 3), dtype=np.int64)
        for i in range(self.num_envs):
            if self.player[i]:
                s = sum(self.player[i])
                ace = int(1 in self.player[i] and s + 10 <= 21)
                out[i] = (s + 10 * ace, self.dealer[i][0], ace)
        return out
