"""InvertedPendulum-v5 oracle (oracle/inverted_pendulum.c -> mjc_planar.h).  Oracle only; PARITY UNPINNED against the MuJoCo
wheel, but pinned analytically: tests/test_oracle_hopper.py integrates the cart-pole equations of motion independently."""
from .mjc_planar import OraclePlanar

NB, NQ, NV, NU, OBS = 3, 2, 2, 1, 4


class OracleInvertedPendulum(OraclePlanar):
    robot = "inverted_pendulum"

    def __init__(self, num_envs, max_episode_steps=1000, reset_noise_scale=0.01):
        super().__init__(num_envs, max_episode_steps, reset_noise_scale)

    def _info_dict(self):  # inverted_pendulum_v5.py:160: {"reward_survive": reward}; reset: {}
        return {"reward_survive": self._info[:, 5].copy()}
