"""HalfCheetah-v5 oracle (oracle/half_cheetah.c -> mjc_planar.h).  Oracle only; PARITY UNPINNED against the MuJoCo wheel."""
from .mjc_planar import OraclePlanar

NB, NQ, NV, NU, OBS = 8, 9, 9, 6, 17
INFO_ROWS = {"x_position": 0, "x_velocity": 2, "reward_forward": 3, "reward_ctrl": 4}


class OracleHalfCheetah(OraclePlanar):
    robot = "half_cheetah"

    def __init__(self, num_envs, max_episode_steps=1000, reset_noise_scale=0.1):
        super().__init__(num_envs, max_episode_steps, reset_noise_scale)

    def _info_dict(self):  # half_cheetah_v5.py:232, :245-248
        return {k: self._info[:, r].copy() for k, r in INFO_ROWS.items()}
