"""Humanoid-v5 oracle (oracle/humanoid.c, MuJoCo-subset restatement; PARITY UNPINNED -- mujoco is not installable
here).  Pins that do not need the wheel: the reference's own structural/behavioural tests for this boundary
(tests/envs/mujoco/test_mujoco_v5.py) and well-known constants of humanoid.xml."""
import numpy as np
import pytest

from oracle.humanoid import OracleHumanoid


def test_model_constants():
    env = OracleHumanoid(1)
    mass, misc, inv = env.model_info()
    # body masses MuJoCo derives from humanoid.xml's geoms at density 1000 (well-known mjModel.body_mass of this model)
    known = [0, 8.90746237, 2.26194671, 6.61619413, 4.75175093, 2.75569617, 1.76714587, 4.75175093, 2.75569617,
             1.76714587, 1.66108048, 1.22954019, 1.66108048, 1.22954019]
    np.testing.assert_allclose(mass, known, rtol=2e-8)
    assert misc[1] > 100  # geom pairs that may collide
    assert (inv[2:28] > 0).all() and (inv[28:] > 0).all()


def test_reset_noise_zero_gives_init_state():
    """tests/envs/mujoco/test_mujoco_v5.py:693-710."""
    env = OracleHumanoid(2, reset_noise_scale=0.0)
    obs, info = env.reset(seed=0)
    qpos, qvel, _, counts = env.debug(0)
    np.testing.assert_array_equal(qpos, [0, 0, 1.4, 1, 0, 0, 0] + [0] * 17)
    np.testing.assert_array_equal(qvel, np.zeros(23))
    assert obs.shape == (2, 348) and obs.dtype == np.float64  # :645-649 / humanoid_v5.py:395-397
    np.testing.assert_array_equal(obs[0], obs[1])
    assert info["x_position"][0] == 0 and info["distance_from_origin"][0] == 0
    # cfrc_ext and qfrc_actuator are zero right after reset (mj_resetData + mj_forward with ctrl = 0)
    assert (obs[0, 22 + 23 + 130 + 78:] == 0).all()
    # cinert block: last entry of each body's 10 numbers is its mass
    np.testing.assert_allclose(obs[0, 45:175].reshape(13, 10)[:, 9], env.model_info()[0][1:], rtol=1e-12)


def test_reward_identities_and_velocity():
    """reward == sum of info terms (:221-231); x_velocity from the mass centre (:132-151); info x == qpos[0] (:32-57)."""
    env = OracleHumanoid(3)
    env.reset(seed=7)
    rs = np.random.default_rng(0)
    prev_done = np.zeros(3, dtype=bool)
    for _ in range(30):
        a = rs.uniform(-0.4, 0.4, size=(3, 17)).astype(np.float32)
        obs, r, te, tr, info = env.step(a)
        live = ~prev_done  # lanes on their autoreset call return the reset info instead
        total = info["reward_survive"] + info["reward_forward"] + info["reward_ctrl"] + info["reward_contact"]
        np.testing.assert_allclose(r[live], total[live], rtol=0, atol=1e-12)
        np.testing.assert_allclose(info["reward_forward"][live], 1.25 * info["x_velocity"][live], rtol=0, atol=1e-15)
        np.testing.assert_allclose(info["reward_ctrl"][live], -0.1 * (a.astype(np.float64) ** 2).sum(1)[live], rtol=1e-12)
        for i in range(3):
            qpos, *_ = env.debug(i)
            assert info["x_position"][i] == qpos[0] and obs[i, 0] == qpos[2]
            assert info["distance_from_origin"][i] == np.sqrt(qpos[0] ** 2 + qpos[1] ** 2)
        prev_done = te | tr


def test_verify_reward_survive_behaviour():
    """tests/envs/mujoco/test_mujoco_v5.py:159-191: noise-free reset, random actions => leaves z in (1, 2) within 80
    steps; reward_survive is 0 on the terminal step and 5 before."""
    env = OracleHumanoid(1, reset_noise_scale=0.0)
    env.reset(seed=0)
    rs = np.random.default_rng(2)
    for step in range(80):
        obs, r, te, tr, info = env.step(rs.uniform(-0.4, 0.4, size=(1, 17)).astype(np.float32))
        if te[0]:
            assert info["reward_survive"][0] == 0
            z = obs[0, 0]
            assert not (1.0 < z < 2.0)
            break
        assert info["reward_survive"][0] == 5.0
        assert np.isfinite(obs).all()
    else:
        raise AssertionError("Humanoid did not terminate within 80 random steps")


def test_out_of_bound_actions_are_clamped_for_dynamics():
    """tests/envs/test_action_dim_check.py:88-147: an out-of-bound Box action gives the same observation as the clipped
    action (MuJoCo clamps ctrl), while the control cost uses the raw value."""
    a, b = OracleHumanoid(1), OracleHumanoid(1)
    a.reset(seed=3); b.reset(seed=3)
    big = np.full((1, 17), 0.9, dtype=np.float32)
    oa, ra, *_ = a.step(big)
    ob, rb, *_ = b.step(np.clip(big, -0.4, 0.4))
    np.testing.assert_array_equal(oa, ob)
    assert ra[0] < rb[0]
    with pytest.raises(ValueError):
        a.step(np.zeros((1, 16), dtype=np.float32))


def test_determinism_autoreset_and_stability():
    a, b = OracleHumanoid(4), OracleHumanoid(4)
    oa, _ = a.reset(seed=11); ob, _ = b.reset(seed=11)
    np.testing.assert_array_equal(oa, ob)
    rs = np.random.default_rng(5)
    prev_done = np.zeros(4, dtype=bool)
    n_done = 0
    for t in range(150):
        act = rs.uniform(-0.4, 0.4, size=(4, 17)).astype(np.float32)
        xa, xb = a.step(act), b.step(act)
        for k in range(4):
            np.testing.assert_array_equal(xa[k], xb[k])
        assert np.isfinite(xa[0]).all() and np.abs(xa[0][:, :45]).max() < 100
        assert (xa[1][prev_done] == 0).all() and not xa[2][prev_done].any()  # the call after a done is the reset
        assert (np.abs(xa[0][prev_done, 0] - 1.4) < 0.011).all()
        prev_done = xa[2] | xa[3]
        n_done += prev_done.sum()
    assert n_done >= 4
