"""Drop-in through Gymnasium's own registry (needs gymnasium importable: baseline/_ref travels with the snapshot)."""
import os
import sys

import numpy as np
import pytest

from conftest import ROOT, have_cuda

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not have_cuda(), reason="needs a CUDA device")]

gym = pytest.importorskip("gymnasium")  # conftest.py puts baseline/_ref on sys.path when it exists


def test_make_vec_through_the_registry():
    import gymnasium_b200
    from gymnasium.vector import AutoresetMode, VectorEnv

    assert gymnasium_b200.HAVE_GYMNASIUM
    env = gym.make_vec("B200/CartPole-v1", num_envs=5, output="numpy")
    assert isinstance(env, VectorEnv) and isinstance(env, gymnasium_b200.B200VectorEnv)
    assert env.spec.id == "B200/CartPole-v1" and env.spec.kwargs["num_envs"] == 5
    assert env.max_episode_steps == 500 and env.metadata["autoreset_mode"] is AutoresetMode.NEXT_STEP
    ref = gym.make_vec("CartPole-v1", num_envs=5, vectorization_mode="sync")
    assert env.observation_space == ref.observation_space and env.action_space == ref.action_space
    assert env.single_observation_space == ref.single_observation_space
    o1, _ = env.reset(seed=123)
    o2, _ = ref.reset(seed=123)
    np.testing.assert_array_equal(o1, o2)
    rs = np.random.default_rng(0)
    for _ in range(150):
        a = ref.action_space.sample()
        x, y = env.step(a), ref.step(a)
        np.testing.assert_allclose(x[0], y[0], rtol=1e-5, atol=1e-5)
        for k in (1, 2, 3):
            np.testing.assert_array_equal(x[k], y[k])
            assert x[k].dtype == y[k].dtype
    env.close(); ref.close()


def test_install_routes_stock_ids():
    import gymnasium_b200

    gymnasium_b200.install(["FrozenLake-v1"])
    try:
        env = gym.make_vec("FrozenLake-v1", num_envs=6, map_name="8x8", output="numpy")
        assert isinstance(env, gymnasium_b200.B200VectorEnv) and env.max_episode_steps == 100
        ref = gym.make_vec("FrozenLake-v1", num_envs=6, map_name="8x8", vectorization_mode="sync")
        np.testing.assert_array_equal(env.reset(seed=5)[0], ref.reset(seed=5)[0])
        for _ in range(200):
            a = ref.action_space.sample()
            x, y = env.step(a), ref.step(a)
            for k in range(4):
                np.testing.assert_array_equal(x[k], y[k])
    finally:
        gymnasium_b200.uninstall()


def test_compat_mode_without_gymnasium():
    """The engine also runs on a box without the host framework (stand-in VectorEnv/spaces, gymnasium_b200/_compat.py)."""
    import subprocess

    code = (
        "import sys, numpy as np\n"
        "import gymnasium_b200\n"
        "assert not gymnasium_b200.HAVE_GYMNASIUM and 'gymnasium' not in sys.modules\n"
        "e = gymnasium_b200.make_vec('CartPole-v1', num_envs=3, output='numpy')\n"
        "o, _ = e.reset(seed=42)\n"
        "assert abs(float(o[0, 0]) - 0.0273956) < 1e-7, o\n"
        "o, r, te, tr, _ = e.step(np.array([1, 0, 1]))\n"
        "assert abs(float(o[0, 1]) - 0.18847767) < 1e-6 and r.dtype == np.float64\n"
        "assert e.action_space.shape == (3,) and e.observation_space.shape == (3, 4)\n"
        "print('compat ok')\n"
    )
    env = dict(os.environ, B200ENV_FORCE_COMPAT="1", PYTHONPATH=ROOT)
    p = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=300)
    assert p.returncode == 0 and "compat ok" in p.stdout, p.stderr[-2000:]
