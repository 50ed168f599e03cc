"""On-device vector wrappers (gymnasium_b200/wrappers.py) against the reference's NumPy wrappers, driven by the same
scripted vector env (CPU tensors: the wrappers are device-agnostic torch code)."""
import numpy as np
import pytest
import torch

gym = pytest.importorskip("gymnasium")
from gymnasium.vector import AutoresetMode, VectorEnv  # noqa: E402

from gymnasium_b200 import wrappers as W  # noqa: E402


class ScriptedEnv(VectorEnv):
    """Replays fixed (obs, reward, terminated, truncated) tapes with NEXT_STEP semantics; numpy or torch outputs."""

    def __init__(self, n, T, seed, as_torch, mode=AutoresetMode.NEXT_STEP):
        rs = np.random.default_rng(seed)
        self.num_envs = n
        self.metadata = {"autoreset_mode": mode}
        self.single_observation_space = gym.spaces.Box(-np.inf, np.inf, (3,), np.float32)
        self.single_action_space = gym.spaces.Discrete(2)
        self.observation_space = gym.vector.utils.batch_space(self.single_observation_space, n)
        self.action_space = gym.vector.utils.batch_space(self.single_action_space, n)
        self.obs = rs.normal(1.5, 2.0, size=(T + 1, n, 3)).astype(np.float32)
        self.rew = rs.normal(0.3, 1.0, size=(T, n))
        self.term = rs.random((T, n)) < 0.07
        self.trunc = rs.random((T, n)) < 0.04
        self.t, self.as_torch = 0, as_torch
        self.prev_done = np.zeros(n, dtype=bool)
        self.device = torch.device("cpu")

    def _cv(self, x):
        return torch.from_numpy(np.ascontiguousarray(x)) if self.as_torch else x

    def reset(self, *, seed=None, options=None):
        self.t = 0
        self.prev_done[:] = False
        return self._cv(self.obs[0]), {}

    def step(self, actions):
        t = self.t
        r, te, tr = self.rew[t].copy(), self.term[t].copy(), self.trunc[t].copy()
        if self.metadata["autoreset_mode"] == AutoresetMode.NEXT_STEP:
            r[self.prev_done] = 0.0; te[self.prev_done] = False; tr[self.prev_done] = False
        self.prev_done = te | tr
        self.t += 1
        return self._cv(self.obs[t + 1]), self._cv(r), self._cv(te), self._cv(tr), {}


@pytest.mark.parametrize("mode", [AutoresetMode.NEXT_STEP, AutoresetMode.SAME_STEP])
def test_record_episode_statistics_matches_reference(mode):
    from gymnasium.wrappers.vector import RecordEpisodeStatistics as Ref

    n, T = 6, 120
    ref = Ref(ScriptedEnv(n, T, 1, False, mode), buffer_length=50)
    mine = W.RecordEpisodeStatistics(ScriptedEnv(n, T, 1, True, mode), buffer_length=50)
    ref.reset(); mine.reset()
    for t in range(T):
        a = np.zeros(n, dtype=np.int64)
        _, _, te, tr, ri = ref.step(a)
        _, _, te2, tr2, mi = mine.step(a)
        done = te | tr
        np.testing.assert_array_equal(mi["_episode"].numpy(), done)
        if done.any():
            np.testing.assert_allclose(mi["episode"]["r"].numpy(), ri["episode"]["r"], rtol=0, atol=1e-12)
            np.testing.assert_array_equal(mi["episode"]["l"].numpy(), ri["episode"]["l"])
            np.testing.assert_array_equal(ri["_episode"], done)
        else:
            assert "episode" not in ri and float(mi["episode"]["r"].abs().sum()) == 0.0
    assert mine.episode_count == ref.episode_count
    np.testing.assert_allclose(np.sort(mine.return_queue.numpy()), np.sort(np.array(ref.return_queue)), atol=1e-12)
    np.testing.assert_array_equal(np.sort(mine.length_queue.numpy()), np.sort(np.array(ref.length_queue)))
    # a buffer smaller than the number of episodes (and than the finishers of one step): same entries, same (chronological)
    # order as the reference's deques
    ref = Ref(ScriptedEnv(n, T, 1, False, mode), buffer_length=4)
    mine = W.RecordEpisodeStatistics(ScriptedEnv(n, T, 1, True, mode), buffer_length=4)
    ref.reset(); mine.reset()
    for t in range(T):
        ref.step(np.zeros(n, dtype=np.int64)); mine.step(np.zeros(n, dtype=np.int64))
        np.testing.assert_allclose(mine.return_queue.numpy(), np.array(ref.return_queue), atol=1e-12, err_msg=f"step {t}")
        np.testing.assert_array_equal(mine.length_queue.numpy(), np.array(ref.length_queue))


def test_normalize_observation_and_reward_match_reference():
    from gymnasium.wrappers.vector import NormalizeObservation as RefO
    from gymnasium.wrappers.vector import NormalizeReward as RefR

    n, T = 8, 90
    ro, mo = RefO(ScriptedEnv(n, T, 2, False)), W.NormalizeObservation(ScriptedEnv(n, T, 2, True))
    rr, mr = RefR(ScriptedEnv(n, T, 3, False), gamma=0.97), W.NormalizeReward(ScriptedEnv(n, T, 3, True), gamma=0.97)
    o1, _ = ro.reset(); o2, _ = mo.reset()
    np.testing.assert_allclose(o2.numpy(), o1, rtol=1e-4, atol=1e-5)
    rr.reset(); mr.reset()
    for t in range(T):
        a = np.zeros(n, dtype=np.int64)
        o1, *_ = ro.step(a)
        o2, *_ = mo.step(a)
        assert o2.dtype == torch.float32
        np.testing.assert_allclose(o2.numpy(), o1, rtol=1e-4, atol=1e-5)
        _, r1, *_ = rr.step(a)
        _, r2, *_ = mr.step(a)
        # the reference takes np.mean/np.var of a float32 array (float32 pairwise sums); ours reduces in float64
        np.testing.assert_allclose(r2.numpy(), r1, rtol=2e-6, atol=1e-9)
    np.testing.assert_allclose(mo.obs_rms.mean.numpy(), ro.obs_rms.mean, rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(float(mr.return_rms.var), float(rr.return_rms.var), rtol=2e-6)
    with pytest.raises(ValueError):
        W.NormalizeObservation(ScriptedEnv(2, 3, 0, True), epsilon=0)
    with pytest.raises(ValueError):
        W.NormalizeReward(ScriptedEnv(2, 3, 0, True), gamma=1.5)


class InfoEnv(ScriptedEnv):
    """ScriptedEnv whose infos carry masked arrays and a nested dict, as SyncVectorEnv builds them (vector_env.py:277-338)."""

    def _info(self, t):
        n = self.num_envs
        rs = np.random.default_rng(1000 + t)
        m1, m2 = rs.random(n) < 0.5, rs.random(n) < 0.3
        info = {"k": rs.normal(size=n), "_k": m1, "always": np.arange(n, dtype=np.int64) + t,
                "nest": {"z": rs.normal(size=n).astype(np.float32), "_z": m2}, "_nest": m2}
        if self.as_torch:
            info = {"k": torch.from_numpy(info["k"]), "_k": torch.from_numpy(m1), "always": torch.from_numpy(info["always"]),
                    "nest": {"z": torch.from_numpy(info["nest"]["z"]), "_z": torch.from_numpy(m2)}, "_nest": torch.from_numpy(m2)}
        return info

    def reset(self, *, seed=None, options=None):
        o, _ = super().reset(seed=seed, options=options)
        return o, self._info(0)

    def step(self, actions):
        out = super().step(actions)
        return out[:4] + (self._info(self.t),)


def test_dict_info_to_list_matches_reference():
    from gymnasium.wrappers.vector import DictInfoToList as Ref

    n, T = 6, 12
    ref, ours = Ref(InfoEnv(n, T, 3, False)), W.DictInfoToList(InfoEnv(n, T, 3, True))
    (_, ia), (_, ib) = ref.reset(seed=0), ours.reset(seed=0)
    for t in range(T + 1):
        assert isinstance(ib, list) and len(ib) == n
        for a, b in zip(ia, ib):
            assert set(a) == set(b)
            for k in a:
                if isinstance(a[k], dict):
                    assert set(a[k]) == set(b[k])
                    for kk in a[k]:
                        assert a[k][kk] == b[k][kk] and np.asarray(b[k][kk]).dtype == np.asarray(a[k][kk]).dtype
                else:
                    assert a[k] == b[k] and np.asarray(b[k]).dtype == np.asarray(a[k]).dtype
        if t < T:
            ia, ib = ref.step(np.zeros(n, dtype=np.int64))[4], ours.step(np.zeros(n, dtype=np.int64))[4]


def test_numpy_to_torch_outputs_are_the_numpy_env_outputs_as_tensors():
    """The reference wrapper (numpy_to_torch.py:15-80) needs `array_api_compat`, which is not installed here; its
    contract is "same values, torch types" (tests/wrappers/vector/test_numpy_to_torch.py), checked against the raw env."""
    n, T = 5, 8
    raw, ours = InfoEnv(n, T, 4, False), W.NumpyToTorch(InfoEnv(n, T, 4, True))
    (oa, ia), (ob, ib) = raw.reset(seed=0), ours.reset(seed=0)
    assert isinstance(ob, torch.Tensor) and np.array_equal(oa, ob.numpy()) and np.array_equal(ia["k"], ib["k"].numpy())
    for t in range(T):
        a = raw.step(np.zeros(n, dtype=np.int64))
        b = ours.step(torch.zeros(n, dtype=torch.int64))
        for x, y in zip(a[:4], b[:4]):
            assert isinstance(y, torch.Tensor) and np.array_equal(x, y.numpy()) and y.numpy().dtype == x.dtype
        assert np.array_equal(a[4]["nest"]["z"], b[4]["nest"]["z"].numpy()) and np.array_equal(a[4]["_k"], b[4]["_k"].numpy())


def test_numpy_to_torch_switches_a_numpy_engine_env_to_torch_outputs():
    class Fake(ScriptedEnv):
        output = "numpy"

        @property
        def unwrapped(self):
            return self

    e = Fake(3, 4, 0, False)
    w = W.NumpyToTorch(e)
    assert e.output == "torch"
    assert isinstance(w.reset()[0], torch.Tensor)  # host arrays a non-engine env still returns are converted
