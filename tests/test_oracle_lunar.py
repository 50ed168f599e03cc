"""LunarLander-v3 oracle (oracle/lunar_lander.c, Box2D-subset restatement; PARITY UNPINNED -- Box2D is not installable
here).  What CAN be pinned: the reference's own behavioural test for this env, structural facts of the model, and
self-consistency."""
import numpy as np
import pytest

from oracle.lunar_lander import OracleLunarLander, heuristic, toi_probe


def run_heuristic(seed, max_steps=1000):
    env = OracleLunarLander(1)
    obs, _ = env.reset(seed=seed)
    total, steps = 0.0, 0
    while True:
        obs, r, te, tr, _ = env.step([heuristic(obs[0])])
        total += r[0]
        steps += 1
        if te[0] or tr[0] or steps >= max_steps:
            return total, steps, obs[0], env


def test_heuristic_lands_like_the_reference_test():
    """tests/envs/test_env_implementation.py:12-16: demo_heuristic_lander(env, seed=1) must score > 100."""
    total, steps, obs, env = run_heuristic(1)
    assert total > 100, total
    assert obs[6] == 1.0 and obs[7] == 1.0  # both legs on the ground
    bodies, _ = env.debug_state()
    assert (bodies[:, 6] == 0).all()  # the island went to sleep -> +100 (lunar_lander.py:658-660)


@pytest.mark.parametrize("seed", [0, 2, 3, 7, 11, 42])
def test_heuristic_usually_lands(seed):
    total, steps, obs, _ = run_heuristic(seed)
    assert total > 0 and steps < 1000, (seed, total, steps)


def test_model_constants():
    env = OracleLunarLander(2)
    obs, _ = env.reset(seed=5)
    bodies, misc = env.debug_state(0)
    # lander mass ~ 867/900 m^2 * 5 = 4.817; leg mass (4/30 * 16/30) * 1 = 0.0711
    assert abs(misc[0] - 4.81666) < 1e-4 and abs(misc[2] - 0.071111) < 1e-6
    assert misc[4] == 0.0 and misc[5] > 0.05  # centroid above the origin on the symmetry axis
    terr = env.terrain(0)
    np.testing.assert_allclose(terr[1:, 0], np.arange(10) * 2.0, atol=1e-6)  # chunk_x = W/10 * i
    helipad_y = np.float32(400 / 30.0 / 4)
    # smooth_y of chunks 4..6 averages three helipad heights: 0.33 * 3 * H/4
    np.testing.assert_allclose(terr[5:7, 1], 0.33 * 3 * helipad_y, rtol=1e-6)
    assert obs.dtype == np.float32 and obs.shape == (2, 8)
    # lunar_lander.py:447: reset returns the observation after one step(0): x ~ 0, y ~ 1.4, no leg contact
    assert abs(obs[0, 0]) < 0.05 and 1.3 < obs[0, 1] < 1.5 and obs[0, 6] == 0 and obs[0, 7] == 0


def test_determinism_and_autoreset():
    a, b = OracleLunarLander(4), OracleLunarLander(4)
    oa, _ = a.reset(seed=9)
    ob, _ = b.reset(seed=9)
    np.testing.assert_array_equal(oa, ob)
    rs = np.random.default_rng(0)
    n_term = 0
    for t in range(400):
        act = rs.integers(0, 4, 4)
        xa, xb = a.step(act), b.step(act)
        for k in range(4):
            np.testing.assert_array_equal(xa[k], xb[k])
        after = xa[2] | xa[3]
        n_term += after.sum()
        if t > 0:
            # the call after a done is the reset: reward 0, flags False, fresh observation near the top
            assert (xa[1][prev_done] == 0).all() and not xa[2][prev_done].any()
            assert (xa[0][prev_done, 1] > 1.3).all()
        prev_done = after
    assert n_term >= 4  # random policies crash within ~100 steps
    # random-policy crashes give -100 on the terminating step
    # different seeds give different terrain/impulses
    c = OracleLunarLander(1)
    oc, _ = c.reset(seed=10)
    assert not np.array_equal(oc[0], oa[0])


def test_time_of_impact_known_answers():
    """b2TimeOfImpact on cases with a closed form.  Both shapes carry b2_polygonRadius (0.01), so the routine reports
    'touching' when the distance between the cores reaches target = max(linearSlop, 0.02 - 3 linearSlop) = 0.005 (within
    0.25 linearSlop).  A box translating straight down onto a horizontal edge: gap(t) = gap0 - drop * t."""
    target, tol = 0.005, 0.00125
    hx, hy = 2 / 30.0, 8 / 30.0  # a leg (lunar_lander.py:413)
    edge = [0.0, 1.0, 4.0, 1.0]
    for y0, y1 in [(2.0, 1.0), (1.5, 1.2), (1.30, 1.25), (3.0, 0.0)]:
        state, t = toi_probe(edge, (hx, hy), (2.0, y0), 0.0, (2.0, y1), 0.0)
        gap0, drop = (y0 - hy) - 1.0, y0 - y1
        assert state == 3, (y0, y1, state)
        assert abs((gap0 - drop * t) - target) <= tol * 1.01, (y0, y1, t)
    # never comes close: separated at t = 1
    assert toi_probe(edge, (hx, hy), (2.0, 3.0), 0.0, (2.0, 2.0), 0.0) == (4, 1.0)
    # starts inside the skin: touching at t = 0; starts overlapped: state 2
    assert toi_probe(edge, (hx, hy), (2.0, 1.0 + hy + 0.004), 0.0, (2.0, 0.5), 0.0) == (3, 0.0)
    assert toi_probe(edge, (hx, hy), (2.0, 1.0 + hy - 0.05), 0.0, (2.0, 0.5), 0.0)[0] == 2
    # pure rotation about the centroid above the edge: the corner radius is r = |(hx, hy)|; with the centre at height h
    # the lowest corner reaches y = h - r cos(phi - angle) ... touching when that equals 1 + target
    r, phi = np.hypot(hx, hy), np.arctan2(hx, hy)
    h = 1.0 + hy + 0.007  # 7 mm above the edge when upright; the corner dips below the edge line as the box turns
    state, t = toi_probe(edge, (hx, hy), (2.0, h), 0.0, (2.0, h), 0.3)
    assert state == 3
    ang = 0.3 * t
    low = h - r * np.cos(phi - ang) if ang <= phi else h - r
    assert abs((low - 1.0) - target) <= tol * 1.05, (t, low)
    # a sweep that dips through the edge line and comes back out is reported separated: conservative advancement only
    # looks at the final configuration first (real Box2D behaviour, kept)
    assert toi_probe(edge, (hx, hy), (2.0, h), 0.0, (2.0, h), 0.6) == (4, 1.0)


def test_continuous_collision_runs_during_landings_and_prevents_deep_first_penetration():
    """SolveTOI is exercised whenever a body arrives at the terrain: every first touch of a fast random-policy crash or of a
    heuristic landing goes through a TOI sub-step, so a body is never found deeper than the solver's slop on the step a
    contact begins (without continuous collision a leg moving at ~3 m/s would be ~5 cm inside)."""
    env = OracleLunarLander(64)
    env.reset(seed=123)
    rs = np.random.default_rng(5)
    events = calls = 0
    for t in range(260):
        env.step(rs.integers(0, 4, 64))
        if t % 20 == 19:
            st = [env.toi_stats(i) for i in range(64)]
            calls, events = max(calls, sum(s[0] for s in st)), max(events, sum(s[1] for s in st))
    assert calls > 64 and events >= 16, (calls, events)


def test_toi_clearly_separated_shortcut_never_changes_a_result():
    """The CUDA engine skips b2TimeOfImpact for edge/polygon pairs that provably stay apart during the sweep
    (lunarlander.cu: toi_clearly_separated).  The oracle always runs b2TimeOfImpact and records, for every evaluation the
    predicate covers, whether the answer was alpha = 1: it must be, every time, for random and for landing trajectories."""
    n = 96
    covered = wrong = calls = 0
    for policy in ("random", "heuristic"):
        env = OracleLunarLander(n)
        obs = env.reset(seed=31)[0]
        rs = np.random.default_rng(11)
        prev = np.zeros((n, 3), dtype=np.int64)
        for _ in range(420):
            a = rs.integers(0, 4, n) if policy == "random" else np.array([heuristic(o) for o in obs])
            obs = env.step(a)[0]
            cur = np.array([(env.toi_stats(i)[0],) + env.toi_shortcut_stats(i) for i in range(n)], dtype=np.int64)
            d = cur - prev
            d[d[:, 0] < 0] = cur[d[:, 0] < 0]  # counters restart with the episode
            prev = cur
            calls, covered, wrong = calls + d[:, 0].sum(), covered + d[:, 1].sum(), wrong + d[:, 2].sum()
    assert wrong == 0
    assert calls > 2000 and covered > 0.25 * calls, (calls, covered)


def _lander_masses():
    poly = np.array([(-14, 17), (-17, 0), (-17, -10), (17, -10), (17, 0), (14, 17)], dtype=np.float64) / 30.0
    x, y = poly[:, 0], poly[:, 1]
    area = 0.5 * abs(np.dot(x, np.roll(y, -1)) - np.dot(y, np.roll(x, -1)))
    return np.array([5.0 * area, 1.0 * (2 * 2 / 30.0) * (2 * 8 / 30.0), 1.0 * (2 * 2 / 30.0) * (2 * 8 / 30.0)])


def test_wind_and_continuous_free_flight_known_answer():
    """LunarLander(continuous=True, enable_wind=True), engines off: before any contact the external forces on (lander + legs)
    are gravity, the initial random force and the wind, F_wind = tanh(sin(0.02 i) + sin(pi 0.01 i)) * 15 with i drawn by
    np_random.integers(-9999, 9999) at reset and advanced once per step (lunar_lander.py:401-403, :476-490).  The total
    momentum then has a closed form; the draws come from numpy itself and the wind from Python's math module."""
    import math

    n, seed, h, g = 16, 77, 1.0 / 50.0, -10.0
    env = OracleLunarLander(n, continuous=True, enable_wind=True, wind_power=15.0, turbulence_power=1.5)
    env.reset(seed=seed)
    masses = _lander_masses()
    F, idx = np.zeros((n, 2)), np.zeros(n, dtype=np.int64)
    for i in range(n):
        gen = np.random.Generator(np.random.PCG64(np.random.SeedSequence(seed + i)))
        gen.uniform(0, 400 / 30.0 / 2, size=(12,))
        F[i] = [gen.uniform(-1000.0, 1000.0), gen.uniform(-1000.0, 1000.0)]
        idx[i] = gen.integers(-9999, 9999)
        t_idx = int(gen.integers(-9999, 9999))
        assert env.wind_state(i) == (int(idx[i]) + 1, t_idx + 1)  # reset() ends with one step

    def wind(i):
        return math.tanh(math.sin(0.02 * i) + math.sin(math.pi * 0.01 * i)) * 15.0

    def momentum():
        b = np.stack([env.debug_state(i)[0] for i in range(n)]).astype(np.float64)  # (n, 3, 7)
        return (masses[None, :, None] * b[:, :, 3:5]).sum(axis=1)

    px = h * F[:, 0] + h * np.array([wind(int(i)) for i in idx])
    py = h * F[:, 1] + masses.sum() * h * g
    p = momentum()
    np.testing.assert_allclose(p[:, 0], px, rtol=5e-5, atol=5e-4)
    np.testing.assert_allclose(p[:, 1], py, rtol=5e-5, atol=5e-4)
    a = np.zeros((n, 2), dtype=np.float32)
    a[:, 1] = 0.5  # |a1| == 0.5 is NOT above the side-engine threshold; a0 == 0 is not above the main-engine one
    for k in range(1, 22):
        o, r, te, tr, _ = env.step(a)
        assert not te.any() and (o[:, 6:8] == 0).all()
        px = px + h * np.array([wind(int(i) + k) for i in idx])
        py = py + masses.sum() * h * g
        p = momentum()
        np.testing.assert_allclose(p[:, 0], px, rtol=1e-4, atol=1e-3, err_msg=f"step {k}")
        np.testing.assert_allclose(p[:, 1], py, rtol=1e-4, atol=1e-3, err_msg=f"step {k}")


def test_continuous_actions_throttle_and_fuel_cost():
    """continuous=True (lunar_lander.py:509-616): the main engine fires for a0 > 0 with power (clip(a0, 0, 1) + 1) / 2, the
    side engines for |a1| > 0.5 with power clip(|a1|, 0.5, 1); fuel costs 0.30 / 0.03 per unit power; actions are clipped to
    [-1, 1] first.  Two copies of the same env that differ only in one throttle differ in reward by the fuel term plus the
    shaping change of the extra impulse; an action beyond the box equals its clipped value exactly."""
    def run(act, steps=3):
        env = OracleLunarLander(1, continuous=True)
        env.reset(seed=5)
        out = [env.step(np.array([act], dtype=np.float32)) for _ in range(steps)]
        return out, env.debug_state(0)[0]

    idle, b_idle = run([0.0, 0.0])
    clipped, b_c = run([1.0, -1.0])
    beyond, b_b = run([7.5, -3.0])
    np.testing.assert_array_equal(clipped[-1][0], beyond[-1][0])
    np.testing.assert_array_equal(b_c, b_b)
    assert all(x[1][0] == y[1][0] for x, y in zip(clipped, beyond))
    # below the thresholds nothing fires: identical to idling
    quiet, b_q = run([-0.3, 0.5])
    np.testing.assert_array_equal(quiet[-1][0], idle[-1][0])
    np.testing.assert_array_equal(b_q, b_idle)
    # the main engine pushes the lander up relative to idling, more so at full throttle than at the minimum 50 %
    half, b_h = run([1e-6, 0.0])
    full, b_f = run([1.0, 0.0])
    assert b_idle[0, 4] < b_h[0, 4] < b_f[0, 4]
    gain_half, gain_full = b_h[0, 4] - b_idle[0, 4], b_f[0, 4] - b_idle[0, 4]
    assert 1.7 < gain_full / gain_half < 2.3  # power 1.0 against 0.5 (dispersion draws are the same in both runs)
    # the side engine turns the lander, in opposite directions for opposite signs
    left, b_l = run([0.0, -1.0])
    right, b_r = run([0.0, 1.0])
    assert (b_l[0, 5] - b_idle[0, 5]) * (b_r[0, 5] - b_idle[0, 5]) < 0


def test_discrete_env_is_unchanged_by_the_wind_and_continuous_code():
    a = OracleLunarLander(4)
    b = OracleLunarLander(4, continuous=False, enable_wind=False, wind_power=3.0, turbulence_power=0.2)
    np.testing.assert_array_equal(a.reset(seed=9)[0], b.reset(seed=9)[0])
    rs = np.random.default_rng(1)
    for _ in range(120):
        act = rs.integers(0, 4, 4)
        x, y = a.step(act), b.step(act)
        for k in range(4):
            np.testing.assert_array_equal(x[k], y[k])


def test_turbulence_changes_the_angular_momentum_by_the_applied_torque():
    """enable_wind=True in free flight: per step the angular momentum of (lander + legs) about their common mass centre changes
    by h (r x F_wind + tau) with r = lander centre - mass centre, F_wind and tau the reference's wind / turbulence pattern
    (lunar_lander.py:476-506) evaluated with Python's math -- gravity has no moment about the mass centre and the joint
    impulses are internal.  Holds step by step under Box2D's semi-implicit Euler (positions move along the NEW velocities,
    which changes no cross product); checks ApplyTorque -> integration, which the momentum test cannot see."""
    import math

    n, h = 8, 1.0 / 50.0
    env = OracleLunarLander(n, continuous=True, enable_wind=True, wind_power=15.0, turbulence_power=2.0)
    env.reset(seed=5)
    _, misc = env.debug_state(0)
    m = np.array([misc[0], misc[2], misc[2]], dtype=np.float64)
    inertia = np.array([misc[1], misc[3], misc[3]], dtype=np.float64)
    lc = np.array([[misc[4], misc[5]], [0, 0], [0, 0]], dtype=np.float64)  # local centres of mass (legs: box centres)

    def state():
        b = np.stack([env.debug_state(i)[0] for i in range(n)]).astype(np.float64)  # (n, 3, 7): origin x y, angle, v, w
        p, a, v, w = b[:, :, 0:2], b[:, :, 2], b[:, :, 3:5], b[:, :, 5]
        c = p + np.stack([np.cos(a) * lc[None, :, 0] - np.sin(a) * lc[None, :, 1],
                          np.sin(a) * lc[None, :, 0] + np.cos(a) * lc[None, :, 1]], axis=-1)
        return c, v, w

    def l_com(c, v, w):
        big_m = m.sum()
        cm, vm = (m[None, :, None] * c).sum(1) / big_m, (m[None, :, None] * v).sum(1) / big_m
        r, u = c - cm[:, None, :], v - vm[:, None, :]
        return (m[None, :] * (r[:, :, 0] * u[:, :, 1] - r[:, :, 1] * u[:, :, 0])).sum(1) + (inertia[None, :] * w).sum(1), cm

    def pattern(i, power):
        return math.tanh(math.sin(0.02 * i) + math.sin(math.pi * 0.01 * i)) * power

    idx = np.array([env.wind_state(i) for i in range(n)])  # the indices the NEXT step uses
    c, v, w = state()
    l0, cm = l_com(c, v, w)
    a = np.zeros((n, 2), dtype=np.float32)
    signal = 0.0
    for k in range(20):
        o, r, te, tr, _ = env.step(a)
        assert not te.any() and (o[:, 6:8] == 0).all()
        force = np.array([pattern(int(idx[i, 0]) + k, 15.0) for i in range(n)])
        torque = np.array([pattern(int(idx[i, 1]) + k, 2.0) for i in range(n)])
        arm = c[:, 0, :] - cm
        expect = h * (-arm[:, 1] * force + torque)  # (r x (F, 0))_z = -r_y F
        c, v, w = state()
        l1, cm1 = l_com(c, v, w)
        np.testing.assert_allclose(l1 - l0, expect, rtol=0, atol=2e-4, err_msg=f"step {k}")
        signal = max(signal, np.abs(expect).max())
        l0, cm = l1, cm1
    assert signal > 0.02  # the check resolves the turbulence term (h * 2.0 * tanh(...) ~ 0.04) 100x above its tolerance


@pytest.mark.parametrize("continuous", [False, True])
def test_engine_impulses_change_the_momentum_as_the_reference_formulas_say(continuous):
    """Free flight with the engines firing: every step the total momentum of (lander + legs) changes by the engine impulses of
    lunar_lander.py:536-616 plus h M g.  The impulses are recomputed here from the reference's expressions -- tip / side from
    the lander angle, the two dispersion draws from numpy's own stream in the reference's order (12 terrain heights, fx, fy,
    then 2 per step, including the step inside reset), MAIN_ENGINE_POWER 13, SIDE_ENGINE_POWER 0.6, the continuous throttle
    rules -- not from oracle/lunar_lander.c."""
    import math

    n, seed, h, g, scale = 12, 31, 1.0 / 50.0, -10.0, 30.0
    env = OracleLunarLander(n, continuous=continuous)
    env.reset(seed=seed)
    masses = _lander_masses()
    gens = []
    for i in range(n):
        gen = np.random.Generator(np.random.PCG64(np.random.SeedSequence(seed + i)))
        gen.uniform(0, 400 / 30.0 / 2, size=(12,))
        gen.uniform(-1000.0, 1000.0), gen.uniform(-1000.0, 1000.0)
        gen.uniform(-1.0, 1.0), gen.uniform(-1.0, 1.0)  # the dispersion draws of reset()'s own step
        gens.append(gen)

    def bodies():
        return np.stack([env.debug_state(i)[0] for i in range(n)]).astype(np.float64)

    def momentum(b):
        return (masses[None, :, None] * b[:, :, 3:5]).sum(axis=1)

    rs = np.random.default_rng(4)
    b = bodies()
    p = momentum(b)
    fired_main = fired_side = 0
    for k in range(18):
        if continuous:
            act = rs.uniform(-1.0, 1.0, size=(n, 2)).astype(np.float32)
        else:
            act = rs.integers(0, 4, n)
        expect = np.zeros((n, 2))
        for i in range(n):
            angle = float(b[i, 0, 2])
            tip, side = (math.sin(angle), math.cos(angle)), (-math.cos(angle), math.sin(angle))
            d = [gens[i].uniform(-1.0, +1.0) / scale for _ in range(2)]
            a = np.clip(act[i], -1, +1).astype(np.float64) if continuous else int(act[i])
            if (continuous and a[0] > 0.0) or (not continuous and a == 2):
                m_power = (np.clip(a[0], 0.0, 1.0) + 1.0) * 0.5 if continuous else 1.0
                ox = tip[0] * (4 / scale + 2 * d[0]) + side[0] * d[1]  # MAIN_ENGINE_Y_LOCATION = 4
                oy = -tip[1] * (4 / scale + 2 * d[0]) - side[1] * d[1]
                expect[i] += [-ox * 13.0 * m_power, -oy * 13.0 * m_power]
                fired_main += 1
            if (continuous and abs(a[1]) > 0.5) or (not continuous and a in (1, 3)):
                direction = np.sign(a[1]) if continuous else a - 2
                s_power = np.clip(abs(a[1]), 0.5, 1.0) if continuous else 1.0
                ox = tip[0] * d[0] + side[0] * (3 * d[1] + direction * 12 / scale)  # SIDE_ENGINE_AWAY = 12
                oy = -tip[1] * d[0] - side[1] * (3 * d[1] + direction * 12 / scale)
                expect[i] += [-ox * 0.6 * s_power, -oy * 0.6 * s_power]
                fired_side += 1
        expect[:, 1] += masses.sum() * h * g
        o, r, te, tr, _ = env.step(act)
        assert not te.any() and (o[:, 6:8] == 0).all()
        b = bodies()
        p1 = momentum(b)
        np.testing.assert_allclose(p1 - p, expect, rtol=0, atol=3e-4, err_msg=f"step {k}")
        p = p1
    assert fired_main > 30 and fired_side > 30


@pytest.mark.parametrize("continuous", [False, True])
def test_observation_and_reward_follow_the_reference_expressions_exactly(continuous):
    """lunar_lander.py:620-662 re-evaluated in Python from the body states the oracle exposes (position, velocity, angle,
    angular velocity as float32, like pybox2d hands them out): the 8 observation entries, the shaping reward, the fuel costs
    (0.30 / 0.03 per unit power), -100 on a crash or when |x| >= 1, +100 when the lander sleeps.  Exact equality: both sides
    are the same IEEE double expressions."""
    n = 24
    env = OracleLunarLander(n, continuous=continuous, max_episode_steps=0)
    obs, _ = env.reset(seed=13)
    w, hgt, scale, fps, leg_down = 600, 400, 30.0, 50, 18
    helipad_y = hgt / scale / 4

    def state(i, legs):
        b = env.debug_state(i)[0]
        px, py, ang, vx, vy, om = (float(b[0, k]) for k in range(6))
        return [(px - w / scale / 2) / (w / scale / 2), (py - (helipad_y + leg_down / scale)) / (hgt / scale / 2),
                vx * (w / scale / 2) / fps, vy * (hgt / scale / 2) / fps, ang, 20.0 * om / fps, legs[0], legs[1]]

    def shaping(s):
        return (-100 * np.sqrt(s[0] * s[0] + s[1] * s[1]) - 100 * np.sqrt(s[2] * s[2] + s[3] * s[3]) - 100 * abs(s[4])
                + 10 * s[6] + 10 * s[7])

    prev = [shaping(state(i, [float(obs[i, 6]), float(obs[i, 7])])) for i in range(n)]
    for i in range(n):
        np.testing.assert_array_equal(obs[i], np.array(state(i, obs[i, 6:8].astype(np.float64)), dtype=np.float32))
    rs = np.random.default_rng(2)
    pending = np.zeros(n, dtype=bool)
    ended = 0
    for t in range(400):
        act = rs.uniform(-1, 1, size=(n, 2)).astype(np.float32) if continuous else rs.integers(0, 4, n)
        o, r, te, tr, _ = env.step(act)
        for i in range(n):
            legs = [float(o[i, 6]), float(o[i, 7])]
            s = state(i, legs)
            np.testing.assert_array_equal(o[i], np.array(s, dtype=np.float32), err_msg=f"obs of env {i} at step {t}")
            if pending[i]:  # this call was the env's reset: reward 0, and the shaping memory restarts
                assert r[i] == 0.0 and not te[i]
                prev[i] = shaping(s)
                continue
            if continuous:
                a = np.clip(act[i], -1, +1).astype(np.float64)
                m_power = (np.clip(a[0], 0.0, 1.0) + 1.0) * 0.5 if a[0] > 0.0 else 0.0
                s_power = np.clip(np.abs(a[1]), 0.5, 1.0) if np.abs(a[1]) > 0.5 else 0.0
            else:
                m_power, s_power = (1.0 if act[i] == 2 else 0.0), (1.0 if act[i] in (1, 3) else 0.0)
            sh = shaping(s)
            reward = sh - prev[i]
            prev[i] = sh
            reward -= m_power * 0.30
            reward -= s_power * 0.03
            if te[i]:
                assert r[i] in (-100.0, 100.0)
                awake = env.debug_state(i)[0][0, 6] != 0
                assert r[i] == (-100.0 if awake else 100.0)
                ended += 1
            else:
                assert abs(s[0]) < 1.0
                assert r[i] == reward, (t, i, r[i], reward)
        pending = te | tr
    assert ended > n  # every env crashed or landed at least once on average
