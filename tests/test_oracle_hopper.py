"""Hopper-v5 oracle (oracle/hopper.c, MuJoCo-subset restatement; PARITY UNPINNED -- mujoco is not installable here).  What CAN
be pinned: the reference's structural checks for this boundary, analytic known answers, self-consistency."""
import numpy as np

from oracle.hopper import NB, OracleHopper


def test_model_sizes_and_masses():
    """nq = nv = 6, nu = 3, nbody = 5 (world + 4), observation 11 (tests/envs/mujoco/test_mujoco_v5.py:538-547 lists the
    model sizes of every v5 env; hopper_v5.py:232-236 the observation).  Body masses = capsule volumes x density 1000."""
    env = OracleHopper(2)
    obs, info = env.reset(seed=3)
    assert obs.shape == (2, 11) and set(info) >= {"x_position", "z_distance_from_origin"}
    mass, misc, inv = env.model_info()
    assert len(mass) == NB == 5 and mass[0] == 0

    def capsule(r, half):
        return 1000.0 * (np.pi * r * r * 2 * half + 4.0 / 3.0 * np.pi * r ** 3)

    np.testing.assert_allclose(mass[1:], [capsule(0.05, 0.2), capsule(0.05, 0.225), capsule(0.04, 0.25), capsule(0.06, 0.195)],
                               rtol=1e-12)
    assert misc[1] == 7  # floor x 4 body geoms + torso-leg, torso-foot, thigh-foot


def test_noise_free_reset_is_the_init_state_and_free_fall_known_answer():
    """reset_noise_scale = 0 => qpos = qpos0 = (0, 1.25, 0, 0, 0, 0) (the `ref` of rootz), qvel = 0.  Until the foot reaches the
    floor the only external force is gravity: the mass centre falls with z(t) = z0 - g t^2 / 2 exactly under RK4 (the joint
    armature of 1 kg m^2 on the three leg hinges is rotor inertia outside the body masses: identity to ~1e-6)."""
    env = OracleHopper(1, reset_noise_scale=0.0)
    obs, info = env.reset(seed=0)
    np.testing.assert_array_equal(obs[0], [1.25] + [0.0] * 10)
    mass, _, _ = env.model_info()
    M = mass.sum()
    xipos0 = env.debug(0)[4]
    z0 = (mass * xipos0[:, 2]).sum() / M
    free = 0
    for k in range(12):
        obs, r, te, tr, info = env.step(np.zeros((1, 3), dtype=np.float32))
        qpos, qvel, qacc, counts, xipos = env.debug(0)
        if counts[0] > 0:
            break
        t = (k + 1) * 4 * 0.002
        # xipos is from the last mj_forward of the step = RK4 stage 4 of the last sub-step, evaluated at time t
        zc = (mass * xipos[:, 2]).sum() / M
        assert abs(zc - (z0 - 0.5 * 9.81 * t * t)) < 5e-6, (k, zc)
        assert abs((mass * xipos[:, 0]).sum() / M - (mass * xipos0[:, 0]).sum() / M) < 1e-9
        assert r[0] == 1.0 and info["reward_ctrl"][0] == 0.0  # healthy, zero control
        free += 1
    assert free >= 8


def test_reward_identity_health_and_termination():
    """reward == reward_forward + reward_survive + reward_ctrl (tests/envs/mujoco/test_mujoco_v5.py:236-241, exact ==);
    unhealthy (z <= 0.7 or |angle| >= 0.2) terminates with reward_survive 0; random policies fall within ~20-30 steps."""
    n = 48
    env = OracleHopper(n)
    env.reset(seed=5)
    rs = np.random.default_rng(1)
    lens, cur, nterm = [], np.zeros(n, int), 0
    prev_done = np.zeros(n, bool)
    for t in range(150):
        a = rs.uniform(-1, 1, size=(n, 3)).astype(np.float32)
        obs, r, te, tr, info = env.step(a)
        live = ~prev_done
        total = info["reward_forward"] + info["reward_survive"] + info["reward_ctrl"]
        assert (r[live] == total[live]).all()
        assert (r[prev_done] == 0).all() and not te[prev_done].any()
        assert (info["reward_survive"][live & te] == 0).all() and (info["reward_survive"][live & ~te] == 1.0).all()
        z, ang = obs[:, 0], obs[:, 1]
        assert ((z[live & ~te] > 0.7) & (np.abs(ang[live & ~te]) < 0.2)).all()
        # float32 control cost of the float32 action (NumPy 2 / NEP 50): -1e-3 * sum(a^2) in float32
        want = -(np.float32(1e-3) * (a[:, 0] * a[:, 0] + a[:, 1] * a[:, 1] + a[:, 2] * a[:, 2])).astype(np.float64)
        np.testing.assert_array_equal(info["reward_ctrl"][live], want[live])
        assert np.abs(obs[:, 5:]).max() <= 10.0
        cur += 1
        for i in np.nonzero(te | tr)[0]:
            lens.append(cur[i]); cur[i] = 0
        nterm += int(te.sum())
        prev_done = te | tr
    assert nterm > n and 8 < np.mean(lens) < 60


def test_determinism_and_seeding():
    a, b, c = OracleHopper(3), OracleHopper(3), OracleHopper(3)
    oa, _ = a.reset(seed=11)
    ob, _ = b.reset(seed=11)
    oc, _ = c.reset(seed=12)
    np.testing.assert_array_equal(oa, ob)
    assert not np.array_equal(oa, oc)
    np.testing.assert_array_equal(oa[1], oc[0])  # sub-env i is seeded seed + i
    # reset noise: qpos0 + U(-5e-3, 5e-3) from numpy's stream, 6 qpos draws then 6 qvel draws (hopper_v5.py:322-337)
    g = np.random.Generator(np.random.PCG64(np.random.SeedSequence(11)))
    qpos = np.array([0, 1.25, 0, 0, 0, 0]) + g.uniform(-5e-3, 5e-3, size=6)
    qvel = g.uniform(-5e-3, 5e-3, size=6)
    np.testing.assert_array_equal(oa[0], np.concatenate([qpos[1:], np.clip(qvel, -10, 10)]))


# ---------------------------------------------------------------------------------------------------------------------
# Walker2d-v5 on the same core (oracle/walker2d.c -> mjc_planar.h)
def test_walker2d_model_free_fall_and_rewards():
    from oracle.walker2d import OracleWalker2d

    env = OracleWalker2d(1, reset_noise_scale=0.0)
    obs, info = env.reset(seed=0)
    assert obs.shape == (1, 17)  # nq = nv = 9 (tests/envs/mujoco/test_mujoco_v5.py:538-547), observation 8 + 9
    np.testing.assert_array_equal(obs[0], [1.25] + [0.0] * 16)
    mass, misc, inv = env.model_info()

    def capsule(r, half):
        return 1000.0 * (np.pi * r * r * 2 * half + 4.0 / 3.0 * np.pi * r ** 3)

    leg = [capsule(0.05, 0.225), capsule(0.04, 0.25), capsule(0.06, 0.1)]
    np.testing.assert_allclose(mass[1:], [capsule(0.05, 0.2)] + leg + leg, rtol=1e-12)
    assert misc[1] == 7  # contype 1 / conaffinity 0 on the robot's geoms: floor x 7 geoms, no self-collision (walker2d_v5.xml:11)
    M = mass.sum()
    xipos0 = env.debug(0)[4]
    z0 = (mass * xipos0[:, 2]).sum() / M
    free = 0
    for k in range(14):
        obs, r, te, tr, info = env.step(np.zeros((1, 6), dtype=np.float32))
        qpos, qvel, qacc, counts, xipos = env.debug(0)
        if counts[0] > 0:
            break
        t = (k + 1) * 4 * 0.002
        assert abs((mass * xipos[:, 2]).sum() / M - (z0 - 0.5 * 9.81 * t * t)) < 5e-6
        free += 1
    assert free >= 8
    n = 32
    env = OracleWalker2d(n)
    env.reset(seed=2)
    rs = np.random.default_rng(0)
    prev = np.zeros(n, bool)
    nterm = 0
    for t in range(120):
        a = rs.uniform(-1, 1, size=(n, 6)).astype(np.float32)
        obs, r, te, tr, info = env.step(a)
        live = ~prev
        # walker2d_v5.py:305-312 and tests/envs/mujoco/test_mujoco_v5.py reward identity, exact
        assert (r[live] == (info["reward_forward"] + info["reward_survive"] + info["reward_ctrl"])[live]).all()
        z, ang = obs[:, 0], obs[:, 1]
        healthy = (0.8 < z) & (z < 2.0) & (-1.0 < ang) & (ang < 1.0)  # walker2d_v5.py:261-272
        np.testing.assert_array_equal(te[live], ~healthy[live])
        nterm += int(te.sum())
        prev = te | tr
    assert nterm > n // 2


# InvertedPendulum-v5 on the same core (oracle/inverted_pendulum.c -> mjc_planar.h)
def _capsule(r, half, rho=1000.0):
    h = 2 * half
    ms, mc = rho * 4 / 3 * np.pi * r ** 3, rho * np.pi * r * r * h
    return ms + mc, mc * (3 * r * r + h * h) / 12 + ms * (0.4 * r * r + 0.25 * h * h + 0.375 * r * h)


def test_inverted_pendulum_matches_the_cart_pole_equations_of_motion():
    """Analytic pin of the multibody core (kinematics, composite inertia, bias forces, factorisation, damping, actuation,
    RK4): the cart-pole Lagrangian for inverted_pendulum.xml -- cart 10.47 kg on a damped slider, pole = capsule r 0.049
    from (0,0,0) to (0.001,0,0.6) on a damped hinge, motor gear 100 -- integrated here with an independent RK4 at the
    model's timestep must give the oracle's observations."""
    from oracle.inverted_pendulum import OracleInvertedPendulum

    env = OracleInvertedPendulum(1, max_episode_steps=0)
    mass, misc, _ = env.model_info()
    m_cart, _ = _capsule(0.1, 0.1)
    length = np.hypot(0.001, 0.6)
    m_pole, i_pole = _capsule(0.049, length / 2)
    np.testing.assert_allclose(mass, [0.0, m_cart, m_pole], rtol=1e-13)
    assert misc[1] == 0  # contype 0: no collision pairs
    com = np.array([0.0005, 0.3])
    l, phi0, g = np.hypot(*com), np.arctan2(com[0], com[1]), 9.81

    def f(y, force):
        x, th, xd, thd = y
        a = th + phi0
        M = np.array([[m_cart + m_pole, m_pole * l * np.cos(a)], [m_pole * l * np.cos(a), i_pole + m_pole * l * l]])
        rhs = np.array([force - 1.0 * xd + m_pole * l * np.sin(a) * thd * thd, -1.0 * thd + m_pole * g * l * np.sin(a)])
        return np.concatenate([[xd, thd], np.linalg.solve(M, rhs)])

    def rk4(y, force, h=0.02):
        k1 = f(y, force); k2 = f(y + 0.5 * h * k1, force); k3 = f(y + 0.5 * h * k2, force); k4 = f(y + h * k3, force)
        return y + h / 6 * (k1 + 2 * k2 + 2 * k3 + k4)

    obs, info = env.reset(seed=3)
    assert info == {} or all(len(v) == 1 for v in info.values())
    assert obs.shape == (1, 4) and np.abs(obs).max() <= 0.01  # init_qpos = init_qvel = 0, noise 0.01
    y, rs, worst, terminated = obs[0].copy(), np.random.default_rng(0), 0.0, False
    for t in range(60):
        a = rs.uniform(-1.5, 1.5, size=(1, 1)).astype(np.float32)
        o, r, te, tr, info = env.step(a)
        if terminated:  # NEXT_STEP: this call was the reset
            assert r[0] == 0.0 and np.abs(o).max() <= 0.01
            y, terminated = o[0].copy(), False
            continue
        for _ in range(2):  # frame_skip
            y = rk4(y, 100.0 * float(a[0, 0]))
        worst = max(worst, np.abs(o[0] - y).max())
        assert te[0] == (abs(y[1]) > 0.2) and r[0] == (0.0 if te[0] else 1.0) and info["reward_survive"][0] == r[0]
        terminated = bool(te[0])
    assert worst < 1e-12, worst


def test_inverted_pendulum_slider_limit_and_ctrl_clamp():
    from oracle.inverted_pendulum import OracleInvertedPendulum

    env = OracleInvertedPendulum(1, max_episode_steps=0)
    env.reset(seed=0)
    env.set_state(0, [0.99, 0.0], [1.5, 0.0])  # the cart runs into the +1 end of the rail
    xs = []
    for _ in range(8):
        o = env.step(np.array([[0.0]], dtype=np.float32))[0]
        xs.append(o[0, 0])
    assert max(xs) < 1.05 and env.debug(0)[3][1] >= 0  # the soft limit holds the cart near the range
    env.set_state(0, [0.0, 0.0], [0.0, 0.0])
    a1 = env.step(np.array([[3.0]], dtype=np.float32))[0].copy()
    env.set_state(0, [0.0, 0.0], [0.0, 0.0])
    a2 = env.step(np.array([[30.0]], dtype=np.float32))[0]
    np.testing.assert_array_equal(a1, a2)  # ctrlrange -3 3 with ctrllimited: the force saturates at gear * 3


def test_model_sizes_follow_the_reference_structural_pins():
    """tests/envs/mujoco/test_mujoco_v5.py:516-640 lists nq / nv / nu / nbody / njnt / ngeom of every v5 model.  The oracles
    (and the kernels, whose host-side compile is compared with the oracles' bit for bit in tests/test_host_cpu.py) carry:
    Hopper 6/6/3/5/6/5, Walker2d 9/9/6/8/9/8, HalfCheetah 9/9/6/8/9/9, InvertedPendulum 2/2/1/3/2 and 2 of its 3 geoms (the rail, a world geom with
    contype 0, takes part in nothing and is not modelled)."""
    import oracle.half_cheetah as hc
    import oracle.hopper as hp
    import oracle.inverted_pendulum as ip
    import oracle.walker2d as w2
    from oracle.mjc_planar import ROBOTS

    assert (hp.NQ, hp.NV, hp.NU, hp.NB, hp.OBS) == (6, 6, 3, 5, 11)
    assert (w2.NQ, w2.NV, w2.NU, w2.NB, w2.OBS) == (9, 9, 6, 8, 17)
    assert (ip.NQ, ip.NV, ip.NU, ip.NB, ip.OBS) == (2, 2, 1, 3, 4)
    assert (hc.NQ, hc.NV, hc.NU, hc.NB, hc.OBS) == (9, 9, 6, 8, 17)  # HalfCheetah: nbody 8, njnt 9, ngeom 9
    for robot, (_, _, nb, nq, nu) in ROBOTS.items():
        if robot.endswith("_euler"):  # test-only instance (see test_euler_integrator_...)
            continue
        env = {"hopper": hp.OracleHopper, "walker2d": w2.OracleWalker2d, "inverted_pendulum": ip.OracleInvertedPendulum,
               "half_cheetah": hc.OracleHalfCheetah}[robot](1)
        obs, _ = env.reset(seed=0)
        mass, misc, inv = env.model_info()
        assert len(mass) == nb and mass[0] == 0 and (mass[1:] > 0).all()
        assert obs.shape[1] == env.obs_size and len(inv) == 2 * nb + nq
        qpos, qvel, qacc, counts, xipos = env.debug(0)
        assert qpos.shape == (nq,) and qvel.shape == (nq,) and xipos.shape == (nb, 3)
        a = np.zeros((1, nu), dtype=np.float32)
        o, r, te, tr, info = env.step(a)
        assert o.shape == obs.shape and np.isfinite(o).all()


# HalfCheetah-v5 on the same core (oracle/half_cheetah.c -> mjc_planar.h): Euler integrator with implicit joint damping, joint
# springs, settotalmass, axisangle / fromto geoms, standard_normal reset noise
def test_half_cheetah_masses_match_mujoco_and_reset_noise_matches_numpy():
    from oracle.half_cheetah import OracleHalfCheetah

    n, seed = 400, 123
    env = OracleHalfCheetah(n)
    mass, misc, _ = env.model_info()
    # mjModel.body_mass of the stock half_cheetah.xml (inertiafromgeom, then settotalmass = 14)
    known = [0.0, 6.25020921, 1.54351464, 1.5874477, 1.09539749, 1.43807531, 1.20083682, 0.88451883]
    np.testing.assert_allclose(mass, known, rtol=1e-8)  # the literature values carry 8 decimals
    assert abs(mass.sum() - 14.0) < 1e-12 and misc[1] == 8  # 8 capsules against the floor, nothing else collides
    obs, info = env.reset(seed=seed)
    assert obs.shape == (n, 17) and set(info) == {"x_position", "x_velocity", "reward_forward", "reward_ctrl"}
    for i in range(n):  # half_cheetah_v5.py:261-276: uniform(-0.1, 0.1, nq) then 0.1 * standard_normal(nv)
        gen = np.random.Generator(np.random.PCG64(np.random.SeedSequence(seed + i)))
        qpos = gen.uniform(low=-0.1, high=0.1, size=9)
        qvel = 0.1 * gen.standard_normal(9)
        np.testing.assert_array_equal(obs[i], np.concatenate([qpos[1:], qvel]))
        assert info["x_position"][i] == qpos[0]


def test_half_cheetah_free_fall_reward_identity_and_time_limit():
    from oracle.half_cheetah import OracleHalfCheetah

    env = OracleHalfCheetah(1, reset_noise_scale=0.0, max_episode_steps=7)
    obs, _ = env.reset(seed=0)
    assert (obs == 0).all()  # qpos0 = 0 (the torso's height 0.7 is the body position, rootz itself starts at 0)
    h, g = 0.01, 9.81
    zero = np.zeros((1, 6), dtype=np.float32)
    for step in (1, 2):  # 10 simulation steps before a foot reaches the floor: semi-implicit Euler, z_k = -g h^2 k (k + 1) / 2
        o, r, te, tr, info = env.step(zero)
        k = 5 * step
        np.testing.assert_allclose(o[0, 0], -g * h * h * k * (k + 1) / 2, rtol=1e-12)
        np.testing.assert_allclose(o[0, 9], -g * h * k, rtol=1e-12)  # qvel of rootz
        assert np.abs(o[0, 1:8]).max() < 1e-12 and env.debug(0)[3][0] == 0  # nothing bends, no contact yet
    rs = np.random.default_rng(3)
    env.reset(seed=1)
    for t in range(7):
        a = rs.uniform(-1, 1, size=(1, 6)).astype(np.float32)
        x0 = env.debug(0)[0][0]
        o, r, te, tr, info = env.step(a)
        x1 = env.debug(0)[0][0]
        assert not te[0] and tr[0] == (t == 6)  # never terminates; TimeLimit truncates
        assert info["x_velocity"][0] == (x1 - x0) / 0.05 and info["x_position"][0] == x1
        ctrl = np.float32(0.1) * np.sum(np.square(a[0]))  # float32, as NumPy 2 evaluates it for float32 actions
        assert info["reward_ctrl"][0] == -float(ctrl) and info["reward_forward"][0] == 1.0 * info["x_velocity"][0]
        assert r[0] == info["reward_forward"][0] + info["reward_ctrl"][0]  # half_cheetah_v5.py:243
    o, r, te, tr, info = env.step(zero)  # NEXT_STEP: the reset call
    assert r[0] == 0.0 and not te[0] and not tr[0] and np.abs(o[0, :8]).max() <= 0.1 + 1e-12


def test_euler_integrator_with_implicit_damping_matches_the_cart_pole_equations():
    """mj_Euler as HalfCheetah-v5 runs it -- qvel += h (M + h B)^-1 (qfrc_smooth + qfrc_constraint), qpos += h qvel_new -- on a
    model whose equations of motion are known: a TEST-ONLY instance of the planar core compiles inverted_pendulum.xml with the
    Euler integrator (oracle/inverted_pendulum_euler.c); the same scheme written out for the cart-pole Lagrangian must give
    the same trajectory."""
    from oracle.mjc_planar import OraclePlanar

    class EulerPendulum(OraclePlanar):
        robot = "inverted_pendulum_euler"

    env = EulerPendulum(1, max_episode_steps=0, reset_noise_scale=0.01)
    m_cart, _ = _capsule(0.1, 0.1)
    m_pole, i_pole = _capsule(0.049, np.hypot(0.001, 0.6) / 2)
    com = np.array([0.0005, 0.3])
    l, phi0, g, h, damping = np.hypot(*com), np.arctan2(com[0], com[1]), 9.81, 0.02, np.array([1.0, 1.0])

    def step(y, force):
        x, th, xd, thd = y
        a = th + phi0
        M = np.array([[m_cart + m_pole, m_pole * l * np.cos(a)], [m_pole * l * np.cos(a), i_pole + m_pole * l * l]])
        smooth = np.array([force - damping[0] * xd + m_pole * l * np.sin(a) * thd * thd,
                           -damping[1] * thd + m_pole * g * l * np.sin(a)])
        v = np.array([xd, thd]) + h * np.linalg.solve(M + h * np.diag(damping), smooth)
        return np.concatenate([np.array([x, th]) + h * v, v])

    obs, _ = env.reset(seed=4)
    y, rs, worst = obs[0].copy(), np.random.default_rng(1), 0.0
    for _ in range(9):
        a = rs.uniform(-1, 1, size=(1, 1)).astype(np.float32)
        o, r, te, tr, _ = env.step(a)
        for _ in range(2):
            y = step(y, 100.0 * float(a[0, 0]))
        worst = max(worst, np.abs(o[0] - y).max())
        if te[0]:
            break
    assert worst < 1e-13, worst
    # and it is NOT what RK4 gives: the two integrators separate by far more than the tolerance above
    from oracle.inverted_pendulum import OracleInvertedPendulum

    rk = OracleInvertedPendulum(1, max_episode_steps=0)
    np.testing.assert_array_equal(rk.reset(seed=4)[0], obs)
    o_rk = rk.step(np.array([[0.7]], dtype=np.float32))[0]
    env.reset(seed=4)
    o_eu = env.step(np.array([[0.7]], dtype=np.float32))[0]
    assert np.abs(o_rk - o_eu).max() > 1e-4


def test_half_cheetah_at_rest_is_carried_by_its_contacts():
    """Statics of the contact pipeline (collision, pyramidal friction cone rows, contact Jacobians, the constraint solver): with
    zero control the cheetah settles on the floor, and whatever the soft-constraint parameters, a body at rest is carried by
    its contacts -- the normal forces (each the sum of its four cone rows) add up to M g, on floor normals (0, 0, 1)."""
    from oracle.half_cheetah import OracleHalfCheetah

    env = OracleHalfCheetah(1, reset_noise_scale=0.0, max_episode_steps=0)
    env.reset(seed=0)
    zero = np.zeros((1, 6), dtype=np.float32)
    for _ in range(400):
        o, r, te, tr, _ = env.step(zero)
        assert not te[0] and not tr[0]
    qpos, qvel, qacc, counts, _ = env.debug(0)
    force, normal = env.contact_forces(0)
    weight = env.model_info()[0].sum() * 9.81
    assert len(force) >= 2 and (force > 0).all() and np.abs(qvel).max() < 1e-4 and np.abs(qacc).max() < 1e-2
    np.testing.assert_array_equal(normal, np.tile([0.0, 0.0, 1.0], (len(force), 1)))
    np.testing.assert_allclose(force.sum(), weight, rtol=1e-5)
    assert counts[1] == 4 * len(force)  # condim 3: four pyramid rows per contact, no joint at its limit
