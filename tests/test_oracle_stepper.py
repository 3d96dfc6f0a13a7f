"""Walker3DStepperEnv oracle pins: terrain / reset draws bit-exact against a NumPy replay of
env_locomotion.py:395-441,481-513 (NumPy RandomState is the reference's RNG), plank geometry, episode sanity."""
import numpy as np

DEG2RAD = np.pi / 180


def numpy_step_placements(rs, curriculum):
    """Direct NumPy replay of Walker3DStepperEnv.generate_step_placements (env_locomotion.py:395-441)."""
    max_curriculum, n_steps, step_radius = 9, 20, 0.25
    dist_range0 = np.array([0.65, 1.25])
    ratio = curriculum / max_curriculum
    dist_upper = np.linspace(*dist_range0, max_curriculum + 1)
    dist_range = np.array([dist_range0[0], dist_upper[curriculum]])
    yaw_range = np.array([-20, 20]) * ratio * DEG2RAD
    pitch_range = np.array([-30, +30]) * ratio * DEG2RAD + np.pi / 2
    tilt_range = np.array([-15, 15]) * ratio * DEG2RAD
    N = n_steps
    dr = rs.uniform(*dist_range, size=N)
    dphi = rs.uniform(*yaw_range, size=N)
    dtheta = rs.uniform(*pitch_range, size=N)
    x_tilt = rs.uniform(*tilt_range, size=N)
    y_tilt = rs.uniform(*tilt_range, size=N)
    dr[0] = 0.0
    dphi[0] = 0.0
    dtheta[0] = np.pi / 2
    dr[1:3] = 0.75
    dphi[1:3] = 0.0
    dtheta[1:3] = np.pi / 2
    x_tilt[0:3] = 0
    y_tilt[0:3] = 0
    dphi = np.cumsum(dphi)
    dx = dr * np.sin(dtheta) * np.cos(dphi)
    dy = dr * np.sin(dtheta) * np.sin(dphi)
    dz = dr * np.cos(dtheta)
    dx_max = np.maximum(np.abs(dx[2:]), step_radius * 2.5)
    dx[2:] = np.sign(dx[2:]) * np.minimum(dx_max, dist_range0[1])
    x, y, z = np.cumsum(dx), np.cumsum(dy), np.cumsum(dz)
    return np.stack((x, y, z, dphi, x_tilt, y_tilt), axis=1)


def test_terrain_and_reset_bit_exact(walker_table, oracle_mod):
    O, t = oracle_mod, walker_table
    for seed, cur in ((0, 0), (1, 5), (2, 9), (3, 3)):
        env = O.Walker3DStepperOracle(t, seed=seed, curriculum=cur)
        rs = np.random.RandomState(O.gym_seed_words(seed))
        for episode in range(2):
            obs = env.reset()
            # robot.reset draws first (robots.py:182-194), then the terrain (env_locomotion.py:498)
            mirrored = rs.rand() < 0.5
            rs.uniform(low=-0.1, high=0.1, size=21)
            ref = numpy_step_placements(rs, cur)
            got = np.array(env.e.terrain[:])
            assert np.array_equal(got, ref), (seed, cur, np.abs(got - ref).max())
            assert bool(env.e.base.mirrored) == bool(mirrored)
            assert env.e.next_step_index == 1
            assert obs.shape == (65,)
            # walk_target = terrain row next+1 (lookahead 2): targets[-1]
            assert list(env.e.base.walk_target[:]) == list(ref[2, 0:3])


def test_plank_geometry(walker_table, oracle_mod):
    """Top surface of a flat plank is at the step origin; cover is the top 2.5 cm."""
    O, t = oracle_mod, walker_table
    env = O.Walker3DStepperOracle(t, seed=0, curriculum=0)
    env.reset()
    for p in range(3):
        base, cover = env.e.boxes[2 * p], env.e.boxes[2 * p + 1]
        step = np.array(env.e.terrain[p][:3])
        assert abs(cover.center[2] + cover.half[2] - step[2]) < 1e-12
        assert abs(base.center[2] + base.half[2] - (step[2] - 0.025)) < 1e-12
        assert (base.half[0], base.half[1]) == (0.25, 5.0)
    assert abs(env.e.boxes[2].center[0] - 0.75) < 1e-12


def test_passive_episode_and_curriculum(walker_table, oracle_mod):
    O, t = oracle_mod, walker_table
    lens = []
    for cur in (0, 9):
        env = O.Walker3DStepperOracle(t, seed=4, curriculum=cur)
        env.reset()
        for i in range(1000):
            obs, r, d, info = env.step(np.zeros(21))
            if d:
                break
        assert d and "steps_reached" in info and 1 <= info["steps_reached"] <= 19
        lens.append(i)
    # the lower terminal height of curriculum 9 (0.45 vs 0.75) lets the collapsing walker live longer
    assert lens[1] >= lens[0]
