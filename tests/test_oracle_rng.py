"""Bit-exactness of the oracle's MT19937 / legacy RandomState draws against NumPy."""
import ctypes as C

import numpy as np


def test_mt19937_matches_numpy(oracle_mod):
    O = oracle_mod
    for seed in (0, 1, 12345, 2 ** 40 + 7):
        words = O.gym_seed_words(seed)
        ref = np.random.RandomState(words)
        r = O.Rng()
        key = (C.c_uint32 * len(words))(*words)
        O.lib().orc_rng_seed_array(C.byref(r), key, len(words))
        st = ref.get_state()
        assert np.array_equal(np.array(r.mt[:], dtype=np.uint32), st[1])
        a = [O.lib().orc_rng_double(C.byref(r)) for _ in range(700)]
        b = ref.random_sample(700)
        assert np.array_equal(np.array(a), b)
        assert O.lib().orc_rng_uniform(C.byref(r), -0.1, 0.1) == ref.uniform(-0.1, 0.1)
        c = [30.0, 60.0][O.lib().orc_rng_u32(C.byref(r)) & 1]
        assert c == ref.choice([30.0, 60.0])
        assert O.lib().orc_rng_double(C.byref(r)) == ref.rand()


def test_gym_seed_words_shape():
    from oracle import oracle as O

    w = O.gym_seed_words(0)
    # sha512("0")[:8] as two little-endian uint32 words
    import hashlib
    import struct

    h = hashlib.sha512(b"0").digest()[:8]
    assert w == list(struct.unpack("2I", h))


def test_reset_state_bit_exact_vs_numpy(walker_table, oracle_mod):
    """Walker3DCustomEnv.reset() draw order (env_locomotion.py:79-109, robots.py:179-210) replayed with NumPy."""
    O, t = oracle_mod, walker_table
    for seed in range(4):
        env = O.Walker3DCustomOracle(t, seed=seed)
        rs = np.random.RandomState(O.gym_seed_words(seed))
        for episode in range(3):
            env.reset()
            dist = rs.uniform(3, 5)
            angle = rs.uniform(-np.pi / 2, np.pi / 2)
            stop = rs.choice([30.0, 60.0])
            ang = np.array(t["base_joint_angles"], dtype=np.float64)
            mirrored = rs.rand() < 0.5
            if mirrored:
                rl = t["right_joint_indices"] + t["left_joint_indices"]
                lr = t["left_joint_indices"] + t["right_joint_indices"]
                ang[rl] = ang[lr]
                ang[t["negation_joint_indices"]] *= -1
            ds = rs.uniform(low=-0.1, high=0.1, size=t["n_dof"])
            weight = np.array([u - l for u, l in zip(t["upper"], t["lower"])], dtype=np.float32)
            bias = np.array(t["lower"], dtype=np.float32)
            ps = 2 * (ang + ds - bias) / weight - 1
            q = weight * (np.clip(ps, -0.95, 0.95) + 1) / 2 + bias
            assert env.e.dist == dist and env.e.angle == angle and env.e.stop_frames == stop
            assert bool(env.e.mirrored) == bool(mirrored)
            assert np.array_equal(np.array(env.e.s.q[: t["n_dof"]]), q)
            assert list(env.e.walk_target) == [dist * np.cos(angle), dist * np.sin(angle), 1.0]
