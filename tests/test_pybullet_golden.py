"""PyBullet golden vectors (tests/golden/walker3d_pybullet_*.npz, made by tools/gen_pybullet_golden.py on a host
that has PyBullet).  None is committed: PyBullet cannot be installed in the build container (no wheel, no
network), so PyBullet parity is UNVERIFIED and these checks skip; they run automatically once a file is added."""
import glob
import os

import numpy as np
import pytest

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "walker3d_pybullet_*.npz")))


@pytest.mark.skipif(not GOLDEN, reason="no PyBullet golden file (PyBullet parity unverified; see DESIGN.md section 5)")
def test_oracle_mass_matrix_vs_pybullet(walker_table, oracle_mod):
    O, t = oracle_mod, walker_table
    g = np.load(GOLDEN[-1])
    m = O.model_from_table(t)
    for st, Mpb in zip(g["free_states"], g["mass_matrix"]):
        s = O.make_state(21, st[0:3], st[3:7], st[7:10], st[10:13], st[13:34], st[34:55])
        M = O.mass_matrix(m, s)
        assert np.abs(M - Mpb).max() / np.abs(Mpb).max() < 1e-4


@pytest.mark.skipif(not GOLDEN, reason="no PyBullet golden file (PyBullet parity unverified; see DESIGN.md section 5)")
def test_oracle_contact_free_step_vs_pybullet(walker_table, oracle_mod):
    O, t = oracle_mod, walker_table
    g = np.load(GOLDEN[-1])
    m = O.model_from_table(t)
    p = O.default_params()
    for st, step in zip(g["free_states"], g["free_steps"]):
        s = O.make_state(21, st[0:3], st[3:7], st[7:10], st[10:13], st[13:34], st[34:55])
        O.step_physics(m, p, s, step[:21])
        ref = step[21:]
        out = O.state_vector(s, 21)
        assert np.max(np.abs(out - ref) / np.maximum(1.0, np.abs(ref))) < 1e-4
