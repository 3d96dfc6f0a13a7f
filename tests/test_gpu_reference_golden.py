"""The CUDA path against the reference-generated fixtures (tests/golden/ref_*.npz, recorded from the reference's own
env code by tools/gen_reference_golden.py).  The oracle replays a fixture exactly (tests/test_reference_golden.py), so
it can hand the device the state and bookkeeping the reference had before every step; the device's observation,
reward and done for that step are then compared with the RECORDED reference values (not with the oracle's)."""
import glob
import os

import pytest

pytestmark = pytest.mark.gpu
_G = os.path.join(os.path.dirname(__file__), "golden")
ALL_TRACES = sorted(glob.glob(os.path.join(_G, "ref_*.npz")))


@pytest.mark.parametrize("path", ALL_TRACES, ids=[os.path.basename(p) for p in ALL_TRACES])
def test_device_env_step_vs_reference_trace(path, oracle_mod):
    """Every env kind on the device (through the C ABI), teacher-forced along every reference trace: observation /
    reward / done per step within 1e-3 / 1e-2 of the RECORDED values (Cassie 1e-2 / 2e-3), zero unexplained outliers:
    a step outside the tolerance must show different constraint-row / contact counts on oracle and device, a
    numerically singular mass matrix or an unstable step map, and stay within the bound that explanation gives
    (tests/teacher.py); the device's integer bookkeeping is read back after every structurally identical step and
    compared exactly with the oracle's, Monkey3D grab steps included."""
    from tests import teacher as T

    j = T.run_golden_trace(oracle_mod, path, "gpu")
    assert j.book_checked >= 0.6 * j.n


def test_device_follows_config1_trace(oracle_mod):
    """BASELINE configs[0]: the 1000-step single-env random-action trace (recorded from the oracle, labelled
    "restatement": tools/gen_config1_trace.py) teacher-forced on the device, explained-or-fatal."""
    from tests import teacher as T

    j = T.run_golden_trace(oracle_mod, os.path.join(_G, "restatement_walker3d_custom_config1.npz"), "gpu")
    assert j.n == 1000 and j.book_checked >= 900
