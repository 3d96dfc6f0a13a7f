import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def walker_table():
    from mocca_envs_b200.model_compiler import load_table

    return load_table(os.path.join(ROOT, "mocca_envs_b200", "models", "walker3d.json"))


@pytest.fixture(scope="session")
def oracle_mod():
    from oracle import oracle as O

    O.build()
    return O


@pytest.fixture(scope="session")
def monkey_table():
    from mocca_envs_b200.model_compiler import load_table

    return load_table(os.path.join(ROOT, "mocca_envs_b200", "models", "monkey3d.json"))


@pytest.fixture(scope="session")
def child_table():
    from mocca_envs_b200.model_compiler import load_table

    return load_table(os.path.join(ROOT, "mocca_envs_b200", "models", "child3d.json"))


@pytest.fixture(scope="session")
def mike_table():
    from mocca_envs_b200.model_compiler import load_table

    return load_table(os.path.join(ROOT, "mocca_envs_b200", "models", "mike.json"))


@pytest.fixture(scope="session")
def walker2d_table():
    from mocca_envs_b200.model_compiler import load_table

    return load_table(os.path.join(ROOT, "mocca_envs_b200", "models", "walker2d.json"))


@pytest.fixture(scope="session")
def crab2d_table():
    from mocca_envs_b200.model_compiler import load_table

    return load_table(os.path.join(ROOT, "mocca_envs_b200", "models", "crab2d.json"))


@pytest.fixture(scope="session")
def cassie_table():
    from mocca_envs_b200.model_compiler import load_table

    return load_table(os.path.join(ROOT, "mocca_envs_b200", "models", "cassie.json"))
