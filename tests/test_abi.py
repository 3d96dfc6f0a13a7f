"""The C-ABI library loads and exports every symbol include/mocca_b200.h declares; no CPU fallback exists."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_header_symbols_exported():
    from mocca_envs_b200 import _lib

    _lib.build()
    L = C.CDLL(_lib.LIB_PATH)
    hdr = open(os.path.join(ROOT, "include", "mocca_b200.h")).read()
    declared = set(re.findall(r"\b(mb200_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    for sym in declared:
        assert hasattr(L, sym), sym
    assert declared == set(_lib.SYMBOLS)


def test_create_fails_loudly_without_gpu():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from mocca_envs_b200 import _lib

    L = _lib.lib()
    h = C.c_void_p()
    rc = L.mb200_create(b"Walker3DCustomEnv-v0", 4, 0, None, C.byref(h))
    assert rc != 0 and h.value is None
    assert len(L.mb200_last_error()) > 0
    rc = L.mb200_create(b"NoSuchEnv-v0", 4, 0, None, C.byref(h))
    assert rc != 0 and b"unsupported env id" in L.mb200_last_error()


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "mocca_envs_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("CPU oracle", "").replace("the oracle", "").lower() or f in (
                    "seeding.py",), (f, "product code must not reference oracle/")


def test_generated_header_is_current():
    """csrc/generated/walker3d_model.h matches what codegen emits from the committed JSON table."""
    from mocca_envs_b200 import codegen
    from mocca_envs_b200.model_compiler import load_table

    t = load_table(os.path.join(ROOT, "mocca_envs_b200", "models", "walker3d.json"))
    want = codegen.emit_header(t, "W3D")
    have = open(os.path.join(ROOT, "mocca_envs_b200", "csrc", "generated", "walker3d_model.h")).read()
    assert want == have


def test_model_table_matches_reference_files():
    ref = "/root/reference/mocca_envs/data"
    if not os.path.isdir(ref):
        pytest.skip("reference data not present on this box")
    from mocca_envs_b200 import model_compiler as mc

    t = mc.compile_walker3d(ref)
    have = mc.load_table(os.path.join(ROOT, "mocca_envs_b200", "models", "walker3d.json"))
    import json

    assert json.loads(json.dumps(t, default=float)) == have
