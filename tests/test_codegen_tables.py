"""Structural invariants of the generated model tables the round-2 kernels rely on (mocca_envs_b200/codegen.py), checked
for every committed model independently of the generator's own assertions:

  * affine factorisation addressing: for pivot row k, slot t of its support is coordinate t + (column steps) and the
    compact row that slot updates starts at word t (t + 1) / 2 + (offset steps) -- the packed-triangle property
    Sim::pivot / row_build / mass_matrix_and_rhs address through instead of per-slot tables;
  * the chain-walk table: every joint is stored by exactly one (step, chain) record, a record's joint is the child of the
    previous step's, and the axis is a unit vector;
  * the body forest: accumulating the records leaf to root makes record bstart(j) the sum over joint j's subtree."""
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MODELS = ["walker3d", "monkey3d", "cassie", "child3d", "mike", "walker2d", "crab2d"]


def _reduced(name):
    from mocca_envs_b200 import codegen
    from mocca_envs_b200.model_compiler import load_table

    return codegen.reduce_table(load_table(os.path.join(ROOT, "mocca_envs_b200", "models", name + ".json")))


def _header(name):
    return open(os.path.join(ROOT, "mocca_envs_b200", "csrc", "generated", name + "_model.h")).read()


def _iarr(h, pattern):
    m = re.search(pattern + r"\[\d+\] = \{([^}]*)\}", h)
    assert m, pattern
    return [int(x) for x in m.group(1).split(",")]


@pytest.mark.parametrize("name", MODELS)
def test_affine_addressing_reproduces_the_compact_layout(name):
    r, h = _reduced(name), _header(name)
    nu, rowoff, rowlen, mask = r["nu"], r["rowoff"], r["rowlen"], r["rowmask_rt"]
    assert _iarr(h, r"static constexpr int k_rowoff") == rowoff and _iarr(h, r"static constexpr int k_rowlen") == rowlen
    tri = lambda t: t * (t + 1) // 2
    kt = {(nm, i): _iarr(h, r"static constexpr int k_%s%d" % (nm, i)) for nm in ("ft", "fd", "fc") for i in (1, 2)}
    rt = {(nm, i): _iarr(h, r"MB_TABLE int \w+_%s%d" % (nm, i)) for nm in ("ft", "fd", "fc") for i in (1, 2)}
    for k in range(nu):
        cols = sorted(j for j in range(k) if (mask[k] >> j) & 1)
        assert len(cols) == rowlen[k] - 1
        for tab, slots in ((kt, cols), (rt, cols + [k])):
            for t, col in enumerate(slots):
                dc = sum(tab[("fc", i)][k] for i in (1, 2) if t >= tab[("ft", i)][k])
                do = sum(tab[("fd", i)][k] for i in (1, 2) if t >= tab[("ft", i)][k])
                assert t + dc == col, (name, k, t)
                assert tri(t) + do == rowoff[col], (name, k, t)
    # the pair index p = t (t + 1) / 2 + s of a pivot's update never leaves the L block by more than the documented slack
    assert r["lsize"] == sum(rowlen) and max(rowlen) <= 14


@pytest.mark.parametrize("name", MODELS)
def test_chain_walk_table_covers_every_joint_once(name):
    r, h = _reduced(name), _header(name)
    m = re.search(r"MbKinRec \w+_kin\[(\d+)\]\[(\d+)\] = \{\n(.*?)\};", h, re.S)
    nlev, nch1 = int(m.group(1)), int(m.group(2))
    recs = re.findall(r"\{(\d+), \{([^}]*)\}, \{([^}]*)\}, (\d)\}", m.group(3))
    assert len(recs) == nlev * nch1 and nlev == r["nlevel"] and 3 * (nch1 - 1) <= 30
    stored = []
    for ch in range(nch1):
        for st in range(nlev):
            j, off, ax, store = recs[st * nch1 + ch]
            j, store = int(j), int(store)
            ax = np.array([float(x.rstrip("f")) for x in ax.split(",")])
            assert abs(np.linalg.norm(ax) - 1.0) < 1e-6
            if store:
                stored.append(j)
            if ch == nch1 - 1:
                assert store == 0  # the idle chain of the unused lanes
    assert sorted(stored) == list(range(r["nj"]))
    # a stored joint's parent is the joint of the previous step of the same chain
    for ch in range(nch1 - 1):
        for st in range(nlev):
            j, _, _, store = recs[st * nch1 + ch]
            if int(store):
                pj = r["jparent"][int(j)]
                assert (st == 0 and pj < 0) or (st > 0 and int(recs[(st - 1) * nch1 + ch][0]) == pj), (name, ch, st)


@pytest.mark.parametrize("name", MODELS)
def test_body_forest_accumulates_to_subtree_sums(name):
    r, h = _reduced(name), _header(name)
    par = _iarr(h, r"static constexpr int k_bparent")
    nb = r["nb"]
    assert len(par) == nb and par[0] == 0 and all(par[b] < b for b in range(1, nb))
    rng = np.random.RandomState(0)
    rec = rng.rand(nb, 16)
    acc = rec.copy()
    for b in range(nb - 1, 0, -1):  # Sim::composites
        acc[par[b]] += acc[b]
    np.testing.assert_allclose(acc[0], rec.sum(0), rtol=1e-12)
    for j in range(r["nj"]):
        np.testing.assert_allclose(acc[r["bstart"][j]], rec[r["bstart"][j]:r["bend"][j]].sum(0), rtol=1e-12)
