#!/usr/bin/env python
"""Regenerate the committed model tables from the reference's data files.

Reads  /root/reference/mocca_envs/data/robots/*.xml  (only available in the build container)
Writes mocca_envs_b200/models/<robot>.json            (committed; what the GPU box uses)
       mocca_envs_b200/csrc/generated/<robot>_model.h (committed; compile-time tables for the CUDA kernels)
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))

from mocca_envs_b200 import model_compiler as mc  # noqa: E402


def main(data_dir="/root/reference/mocca_envs/data"):
    out = os.path.join(os.path.dirname(HERE), "mocca_envs_b200", "models")
    os.makedirs(out, exist_ok=True)
    w = mc.compile_walker3d(data_dir)
    mc.save_table(w, os.path.join(out, "walker3d.json"))
    m = mc.compile_monkey3d(data_dir)
    mc.save_table(m, os.path.join(out, "monkey3d.json"))
    for fn, nm in ((mc.compile_child3d, "child3d"), (mc.compile_mike, "mike"), (mc.compile_walker2d, "walker2d"),
                   (mc.compile_crab2d, "crab2d")):
        t = fn(data_dir)
        mc.save_table(t, os.path.join(out, nm + ".json"))
        print("%-9s links=%d dof=%d mass=%.3f self pairs %d of %d" % (nm + ":", t["n_links"], t["n_dof"], t["total_mass"],
                                                                     len(t["self_pairs"]), t["self_pairs_candidates"]))
    from mocca_envs_b200 import urdf_compiler as uc

    c = uc.compile_cassie(data_dir)
    mc.save_table(c, os.path.join(out, "cassie.json"))
    print("cassie:   links=%d dof=%d mass=%.3f points=%d" % (c["n_links"], c["n_dof"], c["total_mass"], len(c["geoms"])))
    try:
        from mocca_envs_b200 import codegen
        codegen.emit_all(os.path.dirname(HERE))
    except ImportError:
        pass
    print("walker3d: links=%d dof=%d mass=%.3f" % (w["n_links"], w["n_dof"], w["total_mass"]))
    print("monkey3d: links=%d dof=%d mass=%.3f" % (m["n_links"], m["n_dof"], m["total_mass"]))


if __name__ == "__main__":
    main(*sys.argv[1:])
