#!/usr/bin/env python
"""Diagnostic: per-step observation errors of the teacher-forced device-vs-oracle run (tests/teacher.py) with the
verdicts switched off -- prints percentiles and the largest steps.  usage: dbg_teacher_errs.py <env name> [seeds] [steps]
(MB200_LIB selects the library variant)"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from oracle import oracle as O
from tests import teacher as T

env = sys.argv[1]; seeds = int(sys.argv[2]) if len(sys.argv) > 2 else 8; steps = int(sys.argv[3]) if len(sys.argv) > 3 else 25
log = []
def step(self, t, e_obs, rew, ref_rew, done, ref_done, rows, ref_rows, nc, ref_nc, cond, sens=None):
    log.append((self.name, t, e_obs, int(rows) == int(ref_rows) and int(nc) == int(ref_nc), cond))
    return int(rows) == int(ref_rows) and int(nc) == int(ref_nc)
T.Judge.step = step
T.Judge.book = lambda *a, **k: None
T.Judge.finish = lambda self, median_below=None: self
if env == "cassie":
    fn = lambda rng, k: (0.3 if (k // 5) % 2 == 0 else 0.6) * rng.uniform(-1, 1, 10)
else:
    A = T.table_of(T.SPECS[env][1])["n_dof"]
    fn = lambda rng, k: rng.uniform(-1, 1, A)
T.run_vs_oracle(O, env, "gpu", range(seeds), steps, fn)
e = np.array([x[2] for x in log if x[3]])
print(os.environ.get("MB200_LIB", "cur"), env, "steps", len(log), "same-structure", len(e),
      "median %.2e p90 %.2e p99 %.2e max %.2e" % (np.median(e), np.percentile(e, 90), np.percentile(e, 99), e.max()))
for x in sorted(log, key=lambda x: -x[2])[:6]:
    print("   ", x)
