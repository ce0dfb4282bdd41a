"""Times GMRES preconditioned by one MLMG V-cycle against plain MLMG V-cycles on the benchmark operator (variable-coefficient
MLABecLaplacian, max_grid_size 128) to the same relative tolerance.  One JSON line per solver.

  python scripts/gmres_vs_mlmg.py [n_cell]
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import amrex_b200 as ab  # noqa: E402
from common import synth_abeclap  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
ab.init(0)
P = synth_abeclap(ab, n, min(128, n))
mlmg = ab.MLMG(P["op"])
mlmg.setVerbose(0)


def timed(fn, reps=2):
    best = None
    for _ in range(reps):
        P["sol"].copy_from(P["sol0"], ng=1)
        ab.lib.amrex_b200_synchronize()
        t0 = time.perf_counter()
        fn()
        ab.lib.amrex_b200_synchronize()
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return best


t = timed(lambda: mlmg.solve([P["sol"]], [P["rhs"]], 1e-10, 0.0))
ref = P["sol"].norm0()
print(json.dumps(dict(solver="MLMG V-cycles", n_cell=n, seconds=t, iters=mlmg.numIters(), norm0=ref)), flush=True)
gm = ab.GMRESMLMG(mlmg)
gm.setVerbose(0)
t = timed(lambda: gm.solve(P["sol"], P["rhs"], 1e-10, 0.0))
print(json.dumps(dict(solver="GMRES + 1 V-cycle", n_cell=n, seconds=t, iters=gm.numIters(), status=gm.status(),
                      resid_2norm=gm.residualNorm(), norm0=P["sol"].norm0(), rel_diff_norm0=abs(P["sol"].norm0() - ref) / ref)), flush=True)
