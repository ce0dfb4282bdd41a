"""Element-wise comparison of two smooths: fused pass generation 4 (every compiled plan) vs the pair colour sweeps."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import amrex_b200 as ab  # noqa: E402
from common import build_problem, run_ref  # noqa: E402

ab.init(0)
PLANS = [(8, 4, 2, 0), (8, 4, 2, 1), (8, 4, 3, 1), (6, 5, 3, 1), (6, 4, 2, 1), (4, 4, 4, 1)]   # (tile_y, EARLY, LATE, decoupled warps)
cases = ((2, 64, 32), (1, 64, 32), (2, 128, 64), (2, 256, 128), (1, 96, 40))
if len(sys.argv) > 1:
    cases = cases[:int(sys.argv[1])]
nbad = 0
for prob, n, mgs in cases:
    ref, dump = run_ref(dump=True, mode="solve", prob_type=prob, n_cell=n, max_grid_size=mgs, linop_maxorder=2, agg_grid_size=32)
    res = {}
    for plan in [None] + PLANS:
        P = build_problem(ab, prob, n, mgs, dump, maxorder=2, fusion=0 if plan is None else 1)
        op = P["op"]
        if plan is not None:
            op.setFusedVersion(4)
            assert ab.lib.amrex_b200_set_fused4_plan(*plan[:3]) == 0
            ab.lib.b200mg_set_gsrb4_sync(plan[3])
            op.setFusedMinBoxCells(32 ** 3)
        op.prepareForSolve()
        for mglev in (0, 1):
            x = op.make(0, mglev, 1)
            b = op.make(0, mglev, 0)
            nn = n >> mglev
            rng = np.random.default_rng(5 + mglev)
            b.upload(rng.standard_normal((nn, nn, nn)), (0, 0, 0))
            x.setVal(0.0, ng=1)
            x.upload(rng.standard_normal((nn, nn, nn)), (0, 0, 0))
            ab.profile_enable(True)
            op.smooth(0, mglev, x, b)
            op.smooth(0, mglev, x, b)
            names = sorted(set(q[0] for q in ab.profile_report()))
            ab.profile_enable(False)
            res[(plan, mglev)] = (x.download((0, 0, 0), (nn, nn, nn)), names)
    for mglev in (0, 1):
        for plan in PLANS:
            got, names = res[(plan, mglev)]
            d = np.abs(got - res[(None, mglev)][0])
            bad = np.argwhere(d > 0)
            nbad += len(bad)
            used = "gsrb4" if "b200mg_gsrb4" in names else ("gsrb3" if "b200mg_gsrb3" in names else "pairs")
            print(f"prob {prob} n {n} mgs {mgs} mglev {mglev} plan {plan} [{used}]: max|diff| {d.max():.3e}, differing cells {len(bad)}",
                  ("first: " + str(bad[:6].tolist()) + " ... last: " + str(bad[-3:].tolist())) if len(bad) else "", flush=True)
            if len(bad):
                for ax in range(3):
                    vals, cnt = np.unique(bad[:, ax] % (mgs >> mglev), return_counts=True)
                    print(f"   axis {ax} (index mod box size) histogram:", dict(zip(vals.tolist()[:12], cnt.tolist()[:12])))
ab.lib.b200mg_set_gsrb4_sync(0)
print("TOTAL differing cells", nbad)
sys.exit(1 if nbad else 0)
