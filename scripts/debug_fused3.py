"""Element-wise comparison of one smooth: fused pass generation 3 vs the pair colour sweeps."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import amrex_b200 as ab  # noqa: E402
from common import build_problem, run_ref  # noqa: E402

ab.init(0)
for prob, n, mgs in ((1, 128, 64), (2, 64, 32), (1, 64, 32), (2, 128, 64)):
    ref, dump = run_ref(dump=True, mode="solve", prob_type=prob, n_cell=n, max_grid_size=mgs, linop_maxorder=2, agg_grid_size=32)
    res = {}
    for ver in (0, 2, 3):
        P = build_problem(ab, prob, n, mgs, dump, maxorder=2, fusion=0 if ver == 0 else 1)
        op = P["op"]
        if ver:
            op.setFusedVersion(ver)
        op.prepareForSolve()
        for mglev in (0, 1):
            x = op.make(0, mglev, 1)
            b = op.make(0, mglev, 0)
            nn = n >> mglev
            rng = np.random.default_rng(5 + mglev)
            b.upload(rng.standard_normal((nn, nn, nn)), (0, 0, 0))
            x.setVal(0.0, ng=1)
            x.upload(rng.standard_normal((nn, nn, nn)), (0, 0, 0))
            op.smooth(0, mglev, x, b)
            op.smooth(0, mglev, x, b)
            res[(ver, mglev)] = x.download((0, 0, 0), (nn, nn, nn))
    for mglev in (0, 1):
        for ver in (2, 3):
            d = np.abs(res[(ver, mglev)] - res[(0, mglev)])
            bad = np.argwhere(d > 0)
            print(f"prob {prob} n {n} mgs {mgs} mglev {mglev} fused v{ver}: max|diff| {d.max():.3e}, differing cells {len(bad)}",
                  ("first: " + str(bad[:8].tolist()) + " ... last: " + str(bad[-4:].tolist())) if len(bad) else "")
            if len(bad):
                for ax in range(3):
                    vals, cnt = np.unique(bad[:, ax] % (mgs >> mglev), return_counts=True)
                    print(f"   axis {ax} (index mod box size) histogram:", dict(zip(vals.tolist()[:12], cnt.tolist()[:12])))
