"""Stage-by-stage comparison of the 2-level composite solve with the reference (oracle/_ref/ref_driver mode=amr)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import amrex_b200 as ab  # noqa: E402
from common import build_problem_amr, run_ref  # noqa: E402

prob, n, mgs, mo = [int(v) for v in (sys.argv[1:5] if len(sys.argv) >= 5 else (1, 32, 16, 3))]
ab.init(0)
ref, dump = run_ref(dump=True, mode="amr", prob_type=prob, n_cell=n, max_grid_size=mgs, linop_maxorder=mo, agg_grid_size=32, max_level=1)
print("reference resid after 1, 2 iterations:", ref["resid_after_iter"])


def cmp(name, mf, key, ng=0):
    lo, r = dump[key]
    mine = mf.download(tuple(lo), r.shape)
    d = np.abs(mine - r)
    idx = np.unravel_index(np.argmax(d), d.shape)
    print(f"{name}: max|diff| {d.max():.3e} (ref max {np.abs(r).max():.3e}) at index {tuple(int(i) + int(l) for i, l in zip(idx, lo))} "
          f"mine {mine[idx]:.6e} ref {r[idx]:.6e}; cells off by >1e-9 rel: {int((d > 1e-9 * np.abs(r).max()).sum())}")
    return d


P = build_problem_amr(ab, prob, n, mgs, dump, max_level=1, maxorder=mo)
mlmg = ab.MLMG(P["op"])
mlmg.setVerbose(0)
res = [ab.MultiFab(P["ba"][l], P["dm"][l], 1, 0) for l in range(2)]
mlmg.compResidual(res, P["sol"], P["rhs"])
for l in range(2):
    d = cmp(f"compResidual lev{l}", res[l], f"amr_res_lev{l}")
    if l == 0 and d.max() > 1e-9:
        bad = np.argwhere(d > 1e-9 * max(1.0, d.max()))
        print("  first bad cells (array index):", bad[:12].tolist(), "count", len(bad))

for nit in (1, 2):
    P = build_problem_amr(ab, prob, n, mgs, dump, max_level=1, maxorder=mo)
    mlmg = ab.MLMG(P["op"])
    mlmg.setVerbose(0)
    mlmg.setFixedIter(nit)
    try:
        mlmg.solve(P["sol"], P["rhs"], 1e-10, 0.0)
    except Exception as e:  # noqa: BLE001
        print("solve raised:", e)
    print(f"my residual history after {nit} fixed iteration(s):", mlmg.residualHistory())
    for l in range(2):
        cmp(f"sol after {nit} iter lev{l}", P["sol"][l], f"amr_sol{nit}_lev{l}")
