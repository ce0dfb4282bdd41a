"""Developer check run on the GPU box: solve a few configs with verbose output and compare with the reference."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import amrex_b200 as ab
from common import build_problem, run_ref, rel_maxdiff

ab.init(0)
cases = [(1, 32, 16, 2), (1, 128, 64, 2), (2, 64, 32, 2), (2, 128, 64, 3), (5, 64, 32, 2), (2, 256, 128, 2)]
if len(sys.argv) > 1:
    cases = cases[:int(sys.argv[1])]
for prob, n, mgs, mo in cases:
    ref, dump = run_ref(dump=True, mode="solve", prob_type=prob, n_cell=n, max_grid_size=mgs, linop_maxorder=mo, agg_grid_size=32)
    for fusion in (0, 1):
        try:
            P = build_problem(ab, prob, n, mgs, dump, maxorder=mo, fusion=fusion)
            mlmg = ab.MLMG(P["op"]); mlmg.setVerbose(1 if fusion == 0 else 0); mlmg.setMaxIter(30)
            t0 = time.time()
            ab.lib.amrex_b200_reset_launch_count()
            mlmg.solve([P["sol"]], [P["rhs"]], 1e-10, 0.0)
            dt = time.time() - t0
            mine = P["sol"].download((0, 0, 0), (n, n, n)); refv = dump["sol_lev0"][1][1:-1, 1:-1, 1:-1]
            if prob == 5: mine = mine - mine.mean(); refv = refv - refv.mean()
            print(f"CASE prob={prob} n={n} mgs={mgs} maxorder={mo} fusion={fusion}: iters {mlmg.numIters()} (ref {ref['iters']}) "
                  f"reldiff {rel_maxdiff(mine, refv):.3e} time {dt:.4f}s ref_time {ref['solve_times'][0]:.3f}s launches {ab.lib.amrex_b200_launch_count()} "
                  f"cg {mlmg.cgIters()[:4]} ref_cg {ref['cg_iters'][:4]}", flush=True)
            h, rh = mlmg.residualHistory(), ref["history"]
            print("   hist mine", ["%.6e" % x for x in h[:4]], "\n   hist ref ", ["%.6e" % x for x in rh[:4]], flush=True)
        except Exception as e:
            print(f"CASE prob={prob} n={n} fusion={fusion} FAILED: {e}", flush=True)
