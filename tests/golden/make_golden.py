"""Generates the committed fixtures under tests/golden/ by running the UNMODIFIED reference (oracle/_ref/ref_driver,
built by oracle/Makefile from /root/reference).  Run in the build container:  python tests/golden/make_golden.py [--big]

  meta_*.json   MG hierarchy (boxes, owner ranks, domains per amr x mg level), FillBoundary LocTags (cross and full
                stencil) and SFC processor maps for 1/2/3/4/8 ranks             -> bit-exact integer parity targets
  solve_*.json  iteration count, residual history, initial norms, bottom-solver iteration counts, max-norm error against
                the analytic solution                                           -> solver parity targets
  prim_*.npz    inputs and outputs of single primitives (apply, smooth, residual, restriction, prolongation) on a
                16^3 / 8^3-box problem                                          -> kernel parity targets (fp64)
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from common import run_ref  # noqa: E402

META = [  # (name, kwargs)
    ("p1_n128_g64", dict(prob_type=1, n_cell=128, max_grid_size=64)),
    ("p1_n64_g32", dict(prob_type=1, n_cell=64, max_grid_size=32)),
    ("p2_n256_g128", dict(prob_type=2, n_cell=256, max_grid_size=128)),
    ("p2_n512_g128", dict(prob_type=2, n_cell=512, max_grid_size=128)),
    ("p5_n64_g32", dict(prob_type=5, n_cell=64, max_grid_size=32)),
    ("p5_n128_g32", dict(prob_type=5, n_cell=128, max_grid_size=32)),
    ("p1_n96_g40", dict(prob_type=1, n_cell=96, max_grid_size=40)),
    ("p1_n128_g64_lev1", dict(prob_type=1, n_cell=128, max_grid_size=64, max_level=1)),
    ("p2_n64_g32_lev1", dict(prob_type=2, n_cell=64, max_grid_size=32, max_level=1)),
]
SOLVE = [
    ("p1_n128_g64", dict(prob_type=1, n_cell=128, max_grid_size=64, linop_maxorder=2)),
    ("p1_n64_g32", dict(prob_type=1, n_cell=64, max_grid_size=32, linop_maxorder=2)),
    ("p1_n64_g32_smoother", dict(prob_type=1, n_cell=64, max_grid_size=32, linop_maxorder=2, bottom="smoother")),
    ("p1_n64_g32_cg", dict(prob_type=1, n_cell=64, max_grid_size=32, linop_maxorder=2, bottom="cg")),
    ("p2_n64_g32", dict(prob_type=2, n_cell=64, max_grid_size=32, linop_maxorder=2)),
    ("p2_n64_g32_mo3", dict(prob_type=2, n_cell=64, max_grid_size=32, linop_maxorder=3)),
    ("p2_n128_g64", dict(prob_type=2, n_cell=128, max_grid_size=64, linop_maxorder=2)),
    ("p2_n256_g128", dict(prob_type=2, n_cell=256, max_grid_size=128, linop_maxorder=2)),
    ("p5_n64_g32", dict(prob_type=5, n_cell=64, max_grid_size=32, linop_maxorder=2)),
    ("p1_n128_g64_lev1", dict(prob_type=1, n_cell=128, max_grid_size=64, linop_maxorder=2, max_level=1)),
    ("p2_n64_g32_lev1_mo3", dict(prob_type=2, n_cell=64, max_grid_size=32, linop_maxorder=3, max_level=1)),
    ("p2_n128_g64_lev1_mo3", dict(prob_type=2, n_cell=128, max_grid_size=64, linop_maxorder=3, max_level=1)),
]
# the options of section 8f: GMRES preconditioned by MLMG, damped Jacobi smoother, inhomogeneous Neumann data (prob_type 3:
# the fields of problem 2 with d(phi)/dn prescribed on every face), F-cycles
SOLVE += [
    ("p2_n64_g32_gmres", dict(prob_type=2, n_cell=64, max_grid_size=32, linop_maxorder=2, use_gmres=1)),
    ("p1_n64_g32_gmres", dict(prob_type=1, n_cell=64, max_grid_size=32, linop_maxorder=2, use_gmres=1)),
    ("p2_n64_g32_jacobi", dict(prob_type=2, n_cell=64, max_grid_size=32, linop_maxorder=2, gauss_seidel=0)),
    ("p1_n64_g32_jacobi", dict(prob_type=1, n_cell=64, max_grid_size=32, linop_maxorder=2, gauss_seidel=0)),
    ("p3_n64_g32", dict(prob_type=3, n_cell=64, max_grid_size=32, linop_maxorder=2)),
    ("p2_n64_g32_fmg2", dict(prob_type=2, n_cell=64, max_grid_size=32, linop_maxorder=2, max_fmg_iter=2)),
    ("p1_n64_g32_fmg2", dict(prob_type=1, n_cell=64, max_grid_size=32, linop_maxorder=2, max_fmg_iter=2)),
    # level-by-level solves: each fine level is a single-level operator with setCoarseFineBC data from the level below
    ("p2_n64_g32_lev1_levelsolve", dict(prob_type=2, n_cell=64, max_grid_size=32, linop_maxorder=2, max_level=1, composite_solve=0)),
    ("p1_n64_g32_lev1_levelsolve", dict(prob_type=1, n_cell=64, max_grid_size=32, linop_maxorder=2, max_level=1, composite_solve=0)),
]
BIG = [("p2_n512_g128", dict(prob_type=2, n_cell=512, max_grid_size=128, linop_maxorder=2)),
       # BASELINE configs 4 and 5 at the sizes bench.py runs on one GPU: two-level composite solve with a 256^3 base, and the
       # fully periodic Poisson problem at 512^3 (one GPU's share of the 1024^3 weak-scaling run)
       ("p2_n256_g128_lev1_mo3", dict(prob_type=2, n_cell=256, max_grid_size=128, linop_maxorder=3, max_level=1)),
       ("p5_n512_g128", dict(prob_type=5, n_cell=512, max_grid_size=128, linop_maxorder=2)),
       ("p5_n256_g64", dict(prob_type=5, n_cell=256, max_grid_size=64, linop_maxorder=2))]
PRIM = [
    ("p2_n16_g8_m0", dict(prob_type=2, n_cell=16, max_grid_size=8, linop_maxorder=2, prim_mglev=0, agg_grid_size=4)),
    ("p1_n16_g8_m0", dict(prob_type=1, n_cell=16, max_grid_size=8, linop_maxorder=3, prim_mglev=0, agg_grid_size=4)),
    ("p2_n16_g8_m1", dict(prob_type=2, n_cell=16, max_grid_size=8, linop_maxorder=2, prim_mglev=1, agg_grid_size=4)),
]


def main():
    big = "--big" in sys.argv
    for name, kw in ([] if any(a.startswith("--only=") for a in sys.argv) else META):
        res, _ = run_ref(mode="meta", agg_grid_size=32, **kw)
        res["_args"] = dict(kw, agg_grid_size=32)
        json.dump(res, open(os.path.join(HERE, f"meta_{name}.json"), "w"), separators=(",", ":"))
        print("meta", name, [len(l) for l in res["hierarchy"]])
    only = [a.split("=", 1)[1] for a in sys.argv if a.startswith("--only=")]
    for name, kw in SOLVE + (BIG if big else []):
        if only and not any(o in name for o in only):
            continue
        res, _ = run_ref(mode="solve", agg_grid_size=32, **kw)
        res["_args"] = dict(kw, agg_grid_size=32)
        res.pop("solve_times", None)
        json.dump(res, open(os.path.join(HERE, f"solve_{name}.json"), "w"), indent=0)
        print("solve", name, res["iters"], res["err_inf"])
    for name, kw in ([] if any(a.startswith("--only=") for a in sys.argv) else PRIM):
        res, dump = run_ref(dump=True, mode="prim", **kw)
        arrays = {}
        for k, (lo, a) in dump.items():
            arrays[k] = a
            arrays[k + "__lo"] = np.array(lo, dtype=np.int64)
        arrays["_args"] = np.frombuffer(json.dumps(kw).encode(), dtype=np.uint8)
        np.savez_compressed(os.path.join(HERE, f"prim_{name}.npz"), **arrays)
        print("prim", name, sorted(dump.keys()))


if __name__ == "__main__":
    main()
