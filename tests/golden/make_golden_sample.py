"""Generates tests/golden/sol_sample_p2_n512_g128.npz: the reference's solution of the full-size benchmark workload (512^3
variable-coefficient MLABecLaplacian, max_grid_size 128, tol 1e-10) sampled at every 8th cell (offset 3) -> 64^3 values,
plus the sampled analytic solution.  bench.py compares its own solution with it at every N (parity_vs_reference).
Run in the build container:  python tests/golden/make_golden_sample.py   (needs ~25 GB of RAM + 10 GB in /tmp, ~3 min)"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from common import run_ref  # noqa: E402

STRIDE, OFFSET = 8, 3


def main():
    n, mgs = 512, 128
    ref, dump = run_ref(dump=True, mode="solve", prob_type=2, n_cell=n, max_grid_size=mgs, linop_maxorder=2, agg_grid_size=32)
    lo, sol = dump["sol_lev0"]
    sol = sol[1:-1, 1:-1, 1:-1]
    _, exact = dump["exact_lev0"]
    s = (slice(OFFSET, None, STRIDE),) * 3
    np.savez_compressed(os.path.join(HERE, f"sol_sample_p2_n{n}_g{mgs}.npz"), sol=np.ascontiguousarray(sol[s]),
                        exact=np.ascontiguousarray(exact[s]), stride=STRIDE, offset=OFFSET, iters=ref["iters"],
                        history=np.array(ref["history"]), solmax=float(np.max(np.abs(sol))))
    print("iters", ref["iters"], "history", ref["history"], "sample", sol[s].shape)


if __name__ == "__main__":
    main()
