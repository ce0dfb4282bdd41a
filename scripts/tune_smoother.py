"""A/B timing of the finest-level smoother variants on the benchmark workload (run on the GPU box):
    python scripts/tune_smoother.py [n_cell] [reps] [variant ...]
variant = "ty,early,late,pairs" (cell pairs per thread on rows of more than 64 cells: 0 by launch size, 1 = generation 4,
2 = generation 5; default: a few of the compiled plans).  Prints one JSON line per variant: time of one
smooth and of its kernels (CUDA events around every launch, amrex_b200 profile report), and whether the smoothed field has
the bits of the first variant."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
import amrex_b200 as ab  # noqa: E402
from common import synth_abeclap  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
variants = [tuple(int(v) for v in a.split(",")) for a in sys.argv[3:]] or [(8, 4, 2, 2), (8, 4, 2, 1), (8, 4, 2, 0)]
mgs = 128 if n >= 128 else n
ab.init(0)
first = None
for var in variants:
    assert ab.lib.amrex_b200_set_fused4_plan(*var[:3]) == 0, var
    ab.lib.b200mg_set_gsrb4_sync(var[3])
    P = synth_abeclap(ab, n, mgs, fusion=1)
    op = P["op"]
    op.prepareForSolve()
    x = op.make(0, 0, 1)
    b = op.make(0, 0, 0)
    b.copy_from(P["rhs"])
    x.setVal(0.0, ng=1)
    for _ in range(3):
        op.smooth(0, 0, x, b)
    ab.lib.amrex_b200_synchronize()
    ab.profile_enable(True)
    for _ in range(reps):
        op.smooth(0, 0, x, b)
    rep = ab.profile_report()
    ab.profile_enable(False)
    kern = {}
    for name, scope, launches, total, tmin, tmax in rep:
        kern[name] = kern.get(name, 0.0) + total / reps
    cells = float(n) ** 3
    k = kern.get("b200mg_gsrb4", 0.0)
    line = {"variant": "ty=%d early=%d late=%d pairs=%d" % var, "n": n, "ms_per_smooth": round(sum(kern.values()), 4),
            "gsrb4_ms": round(k, 4), "gsrb4_gbs_56": round(56.0 * cells / (k * 1e-3) / 1e9, 1) if k > 0 else None,
            "kernels": {q: round(v, 4) for q, v in sorted(kern.items())}}
    if n <= 256:
        got = x.download((0, 0, 0), (n, n, n))
        if first is None:
            first = got
        line["same_bits_as_first"] = bool(np.array_equal(got, first))
    line["norm0"] = x.norm0()
    print(json.dumps(line), flush=True)
    del x, b, op, P
