"""Profiling driver: runs the finest-level primitives of the benchmark workload a few times so that ncu can capture them
(ncu --set full -k regex:<kernel> ... python scripts/prof_kernels.py [n_cell] [what ...])."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import amrex_b200 as ab  # noqa: E402
from common import synth_abeclap  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
what = sys.argv[2:] or ["smooth", "residual", "restrict", "interp", "vec"]
ab.init(0)
P = synth_abeclap(ab, n, 128 if n >= 128 else n, fusion=int(os.environ.get("FUSION", "1")))
op = P["op"]
op.prepareForSolve()
x = op.make(0, 0, 1)
b = op.make(0, 0, 0)
y = op.make(0, 0, 0)
b.copy_from(P["rhs"])
x.setVal(0.0, ng=1)
reps = int(os.environ.get("REPS", "3"))
import torch  # noqa: E402
ab.lib.amrex_b200_synchronize()
torch.cuda.cudart().cudaProfilerStart()   # ncu --profile-from-start off: capture only the loop below
for _ in range(reps):
    if "smooth" in what:
        op.smooth(0, 0, x, b)
    if "residual" in what:
        op.residual(0, 0, y, x, b)
    if "restrict" in what:
        c = op.make(0, 1, 0)
        op.restriction(0, 1, c, y)
        if "interp" in what:
            f = op.make(0, 0, 0)
            op.interp_add(0, 0, f, c)
    if "vec" in what:
        y.copy_from(b)
        y.norm0()
        y.dot(b)
ab.lib.amrex_b200_synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("done", x.norm0())
