"""Times one finest-level smooth of the benchmark workload with the lean pair sweeps, the generation-3 fused pass and the
generation-4 (bulk-async-copy staged) fused pass for every compiled (tile_y, EARLY, LATE) plan.  One line per variant.

  python scripts/tune_fused4.py [n_cell] [reps] [max_grid_size]
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402
import amrex_b200 as ab  # noqa: E402
from common import synth_abeclap  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
mgs = int(sys.argv[3]) if len(sys.argv) > 3 else (128 if n >= 128 else n)
ab.init(0)
P = synth_abeclap(ab, n, mgs, fusion=0)
op = P["op"]
op.prepareForSolve()
x = op.make(0, 0, 1)
b = op.make(0, 0, 0)
b.copy_from(P["rhs"])
stream = torch.cuda.ExternalStream(ab.lib.amrex_b200_stream(), device=torch.device("cuda", 0))
cells = n ** 3


def time_smooth():
    x.setVal(0.0, ng=1)
    for _ in range(3):
        op.smooth(0, 0, x, b)
    ab.lib.amrex_b200_synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(reps):
        op.smooth(0, 0, x, b)
    e1.record(stream)
    ab.lib.amrex_b200_synchronize()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def kernel_ms():
    ab.profile_enable(True)
    for _ in range(reps):
        op.smooth(0, 0, x, b)
    rep = ab.profile_report()
    ab.profile_enable(False)
    names = sorted(set(q[0] for q in rep))
    return {nm: sum(q[3] for q in rep if q[0] == nm) / sum(q[2] for q in rep if q[0] == nm) for nm in names}


op.setSmootherFusion(0)
ab.lib.b200mg_set_gsrb_lean_occupancy(4)
ms = time_smooth()
km = kernel_ms()
ref = x.norm0()
k = km.get("b200mg_gsrb_abec_pairs_lean")
print(json.dumps(dict(variant="lean pair sweeps", ms_per_smooth=ms, kernel_ms=k, kernel_gbs_44=44.0 * cells / (k * 1e-3) / 1e9)), flush=True)

op.setSmootherFusion(1)
plans = [(3, 8, 0, 0, 0), (4, 8, 4, 2, 0), (4, 8, 4, 2, 1), (4, 8, 4, 3, 1), (4, 6, 5, 3, 1), (4, 6, 4, 2, 1)]   # (version, tile_y, EARLY, LATE, decoupled)
if os.environ.get("TUNE_PLANS"):
    plans = [tuple(int(v) for v in g.split(",")) for g in os.environ["TUNE_PLANS"].split(";")]
for ver, ty, se, sl, dec in plans:
    op.setFusedVersion(ver)
    op.setFusedMinBoxCells(32 ** 3)
    if ver >= 4:
        assert ab.lib.amrex_b200_set_fused4_plan(ty, se, sl) == 0
        ab.lib.b200mg_set_gsrb4_sync(dec)
    else:
        op.setFusedPlan(ty, 128, 0)
    ms = time_smooth()
    km = kernel_ms()
    k = km.get("b200mg_gsrb4") if ver >= 4 else km.get("b200mg_gsrb3")
    sh = km.get("b200mg_gsrb_shell_abec")
    print(json.dumps(dict(variant=f"fused v{ver} ty={ty} stages={se},{sl} decoupled={dec}", ms_per_smooth=ms, kernel_ms=k, shell_ms=sh,
                          kernel_gbs_56=56.0 * cells / (k * 1e-3) / 1e9 if k else None, same_norm_as_pair=bool(x.norm0() == ref),
                          kernels={a: round(v, 4) for a, v in km.items()})), flush=True)
