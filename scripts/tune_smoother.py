"""Times one finest-level smooth (red+black, incl. halo refresh and BCs) of the benchmark workload under every smoother
variant in ONE process: the reference schedule with the pair colour kernel, and the fused pass for a grid of
(tile_y, chunk_z, L2-prefetch distance).  Prints one line per variant: ms per smooth and the bandwidth figures.

  python scripts/tune_smoother.py [n_cell] [reps]
"""
import itertools
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
ab.init(0)
P = synth_abeclap(ab, n, 128 if n >= 128 else n, fusion=0)
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


def kernel_ms(name):
    """average duration of kernel `name` on the finest level from the per-launch CUDA events of the launch layer"""
    ab.profile_enable(True)
    for _ in range(reps):
        op.smooth(0, 0, x, b)
    rep = ab.profile_report()
    ab.profile_enable(False)
    if name is None:
        names = sorted(set(q[0] for q in rep))
        return {nm: sum(q[3] for q in rep if q[0] == nm) / sum(q[2] for q in rep if q[0] == nm) for nm in names}
    r = [q for q in rep if q[0] == name]
    return (sum(q[3] for q in r) / sum(q[2] for q in r)) if r else None


out = []
op.setSmootherFusion(0)
ref = None
for minb, label in ((0, "generic pair colour sweeps"), (4, "lean pair sweeps, 4 CTAs/SM (64 regs)"), (3, "lean pair sweeps, 3 CTAs/SM (80 regs)")):
    ab.lib.b200mg_set_gsrb_lean_occupancy(minb)
    ms = time_smooth()
    k = kernel_ms("b200mg_gsrb_abec_pairs_lean")
    nrm = x.norm0()
    ref = nrm if ref is None else ref
    out.append(dict(variant=label, ms_per_smooth=ms, kernel_ms=k, kernel_gbs_44=44.0 * cells / (k * 1e-3) / 1e9 if k else None,
                    same_norm=bool(nrm == ref)))
    print(json.dumps(out[-1]), flush=True)
if os.environ.get("TUNE_SKIP_FUSED"):
    sys.exit(0)

op.setSmootherFusion(1)
ab.lib.b200mg_set_gsrb_lean_occupancy(4)
grid = [(3, 4, 128), (3, 8, 128), (3, 12, 128), (3, 8, 32), (3, 8, 64), (2, 8, 128)]
if os.environ.get("TUNE_GRID"):
    grid = [tuple(int(v) for v in g.split(",")) for g in os.environ["TUNE_GRID"].split(";")]
for ver, ty, cz in grid:
    op.setFusedVersion(ver)
    op.setFusedPlan(ty, cz, 0)
    ms = time_smooth()
    km = kernel_ms(None)
    k = km.get("b200mg_gsrb3" if ver >= 3 else "b200mg_gsrb2_abec")
    sh = km.get("b200mg_gsrb_shell_abec")
    same = (x.norm0() == ref)
    out.append(dict(variant=f"fused v{ver} ty={ty} cz={cz}", ms_per_smooth=ms, kernel_ms=k, shell_ms=sh,
                    kernel_gbs_56=56.0 * cells / (k * 1e-3) / 1e9 if k else None, same_norm_as_pair=bool(same)))
    print(json.dumps(out[-1]), flush=True)
