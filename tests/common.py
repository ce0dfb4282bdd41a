"""Shared helpers for the parity tests.

`run_ref` executes oracle/_ref/ref_driver -- the UNMODIFIED reference (AMReX 24.10, CPU/OpenMP) built by oracle/Makefile,
which travels to the GPU box as a prebuilt binary.  Nothing here reads /root/reference.
"""
import json
import os
import subprocess
import tempfile

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_DRIVER = os.path.join(REPO, "oracle", "_ref", "ref_driver")
GOLDEN = os.path.join(REPO, "tests", "golden")


def have_ref():
    return os.path.exists(REF_DRIVER) and os.access(REF_DRIVER, os.X_OK)


def run_ref(dump=False, threads=None, **kw):
    """Run the reference driver; returns (result dict, dump dict name->(lo, ndarray[F])) ."""
    env = dict(os.environ)
    env["OMP_NUM_THREADS"] = str(threads or os.cpu_count() or 1)
    args = [REF_DRIVER] + [f"{k}={v}" for k, v in kw.items()]
    tmp = None
    if dump:
        tmp = tempfile.mkdtemp(prefix="refdump_")
        args.append(f"dump_dir={tmp}")
    out = subprocess.run(args, capture_output=True, text=True, env=env, cwd=tempfile.gettempdir(), timeout=1800)
    if out.returncode != 0:
        raise RuntimeError(f"ref_driver failed: {out.stdout[-2000:]}\n{out.stderr[-2000:]}")
    res = None
    for line in out.stdout.splitlines():
        if line.startswith("RESULT "):
            res = json.loads(line[7:])
    if kw.get("mode") == "meta":
        s = out.stdout
        res = json.loads(s[s.index("META") + 5:])
    d = load_dump(tmp) if dump else None
    if tmp:
        import shutil
        shutil.rmtree(tmp, ignore_errors=True)
    return res, d


def load_dump(path):
    man = json.load(open(os.path.join(path, "manifest.json")))
    out = {}
    for name, info in man.items():
        if name.startswith("_"):
            continue
        a = np.fromfile(os.path.join(path, name + ".bin"), dtype=np.float64).reshape(info["shape"], order="F")
        out[name] = (tuple(info["lo"]), a)
    return out


def build_problem(ab, prob_type, n_cell, max_grid_size, dump, maxorder=2, agg_grid_size=-1, nprocs=None, fusion=None,
                  max_coarsening_level=30):
    """Create geometry / grids / operator for one AMR level from a reference dump (bit-identical inputs)."""
    per = 1 if prob_type == 5 else 0
    ab.Geometry.setup((0., 0., 0.), (1., 1., 1.), (per, per, per))
    geom = ab.Geometry((0, 0, 0), (n_cell - 1,) * 3)
    ba = ab.BoxArray((0, 0, 0), (n_cell - 1,) * 3).maxSize(max_grid_size)
    dm = ab.DistributionMapping(ba) if nprocs is None else ab.DistributionMapping(ba, nprocs=nprocs)
    sol = ab.MultiFab(ba, dm, 1, 1)
    rhs = ab.MultiFab(ba, dm, 1, 0)
    lo, a = dump["sol0_lev0"]
    sol.upload(a, lo, ng=1)
    lo, a = dump["rhs_lev0"]
    rhs.upload(a, lo)
    D, N, P = ab.LinOpBCType.Dirichlet, ab.LinOpBCType.Neumann, ab.LinOpBCType.Periodic
    keep = []
    if prob_type in (2, 3, 6):
        op = ab.MLABecLaplacian([geom], [ba], [dm], agg_grid_size=agg_grid_size, con_grid_size=agg_grid_size,
                                max_coarsening_level=max_coarsening_level)
        op.setMaxOrder(maxorder)
        if prob_type == 3:     # the fields of problem 2 with inhomogeneous Neumann data (ghost cells of sol0) on every face
            IN = ab.LinOpBCType.inhomogNeumann
            op.setDomainBC((IN, IN, IN), (IN, IN, IN))
        elif prob_type == 6:   # the fields of problem 2 with Robin data on the x and z faces (reference driver prob_type 6)
            R = ab.LinOpBCType.Robin
            op.setDomainBC((R, D, R), (R, N, R))
        else:
            op.setDomainBC((D, N, N), (N, D, N))
        if prob_type == 6:
            robin = []
            for nm in ("robin_a", "robin_b", "robin_f"):
                f = ab.MultiFab(ba, dm, 1, 1)
                lo, a = dump[nm + "_lev0"]
                f.upload(a, lo, ng=1)
                robin.append(f)
            keep += robin
            op.setLevelBC(0, sol, robin=robin)
        else:
            op.setLevelBC(0, sol)
        op.setScalars(1.e-3, 1.0)
        acoef = ab.MultiFab(ba, dm, 1, 0)
        lo, a = dump["acoef_lev0"]
        acoef.upload(a, lo)
        op.setACoeffs(0, acoef)
        faces = []
        for d, nm in enumerate(("bx", "by", "bz")):
            nodal = [0, 0, 0]
            nodal[d] = 1
            f = ab.MultiFab(ba, dm, 1, 0, nodal=nodal)
            lo, a = dump[nm + "_lev0"]
            f.upload(a, lo)
            faces.append(f)
        op.setBCoeffs(0, faces)
        keep += [acoef] + faces
    elif prob_type == 7:      # MLALaplacian: alpha*a(x) - beta*Laplacian, homogeneous Dirichlet (reference driver prob_type 7)
        op = ab.MLALaplacian([geom], [ba], [dm], agg_grid_size=agg_grid_size, con_grid_size=agg_grid_size,
                             max_coarsening_level=max_coarsening_level)
        op.setMaxOrder(maxorder)
        op.setDomainBC((D, D, D), (D, D, D))
        op.setLevelBC(0, sol)
        op.setScalars(1.0, 1.0)
        acoef = ab.MultiFab(ba, dm, 1, 0)
        lo, a = dump["acoef_lev0"]
        acoef.upload(a, lo)
        op.setACoeffs(0, acoef)
        keep += [acoef]
    else:
        op = ab.MLPoisson([geom], [ba], [dm], agg_grid_size=agg_grid_size, con_grid_size=agg_grid_size,
                          max_coarsening_level=max_coarsening_level)
        op.setMaxOrder(maxorder)
        t = P if prob_type == 5 else D
        op.setDomainBC((t, t, t), (t, t, t))
        op.setLevelBC(0, sol)
    if fusion is not None:
        op.setSmootherFusion(fusion)
        if fusion:
            op.setFusedMinBoxCells(32 ** 3)     # tests exercise the fused pass on small boxes too (default: 64^3 and up)
    return dict(geom=geom, ba=ba, dm=dm, sol=sol, rhs=rhs, op=op, keep=keep, n=n_cell)


def build_problem_amr(ab, prob_type, n_cell, max_grid_size, dump, max_level=1, maxorder=3, agg_grid_size=-1, fusion=None):
    """Multi-level version of build_problem: level l has the domain refined by 2^l, grids = the central half of the
    coarser level's grids refined by 2 and chopped at max_grid_size (as oracle/ref_driver.cpp build_problem)."""
    assert prob_type in (1, 2)
    ab.Geometry.setup((0., 0., 0.), (1., 1., 1.), (0, 0, 0))
    geoms, bas, dms, sols, rhss, keep = [], [], [], [], [], []
    dlo, dhi = [0, 0, 0], [n_cell - 1] * 3      # domain of the level
    glo, ghi = [0, 0, 0], [n_cell - 1] * 3      # region covered by the level's grids
    for l in range(max_level + 1):
        geoms.append(ab.Geometry(tuple(dlo), tuple(dhi)))
        bas.append(ab.BoxArray(tuple(glo), tuple(ghi)).maxSize(max_grid_size))
        dms.append(ab.DistributionMapping(bas[-1]))
        sol = ab.MultiFab(bas[-1], dms[-1], 1, 1)
        rhs = ab.MultiFab(bas[-1], dms[-1], 1, 0)
        lo, a = dump[f"sol0_lev{l}"]
        sol.upload(a, lo, ng=1)
        lo, a = dump[f"rhs_lev{l}"]
        rhs.upload(a, lo)
        sols.append(sol)
        rhss.append(rhs)
        glo = [2 * (v + n_cell // 4) for v in glo]
        ghi = [2 * (v - n_cell // 4) + 1 for v in ghi]
        dlo = [2 * v for v in dlo]
        dhi = [2 * v + 1 for v in dhi]
    D, N = ab.LinOpBCType.Dirichlet, ab.LinOpBCType.Neumann
    if prob_type == 2:
        op = ab.MLABecLaplacian(geoms, bas, dms, agg_grid_size=agg_grid_size, con_grid_size=agg_grid_size)
        op.setMaxOrder(maxorder)
        op.setDomainBC((D, N, N), (N, D, N))
        for l in range(max_level + 1):
            op.setLevelBC(l, sols[l])
        op.setScalars(1.e-3, 1.0)
        for l in range(max_level + 1):
            acoef = ab.MultiFab(bas[l], dms[l], 1, 0)
            lo, a = dump[f"acoef_lev{l}"]
            acoef.upload(a, lo)
            op.setACoeffs(l, acoef)
            faces = []
            for d, nm in enumerate(("bx", "by", "bz")):
                nodal = [0, 0, 0]
                nodal[d] = 1
                f = ab.MultiFab(bas[l], dms[l], 1, 0, nodal=nodal)
                lo, a = dump[f"{nm}_lev{l}"]
                f.upload(a, lo)
                faces.append(f)
            op.setBCoeffs(l, faces)
            keep += [acoef] + faces
    else:
        op = ab.MLPoisson(geoms, bas, dms, agg_grid_size=agg_grid_size, con_grid_size=agg_grid_size)
        op.setMaxOrder(maxorder)
        op.setDomainBC((D, D, D), (D, D, D))
        for l in range(max_level + 1):
            op.setLevelBC(l, sols[l])
    if fusion is not None:
        op.setSmootherFusion(fusion)
        if fusion:
            op.setFusedMinBoxCells(32 ** 3)     # tests exercise the fused pass on small boxes too (default: 64^3 and up)
    return dict(geom=geoms, ba=bas, dm=dms, sol=sols, rhs=rhss, op=op, keep=keep, n=n_cell)


def rel_maxdiff(a, b):
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))


from amrex_b200.synth import (abeclap_beta_and_bc, abeclap_fields, synth_abeclap, synth_poisson)  # noqa: E402,F401
