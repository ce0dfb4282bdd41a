"""Synthetic inputs of the reference's own test problems (Tests/LinearSolvers/ABecLaplacian_C/initProb_K.H:7-38 Poisson,
:74-140 ABecLap), generated box by box with numpy so that no run-time code reads the reference tree.
Used by bench.py, __graft_entry__.smoke() and the tests; nothing here touches the checker."""
import numpy as np


def _centres(lo, hi, n):
    h = 1.0 / n
    return [(np.arange(lo[d], hi[d] + 1) + 0.5) * h for d in range(3)]


def abeclap_fields(box, n, a=1.e-3, b=1.0):
    """rhs, exact on the cells of `box` (lo0,lo1,lo2,hi0,hi1,hi2) of an n^3 unit-cube domain (prob_type 2)."""
    x, y, z = np.meshgrid(*_centres(box[:3], box[3:], n), indexing="ij")
    w, sigma = 0.05, 10.0
    theta = 0.5 * np.log(3.0) / (w + 1.e-50)
    pi, tpi, fpi = np.pi, 2 * np.pi, 4 * np.pi
    fac = 12.0 * pi * pi
    xc = yc = zc = 0.5
    r = np.sqrt((x - xc) ** 2 + (y - yc) ** 2 + (z - zc) ** 2)
    beta = (sigma - 1.) / 2. * np.tanh(theta * (r - 0.25)) + (sigma + 1.) / 2.
    tmp = np.cosh(theta * (r - 0.25))
    dbdrfac = (sigma - 1.) / 2. / (tmp * tmp) * theta / r * b
    cx, cy, cz, sx, sy, sz = np.cos(tpi * x), np.cos(tpi * y), np.cos(tpi * z), np.sin(tpi * x), np.sin(tpi * y), np.sin(tpi * z)
    cx4, cy4, cz4, sx4, sy4, sz4 = np.cos(fpi * x), np.cos(fpi * y), np.cos(fpi * z), np.sin(fpi * x), np.sin(fpi * y), np.sin(fpi * z)
    exact = cx * cy * cz + .25 * cx4 * cy4 * cz4
    rhs = beta * b * fac * (cx * cy * cz + cx4 * cy4 * cz4) \
        + dbdrfac * ((x - xc) * (tpi * sx * cy * cz + pi * sx4 * cy4 * cz4)
                     + (y - yc) * (tpi * cx * sy * cz + pi * cx4 * sy4 * cz4)
                     + (z - zc) * (tpi * cx * cy * sz + pi * cx4 * cy4 * sz4)) \
        + a * (cx * cy * cz + 0.25 * cx4 * cy4 * cz4)
    return rhs, exact


def abeclap_beta_and_bc(box, n):
    """Cell-centred beta and the Dirichlet data (exact solution at clamped coordinates) on `box` grown by one cell."""
    lo = [v - 1 for v in box[:3]]
    hi = [v + 1 for v in box[3:]]
    x, y, z = np.meshgrid(*_centres(lo, hi, n), indexing="ij")
    w, sigma = 0.05, 10.0
    theta = 0.5 * np.log(3.0) / (w + 1.e-50)
    r = np.sqrt((x - 0.5) ** 2 + (y - 0.5) ** 2 + (z - 0.5) ** 2)
    beta = (sigma - 1.) / 2. * np.tanh(theta * (r - 0.25)) + (sigma + 1.) / 2.
    xb, yb, zb = np.clip(x, 0., 1.), np.clip(y, 0., 1.), np.clip(z, 0., 1.)
    tpi, fpi = 2 * np.pi, 4 * np.pi
    bc = np.cos(tpi * xb) * np.cos(tpi * yb) * np.cos(tpi * zb) + .25 * np.cos(fpi * xb) * np.cos(fpi * yb) * np.cos(fpi * zb)
    # Dirichlet data lives in the ghost cells OUTSIDE the domain only.  Everything inside is the zero initial guess: a
    # box's ghost layer overlaps its neighbours' valid cells, and MultiFab.upload writes every local fab the region
    # meets, so non-zero values there would leak into the neighbours' initial guess (depending on which boxes are local).
    inside = (x > 0.) & (x < 1.) & (y > 0.) & (y < 1.) & (z > 0.) & (z < 1.)
    bc[inside] = 0.0
    return beta, bc, lo


def synth_abeclap(ab, n, mgs, maxorder=2, fusion=None, a=1.e-3, b=1.0, keep_host=False):
    """Variable-coefficient MLABecLaplacian problem (prob_type 2) on an n^3 domain chopped into mgs^3 boxes, distributed
    over the ranks of the library's communicator.  Returns dict with geom/ba/dm/sol/sol0/rhs/op and, with keep_host,
    per-local-box host arrays (rhs, exact)."""
    ab.Geometry.setup((0., 0., 0.), (1., 1., 1.), (0, 0, 0))
    geom = ab.Geometry((0, 0, 0), (n - 1,) * 3)
    ba = ab.BoxArray((0, 0, 0), (n - 1,) * 3).maxSize(mgs)
    dm = ab.DistributionMapping(ba)
    me = ab.lib.amrex_b200_myproc()
    pmap = dm.pmap(ba.size())
    sol = ab.MultiFab(ba, dm, 1, 1)
    sol0 = ab.MultiFab(ba, dm, 1, 1)
    rhs = ab.MultiFab(ba, dm, 1, 0)
    acoef = ab.MultiFab(ba, dm, 1, 0)
    bcc = ab.MultiFab(ba, dm, 1, 1)
    faces = []
    for d in range(3):
        nodal = [0, 0, 0]
        nodal[d] = 1
        faces.append(ab.MultiFab(ba, dm, 1, 0, nodal=nodal))
    host = {}
    for g, box in enumerate(ba.boxes()):
        if pmap[g] != me:
            continue
        r, ex = abeclap_fields(box, n, a, b)
        beta, bc, glo = abeclap_beta_and_bc(box, n)
        rhs.upload(r, box[:3])
        bcc.upload(beta, glo, ng=1)
        sol0.upload(bc, glo, ng=1)
        if keep_host:
            host[g] = dict(box=box, rhs=r, exact=ex)
    acoef.setVal(1.0)
    ab.lib.amrex_b200_average_cellcenter_to_face(faces[0].ptr, faces[1].ptr, faces[2].ptr, bcc.ptr, geom.ptr)
    ab.check()
    sol.copy_from(sol0, ng=1)
    D, N = ab.LinOpBCType.Dirichlet, ab.LinOpBCType.Neumann
    op = ab.MLABecLaplacian([geom], [ba], [dm])
    op.setMaxOrder(maxorder)
    op.setDomainBC((D, N, N), (N, D, N))
    op.setLevelBC(0, sol0)
    op.setScalars(a, b)
    op.setACoeffs(0, acoef)
    op.setBCoeffs(0, faces)
    if fusion is not None:
        op.setSmootherFusion(fusion)
        if fusion:
            op.setFusedMinBoxCells(32 ** 3)     # tests exercise the fused pass on small boxes too (default: 64^3 and up)
    return dict(geom=geom, ba=ba, dm=dm, sol=sol, sol0=sol0, rhs=rhs, op=op, keep=[acoef, bcc] + faces, n=n, host=host,
                pmap=pmap, me=me)


def synth_poisson(ab, n, mgs, maxorder=2, fusion=None):
    """MLPoisson on an n^3 unit cube chopped into mgs^3 boxes, homogeneous Dirichlet on every face (prob_type 1 BCs);
    fields are supplied by the caller.  Returns dict with geom/ba/dm/sol0/op."""
    ab.Geometry.setup((0., 0., 0.), (1., 1., 1.), (0, 0, 0))
    geom = ab.Geometry((0, 0, 0), (n - 1,) * 3)
    ba = ab.BoxArray((0, 0, 0), (n - 1,) * 3).maxSize(mgs)
    dm = ab.DistributionMapping(ba)
    sol0 = ab.MultiFab(ba, dm, 1, 1)
    sol0.setVal(0.0, ng=1)
    D = ab.LinOpBCType.Dirichlet
    op = ab.MLPoisson([geom], [ba], [dm])
    op.setMaxOrder(maxorder)
    op.setDomainBC((D, D, D), (D, D, D))
    op.setLevelBC(0, sol0)
    if fusion is not None:
        op.setSmootherFusion(fusion)
        if fusion:
            op.setFusedMinBoxCells(32 ** 3)
    return dict(geom=geom, ba=ba, dm=dm, sol0=sol0, op=op, n=n)


def synth_poisson_periodic(ab, ncell, mgs, maxorder=2, fusion=None, keep_host=False):
    """Fully periodic MLPoisson (BASELINE config 5) on an ncell = (n0, n1, n2) domain of cubic cells of size 1 / min(ncell)
    (so the physical box grows with the cell count: weak scaling keeps the per-cell problem), chopped into mgs^3 boxes.
    rhs = the Laplacian of the analytic solution of the reference's Poisson test (initProb_K.H:7-38), which has period 1 and
    zero mean.  The operator is singular: MLMG makes the right-hand side solvable and the solution is defined up to a
    constant.  Returns dict with geom/ba/dm/sol/sol0/rhs/op and, with keep_host, per-local-box host arrays (rhs, exact)."""
    nb = min(ncell)
    ab.Geometry.setup((0., 0., 0.), tuple(float(n) / nb for n in ncell), (1, 1, 1))
    geom = ab.Geometry((0, 0, 0), tuple(n - 1 for n in ncell))
    ba = ab.BoxArray((0, 0, 0), tuple(n - 1 for n in ncell)).maxSize(mgs)
    dm = ab.DistributionMapping(ba)
    me = ab.lib.amrex_b200_myproc()
    pmap = dm.pmap(ba.size())
    sol = ab.MultiFab(ba, dm, 1, 1)
    sol0 = ab.MultiFab(ba, dm, 1, 1)
    rhs = ab.MultiFab(ba, dm, 1, 0)
    sol0.setVal(0.0, ng=1)
    sol.setVal(0.0, ng=1)
    host = {}
    h = 1.0 / nb
    tpi, fpi = 2 * np.pi, 4 * np.pi
    fac = tpi * tpi * 3.0
    for g, box in enumerate(ba.boxes()):
        if pmap[g] != me:
            continue
        x, y, z = np.meshgrid(*[(np.arange(box[d], box[d + 3] + 1) + 0.5) * h for d in range(3)], indexing="ij")
        s2, s4 = np.sin(tpi * x) * np.sin(tpi * y) * np.sin(tpi * z), np.sin(fpi * x) * np.sin(fpi * y) * np.sin(fpi * z)
        r = -fac * s2 - fac * s4
        rhs.upload(r, box[:3])
        if keep_host:
            host[g] = dict(box=box, rhs=r, exact=s2 + .25 * s4)
    P = ab.LinOpBCType.Periodic
    op = ab.MLPoisson([geom], [ba], [dm])
    op.setMaxOrder(maxorder)
    op.setDomainBC((P, P, P), (P, P, P))
    op.setLevelBC(0, sol0)
    if fusion is not None:
        op.setSmootherFusion(fusion)
    return dict(geom=geom, ba=ba, dm=dm, sol=sol, sol0=sol0, rhs=rhs, op=op, keep=[], n=ncell, host=host, pmap=pmap, me=me,
                ncells=int(np.prod(ncell)))


def synth_abeclap_amr(ab, n, mgs, max_level=1, maxorder=3, fusion=None, a=1.e-3, b=1.0, keep_host=False):
    """Two-level (max_level = 1) composite problem of BASELINE config 4: the variable-coefficient MLABecLaplacian of
    synth_abeclap on an n^3 base level plus a refined patch (ratio 2) over the central half of the domain, i.e. (n/2)^3 coarse
    cells refined to n^3 fine cells - the grids of the reference's test (Tests/LinearSolvers/ABecLaplacian_C/MyTest.cpp:612-619).
    Returns dict with per-level lists geom/ba/dm/sol/sol0/rhs and op, host[level][g] = dict(box, rhs, exact)."""
    ab.Geometry.setup((0., 0., 0.), (1., 1., 1.), (0, 0, 0))
    geoms, bas, dms, sols, sol0s, rhss, keep, hosts = [], [], [], [], [], [], [], []
    me = ab.lib.amrex_b200_myproc()
    dlo, dhi = [0, 0, 0], [n - 1] * 3
    glo, ghi = [0, 0, 0], [n - 1] * 3
    acoefs, facess = [], []
    for l in range(max_level + 1):
        nl = n << l
        geom = ab.Geometry(tuple(dlo), tuple(dhi))
        ba = ab.BoxArray(tuple(glo), tuple(ghi)).maxSize(mgs)
        dm = ab.DistributionMapping(ba)
        pmap = dm.pmap(ba.size())
        sol = ab.MultiFab(ba, dm, 1, 1)
        sol0 = ab.MultiFab(ba, dm, 1, 1)
        rhs = ab.MultiFab(ba, dm, 1, 0)
        acoef = ab.MultiFab(ba, dm, 1, 0)
        bcc = ab.MultiFab(ba, dm, 1, 1)
        faces = []
        for d in range(3):
            nodal = [0, 0, 0]
            nodal[d] = 1
            faces.append(ab.MultiFab(ba, dm, 1, 0, nodal=nodal))
        host = {}
        for g, box in enumerate(ba.boxes()):
            if pmap[g] != me:
                continue
            r, ex = abeclap_fields(box, nl, a, b)
            beta, bc, glo_ = abeclap_beta_and_bc(box, nl)
            rhs.upload(r, box[:3])
            bcc.upload(beta, glo_, ng=1)
            sol0.upload(bc, glo_, ng=1)
            if keep_host:
                host[g] = dict(box=box, rhs=r, exact=ex)
        acoef.setVal(1.0)
        ab.lib.amrex_b200_average_cellcenter_to_face(faces[0].ptr, faces[1].ptr, faces[2].ptr, bcc.ptr, geom.ptr)
        ab.check()
        sol.copy_from(sol0, ng=1)
        geoms.append(geom); bas.append(ba); dms.append(dm); sols.append(sol); sol0s.append(sol0); rhss.append(rhs)
        acoefs.append(acoef); facess.append(faces); keep += [acoef, bcc] + faces; hosts.append(host)
        glo = [2 * (v + n // 4) for v in glo]
        ghi = [2 * (v - n // 4) + 1 for v in ghi]
        dlo = [2 * v for v in dlo]
        dhi = [2 * v + 1 for v in dhi]
    D, N = ab.LinOpBCType.Dirichlet, ab.LinOpBCType.Neumann
    op = ab.MLABecLaplacian(geoms, bas, dms)
    op.setMaxOrder(maxorder)
    op.setDomainBC((D, N, N), (N, D, N))
    for l in range(max_level + 1):
        op.setLevelBC(l, sol0s[l])
    op.setScalars(a, b)
    for l in range(max_level + 1):
        op.setACoeffs(l, acoefs[l])
        op.setBCoeffs(l, facess[l])
    if fusion is not None:
        op.setSmootherFusion(fusion)
    return dict(geom=geoms, ba=bas, dm=dms, sol=sols, sol0=sol0s, rhs=rhss, op=op, keep=keep, n=n, host=hosts, me=me,
                ncells=int(sum(b_.numPts() for b_ in bas)))
