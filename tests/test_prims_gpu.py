"""Kernel-level parity (through the C ABI) of the operator primitives with the reference: apply (FillBoundary + BC +
stencil), red-black smooth, correction / solution residual, restriction, prolongation-add, coefficient average-down.

Targets: committed golden dumps of the reference at 16^3 (tests/golden/prim_*.npz) and live runs of oracle/_ref/ref_driver
at larger sizes.  fp64 tolerance 1e-13 relative to the field's max norm (the kernels keep the reference's association
order, so most outputs are bit-identical)."""
import json
import os

import numpy as np
import pytest

from common import GOLDEN, build_problem, have_ref, rel_maxdiff, run_ref

pytestmark = pytest.mark.gpu
TOL = 1e-13


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, f"prim_{name}.npz"))
    args = json.loads(bytes(z["_args"]).decode())
    dump = {k: (tuple(int(v) for v in z[k + "__lo"]), z[k]) for k in z.files if not k.endswith("__lo") and k != "_args"}
    return args, dump


def run_prims(ab, args, dump, fusion):
    n, mgs, m = args["n_cell"], args["max_grid_size"], args["prim_mglev"]
    P = build_problem(ab, args["prob_type"], n, mgs, dump, maxorder=args["linop_maxorder"],
                      agg_grid_size=args.get("agg_grid_size", -1), fusion=fusion)
    op = P["op"]
    op.prepareForSolve()
    nm = n >> m
    up = lambda mf, key, ng: mf.upload(dump[key][1], dump[key][0], ng=ng)
    val = lambda mf, nn=nm: mf.download((0, 0, 0), (nn, nn, nn))
    ref = lambda key: dump[key][1]
    out = {}
    if args["prob_type"] == 2:   # averaged-down coefficients of the level
        a = op.coeff(0, m, 0)
        out["acoef"] = rel_maxdiff(val(a), ref("prim_acoef_mg"))
        for d, k in enumerate(("prim_bx_mg", "prim_by_mg", "prim_bz_mg")):
            shp = [nm, nm, nm]
            shp[d] += 1
            out[k] = rel_maxdiff(op.coeff(0, m, 1 + d).download((0, 0, 0), tuple(shp)), ref(k))
    x = op.make(0, m, 1)
    b = op.make(0, m, 0)
    y = op.make(0, m, 0)
    up(x, "prim_x", 1)
    up(b, "prim_b", 0)
    op.apply(0, m, y, x)
    out["apply"] = rel_maxdiff(val(y), ref("prim_apply_homog"))
    xg = x.download((-1, -1, -1), (nm + 2,) * 3, ng=1)
    rg = ref("prim_x_after_bc_homog")
    # Domain faces only, and on them only the cells a cross stencil reads: the ghost cells straight across a box face.  The rim
    # of every box face is left out - in the domain-wide download those cells may come from the EDGE ghost cells of the
    # neighbouring box, which the cross-stencil exchange does not fill (AMReX_FabArrayBase.cpp:835-870 clips tags to face slabs).
    boxes, _, _ = op.level(0, m)
    for d in range(3):
        t1, t2 = [a for a in range(3) if a != d]
        for s in (0, -1):
            keep = np.zeros((nm, nm), dtype=bool)
            for bx in boxes:
                on_face = (bx[d] == 0) if s == 0 else (bx[d + 3] == nm - 1)
                if on_face and bx[t1 + 3] - bx[t1] >= 2 and bx[t2 + 3] - bx[t2] >= 2:
                    keep[bx[t1] + 1:bx[t1 + 3], bx[t2] + 1:bx[t2 + 3]] = True
            sl = [slice(1, -1)] * 3
            sl[d] = s
            mine_f, ref_f = xg[tuple(sl)], rg[tuple(sl)]
            out[f"bc_homog_{d}{s}"] = rel_maxdiff(np.where(keep, mine_f, 0.0), np.where(keep, ref_f, 0.0)) if keep.any() else 0.0
    up(x, "prim_x", 1)
    op.smooth(0, m, x, b)
    out["smooth1"] = rel_maxdiff(val(x), ref("prim_smooth1"))
    op.smooth(0, m, x, b)
    out["smooth2"] = rel_maxdiff(val(x), ref("prim_smooth2"))
    op.residual(0, m, y, x, b)
    out["corres"] = rel_maxdiff(val(y), ref("prim_corres"))
    if "prim_restrict" in dump:
        c = op.make(0, m + 1, 0)
        op.restriction(0, m + 1, c, y)
        out["restrict"] = rel_maxdiff(val(c, nm // 2), ref("prim_restrict"))
        f = op.make(0, m, 0)
        f.copy_from(x)
        op.interp_add(0, m, f, c)
        out["interp_add"] = rel_maxdiff(val(f), ref("prim_interp_add"))
    if m == 0:
        up(x, "prim_x", 1)
        op.residual(0, 0, y, x, b, inhomog=True)
        out["solres"] = rel_maxdiff(val(y), ref("prim_solres"))
    return out


@pytest.mark.parametrize("fusion", [0, 1])
@pytest.mark.parametrize("name", ["p2_n16_g8_m0", "p1_n16_g8_m0", "p2_n16_g8_m1"])
def test_prims_vs_golden(ab, name, fusion):
    args, dump = load_golden(name)
    out = run_prims(ab, args, dump, fusion)
    bad = {k: v for k, v in out.items() if not v <= TOL}
    assert not bad, bad


@pytest.mark.skipif(not have_ref(), reason="oracle/_ref/ref_driver not built")
@pytest.mark.parametrize("fusion", [0, 1])
@pytest.mark.parametrize("kw", [
    dict(prob_type=2, n_cell=128, max_grid_size=64, linop_maxorder=2, prim_mglev=0, agg_grid_size=32),
    dict(prob_type=2, n_cell=128, max_grid_size=32, linop_maxorder=3, prim_mglev=1, agg_grid_size=32),
    dict(prob_type=1, n_cell=96, max_grid_size=32, linop_maxorder=3, prim_mglev=0, agg_grid_size=32),
    dict(prob_type=5, n_cell=64, max_grid_size=32, linop_maxorder=2, prim_mglev=0, agg_grid_size=32),
    dict(prob_type=2, n_cell=64, max_grid_size=32, linop_maxorder=2, prim_mglev=1, agg_grid_size=32),   # re-gridded coarse level
], ids=["abec128", "abec128_m1_mo3", "pois96_mo3", "periodic64", "abec64_m1_agg"])
def test_prims_vs_live_reference(ab, kw, fusion):
    _, dump = run_ref(dump=True, mode="prim", **kw)
    out = run_prims(ab, kw, dump, fusion)
    bad = {k: v for k, v in out.items() if not v <= TOL}
    assert not bad, bad


# ---- fused smoother, generation 4 (bulk-async-copy staged pass): every compiled launch plan must reproduce the
#      reference schedule (one kernel per colour) BIT FOR BIT, on both MG levels that are eligible (64^3 and 32^3 boxes)
# (tile_y, EARLY, LATE, cell pairs per thread on rows of > 64 cells [0: by launch size, 1, 2], box size): 64^3 / 32^3 boxes
# always run one pair per thread (generation 4), a 128^3 box the kernel the switch names (2: generation 5)
FUSED4_PLANS = [(8, 4, 2, 0, 64), (8, 4, 3, 0, 64), (6, 5, 3, 0, 64), (6, 4, 2, 0, 64), (4, 4, 4, 0, 64),
                (8, 4, 2, 2, 128), (8, 4, 2, 1, 128), (8, 4, 2, 0, 128), (8, 4, 3, 2, 128), (6, 5, 3, 2, 128), (6, 4, 2, 2, 128), (4, 4, 4, 2, 128)]


def _two_smooths(ab, op, n, mglev, seed):
    nn = n >> mglev
    x = op.make(0, mglev, 1)
    b = op.make(0, mglev, 0)
    rng = np.random.default_rng(seed)
    b.upload(rng.standard_normal((nn, nn, nn)), (0, 0, 0))
    x.setVal(0.0, ng=1)
    x.upload(rng.standard_normal((nn, nn, nn)), (0, 0, 0))
    ab.profile_enable(True)
    op.smooth(0, mglev, x, b)
    op.smooth(0, mglev, x, b)
    names = set(q[0] for q in ab.profile_report())
    ab.profile_enable(False)
    return x.download((0, 0, 0), (nn, nn, nn)), names


@pytest.mark.parametrize("plan", FUSED4_PLANS)
def test_fused4_abeclap_bitwise(ab, plan):
    from common import synth_abeclap
    n, mgs = 128, plan[4]
    want = {}
    P = synth_abeclap(ab, n, mgs, fusion=0)
    P["op"].prepareForSolve()
    for mglev in (0, 1):
        want[mglev], names = _two_smooths(ab, P["op"], n, mglev, 11 + mglev)
        assert "b200mg_gsrb4" not in names
    P = synth_abeclap(ab, n, mgs, fusion=1)
    op = P["op"]
    assert ab.lib.amrex_b200_set_fused4_plan(*plan[:3]) == 0
    ab.lib.b200mg_set_gsrb4_sync(plan[3])
    try:
        op.prepareForSolve()
        for mglev in (0, 1):
            got, names = _two_smooths(ab, op, n, mglev, 11 + mglev)
            assert "b200mg_gsrb4" in names, names          # the staged pass really ran (no silent fallback)
            assert np.array_equal(got, want[mglev]), f"plan {plan} mglev {mglev}: max|diff| {np.abs(got - want[mglev]).max():.3e}"
    finally:
        ab.lib.amrex_b200_set_fused4_plan(8, 4, 2)
        ab.lib.b200mg_set_gsrb4_sync(0)


@pytest.mark.parametrize("kind,n,mgs", [("abeclap", 96, 96), ("abeclap", 72, 72), ("poisson", 96, 96), ("abeclap", 192, 96), ("abeclap", 104, 104)])
def test_fused5_partial_rows_bitwise(ab, kind, n, mgs):
    """Generation 5 of the fused pass (two cell pairs per thread, the second 64 cells to the right of the first) on rows
    shorter than 128 cells, where only some lanes own a second pair and the last cell of the row sits in a second pair of a
    lane other than 31: 96-, 72- and 104-cell rows (32, 4 and 20 second pairs), one box and eight boxes.  Bit for bit the
    colour sweeps, two smooths in a row."""
    from common import synth_abeclap, synth_poisson
    synth = synth_abeclap if kind == "abeclap" else synth_poisson
    P = synth(ab, n, mgs, fusion=0)
    P["op"].prepareForSolve()
    want, names = _two_smooths(ab, P["op"], n, 0, 7)
    assert "b200mg_gsrb4" not in names
    P = synth(ab, n, mgs, fusion=1)
    ab.lib.b200mg_set_gsrb4_sync(2)
    try:
        P["op"].prepareForSolve()
        got, names = _two_smooths(ab, P["op"], n, 0, 7)
        assert "b200mg_gsrb4" in names, names
        assert np.array_equal(got, want), f"max|diff| {np.abs(got - want).max():.3e}"
    finally:
        ab.lib.b200mg_set_gsrb4_sync(0)


@pytest.mark.parametrize("plan", [(8, 4, 2), (6, 5, 3), (4, 4, 4), (8, 4, 2, 128)])
@pytest.mark.parametrize("kind", ["abeclap", "poisson"])
def test_fused4_zero_input_bitwise(ab, kind, plan):
    """smooth(zero_input) - MLMG's cor.setVal(0) + first pre-smooth as ONE pass that never reads cor - must give the bits
    of setVal(0) followed by the ordinary smooth, also when cor holds garbage (NaN) on entry; a second smooth follows to
    show that the ghost cells the pass leaves behind are handled."""
    from common import synth_abeclap, synth_poisson
    n, mgs = 128, (plan[3] if len(plan) > 3 else 64)      # 128: two cell pairs per thread on the finer level (forced)
    synth = synth_abeclap if kind == "abeclap" else synth_poisson
    assert ab.lib.amrex_b200_set_fused4_plan(*plan[:3]) == 0
    ab.lib.b200mg_set_gsrb4_sync(2 if mgs == 128 else 0)
    try:
        out = {}
        for zero_input in (False, True):
            P = synth(ab, n, mgs, fusion=1)
            op = P["op"]
            op.prepareForSolve()
            for mglev in (0, 1):
                nn = n >> mglev
                x = op.make(0, mglev, 1)
                b = op.make(0, mglev, 0)
                b.upload(np.random.default_rng(5 + mglev).standard_normal((nn, nn, nn)), (0, 0, 0))
                x.setVal(float("nan") if zero_input else 0.0, ng=1)
                ab.profile_enable(True)
                if zero_input:
                    op.smooth(0, mglev, x, b, zero_input=True)
                else:
                    op.smooth(0, mglev, x, b, skip_fillboundary=True)
                op.smooth(0, mglev, x, b)
                names = set(q[0] for q in ab.profile_report())
                ab.profile_enable(False)
                assert "b200mg_gsrb4" in names, names
                if zero_input:
                    assert "b200mg_setval" not in names, names      # the zeroing really was skipped
                out[(zero_input, mglev)] = x.download((0, 0, 0), (nn, nn, nn))
        for mglev in (0, 1):
            assert np.isfinite(out[(True, mglev)]).all()
            assert np.array_equal(out[(False, mglev)], out[(True, mglev)]), f"mglev {mglev}"
    finally:
        ab.lib.amrex_b200_set_fused4_plan(8, 4, 2)
        ab.lib.b200mg_set_gsrb4_sync(0)


@pytest.mark.skipif(not have_ref(), reason="oracle/_ref/ref_driver not built")
@pytest.mark.parametrize("prob,n,mgs", [(1, 64, 32), (1, 96, 40), (2, 64, 32)])
def test_fused4_vs_pairs_reference_problems(ab, prob, n, mgs):
    """Poisson (prob 1) and ABecLap (prob 2) on the reference's own coefficient / BC data, incl. a non-power-of-two domain."""
    _, dump = run_ref(dump=True, mode="solve", prob_type=prob, n_cell=n, max_grid_size=mgs, linop_maxorder=2, agg_grid_size=32)
    res = {}
    for fusion in (0, 1):
        P = build_problem(ab, prob, n, mgs, dump, maxorder=2, fusion=fusion)
        op = P["op"]
        op.prepareForSolve()
        res[fusion] = _two_smooths(ab, op, n, 0, 3)
    assert "b200mg_gsrb4" in res[1][1]
    assert np.array_equal(res[0][0], res[1][0])


def test_base_fi_entries(ab):
    """The Base amrex_fi_* entries the reference's Fortran modules bind beyond the solver path (MFIter family, reductions,
    element-wise products, lincomb, SumBoundary, iMultiFab, geometry / distromap getters), against numpy on small data."""
    import ctypes as C
    lib = ab.lib
    n, mgs = 32, 16
    ab.Geometry.setup((0., 0., 0.), (1., 2., 4.), (1, 0, 1))
    geom = ab.Geometry((0, 0, 0), (n - 1,) * 3)
    ba = ab.BoxArray((0, 0, 0), (n - 1,) * 3).maxSize(mgs)
    dm = ab.DistributionMapping(ba)
    rng = np.random.default_rng(7)
    A, B = rng.standard_normal((n, n, n)), rng.standard_normal((n, n, n)) + 3.0
    x, y, z = ab.MultiFab(ba, dm, 1, 1), ab.MultiFab(ba, dm, 1, 1), ab.MultiFab(ba, dm, 1, 1)
    x.setVal(0.0, ng=1); y.setVal(0.0, ng=1); z.setVal(0.0, ng=1)
    x.upload(A, (0, 0, 0)); y.upload(B, (0, 0, 0))
    for f, rt in (("amrex_fi_multifab_min", C.c_double), ("amrex_fi_multifab_max", C.c_double), ("amrex_fi_multifab_norm1", C.c_double),
                  ("amrex_fi_multifab_norm2", C.c_double), ("amrex_fi_distromap_issame", C.c_int), ("amrex_fi_mfiter_grid_index", C.c_int),
                  ("amrex_fi_boxarray_intersects_box", C.c_int)):
        getattr(lib, f).restype = rt
    P, I = C.c_void_p, C.c_int
    i3 = lambda v: (C.c_int * 3)(*v)
    assert lib.amrex_fi_multifab_min(P(x.ptr.value), 0, 0) == pytest.approx(A.min(), rel=0, abs=0)
    assert lib.amrex_fi_multifab_max(P(x.ptr.value), 0, 0) == pytest.approx(A.max(), rel=0, abs=0)
    assert lib.amrex_fi_multifab_norm1(P(x.ptr.value), 0) == pytest.approx(np.abs(A).sum(), rel=1e-13)
    assert lib.amrex_fi_multifab_norm2(P(x.ptr.value), 0) == pytest.approx(np.sqrt((A * A).sum()), rel=1e-13)
    ng0 = i3((0, 0, 0))
    lib.amrex_fi_multifab_copy(P(z.ptr.value), P(x.ptr.value), 0, 0, 1, ng0)
    lib.amrex_fi_multifab_multiply(P(z.ptr.value), P(y.ptr.value), 0, 0, 1, ng0)
    assert np.array_equal(z.download((0, 0, 0), (n, n, n)), A * B)
    lib.amrex_fi_multifab_divide(P(z.ptr.value), P(y.ptr.value), 0, 0, 1, ng0)
    assert np.array_equal(z.download((0, 0, 0), (n, n, n)), (A * B) / B)
    lib.amrex_fi_multifab_lincomb(P(z.ptr.value), C.c_double(2.0), P(x.ptr.value), 0, C.c_double(-0.5), P(y.ptr.value), 0, 0, 1, ng0)
    assert np.array_equal(z.download((0, 0, 0), (n, n, n)), 2.0 * A + (-0.5) * B)
    ab.check()
    # MFIter over the local boxes: grid indices, valid / grown boxes, data pointer bounds
    it = P()
    lib.amrex_fi_new_mfiter_r(C.byref(it), P(x.ptr.value), 0, 0)
    valid, seen = C.c_int(0), []
    lib.amrex_fi_mfiter_is_valid(it, C.byref(valid))
    while valid.value:
        lo, hi, nod = i3((0, 0, 0)), i3((0, 0, 0)), i3((9, 9, 9))
        lib.amrex_fi_mfiter_validbox(it, lo, hi, nod)
        g = lib.amrex_fi_mfiter_grid_index(it)
        assert tuple(lo) + tuple(hi) == ba.boxes()[g] and tuple(nod) == (0, 0, 0)
        glo, ghi = i3((0, 0, 0)), i3((0, 0, 0))
        lib.amrex_fi_mfiter_growntilebox(it, glo, ghi, 1, nod)
        assert tuple(glo) == tuple(v - 1 for v in lo) and tuple(ghi) == tuple(v + 1 for v in hi)
        dp, dlo, dhi = P(), i3((0, 0, 0)), i3((0, 0, 0))
        lib.amrex_fi_multifab_dataptr_iter(P(x.ptr.value), it, C.byref(dp), dlo, dhi)
        assert dp.value and tuple(dlo) == tuple(glo) and tuple(dhi) == tuple(ghi)
        seen.append(g)
        lib.amrex_fi_increment_mfiter(it, C.byref(valid))
    lib.amrex_fi_delete_mfiter(it)
    assert sorted(seen) == list(range(ba.size()))
    # SumBoundary: ghost cells are added to the valid cells they overlap (periodic in x and z)
    w = ab.MultiFab(ba, dm, 1, 1)
    w.setVal(1.0, ng=1)
    lib.amrex_fi_multifab_sum_boundary(P(w.ptr.value), P(geom.ptr.value), 0, 1)
    ab.check()
    got = w.download((0, 0, 0), (n, n, n))
    cnt = np.ones((n, n, n))
    for d, per in enumerate((1, 0, 1)):       # every cell collects one extra contribution per box face it touches (domain faces only when periodic)
        for c in range(n):
            k = (1 if c % mgs == 0 and (c > 0 or per) else 0) + (1 if c % mgs == mgs - 1 and (c < n - 1 or per) else 0)
            sl = [slice(None)] * 3
            sl[d] = c
            cnt[tuple(sl)] *= (1 + k)
    # contributions multiply along the three directions (a corner ghost cell of a box overlaps the diagonal neighbour)
    assert np.array_equal(got, cnt)
    # iMultiFab, geometry / distromap getters
    imf, pba, pdm = P(), P(ba.ptr.value), P(dm.ptr.value)
    lib.amrex_fi_new_imultifab(C.byref(imf), C.byref(pba), C.byref(pdm), 1, i3((1, 1, 1)), i3((0, 0, 0)))
    lib.amrex_fi_imultifab_setval(imf, 7, 0, 1, i3((1, 1, 1)))
    lib.amrex_fi_delete_imultifab(imf)
    ab.check()
    pm = i3((0, 0, 0))
    lib.amrex_fi_geometry_get_pmask(P(geom.ptr.value), pm)
    assert tuple(pm) == (1, 0, 1)
    plo, phi = (C.c_double * 3)(), (C.c_double * 3)()
    lib.amrex_fi_geometry_get_probdomain(P(geom.ptr.value), plo, phi)
    assert tuple(plo) == (0., 0., 0.) and tuple(phi) == (1., 2., 4.)
    dm2 = P()
    lib.amrex_fi_clone_distromap(C.byref(dm2), P(dm.ptr.value))
    assert lib.amrex_fi_distromap_issame(dm2, P(dm.ptr.value)) == 1
    lib.amrex_fi_delete_distromap(dm2)
    assert lib.amrex_fi_boxarray_intersects_box(P(ba.ptr.value), i3((5, 5, 5)), i3((6, 6, 6))) == 1
    assert lib.amrex_fi_boxarray_intersects_box(P(ba.ptr.value), i3((n, n, n)), i3((n + 2, n + 2, n + 2))) == 0
