"""Bit-exact parity of the host-side integer metadata with the reference (a9, a26, a30).

Targets come from tests/golden/meta_*.json (written by tests/golden/make_golden.py from the UNMODIFIED reference) and,
when oracle/_ref/ref_driver is present, from a live run of it.  No GPU is needed: these entry points are pure host code.
"""
import glob
import json
import os

import pytest

import amrex_b200 as ab
from common import GOLDEN, have_ref, run_ref


def _box6(b):
    return tuple(b[:6])


def _setup(args):
    per = 1 if args["prob_type"] == 5 else 0
    ab.Geometry.setup((0., 0., 0.), (1., 1., 1.), (per, per, per))
    n, mgs = args["n_cell"], args["max_grid_size"]
    nlev = args.get("max_level", 0) + 1
    geoms, bas = [], []
    lo, hi = [0, 0, 0], [n - 1] * 3
    glo, ghi = list(lo), list(hi)
    for l in range(nlev):
        geoms.append(ab.Geometry(glo, ghi))
        bas.append(ab.BoxArray(lo, hi).maxSize(mgs))
        # next level: central half of this level, refined by 2 (Tests/LinearSolvers/ABecLaplacian_C/MyTest.cpp:612-619)
        g = n // 4
        lo = [2 * (x + g) for x in lo]
        hi = [2 * (x - g + 1) - 1 for x in hi]
        glo = [2 * x for x in glo]
        ghi = [2 * (x + 1) - 1 for x in ghi]
    return geoms, bas, per


def check_meta(meta):
    args = meta["_args"]
    geoms, bas, per = _setup(args)
    nlev = len(bas)
    dms = [ab.DistributionMapping(ba, nprocs=1) for ba in bas]
    # --- MG hierarchy (MLLinOpT::defineGrids)
    H = ab.hierarchy(geoms, bas, dms, nprocs=1, agg_grid_size=args["agg_grid_size"], con_grid_size=args["agg_grid_size"])
    ref = meta["hierarchy"]
    assert len(H) == len(ref)
    for a in range(nlev):
        assert len(H[a]) == len(ref[a]), f"amr level {a}: {len(H[a])} MG levels, reference {len(ref[a])}"
        for m, (mine, r) in enumerate(zip(H[a], ref[a])):
            assert mine["domain"] == _box6(r["domain"]), (a, m)
            assert mine["boxes"] == [_box6(b) for b in r["boxes"]], (a, m)
            assert mine["dmap"] == r["dmap"], (a, m)
        # amrex::isMFIterSafe between consecutive MG levels (direct vs temporary + ParallelCopy paths; the F-cycle's
        # trilinear interpolation reads different ghost cells on the two paths).  Older golden files do not carry it.
        if "mfiter_safe" in meta:
            assert [int(lv["safe_with_next"]) for lv in H[a][:-1]] == meta["mfiter_safe"][a], a
    # --- FillBoundary local tags, cross and full stencil, in the reference's order
    for l in range(nlev):
        dom = meta["hierarchy"][l][0]["domain"]
        period = [(dom[3 + d] - dom[d] + 1) * per for d in range(3)]
        for key, cross in ((f"fb_cross_ng1_lev{l}", True), (f"fb_full_ng1_lev{l}", False)):
            mine = ab.fb_tags(bas[l], dms[l], 1, cross, period, 0, 0)
            r = meta[key]
            assert len(mine) == len(r), key
            for t, u in zip(mine, r):
                assert t["dbox"] == _box6(u["dbox"]) and t["sbox"] == _box6(u["sbox"]), key
                assert t["dst"] == u["dst"] and t["src"] == u["src"], key
    # --- SFC processor maps
    for key, pm in meta["sfc"].items():
        np_, lev = int(key.split("_")[0][2:]), int(key.split("lev")[1])
        assert ab.make_sfc(bas[lev], np_) == pm, key


GOLDEN_META = sorted(glob.glob(os.path.join(GOLDEN, "meta_*.json")))


@pytest.mark.parametrize("path", GOLDEN_META, ids=[os.path.basename(p)[5:-5] for p in GOLDEN_META])
def test_metadata_vs_golden(path):
    check_meta(json.load(open(path)))


@pytest.mark.skipif(not have_ref(), reason="oracle/_ref/ref_driver not built")
@pytest.mark.parametrize("kw", [dict(prob_type=1, n_cell=48, max_grid_size=16), dict(prob_type=2, n_cell=160, max_grid_size=32),
                                dict(prob_type=5, n_cell=96, max_grid_size=24), dict(prob_type=1, n_cell=64, max_grid_size=16, max_level=1)],
                         ids=["p1_48_16", "p2_160_32", "p5_96_24", "p1_64_16_lev1"])
def test_metadata_vs_live_reference(kw):
    meta, _ = run_ref(mode="meta", agg_grid_size=32, **kw)
    meta["_args"] = dict(kw, agg_grid_size=32)
    check_meta(meta)


@pytest.mark.skipif(not have_ref(), reason="oracle/_ref/ref_driver not built")
def test_mfiter_safe_flags_cover_both_outcomes():
    """The live-reference cases above compare the flags; this one makes sure both values occur (levels coarsened from the
    user's grids share their box list, agglomerated levels never do)."""
    meta, _ = run_ref(mode="meta", agg_grid_size=32, prob_type=2, n_cell=256, max_grid_size=64)
    flags = meta["mfiter_safe"][0]
    assert 1 in flags and 0 in flags and flags == sorted(flags, reverse=True), flags
    meta["_args"] = dict(prob_type=2, n_cell=256, max_grid_size=64, agg_grid_size=32)
    check_meta(meta)


def test_multirank_tags_are_symmetric():
    """For nprocs > 1 the reference cannot run here (no MPI); the pinned property is the one its own debug build checks
    (CheckRcvStats, AMReX_FabArrayCommI.H:195-200): rank r's send list to q equals rank q's receive list from r."""
    ab.Geometry.setup((0., 0., 0.), (1., 1., 1.), (1, 1, 0))
    ba = ab.BoxArray((0, 0, 0), (63, 63, 63)).maxSize(16)
    for nprocs in (2, 3, 8):
        dm = ab.DistributionMapping(ba, nprocs=nprocs)
        for cross in (True, False):
            snd = {r: ab.fb_tags(ba, dm, 1, cross, (64, 64, 0), r, 1) for r in range(nprocs)}
            rcv = {r: ab.fb_tags(ba, dm, 1, cross, (64, 64, 0), r, 2) for r in range(nprocs)}
            for r in range(nprocs):
                for q in range(nprocs):
                    s = [(t["dbox"], t["sbox"], t["dst"], t["src"]) for t in snd[r] if t["peer"] == q]
                    v = [(t["dbox"], t["sbox"], t["dst"], t["src"]) for t in rcv[q] if t["peer"] == r]
                    assert s == v, (nprocs, cross, r, q)
            # full stencil: every ghost cell the serial pattern fills is filled by local + received tags (the cross
            # stencil trims only the remote tags to faces, AMReX_FabArrayBase.cpp:835-854, so volumes differ there)
            if cross:
                continue
            serial = ab.fb_tags(ba, ab.DistributionMapping(ba, nprocs=1), 1, cross, (64, 64, 0), 0, 0)
            vol = lambda b: (b[3] - b[0] + 1) * (b[4] - b[1] + 1) * (b[5] - b[2] + 1)
            tot_serial = sum(vol(t["dbox"]) for t in serial)
            tot = sum(vol(t["dbox"]) for r in range(nprocs) for t in ab.fb_tags(ba, dm, 1, cross, (64, 64, 0), r, 0)) \
                + sum(vol(t["dbox"]) for r in range(nprocs) for t in rcv[r])
            assert tot == tot_serial, (nprocs, cross)


# ---- send/recv symmetry of ParallelCopy tags for FACE-centred arrays across an MG agglomeration transition: the
# one-node-thick face plane on the rank boundary is owned by boxes of both ranks (regression: 2-GPU 512^3 run)
@pytest.mark.parametrize("nprocs", [2, 4, 8])
@pytest.mark.parametrize("d", [0, 1, 2])
def test_cpc_face_symmetry_across_agglomeration(d, nprocs):
    import amrex_b200 as ab
    from amrex_b200 import capi
    ab.Geometry.setup((0., 0., 0.), (1., 1., 1.), (0, 0, 0))
    nodal = [0, 0, 0]
    nodal[d] = 1
    src = ab.BoxArray((0, 0, 0), (127, 127, 127)).maxSize(32)
    dms = ab.DistributionMapping(src, nprocs=nprocs)
    src.convert(nodal)
    src.coarsen(2)
    dst = ab.BoxArray((0, 0, 0), (63, 63, 63)).maxSize(32)
    dmd = ab.DistributionMapping(dst, nprocs=nprocs)
    dst.convert(nodal)

    def vol(b):
        return (b[3] - b[0] + 1) * (b[4] - b[1] + 1) * (b[5] - b[2] + 1)

    snd = {}
    rcv = {}
    covered = 0
    for me in range(nprocs):
        for t in capi.cpc_tags(dst, dmd, 0, src, dms, 0, (0, 0, 0), me, 1):
            snd[(me, t["peer"])] = snd.get((me, t["peer"]), 0) + vol(t["dbox"])
        for t in capi.cpc_tags(dst, dmd, 0, src, dms, 0, (0, 0, 0), me, 2):
            rcv[(t["peer"], me)] = rcv.get((t["peer"], me), 0) + vol(t["dbox"])
        covered += sum(vol(t["dbox"]) for t in capi.cpc_tags(dst, dmd, 0, src, dms, 0, (0, 0, 0), me, 0))
    assert snd == rcv          # what rank a packs for rank b is exactly what b expects from a
    if d == 2:   # the SFC split of 2/4/8 ranks always cuts across z: the shared face plane must be exchanged
        assert any(v > 0 for v in snd.values())


# ---- face links: the form in which the fused smoother's shell sweep replaces the copies between the boxes of one GPU
def _links_case(ab, ba, nprocs, period):
    dm = ab.DistributionMapping(ba, nprocs=nprocs)
    boxes = ba.boxes()
    pmap = dm.pmap(ba.size())
    out = {}
    for me in range(nprocs):
        out[me] = (ab.fb_face_links(ba, dm, period, me), [g for g in range(len(boxes)) if pmap[g] == me])
    return boxes, out


@pytest.mark.parametrize("n,mgs,nprocs,period", [(128, 64, 1, (0, 0, 0)), (128, 64, 1, (128, 128, 128)), (128, 64, 2, (0, 0, 0)),
                                                 (128, 32, 4, (128, 0, 128)), (64, 64, 1, (64, 64, 64)), ((128, 64, 64), 64, 1, (128, 64, 64)),
                                                 (96, 40, 3, (0, 96, 0))])
def test_face_links_geometry_and_symmetry(n, mgs, nprocs, period):
    """Every link (box b, face f) -> (box c, shift s) of amrex_b200_fb_face_links: the slab of ghost cells behind face f of b,
    shifted by s, lies inside c's valid box; c is on the same rank; c's opposite face links back to b with shift -s; a face
    without link has no same-rank box behind it (it is a domain face, or another rank's).  Uniform tilings always have the
    face-link form."""
    import amrex_b200 as ab
    ab.load_library()
    hi = tuple(x - 1 for x in n) if isinstance(n, tuple) else (n - 1,) * 3
    ba = ab.BoxArray((0, 0, 0), hi).maxSize(mgs)
    boxes, per_rank = _links_case(ab, ba, nprocs, period)
    dom = (0, 0, 0) + hi
    nlinks = 0
    for me, (links, mine) in per_rank.items():
        assert links is not None and len(links) == len(mine)
        for lb, g in enumerate(mine):
            b = boxes[g]
            for f in range(6):
                d, side = f % 3, f // 3
                slab = list(b)
                slab[d] = slab[d + 3] = (b[d] - 1) if side == 0 else (b[d + 3] + 1)
                c_local, s = links[lb][f]
                # the box (any rank) that holds the cells behind the face, through the periodic image if there is one
                img = list(slab)
                for q in range(3):
                    if period[q] and img[q] < dom[q]:
                        img[q] += period[q]; img[q + 3] += period[q]
                    elif period[q] and img[q + 3] > dom[q + 3]:
                        img[q] -= period[q]; img[q + 3] -= period[q]
                owners = [h for h, c in enumerate(boxes) if all(c[q] <= img[q] and img[q + 3] <= c[q + 3] for q in range(3))]
                inside = all(dom[q] <= img[q] and img[q + 3] <= dom[q + 3] for q in range(3))
                if c_local < 0:
                    assert s == (0, 0, 0)
                    assert (not inside) or (owners and owners[0] not in mine), (me, g, f)
                    continue
                nlinks += 1
                c = boxes[mine[c_local]]
                assert all(c[q] <= slab[q] + s[q] and slab[q + 3] + s[q] <= c[q + 3] for q in range(3)), (me, g, f, s)
                assert owners == [mine[c_local]], (me, g, f, owners)
                back = links[c_local][(d + 3) if side == 0 else d]
                assert back == (lb, tuple(-x for x in s)), (me, g, f, back)
    assert nlinks > 0 or (len(boxes) == 1 and not any(period))


def test_face_links_refuse_partial_faces():
    """A face fed by two smaller boxes, or covered in part (box lists that are not a tensor-product tiling - AMR patches):
    no face-link form; the exchange keeps its copy kernel."""
    import amrex_b200 as ab
    ab.load_library()
    # one 64x64x64 box next to two 64x32x64 boxes
    ba = ab.BoxArray(boxes=[(0, 0, 0, 63, 63, 63), (64, 0, 0, 127, 31, 63), (64, 32, 0, 127, 63, 63)])
    dm = ab.DistributionMapping(ba, nprocs=1)
    assert ab.fb_face_links(ba, dm, (0, 0, 0), 0) is None
    # a face covered in part: the neighbour is shorter than the face
    ba = ab.BoxArray(boxes=[(0, 0, 0, 63, 63, 63), (64, 0, 0, 127, 31, 63)])
    dm = ab.DistributionMapping(ba, nprocs=1)
    assert ab.fb_face_links(ba, dm, (0, 0, 0), 0) is None
    # the two halves on different ranks: the big box's face is another rank's business on rank 0 ... but rank 1 holds both
    # small boxes, whose faces towards the big box are remote and whose common face is a whole-face link
    ba = ab.BoxArray(boxes=[(0, 0, 0, 63, 63, 63), (64, 0, 0, 127, 31, 63), (64, 32, 0, 127, 63, 63)])
    dm = ab.DistributionMapping(ba, nprocs=2)
    pm = dm.pmap(3)
    for me in (0, 1):
        mine = [g for g in range(3) if pm[g] == me]
        links = ab.fb_face_links(ba, dm, (0, 0, 0), me)
        if set(mine) == {1, 2}:
            assert links is not None and links[0][4] == (1, (0, 0, 0)) and links[1][1] == (0, (0, 0, 0))
