// MG hierarchy, boundary bookkeeping and level operations of the cell-centred operators.
#include "AMReX_MLMG.H"

#include <cmath>
#include <cstring>
#include <cuda_runtime.h>

namespace amrex {


// ================================================================================ MGHierarchy::define
// Restates MLLinOpT::defineGrids (AMReX_MLLinOp.H:795-1165) without semicoarsening / hidden dimensions / EB.
void MGHierarchy::define (Vector<Geometry> const& a_geom, Vector<BoxArray> const& a_grids,
                          Vector<DistributionMapping> const& a_dmap, LPInfo info, int nprocs)
{
    if (info.agg_grid_size <= 0) { info.agg_grid_size = LPInfo::getDefaultAgglomerationGridSize(); }
    if (info.con_grid_size <= 0) { info.con_grid_size = LPInfo::getDefaultConsolidationGridSize(); }
    AMREX_ALWAYS_ASSERT_WITH_MESSAGE(!info.do_semicoarsening && info.hidden_direction < 0,
                                     "semicoarsening / hidden dimension are outside this library's scope");
    constexpr int mg_coarsen_ratio = 2, mg_box_min_width = 2, mg_domain_min_width = 2;

    num_amr_levels = 0;
    for (std::size_t a = 0; a < a_geom.size(); ++a) { if (!a_grids[a].empty()) { ++num_amr_levels; } }
    amr_ref_ratio.assign(num_amr_levels, 0);   // entry amrlev: ratio between amrlev and amrlev+1
    num_mg_levels.assign(num_amr_levels, 0);
    geom.assign(num_amr_levels, {}); grids.assign(num_amr_levels, {}); dmap.assign(num_amr_levels, {});
    mg_coarsen_ratio_vec.clear();

    const RealBox rb = a_geom[0].ProbDomain();
    const int coord = a_geom[0].Coord();
    const Array<int, 3> is_per = a_geom[0].isPeriodic();
    const IntVect rr2(mg_coarsen_ratio);

    for (int amrlev = num_amr_levels - 1; amrlev > 0; --amrlev) {
        num_mg_levels[amrlev] = 1;
        geom[amrlev].push_back(a_geom[amrlev]);
        grids[amrlev].push_back(a_grids[amrlev]);
        dmap[amrlev].push_back(a_dmap[amrlev]);
        IntVect rr = rr2;
        const Box dom = a_geom[amrlev].Domain();
        for (int i = 0; i < 2; ++i) {
            if (!dom.coarsenable(rr)) { Abort("MLLinOp: Uncoarsenable domain"); }
            const Box cdom = amrex::coarsen(dom, rr);
            if (cdom == a_geom[amrlev - 1].Domain()) { break; }
            ++num_mg_levels[amrlev];
            geom[amrlev].emplace_back(cdom, rb, coord, is_per);
            grids[amrlev].push_back(amrex::coarsen(a_grids[amrlev], rr));
            dmap[amrlev].push_back(a_dmap[amrlev]);
            rr *= rr2;
        }
        amr_ref_ratio[amrlev - 1] = rr[0];
    }

    num_mg_levels[0] = 1;
    geom[0].push_back(a_geom[0]);
    grids[0].push_back(a_grids[0]);
    dmap[0].push_back(a_dmap[0]);

    domain_covered.assign(num_amr_levels, 0);
    const Long npts0 = grids[0][0].numPts();
    domain_covered[0] = (npts0 == geom[0][0].Domain().numPts());
    for (int amrlev = 1; amrlev < num_amr_levels; ++amrlev) {
        if (!domain_covered[amrlev - 1]) { break; }
        domain_covered[amrlev] = (grids[amrlev][0].numPts() == geom[amrlev][0].Domain().numPts());
    }

    Box aggbox;
    bool aggable = false;
    if (grids[0][0].size() > 1 && info.do_agglomeration) {
        if (domain_covered[0]) { aggbox = geom[0][0].Domain(); aggable = true; }
        else { aggbox = grids[0][0].minimalBox(); aggable = (aggbox.numPts() == npts0); }
    }

    agged = false; coned = false; agg_lev = 0; con_lev = 0;

    if (info.do_agglomeration && aggable) {
        Box dbx = geom[0][0].Domain();
        Box bbx = aggbox;
        const Real nbxs = Real(grids[0][0].size());
        const Real threshold_npts = Real(info.agg_grid_size) * Real(info.agg_grid_size) * Real(info.agg_grid_size);
        Vector<Box> domainboxes{dbx}, boundboxes{bbx};
        Vector<int> agg_flag{0};
        Vector<IntVect> accum{IntVect(1)};
        for (int lev = 0; lev < info.max_coarsening_level; ++lev) {
            bool ok = true;
            for (int d = 0; d < 3; ++d) {
                IntVect rr_dir(1); rr_dir[d] = mg_coarsen_ratio;
                ok = ok && dbx.coarsenable(rr_dir, IntVect(mg_domain_min_width)) && bbx.coarsenable(rr_dir, IntVect(mg_box_min_width));
            }
            if (!ok) { break; }
            accum.push_back(accum.back() * rr2);
            domainboxes.push_back(dbx.coarsen(rr2));
            boundboxes.push_back(bbx.coarsen(rr2));
            const bool to_agg = (bbx.d_numPts() / nbxs) < 0.999 * threshold_npts;
            agg_flag.push_back(to_agg);
        }
        for (int lev = 1, nlevs = int(domainboxes.size()); lev < nlevs; ++lev) {
            if (!agged && !agg_flag[lev] && a_grids[0].coarsenable(accum[lev], IntVect(mg_box_min_width))) {
                grids[0].push_back(amrex::coarsen(a_grids[0], accum[lev]));
                dmap[0].push_back(a_dmap[0]);
            } else {
                IntVect cr = domainboxes[lev - 1].length() / domainboxes[lev].length();
                if (!grids[0].back().coarsenable(cr)) { break; }
                BoxArray nba(boundboxes[lev]);
                nba.maxSize(IntVect(info.agg_grid_size));
                grids[0].push_back(nba);
                dmap[0].push_back(DistributionMapping());
                if (!agged) { agged = true; agg_lev = lev; }
            }
            geom[0].emplace_back(domainboxes[lev], rb, coord, is_per);
        }
    } else {
        Long consolidation_threshold = 0;
        Real avg_npts = 0.0;
        if (info.do_consolidation) {
            avg_npts = Real(a_grids[0].d_numPts()) / Real(nprocs);
            consolidation_threshold = Long(info.con_grid_size) * info.con_grid_size * info.con_grid_size;
        }
        const Box dom0 = a_geom[0].Domain();
        IntVect rr_vec(1);
        for (int lev = 0; lev < info.max_coarsening_level; ++lev) {
            bool ok = true;
            for (int d = 0; d < 3; ++d) {
                IntVect rr_dir(1); rr_dir[d] = rr_vec[d] * mg_coarsen_ratio;
                ok = ok && dom0.coarsenable(rr_dir, IntVect(mg_domain_min_width)) && a_grids[0].coarsenable(rr_dir, IntVect(mg_box_min_width));
            }
            if (!ok) { break; }
            rr_vec *= rr2;
            geom[0].emplace_back(amrex::coarsen(dom0, rr_vec), rb, coord, is_per);
            grids[0].push_back(amrex::coarsen(a_grids[0], rr_vec));
            if (info.do_consolidation) {
                if (avg_npts / Real(Long(rr_vec[0]) * rr_vec[1] * rr_vec[2]) < Real(0.999) * Real(consolidation_threshold)) {
                    coned = true; con_lev = int(dmap[0].size());
                    dmap[0].push_back(DistributionMapping());
                } else {
                    dmap[0].push_back(dmap[0].back());
                }
            } else {
                dmap[0].push_back(a_dmap[0]);
            }
        }
    }

    num_mg_levels[0] = int(grids[0].size());
    for (int mglev = 0; mglev < num_mg_levels[0] - 1; ++mglev) {
        mg_coarsen_ratio_vec.push_back(geom[0][mglev].Domain().length() / geom[0][mglev + 1].Domain().length());
    }
    for (int amrlev = 0; amrlev + 1 < num_amr_levels; ++amrlev) {
        if (amr_ref_ratio[amrlev] == 4 && mg_coarsen_ratio_vec.empty()) { mg_coarsen_ratio_vec.push_back(IntVect(2)); }
    }

    if (agged) {   // makeAgglomeratedDMap, AMReX_MLLinOp.H:1294-1319
        for (std::size_t i = 1; i < grids[0].size(); ++i) {
            if (dmap[0][i].empty()) {
                auto sfc = DistributionMapping::makeSFC(grids[0][i], true, nprocs);
                Vector<int> pmap(grids[0][i].size());
                for (int ip = 0; ip < nprocs; ++ip) { for (int ib : sfc[ip]) { pmap[ib] = ip; } }
                dmap[0][i].define(std::move(pmap));
            }
        }
    } else if (coned) {   // makeConsolidatedDMap, AMReX_MLLinOp.H:1323-1375
        int factor = 1;
        const int ratio = info.con_ratio, strategy = info.con_strategy;
        for (std::size_t i = 1; i < grids[0].size(); ++i) {
            if (!dmap[0][i].empty()) { continue; }
            factor *= ratio;
            Vector<int> pmap = dmap[0][i - 1].ProcessorMap();
            if (strategy == 1) { for (auto& x : pmap) { x /= ratio; } }
            else if (strategy == 2) {
                const int nprocs_con = int(std::ceil(Real(nprocs) / Real(factor)));
                for (auto& x : pmap) { x = x % nprocs_con; }
            } else if (strategy == 3) {
                if (factor == ratio) {
                    auto sfc = DistributionMapping::makeSFC(grids[0][i], true, nprocs);
                    for (int ip = 0; ip < nprocs; ++ip) { for (int ib : sfc[ip]) { pmap[ib] = ip; } }
                }
                for (auto& x : pmap) { x /= ratio; }
            }
            dmap[0][i].define(std::move(pmap));
        }
    }

    for (int amrlev = 1; amrlev < num_amr_levels; ++amrlev) {
        AMREX_ALWAYS_ASSERT_WITH_MESSAGE(grids[amrlev][0].coarsenable(amr_ref_ratio[amrlev - 1]),
                                         "MLLinOp: grids not coarsenable between AMR levels");
    }
}

// ========================================================================================= BndrySlabs
template <class T>
void BndrySlabs<T>::define (LevelLayout const& L, bool inside, int rad, int extent)
{
    clear();
    const int nl = L.numLocal();
    m_boxes.resize(nl * 6); m_offs.resize(nl * 6); m_hdesc.resize(nl * 6);
    std::size_t total = 0;
    for (int li = 0; li < nl; ++li) {
        for (int f = 0; f < 6; ++f) {
            const Orientation face(f);
            Box b = inside ? insideCell(L.box(li), face, rad) : adjCell(L.box(li), face, rad);
            for (int d = 0; d < 3; ++d) { if (d != face.coordDir()) { b.grow(d, extent); } }
            m_boxes[li * 6 + f] = b;
            m_offs[li * 6 + f] = total;
            total += std::size_t(b.numPts());
            total = (total + 1) / 2 * 2;
        }
    }
    m_nelem = total; m_defined = true;
    if (total) {
        m_data = static_cast<T*>(The_Arena()->alloc(total * sizeof(T)));
        Gpu::memset_async(m_data, 0, total * sizeof(T));
    }
    for (int n = 0; n < nl * 6; ++n) {
        Desc& d = m_hdesc[n]; Box const& b = m_boxes[n];
        d.p = m_data + m_offs[n];
        for (int a = 0; a < 3; ++a) { d.lo[a] = b.smallEnd(a); d.hi[a] = b.bigEnd(a); }
        d.jstride = b.length(0); d.kstride = Long(b.length(0)) * b.length(1); d.nstride = d.kstride * b.length(2);
    }
    m_ddesc.assign(m_hdesc);
}

template <class T>
void BndrySlabs<T>::clear ()
{
    if (m_data) { The_Arena()->free(m_data); m_data = nullptr; }
    m_boxes.clear(); m_offs.clear(); m_hdesc.clear(); m_ddesc.clear(); m_nelem = 0; m_defined = false;
}

template <class T>
void BndrySlabs<T>::upload (std::vector<T> const& h)
{
    AMREX_ALWAYS_ASSERT(h.size() == m_nelem);
    if (m_nelem) { Gpu::htod_memcpy_async(m_data, h.data(), m_nelem * sizeof(T)); Gpu::streamSynchronize(); }
}

template <class T>
void BndrySlabs<T>::setVal (T v)
{
    std::vector<T> h(m_nelem, v); upload(h);
}

template class BndrySlabs<double>;
template class BndrySlabs<int>;

// ============================================================================================= MLLinOp
void MLLinOp::define (Vector<Geometry> const& a_geom, Vector<BoxArray> const& a_grids,
                      Vector<DistributionMapping> const& a_dmap, LPInfo const& a_info)
{
    AMREX_ALWAYS_ASSERT_WITH_MESSAGE(Gpu::Initialized(), "amrex::Initialize (GPU) must be called before defining a linear operator");
    info = a_info;
    H.define(a_geom, a_grids, a_dmap, a_info, ParallelDescriptor::NProcs());
    m_needs_coarse_data_for_bc = !H.domain_covered[0];
    defineAuxData();
}

void MLLinOp::defineAuxData ()
{
    m_lev.resize(H.num_amr_levels);
    for (int a = 0; a < H.num_amr_levels; ++a) {
        m_lev[a].resize(H.num_mg_levels[a]);
        for (int m = 0; m < H.num_mg_levels[a]; ++m) {
            m_lev[a][m] = std::make_unique<LevelData>();
            LevelData& L = *m_lev[a][m];
            L.layout = LevelLayout::get(H.grids[a][m], H.dmap[a][m]);
            L.undrrelxr.define(*L.layout, true);
            buildMasks(a, m);
        }
    }
    m_bndry_sol.resize(H.num_amr_levels);
    m_bndry_cor.resize(H.num_amr_levels);
    for (int a = 0; a < H.num_amr_levels; ++a) {
        m_bndry_sol[a] = std::make_unique<BndrySlabs<double>>();
        m_bndry_sol[a]->define(*lev(a, 0).layout, false);
        if (a > 0) {
            m_bndry_cor[a] = std::make_unique<BndrySlabs<double>>();
            m_bndry_cor[a]->define(*lev(a, 0).layout, false);
        }
    }
    m_norm_fine_mask.resize(std::max(0, H.num_amr_levels - 1));
    for (int a = 0; a + 1 < H.num_amr_levels; ++a) {
        m_norm_fine_mask[a] = std::make_unique<iMultiFab>(
            makeFineMask(H.grids[a][0], H.dmap[a][0], H.grids[a + 1][0], H.amr_ref_ratio[a], 1, 0));
    }
    defineAmrData();
}

// MultiMask::define (AMReX_MultiMask.cpp:24-70) on the slabs of M: 2 outside the (periodically grown) domain, 1 inside and
// not covered by the level's grids, 0 covered.  ngrow = max(out_rad, extent_rad) of the slabs.
namespace {
void fill_masks (BndrySlabs<int> const& M, LevelLayout const& layout, Geometry const& geom, BoxArray const& ba, int ngrow, std::vector<int>& h)
{
    Box domain = geom.Domain();
    for (int d = 0; d < 3; ++d) { if (geom.isPeriodic(d)) { domain.grow(d, ngrow); } }
    const auto pshifts = geom.periodicity().shiftIntVect();
    h.assign(M.numElements(), 0);
    std::vector<std::pair<int, Box>> isects;
    for (int li = 0; li < layout.numLocal(); ++li) {
        for (int f = 0; f < 6; ++f) {
            Box const& sb = M.box(li, f);
            int* p = h.data() + M.offset(li, f);
            const Long nx = sb.length(0), ny = sb.length(1);
            auto at = [&] (int i, int j, int k) -> int& {
                return p[(i - sb.smallEnd(0)) + nx * ((j - sb.smallEnd(1)) + ny * (k - sb.smallEnd(2)))];
            };
            for (int k = sb.smallEnd(2); k <= sb.bigEnd(2); ++k) for (int j = sb.smallEnd(1); j <= sb.bigEnd(1); ++j)
                for (int i = sb.smallEnd(0); i <= sb.bigEnd(0); ++i) { at(i, j, k) = domain.contains(i, j, k) ? 1 : 2; }
            for (auto const& pit : pshifts) {
                ba.intersections(sb + pit, isects);
                for (auto const& is : isects) {
                    const Box b = is.second - pit;
                    for (int k = b.smallEnd(2); k <= b.bigEnd(2); ++k) for (int j = b.smallEnd(1); j <= b.bigEnd(1); ++j)
                        for (int i = b.smallEnd(0); i <= b.bigEnd(0); ++i) { at(i, j, k) = 0; }
                }
            }
        }
    }
}
}

// m_maskvals: MultiMask with in_rad=0, out_rad=1, extent_rad=0 (cross stencil, AMReX_MLCellLinOp.H:393-408)
void MLLinOp::buildMasks (int a, int m)
{
    LevelData& L = lev(a, m);
    L.mask.define(*L.layout, false);
    std::vector<int> h;
    fill_masks(L.mask, *L.layout, H.geom[a][m], H.grids[a][m], 1, h);
    L.mask.upload(h);
    // which faces have any uncovered ghost cell (only those need boundary-condition work)
    L.bcfaces_h.clear();
    for (int li = 0; li < L.layout->numLocal(); ++li) {
        for (int f = 0; f < 6; ++f) {
            Box const& sb = L.mask.box(li, f);
            const int* p = h.data() + L.mask.offset(li, f);
            bool any = false;
            for (Long n = 0, N = sb.numPts(); n < N && !any; ++n) { any = p[n] > 0; }
            if (any) {
                b200mg_bcface fc; fc.box = li; fc.face = f; fc.bctype = 0; fc.blen = L.layout->box(li).length(f % 3); fc.bcloc = 0.0;
                L.bcfaces_h.push_back(fc);
            }
        }
    }
}

void MLLinOp::setDomainBC (Array<BCType, 3> const& a_lobc, Array<BCType, 3> const& a_hibc)
{
    m_lobc = a_lobc; m_hibc = a_hibc; m_lobc_orig = m_lobc; m_hibc_orig = m_hibc;
    for (int d = 0; d < 3; ++d) {
        if (H.geom[0][0].isPeriodic(d)) {
            AMREX_ALWAYS_ASSERT(m_lobc[d] == BCType::Periodic && m_hibc[d] == BCType::Periodic);
        } else {
            AMREX_ALWAYS_ASSERT(m_lobc[d] != BCType::Periodic && m_hibc[d] != BCType::Periodic);
        }
        if (m_lobc[d] == BCType::Robin || m_hibc[d] == BCType::Robin) {
            AMREX_ALWAYS_ASSERT_WITH_MESSAGE(supportRobinBC(), "Robin BC not supported");
        }
        // inhomogeneous Neumann and Robin act as Neumann inside the cycle; their data moves into the right-hand side
        // (and, for Robin, the diagonal): MLLinOpT::setDomainBC, AMReX_MLLinOp.H:1211-1221
        if (m_lobc[d] == BCType::inhomogNeumann || m_lobc[d] == BCType::Robin) { m_lobc[d] = BCType::Neumann; }
        if (m_hibc[d] == BCType::inhomogNeumann || m_hibc[d] == BCType::Robin) { m_hibc[d] = BCType::Neumann; }
    }
}

bool MLLinOp::hasInhomogNeumannBC () const noexcept
{
    for (int d = 0; d < 3; ++d) { if (m_lobc_orig[d] == BCType::inhomogNeumann || m_hibc_orig[d] == BCType::inhomogNeumann) { return true; } }
    return false;
}

bool MLLinOp::hasRobinBC () const noexcept
{
    for (int d = 0; d < 3; ++d) { if (m_lobc_orig[d] == BCType::Robin || m_hibc_orig[d] == BCType::Robin) { return true; } }
    return false;
}

namespace {
void innu_faces (Array<LinOpBCType, 3> const& lo, Array<LinOpBCType, 3> const& hi, int on_face[6], LinOpBCType which = LinOpBCType::inhomogNeumann)
{
    for (int d = 0; d < 3; ++d) { on_face[d] = (lo[d] == which); on_face[d + 3] = (hi[d] == which); }
}
}

// MLCellABecLapT::applyInhomogNeumannTerm (AMReX_MLCellABecLap.H:295-513)
void MLLinOp::applyInhomogNeumannTerm (int amrlev, MultiFab& rhs) const
{
    if (hasRobinBC()) {       // rhs += beta*dxinv^2*b*A next to Robin faces (AMReX_MLCellABecLap.H:448-510)
        LevelData const& L = lev(amrlev, 0);
        const int nf = int(L.bcfaces_h.size());
        if (nf > 0) {
            AMREX_ALWAYS_ASSERT_WITH_MESSAGE(int(m_robin.size()) > amrlev && m_robin[amrlev][0], "Robin BC: setLevelBC must supply robinbc_a / _b / _f");
            Array<MultiFab const*, 3> b; Real bscalar;
            getFluxCoeffs(amrlev, b, bscalar);
            const Real* dxi = H.geom[amrlev][0].InvCellSize();
            const b200mg_fab* out3[3] = {rhs.d_fabs(), rhs.d_fabs(), rhs.d_fabs()};
            const b200mg_fab* b3[3] = {b[0] ? b[0]->d_fabs() : nullptr, b[1] ? b[1]->d_fabs() : nullptr, b[2] ? b[2]->d_fabs() : nullptr};
            const double fac[3] = {bscalar * dxi[0] * dxi[0], bscalar * dxi[1] * dxi[1], bscalar * dxi[2] * dxi[2]};
            const double dx3[3] = {dxi[0], dxi[1], dxi[2]};
            int on_face[6]; innu_faces(m_lobc_orig, m_hibc_orig, on_face, LinOpBCType::Robin);
            for (int f = 0; f < 6; ++f) {
                if (!on_face[f]) { continue; }
                int only[6] = {0, 0, 0, 0, 0, 0}; only[f] = 1;
                B200_KCALL(b200mg_robin(nf, L.bcfaces.data(), L.layout->d_vbox(), out3, b3, nullptr, L.mask.d_table(), m_robin[amrlev][0]->d_table(),
                                        m_robin[amrlev][1]->d_table(), m_robin[amrlev][2]->d_table(), fac, dx3, only, 1, Gpu::gpuStream()));
            }
        }
    }
    if (!hasInhomogNeumannBC()) { return; }
    LevelData const& L = lev(amrlev, 0);
    const int nf = int(L.bcfaces_h.size());
    if (nf == 0) { return; }
    AMREX_ALWAYS_ASSERT_WITH_MESSAGE(m_bndry_sol[amrlev] != nullptr, "inhomogeneous Neumann BC: setLevelBC must supply the boundary data");
    Array<MultiFab const*, 3> b; Real bscalar;
    getFluxCoeffs(amrlev, b, bscalar);
    const Real* dxi = H.geom[amrlev][0].InvCellSize();
    const b200mg_fab* out3[3] = {rhs.d_fabs(), rhs.d_fabs(), rhs.d_fabs()};
    const b200mg_fab* b3[3] = {b[0] ? b[0]->d_fabs() : nullptr, b[1] ? b[1]->d_fabs() : nullptr, b[2] ? b[2]->d_fabs() : nullptr};
    const double fac[3] = {bscalar * dxi[0], bscalar * dxi[1], bscalar * dxi[2]};
    int on_face[6]; innu_faces(m_lobc_orig, m_hibc_orig, on_face);
    // a cell on a domain edge or corner collects the terms of two or three faces: one launch per face orientation, in the
    // reference's order (x low, x high, y low, ...), keeps the updates race free and the roundings identical
    for (int d = 0; d < 3; ++d) {
        for (int side = 0; side < 2; ++side) {
            const int f = d + 3 * side;
            if (!on_face[f]) { continue; }
            int only[6] = {0, 0, 0, 0, 0, 0}; only[f] = 1;
            B200_KCALL(b200mg_apply_innu(nf, L.bcfaces.data(), L.layout->d_vbox(), out3, b3, L.mask.d_table(), m_bndry_sol[amrlev]->d_table(),
                                         fac, only, 0, Gpu::gpuStream()));
        }
    }
}

// MLCellABecLapT::addInhomogNeumannFlux (AMReX_MLCellABecLap.H:517-620): mult_bcoef: grad holds -b grad(phi), else grad(phi)
void MLLinOp::addInhomogNeumannFlux (int amrlev, Array<MultiFab*, 3> const& grad, MultiFab const& sol, bool mult_bcoef) const
{
    const bool has_innu = hasInhomogNeumannBC(), has_robin = hasRobinBC();
    if (!has_innu && !has_robin) { return; }
    LevelData const& L = lev(amrlev, 0);
    const int nf = int(L.bcfaces_h.size());
    if (nf == 0) { return; }
    Array<MultiFab const*, 3> b{{nullptr, nullptr, nullptr}}; Real bscalar = 1.0;
    if (mult_bcoef) { getFluxCoeffs(amrlev, b, bscalar); }
    const b200mg_fab* out3[3] = {grad[0]->d_fabs(), grad[1]->d_fabs(), grad[2]->d_fabs()};
    const b200mg_fab* b3[3] = {b[0] ? b[0]->d_fabs() : nullptr, b[1] ? b[1]->d_fabs() : nullptr, b[2] ? b[2]->d_fabs() : nullptr};
    const double f = mult_bcoef ? -1.0 : 1.0;
    const double fac[3] = {f, f, f};
    int on_face[6];
    if (has_innu) {
        innu_faces(m_lobc_orig, m_hibc_orig, on_face);
        B200_KCALL(b200mg_apply_innu(nf, L.bcfaces.data(), L.layout->d_vbox(), out3, b3, L.mask.d_table(), m_bndry_sol[amrlev]->d_table(),
                                     fac, on_face, 1, Gpu::gpuStream()));
    }
    if (has_robin) {
        AMREX_ALWAYS_ASSERT_WITH_MESSAGE(int(m_robin.size()) > amrlev && m_robin[amrlev][0], "Robin BC: setLevelBC must supply robinbc_a / _b / _f");
        const Real* dxi = H.geom[amrlev][0].InvCellSize();
        const double dx3[3] = {dxi[0], dxi[1], dxi[2]};
        innu_faces(m_lobc_orig, m_hibc_orig, on_face, LinOpBCType::Robin);
        B200_KCALL(b200mg_robin(nf, L.bcfaces.data(), L.layout->d_vbox(), out3, b3, sol.d_fabs(), L.mask.d_table(), m_robin[amrlev][0]->d_table(),
                                m_robin[amrlev][1]->d_table(), m_robin[amrlev][2]->d_table(), fac, dx3, on_face, 2, Gpu::gpuStream()));
    }
}

void MLLinOp::setCoarseFineBC (const MultiFab* crse, int crse_ratio, LinOpBCType bc_type)
{
    m_coarse_data_for_bc = crse; m_coarse_data_crse_ratio = crse_ratio; m_coarse_fine_bc_type = bc_type;
}

// MLMGBndry::setBoxBC (AMReX_MLMGBndry.H:113-161) for every local box of the level
void MLLinOp::buildBCFaces (int a, int m)
{
    LevelData& L = lev(a, m);
    Geometry const& geom = H.geom[a][m];
    const Box domain = geom.Domain();
    const Real* dx0 = H.geom[a][0].CellSize();
    const int ratio = (a == 0) ? (m_needs_coarse_data_for_bc ? m_coarse_data_crse_ratio : 1) : H.amr_ref_ratio[a - 1];
    const int cf_type = (a == 0) ? int(m_coarse_fine_bc_type) : int(LinOpBCType::Dirichlet);
    const int nl = L.layout->numLocal();
    L.bcond.resize(nl); L.bcloc.resize(nl);
    for (int li = 0; li < nl; ++li) {
        Box const& bx = L.layout->box(li);
        for (int f = 0; f < 6; ++f) {
            const int d = f % 3; const bool low = f < 3;
            const int dface = low ? domain.smallEnd(d) : domain.bigEnd(d);
            const int bface = low ? bx.smallEnd(d) : bx.bigEnd(d);
            if (dface == bface && !geom.isPeriodic(d)) {
                L.bcloc[li][f] = low ? m_domain_bloc_lo[d] : m_domain_bloc_hi[d];
                const BCType t = low ? m_lobc[d] : m_hibc[d];
                if (t == BCType::Dirichlet) { L.bcond[li][f] = 101; }
                else if (t == BCType::Neumann) { L.bcond[li][f] = 102; }
                else if (t == BCType::reflect_odd) { L.bcond[li][f] = 103; }
                else { Abort("MLMGBndry::setBoxBC: Unknown LinOpBCType"); }
            } else {
                L.bcond[li][f] = cf_type;
                L.bcloc[li][f] = (ratio > 0) ? Real(0.5) * Real(ratio) * dx0[d] : 0.0;
            }
        }
    }
    for (auto& fc : L.bcfaces_h) { fc.bctype = L.bcond[fc.box][fc.face]; fc.bcloc = L.bcloc[fc.box][fc.face]; }
    L.bcfaces.assign(L.bcfaces_h);
}

// MLCellLinOpT::setLevelBC (AMReX_MLCellLinOp.H:513-642): physical-boundary ghost values of levelbcdata are copied
// into the level's boundary slabs (InterpBndryData::setPhysBndryValues, AMReX_InterpBndryData.H:128-156).
void MLLinOp::setLevelBC (int amrlev, const MultiFab* levelbcdata, const MultiFab* robinbc_a, const MultiFab* robinbc_b,
                          const MultiFab* robinbc_f)
{
    AMREX_ALWAYS_ASSERT(amrlev >= 0 && amrlev < H.num_amr_levels);
    AMREX_ALWAYS_ASSERT_WITH_MESSAGE(m_lobc[0] != BCType::bogus, "setDomainBC must be called before setLevelBC");
    if (levelbcdata) {
        AMREX_ALWAYS_ASSERT(levelbcdata->nGrow() >= 1);
        AMREX_ALWAYS_ASSERT_WITH_MESSAGE(levelbcdata->boxArray() == H.grids[amrlev][0] && levelbcdata->DistributionMap() == H.dmap[amrlev][0],
                                         "MLLinOp::setLevelBC: levelbcdata must live on the level's grids and distribution");
    }
    LevelData& L = lev(amrlev, 0);
    BndrySlabs<double>& B = *m_bndry_sol[amrlev];
    B.setVal(0.0);
    if (amrlev == 0 && m_needs_coarse_data_for_bc) {
        // Level solve of a level that does not cover its domain (AMReX_MLCellLinOp.H:536-566): the faces that are not
        // physical boundaries take Dirichlet data interpolated from the coarse MultiFab of setCoarseFineBC (zero without one)
        const int ratio = (m_coarse_data_crse_ratio > 0) ? m_coarse_data_crse_ratio : 2;
        if (!m_amr_bndry[0] || m_amr_bndry[0]->ratio != ratio) { defineAmrBndry(0, ratio); }
        AmrBndry& A = *m_amr_bndry[0];
        if (m_coarse_data_for_bc != nullptr) {
            AMREX_ALWAYS_ASSERT(m_coarse_data_crse_ratio > 0);
            const Box cbx = amrex::coarsen(H.geom[0][0].Domain(), ratio);
            A.crse_sol_br.copyFrom(*m_coarse_data_for_bc, H.geom[0][0].periodicity(cbx));
        } else {
            for (int f = 0; f < 6; ++f) { A.crse_sol_br.mf[f].setVal(0.0); }
        }
        interpBndry(0, B, A.crse_sol_br);
    }
    if (levelbcdata) {
        Geometry const& geom = H.geom[amrlev][0];
        const Box domain = geom.Domain();
        std::vector<b200mg_copytag> tags;
        for (int li = 0; li < L.layout->numLocal(); ++li) {
            Box const& bx = L.layout->box(li);
            for (int f = 0; f < 6; ++f) {
                const int d = f % 3; const bool low = f < 3;
                const int dface = low ? domain.smallEnd(d) : domain.bigEnd(d);
                const int bface = low ? bx.smallEnd(d) : bx.bigEnd(d);
                if (dface == bface && !geom.isPeriodic(d)) {
                    Box const& sb = B.box(li, f);
                    b200mg_copytag t;
                    for (int x = 0; x < 3; ++x) { t.lo[x] = sb.smallEnd(x); t.hi[x] = sb.bigEnd(x); t.shift[x] = 0; }
                    t.dst_fab = li * 6 + f; t.src_fab = li; t.pad = 0; t.buf_offset = 0;
                    tags.push_back(t);
                }
            }
        }
        if (!tags.empty()) {
            DeviceTable<b200mg_copytag> dt(tags);
            B200_KCALL(b200mg_copy_tags(int(tags.size()), dt.data(), B.d_table(), levelbcdata->d_fabs(), nullptr, 1, 0, 0, 0, 0, Gpu::gpuStream()));
            Gpu::streamSynchronize();
        }
    }
    for (int m = 0; m < H.num_mg_levels[amrlev]; ++m) { buildBCFaces(amrlev, m); }
    if (hasRobinBC()) {
        // m_robin_bcval (AMReX_MLCellLinOp.H:596-640): a, b, f of the ghost cells outside the Robin faces of the domain
        AMREX_ALWAYS_ASSERT_WITH_MESSAGE(robinbc_a != nullptr && robinbc_b != nullptr && robinbc_f != nullptr,
                                         "MLLinOp::setLevelBC: Robin BC needs robinbc_a, robinbc_b and robinbc_f");
        if (int(m_robin.size()) < H.num_amr_levels) { m_robin.resize(H.num_amr_levels); }
        Geometry const& geom = H.geom[amrlev][0];
        const Box domain = geom.Domain();
        const MultiFab* src[3] = {robinbc_a, robinbc_b, robinbc_f};
        for (int q = 0; q < 3; ++q) {
            AMREX_ALWAYS_ASSERT_WITH_MESSAGE(src[q]->nGrow() >= 1 && src[q]->boxArray() == H.grids[amrlev][0] && src[q]->DistributionMap() == H.dmap[amrlev][0],
                                             "MLLinOp::setLevelBC: Robin data must live on the level's grids with one ghost cell");
            m_robin[amrlev][q] = std::make_unique<BndrySlabs<double>>();
            BndrySlabs<double>& B = *m_robin[amrlev][q];
            B.define(*L.layout, false);
            std::vector<b200mg_copytag> tags;
            for (int li = 0; li < L.layout->numLocal(); ++li) {
                Box const& bx = L.layout->box(li);
                for (int f = 0; f < 6; ++f) {
                    const int d = f % 3; const bool low = f < 3;
                    const bool robin = (low ? m_lobc_orig[d] : m_hibc_orig[d]) == BCType::Robin;
                    const int dface = low ? domain.smallEnd(d) : domain.bigEnd(d);
                    const int bface = low ? bx.smallEnd(d) : bx.bigEnd(d);
                    if (robin && dface == bface && !geom.isPeriodic(d)) {
                        Box const& sb = B.box(li, f);
                        b200mg_copytag t;
                        for (int x = 0; x < 3; ++x) { t.lo[x] = sb.smallEnd(x); t.hi[x] = sb.bigEnd(x); t.shift[x] = 0; }
                        t.dst_fab = li * 6 + f; t.src_fab = li; t.pad = 0; t.buf_offset = 0;
                        tags.push_back(t);
                    }
                }
            }
            if (!tags.empty()) {
                DeviceTable<b200mg_copytag> dt(tags);
                B200_KCALL(b200mg_copy_tags(int(tags.size()), dt.data(), B.d_table(), src[q]->d_fabs(), nullptr, 1, 0, 0, 0, 0, Gpu::gpuStream()));
                Gpu::streamSynchronize();
            }
        }
    }
}

// ---- single-kernel BiCGStab bottom solve (kernels/bottom.cu)
bool MLLinOp::bottomKernelEligible (int mglev) const
{
    if (!m_bottom_kernel || Gpu::debugSync()) { return false; }
    BoxArray const& ba = H.grids[0][mglev];
    Geometry const& geom = H.geom[0][mglev];
    if (ba.size() != 1 || ba[0] != geom.Domain() || ba[0].numPts() > 32768) { return false; }
    int nper = 0;
    for (int d = 0; d < 3; ++d) { if (geom.isPeriodic(d)) { ++nper; } }
    LevelData const& L = lev(0, mglev);
    if (L.layout->numLocal() == 1) {
        if (int(L.bcfaces_h.size()) != 6 - 2 * nper) { return false; }
        for (auto const& fc : L.bcfaces_h) { if (fc.box != 0 || fc.bctype < 101 || fc.bctype > 103) { return false; } }
    }
    return true;
}

bool MLLinOp::coarseLegLevelEligible (int mglev, Long max_cells) const
{
    if (!m_coarse_leg || !m_use_gauss_seidel || Gpu::debugSync() || m_needs_coarse_data_for_bc) { return false; }
    BoxArray const& ba = H.grids[0][mglev];
    Geometry const& geom = H.geom[0][mglev];
    if (ba.size() != 1 || ba[0] != geom.Domain() || ba[0].numPts() > max_cells || !ba[0].cellCentered()) { return false; }
    for (int d = 0; d < 3; ++d) {
        if (geom.isPeriodic(d)) { continue; }
        for (BCType t : {m_lobc[d], m_hibc[d]}) {
            if (t != BCType::Dirichlet && t != BCType::Neumann && t != BCType::reflect_odd) { return false; }
        }
    }
    if (mglev + 1 < H.num_mg_levels[0] && H.mg_coarsen_ratio_vec[mglev] != IntVect(2)) { return false; }
    return true;
}

Real MLLinOp::bottomVolInv () const
{
    ensureVolInv();
    return m_volinv[0][H.num_mg_levels[0] - 1];
}

void MLLinOp::fillLegLevel (int mglev, b200mg_leg_level& out) const
{
    LevelData const& L = lev(0, mglev);
    AMREX_ALWAYS_ASSERT(L.layout->numLocal() == 1);
    Geometry const& geom = H.geom[0][mglev];
    Box const& bx = L.layout->box(0);
    for (int d = 0; d < 3; ++d) { out.vb.lo[d] = bx.smallEnd(d); out.vb.hi[d] = bx.bigEnd(d); out.periodic[d] = geom.isPeriodic(d) ? 1 : 0; }
    const MultiFab* a = nullptr; Array<MultiFab const*, 3> b{{nullptr, nullptr, nullptr}}; Real alpha = 0.0, beta = 1.0;
    getLevelCoeffs(0, mglev, a, b, alpha, beta);
    const Real* dxi = geom.InvCellSize();
    const Real* h = geom.CellSize();
    for (int d = 0; d < 3; ++d) {
        out.dxi[d] = dxi[d];
        out.adh[d] = beta * dxi[d] * dxi[d];                               // Fapply: beta*dxinv^2 (AMReX_MLABecLap_3D_K.H:18-20)
        out.dh[d] = a ? beta / (h[d] * h[d]) : dxi[d] * dxi[d];            // Fsmooth: beta/h^2 (AMReX_MLABecLaplacian.H:907-909)
    }
    if (a) { out.a = a->desc(0); out.bx = b[0]->desc(0); out.by = b[1]->desc(0); out.bz = b[2]->desc(0); }
    else { std::memset(&out.a, 0, sizeof(out.a)); out.bx = out.by = out.bz = out.a; }
    for (int f = 0; f < 6; ++f) { out.f[f] = L.undrrelxr.h_table()[f]; out.m[f] = L.mask.h_table()[f]; }
    AMREX_ALWAYS_ASSERT(L.bcfaces_h.size() <= 6);
    out.nfaces = int(L.bcfaces_h.size());
    std::memset(out.faces, 0, sizeof(out.faces));
    for (int n = 0; n < out.nfaces; ++n) {
        out.faces[n] = L.bcfaces_h[n];
        AMREX_ALWAYS_ASSERT(out.faces[n].box == 0 && out.faces[n].bctype >= 101 && out.faces[n].bctype <= 103);
    }
}

void MLLinOp::copyOptionsTo (MLLinOp& op) const
{
    op.setMaxOrder(maxorder);
    op.setGaussSeidel(m_use_gauss_seidel);
    op.setEnforceSingularSolvable(enforceSingularSolvable);
    op.setSmootherFusion(m_fuse_colors);
    op.setDomainBC(m_lobc, m_hibc);           // (inhomogeneous Neumann already acts as Neumann below the finest level)
    op.m_domain_bloc_lo = m_domain_bloc_lo; op.m_domain_bloc_hi = m_domain_bloc_hi;
    op.setLevelBC(0, nullptr);
}

void MLLinOp::buildMergedLeg ()
{
    m_merged.reset(); m_merged_lev = -1;
    const bool off = std::getenv("B200MG_NO_MERGED_LEG") != nullptr;            // read per build: tests toggle it
    if (off || m_is_merged_copy || !m_coarse_leg || !m_use_gauss_seidel || Gpu::debugSync() || m_needs_coarse_data_for_bc) { return; }
    if (!H.domain_covered[0]) { return; }
    for (int d = 0; d < 3; ++d) {
        if (H.geom[0][0].isPeriodic(d)) { continue; }
        for (BCType t : {m_lobc[d], m_hibc[d]}) {
            if (t != BCType::Dirichlet && t != BCType::Neumann && t != BCType::reflect_odd) { return; }
        }
    }
    const char* mc = std::getenv("B200MG_MERGED_MAX_CELLS");
    const Long max_cells = mc ? Long(std::atoll(mc)) : (Long(1) << 21);
    const int nm = H.num_mg_levels[0];
    int m0 = -1;
    for (int m = 0; m < nm; ++m) {
        if (H.grids[0][m].size() == 1) { break; }                  // one box already: the leg kernel takes the level as it is
        if (H.geom[0][m].Domain().numPts() <= max_cells && H.grids[0][m].ixType().cellCentered()) { m0 = m; break; }
    }
    if (m0 < 0) { return; }
    for (int m = m0; m < nm; ++m) {
        if (m + 1 < nm && H.mg_coarsen_ratio_vec[m] != IntVect(2)) { return; }
        BoxArray const& ba = H.grids[0][m];
        if (ba.numPts() != H.geom[0][m].Domain().numPts()) { return; }
        if (ba.size() == 1) { continue; }
        // the order of the Dirichlet ghost-value interpolation is min(box length + 1, maxorder): a chopped level and its
        // one-box copy agree only where no box is shorter than that
        for (int i = 0, N = int(ba.size()); i < N; ++i) {
            for (int d = 0; d < 3; ++d) { if (ba[i].length(d) + 1 < maxorder) { return; } }
        }
    }
    const int owner = (H.grids[0][nm - 1].size() == 1) ? H.dmap[0][nm - 1][0] : 0;
    LPInfo inf;
    inf.setAgglomeration(false).setConsolidation(false).setMaxCoarseningLevel(nm - 1 - m0);
    std::unique_ptr<MLLinOp> op = makeMergedOp(H.geom[0][m0], BoxArray(H.geom[0][m0].Domain()), DistributionMapping(Vector<int>{owner}), inf, m0);
    if (!op || op->NMGLevels(0) != nm - m0) { return; }
    for (int l = 0; l < nm - m0; ++l) { if (op->Geom(0, l).Domain() != H.geom[0][m0 + l].Domain()) { return; } }
    if (op->isBottomSingular() != isBottomSingular()) { return; }
    m_merged = std::move(op); m_merged_lev = m0;
}

int MLLinOp::bottomBiCGStabKernel (int mglev, MultiFab& sol, MultiFab const& rhs, MultiFab& r, MultiFab& p, MultiFab& v, MultiFab& t,
                                   MultiFab& rh, Real eps_rel, Real eps_abs, int maxiter, int& iter) const
{
    Gpu::ProfScope prof_scope__(mglev);
    LevelData const& L = lev(0, mglev);
    double res[2] = {0.0, 0.0};                                      // return code, iterations (0 on ranks without the box)
    if (L.layout->numLocal() == 1) {
        static double* d_out = nullptr; static double* h_out = nullptr;
        if (!d_out) {
            d_out = static_cast<double*>(The_Arena()->alloc(4 * sizeof(double)));
            h_out = static_cast<double*>(pinned_alloc(4 * sizeof(double)));
        }
        const MultiFab* a = nullptr; Array<MultiFab const*, 3> b{{nullptr, nullptr, nullptr}}; Real alpha = 0.0, beta = 1.0;
        getLevelCoeffs(0, mglev, a, b, alpha, beta);
        const bool abec = (a != nullptr);
        const Real* dxi = H.geom[0][mglev].InvCellSize();
        Box const& bx = L.layout->box(0);
        b200mg_box vb; for (int d = 0; d < 3; ++d) { vb.lo[d] = bx.smallEnd(d); vb.hi[d] = bx.bigEnd(d); }
        const int per[3] = {H.geom[0][mglev].isPeriodic(0) ? 1 : 0, H.geom[0][mglev].isPeriodic(1) ? 1 : 0, H.geom[0][mglev].isPeriodic(2) ? 1 : 0};
        B200_KCALL(b200mg_bottom_bicgstab(abec ? 1 : 0, &vb, &sol.desc(0), &rhs.desc(0), &r.desc(0), &p.desc(0), &v.desc(0), &t.desc(0), &rh.desc(0),
                                          abec ? &a->desc(0) : nullptr, abec ? &b[0]->desc(0) : nullptr, abec ? &b[1]->desc(0) : nullptr,
                                          abec ? &b[2]->desc(0) : nullptr, alpha, beta * dxi[0] * dxi[0], beta * dxi[1] * dxi[1], beta * dxi[2] * dxi[2],
                                          int(L.bcfaces_h.size()), L.bcfaces_h.data(), L.mask.h_table(), per, maxorder, dxi[0], dxi[1], dxi[2],
                                          eps_rel, eps_abs, maxiter, d_out, Gpu::gpuStream()));
        Gpu::dtoh_memcpy_async(h_out, d_out, 4 * sizeof(double));
        Gpu::streamSynchronize();
        res[0] = h_out[0]; res[1] = h_out[1];
    }
    ParallelDescriptor::ReduceRealMax(res, 2);                        // the owner's values reach every rank
    iter = int(res[1]);
    return int(res[0]);
}

bool MLLinOp::isMFIterSafe (int amrlev, int mglev1, int mglev2) const
{
    if (!(H.dmap[amrlev][mglev1] == H.dmap[amrlev][mglev2])) { return false; }
    BoxArray const& f = H.grids[amrlev][mglev1];
    BoxArray const& c = H.grids[amrlev][mglev2];
    if (f.size() != c.size()) { return false; }
    const IntVect r = (amrlev > 0) ? IntVect(2) : H.mg_coarsen_ratio_vec[mglev1];
    for (int i = 0, N = int(f.size()); i < N; ++i) { if (amrex::coarsen(f[i], r) != c[i]) { return false; } }
    return true;
}

// MLCellLinOpT::prepareForSolve (AMReX_MLCellLinOp.H:1617-1926): fill the relaxation-coefficient slabs
void MLLinOp::prepareForSolve ()
{
    for (int a = 0; a < H.num_amr_levels; ++a) {
        for (int m = 0; m < H.num_mg_levels[a]; ++m) {
            LevelData& L = lev(a, m);
            AMREX_ALWAYS_ASSERT_WITH_MESSAGE(L.bcond.size() == std::size_t(L.layout->numLocal()), "setLevelBC must be called on every AMR level before solve");
            // every face (not only uncovered ones): Neumann faces get 1 regardless of the mask
            std::vector<b200mg_bcface> all;
            for (int li = 0; li < L.layout->numLocal(); ++li) for (int f = 0; f < 6; ++f) {
                b200mg_bcface fc; fc.box = li; fc.face = f; fc.bctype = L.bcond[li][f]; fc.blen = L.layout->box(li).length(f % 3); fc.bcloc = L.bcloc[li][f];
                all.push_back(fc);
            }
            if (all.empty()) { continue; }
            DeviceTable<b200mg_bcface> dall(all);
            const Real* dxi = H.geom[a][m].InvCellSize();
            B200_KCALL(b200mg_comp_interp_coef0(int(all.size()), dall.data(), L.layout->d_vbox(), L.undrrelxr.d_table(), L.mask.d_table(),
                                                maxorder, dxi[0], dxi[1], dxi[2], Gpu::gpuStream()));
            Gpu::streamSynchronize();
        }
    }
}

// MLCellLinOpT::applyBC (AMReX_MLCellLinOp.H:684-893), cross stencil
void MLLinOp::applyBC (int amrlev, int mglev, MultiFab& in, BCMode bc_mode, StateMode, const BndrySlabs<double>* bndry,
                       bool skip_fillboundary, bool nowait, int halo_parity, bool halo_remote_only) const
{
    if (!m_colour_halo) { halo_parity = -1; }
    Gpu::ProfScope prof_scope__(amrlev * 100 + mglev);
    AMREX_ALWAYS_ASSERT(mglev == 0 || bc_mode == BCMode::Homogeneous);
    AMREX_ALWAYS_ASSERT(bndry != nullptr || bc_mode == BCMode::Homogeneous);
    LevelData const& L = lev(amrlev, mglev);
    const int nf = int(L.bcfaces_h.size());
    const Real* dxi = H.geom[amrlev][mglev].InvCellSize();
    const int flagbc = (bc_mode == BCMode::Inhomogeneous);
    // The physical-boundary fill reads valid cells only and writes ghost cells outside the domain; the halo exchange
    // writes ghost cells inside it: the two are independent (cross stencil) and both are latency-bound O(n^2) launches,
    // so on big levels the BC kernel runs on a second stream next to the halo copies (fork / join through events).
    bool concurrent = false;
    if (!skip_fillboundary && nf > 0 && m_bc_overlap && maxFaceCells(L) >= 4096 && !Gpu::profiling() && !Gpu::debugSync()) {
        cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
        concurrent = (cudaStreamIsCapturing(Gpu::gpuStream(), &cs) == cudaSuccess) && cs == cudaStreamCaptureStatusNone;
    }
    if (concurrent) {
        static cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
        if (!ev_fork) {
            AMREX_CUDA_SAFE_CALL(cudaEventCreateWithFlags(&ev_fork, cudaEventDisableTiming));
            AMREX_CUDA_SAFE_CALL(cudaEventCreateWithFlags(&ev_join, cudaEventDisableTiming));
        }
        cudaStream_t s = Gpu::gpuStream(), aux = Gpu::auxStream();
        AMREX_CUDA_SAFE_CALL(cudaEventRecord(ev_fork, s));
        AMREX_CUDA_SAFE_CALL(cudaStreamWaitEvent(aux, ev_fork, 0));
        B200_KCALL(b200mg_apply_bc(nf, L.bcfaces.data(), L.layout->d_vbox(), in.d_fabs(), L.mask.d_table(),
                                   bndry ? bndry->d_table() : nullptr, maxorder, dxi[0], dxi[1], dxi[2], flagbc, maxFaceCells(L), aux));
        AMREX_CUDA_SAFE_CALL(cudaEventRecord(ev_join, aux));
        if (nowait) { in.FillBoundary_nowait(0, 1, IntVect(1), H.geom[amrlev][mglev].periodicity(), true, halo_parity, halo_remote_only); }
        else { in.FillBoundary(0, 1, IntVect(1), H.geom[amrlev][mglev].periodicity(), true, halo_parity, halo_remote_only); }
        AMREX_CUDA_SAFE_CALL(cudaStreamWaitEvent(s, ev_join, 0));
        return;
    }
    if (!skip_fillboundary) {
        if (nowait) { in.FillBoundary_nowait(0, 1, IntVect(1), H.geom[amrlev][mglev].periodicity(), true, halo_parity, halo_remote_only); }
        else { in.FillBoundary(0, 1, IntVect(1), H.geom[amrlev][mglev].periodicity(), true, halo_parity, halo_remote_only); }
    }
    if (nf == 0) { return; }
    B200_KCALL(b200mg_apply_bc(nf, L.bcfaces.data(), L.layout->d_vbox(), in.d_fabs(), L.mask.d_table(),
                               bndry ? bndry->d_table() : nullptr, maxorder, dxi[0], dxi[1], dxi[2], flagbc, maxFaceCells(L), Gpu::gpuStream()));
}

void MLLinOp::apply (int amrlev, int mglev, MultiFab& out, MultiFab& in, BCMode bc_mode, StateMode s_mode, const BndrySlabs<double>* bndry) const
{
    Gpu::ProfScope prof_scope__(amrlev * 100 + mglev);
    applyBC(amrlev, mglev, in, bc_mode, s_mode, bndry);
    Fapply(amrlev, mglev, out, in);
}

// MLCellLinOpT::smooth (AMReX_MLCellLinOp.H:1206-1217): for each colour, refresh halos + BCs, then sweep.
// Fused schedule (default): identical arithmetic and identical update order, but the red sweep and the part of the
// black sweep that cannot depend on other boxes' red values run in ONE pass over memory (out of place); the black
// sweep on the 1-cell surface shell runs after the second halo refresh.
// Tile table of the fused smoother for one level.  Eligible: every local box has an even x extent <= 256 and the level
// is big enough for the tiling to pay (small levels are launch-latency bound either way).
// cells of the largest face of the level's local boxes (grid sizing of the O(n^2) kernels: BC fill, surface shell)
int MLLinOp::maxFaceCells (LevelData const& L) const
{
    if (L.max_face_cells < 0) {
        int m = 1;
        for (int li = 0; li < L.layout->numLocal(); ++li) {
            Box const& b = L.layout->box(li);
            const int n0 = b.length(0), n1 = b.length(1), n2 = b.length(2);
            m = std::max(m, std::max(n0 * n1, std::max(n0 * n2, n1 * n2)));
        }
        L.max_face_cells = m;
    }
    return L.max_face_cells;
}

bool MLLinOp::planFused (LevelData const& L) const
{
    if (L.fused_state >= 0) { return L.fused_state == 1; }
    L.fused_state = 0;
    // The fused pass defers the black cells of the 1-cell surface shell until after the second boundary fill.  With
    // maxorder 4 that fill reads x[lo+2], an interior black cell the pass has already updated where the reference still
    // sees its pre-sweep value: only orders <= 3 give the bits of two colour sweeps.
    if (maxorder > 3) { return false; }
    const int nl = L.layout->numLocal();
    if (nl == 0 || L.layout->localCells() < Long(32768) * Long(nl)) { return false; }
    if (m_fused_min_box_cells > 0) {
        // explicit threshold (tests): fused wherever the average local box has at least this many cells
        if (L.layout->localCells() < m_fused_min_box_cells * Long(nl)) { return false; }
    } else {
        // Cost model fitted to the measured launches (profiles/r01_s25_kernel_times_per_level_1gpu_fused4.txt, 512^3 on
        // 1 / 2 GPUs): the fused pass (generation 4) runs ceil(CTAs / resident CTAs) rounds of (nz+1) barrier-separated
        // steps of ~1.8 us (one 576..640-thread CTA per SM: 8 rows at nx = 128, 16 rows at nx = 64) + the surface
        // shell; a colour sweep takes cells x 44 B / 4.65 TB/s + ~40 us of ramp and tail.  Few boxes per GPU or short boxes favour the sweeps.
        int nx0 = 0, ny0 = 0, nz0 = 0; Long surf = 0;
        for (int li = 0; li < nl; ++li) {
            Box const& b = L.layout->box(li);
            nx0 = std::max(nx0, b.length(0)); ny0 = std::max(ny0, b.length(1)); nz0 = std::max(nz0, b.length(2));
            surf += 2 * (Long(b.length(0)) * b.length(1) + Long(b.length(0)) * b.length(2) + Long(b.length(1)) * b.length(2));
        }
        if (nx0 > 128 || nx0 < 64) { return false; }
        // one CTA per SM: 8 rows x 128 cells, or 16 rows x 64 cells (b200mg_gsrb4 picks the tile from nx)
        const int tile_y = (nx0 > 64) ? 8 : 16;
        const double ctas = double(nl) * double((ny0 + tile_y - 1) / tile_y);
        const double rounds = std::ceil(ctas / 148.0);
        const double fused_us = rounds * (nz0 + 1) * 1.8 + 320.0 * double(surf) / 6.3e6 + 5.0;
        const double pairs_us = 2.0 * (double(L.layout->localCells()) * 44.0 / 4.65e6 + 40.0);
        if (fused_us > 0.95 * pairs_us) { return false; }
    }
    for (int li = 0; li < nl; ++li) {
        Box const& b = L.layout->box(li);
        if (b.length(0) % 2 != 0 || b.length(0) > 128 || b.length(0) < 16 || b.length(1) < 2) { return false; }
    }
    L.h_vbox.resize(nl);
    for (int li = 0; li < nl; ++li) {
        Box const& b = L.layout->box(li);
        for (int d = 0; d < 3; ++d) { L.h_vbox[li].lo[d] = b.smallEnd(d); L.h_vbox[li].hi[d] = b.bigEnd(d); }
    }
    L.fused_state = 1;
    return true;
}


namespace {
inline void hash_mix (std::size_t& h, std::size_t v) { h ^= v; h *= 1099511628211ull; }
inline std::size_t bits_of (Real v) { std::size_t b = 0; std::memcpy(&b, &v, sizeof(Real) < sizeof(b) ? sizeof(Real) : sizeof(b)); return b; }
}

std::size_t MLLinOp::graphKey (int amrlev, int mglev) const
{
    LevelData const& L = lev(amrlev, mglev);
    std::size_t h = 1469598103934665603ull;
    hash_mix(h, reinterpret_cast<std::size_t>(L.mask.d_table()));
    hash_mix(h, reinterpret_cast<std::size_t>(L.undrrelxr.d_table()));
    hash_mix(h, reinterpret_cast<std::size_t>(L.layout.get()));
    hash_mix(h, std::size_t(m_fuse_colors)); hash_mix(h, std::size_t(m_use_gauss_seidel));
    hash_mix(h, std::size_t(maxorder));
    return h;
}

std::size_t MLABecLaplacian::graphKey (int amrlev, int mglev) const
{
    std::size_t h = MLLinOp::graphKey(amrlev, mglev);
    hash_mix(h, reinterpret_cast<std::size_t>(m_a_coeffs[amrlev][mglev].dataPtr()));
    for (int d = 0; d < 3; ++d) { hash_mix(h, reinterpret_cast<std::size_t>(m_b_coeffs[amrlev][mglev][d].dataPtr())); }
    hash_mix(h, bits_of(m_a_scalar)); hash_mix(h, bits_of(m_b_scalar));
    return h;
}

void MLLinOp::smooth (int amrlev, int mglev, MultiFab& sol, MultiFab const& rhs, bool skip_fillboundary, bool zero_input,
                      bool consecutive) const
{
    Gpu::ProfScope prof_scope__(amrlev * 100 + mglev);
    LevelData const& L = lev(amrlev, mglev);
    const bool fuse = m_fuse_colors && m_use_gauss_seidel && planFused(L) && L.fused4_ok != 0;
    // did the previous smooth of this loop leave the black surface values in the neighbours' ghost cells (face links)?
    const bool ghosts_pushed = consecutive && L.shell_pushed;
    L.shell_pushed = false;
    // zero input without a prior setVal: the fused pass can do without reading sol
    const bool zero4 = zero_input && fuse && m_zero_input_opt;
    if (zero_input && !zero4) { sol.setVal(0.0); skip_fillboundary = true; }
    if (!m_use_gauss_seidel) {
        // Jacobi: the reference's two "colour" calls are two full damped sweeps (AMReX_MLCellLinOp.H:1206-1217 with
        // Fsmooth ignoring redblack); each is one out-of-place pass that ping-pongs with the level's scratch field
        if (!L.scratch) { L.scratch = std::make_unique<MultiFab>(sol.boxArray(), sol.DistributionMap(), 1, sol.nGrow()); }
        for (int sweep = 0; sweep < 2; ++sweep) {
            applyBC(amrlev, mglev, sol, BCMode::Homogeneous, StateMode::Solution, nullptr, skip_fillboundary);
            Fjacobi(amrlev, mglev, *L.scratch, sol, rhs);
            sol.swap(*L.scratch);
            skip_fillboundary = false;
        }
        return;
    }
    if (fuse) {
        if (!L.scratch) { L.scratch = std::make_unique<MultiFab>(sol.boxArray(), sol.DistributionMap(), 1, sol.nGrow()); }
        // (homogeneous BCs of a zero field are zero ghost cells: nothing to fill, and the pass does not read them)
        // FillBoundary overlapped with interior work (FillBoundary_nowait / _finish of the reference, AMReX_FabArray.H:1023-1042):
        // the pass over the boxes whose halo comes from this GPU only runs while the NVLink transfer is in flight, the pass
        // over the boxes with remote neighbours after the unpack.  Out of place, so the two halves do not interact.
        // Face links (a level whose exchange between boxes of this GPU is nothing but whole faces): the shell reads the red
        // cells behind a linked face from the neighbouring box itself, so the exchange ahead of it only has to bring the
        // faces other GPUs feed, and it stores its black results into the neighbours' ghost cells as well, so the same holds
        // for the exchange ahead of the NEXT pass when that follows at once (consecutive).  Same values, same bits.
        if (L.face_links_state < 0) {
            L.face_links = (m_face_links && L.fused4_ok != 0) ? sol.FillBoundaryFaceLinks(IntVect(1), H.geom[amrlev][mglev].periodicity(), true) : nullptr;
            L.face_links_state = (L.face_links != nullptr) ? 1 : 0;
        }
        const b200mg_facelink* links = (L.face_links_state == 1 && L.fused4_ok == 1) ? L.face_links : nullptr;
        bool split = false;
        // (not under the per-kernel profiler: it times whole-level launches on the main stream)
        if (!zero4 && !skip_fillboundary && m_halo_overlap && L.fused4_ok == 1 && ParallelDescriptor::NProcs() > 1 && !Gpu::debugSync()
            && !Gpu::profiling()) {
            if (L.halo_split < 0) {
                auto const& rf = sol.FillBoundaryRemoteFabs(IntVect(1), H.geom[amrlev][mglev].periodicity(), true);
                L.boxes_remote_halo = rf; L.boxes_local_halo.clear();
                for (int li = 0, q = 0; li < L.layout->numLocal(); ++li) {
                    if (q < int(rf.size()) && rf[q] == li) { ++q; } else { L.boxes_local_halo.push_back(li); }
                }
                L.halo_split = (!L.boxes_remote_halo.empty() && !L.boxes_local_halo.empty()) ? 1 : 0;
            }
            split = (L.halo_split == 1);
        }
        if (split) {
            // main stream: physical-boundary fill, intra-GPU copies, then the pass over the boxes whose halo is complete;
            // second stream: wait for the transfer, unpack, the pass over the other boxes - the two passes share the SMs, so
            // there is no tail between them, and the first one hides the transfer
            static cudaEvent_t ev_ready = nullptr, ev_done = nullptr;
            if (!ev_ready) {
                AMREX_CUDA_SAFE_CALL(cudaEventCreateWithFlags(&ev_ready, cudaEventDisableTiming));
                AMREX_CUDA_SAFE_CALL(cudaEventCreateWithFlags(&ev_done, cudaEventDisableTiming));
            }
            applyBC(amrlev, mglev, sol, BCMode::Homogeneous, StateMode::Solution, nullptr, false, true, 1, links && ghosts_pushed);
            cudaStream_t s = Gpu::gpuStream(), aux = Gpu::auxStream();
            const bool side = true;
            bool ok2 = true;
            if (side) {
                AMREX_CUDA_SAFE_CALL(cudaEventRecord(ev_ready, s));
                AMREX_CUDA_SAFE_CALL(cudaStreamWaitEvent(aux, ev_ready, 0));
                Gpu::setStream(aux);
                sol.FillBoundary_finish();
                ok2 = Fsmooth2(amrlev, mglev, *L.scratch, sol, rhs, false, &L.boxes_remote_halo);
                Gpu::setStream(nullptr);
                AMREX_CUDA_SAFE_CALL(cudaEventRecord(ev_done, aux));
            }
            const bool ok1 = Fsmooth2(amrlev, mglev, *L.scratch, sol, rhs, false, &L.boxes_local_halo);
            if (side) { AMREX_CUDA_SAFE_CALL(cudaStreamWaitEvent(s, ev_done, 0)); }
            else {
                sol.FillBoundary_finish();
                ok2 = Fsmooth2(amrlev, mglev, *L.scratch, sol, rhs, false, &L.boxes_remote_halo);
            }
            AMREX_ALWAYS_ASSERT_WITH_MESSAGE(ok1 && ok2, "fused smoother: a level that took the pass before refused it");
            sol.swap(*L.scratch);
            applyBC(amrlev, mglev, sol, BCMode::Homogeneous, StateMode::Solution, nullptr, false, false, 0, links != nullptr);
            FsmoothShell(amrlev, mglev, sol, rhs, 1, links, links != nullptr);
            L.shell_pushed = (links != nullptr);
            return;
        }
        // the pass updates the red cells everywhere (they read black ghost cells: parity 1) and the black cells off the box
        // surface (no ghost cells); the black shell afterwards reads red ghost cells (parity 0)
        if (!zero4) {
            applyBC(amrlev, mglev, sol, BCMode::Homogeneous, StateMode::Solution, nullptr, skip_fillboundary, false, 1, links && ghosts_pushed);
        }
        if (Fsmooth2(amrlev, mglev, *L.scratch, sol, rhs, zero4)) {
            L.fused4_ok = 1;
            sol.swap(*L.scratch);
            applyBC(amrlev, mglev, sol, BCMode::Homogeneous, StateMode::Solution, nullptr, false, false, 0, links != nullptr);
            FsmoothShell(amrlev, mglev, sol, rhs, 1, links, links != nullptr);
            L.shell_pushed = (links != nullptr);
            return;
        }
        // the layout cannot take the bulk copies (L.fused4_ok is 0 from now on): colour sweeps; the boundary fill above is
        // repeated below, which is harmless
        if (zero4) { sol.setVal(0.0); skip_fillboundary = true; }
    }
    for (int redblack = 0; redblack < 2; ++redblack) {
        // the sweep of colour redblack updates cells with (i+j+k+redblack) even and reads their neighbours: the other colour
        applyBC(amrlev, mglev, sol, BCMode::Homogeneous, StateMode::Solution, nullptr, skip_fillboundary, false, 1 - redblack);
        Fsmooth(amrlev, mglev, sol, rhs, redblack);
        skip_fillboundary = false;
    }
}

bool MLLinOp::solutionResidual (int amrlev, MultiFab& resid, MultiFab& x, MultiFab const& b, const MultiFab* crse_bcdata, Real* resnorm)
{
    Gpu::ProfScope prof_scope__(amrlev * 100 + 0);
    if (crse_bcdata != nullptr) { updateSolBC(amrlev, *crse_bcdata); }
    applyBC(amrlev, 0, x, BCMode::Inhomogeneous, StateMode::Solution, m_bndry_sol[amrlev].get());
    // resid = b - L(x), fused (== Fapply + Xpay(resid,-1,b)); the unmasked inf-norm rides along when asked for
    const bool want = resnorm != nullptr && m_fuse_resnorm && amrlev == H.num_amr_levels - 1;
    const bool got = Fapply(amrlev, 0, resid, x, &b, want ? reduce_result_slot(1) : nullptr);
    if (got) {
        double r = fetch_reduce_result(1);
        ParallelDescriptor::ReduceRealMax(&r, 1);
        *resnorm = r;
    }
    return got;
}

void MLLinOp::correctionResidual (int amrlev, int mglev, MultiFab& resid, MultiFab& x, MultiFab const& b, BCMode bc_mode, const MultiFab* crse_bcdata)
{
    Gpu::ProfScope prof_scope__(amrlev * 100 + mglev);
    if (bc_mode == BCMode::Inhomogeneous) {
        if (crse_bcdata) { AMREX_ALWAYS_ASSERT(mglev == 0 && amrlev > 0); updateCorBC(amrlev, *crse_bcdata); }
        applyBC(amrlev, mglev, x, BCMode::Inhomogeneous, StateMode::Correction, m_bndry_cor[amrlev].get());
    } else {
        applyBC(amrlev, mglev, x, BCMode::Homogeneous, StateMode::Correction, nullptr);
    }
    Fapply(amrlev, mglev, resid, x, &b);
}

bool MLLinOp::correctionResidualRestrict (int amrlev, int mglev, MultiFab& crse, MultiFab& x, MultiFab const& b)
{
    if (!m_fuse_restrict || mglev + 1 >= H.num_mg_levels[amrlev] || Gpu::debugSync()) { return false; }
    const int ratio = (amrlev > 0) ? 2 : H.mg_coarsen_ratio_vec[mglev][0];
    if (ratio != 2 || (amrlev == 0 && H.mg_coarsen_ratio_vec[mglev] != IntVect(2))) { return false; }
    LevelLayout const& FL = x.layout();               // (its tiles are 4, 8 or 16 planes deep: always whole plane pairs)
    if (!FL.pairable()) { return false; }
    LevelData const& LD = lev(amrlev, mglev);
    if (LD.even_boxes < 0) {                          // every box of the level (all ranks decide alike)
        LD.even_boxes = 1;
        BoxArray const& ba = x.boxArray();
        for (int i = 0, N = int(ba.size()); i < N && LD.even_boxes; ++i) {
            for (int d = 0; d < 3; ++d) { if (ba[i].smallEnd(d) % 2 != 0 || ba[i].length(d) % 2 != 0) { LD.even_boxes = 0; } }
        }
    }
    if (!LD.even_boxes) { return false; }
    Gpu::ProfScope prof_scope__(amrlev * 100 + mglev);
    applyBC(amrlev, mglev, x, BCMode::Homogeneous, StateMode::Correction, nullptr);
    if (isMFIterSafe(amrlev, mglev, mglev + 1)) {
        FresidualRestrict(amrlev, mglev, crse, x, b);
    } else {   // the coarse level is re-gridded: through the coarsened-fine temporary + ParallelCopy, as restriction() does
        LevelData const& L = lev(amrlev, mglev);
        if (!L.restrict_tmp) { L.restrict_tmp = std::make_unique<MultiFab>(amrex::coarsen(x.boxArray(), ratio), x.DistributionMap(), 1, 0); }
        FresidualRestrict(amrlev, mglev, *L.restrict_tmp, x, b);
        crse.ParallelCopy(*L.restrict_tmp, 0, 0, 1);
    }
    return true;
}

// ---- post-solve API: face-centred gradient and flux of the solution
void MLLinOp::compGrad (int amrlev, Array<MultiFab*, 3> const& grad, MultiFab& sol)
{
    Gpu::ProfScope prof_scope__(amrlev * 100 + 0);
    AMREX_ALWAYS_ASSERT_WITH_MESSAGE(sol.nGrow() >= 1, "compGrad: the solution needs one ghost cell");
    applyBC(amrlev, 0, sol, BCMode::Inhomogeneous, StateMode::Solution, m_bndry_sol[amrlev].get());
    const Real* dxi = H.geom[amrlev][0].InvCellSize();
    for (int d = 0; d < 3; ++d) {
        MultiFab& g = *grad[d];
        AMREX_ALWAYS_ASSERT_WITH_MESSAGE(g.boxArray() == amrex::convert(sol.boxArray(), IntVect::TheDimensionVector(d))
                                         && g.DistributionMap() == sol.DistributionMap(), "compGrad: grad[d] must live on the faces of sol's grids");
        auto const& T = g.layout().tiles(0);
        B200_KCALL(b200mg_face_flux(T.n, T.d.data(), g.layout().d_vbox(), g.d_fabs(), sol.d_fabs(), nullptr, dxi[d], 1.0, d, 0, Gpu::gpuStream()));
    }
    addInhomogNeumannFlux(amrlev, grad, sol, false);         // AMReX_MLCellLinOp.H:1441
}

void MLLinOp::compFlux (int amrlev, Array<MultiFab*, 3> const& fluxes, MultiFab& sol)
{
    Gpu::ProfScope prof_scope__(amrlev * 100 + 0);
    AMREX_ALWAYS_ASSERT_WITH_MESSAGE(sol.nGrow() >= 1, "compFlux: the solution needs one ghost cell");
    applyBC(amrlev, 0, sol, BCMode::Inhomogeneous, StateMode::Solution, m_bndry_sol[amrlev].get());
    Array<MultiFab const*, 3> b; Real bscalar;
    getFluxCoeffs(amrlev, b, bscalar);
    const Real betainv = Real(1.0) / bscalar;
    const Real* dxi = H.geom[amrlev][0].InvCellSize();
    for (int d = 0; d < 3; ++d) {
        MultiFab& f = *fluxes[d];
        AMREX_ALWAYS_ASSERT_WITH_MESSAGE(f.boxArray() == amrex::convert(sol.boxArray(), IntVect::TheDimensionVector(d))
                                         && f.DistributionMap() == sol.DistributionMap(), "compFlux: fluxes[d] must live on the faces of sol's grids");
        auto const& T = f.layout().tiles(0);
        if (b[d]) {   // ABecLap: fac = b_scalar * dxinv (MLABecLaplacianT::FFlux, AMReX_MLABecLaplacian.H:1149-1180)
            B200_KCALL(b200mg_face_flux(T.n, T.d.data(), f.layout().d_vbox(), f.d_fabs(), sol.d_fabs(), b[d]->d_fabs(),
                                        bscalar * dxi[d], betainv, d, 1, Gpu::gpuStream()));
        } else {      // Poisson: dxinv * (s - s^-), then * (1 / -1) (AMReX_MLPoisson.H:859-930, AMReX_MLCellABecLap.H:279-288)
            B200_KCALL(b200mg_face_flux(T.n, T.d.data(), f.layout().d_vbox(), f.d_fabs(), sol.d_fabs(), nullptr,
                                        dxi[d], betainv, d, 2, Gpu::gpuStream()));
        }
    }
}

void MLLinOp::getFluxes (Vector<Array<MultiFab*, 3>> const& a_flux, Vector<MultiFab*> const& a_sol)
{
    for (int alev = 0; alev < H.num_amr_levels; ++alev) {
        compFlux(alev, a_flux[alev], *a_sol[alev]);
        addInhomogNeumannFlux(alev, a_flux[alev], *a_sol[alev], true);     // AMReX_MLCellABecLap.H:289
    }
}

void MLLinOp::restriction (int amrlev, int cmglev, MultiFab& crse, MultiFab& fine) const
{
    Gpu::ProfScope prof_scope__(amrlev * 100 + (cmglev - 1));
    const int ratio = (amrlev > 0) ? 2 : H.mg_coarsen_ratio_vec[cmglev - 1][0];
    if (isMFIterSafe(amrlev, cmglev - 1, cmglev)) {
        auto const& T = crse.layout().tiles(0);
        B200_KCALL(b200mg_restrict_cc(T.n, T.d.data(), crse.layout().d_vbox(), crse.d_fabs(), fine.d_fabs(), ratio, Gpu::gpuStream()));
    } else {   // average_down through a coarsened-fine temporary + ParallelCopy (AMReX_MultiFabUtil.H:582-650)
        LevelData const& L = lev(amrlev, cmglev - 1);
        if (!L.restrict_tmp) { L.restrict_tmp = std::make_unique<MultiFab>(amrex::coarsen(fine.boxArray(), ratio), fine.DistributionMap(), 1, 0); }
        MultiFab& tmp = *L.restrict_tmp;
        auto const& T = tmp.layout().tiles(0);
        B200_KCALL(b200mg_restrict_cc(T.n, T.d.data(), tmp.layout().d_vbox(), tmp.d_fabs(), fine.d_fabs(), ratio, Gpu::gpuStream()));
        crse.ParallelCopy(tmp, 0, 0, 1);
    }
}

void MLLinOp::avgDownResMG (int clev, MultiFab& cres, MultiFab const& fres) const
{
    restriction(0, clev, cres, const_cast<MultiFab&>(fres));
}

// piecewise-constant V-cycle prolongation (AMReX_MLCellLinOp.H:956-999); crse must be on fine's coarsened layout
void MLLinOp::interpolation (int amrlev, int fmglev, MultiFab& fine, MultiFab const& crse) const
{
    Gpu::ProfScope prof_scope__(amrlev * 100 + fmglev);
    const int ratio = (amrlev > 0) ? 2 : H.mg_coarsen_ratio_vec[fmglev][0];
    AMREX_ALWAYS_ASSERT(ratio == 2);
    auto const& T = fine.layout().tiles(0);
    B200_KCALL(b200mg_prolong_add(T.n, T.d.data(), fine.layout().d_vbox(), fine.d_fabs(), crse.d_fabs(), Gpu::gpuStream()));
}

// trilinear interpolation used by the F-cycle (AMReX_MLCellLinOp.H:1003-1092)
// The reference takes the direct path only when the two levels' BoxArrays share their box list (amrex::isMFIterSafe ->
// BoxArray::SameRefs, AMReX_MLCellLinOp.H:1014): true for levels made by coarsening the user's grids, never for levels
// rebuilt by agglomeration (AMReX_MLLinOp.H:996-1020, each a fresh BoxArray).  For restriction / prolongation the two
// paths give the same bits and the geometric test (isMFIterSafe) picks the cheaper one; the trilinear F-cycle
// interpolation however reads ghost cells, and the temporary's ghost cells outside the domain are ZERO while the
// coarse field's own hold stale boundary fills - so here the reference's criterion decides.
bool MLLinOp::sharesBoxList (int amrlev, int mglev1, int mglev2) const { return H.sharesBoxList(amrlev, mglev1, mglev2); }

void MLLinOp::interpAssign (int amrlev, int fmglev, MultiFab& fine, MultiFab& crse) const
{
    Gpu::ProfScope prof_scope__(amrlev * 100 + fmglev);
    Geometry const& cgeom = H.geom[amrlev][fmglev + 1];
    const MultiFab* cmf = &crse;
    if (isMFIterSafe(amrlev, fmglev, fmglev + 1) && sharesBoxList(amrlev, fmglev, fmglev + 1)) {
        crse.FillBoundary(0, 1, IntVect(crse.nGrow()), cgeom.periodicity(), false);
    } else {
        LevelData const& L = lev(amrlev, fmglev);
        if (!L.interp_tmp) { L.interp_tmp = std::make_unique<MultiFab>(amrex::coarsen(fine.boxArray(), 2), fine.DistributionMap(), 1, crse.nGrow()); }
        L.interp_tmp->setVal(0.0);
        L.interp_tmp->ParallelCopy(crse, 0, 0, 1, 0, crse.nGrow(), cgeom.periodicity());
        cmf = L.interp_tmp.get();
    }
    auto const& T = fine.layout().tiles(0);
    B200_KCALL(b200mg_interp_cc_r2(T.n, T.d.data(), fine.layout().d_vbox(), fine.d_fabs(), cmf->d_fabs(), 0, Gpu::gpuStream()));
}


// =========================================================================== multi-level (AMR) coupling
void MLLinOp::CrseBndryReg::define (BoxArray const& cba, DistributionMapping const& dm)
{
    for (int f = 0; f < 6; ++f) {
        const Orientation face(f);
        std::vector<Box> bl;
        bl.reserve(cba.size());
        for (int i = 0, N = int(cba.size()); i < N; ++i) {
            Box b = adjCell(cba[i], face, 1);
            for (int d = 0; d < 3; ++d) { if (d != face.coordDir()) { b.grow(d, 2); } }
            bl.push_back(b);
        }
        mf[f].define(BoxArray(std::move(bl)), dm, 1, 0);
        mf[f].setVal(0.0);
    }
    const int nl = mf[0].local_size();
    std::vector<b200mg_fab> h(std::size_t(nl) * 6);
    for (int li = 0; li < nl; ++li) { for (int f = 0; f < 6; ++f) { h[li * 6 + f] = mf[f].desc(li); } }
    table.assign(h);
}

// BndryRegisterT::copyFrom (AMReX_BndryRegister.H:266-275): one ParallelCopy per face
void MLLinOp::CrseBndryReg::copyFrom (MultiFab const& crse, Periodicity const& period)
{
    for (int f = 0; f < 6; ++f) { mf[f].ParallelCopy(crse, 0, 0, 1, 0, 0, period); }
}

// The coarse/fine boundary machinery of AMR level a against data `ratio` times coarser: BndryData masks, the coarse
// boundary registers and the list of faces that take interpolated coarse data (MLCellLinOpT::defineBC,
// AMReX_MLCellLinOp.H:431-509; for level 0 built on demand by setLevelBC, :536-566)
void MLLinOp::defineAmrBndry (int a, int ratio)
{
    m_amr_bndry[a] = std::make_unique<AmrBndry>();
    AmrBndry& A = *m_amr_bndry[a];
    LevelLayout const& layout = *lev(a, 0).layout;
    Geometry const& geom = H.geom[a][0];
    // BndryData masks: in_rad 0, out_rad 2, extent NTangHalfWidth = 5 (AMReX_BndryData.H:156,265)
    A.bmask.define(layout, false, 2, 5);
    std::vector<int> h;
    fill_masks(A.bmask, layout, geom, H.grids[a][0], 5, h);
    A.bmask.upload(h);
    const BoxArray cba = amrex::coarsen(H.grids[a][0], ratio);
    A.crse_sol_br.define(cba, H.dmap[a][0]);
    A.crse_cor_br.define(cba, H.dmap[a][0]);
    // faces that get interpolated coarse data: everything but non-periodic physical boundaries
    // (InterpBndryDataT::setBndryValues, AMReX_InterpBndryData.H:177-181)
    const Box domain = geom.Domain();
    for (int li = 0; li < layout.numLocal(); ++li) {
        Box const& bx = layout.box(li);
        for (int f = 0; f < 6; ++f) {
            const int d = f % 3; const bool low = f < 3;
            const int dface = low ? domain.smallEnd(d) : domain.bigEnd(d);
            const int bface = low ? bx.smallEnd(d) : bx.bigEnd(d);
            if (bface != dface || geom.isPeriodic(d)) {
                b200mg_bcface fc; fc.box = li; fc.face = f; fc.bctype = 0; fc.blen = bx.length(d); fc.bcloc = 0.0;
                A.cf_faces_h.push_back(fc);
            }
        }
    }
    A.cf_faces.assign(A.cf_faces_h);
    A.ratio = ratio;
}

void MLLinOp::defineAmrData ()
{
    const int nlev = H.num_amr_levels;
    m_amr_bndry.resize(nlev);
    m_fluxreg.resize(std::max(0, nlev - 1));
    for (int a = 1; a < nlev; ++a) { defineAmrBndry(a, H.amr_ref_ratio[a - 1]); }

    for (int a = 0; a + 1 < nlev; ++a) {   // YAFluxRegisterT::define
        m_fluxreg[a] = std::make_unique<FluxReg>();
        FluxReg& R = *m_fluxreg[a];
        BoxArray const& cba = H.grids[a][0];
        DistributionMapping const& cdm = H.dmap[a][0];
        DistributionMapping const& fdm = H.dmap[a + 1][0];
        Geometry const& cgeom = H.geom[a][0];
        const int ratio = H.amr_ref_ratio[a];
        R.crse_data.define(cba, cdm, 1, 0);
        R.crse_flag.define(cba, cdm, 1, 1);
        const auto pshifts = cgeom.periodicity().shiftIntVect();
        const BoxArray cfba = amrex::coarsen(H.grids[a + 1][0], ratio);
        Box cdomain = cgeom.Domain();
        for (int d = 0; d < 3; ++d) { if (cgeom.isPeriodic(d)) { cdomain.grow(d, 1); } }
        std::vector<std::pair<int, Box>> isects;
        // flags: 0 coarse cell, 1 coarse cell next to the fine level, 2 covered by the fine level
        for (int li = 0; li < R.crse_flag.local_size(); ++li) {
            const Box gbx = R.crse_flag.fabbox(li);
            std::vector<int> h(gbx.numPts(), 0);
            auto fill = [&] (Box const& b, int v) {
                for (int k = b.smallEnd(2); k <= b.bigEnd(2); ++k) for (int j = b.smallEnd(1); j <= b.bigEnd(1); ++j)
                    for (int i = b.smallEnd(0); i <= b.bigEnd(0); ++i) {
                        h[(i - gbx.smallEnd(0)) + Long(gbx.length(0)) * ((j - gbx.smallEnd(1)) + Long(gbx.length(1)) * (k - gbx.smallEnd(2)))] = v;
                    }
            };
            for (int pass = 0; pass < 2; ++pass) {
                for (auto const& iv : pshifts) {
                    cfba.intersections(gbx + iv, isects, false, IntVect(pass == 0 ? 1 : 0));
                    for (auto const& is : isects) { fill(is.second - iv, pass == 0 ? 1 : 2); }
                }
            }
            R.crse_flag.copyFromHost(h.data(), gbx, 0, 1);
        }
        // coarse/fine patches: the one-cell shell around every coarsened fine box that no fine box covers
        std::vector<Box> cfp_boxes; Vector<int> cfp_pmap;
        std::vector<b200mg_box> cfbox_h; std::vector<int> findex_h;
        const int myproc = ParallelDescriptor::MyProc();
        int nlocal_fine = 0;
        for (int i = 0, N = int(cfba.size()); i < N; ++i) {
            Box bx = amrex::grow(cfba[i], 1);
            bx &= cdomain;
            const BoxList bl = cfba.complementIn(bx);
            const int proc = fdm[i];
            for (auto const& b : bl) {
                cfp_boxes.push_back(b); cfp_pmap.push_back(proc);
                if (proc == myproc) {
                    b200mg_box cb;
                    for (int d = 0; d < 3; ++d) { cb.lo[d] = cfba[i].smallEnd(d); cb.hi[d] = cfba[i].bigEnd(d); }
                    cfbox_h.push_back(cb); findex_h.push_back(nlocal_fine);
                }
            }
            if (proc == myproc) { ++nlocal_fine; }
        }
        if (!cfp_boxes.empty()) {
            const BoxArray cfp_ba{std::vector<Box>(cfp_boxes)};
            const DistributionMapping cfp_dm{Vector<int>(cfp_pmap)};
            R.cfpatch.define(cfp_ba, cfp_dm, 1, 0);
            AMREX_ALWAYS_ASSERT(R.cfpatch.local_size() == int(cfbox_h.size()));
            R.cfbox.assign(cfbox_h); R.fine_index.assign(findex_h);
            if (cgeom.isAnyPeriodic()) {   // patch cells beyond a periodic boundary that a fine box covers do not count
                R.cfp_mask = std::make_unique<MultiFab>(cfp_ba, cfp_dm, 1, 0);
                const Box domainbox = cgeom.Domain();
                for (int li = 0; li < R.cfp_mask->local_size(); ++li) {
                    const Box bx = R.cfp_mask->validbox(li);
                    std::vector<double> h(bx.numPts(), 1.0);
                    if (!domainbox.contains(bx)) {
                        for (auto const& iv : pshifts) {
                            if (iv == IntVect(0)) { continue; }
                            cfba.intersections(bx + iv, isects);
                            for (auto const& is : isects) {
                                const Box b = is.second - iv;
                                for (int k = b.smallEnd(2); k <= b.bigEnd(2); ++k) for (int j = b.smallEnd(1); j <= b.bigEnd(1); ++j)
                                    for (int i = b.smallEnd(0); i <= b.bigEnd(0); ++i) {
                                        h[(i - bx.smallEnd(0)) + Long(bx.length(0)) * ((j - bx.smallEnd(1)) + Long(bx.length(1)) * (k - bx.smallEnd(2)))] = 0.0;
                                    }
                            }
                        }
                    }
                    R.cfp_mask->copyFromHost(h.data(), bx, 0, 0);
                }
            }
        }
    }
}

// InterpBndryDataT::updateBndryValues (AMReX_InterpBndryData.H:161-268), max_order 3
void MLLinOp::interpBndry (int amrlev, BndrySlabs<double>& bndry, CrseBndryReg const& br) const
{
    AmrBndry const& A = *m_amr_bndry[amrlev];
    const int nf = int(A.cf_faces_h.size());
    if (nf == 0) { return; }
    B200_KCALL(b200mg_interp_bndry_o3(nf, A.cf_faces.data(), lev(amrlev, 0).layout->d_vbox(), bndry.d_table(), br.table.data(),
                                      A.bmask.d_table(), A.ratio, Gpu::gpuStream()));
}

void MLLinOp::updateSolBC (int amrlev, MultiFab const& crse_bcdata) const
{
    AMREX_ALWAYS_ASSERT(amrlev > 0);
    Gpu::ProfScope prof_scope__(amrlev * 100);
    AmrBndry& A = *m_amr_bndry[amrlev];
    A.crse_sol_br.copyFrom(crse_bcdata, H.geom[amrlev - 1][0].periodicity());
    interpBndry(amrlev, *m_bndry_sol[amrlev], A.crse_sol_br);
}

void MLLinOp::updateCorBC (int amrlev, MultiFab const& crse_bcdata) const
{
    AMREX_ALWAYS_ASSERT(amrlev > 0);
    Gpu::ProfScope prof_scope__(amrlev * 100);
    AmrBndry& A = *m_amr_bndry[amrlev];
    A.crse_cor_br.copyFrom(crse_bcdata, H.geom[amrlev - 1][0].periodicity());
    interpBndry(amrlev, *m_bndry_cor[amrlev], A.crse_cor_br);
}

// MLCellLinOpT::reflux (AMReX_MLCellLinOp.H:1274-1344): res(crse) += [coarse flux - average of fine fluxes] / dx on the
// coarse cells that touch the fine level from outside.
void MLLinOp::reflux (int crse_amrlev, MultiFab& res, MultiFab const& crse_sol, MultiFab& fine_sol) const
{
    Gpu::ProfScope prof_scope__(crse_amrlev * 100);
    FluxReg& R = *m_fluxreg[crse_amrlev];
    const int fine_amrlev = crse_amrlev + 1;
    const int ratio = H.amr_ref_ratio[crse_amrlev];
    applyBC(fine_amrlev, 0, fine_sol, BCMode::Inhomogeneous, StateMode::Solution, m_bndry_sol[fine_amrlev].get());
    const Real* cdx = H.geom[crse_amrlev][0].CellSize();
    const Real* fdx = H.geom[fine_amrlev][0].CellSize();
    const Real* cdxi = H.geom[crse_amrlev][0].InvCellSize();
    const Real* fdxi = H.geom[fine_amrlev][0].InvCellSize();
    const Real dt = 1.0;
    Array<MultiFab const*, 3> cb, fb; Real bscalar = 0.0;
    getFluxCoeffs(crse_amrlev, cb, bscalar);
    getFluxCoeffs(fine_amrlev, fb, bscalar);
    {
        auto const& T = R.crse_data.layout().tiles(0);
        B200_KCALL(b200mg_reflux_crse(T.n, T.d.data(), R.crse_data.layout().d_vbox(), R.crse_data.d_fabs(), R.crse_flag.d_fabs(), crse_sol.d_fabs(),
                                      cb[0] ? cb[0]->d_fabs() : nullptr, cb[1] ? cb[1]->d_fabs() : nullptr, cb[2] ? cb[2]->d_fabs() : nullptr,
                                      bscalar * cdxi[0], bscalar * cdxi[1], bscalar * cdxi[2], dt / cdx[0], dt / cdx[1], dt / cdx[2], Gpu::gpuStream()));
    }
    if (!R.cfpatch.empty()) {
        const Real r3 = Real(ratio) * ratio * ratio;
        B200_KCALL(b200mg_reflux_fine(R.cfpatch.local_size(), R.cfpatch.d_fabs(), R.cfbox.data(), R.fine_index.data(),
                                      R.cfp_mask ? R.cfp_mask->d_fabs() : nullptr, fine_sol.d_fabs(),
                                      fb[0] ? fb[0]->d_fabs() : nullptr, fb[1] ? fb[1]->d_fabs() : nullptr, fb[2] ? fb[2]->d_fabs() : nullptr,
                                      bscalar * fdxi[0], bscalar * fdxi[1], bscalar * fdxi[2],
                                      dt / (fdx[0] * r3), dt / (fdx[1] * r3), dt / (fdx[2] * r3), ratio, Gpu::gpuStream()));
        R.crse_data.ParallelCopy(R.cfpatch, 0, 0, 1, 0, 0, H.geom[crse_amrlev][0].periodicity(), CpOp::ADD);
    }
    MultiFab::Add(res, R.crse_data, 0, 0, 1, 0);
}

MultiFab MLLinOp::makeCoarseAmr (int famrlev, int ng) const
{
    return MultiFab(amrex::coarsen(H.grids[famrlev][0], H.amr_ref_ratio[famrlev - 1]), H.dmap[famrlev][0], 1, ng);
}

// trilinear cell-centred interpolation between AMR levels (AMReX_MLCellLinOp.H:1096-1181, mlmg_lin_cc_interp_r2)
void MLLinOp::interpolationAmr (int famrlev, MultiFab& fine, MultiFab const& crse) const
{
    Gpu::ProfScope prof_scope__(famrlev * 100);
    AMREX_ALWAYS_ASSERT_WITH_MESSAGE(H.amr_ref_ratio[famrlev - 1] == 2, "interpolationAmr: only refinement ratio 2 is implemented");
    auto const& T = fine.layout().tiles(0);
    B200_KCALL(b200mg_interp_cc_r2(T.n, T.d.data(), fine.layout().d_vbox(), fine.d_fabs(), crse.d_fabs(), 0, Gpu::gpuStream()));
}

void MLLinOp::avgDownResAmr (int clev, MultiFab& cres, MultiFab const& fres) const
{
    average_down(fres, cres, 0, 1, H.amr_ref_ratio[clev]);
}

Real MLLinOp::xdoty (int, int, MultiFab const& x, MultiFab const& y, bool local) const { return MultiFab::Dot(x, y, local); }

Real MLLinOp::normInf (int amrlev, MultiFab const& mf, bool local) const
{
    const int finest = H.num_amr_levels - 1;
    return (amrlev == finest) ? mf.norminf(local) : mf.norminf(*m_norm_fine_mask[amrlev], local);
}

void MLLinOp::ensureVolInv () const
{
    if (m_volinv.empty()) {   // computeVolInv, AMReX_MLCellLinOp.H:1944-2004
        m_volinv.resize(H.num_amr_levels);
        for (int a = 0; a < H.num_amr_levels; ++a) { m_volinv[a].assign(H.num_mg_levels[a], 0.0); }
        auto f = [&] (int a, int m) {
            if (m_coarse_fine_bc_type == LinOpBCType::Dirichlet) { m_volinv[a][m] = Real(1.0 / H.geom[a][m].Domain().d_numPts()); }
            else { m_volinv[a][m] = Real(1.0 / H.grids[a][m].d_numPts()); }
        };
        f(0, 0); f(0, H.num_mg_levels[0] - 1);
    }
}

Vector<Real> MLLinOp::getSolvabilityOffset (int amrlev, int mglev, MultiFab const& rhs) const
{
    ensureVolInv();
    Vector<Real> offset(1);
    offset[0] = rhs.sum(true) * m_volinv[amrlev][mglev];
    ParallelDescriptor::ReduceRealSum(offset.data(), 1);
    return offset;
}

void MLLinOp::fixSolvabilityByOffset (int, int, MultiFab& rhs, Vector<Real> const& offset) const { rhs.plus(-offset[0], 0); }

void MLLinOp::averageDownAndSync (Vector<MultiFab>& sol) const
{
    for (int falev = H.num_amr_levels - 1; falev > 0; --falev) { average_down(sol[falev], sol[falev - 1], 0, 1, H.amr_ref_ratio[falev - 1]); }
}

// ====================================================================================== MLABecLaplacian
void MLABecLaplacian::define (Vector<Geometry> const& a_geom, Vector<BoxArray> const& a_grids,
                              Vector<DistributionMapping> const& a_dmap, LPInfo const& a_info)
{
    MLLinOp::define(a_geom, a_grids, a_dmap, a_info);
    m_a_coeffs.resize(H.num_amr_levels); m_b_coeffs.resize(H.num_amr_levels);
    for (int a = 0; a < H.num_amr_levels; ++a) {
        m_a_coeffs[a].resize(H.num_mg_levels[a]); m_b_coeffs[a].resize(H.num_mg_levels[a]);
        for (int m = 0; m < H.num_mg_levels[a]; ++m) {
            m_a_coeffs[a][m].define(H.grids[a][m], H.dmap[a][m], 1, 0);
            for (int d = 0; d < 3; ++d) {
                m_b_coeffs[a][m][d].define(amrex::convert(H.grids[a][m], IntVect::TheDimensionVector(d)), H.dmap[a][m], 1, 0);
            }
        }
    }
}

void MLABecLaplacian::setScalars (Real a, Real b) noexcept
{
    m_scalars_set = true;
    m_a_scalar = a; m_b_scalar = b;
    if (a == 0.0) { for (int l = 0; l < H.num_amr_levels; ++l) { m_a_coeffs[l][0].setVal(0.0); } }
}

void MLABecLaplacian::setACoeffs (int amrlev, MultiFab const& alpha)
{
    AMREX_ALWAYS_ASSERT_WITH_MESSAGE(alpha.nComp() == 1, "MLABecLaplacian::setACoeffs: alpha is supposed to be single component.");
    m_a_coeffs[amrlev][0].ParallelCopy(alpha, 0, 0, 1);   // LocalCopy when layouts agree
    m_needs_update = true; m_acoef_set = true;
}

void MLABecLaplacian::setACoeffs (int amrlev, Real alpha) { m_a_coeffs[amrlev][0].setVal(alpha); m_needs_update = true; m_acoef_set = true; }

void MLABecLaplacian::setBCoeffs (int amrlev, Array<MultiFab const*, 3> const& beta)
{
    for (int d = 0; d < 3; ++d) { m_b_coeffs[amrlev][0][d].ParallelCopy(*beta[d], 0, 0, 1); }
    m_needs_update = true;
}

void MLABecLaplacian::setBCoeffs (int amrlev, Real beta)
{
    for (int d = 0; d < 3; ++d) { m_b_coeffs[amrlev][0][d].setVal(beta); }
    m_needs_update = true;
}

void MLABecLaplacian::averageDownCoeffs ()
{
    for (int amrlev = H.num_amr_levels - 1; amrlev >= 0; --amrlev) {
        auto& a = m_a_coeffs[amrlev]; auto& b = m_b_coeffs[amrlev];
        for (int mglev = 1; mglev < int(a.size()); ++mglev) {   // averageDownCoeffsSameAmrLevel, AMReX_MLABecLaplacian.H:631-655
            const int ratio = (amrlev > 0) ? 2 : H.mg_coarsen_ratio_vec[mglev - 1][0];
            if (m_a_scalar == 0.0) { a[mglev].setVal(0.0); }
            else { average_down(a[mglev - 1], a[mglev], 0, 1, ratio); }
            for (int d = 0; d < 3; ++d) { average_down_faces(b[mglev - 1][d], b[mglev][d], d, ratio); }
        }
        if (amrlev > 0) {   // averageDownCoeffsToCoarseAmrLevel, :693-712
            if (m_a_scalar != 0.0) { average_down(m_a_coeffs[amrlev].back(), m_a_coeffs[amrlev - 1].front(), 0, 1, 2); }
            for (int d = 0; d < 3; ++d) { average_down_faces(m_b_coeffs[amrlev].back()[d], m_b_coeffs[amrlev - 1].front()[d], d, 2); }
        }
    }
}

void MLABecLaplacian::update_singular_flags ()
{
    m_is_singular.assign(H.num_amr_levels, 0);
    bool no_dirichlet = true;
    for (int d = 0; d < 3; ++d) { if (m_lobc[d] == BCType::Dirichlet || m_hibc[d] == BCType::Dirichlet) { no_dirichlet = false; } }
    if (no_dirichlet) {
        for (int alev = 0; alev < H.num_amr_levels; ++alev) {
            if (H.domain_covered[alev]) {
                if (m_a_scalar == 0.0) { m_is_singular[alev] = 1; }
                else {
                    const Real asum = m_a_coeffs[alev].back().sum();
                    const Real amax = m_a_coeffs[alev].back().norminf();
                    m_is_singular[alev] = (std::abs(asum) <= amax * Real(1.e-12));
                }
            }
        }
    }
}

// MLABecLaplacianT::applyRobinBCTermsCoeffs (AMReX_MLABecLaplacian.H:459-609): a Robin face acts as a homogeneous Neumann face
// with a larger diagonal: acoef(cell inside) += (b_scalar / a_scalar) * dxinv^2 * bcoef(face) * (1 - B)
void MLABecLaplacian::applyRobinBCTermsCoeffs ()
{
    if (!hasRobinBC()) { return; }
    bool reset_alpha = false;
    if (m_a_scalar == Real(0.0)) { m_a_scalar = Real(1.0); reset_alpha = true; }
    const Real bovera = m_b_scalar / m_a_scalar;
    if (!reset_alpha) {
        AMREX_ALWAYS_ASSERT_WITH_MESSAGE(m_scalars_set && m_acoef_set,
                                         "To reuse solver With Robin BC, one must re-call setScalars (and setACoeffs if the scalar is not zero)");
    }
    m_scalars_set = false; m_acoef_set = false;
    for (int amrlev = 0; amrlev < H.num_amr_levels; ++amrlev) {
        if (reset_alpha) { m_a_coeffs[amrlev][0].setVal(0.0); }
        LevelData const& L = lev(amrlev, 0);
        const int nf = int(L.bcfaces_h.size());
        if (nf == 0) { continue; }
        AMREX_ALWAYS_ASSERT_WITH_MESSAGE(int(m_robin.size()) > amrlev && m_robin[amrlev][0], "Robin BC: setLevelBC must supply robinbc_a / _b / _f");
        const Real* dxi = H.geom[amrlev][0].InvCellSize();
        MultiFab& ac = m_a_coeffs[amrlev][0];
        const b200mg_fab* out3[3] = {ac.d_fabs(), ac.d_fabs(), ac.d_fabs()};
        const b200mg_fab* b3[3] = {m_b_coeffs[amrlev][0][0].d_fabs(), m_b_coeffs[amrlev][0][1].d_fabs(), m_b_coeffs[amrlev][0][2].d_fabs()};
        const double fac[3] = {bovera * dxi[0] * dxi[0], bovera * dxi[1] * dxi[1], bovera * dxi[2] * dxi[2]};
        const double dx3[3] = {dxi[0], dxi[1], dxi[2]};
        for (int f = 0; f < 6; ++f) {
            const int d = f % 3;
            if ((f < 3 ? m_lobc_orig[d] : m_hibc_orig[d]) != BCType::Robin) { continue; }
            int only[6] = {0, 0, 0, 0, 0, 0}; only[f] = 1;
            B200_KCALL(b200mg_robin(nf, L.bcfaces.data(), L.layout->d_vbox(), out3, b3, nullptr, L.mask.d_table(), m_robin[amrlev][0]->d_table(),
                                    m_robin[amrlev][1]->d_table(), m_robin[amrlev][2]->d_table(), fac, dx3, only, 0, Gpu::gpuStream()));
        }
    }
}

void MLABecLaplacian::prepareForSolve ()
{
    MLLinOp::prepareForSolve();
    applyRobinBCTermsCoeffs();
    averageDownCoeffs();
    update_singular_flags();
    m_needs_update = false;
    buildMergedLeg();
}

void MLABecLaplacian::update ()
{
    applyRobinBCTermsCoeffs();
    averageDownCoeffs();
    update_singular_flags();
    m_needs_update = false;
    buildMergedLeg();
}

std::unique_ptr<MLLinOp> MLABecLaplacian::makeMergedOp (Geometry const& geom, BoxArray const& ba, DistributionMapping const& dm,
                                                        LPInfo const& inf, int mglev) const
{
    auto op = std::make_unique<MLABecLaplacian>();
    op->m_is_merged_copy = true;
    op->define({geom}, {ba}, {dm}, inf);
    copyOptionsTo(*op);
    op->setScalars(m_a_scalar, m_b_scalar);
    op->setACoeffs(0, m_a_coeffs[0][mglev]);                     // ParallelCopy onto the one box
    op->setBCoeffs(0, {{&m_b_coeffs[0][mglev][0], &m_b_coeffs[0][mglev][1], &m_b_coeffs[0][mglev][2]}});
    op->prepareForSolve();
    return op;
}

void MLABecLaplacian::normalize (int amrlev, int mglev, MultiFab& mf) const
{
    const Real* dxi = H.geom[amrlev][mglev].InvCellSize();
    auto const& T = mf.layout().tiles(0);
    B200_KCALL(b200mg_normalize_abec(T.n, T.d.data(), mf.layout().d_vbox(), mf.d_fabs(), m_a_coeffs[amrlev][mglev].d_fabs(),
                                     m_b_coeffs[amrlev][mglev][0].d_fabs(), m_b_coeffs[amrlev][mglev][1].d_fabs(), m_b_coeffs[amrlev][mglev][2].d_fabs(),
                                     m_a_scalar, m_b_scalar * dxi[0] * dxi[0], m_b_scalar * dxi[1] * dxi[1], m_b_scalar * dxi[2] * dxi[2], Gpu::gpuStream()));
}

bool MLABecLaplacian::Fapply (int amrlev, int mglev, MultiFab& out, MultiFab const& in, const MultiFab* rhs, Real* norm_dev) const
{
    const Real* dxi = H.geom[amrlev][mglev].InvCellSize();   // dh = beta*dxinv^2 (AMReX_MLABecLap_3D_K.H:18-20)
    auto const& T = out.layout().tiles(0);
    if (out.layout().pairable()) {
        B200_KCALL(b200mg_adotx_abec_pairs(T.n, T.d.data(), out.layout().d_vbox(), out.d_fabs(), in.d_fabs(), rhs ? rhs->d_fabs() : nullptr,
                                           m_a_coeffs[amrlev][mglev].d_fabs(), m_b_coeffs[amrlev][mglev][0].d_fabs(), m_b_coeffs[amrlev][mglev][1].d_fabs(),
                                           m_b_coeffs[amrlev][mglev][2].d_fabs(), m_a_scalar, m_b_scalar * dxi[0] * dxi[0], m_b_scalar * dxi[1] * dxi[1],
                                           m_b_scalar * dxi[2] * dxi[2], norm_dev, Gpu::gpuStream()));
        return norm_dev != nullptr;
    }
    B200_KCALL(b200mg_adotx_abec(T.n, T.d.data(), out.layout().d_vbox(), out.d_fabs(), in.d_fabs(), rhs ? rhs->d_fabs() : nullptr,
                                 m_a_coeffs[amrlev][mglev].d_fabs(), m_b_coeffs[amrlev][mglev][0].d_fabs(), m_b_coeffs[amrlev][mglev][1].d_fabs(),
                                 m_b_coeffs[amrlev][mglev][2].d_fabs(), m_a_scalar, m_b_scalar * dxi[0] * dxi[0], m_b_scalar * dxi[1] * dxi[1],
                                 m_b_scalar * dxi[2] * dxi[2], Gpu::gpuStream()));
    return false;
}

namespace { inline void gsrb_dh (Geometry const& g, Real b, Real dh[3]) { const Real* h = g.CellSize(); for (int d = 0; d < 3; ++d) { dh[d] = b / (h[d] * h[d]); } } }

void MLABecLaplacian::FresidualRestrict (int amrlev, int mglev, MultiFab& crse, MultiFab const& x, MultiFab const& b) const
{
    const Real* dxi = H.geom[amrlev][mglev].InvCellSize();
    auto const& T = x.layout().tiles(0);
    B200_KCALL(b200mg_residual_restrict_abec(T.n, T.d.data(), x.layout().d_vbox(), crse.d_fabs(), x.d_fabs(), b.d_fabs(),
                                             m_a_coeffs[amrlev][mglev].d_fabs(), m_b_coeffs[amrlev][mglev][0].d_fabs(),
                                             m_b_coeffs[amrlev][mglev][1].d_fabs(), m_b_coeffs[amrlev][mglev][2].d_fabs(), m_a_scalar,
                                             m_b_scalar * dxi[0] * dxi[0], m_b_scalar * dxi[1] * dxi[1], m_b_scalar * dxi[2] * dxi[2], Gpu::gpuStream()));
}

void MLABecLaplacian::Fsmooth (int amrlev, int mglev, MultiFab& sol, MultiFab const& rhs, int redblack) const
{
    LevelData const& L = lev(amrlev, mglev);
    Real dh[3]; gsrb_dh(H.geom[amrlev][mglev], m_b_scalar, dh);   // dh = beta/h^2 (AMReX_MLABecLaplacian.H:907-909)
    auto const& T = sol.layout().tiles(0);
    if (sol.layout().leanable()) {
        B200_KCALL(b200mg_gsrb_abec_pairs_lean(T.n, T.d.data(), L.layout->d_vbox(), sol.d_fabs(), rhs.d_fabs(), m_a_coeffs[amrlev][mglev].d_fabs(),
                                               m_b_coeffs[amrlev][mglev][0].d_fabs(), m_b_coeffs[amrlev][mglev][1].d_fabs(), m_b_coeffs[amrlev][mglev][2].d_fabs(),
                                               L.undrrelxr.d_table(), L.mask.d_table(), m_a_scalar, dh[0], dh[1], dh[2], redblack, Gpu::gpuStream()));
        return;
    }
    if (sol.layout().pairable()) {
        B200_KCALL(b200mg_gsrb_abec_pairs(T.n, T.d.data(), L.layout->d_vbox(), sol.d_fabs(), rhs.d_fabs(), m_a_coeffs[amrlev][mglev].d_fabs(),
                                          m_b_coeffs[amrlev][mglev][0].d_fabs(), m_b_coeffs[amrlev][mglev][1].d_fabs(), m_b_coeffs[amrlev][mglev][2].d_fabs(),
                                          L.undrrelxr.d_table(), L.mask.d_table(), m_a_scalar, dh[0], dh[1], dh[2], redblack, Gpu::gpuStream()));
        return;
    }
    B200_KCALL(b200mg_gsrb_abec(T.n, T.d.data(), L.layout->d_vbox(), sol.d_fabs(), rhs.d_fabs(), m_a_coeffs[amrlev][mglev].d_fabs(),
                                m_b_coeffs[amrlev][mglev][0].d_fabs(), m_b_coeffs[amrlev][mglev][1].d_fabs(), m_b_coeffs[amrlev][mglev][2].d_fabs(),
                                L.undrrelxr.d_table(), L.mask.d_table(), m_a_scalar, dh[0], dh[1], dh[2], redblack, Gpu::gpuStream()));
}

bool MLABecLaplacian::Fsmooth2 (int amrlev, int mglev, MultiFab& sol_out, MultiFab& sol_in, MultiFab const& rhs, bool zero_input,
                                const std::vector<int>* boxes) const
{
    LevelData const& L = lev(amrlev, mglev);
    Real dh[3]; gsrb_dh(H.geom[amrlev][mglev], m_b_scalar, dh);
    auto const& ac = m_a_coeffs[amrlev][mglev]; auto const& bc = m_b_coeffs[amrlev][mglev];
    int e;
    {
        Gpu::KernelScope ks__("b200mg_gsrb4(abec)");
        e = b200mg_gsrb4_subset(1, boxes ? int(boxes->size()) : L.layout->numLocal(), boxes ? boxes->data() : nullptr, L.h_vbox.data(),
                                &sol_in.desc(0), &sol_out.desc(0), &rhs.desc(0), &ac.desc(0),
                                &bc[0].desc(0), &bc[1].desc(0), &bc[2].desc(0), L.undrrelxr.h_table(), L.mask.h_table(),
                                m_a_scalar, dh[0], dh[1], dh[2], zero_input ? 1 : 0, Gpu::gpuStream());
    }
    if (e == 0) { return true; }
    if (e != int(cudaErrorInvalidValue)) { Gpu::check(e, "b200mg_gsrb4", __FILE__, __LINE__); }
    L.fused4_ok = 0;
    return false;
}

void MLABecLaplacian::Fjacobi (int amrlev, int mglev, MultiFab& sol_out, MultiFab const& sol_in, MultiFab const& rhs) const
{
    LevelData const& L = lev(amrlev, mglev);
    Real dh[3]; gsrb_dh(H.geom[amrlev][mglev], m_b_scalar, dh);
    const Real* dxi = H.geom[amrlev][mglev].InvCellSize();
    auto const& T = sol_in.layout().tiles(0);
    B200_KCALL(b200mg_jacobi_abec(T.n, T.d.data(), L.layout->d_vbox(), sol_out.d_fabs(), sol_in.d_fabs(), rhs.d_fabs(),
                                  m_a_coeffs[amrlev][mglev].d_fabs(), m_b_coeffs[amrlev][mglev][0].d_fabs(), m_b_coeffs[amrlev][mglev][1].d_fabs(),
                                  m_b_coeffs[amrlev][mglev][2].d_fabs(), L.undrrelxr.d_table(), L.mask.d_table(), m_a_scalar, dh[0], dh[1], dh[2],
                                  m_b_scalar * dxi[0] * dxi[0], m_b_scalar * dxi[1] * dxi[1], m_b_scalar * dxi[2] * dxi[2], Gpu::gpuStream()));
}

void MLABecLaplacian::FsmoothShell (int amrlev, int mglev, MultiFab& sol, MultiFab const& rhs, int redblack,
                                    const b200mg_facelink* links, bool push) const
{
    LevelData const& L = lev(amrlev, mglev);
    Real dh[3]; gsrb_dh(H.geom[amrlev][mglev], m_b_scalar, dh);
    B200_KCALL(b200mg_gsrb_shell_abec_linked(L.layout->numLocal(), L.layout->d_vbox(), sol.d_fabs(), rhs.d_fabs(), m_a_coeffs[amrlev][mglev].d_fabs(),
                                             m_b_coeffs[amrlev][mglev][0].d_fabs(), m_b_coeffs[amrlev][mglev][1].d_fabs(), m_b_coeffs[amrlev][mglev][2].d_fabs(),
                                             L.undrrelxr.d_table(), L.mask.d_table(), m_a_scalar, dh[0], dh[1], dh[2], redblack, maxFaceCells(L),
                                             links, push ? 1 : 0, Gpu::gpuStream()));
}

// ============================================================================================ MLPoisson
void MLPoisson::prepareForSolve ()
{
    MLLinOp::prepareForSolve();
    m_is_singular.assign(H.num_amr_levels, 0);
    bool no_dirichlet = true;
    for (int d = 0; d < 3; ++d) { if (m_lobc[d] == BCType::Dirichlet || m_hibc[d] == BCType::Dirichlet) { no_dirichlet = false; } }
    if (no_dirichlet) { for (int alev = 0; alev < H.num_amr_levels; ++alev) { if (H.domain_covered[alev]) { m_is_singular[alev] = 1; } } }
    buildMergedLeg();
}

std::unique_ptr<MLLinOp> MLPoisson::makeMergedOp (Geometry const& geom, BoxArray const& ba, DistributionMapping const& dm,
                                                  LPInfo const& inf, int) const
{
    auto op = std::make_unique<MLPoisson>();
    op->m_is_merged_copy = true;
    op->define({geom}, {ba}, {dm}, inf);
    copyOptionsTo(*op);
    op->prepareForSolve();
    return op;
}

bool MLPoisson::Fapply (int amrlev, int mglev, MultiFab& out, MultiFab const& in, const MultiFab* rhs, Real* norm_dev) const
{
    const Real* dxi = H.geom[amrlev][mglev].InvCellSize();
    auto const& T = out.layout().tiles(0);
    if (out.layout().pairable()) {
        B200_KCALL(b200mg_adotx_poisson_pairs(T.n, T.d.data(), out.layout().d_vbox(), out.d_fabs(), in.d_fabs(), rhs ? rhs->d_fabs() : nullptr,
                                              dxi[0] * dxi[0], dxi[1] * dxi[1], dxi[2] * dxi[2], norm_dev, Gpu::gpuStream()));
        return norm_dev != nullptr;
    }
    B200_KCALL(b200mg_adotx_poisson(T.n, T.d.data(), out.layout().d_vbox(), out.d_fabs(), in.d_fabs(), rhs ? rhs->d_fabs() : nullptr,
                                    dxi[0] * dxi[0], dxi[1] * dxi[1], dxi[2] * dxi[2], Gpu::gpuStream()));
    return false;
}

void MLPoisson::FresidualRestrict (int amrlev, int mglev, MultiFab& crse, MultiFab const& x, MultiFab const& b) const
{
    const Real* dxi = H.geom[amrlev][mglev].InvCellSize();
    auto const& T = x.layout().tiles(0);
    B200_KCALL(b200mg_residual_restrict_poisson(T.n, T.d.data(), x.layout().d_vbox(), crse.d_fabs(), x.d_fabs(), b.d_fabs(),
                                                dxi[0] * dxi[0], dxi[1] * dxi[1], dxi[2] * dxi[2], Gpu::gpuStream()));
}

void MLPoisson::Fsmooth (int amrlev, int mglev, MultiFab& sol, MultiFab const& rhs, int redblack) const
{
    LevelData const& L = lev(amrlev, mglev);
    const Real* dxi = H.geom[amrlev][mglev].InvCellSize();
    auto const& T = sol.layout().tiles(0);
    if (sol.layout().leanable()) {
        B200_KCALL(b200mg_gsrb_poisson_pairs_lean(T.n, T.d.data(), L.layout->d_vbox(), sol.d_fabs(), rhs.d_fabs(), L.undrrelxr.d_table(), L.mask.d_table(),
                                                  dxi[0] * dxi[0], dxi[1] * dxi[1], dxi[2] * dxi[2], redblack, Gpu::gpuStream()));
        return;
    }
    if (sol.layout().pairable()) {
        B200_KCALL(b200mg_gsrb_poisson_pairs(T.n, T.d.data(), L.layout->d_vbox(), sol.d_fabs(), rhs.d_fabs(), L.undrrelxr.d_table(), L.mask.d_table(),
                                             dxi[0] * dxi[0], dxi[1] * dxi[1], dxi[2] * dxi[2], redblack, Gpu::gpuStream()));
        return;
    }
    B200_KCALL(b200mg_gsrb_poisson(T.n, T.d.data(), L.layout->d_vbox(), sol.d_fabs(), rhs.d_fabs(), L.undrrelxr.d_table(), L.mask.d_table(),
                                   dxi[0] * dxi[0], dxi[1] * dxi[1], dxi[2] * dxi[2], redblack, Gpu::gpuStream()));
}

bool MLPoisson::Fsmooth2 (int amrlev, int mglev, MultiFab& sol_out, MultiFab& sol_in, MultiFab const& rhs, bool zero_input,
                          const std::vector<int>* boxes) const
{
    LevelData const& L = lev(amrlev, mglev);
    const Real* dxi = H.geom[amrlev][mglev].InvCellSize();
    int e;
    {
        Gpu::KernelScope ks__("b200mg_gsrb4(poisson)");
        e = b200mg_gsrb4_subset(0, boxes ? int(boxes->size()) : L.layout->numLocal(), boxes ? boxes->data() : nullptr, L.h_vbox.data(),
                                &sol_in.desc(0), &sol_out.desc(0), &rhs.desc(0), nullptr,
                                nullptr, nullptr, nullptr, L.undrrelxr.h_table(), L.mask.h_table(),
                                0.0, dxi[0] * dxi[0], dxi[1] * dxi[1], dxi[2] * dxi[2], zero_input ? 1 : 0, Gpu::gpuStream());
    }
    if (e == 0) { return true; }
    if (e != int(cudaErrorInvalidValue)) { Gpu::check(e, "b200mg_gsrb4", __FILE__, __LINE__); }
    L.fused4_ok = 0;
    return false;
}

void MLPoisson::Fjacobi (int amrlev, int mglev, MultiFab& sol_out, MultiFab const& sol_in, MultiFab const& rhs) const
{
    LevelData const& L = lev(amrlev, mglev);
    const Real* dxi = H.geom[amrlev][mglev].InvCellSize();
    auto const& T = sol_in.layout().tiles(0);
    B200_KCALL(b200mg_jacobi_poisson(T.n, T.d.data(), L.layout->d_vbox(), sol_out.d_fabs(), sol_in.d_fabs(), rhs.d_fabs(),
                                     L.undrrelxr.d_table(), L.mask.d_table(), dxi[0] * dxi[0], dxi[1] * dxi[1], dxi[2] * dxi[2], Gpu::gpuStream()));
}

void MLPoisson::FsmoothShell (int amrlev, int mglev, MultiFab& sol, MultiFab const& rhs, int redblack,
                              const b200mg_facelink* links, bool push) const
{
    LevelData const& L = lev(amrlev, mglev);
    const Real* dxi = H.geom[amrlev][mglev].InvCellSize();
    B200_KCALL(b200mg_gsrb_shell_poisson_linked(L.layout->numLocal(), L.layout->d_vbox(), sol.d_fabs(), rhs.d_fabs(), L.undrrelxr.d_table(), L.mask.d_table(),
                                                dxi[0] * dxi[0], dxi[1] * dxi[1], dxi[2] * dxi[2], redblack, maxFaceCells(L),
                                                links, push ? 1 : 0, Gpu::gpuStream()));
}

} // namespace amrex
