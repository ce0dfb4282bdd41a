// extern "C" boundary: see include/amrex_b200_fi.h for the contract and the reference citations.
#include "../mlmg/AMReX_MLMG.H"
#include "../mlmg/AMReX_GMRESMLMG.H"
#include "../base/AMReX_PlotFileUtil.H"
#include "amrex_b200_fi.h"

#include <cstring>
#include <iostream>
#include <sstream>
#include <limits>
#include <string>

using namespace amrex;

namespace amrex { void clear_comm_caches (); }

namespace {
void Print0 (std::string const& s) { if (ParallelDescriptor::IOProcessor()) { std::cout << s << std::flush; } }
std::string g_err;
bool g_has_err = false;
void set_err (const char* where, const char* what) { g_err = std::string(where) + ": " + what; g_has_err = true; }
constexpr double kNaN = std::numeric_limits<double>::quiet_NaN();
}

#define FI_TRY try {
#define FI_CATCH(ret) } catch (std::exception const& e) { set_err(__func__, e.what()); ret; } catch (...) { set_err(__func__, "unknown exception"); ret; }
#define FI_VOID(body) FI_TRY body FI_CATCH(return)

extern "C" {

// ------------------------------------------------------------------------------------------ runtime
int amrex_b200_init (int device_id) { FI_TRY Gpu::Initialize(device_id); return 0; FI_CATCH(return 1) }
void amrex_b200_finalize (void)
{
    FI_VOID( clear_comm_caches(); LevelLayout::clearCache(); ParallelDescriptor::FinalizeComm(); Gpu::Finalize(); )
}
int amrex_b200_initialized (void) { return Gpu::Initialized() ? 1 : 0; }
int amrex_b200_nccl_unique_id_bytes (void) { return ParallelDescriptor::NcclUniqueIdBytes(); }
int amrex_b200_nccl_get_unique_id (void* out) { FI_TRY ParallelDescriptor::NcclGetUniqueId(out); return 0; FI_CATCH(return 1) }
int amrex_b200_comm_init (int rank, int nranks, const void* uid) { FI_TRY ParallelDescriptor::InitComm(rank, nranks, uid); return 0; FI_CATCH(return 1) }
void amrex_b200_comm_finalize (void) { FI_VOID( ParallelDescriptor::FinalizeComm(); ) }
int amrex_b200_myproc (void) { return ParallelDescriptor::MyProc(); }
int amrex_b200_nprocs (void) { return ParallelDescriptor::NProcs(); }
const char* amrex_b200_last_error (void) { return g_has_err ? g_err.c_str() : nullptr; }
void amrex_b200_clear_error (void) { g_has_err = false; g_err.clear(); }
void amrex_b200_synchronize (void) { FI_VOID( Gpu::streamSynchronize(); ) }
long long amrex_b200_launch_count (void) { return Gpu::launchCount(); }
void amrex_b200_reset_launch_count (void) { Gpu::resetLaunchCount(); }
void* amrex_b200_stream (void) { return Gpu::gpuStream(); }
void amrex_b200_profile_enable (int on) { Gpu::profileEnable(on != 0); }
int amrex_b200_profile_report (char* buf, int capacity)
{
    static std::string pending;
    FI_TRY
        if (pending.empty()) { pending = Gpu::profileReport(); }
        const int n = int(pending.size());
        if (buf == nullptr) { return n; }
        const int m = std::min(n, capacity - 1);
        std::memcpy(buf, pending.data(), m); buf[m] = 0;
        pending.clear();
        return n;
    FI_CATCH(return -1)
}

// ----------------------------------------------------------------------------------------- Geometry
void amrex_b200_geometry_setup (const Real problo[3], const Real probhi[3], const int is_periodic[3])
{
    RealBox rb({problo[0], problo[1], problo[2]}, {probhi[0], probhi[1], probhi[2]});
    Geometry::Setup(&rb, 0, is_periodic);
}
void amrex_fi_new_geometry (Geometry** geom, int lo[3], int hi[3])
{
    FI_VOID( *geom = new Geometry(Box(IntVect(lo[0], lo[1], lo[2]), IntVect(hi[0], hi[1], hi[2]))); )
}
void amrex_fi_delete_geometry (Geometry* geom) { delete geom; }
void amrex_fi_geometry_get_intdomain (const Geometry* geom, int lo[3], int hi[3])
{
    for (int d = 0; d < 3; ++d) { lo[d] = geom->Domain().smallEnd(d); hi[d] = geom->Domain().bigEnd(d); }
}

// ----------------------------------------------------------------------------------------- BoxArray
void amrex_fi_new_boxarray (BoxArray** ba, int lo[3], int hi[3])
{
    FI_VOID( *ba = new BoxArray(Box(IntVect(lo[0], lo[1], lo[2]), IntVect(hi[0], hi[1], hi[2]))); )
}
void amrex_fi_new_boxarray_from_bxfarr (BoxArray** ba, const int* bxs, const int nsides, const int ndims, const int nbxs)
{
    FI_VOID(
        AMREX_ALWAYS_ASSERT(nsides == 2 && ndims >= 3);
        BoxList bl;
        for (int i = 0; i < nbxs; ++i) {
            bl.push_back(Box(IntVect(bxs[0], bxs[2], bxs[4]), IntVect(bxs[1], bxs[3], bxs[5])));
            bxs += 2 * ndims;
        }
        *ba = new BoxArray(bl); )
}
void amrex_fi_delete_boxarray (BoxArray* ba) { delete ba; }
void amrex_fi_clone_boxarray (BoxArray** bao, const BoxArray* bai) { FI_VOID( *bao = new BoxArray(*bai); ) }
void amrex_fi_boxarray_maxsize (BoxArray* ba, int sz[]) { FI_VOID( ba->maxSize(IntVect(sz[0], sz[1], sz[2])); ) }
long long amrex_fi_boxarray_nboxes (const BoxArray* ba) { return ba->size(); }
void amrex_fi_boxarray_get_box (const BoxArray* ba, int i, int lo[3], int hi[3])
{
    Box const& b = (*ba)[i];
    for (int d = 0; d < 3; ++d) { lo[d] = b.smallEnd(d); hi[d] = b.bigEnd(d); }
}
void amrex_fi_boxarray_nodal_type (const BoxArray* ba, int inodal[3]) { for (int d = 0; d < 3; ++d) { inodal[d] = ba->ixType()[d]; } }
long long amrex_fi_boxarray_numpts (const BoxArray* ba) { return ba->numPts(); }
int amrex_fi_boxarray_issame (const BoxArray* a, const BoxArray* b) { return *a == *b; }
void amrex_b200_boxarray_coarsen (BoxArray* ba, int ratio) { FI_VOID( ba->coarsen(ratio); ) }
void amrex_b200_boxarray_refine (BoxArray* ba, int ratio) { FI_VOID( ba->refine(ratio); ) }
void amrex_b200_boxarray_convert (BoxArray* ba, const int nodal[3]) { FI_VOID( ba->convert(IntVect(nodal[0], nodal[1], nodal[2])); ) }

// ------------------------------------------------------------------------------ DistributionMapping
void amrex_fi_new_distromap (DistributionMapping** dm, const BoxArray* ba) { FI_VOID( *dm = new DistributionMapping(*ba); ) }
void amrex_fi_new_distromap_from_pmap (DistributionMapping** dm, const int* pmap, const int plen)
{
    FI_VOID( *dm = new DistributionMapping(Vector<int>(pmap, pmap + plen)); )
}
void amrex_fi_delete_distromap (DistributionMapping* dm) { delete dm; }
void amrex_fi_distromap_get_pmap (const DistributionMapping* dm, int* pmap, const int plen)
{
    auto const& p = dm->ProcessorMap();
    for (int i = 0; i < plen && i < int(p.size()); ++i) { pmap[i] = p[i]; }
}
void amrex_b200_new_distromap_sfc (DistributionMapping** dm, const BoxArray* ba, int nprocs) { FI_VOID( *dm = new DistributionMapping(*ba, nprocs); ) }

void amrex_b200_make_sfc (const BoxArray* ba, int nprocs, int* bucket_of_box)
{
    FI_VOID(
        auto buckets = DistributionMapping::makeSFC(*ba, true, nprocs);
        for (int r = 0; r < nprocs; ++r) { for (int b : buckets[r]) { bucket_of_box[b] = r; } } )
}

// ----------------------------------------------------------------------------------------- MultiFab
void amrex_fi_new_multifab (MultiFab** mf, const BoxArray** ba, const DistributionMapping** dm, int nc, const int* ng, const int* nodal)
{
    FI_VOID(
        AMREX_ALWAYS_ASSERT_WITH_MESSAGE(ng[0] == ng[1] && ng[1] == ng[2], "uniform ghost width required");
        *mf = new MultiFab(amrex::convert(**ba, IntVect(nodal[0], nodal[1], nodal[2])), **dm, nc, ng[0]);
        *ba = &((*mf)->boxArray()); *dm = &((*mf)->DistributionMap()); )
}
void amrex_fi_delete_multifab (MultiFab* mf) { FI_VOID( delete mf; ) }
int amrex_fi_multifab_ncomp (const MultiFab* mf) { return mf->nComp(); }
void amrex_fi_multifab_ngrow (const MultiFab* mf, int* ngv) { for (int d = 0; d < 3; ++d) { ngv[d] = mf->nGrow(); } }
const BoxArray* amrex_fi_multifab_boxarray (const MultiFab* mf) { return &mf->boxArray(); }
const DistributionMapping* amrex_fi_multifab_distromap (const MultiFab* mf) { return &mf->DistributionMap(); }
void amrex_fi_multifab_dataptr_int (MultiFab* mf, int igrd, Real** dp, int lo[3], int hi[3])
{
    FI_VOID(
        const int li = mf->layout().localIndex(igrd);
        AMREX_ALWAYS_ASSERT_WITH_MESSAGE(li >= 0, "grid is not local to this rank");
        auto const& d = mf->desc(li);
        *dp = d.p; for (int a = 0; a < 3; ++a) { lo[a] = d.lo[a]; hi[a] = d.hi[a]; } )
}
void amrex_b200_multifab_strides (const MultiFab* mf, int igrd, long long strides[3])
{
    FI_VOID( const int li = mf->layout().localIndex(igrd); AMREX_ALWAYS_ASSERT(li >= 0);
             auto const& d = mf->desc(li); strides[0] = d.jstride; strides[1] = d.kstride; strides[2] = d.nstride; )
}
Real amrex_fi_multifab_sum (const MultiFab* mf, int comp) { FI_TRY AMREX_ALWAYS_ASSERT(comp == 0); return mf->sum(); FI_CATCH(return kNaN) }
Real amrex_fi_multifab_norm0 (const MultiFab* mf, int comp) { FI_TRY AMREX_ALWAYS_ASSERT(comp == 0); return mf->norminf(); FI_CATCH(return kNaN) }
void amrex_fi_multifab_setval (MultiFab* mf, Real val, int ic, int nc, const int* ng) { FI_VOID( AMREX_ALWAYS_ASSERT(ic == 0 && nc == 1); mf->setVal(val, ng[0]); ) }
void amrex_fi_multifab_plus (MultiFab* mf, Real val, int ic, int nc, int ng) { FI_VOID( AMREX_ALWAYS_ASSERT(ic == 0 && nc == 1); mf->plus(val, ng); ) }
void amrex_fi_multifab_mult (MultiFab* mf, Real val, int ic, int nc, int ng) { FI_VOID( AMREX_ALWAYS_ASSERT(ic == 0 && nc == 1); mf->mult(val, ng); ) }
void amrex_fi_multifab_add (MultiFab* d, const MultiFab* s, int sc, int dc, int nc, const int* ng) { FI_VOID( MultiFab::Add(*d, *s, sc, dc, nc, ng[0]); ) }
void amrex_fi_multifab_subtract (MultiFab* d, const MultiFab* s, int sc, int dc, int nc, const int* ng) { FI_VOID( MultiFab::Subtract(*d, *s, sc, dc, nc, ng[0]); ) }
void amrex_fi_multifab_saxpy (MultiFab* d, Real a, const MultiFab* s, int sc, int dc, int nc, const int* ng) { FI_VOID( MultiFab::Saxpy(*d, a, *s, sc, dc, nc, ng[0]); ) }
void amrex_fi_multifab_copy (MultiFab* d, const MultiFab* s, int sc, int dc, int nc, const int* ng) { FI_VOID( MultiFab::Copy(*d, *s, sc, dc, nc, ng[0]); ) }
void amrex_fi_multifab_parallelcopy (MultiFab* d, const MultiFab* s, int sc, int dc, int nc, int srcng, int dstng, const Geometry* geom)
{
    FI_VOID( d->ParallelCopy(*s, sc, dc, nc, srcng, dstng, geom->periodicity()); )
}
void amrex_fi_multifab_fill_boundary (MultiFab* mf, const Geometry* geom, int c, int nc, int cross)
{
    FI_VOID( mf->FillBoundary(c, nc, geom->periodicity(), cross != 0); )
}
// ---- the rest of the reference's MultiFab / iMultiFab / MFIter / Geometry entries (Src/F_Interfaces/Base/AMReX_multifab_fi.cpp,
//      AMReX_geometry_fi.cpp, AMReX_distromap_fi.cpp, AMReX_boxarray_fi.cpp): same names, same argument order.  Data pointers
//      are DEVICE pointers (the fields live in HBM); the iteration helpers are host-side bookkeeping over the local boxes.
namespace { inline int uniform_ng (const int* ng, const char* who) { if (!(ng[0] == ng[1] && ng[1] == ng[2])) { throw std::runtime_error(std::string(who) + ": uniform ghost width required"); } return ng[0]; } }
Real amrex_fi_multifab_min (const MultiFab* mf, int comp, int nghost) { FI_TRY return mf->min(comp, nghost); FI_CATCH(return kNaN) }
Real amrex_fi_multifab_max (const MultiFab* mf, int comp, int nghost) { FI_TRY return mf->max(comp, nghost); FI_CATCH(return kNaN) }
Real amrex_fi_multifab_norm1 (const MultiFab* mf, int comp) { FI_TRY return mf->norm1(comp); FI_CATCH(return kNaN) }
Real amrex_fi_multifab_norm2 (const MultiFab* mf, int comp) { FI_TRY return mf->norm2(comp); FI_CATCH(return kNaN) }
void amrex_fi_multifab_multiply (MultiFab* d, const MultiFab* s, int sc, int dc, int nc, const int* ng) { FI_VOID( MultiFab::Multiply(*d, *s, sc, dc, nc, uniform_ng(ng, "multiply")); ) }
void amrex_fi_multifab_divide (MultiFab* d, const MultiFab* s, int sc, int dc, int nc, const int* ng) { FI_VOID( MultiFab::Divide(*d, *s, sc, dc, nc, uniform_ng(ng, "divide")); ) }
void amrex_fi_multifab_lincomb (MultiFab* d, Real a, const MultiFab* s1, int sc1, Real b, const MultiFab* s2, int sc2, int dc, int nc, const int* ng)
{
    // dst = a*src1 + b*src2 (MultiFab::LinComb, AMReX_MultiFab.cpp): copy src2, then dst = a*src1 + b*dst
    FI_VOID( const int g = uniform_ng(ng, "lincomb");
             if (d != s2) { MultiFab::Copy(*d, *s2, sc2, dc, nc, g); }
             AMREX_ALWAYS_ASSERT_WITH_MESSAGE(d != s1 || d == s2, "lincomb: dst may alias src2 only");
             MultiFab::LinComb(*d, a, *s1, b, sc1, dc, nc, g); )
}
void amrex_fi_multifab_parallelcopy_gv (MultiFab* d, const MultiFab* s, int sc, int dc, int nc, const int* srcng, const int* dstng, const Geometry* geom)
{
    FI_VOID( d->ParallelCopy(*s, sc, dc, nc, uniform_ng(srcng, "parallelcopy_gv"), uniform_ng(dstng, "parallelcopy_gv"), geom->periodicity()); )
}
void amrex_fi_multifab_sum_boundary (MultiFab* mf, const Geometry* geom, int icomp, int ncomp) { FI_VOID( mf->SumBoundary(icomp, ncomp, geom->periodicity()); ) }
// node-centred synchronisation (OwnerMask / OverrideSync / AverageSync) belongs to the nodal solvers: outside the cell-centred path
void amrex_fi_build_owner_imultifab (iMultiFab**, const BoxArray**, const DistributionMapping**, const MultiFab*, const Geometry*)
{ set_err(__func__, "nodal owner masks are outside the cell-centred MLMG path of this library"); }
void amrex_fi_multifab_override_sync (MultiFab*, const Geometry*) { set_err(__func__, "nodal OverrideSync is outside the cell-centred MLMG path of this library"); }
void amrex_fi_multifab_override_sync_mask (MultiFab*, const Geometry*, const iMultiFab*) { set_err(__func__, "nodal OverrideSync is outside the cell-centred MLMG path of this library"); }
void amrex_fi_multifab_average_sync (MultiFab*, const Geometry*) { set_err(__func__, "nodal AverageSync is outside the cell-centred MLMG path of this library"); }
void amrex_fi_new_multifab_alias (MultiFab**, const MultiFab*, int, int) { set_err(__func__, "aliased MultiFabs are not supported: copy the components (amrex_fi_multifab_copy)"); }

// iMultiFab
void amrex_fi_new_imultifab (iMultiFab** imf, const BoxArray** ba, const DistributionMapping** dm, int nc, const int* ng, const int* nodal)
{
    FI_VOID( *imf = new iMultiFab(amrex::convert(**ba, IntVect(nodal[0], nodal[1], nodal[2])), **dm, nc, uniform_ng(ng, "new_imultifab"));
             *ba = &((*imf)->boxArray()); *dm = &((*imf)->DistributionMap()); )
}
void amrex_fi_new_imultifab_alias (iMultiFab**, const iMultiFab*, int, int) { set_err(__func__, "aliased iMultiFabs are not supported"); }
void amrex_fi_delete_imultifab (iMultiFab* imf) { delete imf; }
void amrex_fi_imultifab_setval (iMultiFab* imf, int val, int ic, int nc, const int* ng)
{
    FI_VOID( const int g = uniform_ng(ng, "imultifab_setval");
             AMREX_ALWAYS_ASSERT(g <= imf->nGrow() && ic >= 0 && ic + nc <= imf->nComp());
             for (int li = 0; li < imf->local_size(); ++li) {      // device memset per row range: the whole grown fab when g == nGrow
                 auto const& d = imf->desc(li);
                 const Box b = amrex::grow(imf->layout().box(li), g);
                 std::vector<int> row(std::size_t(b.length(0)), val);
                 for (int n = ic; n < ic + nc; ++n) for (int k = b.smallEnd(2); k <= b.bigEnd(2); ++k) for (int j = b.smallEnd(1); j <= b.bigEnd(1); ++j) {
                     int* p = d.p + (b.smallEnd(0) - d.lo[0]) + (j - d.lo[1]) * d.jstride + (k - d.lo[2]) * d.kstride + n * d.nstride;
                     Gpu::htod_memcpy_async(p, row.data(), row.size() * sizeof(int));
                 }
                 Gpu::streamSynchronize();
             } )
}

// MFIter: iteration over the local boxes (no tiling on the device: a tile is the whole box)
int amrex_fi_mfiter_allow_multiple (int allow) { return allow; }
void amrex_fi_new_mfiter_r (MFIter** mfi, MultiFab* mf, int, int) { FI_VOID( *mfi = new MFIter(*mf); ) }
void amrex_fi_new_mfiter_i (MFIter** mfi, iMultiFab* imf, int, int) { FI_VOID( *mfi = new MFIter(*imf); ) }
void amrex_fi_new_mfiter_rs (MFIter** mfi, MultiFab* mf, const int*, int) { FI_VOID( *mfi = new MFIter(*mf); ) }
void amrex_fi_new_mfiter_is (MFIter** mfi, iMultiFab* imf, const int*, int) { FI_VOID( *mfi = new MFIter(*imf); ) }
void amrex_fi_new_mfiter_badm (MFIter** mfi, BoxArray* ba, DistributionMapping* dm, int, int) { FI_VOID( *mfi = new MFIter(*ba, *dm); ) }
void amrex_fi_new_mfiter_badm_s (MFIter** mfi, BoxArray* ba, DistributionMapping* dm, const int*, int) { FI_VOID( *mfi = new MFIter(*ba, *dm); ) }
void amrex_fi_delete_mfiter (MFIter* mfi) { delete mfi; }
void amrex_fi_increment_mfiter (MFIter* mfi, int* isvalid) { ++(*mfi); *isvalid = mfi->isValid() ? 1 : 0; }
void amrex_fi_mfiter_is_valid (MFIter* mfi, int* isvalid) { *isvalid = mfi->isValid() ? 1 : 0; }
int amrex_fi_mfiter_grid_index (MFIter* mfi) { return mfi->index(); }
int amrex_fi_mfiter_local_tile_index (MFIter*) { return 0; }
namespace {
inline void box_out (Box const& bx, int lo[3], int hi[3], int* nodal)
{
    for (int d = 0; d < 3; ++d) { lo[d] = bx.smallEnd(d); hi[d] = bx.bigEnd(d); if (nodal) { nodal[d] = bx.ixType().test(d) ? 1 : 0; } }
}
}
void amrex_fi_mfiter_tilebox (MFIter* mfi, int lo[3], int hi[3], int nodal[3]) { box_out(mfi->tilebox(), lo, hi, nodal); }
void amrex_fi_mfiter_tilebox_iv (MFIter* mfi, int lo[3], int hi[3], const int nodal[3])
{
    box_out(amrex::convert(amrex::enclosedCells(mfi->tilebox()), IntVect(nodal[0], nodal[1], nodal[2])), lo, hi, nullptr);
}
void amrex_fi_mfiter_nodaltilebox (MFIter* mfi, int dir, int lo[3], int hi[3], int nodal[3])
{
    box_out(amrex::convert(amrex::enclosedCells(mfi->tilebox()), IntVect::TheDimensionVector(dir)), lo, hi, nodal);
}
void amrex_fi_mfiter_growntilebox (MFIter* mfi, int lo[3], int hi[3], int ng, int nodal[3]) { box_out(mfi->growntilebox(ng), lo, hi, nodal); }
void amrex_fi_mfiter_grownnodaltilebox (MFIter* mfi, int lo[3], int hi[3], int dir, int ng, int nodal[3])
{
    box_out(amrex::grow(amrex::convert(amrex::enclosedCells(mfi->tilebox()), IntVect::TheDimensionVector(dir)), ng), lo, hi, nodal);
}
void amrex_fi_mfiter_validbox (MFIter* mfi, int lo[3], int hi[3], int nodal[3]) { box_out(mfi->validbox(), lo, hi, nodal); }
void amrex_fi_mfiter_fabbox (MFIter* mfi, int lo[3], int hi[3], int nodal[3]) { box_out(mfi->fabbox(), lo, hi, nodal); }
void amrex_fi_multifab_dataptr_iter (MultiFab* mf, MFIter* mfi, Real** dp, int lo[3], int hi[3])
{
    FI_VOID( auto const& d = mf->desc(mfi->LocalIndex()); *dp = d.p; for (int a = 0; a < 3; ++a) { lo[a] = d.lo[a]; hi[a] = d.hi[a]; } )
}
void amrex_fi_imultifab_dataptr (iMultiFab* imf, MFIter* mfi, int** dp, int lo[3], int hi[3])
{
    FI_VOID( auto const& d = imf->desc(mfi->LocalIndex()); *dp = d.p; for (int a = 0; a < 3; ++a) { lo[a] = d.lo[a]; hi[a] = d.hi[a]; } )
}

// Geometry / DistributionMapping / BoxArray leftovers
void amrex_fi_geometry_get_pmask (const Geometry* geom, int is_per[3]) { for (int d = 0; d < 3; ++d) { is_per[d] = geom->isPeriodic(d) ? 1 : 0; } }
void amrex_fi_geometry_get_probdomain (const Geometry* geom, Real problo[3], Real probhi[3])
{
    for (int d = 0; d < 3; ++d) { problo[d] = geom->ProbLo()[d]; probhi[d] = geom->ProbHi()[d]; }
}
void amrex_fi_clone_distromap (DistributionMapping** dmo, const DistributionMapping* dmi) { FI_VOID( *dmo = new DistributionMapping(*dmi); ) }
int amrex_fi_distromap_issame (const DistributionMapping* a, const DistributionMapping* b) { return (*a == *b) ? 1 : 0; }
void amrex_fi_print_distromap (const DistributionMapping* dm)
{
    std::string o = "(DistributionMapping"; for (int p : dm->ProcessorMap()) { o += " " + std::to_string(p); } Print0(o + ")\n");
}
int amrex_fi_boxarray_intersects_box (const BoxArray* ba, const int lo[3], const int hi[3])
{
    FI_TRY std::vector<std::pair<int, Box>> is; ba->intersections(Box(IntVect(lo[0], lo[1], lo[2]), IntVect(hi[0], hi[1], hi[2]), ba->ixType()), is);
           return is.empty() ? 0 : 1; FI_CATCH(return 0)
}
void amrex_fi_print_boxarray (const BoxArray* ba)
{
    std::string o = "(BoxArray maxbox(" + std::to_string(ba->size()) + ")\n";
    for (int i = 0, N = int(ba->size()); i < N; ++i) {
        Box const& b = (*ba)[i];
        o += "  ((" + std::to_string(b.smallEnd(0)) + "," + std::to_string(b.smallEnd(1)) + "," + std::to_string(b.smallEnd(2)) + ") ("
           + std::to_string(b.bigEnd(0)) + "," + std::to_string(b.bigEnd(1)) + "," + std::to_string(b.bigEnd(2)) + "))\n";
    }
    Print0(o + ")\n");
}
void amrex_fi_print_box (const int lo[3], const int hi[3], const int nodal[3])
{
    std::ostringstream o;
    o << "((" << lo[0] << "," << lo[1] << "," << lo[2] << ") (" << hi[0] << "," << hi[1] << "," << hi[2] << ") (" << nodal[0] << "," << nodal[1] << "," << nodal[2] << "))\n";
    Print0(o.str());
}
// multi-level averaging (Src/F_Interfaces/Base/AMReX_multifabutil_fi.cpp)
void amrex_fi_average_down (const MultiFab* S_fine, MultiFab* S_crse, const Geometry*, const Geometry*, int scomp, int ncomp, int rr)
{
    FI_VOID( amrex::average_down(*S_fine, *S_crse, scomp, ncomp, rr); )
}
void amrex_fi_average_down_faces (MultiFab const* fmf[], MultiFab* cmf[], const Geometry*, int scomp, int ncomp, int rr)
{
    FI_VOID( AMREX_ALWAYS_ASSERT_WITH_MESSAGE(scomp == 0 && ncomp == 1, "average_down_faces: one component");
             for (int d = 0; d < 3; ++d) { amrex::average_down_faces(*fmf[d], *cmf[d], d, rr); } )
}
void amrex_fi_average_cellcenter_to_face (MultiFab* fc[], const MultiFab* cc, const Geometry* geom)
{
    FI_VOID( amrex::average_cellcenter_to_face({{fc[0], fc[1], fc[2]}}, *cc, *geom); )
}
void amrex_fi_average_down_cell_node (const MultiFab* S_fine, MultiFab* S_crse, int scomp, int ncomp, int rr)
{
    FI_VOID( AMREX_ALWAYS_ASSERT_WITH_MESSAGE(S_fine->ixType().cellCentered(), "average_down_cell_node: cell-centred data only (nodal averaging is outside this library's path)");
             amrex::average_down(*S_fine, *S_crse, scomp, ncomp, rr); )
}

Real amrex_b200_multifab_dot (const MultiFab* x, const MultiFab* y) { FI_TRY return MultiFab::Dot(*x, *y); FI_CATCH(return kNaN) }
void amrex_b200_multifab_upload (MultiFab* mf, const Real* h, const int lo[3], const int hi[3], int comp, int ng)
{
    FI_VOID( mf->copyFromHost(h, Box(IntVect(lo[0], lo[1], lo[2]), IntVect(hi[0], hi[1], hi[2]), mf->ixType()), comp, ng); )
}
void amrex_b200_multifab_download (const MultiFab* mf, Real* h, const int lo[3], const int hi[3], int comp, int ng)
{
    FI_VOID( mf->copyToHost(h, Box(IntVect(lo[0], lo[1], lo[2]), IntVect(hi[0], hi[1], hi[2]), mf->ixType()), comp, ng); )
}
// local grid igrd (global box index) alone: its valid cells and ng ghost layers, Fortran order over the grown box
void amrex_b200_multifab_download_fab (const MultiFab* mf, int igrd, Real* h, int comp, int ng)
{
    FI_VOID( const int li = mf->layout().localIndex(igrd);
             AMREX_ALWAYS_ASSERT_WITH_MESSAGE(li >= 0, "amrex_b200_multifab_download_fab: the grid is not local");
             mf->copyFabToHost(li, h, comp, ng); )
}
// the same transfers enqueued on a stream of the caller's (pinned host memory), without synchronisation: a streaming
// application uploads the inputs of the next solve and downloads the previous solution while the current solve runs
void amrex_b200_multifab_upload_async (MultiFab* mf, const Real* h, const int lo[3], const int hi[3], int comp, int ng, void* stream)
{
    FI_TRY Gpu::setStream(static_cast<cudaStream_t>(stream));
           mf->copyFromHost(h, Box(IntVect(lo[0], lo[1], lo[2]), IntVect(hi[0], hi[1], hi[2]), mf->ixType()), comp, ng, false);
           Gpu::setStream(nullptr);
    FI_CATCH(Gpu::setStream(nullptr); return)
}
void amrex_b200_multifab_download_async (const MultiFab* mf, Real* h, const int lo[3], const int hi[3], int comp, int ng, void* stream)
{
    FI_TRY Gpu::setStream(static_cast<cudaStream_t>(stream));
           mf->copyToHost(h, Box(IntVect(lo[0], lo[1], lo[2]), IntVect(hi[0], hi[1], hi[2]), mf->ixType()), comp, ng, true, false);
           Gpu::setStream(nullptr);
    FI_CATCH(Gpu::setStream(nullptr); return)
}
void amrex_b200_average_cellcenter_to_face (MultiFab* fx, MultiFab* fy, MultiFab* fz, const MultiFab* cc, const Geometry* geom)
{
    FI_VOID( average_cellcenter_to_face({fx, fy, fz}, *cc, *geom); )
}

// --------------------------------------------------------------------------------- plotfile output
// Src/F_Interfaces/Base/AMReX_plotfile_fi.cpp:8-26, same name and arguments
void amrex_fi_write_plotfile (const char* name, int nlevs, const MultiFab* mf[], const char* varname[], const Geometry* geom[],
                              Real time, const int level_steps[], const int ref_ratio[])
{
    FI_VOID(
        Vector<const MultiFab*> mfarr(mf, mf + nlevs);
        Vector<std::string> names(varname, varname + mf[0]->nComp());
        Vector<Geometry> geomarr;
        for (int lev = 0; lev < nlevs; ++lev) { geomarr.push_back(*geom[lev]); }
        Vector<int> steps(level_steps, level_steps + nlevs);
        Vector<IntVect> rr;
        for (int lev = 0; lev < nlevs - 1; ++lev) { rr.push_back(IntVect(ref_ratio[lev])); }
        WriteMultiLevelPlotfile(name, nlevs, mfarr, names, geomarr, time, steps, rr); )
}
// VisMF::Write (Src/Base/AMReX_VisMF.H:89): <name>_H + <name>_D_<rank>, ghost cells included
void amrex_b200_vismf_write (const MultiFab* mf, const char* name) { FI_VOID( VisMF::Write(*mf, name); ) }

// --------------------------------------------------------------------------------- linear operators
namespace {
MLLinOp* make_linop (int kind, int nlevels, const Geometry* geom[], const BoxArray* ba[], const DistributionMapping* dm[], LPInfo const& info)
{
    Vector<Geometry> g; Vector<BoxArray> b; Vector<DistributionMapping> d;
    for (int i = 0; i < nlevels; ++i) { g.push_back(*geom[i]); b.push_back(*ba[i]); d.push_back(*dm[i]); }
    if (kind == 0) { return new MLABecLaplacian(g, b, d, info); }
    if (kind == 2) { return new MLALaplacian(g, b, d, info); }
    return new MLPoisson(g, b, d, info);
}
LPInfo make_info (int agglomeration, int consolidation, int max_coarsening_level)
{
    LPInfo info;
    if (agglomeration >= 0) { info.setAgglomeration(agglomeration); }
    if (consolidation >= 0) { info.setConsolidation(consolidation); }
    info.setMaxCoarseningLevel(max_coarsening_level);
    return info;
}
}

void amrex_fi_new_abeclaplacian (MLLinOp** linop, int nlevels, const Geometry* geom[], const BoxArray* ba[], const DistributionMapping* dm[],
                                 int, int agglomeration, int consolidation, int max_coarsening_level)
{
    FI_VOID( *linop = make_linop(0, nlevels, geom, ba, dm, make_info(agglomeration, consolidation, max_coarsening_level)); )
}
void amrex_fi_new_poisson (MLLinOp** linop, int nlevels, const Geometry* geom[], const BoxArray* ba[], const DistributionMapping* dm[],
                           int, int agglomeration, int consolidation, int max_coarsening_level)
{
    FI_VOID( *linop = make_linop(1, nlevels, geom, ba, dm, make_info(agglomeration, consolidation, max_coarsening_level)); )
}
void amrex_b200_new_linop (MLLinOp** linop, int kind, int nlevels, const Geometry* geom[], const BoxArray* ba[], const DistributionMapping* dm[],
                           int agglomeration, int consolidation, int max_coarsening_level, int agg_grid_size, int con_grid_size)
{
    FI_VOID( LPInfo info = make_info(agglomeration, consolidation, max_coarsening_level);
             info.setAgglomerationGridSize(agg_grid_size); info.setConsolidationGridSize(con_grid_size);
             *linop = make_linop(kind, nlevels, geom, ba, dm, info); )
}
void amrex_fi_delete_linop (MLLinOp* linop) { FI_VOID( delete linop; ) }
void amrex_fi_linop_set_maxorder (MLLinOp* linop, int ord) { linop->setMaxOrder(ord); }
void amrex_fi_linop_set_domain_bc (MLLinOp* linop, const int* ilobc, const int* ihibc)
{
    FI_VOID( linop->setDomainBC({LinOpBCType(ilobc[0]), LinOpBCType(ilobc[1]), LinOpBCType(ilobc[2])},
                                {LinOpBCType(ihibc[0]), LinOpBCType(ihibc[1]), LinOpBCType(ihibc[2])}); )
}
void amrex_fi_linop_set_coarse_fine_bc (MLLinOp* linop, const MultiFab* crse, int crse_ratio) { FI_VOID( linop->setCoarseFineBC(crse, crse_ratio); ) }
void amrex_fi_linop_set_level_bc (MLLinOp* linop, int amrlev, const MultiFab* levelbcdata) { FI_VOID( linop->setLevelBC(amrlev, levelbcdata); ) }
void amrex_b200_linop_set_level_bc_robin (MLLinOp* linop, int amrlev, const MultiFab* levelbcdata, const MultiFab* robinbc_a,
                                          const MultiFab* robinbc_b, const MultiFab* robinbc_f)
{
    FI_VOID( linop->setLevelBC(amrlev, levelbcdata, robinbc_a, robinbc_b, robinbc_f); )
}
void amrex_fi_abeclap_set_scalars (MLLinOp* linop, Real a, Real b) { FI_VOID( dynamic_cast<MLABecLaplacian&>(*linop).setScalars(a, b); ) }
void amrex_fi_abeclap_set_acoeffs (MLLinOp* linop, int amrlev, const MultiFab* alpha) { FI_VOID( dynamic_cast<MLABecLaplacian&>(*linop).setACoeffs(amrlev, *alpha); ) }
void amrex_fi_abeclap_set_bcoeffs (MLLinOp* linop, int amrlev, const MultiFab* beta[])
{
    FI_VOID( dynamic_cast<MLABecLaplacian&>(*linop).setBCoeffs(amrlev, {beta[0], beta[1], beta[2]}); )
}
void amrex_b200_linop_set_smoother_fusion (MLLinOp* linop, int fuse) { linop->setSmootherFusion(fuse); }
void amrex_b200_linop_set_gauss_seidel (MLLinOp* linop, int flag) { linop->setGaussSeidel(flag != 0); }   // MLCellLinOpT::setGaussSeidel, AMReX_MLCellLinOp.H:58
void amrex_b200_linop_set_fused_min_box_cells (MLLinOp* linop, long long n) { linop->setFusedMinBoxCells(Long(n)); }
int amrex_b200_set_fused4_plan (int tile_y, int early_stages, int late_stages) { return b200mg_set_gsrb4_plan(tile_y, early_stages, late_stages); }
int amrex_b200_linop_num_mg_levels (const MLLinOp* linop, int amrlev) { return linop->NMGLevels(amrlev); }
void amrex_b200_linop_prepare (MLLinOp* linop) { FI_VOID( linop->prepareForSolve(); ) }
void amrex_b200_linop_make (MLLinOp* linop, MultiFab** mf, int amrlev, int mglev, int ng) { FI_VOID( *mf = new MultiFab(linop->make(amrlev, mglev, ng)); ) }
void amrex_b200_linop_smooth (MLLinOp* linop, int amrlev, int mglev, MultiFab* sol, const MultiFab* rhs, int skip)
{
    FI_VOID( linop->smooth(amrlev, mglev, *sol, *rhs, (skip & 1) != 0, (skip & 2) != 0); )   // bit 1: zero_input
}
void amrex_b200_linop_apply (MLLinOp* linop, int amrlev, int mglev, MultiFab* out, MultiFab* in, int inhomog)
{
    FI_VOID(
        if (inhomog) { Abort("use amrex_b200_linop_residual for inhomogeneous apply"); }
        linop->apply(amrlev, mglev, *out, *in, MLLinOp::BCMode::Homogeneous, MLLinOp::StateMode::Correction); )
}
void amrex_b200_linop_residual (MLLinOp* linop, int amrlev, int mglev, MultiFab* resid, MultiFab* x, const MultiFab* b, int inhomog)
{
    FI_VOID(
        if (inhomog) { AMREX_ALWAYS_ASSERT(mglev == 0); linop->solutionResidual(amrlev, *resid, *x, *b, nullptr); }
        else { linop->correctionResidual(amrlev, mglev, *resid, *x, *b, MLLinOp::BCMode::Homogeneous); } )
}
void amrex_b200_linop_restriction (MLLinOp* linop, int amrlev, int cmglev, MultiFab* crse, MultiFab* fine) { FI_VOID( linop->restriction(amrlev, cmglev, *crse, *fine); ) }
void amrex_b200_linop_interp_add (MLLinOp* linop, int amrlev, int fmglev, MultiFab* fine, const MultiFab* crse)
{
    FI_VOID(
        if (linop->isMFIterSafe(amrlev, fmglev, fmglev + 1)) { linop->interpolation(amrlev, fmglev, *fine, *crse); }
        else {
            MultiFab cfine(amrex::coarsen(fine->boxArray(), 2), fine->DistributionMap(), 1, 0);
            cfine.ParallelCopy(*crse, 0, 0, 1);
            linop->interpolation(amrlev, fmglev, *fine, cfine);
            Gpu::streamSynchronize();
        } )
}
void amrex_b200_linop_get_coeff (MLLinOp* linop, int amrlev, int mglev, int which, const MultiFab** mf)
{
    FI_VOID( auto& ab = dynamic_cast<MLABecLaplacian&>(*linop);
             *mf = (which == 0) ? &ab.getACoeffs(amrlev, mglev) : &ab.getBCoeffs(amrlev, mglev, which - 1); )
}

namespace {
void level_out (MGHierarchy const& H, int a, int m, int* boxes6, int* pmap, int* domain6)
{
    BoxArray const& ba = H.grids[a][m];
    for (int i = 0, N = int(ba.size()); i < N; ++i) {
        for (int d = 0; d < 3; ++d) { boxes6[6 * i + d] = ba[i].smallEnd(d); boxes6[6 * i + 3 + d] = ba[i].bigEnd(d); }
        if (pmap) { pmap[i] = H.dmap[a][m][i]; }
    }
    if (domain6) { Box const& dom = H.geom[a][m].Domain(); for (int d = 0; d < 3; ++d) { domain6[d] = dom.smallEnd(d); domain6[3 + d] = dom.bigEnd(d); } }
}
}
int amrex_b200_linop_level_nboxes (const MLLinOp* linop, int a, int m) { return int(linop->Grids(a, m).size()); }
void amrex_b200_linop_level_boxes (const MLLinOp* linop, int a, int m, int* boxes6, int* pmap, int* domain6) { level_out(linop->hierarchy(), a, m, boxes6, pmap, domain6); }

// --------------------------------------------------------------------------------------------- MLMG
void amrex_fi_new_multigrid (MLMG** mlmg, MLLinOp* lp) { FI_VOID( *mlmg = new MLMG(*lp); (*mlmg)->setThrowException(true); ) }
void amrex_fi_delete_multigrid (MLMG* mlmg) { FI_VOID( delete mlmg; ) }
Real amrex_fi_multigrid_solve (MLMG* mlmg, MultiFab* a_sol[], MultiFab* a_rhs[], Real a_tol_rel, Real a_tol_abs)
{
    FI_TRY
        const int n = mlmg->numAMRLevels();
        return mlmg->solve(Vector<MultiFab*>(a_sol, a_sol + n), Vector<const MultiFab*>(a_rhs, a_rhs + n), a_tol_rel, a_tol_abs);
    FI_CATCH(return kNaN)
}
void amrex_fi_multigrid_comp_residual (MLMG* mlmg, MultiFab* a_res[], MultiFab* a_sol[], MultiFab* a_rhs[])
{
    FI_VOID( const int n = mlmg->numAMRLevels();
             mlmg->compResidual(Vector<MultiFab*>(a_res, a_res + n), Vector<MultiFab*>(a_sol, a_sol + n), Vector<const MultiFab*>(a_rhs, a_rhs + n)); )
}
// a_grad_sol / a_fluxes: [level*3 + dir], face-centred MultiFabs (AMReX_multigrid_fi.cpp:27-51)
void amrex_fi_multigrid_get_grad_solution (MLMG* mlmg, MultiFab* a_grad_sol[])
{
    FI_TRY
        const int n = mlmg->numAMRLevels();
        Vector<Array<MultiFab*, 3>> g(n);
        for (int l = 0; l < n; ++l) { for (int d = 0; d < 3; ++d) { g[l][d] = a_grad_sol[l * 3 + d]; } }
        mlmg->getGradSolution(g);
    FI_CATCH(return)
}
void amrex_fi_multigrid_get_fluxes (MLMG* mlmg, MultiFab* a_fluxes[])
{
    FI_TRY
        const int n = mlmg->numAMRLevels();
        Vector<Array<MultiFab*, 3>> f(n);
        for (int l = 0; l < n; ++l) { for (int d = 0; d < 3; ++d) { f[l][d] = a_fluxes[l * 3 + d]; } }
        mlmg->getFluxes(f);
    FI_CATCH(return)
}
void amrex_fi_multigrid_set_verbose (MLMG* mlmg, int v) { mlmg->setVerbose(v); }
void amrex_fi_multigrid_set_max_iter (MLMG* mlmg, int n) { mlmg->setMaxIter(n); }
void amrex_fi_multigrid_set_max_fmg_iter (MLMG* mlmg, int n) { mlmg->setMaxFmgIter(n); }
void amrex_fi_multigrid_set_fixed_iter (MLMG* mlmg, int n) { mlmg->setFixedIter(n); }
void amrex_fi_multigrid_set_bottom_solver (MLMG* mlmg, int s)
{
    FI_VOID(
        if (s == 0) { mlmg->setBottomSolver(BottomSolver::smoother); }
        else if (s == 1) { mlmg->setBottomSolver(BottomSolver::bicgstab); }
        else if (s == 2) { mlmg->setBottomSolver(BottomSolver::cg); }
        else if (s == 5) { mlmg->setBottomSolver(BottomSolver::bicgcg); }     // 5 / 6: the two-stage solvers (C++ only in the reference;
        else if (s == 6) { mlmg->setBottomSolver(BottomSolver::cgbicg); }     //        its codes 3 / 4 are hypre / petsc: not available)
        else { Abort("amrex_fi_multigrid_set_bottom_solver: unknown or unavailable bottom solver"); } )
}
void amrex_fi_multigrid_set_bottom_verbose (MLMG* mlmg, int n) { mlmg->setBottomVerbose(n); }
void amrex_fi_multigrid_set_always_use_bnorm (MLMG* mlmg, int f) { mlmg->setAlwaysUseBNorm(f); }
void amrex_fi_multigrid_set_final_fill_bc (MLMG* mlmg, int f) { mlmg->setFinalFillBC(f); }
int amrex_b200_multigrid_num_iters (const MLMG* mlmg) { return mlmg->getNumIters(); }
int amrex_b200_multigrid_residual_history (const MLMG* mlmg, Real* hist, int capacity)
{
    auto const& h = mlmg->getResidualHistory();
    for (int i = 0; i < capacity && i < int(h.size()); ++i) { hist[i] = h[i]; }
    return int(h.size());
}
Real amrex_b200_multigrid_init_rhs (const MLMG* mlmg) { return mlmg->getInitRHS(); }
Real amrex_b200_multigrid_init_residual (const MLMG* mlmg) { return mlmg->getInitResidual(); }
int amrex_b200_multigrid_cg_iters (const MLMG* mlmg, int* iters, int capacity)
{
    auto const& h = mlmg->getNumCGIters();
    for (int i = 0; i < capacity && i < int(h.size()); ++i) { iters[i] = h[i]; }
    return int(h.size());
}
void amrex_b200_multigrid_timers (const MLMG* mlmg, double t[3]) { auto a = mlmg->getTimers(); t[0] = a[0]; t[1] = a[1]; t[2] = a[2]; }

// --------------------------------------------------------------------------------------------- GMRES + MLMG
// The reference has no Fortran/C interface for GMRESMLMG (AMReX_GMRES_MLMG.H:19-110 is C++ only); these entries expose
// the same members one to one.
void amrex_b200_new_gmres_mlmg (GMRESMLMG** g, MLMG* mlmg) { FI_VOID( *g = new GMRESMLMG(*mlmg); ) }
void amrex_b200_delete_gmres_mlmg (GMRESMLMG* g) { FI_VOID( delete g; ) }
void amrex_b200_gmres_mlmg_solve (GMRESMLMG* g, MultiFab* sol, const MultiFab* rhs, Real tol_rel, Real tol_abs)
{
    FI_VOID( g->solve(*sol, *rhs, tol_rel, tol_abs); )
}
void amrex_b200_gmres_mlmg_set_verbose (GMRESMLMG* g, int v) { g->setVerbose(v); }
void amrex_b200_gmres_mlmg_set_max_iters (GMRESMLMG* g, int n) { g->setMaxIters(n); }
void amrex_b200_gmres_mlmg_set_restart_length (GMRESMLMG* g, int n) { FI_VOID( g->setRestartLength(n); ) }
void amrex_b200_gmres_mlmg_use_precond (GMRESMLMG* g, int f) { g->usePrecond(f != 0); }
void amrex_b200_gmres_mlmg_set_precond_num_iters (GMRESMLMG* g, int n) { g->setPrecondNumIters(n); }
void amrex_b200_gmres_mlmg_set_property_of_zero (GMRESMLMG* g, int f) { g->setPropertyOfZero(f != 0); }
int amrex_b200_gmres_mlmg_num_iters (const GMRESMLMG* g) { return g->getNumIters(); }
int amrex_b200_gmres_mlmg_status (const GMRESMLMG* g) { return g->getStatus(); }
Real amrex_b200_gmres_mlmg_residual_norm (const GMRESMLMG* g) { return g->getResidualNorm(); }
int amrex_b200_gmres_mlmg_residual_history (const GMRESMLMG* g, Real* hist, int capacity)
{
    auto const& h = g->getResidualHistory();
    for (int i = 0; i < capacity && i < int(h.size()); ++i) { hist[i] = h[i]; }
    return int(h.size());
}

// ------------------------------------------------------------------------------- host-only metadata
void* amrex_b200_hierarchy_new (int nlevels, const Geometry* geom[], const BoxArray* ba[], const DistributionMapping* dm[],
                                int agglomeration, int consolidation, int max_coarsening_level, int agg_grid_size, int con_grid_size, int nprocs)
{
    FI_TRY
        Vector<Geometry> g; Vector<BoxArray> b; Vector<DistributionMapping> d;
        for (int i = 0; i < nlevels; ++i) { g.push_back(*geom[i]); b.push_back(*ba[i]); d.push_back(*dm[i]); }
        LPInfo info = make_info(agglomeration, consolidation, max_coarsening_level);
        info.setAgglomerationGridSize(agg_grid_size); info.setConsolidationGridSize(con_grid_size);
        auto* H = new MGHierarchy;
        H->define(g, b, d, info, nprocs);
        return H;
    FI_CATCH(return nullptr)
}
void amrex_b200_hierarchy_delete (void* h) { delete static_cast<MGHierarchy*>(h); }
int amrex_b200_hierarchy_num_mg_levels (const void* h, int amrlev) { return static_cast<const MGHierarchy*>(h)->num_mg_levels[amrlev]; }
int amrex_b200_hierarchy_nboxes (const void* h, int a, int m) { return int(static_cast<const MGHierarchy*>(h)->grids[a][m].size()); }
int amrex_b200_hierarchy_shares_box_list (const void* h, int a, int m1, int m2) { return static_cast<const MGHierarchy*>(h)->sharesBoxList(a, m1, m2) ? 1 : 0; }
void amrex_b200_hierarchy_level (const void* h, int a, int m, int* boxes6, int* pmap, int* domain6) { level_out(*static_cast<const MGHierarchy*>(h), a, m, boxes6, pmap, domain6); }

namespace {
int write_tags (CommMetaData const& cmd, int kind, int* out, int capacity)
{
    int n = 0;
    auto put = [&] (CopyComTag const& t, int peer) {
        if (out && n < capacity) {
            int* p = out + 15 * n;
            for (int d = 0; d < 3; ++d) { p[d] = t.dbox.smallEnd(d); p[3 + d] = t.dbox.bigEnd(d); p[6 + d] = t.sbox.smallEnd(d); p[9 + d] = t.sbox.bigEnd(d); }
            p[12] = t.dstIndex; p[13] = t.srcIndex; p[14] = peer;
        }
        ++n;
    };
    if (kind == 0) { for (auto const& t : cmd.LocTags) { put(t, -1); } }
    else { for (auto const& kv : (kind == 1 ? cmd.SndTags : cmd.RcvTags)) { for (auto const& t : kv.second) { put(t, kv.first); } } }
    return n;
}
}
int amrex_b200_fb_tags (const BoxArray* ba, const DistributionMapping* dm, int ng, int cross, const int period[3], int myproc, int kind, int* out, int capacity)
{
    FI_TRY
        CommMetaData cmd;
        define_fb_metadata(cmd, *ba, *dm, IntVect(ng), cross != 0, Periodicity(IntVect(period[0], period[1], period[2])), myproc);
        return write_tags(cmd, kind, out, capacity);
    FI_CATCH(return -1)
}
int amrex_b200_fb_face_links (const BoxArray* ba, const DistributionMapping* dm, const int period[3], int myproc, int* out, int capacity)
{
    FI_TRY
        CommMetaData cmd;
        define_fb_metadata(cmd, *ba, *dm, IntVect(1), true, Periodicity(IntVect(period[0], period[1], period[2])), myproc);
        std::vector<b200mg_facelink> links;
        if (!define_fb_face_links(links, cmd, *ba, *dm, myproc)) { return -2; }
        if (out != nullptr) {
            for (std::size_t q = 0; q < links.size() && int(q) < capacity; ++q) {
                out[4 * q] = links[q].fab; out[4 * q + 1] = links[q].shift[0]; out[4 * q + 2] = links[q].shift[1]; out[4 * q + 3] = links[q].shift[2];
            }
        }
        return int(links.size() / 6);
    FI_CATCH(return -1)
}
int amrex_b200_cpc_tags (const BoxArray* ba_dst, const DistributionMapping* dm_dst, int ng_dst, const BoxArray* ba_src, const DistributionMapping* dm_src,
                         int ng_src, const int period[3], int myproc, int kind, int* out, int capacity)
{
    FI_TRY
        CommMetaData cmd;
        define_cpc_metadata(cmd, *ba_dst, *dm_dst, IntVect(ng_dst), *ba_src, *dm_src, IntVect(ng_src),
                            Periodicity(IntVect(period[0], period[1], period[2])), false, myproc);
        return write_tags(cmd, kind, out, capacity);
    FI_CATCH(return -1)
}

} // extern "C"
