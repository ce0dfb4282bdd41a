#include "AMReX_MultiFab.H"
#include <cstdio>
#include <cstdlib>

#include <cuda_runtime.h>
#include <nccl.h>

#include <cstring>
#include <tuple>

namespace amrex {


// =============================================================================== communication metadata
void define_fb_metadata (CommMetaData& cmd, BoxArray const& ba, DistributionMapping const& dm, IntVect const& nghost,
                         bool cross, Periodicity const& period, int MyProc, IntVect const& comm_tile_size)
{
    cmd = CommMetaData();
    std::vector<int> imap;
    for (int i = 0, N = int(ba.size()); i < N; ++i) { if (dm[i] == MyProc) { imap.push_back(i); } }
    const int nlocal = int(imap.size());
    const IntVect ng = nghost;
    std::vector<std::pair<int, Box>> isects;
    const std::vector<IntVect> pshifts = period.shiftIntVect(nghost);

    // send side
    for (int i = 0; i < nlocal; ++i) {
        const int ksnd = imap[i];
        const Box vbx = ba[ksnd];
        for (auto const& pit : pshifts) {
            ba.intersections(vbx + pit, isects, false, ng);
            for (auto const& is : isects) {
                const int krcv = is.first;
                const Box& bx = is.second;
                const int dst_owner = dm[krcv];
                if (dst_owner == MyProc) { continue; }
                BoxList bl = boxDiff(bx, ba[krcv]);
                for (auto const& lit : bl) { cmd.SndTags[dst_owner].emplace_back(lit, lit - pit, krcv, ksnd); }
            }
        }
    }

    // receive side
    for (int i = 0; i < nlocal; ++i) {
        BoxList bl_local(ba.ixType()), bl_remote(ba.ixType());
        const int krcv = imap[i];
        const Box vbx = ba[krcv];
        const Box bxrcv = amrex::grow(vbx, ng);
        for (auto const& pit : pshifts) {
            ba.intersections(bxrcv + pit, isects);
            for (auto const& is : isects) {
                const int ksnd = is.first;
                const Box dst_bx = is.second - pit;
                const int src_owner = dm[ksnd];
                BoxList bl = boxDiff(dst_bx, vbx);
                for (auto const& blbx : bl) {
                    if (src_owner == MyProc) {
                        const BoxList tilelist(blbx, comm_tile_size);
                        for (auto const& t : tilelist) { cmd.LocTags.emplace_back(t, t + pit, krcv, ksnd); }
                        bl_local.push_back(blbx);
                    } else {
                        cmd.RcvTags[src_owner].emplace_back(blbx, blbx + pit, krcv, ksnd);
                        bl_remote.push_back(blbx);
                    }
                }
            }
        }
        if (cmd.threadsafe_loc && bl_local.size() > 1 && !BoxArray(bl_local).isDisjoint()) { cmd.threadsafe_loc = false; }
        if (cmd.threadsafe_rcv && bl_remote.size() > 1 && !BoxArray(bl_remote).isDisjoint()) { cmd.threadsafe_rcv = false; }
    }

    for (int ipass = 0; ipass < 2; ++ipass) {
        auto& Tags = (ipass == 0) ? cmd.SndTags : cmd.RcvTags;
        for (auto& kv : Tags) {
            auto& cctv = kv.second;
            std::sort(cctv.begin(), cctv.end());
            if (!cross) { continue; }
            std::vector<CopyComTag> cross_tags;
            cross_tags.reserve(cctv.size());
            for (auto const& tag : cctv) {
                const Box& bx = tag.dbox;
                const IntVect d2s = tag.sbox.smallEnd() - tag.dbox.smallEnd();
                const Box dstvbx = ba[tag.dstIndex];
                for (int dir = 0; dir < 3; ++dir) {
                    Box lo = dstvbx;
                    lo.setSmall(dir, dstvbx.smallEnd(dir) - ng[dir]); lo.setBig(dir, dstvbx.smallEnd(dir) - 1);
                    lo &= bx;
                    if (lo.ok()) { cross_tags.emplace_back(lo, lo + d2s, tag.dstIndex, tag.srcIndex); }
                    Box hi = dstvbx;
                    hi.setSmall(dir, dstvbx.bigEnd(dir) + 1); hi.setBig(dir, dstvbx.bigEnd(dir) + ng[dir]);
                    hi &= bx;
                    if (hi.ok()) { cross_tags.emplace_back(hi, hi + d2s, tag.dstIndex, tag.srcIndex); }
                }
            }
            if (!cross_tags.empty()) { cctv.swap(cross_tags); }
        }
    }
}

void define_cpc_metadata (CommMetaData& cmd, BoxArray const& ba_dst, DistributionMapping const& dm_dst, IntVect const& ng_dst,
                          BoxArray const& ba_src, DistributionMapping const& dm_src, IntVect const& ng_src,
                          Periodicity const& period, bool tgco, int MyProc, IntVect const& comm_tile_size)
{
    cmd = CommMetaData();
    std::vector<int> imap_src, imap_dst;
    for (int i = 0, N = int(ba_src.size()); i < N; ++i) { if (dm_src[i] == MyProc) { imap_src.push_back(i); } }
    for (int i = 0, N = int(ba_dst.size()); i < N; ++i) { if (dm_dst[i] == MyProc) { imap_dst.push_back(i); } }
    if (imap_src.empty() && imap_dst.empty()) { return; }
    std::vector<std::pair<int, Box>> isects;
    const std::vector<IntVect> pshifts = period.shiftIntVect(ng_dst);

    for (int k_src : imap_src) {
        const Box bx_src = amrex::grow(ba_src[k_src], ng_src);
        for (auto const& pit : pshifts) {
            ba_dst.intersections(bx_src + pit, isects, false, ng_dst);
            for (auto const& is : isects) {
                const int k_dst = is.first;
                const Box& bx = is.second;
                const int dst_owner = dm_dst[k_dst];
                if (dst_owner == MyProc) { continue; }
                BoxList const bl_dst = tgco ? boxDiff(bx, ba_dst[k_dst]) : BoxList(bx);
                for (auto const& b : bl_dst) { cmd.SndTags[dst_owner].emplace_back(b, b - pit, k_dst, k_src); }
            }
        }
    }

    for (int k_dst : imap_dst) {
        BoxList bl_local(ba_dst.ixType()), bl_remote(ba_dst.ixType());
        const Box bx_dst_valid = ba_dst[k_dst];
        const Box bx_dst = amrex::grow(bx_dst_valid, ng_dst);
        for (auto const& pit : pshifts) {
            ba_src.intersections(bx_dst + pit, isects, false, ng_src);
            for (auto const& is : isects) {
                const int k_src = is.first;
                const Box bx = is.second - pit;
                const int src_owner = dm_src[k_src];
                BoxList const bl_dst = tgco ? boxDiff(bx, bx_dst_valid) : BoxList(bx);
                for (auto const& b : bl_dst) {
                    if (src_owner == MyProc) {
                        const BoxList tilelist(b, comm_tile_size);
                        for (auto const& t : tilelist) { cmd.LocTags.emplace_back(t, t + pit, k_dst, k_src); }
                        bl_local.push_back(b);
                    } else {
                        cmd.RcvTags[src_owner].emplace_back(b, b + pit, k_dst, k_src);
                        bl_remote.push_back(b);
                    }
                }
            }
        }
        if (cmd.threadsafe_loc && bl_local.size() > 1 && !BoxArray(bl_local).isDisjoint()) { cmd.threadsafe_loc = false; }
        if (cmd.threadsafe_rcv && bl_remote.size() > 1 && !BoxArray(bl_remote).isDisjoint()) { cmd.threadsafe_rcv = false; }
    }

    for (auto* Tags : {&cmd.SndTags, &cmd.RcvTags}) {
        for (auto& kv : *Tags) { std::sort(kv.second.begin(), kv.second.end()); }
    }
}

// ========================================================================================= LevelLayout
// The cache holds weak references: a layout (device box and tile tables) lives as long as a FabArray or an operator level
// uses it, and its entry goes when the BoxArray / DistributionMapping it is keyed by dies (evict_by_id below).  The maps
// are heap objects that are never destroyed, so Ref destructors running during static destruction find them intact.
namespace {
    using LayoutMap = std::map<std::pair<std::uint64_t, std::uint64_t>, std::weak_ptr<LevelLayout>>;
    LayoutMap& layouts () { static LayoutMap* m = new LayoutMap; return *m; }
    void evict_by_id (std::uint64_t id, int kind);
    const bool g_hook_set = (setRefDeathHook(&evict_by_id), true);
}

std::shared_ptr<LevelLayout> LevelLayout::get (BoxArray const& ba, DistributionMapping const& dm)
{
    AMREX_ALWAYS_ASSERT(ba.size() == dm.size());
    auto key = std::make_pair(ba.id(), dm.id());
    auto it = layouts().find(key);
    if (it != layouts().end()) { if (auto sp = it->second.lock()) { return sp; } }
    auto L = std::shared_ptr<LevelLayout>(new LevelLayout);
    const int me = ParallelDescriptor::MyProc();
    std::vector<b200mg_box> hb;
    for (int i = 0, N = int(ba.size()); i < N; ++i) {
        if (dm[i] != me) { continue; }
        L->m_g2l[i] = int(L->m_index.size());
        L->m_index.push_back(i);
        L->m_boxes.push_back(ba[i]);
        L->m_cells += ba[i].numPts();
        if (ba[i].length(0) % 2 != 0 || ba[i].length(0) > 128 || !ba[i].cellCentered()) { L->m_pairable = false; }
        if (ba[i].length(0) < 4 || ba[i].length(1) < 2) { L->m_lean = false; }
        b200mg_box b; for (int d = 0; d < 3; ++d) { b.lo[d] = ba[i].smallEnd(d); b.hi[d] = ba[i].bigEnd(d); }
        hb.push_back(b);
    }
    L->m_dvbox.assign(hb);
    layouts()[key] = L;
    return L;
}

void LevelLayout::clearCache () { layouts().clear(); }

LevelLayout::Tiles const& LevelLayout::tiles (int ng)
{
    auto it = m_tiles.find(ng);
    if (it != m_tiles.end()) { return it->second; }
    // tile depth: deep tiles on big levels (fewer, longer-running CTAs that stream along z), 4 planes on small ones
    // so that coarse levels still spread over the SMs
    int tz = B200MG_TILE_Z;
    while (tz < 16 && m_cells / (Long(64) * B200MG_TILE_Y * 2 * tz) >= 8 * 148) { tz *= 2; }
    std::vector<b200mg_tile> ht;
    for (int li = 0; li < numLocal(); ++li) {
        const Box g = amrex::grow(m_boxes[li], ng);
        for (int k = g.smallEnd(2); k <= g.bigEnd(2); k += tz)
            for (int j = g.smallEnd(1); j <= g.bigEnd(1); j += B200MG_TILE_Y) { ht.push_back(b200mg_tile{li, j, k, tz}); }
    }
    Tiles& T = m_tiles[ng];
    T.n = int(ht.size());
    T.d.assign(ht);
    return T;
}

// ============================================================================================ FabArray
template <class T>
void FabArray<T>::define (BoxArray const& ba, DistributionMapping const& dm, int ncomp, int ngrow)
{
    clear();
    m_ba = ba; m_dm = dm; m_ncomp = ncomp; m_ngrow = ngrow;
    m_layout = LevelLayout::get(ba, dm);
    const int nl = m_layout->numLocal();
    m_hdesc.resize(nl);
    std::vector<std::size_t> offs(nl);
    const int xoff = (16 - ngrow % 16) % 16;     // first VALID cell of each row lands on a 128-byte boundary
    std::size_t total = 0;
    for (int li = 0; li < nl; ++li) {
        const Box g = amrex::grow(m_layout->box(li), ngrow);
        const long long pitch = ((xoff + g.length(0)) + 15) / 16 * 16;
        Desc& d = m_hdesc[li];
        for (int a = 0; a < 3; ++a) { d.lo[a] = g.smallEnd(a); d.hi[a] = g.bigEnd(a); }
        d.jstride = pitch; d.kstride = pitch * g.length(1); d.nstride = d.kstride * g.length(2);
        offs[li] = total + xoff;
        total += std::size_t(d.nstride) * ncomp;
    }
    m_nelem = total;
    if (total > 0) {
        m_data = static_cast<T*>(The_Arena()->alloc(total * sizeof(T)));
        // padding and ghost cells start as zero so no kernel ever reads indeterminate bits
        Gpu::memset_async(m_data, 0, total * sizeof(T));
    }
    for (int li = 0; li < nl; ++li) { m_hdesc[li].p = m_data + offs[li]; }
    // device table: [comp][local fab]; the entries of component c point at that component's first element, so the
    // single-component kernels work on any component of a multi-component array (d_fabs(c))
    std::vector<Desc> all(std::size_t(nl) * std::max(ncomp, 1));
    for (int c = 0; c < std::max(ncomp, 1); ++c) {
        for (int li = 0; li < nl; ++li) { Desc d = m_hdesc[li]; if (d.p) { d.p += std::size_t(c) * d.nstride; } all[std::size_t(c) * nl + li] = d; }
    }
    m_ddesc.assign(all);
}

template <class T>
void FabArray<T>::clear ()
{
    if (m_data) { The_Arena()->free(m_data); m_data = nullptr; }
    m_ddesc.clear(); m_hdesc.clear(); m_layout.reset(); m_nelem = 0;
}

template <class T>
void FabArray<T>::swap (FabArray& o) noexcept
{
    std::swap(m_ba, o.m_ba); std::swap(m_dm, o.m_dm); std::swap(m_ncomp, o.m_ncomp); std::swap(m_ngrow, o.m_ngrow);
    std::swap(m_layout, o.m_layout); std::swap(m_data, o.m_data); std::swap(m_nelem, o.m_nelem);
    std::swap(m_hdesc, o.m_hdesc); std::swap(m_ddesc, o.m_ddesc);
}

namespace {
template <class T, class DESC>
void copy3d (DESC const& d, T* h, Box const& region, Box const& isect, int comp, bool to_host)
{
    cudaMemcpy3DParms p; std::memset(&p, 0, sizeof(p));
    const std::size_t hnx = region.length(0), hny = region.length(1);
    T* hp = h + (isect.smallEnd(0) - region.smallEnd(0)) + hnx * ((isect.smallEnd(1) - region.smallEnd(1))
            + hny * std::size_t(isect.smallEnd(2) - region.smallEnd(2)));
    const int ny = d.hi[1] - d.lo[1] + 1;
    T* dp = d.p + (isect.smallEnd(0) - d.lo[0]) + (isect.smallEnd(1) - d.lo[1]) * d.jstride
            + (isect.smallEnd(2) - d.lo[2]) * d.kstride + comp * d.nstride;
    cudaPitchedPtr hpp = make_cudaPitchedPtr(hp, hnx * sizeof(T), hnx, hny);
    cudaPitchedPtr dpp = make_cudaPitchedPtr(dp, d.jstride * sizeof(T), d.jstride, ny);
    p.extent = make_cudaExtent(isect.length(0) * sizeof(T), isect.length(1), isect.length(2));
    if (to_host) { p.srcPtr = dpp; p.dstPtr = hpp; p.kind = cudaMemcpyDeviceToHost; }
    else { p.srcPtr = hpp; p.dstPtr = dpp; p.kind = cudaMemcpyHostToDevice; }
    AMREX_CUDA_SAFE_CALL(cudaMemcpy3DAsync(&p, Gpu::gpuStream()));
}
}

template <class T>
void FabArray<T>::copyFromHost (const T* h, Box const& region, int comp, int ng, bool sync)
{
    AMREX_ALWAYS_ASSERT(ng <= m_ngrow && region.ixType() == ixType());
    for (int li = 0; li < local_size(); ++li) {
        Box isect = amrex::grow(validbox(li), ng) & region;
        if (isect.ok()) { copy3d(m_hdesc[li], const_cast<T*>(h), region, isect, comp, false); }
    }
    if (sync) { Gpu::streamSynchronize(); }
}

template <class T>
void FabArray<T>::copyToHost (T* h, Box const& region, int comp, int ng, bool valid_wins, bool sync) const
{
    AMREX_ALWAYS_ASSERT(ng <= m_ngrow && region.ixType() == ixType());
    const int npass = (ng > 0 && valid_wins) ? 2 : 1;
    for (int pass = 0; pass < npass; ++pass) {
        for (int li = 0; li < local_size(); ++li) {
            Box isect = ((pass == 0) ? amrex::grow(validbox(li), ng) : validbox(li)) & region;
            if (isect.ok()) { copy3d(m_hdesc[li], h, region, isect, comp, true); }
        }
    }
    if (sync) { Gpu::streamSynchronize(); }
}

// one local fab (its own valid + ng ghost cells, nothing from its neighbours), Fortran order over the grown box
template <class T>
void FabArray<T>::copyFabToHost (int li, T* h, int comp, int ng) const
{
    AMREX_ALWAYS_ASSERT(ng <= m_ngrow && li >= 0 && li < local_size() && comp >= 0 && comp < m_ncomp);
    const Box g = amrex::grow(validbox(li), ng);
    copy3d(m_hdesc[li], h, g, g, comp, true);
    Gpu::streamSynchronize();
}

template class FabArray<double>;
template class FabArray<int>;

// ============================================================================================ reductions
namespace {
    double* g_red_slots = nullptr;       // device: 8 result slots
    double* g_red_host = nullptr;        // pinned
    double* g_red_scratch = nullptr; long long g_red_scratch_n = 0;
}

double* reduce_result_slot (int which)
{
    if (!g_red_slots) {
        g_red_slots = static_cast<double*>(The_Arena()->alloc(8 * sizeof(double)));
        g_red_host = static_cast<double*>(pinned_alloc(8 * sizeof(double)));
    }
    return g_red_slots + which;
}

double* reduce_scratch (int ntiles)
{
    const long long need = b200mg_reduce_scratch_doubles(ntiles);
    if (need > g_red_scratch_n) {
        if (g_red_scratch) { Gpu::streamSynchronize(); The_Arena()->free(g_red_scratch); }
        g_red_scratch_n = std::max(need, (long long)(1 << 16));
        g_red_scratch = static_cast<double*>(The_Arena()->alloc(g_red_scratch_n * sizeof(double)));
        Gpu::memset_async(g_red_scratch, 0, g_red_scratch_n * sizeof(double));
    } else {
        // the ticket slot of this launch size must be zero: kernels reset it, but sizes vary between calls
        Gpu::memset_async(g_red_scratch + ntiles, 0, sizeof(double));
    }
    return g_red_scratch;
}

double fetch_reduce_result (int which)
{
    reduce_result_slot(0);
    Gpu::dtoh_memcpy_async(g_red_host + which, g_red_slots + which, sizeof(double));
    Gpu::streamSynchronize();
    return g_red_host[which];
}

// ============================================================================================== MultiFab
void MultiFab::setVal (Real v, int comp, int ncomp, int ng)
{
    AMREX_ALWAYS_ASSERT(comp >= 0 && ncomp >= 1 && comp + ncomp <= m_ncomp && ng <= m_ngrow);
    auto const& T = layout().tiles(ng);
    for (int n = 0; n < ncomp; ++n) {
        B200_KCALL(b200mg_setval(T.n, T.d.data(), layout().d_vbox(), d_fabs(comp + n), v, ng, Gpu::gpuStream()));
    }
}

void MultiFab::setBndry (Real v)
{
    if (m_ngrow == 0) { return; }
    auto const& T = layout().tiles(m_ngrow);
    B200_KCALL(b200mg_setbndry(T.n, T.d.data(), layout().d_vbox(), d_fabs(), v, m_ngrow, Gpu::gpuStream()));
}

// every component, like the reference's FabArray::plus / mult (AMReX_FabArray.H:2932-2990)
void MultiFab::plus (Real v, int ng)
{
    auto const& T = layout().tiles(ng);
    for (int n = 0; n < m_ncomp; ++n) {
        B200_KCALL(b200mg_plus(T.n, T.d.data(), layout().d_vbox(), d_fabs(n), v, ng, Gpu::gpuStream()));
    }
}

void MultiFab::mult (Real v, int ng)
{
    auto const& T = layout().tiles(ng);
    for (int n = 0; n < m_ncomp; ++n) {
        B200_KCALL(b200mg_lincomb(T.n, T.d.data(), layout().d_vbox(), d_fabs(n), 0.0, d_fabs(n), v, ng, Gpu::gpuStream()));
    }
}

Real MultiFab::norminf (int comp, bool local) const
{
    AMREX_ALWAYS_ASSERT(comp >= 0 && comp < m_ncomp);
    auto const& T = layout().tiles(0);
    B200_KCALL(b200mg_norminf(T.n, T.d.data(), layout().d_vbox(), d_fabs(comp), nullptr, reduce_result_slot(0), reduce_scratch(T.n), Gpu::gpuStream()));
    double r = fetch_reduce_result(0);
    if (!local) { ParallelDescriptor::ReduceRealMax(&r, 1); }
    return r;
}

Real MultiFab::norm1 (int comp, bool local) const
{
    AMREX_ALWAYS_ASSERT(comp >= 0 && comp < m_ncomp);
    auto const& T = layout().tiles(0);
    B200_KCALL(b200mg_asum(T.n, T.d.data(), layout().d_vbox(), d_fabs(comp), reduce_result_slot(0), reduce_scratch(T.n), Gpu::gpuStream()));
    double r = fetch_reduce_result(0);
    if (!local) { ParallelDescriptor::ReduceRealSum(&r, 1); }
    return r;
}

Real MultiFab::norm2 (int comp) const
{
    AMREX_ALWAYS_ASSERT(comp >= 0 && comp < m_ncomp);
    auto const& T = layout().tiles(0);
    B200_KCALL(b200mg_dot(T.n, T.d.data(), layout().d_vbox(), d_fabs(comp), d_fabs(comp), reduce_result_slot(0), reduce_scratch(T.n), Gpu::gpuStream()));
    double r = fetch_reduce_result(0);
    ParallelDescriptor::ReduceRealSum(&r, 1);
    return std::sqrt(r);
}

Real MultiFab::norminf (iMultiFab const& mask, bool local) const
{
    AMREX_ALWAYS_ASSERT_WITH_MESSAGE(m_ncomp == 1, "MultiFab::norminf(mask): single-component arrays only");
    auto const& T = layout().tiles(0);
    B200_KCALL(b200mg_norminf(T.n, T.d.data(), layout().d_vbox(), d_fabs(), mask.d_fabs(), reduce_result_slot(0), reduce_scratch(T.n), Gpu::gpuStream()));
    double r = fetch_reduce_result(0);
    if (!local) { ParallelDescriptor::ReduceRealMax(&r, 1); }
    return r;
}

Real MultiFab::sum (bool local) const
{
    AMREX_ALWAYS_ASSERT_WITH_MESSAGE(m_ncomp == 1, "MultiFab::sum: single-component arrays only");
    auto const& T = layout().tiles(0);
    B200_KCALL(b200mg_sum(T.n, T.d.data(), layout().d_vbox(), d_fabs(), reduce_result_slot(0), reduce_scratch(T.n), Gpu::gpuStream()));
    double r = fetch_reduce_result(0);
    if (!local) { ParallelDescriptor::ReduceRealSum(&r, 1); }
    return r;
}

Real MultiFab::min (int comp, int nghost, bool local) const
{
    auto const& T = layout().tiles(nghost);
    B200_KCALL(b200mg_minmax(T.n, T.d.data(), layout().d_vbox(), d_fabs(comp), 0, nghost, reduce_result_slot(0), reduce_scratch(T.n), Gpu::gpuStream()));
    double r = fetch_reduce_result(0);
    if (!local) { ParallelDescriptor::ReduceRealMin(&r, 1); }
    return r;
}

Real MultiFab::max (int comp, int nghost, bool local) const
{
    auto const& T = layout().tiles(nghost);
    B200_KCALL(b200mg_minmax(T.n, T.d.data(), layout().d_vbox(), d_fabs(comp), 1, nghost, reduce_result_slot(0), reduce_scratch(T.n), Gpu::gpuStream()));
    double r = fetch_reduce_result(0);
    if (!local) { ParallelDescriptor::ReduceRealMax(&r, 1); }
    return r;
}

void MultiFab::Multiply (MultiFab& dst, MultiFab const& src, int scomp, int dcomp, int ncomp, int ng)
{
    auto const& T = dst.layout().tiles(ng);
    for (int n = 0; n < ncomp; ++n) {
        B200_KCALL(b200mg_multiply(T.n, T.d.data(), dst.layout().d_vbox(), dst.d_fabs(dcomp + n), src.d_fabs(scomp + n), ng, Gpu::gpuStream()));
    }
}

void MultiFab::Divide (MultiFab& dst, MultiFab const& src, int scomp, int dcomp, int ncomp, int ng)
{
    auto const& T = dst.layout().tiles(ng);
    for (int n = 0; n < ncomp; ++n) {
        B200_KCALL(b200mg_divide(T.n, T.d.data(), dst.layout().d_vbox(), dst.d_fabs(dcomp + n), src.d_fabs(scomp + n), ng, Gpu::gpuStream()));
    }
}

void MultiFab::SumBoundary (int scomp, int ncomp, Periodicity const& period)
{
    if (m_ngrow == 0 && boxArray().ixType().cellCentered()) { return; }
    MultiFab tmp(boxArray(), DistributionMap(), ncomp, m_ngrow);
    MultiFab::Copy(tmp, *this, scomp, 0, ncomp, m_ngrow);
    setVal(0.0, scomp, ncomp, 0);
    ParallelCopy(tmp, 0, scomp, ncomp, m_ngrow, 0, period, CpOp::ADD);
}

Real MultiFab::Dot (MultiFab const& x, MultiFab const& y, bool local)
{
    AMREX_ALWAYS_ASSERT_WITH_MESSAGE(x.nComp() == 1 && y.nComp() == 1, "MultiFab::Dot: single-component arrays only");
    auto const& T = x.layout().tiles(0);
    B200_KCALL(b200mg_dot(T.n, T.d.data(), x.layout().d_vbox(), x.d_fabs(), y.d_fabs(), reduce_result_slot(0), reduce_scratch(T.n), Gpu::gpuStream()));
    double r = fetch_reduce_result(0);
    if (!local) { ParallelDescriptor::ReduceRealSum(&r, 1); }
    return r;
}

void MultiFab::MultiDot (MultiFab const& x, Vector<MultiFab const*> const& v, Real* out, bool local)
{
    const int nv = int(v.size());
    if (nv == 0) { return; }
    auto const& T = x.layout().tiles(0);
    static double* d_res = nullptr; static double* h_res = nullptr;
    static double* d_scr = nullptr; static long long scr_n = 0;
    if (!d_res) {
        d_res = static_cast<double*>(The_Arena()->alloc(B200MG_KRYLOV_GROUP * sizeof(double)));
        h_res = static_cast<double*>(pinned_alloc(B200MG_KRYLOV_GROUP * sizeof(double)));
    }
    const long long need = b200mg_multi_dot_scratch_doubles(T.n);
    if (need > scr_n) {
        if (d_scr) { Gpu::streamSynchronize(); The_Arena()->free(d_scr); }
        scr_n = need; d_scr = static_cast<double*>(The_Arena()->alloc(scr_n * sizeof(double)));
    }
    for (int n0 = 0; n0 < nv; n0 += B200MG_KRYLOV_GROUP) {
        const int g = std::min(B200MG_KRYLOV_GROUP, nv - n0);
        const b200mg_fab* tabs[B200MG_KRYLOV_GROUP];
        for (int n = 0; n < g; ++n) {
            AMREX_ALWAYS_ASSERT(v[n0 + n]->layoutPtr() == x.layoutPtr() || (v[n0 + n]->boxArray() == x.boxArray() && v[n0 + n]->DistributionMap() == x.DistributionMap()));
            tabs[n] = v[n0 + n]->d_fabs();
        }
        B200_KCALL(b200mg_multi_dot(T.n, T.d.data(), x.layout().d_vbox(), x.d_fabs(), g, tabs, d_res, d_scr, Gpu::gpuStream()));
        Gpu::dtoh_memcpy_async(h_res, d_res, g * sizeof(double));
        Gpu::streamSynchronize();
        for (int n = 0; n < g; ++n) { out[n0 + n] = h_res[n]; }
    }
    if (!local) { ParallelDescriptor::ReduceRealSum(out, nv); }
}

void MultiFab::MultiSaxpy (MultiFab& w, Vector<MultiFab const*> const& v, const Real* a)
{
    const int nv = int(v.size());
    auto const& T = w.layout().tiles(0);
    for (int n0 = 0; n0 < nv; n0 += B200MG_KRYLOV_GROUP) {
        const int g = std::min(B200MG_KRYLOV_GROUP, nv - n0);
        const b200mg_fab* tabs[B200MG_KRYLOV_GROUP];
        for (int n = 0; n < g; ++n) {
            AMREX_ALWAYS_ASSERT(v[n0 + n]->layoutPtr() == w.layoutPtr() || (v[n0 + n]->boxArray() == w.boxArray() && v[n0 + n]->DistributionMap() == w.DistributionMap()));
            tabs[n] = v[n0 + n]->d_fabs();
        }
        B200_KCALL(b200mg_multi_axpy(T.n, T.d.data(), w.layout().d_vbox(), w.d_fabs(), g, tabs, a + n0, Gpu::gpuStream()));
    }
}

namespace {
void check_same (MultiFab const& a, MultiFab const& b, int scomp, int dcomp, int ncomp, int ng)
{
    AMREX_ALWAYS_ASSERT_WITH_MESSAGE(a.layoutPtr() == b.layoutPtr() || (a.boxArray() == b.boxArray() && a.DistributionMap() == b.DistributionMap()),
                                     "MultiFab op: operands must share BoxArray and DistributionMapping");
    AMREX_ALWAYS_ASSERT(scomp >= 0 && dcomp >= 0 && ncomp >= 1 && scomp + ncomp <= b.nComp() && dcomp + ncomp <= a.nComp()
                        && ng <= a.nGrow() && ng <= b.nGrow());
}
}

void MultiFab::Copy (MultiFab& dst, MultiFab const& src, int scomp, int dcomp, int ncomp, int ng)
{
    check_same(dst, src, scomp, dcomp, ncomp, ng);
    auto const& T = dst.layout().tiles(ng);
    for (int n = 0; n < ncomp; ++n) {
        B200_KCALL(b200mg_copy(T.n, T.d.data(), dst.layout().d_vbox(), dst.d_fabs(dcomp + n), src.d_fabs(scomp + n), ng, Gpu::gpuStream()));
    }
}

// dst(dcomp..) = a * x(scomp..) + b * dst(dcomp..), component by component
void MultiFab::LinComb (MultiFab& dst, Real a, MultiFab const& x, Real b, int scomp, int dcomp, int ncomp, int ng)
{
    check_same(dst, x, scomp, dcomp, ncomp, ng);
    auto const& T = dst.layout().tiles(ng);
    for (int n = 0; n < ncomp; ++n) {
        B200_KCALL(b200mg_lincomb(T.n, T.d.data(), dst.layout().d_vbox(), dst.d_fabs(dcomp + n), a, x.d_fabs(scomp + n), b, ng, Gpu::gpuStream()));
    }
}

void MultiFab::LinComb (MultiFab& dst, Real a, MultiFab const& x, Real b, int ng) { LinComb(dst, a, x, b, 0, 0, 1, ng); }

void MultiFab::Add (MultiFab& dst, MultiFab const& src, int scomp, int dcomp, int ncomp, int ng)
{ LinComb(dst, 1.0, src, 1.0, scomp, dcomp, ncomp, ng); }
void MultiFab::Subtract (MultiFab& dst, MultiFab const& src, int scomp, int dcomp, int ncomp, int ng)
{ LinComb(dst, -1.0, src, 1.0, scomp, dcomp, ncomp, ng); }
void MultiFab::Saxpy (MultiFab& dst, Real a, MultiFab const& src, int scomp, int dcomp, int ncomp, int ng)
{ LinComb(dst, a, src, 1.0, scomp, dcomp, ncomp, ng); }
void MultiFab::Xpay (MultiFab& dst, Real a, MultiFab const& src, int scomp, int dcomp, int ncomp, int ng)
{ LinComb(dst, 1.0, src, a, scomp, dcomp, ncomp, ng); }

// ================================================================================== halo exchange plans
bool define_fb_face_links (std::vector<b200mg_facelink>& links, CommMetaData const& cmd, std::vector<int> const& local_index,
                           std::vector<Box> const& local_boxes)
{
    links.assign(6 * local_index.size(), b200mg_facelink{-1, {0, 0, 0}});
    auto local_of = [&] (int g) {
        auto it = std::lower_bound(local_index.begin(), local_index.end(), g);
        return (it != local_index.end() && *it == g) ? int(it - local_index.begin()) : -1;
    };
    bool ok = true;
    for (auto const& t : cmd.LocTags) {
        const int ld = local_of(t.dstIndex), ls = local_of(t.srcIndex);
        if (ld < 0 || ls < 0) { return false; }
        Box const& vbx = local_boxes[ld];
        const IntVect d2s = t.sbox.smallEnd() - t.dbox.smallEnd();
        for (int dir = 0; dir < 3; ++dir) {
            for (int side = 0; side < 2; ++side) {
                Box face = vbx;                          // the one-cell slab of ghost cells behind this face
                if (side == 0) { face.setSmall(dir, vbx.smallEnd(dir) - 1); face.setBig(dir, vbx.smallEnd(dir) - 1); }
                else { face.setSmall(dir, vbx.bigEnd(dir) + 1); face.setBig(dir, vbx.bigEnd(dir) + 1); }
                Box slab = face; slab &= t.dbox;
                if (!slab.ok()) { continue; }            // (edge / corner pieces of a tag: a cross stencil never reads them)
                b200mg_facelink& l = links[std::size_t(ld) * 6 + dir + 3 * side];
                if (slab == face && l.fab < 0) { l.fab = ls; for (int d = 0; d < 3; ++d) { l.shift[d] = d2s[d]; } }
                else { ok = false; }                     // a face fed by several boxes, or in part
            }
        }
    }
    return ok;
}

bool define_fb_face_links (std::vector<b200mg_facelink>& links, CommMetaData const& cmd, BoxArray const& ba,
                           DistributionMapping const& dm, int myproc)
{
    std::vector<int> idx; std::vector<Box> boxes;
    for (int g = 0; g < int(ba.size()); ++g) { if (dm[g] == myproc) { idx.push_back(g); boxes.push_back(ba[g]); } }
    return define_fb_face_links(links, cmd, idx, boxes);
}

namespace {

struct CommPlan {
    CommMetaData meta;
    DeviceTable<b200mg_copytag> d_loc, d_snd, d_rcv;
    int nloc = 0, nsnd = 0, nrcv = 0;
    int maxloc = 0, maxsnd = 0, maxrcv = 0;              // points of the largest tag of each list (grid sizing)
    struct Peer { int rank; long long offset, count; };
    std::vector<Peer> snd_peers, rcv_peers;
    long long snd_total = 0, rcv_total = 0;
    std::vector<int> remote_fabs;                        // local indices of the destination fabs that receive remote data (sorted)
    // cross-stencil FillBoundary with one ghost cell whose every LOCAL tag is a whole box face copied from one local fab: the
    // local part of the exchange as a table of face links [local fab * 6 + face] (b200mg_facelink); faces fed by other ranks
    // (receive tags) have no link
    DeviceTable<b200mg_facelink> d_links;
    bool links_ok = false;
    double *sndbuf = nullptr, *rcvbuf = nullptr;
    long long buf_ncomp = 0;
    cudaEvent_t ev_packed = nullptr, ev_arrived = nullptr;
    ~CommPlan ()
    {
        if (sndbuf) { The_Arena()->free(sndbuf); } if (rcvbuf) { The_Arena()->free(rcvbuf); }
        if (ev_packed) { cudaEventDestroy(ev_packed); } if (ev_arrived) { cudaEventDestroy(ev_arrived); }
    }
};

b200mg_copytag make_tag (CopyComTag const& t, int dst_fab, int src_fab, long long off)
{
    b200mg_copytag r;
    for (int d = 0; d < 3; ++d) {
        r.lo[d] = t.dbox.smallEnd(d); r.hi[d] = t.dbox.bigEnd(d);
        r.shift[d] = t.sbox.smallEnd(d) - t.dbox.smallEnd(d);
    }
    r.dst_fab = dst_fab; r.src_fab = src_fab; r.pad = 0; r.buf_offset = off;
    return r;
}

// cross_ng != nullptr: the plan serves a cross-stencil FillBoundary with that many ghost cells - local tags are clipped to the
// face slabs of the destination box, as the reference does for its send / receive tags (AMReX_FabArrayBase.cpp:835-870); the
// edge and corner pieces (which its local list keeps, and which a cross stencil never reads) are not copied
void finish_plan (CommPlan& P, LevelLayout const& ldst, LevelLayout const& lsrc, const IntVect* cross_ng = nullptr)
{
    std::vector<b200mg_copytag> h;
    auto npts = [] (CopyComTag const& t) { return int(std::min<Long>(t.dbox.numPts(), Long(1) << 30)); };
    for (auto const& t : P.meta.LocTags) {
        const int ld = ldst.localIndex(t.dstIndex), ls = lsrc.localIndex(t.srcIndex);
        if (cross_ng == nullptr) { h.push_back(make_tag(t, ld, ls, 0)); P.maxloc = std::max(P.maxloc, npts(t)); continue; }
        Box const& vbx = ldst.box(ld);
        const IntVect d2s = t.sbox.smallEnd() - t.dbox.smallEnd();
        for (int dir = 0; dir < 3; ++dir) {
            for (int side = 0; side < 2; ++side) {
                Box face = vbx;
                if (side == 0) { face.setSmall(dir, vbx.smallEnd(dir) - (*cross_ng)[dir]); face.setBig(dir, vbx.smallEnd(dir) - 1); }
                else { face.setSmall(dir, vbx.bigEnd(dir) + 1); face.setBig(dir, vbx.bigEnd(dir) + (*cross_ng)[dir]); }
                Box slab = face; slab &= t.dbox;
                if (!slab.ok()) { continue; }
                CopyComTag c = t;
                c.dbox = slab; c.sbox = slab + d2s;
                h.push_back(make_tag(c, ld, ls, 0)); P.maxloc = std::max(P.maxloc, npts(c));
            }
        }
    }
    P.nloc = int(h.size()); P.d_loc.assign(h);
    if (cross_ng != nullptr && *cross_ng == IntVect(1) && &ldst == &lsrc && P.nloc > 0) {
        std::vector<b200mg_facelink> links;
        std::vector<Box> boxes;
        for (int li = 0; li < ldst.numLocal(); ++li) { boxes.push_back(ldst.box(li)); }
        P.links_ok = define_fb_face_links(links, P.meta, ldst.indexArray(), boxes);
        if (P.links_ok) { P.d_links.assign(links); }
    }
    h.clear();
    long long off = 0;
    for (auto const& kv : P.meta.SndTags) {
        const long long start = off;
        for (auto const& t : kv.second) { h.push_back(make_tag(t, -1, lsrc.localIndex(t.srcIndex), off)); off += t.dbox.numPts(); P.maxsnd = std::max(P.maxsnd, npts(t)); }
        P.snd_peers.push_back({kv.first, start, off - start});
    }
    P.snd_total = off; P.nsnd = int(h.size()); P.d_snd.assign(h);
    h.clear(); off = 0;
    for (auto const& kv : P.meta.RcvTags) {
        const long long start = off;
        for (auto const& t : kv.second) {
            h.push_back(make_tag(t, ldst.localIndex(t.dstIndex), -1, off)); off += t.dbox.numPts(); P.maxrcv = std::max(P.maxrcv, npts(t));
            P.remote_fabs.push_back(ldst.localIndex(t.dstIndex));
        }
        P.rcv_peers.push_back({kv.first, start, off - start});
    }
    std::sort(P.remote_fabs.begin(), P.remote_fabs.end());
    P.remote_fabs.erase(std::unique(P.remote_fabs.begin(), P.remote_fabs.end()), P.remote_fabs.end());
    P.rcv_total = off; P.nrcv = int(h.size()); P.d_rcv.assign(h);
}

// Halo exchange of one plan in two halves (FillBoundary_nowait / FillBoundary_finish of the reference,
// AMReX_FabArrayCommI.H:8-247, with the local copies in between as in FBEP_nowait :118-135).  start: pack -> [event] ->
// grouped ncclSend/ncclRecv on the communication stream (NVLink), and meanwhile the intra-GPU copies on the compute stream;
// finish: the compute stream waits for the transfer and unpacks.  Whatever the caller launches on the compute stream between
// the two halves (the smoother on the boxes without remote neighbours) overlaps the transfer.
void start_plan (CommPlan& P, MultiFab& dst, MultiFab const& src, int scomp, int dcomp, int ncomp, CpOp op, int parity = -1,
                 bool remote_only = false)
{
    cudaStream_t s = Gpu::gpuStream();
    const bool remote = (P.snd_total + P.rcv_total) > 0;
    static const bool overlap = std::getenv("B200MG_NO_COMM_OVERLAP") == nullptr;
    if (remote) {
        if (P.buf_ncomp < ncomp) {
            Gpu::streamSynchronize();
            if (P.sndbuf) { The_Arena()->free(P.sndbuf); } if (P.rcvbuf) { The_Arena()->free(P.rcvbuf); }
            P.sndbuf = static_cast<double*>(The_Arena()->alloc(std::max<long long>(1, P.snd_total * ncomp) * sizeof(double)));
            P.rcvbuf = static_cast<double*>(The_Arena()->alloc(std::max<long long>(1, P.rcv_total * ncomp) * sizeof(double)));
            P.buf_ncomp = ncomp;
        }
        if (!P.ev_packed) {
            AMREX_CUDA_SAFE_CALL(cudaEventCreateWithFlags(&P.ev_packed, cudaEventDisableTiming));
            AMREX_CUDA_SAFE_CALL(cudaEventCreateWithFlags(&P.ev_arrived, cudaEventDisableTiming));
        }
        B200_KCALL(b200mg_copy_tags_colour(P.nsnd, P.d_snd.data(), nullptr, src.d_fabs(), P.sndbuf, ncomp, scomp, dcomp, 0, P.maxsnd, parity, s));
        ncclComm_t comm = static_cast<ncclComm_t>(ParallelDescriptor::Comm());
        AMREX_ALWAYS_ASSERT_WITH_MESSAGE(comm != nullptr, "multi-rank exchange without an NCCL communicator");
        cudaStream_t cs = overlap ? Gpu::commStream() : s;
        if (overlap) {
            AMREX_CUDA_SAFE_CALL(cudaEventRecord(P.ev_packed, s));
            AMREX_CUDA_SAFE_CALL(cudaStreamWaitEvent(cs, P.ev_packed, 0));
        }
        if (Gpu::debugSync()) {
            static long long opno = 0;
            std::fprintf(stderr, "[comm %lld] rank %d ncomp %d op %d dst ba %llu (%ld boxes, ng %d, type %d%d%d) src ba %llu (%ld boxes): snd_total %lld rcv_total %lld",
                         opno++, ParallelDescriptor::MyProc(), ncomp, int(op), (unsigned long long)dst.boxArray().id(), long(dst.boxArray().size()), dst.nGrow(),
                         int(dst.ixType().test(0)), int(dst.ixType().test(1)), int(dst.ixType().test(2)),
                         (unsigned long long)src.boxArray().id(), long(src.boxArray().size()), P.snd_total, P.rcv_total);
            for (auto const& p : P.rcv_peers) { std::fprintf(stderr, " | recv from %d off %lld n %lld", p.rank, p.offset, p.count); }
            for (auto const& p : P.snd_peers) { std::fprintf(stderr, " | send to %d off %lld n %lld", p.rank, p.offset, p.count); }
            std::fprintf(stderr, "\n");
        }
        ncclGroupStart();
        for (auto const& p : P.rcv_peers) { ncclRecv(P.rcvbuf + p.offset * ncomp, p.count * ncomp, ncclDouble, p.rank, comm, cs); }
        for (auto const& p : P.snd_peers) { ncclSend(P.sndbuf + p.offset * ncomp, p.count * ncomp, ncclDouble, p.rank, comm, cs); }
        ncclGroupEnd();
        if (overlap) { AMREX_CUDA_SAFE_CALL(cudaEventRecord(P.ev_arrived, cs)); }
        if (Gpu::debugSync()) { Gpu::check(Gpu::debugSyncNow(), "[B200MG_DEBUG_SYNC] ncclSend/ncclRecv group", __FILE__, __LINE__); }
    }
    if (!remote_only) {
        B200_KCALL(b200mg_copy_tags_colour(P.nloc, P.d_loc.data(), dst.d_fabs(), src.d_fabs(), nullptr, ncomp, scomp, dcomp, int(op), P.maxloc, parity, s));
    }
}

void finish_plan_exchange (CommPlan& P, MultiFab& dst, int scomp, int dcomp, int ncomp, CpOp op, int parity = -1)
{
    cudaStream_t s = Gpu::gpuStream();
    const bool remote = (P.snd_total + P.rcv_total) > 0;
    static const bool overlap = std::getenv("B200MG_NO_COMM_OVERLAP") == nullptr;
    if (remote) {
        if (overlap) { AMREX_CUDA_SAFE_CALL(cudaStreamWaitEvent(s, P.ev_arrived, 0)); }
        B200_KCALL(b200mg_copy_tags_colour(P.nrcv, P.d_rcv.data(), dst.d_fabs(), nullptr, P.rcvbuf, ncomp, scomp, dcomp, int(op), P.maxrcv, parity, s));
    }
}

void execute_plan (CommPlan& P, MultiFab& dst, MultiFab const& src, int scomp, int dcomp, int ncomp, CpOp op, int parity = -1,
                   bool remote_only = false)
{
    start_plan(P, dst, src, scomp, dcomp, ncomp, op, parity, remote_only);
    finish_plan_exchange(P, dst, scomp, dcomp, ncomp, op, parity);
}

using FBKey = std::tuple<std::uint64_t, std::uint64_t, int, int, int, int, int, int, int>;
using FBCache = std::map<FBKey, std::unique_ptr<CommPlan>>;
FBCache& fb_cache () { static FBCache* m = new FBCache; return *m; }
using CPCKey = std::tuple<std::uint64_t, std::uint64_t, int, std::uint64_t, std::uint64_t, int, int, int, int>;
using CPCCache = std::map<CPCKey, std::unique_ptr<CommPlan>>;
CPCCache& cpc_cache () { static CPCCache* m = new CPCCache; return *m; }

// a BoxArray (kind 0) or DistributionMapping (kind 1) box list died: nothing can ask for its plans or layouts again
void evict_by_id (std::uint64_t id, int kind)
{
    for (auto it = layouts().begin(); it != layouts().end(); ) {
        if ((kind == 0 ? it->first.first : it->first.second) == id) { it = layouts().erase(it); } else { ++it; }
    }
    for (auto it = fb_cache().begin(); it != fb_cache().end(); ) {
        if ((kind == 0 ? std::get<0>(it->first) : std::get<1>(it->first)) == id) { it = fb_cache().erase(it); } else { ++it; }
    }
    for (auto it = cpc_cache().begin(); it != cpc_cache().end(); ) {
        const bool hit = (kind == 0) ? (std::get<0>(it->first) == id || std::get<3>(it->first) == id)
                                     : (std::get<1>(it->first) == id || std::get<4>(it->first) == id);
        if (hit) { it = cpc_cache().erase(it); } else { ++it; }
    }
}

} // namespace

void clear_comm_caches () { fb_cache().clear(); cpc_cache().clear(); }
std::size_t comm_cache_size () { return fb_cache().size() + cpc_cache().size() + layouts().size(); }

namespace {
CommPlan& fb_plan (MultiFab& mf, IntVect const& nghost, Periodicity const& period, bool cross)
{
    FBKey key{mf.boxArray().id(), mf.DistributionMap().id(), nghost[0], nghost[1], nghost[2], int(cross),
              period.intVect()[0], period.intVect()[1], period.intVect()[2]};
    auto it = fb_cache().find(key);
    if (it == fb_cache().end()) {
        auto P = std::make_unique<CommPlan>();
        define_fb_metadata(P->meta, mf.boxArray(), mf.DistributionMap(), nghost, cross, period, ParallelDescriptor::MyProc());
        finish_plan(*P, mf.layout(), mf.layout(), cross ? &nghost : nullptr);
        it = fb_cache().emplace(key, std::move(P)).first;
    }
    return *it->second;
}
}

void MultiFab::FillBoundary (int scomp, int ncomp, IntVect const& nghost, Periodicity const& period, bool cross, int parity,
                             bool remote_only)
{
    if (m_ngrow == 0 || nghost.max() == 0) { return; }
    AMREX_ALWAYS_ASSERT(nghost.allLE(IntVect(m_ngrow)));
    AMREX_ALWAYS_ASSERT_WITH_MESSAGE(m_fb_pending == nullptr, "FillBoundary while a FillBoundary_nowait is pending on this MultiFab");
    execute_plan(fb_plan(*this, nghost, period, cross), *this, *this, scomp, scomp, ncomp, CpOp::COPY, parity, remote_only);
}

void MultiFab::FillBoundary_nowait (int scomp, int ncomp, IntVect const& nghost, Periodicity const& period, bool cross, int parity,
                                    bool remote_only)
{
    if (m_ngrow == 0 || nghost.max() == 0) { return; }
    AMREX_ALWAYS_ASSERT(nghost.allLE(IntVect(m_ngrow)));
    AMREX_ALWAYS_ASSERT_WITH_MESSAGE(m_fb_pending == nullptr, "FillBoundary_nowait: the previous one was not finished");
    CommPlan& P = fb_plan(*this, nghost, period, cross);
    start_plan(P, *this, *this, scomp, scomp, ncomp, CpOp::COPY, parity, remote_only);
    m_fb_pending = &P; m_fb_scomp = scomp; m_fb_ncomp = ncomp; m_fb_parity = parity;
}

void MultiFab::FillBoundary_finish ()
{
    if (m_fb_pending == nullptr) { return; }
    finish_plan_exchange(*static_cast<CommPlan*>(m_fb_pending), *this, m_fb_scomp, m_fb_scomp, m_fb_ncomp, CpOp::COPY, m_fb_parity);
    m_fb_pending = nullptr;
}

std::vector<int> const& MultiFab::FillBoundaryRemoteFabs (IntVect const& nghost, Periodicity const& period, bool cross)
{
    return fb_plan(*this, nghost, period, cross).remote_fabs;
}

const b200mg_facelink* MultiFab::FillBoundaryFaceLinks (IntVect const& nghost, Periodicity const& period, bool cross)
{
    CommPlan& P = fb_plan(*this, nghost, period, cross);
    return P.links_ok ? P.d_links.data() : nullptr;
}

void MultiFab::ParallelCopy (MultiFab const& src, int scomp, int dcomp, int ncomp, int src_ng, int dst_ng,
                             Periodicity const& period, CpOp op)
{
    AMREX_ALWAYS_ASSERT(src_ng <= src.nGrow() && dst_ng <= m_ngrow && ixType() == src.ixType());
    // same layout, no ghost cells, non-periodic: plain local copy/add (AMReX_FabArrayCommI.H:336-355)
    if (m_ba == src.boxArray() && m_dm == src.DistributionMap() && src_ng == 0 && dst_ng == 0 && !period.isAnyPeriodic()) {
        if (op == CpOp::COPY) { Copy(*this, src, scomp, dcomp, ncomp, 0); } else { Add(*this, src, scomp, dcomp, ncomp, 0); }
        return;
    }
    CPCKey key{m_ba.id(), m_dm.id(), dst_ng, src.boxArray().id(), src.DistributionMap().id(), src_ng,
               period.intVect()[0], period.intVect()[1], period.intVect()[2]};
    auto it = cpc_cache().find(key);
    if (it == cpc_cache().end()) {
        auto P = std::make_unique<CommPlan>();
        define_cpc_metadata(P->meta, m_ba, m_dm, IntVect(dst_ng), src.boxArray(), src.DistributionMap(), IntVect(src_ng),
                            period, false, ParallelDescriptor::MyProc());
        finish_plan(*P, layout(), src.layout());
        it = cpc_cache().emplace(key, std::move(P)).first;
    }
    execute_plan(*it->second, *this, src, scomp, dcomp, ncomp, op);
}

// ============================================================================================= utilities
void average_down (MultiFab const& fine, MultiFab& crse, int scomp, int ncomp, int ratio)
{
    AMREX_ALWAYS_ASSERT(scomp == 0 && ncomp == 1);
    BoxArray cfba = amrex::coarsen(fine.boxArray(), ratio);
    if (cfba == crse.boxArray() && fine.DistributionMap() == crse.DistributionMap()) {
        auto const& T = crse.layout().tiles(0);
        B200_KCALL(b200mg_restrict_cc(T.n, T.d.data(), crse.layout().d_vbox(), crse.d_fabs(), fine.d_fabs(), ratio, Gpu::gpuStream()));
    } else {   // AMReX_MultiFabUtil.H:582-650: coarsen onto a temporary on the fine layout, then redistribute
        MultiFab tmp(cfba, fine.DistributionMap(), 1, 0);
        auto const& T = tmp.layout().tiles(0);
        B200_KCALL(b200mg_restrict_cc(T.n, T.d.data(), tmp.layout().d_vbox(), tmp.d_fabs(), fine.d_fabs(), ratio, Gpu::gpuStream()));
        crse.ParallelCopy(tmp, 0, 0, 1);
    }
}

void average_down_faces (MultiFab const& fine, MultiFab& crse, int dir, int ratio)
{
    BoxArray cfba = amrex::coarsen(fine.boxArray(), ratio);
    if (cfba == crse.boxArray() && fine.DistributionMap() == crse.DistributionMap()) {
        auto const& T = crse.layout().tiles(0);
        B200_KCALL(b200mg_restrict_faces(T.n, T.d.data(), crse.layout().d_vbox(), crse.d_fabs(), fine.d_fabs(), dir, ratio, Gpu::gpuStream()));
    } else {
        MultiFab tmp(cfba, fine.DistributionMap(), 1, 0);
        auto const& T = tmp.layout().tiles(0);
        B200_KCALL(b200mg_restrict_faces(T.n, T.d.data(), tmp.layout().d_vbox(), tmp.d_fabs(), fine.d_fabs(), dir, ratio, Gpu::gpuStream()));
        crse.ParallelCopy(tmp, 0, 0, 1);
    }
}

void average_cellcenter_to_face (Array<MultiFab*, 3> const& fc, MultiFab const& cc, Geometry const&)
{
    AMREX_ALWAYS_ASSERT(cc.nGrow() >= 1 && cc.nComp() == 1);
    for (int d = 0; d < 3; ++d) {
        auto const& T = fc[d]->layout().tiles(0);
        B200_KCALL(b200mg_cc_to_face(T.n, T.d.data(), fc[d]->layout().d_vbox(), fc[d]->d_fabs(), cc.d_fabs(), d, Gpu::gpuStream()));
    }
}

iMultiFab makeFineMask (BoxArray const& cba, DistributionMapping const& cdm, BoxArray const& fba, int ratio,
                        int crse_value, int fine_value, Periodicity const& period)
{
    iMultiFab mask(cba, cdm, 1, 0);
    const BoxArray cfba = amrex::coarsen(fba, ratio);
    const auto pshifts = period.shiftIntVect();
    std::vector<std::pair<int, Box>> isects;
    for (int li = 0; li < mask.local_size(); ++li) {
        const Box bx = mask.validbox(li);
        std::vector<int> h(bx.numPts(), crse_value);
        for (auto const& iv : pshifts) {
            cfba.intersections(bx + iv, isects);
            for (auto const& is : isects) {
                const Box b = is.second - iv;
                for (int k = b.smallEnd(2); k <= b.bigEnd(2); ++k) for (int j = b.smallEnd(1); j <= b.bigEnd(1); ++j)
                    for (int i = b.smallEnd(0); i <= b.bigEnd(0); ++i) {
                        h[(i - bx.smallEnd(0)) + Long(bx.length(0)) * ((j - bx.smallEnd(1)) + Long(bx.length(1)) * (k - bx.smallEnd(2)))] = fine_value;
                    }
            }
        }
        mask.copyFromHost(h.data(), bx, 0, 0);
    }
    return mask;
}

} // namespace amrex
