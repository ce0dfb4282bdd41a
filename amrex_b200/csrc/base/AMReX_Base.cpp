// Host-side index algebra: see AMReX_Base.H for the reference citations of each algorithm.
#include "AMReX_Base.H"
#include "AMReX_Parallel.H"

#include <atomic>
#include <iostream>

namespace amrex {

static bool s_throw_on_abort = true;

void Abort (const std::string& msg)
{
    std::fprintf(stderr, "amrex::Abort::%d::%s !!!\n", ParallelDescriptor::MyProc(), msg.c_str());
    std::fflush(stderr);
    if (s_throw_on_abort) { throw std::runtime_error(msg); }
    std::abort();
}

void Assert_fail (const char* expr, const char* file, int line, const char* msg)
{
    std::string s = std::string("Assertion `") + expr + "' failed, file \"" + file + "\", line " + std::to_string(line);
    if (msg) { s += std::string(", Msg: ") + msg; }
    Abort(s);
}

Box adjCell (Box const& b, Orientation f, int len) noexcept
{
    IntVect lo = b.smallEnd(), hi = b.bigEnd();
    const int d = f.coordDir();
    if (f.isLow()) { const int sm = lo[d]; lo[d] = sm - len; hi[d] = sm - 1; }
    else { const int bg = hi[d] + 1 - (b.type()[d] % 2); lo[d] = bg; hi[d] = bg + len - 1; }
    IndexType t = b.ixType(); t.unset(d);
    return Box(lo, hi, t);
}

Box insideCell (Box const& b, Orientation f, int len) noexcept
{
    IntVect lo = b.smallEnd(), hi = b.bigEnd();
    const int d = f.coordDir();
    if (f.isLow()) { hi[d] = lo[d] + len - 1; } else { lo[d] = hi[d] - len + 1; }
    return Box(lo, hi, b.ixType());
}

// ------------------------------------------------------------------------------------- Periodicity
Box Periodicity::Domain () const noexcept
{
    Box pd;
    IntVect lo, hi;
    for (int d = 0; d < 3; ++d) {
        if (isPeriodic(d)) { lo[d] = 0; hi[d] = period[d] - 1; }
        else { lo[d] = std::numeric_limits<int>::min(); hi[d] = std::numeric_limits<int>::max() - 1; }
    }
    return Box(lo, hi);
}

std::vector<IntVect> Periodicity::shiftIntVect (IntVect const& nghost) const
{
    std::vector<IntVect> r;
    int per[3] = {0, 0, 0}, jmp[3] = {1, 1, 1};
    for (int d = 0; d < 3; ++d) {
        if (isPeriodic(d)) {
            per[d] = jmp[d] = period[d];
            while (per[d] < nghost[d]) { per[d] += period[d]; }
        }
    }
    for (int i = -per[0]; i <= per[0]; i += jmp[0])
        for (int j = -per[1]; j <= per[1]; j += jmp[1])
            for (int k = -per[2]; k <= per[2]; k += jmp[2])
                r.emplace_back(i, j, k);
    return r;
}

// ---------------------------------------------------------------------------------------- Geometry
namespace {
    RealBox g_default_rb({0., 0., 0.}, {1., 1., 1.});
    int g_default_coord = 0;
    Array<int, 3> g_default_per{{0, 0, 0}};
}

void Geometry::Setup (const RealBox* rb, int coord, int const* is_per)
{
    if (rb) { g_default_rb = *rb; }
    if (coord >= 0) { g_default_coord = coord; }
    if (is_per) { for (int d = 0; d < 3; ++d) { g_default_per[d] = is_per[d]; } }
}

void Geometry::define (Box const& dom) { define(dom, g_default_rb, g_default_coord, g_default_per); }

void Geometry::define (Box const& dom, RealBox const& rb, int a_coord, Array<int, 3> const& is_per)
{
    domain = dom; prob_domain = rb; coord = a_coord; is_periodic = is_per;
    for (int d = 0; d < 3; ++d) {   // AMReX_Geometry.cpp:520-521
        dx[d] = prob_domain.length(d) / static_cast<Real>(domain.length(d));
        inv_dx[d] = 1.0 / dx[d];
    }
}

// ----------------------------------------------------------------------------------------- BoxList
BoxList::BoxList (Box const& bx, IntVect const& tilesize) : btype(bx.ixType())
{
    IntVect nt;
    for (int d = 0; d < 3; ++d) { nt[d] = (bx.length(d) + tilesize[d] - 1) / tilesize[d]; }
    for (int k = 0; k < nt[2]; ++k) for (int j = 0; j < nt[1]; ++j) for (int i = 0; i < nt[0]; ++i) {
        IntVect ijk(i, j, k), sm, bg;
        for (int d = 0; d < 3; ++d) {
            sm[d] = ijk[d] * tilesize[d];
            bg[d] = std::min(sm[d] + tilesize[d] - 1, bx.length(d) - 1);
        }
        Box t(sm, bg, btype); t.shift(bx.smallEnd()); m_lbox.push_back(t);
    }
}

BoxList& BoxList::maxSize (IntVect const& chunk)
{
    std::vector<Box> out;
    for (auto const& bx : m_lbox) {
        const IntVect len = amrex::enclosedCells(bx).size();
        const IntVect lo = bx.smallEnd();
        IntVect ratio(1), numblk(1), extra(0), sz = len;
        for (int d = 0; d < 3; ++d) {
            if (len[d] > chunk[d]) {
                int bs = chunk[d], nlen = len[d];
                while ((bs % 2 == 0) && (nlen % 2 == 0)) { ratio[d] *= 2; bs /= 2; nlen /= 2; }
                numblk[d] = (nlen + bs - 1) / bs;
                sz[d] = nlen / numblk[d];
                extra[d] = nlen - sz[d] * numblk[d];
            }
        }
        if (numblk == 1) { out.push_back(bx); continue; }
        auto span = [&] (int d, int n, int& a, int& b) {
            a = (n < extra[d]) ? n * (sz[d] + 1) * ratio[d] : (n * sz[d] + extra[d]) * ratio[d];
            b = (n < extra[d]) ? a + (sz[d] + 1) * ratio[d] - 1 : a + sz[d] * ratio[d] - 1;
            a += lo[d]; b += lo[d];
        };
        for (int k = 0; k < numblk[2]; ++k) { int klo, khi; span(2, k, klo, khi);
            for (int j = 0; j < numblk[1]; ++j) { int jlo, jhi; span(1, j, jlo, jhi);
                for (int i = 0; i < numblk[0]; ++i) { int ilo, ihi; span(0, i, ilo, ihi);
                    out.push_back(Box(IntVect(ilo, jlo, klo), IntVect(ihi, jhi, khi)).convert(ixType()));
                } } }
    }
    m_lbox.swap(out);
    return *this;
}

Box BoxList::minimalBox () const
{
    Box mb;
    if (!m_lbox.empty()) { mb = m_lbox[0]; for (auto const& b : m_lbox) { mb.minBox(b); } }
    return mb;
}

void boxDiff (BoxList& out, Box const& b1in, Box const& b2)
{
    out.clear(); out.set(b2.ixType());
    if (b2.contains(b1in)) { return; }
    Box b1(b1in);
    if (!b1.intersects(b2)) { out.push_back(b1); return; }
    for (int d = 2; d >= 0; --d) {
        if (b1.smallEnd(d) < b2.smallEnd(d) && b2.smallEnd(d) <= b1.bigEnd(d)) {
            Box bn(b1); bn.setBig(d, b2.smallEnd(d) - 1); out.push_back(bn); b1.setSmall(d, b2.smallEnd(d));
        }
        if (b1.smallEnd(d) <= b2.bigEnd(d) && b2.bigEnd(d) < b1.bigEnd(d)) {
            Box bn(b1); bn.setSmall(d, b2.bigEnd(d) + 1); out.push_back(bn); b1.setBig(d, b2.bigEnd(d));
        }
    }
}

BoxList boxDiff (Box const& b1, Box const& b2) { BoxList bl(b1.ixType()); boxDiff(bl, b1, b2); return bl; }

// ---------------------------------------------------------------------------------------- BoxArray
static std::atomic<RefDeathHook> s_ref_death_hook{nullptr};
void setRefDeathHook (RefDeathHook h) noexcept { s_ref_death_hook = h; }
void notifyRefDeath (std::uint64_t id, int kind) noexcept { if (RefDeathHook h = s_ref_death_hook.load()) { h(id, kind); } }

static std::atomic<std::uint64_t> s_next_id{1};

void BoxArray::newId () { m_ref->id = s_next_id++; }

void BoxArray::uniqify ()
{
    auto p = std::make_shared<Ref>();
    p->boxes = m_ref->boxes; p->ixtype = m_ref->ixtype;
    m_ref = std::move(p); newId();
}

Long BoxArray::numPts () const noexcept { Long n = 0; for (auto const& b : m_ref->boxes) { n += b.numPts(); } return n; }
double BoxArray::d_numPts () const noexcept { double n = 0; for (auto const& b : m_ref->boxes) { n += b.d_numPts(); } return n; }

Box BoxArray::minimalBox () const
{
    Box mb;
    if (!empty()) { mb = m_ref->boxes[0]; for (auto const& b : m_ref->boxes) { mb.minBox(b); } }
    return mb;
}

BoxArray& BoxArray::maxSize (IntVect const& chunk)
{
    BoxList bl = boxList(); bl.maxSize(chunk);
    if (bl.size() != m_ref->boxes.size()) { uniqify(); m_ref->boxes = bl.data(); }
    return *this;
}

BoxArray& BoxArray::coarsen (IntVect const& r) { if (r != 1) { uniqify(); for (auto& b : m_ref->boxes) { b.coarsen(r); } } return *this; }
BoxArray& BoxArray::refine (IntVect const& r) { if (r != 1) { uniqify(); for (auto& b : m_ref->boxes) { b.refine(r); } } return *this; }
BoxArray& BoxArray::convert (IndexType t)
{
    if (t != m_ref->ixtype) { uniqify(); for (auto& b : m_ref->boxes) { b.convert(t); } m_ref->ixtype = t; }
    return *this;
}
BoxArray& BoxArray::grow (int n) { uniqify(); for (auto& b : m_ref->boxes) { b.grow(n); } return *this; }

bool BoxArray::coarsenable (IntVect const& r, IntVect const& mw) const
{
    if (empty()) { return false; }
    for (auto const& b : m_ref->boxes) { if (!b.coarsenable(r, mw)) { return false; } }
    return true;
}

void BoxArray::buildHash () const
{
    Ref& R = *m_ref;
    if (R.has_hash) { return; }
    IntVect maxext(1);
    Box bb = amrex::enclosedCells(R.boxes[0]);
    for (auto const& b0 : R.boxes) {
        Box b = amrex::enclosedCells(b0); b.normalize();
        maxext = amrex::max(maxext, b.size()); bb.minBox(b);
    }
    for (int i = 0, N = int(R.boxes.size()); i < N; ++i) {
        R.hash[key(amrex::coarsen(R.boxes[i].smallEnd(), maxext))].push_back(i);
    }
    R.crsn = maxext;
    R.bbox = bb.coarsen(maxext); R.bbox.normalize();
    R.has_hash = true;
}

void BoxArray::intersections (Box const& bx, std::vector<std::pair<int, Box>>& isects, bool first_only, IntVect const& ng) const
{
    isects.clear();
    if (empty()) { return; }
    buildHash();
    Ref const& R = *m_ref;
    Box gbx = amrex::grow(bx, ng);
    // a nodal query box can touch a cell-keyed bin one further out (role of doiHi, AMReX_BoxArray.H:163)
    gbx.setSmall(gbx.smallEnd() - R.ixtype.ixType());
    // coarsened as a box of the array's own index type (AMReX_BoxArray.cpp:1240): a nodal upper bound that sits exactly on a
    // bin boundary must reach the bin above -- a one-node-thick face plane shared by two boxes is found from both sides
    gbx.coarsen(R.crsn);
    IntVect sm = amrex::max(gbx.smallEnd() - 1, R.bbox.smallEnd());
    IntVect bg = amrex::min(gbx.bigEnd(), R.bbox.bigEnd());
    Box cbx(sm, bg);
    if (!cbx.ok()) { return; }
    for (int k = sm[2]; k <= bg[2]; ++k) for (int j = sm[1]; j <= bg[1]; ++j) for (int i = sm[0]; i <= bg[0]; ++i) {
        auto it = R.hash.find(key(IntVect(i, j, k)));
        if (it == R.hash.end()) { continue; }
        for (int idx : it->second) {
            Box isect = bx & amrex::grow(R.boxes[idx], ng);
            if (isect.ok()) { isects.emplace_back(idx, isect); if (first_only) { return; } }
        }
    }
}

BoxList BoxArray::complementIn (Box const& bx) const
{
    BoxList bl(bx.ixType()); bl.push_back(bx);
    if (empty()) { return bl; }
    std::vector<std::pair<int, Box>> isects;
    intersections(bx, isects);
    BoxList newbl(bx.ixType()), diff(bx.ixType());
    for (auto const& is : isects) {
        Box const& ibox = m_ref->boxes[is.first];
        newbl.clear();
        for (Box const& b : bl) { boxDiff(diff, b, ibox); newbl.join(diff); }
        bl.swap(newbl);
        if (bl.isEmpty()) { break; }
    }
    return bl;
}

bool BoxArray::contains (Box const& bx) const { return !empty() && complementIn(bx).isEmpty(); }
bool BoxArray::contains (BoxArray const& ba) const
{
    if (empty() || ba.empty()) { return false; }
    for (auto const& b : ba.boxes()) { if (!contains(b)) { return false; } }
    return true;
}

bool BoxArray::isDisjoint () const
{
    std::vector<std::pair<int, Box>> isects;
    for (int i = 0, N = int(size()); i < N; ++i) {
        intersections(m_ref->boxes[i], isects);
        if (isects.size() > 1) { return false; }
    }
    return true;
}

// ----------------------------------------------------------------------------- DistributionMapping
namespace {
struct SFCToken { int box; std::uint32_t m[3]; };

inline std::uint32_t make_space (std::uint32_t x) noexcept   // spread the low 10 bits, 2 zero bits between
{
    x = (x | (x << 16)) & 0x030000FFu;
    x = (x | (x << 8)) & 0x0300F00Fu;
    x = (x | (x << 4)) & 0x030C30C3u;
    x = (x | (x << 2)) & 0x09249249u;
    return x;
}

SFCToken make_token (int idx, IntVect const& iv)
{
    SFCToken t; t.box = idx;
    constexpr int imin = -(1 << 29);
    std::uint32_t x = std::uint32_t(iv[0] - imin), y = std::uint32_t(iv[1] - imin), z = std::uint32_t(iv[2] - imin);
    for (int g = 0; g < 3; ++g) {
        t.m[g] = make_space(x & 0x3FF) | (make_space(y & 0x3FF) << 1) | (make_space(z & 0x3FF) << 2);
        x >>= 10; y >>= 10; z >>= 10;
    }
    return t;
}

bool token_less (SFCToken const& l, SFCToken const& r)
{
    return (l.m[2] < r.m[2]) || ((l.m[2] == r.m[2]) && ((l.m[1] < r.m[1]) || ((l.m[1] == r.m[1]) && (l.m[0] < r.m[0]))));
}

void distribute (std::vector<SFCToken> const& tokens, std::vector<Long> const& wgts, int nprocs, Real volpercpu,
                 std::vector<std::vector<int>>& v)
{
    int K = 0; Real totalvol = 0;
    const int TSZ = int(tokens.size());
    for (int i = 0; i < nprocs; ++i) {
        int cnt = 0; Real vol = 0;
        for (; K < TSZ && (i == (nprocs - 1) || (vol < volpercpu)); ++K) {
            vol += Real(wgts[tokens[K].box]); ++cnt; v[i].push_back(tokens[K].box);
        }
        totalvol += vol;
        if ((totalvol / Real(i + 1)) > volpercpu && cnt > 1 && i < nprocs - 1) {
            --K; v[i].pop_back(); totalvol -= Real(wgts[tokens[K].box]);
        }
    }
}
} // namespace

std::vector<std::vector<int>> DistributionMapping::makeSFC (BoxArray const& ba, bool use_box_vol, int nprocs)
{
    if (nprocs < 0) { nprocs = ParallelDescriptor::NProcs(); }
    const int N = int(ba.size());
    std::vector<SFCToken> tokens; tokens.reserve(N);
    std::vector<Long> wgts; wgts.reserve(N);
    Long vol_sum = 0;
    for (int i = 0; i < N; ++i) {
        tokens.push_back(make_token(i, ba[i].smallEnd()));
        const Long v = use_box_vol ? ba[i].numPts() : Long(1);
        vol_sum += v; wgts.push_back(v);
    }
    std::sort(tokens.begin(), tokens.end(), token_less);
    Real volper = Real(vol_sum) / Real(nprocs);
    std::vector<std::vector<int>> r(nprocs);
    distribute(tokens, wgts, nprocs, volper, r);
    return r;
}

static std::atomic<std::uint64_t> s_next_dm_id{1};

void DistributionMapping::define (BoxArray const& ba) { define(ba, ParallelDescriptor::NProcs()); }

void DistributionMapping::define (BoxArray const& ba, int nprocs)
{
    // SFCProcessorMapDoIt (AMReX_DistributionMapping.cpp:1262-1432): SFC buckets, heaviest bucket first
    // (stable), bucket i -> rank ord[i]; ord is the identity here (see class comment).
    const int N = int(ba.size());
    Vector<int> pmap(N, 0);
    if (nprocs > 1 && N > 0) {
        auto buckets = makeSFC(ba, true, nprocs);
        std::vector<std::pair<Long, int>> wi;
        for (int i = 0; i < nprocs; ++i) { Long w = 0; for (int b : buckets[i]) { w += ba[b].numPts(); } wi.emplace_back(w, i); }
        std::stable_sort(wi.begin(), wi.end(), [] (auto const& a, auto const& b) { return a.first > b.first; });
        for (int i = 0; i < nprocs; ++i) { for (int b : buckets[wi[i].second]) { pmap[b] = i; } }
    }
    define(std::move(pmap));
}

void DistributionMapping::define (Vector<int> pmap)
{
    m_ref = std::make_shared<Ref>();
    m_ref->pmap = std::move(pmap);
    m_ref->id = s_next_dm_id++;
}

} // namespace amrex
