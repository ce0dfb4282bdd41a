// Plotfile / VisMF output - see AMReX_PlotFileUtil.H for the on-disk layout and the reference entry points.
#include "AMReX_PlotFileUtil.H"

#include <algorithm>
#include <cerrno>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <iomanip>
#include <limits>
#include <sstream>
#include <sys/stat.h>

namespace amrex {

namespace {

// ((lo) (hi) (type)) - the reference's text form of a Box (operator<<, Src/Base/AMReX_Box.cpp:20-40)
std::string box_text (Box const& b)
{
    const IntVect t = b.ixType().ixType();
    std::ostringstream s;
    s << "((" << b.smallEnd(0) << ',' << b.smallEnd(1) << ',' << b.smallEnd(2) << ") ("
      << b.bigEnd(0) << ',' << b.bigEnd(1) << ',' << b.bigEnd(2) << ") ("
      << t[0] << ',' << t[1] << ',' << t[2] << "))";
    return s.str();
}

void make_dir (std::string const& path)
{
    std::string cur;
    for (std::size_t i = 0; i <= path.size(); ++i) {
        if (i == path.size() || path[i] == '/') {
            if (!cur.empty() && cur != "." && ::mkdir(cur.c_str(), 0755) != 0 && errno != EEXIST) {
                Abort("WritePlotfile: cannot create directory " + cur + ": " + std::strerror(errno));
            }
        }
        if (i < path.size()) { cur.push_back(path[i]); }
    }
}

std::string base_name (std::string const& p) { const auto k = p.rfind('/'); return k == std::string::npos ? p : p.substr(k + 1); }

std::string rank_file (std::string const& name, int rank)
{
    char buf[16]; std::snprintf(buf, sizeof(buf), "_D_%05d", rank);
    return name + buf;
}

// values every rank contributes for its own fabs and all ranks need: sum-reduce of arrays that are zero elsewhere
void share (std::vector<double>& v)
{
    if (ParallelDescriptor::NProcs() == 1) { return; }
    constexpr int chunk = 32;
    for (std::size_t i = 0; i < v.size(); i += chunk) {
        ParallelDescriptor::ReduceRealSum(v.data() + i, int(std::min<std::size_t>(chunk, v.size() - i)));
    }
}

} // namespace

// VisMF::Write (Src/Base/AMReX_VisMF.cpp:958-1190) with header version 1 and one data file per rank
Long VisMF::Write (MultiFab const& mf, std::string const& name)
{
    const int ncomp = mf.nComp(), ng = mf.nGrow();
    const int nboxes = int(mf.boxArray().size());
    const int me = ParallelDescriptor::MyProc();
    auto const& idx = mf.layout().indexArray();

    // [offset | min(ncomp) | max(ncomp)] per global box, filled by the owner
    const int rec = 1 + 2 * ncomp;
    std::vector<double> info(std::size_t(nboxes) * rec, 0.0);
    Long bytes = 0;
    {
        std::ofstream data;
        if (!idx.empty()) {
            data.open(rank_file(name, me), std::ios::binary | std::ios::trunc);
            if (!data.good()) { Abort("VisMF::Write: cannot open " + rank_file(name, me)); }
        }
        std::vector<double> buf;
        for (int li = 0; li < int(idx.size()); ++li) {
            const Box g = amrex::grow(mf.validbox(li), ng);
            const std::size_t npts = std::size_t(g.numPts());
            buf.resize(npts * ncomp);
            for (int c = 0; c < ncomp; ++c) { mf.copyFabToHost(li, buf.data() + c * npts, c, ng); }
            double* r = info.data() + std::size_t(idx[li]) * rec;
            r[0] = double(bytes);
            for (int c = 0; c < ncomp; ++c) {
                auto mm = std::minmax_element(buf.begin() + c * npts, buf.begin() + (c + 1) * npts);
                r[1 + c] = *mm.first; r[1 + ncomp + c] = *mm.second;
            }
            // FAB header: native little-endian IEEE double descriptor (FPC::NativeRealDescriptor), box, component count
            std::ostringstream h;
            h << "FAB ((8, (64 11 52 0 1 12 0 1023)),(8, (8 7 6 5 4 3 2 1)))" << box_text(g) << ' ' << ncomp << '\n';
            const std::string hs = h.str();
            data.write(hs.data(), std::streamsize(hs.size()));
            data.write(reinterpret_cast<const char*>(buf.data()), std::streamsize(buf.size() * sizeof(double)));
            bytes += Long(hs.size()) + Long(buf.size() * sizeof(double));
        }
        if (data.is_open()) { data.close(); if (!data.good()) { Abort("VisMF::Write: write to " + rank_file(name, me) + " failed"); } }
    }
    share(info);

    if (ParallelDescriptor::IOProcessor()) {
        std::ofstream h(name + "_H", std::ios::trunc);
        if (!h.good()) { Abort("VisMF::Write: cannot open " + name + "_H"); }
        h << "1\n" << "1\n" << ncomp << '\n' << ng << '\n';            // version 1, NFiles, ncomp, ngrow
        h << '(' << nboxes << " 0\n";
        for (int i = 0; i < nboxes; ++i) { h << box_text(mf.boxArray()[i]) << '\n'; }
        h << ")\n";
        h << nboxes << '\n';
        const std::string bn = base_name(name);
        for (int i = 0; i < nboxes; ++i) {
            h << "FabOnDisk: " << rank_file(bn, mf.DistributionMap()[i]) << ' ' << (long long)(info[std::size_t(i) * rec]) << '\n';
        }
        h << '\n';
        h << std::scientific << std::setprecision(17);
        for (int which = 0; which < 2; ++which) {
            h << nboxes << ',' << ncomp << '\n';
            for (int i = 0; i < nboxes; ++i) {
                for (int c = 0; c < ncomp; ++c) { h << info[std::size_t(i) * rec + 1 + which * ncomp + c] << ','; }
                h << '\n';
            }
            h << '\n';
        }
        h.close();
        if (!h.good()) { Abort("VisMF::Write: write to " + name + "_H failed"); }
    }
    ParallelDescriptor::Barrier();
    return bytes;
}

// WriteGenericPlotfileHeader + WriteMultiLevelPlotfile (Src/Base/AMReX_PlotFileUtil.cpp:73-270)
void WriteMultiLevelPlotfile (std::string const& plotfilename, int nlevels, Vector<const MultiFab*> const& mf,
                              Vector<std::string> const& varnames, Vector<Geometry> const& geom, Real time,
                              Vector<int> const& level_steps, Vector<IntVect> const& ref_ratio,
                              std::string const& versionName, std::string const& levelPrefix, std::string const& mfPrefix)
{
    AMREX_ALWAYS_ASSERT(nlevels >= 1 && nlevels <= int(mf.size()) && nlevels <= int(geom.size()) && nlevels <= int(level_steps.size())
                        && nlevels <= int(ref_ratio.size()) + 1);
    AMREX_ALWAYS_ASSERT(mf[0]->nComp() == int(varnames.size()));
    const int finest_level = nlevels - 1;
    auto level_dir = [&] (int l) { return levelPrefix + std::to_string(l); };

    if (ParallelDescriptor::IOProcessor()) {
        for (int l = 0; l <= finest_level; ++l) { make_dir(plotfilename + "/" + level_dir(l)); }
        std::ofstream h(plotfilename + "/Header", std::ios::trunc);
        if (!h.good()) { Abort("WriteMultiLevelPlotfile: cannot open " + plotfilename + "/Header"); }
        h << std::setprecision(17);
        h << versionName << '\n' << varnames.size() << '\n';
        for (auto const& v : varnames) { h << v << '\n'; }
        h << 3 << '\n' << time << '\n' << finest_level << '\n';
        for (int d = 0; d < 3; ++d) { h << geom[0].ProbLo()[d] << ' '; }
        h << '\n';
        for (int d = 0; d < 3; ++d) { h << geom[0].ProbHi()[d] << ' '; }
        h << '\n';
        for (int l = 0; l < finest_level; ++l) { h << ref_ratio[l][0] << ' '; }
        h << '\n';
        for (int l = 0; l <= finest_level; ++l) { h << box_text(geom[l].Domain()) << ' '; }
        h << '\n';
        for (int l = 0; l <= finest_level; ++l) { h << level_steps[l] << ' '; }
        h << '\n';
        for (int l = 0; l <= finest_level; ++l) {
            for (int d = 0; d < 3; ++d) { h << geom[l].CellSize()[d] << ' '; }
            h << '\n';
        }
        h << geom[0].Coord() << '\n' << "0\n";
        for (int l = 0; l <= finest_level; ++l) {
            BoxArray const& ba = mf[l]->boxArray();
            h << l << ' ' << ba.size() << ' ' << time << '\n' << level_steps[l] << '\n';
            const IntVect dlo = geom[l].Domain().smallEnd();
            const Real* dx = geom[l].CellSize(); const Real* plo = geom[l].ProbLo();
            for (int i = 0, N = int(ba.size()); i < N; ++i) {
                Box const& b = ba[i];
                for (int d = 0; d < 3; ++d) {
                    // physical extent of the box (RealBox(Box, dx, base), Src/Base/AMReX_RealBox.cpp:9-20)
                    h << plo[d] + dx[d] * (b.smallEnd(d) - dlo[d]) << ' ' << plo[d] + dx[d] * (b.bigEnd(d) - dlo[d] + 1) << '\n';
                }
            }
            h << level_dir(l) << '/' << mfPrefix << '\n';
        }
        h.close();
        if (!h.good()) { Abort("WriteMultiLevelPlotfile: write of the Header failed"); }
    }
    ParallelDescriptor::Barrier();     // directories exist before any rank opens its data file
    for (int l = 0; l <= finest_level; ++l) {
        if (mf[l]->nGrow() > 0) {
            // plotfiles carry valid cells only (AMReX_PlotFileUtil.cpp:225-240 strips the ghost cells through a copy)
            MultiFab tmp(mf[l]->boxArray(), mf[l]->DistributionMap(), mf[l]->nComp(), 0);
            MultiFab::Copy(tmp, *mf[l], 0, 0, mf[l]->nComp(), 0);
            VisMF::Write(tmp, plotfilename + "/" + level_dir(l) + "/" + mfPrefix);
        } else {
            VisMF::Write(*mf[l], plotfilename + "/" + level_dir(l) + "/" + mfPrefix);
        }
    }
}

void WriteSingleLevelPlotfile (std::string const& plotfilename, MultiFab const& mf, Vector<std::string> const& varnames,
                               Geometry const& geom, Real time, int level_step,
                               std::string const& versionName, std::string const& levelPrefix, std::string const& mfPrefix)
{
    WriteMultiLevelPlotfile(plotfilename, 1, {&mf}, varnames, {geom}, time, {level_step}, {IntVect(1)}, versionName, levelPrefix, mfPrefix);
}

} // namespace amrex
