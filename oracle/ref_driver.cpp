// oracle/ref_driver.cpp -- TEST INFRASTRUCTURE ONLY (never linked into or called by amrex_b200/).
//
// A small driver linked against the UNMODIFIED reference (oracle/_ref/libamrex_ref.a, built from
// /root/reference by oracle/Makefile).  It sets up the problems of
// Tests/LinearSolvers/ABecLaplacian_C (prob_type 1: MyTest.cpp:51-148 + initProb_K.H:7-38;
// prob_type 2: MyTest.cpp:150-299 + initProb_K.H:74-140) plus a fully periodic Poisson case
// (prob_type 5, BASELINE config #5, no such test in the tree), runs the reference's own
// MLMG/MLLinOp on them and writes
//   mode=solve : "RESULT {json}" (iterations, residual history, norms, timers) and, with dump_dir=...,
//                raw little-endian fp64 arrays (inputs AND the reference's solution) so the CUDA path
//                can be fed bit-identical inputs;
//   mode=meta  : MG hierarchy (boxes + dmap per amr x mg level), FillBoundary LocTags, SFC maps;
//   mode=prim  : outputs of single primitives (smooth, apply, restriction, interpolation, applyBC).
// Arguments are key=value pairs (ParmParse).
#include <AMReX.H>
#include <AMReX_ParmParse.H>
#include <AMReX_MultiFab.H>
#include <AMReX_MultiFabUtil.H>
#include <AMReX_MLMG.H>
#include <AMReX_MLABecLaplacian.H>
#include <AMReX_MLPoisson.H>
#include <AMReX_MLALaplacian.H>
#include <AMReX_GMRES_MLMG.H>
#include <AMReX_PlotFileUtil.H>
#include <AMReX_Print.H>

#include <chrono>
#include <cmath>
#include <cstdio>
#include <fstream>
#include <sstream>
#include <string>

using namespace amrex;

namespace {

struct Params {
    std::string mode = "solve";
    int prob_type = 1, n_cell = 128, max_grid_size = 64, max_level = 0, ref_ratio = 2;
    int maxorder = 2, agg_grid_size = 32, max_iter = 100, max_fmg_iter = 0, verbose = 0, bottom_verbose = 0;
    int nsolve = 1, max_coarsening_level = 30, agglomeration = 1, consolidation = 1;
    int nprocs = 1;           // meta: rank count for the SFC map dump
    int prim_mglev = 0;       // prim: MG level on which primitives run
    int use_gmres = 0, gmres_precond = 1, gmres_precond_iters = 1;   // solve: GMRESMLMG instead of MLMG::solve
    int gauss_seidel = 1;     // 0: damped Jacobi smoother (MLCellLinOp::setGaussSeidel(false))
    int composite_solve = 1;  // 0: level by level, each fine level with setCoarseFineBC data from the level below (MyTest.cpp:104-141)
    std::string bottom = "default";
    std::string dump_dir;     // empty: no dump
    std::string plotfile;     // solve: write solution / rhs / exact_solution / error as the reference's test driver does
    Real tol_rel = 1.e-10, tol_abs = 0.0, ascalar = 1.e-3, bscalar = 1.0;
};

Params read_params ()
{
    Params p; ParmParse pp;
    pp.query("mode", p.mode); pp.query("prob_type", p.prob_type); pp.query("n_cell", p.n_cell);
    pp.query("max_grid_size", p.max_grid_size); pp.query("max_level", p.max_level);
    pp.query("linop_maxorder", p.maxorder); pp.query("agg_grid_size", p.agg_grid_size);
    pp.query("max_iter", p.max_iter); pp.query("max_fmg_iter", p.max_fmg_iter);
    pp.query("verbose", p.verbose); pp.query("bottom_verbose", p.bottom_verbose);
    pp.query("nsolve", p.nsolve); pp.query("max_coarsening_level", p.max_coarsening_level);
    pp.query("agglomeration", p.agglomeration); pp.query("consolidation", p.consolidation);
    pp.query("nprocs", p.nprocs); pp.query("prim_mglev", p.prim_mglev);
    pp.query("bottom", p.bottom); pp.query("dump_dir", p.dump_dir);
    pp.query("tol_rel", p.tol_rel); pp.query("tol_abs", p.tol_abs);
    pp.query("gauss_seidel", p.gauss_seidel); pp.query("composite_solve", p.composite_solve); pp.query("plotfile", p.plotfile);
    pp.query("use_gmres", p.use_gmres); pp.query("gmres_precond", p.gmres_precond); pp.query("gmres_precond_iters", p.gmres_precond_iters);
    return p;
}

// ---- problem data -------------------------------------------------------------------------------------
struct Problem {
    Vector<Geometry> geom; Vector<BoxArray> grids; Vector<DistributionMapping> dmap;
    Vector<MultiFab> sol, rhs, exact, acoef, bcoef;
    Vector<Array<MultiFab,3>> bface;
    Vector<MultiFab> robin_a, robin_b, robin_f;      // prob_type 6: Robin data a*phi + b*dphi/dn = f in the ghost cells
};

// 2: variable-coefficient ABecLap; 3: its fields with inhomogeneous Neumann data on every face; 6: with Robin data on the x and z faces
inline bool is_abec (Params const& p) { return p.prob_type == 2 || p.prob_type == 3 || p.prob_type == 6; }

constexpr double kPi = 3.1415926535897932;

double beta_bubble (double x, double y, double z)
{
    const double w = 0.05, sigma = 10.0, theta = 0.5*std::log(3.0)/(w+1.e-50);
    const double r = std::sqrt((x-0.5)*(x-0.5)+(y-0.5)*(y-0.5)+(z-0.5)*(z-0.5));
    return (sigma-1.)/2.*std::tanh(theta*(r-0.25)) + (sigma+1.)/2.;
}

void build_problem (Params const& p, Problem& P)
{
    const int nlev = p.max_level+1;
    P.geom.resize(nlev); P.grids.resize(nlev); P.dmap.resize(nlev);
    P.sol.resize(nlev); P.rhs.resize(nlev); P.exact.resize(nlev);
    const bool abec = is_abec(p);
    if (abec) { P.acoef.resize(nlev); P.bcoef.resize(nlev); P.bface.resize(nlev); }
    if (p.prob_type == 6) { P.robin_a.resize(nlev); P.robin_b.resize(nlev); P.robin_f.resize(nlev); }

    RealBox rb({0.,0.,0.},{1.,1.,1.});
    const int per = (p.prob_type == 5) ? 1 : 0;
    Array<int,3> is_per{per,per,per};
    Geometry::Setup(&rb, 0, is_per.data());
    Box dom0(IntVect(0), IntVect(p.n_cell-1));
    Box dom = dom0;
    for (int l = 0; l < nlev; ++l) { P.geom[l].define(dom); dom.refine(p.ref_ratio); }
    dom = dom0;
    for (int l = 0; l < nlev; ++l) {
        P.grids[l].define(dom); P.grids[l].maxSize(p.max_grid_size);
        dom.grow(-p.n_cell/4); dom.refine(p.ref_ratio);
    }
    for (int l = 0; l < nlev; ++l) {
        P.dmap[l].define(P.grids[l]);
        P.sol[l].define(P.grids[l], P.dmap[l], 1, 1);
        P.rhs[l].define(P.grids[l], P.dmap[l], 1, 0);
        P.exact[l].define(P.grids[l], P.dmap[l], 1, 0);
        if (abec) {
            P.acoef[l].define(P.grids[l], P.dmap[l], 1, 0);
            P.bcoef[l].define(P.grids[l], P.dmap[l], 1, 1);
        }
        if (p.prob_type == 7) {      // MLALaplacian: alpha*a(x)*phi - beta*Lap(phi), a = the bubble field of problem 2, fields of problem 1
            if (int(P.acoef.size()) < nlev) { P.acoef.resize(nlev); }
            P.acoef[l].define(P.grids[l], P.dmap[l], 1, 0);
            const auto dxa = P.geom[l].CellSizeArray();
            for (MFIter mfi(P.acoef[l]); mfi.isValid(); ++mfi) {
                auto al = P.acoef[l].array(mfi);
                amrex::LoopOnCpu(mfi.validbox(), [&] (int i, int j, int k) { al(i,j,k) = beta_bubble(dxa[0]*(i+0.5), dxa[1]*(j+0.5), dxa[2]*(k+0.5)); });
            }
        }
        if (p.prob_type == 6) {
            // smooth positive a, b and a smooth f everywhere (only the ghost cells outside the Robin faces are read)
            P.robin_a[l].define(P.grids[l], P.dmap[l], 1, 1); P.robin_b[l].define(P.grids[l], P.dmap[l], 1, 1); P.robin_f[l].define(P.grids[l], P.dmap[l], 1, 1);
            const auto dxr = P.geom[l].CellSizeArray();
            for (MFIter mfi(P.robin_a[l]); mfi.isValid(); ++mfi) {
                auto ra = P.robin_a[l].array(mfi); auto rb = P.robin_b[l].array(mfi); auto rf = P.robin_f[l].array(mfi);
                amrex::LoopOnCpu(amrex::grow(mfi.validbox(),1), [&] (int i, int j, int k) {
                    const double x = dxr[0]*(i+0.5), y = dxr[1]*(j+0.5), z = dxr[2]*(k+0.5);
                    ra(i,j,k) = 1.0 + 0.5*std::cos(2.*kPi*y)*std::cos(2.*kPi*z);
                    rb(i,j,k) = 1.0 + 0.25*std::sin(2.*kPi*(x+z));
                    rf(i,j,k) = std::sin(2.*kPi*(x+y)) + 0.3;
                });
            }
        }
        const auto dx = P.geom[l].CellSizeArray();
        const double a = p.ascalar, b = p.bscalar;
        for (MFIter mfi(P.rhs[l]); mfi.isValid(); ++mfi) {
            const Box vbx = mfi.validbox();
            const Box gbx = amrex::grow(vbx,1);
            auto r = P.rhs[l].array(mfi); auto s = P.sol[l].array(mfi); auto e = P.exact[l].array(mfi);
            if (!abec) {
                const double tpi = 2.*kPi, fpi = 4.*kPi, fac = tpi*tpi*3.0;
                amrex::LoopOnCpu(gbx, [&] (int i, int j, int k) { s(i,j,k) = 0.0; });
                amrex::LoopOnCpu(vbx, [&] (int i, int j, int k) {
                    double x = dx[0]*(i+0.5), y = dx[1]*(j+0.5), z = dx[2]*(k+0.5);
                    e(i,j,k) = (std::sin(tpi*x)*std::sin(tpi*y)*std::sin(tpi*z))
                        + .25*(std::sin(fpi*x)*std::sin(fpi*y)*std::sin(fpi*z));
                    r(i,j,k) = -fac*(std::sin(tpi*x)*std::sin(tpi*y)*std::sin(tpi*z))
                               -fac*(std::sin(fpi*x)*std::sin(fpi*y)*std::sin(fpi*z));
                });
            } else {
                auto al = P.acoef[l].array(mfi); auto be = P.bcoef[l].array(mfi);
                const double w = 0.05, sigma = 10.0, theta = 0.5*std::log(3.0)/(w+1.e-50);
                const double pi = kPi, tpi = 2.*pi, fpi = 4.*pi, fac = 12.0*pi*pi;
                amrex::LoopOnCpu(gbx, [&] (int i, int j, int k) {
                    double x = dx[0]*(i+0.5), y = dx[1]*(j+0.5), z = dx[2]*(k+0.5);
                    const double xc = 0.5, yc = 0.5, zc = 0.5;
                    be(i,j,k) = beta_bubble(x,y,z);
                    if (vbx.contains(i,j,k)) {
                        double rr = std::sqrt((x-xc)*(x-xc)+(y-yc)*(y-yc)+(z-zc)*(z-zc));
                        double tmp = std::cosh(theta*(rr-0.25));
                        double dbdrfac = (sigma-1.)/2./(tmp*tmp)*theta/rr;
                        dbdrfac *= b;
                        al(i,j,k) = 1.0;
                        s(i,j,k) = 0.0;
                        e(i,j,k) = std::cos(tpi*x)*std::cos(tpi*y)*std::cos(tpi*z)
                            + .25*std::cos(fpi*x)*std::cos(fpi*y)*std::cos(fpi*z);
                        r(i,j,k) = be(i,j,k)*b*fac*(std::cos(tpi*x)*std::cos(tpi*y)*std::cos(tpi*z)
                                                   + std::cos(fpi*x)*std::cos(fpi*y)*std::cos(fpi*z))
                            + dbdrfac*((x-xc)*(tpi*std::sin(tpi*x)*std::cos(tpi*y)*std::cos(tpi*z)
                                              + pi*std::sin(fpi*x)*std::cos(fpi*y)*std::cos(fpi*z))
                                     + (y-yc)*(tpi*std::cos(tpi*x)*std::sin(tpi*y)*std::cos(tpi*z)
                                              + pi*std::cos(fpi*x)*std::sin(fpi*y)*std::cos(fpi*z))
                                     + (z-zc)*(tpi*std::cos(tpi*x)*std::cos(tpi*y)*std::sin(tpi*z)
                                              + pi*std::cos(fpi*x)*std::cos(fpi*y)*std::sin(fpi*z)))
                            + a*(std::cos(tpi*x)*std::cos(tpi*y)*std::cos(tpi*z)
                                 + 0.25*std::cos(fpi*x)*std::cos(fpi*y)*std::cos(fpi*z));
                    } else if (p.prob_type == 3) {
                        // inhomogeneous Neumann: the ghost cell holds d(phi)/dn on the boundary face (the reference's
                        // convention, Tests/LinearSolvers/ABecLaplacian_C/initProb.cpp:105-143); any smooth data will do
                        double xb = std::min(std::max(x,0.0),1.0), yb = std::min(std::max(y,0.0),1.0),
                               zb = std::min(std::max(z,0.0),1.0);
                        s(i,j,k) = 0.5*std::sin(tpi*(xb + 2.0*yb))*std::cos(tpi*zb) + 0.1;
                    } else {
                        double xb = std::min(std::max(x,0.0),1.0), yb = std::min(std::max(y,0.0),1.0),
                               zb = std::min(std::max(z,0.0),1.0);
                        s(i,j,k) = std::cos(tpi*xb)*std::cos(tpi*yb)*std::cos(tpi*zb)
                            + .25*std::cos(fpi*xb)*std::cos(fpi*yb)*std::cos(fpi*zb);
                    }
                });
            }
        }
        if (abec) {
            for (int d = 0; d < 3; ++d) {
                P.bface[l][d].define(amrex::convert(P.grids[l], IntVect::TheDimensionVector(d)), P.dmap[l], 1, 0);
            }
            amrex::average_cellcenter_to_face(GetArrOfPtrs(P.bface[l]), P.bcoef[l], P.geom[l]);
        }
    }
}

// ---- raw dumps -------------------------------------------------------------------------------------------
// Writes mf (component 0, incl. ng ghost layers) as ONE Fortran-order fp64 array over the bounding box of
// the level's (index-type-converted) BoxArray grown by ng.  Cells no fab covers are left 0.  Where
// grown boxes overlap, VALID data wins over ghost data.
void dump_mf (std::string const& dir, std::string const& name, MultiFab const& mf, int ng, std::ostream& manifest)
{
    Box bb = mf.boxArray().minimalBox(); bb.grow(ng);
    const Long nx = bb.length(0), ny = bb.length(1), nz = bb.length(2);
    std::vector<double> buf(nx*ny*nz, 0.0);
    const auto lo = bb.smallEnd();
    for (int pass = 0; pass < 2; ++pass) {
        for (MFIter mfi(mf); mfi.isValid(); ++mfi) {
            auto a = mf.const_array(mfi);
            Box b = (pass == 0) ? amrex::grow(mfi.validbox(), ng) : mfi.validbox();
            amrex::LoopOnCpu(b, [&] (int i, int j, int k) {
                buf[(i-lo[0]) + nx*((j-lo[1]) + ny*(k-lo[2]))] = a(i,j,k);
            });
        }
    }
    std::ofstream f(dir+"/"+name+".bin", std::ios::binary);
    f.write(reinterpret_cast<const char*>(buf.data()), buf.size()*sizeof(double));
    manifest << "\"" << name << "\": {\"lo\": [" << lo[0] << "," << lo[1] << "," << lo[2] << "], \"shape\": ["
             << nx << "," << ny << "," << nz << "]},\n";
}

std::string box_json (Box const& b)
{
    std::ostringstream s;
    s << "[" << b.smallEnd(0) << "," << b.smallEnd(1) << "," << b.smallEnd(2) << ","
      << b.bigEnd(0) << "," << b.bigEnd(1) << "," << b.bigEnd(2) << ","
      << b.ixType()[0] << "," << b.ixType()[1] << "," << b.ixType()[2] << "]";
    return s.str();
}

template <class V> std::string vec_json (V const& v)
{
    std::ostringstream s; s.precision(17); s << "[";
    for (size_t i = 0; i < v.size(); ++i) { if (i) s << ","; s << v[i]; }
    s << "]"; return s.str();
}

// ---- linop factories -------------------------------------------------------------------------------------
struct ProbeABec : public MLABecLaplacian {
    using MLABecLaplacian::MLABecLaplacian;
    using MLABecLaplacian::m_grids; using MLABecLaplacian::m_dmap; using MLABecLaplacian::m_geom;
    using MLABecLaplacian::m_num_mg_levels; using MLABecLaplacian::m_a_coeffs; using MLABecLaplacian::m_b_coeffs;
};
struct ProbePoisson : public MLPoisson {
    using MLPoisson::MLPoisson;
    using MLPoisson::m_grids; using MLPoisson::m_dmap; using MLPoisson::m_geom; using MLPoisson::m_num_mg_levels;
};

LPInfo make_info (Params const& p)
{
    LPInfo info;
    info.setAgglomeration(p.agglomeration); info.setConsolidation(p.consolidation);
    info.setMaxCoarseningLevel(p.max_coarsening_level);
    info.setAgglomerationGridSize(p.agg_grid_size); info.setConsolidationGridSize(p.agg_grid_size);
    return info;
}

void setup_abec (Params const& p, Problem& P, MLABecLaplacian& op)
{
    op.setMaxOrder(p.maxorder);
    op.setGaussSeidel(p.gauss_seidel != 0);
    if (p.prob_type == 3) {
        const auto t = LinOpBCType::inhomogNeumann;
        op.setDomainBC({t,t,t},{t,t,t});
    } else if (p.prob_type == 6) {
        op.setDomainBC({LinOpBCType::Robin, LinOpBCType::Dirichlet, LinOpBCType::Robin},
                       {LinOpBCType::Robin, LinOpBCType::Neumann, LinOpBCType::Robin});
    } else {
        op.setDomainBC({LinOpBCType::Dirichlet, LinOpBCType::Neumann, LinOpBCType::Neumann},
                       {LinOpBCType::Neumann, LinOpBCType::Dirichlet, LinOpBCType::Neumann});
    }
    for (int l = 0; l <= p.max_level; ++l) {
        if (p.prob_type == 6) { op.setLevelBC(l, &P.sol[l], &P.robin_a[l], &P.robin_b[l], &P.robin_f[l]); }
        else { op.setLevelBC(l, &P.sol[l]); }
    }
    op.setScalars(p.ascalar, p.bscalar);
    for (int l = 0; l <= p.max_level; ++l) {
        op.setACoeffs(l, P.acoef[l]);
        op.setBCoeffs(l, amrex::GetArrOfConstPtrs(P.bface[l]));
    }
}

void setup_poisson (Params const& p, Problem& P, MLPoisson& op)
{
    op.setMaxOrder(p.maxorder);
    op.setGaussSeidel(p.gauss_seidel != 0);
    const auto t = (p.prob_type == 5) ? LinOpBCType::Periodic : LinOpBCType::Dirichlet;
    op.setDomainBC({t,t,t},{t,t,t});
    for (int l = 0; l <= p.max_level; ++l) { op.setLevelBC(l, &P.sol[l]); }
}

void set_bottom (Params const& p, MLMG& mlmg)
{
    if (p.bottom == "smoother") mlmg.setBottomSolver(BottomSolver::smoother);
    else if (p.bottom == "bicgstab") mlmg.setBottomSolver(BottomSolver::bicgstab);
    else if (p.bottom == "cg") mlmg.setBottomSolver(BottomSolver::cg);
    else if (p.bottom == "bicgcg") mlmg.setBottomSolver(BottomSolver::bicgcg);
    else if (p.bottom == "cgbicg") mlmg.setBottomSolver(BottomSolver::cgbicg);
}

void dump_inputs (Params const& p, Problem& P, std::ostream& man)
{
    for (int l = 0; l <= p.max_level; ++l) {
        std::string s = "_lev" + std::to_string(l);
        dump_mf(p.dump_dir, "sol0"+s, P.sol[l], 1, man);
        dump_mf(p.dump_dir, "rhs"+s, P.rhs[l], 0, man);
        dump_mf(p.dump_dir, "exact"+s, P.exact[l], 0, man);
        if (is_abec(p)) {
            dump_mf(p.dump_dir, "acoef"+s, P.acoef[l], 0, man);
            dump_mf(p.dump_dir, "bx"+s, P.bface[l][0], 0, man);
            dump_mf(p.dump_dir, "by"+s, P.bface[l][1], 0, man);
            dump_mf(p.dump_dir, "bz"+s, P.bface[l][2], 0, man);
        }
        if (p.prob_type == 7) { dump_mf(p.dump_dir, "acoef"+s, P.acoef[l], 0, man); }
        if (p.prob_type == 6) {
            dump_mf(p.dump_dir, "robin_a"+s, P.robin_a[l], 1, man);
            dump_mf(p.dump_dir, "robin_b"+s, P.robin_b[l], 1, man);
            dump_mf(p.dump_dir, "robin_f"+s, P.robin_f[l], 1, man);
        }
    }
}

// ---- mode=solve --------------------------------------------------------------------------------------------
int run_solve (Params const& p)
{
    Problem P; build_problem(p, P);
    std::ofstream man;
    if (!p.dump_dir.empty()) { man.open(p.dump_dir+"/manifest.json"); man << "{\n"; dump_inputs(p, P, man); }

    std::unique_ptr<MLLinOp> op;
    LPInfo info = make_info(p);
    if (is_abec(p)) {
        auto o = std::make_unique<MLABecLaplacian>(P.geom, P.grids, P.dmap, info); setup_abec(p, P, *o); op = std::move(o);
    } else if (p.prob_type == 7) {
        auto o = std::make_unique<MLALaplacian>(P.geom, P.grids, P.dmap, info);
        o->setMaxOrder(p.maxorder);
        o->setDomainBC({LinOpBCType::Dirichlet,LinOpBCType::Dirichlet,LinOpBCType::Dirichlet},{LinOpBCType::Dirichlet,LinOpBCType::Dirichlet,LinOpBCType::Dirichlet});
        for (int l = 0; l <= p.max_level; ++l) { o->setLevelBC(l, &P.sol[l]); }
        o->setScalars(1.0, 1.0);
        for (int l = 0; l <= p.max_level; ++l) { o->setACoeffs(l, P.acoef[l]); }
        op = std::move(o);
    } else {
        auto o = std::make_unique<MLPoisson>(P.geom, P.grids, P.dmap, info); setup_poisson(p, P, *o); op = std::move(o);
    }
    Vector<MultiFab> sol0(p.max_level+1);
    for (int l = 0; l <= p.max_level; ++l) {
        sol0[l].define(P.grids[l], P.dmap[l], 1, 1); MultiFab::Copy(sol0[l], P.sol[l], 0, 0, 1, 1);
    }
    std::vector<double> times; int iters = 0; Vector<Real> hist; Real rhs0 = 0, res0 = 0, fin = 0; Vector<int> cgit;
    if (!p.composite_solve) {
        // level-by-level solves as the reference's test driver does them (Tests/LinearSolvers/ABecLaplacian_C/MyTest.cpp:104-141,
        // 226-279): one single-level operator per AMR level, fine levels take Dirichlet data from the level below
        auto t0 = std::chrono::steady_clock::now();
        for (int l = 0; l <= p.max_level; ++l) {
            std::unique_ptr<MLLinOp> lop;
            if (is_abec(p)) {
                auto o = std::make_unique<MLABecLaplacian>(Vector<Geometry>{P.geom[l]}, Vector<BoxArray>{P.grids[l]}, Vector<DistributionMapping>{P.dmap[l]}, info);
                o->setMaxOrder(p.maxorder); o->setGaussSeidel(p.gauss_seidel != 0);
                o->setDomainBC({LinOpBCType::Dirichlet, LinOpBCType::Neumann, LinOpBCType::Neumann},
                               {LinOpBCType::Neumann, LinOpBCType::Dirichlet, LinOpBCType::Neumann});
                if (l > 0) { o->setCoarseFineBC(&P.sol[l-1], p.ref_ratio); }
                o->setLevelBC(0, &P.sol[l]);
                o->setScalars(p.ascalar, p.bscalar);
                o->setACoeffs(0, P.acoef[l]);
                o->setBCoeffs(0, amrex::GetArrOfConstPtrs(P.bface[l]));
                lop = std::move(o);
            } else {
                auto o = std::make_unique<MLPoisson>(Vector<Geometry>{P.geom[l]}, Vector<BoxArray>{P.grids[l]}, Vector<DistributionMapping>{P.dmap[l]}, info);
                o->setMaxOrder(p.maxorder); o->setGaussSeidel(p.gauss_seidel != 0);
                o->setDomainBC({LinOpBCType::Dirichlet, LinOpBCType::Dirichlet, LinOpBCType::Dirichlet},
                               {LinOpBCType::Dirichlet, LinOpBCType::Dirichlet, LinOpBCType::Dirichlet});
                if (l > 0) { o->setCoarseFineBC(&P.sol[l-1], p.ref_ratio); }
                o->setLevelBC(0, &P.sol[l]);
                lop = std::move(o);
            }
            MLMG mlmg(*lop);
            mlmg.setMaxIter(p.max_iter); mlmg.setMaxFmgIter(p.max_fmg_iter);
            mlmg.setVerbose(p.verbose); mlmg.setBottomVerbose(p.bottom_verbose);
            fin = mlmg.solve({&P.sol[l]}, {&P.rhs[l]}, p.tol_rel, p.tol_abs);
            iters = mlmg.getNumIters(); hist = mlmg.getResidualHistory();       // of the finest level solved
            rhs0 = mlmg.getInitRHS(); res0 = mlmg.getInitResidual(); cgit = mlmg.getNumCGIters();
        }
        times.push_back(std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count());
    }
    for (int is = 0; is < (p.composite_solve ? p.nsolve : 0); ++is) {
        for (int l = 0; l <= p.max_level; ++l) { MultiFab::Copy(P.sol[l], sol0[l], 0, 0, 1, 1); }
        MLMG mlmg(*op);
        mlmg.setMaxIter(p.max_iter); mlmg.setMaxFmgIter(p.max_fmg_iter);
        mlmg.setVerbose(p.verbose); mlmg.setBottomVerbose(p.bottom_verbose);
        set_bottom(p, mlmg);
        if (p.use_gmres) {
            // GMRES preconditioned by one MLMG V-cycle (Tests/LinearSolvers/ABecLaplacian_C/MyTest.cpp:466-532, inputs.gmres)
            AMREX_ALWAYS_ASSERT(p.max_level == 0);
            GMRESMLMG gm(mlmg);
            gm.usePrecond(p.gmres_precond != 0);
            gm.setPrecondNumIters(p.gmres_precond_iters);
            gm.setVerbose(p.verbose);
            auto t0 = std::chrono::steady_clock::now();
            gm.solve(P.sol[0], P.rhs[0], p.tol_rel, p.tol_abs);
            auto t1 = std::chrono::steady_clock::now();
            times.push_back(std::chrono::duration<double>(t1-t0).count());
            iters = gm.getNumIters(); fin = gm.getResidualNorm(); hist.clear();
            continue;
        }
        auto t0 = std::chrono::steady_clock::now();
        fin = mlmg.solve(GetVecOfPtrs(P.sol), GetVecOfConstPtrs(P.rhs), p.tol_rel, p.tol_abs);
        auto t1 = std::chrono::steady_clock::now();
        times.push_back(std::chrono::duration<double>(t1-t0).count());
        iters = mlmg.getNumIters(); hist = mlmg.getResidualHistory();
        rhs0 = mlmg.getInitRHS(); res0 = mlmg.getInitResidual(); cgit = mlmg.getNumCGIters();
        if (!p.dump_dir.empty() && is == p.nsolve - 1) {
            // post-solve API of the reference (MLMG::getFluxes / getGradSolution, face-centred): AMReX_MLMG.H:556-640
            Vector<Array<MultiFab,3>> flux(p.max_level+1), grad(p.max_level+1);
            Vector<Array<MultiFab*,3>> pf(p.max_level+1), pg(p.max_level+1);
            for (int l = 0; l <= p.max_level; ++l) for (int d = 0; d < 3; ++d) {
                BoxArray fba = amrex::convert(P.grids[l], IntVect::TheDimensionVector(d));
                flux[l][d].define(fba, P.dmap[l], 1, 0); grad[l][d].define(fba, P.dmap[l], 1, 0);
                pf[l][d] = &flux[l][d]; pg[l][d] = &grad[l][d];
            }
            mlmg.getFluxes(pf);
            mlmg.getGradSolution(pg);
            for (int l = 0; l <= p.max_level; ++l) for (int d = 0; d < 3; ++d) {
                dump_mf(p.dump_dir, "flux"+std::to_string(d)+"_lev"+std::to_string(l), flux[l][d], 0, man);
                dump_mf(p.dump_dir, "grad"+std::to_string(d)+"_lev"+std::to_string(l), grad[l][d], 0, man);
            }
        }
    }
    std::vector<double> err; Long ncells = 0;
    for (int l = 0; l <= p.max_level; ++l) {
        MultiFab d(P.grids[l], P.dmap[l], 1, 0);
        MultiFab::Copy(d, P.sol[l], 0, 0, 1, 0);
        if (p.prob_type == 5) { // singular: compare up to a constant
            Real m = d.sum(0)/P.grids[l].numPts(); d.plus(-m, 0, 1, 0);
        }
        MultiFab::Subtract(d, P.exact[l], 0, 0, 1, 0);
        err.push_back(d.norminf()); ncells += P.grids[l].numPts();
    }
    if (!p.plotfile.empty()) {
        // Tests/LinearSolvers/ABecLaplacian_C/MyTestPlotfile.cpp:52-84 (without the coefficient components)
        const int nlevels = p.max_level + 1;
        Vector<MultiFab> plotmf(nlevels);
        for (int l = 0; l < nlevels; ++l) {
            plotmf[l].define(P.grids[l], P.dmap[l], 4, 0);
            MultiFab::Copy(plotmf[l], P.sol[l], 0, 0, 1, 0);
            MultiFab::Copy(plotmf[l], P.rhs[l], 0, 1, 1, 0);
            MultiFab::Copy(plotmf[l], P.exact[l], 0, 2, 1, 0);
            MultiFab::Copy(plotmf[l], P.sol[l], 0, 3, 1, 0);
            MultiFab::Subtract(plotmf[l], plotmf[l], 2, 3, 1, 0);
        }
        WriteMultiLevelPlotfile(p.plotfile, nlevels, amrex::GetVecOfConstPtrs(plotmf),
                                {"solution", "rhs", "exact_solution", "error"}, P.geom, 0.0, Vector<int>(nlevels, 0),
                                Vector<IntVect>(nlevels, IntVect(p.ref_ratio)));
    }
    if (!p.dump_dir.empty()) {
        for (int l = 0; l <= p.max_level; ++l) { dump_mf(p.dump_dir, "sol_lev"+std::to_string(l), P.sol[l], 1, man); }
        man << "\"_end\": 0\n}\n";
    }
    std::printf("RESULT {\"mode\":\"solve\",\"prob_type\":%d,\"n_cell\":%d,\"max_grid_size\":%d,\"max_level\":%d,"
                "\"maxorder\":%d,\"iters\":%d,\"rhsnorm0\":%.17g,\"resnorm0\":%.17g,\"final_resnorm\":%.17g,"
                "\"history\":%s,\"cg_iters\":%s,\"err_inf\":%s,\"solve_times\":%s,\"ncells\":%lld,\"omp_threads\":%d}\n",
                p.prob_type, p.n_cell, p.max_grid_size, p.max_level, p.maxorder, iters, rhs0, res0, fin,
                vec_json(hist).c_str(), vec_json(cgit).c_str(), vec_json(err).c_str(), vec_json(times).c_str(),
                (long long)ncells, OpenMP::get_max_threads());
    return 0;
}

// ---- mode=meta ----------------------------------------------------------------------------------------------
template <class OP>
void print_hierarchy (OP const& op, int namr)
{
    std::printf("\"hierarchy\": [");
    for (int a = 0; a < namr; ++a) {
        std::printf("%s[", a ? "," : "");
        for (int m = 0; m < op.m_num_mg_levels[a]; ++m) {
            auto const& ba = op.m_grids[a][m]; auto const& dm = op.m_dmap[a][m];
            std::printf("%s{\"domain\":%s,\"boxes\":[", m ? "," : "", box_json(op.m_geom[a][m].Domain()).c_str());
            for (int i = 0; i < ba.size(); ++i) { std::printf("%s%s", i ? "," : "", box_json(ba[i]).c_str()); }
            std::printf("],\"dmap\":%s}", vec_json(dm.ProcessorMap()).c_str());
        }
        std::printf("]");
    }
    std::printf("],\n");
    // amrex::isMFIterSafe between consecutive MG levels (AMReX_FabArrayBase.H: same DistributionMapping AND the two
    // BoxArrays share their box list): decides direct vs temporary+ParallelCopy paths, e.g. in interpAssign
    std::printf("\"mfiter_safe\": [");
    for (int a = 0; a < namr; ++a) {
        std::printf("%s[", a ? "," : "");
        for (int m = 0; m + 1 < op.m_num_mg_levels[a]; ++m) {
            const bool safe = (op.m_dmap[a][m] == op.m_dmap[a][m+1]) && BoxArray::SameRefs(op.m_grids[a][m], op.m_grids[a][m+1]);
            std::printf("%s%d", m ? "," : "", safe ? 1 : 0);
        }
        std::printf("]");
    }
    std::printf("],\n");
}

void print_fb (const char* key, MultiFab const& mf, int ng, Periodicity const& per, bool cross)
{
    auto const& fb = mf.getFB(IntVect(ng), per, cross, false);
    std::printf("\"%s\": [", key);
    bool first = true;
    for (auto const& t : *fb.m_LocTags) {
        std::printf("%s{\"dbox\":%s,\"sbox\":%s,\"dst\":%d,\"src\":%d}", first ? "" : ",",
                    box_json(t.dbox).c_str(), box_json(t.sbox).c_str(), t.dstIndex, t.srcIndex);
        first = false;
    }
    std::printf("],\n");
}

int run_meta (Params const& p)
{
    Problem P; build_problem(p, P);
    LPInfo info = make_info(p);
    std::printf("META {\n");
    if (p.prob_type == 2) { ProbeABec op(P.geom, P.grids, P.dmap, info); print_hierarchy(op, p.max_level+1); }
    else { ProbePoisson op(P.geom, P.grids, P.dmap, info); print_hierarchy(op, p.max_level+1); }
    for (int l = 0; l <= p.max_level; ++l) {
        std::string k = "fb_cross_ng1_lev"+std::to_string(l);
        print_fb(k.c_str(), P.sol[l], 1, P.geom[l].periodicity(), true);
        k = "fb_full_ng1_lev"+std::to_string(l);
        print_fb(k.c_str(), P.sol[l], 1, P.geom[l].periodicity(), false);
    }
    // SFC maps for several rank counts (Appendix B-2; AMReX_DistributionMapping.cpp:1891-1921)
    std::printf("\"sfc\": {");
    bool first = true;
    for (int np : {1,2,3,4,8}) {
        for (int l = 0; l <= p.max_level; ++l) {
            auto buckets = DistributionMapping::makeSFC(P.grids[l], true, np);
            std::vector<int> pmap(P.grids[l].size(), -1);
            for (int r = 0; r < np; ++r) { for (int b : buckets[r]) { pmap[b] = r; } }
            std::printf("%s\"np%d_lev%d\":%s", first ? "" : ",", np, l, vec_json(pmap).c_str());
            first = false;
        }
    }
    std::printf("},\n\"_end\":0}\n");
    return 0;
}

// ---- mode=amr -----------------------------------------------------------------------------------------------
// Stages of the multi-level composite solve through the reference's public API: composite residual of the initial
// guess (coarse/fine boundary interpolation + reflux), and the solution after 1 and 2 fixed MLMG iterations.
int run_amr (Params const& p)
{
    Problem P; build_problem(p, P);
    std::ofstream man(p.dump_dir+"/manifest.json"); man << "{\n";
    dump_inputs(p, P, man);
    std::unique_ptr<MLLinOp> op;
    LPInfo info = make_info(p);
    if (is_abec(p)) {
        auto o = std::make_unique<MLABecLaplacian>(P.geom, P.grids, P.dmap, info); setup_abec(p, P, *o); op = std::move(o);
    } else {
        auto o = std::make_unique<MLPoisson>(P.geom, P.grids, P.dmap, info); setup_poisson(p, P, *o); op = std::move(o);
    }
    const int nlev = p.max_level+1;
    Vector<MultiFab> sol0(nlev), res(nlev);
    for (int l = 0; l < nlev; ++l) {
        sol0[l].define(P.grids[l], P.dmap[l], 1, 1); MultiFab::Copy(sol0[l], P.sol[l], 0, 0, 1, 1);
        res[l].define(P.grids[l], P.dmap[l], 1, 0);
    }
    {
        MLMG mlmg(*op);
        mlmg.setVerbose(0);
        mlmg.compResidual(GetVecOfPtrs(res), GetVecOfPtrs(P.sol), GetVecOfConstPtrs(P.rhs));
        for (int l = 0; l < nlev; ++l) { dump_mf(p.dump_dir, "amr_res_lev"+std::to_string(l), res[l], 0, man); }
    }
    std::vector<double> h1;
    for (int nit = 1; nit <= 2; ++nit) {
        for (int l = 0; l < nlev; ++l) { MultiFab::Copy(P.sol[l], sol0[l], 0, 0, 1, 1); }
        MLMG mlmg(*op);
        mlmg.setVerbose(0); mlmg.setFixedIter(nit); set_bottom(p, mlmg);
        mlmg.solve(GetVecOfPtrs(P.sol), GetVecOfConstPtrs(P.rhs), p.tol_rel, p.tol_abs);
        for (int l = 0; l < nlev; ++l) { dump_mf(p.dump_dir, "amr_sol"+std::to_string(nit)+"_lev"+std::to_string(l), P.sol[l], 0, man); }
        h1.push_back(mlmg.getResidualHistory().back());
    }
    man << "\"_end\": 0\n}\n";
    std::printf("RESULT {\"mode\":\"amr\",\"resid_after_iter\":[%.17g,%.17g]}\n", h1[0], h1[1]);
    return 0;
}

// ---- mode=prim ----------------------------------------------------------------------------------------------
// Runs single primitives of the linop on deterministic pseudo-random data and dumps in/outputs.
void fill_pseudo (MultiFab& mf, int ng, unsigned seed)
{
    for (MFIter mfi(mf); mfi.isValid(); ++mfi) {
        auto a = mf.array(mfi);
        amrex::LoopOnCpu(amrex::grow(mfi.validbox(), ng), [&] (int i, int j, int k) {
            // integer hash of the GLOBAL index => decomposition independent, reproducible in numpy
            unsigned long long h = (unsigned long long)(i+7)*73856093ULL ^ (unsigned long long)(j+11)*19349663ULL
                                 ^ (unsigned long long)(k+13)*83492791ULL ^ (unsigned long long)seed*2654435761ULL;
            h ^= h >> 13; h *= 0x9E3779B97F4A7C15ULL; h ^= h >> 31;
            a(i,j,k) = double(h % 2000001ULL)/1000000.0 - 1.0;
        });
    }
}

template <class OP>
int run_prim_T (Params const& p, Problem& P, OP& op)
{
    const int m = p.prim_mglev;
    op.prepareForSolve();
    auto const& ba = op.m_grids[0][m]; auto const& dm = op.m_dmap[0][m];
    std::ofstream man(p.dump_dir+"/manifest.json"); man << "{\n";
    dump_inputs(p, P, man);
    MultiFab x(ba, dm, 1, 1), b(ba, dm, 1, 0), y(ba, dm, 1, 0);
    fill_pseudo(x, 0, 1u); x.setBndry(0.0); fill_pseudo(b, 0, 2u);
    dump_mf(p.dump_dir, "prim_x", x, 1, man); dump_mf(p.dump_dir, "prim_b", b, 0, man);
    using BCMode = LinOpEnumType::BCMode; using StateMode = LinOpEnumType::StateMode;
    // homogeneous apply (FillBoundary + BC + Fapply)
    { MultiFab xx(ba, dm, 1, 1); MultiFab::Copy(xx, x, 0, 0, 1, 1);
      op.apply(0, m, y, xx, BCMode::Homogeneous, StateMode::Correction);
      dump_mf(p.dump_dir, "prim_apply_homog", y, 0, man);
      dump_mf(p.dump_dir, "prim_x_after_bc_homog", xx, 1, man); }
    // one smooth (red then black), homogeneous BC, as inside the V-cycle
    { MultiFab xx(ba, dm, 1, 1); MultiFab::Copy(xx, x, 0, 0, 1, 1);
      op.smooth(0, m, xx, b, false);
      dump_mf(p.dump_dir, "prim_smooth1", xx, 0, man);
      op.smooth(0, m, xx, b, false);
      dump_mf(p.dump_dir, "prim_smooth2", xx, 0, man);
      // correction residual
      op.correctionResidual(0, m, y, xx, b, BCMode::Homogeneous);
      dump_mf(p.dump_dir, "prim_corres", y, 0, man);
      if (m+1 < op.m_num_mg_levels[0]) {
          MultiFab c(op.m_grids[0][m+1], op.m_dmap[0][m+1], 1, 0);
          op.restriction(0, m+1, c, y);
          dump_mf(p.dump_dir, "prim_restrict", c, 0, man);
          MultiFab f(ba, dm, 1, 0); MultiFab::Copy(f, xx, 0, 0, 1, 0);
          // as MLMGT::addInterpCorrection (AMReX_MLMG.H:1766-1788): go through a coarsened-fine temp
          // when the coarse level was agglomerated
          BoxArray cba = ba; cba.coarsen(2);
          MultiFab cfine(cba, dm, 1, 0);
          cfine.ParallelCopy(c, 0, 0, 1);
          op.interpolation(0, m, f, cfine);
          dump_mf(p.dump_dir, "prim_interp_add", f, 0, man);
      } }
    if (m == 0) { // inhomogeneous solution residual with the level BC data
        MultiFab xx(ba, dm, 1, 1); MultiFab::Copy(xx, x, 0, 0, 1, 1);
        op.solutionResidual(0, y, xx, b);
        dump_mf(p.dump_dir, "prim_solres", y, 0, man);
        dump_mf(p.dump_dir, "prim_x_after_bc_inhomog", xx, 1, man);
    }
    if constexpr (std::is_same_v<OP,ProbeABec>) {
        dump_mf(p.dump_dir, "prim_acoef_mg", op.m_a_coeffs[0][m], 0, man);
        dump_mf(p.dump_dir, "prim_bx_mg", op.m_b_coeffs[0][m][0], 0, man);
        dump_mf(p.dump_dir, "prim_by_mg", op.m_b_coeffs[0][m][1], 0, man);
        dump_mf(p.dump_dir, "prim_bz_mg", op.m_b_coeffs[0][m][2], 0, man);
    }
    man << "\"_end\": 0\n}\n";
    std::printf("RESULT {\"mode\":\"prim\",\"mglev\":%d,\"nmg\":%d}\n", m, op.m_num_mg_levels[0]);
    return 0;
}

int run_prim (Params const& p)
{
    Problem P; build_problem(p, P);
    LPInfo info = make_info(p);
    if (p.prob_type == 2) { ProbeABec op(P.geom, P.grids, P.dmap, info); setup_abec(p, P, op); return run_prim_T(p, P, op); }
    ProbePoisson op(P.geom, P.grids, P.dmap, info); setup_poisson(p, P, op); return run_prim_T(p, P, op);
}

} // namespace

int main (int argc, char* argv[])
{
    amrex::Initialize(argc, argv, true, MPI_COMM_WORLD, [] () {
        ParmParse pp("amrex"); pp.add("v", 0); pp.add("verbose", 0);
        // GPU-build semantics for FillBoundary LocTags
        ParmParse pf("fabarray"); pf.addarr("comm_tile_size", std::vector<int>{1024000,1024000,1024000});
        ParmParse pt("tiny_profiler"); pt.add("enabled", 0);
    });
    int rc = 0;
    {
        Params p = read_params();
        if (p.mode == "solve") rc = run_solve(p);
        else if (p.mode == "meta") rc = run_meta(p);
        else if (p.mode == "prim") rc = run_prim(p);
        else if (p.mode == "amr") rc = run_amr(p);
        else { std::fprintf(stderr, "unknown mode\n"); rc = 2; }
    }
    amrex::Finalize();
    return rc;
}
