// MLMG cycle driver and Krylov bottom solver (host control flow; all arithmetic is in device kernels).
// Follows the algorithm of MLMGT (AMReX_MLMG.H:356-541, 937-1116, 1228-1629, 1634-1873) and
// MLCGSolverT (AMReX_MLCGSolver.H:98-410).
#include "AMReX_MLMG.H"

#include <cuda_runtime.h>
#include <cstdlib>
#include <cstring>
#include <iomanip>
#include <iostream>
#include <sstream>

namespace amrex {

namespace {
void Print0 (std::string const& s) { if (ParallelDescriptor::IOProcessor()) { std::cout << s << std::flush; } }
template <class... A> std::string cat (A const&... a) { std::ostringstream o; o << std::setprecision(10); (o << ... << a); return o.str(); }
}

// ========================================================================================== MLCGSolver
void MLCGSolver::ensure_temps (MultiFab const& sol)
{
    if (p.empty() || !(p.boxArray() == sol.boxArray()) || !(p.DistributionMap() == sol.DistributionMap())) {
        const int ng = sol.nGrow();
        p = Lp.make(amrlev, mglev, ng); r = Lp.make(amrlev, mglev, ng);
        rh = Lp.make(amrlev, mglev, 0); v = Lp.make(amrlev, mglev, 0); t = Lp.make(amrlev, mglev, 0);
        q = Lp.make(amrlev, mglev, 0); sorig = Lp.make(amrlev, mglev, 0);
    }
}

int MLCGSolver::solve (MultiFab& sol, MultiFab const& rhs, Real eps_rel, Real eps_abs)
{
    return (solver_type == Type::BiCGStab) ? solve_bicgstab(sol, rhs, eps_rel, eps_abs) : solve_cg(sol, rhs, eps_rel, eps_abs);
}

int MLCGSolver::solve_bicgstab (MultiFab& sol, MultiFab const& rhs, Real eps_rel, Real eps_abs)
{
    using BCMode = MLLinOp::BCMode; using StateMode = MLLinOp::StateMode;
    ensure_temps(sol);
    if (initial_vec_zeroed && verbose <= 0 && sol.nGrow() >= 1 && Lp.bottomKernelEligible(mglev)) {
        // B200: the whole solve below as one single-CTA kernel (kernels/bottom.cu), one host read-back instead of ~6 per iteration
        return Lp.bottomBiCGStabKernel(mglev, sol, rhs, r, p, v, t, rh, eps_rel, eps_abs, maxiter, iter);
    }
    p.setVal(0.0); r.setVal(0.0);
    if (initial_vec_zeroed) { MultiFab::Copy(r, rhs, 0, 0, 1, 0); }
    else {
        Lp.correctionResidual(amrlev, mglev, r, sol, rhs, BCMode::Homogeneous);
        MultiFab::Copy(sorig, sol, 0, 0, 1, 0);
        sol.setVal(0.0);
    }
    Lp.normalize(amrlev, mglev, r);
    MultiFab::Copy(rh, r, 0, 0, 1, 0);

    Real rnorm = norm_inf(r);
    const Real rnorm0 = rnorm;
    if (verbose > 0) { Print0(cat("MLCGSolver_BiCGStab: Initial error (error0) =        ", rnorm0, "\n")); }
    int ret = 0;
    iter = 1;
    Real rho_1 = 0, alpha = 0, omega = 0;
    if (rnorm0 == 0 || rnorm0 < eps_abs) { return ret; }

    for (; iter <= maxiter; ++iter) {
        const Real rho = dotxy(rh, r);
        if (rho == 0) { ret = 1; break; }
        if (iter == 1) { MultiFab::Copy(p, r, 0, 0, 1, 0); }
        else {
            const Real beta = (rho / rho_1) * (alpha / omega);
            MultiFab::Saxpy(p, -omega, v, 0, 0, 1, 0);
            MultiFab::Xpay(p, beta, r, 0, 0, 1, 0);
        }
        Lp.apply(amrlev, mglev, v, p, BCMode::Homogeneous, StateMode::Correction);
        Lp.normalize(amrlev, mglev, v);
        const Real rhTv = dotxy(rh, v);
        if (rhTv != Real(0.0)) { alpha = rho / rhTv; } else { ret = 2; break; }
        MultiFab::Saxpy(sol, alpha, p, 0, 0, 1, 0);
        MultiFab::Saxpy(r, -alpha, v, 0, 0, 1, 0);
        rnorm = norm_inf(r);
        if (verbose > 2) { Print0(cat("MLCGSolver_BiCGStab: Half Iter ", std::setw(11), iter, " rel. err. ", rnorm / rnorm0, "\n")); }
        if (rnorm < eps_rel * rnorm0 || rnorm < eps_abs) { break; }
        Lp.apply(amrlev, mglev, t, r, BCMode::Homogeneous, StateMode::Correction);
        Lp.normalize(amrlev, mglev, t);
        Real tvals[2] = {dotxy(t, t, true), dotxy(t, r, true)};
        ParallelDescriptor::ReduceRealSum(tvals, 2);
        if (tvals[0] != Real(0.0)) { omega = tvals[1] / tvals[0]; } else { ret = 3; break; }
        MultiFab::Saxpy(sol, omega, r, 0, 0, 1, 0);
        MultiFab::Saxpy(r, -omega, t, 0, 0, 1, 0);
        rnorm = norm_inf(r);
        if (verbose > 2) { Print0(cat("MLCGSolver_BiCGStab: Iteration ", std::setw(11), iter, " rel. err. ", rnorm / rnorm0, "\n")); }
        if (rnorm < eps_rel * rnorm0 || rnorm < eps_abs) { break; }
        if (omega == 0) { ret = 4; break; }
        rho_1 = rho;
    }
    if (verbose > 0) { Print0(cat("MLCGSolver_BiCGStab: Final: Iteration ", std::setw(4), iter, " rel. err. ", rnorm / rnorm0, "\n")); }
    if (ret == 0 && rnorm > eps_rel * rnorm0 && rnorm > eps_abs) { ret = 8; }
    if ((ret == 0 || ret == 8) && (rnorm < rnorm0)) {
        if (!initial_vec_zeroed) { MultiFab::Add(sol, sorig, 0, 0, 1, 0); }
        if (ret == 8) { ret = 9; }
    } else {
        sol.setVal(0.0);
        if (!initial_vec_zeroed) { MultiFab::Add(sol, sorig, 0, 0, 1, 0); }
    }
    return ret;
}

int MLCGSolver::solve_cg (MultiFab& sol, MultiFab const& rhs, Real eps_rel, Real eps_abs)
{
    using BCMode = MLLinOp::BCMode; using StateMode = MLLinOp::StateMode;
    ensure_temps(sol);
    p.setVal(0.0);
    MultiFab& rr = rh;   // ng = 0 residual
    if (initial_vec_zeroed) { MultiFab::Copy(rr, rhs, 0, 0, 1, 0); }
    else {
        Lp.correctionResidual(amrlev, mglev, rr, sol, rhs, BCMode::Homogeneous);
        MultiFab::Copy(sorig, sol, 0, 0, 1, 0);
        sol.setVal(0.0);
    }
    Real rnorm = norm_inf(rr);
    const Real rnorm0 = rnorm;
    if (verbose > 0) { Print0(cat("MLCGSolver_CG: Initial error (error0) :        ", rnorm0, "\n")); }
    Real rho_1 = 0; int ret = 0; iter = 1;
    if (rnorm0 == 0 || rnorm0 < eps_abs) { return ret; }
    for (; iter <= maxiter; ++iter) {
        const Real rho = dotxy(rr, rr);
        if (rho == 0) { ret = 1; break; }
        if (iter == 1) { MultiFab::Copy(p, rr, 0, 0, 1, 0); }
        else { const Real beta = rho / rho_1; MultiFab::Xpay(p, beta, rr, 0, 0, 1, 0); }
        Lp.apply(amrlev, mglev, q, p, BCMode::Homogeneous, StateMode::Correction);
        Real alpha;
        const Real pw = dotxy(p, q);
        if (pw != Real(0.0)) { alpha = rho / pw; } else { ret = 1; break; }
        MultiFab::Saxpy(sol, alpha, p, 0, 0, 1, 0);
        MultiFab::Saxpy(rr, -alpha, q, 0, 0, 1, 0);
        rnorm = norm_inf(rr);
        if (verbose > 2) { Print0(cat("MLCGSolver_cg:       Iteration", std::setw(4), iter, " rel. err. ", rnorm / rnorm0, "\n")); }
        if (rnorm < eps_rel * rnorm0 || rnorm < eps_abs) { break; }
        rho_1 = rho;
    }
    if (verbose > 0) { Print0(cat("MLCGSolver_cg: Final Iteration", std::setw(4), iter, " rel. err. ", rnorm / rnorm0, "\n")); }
    if (ret == 0 && rnorm > eps_rel * rnorm0 && rnorm > eps_abs) { ret = 8; }
    if ((ret == 0 || ret == 8) && (rnorm < rnorm0)) {
        if (!initial_vec_zeroed) { MultiFab::Add(sol, sorig, 0, 0, 1, 0); }
        if (ret == 8) { ret = 9; }
    } else {
        sol.setVal(0.0);
        if (!initial_vec_zeroed) { MultiFab::Add(sol, sorig, 0, 0, 1, 0); }
    }
    return ret;
}

// ================================================================================================ MLMG
MLMG::MLMG (MLLinOp& a_lp) : linop(a_lp), namrlevs(a_lp.NAMRLevels()), finest_amr_lev(a_lp.NAMRLevels() - 1)
{
    if (const char* e = std::getenv("B200MG_GRAPHS")) { m_use_graphs = (e[0] != '0'); }
    if (const char* e = std::getenv("B200MG_LEG_CTAS")) { m_leg_max_ctas = std::max(1, std::atoi(e)); }
    if (const char* e = std::getenv("B200MG_LEG_NARROW")) { m_leg_narrow_cells = std::max(0, std::atoi(e)); }
}

MLMG::LegPlan::~LegPlan ()
{
    for (auto& kv : args) { if (kv.second.first) { The_Arena()->free(kv.second.first); } }
    if (d_log) { The_Arena()->free(d_log); }
    if (d_stamps) { The_Arena()->free(d_stamps); }
    if (h_log) { pinned_free(h_log); }
}

MLMG::~MLMG ()
{
    for (auto* m : {&m_graph_down, &m_graph_up}) {
        for (auto& kv : *m) { if (kv.second.exec) { cudaGraphExecDestroy(static_cast<cudaGraphExec_t>(kv.second.exec)); } }
    }
}

// ---- CUDA graphs of the coarse V-cycle legs (see AMReX_MLMG.H)
namespace { constexpr Long kGraphLevelCells = Long(128) * 128 * 128; }

// first MG level (>= mglev_top) from which every level down to the bottom has <= 128^3 cells and uses the plain colour
// sweeps (the fused pass ping-pongs arrays: their addresses would alternate under the graph); mglev_bottom if none
int MLMG::graphFirstLevel (int amrlev, int mglev_top, int mglev_bottom) const
{
    int g = mglev_bottom;
    for (int m = mglev_bottom - 1; m >= mglev_top; --m) {
        if (linop.Grids(amrlev, m).numPts() > kGraphLevelCells || linop.usesFusedSmoother(amrlev, m)) { break; }
        g = m;
    }
    return g;
}

// every device address a captured leg bakes in: the cycle's work arrays on levels lev0..lev1 (+ schedule parameters)
std::size_t MLMG::graphKey (int amrlev, int lev0, int lev1) const
{
    std::size_t h = 1469598103934665603ull;
    auto mix = [&] (std::size_t v) { h ^= v; h *= 1099511628211ull; };
    for (int m = lev0; m <= lev1; ++m) {
        mix(reinterpret_cast<std::size_t>(cor[amrlev][m].dataPtr()));
        mix(reinterpret_cast<std::size_t>(res[amrlev][m].dataPtr()));
        mix(reinterpret_cast<std::size_t>(rescor[amrlev][m].dataPtr()));
        if (m < int(cfine_mg.size()) && cfine_mg[m]) { mix(reinterpret_cast<std::size_t>(cfine_mg[m]->dataPtr())); }
    }
    for (int m = lev0; m <= lev1; ++m) { mix(linop.graphKey(amrlev, m)); }
    mix(std::size_t(nu1)); mix(std::size_t(nu2)); mix(reinterpret_cast<std::size_t>(&linop));
    return h;
}

template <class F>
void MLMG::runGraphed (CycleGraph& g, std::size_t key, bool warm, F&& body)
{
    const bool usable = m_use_graphs && !m_graphs_broken && warm && ParallelDescriptor::NProcs() == 1
        && !Gpu::profiling() && !Gpu::debugSync();
    if (!usable) { body(); return; }
    cudaStream_t s = Gpu::gpuStream();
    if (g.exec && g.key == key) {
        AMREX_CUDA_SAFE_CALL(cudaGraphLaunch(static_cast<cudaGraphExec_t>(g.exec), s));
        Gpu::countLaunch(int(g.launches));
        return;
    }
    if (g.exec) { cudaGraphExecDestroy(static_cast<cudaGraphExec_t>(g.exec)); g.exec = nullptr; }
    const long long l0 = Gpu::launchCount();
    if (cudaStreamBeginCapture(s, cudaStreamCaptureModeRelaxed) != cudaSuccess) { cudaGetLastError(); m_graphs_broken = true; body(); return; }
    body();                                              // records the launches of the leg, executes nothing
    cudaGraph_t graph = nullptr;
    cudaError_t e = cudaStreamEndCapture(s, &graph);
    cudaGraphExec_t exec = nullptr;
    if (e == cudaSuccess && graph) { e = cudaGraphInstantiate(&exec, graph, 0); }
    if (graph) { cudaGraphDestroy(graph); }
    const long long nk = Gpu::launchCount() - l0;        // kernels in the leg
    if (e != cudaSuccess || !exec) {                     // not capturable here: stay eager from now on
        cudaGetLastError(); m_graphs_broken = true;
        Gpu::countLaunch(-int(nk));
        body();
        return;
    }
    g.exec = exec; g.key = key; g.launches = nk;
    AMREX_CUDA_SAFE_CALL(cudaGraphLaunch(exec, s));      // the launches counted during the capture stand for this replay
}

Real MLMG::solve (Vector<MultiFab*> const& a_sol, Vector<MultiFab const*> const& a_rhs, Real a_tol_rel, Real a_tol_abs)
{
    if (bottom_solver == BottomSolver::Default) { bottom_solver = linop.getDefaultBottomSolver(); }
    if (bottom_solver == BottomSolver::hypre || bottom_solver == BottomSolver::petsc) { Abort("hypre/petsc bottom solvers are not available"); }

    const double solve_start_time = ParallelDescriptor::second();
    Real& composite_norminf = m_final_resnorm0;
    m_niters_cg.clear();
    m_iter_fine_resnorm0.clear();
    m_leg.slots.clear(); m_leg.launches = 0;

    prepareForSolve(a_sol, a_rhs);
    computeMLResidual(finest_amr_lev);

    Real norms[2] = {MLResNormInf(finest_amr_lev, true), MLRhsNormInf(true)};
    ParallelDescriptor::ReduceRealMax(norms, 2);
    const Real resnorm0 = norms[0], rhsnorm0 = norms[1];
    if (verbose >= 1) { Print0(cat("MLMG: Initial rhs               = ", rhsnorm0, "\n", "MLMG: Initial residual (resid0) = ", resnorm0, "\n")); }
    m_init_resnorm0 = resnorm0; m_rhsnorm0 = rhsnorm0;

    Real max_norm; std::string norm_name;
    if (always_use_bnorm || rhsnorm0 >= resnorm0) { norm_name = "bnorm"; max_norm = rhsnorm0; }
    else { norm_name = "resid0"; max_norm = resnorm0; }
    const Real res_target = std::max(a_tol_abs, std::max(a_tol_rel, Real(1.e-16)) * max_norm);

    timer[1] = 0.0;
    if (resnorm0 <= res_target) {
        composite_norminf = resnorm0;
        if (verbose >= 1) { Print0("MLMG: No iterations needed\n"); }
    } else {
        const double iter_start_time = ParallelDescriptor::second();
        bool converged = false;
        const int niters = do_fixed_number_of_iters ? do_fixed_number_of_iters : max_iters;
        for (int iter = 0; iter < niters; ++iter) {
            oneIter(iter);
            converged = false;
            // the finest level's residual and its (unmasked) norm come out of one kernel where the operator can do that
            Real fused_norm = 0.0;
            const bool have_norm = computeResidual(finest_amr_lev, &fused_norm);
            const Real fine_norminf = have_norm ? fused_norm : ResNormInf(finest_amr_lev);
            m_iter_fine_resnorm0.push_back(fine_norminf);
            composite_norminf = fine_norminf;
            if (verbose >= 2) { Print0(cat("MLMG: Iteration ", std::setw(3), iter + 1, " Fine resid/", norm_name, " = ", fine_norminf / max_norm, "\n")); }
            const bool fine_converged = (fine_norminf <= res_target);
            if (namrlevs == 1 && fine_converged) { converged = true; }
            else if (fine_converged) {
                computeMLResidual(finest_amr_lev - 1);
                const Real crse_norminf = MLResNormInf(finest_amr_lev - 1);
                if (verbose >= 2) { Print0(cat("MLMG: Iteration ", std::setw(3), iter + 1, " Crse resid/", norm_name, " = ", crse_norminf / max_norm, "\n")); }
                converged = (crse_norminf <= res_target);
                composite_norminf = std::max(fine_norminf, crse_norminf);
            }
            if (converged) {
                if (verbose >= 1) { Print0(cat("MLMG: Final Iter. ", iter + 1, " resid, resid/", norm_name, " = ", composite_norminf, ", ", composite_norminf / max_norm, "\n")); }
                break;
            } else if (composite_norminf > Real(1.e20) * max_norm) {
                if (verbose > 0) { Print0(cat("MLMG: Failing to converge after ", iter + 1, " iterations. resid, resid/", norm_name, " = ", composite_norminf, ", ", composite_norminf / max_norm, "\n")); }
                if (throw_exception) { throw error("MLMG blew up."); } else { Abort("MLMG failing so lets stop here"); }
            }
        }
        if (!converged && do_fixed_number_of_iters == 0) {
            if (verbose > 0) { Print0(cat("MLMG: Failed to converge after ", max_iters, " iterations. resid, resid/", norm_name, " = ", composite_norminf, ", ", composite_norminf / max_norm, "\n")); }
            if (throw_exception) { throw error("MLMG failed to converge."); } else { Abort("MLMG failed."); }
        }
        timer[1] = ParallelDescriptor::second() - iter_start_time;
    }

    const int ng_back = final_fill_bc ? 1 : 0;
    for (int alev = 0; alev < namrlevs; ++alev) {
        MultiFab::Copy(*a_sol[alev], sol[alev], 0, 0, 1, std::min(ng_back, a_sol[alev]->nGrow()));
    }
    Gpu::streamSynchronize();
    collectLegLog();
    timer[0] = ParallelDescriptor::second() - solve_start_time;
    if (verbose >= 1) { Print0(cat("MLMG: Timers: Solve = ", timer[0], " Iter = ", timer[1], " Bottom = ", timer[2], "\n")); }
    ++solve_called;
    return composite_norminf;
}

// The reference aliases the user's sol when it has exactly one ghost cell (AMReX_MLMG.H:983-987).  Here the solver
// always works on its own (aligned, ping-pong capable) copy and copies the result back: two extra passes per solve.
void MLMG::prepareForSolve (Vector<MultiFab*> const& a_sol, Vector<MultiFab const*> const& a_rhs)
{
    AMREX_ALWAYS_ASSERT(namrlevs <= int(a_sol.size()) && namrlevs <= int(a_rhs.size()));
    timer[0] = timer[1] = timer[2] = 0.0;
    prepareLinOp();
    if (sol.empty()) {
        sol.resize(namrlevs); rhs.resize(namrlevs);
        for (int alev = 0; alev < namrlevs; ++alev) {
            sol[alev] = linop.make(alev, 0, 1);
            rhs[alev] = linop.make(alev, 0, 0);
        }
    }
    const bool cycle_data_fresh = res.empty();
    prepareMGcycle();
    for (int alev = 0; alev < namrlevs; ++alev) {
        AMREX_ALWAYS_ASSERT_WITH_MESSAGE(a_sol[alev]->boxArray() == linop.Grids(alev) && a_sol[alev]->DistributionMap() == linop.DMap(alev),
                                         "MLMG::solve: sol must live on the operator's grids");
        MultiFab::Copy(sol[alev], *a_sol[alev], 0, 0, 1, 0);
        sol[alev].setBndry(0.0);
        MultiFab::Copy(rhs[alev], *a_rhs[alev], 0, 0, 1, 0);
        linop.applyInhomogNeumannTerm(alev, rhs[alev]);          // AMReX_MLMG.H:1009
    }
    for (int falev = finest_amr_lev; falev > 0; --falev) {
        average_down(sol[falev], sol[falev - 1], 0, 1, linop.AMRRefRatio(falev - 1));
        average_down(rhs[falev], rhs[falev - 1], 0, 1, linop.AMRRefRatio(falev - 1));
    }
    if (linop.isSingular(0) && linop.getEnforceSingularSolvable()) { makeSolvable(); }
    if (!cycle_data_fresh) { zeroCycleData(); }
    if (verbose >= 2) {
        Print0(cat("MLMG: # of AMR levels: ", namrlevs, "\n", "      # of MG levels on the coarsest AMR level: ", linop.NMGLevels(0), "\n"));
    }
}

// MLMGT::oneIter (AMReX_MLMG.H:1228-1293): down the AMR hierarchy (fine smooth, coarse composite residual), MG cycle on
// level 0, then back up (interpolate the coarse correction, fine residual with coarse BC, fine smooth).
void MLMG::prepareLinOp ()
{
    if (!linop_prepared) { linop.prepareForSolve(); linop_prepared = true; }
    else if (linop.needsUpdate()) { linop.update(); }
}

// MLMGT::prepareMGcycle (AMReX_MLMG.H:1138-1190): the fields of the V-cycle, allocated once and zeroed
void MLMG::prepareMGcycle ()
{
    if (!res.empty()) { return; }
    res.resize(namrlevs); rescor.resize(namrlevs); cor.resize(namrlevs); cor_hold.resize(std::max(namrlevs - 1, 1));
    for (int alev = 0; alev < namrlevs; ++alev) {
        const int nmg = linop.NMGLevels(alev);
        res[alev].resize(nmg); rescor[alev].resize(nmg); cor[alev].resize(nmg);
        for (int m = 0; m < nmg; ++m) {
            res[alev][m] = linop.make(alev, m, 0); rescor[alev][m] = linop.make(alev, m, 0); cor[alev][m] = linop.make(alev, m, 1);
        }
    }
    const int nmg0 = linop.NMGLevels(0);
    cor_hold[0].resize(nmg0);
    for (int m = 0; m < nmg0 - 1; ++m) { cor_hold[0][m] = linop.make(0, m, 1); }
    for (int alev = 1; alev < finest_amr_lev; ++alev) { cor_hold[alev].resize(1); cor_hold[alev][0] = linop.make(alev, 0, 1); }
    cfine_mg.resize(nmg0);
    zeroCycleData();
}

void MLMG::zeroCycleData ()
{
    for (int alev = 0; alev <= finest_amr_lev; ++alev) {
        for (int m = 0; m < linop.NMGLevels(alev); ++m) {
            res[alev][m].setVal(0.0); rescor[alev][m].setVal(0.0); cor[alev][m].setVal(0.0);
            if (alev == 0 && m < linop.NMGLevels(0) - 1) { cor_hold[0][m].setVal(0.0); }
        }
    }
    for (int alev = 1; alev < finest_amr_lev; ++alev) { cor_hold[alev][0].setVal(0.0); }
}

void MLMG::oneIter (int iter)
{
    for (int alev = finest_amr_lev; alev > 0; --alev) {
        miniCycle(alev);
        MultiFab::Add(sol[alev], cor[alev][0], 0, 0, 1, 0);
        computeResWithCrseSolFineCor(alev - 1, alev);
        if (alev != finest_amr_lev) { std::swap(cor_hold[alev][0], cor[alev][0]); }   // saved for the up cycle
    }
    {
        Gpu::ProfScope prof_scope__(0);
        if (linop.isSingular(0) && linop.getEnforceSingularSolvable()) { makeSolvable(0, 0, res[0][0]); }
        if (iter < max_fmg_iters) { mgFcycle(); } else { mgVcycle(0, 0); }
        MultiFab::Add(sol[0], cor[0][0], 0, 0, 1, 0);
    }
    for (int alev = 1; alev <= finest_amr_lev; ++alev) {
        interpCorrection(alev);                                   // (fine AMR correction) = I(coarse AMR correction)
        MultiFab::Add(sol[alev], cor[alev][0], 0, 0, 1, 0);
        if (alev != finest_amr_lev) { MultiFab::Add(cor_hold[alev][0], cor[alev][0], 0, 0, 1, 0); }
        computeResWithCrseCorFineCor(alev);
        miniCycle(alev);
        MultiFab::Add(sol[alev], cor[alev][0], 0, 0, 1, 0);
        if (alev != finest_amr_lev) { MultiFab::Add(cor[alev][0], cor_hold[alev][0], 0, 0, 1, 0); }
    }
    linop.averageDownAndSync(sol);
}

void MLMG::mgVcycle (int amrlev, int mglev_top)
{
    const int mglev_bottom = linop.NMGLevels(amrlev) - 1;
    auto down = [&] (int m0, int m1) {                   // levels [m0, m1): pre-smooth, residual, restriction
        for (int mglev = m0; mglev < m1; ++mglev) {
            Gpu::ProfScope prof_scope__(amrlev * 100 + mglev);
            // cor = 0, then nu1 smooths: the zeroing is handed to the first smooth (the fused pass needs neither the
            // setVal nor the read of its zero input; every other path zeroes inside smooth)
            if (nu1 <= 0) { cor[amrlev][mglev].setVal(0.0); }
            bool skip_fillboundary = true;
            for (int i = 0; i < nu1; ++i) {
                linop.smooth(amrlev, mglev, cor[amrlev][mglev], res[amrlev][mglev], skip_fillboundary, i == 0, i > 0);
                skip_fillboundary = false;
            }
            // residual + restriction in one pass when the level allows it (rescor of this level is then not formed)
            if (!linop.correctionResidualRestrict(amrlev, mglev, res[amrlev][mglev + 1], cor[amrlev][mglev], res[amrlev][mglev])) {
                computeResOfCorrection(amrlev, mglev);
                linop.restriction(amrlev, mglev + 1, res[amrlev][mglev + 1], rescor[amrlev][mglev]);
            }
        }
    };
    auto up = [&] (int m1, int m0) {                     // levels m1 down to m0: prolongation-add, post-smooth
        for (int mglev = m1; mglev >= m0; --mglev) {
            Gpu::ProfScope prof_scope__(amrlev * 100 + mglev);
            addInterpCorrection(amrlev, mglev);
            for (int i = 0; i < nu2; ++i) { linop.smooth(amrlev, mglev, cor[amrlev][mglev], res[amrlev][mglev], false, false, i > 0); }
        }
    };
    // B200: every level from the first single-box level down to the bottom solve and back up is ONE kernel
    const int leg0 = coarseLegFirstLevel(amrlev, mglev_top, mglev_bottom);
    if (leg0 >= 0) {
        down(mglev_top, leg0);
        runCoarseLeg(leg0, mglev_bottom, legIsMerged(leg0));
        up(leg0 - 1, mglev_top);
        ++m_cycles_done[amrlev];
        return;
    }
    // the launch-bound small levels [g, bottom) replay as CUDA graphs (see AMReX_MLMG.H); the big ones run eagerly
    const int g = graphFirstLevel(amrlev, mglev_top, mglev_bottom);
    down(mglev_top, g);
    if (g < mglev_bottom) {
        runGraphed(m_graph_down[amrlev * 1000 + g], graphKey(amrlev, g, mglev_bottom), m_cycles_done[amrlev] >= 1, [&] { down(g, mglev_bottom); });
    }
    if (amrlev == 0) { bottomSolve(); }
    else {
        cor[amrlev][mglev_bottom].setVal(0.0);
        bool skip_fillboundary = true;
        for (int i = 0; i < nu1; ++i) {
            linop.smooth(amrlev, mglev_bottom, cor[amrlev][mglev_bottom], res[amrlev][mglev_bottom], skip_fillboundary);
            skip_fillboundary = false;
        }
    }
    if (g < mglev_bottom) {
        runGraphed(m_graph_up[amrlev * 1000 + g], graphKey(amrlev, g, mglev_bottom), m_cycles_done[amrlev] >= 1, [&] { up(mglev_bottom - 1, g); });
    }
    up(g - 1, mglev_top);
    ++m_cycles_done[amrlev];
}

int MLMG::coarseLegFirstLevel (int amrlev, int mglev_top, int mglev_bottom) const
{
    if (amrlev != 0) { return -1; }
    const BottomSolver bs = (bottom_solver == BottomSolver::Default) ? linop.getDefaultBottomSolver() : bottom_solver;
    if ((bs != BottomSolver::bicgstab && bs != BottomSolver::smoother) || bottom_verbose > 0) { return -1; }
    constexpr Long kWideMax = Long(1) << 21, kBiCGMax = 32768;        // the bottom BiCGStab is one CTA's work
    // the merged copy of the operator (every level from mergedLegLevel() down re-gridded as one box)
    const int mm = linop.mergedLegLevel();
    if (mm >= mglev_top && mm <= mglev_bottom && mglev_bottom - mm + 1 <= B200MG_LEG_MAX_LEVELS && mglev_bottom == linop.NMGLevels(0) - 1) {
        MLLinOp const& op = linop.mergedOp();
        bool ok = true;
        for (int l = 0; l <= mglev_bottom - mm && ok; ++l) {
            const bool bottom = (l == mglev_bottom - mm);
            ok = op.coarseLegLevelEligible(l, (bottom && bs == BottomSolver::bicgstab) ? kBiCGMax : kWideMax);
        }
        if (ok) { return mm; }
    }
    int leg0 = -1;
    for (int m = mglev_bottom; m >= mglev_top; --m) {
        const bool bottom = (m == mglev_bottom);
        if (!linop.coarseLegLevelEligible(m, (bottom && bs == BottomSolver::bicgstab) ? kBiCGMax : kWideMax)) { break; }
        leg0 = m;
    }
    if (leg0 < 0 || mglev_bottom - leg0 + 1 > B200MG_LEG_MAX_LEVELS) { return -1; }
    return leg0;
}

namespace {
inline void leg_mix (std::size_t& h, std::size_t v) { h ^= v; h *= 1099511628211ull; }
inline std::size_t leg_bits (Real v) { std::size_t b = 0; std::memcpy(&b, &v, sizeof(Real)); return b; }
}

// merged: the leg runs on linop.mergedOp() (level l of it == MG level leg0 + l here): the residual of level leg0 moves onto
// the one box by ParallelCopy, the correction comes back the same way.
void MLMG::runCoarseLeg (int leg0, int mglev_bottom, bool merged)
{
    Gpu::ProfScope prof_scope__(leg0);
    const bool bicg = (bottom_solver != BottomSolver::smoother);
    if (bicg) {
        if (int(m_leg.slots.size()) < kLegLogMax) { m_leg.slots.push_back(int(m_niters_cg.size())); }
        m_niters_cg.push_back(-1);
    }
    MLLinOp& op = merged ? linop.mergedOp() : linop;
    const int off = merged ? leg0 : 0;                       // MG level of op = MG level of linop - off
    const int nlev = mglev_bottom - leg0 + 1;
    if (merged) {
        if (m_leg.mop != &op) {                              // first use, or update() rebuilt the copy
            m_leg.mcor.clear(); m_leg.mres.clear(); m_leg.mrescor.clear(); m_leg.mcg.reset(); m_leg.mbb.reset();
            for (int l = 0; l < nlev; ++l) {
                m_leg.mcor.push_back(op.make(0, l, 1)); m_leg.mres.push_back(op.make(0, l, 0)); m_leg.mrescor.push_back(op.make(0, l, 0));
            }
            m_leg.mop = &op;
        }
        m_leg.mres[0].ParallelCopy(res[0][leg0], 0, 0, 1);
    }
    auto Cor = [&] (int m) -> MultiFab& { return merged ? m_leg.mcor[m - off] : cor[0][m]; };
    auto Res = [&] (int m) -> MultiFab& { return merged ? m_leg.mres[m - off] : res[0][m]; };
    auto Rescor = [&] (int m) -> MultiFab& { return merged ? m_leg.mrescor[m - off] : rescor[0][m]; };
    if (op.ownsSingleBox(leg0 - off)) {                      // else another rank owns the box of these levels
        const bool singular = bicg && op.isBottomSingular() && op.getEnforceSingularSolvable();
        MLCGSolver::Temps tmp{nullptr, nullptr, nullptr, nullptr, nullptr};
        MultiFab* bb = nullptr;
        if (bicg) {
            std::unique_ptr<MLCGSolver>& cg = merged ? m_leg.mcg : cg_solver;
            if (!cg) { cg = std::make_unique<MLCGSolver>(op); }
            tmp = cg->temps(Cor(mglev_bottom));
            if (singular) {
                std::unique_ptr<MultiFab>& b = merged ? m_leg.mbb : bottom_b;
                if (!b) { b = std::make_unique<MultiFab>(op.make(0, mglev_bottom - off, 0)); }
                bb = b.get();
            }
        }
        Long top_cells = op.Geom(0, leg0 - off).Domain().numPts();
        const int ctas = (top_cells <= Long(m_leg_narrow_cells)) ? 1
                       : int(std::min<Long>(Long(m_leg_max_ctas), (top_cells / 2 + 511) / 512));
        // every device address and parameter the kernel bakes in
        std::size_t key = 1469598103934665603ull;
        for (int m = leg0; m <= mglev_bottom; ++m) {
            leg_mix(key, reinterpret_cast<std::size_t>(Cor(m).dataPtr())); leg_mix(key, reinterpret_cast<std::size_t>(Res(m).dataPtr()));
            leg_mix(key, reinterpret_cast<std::size_t>(Rescor(m).dataPtr())); leg_mix(key, op.graphKey(0, m - off));
        }
        if (bicg) {
            for (MultiFab* q : {tmp.p, tmp.r, tmp.rh, tmp.v, tmp.t}) { leg_mix(key, reinterpret_cast<std::size_t>(q->dataPtr())); }
            if (singular) { leg_mix(key, reinterpret_cast<std::size_t>(bb->dataPtr())); }
        }
        for (int v : {nu1, nu2, nuf, nub, bottom_maxiter, int(bicg), int(singular), leg0, nlev, int(merged), m_leg_narrow_cells}) { leg_mix(key, std::size_t(v)); }
        leg_mix(key, leg_bits(bottom_reltol)); leg_mix(key, leg_bits(bottom_abstol));
        leg_mix(key, leg_bits(op.getAScalar())); leg_mix(key, leg_bits(op.getBScalar()));
        auto& plan = m_leg.args[leg0];
        if (!plan.first || plan.second != key) {
            static b200mg_leg_args A;                         // ~10 KB: kept off the stack
            std::memset(&A, 0, sizeof(A));
            A.nlev = nlev; A.maxorder = op.getMaxOrder(); A.nu1 = nu1; A.nu2 = nu2; A.nuf = nuf; A.nub = nub;
            A.bottom_mode = bicg ? 0 : 1; A.singular = singular ? 1 : 0; A.maxiter = bottom_maxiter; A.narrow_cells = m_leg_narrow_cells;
            A.alpha = op.getAScalar(); A.volinv = singular ? op.bottomVolInv() : 0.0;
            A.eps_rel = bottom_reltol; A.eps_abs = bottom_abstol;
            if (bicg) {
                A.r = tmp.r->desc(0); A.p = tmp.p->desc(0); A.v = tmp.v->desc(0); A.t = tmp.t->desc(0); A.rh = tmp.rh->desc(0);
                if (singular) { A.bb = bb->desc(0); }
            }
            for (int l = 0; l < nlev; ++l) {
                const int m = leg0 + l;
                op.fillLegLevel(m - off, A.lev[l]);
                A.lev[l].cor = Cor(m).desc(0); A.lev[l].res = Res(m).desc(0); A.lev[l].rescor = Rescor(m).desc(0);
            }
            if (std::getenv("B200MG_LEG_STAMPS")) {               // tuning aid: per-phase SM clock stamps of the last launch
                if (!m_leg.d_stamps) { m_leg.d_stamps = static_cast<unsigned long long*>(The_Arena()->alloc(B200MG_LEG_MAX_STAMPS * sizeof(unsigned long long))); }
                A.stamps = m_leg.d_stamps;
            }
            if (!plan.first) { plan.first = static_cast<b200mg_leg_args*>(The_Arena()->alloc(sizeof(b200mg_leg_args))); }
            Gpu::htod_memcpy_async(plan.first, &A, sizeof(A));
            Gpu::streamSynchronize();                          // A is reused by the next plan
            plan.second = key;
        }
        if (!m_leg.d_log) {
            m_leg.d_log = static_cast<double*>(The_Arena()->alloc(2 * kLegLogMax * sizeof(double)));
            m_leg.h_log = static_cast<double*>(pinned_alloc(2 * kLegLogMax * sizeof(double)));
        }
        double* out = (bicg && m_leg.launches < kLegLogMax) ? m_leg.d_log + 2 * m_leg.launches : nullptr;
        const MultiFab* a = nullptr; Array<MultiFab const*, 3> b{{nullptr, nullptr, nullptr}}; Real alpha = 0.0, beta = 1.0;
        op.getLevelCoeffs(0, leg0 - off, a, b, alpha, beta);
        B200_KCALL(b200mg_coarse_leg(a ? 1 : 0, plan.first, out, ctas, Gpu::gpuStream()));
        if (bicg) { ++m_leg.launches; }
        if (bicg && verbose > 1 && out) {                      // the reference reports a failed bottom solve right away
            double h[2];
            Gpu::dtoh_memcpy_async(h, out, 2 * sizeof(double)); Gpu::streamSynchronize();
            if (int(h[0]) != 0) { Print0("MLMG: Bottom solve failed.\n"); }
        }
    }
    if (merged) { cor[0][leg0].ParallelCopy(m_leg.mcor[0], 0, 0, 1); }
}

// The leg kernel leaves {return code, iterations} of each bottom solve in device memory; one read-back per solve and one
// small all-reduce (only the rank that owns the coarse box knows the numbers) instead of a host round trip per V-cycle.
void MLMG::collectLegLog ()
{
    if (m_leg.d_stamps && m_leg.launches > 0) {
        std::vector<unsigned long long> h(B200MG_LEG_MAX_STAMPS);
        Gpu::dtoh_memcpy_async(h.data(), m_leg.d_stamps, h.size() * sizeof(unsigned long long)); Gpu::streamSynchronize();
        const int ns = int(std::min<unsigned long long>(h[0], B200MG_LEG_MAX_STAMPS));
        std::fprintf(stderr, "[leg stamps] %d records (level*16+phase: cycles since the previous record)\n", ns - 1);
        for (int i = 2; i < ns; ++i) {
            std::fprintf(stderr, " %d:%lld", int(h[i] >> 48), (long long)((h[i] & 0xffffffffffffull) - (h[i - 1] & 0xffffffffffffull)));
        }
        std::fprintf(stderr, "\n");
    }
    const int n = int(m_leg.slots.size());
    if (n == 0) { return; }
    std::vector<double> its(n, 0.0);
    if (m_leg.launches > 0) {
        const int m = std::min(m_leg.launches, kLegLogMax);
        Gpu::dtoh_memcpy_async(m_leg.h_log, m_leg.d_log, 2 * m * sizeof(double));
        Gpu::streamSynchronize();
        for (int i = 0; i < std::min(n, m); ++i) { its[i] = m_leg.h_log[2 * i + 1]; }
    }
    for (int i0 = 0; i0 < n; i0 += 32) { ParallelDescriptor::ReduceRealMax(its.data() + i0, std::min(32, n - i0)); }
    for (int i = 0; i < n; ++i) { m_niters_cg[m_leg.slots[i]] = int(its[i]); }
    m_leg.slots.clear(); m_leg.launches = 0;
}

void MLMG::getGradSolution (Vector<Array<MultiFab*, 3>> const& a_grad_sol)
{
    AMREX_ALWAYS_ASSERT_WITH_MESSAGE(int(sol.size()) == namrlevs && int(a_grad_sol.size()) == namrlevs, "getGradSolution: call solve first; one entry per AMR level");
    for (int alev = 0; alev <= finest_amr_lev; ++alev) { linop.compGrad(alev, a_grad_sol[alev], sol[alev]); }
}

void MLMG::getFluxes (Vector<Array<MultiFab*, 3>> const& a_flux)
{
    AMREX_ALWAYS_ASSERT_WITH_MESSAGE(int(sol.size()) == namrlevs && int(a_flux.size()) == namrlevs, "getFluxes: call solve first; one entry per AMR level");
    Vector<MultiFab*> ps(namrlevs);
    for (int alev = 0; alev < namrlevs; ++alev) { ps[alev] = &sol[alev]; }
    linop.getFluxes(a_flux, ps);
}

void MLMG::mgFcycle ()
{
    const int amrlev = 0;
    const int mg_bottom_lev = linop.NMGLevels(amrlev) - 1;
    for (int mglev = 1; mglev <= mg_bottom_lev; ++mglev) { linop.avgDownResMG(mglev, res[amrlev][mglev], res[amrlev][mglev - 1]); }
    bottomSolve();
    for (int mglev = mg_bottom_lev - 1; mglev >= 0; --mglev) {
        interpCorrection(amrlev, mglev);
        computeResOfCorrection(amrlev, mglev);
        MultiFab::Copy(res[amrlev][mglev], rescor[amrlev][mglev], 0, 0, 1, 0);
        std::swap(cor[amrlev][mglev], cor_hold[amrlev][mglev]);
        mgVcycle(amrlev, mglev);
        MultiFab::Add(cor[amrlev][mglev], cor_hold[amrlev][mglev], 0, 0, 1, 0);
    }
}

void MLMG::bottomSolve ()
{
    const double t0 = ParallelDescriptor::second();
    const int amrlev = 0;
    const int mglev = linop.NMGLevels(amrlev) - 1;
    Gpu::ProfScope prof_scope__(amrlev * 100 + mglev);
    MultiFab& x = cor[amrlev][mglev];
    MultiFab& b = res[amrlev][mglev];
    x.setVal(0.0);
    if (bottom_solver == BottomSolver::smoother) {
        bool skip_fillboundary = true;
        for (int i = 0; i < nuf; ++i) { linop.smooth(amrlev, mglev, x, b, skip_fillboundary, false, i > 0); skip_fillboundary = false; }
    } else {
        MultiFab* pb = &b;
        if (linop.isBottomSingular() && linop.getEnforceSingularSolvable()) {
            if (!bottom_b) { bottom_b = std::make_unique<MultiFab>(linop.make(amrlev, mglev, 0)); }
            MultiFab::Copy(*bottom_b, b, 0, 0, 1, 0);
            pb = bottom_b.get();
            makeSolvable(amrlev, mglev, *pb);
        }
        MLCGSolver::Type cg_type = (bottom_solver == BottomSolver::cg || bottom_solver == BottomSolver::cgbicg)
            ? MLCGSolver::Type::CG : MLCGSolver::Type::BiCGStab;
        int ret = bottomSolveWithCG(x, *pb, cg_type);
        if (ret != 0 && (bottom_solver == BottomSolver::cgbicg || bottom_solver == BottomSolver::bicgcg)) {
            cg_type = (bottom_solver == BottomSolver::cgbicg) ? MLCGSolver::Type::BiCGStab : MLCGSolver::Type::CG;
            cor[amrlev][mglev].setVal(0.0);
            ret = bottomSolveWithCG(x, *pb, cg_type);
            if (ret == 0) { bottom_solver = (cg_type == MLCGSolver::Type::CG) ? BottomSolver::cg : BottomSolver::bicgstab; }
        }
        if (ret != 0 && ret != 9) { cor[amrlev][mglev].setVal(0.0); }
        const int n = (ret == 0) ? nub : nuf;
        for (int i = 0; i < n; ++i) { linop.smooth(amrlev, mglev, x, b, false, false, i > 0); }
    }
    timer[2] += ParallelDescriptor::second() - t0;
}

int MLMG::bottomSolveWithCG (MultiFab& x, MultiFab const& b, MLCGSolver::Type type)
{
    if (!cg_solver) { cg_solver = std::make_unique<MLCGSolver>(linop); }
    cg_solver->setSolver(type);
    cg_solver->setVerbose(bottom_verbose);
    cg_solver->setMaxIter(bottom_maxiter);
    cg_solver->setInitSolnZeroed(true);
    const int ret = cg_solver->solve(x, b, bottom_reltol, bottom_abstol);
    if (ret != 0 && verbose > 1) { Print0("MLMG: Bottom solve failed.\n"); }
    m_niters_cg.push_back(cg_solver->getNumIters());
    return ret;
}

void MLMG::computeMLResidual (int amrlevmax)
{
    for (int alev = amrlevmax; alev >= 0; --alev) {
        const MultiFab* crse_bcdata = (alev > 0) ? &sol[alev - 1] : nullptr;
        linop.solutionResidual(alev, res[alev][0], sol[alev], rhs[alev], crse_bcdata);
        if (alev < finest_amr_lev) { linop.reflux(alev, res[alev][0], sol[alev], sol[alev + 1]); }
    }
}

bool MLMG::computeResidual (int alev, Real* resnorm)
{
    const MultiFab* crse_bcdata = (alev > 0) ? &sol[alev - 1] : nullptr;
    return linop.solutionResidual(alev, res[alev][0], sol[alev], rhs[alev], crse_bcdata, resnorm);
}

void MLMG::computeResOfCorrection (int amrlev, int mglev)
{
    linop.correctionResidual(amrlev, mglev, rescor[amrlev][mglev], cor[amrlev][mglev], res[amrlev][mglev], MLLinOp::BCMode::Homogeneous);
}

// AMReX_MLMG.H:1662-1690: coarse composite residual from the coarse solution and the fine correction
void MLMG::computeResWithCrseSolFineCor (int calev, int falev)
{
    const MultiFab* crse_bcdata = (calev > 0) ? &sol[calev - 1] : nullptr;
    linop.solutionResidual(calev, res[calev][0], sol[calev], rhs[calev], crse_bcdata);
    linop.correctionResidual(falev, 0, rescor[falev][0], cor[falev][0], res[falev][0], MLLinOp::BCMode::Homogeneous);
    MultiFab::Copy(res[falev][0], rescor[falev][0], 0, 0, 1, 0);
    linop.reflux(calev, res[calev][0], sol[calev], sol[falev]);
    linop.avgDownResAmr(calev, res[calev][0], res[falev][0]);
}

// AMReX_MLMG.H:1695-1714: fine_res -= L(fine_cor) with the coarse correction as boundary data
void MLMG::computeResWithCrseCorFineCor (int falev)
{
    linop.correctionResidual(falev, 0, rescor[falev][0], cor[falev][0], res[falev][0], MLLinOp::BCMode::Inhomogeneous, &cor[falev - 1][0]);
    MultiFab::Copy(res[falev][0], rescor[falev][0], 0, 0, 1, 0);
}

// AMReX_MLMG.H:1719-1747: trilinear interpolation of the coarse AMR correction through a ghosted coarsened-fine temporary
void MLMG::interpCorrection (int alev)
{
    if (int(cfine_amr.size()) <= alev) { cfine_amr.resize(alev + 1); }
    if (!cfine_amr[alev]) { cfine_amr[alev] = std::make_unique<MultiFab>(linop.makeCoarseAmr(alev, 1)); }
    MultiFab& cfine = *cfine_amr[alev];
    cfine.setVal(0.0);
    cfine.ParallelCopy(cor[alev - 1][0], 0, 0, 1, 0, 1, linop.Geom(alev - 1, 0).periodicity());
    linop.interpolationAmr(alev, cor[alev][0], cfine);
}

void MLMG::interpCorrection (int alev, int mglev)
{
    linop.interpAssign(alev, mglev, cor[alev][mglev], cor[alev][mglev + 1]);
}

void MLMG::addInterpCorrection (int alev, int mglev)
{
    MultiFab const& crse_cor = cor[alev][mglev + 1];
    MultiFab& fine_cor = cor[alev][mglev];
    const MultiFab* cmf = &crse_cor;
    if (!linop.isMFIterSafe(alev, mglev, mglev + 1)) {
        AMREX_ALWAYS_ASSERT(alev == 0);
        if (!cfine_mg[mglev]) {
            cfine_mg[mglev] = std::make_unique<MultiFab>(amrex::coarsen(linop.Grids(alev, mglev), 2), linop.DMap(alev, mglev), 1, 0);
        }
        cfine_mg[mglev]->ParallelCopy(crse_cor, 0, 0, 1);
        cmf = cfine_mg[mglev].get();
    }
    linop.interpolation(alev, mglev, fine_cor, *cmf);
}

Real MLMG::ResNormInf (int alev, bool local) { return linop.normInf(alev, res[alev][0], local); }

Real MLMG::MLResNormInf (int alevmax, bool local)
{
    Real r = 0.0;
    for (int alev = 0; alev <= alevmax; ++alev) { r = std::max(r, ResNormInf(alev, true)); }
    if (!local) { ParallelDescriptor::ReduceRealMax(&r, 1); }
    return r;
}

Real MLMG::MLRhsNormInf (bool local)
{
    Real r = 0.0;
    for (int alev = 0; alev <= finest_amr_lev; ++alev) { r = std::max(r, linop.normInf(alev, rhs[alev], true)); }
    if (!local) { ParallelDescriptor::ReduceRealMax(&r, 1); }
    return r;
}

void MLMG::makeSolvable ()
{
    auto const offset = linop.getSolvabilityOffset(0, 0, rhs[0]);
    for (int alev = 0; alev < namrlevs; ++alev) { linop.fixSolvabilityByOffset(alev, 0, rhs[alev], offset); }
}

void MLMG::makeSolvable (int amrlev, int mglev, MultiFab& mf)
{
    auto const offset = linop.getSolvabilityOffset(amrlev, mglev, mf);
    linop.fixSolvabilityByOffset(amrlev, mglev, mf, offset);
}

// MLMGT::compResidual (AMReX_MLMG.H:792-856): composite residual b - L(sol) on every AMR level
void MLMG::compResidual (Vector<MultiFab*> const& a_res, Vector<MultiFab*> const& a_sol, Vector<MultiFab const*> const& a_rhs)
{
    if (!linop_prepared) { linop.prepareForSolve(); linop_prepared = true; }
    else if (linop.needsUpdate()) { linop.update(); }
    Vector<MultiFab> s(namrlevs);
    for (int alev = 0; alev < namrlevs; ++alev) {
        s[alev] = linop.make(alev, 0, 1);
        MultiFab::Copy(s[alev], *a_sol[alev], 0, 0, 1, 0);
    }
    const bool innu = linop.hasInhomogNeumannBC();
    for (int alev = finest_amr_lev; alev >= 0; --alev) {
        const MultiFab* crse_bcdata = (alev > 0) ? &s[alev - 1] : nullptr;
        MultiFab rhstmp;                                         // AMReX_MLMG.H:825-835: the Neumann data enters through the rhs
        if (innu) {
            rhstmp = linop.make(alev, 0, 0);
            MultiFab::Copy(rhstmp, *a_rhs[alev], 0, 0, 1, 0);
            linop.applyInhomogNeumannTerm(alev, rhstmp);
        }
        linop.solutionResidual(alev, *a_res[alev], s[alev], innu ? rhstmp : *a_rhs[alev], crse_bcdata);
        if (alev < finest_amr_lev) {
            linop.reflux(alev, *a_res[alev], s[alev], s[alev + 1]);
            average_down(*a_res[alev + 1], *a_res[alev], 0, 1, linop.AMRRefRatio(alev));
        }
    }
    Gpu::streamSynchronize();
}

void MLMG::apply (Vector<MultiFab*> const& out, Vector<MultiFab*> const& in)
{
    if (!linop_prepared) { linop.prepareForSolve(); linop_prepared = true; }
    AMREX_ALWAYS_ASSERT_WITH_MESSAGE(namrlevs == 1, "apply: single level only for now");
    MultiFab s = linop.make(0, 0, 1);
    MultiFab::Copy(s, *in[0], 0, 0, 1, 0);
    if (linop.hasInhomogNeumannBC()) {
        // AMReX_MLMG.H:866-932: out = -(rh - L(in)) with rh = 0 + the Neumann boundary term
        MultiFab rh = linop.make(0, 0, 0);
        rh.setVal(0.0);
        linop.applyInhomogNeumannTerm(0, rh);
        linop.solutionResidual(0, *out[0], s, rh);
        out[0]->mult(-1.0);
        return;
    }
    linop.apply(0, 0, *out[0], s, MLLinOp::BCMode::Inhomogeneous, MLLinOp::StateMode::Solution, linop.m_bndry_sol[0].get());
}

} // namespace amrex
