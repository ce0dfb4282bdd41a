// amrex::GMRESMLMG - see AMReX_GMRESMLMG.H.  Reference: LinearSolvers/AMReX_GMRES_MLMG.H, LinearSolvers/AMReX_GMRES.H.
#include "AMReX_GMRESMLMG.H"

#include <cmath>
#include <iomanip>

#include <iostream>
#include <sstream>

namespace amrex {

namespace {
void Print0 (std::string const& s) { if (ParallelDescriptor::IOProcessor()) { std::cout << s << std::flush; } }
template <class... A> std::string cat (A const&... a) { std::ostringstream o; o << std::setprecision(10); (o << ... << a); return o.str(); }
}

GMRESMLMG::GMRESMLMG (MLMG& mlmg) : m_mlmg(&mlmg), m_linop(&mlmg.getLinOp())
{
    AMREX_ALWAYS_ASSERT_WITH_MESSAGE(m_linop->NAMRLevels() == 1, "GMRESMLMG: only single-level solves (AMReX_GMRES_MLMG.H:116)");
    m_mlmg->setVerbose(0);
    m_mlmg->setBottomVerbose(0);
    m_mlmg->prepareForGMRES();
    setRestartLength(m_restrtlen);
}

void GMRESMLMG::setRestartLength (int rl)
{
    m_restrtlen = rl;
    m_hh.assign(std::size_t(rl + 2) * (rl + 1), 0.0);    // Hessenberg, rows 0..rl+1, columns 0..rl
    m_grs.assign(rl + 2, 0.0); m_cc.assign(rl + 1, 0.0); m_ss.assign(rl + 1, 0.0);
    m_vv.clear();
}

MultiFab GMRESMLMG::makeVecRHS () const { return m_linop->make(0, 0, 0); }

MultiFab GMRESMLMG::makeVecLHS () const
{
    MultiFab mf = m_linop->make(0, 0, 1);
    mf.setBndry(0.0);
    return mf;
}

Real GMRESMLMG::norm2 (MultiFab const& mf) const { return std::sqrt(m_linop->xdoty(0, 0, mf, mf, false)); }

void GMRESMLMG::apply (MultiFab& lhs, MultiFab& rhs) const
{
    m_linop->apply(0, 0, lhs, rhs, MLLinOp::BCMode::Homogeneous, MLLinOp::StateMode::Correction);
}

// AMReX_GMRES_MLMG.H:188-214: m_precond_niters V-cycles on L z = r starting from z = 0
void GMRESMLMG::precond (MultiFab& lhs, MultiFab const& rhs)
{
    if (!m_use_precond) { MultiFab::Copy(lhs, rhs, 0, 0, 1, 0); return; }
    m_mlmg->prepareMGcycle();
    for (int icycle = 0; icycle < m_precond_niters; ++icycle) {
        if (icycle == 0) {
            MultiFab::Copy(m_mlmg->res[0][0], rhs, 0, 0, 1, 0);
        } else {
            m_mlmg->computeResOfCorrection(0, 0);
            MultiFab::Copy(m_mlmg->res[0][0], m_mlmg->rescor[0][0], 0, 0, 1, 0);
        }
        m_mlmg->mgVcycle(0, 0);
        if (icycle == 0) { MultiFab::Copy(lhs, m_mlmg->cor[0][0], 0, 0, 1, 0); }
        else { MultiFab::Saxpy(lhs, 1.0, m_mlmg->cor[0][0], 0, 0, 1, 0); }
    }
}

// AMReX_GMRES_MLMG.H:216-232
void GMRESMLMG::solve (MultiFab& a_sol, MultiFab const& a_rhs, RT a_tol_rel, RT a_tol_abs)
{
    m_rtol = a_tol_rel; m_atol = a_tol_abs;
    if (m_prop_zero) {
        MultiFab rhs = makeVecRHS();
        MultiFab::Copy(rhs, a_rhs, 0, 0, 1, 0);
        krylov(a_sol, rhs);
    } else {
        // the boundary data lives in the affine part: solve L(cor) = L(sol) - rhs with homogeneous BCs, sol -= cor
        MultiFab res = makeVecRHS();
        m_mlmg->apply({&res}, {&a_sol});
        MultiFab::Saxpy(res, -1.0, a_rhs, 0, 0, 1, 0);
        MultiFab cor = makeVecLHS();
        krylov(cor, res);
        MultiFab::Saxpy(a_sol, -1.0, cor, 0, 0, 1, 0);
    }
    Gpu::streamSynchronize();
}

// GMRES::solve (AMReX_GMRES.H:165-215)
void GMRESMLMG::krylov (MultiFab& x, MultiFab const& b)
{
    const double t0 = ParallelDescriptor::second();
    m_tmp_rhs = makeVecRHS();
    m_tmp_lhs = makeVecLHS();
    if (m_vv.empty()) { for (int i = 0; i < 2; ++i) { m_vv.emplace_back(makeVecRHS()); } }
    m_history.clear();

    RT rnorm0 = 0.0;
    MultiFab::Copy(m_vv[0], b, 0, 0, 1, 0);
    x.setVal(0.0);
    m_its = 0; m_status = -1;
    cycle(x, rnorm0);
    while (m_status == -1 && m_its < m_maxiter) {
        // restart: r = b - L x
        MultiFab::Copy(m_tmp_lhs, x, 0, 0, 1, 0);
        apply(m_tmp_rhs, m_tmp_lhs);
        MultiFab::Copy(m_vv[0], m_tmp_rhs, 0, 0, 1, 0);
        MultiFab::LinComb(m_vv[0], 1.0, b, -1.0, 0);
        cycle(x, rnorm0);
    }
    if (m_status == -1 && m_its >= m_maxiter) { m_status = 1; }
    m_tmp_rhs = MultiFab(); m_tmp_lhs = MultiFab(); m_vv.clear();
    if (m_verbose > 0) { Print0(cat("GMRES: Solve Time = ", ParallelDescriptor::second() - t0, "\n")); }
}

// one restart cycle (AMReX_GMRES.H:217-298)
void GMRESMLMG::cycle (MultiFab& x, RT& rnorm0)
{
    m_res = norm2(m_vv[0]);
    m_grs[0] = m_res;
    if (m_res == 0.0) { m_status = 0; return; }
    m_vv[0].mult(1.0 / m_res);
    if (m_its == 0) { rnorm0 = m_res; m_history.push_back(m_res); }
    m_status = converged(rnorm0, m_res) ? 0 : -1;

    auto report = [&] { Print0(cat("GMRES: iter = ", m_its, ", residual = ", m_res, ", ", m_res / rnorm0, " (rel.)\n")); };
    int it = 0;
    while (it < m_restrtlen && m_its < m_maxiter) {
        if (m_verbose > 1) { report(); }
        if (m_status == 0) { break; }
        while (int(m_vv.size()) < it + 2) { m_vv.emplace_back(makeVecRHS()); }

        precond(m_tmp_lhs, m_vv[it]);
        apply(m_vv[it + 1], m_tmp_lhs);
        orthogonalize(it);

        const RT tt = norm2(m_vv[it + 1]);
        const bool happyend = (tt < 1.e-99);
        if (!happyend) { m_vv[it + 1].mult(1.0 / tt); }
        hh(it + 1, it) = tt;
        rotate(it, happyend);

        ++it; ++m_its;
        m_history.push_back(m_res);
        m_status = converged(rnorm0, m_res) ? 0 : -1;
        if (happyend) { break; }
    }
    if (m_verbose > 1 && (m_status != 0 || m_its >= m_maxiter)) { report(); }
    buildSolution(x, it - 1);
}

// classical Gram-Schmidt, twice (AMReX_GMRES.H:322-348); each pass = one batched reduction + one batched update
void GMRESMLMG::orthogonalize (int it)
{
    MultiFab& w = m_vv[it + 1];
    Vector<MultiFab const*> basis(it + 1);
    for (int j = 0; j <= it; ++j) { basis[j] = &m_vv[j]; hh(j, it) = 0.0; }
    Vector<RT> lhh(it + 1), neg(it + 1);
    for (int pass = 0; pass < 2; ++pass) {
        MultiFab::MultiDot(w, basis, lhh.data(), false);
        for (int j = 0; j <= it; ++j) { neg[j] = -lhh[j]; hh(j, it) += lhh[j]; }
        MultiFab::MultiSaxpy(w, basis, neg.data());
    }
}

// apply the previous Givens rotations to column `it`, then the new one (AMReX_GMRES.H:350-376)
void GMRESMLMG::rotate (int it, bool happyend)
{
    for (int j = 1; j <= it; ++j) {
        const RT tt = hh(j - 1, it);
        hh(j - 1, it) = m_cc[j - 1] * tt + m_ss[j - 1] * hh(j, it);
        hh(j, it) = m_cc[j - 1] * hh(j, it) - m_ss[j - 1] * tt;
    }
    if (happyend) { m_res = 0.0; return; }
    const RT tt = std::sqrt(hh(it, it) * hh(it, it) + hh(it + 1, it) * hh(it + 1, it));
    m_cc[it] = hh(it, it) / tt;
    m_ss[it] = hh(it + 1, it) / tt;
    m_grs[it + 1] = -(m_ss[it] * m_grs[it]);
    m_grs[it] = m_cc[it] * m_grs[it];
    hh(it, it) = m_cc[it] * hh(it, it) + m_ss[it] * hh(it + 1, it);
    m_res = std::abs(m_grs[it + 1]);
}

// back substitution, x += M^-1 (V y)  (AMReX_GMRES.H:378-400)
void GMRESMLMG::buildSolution (MultiFab& x, int it)
{
    if (it < 0) { return; }
    m_grs[it] = (hh(it, it) != 0.0) ? m_grs[it] / hh(it, it) : 0.0;
    for (int ii = 1; ii <= it; ++ii) {
        const int k = it - ii;
        RT tt = m_grs[k];
        for (int j = k + 1; j <= it; ++j) { tt -= hh(k, j) * m_grs[j]; }
        m_grs[k] = tt / hh(k, k);
    }
    m_tmp_rhs.setVal(0.0);
    Vector<MultiFab const*> basis(it + 1);
    for (int j = 0; j <= it; ++j) { basis[j] = &m_vv[j]; }
    MultiFab::MultiSaxpy(m_tmp_rhs, basis, m_grs.data());
    precond(m_tmp_lhs, m_tmp_rhs);
    MultiFab::Saxpy(x, 1.0, m_tmp_lhs, 0, 0, 1, 0);
}

} // namespace amrex
