// Device-side BiCGStab of ONE box by ONE CTA (reference algorithm: MLCGSolverT::solve_bicgstab, AMReX_MLCGSolver.H:98-273, with
// initial_vec_zeroed = true as MLMG's bottom solve calls it).  Shared by the stand-alone bottom kernel (bottom.cu) and the
// coarse-leg kernel (coarse_leg.cu), so both produce the same bits: per-cell arithmetic from stencil_math.cuh
// (-fmad=false), dot products summed in a fixed order (per-thread strided partial sums, warp shuffles, then the warp
// partials in index order).
//
// Requirements (checked by the callers): the box is the whole level, every face is a physical Dirichlet / Neumann /
// reflect-odd boundary or a periodic boundary of the box onto itself, p / r / sol carry one ghost cell, and the calling
// CTA has exactly kBottomThreads threads.
#ifndef AMREX_B200_BOTTOM_SOLVE_CUH_
#define AMREX_B200_BOTTOM_SOLVE_CUH_

#include "common.cuh"
#include "stencil_math.cuh"

namespace b200mg {

constexpr int kBcDirichletB = 101, kBcNeumannB = 102, kBcReflectOddB = 103;
constexpr int kBottomThreads = 512;

// Boundary description of a level that is ONE box covering the domain.  The host fills faces / mask / periodic; nxo / coef
// (interpolation order and Lagrange weights of the Dirichlet ghost value, one set per listed face) are computed once on
// the device by bc_prepare - the same poly_interp_coeff call k_apply_bc makes per launch, so the same bits.
struct BoxBC {
    int nfaces;
    b200mg_bcface faces[6];                 // faces with uncovered ghost cells (box == 0)
    b200mg_ifab mask[6];                    // [face orientation] mask slabs of the box
    int periodic[3];                        // direction wraps onto the box itself (single box == periodic domain)
    int nxo[6];
    double coef[6][4];
};

struct BottomArgs {
    b200mg_box vb;
    b200mg_fab sol, rhs, r, p, v, t, rh;
    b200mg_fab a, bx, by, bz;               // abec only
    int abec;
    double alpha, dhx, dhy, dhz;            // operator scalings of apply / normalize (beta*dxinv^2; Poisson: dxinv^2)
    BoxBC bc;
    int maxorder;
    double dxi[3];
    double eps_rel, eps_abs;
    int maxiter;
    double* out;                            // [0] return code, [1] iterations, [2] rnorm, [3] rnorm0
};

// threads tid = 0..5 fill the weights of listed face tid (callers synchronise afterwards)
__device__ __forceinline__ void bc_prepare (BoxBC& B, int maxorder, const double* dxi, int tid)
{
    if (tid < 6) {
        int NX = 0;
        double coef[4] = {0., 0., 0., 0.};
        if (tid < B.nfaces && B.faces[tid].bctype == kBcDirichletB) {
            const b200mg_bcface fc = B.faces[tid];
            NX = min(fc.blen + 1, maxorder);
            double xs[4] = {-fc.bcloc * dxi[fc.face % 3], 0.5, 1.5, 2.5};
            poly_interp_coeff(-0.5, xs, NX, coef);
        }
        B.nxo[tid] = NX;
        for (int m = 0; m < 4; ++m) { B.coef[tid][m] = coef[m]; }
    }
}

// homogeneous boundary fill of x's ghost faces by a team of nth threads (this thread: tid), no synchronisation inside:
// mllinop_apply_bc_* (AMReX_MLLinOp_K.H:14-327), as k_apply_bc with inhomog = 0; periodic directions copy the opposite
// valid plane (FillBoundary of the box onto itself, cross stencil).  Reads valid cells, writes face ghost cells only.
__device__ inline void bc_fill (const b200mg_box& vb, const BoxBC& B, const View<double>& x, int tid, int nth)
{
    for (int d = 0; d < 3; ++d) {
        if (!B.periodic[d]) { continue; }
        const long long st = (d == 0) ? 1 : ((d == 1) ? x.js : x.ks);
        const int n = vb.hi[d] - vb.lo[d] + 1;
        const int d1 = (d == 0) ? 1 : 0, d2 = (d == 2) ? 1 : 2;
        const int n1 = vb.hi[d1] - vb.lo[d1] + 1, n2 = vb.hi[d2] - vb.lo[d2] + 1;
        for (int q = tid; q < 2 * n1 * n2; q += nth) {
            const int side = q / (n1 * n2), qq = q - side * n1 * n2;
            int idx[3];
            idx[d] = side ? vb.hi[d] + 1 : vb.lo[d] - 1; idx[d1] = vb.lo[d1] + qq % n1; idx[d2] = vb.lo[d2] + qq / n1;
            double* p = x.ptr(idx[0], idx[1], idx[2]);
            *p = side ? p[-(long long)n * st] : p[(long long)n * st];
        }
    }
    for (int n = 0; n < B.nfaces; ++n) {
        const b200mg_bcface fc = B.faces[n];
        const auto mask = view(B.mask[fc.face]);
        const int d = fc.face % 3;
        const int s = (fc.face < 3) ? 1 : -1;
        const long long st = (d == 0) ? 1 : ((d == 1) ? x.js : x.ks);
        const int g = (fc.face < 3) ? vb.lo[d] - 1 : vb.hi[d] + 1;
        const int d1 = (d == 0) ? 1 : 0, d2 = (d == 2) ? 1 : 2;
        const int n1 = vb.hi[d1] - vb.lo[d1] + 1, n2 = vb.hi[d2] - vb.lo[d2] + 1;
        const int NX = B.nxo[n];
        for (int q = tid; q < n1 * n2; q += nth) {
            int idx[3];
            idx[d] = g; idx[d1] = vb.lo[d1] + q % n1; idx[d2] = vb.lo[d2] + q / n1;
            if (mask(idx[0], idx[1], idx[2]) > 0) {
                double* p = x.ptr(idx[0], idx[1], idx[2]);
                if (fc.bctype == kBcNeumannB) { *p = p[s * st]; }
                else if (fc.bctype == kBcReflectOddB) { *p = -p[s * st]; }
                else if (fc.bctype == kBcDirichletB) {
                    double tmp = 0.0;
                    for (int m = 1; m < NX; ++m) { tmp += p[m * s * st] * B.coef[n][m]; }
                    *p = tmp;
                }
            }
        }
    }
}

struct BottomCtx {
    int nx, ny, nz, ncells, tid;
    double* sh;                             // kBottomThreads / 32 + 1 doubles of shared memory
};

__device__ __forceinline__ BottomCtx make_bottom_ctx (const b200mg_box& vb, double* sh)
{
    BottomCtx C;
    C.nx = vb.hi[0] - vb.lo[0] + 1; C.ny = vb.hi[1] - vb.lo[1] + 1; C.nz = vb.hi[2] - vb.lo[2] + 1;
    C.ncells = C.nx * C.ny * C.nz; C.tid = int(threadIdx.x); C.sh = sh;
    return C;
}

// sum over the CTA, same value in every thread; two barriers, fixed summation order
__device__ __forceinline__ double cta_sum (double v, const BottomCtx& C)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { v += __shfl_down_sync(0xffffffffu, v, o); }
    if ((C.tid & 31) == 0) { C.sh[C.tid >> 5] = v; }
    __syncthreads();
    double r = 0.0;
    for (int w = 0; w < kBottomThreads / 32; ++w) { r += C.sh[w]; }
    __syncthreads();
    return r;
}
__device__ __forceinline__ double cta_max (double v, const BottomCtx& C)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { v = fmax(v, __shfl_down_sync(0xffffffffu, v, o)); }
    if ((C.tid & 31) == 0) { C.sh[C.tid >> 5] = v; }
    __syncthreads();
    double r = 0.0;
    for (int w = 0; w < kBottomThreads / 32; ++w) { r = fmax(r, C.sh[w]); }
    __syncthreads();
    return r;
}

template <class F>
__device__ __forceinline__ void bottom_for_cells (const b200mg_box& vb, const BottomCtx& C, F&& f)
{
    for (int c = C.tid; c < C.ncells; c += kBottomThreads) {
        const int i = c % C.nx, jk = c / C.nx;
        f(vb.lo[0] + i, vb.lo[1] + jk % C.ny, vb.lo[2] + jk / C.ny);
    }
}

__device__ __forceinline__ void bottom_fill_bc (const BottomArgs& A, const BoxBC& bc, const BottomCtx& C, const View<double>& x)
{
    bc_fill(A.vb, bc, x, C.tid, kBottomThreads);
    __syncthreads();
}

// y = normalize(L x): Lp.apply(Homogeneous) followed by Lp.normalize (mlabeclap_adotx + mlabeclap_normalize; Poisson: adotx only)
__device__ inline void bottom_apply_normalized (const BottomArgs& A, const BoxBC& bc, const BottomCtx& C, const View<double>& y, const View<double>& x)
{
    bottom_fill_bc(A, bc, C, x);
    const int js = int(x.js), ks = int(x.ks);
    if (A.abec) {
        const auto a = view(A.a); const auto bx = view(A.bx); const auto by = view(A.by); const auto bz = view(A.bz);
        bottom_for_cells(A.vb, C, [&] (int i, int j, int k) {
            const double* pc = x.ptr(i, j, k);
            const double av = a(i, j, k);
            const double bxm = bx(i, j, k), bxp = bx(i + 1, j, k), bym = by(i, j, k), byp = by(i, j + 1, k), bzm = bz(i, j, k), bzp = bz(i, j, k + 1);
            double v = adotx_abec_cell(*pc, pc[-1], pc[1], pc[-js], pc[js], pc[-ks], pc[ks], av, bxm, bxp, bym, byp, bzm, bzp,
                                       A.alpha, A.dhx, A.dhy, A.dhz);
            v /= A.alpha * av + A.dhx * (bxm + bxp) + A.dhy * (bym + byp) + A.dhz * (bzm + bzp);
            y(i, j, k) = v;
        });
    } else {
        bottom_for_cells(A.vb, C, [&] (int i, int j, int k) {
            const double* pc = x.ptr(i, j, k);
            y(i, j, k) = adotx_poisson_cell(*pc, pc[-1], pc[1], pc[-js], pc[js], pc[-ks], pc[ks], A.dhx, A.dhy, A.dhz);
        });
    }
    __syncthreads();
}

// The solve.  Every thread of the CTA returns the same {return code, iterations}; thread 0 also stores them in A.out.
// sol must be zero on entry (valid and ghost cells).
// bc: the box's boundary description with nxo / coef prepared (bc_prepare), e.g. a shared-memory copy of A.bc.
__device__ inline void bottom_bicgstab (const BottomArgs& A, const BoxBC& bc, const BottomCtx& C, int& ret_out, int& iter_out)
{
    const auto sol = view(A.sol); const auto rhs = view(A.rhs); const auto r = view(A.r); const auto p = view(A.p);
    const auto v = view(A.v); const auto t = view(A.t); const auto rh = view(A.rh);

    // p = 0, r = 0 on their whole (ghosted) boxes; r = rhs; normalize(r); rh = r
    {
        const int gx = C.nx + 2, gy = C.ny + 2, gz = C.nz + 2;
        for (int c = C.tid; c < gx * gy * gz; c += kBottomThreads) {
            const int i = A.vb.lo[0] - 1 + c % gx, j = A.vb.lo[1] - 1 + (c / gx) % gy, k = A.vb.lo[2] - 1 + c / (gx * gy);
            p(i, j, k) = 0.0; r(i, j, k) = 0.0;
        }
    }
    __syncthreads();
    double nrm = 0.0;
    if (A.abec) {
        const auto a = view(A.a); const auto bx = view(A.bx); const auto by = view(A.by); const auto bz = view(A.bz);
        bottom_for_cells(A.vb, C, [&] (int i, int j, int k) {
            double x = rhs(i, j, k);
            x /= A.alpha * a(i, j, k) + A.dhx * (bx(i, j, k) + bx(i + 1, j, k)) + A.dhy * (by(i, j, k) + by(i, j + 1, k))
                + A.dhz * (bz(i, j, k) + bz(i, j, k + 1));
            r(i, j, k) = x; rh(i, j, k) = x;
            nrm = fmax(nrm, fabs(x));
        });
    } else {
        bottom_for_cells(A.vb, C, [&] (int i, int j, int k) { const double x = rhs(i, j, k); r(i, j, k) = x; rh(i, j, k) = x; nrm = fmax(nrm, fabs(x)); });
    }
    double rnorm = cta_max(nrm, C);             // (the barriers inside also publish r / rh)
    const double rnorm0 = rnorm;
    int ret = 0, iter = 1;
    double rho_1 = 0.0, alpha = 0.0, omega = 0.0;

    if (!(rnorm0 == 0.0 || rnorm0 < A.eps_abs)) {
        for (; iter <= A.maxiter; ++iter) {
            double acc = 0.0;
            bottom_for_cells(A.vb, C, [&] (int i, int j, int k) { acc += rh(i, j, k) * r(i, j, k); });
            const double rho = cta_sum(acc, C);
            if (rho == 0.0) { ret = 1; break; }
            if (iter == 1) {
                bottom_for_cells(A.vb, C, [&] (int i, int j, int k) { p(i, j, k) = r(i, j, k); });
            } else {
                const double beta = (rho / rho_1) * (alpha / omega);
                const double momega = -omega;
                bottom_for_cells(A.vb, C, [&] (int i, int j, int k) {
                    double pv = momega * v(i, j, k) + 1.0 * p(i, j, k);      // Saxpy(p, -omega, v)
                    pv = 1.0 * r(i, j, k) + beta * pv;                        // Xpay(p, beta, r)
                    p(i, j, k) = pv;
                });
            }
            __syncthreads();
            bottom_apply_normalized(A, bc, C, v, p);
            acc = 0.0;
            bottom_for_cells(A.vb, C, [&] (int i, int j, int k) { acc += rh(i, j, k) * v(i, j, k); });
            const double rhTv = cta_sum(acc, C);
            if (rhTv != 0.0) { alpha = rho / rhTv; } else { ret = 2; break; }
            nrm = 0.0;
            {
                const double malpha = -alpha;
                bottom_for_cells(A.vb, C, [&] (int i, int j, int k) {
                    sol(i, j, k) = alpha * p(i, j, k) + 1.0 * sol(i, j, k);
                    const double rv = malpha * v(i, j, k) + 1.0 * r(i, j, k);
                    r(i, j, k) = rv;
                    nrm = fmax(nrm, fabs(rv));
                });
            }
            rnorm = cta_max(nrm, C);
            if (rnorm < A.eps_rel * rnorm0 || rnorm < A.eps_abs) { break; }
            bottom_apply_normalized(A, bc, C, t, r);
            double acc2 = 0.0; acc = 0.0;
            bottom_for_cells(A.vb, C, [&] (int i, int j, int k) { const double tv = t(i, j, k); acc += tv * tv; acc2 += tv * r(i, j, k); });
            const double tt = cta_sum(acc, C);
            const double tr = cta_sum(acc2, C);
            if (tt != 0.0) { omega = tr / tt; } else { ret = 3; break; }
            nrm = 0.0;
            {
                const double momega = -omega;
                bottom_for_cells(A.vb, C, [&] (int i, int j, int k) {
                    const double rold = r(i, j, k);
                    sol(i, j, k) = omega * rold + 1.0 * sol(i, j, k);
                    const double rv = momega * t(i, j, k) + 1.0 * rold;
                    r(i, j, k) = rv;
                    nrm = fmax(nrm, fabs(rv));
                });
            }
            rnorm = cta_max(nrm, C);
            if (rnorm < A.eps_rel * rnorm0 || rnorm < A.eps_abs) { break; }
            if (omega == 0.0) { ret = 4; break; }
            rho_1 = rho;
        }
        if (ret == 0 && rnorm > A.eps_rel * rnorm0 && rnorm > A.eps_abs) { ret = 8; }
        if ((ret == 0 || ret == 8) && (rnorm < rnorm0)) {
            if (ret == 8) { ret = 9; }
        } else {
            __syncthreads();
            const int gx = C.nx + 2, gy = C.ny + 2, gz = C.nz + 2;                // sol.setVal(0.0)
            for (int c = C.tid; c < gx * gy * gz; c += kBottomThreads) {
                sol(A.vb.lo[0] - 1 + c % gx, A.vb.lo[1] - 1 + (c / gx) % gy, A.vb.lo[2] - 1 + c / (gx * gy)) = 0.0;
            }
        }
    }
    if (C.tid == 0 && A.out != nullptr) { A.out[0] = double(ret); A.out[1] = double(iter); A.out[2] = rnorm; A.out[3] = rnorm0; }
    __syncthreads();
    ret_out = ret; iter_out = iter;
}

} // namespace b200mg
#endif
