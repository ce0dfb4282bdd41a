// Whole BiCGStab bottom solve in ONE single-CTA kernel (reference algorithm: MLCGSolverT::solve_bicgstab,
// AMReX_MLCGSolver.H:98-273, with initial_vec_zeroed = true as MLMG's bottom solve calls it).
//
// On the bottom MG level (one box of 2^3 .. 32^3 cells after agglomeration) the launch-per-operation schedule costs
// ~22 launches and 6 host round trips (dot products, norms) per BiCGStab iteration for a few thousand flops.  Here
// the iteration loop, its scalar control flow, the homogeneous boundary fill, the operator apply, the normalisation,
// the vector updates and the reductions all run inside one CTA; the host reads back {return code, iterations} once.
// Per-cell arithmetic is the same code as the multi-kernel path (stencil_math.cuh, -fmad=false), only the summation
// order of the dot products differs (fixed order: warp shuffles, then the warp partials in index order).
//
// Requirements (checked by the caller): ONE box that is the whole level, no periodic direction (no self halo),
// every face a physical Dirichlet / Neumann / reflect-odd boundary, p and r carry one ghost cell.
#include "common.cuh"
#include "stencil_math.cuh"

using namespace b200mg;

namespace {

constexpr int kBcDirichlet = 101, kBcNeumann = 102, kBcReflectOdd = 103;
constexpr int kThreads = 512;

struct BottomArgs {
    b200mg_box vb;
    b200mg_fab sol, rhs, r, p, v, t, rh;
    b200mg_fab a, bx, by, bz;               // abec only
    int abec;
    double alpha, dhx, dhy, dhz;            // operator scalings of apply / normalize (beta*dxinv^2; Poisson: dxinv^2)
    int nfaces;
    b200mg_bcface faces[6];
    const b200mg_ifab* mask;                // [face] table of the box
    int maxorder;
    double dxi[3];
    double eps_rel, eps_abs;
    int maxiter;
    double* out;                            // [0] return code, [1] iterations, [2] rnorm, [3] rnorm0
};

struct Ctx {
    int nx, ny, nz, ncells, tid;
    double* sh;                             // kThreads / 32 + 1 doubles
};

// sum over the CTA, same value in every thread; two barriers, fixed summation order
__device__ __forceinline__ double cta_sum (double v, const Ctx& C)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { v += __shfl_down_sync(0xffffffffu, v, o); }
    if ((C.tid & 31) == 0) { C.sh[C.tid >> 5] = v; }
    __syncthreads();
    double r = 0.0;
    for (int w = 0; w < kThreads / 32; ++w) { r += C.sh[w]; }
    __syncthreads();
    return r;
}
__device__ __forceinline__ double cta_max (double v, const Ctx& C)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { v = fmax(v, __shfl_down_sync(0xffffffffu, v, o)); }
    if ((C.tid & 31) == 0) { C.sh[C.tid >> 5] = v; }
    __syncthreads();
    double r = 0.0;
    for (int w = 0; w < kThreads / 32; ++w) { r = fmax(r, C.sh[w]); }
    __syncthreads();
    return r;
}

template <class F>
__device__ __forceinline__ void for_cells (const BottomArgs& A, const Ctx& C, F&& f)
{
    for (int c = C.tid; c < C.ncells; c += kThreads) {
        const int i = c % C.nx, jk = c / C.nx;
        f(A.vb.lo[0] + i, A.vb.lo[1] + jk % C.ny, A.vb.lo[2] + jk / C.ny);
    }
}

// homogeneous boundary fill of x's ghost faces: mllinop_apply_bc_* (AMReX_MLLinOp_K.H:14-327), as k_apply_bc with inhomog = 0
__device__ void fill_bc (const BottomArgs& A, const Ctx& C, const View<double>& x)
{
    for (int n = 0; n < A.nfaces; ++n) {
        const b200mg_bcface fc = A.faces[n];
        const auto mask = view(A.mask[fc.face]);
        const int d = fc.face % 3;
        const int s = (fc.face < 3) ? 1 : -1;
        const long long st = (d == 0) ? 1 : ((d == 1) ? x.js : x.ks);
        const int g = (fc.face < 3) ? A.vb.lo[d] - 1 : A.vb.hi[d] + 1;
        const int d1 = (d == 0) ? 1 : 0, d2 = (d == 2) ? 1 : 2;
        const int n1 = A.vb.hi[d1] - A.vb.lo[d1] + 1, n2 = A.vb.hi[d2] - A.vb.lo[d2] + 1;
        int NX = 0;
        double coef[4] = {0., 0., 0., 0.};
        if (fc.bctype == kBcDirichlet) {
            NX = min(fc.blen + 1, A.maxorder);
            double xs[4] = {-fc.bcloc * A.dxi[d], 0.5, 1.5, 2.5};
            poly_interp_coeff(-0.5, xs, NX, coef);
        }
        for (int q = C.tid; q < n1 * n2; q += kThreads) {
            int idx[3];
            idx[d] = g; idx[d1] = A.vb.lo[d1] + q % n1; idx[d2] = A.vb.lo[d2] + q / n1;
            if (mask(idx[0], idx[1], idx[2]) > 0) {
                double* p = x.ptr(idx[0], idx[1], idx[2]);
                if (fc.bctype == kBcNeumann) { *p = p[s * st]; }
                else if (fc.bctype == kBcReflectOdd) { *p = -p[s * st]; }
                else if (fc.bctype == kBcDirichlet) {
                    double tmp = 0.0;
                    for (int m = 1; m < NX; ++m) { tmp += p[m * s * st] * coef[m]; }
                    *p = tmp;
                }
            }
        }
    }
    __syncthreads();
}

// y = normalize(L x): Lp.apply(Homogeneous) followed by Lp.normalize (mlabeclap_adotx + mlabeclap_normalize; Poisson: adotx only)
__device__ void apply_normalized (const BottomArgs& A, const Ctx& C, const View<double>& y, const View<double>& x)
{
    fill_bc(A, C, x);
    const int js = int(x.js), ks = int(x.ks);
    if (A.abec) {
        const auto a = view(A.a); const auto bx = view(A.bx); const auto by = view(A.by); const auto bz = view(A.bz);
        for_cells(A, C, [&] (int i, int j, int k) {
            const double* pc = x.ptr(i, j, k);
            const double av = a(i, j, k);
            const double bxm = bx(i, j, k), bxp = bx(i + 1, j, k), bym = by(i, j, k), byp = by(i, j + 1, k), bzm = bz(i, j, k), bzp = bz(i, j, k + 1);
            double v = adotx_abec_cell(*pc, pc[-1], pc[1], pc[-js], pc[js], pc[-ks], pc[ks], av, bxm, bxp, bym, byp, bzm, bzp,
                                       A.alpha, A.dhx, A.dhy, A.dhz);
            v /= A.alpha * av + A.dhx * (bxm + bxp) + A.dhy * (bym + byp) + A.dhz * (bzm + bzp);
            y(i, j, k) = v;
        });
    } else {
        for_cells(A, C, [&] (int i, int j, int k) {
            const double* pc = x.ptr(i, j, k);
            y(i, j, k) = adotx_poisson_cell(*pc, pc[-1], pc[1], pc[-js], pc[js], pc[-ks], pc[ks], A.dhx, A.dhy, A.dhz);
        });
    }
    __syncthreads();
}

__global__ void __launch_bounds__(kThreads, 1)
k_bottom_bicgstab (const __grid_constant__ BottomArgs A)
{
    __shared__ double sh[kThreads / 32 + 1];
    Ctx C;
    C.nx = A.vb.hi[0] - A.vb.lo[0] + 1; C.ny = A.vb.hi[1] - A.vb.lo[1] + 1; C.nz = A.vb.hi[2] - A.vb.lo[2] + 1;
    C.ncells = C.nx * C.ny * C.nz; C.tid = int(threadIdx.x); C.sh = sh;
    const auto sol = view(A.sol); const auto rhs = view(A.rhs); const auto r = view(A.r); const auto p = view(A.p);
    const auto v = view(A.v); const auto t = view(A.t); const auto rh = view(A.rh);

    // p = 0, r = 0 on their whole (ghosted) boxes; r = rhs; normalize(r); rh = r
    {
        const int gx = C.nx + 2, gy = C.ny + 2, gz = C.nz + 2;
        for (int c = C.tid; c < gx * gy * gz; c += kThreads) {
            const int i = A.vb.lo[0] - 1 + c % gx, j = A.vb.lo[1] - 1 + (c / gx) % gy, k = A.vb.lo[2] - 1 + c / (gx * gy);
            p(i, j, k) = 0.0; r(i, j, k) = 0.0;
        }
    }
    __syncthreads();
    double nrm = 0.0;
    if (A.abec) {
        const auto a = view(A.a); const auto bx = view(A.bx); const auto by = view(A.by); const auto bz = view(A.bz);
        for_cells(A, C, [&] (int i, int j, int k) {
            double x = rhs(i, j, k);
            x /= A.alpha * a(i, j, k) + A.dhx * (bx(i, j, k) + bx(i + 1, j, k)) + A.dhy * (by(i, j, k) + by(i, j + 1, k))
                + A.dhz * (bz(i, j, k) + bz(i, j, k + 1));
            r(i, j, k) = x; rh(i, j, k) = x;
            nrm = fmax(nrm, fabs(x));
        });
    } else {
        for_cells(A, C, [&] (int i, int j, int k) { const double x = rhs(i, j, k); r(i, j, k) = x; rh(i, j, k) = x; nrm = fmax(nrm, fabs(x)); });
    }
    double rnorm = cta_max(nrm, C);             // (the barriers inside also publish r / rh)
    const double rnorm0 = rnorm;
    int ret = 0, iter = 1;
    double rho_1 = 0.0, alpha = 0.0, omega = 0.0;

    if (!(rnorm0 == 0.0 || rnorm0 < A.eps_abs)) {
        for (; iter <= A.maxiter; ++iter) {
            double acc = 0.0;
            for_cells(A, C, [&] (int i, int j, int k) { acc += rh(i, j, k) * r(i, j, k); });
            const double rho = cta_sum(acc, C);
            if (rho == 0.0) { ret = 1; break; }
            if (iter == 1) {
                for_cells(A, C, [&] (int i, int j, int k) { p(i, j, k) = r(i, j, k); });
            } else {
                const double beta = (rho / rho_1) * (alpha / omega);
                const double momega = -omega;
                for_cells(A, C, [&] (int i, int j, int k) {
                    double pv = momega * v(i, j, k) + 1.0 * p(i, j, k);      // Saxpy(p, -omega, v)
                    pv = 1.0 * r(i, j, k) + beta * pv;                        // Xpay(p, beta, r)
                    p(i, j, k) = pv;
                });
            }
            __syncthreads();
            apply_normalized(A, C, v, p);
            acc = 0.0;
            for_cells(A, C, [&] (int i, int j, int k) { acc += rh(i, j, k) * v(i, j, k); });
            const double rhTv = cta_sum(acc, C);
            if (rhTv != 0.0) { alpha = rho / rhTv; } else { ret = 2; break; }
            nrm = 0.0;
            {
                const double malpha = -alpha;
                for_cells(A, C, [&] (int i, int j, int k) {
                    sol(i, j, k) = alpha * p(i, j, k) + 1.0 * sol(i, j, k);
                    const double rv = malpha * v(i, j, k) + 1.0 * r(i, j, k);
                    r(i, j, k) = rv;
                    nrm = fmax(nrm, fabs(rv));
                });
            }
            rnorm = cta_max(nrm, C);
            if (rnorm < A.eps_rel * rnorm0 || rnorm < A.eps_abs) { break; }
            apply_normalized(A, C, t, r);
            double acc2 = 0.0; acc = 0.0;
            for_cells(A, C, [&] (int i, int j, int k) { const double tv = t(i, j, k); acc += tv * tv; acc2 += tv * r(i, j, k); });
            const double tt = cta_sum(acc, C);
            const double tr = cta_sum(acc2, C);
            if (tt != 0.0) { omega = tr / tt; } else { ret = 3; break; }
            nrm = 0.0;
            {
                const double momega = -omega;
                for_cells(A, C, [&] (int i, int j, int k) {
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
            for (int c = C.tid; c < gx * gy * gz; c += kThreads) {
                sol(A.vb.lo[0] - 1 + c % gx, A.vb.lo[1] - 1 + (c / gx) % gy, A.vb.lo[2] - 1 + c / (gx * gy)) = 0.0;
            }
        }
    }
    if (C.tid == 0) { A.out[0] = double(ret); A.out[1] = double(iter); A.out[2] = rnorm; A.out[3] = rnorm0; }
}

} // namespace

extern "C" {

int b200mg_bottom_bicgstab (int abec, const b200mg_box* h_vbox,
                            const b200mg_fab* h_sol, const b200mg_fab* h_rhs, const b200mg_fab* h_r, const b200mg_fab* h_p,
                            const b200mg_fab* h_v, const b200mg_fab* h_t, const b200mg_fab* h_rh,
                            const b200mg_fab* h_a, const b200mg_fab* h_bx, const b200mg_fab* h_by, const b200mg_fab* h_bz,
                            double alpha, double dhx, double dhy, double dhz,
                            int nfaces, const b200mg_bcface* h_faces, const b200mg_ifab* d_mask, int maxorder,
                            double dxinv0, double dxinv1, double dxinv2, double eps_rel, double eps_abs, int maxiter,
                            double* d_out, cudaStream_t s)
{
    if (nfaces < 0 || nfaces > 6) { return int(cudaErrorInvalidValue); }
    const long long nc = (long long)(h_vbox->hi[0] - h_vbox->lo[0] + 1) * (h_vbox->hi[1] - h_vbox->lo[1] + 1) * (h_vbox->hi[2] - h_vbox->lo[2] + 1);
    if (nc <= 0 || nc > 32768) { return int(cudaErrorInvalidValue); }
    // p, r, sol: one ghost cell around the valid box
    for (const b200mg_fab* f : {h_sol, h_r, h_p}) {
        for (int d = 0; d < 3; ++d) { if (f->lo[d] > h_vbox->lo[d] - 1 || f->hi[d] < h_vbox->hi[d] + 1) { return int(cudaErrorInvalidValue); } }
    }
    static BottomArgs A;                     // kept off the stack; copied by value at the launch
    A.vb = *h_vbox;
    A.sol = *h_sol; A.rhs = *h_rhs; A.r = *h_r; A.p = *h_p; A.v = *h_v; A.t = *h_t; A.rh = *h_rh;
    A.abec = abec ? 1 : 0;
    if (abec) { A.a = *h_a; A.bx = *h_bx; A.by = *h_by; A.bz = *h_bz; }
    A.alpha = alpha; A.dhx = dhx; A.dhy = dhy; A.dhz = dhz;
    A.nfaces = nfaces;
    for (int n = 0; n < nfaces; ++n) { A.faces[n] = h_faces[n]; if (h_faces[n].box != 0) { return int(cudaErrorInvalidValue); } }
    A.mask = d_mask; A.maxorder = maxorder;
    A.dxi[0] = dxinv0; A.dxi[1] = dxinv1; A.dxi[2] = dxinv2;
    A.eps_rel = eps_rel; A.eps_abs = eps_abs; A.maxiter = maxiter;
    A.out = d_out;
    k_bottom_bicgstab<<<1, kThreads, 0, s>>>(A);
    return last_error();
}

} // extern "C"
