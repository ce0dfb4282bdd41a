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
// Requirements (checked by the caller): ONE box that is the whole level; every face a physical Dirichlet / Neumann /
// reflect-odd boundary or a periodic boundary of the box onto itself; p and r carry one ghost cell.
// The device code lives in bottom_solve.cuh (shared with the coarse-leg kernel, coarse_leg.cu).
#include "bottom_solve.cuh"

using namespace b200mg;

namespace {

__global__ void __launch_bounds__(kBottomThreads, 1)
k_bottom_bicgstab (const __grid_constant__ BottomArgs A)
{
    __shared__ double sh[kBottomThreads / 32 + 1];
    __shared__ BoxBC bc;
    if (threadIdx.x == 0) { bc = A.bc; }
    __syncthreads();
    bc_prepare(bc, A.maxorder, A.dxi, int(threadIdx.x));
    __syncthreads();
    const BottomCtx C = make_bottom_ctx(A.vb, sh);
    int ret, iter;
    bottom_bicgstab(A, bc, C, ret, iter);
}

} // namespace

extern "C" {

int b200mg_bottom_bicgstab (int abec, const b200mg_box* h_vbox,
                            const b200mg_fab* h_sol, const b200mg_fab* h_rhs, const b200mg_fab* h_r, const b200mg_fab* h_p,
                            const b200mg_fab* h_v, const b200mg_fab* h_t, const b200mg_fab* h_rh,
                            const b200mg_fab* h_a, const b200mg_fab* h_bx, const b200mg_fab* h_by, const b200mg_fab* h_bz,
                            double alpha, double dhx, double dhy, double dhz,
                            int nfaces, const b200mg_bcface* h_faces, const b200mg_ifab* h_mask, const int* periodic, int maxorder,
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
    A.bc.nfaces = nfaces;
    for (int n = 0; n < nfaces; ++n) { A.bc.faces[n] = h_faces[n]; if (h_faces[n].box != 0) { return int(cudaErrorInvalidValue); } }
    for (int f = 0; f < 6; ++f) { A.bc.mask[f] = h_mask[f]; }
    for (int d = 0; d < 3; ++d) { A.bc.periodic[d] = periodic ? periodic[d] : 0; }
    A.maxorder = maxorder;
    A.dxi[0] = dxinv0; A.dxi[1] = dxinv1; A.dxi[2] = dxinv2;
    A.eps_rel = eps_rel; A.eps_abs = eps_abs; A.maxiter = maxiter;
    A.out = d_out;
    k_bottom_bicgstab<<<1, kBottomThreads, 0, s>>>(A);
    return last_error();
}

} // extern "C"
