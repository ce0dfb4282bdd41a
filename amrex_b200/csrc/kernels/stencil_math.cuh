// Per-cell arithmetic of the cell-centred operators, written once and shared by every kernel variant
// (plain colour sweep, surface shell, fused two-colour pass) so that all of them produce identical bits.
// The association order of every expression follows the reference's kernels (it is part of the parity
// contract); compile with -fmad=false to keep it.
#ifndef AMREX_B200_STENCIL_MATH_CUH_
#define AMREX_B200_STENCIL_MATH_CUH_

#include "common.cuh"

namespace b200mg {

struct FaceCoefs { double c[6]; };

// cf0..cf5 of the reference GSRB kernels: slab value f at the cell if the cell lies on that face of its
// box AND the ghost cell beyond the face is an uncovered (mask>0) boundary cell.
__device__ __forceinline__ FaceCoefs
face_coefs (int i, int j, int k, const b200mg_box& vb, const b200mg_fab* f6, const b200mg_ifab* m6)
{
    FaceCoefs r;
#pragma unroll
    for (int n = 0; n < 6; ++n) { r.c[n] = 0.0; }
    if (i == vb.lo[0]) { if (view(m6[0])(i - 1, j, k) > 0) { r.c[0] = view(f6[0])(i, j, k); } }
    if (j == vb.lo[1]) { if (view(m6[1])(i, j - 1, k) > 0) { r.c[1] = view(f6[1])(i, j, k); } }
    if (k == vb.lo[2]) { if (view(m6[2])(i, j, k - 1) > 0) { r.c[2] = view(f6[2])(i, j, k); } }
    if (i == vb.hi[0]) { if (view(m6[3])(i + 1, j, k) > 0) { r.c[3] = view(f6[3])(i, j, k); } }
    if (j == vb.hi[1]) { if (view(m6[4])(i, j + 1, k) > 0) { r.c[4] = view(f6[4])(i, j, k); } }
    if (k == vb.hi[2]) { if (view(m6[5])(i, j, k + 1) > 0) { r.c[5] = view(f6[5])(i, j, k); } }
    return r;
}

__device__ __forceinline__ bool on_surface (int i, int j, int k, const b200mg_box& vb)
{
    return i == vb.lo[0] || i == vb.hi[0] || j == vb.lo[1] || j == vb.hi[1] || k == vb.lo[2] || k == vb.hi[2];
}

constexpr double kOmega = 1.15;   // over-relaxation factor of both GSRB kernels

// abec_gsrb, AMReX_MLABecLap_3D_K.H:225-263.  b*m / b*p: face coefficients on the low / high side.
__device__ __forceinline__ double
gsrb_abec_cell (double phi, double pxm, double pxp, double pym, double pyp, double pzm, double pzp,
                double rhs, double a, double bxm, double bxp, double bym, double byp, double bzm, double bzp,
                double cf0, double cf1, double cf2, double cf3, double cf4, double cf5,
                double alpha, double dhx, double dhy, double dhz)
{
    const double gamma = alpha * a + dhx * (bxm + bxp) + dhy * (bym + byp) + dhz * (bzm + bzp);
    const double g_m_d = gamma - (dhx * (bxm * cf0 + bxp * cf3) + dhy * (bym * cf1 + byp * cf4) + dhz * (bzm * cf2 + bzp * cf5));
    const double rho = dhx * (bxm * pxm + bxp * pxp) + dhy * (bym * pym + byp * pyp) + dhz * (bzm * pzm + bzp * pzp);
    const double res = rhs - (gamma * phi - rho);
    return phi + kOmega / g_m_d * res;
}

// interior fast path: all cf == 0  =>  g_m_d = gamma - (dhx*(0+0) + ...) = gamma - 0 = gamma exactly
__device__ __forceinline__ double
gsrb_abec_cell_interior (double phi, double pxm, double pxp, double pym, double pyp, double pzm, double pzp,
                         double rhs, double a, double bxm, double bxp, double bym, double byp, double bzm, double bzp,
                         double alpha, double dhx, double dhy, double dhz)
{
    const double gamma = alpha * a + dhx * (bxm + bxp) + dhy * (bym + byp) + dhz * (bzm + bzp);
    const double rho = dhx * (bxm * pxm + bxp * pxp) + dhy * (bym * pym + byp * pyp) + dhz * (bzm * pzm + bzp * pzp);
    const double res = rhs - (gamma * phi - rho);
    return phi + kOmega / gamma * res;
}

// mlpoisson_gsrb, AMReX_MLPoisson_3D_K.H:171-195
__device__ __forceinline__ double
gsrb_poisson_cell (double phi, double pxm, double pxp, double pym, double pyp, double pzm, double pzp,
                   double rhs, double cf0, double cf1, double cf2, double cf3, double cf4, double cf5,
                   double dhx, double dhy, double dhz)
{
    const double gamma = -2.0 * (dhx + dhy + dhz);
    const double g_m_d = gamma + dhx * (cf0 + cf3) + dhy * (cf1 + cf4) + dhz * (cf2 + cf5);
    const double res = rhs - gamma * phi - dhx * (pxm + pxp) - dhy * (pym + pyp) - dhz * (pzm + pzp);
    return phi + kOmega / g_m_d * res;
}

// mlabeclap_adotx, AMReX_MLABecLap_3D_K.H:21-27 (dh* = beta*dxinv^2)
__device__ __forceinline__ double
adotx_abec_cell (double x, double xxm, double xxp, double xym, double xyp, double xzm, double xzp,
                 double a, double bxm, double bxp, double bym, double byp, double bzm, double bzp,
                 double alpha, double dhx, double dhy, double dhz)
{
    return alpha * a * x
        - dhx * (bxp * (xxp - x) - bxm * (x - xxm))
        - dhy * (byp * (xyp - x) - bym * (x - xym))
        - dhz * (bzp * (xzp - x) - bzm * (x - xzm));
}

// mlpoisson_adotx, AMReX_MLPoisson_3D_K.H:13-15
__device__ __forceinline__ double
adotx_poisson_cell (double x, double xxm, double xxp, double xym, double xyp, double xzm, double xzp,
                    double dhx, double dhy, double dhz)
{
    return dhx * (xxm - 2.0 * x + xxp) + dhy * (xym - 2.0 * x + xyp) + dhz * (xzm - 2.0 * x + xzp);
}

// Lagrange weights at xInt for nodes x[0..N) (poly_interp_coeff, Src/Boundary/AMReX_LOUtil_K.H:24-36)
__device__ __forceinline__ void poly_interp_coeff (double xInt, const double* x, int N, double* c)
{
    for (int j = 0; j < N; ++j) {
        double num = 1.0, den = 1.0;
        for (int i = 0; i < N; ++i) {
            if (i != j) { num *= xInt - x[i]; den *= x[j] - x[i]; }
        }
        c[j] = num / den;
    }
}

} // namespace b200mg
#endif
