// Colour-sweep smoother, operator apply / residual, normalisation and boundary-condition kernels.
// Reference rows: K1/K2 (GSRB), K4/K5 (adotx), K9 (apply_bc), K11 (comp_interp_coef0), K13 (normalize).
#include "common.cuh"
#include "stencil_math.cuh"

using namespace b200mg;

namespace {

struct AbecArgs {
    const b200mg_fab *phi, *rhs, *a, *bx, *by, *bz, *f;
    const b200mg_ifab* m;
    double alpha, dhx, dhy, dhz;
};
struct PoisArgs {
    const b200mg_fab *phi, *rhs, *f;
    const b200mg_ifab* m;
    double dhx, dhy, dhz;
};

// descriptors of one box, loaded once per kernel (not per cell)
struct AbecViews {
    View<double> phi, rhs, a, bx, by, bz;
    const b200mg_fab* f6; const b200mg_ifab* m6;
    double alpha, dhx, dhy, dhz;
    __device__ __forceinline__ AbecViews (const AbecArgs& A, int box)
        : phi(view(A.phi[box])), rhs(view(A.rhs[box])), a(view(A.a[box])), bx(view(A.bx[box])), by(view(A.by[box])), bz(view(A.bz[box])),
          f6(A.f + 6 * box), m6(A.m + 6 * box), alpha(A.alpha), dhx(A.dhx), dhy(A.dhy), dhz(A.dhz) {}
};
struct PoisViews {
    View<double> phi, rhs;
    const b200mg_fab* f6; const b200mg_ifab* m6;
    double dhx, dhy, dhz;
    __device__ __forceinline__ PoisViews (const PoisArgs& A, int box)
        : phi(view(A.phi[box])), rhs(view(A.rhs[box])), f6(A.f + 6 * box), m6(A.m + 6 * box), dhx(A.dhx), dhy(A.dhy), dhz(A.dhz) {}
};

__device__ __forceinline__ void
gsrb_at (int i, int j, int k, const b200mg_box& vb, const AbecViews& V)
{
    double* pc = V.phi.ptr(i, j, k);
    const double p = *pc;
    const int js = int(V.phi.js), ks = int(V.phi.ks);
    const double* pbx = V.bx.ptr(i, j, k); const double* pby = V.by.ptr(i, j, k); const double* pbz = V.bz.ptr(i, j, k);
    double r;
    if (on_surface(i, j, k, vb)) {
        const FaceCoefs cf = face_coefs(i, j, k, vb, V.f6, V.m6);
        r = gsrb_abec_cell(p, pc[-1], pc[1], pc[-js], pc[js], pc[-ks], pc[ks],
                           V.rhs(i, j, k), V.a(i, j, k), pbx[0], pbx[1], pby[0], pby[V.by.js], pbz[0], pbz[V.bz.ks],
                           cf.c[0], cf.c[1], cf.c[2], cf.c[3], cf.c[4], cf.c[5], V.alpha, V.dhx, V.dhy, V.dhz);
    } else {
        r = gsrb_abec_cell_interior(p, pc[-1], pc[1], pc[-js], pc[js], pc[-ks], pc[ks],
                                    V.rhs(i, j, k), V.a(i, j, k), pbx[0], pbx[1], pby[0], pby[V.by.js], pbz[0], pbz[V.bz.ks],
                                    V.alpha, V.dhx, V.dhy, V.dhz);
    }
    *pc = r;
}

__device__ __forceinline__ void
gsrb_at (int i, int j, int k, const b200mg_box& vb, const PoisViews& V)
{
    double* pc = V.phi.ptr(i, j, k);
    const int js = int(V.phi.js), ks = int(V.phi.ks);
    FaceCoefs cf;
    if (on_surface(i, j, k, vb)) { cf = face_coefs(i, j, k, vb, V.f6, V.m6); }
    else {
#pragma unroll
        for (int n = 0; n < 6; ++n) { cf.c[n] = 0.0; }
    }
    *pc = gsrb_poisson_cell(*pc, pc[-1], pc[1], pc[-js], pc[js], pc[-ks], pc[ks], V.rhs(i, j, k),
                            cf.c[0], cf.c[1], cf.c[2], cf.c[3], cf.c[4], cf.c[5], V.dhx, V.dhy, V.dhz);
}

// one colour; each thread owns the cell pair (i0,i0+1) and updates the one with the right parity
__global__ void __launch_bounds__(kTileTX * B200MG_TILE_Y)
k_gsrb_abec (const b200mg_tile* __restrict__ tiles, const b200mg_box* __restrict__ vbox, AbecArgs A, int redblack)
{
    const b200mg_tile t = tiles[blockIdx.x];
    const b200mg_box vb = vbox[t.box];
    const int j = t.j0 + int(threadIdx.y);
    if (j > vb.hi[1]) { return; }
    const int khi = min(t.k0 + tile_nk(t) - 1, vb.hi[2]);
    const AbecViews V(A, t.box);
    for (int k = t.k0; k <= khi; ++k) {
        const int off = (vb.lo[0] + j + k + redblack) & 1;
        for (int i = vb.lo[0] + off + 2 * int(threadIdx.x); i <= vb.hi[0]; i += 2 * int(blockDim.x)) {
            gsrb_at(i, j, k, vb, V);
        }
    }
}

__global__ void __launch_bounds__(kTileTX * B200MG_TILE_Y)
k_gsrb_poisson (const b200mg_tile* __restrict__ tiles, const b200mg_box* __restrict__ vbox, PoisArgs A, int redblack)
{
    const b200mg_tile t = tiles[blockIdx.x];
    const b200mg_box vb = vbox[t.box];
    const int j = t.j0 + int(threadIdx.y);
    if (j > vb.hi[1]) { return; }
    const int khi = min(t.k0 + tile_nk(t) - 1, vb.hi[2]);
    const PoisViews V(A, t.box);
    for (int k = t.k0; k <= khi; ++k) {
        const int off = (vb.lo[0] + j + k + redblack) & 1;
        for (int i = vb.lo[0] + off + 2 * int(threadIdx.x); i <= vb.hi[0]; i += 2 * int(blockDim.x)) {
            gsrb_at(i, j, k, vb, V);
        }
    }
}

// Damped Jacobi sweep (abec_jacobi AMReX_MLABecLap_3D_K.H:332-375, mlpoisson_jacobi AMReX_MLPoisson_3D_K.H:250-281): the
// reference first stores Ax = L(phi) (Fapply, AMReX_MLABecLaplacian.H:866-871) and then updates every cell; here both
// happen in one out-of-place pass (phi_in -> phi_out), so Ax is never written.  ad*: the apply's beta*dxinv^2
// (AMReX_MLABecLap_3D_K.H:18-20), dh*: the smoother's beta/h^2 (AMReX_MLABecLaplacian.H:907-909).
struct JacobiDh { double adx, ady, adz; };

__global__ void __launch_bounds__(kTileTX * B200MG_TILE_Y)
k_jacobi_abec (const b200mg_tile* __restrict__ tiles, const b200mg_box* __restrict__ vbox, const b200mg_fab* outf, AbecArgs A, JacobiDh D)
{
    const b200mg_tile t = tiles[blockIdx.x];
    const b200mg_box vb = vbox[t.box];
    const AbecViews V(A, t.box);
    const auto out = view(outf[t.box]);
    const int js = int(V.phi.js), ks = int(V.phi.ks);
    tile_for(t, vb, 0, [&] (int i, int j, int k) {
        const double* pc = V.phi.ptr(i, j, k);
        const double p = *pc;
        const double* pbx = V.bx.ptr(i, j, k); const double* pby = V.by.ptr(i, j, k); const double* pbz = V.bz.ptr(i, j, k);
        const double a = V.a(i, j, k);
        const double bxm = pbx[0], bxp = pbx[1], bym = pby[0], byp = pby[V.by.js], bzm = pbz[0], bzp = pbz[V.bz.ks];
        const double ax = adotx_abec_cell(p, pc[-1], pc[1], pc[-js], pc[js], pc[-ks], pc[ks], a, bxm, bxp, bym, byp, bzm, bzp,
                                          V.alpha, D.adx, D.ady, D.adz);
        FaceCoefs cf;
        if (on_surface(i, j, k, vb)) { cf = face_coefs(i, j, k, vb, V.f6, V.m6); }
        else {
#pragma unroll
            for (int n = 0; n < 6; ++n) { cf.c[n] = 0.0; }
        }
        const double gamma = V.alpha * a + V.dhx * (bxm + bxp) + V.dhy * (bym + byp) + V.dhz * (bzm + bzp);
        const double g_m_d = gamma - (V.dhx * (bxm * cf.c[0] + bxp * cf.c[3]) + V.dhy * (bym * cf.c[1] + byp * cf.c[4]) + V.dhz * (bzm * cf.c[2] + bzp * cf.c[5]));
        out(i, j, k) = p + (2.0 / 3.0) * (V.rhs(i, j, k) - ax) / g_m_d;
    });
}

__global__ void __launch_bounds__(kTileTX * B200MG_TILE_Y)
k_jacobi_poisson (const b200mg_tile* __restrict__ tiles, const b200mg_box* __restrict__ vbox, const b200mg_fab* outf, PoisArgs A)
{
    const b200mg_tile t = tiles[blockIdx.x];
    const b200mg_box vb = vbox[t.box];
    const PoisViews V(A, t.box);
    const auto out = view(outf[t.box]);
    const int js = int(V.phi.js), ks = int(V.phi.ks);
    tile_for(t, vb, 0, [&] (int i, int j, int k) {
        const double* pc = V.phi.ptr(i, j, k);
        const double p = *pc;
        const double ax = adotx_poisson_cell(p, pc[-1], pc[1], pc[-js], pc[js], pc[-ks], pc[ks], V.dhx, V.dhy, V.dhz);
        FaceCoefs cf;
        if (on_surface(i, j, k, vb)) { cf = face_coefs(i, j, k, vb, V.f6, V.m6); }
        else {
#pragma unroll
            for (int n = 0; n < 6; ++n) { cf.c[n] = 0.0; }
        }
        const double gamma = -2.0 * (V.dhx + V.dhy + V.dhz);
        const double g_m_d = gamma + V.dhx * (cf.c[0] + cf.c[3]) + V.dhy * (cf.c[1] + cf.c[4]) + V.dhz * (cf.c[2] + cf.c[5]);
        out(i, j, k) = p + (2.0 / 3.0) * (V.rhs(i, j, k) - ax) / g_m_d;
    });
}

// Surface shell sweep: blockIdx.x = box*6 + face, blockIdx.y = chunk of the face; every shell cell belongs to exactly one
// face (x faces own their edges/corners, y faces exclude the x extremes, z faces exclude x and y extremes).  Threads run
// along the face's fastest-varying tangential direction (x for y/z faces: coalesced).
constexpr int kShellThreads = 256;
// blockIdx.y chunks so that every thread owns ONE cell of the swept colour (the launch is latency bound otherwise)
inline int shell_chunks (int face_cells) { const int c = (face_cells / 2 + kShellThreads - 1) / kShellThreads; return c < 1 ? 1 : (c > 64 ? 64 : c); }

template <class F>
__device__ __forceinline__ void shell_loop (const b200mg_box& vb, int face, int redblack, F&& f)
{
    const int d = face % 3;
    const int fix = (face < 3) ? vb.lo[d] : vb.hi[d];
    if (face >= 3 && vb.lo[d] == vb.hi[d]) { return; }   // one-cell-thick box: low face owns it
    int lo[3] = {vb.lo[0], vb.lo[1], vb.lo[2]}, hi[3] = {vb.hi[0], vb.hi[1], vb.hi[2]};
    if (d >= 1) { lo[0] += 1; hi[0] -= 1; }
    if (d == 2) { lo[1] += 1; hi[1] -= 1; }
    lo[d] = hi[d] = fix;
    const int du = (d == 0) ? 1 : 0, dv = (d == 2) ? 1 : 2;     // tangential directions, u fastest in memory
    const int nu = hi[du] - lo[du] + 1, nv = hi[dv] - lo[dv] + 1;
    if (nu <= 0 || nv <= 0) { return; }
    // only cells of the swept colour are enumerated: row v holds them at u = 2*uu + (parity of the row's first cell)
    const unsigned nuh = unsigned(nu + 1) / 2u;
    const unsigned n = nuh * unsigned(nv);
    for (unsigned t = threadIdx.x + blockIdx.y * blockDim.x; t < n; t += blockDim.x * gridDim.y) {
        const unsigned v = t / nuh, uu = t - v * nuh;
        int idx[3];
        idx[d] = fix; idx[du] = lo[du]; idx[dv] = lo[dv] + int(v);
        const int u = 2 * int(uu) + ((idx[0] + idx[1] + idx[2] + redblack) & 1);
        idx[du] += u;
        if (u < nu) { f(idx[0], idx[1], idx[2]); }
    }
}

__global__ void __launch_bounds__(256)
k_gsrb_shell_abec (const b200mg_box* __restrict__ vbox, AbecArgs A, int redblack)
{
    const int box = blockIdx.x / 6, face = blockIdx.x % 6;
    const b200mg_box vb = vbox[box];
    const AbecViews V(A, box);
    shell_loop(vb, face, redblack, [&] (int i, int j, int k) { gsrb_at(i, j, k, vb, V); });
}

__global__ void __launch_bounds__(256)
k_gsrb_shell_poisson (const b200mg_box* __restrict__ vbox, PoisArgs A, int redblack)
{
    const int box = blockIdx.x / 6, face = blockIdx.x % 6;
    const b200mg_box vb = vbox[box];
    const PoisViews V(A, box);
    shell_loop(vb, face, redblack, [&] (int i, int j, int k) { gsrb_at(i, j, k, vb, V); });
}

// Shell sweep with face links (b200mg_facelink): the six neighbours of a shell cell, the one(s) beyond a linked face taken
// from the neighbouring fab's valid cell - the value a halo exchange would have copied into the ghost cell, so the bits do
// not change - and, with push, the new value stored into the ghost cell(s) that fab keeps for this cell.  A cell on an edge
// or a corner of its box lies on two or three faces.  No races within one colour: a cell of the swept colour reads cells of
// the other colour only, and the pushed ghost cells (swept colour) are read by nobody during the sweep.
struct Nbr6 { double xm, xp, ym, yp, zm, zp; };

// the link of the face a block sweeps, resolved once per block: every cell of the block lies on that face
struct OwnLink { int face; bool linked; View<double> nbr; int s0, s1, s2; };

__device__ __forceinline__ OwnLink
own_link (int face, const b200mg_fab* phif, const b200mg_facelink* __restrict__ L6)
{
    OwnLink o;
    const b200mg_facelink l = L6[face];
    o.face = face; o.linked = (l.fab >= 0);
    o.nbr = view(phif[o.linked ? l.fab : 0]);
    o.s0 = l.shift[0]; o.s1 = l.shift[1]; o.s2 = l.shift[2];
    return o;
}

__device__ __forceinline__ Nbr6
linked_neighbours (int i, int j, int k, const b200mg_box& vb, const double* pc, int js, int ks,
                   const b200mg_fab* phif, const b200mg_facelink* __restrict__ L6, const OwnLink& O)
{
    Nbr6 n;
    auto beyond = [&] (int face, int ii, int jj, int kk, const double* own) {
        if (face == O.face) { return O.linked ? O.nbr(ii + O.s0, jj + O.s1, kk + O.s2) : *own; }
        const b200mg_facelink l = L6[face];              // a cell on an edge or a corner of the box: its other face(s)
        return (l.fab >= 0) ? view(phif[l.fab])(ii + l.shift[0], jj + l.shift[1], kk + l.shift[2]) : *own;
    };
    n.xm = (i == vb.lo[0]) ? beyond(0, i - 1, j, k, pc - 1) : pc[-1];
    n.xp = (i == vb.hi[0]) ? beyond(3, i + 1, j, k, pc + 1) : pc[1];
    n.ym = (j == vb.lo[1]) ? beyond(1, i, j - 1, k, pc - js) : pc[-js];
    n.yp = (j == vb.hi[1]) ? beyond(4, i, j + 1, k, pc + js) : pc[js];
    n.zm = (k == vb.lo[2]) ? beyond(2, i, j, k - 1, pc - ks) : pc[-ks];
    n.zp = (k == vb.hi[2]) ? beyond(5, i, j, k + 1, pc + ks) : pc[ks];
    return n;
}

__device__ __forceinline__ void
push_to_links (int i, int j, int k, const b200mg_box& vb, double v, const b200mg_fab* phif, const b200mg_facelink* __restrict__ L6,
               const OwnLink& O)
{
    // this cell is the ghost cell (i,j,k) + shift of the fab behind every linked face it lies on
    auto put = [&] (int face) {
        if (face == O.face) { if (O.linked) { O.nbr(i + O.s0, j + O.s1, k + O.s2) = v; } return; }
        const b200mg_facelink l = L6[face];
        if (l.fab >= 0) { view(phif[l.fab])(i + l.shift[0], j + l.shift[1], k + l.shift[2]) = v; }
    };
    if (i == vb.lo[0]) { put(0); }
    if (i == vb.hi[0]) { put(3); }
    if (j == vb.lo[1]) { put(1); }
    if (j == vb.hi[1]) { put(4); }
    if (k == vb.lo[2]) { put(2); }
    if (k == vb.hi[2]) { put(5); }
}

__global__ void __launch_bounds__(256)
k_gsrb_shell_abec_linked (const b200mg_box* __restrict__ vbox, AbecArgs A, int redblack, const b200mg_facelink* __restrict__ links, int push)
{
    const int box = blockIdx.x / 6, face = blockIdx.x % 6;
    const b200mg_box vb = vbox[box];
    const AbecViews V(A, box);
    const b200mg_facelink* L6 = links + 6 * box;
    const OwnLink O = own_link(face, A.phi, L6);
    shell_loop(vb, face, redblack, [&] (int i, int j, int k) {
        double* pc = V.phi.ptr(i, j, k);
        const double p = *pc;
        const Nbr6 n = linked_neighbours(i, j, k, vb, pc, int(V.phi.js), int(V.phi.ks), A.phi, L6, O);
        const double* pbx = V.bx.ptr(i, j, k); const double* pby = V.by.ptr(i, j, k); const double* pbz = V.bz.ptr(i, j, k);
        const FaceCoefs cf = face_coefs(i, j, k, vb, V.f6, V.m6);      // (every shell cell is a surface cell)
        const double r = gsrb_abec_cell(p, n.xm, n.xp, n.ym, n.yp, n.zm, n.zp,
                                        V.rhs(i, j, k), V.a(i, j, k), pbx[0], pbx[1], pby[0], pby[V.by.js], pbz[0], pbz[V.bz.ks],
                                        cf.c[0], cf.c[1], cf.c[2], cf.c[3], cf.c[4], cf.c[5], V.alpha, V.dhx, V.dhy, V.dhz);
        *pc = r;
        if (push) { push_to_links(i, j, k, vb, r, A.phi, L6, O); }
    });
}

__global__ void __launch_bounds__(256)
k_gsrb_shell_poisson_linked (const b200mg_box* __restrict__ vbox, PoisArgs A, int redblack, const b200mg_facelink* __restrict__ links, int push)
{
    const int box = blockIdx.x / 6, face = blockIdx.x % 6;
    const b200mg_box vb = vbox[box];
    const PoisViews V(A, box);
    const b200mg_facelink* L6 = links + 6 * box;
    const OwnLink O = own_link(face, A.phi, L6);
    shell_loop(vb, face, redblack, [&] (int i, int j, int k) {
        double* pc = V.phi.ptr(i, j, k);
        const Nbr6 n = linked_neighbours(i, j, k, vb, pc, int(V.phi.js), int(V.phi.ks), A.phi, L6, O);
        const FaceCoefs cf = face_coefs(i, j, k, vb, V.f6, V.m6);
        const double r = gsrb_poisson_cell(*pc, n.xm, n.xp, n.ym, n.yp, n.zm, n.zp, V.rhs(i, j, k),
                                           cf.c[0], cf.c[1], cf.c[2], cf.c[3], cf.c[4], cf.c[5], V.dhx, V.dhy, V.dhz);
        *pc = r;
        if (push) { push_to_links(i, j, k, vb, r, A.phi, L6, O); }
    });
}

// ---------------------------------------------------------------------------------------- adotx
__global__ void __launch_bounds__(kTileTX * B200MG_TILE_Y)
k_adotx_abec (const b200mg_tile* __restrict__ tiles, const b200mg_box* __restrict__ vbox,
              const b200mg_fab* yf, const b200mg_fab* xf, const b200mg_fab* rf, const b200mg_fab* af,
              const b200mg_fab* bxf, const b200mg_fab* byf, const b200mg_fab* bzf,
              double alpha, double dhx, double dhy, double dhz)
{
    const b200mg_tile t = tiles[blockIdx.x];
    const b200mg_box vb = vbox[t.box];
    const auto y = view(yf[t.box]); const auto x = view(xf[t.box]); const auto a = view(af[t.box]);
    const auto bx = view(bxf[t.box]); const auto by = view(byf[t.box]); const auto bz = view(bzf[t.box]);
    const bool has_r = (rf != nullptr);
    View<double> r = has_r ? view(rf[t.box]) : y;
    tile_for(t, vb, 0, [&] (int i, int j, int k) {
        const double* xc = x.ptr(i, j, k);
        const double v = adotx_abec_cell(*xc, xc[-1], xc[1], xc[-x.js], xc[x.js], xc[-x.ks], xc[x.ks],
                                         a(i, j, k), bx(i, j, k), bx(i + 1, j, k), by(i, j, k), by(i, j + 1, k),
                                         bz(i, j, k), bz(i, j, k + 1), alpha, dhx, dhy, dhz);
        y(i, j, k) = has_r ? (r(i, j, k) + (-1.0) * v) : v;   // Xpay(y,-1,b): y = b + (-1)*y
    });
}

__global__ void __launch_bounds__(kTileTX * B200MG_TILE_Y)
k_adotx_poisson (const b200mg_tile* __restrict__ tiles, const b200mg_box* __restrict__ vbox,
                 const b200mg_fab* yf, const b200mg_fab* xf, const b200mg_fab* rf,
                 double dhx, double dhy, double dhz)
{
    const b200mg_tile t = tiles[blockIdx.x];
    const b200mg_box vb = vbox[t.box];
    const auto y = view(yf[t.box]); const auto x = view(xf[t.box]);
    const bool has_r = (rf != nullptr);
    View<double> r = has_r ? view(rf[t.box]) : y;
    tile_for(t, vb, 0, [&] (int i, int j, int k) {
        const double* xc = x.ptr(i, j, k);
        const double v = adotx_poisson_cell(*xc, xc[-1], xc[1], xc[-x.js], xc[x.js], xc[-x.ks], xc[x.ks], dhx, dhy, dhz);
        y(i, j, k) = has_r ? (r(i, j, k) + (-1.0) * v) : v;
    });
}

// z-marching operator apply / residual: each thread owns the cell pair (i0, i0+1) of one row and streams through the
// planes of its tile; the x pair of planes k-1, k, k+1 and the z-face coefficient stay in registers, everything is read with
// 16-byte loads (x/y neighbour loads hit L1).  Requires even x extents (16-byte aligned pairs); nx <= 2 * blockDim.x.
// NORM: max |y| over the launch is folded into *norm (which the host zeroes ahead of the launch) - the residual and its
// inf-norm (MLMGT::ResNormInf, AMReX_MLMG.H:1808-1812) in ONE pass: per-thread running maximum, warp shuffles, one atomic
// per warp.  max is order independent, so the result is the same bits as the separate reduction kernel.
template <bool ABEC, bool NORM>
__global__ void __launch_bounds__(kTileTX * B200MG_TILE_Y, 4)
k_adotx_pair (const b200mg_tile* __restrict__ tiles, const b200mg_box* __restrict__ vbox,
              const b200mg_fab* yf, const b200mg_fab* xf, const b200mg_fab* rf, const b200mg_fab* af,
              const b200mg_fab* bxf, const b200mg_fab* byf, const b200mg_fab* bzf,
              double alpha, double dhx, double dhy, double dhz, double* __restrict__ norm)
{
    const b200mg_tile t = tiles[blockIdx.x];
    const b200mg_box vb = vbox[t.box];
    const int i0 = vb.lo[0] + 2 * int(threadIdx.x);
    const int j = t.j0 + int(threadIdx.y);
    const bool idle = (i0 >= vb.hi[0] || j > vb.hi[1]);
    if (!NORM && idle) { return; }
    double nrm = 0.0;
    if (!idle) {
    const int k0 = t.k0, k1 = min(t.k0 + tile_nk(t) - 1, vb.hi[2]);
    const auto x = view(xf[t.box]); const auto y = view(yf[t.box]);
    const int x_js = int(x.js), x_ks = int(x.ks), y_ks = int(y.ks);
    const double* px = x.ptr(i0, j, k0);
    double* py = y.ptr(i0, j, k0);
    const bool has_r = (rf != nullptr);
    const double* pr = nullptr; int r_ks = 0;
    if (has_r) { const auto r = view(rf[t.box]); pr = r.ptr(i0, j, k0); r_ks = int(r.ks); }
    const double *pa = nullptr, *pbx = nullptr, *pby = nullptr, *pbz = nullptr;
    int a_ks = 0, bx_ks = 0, by_ks = 0, by_js = 0, bz_ks = 0;
    double2 bzlo = make_double2(0.0, 0.0);
    if constexpr (ABEC) {
        const auto a = view(af[t.box]); const auto bx = view(bxf[t.box]); const auto by = view(byf[t.box]); const auto bz = view(bzf[t.box]);
        pa = a.ptr(i0, j, k0); pbx = bx.ptr(i0, j, k0); pby = by.ptr(i0, j, k0); pbz = bz.ptr(i0, j, k0);
        a_ks = int(a.ks); bx_ks = int(bx.ks); by_ks = int(by.ks); by_js = int(by.js); bz_ks = int(bz.ks);
        bzlo = __ldg(reinterpret_cast<const double2*>(pbz));
    }
    double2 xm = *reinterpret_cast<const double2*>(px - x_ks);
    double2 xc = *reinterpret_cast<const double2*>(px);
    for (int k = k0; k <= k1; ++k) {
        const double2 xp = *reinterpret_cast<const double2*>(px + x_ks);
        const double xl = px[-1], xr = px[2];
        const double2 ylo = *reinterpret_cast<const double2*>(px - x_js);
        const double2 yhi = *reinterpret_cast<const double2*>(px + x_js);
        double2 v;
        if constexpr (ABEC) {
            const double2 a = __ldg(reinterpret_cast<const double2*>(pa));
            const double2 bx = __ldg(reinterpret_cast<const double2*>(pbx));
            const double bx2 = __ldg(pbx + 2);
            const double2 bylo = __ldg(reinterpret_cast<const double2*>(pby));
            const double2 byhi = __ldg(reinterpret_cast<const double2*>(pby + by_js));
            const double2 bzhi = __ldg(reinterpret_cast<const double2*>(pbz + bz_ks));
            v.x = adotx_abec_cell(xc.x, xl, xc.y, ylo.x, yhi.x, xm.x, xp.x, a.x, bx.x, bx.y, bylo.x, byhi.x, bzlo.x, bzhi.x, alpha, dhx, dhy, dhz);
            v.y = adotx_abec_cell(xc.y, xc.x, xr, ylo.y, yhi.y, xm.y, xp.y, a.y, bx.y, bx2, bylo.y, byhi.y, bzlo.y, bzhi.y, alpha, dhx, dhy, dhz);
            bzlo = bzhi;
            pa += a_ks; pbx += bx_ks; pby += by_ks; pbz += bz_ks;
        } else {
            v.x = adotx_poisson_cell(xc.x, xl, xc.y, ylo.x, yhi.x, xm.x, xp.x, dhx, dhy, dhz);
            v.y = adotx_poisson_cell(xc.y, xc.x, xr, ylo.y, yhi.y, xm.y, xp.y, dhx, dhy, dhz);
        }
        if (has_r) {   // Xpay(y,-1,b): y = b + (-1)*y
            const double2 r = __ldg(reinterpret_cast<const double2*>(pr));
            v.x = r.x + (-1.0) * v.x; v.y = r.y + (-1.0) * v.y;
            pr += r_ks;
        }
        *reinterpret_cast<double2*>(py) = v;
        if constexpr (NORM) { nrm = fmax(nrm, fmax(fabs(v.x), fabs(v.y))); }
        xm = xc; xc = xp;
        px += x_ks; py += y_ks;
    }
    }
    if constexpr (NORM) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { nrm = fmax(nrm, __shfl_xor_sync(0xffffffffu, nrm, o)); }
        // non-negative doubles order like their bit patterns
        if (((threadIdx.x + threadIdx.y * blockDim.x) & 31u) == 0u && nrm > 0.0) {
            atomicMax(reinterpret_cast<unsigned long long*>(norm), static_cast<unsigned long long>(__double_as_longlong(nrm)));
        }
    }
}

// Residual fused with its restriction (MLMGT::mgVcycle, AMReX_MLMG.H:1332-1345: computeResOfCorrection followed by
// restriction of rescor; kernels mlabeclap_adotx + Xpay + amrex_avgdown, AMReX_MultiFabUtil_3D_C.H:381-394): the fine residual
// r - L(x) is never stored - the cycle does not read it again - only its 2x2x2 averages, so the pass moves 49 B/cell
// instead of 56 + 9.  Thread map of k_adotx_pair; the thread of an even row collects the pair of the odd row above it
// through shared memory (two plane buffers, one barrier per plane pair) and adds the eight values in amrex_avgdown's order
// (x fastest, then y, then z), so the coarse field has the bits of the two-kernel sequence.  Requires even box corners and
// extents (the boxes of a level that coarsens by 2) and an even tile depth.
template <bool ABEC>
__global__ void __launch_bounds__(kTileTX * B200MG_TILE_Y, 4)
k_adotx_pair_restrict (const b200mg_tile* __restrict__ tiles, const b200mg_box* __restrict__ vbox,
                       const b200mg_fab* cf, const b200mg_fab* xf, const b200mg_fab* rf, const b200mg_fab* af,
                       const b200mg_fab* bxf, const b200mg_fab* byf, const b200mg_fab* bzf,
                       double alpha, double dhx, double dhy, double dhz)
{
    __shared__ double2 sh[2][2][B200MG_TILE_Y / 2][kTileTX];      // [pair parity][plane of the pair][odd row][x pair]
    const b200mg_tile t = tiles[blockIdx.x];
    const b200mg_box vb = vbox[t.box];
    const int tx = int(threadIdx.x), ty = int(threadIdx.y);
    const int i0 = vb.lo[0] + 2 * tx;
    const int j = t.j0 + ty;
    const bool idle = (i0 >= vb.hi[0] || j > vb.hi[1]);
    const int k0 = t.k0, k1 = min(t.k0 + tile_nk(t) - 1, vb.hi[2]);
    const auto x = view(xf[t.box]); const auto crse = view(cf[t.box]);
    const int x_js = int(x.js), x_ks = int(x.ks);
    const double* px = idle ? nullptr : x.ptr(i0, j, k0);
    const double* pr = nullptr; int r_ks = 0;
    { const auto r = view(rf[t.box]); if (!idle) { pr = r.ptr(i0, j, k0); } r_ks = int(r.ks); }
    const double *pa = nullptr, *pbx = nullptr, *pby = nullptr, *pbz = nullptr;
    int a_ks = 0, bx_ks = 0, by_ks = 0, by_js = 0, bz_ks = 0;
    double2 bzlo = make_double2(0.0, 0.0);
    if constexpr (ABEC) {
        const auto a = view(af[t.box]); const auto bx = view(bxf[t.box]); const auto by = view(byf[t.box]); const auto bz = view(bzf[t.box]);
        a_ks = int(a.ks); bx_ks = int(bx.ks); by_ks = int(by.ks); by_js = int(by.js); bz_ks = int(bz.ks);
        if (!idle) {
            pa = a.ptr(i0, j, k0); pbx = bx.ptr(i0, j, k0); pby = by.ptr(i0, j, k0); pbz = bz.ptr(i0, j, k0);
            bzlo = __ldg(reinterpret_cast<const double2*>(pbz));
        }
    }
    double2 xm = make_double2(0.0, 0.0), xc = xm;
    if (!idle) { xm = *reinterpret_cast<const double2*>(px - x_ks); xc = *reinterpret_cast<const double2*>(px); }
    double2 v0 = make_double2(0.0, 0.0);                          // this thread's residual pair of the first plane of the pair
    for (int k = k0; k <= k1; ++k) {
        double2 v = make_double2(0.0, 0.0);
        if (!idle) {
            const double2 xp = *reinterpret_cast<const double2*>(px + x_ks);
            const double xl = px[-1], xr = px[2];
            const double2 ylo = *reinterpret_cast<const double2*>(px - x_js);
            const double2 yhi = *reinterpret_cast<const double2*>(px + x_js);
            if constexpr (ABEC) {
                const double2 a = __ldg(reinterpret_cast<const double2*>(pa));
                const double2 bx = __ldg(reinterpret_cast<const double2*>(pbx));
                const double bx2 = __ldg(pbx + 2);
                const double2 bylo = __ldg(reinterpret_cast<const double2*>(pby));
                const double2 byhi = __ldg(reinterpret_cast<const double2*>(pby + by_js));
                const double2 bzhi = __ldg(reinterpret_cast<const double2*>(pbz + bz_ks));
                v.x = adotx_abec_cell(xc.x, xl, xc.y, ylo.x, yhi.x, xm.x, xp.x, a.x, bx.x, bx.y, bylo.x, byhi.x, bzlo.x, bzhi.x, alpha, dhx, dhy, dhz);
                v.y = adotx_abec_cell(xc.y, xc.x, xr, ylo.y, yhi.y, xm.y, xp.y, a.y, bx.y, bx2, bylo.y, byhi.y, bzlo.y, bzhi.y, alpha, dhx, dhy, dhz);
                bzlo = bzhi;
                pa += a_ks; pbx += bx_ks; pby += by_ks; pbz += bz_ks;
            } else {
                v.x = adotx_poisson_cell(xc.x, xl, xc.y, ylo.x, yhi.x, xm.x, xp.x, dhx, dhy, dhz);
                v.y = adotx_poisson_cell(xc.y, xc.x, xr, ylo.y, yhi.y, xm.y, xp.y, dhx, dhy, dhz);
            }
            const double2 r = __ldg(reinterpret_cast<const double2*>(pr));
            v.x = r.x + (-1.0) * v.x; v.y = r.y + (-1.0) * v.y;       // Xpay(y,-1,b)
            pr += r_ks;
            xm = xc; xc = xp;
            px += x_ks;
        }
        const int kk = k - k0;                                     // even: first plane of a pair
        const int par = (kk >> 1) & 1;
        if (ty & 1) { sh[par][kk & 1][ty >> 1][tx] = v; }
        if ((kk & 1) == 0) { v0 = v; continue; }
        __syncthreads();
        if (!(ty & 1) && !idle) {
            const double2 b0 = sh[par][0][ty >> 1][tx], b1 = sh[par][1][ty >> 1][tx];
            double c = 0.0;
            c += v0.x; c += v0.y; c += b0.x; c += b0.y; c += v.x; c += v.y; c += b1.x; c += b1.y;
            crse(i0 >> 1, j >> 1, (k - 1) >> 1) = 0.125 * c;
        }
    }
}

__global__ void __launch_bounds__(kTileTX * B200MG_TILE_Y)
k_normalize_abec (const b200mg_tile* __restrict__ tiles, const b200mg_box* __restrict__ vbox,
                  const b200mg_fab* xf, const b200mg_fab* af,
                  const b200mg_fab* bxf, const b200mg_fab* byf, const b200mg_fab* bzf,
                  double alpha, double dhx, double dhy, double dhz)
{
    const b200mg_tile t = tiles[blockIdx.x];
    const b200mg_box vb = vbox[t.box];
    const auto x = view(xf[t.box]); const auto a = view(af[t.box]);
    const auto bx = view(bxf[t.box]); const auto by = view(byf[t.box]); const auto bz = view(bzf[t.box]);
    tile_for(t, vb, 0, [&] (int i, int j, int k) {   // AMReX_MLABecLap_3D_K.H:71-74
        x(i, j, k) /= alpha * a(i, j, k) + dhx * (bx(i, j, k) + bx(i + 1, j, k))
            + dhy * (by(i, j, k) + by(i, j + 1, k)) + dhz * (bz(i, j, k) + bz(i, j, k + 1));
    });
}

// --------------------------------------------------------------------------------- boundary conditions
constexpr int kBcDirichlet = 101, kBcNeumann = 102, kBcReflectOdd = 103;

// one block per (box, face) item; threads sweep the face's ghost cells (tangential extent = valid box)
template <class F>
__device__ __forceinline__ void face_loop (const b200mg_box& vb, int face, F&& f)
{
    const int d = face % 3;
    const int g = (face < 3) ? vb.lo[d] - 1 : vb.hi[d] + 1;   // ghost index in the normal direction
    const int d1 = (d == 0) ? 1 : 0, d2 = (d == 2) ? 1 : 2;    // tangential dirs, d1 fastest
    const int n1 = vb.hi[d1] - vb.lo[d1] + 1, n2 = vb.hi[d2] - vb.lo[d2] + 1;
    for (int t = threadIdx.x + blockIdx.y * blockDim.x; t < n1 * n2; t += blockDim.x * gridDim.y) {
        int idx[3];
        idx[d] = g; idx[d1] = vb.lo[d1] + t % n1; idx[d2] = vb.lo[d2] + t / n1;
        f(idx[0], idx[1], idx[2]);
    }
}

__global__ void __launch_bounds__(128)
k_apply_bc (const b200mg_bcface* __restrict__ faces, const b200mg_box* __restrict__ vbox,
            const b200mg_fab* phif, const b200mg_ifab* mf, const b200mg_fab* bvf,
            int maxorder, double dxi0, double dxi1, double dxi2, int inhomog)
{
    const b200mg_bcface fc = faces[blockIdx.x];
    const b200mg_box vb = vbox[fc.box];
    const auto phi = view(phif[fc.box]);
    const auto mask = view(mf[fc.box * 6 + fc.face]);
    const int d = fc.face % 3;
    const int s = (fc.face < 3) ? 1 : -1;                 // towards the interior
    const long long st = (d == 0) ? 1 : ((d == 1) ? phi.js : phi.ks);
    const double dxinv = (d == 0) ? dxi0 : ((d == 1) ? dxi1 : dxi2);
    if (fc.bctype == kBcNeumann) {
        face_loop(vb, fc.face, [&] (int i, int j, int k) {
            if (mask(i, j, k) > 0) { double* p = phi.ptr(i, j, k); *p = p[s * st]; } });
    } else if (fc.bctype == kBcReflectOdd) {
        face_loop(vb, fc.face, [&] (int i, int j, int k) {
            if (mask(i, j, k) > 0) { double* p = phi.ptr(i, j, k); *p = -p[s * st]; } });
    } else if (fc.bctype == kBcDirichlet) {
        const int NX = min(fc.blen + 1, maxorder);
        double x[4] = {-fc.bcloc * dxinv, 0.5, 1.5, 2.5};
        double coef[4] = {0., 0., 0., 0.};
        poly_interp_coeff(-0.5, x, NX, coef);
        const bool inh = inhomog && (bvf != nullptr);
        face_loop(vb, fc.face, [&] (int i, int j, int k) {
            if (mask(i, j, k) > 0) {
                double* p = phi.ptr(i, j, k);
                double tmp = 0.0;
                for (int m = 1; m < NX; ++m) { tmp += p[m * s * st] * coef[m]; }
                if (inh) { tmp += view(bvf[fc.box * 6 + fc.face])(i, j, k) * coef[0]; }
                *p = tmp;
            } });
    }
}

// Inhomogeneous Neumann data (ghost cell = d(phi)/dn on the domain face).  mode 0: the boundary flux moves into the right-hand
// side of the cell inside the face (mllinop_apply_innu_*, AMReX_MLLinOp_K.H:930-1075: rhs -= fac*b*bcval on low faces,
// += on high faces, fac = beta*dxinv); mode 1: the face value of a flux / gradient array is overwritten by fac*b*bcval
// (MLCellABecLapT::addInhomogNeumannFlux, AMReX_MLCellABecLap.H:517-620).  Only faces flagged in on_face[] whose ghost
// cells are outside the domain (mask == 2) take part.  out3 / b3: per-direction fab tables (mode 0 passes rhs three times).
struct Innu { const b200mg_fab* out[3]; const b200mg_fab* b[3]; double fac[3]; int on_face[6]; int mode; };

__global__ void __launch_bounds__(128)
k_apply_innu (const b200mg_bcface* __restrict__ faces, const b200mg_box* __restrict__ vbox,
              const b200mg_ifab* mf, const b200mg_fab* bvf, Innu P)
{
    const b200mg_bcface fc = faces[blockIdx.x];
    if (!P.on_face[fc.face] || fc.bctype != kBcNeumann) { return; }
    const b200mg_box vb = vbox[fc.box];
    const int d = fc.face % 3;
    const bool low = fc.face < 3;
    const auto mask = view(mf[fc.box * 6 + fc.face]);
    const auto bv = view(bvf[fc.box * 6 + fc.face]);
    const auto out = view(P.out[d][fc.box]);
    const bool has_b = (P.b[d] != nullptr);
    View<double> bc = out;
    if (has_b) { bc = view(P.b[d][fc.box]); }
    const double fac = P.fac[d];
    face_loop(vb, fc.face, [&] (int i, int j, int k) {
        if (mask(i, j, k) != 2) { return; }
        // face index of the domain face = first valid cell (low side) or the ghost cell itself (high side)
        const int fi = i + ((low && d == 0) ? 1 : 0), fj = j + ((low && d == 1) ? 1 : 0), fk = k + ((low && d == 2) ? 1 : 0);
        const double b = has_b ? bc(fi, fj, fk) : 1.0;
        if (P.mode == 0) {
            if (low) { out(fi, fj, fk) -= fac * b * bv(i, j, k); }
            else { out(i - (d == 0), j - (d == 1), k - (d == 2)) += fac * b * bv(i, j, k); }
        } else {
            out(fi, fj, fk) = fac * b * bv(i, j, k);
        }
    });
}

// Robin boundary condition a*phi + b*dphi/dn = f on flagged domain faces; the data sits in the ghost cells of the face slabs
// ra / rb / rf.  With u_ghost = A + B*u_in, A = f/(b/h + a/2), B = (b/h - a/2)/(b/h + a/2), the face acts as a homogeneous
// Neumann face with a modified diagonal and right-hand side (AMReX_MLABecLaplacian.H:459-600, AMReX_MLCellABecLap.H:448-510):
//   mode 0: acoef(cell inside) += fac[d]*bcoef(face)*(1-B)      fac = (b_scalar/a_scalar)*dxinv^2   (applyRobinBCTermsCoeffs)
//   mode 1: rhs(cell inside)   += fac[d]*bcoef(face)*A          fac = b_scalar*dxinv^2             (applyInhomogNeumannTerm)
//   mode 2: face value of out3[d] := fac[d]*bcoef*dxinv*((1-B)*phi_in - A) on low faces, fac*b*dxinv*(A + (B-1)*phi_in) on high
//           faces                                               (addInhomogNeumannFlux, AMReX_MLCellABecLap.H:579-612)
// Modes 0 / 1 are read-modify-writes of cells that may belong to several faces: one face orientation per launch.
struct Robin { const b200mg_fab* out[3]; const b200mg_fab* b[3]; const b200mg_fab* phi; double fac[3]; double dxi[3]; int on_face[6]; int mode; };

__global__ void __launch_bounds__(128)
k_robin (const b200mg_bcface* __restrict__ faces, const b200mg_box* __restrict__ vbox, const b200mg_ifab* mf,
         const b200mg_fab* raf, const b200mg_fab* rbf, const b200mg_fab* rff, Robin P)
{
    const b200mg_bcface fc = faces[blockIdx.x];
    if (!P.on_face[fc.face]) { return; }
    const b200mg_box vb = vbox[fc.box];
    const int d = fc.face % 3;
    const bool low = fc.face < 3;
    const auto mask = view(mf[fc.box * 6 + fc.face]);
    const auto ra = view(raf[fc.box * 6 + fc.face]); const auto rb = view(rbf[fc.box * 6 + fc.face]); const auto rf = view(rff[fc.box * 6 + fc.face]);
    const auto out = view(P.out[d][fc.box]);
    const bool has_b = (P.b[d] != nullptr);
    View<double> bc = out, phi = out;
    if (has_b) { bc = view(P.b[d][fc.box]); }
    if (P.mode == 2) { phi = view(P.phi[fc.box]); }
    const double fac = P.fac[d], dxi = P.dxi[d];
    face_loop(vb, fc.face, [&] (int i, int j, int k) {
        if (mask(i, j, k) != 2) { return; }
        const int ci = i + (d == 0 ? (low ? 1 : -1) : 0), cj = j + (d == 1 ? (low ? 1 : -1) : 0), ck = k + (d == 2 ? (low ? 1 : -1) : 0);   // cell inside
        const int fi = low ? ci : i, fj = low ? cj : j, fk = low ? ck : k;                                                          // the domain face
        const double b = has_b ? bc(fi, fj, fk) : 1.0;
        const double a_ = ra(i, j, k), b_ = rb(i, j, k);
        if (P.mode == 0) {
            const double B = (b_ * dxi - a_ * 0.5) / (b_ * dxi + a_ * 0.5);
            out(ci, cj, ck) += fac * b * (1.0 - B);
        } else if (P.mode == 1) {
            const double A = rf(i, j, k) / (b_ * dxi + a_ * 0.5);
            out(ci, cj, ck) += fac * b * A;
        } else {
            const double tmp = 1.0 / (b_ * dxi + a_ * 0.5);
            const double RA = rf(i, j, k) * tmp;
            const double RB = (b_ * dxi - a_ * 0.5) * tmp;
            if (low) { out(fi, fj, fk) = fac * b * dxi * ((1.0 - RB) * phi(ci, cj, ck) - RA); }
            else { out(fi, fj, fk) = fac * b * dxi * (RA + (RB - 1.0) * phi(ci, cj, ck)); }
        }
    });
}

__global__ void __launch_bounds__(128)
k_comp_interp_coef0 (const b200mg_bcface* __restrict__ faces, const b200mg_box* __restrict__ vbox,
                     const b200mg_fab* ff, const b200mg_ifab* mf,
                     int maxorder, double dxi0, double dxi1, double dxi2)
{
    const b200mg_bcface fc = faces[blockIdx.x];
    const b200mg_box vb = vbox[fc.box];
    const auto f = view(ff[fc.box * 6 + fc.face]);
    const auto mask = view(mf[fc.box * 6 + fc.face]);
    const int d = fc.face % 3;
    const int s = (fc.face < 3) ? 1 : -1;
    const double dxinv = (d == 0) ? dxi0 : ((d == 1) ? dxi1 : dxi2);
    double c1 = 0.0;
    if (fc.bctype == kBcDirichlet) {
        const int NX = min(fc.blen + 1, maxorder);
        double x[4] = {-fc.bcloc * dxinv, 0.5, 1.5, 2.5};
        double coef[4] = {0., 0., 0., 0.};
        poly_interp_coeff(-0.5, x, NX, coef);
        c1 = coef[1];
    }
    const int bct = fc.bctype;
    face_loop(vb, fc.face, [&] (int i, int j, int k) {
        int ii = i, jj = j, kk = k;
        if (d == 0) { ii += s; } else if (d == 1) { jj += s; } else { kk += s; }
        if (bct == kBcNeumann) { f(ii, jj, kk) = 1.0; }
        else if (bct == kBcReflectOdd) { f(ii, jj, kk) = (mask(i, j, k) > 0) ? 1.0 : 0.0; }
        else if (bct == kBcDirichlet) { f(ii, jj, kk) = (mask(i, j, k) > 0) ? c1 : 0.0; }
    });
}

} // namespace

extern "C" {

int b200mg_gsrb_abec (int ntiles, const b200mg_tile* tiles, const b200mg_box* vbox,
                      const b200mg_fab* phi, const b200mg_fab* rhs, const b200mg_fab* a,
                      const b200mg_fab* bx, const b200mg_fab* by, const b200mg_fab* bz,
                      const b200mg_fab* f, const b200mg_ifab* m,
                      double alpha, double dhx, double dhy, double dhz, int redblack, cudaStream_t s)
{
    if (ntiles <= 0) { return 0; }
    AbecArgs A{phi, rhs, a, bx, by, bz, f, m, alpha, dhx, dhy, dhz};
    k_gsrb_abec<<<ntiles, tile_block(), 0, s>>>(tiles, vbox, A, redblack);
    return last_error();
}

int b200mg_gsrb_poisson (int ntiles, const b200mg_tile* tiles, const b200mg_box* vbox,
                         const b200mg_fab* phi, const b200mg_fab* rhs,
                         const b200mg_fab* f, const b200mg_ifab* m,
                         double dhx, double dhy, double dhz, int redblack, cudaStream_t s)
{
    if (ntiles <= 0) { return 0; }
    PoisArgs A{phi, rhs, f, m, dhx, dhy, dhz};
    k_gsrb_poisson<<<ntiles, tile_block(), 0, s>>>(tiles, vbox, A, redblack);
    return last_error();
}

int b200mg_gsrb_shell_abec (int nboxes, const b200mg_box* vbox,
                            const b200mg_fab* phi, const b200mg_fab* rhs, const b200mg_fab* a,
                            const b200mg_fab* bx, const b200mg_fab* by, const b200mg_fab* bz,
                            const b200mg_fab* f, const b200mg_ifab* m,
                            double alpha, double dhx, double dhy, double dhz, int redblack, int max_face_cells, cudaStream_t s)
{
    if (nboxes <= 0) { return 0; }
    AbecArgs A{phi, rhs, a, bx, by, bz, f, m, alpha, dhx, dhy, dhz};
    k_gsrb_shell_abec<<<dim3(nboxes * 6, shell_chunks(max_face_cells)), kShellThreads, 0, s>>>(vbox, A, redblack);
    return last_error();
}

int b200mg_gsrb_shell_poisson (int nboxes, const b200mg_box* vbox,
                               const b200mg_fab* phi, const b200mg_fab* rhs,
                               const b200mg_fab* f, const b200mg_ifab* m,
                               double dhx, double dhy, double dhz, int redblack, int max_face_cells, cudaStream_t s)
{
    if (nboxes <= 0) { return 0; }
    PoisArgs A{phi, rhs, f, m, dhx, dhy, dhz};
    k_gsrb_shell_poisson<<<dim3(nboxes * 6, shell_chunks(max_face_cells)), kShellThreads, 0, s>>>(vbox, A, redblack);
    return last_error();
}

int b200mg_gsrb_shell_abec_linked (int nboxes, const b200mg_box* vbox,
                                   const b200mg_fab* phi, const b200mg_fab* rhs, const b200mg_fab* a,
                                   const b200mg_fab* bx, const b200mg_fab* by, const b200mg_fab* bz,
                                   const b200mg_fab* f, const b200mg_ifab* m,
                                   double alpha, double dhx, double dhy, double dhz, int redblack, int max_face_cells,
                                   const b200mg_facelink* links, int push, cudaStream_t s)
{
    if (links == nullptr) { return b200mg_gsrb_shell_abec(nboxes, vbox, phi, rhs, a, bx, by, bz, f, m, alpha, dhx, dhy, dhz, redblack, max_face_cells, s); }
    if (nboxes <= 0) { return 0; }
    AbecArgs A{phi, rhs, a, bx, by, bz, f, m, alpha, dhx, dhy, dhz};
    k_gsrb_shell_abec_linked<<<dim3(nboxes * 6, shell_chunks(max_face_cells)), kShellThreads, 0, s>>>(vbox, A, redblack, links, push);
    return last_error();
}

int b200mg_gsrb_shell_poisson_linked (int nboxes, const b200mg_box* vbox,
                                      const b200mg_fab* phi, const b200mg_fab* rhs,
                                      const b200mg_fab* f, const b200mg_ifab* m,
                                      double dhx, double dhy, double dhz, int redblack, int max_face_cells,
                                      const b200mg_facelink* links, int push, cudaStream_t s)
{
    if (links == nullptr) { return b200mg_gsrb_shell_poisson(nboxes, vbox, phi, rhs, f, m, dhx, dhy, dhz, redblack, max_face_cells, s); }
    if (nboxes <= 0) { return 0; }
    PoisArgs A{phi, rhs, f, m, dhx, dhy, dhz};
    k_gsrb_shell_poisson_linked<<<dim3(nboxes * 6, shell_chunks(max_face_cells)), kShellThreads, 0, s>>>(vbox, A, redblack, links, push);
    return last_error();
}

int b200mg_adotx_abec (int ntiles, const b200mg_tile* tiles, const b200mg_box* vbox,
                       const b200mg_fab* y, const b200mg_fab* x, const b200mg_fab* rhs, const b200mg_fab* a,
                       const b200mg_fab* bx, const b200mg_fab* by, const b200mg_fab* bz,
                       double alpha, double dhx, double dhy, double dhz, cudaStream_t s)
{
    if (ntiles <= 0) { return 0; }
    k_adotx_abec<<<ntiles, tile_block(), 0, s>>>(tiles, vbox, y, x, rhs, a, bx, by, bz, alpha, dhx, dhy, dhz);
    return last_error();
}

int b200mg_adotx_poisson (int ntiles, const b200mg_tile* tiles, const b200mg_box* vbox,
                          const b200mg_fab* y, const b200mg_fab* x, const b200mg_fab* rhs,
                          double dhx, double dhy, double dhz, cudaStream_t s)
{
    if (ntiles <= 0) { return 0; }
    k_adotx_poisson<<<ntiles, tile_block(), 0, s>>>(tiles, vbox, y, x, rhs, dhx, dhy, dhz);
    return last_error();
}

int b200mg_normalize_abec (int ntiles, const b200mg_tile* tiles, const b200mg_box* vbox,
                           const b200mg_fab* x, const b200mg_fab* a,
                           const b200mg_fab* bx, const b200mg_fab* by, const b200mg_fab* bz,
                           double alpha, double dhx, double dhy, double dhz, cudaStream_t s)
{
    if (ntiles <= 0) { return 0; }
    k_normalize_abec<<<ntiles, tile_block(), 0, s>>>(tiles, vbox, x, a, bx, by, bz, alpha, dhx, dhy, dhz);
    return last_error();
}

int b200mg_apply_bc (int nfaces, const b200mg_bcface* faces, const b200mg_box* vbox,
                     const b200mg_fab* phi, const b200mg_ifab* m, const b200mg_fab* bcval,
                     int maxorder, double dxinv0, double dxinv1, double dxinv2, int inhomog, int max_face_cells, cudaStream_t s)
{
    if (nfaces <= 0) { return 0; }
    // blockIdx.y chunks: 2 ghost cells per thread on the largest face (latency bound otherwise); <= 0: unknown, 8 chunks
    int chunks = 8;
    if (max_face_cells > 0) { chunks = (max_face_cells + 255) / 256; chunks = chunks < 1 ? 1 : (chunks > 64 ? 64 : chunks); }
    k_apply_bc<<<dim3(nfaces, chunks), 128, 0, s>>>(faces, vbox, phi, m, bcval, maxorder, dxinv0, dxinv1, dxinv2, inhomog);
    return last_error();
}

int b200mg_apply_innu (int nfaces, const b200mg_bcface* faces, const b200mg_box* vbox,
                       const b200mg_fab* const out3[3], const b200mg_fab* const b3[3],
                       const b200mg_ifab* m, const b200mg_fab* bcval, const double fac[3], const int on_face[6], int mode, cudaStream_t s)
{
    if (nfaces <= 0) { return 0; }
    Innu P;
    for (int d = 0; d < 3; ++d) { P.out[d] = out3[d]; P.b[d] = b3 ? b3[d] : nullptr; P.fac[d] = fac[d]; }
    for (int f = 0; f < 6; ++f) { P.on_face[f] = on_face[f]; }
    P.mode = mode;
    k_apply_innu<<<dim3(nfaces, 8), 128, 0, s>>>(faces, vbox, m, bcval, P);
    return last_error();
}

int b200mg_robin (int nfaces, const b200mg_bcface* faces, const b200mg_box* vbox,
                  const b200mg_fab* const out3[3], const b200mg_fab* const b3[3], const b200mg_fab* phi,
                  const b200mg_ifab* m, const b200mg_fab* ra, const b200mg_fab* rb, const b200mg_fab* rf,
                  const double fac[3], const double dxinv[3], const int on_face[6], int mode, cudaStream_t s)
{
    if (nfaces <= 0) { return 0; }
    Robin P;
    for (int d = 0; d < 3; ++d) { P.out[d] = out3[d]; P.b[d] = b3 ? b3[d] : nullptr; P.fac[d] = fac[d]; P.dxi[d] = dxinv[d]; }
    for (int f = 0; f < 6; ++f) { P.on_face[f] = on_face[f]; }
    P.phi = phi; P.mode = mode;
    k_robin<<<dim3(nfaces, 8), 128, 0, s>>>(faces, vbox, m, ra, rb, rf, P);
    return last_error();
}

int b200mg_comp_interp_coef0 (int nfaces, const b200mg_bcface* faces, const b200mg_box* vbox,
                              const b200mg_fab* f, const b200mg_ifab* m,
                              int maxorder, double dxinv0, double dxinv1, double dxinv2, cudaStream_t s)
{
    if (nfaces <= 0) { return 0; }
    k_comp_interp_coef0<<<dim3(nfaces, 8), 128, 0, s>>>(faces, vbox, f, m, maxorder, dxinv0, dxinv1, dxinv2);
    return last_error();
}

int b200mg_adotx_abec_pairs (int ntiles, const b200mg_tile* tiles, const b200mg_box* vbox,
                             const b200mg_fab* y, const b200mg_fab* x, const b200mg_fab* rhs, const b200mg_fab* a,
                             const b200mg_fab* bx, const b200mg_fab* by, const b200mg_fab* bz,
                             double alpha, double dhx, double dhy, double dhz, double* norminf, cudaStream_t s)
{
    if (norminf) { const cudaError_t e = cudaMemsetAsync(norminf, 0, sizeof(double), s); if (e != cudaSuccess) { return int(e); } }
    if (ntiles <= 0) { return 0; }
    if (norminf) { k_adotx_pair<true, true><<<ntiles, tile_block(), 0, s>>>(tiles, vbox, y, x, rhs, a, bx, by, bz, alpha, dhx, dhy, dhz, norminf); }
    else { k_adotx_pair<true, false><<<ntiles, tile_block(), 0, s>>>(tiles, vbox, y, x, rhs, a, bx, by, bz, alpha, dhx, dhy, dhz, nullptr); }
    return last_error();
}

int b200mg_adotx_poisson_pairs (int ntiles, const b200mg_tile* tiles, const b200mg_box* vbox,
                                const b200mg_fab* y, const b200mg_fab* x, const b200mg_fab* rhs,
                                double dhx, double dhy, double dhz, double* norminf, cudaStream_t s)
{
    if (norminf) { const cudaError_t e = cudaMemsetAsync(norminf, 0, sizeof(double), s); if (e != cudaSuccess) { return int(e); } }
    if (ntiles <= 0) { return 0; }
    if (norminf) { k_adotx_pair<false, true><<<ntiles, tile_block(), 0, s>>>(tiles, vbox, y, x, rhs, nullptr, nullptr, nullptr, nullptr, 0.0, dhx, dhy, dhz, norminf); }
    else { k_adotx_pair<false, false><<<ntiles, tile_block(), 0, s>>>(tiles, vbox, y, x, rhs, nullptr, nullptr, nullptr, nullptr, 0.0, dhx, dhy, dhz, nullptr); }
    return last_error();
}

int b200mg_residual_restrict_abec (int ntiles, const b200mg_tile* tiles, const b200mg_box* vbox,
                                   const b200mg_fab* crse, const b200mg_fab* x, const b200mg_fab* rhs, const b200mg_fab* a,
                                   const b200mg_fab* bx, const b200mg_fab* by, const b200mg_fab* bz,
                                   double alpha, double dhx, double dhy, double dhz, cudaStream_t s)
{
    if (ntiles <= 0) { return 0; }
    k_adotx_pair_restrict<true><<<ntiles, tile_block(), 0, s>>>(tiles, vbox, crse, x, rhs, a, bx, by, bz, alpha, dhx, dhy, dhz);
    return last_error();
}

int b200mg_residual_restrict_poisson (int ntiles, const b200mg_tile* tiles, const b200mg_box* vbox,
                                      const b200mg_fab* crse, const b200mg_fab* x, const b200mg_fab* rhs,
                                      double dhx, double dhy, double dhz, cudaStream_t s)
{
    if (ntiles <= 0) { return 0; }
    k_adotx_pair_restrict<false><<<ntiles, tile_block(), 0, s>>>(tiles, vbox, crse, x, rhs, nullptr, nullptr, nullptr, nullptr, 0.0, dhx, dhy, dhz);
    return last_error();
}

int b200mg_jacobi_abec (int ntiles, const b200mg_tile* tiles, const b200mg_box* vbox,
                        const b200mg_fab* phi_out, const b200mg_fab* phi_in, const b200mg_fab* rhs, const b200mg_fab* a,
                        const b200mg_fab* bx, const b200mg_fab* by, const b200mg_fab* bz,
                        const b200mg_fab* f, const b200mg_ifab* m, double alpha, double dhx, double dhy, double dhz,
                        double adx, double ady, double adz, cudaStream_t s)
{
    if (ntiles <= 0) { return 0; }
    const AbecArgs A{phi_in, rhs, a, bx, by, bz, f, m, alpha, dhx, dhy, dhz};
    k_jacobi_abec<<<ntiles, tile_block(), 0, s>>>(tiles, vbox, phi_out, A, JacobiDh{adx, ady, adz});
    return last_error();
}

int b200mg_jacobi_poisson (int ntiles, const b200mg_tile* tiles, const b200mg_box* vbox,
                           const b200mg_fab* phi_out, const b200mg_fab* phi_in, const b200mg_fab* rhs,
                           const b200mg_fab* f, const b200mg_ifab* m, double dhx, double dhy, double dhz, cudaStream_t s)
{
    if (ntiles <= 0) { return 0; }
    const PoisArgs A{phi_in, rhs, f, m, dhx, dhy, dhz};
    k_jacobi_poisson<<<ntiles, tile_block(), 0, s>>>(tiles, vbox, phi_out, A);
    return last_error();
}

const char* b200mg_version (void) { return "amrex_b200 0.1 (sm_100a)"; }

} // extern "C"
