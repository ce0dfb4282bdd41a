// z-marching colour sweep of the red-black Gauss-Seidel smoother (one colour per launch, the reference's own schedule:
// MLCellLinOpT::smooth AMReX_MLCellLinOp.H:1206-1217; kernels abec_gsrb AMReX_MLABecLap_3D_K.H:210-264 and
// mlpoisson_gsrb AMReX_MLPoisson_3D_K.H:155-196).
//
// Same thread map as the z-marching residual kernel (stencil.cu, k_adotx_pair): a thread owns the cell pair (i0, i0+1) of
// one row and streams through the planes of its tile.  In every plane exactly one cell of the pair has the sweep's colour,
// so no lane idles.  The phi pairs of planes k-1, k, k+1 travel in registers (16-byte loads, one new plane per step); x/y
// neighbours and the coefficients of the active cell are 8-byte loads that hit lines the other colour / neighbouring rows
// bring into L1.  No shared memory and no block synchronisation: CTAs drift apart and keep HBM saturated.
//
// Face relaxation coefficients (cf0..cf5 of the reference): the x faces only concern the first / last lane of a row and
// are handled with predicated loads; y / z faces are warp-uniform (a warp covers half a row of one plane) and take a
// uniform branch.  The expression tree of g_m_d is the reference's: adding the y / z terms as exact zeros leaves the bits
// unchanged, so this kernel, the generic colour kernel and the fused kernel agree bit for bit.
//
// HBM traffic per cell of the level and colour: phi 8 (read, line granular) + phi write-back 8 + rhs 8 + a 8 + b 24 = 56 B
// line-granular (44 B strictly algorithmic); Poisson 24 (16).
#include "common.cuh"
#include "stencil_math.cuh"

using namespace b200mg;

namespace {

struct PairArgs {
    const b200mg_fab *phi, *rhs, *a, *bx, *by, *bz, *f;
    const b200mg_ifab* m;
    double alpha, dhx, dhy, dhz;
};

// slab lookups for one face: value f at the face cell if the ghost cell beyond it is an uncovered boundary cell
__device__ __forceinline__ double face_cf (const b200mg_fab& f, const b200mg_ifab& m, int i, int j, int k, int gi, int gj, int gk)
{
    return (view(m)(gi, gj, gk) > 0) ? view(f)(i, j, k) : 0.0;
}

template <bool ABEC>
__global__ void __launch_bounds__(kTileTX * B200MG_TILE_Y, 4)
k_gsrb_pair (const b200mg_tile* __restrict__ tiles, const b200mg_box* __restrict__ vbox, PairArgs A, int redblack)
{
    const b200mg_tile t = tiles[blockIdx.x];
    const b200mg_box vb = vbox[t.box];
    const int i0 = vb.lo[0] + 2 * int(threadIdx.x);
    const int j = t.j0 + int(threadIdx.y);
    if (i0 >= vb.hi[0] || j > vb.hi[1]) { return; }
    const int k0 = t.k0, k1 = min(t.k0 + tile_nk(t) - 1, vb.hi[2]);

    const auto phi = view(A.phi[t.box]);
    const int p_js = int(phi.js), p_ks = int(phi.ks);
    double* pp = phi.ptr(i0, j, k0);
    const auto rv = view(A.rhs[t.box]);
    const double* pr = rv.ptr(i0, j, k0); const int r_ks = int(rv.ks);
    const double *pa = nullptr, *pbx = nullptr, *pby = nullptr, *pbz = nullptr;
    int a_ks = 0, bx_ks = 0, by_ks = 0, by_js = 0, bz_ks = 0;
    if constexpr (ABEC) {
        const auto a = view(A.a[t.box]); const auto bx = view(A.bx[t.box]); const auto by = view(A.by[t.box]); const auto bz = view(A.bz[t.box]);
        pa = a.ptr(i0, j, k0); pbx = bx.ptr(i0, j, k0); pby = by.ptr(i0, j, k0); pbz = bz.ptr(i0, j, k0);
        a_ks = int(a.ks); bx_ks = int(bx.ks); by_ks = int(by.ks); by_js = int(by.js); bz_ks = int(bz.ks);
    }
    const b200mg_fab* f6 = A.f + 6 * t.box;
    const b200mg_ifab* m6 = A.m + 6 * t.box;
    const bool first = (i0 == vb.lo[0]), last = (i0 + 1 == vb.hi[0]);
    const bool jlo = (j == vb.lo[1]), jhi = (j == vb.hi[1]);

    int c = (i0 + j + k0 + redblack) & 1;            // which cell of the pair carries the colour in plane k0
    double2 zm = *reinterpret_cast<const double2*>(pp - p_ks);
    double2 xc = *reinterpret_cast<const double2*>(pp);
    for (int k = k0; k <= k1; ++k) {
        const double2 zp = *reinterpret_cast<const double2*>(pp + p_ks);
        const double xo = pp[c ? 2 : -1];            // the x neighbour outside the pair
        const double ym = pp[c - p_js], yp = pp[c + p_js];
        const double p = c ? xc.y : xc.x;
        const double xm = c ? xc.x : xo, xp = c ? xo : xc.y;
        const double zlo = c ? zm.y : zm.x, zhi = c ? zp.y : zp.x;
        const double rhs = __ldg(pr + c);
        const int i = i0 + c;

        double cf0 = 0.0, cf3 = 0.0;
        if (first && c == 0) { cf0 = face_cf(f6[0], m6[0], i, j, k, i - 1, j, k); }
        if (last && c == 1) { cf3 = face_cf(f6[3], m6[3], i, j, k, i + 1, j, k); }
        const bool klo = (k == vb.lo[2]), khi = (k == vb.hi[2]);
        const bool yz_surface = jlo || jhi || klo || khi;      // warp-uniform
        double cf1 = 0.0, cf2 = 0.0, cf4 = 0.0, cf5 = 0.0;
        if (yz_surface) {
            if (jlo) { cf1 = face_cf(f6[1], m6[1], i, j, k, i, j - 1, k); }
            if (klo) { cf2 = face_cf(f6[2], m6[2], i, j, k, i, j, k - 1); }
            if (jhi) { cf4 = face_cf(f6[4], m6[4], i, j, k, i, j + 1, k); }
            if (khi) { cf5 = face_cf(f6[5], m6[5], i, j, k, i, j, k + 1); }
        }

        double v;
        if constexpr (ABEC) {
            const double a = __ldg(pa + c);
            const double bxm = __ldg(pbx + c), bxp = __ldg(pbx + c + 1);
            const double bym = __ldg(pby + c), byp = __ldg(pby + c + by_js);
            const double bzm = __ldg(pbz + c), bzp = __ldg(pbz + c + bz_ks);
            // abec_gsrb (AMReX_MLABecLap_3D_K.H:225-263) with the reference's association order
            const double gamma = A.alpha * a + A.dhx * (bxm + bxp) + A.dhy * (bym + byp) + A.dhz * (bzm + bzp);
            double corr = A.dhx * (bxm * cf0 + bxp * cf3);
            if (yz_surface) { corr = corr + A.dhy * (bym * cf1 + byp * cf4) + A.dhz * (bzm * cf2 + bzp * cf5); }
            const double g_m_d = gamma - corr;
            const double rho = A.dhx * (bxm * xm + bxp * xp) + A.dhy * (bym * ym + byp * yp) + A.dhz * (bzm * zlo + bzp * zhi);
            const double res = rhs - (gamma * p - rho);
            v = p + kOmega / g_m_d * res;
            pa += a_ks; pbx += bx_ks; pby += by_ks; pbz += bz_ks;
        } else {
            // mlpoisson_gsrb (AMReX_MLPoisson_3D_K.H:171-195)
            const double gamma = -2.0 * (A.dhx + A.dhy + A.dhz);
            double g_m_d = gamma + A.dhx * (cf0 + cf3);
            if (yz_surface) { g_m_d = g_m_d + A.dhy * (cf1 + cf4) + A.dhz * (cf2 + cf5); }
            const double res = rhs - gamma * p - A.dhx * (xm + xp) - A.dhy * (ym + yp) - A.dhz * (zlo + zhi);
            v = p + kOmega / g_m_d * res;
        }
        pp[c] = v;
        zm = xc; xc = zp;                            // the other colour of these planes is not touched by this sweep
        pp += p_ks; pr += r_ks;
        c ^= 1;
    }
}


// ------------------------------------------------------------------------------------------------------------------
// Lean variant of the same sweep for levels whose boxes all have nx >= 4 and ny >= 2 (everything but the last coarse
// levels).  ncu on the kernel above (profiles/r01_s9_gsrb_pair_ncu.txt): 237 instructions per updated cell, 30 % issue
// utilisation, long-scoreboard bound -- the x-face lanes drag every warp through descriptor -> mask -> coefficient load
// chains with 64-bit index arithmetic every other plane, selects on the runtime colour bit, and 64-bit cursors that
// spill.  Here: the colour bit is a template parameter (the k loop is unrolled by two), every array is addressed by a
// uniform base and ONE 32-bit element offset per thread, and the x / y face slabs get per-thread cursors set up before
// the loop, so the steady-state plane costs loads + arithmetic only.  z faces (first / last plane of a box) keep the
// descriptor path.  Arithmetic and association order are unchanged => identical bits.
struct LeanCursor {
    // element offsets of (i0, j, k) from the fab bases, advanced by one plane per step.  rhs and a share a layout
    // (cell-centred, no ghost cells); a face's mask slab (outside) and coefficient slab (inside) have the same shape.
    int phi, cc, bx, by, bz;
    int xs, ys;                  // x-face / y-face slab cursors
};

template <bool ABEC, int C>
__device__ __forceinline__ void
lean_step (const PairArgs& A, double* __restrict__ phi, int p_js, int p_ks,
           const double* __restrict__ rhs, const double* __restrict__ a, const double* __restrict__ bx,
           const double* __restrict__ by, const double* __restrict__ bz, int by_js, int bz_ks,
           const int* __restrict__ mxp, const double* __restrict__ fxp, const int* __restrict__ myp, const double* __restrict__ fyp,
           bool xface_lo, bool xface_hi, bool jlo, bool jhi, double cf2, double cf5, bool z_surface,
           const LeanCursor& o, const double2& zm, const double2& xc, double2& zp)
{
    double* pp = phi + o.phi;
    zp = *reinterpret_cast<const double2*>(pp + p_ks);
    const double xo = pp[C ? 2 : -1];
    const double ym = pp[C - p_js], yp = pp[C + p_js];
    const double p = C ? xc.y : xc.x;
    const double xm = C ? xc.x : xo, xp = C ? xo : xc.y;
    const double zlo = C ? zm.y : zm.x, zhi = C ? zp.y : zp.x;
    const double r = __ldg(rhs + o.cc + C);

    // x faces: cell 0 of the first pair / cell 1 of the last pair (independent loads, select afterwards)
    double cf0 = 0.0, cf3 = 0.0;
    if (C == 0) { if (xface_lo) { const int mk = mxp[o.xs]; const double f = fxp[o.xs]; cf0 = (mk > 0) ? f : 0.0; } }
    else        { if (xface_hi) { const int mk = mxp[o.xs]; const double f = fxp[o.xs]; cf3 = (mk > 0) ? f : 0.0; } }
    const bool yz_surface = jlo || jhi || z_surface;       // warp-uniform
    double cf1 = 0.0, cf4 = 0.0;
    if (jlo || jhi) {
        const int mk = myp[o.ys + C]; const double f = fyp[o.ys + C];
        const double cf = (mk > 0) ? f : 0.0;
        if (jlo) { cf1 = cf; } else { cf4 = cf; }
    }

    double v;
    if constexpr (ABEC) {
        const double av = __ldg(a + o.cc + C);
        const double bxm = __ldg(bx + o.bx + C), bxp = __ldg(bx + o.bx + C + 1);
        const double bym = __ldg(by + o.by + C), byp = __ldg(by + o.by + C + by_js);
        const double bzm = __ldg(bz + o.bz + C), bzp = __ldg(bz + o.bz + C + bz_ks);
        const double gamma = A.alpha * av + A.dhx * (bxm + bxp) + A.dhy * (bym + byp) + A.dhz * (bzm + bzp);
        double corr = A.dhx * (bxm * cf0 + bxp * cf3);
        if (yz_surface) { corr = corr + A.dhy * (bym * cf1 + byp * cf4) + A.dhz * (bzm * cf2 + bzp * cf5); }
        const double g_m_d = gamma - corr;
        const double rho = A.dhx * (bxm * xm + bxp * xp) + A.dhy * (bym * ym + byp * yp) + A.dhz * (bzm * zlo + bzp * zhi);
        const double res = r - (gamma * p - rho);
        v = p + kOmega / g_m_d * res;
    } else {
        const double gamma = -2.0 * (A.dhx + A.dhy + A.dhz);
        double g_m_d = gamma + A.dhx * (cf0 + cf3);
        if (yz_surface) { g_m_d = g_m_d + A.dhy * (cf1 + cf4) + A.dhz * (cf2 + cf5); }
        const double res = r - gamma * p - A.dhx * (xm + xp) - A.dhy * (ym + yp) - A.dhz * (zlo + zhi);
        v = p + kOmega / g_m_d * res;
    }
    pp[C] = v;
}

template <bool ABEC, int MINB>
__global__ void __launch_bounds__(kTileTX * B200MG_TILE_Y, MINB)
k_gsrb_pair_lean (const b200mg_tile* __restrict__ tiles, const b200mg_box* __restrict__ vbox, PairArgs A, int redblack)
{
    const b200mg_tile t = tiles[blockIdx.x];
    const b200mg_box vb = vbox[t.box];
    const int i0 = vb.lo[0] + 2 * int(threadIdx.x);
    const int j = t.j0 + int(threadIdx.y);
    if (i0 >= vb.hi[0] || j > vb.hi[1]) { return; }
    const int k0 = t.k0, k1 = min(t.k0 + tile_nk(t) - 1, vb.hi[2]);

    const auto phi = view(A.phi[t.box]);
    const int p_js = int(phi.js), p_ks = int(phi.ks);
    const auto rv = view(A.rhs[t.box]);
    LeanCursor o;
    o.phi = int(phi.ptr(i0, j, k0) - phi.p);
    o.cc = int(rv.ptr(i0, j, k0) - rv.p);
    const int cc_ks = int(rv.ks);
    const double *pa = nullptr, *pbx = nullptr, *pby = nullptr, *pbz = nullptr;
    int bx_ks = 0, by_ks = 0, by_js = 0, bz_ks = 0;
    o.bx = o.by = o.bz = 0;
    if constexpr (ABEC) {
        const auto a = view(A.a[t.box]); const auto bx = view(A.bx[t.box]); const auto by = view(A.by[t.box]); const auto bz = view(A.bz[t.box]);
        pa = a.p; pbx = bx.p; pby = by.p; pbz = bz.p;
        o.bx = int(bx.ptr(i0, j, k0) - bx.p); o.by = int(by.ptr(i0, j, k0) - by.p); o.bz = int(bz.ptr(i0, j, k0) - bz.p);
        bx_ks = int(bx.ks); by_ks = int(by.ks); by_js = int(by.js); bz_ks = int(bz.ks);
    }
    const b200mg_fab* f6 = A.f + 6 * t.box;
    const b200mg_ifab* m6 = A.m + 6 * t.box;
    // nx >= 4: a pair is first or last, never both; ny >= 2: a row is the low or the high y face, never both
    const bool first = (i0 == vb.lo[0]), last = (i0 + 1 == vb.hi[0]);
    const bool jlo = (j == vb.lo[1]), jhi = (j == vb.hi[1]);
    const int* mxp = nullptr; const double* fxp = nullptr; int xs_ks = 0;
    o.xs = o.ys = 0;
    if (first || last) {
        const int fc = last ? 3 : 0;
        const auto fv = view(f6[fc]);
        mxp = m6[fc].p; fxp = fv.p; xs_ks = int(fv.ks);
        o.xs = int(fv.ptr(last ? i0 + 1 : i0, j, k0) - fv.p);
    }
    const int* myp = nullptr; const double* fyp = nullptr; int ys_ks = 0;
    if (jlo || jhi) {
        const int fc = jhi ? 4 : 1;
        const auto fv = view(f6[fc]);
        myp = m6[fc].p; fyp = fv.p; ys_ks = int(fv.ks);
        o.ys = int(fv.ptr(i0, j, k0) - fv.p);
    }
    // z faces: coefficient of this thread's coloured cell in the first / last plane of the box, if the tile holds them
    const int c0 = (i0 + j + k0 + redblack) & 1;          // which cell of the pair carries the colour in plane k0
    double cf2v = 0.0, cf5v = 0.0;
    if (k0 == vb.lo[2]) { cf2v = face_cf(f6[2], m6[2], i0 + c0, j, k0, i0 + c0, j, k0 - 1); }
    if (k1 == vb.hi[2]) { const int c1 = (c0 + k1 - k0) & 1; cf5v = face_cf(f6[5], m6[5], i0 + c1, j, k1, i0 + c1, j, k1 + 1); }
    const int kzlo = (k0 == vb.lo[2]) ? k0 : -(1 << 30), kzhi = (k1 == vb.hi[2]) ? k1 : -(1 << 30);

    double* pphi = phi.p;
    double2 zm = *reinterpret_cast<const double2*>(pphi + o.phi - p_ks);
    double2 xc = *reinterpret_cast<const double2*>(pphi + o.phi);
    double2 zp;
    auto advance = [&] () {
        o.phi += p_ks; o.cc += cc_ks;
        if constexpr (ABEC) { o.bx += bx_ks; o.by += by_ks; o.bz += bz_ks; }
        o.xs += xs_ks; o.ys += ys_ks;
        zm = xc; xc = zp;
    };
#define B200MG_LEAN_STEP(CC, K) lean_step<ABEC, CC>(A, pphi, p_js, p_ks, rv.p, pa, pbx, pby, pbz, by_js, bz_ks, mxp, fxp, myp, fyp, \
                                                    first, last, jlo, jhi, ((K) == kzlo) ? cf2v : 0.0, ((K) == kzhi) ? cf5v : 0.0, \
                                                    ((K) == kzlo) || ((K) == kzhi), o, zm, xc, zp)
    int k = k0;
    if (c0) {
        for (; k + 1 <= k1; k += 2) { B200MG_LEAN_STEP(1, k); advance(); B200MG_LEAN_STEP(0, k + 1); advance(); }
        if (k <= k1) { B200MG_LEAN_STEP(1, k); }
    } else {
        for (; k + 1 <= k1; k += 2) { B200MG_LEAN_STEP(0, k); advance(); B200MG_LEAN_STEP(1, k + 1); advance(); }
        if (k <= k1) { B200MG_LEAN_STEP(0, k); }
    }
#undef B200MG_LEAN_STEP
}

int g_lean_minb = 4;     // measured (profiles/r01_s11_tune_lean.txt): 1.26 ms at 4 CTAs/SM vs 1.35 ms at 3, generic sweep 1.43 ms

} // namespace

extern "C" {

int b200mg_gsrb_abec_pairs (int ntiles, const b200mg_tile* tiles, const b200mg_box* vbox,
                            const b200mg_fab* phi, const b200mg_fab* rhs, const b200mg_fab* a,
                            const b200mg_fab* bx, const b200mg_fab* by, const b200mg_fab* bz,
                            const b200mg_fab* f, const b200mg_ifab* m,
                            double alpha, double dhx, double dhy, double dhz, int redblack, cudaStream_t s)
{
    if (ntiles <= 0) { return 0; }
    PairArgs A{phi, rhs, a, bx, by, bz, f, m, alpha, dhx, dhy, dhz};
    k_gsrb_pair<true><<<ntiles, tile_block(), 0, s>>>(tiles, vbox, A, redblack);
    return last_error();
}

int b200mg_gsrb_abec_pairs_lean (int ntiles, const b200mg_tile* tiles, const b200mg_box* vbox,
                                 const b200mg_fab* phi, const b200mg_fab* rhs, const b200mg_fab* a,
                                 const b200mg_fab* bx, const b200mg_fab* by, const b200mg_fab* bz,
                                 const b200mg_fab* f, const b200mg_ifab* m,
                                 double alpha, double dhx, double dhy, double dhz, int redblack, cudaStream_t s)
{
    if (ntiles <= 0) { return 0; }
    PairArgs A{phi, rhs, a, bx, by, bz, f, m, alpha, dhx, dhy, dhz};
    if (g_lean_minb <= 0) { k_gsrb_pair<true><<<ntiles, tile_block(), 0, s>>>(tiles, vbox, A, redblack); }
    else if (g_lean_minb >= 4) { k_gsrb_pair_lean<true, 4><<<ntiles, tile_block(), 0, s>>>(tiles, vbox, A, redblack); }
    else { k_gsrb_pair_lean<true, 3><<<ntiles, tile_block(), 0, s>>>(tiles, vbox, A, redblack); }
    return last_error();
}

int b200mg_gsrb_poisson_pairs (int ntiles, const b200mg_tile* tiles, const b200mg_box* vbox,
                               const b200mg_fab* phi, const b200mg_fab* rhs,
                               const b200mg_fab* f, const b200mg_ifab* m,
                               double dhx, double dhy, double dhz, int redblack, cudaStream_t s)
{
    if (ntiles <= 0) { return 0; }
    PairArgs A{phi, rhs, nullptr, nullptr, nullptr, nullptr, f, m, 0.0, dhx, dhy, dhz};
    k_gsrb_pair<false><<<ntiles, tile_block(), 0, s>>>(tiles, vbox, A, redblack);
    return last_error();
}

int b200mg_gsrb_poisson_pairs_lean (int ntiles, const b200mg_tile* tiles, const b200mg_box* vbox,
                                    const b200mg_fab* phi, const b200mg_fab* rhs,
                                    const b200mg_fab* f, const b200mg_ifab* m,
                                    double dhx, double dhy, double dhz, int redblack, cudaStream_t s)
{
    if (ntiles <= 0) { return 0; }
    PairArgs A{phi, rhs, nullptr, nullptr, nullptr, nullptr, f, m, 0.0, dhx, dhy, dhz};
    if (g_lean_minb <= 0) { k_gsrb_pair<false><<<ntiles, tile_block(), 0, s>>>(tiles, vbox, A, redblack); }
    else if (g_lean_minb >= 4) { k_gsrb_pair_lean<false, 4><<<ntiles, tile_block(), 0, s>>>(tiles, vbox, A, redblack); }
    else { k_gsrb_pair_lean<false, 3><<<ntiles, tile_block(), 0, s>>>(tiles, vbox, A, redblack); }
    return last_error();
}

/* resident CTAs per SM the lean sweep is compiled for: 4 (64 registers) or 3 (80 registers); <= 0: use the generic pair sweep */
void b200mg_set_gsrb_lean_occupancy (int min_blocks) { g_lean_minb = min_blocks; }

} // extern "C"
