// The coarse leg of a V-cycle in ONE kernel (reference control flow: MLMGT::mgVcycle, AMReX_MLMG.H:1308-1415, and
// MLMGT::bottomSolve / actualBottomSolve, :1460-1576).
//
// From the first MG level that is (or has been merged into) a single box covering the whole domain down to the bottom and
// back up, the launch-per-operation schedule issues ~35 kernels of 2-10 us per level and V-cycle (boundary fill, two
// colour sweeps per smooth, residual, restriction, prolongation), every one bound by launch and drain latency, not by
// work; on 8 GPUs these levels cost as much as on one.  Here ONE cooperative grid runs the whole leg: pre-smooths,
// residual and restriction level by level, the BiCGStab bottom solve (bottom_solve.cuh, by CTA 0), then prolongation and
// post-smooths back up.  "Wide" levels (more than narrow_cells cells) are worked on by every CTA with a grid barrier
// between phases; the small levels below are left to CTA 0 alone, whose phases cost a block barrier each, while the other
// CTAs wait at the one grid barrier that precedes the first wide level on the way up.  Fields stay in L2.  Per-cell
// arithmetic is the shared stencil_math.cuh code and the sequence of operations is the host schedule's, so the leg
// leaves the bits the launch-per-operation path leaves.
#include "bottom_solve.cuh"

#include <cooperative_groups.h>
#include <cstdlib>

namespace cg = cooperative_groups;
using namespace b200mg;

namespace {

constexpr int kLegThreads = kBottomThreads;     // CTA size (the bottom solve's reduction order is tied to it)

struct Team { int tid, nth; };

// phase time stamps (tuning aid): thread 0 of CTA 0 appends (id << 48 | clock) after a phase's barrier
struct Stamps {
    unsigned long long* buf; int n;
    __device__ __forceinline__ void mark (int id)
    {
        if (buf != nullptr && n < B200MG_LEG_MAX_STAMPS) { buf[n++] = ((unsigned long long)id << 48) | ((unsigned long long)clock64() & 0xffffffffffffull); buf[0] = (unsigned long long)n; }
    }
};

// Grid barrier of the cooperative launch (all CTAs co-resident): one counter in global memory that only grows - CTA b's
// thread 0 adds 1 and spins until the count reaches the barrier's target (epoch * CTAs), block barriers on both sides, the
// fences make the CTA's writes visible to the grid and invalidate this SM's L1 before the CTA reads other CTAs' data.
// Measured ~3x cheaper than cooperative_groups' grid.sync() here (B200MG_LEG_CG_SYNC=1 selects that one for comparison).
struct GridBar { unsigned* counter; unsigned target; unsigned nctas; int use_cg; };

__device__ __forceinline__ void grid_barrier (GridBar& B)
{
    if (B.use_cg) { cg::this_grid().sync(); return; }
    __syncthreads();
    if (threadIdx.x == 0) {
        B.target += B.nctas;
        __threadfence();
        atomicAdd(B.counter, 1u);
        unsigned v;
        do { asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(B.counter) : "memory"); } while (v < B.target);
        __threadfence();
    }
    __syncthreads();
}

// WIDE: the team is the whole grid, else the calling CTA
template <bool WIDE> __device__ __forceinline__ void team_sync (GridBar& B)
{
    if constexpr (WIDE) { grid_barrier(B); } else { __syncthreads(); }
}

template <class F>
__device__ __forceinline__ void for_box (const b200mg_box& b, int ng, const Team& T, F&& f)
{
    const int nx = b.hi[0] - b.lo[0] + 1 + 2 * ng, ny = b.hi[1] - b.lo[1] + 1 + 2 * ng, nz = b.hi[2] - b.lo[2] + 1 + 2 * ng;
    const int n = nx * ny * nz;
    for (int c = T.tid; c < n; c += T.nth) {
        const int i = c % nx, jk = c / nx;
        f(b.lo[0] - ng + i, b.lo[1] - ng + jk % ny, b.lo[2] - ng + jk / ny);
    }
}

__device__ __forceinline__ void zero_field (const b200mg_leg_level& L, const Team& T)
{
    const auto x = view(L.cor);
    for_box(L.vb, 1, T, [&] (int i, int j, int k) { x(i, j, k) = 0.0; });
}

// Boundary conditions of a one-box level by face orientation.  The leg never stores ghost values: a cell on the box surface
// computes the value the boundary fill (bc_fill = mllinop_apply_bc_*, or the periodic wrap) would have left in the ghost
// cell beyond its face, from the same operands in the same order - the same bits, with half the barriers per smooth and no
// O(n^2) phases.  (The ghost cells of cor therefore hold stale values after the leg; nothing reads them: prolongation and
// the copy back to a chopped level take valid cells only.)
struct FaceBC {
    int type[6];                            // kBcDirichletB / kBcNeumannB / kBcReflectOddB, or kPeriodicFace
    int nx[6];                              // Dirichlet: interpolation nodes (ghost-side node included)
    double coef[6][4];
    double w[6][3];                         // the same fill as ONE expression: ghost = 0 + x0*w0 + x1*w1 + x2*w2 over the first
                                            // three cells inside the face (Neumann 1,0,0; reflect-odd -1,0,0; Dirichlet coef[1..3],
                                            // zeros beyond the order: adding x*0 changes no bit of a finite sum)
};
constexpr int kPeriodicFace = 200;

__device__ __forceinline__ void make_face_bc (FaceBC& F, const BoxBC& B)
{
    for (int o = 0; o < 6; ++o) { F.type[o] = B.periodic[o % 3] ? kPeriodicFace : kBcNeumannB; F.nx[o] = 0; for (int m = 0; m < 4; ++m) { F.coef[o][m] = 0.0; } }
    for (int n = 0; n < B.nfaces; ++n) {
        const int o = B.faces[n].face;
        F.type[o] = B.faces[n].bctype; F.nx[o] = B.nxo[n];
        for (int m = 0; m < 4; ++m) { F.coef[o][m] = B.coef[n][m]; }
    }
    for (int o = 0; o < 6; ++o) {
        F.w[o][0] = F.w[o][1] = F.w[o][2] = 0.0;
        if (F.type[o] == kBcNeumannB) { F.w[o][0] = 1.0; }
        else if (F.type[o] == kBcReflectOddB) { F.w[o][0] = -1.0; }
        else if (F.type[o] == kBcDirichletB) { for (int m = 1; m < F.nx[o] && m < 4; ++m) { F.w[o][m - 1] = F.coef[o][m]; } }
    }
}

// ghost value beyond face o next to the inside cell *pc; sst: signed stride from the ghost cell towards the inside,
// n: cells of the box along that direction
__device__ __forceinline__ double ghost_value (const FaceBC& F, int o, const double* pc, long long sst, int n)
{
    const int t = F.type[o];
    if (t == kPeriodicFace) { return pc[(long long)(n - 1) * sst]; }
    if (t == kBcNeumannB) { return *pc; }
    if (t == kBcReflectOddB) { return -*pc; }
    double tmp = 0.0;
    for (int m = 1; m < F.nx[o]; ++m) { tmp += pc[(long long)(m - 1) * sst] * F.coef[o][m]; }
    return tmp;
}

struct Nbrs { double c, xm, xp, ym, yp, zm, zp; };

__device__ __forceinline__ Nbrs load_nbrs (const double* pc, int i, int j, int k, const b200mg_box& vb, const FaceBC& F, long long js, long long ks)
{
    Nbrs q;
    q.c = *pc;
    q.xm = (i > vb.lo[0]) ? pc[-1] : ghost_value(F, 0, pc, 1, vb.hi[0] - vb.lo[0] + 1);
    q.xp = (i < vb.hi[0]) ? pc[1] : ghost_value(F, 3, pc, -1, vb.hi[0] - vb.lo[0] + 1);
    q.ym = (j > vb.lo[1]) ? pc[-js] : ghost_value(F, 1, pc, js, vb.hi[1] - vb.lo[1] + 1);
    q.yp = (j < vb.hi[1]) ? pc[js] : ghost_value(F, 4, pc, -js, vb.hi[1] - vb.lo[1] + 1);
    q.zm = (k > vb.lo[2]) ? pc[-ks] : ghost_value(F, 2, pc, ks, vb.hi[2] - vb.lo[2] + 1);
    q.zp = (k < vb.hi[2]) ? pc[ks] : ghost_value(F, 5, pc, -ks, vb.hi[2] - vb.lo[2] + 1);
    return q;
}

// operands of one cell update, gathered before any store so that the loads of a batch of cells overlap
template <bool ABEC>
struct CellIn {
    double* pc; Nbrs q; double rhs, a, bxm, bxp, bym, byp, bzm, bzp; int i, j, k; bool on;
};

// one colour of MLCellLinOpT::smooth's Fsmooth (abec_gsrb / mlpoisson_gsrb), cells with (i+j+k+redblack) even
template <bool ABEC>
__device__ __noinline__ void sweep (const b200mg_leg_level& L, const FaceBC& F, double alpha, int redblack, const Team& T)
{
    const auto phi = view(L.cor); const auto rhs = view(L.res);
    const long long js = phi.js, ks = phi.ks;
    const b200mg_box vb = L.vb;
    View<double> a = phi, bx = phi, by = phi, bz = phi;
    if constexpr (ABEC) { a = view(L.a); bx = view(L.bx); by = view(L.by); bz = view(L.bz); }
    // the cells of this colour only: row (j,k) holds ceil(nx/2) candidates i = i0 + 2*ii, so a warp covers 64 consecutive cells
    const int nx = vb.hi[0] - vb.lo[0] + 1, ny = vb.hi[1] - vb.lo[1] + 1, nz = vb.hi[2] - vb.lo[2] + 1;
    const int nxh = (nx + 1) >> 1;
    const int n = nxh * ny * nz;
    auto gather = [&] (int c) {
        CellIn<ABEC> in;
        in.on = false;
        if (c >= n) { return in; }
        const int ii = c % nxh, jk = c / nxh;
        const int j = vb.lo[1] + jk % ny, k = vb.lo[2] + jk / ny;
        const int i = vb.lo[0] + ((vb.lo[0] + j + k + redblack) & 1) + 2 * ii;
        if (i > vb.hi[0]) { return in; }
        in.on = true;
        in.pc = phi.ptr(i, j, k);
        in.q = load_nbrs(in.pc, i, j, k, vb, F, js, ks);
        in.rhs = rhs(i, j, k);
        if constexpr (ABEC) {
            in.a = a(i, j, k); in.bxm = bx(i, j, k); in.bxp = bx(i + 1, j, k); in.bym = by(i, j, k); in.byp = by(i, j + 1, k);
            in.bzm = bz(i, j, k); in.bzp = bz(i, j, k + 1);
        }
        in.i = i; in.j = j; in.k = k;
        return in;
    };
    auto update = [&] (const CellIn<ABEC>& in) {
        if (!in.on) { return; }
        // relaxation coefficients of the boundary faces: surface cells only (rare), fetched here to keep the batch in registers
        FaceCoefs cf;
        if (on_surface(in.i, in.j, in.k, vb)) { cf = face_coefs(in.i, in.j, in.k, vb, L.f, L.m); }
        else {
#pragma unroll
            for (int q = 0; q < 6; ++q) { cf.c[q] = 0.0; }
        }
        if constexpr (ABEC) {
            *in.pc = gsrb_abec_cell(in.q.c, in.q.xm, in.q.xp, in.q.ym, in.q.yp, in.q.zm, in.q.zp, in.rhs, in.a,
                                    in.bxm, in.bxp, in.bym, in.byp, in.bzm, in.bzp,
                                    cf.c[0], cf.c[1], cf.c[2], cf.c[3], cf.c[4], cf.c[5], alpha, L.dh[0], L.dh[1], L.dh[2]);
        } else {
            *in.pc = gsrb_poisson_cell(in.q.c, in.q.xm, in.q.xp, in.q.ym, in.q.yp, in.q.zm, in.q.zp, in.rhs,
                                       cf.c[0], cf.c[1], cf.c[2], cf.c[3], cf.c[4], cf.c[5], L.dh[0], L.dh[1], L.dh[2]);
        }
    };
    // cells of one colour do not read each other: two per iteration, all loads ahead of the stores
    for (int c = T.tid; c < n; c += 2 * T.nth) {
        const CellIn<ABEC> A = gather(c);
        const CellIn<ABEC> B = gather(c + T.nth);
        update(A); update(B);
    }
}

// ---- lean versions for levels with at least 4 cells per direction (everything above the last two or three levels): plain
// int index arithmetic on register copies of the descriptors, boundary values through FaceBC::w, no mask look-ups (a one-box
// level has no covered ghost cells except across periodic faces).  Same per-cell arithmetic, same bits.
struct Arr { const double* p; int js, ks; };                // p: element (lo0, lo1, lo2) of the array

__device__ __forceinline__ Arr arr_of (const b200mg_fab& f, const b200mg_box& vb)
{
    const auto v = view(f);
    return Arr{v.ptr(vb.lo[0], vb.lo[1], vb.lo[2]), int(v.js), int(v.ks)};
}

// ghost value beyond face o for the inside cell *pc (sst: stride towards the inside, n: box cells along the direction)
__device__ __forceinline__ double ghost_w (const FaceBC& F, int o, const double* pc, int sst, int n)
{
    if (F.type[o] == kPeriodicFace) { return pc[(n - 1) * sst]; }
    double tmp = 0.0;
    tmp += pc[0] * F.w[o][0]; tmp += pc[sst] * F.w[o][1]; tmp += pc[2 * sst] * F.w[o][2];
    return tmp;
}

struct FastBox { int nx, ny, nz, par0; };                   // par0: parity of lo0 + lo1 + lo2

template <bool ABEC>
struct FastIn { double* pc; double c, xm, xp, ym, yp, zm, zp, rhs, a, bxm, bxp, bym, byp, bzm, bzp; int i, j, k; bool on; };

template <bool ABEC>
__device__ __forceinline__ FastIn<ABEC> fast_gather (int i, int j, int k, bool on, const FastBox& B, const FaceBC& F, const Arr& P, const Arr& R,
                                                     const Arr& A, const Arr& BX, const Arr& BY, const Arr& BZ)
{
    FastIn<ABEC> in;
    in.on = on; in.i = i; in.j = j; in.k = k;
    if (!on) { return in; }
    double* pc = const_cast<double*>(P.p) + i + j * P.js + k * P.ks;
    in.pc = pc;
    in.c = *pc;
    in.xm = (i > 0) ? pc[-1] : ghost_w(F, 0, pc, 1, B.nx);
    in.xp = (i < B.nx - 1) ? pc[1] : ghost_w(F, 3, pc, -1, B.nx);
    in.ym = (j > 0) ? pc[-P.js] : ghost_w(F, 1, pc, P.js, B.ny);
    in.yp = (j < B.ny - 1) ? pc[P.js] : ghost_w(F, 4, pc, -P.js, B.ny);
    in.zm = (k > 0) ? pc[-P.ks] : ghost_w(F, 2, pc, P.ks, B.nz);
    in.zp = (k < B.nz - 1) ? pc[P.ks] : ghost_w(F, 5, pc, -P.ks, B.nz);
    in.rhs = R.p[i + j * R.js + k * R.ks];
    if constexpr (ABEC) {
        in.a = A.p[i + j * A.js + k * A.ks];
        const double* q = BX.p + i + j * BX.js + k * BX.ks; in.bxm = q[0]; in.bxp = q[1];
        q = BY.p + i + j * BY.js + k * BY.ks; in.bym = q[0]; in.byp = q[BY.js];
        q = BZ.p + i + j * BZ.js + k * BZ.ks; in.bzm = q[0]; in.bzp = q[BZ.ks];
    }
    return in;
}

// relaxation coefficient of face o at a surface cell (absolute indices): the slab value where the ghost cell beyond is a
// boundary cell, i.e. everywhere except across periodic faces
__device__ __forceinline__ double face_cf (const b200mg_leg_level& L, const FaceBC& F, int o, int i, int j, int k)
{
    return (F.type[o] == kPeriodicFace) ? 0.0 : view(L.f[o])(i, j, k);
}

// interior cell (no face of the box next to it): straight loads
template <bool ABEC>
__device__ __forceinline__ FastIn<ABEC> interior_gather (int i, int j, int k, bool on, const Arr& P, const Arr& R,
                                                         const Arr& A, const Arr& BX, const Arr& BY, const Arr& BZ)
{
    FastIn<ABEC> in;
    in.on = on; in.i = i; in.j = j; in.k = k;
    if (!on) { return in; }
    double* pc = const_cast<double*>(P.p) + i + j * P.js + k * P.ks;
    in.pc = pc;
    in.c = pc[0]; in.xm = pc[-1]; in.xp = pc[1]; in.ym = pc[-P.js]; in.yp = pc[P.js]; in.zm = pc[-P.ks]; in.zp = pc[P.ks];
    in.rhs = R.p[i + j * R.js + k * R.ks];
    if constexpr (ABEC) {
        in.a = A.p[i + j * A.js + k * A.ks];
        const double* q = BX.p + i + j * BX.js + k * BX.ks; in.bxm = q[0]; in.bxp = q[1];
        q = BY.p + i + j * BY.js + k * BY.ks; in.bym = q[0]; in.byp = q[BY.js];
        q = BZ.p + i + j * BZ.js + k * BZ.ks; in.bzm = q[0]; in.bzp = q[BZ.ks];
    }
    return in;
}

// the cells on the surface of an nx x ny x nz box, enumerated once each: x faces own their edges and corners, y faces
// exclude the x extremes, z faces the x and y extremes (nx, ny, nz >= 3)
struct SurfaceMap {
    unsigned n_x, n_y, n_z, nx, ny, nz;          // cells of the two x / y / z faces
    __device__ __forceinline__ SurfaceMap (int ax, int ay, int az)
        : n_x(2u * unsigned(ay) * unsigned(az)), n_y(2u * unsigned(ax - 2) * unsigned(az)), n_z(2u * unsigned(ax - 2) * unsigned(ay - 2)),
          nx(unsigned(ax)), ny(unsigned(ay)), nz(unsigned(az)) {}
    __device__ __forceinline__ unsigned size () const { return n_x + n_y + n_z; }
    __device__ __forceinline__ void locate (unsigned t, int& i, int& j, int& k) const
    {
        if (t < n_x) {
            const unsigned side = t & 1u, q = t >> 1;
            i = side ? int(nx) - 1 : 0; j = int(q % ny); k = int(q / ny);
        } else if (t < n_x + n_y) {
            const unsigned u = t - n_x, side = u & 1u, q = u >> 1;
            j = side ? int(ny) - 1 : 0; i = 1 + int(q % (nx - 2u)); k = int(q / (nx - 2u));
        } else {
            const unsigned u = t - n_x - n_y, side = u & 1u, q = u >> 1;
            k = side ? int(nz) - 1 : 0; i = 1 + int(q % (nx - 2u)); j = 1 + int(q / (nx - 2u));
        }
    }
};

template <bool ABEC>
__device__ __noinline__ void sweep_fast (const b200mg_leg_level& L, const FaceBC& F, double alpha, int redblack, const Team& T)
{
    const b200mg_box vb = L.vb;
    const FastBox B{vb.hi[0] - vb.lo[0] + 1, vb.hi[1] - vb.lo[1] + 1, vb.hi[2] - vb.lo[2] + 1, (vb.lo[0] + vb.lo[1] + vb.lo[2]) & 1};
    const Arr P = arr_of(L.cor, vb), R = arr_of(L.res, vb);
    Arr A = P, BX = P, BY = P, BZ = P;
    if constexpr (ABEC) { A = arr_of(L.a, vb); BX = arr_of(L.bx, vb); BY = arr_of(L.by, vb); BZ = arr_of(L.bz, vb); }
    const double dhx = L.dh[0], dhy = L.dh[1], dhz = L.dh[2];
    const int par = (B.par0 + redblack) & 1;
    // ---- interior cells of the colour: 1 <= i <= nx-2 etc.; row (j,k) holds them at i = 1 + ((par + 1 + j + k) & 1) + 2*ii
    {
        const unsigned mxh = unsigned(B.nx - 1) >> 1, my = unsigned(B.ny - 2);        // ceil((nx-2)/2) candidates per row
        const unsigned n = mxh * my * unsigned(B.nz - 2);
        auto locate = [&] (unsigned c, int& i, int& j, int& k) -> bool {
            if (c >= n) { i = j = k = 1; return false; }
            const unsigned ii = c % mxh, jk = c / mxh;
            j = 1 + int(jk % my); k = 1 + int(jk / my);
            i = 1 + ((par + 1 + j + k) & 1) + 2 * int(ii);
            return i <= B.nx - 2;
        };
        auto update = [&] (const FastIn<ABEC>& in) {
            if (!in.on) { return; }
            if constexpr (ABEC) {
                *in.pc = gsrb_abec_cell_interior(in.c, in.xm, in.xp, in.ym, in.yp, in.zm, in.zp, in.rhs, in.a, in.bxm, in.bxp, in.bym, in.byp,
                                                 in.bzm, in.bzp, alpha, dhx, dhy, dhz);
            } else {
                *in.pc = gsrb_poisson_cell(in.c, in.xm, in.xp, in.ym, in.yp, in.zm, in.zp, in.rhs, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, dhx, dhy, dhz);
            }
        };
        // cells of one colour do not read each other: two per iteration, all loads ahead of the stores
        for (unsigned c = unsigned(T.tid); c < n; c += 2u * unsigned(T.nth)) {
            int i0, j0, k0, i1, j1, k1;
            const bool on0 = locate(c, i0, j0, k0), on1 = locate(c + unsigned(T.nth), i1, j1, k1);
            const FastIn<ABEC> X = interior_gather<ABEC>(i0, j0, k0, on0, P, R, A, BX, BY, BZ);
            const FastIn<ABEC> Y = interior_gather<ABEC>(i1, j1, k1, on1, P, R, A, BX, BY, BZ);
            update(X); update(Y);
        }
    }
    // ---- surface cells of the colour: boundary values through FaceBC, relaxation coefficients from the face slabs
    {
        const SurfaceMap M(B.nx, B.ny, B.nz);
        const unsigned n = M.size();
        for (unsigned t = unsigned(T.tid); t < n; t += unsigned(T.nth)) {
            int i, j, k;
            M.locate(t, i, j, k);
            if ((i + j + k + par) & 1) { continue; }
            const FastIn<ABEC> in = fast_gather<ABEC>(i, j, k, true, B, F, P, R, A, BX, BY, BZ);
            double cf0 = 0.0, cf1 = 0.0, cf2 = 0.0, cf3 = 0.0, cf4 = 0.0, cf5 = 0.0;
            if (i == 0) { cf0 = face_cf(L, F, 0, vb.lo[0], vb.lo[1] + j, vb.lo[2] + k); }
            if (j == 0) { cf1 = face_cf(L, F, 1, vb.lo[0] + i, vb.lo[1], vb.lo[2] + k); }
            if (k == 0) { cf2 = face_cf(L, F, 2, vb.lo[0] + i, vb.lo[1] + j, vb.lo[2]); }
            if (i == B.nx - 1) { cf3 = face_cf(L, F, 3, vb.hi[0], vb.lo[1] + j, vb.lo[2] + k); }
            if (j == B.ny - 1) { cf4 = face_cf(L, F, 4, vb.lo[0] + i, vb.hi[1], vb.lo[2] + k); }
            if (k == B.nz - 1) { cf5 = face_cf(L, F, 5, vb.lo[0] + i, vb.lo[1] + j, vb.hi[2]); }
            if constexpr (ABEC) {
                *in.pc = gsrb_abec_cell(in.c, in.xm, in.xp, in.ym, in.yp, in.zm, in.zp, in.rhs, in.a, in.bxm, in.bxp, in.bym, in.byp, in.bzm, in.bzp,
                                        cf0, cf1, cf2, cf3, cf4, cf5, alpha, dhx, dhy, dhz);
            } else {
                *in.pc = gsrb_poisson_cell(in.c, in.xm, in.xp, in.ym, in.yp, in.zm, in.zp, in.rhs, cf0, cf1, cf2, cf3, cf4, cf5, dhx, dhy, dhz);
            }
        }
    }
}

template <bool ABEC>
__device__ __noinline__ void residual_fast (const b200mg_leg_level& L, const FaceBC& F, double alpha, const Team& T)
{
    const b200mg_box vb = L.vb;
    const FastBox B{vb.hi[0] - vb.lo[0] + 1, vb.hi[1] - vb.lo[1] + 1, vb.hi[2] - vb.lo[2] + 1, 0};
    const Arr P = arr_of(L.cor, vb), R = arr_of(L.res, vb), Y = arr_of(L.rescor, vb);
    Arr A = P, BX = P, BY = P, BZ = P;
    if constexpr (ABEC) { A = arr_of(L.a, vb); BX = arr_of(L.bx, vb); BY = arr_of(L.by, vb); BZ = arr_of(L.bz, vb); }
    const double dhx = L.adh[0], dhy = L.adh[1], dhz = L.adh[2];
    auto finish = [&] (const FastIn<ABEC>& in) {
        if (!in.on) { return; }
        double v;
        if constexpr (ABEC) {
            v = adotx_abec_cell(in.c, in.xm, in.xp, in.ym, in.yp, in.zm, in.zp, in.a, in.bxm, in.bxp, in.bym, in.byp, in.bzm, in.bzp, alpha, dhx, dhy, dhz);
        } else {
            v = adotx_poisson_cell(in.c, in.xm, in.xp, in.ym, in.yp, in.zm, in.zp, dhx, dhy, dhz);
        }
        const_cast<double*>(Y.p)[in.i + in.j * Y.js + in.k * Y.ks] = in.rhs + (-1.0) * v;      // Xpay(y,-1,b)
    };
    {   // interior
        const unsigned mx = unsigned(B.nx - 2), my = unsigned(B.ny - 2);
        const unsigned n = mx * my * unsigned(B.nz - 2);
        auto locate = [&] (unsigned c, int& i, int& j, int& k) -> bool {
            if (c >= n) { i = j = k = 1; return false; }
            const unsigned jk = c / mx;
            i = 1 + int(c % mx); j = 1 + int(jk % my); k = 1 + int(jk / my);
            return true;
        };
        for (unsigned c = unsigned(T.tid); c < n; c += 2u * unsigned(T.nth)) {
            int i0, j0, k0, i1, j1, k1;
            const bool on0 = locate(c, i0, j0, k0), on1 = locate(c + unsigned(T.nth), i1, j1, k1);
            const FastIn<ABEC> X = interior_gather<ABEC>(i0, j0, k0, on0, P, R, A, BX, BY, BZ);
            const FastIn<ABEC> Z = interior_gather<ABEC>(i1, j1, k1, on1, P, R, A, BX, BY, BZ);
            finish(X); finish(Z);
        }
    }
    {   // surface
        const SurfaceMap M(B.nx, B.ny, B.nz);
        const unsigned n = M.size();
        for (unsigned t = unsigned(T.tid); t < n; t += unsigned(T.nth)) {
            int i, j, k;
            M.locate(t, i, j, k);
            finish(fast_gather<ABEC>(i, j, k, true, B, F, P, R, A, BX, BY, BZ));
        }
    }
}

__device__ __forceinline__ bool fast_level (const b200mg_leg_level& L)
{
    return (L.vb.hi[0] - L.vb.lo[0] >= 3) && (L.vb.hi[1] - L.vb.lo[1] >= 3) && (L.vb.hi[2] - L.vb.lo[2] >= 3);
}

// MLCellLinOpT::smooth (AMReX_MLCellLinOp.H:1206-1217): per colour, boundary values (inline) + sweep
template <bool ABEC, bool WIDE>
__device__ __forceinline__ void smooth (const b200mg_leg_level& L, const FaceBC& F, double alpha, const Team& T, Stamps& st, int id, GridBar& bar)
{
    const bool fast = fast_level(L);
    for (int redblack = 0; redblack < 2; ++redblack) {
        if (fast) { sweep_fast<ABEC>(L, F, alpha, redblack, T); } else { sweep<ABEC>(L, F, alpha, redblack, T); }
        team_sync<WIDE>(bar);
        st.mark(id + redblack);
    }
}

// rescor = res - L(cor) with homogeneous BCs (MLCellLinOpT::correctionResidual, AMReX_MLCellLinOp.H:1248-1270); no barrier inside
template <bool ABEC>
__device__ __noinline__ void residual_generic (const b200mg_leg_level& L, const FaceBC& F, double alpha, const Team& T)
{
    const auto x = view(L.cor); const auto b = view(L.res); const auto y = view(L.rescor);
    const long long js = x.js, ks = x.ks;
    const b200mg_box vb = L.vb;
    View<double> a = x, bx = x, by = x, bz = x;
    if constexpr (ABEC) { a = view(L.a); bx = view(L.bx); by = view(L.by); bz = view(L.bz); }
    for_box(vb, 0, T, [&] (int i, int j, int k) {
        const Nbrs q = load_nbrs(x.ptr(i, j, k), i, j, k, vb, F, js, ks);
        double v;
        if constexpr (ABEC) {
            v = adotx_abec_cell(q.c, q.xm, q.xp, q.ym, q.yp, q.zm, q.zp, a(i, j, k), bx(i, j, k), bx(i + 1, j, k),
                                by(i, j, k), by(i, j + 1, k), bz(i, j, k), bz(i, j, k + 1), alpha, L.adh[0], L.adh[1], L.adh[2]);
        } else {
            v = adotx_poisson_cell(q.c, q.xm, q.xp, q.ym, q.yp, q.zm, q.zp, L.adh[0], L.adh[1], L.adh[2]);
        }
        y(i, j, k) = b(i, j, k) + (-1.0) * v;      // Xpay(y,-1,b)
    });
}

template <bool ABEC>
__device__ __forceinline__ void residual (const b200mg_leg_level& L, const FaceBC& F, double alpha, const Team& T)
{
    if (fast_level(L)) { residual_fast<ABEC>(L, F, alpha, T); } else { residual_generic<ABEC>(L, F, alpha, T); }
}

// res(coarse) = average of rescor(fine) (amrex_avgdown, AMReX_MultiFabUtil_3D_C.H:381-394); also cor(coarse) = 0
__device__ __forceinline__ void restrict_and_zero (const b200mg_leg_level& F, const b200mg_leg_level& C, const Team& T)
{
    const auto fine = view(F.rescor); const auto crse = view(C.res);
    for_box(C.vb, 0, T, [&] (int i, int j, int k) {
        const double* p = fine.ptr(2 * i, 2 * j, 2 * k);
        double c = 0.0;
        c += p[0]; c += p[1]; c += p[fine.js]; c += p[fine.js + 1];
        c += p[fine.ks]; c += p[fine.ks + 1]; c += p[fine.ks + fine.js]; c += p[fine.ks + fine.js + 1];
        crse(i, j, k) = 0.125 * c;
    });
    zero_field(C, T);
}

// cor(fine) += cor(coarse) (AMReX_MLCellLinOp.H:973-976)
__device__ __forceinline__ void prolong_add (const b200mg_leg_level& F, const b200mg_leg_level& C, const Team& T)
{
    const auto fine = view(F.cor); const auto crse = view(C.cor);
    for_box(F.vb, 0, T, [&] (int i, int j, int k) { fine(i, j, k) += crse(i >> 1, j >> 1, k >> 1); });
}

__device__ __forceinline__ void load_bc (BoxBC& bc, const b200mg_leg_level& L)
{
    bc.nfaces = L.nfaces;
    for (int n = 0; n < 6; ++n) { bc.faces[n] = L.faces[n]; bc.mask[n] = L.m[n]; }
    for (int d = 0; d < 3; ++d) { bc.periodic[d] = L.periodic[d]; }
}

template <bool ABEC>
__global__ void __launch_bounds__(kLegThreads, 1)
k_coarse_leg (const b200mg_leg_args* __restrict__ gA, double* __restrict__ out, unsigned* __restrict__ bar_counter, int use_cg)
{
    __shared__ GridBar bar;
    if (threadIdx.x == 0) { bar.counter = bar_counter; bar.target = 0u; bar.nctas = gridDim.x; bar.use_cg = use_cg; }
    __shared__ b200mg_leg_args S;
    __shared__ BoxBC sbc[B200MG_LEG_MAX_LEVELS];
    __shared__ BottomArgs B;
    __shared__ double sh[kLegThreads / 32 + 1];
    const int tid = int(threadIdx.x);
    {
        static_assert(sizeof(b200mg_leg_args) % 8 == 0, "leg arguments are copied in 8-byte words");
        const unsigned long long* src = reinterpret_cast<const unsigned long long*>(gA);
        unsigned long long* dst = reinterpret_cast<unsigned long long*>(&S);
        for (int w = tid; w < int(sizeof(b200mg_leg_args) / 8); w += kLegThreads) { dst[w] = src[w]; }
    }
    __syncthreads();
    const int nl = S.nlev;
    if (tid < nl) { load_bc(sbc[tid], S.lev[tid]); }
    __syncthreads();
    for (int l = 0; l < nl; ++l) { bc_prepare(sbc[l], S.maxorder, S.lev[l].dxi, tid); }
    __syncthreads();
    __shared__ FaceBC fbc[B200MG_LEG_MAX_LEVELS];
    if (tid < nl) { make_face_bc(fbc[tid], sbc[tid]); }

    __shared__ int wide[B200MG_LEG_MAX_LEVELS];
    if (tid < nl) {
        const b200mg_box& b = S.lev[tid].vb;
        const long long cells = (long long)(b.hi[0] - b.lo[0] + 1) * (b.hi[1] - b.lo[1] + 1) * (b.hi[2] - b.lo[2] + 1);
        wide[tid] = (gridDim.x > 1 && cells > (long long)S.narrow_cells) ? 1 : 0;
    }
    __syncthreads();
    const int cta = int(blockIdx.x);
    const Team TW{cta * kLegThreads + tid, int(gridDim.x) * kLegThreads};
    const Team T0{tid, kLegThreads};
    const double alpha = S.alpha;
    Stamps st{(cta == 0 && tid == 0) ? S.stamps : nullptr, 1};
    st.mark(0);

    // ---- down: cor = 0, nu1 smooths, residual, restriction (mgVcycle, AMReX_MLMG.H:1318-1345)
    if (wide[0]) { zero_field(S.lev[0], TW); grid_barrier(bar); }
    else if (cta == 0) { zero_field(S.lev[0], T0); __syncthreads(); }
    for (int l = 0; l < nl - 1; ++l) {
        const b200mg_leg_level& L = S.lev[l];
        if (wide[l]) {
            for (int i = 0; i < S.nu1; ++i) { smooth<ABEC, true>(L, fbc[l], alpha, TW, st, l * 16, bar); }
            residual<ABEC>(L, fbc[l], alpha, TW);
            grid_barrier(bar);
            st.mark(l * 16 + 2);
            restrict_and_zero(L, S.lev[l + 1], TW);
            grid_barrier(bar);
            st.mark(l * 16 + 3);
        } else if (cta == 0) {
            for (int i = 0; i < S.nu1; ++i) { smooth<ABEC, false>(L, fbc[l], alpha, T0, st, l * 16, bar); }
            residual<ABEC>(L, fbc[l], alpha, T0);
            __syncthreads();
            st.mark(l * 16 + 2);
            restrict_and_zero(L, S.lev[l + 1], T0);
            __syncthreads();
            st.mark(l * 16 + 3);
        }
    }

    // ---- bottom (bottomSolve, AMReX_MLMG.H:1460-1576): BiCGStab by CTA 0 alone, block barriers only
    const b200mg_leg_level& LB = S.lev[nl - 1];
    if (S.bottom_mode == 1 && wide[nl - 1]) {                       // BottomSolver::smoother on a big bottom level
        for (int i = 0; i < S.nuf; ++i) { smooth<ABEC, true>(LB, fbc[nl - 1], alpha, TW, st, (nl - 1) * 16 + 8, bar); }
    } else if (cta == 0) {
        int ret = 0, iter = 0;
        if (S.bottom_mode == 1) {
            for (int i = 0; i < S.nuf; ++i) { smooth<ABEC, false>(LB, fbc[nl - 1], alpha, T0, st, (nl - 1) * 16 + 8, bar); }
        } else {
            if (tid == 0) {
                B.vb = LB.vb;
                B.sol = LB.cor; B.rhs = S.singular ? S.bb : LB.res; B.r = S.r; B.p = S.p; B.v = S.v; B.t = S.t; B.rh = S.rh;
                B.a = LB.a; B.bx = LB.bx; B.by = LB.by; B.bz = LB.bz;
                B.abec = ABEC ? 1 : 0;
                B.alpha = alpha; B.dhx = LB.adh[0]; B.dhy = LB.adh[1]; B.dhz = LB.adh[2];
                B.maxorder = S.maxorder;
                for (int d = 0; d < 3; ++d) { B.dxi[d] = LB.dxi[d]; }
                B.eps_rel = S.eps_rel; B.eps_abs = S.eps_abs; B.maxiter = S.maxiter;
                B.out = nullptr;
            }
            __syncthreads();
            const BottomCtx C = make_bottom_ctx(LB.vb, sh);
            if (S.singular) {
                // makeSolvable on a copy of the bottom right-hand side (AMReX_MLMG.H:1489-1499, AMReX_MLCellLinOp.H:2008-2060)
                const auto b = view(LB.res); const auto bb = view(S.bb);
                double acc = 0.0;
                bottom_for_cells(LB.vb, C, [&] (int i, int j, int k) { const double v = b(i, j, k); bb(i, j, k) = v; acc += v; });
                const double off = cta_sum(acc, C) * S.volinv;
                const double moff = -off;
                bottom_for_cells(LB.vb, C, [&] (int i, int j, int k) { bb(i, j, k) += moff; });
                __syncthreads();
            }
            bottom_bicgstab(B, sbc[nl - 1], C, ret, iter);
            st.mark((nl - 1) * 16 + 7);
            if (ret != 0 && ret != 9) {                             // the solve failed: start the smooths from zero
                zero_field(LB, T0);
                __syncthreads();
            }
            const int n = (ret == 0) ? S.nub : S.nuf;
            for (int i = 0; i < n; ++i) { smooth<ABEC, false>(LB, fbc[nl - 1], alpha, T0, st, (nl - 1) * 16 + 8, bar); }
        }
        if (tid == 0 && out != nullptr) { out[0] = double(ret); out[1] = double(iter); }
    }

    // ---- up: prolongation-add, nu2 smooths (AMReX_MLMG.H:1392-1413)
    bool joined = false;                                            // the grid has met again after CTA 0's solo part
    for (int l = nl - 2; l >= 0; --l) {
        const b200mg_leg_level& L = S.lev[l];
        if (wide[l]) {
            if (!joined) { __threadfence(); grid_barrier(bar); joined = true; }
            prolong_add(L, S.lev[l + 1], TW);
            grid_barrier(bar);
            st.mark(l * 16 + 4);
            for (int i = 0; i < S.nu2; ++i) { smooth<ABEC, true>(L, fbc[l], alpha, TW, st, l * 16 + 5, bar); }
        } else if (cta == 0) {
            prolong_add(L, S.lev[l + 1], T0);
            __syncthreads();
            st.mark(l * 16 + 4);
            for (int i = 0; i < S.nu2; ++i) { smooth<ABEC, false>(L, fbc[l], alpha, T0, st, l * 16 + 5, bar); }
        }
    }
}

int g_max_ctas[2] = {0, 0};      // co-resident CTAs of the cooperative grid, per kernel instance

} // namespace

extern "C" {

int b200mg_coarse_leg (int abec, const b200mg_leg_args* d_args, double* d_out, int ctas, cudaStream_t s)
{
    static unsigned* d_counter = nullptr;                 // the grid barrier's counter, zeroed ahead of every launch
    static const int use_cg = std::getenv("B200MG_LEG_CG_SYNC") != nullptr;
    if (!d_counter) {
        const cudaError_t e = cudaMalloc(&d_counter, 64);
        if (e != cudaSuccess) { return int(e); }
    }
    auto kern = abec ? k_coarse_leg<true> : k_coarse_leg<false>;
    int& maxc = g_max_ctas[abec ? 1 : 0];
    if (maxc == 0) {
        int dev = 0, sms = 0, per_sm = 0, coop = 0;
        cudaError_t e = cudaGetDevice(&dev);
        if (e == cudaSuccess) { e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev); }
        if (e == cudaSuccess) { e = cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev); }
        if (e == cudaSuccess) { e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kLegThreads, 0); }
        if (e != cudaSuccess) { return int(e); }
        if (!coop || per_sm < 1) { return int(cudaErrorCooperativeLaunchTooLarge); }
        maxc = sms;                                                 // one CTA per SM: more only lengthens the grid barrier
    }
    if (ctas < 1) { return int(cudaErrorInvalidValue); }
    if (ctas > maxc) { ctas = maxc; }
    {
        const cudaError_t e = cudaMemsetAsync(d_counter, 0, sizeof(unsigned), s);
        if (e != cudaSuccess) { return int(e); }
    }
    int cgflag = use_cg;
    void* args[4] = {(void*)&d_args, (void*)&d_out, (void*)&d_counter, (void*)&cgflag};
    const cudaError_t e = cudaLaunchCooperativeKernel((const void*)kern, dim3(unsigned(ctas), 1, 1), dim3(kLegThreads, 1, 1), args, 0, s);
    if (e != cudaSuccess) { return int(e); }
    return last_error();
}

} // extern "C"
