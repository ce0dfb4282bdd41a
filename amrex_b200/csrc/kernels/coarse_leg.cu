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

namespace cg = cooperative_groups;
using namespace b200mg;

namespace {

constexpr int kLegThreads = kBottomThreads;     // CTA size (the bottom solve's reduction order is tied to it)

struct Team { int tid, nth; };

// WIDE: the team is the whole grid, else the calling CTA
template <bool WIDE> __device__ __forceinline__ void team_sync ()
{
    if constexpr (WIDE) { cg::this_grid().sync(); } else { __syncthreads(); }
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

// one colour of MLCellLinOpT::smooth's Fsmooth (abec_gsrb / mlpoisson_gsrb), cells with (i+j+k+redblack) even
template <bool ABEC>
__device__ __forceinline__ void sweep (const b200mg_leg_level& L, double alpha, int redblack, const Team& T)
{
    const auto phi = view(L.cor); const auto rhs = view(L.res);
    const int js = int(phi.js), ks = int(phi.ks);
    const b200mg_box vb = L.vb;
    View<double> a = phi, bx = phi, by = phi, bz = phi;
    if constexpr (ABEC) { a = view(L.a); bx = view(L.bx); by = view(L.by); bz = view(L.bz); }
    // the cells of this colour only: row (j,k) holds ceil(nx/2) candidates i = i0 + 2*ii, so a warp covers 64 consecutive cells
    const int nx = vb.hi[0] - vb.lo[0] + 1, ny = vb.hi[1] - vb.lo[1] + 1, nz = vb.hi[2] - vb.lo[2] + 1;
    const int nxh = (nx + 1) >> 1;
    const int n = nxh * ny * nz;
    for (int c = T.tid; c < n; c += T.nth) {
        const int ii = c % nxh, jk = c / nxh;
        const int j = vb.lo[1] + jk % ny, k = vb.lo[2] + jk / ny;
        const int i = vb.lo[0] + ((vb.lo[0] + j + k + redblack) & 1) + 2 * ii;
        if (i > vb.hi[0]) { continue; }
        double* pc = phi.ptr(i, j, k);
        FaceCoefs cf;
        if (on_surface(i, j, k, vb)) { cf = face_coefs(i, j, k, vb, L.f, L.m); }
        else {
#pragma unroll
            for (int q = 0; q < 6; ++q) { cf.c[q] = 0.0; }
        }
        if constexpr (ABEC) {
            *pc = gsrb_abec_cell(*pc, pc[-1], pc[1], pc[-js], pc[js], pc[-ks], pc[ks], rhs(i, j, k), a(i, j, k),
                                 bx(i, j, k), bx(i + 1, j, k), by(i, j, k), by(i, j + 1, k), bz(i, j, k), bz(i, j, k + 1),
                                 cf.c[0], cf.c[1], cf.c[2], cf.c[3], cf.c[4], cf.c[5], alpha, L.dh[0], L.dh[1], L.dh[2]);
        } else {
            *pc = gsrb_poisson_cell(*pc, pc[-1], pc[1], pc[-js], pc[js], pc[-ks], pc[ks], rhs(i, j, k),
                                    cf.c[0], cf.c[1], cf.c[2], cf.c[3], cf.c[4], cf.c[5], L.dh[0], L.dh[1], L.dh[2]);
        }
    }
}

// MLCellLinOpT::smooth (AMReX_MLCellLinOp.H:1206-1217): boundary fill + sweep, per colour
template <bool ABEC, bool WIDE>
__device__ __forceinline__ void smooth (const b200mg_leg_level& L, const BoxBC& bc, double alpha, const Team& T)
{
    const auto phi = view(L.cor);
    for (int redblack = 0; redblack < 2; ++redblack) {
        bc_fill(L.vb, bc, phi, T.tid, T.nth);
        team_sync<WIDE>();
        sweep<ABEC>(L, alpha, redblack, T);
        team_sync<WIDE>();
    }
}

// rescor = res - L(cor) with homogeneous BCs (MLCellLinOpT::correctionResidual, AMReX_MLCellLinOp.H:1248-1270)
template <bool ABEC, bool WIDE>
__device__ __forceinline__ void residual (const b200mg_leg_level& L, const BoxBC& bc, double alpha, const Team& T)
{
    const auto x = view(L.cor); const auto b = view(L.res); const auto y = view(L.rescor);
    bc_fill(L.vb, bc, x, T.tid, T.nth);
    team_sync<WIDE>();
    const int js = int(x.js), ks = int(x.ks);
    View<double> a = x, bx = x, by = x, bz = x;
    if constexpr (ABEC) { a = view(L.a); bx = view(L.bx); by = view(L.by); bz = view(L.bz); }
    for_box(L.vb, 0, T, [&] (int i, int j, int k) {
        const double* xc = x.ptr(i, j, k);
        double v;
        if constexpr (ABEC) {
            v = adotx_abec_cell(*xc, xc[-1], xc[1], xc[-js], xc[js], xc[-ks], xc[ks], a(i, j, k), bx(i, j, k), bx(i + 1, j, k),
                                by(i, j, k), by(i, j + 1, k), bz(i, j, k), bz(i, j, k + 1), alpha, L.adh[0], L.adh[1], L.adh[2]);
        } else {
            v = adotx_poisson_cell(*xc, xc[-1], xc[1], xc[-js], xc[js], xc[-ks], xc[ks], L.adh[0], L.adh[1], L.adh[2]);
        }
        y(i, j, k) = b(i, j, k) + (-1.0) * v;      // Xpay(y,-1,b)
    });
    team_sync<WIDE>();
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
k_coarse_leg (const b200mg_leg_args* __restrict__ gA, double* __restrict__ out)
{
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

    // ---- down: cor = 0, nu1 smooths, residual, restriction (mgVcycle, AMReX_MLMG.H:1318-1345)
    if (wide[0]) { zero_field(S.lev[0], TW); cg::this_grid().sync(); }
    else if (cta == 0) { zero_field(S.lev[0], T0); __syncthreads(); }
    for (int l = 0; l < nl - 1; ++l) {
        const b200mg_leg_level& L = S.lev[l];
        if (wide[l]) {
            for (int i = 0; i < S.nu1; ++i) { smooth<ABEC, true>(L, sbc[l], alpha, TW); }
            residual<ABEC, true>(L, sbc[l], alpha, TW);
            restrict_and_zero(L, S.lev[l + 1], TW);
            cg::this_grid().sync();
        } else if (cta == 0) {
            for (int i = 0; i < S.nu1; ++i) { smooth<ABEC, false>(L, sbc[l], alpha, T0); }
            residual<ABEC, false>(L, sbc[l], alpha, T0);
            restrict_and_zero(L, S.lev[l + 1], T0);
            __syncthreads();
        }
    }

    // ---- bottom (bottomSolve, AMReX_MLMG.H:1460-1576): BiCGStab by CTA 0 alone, block barriers only
    const b200mg_leg_level& LB = S.lev[nl - 1];
    if (S.bottom_mode == 1 && wide[nl - 1]) {                       // BottomSolver::smoother on a big bottom level
        for (int i = 0; i < S.nuf; ++i) { smooth<ABEC, true>(LB, sbc[nl - 1], alpha, TW); }
    } else if (cta == 0) {
        int ret = 0, iter = 0;
        if (S.bottom_mode == 1) {
            for (int i = 0; i < S.nuf; ++i) { smooth<ABEC, false>(LB, sbc[nl - 1], alpha, T0); }
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
            if (ret != 0 && ret != 9) {                             // the solve failed: start the smooths from zero
                zero_field(LB, T0);
                __syncthreads();
            }
            const int n = (ret == 0) ? S.nub : S.nuf;
            for (int i = 0; i < n; ++i) { smooth<ABEC, false>(LB, sbc[nl - 1], alpha, T0); }
        }
        if (tid == 0 && out != nullptr) { out[0] = double(ret); out[1] = double(iter); }
    }

    // ---- up: prolongation-add, nu2 smooths (AMReX_MLMG.H:1392-1413)
    bool joined = false;                                            // the grid has met again after CTA 0's solo part
    for (int l = nl - 2; l >= 0; --l) {
        const b200mg_leg_level& L = S.lev[l];
        if (wide[l]) {
            if (!joined) { __threadfence(); cg::this_grid().sync(); joined = true; }
            prolong_add(L, S.lev[l + 1], TW);
            cg::this_grid().sync();
            for (int i = 0; i < S.nu2; ++i) { smooth<ABEC, true>(L, sbc[l], alpha, TW); }
        } else if (cta == 0) {
            prolong_add(L, S.lev[l + 1], T0);
            __syncthreads();
            for (int i = 0; i < S.nu2; ++i) { smooth<ABEC, false>(L, sbc[l], alpha, T0); }
        }
    }
}

int g_max_ctas[2] = {0, 0};      // co-resident CTAs of the cooperative grid, per kernel instance

} // namespace

extern "C" {

int b200mg_coarse_leg (int abec, const b200mg_leg_args* d_args, double* d_out, int ctas, cudaStream_t s)
{
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
    void* args[2] = {(void*)&d_args, (void*)&d_out};
    const cudaError_t e = cudaLaunchCooperativeKernel((const void*)kern, dim3(unsigned(ctas), 1, 1), dim3(kLegThreads, 1, 1), args, 0, s);
    if (e != cudaSuccess) { return int(e); }
    return last_error();
}

} // extern "C"
