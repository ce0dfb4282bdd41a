// The coarse leg of a V-cycle in ONE kernel (reference control flow: MLMGT::mgVcycle, AMReX_MLMG.H:1308-1415, and
// MLMGT::bottomSolve / actualBottomSolve, :1460-1576).
//
// From the first MG level that is a single box covering the whole domain (<= 32^3 cells, the product of agglomeration)
// down to the bottom and back up, the launch-per-operation schedule issues ~35 kernels of 2-10 us per level and V-cycle
// (boundary fill, two colour sweeps per smooth, residual, restriction, prolongation), every one bound by launch and
// drain latency, not by work; on 8 GPUs these levels cost as much as on one.  Here one thread-block CLUSTER runs the whole
// leg: pre-smooths, residual and restriction level by level, the BiCGStab bottom solve (bottom_solve.cuh, by CTA 0),
// then prolongation and post-smooths back up.  Phases are separated by the hardware cluster barrier (release / acquire,
// ~0.2 us), fields stay in L2.  Per-cell arithmetic is the shared stencil_math.cuh code and the sequence of operations
// is the host schedule's, so the leg leaves the bits the launch-per-operation path leaves.
#include "bottom_solve.cuh"

using namespace b200mg;

namespace {

constexpr int kLegThreads = kBottomThreads;     // CTA size (the bottom solve's reduction order is tied to it)

struct Team { int tid, nth; };

__device__ __forceinline__ unsigned cluster_ctarank () { unsigned r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ unsigned cluster_nctarank () { unsigned r; asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_barrier ()
{
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// CL: the team is the whole cluster, else the calling CTA
template <bool CL> __device__ __forceinline__ void team_sync () { if constexpr (CL) { cluster_barrier(); } else { __syncthreads(); } }

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
    for_box(vb, 0, T, [&] (int i, int j, int k) {
        if ((i + j + k + redblack) & 1) { return; }
        double* pc = phi.ptr(i, j, k);
        FaceCoefs cf;
        if (on_surface(i, j, k, vb)) { cf = face_coefs(i, j, k, vb, L.f, L.m); }
        else {
#pragma unroll
            for (int n = 0; n < 6; ++n) { cf.c[n] = 0.0; }
        }
        if constexpr (ABEC) {
            *pc = gsrb_abec_cell(*pc, pc[-1], pc[1], pc[-js], pc[js], pc[-ks], pc[ks], rhs(i, j, k), a(i, j, k),
                                 bx(i, j, k), bx(i + 1, j, k), by(i, j, k), by(i, j + 1, k), bz(i, j, k), bz(i, j, k + 1),
                                 cf.c[0], cf.c[1], cf.c[2], cf.c[3], cf.c[4], cf.c[5], alpha, L.dh[0], L.dh[1], L.dh[2]);
        } else {
            *pc = gsrb_poisson_cell(*pc, pc[-1], pc[1], pc[-js], pc[js], pc[-ks], pc[ks], rhs(i, j, k),
                                    cf.c[0], cf.c[1], cf.c[2], cf.c[3], cf.c[4], cf.c[5], L.dh[0], L.dh[1], L.dh[2]);
        }
    });
}

// MLCellLinOpT::smooth (AMReX_MLCellLinOp.H:1206-1217): boundary fill + sweep, per colour
template <bool ABEC, bool CL>
__device__ __forceinline__ void smooth (const b200mg_leg_level& L, const BoxBC& bc, double alpha, const Team& T)
{
    const auto phi = view(L.cor);
    for (int redblack = 0; redblack < 2; ++redblack) {
        bc_fill(L.vb, bc, phi, T.tid, T.nth);
        team_sync<CL>();
        sweep<ABEC>(L, alpha, redblack, T);
        team_sync<CL>();
    }
}

// rescor = res - L(cor) with homogeneous BCs (MLCellLinOpT::correctionResidual, AMReX_MLCellLinOp.H:1248-1270)
template <bool ABEC, bool CL>
__device__ __forceinline__ void residual (const b200mg_leg_level& L, const BoxBC& bc, double alpha, const Team& T)
{
    const auto x = view(L.cor); const auto b = view(L.res); const auto y = view(L.rescor);
    bc_fill(L.vb, bc, x, T.tid, T.nth);
    team_sync<CL>();
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
    team_sync<CL>();
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

    const int rank = int(cluster_ctarank());
    const Team T{rank * kLegThreads + tid, int(cluster_nctarank()) * kLegThreads};
    const Team T0{tid, kLegThreads};
    const double alpha = S.alpha;

    // ---- down: cor = 0, nu1 smooths, residual, restriction (mgVcycle, AMReX_MLMG.H:1318-1345)
    zero_field(S.lev[0], T);
    cluster_barrier();
    for (int l = 0; l < nl - 1; ++l) {
        const b200mg_leg_level& L = S.lev[l];
        for (int i = 0; i < S.nu1; ++i) { smooth<ABEC, true>(L, sbc[l], alpha, T); }
        residual<ABEC, true>(L, sbc[l], alpha, T);
        restrict_and_zero(L, S.lev[l + 1], T);
        cluster_barrier();
    }

    // ---- bottom (bottomSolve, AMReX_MLMG.H:1460-1576): CTA 0 alone, block barriers only
    if (rank == 0) {
        const b200mg_leg_level& L = S.lev[nl - 1];
        int ret = 0, iter = 0;
        if (S.bottom_mode == 1) {                                   // BottomSolver::smoother
            for (int i = 0; i < S.nuf; ++i) { smooth<ABEC, false>(L, sbc[nl - 1], alpha, T0); }
        } else {
            if (tid == 0) {
                B.vb = L.vb;
                B.sol = L.cor; B.rhs = S.singular ? S.bb : L.res; B.r = S.r; B.p = S.p; B.v = S.v; B.t = S.t; B.rh = S.rh;
                B.a = L.a; B.bx = L.bx; B.by = L.by; B.bz = L.bz;
                B.abec = ABEC ? 1 : 0;
                B.alpha = alpha; B.dhx = L.adh[0]; B.dhy = L.adh[1]; B.dhz = L.adh[2];
                B.maxorder = S.maxorder;
                for (int d = 0; d < 3; ++d) { B.dxi[d] = L.dxi[d]; }
                B.eps_rel = S.eps_rel; B.eps_abs = S.eps_abs; B.maxiter = S.maxiter;
                B.out = nullptr;
            }
            __syncthreads();
            const BottomCtx C = make_bottom_ctx(L.vb, sh);
            if (S.singular) {
                // makeSolvable on a copy of the bottom right-hand side (AMReX_MLMG.H:1489-1499, AMReX_MLCellLinOp.H:2008-2060)
                const auto b = view(L.res); const auto bb = view(S.bb);
                double acc = 0.0;
                bottom_for_cells(L.vb, C, [&] (int i, int j, int k) { const double v = b(i, j, k); bb(i, j, k) = v; acc += v; });
                const double off = cta_sum(acc, C) * S.volinv;
                const double moff = -off;
                bottom_for_cells(L.vb, C, [&] (int i, int j, int k) { bb(i, j, k) += moff; });
                __syncthreads();
            }
            bottom_bicgstab(B, sbc[nl - 1], C, ret, iter);
            if (ret != 0 && ret != 9) {                             // the solve failed: start the smooths from zero
                zero_field(L, T0);
                __syncthreads();
            }
            const int n = (ret == 0) ? S.nub : S.nuf;
            for (int i = 0; i < n; ++i) { smooth<ABEC, false>(L, sbc[nl - 1], alpha, T0); }
        }
        if (tid == 0 && out != nullptr) { out[0] = double(ret); out[1] = double(iter); }
    }
    cluster_barrier();

    // ---- up: prolongation-add, nu2 smooths (AMReX_MLMG.H:1392-1413)
    for (int l = nl - 2; l >= 0; --l) {
        const b200mg_leg_level& L = S.lev[l];
        prolong_add(L, S.lev[l + 1], T);
        cluster_barrier();
        for (int i = 0; i < S.nu2; ++i) { smooth<ABEC, true>(L, sbc[l], alpha, T); }
    }
}

bool g_attr_set[2] = {false, false};

} // namespace

extern "C" {

int b200mg_coarse_leg (int abec, const b200mg_leg_args* d_args, double* d_out, int cluster_ctas, cudaStream_t s)
{
    if (cluster_ctas < 1 || cluster_ctas > 16) { return int(cudaErrorInvalidValue); }
    auto kern = abec ? k_coarse_leg<true> : k_coarse_leg<false>;
    if (cluster_ctas > 8 && !g_attr_set[abec ? 1 : 0]) {
        const cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
        if (e != cudaSuccess) { return int(e); }
        g_attr_set[abec ? 1 : 0] = true;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(unsigned(cluster_ctas), 1, 1);
    cfg.blockDim = dim3(kLegThreads, 1, 1);
    cfg.dynamicSmemBytes = 0;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = unsigned(cluster_ctas); attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    const cudaError_t e = cudaLaunchKernelEx(&cfg, kern, d_args, d_out);
    if (e != cudaSuccess) { return int(e); }
    return last_error();
}

} // extern "C"
