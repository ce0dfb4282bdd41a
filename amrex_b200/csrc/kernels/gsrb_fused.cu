// Fused red+black Gauss-Seidel pass: ONE sweep over memory per smooth instead of two.
//
// Reference schedule (MLCellLinOpT::smooth, AMReX_MLCellLinOp.H:1206-1217; kernels abec_gsrb AMReX_MLABecLap_3D_K.H:210-264,
// mlpoisson_gsrb AMReX_MLPoisson_3D_K.H:155-196): halo refresh, red sweep, halo refresh, black sweep -- every array is
// streamed from HBM twice.  Here a CTA owns a tile (all x, tile_y rows, chunk_z planes) of one box and streams it in z:
// at step k it finishes the RED update of plane k+1 and then the BLACK update of plane k, which by then sees new red
// values on all six sides.  x/y neighbours come from three rotating shared-memory planes, z neighbours from registers.
// Red values on the one-cell ring around the tile (inside the box) are recomputed redundantly, so tiles do not
// communicate; the pass is out of place (phi_in -> phi_out) so no tile ever reads a value another tile has overwritten.
// Black cells ON the box surface need red ghost values of neighbouring boxes: they are copied through unchanged and
// finished by b200mg_gsrb_shell_* after the second halo refresh.  Per-cell arithmetic is the shared code of
// stencil_math.cuh, and every cell sees exactly the operands the two-sweep schedule gives it => identical bits.
//
// Thread map: blockDim = (TX, tile_y + 4).  Thread (tx, ty) owns the cell pair (lo_x + 2 tx, +1) of row j0 - 2 + ty: one
// red and one black cell in every plane (no colour divergence).  Rows 0 / last only feed y neighbours, rows 1 / last-1 are
// the redundant red ring.  HBM traffic per cell: phi 8 + rhs 8 + a 8 + b 24 + phi_out 8 = 56 B (ABecLap), 24 B (Poisson).
#include "common.cuh"
#include "stencil_math.cuh"

using namespace b200mg;

namespace {

struct FusedArgs {
    const b200mg_fab *pin, *pout, *rhs, *a, *bx, *by, *bz, *f;
    const b200mg_ifab* m;
    double alpha, dhx, dhy, dhz;
    int tile_y, chunk_z;
    int pf;              // L2 prefetch distance in planes (0: off)
};

int g_prefetch_planes = 0;     // measured on B200 (profiles/r01_s8_tune_smoother.txt): the fused pass is not DRAM-latency bound, prefetch costs 3-4 %

// One instruction pulls `bytes` (multiple of 16, 16-byte aligned address) of global memory into L2 without occupying a
// register or a shared-memory slot: the DRAM latency of the plane that is D steps ahead is paid here, the real loads
// later only see L2 latency.
__device__ __forceinline__ void l2_prefetch (const void* p, unsigned bytes)
{
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" :: "l"(p), "r"(bytes) : "memory");
}

__device__ __forceinline__ double sel (const double2& v, int c) { return c ? v.y : v.x; }
__device__ __forceinline__ void put (double2& v, int c, double x) { if (c) { v.y = x; } else { v.x = x; } }

// Per-thread cursors into the coefficient arrays: 32-bit element offsets of (i0, j, red plane) from each fab's base,
// advanced by one plane per step; bases and strides are CTA-uniform.  Loaded ONCE from the descriptor tables (no
// per-access descriptor traffic or 64-bit index arithmetic in the loop).  The black plane is one plane stride behind.
template <bool ABEC>
struct Cursors {
    const double *rhs, *a, *bx, *by, *bz;            // fab bases (uniform)
    int rhs_ks, a_ks, bx_ks, by_ks, bz_ks, by_js;    // strides (uniform)
    int o_rhs, o_a, o_bx, o_by, o_bz;                // per-thread offsets at the red plane
    __device__ __forceinline__ void init (const FusedArgs& A, int box, int i0, int j, int k)
    {
        const auto r = view(A.rhs[box]);
        rhs = r.p; rhs_ks = int(r.ks); o_rhs = int(r.ptr(i0, j, k) - r.p);
        if constexpr (ABEC) {
            const auto va = view(A.a[box]); const auto vx = view(A.bx[box]); const auto vy = view(A.by[box]); const auto vz = view(A.bz[box]);
            a = va.p; bx = vx.p; by = vy.p; bz = vz.p;
            a_ks = int(va.ks); bx_ks = int(vx.ks); by_ks = int(vy.ks); bz_ks = int(vz.ks); by_js = int(vy.js);
            o_a = int(va.ptr(i0, j, k) - va.p); o_bx = int(vx.ptr(i0, j, k) - vx.p);
            o_by = int(vy.ptr(i0, j, k) - vy.p); o_bz = int(vz.ptr(i0, j, k) - vz.p);
        }
    }
    __device__ __forceinline__ void advance ()
    {
        o_rhs += rhs_ks;
        if constexpr (ABEC) { o_a += a_ks; o_bx += bx_ks; o_by += by_ks; o_bz += bz_ks; }
    }
};

// update of cell c of the pair at the cursors' current plane
template <bool ABEC>
__device__ __forceinline__ double
update_cell (const Cursors<ABEC>& C, int back, int c, int i, int j, int k, int box, const b200mg_box& vb, const FusedArgs& A, bool surface,
             double p, double xm, double xp, double ym, double yp, double zm, double zp)
{
    const double rhs = __ldg(C.rhs + (C.o_rhs - back * C.rhs_ks + c));
    if constexpr (ABEC) {
        const double a = __ldg(C.a + (C.o_a - back * C.a_ks + c));
        const double* pbx = C.bx + (C.o_bx - back * C.bx_ks + c);
        const double* pby = C.by + (C.o_by - back * C.by_ks + c);
        const double* pbz = C.bz + (C.o_bz - back * C.bz_ks + c);
        const double bxm = __ldg(pbx), bxp = __ldg(pbx + 1);
        const double bym = __ldg(pby), byp = __ldg(pby + C.by_js);
        const double bzm = __ldg(pbz), bzp = __ldg(pbz + C.bz_ks);
        if (surface) {
            const FaceCoefs cf = face_coefs(i, j, k, vb, A.f + 6 * box, A.m + 6 * box);
            return gsrb_abec_cell(p, xm, xp, ym, yp, zm, zp, rhs, a, bxm, bxp, bym, byp, bzm, bzp,
                                  cf.c[0], cf.c[1], cf.c[2], cf.c[3], cf.c[4], cf.c[5], A.alpha, A.dhx, A.dhy, A.dhz);
        }
        return gsrb_abec_cell_interior(p, xm, xp, ym, yp, zm, zp, rhs, a, bxm, bxp, bym, byp, bzm, bzp, A.alpha, A.dhx, A.dhy, A.dhz);
    } else {
        FaceCoefs cf;
        if (surface) { cf = face_coefs(i, j, k, vb, A.f + 6 * box, A.m + 6 * box); }
        else {
#pragma unroll
            for (int n = 0; n < 6; ++n) { cf.c[n] = 0.0; }
        }
        return gsrb_poisson_cell(p, xm, xp, ym, yp, zm, zp, rhs, cf.c[0], cf.c[1], cf.c[2], cf.c[3], cf.c[4], cf.c[5], A.dhx, A.dhy, A.dhz);
    }
}

template <bool ABEC, int MAXT>
__global__ void __launch_bounds__(MAXT)
k_gsrb2 (const b200mg_tile* __restrict__ tiles, const b200mg_box* __restrict__ vbox, FusedArgs A)
{
    extern __shared__ double sm[];
    const b200mg_tile t = tiles[blockIdx.x];
    const b200mg_box vb = vbox[t.box];
    const int tx = threadIdx.x, ty = threadIdx.y;
    const int SX = 2 * int(blockDim.x) + 4;             // cell i lives at column i - lo_x + 2 (even => 16 B aligned pairs)
    const int psz = int(blockDim.y) * SX;
    double* sA = sm; double* sB = sm + psz; double* sC = sm + 2 * psz;

    const int i0 = vb.lo[0] + 2 * tx;
    const bool xact = (i0 < vb.hi[0]);                  // nx is even: both cells of the pair are valid
    const int j = t.j0 - 2 + ty;
    const int j1 = min(t.j0 + A.tile_y - 1, vb.hi[1]);
    const int k0 = t.k0, k1 = min(t.k0 + A.chunk_z - 1, vb.hi[2]);
    const bool row_load = xact && (j >= vb.lo[1] - 1) && (j <= min(j1 + 2, vb.hi[1] + 1));
    const bool row_red = xact && (j >= max(t.j0 - 1, vb.lo[1])) && (j <= min(j1 + 1, vb.hi[1]));
    const bool row_black = xact && (j >= t.j0) && (j <= j1);
    const bool gl = row_load && (tx == 0), gr = row_load && (i0 + 1 == vb.hi[0]);   // who carries the x ghost cells
    const bool jsurf = (j == vb.lo[1]) || (j == vb.hi[1]);
    const bool isurf0 = (i0 == vb.lo[0]), isurf1 = (i0 + 1 == vb.hi[0]);
    const int par0 = (i0 + j) & 1;

    const auto pin = view(A.pin[t.box]);
    const auto pout = view(A.pout[t.box]);
    const int pin_ks = int(pin.ks), pout_ks = int(pout.ks);
    const double* qin = pin.ptr(i0, j, k0 + 1);          // plane kk+3 of loop step kk (k0+1 at the first step)
    double* qout = pout.ptr(i0, j, k0);
    const int srow = ty * SX + 2 * tx + 2;

    // cursor at the red plane (kk+1); the black plane (kk) is one stride behind; the loop starts at kk = k0-2
    Cursors<ABEC> cr;
    cr.init(A, t.box, i0, j, k0 - 1);

    auto load_at = [&] (const double* p, int k, double2& v, double& vl, double& vr) {
        v = make_double2(0.0, 0.0); vl = 0.0; vr = 0.0;
        if (row_load && k >= vb.lo[2] - 1 && k <= vb.hi[2] + 1) {
            v = *reinterpret_cast<const double2*>(p);
            if (gl) { vl = p[-1]; }
            if (gr) { vr = p[2]; }
        }
    };
    auto store_plane = [&] (double* s, const double2& v, double vl, double vr) {
        if (row_load) { *reinterpret_cast<double2*>(s + srow) = v; }   // idle lanes must not touch the ghost column
        if (gl) { s[srow - 1] = vl; }
        if (gr) { s[srow + 2] = vr; }
    };

    double2 pm1 = make_double2(0.0, 0.0), pk, pp1, pp2;
    double gl0, gr0, gl1, gr1;
    load_at(qin - 3 * pin_ks, k0 - 2, pk, gl0, gr0);
    load_at(qin - 2 * pin_ks, k0 - 1, pp1, gl0, gr0);
    store_plane(sB, pp1, gl0, gr0);
    load_at(qin - pin_ks, k0, pp2, gl1, gr1);            // ghost x values of the plane that sits in registers travel with it
    __syncthreads();

    const int kr_lo = max(k0 - 1, vb.lo[2]), kr_hi = min(k1 + 1, vb.hi[2]);
    // L2 prefetch: lanes 0..5 of every row pull that row of one array each, A.pf planes ahead of the loads below
    const unsigned pf_bytes = unsigned(vb.hi[0] - vb.lo[0] + 1) * 8u;
    const bool pf_phi = (A.pf > 0) && (tx == 0) && (j >= vb.lo[1] - 1) && (j <= min(j1 + 2, vb.hi[1] + 1));
    const bool pf_coef = (A.pf > 0) && ((j >= max(t.j0 - 1, vb.lo[1])) && (j <= min(j1 + 1, vb.hi[1]))) && (tx >= 1) && (tx <= (ABEC ? 5 : 1));
    const double* pf_ptr = nullptr; int pf_ks = 0;
    if (pf_phi) { pf_ptr = pin.ptr(vb.lo[0], j, k0 + 1 + A.pf); pf_ks = pin_ks; }
    if (pf_coef) {
        const b200mg_fab* fab = (tx == 1) ? A.rhs : (tx == 2) ? A.a : (tx == 3) ? A.bx : (tx == 4) ? A.by : A.bz;
        const auto v = view(fab[t.box]);
        pf_ptr = v.ptr(vb.lo[0], j, k0 - 1 + A.pf); pf_ks = int(v.ks);
    }
    const int pf_klast = min(k1 + 1, vb.hi[2]);
    for (int kk = k0 - 2; kk <= k1; ++kk) {
        if (pf_phi || pf_coef) {
            const int kp = (pf_phi ? kk + 3 : kk + 1) + A.pf;      // plane the pointer addresses at this step
            if (kp <= pf_klast) { l2_prefetch(pf_ptr, pf_bytes); }
            pf_ptr += pf_ks;
        }
        // ---- phase 1: red update of plane kk+1, in place in sB and pp1
        const int kr = kk + 1;
        if (row_red && kr >= kr_lo && kr <= kr_hi) {
            const int c = (par0 + kr) & 1;              // which cell of the pair is red
            const double* s = sB + srow + c;
            const bool surf = jsurf || (c ? isurf1 : isurf0) || (kr == vb.lo[2]) || (kr == vb.hi[2]);
            const double v = update_cell<ABEC>(cr, 0, c, i0 + c, j, kr, t.box, vb, A, surf, sel(pp1, c), s[-1], s[1], s[-SX], s[SX],
                                               sel(pk, c), sel(pp2, c));
            put(pp1, c, v);
            sB[srow + c] = v;
        }
        // (no barrier here: phase 1 touches sB only -- red cells written, black cells read -- and phase 2 reads sA only)
        // ---- phase 2: black update of plane kk (reads sA: new red neighbours), write the finished plane
        if (row_black && kk >= k0) {
            const int c = 1 - ((par0 + kk) & 1);        // which cell of the pair is black
            const bool surf = jsurf || (c ? isurf1 : isurf0) || (kk == vb.lo[2]) || (kk == vb.hi[2]);
            double2 o = pk;
            if (!surf) {
                const double* s = sA + srow + c;
                put(o, c, update_cell<ABEC>(cr, 1, c, i0 + c, j, kk, t.box, vb, A, false, sel(pk, c), s[-1], s[1], s[-SX], s[SX],
                                            sel(pm1, c), sel(pp1, c)));
            }
            *reinterpret_cast<double2*>(qout) = o;
            qout += pout_ks;
        }
        store_plane(sC, pp2, gl1, gr1);                 // old plane kk+2 for the next step's red phase
        pm1 = pk; pk = pp1; pp1 = pp2;
        load_at(qin, kk + 3, pp2, gl1, gr1);
        qin += pin_ks;
        cr.advance();
        __syncthreads();
        double* tmp = sA; sA = sB; sB = sC; sC = tmp;
    }
}

template <bool ABEC>
int launch (int nblocks, const b200mg_tile* tiles, const b200mg_box* vbox, const FusedArgs& A, int tx, cudaStream_t s)
{
    if (nblocks <= 0) { return 0; }
    if (tx <= 0 || tx > 128 || (tx % 32) != 0 || A.tile_y < 1 || A.chunk_z < 1) { return int(cudaErrorInvalidValue); }
    const dim3 block(tx, A.tile_y + 4, 1);
    const int nthreads = int(block.x * block.y);
    if (nthreads > 1024) { return int(cudaErrorInvalidValue); }
    const size_t smem = size_t(3) * block.y * (2 * block.x + 4) * sizeof(double);
    if (nthreads <= 512) {
        auto kern = k_gsrb2<ABEC, 512>;
        if (smem > 48 * 1024) { cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)); }
        kern<<<nblocks, block, smem, s>>>(tiles, vbox, A);
    } else if (nthreads <= 768) {
        auto kern = k_gsrb2<ABEC, 768>;
        if (smem > 48 * 1024) { cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)); }
        kern<<<nblocks, block, smem, s>>>(tiles, vbox, A);
    } else {
        auto kern = k_gsrb2<ABEC, 1024>;
        if (smem > 48 * 1024) { cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)); }
        kern<<<nblocks, block, smem, s>>>(tiles, vbox, A);
    }
    return last_error();
}

} // namespace

extern "C" {

int b200mg_gsrb2_abec (int nblocks, const b200mg_tile* tiles, const b200mg_box* vbox,
                       const b200mg_fab* phi_in, const b200mg_fab* phi_out, const b200mg_fab* rhs, const b200mg_fab* a,
                       const b200mg_fab* bx, const b200mg_fab* by, const b200mg_fab* bz,
                       const b200mg_fab* f, const b200mg_ifab* m,
                       double alpha, double dhx, double dhy, double dhz, int tx, int tile_y, int chunk_z, cudaStream_t s)
{
    FusedArgs A{phi_in, phi_out, rhs, a, bx, by, bz, f, m, alpha, dhx, dhy, dhz, tile_y, chunk_z, g_prefetch_planes};
    return launch<true>(nblocks, tiles, vbox, A, tx, s);
}

int b200mg_gsrb2_poisson (int nblocks, const b200mg_tile* tiles, const b200mg_box* vbox,
                          const b200mg_fab* phi_in, const b200mg_fab* phi_out, const b200mg_fab* rhs,
                          const b200mg_fab* f, const b200mg_ifab* m,
                          double dhx, double dhy, double dhz, int tx, int tile_y, int chunk_z, cudaStream_t s)
{
    FusedArgs A{phi_in, phi_out, rhs, nullptr, nullptr, nullptr, nullptr, f, m, 0.0, dhx, dhy, dhz, tile_y, chunk_z, g_prefetch_planes};
    return launch<false>(nblocks, tiles, vbox, A, tx, s);
}

void b200mg_set_gsrb2_prefetch (int planes) { g_prefetch_planes = planes < 0 ? 0 : planes; }

} // extern "C"
