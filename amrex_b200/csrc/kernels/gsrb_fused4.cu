// Fused red+black Gauss-Seidel pass, fourth generation: the same algorithm, tiling and bit-exact results as
// gsrb_fused3.cu (read the header of gsrb_fused.cu for the algorithm), but memory latency is taken off the compute
// path.  Generations 2/3 are latency bound (profiles/r01_s9_gsrb_fused_v2_ncu.txt: ideal DRAM traffic, 36 % of peak
// throughput, one exposed load round trip + one CTA barrier per plane).  Here every operand plane travels
//     HBM --cp.async.bulk (TMA engine, ONE 1-D copy per array and plane)--> shared-memory ring --> compute warps,
// several planes ahead of its use, completion signalled through mbarriers with transaction counts.  The compute warps
// never issue a global load on the main path (only the O(n^2) face-relaxation slabs of surface cells) and store their
// black results straight from registers as 16-byte pairs.
//
// CTA = (all x of one box) x TY rows, marching over every z plane of the box; (nx/2 -> multiple of 32) x (TY+2)
// threads (one cell pair each; rows j0-1 .. j1+1, the two ring rows recompute red only).  The shared-memory planes keep
// the row pitch of the arrays in HBM, so the rows a tile needs from one plane of one array are ONE contiguous byte range
// = one bulk copy (6 per step, issued by thread 0 right after the barrier that frees the slots).
// Two rings, plane index q = k - (lo_z - 1):
//     EARLY[q] = phi rows j0-2..j1+2 (x from lo-2 to hi+2, 16-byte aligned) + bz rows j0-1..j1+1      lifetime: steps q-2..q
//     LATE[q]  = rhs, a, bx (rows j0-1..j1+1), by (rows j0-1..j1+2)                                   lifetime: step  q-1
// Step t does the red update of plane t+1 in place in EARLY[t+1].phi, then the black update of plane t reading the red
// x / y neighbours of EARLY[t].phi written during step t-1.  z neighbours and the z-face coefficients stay in registers.
// The coefficient pairs of plane t+1 are read ONCE as 16-byte shared-memory loads: the red half is used at once, the
// black half is carried in registers to step t+1 - which is why a LATE slot is free again after a single step.
// One CTA barrier per step (it also hands the freed slots back to the issuing thread).
//
// Requirements (checked on the host, cudaErrorInvalidValue otherwise): every box has an even x extent, 4 <= nx <= 128,
// ny >= 2; rows of every array are 16-byte aligned at the first valid cell with even strides, phi rows readable from
// lo-2 to hi+2 and bx rows up to nx+2 doubles (the FabArray allocator pads rows to 128-byte multiples: AMReX_MultiFab.cpp);
// all boxes of the launch share the row pitches, and rhs / a / by / bz share one pitch.
#include "common.cuh"
#include "stencil_math.cuh"
#include "gsrb_fused_stage.cuh"

using namespace b200mg;
using namespace b200mg::fused;

namespace {


template <bool ABEC, int TY, int SE, int SL, int C>
__device__ __forceinline__ void
step4 (const FusedParams4& P, const FusedBox4& B, const Lay<ABEC, TY>& Y, double* __restrict__ smE, double* __restrict__ smL,
       const Header* H, uint32_t barE, uint32_t barL, Ring<SE, SL>& R, int t, int nz,
       bool row_load, bool row_red, bool row_black, bool first, bool last, bool jlo, bool jhi, int tx2, int jrel,
       int prow, int crow, int xrow, int& out_cur,
       double& zlo_b, double2& pk, double2& pp1, double& bzm_b, double2& bz1, Carry& cb, int& xmk, double& xf)
{
    // The arithmetic below is straight-line code executed by EVERY thread (threads without a cell to update compute on
    // whatever their in-range shared-memory addresses hold and only their stores are predicated): the two dependent
    // chains of a step - the black cell's divide and partial sums, which do not depend on this step's red result, and
    // the red update - then sit in the same scheduling region and overlap.
    const bool do_red = row_red && (t + 1 <= nz);
    const bool do_black = row_black && (t >= 1);

    // ---- rare, warp-uniform: face relaxation coefficients (AMReX_MLABecLap_3D_K.H:228-245) of red cells on the y / z
    //      box surface, global slab lookups; the x faces (first / last lane of every row) are fetched one step ahead
    const int xmk_now = xmk; const double xf_now = xf;
    if (((C == 1) ? first : last) && row_red && t + 2 <= nz) {
        const int xs = jrel + (t + 1) * (B.hi[1] - B.lo[1] + 1);         // x slabs: 1 x ny x nz
        xmk = B.m[C ? 0 : 3][xs]; xf = B.f[C ? 0 : 3][xs];
    }
    const int kr_rel = t;                                                // red plane - lo_z
    const bool klo = (kr_rel == 0), khi = (kr_rel == nz - 1);
    const bool yz_surface = do_red && (jlo || jhi || klo || khi);
    double cf1 = 0.0, cf2 = 0.0, cf4 = 0.0, cf5 = 0.0;
    if (yz_surface) {
        const int nx = B.hi[0] - B.lo[0] + 1;
        const int ys = tx2 + C + kr_rel * nx;                            // y slabs: nx x 1 x nz
        if (jlo) { const int mk = B.m[1][ys]; const double f = B.f[1][ys]; cf1 = (mk > 0) ? f : 0.0; }
        if (jhi) { const int mk = B.m[4][ys]; const double f = B.f[4][ys]; cf4 = (mk > 0) ? f : 0.0; }
        if (klo || khi) {
            const int zo = (tx2 + C) + jrel * nx;
            if (klo) { const int mk = B.m[2][zo]; const double f = B.f[2][zo]; cf2 = (mk > 0) ? f : 0.0; }
            if (khi) { const int mk = B.m[5][zo]; const double f = B.f[5][zo]; cf5 = (mk > 0) ? f : 0.0; }
        }
    }
    double cf0 = 0.0, cf3 = 0.0;
    if (C == 0) { cf0 = (first && xmk_now > 0) ? xf_now : 0.0; } else { cf3 = (last && xmk_now > 0) ? xf_now : 0.0; }

    // ---- operands that land during this step's first use
    const uint32_t s0 = R.s0(t), s1 = R.s1(t), s2 = R.s2(t), sl = R.l(t);
    if (t + 2 <= nz + 1) { mbar_wait(barE + 8u * s2, R.par2(t)); }
    if (t + 1 <= nz)     { mbar_wait(barL + 8u * sl, R.parl(t)); }
    double* __restrict__ e2 = smE + s2 * Y.e_size;
    double* __restrict__ e1 = smE + s1 * Y.e_size;
    double* __restrict__ e0 = smE + s0 * Y.e_size;
    const double* __restrict__ l1 = smL + sl * Y.l_size;

    // zero input: the red value this thread left in plane t-1 two steps ago (last read during step t-1) is cleared before
    // the slot is used again - the copy engine does not touch the phi part of the planes in this mode
    if (P.phi_zero && t >= 2) { smE[R.sm1(t) * Y.e_size + Y.e_phi + prow + C] = 0.0; }

    // ---- black cell of plane t, part 1: everything that does not need the red value above it (computed below).
    //      New red values on five sides: EARLY[t].phi in x / y (written during step t-1), zlo_b below.
    const double pb = C ? pk.y : pk.x;
    double b_q, b_part, b_w = 0.0, b_gp = 0.0, b_bzp = 0.0;
    const double b_rhs = cb.rhs;
    {
        const double* sb = e0 + Y.e_phi + prow + C;
        const double xm = C ? pk.x : sb[-1];
        const double xp = C ? sb[1] : pk.y;
        const double ym = sb[-Y.PS], yp = sb[Y.PS];
        if constexpr (ABEC) {
            const double b_bzm = bzm_b;
            b_bzp = C ? bz1.y : bz1.x;
            const double gamma = P.alpha * cb.a + P.dhx * (cb.bxm + cb.bxp) + P.dhy * (cb.bym + cb.byp) + P.dhz * (b_bzm + b_bzp);
            b_q = kOmega / gamma;
            b_part = P.dhx * (cb.bxm * xm + cb.bxp * xp) + P.dhy * (cb.bym * ym + cb.byp * yp);
            b_w = b_bzm * zlo_b;
            b_gp = gamma * pb;
        } else {
            const double gamma = -2.0 * (P.dhx + P.dhy + P.dhz);
            b_q = kOmega / gamma;
            b_part = b_rhs - gamma * pb - P.dhx * (xm + xp) - P.dhy * (ym + yp);
        }
    }

    // ---- red update of plane t+1, in place in EARLY[t+1].phi and pp1; the coefficient pairs are read once, the black
    //      halves are carried to the next step
    const double zhi_r = e2[Y.e_phi + prow + C];
    double* sr = e1 + Y.e_phi + prow + C;
    const double pr = C ? pp1.y : pp1.x;
    double vr;
    {
        const double xm = C ? pp1.x : sr[-1];
        const double xp = C ? sr[1] : pp1.y;
        const double ym = sr[-Y.PS], yp = sr[Y.PS];
        const double zlo = C ? pk.y : pk.x;
        const double2 vrhs = *reinterpret_cast<const double2*>(l1 + Y.l_rhs + crow);
        const double r_rhs = C ? vrhs.y : vrhs.x;
        const double n_rhs = C ? vrhs.x : vrhs.y;
        if constexpr (ABEC) {
            const double bzp_r = e2[Y.e_bz + crow + C];
            const double2 va = *reinterpret_cast<const double2*>(l1 + Y.l_a + crow);
            const double2 vx = *reinterpret_cast<const double2*>(l1 + Y.l_bx + xrow);
            const double vx2 = l1[Y.l_bx + xrow + 2];
            const double2 vy0 = *reinterpret_cast<const double2*>(l1 + Y.l_by + crow);
            const double2 vy1 = *reinterpret_cast<const double2*>(l1 + Y.l_by + crow + Y.NX);
            const double r_a = C ? va.y : va.x;
            const double r_bxm = C ? vx.y : vx.x, r_bxp = C ? vx2 : vx.y;
            const double r_bym = C ? vy0.y : vy0.x, r_byp = C ? vy1.y : vy1.x;
            const double r_bzm = C ? bz1.y : bz1.x, r_bzp = bzp_r;
            const double gamma = P.alpha * r_a + P.dhx * (r_bxm + r_bxp) + P.dhy * (r_bym + r_byp) + P.dhz * (r_bzm + r_bzp);
            double corr = P.dhx * (r_bxm * cf0 + r_bxp * cf3);
            if (yz_surface) { corr = corr + P.dhy * (r_bym * cf1 + r_byp * cf4) + P.dhz * (r_bzm * cf2 + r_bzp * cf5); }
            const double g_m_d = gamma - corr;
            const double rho = P.dhx * (r_bxm * xm + r_bxp * xp) + P.dhy * (r_bym * ym + r_byp * yp) + P.dhz * (r_bzm * zlo + r_bzp * zhi_r);
            const double res = r_rhs - (gamma * pr - rho);
            vr = pr + kOmega / g_m_d * res;
            if (do_red) {
                cb.rhs = n_rhs; cb.a = C ? va.x : va.y;
                cb.bxm = C ? vx.x : vx.y; cb.bxp = C ? vx.y : vx2;
                cb.bym = C ? vy0.x : vy0.y; cb.byp = C ? vy1.x : vy1.y;
            }
        } else {
            const double gamma = -2.0 * (P.dhx + P.dhy + P.dhz);
            double g_m_d = gamma + P.dhx * (cf0 + cf3);
            if (yz_surface) { g_m_d = g_m_d + P.dhy * (cf1 + cf4) + P.dhz * (cf2 + cf5); }
            const double res = r_rhs - gamma * pr - P.dhx * (xm + xp) - P.dhy * (ym + yp) - P.dhz * (zlo + zhi_r);
            vr = pr + kOmega / g_m_d * res;
            if (do_red) { cb.rhs = n_rhs; }
        }
    }
    if (do_red) {
        if (C) { pp1.y = vr; } else { pp1.x = vr; }
        sr[0] = vr;
    }
    // ---- black cell of plane t, part 2 (needs the red value above: pp1); box-surface cells pass through unchanged and
    //      are finished by the shell kernel after the second halo refresh
    if (do_black) {
        const int kb_rel = t - 1;                                // black plane - lo_z
        const bool surf_b = jlo || jhi || (C ? last : first) || (kb_rel == 0) || (kb_rel == nz - 1);
        const double zhi = C ? pp1.y : pp1.x;
        double vb;
        if constexpr (ABEC) {
            const double rho = b_part + P.dhz * (b_w + b_bzp * zhi);
            const double res = b_rhs - (b_gp - rho);
            vb = pb + b_q * res;
        } else {
            const double res = b_part - P.dhz * (zlo_b + zhi);
            vb = pb + b_q * res;
        }
        double2 out = pk;
        if (!surf_b) { if (C) { out.y = vb; } else { out.x = vb; } }
        *reinterpret_cast<double2*>(B.pout.p + out_cur) = out;
        out_cur += B.pout.ks;
    }

    // ---- rotate, advance; generic-proxy accesses of the slots freed by this step are ordered before the refill
    // (the next step works at pair position 1-C: its black cell needs the red value below it and the z face below it)
    zlo_b = C ? pk.x : pk.y; bzm_b = C ? bz1.x : bz1.y;
    pk = pp1;
    pp1 = make_double2(0.0, 0.0); bz1 = pp1;
    if (row_load && (t + 2 <= nz + 1)) {
        pp1 = *reinterpret_cast<const double2*>(e2 + Y.e_phi + prow);
        if constexpr (ABEC) { if (row_red) { bz1 = *reinterpret_cast<const double2*>(e2 + Y.e_bz + crow); } }
    }
    fence_proxy_async();
    cta_sync();
    // ---- EARLY[t] and LATE[t+1] are free: refill them with planes t+SE and t+1+SL, then rotate the ring
    if (threadIdx.x == 0) {
        if (t + SE <= nz + 1) {
            produce(H, 0, 2, barE + 8u * s0, smem_u32(smE) + s0 * uint32_t(8 * Y.e_size), t + SE, H->bytesE);
        }
        if (t + 1 + SL <= nz) {
            produce(H, 2, 6, barL + 8u * sl, smem_u32(smL) + sl * uint32_t(8 * Y.l_size), t + 1 + SL, H->bytesL);
        }
    }
    R.advance();
}

template <bool ABEC, int TY, int SE, int SL, int MAXT>
__global__ void __launch_bounds__(MAXT, 1)
k_gsrb4 (const __grid_constant__ FusedParams4 P)
{
    extern __shared__ __align__(128) unsigned char sm_raw[];
    const FusedBox4& B = P.box[blockIdx.y];
    const int j0 = B.lo[1] + int(blockIdx.x) * TY;
    if (j0 > B.hi[1]) { return; }                                   // uniform: the whole CTA leaves
    const int j1 = min(j0 + TY - 1, B.hi[1]);
    const int nx = B.hi[0] - B.lo[0] + 1, nz = B.hi[2] - B.lo[2] + 1;
    const Lay<ABEC, TY> Y(P.nxs, P.ps, P.cs, P.xs);
    Header* H = reinterpret_cast<Header*>(sm_raw + kBarBytes);
    double* smE = reinterpret_cast<double*>(sm_raw + kHdrBytes);
    double* smL = smE + SE * Y.e_size;
    const uint32_t barE = smem_u32(sm_raw), barL = barE + 8u * SE;
    static_assert(8 * (SE + SL) <= kBarBytes, "too many stages for the mbarrier block");
    static_assert(SE >= 4 && SL >= 2, "ring depths: EARLY planes live three steps, LATE planes one");

    const int tid = int(threadIdx.x);
    if (P.phi_zero) {                                               // phi part of every EARLY slot := 0 (see FusedParams4)
        const int nphi = Y.e_bz - Y.e_phi;
        for (int i = tid; i < SE * nphi; i += int(blockDim.x)) { smE[(i / nphi) * Y.e_size + Y.e_phi + (i % nphi)] = 0.0; }
    }
    if (tid == 0) {
        for (int s = 0; s < SE + SL; ++s) { mbar_init(barE + 8u * s, 1u); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        // rows this tile needs (clipped to what exists); one contiguous range per array and plane
        const int pj_lo = max(j0 - 2, B.lo[1] - 1), pj_hi = min(j1 + 2, B.hi[1] + 1);    // phi rows (ghost rows exist)
        const int cj_lo = max(j0 - 1, B.lo[1]),     cj_hi = min(j1 + 1, B.hi[1]);        // cell-centred coefficient rows
        const int yj_hi = min(j1 + 2, B.hi[1] + 1);                                      // y-face rows cj_lo .. yj_hi
        const int kbase = B.lo[2] - 1;                                                   // plane of q = 0
        auto set = [&] (int d, const double* g, long long gstep, int soff, int nelem, int qmin, int qmax) {
            H->d[d].g = reinterpret_cast<const char*>(g); H->d[d].gstep = 8 * gstep; H->d[d].soff = uint32_t(8 * soff);
            H->d[d].bytes = uint32_t(8 * nelem); H->d[d].qmin = qmin; H->d[d].qmax = qmax;
        };
        set(0, B.pin.p + ((B.lo[0] - 2 - B.glo_in[0]) + (long long)(pj_lo - B.glo_in[1]) * B.pin.js + (long long)(kbase - B.glo_in[2]) * B.pin.ks),
            B.pin.ks, Y.e_phi + (pj_lo - (j0 - 2)) * Y.PS, P.phi_zero ? 0 : (pj_hi - pj_lo) * Y.PS + nx + 4, 0, nz + 1);
        const int ccn = (cj_hi - cj_lo) * Y.NX + nx, ccoff = (cj_lo - (j0 - 1)) * Y.NX;
        set(2, B.rhs.p + ((long long)(cj_lo - B.lo[1]) * B.rhs.js + (long long)(kbase - B.lo[2]) * B.rhs.ks), B.rhs.ks, Y.l_rhs + ccoff, ccn, 1, nz);
        if constexpr (ABEC) {
            set(1, B.bz.p + ((B.lo[0] - B.glo_b[2][0]) + (long long)(cj_lo - B.glo_b[2][1]) * B.bz.js + (long long)(kbase - B.glo_b[2][2]) * B.bz.ks),
                B.bz.ks, Y.e_bz + ccoff, ccn, 1, nz + 1);
            set(3, B.a.p + ((long long)(cj_lo - B.lo[1]) * B.a.js + (long long)(kbase - B.lo[2]) * B.a.ks), B.a.ks, Y.l_a + ccoff, ccn, 1, nz);
            set(4, B.bx.p + ((B.lo[0] - B.glo_b[0][0]) + (long long)(cj_lo - B.glo_b[0][1]) * B.bx.js + (long long)(kbase - B.glo_b[0][2]) * B.bx.ks),
                B.bx.ks, Y.l_bx + (cj_lo - (j0 - 1)) * Y.XS, (cj_hi - cj_lo) * Y.XS + nx + 2, 1, nz);
            set(5, B.by.p + ((B.lo[0] - B.glo_b[1][0]) + (long long)(cj_lo - B.glo_b[1][1]) * B.by.js + (long long)(kbase - B.glo_b[1][2]) * B.by.ks),
                B.by.ks, Y.l_by + ccoff, (yj_hi - cj_lo) * Y.NX + nx, 1, nz);
        } else {
            set(1, nullptr, 0, 0, 0, 1, 0); set(3, nullptr, 0, 0, 0, 1, 0); set(4, nullptr, 0, 0, 0, 1, 0); set(5, nullptr, 0, 0, 0, 1, 0);
        }
        H->bytesE0 = H->d[0].bytes;
        H->bytesE = H->d[0].bytes + H->d[1].bytes;
        H->bytesL = H->d[2].bytes + H->d[3].bytes + H->d[4].bytes + H->d[5].bytes;
        fence_proxy_async();
        // prologue: fill both rings
        for (int q = 0; q < SE && q <= nz + 1; ++q) {
            produce(H, 0, 2, barE + 8u * q, smem_u32(smE) + uint32_t(q) * uint32_t(8 * Y.e_size), q, q == 0 ? H->bytesE0 : H->bytesE);
        }
        for (int q = 1; q <= SL && q <= nz; ++q) {
            produce(H, 2, 6, barL + 8u * (q - 1), smem_u32(smL) + uint32_t(q - 1) * uint32_t(8 * Y.l_size), q, H->bytesL);
        }
    }
    __syncthreads();

    const int tx = tid % P.txp, ty = tid / P.txp;
    const int i0 = B.lo[0] + 2 * tx;
    const int j = j0 - 1 + ty;
    const bool xact = (2 * tx < nx);
    const bool row_load = xact && (j <= min(j1 + 1, B.hi[1] + 1));
    const bool row_red = xact && (j >= B.lo[1]) && (j <= min(j1 + 1, B.hi[1]));
    const bool row_black = xact && (ty >= 1) && (j <= j1);
    const bool first = (tx == 0), last = (i0 + 1 == B.hi[0]);
    const bool jlo = (j == B.lo[1]), jhi = (j == B.hi[1]);
    const int txe = xact ? tx : 0;                                 // idle lanes compute on in-range addresses (see step4)
    const int prow = (ty + 1) * Y.PS + 2 * txe + 2;                // own pair inside a phi plane
    const int crow = ty * Y.NX + 2 * txe;                          // ... inside rhs / a / by / bz planes
    const int xrow = ty * Y.XS + 2 * txe;                          // ... inside a bx plane

    int out_cur = (i0 - B.glo_out[0]) + (j - B.glo_out[1]) * B.pout.js + (B.lo[2] - B.glo_out[2]) * B.pout.ks;
    const int tx2 = 2 * tx, jrel = j - B.lo[1];

    // prologue: planes q = 0 (-> pk) and q = 1 (-> pp1, bz1)
    mbar_wait(barE, 0u);
    mbar_wait(barE + 8u, 0u);
    double2 pk = make_double2(0.0, 0.0), pp1 = pk, bz1 = pk;
    double zlo_b = 0.0, bzm_b = 0.0;
    if (row_load) {
        pk = *reinterpret_cast<const double2*>(smE + Y.e_phi + prow);
        pp1 = *reinterpret_cast<const double2*>(smE + Y.e_size + Y.e_phi + prow);
        if constexpr (ABEC) { if (row_red) { bz1 = *reinterpret_cast<const double2*>(smE + Y.e_size + Y.e_bz + crow); } }
    }
    Carry cb{0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
    Ring<SE, SL> R;

    const int c_first = (i0 + j + B.lo[2]) & 1;                    // pair position of the red cell of plane lo_z (step 0)
    int xmk = 0; double xf = 0.0;                                  // x-face slab values of step 0 (see step4)
    if ((c_first ? last : first) && row_red) { xmk = B.m[c_first ? 3 : 0][jrel]; xf = B.f[c_first ? 3 : 0][jrel]; }
#define B200MG_STEP4(CC, TT) step4<ABEC, TY, SE, SL, CC>(P, B, Y, smE, smL, H, barE, barL, R, TT, nz, row_load, row_red, row_black, first, last, \
                                                         jlo, jhi, tx2, jrel, prow, crow, xrow, out_cur, zlo_b, pk, pp1, bzm_b, bz1, cb, xmk, xf)
    int t = 0;
    if (c_first) {
        for (; t + 1 <= nz; t += 2) { B200MG_STEP4(1, t); B200MG_STEP4(0, t + 1); }
        if (t <= nz) { B200MG_STEP4(1, t); }
    } else {
        for (; t + 1 <= nz; t += 2) { B200MG_STEP4(0, t); B200MG_STEP4(1, t + 1); }
        if (t <= nz) { B200MG_STEP4(0, t); }
    }
#undef B200MG_STEP4
}

int g_plan_ty = 8, g_plan_se = 4, g_plan_sl = 2;                   // launch plan (b200mg_set_gsrb4_plan)
int g_plan_pairs = 0;                                               // cell pairs per thread on rows of > 64 cells: 0 auto, 1 one, 2 two

template <bool ABEC, int TY, int SE, int SL>
int launch4 (const FusedParams4& P, int nboxes, cudaStream_t s)
{
    const Lay<ABEC, TY> Y(P.nxs, P.ps, P.cs, P.xs);
    const size_t smem = size_t(kHdrBytes) + size_t(8) * (size_t(SE) * Y.e_size + size_t(SL) * Y.l_size);
    if (smem > 227 * 1024) { return int(cudaErrorInvalidValue); }
    const int nthreads = P.txp * (TY + 2);
    const dim3 grid(P.nty, nboxes, 1);
    cudaError_t e = cudaSuccess;
    if (nthreads <= 384) {
        auto kern = k_gsrb4<ABEC, TY, SE, SL, 384>;
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
        if (e != cudaSuccess) { return int(e); }
        kern<<<grid, nthreads, smem, s>>>(P);
    } else if (nthreads <= 512) {
        auto kern = k_gsrb4<ABEC, TY, SE, SL, 512>;
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
        if (e != cudaSuccess) { return int(e); }
        kern<<<grid, nthreads, smem, s>>>(P);
    } else {
        auto kern = k_gsrb4<ABEC, TY, SE, SL, 640>;
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
        if (e != cudaSuccess) { return int(e); }
        kern<<<grid, nthreads, smem, s>>>(P);
    }
    return last_error();
}

// rows per CTA tile actually used: boxes of nx <= 64 under the default 8-row plan take 16 rows, so that a CTA keeps the
// footprint it has at nx = 128 (18 warps, one CTA per SM) instead of two half-size CTAs per SM paying the per-plane
// overheads twice
int effective_tile_y (int nxmax) { return (g_plan_ty == 8 && nxmax <= 64) ? 16 : g_plan_ty; }

// cell pairs per thread: two on rows of 65 .. 128 cells (one warp per row, the kernel of gsrb_fused5.cu) when the launch
// runs several waves of CTAs.  The tiles that hold the first / last row of a box look up their y-face coefficients in
// global memory in every step; with two pairs per thread that load sits on the step's critical path (0.26 ms for such a
// tile against 0.20 ms for the others, with the producer warp's L1 prefetch; 0.33 ms without), which the average over waves
// hides (1.47 against 1.60 ms at 6.9 waves) but a launch of one or two waves shows in full: 0.260 against 0.247 ms for the
// 8 boxes of 128^3 a GPU owns in the 8-GPU run (profiles/r02_s30_single_wave.txt).
int g_num_sms = 0;
int pairs_per_thread (int nxmax, long long nctas)
{
    if (g_num_sms == 0) {
        int dev = 0; cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
        if (g_num_sms <= 0) { g_num_sms = 148; }
    }
    if (nxmax <= 64 || g_plan_pairs == 1) { return 1; }
    return (g_plan_pairs == 2 || nctas > 3LL * g_num_sms) ? 2 : 1;
}

template <bool ABEC>
int dispatch4 (const FusedParams4& P, int nboxes, cudaStream_t s)
{
    const int key = effective_tile_y(P.nxs) * 100 + g_plan_se * 10 + g_plan_sl;
#define B200MG_PLAN4(K, TYv, SEv, SLv) case K: return launch4<ABEC, TYv, SEv, SLv>(P, nboxes, s)
    switch (key) {
        B200MG_PLAN4(842, 8, 4, 2);
        B200MG_PLAN4(843, 8, 4, 3);
        B200MG_PLAN4(1642, 16, 4, 2);
        B200MG_PLAN4(1643, 16, 4, 3);
        B200MG_PLAN4(653, 6, 5, 3);
        B200MG_PLAN4(642, 6, 4, 2);
        B200MG_PLAN4(444, 4, 4, 4);
        default: return int(cudaErrorInvalidValue);
    }
#undef B200MG_PLAN4
}

FArr4 farr4 (const b200mg_fab& f) { return FArr4{f.p, int(f.jstride), int(f.kstride)}; }

bool aligned16 (const double* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

} // namespace

extern "C" {

// launch plan of the fourth-generation fused pass: rows per CTA tile (4 or 8) and ring depths (EARLY, LATE);
// supported: (8,4,2) default, (8,4,3), (6,5,3), (6,4,2), (4,4,4).  Returns 0 when the combination exists.
int b200mg_set_gsrb4_plan (int tile_y, int early_stages, int late_stages)
{
    const int key = tile_y * 100 + early_stages * 10 + late_stages;
    if (key != 843 && key != 842 && key != 653 && key != 642 && key != 444) { return int(cudaErrorInvalidValue); }
    g_plan_ty = tile_y; g_plan_se = early_stages; g_plan_sl = late_stages;
    return 0;
}

// cell pairs per thread on rows of more than 64 cells: 0 = by the size of the launch (default, see pairs_per_thread),
// 1 = one (this file), 2 = two (gsrb_fused5.cu)
void b200mg_set_gsrb4_sync (int pairs) { g_plan_pairs = (pairs == 1 || pairs == 2) ? pairs : 0; }

// HOST descriptor tables as in b200mg_gsrb3.  abec == 0: Poisson (a, bx, by, bz ignored).
int b200mg_gsrb4 (int abec, int nboxes, const b200mg_box* h_vbox,
                  const b200mg_fab* h_phi_in, const b200mg_fab* h_phi_out, const b200mg_fab* h_rhs, const b200mg_fab* h_a,
                  const b200mg_fab* h_bx, const b200mg_fab* h_by, const b200mg_fab* h_bz,
                  const b200mg_fab* h_f, const b200mg_ifab* h_m,
                  double alpha, double dhx, double dhy, double dhz, int phi_zero, cudaStream_t s)
{
    return b200mg_gsrb4_subset(abec, nboxes, nullptr, h_vbox, h_phi_in, h_phi_out, h_rhs, h_a, h_bx, h_by, h_bz, h_f, h_m,
                               alpha, dhx, dhy, dhz, phi_zero, s);
}

// the same pass over the listed local boxes only (ids == NULL: boxes 0 .. nboxes-1)
int b200mg_gsrb4_subset (int abec, int nboxes, const int* ids, const b200mg_box* h_vbox,
                         const b200mg_fab* h_phi_in, const b200mg_fab* h_phi_out, const b200mg_fab* h_rhs, const b200mg_fab* h_a,
                         const b200mg_fab* h_bx, const b200mg_fab* h_by, const b200mg_fab* h_bz,
                         const b200mg_fab* h_f, const b200mg_ifab* h_m,
                         double alpha, double dhx, double dhy, double dhz, int phi_zero, cudaStream_t s)
{
    if (nboxes <= 0) { return 0; }
    auto id = [&] (int n) { return ids ? ids[n] : n; };
    static FusedParams4 P;                              // kept off the stack; copied by value at every launch
    P.phi_zero = phi_zero ? 1 : 0;
    P.alpha = alpha; P.dhx = dhx; P.dhy = dhy; P.dhz = dhz;
    int nxmax = 0, nymax = 0;
    for (int q = 0; q < nboxes; ++q) {
        const int b = id(q);
        const int nx = h_vbox[b].hi[0] - h_vbox[b].lo[0] + 1, ny = h_vbox[b].hi[1] - h_vbox[b].lo[1] + 1;
        if (nx % 2 != 0 || nx < 4 || nx > 128 || ny < 2) { return int(cudaErrorInvalidValue); }
        nxmax = nx > nxmax ? nx : nxmax; nymax = ny > nymax ? ny : nymax;
    }
    const int ty = effective_tile_y(nxmax);
    P.nty = (nymax + ty - 1) / ty;
    const bool two_pairs = pairs_per_thread(nxmax, (long long)P.nty * nboxes) == 2;
    P.txp = two_pairs ? 32 : ((nxmax / 2 + 31) / 32) * 32;
    P.nxs = nxmax;
    P.ps = int(h_phi_in[id(0)].jstride); P.cs = int(h_rhs[id(0)].jstride); P.xs = abec ? int(h_bx[id(0)].jstride) : 0;
    for (int b0 = 0; b0 < nboxes; b0 += kMaxBoxes4) {
        const int nb = (nboxes - b0 < kMaxBoxes4) ? nboxes - b0 : kMaxBoxes4;
        for (int n = 0; n < nb; ++n) {
            const int b = id(b0 + n);
            FusedBox4& B = P.box[n];
            const int nx = h_vbox[b].hi[0] - h_vbox[b].lo[0] + 1;
            B.pin = farr4(h_phi_in[b]); B.pout = farr4(h_phi_out[b]); B.rhs = farr4(h_rhs[b]);
            for (int d = 0; d < 3; ++d) {
                B.lo[d] = h_vbox[b].lo[d]; B.hi[d] = h_vbox[b].hi[d];
                B.glo_in[d] = h_phi_in[b].lo[d]; B.glo_out[d] = h_phi_out[b].lo[d];
            }
            // bulk copies: 16-byte aligned rows, even strides, phi readable on [lo-2, hi+2]
            auto ok = [&] (const b200mg_fab& f, int need) {
                return aligned16(f.p + (B.lo[0] - f.lo[0])) && f.jstride % 2 == 0 && f.kstride % 2 == 0 && f.jstride >= need;
            };
            if (h_rhs[b].lo[0] != B.lo[0] || h_rhs[b].lo[1] != B.lo[1] || h_rhs[b].lo[2] != B.lo[2]) { return int(cudaErrorInvalidValue); }
            if (B.glo_in[0] != B.lo[0] - 1 || B.glo_in[1] != B.lo[1] - 1 || B.glo_in[2] != B.lo[2] - 1) { return int(cudaErrorInvalidValue); }
            if (!ok(h_phi_in[b], nx + 4) || !ok(h_phi_out[b], nx + 2) || !ok(h_rhs[b], nx)) { return int(cudaErrorInvalidValue); }
            if (B.pin.js != P.ps || B.rhs.js != P.cs) { return int(cudaErrorInvalidValue); }          // one pitch per launch
            if (abec) {
                B.a = farr4(h_a[b]); B.bx = farr4(h_bx[b]); B.by = farr4(h_by[b]); B.bz = farr4(h_bz[b]);
                if (h_a[b].lo[0] != B.lo[0] || h_a[b].lo[1] != B.lo[1] || h_a[b].lo[2] != B.lo[2]) { return int(cudaErrorInvalidValue); }
                if (!ok(h_a[b], nx) || !ok(h_bx[b], nx + 2) || !ok(h_by[b], nx) || !ok(h_bz[b], nx)) { return int(cudaErrorInvalidValue); }
                if (B.a.js != P.cs || B.by.js != P.cs || B.bz.js != P.cs || B.bx.js != P.xs) { return int(cudaErrorInvalidValue); }
                for (int d = 0; d < 3; ++d) { B.glo_b[0][d] = h_bx[b].lo[d]; B.glo_b[1][d] = h_by[b].lo[d]; B.glo_b[2][d] = h_bz[b].lo[d]; }
            } else {
                B.a = B.bx = B.by = B.bz = FArr4{nullptr, 0, 0};
            }
            for (int f = 0; f < 6; ++f) { B.m[f] = h_m[b * 6 + f].p; B.f[f] = h_f[b * 6 + f].p; }
        }
        // the default (8,4,2) plan runs its rows of 65 .. 128 cells with three LATE stages: two pairs per thread make the step
        // short enough that one step of prefetch distance no longer covers the HBM latency (profiles/r02_s24_tune_gsrb5.txt)
        const int sl5 = (g_plan_ty == 8 && g_plan_se == 4 && g_plan_sl == 2) ? 3 : g_plan_sl;
        int e = two_pairs ? dispatch5(abec != 0, P, nb, effective_tile_y(nxmax), g_plan_se, sl5, s)
                          : (abec ? dispatch4<true>(P, nb, s) : dispatch4<false>(P, nb, s));
        if (two_pairs && e == int(cudaErrorInvalidValue)) {
            // (row pitches under which the deeper rings of generation 5 exceed the shared memory, or a plan it does not
            //  compile: nothing was launched - the one-pair kernel takes the launch)
            P.txp = ((nxmax / 2 + 31) / 32) * 32;
            e = abec ? dispatch4<true>(P, nb, s) : dispatch4<false>(P, nb, s);
        }
        if (e != 0) { return e; }
    }
    return 0;
}

} // extern "C"
