// Fused red+black Gauss-Seidel pass, third generation.  Same algorithm, tiling and bit-exact results as gsrb_fused.cu
// (read its header first); what changed is everything that cost instructions or registers there
// (profiles/r01_s9_gsrb_fused_v2_ncu.txt: 1.15e9 warp instructions per 512^3 pass = as many as two separate colour sweeps,
// 38 % issue utilisation, barrier-stall bound, 80 registers at 768 threads):
//   * per-box descriptors travel in KERNEL PARAMETERS (constant bank) and the box index is blockIdx.y, so bases, strides
//     and box bounds live in uniform registers instead of 30+ vector registers per thread;
//   * the colour bit is a template parameter (the z loop is unrolled by two): no selects on a runtime parity;
//   * every array is addressed by one 32-bit element cursor per thread that advances by a uniform stride per plane;
//   * the red cell of plane k+1 and the black cell of plane k sit at the SAME x position of the pair, so the z-face
//     coefficient between them is loaded once and both updates share the cursors;
//   * face relaxation coefficients: x / y face slabs get cursors set up before the loop (independent mask / value loads,
//     select afterwards), z faces are looked up only in the first / last plane of a box;
//   * one barrier per plane.
// Requirements: every box has an even x extent, 4 <= nx <= 256 (TX = 32..128 threads in x), ny >= 2.
#include "common.cuh"
#include "stencil_math.cuh"

using namespace b200mg;

namespace {

struct FArr { double* p; int js, ks; };                  // fab base (element (lo) of the GROWN box), strides in elements

struct FusedBox {                                       // 232 bytes
    FArr pin, pout, rhs, a, bx, by, bz;
    const int* m[6];                                    // mask slabs (one cell outside each face), [face]
    const double* f[6];                                 // relaxation-coefficient slabs (one cell inside each face)
    int lo[3], hi[3];                                   // valid box
    int glo_in[3], glo_out[3];                          // lower corner of the grown boxes of pin / pout
    int glo_b[3][3];                                    // lower corners of bx, by, bz (valid face boxes, no ghosts)
};

constexpr int kMaxBoxes = 64;                          // 64 x 296 B = 18.5 KB of the 32 KB kernel-parameter space

struct FusedParams {
    FusedBox box[kMaxBoxes];
    double alpha, dhx, dhy, dhz;
    int tile_y, chunk_z, nty;                           // rows / planes per CTA tile, tiles per box in y
};

struct Cursors {
    int in, out, cc, bx, by, bz;                        // element offsets at (i0, j, red plane); `out` at the black plane
    int xs, ys;                                         // x-face / y-face slab cursors at the red plane
};

// One z step: red update of plane kr = kk+1 at pair position C, then black update of plane kk at the same position.
template <bool ABEC, int C>
__device__ __forceinline__ void
fused_step (const FusedParams& P, const FusedBox& B, int kk, int k0, int k1, int kr_lo, int kr_hi,
            bool row_load, bool row_red, bool row_black, bool first, bool last, bool jlo, bool jhi, int i0, int j,
            double* __restrict__ sA, double* __restrict__ sB, double* __restrict__ sC, int srow, int SX,
            Cursors& o, double2& pm1, double2& pk, double2& pp1, double2& pp2, double& g2)
{
    const int kr = kk + 1;
    // ---- loads of this step, all independent: next phi plane (kk+3) and its x ghost value
    double2 nq = make_double2(0.0, 0.0); double gq = 0.0;
    {
        const int kl = kk + 3;
        if (row_load && kl <= B.hi[2] + 1) {
            const double* q = B.pin.p + (o.in + 2 * B.pin.ks);
            nq = *reinterpret_cast<const double2*>(q);
            if (first) { gq = q[-1]; }
            if (last) { gq = q[2]; }
        }
    }
    const bool do_red = row_red && kr >= kr_lo && kr <= kr_hi;
    const bool do_black = row_black && kk >= k0;
    const bool zsurf_b = (kk == B.lo[2]) || (kk == B.hi[2]);
    const bool surf_b = jlo || jhi || (C ? last : first) || zsurf_b;
    const bool upd_black = do_black && !surf_b;

    double r_rhs = 0, r_a = 0, r_bxm = 0, r_bxp = 0, r_bym = 0, r_byp = 0, r_bzm = 0, r_bzp = 0;
    double b_rhs = 0, b_a = 0, b_bxm = 0, b_bxp = 0, b_bym = 0, b_byp = 0, b_bzm = 0;
    if (do_red) {
        r_rhs = __ldg(B.rhs.p + o.cc + C);
        if constexpr (ABEC) {
            r_a = __ldg(B.a.p + o.cc + C);
            r_bxm = __ldg(B.bx.p + o.bx + C); r_bxp = __ldg(B.bx.p + o.bx + C + 1);
            r_bym = __ldg(B.by.p + o.by + C); r_byp = __ldg(B.by.p + o.by + C + B.by.js);
            r_bzm = __ldg(B.bz.p + o.bz + C); r_bzp = __ldg(B.bz.p + o.bz + C + B.bz.ks);
        }
    }
    if (upd_black) {
        b_rhs = __ldg(B.rhs.p + o.cc - B.rhs.ks + C);
        if constexpr (ABEC) {
            b_a = __ldg(B.a.p + o.cc - B.a.ks + C);
            b_bxm = __ldg(B.bx.p + o.bx - B.bx.ks + C); b_bxp = __ldg(B.bx.p + o.bx - B.bx.ks + C + 1);
            b_bym = __ldg(B.by.p + o.by - B.by.ks + C); b_byp = __ldg(B.by.p + o.by - B.by.ks + C + B.by.js);
            b_bzm = __ldg(B.bz.p + o.bz - B.bz.ks + C);
            if (!do_red) { r_bzm = __ldg(B.bz.p + o.bz + C); }     // the face between the two cells (normally loaded by red)
        }
    }

    // ---- red update of plane kr, in place in sB and pp1
    if (do_red) {
        const double* s = sB + srow + C;
        const double p = C ? pp1.y : pp1.x;
        const double xm = s[-1], xp = s[1], ym = s[-SX], yp = s[SX];
        const double zlo = C ? pk.y : pk.x, zhi = C ? pp2.y : pp2.x;
        // face relaxation coefficients (AMReX_MLABecLap_3D_K.H:228-245)
        double cf0 = 0.0, cf3 = 0.0;
        if (C == 0) { if (first) { const int mk = B.m[0][o.xs]; const double f = B.f[0][o.xs]; cf0 = (mk > 0) ? f : 0.0; } }
        else        { if (last)  { const int mk = B.m[3][o.xs]; const double f = B.f[3][o.xs]; cf3 = (mk > 0) ? f : 0.0; } }
        const bool klo = (kr == B.lo[2]), khi = (kr == B.hi[2]);
        const bool yz_surface = jlo || jhi || klo || khi;      // warp-uniform
        double cf1 = 0.0, cf2 = 0.0, cf4 = 0.0, cf5 = 0.0;
        if (yz_surface) {
            if (jlo) { const int mk = B.m[1][o.ys + C]; const double f = B.f[1][o.ys + C]; cf1 = (mk > 0) ? f : 0.0; }
            if (jhi) { const int mk = B.m[4][o.ys + C]; const double f = B.f[4][o.ys + C]; cf4 = (mk > 0) ? f : 0.0; }
            if (klo || khi) {
                const int nx = B.hi[0] - B.lo[0] + 1;
                const int zo = (i0 + C - B.lo[0]) + (j - B.lo[1]) * nx;
                if (klo) { const int mk = B.m[2][zo]; const double f = B.f[2][zo]; cf2 = (mk > 0) ? f : 0.0; }
                if (khi) { const int mk = B.m[5][zo]; const double f = B.f[5][zo]; cf5 = (mk > 0) ? f : 0.0; }
            }
        }
        double v;
        if constexpr (ABEC) {
            const double gamma = P.alpha * r_a + P.dhx * (r_bxm + r_bxp) + P.dhy * (r_bym + r_byp) + P.dhz * (r_bzm + r_bzp);
            double corr = P.dhx * (r_bxm * cf0 + r_bxp * cf3);
            if (yz_surface) { corr = corr + P.dhy * (r_bym * cf1 + r_byp * cf4) + P.dhz * (r_bzm * cf2 + r_bzp * cf5); }
            const double g_m_d = gamma - corr;
            const double rho = P.dhx * (r_bxm * xm + r_bxp * xp) + P.dhy * (r_bym * ym + r_byp * yp) + P.dhz * (r_bzm * zlo + r_bzp * zhi);
            const double res = r_rhs - (gamma * p - rho);
            v = p + kOmega / g_m_d * res;
        } else {
            const double gamma = -2.0 * (P.dhx + P.dhy + P.dhz);
            double g_m_d = gamma + P.dhx * (cf0 + cf3);
            if (yz_surface) { g_m_d = g_m_d + P.dhy * (cf1 + cf4) + P.dhz * (cf2 + cf5); }
            const double res = r_rhs - gamma * p - P.dhx * (xm + xp) - P.dhy * (ym + yp) - P.dhz * (zlo + zhi);
            v = p + kOmega / g_m_d * res;
        }
        if (C) { pp1.y = v; } else { pp1.x = v; }
        sB[srow + C] = v;
    }

    // ---- black update of plane kk (new red values on all six sides: sA in x/y, pm1 / pp1 in z); box-surface cells pass
    //      through unchanged and are finished by the shell kernel after the second halo refresh
    if (do_black) {
        double2 out = pk;
        if (upd_black) {
            const double* s = sA + srow + C;
            const double p = C ? pk.y : pk.x;
            const double xm = s[-1], xp = s[1], ym = s[-SX], yp = s[SX];
            const double zlo = C ? pm1.y : pm1.x, zhi = C ? pp1.y : pp1.x;
            double v;
            if constexpr (ABEC) {
                const double b_bzp = r_bzm;
                const double gamma = P.alpha * b_a + P.dhx * (b_bxm + b_bxp) + P.dhy * (b_bym + b_byp) + P.dhz * (b_bzm + b_bzp);
                const double rho = P.dhx * (b_bxm * xm + b_bxp * xp) + P.dhy * (b_bym * ym + b_byp * yp) + P.dhz * (b_bzm * zlo + b_bzp * zhi);
                const double res = b_rhs - (gamma * p - rho);
                v = p + kOmega / gamma * res;
            } else {
                const double gamma = -2.0 * (P.dhx + P.dhy + P.dhz);
                const double res = b_rhs - gamma * p - P.dhx * (xm + xp) - P.dhy * (ym + yp) - P.dhz * (zlo + zhi);
                v = p + kOmega / gamma * res;
            }
            if (C) { out.y = v; } else { out.x = v; }
        }
        *reinterpret_cast<double2*>(B.pout.p + o.out) = out;
        o.out += B.pout.ks;
    }

    // ---- old plane kk+2 into the free buffer for the next step's red phase, rotate, advance
    // (only lanes that own cells store: with nx < 2*blockDim.x the first idle lane's pair would land on the ghost column)
    if (row_load) { *reinterpret_cast<double2*>(sC + srow) = pp2; }
    if (first) { sC[srow - 1] = g2; }
    if (last) { sC[srow + 2] = g2; }
    pm1 = pk; pk = pp1; pp1 = pp2; pp2 = nq; g2 = gq;
    o.in += B.pin.ks; o.cc += B.rhs.ks;
    if constexpr (ABEC) { o.bx += B.bx.ks; o.by += B.by.ks; o.bz += B.bz.ks; }
    const int ny = B.hi[1] - B.lo[1] + 1, nx = B.hi[0] - B.lo[0] + 1;
    o.xs += ny; o.ys += nx;
    __syncthreads();
}

template <bool ABEC, int MAXT>
__global__ void __launch_bounds__(MAXT)
k_gsrb3 (const __grid_constant__ FusedParams P)
{
    extern __shared__ double sm[];
    const FusedBox& B = P.box[blockIdx.y];
    const int tx = threadIdx.x, ty = threadIdx.y;
    const int j0 = B.lo[1] + int(blockIdx.x % P.nty) * P.tile_y;
    const int k0 = B.lo[2] + int(blockIdx.x / P.nty) * P.chunk_z;
    if (j0 > B.hi[1] || k0 > B.hi[2]) { return; }                   // uniform: the whole CTA leaves
    const int SX = 2 * int(blockDim.x) + 4;                         // cell i lives at column i - lo_x + 2
    const int psz = int(blockDim.y) * SX;
    double* sA = sm; double* sB = sm + psz; double* sC = sm + 2 * psz;

    const int i0 = B.lo[0] + 2 * tx;
    const bool xact = (i0 < B.hi[0]);
    const int j = j0 - 2 + ty;
    const int j1 = min(j0 + P.tile_y - 1, B.hi[1]);
    const int k1 = min(k0 + P.chunk_z - 1, B.hi[2]);
    const bool row_load = xact && (j >= B.lo[1] - 1) && (j <= min(j1 + 2, B.hi[1] + 1));
    const bool row_red = xact && (j >= max(j0 - 1, B.lo[1])) && (j <= min(j1 + 1, B.hi[1]));
    const bool row_black = xact && (j >= j0) && (j <= j1);
    const bool first = row_load && (tx == 0), last = row_load && (i0 + 1 == B.hi[0]);
    const bool jlo = (j == B.lo[1]), jhi = (j == B.hi[1]);
    const int srow = ty * SX + 2 * tx + 2;
    const int nx = B.hi[0] - B.lo[0] + 1, ny = B.hi[1] - B.lo[1] + 1;

    // cursors at the first red plane of the loop, k0 - 1 (the loop starts at kk = k0 - 2); never dereferenced out of range
    Cursors o;
    const int kr0 = k0 - 1;
    o.in = (i0 - B.glo_in[0]) + (j - B.glo_in[1]) * B.pin.js + (kr0 - B.glo_in[2]) * B.pin.ks;
    o.out = (i0 - B.glo_out[0]) + (j - B.glo_out[1]) * B.pout.js + (k0 - B.glo_out[2]) * B.pout.ks;
    o.cc = (i0 - B.lo[0]) + (j - B.lo[1]) * B.rhs.js + (kr0 - B.lo[2]) * B.rhs.ks;
    o.bx = o.by = o.bz = 0;
    if constexpr (ABEC) {
        o.bx = (i0 - B.glo_b[0][0]) + (j - B.glo_b[0][1]) * B.bx.js + (kr0 - B.glo_b[0][2]) * B.bx.ks;
        o.by = (i0 - B.glo_b[1][0]) + (j - B.glo_b[1][1]) * B.by.js + (kr0 - B.glo_b[1][2]) * B.by.ks;
        o.bz = (i0 - B.glo_b[2][0]) + (j - B.glo_b[2][1]) * B.bz.js + (kr0 - B.glo_b[2][2]) * B.bz.ks;
    }
    o.xs = (j - B.lo[1]) + (kr0 - B.lo[2]) * ny;                    // x slabs: 1 x ny x nz
    o.ys = (i0 - B.lo[0]) + (kr0 - B.lo[2]) * nx;                   // y slabs: nx x 1 x nz

    // prologue: planes k0-2 (-> pk), k0-1 (-> pp1 and sB), k0 (-> pp2, stored by the first step)
    auto load_at = [&] (int k, double2& v, double& g) {
        v = make_double2(0.0, 0.0); g = 0.0;
        if (row_load && k >= B.lo[2] - 1 && k <= B.hi[2] + 1) {
            const double* q = B.pin.p + (o.in + (k - kr0) * B.pin.ks);
            v = *reinterpret_cast<const double2*>(q);
            if (first) { g = q[-1]; }
            if (last) { g = q[2]; }
        }
    };
    double2 pm1 = make_double2(0.0, 0.0), pk, pp1, pp2;
    double g0, g1, g2;
    load_at(k0 - 2, pk, g0);
    load_at(k0 - 1, pp1, g1);
    if (row_load) { *reinterpret_cast<double2*>(sB + srow) = pp1; }
    if (first) { sB[srow - 1] = g1; }
    if (last) { sB[srow + 2] = g1; }
    load_at(k0, pp2, g2);
    __syncthreads();

    const int kr_lo = max(k0 - 1, B.lo[2]), kr_hi = min(k1 + 1, B.hi[2]);
    const int c_first = ((i0 + j) + (k0 - 2) + 1) & 1;              // pair position of the red cell of plane kr at kk = k0-2
#define B200MG_FSTEP(CC, KK) do { fused_step<ABEC, CC>(P, B, KK, k0, k1, kr_lo, kr_hi, row_load, row_red, row_black, first, last, jlo, jhi, i0, j, \
                                                       sA, sB, sC, srow, SX, o, pm1, pk, pp1, pp2, g2); \
                                  double* t_ = sA; sA = sB; sB = sC; sC = t_; } while (0)
    int kk = k0 - 2;
    if (c_first) {
        for (; kk + 1 <= k1; kk += 2) { B200MG_FSTEP(1, kk); B200MG_FSTEP(0, kk + 1); }
        if (kk <= k1) { B200MG_FSTEP(1, kk); }
    } else {
        for (; kk + 1 <= k1; kk += 2) { B200MG_FSTEP(0, kk); B200MG_FSTEP(1, kk + 1); }
        if (kk <= k1) { B200MG_FSTEP(0, kk); }
    }
#undef B200MG_FSTEP
}

template <bool ABEC>
int launch3 (FusedParams& P, int nboxes, int tx, int ntz, cudaStream_t s)
{
    const dim3 block(tx, P.tile_y + 4, 1);
    const dim3 grid(P.nty * ntz, nboxes, 1);
    const int nthreads = int(block.x * block.y);
    const size_t smem = size_t(3) * block.y * (2 * block.x + 4) * sizeof(double);
    if (nthreads <= 512) {
        auto kern = k_gsrb3<ABEC, 512>;
        if (smem > 48 * 1024) { cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)); }
        kern<<<grid, block, smem, s>>>(P);
    } else if (nthreads <= 768) {
        auto kern = k_gsrb3<ABEC, 768>;
        if (smem > 48 * 1024) { cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)); }
        kern<<<grid, block, smem, s>>>(P);
    } else {
        auto kern = k_gsrb3<ABEC, 1024>;
        if (smem > 48 * 1024) { cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)); }
        kern<<<grid, block, smem, s>>>(P);
    }
    return last_error();
}

FArr farr (const b200mg_fab& f) { return FArr{f.p, int(f.jstride), int(f.kstride)}; }

} // namespace

extern "C" {

// HOST descriptor tables (one entry per local box; f / m: [box*6+face]).  abec == 0: Poisson (a, bx, by, bz ignored).
int b200mg_gsrb3 (int abec, int nboxes, const b200mg_box* h_vbox,
                  const b200mg_fab* h_phi_in, const b200mg_fab* h_phi_out, const b200mg_fab* h_rhs, const b200mg_fab* h_a,
                  const b200mg_fab* h_bx, const b200mg_fab* h_by, const b200mg_fab* h_bz,
                  const b200mg_fab* h_f, const b200mg_ifab* h_m,
                  double alpha, double dhx, double dhy, double dhz, int tile_y, int chunk_z, cudaStream_t s)
{
    if (nboxes <= 0) { return 0; }
    if (tile_y < 1 || chunk_z < 1) { return int(cudaErrorInvalidValue); }
    static FusedParams P;                               // 18.5 KB: kept off the stack; copied by value at every launch
    P.alpha = alpha; P.dhx = dhx; P.dhy = dhy; P.dhz = dhz; P.tile_y = tile_y; P.chunk_z = chunk_z;
    int nxmax = 0, nymax = 0, nzmax = 0;
    for (int b = 0; b < nboxes; ++b) {
        const int nx = h_vbox[b].hi[0] - h_vbox[b].lo[0] + 1, ny = h_vbox[b].hi[1] - h_vbox[b].lo[1] + 1, nz = h_vbox[b].hi[2] - h_vbox[b].lo[2] + 1;
        if (nx % 2 != 0 || nx < 4 || nx > 256 || ny < 2) { return int(cudaErrorInvalidValue); }
        nxmax = nx > nxmax ? nx : nxmax; nymax = ny > nymax ? ny : nymax; nzmax = nz > nzmax ? nz : nzmax;
    }
    const int tx = ((nxmax / 2 + 31) / 32) * 32;
    if (tx * (tile_y + 4) > 1024) { return int(cudaErrorInvalidValue); }
    P.nty = (nymax + tile_y - 1) / tile_y;
    const int ntz = (nzmax + chunk_z - 1) / chunk_z;
    for (int b0 = 0; b0 < nboxes; b0 += kMaxBoxes) {
        const int nb = (nboxes - b0 < kMaxBoxes) ? nboxes - b0 : kMaxBoxes;
        for (int n = 0; n < nb; ++n) {
            const int b = b0 + n;
            FusedBox& B = P.box[n];
            B.pin = farr(h_phi_in[b]); B.pout = farr(h_phi_out[b]); B.rhs = farr(h_rhs[b]);
            for (int d = 0; d < 3; ++d) {
                B.lo[d] = h_vbox[b].lo[d]; B.hi[d] = h_vbox[b].hi[d];
                B.glo_in[d] = h_phi_in[b].lo[d]; B.glo_out[d] = h_phi_out[b].lo[d];
            }
            if (h_rhs[b].lo[0] != B.lo[0] || h_rhs[b].lo[1] != B.lo[1] || h_rhs[b].lo[2] != B.lo[2]) { return int(cudaErrorInvalidValue); }
            if (abec) {
                B.a = farr(h_a[b]); B.bx = farr(h_bx[b]); B.by = farr(h_by[b]); B.bz = farr(h_bz[b]);
                if (B.a.js != B.rhs.js || B.a.ks != B.rhs.ks || h_a[b].lo[0] != B.lo[0] || h_a[b].lo[1] != B.lo[1] || h_a[b].lo[2] != B.lo[2]) {
                    return int(cudaErrorInvalidValue);          // rhs and a must share a layout (one cursor)
                }
                for (int d = 0; d < 3; ++d) { B.glo_b[0][d] = h_bx[b].lo[d]; B.glo_b[1][d] = h_by[b].lo[d]; B.glo_b[2][d] = h_bz[b].lo[d]; }
            } else {
                B.a = B.bx = B.by = B.bz = FArr{nullptr, 0, 0};
            }
            for (int f = 0; f < 6; ++f) { B.m[f] = h_m[b * 6 + f].p; B.f[f] = h_f[b * 6 + f].p; }
        }
        const int e = abec ? launch3<true>(P, nb, tx, ntz, s) : launch3<false>(P, nb, tx, ntz, s);
        if (e != 0) { return e; }
    }
    return 0;
}

} // extern "C"
