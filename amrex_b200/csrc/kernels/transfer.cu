// Grid-transfer kernels: restriction (cell and face averages), V-cycle prolongation, trilinear interpolation.
// Reference rows K7, K8, K8b.
#include "common.cuh"

using namespace b200mg;

namespace {

// amrex_avgdown, AMReX_MultiFabUtil_3D_C.H:381-394: sum with iref fastest, then volfrac * c
__global__ void __launch_bounds__(kTileTX * B200MG_TILE_Y)
k_restrict_cc (const b200mg_tile* __restrict__ tiles, const b200mg_box* __restrict__ cbox,
               const b200mg_fab* cf, const b200mg_fab* ff, int ratio)
{
    const b200mg_tile t = tiles[blockIdx.x];
    const b200mg_box cb = cbox[t.box];
    const auto crse = view(cf[t.box]); const auto fine = view(ff[t.box]);
    const double volfrac = 1.0 / double(ratio * ratio * ratio);
    tile_for(t, cb, 0, [&] (int i, int j, int k) {
        const int ii = i * ratio, jj = j * ratio, kk = k * ratio;
        double c = 0.0;
        for (int kr = 0; kr < ratio; ++kr)
            for (int jr = 0; jr < ratio; ++jr)
                for (int ir = 0; ir < ratio; ++ir) { c += fine(ii + ir, jj + jr, kk + kr); }
        crse(i, j, k) = volfrac * c;
    });
}

// ratio-2 specialisation: each thread reads its 2x2x2 fine cells as four 16-byte loads
__global__ void __launch_bounds__(kTileTX * B200MG_TILE_Y)
k_restrict_cc_r2 (const b200mg_tile* __restrict__ tiles, const b200mg_box* __restrict__ cbox,
                  const b200mg_fab* cf, const b200mg_fab* ff)
{
    const b200mg_tile t = tiles[blockIdx.x];
    const b200mg_box cb = cbox[t.box];
    const auto crse = view(cf[t.box]); const auto fine = view(ff[t.box]);
    tile_for(t, cb, 0, [&] (int i, int j, int k) {
        const double* p = fine.ptr(2 * i, 2 * j, 2 * k);
        double c = 0.0;
        if ((reinterpret_cast<unsigned long long>(p) & 15ull) == 0 && (fine.js & 1) == 0 && (fine.ks & 1) == 0) {
            const double2 a = *reinterpret_cast<const double2*>(p);
            const double2 b = *reinterpret_cast<const double2*>(p + fine.js);
            const double2 d = *reinterpret_cast<const double2*>(p + fine.ks);
            const double2 e = *reinterpret_cast<const double2*>(p + fine.ks + fine.js);
            c += a.x; c += a.y; c += b.x; c += b.y; c += d.x; c += d.y; c += e.x; c += e.y;
        } else {
            c += p[0]; c += p[1]; c += p[fine.js]; c += p[fine.js + 1];
            c += p[fine.ks]; c += p[fine.ks + 1]; c += p[fine.ks + fine.js]; c += p[fine.ks + fine.js + 1];
        }
        crse(i, j, k) = 0.125 * c;
    });
}

// amrex_avgdown_faces, AMReX_MultiFabUtil_3D_C.H:173-217.  cbox = coarse FACE boxes of direction dir.
__global__ void __launch_bounds__(kTileTX * B200MG_TILE_Y)
k_restrict_faces (const b200mg_tile* __restrict__ tiles, const b200mg_box* __restrict__ cbox,
                  const b200mg_fab* cf, const b200mg_fab* ff, int dir, int ratio)
{
    const b200mg_tile t = tiles[blockIdx.x];
    const b200mg_box cb = cbox[t.box];
    const auto crse = view(cf[t.box]); const auto fine = view(ff[t.box]);
    const double facInv = 1.0 / double(ratio * ratio);
    tile_for(t, cb, 0, [&] (int i, int j, int k) {
        const int ii = i * ratio, jj = j * ratio, kk = k * ratio;
        double c = 0.0;
        if (dir == 0) {
            for (int kr = 0; kr < ratio; ++kr) for (int jr = 0; jr < ratio; ++jr) { c += fine(ii, jj + jr, kk + kr); }
        } else if (dir == 1) {
            for (int kr = 0; kr < ratio; ++kr) for (int ir = 0; ir < ratio; ++ir) { c += fine(ii + ir, jj, kk + kr); }
        } else {
            for (int jr = 0; jr < ratio; ++jr) for (int ir = 0; ir < ratio; ++ir) { c += fine(ii + ir, jj + jr, kk); }
        }
        crse(i, j, k) = c * facInv;
    });
}

__device__ __forceinline__ int floor_half (int i) { return i >> 1; }   // amrex::coarsen(i,2): floor division

// V-cycle correction add, AMReX_MLCellLinOp.H:973-976: fine += crse(coarsen(i,2), ...)
__global__ void __launch_bounds__(kTileTX * B200MG_TILE_Y)
k_prolong_add (const b200mg_tile* __restrict__ tiles, const b200mg_box* __restrict__ fbox,
               const b200mg_fab* ff, const b200mg_fab* cf)
{
    const b200mg_tile t = tiles[blockIdx.x];
    const b200mg_box fb = fbox[t.box];
    const auto fine = view(ff[t.box]); const auto crse = view(cf[t.box]);
    // rows that start on an even cell and hold an even number of cells (16-byte aligned, as the FabArray allocator lays them
    // out): a thread adds ONE coarse value to a cell pair, 16-byte read-modify-write (17 B/cell at 0.67 -> 0.9 of the HBM peak)
    const int nx = fb.hi[0] - fb.lo[0] + 1;
    const bool pairs = ((fb.lo[0] | nx) & 1) == 0 && ((fine.js | fine.ks) & 1) == 0
        && (reinterpret_cast<unsigned long long>(fine.ptr(fb.lo[0], fb.lo[1], fb.lo[2])) & 15ull) == 0ull;
    if (pairs) {
        const int j = t.j0 + int(threadIdx.y);
        if (j > fb.hi[1]) { return; }
        const int khi = min(t.k0 + tile_nk(t) - 1, fb.hi[2]);
        const int ic0 = floor_half(fb.lo[0]), jc = floor_half(j);
        for (int k = t.k0; k <= khi; ++k) {
            double2* frow = reinterpret_cast<double2*>(fine.ptr(fb.lo[0], j, k));
            const double* crow = crse.ptr(ic0, jc, floor_half(k));
            for (int ip = int(threadIdx.x); 2 * ip < nx; ip += int(blockDim.x)) {
                double2 v = frow[ip];
                const double c = crow[ip];
                v.x += c; v.y += c;
                frow[ip] = v;
            }
        }
        return;
    }
    tile_for(t, fb, 0, [&] (int i, int j, int k) {
        fine(i, j, k) += crse(floor_half(i), floor_half(j), floor_half(k));
    });
}

// mlmg_lin_cc_interp_r2, AMReX_MLMG_3D_K.H:17-36 (C integer division, flat left-to-right sum)
__global__ void __launch_bounds__(kTileTX * B200MG_TILE_Y)
k_interp_cc_r2 (const b200mg_tile* __restrict__ tiles, const b200mg_box* __restrict__ fbox,
                const b200mg_fab* ff, const b200mg_fab* cf, int add)
{
    const b200mg_tile t = tiles[blockIdx.x];
    const b200mg_box fb = fbox[t.box];
    const auto fine = view(ff[t.box]); const auto cc = view(cf[t.box]);
    tile_for(t, fb, 0, [&] (int i, int j, int k) {
        const int ic = i / 2, jc = j / 2, kc = k / 2;
        const int ioff = 2 * (i - ic * 2) - 1, joff = 2 * (j - jc * 2) - 1, koff = 2 * (k - kc * 2) - 1;
        const double v = 0.421875 * cc(ic, jc, kc)
            + 0.140625 * cc(ic + ioff, jc, kc)
            + 0.140625 * cc(ic, jc + joff, kc)
            + 0.140625 * cc(ic, jc, kc + koff)
            + 0.046875 * cc(ic, jc + joff, kc + koff)
            + 0.046875 * cc(ic + ioff, jc, kc + koff)
            + 0.046875 * cc(ic + ioff, jc + joff, kc)
            + 0.015625 * cc(ic + ioff, jc + joff, kc + koff);
        if (add) { fine(i, j, k) += v; } else { fine(i, j, k) = v; }
    });
}

// average_cellcenter_to_face, arithmetic mean (AMReX_MultiFabUtil_3D_C.H:81-89); fbox = FACE boxes of direction dir
__global__ void __launch_bounds__(kTileTX * B200MG_TILE_Y)
k_cc_to_face (const b200mg_tile* __restrict__ tiles, const b200mg_box* __restrict__ fbox,
              const b200mg_fab* ff, const b200mg_fab* cf, int dir)
{
    const b200mg_tile t = tiles[blockIdx.x];
    const b200mg_box fb = fbox[t.box];
    const auto fc = view(ff[t.box]); const auto cc = view(cf[t.box]);
    const int di = (dir == 0), dj = (dir == 1), dk = (dir == 2);
    tile_for(t, fb, 0, [&] (int i, int j, int k) { fc(i, j, k) = 0.5 * (cc(i - di, j - dj, k - dk) + cc(i, j, k)); });
}

// Face-centred difference of a cell-centred field along dir on the FACE boxes of that direction:
//   mode 0  grad  : out = fac*(s(i) - s(i-e))                        compGrad,          AMReX_MLCellLinOp.H:1421-1436
//   mode 1  abec  : out = -fac*b(i)*(s(i) - s(i-e))   [* post]       mlabeclap_flux_*,  AMReX_MLABecLap_3D_K.H:79-135
//   mode 2  poiss : out =  fac*(s(i) - s(i-e))        [* post]       mlpoisson_flux_*,  AMReX_MLPoisson_3D_K.H:36-98
// post (= 1/b_scalar) is applied as a separate multiplication when != 1, as MLCellABecLap::getFluxes does (:279-288).
__global__ void __launch_bounds__(kTileTX * B200MG_TILE_Y)
k_face_flux (const b200mg_tile* __restrict__ tiles, const b200mg_box* __restrict__ fbox,
             const b200mg_fab* of, const b200mg_fab* sf, const b200mg_fab* bf, double fac, double post, int dir, int mode)
{
    const b200mg_tile t = tiles[blockIdx.x];
    const b200mg_box fb = fbox[t.box];
    const auto o = view(of[t.box]); const auto s = view(sf[t.box]);
    const int di = (dir == 0), dj = (dir == 1), dk = (dir == 2);
    if (mode == 1) {
        const auto b = view(bf[t.box]);
        tile_for(t, fb, 0, [&] (int i, int j, int k) {
            double v = -fac * b(i, j, k) * (s(i, j, k) - s(i - di, j - dj, k - dk));
            if (post != 1.0) { v = v * post; }
            o(i, j, k) = v;
        });
    } else {
        tile_for(t, fb, 0, [&] (int i, int j, int k) {
            double v = fac * (s(i, j, k) - s(i - di, j - dj, k - dk));
            if (mode == 2 && post != 1.0) { v = v * post; }
            o(i, j, k) = v;
        });
    }
}

} // namespace

extern "C" {

int b200mg_face_flux (int ntiles, const b200mg_tile* tiles, const b200mg_box* fbox,
                      const b200mg_fab* out, const b200mg_fab* sol, const b200mg_fab* b,
                      double fac, double post, int dir, int mode, cudaStream_t s)
{
    if (ntiles <= 0) { return 0; }
    if (mode == 1 && b == nullptr) { return int(cudaErrorInvalidValue); }
    k_face_flux<<<ntiles, tile_block(), 0, s>>>(tiles, fbox, out, sol, b, fac, post, dir, mode);
    return last_error();
}

int b200mg_cc_to_face (int ntiles, const b200mg_tile* tiles, const b200mg_box* fbox,
                       const b200mg_fab* face, const b200mg_fab* cc, int dir, cudaStream_t s)
{
    if (ntiles <= 0) { return 0; }
    k_cc_to_face<<<ntiles, tile_block(), 0, s>>>(tiles, fbox, face, cc, dir);
    return last_error();
}

int b200mg_restrict_cc (int ntiles, const b200mg_tile* tiles, const b200mg_box* cbox,
                        const b200mg_fab* crse, const b200mg_fab* fine, int ratio, cudaStream_t s)
{
    if (ntiles <= 0) { return 0; }
    if (ratio == 2) { k_restrict_cc_r2<<<ntiles, tile_block(), 0, s>>>(tiles, cbox, crse, fine); }
    else { k_restrict_cc<<<ntiles, tile_block(), 0, s>>>(tiles, cbox, crse, fine, ratio); }
    return last_error();
}

int b200mg_restrict_faces (int ntiles, const b200mg_tile* tiles, const b200mg_box* cbox,
                           const b200mg_fab* crse, const b200mg_fab* fine, int dir, int ratio, cudaStream_t s)
{
    if (ntiles <= 0) { return 0; }
    k_restrict_faces<<<ntiles, tile_block(), 0, s>>>(tiles, cbox, crse, fine, dir, ratio);
    return last_error();
}

int b200mg_prolong_add (int ntiles, const b200mg_tile* tiles, const b200mg_box* fbox,
                        const b200mg_fab* fine, const b200mg_fab* crse, cudaStream_t s)
{
    if (ntiles <= 0) { return 0; }
    k_prolong_add<<<ntiles, tile_block(), 0, s>>>(tiles, fbox, fine, crse);
    return last_error();
}

int b200mg_interp_cc_r2 (int ntiles, const b200mg_tile* tiles, const b200mg_box* fbox,
                         const b200mg_fab* fine, const b200mg_fab* crse, int add, cudaStream_t s)
{
    if (ntiles <= 0) { return 0; }
    k_interp_cc_r2<<<ntiles, tile_block(), 0, s>>>(tiles, fbox, fine, crse, add);
    return last_error();
}

} // extern "C"
