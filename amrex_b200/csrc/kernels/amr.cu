// Coarse/fine coupling kernels of the multi-level (AMR composite) solve.
// Reference rows: a24 interpbndrydata_{x,y,z}_o3 (Src/Boundary/AMReX_InterpBndryData_3D_K.H:22-119), a23 reflux:
// yafluxreg_crseadd / yafluxreg_fineadd (Src/Boundary/AMReX_YAFluxRegister_3D_K.H:11-199) fused with the flux kernels
// mlabeclap_flux_* (AMReX_MLABecLap_3D_K.H:79-200) / mlpoisson_flux_* (AMReX_MLPoisson_3D_K.H:36-120): the fluxes are
// evaluated on the fly at the faces the register needs instead of being stored in face arrays first.
#include "common.cuh"

using namespace b200mg;

namespace {

constexpr int kNotCovered = 1;           // BndryData::not_covered (AMReX_BndryData.H:44)
constexpr int kCrseFineBoundaryCell = 1; // amrex_yafluxreg_crse_fine_boundary_cell
constexpr int kFineCell = 2;             // amrex_yafluxreg_fine_cell

__device__ __forceinline__ int coarsen_idx (int i, int r) { return (i < 0) ? -((-i + r - 1) / r) : i / r; }

// ---------------------------------------------------------------------------------- interp_bndry_o3
// One block column per (box, face) item.  bdry: one-cell slab outside the face of the FINE box (extent 0); crse: the
// boundary register of the coarsened box (one cell outside, two cells of tangential extent); mask: two cells outside, five
// cells of tangential extent, not_covered where the fine level does not cover the cell.
__global__ void __launch_bounds__(128)
k_interp_bndry_o3 (const b200mg_bcface* __restrict__ faces, const b200mg_box* __restrict__ vbox,
                   const b200mg_fab* bdryf, const b200mg_fab* crsef, const b200mg_ifab* maskf, int r)
{
    const b200mg_bcface fc = faces[blockIdx.x];
    const b200mg_box vb = vbox[fc.box];
    const auto bdry = view(bdryf[fc.box * 6 + fc.face]);
    const auto crse = view(crsef[fc.box * 6 + fc.face]);
    const auto mask = view(maskf[fc.box * 6 + fc.face]);
    const int d = fc.face % 3;
    const int g = (fc.face < 3) ? vb.lo[d] - 1 : vb.hi[d] + 1;
    const int d1 = (d == 0) ? 1 : 0, d2 = (d == 2) ? 1 : 2;      // tangential directions, d1 < d2
    const int n1 = vb.hi[d1] - vb.lo[d1] + 1, n2 = vb.hi[d2] - vb.lo[d2] + 1;
    const double rr = double(r);
    for (int t = threadIdx.x + blockIdx.y * blockDim.x; t < n1 * n2; t += blockDim.x * gridDim.y) {
        int idx[3];
        idx[d] = g; idx[d1] = vb.lo[d1] + t % n1; idx[d2] = vb.lo[d2] + t / n1;
        int ic[3] = {coarsen_idx(idx[0], r), coarsen_idx(idx[1], r), coarsen_idx(idx[2], r)};
        auto M = [&] (int o1, int o2) {   // mask at tangential offsets (in fine cells)
            int q[3] = {idx[0], idx[1], idx[2]}; q[d1] += o1; q[d2] += o2;
            return mask(q[0], q[1], q[2]) == kNotCovered;
        };
        auto C = [&] (int o1, int o2) {   // coarse value at tangential offsets (in coarse cells)
            int q[3] = {ic[0], ic[1], ic[2]}; q[d1] += o1; q[d2] += o2;
            return crse(q[0], q[1], q[2]);
        };
        // first tangential direction (the reference's dy for x faces, dx for y and z faces)
        int lo = M(-r, 0) ? -1 : 0;
        int hi = M(r, 0) ? 1 : 0;
        double fac = (hi == lo + 1) ? 1.0 : 0.5;
        const double da = fac * (C(hi, 0) - C(lo, 0));
        const double da2 = (hi == lo + 2) ? 0.5 * (C(1, 0) - 2. * C(0, 0) + C(-1, 0)) : 0.;
        // second tangential direction
        lo = M(0, -r) ? -1 : 0;
        hi = M(0, r) ? 1 : 0;
        fac = (hi == lo + 1) ? 1.0 : 0.5;
        const double db = fac * (C(0, hi) - C(0, lo));
        const double db2 = (hi == lo + 2) ? 0.5 * (C(0, 1) - 2. * C(0, 0) + C(0, -1)) : 0.;
        const double dab = (M(-r, -r) && M(r, -r) && M(-r, r) && M(r, r))
            ? 0.25 * (C(1, 1) - C(-1, 1) + C(-1, -1) - C(1, -1)) : 0.0;
        const double a = -0.5 + (idx[d1] - ic[d1] * r + 0.5) / rr;
        const double b = -0.5 + (idx[d2] - ic[d2] * r + 0.5) / rr;
        bdry(idx[0], idx[1], idx[2]) = C(0, 0) + a * da + (a * a) * da2 + b * db + (b * b) * db2 + a * b * dab;
    }
}

// ---------------------------------------------------------------------------------- reflux, coarse side
struct RefluxArgs {
    const b200mg_fab *sol, *bx, *by, *bz;    // b* == nullptr: constant coefficient 1 (Poisson)
    double fac[3];                           // b_scalar * dxinv[d]
    double dtdx[3];
};

template <bool ABEC>
__device__ __forceinline__ double face_flux (const View<double>& sol, const View<double>* b, int d, double fac, int i, int j, int k)
{
    // flux through the low face of cell (i,j,k) in direction d: -fac*b*(sol(i) - sol(i-1))   (Poisson: dxinv*(...), fac = -dxinv)
    const int im = i - (d == 0), jm = j - (d == 1), km = k - (d == 2);
    const double diff = sol(i, j, k) - sol(im, jm, km);
    if constexpr (ABEC) { return -fac * (*b)(i, j, k) * diff; }
    else { return -fac * diff; }
}

template <bool ABEC>
__global__ void __launch_bounds__(kTileTX * B200MG_TILE_Y)
k_reflux_crse (const b200mg_tile* __restrict__ tiles, const b200mg_box* __restrict__ vbox,
               const b200mg_fab* dstf, const b200mg_ifab* flagf, RefluxArgs A)
{
    const b200mg_tile t = tiles[blockIdx.x];
    const b200mg_box vb = vbox[t.box];
    const auto dst = view(dstf[t.box]);
    const auto flag = view(flagf[t.box]);
    const auto sol = view(A.sol[t.box]);
    View<double> b[3];
    if constexpr (ABEC) { b[0] = view(A.bx[t.box]); b[1] = view(A.by[t.box]); b[2] = view(A.bz[t.box]); }
    tile_for(t, vb, 0, [&] (int i, int j, int k) {
        double d = 0.0;
        if (flag(i, j, k) == kCrseFineBoundaryCell) {
            if (flag(i - 1, j, k) == kFineCell) { d -= A.dtdx[0] * face_flux<ABEC>(sol, &b[0], 0, A.fac[0], i, j, k); }
            if (flag(i + 1, j, k) == kFineCell) { d += A.dtdx[0] * face_flux<ABEC>(sol, &b[0], 0, A.fac[0], i + 1, j, k); }
            if (flag(i, j - 1, k) == kFineCell) { d -= A.dtdx[1] * face_flux<ABEC>(sol, &b[1], 1, A.fac[1], i, j, k); }
            if (flag(i, j + 1, k) == kFineCell) { d += A.dtdx[1] * face_flux<ABEC>(sol, &b[1], 1, A.fac[1], i, j + 1, k); }
            if (flag(i, j, k - 1) == kFineCell) { d -= A.dtdx[2] * face_flux<ABEC>(sol, &b[2], 2, A.fac[2], i, j, k); }
            if (flag(i, j, k + 1) == kFineCell) { d += A.dtdx[2] * face_flux<ABEC>(sol, &b[2], 2, A.fac[2], i, j, k + 1); }
        }
        dst(i, j, k) = d;
    });
}

// ---------------------------------------------------------------------------------- reflux, fine side
// One block column per coarse/fine patch fab.  Every patch cell that touches a face of the (coarsened) fine box it belongs
// to receives -/+ dtdx * (sum of the r*r fine fluxes through that coarse face); all other cells are set to zero.
template <bool ABEC>
__global__ void __launch_bounds__(128)
k_reflux_fine (const b200mg_fab* __restrict__ cfpf, const b200mg_box* __restrict__ cfbox, const int* __restrict__ fine_index,
               const b200mg_fab* maskf, RefluxArgs A, int r)
{
    const int n = blockIdx.x;
    const b200mg_fab pf = cfpf[n];
    const b200mg_box cb = cfbox[n];          // coarsened fine box
    const int fi = fine_index[n];
    const auto cfp = view(pf);
    const auto sol = view(A.sol[fi]);
    View<double> b[3];
    if constexpr (ABEC) { b[0] = view(A.bx[fi]); b[1] = view(A.by[fi]); b[2] = view(A.bz[fi]); }
    const int nx = pf.hi[0] - pf.lo[0] + 1, ny = pf.hi[1] - pf.lo[1] + 1, nz = pf.hi[2] - pf.lo[2] + 1;
    for (int t = threadIdx.x + blockIdx.y * blockDim.x; t < nx * ny * nz; t += blockDim.x * gridDim.y) {
        const int c[3] = {pf.lo[0] + t % nx, pf.lo[1] + (t / nx) % ny, pf.lo[2] + t / (nx * ny)};
        double v = 0.0;
        for (int d = 0; d < 3; ++d) {
            const int d1 = (d == 0) ? 1 : 0, d2 = (d == 2) ? 1 : 2;
            const bool tang = c[d1] >= cb.lo[d1] && c[d1] <= cb.hi[d1] && c[d2] >= cb.lo[d2] && c[d2] <= cb.hi[d2];
            if (!tang) { continue; }
            const bool lo = (c[d] == cb.lo[d] - 1), hi = (c[d] == cb.hi[d] + 1);
            if (!lo && !hi) { continue; }
            int f[3];
            f[d] = lo ? (c[d] + 1) * r : c[d] * r;          // fine index of the face (== low face of that fine cell)
            const double s = lo ? -A.dtdx[d] : A.dtdx[d];
            for (int o2 = 0; o2 < r; ++o2) {
                f[d2] = c[d2] * r + o2;
                for (int o1 = 0; o1 < r; ++o1) {
                    f[d1] = c[d1] * r + o1;
                    v += s * face_flux<ABEC>(sol, &b[d], d, A.fac[d], f[0], f[1], f[2]);
                }
            }
        }
        if (maskf) { v *= view(maskf[n])(c[0], c[1], c[2]); }
        cfp(c[0], c[1], c[2]) = v;
    }
}

} // namespace

extern "C" {

int b200mg_interp_bndry_o3 (int nfaces, const b200mg_bcface* faces, const b200mg_box* vbox,
                            const b200mg_fab* bdry, const b200mg_fab* crse, const b200mg_ifab* mask, int ratio, cudaStream_t s)
{
    if (nfaces <= 0) { return 0; }
    k_interp_bndry_o3<<<dim3(nfaces, 8), 128, 0, s>>>(faces, vbox, bdry, crse, mask, ratio);
    return last_error();
}

int b200mg_reflux_crse (int ntiles, const b200mg_tile* tiles, const b200mg_box* vbox,
                        const b200mg_fab* crse_data, const b200mg_ifab* flag, const b200mg_fab* sol,
                        const b200mg_fab* bx, const b200mg_fab* by, const b200mg_fab* bz,
                        double facx, double facy, double facz, double dtdx, double dtdy, double dtdz, cudaStream_t s)
{
    if (ntiles <= 0) { return 0; }
    RefluxArgs A{sol, bx, by, bz, {facx, facy, facz}, {dtdx, dtdy, dtdz}};
    if (bx) { k_reflux_crse<true><<<ntiles, tile_block(), 0, s>>>(tiles, vbox, crse_data, flag, A); }
    else { k_reflux_crse<false><<<ntiles, tile_block(), 0, s>>>(tiles, vbox, crse_data, flag, A); }
    return last_error();
}

int b200mg_reflux_fine (int npatches, const b200mg_fab* cfpatch, const b200mg_box* cfbox, const int* fine_index,
                        const b200mg_fab* mask, const b200mg_fab* fine_sol,
                        const b200mg_fab* bx, const b200mg_fab* by, const b200mg_fab* bz,
                        double facx, double facy, double facz, double dtdx, double dtdy, double dtdz, int ratio, cudaStream_t s)
{
    if (npatches <= 0) { return 0; }
    RefluxArgs A{fine_sol, bx, by, bz, {facx, facy, facz}, {dtdx, dtdy, dtdz}};
    if (bx) { k_reflux_fine<true><<<dim3(npatches, 4), 128, 0, s>>>(cfpatch, cfbox, fine_index, mask, A, ratio); }
    else { k_reflux_fine<false><<<dim3(npatches, 4), 128, 0, s>>>(cfpatch, cfbox, fine_index, mask, A, ratio); }
    return last_error();
}

} // extern "C"
