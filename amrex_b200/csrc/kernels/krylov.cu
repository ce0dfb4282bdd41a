// Batched vector kernels of the Krylov solvers (GMRES orthogonalisation, reference role: the dotProduct / increment
// loops of GMRES::gram_schmidt_orthogonalization, AMReX_GMRES.H:322-348, which launch one reduction and one axpy per
// basis vector and re-read the new vector every time).  Here the new vector is read ONCE per group of up to 8 basis
// vectors: all inner products of a group come out of one pass, and the whole update w -= sum_j h_j v_j is one pass that
// applies the axpys in the reference's order in registers (so the result is bit-identical to the one-by-one sequence).
#include "common.cuh"

using namespace b200mg;

namespace {

constexpr int kGroup = B200MG_KRYLOV_GROUP;

struct FabTables { const b200mg_fab* v[kGroup]; };
struct Coefs { double a[kGroup]; };

// partial[n * gridDim.x + block] = sum over the block's tile of x * v_n
__global__ void __launch_bounds__(kTileTX * B200MG_TILE_Y)
k_multi_dot (const b200mg_tile* __restrict__ tiles, const b200mg_box* __restrict__ vbox, const b200mg_fab* xf,
             FabTables V, int nv, double* __restrict__ partial)
{
    const b200mg_tile t = tiles[blockIdx.x];
    const b200mg_box vb = vbox[t.box];
    const auto x = view(xf[t.box]);
    View<double> y[kGroup];
#pragma unroll
    for (int n = 0; n < kGroup; ++n) { y[n] = view(V.v[n < nv ? n : 0][t.box]); }
    double acc[kGroup];
#pragma unroll
    for (int n = 0; n < kGroup; ++n) { acc[n] = 0.0; }
    tile_for(t, vb, 0, [&] (int i, int j, int k) {
        const double xv = x(i, j, k);
#pragma unroll
        for (int n = 0; n < kGroup; ++n) { if (n < nv) { acc[n] += xv * y[n](i, j, k); } }
    });
    const int tid = threadIdx.x + threadIdx.y * blockDim.x;
#pragma unroll
    for (int n = 0; n < kGroup; ++n) {
        if (n < nv) {                                   // uniform
            const double r = block_reduce<OpSum>(acc[n]);
            if (tid == 0) { partial[(long long)n * gridDim.x + blockIdx.x] = r; }
            __syncthreads();                            // block_reduce's staging array is reused
        }
    }
}

// result[n] = sum of partial[n * nblocks + b] in a fixed order (thread t folds b = t, t + 256, ...; then a block tree)
__global__ void __launch_bounds__(256)
k_fold_partials (const double* __restrict__ partial, int nblocks, double* __restrict__ result)
{
    const double* p = partial + (long long)blockIdx.x * nblocks;
    double v = 0.0;
    for (int b = threadIdx.x; b < nblocks; b += blockDim.x) { v += p[b]; }
    v = block_reduce<OpSum>(v);
    if (threadIdx.x == 0) { result[blockIdx.x] = v; }
}

// w = (...((a_0 v_0 + w) + a_1 v_1 ...) : the same sequence of roundings as nv calls of y = a*x + 1.0*y (k_lincomb)
__global__ void __launch_bounds__(kTileTX * B200MG_TILE_Y)
k_multi_axpy (const b200mg_tile* __restrict__ tiles, const b200mg_box* __restrict__ vbox, const b200mg_fab* wf,
              FabTables V, Coefs C, int nv)
{
    const b200mg_tile t = tiles[blockIdx.x];
    const b200mg_box vb = vbox[t.box];
    const auto w = view(wf[t.box]);
    View<double> y[kGroup];
#pragma unroll
    for (int n = 0; n < kGroup; ++n) { y[n] = view(V.v[n < nv ? n : 0][t.box]); }
    tile_for(t, vb, 0, [&] (int i, int j, int k) {
        double r = w(i, j, k);
#pragma unroll
        for (int n = 0; n < kGroup; ++n) { if (n < nv) { r = C.a[n] * y[n](i, j, k) + 1.0 * r; } }
        w(i, j, k) = r;
    });
}

} // namespace

extern "C" {

long long b200mg_multi_dot_scratch_doubles (int ntiles) { return (long long)kGroup * (ntiles > 0 ? ntiles : 1); }

int b200mg_multi_dot (int ntiles, const b200mg_tile* tiles, const b200mg_box* vbox, const b200mg_fab* x,
                      int nv, const b200mg_fab* const* v, double* result, double* scratch, cudaStream_t s)
{
    if (nv < 0 || nv > kGroup) { return int(cudaErrorInvalidValue); }
    if (nv == 0) { return 0; }
    if (ntiles <= 0) { return int(cudaMemsetAsync(result, 0, sizeof(double) * nv, s)); }
    FabTables V;
    for (int n = 0; n < kGroup; ++n) { V.v[n] = v[n < nv ? n : 0]; }
    k_multi_dot<<<ntiles, tile_block(), 0, s>>>(tiles, vbox, x, V, nv, scratch);
    k_fold_partials<<<nv, 256, 0, s>>>(scratch, ntiles, result);
    return last_error();
}

int b200mg_multi_axpy (int ntiles, const b200mg_tile* tiles, const b200mg_box* vbox, const b200mg_fab* w,
                       int nv, const b200mg_fab* const* v, const double* a, cudaStream_t s)
{
    if (nv < 0 || nv > kGroup) { return int(cudaErrorInvalidValue); }
    if (nv == 0 || ntiles <= 0) { return 0; }
    FabTables V; Coefs C;
    for (int n = 0; n < kGroup; ++n) { V.v[n] = v[n < nv ? n : 0]; C.a[n] = (n < nv) ? a[n] : 0.0; }
    k_multi_axpy<<<ntiles, tile_block(), 0, s>>>(tiles, vbox, w, V, C, nv);
    return last_error();
}

} // extern "C"
