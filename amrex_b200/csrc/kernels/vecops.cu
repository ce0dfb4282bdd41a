// Vector operations, reductions and halo copy kernels.  Reference rows K6, K10, K12.
#include "common.cuh"

using namespace b200mg;

namespace {

__global__ void __launch_bounds__(kTileTX * B200MG_TILE_Y)
k_setval (const b200mg_tile* __restrict__ tiles, const b200mg_box* __restrict__ vbox, const b200mg_fab* yf, double v, int ng)
{
    const b200mg_tile t = tiles[blockIdx.x];
    const b200mg_box vb = vbox[t.box];
    const auto y = view(yf[t.box]);
    tile_for(t, vb, ng, [&] (int i, int j, int k) { y(i, j, k) = v; });
}

__global__ void __launch_bounds__(kTileTX * B200MG_TILE_Y)
k_setbndry (const b200mg_tile* __restrict__ tiles, const b200mg_box* __restrict__ vbox, const b200mg_fab* yf, double v, int ng)
{
    const b200mg_tile t = tiles[blockIdx.x];
    const b200mg_box vb = vbox[t.box];
    const auto y = view(yf[t.box]);
    tile_for(t, vb, ng, [&] (int i, int j, int k) {
        const bool inside = i >= vb.lo[0] && i <= vb.hi[0] && j >= vb.lo[1] && j <= vb.hi[1] && k >= vb.lo[2] && k <= vb.hi[2];
        if (!inside) { y(i, j, k) = v; }
    });
}

__global__ void __launch_bounds__(kTileTX * B200MG_TILE_Y)
k_copy (const b200mg_tile* __restrict__ tiles, const b200mg_box* __restrict__ vbox, const b200mg_fab* yf, const b200mg_fab* xf, int ng)
{
    const b200mg_tile t = tiles[blockIdx.x];
    const b200mg_box vb = vbox[t.box];
    const auto y = view(yf[t.box]); const auto x = view(xf[t.box]);
    tile_for(t, vb, ng, [&] (int i, int j, int k) { y(i, j, k) = x(i, j, k); });
}

__global__ void __launch_bounds__(kTileTX * B200MG_TILE_Y)
k_lincomb (const b200mg_tile* __restrict__ tiles, const b200mg_box* __restrict__ vbox, const b200mg_fab* yf,
           double a, const b200mg_fab* xf, double b, int ng)
{
    const b200mg_tile t = tiles[blockIdx.x];
    const b200mg_box vb = vbox[t.box];
    const auto y = view(yf[t.box]); const auto x = view(xf[t.box]);
    tile_for(t, vb, ng, [&] (int i, int j, int k) { y(i, j, k) = a * x(i, j, k) + b * y(i, j, k); });
}

__global__ void __launch_bounds__(kTileTX * B200MG_TILE_Y)
k_plus (const b200mg_tile* __restrict__ tiles, const b200mg_box* __restrict__ vbox, const b200mg_fab* yf, double v, int ng)
{
    const b200mg_tile t = tiles[blockIdx.x];
    const b200mg_box vb = vbox[t.box];
    const auto y = view(yf[t.box]);
    tile_for(t, vb, ng, [&] (int i, int j, int k) { y(i, j, k) += v; });
}

// y = y * x (op 0) or y / x (op 1): MultiFab::Multiply / Divide (AMReX_MultiFab.H), one component per launch
template <int OP>
__global__ void __launch_bounds__(kTileTX * B200MG_TILE_Y)
k_binop (const b200mg_tile* __restrict__ tiles, const b200mg_box* __restrict__ vbox, const b200mg_fab* yf, const b200mg_fab* xf, int ng)
{
    const b200mg_tile t = tiles[blockIdx.x];
    const b200mg_box vb = vbox[t.box];
    const auto y = view(yf[t.box]); const auto x = view(xf[t.box]);
    tile_for(t, vb, ng, [&] (int i, int j, int k) { if (OP == 0) { y(i, j, k) *= x(i, j, k); } else { y(i, j, k) /= x(i, j, k); } });
}

// Two-stage deterministic reduction: per-block partials in scratch[0..n), the last block to finish
// (ticket in scratch[n]) folds them in index order and resets the ticket.
template <class Op, class F>
__device__ __forceinline__ void reduce_tiles (const b200mg_tile t, const b200mg_box& vb, double* result, double* scratch, F&& f, int ng = 0)
{
    double acc = Op::id();
    tile_for(t, vb, ng, [&] (int i, int j, int k) { acc = Op::ap(acc, f(i, j, k)); });
    acc = block_reduce<Op>(acc);
    __shared__ bool is_last;
    const int tid = threadIdx.x + threadIdx.y * blockDim.x;
    const int n = gridDim.x;
    if (tid == 0) {
        scratch[blockIdx.x] = acc;
        __threadfence();
        unsigned int* ticket = reinterpret_cast<unsigned int*>(scratch + n);
        const unsigned int prev = atomicAdd(ticket, 1u);
        is_last = (prev == unsigned(n - 1));
    }
    __syncthreads();
    if (is_last) {
        __threadfence();
        double v = Op::id();
        const int nt = blockDim.x * blockDim.y;
        // fixed partition: thread t folds partials t, t+nt, ... in order; then a block tree
        for (int b = tid; b < n; b += nt) { v = Op::ap(v, reinterpret_cast<volatile double*>(scratch)[b]); }
        __syncthreads();
        v = block_reduce<Op>(v);
        if (tid == 0) {
            result[0] = v;
            *reinterpret_cast<unsigned int*>(scratch + n) = 0u;
        }
    }
}

__global__ void __launch_bounds__(kTileTX * B200MG_TILE_Y)
k_norminf (const b200mg_tile* __restrict__ tiles, const b200mg_box* __restrict__ vbox, const b200mg_fab* xf,
           const b200mg_ifab* mf, double* result, double* scratch)
{
    const b200mg_tile t = tiles[blockIdx.x];
    const b200mg_box vb = vbox[t.box];
    const auto x = view(xf[t.box]);
    if (mf) {
        const auto m = view(mf[t.box]);
        reduce_tiles<OpMax>(t, vb, result, scratch, [&] (int i, int j, int k) { return m(i, j, k) ? fabs(x(i, j, k)) : 0.0; });
    } else {
        reduce_tiles<OpMax>(t, vb, result, scratch, [&] (int i, int j, int k) { return fabs(x(i, j, k)); });
    }
}

__global__ void __launch_bounds__(kTileTX * B200MG_TILE_Y)
k_dot (const b200mg_tile* __restrict__ tiles, const b200mg_box* __restrict__ vbox, const b200mg_fab* xf,
       const b200mg_fab* yf, double* result, double* scratch)
{
    const b200mg_tile t = tiles[blockIdx.x];
    const b200mg_box vb = vbox[t.box];
    const auto x = view(xf[t.box]); const auto y = view(yf[t.box]);
    reduce_tiles<OpSum>(t, vb, result, scratch, [&] (int i, int j, int k) { return x(i, j, k) * y(i, j, k); });
}

__global__ void __launch_bounds__(kTileTX * B200MG_TILE_Y)
k_sum (const b200mg_tile* __restrict__ tiles, const b200mg_box* __restrict__ vbox, const b200mg_fab* xf,
       double* result, double* scratch)
{
    const b200mg_tile t = tiles[blockIdx.x];
    const b200mg_box vb = vbox[t.box];
    const auto x = view(xf[t.box]);
    reduce_tiles<OpSum>(t, vb, result, scratch, [&] (int i, int j, int k) { return x(i, j, k); });
}

// signed extrema over the cells grown by ng (FabArray::min / max, AMReX_MultiFab.cpp)
struct OpMinS { __device__ static double id () { return 1.7976931348623157e308; } __device__ static double ap (double a, double b) { return fmin(a, b); } };
struct OpMaxS { __device__ static double id () { return -1.7976931348623157e308; } __device__ static double ap (double a, double b) { return fmax(a, b); } };

template <class Op>
__global__ void __launch_bounds__(kTileTX * B200MG_TILE_Y)
k_extremum (const b200mg_tile* __restrict__ tiles, const b200mg_box* __restrict__ vbox, const b200mg_fab* xf,
            double* result, double* scratch, int ng)
{
    const b200mg_tile t = tiles[blockIdx.x];
    const b200mg_box vb = vbox[t.box];
    const auto x = view(xf[t.box]);
    reduce_tiles<Op>(t, vb, result, scratch, [&] (int i, int j, int k) { return x(i, j, k); }, ng);
}

__global__ void __launch_bounds__(kTileTX * B200MG_TILE_Y)
k_asum (const b200mg_tile* __restrict__ tiles, const b200mg_box* __restrict__ vbox, const b200mg_fab* xf,
        double* result, double* scratch)
{
    const b200mg_tile t = tiles[blockIdx.x];
    const b200mg_box vb = vbox[t.box];
    const auto x = view(xf[t.box]);
    reduce_tiles<OpSum>(t, vb, result, scratch, [&] (int i, int j, int k) { return fabs(x(i, j, k)); });
}

// fab_to_fab / pack / unpack (AMReX_FBI.H:53-70, 729-893).  blockIdx.x = tag; threads sweep the tag box,
// x fastest, then y, z, component -- the order that also defines the linear buffer layout (AMReX_FBI.H:765-771).
__global__ void __launch_bounds__(256)
k_copy_tags (const b200mg_copytag* __restrict__ tags, const b200mg_fab* dstf, const b200mg_fab* srcf,
             double* __restrict__ buf, int ncomp, int scomp, int dcomp, int op, int parity)
{
    const b200mg_copytag t = tags[blockIdx.x];
    const unsigned n0 = unsigned(t.hi[0] - t.lo[0] + 1), n1 = unsigned(t.hi[1] - t.lo[1] + 1), n2 = unsigned(t.hi[2] - t.lo[2] + 1);
    const unsigned n01 = n0 * n1;
    const unsigned npts = n01 * n2;                      // a tag never exceeds one fab (< 2^31 points)
    View<double> src, dst;
    if (t.src_fab >= 0) { src = view(srcf[t.src_fab]); }
    if (t.dst_fab >= 0) { dst = view(dstf[t.dst_fab]); }
    for (int n = 0; n < ncomp; ++n) {
        // buffer layout: the tag's block starts at buf_offset * ncomp (offsets count points of one component), components follow
        // each other inside the block (AMReX_FBI.H:765-771) - so the per-peer ranges [offset, offset + count) * ncomp are contiguous
        double* b = buf ? buf + t.buf_offset * (long long)ncomp + (long long)n * npts : nullptr;
        for (unsigned r = threadIdx.x + blockIdx.y * blockDim.x; r < npts; r += blockDim.x * gridDim.y) {
            const unsigned k = r / n01, r2 = r - k * n01, j = r2 / n0, i = r2 - j * n0;
            const int ii = t.lo[0] + int(i), jj = t.lo[1] + int(j), kk = t.lo[2] + int(k);
            // one colour of the red-black lattice only (destination indices: the same cells on the packing and the unpacking side)
            if (parity >= 0 && ((ii + jj + kk) & 1) != parity) { continue; }
            double v;
            if (t.src_fab >= 0) { v = src(ii + t.shift[0], jj + t.shift[1], kk + t.shift[2], n + scomp); }
            else { v = b[r]; }
            if (t.dst_fab >= 0) {
                double& d = dst(ii, jj, kk, n + dcomp);
                // ADD: destination boxes of different tags may overlap (flux-register patches): no lost updates
                if (op == 0) { d = v; } else if (v != 0.0) { atomicAdd(&d, v); }
            } else {
                b[r] = v;
            }
        }
    }
}

} // namespace

extern "C" {

int b200mg_setval (int ntiles, const b200mg_tile* tiles, const b200mg_box* vbox, const b200mg_fab* y, double v, int ng, cudaStream_t s)
{
    if (ntiles <= 0) { return 0; }
    k_setval<<<ntiles, tile_block(), 0, s>>>(tiles, vbox, y, v, ng);
    return last_error();
}

int b200mg_setbndry (int ntiles, const b200mg_tile* tiles, const b200mg_box* vbox, const b200mg_fab* y, double v, int ng, cudaStream_t s)
{
    if (ntiles <= 0) { return 0; }
    k_setbndry<<<ntiles, tile_block(), 0, s>>>(tiles, vbox, y, v, ng);
    return last_error();
}

int b200mg_copy (int ntiles, const b200mg_tile* tiles, const b200mg_box* vbox, const b200mg_fab* y, const b200mg_fab* x, int ng, cudaStream_t s)
{
    if (ntiles <= 0) { return 0; }
    k_copy<<<ntiles, tile_block(), 0, s>>>(tiles, vbox, y, x, ng);
    return last_error();
}

int b200mg_lincomb (int ntiles, const b200mg_tile* tiles, const b200mg_box* vbox, const b200mg_fab* y,
                    double a, const b200mg_fab* x, double b, int ng, cudaStream_t s)
{
    if (ntiles <= 0) { return 0; }
    k_lincomb<<<ntiles, tile_block(), 0, s>>>(tiles, vbox, y, a, x, b, ng);
    return last_error();
}

int b200mg_plus (int ntiles, const b200mg_tile* tiles, const b200mg_box* vbox, const b200mg_fab* y, double v, int ng, cudaStream_t s)
{
    if (ntiles <= 0) { return 0; }
    k_plus<<<ntiles, tile_block(), 0, s>>>(tiles, vbox, y, v, ng);
    return last_error();
}

int b200mg_multiply (int ntiles, const b200mg_tile* tiles, const b200mg_box* vbox, const b200mg_fab* y, const b200mg_fab* x, int ng, cudaStream_t s)
{
    if (ntiles <= 0) { return 0; }
    k_binop<0><<<ntiles, tile_block(), 0, s>>>(tiles, vbox, y, x, ng);
    return last_error();
}

int b200mg_divide (int ntiles, const b200mg_tile* tiles, const b200mg_box* vbox, const b200mg_fab* y, const b200mg_fab* x, int ng, cudaStream_t s)
{
    if (ntiles <= 0) { return 0; }
    k_binop<1><<<ntiles, tile_block(), 0, s>>>(tiles, vbox, y, x, ng);
    return last_error();
}

int b200mg_minmax (int ntiles, const b200mg_tile* tiles, const b200mg_box* vbox, const b200mg_fab* x, int want_max, int ng,
                   double* result, double* scratch, cudaStream_t s)
{
    if (ntiles <= 0) {      // no local cells: the identity of the reduction
        const double v = want_max ? -1.7976931348623157e308 : 1.7976931348623157e308;
        return int(cudaMemcpyAsync(result, &v, sizeof(double), cudaMemcpyHostToDevice, s));
    }
    if (want_max) { k_extremum<OpMaxS><<<ntiles, tile_block(), 0, s>>>(tiles, vbox, x, result, scratch, ng); }
    else { k_extremum<OpMinS><<<ntiles, tile_block(), 0, s>>>(tiles, vbox, x, result, scratch, ng); }
    return last_error();
}

long long b200mg_reduce_scratch_doubles (int ntiles) { return (long long)ntiles + 2; }

int b200mg_norminf (int ntiles, const b200mg_tile* tiles, const b200mg_box* vbox, const b200mg_fab* x,
                    const b200mg_ifab* mask, double* result, double* scratch, cudaStream_t s)
{
    if (ntiles <= 0) { return int(cudaMemsetAsync(result, 0, sizeof(double), s)); }
    k_norminf<<<ntiles, tile_block(), 0, s>>>(tiles, vbox, x, mask, result, scratch);
    return last_error();
}

int b200mg_dot (int ntiles, const b200mg_tile* tiles, const b200mg_box* vbox, const b200mg_fab* x,
                const b200mg_fab* y, double* result, double* scratch, cudaStream_t s)
{
    if (ntiles <= 0) { return int(cudaMemsetAsync(result, 0, sizeof(double), s)); }
    k_dot<<<ntiles, tile_block(), 0, s>>>(tiles, vbox, x, y, result, scratch);
    return last_error();
}

int b200mg_sum (int ntiles, const b200mg_tile* tiles, const b200mg_box* vbox, const b200mg_fab* x,
                double* result, double* scratch, cudaStream_t s)
{
    if (ntiles <= 0) { return int(cudaMemsetAsync(result, 0, sizeof(double), s)); }
    k_sum<<<ntiles, tile_block(), 0, s>>>(tiles, vbox, x, result, scratch);
    return last_error();
}

int b200mg_asum (int ntiles, const b200mg_tile* tiles, const b200mg_box* vbox, const b200mg_fab* x,
                 double* result, double* scratch, cudaStream_t s)
{
    if (ntiles <= 0) { return int(cudaMemsetAsync(result, 0, sizeof(double), s)); }
    k_asum<<<ntiles, tile_block(), 0, s>>>(tiles, vbox, x, result, scratch);
    return last_error();
}

int b200mg_copy_tags_colour (int ntags, const b200mg_copytag* tags, const b200mg_fab* dst, const b200mg_fab* src,
                             double* buf, int ncomp, int scomp, int dcomp, int op, int max_pts, int parity, cudaStream_t s)
{
    if (ntags <= 0) { return 0; }
    // blockIdx.y chunks: ~2 points per thread on the largest tag (these launches are latency bound: few tags on a GPU that
    // owns few boxes), one chunk for the small tags of coarse levels; max_pts <= 0: unknown, 4 chunks
    int chunks = 4;
    if (max_pts > 0) { chunks = (max_pts + 511) / 512; chunks = chunks < 1 ? 1 : (chunks > 32 ? 32 : chunks); }
    k_copy_tags<<<dim3(ntags, chunks), 256, 0, s>>>(tags, dst, src, buf, ncomp, scomp, dcomp, op, parity);
    return last_error();
}

int b200mg_copy_tags (int ntags, const b200mg_copytag* tags, const b200mg_fab* dst, const b200mg_fab* src,
                      double* buf, int ncomp, int scomp, int dcomp, int op, int max_pts, cudaStream_t s)
{
    return b200mg_copy_tags_colour(ntags, tags, dst, src, buf, ncomp, scomp, dcomp, op, max_pts, -1, s);
}

} // extern "C"
