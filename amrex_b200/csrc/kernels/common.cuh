// Shared device helpers for the sm_100a kernels of the MLMG path.
#ifndef AMREX_B200_KERNELS_COMMON_CUH_
#define AMREX_B200_KERNELS_COMMON_CUH_

#include <cuda_runtime.h>
#include "amrex_b200_kernels.h"

namespace b200mg {

// Array4-style accessor built from a descriptor (reference contract: AMReX_Array4.H:85-137)
template <class T>
struct View {
    T* __restrict__ p;
    int l0, l1, l2;
    long long js, ks, ns;
    __device__ __forceinline__ T& operator() (int i, int j, int k) const noexcept {
        return p[(i - l0) + (j - l1) * js + (k - l2) * ks];
    }
    __device__ __forceinline__ T& operator() (int i, int j, int k, int n) const noexcept {
        return p[(i - l0) + (j - l1) * js + (k - l2) * ks + n * ns];
    }
    __device__ __forceinline__ T* ptr (int i, int j, int k) const noexcept {
        return p + (i - l0) + (j - l1) * js + (k - l2) * ks;
    }
};

__device__ __forceinline__ View<double> view (const b200mg_fab& f) noexcept {
    return View<double>{f.p, f.lo[0], f.lo[1], f.lo[2], f.jstride, f.kstride, f.nstride};
}
__device__ __forceinline__ View<int> view (const b200mg_ifab& f) noexcept {
    return View<int>{f.p, f.lo[0], f.lo[1], f.lo[2], f.jstride, f.kstride, f.nstride};
}

// Tile loop: blockDim = (TX, B200MG_TILE_Y); thread row j = j0 + threadIdx.y, planes k0..k0+TILE_Z-1,
// i strides by blockDim.x starting at the (grown) lower x bound -> consecutive lanes touch consecutive
// doubles of one row (coalesced 256 B per warp-load).
__device__ __forceinline__ int tile_nk (const b200mg_tile& t) noexcept { return t.nk > 0 ? t.nk : B200MG_TILE_Z; }

template <class F>
__device__ __forceinline__ void tile_for (const b200mg_tile t, const b200mg_box& b, int ng, F&& f)
{
    const int j = t.j0 + int(threadIdx.y);
    const int jhi = b.hi[1] + ng;
    if (j > jhi) { return; }
    const int khi = min(t.k0 + tile_nk(t) - 1, b.hi[2] + ng);
    const int ilo = b.lo[0] - ng, ihi = b.hi[0] + ng;
    for (int k = t.k0; k <= khi; ++k) {
        for (int i = ilo + int(threadIdx.x); i <= ihi; i += int(blockDim.x)) { f(i, j, k); }
    }
}

constexpr int kTileTX = 64;
inline dim3 tile_block () { return dim3(kTileTX, B200MG_TILE_Y, 1); }

// warp / block reductions (shuffles; reference role: AMReX_GpuReduce.H:55-120)
struct OpSum { __device__ static double id () { return 0.0; } __device__ static double ap (double a, double b) { return a + b; } };
struct OpMax { __device__ static double id () { return 0.0; } __device__ static double ap (double a, double b) { return fmax(a, b); } };

template <class Op>
__device__ __forceinline__ double warp_reduce (double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { v = Op::ap(v, __shfl_down_sync(0xffffffffu, v, o)); }
    return v;
}

template <class Op>
__device__ __forceinline__ double block_reduce (double v)   // result valid in thread 0
{
    __shared__ double sh[32];
    const int tid = threadIdx.x + threadIdx.y * blockDim.x;
    const int lane = tid & 31, wid = tid >> 5;
    v = warp_reduce<Op>(v);
    if (lane == 0) { sh[wid] = v; }
    __syncthreads();
    const int nw = (blockDim.x * blockDim.y + 31) >> 5;
    v = (tid < nw) ? sh[tid] : Op::id();
    if (wid == 0) { v = warp_reduce<Op>(v); }
    return v;
}

inline int last_error () { return int(cudaGetLastError()); }

} // namespace b200mg
#endif
