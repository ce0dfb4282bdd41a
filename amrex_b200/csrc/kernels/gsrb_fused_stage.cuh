// Shared pieces of the fused red+black Gauss-Seidel passes (gsrb_fused4.cu: one cell pair per thread, gsrb_fused5.cu: two):
// kernel-parameter descriptors, the PTX wrappers of the bulk-copy pipeline (cp.async.bulk + mbarrier transaction counts),
// the shared-memory layout of the two plane rings and the ring arithmetic.  Read the header of gsrb_fused4.cu first.
#ifndef AMREX_B200_GSRB_FUSED_STAGE_CUH_
#define AMREX_B200_GSRB_FUSED_STAGE_CUH_

#include "common.cuh"

#include <cstdint>

namespace b200mg { namespace fused {

struct FArr4 { double* p; int js, ks; };                // fab base (element (lo) of the fab's own box), strides in elements

struct FusedBox4 {
    FArr4 pin, pout, rhs, a, bx, by, bz;
    const int* m[6];                                    // mask slabs (one cell outside each face), [face]
    const double* f[6];                                 // relaxation-coefficient slabs (one cell inside each face)
    int lo[3], hi[3];                                   // valid box
    int glo_in[3], glo_out[3];                          // lower corner of the grown boxes of pin / pout
    int glo_b[3][3];                                    // lower corners of bx, by, bz
};

constexpr int kMaxBoxes4 = 64;

struct FusedParams4 {
    FusedBox4 box[kMaxBoxes4];
    double alpha, dhx, dhy, dhz;
    int nty;                                            // y tiles per box
    int txp;                                            // compute threads per row (multiple of 32)
    int nxs;                                            // longest row (max nx, even)
    int ps, cs, xs;                                     // row pitches (elements) of phi, of rhs / a / by / bz, and of bx
    int phi_zero;                                       // 1: the input is identically zero (first smooth of a V-cycle): it is
                                                        // not read from HBM, the shared-memory planes are zero-filled instead
};

// ---------------------------------------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ uint32_t smem_u32 (const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void mbar_init (uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx (uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait (uint32_t bar, uint32_t parity)
{
    uint32_t done;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    } while (!done);
}
__device__ __forceinline__ void bulk_g2s (uint32_t dst, const void* src, uint32_t bytes, uint32_t bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void fence_proxy_async () { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
// CTA-wide barrier reached from two different code paths (compute warps inside step4, the producer warp in its loop)
__device__ __forceinline__ void cta_sync () { asm volatile("bar.sync 1, %0;" :: "r"(int(blockDim.x)) : "memory"); }

// ---------------------------------------------------------------------------------------------- shared-memory layout
template <bool ABEC, int TY>
struct Lay {
    int PS, XS, NX;                                     // row pitches: phi, bx, the others (NX: rhs / a / by / bz)
    int e_phi, e_bz, e_size;                            // element offsets inside an EARLY stage
    int l_rhs, l_a, l_bx, l_by, l_size;                 // ... inside a LATE stage
    __host__ __device__ Lay (int nxs, int ps, int cs, int xs)
    {
        NX = cs; PS = ps; XS = xs;
        const int cc = (TY + 1) * cs + nxs;             // TY+2 rows of a cell-centred array, last row without its padding
        e_phi = 0; e_bz = (TY + 3) * ps + nxs + 4; e_size = e_bz + (ABEC ? cc : 0);
        l_rhs = 0; l_a = cc;
        l_bx = l_a + (ABEC ? cc : 0);
        l_by = l_bx + (ABEC ? (TY + 1) * xs + nxs + 2 : 0);
        l_size = l_by + (ABEC ? (TY + 2) * cs + nxs : 0);
    }
};

constexpr int kBarBytes = 128;                           // mbarriers live in the first 128 bytes of dynamic shared memory
constexpr int kHdrBytes = 384;                           // ... followed by the copy descriptors; the rings start here

// one bulk copy per array and plane: global range of plane q = g + q*gstep (bytes), `bytes` into slot offset soff
struct CopyDesc { const char* g; long long gstep; uint32_t soff; uint32_t bytes; int qmin, qmax; };   // 32 bytes
struct Header {                                          // shared memory behind the mbarriers
    CopyDesc d[6];                                       // EARLY: phi, bz; LATE: rhs, a, bx, by
    uint32_t bytesE0, bytesE, bytesL;                    // transaction bytes of EARLY[0], EARLY[q >= 1], LATE[q]
};
static_assert(kBarBytes + sizeof(Header) <= kHdrBytes, "header does not fit");

// Ring positions of step t: EARLY slots of planes t, t+1, t+2 (+ phase parity of t+2); LATE slot of plane t+1 (+ parity).
// Power-of-two depths are computed from t (no state); other depths keep counters.
template <int SE, int SL>
struct Ring {
    static constexpr bool pe = (SE & (SE - 1)) == 0, pl = (SL & (SL - 1)) == 0;
    uint32_t c2 = 2u, cp2 = 0u, cl = 0u, cpl = 0u;
    __device__ __forceinline__ uint32_t sm1 (int t) const { return pe ? uint32_t(t - 1) & (SE - 1) : (c2 >= 3u ? c2 - 3u : c2 + SE - 3u); }
    __device__ __forceinline__ uint32_t s0 (int t) const { return pe ? uint32_t(t) & (SE - 1) : (c2 >= 2u ? c2 - 2u : c2 + SE - 2u); }
    __device__ __forceinline__ uint32_t s1 (int t) const { return pe ? uint32_t(t + 1) & (SE - 1) : (c2 >= 1u ? c2 - 1u : c2 + SE - 1u); }
    __device__ __forceinline__ uint32_t s2 (int t) const { return pe ? uint32_t(t + 2) & (SE - 1) : c2; }
    __device__ __forceinline__ uint32_t par2 (int t) const { return pe ? (uint32_t(t + 2) / SE) & 1u : cp2; }
    __device__ __forceinline__ uint32_t l (int t) const { return pl ? uint32_t(t) & (SL - 1) : cl; }
    __device__ __forceinline__ uint32_t parl (int t) const { return pl ? (uint32_t(t) / SL) & 1u : cpl; }
    __device__ __forceinline__ void advance ()
    {
        if (!pe) { if (++c2 == uint32_t(SE)) { c2 = 0u; cp2 ^= 1u; } }
        if (!pl) { if (++cl == uint32_t(SL)) { cl = 0u; cpl ^= 1u; } }
    }
};

struct Carry { double rhs, a, bxm, bxp, bym, byp; };     // coefficients of the black cell of the pair, read one step ahead

// ---------------------------------------------------------------------------------------------- one z step (compute threads)
// thread 0: arm the slot's mbarrier with the plane's byte count, then issue the plane's copies (descriptors d0 .. d1-1)
__device__ __forceinline__ void
produce (const Header* H, int d0, int d1, uint32_t bar, uint32_t stage_base, int q, uint32_t total_bytes)
{
#if defined(B200MG_G4_DIAG) && B200MG_G4_DIAG == 1        // timing diagnostics: no copies, the barrier completes at once
    mbar_arrive_expect_tx(bar, 0u);
#else
    mbar_arrive_expect_tx(bar, total_bytes);
    for (int d = d0; d < d1; ++d) {
        const CopyDesc c = H->d[d];
        if (c.bytes != 0u && q >= c.qmin && q <= c.qmax) { bulk_g2s(stage_base + c.soff, c.g + q * c.gstep, c.bytes, bar); }
    }
#endif
}

// the pass with two cell pairs per thread (gsrb_fused5.cu): rows of 65 .. 128 cells, P.txp == 32
int dispatch5 (bool abec, const FusedParams4& P, int nboxes, int tile_y, int early_stages, int late_stages, cudaStream_t s);

} } // namespace b200mg::fused

#endif
