/* amrex_b200_kernels.h -- kernel-level C ABI of the B200 MLMG path (libamrex_b200.so).
 *
 * One entry per device lambda of the reference's cell-centred MLMG path.
 * Every entry takes plain device pointers / POD descriptor tables and a cudaStream_t, launches hand-written
 * sm_100a kernels asynchronously on that stream and returns 0 or a cudaError_t value.  Nothing here
 * allocates, synchronises or falls back to the CPU.
 *
 * Descriptor tables (b200mg_fab etc.) live in DEVICE memory, one element per local box of a level, all
 * tables of one call indexed by the same local box number.  Layout of a fab: Fortran order,
 * element (i,j,k,n) at p[(i-lo[0]) + (j-lo[1])*jstride + (k-lo[2])*kstride + n*nstride], where lo/hi bound
 * the GROWN (valid + ghost) box -- the contract of Array4 (reference Src/Base/AMReX_Array4.H:85-137), except
 * that jstride may exceed the grown x-extent (rows are padded so that the first VALID cell of every row is
 * 128-byte aligned).
 */
#ifndef AMREX_B200_KERNELS_H_
#define AMREX_B200_KERNELS_H_

#ifdef __cplusplus
extern "C" {
#endif

#ifndef __CUDA_RUNTIME_H__
struct CUstream_st;
typedef struct CUstream_st* cudaStream_t;
#endif

typedef struct b200mg_fab {
    double* p;                 /* address of element (lo[0],lo[1],lo[2],0) */
    int lo[3], hi[3];          /* grown box, inclusive */
    long long jstride, kstride, nstride;   /* in elements */
} b200mg_fab;

typedef struct b200mg_ifab {
    int* p;
    int lo[3], hi[3];
    long long jstride, kstride, nstride;
} b200mg_ifab;

typedef struct b200mg_box { int lo[3], hi[3]; } b200mg_box;

/* A unit of work of a cell kernel: rows [j0, j0+B200MG_TILE_Y) x planes [k0, k0+nk) x all i of local box `box`
 * (clipped to the box); nk == 0 means B200MG_TILE_Z.  Tables of tiles are built once per level (amrex::LevelLayout). */
typedef struct b200mg_tile { int box, j0, k0, nk; } b200mg_tile;
#define B200MG_TILE_Y 4
#define B200MG_TILE_Z 4

/* boundary-condition work item: one face of one box (reference MLMGABCTag, AMReX_MLCellLinOp.H:724-760) */
typedef struct b200mg_bcface {
    int box;          /* local box number */
    int face;         /* Orientation value: dir + 3*side */
    int bctype;       /* 101 Dirichlet, 102 Neumann, 103 reflect-odd (AMReX_LO_BCTYPES.H:5-15) */
    int blen;         /* valid-box length in the face-normal direction */
    double bcloc;     /* distance of the BC point from the face */
} b200mg_bcface;

/* copy work item of FillBoundary / ParallelCopy (reference CopyComTag + Array4CopyTag, AMReX_FBI.H:53-70).
 * Copies the box [lo,hi] of the DESTINATION index space; source index = destination index + shift.
 * dst_fab / src_fab are local box numbers into the fab tables, or -1 when that side is the linear buffer,
 * in which case buf_offset is the element offset of this item's first value (x fastest, then y, z, comp). */
typedef struct b200mg_copytag {
    int lo[3], hi[3];
    int shift[3];
    int dst_fab, src_fab;
    int pad;
    long long buf_offset;
} b200mg_copytag;

/* Face link of a local box: the one local fab whose valid cells lie behind the whole face (fab < 0: none - physical boundary,
 * coarse/fine boundary, another rank, or several neighbours), and the index shift from this box's ghost cell to the cell of
 * that fab that holds its value (non-zero across a periodic boundary).  Table layout [box * 6 + face], faces 0,1,2 = x,y,z low,
 * 3,4,5 = high. */
typedef struct b200mg_facelink {
    int fab;
    int shift[3];
} b200mg_facelink;

/* ---- smoother: one red or black sweep (K1 abec_gsrb AMReX_MLABecLap_3D_K.H:210-264,
 *      K2 mlpoisson_gsrb AMReX_MLPoisson_3D_K.H:155-196).  f / m: tables of 6*nboxes slabs, [box*6+face]. */
int b200mg_gsrb_abec(int ntiles, const b200mg_tile* tiles, const b200mg_box* vbox,
                     const b200mg_fab* phi, const b200mg_fab* rhs, const b200mg_fab* a,
                     const b200mg_fab* bx, const b200mg_fab* by, const b200mg_fab* bz,
                     const b200mg_fab* f, const b200mg_ifab* m,
                     double alpha, double dhx, double dhy, double dhz, int redblack, cudaStream_t s);
int b200mg_gsrb_poisson(int ntiles, const b200mg_tile* tiles, const b200mg_box* vbox,
                        const b200mg_fab* phi, const b200mg_fab* rhs,
                        const b200mg_fab* f, const b200mg_ifab* m,
                        double dhx, double dhy, double dhz, int redblack, cudaStream_t s);
/* Same colour sweeps for levels whose boxes all have an even x extent <= 128: one thread per cell PAIR streaming along z
 * (the fast path; results are bit-identical to the entries above). */
int b200mg_gsrb_abec_pairs(int ntiles, const b200mg_tile* tiles, const b200mg_box* vbox,
                           const b200mg_fab* phi, const b200mg_fab* rhs, const b200mg_fab* a,
                           const b200mg_fab* bx, const b200mg_fab* by, const b200mg_fab* bz,
                           const b200mg_fab* f, const b200mg_ifab* m,
                           double alpha, double dhx, double dhy, double dhz, int redblack, cudaStream_t s);
int b200mg_gsrb_poisson_pairs(int ntiles, const b200mg_tile* tiles, const b200mg_box* vbox,
                              const b200mg_fab* phi, const b200mg_fab* rhs,
                              const b200mg_fab* f, const b200mg_ifab* m,
                              double dhx, double dhy, double dhz, int redblack, cudaStream_t s);
/* Lean variants of the pair sweeps: additionally require nx >= 4 and ny >= 2 for every box (identical results). */
int b200mg_gsrb_abec_pairs_lean(int ntiles, const b200mg_tile* tiles, const b200mg_box* vbox,
                                const b200mg_fab* phi, const b200mg_fab* rhs, const b200mg_fab* a,
                                const b200mg_fab* bx, const b200mg_fab* by, const b200mg_fab* bz,
                                const b200mg_fab* f, const b200mg_ifab* m,
                                double alpha, double dhx, double dhy, double dhz, int redblack, cudaStream_t s);
int b200mg_gsrb_poisson_pairs_lean(int ntiles, const b200mg_tile* tiles, const b200mg_box* vbox,
                                   const b200mg_fab* phi, const b200mg_fab* rhs,
                                   const b200mg_fab* f, const b200mg_ifab* m,
                                   double dhx, double dhy, double dhz, int redblack, cudaStream_t s);
/* resident CTAs per SM the lean sweeps are compiled for: 4 (64 registers, default) or 3 (80 registers); <= 0 selects the generic pair sweep */
void b200mg_set_gsrb_lean_occupancy(int min_blocks);
/* Fused red+black pass (one sweep over memory per smooth): red update of every valid cell and black update of the
 * cells that do not touch the box surface, out of place (phi_in -> phi_out).  The black surface shell is finished by
 * b200mg_gsrb_shell_* after the halo refresh.  Same arithmetic and update order as two colour sweeps.
 * (Two earlier generations of this pass - plain global loads with L2 prefetch - were measured at 36-44 % of the HBM
 * peak in round 1 and have been removed; levels the staged pass cannot take run the colour sweeps.) */
/* The fused pass: HOST descriptor tables (one entry per local box; f / m: [box*6+face]) travel to the device as kernel
 * parameters, <= 64 boxes per launch; abec == 0: Poisson.  Every operand plane is staged
 * into a shared-memory ring by 1-D bulk async copies (TMA engine, mbarrier transaction counts) several planes ahead of
 * its use by a dedicated producer warp; compute warps read shared memory only.  Whole-z CTAs of (all x) x tile_y rows.
 * Requirements: even x extent, 4 <= nx <= 128, ny >= 2, rows 16-byte aligned at the first valid cell, even strides,
 * phi rows readable on [lo-2, hi+2], bx rows nx+2 doubles long (all guaranteed by the FabArray allocator);
 * cudaErrorInvalidValue otherwise (the caller falls back to the colour sweeps).
 * phi_zero != 0: the input is identically zero INCLUDING its ghost cells (first smooth after cor.setVal(0),
 * AMReX_MLMG.H:1318-1326): h_phi_in is not read at all - the shared-memory planes are zero-filled instead - so the
 * caller may skip the setVal; the output is the same bits as with a zeroed input. */
int b200mg_gsrb4(int abec, int nboxes, const b200mg_box* h_vbox,
                 const b200mg_fab* h_phi_in, const b200mg_fab* h_phi_out, const b200mg_fab* h_rhs, const b200mg_fab* h_a,
                 const b200mg_fab* h_bx, const b200mg_fab* h_by, const b200mg_fab* h_bz,
                 const b200mg_fab* h_f, const b200mg_ifab* h_m,
                 double alpha, double dhx, double dhy, double dhz, int phi_zero, cudaStream_t s);
/* the same pass over the nboxes local boxes listed in ids (host array; NULL: boxes 0 .. nboxes-1); the tables stay indexed
 * by local box.  Used to run the boxes whose halo is complete while the remote part of a FillBoundary is in flight. */
int b200mg_gsrb4_subset(int abec, int nboxes, const int* ids, const b200mg_box* h_vbox,
                        const b200mg_fab* h_phi_in, const b200mg_fab* h_phi_out, const b200mg_fab* h_rhs, const b200mg_fab* h_a,
                        const b200mg_fab* h_bx, const b200mg_fab* h_by, const b200mg_fab* h_bz,
                        const b200mg_fab* h_f, const b200mg_ifab* h_m,
                        double alpha, double dhx, double dhy, double dhz, int phi_zero, cudaStream_t s);
/* launch plan of b200mg_gsrb4: rows per CTA tile and ring depths; (8,4,2) default, (8,4,3), (6,5,3), (6,4,2), (4,4,4) */
int b200mg_set_gsrb4_plan(int tile_y, int early_stages, int late_stages);
/* cell pairs per thread of b200mg_gsrb4 on rows of more than 64 cells: 0 = chosen by the size of the launch (two when it runs
 * more than three waves of CTAs), 1 = one pair per thread, 2 = two (one warp per 128-cell row + a producer warp) */
void b200mg_set_gsrb4_sync(int pairs);
/* black (redblack=1) or red sweep restricted to the 1-cell surface shell of every box; max_face_cells = cells of the
 * largest box face of the level (sizes the grid: one thread per swept cell) */
int b200mg_gsrb_shell_abec(int nboxes, const b200mg_box* vbox,
                           const b200mg_fab* phi, const b200mg_fab* rhs, const b200mg_fab* a,
                           const b200mg_fab* bx, const b200mg_fab* by, const b200mg_fab* bz,
                           const b200mg_fab* f, const b200mg_ifab* m,
                           double alpha, double dhx, double dhy, double dhz, int redblack, int max_face_cells, cudaStream_t s);
int b200mg_gsrb_shell_poisson(int nboxes, const b200mg_box* vbox,
                              const b200mg_fab* phi, const b200mg_fab* rhs,
                              const b200mg_fab* f, const b200mg_ifab* m,
                              double dhx, double dhy, double dhz, int redblack, int max_face_cells, cudaStream_t s);
/* the same with face links: a shell cell reads the value beyond a linked face from the neighbouring fab's valid cell instead
 * of its own ghost cell (so the halo exchange ahead of the shell sweep has nothing to copy between local boxes), and with
 * push != 0 it stores its new value into that fab's ghost cell as well (so the exchange ahead of the NEXT sweep of the other
 * colour has nothing to copy either).  links == NULL: the plain shell sweep. */
int b200mg_gsrb_shell_abec_linked(int nboxes, const b200mg_box* vbox,
                                  const b200mg_fab* phi, const b200mg_fab* rhs, const b200mg_fab* a,
                                  const b200mg_fab* bx, const b200mg_fab* by, const b200mg_fab* bz,
                                  const b200mg_fab* f, const b200mg_ifab* m,
                                  double alpha, double dhx, double dhy, double dhz, int redblack, int max_face_cells,
                                  const b200mg_facelink* links, int push, cudaStream_t s);
int b200mg_gsrb_shell_poisson_linked(int nboxes, const b200mg_box* vbox,
                                     const b200mg_fab* phi, const b200mg_fab* rhs,
                                     const b200mg_fab* f, const b200mg_ifab* m,
                                     double dhx, double dhy, double dhz, int redblack, int max_face_cells,
                                     const b200mg_facelink* links, int push, cudaStream_t s);

/* ---- damped Jacobi sweep (abec_jacobi AMReX_MLABecLap_3D_K.H:332-375, mlpoisson_jacobi AMReX_MLPoisson_3D_K.H:250-281):
 *      phi_out = phi_in + 2/3 * (rhs - L(phi_in)) / (gamma - face terms) on every valid cell, out of place, with L(phi_in)
 *      evaluated in the same pass (the reference stores it first).  dh*: beta/h^2 of the smoother, ad*: beta*dxinv^2 of
 *      the apply (Poisson: both are dxinv^2).  The PoisArgs mirror b200mg_gsrb_poisson. */
int b200mg_jacobi_abec(int ntiles, const b200mg_tile* tiles, const b200mg_box* vbox,
                       const b200mg_fab* phi_out, const b200mg_fab* phi_in, const b200mg_fab* rhs, const b200mg_fab* a,
                       const b200mg_fab* bx, const b200mg_fab* by, const b200mg_fab* bz,
                       const b200mg_fab* f, const b200mg_ifab* m, double alpha, double dhx, double dhy, double dhz,
                       double adx, double ady, double adz, cudaStream_t s);
int b200mg_jacobi_poisson(int ntiles, const b200mg_tile* tiles, const b200mg_box* vbox,
                          const b200mg_fab* phi_out, const b200mg_fab* phi_in, const b200mg_fab* rhs,
                          const b200mg_fab* f, const b200mg_ifab* m, double dhx, double dhy, double dhz, cudaStream_t s);

/* ---- operator apply / residual (K4 mlabeclap_adotx AMReX_MLABecLap_3D_K.H:9-28, K5 mlpoisson_adotx
 *      AMReX_MLPoisson_3D_K.H:9-16).  If rhs != NULL writes y = rhs - L(x)  (== Xpay(y,-1,rhs),
 *      AMReX_MLCellLinOp.H:1234), else y = L(x).  dx*: beta*dxinv^2 (abec) or dxinv^2 (poisson). */
int b200mg_adotx_abec(int ntiles, const b200mg_tile* tiles, const b200mg_box* vbox,
                      const b200mg_fab* y, const b200mg_fab* x, const b200mg_fab* rhs, const b200mg_fab* a,
                      const b200mg_fab* bx, const b200mg_fab* by, const b200mg_fab* bz,
                      double alpha, double dhx, double dhy, double dhz, cudaStream_t s);
int b200mg_adotx_poisson(int ntiles, const b200mg_tile* tiles, const b200mg_box* vbox,
                         const b200mg_fab* y, const b200mg_fab* x, const b200mg_fab* rhs,
                         double dhx, double dhy, double dhz, cudaStream_t s);
/* Same operators for levels whose boxes all have an even x extent <= 128: one thread per cell PAIR, streaming along z with
 * 16-byte loads (the fast path; results are bit-identical to the entries above).
 * norminf != NULL (device pointer): *norminf = max |y| over the launch, computed in the same pass (residual + ResNormInf,
 * AMReX_MLMG.H:1808-1812, without re-reading y); NULL: not computed. */
int b200mg_adotx_abec_pairs(int ntiles, const b200mg_tile* tiles, const b200mg_box* vbox,
                            const b200mg_fab* y, const b200mg_fab* x, const b200mg_fab* rhs, const b200mg_fab* a,
                            const b200mg_fab* bx, const b200mg_fab* by, const b200mg_fab* bz,
                            double alpha, double dhx, double dhy, double dhz, double* norminf, cudaStream_t s);
int b200mg_adotx_poisson_pairs(int ntiles, const b200mg_tile* tiles, const b200mg_box* vbox,
                               const b200mg_fab* y, const b200mg_fab* x, const b200mg_fab* rhs,
                               double dhx, double dhy, double dhz, double* norminf, cudaStream_t s);
/* residual fused with its restriction: crse = average_down(rhs - L(x)), the fine residual is not stored (MLMGT::mgVcycle
 * AMReX_MLMG.H:1332-1345: computeResOfCorrection + restriction; amrex_avgdown AMReX_MultiFabUtil_3D_C.H:381-394).
 * tiles / vbox: the FINE level's (ng = 0); crse: fab table on the coarsened fine layout (same local box order).  Requires
 * the pair layout of b200mg_adotx_*_pairs, even box corners / extents and an even tile depth. */
int b200mg_residual_restrict_abec(int ntiles, const b200mg_tile* tiles, const b200mg_box* vbox,
                                  const b200mg_fab* crse, const b200mg_fab* x, const b200mg_fab* rhs, const b200mg_fab* a,
                                  const b200mg_fab* bx, const b200mg_fab* by, const b200mg_fab* bz,
                                  double alpha, double dhx, double dhy, double dhz, cudaStream_t s);
int b200mg_residual_restrict_poisson(int ntiles, const b200mg_tile* tiles, const b200mg_box* vbox,
                                     const b200mg_fab* crse, const b200mg_fab* x, const b200mg_fab* rhs,
                                     double dhx, double dhy, double dhz, cudaStream_t s);
/* K13 mlabeclap_normalize AMReX_MLABecLap_3D_K.H:60-75 */
int b200mg_normalize_abec(int ntiles, const b200mg_tile* tiles, const b200mg_box* vbox,
                          const b200mg_fab* x, const b200mg_fab* a,
                          const b200mg_fab* bx, const b200mg_fab* by, const b200mg_fab* bz,
                          double alpha, double dhx, double dhy, double dhz, cudaStream_t s);

/* ---- boundary conditions (K9 mllinop_apply_bc_* AMReX_MLLinOp_K.H:14-327; K11 comp_interp_coef0 :329-571).
 *      m, f, bcval: [box*6+face] tables.  bcval may be NULL (homogeneous).  max_face_cells: cells of the largest box face
 *      (sizes the grid: 2 ghost cells per thread; <= 0 if unknown). */
int b200mg_apply_bc(int nfaces, const b200mg_bcface* faces, const b200mg_box* vbox,
                    const b200mg_fab* phi, const b200mg_ifab* m, const b200mg_fab* bcval,
                    int maxorder, double dxinv0, double dxinv1, double dxinv2, int inhomog, int max_face_cells, cudaStream_t s);
/* Inhomogeneous Neumann data, stored as d(phi)/dn in the ghost cells of the BC values (bcval, [box*6+face]).
 *   mode 0 (mllinop_apply_innu_*, AMReX_MLLinOp_K.H:930-1075): rhs of the cell inside a flagged domain face
 *          -= (low) / += (high) fac[d]*b*bcval, fac = beta*dxinv; out3 = {rhs, rhs, rhs};
 *   mode 1 (MLCellABecLapT::addInhomogNeumannFlux, AMReX_MLCellABecLap.H:517-620): the domain-face value of the
 *          face-centred array out3[d] := fac[d]*b*bcval.
 * b3: face coefficients per direction (NULL or NULL entries: b = 1); on_face[6]: which domain faces (orientation values)
 * carry inhomogeneous Neumann data; only ghost cells with mask == 2 (outside the domain) take part.  Mode 0 updates are
 * plain read-modify-writes: flag ONE face orientation per launch (edge / corner cells belong to several faces). */
int b200mg_apply_innu(int nfaces, const b200mg_bcface* faces, const b200mg_box* vbox,
                      const b200mg_fab* const out3[3], const b200mg_fab* const b3[3],
                      const b200mg_ifab* m, const b200mg_fab* bcval, const double fac[3], const int on_face[6], int mode, cudaStream_t s);
/* Robin boundary condition a*phi + b*dphi/dn = f on the flagged domain faces (data in the ghost cells of the slabs ra / rb / rf,
 * [box*6+face]): mode 0 adds the diagonal term to the a coefficient (MLABecLaplacian applyRobinBCTermsCoeffs,
 * AMReX_MLABecLaplacian.H:459-600; fac = (b_scalar/a_scalar)*dxinv^2), mode 1 the right-hand-side term
 * (AMReX_MLCellABecLap.H:448-510; fac = b_scalar*dxinv^2), mode 2 writes the domain-face flux (:579-612; phi: the solution).
 * Modes 0 / 1: flag ONE face orientation per launch. */
int b200mg_robin(int nfaces, const b200mg_bcface* faces, const b200mg_box* vbox,
                 const b200mg_fab* const out3[3], const b200mg_fab* const b3[3], const b200mg_fab* phi,
                 const b200mg_ifab* m, const b200mg_fab* ra, const b200mg_fab* rb, const b200mg_fab* rf,
                 const double fac[3], const double dxinv[3], const int on_face[6], int mode, cudaStream_t s);
int b200mg_comp_interp_coef0(int nfaces, const b200mg_bcface* faces, const b200mg_box* vbox,
                             const b200mg_fab* f, const b200mg_ifab* m,
                             int maxorder, double dxinv0, double dxinv1, double dxinv2, cudaStream_t s);

/* ---- grid transfer (K7 amrex_avgdown AMReX_MultiFabUtil_3D_C.H:377-395, amrex_avgdown_faces :173-217;
 *      K8 prolongation AMReX_MLCellLinOp.H:970-977; K8b mlmg_lin_cc_interp_r2 AMReX_MLMG_3D_K.H:9-39).
 *      tiles/vbox describe the COARSE boxes for restriction and the FINE boxes for prolongation. */
int b200mg_restrict_cc(int ntiles, const b200mg_tile* tiles, const b200mg_box* cbox,
                       const b200mg_fab* crse, const b200mg_fab* fine, int ratio, cudaStream_t s);
int b200mg_restrict_faces(int ntiles, const b200mg_tile* tiles, const b200mg_box* cbox,
                          const b200mg_fab* crse, const b200mg_fab* fine, int dir, int ratio, cudaStream_t s);
int b200mg_prolong_add(int ntiles, const b200mg_tile* tiles, const b200mg_box* fbox,
                       const b200mg_fab* fine, const b200mg_fab* crse, cudaStream_t s);
int b200mg_interp_cc_r2(int ntiles, const b200mg_tile* tiles, const b200mg_box* fbox,
                        const b200mg_fab* fine, const b200mg_fab* crse, int add, cudaStream_t s);

/* ---- coarse/fine coupling of the multi-level solve.
 *      interp_bndry_o3: interpbndrydata_{x,y,z}_o3 (Src/Boundary/AMReX_InterpBndryData_3D_K.H:22-119) for every listed
 *      (box, face): bdry / crse / mask are [box*6+face] tables (bdry: 1-cell slab outside the fine face; crse: boundary
 *      register of the coarsened box, tangential extent 2; mask: 2 cells outside, tangential extent 5, 1 = not covered).
 *      reflux_crse / reflux_fine: YAFluxRegister CrseAdd / FineAdd (Src/Boundary/AMReX_YAFluxRegister_3D_K.H:11-199)
 *      with the operator's face fluxes -fac*b*(sol(i)-sol(i-1)) (AMReX_MLABecLap_3D_K.H:79-95; b tables NULL: b = 1)
 *      evaluated in place.  fac* = b_scalar*dxinv; crse: dtd* = dt/dx_crse; fine: dtd* = dt/(dx_fine*ratio^3).
 *      reflux_fine: one entry per coarse/fine patch fab: cfbox = coarsened fine box it surrounds, fine_index = local
 *      index of that fine box; mask (may be NULL) multiplies the result (periodic self-overlap). */
int b200mg_interp_bndry_o3(int nfaces, const b200mg_bcface* faces, const b200mg_box* vbox,
                           const b200mg_fab* bdry, const b200mg_fab* crse, const b200mg_ifab* mask, int ratio, cudaStream_t s);
int b200mg_reflux_crse(int ntiles, const b200mg_tile* tiles, const b200mg_box* vbox,
                       const b200mg_fab* crse_data, const b200mg_ifab* flag, const b200mg_fab* sol,
                       const b200mg_fab* bx, const b200mg_fab* by, const b200mg_fab* bz,
                       double facx, double facy, double facz, double dtdx, double dtdy, double dtdz, cudaStream_t s);
int b200mg_reflux_fine(int npatches, const b200mg_fab* cfpatch, const b200mg_box* cfbox, const int* fine_index,
                       const b200mg_fab* mask, const b200mg_fab* fine_sol,
                       const b200mg_fab* bx, const b200mg_fab* by, const b200mg_fab* bz,
                       double facx, double facy, double facz, double dtdx, double dtdy, double dtdz, int ratio, cudaStream_t s);

/* face-centred fluxes / gradients of a cell-centred solution (post-solve API: MLMG::getFluxes / getGradSolution).
 * fbox / tiles: FACE boxes of direction dir; sol needs one filled ghost cell.  mode 0: fac*(s-s^-) (compGrad,
 * AMReX_MLCellLinOp.H:1421-1436); 1: -fac*b*(s-s^-) (mlabeclap_flux_*, AMReX_MLABecLap_3D_K.H:79-135); 2: fac*(s-s^-)
 * (mlpoisson_flux_*, AMReX_MLPoisson_3D_K.H:36-98); modes 1, 2 multiply by post afterwards when post != 1. */
int b200mg_face_flux(int ntiles, const b200mg_tile* tiles, const b200mg_box* fbox,
                     const b200mg_fab* out, const b200mg_fab* sol, const b200mg_fab* b,
                     double fac, double post, int dir, int mode, cudaStream_t s);

/* arithmetic cell-centre -> face average (amrex::average_cellcenter_to_face, AMReX_MultiFabUtil_3D_C.H:81-89) */
int b200mg_cc_to_face(int ntiles, const b200mg_tile* tiles, const b200mg_box* fbox,
                      const b200mg_fab* face, const b200mg_fab* cc, int dir, cudaStream_t s);

/* ---- vector operations on the valid box grown by ng (K6; AMReX_FabArray.H:179-260,2918-3006) */
int b200mg_setval(int ntiles, const b200mg_tile* tiles, const b200mg_box* vbox, const b200mg_fab* y,
                  double v, int ng, cudaStream_t s);
int b200mg_copy(int ntiles, const b200mg_tile* tiles, const b200mg_box* vbox, const b200mg_fab* y,
                const b200mg_fab* x, int ng, cudaStream_t s);
/* y = a*x + b*y   (a=1,b=1: LocalAdd; b=1: Saxpy; a=1: Xpay; a=0: scale) */
int b200mg_lincomb(int ntiles, const b200mg_tile* tiles, const b200mg_box* vbox, const b200mg_fab* y,
                   double a, const b200mg_fab* x, double b, int ng, cudaStream_t s);
int b200mg_plus(int ntiles, const b200mg_tile* tiles, const b200mg_box* vbox, const b200mg_fab* y,
                double v, int ng, cudaStream_t s);
/* y *= x / y /= x on the cells grown by ng (MultiFab::Multiply / Divide, Src/Base/AMReX_MultiFab.H) */
int b200mg_multiply(int ntiles, const b200mg_tile* tiles, const b200mg_box* vbox, const b200mg_fab* y, const b200mg_fab* x, int ng, cudaStream_t s);
int b200mg_divide(int ntiles, const b200mg_tile* tiles, const b200mg_box* vbox, const b200mg_fab* y, const b200mg_fab* x, int ng, cudaStream_t s);
/* signed minimum (want_max == 0) / maximum over the cells grown by ng (FabArray min / max); tiles built for that ng */
int b200mg_minmax(int ntiles, const b200mg_tile* tiles, const b200mg_box* vbox, const b200mg_fab* x, int want_max, int ng,
                  double* result, double* scratch, cudaStream_t s);
/* ghost cells only (setBndry) */
int b200mg_setbndry(int ntiles, const b200mg_tile* tiles, const b200mg_box* vbox, const b200mg_fab* y,
                    double v, int ng, cudaStream_t s);

/* ---- reductions over valid cells (K12).  result: DEVICE double[ntiles>0 ? 1 : 1]; the kernel leaves the
 *      final value in result[0] (two-stage: warp shuffles + last-block finish), no host sync.
 *      mask (may be NULL): only cells with mask != 0 count (AMReX_FabArray.H:3577-3625). */
int b200mg_norminf(int ntiles, const b200mg_tile* tiles, const b200mg_box* vbox, const b200mg_fab* x,
                   const b200mg_ifab* mask, double* result, double* scratch, cudaStream_t s);
int b200mg_dot(int ntiles, const b200mg_tile* tiles, const b200mg_box* vbox, const b200mg_fab* x,
               const b200mg_fab* y, double* result, double* scratch, cudaStream_t s);
int b200mg_sum(int ntiles, const b200mg_tile* tiles, const b200mg_box* vbox, const b200mg_fab* x,
               double* result, double* scratch, cudaStream_t s);
/* sum of |x| over the valid cells (FabArray::norm1, AMReX_FabArray.H) */
int b200mg_asum(int ntiles, const b200mg_tile* tiles, const b200mg_box* vbox, const b200mg_fab* x,
                double* result, double* scratch, cudaStream_t s);
/* scratch must hold b200mg_reduce_scratch_doubles(ntiles) doubles */
long long b200mg_reduce_scratch_doubles(int ntiles);

/* ---- whole BiCGStab bottom solve in one single-CTA kernel (MLCGSolverT::solve_bicgstab with a zeroed initial vector,
 *      AMReX_MLCGSolver.H:98-273): for a bottom MG level that is ONE box (<= 32^3 cells) covering the domain.
 *      h_*: HOST descriptors of the box's fabs (sol, r, p with one ghost cell; sol zeroed by the caller; r, p, v, t, rh
 *      scratch); abec == 0: Poisson (a, bx, by, bz ignored; dh* = dxinv^2), else dh* = beta*dxinv^2.  h_faces: the box's
 *      physical faces (<= 6, box == 0), h_mask: HOST table of the box's 6 mask slabs, periodic[3]: directions in which the
 *      box wraps onto itself (NULL: none).  d_out (device, 4 doubles):
 *      return code of solve_bicgstab, iteration count, final and initial max-norm of the residual. */
int b200mg_bottom_bicgstab(int abec, const b200mg_box* h_vbox,
                           const b200mg_fab* h_sol, const b200mg_fab* h_rhs, const b200mg_fab* h_r, const b200mg_fab* h_p,
                           const b200mg_fab* h_v, const b200mg_fab* h_t, const b200mg_fab* h_rh,
                           const b200mg_fab* h_a, const b200mg_fab* h_bx, const b200mg_fab* h_by, const b200mg_fab* h_bz,
                           double alpha, double dhx, double dhy, double dhz,
                           int nfaces, const b200mg_bcface* h_faces, const b200mg_ifab* h_mask, const int* periodic, int maxorder,
                           double dxinv0, double dxinv1, double dxinv2, double eps_rel, double eps_abs, int maxiter,
                           double* d_out, cudaStream_t s);

/* ---- the coarse leg of a V-cycle in ONE cooperative kernel (MLMGT::mgVcycle AMReX_MLMG.H:1308-1415 from the first MG
 *      level that is - or has been merged into - a single box covering the domain down to the bottom solve,
 *      MLMGT::bottomSolve :1460-1576, BiCGStab as b200mg_bottom_bicgstab or nuf smooths, and back up).  lev[0] is the top
 *      level of the leg, lev[nlev-1] the bottom (<= 32^3 cells when solved by BiCGStab); every level is ONE box, coarsened
 *      by 2 from the level above, whose faces are physical Dirichlet / Neumann / reflect-odd boundaries or wrap onto the
 *      box itself (periodic).  Levels of more than narrow_cells cells are worked on by the whole grid (grid barrier between
 *      phases), smaller ones by CTA 0 alone (block barriers).  res[0] holds the right-hand side on entry, cor[0] the
 *      correction on exit; all other fields of the leg are scratch.  Same bits as the launch-per-operation schedule
 *      (shared per-cell code, same operation order). */
#define B200MG_LEG_MAX_LEVELS 8
typedef struct b200mg_leg_level {
    b200mg_box vb;                  /* the level's box == its domain */
    b200mg_fab cor, res, rescor;    /* cor: one ghost cell */
    b200mg_fab a, bx, by, bz;       /* MLABecLaplacian coefficients (unused for Poisson) */
    b200mg_fab f[6];                /* relaxation-coefficient slabs (m_undrrelxr), [face] */
    b200mg_ifab m[6];               /* mask slabs (m_maskvals), [face] */
    int nfaces;                     /* faces with uncovered ghost cells = physical boundaries */
    int periodic[3];
    b200mg_bcface faces[6];
    double dxi[3];                  /* inverse cell size */
    double dh[3];                   /* smoother scaling: beta/h^2 (Poisson: dxinv^2) */
    double adh[3];                  /* apply scaling: beta*dxinv^2 (Poisson: dxinv^2) */
} b200mg_leg_level;
typedef struct b200mg_leg_args {
    int nlev, maxorder, nu1, nu2, nuf, nub;
    int bottom_mode;                /* 0: BiCGStab, 1: smoother */
    int singular;                   /* bottom right-hand side is made solvable on the copy bb first */
    int maxiter, narrow_cells;
    double alpha, volinv, eps_rel, eps_abs;
    b200mg_fab r, p, v, t, rh, bb;  /* bottom-level scratch: r, p with one ghost cell */
    unsigned long long* stamps;     /* NULL, or device array of B200MG_LEG_MAX_STAMPS words: CTA 0 records (id << 48 | SM clock) after
                                       every phase (id = level * 16 + phase), entry 0 = number of records (tuning aid) */
    b200mg_leg_level lev[B200MG_LEG_MAX_LEVELS];
} b200mg_leg_args;
#define B200MG_LEG_MAX_STAMPS 2048
/* d_args: the arguments in DEVICE memory; d_out (device, 2 doubles, may be NULL): return code and iteration count of the
 * BiCGStab bottom solve; ctas: CTAs of the cooperative grid (512 threads each; clamped to one per SM). */
int b200mg_coarse_leg(int abec, const b200mg_leg_args* d_args, double* d_out, int ctas, cudaStream_t s);

/* ---- batched Krylov vector kernels (GMRES Gram-Schmidt: the dotProduct / increment loops of
 *      GMRES::gram_schmidt_orthogonalization, AMReX_GMRES.H:322-348).  v: HOST array of nv <= B200MG_KRYLOV_GROUP device
 *      fab tables living on the same layout as x / w.
 *      multi_dot : result[n] = sum_valid x * v[n]  (device array of nv doubles; deterministic two-stage reduction,
 *                  scratch >= b200mg_multi_dot_scratch_doubles(ntiles) doubles), x is read once.
 *      multi_axpy: w = a[nv-1]*v[nv-1] + (... (a[0]*v[0] + w)), one pass, the roundings of nv successive Saxpy calls. */
#define B200MG_KRYLOV_GROUP 8
long long b200mg_multi_dot_scratch_doubles(int ntiles);
int b200mg_multi_dot(int ntiles, const b200mg_tile* tiles, const b200mg_box* vbox, const b200mg_fab* x,
                     int nv, const b200mg_fab* const* v, double* result, double* scratch, cudaStream_t s);
int b200mg_multi_axpy(int ntiles, const b200mg_tile* tiles, const b200mg_box* vbox, const b200mg_fab* w,
                      int nv, const b200mg_fab* const* v, const double* a, cudaStream_t s);

/* ---- halo / redistribution copies (K10 fab_to_fab, pack, unpack; AMReX_FBI.H:53-70,272-328,729-893).
 *      op: 0 = copy, 1 = add.  buf: linear staging buffer (may be NULL when no tag uses it).
 *      max_pts: points of the largest tag (sizes the grid: ~2 points per thread; <= 0 if unknown). */
int b200mg_copy_tags(int ntags, const b200mg_copytag* tags, const b200mg_fab* dst, const b200mg_fab* src,
                     double* buf, int ncomp, int scomp, int dcomp, int op, int max_pts, cudaStream_t s);
/* the same for one colour of the red-black lattice: only destination cells with (i+j+k) & 1 == parity (parity < 0: all).  A
 * colour sweep reads ghost cells of the other colour only, so the exchange ahead of it moves half the cells - which halves
 * the sector traffic of the x faces, where every 8-byte value sits in a sector of its own. */
int b200mg_copy_tags_colour(int ntags, const b200mg_copytag* tags, const b200mg_fab* dst, const b200mg_fab* src,
                            double* buf, int ncomp, int scomp, int dcomp, int op, int max_pts, int parity, cudaStream_t s);

/* library identification */
const char* b200mg_version(void);

#ifdef __cplusplus
}
#endif
#endif
