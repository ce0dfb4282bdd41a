/* amrex_b200_fi.h -- object-level C ABI of libamrex_b200.so: the drop-in boundary of the MLMG path.
 *
 * The amrex_fi_* entries have the names, argument order and meaning of the reference's own C ABI
 * (Src/F_Interfaces, the layer its Fortran module binds): each declaration cites the reference definition it
 * replaces.  `T**` here is ABI-identical to the reference's `T*&`.  Objects are opaque heap pointers created
 * and destroyed by new_/delete_ pairs; ownership rules are the reference's (MLMG keeps a
 * reference to the linop; the linop copies coefficients and BC values).
 *
 * The amrex_b200_* entries are what a process-per-GPU launcher needs and the reference gets from MPI / the
 * host: runtime init, NCCL bootstrap, host<->device field transfer, results of the last solve, and the
 * host-only metadata queries used for bit-exact parity checks.
 *
 * Error convention: the reference aborts the process (amrex::Abort).  Here every failure (assertion, CUDA
 * error, non-convergence) is caught at this boundary: the call returns, amrex_b200_last_error() returns a
 * non-NULL message until amrex_b200_clear_error(), and value-returning calls return NaN / -1 / NULL.
 * There is no CPU fallback: any call that needs the device fails with such an error when no GPU is present.
 */
#ifndef AMREX_B200_FI_H_
#define AMREX_B200_FI_H_

#ifdef __cplusplus
namespace amrex { template <class T> class FabArray; }
extern "C" {
#define B200_OPAQUE(T) namespace amrex { class T; } typedef amrex::T T
B200_OPAQUE(BoxArray); B200_OPAQUE(DistributionMapping); B200_OPAQUE(Geometry); B200_OPAQUE(MultiFab);
B200_OPAQUE(MLLinOp); B200_OPAQUE(MLMG); B200_OPAQUE(GMRESMLMG); B200_OPAQUE(MFIter);
typedef amrex::FabArray<int> iMultiFab;
#undef B200_OPAQUE
#else
typedef struct BoxArray BoxArray; typedef struct DistributionMapping DistributionMapping;
typedef struct Geometry Geometry; typedef struct MultiFab MultiFab;
typedef struct MLLinOp MLLinOp; typedef struct MLMG MLMG; typedef struct GMRESMLMG GMRESMLMG;
typedef struct MFIter MFIter; typedef struct iMultiFab iMultiFab;
#endif
typedef double Real;

/* ---- runtime (reference: amrex::Initialize/Finalize Src/Base/AMReX.cpp:332,725; ParallelDescriptor) ---- */
int  amrex_b200_init(int device_id);                 /* select GPU, create streams; 0 on success */
void amrex_b200_finalize(void);
int  amrex_b200_initialized(void);
int  amrex_b200_nccl_unique_id_bytes(void);
int  amrex_b200_nccl_get_unique_id(void* out);       /* rank 0 */
int  amrex_b200_comm_init(int rank, int nranks, const void* unique_id);   /* all ranks; nranks==1: no NCCL */
void amrex_b200_comm_finalize(void);
int  amrex_b200_myproc(void);
int  amrex_b200_nprocs(void);
const char* amrex_b200_last_error(void);
void amrex_b200_clear_error(void);
void amrex_b200_synchronize(void);
long long amrex_b200_launch_count(void);             /* kernels launched by this library since the last reset */
void amrex_b200_reset_launch_count(void);
void* amrex_b200_stream(void);
/* per-kernel device timing (CUDA events around every launch); report: lines "kernel scope launches total_ms min_ms max_ms",
 * scope = amrlev*100 + mglev of the level the launch worked on (-1: none).  Returns the report length; call with
 * buf == NULL to size the buffer (the report is then kept until read). */
void amrex_b200_profile_enable(int on);
int  amrex_b200_profile_report(char* buf, int capacity);                       /* the cudaStream_t every kernel is launched on */

/* ---- Geometry (Src/F_Interfaces/Base/AMReX_geometry_fi.cpp:7-40) ---- */
void amrex_b200_geometry_setup(const Real problo[3], const Real probhi[3], const int is_periodic[3]); /* Geometry::Setup */
void amrex_fi_new_geometry(Geometry** geom, int lo[3], int hi[3]);
void amrex_fi_delete_geometry(Geometry* geom);
void amrex_fi_geometry_get_intdomain(const Geometry* geom, int lo[3], int hi[3]);

/* ---- BoxArray (AMReX_boxarray_fi.cpp:10-93) ---- */
void amrex_fi_new_boxarray(BoxArray** ba, int lo[3], int hi[3]);
void amrex_fi_new_boxarray_from_bxfarr(BoxArray** ba, const int* bxs, const int nsides, const int ndims, const int nbxs);
void amrex_fi_delete_boxarray(BoxArray* ba);
void amrex_fi_clone_boxarray(BoxArray** bao, const BoxArray* bai);
void amrex_fi_boxarray_maxsize(BoxArray* ba, int sz[]);
long long amrex_fi_boxarray_nboxes(const BoxArray* ba);
void amrex_fi_boxarray_get_box(const BoxArray* ba, int i, int lo[3], int hi[3]);
void amrex_fi_boxarray_nodal_type(const BoxArray* ba, int inodal[3]);
long long amrex_fi_boxarray_numpts(const BoxArray* ba);
int  amrex_fi_boxarray_issame(const BoxArray* baa, const BoxArray* bab);
void amrex_b200_boxarray_coarsen(BoxArray* ba, int ratio);   /* BoxArray::coarsen */
void amrex_b200_boxarray_refine(BoxArray* ba, int ratio);
void amrex_b200_boxarray_convert(BoxArray* ba, const int nodal[3]);   /* BoxArray::convert(IndexType) */

/* ---- DistributionMapping (AMReX_distromap_fi.cpp:9-50) ---- */
void amrex_fi_new_distromap(DistributionMapping** dm, const BoxArray* ba);
void amrex_fi_new_distromap_from_pmap(DistributionMapping** dm, const int* pmap, const int plen);
void amrex_fi_delete_distromap(DistributionMapping* dm);
void amrex_fi_distromap_get_pmap(const DistributionMapping* dm, int* pmap, const int plen);
/* SFC map for an explicit rank count, no device needed (DistributionMapping::makeSFC, AMReX_DistributionMapping.cpp:1891) */
void amrex_b200_new_distromap_sfc(DistributionMapping** dm, const BoxArray* ba, int nprocs);
/* raw bucket number of every box from DistributionMapping::makeSFC(ba, true, nprocs) (same file :1891-1921); host only */
void amrex_b200_make_sfc(const BoxArray* ba, int nprocs, int* bucket_of_box);

/* ---- MultiFab (AMReX_multifab_fi.cpp:10-216) ---- */
void amrex_fi_new_multifab(MultiFab** mf, const BoxArray** ba, const DistributionMapping** dm, int nc, const int* ng, const int* nodal);
void amrex_fi_delete_multifab(MultiFab* mf);
int  amrex_fi_multifab_ncomp(const MultiFab* mf);
void amrex_fi_multifab_ngrow(const MultiFab* mf, int* ngv);
const BoxArray* amrex_fi_multifab_boxarray(const MultiFab* mf);
const DistributionMapping* amrex_fi_multifab_distromap(const MultiFab* mf);
/* device pointer and GROWN bounds of local grid igrd; rows are padded: use amrex_b200_multifab_strides */
void amrex_fi_multifab_dataptr_int(MultiFab* mf, int igrd, Real** dp, int lo[3], int hi[3]);
void amrex_b200_multifab_strides(const MultiFab* mf, int igrd, long long strides[3]);
Real amrex_fi_multifab_sum(const MultiFab* mf, int comp);
Real amrex_fi_multifab_norm0(const MultiFab* mf, int comp);
void amrex_fi_multifab_setval(MultiFab* mf, Real val, int ic, int nc, const int* ng);
void amrex_fi_multifab_plus(MultiFab* mf, Real val, int ic, int nc, int ng);
void amrex_fi_multifab_mult(MultiFab* mf, Real val, int ic, int nc, int ng);
void amrex_fi_multifab_add(MultiFab* dstmf, const MultiFab* srcmf, int srccomp, int dstcomp, int nc, const int* ng);
void amrex_fi_multifab_subtract(MultiFab* dstmf, const MultiFab* srcmf, int srccomp, int dstcomp, int nc, const int* ng);
void amrex_fi_multifab_saxpy(MultiFab* dstmf, Real a, const MultiFab* srcmf, int srccomp, int dstcomp, int nc, const int* ng);
void amrex_fi_multifab_copy(MultiFab* dstmf, const MultiFab* srcmf, int srccomp, int dstcomp, int nc, const int* ng);
void amrex_fi_multifab_parallelcopy(MultiFab* dstmf, const MultiFab* srcmf, int srccomp, int dstcomp, int nc, int srcng, int dstng, const Geometry* geom);
void amrex_fi_multifab_fill_boundary(MultiFab* mf, const Geometry* geom, int c, int nc, int cross);
/* ---- the rest of the reference's Base entries (Src/F_Interfaces/Base/AMReX_multifab_fi.cpp:96-300, AMReX_multifabutil_fi.cpp,
 *      AMReX_geometry_fi.cpp:27-47, AMReX_distromap_fi.cpp, AMReX_boxarray_fi.cpp, AMReX_box_fi.cpp): same names and argument order,
 *      so that the reference's Fortran modules (amrex_multifab_mod, amrex_multifabutil_mod, ...) link unchanged.  Data pointers
 *      are device pointers.  The node-centred synchronisation entries (owner masks, override / average sync) and aliased
 *      MultiFabs exist and report an error through amrex_b200_last_error: they belong to the nodal solvers, outside this path. */
Real amrex_fi_multifab_min(const MultiFab* mf, int comp, int nghost);
Real amrex_fi_multifab_max(const MultiFab* mf, int comp, int nghost);
Real amrex_fi_multifab_norm1(const MultiFab* mf, int comp);
Real amrex_fi_multifab_norm2(const MultiFab* mf, int comp);
void amrex_fi_multifab_multiply(MultiFab* dstmf, const MultiFab* srcmf, int srccomp, int dstcomp, int nc, const int* ng);
void amrex_fi_multifab_divide(MultiFab* dstmf, const MultiFab* srcmf, int srccomp, int dstcomp, int nc, const int* ng);
void amrex_fi_multifab_lincomb(MultiFab* dstmf, Real a, const MultiFab* srcmf1, int srccomp1, Real b, const MultiFab* srcmf2, int srccomp2, int dstcomp, int nc, const int* ng);
void amrex_fi_multifab_parallelcopy_gv(MultiFab* dstmf, const MultiFab* srcmf, int srccomp, int dstcomp, int nc, const int* srcng, const int* dstng, const Geometry* geom);
void amrex_fi_multifab_sum_boundary(MultiFab* mf, const Geometry* geom, int icomp, int ncomp);
void amrex_fi_build_owner_imultifab(iMultiFab** msk, const BoxArray** ba, const DistributionMapping** dm, const MultiFab* data, const Geometry* geom);
void amrex_fi_multifab_override_sync(MultiFab* mf, const Geometry* geom);
void amrex_fi_multifab_override_sync_mask(MultiFab* mf, const Geometry* geom, const iMultiFab* msk);
void amrex_fi_multifab_average_sync(MultiFab* mf, const Geometry* geom);
void amrex_fi_new_multifab_alias(MultiFab** mf, const MultiFab* srcmf, int comp, int ncomp);
void amrex_fi_new_imultifab(iMultiFab** imf, const BoxArray** ba, const DistributionMapping** dm, int nc, const int* ng, const int* nodal);
void amrex_fi_new_imultifab_alias(iMultiFab** mf, const iMultiFab* srcmf, int comp, int ncomp);
void amrex_fi_delete_imultifab(iMultiFab* imf);
void amrex_fi_imultifab_setval(iMultiFab* imf, int val, int ic, int nc, const int* ng);
void amrex_fi_imultifab_dataptr(iMultiFab* imf, MFIter* mfi, int** dp, int lo[3], int hi[3]);
void amrex_fi_multifab_dataptr_iter(MultiFab* mf, MFIter* mfi, Real** dp, int lo[3], int hi[3]);
int  amrex_fi_mfiter_allow_multiple(int allow);
void amrex_fi_new_mfiter_r(MFIter** mfi, MultiFab* mf, int tiling, int dynamic);
void amrex_fi_new_mfiter_i(MFIter** mfi, iMultiFab* imf, int tiling, int dynamic);
void amrex_fi_new_mfiter_rs(MFIter** mfi, MultiFab* mf, const int* tilesize, int dynamic);
void amrex_fi_new_mfiter_is(MFIter** mfi, iMultiFab* imf, const int* tilesize, int dynamic);
void amrex_fi_new_mfiter_badm(MFIter** mfi, BoxArray* ba, DistributionMapping* dm, int tiling, int dynamic);
void amrex_fi_new_mfiter_badm_s(MFIter** mfi, BoxArray* ba, DistributionMapping* dm, const int* tilesize, int dynamic);
void amrex_fi_delete_mfiter(MFIter* mfi);
void amrex_fi_increment_mfiter(MFIter* mfi, int* isvalid);
void amrex_fi_mfiter_is_valid(MFIter* mfi, int* isvalid);
int  amrex_fi_mfiter_grid_index(MFIter* mfi);
int  amrex_fi_mfiter_local_tile_index(MFIter* mfi);
void amrex_fi_mfiter_tilebox(MFIter* mfi, int lo[3], int hi[3], int nodal[3]);
void amrex_fi_mfiter_tilebox_iv(MFIter* mfi, int lo[3], int hi[3], const int nodal[3]);
void amrex_fi_mfiter_nodaltilebox(MFIter* mfi, int dir, int lo[3], int hi[3], int nodal[3]);
void amrex_fi_mfiter_growntilebox(MFIter* mfi, int lo[3], int hi[3], int ng, int nodal[3]);
void amrex_fi_mfiter_grownnodaltilebox(MFIter* mfi, int lo[3], int hi[3], int dir, int ng, int nodal[3]);
void amrex_fi_mfiter_validbox(MFIter* mfi, int lo[3], int hi[3], int nodal[3]);
void amrex_fi_mfiter_fabbox(MFIter* mfi, int lo[3], int hi[3], int nodal[3]);
void amrex_fi_geometry_get_pmask(const Geometry* geom, int is_per[3]);
void amrex_fi_geometry_get_probdomain(const Geometry* geom, Real problo[3], Real probhi[3]);
void amrex_fi_clone_distromap(DistributionMapping** dmo, const DistributionMapping* dmi);
int  amrex_fi_distromap_issame(const DistributionMapping* dma, const DistributionMapping* dmb);
void amrex_fi_print_distromap(const DistributionMapping* dm);
int  amrex_fi_boxarray_intersects_box(const BoxArray* ba, const int lo[3], const int hi[3]);
void amrex_fi_print_boxarray(const BoxArray* ba);
void amrex_fi_print_box(const int lo[3], const int hi[3], const int nodal[3]);
void amrex_fi_average_down(const MultiFab* S_fine, MultiFab* S_crse, const Geometry* fgeom, const Geometry* cgeom, int scomp, int ncomp, int rr);
void amrex_fi_average_down_cell_node(const MultiFab* S_fine, MultiFab* S_crse, int scomp, int ncomp, int rr);
void amrex_fi_average_down_faces(MultiFab const* fmf[], MultiFab* cmf[], const Geometry* cgeom, int scomp, int ncomp, int rr);
void amrex_fi_average_cellcenter_to_face(MultiFab* fc[], const MultiFab* cc, const Geometry* geom);
Real amrex_b200_multifab_dot(const MultiFab* x, const MultiFab* y);   /* amrex::Dot, AMReX_FabArrayUtility.H:1554 */
/* host <-> device: `h` is a Fortran-order array covering exactly the index box [lo,hi] (of the MultiFab's index
 * type); cells of every local fab inside grow(validbox,ng) are transferred.  Host memory may be pinned. */
void amrex_b200_multifab_upload(MultiFab* mf, const Real* h, const int lo[3], const int hi[3], int comp, int ng);
void amrex_b200_multifab_download(const MultiFab* mf, Real* h, const int lo[3], const int hi[3], int comp, int ng);
/* one local grid alone (global box index igrd): valid cells + ng ghost layers of THAT fab, Fortran order over the grown box */
void amrex_b200_multifab_download_fab(const MultiFab* mf, int igrd, Real* h, int comp, int ng);
/* the same, enqueued on the caller's cudaStream_t without synchronisation (pinned host memory): lets an application overlap the
 * upload of the next solve's inputs and the download of the previous solution with the running solve */
void amrex_b200_multifab_upload_async(MultiFab* mf, const Real* h, const int lo[3], const int hi[3], int comp, int ng, void* stream);
void amrex_b200_multifab_download_async(const MultiFab* mf, Real* h, const int lo[3], const int hi[3], int comp, int ng, void* stream);
/* cell centres -> faces (amrex::average_cellcenter_to_face, Src/Base/AMReX_MultiFabUtil.cpp:226) */
void amrex_b200_average_cellcenter_to_face(MultiFab* fx, MultiFab* fy, MultiFab* fz, const MultiFab* cc, const Geometry* geom);
/* ---- plotfile output in the reference's on-disk format (readable by its Tools/Plotfile: fcompare, fextrema, ...)
 *      amrex_fi_write_plotfile: Src/F_Interfaces/Base/AMReX_plotfile_fi.cpp:8-26 (same name and arguments); mf[lev] holds
 *      one component per entry of varname[]; ref_ratio[lev] for lev < nlevs-1.  One data file per rank, rank 0 writes
 *      the headers.  amrex_b200_vismf_write: amrex::VisMF::Write(mf, name) (ghost cells included). */
void amrex_fi_write_plotfile(const char* name, int nlevs, const MultiFab* mf[], const char* varname[], const Geometry* geom[],
                             Real time, const int level_steps[], const int ref_ratio[]);
void amrex_b200_vismf_write(const MultiFab* mf, const char* name);

/* ---- linear operators (Src/F_Interfaces/LinearSolvers/AMReX_abeclaplacian_fi.cpp:8-48, AMReX_poisson_fi.cpp:8,
 *      AMReX_linop_fi.cpp:8-40) ---- */
void amrex_fi_new_abeclaplacian(MLLinOp** linop, int nlevels, const Geometry* geom[], const BoxArray* ba[],
                                const DistributionMapping* dm[], int metric_term, int agglomeration,
                                int consolidation, int max_coarsening_level);
void amrex_fi_new_poisson(MLLinOp** linop, int nlevels, const Geometry* geom[], const BoxArray* ba[],
                          const DistributionMapping* dm[], int metric_term, int agglomeration,
                          int consolidation, int max_coarsening_level);
void amrex_fi_delete_linop(MLLinOp* linop);
void amrex_fi_linop_set_maxorder(MLLinOp* linop, int ord);
void amrex_fi_linop_set_domain_bc(MLLinOp* linop, const int* ilobc, const int* ihibc);   /* LinOpBCType values */
void amrex_fi_linop_set_coarse_fine_bc(MLLinOp* linop, const MultiFab* crse, int crse_ratio);
void amrex_fi_linop_set_level_bc(MLLinOp* linop, int amrlev, const MultiFab* levelbcdata);
/* MLLinOpT::setLevelBC with Robin data a*phi + b*dphi/dn = f in the ghost cells outside the Robin faces (AMReX_MLLinOp.H:220-241,
 * AMReX_MLCellLinOp.H:513-642); the reference's Fortran interface has no such entry */
void amrex_b200_linop_set_level_bc_robin(MLLinOp* linop, int amrlev, const MultiFab* levelbcdata, const MultiFab* robinbc_a,
                                         const MultiFab* robinbc_b, const MultiFab* robinbc_f);
void amrex_fi_abeclap_set_scalars(MLLinOp* linop, Real a, Real b);
void amrex_fi_abeclap_set_acoeffs(MLLinOp* linop, int amrlev, const MultiFab* alpha);
void amrex_fi_abeclap_set_bcoeffs(MLLinOp* linop, int amrlev, const MultiFab* beta[]);
/* LPInfo::setAgglomerationGridSize / setConsolidationGridSize variant of the constructors (grid size <= 0: default 32) */
void amrex_b200_new_linop(MLLinOp** linop, int kind /*0 abeclap, 1 poisson*/, int nlevels, const Geometry* geom[],
                          const BoxArray* ba[], const DistributionMapping* dm[], int agglomeration, int consolidation,
                          int max_coarsening_level, int agg_grid_size, int con_grid_size);
void amrex_b200_linop_set_smoother_fusion(MLLinOp* linop, int fuse);   /* 0: reference schedule, 1: fused colours */
/* MLCellLinOpT::setGaussSeidel (AMReX_MLCellLinOp.H:58): 1 red-black Gauss-Seidel (default), 0 damped Jacobi */
void amrex_b200_linop_set_gauss_seidel(MLLinOp* linop, int flag);
/* launch plan of the fused smoother: rows / planes per CTA tile (<= 0: automatic), L2 prefetch distance in planes (< 0: keep) */
/* n > 0: fused pass on every level whose local boxes average >= n cells (floor 32^3); n <= 0 (default): per-level cost model */
void amrex_b200_linop_set_fused_min_box_cells(MLLinOp* linop, long long n);
/* launch plan of the generation-4 fused pass (process-wide): rows per CTA tile, EARLY / LATE ring depths; 0 = accepted */
int amrex_b200_set_fused4_plan(int tile_y, int early_stages, int late_stages);
/* primitives of the operator on (amrlev 0, mglev), exposed for parity tests (MLCellLinOp::smooth/apply/...) */
int  amrex_b200_linop_num_mg_levels(const MLLinOp* linop, int amrlev);
void amrex_b200_linop_prepare(MLLinOp* linop);
void amrex_b200_linop_make(MLLinOp* linop, MultiFab** mf, int amrlev, int mglev, int ng);
/* flags: bit 0 = skip_fillboundary (AMReX_MLCellLinOp.H:1206), bit 1 = zero_input (sol is taken as identically zero
 * without being read: MLMG's cor.setVal(0) + first pre-smooth in one pass) */
void amrex_b200_linop_smooth(MLLinOp* linop, int amrlev, int mglev, MultiFab* sol, const MultiFab* rhs, int flags);
void amrex_b200_linop_apply(MLLinOp* linop, int amrlev, int mglev, MultiFab* out, MultiFab* in, int inhomog);
void amrex_b200_linop_residual(MLLinOp* linop, int amrlev, int mglev, MultiFab* resid, MultiFab* x, const MultiFab* b, int inhomog);
void amrex_b200_linop_restriction(MLLinOp* linop, int amrlev, int cmglev, MultiFab* crse, MultiFab* fine);
void amrex_b200_linop_interp_add(MLLinOp* linop, int amrlev, int fmglev, MultiFab* fine, const MultiFab* crse);
void amrex_b200_linop_get_coeff(MLLinOp* linop, int amrlev, int mglev, int which /*0 a, 1..3 b*/, const MultiFab** mf);
/* MG hierarchy of the operator: boxes (6 ints each) and owner ranks of (amrlev, mglev) */
int  amrex_b200_linop_level_nboxes(const MLLinOp* linop, int amrlev, int mglev);
void amrex_b200_linop_level_boxes(const MLLinOp* linop, int amrlev, int mglev, int* boxes6, int* pmap, int* domain6);

/* ---- MLMG (AMReX_multigrid_fi.cpp:8-113) ---- */
void amrex_fi_new_multigrid(MLMG** mlmg, MLLinOp* lp);
void amrex_fi_delete_multigrid(MLMG* mlmg);
Real amrex_fi_multigrid_solve(MLMG* mlmg, MultiFab* a_sol[], MultiFab* a_rhs[], Real a_tol_rel, Real a_tol_abs);
void amrex_fi_multigrid_comp_residual(MLMG* mlmg, MultiFab* a_res[], MultiFab* a_sol[], MultiFab* a_rhs[]);
/* post-solve: face-centred grad(phi) / fluxes -b*grad(phi) of the solution of the last solve, arrays [level*3 + dir] of
 * MultiFabs on the faces of the level's grids (reference: AMReX_multigrid_fi.cpp:27-51, MLMG::getGradSolution / getFluxes) */
void amrex_fi_multigrid_get_grad_solution(MLMG* mlmg, MultiFab* a_grad_sol[]);
void amrex_fi_multigrid_get_fluxes(MLMG* mlmg, MultiFab* a_fluxes[]);
void amrex_fi_multigrid_set_verbose(MLMG* mlmg, int v);
void amrex_fi_multigrid_set_max_iter(MLMG* mlmg, int n);
void amrex_fi_multigrid_set_max_fmg_iter(MLMG* mlmg, int n);
void amrex_fi_multigrid_set_fixed_iter(MLMG* mlmg, int n);
void amrex_fi_multigrid_set_bottom_solver(MLMG* mlmg, int s);   /* 0 smoother, 1 bicgstab, 2 cg */
void amrex_fi_multigrid_set_bottom_verbose(MLMG* mlmg, int n);
void amrex_fi_multigrid_set_always_use_bnorm(MLMG* mlmg, int f);
void amrex_fi_multigrid_set_final_fill_bc(MLMG* mlmg, int f);
/* results of the last solve (MLMG::getNumIters/getResidualHistory/getInitRHS/getInitResidual/getNumCGIters) */
int  amrex_b200_multigrid_num_iters(const MLMG* mlmg);
int  amrex_b200_multigrid_residual_history(const MLMG* mlmg, Real* hist, int capacity);
Real amrex_b200_multigrid_init_rhs(const MLMG* mlmg);
Real amrex_b200_multigrid_init_residual(const MLMG* mlmg);
int  amrex_b200_multigrid_cg_iters(const MLMG* mlmg, int* iters, int capacity);
void amrex_b200_multigrid_timers(const MLMG* mlmg, double t[3]);   /* solve, iter, bottom wall seconds */

/* ---- GMRES preconditioned by MLMG V-cycles: amrex::GMRESMLMG (LinearSolvers/AMReX_GMRES_MLMG.H:19-236; the reference has
 *      no C interface for it, members are exposed one to one).  Single AMR level; the MLMG object must outlive it. */
void amrex_b200_new_gmres_mlmg(GMRESMLMG** g, MLMG* mlmg);
void amrex_b200_delete_gmres_mlmg(GMRESMLMG* g);
void amrex_b200_gmres_mlmg_solve(GMRESMLMG* g, MultiFab* sol, const MultiFab* rhs, Real tol_rel, Real tol_abs);
void amrex_b200_gmres_mlmg_set_verbose(GMRESMLMG* g, int v);
void amrex_b200_gmres_mlmg_set_max_iters(GMRESMLMG* g, int n);
void amrex_b200_gmres_mlmg_set_restart_length(GMRESMLMG* g, int n);
void amrex_b200_gmres_mlmg_use_precond(GMRESMLMG* g, int f);
void amrex_b200_gmres_mlmg_set_precond_num_iters(GMRESMLMG* g, int n);
void amrex_b200_gmres_mlmg_set_property_of_zero(GMRESMLMG* g, int f);
int  amrex_b200_gmres_mlmg_num_iters(const GMRESMLMG* g);
int  amrex_b200_gmres_mlmg_status(const GMRESMLMG* g);             /* 0 converged, 1 iteration limit, -1 not run */
Real amrex_b200_gmres_mlmg_residual_norm(const GMRESMLMG* g);      /* 2-norm estimate of the last iteration */
int  amrex_b200_gmres_mlmg_residual_history(const GMRESMLMG* g, Real* hist, int capacity);

/* ---- host-only metadata (no device): bit-exact parity targets ---- */
/* MG hierarchy as MLLinOpT::defineGrids would build it for `nprocs` ranks (AMReX_MLLinOp.H:795-1165).
 * Returns a handle; query with the functions below; free with amrex_b200_hierarchy_delete. */
void* amrex_b200_hierarchy_new(int nlevels, const Geometry* geom[], const BoxArray* ba[], const DistributionMapping* dm[],
                               int agglomeration, int consolidation, int max_coarsening_level,
                               int agg_grid_size, int con_grid_size, int nprocs);
void amrex_b200_hierarchy_delete(void* h);
int  amrex_b200_hierarchy_num_mg_levels(const void* h, int amrlev);
int  amrex_b200_hierarchy_nboxes(const void* h, int amrlev, int mglev);
/* amrex::isMFIterSafe between MG levels m1, m2 of AMR level a (equal dmaps and shared box list, BoxArray::SameRefs) */
int  amrex_b200_hierarchy_shares_box_list(const void* h, int a, int m1, int m2);
void amrex_b200_hierarchy_level(const void* h, int amrlev, int mglev, int* boxes6, int* pmap, int* domain6);
/* FillBoundary / ParallelCopy tag lists as rank `myproc` sees them (AMReX_FabArrayBase.cpp:658-877, 324-467).
 * kind: 0 LocTags, 1 SndTags, 2 RcvTags.  Each tag is written as 15 ints:
 * dbox lo[3] hi[3], sbox lo[3] hi[3], dstIndex, srcIndex, peer rank (-1 for local).  Returns the tag count
 * (call with out == NULL to size the buffer). */
int  amrex_b200_fb_tags(const BoxArray* ba, const DistributionMapping* dm, int ng, int cross, const int period[3],
                        int myproc, int kind, int* out, int capacity);
int  amrex_b200_cpc_tags(const BoxArray* ba_dst, const DistributionMapping* dm_dst, int ng_dst,
                         const BoxArray* ba_src, const DistributionMapping* dm_src, int ng_src,
                         const int period[3], int myproc, int kind, int* out, int capacity);
/* Face links of the cross-stencil, one-ghost-cell FillBoundary as rank `myproc` sees it (the form in which the shell sweep
 * of the fused smoother replaces the copies between boxes of one GPU): 4 ints per (local box, face) in
 * out - linked local box index (-1: none) and the index shift[3]; local boxes in ascending global index, faces 0,1,2 = x,y,z
 * low, 3,4,5 = high.  Returns the number of local boxes when every local tag is one whole face from one box, -2 when the
 * pattern has no such form, -1 on error (out == NULL: only the return value).  Host only. */
int  amrex_b200_fb_face_links(const BoxArray* ba, const DistributionMapping* dm, const int period[3], int myproc,
                              int* out, int capacity);

#ifdef __cplusplus
}
#endif
#endif
