"""ctypes binding of libamrex_b200.so (include/amrex_b200_fi.h).  Thin by design: every method is one C-ABI call."""
import ctypes as C
import os

import numpy as np

# (AMREX_B200_LIB: another build of the same library, for A/B timing of kernel variants on the GPU box)
LIB_PATH = os.environ.get("AMREX_B200_LIB") or os.path.join(os.path.dirname(os.path.abspath(__file__)), "lib", "libamrex_b200.so")

_lib = None


class AmrexError(RuntimeError):
    pass


class LinOpBCType:
    interior, Dirichlet, Neumann, reflect_odd = 0, 101, 102, 103
    inhomogNeumann, Robin, Periodic = 107, 108, 200


_P = C.c_void_p
_PP = C.POINTER(C.c_void_p)
_I = C.c_int
_D = C.c_double
_IP = C.POINTER(C.c_int)
_DP = C.POINTER(C.c_double)
_LL = C.c_longlong

# name -> (restype, argtypes); only what the Python side calls. tests/test_abi.py checks EVERY header symbol exports.
_SIGS = {
    "amrex_b200_init": (_I, [_I]), "amrex_b200_finalize": (None, []), "amrex_b200_initialized": (_I, []),
    "amrex_b200_nccl_unique_id_bytes": (_I, []), "amrex_b200_nccl_get_unique_id": (_I, [_P]),
    "amrex_b200_comm_init": (_I, [_I, _I, _P]), "amrex_b200_comm_finalize": (None, []),
    "amrex_b200_myproc": (_I, []), "amrex_b200_nprocs": (_I, []),
    "amrex_b200_last_error": (C.c_char_p, []), "amrex_b200_clear_error": (None, []),
    "amrex_b200_synchronize": (None, []), "amrex_b200_launch_count": (_LL, []), "amrex_b200_reset_launch_count": (None, []),
    "amrex_b200_stream": (_P, []),
    "amrex_b200_profile_enable": (None, [_I]), "amrex_b200_profile_report": (_I, [C.c_char_p, _I]),
    "amrex_b200_geometry_setup": (None, [_DP, _DP, _IP]),
    "amrex_fi_new_geometry": (None, [_PP, _IP, _IP]), "amrex_fi_delete_geometry": (None, [_P]),
    "amrex_fi_new_boxarray": (None, [_PP, _IP, _IP]), "amrex_fi_new_boxarray_from_bxfarr": (None, [_PP, _IP, _I, _I, _I]),
    "amrex_fi_delete_boxarray": (None, [_P]), "amrex_fi_clone_boxarray": (None, [_PP, _P]),
    "amrex_fi_boxarray_maxsize": (None, [_P, _IP]), "amrex_fi_boxarray_nboxes": (_LL, [_P]),
    "amrex_fi_boxarray_get_box": (None, [_P, _I, _IP, _IP]), "amrex_fi_boxarray_numpts": (_LL, [_P]),
    "amrex_b200_boxarray_coarsen": (None, [_P, _I]), "amrex_b200_boxarray_refine": (None, [_P, _I]),
    "amrex_b200_boxarray_convert": (None, [_P, C.POINTER(C.c_int)]),
    "amrex_fi_new_distromap": (None, [_PP, _P]), "amrex_fi_new_distromap_from_pmap": (None, [_PP, _IP, _I]),
    "amrex_fi_delete_distromap": (None, [_P]), "amrex_fi_distromap_get_pmap": (None, [_P, _IP, _I]),
    "amrex_b200_new_distromap_sfc": (None, [_PP, _P, _I]), "amrex_b200_make_sfc": (None, [_P, _I, _IP]),
    "amrex_fi_new_multifab": (None, [_PP, _PP, _PP, _I, _IP, _IP]), "amrex_fi_delete_multifab": (None, [_P]),
    "amrex_fi_multifab_sum": (_D, [_P, _I]), "amrex_fi_multifab_norm0": (_D, [_P, _I]),
    "amrex_fi_multifab_setval": (None, [_P, _D, _I, _I, _IP]),
    "amrex_fi_multifab_copy": (None, [_P, _P, _I, _I, _I, _IP]),
    "amrex_fi_multifab_saxpy": (None, [_P, _D, _P, _I, _I, _I, _IP]),
    "amrex_fi_multifab_parallelcopy": (None, [_P, _P, _I, _I, _I, _I, _I, _P]),
    "amrex_fi_multifab_fill_boundary": (None, [_P, _P, _I, _I, _I]),
    "amrex_b200_multifab_dot": (_D, [_P, _P]),
    "amrex_b200_multifab_upload": (None, [_P, _P, _IP, _IP, _I, _I]),
    "amrex_b200_multifab_upload_async": (None, [_P, _P, _IP, _IP, _I, _I, _P]),
    "amrex_b200_multifab_download_fab": (None, [_P, _I, _P, _I, _I]),
    "amrex_b200_multifab_download_async": (None, [_P, _P, _IP, _IP, _I, _I, _P]),
    "amrex_b200_multifab_download": (None, [_P, _P, _IP, _IP, _I, _I]),
    "amrex_b200_average_cellcenter_to_face": (None, [_P, _P, _P, _P, _P]),
    "amrex_fi_write_plotfile": (None, [C.c_char_p, _I, _PP, C.POINTER(C.c_char_p), _PP, _D, _IP, _IP]),
    "amrex_b200_vismf_write": (None, [_P, C.c_char_p]),
    "amrex_fi_multifab_subtract": (None, [_P, _P, _I, _I, _I, _IP]),
    "amrex_b200_new_linop": (None, [_PP, _I, _I, _PP, _PP, _PP, _I, _I, _I, _I, _I]),
    "amrex_fi_new_abeclaplacian": (None, [_PP, _I, _PP, _PP, _PP, _I, _I, _I, _I]),
    "amrex_fi_new_poisson": (None, [_PP, _I, _PP, _PP, _PP, _I, _I, _I, _I]),
    "amrex_fi_delete_linop": (None, [_P]), "amrex_fi_linop_set_maxorder": (None, [_P, _I]),
    "amrex_fi_linop_set_domain_bc": (None, [_P, _IP, _IP]), "amrex_fi_linop_set_level_bc": (None, [_P, _I, _P]),
    "amrex_b200_linop_set_level_bc_robin": (None, [_P, _I, _P, _P, _P, _P]),
    "amrex_fi_linop_set_coarse_fine_bc": (None, [_P, _P, _I]),
    "amrex_fi_abeclap_set_scalars": (None, [_P, _D, _D]), "amrex_fi_abeclap_set_acoeffs": (None, [_P, _I, _P]),
    "amrex_fi_abeclap_set_bcoeffs": (None, [_P, _I, _PP]),
    "amrex_b200_linop_set_smoother_fusion": (None, [_P, _I]),
    "amrex_b200_linop_set_gauss_seidel": (None, [_P, _I]),
    "amrex_b200_set_fused4_plan": (_I, [_I, _I, _I]),
    "b200mg_set_gsrb4_sync": (None, [_I]),
    "amrex_b200_linop_set_fused_min_box_cells": (None, [_P, C.c_longlong]),
    "amrex_b200_linop_num_mg_levels": (_I, [_P, _I]), "amrex_b200_linop_prepare": (None, [_P]),
    "amrex_b200_linop_make": (None, [_P, _PP, _I, _I, _I]),
    "amrex_b200_linop_smooth": (None, [_P, _I, _I, _P, _P, _I]),
    "amrex_b200_linop_apply": (None, [_P, _I, _I, _P, _P, _I]),
    "amrex_b200_linop_residual": (None, [_P, _I, _I, _P, _P, _P, _I]),
    "amrex_b200_linop_restriction": (None, [_P, _I, _I, _P, _P]),
    "amrex_b200_linop_interp_add": (None, [_P, _I, _I, _P, _P]),
    "amrex_b200_linop_get_coeff": (None, [_P, _I, _I, _I, _PP]),
    "amrex_b200_linop_level_nboxes": (_I, [_P, _I, _I]), "amrex_b200_linop_level_boxes": (None, [_P, _I, _I, _IP, _IP, _IP]),
    "amrex_fi_new_multigrid": (None, [_PP, _P]), "amrex_fi_delete_multigrid": (None, [_P]),
    "amrex_fi_multigrid_solve": (_D, [_P, _PP, _PP, _D, _D]),
    "amrex_fi_multigrid_comp_residual": (None, [_P, _PP, _PP, _PP]),
    "amrex_fi_multigrid_get_grad_solution": (None, [_P, _PP]), "amrex_fi_multigrid_get_fluxes": (None, [_P, _PP]),
    "amrex_fi_multigrid_set_verbose": (None, [_P, _I]), "amrex_fi_multigrid_set_max_iter": (None, [_P, _I]),
    "amrex_fi_multigrid_set_max_fmg_iter": (None, [_P, _I]), "amrex_fi_multigrid_set_fixed_iter": (None, [_P, _I]),
    "amrex_fi_multigrid_set_bottom_solver": (None, [_P, _I]), "amrex_fi_multigrid_set_bottom_verbose": (None, [_P, _I]),
    "amrex_fi_multigrid_set_always_use_bnorm": (None, [_P, _I]), "amrex_fi_multigrid_set_final_fill_bc": (None, [_P, _I]),
    "amrex_b200_multigrid_num_iters": (_I, [_P]), "amrex_b200_multigrid_residual_history": (_I, [_P, _DP, _I]),
    "amrex_b200_multigrid_init_rhs": (_D, [_P]), "amrex_b200_multigrid_init_residual": (_D, [_P]),
    "amrex_b200_multigrid_cg_iters": (_I, [_P, _IP, _I]), "amrex_b200_multigrid_timers": (None, [_P, _DP]),
    "amrex_b200_new_gmres_mlmg": (None, [_PP, _P]), "amrex_b200_delete_gmres_mlmg": (None, [_P]),
    "amrex_b200_gmres_mlmg_solve": (None, [_P, _P, _P, _D, _D]),
    "amrex_b200_gmres_mlmg_set_verbose": (None, [_P, _I]), "amrex_b200_gmres_mlmg_set_max_iters": (None, [_P, _I]),
    "amrex_b200_gmres_mlmg_set_restart_length": (None, [_P, _I]), "amrex_b200_gmres_mlmg_use_precond": (None, [_P, _I]),
    "amrex_b200_gmres_mlmg_set_precond_num_iters": (None, [_P, _I]),
    "amrex_b200_gmres_mlmg_set_property_of_zero": (None, [_P, _I]),
    "amrex_b200_gmres_mlmg_num_iters": (_I, [_P]), "amrex_b200_gmres_mlmg_status": (_I, [_P]),
    "amrex_b200_gmres_mlmg_residual_norm": (_D, [_P]), "amrex_b200_gmres_mlmg_residual_history": (_I, [_P, _DP, _I]),
    "amrex_b200_hierarchy_new": (_P, [_I, _PP, _PP, _PP, _I, _I, _I, _I, _I, _I]),
    "amrex_b200_hierarchy_delete": (None, [_P]), "amrex_b200_hierarchy_num_mg_levels": (_I, [_P, _I]),
    "amrex_b200_hierarchy_nboxes": (_I, [_P, _I, _I]), "amrex_b200_hierarchy_level": (None, [_P, _I, _I, _IP, _IP, _IP]),
    "amrex_b200_hierarchy_shares_box_list": (_I, [_P, _I, _I, _I]),
    "amrex_b200_fb_tags": (_I, [_P, _P, _I, _I, _IP, _I, _I, _IP, _I]),
    "amrex_b200_cpc_tags": (_I, [_P, _P, _I, _P, _P, _I, _IP, _I, _I, _IP, _I]),
    "amrex_b200_fb_face_links": (_I, [_P, _P, _IP, _I, _IP, _I]),
}


def load_library(path=LIB_PATH):
    """Load libamrex_b200.so.  torch (if importable) is imported first so that one NCCL is shared by both."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(path):
        raise AmrexError(f"{path} not found: build it with `make -C amrex_b200/csrc` (there is no Python/CPU fallback)")
    try:
        import torch  # noqa: F401  (loads libcudart/libnccl with torch's RPATH before ours resolves the sonames)
    except Exception:
        pass
    _lib = C.CDLL(path, mode=C.RTLD_GLOBAL)
    for name, (res, args) in _SIGS.items():
        f = getattr(_lib, name)
        f.restype = res
        f.argtypes = args
    return _lib


class _LibProxy:
    def __getattr__(self, name):
        return getattr(load_library(), name)


lib = _LibProxy()


def check():
    e = lib.amrex_b200_last_error()
    if e:
        msg = e.decode()
        lib.amrex_b200_clear_error()
        raise AmrexError(msg)


def _i3(v):
    return (C.c_int * 3)(*[int(x) for x in v])


def _d3(v):
    return (C.c_double * 3)(*[float(x) for x in v])


def init(device_id=None):
    """amrex::Initialize for one process-per-GPU rank: selects the GPU (LOCAL_RANK by default)."""
    if device_id is None:
        device_id = int(os.environ.get("LOCAL_RANK", "0"))
    if lib.amrex_b200_init(device_id) != 0:
        check()
        raise AmrexError("amrex_b200_init failed")


def profile_enable(on=True):
    lib.amrex_b200_profile_enable(int(on))


def profile_report():
    """[(kernel, scope, launches, total_ms, min_ms, max_ms)] of the launches since profiling was enabled / last report."""
    n = lib.amrex_b200_profile_report(None, 0)
    check()
    buf = C.create_string_buffer(n + 1)
    lib.amrex_b200_profile_report(buf, n + 1)
    out = []
    for line in buf.value.decode().splitlines():
        f = line.split()
        out.append((f[0], int(f[1]), int(f[2]), float(f[3]), float(f[4]), float(f[5])))
    return out


def finalize():
    lib.amrex_b200_finalize()
    check()


def comm_init_from_torch():
    """Create the library's NCCL communicator over the ranks of an initialised torch.distributed process group.
    torch.distributed is only the bootstrap (broadcast of the NCCL unique id)."""
    import torch
    import torch.distributed as dist
    rank, world = dist.get_rank(), dist.get_world_size()
    n = lib.amrex_b200_nccl_unique_id_bytes()
    buf = (C.c_ubyte * n)()
    if rank == 0 and world > 1:
        if lib.amrex_b200_nccl_get_unique_id(buf) != 0:
            check()
    if world > 1:
        dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
        t = torch.tensor(list(buf), dtype=torch.uint8, device=dev)
        dist.broadcast(t, 0)
        buf = (C.c_ubyte * n)(*t.cpu().tolist())
    if lib.amrex_b200_comm_init(rank, world, buf) != 0:
        check()
    check()


class _Obj:
    _deleter = None

    def __init__(self, ptr, own=True):
        self.ptr = ptr
        self._own = own

    def __del__(self):
        try:
            if self._own and self.ptr and self._deleter and _lib is not None:
                getattr(_lib, self._deleter)(self.ptr)
        except Exception:
            pass
        self.ptr = None


class Geometry(_Obj):
    _deleter = "amrex_fi_delete_geometry"

    @staticmethod
    def setup(problo=(0., 0., 0.), probhi=(1., 1., 1.), is_periodic=(0, 0, 0)):
        lib.amrex_b200_geometry_setup(_d3(problo), _d3(probhi), _i3(is_periodic))

    def __init__(self, lo, hi):
        p = C.c_void_p()
        lib.amrex_fi_new_geometry(C.byref(p), _i3(lo), _i3(hi))
        check()
        super().__init__(p)
        self.lo, self.hi = tuple(lo), tuple(hi)


class BoxArray(_Obj):
    _deleter = "amrex_fi_delete_boxarray"

    def __init__(self, lo=None, hi=None, boxes=None, ptr=None):
        p = C.c_void_p()
        if ptr is not None:
            super().__init__(ptr, own=False)
            return
        if boxes is not None:
            arr = []
            for b in boxes:   # (lo0,lo1,lo2,hi0,hi1,hi2) -> Fortran (2,3) layout: lo0,hi0,lo1,hi1,lo2,hi2
                arr += [b[0], b[3], b[1], b[4], b[2], b[5]]
            a = (C.c_int * len(arr))(*arr)
            lib.amrex_fi_new_boxarray_from_bxfarr(C.byref(p), a, 2, 3, len(boxes))
        else:
            lib.amrex_fi_new_boxarray(C.byref(p), _i3(lo), _i3(hi))
        check()
        super().__init__(p)

    def maxSize(self, n):
        sz = (n, n, n) if np.isscalar(n) else n
        lib.amrex_fi_boxarray_maxsize(self.ptr, _i3(sz))
        check()
        return self

    def convert(self, nodal):
        lib.amrex_b200_boxarray_convert(self.ptr, _i3(nodal))
        check()
        return self

    def coarsen(self, r):
        lib.amrex_b200_boxarray_coarsen(self.ptr, r)
        check()
        return self

    def clone(self):
        p = C.c_void_p()
        lib.amrex_fi_clone_boxarray(C.byref(p), self.ptr)
        b = BoxArray.__new__(BoxArray)
        _Obj.__init__(b, p)
        return b

    def size(self):
        return int(lib.amrex_fi_boxarray_nboxes(self.ptr))

    def numPts(self):
        return int(lib.amrex_fi_boxarray_numpts(self.ptr))

    def box(self, i):
        lo, hi = _i3((0, 0, 0)), _i3((0, 0, 0))
        lib.amrex_fi_boxarray_get_box(self.ptr, i, lo, hi)
        return tuple(lo) + tuple(hi)

    def boxes(self):
        return [self.box(i) for i in range(self.size())]


class DistributionMapping(_Obj):
    _deleter = "amrex_fi_delete_distromap"

    def __init__(self, ba=None, pmap=None, nprocs=None):
        p = C.c_void_p()
        if pmap is not None:
            a = (C.c_int * len(pmap))(*[int(x) for x in pmap])
            lib.amrex_fi_new_distromap_from_pmap(C.byref(p), a, len(pmap))
        elif nprocs is not None:
            lib.amrex_b200_new_distromap_sfc(C.byref(p), ba.ptr, nprocs)
        else:
            lib.amrex_fi_new_distromap(C.byref(p), ba.ptr)
        check()
        super().__init__(p)

    def pmap(self, n):
        a = (C.c_int * n)()
        lib.amrex_fi_distromap_get_pmap(self.ptr, a, n)
        return list(a)


class MultiFab(_Obj):
    _deleter = "amrex_fi_delete_multifab"

    def __init__(self, ba=None, dm=None, ncomp=1, ngrow=0, nodal=(0, 0, 0), ptr=None, own=True):
        if ptr is not None:
            super().__init__(ptr, own=own)
            return
        p = C.c_void_p()
        pba, pdm = C.c_void_p(ba.ptr.value), C.c_void_p(dm.ptr.value)
        lib.amrex_fi_new_multifab(C.byref(p), C.byref(pba), C.byref(pdm), ncomp, _i3((ngrow,) * 3), _i3(nodal))
        check()
        super().__init__(p)
        self.ngrow, self.nodal = ngrow, tuple(nodal)

    def setVal(self, v, ng=0):
        lib.amrex_fi_multifab_setval(self.ptr, float(v), 0, 1, _i3((ng,) * 3))
        check()

    def norm0(self):
        r = lib.amrex_fi_multifab_norm0(self.ptr, 0)
        check()
        return r

    def sum(self):
        r = lib.amrex_fi_multifab_sum(self.ptr, 0)
        check()
        return r

    def dot(self, other):
        r = lib.amrex_b200_multifab_dot(self.ptr, other.ptr)
        check()
        return r

    def copy_from(self, src, ng=0, scomp=0, dcomp=0, ncomp=1):
        lib.amrex_fi_multifab_copy(self.ptr, src.ptr, scomp, dcomp, ncomp, _i3((ng,) * 3))
        check()

    def subtract(self, src, scomp=0, dcomp=0, ncomp=1, ng=0):
        """self(dcomp..) -= src(scomp..)"""
        lib.amrex_fi_multifab_subtract(self.ptr, src.ptr, scomp, dcomp, ncomp, _i3((ng,) * 3))
        check()

    def fill_boundary(self, geom, cross=False, comp=0, ncomp=1):
        lib.amrex_fi_multifab_fill_boundary(self.ptr, geom.ptr, comp, ncomp, int(cross))
        check()

    def parallel_copy(self, src, geom, srcng=0, dstng=0, scomp=0, dcomp=0, ncomp=1):
        lib.amrex_fi_multifab_parallelcopy(self.ptr, src.ptr, scomp, dcomp, ncomp, srcng, dstng, geom.ptr)
        check()

    def upload(self, arr, lo, ng=0, comp=0):
        """arr: numpy float64 array indexed [i,j,k] (any memory order) whose [0,0,0] element is index `lo`."""
        a = np.asfortranarray(arr, dtype=np.float64)
        hi = [lo[d] + a.shape[d] - 1 for d in range(3)]
        lib.amrex_b200_multifab_upload(self.ptr, a.ctypes.data_as(C.c_void_p), _i3(lo), _i3(hi), comp, ng)
        check()

    def upload_ptr(self, host_ptr, lo, hi, ng=0):
        lib.amrex_b200_multifab_upload(self.ptr, C.c_void_p(host_ptr), _i3(lo), _i3(hi), 0, ng)
        check()

    def download(self, lo, shape, ng=0, comp=0):
        a = np.zeros(shape, dtype=np.float64, order="F")
        hi = [lo[d] + shape[d] - 1 for d in range(3)]
        lib.amrex_b200_multifab_download(self.ptr, a.ctypes.data_as(C.c_void_p), _i3(lo), _i3(hi), comp, ng)
        check()
        return a

    def download_fab(self, igrd, box, ng=0, comp=0):
        """local grid igrd alone (its own ghost cells, nothing from its neighbours); box = (lo0,lo1,lo2,hi0,hi1,hi2) of the grid"""
        shape = tuple(box[d + 3] - box[d] + 1 + 2 * ng for d in range(3))
        a = np.zeros(shape, dtype=np.float64, order="F")
        lib.amrex_b200_multifab_download_fab(self.ptr, igrd, a.ctypes.data_as(C.c_void_p), comp, ng)
        check()
        return a

    def upload_ptr_async(self, host_ptr, lo, hi, stream, ng=0):
        """enqueue on the CUDA stream `stream` (a cudaStream_t as int), no synchronisation; pinned host memory"""
        lib.amrex_b200_multifab_upload_async(self.ptr, C.c_void_p(host_ptr), _i3(lo), _i3(hi), 0, ng, C.c_void_p(stream))
        check()

    def download_ptr_async(self, host_ptr, lo, hi, stream, ng=0):
        lib.amrex_b200_multifab_download_async(self.ptr, C.c_void_p(host_ptr), _i3(lo), _i3(hi), 0, ng, C.c_void_p(stream))
        check()

    def download_ptr(self, host_ptr, lo, hi, ng=0):
        lib.amrex_b200_multifab_download(self.ptr, C.c_void_p(host_ptr), _i3(lo), _i3(hi), 0, ng)
        check()


def _ptr_array(objs):
    return (C.c_void_p * len(objs))(*[o.ptr.value if isinstance(o.ptr, C.c_void_p) else o.ptr for o in objs])


class MLLinOp(_Obj):
    _deleter = "amrex_fi_delete_linop"
    _kind = None

    def __init__(self, geom, ba, dm, agglomeration=1, consolidation=1, max_coarsening_level=30,
                 agg_grid_size=-1, con_grid_size=-1):
        p = C.c_void_p()
        self._keep = (list(geom), list(ba), list(dm))
        lib.amrex_b200_new_linop(C.byref(p), self._kind, len(geom), _ptr_array(geom), _ptr_array(ba), _ptr_array(dm),
                                 agglomeration, consolidation, max_coarsening_level, agg_grid_size, con_grid_size)
        check()
        super().__init__(p)

    def setMaxOrder(self, o):
        lib.amrex_fi_linop_set_maxorder(self.ptr, o)

    def setDomainBC(self, lo, hi):
        lib.amrex_fi_linop_set_domain_bc(self.ptr, _i3(lo), _i3(hi))
        check()

    def setLevelBC(self, amrlev, mf, robin=None):
        """robin: (a, b, f) MultiFabs with the Robin data a*phi + b*dphi/dn = f in their ghost cells"""
        if robin is not None:
            lib.amrex_b200_linop_set_level_bc_robin(self.ptr, amrlev, mf.ptr if mf is not None else None, robin[0].ptr, robin[1].ptr, robin[2].ptr)
        else:
            lib.amrex_fi_linop_set_level_bc(self.ptr, amrlev, mf.ptr if mf is not None else None)
        check()

    def setCoarseFineBC(self, crse, ratio):
        lib.amrex_fi_linop_set_coarse_fine_bc(self.ptr, crse.ptr, ratio)
        check()

    def setSmootherFusion(self, f):
        lib.amrex_b200_linop_set_smoother_fusion(self.ptr, int(f))

    def setGaussSeidel(self, flag):
        lib.amrex_b200_linop_set_gauss_seidel(self.ptr, int(bool(flag)))

    def setFusedMinBoxCells(self, n):
        lib.amrex_b200_linop_set_fused_min_box_cells(self.ptr, int(n))

    def NMGLevels(self, amrlev=0):
        return lib.amrex_b200_linop_num_mg_levels(self.ptr, amrlev)

    def prepareForSolve(self):
        lib.amrex_b200_linop_prepare(self.ptr)
        check()

    def make(self, amrlev, mglev, ng):
        p = C.c_void_p()
        lib.amrex_b200_linop_make(self.ptr, C.byref(p), amrlev, mglev, ng)
        check()
        return MultiFab(ptr=p)

    def smooth(self, amrlev, mglev, sol, rhs, skip_fillboundary=False, zero_input=False):
        lib.amrex_b200_linop_smooth(self.ptr, amrlev, mglev, sol.ptr, rhs.ptr, int(bool(skip_fillboundary)) | (2 if zero_input else 0))
        check()

    def apply(self, amrlev, mglev, out, inp):
        lib.amrex_b200_linop_apply(self.ptr, amrlev, mglev, out.ptr, inp.ptr, 0)
        check()

    def residual(self, amrlev, mglev, resid, x, b, inhomog=False):
        lib.amrex_b200_linop_residual(self.ptr, amrlev, mglev, resid.ptr, x.ptr, b.ptr, int(inhomog))
        check()

    def restriction(self, amrlev, cmglev, crse, fine):
        lib.amrex_b200_linop_restriction(self.ptr, amrlev, cmglev, crse.ptr, fine.ptr)
        check()

    def interp_add(self, amrlev, fmglev, fine, crse):
        lib.amrex_b200_linop_interp_add(self.ptr, amrlev, fmglev, fine.ptr, crse.ptr)
        check()

    def coeff(self, amrlev, mglev, which):
        p = C.c_void_p()
        lib.amrex_b200_linop_get_coeff(self.ptr, amrlev, mglev, which, C.byref(p))
        check()
        return MultiFab(ptr=p, own=False)

    def level(self, amrlev, mglev):
        n = lib.amrex_b200_linop_level_nboxes(self.ptr, amrlev, mglev)
        boxes, pmap, dom = (C.c_int * (6 * n))(), (C.c_int * n)(), (C.c_int * 6)()
        lib.amrex_b200_linop_level_boxes(self.ptr, amrlev, mglev, boxes, pmap, dom)
        return [tuple(boxes[6 * i:6 * i + 6]) for i in range(n)], list(pmap), tuple(dom)


class MLABecLaplacian(MLLinOp):
    _kind = 0

    def setScalars(self, a, b):
        lib.amrex_fi_abeclap_set_scalars(self.ptr, float(a), float(b))
        check()

    def setACoeffs(self, amrlev, mf):
        lib.amrex_fi_abeclap_set_acoeffs(self.ptr, amrlev, mf.ptr)
        check()

    def setBCoeffs(self, amrlev, beta):
        lib.amrex_fi_abeclap_set_bcoeffs(self.ptr, amrlev, _ptr_array(beta))
        check()


class MLPoisson(MLLinOp):
    _kind = 1


class MLALaplacian(MLABecLaplacian):
    """(alpha a - beta Laplacian): the reference's MLALaplacian (kind 2 of amrex_b200_new_linop)"""
    _kind = 2


class MLMG(_Obj):
    _deleter = "amrex_fi_delete_multigrid"

    def __init__(self, linop):
        p = C.c_void_p()
        self._linop = linop
        lib.amrex_fi_new_multigrid(C.byref(p), linop.ptr)
        check()
        super().__init__(p)

    def setVerbose(self, v): lib.amrex_fi_multigrid_set_verbose(self.ptr, v)
    def setBottomVerbose(self, v): lib.amrex_fi_multigrid_set_bottom_verbose(self.ptr, v)
    def setMaxIter(self, n): lib.amrex_fi_multigrid_set_max_iter(self.ptr, n)
    def setMaxFmgIter(self, n): lib.amrex_fi_multigrid_set_max_fmg_iter(self.ptr, n)
    def setFixedIter(self, n): lib.amrex_fi_multigrid_set_fixed_iter(self.ptr, n)

    def setBottomSolver(self, s):
        lib.amrex_fi_multigrid_set_bottom_solver(self.ptr, {"smoother": 0, "bicgstab": 1, "cg": 2, "bicgcg": 5, "cgbicg": 6}[s])
        check()

    def solve(self, sol, rhs, tol_rel, tol_abs):
        r = lib.amrex_fi_multigrid_solve(self.ptr, _ptr_array(sol), _ptr_array(rhs), float(tol_rel), float(tol_abs))
        check()
        return r

    def getGradSolution(self, grads):
        """grads: per AMR level a list of 3 face-centred MultiFabs"""
        lib.amrex_fi_multigrid_get_grad_solution(self.ptr, _ptr_array([m for lev in grads for m in lev]))
        check()

    def getFluxes(self, fluxes):
        lib.amrex_fi_multigrid_get_fluxes(self.ptr, _ptr_array([m for lev in fluxes for m in lev]))
        check()

    def compResidual(self, res, sol, rhs):
        lib.amrex_fi_multigrid_comp_residual(self.ptr, _ptr_array(res), _ptr_array(sol), _ptr_array(rhs))
        check()

    def numIters(self):
        return lib.amrex_b200_multigrid_num_iters(self.ptr)

    def residualHistory(self):
        n = self.numIters()
        a = (C.c_double * max(n, 1))()
        lib.amrex_b200_multigrid_residual_history(self.ptr, a, n)
        return list(a)[:n]

    def cgIters(self):
        a = (C.c_int * 512)()
        n = lib.amrex_b200_multigrid_cg_iters(self.ptr, a, 512)
        return list(a)[:min(n, 512)]

    def initRHS(self): return lib.amrex_b200_multigrid_init_rhs(self.ptr)
    def initResidual(self): return lib.amrex_b200_multigrid_init_residual(self.ptr)

    def timers(self):
        t = (C.c_double * 3)()
        lib.amrex_b200_multigrid_timers(self.ptr, t)
        return list(t)


# ------------------------------------------------------------------------------------ host-only metadata helpers
def hierarchy(geom, ba, dm, nprocs, agglomeration=1, consolidation=1, max_coarsening_level=30, agg_grid_size=-1, con_grid_size=-1):
    """MG hierarchy (list over amr levels of list over mg levels of dict(boxes, dmap, domain)). No GPU needed."""
    h = lib.amrex_b200_hierarchy_new(len(geom), _ptr_array(geom), _ptr_array(ba), _ptr_array(dm), agglomeration, consolidation,
                                     max_coarsening_level, agg_grid_size, con_grid_size, nprocs)
    check()
    out = []
    for a in range(len(geom)):
        levs = []
        for m in range(lib.amrex_b200_hierarchy_num_mg_levels(h, a)):
            n = lib.amrex_b200_hierarchy_nboxes(h, a, m)
            boxes, pmap, dom = (C.c_int * (6 * n))(), (C.c_int * n)(), (C.c_int * 6)()
            lib.amrex_b200_hierarchy_level(h, a, m, boxes, pmap, dom)
            levs.append({"boxes": [tuple(boxes[6 * i:6 * i + 6]) for i in range(n)], "dmap": list(pmap), "domain": tuple(dom),
                         "safe_with_next": bool(lib.amrex_b200_hierarchy_shares_box_list(h, a, m, m + 1))
                         if m + 1 < lib.amrex_b200_hierarchy_num_mg_levels(h, a) else None})
        out.append(levs)
    lib.amrex_b200_hierarchy_delete(h)
    return out


def make_sfc(ba, nprocs):
    """Bucket (rank before the weight sort) of every box, DistributionMapping::makeSFC. No GPU needed."""
    n = ba.size()
    a = (C.c_int * n)()
    lib.amrex_b200_make_sfc(ba.ptr, nprocs, a)
    check()
    return list(a)


def _tags(n, buf):
    out = []
    for i in range(n):
        t = buf[15 * i:15 * i + 15]
        out.append({"dbox": tuple(t[0:6]), "sbox": tuple(t[6:12]), "dst": t[12], "src": t[13], "peer": t[14]})
    return out


def fb_tags(ba, dm, ng, cross, period, myproc, kind):
    n = lib.amrex_b200_fb_tags(ba.ptr, dm.ptr, ng, int(cross), _i3(period), myproc, kind, None, 0)
    check()
    buf = (C.c_int * (15 * max(n, 1)))()
    lib.amrex_b200_fb_tags(ba.ptr, dm.ptr, ng, int(cross), _i3(period), myproc, kind, buf, n)
    return _tags(n, buf)


def fb_face_links(ba, dm, period, myproc):
    """Face links of the cross-stencil one-ghost-cell FillBoundary as rank myproc sees it: list over its boxes (ascending
    global index) of 6 (linked local box or -1, shift) pairs, or None when the pattern has no face-link form."""
    n = lib.amrex_b200_fb_face_links(ba.ptr, dm.ptr, _i3(period), myproc, None, 0)
    check()
    if n < 0:
        return None
    buf = (C.c_int * (24 * max(n, 1)))()
    lib.amrex_b200_fb_face_links(ba.ptr, dm.ptr, _i3(period), myproc, buf, 6 * n)
    return [[(buf[4 * (6 * b + f)], tuple(buf[4 * (6 * b + f) + 1: 4 * (6 * b + f) + 4])) for f in range(6)] for b in range(n)]


def cpc_tags(ba_dst, dm_dst, ng_dst, ba_src, dm_src, ng_src, period, myproc, kind):
    n = lib.amrex_b200_cpc_tags(ba_dst.ptr, dm_dst.ptr, ng_dst, ba_src.ptr, dm_src.ptr, ng_src, _i3(period), myproc, kind, None, 0)
    check()
    buf = (C.c_int * (15 * max(n, 1)))()
    lib.amrex_b200_cpc_tags(ba_dst.ptr, dm_dst.ptr, ng_dst, ba_src.ptr, dm_src.ptr, ng_src, _i3(period), myproc, kind, buf, n)
    return _tags(n, buf)


class GMRESMLMG(_Obj):
    """GMRES preconditioned by MLMG V-cycles (amrex::GMRESMLMG, LinearSolvers/AMReX_GMRES_MLMG.H)."""
    _deleter = "amrex_b200_delete_gmres_mlmg"

    def __init__(self, mlmg):
        p = C.c_void_p()
        self._mlmg = mlmg
        lib.amrex_b200_new_gmres_mlmg(C.byref(p), mlmg.ptr)
        check()
        super().__init__(p)

    def setVerbose(self, v): lib.amrex_b200_gmres_mlmg_set_verbose(self.ptr, v)
    def setMaxIters(self, n): lib.amrex_b200_gmres_mlmg_set_max_iters(self.ptr, n)
    def usePrecond(self, f): lib.amrex_b200_gmres_mlmg_use_precond(self.ptr, int(bool(f)))
    def setPrecondNumIters(self, n): lib.amrex_b200_gmres_mlmg_set_precond_num_iters(self.ptr, n)
    def setPropertyOfZero(self, f): lib.amrex_b200_gmres_mlmg_set_property_of_zero(self.ptr, int(bool(f)))

    def setRestartLength(self, n):
        lib.amrex_b200_gmres_mlmg_set_restart_length(self.ptr, n)
        check()

    def solve(self, sol, rhs, tol_rel, tol_abs):
        lib.amrex_b200_gmres_mlmg_solve(self.ptr, sol.ptr, rhs.ptr, float(tol_rel), float(tol_abs))
        check()

    def numIters(self): return lib.amrex_b200_gmres_mlmg_num_iters(self.ptr)
    def status(self): return lib.amrex_b200_gmres_mlmg_status(self.ptr)
    def residualNorm(self): return lib.amrex_b200_gmres_mlmg_residual_norm(self.ptr)

    def residualHistory(self):
        a = (C.c_double * 4096)()
        n = lib.amrex_b200_gmres_mlmg_residual_history(self.ptr, a, 4096)
        return list(a)[:min(n, 4096)]


def write_plotfile(name, mfs, varnames, geoms, time=0.0, level_steps=None, ref_ratio=None):
    """amrex::WriteMultiLevelPlotfile: mfs[lev] has one component per variable name."""
    n = len(mfs)
    names = (C.c_char_p * len(varnames))(*[v.encode() for v in varnames])
    steps = (C.c_int * n)(*(level_steps or [0] * n))
    rr = (C.c_int * max(n - 1, 1))(*((ref_ratio or [2] * (n - 1)) + [0])[:max(n - 1, 1)])
    lib.amrex_fi_write_plotfile(str(name).encode(), n, _ptr_array(mfs), names, _ptr_array(geoms), float(time), steps, rr)
    check()
