"""amrex_b200: B200-native (sm_100a) drop-in for AMReX's cell-centred MLMG solve path.

The product is the C++/CUDA shared library ``amrex_b200/lib/libamrex_b200.so`` (built by
``make -C amrex_b200/csrc`` or ``__graft_entry__.build()``).  This Python package is only a ctypes binding of its
C ABI (``include/amrex_b200_fi.h`` -- the reference's own ``amrex_fi_*`` interface, Src/F_Interfaces) used by the
tests, ``bench.py`` and launch scripts.  There is no Python or CPU compute path: if the library is missing, or no GPU
is visible when a device operation is requested, calls raise.
"""
from .capi import (  # noqa: F401
    lib, load_library, LIB_PATH, AmrexError, check,
    init, finalize, comm_init_from_torch, profile_enable, profile_report,
    Geometry, BoxArray, DistributionMapping, MultiFab, MLLinOp, MLABecLaplacian, MLALaplacian, MLPoisson, MLMG, GMRESMLMG,
    hierarchy, fb_tags, fb_face_links, cpc_tags, make_sfc, LinOpBCType, write_plotfile,
)
