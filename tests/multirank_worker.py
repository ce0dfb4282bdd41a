"""One rank of a multi-GPU parity run (launched by tests/test_multirank_gpu.py through torch.distributed.run, one process
per GPU, NCCL).  Every rank runs the reference (oracle/_ref/ref_driver, CPU) on the case to get bit-identical inputs and the
reference's answer, builds the distributed problem through the C ABI (boxes spread by the SFC DistributionMapping, remote
halos over ncclSend/ncclRecv, agglomerated MG levels reached by ParallelCopy), solves, and compares ITS boxes of the
solution with the reference's.  Rank 0 writes the verdict as one JSON line to argv[2].

    python -m torch.distributed.run --nproc-per-node N tests/multirank_worker.py <case> <out.json>
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)

CASES = {
    # name: (prob_type, n_cell, max_grid_size, max_level, maxorder)
    "abeclap128": (2, 128, 32, 0, 2),        # 64 boxes: agglomeration / consolidation across ranks
    "abeclap128_g64": (2, 128, 64, 0, 3),    # 8 boxes of 64^3: fused smoother + remote halos
    "poisson_periodic64": (5, 64, 16, 0, 2),
    "poisson128": (1, 128, 64, 0, 2),
    "amr64": (2, 64, 32, 1, 3),              # two-level composite solve
    "amr128": (1, 128, 32, 1, 3),
}


def exchange_case(ab, world, rank, out_path):
    """Two-component halo exchange and ParallelCopy across ranks against the known answer (the reference's
    Tests/MultiPeriod/main.cpp:33-69: every ghost cell holds the value of its periodic image): pack -> ncclSend/Recv ->
    unpack with the components of a tag back to back in the buffer (AMReX_FBI.H:765-771)."""
    import torch
    import torch.distributed as dist
    n, mgs = 64, 16
    ab.Geometry.setup((0., 0., 0.), (1., 1., 1.), (1, 1, 0))
    geom = ab.Geometry((0, 0, 0), (n - 1,) * 3)
    ba = ab.BoxArray((0, 0, 0), (n - 1,) * 3).maxSize(mgs)
    dm = ab.DistributionMapping(ba)
    pmap = dm.pmap(ba.size())
    me = ab.lib.amrex_b200_myproc()
    f = lambda i, j, k, c: (i % n) + n * ((j % n) + n * k) + 0.5 * c * n ** 3      # periodic in x and y
    bad = 0
    for cross in (False, True):
        mf = ab.MultiFab(ba, dm, 2, 1)
        mf.setVal(-1.0, ng=1)
        boxes = ba.boxes()
        for g, b in enumerate(boxes):
            if pmap[g] != me:
                continue
            I, J, K = np.meshgrid(*[np.arange(b[d], b[d + 3] + 1) for d in range(3)], indexing="ij")
            for c in range(2):
                mf.upload(f(I, J, K, c), b[:3], comp=c)
        mf.fill_boundary(geom, cross=cross, comp=0, ncomp=2)
        for g, b in enumerate(boxes):
            if pmap[g] != me:
                continue
            glo = [v - 1 for v in b[:3]]
            shp = tuple(b[d + 3] - b[d] + 3 for d in range(3))
            I, J, K = np.meshgrid(*[np.arange(glo[d], glo[d] + shp[d]) for d in range(3)], indexing="ij")
            inside_z = (K >= 0) & (K <= n - 1)
            nout = (I < b[0]).astype(int) + (I > b[3]) + (J < b[1]) + (J > b[4]) + (K < b[2]) + (K > b[5])
            check = inside_z & ((nout <= 1) if cross else (nout >= 0))          # cross: faces only
            for c in range(2):
                # a region that is this box grown by one: only this fab's own ghost cells are read where it covers them last
                got = mf.download_fab(g, b, ng=1, comp=c)          # this fab's own ghost cells
                bad += int(np.count_nonzero(check & (got != f(I, J, K, c))))
    # ParallelCopy of both components onto another layout (boxes of 32^3, other owners)
    ba2 = ab.BoxArray((0, 0, 0), (n - 1,) * 3).maxSize(32)
    dm2 = ab.DistributionMapping(ba2)
    src = ab.MultiFab(ba, dm, 2, 0)
    for g, b in enumerate(ba.boxes()):
        if pmap[g] != me:
            continue
        I, J, K = np.meshgrid(*[np.arange(b[d], b[d + 3] + 1) for d in range(3)], indexing="ij")
        for c in range(2):
            src.upload(f(I, J, K, c), b[:3], comp=c)
    dst = ab.MultiFab(ba2, dm2, 2, 0)
    dst.setVal(-1.0)
    dst.parallel_copy(src, geom, scomp=0, dcomp=0, ncomp=2)
    pmap2 = dm2.pmap(ba2.size())
    for g, b in enumerate(ba2.boxes()):
        if pmap2[g] != me:
            continue
        I, J, K = np.meshgrid(*[np.arange(b[d], b[d + 3] + 1) for d in range(3)], indexing="ij")
        for c in range(2):
            bad += int(np.count_nonzero(dst.download(b[:3], tuple(b[d + 3] - b[d] + 1 for d in range(3)), comp=c) != f(I, J, K, c)))
    t = torch.tensor([float(bad)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t)
    if rank == 0:
        with open(out_path, "w") as fh:
            fh.write(json.dumps(dict(case="exchange2", world=world, mismatches=int(t.item()), comm_nranks=int(ab.lib.amrex_b200_nprocs()))) + "\n")


def main():
    case, out_path = sys.argv[1], sys.argv[2]
    if case == "exchange2":
        import torch
        import torch.distributed as dist
        import amrex_b200 as ab
        world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(local)
        ab.init(local)
        if world > 1:
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
            ab.comm_init_from_torch()
        exchange_case(ab, world, rank, out_path)
        if world > 1:
            dist.barrier()
            ab.lib.amrex_b200_comm_finalize()
            dist.destroy_process_group()
        return
    prob_type, n, mgs, max_level, maxorder = CASES[case]
    import torch
    import torch.distributed as dist
    import amrex_b200 as ab
    from common import build_problem, build_problem_amr, rel_maxdiff, run_ref

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    ab.init(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        ab.comm_init_from_torch()
    kw = dict(mode="solve", prob_type=prob_type, n_cell=n, max_grid_size=mgs, linop_maxorder=maxorder, agg_grid_size=32)
    if max_level:
        kw["max_level"] = max_level
    ref, dump = run_ref(dump=True, threads=max(1, (os.cpu_count() or 2) // max(world, 1)), **kw)
    if max_level:
        P = build_problem_amr(ab, prob_type, n, mgs, dump, max_level=max_level, maxorder=maxorder)
        sols, rhss = P["sol"], P["rhs"]
        bas, dms = P["ba"], P["dm"]
    else:
        fusion = int(os.environ["WORKER_FUSION"]) if "WORKER_FUSION" in os.environ else None    # 1: fused pass on boxes >= 32^3
        P = build_problem(ab, prob_type, n, mgs, dump, maxorder=maxorder, fusion=fusion)
        sols, rhss = [P["sol"]], [P["rhs"]]
        bas, dms = [P["ba"]], [P["dm"]]
    mlmg = ab.MLMG(P["op"])
    mlmg.setVerbose(0)
    mlmg.setMaxIter(100)
    ab.profile_enable(True)
    mlmg.solve(sols, rhss, 1e-10, 0.0)
    names = sorted(set(q[0] for q in ab.profile_report()))
    ab.profile_enable(False)
    me = ab.lib.amrex_b200_myproc()
    diff, refmax, nlocal = 0.0, 0.0, 0
    means = []
    for lev in range(max_level + 1):
        lo, refsol = dump[f"sol_lev{lev}"]
        pmap = dms[lev].pmap(bas[lev].size())
        for g, b in enumerate(bas[lev].boxes()):
            sl = tuple(slice(b[d] - lo[d], b[d + 3] - lo[d] + 1) for d in range(3))
            r = refsol[sl]
            refmax = max(refmax, float(np.max(np.abs(r))))
            if pmap[g] != me:
                continue
            nlocal += 1
            mine = P["sol"][lev].download(b[:3], r.shape) if max_level else P["sol"].download(b[:3], r.shape)
            if prob_type == 5:          # singular: compare up to the constant (means gathered below)
                means.append((float(mine.sum()), float(r.sum()), mine, r))
            else:
                diff = max(diff, float(np.max(np.abs(mine - r))))
    if prob_type == 5:
        t = torch.tensor([sum(m[0] for m in means), sum(m[1] for m in means)], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t)
        shift = (t[0].item() - t[1].item()) / float(n ** 3)
        for _, _, mine, r in means:
            diff = max(diff, float(np.max(np.abs(mine - shift - r))))
    t = torch.tensor([diff, float(nlocal)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    hist, rh = list(mlmg.residualHistory()), ref["history"]
    # deviation of the residual history in units of the bar of tests/test_solve_gpu.py: 1e-5 relative, or the rounding floor
    # 1e-13 * initial norm under every residual (summation orders differ), whichever is larger
    floor = 1e-13 * max(ref["rhsnorm0"], ref["resnorm0"])
    hist_rel = max((abs(a - b) / max(1e-5 * abs(b), floor) for a, b in zip(hist, rh)), default=0.0)
    res = dict(case=case, world=world, iters=mlmg.numIters(), ref_iters=ref["iters"], sol_rel_maxdiff=t[0].item() / max(refmax, 1e-300),
               history_max_rel_diff=hist_rel, history=hist, ref_history=rh, cg_iters=list(mlmg.cgIters()), ref_cg_iters=ref.get("cg_iters"),
               init_resnorm=mlmg.initResidual(), ref_init_resnorm=ref["resnorm0"], max_local_boxes=int(t[1].item()),
               kernels=names, comm_nranks=int(ab.lib.amrex_b200_nprocs()))
    if rank == 0:
        with open(out_path, "w") as fh:
            fh.write(json.dumps(res) + "\n")
    if world > 1:
        dist.barrier()
        ab.lib.amrex_b200_comm_finalize()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
