"""world_size-2 (gloo, CPU) test of the multi-rank halo-exchange plan: every rank builds ITS FillBoundary tag lists through
the C ABI (host-only entry points), packs with numpy, exchanges the per-peer buffers over gloo exactly as the device path
does over NCCL (one message per neighbour rank, tags concatenated in list order), unpacks, and every ghost cell must hold
the value of its periodic image -- the known-answer check of the reference's Tests/MultiPeriod/main.cpp:33-69."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, cross, q):
    try:
        sys.path.insert(0, REPO)
        os.environ["MASTER_ADDR"] = "127.0.0.1"
        os.environ["MASTER_PORT"] = str(port)
        dist.init_process_group("gloo", rank=rank, world_size=world)
        import amrex_b200 as ab
        n, mgs, ng = 32, 8, 1
        period = (n, n, 0)
        ab.Geometry.setup((0., 0., 0.), (1., 1., 1.), (1, 1, 0))
        ba = ab.BoxArray((0, 0, 0), (n - 1,) * 3).maxSize(mgs)
        dm = ab.DistributionMapping(ba, nprocs=world)
        pmap = dm.pmap(ba.size())
        boxes = ba.boxes()
        f = lambda i, j, k: (i % n) + n * ((j % n) + n * k)     # global linear index of the periodic image
        fabs = {}
        for g, b in enumerate(boxes):
            if pmap[g] != rank:
                continue
            a = np.full((mgs + 2,) * 3, -1.0)
            I, J, K = np.meshgrid(*[np.arange(b[d], b[d + 3] + 1) for d in range(3)], indexing="ij")
            a[1:-1, 1:-1, 1:-1] = f(I, J, K)
            fabs[g] = a

        def region(g, bx):
            b = boxes[g]
            return tuple(slice(bx[d] - b[d] + ng, bx[d + 3] - b[d] + ng + 1) for d in range(3))

        loc = ab.fb_tags(ba, dm, ng, cross, period, rank, 0)
        snd = ab.fb_tags(ba, dm, ng, cross, period, rank, 1)
        rcv = ab.fb_tags(ba, dm, ng, cross, period, rank, 2)
        reqs, rbufs = [], {}
        for peer in sorted({t["peer"] for t in rcv}):
            cnt = sum(np.prod([t["dbox"][d + 3] - t["dbox"][d] + 1 for d in range(3)]) for t in rcv if t["peer"] == peer)
            rbufs[peer] = torch.empty(int(cnt), dtype=torch.float64)
            reqs.append(dist.irecv(rbufs[peer], src=peer))
        for peer in sorted({t["peer"] for t in snd}):
            parts = [fabs[t["src"]][region(t["src"], t["sbox"])].ravel(order="F") for t in snd if t["peer"] == peer]
            reqs.append(dist.isend(torch.from_numpy(np.concatenate(parts)), dst=peer))
        for t in loc:
            fabs[t["dst"]][region(t["dst"], t["dbox"])] = fabs[t["src"]][region(t["src"], t["sbox"])]
        for r in reqs:
            r.wait()
        for peer, buf in rbufs.items():
            off = 0
            for t in [t for t in rcv if t["peer"] == peer]:
                shp = tuple(t["dbox"][d + 3] - t["dbox"][d] + 1 for d in range(3))
                cnt = int(np.prod(shp))
                fabs[t["dst"]][region(t["dst"], t["dbox"])] = buf[off:off + cnt].numpy().reshape(shp, order="F")
                off += cnt
        bad = 0
        for g, a in fabs.items():
            b = boxes[g]
            I, J, K = np.meshgrid(*[np.arange(b[d] - 1, b[d + 3] + 2) for d in range(3)], indexing="ij")
            inside = (K >= 0) & (K < n)                      # z is not periodic
            nout = (I < b[0]).astype(int) + (I > b[3]) + (J < b[1]) + (J > b[4]) + (K < b[2]) + (K > b[5])
            want = inside & ((nout <= 1) if cross else (nout >= 0))   # cross stencil guarantees faces only
            bad += int(np.sum(want & (a != f(I, J, K))))
        q.put((rank, bad, len(snd), len(rcv)))
        dist.barrier()
        dist.destroy_process_group()
    except Exception as e:  # pragma: no cover
        import traceback
        q.put((rank, f"EXC {e}\n{traceback.format_exc()}", 0, 0))


@pytest.mark.parametrize("cross", [True, False])
def test_fill_boundary_plan_two_ranks_gloo(cross):
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, cross, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    for rank, bad, ns, nr in res:
        assert bad == 0, (rank, bad)
        assert ns > 0 and nr > 0
