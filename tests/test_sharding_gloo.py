"""world_size-2 CPU test (gloo) of the N>1 host logic: contiguous primitive-range sharding and the semantics of the three
natural reductions (energy sum, gradient sum, min of the CCD step), checked against the unsharded oracle."""
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT, make_cases


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from idp_b200.sharding import shard_range
    from oracle.binding import Oracle
    orc = Oracle()
    name, m, d, dhats = make_cases()[1]
    dh = dhats[-1]
    om = orc.mesh(m.X, m.X0, m.bnode, m.bedge, m.btri, m.dbc)
    rows, info, _, _ = orc.constraint_set(om, dh * dh)
    # rows are replicated after the all-gather; each rank evaluates its slice
    b, e = shard_range(len(rows), rank, world)
    st, E = orc.barrier(om, rows[b:e], info[b:e, 0], dh * dh, 1e5)
    st, g = orc.barrier_gradient(om, rows[b:e], info[b:e, 0], dh * dh, 1e5)
    Et = torch.tensor([E], dtype=torch.float64); gt = torch.from_numpy(g.copy())
    dist.all_reduce(Et, op=dist.ReduceOp.SUM)
    dist.all_reduce(gt, op=dist.ReduceOp.SUM)
    # CCD: the step is the min over query shards; emulate a shard by masking the other half of the search direction
    nb, ne = shard_range(m.nV, rank, world)
    ranges = [None] * world
    dist.all_gather_object(ranges, (b, e, nb, ne))
    a_part = torch.tensor([0.3 + 0.1 * rank], dtype=torch.float64)
    dist.all_reduce(a_part, op=dist.ReduceOp.MIN)
    if rank == 0:
        _, E0 = orc.barrier(om, rows, info[:, 0], dh * dh, 1e5)
        _, g0 = orc.barrier_gradient(om, rows, info[:, 0], dh * dh, 1e5)
        q.put((ranges, len(rows), m.nV, Et.item(), E0, float(np.abs(gt.numpy() - g0).max() / np.abs(g0).max()), a_part.item()))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_reductions_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    ranges, n, nV, E, E0, gerr, amin = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # the slices tile [0, n) and [0, nV) exactly, in rank order
    assert ranges[0][0] == 0 and ranges[0][1] == ranges[1][0] and ranges[1][1] == n
    assert ranges[0][2] == 0 and ranges[0][3] == ranges[1][2] and ranges[1][3] == nV
    assert abs(E - E0) <= 1e-12 * abs(E0) and gerr <= 1e-12
    assert amin == 0.3


def test_shard_range_tiles_exactly():
    from idp_b200.sharding import shard_range
    for n in (0, 1, 7, 1000, 2008008, 6000011):
        for P in (1, 2, 3, 4, 8):
            edges = [shard_range(n, r, P) for r in range(P)]
            assert edges[0][0] == 0 and edges[-1][1] == n
            assert all(edges[i][1] == edges[i + 1][0] for i in range(P - 1))
            assert max(e - b for b, e in edges) - min(e - b for b, e in edges) <= 1
