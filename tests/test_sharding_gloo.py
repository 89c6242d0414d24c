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
    from sharding_np import row_owner, shard_range
    from oracle.binding import Oracle
    orc = Oracle()
    name, m, d, dhats = make_cases()[1]
    dh = dhats[-1]
    om = orc.mesh(m.X, m.X0, m.bnode, m.bedge, m.btri, m.dbc)
    rows, info, _, _ = orc.constraint_set(om, dh * dh)
    # rows are replicated after the all-gather; each rank evaluates the rows it owns (vertex-chunk ownership)
    b, e = shard_range(len(rows), rank, world)
    mine = row_owner(rows, m.nV, world) == rank
    st, E = orc.barrier(om, rows[mine], info[mine, 0], dh * dh, 1e5)
    st, g = orc.barrier_gradient(om, rows[mine], info[mine, 0], dh * dh, 1e5)
    Et = torch.tensor([E], dtype=torch.float64); gt = torch.from_numpy(g.copy())
    dist.all_reduce(Et, op=dist.ReduceOp.SUM)
    dist.all_reduce(gt, op=dist.ReduceOp.SUM)
    # CCD: the step is the min over query shards; emulate a shard by masking the other half of the search direction
    nb, ne = shard_range(m.nV, rank, world)
    ranges = [None] * world
    dist.all_gather_object(ranges, (b, e, nb, ne, int(mine.sum())))
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
    assert ranges[0][4] + ranges[1][4] == n and min(ranges[0][4], ranges[1][4]) > 0.2 * n  # every row owned once, both busy
    assert abs(E - E0) <= 1e-12 * abs(E0) and gerr <= 1e-12
    assert amin == 0.3


def test_row_owner_decodes_every_row_kind():
    from sharding_np import owner_shift, row_owner, row_vertices
    rows = np.array([[5, 6, 7, 8], [5, 6, -8, 9], [5, 7, 8, -10], [5, 7, -9, -10],
                     [-4, 10, 11, 12], [-4, 10, 11, -2], [-4, 10, -1, -3]], np.int32)
    v = row_vertices(rows)
    assert v.tolist() == [[5, 6, 7, 8], [5, 6, 7, 9], [5, 7, 8, 9], [5, 7, 8, 9], [3, 10, 11, 12], [3, 10, 11, -1], [3, 10, -1, -1]]
    assert owner_shift(2008008, 8) == 14 and owner_shift(12000, 2) == 9 and owner_shift(100, 8) == 8
    own = row_owner(np.array([[70000, 70001, 70002, 70003], [-1, 5, 6, 7]], np.int32), 2008008, 8)
    assert own.tolist() == [(70000 >> 14) % 8, 0]


def test_shard_range_tiles_exactly():
    from sharding_np import shard_range
    for n in (0, 1, 7, 1000, 2008008, 6000011):
        for P in (1, 2, 3, 4, 8):
            edges = [shard_range(n, r, P) for r in range(P)]
            assert edges[0][0] == 0 and edges[-1][1] == n
            assert all(edges[i][1] == edges[i + 1][0] for i in range(P - 1))
            assert max(e - b for b, e in edges) - min(e - b for b, e in edges) <= 1
