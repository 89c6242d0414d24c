"""Parity of the SHARDED path with the unsharded one, on ONE GPU, through the C ABI.

Two transports run the same sharded code (build_constraint_set / barrier_eval / ccd_step / min_dist2 with nranks > 1):
* idp_comm_init_local: P contexts of this process, one host thread each, collectives over peer memory -- the full
  sharded semantics (distributed duplicate merge with key routing, gather order, all-reduces);
* idp_set_shard: P communicator-less shards run one after the other -- partial results, checked as a partition.
The NCCL transport itself (one process per GPU) is covered by scripts/mgpu_check.py on a multi-GPU box and by the
`parity_ok` field of bench.py --gpus N; everything above the transport is what these tests pin.
"""
import os
import threading

import numpy as np
import pytest
import scipy.sparse as sp

from conftest import lexsorted

pytestmark = pytest.mark.gpu
KAPPA = 1e5


def _meshes():
    from idp_b200 import meshgen
    out = []
    m, d = meshgen.sheet_stack(n_sheets=4, nx=60, ny=50, h=4e-3, A=1.5e-3, extent=(0.24, 0.2))
    out.append(("sheets4x60x50", m, d, (2e-3) ** 2, 0.0))
    m, d = meshgen.nested_icospheres(nu=16, gap=4e-2, jitter=1e-3, seed=11)
    m.dbc[::7] = 1  # Dirichlet mask: all-DBC pairs are dropped in the broad phase
    out.append(("icospheres16_dbc_xi", m, d, 0.05 ** 2, 1e-3))
    return out


def _unsharded(mesh, direction, dh2, xi):
    from idp_b200 import ContactContext
    c = ContactContext(0)
    c.set_surface_mesh(mesh)
    n = c.constraint_set(dh2, xi)
    rows, info = c.get_constraints()
    r = {"n": n, "rows": rows, "info": info, "E": c.barrier_energy(dh2, KAPPA, xi), "g": c.barrier_gradient(dh2, KAPPA, xi)}
    ptr, col, val = c.barrier_hessian(dh2, KAPPA, xi)
    N = 3 * mesh.nV
    r["H"] = sp.csr_matrix((val, col, ptr), shape=(N, N))
    r["alpha"] = c.ccd_step(direction, 1.0, xi)
    r["ccd_pt"], r["ccd_ee"] = c.get_candidates(2), c.get_candidates(3)
    r["dist2"], r["min"] = c.min_dist2(xi)
    c.close()
    return r


def _run_group(P, mesh, direction, dh2, xi):
    """P ranks of an in-process group, one thread each; returns the per-rank result dicts."""
    from idp_b200 import ContactContext
    ctxs = [ContactContext(0) for _ in range(P)]
    ContactContext.comm_init_local(ctxs)
    res, errs = [None] * P, [None] * P
    N = 3 * mesh.nV

    def work(r):
        c = ctxs[r]
        try:
            c.set_surface_mesh(mesh)
            o = {"n": c.constraint_set(dh2, xi)}
            o["local_rows"], _ = c.get_constraints()
            o["rows"], o["info"] = c.gather_constraints()
            o["E"] = c.barrier_energy(dh2, KAPPA, xi)
            o["g"] = c.barrier_gradient(dh2, KAPPA, xi)
            ptr, col, val = c.barrier_hessian(dh2, KAPPA, xi)
            o["H"] = sp.csr_matrix((val, col, ptr), shape=(N, N))
            E2, nnz = c.barrier_all(dh2, KAPPA, xi)
            o["E_all"] = E2
            o["alpha"] = c.ccd_step(direction, 1.0, xi)
            o["ccd_pt"], o["ccd_ee"] = c.get_candidates(2), c.get_candidates(3)
            _, o["min_only"] = c.min_dist2(xi, want_all=False)
            o["local_dist2"], o["min"] = c.min_dist2(xi)
            _, _, o["dist2"] = c.gather_constraints(want_dist2=True)
            res[r] = o
        except BaseException as e:  # noqa: BLE001 - wake the peers instead of leaving them in a barrier
            errs[r] = e
            c.comm_abort()

    th = [threading.Thread(target=work, args=(r,)) for r in range(P)]
    for t in th:
        t.start()
    for t in th:
        t.join(timeout=600)
    alive = [t.is_alive() for t in th]
    for c in ctxs:
        if not any(alive):
            c.close()
    assert not any(alive), "sharded ranks hung"
    for e in errs:
        if e is not None:
            raise e
    return res


def _check_against(ref, res, P):
    # every rank returns the global constraint list, bit-identical to the unsharded one, in the same order
    for o in res:
        assert o["n"] == ref["n"]
        assert np.array_equal(o["rows"], ref["rows"]) and np.array_equal(o["info"], ref["info"])
        assert np.array_equal(o["dist2"], ref["dist2"]) and o["min"] == ref["min"] and o["min_only"] == ref["min"]
        assert o["alpha"] == ref["alpha"]
        assert abs(o["E"] - ref["E"]) <= 1e-12 * abs(ref["E"]) and o["E"] == res[0]["E"] and abs(o["E_all"] - ref["E"]) <= 1e-12 * abs(ref["E"])
        assert np.abs(o["g"] - ref["g"]).max() <= 1e-12 * np.abs(ref["g"]).max() and np.array_equal(o["g"], res[0]["g"])
    # the local lists partition the global one: group k of the global list = concatenation of the ranks' groups
    # (merged group: descending rank order, the owner ranges follow the vertex slabs)
    assert sum(len(o["local_rows"]) for o in res) == ref["n"]
    assert np.array_equal(lexsorted(np.concatenate([o["local_rows"] for o in res])), lexsorted(ref["rows"]))
    assert sum(len(o["local_dist2"]) for o in res) == ref["n"]
    # sum of the partial CSRs = the global Hessian
    H = res[0]["H"]
    for o in res[1:]:
        H = H + o["H"]
    assert abs(H - ref["H"]).max() <= 1e-12 * abs(ref["H"]).max()
    assert sum(o["H"].nnz for o in res) < 1.5 * ref["H"].nnz  # nearly disjoint patterns (vertex slabs)
    # CCD candidates: the shards' query ranges partition the candidate sets
    for k in ("ccd_pt", "ccd_ee"):
        assert np.array_equal(lexsorted(np.concatenate([o[k] for o in res])), ref[k])


@pytest.mark.parametrize("P", [2, 8])
def test_in_process_group_matches_unsharded(lib_built, P):
    for name, mesh, direction, dh2, xi in _meshes():
        ref = _unsharded(mesh, direction, dh2, xi)
        assert ref["n"] > 1000, name
        res = _run_group(P, mesh, direction, dh2, xi)
        _check_against(ref, res, P)


def test_in_process_group_row_merge_fallback(lib_built):
    """Meshes with more than 2^21 vertices cannot pack a PP/PE row into 64 bits: the 16-byte row merge is forced here;
    sharded, it replicates the rows (all-gather of the three groups) and every rank evaluates its vertex chunks."""
    name, mesh, direction, dh2, xi = _meshes()[0]
    os.environ["IDP_FORCE_ROW_MERGE"] = "1"
    try:
        ref = _unsharded(mesh, direction, dh2, xi)
        res = _run_group(2, mesh, direction, dh2, xi)
    finally:
        del os.environ["IDP_FORCE_ROW_MERGE"]
    for o in res:
        assert np.array_equal(o["rows"], ref["rows"]) and np.array_equal(o["local_rows"], ref["rows"])  # replicated
        assert np.array_equal(o["dist2"], ref["dist2"]) and o["min"] == ref["min"] and o["alpha"] == ref["alpha"]
        assert abs(o["E"] - ref["E"]) <= 1e-12 * abs(ref["E"])
        assert np.abs(o["g"] - ref["g"]).max() <= 1e-12 * np.abs(ref["g"]).max()
    H = res[0]["H"] + res[1]["H"]
    assert abs(H - ref["H"]).max() <= 1e-12 * abs(ref["H"]).max()


@pytest.mark.parametrize("P", [2, 8])
def test_communicatorless_shards_partition_the_work(lib_built, P):
    """idp_set_shard(r, P): P shards one after the other on one context. Direct rows partition exactly; duplicate PP/PE
    rows are merged per shard only, so their multiplicities add up; E, g, H are partial sums; step and distance minima."""
    from idp_b200 import ContactContext
    name, mesh, direction, dh2, xi = _meshes()[0]
    ref = _unsharded(mesh, direction, dh2, xi)
    N = 3 * mesh.nV
    c = ContactContext(0)
    c.set_surface_mesh(mesh)
    E, g, H, alpha, mn = 0.0, np.zeros((mesh.nV, 3)), None, 1.0, np.inf
    direct, merged, n_merged_rows = [], {}, 0
    for r in range(P):
        c.set_shard(r, P)
        c.constraint_set(dh2, xi)
        rows, _ = c.get_constraints()
        du = (rows[:, 0] < 0) & (rows[:, 3] < 0)
        direct.append(rows[~du])
        n_merged_rows += int(du.sum())
        for row in rows[du]:
            k = (int(row[0]), int(row[1]), int(row[2]))
            merged[k] = merged.get(k, 0) - int(row[3])
        E += c.barrier_energy(dh2, KAPPA, xi)
        g += c.barrier_gradient(dh2, KAPPA, xi)
        ptr, col, val = c.barrier_hessian(dh2, KAPPA, xi)
        Hr = sp.csr_matrix((val, col, ptr), shape=(N, N))
        H = Hr if H is None else H + Hr
        alpha = min(alpha, c.ccd_step(direction, 1.0, xi))
        d, m = c.min_dist2(xi)
        if len(d):
            mn = min(mn, m)
    c.close()
    rdu = (ref["rows"][:, 0] < 0) & (ref["rows"][:, 3] < 0)
    assert np.array_equal(lexsorted(np.concatenate(direct)), lexsorted(ref["rows"][~rdu]))
    want = {(int(a), int(b), int(cc)): -int(dd) for a, b, cc, dd in ref["rows"][rdu]}
    assert merged == want
    assert n_merged_rows > len(want)  # the same PP/PE key IS produced on different shards: the sharded path must route keys
    # a multiplicity-k row is k * (weight * b(d)); split over shards the sum is the same
    assert abs(E - ref["E"]) <= 1e-12 * abs(ref["E"])
    assert np.abs(g - ref["g"]).max() <= 1e-12 * np.abs(ref["g"]).max()
    assert alpha == ref["alpha"] and mn == ref["min"]
    # makePD is positively homogeneous, so the projected Hessians of a multiplicity split over shards add up as well
    assert abs(H - ref["H"]).max() <= 1e-10 * abs(ref["H"]).max()
