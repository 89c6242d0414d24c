"""The oracle against the REFERENCE's own contact loops.

FEM/IPC.h and Grid/SPATIAL_HASH.h themselves -- not a restatement -- are compiled from /root/reference/Library into
oracle/_ref/libidp_ref_ipc.so (oracle/ref_shim: std::vector stand-ins for the Cabana storages, the repo's Eigen subset, an
inert pybind11). tests/golden/ref_loops.npz holds what the reference's six operators return on the small test meshes
(made by tests/golden/make_golden_loops.py); it travels to machines without /root/reference. Where the library is present
the comparison is also made live, including a Dirichlet mask and a thickness offset.

Bars: constraint sets (sorted) and the merged PP/PE group in the reference's own order bit for bit; energy, per-row
distances and the CCD step bit for bit (same expressions, same rounding); gradient and projected Hessian to 1e-12
(summation order / eigen-solver differ)."""
import hashlib
import os

import numpy as np
import pytest
import scipy.sparse as sp

from conftest import ROOT, lexsorted, make_cases

G = np.load(os.path.join(ROOT, "tests", "golden", "ref_loops.npz"))
KAPPA = 1e5
CCD_CONFIGS = ((1.0, 0.0, 1.0), (0.3, 0.0, 1.0), (4.0, 0.0, 1.0), (1.0, 1e-4, 0.7))


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def probe_vector(n):
    return np.random.default_rng(20260118).normal(size=n)


def omesh(orc, m):
    return orc.mesh(m.X, m.X0, m.bnode, m.bedge, m.btri, m.dbc)


@pytest.fixture(scope="module")
def cases():
    return make_cases()


def test_oracle_constraint_sets_match_reference_golden(orc, cases):
    for name, m, d, dhats in cases:
        om = omesh(orc, m)
        for k, dh in enumerate(dhats):
            rows, info, _, _ = orc.constraint_set(om, dh * dh)
            assert len(rows) == int(G["%s/cs%d/n" % (name, k)]), (name, dh)
            assert sha(lexsorted(rows).astype(np.int32)) == str(G["%s/cs%d/sorted_sha" % (name, k)]), (name, dh)
            dup = (rows[:, 0] < 0) & (rows[:, 3] < 0)
            assert sha(rows[dup].astype(np.int32)) == str(G["%s/cs%d/merged_sha" % (name, k)]), (name, dh)
            if len(info):
                assert np.array_equal(info[0], G["%s/cs%d/info" % (name, k)])


def test_oracle_barrier_min_dist_ccd_match_reference_golden(orc, cases):
    for name, m, d, dhats in cases:
        om = omesh(orc, m)
        dh = dhats[-1]
        rows = G["%s/rows" % name]
        w = np.ones(len(rows))
        _, E = orc.barrier(om, rows, w, dh * dh, KAPPA)
        _, g = orc.barrier_gradient(om, rows, w, dh * dh, KAPPA)
        assert E == float(G["%s/E" % name]), name                      # same row order, same serial accumulation (IPC.h:940)
        assert np.abs(g - G["%s/g" % name]).max() <= 1e-12 * np.abs(G["%s/g" % name]).max()
        N = 3 * m.nV
        for spd in (0, 1):
            ptr, col, val = orc.barrier_hessian(om, rows, w, dh * dh, KAPPA, project_spd=bool(spd))["csr"]
            H = sp.csr_matrix((val, col, ptr), shape=(N, N))
            ref = G["%s/H%d_probe" % (name, spd)]
            assert np.abs(H @ probe_vector(N) - ref).max() <= 1e-12 * np.abs(ref).max(), (name, spd)
            assert abs(np.sqrt((val ** 2).sum()) - float(G["%s/H%d_fro" % (name, spd)])) <= 1e-12 * float(G["%s/H%d_fro" % (name, spd)])
            assert H.nnz == int(G["%s/H%d_nnz" % (name, spd)])        # Eigen setFromTriplets keeps explicit zeros, as does the oracle
        d2, mn = orc.min_dist2(om, rows, 1e-4)
        assert sha(d2) == str(G["%s/dist2_sha" % name]) and mn == float(G["%s/min_dist2" % name])
        steps = np.array([orc.ccd(om, d * s, a0, xi)["step"] for s, xi, a0 in CCD_CONFIGS])
        assert np.array_equal(steps, G["%s/ccd" % name]), (name, steps, G["%s/ccd" % name])


def test_oracle_matches_reference_loops_live(orc, cases):
    from oracle import ref_binding
    if not ref_binding.ipc_available():
        pytest.skip("oracle/_ref/libidp_ref_ipc.so not built (needs /root/reference at build time)")
    ref = ref_binding.ReferenceIPC()
    name, m, d, dhats = cases[1]
    dbc = np.zeros(m.nV, np.uint8)
    dbc[: m.nV // 3] = 1
    m.dbc = dbc
    try:
        om = omesh(orc, m)
        for thickness in (0.0, 2e-3):
            rrows, rinfo = ref.constraint_set(m, dhats[-1] ** 2, thickness)
            orows, oinfo, _, _ = orc.constraint_set(om, dhats[-1] ** 2, thickness)
            assert len(rrows) == len(orows) > 0 and np.array_equal(lexsorted(rrows), lexsorted(orows))
            dup = (rrows[:, 0] < 0) & (rrows[:, 3] < 0)
            assert np.array_equal(rrows[dup], orows[(orows[:, 0] < 0) & (orows[:, 3] < 0)]) and np.array_equal(rinfo, oinfo)
            E, g, (tr, tc, tv) = ref.barrier(m, orows, oinfo[:, 0], dhats[-1] ** 2, KAPPA, thickness)
            _, oE = orc.barrier(om, orows, oinfo[:, 0], dhats[-1] ** 2, KAPPA, thickness)
            _, og = orc.barrier_gradient(om, orows, oinfo[:, 0], dhats[-1] ** 2, KAPPA, thickness)
            optr, ocol, oval = orc.barrier_hessian(om, orows, oinfo[:, 0], dhats[-1] ** 2, KAPPA, thickness)["csr"]
            N = 3 * m.nV
            A = sp.coo_matrix((tv, (tr, tc)), shape=(N, N)).tocsr()
            B = sp.csr_matrix((oval, ocol, optr), shape=(N, N))
            assert E == oE and np.abs(g - og).max() <= 1e-12 * np.abs(og).max() and abs(A - B).max() <= 1e-12 * abs(B).max()
            rd, rmn = ref.min_dist2(m, orows, thickness)
            od, omn = orc.min_dist2(om, orows, thickness)
            assert np.array_equal(rd, od) and rmn == omn
            assert ref.ccd(m, d, 1.0, thickness) == orc.ccd(om, d, 1.0, thickness)["step"]
    finally:
        m.dbc = np.zeros(m.nV, np.uint8)


# ---- real geometry of the paper examples: bunny3K and hand in a normal-flow-like configuration ----------
def paper_case(P, name):
    from idp_b200 import meshgen
    V, F, e = P[name + "/V"], P[name + "/F"], float(P[name + "/edge"])
    m0 = meshgen.SurfaceMesh(V, F)
    n = m0.vertex_normals()
    m = meshgen.SurfaceMesh(V - 0.2 * e * n, F, X0=V)
    return m, np.ascontiguousarray(-2.0 * e * n), e


@pytest.mark.parametrize("name", ["bunny3K", "hand"])
def test_oracle_matches_reference_on_paper_meshes(orc, name):
    P = np.load(os.path.join(ROOT, "tests", "golden", "ref_paper_meshes.npz"))
    m, d, e = paper_case(P, name)
    om = omesh(orc, m)
    for k, f in enumerate((0.6, 1.2)):
        dh = f * e
        rows, info, _, _ = orc.constraint_set(om, dh * dh)
        assert len(rows) == int(P["%s/cs%d/n" % (name, k)])
        assert sha(lexsorted(rows).astype(np.int32)) == str(P["%s/cs%d/sorted_sha" % (name, k)])
        dup = (rows[:, 0] < 0) & (rows[:, 3] < 0)
        assert sha(rows[dup].astype(np.int32)) == str(P["%s/cs%d/merged_sha" % (name, k)])
    # rows inside the PT / EE groups are in a different (sorted) order here, so sums agree to rounding, not bit for bit
    _, E = orc.barrier(om, rows, info[:, 0], dh * dh, KAPPA)
    _, g = orc.barrier_gradient(om, rows, info[:, 0], dh * dh, KAPPA)
    assert abs(E - float(P[name + "/E"])) <= 1e-12 * abs(float(P[name + "/E"]))
    assert np.abs(g - P[name + "/g"]).max() <= 1e-12 * np.abs(P[name + "/g"]).max()
    N = 3 * m.nV
    ptr, col, val = orc.barrier_hessian(om, rows, info[:, 0], dh * dh, KAPPA, project_spd=True)["csr"]
    ref = P[name + "/H_probe"]
    assert np.abs(sp.csr_matrix((val, col, ptr), shape=(N, N)) @ probe_vector(N) - ref).max() <= 1e-11 * np.abs(ref).max()
    d2, mn = orc.min_dist2(om, rows)
    assert mn == float(P[name + "/min_dist2"])
    steps = np.array([orc.ccd(om, d, 1.0, 0.0)["step"], orc.ccd(om, 0.25 * d, 0.5, 0.0)["step"]])
    assert np.array_equal(steps, P[name + "/ccd"]), (steps, P[name + "/ccd"])


def test_oracle_matches_reference_loops_on_bench_geometry(orc):
    """A crop of the bench sheets (same spacing, waves and dHat as BASELINE configs[3]): the regime where 98 % of the
    four-vertex rows are mollified PE / PP rows of parallel in-sheet edges. Live only (needs oracle/_ref)."""
    from idp_b200 import meshgen
    from oracle import ref_binding
    if not ref_binding.ipc_available():
        pytest.skip("oracle/_ref/libidp_ref_ipc.so not built (needs /root/reference at build time)")
    ref = ref_binding.ReferenceIPC()
    m, d = meshgen.sheet_stack(n_sheets=8, nx=24, ny=24, h=4e-3, A=1.5e-3, extent=(0.048, 0.048))
    dh = 2e-3
    om = omesh(orc, m)
    rrows, rinfo = ref.constraint_set(m, dh * dh)
    orows, oinfo, _, _ = orc.constraint_set(om, dh * dh)
    assert len(rrows) == len(orows) > 20000 and np.array_equal(lexsorted(rrows), lexsorted(orows))
    moll = (orows[:, 0] >= 0) & ((orows[:, 2] < 0) | (orows[:, 3] < 0))
    assert moll.sum() > 0.3 * len(orows)
    dup = (rrows[:, 0] < 0) & (rrows[:, 3] < 0)
    assert np.array_equal(rrows[dup], orows[(orows[:, 0] < 0) & (orows[:, 3] < 0)])
    E, g, (tr, tc, tv) = ref.barrier(m, orows, oinfo[:, 0], dh * dh, KAPPA)
    _, oE = orc.barrier(om, orows, oinfo[:, 0], dh * dh, KAPPA)
    _, og = orc.barrier_gradient(om, orows, oinfo[:, 0], dh * dh, KAPPA)
    optr, ocol, oval = orc.barrier_hessian(om, orows, oinfo[:, 0], dh * dh, KAPPA)["csr"]
    N = 3 * m.nV
    A = sp.coo_matrix((tv, (tr, tc)), shape=(N, N)).tocsr()
    B = sp.csr_matrix((oval, ocol, optr), shape=(N, N))
    assert E == oE and np.abs(g - og).max() <= 1e-12 * np.abs(og).max() and abs(A - B).max() <= 1e-12 * abs(B).max()
    assert ref.ccd(m, d, 1.0, 0.0) == orc.ccd(om, d, 1.0, 0.0)["step"]
