"""CPU tests of the oracle restatement of the caller-side steps (oracle/orc_system.hpp): flow-term triplets, `+= M`,
Project_DBC and the std::map surface extraction, against independent numpy / scipy formulations."""
import numpy as np
import scipy.sparse as sp

from conftest import make_cases


def test_flow_mass_dbc_restatement_against_scipy(orc):
    name, m, _d, _ = make_cases()[1]
    F = np.ascontiguousarray(m.btri[:, :3], np.int32)
    rng = np.random.default_rng(1)
    vol = rng.uniform(0.5, 2.0, len(F))
    mass = rng.uniform(0.1, 1.0, m.nV)
    mass[::9] = 0.0
    dbc = (rng.uniform(size=m.nV) < 0.1).astype(np.uint8)
    h = 0.02
    om = orc.mesh(m.X, m.X0, m.bnode, m.bedge, m.btri, dbc)
    rows = np.zeros((0, 4), np.int32)
    n = 3 * m.nV
    ptr, col, val = orc.system_matrix(om, rows, np.ones(0), 1e-4, 1e5, 0.0, True, F, vol, h, mass, False)
    A = sp.csr_matrix((val, col, ptr), shape=(n, n))
    # independent formulation: per-axis graph Laplacian weighted by h vol / 6 (INC_POTENTIAL.h:323-339)
    L = sp.lil_matrix((m.nV, m.nV))
    for e, (a, b, c) in enumerate(F):
        w = h * vol[e] / 6
        for i, j in ((a, b), (a, c), (b, a), (b, c), (c, a), (c, b)):
            L[i, j] -= w
        for i in (a, b, c):
            L[i, i] += 2 * w
    B = sp.kron(L.tocsr(), sp.identity(3)) + sp.diags(np.repeat(mass, 3))
    assert abs(A - B).max() <= 1e-15 * abs(B).max()
    for r in range(n):                       # columns ascending, no duplicates
        assert np.all(np.diff(col[ptr[r]:ptr[r + 1]]) > 0)
    ptr2, col2, val2 = orc.system_matrix(om, rows, np.ones(0), 1e-4, 1e5, 0.0, True, F, vol, h, mass, True)
    assert np.array_equal(ptr, ptr2) and np.array_equal(col, col2)
    P = sp.csr_matrix((val2, col2, ptr2), shape=(n, n)).toarray()
    fixed = np.repeat(dbc.astype(bool), 3)
    want = B.toarray()
    want[fixed, :] = 0; want[:, fixed] = 0
    want[fixed, fixed] = 1                   # CSR_MATRIX.h:130-141 (only stored entries change; the diagonal is stored)
    assert np.abs(P - want).max() <= 1e-15 * np.abs(want).max()


def test_surface_restatement_against_numpy_mirror(orc):
    from idp_b200 import meshgen
    X, F = meshgen.icosphere(5)
    s = orc.surface(len(X), F, X)
    p = meshgen.find_surface_primitives(X, F)
    for k in ("bnode", "bedge", "btri"):
        assert np.array_equal(s[k], p[k]), k
    for k in ("BNArea", "BEArea", "BTArea"):
        assert np.allclose(s[k], p[k], rtol=1e-14, atol=0), k
    # the reference keeps the orientation of the FIRST directed edge and lists the pairs in std::map order
    assert np.array_equal(s["bedge"], s["bedge"][np.lexsort((s["bedge"][:, 1], s["bedge"][:, 0]))])


def test_oracle_system_matrix_matches_reference_csr_matrix(orc):
    """The oracle's triplet -> CSR, `+= M` and Project_DBC against the reference's own Math/CSR_MATRIX.h compiled in
    oracle/_ref (Eigen::SparseMatrix stand-in): pattern identical, values bit for bit -- on the triplets of a real
    constraint set (barrier rows with PSD projection + flow term)."""
    import pytest
    from oracle import ref_binding
    if not ref_binding.csr_available():
        pytest.skip("oracle/_ref/libidp_ref_csr.so not built (needs /root/reference at build time)")
    name, m, _d, dhats = make_cases()[1]
    F = np.ascontiguousarray(m.btri[:, :3], np.int32)
    rng = np.random.default_rng(4)
    vol = rng.uniform(0.5, 2.0, len(F))
    mass = rng.uniform(0.1, 1.0, m.nV)
    mass[::5] = 0.0
    dbc = (rng.uniform(size=m.nV) < 0.1).astype(np.uint8)
    om = orc.mesh(m.X, m.X0, m.bnode, m.bedge, m.btri, dbc)
    rows, _info, _, _ = orc.constraint_set(om, dhats[-1] ** 2)
    assert len(rows) > 100
    for use_mass, use_dbc in ((False, False), (True, False), (True, True)):
        ptr, col, val, (tr, tc, tv) = orc.system_matrix(om, rows, np.ones(len(rows)), dhats[-1] ** 2, 1e5, 0.0, True, F, vol, 0.02,
                                                        mass if use_mass else None, use_dbc, want_triplets=True)
        rptr, rcol, rval = ref_binding.ref_csr_system(3 * m.nV, tr, tc, tv, np.repeat(mass, 3) if use_mass else None, dbc if use_dbc else None, 3)
        assert np.array_equal(ptr, rptr) and np.array_equal(col, rcol), (use_mass, use_dbc)
        assert np.array_equal(val, rval), (use_mass, use_dbc, np.abs(val - rval).max())
