"""GPU parity tests of the elastic terms kept on the device, through the C ABI: membrane and hinge
energy / gradient and their PSD-projected Hessians assembled into the same device CSR as the barrier rows, the lumped mass
and the Dirichlet projection, against the oracle (oracle/orc_elastic.hpp + the barrier / system restatements).
Bar: 1e-10 relative (BASELINE.json's E/g/H tolerance)."""
import os
import sys

import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as spla

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from conftest import make_cases  # noqa: E402
from shell_np import first_fundamental_forms, hinges  # noqa: E402

pytestmark = pytest.mark.gpu
KAPPA = 1e5
RTOL = 1e-10


def _dense_blocks_to_coo(H, verts, n):
    """per-element dense Hessians (nE x 3k x 3k) over vertex lists (nE x k) -> scipy COO of size n"""
    k = verts.shape[1]
    dof = (3 * verts[:, :, None] + np.arange(3)[None, None, :]).reshape(len(verts), 3 * k)
    r = np.repeat(dof[:, :, None], 3 * k, axis=2).ravel()
    c = np.repeat(dof[:, None, :], 3 * k, axis=1).ravel()
    return sp.coo_matrix((H.ravel(), (r, c)), shape=(n, n))


@pytest.mark.parametrize("with_contact", [False, True])
def test_elastic_terms_match_oracle(lib_built, orc, with_contact):
    from idp_b200 import ContactContext
    for name, m, _d, dhats in make_cases()[:2]:
        rng = np.random.default_rng(11)
        F = np.ascontiguousarray(m.btri[:, :3], np.int32)
        X0 = m.X0.copy()
        X = m.X + rng.normal(0, 0.02 * np.sqrt(first_fundamental_forms(X0, F)[:, 0].mean()), m.X.shape) * (0 if with_contact else 1)
        dbc = (rng.uniform(size=m.nV) < 0.1).astype(np.uint8)
        ib = first_fundamental_forms(X0, F)
        thickness = 1e-2
        area = 0.5 * np.linalg.norm(np.cross(X0[F[:, 1]] - X0[F[:, 0]], X0[F[:, 2]] - X0[F[:, 0]]), axis=1)
        vol = area * thickness
        Eym, nu, h = 1e4, 0.4, 0.04
        lam, mu = Eym * nu / (1 - nu * nu), Eym / (2 * (1 + nu))
        st, info = hinges(X0, F)
        assert len(st) > 0
        info[:, 0] += rng.normal(0, 0.05, len(info))  # rest angles that differ from the current ones
        k = Eym * thickness ** 3 / (24 * (1 - nu * nu))
        mass = np.zeros(m.nV)
        np.add.at(mass, F.ravel(), np.repeat(area * thickness * 1000.0 / 3, 3))
        n = 3 * m.nV
        c = ContactContext(0)
        try:
            c.set_mesh(m.nV, m.bnode, m.bedge, m.btri, dbc)
            c.set_rest_positions(X0)
            c.set_positions(X)
            c.set_membrane(F, ib, vol, lam, mu, h)
            c.set_hinges(st, info, k, h)
            c.set_mass(mass)
            # energy and gradient
            mE, mg, mH, mact = orc.membrane_batch(X, F, ib, h * h * vol, lam, mu, dbc=dbc, project_spd=True)
            hE, hg, hH, hact = orc.hinge_batch(X, st, info, h * h * k, dbc=dbc, project_spd=True)
            assert 0 < mact.sum() < len(F) or dbc.sum() == 0
            E = c.elastic_energy(E0=0.25)
            assert abs(E - 0.25 - (mE.sum() + hE.sum())) <= RTOL * (np.abs(mE).sum() + np.abs(hE).sum())
            g = c.elastic_gradient()
            og = mg + hg
            assert np.abs(g - og).max() <= RTOL * np.abs(og).max(), (name, np.abs(g - og).max() / np.abs(og).max())
            # system matrix: [membrane][hinges][barrier rows] + M, then Project_DBC
            om = orc.mesh(X, X0, m.bnode, m.bedge, m.btri, dbc)
            dh2 = dhats[-1] ** 2
            if with_contact:
                c.constraint_set(dh2)
                rows, _ = c.get_constraints()
                assert len(rows) > 0
            else:
                rows = np.zeros((0, 4), np.int32)
                c.set_constraints(rows)
            ptr, col, val = c.barrier_hessian(dh2, KAPPA, project_spd=True)
            A = sp.csr_matrix((val, col, ptr), shape=(n, n))
            B = _dense_blocks_to_coo(mH, F, n).tocsr() + _dense_blocks_to_coo(hH, st, n).tocsr() + sp.diags(np.repeat(mass, 3))
            if len(rows):
                o = orc.barrier_hessian(om, rows, np.ones(len(rows)), dh2, KAPPA, project_spd=True)["csr"]
                B = B + sp.csr_matrix((o[2], o[1], o[0]), shape=(n, n))
            assert spla.norm(A - B) <= RTOL * spla.norm(B), (name, with_contact, spla.norm(A - B) / spla.norm(B))
            assert abs(A - A.T).max() <= 1e-12 * abs(A).max()
            assert all(np.all(np.diff(col[ptr[r]:ptr[r + 1]]) > 0) for r in range(0, n, 5))
            # positive definite after mass + projection: the PCG converges to the direct solution
            c.project_dbc()
            rhs = rng.normal(size=n)
            rhs[np.repeat(dbc.astype(bool), 3)] = 0
            sol, iters, res = c.solve_pcg(rhs, rel_tol=1e-12, max_iter=20000)
            p2, c2, v2 = c.get_hessian_csr()
            ref = spla.spsolve(sp.csr_matrix((v2, c2, p2), shape=(n, n)).tocsc(), rhs)
            assert np.abs(sol - ref).max() <= 1e-7 * np.abs(ref).max(), (iters, res)
            # run-to-run bit identity of the assembled values (deterministic summation order)
            _, _, val_b = c.barrier_hessian(dh2, KAPPA, project_spd=True)
            assert np.array_equal(val, val_b)
            # removing the terms gives back the plain system
            c.set_membrane(None, None, None, None, None, 0.0)
            c.set_hinges(None, None, 0.0, 0.0)
            assert c.elastic_energy() == 0.0
        finally:
            c.close()
