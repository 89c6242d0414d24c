"""A/B timing of idp_solve_pcg on a fixed system: sheets8x160 (409,600 triangles) with contact rows, lumped mass and a Dirichlet mask.
usage: python scripts/exp/pcg_time.py [path/to/libidp_contact.so ...]   (default: the in-tree library)"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from idp_b200 import ContactContext, contact  # noqa: E402

mesh, direction, dhat = bench.build_workload("sheets8x160")
F = np.ascontiguousarray(mesh.btri[:, :3], np.int32)
area = 0.5 * np.linalg.norm(np.cross(mesh.X[F[:, 1]] - mesh.X[F[:, 0]], mesh.X[F[:, 2]] - mesh.X[F[:, 0]]), axis=1)
mass = np.zeros(mesh.nV)
np.add.at(mass, F.ravel(), np.repeat(area * 1e-3 * 1000.0 / 3, 3))
rng = np.random.default_rng(1)
rhs = rng.normal(size=3 * mesh.nV)
for path in (sys.argv[1:] or [contact.LIB_PATH]):
    c = ContactContext(0, contact.load_library(path))
    c.set_surface_mesh(mesh)
    n = c.constraint_set(dhat * dhat)
    c.set_mass(mass)
    ptr, col, val = c.barrier_hessian(dhat * dhat, 1e5, project_spd=True)
    for iters in (200, 1000):
        c.solve_pcg(rhs, rel_tol=0.0, max_iter=iters)
        t0 = time.perf_counter()
        sol, it, res = c.solve_pcg(rhs, rel_tol=0.0, max_iter=iters)
        dt = time.perf_counter() - t0
        print("%s: rows %d nnz %d unknowns %d: %d iterations in %.2f ms = %.1f us / iteration (rel. residual %.3e)" %
              (os.path.basename(path), n, len(val), 3 * mesh.nV, it, 1e3 * dt, 1e6 * dt / it, res), flush=True)
    c.close()
