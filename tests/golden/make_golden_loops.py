"""Generates tests/golden/ref_loops.npz from the REFERENCE's own contact loops (oracle/_ref/libidp_ref_ipc.so = FEM/IPC.h
and Grid/SPATIAL_HASH.h compiled from /root/reference/Library by oracle/ref_shim/Makefile). Run in the authoring container:

    python tests/golden/make_golden_loops.py

For every small test mesh (tests/conftest.py make_cases) it stores what the reference's six operators return: the
constraint set per dHat (row count, SHA-256 of the lexicographically sorted rows and of the merged PP/PE group in the
reference's own order), and at the largest dHat the barrier energy, the full gradient, the projected Hessian applied to
a seeded vector (and its Frobenius norm), the per-row squared distances (SHA-256 + minimum), and the intersection-free
step for several directions / thicknesses. The fixtures travel to the GPU box, where /root/reference does not exist."""
import hashlib
import os
import sys

import numpy as np
import scipy.sparse as sp

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import lexsorted, make_cases  # noqa: E402
from oracle import ref_binding  # noqa: E402

KAPPA = 1e5
CCD_CONFIGS = ((1.0, 0.0, 1.0), (0.3, 0.0, 1.0), (4.0, 0.0, 1.0), (1.0, 1e-4, 0.7))  # (direction scale, thickness, input step)


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def probe_vector(n):
    return np.random.default_rng(20260118).normal(size=n)


def main():
    assert ref_binding.build() and ref_binding.ipc_available(), "oracle/_ref could not be built (needs /root/reference)"
    ref = ref_binding.ReferenceIPC()
    out = {}
    for name, m, d, dhats in make_cases():
        for k, dh in enumerate(dhats):
            rows, info = ref.constraint_set(m, dh * dh)
            dup = (rows[:, 0] < 0) & (rows[:, 3] < 0)
            out["%s/cs%d/n" % (name, k)] = np.int64(len(rows))
            out["%s/cs%d/sorted_sha" % (name, k)] = sha(lexsorted(rows).astype(np.int32))
            out["%s/cs%d/merged_sha" % (name, k)] = sha(rows[dup].astype(np.int32))
            out["%s/cs%d/info" % (name, k)] = info[0] if len(info) else np.zeros(2)
        # barrier on the reference's own rows, in the reference's own order
        dh = dhats[-1]
        w = info[:, 0]
        for spd in (0, 1):
            E, g, (tr, tc, tv) = ref.barrier(m, rows, w, dh * dh, KAPPA, project_spd=bool(spd))
            N = 3 * m.nV
            H = sp.coo_matrix((tv, (tr, tc)), shape=(N, N)).tocsr()
            out["%s/H%d_probe" % (name, spd)] = H @ probe_vector(N)
            out["%s/H%d_fro" % (name, spd)] = np.float64(np.sqrt((H.data ** 2).sum()))
            out["%s/H%d_nnz" % (name, spd)] = np.int64(H.nnz)
        out["%s/rows" % name] = rows.astype(np.int32)
        out["%s/E" % name] = np.float64(E)
        out["%s/g" % name] = g
        d2, mn = ref.min_dist2(m, rows, 1e-4)
        out["%s/dist2_sha" % name] = sha(d2)
        out["%s/min_dist2" % name] = np.float64(mn)
        out["%s/ccd" % name] = np.array([ref.ccd(m, d * s, a0, xi) for s, xi, a0 in CCD_CONFIGS])
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "ref_loops.npz"), **out)
    print("wrote", len(out), "arrays")


if __name__ == "__main__":
    main()
