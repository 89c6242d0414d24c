"""Generates tests/golden/ref_paper_meshes.npz: what the REFERENCE's own contact loops (oracle/_ref/libidp_ref_ipc.so)
return on real geometry of the paper examples (Projects/FEMShell/input/{bunny3K,hand}.obj), in a
normal-flow-like configuration (BASELINE configs[0]: every vertex displaced along its normal). The vertex / triangle
arrays of the two meshes are stored with the results so that the GPU test can run where /root/reference does not exist.

    python tests/golden/make_golden_paper.py
"""
import hashlib
import os
import sys

import numpy as np
import scipy.sparse as sp

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import lexsorted  # noqa: E402
from idp_b200 import meshgen  # noqa: E402
from oracle import ref_binding  # noqa: E402

INPUT = "/root/reference/Projects/FEMShell/input"
MESHES = ("bunny3K", "hand")
KAPPA = 1e5


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def load_obj(path):
    V, F = [], []
    for line in open(path):
        t = line.split()
        if not t:
            continue
        if t[0] == "v":
            V.append([float(x) for x in t[1:4]])
        elif t[0] == "f":
            F.append([int(x.split("/")[0]) - 1 for x in t[1:4]])
    return np.array(V, np.float64), np.array(F, np.int32)


def normal_flow_case(V, F):
    """rest shape = the mesh; current shape = every vertex moved inward along its normal by 20 % of the mean edge length
    (thin features approach each other); search direction = a further inward step of 2 mean edge lengths (forces a CCD hit)."""
    m0 = meshgen.SurfaceMesh(V, F)
    n = m0.vertex_normals()
    e = np.linalg.norm(V[m0.bedge[:, 0]] - V[m0.bedge[:, 1]], axis=1).mean()
    X = V - 0.2 * e * n
    m = meshgen.SurfaceMesh(X, F, X0=V)
    return m, np.ascontiguousarray(-2.0 * e * n), e


def main():
    assert ref_binding.build() and ref_binding.ipc_available()
    ref = ref_binding.ReferenceIPC()
    out = {}
    for name in MESHES:
        V, F = load_obj(os.path.join(INPUT, name + ".obj"))
        m, d, e = normal_flow_case(V, F)
        out[name + "/V"], out[name + "/F"], out[name + "/edge"] = V, F, np.float64(e)
        for k, f in enumerate((0.6, 1.2)):
            dh = f * e
            rows, info = ref.constraint_set(m, dh * dh)
            dup = (rows[:, 0] < 0) & (rows[:, 3] < 0)
            out["%s/cs%d/n" % (name, k)] = np.int64(len(rows))
            out["%s/cs%d/sorted_sha" % (name, k)] = sha(lexsorted(rows).astype(np.int32))
            out["%s/cs%d/merged_sha" % (name, k)] = sha(rows[dup].astype(np.int32))
        E, g, (tr, tc, tv) = ref.barrier(m, rows, info[:, 0], dh * dh, KAPPA, project_spd=True)
        N = 3 * m.nV
        H = sp.coo_matrix((tv, (tr, tc)), shape=(N, N)).tocsr()
        out[name + "/E"], out[name + "/g"] = np.float64(E), g
        out[name + "/H_probe"] = H @ np.random.default_rng(20260118).normal(size=N)
        d2, mn = ref.min_dist2(m, rows)
        out[name + "/dist2_sha"], out[name + "/min_dist2"] = sha(d2), np.float64(mn)
        out[name + "/ccd"] = np.array([ref.ccd(m, d, 1.0, 0.0), ref.ccd(m, 0.25 * d, 0.5, 0.0)])
        print(name, m.nV, m.nF, "rows", len(rows), "E", E, "ccd", out[name + "/ccd"])
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "ref_paper_meshes.npz"), **out)


if __name__ == "__main__":
    main()
