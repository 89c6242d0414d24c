"""Generates tests/golden/ref_pair_math.npz from the REFERENCE's own per-pair code (oracle/_ref, compiled from
/root/reference/Library/Math by oracle/ref_shim/Makefile). Run in the authoring container:

    python tests/golden/make_golden.py

Inputs are seeded; outputs are what the reference's generated g_*/H_* code, classifiers, mollifier, barrier scalars,
AABB predicates and additive CCD return for them."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_binding  # noqa: E402

assert ref_binding.build(), "oracle/_ref could not be built (needs /root/reference)"
ref = ref_binding.Reference()
rng = np.random.default_rng(20260117)
out = {}
NP = {0: 2, 1: 3, 2: 4, 3: 4, 6: 4}
for kind, n in NP.items():
    X = rng.normal(size=(48, n, 3))
    X[::4] *= 1e-2  # small scales as in the paper configs
    out["x_%d" % kind] = X
    out["d_%d" % kind] = np.array([ref.dist2(kind, x) for x in X])
    gh = [ref.grad_hess(kind, x) for x in X]
    out["g_%d" % kind] = np.array([a for a, _ in gh])
    out["H_%d" % kind] = np.array([b for _, b in gh])
# classification: random, near-degenerate and exactly degenerate configurations
T = rng.normal(size=(400, 4, 3))
T[50:100, 0] = T[50:100, 1] + 1e-6 * rng.normal(size=(50, 3))                       # point on a vertex
T[100:150, 3] = T[100:150, 2] + (T[100:150, 1] - T[100:150, 0])                     # exactly parallel edges
T[150:200, 3] = T[150:200, 2] + (T[150:200, 1] - T[150:200, 0]) * (1 + 1e-12)       # nearly parallel
T[200:250, 0] = 0.5 * (T[200:250, 1] + T[200:250, 2])                               # point on an edge
T[250:300, :, 2] = 0                                                                # coplanar
out["type_x"] = T
out["pt_type"] = np.array([ref.pt_type(x) for x in T])
out["ee_type"] = np.array([ref.ee_type(x) for x in T])
out["pt_unclassified"] = np.array([ref.dist2(4, x) for x in T])
out["ee_unclassified"] = np.array([ref.dist2(5, x) for x in T])
# mollifier on the reference's own derivative-test stencil and random ones
M = np.concatenate([np.array([[[0, 0, 0], [1, .1, 0], [0, 1.1, -.1], [0, .1, -1.1]]], float), rng.normal(size=(31, 4, 3))])
eps = np.concatenate([[10.0], rng.uniform(0.5, 20.0, 31)])
mo = [ref.mollifier(x, e) for x, e in zip(M, eps)]
out["moll_x"], out["moll_eps"] = M, eps
out["moll_e"] = np.array([a for a, _, _ in mo]); out["moll_g"] = np.array([b for _, b, _ in mo]); out["moll_H"] = np.array([c for _, _, c in mo])
out["moll_thr"] = np.array([ref.mollifier_threshold(x) for x in M])
# barrier scalars
D = np.concatenate([10.0 ** rng.uniform(-12, -4.01, 60), [9.99e-5, 1e-4]])
out["bar_d"] = D
out["bar"] = np.array([ref.barrier_scalar(d, 1e-4, 1e5) for d in D])
# AABB predicates incl. exact ties (gap == dist must pass: strict > rejects)
A = rng.normal(size=(200, 4, 3)); DA = 0.3 * rng.normal(size=(200, 4, 3)); dist = rng.uniform(0.01, 1.0, 200)
out["aabb_x"], out["aabb_d"], out["aabb_dist"] = A, DA, dist
out["aabb"] = np.array([[ref.aabb(k, x, d, t) for k in range(4)] for x, d, t in zip(A, DA, dist)])
# additive CCD: random motions plus constructed approaching pairs (point dropping onto a triangle, crossing edges)
for kind, name in ((0, "pt"), (1, "ee")):
    X = rng.normal(size=(160, 4, 3)); Dd = rng.normal(size=(160, 4, 3)) * rng.uniform(0.1, 3, (160, 1, 1))
    for i in range(0, 160, 2):
        if kind == 0:
            tri = X[i, 1:]
            n = np.cross(tri[1] - tri[0], tri[2] - tri[0]); n /= np.linalg.norm(n)
            X[i, 0] = rng.dirichlet([1, 1, 1]) @ tri + n * rng.uniform(0.05, 0.6)
            Dd[i] = 0.05 * rng.normal(size=(4, 3)); Dd[i, 0] -= n * rng.uniform(0.2, 2.0)
        else:
            mid = 0.5 * (X[i, 0] + X[i, 1])
            dirb = np.cross(X[i, 1] - X[i, 0], rng.normal(size=3)); dirb /= np.linalg.norm(dirb)
            off = np.cross(X[i, 1] - X[i, 0], dirb); off /= np.linalg.norm(off)
            X[i, 2] = mid - dirb + off * rng.uniform(0.05, 0.5); X[i, 3] = mid + dirb + off * rng.uniform(0.05, 0.5)
            Dd[i] = 0.05 * rng.normal(size=(4, 3)); Dd[i, 2:] -= off * rng.uniform(0.2, 2.0)
    res = [ref.accd(kind, x, d, 1.0, 0.1, th) for x, d, th in zip(X, Dd, np.where(np.arange(160) % 3 == 0, 1e-3, 0.0))]
    out["accd_%s_x" % name], out["accd_%s_d" % name] = X, Dd
    out["accd_%s_hit" % name] = np.array([h for h, _ in res]); out["accd_%s_toc" % name] = np.array([t for _, t in res])
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "ref_pair_math.npz"), **out)
print("wrote", len(out), "arrays")
