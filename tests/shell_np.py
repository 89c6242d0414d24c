"""numpy restatement of the rest-state quantities the elastic terms need (test helper): first fundamental forms and the
hinge stencils / rest angles / lengths / heights of Compute_Discrete_Shell_Inv_Basis<KL=false>
(Library/FEM/Shell/DISCRETE_SHELL.h:138-216), with the reference's stencil orientation (v0; v1, v2; v3): (v0, v1, v2) is the
triangle that owns the directed edge (v1, v2), (v3, v2, v1) its neighbour."""
import numpy as np


def first_fundamental_forms(X, F):
    e1, e2 = X[F[:, 1]] - X[F[:, 0]], X[F[:, 2]] - X[F[:, 0]]
    return np.stack([(e1 * e1).sum(1), (e1 * e2).sum(1), (e2 * e2).sum(1)], axis=1)


def dihedral(x0, x1, x2, x3):
    n1, n2 = np.cross(x1 - x0, x2 - x0), np.cross(x2 - x3, x1 - x3)
    a = np.arccos(np.clip(n1 @ n2 / np.sqrt((n1 @ n1) * (n2 @ n2)), -1.0, 1.0))
    return -a if np.cross(n2, n1) @ (x1 - x2) < 0 else a


def hinges(X, F):
    e2t = {}
    for t, f in enumerate(F.tolist()):
        for i in range(3):
            e2t[(f[i], f[(i + 1) % 3])] = t
    st, info = [], []
    seen = set()
    for (a, b) in sorted(e2t):
        if (b, a) not in e2t or (b, a) in seen:
            continue
        seen.add((a, b))
        f = F[e2t[(a, b)]].tolist()
        v0 = f[(f.index(b) + 1) % 3]
        o = F[e2t[(b, a)]].tolist()
        v3 = o[(o.index(a) + 1) % 3]
        st.append([v0, a, b, v3])
        l = np.linalg.norm(X[a] - X[b])
        n1, n2 = np.cross(X[a] - X[v0], X[b] - X[v0]), np.cross(X[b] - X[v3], X[a] - X[v3])
        info.append([dihedral(X[v0], X[a], X[b], X[v3]), l, (np.linalg.norm(n1) + np.linalg.norm(n2)) / (l * 6)])
    return np.array(st, np.int32).reshape(-1, 4), np.array(info, np.float64).reshape(-1, 3)
