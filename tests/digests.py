"""Order-canonical digests of constraint sets / candidate sets, shared by the full-size golden generator
(tests/golden/make_golden_fullsize.py, run once where /root/reference is mounted) and the GPU tests / bench.py.

Canonical row order: [PT rows, lexicographic][EE + mollified rows, lexicographic][merged PP/PE rows AS PRODUCED] --
the reference's order inside the first two groups depends on unordered_set iteration (FEM/IPC.h:158-267, 371-564), the
merged group comes out of a std::map and is compared in the reference's own order (IPC.h:599-654).
"""
import hashlib

import numpy as np


def _lex(a):
    return np.lexsort(a.T[::-1]) if len(a) else np.zeros(0, np.int64)


def canonical_row_perm(rows):
    rows = np.asarray(rows, np.int32).reshape(-1, 4)
    a, d = rows[:, 0], rows[:, 3]
    pt = np.nonzero((a < 0) & (d >= 0))[0]
    ee = np.nonzero(a >= 0)[0]
    du = np.nonzero((a < 0) & (d < 0))[0]
    return np.concatenate([pt[_lex(rows[pt])], ee[_lex(rows[ee])], du]), (len(pt), len(ee), len(du))


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def rows_digest(rows, dist2=None):
    """-> dict(n, groups, rows_sha256[, dist2_sha256]) of the canonically ordered rows (and their per-row values)."""
    rows = np.asarray(rows, np.int32).reshape(-1, 4)
    perm, groups = canonical_row_perm(rows)
    out = {"n": int(len(rows)), "groups": [int(g) for g in groups], "rows_sha256": sha(rows[perm])}
    if dist2 is not None:
        out["dist2_sha256"] = sha(np.asarray(dist2, np.float64)[perm])
    return out


def pairs_digest(pairs):
    """(n, 2) int32 candidate pairs, already sorted lexicographically (idp_get_candidates order)."""
    pairs = np.asarray(pairs, np.int32).reshape(-1, 2)
    return {"n": int(len(pairs)), "sha256": sha(pairs)}
