"""Full-size golden digests: the REFERENCE's own loops (oracle/_ref/libidp_ref_ipc.so = FEM/IPC.h + SPATIAL_HASH.h
compiled from /root/reference) and the oracle run ONCE, offline, on the BASELINE full-size geometries; only sha256
digests, counts and scalars are committed (tests/golden/ref_fullsize.json). Run where /root/reference is mounted:

    python tests/golden/make_golden_fullsize.py [config4] [config5crop] [config4crop]

* config4      sheets8x500 (4,000,000 triangles, BASELINE configs[3], the bench workload): constraint set (reference
               loops), per-row dist2 + min (reference loops), CCD step (reference loops), static + CCD candidate sets
               (oracle: the reference does not expose its candidate lists; the oracle's are pinned to the reference on the
               small fixtures, tests/test_ref_loops.py);
* config5crop  a 1,000,000-triangle crop of the config-5 CCD stress sheets (16 sheets, h = 2e-3): CCD steps of the
               step-filter sweep (reference loops) + candidate sets (oracle);
* config4crop  sheets8x160 (409,600 triangles): same as config4, a cheaper mid-size anchor.
The barrier triplets of the 29 M rows (67 GB in the reference's format) are not produced.
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
OUT = os.path.join(ROOT, "tests", "golden", "ref_fullsize.json")

from digests import pairs_digest, rows_digest  # noqa: E402


def quiet():
    sys.stdout.flush()
    devnull, saved = os.open(os.devnull, os.O_WRONLY), os.dup(1)
    os.dup2(devnull, 1)
    return devnull, saved


def loud(h):
    sys.stdout.flush()
    os.dup2(h[1], 1)
    os.close(h[0])


def config5_mesh(nx, ny, extent):
    from idp_b200 import meshgen
    h = 2e-3
    m, d = meshgen.sheet_stack(n_sheets=16, nx=nx, ny=ny, h=h, A=0.75e-3, seed=20260104, dir_sigma=1.0, dir_seed=20260105, extent=extent)
    d[:, 2] -= np.where((np.arange(len(d)) // ((nx + 1) * (ny + 1))) % 2 == 1, -1.0, 1.0) * 0.25 * h  # unit Gaussian part only
    return m, d, h


def count_vertex_pairs(rows, nV):
    """Number of distinct ordered vertex pairs (vi, vj) that appear together in a row = block non-zeros of the Hessian."""
    r = rows.astype(np.int64)
    v = np.where(r < 0, -r - 1, r)             # decoded vertex ids; -1 (unused slot) decodes to 0 and is masked below
    used = np.ones(r.shape, bool)
    plain = r[:, 0] < 0                         # PT / PE / PP rows: slots holding -1 are unused (PE: slot 3, PP: slots 2, 3)
    used[plain, 3] = r[plain, 3] >= 0
    used[plain, 2] = r[plain, 2] >= 0
    chunks = []
    for c0 in range(0, len(r), 4000000):
        vv, uu = v[c0:c0 + 4000000], used[c0:c0 + 4000000]
        keys = []
        for i in range(4):
            for j in range(i, 4):
                ok = uu[:, i] & uu[:, j]
                lo = np.minimum(vv[ok, i], vv[ok, j]); hi = np.maximum(vv[ok, i], vv[ok, j])
                keys.append(lo * nV + hi)
        chunks.append(np.unique(np.concatenate(keys)))
    allk = np.unique(np.concatenate(chunks))
    diag = int(((allk // nV) == (allk % nV)).sum())
    return 2 * (len(allk) - diag) + diag


def static_and_ccd(name, m, d, dhat, out, log):
    from oracle import ref_binding
    from oracle.binding import Oracle
    ref = ref_binding.ReferenceIPC()
    orc = Oracle("parity")
    om = orc.mesh(m.X, m.X0, m.bnode, m.bedge, m.btri, m.dbc)
    rec = {"triangles": int(m.nF), "vertices": int(m.nV), "dhat": dhat}
    t0 = time.time()
    h = quiet()
    rows, info = ref.constraint_set(m, dhat * dhat, cap=max(4000000, 10 * m.nF))
    loud(h)
    log("%s reference constraint set: %d rows, %.0f s" % (name, len(rows), time.time() - t0))
    t0 = time.time()
    dist2, mn = ref.min_dist2(m, rows)
    rec["constraint_set"] = rows_digest(rows, dist2)
    rec["min_dist2"] = float(mn)
    rec["info_all"] = [float(info[0, 0]), float(info[0, 1])] if len(info) and (info == info[0]).all() else None
    log("%s reference min dist: %.17g, %.0f s" % (name, mn, time.time() - t0))
    # barrier energy and gradient of the whole set through the reference's Compute_Barrier / _Gradient (kappa = 1e5, the
    # bench's), and the size of the Hessian pattern (9 x number of distinct ordered vertex pairs that share a row)
    t0 = time.time()
    E, g, _ = ref.barrier(m, rows, info[:, 0], dhat * dhat, 1e5, want_h=False)
    rec["barrier"] = {"kappa": 1e5, "E_reference": float(E), "g_l2": float(np.linalg.norm(g)), "g_l1": float(np.abs(g).sum()),
                      "g_sample_index": [int(i) for i in np.linspace(0, m.nV - 1, 64).astype(np.int64)],
                      "g_sample": [[float(x) for x in g[i]] for i in np.linspace(0, m.nV - 1, 64).astype(np.int64)]}
    log("%s reference E = %.17g, |g| = %.17g, %.0f s" % (name, E, rec["barrier"]["g_l2"], time.time() - t0))
    t0 = time.time()
    rec["barrier"]["nnz"] = int(9 * count_vertex_pairs(rows, m.nV))
    log("%s Hessian pattern nnz = %d, %.0f s" % (name, rec["barrier"]["nnz"], time.time() - t0))
    t0 = time.time()
    h = quiet()
    a = ref.ccd(m, d, 1.0, 0.0)
    loud(h)
    rec["ccd"] = {"alpha0": 1.0, "thickness": 0.0, "alpha_reference": float(a)}
    log("%s reference CCD: alpha %.17g, %.0f s" % (name, a, time.time() - t0))
    t0 = time.time()
    orows, _, cpt, cee = orc.constraint_set(om, dhat * dhat, want_cand=True)
    od = rows_digest(orows)
    assert od["rows_sha256"] == rec["constraint_set"]["rows_sha256"], "oracle rows differ from the reference's at full size"
    rec["static_candidates"] = {"pt": pairs_digest(cpt), "ee": pairs_digest(cee)}
    del orows, cpt, cee
    log("%s oracle constraint set + candidates (rows identical to the reference's): %.0f s" % (name, time.time() - t0))
    t0 = time.time()
    o = orc.ccd(om, d, 1.0, 0.0, want_cand=True)
    assert o["status"] == 0 and o["step"] == a, (o["step"], a)
    rec["ccd"]["alpha_oracle"] = float(o["step"])
    rec["ccd"]["step_after_clamp"] = float(o["step_after_clamp"])
    rec["ccd"]["candidates"] = {"pt": pairs_digest(o["cand_pt"]), "ee": pairs_digest(o["cand_ee"])}
    log("%s oracle CCD + candidates: %.0f s" % (name, time.time() - t0))
    out[name] = rec


def main():
    which = sys.argv[1:] or ["config4crop", "config5crop", "config4"]
    out = {}
    if os.path.exists(OUT):
        with open(OUT) as f:
            out = json.load(f)
    t_start = time.time()

    def log(s):
        print("[%6.0f s] %s" % (time.time() - t_start, s), file=sys.stderr, flush=True)

    def save():
        out["_generator"] = "tests/golden/make_golden_fullsize.py (reference loops: oracle/_ref/libidp_ref_ipc.so; candidates: oracle)"
        with open(OUT, "w") as f:
            json.dump(out, f, indent=1, sort_keys=True)

    import bench
    for w in which:
        if w == "config4":
            m, d, dh = bench.build_workload("sheets8x500")
            static_and_ccd("sheets8x500", m, d, dh, out, log)
        elif w == "config4crop":
            m, d, dh = bench.build_workload("sheets8x160")
            static_and_ccd("sheets8x160", m, d, dh, out, log)
        elif w == "config5crop":
            from oracle import ref_binding
            from oracle.binding import Oracle
            ref = ref_binding.ReferenceIPC()
            orc = Oracle("parity")
            m, d, h = config5_mesh(250, 125, (0.5, 0.25))
            assert m.nF == 1000000
            om = orc.mesh(m.X, m.X0, m.bnode, m.bedge, m.btri, m.dbc)
            rec = {"triangles": int(m.nF), "vertices": int(m.nV), "h": h, "sweep": []}
            for sigma_h, a0, xi in ((0.5, 1.0, 0.0), (0.5, 0.25, 1e-4), (1.0, 1.0, 1e-4), (4.0, 1.0, 0.0)):
                dd = np.ascontiguousarray(d * (sigma_h * h))
                t0 = time.time()
                hq = quiet()
                a = ref.ccd(m, dd, a0, xi)
                loud(hq)
                o = orc.ccd(om, dd, a0, xi, want_cand=True)
                assert o["status"] == 0 and o["step"] == a, (o["step"], a)
                rec["sweep"].append({"sigma_over_h": sigma_h, "alpha0": a0, "thickness": xi, "alpha_reference": float(a),
                                     "step_after_clamp": float(o["step_after_clamp"]),
                                     "candidates": {"pt": pairs_digest(o["cand_pt"]), "ee": pairs_digest(o["cand_ee"])}})
                log("config5crop sigma=%.1fh a0=%g xi=%g: alpha %.17g, %d+%d candidates, %.0f s"
                    % (sigma_h, a0, xi, a, len(o["cand_pt"]), len(o["cand_ee"]), time.time() - t0))
            out["ccd_stress_16x250x125"] = rec
        save()
    log("done")


if __name__ == "__main__":
    main()
