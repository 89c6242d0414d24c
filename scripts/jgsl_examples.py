"""End-to-end timing of the two paper examples (BASELINE configs[0] and configs[1]) through the `JGSL` module: the B200 build
(idp_b200/jgsl/JGSL.so) and, beside it, the same module on the REFERENCE's own CPU contact loops (tests/host_shim/jgsl_ref,
test infrastructure; all host threads) on the same machine. One JSON line per example; wall clock of the whole driver process
(module import, mesh I/O, every frame), inputs from the committed fixtures.

usage: python scripts/jgsl_examples.py [--skip-reference] [--examples bunny3K,hand,rumba]
"""
import argparse
import json
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
from jgsl_common import (PRODUCT_DIR, REFLOOPS_DIR, SEQ_TRACE, TRACE, build_product, read_counter, run_own_driver, run_own_seq_driver, write_obj,  # noqa: E402
                         write_sequence)


def run(example, module_dir, tmp, threads):
    t0 = time.time()
    if example == "rumba":
        z = np.load(SEQ_TRACE)
        rest, seq, n = write_sequence(tmp, z)
        out = os.path.join(tmp, "out_" + os.path.basename(module_dir))
        t0 = time.time()
        rc, log = run_own_seq_driver(module_dir, rest, seq, n, out, threads=threads, timeout=7200)
    else:
        z = np.load(TRACE)
        obj = os.path.join(tmp, example + ".obj")
        write_obj(obj, z[example + "/V"], z[example + "/F"])
        smooth, mag, frames = z[example + "/args"]
        out = os.path.join(tmp, "out_%s_%s" % (example, os.path.basename(module_dir)))
        t0 = time.time()
        rc, log = run_own_driver(module_dir, obj, smooth, mag, frames, out, threads=threads, timeout=7200)
    wall = time.time() - t0
    assert rc == 0, open(log).read()[-2000:]
    c = read_counter(os.path.join(out, "counter.txt"))
    text = open(log).read()
    pcg = [int(l.split()[4]) for l in text.splitlines() if l.startswith("linear solve")]
    return dict(wall_s=wall, steps=int(len(c)), pn_iterations=int(c[:, 0].sum()), contacts_last_step=int(c[-1, 1]),
                linear_solver_iterations_mean=float(np.mean(pcg)) if pcg else 0.0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--skip-reference", action="store_true")
    ap.add_argument("--examples", default="bunny3K,hand,rumba")
    a = ap.parse_args()
    build_product()
    cores = os.cpu_count() or 1
    desc = {"bunny3K": "configs[0] 12-14_normal_flow.py bunny3K 0.5 -5e-3 50 (3,135 vertices)", "hand": "configs[0] 12-14_normal_flow.py hand 0.5 5e-3 3",
            "rumba": "configs[1] 16_fix_char_seq.py wm2_15k on Rumba_Dancing_unfixed, first 6 frames (12,811 vertices, 25,472 triangles, dHat 1e-2)"}
    for ex in a.examples.split(","):
        with tempfile.TemporaryDirectory() as tmp:
            line = {"example": ex, "what": desc[ex], "b200": run(ex, PRODUCT_DIR, tmp, "8")}
            if not a.skip_reference and os.path.exists(os.path.join(REFLOOPS_DIR, "JGSL.so")):
                line["reference_loops"] = dict(run(ex, REFLOOPS_DIR, tmp, str(cores)), cores=cores)
                line["speedup_wall"] = line["reference_loops"]["wall_s"] / line["b200"]["wall_s"]
            print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
