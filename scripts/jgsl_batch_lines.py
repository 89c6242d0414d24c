"""The further lines of the reference's Projects/FEMShell/batch.py (cat 10 frames, font_Tao 10 frames, feline 50 frames, the first
frames of the Kick_unfixed sequence) through the B200 build of the `JGSL` module, compared with the traces the reference's
unchanged scripts + the reference's own Newton driver and operators produced on the CPU (tests/golden/batch_lines_trace.npz,
tests/golden/make_golden_normal_flow.py batch). One JSON line per example: wall clock and the deviations the GPU test bounds.

usage: python scripts/jgsl_batch_lines.py [--examples cat,font_Tao,feline,kick]
"""
import argparse
import json
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
from jgsl_common import PRODUCT_DIR, SEQ_TRACE, build_product, read_counter, read_obj, run_own_driver, run_own_seq_driver, write_obj, write_sequence  # noqa: E402

BATCH_TRACE = os.path.join(ROOT, "tests", "golden", "batch_lines_trace.npz")


def run_example(example, tmp, module_dir=PRODUCT_DIR):
    """-> dict of measured deviations from the golden trace of `example`"""
    z = np.load(BATCH_TRACE)
    out = os.path.join(tmp, "out_" + example)
    if example == "kick":
        rest = np.load(SEQ_TRACE)
        zz = {"rest/V": rest["rest/V"], "rest/F": rest["rest/F"], "counter": z["kick/counter"]}
        for k in z.files:
            if k.startswith("kick/frame"):
                zz[k[5:]] = z[k]
        rest_obj, seq, n = write_sequence(tmp, zz)
        t0 = time.time()
        rc, log = run_own_seq_driver(module_dir, rest_obj, seq, n, out, timeout=3000)
        golden, V0, Vg, last = z["kick/counter"], z["kick/V_start"], z["kick/V_end"], "shell%d.obj" % n
    else:
        obj = os.path.join(tmp, example + ".obj")
        write_obj(obj, z[example + "/V"], z[example + "/F"])
        smooth, mag, frames = z[example + "/args"]
        t0 = time.time()
        rc, log = run_own_driver(module_dir, obj, smooth, mag, frames, out, timeout=3000)
        golden, V0, Vg, last = z[example + "/counter"], z[example + "/V"], z[example + "/V_end"], "shell%s.obj" % frames
    wall = time.time() - t0
    text = open(log).read()
    assert rc == 0, text[-3000:]
    counter = read_counter(os.path.join(out, "counter.txt"))
    Vend, _ = read_obj(os.path.join(out, last))
    same = (counter == golden).all(axis=1) if counter.shape == golden.shape else np.zeros(0, bool)
    moved = float(np.median(np.linalg.norm(Vg - V0, axis=1)))
    dev = np.linalg.norm(Vend - Vg, axis=1)
    mins = [float(l.split()[2].rstrip(",")) for l in text.splitlines() if l.startswith("minDist2 =")]
    solves = [(int(l.split()[4]), float(l.split()[-1])) for l in text.splitlines() if l.startswith("linear solve")]
    return dict(linear_solves=len(solves), linear_iterations_max=max([s[0] for s in solves], default=0),
                linear_rel_residual_max=max([s[1] for s in solves], default=0.0), descent_fallbacks=text.count("gradient descent"),
                example=example, wall_s=round(wall, 2), steps=int(len(counter)), golden_steps=int(len(golden)),
                identical_leading_steps=int(len(same) if same.all() else np.argmin(same)) if len(same) else 0,
                pn_iterations=int(counter[:, 0].sum()), golden_pn_iterations=int(golden[:, 0].sum()),
                max_rel_contact_dev=float(np.max(np.abs(counter[:, 1] - golden[:, 1]) / np.maximum(golden[:, 1], 50.0))) if len(same) else None,
                max_iter_dev_per_step=int(np.max(np.abs(counter[:, 0] - golden[:, 0]))) if len(same) else None,
                median_dev_over_moved=float(np.median(dev) / moved), p99_dev_over_moved=float(np.quantile(dev, 0.99) / moved),
                max_dev_over_moved=float(dev.max() / moved), min_minDist2=min(mins) if mins else None,
                device_path="(B200 backend)" in text and "linear solve (device PCG)" in text,
                counter=counter.tolist(), golden_counter=golden.tolist())


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--examples", default="cat,font_Tao,feline,kick")
    a = ap.parse_args()
    build_product()
    for ex in a.examples.split(","):
        with tempfile.TemporaryDirectory() as tmp:
            print(json.dumps(run_example(ex, tmp)), flush=True)


if __name__ == "__main__":
    main()
