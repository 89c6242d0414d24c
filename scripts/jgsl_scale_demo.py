"""The device-resident Newton iteration at a size the device path is for: BASELINE configs[2] geometry (two nested geodesic
icospheres, 501,760 triangles at nu = 112, gap 5e-3) driven through the `JGSL` module's normal-flow time step -- constraint set,
barrier E / g / H with PSD projection, flow + mass terms, CSR assembly, Project_DBC, PCG solve, CCD line search -- with the
outer sphere's orientation flipped so that the normal flow closes the gap and the surfaces go into contact.
Prints one JSON line: wall time per frame and per Newton iteration, contact rows, PCG iterations.

usage: python scripts/jgsl_scale_demo.py [--nu 112] [--frames 2]
"""
import argparse
import json
import os
import re
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
from idp_b200 import meshgen  # noqa: E402
from jgsl_common import PRODUCT_DIR, build_product, read_counter, run_own_driver, write_obj  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nu", type=int, default=112)
    ap.add_argument("--frames", type=int, default=2)
    ap.add_argument("--mag", type=float, default=4e-3)
    a = ap.parse_args()
    build_product()
    mesh, _ = meshgen.nested_icospheres(nu=a.nu, gap=5e-3, jitter=1e-4)
    F = np.ascontiguousarray(mesh.btri[:, :3], np.int32).copy()
    half = len(F) // 2
    F[half:] = F[half:, ::-1]  # outer sphere: normals point inwards, so a positive normal flow moves the two surfaces together
    with tempfile.TemporaryDirectory() as tmp:
        obj = os.path.join(tmp, "spheres.obj")
        write_obj(obj, mesh.X, F)
        out = os.path.join(tmp, "out")
        os.environ["IDP_PROFILE"] = "1"  # the backend prints its per-operator wall clock at exit
        t0 = time.time()
        rc, log = run_own_driver(PRODUCT_DIR, obj, 0.5, a.mag, a.frames, out, timeout=3000)
        wall = time.time() - t0
        text = open(log).read()
        assert rc == 0, text[-3000:]
        c = read_counter(os.path.join(out, "counter.txt"))
        frame_s = [float(x) for x in re.findall(r"([0-9.]+) s since the previous flush", text)]
        pcg = [int(l.split()[4]) for l in text.splitlines() if l.startswith("linear solve")]
        mins = [float(l.split()[2].rstrip(",")) for l in text.splitlines() if l.startswith("minDist2 =")]
        prof = [l for l in text.splitlines() if l.startswith("[B200 backend]") and "inside the C ABI" in l]
    print(json.dumps({"workload": "nested icospheres nu=%d: %d triangles, %d vertices, normal flow through the JGSL module (B200 backend)" % (a.nu, len(F), mesh.nV),
                      "frames": int(len(c)), "pn_iterations_per_frame": c[:, 0].tolist(), "contact_rows_per_frame": c[:, 1].tolist(),
                      "wall_s_total": wall, "wall_s_per_frame_after_first": frame_s[1:], "pcg_iterations": pcg,
                      "ms_per_newton_iteration_last_frame": (1e3 * frame_s[-1] / c[-1, 0]) if len(frame_s) > 1 else None,
                      "min_dist2_min": min(mins) if mins else None,
                      "note": "frame times include writing the frame's .obj (0.2-0.3 s at this size) and the host driver's O(nV) algebra",
                      "c_abi_wall_clock_whole_run": prof[-1] if prof else None}))


if __name__ == "__main__":
    main()
