"""Helpers shared by the CPU and GPU tests of the `JGSL` module (idp_b200/jgsl/JGSL.so)."""
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PRODUCT_DIR = os.path.join(ROOT, "idp_b200", "jgsl")
REFLOOPS_DIR = os.path.join(ROOT, "tests", "host_shim", "jgsl_ref")
DRIVER = os.path.join(ROOT, "tests", "jgsl_driver", "normal_flow.py")
MIRROR = os.path.join(ROOT, "baseline", "_ref", "IDP_mirror", "Projects", "FEMShell")
TRACE = os.path.join(ROOT, "tests", "golden", "normal_flow_trace.npz")
SEQ_TRACE = os.path.join(ROOT, "tests", "golden", "fix_char_seq_trace.npz")
SEQ_DRIVER = os.path.join(ROOT, "tests", "jgsl_driver", "fix_char_seq.py")
CLOTH_DRIVER = os.path.join(ROOT, "tests", "jgsl_driver", "cloth_on_ball.py")
CLOTH_TRACE = os.path.join(ROOT, "tests", "golden", "cloth_on_ball_trace.npz")
MIRROR_PYTHON = os.path.join(ROOT, "baseline", "_ref", "IDP_mirror", "Python")
TWO_SHELLS_DRIVER = os.path.join(ROOT, "tests", "jgsl_driver", "two_shells.py")
TWO_SHELLS_TRACE = os.path.join(ROOT, "tests", "golden", "two_shells_friction_trace.npz")


def build_product():
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "idp_b200", "csrc"), "-j4"], stdout=subprocess.DEVNULL)
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "idp_b200", "host", "jgsl")], stdout=subprocess.DEVNULL)
    return os.path.join(PRODUCT_DIR, "JGSL.so")


def write_obj(path, V, F):
    with open(path, "w") as f:
        for v in V:
            f.write("v %.17g %.17g %.17g\n" % tuple(v))
        for t in F:
            f.write("f %d %d %d\n" % tuple(int(i) + 1 for i in t))


def read_obj(path):
    V, F = [], []
    for line in open(path):
        if line.startswith("v "):
            V.append([float(t) for t in line.split()[1:4]])
        elif line.startswith("f"):
            F.append([int(t.split("/")[0]) - 1 for t in line.split()[1:4]])
    return np.array(V, np.float64), np.array(F, np.int32)


def read_counter(path):
    return np.array([[int(t) for t in l.split()] for l in open(path)], np.int64)


def run_own_driver(module_dir, mesh_obj, smooth, mag, frames, out, threads="8", timeout=3000, mu=None, fric_iter=None):
    """tests/jgsl_driver/normal_flow.py in a fresh interpreter with `module_dir` first on the import path."""
    env = dict(os.environ, PYTHONPATH=module_dir, OMP_NUM_THREADS=threads)
    log = os.path.join(out, "log.txt")
    os.makedirs(out, exist_ok=True)
    with open(log, "w") as lf:
        rc = subprocess.call([sys.executable, DRIVER, mesh_obj, str(smooth), str(mag), str(frames), out] + ([str(mu)] if mu is not None else []) + ([str(fric_iter)] if fric_iter is not None else []), env=env,
                             stdout=lf, stderr=subprocess.STDOUT, timeout=timeout)
    return rc, log


def run_two_shells(module_dir, folder, z, out, threads="8", timeout=3000, ref_driver=False):
    """tests/jgsl_driver/two_shells.py on the fixture z (tests/golden/two_shells_friction_trace.npz) -> (rc, log text, counter, end positions)"""
    os.makedirs(out, exist_ok=True)
    for k in ("inner", "outer"):
        write_obj(os.path.join(folder, k + ".obj"), z[k + "/V"], z[k + "/F"])
    a = [str(t) for t in z["args"]]
    env = dict(os.environ, PYTHONPATH=module_dir, OMP_NUM_THREADS=threads)
    env.pop("JGSL_REF_DRIVER", None)
    if ref_driver:
        env["JGSL_REF_DRIVER"] = "1"
    log = os.path.join(out, "log.txt")
    with open(log, "w") as lf:
        rc = subprocess.call([sys.executable, TWO_SHELLS_DRIVER, os.path.join(folder, "inner.obj"), os.path.join(folder, "outer.obj")] + a[:3] + [out] + a[3:],
                             env=env, stdout=lf, stderr=subprocess.STDOUT, timeout=timeout)
    text = open(log).read()
    if rc != 0:
        return rc, text, None, None
    return rc, text, read_counter(os.path.join(out, "counter.txt")), read_obj(os.path.join(out, "shell%s.obj" % a[2]))[0]


def run_cloth_on_ball(module_dir, folder, z, threads="8", timeout=3000, ref_driver=False):
    """tests/jgsl_driver/cloth_on_ball.py (the reference's unchanged Python/Drivers from the mirror) on the fixture z, working
    directory `folder` -> (rc, log text, counter, end positions)"""
    os.makedirs(folder, exist_ok=True)
    for k in ("cloth", "ball"):
        write_obj(os.path.join(folder, k + ".obj"), z[k + "/V"], z[k + "/F"])
    frames, mu = [str(t) for t in z["args"]]
    env = dict(os.environ, PYTHONPATH=module_dir, OMP_NUM_THREADS=threads)
    env.pop("JGSL_REF_DRIVER", None)
    if ref_driver:
        env["JGSL_REF_DRIVER"] = "1"
    log = os.path.join(folder, "log.txt")
    with open(log, "w") as lf:
        rc = subprocess.call([sys.executable, CLOTH_DRIVER, MIRROR_PYTHON, os.path.join(folder, "cloth.obj"), os.path.join(folder, "ball.obj"), frames, mu],
                             cwd=folder, env=env, stdout=lf, stderr=subprocess.STDOUT, timeout=timeout)
    text = open(log).read()
    out = os.path.join(folder, "output", "cloth_on_ball", "run")
    if rc != 0:
        return rc, text, None, None
    return rc, text, read_counter(os.path.join(out, "counter.txt")), read_obj(os.path.join(out, "shell%s.obj" % frames))[0]


def write_sequence(folder, z):
    """rest mesh + target frames of the animation-fix fixture as the files the example reads"""
    os.makedirs(os.path.join(folder, "seq"), exist_ok=True)
    write_obj(os.path.join(folder, "rest.obj"), z["rest/V"], z["rest/F"])
    n = len(z["counter"])
    for f in range(1, n + 1):
        write_obj(os.path.join(folder, "seq", "%d.obj" % f), z["frame%d/V" % f], z["rest/F"])
    return os.path.join(folder, "rest.obj"), os.path.join(folder, "seq"), n


def run_own_seq_driver(module_dir, rest_obj, seq, frames, out, threads="8", timeout=3000):
    env = dict(os.environ, PYTHONPATH=module_dir, OMP_NUM_THREADS=threads)
    os.makedirs(out, exist_ok=True)
    log = os.path.join(out, "log.txt")
    with open(log, "w") as lf:
        rc = subprocess.call([sys.executable, SEQ_DRIVER, rest_obj, seq, str(frames), out], env=env, stdout=lf, stderr=subprocess.STDOUT, timeout=timeout)
    return rc, log


def run_reference_seq_script(module_dir, threads="8", timeout=3000):
    """The reference's UNCHANGED Projects/FEMShell/16_fix_char_seq.py from the mirror. It asks for 180 frames; the mirror holds the
    first targets only, so the process ends when a frame cannot be read (as the reference's would): the return code is not
    checked, the completed steps are."""
    folder = os.path.join(MIRROR, "output", "16_fix_char_seq")
    subprocess.call(["rm", "-rf", folder])
    env = dict(os.environ, PYTHONPATH=module_dir, OMP_NUM_THREADS=threads)
    subprocess.call([sys.executable, "16_fix_char_seq.py"], cwd=MIRROR, env=env, stdout=subprocess.DEVNULL, stderr=subprocess.STDOUT, timeout=timeout)
    return folder


def run_reference_script(module_dir, mesh, smooth, mag, frames, threads="8", timeout=3000):
    """The reference's UNCHANGED Projects/FEMShell/12-14_normal_flow.py from the mirror (scripts/make_ref_mirror.sh)."""
    folder = os.path.join(MIRROR, "output", "12-14_normal_flow", "%s_%s_%s_%s" % (mesh, smooth, mag, frames))
    subprocess.call(["rm", "-rf", folder])
    env = dict(os.environ, PYTHONPATH=module_dir, OMP_NUM_THREADS=threads)
    rc = subprocess.call([sys.executable, "12-14_normal_flow.py", mesh, smooth, mag, frames], cwd=MIRROR, env=env, stdout=subprocess.DEVNULL,
                         stderr=subprocess.STDOUT, timeout=timeout)
    return rc, folder


def compare_trace(counter, golden, exact_steps, rel_contacts=0.08, rel_iters=0.15):
    """The flow is chaotic in the last bits: two runs of the REFERENCE loops themselves with a different summation order agree
    exactly for the first ~14 steps of the bunny example and within a few percent after that, so: same number of steps, the
    first `exact_steps` rows identical, later contact counts within rel_contacts, total PN iterations within rel_iters."""
    assert counter.shape == golden.shape, (counter.shape, golden.shape)
    k = min(exact_steps, len(golden))
    assert np.array_equal(counter[:k], golden[:k]), (counter[:k].tolist(), golden[:k].tolist())
    c, g = counter[k:, 1].astype(float), golden[k:, 1].astype(float)
    if len(g):
        assert np.all(np.abs(c - g) <= rel_contacts * np.maximum(g, 50.0)), np.abs(c - g).max()
    assert abs(counter[:, 0].sum() - golden[:, 0].sum()) <= rel_iters * golden[:, 0].sum()
