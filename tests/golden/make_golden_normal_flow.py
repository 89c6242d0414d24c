"""Golden trace of BASELINE configs[0] (the paper's normal-flow example): the reference's UNCHANGED script
Projects/FEMShell/12-14_normal_flow.py + Python/Drivers run from the writable mirror (scripts/make_ref_mirror.sh) on the
`JGSL` module of this repository built with the REFERENCE's own CPU contact loops as its backend
(tests/host_shim/jgsl_ref/JGSL.so: FEM/IPC.h + Grid/SPATIAL_HASH.h + Math/CSR_MATRIX.h + FEM/Shell/MEMBRANE.h + BENDING.h +
FEM/FRICTION.h compiled from /root/reference).
With JGSL_REF_DRIVER=1 that build hands every time step to the REFERENCE's own Newton driver (tests/host_shim/libref_driver.so:
Advance_One_Step_IE_Discrete_Shell + Line_Search of FEM/Shell/IMPLICIT_EULER.h, Compute_IncPotential* of INC_POTENTIAL.h), so the
traces stored here are produced by the reference's scripts, driver and operators; only the storages, Eigen, the linear solver and
the module's set-up functions are this repository's. Stores the input mesh, counter.txt (PN iterations and contact # per time
step, Shell/IMPLICIT_EULER.h:857-864) and the final vertex positions. The B200 build of the same module must reproduce the trace (tests/test_gpu_jgsl_module.py).

The same for BASELINE configs[1] (the animation-fix example, Projects/FEMShell/16_fix_char_seq.py, unchanged): membrane + hinge
bending + inertia + barrier on wm2_15k following the first frames of Rumba_Dancing_unfixed -> fix_char_seq_trace.npz (the
script asks for 180 frames; the mirror holds the first 6 targets, so the run ends -- like the reference would -- when frame 7
cannot be read; the 6 completed steps are the trace).

Run in the authoring container only (needs /root/reference):  python tests/golden/make_golden_normal_flow.py [flow|seq|batch|components|cloth]
"""
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
MIRROR = os.path.join(ROOT, "baseline", "_ref", "IDP_mirror")
CASES = [("bunny3K", "0.5", "-5e-3", "50"), ("hand", "0.5", "5e-3", "3")]  # batch.py:36-43 and :16-23
# further lines of batch.py (:7-14 cat, :56-63 font_Tao, :46-53 feline, :97-101 second sequence) -> batch_lines_trace.npz
BATCH_CASES = [("cat", "0.5", "5e-3", "10"), ("font_Tao", "0.5", "5e-3", "10"), ("feline", "1", "-5e-3", "50")]


def read_obj(path):
    V, F = [], []
    for line in open(path):
        if line.startswith("v "):
            V.append([float(t) for t in line.split()[1:4]])
        elif line.startswith("f"):
            F.append([int(t.split("/")[0]) - 1 for t in line.split()[1:4]])
    return np.array(V, np.float64), np.array(F, np.int32)


def fix_char_seq(cwd, env, n_frames=6, seq="Rumba_Dancing_unfixed", save=True):
    # the script's output folder carries its arguments; without arguments it runs the Rumba sequence
    args = [] if seq == "Rumba_Dancing_unfixed" else [seq]
    folder = os.path.join(cwd, "output", "16_fix_char_seq", *args)
    subprocess.call(["rm", "-rf", folder])
    subprocess.call([sys.executable, "16_fix_char_seq.py"] + args, cwd=cwd, env=env, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    counter = np.array([[int(t) for t in l.split()] for l in open(os.path.join(folder, "counter.txt"))], np.int64)
    assert len(counter) == n_frames, counter
    V, F = read_obj(os.path.join(cwd, "input", "wm2_15k.obj"))
    out = {"rest/V": V, "rest/F": F, "counter": counter, "V_end": read_obj(os.path.join(folder, "shell%d.obj" % n_frames))[0],
           "V_start": read_obj(os.path.join(folder, "shell0.obj"))[0]}
    for f in range(1, n_frames + 1):
        out["frame%d/V" % f] = read_obj(os.path.join(cwd, "input", seq, "%d.obj" % f))[0]
    print("fix_char_seq", seq, "steps", len(counter), "PN iterations", counter[:, 0].sum(), "contact #", counter[:, 1].tolist())
    if save:
        np.savez_compressed(os.path.join(HERE, "fix_char_seq_trace.npz"), **out)
    return out


def normal_flow_case(cwd, out, mesh, smooth, mag, frames):
    folder = os.path.join(cwd, "output", "12-14_normal_flow", "%s_%s_%s_%s" % (mesh, smooth, mag, frames))
    if not (os.environ.get("GOLDEN_REUSE") and os.path.exists(os.path.join(folder, "shell%s.obj" % frames))):  # GOLDEN_REUSE=1: keep finished runs
        subprocess.call(["rm", "-rf", folder])
        env = dict(os.environ, PYTHONPATH=os.path.join(ROOT, "tests", "host_shim", "jgsl_ref"), OMP_NUM_THREADS="8", JGSL_REF_DRIVER="1")
        subprocess.check_call([sys.executable, "12-14_normal_flow.py", mesh, smooth, mag, frames], cwd=cwd, env=env, stdout=subprocess.DEVNULL)
    V, F = read_obj(os.path.join(cwd, "input", mesh + ".obj"))
    Vend, _ = read_obj(os.path.join(folder, "shell%s.obj" % frames))
    counter = np.array([[int(t) for t in l.split()] for l in open(os.path.join(folder, "counter.txt"))], np.int64)
    out[mesh + "/V"] = V
    out[mesh + "/F"] = F
    out[mesh + "/args"] = np.array([smooth, mag, frames])
    out[mesh + "/counter"] = counter
    out[mesh + "/V_end"] = Vend
    print(mesh, "steps", len(counter), "PN iterations", counter[:, 0].sum(), "last contact #", counter[-1, 1], flush=True)


def two_shells_friction():
    """Lagged friction with one coefficient per pair of components (muComp, Shell/IMPLICIT_EULER.h:435-438): two nested geodesic
    spheres (2,004 vertices), the outer one with inward normals so that the flow presses them together; mu = 0.1 inside a
    component, 0.6 between the two; the reference's driver and operators through tests/jgsl_driver/two_shells.py."""
    import tempfile
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from idp_b200 import meshgen
    from jgsl_common import REFLOOPS_DIR, read_counter, write_obj
    mesh, _ = meshgen.nested_icospheres(nu=10, gap=5e-3, jitter=1e-4)
    F = np.ascontiguousarray(mesh.btri[:, :3], np.int32)
    n1, half = mesh.nV // 2, len(F) // 2
    out = {"inner/V": mesh.X[:n1], "inner/F": F[:half], "outer/V": mesh.X[n1:], "outer/F": np.ascontiguousarray(F[half:, ::-1] - n1),
           "args": np.array(["0.5", "4e-3", "4", "0.1", "0.6", "2"])}  # smooth, magnitude, frames, mu same / cross component, friction iterations
    with tempfile.TemporaryDirectory() as tmp:
        for k in ("inner", "outer"):
            write_obj(os.path.join(tmp, k + ".obj"), out[k + "/V"], out[k + "/F"])
        env = dict(os.environ, PYTHONPATH=REFLOOPS_DIR, OMP_NUM_THREADS="8", JGSL_REF_DRIVER="1")
        log = subprocess.check_output([sys.executable, os.path.join(ROOT, "tests", "jgsl_driver", "two_shells.py"), os.path.join(tmp, "inner.obj"),
                                       os.path.join(tmp, "outer.obj")] + list(out["args"][:3]) + [os.path.join(tmp, "out")] + list(out["args"][3:]), env=env).decode()
        out["counter"] = read_counter(os.path.join(tmp, "out", "counter.txt"))
        out["V_end"] = read_obj(os.path.join(tmp, "out", "shell4.obj"))[0]
        out["friction_updates"] = np.array(log.count("friction updated Newton res"))
    print("two shells with per-component friction", out["counter"].tolist(), "friction updates", int(out["friction_updates"]))
    np.savez_compressed(os.path.join(HERE, "two_shells_friction_trace.npz"), **out)


def cloth_on_ball():
    """A 441-vertex cloth falling under gravity onto a 252-vertex ball that is a moving Dirichlet body, friction 0.3, 12 steps of
    Advance_One_Step_IE_Hinge through the reference's unchanged Python/Drivers (tests/jgsl_driver/cloth_on_ball.py) and the
    reference's own Newton driver and operators."""
    import tempfile
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from idp_b200 import meshgen
    from jgsl_common import REFLOOPS_DIR, run_cloth_on_ball
    n = 20
    g = np.linspace(-0.5, 0.5, n + 1)
    xx, zz = np.meshgrid(g, g, indexing="ij")
    V = np.stack([xx.ravel(), 0.30 + 1e-3 * np.random.default_rng(5).standard_normal(xx.size), zz.ravel()], 1)
    idx = np.arange((n + 1) ** 2).reshape(n + 1, n + 1)
    t1 = np.stack([idx[:-1, :-1], idx[:-1, 1:], idx[1:, :-1]], -1).reshape(-1, 3)
    t2 = np.stack([idx[1:, 1:], idx[1:, :-1], idx[:-1, 1:]], -1).reshape(-1, 3)
    Vb, Fb = meshgen.icosphere(5, 0.25)
    out = {"cloth/V": V, "cloth/F": np.concatenate([t1, t2]).astype(np.int32), "ball/V": Vb, "ball/F": np.ascontiguousarray(Fb[:, :3], np.int32),
           "args": np.array(["12", "0.3"])}  # frames, mu
    with tempfile.TemporaryDirectory() as tmp:
        rc, text, counter, Vend = run_cloth_on_ball(REFLOOPS_DIR, tmp, out, ref_driver=True)
        assert rc == 0, text[-2000:]
    out["counter"], out["V_end"], out["friction_updates"] = counter, Vend, np.array(text.count("friction updated Newton res"))
    print("cloth on ball", counter.tolist(), "friction updates", int(out["friction_updates"]))
    np.savez_compressed(os.path.join(HERE, "cloth_on_ball_trace.npz"), **out)


def main():
    subprocess.check_call([os.path.join(ROOT, "scripts", "make_ref_mirror.sh")])
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "tests", "host_shim"), "jgsl_ref/JGSL.so"])
    cwd = os.path.join(MIRROR, "Projects", "FEMShell")
    subprocess.check_call(["chmod", "-R", "u+w", cwd])
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    if which in ("seq", "all"):
        fix_char_seq(cwd, dict(os.environ, PYTHONPATH=os.path.join(ROOT, "tests", "host_shim", "jgsl_ref"), OMP_NUM_THREADS="8", JGSL_REF_DRIVER="1"))
    if which == "seq":
        return
    if which == "components":
        two_shells_friction()
        return
    if which == "cloth":
        cloth_on_ball()
        return
    if which == "batch":
        out = {}
        for case in BATCH_CASES:
            normal_flow_case(cwd, out, *case)
        kick = fix_char_seq(cwd, dict(os.environ, PYTHONPATH=os.path.join(ROOT, "tests", "host_shim", "jgsl_ref"), OMP_NUM_THREADS="8", JGSL_REF_DRIVER="1"),
                            n_frames=3, seq="Kick_unfixed", save=False)
        for k, v in kick.items():
            if not k.startswith("rest/"):  # the rest mannequin is in fix_char_seq_trace.npz already
                out["kick/" + k] = v
        np.savez_compressed(os.path.join(HERE, "batch_lines_trace.npz"), **out)
        return
    out = {}
    for mesh, smooth, mag, frames in CASES:
        normal_flow_case(cwd, out, mesh, smooth, mag, frames)
    # friction (mu = 0.3, two friction iterations): the paper scripts leave sim.mu at 0, so this case is driven by the repository's own
    # caller (tests/jgsl_driver/normal_flow.py, the same module calls with mu / fricIterAmt passed through) on the same checker build
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from jgsl_common import REFLOOPS_DIR, read_counter, run_own_driver, write_obj
    import tempfile
    with tempfile.TemporaryDirectory() as tmp:
        obj = os.path.join(tmp, "hand.obj")
        write_obj(obj, out["hand/V"], out["hand/F"])
        os.environ["JGSL_REF_DRIVER"] = "1"
        rc, log = run_own_driver(REFLOOPS_DIR, obj, "0.5", "5e-3", "3", os.path.join(tmp, "out"), mu=0.3, fric_iter=2)
        assert rc == 0, open(log).read()[-2000:]
        out["hand_friction/args"] = np.array(["0.5", "5e-3", "3", "0.3", "2"])
        out["hand_friction/counter"] = read_counter(os.path.join(tmp, "out", "counter.txt"))
        out["hand_friction/V_end"] = read_obj(os.path.join(tmp, "out", "shell3.obj"))[0]
        out["hand_friction/friction_updates"] = np.array(open(log).read().count("friction updated Newton res"))
        print("hand with friction", out["hand_friction/counter"].tolist(), "friction updates", int(out["hand_friction/friction_updates"]))
    np.savez_compressed(os.path.join(HERE, "normal_flow_trace.npz"), **out)


if __name__ == "__main__":
    main()
