"""The restated Newton driver (idp_b200/host/jgsl/shell_flow.h) against the REFERENCE's own driver.

tests/host_shim/libref_driver.so is the reference's `Advance_One_Step_IE_Discrete_Shell` (FEM/Shell/IMPLICIT_EULER.h:151-891, with
Line_Search, Compute_IncPotential / _Gradient / _Hessian of INC_POTENTIAL.h and every operator header they include) compiled in
place from /root/reference against the stand-ins of oracle/ref_shim/include. The checker build of the module
(tests/host_shim/jgsl_ref/JGSL.so) hands each time step to it when JGSL_REF_DRIVER is set; the golden traces in tests/golden were
generated that way (tests/golden/make_golden_normal_flow.py). With the variable unset, the same build runs the restated driver on
the same reference operators: both must give the same counter.txt (PN iterations and contact # per step) and end state.
No GPU and no product code is involved here; this pins the host logic the B200 module shares with the checker build.
"""
import os
import re
import subprocess
import sys

import numpy as np
import pytest

from jgsl_common import (CLOTH_DRIVER, CLOTH_TRACE, MIRROR_PYTHON, REFLOOPS_DIR, ROOT, TRACE, TWO_SHELLS_TRACE, read_counter, read_obj, run_cloth_on_ball, run_own_driver,
                         run_two_shells, write_obj)

HAVE = os.path.exists(os.path.join(REFLOOPS_DIR, "JGSL.so")) and os.path.exists(os.path.join(ROOT, "tests", "host_shim", "libref_driver.so"))
pytestmark = pytest.mark.skipif(not HAVE, reason="reference driver / checker build absent (needs /root/reference)")


def _run(tmp_path, name, ref_driver, mu=None, fric_iter=None):
    z = np.load(TRACE)
    obj = str(tmp_path / "hand.obj")
    write_obj(obj, z["hand/V"], z["hand/F"])
    smooth, mag, frames = z["hand/args"]
    out = str(tmp_path / name)
    old = os.environ.pop("JGSL_REF_DRIVER", None)
    try:
        if ref_driver:
            os.environ["JGSL_REF_DRIVER"] = "1"
        rc, log = run_own_driver(REFLOOPS_DIR, obj, smooth, mag, frames, out, mu=mu, fric_iter=fric_iter)
    finally:
        os.environ.pop("JGSL_REF_DRIVER", None)
        if old is not None:
            os.environ["JGSL_REF_DRIVER"] = old
    assert rc == 0, open(log).read()[-2000:]
    Vend, _ = read_obj(os.path.join(out, "shell%s.obj" % frames))
    return read_counter(os.path.join(out, "counter.txt")), Vend, open(log).read()


def test_reference_driver_reproduces_the_golden_trace(tmp_path):
    """the golden really is the reference driver's output (regenerating it here gives the same bits)"""
    z = np.load(TRACE)
    counter, Vend, _ = _run(tmp_path, "ref", True)
    assert np.array_equal(counter, z["hand/counter"]), (counter.tolist(), z["hand/counter"].tolist())
    assert np.array_equal(Vend, z["hand/V_end"])


def test_restated_driver_matches_reference_driver_with_friction(tmp_path):
    """lagged friction (mu = 0.3, two friction iterations per step): restated driver == reference driver, row by row"""
    z = np.load(TRACE)
    mu, it = 0.3, 2
    c_ref, V_ref, _ = _run(tmp_path, "ref", True, mu=mu, fric_iter=it)
    c_own, V_own, _ = _run(tmp_path, "own", False, mu=mu, fric_iter=it)
    assert np.array_equal(c_ref, z["hand_friction/counter"]), (c_ref.tolist(), z["hand_friction/counter"].tolist())
    assert np.array_equal(c_own, c_ref), (c_own.tolist(), c_ref.tolist())
    assert np.array_equal(V_own, V_ref) and np.array_equal(V_ref, z["hand_friction/V_end"])


def test_restated_driver_matches_reference_driver_with_component_friction(tmp_path):
    """two components, one friction coefficient per pair of components (muComp table -> Compute_Friction_Coef, mu = 1,
    IMPLICIT_EULER.h:435-438): the golden is the reference driver's run; the restated driver on the reference's FRICTION.h gives the
    same rows, and the end state to the digits the obj files carry"""
    z = np.load(TWO_SHELLS_TRACE)
    rc, text, counter, Vend = run_two_shells(REFLOOPS_DIR, str(tmp_path), z, str(tmp_path / "own"))
    assert rc == 0, text[-2000:]
    assert np.array_equal(counter, z["counter"]), (counter.tolist(), z["counter"].tolist())
    assert text.count("friction updated Newton res") == int(z["friction_updates"]) > 0
    assert np.abs(Vend - z["V_end"]).max() <= 1e-10


@pytest.mark.skipif(not os.path.isdir(MIRROR_PYTHON), reason="mirror of the reference's Python/Drivers absent (scripts/make_ref_mirror.sh)")
def test_restated_driver_matches_reference_driver_on_cloth_on_ball(tmp_path):
    """Advance_One_Step_IE_Hinge beyond the paper scripts: gravity, a Dirichlet subset moving with a velocity (Step_Dirichlet every
    step, the ball), membrane + bending + inertia, friction in the elastic step; scene driven by the reference's unchanged
    Python/Drivers. The golden is the reference driver's run: the restated driver gives the same rows and the same end state bits."""
    z = np.load(CLOTH_TRACE)
    rc, text, counter, Vend = run_cloth_on_ball(REFLOOPS_DIR, str(tmp_path), z)
    assert rc == 0, text[-2000:]
    assert np.array_equal(counter, z["counter"]), (counter.tolist(), z["counter"].tolist())
    assert counter[-1, 1] > 0 and text.count("friction updated Newton res") == int(z["friction_updates"])
    assert np.array_equal(Vend, z["V_end"])


def _squeeze_log(folder, z, ref_driver, seconds):
    """cloth_on_ball.py with the fixed plate above the cloth, stopped after `seconds` (the squeezed step does not end)"""
    os.makedirs(folder, exist_ok=True)
    for k in ("cloth", "ball"):
        write_obj(os.path.join(folder, k + ".obj"), z[k + "/V"], z[k + "/F"])
    n = 6
    g = np.linspace(-0.3, 0.3, n + 1)
    xx, zz = np.meshgrid(g, g, indexing="ij")
    idx = np.arange((n + 1) ** 2).reshape(n + 1, n + 1)
    F = np.concatenate([np.stack([idx[:-1, :-1], idx[:-1, 1:], idx[1:, :-1]], -1).reshape(-1, 3), np.stack([idx[1:, 1:], idx[1:, :-1], idx[:-1, 1:]], -1).reshape(-1, 3)])
    write_obj(os.path.join(folder, "plate.obj"), np.stack([xx.ravel(), np.full(xx.size, 0.33), zz.ravel()], 1), F)
    env = dict(os.environ, PYTHONPATH=REFLOOPS_DIR, OMP_NUM_THREADS="8")
    env.pop("JGSL_REF_DRIVER", None)
    if ref_driver:
        env["JGSL_REF_DRIVER"] = "1"
    log = os.path.join(folder, "log.txt")
    with open(log, "w") as lf:
        try:
            subprocess.run([sys.executable, CLOTH_DRIVER, MIRROR_PYTHON, os.path.join(folder, "cloth.obj"), os.path.join(folder, "ball.obj"), "22", "0.0",
                            os.path.join(folder, "plate.obj")], cwd=folder, env=env, stdout=lf, stderr=subprocess.STDOUT, timeout=seconds)
        except subprocess.TimeoutExpired:
            pass
    text = open(log, errors="replace").read()
    kappa = re.findall(r"minDist2 = [0-9.e+-]+, kappa = ([0-9.e+-]+)", text)
    return kappa, re.findall(r"updated DBCStiff to ([0-9.e+-]+)", text), re.findall(r"PNIter(\d+): Newton res = ([0-9.]+)e", text)


@pytest.mark.skipif(not os.path.isdir(MIRROR_PYTHON), reason="mirror of the reference's Python/Drivers absent (scripts/make_ref_mirror.sh)")
def test_restated_driver_matches_reference_driver_when_squeezed(tmp_path):
    """The barrier-stiffness rule (kappa doubles when a row closer than 1e-9 gets closer, IMPLICIT_EULER.h:568-598) and the
    augmented-Lagrangian Dirichlet path (DBCStiff doubling) only show under extreme compression: the rising ball squeezes the cloth
    against a fixed plate. The squeezed step never converges (in the reference either), so both drivers run for a fixed time and the
    common prefix of their Newton iterations is compared: same kappa at every iteration, same stiffness updates."""
    z = np.load(CLOTH_TRACE)
    k_ref, d_ref, it_ref = _squeeze_log(str(tmp_path / "ref"), z, True, 18)
    k_own, d_own, it_own = _squeeze_log(str(tmp_path / "own"), z, False, 18)
    n = min(len(k_ref), len(k_own))
    if n < 135:
        pytest.skip("machine too slow to reach the squeezed step in the time given (%d Newton iterations)" % n)
    assert k_ref[:n] == k_own[:n]
    changes = [i for i in range(1, n) if k_own[i] != k_own[i - 1]]
    assert len(changes) >= 3, changes  # the rule fired, at the same iterations in both (the sequences are equal)
    m = min(len(d_ref), len(d_own))
    assert m >= 1 and d_ref[:m] == d_own[:m]
    # Newton iteration indices and the leading digits of the residuals agree over the prefix (the last printed digit may differ:
    # distances of 1e-9 are differences of O(1) coordinates)
    m = min(len(it_ref), len(it_own), 3 * n // 4)
    assert [a[0] for a in it_ref[:m]] == [a[0] for a in it_own[:m]]
    assert all(abs(float(a[1]) - float(b[1])) <= 2e-3 * max(1.0, float(a[1])) for a, b in zip(it_ref[:m], it_own[:m]))
