"""GPU tests: the further lines of the reference's Projects/FEMShell/batch.py through the B200 build of the `JGSL` module --
12-14_normal_flow.py on cat (10 frames), font_Tao (10 frames) and feline (50 frames), and the first frames of the second
animation-fix sequence (16_fix_char_seq.py Kick_unfixed) -- against tests/golden/batch_lines_trace.npz, the traces of the
reference's unchanged scripts on the reference's own Newton driver and CPU operators (tests/golden/make_golden_normal_flow.py
batch). Same kind of bar as tests/test_gpu_jgsl_module.py: the flow is chaotic in the last bits, so the leading steps are
compared exactly and the rest within a few percent; every run stays intersection free. Two further scenes whose goldens come from the
reference's own driver: per-component friction (two shells, muComp table) and a cloth falling onto a moving Dirichlet ball."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "scripts"))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from jgsl_batch_lines import BATCH_TRACE, run_example  # noqa: E402
from jgsl_common import CLOTH_TRACE, MIRROR_PYTHON, PRODUCT_DIR, TWO_SHELLS_TRACE, build_product, run_cloth_on_ball, run_two_shells  # noqa: E402

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not os.path.exists(BATCH_TRACE), reason="batch_lines_trace.npz absent")]

# example: (identical leading steps, contact # within, PN iterations in total within, median and 99th-percentile end deviation / median
# motion). Measured on a B200 (profiles/r2_jgsl_batch_lines.jsonl): font_Tao identical in all 10 steps, feline in its first 11 (then
# within 2.1 % / 2.6 %), kick in 2 of 3 (third: one contact row of 46,598), cat -- an ill-conditioned system, 50 K PCG iterations per
# solve -- in its first 3 (then within 8.4 % / 4.6 %; the restated driver on the reference's CPU operators differs from the reference's
# driver by 4.7 % / 4.3 % on it). The bars leave a margin over that.
BARS = {"cat": (0, 0.20, 0.20, 0.03, 0.15), "font_Tao": (5, 0.02, 0.05, 0.01, 0.02), "feline": (5, 0.08, 0.10, 0.01, 0.10), "kick": (1, 0.01, 0.10, 0.01, 0.01)}


@pytest.mark.parametrize("example", sorted(BARS))
def test_b200_module_batch_line(tmp_path, example):
    build_product()
    lead, rel_contacts, rel_iters, med, p99 = BARS[example]
    r = run_example(example, str(tmp_path))
    assert r["device_path"] and r["descent_fallbacks"] == 0 and r["linear_rel_residual_max"] <= 1e-9, r  # every solve converged on the device
    assert r["steps"] == r["golden_steps"], r
    assert r["identical_leading_steps"] >= lead, r
    assert r["max_rel_contact_dev"] <= rel_contacts, r
    assert abs(r["pn_iterations"] - r["golden_pn_iterations"]) <= rel_iters * r["golden_pn_iterations"], r
    assert r["median_dev_over_moved"] <= med and r["p99_dev_over_moved"] <= p99, r
    assert r["min_minDist2"] is not None and r["min_minDist2"] > 0, r


@pytest.mark.skipif(not os.path.exists(TWO_SHELLS_TRACE), reason="two_shells_friction_trace.npz absent")
def test_b200_module_component_friction(tmp_path):
    """Two shells pressed together by the flow, lagged friction with the muComp table (0.1 inside a component, 0.6 across): the device
    path (idp_friction_set_components) against the trace of the reference's own driver and FRICTION.h. 18 K contact rows from the
    second step on; bars as for the single-component friction test."""
    build_product()
    z = np.load(TWO_SHELLS_TRACE)
    rc, text, counter, Vend = run_two_shells(PRODUCT_DIR, str(tmp_path), z, str(tmp_path / "out"))
    assert rc == 0, text[-3000:]
    assert "(B200 backend)" in text and "linear solve (device PCG)" in text
    g = z["counter"]
    assert counter.shape == g.shape and np.array_equal(counter[:2], g[:2]), (counter.tolist(), g.tolist())
    assert np.all(np.abs(counter[:, 1] - g[:, 1]) <= 0.03 * np.maximum(g[:, 1], 50)) and np.all(np.abs(counter[:, 0] - g[:, 0]) <= 3), (counter.tolist(), g.tolist())
    assert text.count("friction updated Newton res") >= 1
    V0 = np.concatenate([z["inner/V"], z["outer/V"]])
    moved = np.median(np.linalg.norm(z["V_end"] - V0, axis=1))
    dev = np.linalg.norm(Vend - z["V_end"], axis=1)
    assert np.median(dev) <= 0.01 * moved and np.quantile(dev, 0.99) <= 0.05 * moved, (np.median(dev), np.quantile(dev, 0.99), moved)
    mins = [float(l.split()[2].rstrip(",")) for l in text.splitlines() if l.startswith("minDist2 =")]
    assert mins and min(mins) > 0


@pytest.mark.skipif(not (os.path.exists(CLOTH_TRACE) and os.path.isdir(MIRROR_PYTHON)), reason="fixture / mirror of the reference's Python/Drivers absent")
def test_b200_module_cloth_on_ball(tmp_path):
    """The hinge time step beyond the paper scripts (gravity, a moving Dirichlet body, friction 0.3; the reference's unchanged
    Python/Drivers): B200 module against the trace of the reference's own driver and operators. A small, well-conditioned scene:
    contact counts and PN iterations per step equal or next to equal."""
    build_product()
    z = np.load(CLOTH_TRACE)
    rc, text, counter, Vend = run_cloth_on_ball(PRODUCT_DIR, str(tmp_path), z)
    assert rc == 0, text[-3000:]
    assert "(B200 backend)" in text and "linear solve (device PCG)" in text
    g = z["counter"]
    assert counter.shape == g.shape and np.array_equal(counter[:6], g[:6]), (counter.tolist(), g.tolist())
    assert np.all(np.abs(counter[:, 1] - g[:, 1]) <= np.maximum(3, 0.05 * g[:, 1])) and np.all(np.abs(counter[:, 0] - g[:, 0]) <= 3), (counter.tolist(), g.tolist())
    n_cloth = len(z["cloth/V"])
    V0 = np.concatenate([z["cloth/V"], z["ball/V"]])
    moved = np.median(np.linalg.norm(z["V_end"] - V0, axis=1))
    dev = np.linalg.norm(Vend - z["V_end"], axis=1)
    assert np.abs(Vend[n_cloth:] - z["V_end"][n_cloth:]).max() <= 1e-12  # the scripted ball
    assert np.median(dev) <= 0.01 * moved and np.quantile(dev, 0.99) <= 0.05 * moved, (np.median(dev), np.quantile(dev, 0.99), moved)
    mins = [float(l.split()[2].rstrip(",")) for l in text.splitlines() if l.startswith("minDist2 =")]
    assert mins and min(mins) > 0
