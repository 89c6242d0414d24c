"""CPU tests of the `JGSL` module: the module loads and exports the surface the normal-flow drivers
touch, the host-side set-up functions agree with independent numpy formulations of the reference's definitions
(FEM/Shell/DISCRETE_SHELL.h:218-395,555-577; FEM/BOUNDARY_CONDITION.h), and -- where the reference-loops checker build
exists -- the repository's own driver reproduces the golden trace produced by the reference's unchanged scripts."""
import importlib.util
import os
import subprocess
import sys

import numpy as np
import pytest

from jgsl_common import (DRIVER, PRODUCT_DIR, REFLOOPS_DIR, SEQ_TRACE, TRACE, build_product, compare_trace, read_counter, read_obj, run_own_driver,
                         run_own_seq_driver, write_obj, write_sequence)

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def J():
    path = build_product()
    spec = importlib.util.spec_from_file_location("JGSL", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)  # dlopen of libidp_contact.so only; no device is touched before the first time step
    return mod


def test_module_surface(J):
    for name in ["Kokkos_Initialize", "Set_Parameter", "Get_Parameter", "TIMER_FLUSH", "Scalard", "Vector2d", "Vector3d", "Vector4d", "Vector2i",
                 "Vector3i", "Vector4i", "Matrix2d", "Matrix3d", "Vector3f", "StdVectorXd", "StdVectorXi", "StdVectorVector2i", "StdVectorVector3i",
                 "StdVectorVector4i", "StdVectorVector3d", "StdMapPairiToi", "CSR_MATRIX_D", "FIXED_COROTATED_2", "FIXED_COROTATED_3", "Storage",
                 "MeshIO", "FEM"]:
        assert hasattr(J, name), name
    for name in ["V2dStorage", "V3dStorage", "V4dStorage", "V2iStorage", "V3iStorage", "V4iStorage", "SiStorage", "SdStorage", "V3dV3dV3dSdStorage",
                 "M2dM2dSdStorage", "M3dM3dSdStorage", "V2iV3dV3dV3dSdStorage"]:
        getattr(J.Storage, name)()
    for name in ["Add_Shell", "Initialize_Shell_Hinge_EIPC", "Initialize_OIPC", "Update_Normal_Flow_Neumann", "Advance_One_Step_IE_Flow", "Advance_One_Step_IE_Hinge"]:
        assert hasattr(J.FEM.DiscreteShell, name), name
    for name in ["Boundary_Dirichlet", "Init_Dirichlet", "Step_Dirichlet", "Turn_Dirichlet", "Reset_Dirichlet", "Load_Dirichlet"]:
        assert hasattr(J.FEM, name), name
    v = J.Vector3d(1, 2, 3) + J.Vector3d(1, 1, 1) * 2.0
    assert [v[0], v[1], v[2]] == [3.0, 4.0, 5.0] and abs(v.length2() - 50.0) < 1e-15
    J.Set_Parameter("Terminate", False)
    assert J.Get_Parameter("Terminate", True) is False and J.Get_Parameter("missing", 7) == 7
    with pytest.raises(NotImplementedError):  # outside the hosted path: named, not silently absent
        J.FEM.DiscreteShell.Advance_One_Step_SIE_Hinge()


def _setup(J, tmp_path, V, F, h=0.5):
    obj = str(tmp_path / "m.obj")
    write_obj(obj, V, F)
    X, Elem, rng = J.Storage.V3dStorage(), J.Storage.V3iStorage(), J.StdVectorXi()
    z = J.Vector3d(0, 0, 0)
    cnt = J.FEM.DiscreteShell.Add_Shell(obj, z, J.Vector3d(1, 1, 1), z, z, 0, X, Elem, rng)
    assert [cnt[i] for i in range(4)] == [0, 0, len(V), len(F)] and list(rng) == [len(V)]
    nodeAttr, M, b = J.Storage.V3dV3dV3dSdStorage(), J.CSR_MATRIX_D(), J.StdVectorXd()
    elemAttr, fcr, kappa = J.Storage.M2dM2dSdStorage(), J.FIXED_COROTATED_2.Create(), J.Vector3d(1e5, 0, 0)
    e2t, st, info = J.StdMapPairiToi(), J.StdVectorVector4i(), J.StdVectorVector3d()
    d2 = J.FEM.DiscreteShell.Initialize_Shell_Hinge_EIPC(1, 0, 0, 1, h, 1e-6, X, Elem, J.StdVectorVector2i(), e2t, st, info, nodeAttr, M, z, b, elemAttr, fcr,
                                                         kappa)
    return dict(X=X, Elem=Elem, nodeAttr=nodeAttr, M=M, b=b, kappa=kappa, e2t=e2t, st=st, info=info, d2=d2)


def test_setup_functions_against_numpy(J, tmp_path):
    z = np.load(TRACE)
    V, F = z["hand/V"], z["hand/F"]
    s = _setup(J, tmp_path, V, F)
    area = 0.5 * np.linalg.norm(np.cross(V[F[:, 1]] - V[F[:, 0]], V[F[:, 2]] - V[F[:, 0]]), axis=1)
    m = np.zeros(len(V))
    np.add.at(m, F.ravel(), np.repeat(area / 3, 3))  # rho = thickness = 1 (DISCRETE_SHELL.h:292)
    got = np.array([s["M"].coeff(3 * v, 3 * v) for v in range(len(V))])
    assert np.allclose(got, m, rtol=1e-13, atol=0)
    assert s["d2"] == 1.0 and len(s["e2t"]) == 3 * len(F)
    directed = {(int(t[i]), int(t[(i + 1) % 3])) for t in F for i in range(3)}
    interior = sum(1 for (a, b) in directed if a < b and (b, a) in directed)
    assert len(s["st"]) == interior  # one hinge per edge shared by two oppositely oriented triangles
    # hinge stencil (v0, v1, v2, v3): (v0,v1,v2) and (v3,v2,v1) are triangles of the mesh up to rotation
    tris = {tuple(np.roll(t, k)) for t in F.tolist() for k in range(3)}
    for k in range(0, len(s["st"]), 97):
        q = s["st"][k]
        assert (q[0], q[1], q[2]) in tris and (q[3], q[2], q[1]) in tris
    # Initialize_OIPC: kappa = 1e11 * mean(diag M) * 3 / (4e-16 * b''(1e-16)) with dHat2 = thickness^2 (DISCRETE_SHELL.h:566-573)
    d2 = J.FEM.DiscreteShell.Initialize_OIPC(0.0, 0.0, 1e-3, 0.0, s["M"], s["kappa"], 1)
    assert d2 == 1e-6
    d, dh = 1e-16, 1e-6
    Hb = (np.log(d / dh) * -2.0 - (d - dh) * 4.0 / d) + (d - dh) ** 2 / (d * d)
    assert abs(s["kappa"][0] - 1e11 * m.mean() * 3 / (4e-16 * Hb)) <= 1e-12 * s["kappa"][0] and s["kappa"][1] == 100 * s["kappa"][0]
    # Update_Normal_Flow_Neumann: b_v = mag * m_v * unit(area-weighted normal)
    J.FEM.DiscreteShell.Update_Normal_Flow_Neumann(s["X"], s["Elem"], s["M"], -0.02, s["b"])
    n = np.zeros_like(V)
    np.add.at(n, F.ravel(), np.repeat(np.cross(V[F[:, 1]] - V[F[:, 0]], V[F[:, 2]] - V[F[:, 0]]), 3, axis=0))
    want = -0.02 * m[:, None] * n / np.linalg.norm(n, axis=1)[:, None]
    assert np.allclose(np.array(list(s["b"])).reshape(-1, 3), want, rtol=1e-12, atol=1e-18)
    # Boundary_Dirichlet pins the end points of edges without an opposite twin, at their current position
    DBC = J.Storage.V4dStorage()
    J.FEM.Boundary_Dirichlet(s["X"], s["Elem"], DBC)
    rim = sorted({v for (a, b) in directed if (b, a) not in directed for v in (a, b)})
    assert [int(DBC.get(i)[0]) for i in range(len(DBC))] == rim
    for i in range(0, len(DBC), 7):
        assert [DBC.get(i)[k] for k in (1, 2, 3)] == V[rim[i]].tolist()


def test_obj_round_trip_and_dirichlet_scripting(J, tmp_path):
    V = np.array([[0, 0, 0], [1, 0, 0], [1, 1, 0], [0, 1, 0.25]], float)
    p = str(tmp_path / "q.obj")
    with open(p, "w") as f:
        for v in V:
            f.write("v %.17g %.17g %.17g\n" % tuple(v))
        f.write("f 1/1/1 2/2/2 3/3/3 4/4/4\n")  # a quad with texture / normal indices: split into (0,1,2), (0,2,3)
    X, E = J.Storage.V3dStorage(), J.Storage.V3iStorage()
    c = J.MeshIO.Read_TriMesh_Obj(p, X, E)
    assert [c[i] for i in range(4)] == [0, 0, 4, 2]
    assert [[E.get(i)[k] for k in range(3)] for i in range(2)] == [[0, 1, 2], [0, 2, 3]]
    J.MeshIO.Write_TriMesh_Obj(X, E, str(tmp_path / "w.obj"))
    V2, F2 = read_obj(str(tmp_path / "w.obj"))
    assert np.allclose(V2, V, rtol=1e-6) and F2.tolist() == [[0, 1, 2], [0, 2, 3]]
    DBC, mot = J.Storage.V4dStorage(), J.Storage.V2iV3dV3dV3dSdStorage()
    J.FEM.Init_Dirichlet(X, J.Vector3d(-0.1, -0.1, -0.1), J.Vector3d(0.5, 1.1, 1.1), J.Vector3d(0, 0, 1), J.Vector3d(0, 0, 0), J.Vector3d(0, 0, 1), 90.0,
                         DBC, mot, J.Vector4i(0, 0, 1000000000, -1))
    assert sorted(int(DBC.get(i)[0]) for i in range(len(DBC))) == [0, 3]
    J.FEM.Step_Dirichlet(mot, 1.0, DBC)  # quarter turn about z through the origin, then +1 in z
    got = {int(DBC.get(i)[0]): [DBC.get(i)[k] for k in (1, 2, 3)] for i in range(len(DBC))}
    assert np.allclose(got[3], [-1.0, 0.0, 1.25], atol=1e-15) and np.allclose(got[0], [0, 0, 1.0], atol=1e-15)
    J.FEM.Reset_Dirichlet(X, DBC)
    assert np.allclose([DBC.get(1)[k] for k in (1, 2, 3)], V[3])


@pytest.mark.skipif(not os.path.exists(os.path.join(REFLOOPS_DIR, "JGSL.so")), reason="reference-loops checker build absent (needs /root/reference)")
def test_own_driver_reproduces_reference_script_trace(tmp_path):
    """The golden trace was produced by the reference's unchanged 12-14_normal_flow.py + Python/Drivers; the repository's own
    caller (tests/jgsl_driver/normal_flow.py) on the same checker build must give the identical counter.txt and end state."""
    z = np.load(TRACE)
    obj = str(tmp_path / "hand.obj")
    write_obj(obj, z["hand/V"], z["hand/F"])
    smooth, mag, frames = z["hand/args"]
    rc, log = run_own_driver(REFLOOPS_DIR, obj, smooth, mag, frames, str(tmp_path / "out"))
    assert rc == 0, open(log).read()[-2000:]
    assert np.array_equal(read_counter(str(tmp_path / "out" / "counter.txt")), z["hand/counter"])
    Vend, _ = read_obj(str(tmp_path / "out" / ("shell%s.obj" % frames)))
    assert np.array_equal(Vend, z["hand/V_end"])


@pytest.mark.skipif(not (os.path.exists(os.path.join(REFLOOPS_DIR, "JGSL.so")) and os.path.exists(SEQ_TRACE)),
                    reason="reference-loops checker build / fixture absent (needs /root/reference)")
def test_own_seq_driver_reproduces_reference_script_trace(tmp_path):
    """Animation-fix example (membrane + hinge bending + barrier): the golden counter.txt came from the reference's unchanged
    16_fix_char_seq.py; the repository's own caller on the same all-reference checker build gives the same first two steps."""
    z = np.load(SEQ_TRACE)
    rest, seq, _n = write_sequence(str(tmp_path), z)
    rc, log = run_own_seq_driver(REFLOOPS_DIR, rest, seq, 2, str(tmp_path / "out"))
    assert rc == 0, open(log).read()[-2000:]
    assert np.array_equal(read_counter(str(tmp_path / "out" / "counter.txt")), z["counter"][:2])
