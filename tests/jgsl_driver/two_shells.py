"""Two shells, normal flow, lagged friction with one coefficient per pair of components (the muComp table of
FEM/Shell/IMPLICIT_EULER.h:435-438 and FEM/FRICTION.h Compute_Friction_Coef): the same module calls as normal_flow.py, with two
Add_Shell calls so that compNodeRange holds two components.

usage: two_shells.py inner.obj outer.obj smooth mag frames out mu_same mu_cross fricIterAmt
"""
import os
import sys

from JGSL import *  # noqa: F401,F403  (whichever build of the module is first on the import path)


def run(inner, outer, smooth, mag, frames, out, mu_same, mu_cross, fric_iter):
    os.makedirs(out, exist_ok=True)
    if not out.endswith("/"):
        out += "/"
    Kokkos_Initialize()
    Set_Parameter("Basic.log_folder", out)
    X, X0, Elem = Storage.V3dStorage(), Storage.V3dStorage(), Storage.V3iStorage()
    nodeAttr, massMatrix = Storage.V3dV3dV3dSdStorage(), CSR_MATRIX_D()
    elemAttr, elasticity = Storage.M2dM2dSdStorage(), FIXED_COROTATED_2.Create()
    DBC = Storage.V4dStorage()
    segs, edge2tri, edgeStencil, edgeInfo = StdVectorVector2i(), StdMapPairiToi(), StdVectorVector4i(), StdVectorVector3d()
    bodyForce, compNodeRange, muComp = StdVectorXd(), StdVectorXi(), StdVectorXd()
    tet, tetAttr, tetElasticity = Storage.V4iStorage(), Storage.M3dM3dSdStorage(), FIXED_COROTATED_3.Create()
    rod, rodInfo, rodHinge, rodHingeInfo = StdVectorVector2i(), StdVectorVector3d(), StdVectorVector3i(), StdVectorVector3d()
    stitchInfo, stitchRatio, particle = StdVectorVector3i(), StdVectorXd(), StdVectorXi()
    kappa, gravity = Vector3d(1e5, 0, 0), Vector3d(0, 0, 0)
    zero = Vector3d(0, 0, 0)

    for path in (inner, outer):
        FEM.DiscreteShell.Add_Shell(path, zero, Vector3d(1, 1, 1), zero, zero, 0, X, Elem, compNodeRange)
    for c1 in range(2):
        for c0 in range(2):
            muComp.append(mu_same if c0 == c1 else mu_cross)
    dt = smooth
    mag /= smooth * smooth
    MeshIO.Append_Attribute(X, X0)
    dHat2 = FEM.DiscreteShell.Initialize_Shell_Hinge_EIPC(1, 0, 0, 1, dt, 1e-6, X, Elem, segs, edge2tri, edgeStencil, edgeInfo, nodeAttr, massMatrix,
                                                          gravity, bodyForce, elemAttr, elasticity, kappa)
    FEM.Boundary_Dirichlet(X, Elem, DBC)
    dHat2 = FEM.DiscreteShell.Initialize_OIPC(0.0, 0.0, 1e-3, 0.0, massMatrix, kappa, 1)
    MeshIO.Write_TriMesh_Obj(X, Elem, out + "shell0.obj")
    total = 0
    for f in range(1, frames + 1):
        FEM.DiscreteShell.Update_Normal_Flow_Neumann(X, Elem, massMatrix, mag, bodyForce)
        total += FEM.DiscreteShell.Advance_One_Step_IE_Flow(
            Elem, segs, DBC, edge2tri, edgeStencil, edgeInfo, 0, 0, Vector4d(0, 0, 0, 0), Vector3d(0, 0, 0), Vector2d(1.01, 0), Vector2d(1, 1),
            Vector2d(0, 0), bodyForce, dt, 1e-3, True, dHat2, kappa, 0.0, 1e-6, fric_iter, compNodeRange, muComp, False, X, nodeAttr, massMatrix, elemAttr,
            elasticity, tet, tetAttr, tetElasticity, rod, rodInfo, rodHinge, rodHingeInfo, stitchInfo, stitchRatio, 10, particle, out)
        print("Total PN iteration count: ", total, "\n")
        TIMER_FLUSH(f, frames, dt, dt)
        MeshIO.Write_TriMesh_Obj(X, Elem, out + "shell%d.obj" % f)
        if Get_Parameter("Terminate", False):
            break
    return total


if __name__ == "__main__":
    a = sys.argv
    run(a[1], a[2], float(a[3]), float(a[4]), int(a[5]), a[6], float(a[7]), float(a[8]), int(a[9]))
