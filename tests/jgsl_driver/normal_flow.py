"""Stand-alone caller of the `JGSL` module for the normal-flow example (BASELINE configs[0]). It issues the module calls of
the reference's driver stack for this example -- Projects/FEMShell/12-14_normal_flow.py:7-44 ->
Python/Drivers/FEMDiscreteShellBase.py:143-147,190-204,238-242,266-273,333-341,379-380 -> SimulationBase.py:73-104 -- in the
same order with the same arguments, so the GPU box (where /root/reference does not exist) can drive the module without the
reference's scripts. Where the mirror of the unchanged scripts is present the tests run those as well.

usage: python normal_flow.py <mesh.obj> <smoothIntensity> <normalFlowMag> <frames> <output folder> [mu] [fricIterAmt]   (the drivers' sim.mu / sim.fricIterAmt, defaults 0 / 1)
"""
import os
import sys

from JGSL import *  # noqa: F401,F403  (whichever build of the module is first on the import path)


def run(mesh_path, smooth, mag, frames, out, mu=0.0, fric_iter=1):
    os.makedirs(out, exist_ok=True)
    if not out.endswith("/"):
        out += "/"
    Kokkos_Initialize()
    Set_Parameter("Basic.log_folder", out)
    if os.environ.get("IDP_PCG_REL_TOL"):  # experiments on the stopping rule of the device solve (the default is used otherwise)
        Set_Parameter("B200.pcg_rel_tol", float(os.environ["IDP_PCG_REL_TOL"]))
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

    FEM.DiscreteShell.Add_Shell(mesh_path, zero, Vector3d(1, 1, 1), zero, zero, 0, X, Elem, compNodeRange)
    dt, flow = 1.0, smooth > 0
    if flow:
        dt *= smooth
        mag /= smooth * smooth
    MeshIO.Append_Attribute(X, X0)
    dHat2 = FEM.DiscreteShell.Initialize_Shell_Hinge_EIPC(1, 0, 0, 1, dt, 1e-6, X, Elem, segs, edge2tri, edgeStencil, edgeInfo, nodeAttr, massMatrix,
                                                          gravity, bodyForce, elemAttr, elasticity, kappa)
    if flow:
        FEM.Boundary_Dirichlet(X, Elem, DBC)
    dHat2 = FEM.DiscreteShell.Initialize_OIPC(0.0, 0.0, 1e-3, 0.0, massMatrix, kappa, 1)
    MeshIO.Write_TriMesh_Obj(X, Elem, out + "shell0.obj")
    total = 0
    for f in range(1, frames + 1):
        FEM.DiscreteShell.Update_Normal_Flow_Neumann(X, Elem, massMatrix, mag, bodyForce)
        total += FEM.DiscreteShell.Advance_One_Step_IE_Flow(
            Elem, segs, DBC, edge2tri, edgeStencil, edgeInfo, 0, 0, Vector4d(0, 0, 0, 0), Vector3d(0, 0, 0), Vector2d(1.01, 0), Vector2d(1, 1),
            Vector2d(0, 0), bodyForce, dt, 1e-3, True, dHat2, kappa, mu, 1e-6, fric_iter, compNodeRange, muComp, False, X, nodeAttr, massMatrix, elemAttr,
            elasticity, tet, tetAttr, tetElasticity, rod, rodInfo, rodHinge, rodHingeInfo, stitchInfo, stitchRatio, 10, particle, out)
        print("Total PN iteration count: ", total, "\n")
        TIMER_FLUSH(f, frames, dt, dt)
        MeshIO.Write_TriMesh_Obj(X, Elem, out + "shell%d.obj" % f)
        if Get_Parameter("Terminate", False):
            break
    return total


if __name__ == "__main__":
    run(sys.argv[1], float(sys.argv[2]), float(sys.argv[3]), int(sys.argv[4]), sys.argv[5], float(sys.argv[6]) if len(sys.argv) > 6 else 0.0,
        int(sys.argv[7]) if len(sys.argv) > 7 else 1)
