"""Stand-alone caller of the `JGSL` module for the animation-fix example (BASELINE configs[1]). It issues the module calls of
the reference's driver stack for this example -- Projects/FEMShell/16_fix_char_seq.py:6-42 ->
Python/Drivers/FEMDiscreteShellBase.py:146-147,190-204,238-242,250-251,266-311,353-361,379-380 -> SimulationBase.py:73-104 -- in
the same order with the same arguments (membrane + hinge bending + inertia + barrier, rest shape and Dirichlet targets
re-loaded from the next frame of the sequence every step), for the GPU box where /root/reference does not exist.

usage: python fix_char_seq.py <rest.obj> <sequence folder with 1.obj, 2.obj, ...> <frames> <output folder> [dhat] [kappaMult]
"""
import math
import os
import sys

from JGSL import *  # noqa: F401,F403


def run(rest_obj, seq, frames, out, dhat=1e-2, kappa_mult=1.0):
    os.makedirs(out, exist_ok=True)
    if not out.endswith("/"):
        out += "/"
    Kokkos_Initialize()
    X, X0, Elem = Storage.V3dStorage(), Storage.V3dStorage(), Storage.V3iStorage()
    nodeAttr, massMatrix = Storage.V3dV3dV3dSdStorage(), CSR_MATRIX_D()
    elemAttr, elasticity = Storage.M2dM2dSdStorage(), FIXED_COROTATED_2.Create()
    DBC, DBCMotion = Storage.V4dStorage(), Storage.V2iV3dV3dV3dSdStorage()
    segs, edge2tri, edgeStencil, edgeInfo = StdVectorVector2i(), StdMapPairiToi(), StdVectorVector4i(), StdVectorVector3d()
    bodyForce, compNodeRange, muComp = StdVectorXd(), StdVectorXi(), StdVectorXd()
    tet, tetAttr, tetElasticity = Storage.V4iStorage(), Storage.M3dM3dSdStorage(), FIXED_COROTATED_3.Create()
    rod, rodInfo, rodHinge, rodHingeInfo = StdVectorVector2i(), StdVectorVector3d(), StdVectorVector3i(), StdVectorVector3d()
    stitchInfo, stitchRatio, particle = StdVectorVector3i(), StdVectorXd(), StdVectorXi()
    kappa, gravity, zero = Vector3d(1e5, 0, 0), Vector3d(0, 0, 0), Vector3d(0, 0, 0)
    density, young, nu, shell_thickness, dt = 1000, 100, 0.4, 0.01, 0.04

    FEM.DiscreteShell.Add_Shell(rest_obj, Vector3d(0, 0.75, 0), Vector3d(1, 1, 1), zero, Vector3d(1, 0, 0), -90, X, Elem, compNodeRange)
    FEM.Init_Dirichlet(X, Vector3d(-0.1, 0.867, -0.1), Vector3d(1.1, 1.1, 1.1), zero, zero, Vector3d(1, 0, 0), 0, DBC, DBCMotion,
                       Vector4i(0, 0, 1000000000, -1))
    MeshIO.Append_Attribute(X, X0)

    def init_shell(rest):
        return FEM.DiscreteShell.Initialize_Shell_Hinge_EIPC(density, young, nu, shell_thickness, dt, 1e-6, rest, Elem, segs, edge2tri, edgeStencil, edgeInfo,
                                                             nodeAttr, massMatrix, gravity, bodyForce, elemAttr, elasticity, kappa)
    init_shell(X)
    dHat2 = FEM.DiscreteShell.Initialize_OIPC(0.0, 0.0, dhat, 0.0, massMatrix, kappa, kappa_mult)
    offset = 0
    MeshIO.Write_TriMesh_Obj(X, Elem, out + "shell0.obj")
    total, lv_fn = 0, 1
    for f in range(1, frames + 1):
        FEM.Step_Dirichlet(DBCMotion, dt, DBC)
        target = "%s/%d.obj" % (seq, lv_fn)
        newX, newElem = Storage.V3dStorage(), Storage.V3iStorage()
        MeshIO.Read_TriMesh_Obj(target, newX, newElem)        # the next frame becomes the rest shape ...
        init_shell(newX)
        dHat2 = FEM.DiscreteShell.Initialize_OIPC(0.0, 0.0, math.sqrt(dHat2), 0.0, massMatrix, kappa, 1)
        MeshIO.Load_Velocity_X0(seq, lv_fn, dt, X, nodeAttr)  # ... the velocity points at it ...
        FEM.Load_Dirichlet(target, 0, zero, DBC)              # ... and the pinned nodes follow it
        lv_fn += 1
        total += FEM.DiscreteShell.Advance_One_Step_IE_Hinge(
            Elem, segs, DBC, edge2tri, edgeStencil, edgeInfo, offset, 1, Vector4d(0, 0, 0, 0), Vector3d(0, 0, 0), Vector2d(1.01, 0), Vector2d(1, 1),
            Vector2d(0, 0), bodyForce, dt, 1e-3, True, dHat2, kappa, 0, 1e-6, 1, compNodeRange, muComp, False, X, nodeAttr, massMatrix, elemAttr,
            elasticity, tet, tetAttr, tetElasticity, rod, rodInfo, rodHinge, rodHingeInfo, stitchInfo, stitchRatio, 10, particle, out)
        print("Total PN iteration count: ", total, "\n")
        TIMER_FLUSH(f, frames, dt, dt)
        MeshIO.Write_TriMesh_Obj(X, Elem, out + "shell%d.obj" % f)
    return total


if __name__ == "__main__":
    run(sys.argv[1], sys.argv[2], int(sys.argv[3]), sys.argv[4], *(float(a) for a in sys.argv[5:7]))
