"""A hinge-shell scene in the style of the reference's example scripts, driven by the reference's UNCHANGED Python/Drivers
(FEMDiscreteShellBase) from the mirror: a square cloth falls under gravity onto a ball that is a moving Dirichlet body (every ball
vertex scripted upwards, FEM.Init_Dirichlet / Step_Dirichlet), membrane + hinge bending + inertia + barrier + friction through
Advance_One_Step_IE_Hinge. Exercises what the two paper scripts do not: gravity as body force, a Dirichlet set that is a strict
subset of the vertices and moves with a velocity, lagged friction in the elastic time step.

usage (from a scratch working directory; the driver writes to ./output/cloth_on_ball/run/):
    cloth_on_ball.py <mirror Python dir> cloth.obj ball.obj frames mu [plate.obj]
With a plate (a third, fixed Dirichlet shell above the cloth) the rising ball squeezes the cloth against it: distances fall below
1e-9, the barrier stiffness doubles (IMPLICIT_EULER.h:568-598) and, once the ball's targets cannot be reached without intersection,
the augmented-Lagrangian Dirichlet path takes over -- that step never ends, in the reference either (tests compare a prefix).
"""
import os
import sys

sys.path.insert(0, sys.argv[1])
import Drivers  # noqa: E402
from JGSL import *  # noqa: E402,F401,F403

if __name__ == "__main__":
    cloth, ball, frames, mu = sys.argv[2], sys.argv[3], int(sys.argv[4]), float(sys.argv[5])
    plate = sys.argv[6] if len(sys.argv) > 6 else None
    os.makedirs("output", exist_ok=True)
    sys.argv = [sys.argv[0], "run"]  # the reference's SimulationBase names its output folder after the script and its arguments
    sim = Drivers.FEMDiscreteShellBase("double", 3)
    zero = Vector3d(0, 0, 0)
    sim.add_shell_3D(cloth, zero, zero, Vector3d(1, 0, 0), 0)
    n_cloth = sim.compNodeRange[-1]
    sim.add_shell_3D(ball, zero, zero, Vector3d(1, 0, 0), 0)
    n_all = sim.compNodeRange[-1]
    # the ball: every vertex of the second component, moving up at 0.5 per unit time
    sim.set_DBC_with_range(Vector3d(-0.1, -0.1, -0.1), Vector3d(1.1, 1.1, 1.1), Vector3d(0, 0.5, 0), zero, Vector3d(1, 0, 0), 0, Vector4i(n_cloth, 0, n_all, -1))
    if plate:
        sim.add_shell_3D(plate, zero, zero, Vector3d(1, 0, 0), 0)
        sim.set_DBC_with_range(Vector3d(-0.1, -0.1, -0.1), Vector3d(1.1, 1.1, 1.1), zero, zero, Vector3d(1, 0, 0), 0, Vector4i(n_all, 0, sim.compNodeRange[-1], -1))
    sim.dt = 0.01
    sim.frame_dt = 0.01
    sim.frame_num = frames
    sim.withCollision = True
    sim.mu = mu
    sim.fricIterAmt = 2
    sim.initialize(1000, 1e5, 0.4, 1e-3, 0)
    sim.initialize_OIPC(1e-3, 0)
    sim.run()
