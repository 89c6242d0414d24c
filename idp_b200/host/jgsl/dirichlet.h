// Dirichlet scripting the drivers call between time steps (Library/FEM/BOUNDARY_CONDITION.h:22-266): pure host bookkeeping on
// the (vertex, target) rows the time step consumes.
#pragma once
#include <climits>
#include <cstdio>
#include <cstdlib>
#include <iostream>

#include "shell_flow.h"

namespace jgsl {

inline Vec<double, 3> rotate_axis_angle(const Vec<double, 3>& axis, double ang, const Vec<double, 3>& q)
{
    Vec<double, 3> k = axis;
    const double kn = k.length();
    if (!(kn > 0) || ang == 0) return q;
    k /= kn;
    return q * std::cos(ang) + cross(k, q) * std::sin(ang) + k * (k.dot(q) * (1 - std::cos(ang)));
}

// Init_Dirichlet (:22-107): nodes of [vIndRange[0], vIndRange[2]) inside the relative box of their bounding box
inline void init_dirichlet(NodeStorage& X, const Vec<double, 3>& relBoxMin, const Vec<double, 3>& relBoxMax, const Vec<double, 3>& v,
    const Vec<double, 3>& rotCenter, const Vec<double, 3>& rotAxis, double angVelDeg, DbcStorage& DBC, DbcMotionStorage& DBCMotion,
    const Vec<int, 4>& vIndRange, double ring)
{
    if (!X.size()) {
        puts("no nodes in the model!");
        exit(-1);
    }
    Vec<double, 3> lo, hi;
    for (int id = 0; id < X.size(); ++id)
        if (id >= vIndRange[0] && id < vIndRange[2]) {
            const Vec<double, 3>& x = std::get<0>(X.rows[id]);
            if (id == vIndRange[0]) lo = hi = x;
            else for (int d = 0; d < 3; ++d) { if (hi[d] < x[d]) hi[d] = x[d]; if (lo[d] > x[d]) lo[d] = x[d]; }
        }
    Vec<double, 3> rmin = relBoxMin, rmax = relBoxMax;
    for (int d = 0; d < 3; ++d) {
        rmin[d] = rmin[d] * (hi[d] - lo[d]) + lo[d];
        rmax[d] = rmax[d] * (hi[d] - lo[d]) + lo[d];
    }
    std::cout << "DBC node inds: ";
    Vec<int, 2> range;
    range[0] = DBC.size();
    for (int id = 0; id < X.size(); ++id)
        if (id >= vIndRange[0] && id < vIndRange[2]) {
            const Vec<double, 3>& x = std::get<0>(X.rows[id]);
            if (x[0] >= rmin[0] && x[0] <= rmax[0] && x[1] >= rmin[1] && x[1] <= rmax[1] && x[2] >= rmin[2] && x[2] <= rmax[2]) {
                if (x.length() >= ring) DBC.append(Vec<double, 4>((double)id, x[0], x[1], x[2]));
                std::cout << " " << id;
            }
        }
    range[1] = DBC.size();
    DBCMotion.append(range, v, rotCenter, rotAxis, angVelDeg);
    printf("\nvelocity %le %le %le, rotCenter %le %le %le, rotAxis %le %le %le, angVelDeg %le\n", v[0], v[1], v[2], rotCenter[0], rotCenter[1],
        rotCenter[2], rotAxis[0], rotAxis[1], rotAxis[2], angVelDeg);
}

// Step_Dirichlet (:109-155): advance the targets by one step of their rigid motion
inline void step_dirichlet(DbcMotionStorage& DBCMotion, double h, DbcStorage& DBC)
{
    for (auto& m : DBCMotion.rows) {
        const Vec<int, 2>& range = std::get<0>(m);
        const Vec<double, 3>&v = std::get<1>(m), &c = std::get<2>(m), &axis = std::get<3>(m);
        const double angVelDeg = std::get<4>(m);
        for (int i = range[0]; i < range[1]; ++i) {
            Vec<double, 4>& dI = std::get<0>(DBC.rows[i]);
            if (angVelDeg) {
                const Vec<double, 3> r = rotate_axis_angle(axis, angVelDeg / 180 * M_PI * h, Vec<double, 3>(dI[1] - c[0], dI[2] - c[1], dI[3] - c[2]));
                for (int d = 0; d < 3; ++d) dI[d + 1] = r[d] + c[d];
            }
            for (int d = 0; d < 3; ++d) dI[d + 1] += v[d] * h;
        }
    }
}

inline void turn_dirichlet(DbcMotionStorage& DBCMotion) // :215-223
{
    for (auto& m : DBCMotion.rows) std::get<1>(m) = std::get<1>(m) * -1.0;
}

inline void reset_dirichlet(NodeStorage& X, DbcStorage& DBC) // :157-171
{
    for (auto& r : DBC.rows) {
        Vec<double, 4>& dI = std::get<0>(r);
        const Vec<double, 3>& x = std::get<0>(X.rows[(int)dI[0]]);
        for (int d = 0; d < 3; ++d) dI[d + 1] = x[d];
    }
}

inline void load_dirichlet(const std::string& path, int vIndOffset, const Vec<double, 3>& translate, DbcStorage& DBC) // :173-197
{
    NodeStorage X;
    TriStorage tris;
    if (read_trimesh_obj(path, X, tris)[0] < 0) return;
    for (auto& r : DBC.rows) {
        Vec<double, 4>& dI = std::get<0>(r);
        const Vec<double, 3>& x = std::get<0>(X.rows[(int)dI[0] - vIndOffset]);
        for (int d = 0; d < 3; ++d) dI[d + 1] = x[d] + translate[d];
    }
}

} // namespace jgsl
