// Host Newton driver of the discrete-shell time step around the device-resident contact path.
//
// What the reference does in Library/FEM/Shell/IMPLICIT_EULER.h:151-891 (Advance_One_Step_IE_Discrete_Shell<double,3,KL=false,
// elasticIPC=false,flow>, exported as FEM.DiscreteShell.Advance_One_Step_IE_Flow / _IE_Hinge, DISCRETE_SHELL.h:1110,1113) with
// Line_Search (:9-149) and Compute_IncPotential / _Gradient / _Hessian (INC_POTENTIAL.h:14-394: the `flow` branch, or membrane +
// hinge bending + inertia), restated here on top of a small backend interface: every contact operator, the system-matrix assembly, Project_DBC and
// the linear solve are ONE backend call each, so that the B200 backend (backend_b200.h) keeps them on the device and the
// host only runs the O(nV) vector algebra and the control flow. Same printed lines, same files (residual.txt, counter.txt,
// stretch.txt, Hessian_info.txt), same error behaviour (message + exit(-1)).
#pragma once
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <fstream>
#include <iostream>
#include <set>
#include <sstream>
#include <string>
#include <vector>

#include "jgsl_types.h"

namespace jgsl {

// ---------------------------------------------------------------------------------------------------------------------------
// The operators the time step needs from whoever owns the contact path. Positions / directions / gradients are dense
// host arrays of 3 doubles per vertex.
struct ContactBackend {
    virtual ~ContactBackend() {}
    virtual const char* name() const = 0;
    // Find_Surface_Primitives_And_Compute_Area (Utils/MESHIO.h:768-834) + the Dirichlet mask DBCb (IMPLICIT_EULER.h:291-316)
    virtual void set_mesh(int nV, const std::vector<int>& tri3, const double* x, const std::vector<uint8_t>& dbc) = 0;
    virtual void set_rest_positions(const double* x0) = 0;
    // the constant terms of the system matrix: Laplacian flow blocks (INC_POTENTIAL.h:323-339), lumped mass (:383-386)
    virtual void set_system_terms(const std::vector<int>& elem3, const std::vector<double>& vol, double h, const std::vector<double>& mass) = 0;
    // elastic terms of the non-flow step: membrane triangles (MEMBRANE.h) and bending hinges (BENDING.h, KL = false); empty = none
    virtual void set_elastic_terms(const std::vector<int>& elem3, const std::vector<double>& ib3, const std::vector<double>& vol,
        const std::vector<double>& lambda, const std::vector<double>& mu, const std::vector<int>& stencil4, const std::vector<double>& info3, double k,
        double h) = 0;
    virtual void elastic_energy(double& E) = 0;   // adds Compute_Membrane_Energy + Compute_Bending_Energy
    virtual void elastic_gradient(double* g) = 0; // adds
    virtual void set_positions(const double* x) = 0;
    virtual int constraint_set(double dHat2, double thickness) = 0;                      // Compute_Constraint_Set, returns #rows
    virtual void barrier_energy(double dHat2, double kappa, double thickness, double& E) = 0;          // adds
    virtual void barrier_gradient(double dHat2, double kappa, double thickness, double* g) = 0;        // adds
    // Compute_IncPotential_Hessian (flow) + Solve_Direct: assemble flow + mass + projected barrier Hessians, Project_DBC, solve
    // projMask: the vertices Project_DBC fixes (null: the Dirichlet mask of set_mesh)
    virtual bool solve_newton_system(double dHat2, double kappa, double thickness, const std::vector<uint8_t>* projMask, const double* rhs, double* sol) = 0;
    // lagged friction (FEM/FRICTION.h): freeze the current rows at the current positions; E / g at the current positions relative
    // to xn; the Hessian blocks are part of solve_newton_system while mu > 0
    virtual long friction_update(double dHat2, double kappa, double thickness) = 0;     // Compute_Friction_Basis, returns #friction rows
    virtual void friction_set(const double* xn, double epsv2h2, double mu) = 0;
    // Compute_Friction_Coef: per-component coefficients applied to the normal forces at every friction_update (empty: off)
    virtual void friction_set_components(const std::vector<int>& compNodeRange, const std::vector<double>& muComp) = 0;
    virtual void friction_energy(double& E) = 0;   // adds
    virtual void friction_gradient(double* g) = 0; // adds
    virtual double ccd(const double* dir, double thickness, double alpha) = 0;          // Compute_Intersection_Free_StepSize
    virtual bool min_dist2(double thickness, std::vector<double>* dist2, double& minDist2) = 0; // false: no rows
    virtual void get_rows(std::vector<int>& rows4, std::vector<double>& info2) = 0;
    virtual void set_rows(const std::vector<int>& rows4, const std::vector<double>& info2) = 0;
};

// ---------------------------------------------------------------------------------------------------------------------------
// Mesh files (Utils/MESHIO.h:17-80, 185-214): the same tolerant "v x y z" / "f a/b/c ..." reader (quads split), "%le" writer.
inline Vec<int, 4> read_trimesh_obj(const std::string& path, NodeStorage& X, TriStorage& tris)
{
    std::ifstream is(path);
    if (!is.is_open()) {
        puts((path + " not found!").c_str());
        return Vec<int, 4>(-1, -1, -1, -1);
    }
    Vec<int, 4> counter(X.size(), tris.size(), 0, 0);
    std::string line;
    auto next_index = [&](const std::string& s, size_t& p, int& out) {
        while (p < s.size() && (s[p] < '0' || s[p] > '9')) ++p;
        if (p >= s.size()) return false;
        int v = 0;
        while (p < s.size() && s[p] >= '0' && s[p] <= '9') v = v * 10 + (s[p++] - '0');
        out = v;
        return true;
    };
    while (std::getline(is, line)) {
        if (line.size() > 1 && line[0] == 'v' && line[1] == ' ') {
            std::stringstream ss(line.substr(1));
            Vec<double, 3> p;
            ss >> p[0] >> p[1] >> p[2];
            X.append(p);
        }
        else if (!line.empty() && line[0] == 'f') {
            size_t p = 0;
            int id[4] = {0, 0, 0, 0};
            int got = 0;
            for (; got < 4; ++got) {
                if (!next_index(line, p, id[got])) break;
                while (p < line.size() && line[p] != ' ') ++p; // skip "/vt/vn"
            }
            if (got >= 3) {
                Vec<int, 3> t(id[0] - 1 + counter[0], id[1] - 1 + counter[0], id[2] - 1 + counter[0]);
                tris.append(t);
                if (got == 4) tris.append(Vec<int, 3>(t[0], t[2], id[3] - 1 + counter[0]));
            }
        }
    }
    counter[2] = X.size();
    counter[3] = tris.size();
    return counter;
}

inline void write_trimesh_obj(const NodeStorage& X, const TriStorage& tris, const std::string& path)
{
    FILE* f = fopen(path.c_str(), "w");
    if (!f) {
        puts("failed to create file");
        exit(-1);
    }
    for (const auto& r : X.rows) fprintf(f, "v %le %le %le\n", std::get<0>(r)[0], std::get<0>(r)[1], std::get<0>(r)[2]);
    for (const auto& r : tris.rows) fprintf(f, "f %d %d %d\n", std::get<0>(r)[0] + 1, std::get<0>(r)[1] + 1, std::get<0>(r)[2] + 1);
    fclose(f);
}

// Add_Discrete_Shell_3D (DISCRETE_SHELL.h:14-64): read, scale about rotCenter, rotate (axis-angle, Rodrigues), translate, append.
inline Vec<int, 4> add_shell(const std::string& path, const Vec<double, 3>& trans, const Vec<double, 3>& scale, const Vec<double, 3>& rotCenter,
    const Vec<double, 3>& rotAxis, double rotAngDeg, NodeStorage& X, TriStorage& Elem, std::vector<int>& compNodeRange)
{
    NodeStorage newX;
    TriStorage newElem;
    Vec<int, 4> counter = read_trimesh_obj(path, newX, newElem);
    counter[0] += X.size(); counter[2] += X.size();
    counter[1] += Elem.size(); counter[3] += Elem.size();
    const double ang = rotAngDeg / 180 * M_PI;
    const double c = std::cos(ang), s = std::sin(ang);
    Vec<double, 3> k = rotAxis;
    const double kn = k.length();
    if (kn > 0) k /= kn; // Eigen::AngleAxis expects a unit axis; a zero axis with a zero angle is the identity
    for (auto& r : newX.rows) {
        Vec<double, 3>& p = std::get<0>(r);
        Vec<double, 3> q = p - rotCenter;
        for (int i = 0; i < 3; ++i) q[i] *= scale[i];
        Vec<double, 3> rot = q;
        if (kn > 0 && ang != 0) rot = q * c + cross(k, q) * s + k * (k.dot(q) * (1 - c));
        for (int i = 0; i < 3; ++i) p[i] = rot[i] + rotCenter[i] + trans[i];
    }
    const int base = X.size();
    for (auto& r : newElem.rows) for (int i = 0; i < 3; ++i) std::get<0>(r)[i] += base;
    X.rows.insert(X.rows.end(), newX.rows.begin(), newX.rows.end());
    Elem.rows.insert(Elem.rows.end(), newElem.rows.begin(), newElem.rows.end());
    compNodeRange.emplace_back(X.size());
    return counter;
}

// Compute_Dihedral_Angle (Math/DIHEDRAL_ANGLE.h:9-24)
inline double dihedral_angle(const Vec<double, 3>& v0, const Vec<double, 3>& v1, const Vec<double, 3>& v2, const Vec<double, 3>& v3)
{
    const Vec<double, 3> n1 = cross(v1 - v0, v2 - v0), n2 = cross(v2 - v3, v1 - v3);
    double a = std::acos(std::max(-1.0, std::min(1.0, n1.dot(n2) / std::sqrt(n1.length2() * n2.length2()))));
    if (cross(n2, n1).dot(v1 - v2) < 0) a = -a;
    return a;
}

// Initialize_Discrete_Shell<double,3,KL=false,elasticIPC=true> (DISCRETE_SHELL.h:218-357; exported as
// Initialize_Shell_Hinge_EIPC): drops degenerate triangles, rest state, directed-edge map, hinge stencils with rest angle /
// rest length / height, lumped mass + gravity body force, per-element volume and Lame parameters, hinge stiffness,
// elastic-IPC kappa; returns dHat2 = thickness^2.
inline double initialize_shell_hinge(double rho0, double E, double nu, double thickness, double h, double dHat2, NodeStorage& X, TriStorage& Elem,
    std::vector<Vec<int, 2>>& seg, EdgeToTri& edge2tri, std::vector<Vec<int, 4>>& edgeStencil, std::vector<Vec<double, 3>>& edgeInfo,
    NodeAttrStorage& nodeAttr, CsrMatrix& M, const Vec<double, 3>& gravity, std::vector<double>& b, ElemAttrStorage& elemAttr,
    Fcr2Storage& elasticityAttr, Vec<double, 3>& kappa, bool elasticIPC = true)
{
    for (const auto& r : Elem.rows)
        for (int k = 0; k < 3; ++k)
            if (std::get<0>(r)[k] < 0 || std::get<0>(r)[k] >= X.size()) { // e.g. the rest shape of the next frame could not be read
                printf("Initialize_Shell: element vertex %d outside the %d nodes given (mesh file missing?)\n", std::get<0>(r)[k], X.size());
                exit(-1);
            }
    auto P = [&](int v) -> const Vec<double, 3>& { return std::get<0>(X.rows[v]); };
    TriStorage kept;
    for (const auto& r : Elem.rows) {
        const Vec<int, 3>& t = std::get<0>(r);
        const Vec<double, 3> e1 = P(t[1]) - P(t[0]), e2 = P(t[2]) - P(t[0]);
        const double a = e1.length2(), bb = e1.dot(e2), c = e2.length2();
        if (a * c - bb * bb != 0) kept.rows.push_back(r);
    }
    Elem = kept;

    nodeAttr.clear();
    for (int i = 0; i < X.size(); ++i) nodeAttr.append(P(i), Vec<double, 3>(), Vec<double, 3>(), 0.0);

    edge2tri.clear();
    for (int e = 0; e < Elem.size(); ++e) {
        const Vec<int, 3>& t = std::get<0>(Elem.rows[e]);
        edge2tri[std::make_pair(t[0], t[1])] = e;
        edge2tri[std::make_pair(t[1], t[2])] = e;
        edge2tri[std::make_pair(t[2], t[0])] = e;
    }

    // Find_Surface_Primitives (MESHIO.h:728-766): one entry per undirected edge, orientation of the first triangle, set order
    std::set<std::pair<int, int>> edges;
    for (const auto& r : Elem.rows) {
        const Vec<int, 3>& t = std::get<0>(r);
        for (int i = 0; i < 3; ++i) {
            const int u = t[i], v = t[(i + 1) % 3];
            if (edges.find(std::make_pair(v, u)) == edges.end()) edges.insert(std::make_pair(u, v));
        }
    }

    // Compute_Discrete_Shell_Inv_Basis<KL=false> (DISCRETE_SHELL.h:138-216): first fundamental form + hinge stencils
    elemAttr.clear();
    for (const auto& r : Elem.rows) {
        const Vec<int, 3>& t = std::get<0>(r);
        const Vec<double, 3> e1 = P(t[1]) - P(t[0]), e2 = P(t[2]) - P(t[0]);
        Mat<double, 2> IB, D;
        IB(0, 0) = e1.length2(); IB(1, 0) = IB(0, 1) = e1.dot(e2); IB(1, 1) = e2.length2();
        elemAttr.append(IB, D);
    }
    edgeStencil.clear();
    edgeInfo.clear();
    for (const auto& e : edges) {
        const auto tf = edge2tri.find(e);
        if (tf == edge2tri.end()) continue;
        const Vec<int, 3>& t = std::get<0>(Elem.rows[tf->second]);
        int v0 = -1;
        for (int j = 0; j < 3; ++j) if (t[j] == e.second) { v0 = t[(j + 1) % 3]; break; }
        const auto of = edge2tri.find(std::make_pair(e.second, e.first));
        if (of == edge2tri.end()) continue; // boundary edge: no hinge
        const Vec<int, 3>& o = std::get<0>(Elem.rows[of->second]);
        int v3 = 0;
        for (int j = 0; j < 3; ++j) if (o[j] == e.first) { v3 = o[(j + 1) % 3]; break; }
        edgeStencil.emplace_back(v0, e.first, e.second, v3);
        Vec<double, 3> info;
        info[0] = dihedral_angle(P(v0), P(e.first), P(e.second), P(v3));
        info[1] = (P(e.first) - P(e.second)).length();
        const Vec<double, 3> n1 = cross(P(e.first) - P(v0), P(e.second) - P(v0)), n2 = cross(P(e.second) - P(v3), P(e.first) - P(v3));
        info[2] = (n1.length() + n2.length()) / (info[1] * 6);
        edgeInfo.push_back(info);
    }
    std::cout << edgeStencil.size() << " hinges" << std::endl;
    std::cout << "IB and D computed" << std::endl;

    // lumped mass and body force (:279-318)
    std::vector<double> diag(3 * (size_t)X.size(), 0.0);
    b.assign(3 * (size_t)X.size(), 0.0);
    double massPortionMean = 0;
    for (const auto& r : Elem.rows) {
        const Vec<int, 3>& t = std::get<0>(r);
        const double mp = cross(P(t[1]) - P(t[0]), P(t[2]) - P(t[0])).length() / 2 * thickness * rho0 / 3;
        massPortionMean += mp;
        for (int k = 0; k < 3; ++k) {
            std::get<3>(nodeAttr.rows[t[k]]) += mp;
            for (int d = 0; d < 3; ++d) {
                diag[3 * (size_t)t[k] + d] += mp;
                b[3 * (size_t)t[k] + d] += mp * gravity[d];
            }
        }
    }
    if (Elem.size()) massPortionMean /= Elem.size();
    for (const auto& s : seg)
        for (int k = 0; k < 2; ++k)
            for (int d = 0; d < 3; ++d) diag[3 * (size_t)s[k] + d] += massPortionMean * 3;
    M.set_diagonal(diag);

    // quadratures (:320-334) and hinge stiffness (:335-341)
    elasticityAttr.clear();
    const double lambda = E * nu / (1.0 - nu * nu), mu = E / (2.0 * (1.0 + nu));
    for (const auto& r : Elem.rows) {
        const Vec<int, 3>& t = std::get<0>(r);
        const double area = cross(P(t[1]) - P(t[0]), P(t[2]) - P(t[0])).length() / 2;
        elasticityAttr.append(Mat<double, 2>(), area * thickness, lambda, mu);
    }
    if (elemAttr.size()) {
        double& k = std::get<1>(elemAttr.rows[0])(0, 0);
        k = E * std::pow(thickness, 3) / (24 * (1.0 - nu * nu));
        std::cout << "hinge k = " << k << std::endl;
    }
    if (elasticIPC) { // :343-349; the non-EIPC instantiation leaves kappa and dHat2 alone
        kappa[0] = h * h * mu; kappa[1] = h * h * lambda; kappa[2] = nu;
        dHat2 = thickness * thickness;
    }
    std::cout << "shell initialized" << std::endl;
    return dHat2;
}

// the barrier's second derivative b''(d) of Math/BARRIER.h:45-53 (non-elastic form), needed for the kappa heuristic
inline double barrier_hessian_scalar(double d, double dHat, double kappa0)
{
    const double t2 = d - dHat;
    return kappa0 * ((std::log(d / dHat) * -2.0 - t2 * 4.0 / d) + 1.0 / (d * d) * (t2 * t2));
}

// Initialize_EIPC<double, elasticIPC=false> (DISCRETE_SHELL.h:555-577; exported as Initialize_OIPC)
inline double initialize_oipc(double E, double nu, double thickness, double h, CsrMatrix& M, Vec<double, 3>& kappa, double stiffMult)
{
    const double lambda = E * nu / (1.0 - nu * nu), mu = E / (2.0 * (1.0 + nu));
    kappa[0] = h * h * mu; kappa[1] = h * h * lambda; kappa[2] = nu;
    const double dHat2 = thickness * thickness;
    const double Hb = barrier_hessian_scalar(1.0e-16, dHat2, 1.0);
    kappa[0] = stiffMult * 1.0e11 * M.diagonal_mean() * 3 / (4.0e-16 * Hb);
    kappa[1] = 100 * kappa[0];
    printf("original IPC kappa = %le\n", kappa[0]);
    return dHat2;
}

// Initialize_EIPC<double, elasticIPC=true> (DISCRETE_SHELL.h:555-564; exported as Initialize_EIPC)
inline double initialize_eipc(double E, double nu, double thickness, double h, CsrMatrix& M, Vec<double, 3>& kappa, double stiffMult)
{
    (void)M; (void)stiffMult;
    const double lambda = E * nu / (1.0 - nu * nu), mu = E / (2.0 * (1.0 + nu));
    kappa[0] = h * h * mu; kappa[1] = h * h * lambda; kappa[2] = nu;
    return thickness * thickness;
}

// Initialize_OIPC_VecM (DISCRETE_SHELL.h:579-600; exported as Initialize_OIPC_VM): the kappa heuristic from the mean nodal mass
inline double initialize_oipc_vm(double dHat2, NodeAttrStorage& nodeAttr, Vec<double, 3>& kappa, double stiffMult)
{
    double avg = 0;
    for (const auto& r : nodeAttr.rows) avg += std::get<3>(r);
    avg /= nodeAttr.size();
    kappa[0] = stiffMult * 1.0e11 * avg / (4.0e-16 * barrier_hessian_scalar(1.0e-16, dHat2, 1.0));
    kappa[1] = 100 * kappa[0];
    printf("original IPC kappa = %le\n", kappa[0]);
    return dHat2;
}

// Update_Normal_Flow_Neumann (DISCRETE_SHELL.h:359-395): b_i = magnitude * M_ii * (area-weighted vertex normal, normalised)
inline void update_normal_flow_neumann(const NodeStorage& X, const TriStorage& Elem, const CsrMatrix& M, double magnitude, std::vector<double>& b)
{
    b.assign(3 * (size_t)X.size(), 0.0);
    for (const auto& r : Elem.rows) {
        const Vec<int, 3>& t = std::get<0>(r);
        const Vec<double, 3>&x1 = std::get<0>(X.rows[t[0]]), &x2 = std::get<0>(X.rows[t[1]]), &x3 = std::get<0>(X.rows[t[2]]);
        const Vec<double, 3> n = cross(x2 - x1, x3 - x1);
        for (int k = 0; k < 3; ++k) for (int d = 0; d < 3; ++d) b[3 * (size_t)t[k] + d] += n[d];
    }
    for (int v = 0; v < X.size(); ++v) {
        double sq = 0;
        for (int d = 0; d < 3; ++d) sq += b[3 * (size_t)v + d] * b[3 * (size_t)v + d];
        const double w = magnitude * M.coeff(3 * v, 3 * v) / std::sqrt(sq);
        for (int d = 0; d < 3; ++d) b[3 * (size_t)v + d] *= w;
    }
}

// Boundary_Dirichlet (FEM/BOUNDARY_CONDITION.h:199-213): pin the nodes of open-boundary edges where they are
inline void boundary_dirichlet(const NodeStorage& X, const TriStorage& Tri, DbcStorage& DBC)
{
    std::set<std::pair<int, int>> es;
    for (const auto& r : Tri.rows) {
        const Vec<int, 3>& t = std::get<0>(r);
        for (int i = 0; i < 3; ++i) es.insert(std::make_pair(t[i], t[(i + 1) % 3]));
    }
    std::vector<bool> isB((size_t)X.size(), false);
    for (const auto& e : es)
        if (es.find(std::make_pair(e.second, e.first)) == es.end()) isB[e.first] = isB[e.second] = true;
    for (int v = 0; v < X.size(); ++v)
        if (isB[v]) {
            const Vec<double, 3>& x = std::get<0>(X.rows[v]);
            DBC.append(Vec<double, 4>((double)v, x[0], x[1], x[2]));
        }
}

// Compute_Max_And_Avg_Stretch (DISCRETE_SHELL.h:601-671): extreme principal stretches of F = A B^-1 per triangle, where A / B
// are the upper-triangular factors of the current / rest first fundamental forms; the singular values of a 2x2 matrix in
// closed form instead of the reference's SVD routine.
inline void max_and_avg_stretch(const TriStorage& Elem, const std::vector<bool>& DBCb, const double* x, const ElemAttrStorage& elemAttr, double& maxs,
    double& avgs, double& minc, double& avgc)
{
    maxs = 1.0; avgs = 0.0; minc = 1.0; avgc = 0.0;
    int ns = 0, nc = 0;
    for (int e = 0; e < Elem.size(); ++e) {
        const Vec<int, 3>& t = std::get<0>(Elem.rows[e]);
        if (DBCb[t[0]] && DBCb[t[1]] && DBCb[t[2]]) continue;
        const Mat<double, 2>& G = std::get<0>(elemAttr.rows[e]);
        const double b00 = std::sqrt(G(0, 0)), b01 = G(0, 1) / b00, b11 = std::sqrt(G(1, 1) - G(0, 1) * G(0, 1) / G(0, 0));
        if (b00 * b11 == 0.0) continue;
        Vec<double, 3> x1(x[3 * t[0]], x[3 * t[0] + 1], x[3 * t[0] + 2]), x2(x[3 * t[1]], x[3 * t[1] + 1], x[3 * t[1] + 2]),
            x3(x[3 * t[2]], x[3 * t[2] + 1], x[3 * t[2] + 2]);
        const double a00 = (x2 - x1).length(), a01 = (x2 - x1).dot(x3 - x1) / a00, a11 = cross(x2 - x1, x3 - x1).length() / a00;
        // F = A B^-1 (both upper triangular)
        const double f00 = a00 / b00, f01 = (a01 - f00 * b01) / b11, f11 = a11 / b11;
        const double p = f00 * f00 + f01 * f01 + f11 * f11, q = f00 * f11; // trace(F^T F), det F
        const double disc = std::sqrt(std::max(0.0, p * p - 4 * q * q));
        const double smax = std::sqrt((p + disc) / 2), smin = std::sqrt(std::max(0.0, (p - disc) / 2));
        if (smax > maxs) maxs = smax;
        if (smax > 1) { ++ns; avgs += smax; }
        if (smin < minc) minc = smin;
        if (smin < 1) { ++nc; avgc += smin; }
    }
    if (ns) avgs /= ns;
    if (nc) avgc /= nc;
}

// ---------------------------------------------------------------------------------------------------------------------------
// The incremental potential without the barrier term (INC_POTENTIAL.h:14-150 energy, :152-276 gradient):
//   flow:   -h/2 x^T L x, L assembled per triangle with weight vol/6 (host, O(nF))             (:53-73, 191-211)
//   else:   membrane + hinge bending (backend: on the device for the B200 build)                (:75-96, 213-232)
//   both:   + 1/2 (x - xtilde)^T M (x - xtilde), M lumped                                       (:131-144, 250-268)
//   + the augmented-Lagrangian Dirichlet penalty 1/2 k sum m_v |x_v - target_v|^2 while it is active (DIRICHLET.h:25-71)
struct StepPotential {
    bool flow = false;
    ContactBackend* be = nullptr;
    const TriStorage* Elem = nullptr;
    const Fcr2Storage* fcr = nullptr;
    const std::vector<double>* massDiag = nullptr; // 3 nV
    const DbcStorage* DBC = nullptr;
    const NodeAttrStorage* nodeAttr = nullptr;
    double h = 0, DBCStiff = 0;
    bool friction = false; // mu > 0 and contact on: Compute_Friction_Potential joins every energy evaluation
    std::vector<double> Lx;

    void laplacian(const std::vector<double>& x)
    {
        Lx.assign(x.size(), 0.0);
        for (int e = 0; e < Elem->size(); ++e) {
            const Vec<int, 3>& t = std::get<0>(Elem->rows[e]);
            const double w = std::get<1>(fcr->rows[e]) / 6;
            for (int i = 0; i < 3; ++i) {
                const size_t a = 3 * (size_t)t[i], p = 3 * (size_t)t[(i + 1) % 3], q = 3 * (size_t)t[(i + 2) % 3];
                for (int d = 0; d < 3; ++d) Lx[a + d] += w * (x[p + d] + x[q + d] - 2 * x[a + d]);
            }
        }
    }
    double dbc_dist2(const std::vector<double>& x, bool massWeighted) const
    {
        double s = 0;
        for (const auto& r : DBC->rows) {
            const Vec<double, 4>& dI = std::get<0>(r);
            const int v = (int)dI[0];
            double p = 0;
            for (int d = 0; d < 3; ++d) p += (dI[d + 1] - x[3 * (size_t)v + d]) * (dI[d + 1] - x[3 * (size_t)v + d]);
            s += massWeighted ? p * std::get<3>(nodeAttr->rows[v]) : p;
        }
        return s;
    }
    // the backend must hold x as its current positions when flow == false
    double energy(const std::vector<double>& x, const std::vector<double>& xtilde)
    {
        double E = 0;
        if (flow) {
            laplacian(x);
            for (size_t i = 0; i < x.size(); ++i) E += -h * 0.5 * x[i] * Lx[i];
        }
        else be->elastic_energy(E);
        double I = 0;
        for (size_t i = 0; i < x.size(); ++i) {
            const double dx = x[i] - xtilde[i];
            I += (*massDiag)[i] * dx * dx;
        }
        return E + 0.5 * I;
    }
    void gradient(const std::vector<double>& x, const std::vector<double>& xtilde, std::vector<double>& g)
    {
        g.assign(x.size(), 0.0);
        if (flow) {
            laplacian(x);
            for (size_t i = 0; i < x.size(); ++i) g[i] = -h * Lx[i];
        }
        else be->elastic_gradient(g.data());
        for (size_t i = 0; i < x.size(); ++i) g[i] += (*massDiag)[i] * (x[i] - xtilde[i]);
    }
    void add_dbc_energy(const std::vector<double>& x, double& E) const
    {
        if (DBCStiff) E += 0.5 * DBCStiff * dbc_dist2(x, true);
    }
};

struct StepState { // what Line_Search shares with the Newton loop
    std::vector<double> x, xtilde, sol, rhs, g;
    int nRows = 0;
    double Eprev = 0;
};

// Line_Search (IMPLICIT_EULER.h:9-149) for the configurations hosted here (no inextensibility, no fibers, no tets, mu = 0)
inline void step_line_search(ContactBackend& be, StepPotential& pot, StepState& s, bool withCollision, double dHat2, double kappa, double thickness,
    double& alpha, double& feasibleAlpha)
{
    const std::vector<double> xprev = s.x;
    alpha = 1.0;
    if (withCollision) {
        alpha = be.ccd(s.sol.data(), thickness, alpha);
        printf("intersection free step size = %le\n", alpha);
    }
    feasibleAlpha = alpha;
    double E;
    do {
        for (size_t i = 0; i < s.x.size(); ++i) s.x[i] = xprev[i] + alpha * s.sol[i];
        be.set_positions(s.x.data());
        E = pot.energy(s.x, s.xtilde);
        if (withCollision) {
            s.nRows = be.constraint_set(dHat2, thickness);
            if (s.nRows) {
                double minDist2 = 0;
                be.min_dist2(thickness, nullptr, minDist2);
                if (minDist2 <= 0) {
                    std::cout << "safe guard backtrack!" << std::endl;
                    alpha /= 2;
                    E = s.Eprev + 1;
                    continue;
                }
            }
            be.barrier_energy(dHat2, kappa, thickness, E);
            if (pot.friction) be.friction_energy(E);
        }
        pot.add_dbc_energy(s.x, E);
        alpha /= 2.0;
        printf("E %le, Eprev %le, alpha %le, valid %d\n", E, s.Eprev, alpha * 2, 1);
    } while (E > s.Eprev);
    printf("alpha = %le\n", alpha * 2.0);
    s.Eprev = E;
}

struct ShellStepInputs { // the arguments of Advance_One_Step_IE_Discrete_Shell the hosted variants read (IMPLICIT_EULER.h:151-187)
    bool flow = false;
    double thickness = 0, bendingStiffMult = 0, h = 0, NewtonTol = 1e-3, dHat2 = 0, mu = 0, epsv2 = 0;
    int fricIterAmt = 1;
    std::vector<int> compNodeRange; // Compute_Friction_Coef: used when muComp holds one coefficient per pair of components
    std::vector<double> muComp;
    bool withCollision = false, staticSolve = false;
    int nTet = 0, nRod = 0, nStitch = 0, nParticle = 0;
    std::string outputFolder;
};

// Advance_One_Step_IE_Discrete_Shell<double, 3, KL=false, elasticIPC=false, flow> (IMPLICIT_EULER.h:151-891).
// Unsupported inputs are rejected like the contact path rejects them (message + exit(-1)): segments, rods, particles, tets,
// stitches, strain limiting, fibers, static solves.
inline int advance_one_step_ie(ContactBackend& be, const ShellStepInputs& in, TriStorage& Elem, const std::vector<Vec<int, 2>>& seg, DbcStorage& DBC,
    const std::vector<Vec<int, 4>>& edgeStencil, const std::vector<Vec<double, 3>>& edgeInfo, const Vec<double, 4>& fiberStiffMult,
    const Vec<double, 2>& kappa_s, const std::vector<double>& b, Vec<double, 3>& kappaVec, NodeStorage& X, NodeAttrStorage& nodeAttr, CsrMatrix& M,
    ElemAttrStorage& elemAttr, Fcr2Storage& elasticityAttr)
{
    const bool flow = in.flow, withCollision = in.withCollision;
    const double h = in.h, thickness = in.thickness, dHat2 = in.dHat2, NewtonTol = in.NewtonTol;
    const std::string& outputFolder = in.outputFolder;
    if (!seg.empty() || in.nTet || in.nRod || in.nStitch || in.nParticle || kappa_s[0] > 0 || fiberStiffMult[0] > 0 || fiberStiffMult[1] > 0 ||
        in.staticSolve) {
        printf("Advance_One_Step_IE (%s): segments / rods / particles / tets / stitches / strain limiting / fibers / static "
               "solves are outside the device-resident shell step\n", be.name());
        exit(-1);
    }
    const int nV = X.size();
    const size_t n3 = 3 * (size_t)nV;
    double kappa[3] = {kappaVec[0], kappaVec[1], kappaVec[2]};

    StepState s;
    s.x.resize(n3);
    for (int v = 0; v < nV; ++v) for (int d = 0; d < 3; ++d) s.x[3 * (size_t)v + d] = std::get<0>(X.rows[v])[d];
    const std::vector<double> xn = s.x;

    // Xtilde = Xn + h v + h^2 M^-1 b, v zeroed by the flow variant (:196-217); M is the lumped (diagonal) mass
    std::vector<double> massDiag(n3);
    for (size_t i = 0; i < n3; ++i) {
        massDiag[i] = M.coeff((int)i, (int)i);
        if (!(massDiag[i] > 0)) {
            std::cout << "mass matrix factorization failed!" << std::endl;
            exit(-1);
        }
    }
    s.xtilde = xn;
    for (int v = 0; v < nV; ++v) {
        if (flow) std::get<1>(nodeAttr.rows[v]) = Vec<double, 3>();
        const Vec<double, 3>& vel = std::get<1>(nodeAttr.rows[v]);
        for (int d = 0; d < 3; ++d) s.xtilde[3 * (size_t)v + d] += h * vel[d] + h * h * (b[3 * (size_t)v + d] / massDiag[3 * (size_t)v + d]);
    }
    std::cout << "Xn and Xtilde prepared" << std::endl;

    // Dirichlet data (:291-316): mask, displacement to the targets, nodes whose target is where they are
    std::vector<uint8_t> dbcMask((size_t)nV, 0), dbcFixed((size_t)nV, 0);
    std::vector<bool> DBCb((size_t)nV, false);
    std::vector<double> DBCDisp(n3, 0.0);
    for (const auto& r : DBC.rows) {
        const Vec<double, 4>& dI = std::get<0>(r);
        const int v = (int)dI[0];
        bool still = true;
        for (int d = 0; d < 3; ++d) {
            DBCDisp[3 * (size_t)v + d] = dI[d + 1] - s.x[3 * (size_t)v + d];
            still = still && !DBCDisp[3 * (size_t)v + d];
        }
        dbcFixed[v] = still ? 1 : 0;
        dbcMask[v] = 1;
        DBCb[v] = true;
    }

    // surface primitives: once per step in the reference (:222-241, "TODO: only once"), once per mesh on the device here
    std::vector<int> tri3(3 * (size_t)Elem.size());
    std::vector<double> vol((size_t)Elem.size());
    for (int e = 0; e < Elem.size(); ++e) {
        for (int k = 0; k < 3; ++k) tri3[3 * (size_t)e + k] = std::get<0>(Elem.rows[e])[k];
        vol[e] = std::get<1>(elasticityAttr.rows[e]);
    }
    std::vector<double> x0(n3);
    for (int v = 0; v < nV; ++v) for (int d = 0; d < 3; ++d) x0[3 * (size_t)v + d] = std::get<0>(nodeAttr.rows[v])[d];
    be.set_mesh(nV, tri3, s.x.data(), dbcMask);
    be.set_rest_positions(x0.data());
    std::vector<double> massVertex((size_t)nV);
    for (int v = 0; v < nV; ++v) massVertex[v] = massDiag[3 * (size_t)v];
    const std::vector<int> noElem;
    const std::vector<double> noVol;
    be.set_system_terms(flow ? tri3 : noElem, flow ? vol : noVol, h, massVertex);
    if (!flow) {
        std::vector<double> ib3(3 * (size_t)Elem.size()), lam((size_t)Elem.size()), mu((size_t)Elem.size());
        for (int e = 0; e < Elem.size(); ++e) {
            const Mat<double, 2>& IB = std::get<0>(elemAttr.rows[e]);
            ib3[3 * (size_t)e] = IB(0, 0); ib3[3 * (size_t)e + 1] = IB(0, 1); ib3[3 * (size_t)e + 2] = IB(1, 1);
            lam[e] = std::get<2>(elasticityAttr.rows[e]); mu[e] = std::get<3>(elasticityAttr.rows[e]);
        }
        std::vector<int> st;
        std::vector<double> info;
        double k = 0;
        if (in.bendingStiffMult && elemAttr.size()) { // BENDING.h:53-54
            k = in.bendingStiffMult * std::get<1>(elemAttr.rows[0])(0, 0);
            st.resize(4 * edgeStencil.size()); info.resize(3 * edgeStencil.size());
            for (size_t e = 0; e < edgeStencil.size(); ++e) {
                for (int i = 0; i < 4; ++i) st[4 * e + i] = edgeStencil[e][i];
                for (int i = 0; i < 3; ++i) info[3 * e + i] = edgeInfo[e][i];
            }
        }
        be.set_elastic_terms(tri3, ib3, vol, lam, mu, st, info, k, h);
    }
    else be.set_elastic_terms(noElem, noVol, noVol, noVol, noVol, noElem, noVol, 0.0, h);
    be.set_positions(s.x.data());
    std::cout << "surface primitives found" << std::endl;

    double DBCAlpha = 1;
    if (withCollision) {
        DBCAlpha = be.ccd(DBCDisp.data(), thickness, DBCAlpha);
        printf("DBCAlpha under contact: %le\n", DBCAlpha);
    }
    StepPotential pot;
    pot.flow = flow; pot.be = &be; pot.Elem = &Elem; pot.fcr = &elasticityAttr; pot.massDiag = &massDiag; pot.DBC = &DBC; pot.nodeAttr = &nodeAttr; pot.h = h;
    double DBCPenaltyXn = 0;
    if (DBCAlpha == 1) {
        for (const auto& r : DBC.rows) {
            const Vec<double, 4>& dI = std::get<0>(r);
            for (int d = 0; d < 3; ++d) s.x[3 * (size_t)((int)dI[0]) + d] = dI[d + 1];
        }
        printf("DBC handled\n");
    }
    else { // the targets cannot be reached without intersection: augmented-Lagrangian penalty (:413-417, DIRICHLET.h)
        printf("moved DBC by %le, turn on Augmented Lagrangian\n", DBCAlpha);
        pot.DBCStiff = 1e6;
        DBCPenaltyXn = pot.dbc_dist2(xn, false);
    }
    // the matrix terms that depend on the penalty: diagonal k m_v on the Dirichlet nodes (DIRICHLET.h:73-94) and the
    // projection mask (all Dirichlet nodes, or only those that do not move while the penalty is active; INC_POTENTIAL.h:387-394)
    auto set_penalty_terms = [&]() {
        std::vector<double> mv = massVertex;
        if (pot.DBCStiff) for (int v = 0; v < nV; ++v) if (dbcMask[v]) mv[v] += pot.DBCStiff * std::get<3>(nodeAttr.rows[v]);
        be.set_system_terms(flow ? tri3 : noElem, flow ? vol : noVol, h, mv);
    };
    if (pot.DBCStiff) set_penalty_terms();

    auto total_energy = [&]() { // Compute_IncPotential (+ Compute_DBC_Energy) at the backend's current positions
        double E = pot.energy(s.x, s.xtilde);
        if (withCollision) be.barrier_energy(dHat2, kappa[0], thickness, E);
        if (pot.friction) be.friction_energy(E);
        pot.add_dbc_energy(s.x, E);
        return E;
    };
    // gradient of the same potential with the Dirichlet handling of :459-482, rhs = -g
    auto gradient_and_rhs = [&]() {
        pot.gradient(s.x, s.xtilde, s.g);
        if (withCollision) be.barrier_gradient(dHat2, kappa[0], thickness, s.g.data());
        if (pot.friction) be.friction_gradient(s.g.data());
        if (pot.DBCStiff) {
            for (const auto& r : DBC.rows) {
                const Vec<double, 4>& dI = std::get<0>(r);
                const int v = (int)dI[0];
                const double km = pot.DBCStiff * std::get<3>(nodeAttr.rows[v]);
                for (int d = 0; d < 3; ++d) s.g[3 * (size_t)v + d] += km * (s.x[3 * (size_t)v + d] - dI[d + 1]);
            }
            for (int v = 0; v < nV; ++v)
                if (dbcFixed[v]) s.g[3 * (size_t)v] = s.g[3 * (size_t)v + 1] = s.g[3 * (size_t)v + 2] = 0;
        }
        else {
            for (int v = 0; v < nV; ++v)
                if (DBCb[v]) s.g[3 * (size_t)v] = s.g[3 * (size_t)v + 1] = s.g[3 * (size_t)v + 2] = 0;
            std::cout << "project rhs for Dirichlet boundary condition " << DBC.size() << std::endl;
        }
        for (size_t i = 0; i < n3; ++i) s.rhs[i] = -s.g[i];
    };

    int PNIter = 0;
    double L2Norm = 0;
    bool useGD = false;
    std::vector<int> rowsPrev;
    std::vector<double> infoPrev, dist2Prev;
    printf("computing initial energy\n");
    be.set_positions(s.x.data());
    if (withCollision) s.nRows = be.constraint_set(dHat2, thickness);
    // per-component coefficients (Compute_Friction_Coef) are in force when muComp is an nComp x nComp table; mu is then 1 (:435-438)
    const bool perComp = !in.muComp.empty() && in.muComp.size() == in.compNodeRange.size() * in.compNodeRange.size();
    const double muEff = perComp ? 1.0 : in.mu;
    pot.friction = withCollision && muEff > 0;
    if (pot.friction) { // lagged friction: basis and normal forces from the state the step starts in (:432-439)
        be.friction_set_components(perComp ? in.compNodeRange : std::vector<int>(), perComp ? in.muComp : std::vector<double>());
        be.friction_set(xn.data(), in.epsv2 * h * h, muEff);
        be.friction_update(dHat2, kappa[0], thickness);
    }
    else be.friction_set(nullptr, 0.0, 0.0);
    s.rhs.resize(n3);
    s.sol.resize(n3);
    s.Eprev = total_energy();
    printf("entering Newton loop\n");
    std::deque<double> resRecord, MDBCProgress;
    int fricIterI = 0;
    const int nFree = nV - DBC.size();
    do {
        gradient_and_rhs(); // :459-491

        // Hessian + search direction (:493-556)
        if (useGD) {
            printf("use gradient descent\n");
            s.sol = s.rhs;
        }
        else if (!be.solve_newton_system(dHat2, kappa[0], thickness, pot.DBCStiff ? &dbcFixed : nullptr, s.rhs.data(), s.sol.data())) {
            FILE* out = fopen((outputFolder + "/Hessian_info.txt").c_str(), "a+");
            if (out) { fprintf(out, "Hessian not SPD in PNIter%d\n", PNIter); fclose(out); }
            useGD = true;
            printf("use gradient descent\n");
            s.sol = s.rhs;
        }

        double alpha, feasibleAlpha;
        step_line_search(be, pot, s, withCollision, dHat2, kappa[0], thickness, alpha, feasibleAlpha);

        // kappa adaptation (:568-598): only rows that were closer than 1e-18 can trigger it, so the previous rows are
        // re-evaluated only when such a row exists (same outcome, no hand-over otherwise)
        if (!rowsPrev.empty()) {
            std::vector<int> rowsCur;
            std::vector<double> infoCur, cur;
            be.get_rows(rowsCur, infoCur);
            be.set_rows(rowsPrev, infoPrev);
            double m;
            be.min_dist2(thickness, &cur, m);
            be.set_rows(rowsCur, infoCur);
            bool updateKappa = false;
            for (size_t i = 0; i < cur.size(); ++i)
                if (dist2Prev[i] < 1e-18 && cur[i] < dist2Prev[i]) { updateKappa = true; break; }
            if (updateKappa && kappa[0] < kappa[1]) {
                kappa[0] *= 2;
                kappaVec[0] *= 2;
                s.Eprev = total_energy();
            }
        }
        rowsPrev.clear(); infoPrev.clear(); dist2Prev.clear();
        if (s.nRows) {
            double minDist2 = 0;
            be.min_dist2(thickness, &dist2Prev, minDist2);
            printf("minDist2 = %le, kappa = %le (max %le)\n", minDist2, kappa[0], kappa[1]);
            bool any = false;
            for (double d : dist2Prev) any = any || d < 1e-18;
            if (any) be.get_rows(rowsPrev, infoPrev);
        }

        // stopping criteria (:647-676)
        double maxRes = 0.0, avgResMag = 0.0;
        L2Norm = 0.0;
        for (int v = 0; v < nV; ++v) {
            double cur = 0;
            for (int d = 0; d < 3; ++d) {
                const double c = s.sol[3 * (size_t)v + d];
                cur += c * c;
                maxRes = std::max(maxRes, std::abs(c));
                L2Norm += c * c;
            }
            avgResMag += std::sqrt(cur);
        }
        avgResMag /= nFree; avgResMag /= h;
        maxRes /= h;
        L2Norm = std::sqrt(L2Norm / nFree) / h;
        printf("PNIter%d: Newton res = %le, tol = %le\n", PNIter++, L2Norm, NewtonTol);
        FILE* out = fopen((outputFolder + "/residual.txt").c_str(), "a+");
        if (out) { fprintf(out, "%d %le %le %le %le\n", PNIter, avgResMag, maxRes, s.Eprev, L2Norm); fclose(out); }
        resRecord.push_back(L2Norm);
        if (resRecord.size() > 3) resRecord.pop_front();
        L2Norm = *std::max_element(resRecord.begin(), resRecord.end());
        if (useGD) L2Norm = NewtonTol * 10;
        if (alpha * 2 < 1e-8 && feasibleAlpha > 1e-8) {
            if (!useGD) {
                useGD = true;
                double gp = 0, gg = 0, pp = 0;
                for (size_t i = 0; i < n3; ++i) { gp += s.rhs[i] * s.sol[i]; gg += s.rhs[i] * s.rhs[i]; pp += s.sol[i] * s.sol[i]; }
                printf("-gdotp = %le, -gpcos = %le\n", gp, gp / std::sqrt(gg * pp));
            }
            else printf("GD tiny step size!\n");
        }
        else useGD = false;

        // progress of the penalised Dirichlet nodes towards their targets (:701-760)
        if (pot.DBCStiff) {
            const double progress = 1 - std::sqrt(pot.dbc_dist2(s.x, false) / DBCPenaltyXn);
            printf("MDBC progress: %le, DBCStiff %le\n", progress, pot.DBCStiff);
            MDBCProgress.push_back(progress);
            if (MDBCProgress.size() > 4) MDBCProgress.pop_front();
            if (progress < 0.99) {
                if (L2Norm < NewtonTol * 10 && pot.DBCStiff < 1e8) {
                    pot.DBCStiff *= 2;
                    set_penalty_terms();
                    s.Eprev = total_energy();
                    printf("updated DBCStiff to %le\n", pot.DBCStiff);
                }
                L2Norm = NewtonTol * 10; // ensures not exit Newton loop
            }
            else {
                pot.DBCStiff = 0;
                set_penalty_terms();
                s.Eprev = total_energy();
                printf("DBC moved to target, turn off Augmented Lagrangian\n");
            }
        }

        // converged with the lagged friction forces: refresh them and check the residual again (:762-848)
        if (resRecord.size() == 3 && L2Norm <= NewtonTol) {
            ++fricIterI;
            if ((fricIterI < in.fricIterAmt || in.fricIterAmt <= 0) && pot.friction) {
                be.friction_update(dHat2, kappa[0], thickness);
                gradient_and_rhs();
                if (!be.solve_newton_system(dHat2, kappa[0], thickness, pot.DBCStiff ? &dbcFixed : nullptr, s.rhs.data(), s.sol.data())) {
                    FILE* fo = fopen((outputFolder + "/Hessian_info.txt").c_str(), "a+");
                    if (fo) { fprintf(fo, "Hessian not SPD in PNIter%d\n", PNIter); fclose(fo); }
                    exit(-1);
                }
                L2Norm = 0.0;
                for (size_t i = 0; i < n3; ++i) L2Norm += s.sol[i] * s.sol[i];
                L2Norm = std::sqrt(L2Norm / nFree) / h;
                printf("friction updated Newton res = %le, tol = %le\n", L2Norm, NewtonTol);
                if (L2Norm > NewtonTol) s.Eprev = total_energy();
            }
        }

        if (flow && (!withCollision || s.nRows == 0)) break; // the flow variant leaves after one iteration without contact (:850-854)
    } while (resRecord.size() < 3 || L2Norm > NewtonTol);

    FILE* out = fopen((outputFolder + "/counter.txt").c_str(), "a+");
    if (out) {
        fprintf(out, "%d", PNIter);
        if (withCollision) fprintf(out, " %lu", (unsigned long)s.nRows);
        fprintf(out, "\n");
        fclose(out);
    }
    if (withCollision) printf("contact #: %lu\n", (unsigned long)s.nRows);

    if (Elem.size()) {
        double maxs, avgs, minc, avgc;
        max_and_avg_stretch(Elem, DBCb, s.x.data(), elemAttr, maxs, avgs, minc, avgc);
        printf("maxs = %le, avgs = %le\n", maxs, avgs);
        out = fopen((outputFolder + "/stretch.txt").c_str(), "a+");
        if (out) { fprintf(out, "%le %le %le %le\n", maxs, avgs, minc, avgc); fclose(out); }
    }

    // hand the state back: X, velocity (:879-886), the last gradient
    for (int v = 0; v < nV; ++v) {
        Vec<double, 3>& xv = std::get<0>(X.rows[v]);
        Vec<double, 3>& vel = std::get<1>(nodeAttr.rows[v]);
        Vec<double, 3>& gv = std::get<2>(nodeAttr.rows[v]);
        for (int d = 0; d < 3; ++d) {
            xv[d] = s.x[3 * (size_t)v + d];
            vel[d] = (s.x[3 * (size_t)v + d] - xn[3 * (size_t)v + d]) / h;
            gv[d] = s.g[3 * (size_t)v + d];
        }
    }
    return PNIter;
}

} // namespace jgsl
