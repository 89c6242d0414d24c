// ORACLE -- TEST INFRASTRUCTURE ONLY (never imported by idp_b200/). CPU restatement of the lagged friction of the contact rows
// (SURVEY.md 8(f) rank 4), following Library/FEM/FRICTION.h:
//   Compute_Friction_Basis      :17-124  (closest points / tangent bases of FRICTION_UTILS.h:41-66, 107-142, 184-205, 243-260)
//   Compute_Friction_Potential  :172-252, _Gradient :254-379, _Hessian :381-662 (f0 / f1 / f2 of FRICTION_UTILS.h:10-39)
// written as the reference writes it: per row a 2-vector u = T^T relDX, the dense TT (2 x 3nv) map, Hessian TT^T M2 TT with the
// inner 2x2 matrix projected by makePD where the reference projects it. Pinned by the reference's own FRICTION.h compiled in
// oracle/_ref (tests/test_friction.py::test_oracle_friction_matches_reference). The reference's quirk of reading the second point
// of a point-point row from X instead of X - Xn (:217, 311, 498) is kept.
#pragma once
#include "orc_ipc.hpp"

namespace orc {

struct FrictionRow {
    Row row;          // the contact row as it was
    int nv; int v[4]; // stencil
    double cp[2];     // closest-point parameters
    V3 t0, t1;        // tangent basis columns
    double lam;       // normal force (without multiplicity)
};

static inline V3 normalized(const V3& a) { return a / std::sqrt(sqn(a)); }

static inline void friction_basis(const double* X, const std::vector<Row>& rows, const double* weight, double dHat2, double kappa, double thickness,
    std::vector<FrictionRow>& out)
{
    const double t2 = thickness * thickness;
    dHat2 = dHat2 + 2 * std::sqrt(dHat2) * thickness;
    out.clear();
    for (size_t i = 0; i < rows.size(); ++i) {
        const Row& r = rows[i];
        if (r[0] >= 0 && (r[2] < 0 || r[3] < 0)) continue; // mollified stencils carry no friction (:37-41)
        FrictionRow f;
        f.row = r; f.cp[0] = f.cp[1] = 0;
        double d2;
        if (r[0] >= 0) { // edge-edge
            f.nv = 4; f.v[0] = r[0]; f.v[1] = r[1]; f.v[2] = r[2]; f.v[3] = r[3];
            const V3 v0 = ld3(X + 3 * r[0]), v1 = ld3(X + 3 * r[1]), v2 = ld3(X + 3 * r[2]), v3 = ld3(X + 3 * r[3]);
            const V3 e20 = v0 - v2, e01 = v1 - v0, e23 = v3 - v2;
            ldlt2_solve(sqn(e01), -dot(e23, e01), sqn(e23), -dot(e20, e01), dot(e20, e23), f.cp[0], f.cp[1]);
            f.t0 = normalized(e01); f.t1 = normalized(cross(cross(e01, e23), e01));
            d2 = dist2_ee(v0, v1, v2, v3);
        }
        else {
            const int p = -r[0] - 1;
            const V3 x0 = ld3(X + 3 * p);
            if (r[2] < 0) { // point-point
                f.nv = 2; f.v[0] = p; f.v[1] = r[1]; f.v[2] = f.v[3] = p;
                const V3 v01 = ld3(X + 3 * r[1]) - x0;
                const V3 xc = cross(V3{1, 0, 0}, v01), yc = cross(V3{0, 1, 0}, v01);
                if (sqn(xc) > sqn(yc)) { f.t0 = normalized(xc); f.t1 = normalized(cross(v01, xc)); }
                else { f.t0 = normalized(yc); f.t1 = normalized(cross(v01, yc)); }
                d2 = dist2_pp(x0, ld3(X + 3 * r[1]));
            }
            else if (r[3] < 0) { // point-edge
                f.nv = 3; f.v[0] = p; f.v[1] = r[1]; f.v[2] = r[2]; f.v[3] = p;
                const V3 v1 = ld3(X + 3 * r[1]), v2 = ld3(X + 3 * r[2]), e12 = v2 - v1;
                f.cp[0] = dot(x0 - v1, e12) / sqn(e12);
                f.t0 = normalized(e12); f.t1 = normalized(cross(e12, x0 - v1));
                d2 = dist2_pe(x0, v1, v2);
            }
            else { // point-triangle
                f.nv = 4; f.v[0] = p; f.v[1] = r[1]; f.v[2] = r[2]; f.v[3] = r[3];
                const V3 v1 = ld3(X + 3 * r[1]), v2 = ld3(X + 3 * r[2]), v3 = ld3(X + 3 * r[3]);
                const V3 b0 = v2 - v1, b1 = v3 - v1, po = x0 - v1;
                ldlt2_solve(sqn(b0), dot(b1, b0), sqn(b1), dot(b0, po), dot(b1, po), f.cp[0], f.cp[1]);
                f.t0 = normalized(b0); f.t1 = normalized(cross(cross(b0, b1), b0));
                d2 = dist2_pt(x0, v1, v2, v3);
            }
        }
        f.lam = -barrier_g(d2 - t2, dHat2, kappa) * 2 * std::sqrt(d2) * (weight ? weight[i] : 1.0);
        out.push_back(f);
    }
}

// Compute_Friction_Coef (FRICTION.h:114-170): normal force x coefficient of the (first primitive, opposite primitive) components
static inline bool friction_coef(std::vector<FrictionRow>& rows, const std::vector<int>& compNodeRange, const std::vector<double>& muComp)
{
    auto comp = [&](int v) { for (size_t c = 0; c < compNodeRange.size(); ++c) if (v < compNodeRange[c]) return (int)c; return -1; };
    for (FrictionRow& f : rows) {
        const int c0 = comp(f.v[0]), c1 = comp(f.row[0] >= 0 ? f.v[2] : f.v[1]);
        if (c0 < 0 || c1 < 0) return false; // "can't find node compI"
        f.lam *= muComp[c0 + c1 * compNodeRange.size()];
    }
    return true;
}

// stencil weights of the relative displacement (FRICTION_UTILS.h: *_RelDX / *_TT)
static inline void friction_weights(const FrictionRow& f, double w[4])
{
    w[0] = w[1] = w[2] = w[3] = 0;
    if (f.row[0] >= 0) { w[0] = 1.0 - f.cp[0]; w[1] = f.cp[0]; w[2] = f.cp[1] - 1.0; w[3] = -f.cp[1]; }
    else if (f.nv == 2) { w[0] = 1.0; w[1] = -1.0; }
    else if (f.nv == 3) { w[0] = 1.0; w[1] = f.cp[0] - 1.0; w[2] = -f.cp[0]; }
    else { w[0] = 1.0; w[1] = -1.0 + f.cp[0] + f.cp[1]; w[2] = -f.cp[0]; w[3] = -f.cp[1]; }
}

// E (added), g (nV x 3, added), triplets (appended) at X relative to Xn
static inline void friction_eval(const double* X, const double* Xn, const std::vector<FrictionRow>& rows, double epsvh2, double mu, bool projectSPD,
    double* E, double* g, Triplets* T)
{
    const double eps = std::sqrt(epsvh2);
    for (const FrictionRow& f : rows) {
        double w[4];
        friction_weights(f, w);
        V3 rel{0, 0, 0};
        for (int k = 0; k < f.nv; ++k) {
            const int v = f.v[k];
            V3 dx = ld3(X + 3 * v) - ld3(Xn + 3 * v);
            if (f.nv == 2 && k == 1) dx = ld3(X + 3 * v); // the reference's point-point rows read X here
            rel = rel + w[k] * dx;
        }
        const double u0 = dot(rel, f.t0), u1 = dot(rel, f.t1), x2 = u0 * u0 + u1 * u1, n = std::sqrt(x2);
        const double mult = (f.nv < 4 && f.row[3] < -1) ? (double)(-f.row[3]) : 1.0;
        const double c = mu * f.lam * mult;
        if (E) *E += c * (x2 >= eps * eps ? n : x2 * (-n / 3.0 + eps) / (eps * eps) + eps / 3.0);
        const double f1d = x2 >= eps * eps ? 1.0 / n : (-n + 2.0 * eps) / (eps * eps);
        if (g) {
            const V3 t = (c * f1d) * (u0 * f.t0 + u1 * f.t1);
            for (int k = 0; k < f.nv; ++k) { g[3 * f.v[k]] += w[k] * t.x; g[3 * f.v[k] + 1] += w[k] * t.y; g[3 * f.v[k] + 2] += w[k] * t.z; }
        }
        if (T) {
            double M[4]; // inner 2x2
            if (x2 >= eps * eps) { const double k = f1d / x2; M[0] = k * u1 * u1; M[1] = M[2] = -k * u0 * u1; M[3] = k * u0 * u0; }
            else if (n == 0) { M[0] = M[3] = f1d; M[1] = M[2] = 0; }
            else {
                const double f2 = -1.0 / (eps * eps) / n;
                M[0] = f1d + f2 * u0 * u0; M[1] = M[2] = f2 * u0 * u1; M[3] = f1d + f2 * u1 * u1;
                if (projectSPD) make_pd(2, M);
            }
            const double tb[2][3] = {{f.t0.x, f.t0.y, f.t0.z}, {f.t1.x, f.t1.y, f.t1.z}};
            for (int p = 0; p < f.nv; ++p)
                for (int a = 0; a < 3; ++a)
                    for (int q = 0; q < f.nv; ++q)
                        for (int b = 0; b < 3; ++b) {
                            double s = 0;
                            for (int i = 0; i < 2; ++i) for (int j = 0; j < 2; ++j) s += tb[i][a] * M[2 * i + j] * tb[j][b];
                            T->r.push_back(3 * f.v[p] + a); T->c.push_back(3 * f.v[q] + b); T->v.push_back(c * w[p] * w[q] * s);
                        }
        }
    }
}

} // namespace orc
