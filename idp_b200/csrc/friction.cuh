// Lagged friction of the contact rows (Library/FEM/FRICTION.h:17-662, FRICTION_UTILS.h).
//
// A friction row is a non-mollified contact row (EE, PT, PE, PP) frozen at the iterate where the basis was computed:
// stencil weights w_i (the closest-point parametrisation), an orthonormal tangent basis T (3x2) and the normal force
// lambda = -b'(d) 2 sqrt(d) * weight. With u = T^T (sum_i w_i dx_i), dx = x - x_n:
//   E = mu lambda k f0(|u|),   g_i = w_i T (mu lambda k f1(|u|)/|u|) u,   H_ij = w_i w_j T M2 T^T,
//   M2 = mu lambda k ( f1/|u| I + (f2_term/|u|) u u^T )  for |u| < eps_v h,   (f1/|u|^3) ubar ubar^T  beyond,   f1/|u| I  at u = 0
// (k = multiplicity of merged PP / PE rows). M2 is positive semi-definite in all three cases (eigenvalues f1/|u| and
// 2 (eps - |u|)/eps^2 inside the clamp), so the reference's 2x2 makePD is the identity and no eigen-solve is needed: every
// 3x3 block of a row is the SAME matrix B = T M2 T^T scaled by w_i w_j.
// The reference reads the SECOND point of a PP row from X instead of X - Xn (FRICTION.h:217,311,498); reproduced (pp_abs).
#pragma once
#include "pair_deriv.cuh"

namespace idp {

struct FricRow {
    int v[4]; int nv;    // nv = 0: the contact row carries no friction (mollified kinds)
    int kind, mult;      // RowKind of the contact row, multiplicity of a merged PP / PE row
    double w[4];
    double t0[3], t1[3]; // tangent basis columns
    double lam;          // normal force x multiplicity (0: not a friction row)
    int pp_abs;          // PP row: relative displacement uses x (not x - xn) of the second point
    double cp[2];        // closest-point parameters as the reference stores them (for inspection)
};

IDP_HD V3 normalized3(const V3& a) { return a / sqrt(sqn3(a)); }

// Compute_Friction_Basis for one decoded contact row (FRICTION.h:49-117); dHat2 already includes the thickness offset
IDP_HD void friction_basis(const RowDec& d, const V3* x, double weight, double dHat2, double kappa, double xi2, FricRow& f)
{
    f.nv = 0; f.lam = 0; f.pp_abs = 0; f.cp[0] = f.cp[1] = 0; f.kind = d.kind; f.mult = d.mult;
    for (int i = 0; i < 4; ++i) { f.v[i] = d.v[i]; f.w[i] = 0; }
    V3 b0, b1;
    double dist2;
    if (d.kind == K_EE) {
        const V3 e20 = x[0] - x[2], e01 = x[1] - x[0], e23 = x[3] - x[2];
        double g1, g2;
        ldlt2_solve(sqn3(e01), -dot3(e23, e01), sqn3(e23), -dot3(e20, e01), dot3(e20, e23), g1, g2);
        f.cp[0] = g1; f.cp[1] = g2;
        f.w[0] = 1.0 - g1; f.w[1] = g1; f.w[2] = g2 - 1.0; f.w[3] = -g2;
        b0 = normalized3(e01);
        b1 = normalized3(cross3(cross3(e01, e23), e01));
        dist2 = dist2_ee(x[0], x[1], x[2], x[3]);
        f.nv = 4;
    }
    else if (d.kind == K_PT) {
        const V3 r0 = x[2] - x[1], r1 = x[3] - x[1], po = x[0] - x[1];
        double be1, be2;
        ldlt2_solve(sqn3(r0), dot3(r1, r0), sqn3(r1), dot3(r0, po), dot3(r1, po), be1, be2);
        f.cp[0] = be1; f.cp[1] = be2;
        f.w[0] = 1.0; f.w[1] = -1.0 + be1 + be2; f.w[2] = -be1; f.w[3] = -be2;
        b0 = normalized3(r0);
        b1 = normalized3(cross3(cross3(r0, r1), r0));
        dist2 = dist2_pt(x[0], x[1], x[2], x[3]);
        f.nv = 4;
    }
    else if (d.kind == K_PE) {
        const V3 e12 = x[2] - x[1];
        const double yita = dot3(x[0] - x[1], e12) / sqn3(e12);
        f.cp[0] = yita;
        f.w[0] = 1.0; f.w[1] = yita - 1.0; f.w[2] = -yita;
        b0 = normalized3(e12);
        b1 = normalized3(cross3(e12, x[0] - x[1]));
        dist2 = dist2_pe(x[0], x[1], x[2]);
        f.nv = 3;
    }
    else if (d.kind == K_PP) {
        const V3 v01 = x[1] - x[0];
        const V3 xc = cross3(mk3(1, 0, 0), v01), yc = cross3(mk3(0, 1, 0), v01);
        if (sqn3(xc) > sqn3(yc)) { b0 = normalized3(xc); b1 = normalized3(cross3(v01, xc)); }
        else { b0 = normalized3(yc); b1 = normalized3(cross3(v01, yc)); }
        f.w[0] = 1.0; f.w[1] = -1.0;
        dist2 = dist2_pp(x[0], x[1]);
        f.nv = 2;
        f.pp_abs = 1;
    }
    else return; // mollified kinds carry no friction (FRICTION.h:37-41)
    f.t0[0] = b0.x; f.t0[1] = b0.y; f.t0[2] = b0.z;
    f.t1[0] = b1.x; f.t1[1] = b1.y; f.t1[2] = b1.z;
    double b, bg, bh;
    barrier_all(dist2 - xi2, dHat2, kappa, b, bg, bh);
    f.lam = -bg * 2.0 * sqrt(dist2) * weight * (double)d.mult;
}

// C1-clamped friction functions (FRICTION_UTILS.h:10-39)
IDP_HD double fric_f0(double x2, double eps) { return x2 >= eps * eps ? sqrt(x2) : x2 * (-sqrt(x2) / 3.0 + eps) / (eps * eps) + eps / 3.0; }
IDP_HD double fric_f1_div(double x2, double eps) { return x2 >= eps * eps ? 1.0 / sqrt(x2) : (-sqrt(x2) + 2.0 * eps) / (eps * eps); }

// relative displacement in the tangent plane; dx[i] = x_i - xn_i (for a PP row the caller passes x_1 itself as dx[1])
IDP_HD void fric_rel(const FricRow& f, const V3* dx, double& u0, double& u1)
{
    V3 r = mk3(0, 0, 0);
    for (int i = 0; i < 4; ++i) if (i < f.nv) r = r + f.w[i] * dx[i];
    u0 = f.t0[0] * r.x + (f.t0[1] * r.y + f.t0[2] * r.z);
    u1 = f.t1[0] * r.x + (f.t1[1] * r.y + f.t1[2] * r.z);
}
IDP_HD double friction_energy(const FricRow& f, const V3* dx, double epsvh, double mu)
{
    double u0, u1;
    fric_rel(f, dx, u0, u1);
    return mu * f.lam * fric_f0(u0 * u0 + u1 * u1, epsvh);
}
// g3: the common 3-vector T (mu lambda f1/|u|) u; vertex i receives w_i g3
IDP_HD void friction_gradient(const FricRow& f, const V3* dx, double epsvh, double mu, double* g3)
{
    double u0, u1;
    fric_rel(f, dx, u0, u1);
    const double k = fric_f1_div(u0 * u0 + u1 * u1, epsvh) * mu * f.lam;
    for (int a = 0; a < 3; ++a) g3[a] = k * (f.t0[a] * u0 + f.t1[a] * u1);
}
// B = T M2 T^T (3x3, row major); block (i, j) of the row's Hessian is w_i w_j B
IDP_HD void friction_hessian_core(const FricRow& f, const V3* dx, double epsvh, double mu, double* B9)
{
    double u0, u1;
    fric_rel(f, dx, u0, u1);
    const double x2 = u0 * u0 + u1 * u1, n = sqrt(x2);
    const double f1d = fric_f1_div(x2, epsvh), c = mu * f.lam;
    double m00, m01, m11;
    if (x2 >= epsvh * epsvh) {
        const double k = c * f1d / x2; // ubar = (-u1, u0)
        m00 = k * u1 * u1; m01 = -k * u0 * u1; m11 = k * u0 * u0;
    }
    else if (n == 0) { m00 = m11 = c * f1d; m01 = 0; }
    else {
        const double f2 = -1.0 / (epsvh * epsvh) / n;
        m00 = c * (f1d + f2 * u0 * u0); m01 = c * (f2 * u0 * u1); m11 = c * (f1d + f2 * u1 * u1);
    }
    for (int a = 0; a < 3; ++a)
        for (int b = 0; b < 3; ++b)
            B9[3 * a + b] = f.t0[a] * (m00 * f.t0[b] + m01 * f.t1[b]) + f.t1[a] * (m01 * f.t0[b] + m11 * f.t1[b]);
}

} // namespace idp
