// ORACLE -- TEST INFRASTRUCTURE ONLY (never imported by idp_b200/). CPU restatement of the elastic terms of the discrete-shell
// Newton system:
//   membrane energy  Library/FEM/Shell/MEMBRANE.h:8-52 (useNH = true), gradient :54-122, Hessian + makePD :124-315
//   hinge energy     Library/FEM/Shell/BENDING.h:52-80 (KL = false), gradient :176-213, Hessian + makePD :438-497,
//                    dihedral angle Library/Math/DIHEDRAL_ANGLE.h:9-24
// The ENERGIES are written as the reference defines them; gradients and Hessians are obtained from those expressions by
// second-order forward-mode jets (orc_deriv.hpp), independently of the closed forms the CUDA path uses; the PSD projection
// is the oracle's cyclic-Jacobi makePD (Math/UTILS.h:9-27). Parity status: PINNED by the reference's own code compiled in
// oracle/_ref (tests/test_elastic_host.py): FEM/Shell/MEMBRANE.h and BENDING.h (KL = false) in libidp_ref_shell.so -- energy,
// gradient and PSD-projected Hessian triplets of a deformed mesh with a Dirichlet mask within 1e-10 -- and Math/DIHEDRAL_ANGLE.h
// in libidp_ref.so (angle bit for bit, gradient / Hessian 1e-10). Those builds use the repository's Eigen stand-in (no Eigen
// in this image), like every other reference build here.
#pragma once
#include "orc_deriv.hpp"
#include <cmath>

namespace orc {

template <int N> static inline Jet<N> jscale(double s, const Jet<N>& a)
{
    Jet<N> r; r.v = s * a.v;
    for (int i = 0; i < N; ++i) { r.g[i] = s * a.g[i]; for (int j = 0; j < N; ++j) r.h[i][j] = s * a.h[i][j]; }
    return r;
}
// f(a) for scalar f with first / second derivative f1, f2 at a.v
template <int N> static inline Jet<N> jchain(const Jet<N>& a, double f0, double f1, double f2)
{
    Jet<N> r; r.v = f0;
    for (int i = 0; i < N; ++i) { r.g[i] = f1 * a.g[i]; for (int j = 0; j < N; ++j) r.h[i][j] = f1 * a.h[i][j] + f2 * a.g[i] * a.g[j]; }
    return r;
}
template <int N> static inline Jet<N> jlog(const Jet<N>& a) { return jchain(a, std::log(a.v), 1.0 / a.v, -1.0 / (a.v * a.v)); }
template <int N> static inline Jet<N> jsqrt(const Jet<N>& a)
{
    const double s = std::sqrt(a.v);
    return jchain(a, s, 0.5 / s, -0.25 / (s * a.v));
}
// atan2(y, x): a smooth branch of the angle whose VALUE is replaced by the caller where the reference uses acos
template <int N> static inline Jet<N> jatan2(const Jet<N>& y, const Jet<N>& x)
{
    const double r2 = x.v * x.v + y.v * y.v, r4 = r2 * r2;
    const double fy = x.v / r2, fx = -y.v / r2;
    const double fyy = -2.0 * x.v * y.v / r4, fxx = 2.0 * x.v * y.v / r4, fxy = (y.v * y.v - x.v * x.v) / r4;
    Jet<N> r; r.v = std::atan2(y.v, x.v);
    for (int i = 0; i < N; ++i) {
        r.g[i] = fy * y.g[i] + fx * x.g[i];
        for (int j = 0; j < N; ++j)
            r.h[i][j] = fy * y.h[i][j] + fx * x.h[i][j] + fyy * y.g[i] * y.g[j] + fxx * x.g[i] * x.g[j] + fxy * (y.g[i] * x.g[j] + x.g[i] * y.g[j]);
    }
    return r;
}

// Compute_Dihedral_Angle (DIHEDRAL_ANGLE.h:9-24)
static inline double dihedral_angle(const V3& v0, const V3& v1, const V3& v2, const V3& v3)
{
    const V3 n1 = cross(v1 - v0, v2 - v0), n2 = cross(v2 - v3, v1 - v3);
    double a = std::acos(std::max(-1.0, std::min(1.0, dot(n1, n2) / std::sqrt(sqn(n1) * sqn(n2)))));
    if (dot(cross(n2, n1), v1 - v2) < 0) a = -a;
    return a;
}

// hinge energy h^2 k (theta - thetabar)^2 ebar / hbar (BENDING.h:77) with g (12) and H (12x12, row major); coef = h^2 k ebar / hbar
static inline double hinge_EgH(const V3* x, double thetabar, double coef, bool projectSPD, double* g, double* H)
{
    auto X0 = jvar<12>(x[0], 0), X1 = jvar<12>(x[1], 1), X2 = jvar<12>(x[2], 2), X3 = jvar<12>(x[3], 3);
    auto n1 = jcross(X1 - X0, X2 - X0), n2 = jcross(X2 - X3, X1 - X3);
    auto e = X1 - X2;
    // sin(theta) |n1||n2| = (n2 x n1).(x1 - x2)/|x1 - x2| (the orientation test of the reference), cos(theta) |n1||n2| = n1.n2
    auto th = jatan2(jdot(jcross(n2, n1), e) / jsqrt(jdot(e, e)), jdot(n1, n2));
    th.v = dihedral_angle(x[0], x[1], x[2], x[3]); // the value as the reference computes it
    Jet<12> d = th;
    d.v -= thetabar;
    auto W = jscale(coef, d * d);
    if (g) for (int i = 0; i < 12; ++i) g[i] = W.g[i];
    if (H) {
        for (int i = 0; i < 12; ++i) for (int j = 0; j < 12; ++j) H[i * 12 + j] = W.h[i][j];
        if (projectSPD) make_pd(12, H);
    }
    return W.v;
}

// membrane energy h^2 vol (mu/2 (tr(IB^-1 A) - 2 - 2 lnJ) + lambda/2 lnJ^2) (MEMBRANE.h:43-44); ib = (IB00, IB01, IB11) not inverted,
// coef = h^2 vol; returns false (nothing written) for det IB == 0 (:30-31)
static inline bool membrane_EgH(const V3* x, const double* ib, double coef, double lambda, double mu, bool projectSPD, double* E, double* g, double* H)
{
    const double detB = ib[0] * ib[2] - ib[1] * ib[1];
    if (detB == 0.0) return false;
    const double B00 = ib[2] / detB, B01 = -ib[1] / detB, B11 = ib[0] / detB;
    auto X1 = jvar<9>(x[0], 0), X2 = jvar<9>(x[1], 1), X3 = jvar<9>(x[2], 2);
    auto e01 = X2 - X1, e02 = X3 - X1;
    auto A00 = jdot(e01, e01), A01 = jdot(e01, e02), A11 = jdot(e02, e02);
    auto detA = A00 * A11 - A01 * A01;
    auto lnJ = jscale(0.5, jlog(jscale(B00 * B11 - B01 * B01, detA)));
    auto tr = jscale(B00, A00) + jscale(2.0 * B01, A01) + jscale(B11, A11);
    Jet<9> two(2.0);
    auto W = jscale(coef, jscale(0.5 * mu, tr - two - jscale(2.0, lnJ)) + jscale(0.5 * lambda, lnJ * lnJ));
    if (E) *E = W.v;
    if (g) for (int i = 0; i < 9; ++i) g[i] = W.g[i];
    if (H) {
        for (int i = 0; i < 9; ++i) for (int j = 0; j < 9; ++j) H[i * 9 + j] = W.h[i][j];
        if (projectSPD) make_pd(9, H);
    }
    return true;
}

} // namespace orc
