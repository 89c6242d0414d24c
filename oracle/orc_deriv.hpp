// ORACLE — TEST INFRASTRUCTURE ONLY (see orc_math.hpp header).
//
// Gradients / Hessians of the squared distances and of the edge-edge mollifier, and makePD.
//
// The reference evaluates these with MATLAB-generated scalar code (Math/Distance/POINT_EDGE.h:61-113,
// 267-587; POINT_TRIANGLE.h:25-86,104-549; EDGE_EDGE.h:25-104,122-732; EDGE_EDGE_MOLLIFIER.h:21-79,
// 98-366). That code is not restated line by line: the functions below differentiate the SAME
// closed-form squared distances (orc_math.hpp) analytically, and a second, independent evaluation by
// second-order forward-mode automatic differentiation ("jets", bottom of this file) differentiates the
// reference's distance expressions literally. tests/ checks closed form == jets == the reference's
// generated code (through oracle/_ref and tests/golden/).
//
// Layout: gradients are stacked per vertex in the stencil order of SURVEY.md A.1, Hessians are dense
// row-major n x n (symmetric, so identical to the reference's column-major storage).
#pragma once
#include "orc_math.hpp"

namespace orc {

// skew(a) b = a x b
static inline void skew(const V3& a, double S[3][3])
{
    S[0][0] = 0;    S[0][1] = -a.z; S[0][2] = a.y;
    S[1][0] = a.z;  S[1][1] = 0;    S[1][2] = -a.x;
    S[2][0] = -a.y; S[2][1] = a.x;  S[2][2] = 0;
}
static inline double comp(const V3& a, int i) { return i == 0 ? a.x : (i == 1 ? a.y : a.z); }

// ---------------------------------------------------------------------------------------------
// d = A^2 / B with A = w . (u x v), B = |u x v|^2 in reduced coordinates y = (w, u, v) in R^9.
// Point-triangle: w = p - t0, u = t1 - t0, v = t2 - t0.  Edge-edge: w = b0 - a0, u = a1 - a0, v = b1 - b0.
// ---------------------------------------------------------------------------------------------
static inline void triple_quotient(const V3& w, const V3& u, const V3& v, double g9[9], double H9[9][9])
{
    const V3 n = cross(u, v);
    const double A = dot(w, n), B = sqn(n), r = A / B;
    // grad A = (n, v x w, w x u); grad B = (0, 2 v x n, 2 n x u)
    const V3 Au = cross(v, w), Av = cross(w, u);
    const V3 Bu = 2.0 * cross(v, n), Bv = 2.0 * cross(n, u);
    double gA[9] = {n.x, n.y, n.z, Au.x, Au.y, Au.z, Av.x, Av.y, Av.z};
    double gB[9] = {0, 0, 0, Bu.x, Bu.y, Bu.z, Bv.x, Bv.y, Bv.z};
    double q[9];
    for (int i = 0; i < 9; ++i) {
        g9[i] = 2.0 * r * gA[i] - r * r * gB[i];
        q[i] = gA[i] - r * gB[i];
    }
    // Hessian of A: blocks [w,u] = -skew(v), [w,v] = skew(u), [u,v] = -skew(w) (+ transposes)
    double Sv[3][3], Su[3][3], Sw[3][3], Sn[3][3];
    skew(v, Sv); skew(u, Su); skew(w, Sw); skew(n, Sn);
    double HA[9][9] = {}, HB[9][9] = {};
    const double uu = sqn(u), vv = sqn(v), uv = dot(u, v);
    for (int i = 0; i < 3; ++i) {
        for (int j = 0; j < 3; ++j) {
            HA[i][3 + j] = -Sv[i][j]; HA[3 + j][i] = -Sv[i][j];
            HA[i][6 + j] = Su[i][j];  HA[6 + j][i] = Su[i][j];
            HA[3 + i][6 + j] = -Sw[i][j]; HA[6 + j][3 + i] = -Sw[i][j];
            const double dij = (i == j) ? 1.0 : 0.0;
            HB[3 + i][3 + j] = 2.0 * (vv * dij - comp(v, i) * comp(v, j));
            HB[6 + i][6 + j] = 2.0 * (uu * dij - comp(u, i) * comp(u, j));
            // [u,v] = 2 skew(v) skew(u) - 2 skew(n) = 2 (u v^T - (u.v) I) - 2 skew(n)
            const double buv = 2.0 * (comp(u, i) * comp(v, j) - uv * dij) - 2.0 * Sn[i][j];
            HB[3 + i][6 + j] = buv;
            HB[6 + j][3 + i] = buv;
        }
    }
    for (int i = 0; i < 9; ++i)
        for (int j = 0; j < 9; ++j)
            H9[i][j] = (2.0 / B) * q[i] * q[j] + 2.0 * r * HA[i][j] - r * r * HB[i][j];
}

// embed reduced coordinates into vertex coordinates: y_k = sum_i c[k][i] x_i
template <int K, int NV>
static inline void embed(const double c[K][NV], const double* gk, const double* Hk, double* g, double* H)
{
    const int n = 3 * NV, m = 3 * K;
    for (int i = 0; i < NV; ++i)
        for (int a = 0; a < 3; ++a) {
            double s = 0;
            for (int k = 0; k < K; ++k) s += c[k][i] * gk[3 * k + a];
            g[3 * i + a] = s;
        }
    for (int i = 0; i < NV; ++i)
        for (int j = 0; j < NV; ++j)
            for (int a = 0; a < 3; ++a)
                for (int b = 0; b < 3; ++b) {
                    double s = 0;
                    for (int k = 0; k < K; ++k)
                        for (int l = 0; l < K; ++l)
                            s += c[k][i] * c[l][j] * Hk[(3 * k + a) * m + (3 * l + b)];
                    H[(3 * i + a) * n + (3 * j + b)] = s;
                }
}

// point-triangle, stencil (p, t0, t1, t2)
static inline void pt_grad_hess(const V3& p, const V3& t0, const V3& t1, const V3& t2, double g[12], double H[144])
{
    double g9[9], H9[9][9];
    triple_quotient(p - t0, t1 - t0, t2 - t0, g9, H9);
    static const double c[3][4] = {{1, -1, 0, 0}, {0, -1, 1, 0}, {0, -1, 0, 1}};
    embed<3, 4>(c, g9, &H9[0][0], g, H);
}
// edge-edge, stencil (ea0, ea1, eb0, eb1)
static inline void ee_grad_hess(const V3& a0, const V3& a1, const V3& b0, const V3& b1, double g[12], double H[144])
{
    double g9[9], H9[9][9];
    triple_quotient(b0 - a0, a1 - a0, b1 - b0, g9, H9);
    static const double c[3][4] = {{-1, 0, 1, 0}, {-1, 1, 0, 0}, {0, 0, -1, 1}};
    embed<3, 4>(c, g9, &H9[0][0], g, H);
}

// N = |w x u|^2 in reduced coordinates (w, u): gradient and Hessian
static inline void cross_norm2_reduced(const V3& w, const V3& u, double& N, double g6[6], double H6[6][6])
{
    const V3 c = cross(w, u);
    N = sqn(c);
    const V3 gw = 2.0 * cross(u, c), gu = 2.0 * cross(c, w);
    g6[0] = gw.x; g6[1] = gw.y; g6[2] = gw.z; g6[3] = gu.x; g6[4] = gu.y; g6[5] = gu.z;
    double Sc[3][3];
    skew(c, Sc);
    const double uu = sqn(u), ww = sqn(w), wu = dot(w, u);
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            const double dij = (i == j) ? 1.0 : 0.0;
            H6[i][j] = 2.0 * (uu * dij - comp(u, i) * comp(u, j));
            H6[3 + i][3 + j] = 2.0 * (ww * dij - comp(w, i) * comp(w, j));
            // [w,u] = 2 skew(u) skew(w) - 2 skew(c) = 2 (w u^T - (w.u) I) - 2 skew(c)
            const double b = 2.0 * (comp(w, i) * comp(u, j) - wu * dij) - 2.0 * Sc[i][j];
            H6[i][3 + j] = b;
            H6[3 + j][i] = b;
        }
}

// point-edge, stencil (p, e0, e1): d = |w x u|^2 / |u|^2, w = p - e0, u = e1 - e0
static inline void pe_grad_hess(const V3& p, const V3& e0, const V3& e1, double g[9], double H[81])
{
    const V3 w = p - e0, u = e1 - e0;
    double N, gN[6], HN[6][6];
    cross_norm2_reduced(w, u, N, gN, HN);
    const double B = sqn(u);
    double gB[6] = {0, 0, 0, 2.0 * u.x, 2.0 * u.y, 2.0 * u.z};
    double g6[6], H6[6][6];
    for (int i = 0; i < 6; ++i) g6[i] = gN[i] / B - (N / (B * B)) * gB[i];
    for (int i = 0; i < 6; ++i)
        for (int j = 0; j < 6; ++j) {
            const double hb = (i >= 3 && i == j) ? 2.0 : 0.0;
            H6[i][j] = HN[i][j] / B - (gN[i] * gB[j] + gB[i] * gN[j]) / (B * B) - (N / (B * B)) * hb
                + (2.0 * N / (B * B * B)) * gB[i] * gB[j];
        }
    static const double c[2][3] = {{1, -1, 0}, {0, -1, 1}};
    embed<2, 3>(c, g6, &H6[0][0], g, H);
}

// point-point, stencil (a, b) — Math/Distance/POINT_POINT.h:19-41
static inline void pp_grad_hess(const V3& a, const V3& b, double g[6], double H[36])
{
    const V3 d = 2.0 * (a - b);
    g[0] = d.x; g[1] = d.y; g[2] = d.z; g[3] = -d.x; g[4] = -d.y; g[5] = -d.z;
    for (int i = 0; i < 36; ++i) H[i] = 0;
    for (int i = 0; i < 6; ++i) H[i * 6 + i] = 2.0;
    for (int i = 0; i < 3; ++i) { H[i * 6 + 3 + i] = -2.0; H[(3 + i) * 6 + i] = -2.0; }
}

// c = |u x v|^2, u = a1 - a0, v = b1 - b0, stencil (a0, a1, b0, b1) — EDGE_EDGE_MOLLIFIER.h:10-18 + derivatives
static inline void ee_cross_norm2_grad_hess(const V3& a0, const V3& a1, const V3& b0, const V3& b1, double g[12], double H[144])
{
    double N, g6[6], H6[6][6];
    cross_norm2_reduced(a1 - a0, b1 - b0, N, g6, H6);
    static const double c[2][4] = {{-1, 1, 0, 0}, {0, 0, -1, 1}};
    embed<2, 4>(c, g6, &H6[0][0], g, H);
}

// e, grad e, hess e — EDGE_EDGE_MOLLIFIER.h:461-524
static inline void ee_mollifier_all(const V3& a0, const V3& a1, const V3& b0, const V3& b1, double eps_x,
    double& e, double ge[12], double He[144])
{
    const double c = ee_cross_norm2(a0, a1, b0, b1);
    if (c < eps_x) {
        e = eem(c, eps_x);
        const double qg = eem_g(c, eps_x), qh = eem_h(c, eps_x);
        double gc[12];
        ee_cross_norm2_grad_hess(a0, a1, b0, b1, gc, He);
        for (int i = 0; i < 12; ++i)
            for (int j = 0; j < 12; ++j) He[i * 12 + j] = He[i * 12 + j] * qg + (qh * gc[i]) * gc[j];
        for (int i = 0; i < 12; ++i) ge[i] = gc[i] * qg;
    }
    else {
        e = 1.0;
        for (int i = 0; i < 12; ++i) ge[i] = 0;
        for (int i = 0; i < 144; ++i) He[i] = 0;
    }
}

// ---------------------------------------------------------------------------------------------
// makePD — Math/UTILS.h:9-27. Symmetric eigen-decomposition (cyclic Jacobi here; the reference uses
// Eigen::SelfAdjointEigenSolver = tridiagonal QR; both converge to machine precision), eigenvalues
// ascending, unchanged return when lambda_min >= 0, otherwise negative eigenvalues zeroed and
// H = V diag(lambda) V^T. Only the lower triangle of the input is read.
// ---------------------------------------------------------------------------------------------
static inline void sym_eig_jacobi(int n, const double* Ain, double* lam, double* V)
{
    double A[144];
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) A[i * n + j] = (i >= j) ? Ain[i * n + j] : Ain[j * n + i];
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) V[i * n + j] = (i == j) ? 1.0 : 0.0;
    for (int sweep = 0; sweep < 100; ++sweep) {
        double off = 0, tot = 0;
        for (int i = 0; i < n; ++i)
            for (int j = 0; j < n; ++j) {
                tot += A[i * n + j] * A[i * n + j];
                if (i != j) off += A[i * n + j] * A[i * n + j];
            }
        if (off <= 1e-34 * tot || off == 0) break;
        for (int p = 0; p < n - 1; ++p)
            for (int q = p + 1; q < n; ++q) {
                const double apq = A[p * n + q];
                if (apq == 0) continue;
                const double theta = (A[q * n + q] - A[p * n + p]) / (2.0 * apq);
                const double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
                const double c = 1.0 / std::sqrt(t * t + 1.0), s = t * c;
                for (int k = 0; k < n; ++k) {
                    const double akp = A[k * n + p], akq = A[k * n + q];
                    A[k * n + p] = c * akp - s * akq;
                    A[k * n + q] = s * akp + c * akq;
                }
                for (int k = 0; k < n; ++k) {
                    const double apk = A[p * n + k], aqk = A[q * n + k];
                    A[p * n + k] = c * apk - s * aqk;
                    A[q * n + k] = s * apk + c * aqk;
                }
                for (int k = 0; k < n; ++k) {
                    const double vkp = V[k * n + p], vkq = V[k * n + q];
                    V[k * n + p] = c * vkp - s * vkq;
                    V[k * n + q] = s * vkp + c * vkq;
                }
            }
    }
    // sort ascending
    int idx[12];
    for (int i = 0; i < n; ++i) idx[i] = i;
    std::sort(idx, idx + n, [&](int a, int b) { return A[a * n + a] < A[b * n + b]; });
    double Vs[144];
    for (int j = 0; j < n; ++j) {
        lam[j] = A[idx[j] * n + idx[j]];
        for (int k = 0; k < n; ++k) Vs[k * n + j] = V[k * n + idx[j]];
    }
    std::memcpy(V, Vs, sizeof(double) * n * n);
}

static inline void make_pd(int n, double* H)
{
    double lam[12], V[144];
    sym_eig_jacobi(n, H, lam, V);
    if (lam[0] >= 0) return;
    for (int i = 0; i < n; ++i) {
        if (lam[i] < 0) lam[i] = 0;
        else break;
    }
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) {
            double s = 0;
            for (int k = 0; k < n; ++k) s += V[i * n + k] * lam[k] * V[j * n + k];
            H[i * n + j] = s;
        }
}

// ---------------------------------------------------------------------------------------------
// Second-order forward-mode jets over N variables: independent check of the closed forms above.
// ---------------------------------------------------------------------------------------------
template <int N>
struct Jet {
    double v;
    double g[N];
    double h[N][N];
    Jet() : v(0) { std::memset(g, 0, sizeof(g)); std::memset(h, 0, sizeof(h)); }
    explicit Jet(double c) : Jet() { v = c; }
    static Jet var(double val, int i) { Jet r; r.v = val; r.g[i] = 1.0; return r; }
};
template <int N> static inline Jet<N> operator+(const Jet<N>& a, const Jet<N>& b)
{
    Jet<N> r; r.v = a.v + b.v;
    for (int i = 0; i < N; ++i) { r.g[i] = a.g[i] + b.g[i]; for (int j = 0; j < N; ++j) r.h[i][j] = a.h[i][j] + b.h[i][j]; }
    return r;
}
template <int N> static inline Jet<N> operator-(const Jet<N>& a, const Jet<N>& b)
{
    Jet<N> r; r.v = a.v - b.v;
    for (int i = 0; i < N; ++i) { r.g[i] = a.g[i] - b.g[i]; for (int j = 0; j < N; ++j) r.h[i][j] = a.h[i][j] - b.h[i][j]; }
    return r;
}
template <int N> static inline Jet<N> operator*(const Jet<N>& a, const Jet<N>& b)
{
    Jet<N> r; r.v = a.v * b.v;
    for (int i = 0; i < N; ++i) {
        r.g[i] = a.g[i] * b.v + a.v * b.g[i];
        for (int j = 0; j < N; ++j) r.h[i][j] = a.h[i][j] * b.v + a.g[i] * b.g[j] + a.g[j] * b.g[i] + a.v * b.h[i][j];
    }
    return r;
}
template <int N> static inline Jet<N> operator/(const Jet<N>& a, const Jet<N>& b)
{
    // a * (1/b)
    Jet<N> inv; inv.v = 1.0 / b.v;
    const double i2 = -inv.v * inv.v, i3 = 2.0 * inv.v * inv.v * inv.v;
    for (int i = 0; i < N; ++i) {
        inv.g[i] = i2 * b.g[i];
        for (int j = 0; j < N; ++j) inv.h[i][j] = i2 * b.h[i][j] + i3 * b.g[i] * b.g[j];
    }
    return a * inv;
}
template <int N> struct JV3 { Jet<N> x, y, z; };
template <int N> static inline JV3<N> operator-(const JV3<N>& a, const JV3<N>& b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
template <int N> static inline Jet<N> jdot(const JV3<N>& a, const JV3<N>& b) { return a.x * b.x + (a.y * b.y + a.z * b.z); }
template <int N> static inline JV3<N> jcross(const JV3<N>& a, const JV3<N>& b)
{
    return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
template <int N> static inline JV3<N> jvar(const V3& p, int vertex)
{
    return {Jet<N>::var(p.x, 3 * vertex), Jet<N>::var(p.y, 3 * vertex + 1), Jet<N>::var(p.z, 3 * vertex + 2)};
}
template <int N> static inline void jet_out(const Jet<N>& d, double* g, double* H)
{
    for (int i = 0; i < N; ++i) { g[i] = d.g[i]; for (int j = 0; j < N; ++j) H[i * N + j] = d.h[i][j]; }
}
static inline double pt_jet(const V3& p, const V3& t0, const V3& t1, const V3& t2, double g[12], double H[144])
{
    auto P = jvar<12>(p, 0), T0 = jvar<12>(t0, 1), T1 = jvar<12>(t1, 2), T2 = jvar<12>(t2, 3);
    auto b = jcross(T1 - T0, T2 - T0);
    auto aTb = jdot(P - T0, b);
    auto d = aTb * aTb / jdot(b, b);
    jet_out(d, g, H);
    return d.v;
}
static inline double ee_jet(const V3& a0, const V3& a1, const V3& b0, const V3& b1, double g[12], double H[144])
{
    auto A0 = jvar<12>(a0, 0), A1 = jvar<12>(a1, 1), B0 = jvar<12>(b0, 2), B1 = jvar<12>(b1, 3);
    auto b = jcross(A1 - A0, B1 - B0);
    auto aTb = jdot(B0 - A0, b);
    auto d = aTb * aTb / jdot(b, b);
    jet_out(d, g, H);
    return d.v;
}
static inline double pe_jet(const V3& p, const V3& e0, const V3& e1, double g[9], double H[81])
{
    auto P = jvar<9>(p, 0), E0 = jvar<9>(e0, 1), E1 = jvar<9>(e1, 2);
    auto c = jcross(E0 - P, E1 - P);
    auto e = E1 - E0;
    auto d = jdot(c, c) / jdot(e, e);
    jet_out(d, g, H);
    return d.v;
}
static inline double eecn2_jet(const V3& a0, const V3& a1, const V3& b0, const V3& b1, double g[12], double H[144])
{
    auto A0 = jvar<12>(a0, 0), A1 = jvar<12>(a1, 1), B0 = jvar<12>(b0, 2), B1 = jvar<12>(b1, 3);
    auto c = jcross(A1 - A0, B1 - B0);
    auto d = jdot(c, c);
    jet_out(d, g, H);
    return d.v;
}

} // namespace orc
