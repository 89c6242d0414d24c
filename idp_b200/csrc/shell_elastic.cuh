// Elastic terms of the discrete-shell Newton system evaluated per element in registers, with the same machinery as the
// contact rows (psd_lowrank.cuh): reduced difference coordinates, PSD projection on the translation-free subspace
// (6x6 for a triangle, 9x9 for a hinge), blocks emitted through the row sink.
//
//   membrane  -- Library/FEM/Shell/MEMBRANE.h:8-315 (useNH = true): neo-Hookean energy of the first fundamental form
//                A = J^T J, J = [x2 - x1, x3 - x1]:  W = c ( mu/2 (tr(IB^-1 A) - 2 - 2 lnJ) + lambda/2 lnJ^2 ),
//                lnJ = ln(det A / det IB) / 2, c = h^2 vol.
//   hinge     -- Library/FEM/Shell/BENDING.h:52-80,176-213,438-497 (KL = false) with Math/DIHEDRAL_ANGLE.h:9-24,176-205,1191+:
//                W = c (theta - thetabar)^2, c = h^2 k ebar / hbar.
//
// Formulation (not the reference's): with S = dW/dA the membrane gradient is 2 J S and the reduced Hessian is
//   2 S (x) I3 + c [ (lambda - 4 t1) a a^T + (2 t1 / det A) Q ],  t1 = (-mu + lambda lnJ)/2,  a = vec(J A^-1),
//   Q = the quadratic form of det(dA)  (Q_uu = -v v^T, Q_vv = -u u^T, Q_uv = 2 u v^T - v u^T),
// which follows from d(A^-1) = -A^-1 dA A^-1 and tr(M^2) = tr(M)^2 - 2 det M for 2x2 M. The hinge uses the closed-form
// gradient of the dihedral angle in reduced coordinates (a, e, d) = (x0 - x1, x2 - x1, x3 - x1),
//   dtheta/da = -|e| n1/|n1|^2,  dtheta/dd = -|e| n2/|n2|^2,  dtheta/de = (e.a) n1/(|e||n1|^2) + (e.d) n2/(|e||n2|^2),
//   n1 = e x a, n2 = d x e, and its Jacobian (the angle Hessian) by forward-mode differentiation of that gradient along the
//   nine coordinate directions (dual numbers; the primal parts are shared by the compiler).
// Tolerance against the oracle: 1e-10 relative (tests/test_gpu_elastic.py, tests/test_host.py).
#pragma once
#include "psd_lowrank.cuh"

namespace idp {

// ---- dual numbers (value + one directional derivative) ---------------------------------------------------------------------
struct Du {
    double v, d;
};
IDP_HD Du mkdu(double v, double d) { Du r; r.v = v; r.d = d; return r; }
IDP_HD Du operator+(const Du& a, const Du& b) { return mkdu(a.v + b.v, a.d + b.d); }
IDP_HD Du operator-(const Du& a, const Du& b) { return mkdu(a.v - b.v, a.d - b.d); }
IDP_HD Du operator-(const Du& a) { return mkdu(-a.v, -a.d); }
IDP_HD Du operator*(const Du& a, const Du& b) { return mkdu(a.v * b.v, a.v * b.d + a.d * b.v); }
IDP_HD Du operator/(const Du& a, const Du& b)
{
    const double q = a.v / b.v;
    return mkdu(q, (a.d - q * b.d) / b.v);
}
IDP_HD Du du_sqrt(const Du& a)
{
    const double s = sqrt(a.v);
    return mkdu(s, 0.5 * a.d / s);
}
IDP_HD double du_sqrt(double a) { return sqrt(a); }

template <class S>
struct V3T {
    S x, y, z;
};
template <class S>
IDP_HD V3T<S> v3t(const S& x, const S& y, const S& z) { V3T<S> r; r.x = x; r.y = y; r.z = z; return r; }
template <class S>
IDP_HD S dotT(const V3T<S>& a, const V3T<S>& b) { return a.x * b.x + (a.y * b.y + a.z * b.z); }
template <class S>
IDP_HD V3T<S> crossT(const V3T<S>& a, const V3T<S>& b) { return v3t<S>(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
template <class S>
IDP_HD V3T<S> scaleT(const S& s, const V3T<S>& a) { return v3t<S>(s * a.x, s * a.y, s * a.z); }
template <class S>
IDP_HD V3T<S> addT(const V3T<S>& a, const V3T<S>& b) { return v3t<S>(a.x + b.x, a.y + b.y, a.z + b.z); }

// gradient of the dihedral angle in reduced coordinates; g9 = (d/da, d/de, d/dd)
template <class S>
IDP_HD void hinge_angle_gradient(const V3T<S>& a, const V3T<S>& e, const V3T<S>& d, S* g9)
{
    const V3T<S> n1 = crossT(e, a), n2 = crossT(d, e);
    const S le = du_sqrt(dotT(e, e)), q1 = dotT(n1, n1), q2 = dotT(n2, n2);
    const S ka = -(le / q1), kd = -(le / q2);
    const S k1 = dotT(e, a) / (le * q1), k2 = dotT(e, d) / (le * q2);
    const V3T<S> ga = scaleT(ka, n1), gd = scaleT(kd, n2), ge = addT(scaleT(k1, n1), scaleT(k2, n2));
    g9[0] = ga.x; g9[1] = ga.y; g9[2] = ga.z;
    g9[3] = ge.x; g9[4] = ge.y; g9[5] = ge.z;
    g9[6] = gd.x; g9[7] = gd.y; g9[8] = gd.z;
}

// the angle itself, in the reference's own operation order (Math/DIHEDRAL_ANGLE.h:16-23): acos of the clamped cosine is
// ill-conditioned near 0 and +-pi (errors of 1e-16 in the cosine become 1e-8 in the angle), so the value only agrees with
// the reference's where the same roundings are made; elastic_kernels.cu is compiled with --fmad=false for that reason
IDP_HD double hinge_angle(const V3& v0, const V3& v1, const V3& v2, const V3& v3)
{
    const V3 n1 = cross3(v1 - v0, v2 - v0), n2 = cross3(v2 - v3, v1 - v3);
    double th = acos(fmax(-1.0, fmin(1.0, dot3(n1, n2) / sqrt(sqn3(n1) * sqn3(n2)))));
    if (dot3(cross3(n2, n1), v1 - v2) < 0) th = -th;
    return th;
}

struct ElasticOut {
    double E;
    double g[12];
    bool eigFail;
};

// One hinge (x0; x1, x2; x3): stencil order of edgeStencil (DISCRETE_SHELL.h:169-211). coef = h^2 k ebar / hbar.
template <class VS9, class Emit>
IDP_HD void hinge_eval(const V3* x, double thetabar, double coef, bool projectSPD, bool wantG, bool wantH, VS9& V9, ElasticOut& out, Emit& emit)
{
    out.eigFail = false;
    const V3 a = x[0] - x[1], e = x[2] - x[1], d = x[3] - x[1];
    const double dth = hinge_angle(x[0], x[1], x[2], x[3]) - thetabar;
    out.E = coef * dth * dth;
    if (!wantG && !wantH) return;
    double g9[9];
    hinge_angle_gradient<double>(v3t<double>(a.x, a.y, a.z), v3t<double>(e.x, e.y, e.z), v3t<double>(d.x, d.y, d.z), g9);
    const double K = 2.0 * coef;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const double ga = K * dth * g9[c], ge = K * dth * g9[3 + c], gd = K * dth * g9[6 + c];
        out.g[c] = ga; out.g[6 + c] = ge; out.g[9 + c] = gd; out.g[3 + c] = -(ga + ge + gd);
    }
    if (!wantH) return;
    double H[45];
    const double xv[9] = {a.x, a.y, a.z, e.x, e.y, e.z, d.x, d.y, d.z};
#pragma unroll
    for (int k = 0; k < 9; ++k) {
        Du s[9];
#pragma unroll
        for (int m = 0; m < 9; ++m) s[m] = mkdu(xv[m], m == k ? 1.0 : 0.0);
        Du col[9];
        hinge_angle_gradient<Du>(v3t<Du>(s[0], s[1], s[2]), v3t<Du>(s[3], s[4], s[5]), v3t<Du>(s[6], s[7], s[8]), col);
#pragma unroll
        for (int i = 0; i <= k; ++i) H[SI<9>(i, k)] = (K * dth) * col[i].d + K * g9[i] * g9[k];
    }
    // (w, u, v) = (a, e, d) with base vertex 1: the point-triangle embedding of psd_lowrank.cuh
    const double T[3][3] = {{1, 0, 1}, {1, -1, 0}, {0, -1, 1}};
    double M[45];
    congruence_blocks<3>(T, H, M);
    if (projectSPD && !make_pd_ql<9>(M, V9)) out.eigFail = true;
    const double Hm[4][3] = {{0.5, 0.5, 0.5}, {-0.5, 0.5, -0.5}, {0.5, -0.5, -0.5}, {-0.5, -0.5, 0.5}};
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (!emit.wants(i, j)) continue;
            double blk[9];
            expand_block<4, 3>(Hm, M, i, j, blk);
            emit(i, j, blk);
        }
}

// One membrane triangle (x1, x2, x3 in the reference's naming = x[0..2]); ib = (IB00, IB01, IB11) of the REST first
// fundamental form (not inverted), coef = h^2 vol. Returns false for a degenerate rest triangle (skipped, MEMBRANE.h:30-31).
template <class VS6, class Emit>
IDP_HD bool membrane_eval(const V3* x, const double* ib, double coef, double lambda, double mu, bool projectSPD, bool wantG, bool wantH, VS6& V6,
    ElasticOut& out, Emit& emit)
{
    out.eigFail = false;
    const double detB = ib[0] * ib[2] - ib[1] * ib[1];
    if (detB == 0.0) return false;
    const V3 u = x[1] - x[0], v = x[2] - x[0];
    const double A00 = sqn3(u), A01 = dot3(u, v), A11 = sqn3(v);
    const double detA = A00 * A11 - A01 * A01;
    const double B00 = ib[2] / detB, B01 = -ib[1] / detB, B11 = ib[0] / detB; // IB^-1
    const double lnJ = 0.5 * log(detA * (B00 * B11 - B01 * B01));
    out.E = coef * (0.5 * mu * ((B00 * A00 + 2.0 * B01 * A01 + B11 * A11) - 2.0 - 2.0 * lnJ) + 0.5 * lambda * lnJ * lnJ);
    if (!wantG && !wantH) return true;
    const double I00 = A11 / detA, I01 = -A01 / detA, I11 = A00 / detA; // A^-1
    const double t1 = 0.5 * (-mu + lambda * lnJ);
    const double S00 = coef * (0.5 * mu * B00 + t1 * I00), S01 = coef * (0.5 * mu * B01 + t1 * I01), S11 = coef * (0.5 * mu * B11 + t1 * I11);
    const V3 gu = 2.0 * (S00 * u + S01 * v), gv = 2.0 * (S01 * u + S11 * v);
    out.g[0] = -(gu.x + gv.x); out.g[1] = -(gu.y + gv.y); out.g[2] = -(gu.z + gv.z);
    out.g[3] = gu.x; out.g[4] = gu.y; out.g[5] = gu.z;
    out.g[6] = gv.x; out.g[7] = gv.y; out.g[8] = gv.z;
    if (!wantH) return true;
    const V3 au = I00 * u + I01 * v, av = I01 * u + I11 * v;
    const double av6[6] = {au.x, au.y, au.z, av.x, av.y, av.z};
    const double uu[3] = {u.x, u.y, u.z}, vv[3] = {v.x, v.y, v.z};
    const double ka = coef * (lambda - 4.0 * t1), kq = coef * 2.0 * t1 / detA;
    double H6[21];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            const double id = i == j ? 1.0 : 0.0;
            if (j >= i) {
                H6[SI<6>(i, j)] = 2.0 * S00 * id + ka * av6[i] * av6[j] - kq * vv[i] * vv[j];
                H6[SI<6>(3 + i, 3 + j)] = 2.0 * S11 * id + ka * av6[3 + i] * av6[3 + j] - kq * uu[i] * uu[j];
            }
            H6[SI<6>(i, 3 + j)] = 2.0 * S01 * id + ka * av6[i] * av6[3 + j] + kq * (2.0 * uu[i] * vv[j] - vv[i] * uu[j]);
        }
    // (u, v) = (x1 - x0, x2 - x0): coefficient rows c_u = (-1, 1, 0), c_v = (-1, 0, 1); T[m][k] = sum_i c[m][i] Hm[i][k]
    const double r2 = 1.4142135623730951, ir2 = 0.70710678118654752, r32 = 1.2247448713915890, ir6 = 0.40824829046386302;
    const double T[2][2] = {{-r2, 0.0}, {-ir2, -r32}};
    double M[21];
    congruence_blocks<2>(T, H6, M);
    if (projectSPD && !make_pd_ql<6>(M, V6)) out.eigFail = true;
    const double Hm[3][2] = {{ir2, ir6}, {-ir2, ir6}, {0.0, -2.0 * ir6}};
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            if (!emit.wants(i, j)) continue;
            double blk[9];
            expand_block<3, 2>(Hm, M, i, j, blk);
            emit(i, j, blk);
        }
    return true;
}

} // namespace idp
