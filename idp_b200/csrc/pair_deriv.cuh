// Gradients / Hessians of the squared distances and of the edge-edge mollifier in reduced (difference)
// coordinates, embedding into vertex DOFs, and the dense symmetric PSD projection (makePD semantics).
// Only a 1e-10 relative tolerance is required of these quantities (BASELINE.json), so translation units
// using this header may be compiled with FMA contraction enabled.
//
// Reference functions replaced (relative to /root/reference/Library):
//   Math/Distance/POINT_POINT.h:19-41, POINT_EDGE.h:61-113,267-587, POINT_TRIANGLE.h:25-86,104-549,
//   EDGE_EDGE.h:25-104,122-732, EDGE_EDGE_MOLLIFIER.h:21-79,98-366,441-524, Math/BARRIER.h:10-62,
//   Math/UTILS.h:9-27 (makePD).
// The formulas are derived by hand from the distance definitions (they are not transcriptions of the
// reference's generated code): with y = (w,u,v), A = w.(u x v), B = |u x v|^2, r = A/B, q = grad A - r grad B
//   d = A^2/B,  grad d = 2 r grad A - r^2 grad B,  hess d = (2/B) q q^T + 2 r hess A - r^2 hess B.
#pragma once
#include "pair_exact.cuh"

namespace idp {

IDP_HD double comp3(const V3& a, int i) { return i == 0 ? a.x : (i == 1 ? a.y : a.z); }
// skew(a)[i][j] with skew(a) b = a x b
IDP_HD double skew3(const V3& a, int i, int j)
{
    if (i == j) return 0.0;
    const int k = 3 - i - j;
    const double s = ((j - i + 3) % 3 == 1) ? -1.0 : 1.0; // (0,1),(1,2),(2,0) -> -a_k ; reversed -> +a_k
    return s * comp3(a, k);
}

// ---- barrier scalars on squared distance (BARRIER.h, elastic=false) ----------------------------------
IDP_HD void barrier_all(double d, double dHat2, double kappa, double& b, double& bg, double& bh)
{
    const double t2 = d - dHat2;
    const double lg = log(d / dHat2);
    b = -kappa * t2 * t2 * lg;
    bg = kappa * (t2 * lg * -2.0 - (t2 * t2) / d);
    bh = kappa * ((lg * -2.0 - t2 * 4.0 / d) + 1.0 / (d * d) * (t2 * t2));
}

// ---- d = A^2/B, reduced coordinates (w,u,v): gradient g9 and Hessian H9 (row-major 9x9) ----------------
IDP_HD void triple_quotient(const V3& w, const V3& u, const V3& v, double* g9, double* H9)
{
    const V3 n = cross3(u, v);
    const double A = dot3(w, n), B = sqn3(n), r = A / B;
    const V3 Au = cross3(v, w), Av = cross3(w, u);
    const V3 Bu = 2.0 * cross3(v, n), Bv = 2.0 * cross3(n, u);
    const double gA[9] = {n.x, n.y, n.z, Au.x, Au.y, Au.z, Av.x, Av.y, Av.z};
    const double gB[9] = {0, 0, 0, Bu.x, Bu.y, Bu.z, Bv.x, Bv.y, Bv.z};
    double q[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) {
        g9[i] = 2.0 * r * gA[i] - r * r * gB[i];
        q[i] = gA[i] - r * gB[i];
    }
    const double s2 = 2.0 / B, r2 = r * r, tr = 2.0 * r;
#pragma unroll
    for (int i = 0; i < 9; ++i)
#pragma unroll
        for (int j = 0; j < 9; ++j) H9[i * 9 + j] = s2 * q[i] * q[j];
    const double uu = sqn3(u), vv = sqn3(v), uv = dot3(u, v);
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            const double dij = (i == j) ? 1.0 : 0.0;
            // hess A blocks: [w,u] = -skew(v), [w,v] = skew(u), [u,v] = -skew(w)
            const double awu = -skew3(v, i, j), awv = skew3(u, i, j), auv = -skew3(w, i, j);
            H9[i * 9 + 3 + j] += tr * awu; H9[(3 + j) * 9 + i] += tr * awu;
            H9[i * 9 + 6 + j] += tr * awv; H9[(6 + j) * 9 + i] += tr * awv;
            // hess B blocks
            const double buu = 2.0 * (vv * dij - comp3(v, i) * comp3(v, j));
            const double bvv = 2.0 * (uu * dij - comp3(u, i) * comp3(u, j));
            const double buv = 2.0 * (comp3(u, i) * comp3(v, j) - uv * dij) - 2.0 * skew3(n, i, j);
            H9[(3 + i) * 9 + 3 + j] -= r2 * buu;
            H9[(6 + i) * 9 + 6 + j] -= r2 * bvv;
            H9[(3 + i) * 9 + 6 + j] += tr * auv - r2 * buv;
            H9[(6 + j) * 9 + 3 + i] += tr * auv - r2 * buv;
        }
}

// ---- N = |w x u|^2 in (w,u): value, gradient g6, Hessian H6 (row-major 6x6) -------------------------------
IDP_HD void cross_norm2_reduced(const V3& w, const V3& u, double& N, double* g6, double* H6)
{
    const V3 c = cross3(w, u);
    N = sqn3(c);
    const V3 gw = 2.0 * cross3(u, c), gu = 2.0 * cross3(c, w);
    g6[0] = gw.x; g6[1] = gw.y; g6[2] = gw.z; g6[3] = gu.x; g6[4] = gu.y; g6[5] = gu.z;
    const double uu = sqn3(u), ww = sqn3(w), wu = dot3(w, u);
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            const double dij = (i == j) ? 1.0 : 0.0;
            H6[i * 6 + j] = 2.0 * (uu * dij - comp3(u, i) * comp3(u, j));
            H6[(3 + i) * 6 + 3 + j] = 2.0 * (ww * dij - comp3(w, i) * comp3(w, j));
            const double b = 2.0 * (comp3(w, i) * comp3(u, j) - wu * dij) - 2.0 * skew3(c, i, j);
            H6[i * 6 + 3 + j] = b;
            H6[(3 + j) * 6 + i] = b;
        }
}

// d = |w x u|^2 / |u|^2 in (w,u)
IDP_HD void pe_reduced(const V3& w, const V3& u, double* g6, double* H6)
{
    double N, gN[6], HN[36];
    cross_norm2_reduced(w, u, N, gN, HN);
    const double B = sqn3(u), iB = 1.0 / B, NB2 = N * iB * iB;
    const double gB[6] = {0, 0, 0, 2.0 * u.x, 2.0 * u.y, 2.0 * u.z};
#pragma unroll
    for (int i = 0; i < 6; ++i) g6[i] = gN[i] * iB - NB2 * gB[i];
#pragma unroll
    for (int i = 0; i < 6; ++i)
#pragma unroll
        for (int j = 0; j < 6; ++j) {
            const double hb = (i >= 3 && i == j) ? 2.0 : 0.0;
            H6[i * 6 + j] = HN[i * 6 + j] * iB - (gN[i] * gB[j] + gB[i] * gN[j]) * (iB * iB) - NB2 * hb
                + (2.0 * NB2 * iB) * gB[i] * gB[j];
        }
}

// embed K reduced 3-vectors into NV vertices: y_k = sum_i c[k][i] x_i  (c entries in {-1,0,1})
template <int K, int NV>
IDP_HD void embed_gH(const signed char (&c)[K][NV], const double* gk, const double* Hk, double* g, double* H)
{
    const int n = 3 * NV, m = 3 * K;
#pragma unroll
    for (int i = 0; i < NV; ++i)
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            double s = 0;
#pragma unroll
            for (int k = 0; k < K; ++k)
                if (c[k][i]) s += c[k][i] * gk[3 * k + a];
            g[3 * i + a] = s;
        }
    if (!H) return;
#pragma unroll
    for (int i = 0; i < NV; ++i)
#pragma unroll
        for (int j = 0; j < NV; ++j)
#pragma unroll
            for (int a = 0; a < 3; ++a)
#pragma unroll
                for (int b = 0; b < 3; ++b) {
                    double s = 0;
#pragma unroll
                    for (int k = 0; k < K; ++k)
#pragma unroll
                        for (int l = 0; l < K; ++l)
                            if (c[k][i] * c[l][j] != 0) s += (c[k][i] * c[l][j]) * Hk[(3 * k + a) * m + (3 * l + b)];
                    H[(3 * i + a) * n + (3 * j + b)] = s;
                }
}

// distance gradient (+ Hessian if H != nullptr) for the four geometric kinds, stencil orders of SURVEY.md A.1
IDP_HD void pp_gH(const V3& a, const V3& b, double* g, double* H)
{
    const V3 d = 2.0 * (a - b);
    g[0] = d.x; g[1] = d.y; g[2] = d.z; g[3] = -d.x; g[4] = -d.y; g[5] = -d.z;
    if (!H) return;
#pragma unroll
    for (int i = 0; i < 6; ++i)
#pragma unroll
        for (int j = 0; j < 6; ++j) H[i * 6 + j] = (i == j) ? 2.0 : ((i % 3 == j % 3) ? -2.0 : 0.0);
}
IDP_HD void pe_gH(const V3& p, const V3& e0, const V3& e1, double* g, double* H)
{
    double g6[6], H6[36];
    pe_reduced(p - e0, e1 - e0, g6, H6);
    const signed char c[2][3] = {{1, -1, 0}, {0, -1, 1}};
    embed_gH<2, 3>(c, g6, H6, g, H);
}
IDP_HD void pt_gH(const V3& p, const V3& t0, const V3& t1, const V3& t2, double* g, double* H)
{
    double g9[9], H9[81];
    triple_quotient(p - t0, t1 - t0, t2 - t0, g9, H9);
    const signed char c[3][4] = {{1, -1, 0, 0}, {0, -1, 1, 0}, {0, -1, 0, 1}};
    embed_gH<3, 4>(c, g9, H9, g, H);
}
IDP_HD void ee_gH(const V3& a0, const V3& a1, const V3& b0, const V3& b1, double* g, double* H)
{
    double g9[9], H9[81];
    triple_quotient(b0 - a0, a1 - a0, b1 - b0, g9, H9);
    const signed char c[3][4] = {{-1, 0, 1, 0}, {-1, 1, 0, 0}, {0, 0, -1, 1}};
    embed_gH<3, 4>(c, g9, H9, g, H);
}
// mollifier e(c), grad, hess over (ea0, ea1, eb0, eb1); c = |u x v|^2  (EDGE_EDGE_MOLLIFIER.h:441-524)
IDP_HD void mollifier_all(const V3& a0, const V3& a1, const V3& b0, const V3& b1, double eps_x, double& e, double* ge, double* He)
{
    double c, g6[6], H6[36];
    cross_norm2_reduced(a1 - a0, b1 - b0, c, g6, H6);
    if (c < eps_x) {
        const double q = c / eps_x, ie = 1.0 / eps_x;
        e = (-q + 2.0) * q;
        const double qg = 2.0 * ie * (-ie * c + 1.0), qh = -2.0 / (eps_x * eps_x);
        const signed char cm[2][4] = {{-1, 1, 0, 0}, {0, 0, -1, 1}};
        double gc[12];
        embed_gH<2, 4>(cm, g6, H6, gc, He);
        if (He) {
            for (int i = 0; i < 12; ++i)
                for (int j = 0; j < 12; ++j) He[i * 12 + j] = He[i * 12 + j] * qg + (qh * gc[i]) * gc[j];
        }
        for (int i = 0; i < 12; ++i) ge[i] = gc[i] * qg;
    }
    else {
        e = 1.0;
        for (int i = 0; i < 12; ++i) ge[i] = 0;
        if (He) for (int i = 0; i < 144; ++i) He[i] = 0;
    }
}

// ---- dense symmetric PSD projection (makePD): cyclic Jacobi with eigenvector accumulation ---------------
// Semantics of Math/UTILS.h:9-27: if lambda_min >= 0 the matrix is returned untouched; otherwise negative
// eigenvalues are zeroed and H = V diag(lambda) V^T. A and V are n x n row-major work arrays.
template <int N>
IDP_HD void make_pd_jacobi(double* H)
{
    double A[N * N], V[N * N];
    for (int i = 0; i < N; ++i)
        for (int j = 0; j < N; ++j) {
            A[i * N + j] = (i >= j) ? H[i * N + j] : H[j * N + i]; // lower triangle is read
            V[i * N + j] = (i == j) ? 1.0 : 0.0;
        }
    for (int sweep = 0; sweep < 30; ++sweep) {
        double off = 0, tot = 0;
        for (int i = 0; i < N; ++i)
            for (int j = 0; j < N; ++j) {
                const double a = A[i * N + j];
                tot += a * a;
                if (i != j) off += a * a;
            }
        if (off <= 1e-34 * tot || off == 0) break;
        for (int p = 0; p < N - 1; ++p)
            for (int q = p + 1; q < N; ++q) {
                const double apq = A[p * N + q];
                if (apq == 0) continue;
                const double theta = (A[q * N + q] - A[p * N + p]) / (2.0 * apq);
                const double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
                const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
                for (int k = 0; k < N; ++k) {
                    const double akp = A[k * N + p], akq = A[k * N + q];
                    A[k * N + p] = c * akp - s * akq;
                    A[k * N + q] = s * akp + c * akq;
                }
                for (int k = 0; k < N; ++k) {
                    const double apk = A[p * N + k], aqk = A[q * N + k];
                    A[p * N + k] = c * apk - s * aqk;
                    A[q * N + k] = s * apk + c * aqk;
                }
                for (int k = 0; k < N; ++k) {
                    const double vkp = V[k * N + p], vkq = V[k * N + q];
                    V[k * N + p] = c * vkp - s * vkq;
                    V[k * N + q] = s * vkp + c * vkq;
                }
            }
    }
    double lmin = A[0];
    for (int i = 1; i < N; ++i) lmin = fmin(lmin, A[i * N + i]);
    if (lmin >= 0) return;
    // H = sum_k max(lambda_k, 0) v_k v_k^T
    for (int i = 0; i < N; ++i)
        for (int j = 0; j <= i; ++j) {
            double s = 0;
            for (int k = 0; k < N; ++k) {
                const double l = A[k * N + k];
                if (l > 0) s += V[i * N + k] * l * V[j * N + k];
            }
            H[i * N + j] = s;
            H[j * N + i] = s;
        }
}

// ---- one constraint row: local energy, gradient (12) and Hessian (n x n, n = 3 nv) -------------------------
// x[4]: current positions of the decoded stencil, x0[4]: rest positions (mollified kinds only).
// Implements IPC.h:801-938 (E), 1012-1254 (g), 1390-1729 (H). Returns false when dist2 - xi^2 <= 0.
IDP_HD bool row_EgH(const RowDec& d, const V3* x, const V3* xr, double weight, double dHat2, double kappa, double xi2,
    bool projectSPD, double* E, double* g, double* H)
{
    const double dist2 = row_dist2(d.kind, x[0], x[1], x[2], x[3]) - xi2;
    if (!(dist2 > 0)) return false;
    double b, bg, bh;
    barrier_all(dist2, dHat2, kappa, b, bg, bh);
    const double mu = (double)d.mult;
    const bool moll = (d.kind == K_EE_M || d.kind == K_PE_M || d.kind == K_PP_M);
    if (!moll) {
        if (E) *E = b * (d.mult > 1 ? mu : 1.0) * weight;
        if (!g && !H) return true;
        double dg[12];
        const int n = 3 * d.nv;
        switch (d.kind) {
        case K_EE: ee_gH(x[0], x[1], x[2], x[3], dg, H); break;
        case K_PT: pt_gH(x[0], x[1], x[2], x[3], dg, H); break;
        case K_PE: pe_gH(x[0], x[1], x[2], dg, H); break;
        default: pp_gH(x[0], x[1], dg, H); break;
        }
        if (g) for (int i = 0; i < n; ++i) g[i] = dg[i] * (mu * weight * bg);
        if (H) {
            for (int i = 0; i < n; ++i)
                for (int j = 0; j < n; ++j) H[i * n + j] = (((mu * bh) * dg[i]) * dg[j] + (mu * bg) * H[i * n + j]) * weight;
            if (projectSPD) {
                if (d.nv == 4) make_pd_jacobi<12>(H);
                else if (d.nv == 3) make_pd_jacobi<9>(H);
                else make_pd_jacobi<6>(H);
            }
        }
        return true;
    }
    const double eps_x = ee_mollifier_threshold(xr[0], xr[1], xr[2], xr[3]);
    double e, ge[12];
    double He[144];
    mollifier_all(x[0], x[1], x[2], x[3], eps_x, e, ge, H ? He : nullptr);
    if (E) *E = b * e * weight;
    if (!g && !H) return true;
    double dg[12], dH[144];
    int nd;
    int map[12];
    if (d.kind == K_EE_M) { nd = 12; ee_gH(x[0], x[1], x[2], x[3], dg, H ? dH : nullptr); for (int i = 0; i < 12; ++i) map[i] = i; }
    else if (d.kind == K_PE_M) {
        nd = 9; pe_gH(x[0], x[2], x[3], dg, H ? dH : nullptr);
        for (int i = 0; i < 3; ++i) { map[i] = i; map[3 + i] = 6 + i; map[6 + i] = 9 + i; }
    }
    else {
        nd = 6; pp_gH(x[0], x[2], dg, H ? dH : nullptr);
        for (int i = 0; i < 3; ++i) { map[i] = i; map[3 + i] = 6 + i; }
    }
    double Pg[12];
    for (int i = 0; i < 12; ++i) Pg[i] = 0;
    for (int i = 0; i < nd; ++i) Pg[map[i]] = dg[i];
    if (g) for (int i = 0; i < 12; ++i) g[i] = weight * ((e * bg) * Pg[i] + b * ge[i]);
    if (H) {
        for (int i = 0; i < 144; ++i) H[i] = b * He[i];
        for (int i = 0; i < nd; ++i)
            for (int j = 0; j < nd; ++j) H[map[i] * 12 + map[j]] += ((e * bh) * dg[i]) * dg[j] + (e * bg) * dH[i * nd + j];
        for (int i = 0; i < 12; ++i)
            for (int j = 0; j < 12; ++j) H[i * 12 + j] += (bg * Pg[i]) * ge[j] + (bg * Pg[j]) * ge[i];
        for (int i = 0; i < 144; ++i) H[i] *= weight;
        if (projectSPD) make_pd_jacobi<12>(H);
    }
    return true;
}

} // namespace idp
