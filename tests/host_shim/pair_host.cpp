// Test-only harness: compiles the product's per-pair device headers (idp_b200/csrc/pair_*.cuh) as plain host
// C++ so that their arithmetic can be compared with the oracle on a machine without a GPU. This is NOT a CPU
// fallback of the product: nothing in idp_b200/ links or loads it.
#include "../../idp_b200/csrc/pair_deriv.cuh"
#include "../../idp_b200/csrc/psd_lowrank.cuh"
#include "../../idp_b200/csrc/shell_elastic.cuh"
#include "../../idp_b200/csrc/friction.cuh"
using namespace idp;
static V3 l3(const double* p) { return mk3(p[0], p[1], p[2]); }
extern "C" {
int hs_pt_type(const double* x) { return pt_type(l3(x), l3(x + 3), l3(x + 6), l3(x + 9)); }
int hs_ee_type(const double* x) { return ee_type(l3(x), l3(x + 3), l3(x + 6), l3(x + 9)); }
double hs_dist2_unclassified(int kind, const double* x)
{
    return kind == 0 ? dist2_pt_unclassified(l3(x), l3(x + 3), l3(x + 6), l3(x + 9))
                     : dist2_ee_unclassified(l3(x), l3(x + 3), l3(x + 6), l3(x + 9));
}
// row: 4 ints; X, X0: nV x 3; returns 0 ok / 1 nonpositive distance. H is n x n, n = 3 nv (written to nv_out)
int hs_row_EgH(const int* row, const double* X, const double* X0, double weight, double dHat2, double kappa,
    double xi, int projectSPD, int lowrank, double* E, double* g, double* H, int* nv_out, int* verts)
{
    RowDec d = decode_row(row[0], row[1], row[2], row[3]);
    V3 x[4], xr[4];
    for (int i = 0; i < 4; ++i) { x[i] = l3(X + 3 * (long)d.v[i]); xr[i] = l3(X0 + 3 * (long)d.v[i]); verts[i] = d.v[i]; }
    *nv_out = d.nv;
    const double dh2 = dHat2 + 2 * sqrt(dHat2) * xi;
    bool ok;
    if (lowrank) ok = row_EgH_lowrank(d, x, xr, weight, dh2, kappa, xi * xi, projectSPD != 0, E, g, H);
    else ok = row_EgH(d, x, xr, weight, dh2, kappa, xi * xi, projectSPD != 0, E, g, H);
    return ok ? 0 : 1;
}
int hs_accd(int kind, const double* x, const double* d, double eta, double xi, double bound, double* toc, int* iters)
{
    int it = 0, r;
    auto bf = [bound]() { return bound; };
    if (kind == 0) r = accd_pt(l3(x), l3(x + 3), l3(x + 6), l3(x + 9), l3(d), l3(d + 3), l3(d + 6), l3(d + 9), eta, xi, bf, *toc, it);
    else r = accd_ee(l3(x), l3(x + 3), l3(x + 6), l3(x + 9), l3(d), l3(d + 3), l3(d + 6), l3(d + 9), eta, xi, bf, *toc, it);
    *iters = it;
    return r;
}
// make_pd_ql on a dense symmetric N x N matrix (row-major in / out); N = 6 or 9. Returns 0 when the QL iteration converged.
int hs_make_pd(int n, const double* A, double* out)
{
    double work[QlStore<9, 1>::WORDS];
    bool ok = false;
    if (n == 9) {
        double m[45];
        for (int i = 0; i < 9; ++i) for (int j = i; j < 9; ++j) m[SI<9>(i, j)] = A[i * 9 + j];
        QlStore<9, 1> S{work};
        ok = make_pd_ql<9>(m, S);
        for (int i = 0; i < 9; ++i) for (int j = 0; j < 9; ++j) out[i * 9 + j] = m[SI<9>(i, j)];
    }
    else if (n == 6) {
        double m[21];
        for (int i = 0; i < 6; ++i) for (int j = i; j < 6; ++j) m[SI<6>(i, j)] = A[i * 6 + j];
        QlStore<6, 1> S{work};
        ok = make_pd_ql<6>(m, S);
        for (int i = 0; i < 6; ++i) for (int j = 0; j < 6; ++j) out[i * 6 + j] = m[SI<6>(i, j)];
    }
    else return -1;
    return ok ? 0 : 1;
}
// shell_elastic.cuh on the host: one hinge (x: 4 x 3) / one membrane triangle (x: 3 x 3); dense H (12x12 / 9x9, row major)
int hs_hinge_EgH(const double* x, double thetabar, double coef, int projectSPD, double* E, double* g, double* H)
{
    double work[QlStore<9, 1>::WORDS];
    QlStore<9, 1> V9{work};
    const V3 xv[4] = {l3(x), l3(x + 3), l3(x + 6), l3(x + 9)};
    ElasticOut out;
    DenseEmit em{H, 12};
    hinge_eval(xv, thetabar, coef, projectSPD != 0, true, true, V9, out, em);
    *E = out.E;
    for (int i = 0; i < 12; ++i) g[i] = out.g[i];
    return out.eigFail ? 1 : 0;
}
int hs_membrane_EgH(const double* x, const double* ib3, double coef, double lambda, double mu, int projectSPD, double* E, double* g, double* H)
{
    double work[QlStore<6, 1>::WORDS];
    QlStore<6, 1> V6{work};
    const V3 xv[3] = {l3(x), l3(x + 3), l3(x + 6)};
    ElasticOut out;
    DenseEmit em{H, 9};
    if (!membrane_eval(xv, ib3, coef, lambda, mu, projectSPD != 0, true, true, V6, out, em)) return 2;
    *E = out.E;
    for (int i = 0; i < 9; ++i) g[i] = out.g[i];
    return out.eigFail ? 1 : 0;
}
// friction.cuh on the host: basis of n contact rows at Xb, then E (added), g (nV x 3, added) and the dense per-row Hessians
// (n x 12 x 12 over the row's stencil order, zero padded) at X relative to Xn. Outputs per row: nv, v[4], w[4], basis (6), lam.
void hs_friction(int n, const int* rows4, const double* Xb, const double* X, const double* Xn, double dHat2, double kappa, double xi, double epsvh,
    double mu, int* nv, int* verts, double* w, double* basis, double* lam, double* cp, double* E, double* g, double* H)
{
    for (int i = 0; i < n; ++i) {
        const RowDec d = decode_row(rows4[4 * i], rows4[4 * i + 1], rows4[4 * i + 2], rows4[4 * i + 3]);
        V3 xb[4];
        for (int k = 0; k < 4; ++k) xb[k] = l3(Xb + 3 * (long)d.v[k]);
        FricRow f;
        friction_basis(d, xb, 1.0, dHat2 + 2 * sqrt(dHat2) * xi, kappa, xi * xi, f);
        nv[i] = f.nv; lam[i] = f.lam;
        for (int k = 0; k < 4; ++k) { verts[4 * i + k] = f.v[k]; w[4 * i + k] = f.w[k]; }
        for (int a = 0; a < 3; ++a) { basis[6 * i + a] = f.t0[a]; basis[6 * i + 3 + a] = f.t1[a]; }
        cp[2 * i] = f.cp[0]; cp[2 * i + 1] = f.cp[1];
        if (!X || f.nv == 0) continue;
        V3 dx[4];
        for (int k = 0; k < 4; ++k) {
            dx[k] = mk3(0, 0, 0);
            if (k < f.nv) dx[k] = (f.pp_abs && k == 1) ? l3(X + 3 * (long)f.v[k]) : l3(X + 3 * (long)f.v[k]) - l3(Xn + 3 * (long)f.v[k]);
        }
        *E += friction_energy(f, dx, epsvh, mu);
        double g3[3], B[9];
        friction_gradient(f, dx, epsvh, mu, g3);
        friction_hessian_core(f, dx, epsvh, mu, B);
        for (int k = 0; k < f.nv; ++k) for (int a = 0; a < 3; ++a) g[3 * (long)f.v[k] + a] += f.w[k] * g3[a];
        for (int p = 0; p < f.nv; ++p)
            for (int q = 0; q < f.nv; ++q)
                for (int a = 0; a < 3; ++a)
                    for (int b = 0; b < 3; ++b) H[144 * (long)i + (3 * p + a) * 12 + 3 * q + b] = f.w[p] * f.w[q] * B[3 * a + b];
    }
}
#ifdef IDP_QL_STATS
// development aid (scripts/ql_stats.py): chase lengths of the QL trips of the last 9x9 projection
int hs_ql_negcount() { return g_ql_stats.negcount; }
int hs_ql_trace(int* out) { for (int i = 0; i < g_ql_stats.n; ++i) out[i] = g_ql_stats.chase[i]; return g_ql_stats.n; }
#endif
}
