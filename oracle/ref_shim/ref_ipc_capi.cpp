// ORACLE / TEST INFRASTRUCTURE: compiles the REFERENCE's own contact LOOPS -- FEM/IPC.h and Grid/SPATIAL_HASH.h, from
// where they lie under /root/reference/Library -- against the stand-ins in include/ (std::vector storage instead of
// Cabana, the repo's Eigen subset, an inert pybind11) and exposes the six operators
// <double, 3, shell=false, elasticIPC=false> to ctypes. No reference source is copied into this repository.
#include <FEM/IPC.h>
#include <FEM/FRICTION.h>

using namespace JGSL;
typedef double T;

namespace {
struct Scene {
    MESH_NODE<T, 3> X;
    MESH_NODE_ATTR<T, 3> nodeAttr;
    std::vector<int> bnode, particle;
    std::vector<VECTOR<int, 2>> bedge, rod;
    std::vector<VECTOR<int, 3>> btri;
    std::map<int, std::set<int>> nnExclusion;
    std::vector<T> BNArea, BEArea, BTArea;
    VECTOR<int, 2> codimBNStartInd;
    std::vector<bool> DBCb;
    Scene(int nV, const double* x, const double* x0, int nBN, const int* bn, int nBE, const int* be, int nBT, const int* bt, const unsigned char* dbc)
        : X(nV), nodeAttr(nV), codimBNStartInd(nBN, nBN)
    {
        for (int i = 0; i < nV; ++i) {
            X.Append(VECTOR<T, 3>(x[3 * i], x[3 * i + 1], x[3 * i + 2]));
            const double* r = x0 ? x0 + 3 * i : x + 3 * i;
            nodeAttr.Append(VECTOR<T, 3>(r[0], r[1], r[2]), VECTOR<T, 3>(0.0), VECTOR<T, 3>(0.0), 0.0);
        }
        bnode.assign(bn, bn + nBN);
        for (int i = 0; i < nBE; ++i) bedge.emplace_back(be[2 * i], be[2 * i + 1]);
        for (int i = 0; i < nBT; ++i) btri.emplace_back(bt[3 * i], bt[3 * i + 1], bt[3 * i + 2]);
        BNArea.assign(nBN, 1.0); BEArea.assign(nBE, 1.0); BTArea.assign(nBT, 1.0);
        DBCb.assign(nV, false);
        if (dbc) for (int i = 0; i < nV; ++i) DBCb[i] = dbc[i] != 0;
    }
};
void to_rows(const int* rows4, const double* w, int n, double dHat2, std::vector<VECTOR<int, 4>>& cs, std::vector<VECTOR<T, 2>>& info)
{
    for (int i = 0; i < n; ++i) {
        cs.emplace_back(rows4[4 * i], rows4[4 * i + 1], rows4[4 * i + 2], rows4[4 * i + 3]);
        info.emplace_back(w ? w[i] : 1.0, dHat2);
    }
}
} // namespace

extern "C" {

// returns the number of rows (<= cap are written); info2 = (weight, dHat2) per row
int refipc_constraint_set(int nV, const double* x, const double* x0, int nBN, const int* bn, int nBE, const int* be, int nBT, const int* bt,
    const unsigned char* dbc, double dHat2, double thickness, long cap, int* rows4, double* info2)
{
    Scene s(nV, x, x0, nBN, bn, nBE, be, nBT, bt, dbc);
    std::vector<VECTOR<int, 4>> cs;
    std::vector<VECTOR<int, 2>> ptee;
    std::vector<VECTOR<T, 2>> info;
    Compute_Constraint_Set<T, 3, false, false>(s.X, s.nodeAttr, s.bnode, s.bedge, s.btri, s.particle, s.rod, s.nnExclusion, s.BNArea, s.BEArea,
        s.BTArea, s.codimBNStartInd, s.DBCb, dHat2, thickness, false, cs, ptee, info);
    for (long i = 0; i < (long)cs.size() && i < cap; ++i) {
        for (int k = 0; k < 4; ++k) rows4[4 * i + k] = cs[i][k];
        info2[2 * i] = info[i][0]; info2[2 * i + 1] = info[i][1];
    }
    return (int)cs.size();
}

// E (added to *E), gradient (nV x 3, added), triplets (returns count; <= cap written)
long refipc_barrier(int nV, const double* x, const double* x0, int n, const int* rows4, const double* w, double dHat2, double kappa,
    double thickness, int projectSPD, double* E, double* g, long cap, int* trow, int* tcol, double* tval)
{
    Scene s(nV, x, x0, 0, nullptr, 0, nullptr, 0, nullptr, nullptr);
    std::vector<VECTOR<int, 4>> cs;
    std::vector<VECTOR<T, 2>> info;
    to_rows(rows4, w, n, dHat2, cs, info);
    T kap[3] = {kappa, kappa, kappa};
    if (E) Compute_Barrier<T, 3, false>(s.X, s.nodeAttr, cs, info, dHat2, kap, thickness, *E);
    if (g) {
        Compute_Barrier_Gradient<T, 3, false>(s.X, cs, info, dHat2, kap, thickness, s.nodeAttr);
        for (int i = 0; i < nV; ++i) {
            const VECTOR<T, 3>& gi = std::get<FIELDS<MESH_NODE_ATTR<T, 3>>::g>(s.nodeAttr.Get_Unchecked(i));
            for (int k = 0; k < 3; ++k) g[3 * i + k] += gi[k];
        }
    }
    long nt = 0;
    if (trow) {
        std::vector<Eigen::Triplet<T>> trip;
        Compute_Barrier_Hessian<T, 3, false>(s.X, s.nodeAttr, cs, info, dHat2, kap, thickness, projectSPD != 0, trip);
        nt = (long)trip.size();
        for (long i = 0; i < nt && i < cap; ++i) { trow[i] = trip[i].row(); tcol[i] = trip[i].col(); tval[i] = trip[i].value(); }
    }
    return nt;
}

double refipc_ccd(int nV, const double* x, int nBN, const int* bn, int nBE, const int* be, int nBT, const int* bt, const unsigned char* dbc,
    const double* dir, double thickness, double step)
{
    Scene s(nV, x, nullptr, nBN, bn, nBE, be, nBT, bt, dbc);
    std::vector<T> sd(dir, dir + 3 * (size_t)nV);
    T a = step;
    Compute_Intersection_Free_StepSize<T, 3, false, false>(s.X, s.bnode, s.bedge, s.btri, s.particle, s.rod, s.nnExclusion, s.codimBNStartInd,
        s.DBCb, sd, thickness, a);
    return a;
}

double refipc_min_dist2(int nV, const double* x, int n, const int* rows4, double thickness, double* dist2)
{
    Scene s(nV, x, nullptr, 0, nullptr, 0, nullptr, 0, nullptr, nullptr);
    std::vector<VECTOR<int, 4>> cs;
    std::vector<VECTOR<T, 2>> info;
    to_rows(rows4, nullptr, n, 0.0, cs, info);
    std::vector<T> d;
    T mn = 0;
    Compute_Min_Dist2<T, 3, false>(s.X, cs, thickness, d, mn);
    for (size_t i = 0; i < d.size(); ++i) dist2[i] = d[i];
    return mn;
}

// Lagged friction (FEM/FRICTION.h:17-662): Compute_Friction_Basis at xb with the contact rows, then potential / gradient /
// Hessian triplets at x relative to xn. Outputs: the friction rows (the non-mollified contact rows, in order), closest-point
// parameters (2 per row), tangent bases (6 per row, column major 3x2), normal forces; E (added), g (nV x 3, added), triplets.
long refipc_friction_comp(int nV, const double* xb, const double* x, const double* xn, int n, const int* rows4, const double* w, double dHat2, double kappa,
    double thickness, double epsvh2, double mu, int projectSPD, int* nFric, int* fricRows4, double* closest2, double* basis6, double* normalForce,
    double* E, double* g, long cap, int* trow, int* tcol, double* tval, int nComp, const int* compNodeRange, const double* muComp);

long refipc_friction(int nV, const double* xb, const double* x, const double* xn, int n, const int* rows4, const double* w, double dHat2, double kappa,
    double thickness, double epsvh2, double mu, int projectSPD, int* nFric, int* fricRows4, double* closest2, double* basis6, double* normalForce,
    double* E, double* g, long cap, int* trow, int* tcol, double* tval)
{
    return refipc_friction_comp(nV, xb, x, xn, n, rows4, w, dHat2, kappa, thickness, epsvh2, mu, projectSPD, nFric, fricRows4, closest2, basis6, normalForce, E, g,
        cap, trow, tcol, tval, 0, nullptr, nullptr);
}

// the same with per-component coefficients: Compute_Friction_Coef (FRICTION.h:126-170) scales the normal forces right after the basis,
// as the time step does (Shell/IMPLICIT_EULER.h:435-438); it sets mu to 1
long refipc_friction_comp(int nV, const double* xb, const double* x, const double* xn, int n, const int* rows4, const double* w, double dHat2, double kappa,
    double thickness, double epsvh2, double mu, int projectSPD, int* nFric, int* fricRows4, double* closest2, double* basis6, double* normalForce,
    double* E, double* g, long cap, int* trow, int* tcol, double* tval, int nComp, const int* compNodeRange, const double* muComp)
{
    Scene sb(nV, xb, nullptr, 0, nullptr, 0, nullptr, 0, nullptr, nullptr);
    std::vector<VECTOR<int, 4>> cs, fcs;
    std::vector<VECTOR<T, 2>> info;
    to_rows(rows4, w, n, dHat2, cs, info);
    std::vector<Eigen::Matrix<T, 2, 1>> cp;
    std::vector<Eigen::Matrix<T, 3, 2>> tb;
    std::vector<T> nf;
    T kap[3] = {kappa, kappa, kappa};
    Compute_Friction_Basis<T, 3, false>(sb.X, cs, info, fcs, cp, tb, nf, dHat2, kap, thickness);
    if (nComp > 0) {
        std::vector<int> range(compNodeRange, compNodeRange + nComp);
        std::vector<T> mc(muComp, muComp + (size_t)nComp * nComp);
        Compute_Friction_Coef<T, 3>(fcs, range, mc, nf, mu);
    }
    *nFric = (int)fcs.size();
    for (size_t i = 0; i < fcs.size(); ++i) {
        for (int k = 0; k < 4; ++k) fricRows4[4 * i + k] = fcs[i][k];
        closest2[2 * i] = cp[i][0]; closest2[2 * i + 1] = cp[i][1];
        for (int c = 0; c < 2; ++c) for (int r = 0; r < 3; ++r) basis6[6 * i + 3 * c + r] = tb[i](r, c);
        normalForce[i] = nf[i];
    }
    if (!x) return 0;
    Scene s(nV, x, nullptr, 0, nullptr, 0, nullptr, 0, nullptr, nullptr), sn(nV, xn, nullptr, 0, nullptr, 0, nullptr, 0, nullptr, nullptr);
    if (E) Compute_Friction_Potential<T, 3>(s.X, sn.X, fcs, cp, tb, nf, epsvh2, mu, *E);
    if (g) {
        Compute_Friction_Gradient<T, 3>(s.X, sn.X, fcs, cp, tb, nf, epsvh2, mu, s.nodeAttr);
        for (int i = 0; i < nV; ++i) {
            const VECTOR<T, 3>& gi = std::get<FIELDS<MESH_NODE_ATTR<T, 3>>::g>(s.nodeAttr.Get_Unchecked(i));
            for (int k = 0; k < 3; ++k) g[3 * i + k] += gi[k];
        }
    }
    long nt = 0;
    if (trow) {
        std::vector<Eigen::Triplet<T>> trip;
        Compute_Friction_Hessian<T, 3>(s.X, sn.X, fcs, cp, tb, nf, epsvh2, mu, projectSPD != 0, trip);
        nt = (long)trip.size();
        for (long i = 0; i < nt && i < cap; ++i) { trow[i] = trip[i].row(); tcol[i] = trip[i].col(); tval[i] = trip[i].value(); }
    }
    return nt;
}

// Compute_Friction_Coef (FRICTION.h:126-170) on friction rows and their normal forces (scaled in place); returns the mu it sets (1)
double refipc_friction_coef(int n, const int* fricRows4, int nComp, const int* compNodeRange, const double* muComp, double* normalForce)
{
    std::vector<VECTOR<int, 4>> fcs;
    for (int i = 0; i < n; ++i) fcs.emplace_back(fricRows4[4 * i], fricRows4[4 * i + 1], fricRows4[4 * i + 2], fricRows4[4 * i + 3]);
    std::vector<int> range(compNodeRange, compNodeRange + nComp);
    std::vector<T> mc(muComp, muComp + (size_t)nComp * nComp), nf(normalForce, normalForce + n);
    T mu = 0;
    Compute_Friction_Coef<T, 3>(fcs, range, mc, nf, mu);
    for (int i = 0; i < n; ++i) normalForce[i] = nf[i];
    return mu;
}

} // extern "C"
