// Test-only: the drop-in boundary exercised with the REFERENCE's own types and operators in the same translation unit.
// <FEM/IPC.h> is the reference's header (compiled from /root/reference against the stand-ins of oracle/ref_shim/include);
// "IPC_B200.h" is this repository's host mirror. Both sets of six operators are called with the SAME argument objects
// (MESH_NODE / MESH_NODE_ATTR storages, std::vector<VECTOR<int,4>>, std::vector<Eigen::Triplet<double>> ...), exactly as a
// maintainer would switch the five call sites, and the results are compared. Built only where /root/reference exists
// (tests/host_shim/Makefile); the shared object travels to the GPU box.
#include <FEM/IPC.h>
#include <Math/CSR_MATRIX.h>
#include "../../idp_b200/host/IPC_B200.h"
#include <algorithm>
#include <cmath>

using namespace JGSL;
typedef double T;

namespace {
typedef std::vector<Eigen::Triplet<T>> Trips;
// sum duplicates: (row, col) -> value, sorted
std::vector<std::pair<std::pair<int, int>, T>> merged(const Trips& t)
{
    std::vector<std::pair<std::pair<int, int>, T>> v;
    v.reserve(t.size());
    for (const auto& x : t) v.push_back({{x.row(), x.col()}, x.value()});
    std::sort(v.begin(), v.end(), [](const auto& a, const auto& b) { return a.first < b.first; });
    size_t o = 0;
    for (size_t i = 0; i < v.size(); ++i) {
        if (o && v[o - 1].first == v[i].first) v[o - 1].second += v[i].second;
        else v[o++] = v[i];
    }
    v.resize(o);
    return v;
}
} // namespace

extern "C" int b2_side_by_side(int nV, const double* x, const double* x0, int nBN, const int* bn, int nBE, const int* be, int nBT,
    const int* bt, const double* dir, double dHat2, double kappa, double thickness, double* report)
{
    MESH_NODE<T, 3> X(nV);
    MESH_NODE_ATTR<T, 3> attrRef(nV), attrNew(nV);
    for (int i = 0; i < nV; ++i) {
        X.Append(VECTOR<T, 3>(x[3 * i], x[3 * i + 1], x[3 * i + 2]));
        const VECTOR<T, 3> r(x0[3 * i], x0[3 * i + 1], x0[3 * i + 2]);
        attrRef.Append(r, VECTOR<T, 3>(0.0), VECTOR<T, 3>(0.0), 0.0);
        attrNew.Append(r, VECTOR<T, 3>(0.0), VECTOR<T, 3>(0.0), 0.0);
    }
    std::vector<int> bnode(bn, bn + nBN), particle;
    std::vector<VECTOR<int, 2>> bedge, rod;
    std::vector<VECTOR<int, 3>> btri;
    for (int i = 0; i < nBE; ++i) bedge.emplace_back(be[2 * i], be[2 * i + 1]);
    for (int i = 0; i < nBT; ++i) btri.emplace_back(bt[3 * i], bt[3 * i + 1], bt[3 * i + 2]);
    std::map<int, std::set<int>> nn;
    std::vector<T> BNArea(nBN, 1.0), BEArea(nBE, 1.0), BTArea(nBT, 1.0);
    const VECTOR<int, 2> codim(nBN, nBN);
    std::vector<bool> DBCb(nV, false);
    T kap[3] = {kappa, kappa, kappa};

    // ---- constraint set
    std::vector<VECTOR<int, 4>> csRef, csNew;
    std::vector<VECTOR<int, 2>> ptee;
    std::vector<VECTOR<T, 2>> infoRef, infoNew;
    Compute_Constraint_Set<T, 3, false, false>(X, attrRef, bnode, bedge, btri, particle, rod, nn, BNArea, BEArea, BTArea, codim, DBCb, dHat2,
        thickness, false, csRef, ptee, infoRef);
    B200::Compute_Constraint_Set<T, 3, false, false>(X, attrNew, bnode, bedge, btri, particle, rod, nn, BNArea, BEArea, BTArea, codim, DBCb,
        dHat2, thickness, false, csNew, ptee, infoNew);
    auto key = [](const VECTOR<int, 4>& a, const VECTOR<int, 4>& b) {
        for (int k = 0; k < 4; ++k) if (a[k] != b[k]) return a[k] < b[k];
        return false;
    };
    std::vector<VECTOR<int, 4>> a = csRef, b = csNew;
    std::sort(a.begin(), a.end(), key); std::sort(b.begin(), b.end(), key);
    bool same = a.size() == b.size();
    for (size_t i = 0; same && i < a.size(); ++i) for (int k = 0; k < 4; ++k) same = same && a[i][k] == b[i][k];
    report[0] = (double)csRef.size(); report[1] = (double)csNew.size(); report[2] = same ? 1.0 : 0.0;

    // ---- E, g, H on the reference's rows (both sides get the same constraintSet / stencilInfo objects)
    T eRef = 0.25, eNew = 0.25;
    Compute_Barrier<T, 3, false>(X, attrRef, csRef, infoRef, dHat2, kap, thickness, eRef);
    B200::Compute_Barrier<T, 3, false>(X, attrNew, csRef, infoRef, dHat2, kap, thickness, eNew);
    report[3] = eRef; report[4] = eNew;
    Compute_Barrier_Gradient<T, 3, false>(X, csRef, infoRef, dHat2, kap, thickness, attrRef);
    B200::Compute_Barrier_Gradient<T, 3, false>(X, csRef, infoRef, dHat2, kap, thickness, attrNew);
    T gmax = 0, gdiff = 0;
    for (int i = 0; i < nV; ++i) {
        const VECTOR<T, 3>& g0 = std::get<FIELDS<MESH_NODE_ATTR<T, 3>>::g>(attrRef.Get_Unchecked(i));
        const VECTOR<T, 3>& g1 = std::get<FIELDS<MESH_NODE_ATTR<T, 3>>::g>(attrNew.Get_Unchecked(i));
        for (int k = 0; k < 3; ++k) { gmax = std::max(gmax, std::fabs(g0[k])); gdiff = std::max(gdiff, std::fabs(g0[k] - g1[k])); }
    }
    report[5] = gmax; report[6] = gdiff;
    Trips tRef, tNew;
    tNew.emplace_back(0, 0, 1.0); tRef.emplace_back(0, 0, 1.0); // appended to, not overwritten
    Compute_Barrier_Hessian<T, 3, false>(X, attrRef, csRef, infoRef, dHat2, kap, thickness, true, tRef);
    B200::Compute_Barrier_Hessian<T, 3, false>(X, attrNew, csRef, infoRef, dHat2, kap, thickness, true, tNew);
    const auto mr = merged(tRef), mn = merged(tNew);
    T hmax = 0, hdiff = 0;
    bool pattern = mr.size() == mn.size();
    for (size_t i = 0; pattern && i < mr.size(); ++i) {
        pattern = mr[i].first == mn[i].first;
        hmax = std::max(hmax, std::fabs(mr[i].second)); hdiff = std::max(hdiff, std::fabs(mr[i].second - mn[i].second));
    }
    report[7] = hmax; report[8] = hdiff; report[9] = pattern ? 1.0 : 0.0; report[10] = (double)mr.size();

    // ---- intersection-free step and min distance
    std::vector<T> sd(dir, dir + 3 * (size_t)nV);
    T aRef = 1.0, aNew = 1.0;
    Compute_Intersection_Free_StepSize<T, 3, false, false>(X, bnode, bedge, btri, particle, rod, nn, codim, DBCb, sd, thickness, aRef);
    B200::Compute_Intersection_Free_StepSize<T, 3, false, false>(X, bnode, bedge, btri, particle, rod, nn, codim, DBCb, sd, thickness, aNew);
    report[11] = aRef; report[12] = aNew;
    std::vector<T> dRef, dNew;
    T mRef = 0, mNew = 0;
    Compute_Min_Dist2<T, 3, false>(X, csRef, thickness, dRef, mRef);
    B200::Compute_Min_Dist2<T, 3, false>(X, csRef, thickness, dNew, mNew);
    bool deq = dRef.size() == dNew.size();
    for (size_t i = 0; deq && i < dRef.size(); ++i) deq = dRef[i] == dNew[i];
    report[13] = mRef; report[14] = mNew; report[15] = deq ? 1.0 : 0.0;
    return 0;
}


// The `flow` branch of Compute_IncPotential_Hessian (FEM/Shell/INC_POTENTIAL.h:321-394) side by side: the reference's
// sequence -- flow triplets (:323-339, restated here: INC_POTENTIAL.h itself needs the whole shell stack), the reference's
// Compute_Barrier_Hessian, the reference's CSR_MATRIX::Construct_From_Triplet, `+= M`, CSR_MATRIX::Project_DBC -- against
// B200::Compute_IncPotential_Hessian_Flow, which assembles everything on the device and fills the same CSR_MATRIX type
// through Construct_From_CSR. report: [nnzRef, nnzNew (stored), maxAbs, maxDiff, entries of Ref missing in New]
extern "C" int b2_flow_system(int nV, const double* x, const double* x0, int nF, const int* tri, const double* vol, const double* mass,
    const unsigned char* dbc, int nRows, const int* rows4, double h, double dHat2, double kappa, double* report)
{
    MESH_NODE<T, 3> X(nV);
    MESH_NODE_ATTR<T, 3> attr(nV);
    for (int i = 0; i < nV; ++i) {
        X.Append(VECTOR<T, 3>(x[3 * i], x[3 * i + 1], x[3 * i + 2]));
        attr.Append(VECTOR<T, 3>(x0[3 * i], x0[3 * i + 1], x0[3 * i + 2]), VECTOR<T, 3>(0.0), VECTOR<T, 3>(0.0), 0.0);
    }
    MESH_ELEM<2> Elem(nF);
    for (int i = 0; i < nF; ++i) Elem.Append(VECTOR<int, 3>(tri[3 * i], tri[3 * i + 1], tri[3 * i + 2]));
    std::vector<VECTOR<int, 4>> cs;
    for (int i = 0; i < nRows; ++i) cs.emplace_back(rows4[4 * i], rows4[4 * i + 1], rows4[4 * i + 2], rows4[4 * i + 3]);
    const std::vector<VECTOR<T, 2>> info(cs.size(), VECTOR<T, 2>(1, dHat2));
    std::vector<bool> DBCb(nV, false);
    for (int i = 0; i < nV; ++i) DBCb[i] = dbc[i] != 0;
    T kap[3] = {kappa, kappa, kappa};
    const int dim = 3;
    // reference sequence
    Trips triplets;
    for (int id = 0; id < nF; ++id)
        for (int i = 0; i < dim; ++i)
            for (int d = 0; d < dim; ++d) {
                triplets.emplace_back(tri[3 * id + i] * dim + d, tri[3 * id + (i + 1) % dim] * dim + d, -h * vol[id] / 6);
                triplets.emplace_back(tri[3 * id + i] * dim + d, tri[3 * id + (i + 2) % dim] * dim + d, -h * vol[id] / 6);
                triplets.emplace_back(tri[3 * id + i] * dim + d, tri[3 * id + i] * dim + d, 2 * h * vol[id] / 6);
            }
    Compute_Barrier_Hessian<T, 3, false>(X, attr, cs, info, dHat2, kap, T(0), true, triplets);
    CSR_MATRIX<T> sysRef, M, sysNew;
    sysRef.Construct_From_Triplet(nV * dim, nV * dim, triplets);
    Trips mt;
    for (int v = 0; v < nV; ++v)
        if (mass[v] != 0.0) for (int d = 0; d < dim; ++d) mt.emplace_back(v * dim + d, v * dim + d, mass[v]);
    M.Construct_From_Triplet(nV * dim, nV * dim, mt);
    sysRef.Get_Matrix() += M.Get_Matrix();
    sysRef.Project_DBC(DBCb, dim);
    // device path
    const std::vector<T> volv(vol, vol + nF), massv(mass, mass + nV);
    B200::Compute_IncPotential_Hessian_Flow<T, 3>(Elem, volv, h, X, attr, cs, info, dHat2, kap, T(0), true, massv, DBCb, sysNew);
    auto& A = sysRef.Get_Matrix();
    auto& B = sysNew.Get_Matrix();
    double amax = 0, dmax = 0, missing = 0;
    for (int r = 0; r < nV * dim; ++r) {
        int q = B.outerIndexPtr()[r];
        for (int p = A.outerIndexPtr()[r]; p < A.outerIndexPtr()[r + 1]; ++p) {
            const int c = A.innerIndexPtr()[p];
            while (q < B.outerIndexPtr()[r + 1] && B.innerIndexPtr()[q] < c) { dmax = std::max(dmax, std::fabs(B.valuePtr()[q])); ++q; } // explicit zeros of the 3x3 blocks
            const double a = A.valuePtr()[p];
            amax = std::max(amax, std::fabs(a));
            if (q < B.outerIndexPtr()[r + 1] && B.innerIndexPtr()[q] == c) { dmax = std::max(dmax, std::fabs(a - B.valuePtr()[q])); ++q; }
            else if (a != 0.0) missing += 1;
        }
        for (; q < B.outerIndexPtr()[r + 1]; ++q) dmax = std::max(dmax, std::fabs(B.valuePtr()[q]));
    }
    report[0] = (double)A.nonZeros(); report[1] = (double)B.nonZeros(); report[2] = amax; report[3] = dmax; report[4] = missing;
    return 0;
}
