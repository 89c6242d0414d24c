// Test-only: the drop-in boundary exercised with the REFERENCE's own types and operators in the same translation unit.
// <FEM/IPC.h> is the reference's header (compiled from /root/reference against the stand-ins of oracle/ref_shim/include);
// "IPC_B200.h" is this repository's host mirror. Both sets of six operators are called with the SAME argument objects
// (MESH_NODE / MESH_NODE_ATTR storages, std::vector<VECTOR<int,4>>, std::vector<Eigen::Triplet<double>> ...), exactly as a
// maintainer would switch the five call sites, and the results are compared. Built only where /root/reference exists
// (tests/host_shim/Makefile); the shared object travels to the GPU box.
#include <FEM/IPC.h>
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
