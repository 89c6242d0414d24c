// Test-only driver: calls the six operators of idp_b200/host/IPC_B200.h with the reference's call pattern
// (Shell/IMPLICIT_EULER.h:331,419,430,458,495 and Line_Search :97-132) on mock JGSL containers, returns flat arrays.
#include "jgsl_mock.h"
#include "../../idp_b200/host/IPC_B200.h"
#include <cstring>
using namespace JGSL;
typedef VECTOR<int, 2> VI2; typedef VECTOR<int, 3> VI3; typedef VECTOR<int, 4> VI4; typedef VECTOR<double, 2> VT2;

extern "C" int b2_run(int nV, const double* X, const double* X0, int nBN, const int* bnode, int nBE, const int* bedge, int nBT,
    const int* btri, const double* dir, double dHat2, double kappa0, double thickness,
    int* nRowsOut, int* rowsOut, int rowsCap, double* Eout, double* gOut, long* nTripOut, int* tr, int* tc, double* tv, long tripCap,
    double* alphaOut, double* dist2Out, double* minDist2Out)
{
    MESH_NODE<double, 3> Xs;
    MESH_NODE_ATTR<double, 3> attr;
    for (int i = 0; i < nV; ++i) {
        Xs.Append(VECTOR<double, 3>(X[3 * i], X[3 * i + 1], X[3 * i + 2]));
        attr.Append(VECTOR<double, 3>(X0[3 * i], X0[3 * i + 1], X0[3 * i + 2]), VECTOR<double, 3>(), VECTOR<double, 3>(), 1.0);
    }
    std::vector<int> bn(bnode, bnode + nBN), particle;
    std::vector<VI2> be, rod; std::vector<VI3> bt;
    for (int i = 0; i < nBE; ++i) be.emplace_back(bedge[2 * i], bedge[2 * i + 1]);
    for (int i = 0; i < nBT; ++i) bt.emplace_back(btri[3 * i], btri[3 * i + 1], btri[3 * i + 2]);
    std::map<int, std::set<int>> NNE;
    std::vector<double> BNA(nBN, 1.0), BEA(nBE, 1.0), BTA(nBT, 1.0), searchDir(dir, dir + 3 * (size_t)nV);
    std::vector<bool> DBCb(nV, false);
    VI2 codim(nBN, nBN);
    double kappa[3] = {kappa0, 100 * kappa0, 0};
    // CCD first (IMPLICIT_EULER.h:331), then constraint set, energy, gradient, Hessian, min distance
    double alpha = 1.0;
    B200::Compute_Intersection_Free_StepSize<double, 3>(Xs, bn, be, bt, particle, rod, NNE, codim, DBCb, searchDir, thickness, alpha);
    *alphaOut = alpha;
    std::vector<VI4> cs; std::vector<VI2> ptee; std::vector<VT2> info;
    B200::Compute_Constraint_Set<double, 3>(Xs, attr, bn, be, bt, particle, rod, NNE, BNA, BEA, BTA, codim, DBCb, dHat2, thickness, false, cs, ptee, info);
    *nRowsOut = (int)cs.size();
    for (size_t i = 0; i < cs.size() && (int)i < rowsCap; ++i) for (int k = 0; k < 4; ++k) rowsOut[4 * i + k] = cs[i][k];
    double E = 0.25; // adds to the incoming value
    B200::Compute_Barrier<double, 3>(Xs, attr, cs, info, dHat2, kappa, thickness, E);
    *Eout = E;
    B200::Compute_Barrier_Gradient<double, 3>(Xs, cs, info, dHat2, kappa, thickness, attr);
    for (int i = 0; i < nV; ++i) { auto& g = std::get<2>(attr.Get_Unchecked(i)); gOut[3 * i] = g[0]; gOut[3 * i + 1] = g[1]; gOut[3 * i + 2] = g[2]; }
    std::vector<Eigen::Triplet<double>> trip;
    trip.emplace_back(0, 0, 1.0); // pre-existing content must be preserved (the reference appends)
    B200::Compute_Barrier_Hessian<double, 3>(Xs, attr, cs, info, dHat2, kappa, thickness, true, trip);
    *nTripOut = (long)trip.size();
    for (size_t i = 0; i < trip.size() && (long)i < tripCap; ++i) { tr[i] = trip[i].row(); tc[i] = trip[i].col(); tv[i] = trip[i].value(); }
    std::vector<double> d2; double mn = -1;
    B200::Compute_Min_Dist2<double, 3>(Xs, cs, thickness, d2, mn);
    *minDist2Out = mn;
    for (size_t i = 0; i < d2.size() && (int)i < rowsCap; ++i) dist2Out[i] = d2[i];
    return 0;
}
