// ORACLE / TEST INFRASTRUCTURE: compiles the REFERENCE's own shell energy terms -- FEM/Shell/MEMBRANE.h and FEM/Shell/BENDING.h
// (KL = false: hinge bending), from where they lie under /root/reference/Library -- against the stand-ins in include/ and exposes
// Compute_Membrane_* / Compute_Bending_* <double, 3> to ctypes. Pins oracle/orc_elastic.hpp (and through it the CUDA path).
// No reference source is copied into this repository.
#include <memory>
#include <FEM/DATA_TYPE.h>
#include <Math/CSR_MATRIX.h>
namespace JGSL { template <class T, int dim> using MPM_STRESS = BASE_STORAGE<MATRIX<T, dim>>; } // named by Physics/FIXED_COROTATED.h, never used here
#include <Math/UTILS.h>
#include <Math/DIHEDRAL_ANGLE.h>
#include <FEM/Shell/UTILS.h>
#include <FEM/Shell/MEMBRANE.h>
#include <FEM/Shell/BENDING.h>

using namespace JGSL;
typedef double T;

namespace {
struct Shell {
    MESH_NODE<T, 3> X;
    MESH_NODE_ATTR<T, 3> nodeAttr;
    MESH_ELEM<2> Elem;
    MESH_ELEM_ATTR<T, 2> elemAttr;
    FIXED_COROTATED<T, 2> fcr;
    std::vector<bool> DBCb;
    Shell(int nV, const double* x, int nE, const int* elem3, const double* ib3, const double* vol, const double* lambda, const double* mu,
        const unsigned char* dbc, double hingeK)
    {
        for (int i = 0; i < nV; ++i) {
            X.Append(VECTOR<T, 3>(x[3 * i], x[3 * i + 1], x[3 * i + 2]));
            nodeAttr.Append(VECTOR<T, 3>(0.0), VECTOR<T, 3>(0.0), VECTOR<T, 3>(0.0), 0.0);
        }
        for (int e = 0; e < nE; ++e) {
            Elem.Append(VECTOR<int, 3>(elem3[3 * e], elem3[3 * e + 1], elem3[3 * e + 2]));
            MATRIX<T, 2> IB, P;
            IB(0, 0) = ib3[3 * e]; IB(0, 1) = IB(1, 0) = ib3[3 * e + 1]; IB(1, 1) = ib3[3 * e + 2];
            if (e == 0) P(0, 0) = hingeK; // the hinge stiffness lives in P(0,0) of element 0 (DISCRETE_SHELL.h:336-341)
            elemAttr.Append(IB, P);
            fcr.Append(MATRIX<T, 2>(), vol[e], lambda[e], mu[e]);
        }
        DBCb.assign(nV, false);
        if (dbc) for (int i = 0; i < nV; ++i) DBCb[i] = dbc[i] != 0;
    }
    void add_gradient(double* g)
    {
        for (int i = 0; i < X.size; ++i) {
            const VECTOR<T, 3>& gi = std::get<FIELDS<MESH_NODE_ATTR<T, 3>>::g>(nodeAttr.Get_Unchecked(i));
            for (int k = 0; k < 3; ++k) g[3 * i + k] += gi[k];
        }
    }
};
long copy_triplets(const std::vector<Eigen::Triplet<T>>& trip, long cap, int* tr, int* tc, double* tv)
{
    const long nt = (long)trip.size();
    for (long i = 0; i < nt && i < cap; ++i) { tr[i] = trip[i].row(); tc[i] = trip[i].col(); tv[i] = trip[i].value(); }
    return nt;
}
} // namespace

extern "C" {

// Compute_Membrane_Energy / _Gradient / _Hessian (MEMBRANE.h:8-315): E added, g (nV x 3) added, triplets; returns their count
long refshell_membrane(int nV, const double* x, int nE, const int* elem3, const double* ib3, const double* vol, const double* lambda, const double* mu,
    const unsigned char* dbc, double h, int projectSPD, double* E, double* g, long cap, int* tr, int* tc, double* tv)
{
    Shell s(nV, x, nE, elem3, ib3, vol, lambda, mu, dbc, 0.0);
    if (E) Compute_Membrane_Energy<T, 3>(s.Elem, h, s.DBCb, s.X, s.nodeAttr, s.elemAttr, s.fcr, *E);
    if (g) { Compute_Membrane_Gradient<T, 3>(s.Elem, h, s.DBCb, s.X, s.nodeAttr, s.elemAttr, s.fcr); s.add_gradient(g); }
    if (!tr) return 0;
    std::vector<Eigen::Triplet<T>> trip;
    Compute_Membrane_Hessian<T, 3>(s.Elem, h, projectSPD != 0, s.DBCb, s.X, s.nodeAttr, s.elemAttr, s.fcr, trip);
    return copy_triplets(trip, cap, tr, tc, tv);
}

// Compute_Bending_Energy / _Gradient / _Hessian <T, 3, KL=false> (BENDING.h:10-500): hinge stencils + (thetabar, ebar, hbar), stiffness k
long refshell_hinges(int nV, const double* x, int nH, const int* stencil4, const double* info3, double k, double bendingStiffMult, const unsigned char* dbc,
    double h, int projectSPD, double* E, double* g, long cap, int* tr, int* tc, double* tv)
{
    const int elem[3] = {0, 1, 2};
    const double ib[3] = {1, 0, 1}, one = 1.0;
    Shell s(nV, x, 1, elem, ib, &one, &one, &one, dbc, k);
    std::map<std::pair<int, int>, int> edge2tri;
    std::vector<VECTOR<int, 4>> st;
    std::vector<VECTOR<T, 3>> info;
    for (int i = 0; i < nH; ++i) {
        st.emplace_back(stencil4[4 * i], stencil4[4 * i + 1], stencil4[4 * i + 2], stencil4[4 * i + 3]);
        info.emplace_back(info3[3 * i], info3[3 * i + 1], info3[3 * i + 2]);
    }
    if (E) Compute_Bending_Energy<T, 3, false>(s.Elem, h, edge2tri, st, info, 0.0, bendingStiffMult, s.DBCb, s.X, s.nodeAttr, s.elemAttr, s.fcr, *E);
    if (g) {
        Compute_Bending_Gradient<T, 3, false>(s.Elem, h, edge2tri, st, info, 0.0, bendingStiffMult, s.DBCb, s.X, s.nodeAttr, s.elemAttr, s.fcr);
        s.add_gradient(g);
    }
    if (!tr) return 0;
    std::vector<Eigen::Triplet<T>> trip;
    Compute_Bending_Hessian<T, 3, false>(s.Elem, h, projectSPD != 0, edge2tri, st, info, 0.0, bendingStiffMult, s.DBCb, s.X, s.nodeAttr, s.elemAttr, s.fcr, trip);
    return copy_triplets(trip, cap, tr, tc, tv);
}

} // extern "C"
