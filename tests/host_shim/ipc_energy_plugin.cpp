// Test-only: the reference's PLUGIN boundary for this path -- IPC_ENERGY<T, dim, elasticIPC> : ABSTRACT_ENERGY<T, dim>
// (Library/FEM/Energy/IPC_ENERGY.h:11-58, interface Library/FEM/Energy/ENERGY.h:15-50) -- compiled from the reference's own
// header with its three barrier calls routed to JGSL::B200 (the qualifier switch INTEGRATION.md describes; here done with
// three object-like macros so that the reference header itself is compiled unmodified from /root/reference), called through
// the virtual interface and compared with the reference's free functions on the same argument objects.
// Built only where /root/reference exists (tests/host_shim/Makefile); the shared object travels to the GPU box.
#include <FEM/IPC.h>
#include "../../idp_b200/host/IPC_B200.h"
#include <algorithm>
#include <cmath>
#include <memory>

namespace JGSL {
// what ENERGY.h's signatures need beyond FEM/IPC.h (ENERGY.h itself pulls in every other energy of the volumetric stepper)
template <class T, int dim> using MESH_ELEM_ATTR = BASE_STORAGE<Eigen::Matrix<T, dim, dim>, T>;
template <class T, int dim> using FIXED_COROTATED = BASE_STORAGE<Eigen::Matrix<T, dim, dim>, T, T, T>;
// the abstract plugin interface, ENERGY.h:15-50 (signatures only)
template <class T, int dim>
class ABSTRACT_ENERGY {
public:
    virtual ~ABSTRACT_ENERGY() {}
    virtual void Compute_IncPotential(MESH_ELEM<dim>& Elem, const VECTOR<T, dim>& gravity, T h, MESH_NODE<T, dim>& X, MESH_NODE<T, dim>& Xtilde,
        MESH_NODE_ATTR<T, dim>& nodeAttr, MESH_ELEM_ATTR<T, dim>& elemAttr, FIXED_COROTATED<T, dim>& elasticityAttr,
        std::vector<VECTOR<int, dim + 1>>& constraintSet, T dHat2, T kappa[], double& value) = 0;
    virtual void Compute_IncPotential_Gradient(MESH_ELEM<dim>& Elem, const VECTOR<T, dim>& gravity, T h, MESH_NODE<T, dim>& X, MESH_NODE<T, dim>& Xtilde,
        MESH_NODE_ATTR<T, dim>& nodeAttr, MESH_ELEM_ATTR<T, dim>& elemAttr, FIXED_COROTATED<T, dim>& elasticityAttr,
        std::vector<VECTOR<int, dim + 1>>& constraintSet, T dHat2, T kappa[]) = 0;
    virtual void Compute_IncPotential_Hessian(MESH_ELEM<dim>& Elem, T h, MESH_NODE<T, dim>& X, MESH_NODE_ATTR<T, dim>& nodeAttr,
        MESH_ELEM_ATTR<T, dim>& elemAttr, FIXED_COROTATED<T, dim>& elasticityAttr, std::vector<VECTOR<int, dim + 1>>& constraintSet, T dHat2,
        T kappa[], std::vector<Eigen::Triplet<T>>& triplets) = 0;
};
} // namespace JGSL

// JGSL_USE_B200_CONTACT: the plugin's three call sites get the B200:: qualifier
#define Compute_Barrier B200::Compute_Barrier
#define Compute_Barrier_Gradient B200::Compute_Barrier_Gradient
#define Compute_Barrier_Hessian B200::Compute_Barrier_Hessian
#include <FEM/Energy/IPC_ENERGY.h>
#undef Compute_Barrier
#undef Compute_Barrier_Gradient
#undef Compute_Barrier_Hessian

using namespace JGSL;
typedef double T;

// report: [E_ref, E_plugin, gmax, gdiff, hmax, hdiff, nTripletsRef, nTripletsPlugin]
extern "C" int ipc_energy_plugin(int nV, const double* x, const double* x0, int nRows, const int* rows4, double dHat2, double kappa, double* report)
{
    MESH_NODE<T, 3> X(nV), Xtilde(nV);
    MESH_NODE_ATTR<T, 3> attrRef(nV), attrNew(nV);
    for (int i = 0; i < nV; ++i) {
        X.Append(VECTOR<T, 3>(x[3 * i], x[3 * i + 1], x[3 * i + 2]));
        Xtilde.Append(VECTOR<T, 3>(x[3 * i], x[3 * i + 1], x[3 * i + 2]));
        const VECTOR<T, 3> r(x0[3 * i], x0[3 * i + 1], x0[3 * i + 2]);
        attrRef.Append(r, VECTOR<T, 3>(0.0), VECTOR<T, 3>(0.0), 0.0);
        attrNew.Append(r, VECTOR<T, 3>(0.0), VECTOR<T, 3>(0.0), 0.0);
    }
    std::vector<VECTOR<int, 4>> cs;
    for (int i = 0; i < nRows; ++i) cs.emplace_back(rows4[4 * i], rows4[4 * i + 1], rows4[4 * i + 2], rows4[4 * i + 3]);
    const std::vector<VECTOR<T, 2>> info(cs.size(), VECTOR<T, 2>(1, dHat2)); // what the plugin builds (IPC_ENERGY.h:25-26)
    T kap[3] = {kappa, kappa, kappa};
    MESH_ELEM<3> Elem(0);
    MESH_ELEM_ATTR<T, 3> elemAttr(0);
    FIXED_COROTATED<T, 3> elasticityAttr(0);
    const VECTOR<T, 3> gravity(0.0);

    std::shared_ptr<ABSTRACT_ENERGY<T, 3>> plugin = std::make_shared<IPC_ENERGY<T, 3, false>>(); // as ENERGY<T,dim>::Add stores it (ENERGY.h:55-57)
    double eRef = 0.5, eNew = 0.5;
    Compute_Barrier<T, 3, false>(X, attrRef, cs, info, dHat2, kap, T(0), eRef);
    plugin->Compute_IncPotential(Elem, gravity, 0.01, X, Xtilde, attrNew, elemAttr, elasticityAttr, cs, dHat2, kap, eNew);
    Compute_Barrier_Gradient<T, 3, false>(X, cs, info, dHat2, kap, T(0), attrRef);
    plugin->Compute_IncPotential_Gradient(Elem, gravity, 0.01, X, Xtilde, attrNew, elemAttr, elasticityAttr, cs, dHat2, kap);
    double gmax = 0, gdiff = 0;
    for (int i = 0; i < nV; ++i) {
        const VECTOR<T, 3>& a = std::get<FIELDS<MESH_NODE_ATTR<T, 3>>::g>(attrRef.Get_Unchecked(i));
        const VECTOR<T, 3>& b = std::get<FIELDS<MESH_NODE_ATTR<T, 3>>::g>(attrNew.Get_Unchecked(i));
        for (int d = 0; d < 3; ++d) { gmax = std::max(gmax, std::fabs(a[d])); gdiff = std::max(gdiff, std::fabs(a[d] - b[d])); }
    }
    std::vector<Eigen::Triplet<T>> tRef, tNew;
    Compute_Barrier_Hessian<T, 3, false>(X, attrRef, cs, info, dHat2, kap, T(0), true, tRef);
    plugin->Compute_IncPotential_Hessian(Elem, 0.01, X, attrNew, elemAttr, elasticityAttr, cs, dHat2, kap, tNew);
    auto merged = [](const std::vector<Eigen::Triplet<T>>& t) {
        std::vector<std::pair<std::pair<int, int>, T>> v;
        for (const auto& e : t) v.push_back({{e.row(), e.col()}, e.value()});
        std::sort(v.begin(), v.end(), [](const auto& a, const auto& b) { return a.first < b.first; });
        size_t o = 0;
        for (size_t i = 0; i < v.size(); ++i) {
            if (o && v[o - 1].first == v[i].first) v[o - 1].second += v[i].second;
            else v[o++] = v[i];
        }
        v.resize(o);
        return v;
    };
    const auto mr = merged(tRef), mn = merged(tNew);
    double hmax = 0, hdiff = 0;
    int samePattern = mr.size() == mn.size();
    for (size_t i = 0; i < mr.size() && samePattern; ++i) {
        if (mr[i].first != mn[i].first) { samePattern = 0; break; }
        hmax = std::max(hmax, std::fabs(mr[i].second));
        hdiff = std::max(hdiff, std::fabs(mr[i].second - mn[i].second));
    }
    report[0] = eRef; report[1] = eNew; report[2] = gmax; report[3] = gdiff; report[4] = hmax; report[5] = hdiff;
    report[6] = (double)mr.size(); report[7] = samePattern ? (double)mn.size() : -1.0;
    return 0;
}
