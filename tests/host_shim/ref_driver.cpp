// TEST INFRASTRUCTURE: compiles the REFERENCE's own Newton driver -- Advance_One_Step_IE_Discrete_Shell and Line_Search of
// FEM/Shell/IMPLICIT_EULER.h with Compute_IncPotential* of FEM/Shell/INC_POTENTIAL.h, from where they lie under
// /root/reference/Library, together with its contact loops (FEM/IPC.h), shell energies (MEMBRANE.h, BENDING.h), friction,
// Dirichlet penalty and CSR_MATRIX -- against the stand-ins of oracle/ref_shim/include, and exposes one time step to C.
// What is NOT the reference's here, because it cannot be built in this image: the storages (std::vector stand-in), Eigen (the
// repository's subset), the linear solver behind Solve_Direct (a Jacobi-preconditioned CG to 1e-12 instead of CHOLMOD), the
// surface-primitive extraction (the oracle's restatement of Find_Surface_Primitives_And_Compute_Area) and the features the
// two paper configurations never execute (strain limiting, fibers, tets: no-op / aborting stubs below).
// Used to check the repository's restatement of that driver (idp_b200/host/jgsl/shell_flow.h): tests/test_ref_driver.py.
#include <memory>
#include <deque>
#include <FEM/DATA_TYPE.h>
#include <Math/CSR_MATRIX.h>
namespace JGSL { template <class T, int dim> using MPM_STRESS = BASE_STORAGE<MATRIX<T, dim>>; } // named by Physics/FIXED_COROTATED.h only
#include <Math/UTILS.h>
#include <Math/DIHEDRAL_ANGLE.h>
#include <Physics/FIXED_COROTATED.h>
#include <FEM/IPC.h>
#include <FEM/FRICTION.h>
#include "../../oracle/orc_system.hpp"

namespace JGSL {

[[noreturn]] inline void not_in_the_paper_configs(const char* what)
{
    printf("ref_driver: %s is not part of this build (never executed by the two paper configurations)\n", what);
    exit(-1);
}
// FEM/Shell/INEXT.h, ANISO_INEXT.h (strain limiting, fibers: need JacobiSVD / FullPivLU): reached only with kappa_s > 0 / fiberStiffMult > 0
template <class... A> void Compute_Inextensibility(A&&...) { not_in_the_paper_configs("Compute_Inextensibility"); }
template <class... A> bool Compute_Inextensibility_Energy(A&&...) { not_in_the_paper_configs("Compute_Inextensibility_Energy"); }
template <class... A> void Compute_Inextensibility_Gradient(A&&...) { not_in_the_paper_configs("Compute_Inextensibility_Gradient"); }
template <class... A> void Compute_Inextensibility_Hessian(A&&...) { not_in_the_paper_configs("Compute_Inextensibility_Hessian"); }
template <class... A> bool Check_Fiber_Feasibility(A&&...) { not_in_the_paper_configs("Check_Fiber_Feasibility"); }
template <class... A> void Compute_Fiber_Energy(A&&...) { not_in_the_paper_configs("Compute_Fiber_Energy"); }
template <class... A> void Compute_Fiber_Gradient(A&&...) { not_in_the_paper_configs("Compute_Fiber_Gradient"); }
template <class... A> void Compute_Fiber_Hessian(A&&...) { not_in_the_paper_configs("Compute_Fiber_Hessian"); }
// volumetric elasticity of tets: called unconditionally, on empty storages in a tet-free scene -> nothing to do
template <class T, int dim>
struct NEOHOOKEAN_FUNCTOR {
    typedef BASE_STORAGE<Eigen::Matrix<T, dim * dim, dim * dim>> DIFFERENTIAL;
    template <class... A> static void Compute_Psi(A&&...) {}
    template <class... A> static void Compute_First_PiolaKirchoff_Stress(A&&...) {}
    template <class... A> static void Compute_First_PiolaKirchoff_Stress_Derivative(A&&...) {}
};
template <class... A> void Compute_Deformation_Gradient(A&&...) {}
template <class... A> void Elem_To_Node(A&&...) {}
template <class T, bool b, class... A> void Find_Surface_TriMesh(A&&...) {} // no tets: no extra surface triangles
template <class... Ts> void Append_Attribute(BASE_STORAGE<Ts...>& src, BASE_STORAGE<Ts...>& dst) // Utils/MESHIO.h:1016-1032
{
    for (int i = 0; i < src.size; ++i) { dst.rows.push_back(src.rows[i]); ++dst.size; }
}
// Find_Surface_Primitives_And_Compute_Area (Utils/MESHIO.h:768-834; that header needs the real storages): the oracle's restatement
template <class T>
void Find_Surface_Primitives_And_Compute_Area(MESH_NODE<T, 3>& X, MESH_ELEM<2>& Tri, std::vector<int>& bn, std::vector<VECTOR<int, 2>>& be,
    std::vector<VECTOR<int, 3>>& bt, std::vector<T>& BNArea, std::vector<T>& BEArea, std::vector<T>& BTArea)
{
    std::vector<double> x(3 * (size_t)X.size);
    std::vector<int> tri(3 * (size_t)Tri.size);
    for (int i = 0; i < X.size; ++i) for (int k = 0; k < 3; ++k) x[3 * (size_t)i + k] = std::get<0>(X.Get_Unchecked(i))[k];
    for (int i = 0; i < Tri.size; ++i) for (int k = 0; k < 3; ++k) tri[3 * (size_t)i + k] = std::get<0>(Tri.Get_Unchecked(i))[k];
    orc::SurfacePrimitives S;
    orc::find_surface_primitives(X.size, Tri.size, tri.data(), x.data(), S);
    bn = S.bnode; BNArea = S.BNArea; BEArea = S.BEArea; BTArea = S.BTArea;
    for (size_t i = 0; i < S.bedge.size() / 2; ++i) be.emplace_back(S.bedge[2 * i], S.bedge[2 * i + 1]);
    for (size_t i = 0; i < S.btri.size() / 3; ++i) bt.emplace_back(S.btri[3 * i], S.btri[3 * i + 1], S.btri[3 * i + 2]);
}
// stretch statistics (FEM/Shell/DISCRETE_SHELL.h:601-671, needs the SVD routine): only feed stretch.txt
template <class T, int dim>
void Compute_Max_And_Avg_Stretch(MESH_ELEM<dim - 1>&, T, const VECTOR<T, 4>&, const std::vector<bool>&, MESH_NODE<T, dim>&, MESH_NODE_ATTR<T, dim>&,
    MESH_ELEM_ATTR<T, dim - 1>&, FIXED_COROTATED<T, dim - 1>&, T& maxs, T& avgs, T& minc, T& avgc)
{
    maxs = 1; avgs = 0; minc = 1; avgc = 0;
}
template <class... A> void Compute_Max_And_Avg_Stretch_Rod(A&&...) {}
// Solve_Direct (Math/DIRECT_SOLVER.h:14-88: CHOLMOD / SimplicialLDLT): Jacobi-preconditioned CG to 1e-12 on the CSR arrays
template <class T>
bool Solve_Direct(CSR_MATRIX<T>& A, const std::vector<T>& rhs, std::vector<T>& sol)
{
    auto& S = A.Get_Matrix();
    const int n = S.rows();
    if (n != (int)rhs.size()) { printf("sysMtr dimension does not match with rhs!\n"); return false; }
    const int* ptr = S.outerIndexPtr(); const int* col = S.innerIndexPtr(); const T* val = S.valuePtr();
    std::vector<T> dinv(n), r(rhs), z(n), p(n), Ap(n);
    for (int i = 0; i < n; ++i) {
        T dii = 0;
        for (int k = ptr[i]; k < ptr[i + 1]; ++k) if (col[k] == i) dii = val[k];
        if (!(dii > 0)) return false;
        dinv[i] = 1.0 / dii;
    }
    sol.assign(n, T(0));
    T bnorm = 0;
    for (int i = 0; i < n; ++i) bnorm += rhs[i] * rhs[i];
    if (bnorm == 0) return true;
    T rz = 0, rr = bnorm;
    for (int i = 0; i < n; ++i) { z[i] = dinv[i] * r[i]; p[i] = z[i]; rz += r[i] * z[i]; }
    for (int it = 0; it < 100000 && rr > 1e-24 * bnorm; ++it) {
#pragma omp parallel for
        for (int i = 0; i < n; ++i) { T s = 0; for (int k = ptr[i]; k < ptr[i + 1]; ++k) s += val[k] * p[col[k]]; Ap[i] = s; }
        T pAp = 0;
        for (int i = 0; i < n; ++i) pAp += Ap[i] * p[i];
        if (!(pAp > 0)) return false;
        const T a = rz / pAp;
        T rzn = 0; rr = 0;
        for (int i = 0; i < n; ++i) { sol[i] += a * p[i]; r[i] -= a * Ap[i]; z[i] = dinv[i] * r[i]; rzn += r[i] * z[i]; rr += r[i] * r[i]; }
        const T beta = rzn / rz; rz = rzn;
        for (int i = 0; i < n; ++i) p[i] = z[i] + beta * p[i];
    }
    return true;
}

} // namespace JGSL

// the reference's own headers that include the two unbuildable ones are given those names first
#define JGSL_REF_DRIVER_STUBS 1
#include <FEM/Shell/IMPLICIT_EULER.h>

using namespace JGSL;
typedef double T;

extern "C" int refdrv_advance(int flow, int nV, double* X, double* vel, const double* x0, const double* mass, int nE, const int* elem3, const double* ib3,
    double hingeK, const double* vol, const double* lam, const double* mu, int nH, const int* stencil4, const double* info3, int nDBC, const double* dbc4,
    const double* b3, double h, double tol, int withCollision, double dHat2, double* kappa3, double muFric, double epsv2, int fricIterAmt, double thickness,
    double bendingStiffMult, int nComp, const int* compRange, int nMuComp, const double* muCompIn, const char* outputFolder)
{
    MESH_NODE<T, 3> Xs;
    MESH_NODE_ATTR<T, 3> nodeAttr;
    for (int i = 0; i < nV; ++i) {
        Xs.Append(VECTOR<T, 3>(X[3 * i], X[3 * i + 1], X[3 * i + 2]));
        nodeAttr.Append(VECTOR<T, 3>(x0[3 * i], x0[3 * i + 1], x0[3 * i + 2]), VECTOR<T, 3>(vel[3 * i], vel[3 * i + 1], vel[3 * i + 2]), VECTOR<T, 3>(0.0), mass[i]);
    }
    MESH_ELEM<2> Elem;
    MESH_ELEM_ATTR<T, 2> elemAttr;
    FIXED_COROTATED<T, 2> fcr;
    for (int e = 0; e < nE; ++e) {
        Elem.Append(VECTOR<int, 3>(elem3[3 * e], elem3[3 * e + 1], elem3[3 * e + 2]));
        MATRIX<T, 2> IB, P;
        IB(0, 0) = ib3[3 * e]; IB(0, 1) = IB(1, 0) = ib3[3 * e + 1]; IB(1, 1) = ib3[3 * e + 2];
        if (e == 0) P(0, 0) = hingeK;
        elemAttr.Append(IB, P);
        fcr.Append(MATRIX<T, 2>(), vol[e], lam[e], mu[e]);
    }
    std::vector<VECTOR<int, 4>> edgeStencil;
    std::vector<VECTOR<T, 3>> edgeInfo;
    for (int i = 0; i < nH; ++i) {
        edgeStencil.emplace_back(stencil4[4 * i], stencil4[4 * i + 1], stencil4[4 * i + 2], stencil4[4 * i + 3]);
        edgeInfo.emplace_back(info3[3 * i], info3[3 * i + 1], info3[3 * i + 2]);
    }
    VECTOR_STORAGE<T, 4> DBC;
    for (int i = 0; i < nDBC; ++i) DBC.Append(VECTOR<T, 4>(dbc4[4 * i], dbc4[4 * i + 1], dbc4[4 * i + 2], dbc4[4 * i + 3]));
    std::vector<Eigen::Triplet<T>> mt;
    for (int i = 0; i < nV; ++i) for (int d = 0; d < 3; ++d) mt.emplace_back(3 * i + d, 3 * i + d, mass[i]);
    CSR_MATRIX<T> M;
    M.Construct_From_Triplet(3 * nV, 3 * nV, mt);
    std::vector<T> b(b3, b3 + 3 * (size_t)nV);
    const std::vector<VECTOR<int, 2>> seg, rod;
    const std::map<std::pair<int, int>, int> edge2tri;
    VECTOR<T, 4> fiberStiffMult(0, 0, 0, 0);
    VECTOR<T, 3> fiberLimit(0, 0, 0), kappaVec(kappa3[0], kappa3[1], kappa3[2]);
    VECTOR<T, 2> s(1.01, 0), sHat(1, 1), kappa_s(0, 0);
    const std::vector<int> compNodeRange(compRange, compRange + nComp), particle;
    const std::vector<T> muComp(muCompIn, muCompIn + nMuComp), stitchRatio;
    MESH_ELEM<3> tet;
    MESH_ELEM_ATTR<T, 3> tetAttr;
    FIXED_COROTATED<T, 3> tetFcr;
    const std::vector<VECTOR<T, 3>> rodInfo, rodHingeInfo;
    const std::vector<VECTOR<int, 3>> rodHinge, stitchInfo;
    const std::string out(outputFolder);
    int it;
    if (flow)
        it = Advance_One_Step_IE_Discrete_Shell<T, 3, false, false, true>(Elem, seg, DBC, edge2tri, edgeStencil, edgeInfo, thickness, bendingStiffMult, fiberStiffMult,
            fiberLimit, s, sHat, kappa_s, b, h, tol, withCollision != 0, dHat2, kappaVec, muFric, epsv2, fricIterAmt, compNodeRange, muComp, false, Xs, nodeAttr, M,
            elemAttr, fcr, tet, tetAttr, tetFcr, rod, rodInfo, rodHinge, rodHingeInfo, stitchInfo, stitchRatio, 10.0, particle, out);
    else
        it = Advance_One_Step_IE_Discrete_Shell<T, 3, false, false, false>(Elem, seg, DBC, edge2tri, edgeStencil, edgeInfo, thickness, bendingStiffMult, fiberStiffMult,
            fiberLimit, s, sHat, kappa_s, b, h, tol, withCollision != 0, dHat2, kappaVec, muFric, epsv2, fricIterAmt, compNodeRange, muComp, false, Xs, nodeAttr, M,
            elemAttr, fcr, tet, tetAttr, tetFcr, rod, rodInfo, rodHinge, rodHingeInfo, stitchInfo, stitchRatio, 10.0, particle, out);
    for (int i = 0; i < nV; ++i) {
        const VECTOR<T, 3>& x = std::get<0>(Xs.Get_Unchecked(i));
        const VECTOR<T, 3>& v = std::get<FIELDS<MESH_NODE_ATTR<T, 3>>::v>(nodeAttr.Get_Unchecked(i));
        for (int k = 0; k < 3; ++k) { X[3 * i + k] = x[k]; vel[3 * i + k] = v[k]; }
    }
    for (int k = 0; k < 3; ++k) kappa3[k] = kappaVec[k];
    return it;
}
