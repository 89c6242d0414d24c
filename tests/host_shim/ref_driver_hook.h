// TEST INFRASTRUCTURE: lets the reference-loop build of the JGSL module (tests/host_shim/jgsl_ref/JGSL.so) hand a whole time step
// to the REFERENCE's own Newton driver (tests/host_shim/libref_driver.so: Advance_One_Step_IE_Discrete_Shell of
// FEM/Shell/IMPLICIT_EULER.h compiled from /root/reference) when the environment variable JGSL_REF_DRIVER is set: the scripts,
// the module's storages and set-up functions stay, the time step itself is the reference's code. Used to check the
// repository's restatement of that driver (idp_b200/host/jgsl/shell_flow.h) against the original.
#pragma once
#include <cstdlib>
#include <vector>

#include "shell_flow.h"

extern "C" int refdrv_advance(int flow, int nV, double* X, double* vel, const double* x0, const double* mass, int nE, const int* elem3, const double* ib3,
    double hingeK, const double* vol, const double* lam, const double* mu, int nH, const int* stencil4, const double* info3, int nDBC, const double* dbc4,
    const double* b3, double h, double tol, int withCollision, double dHat2, double* kappa3, double muFric, double epsv2, int fricIterAmt, double thickness,
    double bendingStiffMult, int nComp, const int* compRange, int nMuComp, const double* muComp, const char* outputFolder);

namespace jgsl {

inline bool reference_driver_step(const ShellStepInputs& in, TriStorage& Elem, DbcStorage& DBC, const std::vector<Vec<int, 4>>& edgeStencil,
    const std::vector<Vec<double, 3>>& edgeInfo, const std::vector<double>& b, Vec<double, 3>& kappaVec, NodeStorage& X, NodeAttrStorage& nodeAttr,
    ElemAttrStorage& elemAttr, Fcr2Storage& fcr, int* iterations)
{
    if (!std::getenv("JGSL_REF_DRIVER")) return false;
    const int nV = X.size(), nE = Elem.size(), nH = (int)edgeStencil.size(), nD = DBC.size();
    std::vector<double> x(3 * (size_t)nV), v(3 * (size_t)nV), x0(3 * (size_t)nV), m((size_t)nV);
    for (int i = 0; i < nV; ++i) {
        for (int k = 0; k < 3; ++k) {
            x[3 * (size_t)i + k] = std::get<0>(X.rows[i])[k];
            x0[3 * (size_t)i + k] = std::get<0>(nodeAttr.rows[i])[k];
            v[3 * (size_t)i + k] = std::get<1>(nodeAttr.rows[i])[k];
        }
        m[i] = std::get<3>(nodeAttr.rows[i]);
    }
    std::vector<int> elem(3 * (size_t)nE), st(4 * (size_t)nH);
    std::vector<double> ib(3 * (size_t)nE), vol((size_t)nE), lam((size_t)nE), mu((size_t)nE), info(3 * (size_t)nH), dbc(4 * (size_t)nD);
    for (int e = 0; e < nE; ++e) {
        for (int k = 0; k < 3; ++k) elem[3 * (size_t)e + k] = std::get<0>(Elem.rows[e])[k];
        const Mat<double, 2>& IB = std::get<0>(elemAttr.rows[e]);
        ib[3 * (size_t)e] = IB(0, 0); ib[3 * (size_t)e + 1] = IB(0, 1); ib[3 * (size_t)e + 2] = IB(1, 1);
        vol[e] = std::get<1>(fcr.rows[e]); lam[e] = std::get<2>(fcr.rows[e]); mu[e] = std::get<3>(fcr.rows[e]);
    }
    for (int i = 0; i < nH; ++i) {
        for (int k = 0; k < 4; ++k) st[4 * (size_t)i + k] = edgeStencil[i][k];
        for (int k = 0; k < 3; ++k) info[3 * (size_t)i + k] = edgeInfo[i][k];
    }
    for (int i = 0; i < nD; ++i) for (int k = 0; k < 4; ++k) dbc[4 * (size_t)i + k] = std::get<0>(DBC.rows[i])[k];
    const double hingeK = nE ? std::get<1>(elemAttr.rows[0])(0, 0) : 0.0;
    double kappa[3] = {kappaVec[0], kappaVec[1], kappaVec[2]};
    *iterations = refdrv_advance(in.flow ? 1 : 0, nV, x.data(), v.data(), x0.data(), m.data(), nE, elem.data(), ib.data(), hingeK, vol.data(), lam.data(), mu.data(),
        nH, st.data(), info.data(), nD, dbc.data(), b.data(), in.h, in.NewtonTol, in.withCollision ? 1 : 0, in.dHat2, kappa, in.mu, in.epsv2, in.fricIterAmt,
        in.thickness, in.bendingStiffMult, (int)in.compNodeRange.size(), in.compNodeRange.data(), (int)in.muComp.size(), in.muComp.data(), in.outputFolder.c_str());
    for (int i = 0; i < nV; ++i)
        for (int k = 0; k < 3; ++k) { std::get<0>(X.rows[i])[k] = x[3 * (size_t)i + k]; std::get<1>(nodeAttr.rows[i])[k] = v[3 * (size_t)i + k]; }
    for (int k = 0; k < 3; ++k) kappaVec[k] = kappa[k];
    return true;
}

} // namespace jgsl
#define JGSL_STEP_HOOK jgsl::reference_driver_step
