// Structured (low-rank) evaluation of the projected barrier Hessian. Placeholder: forwards to the dense path.
#pragma once
#include "pair_deriv.cuh"
namespace idp {
IDP_HD bool row_EgH_lowrank(const RowDec& d, const V3* x, const V3* xr, double weight, double dHat2, double kappa,
    double xi2, bool projectSPD, double* E, double* g, double* H)
{
    return row_EgH(d, x, xr, weight, dHat2, kappa, xi2, projectSPD, E, g, H);
}
} // namespace idp
