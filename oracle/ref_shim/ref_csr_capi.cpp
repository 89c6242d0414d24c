// ORACLE / TEST INFRASTRUCTURE: compiles the REFERENCE's own Math/CSR_MATRIX.h (Construct_From_Triplet, Project_DBC) from
// where it lies under /root/reference/Library against the stand-ins in include/ (Eigen::SparseMatrix subset, inert
// pybind11, no-op profiler) and runs the caller's sequence of FEM/Shell/INC_POTENTIAL.h:382-394:
//   sysMtr.Construct_From_Triplet(..., triplets); sysMtr.Get_Matrix() += M.Get_Matrix(); sysMtr.Project_DBC(DBCb, dim).
// No reference source is copied into this repository.
#include <algorithm>
#include <vector>
#include "eigen_shim.hpp"
#include <Math/CSR_MATRIX.h>

extern "C" long ref_csr_system(int n, long nT, const int* r, const int* c, const double* v, const double* mdiag, const unsigned char* dbc, int dim,
    int* ptr, int* col, double* val, long cap)
{
    std::vector<Eigen::Triplet<double>> triplets;
    triplets.reserve(nT);
    for (long k = 0; k < nT; ++k) triplets.emplace_back(r[k], c[k], v[k]);
    JGSL::CSR_MATRIX<double> sysMtr;
    sysMtr.Construct_From_Triplet(n, n, triplets);
    if (mdiag) {
        std::vector<Eigen::Triplet<double>> mt;
        for (int i = 0; i < n; ++i) if (mdiag[i] != 0.0) mt.emplace_back(i, i, mdiag[i]);
        JGSL::CSR_MATRIX<double> M;
        M.Construct_From_Triplet(n, n, mt);
        sysMtr.Get_Matrix() += M.Get_Matrix();
    }
    if (dbc) {
        std::vector<bool> DBCb(n / dim);
        for (int i = 0; i < n / dim; ++i) DBCb[i] = dbc[i] != 0;
        sysMtr.Project_DBC(DBCb, dim);
    }
    auto& A = sysMtr.Get_Matrix();
    const long nnz = A.nonZeros();
    if (nnz > cap) return -nnz;
    std::copy(A.outerIndexPtr(), A.outerIndexPtr() + n + 1, ptr);
    std::copy(A.innerIndexPtr(), A.innerIndexPtr() + nnz, col);
    std::copy(A.valuePtr(), A.valuePtr() + nnz, val);
    return nnz;
}
