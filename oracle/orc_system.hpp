// ORACLE — TEST INFRASTRUCTURE ONLY (see orc_math.hpp header). Never linked into or called by the product.
// CPU restatement of the caller-side steps around the barrier Hessian, following
//   Library/FEM/Shell/INC_POTENTIAL.h:323-339 (Laplacian flow-term triplets), :382-394 (Construct_From_Triplet, `+= M`, Project_DBC)
//   Library/FEM/Shell/DISCRETE_SHELL.h:279-318 (lumped mass matrix: diagonal triplets per element vertex)
//   Library/Math/CSR_MATRIX.h:130-141 (Project_DBC)
//   Library/Utils/MESHIO.h:768-834 (Find_Surface_Primitives_And_Compute_Area: std::map ordering contract, areas)
// Parity pin: Project_DBC and Construct_From_Triplet are additionally checked against the reference's own CSR_MATRIX.h
// compiled in oracle/_ref (ref_shim, Eigen::SparseMatrix stand-in); the flow / mass / surface restatements have no compiled
// reference counterpart here (their headers pull in the whole FEM stack) -- "parity unpinned" for those three.
#pragma once
#include "orc_ipc.hpp"
#include <map>

namespace orc {

// INC_POTENTIAL.h:323-339 — for element `id`, vertex i, axis d: three triplets, in this order:
//   (v_i d, v_{i+1} d, -h vol / 6), (v_i d, v_{i+2} d, -h vol / 6), (v_i d, v_i d, 2 h vol / 6)
static inline void flow_term_triplets(int nElem, const int* elem3, const double* vol, double h, Triplets& T)
{
    const int dim = 3;
    for (int id = 0; id < nElem; ++id)
        for (int i = 0; i < dim; ++i)
            for (int d = 0; d < dim; ++d) {
                const int vi = elem3[3 * id + i], v1 = elem3[3 * id + (i + 1) % dim], v2 = elem3[3 * id + (i + 2) % dim];
                T.r.push_back(vi * dim + d); T.c.push_back(v1 * dim + d); T.v.push_back(-h * vol[id] / 6);
                T.r.push_back(vi * dim + d); T.c.push_back(v2 * dim + d); T.v.push_back(-h * vol[id] / 6);
                T.r.push_back(vi * dim + d); T.c.push_back(vi * dim + d); T.v.push_back(2 * h * vol[id] / 6);
            }
}

// `sysMtr.Get_Matrix() += M.Get_Matrix()` with M = diag(m_v) on the vertices that carry mass (INC_POTENTIAL.h:383-386):
// sparse sum = union of the patterns, columns ascending
static inline void csr_add_mass(CSR& A, int nV, const double* m)
{
    CSR B;
    B.ptr.assign(3 * (size_t)nV + 1, 0);
    for (int r = 0; r < 3 * nV; ++r) {
        const double mv = m[r / 3];
        bool placed = (mv == 0.0);
        for (int p = A.ptr[r]; p < A.ptr[r + 1]; ++p) {
            if (!placed && A.col[p] > r) { B.col.push_back(r); B.val.push_back(mv); placed = true; }
            if (!placed && A.col[p] == r) { B.col.push_back(r); B.val.push_back(A.val[p] + mv); placed = true; continue; }
            B.col.push_back(A.col[p]); B.val.push_back(A.val[p]);
        }
        if (!placed) { B.col.push_back(r); B.val.push_back(mv); }
        B.ptr[r + 1] = (int)B.col.size();
    }
    A = std::move(B);
}

// CSR_MATRIX::Project_DBC — CSR_MATRIX.h:130-141
static inline void project_dbc(CSR& A, const uint8_t* DBCb, int dim)
{
    const int n = (int)A.ptr.size() - 1;
    for (int k = 0; k < n; ++k)
        for (int p = A.ptr[k]; p < A.ptr[k + 1]; ++p)
            if (DBCb[k / dim] || DBCb[A.col[p] / dim]) A.val[p] = (k == A.col[p]) ? 1.0 : 0.0;
}

// Find_Surface_Primitives_And_Compute_Area — MESHIO.h:768-834, statement by statement (std::map keyed by the directed pair)
struct SurfacePrimitives {
    std::vector<int> bnode, bedge, btri;
    std::vector<double> BNArea, BEArea, BTArea;
};
static inline void find_surface_primitives(int nV, int nF, const int* tri, const double* X /* nV x 3 */, SurfacePrimitives& S)
{
    std::map<std::pair<int, int>, double> boundaryEdgeSet;
    std::vector<double> isBoundaryNode(nV, 0.0);
    for (int id = 0; id < nF; ++id) {
        const int t0 = tri[3 * id], t1 = tri[3 * id + 1], t2 = tri[3 * id + 2];
        const V3 v0 = ld3(X + 3 * (long)t0), v1 = ld3(X + 3 * (long)t1), v2 = ld3(X + 3 * (long)t2);
        const V3 n = cross(v1 - v0, v2 - v0);
        S.BTArea.push_back(0.5 * std::sqrt((n.x * n.x + n.y * n.y) + n.z * n.z)); // VECTOR::length: p(0) + p(1) + p(2), Math/VECTOR.h:136-151
        S.btri.push_back(t0); S.btri.push_back(t1); S.btri.push_back(t2);
        const int e[3][2] = {{t0, t1}, {t1, t2}, {t2, t0}};
        for (int k = 0; k < 3; ++k) {
            auto finder = boundaryEdgeSet.find({e[k][1], e[k][0]});
            if (finder == boundaryEdgeSet.end()) boundaryEdgeSet[{e[k][0], e[k][1]}] = S.BTArea.back() / 3;
            else finder->second += S.BTArea.back() / 3;
        }
        isBoundaryNode[t0] += S.BTArea.back() / 3;
        isBoundaryNode[t1] += S.BTArea.back() / 3;
        isBoundaryNode[t2] += S.BTArea.back() / 3;
        S.BTArea.back() /= 2;
    }
    for (const auto& i : boundaryEdgeSet) {
        S.bedge.push_back(i.first.first); S.bedge.push_back(i.first.second);
        S.BEArea.push_back(i.second / 2);
    }
    for (int vI = 0; vI < nV; ++vI)
        if (isBoundaryNode[vI]) { S.bnode.push_back(vI); S.BNArea.push_back(isBoundaryNode[vI]); }
}

} // namespace orc
