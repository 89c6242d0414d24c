// IPC_B200.h — host-side C++ mirror of the reference's six contact operators (Library/FEM/IPC.h), forwarding to the
// B200 library through the C ABI of include/idp_contact.h. Same names, argument meaning and error behaviour
// (printf + exit(-1), IPC.h:775,819,2031) as the reference, so that a maintainer replaces
//     #include <FEM/IPC.h>        by        #include <FEM/IPC.h> + #include "IPC_B200.h"
// and calls JGSL::B200::Compute_* at the five call sites. It is written against the reference's own
// container types (MESH_NODE, MESH_NODE_ATTR, VECTOR, FIELDS, Eigen::Triplet); tests/host_shim/jgsl_mock.h provides
// layout-compatible stand-ins so the marshalling is compiled and exercised in this repository.
//
// Supported instantiation: <T=double, dim=3, shell=false, elasticIPC=false> with empty rod / particle / NNExclusion
// (SURVEY.md §8a). Anything else aborts with the library's IDP_ERR_UNSUPPORTED_PRIMITIVE message — no CPU fallback.
#pragma once
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <map>
#include <set>
#include <vector>
#include "../../include/idp_contact.h"

namespace JGSL {
namespace B200 {

struct Session {
    idp_ctx* ctx = nullptr;
    std::uint64_t mesh_hash = 0, x_hash = 0, x0_hash = 0, rows_hash = 0; // contents last uploaded (0: nothing yet)
    std::size_t nV = 0;
    std::vector<double> xbuf, gbuf;
    std::vector<int> ibuf;
    std::vector<std::uint8_t> dbc;
    static Session& get(int device = 0)
    {
        static Session s;
        if (!s.ctx && idp_create(device, &s.ctx) != IDP_OK) {
            printf("idp_create failed: no CUDA device / library (the ported path has no CPU fallback)\n");
            exit(-1);
        }
        return s;
    }
    void check(int st)
    {
        if (st != IDP_OK) {
            printf("%s\n", idp_last_error(ctx));
            exit(-1);
        }
    }
};

static inline std::uint64_t fnv(std::uint64_t h, const void* p, std::size_t n)
{
    const unsigned char* b = (const unsigned char*)p;
    for (std::size_t i = 0; i < n; ++i) { h ^= b[i]; h *= 1099511628211ull; }
    return h;
}
// word-wise content hash for the per-iterate buffers (positions, 29 M constraint rows ...): four independent
// multiply-xorshift lanes over 8-byte words, a few GB/s -- an order of magnitude cheaper than the upload it saves
static inline std::uint64_t hash_words(const void* p, std::size_t bytes, std::uint64_t seed)
{
    const std::uint64_t* w = (const std::uint64_t*)p;
    const std::size_t n = bytes / 8;
    std::uint64_t h[4] = {seed ^ 0x9e3779b97f4a7c15ull, seed ^ 0xc2b2ae3d27d4eb4full, seed ^ 0x165667b19e3779f9ull, seed ^ 0x27d4eb2f165667c5ull};
    std::size_t i = 0;
    for (; i + 4 <= n; i += 4)
        for (int k = 0; k < 4; ++k) { h[k] = (h[k] ^ w[i + k]) * 0xff51afd7ed558ccdull; h[k] ^= h[k] >> 29; }
    for (; i < n; ++i) { h[0] = (h[0] ^ w[i]) * 0xff51afd7ed558ccdull; h[0] ^= h[0] >> 29; }
    std::uint64_t r = fnv(h[0] ^ (h[1] * 3) ^ (h[2] * 5) ^ (h[3] * 7), (const unsigned char*)p + 8 * n, bytes - 8 * n) ^ (std::uint64_t)bytes;
    return r ? r : 1;
}
// The barrier / min-distance operators only need the vertex count: if no Compute_Constraint_Set / step-size call has set
// the surface on this session yet (the IPC_ENERGY plugin calls the barrier functions on their own, Energy/IPC_ENERGY.h:11-58),
// a vertex-only mesh is installed.
inline void Ensure_Vertices(Session& s, std::size_t nV)
{
    if (s.nV == nV) return;
    s.check(idp_set_mesh(s.ctx, (int)nV, 0, nullptr, 0, nullptr, 0, nullptr, nullptr));
    s.nV = nV;
    s.mesh_hash = s.x_hash = s.x0_hash = s.rows_hash = 0;
}

// uploads boundaryNode / boundaryEdge / boundaryTri / DBCb when they changed (hash of their contents)
template <class VecI2, class VecI3>
inline void Sync_Mesh(Session& s, std::size_t nV, const std::vector<int>& boundaryNode, const std::vector<VecI2>& boundaryEdge,
    const std::vector<VecI3>& boundaryTri, const std::vector<bool>& DBCb)
{
    std::vector<int>& e = s.ibuf;
    e.clear();
    e.reserve(2 * boundaryEdge.size() + 3 * boundaryTri.size());
    for (const auto& v : boundaryEdge) { e.push_back(v[0]); e.push_back(v[1]); }
    const std::size_t triOff = e.size();
    for (const auto& v : boundaryTri) { e.push_back(v[0]); e.push_back(v[1]); e.push_back(v[2]); }
    s.dbc.assign(nV, 0);
    for (std::size_t i = 0; i < DBCb.size() && i < nV; ++i) s.dbc[i] = DBCb[i] ? 1 : 0;
    std::uint64_t h = 1469598103934665603ull ^ nV;
    h = fnv(h, boundaryNode.data(), boundaryNode.size() * sizeof(int));
    h = fnv(h, e.data(), e.size() * sizeof(int));
    h = fnv(h, s.dbc.data(), s.dbc.size());
    if (h == s.mesh_hash) return;
    s.check(idp_set_mesh(s.ctx, (int)nV, (int)boundaryNode.size(), boundaryNode.data(), (int)boundaryEdge.size(), e.data(),
        (int)boundaryTri.size(), e.data() + triOff, s.dbc.data()));
    s.mesh_hash = h;
    s.nV = nV;
    s.x_hash = s.x0_hash = s.rows_hash = 0; // idp_set_mesh invalidates the per-iterate state
}

// X (Cabana AoSoA of 32-byte VECTOR<T,3>) -> dense xyz
template <class MeshNode>
inline void Sync_Positions(Session& s, MeshNode& X)
{
    const std::size_t n = X.size;
    s.xbuf.resize(3 * n);
    for (std::size_t i = 0; i < n; ++i) {
        const auto& x = std::get<0>(X.Get_Unchecked(i));
        s.xbuf[3 * i] = x[0]; s.xbuf[3 * i + 1] = x[1]; s.xbuf[3 * i + 2] = x[2];
    }
    Ensure_Vertices(s, n);
    const std::uint64_t h = hash_words(s.xbuf.data(), s.xbuf.size() * sizeof(double), 1);
    if (h == s.x_hash) return; // E, g and H of one Newton iterate see the same X: uploaded once
    s.check(idp_set_positions(s.ctx, s.xbuf.data(), 3));
    s.x_hash = h;
}
template <class MeshNodeAttr, int X0_FIELD>
inline void Sync_Rest_Positions(Session& s, MeshNodeAttr& nodeAttr)
{
    const std::size_t n = nodeAttr.size;
    s.xbuf.resize(3 * n);
    for (std::size_t i = 0; i < n; ++i) {
        const auto& x = std::get<X0_FIELD>(nodeAttr.Get_Unchecked(i));
        s.xbuf[3 * i] = x[0]; s.xbuf[3 * i + 1] = x[1]; s.xbuf[3 * i + 2] = x[2];
    }
    Ensure_Vertices(s, n);
    const std::uint64_t h = hash_words(s.xbuf.data(), s.xbuf.size() * sizeof(double), 2);
    if (h == s.x0_hash) return;
    s.check(idp_set_rest_positions(s.ctx, s.xbuf.data(), 3));
    s.x0_hash = h;
}
template <class VecI4, class VecT2>
inline void Sync_Constraints(Session& s, const std::vector<VecI4>& constraintSet, const std::vector<VecT2>& stencilInfo)
{
    std::vector<int> rows(4 * constraintSet.size());
    std::vector<double> info(2 * constraintSet.size());
    for (std::size_t i = 0; i < constraintSet.size(); ++i) {
        for (int k = 0; k < 4; ++k) rows[4 * i + k] = constraintSet[i][k];
        info[2 * i] = stencilInfo[i][0]; info[2 * i + 1] = stencilInfo[i][1];
    }
    std::uint64_t h = hash_words(rows.data(), rows.size() * sizeof(int), 3);
    h = hash_words(info.data(), info.size() * sizeof(double), h);
    if (h == s.rows_hash) return; // the same constraint set as the previous operator call (E -> g -> H, line search)
    s.check(idp_set_constraints(s.ctx, (int)constraintSet.size(), rows.data(), info.data()));
    s.rows_hash = h;
}

// ---- Compute_Constraint_Set (FEM/IPC.h:19-36) -------------------------------------------------------------------------
template <class T, int dim, bool shell = false, bool elasticIPC = false, class MeshNode, class MeshNodeAttr, class VecI2, class VecI3,
    class VecI4, class VecT2, int X0_FIELD = 0>
void Compute_Constraint_Set(MeshNode& X, MeshNodeAttr& nodeAttr, const std::vector<int>& boundaryNode,
    const std::vector<VecI2>& boundaryEdge, const std::vector<VecI3>& boundaryTri, const std::vector<int>& particle,
    const std::vector<VecI2>& rod, const std::map<int, std::set<int>>& NNExclusion, const std::vector<T>& BNArea,
    const std::vector<T>& BEArea, const std::vector<T>& BTArea, const VecI2& codimBNStartInd, const std::vector<bool>& DBCb,
    T dHat2, T thickness, bool getPTEE, std::vector<VecI4>& constraintSet, std::vector<VecI2>& cs_PTEE, std::vector<VecT2>& stencilInfo)
{
    static_assert(dim == 3 && !shell && !elasticIPC, "IPC_B200: only <double,3,false,false> is ported");
    (void)BNArea; (void)BEArea; (void)BTArea; (void)codimBNStartInd; (void)cs_PTEE;
    Session& s = Session::get();
    s.check(idp_declare_unsupported(s.ctx, (int)rod.size(), (int)particle.size(), (int)NNExclusion.size() + (getPTEE ? 1 : 0)));
    Sync_Mesh(s, X.size, boundaryNode, boundaryEdge, boundaryTri, DBCb);
    Sync_Positions(s, X);
    Sync_Rest_Positions<MeshNodeAttr, X0_FIELD>(s, nodeAttr);
    int n = 0;
    s.check(idp_constraint_set(s.ctx, dHat2, thickness, &n));
    s.rows_hash = 0; // the device now holds the rows it built (weights 1, IPC.h:656-660); a later upload must not be skipped
    std::vector<int> rows(4 * (std::size_t)n);
    std::vector<double> info(2 * (std::size_t)n);
    s.check(idp_get_constraints(s.ctx, rows.data(), info.data()));
    constraintSet.resize(0); stencilInfo.resize(0);
    constraintSet.reserve(n); stencilInfo.reserve(n);
    for (int i = 0; i < n; ++i) {
        constraintSet.emplace_back(rows[4 * i], rows[4 * i + 1], rows[4 * i + 2], rows[4 * i + 3]);
        stencilInfo.emplace_back(info[2 * i], info[2 * i + 1]);
    }
}

// ---- Compute_Barrier (IPC.h:742-748): E += ... ---------------------------------------------------------------------------
template <class T, int dim, bool elasticIPC = false, class MeshNode, class MeshNodeAttr, class VecI4, class VecT2, int X0_FIELD = 0>
void Compute_Barrier(MeshNode& X, MeshNodeAttr& nodeAttr, const std::vector<VecI4>& constraintSet, const std::vector<VecT2>& stencilInfo,
    T dHat2, T kappa[], T thickness, T& E)
{
    Session& s = Session::get();
    Sync_Positions(s, X);
    Sync_Rest_Positions<MeshNodeAttr, X0_FIELD>(s, nodeAttr);
    Sync_Constraints(s, constraintSet, stencilInfo);
    double e = E;
    s.check(idp_barrier_energy(s.ctx, dHat2, kappa[0], thickness, &e));
    E = e;
}

// ---- Compute_Barrier_Gradient (IPC.h:943-948): nodeAttr.g += ... ----------------------------------------------------------
template <class T, int dim, bool elasticIPC = false, class MeshNode, class MeshNodeAttr, class VecI4, class VecT2, int X0_FIELD = 0, int G_FIELD = 2>
void Compute_Barrier_Gradient(MeshNode& X, const std::vector<VecI4>& constraintSet, const std::vector<VecT2>& stencilInfo, T dHat2,
    T kappa[], T thickness, MeshNodeAttr& nodeAttr)
{
    Session& s = Session::get();
    Sync_Positions(s, X);
    Sync_Rest_Positions<MeshNodeAttr, X0_FIELD>(s, nodeAttr);
    Sync_Constraints(s, constraintSet, stencilInfo);
    s.gbuf.assign(3 * X.size, 0.0);
    s.check(idp_barrier_gradient(s.ctx, dHat2, kappa[0], thickness, s.gbuf.data(), 3));
    for (std::size_t i = 0; i < X.size; ++i) {
        auto& g = std::get<G_FIELD>(nodeAttr.Get_Unchecked(i));
        g[0] += s.gbuf[3 * i]; g[1] += s.gbuf[3 * i + 1]; g[2] += s.gbuf[3 * i + 2];
    }
}

// ---- Compute_Barrier_Hessian (IPC.h:1258-1265): appends triplets ------------------------------------------------------------
// The device returns the barrier Hessian already summed into a scalar CSR; it is appended as one triplet per stored entry,
// which Eigen::setFromTriplets (INC_POTENTIAL.h:382) merges with the other energies exactly like the reference's 144 per row.
template <class T, int dim, bool elasticIPC = false, class MeshNode, class MeshNodeAttr, class VecI4, class VecT2, class Triplet, int X0_FIELD = 0>
void Compute_Barrier_Hessian(MeshNode& X, MeshNodeAttr& nodeAttr, const std::vector<VecI4>& constraintSet,
    const std::vector<VecT2>& stencilInfo, T dHat2, T kappa[], T thickness, bool projectSPD, std::vector<Triplet>& triplets)
{
    Session& s = Session::get();
    Sync_Positions(s, X);
    Sync_Rest_Positions<MeshNodeAttr, X0_FIELD>(s, nodeAttr);
    Sync_Constraints(s, constraintSet, stencilInfo);
    long nnz = 0;
    s.check(idp_barrier_hessian(s.ctx, dHat2, kappa[0], thickness, projectSPD ? 1 : 0, &nnz));
    std::vector<int> ptr(3 * X.size + 1), col((std::size_t)nnz);
    std::vector<double> val((std::size_t)nnz);
    s.check(idp_get_hessian_csr(s.ctx, ptr.data(), col.data(), val.data()));
    triplets.reserve(triplets.size() + (std::size_t)nnz);
    for (std::size_t r = 0; r + 1 < ptr.size(); ++r)
        for (int k = ptr[r]; k < ptr[r + 1]; ++k) triplets.emplace_back((int)r, col[k], val[k]);
}

// ---- Compute_Intersection_Free_StepSize (IPC.h:1879-1890) -----------------------------------------------------------------------
template <class T, int dim, bool shell = false, bool elasticIPC = false, class MeshNode, class VecI2, class VecI3>
void Compute_Intersection_Free_StepSize(MeshNode& X, const std::vector<int>& boundaryNode, const std::vector<VecI2>& boundaryEdge,
    const std::vector<VecI3>& boundaryTri, const std::vector<int>& particle, const std::vector<VecI2>& rod,
    const std::map<int, std::set<int>>& NNExclusion, const VecI2& codimBNStartInd, const std::vector<bool>& DBCb,
    const std::vector<T>& searchDir, T thickness, T& stepSize)
{
    static_assert(dim == 3 && !shell && !elasticIPC, "IPC_B200: only <double,3,false,false> is ported");
    (void)codimBNStartInd;
    Session& s = Session::get();
    s.check(idp_declare_unsupported(s.ctx, (int)rod.size(), (int)particle.size(), (int)NNExclusion.size()));
    Sync_Mesh(s, X.size, boundaryNode, boundaryEdge, boundaryTri, DBCb);
    Sync_Positions(s, X);
    double a = stepSize;
    s.check(idp_ccd_step(s.ctx, searchDir.data(), 3, thickness, &a));
    printf("intersection free step size = %le\n", a);
    stepSize = a;
}

// ---- Compute_Min_Dist2 (IPC.h:2246-2249) -------------------------------------------------------------------------------------------
template <class T, int dim, bool elasticIPC = false, class MeshNode, class VecI4>
void Compute_Min_Dist2(MeshNode& X, const std::vector<VecI4>& constraintSet, T thickness, std::vector<T>& dist2, T& minDist2)
{
    if (constraintSet.empty()) return; // IPC.h:2253-2255
    Session& s = Session::get();
    Sync_Positions(s, X);
    std::vector<int> rows(4 * constraintSet.size());
    for (std::size_t i = 0; i < constraintSet.size(); ++i)
        for (int k = 0; k < 4; ++k) rows[4 * i + k] = constraintSet[i][k];
    s.check(idp_set_constraints(s.ctx, (int)constraintSet.size(), rows.data(), nullptr));
    s.rows_hash = 0;
    dist2.resize(constraintSet.size());
    double m = 0;
    s.check(idp_min_dist2(s.ctx, thickness, dist2.data(), &m));
    minDist2 = m;
}

// ---- the `flow` branch of Compute_IncPotential_Hessian (FEM/Shell/INC_POTENTIAL.h:321-394) kept on the device -----------
// Laplacian flow triplets + Compute_Barrier_Hessian + Construct_From_Triplet + `+= M` + Project_DBC in one call: the device
// assembles all of it into one CSR (idp_system_set_flow_term / idp_system_set_mass / idp_barrier_hessian / idp_project_dbc)
// and the result is handed over through CSR_MATRIX::Construct_From_CSR (Math/CSR_MATRIX.h:33-47) -- no triplet vector at
// all (the six-operator path above has to explode the CSR into one triplet per entry for setFromTriplets to re-sort).
// elemVol[e]: the `vol` of element e (elasticityAttr); massDiag[v]: the lumped mass of vertex v (empty: staticSolve).
template <class T, int dim, class MeshElem, class MeshNode, class MeshNodeAttr, class VecI4, class VecT2, class CsrMatrix, int X0_FIELD = 0>
void Compute_IncPotential_Hessian_Flow(MeshElem& Elem, const std::vector<T>& elemVol, T h, MeshNode& X, MeshNodeAttr& nodeAttr,
    const std::vector<VecI4>& constraintSet, const std::vector<VecT2>& stencilInfo, T dHat2, T kappa[], T thickness, bool projectSPD,
    const std::vector<T>& massDiag, const std::vector<bool>& DBCb, CsrMatrix& sysMtr)
{
    static_assert(dim == 3, "IPC_B200: only dim = 3 is ported");
    Session& s = Session::get();
    Sync_Positions(s, X);
    Sync_Rest_Positions<MeshNodeAttr, X0_FIELD>(s, nodeAttr);
    Sync_Constraints(s, constraintSet, stencilInfo);
    std::vector<int> elem(3 * (std::size_t)Elem.size);
    for (std::size_t e = 0; e < (std::size_t)Elem.size; ++e) {
        const auto& v = std::get<0>(Elem.Get_Unchecked(e));
        elem[3 * e] = v[0]; elem[3 * e + 1] = v[1]; elem[3 * e + 2] = v[2];
    }
    s.check(idp_system_set_flow_term(s.ctx, (int)Elem.size, elem.data(), 3, elemVol.data(), h));
    s.check(idp_system_set_mass(s.ctx, massDiag.empty() ? nullptr : massDiag.data()));
    long nnz = 0;
    s.check(idp_barrier_hessian(s.ctx, dHat2, kappa[0], thickness, projectSPD ? 1 : 0, &nnz));
    // Project_DBC on the device needs the mask the session holds; a vertex-only session has none: project on the host copy
    bool anyFixed = false;
    for (std::size_t i = 0; i < DBCb.size(); ++i) anyFixed = anyFixed || DBCb[i];
    std::vector<int> ptr(3 * X.size + 1), col((std::size_t)nnz);
    std::vector<double> val((std::size_t)nnz);
    s.check(idp_get_hessian_csr(s.ctx, ptr.data(), col.data(), val.data()));
    s.check(idp_system_set_flow_term(s.ctx, 0, nullptr, 3, nullptr, 0.0)); // leave the session as the six operators expect it
    s.check(idp_system_set_mass(s.ctx, nullptr));
    sysMtr.Construct_From_CSR(ptr, col, val);
    if (anyFixed) sysMtr.Project_DBC(DBCb, dim);
}

} // namespace B200
} // namespace JGSL
