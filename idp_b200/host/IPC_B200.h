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
    std::uint64_t mesh_hash = 0;
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
    s.check(idp_set_positions(s.ctx, s.xbuf.data(), 3));
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
    s.check(idp_set_rest_positions(s.ctx, s.xbuf.data(), 3));
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
    s.check(idp_set_constraints(s.ctx, (int)constraintSet.size(), rows.data(), info.data()));
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
    dist2.resize(constraintSet.size());
    double m = 0;
    s.check(idp_min_dist2(s.ctx, thickness, dist2.data(), &m));
    minDist2 = m;
}

} // namespace B200
} // namespace JGSL
