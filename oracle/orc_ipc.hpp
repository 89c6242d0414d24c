// ORACLE — TEST INFRASTRUCTURE ONLY (see orc_math.hpp header).
//
// CPU restatement of the six hot-path operators of Library/FEM/IPC.h, instantiated as
// <T=double, dim=3, shell=false, elasticIPC=false> with empty rod / particle / NNExclusion
// (SURVEY.md F8), and of the two SPATIAL_HASH builds they use (Library/Grid/SPATIAL_HASH.h).
// The parallel structure mirrors the reference: `omp parallel for` where the
// reference uses Par_Each, serial where the reference is serial (hash inserts, merge, E, g, min-dist),
// because the same code can be timed as the CPU baseline ("port").
// Checked against the reference's own FEM/IPC.h + Grid/SPATIAL_HASH.h compiled into oracle/_ref/libidp_ref_ipc.so:
// constraint sets, merged-group order, E, per-row distances and CCD steps bit for bit, g / H to 1e-12 (tests/test_ref_loops.py).
#pragma once
#include "orc_deriv.hpp"
#include <vector>
#include <array>
#include <map>
#include <unordered_map>
#include <unordered_set>
#include <numeric>
#include <cstdio>

namespace orc {

struct Mesh {
    int nV = 0;               // number of rows of X (all vertices)
    const double* X = nullptr;   // nV x 3, xyz interleaved
    const double* X0 = nullptr;  // rest positions, nV x 3
    int nBN = 0; const int* bnode = nullptr;   // ascending vertex ids
    int nBE = 0; const int* bedge = nullptr;   // nBE x 2
    int nBT = 0; const int* btri = nullptr;    // nBT x 3
    const uint8_t* dbc = nullptr;              // nV, 1 = Dirichlet
    V3 x(int v) const { return ld3(X + 3 * (long)v); }
    V3 x0(int v) const { return ld3(X0 + 3 * (long)v); }
};

typedef std::array<int, 4> Row;

enum Status { OK = 0, ERR_NONPOSITIVE_DISTANCE = 1, ERR_CCD_ZERO_STEP = 2 };

// ---------------------------------------------------------------------------------------------
// SPATIAL_HASH — Grid/SPATIAL_HASH.h
// ---------------------------------------------------------------------------------------------
struct SpatialHash {
    V3 lo, hi;
    double inv;
    int n[3];
    int n01;
    int edgeStart, triStart;
    std::unordered_map<int, std::vector<int>> voxel;
    std::vector<std::vector<int>> occupancy; // CCD build only (pointAndEdgeOccupancy)

    void axis_index(const V3& p, int idx[3]) const // :704-708
    {
        idx[0] = (int)std::floor((p.x - lo.x) * inv);
        idx[1] = (int)std::floor((p.y - lo.y) * inv);
        idx[2] = (int)std::floor((p.z - lo.z) * inv);
    }
    int lin(const int i[3]) const { return i[0] + i[1] * n[0] + i[2] * n01; } // :693-703

    void set_grid(double voxelSize) // :71-86, 504-519
    {
        const double range[3] = {hi.x - lo.x, hi.y - lo.y, hi.z - lo.z};
        inv = 1.0 / voxelSize;
        long amt = 1;
        for (int d = 0; d < 3; ++d) amt *= std::max(1L, (long)std::ceil(range[d] * inv));
        if (amt > 1e9) {
            voxelSize *= std::pow(amt / 1.0e9, 1.0 / 3);
            inv = 1.0 / voxelSize;
        }
        int mn = std::numeric_limits<int>::max();
        for (int d = 0; d < 3; ++d) {
            n[d] = std::max(1, (int)std::ceil(range[d] * inv));
            mn = std::min(mn, n[d]);
        }
        if (mn <= 0) {
            inv = 1.0 / (std::max(std::max(range[0], range[1]), range[2]) * 1.01);
            n[0] = n[1] = n[2] = 1;
        }
        n01 = n[0] * n[1];
    }

    static double mean_edge_length(const Mesh& m)
    {
        std::vector<double> eLen(m.nBE);
#pragma omp parallel for
        for (int e = 0; e < m.nBE; ++e)
            eLen[e] = std::sqrt(sqn(m.x(m.bedge[2 * e]) - m.x(m.bedge[2 * e + 1])));
        return tree_sum(eLen.data(), m.nBE) / m.nBE;
    }

    template <class F>
    static void for_cells(const int mins[3], const int maxs[3], int n0, int n01, F f)
    {
        for (int iz = mins[2]; iz <= maxs[2]; ++iz)
            for (int iy = mins[1]; iy <= maxs[1]; ++iy)
                for (int ix = mins[0]; ix <= maxs[0]; ++ix) f(ix + iy * n0 + iz * n01);
    }

    // static build — :28-213
    void build(const Mesh& m, double voxelSize)
    {
        if (m.nBE) voxelSize *= mean_edge_length(m);
        lo = hi = m.x(0);
        for (int v = 1; v < m.nV; ++v) { lo = vmin(lo, m.x(v)); hi = vmax(hi, m.x(v)); }
        set_grid(voxelSize);
        edgeStart = m.nBN;
        triStart = edgeStart + m.nBE;

        std::vector<std::array<int, 3>> svIdx(m.nBN);
        std::vector<int> vI2SVI(m.nV, -1);
#pragma omp parallel for
        for (int s = 0; s < m.nBN; ++s) {
            axis_index(m.x(m.bnode[s]), svIdx[s].data());
            vI2SVI[m.bnode[s]] = s;
        }
        voxel.clear();
        for (int s = 0; s < m.nBN; ++s) voxel[lin(svIdx[s].data())].push_back(s);

        std::vector<std::vector<int>> locE(m.nBE), locT(m.nBT);
#pragma omp parallel for
        for (int e = 0; e < m.nBE; ++e) {
            const auto& a = svIdx[vI2SVI[m.bedge[2 * e]]];
            const auto& b = svIdx[vI2SVI[m.bedge[2 * e + 1]]];
            int mins[3], maxs[3];
            for (int d = 0; d < 3; ++d) { mins[d] = std::min(a[d], b[d]); maxs[d] = std::max(a[d], b[d]); }
            for_cells(mins, maxs, n[0], n01, [&](int c) { locE[e].push_back(c); });
        }
#pragma omp parallel for
        for (int t = 0; t < m.nBT; ++t) {
            const auto& a = svIdx[vI2SVI[m.btri[3 * t]]];
            const auto& b = svIdx[vI2SVI[m.btri[3 * t + 1]]];
            const auto& c = svIdx[vI2SVI[m.btri[3 * t + 2]]];
            int mins[3], maxs[3];
            for (int d = 0; d < 3; ++d) {
                mins[d] = std::min(std::min(a[d], b[d]), c[d]);
                maxs[d] = std::max(std::max(a[d], b[d]), c[d]);
            }
            for_cells(mins, maxs, n[0], n01, [&](int cI) { locT[t].push_back(cI); });
        }
        for (int e = 0; e < m.nBE; ++e)
            for (int c : locE[e]) voxel[c].push_back(e + edgeStart);
        for (int t = 0; t < m.nBT; ++t)
            for (int c : locT[t]) voxel[c].push_back(t + triStart);
    }

    void clamp_range(const V3& a, const V3& b, int mins[3], int maxs[3]) const
    {
        axis_index(a, mins);
        axis_index(b, maxs);
        for (int d = 0; d < 3; ++d) { mins[d] = std::max(mins[d], 0); maxs[d] = std::min(maxs[d], n[d] - 1); }
    }

    // :215-241
    void query_point_for_triangles(const V3& p, double radius, std::unordered_set<int>& out) const
    {
        int mins[3], maxs[3];
        clamp_range({p.x - radius, p.y - radius, p.z - radius}, {p.x + radius, p.y + radius, p.z + radius}, mins, maxs);
        out.clear();
        for_cells(mins, maxs, n[0], n01, [&](int c) {
            auto it = voxel.find(c);
            if (it != voxel.end())
                for (int id : it->second)
                    if (id >= triStart) out.insert(id - triStart);
        });
    }

    // :243-291
    void query_edge_for_edges(const V3& a, const V3& b, double radius, std::vector<int>& out, int eIq) const
    {
        const V3 mn = vmin(a, b), mx = vmax(a, b);
        int mins[3], maxs[3];
        clamp_range({mn.x - radius, mn.y - radius, mn.z - radius}, {mx.x + radius, mx.y + radius, mx.z + radius}, mins, maxs);
        out.resize(0);
        for_cells(mins, maxs, n[0], n01, [&](int c) {
            auto it = voxel.find(c);
            if (it != voxel.end())
                for (int id : it->second)
                    if (id >= edgeStart && id < triStart && id - edgeStart > eIq) out.push_back(id - edgeStart);
        });
        std::sort(out.begin(), out.end());
        out.erase(std::unique(out.begin(), out.end()), out.end());
    }

    // CCD build — :432-622. Mutates step (span clamp, F7).
    void build_ccd(const Mesh& m, const double* dir, double& step, double voxelSize, double thickness)
    {
        if (m.nBE) voxelSize *= mean_edge_length(m);
        double pSize = 0;
        for (int s = 0; s < m.nBN; ++s) {
            const int v = m.bnode[s];
            pSize += std::fabs(dir[3 * (long)v]);
            pSize += std::fabs(dir[3 * (long)v + 1]);
            pSize += std::fabs(dir[3 * (long)v + 2]);
        }
        pSize /= m.nBN * 3;
        const double spanSize = step * pSize / voxelSize;
        if (spanSize > 1) step /= spanSize;

        std::vector<V3> SV(m.nBN), SVt(m.nBN);
        std::vector<int> vI2SVI(m.nV, -1);
        for (int s = 0; s < m.nBN; ++s) {
            const int v = m.bnode[s];
            vI2SVI[v] = s;
            SV[s] = m.x(v);
            SVt[s] = {SV[s].x + step * dir[3 * (long)v], SV[s].y + step * dir[3 * (long)v + 1], SV[s].z + step * dir[3 * (long)v + 2]};
        }
        V3 mn = vmin(SV[0], SVt[0]), mx = vmax(SV[0], SVt[0]);
        {
            // colwise min of SV and of SVt taken separately, then combined (:502-503); min/max are exact
            for (int s = 1; s < m.nBN; ++s) { mn = vmin(mn, vmin(SV[s], SVt[s])); mx = vmax(mx, vmax(SV[s], SVt[s])); }
        }
        const double half = thickness / 2;
        lo = {mn.x - half, mn.y - half, mn.z - half};
        hi = {mx.x + half, mx.y + half, mx.z + half};
        set_grid(voxelSize);
        edgeStart = m.nBN;
        triStart = edgeStart + m.nBE;

        std::vector<std::array<int, 3>> svMin(m.nBN), svMax(m.nBN);
#pragma omp parallel for
        for (int s = 0; s < m.nBN; ++s) {
            const V3 a = vmin(SV[s], SVt[s]), b = vmax(SV[s], SVt[s]);
            axis_index({a.x - half, a.y - half, a.z - half}, svMin[s].data());
            axis_index({b.x + half, b.y + half, b.z + half}, svMax[s].data());
        }
        voxel.clear();
        occupancy.assign(triStart, {});
#pragma omp parallel for
        for (int s = 0; s < m.nBN; ++s)
            for_cells(svMin[s].data(), svMax[s].data(), n[0], n01, [&](int c) { occupancy[s].push_back(c); });
#pragma omp parallel for
        for (int e = 0; e < m.nBE; ++e) {
            const int s0 = vI2SVI[m.bedge[2 * e]], s1 = vI2SVI[m.bedge[2 * e + 1]];
            int mins[3], maxs[3];
            for (int d = 0; d < 3; ++d) { mins[d] = std::min(svMin[s0][d], svMin[s1][d]); maxs[d] = std::max(svMax[s0][d], svMax[s1][d]); }
            for_cells(mins, maxs, n[0], n01, [&](int c) { occupancy[e + edgeStart].push_back(c); });
        }
        std::vector<std::vector<int>> locT(m.nBT);
#pragma omp parallel for
        for (int t = 0; t < m.nBT; ++t) {
            const int s0 = vI2SVI[m.btri[3 * t]], s1 = vI2SVI[m.btri[3 * t + 1]], s2 = vI2SVI[m.btri[3 * t + 2]];
            int mins[3], maxs[3];
            for (int d = 0; d < 3; ++d) {
                mins[d] = std::min(std::min(svMin[s0][d], svMin[s1][d]), svMin[s2][d]);
                maxs[d] = std::max(std::max(svMax[s0][d], svMax[s1][d]), svMax[s2][d]);
            }
            for_cells(mins, maxs, n[0], n01, [&](int c) { locT[t].push_back(c); });
        }
        for (int i = 0; i < (int)occupancy.size(); ++i)
            for (int c : occupancy[i]) voxel[c].push_back(i);
        for (int t = 0; t < m.nBT; ++t)
            for (int c : locT[t]) voxel[c].push_back(t + triStart);
    }

    // :624-647 (triangles only: the point/edge sets are unused on the codimensional-surface path)
    void query_point_for_triangles_ccd(int svI, std::unordered_set<int>& tris) const
    {
        tris.clear();
        for (int c : occupancy[svI]) {
            auto it = voxel.find(c);
            for (int id : it->second)
                if (id >= triStart) tris.insert(id - triStart);
        }
    }
    // :650-662
    void query_edge_for_edges_ccd(int seI, std::unordered_set<int>& edges) const
    {
        edges.clear();
        for (int c : occupancy[seI + edgeStart]) {
            auto it = voxel.find(c);
            for (int id : it->second)
                if (id >= edgeStart && id < triStart && id - edgeStart > seI) edges.insert(id - edgeStart);
        }
    }
};

// ---------------------------------------------------------------------------------------------
// Compute_Constraint_Set — FEM/IPC.h:19-740 (3-D branch :143-661)
// `brute` selects the reference's all-pairs `#else` branches (:166-168, 380-382).
// candPT / candEE (optional) receive the pairs that reach the distance-type switch (:198, :434),
// i.e. the post-AABB candidate set of SURVEY.md A.2.
// ---------------------------------------------------------------------------------------------
struct ConstraintSetResult {
    std::vector<Row> rows;
    std::vector<std::array<double, 2>> info; // weight, dHat2
    std::vector<std::array<int, 2>> candPT, candEE;
};

static inline bool tri_excluded(const Mesh& m, int vI, const int* t)
{
    if (vI == t[0] || vI == t[1] || vI == t[2]) return true;
    if (m.dbc && m.dbc[vI] && m.dbc[t[0]] && m.dbc[t[1]] && m.dbc[t[2]]) return true;
    return false;
}
static inline bool edge_excluded(const Mesh& m, const int* a, const int* b, int eI, int eJ)
{
    if (a[0] == b[0] || a[0] == b[1] || a[1] == b[0] || a[1] == b[1] || eI > eJ) return true;
    if (m.dbc && m.dbc[a[0]] && m.dbc[a[1]] && m.dbc[b[0]] && m.dbc[b[1]]) return true;
    return false;
}

static inline void compute_constraint_set(const Mesh& m, double dHat2, double thickness, bool brute, bool wantCand,
    ConstraintSetResult& R)
{
    SpatialHash sh;
    if (!brute) sh.build(m, 1.0);
    const double dHat = std::sqrt(dHat2) + thickness;
    dHat2 = dHat * dHat;

    std::vector<std::vector<Row>> csPT(m.nBN), csEE(m.nBE);
    std::vector<std::vector<std::array<int, 2>>> cPT(wantCand ? m.nBN : 0), cEE(wantCand ? m.nBE : 0);

#pragma omp parallel for schedule(dynamic, 64)
    for (int svI = 0; svI < m.nBN; ++svI) {
        const int vI = m.bnode[svI];
        const V3 p = m.x(vI);
        std::unordered_set<int> tris;
        if (!brute) sh.query_point_for_triangles(p, dHat, tris);
        auto body = [&](int sfI) {
            const int* t = m.btri + 3 * (long)sfI;
            if (tri_excluded(m, vI, t)) return;
            const V3 t0 = m.x(t[0]), t1 = m.x(t[1]), t2 = m.x(t[2]);
            if (!pt_cd_broadphase(p, t0, t1, t2, dHat)) return;
            if (wantCand) cPT[svI].push_back({svI, sfI});
            double d;
            Row r;
            switch (pt_type(p, t0, t1, t2)) {
            case 0: d = dist2_pp(p, t0); r = {-vI - 1, t[0], -1, -1}; break;
            case 1: d = dist2_pp(p, t1); r = {-vI - 1, t[1], -1, -1}; break;
            case 2: d = dist2_pp(p, t2); r = {-vI - 1, t[2], -1, -1}; break;
            case 3: d = dist2_pe(p, t0, t1); r = {-vI - 1, t[0], t[1], -1}; break;
            case 4: d = dist2_pe(p, t1, t2); r = {-vI - 1, t[1], t[2], -1}; break;
            case 5: d = dist2_pe(p, t2, t0); r = {-vI - 1, t[2], t[0], -1}; break;
            default: d = dist2_pt(p, t0, t1, t2); r = {-vI - 1, t[0], t[1], t[2]}; break;
            }
            if (d < dHat2) csPT[svI].push_back(r);
        };
        if (brute) for (int sfI = 0; sfI < m.nBT; ++sfI) body(sfI);
        else for (int sfI : tris) body(sfI);
    }

#pragma omp parallel for schedule(dynamic, 64)
    for (int eI = 0; eI < m.nBE; ++eI) {
        const int* a = m.bedge + 2 * (long)eI;
        const V3 ea0 = m.x(a[0]), ea1 = m.x(a[1]);
        std::vector<int> partners;
        if (!brute) sh.query_edge_for_edges(ea0, ea1, dHat, partners, eI);
        auto body = [&](int eJ) {
            const int* b = m.bedge + 2 * (long)eJ;
            if (edge_excluded(m, a, b, eI, eJ)) return;
            const V3 eb0 = m.x(b[0]), eb1 = m.x(b[1]);
            if (!ee_cd_broadphase(ea0, ea1, eb0, eb1, dHat)) return;
            if (wantCand) cEE[eI].push_back({eI, eJ});
            const double c = ee_cross_norm2(ea0, ea1, eb0, eb1);
            const double eps_x = ee_mollifier_threshold(m.x0(a[0]), m.x0(a[1]), m.x0(b[0]), m.x0(b[1]));
            const bool moll = c < eps_x;
            double d;
            Row r;
            switch (ee_type(ea0, ea1, eb0, eb1)) {
            case 0: d = dist2_pp(ea0, eb0); r = moll ? Row{a[0], b[0], -a[1] - 1, -b[1] - 1} : Row{-a[0] - 1, b[0], -1, -1}; break;
            case 1: d = dist2_pp(ea0, eb1); r = moll ? Row{a[0], b[1], -a[1] - 1, -b[0] - 1} : Row{-a[0] - 1, b[1], -1, -1}; break;
            case 2: d = dist2_pe(ea0, eb0, eb1); r = moll ? Row{a[0], b[0], b[1], -a[1] - 1} : Row{-a[0] - 1, b[0], b[1], -1}; break;
            case 3: d = dist2_pp(ea1, eb0); r = moll ? Row{a[1], b[0], -a[0] - 1, -b[1] - 1} : Row{-a[1] - 1, b[0], -1, -1}; break;
            case 4: d = dist2_pp(ea1, eb1); r = moll ? Row{a[1], b[1], -a[0] - 1, -b[0] - 1} : Row{-a[1] - 1, b[1], -1, -1}; break;
            case 5: d = dist2_pe(ea1, eb0, eb1); r = moll ? Row{a[1], b[0], b[1], -a[0] - 1} : Row{-a[1] - 1, b[0], b[1], -1}; break;
            case 6: d = dist2_pe(eb0, ea0, ea1); r = moll ? Row{b[0], a[0], a[1], -b[1] - 1} : Row{-b[0] - 1, a[0], a[1], -1}; break;
            case 7: d = dist2_pe(eb1, ea0, ea1); r = moll ? Row{b[1], a[0], a[1], -b[0] - 1} : Row{-b[1] - 1, a[0], a[1], -1}; break;
            default: d = dist2_ee(ea0, ea1, eb0, eb1); r = moll ? Row{a[0], a[1], -b[0] - 1, b[1]} : Row{a[0], a[1], b[0], b[1]}; break;
            }
            if (d < dHat2) csEE[eI].push_back(r);
        };
        if (brute) for (int eJ = eI + 1; eJ < m.nBE; ++eJ) body(eJ);
        else for (int eJ : partners) body(eJ);
    }

    // merge — :571-661 (serial); OIPC: all weights 1 (:656-660)
    R.rows.clear();
    R.info.clear();
    std::map<Row, int> counter;
    for (const auto& cs : csPT)
        for (const auto& r : cs) {
            if (r[3] < 0) ++counter[r];
            else R.rows.push_back(r);
        }
    for (const auto& cs : csEE)
        for (const auto& r : cs) {
            if (r[0] < 0) ++counter[r];
            else R.rows.push_back(r);
        }
    for (const auto& kv : counter) R.rows.push_back({kv.first[0], kv.first[1], kv.first[2], -kv.second});
    R.info.assign(R.rows.size(), {1.0, dHat2});
    if (wantCand) {
        R.candPT.clear(); R.candEE.clear();
        for (auto& v : cPT) R.candPT.insert(R.candPT.end(), v.begin(), v.end());
        for (auto& v : cEE) R.candEE.insert(R.candEE.end(), v.begin(), v.end());
        std::sort(R.candPT.begin(), R.candPT.end());
        std::sort(R.candEE.begin(), R.candEE.end());
    }
}

// ---------------------------------------------------------------------------------------------
// constraint-row decoding (IPC.h:802-936 and the three sibling decoders)
// ---------------------------------------------------------------------------------------------
enum Kind { K_EE = 0, K_EE_M = 1, K_PE_M = 2, K_PP_M = 3, K_PT = 4, K_PE = 5, K_PP = 6 };
struct Decoded {
    Kind kind;
    int v[4];   // stencil vertices in g/H DOF order
    int nv;     // 4, 3 or 2
    int mult;   // multiplicity (>=1)
};
static inline Decoded decode(const Row& r)
{
    Decoded d;
    d.mult = 1;
    if (r[0] >= 0) {
        d.nv = 4;
        if (r[3] >= 0 && r[2] >= 0) { d.kind = K_EE; d.v[0] = r[0]; d.v[1] = r[1]; d.v[2] = r[2]; d.v[3] = r[3]; }
        else if (r[3] >= 0) { d.kind = K_EE_M; d.v[0] = r[0]; d.v[1] = r[1]; d.v[2] = -r[2] - 1; d.v[3] = r[3]; }
        else if (r[2] >= 0) { d.kind = K_PE_M; d.v[0] = r[0]; d.v[1] = -r[3] - 1; d.v[2] = r[1]; d.v[3] = r[2]; }
        else { d.kind = K_PP_M; d.v[0] = r[0]; d.v[1] = -r[2] - 1; d.v[2] = r[1]; d.v[3] = -r[3] - 1; }
    }
    else {
        d.v[0] = -r[0] - 1; d.v[1] = r[1]; d.v[2] = r[2]; d.v[3] = r[3];
        if (r[3] >= 0) { d.kind = K_PT; d.nv = 4; }
        else if (r[2] >= 0) { d.kind = K_PE; d.nv = 3; d.mult = -r[3]; }
        else { d.kind = K_PP; d.nv = 2; d.mult = -r[3]; }
    }
    return d;
}
static inline double row_dist2(const Mesh& m, const Decoded& d)
{
    switch (d.kind) {
    case K_EE: case K_EE_M: return dist2_ee(m.x(d.v[0]), m.x(d.v[1]), m.x(d.v[2]), m.x(d.v[3]));
    case K_PE_M: return dist2_pe(m.x(d.v[0]), m.x(d.v[2]), m.x(d.v[3]));
    case K_PP_M: return dist2_pp(m.x(d.v[0]), m.x(d.v[2]));
    case K_PT: return dist2_pt(m.x(d.v[0]), m.x(d.v[1]), m.x(d.v[2]), m.x(d.v[3]));
    case K_PE: return dist2_pe(m.x(d.v[0]), m.x(d.v[1]), m.x(d.v[2]));
    default: return dist2_pp(m.x(d.v[0]), m.x(d.v[1]));
    }
}

// local (per-row) energy, gradient and Hessian over the decoded stencil. n = 3*nv DOFs.
// Implements IPC.h:801-938 (E), 1012-1254 (g), 1390-1729 (H) for one row.
static inline int row_EgH(const Mesh& m, const Row& row, double weight, double dHat2, double kappa, double thickness2,
    bool projectSPD, double* E, double* g, double* H, Decoded* dec_out = nullptr)
{
    const Decoded d = decode(row);
    if (dec_out) *dec_out = d;
    const int n = 3 * d.nv;
    const double dist2 = row_dist2(m, d) - thickness2;
    if (dist2 <= 0) return ERR_NONPOSITIVE_DISTANCE;
    const double b = barrier(dist2, dHat2, kappa);
    const double bg = barrier_g(dist2, dHat2, kappa), bh = barrier_h(dist2, dHat2, kappa);
    const bool wantD = (g || H);
    double dg[12], dH[144];
    const bool moll = (d.kind == K_EE_M || d.kind == K_PE_M || d.kind == K_PP_M);
    if (!moll) {
        if (E) *E = b * (d.mult > 1 ? (double)d.mult : 1.0) * weight;
        if (!wantD) return OK;
        switch (d.kind) {
        case K_EE: ee_grad_hess(m.x(d.v[0]), m.x(d.v[1]), m.x(d.v[2]), m.x(d.v[3]), dg, dH); break;
        case K_PT: pt_grad_hess(m.x(d.v[0]), m.x(d.v[1]), m.x(d.v[2]), m.x(d.v[3]), dg, dH); break;
        case K_PE: pe_grad_hess(m.x(d.v[0]), m.x(d.v[1]), m.x(d.v[2]), dg, dH); break;
        default: pp_grad_hess(m.x(d.v[0]), m.x(d.v[1]), dg, dH); break;
        }
        const double mu = (double)d.mult;
        if (g) for (int i = 0; i < n; ++i) g[i] = dg[i] * (mu * weight * bg);
        if (H) {
            for (int i = 0; i < n; ++i)
                for (int j = 0; j < n; ++j) H[i * n + j] = (((mu * bh) * dg[i]) * dg[j] + (mu * bg) * dH[i * n + j]) * weight;
            if (projectSPD) make_pd(n, H);
        }
        return OK;
    }
    // mollified rows: 12 DOFs (ea0, ea1, eb0, eb1)
    const V3 a0 = m.x(d.v[0]), a1 = m.x(d.v[1]), b0 = m.x(d.v[2]), b1 = m.x(d.v[3]);
    const double eps_x = ee_mollifier_threshold(m.x0(d.v[0]), m.x0(d.v[1]), m.x0(d.v[2]), m.x0(d.v[3]));
    double e, ge[12], He[144];
    ee_mollifier_all(a0, a1, b0, b1, eps_x, e, ge, He);
    if (E) *E = b * e * weight;
    if (!wantD) return OK;
    // distance derivatives on the sub-stencil, scattered into the 12 DOFs (P of SURVEY.md C.3)
    int map[12], nd;
    if (d.kind == K_EE_M) { nd = 12; ee_grad_hess(a0, a1, b0, b1, dg, dH); for (int i = 0; i < 12; ++i) map[i] = i; }
    else if (d.kind == K_PE_M) {
        nd = 9; pe_grad_hess(a0, b0, b1, dg, dH);
        for (int i = 0; i < 3; ++i) { map[i] = i; map[3 + i] = 6 + i; map[6 + i] = 9 + i; }
    }
    else {
        nd = 6; pp_grad_hess(a0, b0, dg, dH);
        for (int i = 0; i < 3; ++i) { map[i] = i; map[3 + i] = 6 + i; }
    }
    double Pg[12] = {0};
    for (int i = 0; i < nd; ++i) Pg[map[i]] = dg[i];
    if (g) for (int i = 0; i < 12; ++i) g[i] = weight * ((e * bg) * Pg[i] + b * ge[i]);
    if (H) {
        for (int i = 0; i < 144; ++i) H[i] = b * He[i];
        for (int i = 0; i < nd; ++i)
            for (int j = 0; j < nd; ++j)
                H[map[i] * 12 + map[j]] += ((e * bh) * dg[i]) * dg[j] + (e * bg) * dH[i * nd + j];
        for (int i = 0; i < 12; ++i)
            for (int j = 0; j < 12; ++j) H[i * 12 + j] += (bg * Pg[i]) * ge[j] + (bg * Pg[j]) * ge[i];
        for (int i = 0; i < 144; ++i) H[i] *= weight;
        if (projectSPD) make_pd(12, H);
    }
    return OK;
}

static inline double adjusted_dhat2(double dHat2, double thickness) { return dHat2 + 2 * std::sqrt(dHat2) * thickness; } // :757

// Compute_Barrier — IPC.h:742-941 (adds to E; serial loop, std::accumulate in row order)
static inline int compute_barrier(const Mesh& m, const std::vector<Row>& rows, const double* weight, double dHat2,
    double kappa, double thickness, double& E)
{
    const double t2 = thickness * thickness;
    dHat2 = adjusted_dhat2(dHat2, thickness);
    std::vector<double> b(rows.size());
    for (size_t c = 0; c < rows.size(); ++c) {
        const int st = row_EgH(m, rows[c], weight[c], dHat2, kappa, t2, false, &b[c], nullptr, nullptr);
        if (st) return st;
    }
    E += std::accumulate(b.begin(), b.end(), 0.0);
    return OK;
}

// Compute_Barrier_Gradient — IPC.h:943-1256 (adds into g[3*nV]; serial)
static inline int compute_barrier_gradient(const Mesh& m, const std::vector<Row>& rows, const double* weight, double dHat2,
    double kappa, double thickness, double* gOut)
{
    const double t2 = thickness * thickness;
    dHat2 = adjusted_dhat2(dHat2, thickness);
    for (size_t c = 0; c < rows.size(); ++c) {
        double g[12];
        Decoded d;
        const int st = row_EgH(m, rows[c], weight[c], dHat2, kappa, t2, false, nullptr, g, nullptr, &d);
        if (st) return st;
        for (int i = 0; i < d.nv; ++i)
            for (int a = 0; a < 3; ++a) gOut[3 * (long)d.v[i] + a] += g[3 * i + a];
    }
    return OK;
}

// Compute_Barrier_Hessian — IPC.h:1258-1731: appends 144/81/36 triplets per row, row-major within the block
struct Triplets {
    std::vector<int> r, c;
    std::vector<double> v;
};
static inline int compute_barrier_hessian(const Mesh& m, const std::vector<Row>& rows, const double* weight, double dHat2,
    double kappa, double thickness, bool projectSPD, Triplets& T)
{
    const double t2 = thickness * thickness;
    dHat2 = adjusted_dhat2(dHat2, thickness);
    std::vector<size_t> start(rows.size());
    size_t cur = T.v.size();
    for (size_t c = 0; c < rows.size(); ++c) {
        start[c] = cur;
        const Row& r = rows[c];
        cur += (r[0] >= 0 || r[3] >= 0) ? 144 : (r[2] >= 0 ? 81 : 36);
    }
    T.r.resize(cur); T.c.resize(cur); T.v.resize(cur);
    int status = OK;
#pragma omp parallel for schedule(dynamic, 64)
    for (long c = 0; c < (long)rows.size(); ++c) {
        double H[144];
        Decoded d;
        const int st = row_EgH(m, rows[c], weight[c], dHat2, kappa, t2, projectSPD, nullptr, nullptr, H, &d);
        if (st) {
#pragma omp critical
            status = st;
            continue;
        }
        const int n = 3 * d.nv;
        size_t o = start[c];
        for (int i = 0; i < d.nv; ++i)
            for (int a = 0; a < 3; ++a)
                for (int j = 0; j < d.nv; ++j)
                    for (int bq = 0; bq < 3; ++bq) {
                        const size_t k = o + (size_t)(i * 3 + a) * n + j * 3 + bq;
                        T.r[k] = d.v[i] * 3 + a;
                        T.c[k] = d.v[j] * 3 + bq;
                        T.v[k] = H[(i * 3 + a) * n + j * 3 + bq];
                    }
    }
    return status;
}

// CSR_MATRIX::Construct_From_Triplet — Math/CSR_MATRIX.h:49-56 (Eigen setFromTriplets semantics:
// duplicates summed in insertion order, columns sorted per row, explicit zeros kept)
struct CSR {
    std::vector<int> ptr, col;
    std::vector<double> val;
};
static inline void csr_from_triplets(int nrows, const Triplets& T, CSR& A)
{
    std::vector<std::vector<std::pair<int, size_t>>> perRow(nrows);
    for (size_t k = 0; k < T.v.size(); ++k) perRow[T.r[k]].push_back({T.c[k], k});
    A.ptr.assign(nrows + 1, 0);
    A.col.clear(); A.val.clear();
    for (int r = 0; r < nrows; ++r) {
        auto& e = perRow[r];
        std::stable_sort(e.begin(), e.end(), [](const auto& a, const auto& b) { return a.first < b.first; });
        for (size_t k = 0; k < e.size(); ++k) {
            if (k > 0 && e[k].first == e[k - 1].first) A.val.back() += T.v[e[k].second];
            else { A.col.push_back(e[k].first); A.val.push_back(T.v[e[k].second]); }
        }
        A.ptr[r + 1] = (int)A.col.size();
    }
}

// Compute_Min_Dist2 — IPC.h:2246-2388
static inline void compute_min_dist2(const Mesh& m, const std::vector<Row>& rows, double thickness, std::vector<double>& dist2, double& minDist2)
{
    if (rows.empty()) return;
    dist2.resize(rows.size());
    for (size_t c = 0; c < rows.size(); ++c) dist2[c] = row_dist2(m, decode(rows[c]));
    minDist2 = *std::min_element(dist2.begin(), dist2.end());
    minDist2 -= thickness * thickness;
}

// ---------------------------------------------------------------------------------------------
// Compute_Intersection_Free_StepSize — IPC.h:1879-2244 (3-D branch :1957-2243)
// ---------------------------------------------------------------------------------------------
struct CCDResult {
    double step;            // in: initial step; out: filtered step
    double step_after_clamp;
    long iters = 0;         // total ACCD loop trips
    std::vector<std::array<int, 2>> candPT, candEE; // pairs reaching the ACCD call (:2009, :2233), sorted
};
static inline int compute_intersection_free_stepsize(const Mesh& m, const double* dir, double thickness, bool brute,
    bool wantCand, CCDResult& R)
{
    const double eta = 0.1;
    double stepSize = R.step;
    SpatialHash sh;
    if (!brute) sh.build_ccd(m, dir, stepSize, 1.0, thickness);
    R.step_after_clamp = stepSize;
    int status = OK;
    long iters = 0;
    std::vector<std::vector<std::array<int, 2>>> cPT(wantCand ? m.nBN : 0), cEE(wantCand ? m.nBE : 0);

    std::vector<double> alphaPT(m.nBN);
#pragma omp parallel for schedule(dynamic, 64) reduction(+ : iters)
    for (int svI = 0; svI < m.nBN; ++svI) {
        const int vI = m.bnode[svI];
        const V3 p = m.x(vI), dp = ld3(dir + 3 * (long)vI);
        alphaPT[svI] = stepSize;
        std::unordered_set<int> tris;
        if (!brute) sh.query_point_for_triangles_ccd(svI, tris);
        auto body = [&](int sfI) {
            const int* t = m.btri + 3 * (long)sfI;
            if (tri_excluded(m, vI, t)) return;
            const V3 t0 = m.x(t[0]), t1 = m.x(t[1]), t2 = m.x(t[2]);
            const V3 dt0 = ld3(dir + 3 * (long)t[0]), dt1 = ld3(dir + 3 * (long)t[1]), dt2 = ld3(dir + 3 * (long)t[2]);
            if (!pt_ccd_broadphase(p, t0, t1, t2, dp, dt0, dt1, dt2, thickness)) return;
            if (wantCand) cPT[svI].push_back({svI, sfI});
            double a = alphaPT[svI];
            if (accd_pt(p, t0, t1, t2, dp, dt0, dt1, dt2, eta, thickness, a, &iters)) {
                if (alphaPT[svI] > a) alphaPT[svI] = a;
            }
            if (a == 0) {
#pragma omp critical
                status = ERR_CCD_ZERO_STEP;
            }
        };
        if (brute) for (int sfI = 0; sfI < m.nBT; ++sfI) body(sfI);
        else for (int sfI : tris) body(sfI);
    }
    if (m.nBN) stepSize = std::min(stepSize, *std::min_element(alphaPT.begin(), alphaPT.end()));

    std::vector<double> alphaEE(m.nBE);
#pragma omp parallel for schedule(dynamic, 64) reduction(+ : iters)
    for (int eI = 0; eI < m.nBE; ++eI) {
        const int* a = m.bedge + 2 * (long)eI;
        const V3 ea0 = m.x(a[0]), ea1 = m.x(a[1]);
        const V3 dea0 = ld3(dir + 3 * (long)a[0]), dea1 = ld3(dir + 3 * (long)a[1]);
        alphaEE[eI] = stepSize;
        std::unordered_set<int> partners;
        if (!brute) sh.query_edge_for_edges_ccd(eI, partners);
        auto body = [&](int eJ) {
            const int* b = m.bedge + 2 * (long)eJ;
            if (edge_excluded(m, a, b, eI, eJ)) return;
            const V3 eb0 = m.x(b[0]), eb1 = m.x(b[1]);
            const V3 deb0 = ld3(dir + 3 * (long)b[0]), deb1 = ld3(dir + 3 * (long)b[1]);
            if (!ee_ccd_broadphase(ea0, ea1, eb0, eb1, dea0, dea1, deb0, deb1, thickness)) return;
            if (wantCand) cEE[eI].push_back({eI, eJ});
            double al = alphaEE[eI];
            if (accd_ee(ea0, ea1, eb0, eb1, dea0, dea1, deb0, deb1, eta, thickness, al, &iters)) {
                if (alphaEE[eI] > al) alphaEE[eI] = al;
            }
        };
        if (brute) for (int eJ = eI + 1; eJ < m.nBE; ++eJ) body(eJ);
        else for (int eJ : partners) body(eJ);
    }
    if (m.nBE) stepSize = std::min(stepSize, *std::min_element(alphaEE.begin(), alphaEE.end()));
    R.step = stepSize;
    R.iters = iters;
    if (wantCand) {
        R.candPT.clear(); R.candEE.clear();
        for (auto& v : cPT) R.candPT.insert(R.candPT.end(), v.begin(), v.end());
        for (auto& v : cEE) R.candEE.insert(R.candEE.end(), v.begin(), v.end());
        std::sort(R.candPT.begin(), R.candPT.end());
        std::sort(R.candEE.begin(), R.candEE.end());
    }
    return status;
}

} // namespace orc
