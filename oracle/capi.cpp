// ORACLE — TEST INFRASTRUCTURE ONLY (see orc_math.hpp header).
// Flat C entry points for ctypes (oracle/binding.py). Not part of the product's C ABI.
#include "orc_ipc.hpp"
#include "orc_system.hpp"
#include "orc_elastic.hpp"
#include "orc_friction.hpp"
#include <chrono>

using namespace orc;

extern "C" {

struct OrcMesh {
    int nV; const double* X; const double* X0;
    int nBN; const int* bnode; int nBE; const int* bedge; int nBT; const int* btri;
    const uint8_t* dbc;
};
static Mesh to_mesh(const OrcMesh* m)
{
    Mesh r;
    r.nV = m->nV; r.X = m->X; r.X0 = m->X0; r.nBN = m->nBN; r.bnode = m->bnode; r.nBE = m->nBE; r.bedge = m->bedge;
    r.nBT = m->nBT; r.btri = m->btri; r.dbc = m->dbc;
    return r;
}
static std::vector<Row> to_rows(long n, const int* rows)
{
    std::vector<Row> r(n);
    for (long i = 0; i < n; ++i) r[i] = {rows[4 * i], rows[4 * i + 1], rows[4 * i + 2], rows[4 * i + 3]};
    return r;
}

// ---- constraint set -------------------------------------------------------------------------
void* orc_constraint_set(const OrcMesh* m, double dHat2, double thickness, int brute, int wantCand)
{
    auto* R = new ConstraintSetResult();
    compute_constraint_set(to_mesh(m), dHat2, thickness, brute != 0, wantCand != 0, *R);
    return R;
}
long orc_cs_count(void* h, int which)
{
    auto* R = (ConstraintSetResult*)h;
    return which == 0 ? (long)R->rows.size() : (which == 1 ? (long)R->candPT.size() : (long)R->candEE.size());
}
void orc_cs_copy(void* h, int* rows, double* info, int* candPT, int* candEE)
{
    auto* R = (ConstraintSetResult*)h;
    if (rows) for (size_t i = 0; i < R->rows.size(); ++i) for (int k = 0; k < 4; ++k) rows[4 * i + k] = R->rows[i][k];
    if (info) for (size_t i = 0; i < R->info.size(); ++i) { info[2 * i] = R->info[i][0]; info[2 * i + 1] = R->info[i][1]; }
    if (candPT) for (size_t i = 0; i < R->candPT.size(); ++i) { candPT[2 * i] = R->candPT[i][0]; candPT[2 * i + 1] = R->candPT[i][1]; }
    if (candEE) for (size_t i = 0; i < R->candEE.size(); ++i) { candEE[2 * i] = R->candEE[i][0]; candEE[2 * i + 1] = R->candEE[i][1]; }
}
void orc_cs_free(void* h) { delete (ConstraintSetResult*)h; }

// ---- barrier --------------------------------------------------------------------------------
int orc_barrier(const OrcMesh* m, long n, const int* rows, const double* weight, double dHat2, double kappa,
    double thickness, double* E)
{
    return compute_barrier(to_mesh(m), to_rows(n, rows), weight, dHat2, kappa, thickness, *E);
}
int orc_barrier_gradient(const OrcMesh* m, long n, const int* rows, const double* weight, double dHat2, double kappa,
    double thickness, double* g)
{
    return compute_barrier_gradient(to_mesh(m), to_rows(n, rows), weight, dHat2, kappa, thickness, g);
}
struct HessHandle { Triplets T; CSR A; int status; };
void* orc_barrier_hessian(const OrcMesh* m, long n, const int* rows, const double* weight, double dHat2, double kappa,
    double thickness, int projectSPD, int buildCSR)
{
    auto* h = new HessHandle();
    h->status = compute_barrier_hessian(to_mesh(m), to_rows(n, rows), weight, dHat2, kappa, thickness, projectSPD != 0, h->T);
    if (buildCSR && h->status == OK) csr_from_triplets(3 * m->nV, h->T, h->A);
    return h;
}
int orc_hess_status(void* h) { return ((HessHandle*)h)->status; }
long orc_hess_count(void* h, int which) { auto* H = (HessHandle*)h; return which == 0 ? (long)H->T.v.size() : (long)H->A.col.size(); }
void orc_hess_copy(void* h, int* tr, int* tc, double* tv, int* ptr, int* col, double* val)
{
    auto* H = (HessHandle*)h;
    if (tr) std::copy(H->T.r.begin(), H->T.r.end(), tr);
    if (tc) std::copy(H->T.c.begin(), H->T.c.end(), tc);
    if (tv) std::copy(H->T.v.begin(), H->T.v.end(), tv);
    if (ptr) std::copy(H->A.ptr.begin(), H->A.ptr.end(), ptr);
    if (col) std::copy(H->A.col.begin(), H->A.col.end(), col);
    if (val) std::copy(H->A.val.begin(), H->A.val.end(), val);
}
void orc_hess_free(void* h) { delete (HessHandle*)h; }

// ---- elastic terms of the shell system (orc_elastic.hpp) ---------------------------------------------------------
// Per element: E[e], g (accumulated into g3nV), H (81 / 144 doubles per element, row major, zero for skipped elements),
// active[e] = 0 where the reference skips the element (all vertices Dirichlet, MEMBRANE.h:25 / BENDING.h:56-60; det IB == 0).
void orc_membrane_batch(int nElem, const int* elem3, const double* X, const double* ib3, const double* coef, const double* lambda, const double* mu,
    const uint8_t* dbc, int projectSPD, double* E, double* g3nV, double* H81, uint8_t* active)
{
    std::vector<double> ge(g3nV ? 9 * (size_t)nElem : 0, 0.0);
#pragma omp parallel for
    for (int e = 0; e < nElem; ++e) {
        const int* t = elem3 + 3 * e;
        active[e] = 0; E[e] = 0;
        if (H81) std::memset(H81 + 81 * (size_t)e, 0, 81 * sizeof(double));
        if (dbc && dbc[t[0]] && dbc[t[1]] && dbc[t[2]]) continue;
        const V3 x[3] = {ld3(X + 3 * t[0]), ld3(X + 3 * t[1]), ld3(X + 3 * t[2])};
        double g[9];
        if (!membrane_EgH(x, ib3 + 3 * e, coef[e], lambda[e], mu[e], projectSPD != 0, E + e, g, H81 ? H81 + 81 * (size_t)e : nullptr)) continue;
        active[e] = 1;
        if (g3nV) std::memcpy(&ge[9 * (size_t)e], g, sizeof(g));
    }
    if (g3nV) // serial scatter in element order (the reference's loop is serial, MEMBRANE.h:71): reproducible sums
        for (int e = 0; e < nElem; ++e)
            if (active[e])
                for (int k = 0; k < 3; ++k)
                    for (int d = 0; d < 3; ++d) g3nV[3 * elem3[3 * e + k] + d] += ge[9 * (size_t)e + 3 * k + d];
}
void orc_hinge_batch(int nHinge, const int* stencil4, const double* X, const double* info3, double kh2, const uint8_t* dbc, int projectSPD,
    double* E, double* g3nV, double* H144, uint8_t* active)
{
    std::vector<double> ge(g3nV ? 12 * (size_t)nHinge : 0, 0.0);
#pragma omp parallel for
    for (int e = 0; e < nHinge; ++e) {
        const int* t = stencil4 + 4 * e;
        active[e] = 0; E[e] = 0;
        if (H144) std::memset(H144 + 144 * (size_t)e, 0, 144 * sizeof(double));
        if (dbc && dbc[t[0]] && dbc[t[1]] && dbc[t[2]] && dbc[t[3]]) continue;
        const V3 x[4] = {ld3(X + 3 * t[0]), ld3(X + 3 * t[1]), ld3(X + 3 * t[2]), ld3(X + 3 * t[3])};
        double g[12];
        E[e] = hinge_EgH(x, info3[3 * e], kh2 * info3[3 * e + 1] / info3[3 * e + 2], projectSPD != 0, g, H144 ? H144 + 144 * (size_t)e : nullptr);
        active[e] = 1;
        if (g3nV) std::memcpy(&ge[12 * (size_t)e], g, sizeof(g));
    }
    if (g3nV)
        for (int e = 0; e < nHinge; ++e)
            if (active[e])
                for (int k = 0; k < 4; ++k)
                    for (int d = 0; d < 3; ++d) g3nV[3 * stencil4[4 * e + k] + d] += ge[12 * (size_t)e + 3 * k + d];
}
double orc_dihedral_angle(const double* x12) { return dihedral_angle(ld3(x12), ld3(x12 + 3), ld3(x12 + 6), ld3(x12 + 9)); }

// ---- lagged friction (orc_friction.hpp) ---------------------------------------------------------------------------
// basis at Xb from the contact rows (outputs like the reference's containers: rows, closest (2), basis (3x2 column major), normal
// force), then E (added) / g (added) / triplets at X relative to Xn when X is given. Returns the triplet count (<= cap written).
long orc_friction(int nV, const double* Xb, const double* X, const double* Xn, long n, const int* rows4, const double* weight, double dHat2, double kappa,
    double thickness, double epsvh2, double mu, int projectSPD, int* nFric, int* fricRows4, double* closest2, double* basis6, double* normalForce,
    double* E, double* g, long cap, int* tr, int* tc, double* tv)
{
    (void)nV;
    std::vector<FrictionRow> fr;
    friction_basis(Xb, to_rows(n, rows4), weight, dHat2, kappa, thickness, fr);
    *nFric = (int)fr.size();
    for (size_t i = 0; i < fr.size(); ++i) {
        for (int k = 0; k < 4; ++k) fricRows4[4 * i + k] = fr[i].row[k];
        closest2[2 * i] = fr[i].cp[0]; closest2[2 * i + 1] = fr[i].cp[1];
        basis6[6 * i] = fr[i].t0.x; basis6[6 * i + 1] = fr[i].t0.y; basis6[6 * i + 2] = fr[i].t0.z;
        basis6[6 * i + 3] = fr[i].t1.x; basis6[6 * i + 4] = fr[i].t1.y; basis6[6 * i + 5] = fr[i].t1.z;
        normalForce[i] = fr[i].lam;
    }
    if (!X) return 0;
    Triplets T;
    friction_eval(X, Xn, fr, epsvh2, mu, projectSPD != 0, E, g, tr ? &T : nullptr);
    const long nt = (long)T.v.size();
    for (long i = 0; i < nt && i < cap; ++i) { tr[i] = T.r[i]; tc[i] = T.c[i]; tv[i] = T.v[i]; }
    return nt;
}

// Compute_Friction_Coef on friction rows (as returned by orc_friction) and their normal forces, in place; 0 on success
int orc_friction_coef(int n, const int* fricRows4, int nComp, const int* compNodeRange, const double* muComp, double* normalForce)
{
    std::vector<FrictionRow> fr((size_t)n);
    for (int i = 0; i < n; ++i) {
        FrictionRow& f = fr[i];
        for (int k = 0; k < 4; ++k) f.row[k] = fricRows4[4 * i + k];
        f.v[0] = f.row[0] >= 0 ? f.row[0] : -f.row[0] - 1; f.v[1] = f.row[1]; f.v[2] = f.row[2]; f.v[3] = f.row[3];
        f.lam = normalForce[i];
    }
    if (!friction_coef(fr, std::vector<int>(compNodeRange, compNodeRange + nComp), std::vector<double>(muComp, muComp + (size_t)nComp * nComp))) return 1;
    for (int i = 0; i < n; ++i) normalForce[i] = fr[i].lam;
    return 0;
}

// ---- system matrix around the barrier Hessian (orc_system.hpp) -------------------------------------------------
// triplets = [flow term][barrier rows] -> Construct_From_Triplet -> += M -> Project_DBC (INC_POTENTIAL.h:321-394)
void* orc_system_matrix(const OrcMesh* m, long n, const int* rows, const double* weight, double dHat2, double kappa,
    double thickness, int projectSPD, int nElem, const int* elem3, const double* vol, double h, const double* mass, int projectDBC)
{
    auto* H = new HessHandle();
    if (nElem > 0) flow_term_triplets(nElem, elem3, vol, h, H->T);
    H->status = n > 0 ? compute_barrier_hessian(to_mesh(m), to_rows(n, rows), weight, dHat2, kappa, thickness, projectSPD != 0, H->T) : OK;
    if (H->status == OK) {
        csr_from_triplets(3 * m->nV, H->T, H->A);
        if (mass) csr_add_mass(H->A, m->nV, mass);
        if (projectDBC && m->dbc) project_dbc(H->A, m->dbc, 3);
    }
    return H;
}
void* orc_surface(int nV, int nF, const int* tri, const double* X)
{
    auto* S = new SurfacePrimitives();
    find_surface_primitives(nV, nF, tri, X, *S);
    return S;
}
long orc_surface_count(void* h, int which)
{
    auto* S = (SurfacePrimitives*)h;
    return which == 0 ? (long)S->bnode.size() : (which == 1 ? (long)S->bedge.size() / 2 : (long)S->btri.size() / 3);
}
void orc_surface_copy(void* h, int* bnode, int* bedge, int* btri, double* BNArea, double* BEArea, double* BTArea)
{
    auto* S = (SurfacePrimitives*)h;
    std::copy(S->bnode.begin(), S->bnode.end(), bnode);
    std::copy(S->bedge.begin(), S->bedge.end(), bedge);
    std::copy(S->btri.begin(), S->btri.end(), btri);
    std::copy(S->BNArea.begin(), S->BNArea.end(), BNArea);
    std::copy(S->BEArea.begin(), S->BEArea.end(), BEArea);
    std::copy(S->BTArea.begin(), S->BTArea.end(), BTArea);
}
void orc_surface_free(void* h) { delete (SurfacePrimitives*)h; }

// per-row local E / g / H (dense n x n, n = 3*nv) for one row; returns status, writes nv and stencil
int orc_row_EgH(const OrcMesh* m, const int* row, double weight, double dHat2, double kappa, double thickness,
    int projectSPD, double* E, double* g, double* H, int* nv, int* verts)
{
    Decoded d;
    const Row r = {row[0], row[1], row[2], row[3]};
    const int st = row_EgH(to_mesh(m), r, weight, adjusted_dhat2(dHat2, thickness), kappa, thickness * thickness,
        projectSPD != 0, E, g, H, &d);
    *nv = d.nv;
    for (int i = 0; i < 4; ++i) verts[i] = d.v[i];
    return st;
}

void orc_min_dist2(const OrcMesh* m, long n, const int* rows, double thickness, double* dist2, double* minDist2)
{
    std::vector<double> d;
    compute_min_dist2(to_mesh(m), to_rows(n, rows), thickness, d, *minDist2);
    if (dist2) std::copy(d.begin(), d.end(), dist2);
}

// ---- CCD ------------------------------------------------------------------------------------
void* orc_ccd(const OrcMesh* m, const double* dir, double thickness, int brute, int wantCand, double* step,
    double* step_after_clamp, long* iters, int* status)
{
    auto* R = new CCDResult();
    R->step = *step;
    *status = compute_intersection_free_stepsize(to_mesh(m), dir, thickness, brute != 0, wantCand != 0, *R);
    *step = R->step;
    *step_after_clamp = R->step_after_clamp;
    *iters = R->iters;
    return R;
}
long orc_ccd_count(void* h, int which) { auto* R = (CCDResult*)h; return which == 1 ? (long)R->candPT.size() : (long)R->candEE.size(); }
void orc_ccd_copy(void* h, int* candPT, int* candEE)
{
    auto* R = (CCDResult*)h;
    if (candPT) for (size_t i = 0; i < R->candPT.size(); ++i) { candPT[2 * i] = R->candPT[i][0]; candPT[2 * i + 1] = R->candPT[i][1]; }
    if (candEE) for (size_t i = 0; i < R->candEE.size(); ++i) { candEE[2 * i] = R->candEE[i][0]; candEE[2 * i + 1] = R->candEE[i][1]; }
}
void orc_ccd_free(void* h) { delete (CCDResult*)h; }

// ---- per-pair functions for unit tests --------------------------------------------------------
// x: 4 points x 3 (or fewer for PE / PP)
int orc_pt_type(const double* x) { return pt_type(ld3(x), ld3(x + 3), ld3(x + 6), ld3(x + 9)); }
int orc_ee_type(const double* x) { return ee_type(ld3(x), ld3(x + 3), ld3(x + 6), ld3(x + 9)); }
double orc_dist2(int kind, const double* x) // 0 PP, 1 PE, 2 PT, 3 EE, 4 PT unclassified, 5 EE unclassified, 6 EE cross norm2
{
    switch (kind) {
    case 0: return dist2_pp(ld3(x), ld3(x + 3));
    case 1: return dist2_pe(ld3(x), ld3(x + 3), ld3(x + 6));
    case 2: return dist2_pt(ld3(x), ld3(x + 3), ld3(x + 6), ld3(x + 9));
    case 3: return dist2_ee(ld3(x), ld3(x + 3), ld3(x + 6), ld3(x + 9));
    case 4: return dist2_pt_unclassified(ld3(x), ld3(x + 3), ld3(x + 6), ld3(x + 9));
    case 5: return dist2_ee_unclassified(ld3(x), ld3(x + 3), ld3(x + 6), ld3(x + 9));
    default: return ee_cross_norm2(ld3(x), ld3(x + 3), ld3(x + 6), ld3(x + 9));
    }
}
// closed-form (jet=0) or autodiff (jet=1) gradient/Hessian; kind: 0 PP, 1 PE, 2 PT, 3 EE, 6 EE cross norm2
void orc_grad_hess(int kind, int jet, const double* x, double* g, double* H)
{
    const V3 a = ld3(x), b = ld3(x + 3);
    switch (kind) {
    case 0: pp_grad_hess(a, b, g, H); break;
    case 1: if (jet) pe_jet(a, b, ld3(x + 6), g, H); else pe_grad_hess(a, b, ld3(x + 6), g, H); break;
    case 2: if (jet) pt_jet(a, b, ld3(x + 6), ld3(x + 9), g, H); else pt_grad_hess(a, b, ld3(x + 6), ld3(x + 9), g, H); break;
    case 3: if (jet) ee_jet(a, b, ld3(x + 6), ld3(x + 9), g, H); else ee_grad_hess(a, b, ld3(x + 6), ld3(x + 9), g, H); break;
    default: if (jet) eecn2_jet(a, b, ld3(x + 6), ld3(x + 9), g, H); else ee_cross_norm2_grad_hess(a, b, ld3(x + 6), ld3(x + 9), g, H); break;
    }
}
void orc_mollifier(const double* x, double eps_x, double* e, double* g, double* H)
{
    ee_mollifier_all(ld3(x), ld3(x + 3), ld3(x + 6), ld3(x + 9), eps_x, *e, g, H);
}
double orc_mollifier_threshold(const double* x0) { return ee_mollifier_threshold(ld3(x0), ld3(x0 + 3), ld3(x0 + 6), ld3(x0 + 9)); }
void orc_barrier_scalar(double d, double dHat2, double kappa, double* b, double* g, double* h)
{
    *b = barrier(d, dHat2, kappa); *g = barrier_g(d, dHat2, kappa); *h = barrier_h(d, dHat2, kappa);
}
void orc_make_pd(int n, double* H) { make_pd(n, H); }
void orc_sym_eig(int n, const double* A, double* lam, double* V) { sym_eig_jacobi(n, A, lam, V); }
// ACCD: x = 4 points, d = 4 displacements; kind 0 PT, 1 EE. returns 1 if hit
int orc_accd(int kind, const double* x, const double* d, double eta, double thickness, double* toc, long* iters)
{
    long it = 0;
    bool r;
    if (kind == 0) r = accd_pt(ld3(x), ld3(x + 3), ld3(x + 6), ld3(x + 9), ld3(d), ld3(d + 3), ld3(d + 6), ld3(d + 9), eta, thickness, *toc, &it);
    else r = accd_ee(ld3(x), ld3(x + 3), ld3(x + 6), ld3(x + 9), ld3(d), ld3(d + 3), ld3(d + 6), ld3(d + 9), eta, thickness, *toc, &it);
    if (iters) *iters = it;
    return r ? 1 : 0;
}
int orc_aabb(int kind, const double* x, const double* d, double dist) // 0 PT static, 1 EE static, 2 PT swept, 3 EE swept
{
    switch (kind) {
    case 0: return pt_cd_broadphase(ld3(x), ld3(x + 3), ld3(x + 6), ld3(x + 9), dist);
    case 1: return ee_cd_broadphase(ld3(x), ld3(x + 3), ld3(x + 6), ld3(x + 9), dist);
    case 2: return pt_ccd_broadphase(ld3(x), ld3(x + 3), ld3(x + 6), ld3(x + 9), ld3(d), ld3(d + 3), ld3(d + 6), ld3(d + 9), dist);
    default: return ee_ccd_broadphase(ld3(x), ld3(x + 3), ld3(x + 6), ld3(x + 9), ld3(d), ld3(d + 3), ld3(d + 6), ld3(d + 9), dist);
    }
}
double orc_tree_mean(const double* a, long n) { return tree_sum(a, n) / n; }
double orc_mean_edge_length(const OrcMesh* m) { return SpatialHash::mean_edge_length(to_mesh(m)); }

// wall-clock helper for the CPU-baseline leg of bench.py
double orc_now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

} // extern "C"
