// TEST INFRASTRUCTURE (never shipped, never linked into idp_b200/): the same flow time step (idp_b200/host/jgsl/shell_flow.h)
// with every contact operator served by the REFERENCE's own CPU loops -- FEM/IPC.h + Grid/SPATIAL_HASH.h compiled in
// oracle/_ref/libidp_ref_ipc.so --, the membrane / hinge terms by the reference's own FEM/Shell/MEMBRANE.h and BENDING.h
// (oracle/_ref/libidp_ref_shell.so), friction by its FEM/FRICTION.h, and the system matrix built by the reference's own Math/CSR_MATRIX.h
// (oracle/_ref/libidp_ref_csr.so: Construct_From_Triplet, += M, Project_DBC). Used to produce the per-step
// "PN iterations / contact #" trace (counter.txt, Shell/IMPLICIT_EULER.h:857-864) that the B200 build of the module must
// reproduce. The linear solve is a host Jacobi-preconditioned CG to 1e-12 (CHOLMOD is not in this image).
#pragma once
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <map>

#include "shell_flow.h"

extern "C" {
int refipc_constraint_set(int nV, const double* x, const double* x0, int nBN, const int* bn, int nBE, const int* be, int nBT, const int* bt,
    const unsigned char* dbc, double dHat2, double thickness, long cap, int* rows4, double* info2);
long refipc_barrier(int nV, const double* x, const double* x0, int n, const int* rows4, const double* w, double dHat2, double kappa,
    double thickness, int projectSPD, double* E, double* g, long cap, int* trow, int* tcol, double* tval);
double refipc_ccd(int nV, const double* x, int nBN, const int* bn, int nBE, const int* be, int nBT, const int* bt, const unsigned char* dbc,
    const double* dir, double thickness, double step);
double refipc_min_dist2(int nV, const double* x, int n, const int* rows4, double thickness, double* dist2);
// the reference's own shell energy terms (FEM/Shell/MEMBRANE.h, BENDING.h with KL = false; oracle/_ref/libidp_ref_shell.so)
long refshell_membrane(int nV, const double* x, int nE, const int* elem3, const double* ib3, const double* vol, const double* lambda, const double* mu,
    const unsigned char* dbc, double h, int projectSPD, double* E, double* g, long cap, int* tr, int* tc, double* tv);
long refshell_hinges(int nV, const double* x, int nH, const int* stencil4, const double* info3, double k, double bendingStiffMult, const unsigned char* dbc,
    double h, int projectSPD, double* E, double* g, long cap, int* tr, int* tc, double* tv);
long refipc_friction(int nV, const double* xb, const double* x, const double* xn, int n, const int* rows4, const double* w, double dHat2, double kappa,
    double thickness, double epsvh2, double mu, int projectSPD, int* nFric, int* fricRows4, double* closest2, double* basis6, double* normalForce,
    double* E, double* g, long cap, int* trow, int* tcol, double* tval);
long refipc_friction_comp(int nV, const double* xb, const double* x, const double* xn, int n, const int* rows4, const double* w, double dHat2, double kappa,
    double thickness, double epsvh2, double mu, int projectSPD, int* nFric, int* fricRows4, double* closest2, double* basis6, double* normalForce,
    double* E, double* g, long cap, int* trow, int* tcol, double* tval, int nComp, const int* compNodeRange, const double* muComp);
long ref_csr_system(int n, long nT, const int* r, const int* c, const double* v, const double* mdiag, const unsigned char* dbc, int dim,
    int* ptr, int* col, double* val, long cap);
}

namespace jgsl {

class RefLoopsBackend : public ContactBackend {
public:
    double pcg_rel_tol = 1e-12; // same knobs as the B200 backend (the module sets them); this backend always solves to 1e-12
    int pcg_max_iter = 100000;
    explicit RefLoopsBackend(int) {}
    const char* name() const override { return "reference-loops"; }

    void set_mesh(int nV, const std::vector<int>& tri3, const double* x, const std::vector<uint8_t>& dbc) override
    {
        nV_ = nV; dbc_ = dbc;
        // Find_Surface_Primitives_And_Compute_Area's ordering contract (Utils/MESHIO.h:768-834) with a std::map, as the reference does
        btri_ = tri3;
        std::map<std::pair<int, int>, double> es;
        std::vector<double> nodeArea((size_t)nV, 0.0);
        for (size_t f = 0; f < tri3.size() / 3; ++f) {
            const int* t = &tri3[3 * f];
            const double* a = x + 3 * (size_t)t[0]; const double* b = x + 3 * (size_t)t[1]; const double* c = x + 3 * (size_t)t[2];
            const double u[3] = {b[0] - a[0], b[1] - a[1], b[2] - a[2]}, v[3] = {c[0] - a[0], c[1] - a[1], c[2] - a[2]};
            const double n[3] = {u[1] * v[2] - u[2] * v[1], u[2] * v[0] - u[0] * v[2], u[0] * v[1] - u[1] * v[0]};
            const double area = 0.5 * std::sqrt(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);
            for (int i = 0; i < 3; ++i) {
                const int p = t[i], q = t[(i + 1) % 3];
                auto it = es.find(std::make_pair(q, p));
                if (it == es.end()) es[std::make_pair(p, q)] = area / 3;
                else it->second += area / 3;
                nodeArea[p] += area / 3;
            }
        }
        bedge_.clear();
        for (const auto& e : es) { bedge_.push_back(e.first.first); bedge_.push_back(e.first.second); }
        bnode_.clear();
        for (int v = 0; v < nV; ++v) if (nodeArea[v]) bnode_.push_back(v);
    }
    void set_rest_positions(const double* x0) override { x0_.assign(x0, x0 + 3 * (size_t)nV_); }
    void set_system_terms(const std::vector<int>& elem3, const std::vector<double>& vol, double h, const std::vector<double>& mass) override
    {
        elem_ = elem3; vol_ = vol; h_ = h; mass_ = mass;
    }
    void set_elastic_terms(const std::vector<int>& elem3, const std::vector<double>& ib3, const std::vector<double>& vol, const std::vector<double>& lambda,
        const std::vector<double>& mu, const std::vector<int>& stencil4, const std::vector<double>& info3, double k, double h) override
    {
        mElem_ = elem3; mIB_ = ib3; mVol_ = vol; mLam_ = lambda; mMu_ = mu; hSt_ = stencil4; hInfo_ = info3; hK_ = k; eH_ = h;
    }
    // Compute_Membrane_* then Compute_Bending_* of the reference itself: E (added), g (added), triplets (appended)
    void elastic(double* E, double* g, std::vector<int>* tr, std::vector<int>* tc, std::vector<double>* tv)
    {
        const int nM = (int)mElem_.size() / 3, nH = (int)hSt_.size() / 4;
        for (int part = 0; part < 2; ++part) {
            const long cap = tr ? (part == 0 ? 81L * nM : 144L * nH) : 0;
            if ((part == 0 ? nM : nH) == 0) continue;
            std::vector<int> r((size_t)cap), c((size_t)cap);
            std::vector<double> v((size_t)cap);
            long nt;
            if (part == 0)
                nt = refshell_membrane(nV_, x_.data(), nM, mElem_.data(), mIB_.data(), mVol_.data(), mLam_.data(), mMu_.data(), dbc_.data(), eH_, 1, E, g, cap,
                    tr ? r.data() : nullptr, c.data(), v.data());
            else
                nt = refshell_hinges(nV_, x_.data(), nH, hSt_.data(), hInfo_.data(), hK_, 1.0, dbc_.data(), eH_, 1, E, g, cap, tr ? r.data() : nullptr, c.data(), v.data());
            if (tr) { tr->insert(tr->end(), r.begin(), r.begin() + nt); tc->insert(tc->end(), c.begin(), c.begin() + nt); tv->insert(tv->end(), v.begin(), v.begin() + nt); }
        }
    }
    void elastic_energy(double& E) override { elastic(&E, nullptr, nullptr, nullptr, nullptr); }
    void elastic_gradient(double* g) override { elastic(nullptr, g, nullptr, nullptr, nullptr); }
    void set_positions(const double* x) override { x_.assign(x, x + 3 * (size_t)nV_); }
    int constraint_set(double dHat2, double thickness) override
    {
        long cap = 1 << 16;
        for (;;) {
            rows_.resize(4 * (size_t)cap); info_.resize(2 * (size_t)cap);
            const int n = refipc_constraint_set(nV_, x_.data(), x0_.data(), (int)bnode_.size(), bnode_.data(), (int)bedge_.size() / 2, bedge_.data(),
                (int)btri_.size() / 3, btri_.data(), dbc_.data(), dHat2, thickness, cap, rows_.data(), info_.data());
            if (n <= cap) {
                rows_.resize(4 * (size_t)n); info_.resize(2 * (size_t)n);
                return n;
            }
            cap = n;
        }
    }
    void barrier_energy(double dHat2, double kappa, double thickness, double& E) override
    {
        if (rows_.empty()) return;
        std::vector<double> w = weights();
        refipc_barrier(nV_, x_.data(), x0_.data(), (int)w.size(), rows_.data(), w.data(), dHat2, kappa, thickness, 0, &E, nullptr, 0, nullptr, nullptr,
            nullptr);
    }
    void barrier_gradient(double dHat2, double kappa, double thickness, double* g) override
    {
        if (rows_.empty()) return;
        std::vector<double> w = weights();
        refipc_barrier(nV_, x_.data(), x0_.data(), (int)w.size(), rows_.data(), w.data(), dHat2, kappa, thickness, 0, nullptr, g, 0, nullptr, nullptr,
            nullptr);
    }
    bool solve_newton_system(double dHat2, double kappa, double thickness, const std::vector<uint8_t>* projMask, const double* rhs, double* sol) override
    {
        // triplets in the reference's order: flow term (INC_POTENTIAL.h:323-339), then the barrier Hessians
        std::vector<int> tr, tc;
        std::vector<double> tv;
        for (size_t e = 0; e < elem_.size() / 3; ++e)
            for (int i = 0; i < 3; ++i)
                for (int d = 0; d < 3; ++d) {
                    const int a = elem_[3 * e + i] * 3 + d, p = elem_[3 * e + (i + 1) % 3] * 3 + d, q = elem_[3 * e + (i + 2) % 3] * 3 + d;
                    tr.push_back(a); tc.push_back(p); tv.push_back(-h_ * vol_[e] / 6);
                    tr.push_back(a); tc.push_back(q); tv.push_back(-h_ * vol_[e] / 6);
                    tr.push_back(a); tc.push_back(a); tv.push_back(2 * h_ * vol_[e] / 6);
                }
        if (!mElem_.empty() || !hSt_.empty()) elastic(nullptr, nullptr, &tr, &tc, &tv); // membrane, then hinges (INC_POTENTIAL.h:344-352)
        if (!rows_.empty()) {
            std::vector<double> w = weights();
            const long cap = 144 * (long)w.size();
            std::vector<int> br((size_t)cap), bc((size_t)cap);
            std::vector<double> bv((size_t)cap);
            const long nt = refipc_barrier(nV_, x_.data(), x0_.data(), (int)w.size(), rows_.data(), w.data(), dHat2, kappa, thickness, 1, nullptr, nullptr,
                cap, br.data(), bc.data(), bv.data());
            tr.insert(tr.end(), br.begin(), br.begin() + nt); tc.insert(tc.end(), bc.begin(), bc.begin() + nt);
            tv.insert(tv.end(), bv.begin(), bv.begin() + nt);
        }
        if (fMu_ > 0 && !fRows_.empty()) { // friction Hessian (INC_POTENTIAL.h:375-377)
            std::vector<int> fr, fc;
            std::vector<double> fv;
            friction_call(nullptr, nullptr, &fr, &fc, &fv);
            tr.insert(tr.end(), fr.begin(), fr.end()); tc.insert(tc.end(), fc.begin(), fc.end()); tv.insert(tv.end(), fv.begin(), fv.end());
        }
        const int n = 3 * nV_;
        std::vector<double> md((size_t)n);
        for (int v = 0; v < nV_; ++v) md[3 * v] = md[3 * v + 1] = md[3 * v + 2] = mass_[v];
        const long cap = (long)tv.size() + n;
        std::vector<int> ptr((size_t)n + 1), col((size_t)cap);
        std::vector<double> val((size_t)cap);
        const long nnz = ref_csr_system(n, (long)tv.size(), tr.data(), tc.data(), tv.data(), md.data(), projMask ? projMask->data() : dbc_.data(), 3, ptr.data(), col.data(), val.data(), cap);
        if (nnz < 0) return false;
        // Jacobi-preconditioned conjugate gradients
        std::vector<double> dinv((size_t)n), r(rhs, rhs + n), z((size_t)n), p((size_t)n), Ap((size_t)n);
        for (int i = 0; i < n; ++i) {
            double dii = 0;
            for (int k = ptr[i]; k < ptr[i + 1]; ++k) if (col[k] == i) dii = val[k];
            if (!(dii > 0)) return false;
            dinv[i] = 1.0 / dii;
        }
        std::fill(sol, sol + n, 0.0);
        double bnorm = 0;
        for (int i = 0; i < n; ++i) bnorm += rhs[i] * rhs[i];
        if (bnorm == 0) return true;
        double rz = 0;
        for (int i = 0; i < n; ++i) { z[i] = dinv[i] * r[i]; p[i] = z[i]; rz += r[i] * z[i]; }
        int it = 0;
        double rr = bnorm;
        for (; it < 100000 && rr > 1e-24 * bnorm; ++it) {
            double pAp = 0;
#pragma omp parallel for
            for (int i = 0; i < n; ++i) {
                double s = 0;
                for (int k = ptr[i]; k < ptr[i + 1]; ++k) s += val[k] * p[col[k]];
                Ap[i] = s;
            }
            for (int i = 0; i < n; ++i) pAp += Ap[i] * p[i]; // serial: the trace must not depend on the thread count
            if (!(pAp > 0)) return false;
            const double a = rz / pAp;
            double rzn = 0;
            rr = 0;
            for (int i = 0; i < n; ++i) {
                sol[i] += a * p[i];
                r[i] -= a * Ap[i];
                z[i] = dinv[i] * r[i];
                rzn += r[i] * z[i];
                rr += r[i] * r[i];
            }
            const double beta = rzn / rz;
            rz = rzn;
            for (int i = 0; i < n; ++i) p[i] = z[i] + beta * p[i];
        }
        printf("linear solve (host PCG): %d iterations, relative residual %le\n", it, std::sqrt(rr / bnorm));
        return true;
    }
    // the reference's own FEM/FRICTION.h: the basis is recomputed from the frozen state on every call (same values)
    long friction_update(double dHat2, double kappa, double thickness) override
    {
        fRows_ = rows_; fInfo_ = info_; fXb_ = x_; fDHat2_ = dHat2; fKappa_ = kappa; fXi_ = thickness;
        return friction_call(nullptr, nullptr, nullptr, nullptr, nullptr);
    }
    void friction_set(const double* xn, double epsv2h2, double mu) override
    {
        fMu_ = mu; fEps_ = epsv2h2;
        if (xn) fXn_.assign(xn, xn + 3 * (size_t)nV_);
    }
    long friction_call(double* E, double* g, std::vector<int>* tr, std::vector<int>* tc, std::vector<double>* tv)
    {
        const int n = (int)fRows_.size() / 4;
        if (!n) return 0;
        std::vector<double> w((size_t)n);
        for (int i = 0; i < n; ++i) w[i] = fInfo_[2 * i];
        std::vector<int> fr(4 * (size_t)n);
        std::vector<double> cp(2 * (size_t)n), bs(6 * (size_t)n), lam((size_t)n);
        int nf = 0;
        const bool ev = (E || g || tr) && fMu_ > 0;
        const long cap = tr ? 144 * (long)n : 0;
        if (tr) { tr->resize((size_t)cap); tc->resize((size_t)cap); tv->resize((size_t)cap); }
        const long nt = refipc_friction_comp(nV_, fXb_.data(), ev ? x_.data() : nullptr, ev ? fXn_.data() : nullptr, n, fRows_.data(), w.data(), fDHat2_, fKappa_, fXi_, fEps_,
            fMu_, 1, &nf, fr.data(), cp.data(), bs.data(), lam.data(), E, g, cap, tr ? tr->data() : nullptr, tr ? tc->data() : nullptr, tr ? tv->data() : nullptr,
            (int)fComp_.size(), fComp_.data(), fMuComp_.data());
        if (tr) { tr->resize((size_t)nt); tc->resize((size_t)nt); tv->resize((size_t)nt); }
        return nf;
    }
    void friction_set_components(const std::vector<int>& compNodeRange, const std::vector<double>& muComp) override
    {
        fComp_ = compNodeRange; fMuComp_ = muComp; // Compute_Friction_Coef runs right after the basis inside refipc_friction_comp
    }
    void friction_energy(double& E) override { if (fMu_ > 0) friction_call(&E, nullptr, nullptr, nullptr, nullptr); }
    void friction_gradient(double* g) override { if (fMu_ > 0) friction_call(nullptr, g, nullptr, nullptr, nullptr); }
    double ccd(const double* dir, double thickness, double alpha) override
    {
        return refipc_ccd(nV_, x_.data(), (int)bnode_.size(), bnode_.data(), (int)bedge_.size() / 2, bedge_.data(), (int)btri_.size() / 3, btri_.data(),
            dbc_.data(), dir, thickness, alpha);
    }
    bool min_dist2(double thickness, std::vector<double>* dist2, double& minDist2) override
    {
        const int n = (int)rows_.size() / 4;
        if (!n) return false;
        std::vector<double> d((size_t)n);
        minDist2 = refipc_min_dist2(nV_, x_.data(), n, rows_.data(), thickness, d.data());
        if (dist2) *dist2 = d;
        return true;
    }
    void get_rows(std::vector<int>& rows4, std::vector<double>& info2) override { rows4 = rows_; info2 = info_; }
    void set_rows(const std::vector<int>& rows4, const std::vector<double>& info2) override { rows_ = rows4; info_ = info2; }

private:
    std::vector<double> weights() const
    {
        std::vector<double> w(rows_.size() / 4);
        for (size_t i = 0; i < w.size(); ++i) w[i] = info_[2 * i];
        return w;
    }
    int nV_ = 0;
    double h_ = 0;
    std::vector<int> bnode_, bedge_, btri_, elem_, rows_, mElem_, hSt_;
    std::vector<double> mIB_, mVol_, mLam_, mMu_, hInfo_;
    double hK_ = 0, eH_ = 0;
    std::vector<int> fRows_;
    std::vector<double> fInfo_, fXb_, fXn_, fMuComp_;
    std::vector<int> fComp_;
    double fDHat2_ = 0, fKappa_ = 0, fXi_ = 0, fMu_ = 0, fEps_ = 0;
    std::vector<uint8_t> dbc_;
    std::vector<double> x_, x0_, vol_, mass_, info_;
};

} // namespace jgsl
