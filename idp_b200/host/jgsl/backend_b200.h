// The B200 backend of the flow time step (shell_flow.h): every operator is one call into the C ABI of libidp_contact.so
// (include/idp_contact.h). The constraint set, the barrier rows, the system matrix (flow + mass + projected barrier
// Hessians), Project_DBC and the linear solve never leave the device; the host sees scalars, the 3 nV gradient and the
// 3 nV search direction. There is no CPU fallback: without a CUDA device idp_create fails and the step exits like the
// reference does on a fatal condition (message + exit(-1)).
#pragma once
#include <chrono>
#include <cstdio>
#include <cstdlib>

#include "idp_contact.h"
#include "shell_flow.h"

namespace jgsl {

class B200Backend : public ContactBackend {
public:
    double pcg_rel_tol = 1e-12; // Solve_Direct is a factorisation in the reference: an ill-conditioned flow (batch.py's cat) follows it from the first step only below 1e-10
    int pcg_max_iter = 200000; // the reference factorises, so the cap only bounds a broken system: ill-conditioned flows (batch.py's cat) need > 20,000
    long pcg_iters_total = 0, newton_solves = 0;

    explicit B200Backend(int device = 0)
    {
        const int rc = idp_create(device, &ctx_);
        if (rc != IDP_OK) {
            printf("JGSL (B200): no usable CUDA device / libidp_contact context (status %d); this build has no CPU path\n", rc);
            exit(-1);
        }
    }
    ~B200Backend() override
    {
        if (std::getenv("IDP_PROFILE")) { // wall clock per operator, accumulated over the run
            static const char* names[N_OPS] = {"set_mesh", "set_terms", "set_positions", "constraint_set", "barrier_energy", "barrier_all(g+H)", "assemble(H only)",
                "project_dbc", "pcg", "ccd", "min_dist2", "elastic_E", "elastic_g", "friction", "rows_io"};
            double tot = 0;
            for (int i = 0; i < N_OPS; ++i) tot += sec_[i];
            printf("[B200 backend] %.3f s inside the C ABI:", tot);
            for (int i = 0; i < N_OPS; ++i) if (calls_[i]) printf("  %s %.3f s / %ld", names[i], sec_[i], calls_[i]);
            printf("\n");
        }
        if (ctx_) idp_destroy(ctx_);
    }
    const char* name() const override { return "B200"; }
    idp_ctx* context() const { return ctx_; }

    void set_mesh(int nV, const std::vector<int>& tri3, const double* x, const std::vector<uint8_t>& dbc) override
    {
        // Find_Surface_Primitives_And_Compute_Area runs once per mesh content, not once per time step (IMPLICIT_EULER.h:222 "TODO: only once")
        if (nV != nV_ || tri3 != tri_ || dbc != dbc_) {
            Scope t(*this, OP_MESH);
            check(idp_set_mesh_from_triangles(ctx_, nV, (int)(tri3.size() / 3), tri3.data(), 3, x, 3, dbc.data()));
            nV_ = nV; tri_ = tri3; dbc_ = dbc;
            termsSet_ = false; elasticSet_ = false;
        }
        fresh_ = false;
    }
    void set_rest_positions(const double* x0) override { Scope t(*this, OP_POS); check(idp_set_rest_positions(ctx_, x0, 3)); fresh_ = false; }
    void set_system_terms(const std::vector<int>& elem3, const std::vector<double>& vol, double h, const std::vector<double>& mass) override
    {
        if (termsSet_ && elem3 == elem_ && vol == vol_ && h == h_ && mass == mass_) return;
        Scope t(*this, OP_TERMS);
        check(idp_system_set_flow_term(ctx_, (int)(elem3.size() / 3), elem3.data(), 3, vol.data(), h));
        check(idp_system_set_mass(ctx_, mass.data()));
        elem_ = elem3; vol_ = vol; h_ = h; mass_ = mass;
        termsSet_ = true;
        fresh_ = false;
    }
    void set_elastic_terms(const std::vector<int>& elem3, const std::vector<double>& ib3, const std::vector<double>& vol, const std::vector<double>& lambda,
        const std::vector<double>& mu, const std::vector<int>& stencil4, const std::vector<double>& info3, double k, double h) override
    {
        if (elasticSet_ && elem3 == mElem_ && ib3 == mIB_ && vol == mVol_ && lambda == mLam_ && mu == mMu_ && stencil4 == hSt_ && info3 == hInfo_ && k == hK_ &&
            h == eH_)
            return;
        Scope t(*this, OP_TERMS);
        check(idp_system_set_membrane(ctx_, (int)(elem3.size() / 3), elem3.data(), 3, ib3.data(), vol.data(), lambda.data(), mu.data(), h));
        check(idp_system_set_hinges(ctx_, (int)(stencil4.size() / 4), stencil4.data(), info3.data(), k, h));
        mElem_ = elem3; mIB_ = ib3; mVol_ = vol; mLam_ = lambda; mMu_ = mu; hSt_ = stencil4; hInfo_ = info3; hK_ = k; eH_ = h;
        elasticSet_ = true;
        fresh_ = false;
    }
    void elastic_energy(double& E) override { Scope t(*this, OP_EE); check(idp_elastic_energy(ctx_, &E)); }
    void elastic_gradient(double* g) override { Scope t(*this, OP_EG); check(idp_elastic_gradient(ctx_, g, 3)); }
    void set_positions(const double* x) override { Scope t(*this, OP_POS); check(idp_set_positions(ctx_, x, 3)); fresh_ = false; }
    int constraint_set(double dHat2, double thickness) override
    {
        Scope t(*this, OP_CS);
        int n = 0;
        check(idp_constraint_set(ctx_, dHat2, thickness, &n));
        fresh_ = false;
        return n;
    }
    void barrier_energy(double dHat2, double kappa, double thickness, double& E) override
    {
        Scope t(*this, OP_BE);
        check(idp_barrier_energy(ctx_, dHat2, kappa, thickness, &E));
    }
    void barrier_gradient(double dHat2, double kappa, double thickness, double* g) override
    {
        // one pass over the rows produces the gradient AND the assembled system matrix the solve of this iterate needs
        Scope t(*this, OP_BALL);
        long nnz = 0;
        check(idp_barrier_all(ctx_, dHat2, kappa, thickness, 1, nullptr, &nnz));
        check(idp_get_gradient(ctx_, g, 3));
        fresh_ = true; freshKappa_ = kappa; freshDHat2_ = dHat2;
    }
    bool solve_newton_system(double dHat2, double kappa, double thickness, const std::vector<uint8_t>* projMask, const double* rhs, double* sol) override
    {
        if (!(fresh_ && freshKappa_ == kappa && freshDHat2_ == dHat2)) {
            Scope t(*this, OP_H);
            long nnz = 0;
            check(idp_barrier_hessian(ctx_, dHat2, kappa, thickness, 1, &nnz));
        }
        fresh_ = false; // Project_DBC rewrites the values in place
        { Scope t(*this, OP_PROJ); check(idp_project_dbc_mask(ctx_, projMask ? projMask->data() : nullptr)); }
        int iters = 0;
        double rel = 0;
        { Scope t(*this, OP_PCG); check(idp_solve_pcg(ctx_, rhs, sol, pcg_rel_tol, pcg_max_iter, &iters, &rel)); }
        pcg_iters_total += iters; ++newton_solves;
        printf("linear solve (device PCG): %d iterations, relative residual %le\n", iters, rel);
        return rel <= 1e3 * pcg_rel_tol && rel == rel;
    }
    long friction_update(double dHat2, double kappa, double thickness) override
    {
        Scope t(*this, OP_FRIC);
        long n = 0;
        check(idp_friction_update(ctx_, dHat2, kappa, thickness, &n));
        fresh_ = false;
        return n;
    }
    void friction_set(const double* xn, double epsv2h2, double mu) override { check(idp_friction_set(ctx_, xn, 3, epsv2h2, mu)); fresh_ = false; }
    void friction_set_components(const std::vector<int>& compNodeRange, const std::vector<double>& muComp) override
    {
        check(idp_friction_set_components(ctx_, (int)compNodeRange.size(), compNodeRange.data(), muComp.data()));
    }
    void friction_energy(double& E) override { Scope t(*this, OP_FRIC); check(idp_friction_energy(ctx_, &E)); }
    void friction_gradient(double* g) override { Scope t(*this, OP_FRIC); check(idp_friction_gradient(ctx_, g, 3)); }
    double ccd(const double* dir, double thickness, double alpha) override
    {
        Scope t(*this, OP_CCD);
        check(idp_ccd_step(ctx_, dir, 3, thickness, &alpha));
        return alpha;
    }
    bool min_dist2(double thickness, std::vector<double>* dist2, double& minDist2) override
    {
        const long n = idp_last_count(ctx_, 0);
        if (n <= 0) return false;
        Scope t(*this, OP_MIND);
        if (dist2) dist2->resize((size_t)n);
        check(idp_min_dist2(ctx_, thickness, dist2 ? dist2->data() : nullptr, &minDist2));
        return true;
    }
    void get_rows(std::vector<int>& rows4, std::vector<double>& info2) override
    {
        const long n = idp_last_count(ctx_, 0);
        rows4.resize(4 * (size_t)n); info2.resize(2 * (size_t)n);
        Scope t(*this, OP_ROWS);
        if (n) check(idp_get_constraints(ctx_, rows4.data(), info2.data()));
    }
    void set_rows(const std::vector<int>& rows4, const std::vector<double>& info2) override
    {
        Scope t(*this, OP_ROWS);
        check(idp_set_constraints(ctx_, (int)(rows4.size() / 4), rows4.data(), info2.data()));
        fresh_ = false;
    }

private:
    enum { OP_MESH, OP_TERMS, OP_POS, OP_CS, OP_BE, OP_BALL, OP_H, OP_PROJ, OP_PCG, OP_CCD, OP_MIND, OP_EE, OP_EG, OP_FRIC, OP_ROWS, N_OPS };
    double sec_[N_OPS] = {};
    long calls_[N_OPS] = {};
    struct Scope {
        B200Backend& b; int op; std::chrono::steady_clock::time_point t0;
        Scope(B200Backend& bb, int o) : b(bb), op(o), t0(std::chrono::steady_clock::now()) {}
        ~Scope() { b.sec_[op] += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count(); ++b.calls_[op]; }
    };
    void check(int rc)
    {
        if (rc == IDP_OK) return;
        printf("JGSL (B200): %s (status %d)\n", idp_last_error(ctx_), rc); // the reference prints and exits on its fatal conditions
        exit(-1);
    }
    idp_ctx* ctx_ = nullptr;
    int nV_ = -1;
    std::vector<int> tri_, elem_, mElem_, hSt_;
    std::vector<double> mIB_, mVol_, mLam_, mMu_, hInfo_;
    double hK_ = 0, eH_ = 0;
    bool elasticSet_ = false;
    std::vector<uint8_t> dbc_;
    std::vector<double> vol_, mass_;
    double h_ = 0, freshKappa_ = 0, freshDHat2_ = 0;
    bool termsSet_ = false, fresh_ = false;
};

} // namespace jgsl
