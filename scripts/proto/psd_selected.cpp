// R&D prototype (not product code, not built into libidp_contact.so): PSD projection of a small symmetric matrix WITHOUT the
// full eigenvector matrix -- the candidate replacement for the QL-with-vectors solver of idp_b200/csrc/psd_lowrank.cuh
// (DESIGN.md section 8, item 1). Per row it needs 2N words of dynamically indexed state (d, e) instead of N*N + 2N:
//   1. Householder tridiagonalisation  T = Q^T M Q                      (static control flow)
//   2. eigenvalues of T by values-only implicit QL                       (the only dynamically indexed part)
//   3. eigenvectors of the SMALLER side of the spectrum only (<= N/2 of them) by inverse iteration on T - lambda I
//      (LU with partial pivoting, tiny pivots replaced; modified Gram-Schmidt against the vectors already found)
//   4. T+ = sum_{lambda>0} lambda z z^T   or   T - sum_{lambda<0} lambda z z^T,   then   M+ = Q T+ Q^T.
// scripts/proto/test_psd_selected.py checks it against numpy on the hard spectra of tests/test_host.py and on real rows.
//   g++ -O2 -shared -fPIC -o libpsd_selected.so psd_selected.cpp
#include <algorithm>
#include <cmath>
#include <cstring>

namespace {

template <int N>
struct Sel {
    double A[N][N];   // in: symmetric matrix; out: projection
    double v[N][N];   // Householder vectors (row k: v_k over indices k+1..N-1)
    double beta[N];
    double d[N], e[N], lam[N];
    long stats_iters = 0;

    void tridiagonalise()
    {
        for (int k = 0; k < N - 2; ++k) {
            double sigma = 0;
            for (int i = k + 2; i < N; ++i) sigma += A[k][i] * A[k][i];
            const double x0 = A[k][k + 1], n2 = x0 * x0 + sigma;
            double alpha = x0, bt = 0;
            for (int i = 0; i < N; ++i) v[k][i] = 0;
            if (sigma > 0 && n2 > 1e-280) {
                const double nrm = std::sqrt(n2);
                alpha = x0 >= 0 ? -nrm : nrm;
                bt = 1.0 / (nrm * (std::fabs(x0) + nrm));
                v[k][k + 1] = x0 - alpha;
                for (int i = k + 2; i < N; ++i) v[k][i] = A[k][i];
            }
            beta[k] = bt;
            double p[N], vp = 0;
            for (int i = k + 1; i < N; ++i) {
                double s = 0;
                for (int j = k + 1; j < N; ++j) s += A[i][j] * v[k][j];
                p[i] = bt * s;
                vp += p[i] * v[k][i];
            }
            const double K = 0.5 * bt * vp;
            for (int i = k + 1; i < N; ++i) p[i] -= K * v[k][i];
            for (int i = k + 1; i < N; ++i)
                for (int j = k + 1; j < N; ++j) A[i][j] -= v[k][i] * p[j] + p[i] * v[k][j];
            e[k] = alpha;
        }
        e[N - 2] = A[N - 2][N - 1];
        e[N - 1] = 0;
        for (int i = 0; i < N; ++i) d[i] = A[i][i];
    }

    // values-only implicit QL on a copy of (d, e)
    bool eigenvalues(double tol)
    {
        double dd[N], ee[N];
        for (int i = 0; i < N; ++i) { dd[i] = d[i]; ee[i] = e[i]; }
        int l = 0;
        for (int trip = 0; trip < 30 * N; ++trip) {
            while (l < N - 1 && std::fabs(ee[l]) <= tol) ++l;
            if (l >= N - 1) break;
            int m = l + 1;
            while (m < N - 1 && std::fabs(ee[m]) > tol) ++m;
            double g = (dd[l + 1] - dd[l]) / (2.0 * ee[l]);
            const double r0 = std::sqrt(g * g + 1.0);
            g = dd[m] - dd[l] + ee[l] / (g + (g >= 0 ? r0 : -r0));
            double s = 1, c = 1, p = 0;
            bool under = false;
            for (int i = m - 1; i >= l; --i) {
                ++stats_iters;
                const double f = s * ee[i], b = c * ee[i];
                const double r = std::sqrt(f * f + g * g);
                ee[i + 1] = r;
                if (!(r > 1e-145)) { dd[i + 1] -= p; ee[i + 1] = 0; ee[m] = 0; under = true; break; }
                s = f / r; c = g / r;
                g = dd[i + 1] - p;
                const double rr = (dd[i] - g) * s + 2.0 * c * b;
                p = s * rr;
                dd[i + 1] = g + p;
                g = c * rr - b;
            }
            ee[m] = 0;
            if (under) continue;
            dd[l] -= p; ee[l] = g;
        }
        for (int i = 0; i < N; ++i) lam[i] = dd[i];
        return l >= N - 1;
    }

    // one eigenvector of T for the (accurate) eigenvalue mu by inverse iteration; prev: vectors to stay orthogonal to
    void eigenvector(double mu, const double (*prev)[N], int nprev, int seed, double tnorm, double* z)
    {
        // LU of T - mu I with partial pivoting: rows become (u0, u1, u2) with multipliers lm and swap flags
        double u0[N], u1[N], u2[N], lm[N];
        bool sw[N];
        const double tiny = 2.3e-16 * tnorm + 1e-300;
        double a = d[0] - mu, b = (N > 1) ? e[0] : 0.0; // current row k: (a, b, 0)
        double c2 = 0;                                 // second super-diagonal fill of the current row
        for (int k = 0; k < N - 1; ++k) {
            const double sub = e[k], nd = d[k + 1] - mu, ne = (k + 1 < N - 1) ? e[k + 1] : 0.0; // next row: (sub, nd, ne)
            if (std::fabs(sub) > std::fabs(a)) { // swap rows k and k+1
                sw[k] = true;
                u0[k] = sub; u1[k] = nd; u2[k] = ne;
                const double mlt = a / sub;
                lm[k] = mlt;
                a = b - mlt * nd; b = c2 - mlt * ne; c2 = 0;
            }
            else {
                sw[k] = false;
                if (std::fabs(a) < tiny) a = (a < 0 ? -tiny : tiny);
                u0[k] = a; u1[k] = b; u2[k] = c2;
                const double mlt = sub / a;
                lm[k] = mlt;
                a = nd - mlt * b; b = ne - mlt * c2; c2 = 0;
            }
        }
        if (std::fabs(a) < tiny) a = (a < 0 ? -tiny : tiny);
        u0[N - 1] = a; u1[N - 1] = 0; u2[N - 1] = 0;
        // start vector: fixed pseudo-random pattern depending on the index of the eigenvalue
        for (int i = 0; i < N; ++i) z[i] = std::sin(1.0 + 2.399963 * (i + 1) + 0.7 * seed) + 0.3 * std::cos(0.37 * (i + 3) * (seed + 1));
        for (int it = 0; it < 3; ++it) {
            // forward: apply the recorded row operations to the right-hand side
            for (int k = 0; k < N - 1; ++k) {
                if (sw[k]) { const double t = z[k]; z[k] = z[k + 1]; z[k + 1] = t - lm[k] * z[k]; }
                else z[k + 1] -= lm[k] * z[k];
            }
            // back substitution with the two super-diagonals
            for (int k = N - 1; k >= 0; --k) {
                double t = z[k];
                if (k + 1 < N) t -= u1[k] * z[k + 1];
                if (k + 2 < N) t -= u2[k] * z[k + 2];
                z[k] = t / u0[k];
            }
            // modified Gram-Schmidt against the vectors already accepted, then normalise
            for (int q = 0; q < nprev; ++q) {
                double dot = 0;
                for (int i = 0; i < N; ++i) dot += prev[q][i] * z[i];
                for (int i = 0; i < N; ++i) z[i] -= dot * prev[q][i];
            }
            double big = 0;
            for (int i = 0; i < N; ++i) big = std::max(big, std::fabs(z[i]));
            if (!(big > 0)) { z[seed % N] = 1.0; big = 1.0; }
            double nn = 0;
            for (int i = 0; i < N; ++i) { z[i] /= big; nn += z[i] * z[i]; }
            const double inv = 1.0 / std::sqrt(nn);
            for (int i = 0; i < N; ++i) z[i] *= inv;
        }
    }

    bool project()
    {
        double amax = 0;
        for (int i = 0; i < N; ++i) for (int j = 0; j < N; ++j) amax = std::max(amax, std::fabs(A[i][j]));
        if (!(amax >= 2.3e-308)) { std::memset(A, 0, sizeof(A)); return true; }
        int ex; std::frexp(amax, &ex);
        const double fs = std::ldexp(1.0, -ex), fu = std::ldexp(1.0, ex);
        for (int i = 0; i < N; ++i) for (int j = 0; j < N; ++j) A[i][j] *= fs;
        tridiagonalise();
        double dmax = 0, emax = 0;
        for (int i = 0; i < N; ++i) { dmax = std::max(dmax, std::fabs(d[i])); emax = std::max(emax, std::fabs(e[i])); }
        const double tnorm = dmax + emax;
        const bool ok = eigenvalues(2e-15 * tnorm);
        int idx[N], nneg = 0, npos = 0;
        for (int i = 0; i < N; ++i) { idx[i] = i; if (lam[i] < 0) ++nneg; else if (lam[i] > 0) ++npos; }
        const bool usePos = npos <= nneg;
        std::sort(idx, idx + N, [&](int x, int y) { return lam[x] < lam[y]; });
        double Z[N][N], S[N][N];
        std::memset(S, 0, sizeof(S));
        int nz = 0;
        double last = 0;
        for (int t = 0; t < N; ++t) {
            const int j = idx[t];
            if (usePos ? !(lam[j] > 0) : !(lam[j] < 0)) continue;
            double mu = lam[j];
            if (nz && mu - last <= 1e-15 * tnorm) mu = last + 1e-15 * tnorm; // separate (numerically) equal shifts a little
            last = mu;
            eigenvector(mu, Z, nz, nz, tnorm, Z[nz]);
            for (int a = 0; a < N; ++a) for (int b = 0; b < N; ++b) S[a][b] += lam[j] * Z[nz][a] * Z[nz][b];
            ++nz;
        }
        // T+ in the tridiagonal basis
        double Tp[N][N];
        for (int a = 0; a < N; ++a)
            for (int b = 0; b < N; ++b) {
                double t = 0;
                if (!usePos) { t = (a == b) ? d[a] : ((a == b + 1) ? e[b] : ((b == a + 1) ? e[a] : 0.0)); t -= S[a][b]; }
                else t = S[a][b];
                Tp[a][b] = t;
            }
        // M+ = Q T+ Q^T with Q = H_0 H_1 ... H_{N-3}: apply the reflectors from the last to the first on both sides
        for (int k = N - 3; k >= 0; --k) {
            const double bt = beta[k];
            if (bt == 0) continue;
            double p[N];
            for (int i = 0; i < N; ++i) { double s = 0; for (int j = k + 1; j < N; ++j) s += Tp[i][j] * v[k][j]; p[i] = bt * s; }
            double vp = 0;
            for (int i = k + 1; i < N; ++i) vp += v[k][i] * p[i];
            const double K = 0.5 * bt * vp;
            for (int i = 0; i < N; ++i) p[i] -= K * v[k][i];
            for (int i = 0; i < N; ++i) for (int j = 0; j < N; ++j) Tp[i][j] -= v[k][i] * p[j] + p[i] * v[k][j];
        }
        for (int i = 0; i < N; ++i) for (int j = 0; j < N; ++j) A[i][j] = 0.5 * (Tp[i][j] + Tp[j][i]) * fu;
        return ok;
    }
};

} // namespace

extern "C" int psd_selected(int n, const double* A, double* out, long* qlsteps)
{
    if (n == 9) { static Sel<9> s; s.stats_iters = 0; std::memcpy(s.A, A, sizeof(s.A)); const bool ok = s.project(); std::memcpy(out, s.A, sizeof(s.A)); if (qlsteps) *qlsteps = s.stats_iters; return ok ? 0 : 1; }
    if (n == 6) { static Sel<6> s; s.stats_iters = 0; std::memcpy(s.A, A, sizeof(s.A)); const bool ok = s.project(); std::memcpy(out, s.A, sizeof(s.A)); if (qlsteps) *qlsteps = s.stats_iters; return ok ? 0 : 1; }
    return -1;
}
