// R&D prototype for the next round (not built into libidp_contact.so): the selected-eigenvector PSD projection of
// scripts/proto/psd_selected.cpp written the way the device wants it -- every register array statically indexed, the only
// dynamically indexed state (d, e of the values-only QL; the sorted eigenvalues) in a 2N-word shared store per row, the
// Householder vectors parked in a global scratch between the reduction and the back-transformation. Purpose: read the
// register / spill / shared-memory budget off `ptxas -v` before committing to the kernel split:
//   nvcc -std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a -Xptxas -v -DPSD_MINB=3 -c psd_selected_kernel.cu
// The same source compiles for the host (g++ -x c++ -DPSD_HOST) where test_psd_selected.py checks it against numpy.
#include <math.h>
#include <string.h>
#ifdef PSD_HOST
#define HD inline
#else
#define HD __device__ __forceinline__
#endif
#ifndef PSD_MINB
#define PSD_MINB 3
#endif

template <int N> HD constexpr int SI(int r, int c) { return r <= c ? (r * N - (r * (r - 1)) / 2 + (c - r)) : (c * N - (c * (c - 1)) / 2 + (r - c)); }

HD double rcp_(double x)
{
#ifdef PSD_HOST
    return 1.0 / x;
#else
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    r = fma(fma(-x, r, 1.0), r, r);
    return fma(fma(-x, r, 1.0), r, r);
#endif
}
HD double rsqrt_(double x)
{
#ifdef PSD_HOST
    return 1.0 / sqrt(x);
#else
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    const double e = fma(-x * y, y, 1.0);
    return fma(y * e, fma(0.375, e, 0.5), y);
#endif
}

// per-row dynamically indexed words: element w at p[w * STRIDE]
template <int STRIDE> struct Dyn {
    double* p;
    HD double& operator[](int w) const { return p[w * STRIDE]; }
};

// a: packed symmetric N x N (in: matrix, out: projection); vs: N*(N-1)/2 + N words of global scratch, element w at vs[w * vstride]
template <int N, int STRIDE>
HD bool psd_selected(double* a, Dyn<STRIDE> S, double* vs, long vstride)
{
    constexpr int NP = N * (N + 1) / 2, KMAX = N / 2;
    // ---- scale
    double am[4] = {0, 0, 0, 0};
#pragma unroll
    for (int i = 0; i < NP; ++i) am[i & 3] = fmax(am[i & 3], fabs(a[i]));
    const double amax = fmax(fmax(am[0], am[1]), fmax(am[2], am[3]));
    long long eb;
    memcpy(&eb, &amax, 8);
    eb = (eb >> 52) & 0x7ffLL;
    if (eb > 2044) eb = 2044;
    long long fb = (2045LL - eb) << 52, ub = (eb + 1LL) << 52;
    double fs, fu;
    memcpy(&fs, &fb, 8); memcpy(&fu, &ub, 8);
#pragma unroll
    for (int i = 0; i < NP; ++i) a[i] *= fs;
    // ---- Householder; v_k (over a(k, k+1..)) and beta_k go to the global scratch as soon as step k is done
    double d0[N], e0[N];
    double dmax = 0, emax = 0;
    int w = 0;
#pragma unroll
    for (int k = 0; k < N - 2; ++k) {
        const double x0 = a[SI<N>(k, k + 1)];
        double sigma = 0;
#pragma unroll
        for (int i = k + 2; i < N; ++i) sigma += a[SI<N>(k, i)] * a[SI<N>(k, i)];
        const double n2 = x0 * x0 + sigma;
        const bool act = sigma > 0.0 && n2 > 1e-280;
        double alpha = x0, bt = 0.0;
        if (act) {
            const double nrm = n2 * rsqrt_(n2);
            alpha = x0 >= 0 ? -nrm : nrm;
            bt = rcp_(nrm * (fabs(x0) + nrm));
            a[SI<N>(k, k + 1)] = x0 - alpha;
        }
        double pv[N], vp = 0;
#pragma unroll
        for (int i = 0; i < N; ++i) {
            if (i > k) {
                double s = 0;
#pragma unroll
                for (int j = 0; j < N; ++j) if (j > k) s += a[SI<N>(i, j)] * a[SI<N>(k, j)];
                pv[i] = bt * s;
                vp += pv[i] * a[SI<N>(k, i)];
            }
        }
        const double K = 0.5 * bt * vp;
#pragma unroll
        for (int i = 0; i < N; ++i) if (i > k) pv[i] -= K * a[SI<N>(k, i)];
#pragma unroll
        for (int i = 0; i < N; ++i)
#pragma unroll
            for (int j = 0; j < N; ++j)
                if (i > k && j >= i) a[SI<N>(i, j)] -= a[SI<N>(k, i)] * pv[j] + pv[i] * a[SI<N>(k, j)];
#pragma unroll
        for (int i = 0; i < N; ++i) if (i > k) vs[(w++) * vstride] = a[SI<N>(k, i)];
        vs[(w++) * vstride] = bt;
        e0[k] = alpha;
        d0[k] = a[SI<N>(k, k)];
        emax = fmax(emax, fabs(alpha));
    }
    d0[N - 2] = a[SI<N>(N - 2, N - 2)]; d0[N - 1] = a[SI<N>(N - 1, N - 1)];
    e0[N - 2] = a[SI<N>(N - 2, N - 1)]; e0[N - 1] = 0.0;
    emax = fmax(emax, fabs(e0[N - 2]));
#pragma unroll
    for (int i = 0; i < N; ++i) { dmax = fmax(dmax, fabs(d0[i])); S[i] = d0[i]; S[N + i] = e0[i]; }
    const double tnorm = dmax + emax, tol = 2e-15 * tnorm;
    // ---- values-only implicit QL on the shared copy
    int l = 0;
    for (int trip = 0; trip < 30 * N; ++trip) {
        unsigned negl = 0;
#pragma unroll
        for (int i = 0; i < N - 1; ++i) negl |= (fabs(S[N + i]) <= tol) ? (1u << i) : 0u;
        l += __builtin_ffs((int)~(negl >> l)) - 1;
        if (l >= N - 1) break;
        const int m = l + __builtin_ffs((int)((negl | (1u << (N - 1))) >> (l + 1)));
        const double dl = S[l], el = S[N + l];
        double g = (S[l + 1] - dl) * 0.5 * rcp_(fabs(el));
        if (el < 0) g = -g;
        const double r0 = (g * g + 1.0) * rsqrt_(g * g + 1.0);
        g = S[m] - dl + el * (g >= 0 ? rcp_(g + r0) : -rcp_(r0 - g));
        double s = 1.0, c = 1.0, p = 0.0;
        bool under = false;
        for (int i = m - 1; i >= l; --i) {
            const double ei = S[N + i], di = S[i], di1 = S[i + 1];
            const double f = s * ei, b = c * ei, r2 = f * f + g * g;
            if (!(r2 > 1e-290)) { S[i + 1] = di1 - p; S[N + i + 1] = 0.0; under = true; break; }
            const double ir = rsqrt_(r2);
            S[N + i + 1] = r2 * ir;
            s = f * ir; c = g * ir;
            g = di1 - p;
            const double rr = (di - g) * s + 2.0 * c * b;
            p = s * rr;
            S[i + 1] = g + p;
            g = c * rr - b;
        }
        S[N + m] = 0.0;
        if (under) continue;
        S[l] = dl - p; S[N + l] = g;
    }
    const bool converged = l >= N - 1;
    // ---- eigenvalues to registers, sorted ascending by an odd-even transposition network (static), then back to the store
    double lam[N];
#pragma unroll
    for (int i = 0; i < N; ++i) lam[i] = S[i];
#pragma unroll
    for (int pass = 0; pass < N; ++pass)
#pragma unroll
        for (int i = pass & 1; i + 1 < N; i += 2) {
            const double lo = fmin(lam[i], lam[i + 1]), hi = fmax(lam[i], lam[i + 1]);
            lam[i] = lo; lam[i + 1] = hi;
        }
    int nneg = 0, npos = 0;
#pragma unroll
    for (int i = 0; i < N; ++i) { nneg += lam[i] < 0; npos += lam[i] > 0; S[i] = lam[i]; }
    const bool usePos = npos <= nneg;
    const int first = usePos ? N - npos : 0, count = usePos ? npos : nneg; // count <= N/2
    // ---- eigenvectors of the smaller side by inverse iteration, S2 = sum lambda z z^T
    double Z[KMAX][N];
    double S2[NP];
#pragma unroll
    for (int i = 0; i < NP; ++i) S2[i] = 0.0;
    const double tiny = 2.3e-16 * tnorm + 1e-300;
    double last = 0;
#pragma unroll
    for (int k = 0; k < KMAX; ++k) {
        if (k < count) {
            const double lamk = S[first + k];
            double mu = lamk;
            if (k && mu - last <= 1e-15 * tnorm) mu = last + 1e-15 * tnorm;
            last = mu;
            // LU of T - mu I with partial pivoting
            double u0[N], u1[N], u2[N], lm[N];
            unsigned sw = 0;
            double ra = d0[0] - mu, rb = e0[0], rc = 0.0;
#pragma unroll
            for (int i = 0; i < N - 1; ++i) {
                const double sub = e0[i], nd = d0[i + 1] - mu, ne = (i + 1 < N - 1) ? e0[i + 1] : 0.0;
                if (fabs(sub) > fabs(ra)) {
                    sw |= 1u << i;
                    u0[i] = sub; u1[i] = nd; u2[i] = ne;
                    const double mlt = ra * rcp_(fabs(sub)) * (sub < 0 ? -1.0 : 1.0);
                    lm[i] = mlt;
                    ra = rb - mlt * nd; rb = rc - mlt * ne; rc = 0.0;
                }
                else {
                    if (fabs(ra) < tiny) ra = (ra < 0 ? -tiny : tiny);
                    u0[i] = ra; u1[i] = rb; u2[i] = rc;
                    const double mlt = sub * rcp_(fabs(ra)) * (ra < 0 ? -1.0 : 1.0);
                    lm[i] = mlt;
                    ra = nd - mlt * rb; rb = ne - mlt * rc; rc = 0.0;
                }
            }
            if (fabs(ra) < tiny) ra = (ra < 0 ? -tiny : tiny);
            u0[N - 1] = ra; u1[N - 1] = 0.0; u2[N - 1] = 0.0;
            double iu[N];
#pragma unroll
            for (int i = 0; i < N; ++i) iu[i] = rcp_(fabs(u0[i])) * (u0[i] < 0 ? -1.0 : 1.0);
            double z[N];
#pragma unroll
            for (int i = 0; i < N; ++i) z[i] = sin(1.0 + 2.399963 * (i + 1) + 0.7 * k) + 0.3 * cos(0.37 * (i + 3) * (k + 1)); // constants
#pragma unroll 1
            for (int it = 0; it < 3; ++it) {
#pragma unroll
                for (int i = 0; i < N - 1; ++i) {
                    if (sw & (1u << i)) { const double t = z[i]; z[i] = z[i + 1]; z[i + 1] = t - lm[i] * z[i]; }
                    else z[i + 1] -= lm[i] * z[i];
                }
#pragma unroll
                for (int i = N - 1; i >= 0; --i) {
                    double t = z[i];
                    if (i + 1 < N) t -= u1[i] * z[i + 1];
                    if (i + 2 < N) t -= u2[i] * z[i + 2];
                    z[i] = t * iu[i];
                }
#pragma unroll
                for (int q = 0; q < KMAX; ++q) {
                    if (q < k) {
                        double dot = 0;
#pragma unroll
                        for (int i = 0; i < N; ++i) dot += Z[q][i] * z[i];
#pragma unroll
                        for (int i = 0; i < N; ++i) z[i] -= dot * Z[q][i];
                    }
                }
                double big = 0;
#pragma unroll
                for (int i = 0; i < N; ++i) big = fmax(big, fabs(z[i]));
                if (!(big > 0)) { z[k] = 1.0; big = 1.0; }
                const double ib = rcp_(big);
                double nn = 0;
#pragma unroll
                for (int i = 0; i < N; ++i) { z[i] *= ib; nn += z[i] * z[i]; }
                const double inv = rsqrt_(nn);
#pragma unroll
                for (int i = 0; i < N; ++i) z[i] *= inv;
            }
#pragma unroll
            for (int i = 0; i < N; ++i) Z[k][i] = z[i];
#pragma unroll
            for (int i = 0; i < N; ++i)
#pragma unroll
                for (int j = 0; j < N; ++j) if (j >= i) S2[SI<N>(i, j)] += (lamk * z[i]) * z[j];
        }
    }
    // ---- T+ (packed) in the tridiagonal basis
#pragma unroll
    for (int i = 0; i < N; ++i)
#pragma unroll
        for (int j = 0; j < N; ++j)
            if (j >= i) {
                const double t = (j == i) ? d0[i] : ((j == i + 1) ? e0[i] : 0.0);
                a[SI<N>(i, j)] = usePos ? S2[SI<N>(i, j)] : t - S2[SI<N>(i, j)];
            }
    // ---- M+ = Q T+ Q^T: reflectors read back from the scratch, last one first
#pragma unroll
    for (int k = N - 3; k >= 0; --k) {
        w -= (N - k - 1) + 1;
        double v[N];
#pragma unroll
        for (int i = 0; i < N; ++i) v[i] = 0.0;
        int ww = w;
#pragma unroll
        for (int i = 0; i < N; ++i) if (i > k) v[i] = vs[(ww++) * vstride];
        const double bt = vs[ww * vstride];
        double pv[N], vp = 0;
#pragma unroll
        for (int i = 0; i < N; ++i) {
            double s = 0;
#pragma unroll
            for (int j = 0; j < N; ++j) if (j > k) s += a[SI<N>(i, j)] * v[j];
            pv[i] = bt * s;
            vp += v[i] * pv[i];
        }
        const double K = 0.5 * bt * vp;
#pragma unroll
        for (int i = 0; i < N; ++i) pv[i] -= K * v[i];
#pragma unroll
        for (int i = 0; i < N; ++i)
#pragma unroll
            for (int j = 0; j < N; ++j) if (j >= i) a[SI<N>(i, j)] -= v[i] * pv[j] + pv[i] * v[j];
    }
#pragma unroll
    for (int i = 0; i < NP; ++i) a[i] *= fu;
    return converged;
}

#ifdef PSD_HOST
extern "C" int psd_selected_dev(int n, const double* A, double* out)
{
    double store[18], scratch[64];
    bool ok = false;
    if (n == 9) {
        double m[45];
        for (int i = 0; i < 9; ++i) for (int j = i; j < 9; ++j) m[SI<9>(i, j)] = A[i * 9 + j];
        ok = psd_selected<9, 1>(m, Dyn<1>{store}, scratch, 1);
        for (int i = 0; i < 9; ++i) for (int j = 0; j < 9; ++j) out[i * 9 + j] = m[SI<9>(i, j)];
    }
    else {
        double m[21];
        for (int i = 0; i < 6; ++i) for (int j = i; j < 6; ++j) m[SI<6>(i, j)] = A[i * 6 + j];
        ok = psd_selected<6, 1>(m, Dyn<1>{store}, scratch, 1);
        for (int i = 0; i < 6; ++i) for (int j = 0; j < 6; ++j) out[i * 6 + j] = m[SI<6>(i, j)];
    }
    return ok ? 0 : 1;
}
#else
template <int N>
__global__ void __launch_bounds__(128, PSD_MINB) k_psd_selected(const double* __restrict__ Min, double* __restrict__ Mout, double* __restrict__ vscratch, long n)
{
    extern __shared__ double sh[];
    constexpr int NP = N * (N + 1) / 2;
    for (long r = (long)blockIdx.x * 128 + threadIdx.x; r < n; r += (long)gridDim.x * 128) {
        double a[NP];
#pragma unroll
        for (int i = 0; i < NP; ++i) a[i] = Min[i * n + r];
        psd_selected<N, 128>(a, Dyn<128>{sh + threadIdx.x}, vscratch + r, n);
#pragma unroll
        for (int i = 0; i < NP; ++i) Mout[i * n + r] = a[i];
    }
}
template __global__ void k_psd_selected<9>(const double*, double*, double*, long);
template __global__ void k_psd_selected<6>(const double*, double*, double*, long);
#endif
