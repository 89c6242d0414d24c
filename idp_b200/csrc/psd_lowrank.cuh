// Register-resident evaluation of one constraint row: energy, gradient and (PSD-projected) Hessian.
//
// Replaces the dense path of pair_deriv.cuh (kept as the straightforward formulation the tests compare against):
//   * every quantity is formed in REDUCED (difference) coordinates: (w,u,v) in R^9 for four-vertex stencils, (w,u) in R^6
//     for point-edge, because all distances and the mollifier are translation invariant;
//   * the projection onto the PSD cone (makePD, Library/Math/UTILS.h:9-27) is done on the 9x9 / 6x6 matrix
//     M = T^T H_r T, where T maps an ORTHONORMAL basis of the translation-free subspace (Hadamard basis of the stencil
//     vertices) to the difference coordinates. Since H = Q M Q^T with Q orthonormal, eig(H) = eig(M) + three zeros and
//     makePD(H) = Q makePD(M) Q^T exactly;
//   * the symmetric eigen-decomposition is Householder tridiagonalisation (registers, fully unrolled) followed by
//     implicit-shift QL with the eigenvector matrix in shared memory (make_pd_ql below); point-point rows have a closed form.
// Reference semantics: FEM/IPC.h:801-938 (E), 1012-1254 (g), 1390-1729 (H); tolerance 1e-10 relative.
#pragma once
#include "pair_deriv.cuh"
#include <string.h>

namespace idp {

// packed upper triangle of a symmetric N x N matrix
template <int N>
IDP_HD constexpr int SI(int r, int c) { return r <= c ? (r * N - (r * (r - 1)) / 2 + (c - r)) : (c * N - (c * (c - 1)) / 2 + (r - c)); }

// approximate reciprocal / reciprocal square root refined by Newton steps (no IEEE corner cases needed: inputs are
// positive, finite and far from the subnormal range on this path)
IDP_HD double rcp_nr2(double x)
{
#if defined(__CUDA_ARCH__)
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    r = fma(fma(-x, r, 1.0), r, r);
    r = fma(fma(-x, r, 1.0), r, r);
    return r;
#else
    return 1.0 / x;
#endif
}
template <int STEPS>
IDP_HD double rsqrt_nr(double x)
{
#if defined(__CUDA_ARCH__)
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    const double h = 0.5 * x;
#pragma unroll
    for (int i = 0; i < STEPS; ++i) y = fma(fma(-h * y, y, 0.5), y, y);
    return y;
#else
    return 1.0 / sqrt(x);
#endif
}
// one third-order step on the ~2^-20 hardware approximation (MUFU.RSQ64H works on the high word): with e = 1 - x y^2,
// y (1 + e/2 + 3e^2/8) has relative error ~(5/16) e^3 < 2^-60, i.e. correctly rounded up to the last FMA. Four dependent
// operations instead of the nine of three Newton steps: this sits on the serial chain of every Givens rotation.
IDP_HD double rsqrt_h3(double x)
{
#if defined(__CUDA_ARCH__)
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    const double e = fma(-x * y, y, 1.0);
    return fma(y * e, fma(0.375, e, 0.5), y);
#else
    return 1.0 / sqrt(x);
#endif
}

// ---- makePD through Householder tridiagonalisation + implicit QL ----------------------------------------------------------
// Result V max(lambda,0) V^T at roughly a third of the arithmetic of a cyclic Jacobi iteration:
//   1. Householder reduction of the packed matrix to tridiagonal form T = Q^T M Q: fixed control flow, fully unrolled,
//      every index static -> registers only;
//   2. Q is formed explicitly in the work store S (shared memory on the device);
//   3. implicit-shift QL on (d, e) with the Givens rotations accumulated into the columns of S. Every lane of a warp
//      performs ONE QL iteration per trip (deflation scan, Wilkinson shift, bulge chase from its own m down to its own l),
//      so lanes stay in the chase together and a trip costs the longest chase in the warp. The chase is software
//      pipelined: the column rotation of step i+1 (9 independent rows) is issued together with the scalar dependency
//      chain that produces the rotation of step i, and the column shared by consecutive rotations is carried in
//      registers (one column load + one column store per step);
//   4. reconstruction V max(lambda,0) V^T into the packed matrix.
// Deflation threshold: |e| <= 2e-15 * (max|d_i| + max|e_i|). Dropping such an e perturbs T by that much; the projection
// onto the PSD cone is non-expansive, so the result moves by no more than the sum of the dropped values (<= ~2e-14 ||M||).
//
// QlStore: per-row work store of N*N eigenvector entries, N diagonal and N sub-diagonal values; element w of the row
// lives at p[w * STRIDE] (device: STRIDE = threads per block, p = shared base + threadIdx.x -> conflict-free 64-bit accesses
// whatever the per-lane column index is; host: STRIDE = 1 over a local buffer).
template <int N, int STRIDE>
struct QlStore {
    static constexpr int WORDS = N * N + 2 * N;
    static constexpr int STRIDE_V = STRIDE;
    double* p;
    // entry (r, c) of the eigenvector matrix relative to a column pointer col(c)
    IDP_HD double* col(int c) const { return p + c * STRIDE; }
    static IDP_HD double z(const double* colp, int r) { return colp[r * N * STRIDE]; }
    static IDP_HD void zset(double* colp, int r, double x) { colp[r * N * STRIDE] = x; }
    static IDP_HD double d(const double* colp) { return colp[N * N * STRIDE]; }
    static IDP_HD void dset(double* colp, double x) { colp[N * N * STRIDE] = x; }
    static IDP_HD double e(const double* colp) { return colp[(N * N + N) * STRIDE]; }
    static IDP_HD void eset(double* colp, double x) { colp[(N * N + N) * STRIDE] = x; }
};
IDP_HD long long pun_double_bits(double x)
{
#if defined(__CUDA_ARCH__)
    return __double_as_longlong(x);
#else
    long long b;
    memcpy(&b, &x, 8);
    return b;
#endif
}
IDP_HD double pun_bits_double(long long b)
{
#if defined(__CUDA_ARCH__)
    return __longlong_as_double(b);
#else
    double x;
    memcpy(&x, &b, 8);
    return x;
#endif
}
IDP_HD int first_set_bit(unsigned x) // 1-based, 0 when x == 0
{
#if defined(__CUDA_ARCH__)
    return __ffs((int)x);
#else
    return __builtin_ffs((int)x);
#endif
}
#ifdef IDP_QL_STATS
struct QlStats { long rows, trips, givens; int n; int chase[512]; int negcount; };
static QlStats g_ql_stats = {0, 0, 0, 0, {0}, 0};
#endif
template <int N, class ST>
IDP_HD bool make_pd_ql(double* a, ST& S) // false: the QL iteration cap was reached (never observed; reported, not ignored)
{
    // ---- 0. exact power-of-two scaling to [0.5, 1): the squares formed below neither underflow nor overflow whatever the
    // magnitude of the row (mollified rows near e = 0 are tiny, rows at contact are huge); the result is scaled back
    double am[4] = {0, 0, 0, 0}; // four independent chains: a single fmax chain over 45 entries is pure latency
#pragma unroll
    for (int i = 0; i < N * (N + 1) / 2; ++i) am[i & 3] = fmax(am[i & 3], fabs(a[i]));
    const double amax = fmax(fmax(am[0], am[1]), fmax(am[2], am[3]));
    if (!(amax >= 2.3e-308)) { // zero (or denormal) matrix: its projection is zero
#pragma unroll
        for (int i = 0; i < N * (N + 1) / 2; ++i) a[i] = 0.0;
        return true;
    }
    long long ebits = (pun_double_bits(amax) >> 52) & 0x7ffLL;                 // biased exponent of the largest entry
    if (ebits > 2044) ebits = 2044;                                            // (keeps both scale factors finite)
    const double fscale = pun_bits_double((2045LL - ebits) << 52);            // 2^(1022 - e): amax * fscale in [0.5, 1)
    const double funscale = pun_bits_double((ebits + 1LL) << 52);             // its inverse 2^(e - 1022)
    const double unscale = (ebits < 723 || ebits > 1323) ? funscale : 1.0;
    if (ebits < 723 || ebits > 1323) { // |entries| outside [2^-300, 2^300]: only then can a square leave the double range
#pragma unroll
        for (int i = 0; i < N * (N + 1) / 2; ++i) a[i] *= fscale;
    }
    // ---- 1. Householder: for column k annihilate a(k+2.., k); v is stored over a(k, k+1..), beta kept
    double beta[N - 2];
    double emax = 0;
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
            const double nrm = n2 * rsqrt_nr<3>(n2);
            alpha = x0 >= 0 ? -nrm : nrm;
            bt = rcp_nr2(nrm * (fabs(x0) + nrm)); // 2 / v^T v with v0 = x0 - alpha
            a[SI<N>(k, k + 1)] = x0 - alpha;
        }
        beta[k] = bt;
        // p = beta A22 v, K = beta/2 v.p, w = p - K v, A22 -= v w^T + w v^T
        double pv[N];
        double vp = 0;
#pragma unroll
        for (int i = k + 1; i < N; ++i) {
            double s = 0;
#pragma unroll
            for (int j = k + 1; j < N; ++j) s += a[SI<N>(i, j)] * a[SI<N>(k, j)];
            pv[i] = bt * s;
            vp += pv[i] * a[SI<N>(k, i)];
        }
        const double K = 0.5 * bt * vp;
#pragma unroll
        for (int i = k + 1; i < N; ++i) pv[i] -= K * a[SI<N>(k, i)];
#pragma unroll
        for (int i = 0; i < N; ++i)
#pragma unroll
            for (int j = 0; j < N; ++j)
                if (i > k && j >= i) a[SI<N>(i, j)] -= a[SI<N>(k, i)] * pv[j] + pv[i] * a[SI<N>(k, j)];
        ST::eset(S.col(k), alpha);
        emax = fmax(emax, fabs(alpha));
    }
    ST::eset(S.col(N - 2), a[SI<N>(N - 2, N - 1)]);
    ST::eset(S.col(N - 1), 0.0);
    double dmax = 0;
    emax = fmax(emax, fabs(a[SI<N>(N - 2, N - 1)]));
#pragma unroll
    for (int i = 0; i < N; ++i) {
        ST::dset(S.col(i), a[SI<N>(i, i)]);
        dmax = fmax(dmax, fabs(a[SI<N>(i, i)]));
    }
    const double tol = 2e-15 * (dmax + emax);
    // ---- 2. Q = H_0 H_1 ... H_{N-3} (backward accumulation; rows/columns <= k stay those of the identity)
#pragma unroll
    for (int i = 0; i < N; ++i)
#pragma unroll
        for (int j = 0; j < N; ++j)
            if (i == 0 || j == 0 || (i == N - 1 && j == N - 1)) ST::zset(S.col(j), i, i == j ? 1.0 : 0.0);
#pragma unroll
    for (int k = N - 3; k >= 0; --k) {
        const double bt = beta[k];
        { // column k+1 of the current block is e_{k+1}: s = beta v_{k+1}
            const double s = bt * a[SI<N>(k, k + 1)];
#pragma unroll
            for (int i = k + 1; i < N; ++i) ST::zset(S.col(k + 1), i, (i == k + 1 ? 1.0 : 0.0) - s * a[SI<N>(k, i)]);
        }
#pragma unroll
        for (int j = k + 2; j < N; ++j) {
            double cj[N];
            double s = 0;
#pragma unroll
            for (int i = k + 2; i < N; ++i) {
                cj[i] = ST::z(S.col(j), i);
                s += a[SI<N>(k, i)] * cj[i];
            }
            s *= bt;
            ST::zset(S.col(j), k + 1, -s * a[SI<N>(k, k + 1)]);
#pragma unroll
            for (int i = k + 2; i < N; ++i) ST::zset(S.col(j), i, cj[i] - s * a[SI<N>(k, i)]);
        }
    }
    // ---- 3. implicit QL with Wilkinson shift (tql2 / tqli structure), rotations accumulated into S
    int l = 0;
#ifdef IDP_QL_STATS
    ++g_ql_stats.rows; g_ql_stats.n = 0;
#endif
    for (int trip = 0; trip < 30 * N; ++trip) {
        // deflation scan: bit i set <=> e_i negligible (all loads independent, static offsets)
        unsigned negl = 0;
#pragma unroll
        for (int i = 0; i < N - 1; ++i) negl |= (fabs(ST::e(S.col(i))) <= tol) ? (1u << i) : 0u;
        l += first_set_bit(~(negl >> l)) - 1;                                 // skip converged eigenvalues
        if (l >= N - 1) break;
        const int m = l + first_set_bit((negl | (1u << (N - 1))) >> (l + 1)); // first negligible e above l (or the end)
        double* cl = S.col(l);
        const double* cm = S.col(m);
        double zc[N], zl[N]; // zc: column carried between consecutive rotations (starts as column m); zl: column l,
                             // only touched by the last rotation of the chase, fetched here so its latency is hidden
#pragma unroll
        for (int k = 0; k < N; ++k) { zc[k] = ST::z(cm, k); zl[k] = ST::z(cl, k); }
        const double dl = ST::d(cl), el = ST::e(cl);
        double g = (ST::d(cl + ST::STRIDE_V) - dl) * 0.5 * rcp_nr2(fabs(el));
        if (el < 0) g = -g;
        const double r0 = (g * g + 1.0) * rsqrt_nr<2>(g * g + 1.0);
        g = ST::d(cm) - dl + el * (g >= 0 ? rcp_nr2(g + r0) : -rcp_nr2(r0 - g));
        double s = 1.0, c = 1.0, p = 0.0;
        double cp = 1.0, sp = 0.0; // rotation whose column update is still pending
        bool underflow = false;
#ifdef IDP_QL_STATS
        ++g_ql_stats.trips; g_ql_stats.givens += m - l; if (g_ql_stats.n < 512) g_ql_stats.chase[g_ql_stats.n++] = m - l;
#endif
        // The body is straight-line code on purpose: the scalar recurrence (s, c, g, p) -> next rotation is one serial
        // dependency chain (~12 FP64 operations + MUFU); the column rotation of the PREVIOUS step (9 independent rows) has to
        // be issued inside its stalls. With `if (ok)` / `if (pend)` as branches the compiler emitted chain and rotation as
        // separate blocks (BSSY/BSYNC) and nothing overlapped (ncu source view). Now: the rotation arithmetic is
        // unconditional (the first step applies the identity: cp = 1, sp = 0 and z0 == zc), only its stores are predicated;
        // the underflow guard feeds 1.0 into the chain and is handled after the fact. The scalars of the next step are
        // fetched one step ahead (d_{i+1} of the next step is this step's d_i, not yet modified).
        double* ci = S.col(m - 1);
        double ei = ST::e(ci), di = ST::d(ci), di1 = ST::d(ci + ST::STRIDE_V);
        double sn, cn, rr, pn, gg, bb, er; // results of the scalar recurrence of one step
        bool ok;
        auto recurrence = [&]() {
            const double f = s * ei;
            bb = c * ei;
            const double r2 = f * f + g * g;
            ok = r2 > 1e-290;
            const double r2s = ok ? r2 : 1.0;
            const double ir = rsqrt_h3(r2s);
            er = r2s * ir;
            sn = f * ir; cn = g * ir;
            gg = di1 - p;
            rr = (di - gg) * sn + 2.0 * cn * bb;
            pn = sn * rr;
        };
        auto split_on_underflow = [&]() { // tqli's recovery: split here; the carried column is the current column i+1
            ST::dset(ci + ST::STRIDE_V, di1 - p);
            ST::eset(ci + ST::STRIDE_V, 0.0);
#pragma unroll
            for (int k = 0; k < N; ++k) ST::zset(ci + ST::STRIDE_V, k, zc[k]);
            underflow = true;
        };
        auto commit = [&](double eiN, double diN) {
            ST::eset(ci + ST::STRIDE_V, er);
            ST::dset(ci + ST::STRIDE_V, gg + pn);
            g = cn * rr - bb;
            s = sn; c = cn; p = pn;
            cp = cn; sp = sn;
            di1 = di; ei = eiN; di = diN;
        };
        { // first step (i = m - 1): no rotation pending yet
            const bool more = m - 1 > l;
            const double eiN = more ? ST::e(ci - ST::STRIDE_V) : 0.0, diN = more ? ST::d(ci - ST::STRIDE_V) : 0.0;
            recurrence();
            if (!ok) split_on_underflow();
            else commit(eiN, diN);
            ci -= ST::STRIDE_V;
        }
        if (!underflow) {
            for (int i = m - 2; i >= l; --i, ci -= ST::STRIDE_V) {
                const bool more = i > l;
                const double eiN = more ? ST::e(ci - ST::STRIDE_V) : 0.0, diN = more ? ST::d(ci - ST::STRIDE_V) : 0.0;
                double z0[N];
#pragma unroll
                for (int k = 0; k < N; ++k) z0[k] = ST::z(ci + ST::STRIDE_V, k);
                recurrence();
#pragma unroll
                for (int k = 0; k < N; ++k) { // rotation of the previous step: columns i+2 (stored) and i+1 (carried on)
                    ST::zset(ci + 2 * ST::STRIDE_V, k, sp * z0[k] + cp * zc[k]);
                    zc[k] = cp * z0[k] - sp * zc[k];
                }
                if (!ok) { split_on_underflow(); break; }
                commit(eiN, diN);
            }
        }
        ST::eset(S.col(m), 0.0);
        if (underflow) continue;
        // flush the last rotation (columns l, l+1)
#pragma unroll
        for (int k = 0; k < N; ++k) {
            ST::zset(cl + ST::STRIDE_V, k, sp * zl[k] + cp * zc[k]);
            ST::zset(cl, k, cp * zl[k] - sp * zc[k]);
        }
        ST::dset(cl, dl - p);
        ST::eset(cl, g);
    }
    const bool converged = l >= N - 1;
    // ---- 4. V max(lambda, 0) V^T (eigenvalues scaled back by the exact power of two)
    double lam[N];
#ifdef IDP_QL_STATS
    g_ql_stats.negcount = 0;
#endif
#pragma unroll
    for (int i = 0; i < N; ++i) {
        const double x = ST::d(S.col(i));
        lam[i] = x > 0 ? x * unscale : 0.0;
#ifdef IDP_QL_STATS
        if (x < 0) ++g_ql_stats.negcount;
#endif
    }
#pragma unroll
    for (int r = 0; r < N; ++r) {
        double wr[N];
#pragma unroll
        for (int k = 0; k < N; ++k) wr[k] = ST::z(S.col(k), r) * lam[k];
#pragma unroll
        for (int c = r; c < N; ++c) {
            double s = 0;
#pragma unroll
            for (int k = 0; k < N; ++k) s += wr[k] * ST::z(S.col(k), c);
            a[SI<N>(r, c)] = s;
        }
    }
    return converged;
}

// ---- reduced-coordinate pieces ---------------------------------------------------------------------------------
// adds to the packed 9x9 H (blocks w=0, u=1, v=2):  kg g g^T + kq q q^T + kA hess(A) + kB hess(B)
// for d = A^2/B;  g = grad d, q = grad A - r grad B are returned in g9 / internally.
IDP_HD void triple_quotient_packed(const V3& w, const V3& u, const V3& v, double kbh, double kbg, double* g9, double* H)
{
    const V3 n = cross3(u, v);
    const double A = dot3(w, n), B = sqn3(n), iB = 1.0 / B, r = A * iB;
    const V3 Au = cross3(v, w), Av = cross3(w, u);
    const V3 Bu = 2.0 * cross3(v, n), Bv = 2.0 * cross3(n, u);
    const double gA[9] = {n.x, n.y, n.z, Au.x, Au.y, Au.z, Av.x, Av.y, Av.z};
    const double gB[9] = {0, 0, 0, Bu.x, Bu.y, Bu.z, Bv.x, Bv.y, Bv.z};
    double q[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) {
        g9[i] = 2.0 * r * gA[i] - r * r * gB[i];
        q[i] = gA[i] - r * gB[i];
    }
    const double kq = kbg * 2.0 * iB, kA = kbg * 2.0 * r, kB = -kbg * r * r;
#pragma unroll
    for (int i = 0; i < 9; ++i)
#pragma unroll
        for (int j = i; j < 9; ++j) H[SI<9>(i, j)] += kbh * g9[i] * g9[j] + kq * q[i] * q[j];
    const double uu = sqn3(u), vv = sqn3(v), uv = dot3(u, v);
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            const double dij = (i == j) ? 1.0 : 0.0;
            H[SI<9>(i, 3 + j)] += kA * (-skew3(v, i, j));
            H[SI<9>(i, 6 + j)] += kA * skew3(u, i, j);
            H[SI<9>(3 + i, 6 + j)] += kA * (-skew3(w, i, j)) + kB * (2.0 * (comp3(u, i) * comp3(v, j) - uv * dij) - 2.0 * skew3(n, i, j));
            if (j >= i) {
                H[SI<9>(3 + i, 3 + j)] += kB * 2.0 * (vv * dij - comp3(v, i) * comp3(v, j));
                H[SI<9>(6 + i, 6 + j)] += kB * 2.0 * (uu * dij - comp3(u, i) * comp3(u, j));
            }
        }
}

// N = |a x b|^2: value, gradient (ga, gb) and the three Hessian blocks as closures over (a, b, c = a x b)
struct CrossN2 {
    V3 a, b, c;
    double N, aa, bb, ab;
    V3 ga, gb;
    IDP_HD void init(const V3& a_, const V3& b_)
    {
        a = a_; b = b_;
        c = cross3(a, b);
        N = sqn3(c); aa = sqn3(a); bb = sqn3(b); ab = dot3(a, b);
        ga = 2.0 * cross3(b, c);
        gb = 2.0 * cross3(c, a);
    }
    IDP_HD double Haa(int i, int j) const { return 2.0 * (bb * (i == j ? 1.0 : 0.0) - comp3(b, i) * comp3(b, j)); }
    IDP_HD double Hbb(int i, int j) const { return 2.0 * (aa * (i == j ? 1.0 : 0.0) - comp3(a, i) * comp3(a, j)); }
    IDP_HD double Hab(int i, int j) const { return 2.0 * (comp3(a, i) * comp3(b, j) - ab * (i == j ? 1.0 : 0.0)) - 2.0 * skew3(c, i, j); }
};

// d = |a x b|^2 / |b|^2 in coordinates (a, b) placed at block offsets oa, ob (oa < ob) of a packed NxN matrix:
// H += kbh g g^T + kbg hess(d);  returns g (N-vector contributions written into gN at the two blocks)
template <int N>
IDP_HD void pe_quotient_packed(const V3& a, const V3& b, int oa, int ob, double kbh, double kbg, double* gN, double* H)
{
    CrossN2 cn;
    cn.init(a, b);
    const double B = cn.bb, iB = 1.0 / B, NB2 = cn.N * iB * iB;
    const V3 gBb = 2.0 * b;
    double g6[6];
    const double gNn[6] = {cn.ga.x, cn.ga.y, cn.ga.z, cn.gb.x, cn.gb.y, cn.gb.z};
    const double gBv[6] = {0, 0, 0, gBb.x, gBb.y, gBb.z};
#pragma unroll
    for (int i = 0; i < 6; ++i) g6[i] = gNn[i] * iB - NB2 * gBv[i];
#pragma unroll
    for (int i = 0; i < 3; ++i) { gN[oa + i] = g6[i]; gN[ob + i] = g6[3 + i]; }
#pragma unroll
    for (int i = 0; i < 6; ++i)
#pragma unroll
        for (int j = i; j < 6; ++j) {
            const int bi = i / 3, bj = j / 3, ii = i % 3, jj = j % 3;
            const double hn = (bi == 0 && bj == 0) ? cn.Haa(ii, jj) : ((bi == 1 && bj == 1) ? cn.Hbb(ii, jj) : cn.Hab(ii, jj));
            const double hb = (i >= 3 && i == j) ? 2.0 : 0.0;
            const double hd = hn * iB - (gNn[i] * gBv[j] + gBv[i] * gNn[j]) * (iB * iB) - NB2 * hb + (2.0 * NB2 * iB) * gBv[i] * gBv[j];
            const int R = (bi == 0 ? oa : ob) + ii, C = (bj == 0 ? oa : ob) + jj;
            H[SI<N>(R, C)] += kbh * g6[i] * g6[j] + kbg * hd;
        }
}

// M = T^T H T for a block matrix with scalar 3x3 coefficient matrix T (K x K), blocks of size 3: packed in / packed out
template <int K>
IDP_HD void congruence_blocks(const double (&T)[K][K], const double* H, double* M)
{
    constexpr int N = 3 * K;
#pragma unroll
    for (int k = 0; k < K; ++k)
#pragma unroll
        for (int l = k; l < K; ++l)
#pragma unroll
            for (int i = 0; i < 3; ++i)
#pragma unroll
                for (int j = (k == l ? i : 0); j < 3; ++j) {
                    double s = 0;
#pragma unroll
                    for (int m = 0; m < K; ++m)
#pragma unroll
                        for (int n = 0; n < K; ++n) {
                            const double cf = T[m][k] * T[n][l];
                            if (cf != 0.0) s += cf * H[SI<N>(3 * m + i, 3 * n + j)];
                        }
                    M[SI<N>(3 * k + i, 3 * l + j)] = s;
                }
}

// dense block (i,j) of Q M Q^T for Q = Hm (x) I3, Hm: NV x K coefficients; writes 9 doubles row-major
template <int NV, int K>
IDP_HD void expand_block(const double (&Hm)[NV][K], const double* M, int i, int j, double* out9)
{
    constexpr int N = 3 * K;
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int b = 0; b < 3; ++b) {
            double s = 0;
#pragma unroll
            for (int k = 0; k < K; ++k)
#pragma unroll
                for (int l = 0; l < K; ++l) s += (Hm[i][k] * Hm[j][l]) * M[SI<N>(3 * k + a, 3 * l + b)];
            out9[3 * a + b] = s;
        }
}

// ---- row evaluation ------------------------------------------------------------------------------------------------
// Outputs: E (scalar), g (3*nv), and the Hessian through the sink `emit(i, j, block9)` called for all nv^2 blocks.
// VS9 / VS6 are the QL work stores (QlStore) of the 9x9 / 6x6 projections.
struct RowOut {
    double E;
    double g[12];
    bool eigFail; // the PSD projection did not converge (QL iteration cap)
};

// PATH selects which kinds are compiled in: -1 all, 0 four-vertex kinds, 1 point-edge, 2 point-point
template <int PATH, class VS9, class VS6, class Emit>
IDP_HD bool row_eval(const RowDec& d, const V3* x, const V3* xr, double weight, double dHat2, double kappa, double xi2,
    bool projectSPD, bool wantH, VS9& V9, VS6& V6, RowOut& out, Emit& emit)
{
    out.eigFail = false;
    const double dist2 = row_dist2(d.kind, x[0], x[1], x[2], x[3]) - xi2;
    if (!(dist2 > 0)) return false;
    double b, bg, bh;
    barrier_all(dist2, dHat2, kappa, b, bg, bh);
    const double mu = (double)d.mult;

    if ((PATH < 0 || PATH == 2) && d.kind == K_PP) {
        // closed form: H = [[M,-M],[-M,M]], M = w mu (4 bh dd^T + 2 bg I); eigenvalues 2 w mu (4 bh |d|^2 + 2 bg), 2 w mu 2 bg (x2)
        const V3 dd = x[0] - x[1];
        out.E = b * (d.mult > 1 ? mu : 1.0) * weight;
        const double kg = mu * weight * bg * 2.0;
        out.g[0] = kg * dd.x; out.g[1] = kg * dd.y; out.g[2] = kg * dd.z;
        out.g[3] = -out.g[0]; out.g[4] = -out.g[1]; out.g[5] = -out.g[2];
        if (wantH) {
            const double s = weight * mu;
            const double l2 = sqn3(dd);
            double kd = s * 4.0 * bh, ki = s * 2.0 * bg; // M = kd dd^T + ki I
            if (projectSPD) {
                const double lam1 = kd * l2 + ki, lam2 = ki;
                if (fmin(lam1, lam2) < 0) {
                    const double l1 = lam1 > 0 ? lam1 : 0.0, lp = lam2 > 0 ? lam2 : 0.0;
                    // M+ = lp I + (l1 - lp) d d^T / |d|^2
                    kd = (l1 - lp) / l2;
                    ki = lp;
                }
            }
            double Mb[9], Nb[9];
            const double dv[3] = {dd.x, dd.y, dd.z};
#pragma unroll
            for (int i = 0; i < 3; ++i)
#pragma unroll
                for (int j = 0; j < 3; ++j) {
                    Mb[3 * i + j] = kd * dv[i] * dv[j] + (i == j ? ki : 0.0);
                    Nb[3 * i + j] = -Mb[3 * i + j];
                }
            emit(0, 0, Mb); emit(0, 1, Nb); emit(1, 1, Mb);
            if (emit.wants(1, 0)) emit(1, 0, Nb);
        }
        return true;
    }

    if ((PATH < 0 || PATH == 1) && d.kind == K_PE) {
        // reduced (w,u) = (p - e0, e1 - e0); orthonormal basis of the translation-free subspace of 3 points:
        // h1 = (1,-1,0)/sqrt2, h2 = (1,1,-2)/sqrt6
        const V3 w = x[0] - x[1], u = x[2] - x[1];
        double H6[21], g6[6];
#pragma unroll
        for (int i = 0; i < 21; ++i) H6[i] = 0;
        const double s = weight * mu;
        pe_quotient_packed<6>(w, u, 0, 3, wantH ? s * bh : 0.0, wantH ? s * bg : 0.0, g6, H6);
        out.E = b * (d.mult > 1 ? mu : 1.0) * weight;
        const double kg = s * bg;
        // c_w = (1,-1,0), c_u = (0,-1,1)
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            out.g[a] = kg * g6[a];
            out.g[6 + a] = kg * g6[3 + a];
            out.g[3 + a] = -kg * (g6[a] + g6[3 + a]);
        }
        if (wantH) {
            const double r2 = 1.4142135623730951, ir2 = 0.70710678118654752, r32 = 1.2247448713915890, ir6 = 0.40824829046386302;
            const double T[2][2] = {{r2, 0.0}, {ir2, -r32}};
            double M[21];
            congruence_blocks<2>(T, H6, M);
            if (projectSPD && !make_pd_ql<6>(M, V6)) out.eigFail = true;
            const double Hm[3][2] = {{ir2, ir6}, {-ir2, ir6}, {0.0, -2.0 * ir6}};
#pragma unroll
            for (int i = 0; i < 3; ++i)
#pragma unroll
                for (int j = 0; j < 3; ++j) {
                    if (!emit.wants(i, j)) continue;
                    double blk[9];
                    expand_block<3, 2>(Hm, M, i, j, blk);
                    emit(i, j, blk);
                }
        }
        return true;
    }

    if (!(PATH < 0 || PATH == 0) || d.kind == K_PP || d.kind == K_PE) return true;
    // ---- four-vertex kinds: reduced (w,u,v)
    const bool isPT = (d.kind == K_PT);
    const bool moll = (d.kind == K_EE_M || d.kind == K_PE_M || d.kind == K_PP_M);
    V3 w, u, v;
    if (isPT) { w = x[0] - x[1]; u = x[2] - x[1]; v = x[3] - x[1]; }
    else { w = x[2] - x[0]; u = x[1] - x[0]; v = x[3] - x[2]; }
    double H[45], g9[9];
#pragma unroll
    for (int i = 0; i < 45; ++i) H[i] = 0;
#pragma unroll
    for (int i = 0; i < 9; ++i) g9[i] = 0;
    double e = 1.0, qg = 0, qh = 0;
    CrossN2 cm;
    if (moll) {
        const double eps_x = ee_mollifier_threshold(xr[0], xr[1], xr[2], xr[3]);
        cm.init(u, v);
        if (cm.N < eps_x) {
            const double ie = 1.0 / eps_x, qq = cm.N * ie;
            e = (-qq + 2.0) * qq;
            qg = 2.0 * ie * (-ie * cm.N + 1.0);
            qh = -2.0 * ie * ie;
        }
    }
    const double We = weight * e;
    // distance part: H += We (bh g g^T + bg hess d), g9 = grad d
    if (d.kind == K_PT || d.kind == K_EE || d.kind == K_EE_M) triple_quotient_packed(w, u, v, wantH ? We * bh : 0.0, wantH ? We * bg : 0.0, g9, H);
    else if (d.kind == K_PE_M) pe_quotient_packed<9>(w, v, 0, 6, wantH ? We * bh : 0.0, wantH ? We * bg : 0.0, g9, H);
    else { // K_PP_M: d = |w|^2
        g9[0] = 2.0 * w.x; g9[1] = 2.0 * w.y; g9[2] = 2.0 * w.z;
        if (wantH) {
#pragma unroll
            for (int i = 0; i < 3; ++i)
#pragma unroll
                for (int j = i; j < 3; ++j) H[SI<9>(i, j)] += We * bh * g9[i] * g9[j] + (i == j ? We * bg * 2.0 : 0.0);
        }
    }
    double gr[9]; // reduced gradient of the row energy
    if (moll) {
        out.E = b * e * weight;
        const double gc[9] = {0, 0, 0, cm.ga.x, cm.ga.y, cm.ga.z, cm.gb.x, cm.gb.y, cm.gb.z};
#pragma unroll
        for (int i = 0; i < 9; ++i) gr[i] = weight * ((e * bg) * g9[i] + (b * qg) * gc[i]);
        if (wantH && (qg != 0.0 || qh != 0.0)) {
            // + w [ b (qg hess c + qh gc gc^T) + bg qg (g gc^T + gc g^T) ]
            const double k1 = weight * b * qg, k2 = weight * b * qh, k3 = weight * bg * qg;
#pragma unroll
            for (int i = 0; i < 9; ++i)
#pragma unroll
                for (int j = i; j < 9; ++j) H[SI<9>(i, j)] += k2 * gc[i] * gc[j] + k3 * (g9[i] * gc[j] + gc[i] * g9[j]);
#pragma unroll
            for (int i = 0; i < 3; ++i)
#pragma unroll
                for (int j = 0; j < 3; ++j) {
                    H[SI<9>(3 + i, 6 + j)] += k1 * cm.Hab(i, j);
                    if (j >= i) {
                        H[SI<9>(3 + i, 3 + j)] += k1 * cm.Haa(i, j);
                        H[SI<9>(6 + i, 6 + j)] += k1 * cm.Hbb(i, j);
                    }
                }
        }
    }
    else {
        out.E = b * weight;
#pragma unroll
        for (int i = 0; i < 9; ++i) gr[i] = (weight * bg) * g9[i];
    }
    // gradient embedding g_i = sum_m c[m][i] gr[m]
    if (isPT) {
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            out.g[a] = gr[a];
            out.g[6 + a] = gr[3 + a];
            out.g[9 + a] = gr[6 + a];
            out.g[3 + a] = -(gr[a] + gr[3 + a] + gr[6 + a]);
        }
    }
    else {
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            out.g[a] = -(gr[a] + gr[3 + a]);      // a0: -w -u
            out.g[3 + a] = gr[3 + a];             // a1: +u
            out.g[6 + a] = gr[a] - gr[6 + a];     // b0: +w -v
            out.g[9 + a] = gr[6 + a];             // b1: +v
        }
    }
    if (!wantH) return true;
    // Hadamard basis rows Hm[i][k] (k = 1..3) and T = c Hm
    double M[45];
    if (isPT) {
        const double T[3][3] = {{1, 0, 1}, {1, -1, 0}, {0, -1, 1}};
        congruence_blocks<3>(T, H, M);
    }
    else {
        const double T[3][3] = {{0, -1, -1}, {-1, 0, -1}, {-1, 0, 1}};
        congruence_blocks<3>(T, H, M);
    }
    if (projectSPD && !make_pd_ql<9>(M, V9)) out.eigFail = true;
    const double Hm[4][3] = {{0.5, 0.5, 0.5}, {-0.5, 0.5, -0.5}, {0.5, -0.5, -0.5}, {-0.5, -0.5, 0.5}};
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (!emit.wants(i, j)) continue;
            double blk[9];
            expand_block<4, 3>(Hm, M, i, j, blk);
            emit(i, j, blk);
        }
    return true;
}

// dense adaptor with the signature of row_EgH (pair_deriv.cuh): used by the host-side test shim
struct DenseEmit {
    double* H;
    int n;
    IDP_HD bool wants(int, int) const { return true; }
    IDP_HD void operator()(int i, int j, const double* blk)
    {
        for (int a = 0; a < 3; ++a)
            for (int b = 0; b < 3; ++b) H[(3 * i + a) * n + 3 * j + b] = blk[3 * a + b];
    }
};
IDP_HD bool row_EgH_lowrank(const RowDec& d, const V3* x, const V3* xr, double weight, double dHat2, double kappa,
    double xi2, bool projectSPD, double* E, double* g, double* H)
{
    double work[QlStore<9, 1>::WORDS];
    QlStore<9, 1> V9{work};
    QlStore<6, 1> V6{work};
    RowOut out;
    DenseEmit em{H, 3 * d.nv};
    const bool ok = row_eval<-1>(d, x, xr, weight, dHat2, kappa, xi2, projectSPD, H != nullptr, V9, V6, out, em);
    if (!ok || out.eigFail) return false;
    if (E) *E = out.E;
    if (g) for (int i = 0; i < 3 * d.nv; ++i) g[i] = out.g[i];
    return true;
}

} // namespace idp
