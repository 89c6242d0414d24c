// ORACLE — TEST INFRASTRUCTURE ONLY. Nothing under idp_b200/ may include, link or call this.
//
// CPU restatement of the per-pair math of the IPC contact hot path of ipc-sim/IDP ("JGSL").
// Parity status: PINNED AGAINST THE REFERENCE'S OWN CODE, with one stated exception. The reference ships no golden
// vectors, so the pins are produced here: its per-pair headers (Math/Distance/*.h, BARRIER.h, UTILS.h) AND
// its loops (FEM/IPC.h, Grid/SPATIAL_HASH.h) are compiled from /root/reference into oracle/_ref (oracle/ref_shim/) and this
// restatement is checked against them bit for bit / to 1e-12 (tests/test_golden.py, tests/test_ref_loops.py and the
// fixtures under tests/golden/). The exception: Eigen itself is not installed, so the reference code runs on the repo's
// Eigen subset; what is taken from Eigen semantics there (3-term reduction order, pivoted 2x2 LDLT, VectorXd::mean()
// order, the symmetric eigen-solver) remains "parity unpinned" at the Eigen boundary.
//
// All citations are relative to /root/reference/Library.
//
// Build rule: this header is compiled with -ffp-contract=off for the parity build so that every
// quantity feeding a comparison (voxel index, AABB test, distance type, d < dHat^2, ACCD) is a
// fixed sequence of IEEE-754 double operations that the CUDA path reproduces with --fmad=false.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <algorithm>
#include <limits>

namespace orc {

struct V3 {
    double x, y, z;
};
static inline V3 operator+(const V3& a, const V3& b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
static inline V3 operator-(const V3& a, const V3& b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
static inline V3 operator*(double s, const V3& a) { return {s * a.x, s * a.y, s * a.z}; }
static inline V3 operator/(const V3& a, double s) { return {a.x / s, a.y / s, a.z / s}; }
static inline V3 ld3(const double* p) { return {p[0], p[1], p[2]}; }

// Eigen evaluates a fixed-size 3-term reduction as x0 + (x1 + x2) (redux_novec_unroller splits
// [0,3) into [0,1) and [1,3)); SURVEY.md A.5.
static inline double dot(const V3& a, const V3& b) { return a.x * b.x + (a.y * b.y + a.z * b.z); }
static inline double sqn(const V3& a) { return a.x * a.x + (a.y * a.y + a.z * a.z); }
static inline V3 cross(const V3& a, const V3& b)
{
    return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
static inline V3 vmin(const V3& a, const V3& b) { return {std::min(a.x, b.x), std::min(a.y, b.y), std::min(a.z, b.z)}; }
static inline V3 vmax(const V3& a, const V3& b) { return {std::max(a.x, b.x), std::max(a.y, b.y), std::max(a.z, b.z)}; }

// ---------------------------------------------------------------------------------------------
// squared distances — Math/Distance/POINT_POINT.h:11-17, POINT_EDGE.h:12-25 (3-D branch),
// POINT_TRIANGLE.h:12-22, EDGE_EDGE.h:12-22
// ---------------------------------------------------------------------------------------------
static inline double dist2_pp(const V3& a, const V3& b) { return sqn(a - b); }
static inline double dist2_pe(const V3& p, const V3& e0, const V3& e1)
{
    return sqn(cross(e0 - p, e1 - p)) / sqn(e1 - e0);
}
static inline double dist2_pt(const V3& p, const V3& t0, const V3& t1, const V3& t2)
{
    const V3 b = cross(t1 - t0, t2 - t0);
    const double aTb = dot(p - t0, b);
    return aTb * aTb / sqn(b);
}
static inline double dist2_ee(const V3& a0, const V3& a1, const V3& b0, const V3& b1)
{
    const V3 b = cross(a1 - a0, b1 - b0);
    const double aTb = dot(b0 - a0, b);
    return aTb * aTb / sqn(b);
}

// ---------------------------------------------------------------------------------------------
// 2x2 normal-equation solve used by the point-triangle classifier. Transcription of Eigen 3.3's
// pivoted LDLT (ldlt_inplace<Lower>::unblocked + LDLT::_solve_impl) for a 2x2 matrix whose lower
// triangle is (a00, a10, a11). Only the signs / comparisons of the result are consumed
// (Math/Distance/DISTANCE_TYPE.h:46-47).
// ---------------------------------------------------------------------------------------------
static inline void ldlt2_solve(double a00, double a10, double a11, double b0, double b1, double& x0, double& x1)
{
    const bool piv = std::fabs(a11) > std::fabs(a00); // first maximum wins on ties
    double d0 = piv ? a11 : a00;
    double d1 = piv ? a00 : a11;
    double l10 = a10;
    if (std::fabs(d0) > 0.0) {
        l10 = a10 / d0;
    }
    d1 = d1 - l10 * (d0 * l10);
    double y0 = piv ? b1 : b0;
    double y1 = piv ? b0 : b1;
    y1 = y1 - l10 * y0;
    const double tol = std::numeric_limits<double>::min();
    y0 = (std::fabs(d0) > tol) ? y0 / d0 : 0.0;
    y1 = (std::fabs(d1) > tol) ? y1 / d1 : 0.0;
    y0 = y0 - l10 * y1;
    x0 = piv ? y1 : y0;
    x1 = piv ? y0 : y1;
}

// one edge test of Point_Triangle_Distance_Type: basis rows (e, e x n), solve (B B^T) param = B (p - o)
static inline void pt_edge_param(const V3& e, const V3& n, const V3& po, double& u, double& v)
{
    const V3 r1 = cross(e, n);
    const double a00 = dot(e, e), a10 = dot(r1, e), a11 = dot(r1, r1);
    ldlt2_solve(a00, a10, a11, dot(e, po), dot(r1, po), u, v);
}

// Math/Distance/DISTANCE_TYPE.h:31-82 — returns 0,1,2 (PP t0/t1/t2), 3,4,5 (PE t0t1/t1t2/t2t0), 6 (PT)
static inline int pt_type(const V3& p, const V3& t0, const V3& t1, const V3& t2)
{
    const V3 e0 = t1 - t0;
    const V3 n = cross(e0, t2 - t0);
    double u0, v0, u1, v1, u2, v2;
    pt_edge_param(e0, n, p - t0, u0, v0);
    if (u0 > 0.0 && u0 < 1.0 && v0 >= 0.0) return 3;
    pt_edge_param(t2 - t1, n, p - t1, u1, v1);
    if (u1 > 0.0 && u1 < 1.0 && v1 >= 0.0) return 4;
    pt_edge_param(t0 - t2, n, p - t2, u2, v2);
    if (u2 > 0.0 && u2 < 1.0 && v2 >= 0.0) return 5;
    if (u0 <= 0.0 && u2 >= 1.0) return 0;
    if (u1 <= 0.0 && u0 >= 1.0) return 1;
    if (u2 <= 0.0 && u1 >= 1.0) return 2;
    return 6;
}

// Math/Distance/DISTANCE_TYPE.h:86-164 — returns 0..8
static inline int ee_type(const V3& ea0, const V3& ea1, const V3& eb0, const V3& eb1)
{
    const V3 u = ea1 - ea0, v = eb1 - eb0, w = ea0 - eb0;
    const double a = sqn(u), b = dot(u, v), c = sqn(v), d = dot(u, w), e = dot(v, w);
    const double D = a * c - b * b;
    double tD = D, tN;
    int def = 8;
    const double sN = b * e - c * d;
    if (sN <= 0.0) {
        tN = e; tD = c; def = 2;
    }
    else if (sN >= D) {
        tN = e + b; tD = c; def = 5;
    }
    else {
        tN = a * e - b * d;
        if (tN > 0.0 && tN < tD) {
            const V3 uxv = cross(u, v);
            if (dot(uxv, w) == 0.0 || sqn(uxv) < 1.0e-20 * a * c) {
                if (sN < D / 2) { tN = e; tD = c; def = 2; }
                else { tN = e + b; tD = c; def = 5; }
            }
        }
    }
    if (tN <= 0.0) {
        if (-d <= 0.0) return 0;
        else if (-d >= a) return 3;
        else return 6;
    }
    else if (tN >= tD) {
        if ((-d + b) <= 0.0) return 1;
        else if ((-d + b) >= a) return 4;
        else return 7;
    }
    return def;
}

// Math/Distance/DISTANCE_UNCLASSIFIED.h:16-62, 65-121
static inline double dist2_pt_unclassified(const V3& p, const V3& t0, const V3& t1, const V3& t2)
{
    switch (pt_type(p, t0, t1, t2)) {
    case 0: return dist2_pp(p, t0);
    case 1: return dist2_pp(p, t1);
    case 2: return dist2_pp(p, t2);
    case 3: return dist2_pe(p, t0, t1);
    case 4: return dist2_pe(p, t1, t2);
    case 5: return dist2_pe(p, t2, t0);
    default: return dist2_pt(p, t0, t1, t2);
    }
}
static inline double dist2_ee_unclassified(const V3& a0, const V3& a1, const V3& b0, const V3& b1)
{
    switch (ee_type(a0, a1, b0, b1)) {
    case 0: return dist2_pp(a0, b0);
    case 1: return dist2_pp(a0, b1);
    case 2: return dist2_pe(a0, b0, b1);
    case 3: return dist2_pp(a1, b0);
    case 4: return dist2_pp(a1, b1);
    case 5: return dist2_pe(a1, b0, b1);
    case 6: return dist2_pe(b0, a0, a1);
    case 7: return dist2_pe(b1, a0, a1);
    default: return dist2_ee(a0, a1, b0, b1);
    }
}

// ---------------------------------------------------------------------------------------------
// AABB broad-phase predicates — Math/Distance/CCD.h:149-185 (static), 187-235 (swept, full step)
// reject iff any per-axis gap is strictly greater than dist
// ---------------------------------------------------------------------------------------------
static inline bool any_gt(const V3& a, double d) { return a.x > d || a.y > d || a.z > d; }
static inline bool pt_cd_broadphase(const V3& p, const V3& t0, const V3& t1, const V3& t2, double dist)
{
    const V3 mx = vmax(vmax(t0, t1), t2), mn = vmin(vmin(t0, t1), t2);
    return !(any_gt(p - mx, dist) || any_gt(mn - p, dist));
}
static inline bool ee_cd_broadphase(const V3& a0, const V3& a1, const V3& b0, const V3& b1, double dist)
{
    const V3 mxa = vmax(a0, a1), mna = vmin(a0, a1), mxb = vmax(b0, b1), mnb = vmin(b0, b1);
    return !(any_gt(mna - mxb, dist) || any_gt(mnb - mxa, dist));
}
static inline bool pt_ccd_broadphase(const V3& p, const V3& t0, const V3& t1, const V3& t2,
    const V3& dp, const V3& dt0, const V3& dt1, const V3& dt2, double dist)
{
    const V3 pe = p + dp;
    const V3 mxp = vmax(p, pe), mnp = vmin(p, pe);
    const V3 t0e = t0 + dt0, t1e = t1 + dt1, t2e = t2 + dt2;
    const V3 mxt = vmax(vmax(vmax(vmax(vmax(t0, t1), t2), t0e), t1e), t2e);
    const V3 mnt = vmin(vmin(vmin(vmin(vmin(t0, t1), t2), t0e), t1e), t2e);
    return !(any_gt(mnp - mxt, dist) || any_gt(mnt - mxp, dist));
}
static inline bool ee_ccd_broadphase(const V3& a0, const V3& a1, const V3& b0, const V3& b1,
    const V3& da0, const V3& da1, const V3& db0, const V3& db1, double dist)
{
    const V3 a0e = a0 + da0, a1e = a1 + da1, b0e = b0 + db0, b1e = b1 + db1;
    const V3 mxa = vmax(vmax(vmax(a0, a1), a0e), a1e), mna = vmin(vmin(vmin(a0, a1), a0e), a1e);
    const V3 mxb = vmax(vmax(vmax(b0, b1), b0e), b1e), mnb = vmin(vmin(vmin(b0, b1), b0e), b1e);
    return !(any_gt(mna - mxb, dist) || any_gt(mnb - mxa, dist));
}

// ---------------------------------------------------------------------------------------------
// mollifier — Math/Distance/EDGE_EDGE_MOLLIFIER.h:10-18, 441-458, 583-591
// ---------------------------------------------------------------------------------------------
static inline double ee_cross_norm2(const V3& a0, const V3& a1, const V3& b0, const V3& b1)
{
    return sqn(cross(a1 - a0, b1 - b0));
}
static inline double ee_mollifier_threshold(const V3& a0r, const V3& a1r, const V3& b0r, const V3& b1r)
{
    return 1.0e-3 * sqn(a0r - a1r) * sqn(b0r - b1r);
}
static inline double eem(double c, double eps) { const double q = c / eps; return (-q + 2.0) * q; }
static inline double eem_g(double c, double eps) { const double i = 1.0 / eps; return 2.0 * i * (-i * c + 1.0); }
static inline double eem_h(double, double eps) { return -2.0 / (eps * eps); }

// ---------------------------------------------------------------------------------------------
// barrier on squared distance — Math/BARRIER.h:10-62, elastic=false branch
// ---------------------------------------------------------------------------------------------
static inline double barrier(double d, double dHat2, double kappa) { return -kappa * (d - dHat2) * (d - dHat2) * std::log(d / dHat2); }
static inline double barrier_g(double d, double dHat2, double kappa)
{
    const double t2 = d - dHat2;
    return kappa * (t2 * std::log(d / dHat2) * -2.0 - (t2 * t2) / d);
}
static inline double barrier_h(double d, double dHat2, double kappa)
{
    const double t2 = d - dHat2;
    return kappa * ((std::log(d / dHat2) * -2.0 - t2 * 4.0 / d) + 1.0 / (d * d) * (t2 * t2));
}

// ---------------------------------------------------------------------------------------------
// additive CCD — Math/Distance/CCD.h:279-328 (PT), 330-395 (EE). Returns true and writes toc if a
// time of impact bound <= the incoming toc was found. *iters counts loop trips (for the
// measurement model); max_iter guards the reference's unbounded loop.
// ---------------------------------------------------------------------------------------------
static inline bool accd_pt(V3 p, V3 t0, V3 t1, V3 t2, V3 dp, V3 dt0, V3 dt1, V3 dt2,
    double eta, double thickness, double& toc, long* iters = nullptr, long max_iter = 100000000L)
{
    const V3 mov = (((dt0 + dt1) + dt2) + dp) / 4.0;
    dt0 = dt0 - mov; dt1 = dt1 - mov; dt2 = dt2 - mov; dp = dp - mov;
    const double m2 = std::max(std::max(sqn(dt0), sqn(dt1)), sqn(dt2));
    const double maxDispMag = std::sqrt(sqn(dp)) + std::sqrt(m2);
    if (maxDispMag == 0) return false;
    const double xi2 = thickness * thickness;
    double dist2_cur = dist2_pt_unclassified(p, t0, t1, t2);
    double dist_cur = std::sqrt(dist2_cur);
    const double gap = eta * (dist2_cur - xi2) / (dist_cur + thickness);
    const double toc_prev = toc;
    toc = 0;
    long it = 0;
    while (true) {
        if (++it > max_iter) { if (iters) *iters += it; return false; }
        const double tl = (1 - eta) * (dist2_cur - xi2) / ((dist_cur + thickness) * maxDispMag);
        p = p + tl * dp; t0 = t0 + tl * dt0; t1 = t1 + tl * dt1; t2 = t2 + tl * dt2;
        dist2_cur = dist2_pt_unclassified(p, t0, t1, t2);
        dist_cur = std::sqrt(dist2_cur);
        if (toc != 0 && ((dist2_cur - xi2) / (dist_cur + thickness) < gap)) break;
        toc += tl;
        if (toc > toc_prev) { if (iters) *iters += it; return false; }
    }
    if (iters) *iters += it;
    return true;
}

static inline double ee_min_endpoint_dist2(const V3& a0, const V3& a1, const V3& b0, const V3& b1)
{
    return std::min(std::min(sqn(a0 - b0), sqn(a0 - b1)), std::min(sqn(a1 - b0), sqn(a1 - b1)));
}

static inline bool accd_ee(V3 a0, V3 a1, V3 b0, V3 b1, V3 da0, V3 da1, V3 db0, V3 db1,
    double eta, double thickness, double& toc, long* iters = nullptr, long max_iter = 100000000L)
{
    const V3 mov = (((da0 + da1) + db0) + db1) / 4.0;
    da0 = da0 - mov; da1 = da1 - mov; db0 = db0 - mov; db1 = db1 - mov;
    const double maxDispMag = std::sqrt(std::max(sqn(da0), sqn(da1))) + std::sqrt(std::max(sqn(db0), sqn(db1)));
    if (maxDispMag == 0) return false;
    const double xi2 = thickness * thickness;
    double dist2_cur = dist2_ee_unclassified(a0, a1, b0, b1);
    double dFunc = dist2_cur - xi2;
    if (dFunc <= 0) {
        dist2_cur = ee_min_endpoint_dist2(a0, a1, b0, b1);
        dFunc = dist2_cur - xi2;
    }
    double dist_cur = std::sqrt(dist2_cur);
    const double gap = eta * dFunc / (dist_cur + thickness);
    const double toc_prev = toc;
    toc = 0;
    long it = 0;
    while (true) {
        if (++it > max_iter) { if (iters) *iters += it; return false; }
        const double tl = (1 - eta) * dFunc / ((dist_cur + thickness) * maxDispMag);
        a0 = a0 + tl * da0; a1 = a1 + tl * da1; b0 = b0 + tl * db0; b1 = b1 + tl * db1;
        dist2_cur = dist2_ee_unclassified(a0, a1, b0, b1);
        dFunc = dist2_cur - xi2;
        if (dFunc <= 0) {
            dist2_cur = ee_min_endpoint_dist2(a0, a1, b0, b1);
            dFunc = dist2_cur - xi2;
        }
        dist_cur = std::sqrt(dist2_cur);
        if (toc != 0 && (dFunc / (dist_cur + thickness) < gap)) break;
        toc += tl;
        if (toc > toc_prev) { if (iters) *iters += it; return false; }
    }
    if (iters) *iters += it;
    return true;
}

// ---------------------------------------------------------------------------------------------
// deterministic mean used for the voxel size (Grid/SPATIAL_HASH.h:66, 463: Eigen VectorXd::mean(),
// whose packetised summation order is not recoverable offline -> parity unpinned). Definition used
// on both sides of the parity check: perfectly balanced adjacent-pair binary tree over the array
// zero-padded to the next power of two, divided by n.
// ---------------------------------------------------------------------------------------------
static inline double tree_sum_rec(const double* a, long n, long lo, long len)
{
    if (lo >= n) return 0.0;
    if (len == 1) return a[lo];
    return tree_sum_rec(a, n, lo, len / 2) + tree_sum_rec(a, n, lo + len / 2, len / 2);
}
static inline double tree_sum(const double* a, long n)
{
    if (n <= 0) return 0.0;
    long P = 1;
    while (P < n) P <<= 1;
    return tree_sum_rec(a, n, 0, P);
}

} // namespace orc
