// Per-pair arithmetic whose results feed comparisons (AABB tests, distance-type classification,
// d < dHat^2, additive CCD). Every function here is a fixed sequence of IEEE-754 double operations;
// translation units that include this header for those purposes are compiled with --fmad=false so that
// ptxas never contracts a*b+c, which is what makes the constraint / candidate sets bit-exact.
//
// Reference semantics (citations relative to /root/reference/Library):
//   Math/Distance/POINT_POINT.h:11-17, POINT_EDGE.h:12-25, POINT_TRIANGLE.h:12-22, EDGE_EDGE.h:12-22,
//   DISTANCE_TYPE.h:31-164, DISTANCE_UNCLASSIFIED.h:16-121, CCD.h:149-235, 279-395,
//   EDGE_EDGE_MOLLIFIER.h:10-18, 583-591.
// Fixed-size reductions use Eigen's association x0 + (x1 + x2).
#pragma once
#include <math.h>
#include <float.h>

#if defined(__CUDACC__)
#define IDP_HD __host__ __device__ __forceinline__
#else
#define IDP_HD inline
#endif

namespace idp {

struct V3 {
    double x, y, z;
};
IDP_HD V3 mk3(double x, double y, double z) { V3 r; r.x = x; r.y = y; r.z = z; return r; }
IDP_HD V3 operator+(const V3& a, const V3& b) { return mk3(a.x + b.x, a.y + b.y, a.z + b.z); }
IDP_HD V3 operator-(const V3& a, const V3& b) { return mk3(a.x - b.x, a.y - b.y, a.z - b.z); }
IDP_HD V3 operator*(double s, const V3& a) { return mk3(s * a.x, s * a.y, s * a.z); }
IDP_HD V3 operator/(const V3& a, double s) { return mk3(a.x / s, a.y / s, a.z / s); }
IDP_HD double dot3(const V3& a, const V3& b) { return a.x * b.x + (a.y * b.y + a.z * b.z); }
IDP_HD double sqn3(const V3& a) { return a.x * a.x + (a.y * a.y + a.z * a.z); }
IDP_HD V3 cross3(const V3& a, const V3& b)
{
    return mk3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
IDP_HD V3 min3(const V3& a, const V3& b) { return mk3(fmin(a.x, b.x), fmin(a.y, b.y), fmin(a.z, b.z)); }
IDP_HD V3 max3(const V3& a, const V3& b) { return mk3(fmax(a.x, b.x), fmax(a.y, b.y), fmax(a.z, b.z)); }

// ---- squared distances -------------------------------------------------------------------------
IDP_HD double dist2_pp(const V3& a, const V3& b) { return sqn3(a - b); }
IDP_HD double dist2_pe(const V3& p, const V3& e0, const V3& e1) { return sqn3(cross3(e0 - p, e1 - p)) / sqn3(e1 - e0); }
IDP_HD double dist2_pt(const V3& p, const V3& t0, const V3& t1, const V3& t2)
{
    const V3 b = cross3(t1 - t0, t2 - t0);
    const double aTb = dot3(p - t0, b);
    return aTb * aTb / sqn3(b);
}
IDP_HD double dist2_ee(const V3& a0, const V3& a1, const V3& b0, const V3& b1)
{
    const V3 b = cross3(a1 - a0, b1 - b0);
    const double aTb = dot3(b0 - a0, b);
    return aTb * aTb / sqn3(b);
}

// ---- pivoted 2x2 LDLT solve (Eigen 3.3 LDLT semantics; DISTANCE_TYPE.h:46) -----------------------
IDP_HD void ldlt2_solve(double a00, double a10, double a11, double b0, double b1, double& x0, double& x1)
{
    const bool piv = fabs(a11) > fabs(a00);
    const double d0 = piv ? a11 : a00;
    double d1 = piv ? a00 : a11;
    double l10 = a10;
    if (fabs(d0) > 0.0) l10 = a10 / d0;
    d1 = d1 - l10 * (d0 * l10);
    double y0 = piv ? b1 : b0;
    double y1 = piv ? b0 : b1;
    y1 = y1 - l10 * y0;
    y0 = (fabs(d0) > DBL_MIN) ? y0 / d0 : 0.0;
    y1 = (fabs(d1) > DBL_MIN) ? y1 / d1 : 0.0;
    y0 = y0 - l10 * y1;
    x0 = piv ? y1 : y0;
    x1 = piv ? y0 : y1;
}
IDP_HD void pt_edge_param(const V3& e, const V3& n, const V3& po, double& u, double& v)
{
    const V3 r1 = cross3(e, n);
    ldlt2_solve(dot3(e, e), dot3(r1, e), dot3(r1, r1), dot3(e, po), dot3(r1, po), u, v);
}

// 0,1,2: PP with t0/t1/t2; 3,4,5: PE with t0t1/t1t2/t2t0; 6: PT   (DISTANCE_TYPE.h:31-82)
IDP_HD int pt_type(const V3& p, const V3& t0, const V3& t1, const V3& t2)
{
    const V3 e0 = t1 - t0;
    const V3 n = cross3(e0, t2 - t0);
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

// 0..8 (DISTANCE_TYPE.h:86-164; table in SURVEY.md C.2)
IDP_HD int ee_type(const V3& ea0, const V3& ea1, const V3& eb0, const V3& eb1)
{
    const V3 u = ea1 - ea0, v = eb1 - eb0, w = ea0 - eb0;
    const double a = sqn3(u), b = dot3(u, v), c = sqn3(v), d = dot3(u, w), e = dot3(v, w);
    const double D = a * c - b * b;
    double tD = D, tN;
    int def = 8;
    const double sN = b * e - c * d;
    if (sN <= 0.0) { tN = e; tD = c; def = 2; }
    else if (sN >= D) { tN = e + b; tD = c; def = 5; }
    else {
        tN = a * e - b * d;
        if (tN > 0.0 && tN < tD) {
            const V3 uxv = cross3(u, v);
            if (dot3(uxv, w) == 0.0 || sqn3(uxv) < 1.0e-20 * a * c) {
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

IDP_HD double dist2_pt_by_type(int t, const V3& p, const V3& t0, const V3& t1, const V3& t2)
{
    switch (t) {
    case 0: return dist2_pp(p, t0);
    case 1: return dist2_pp(p, t1);
    case 2: return dist2_pp(p, t2);
    case 3: return dist2_pe(p, t0, t1);
    case 4: return dist2_pe(p, t1, t2);
    case 5: return dist2_pe(p, t2, t0);
    default: return dist2_pt(p, t0, t1, t2);
    }
}
IDP_HD double dist2_ee_by_type(int t, const V3& a0, const V3& a1, const V3& b0, const V3& b1)
{
    switch (t) {
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
IDP_HD double dist2_pt_unclassified(const V3& p, const V3& t0, const V3& t1, const V3& t2)
{
    return dist2_pt_by_type(pt_type(p, t0, t1, t2), p, t0, t1, t2);
}
IDP_HD double dist2_ee_unclassified(const V3& a0, const V3& a1, const V3& b0, const V3& b1)
{
    return dist2_ee_by_type(ee_type(a0, a1, b0, b1), a0, a1, b0, b1);
}

// ---- AABB gap predicate (CCD.h:149-235): reject iff any per-axis gap > dist (strict) -------------
IDP_HD bool aabb_gap_ok(const V3& alo, const V3& ahi, const V3& blo, const V3& bhi, double dist)
{
    return !((alo.x - bhi.x > dist) || (alo.y - bhi.y > dist) || (alo.z - bhi.z > dist) ||
             (blo.x - ahi.x > dist) || (blo.y - ahi.y > dist) || (blo.z - ahi.z > dist));
}

// ---- mollifier scalars ------------------------------------------------------------------------------
IDP_HD double ee_cross_norm2(const V3& a0, const V3& a1, const V3& b0, const V3& b1) { return sqn3(cross3(a1 - a0, b1 - b0)); }
IDP_HD double ee_mollifier_threshold(const V3& a0r, const V3& a1r, const V3& b0r, const V3& b1r)
{
    return 1.0e-3 * sqn3(a0r - a1r) * sqn3(b0r - b1r);
}

// ---- additive CCD (CCD.h:279-395) ---------------------------------------------------------------------
// `bound` is read before every comparison so that a caller may pass a location other threads tighten
// concurrently (monotone). Returns: 1 hit (toc written), 0 no hit, -1 iteration cap reached.
#ifndef IDP_ACCD_MAX_ITER
#define IDP_ACCD_MAX_ITER 1000000
#endif
template <class BoundFn>
IDP_HD int accd_pt(V3 p, V3 t0, V3 t1, V3 t2, V3 dp, V3 dt0, V3 dt1, V3 dt2, double eta, double xi, BoundFn bound,
    double& toc_out, int& iters)
{
    const V3 mov = (((dt0 + dt1) + dt2) + dp) / 4.0;
    dt0 = dt0 - mov; dt1 = dt1 - mov; dt2 = dt2 - mov; dp = dp - mov;
    const double m2 = fmax(fmax(sqn3(dt0), sqn3(dt1)), sqn3(dt2));
    const double maxDispMag = sqrt(sqn3(dp)) + sqrt(m2);
    if (maxDispMag == 0) return 0;
    const double xi2 = xi * xi;
    double dist2_cur = dist2_pt_unclassified(p, t0, t1, t2);
    double dist_cur = sqrt(dist2_cur);
    const double gap = eta * (dist2_cur - xi2) / (dist_cur + xi);
    double toc = 0;
    for (int it = 0; it < IDP_ACCD_MAX_ITER; ++it) {
        ++iters;
        const double tl = (1 - eta) * (dist2_cur - xi2) / ((dist_cur + xi) * maxDispMag);
        p = p + tl * dp; t0 = t0 + tl * dt0; t1 = t1 + tl * dt1; t2 = t2 + tl * dt2;
        dist2_cur = dist2_pt_unclassified(p, t0, t1, t2);
        dist_cur = sqrt(dist2_cur);
        if (toc != 0 && ((dist2_cur - xi2) / (dist_cur + xi) < gap)) { toc_out = toc; return 1; }
        toc += tl;
        if (toc > bound()) return 0;
    }
    return -1;
}
IDP_HD double ee_min_endpoint_dist2(const V3& a0, const V3& a1, const V3& b0, const V3& b1)
{
    return fmin(fmin(sqn3(a0 - b0), sqn3(a0 - b1)), fmin(sqn3(a1 - b0), sqn3(a1 - b1)));
}
template <class BoundFn>
IDP_HD int accd_ee(V3 a0, V3 a1, V3 b0, V3 b1, V3 da0, V3 da1, V3 db0, V3 db1, double eta, double xi, BoundFn bound,
    double& toc_out, int& iters)
{
    const V3 mov = (((da0 + da1) + db0) + db1) / 4.0;
    da0 = da0 - mov; da1 = da1 - mov; db0 = db0 - mov; db1 = db1 - mov;
    const double maxDispMag = sqrt(fmax(sqn3(da0), sqn3(da1))) + sqrt(fmax(sqn3(db0), sqn3(db1)));
    if (maxDispMag == 0) return 0;
    const double xi2 = xi * xi;
    double dist2_cur = dist2_ee_unclassified(a0, a1, b0, b1);
    double dFunc = dist2_cur - xi2;
    if (dFunc <= 0) {
        dist2_cur = ee_min_endpoint_dist2(a0, a1, b0, b1);
        dFunc = dist2_cur - xi2;
    }
    double dist_cur = sqrt(dist2_cur);
    const double gap = eta * dFunc / (dist_cur + xi);
    double toc = 0;
    for (int it = 0; it < IDP_ACCD_MAX_ITER; ++it) {
        ++iters;
        const double tl = (1 - eta) * dFunc / ((dist_cur + xi) * maxDispMag);
        a0 = a0 + tl * da0; a1 = a1 + tl * da1; b0 = b0 + tl * db0; b1 = b1 + tl * db1;
        dist2_cur = dist2_ee_unclassified(a0, a1, b0, b1);
        dFunc = dist2_cur - xi2;
        if (dFunc <= 0) {
            dist2_cur = ee_min_endpoint_dist2(a0, a1, b0, b1);
            dFunc = dist2_cur - xi2;
        }
        dist_cur = sqrt(dist2_cur);
        if (toc != 0 && (dFunc / (dist_cur + xi) < gap)) { toc_out = toc; return 1; }
        toc += tl;
        if (toc > bound()) return 0;
    }
    return -1;
}

// ---- constraint-row kinds --------------------------------------------------------------
enum RowKind { K_EE = 0, K_EE_M = 1, K_PE_M = 2, K_PP_M = 3, K_PT = 4, K_PE = 5, K_PP = 6 };
struct RowDec {
    int kind;
    int v[4]; // stencil vertices in g/H DOF order
    int nv;
    int mult;
};
IDP_HD RowDec decode_row(int r0, int r1, int r2, int r3)
{
    RowDec d;
    d.mult = 1;
    if (r0 >= 0) {
        d.nv = 4;
        if (r3 >= 0 && r2 >= 0) { d.kind = K_EE; d.v[0] = r0; d.v[1] = r1; d.v[2] = r2; d.v[3] = r3; }
        else if (r3 >= 0) { d.kind = K_EE_M; d.v[0] = r0; d.v[1] = r1; d.v[2] = -r2 - 1; d.v[3] = r3; }
        else if (r2 >= 0) { d.kind = K_PE_M; d.v[0] = r0; d.v[1] = -r3 - 1; d.v[2] = r1; d.v[3] = r2; }
        else { d.kind = K_PP_M; d.v[0] = r0; d.v[1] = -r2 - 1; d.v[2] = r1; d.v[3] = -r3 - 1; }
    }
    else {
        d.v[0] = -r0 - 1; d.v[1] = r1; d.v[2] = r2; d.v[3] = r3;
        if (r3 >= 0) { d.kind = K_PT; d.nv = 4; }
        else if (r2 >= 0) { d.kind = K_PE; d.nv = 3; d.mult = -r3; d.v[3] = d.v[0]; }
        else { d.kind = K_PP; d.nv = 2; d.mult = -r3; d.v[2] = d.v[0]; d.v[3] = d.v[0]; }
    }
    return d;
}
// distance of a decoded row given its (up to) four stencil positions
IDP_HD double row_dist2(int kind, const V3& x0, const V3& x1, const V3& x2, const V3& x3)
{
    switch (kind) {
    case K_EE: case K_EE_M: return dist2_ee(x0, x1, x2, x3);
    case K_PE_M: return dist2_pe(x0, x2, x3);
    case K_PP_M: return dist2_pp(x0, x2);
    case K_PT: return dist2_pt(x0, x1, x2, x3);
    case K_PE: return dist2_pe(x0, x1, x2);
    default: return dist2_pp(x0, x1);
    }
}

} // namespace idp
