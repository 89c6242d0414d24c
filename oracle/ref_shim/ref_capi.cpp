// ORACLE / TEST INFRASTRUCTURE: compiles the REFERENCE's own per-pair math from where it lies under
// /root/reference/Library (Math/Distance/*.h, Math/BARRIER.h, Math/UTILS.h) against the Eigen / pybind11 stubs in
// include/, and exposes it to ctypes. No reference source is copied into this repository.
#include <Math/Distance/CCD.h>            // pulls DISTANCE_TYPE.h, DISTANCE_UNCLASSIFIED.h, POINT_*.h, EDGE_EDGE.h
#include <Math/Distance/EDGE_EDGE_MOLLIFIER.h>
#include <Math/BARRIER.h>
#include <Math/UTILS.h>
#include <Math/DIHEDRAL_ANGLE.h>         // the hinge angle, its gradient and Hessian (used by FEM/Shell/BENDING.h)

using namespace JGSL;
typedef Eigen::Matrix<double, 3, 1> V3d;
static V3d l3(const double* p) { return V3d(p[0], p[1], p[2]); }

extern "C" {
int ref_pt_type(const double* x) { return Point_Triangle_Distance_Type(l3(x), l3(x + 3), l3(x + 6), l3(x + 9)); }
int ref_ee_type(const double* x) { return Edge_Edge_Distance_Type(l3(x), l3(x + 3), l3(x + 6), l3(x + 9)); }
double ref_dist2(int kind, const double* x) // same kind ids as orc_dist2
{
    double d = 0;
    switch (kind) {
    case 0: Point_Point_Distance(l3(x), l3(x + 3), d); break;
    case 1: Point_Edge_Distance(l3(x), l3(x + 3), l3(x + 6), d); break;
    case 2: Point_Triangle_Distance(l3(x), l3(x + 3), l3(x + 6), l3(x + 9), d); break;
    case 3: Edge_Edge_Distance(l3(x), l3(x + 3), l3(x + 6), l3(x + 9), d); break;
    case 4: Point_Triangle_Distance_Unclassified(l3(x), l3(x + 3), l3(x + 6), l3(x + 9), d); break;
    case 5: Edge_Edge_Distance_Unclassified(l3(x), l3(x + 3), l3(x + 6), l3(x + 9), d); break;
    default: Edge_Edge_Cross_Norm2(l3(x), l3(x + 3), l3(x + 6), l3(x + 9), d); break;
    }
    return d;
}
void ref_grad_hess(int kind, const double* x, double* g, double* H) // H column-major as the reference stores it
{
    switch (kind) {
    case 0: { Eigen::Matrix<double, 6, 1> gg; Eigen::Matrix<double, 6, 6> HH; Point_Point_Distance_Gradient(l3(x), l3(x + 3), gg); Point_Point_Distance_Hessian(l3(x), l3(x + 3), HH); memcpy(g, gg.data(), 48); memcpy(H, HH.data(), 288); break; }
    case 1: { Eigen::Matrix<double, 9, 1> gg; Eigen::Matrix<double, 9, 9> HH; Point_Edge_Distance_Gradient(l3(x), l3(x + 3), l3(x + 6), gg); Point_Edge_Distance_Hessian(l3(x), l3(x + 3), l3(x + 6), HH); memcpy(g, gg.data(), 72); memcpy(H, HH.data(), 648); break; }
    case 2: { Eigen::Matrix<double, 12, 1> gg; Eigen::Matrix<double, 12, 12> HH; Point_Triangle_Distance_Gradient(l3(x), l3(x + 3), l3(x + 6), l3(x + 9), gg); Point_Triangle_Distance_Hessian(l3(x), l3(x + 3), l3(x + 6), l3(x + 9), HH); memcpy(g, gg.data(), 96); memcpy(H, HH.data(), 1152); break; }
    case 3: { Eigen::Matrix<double, 12, 1> gg; Eigen::Matrix<double, 12, 12> HH; Edge_Edge_Distance_Gradient(l3(x), l3(x + 3), l3(x + 6), l3(x + 9), gg); Edge_Edge_Distance_Hessian(l3(x), l3(x + 3), l3(x + 6), l3(x + 9), HH); memcpy(g, gg.data(), 96); memcpy(H, HH.data(), 1152); break; }
    default: { Eigen::Matrix<double, 12, 1> gg; Eigen::Matrix<double, 12, 12> HH; Edge_Edge_Cross_Norm2_Gradient(l3(x), l3(x + 3), l3(x + 6), l3(x + 9), gg); Edge_Edge_Cross_Norm2_Hessian(l3(x), l3(x + 3), l3(x + 6), l3(x + 9), HH); memcpy(g, gg.data(), 96); memcpy(H, HH.data(), 1152); break; }
    }
}
void ref_mollifier(const double* x, double eps_x, double* e, double* g, double* H)
{
    Eigen::Matrix<double, 12, 1> gg; Eigen::Matrix<double, 12, 12> HH;
    Edge_Edge_Mollifier(l3(x), l3(x + 3), l3(x + 6), l3(x + 9), eps_x, *e);
    Edge_Edge_Mollifier_Gradient(l3(x), l3(x + 3), l3(x + 6), l3(x + 9), eps_x, gg);
    Edge_Edge_Mollifier_Hessian(l3(x), l3(x + 3), l3(x + 6), l3(x + 9), eps_x, HH);
    memcpy(g, gg.data(), 96); memcpy(H, HH.data(), 1152);
}
double ref_mollifier_threshold(const double* x0) { double e; Edge_Edge_Mollifier_Threshold(l3(x0), l3(x0 + 3), l3(x0 + 6), l3(x0 + 9), e); return e; }
void ref_barrier_scalar(double d, double dHat2, double kappa, double* b, double* g, double* h)
{
    double k[3] = {kappa, 0, 0};
    Barrier<false>(d, dHat2, k, *b); Barrier_Gradient<false>(d, dHat2, k, *g); Barrier_Hessian<false>(d, dHat2, k, *h);
}
void ref_make_pd12(double* H) { Eigen::Matrix<double, 12, 12> M(H); makePD(M); memcpy(H, M.data(), 1152); }
int ref_accd(int kind, const double* x, const double* d, double eta, double thickness, double* toc)
{
    bool r;
    if (kind == 0) r = Point_Triangle_CCD(l3(x), l3(x + 3), l3(x + 6), l3(x + 9), l3(d), l3(d + 3), l3(d + 6), l3(d + 9), eta, thickness, *toc);
    else r = Edge_Edge_CCD(l3(x), l3(x + 3), l3(x + 6), l3(x + 9), l3(d), l3(d + 3), l3(d + 6), l3(d + 9), eta, thickness, *toc);
    return r ? 1 : 0;
}
int ref_aabb(int kind, const double* x, const double* d, double dist)
{
    switch (kind) {
    case 0: return Point_Triangle_CD_Broadphase(l3(x), l3(x + 3), l3(x + 6), l3(x + 9), dist);
    case 1: return Edge_Edge_CD_Broadphase(l3(x), l3(x + 3), l3(x + 6), l3(x + 9), dist);
    case 2: return Point_Triangle_CCD_Broadphase(l3(x), l3(x + 3), l3(x + 6), l3(x + 9), l3(d), l3(d + 3), l3(d + 6), l3(d + 9), dist);
    default: return Edge_Edge_CCD_Broadphase(l3(x), l3(x + 3), l3(x + 6), l3(x + 9), l3(d), l3(d + 3), l3(d + 6), l3(d + 9), dist);
    }
}
// Compute_Dihedral_Angle / _Gradient / _Hessian (Math/DIHEDRAL_ANGLE.h:9-24, 176-205, 1191-1298) in the argument order
// FEM/Shell/BENDING.h calls them with: the hinge stencil (x0; x1, x2; x3). g: 12, H: 12 x 12 row major.
void ref_dihedral(const double* x, double* theta, double* g, double* H)
{
    Eigen::Matrix<double, 3, 1> v0(x[0], x[1], x[2]), v1(x[3], x[4], x[5]), v2(x[6], x[7], x[8]), v3(x[9], x[10], x[11]);
    JGSL::Compute_Dihedral_Angle(v0, v1, v2, v3, *theta);
    Eigen::Matrix<double, 12, 1> grad;
    JGSL::Compute_Dihedral_Angle_Gradient(v0, v1, v2, v3, grad);
    Eigen::Matrix<double, 12, 12> Hs;
    JGSL::Compute_Dihedral_Angle_Hessian(v0, v1, v2, v3, Hs);
    for (int i = 0; i < 12; ++i) {
        g[i] = grad[i];
        for (int j = 0; j < 12; ++j) H[i * 12 + j] = Hs(i, j);
    }
}
}
