// CPU build of linearsfm_b200/csrc/builder_math.h for the unit test of the builder's per-landmark
// math (tests/test_builder.py compiles this with g++ and calls it through ctypes).  Test helper only:
// nothing in the product links it.
#include "builder_math.h"
extern "C" double bm_feature_blocks(double f, double b, double cx, double cy, double sigma, const double *pose,
                                    const double *X, const double *z0, const double *z1, double *V, double *W,
                                    double *Ub, double *gF, double *gP)
{
    bld::Cam c{f, b, cx, cy, 1.0 / (sigma * sigma)};
    bld::PoseLin P;
    bld::pose_lin(pose, P);
    return bld::feature_blocks(c, P, X, z0, z1, V, W, Ub, gF, gP);
}
extern "C" int bm_solve6(double *S, double *e) { return bld::solve6_spd(S, e) ? 1 : 0; }
extern "C" double bm_inv3(const double *a, double *o) { return bld::inv3_sym(a, o); }
extern "C" void bm_triangulate(double f, double b, double cx, double cy, const double *z, double *X)
{
    bld::Cam c{f, b, cx, cy, 1.0};
    bld::triangulate(c, z, X);
}
