// oracle/ref_gjk_shim.cpp -- C entry point onto the reference's UNMODIFIED openGJK
// (compiled from /root/reference/src/openGJK/openGJK.cpp where it lies; see Makefile).
// TEST INFRASTRUCTURE ONLY: pins orc_min_norm_hull() and generates tests/golden/gjk_golden.npz.
// Call shape follows closestPointsBetweenPointAndConvexHull, include/geometry.hpp:266-296:
// body 1 = the hull, body 2 = the single point (0,0,0); returns gjk()'s distance, v = witness.
#include <openGJK/openGJK.hpp>

extern "C" double ref_gjk_hull_origin(const double *pts, int npts, double *v) {
    struct simplex s;
    struct bd bd1, bd2;
    bd1.numpoints = npts;
    for (int i = 0; i < npts; i++) bd1.coord.push_back({{pts[3 * i], pts[3 * i + 1], pts[3 * i + 2]}});
    bd2.numpoints = 1;
    bd2.coord.push_back({{0.0, 0.0, 0.0}});
    return gjk(bd1, bd2, &s, v);
}
