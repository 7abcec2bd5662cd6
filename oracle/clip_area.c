/* CPU oracle (test infrastructure): float64 area of polygon ∩ axis-aligned square, restating
 * deephisto_b200/csrc/dh_region.cu::edge_term / clip_area operation for operation.
 * Reference call site: patch_samplers/region_samplers.py:125-134,180-189
 *   `self.polygon.intersection(patch_polygon).area` (shapely/GEOS, absent from the reference tree).
 * Build: gcc -O2 -ffp-contract=off -shared -fPIC (oracle/Makefile); no FMA contraction. */
#include <math.h>
#include <stdint.h>

static double dmin(double a, double b) { return a < b ? a : b; }
static double dmax(double a, double b) { return a > b ? a : b; }

static double edge_term(const double* e, double xa, double xb, double ya, double yb) {
    const double xA = e[0], yA = e[1], xB = e[2], yB = e[3], m = e[4], r = e[5], sgn = e[6];
    const double ys = dmax(yA, ya), ye = dmin(yB, yb);
    if (!(ys < ye)) return 0.0;
    const double xs = (ys == yA) ? xA : xA + (ys - yA) * m;
    const double xe = (ye == yB) ? xB : xA + (ye - yA) * m;
    const double wx = xb - xa;
    double val;
    if (xs == xe) {
        double g = dmin(dmax(xs, xa), xb) - xa;
        val = (ye - ys) * g;
    } else {
        const double xmin = dmin(xs, xe), xmax = dmax(xs, xe);
        const double ar = fabs(r);
        val = 0.0;
        const double cl = dmax(xmin, xa), ch = dmin(xmax, xb);
        if (cl < ch) val = ((ch - cl) * ar) * (((cl - xa) + (ch - xa)) * 0.5);
        const double ul = dmax(xmin, xb);
        if (ul < xmax) val = val + ((xmax - ul) * ar) * wx;
    }
    return sgn * val;
}

double oracle_clip_area(const double* edges, int64_t n_edges, double x, double y, double ps) {
    const double xa = x, xb = x + ps, ya = y, yb = y + ps;
    double acc = 0.0;
    for (int64_t e = 0; e < n_edges; ++e) acc = acc + edge_term(edges + 8 * e, xa, xb, ya, yb);
    return fabs(acc);
}

void oracle_clip_area_many(const double* edges, int64_t n_edges, const double* xs, const double* ys, int64_t n, double ps,
                           double* out) {
    for (int64_t i = 0; i < n; ++i) out[i] = oracle_clip_area(edges, n_edges, xs[i], ys[i], ps);
}
