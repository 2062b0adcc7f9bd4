/* Ray caster of the synthetic LiDAR world (lidar_slam_arvc_b200/synth.py World.cast), in C: the numpy version costs
 * 0.6 s per 64-beam scan, which is what a 5 000-scan run (BASELINE configs[4]) cannot afford.  Host-side data
 * generation only - no part of the measured path.  Every expression is evaluated in the order numpy evaluates it
 * (compiled with -ffp-contract=off), NaNs propagate through min / max as in np.minimum / np.maximum, so the ranges are
 * bit-identical to the numpy implementation (tests/test_synth_cpu.py). */
#include <math.h>
#include <omp.h>
#include <stddef.h>

/* worker processes of a generation pool cast single-threaded (the pool is the parallelism) */
void arvc_synth_set_threads(int n) { omp_set_num_threads(n > 0 ? n : 1); }

static inline double np_min(double a, double b) { return (a != a || b != b) ? NAN : (a < b ? a : b); }
static inline double np_max(double a, double b) { return (a != a || b != b) ? NAN : (a > b ? a : b); }

/* o[3]; d[n][3] unit directions (world frame); segments[ns][4] x0 y0 x1 y1; boxes[nb][6] lo xyz hi xyz;
 * cylinders[nc][4] x y r h; t_out[n] = range of the first hit or +inf. */
void arvc_synth_cast(const double* o, const double* d, long n, const double* segments, int ns, double wall_h, const double* boxes, int nb,
                     const double* cylinders, int nc, double* t_out) {
#pragma omp parallel for schedule(static)
    for (long i = 0; i < n; ++i) {
        const double d0 = d[3 * i], d1 = d[3 * i + 1], d2 = d[3 * i + 2];
        double t = INFINITY;
        /* ground */
        {
            const double tg = -o[2] / d2;
            t = np_min(t, (d2 < 0 && tg > 0) ? tg : INFINITY);
        }
        /* walls */
        for (int s = 0; s < ns; ++s) {
            const double x0 = segments[4 * s], y0 = segments[4 * s + 1], x1 = segments[4 * s + 2], y1 = segments[4 * s + 3];
            const double ex = x1 - x0, ey = y1 - y0;
            const double den = d0 * ey - d1 * ex;
            const double wx = x0 - o[0], wy = y0 - o[1];
            const double tt = (wx * ey - wy * ex) / den;
            const double u = (wx * d1 - wy * d0) / den;
            const double z = o[2] + tt * d2;
            const int ok = (fabs(den) > 1e-12) && (tt > 0) && (u >= 0) && (u <= 1) && (z >= 0) && (z <= wall_h);
            t = np_min(t, ok ? tt : INFINITY);
        }
        /* boxes: slab method */
        if (nb > 0) {
            const double i0 = 1.0 / d0, i1 = 1.0 / d1, i2 = 1.0 / d2;
            double best = INFINITY;
            for (int b = 0; b < nb; ++b) {
                const double* B = boxes + 6 * b;
                const double lo0 = (B[0] - o[0]) * i0, lo1 = (B[1] - o[1]) * i1, lo2 = (B[2] - o[2]) * i2;
                const double hi0 = (B[3] - o[0]) * i0, hi1 = (B[4] - o[1]) * i1, hi2 = (B[5] - o[2]) * i2;
                const double tmin = np_max(np_max(np_min(lo0, hi0), np_min(lo1, hi1)), np_min(lo2, hi2));
                const double tmax = np_min(np_min(np_max(lo0, hi0), np_max(lo1, hi1)), np_max(lo2, hi2));
                const int ok = (tmax >= np_max(tmin, 0.0)) && (tmin > 0);
                best = np_min(best, ok ? tmin : INFINITY);
            }
            t = np_min(t, best);
        }
        /* vertical cylinders (side surface only) */
        if (nc > 0) {
            const double a = d0 * d0 + d1 * d1;
            double best = INFINITY;
            for (int c = 0; c < nc; ++c) {
                const double* C = cylinders + 4 * c;
                const double cx = C[0] - o[0], cy = C[1] - o[1], r = C[2], h = C[3];
                const double bb = d0 * cx + d1 * cy;
                const double cc = cx * cx + cy * cy - r * r;
                const double disc = bb * bb - a * cc;
                const double tt = (bb - sqrt(disc > 0 ? disc : NAN)) / a;
                const double z = o[2] + tt * d2;
                const int ok = (disc > 0) && (tt > 0) && (z >= 0) && (z <= h);
                best = np_min(best, ok ? tt : INFINITY);
            }
            t = np_min(t, best);
        }
        t_out[i] = t;
    }
}
