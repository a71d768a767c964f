/* TEST INFRASTRUCTURE — CPU restatement of the two profile-shaping routines of tpl's
 * path/velocity decomposition planner, one problem per call, plain C, strict IEEE.
 *
 *   tplo_rampify_velocity : library/tpl/planning/utils.py:5-65
 *   tplo_rampify_lateral  : library/tpl/planning/path_vel_decomp/path_optim.py:11-55
 *
 * Pinned against outputs of the reference's own (numba) functions: tests/golden/prep_*.npz,
 * recorded by tests/golden/make_golden_prep.py.  Only tests/, smoke() and bench.py's CPU
 * baseline may use this file; the product path is tpl_b200/csrc/prep.cu. */
#include <math.h>
#include <stddef.h>

static double dmax(double a, double b) { return a > b ? a : b; }
static double dmin(double a, double b) { return a < b ? a : b; }

/* lim_v[n] (modified copy is taken internally), profile[n][2]; has_v0 / has_a0 == 0 mean `None` */
void tplo_rampify_velocity(int n, int has_v0, double v0, int has_a0, double a0, const double* lim_v_in,
                           double a_min, double a_max, double j_min, double j_max, double v_min,
                           double step, double* lim_v /* scratch [n] */, double* profile) {
    for (int t = 0; t < n; ++t) {                                  /* utils.py:17-20 */
        lim_v[t] = dmax(lim_v_in[t], v_min);
        profile[2 * t] = 0.0;
        profile[2 * t + 1] = 0.0;
    }
    double current_v = lim_v[n - 1], current_a = 0.0;              /* :24-25 */
    for (int t = n - 1; t > 0; --t) {                              /* :26-35 */
        profile[2 * t] = current_v;
        profile[2 * t + 1] = current_a;
        double lim_a = dmax(a_min, (current_v - lim_v[t - 1]) / step * current_v);
        if (lim_a < 0.0) {
            current_a = dmax(current_a + j_min / current_v * step, lim_a);
        } else {
            current_a = 0.0;
            current_v = lim_v[t];
        }
        current_v += dmin(-current_a / current_v * step, lim_v[t - 1] - current_v);
    }
    if (!has_v0) {                                                 /* :39-43 */
        profile[0] = current_v;
    } else {
        current_v = dmax(v0, v_min);
        profile[0] = dmax(v0, v_min);
    }
    if (!has_a0) {                                                 /* :45-50 */
        current_a = -current_a;
        profile[1] = current_a;
    } else {
        current_a = a0;
        profile[1] = a0;
    }
    double lim_a = 0.0;
    for (int t = 0; t < n; ++t) {                                  /* :52-63 */
        if (t < n - 1) lim_a = dmin(a_max, (profile[2 * (t + 1)] - current_v) / step * current_v);
        if (lim_a > 0.0) {
            current_a = dmin(current_a + j_max / current_v * step, lim_a);
        } else {
            current_a = 0.0;
            current_v = profile[2 * t];
        }
        double next_v = current_v + dmin(current_a / current_v * step, lim_v[t] - current_v);
        current_v = dmin(profile[2 * t], next_v);
        profile[2 * t] = current_v;
        profile[2 * t + 1] = current_a;
    }
}

/* path_v[n] = path[:, 5]; lower, upper [n]; fwd, bwd scratch [n]; out [n] */
void tplo_rampify_lateral(int n, int horizon, double step, double evasion_sharpness, double proj_distance,
                          const double* path_v, double gap, const double* lower, const double* upper,
                          double* fwd, double* bwd, double* out) {
    for (int i = 0; i < n; ++i) fwd[i] = bwd[i] = 0.0 - 10;         /* :21-22 */
    for (int pass_nr = 0; pass_nr < 2; ++pass_nr) {
        double* pd = pass_nr == 0 ? fwd : bwd;
        double d = pass_nr == 0 ? lower[0] : lower[horizon - 1];
        const int first = pass_nr == 0 ? 0 : horizon - 1, dir = pass_nr == 0 ? 1 : -1;
        for (int c = 0, i = first; c < horizon; ++c, i += dir) {
            d = dmax(lower[i], d);
            pd[i] = d;
            const double v = dmax(path_v[i], 1e-8);
            double slope = -(evasion_sharpness / (v * v));
            for (int k = i; k >= 0 && k < horizon; k += dir) {      /* :45-46 */
                const int dist = k > i ? k - i : i - k;
                slope = dmin(slope, (upper[k] - gap - d) / ((dist > 1 ? dist : 1) * step));
            }
            if (pass_nr == 1) slope = dmin(slope, (proj_distance - d) / dmax(1, i * step));
            d += step * slope;
        }
    }
    for (int i = 0; i < n; ++i) out[i] = dmax(fwd[i], bwd[i]);     /* :55 */
}
