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

/* ---- SimCore.update_ego, library/tpl/simulation/core.py:91-134, one vehicle ------------------- */

typedef struct {
    double x, y, yaw, v, a, steer_angle, control_acc, control_steer;
    double acc_dead_time, steer_dead_time, wheel_base, v_ch, max_v, min_v, max_steer_angle;
} tplo_ego;

/* a command history: the reference's Python list of (t, value) */
typedef struct {
    int len;
    double t[64], value[64];
} tplo_history;

/* Python float floor division (CPython floatobject.c float_divmod) */
static double floordiv(double a, double b) {
    double mod = fmod(a, b), div = (a - mod) / b;
    if (mod != 0.0 && ((b < 0) != (mod < 0))) div -= 1.0;
    if (div == 0.0) return copysign(0.0, a / b);
    double f = floor(div);
    return (div - f > 0.5) ? f + 1.0 : f;
}

static double pymod(double a, double m) {
    double r = fmod(a, m);
    if (r != 0.0 && ((r < 0) != (m < 0))) r += m;
    return r;
}

static void history_step(tplo_history* h, double t, double dt, double command, double dead_time, double* applied) {
    if (dt > 0.0) {                                                /* core.py:95-102 */
        h->t[h->len] = t;
        h->value[h->len] = command;
        h->len++;
        while (h->len > floordiv(dead_time, dt) + 1) {
            for (int i = 1; i < h->len; ++i) { h->t[i - 1] = h->t[i]; h->value[i - 1] = h->value[i]; }
            h->len--;
        }
    }
    if (dead_time == 0.0 && h->len > 0) {                          /* :104-110 */
        *applied = h->value[h->len - 1];
    } else {
        for (int i = 0; i < h->len; ++i)
            if (t - h->t[i] <= dead_time) { *applied = h->value[i]; break; }
    }
}

void tplo_update_ego(tplo_ego* e, tplo_history* acc, tplo_history* steer, double t, double dt) {
    const double pi = 3.141592653589793;
    history_step(acc, t, dt, e->control_acc, e->acc_dead_time, &e->a);
    history_step(steer, t, dt, e->control_steer, e->steer_dead_time, &e->steer_angle);
    e->x += dt * e->v * cos(e->yaw);                               /* :122-123 */
    e->y += dt * e->v * sin(e->yaw);
    double r = e->v / e->v_ch;
    e->yaw += dt * e->v / (e->wheel_base * (1 + r * r)) * tan(e->steer_angle);   /* :125-126 */
    double a = pymod(e->yaw, pi * 2);                              /* util.py:95-98 */
    a = pymod(a + pi * 2, pi * 2);
    if (a > pi) a -= pi * 2;
    e->yaw = a;
    e->v += dt * e->a;                                             /* :130-131 */
    e->v = fmin(e->max_v, fmax(e->min_v, e->v));
    e->steer_angle = fmin(e->max_steer_angle, fmax(-e->max_steer_angle, e->steer_angle));
}
