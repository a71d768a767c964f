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

/* ---------------------------------------------------------------------------------------------
 * Row f1: reference-path preparation.  The reference's project / resample are C++ on Eigen
 * (library/src/utils.cpp:257-408, 410-560; Eigen is not in this image, so they cannot be compiled
 * here: PARITY UNPINNED against tplcpp — the restatement follows the source statement by statement and is
 * checked through geometric properties); interp_resampled_path is Python (library/tpl/util.py:155-191)
 * and IS pinned: tests/golden/prep_path.npz holds its outputs, recorded by running the reference's
 * own function.
 * --------------------------------------------------------------------------------------------- */
static long lmod(long n, long m) { n = n % m; if (n < 0) n += m; return n; }           /* utils.cpp:16-23 */

/* out: distance, arc_len, alpha, index, start, end, point x, y, tangent x, y, angle, in_bounds */
void tplo_project(const double* pts, int points_len, double px, double py, int closed, double* out) {
    double best = INFINITY, offset = 0.0, arc_len = 0.0, alpha = 0.0, bx = 0.0, by = 0.0;
    long index = 0;
    int in_b = 0;
    long end = closed ? points_len + 1 : points_len;
    double prevx = pts[0], prevy = pts[1];
    for (long i = 1; i < end; ++i) {                                                    /* :283-321 */
        const double nx = pts[2 * lmod(i, points_len)], ny = pts[2 * lmod(i, points_len) + 1];
        const double pvx = px - prevx, pvy = py - prevy, vx = nx - prevx, vy = ny - prevy;
        const double l = sqrt(vx * vx + vy * vy);
        double q = (pvx * vx + pvy * vy) / (vx * vx + vy * vy);
        double cx, cy;
        int inb = 1;
        if (q < 0) { inb = !closed && i != 1; q = 0.0; cx = prevx; cy = prevy; }
        else if (q > 1) { inb = !closed && i != end - 1; q = 1.0; cx = nx; cy = ny; }
        else { cx = prevx + vx * q; cy = prevy + vy * q; }
        const double dx = px - cx, dy = py - cy, d = dx * dx + dy * dy;
        if (d < best) { in_b = inb; best = d; bx = cx; by = cy; index = i; alpha = q; offset = arc_len; }
        arc_len += l;
        prevx = nx; prevy = ny;
    }
    double distance = sqrt(best);                                                       /* :329 */
    long idx_start, idx_end, idx_next;
    if (closed) { idx_start = lmod(index - 1, points_len); idx_end = lmod(index, points_len);
                  idx_next = lmod(index + 1, points_len); }
    else { idx_start = index - 1 > 0 ? index - 1 : 0; idx_end = index;
           idx_next = index + 1 < points_len - 1 ? index + 1 : points_len - 1; }
    if (alpha < 0.5) index = idx_start;                                                 /* :357-359 */
    const double sx = pts[2 * idx_start], sy = pts[2 * idx_start + 1];
    const double ex = pts[2 * idx_end], ey = pts[2 * idx_end + 1];
    double vx = ex - sx, vy = ey - sy;
    const double l = sqrt(vx * vx + vy * vy);
    vx /= l; vy /= l;
    const double adx = sx - bx, ady = sy - by;
    const double arc = offset + sqrt(adx * adx + ady * ady) * (alpha < 0 ? -1.0 : 1.0);  /* :372-373 */
    double tx = vx, ty = vy;
    if (index < points_len - 2) {                                                       /* :377-386 */
        double nvx = pts[2 * idx_next] - ex, nvy = pts[2 * idx_next + 1] - ey;
        const double nl = sqrt(nvx * nvx + nvy * nvy);
        nvx /= nl; nvy /= nl;
        tx = alpha * nvx + (1.0 - alpha) * vx;
        ty = alpha * nvy + (1.0 - alpha) * vy;
    }
    double ox = bx - px, oy = by - py;                                                  /* :394-403 */
    const double on = sqrt(ox * ox + oy * oy);
    ox /= on; oy /= on;
    const double rx = -oy, ry = ox;
    if (vx * rx + vy * ry <= 0) distance *= -1.0;
    out[0] = distance; out[1] = arc; out[2] = alpha; out[3] = (double)index; out[4] = (double)idx_start;
    out[5] = (double)idx_end; out[6] = bx; out[7] = by; out[8] = tx; out[9] = ty; out[10] = atan2(ty, tx);
    out[11] = (double)in_b;
}

/* rsi: [steps][5] = x, y, alpha, prev, next.  `scratch`: 2 * len_points doubles.  Returns the number of
 * rows written (1 when all points coincide), 0 for an empty input, -1 where the reference throws. */
int tplo_resample(const double* in_pts, int len_points, double dist, int steps, long start, int closed,
                  double* scratch, double* rsi) {
    if (len_points == 0 || steps == 0) return 0;
    double* pts = scratch;
    long count_pts = 1;
    pts[0] = in_pts[0]; pts[1] = in_pts[1];
    for (long k = 1; k < len_points; ++k) {                                             /* :431-437 */
        pts[2 * count_pts] = in_pts[2 * k]; pts[2 * count_pts + 1] = in_pts[2 * k + 1];
        const double dx = pts[2 * count_pts] - pts[2 * count_pts - 2], dy = pts[2 * count_pts + 1] - pts[2 * count_pts - 1];
        if (sqrt(dx * dx + dy * dy) != 0) count_pts += 1;
    }
    for (long i = 0; i < (long)steps * 5; ++i) rsi[i] = 0.0;
    if (count_pts == 1) { rsi[0] = pts[0]; rsi[1] = pts[1]; return 1; }                 /* :439-447 */
    if (closed) start = lmod(start, count_pts);
    else start = start < 0 ? 0 : (start > count_pts - 1 ? count_pts - 1 : start);
    rsi[0] = pts[2 * start]; rsi[1] = pts[2 * start + 1]; rsi[2] = 0.0; rsi[3] = (double)start;
    if (closed) rsi[4] = (double)lmod(start + 1, count_pts);
    else { long n = start + 1; n = n < 0 ? 0 : (n > count_pts - 1 ? count_pts - 1 : n); rsi[4] = (double)n; }
    long count = 1, i = start;
    while (count < steps) {                                                             /* :470-556 */
        const long prev_count = count;
        for (long k = 0; k < count_pts; ++k) {
            long prev_idx = i + k, next_idx = i + k + 1;
            if (closed) { prev_idx = lmod(prev_idx, count_pts); next_idx = lmod(next_idx, count_pts); }
            else {
                prev_idx = prev_idx < 0 ? 0 : (prev_idx > count_pts - 2 ? count_pts - 2 : prev_idx);
                next_idx = next_idx < 0 ? 0 : (next_idx > count_pts - 1 ? count_pts - 1 : next_idx);
            }
            const double ppx = pts[2 * prev_idx], ppy = pts[2 * prev_idx + 1];
            const double npx = pts[2 * next_idx], npy = pts[2 * next_idx + 1];
            const double vx = npx - ppx, vy = npy - ppy;
            const double l = sqrt(vx * vx + vy * vy), ls = l * l;
            const double vnx = vx / l, vny = vy / l;
            const double* c = rsi + 5 * (count - 1);
            const double D = (ppx - c[0]) * (npy - c[1]) - (npx - c[0]) * (ppy - c[1]);
            const double discriminant = dist * dist * ls - D * D;
            if (discriminant < 0) return -1;
            const double sq = sqrt(discriminant), sign_y = (vy < 0.0) ? -1.0 : 1.0;
            const double xp0 = D * vy, yp0 = -D * vx, xp1 = sign_y * vx * sq, yp1 = fabs(vy) * sq;
            double p0x = (xp0 + xp1) / ls + c[0], p0y = (yp0 + yp1) / ls + c[1];
            const double p1x = (xp0 - xp1) / ls + c[0], p1y = (yp0 - yp1) / ls + c[1];
            double q0 = (vnx * (p0x - ppx) + vny * (p0y - ppy)) / l;
            const double q1 = (vnx * (p1x - ppx) + vny * (p1y - ppy)) / l;
            const double tol = 1e-8;
            if (q0 < q1) { q0 = q1; p0x = p1x; p0y = p1y; }
            if ((!closed && next_idx == count_pts - 1) || (q0 > -tol && q0 - 1.0 < tol)) {
                i = prev_idx;
                double* o = rsi + 5 * count;
                o[0] = p0x; o[1] = p0y; o[2] = q0; o[3] = (double)prev_idx; o[4] = (double)next_idx;
                count += 1;
                break;
            }
        }
        if (count == prev_count) return -1;
    }
    return (int)count;
}

static double normalize_angle(double a) {                                               /* util.py:92-100 */
    const double pi = 3.141592653589793;
    a = pymod(a, pi * 2);
    a = pymod(a + pi * 2, pi * 2);
    if (a > pi) a -= pi * 2;
    return a;
}

static double short_angle_dist_py(double x, double y) {                                 /* util.py:71-89 */
    const double pi = 3.141592653589793;
    x = normalize_angle(x);
    y = normalize_angle(y);
    const double a0 = y - x, a1 = y - x + 2 * pi, a2 = y - x - 2 * pi;
    double a = a0;
    if (fabs(a1) < fabs(a)) a = a1;
    if (fabs(a2) < fabs(a)) a = a2;
    return a;
}

/* util.py:155-191.  path [len_path][6] = x, y, orientation, s, curvature, velocity; rsi [n][5]; rs [steps][6]
 * (rows >= n stay zero). */
void tplo_interp_resampled_path(const double* path, int len_path, const double* rsi, int n, double step_size,
                                int steps, int zero_vel_at_end, int closed, double* rs) {
    for (int i = 0; i < steps * 6; ++i) rs[i] = 0.0;
    for (int i = 0; i < n; ++i) {
        rs[6 * i] = rsi[5 * i]; rs[6 * i + 1] = rsi[5 * i + 1];
        const double* prev = path + 6 * (long)rsi[5 * i + 3];
        const double* next = path + 6 * (long)rsi[5 * i + 4];
        const double t = rsi[5 * i + 2];
        if (!closed && rsi[5 * i + 4] == len_path - 1 && t > 1.0) {
            rs[6 * i + 2] = next[2];
            rs[6 * i + 3] = step_size * i;
            rs[6 * i + 5] = zero_vel_at_end ? 0.0 : next[5];
        } else {
            rs[6 * i + 2] = prev[2] + t * short_angle_dist_py(prev[2], next[2]);
            rs[6 * i + 3] = step_size * i;
            rs[6 * i + 5] = (1.0 - t) * prev[5] + t * next[5];
        }
    }
    for (int i = 1; i < n; ++i)
        rs[6 * (i - 1) + 4] = 2 * sin(short_angle_dist_py(rs[6 * (i - 1) + 2], rs[6 * i + 2]) / 2) / step_size;
    const int last = steps - 1;                     /* rs[-1], rs[-2] index the OUTPUT array */
    if (closed) {
        const double gx = rs[0] - rs[6 * last], gy = rs[1] - rs[6 * last + 1], gap = sqrt(gx * gx + gy * gy);
        if (gap == 0.0) rs[6 * last + 4] = rs[6 * (last - 1) + 4];
        else rs[6 * last + 4] = 2 * sin(short_angle_dist_py(rs[6 * last + 2], rs[2]) / 2) / gap;
    } else if (steps >= 2) {
        rs[6 * last + 4] = rs[6 * (last - 1) + 4];
    }
}
