// Batched profile shaping in front of the lateral / velocity solves (include/tplb200_prep.h).
// One thread = one problem (both routines are scans over the samples with a carried state);
// arrays are [sample][problem], so every access of a warp is one 256-byte row.
#include <cmath>
#include <cstdio>
#include <cuda_runtime.h>

#include "../../include/tplb200_prep.h"

namespace {

thread_local char g_error[256] = "";

int fail(int code, const char* msg) {
    std::snprintf(g_error, sizeof g_error, "%s", msg);
    return code;
}

int check_launch(const char* what) {
    const cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) return 0;
    std::snprintf(g_error, sizeof g_error, "%s: %s", what, cudaGetErrorString(e));
    return TPLB_PREP_E_ARG;
}

// planning/utils.py:5-65.  profile: [N][2][B].
__global__ void rampify_velocity_kernel(int B, int N, const double* __restrict__ v0, const double* __restrict__ a0,
                                        const double* __restrict__ lim_v_in, double a_min, double a_max,
                                        double j_min, double j_max, double v_min, double step,
                                        double* __restrict__ profile) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    auto lim = [&](int t) { return fmax(lim_v_in[(size_t)t * B + b], v_min); };   // utils.py:17
    auto pv = [&](int t) -> double& { return profile[((size_t)t * 2 + 0) * B + b]; };
    auto pa = [&](int t) -> double& { return profile[((size_t)t * 2 + 1) * B + b]; };

    // backward pass, utils.py:24-35
    double cur_v = lim(N - 1), cur_a = 0.0;
    pv(0) = 0.0;
    pa(0) = 0.0;
    double lim_t = cur_v;                                    // lim_v[t]
    for (int t = N - 1; t > 0; --t) {
        pv(t) = cur_v;
        pa(t) = cur_a;
        const double lim_prev = lim(t - 1);
        const double lim_a = fmax(a_min, (cur_v - lim_prev) / step * cur_v);
        if (lim_a < 0.0) {
            cur_a = fmax(cur_a + j_min / cur_v * step, lim_a);
        } else {
            cur_a = 0.0;
            cur_v = lim_t;
        }
        cur_v += fmin(-cur_a / cur_v * step, lim_prev - cur_v);
        lim_t = lim_prev;
    }

    // forward pass, utils.py:39-63
    if (v0 == nullptr) {
        pv(0) = cur_v;
    } else {
        cur_v = fmax(v0[b], v_min);
        pv(0) = cur_v;
    }
    if (a0 == nullptr) {
        cur_a = -cur_a;
        pa(0) = cur_a;
    } else {
        cur_a = a0[b];
        pa(0) = cur_a;
    }
    double lim_a = 0.0;
    for (int t = 0; t < N; ++t) {
        const double p_t = pv(t);
        if (t < N - 1) lim_a = fmin(a_max, (pv(t + 1) - cur_v) / step * cur_v);
        if (lim_a > 0.0) {
            cur_a = fmin(cur_a + j_max / cur_v * step, lim_a);
        } else {
            cur_a = 0.0;
            cur_v = p_t;
        }
        const double next_v = cur_v + fmin(cur_a / cur_v * step, lim(t) - cur_v);
        cur_v = fmin(p_t, next_v);
        pv(t) = cur_v;
        pa(t) = cur_a;
    }
}

// path_optim.py:11-55.  The inner loop looks at every sample ahead of (behind) i, so the
// problem's `upper` row is staged in shared memory once: s_upper[k][threadIdx.x].
__global__ void rampify_lateral_kernel(int B, int N, int horizon, double step, double evasion_sharpness,
                                       const double* __restrict__ proj_distance, const double* __restrict__ path_v,
                                       double gap, const double* __restrict__ lower, const double* __restrict__ upper,
                                       double* __restrict__ d_offset) {
    extern __shared__ double s_upper[];                      // [horizon][blockDim.x]
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const int nt = blockDim.x, tx = threadIdx.x;
    for (int k = 0; k < horizon; ++k) s_upper[k * nt + tx] = upper[(size_t)k * B + b] - gap;
    auto out = [&](int i) -> double& { return d_offset[(size_t)i * B + b]; };
    for (int i = 0; i < N; ++i) out(i) = -10.0;               // :21-22
    const double proj = proj_distance[b];

    for (int pass_nr = 0; pass_nr < 2; ++pass_nr) {
        double d = pass_nr == 0 ? lower[b] : lower[(size_t)(horizon - 1) * B + b];
        for (int n = 0; n < horizon; ++n) {
            const int i = pass_nr == 0 ? n : horizon - 1 - n;
            d = fmax(lower[(size_t)i * B + b], d);
            // forward pass writes, backward pass folds with the forward value (np.maximum, :55)
            out(i) = pass_nr == 0 ? d : fmax(out(i), d);
            const double v = fmax(path_v[(size_t)i * B + b], 1e-8);
            double slope = -(evasion_sharpness / (v * v));
            if (pass_nr == 0) {
                for (int k = i; k < horizon; ++k)
                    slope = fmin(slope, (s_upper[k * nt + tx] - d) / (fmax(1.0, (double)(k - i)) * step));
            } else {
                for (int k = i; k >= 0; --k)
                    slope = fmin(slope, (s_upper[k * nt + tx] - d) / (fmax(1.0, (double)(i - k)) * step));
                slope = fmin(slope, (proj - d) / fmax(1.0, i * step));          // :50-51
            }
            d += step * slope;
        }
    }
}

// velocity_optim.py:98-104 through scipy's interp1d.  Grid and query points are formed exactly
// as numpy does (ss[i] = i*step, then + offset: two roundings, no FMA) because they decide
// which interval a query falls into.
__global__ void shift_interp_kernel(int B, int n, int rows, double step, const double* __restrict__ offset,
                                    int kind, const double* __restrict__ in, double* __restrict__ out) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = blockIdx.y;
    if (b >= B) return;
    auto grid = [&](int j) { return __dmul_rn((double)j, step); };
    const double xq = __dadd_rn(grid(i), offset[b]);
    // first j with grid(j) >= xq (numpy.searchsorted, side = left), from a guess
    int j = (int)ceil(xq / step);
    j = j < 0 ? 0 : (j > n ? n : j);
    while (j > 0 && grid(j - 1) >= xq) --j;
    while (j < n && grid(j) < xq) ++j;
    if (kind == TPLB_INTERP_ZERO) {
        // previous sample: last j' with grid(j') <= xq, clamped to the ends
        int p = (j < n && grid(j) == xq) ? j : j - 1;
        p = p < 0 ? 0 : (p > n - 1 ? n - 1 : p);
        for (int r = 0; r < rows; ++r)
            out[((size_t)i * rows + r) * B + b] = in[((size_t)p * rows + r) * B + b];
        return;
    }
    const int hi = j < 1 ? 1 : (j > n - 1 ? n - 1 : j), lo = hi - 1;
    const double x_lo = grid(lo), dx = __dadd_rn(grid(hi), -x_lo), w = __dadd_rn(xq, -x_lo);
    for (int r = 0; r < rows; ++r) {
        const double y_lo = in[((size_t)lo * rows + r) * B + b], y_hi = in[((size_t)hi * rows + r) * B + b];
        const double slope = __ddiv_rn(__dadd_rn(y_hi, -y_lo), dx);
        out[((size_t)i * rows + r) * B + b] = __dadd_rn(__dmul_rn(slope, w), y_lo);
    }
}

// Python's float `//` (CPython float_divmod): not floor(a / b) — 0.18 // 0.01 is 17.0, not 18.0
__host__ __device__ inline double py_floordiv(double a, double b) {
    double mod = fmod(a, b);
    double div = (a - mod) / b;
    if (mod != 0.0 && ((b < 0.0) != (mod < 0.0))) div -= 1.0;
    if (div == 0.0) return copysign(0.0, a / b);
    double fl = floor(div);
    if (div - fl > 0.5) fl += 1.0;
    return fl;
}

// util.py:92-100 with Python's floored float modulo
__device__ __forceinline__ double py_mod(double a, double m) {
    double r = fmod(a, m);
    if (r != 0.0 && ((r < 0.0) != (m < 0.0))) r += m;
    return r;
}

__device__ __forceinline__ double normalize_angle(double a) {
    const double two_pi = 3.141592653589793 * 2;
    a = py_mod(a, two_pi);
    a = py_mod(a + two_pi, two_pi);
    if (a > 3.141592653589793) a -= two_pi;
    return a;
}

// one actuator: append the command, drop what is older than the buffer may hold, pick the
// command to apply (core.py:95-120).  History rows are [slot][B].
__device__ __forceinline__ void dead_time_channel(double* ts, double* vals, int32_t* len_p, int B, int b, int capacity,
                                                  double t, double dt, double command, double dead_time,
                                                  double* applied) {
    int len = len_p[b];
    if (dt > 0.0) {
        if (len < capacity) {
            ts[(size_t)len * B + b] = t;
            vals[(size_t)len * B + b] = command;
            ++len;
        }
        const double keep = py_floordiv(dead_time, dt) + 1;   // core.py:99: dead_time // dt + 1
        while ((double)len > keep) {                          // list.pop(0)
            for (int i = 1; i < len; ++i) {
                ts[(size_t)(i - 1) * B + b] = ts[(size_t)i * B + b];
                vals[(size_t)(i - 1) * B + b] = vals[(size_t)i * B + b];
            }
            --len;
        }
        len_p[b] = len;
    }
    if (dead_time == 0.0 && len > 0) {
        *applied = vals[(size_t)(len - 1) * B + b];
        return;
    }
    for (int i = 0; i < len; ++i)
        if (t - ts[(size_t)i * B + b] <= dead_time) {
            *applied = vals[(size_t)i * B + b];
            return;
        }
}

__global__ void update_ego_kernel(const __grid_constant__ tplb_ego e, double t, double dt) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    const int B = e.batch;
    if (b >= B) return;
    double a = e.a[b], steer = e.steer_angle[b];
    dead_time_channel(e.acc_t, e.acc_value, e.acc_len, B, b, e.capacity, t, dt, e.control_acc[b], e.acc_dead_time, &a);
    dead_time_channel(e.steer_t, e.steer_value, e.steer_len, B, b, e.capacity, t, dt, e.control_steer[b],
                      e.steer_dead_time, &steer);
    double v = e.v[b], yaw = e.yaw[b];
    // core.py:122-134, the reference's association; the products are not fused into the sums
    e.x[b] = __dadd_rn(e.x[b], __dmul_rn(__dmul_rn(dt, v), cos(yaw)));
    e.y[b] = __dadd_rn(e.y[b], __dmul_rn(__dmul_rn(dt, v), sin(yaw)));
    const double q = v / e.v_ch;
    const double denom = __dmul_rn(e.wheel_base, __dadd_rn(1.0, __dmul_rn(q, q)));
    yaw = __dadd_rn(yaw, __dmul_rn(__dmul_rn(dt, v) / denom, tan(steer)));
    e.yaw[b] = normalize_angle(yaw);
    v = __dadd_rn(v, __dmul_rn(dt, a));
    e.v[b] = fmin(e.max_v, fmax(e.min_v, v));
    e.a[b] = a;
    e.steer_angle[b] = fmin(e.max_steer_angle, fmax(-e.max_steer_angle, steer));
}


// ---------------------------------------------------------------------------------------------
// Row f1: reference-path preparation for B (path, vehicle) pairs — what every MPC cycle does on the
// host before its solve (control/model_predictive_controller.py:124-128, 188-189): resample the planned
// trajectory equidistantly (utils.cpp:410-560 + util.py:134-191) and project the vehicle position on
// it (utils.cpp:257-408).  Both walk the polyline sequentially (running minimum / the next sample
// depends on the previous one), so one thread owns one problem; paths are [B][P][stride] rows as
// the reference holds them.  Every statement keeps the reference's association (this file is
// compiled with -fmad=false), so the results follow the C restatement in oracle/prep_oracle.c.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ long lmod(long n, long m) { n = n % m; return n < 0 ? n + m : n; }   // utils.cpp:16-23

// out [12][B]: distance, arc_len, alpha, index, start, end, point x, y, tangent x, y, angle, in_bounds
__global__ void project_kernel(int B, int P, int stride, const double* __restrict__ paths,
                               const double* __restrict__ position, int closed, double* __restrict__ out) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const double* pts = paths + (size_t)b * P * stride;
    const double px = position[2 * b], py = position[2 * b + 1];
    double best = INFINITY, offset = 0.0, arc_len = 0.0, alpha = 0.0, bx = 0.0, by = 0.0;
    long index = 0;
    int in_b = 0;
    const long end = closed ? P + 1 : P;
    double prevx = pts[0], prevy = pts[1];
    for (long i = 1; i < end; ++i) {                                                    // :283-321
        const long k = lmod(i, P);
        const double nx = pts[k * stride], ny = pts[k * stride + 1];
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
    double distance = sqrt(best);
    long idx_start, idx_end, idx_next;
    if (closed) { idx_start = lmod(index - 1, P); idx_end = lmod(index, P); idx_next = lmod(index + 1, P); }
    else { idx_start = index - 1 > 0 ? index - 1 : 0; idx_end = index; idx_next = index + 1 < P - 1 ? index + 1 : P - 1; }
    if (alpha < 0.5) index = idx_start;                                                 // :357-359
    const double sx = pts[idx_start * stride], sy = pts[idx_start * stride + 1];
    const double ex = pts[idx_end * stride], ey = pts[idx_end * stride + 1];
    double vx = ex - sx, vy = ey - sy;
    const double l = sqrt(vx * vx + vy * vy);
    vx /= l; vy /= l;
    const double adx = sx - bx, ady = sy - by;
    const double arc = offset + sqrt(adx * adx + ady * ady) * (alpha < 0 ? -1.0 : 1.0);
    double tx = vx, ty = vy;
    if (index < P - 2) {                                                                // :377-386
        double nvx = pts[idx_next * stride] - ex, nvy = pts[idx_next * stride + 1] - ey;
        const double nl = sqrt(nvx * nvx + nvy * nvy);
        nvx /= nl; nvy /= nl;
        tx = alpha * nvx + (1.0 - alpha) * vx;
        ty = alpha * nvy + (1.0 - alpha) * vy;
    }
    double ox = bx - px, oy = by - py;                                                  // :394-403
    const double on = sqrt(ox * ox + oy * oy);
    ox /= on; oy /= on;
    if (vx * (-oy) + vy * ox <= 0) distance *= -1.0;
    const double vals[12] = {distance, arc, alpha, (double)index, (double)idx_start, (double)idx_end, bx, by, tx, ty,
                             atan2(ty, tx), (double)in_b};
#pragma unroll
    for (int f = 0; f < 12; ++f) out[(size_t)f * B + b] = vals[f];
}

__device__ __forceinline__ double short_angle_dist_py(double x, double y) {            // util.py:71-89
    const double pi = 3.141592653589793;
    x = normalize_angle(x);
    y = normalize_angle(y);
    const double a0 = y - x, a1 = y - x + 2 * pi, a2 = y - x - 2 * pi;
    double a = a0;
    if (fabs(a1) < fabs(a)) a = a1;
    if (fabs(a2) < fabs(a)) a = a2;
    return a;
}

// util.resample_path: paths [B][P][6] -> rs [6][B][steps] (every component plane is a (B, steps) parameter
// array of the solver), ok [B] (0 where the reference returns None).  scratch: [B][2 P + 5 steps].
__global__ void resample_path_kernel(int B, int P, const double* __restrict__ paths, double dist, int steps,
                                     const int32_t* __restrict__ start_index, int zero_vel_at_end, int closed,
                                     double* __restrict__ rs, int32_t* __restrict__ ok, double* __restrict__ scratch) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const double* path = paths + (size_t)b * P * 6;
    double* pts = scratch + (size_t)b * (2 * P + 5 * steps);
    double* rsi = pts + 2 * P;
    auto out = [&](int comp, int i) -> double& { return rs[((size_t)comp * B + b) * steps + i]; };
    for (int i = 0; i < steps; ++i)
        for (int c = 0; c < 6; ++c) out(c, i) = 0.0;
    ok[b] = 0;
    // ---- resample (utils.cpp:410-560): de-duplicate, then march in steps of `dist` along the polyline
    long count_pts = 1;
    pts[0] = path[0]; pts[1] = path[1];
    for (long k = 1; k < P; ++k) {
        pts[2 * count_pts] = path[6 * k]; pts[2 * count_pts + 1] = path[6 * k + 1];
        const double dx = pts[2 * count_pts] - pts[2 * count_pts - 2], dy = pts[2 * count_pts + 1] - pts[2 * count_pts - 1];
        if (sqrt(dx * dx + dy * dy) != 0) count_pts += 1;
    }
    if (count_pts == 1) return;                       // a single point: the reference's result has one row
    long start = start_index ? start_index[b] : 0;
    if (closed) start = lmod(start, count_pts);
    else start = start < 0 ? 0 : (start > count_pts - 1 ? count_pts - 1 : start);
    for (int i = 0; i < 5 * steps; ++i) rsi[i] = 0.0;
    rsi[0] = pts[2 * start]; rsi[1] = pts[2 * start + 1]; rsi[3] = (double)start;
    if (closed) rsi[4] = (double)lmod(start + 1, count_pts);
    else { long n = start + 1; n = n < 0 ? 0 : (n > count_pts - 1 ? count_pts - 1 : n); rsi[4] = (double)n; }
    long count = 1, i = start;
    while (count < steps) {
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
            if (discriminant < 0) return;             // "cannot solve for next sampling point"
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
        if (count == prev_count) return;              // "resampling failed"
    }
    // ---- interp_resampled_path (util.py:155-191): orientation, arc length, velocity, curvature
    for (int k = 0; k < steps; ++k) {
        out(0, k) = rsi[5 * k]; out(1, k) = rsi[5 * k + 1];
        const double* prev = path + 6 * (long)rsi[5 * k + 3];
        const double* next = path + 6 * (long)rsi[5 * k + 4];
        const double t = rsi[5 * k + 2];
        if (!closed && rsi[5 * k + 4] == P - 1 && t > 1.0) {
            out(2, k) = next[2];
            out(3, k) = dist * k;
            out(5, k) = zero_vel_at_end ? 0.0 : next[5];
        } else {
            out(2, k) = prev[2] + t * short_angle_dist_py(prev[2], next[2]);
            out(3, k) = dist * k;
            out(5, k) = (1.0 - t) * prev[5] + t * next[5];
        }
    }
    for (int k = 1; k < steps; ++k) out(4, k - 1) = 2 * sin(short_angle_dist_py(out(2, k - 1), out(2, k)) / 2) / dist;
    const int last = steps - 1;
    if (closed) {
        const double gx = out(0, 0) - out(0, last), gy = out(1, 0) - out(1, last), gap = sqrt(gx * gx + gy * gy);
        if (gap == 0.0) out(4, last) = out(4, last - 1);
        else out(4, last) = 2 * sin(short_angle_dist_py(out(2, last), out(2, 0)) / 2) / gap;
    } else if (steps >= 2) {
        out(4, last) = out(4, last - 1);
    }
    ok[b] = 1;
}

// planning/path_vel_decomp/path_optim.py:303-305: the lateral solution back to Cartesian coordinates —
// x += -sin(phi) d, y += cos(phi) d, phi += atan(v_d) with d = opt.x[:-1, 0], v_d = opt.x[:-1, 1].
// paths [B][n][6] in place; xs: the solver's state buffer [t][X][B].  Thread per (problem, sample).
__global__ void frenet_to_cartesian_kernel(int B, int n, int X, double* __restrict__ paths,
                                           const double* __restrict__ xs) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y;
    if (b >= B) return;
    double* row = paths + ((size_t)b * n + i) * 6;
    const double d = xs[((size_t)i * X + 0) * B + b], vd = xs[((size_t)i * X + 1) * B + b];
    const double phi = row[2];
    row[0] += -sin(phi) * d;
    row[1] += cos(phi) * d;
    row[2] += atan(vd);
}

}  // namespace

extern "C" {

int32_t tplb_prep_abi_version(void) { return TPLB_PREP_ABI_VERSION; }
const char* tplb_prep_last_error(void) { return g_error; }

int32_t tplb_rampify_velocity(int32_t batch, int32_t n, const double* v0, const double* a0, const double* lim_v,
                              double a_min, double a_max, double j_min, double j_max, double v_min, double step,
                              double* profile, void* stream) {
    if (batch <= 0 || n <= 0) return fail(TPLB_PREP_E_ARG, "batch and n must be positive");
    if (!lim_v || !profile) return fail(TPLB_PREP_E_ARG, "lim_v / profile is NULL");
    const int block = 64;
    rampify_velocity_kernel<<<(batch + block - 1) / block, block, 0, static_cast<cudaStream_t>(stream)>>>(
        batch, n, v0, a0, lim_v, a_min, a_max, j_min, j_max, v_min, step, profile);
    return check_launch("tplb_rampify_velocity");
}

int32_t tplb_rampify_lateral(int32_t batch, int32_t n, int32_t horizon, double step, double evasion_sharpness,
                             const double* proj_distance, const double* path_v, double gap, const double* lower,
                             const double* upper, double* d_offset, void* stream) {
    if (batch <= 0 || n <= 0) return fail(TPLB_PREP_E_ARG, "batch and n must be positive");
    if (horizon <= 0 || horizon > n) return fail(TPLB_PREP_E_ARG, "horizon must be in 1..n");
    if (!proj_distance || !path_v || !lower || !upper || !d_offset) return fail(TPLB_PREP_E_ARG, "NULL array");
    const int block = 32;
    const size_t smem = sizeof(double) * (size_t)horizon * block;
    if (smem > 200 * 1024) return fail(TPLB_PREP_E_ARG, "horizon too long for the shared-memory tile (max 800)");
    if (smem > 48 * 1024)       // opt in per launch: the attribute belongs to the current device
        cudaFuncSetAttribute(rampify_lateral_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    rampify_lateral_kernel<<<(batch + block - 1) / block, block, smem, static_cast<cudaStream_t>(stream)>>>(
        batch, n, horizon, step, evasion_sharpness, proj_distance, path_v, gap, lower, upper, d_offset);
    return check_launch("tplb_rampify_lateral");
}

int32_t tplb_shift_interp(int32_t batch, int32_t n, int32_t rows, double step, const double* offset, int32_t kind,
                          const double* in, double* out, void* stream) {
    if (batch <= 0 || n < 2 || rows <= 0) return fail(TPLB_PREP_E_ARG, "batch, rows must be positive and n >= 2");
    if (!(step > 0.0)) return fail(TPLB_PREP_E_ARG, "step must be positive");
    if (kind != TPLB_INTERP_LINEAR && kind != TPLB_INTERP_ZERO) return fail(TPLB_PREP_E_ARG, "unknown kind");
    if (!offset || !in || !out || in == out) return fail(TPLB_PREP_E_ARG, "NULL or aliased array");
    const int block = 128;
    shift_interp_kernel<<<dim3((batch + block - 1) / block, n), block, 0, static_cast<cudaStream_t>(stream)>>>(
        batch, n, rows, step, offset, kind, in, out);
    return check_launch("tplb_shift_interp");
}

int32_t tplb_update_ego(const tplb_ego* ego, double t, double dt, void* stream) {
    if (!ego) return fail(TPLB_PREP_E_ARG, "ego is NULL");
    if (ego->struct_bytes != (int32_t)sizeof(tplb_ego)) return fail(TPLB_PREP_E_ARG, "tplb_ego size mismatch");
    if (ego->batch <= 0 || ego->capacity <= 0) return fail(TPLB_PREP_E_ARG, "batch and capacity must be positive");
    if (dt > 0.0) {
        const double need = fmax(py_floordiv(ego->acc_dead_time, dt), py_floordiv(ego->steer_dead_time, dt)) + 2;
        if ((double)ego->capacity < need) return fail(TPLB_PREP_E_ARG, "history capacity too small for dead_time / dt");
    }
    if (!ego->x || !ego->y || !ego->yaw || !ego->v || !ego->a || !ego->steer_angle || !ego->control_acc ||
        !ego->control_steer || !ego->acc_t || !ego->acc_value || !ego->acc_len || !ego->steer_t ||
        !ego->steer_value || !ego->steer_len)
        return fail(TPLB_PREP_E_ARG, "NULL array");
    const int block = 128;
    update_ego_kernel<<<(ego->batch + block - 1) / block, block, 0, static_cast<cudaStream_t>(stream)>>>(*ego, t, dt);
    return check_launch("tplb_update_ego");
}

int32_t tplb_project(int32_t batch, int32_t points, int32_t stride, const double* paths, const double* position,
                     int32_t closed, double* out, void* stream) {
    if (batch <= 0 || points < 2 || stride < 2) return fail(TPLB_PREP_E_ARG, "batch > 0, points >= 2, stride >= 2");
    if (!paths || !position || !out) return fail(TPLB_PREP_E_ARG, "NULL array");
    const int block = 64;
    project_kernel<<<(batch + block - 1) / block, block, 0, static_cast<cudaStream_t>(stream)>>>(
        batch, points, stride, paths, position, closed, out);
    return check_launch("tplb_project");
}

size_t tplb_resample_scratch_doubles(int32_t batch, int32_t points, int32_t steps) {
    return (size_t)batch * (2 * (size_t)points + 5 * (size_t)steps);
}

int32_t tplb_resample_path(int32_t batch, int32_t points, const double* paths, double step_size, int32_t steps,
                           const int32_t* start_index, int32_t zero_vel_at_end, int32_t closed, double* rs,
                           int32_t* ok, double* scratch, void* stream) {
    if (batch <= 0 || points < 1 || steps < 1) return fail(TPLB_PREP_E_ARG, "batch, points and steps must be positive");
    if (!(step_size > 0.0)) return fail(TPLB_PREP_E_ARG, "step_size must be positive");
    if (!paths || !rs || !ok || !scratch) return fail(TPLB_PREP_E_ARG, "NULL array");
    const int block = 32;
    resample_path_kernel<<<(batch + block - 1) / block, block, 0, static_cast<cudaStream_t>(stream)>>>(
        batch, points, paths, step_size, steps, start_index, zero_vel_at_end, closed, rs, ok, scratch);
    return check_launch("tplb_resample_path");
}

int32_t tplb_frenet_to_cartesian(int32_t batch, int32_t n, int32_t state_dims, double* paths, const double* xs,
                                 void* stream) {
    if (batch <= 0 || n <= 0 || state_dims < 2) return fail(TPLB_PREP_E_ARG, "batch, n > 0 and state_dims >= 2");
    if (!paths || !xs) return fail(TPLB_PREP_E_ARG, "NULL array");
    const int block = 128;
    frenet_to_cartesian_kernel<<<dim3((batch + block - 1) / block, n), block, 0, static_cast<cudaStream_t>(stream)>>>(
        batch, n, state_dims, paths, xs);
    return check_launch("tplb_frenet_to_cartesian");
}

}  // extern "C"
