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

}  // extern "C"
