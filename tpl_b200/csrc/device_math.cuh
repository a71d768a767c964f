// Device helpers the generated model code calls: small powers and the reference's
// interpolation lookups (/root/reference/library/tpl/optim/templates/optim.c:330-408),
// including the x86 quirk that a negative sample position selects the LAST sample.
#pragma once

#include <cstdint>

#include "fast_math.cuh"

namespace tplb {

// Elementary functions the generated code calls.  Default: the straight-line versions of
// fast_math.cuh; -DTPLB_LIBDEVICE_MATH selects CUDA's libdevice for A/B comparisons.
#ifdef TPLB_LIBDEVICE_MATH
__device__ __forceinline__ double m_sin(double x) { return sin(x); }
__device__ __forceinline__ double m_cos(double x) { return cos(x); }
__device__ __forceinline__ double m_tan(double x) { return tan(x); }
__device__ __forceinline__ void m_sincos(double x, double* s, double* c) { sincos(x, s, c); }
__device__ __forceinline__ double m_sqrt(double x) { return sqrt(x); }
__device__ __forceinline__ double m_rsqrt(double x) { return rsqrt(x); }
__device__ __forceinline__ double m_inv(double x) { return 1.0 / x; }
__device__ __forceinline__ double m_div(double a, double b) { return a / b; }
#else
__device__ __forceinline__ double m_sin(double x) { return sin_bf(x); }
__device__ __forceinline__ double m_cos(double x) { return cos_bf(x); }
__device__ __forceinline__ double m_tan(double x) { return tan_bf(x); }
__device__ __forceinline__ void m_sincos(double x, double* s, double* c) { sincos_bf(x, s, c); }
__device__ __forceinline__ double m_sqrt(double x) { return sqrt_bf(x); }
__device__ __forceinline__ double m_rsqrt(double x) { return rsqrt_bf(x); }
__device__ __forceinline__ double m_inv(double x) { return inv_bf(x); }
// a / b of the interpolation lookups: reciprocal (correctly rounded on every argument tried, loop
// invariant for a constant grid step) + one exact-remainder correction — Markstein's sequence,
// which returns the correctly rounded quotient and therefore the same sample interval as the
// reference's division, without the slow-path branch of the built-in operator
__device__ __forceinline__ double m_div(double a, double b) { return div_bf(a, b); }
#endif
// fp32 compute mode: libdevice's single-precision functions (<= 2 ulp, no slow path below 1e5)
__device__ __forceinline__ float m_sin(float x) { return sinf(x); }
__device__ __forceinline__ float m_cos(float x) { return cosf(x); }
__device__ __forceinline__ float m_tan(float x) { return tanf(x); }
__device__ __forceinline__ void m_sincos(float x, float* s, float* c) { sincosf(x, s, c); }
__device__ __forceinline__ float m_sqrt(float x) { return sqrtf(x); }
__device__ __forceinline__ float m_rsqrt(float x) { return rsqrtf(x); }
__device__ __forceinline__ float m_inv(float x) { return 1.0f / x; }
__device__ __forceinline__ float m_div(float a, float b) { return a / b; }

template <typename R> __device__ __forceinline__ R sq(R a) { return a * a; }
template <typename R> __device__ __forceinline__ R pow3h(R a) { return a * m_sqrt(a); }
template <typename R> __device__ __forceinline__ R ipow3(R a) { return a * a * a; }
template <typename R> __device__ __forceinline__ R ipow4(R a) { R b = a * a; return b * b; }
template <typename R> __device__ __forceinline__ R ipow5(R a) { R b = a * a; return b * b * a; }
template <typename R> __device__ __forceinline__ R ipow6(R a) { R b = a * a * a; return b * b; }
template <typename R> __device__ __forceinline__ R ipow7(R a) { R b = a * a * a; return b * b * a; }
template <typename R> __device__ __forceinline__ R ipow8(R a) { R b = a * a; b = b * b; return b * b; }

// optim.c:351-352: min(size-1, max(0lu, (size_t)floor(q))).  On x86-64 the
// double -> size_t conversion of a negative value yields a huge number, so the
// min() picks size-1; CUDA's conversion would saturate to 0, hence explicit.
template <typename R>
__device__ __forceinline__ int sample_index(R v, int n) {
    const bool outside = !(v >= R(0)) || v >= R(n);          // NaN and negative positions select the last sample
    return outside ? n - 1 : static_cast<int>(v);
}

// fmod(a, b) for b > 0: |a| < b returns a itself (what fmod returns, exactly); libdevice's general
// algorithm (a loop of scaled subtractions) runs only for the rare larger arguments
template <typename R>
__device__ __forceinline__ R fmod_small(R a, R b) {
    return (fabs(a) < b) ? a : fmod(a, b);
}

// optim.c:332-338
template <typename R>
__device__ __forceinline__ R short_angle_dist(R from, R to) {
    const R turn = R(3.14159265358979323846) * 2;
    const R d = fmod_small(to - from, turn);
    return fmod_small(2 * d, turn) - d;
}

// ---------------------------------------------------------------------------------
// the reference's lookups on one row of samples (optim.c:330-406).  kGlobal: the row is in
// global memory and read through the read-only path; otherwise it was staged in shared memory.
// ---------------------------------------------------------------------------------
template <typename R, bool kGlobal>
__device__ __forceinline__ R sample_at(const double* p) {
    if constexpr (kGlobal) return R(__ldg(p));
    else return R(*p);
}

// The lookups are branch-free (an empty row gives 0.0 through a select, optim.c:363-365, 380-382):
// inside the rollout chain every branch ends a basic block and with it the overlap of independent work.
template <typename R, bool kGlobal>
__device__ __forceinline__ R sample_or_zero(const double* p, int i, int n) {
    return n > 0 ? sample_at<R, kGlobal>(p + (i > 0 ? i : 0)) : R(0);
}

// optim.c:374-388 (+ initInterp :347-355)
template <typename R, bool kGlobal>
__device__ __forceinline__ R row_lerp(const double* p, int n, R x0, R dx, R x) {
    const R q = m_div(x - x0, dx);
    const int lo = sample_index(floor(q), n);
    const int hi = sample_index(ceil(q), n);
    R w = q - R(lo);
    w = (R(0) > w) ? R(0) : w;
    w = (w < R(1)) ? w : R(1);
    const R v = (R(1) - w) * sample_or_zero<R, kGlobal>(p, lo, n) + w * sample_or_zero<R, kGlobal>(p, hi, n);
    return n > 0 ? v : R(0);
}

// optim.c:392-406
template <typename R, bool kGlobal>
__device__ __forceinline__ R row_lerp_angle(const double* p, int n, R x0, R dx, R x) {
    const R q = m_div(x - x0, dx);
    const int lo = sample_index(floor(q), n);
    const int hi = sample_index(ceil(q), n);
    R w = q - R(lo);
    w = (R(0) > w) ? R(0) : w;
    w = (w < R(1)) ? w : R(1);
    const R v0 = sample_or_zero<R, kGlobal>(p, lo, n);
    const R v = v0 + short_angle_dist(v0, sample_or_zero<R, kGlobal>(p, hi, n)) * w;
    return n > 0 ? v : R(0);
}

// optim.c:357-370
template <typename R, bool kGlobal>
__device__ __forceinline__ R row_box_interp(const double* p, int n, R dx, R x) {
    return sample_or_zero<R, kGlobal>(p, sample_index(floor(m_div(x, dx)), n), n);
}

// optim.c:410-448: periodic lookup over [xs[0], xs[0] + len); the samples end `gap` before the
// period does and the last segment blends arr[n-1] back into arr[0].  The reference leaves its
// indices unclamped (with gap <= 0 a position at or beyond the last sample reads past the
// array); here they clamp to the last sample.
template <typename R, bool kGlobal>
__device__ __forceinline__ R row_lerp_wrap(const double* xs, int nxs, const double* p, int n, R len, R dx, R x) {
    const R first = sample_or_zero<R, kGlobal>(xs, 0, nxs);
    const R last = first + R(n - 1) * dx;
    const R gap = len - (last - first);
    R y = x - first;
    y = (fabs(y) < fabs(len)) ? y : fmod(y, len);          // fmod(a, b) = a for |a| < |b|, exactly
    y = (y < R(0)) ? y + len : y;
    y += first;
    const bool across = (y >= last) && (gap > R(0));
    const R q = m_div(y - first, dx);
    const int lo = across ? n - 1 : sample_index(floor(q), n);
    const int hi = across ? 0 : sample_index(ceil(q), n);
    const R w = across ? m_div(y - last, gap) : q - R(lo);
    const R v = (R(1) - w) * sample_or_zero<R, kGlobal>(p, lo, n) + w * sample_or_zero<R, kGlobal>(p, hi, n);
    return n > 0 ? v : R(0);
}

// optim.c:457-481 (+ initInterp :347-355 per axis): p is rows x cols, row-major; x runs along a row
template <typename R, bool kGlobal>
__device__ __forceinline__ R row_blerp(const double* p, int rows, int cols, R x0, R y0, R dx, R dy, R x, R y) {
    const R qx = m_div(x - x0, dx);
    const R qy = m_div(y - y0, dy);
    const int x_lo = sample_index(floor(qx), cols), x_hi = sample_index(ceil(qx), cols);
    const int y_lo = sample_index(floor(qy), rows), y_hi = sample_index(ceil(qy), rows);
    R ax = qx - R(x_lo);
    ax = (R(0) > ax) ? R(0) : ax;
    ax = (ax < R(1)) ? ax : R(1);
    R ay = qy - R(y_lo);
    ay = (R(0) > ay) ? R(0) : ay;
    ay = (ay < R(1)) ? ay : R(1);
    const int n = rows * cols;                               // an empty map reads as 0.0 (the reference would index out of bounds)
    const R p0 = (R(1) - ay) * sample_or_zero<R, kGlobal>(p, y_lo * cols + x_lo, n)
               + ay * sample_or_zero<R, kGlobal>(p, y_hi * cols + x_lo, n);
    const R p1 = (R(1) - ay) * sample_or_zero<R, kGlobal>(p, y_lo * cols + x_hi, n)
               + ay * sample_or_zero<R, kGlobal>(p, y_hi * cols + x_hi, n);
    return n > 0 ? (R(1) - ax) * p0 + ax * p1 : R(0);
}

// optim.c:330 — the reference indexes unchecked; clamp instead of reading out of bounds
template <typename R, bool kGlobal>
__device__ __forceinline__ R row_value(const double* p, int n, R i) {
    int j = static_cast<int>(i);
    j = j < 0 ? 0 : (j >= n ? n - 1 : j);
    return sample_or_zero<R, kGlobal>(p, j, n);
}

// Parameter view of ONE problem: scalars of its scene and its scene's rows of the
// parameter arrays.  scalars: [num_scalars][S]; array a: [S][len[a]].  Storage is always
// fp64; R is the type the model is evaluated in (double, or float in fp32 compute mode).
template <typename R>
struct ParamView {
    const double* scalars;
    const double* const* arrays;
    const int32_t* len;                 // samples per scene (rows * cols of a 2-D array)
    const int32_t* cols;                // row length of a 2-D array, 0 for a 1-D one
    int32_t num_scenes;
    int32_t scene;

    __device__ __forceinline__ R scalar(int i) const {
        return R(__ldg(scalars + (size_t)i * num_scenes + scene));
    }
    __device__ __forceinline__ const double* row(int a) const {
        return arrays[a] + (size_t)scene * len[a];
    }
    __device__ __forceinline__ R lerp(int a, R x0, R dx, R x) const { return row_lerp<R, true>(row(a), len[a], x0, dx, x); }
    __device__ __forceinline__ R lerp_angle(int a, R x0, R dx, R x) const {
        return row_lerp_angle<R, true>(row(a), len[a], x0, dx, x);
    }
    __device__ __forceinline__ R box_interp(int a, R dx, R x) const { return row_box_interp<R, true>(row(a), len[a], dx, x); }
    __device__ __forceinline__ R array_value(int a, R i) const { return row_value<R, true>(row(a), len[a], i); }
    __device__ __forceinline__ R lerp_wrap(int xs, int a, R period, R dx, R x) const {
        return row_lerp_wrap<R, true>(row(xs), len[xs], row(a), len[a], period, dx, x);
    }
    __device__ __forceinline__ R blerp(int a, R x0, R y0, R dx, R dy, R x, R y) const {
        const int c = cols[a];
        return row_blerp<R, true>(row(a), c > 0 ? len[a] / c : 0, c, x0, y0, dx, dy, x, y);
    }
};


// ParamView whose scalar parameters sit in registers.  A kernel that walks the stages in a loop
// and stores to global memory inside it (the rollouts) would otherwise reload every scalar in
// every stage: the compiler cannot hoist the loads over the stores.  The generated code asks
// for scalars by literal index, so `cached` never becomes an addressable array.
template <typename R, int NS>
struct CachedParamView : ParamView<R> {
    R cached[NS > 0 ? NS : 1];

    __device__ __forceinline__ void preload(const ParamView<R>& base) {
        static_cast<ParamView<R>&>(*this) = base;
#pragma unroll
        for (int i = 0; i < NS; ++i) cached[i] = base.scalar(i);
    }
    __device__ __forceinline__ R scalar(int i) const { return cached[i]; }
};

// Parameters of ONE problem held on chip: scalars in registers, the rows of its parameter arrays
// (reference path / map samples) staged in shared memory — the one-launch kernel (solo.cuh) looks
// them up inside its serial rollout chain, where a global-memory round trip per lookup would
// dominate the stage.
template <typename R, int NS, int NA>
struct StagedParamView {
    R cached[NS > 0 ? NS : 1];
    const double* rows[NA > 0 ? NA : 1];          // shared memory
    int32_t len[NA > 0 ? NA : 1];
    int32_t cols[NA > 0 ? NA : 1];

    __device__ __forceinline__ R scalar(int i) const { return cached[i]; }
    __device__ __forceinline__ R lerp(int a, R x0, R dx, R x) const { return row_lerp<R, false>(rows[a], len[a], x0, dx, x); }
    __device__ __forceinline__ R lerp_angle(int a, R x0, R dx, R x) const {
        return row_lerp_angle<R, false>(rows[a], len[a], x0, dx, x);
    }
    __device__ __forceinline__ R box_interp(int a, R dx, R x) const { return row_box_interp<R, false>(rows[a], len[a], dx, x); }
    __device__ __forceinline__ R array_value(int a, R i) const { return row_value<R, false>(rows[a], len[a], i); }
    __device__ __forceinline__ R lerp_wrap(int xs, int a, R period, R dx, R x) const {
        return row_lerp_wrap<R, false>(rows[xs], len[xs], rows[a], len[a], period, dx, x);
    }
    __device__ __forceinline__ R blerp(int a, R x0, R y0, R dx, R dy, R x, R y) const {
        const int c = cols[a];
        return row_blerp<R, false>(rows[a], c > 0 ? len[a] / c : 0, c, x0, y0, dx, dy, x, y);
    }
};

}  // namespace tplb
