// Batched iLQR kernels for sm_100a.  One instantiation per generated `Model`.
//
// Restates, for B independent problems, the solver of
// /root/reference/library/tpl/optim/templates/optim.c ("optim.c") :
//   rollout_init_kernel   optim.c:1096-1111   initial rollout + trajectory cost
//   multiplier_kernel     optim.c:1115-1136   multiplier update, per-outer-iteration reset
//   linearize_kernel      optim.c:896-912     derivative blocks of every stage   [stage parallel]
//   backward_kernel       optim.c:914-985     Riccati sweep + gains + box limits [problem parallel]
//   line_search_kernel    optim.c:732-792, 840-873, 987-1006   8 step sizes at once, first
//                                             improving one wins; mu schedule; stop test
//   accept_kernel         optim.c:844-848     copy the winning candidate        [stage parallel]
//   finalize_kernel       optim.c:1145-1149   termination flag
//
// Data layout: structure of arrays, problem index fastest — a warp of consecutive
// problems reads/writes 256 contiguous bytes for every (stage, component).
#pragma once

#include <cstdint>
#include <cuda_runtime.h>

#include "../../include/tplb200.h"
#include "device_math.cuh"

namespace tplb {

constexpr int kAlphas = TPLB_LINE_SEARCH_STEPS;

template <typename M>
struct Dims {
    static constexpr int X = M::X, U = M::U, C = M::C;
    static constexpr int Cs = C > 0 ? C : 1;
    // derivative block of one stage
    static constexpr int OFF_FX = 0;
    static constexpr int OFF_FU = OFF_FX + X * X;
    static constexpr int OFF_LX = OFF_FU + X * U;
    static constexpr int OFF_LU = OFF_LX + X;
    static constexpr int OFF_LXX = OFF_LU + U;
    static constexpr int OFF_LUU = OFF_LXX + X * X;
    static constexpr int OFF_LUX = OFF_LUU + U * U;
    static constexpr int STRIDE = OFF_LUX + U * X;
};

// Scratch carved out of tplb_batch.workspace.
struct Workspace {
    double* deriv;       // [t_max][STRIDE][B]
    double* cand_x;      // [8][t_max+1][X][B]
    double* cand_u;      // [8][t_max][U][B]
    double* cand_cost;   // [8][B]
    int32_t* winner;     // [B]  index of the accepted alpha, -1 = none
    int32_t* running;    // [B]  inner loop still active
};

__host__ __device__ inline size_t align_up(size_t v) { return (v + 255) & ~size_t(255); }

template <typename M>
__host__ __device__ inline Workspace carve(void* base, int B, int t_max, size_t* total = nullptr) {
    using D = Dims<M>;
    char* p = static_cast<char*>(base);
    size_t off = 0;
    Workspace w;
    w.deriv = reinterpret_cast<double*>(p + off);     off += align_up(sizeof(double) * (size_t)t_max * D::STRIDE * B);
    w.cand_x = reinterpret_cast<double*>(p + off);    off += align_up(sizeof(double) * (size_t)kAlphas * (t_max + 1) * D::X * B);
    w.cand_u = reinterpret_cast<double*>(p + off);    off += align_up(sizeof(double) * (size_t)kAlphas * t_max * D::U * B);
    w.cand_cost = reinterpret_cast<double*>(p + off); off += align_up(sizeof(double) * (size_t)kAlphas * B);
    w.winner = reinterpret_cast<int32_t*>(p + off);   off += align_up(sizeof(int32_t) * (size_t)B);
    w.running = reinterpret_cast<int32_t*>(p + off);  off += align_up(sizeof(int32_t) * (size_t)B);
    if (total) *total = off;
    return w;
}

__device__ __forceinline__ ParamView<double> param_view(const tplb_batch& q, int b) {
    ParamView<double> P;
    P.scalars = q.scalars;
    P.arrays = q.arrays;
    P.len = q.array_len;
    P.num_scenes = q.scenes;
    P.scene = __ldg(q.scene_index + b);
    return P;
}

// ---------------------------------------------------------------------------------
// integrators (optim.c:657-730); ctDynamics sees the same (t, dt) at all sub-stages
// ---------------------------------------------------------------------------------
template <typename M, typename R, typename PV>
__device__ __forceinline__ void step_state(const PV& P, const R* x, const R* u, R t, R h, int scheme, R* out) {
    constexpr int X = M::X;
    R k1[X], k2[X], y[X];
    M::ct_dynamics(P, x, u, t, h, k1);
    if (scheme == TPLB_EULER) {
#pragma unroll
        for (int i = 0; i < X; ++i) out[i] = x[i] + k1[i] * h;
    } else if (scheme == TPLB_HEUN) {
#pragma unroll
        for (int i = 0; i < X; ++i) y[i] = x[i] + k1[i] * h;
        M::ct_dynamics(P, y, u, t, h, k2);
        const R hh = h / R(2);
#pragma unroll
        for (int i = 0; i < X; ++i) out[i] = x[i] + (k1[i] + k2[i]) * hh;
    } else {
        R k3[X], k4[X];
        const R hh = h / R(2);
#pragma unroll
        for (int i = 0; i < X; ++i) y[i] = x[i] + k1[i] * hh;
        M::ct_dynamics(P, y, u, t, h, k2);
#pragma unroll
        for (int i = 0; i < X; ++i) y[i] = x[i] + k2[i] * hh;
        M::ct_dynamics(P, y, u, t, h, k3);
#pragma unroll
        for (int i = 0; i < X; ++i) y[i] = x[i] + k3[i] * h;
        M::ct_dynamics(P, y, u, t, h, k4);
        const R h6 = h / R(6);
#pragma unroll
        for (int i = 0; i < X; ++i) {
            R acc = k1[i] + R(0);                 // summation order of optim.c:717-724
            acc = k2[i] * R(2) + acc;
            acc = k3[i] * R(2) + acc;
            acc = k4[i] + acc;
            out[i] = x[i] + acc * h6;
        }
    }
}

// ---------------------------------------------------------------------------------
// initial rollout: x[t+1] = F(x[t], u[t]), trajCosts = sum_t l + l_end
// ---------------------------------------------------------------------------------
template <typename M>
__global__ void rollout_init_kernel(const __grid_constant__ tplb_batch q) {
    using D = Dims<M>;
    constexpr int X = D::X, U = D::U, C = D::C;
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    const int B = q.batch;
    if (b >= B) return;
    const ParamView<double> P = param_view(q, b);

    double w[D::Cs];
#pragma unroll
    for (int c = 0; c < C; ++c) w[c] = q.barrier_weight[(size_t)c * B + b];

    double x[X], xn[X], u[U], lam[D::Cs];
#pragma unroll
    for (int i = 0; i < X; ++i) x[i] = q.x[(size_t)i * B + b];

    double total = 0.0;
    const int T = q.horizon;
    for (int t = 0; t < T; ++t) {
#pragma unroll
        for (int i = 0; i < U; ++i) u[i] = q.u[((size_t)t * U + i) * B + b];
#pragma unroll
        for (int c = 0; c < C; ++c) lam[c] = q.lagrange_multiplier[((size_t)t * C + c) * B + b];
        step_state<M>(P, x, u, (double)t, q.dt, q.integrator_type, xn);
        double c;
        M::stage_cost(P, x, u, lam, w, (double)t, q.dt, &c);
        total += c;
#pragma unroll
        for (int i = 0; i < X; ++i) {
            x[i] = xn[i];
            q.x[((size_t)(t + 1) * X + i) * B + b] = xn[i];
        }
    }
    double ce;
    M::end_cost(P, x, (double)T, q.dt, &ce);
    total += ce;
    q.traj_costs[b] = total;
}

// ---------------------------------------------------------------------------------
// multiplier update lambda <- min(limit, max(0, lambda + w g)) for every stage, and
// the per-outer-iteration reset of the solver flags (row t == 0 does it).
// ---------------------------------------------------------------------------------
template <typename M>
__global__ void multiplier_kernel(const __grid_constant__ tplb_batch q, Workspace ws) {
    using D = Dims<M>;
    constexpr int X = D::X, U = D::U, C = D::C;
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    const int t = blockIdx.y;
    const int B = q.batch;
    if (b >= B) return;
    if (t == 0) {
        q.trajectory_changed[b] = 1;
        q.improved[b] = 0;
        q.iterations[b] = 0;
        ws.running[b] = 1;
    }
    if (C == 0) return;
    const ParamView<double> P = param_view(q, b);
    double x[X], u[U], lam[D::Cs], w[D::Cs], g[D::Cs];
#pragma unroll
    for (int i = 0; i < X; ++i) x[i] = q.x[((size_t)t * X + i) * B + b];
#pragma unroll
    for (int i = 0; i < U; ++i) u[i] = q.u[((size_t)t * U + i) * B + b];
#pragma unroll
    for (int c = 0; c < C; ++c) {
        lam[c] = q.lagrange_multiplier[((size_t)t * C + c) * B + b];
        w[c] = q.barrier_weight[(size_t)c * B + b];
        g[c] = 0.0;
    }
    M::constraints(P, x, u, lam, w, (double)t, q.dt, g);
#pragma unroll
    for (int c = 0; c < C; ++c) {
        double v = lam[c] + w[c] * g[c];
        v = (0.0 > v) ? 0.0 : v;
        const double lim = q.lg_mult_limit[(size_t)c * B + b];
        q.lagrange_multiplier[((size_t)t * C + c) * B + b] = (lim < v) ? lim : v;
    }
}

// ---------------------------------------------------------------------------------
// linearisation / quadratisation of every stage — thread per (problem, stage)
// ---------------------------------------------------------------------------------
template <typename M, bool kForce>
__global__ void linearize_kernel(const __grid_constant__ tplb_batch q, Workspace ws) {
    using D = Dims<M>;
    constexpr int X = D::X, U = D::U, C = D::C;
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    const int t = blockIdx.y;
    const int B = q.batch;
    if (b >= B) return;
    if (!kForce && !(ws.running[b] && q.trajectory_changed[b])) return;
    const ParamView<double> P = param_view(q, b);

    double x[X], u[U], lam[D::Cs], w[D::Cs];
#pragma unroll
    for (int i = 0; i < X; ++i) x[i] = q.x[((size_t)t * X + i) * B + b];
#pragma unroll
    for (int i = 0; i < U; ++i) u[i] = q.u[((size_t)t * U + i) * B + b];
#pragma unroll
    for (int c = 0; c < C; ++c) {
        lam[c] = q.lagrange_multiplier[((size_t)t * C + c) * B + b];
        w[c] = q.barrier_weight[(size_t)c * B + b];
    }
    double blk[D::STRIDE];
    if (q.use_quadratic_terms) {
        M::linearize(P, x, u, lam, w, (double)t, q.dt,
                     blk + D::OFF_FX, blk + D::OFF_FU, blk + D::OFF_LX, blk + D::OFF_LU,
                     blk + D::OFF_LXX, blk + D::OFF_LUU, blk + D::OFF_LUX);
    } else {
#pragma unroll
        for (int e = D::OFF_LXX; e < D::STRIDE; ++e) blk[e] = 0.0;
        M::dynamics_jacobians(P, x, u, (double)t, q.dt, blk + D::OFF_FX, blk + D::OFF_FU);
        M::cost_gradients(P, x, u, lam, w, (double)t, q.dt, blk + D::OFF_LX, blk + D::OFF_LU);
    }
    double* out = ws.deriv + (size_t)t * D::STRIDE * B + b;
#pragma unroll
    for (int e = 0; e < D::STRIDE; ++e) out[(size_t)e * B] = blk[e];
}

// ---------------------------------------------------------------------------------
// gains  k = -(Quu + mu I)^-1 Qu,  K = -(Quu + mu I)^-1 Qux   (optim.c:243-291)
// ---------------------------------------------------------------------------------
template <int X, int U>
__device__ __forceinline__ void control_gains(const double (&Quu)[U][U], const double (&Qu)[U],
                                              const double (&Qux)[U][X], double mu,
                                              double (&k)[U], double (&K)[U][X]) {
    static_assert(U == 1 || U == 2, "more than two controls are not supported (genopt.py:420-425)");
    if constexpr (U == 1) {
        double s = 0.0;
        if (Quu[0][0] > 0.0) s = -1.0 / (Quu[0][0] + mu);     // test on the un-regularised value
        k[0] = Qu[0] * s;
#pragma unroll
        for (int j = 0; j < X; ++j) K[0][j] = Qux[0][j] * s;
    } else {
        const double a = Quu[0][0] + mu, bb = Quu[0][1], d = Quu[1][1] + mu;
        const double det = a * d - bb * bb;
        const double s = -1.0 / det;                          // no definiteness check
        double Mi[2][2];
        Mi[0][0] = d * s;
        Mi[0][1] = -bb * s;
        Mi[1][0] = Mi[0][1];
        Mi[1][1] = a * s;
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            double acc = 0.0;
#pragma unroll
            for (int c = 0; c < 2; ++c) acc += Mi[i][c] * Qu[c];
            k[i] = acc;
#pragma unroll
            for (int j = 0; j < X; ++j) {
                double r = 0.0;
#pragma unroll
                for (int c = 0; c < 2; ++c) r += Mi[i][c] * Qux[c][j];
                K[i][j] = r;
            }
        }
    }
}

// ---------------------------------------------------------------------------------
// backward Riccati sweep — thread per problem, value function in registers
// ---------------------------------------------------------------------------------
template <typename M>
__global__ void __launch_bounds__(128)
backward_kernel(const __grid_constant__ tplb_batch q, Workspace ws, int iteration) {
    using D = Dims<M>;
    constexpr int X = D::X, U = D::U;
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    const int B = q.batch;
    if (b >= B) return;
    if (!ws.running[b]) return;
    q.iterations[b] = iteration + 1;             // optim.c:894
    q.trajectory_changed[b] = 0;                 // optim.c:911 (linearize_kernel ran just before)
    const ParamView<double> P = param_view(q, b);
    const int T = q.horizon;
    const double mu = q.mu[b];

    double Vx[X], Vxx[X][X];
    {
        double xT[X];
#pragma unroll
        for (int i = 0; i < X; ++i) xT[i] = q.x[((size_t)T * X + i) * B + b];
        M::end_derivatives(P, xT, (double)T, q.dt, Vx, &Vxx[0][0]);
    }

    for (int t = T - 1; t >= 0; --t) {
        const double* blk = ws.deriv + (size_t)t * D::STRIDE * B + b;
        auto ld = [&](int e) { return blk[(size_t)e * B]; };

        double A[X][X], Bm[X][U];
#pragma unroll
        for (int i = 0; i < X; ++i) {
#pragma unroll
            for (int j = 0; j < X; ++j) A[i][j] = ld(D::OFF_FX + i * X + j);
#pragma unroll
            for (int j = 0; j < U; ++j) Bm[i][j] = ld(D::OFF_FU + i * U + j);
        }

        double Qx[X], Qu[U], Qxx[X][X], Quu[U][U], Qux[U][X];
        double VA[X][X], VB[X][U];

#pragma unroll
        for (int i = 0; i < X; ++i) {                        // Qx = lx + A' Vx
            double acc = 0.0;
#pragma unroll
            for (int r = 0; r < X; ++r) acc += A[r][i] * Vx[r];
            Qx[i] = ld(D::OFF_LX + i) + acc;
        }
#pragma unroll
        for (int i = 0; i < U; ++i) {                        // Qu = lu + B' Vx
            double acc = 0.0;
#pragma unroll
            for (int r = 0; r < X; ++r) acc += Bm[r][i] * Vx[r];
            Qu[i] = ld(D::OFF_LU + i) + acc;
        }
#pragma unroll
        for (int i = 0; i < X; ++i) {                        // VA = Vxx A, VB = Vxx B
#pragma unroll
            for (int j = 0; j < X; ++j) {
                double acc = 0.0;
#pragma unroll
                for (int r = 0; r < X; ++r) acc += Vxx[i][r] * A[r][j];
                VA[i][j] = acc;
            }
#pragma unroll
            for (int j = 0; j < U; ++j) {
                double acc = 0.0;
#pragma unroll
                for (int r = 0; r < X; ++r) acc += Vxx[i][r] * Bm[r][j];
                VB[i][j] = acc;
            }
        }
#pragma unroll
        for (int i = 0; i < X; ++i)                          // Qxx = lxx + A' VA (lower triangle mirrored)
#pragma unroll
            for (int j = 0; j <= i; ++j) {
                double acc = 0.0;
#pragma unroll
                for (int r = 0; r < X; ++r) acc += A[r][i] * VA[r][j];
                Qxx[i][j] = acc;
                Qxx[j][i] = acc;
            }
#pragma unroll
        for (int i = 0; i < X; ++i)
#pragma unroll
            for (int j = 0; j < X; ++j) Qxx[i][j] = ld(D::OFF_LXX + i * X + j) + Qxx[i][j];
#pragma unroll
        for (int i = 0; i < U; ++i)                          // Quu = luu + B' VB (lower triangle mirrored)
#pragma unroll
            for (int j = 0; j <= i; ++j) {
                double acc = 0.0;
#pragma unroll
                for (int r = 0; r < X; ++r) acc += Bm[r][i] * VB[r][j];
                Quu[i][j] = acc;
                Quu[j][i] = acc;
            }
#pragma unroll
        for (int i = 0; i < U; ++i)
#pragma unroll
            for (int j = 0; j < U; ++j) Quu[i][j] = ld(D::OFF_LUU + i * U + j) + Quu[i][j];
#pragma unroll
        for (int i = 0; i < U; ++i)                          // Qux = lux + B' VA
#pragma unroll
            for (int j = 0; j < X; ++j) {
                double acc = 0.0;
#pragma unroll
                for (int r = 0; r < X; ++r) acc += Bm[r][i] * VA[r][j];
                Qux[i][j] = ld(D::OFF_LUX + i * X + j) + acc;
            }

        double k[U], K[U][X];
        control_gains<X, U>(Quu, Qu, Qux, mu, k, K);

        // box limits on the feed-forward step (optim.c:950-963)
#pragma unroll
        for (int d = 0; d < U; ++d) {
            const double ud = q.u[((size_t)t * U + d) * B + b];
            const double cand = ud + k[d];
            const double hi = q.u_max[((size_t)t * U + d) * B + b];
            const double lo = q.u_min[((size_t)t * U + d) * B + b];
            if (cand > hi) {
                k[d] = hi - ud;
#pragma unroll
                for (int j = 0; j < X; ++j) K[d][j] = 0.0;
            }
            if (cand < lo) {
                k[d] = lo - ud;
#pragma unroll
                for (int j = 0; j < X; ++j) K[d][j] = 0.0;
            }
            q.k[((size_t)t * U + d) * B + b] = k[d];
#pragma unroll
            for (int j = 0; j < X; ++j) q.K[((size_t)t * U * X + d * X + j) * B + b] = K[d][j];
        }

        // value function (optim.c:965-984)
        double KtQux[X][X], KtQuu[X][U];
#pragma unroll
        for (int i = 0; i < X; ++i) {
#pragma unroll
            for (int j = 0; j < X; ++j) {
                double acc = 0.0;
#pragma unroll
                for (int c = 0; c < U; ++c) acc += K[c][i] * Qux[c][j];
                KtQux[i][j] = acc;
            }
#pragma unroll
            for (int j = 0; j < U; ++j) {
                double acc = 0.0;
#pragma unroll
                for (int c = 0; c < U; ++c) acc += K[c][i] * Quu[c][j];
                KtQuu[i][j] = acc;
            }
        }
#pragma unroll
        for (int i = 0; i < X; ++i)
#pragma unroll
            for (int j = 0; j < X; ++j) {
                double v = KtQux[j][i] + KtQux[i][j];
#pragma unroll
                for (int c = 0; c < U; ++c) v += KtQuu[i][c] * K[c][j];
                Vxx[i][j] = v + Qxx[i][j];
            }
#pragma unroll
        for (int i = 0; i < X; ++i) {
            double v = 0.0;
#pragma unroll
            for (int c = 0; c < U; ++c) v += KtQuu[i][c] * k[c];
#pragma unroll
            for (int c = 0; c < U; ++c) v += K[c][i] * Qu[c];
#pragma unroll
            for (int c = 0; c < U; ++c) v += Qux[c][i] * k[c];
            Vx[i] = v + Qx[i];
        }
    }
}

// gradient-only sweep (optim.c:1038-1076): costate recursion, clipped descent direction
template <typename M>
__global__ void backward_first_order_kernel(const __grid_constant__ tplb_batch q, Workspace ws, int iteration) {
    using D = Dims<M>;
    constexpr int X = D::X, U = D::U;
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    const int B = q.batch;
    if (b >= B) return;
    if (!ws.running[b]) return;
    q.iterations[b] = iteration + 1;
    q.trajectory_changed[b] = 0;
    const ParamView<double> P = param_view(q, b);
    const int T = q.horizon;
    double Vx[X];
    {
        double xT[X], Vxx[X * X];
#pragma unroll
        for (int i = 0; i < X; ++i) xT[i] = q.x[((size_t)T * X + i) * B + b];
        M::end_derivatives(P, xT, (double)T, q.dt, Vx, Vxx);
    }
    for (int t = T - 1; t >= 0; --t) {
        const double* blk = ws.deriv + (size_t)t * D::STRIDE * B + b;
        auto ld = [&](int e) { return blk[(size_t)e * B]; };
        double Qx[X];
#pragma unroll
        for (int i = 0; i < X; ++i) {
            double acc = 0.0;
#pragma unroll
            for (int r = 0; r < X; ++r) acc += ld(D::OFF_FX + r * X + i) * Vx[r];
            Qx[i] = ld(D::OFF_LX + i) + acc;
        }
#pragma unroll
        for (int i = 0; i < U; ++i) {
            double acc = 0.0;
#pragma unroll
            for (int r = 0; r < X; ++r) acc += ld(D::OFF_FU + r * U + i) * Vx[r];
            const double gq = ld(D::OFF_LU + i) + acc;
            const size_t idx = ((size_t)t * U + i) * B + b;
            const double ui = q.u[idx];
            double kk = gq;
            const double cand = ui - gq;
            if (cand > q.u_max[idx]) kk = ui - q.u_max[idx];
            if (cand < q.u_min[idx]) kk = ui - q.u_min[idx];
            if (q.g) q.g[idx] = gq;
            q.k[idx] = kk;
        }
#pragma unroll
        for (int i = 0; i < X; ++i) Vx[i] = Qx[i];
    }
}

// ---------------------------------------------------------------------------------
// line search: the 8 step sizes alpha_i = 10^-i roll out concurrently.
// Block = PB problems x 8 step sizes; threadIdx.x = problem (coalesced), threadIdx.y = i.
// The lowest i whose cost passes `testImprovement` wins, exactly what the reference's
// sequential early-exit loop selects (optim.c:861-869).
// ---------------------------------------------------------------------------------
template <typename M, int PB>
__global__ void __launch_bounds__(PB * kAlphas)
line_search_kernel(const __grid_constant__ tplb_batch q, Workspace ws) {
    using D = Dims<M>;
    constexpr int X = D::X, U = D::U, C = D::C;
    __shared__ double s_cost[kAlphas][PB];

    const int lane = threadIdx.x;
    const int ai = threadIdx.y;
    const int b = blockIdx.x * PB + lane;
    const int B = q.batch;
    const bool live = (b < B) && ws.running[b];
    const int T = q.horizon;

    // alpha = 1.0 / pow(10, i)   (optim.c:863)
    double tens = 1.0;
    for (int i = 0; i < ai; ++i) tens *= 10.0;
    const double alpha = 1.0 / tens;

    double total = 0.0;
    if (live) {
        const ParamView<double> P = param_view(q, b);
        double w[D::Cs];
#pragma unroll
        for (int c = 0; c < C; ++c) w[c] = q.barrier_weight[(size_t)c * B + b];

        double* cx = ws.cand_x + (size_t)ai * (q.t_max + 1) * X * B + b;
        double* cu = ws.cand_u + (size_t)ai * q.t_max * U * B + b;

        double xn[X], xnext[X], un[U], lam[D::Cs];
#pragma unroll
        for (int i = 0; i < X; ++i) {
            xn[i] = q.x[(size_t)i * B + b];
            cx[(size_t)i * B] = xn[i];
        }
        const bool second_order = q.use_quadratic_terms != 0;
        for (int t = 0; t < T; ++t) {
#pragma unroll
            for (int d = 0; d < U; ++d) {
                const size_t idx = ((size_t)t * U + d) * B + b;
                if (second_order) {
                    double v = q.k[idx] * alpha + q.u[idx];
#pragma unroll
                    for (int j = 0; j < X; ++j)
                        v += q.K[((size_t)t * U * X + d * X + j) * B + b] *
                             (xn[j] - q.x[((size_t)t * X + j) * B + b]);
                    const double hi = q.u_max[idx], lo = q.u_min[idx];
                    const double capped = (hi < v) ? hi : v;           // optim.c:755-758
                    un[d] = (lo > capped) ? lo : capped;
                } else {
                    un[d] = q.u[idx] - q.k[idx] * alpha;               // optim.c:803-804
                }
                cu[((size_t)t * U + d) * B] = un[d];
            }
#pragma unroll
            for (int c = 0; c < C; ++c) lam[c] = q.lagrange_multiplier[((size_t)t * C + c) * B + b];
            step_state<M>(P, xn, un, (double)t, q.dt, q.integrator_type, xnext);
            double c;
            M::stage_cost(P, xn, un, lam, w, (double)t, q.dt, &c);
            total += c;
#pragma unroll
            for (int i = 0; i < X; ++i) {
                xn[i] = xnext[i];
                cx[((size_t)(t + 1) * X + i) * B] = xnext[i];
            }
        }
        double ce;
        M::end_cost(P, xn, (double)T, q.dt, &ce);
        total += ce;
        ws.cand_cost[(size_t)ai * B + b] = total;
    }
    s_cost[ai][lane] = total;
    __syncthreads();

    if (ai != 0 || b >= B) return;
    if (!live) {                                             // stopped earlier: nothing to accept
        ws.winner[b] = -1;
        return;
    }

    // testImprovement (optim.c:842) for i = 0..7 in order
    const double before = q.traj_costs[b];
    int win = -1;
#pragma unroll
    for (int i = kAlphas - 1; i >= 0; --i) {
        const double c = s_cost[i][lane];
        if (c < before && isfinite(c) && c >= 0.0) win = i;
    }
    ws.winner[b] = win;
    double now = before;
    double a_used = 1e-7;                                    // last step tried when none passes
    {
        double tn = 1.0;
        for (int i = 0; i < (win < 0 ? kAlphas - 1 : win); ++i) tn *= 10.0;
        a_used = 1.0 / tn;
    }
    q.alpha[b] = a_used;
    if (win >= 0) {
        now = s_cost[win][lane];
        q.traj_costs[b] = now;
        q.trajectory_changed[b] = 1;
        q.improved[b] = 1;
    }
    if (q.use_quadratic_terms) {                             // regularisation schedule (optim.c:989-999)
        int ms = q.mu_step[b];
        ms = (win >= 0) ? (ms - 1 > 0 ? ms - 1 : 0) : (ms + 1 < 7 ? ms + 1 : 7);
        q.mu_step[b] = ms;
        double m = 0.0;
        if (ms > 0) {
            m = 1.0;
            for (int i = 1; i < ms; ++i) m *= 10.0;          // 10^(ms-1), exact
        }
        q.mu[b] = m;
    }
    const double rel = fabs(now - before) / now;             // optim.c:1001-1006
    if (rel < q.min_rel_cost_change) {
        q.termination_condition[b] = 2;
        ws.running[b] = 0;
    }
}

// copy the accepted candidate into x, u (and keep prev_x, prev_k) — thread per (problem, stage)
template <typename M>
__global__ void accept_kernel(const __grid_constant__ tplb_batch q, Workspace ws) {
    using D = Dims<M>;
    constexpr int X = D::X, U = D::U;
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    const int t = blockIdx.y;                                // 0..T
    const int B = q.batch;
    if (b >= B) return;
    const int win = ws.winner[b];
    if (win < 0) return;
    const double* cx = ws.cand_x + (size_t)win * (q.t_max + 1) * X * B + b;
    const double* cu = ws.cand_u + (size_t)win * q.t_max * U * B + b;
#pragma unroll
    for (int i = 0; i < X; ++i) {
        const size_t idx = ((size_t)t * X + i) * B + b;
        if (q.keep_previous) q.prev_x[idx] = q.x[idx];
        q.x[idx] = cx[((size_t)t * X + i) * B];
    }
    if (t < q.horizon) {
#pragma unroll
        for (int d = 0; d < U; ++d) {
            const size_t idx = ((size_t)t * U + d) * B + b;
            if (q.keep_previous) q.prev_k[idx] = q.k[idx];
            q.u[idx] = cu[((size_t)t * U + d) * B];
        }
    }
}

__global__ void finalize_kernel(const __grid_constant__ tplb_batch q, int lg_done) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= q.batch) return;
    q.lg_iterations[b] = lg_done;
    if (q.iterations[b] == q.max_iterations) q.termination_condition[b] = 1;   // optim.c:1147-1149
}

// ---------------------------------------------------------------------------------
// warm-start shift (optim.c:1162-1177) — thread per problem (in-place, ascending t)
// ---------------------------------------------------------------------------------
template <typename M>
__global__ void shift_kernel(const __grid_constant__ tplb_batch q, int amount, const int32_t* amounts) {
    using D = Dims<M>;
    constexpr int X = D::X, U = D::U, C = D::C;
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    const int B = q.batch;
    if (b >= B) return;
    int n = amounts ? amounts[b] : amount;
    if (n < 0) n = 0;
    if (n == 0) return;
    const int T = q.horizon;
    for (int t = 0; t < T + 1; ++t) {
        const int s = (t + n < T) ? t + n : T;
#pragma unroll
        for (int i = 0; i < X; ++i) q.x[((size_t)t * X + i) * B + b] = q.x[((size_t)s * X + i) * B + b];
    }
    for (int t = 0; t < T; ++t) {
        const int s = (t + n < T - 1) ? t + n : T - 1;
#pragma unroll
        for (int i = 0; i < U; ++i) q.u[((size_t)t * U + i) * B + b] = q.u[((size_t)s * U + i) * B + b];
#pragma unroll
        for (int c = 0; c < C; ++c)
            q.lagrange_multiplier[((size_t)t * C + c) * B + b] = q.lagrange_multiplier[((size_t)s * C + c) * B + b];
    }
}

// ---------------------------------------------------------------------------------
// point evaluations of the dynamics (optim.c:1512-1652)
// ---------------------------------------------------------------------------------
template <typename M>
__global__ void dynamics_kernel(const __grid_constant__ tplb_batch q, const double* x_in, const double* u_in,
                                const int32_t* scene_of_point, int n, int t, double dt, int continuous,
                                double* x_out) {
    constexpr int X = M::X, U = M::U;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    ParamView<double> P;
    P.scalars = q.scalars;
    P.arrays = q.arrays;
    P.len = q.array_len;
    P.num_scenes = q.scenes;
    P.scene = scene_of_point ? scene_of_point[i] : ((n == q.batch && q.scene_index) ? q.scene_index[i] : 0);
    double x[X], u[U], out[X];
#pragma unroll
    for (int j = 0; j < X; ++j) x[j] = x_in[(size_t)j * n + i];
#pragma unroll
    for (int j = 0; j < U; ++j) u[j] = u_in[(size_t)j * n + i];
    if (continuous) M::ct_dynamics(P, x, u, (double)t, dt, out);
    else step_state<M>(P, x, u, (double)t, dt, q.integrator_type, out);
#pragma unroll
    for (int j = 0; j < X; ++j) x_out[(size_t)j * n + i] = out[j];
}

// ---------------------------------------------------------------------------------
// multi-start reduction: smallest finite cost of each contiguous group — warp per group
// ---------------------------------------------------------------------------------
__global__ void argmin_groups_kernel(const double* cost, int groups, int per_group,
                                     double* min_cost, int32_t* arg_min) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (warp >= groups) return;
    double best = INFINITY;
    int arg = -1;
    for (int i = lane; i < per_group; i += 32) {
        const double c = cost[(size_t)warp * per_group + i];
        if (isfinite(c) && (c < best)) { best = c; arg = warp * per_group + i; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double ob = __shfl_down_sync(0xffffffffu, best, o);
        const int oa = __shfl_down_sync(0xffffffffu, arg, o);
        if (oa >= 0 && (ob < best || (ob == best && (arg < 0 || oa < arg)))) { best = ob; arg = oa; }
    }
    if (lane == 0) { min_cost[warp] = best; arg_min[warp] = arg; }
}

// ---------------------------------------------------------------------------------
// FP64 pipe peak: 8 independent register-resident DFMA chains per thread
// ---------------------------------------------------------------------------------
__global__ void dfma_peak_kernel(double* sink, int inner) {
    double a0 = threadIdx.x * 1e-9, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3;
    double a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const double m = 0.999999999, c = 1e-9;
    for (int i = 0; i < inner; ++i) {
        a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
        a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
    }
    const double s = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
    if (s == 123.456) sink[0] = s;
}

}  // namespace tplb
