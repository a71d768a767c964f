// One launch, one CTA per problem: the whole update() (optim.c:1091-1160) for batches that
// cannot fill the GPU — a single planning / control cycle, or a handful of candidate manoeuvres.
//
// The batched sequences of cabi.cu need 50-80 launches per update(); with one problem every
// launch is a latency-bound chain of one thread, and the launch gaps plus the chain of one
// lane per warp add up to more than the reference needs on a CPU core.  Here the problem
// lives in shared memory for the whole solve and each phase uses the parallelism it has:
//
//   linearisation, stage costs      one thread per stage / per (step size, stage)      [stage parallel]
//   Riccati recursion               one warp per problem: every entry of Vxx A, A'Vxx A, the gains
//                                   and the new value function on its own lane, operands in
//                                   shared memory, __syncwarp between the dependent products
//                                   (small X: one lane, the recursion is too narrow to split)
//   line search                     the 8 step sizes alpha = 10^-i on 8 lanes of ONE warp: the
//                                   dependent FP64 chain of the rollouts is issued once for all of them
//   cost sums, decisions            in the reference's order (t = 0..T, then first improving step)
//
// Every floating-point expression is the one the batched kernels evaluate (same device
// functions, same order of the sums), so both paths return identical solutions.
#pragma once

#include "solver.cuh"

namespace tplb {

constexpr int kSoloThreads = 256;

// -DTPLB_SOLO_TIMING: thread 0 of block 0 accumulates the SM cycles spent in each phase
// (read back with tplb_debug_solo_cycles; development aid, not part of the ABI contract)
#ifdef TPLB_SOLO_TIMING
__device__ long long g_solo_cycles[8];
#define SOLO_TICK(slot)                                                   \
    do {                                                                  \
        if (blockIdx.x == 0 && threadIdx.x == 0) {                        \
            const long long now_ = clock64();                             \
            g_solo_cycles[slot] += now_ - tick_;                          \
            tick_ = now_;                                                 \
        }                                                                 \
    } while (0)
#else
#define SOLO_TICK(slot) do { } while (0)
#endif

template <typename M>
struct SoloLayout {
    using D = Dims<M>;
    static constexpr int X = D::X, U = D::U, Cs = D::Cs, NSCs = D::NSCs, NC = D::COMPACT;
    // doubles per stage slot; every array is [component][TP] with TP = T + 1
    static constexpr int PER_STAGE = X + U /*u*/ + U /*k*/ + U * X /*K*/ + U /*u_max*/ + U /*u_min*/ + Cs /*lambda*/ +
                                     NSCs + NC + kAlphas * X + kAlphas * U + kAlphas /*cost terms*/;
    // scratch of the cooperative Riccati step: dense record, V, VA, VB, Q terms, gains
    static constexpr int SCRATCH = 2 * D::DENSE + 3 * X * X + 3 * X * U + 3 * X + 3 * U + U * U + 40;
    // `array_samples`: total length of the problem's parameter arrays (staged behind the scratch)
    __host__ __device__ static size_t bytes(int T, int array_samples) {
        return sizeof(double) * ((size_t)PER_STAGE * (T + 1) + SCRATCH + array_samples) + 512;
    }
};

struct SoloStatus {
    double traj_costs, alpha, mu;
    double total[kAlphas];
    int32_t mu_step, iterations, lg_iterations, trajectory_changed, improved, termination_condition;
    int32_t running, winner, last_tried, c_lin, c_bwd, c_roll, lam_zero;
};

// one lane: the recursion exactly as the batched backward sweep runs it
template <typename M, typename PV>
__device__ __forceinline__ void solo_backward_serial(const PV& P, double mu, double dt, int T, int TP, const double* sx,
                                                     const double* su, const double* shi, const double* slo,
                                                     const double* ssc, const double* srec, double* sk, double* sK) {
    using D = Dims<M>;
    constexpr int X = D::X, U = D::U, NC = D::COMPACT, NSC = D::NSC;
    double Vx[X], Vxx[X][X];
    {
        double xT[X], sc[D::NSCs];
#pragma unroll
        for (int i = 0; i < X; ++i) xT[i] = sx[i * TP + T];
#pragma unroll
        for (int j = 0; j < NSC; ++j) sc[j] = ssc[j * TP + T];
        M::end_derivatives(P, xT, sc, double(T), dt, Vx, &Vxx[0][0]);
    }
    for (int t = T - 1; t >= 0; --t) {
        double rec[NC], ub[U], hib[U], lob[U], k[U], K[U][X];
#pragma unroll
        for (int s = 0; s < M::DERIV_COMPACT; ++s) rec[s] = srec[s * TP + t];
#pragma unroll
        for (int d = 0; d < U; ++d) {
            ub[d] = su[d * TP + t];
            hib[d] = shi[d * TP + t];
            lob[d] = slo[d * TP + t];
        }
        riccati_stage<M, double>(rec, Vx, Vxx, mu, ub, hib, lob, k, K);
#pragma unroll
        for (int d = 0; d < U; ++d) {
            sk[d * TP + t] = k[d];
#pragma unroll
            for (int j = 0; j < X; ++j) sK[(d * X + j) * TP + t] = K[d][j];
        }
    }
}

// Four warps: the same sums, one output entry per lane (riccati_stage is the specification).
// Every product of the step has the form  out = [base +] sum_r L[r] * R[r]  over the X rows, so a
// lane describes its output once, before the sweep, by six offsets (DotJob) and the stage loop is
// branch-free: 96 lanes (three warps, on three schedulers) run the same X-term FMA chain on
// different operands of the stage's scratch in shared memory; the fourth warp expands the compact
// record of the NEXT stage into the dense fx|fu|lx|lu|lxx|luu|lux block meanwhile (two buffers).
// An entry the model's structure table marks as 0 / 1 enters as the literal 0.0 / 1.0, which
// leaves every finite sum unchanged.
constexpr int kCoopLanes = 96, kCoopThreads = 128;

struct DotJob {
    int lo, ls;      // L[r] = dense[lo + r * ls]            (a column of A or B)
    int ro, rs;      // R[r] = scratch[ro + r * rs]
    int bo;          // base = dense[bo], -1: none
    int oo;          // out  = scratch[oo], -1: this lane has no output in this step
};

__device__ __forceinline__ void named_barrier(int id, int threads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}

template <int X>
__device__ __forceinline__ void solo_dot(double* scratch, const double* dense, const DotJob& j) {
    double acc = 0.0;
#pragma unroll
    for (int r = 0; r < X; ++r) acc += dense[j.lo + r * j.ls] * scratch[j.ro + r * j.rs];
    if (j.oo >= 0) scratch[j.oo] = j.bo >= 0 ? dense[j.bo] + acc : acc;
}

template <typename M>
struct SoloScratch {
    using D = Dims<M>;
    static constexpr int X = D::X, U = D::U;
    static constexpr int DENSE0 = 0, DENSE1 = D::DENSE;      // the stage's record, double buffered
    static constexpr int VXX = 2 * D::DENSE, VX = VXX + X * X, VA = VX + X, VB = VA + X * X, QX = VB + X * U,
                         QU = QX + X, QXX = QU + U, QUU = QXX + X * X, QUX = QUU + U * U, KC = QUX + U * X,
                         KS = KC + U * X, END = KS + U;
    static constexpr int A = D::OFF_FX, B = D::OFF_FU;       // within a dense block: A[r][c] = A + r * X + c
};

// called by the first kCoopThreads threads of the block
template <typename M, typename PV>
__device__ __forceinline__ void solo_backward_coop(const PV& P, double mu, double dt, int T, int TP, const double* sx,
                                                   const double* su, const double* shi, const double* slo,
                                                   const double* ssc, const double* srec, double* sk, double* sK,
                                                   double* scratch, const int8_t* s_slot) {
    using D = Dims<M>;
    using S = SoloScratch<M>;
    constexpr int X = D::X, U = D::U, NSC = D::NSC;
    constexpr int N1 = X + U + X * X + X * U;                // Qx, Qu, VA = Vxx A, VB = Vxx B
    constexpr int N2 = X * X + U * U + U * X;                // Qxx, Quu, Qux
    static_assert(N1 <= kCoopLanes && N2 <= kCoopLanes && X * X <= 64, "one output per lane");
    const int tid = threadIdx.x;
    const bool expander = tid >= kCoopLanes;

    // ---- what this lane computes in every stage ------------------------------------------
    DotJob j1{0, 0, 0, 0, -1, -1}, j2{0, 0, 0, 0, -1, -1};
    {
        const int o = tid;
        if (o < X) {                                         // Qx[i] = lx[i] + sum_r A[r][i] Vx[r]
            j1 = DotJob{S::A + o, X, S::VX, 1, D::OFF_LX + o, S::QX + o};
        } else if (o < X + U) {                              // Qu[i] = lu[i] + sum_r B[r][i] Vx[r]
            const int i = o - X;
            j1 = DotJob{S::B + i, U, S::VX, 1, D::OFF_LU + i, S::QU + i};
        } else if (o < X + U + X * X) {                      // VA[i][j] = sum_r A[r][j] Vxx[i][r]
            const int e = o - X - U, i = e / X, c = e % X;
            j1 = DotJob{S::A + c, X, S::VXX + i * X, 1, -1, S::VA + e};
        } else if (o < N1) {                                 // VB[i][j] = sum_r B[r][j] Vxx[i][r]
            const int e = o - X - U - X * X, i = e / U, c = e % U;
            j1 = DotJob{S::B + c, U, S::VXX + i * X, 1, -1, S::VB + e};
        }
        if (o < X * X) {                                     // Qxx[i][j] = lxx[i][j] + sum_r A[r][p] VA[r][q], (p, q) = (max, min)
            const int i = o / X, c = o % X, p = i > c ? i : c, q = i > c ? c : i;
            j2 = DotJob{S::A + p, X, S::VA + q, X, D::OFF_LXX + o, S::QXX + o};
        } else if (o < X * X + U * U) {                      // Quu[i][j] = luu[i][j] + sum_r B[r][p] VB[r][q]
            const int e = o - X * X, i = e / U, c = e % U, p = i > c ? i : c, q = i > c ? c : i;
            j2 = DotJob{S::B + p, U, S::VB + q, U, D::OFF_LUU + e, S::QUU + e};
        } else if (o < N2) {                                 // Qux[i][j] = lux[i][j] + sum_r B[r][i] VA[r][j]
            const int e = o - X * X - U * U, i = e / X, c = e % X;
            j2 = DotJob{S::B + i, U, S::VA + c, X, D::OFF_LUX + e, S::QUX + e};
        }
    }
    // the expanding warp: rows of the compact record behind its dense entries (-1: 0.0, -2: 1.0, -3: none)
    constexpr int RE = (D::DENSE + 31) / 32;
    int esrc[RE];
#pragma unroll
    for (int k = 0; k < RE; ++k) {
        const int e = (tid & 31) + 32 * k;
        esrc[k] = e < D::DENSE ? (int)s_slot[e] : -3;
    }
    auto expand = [&](int t) {
        double* dst = scratch + ((t & 1) ? S::DENSE1 : S::DENSE0) + (tid & 31);
#pragma unroll
        for (int k = 0; k < RE; ++k) {
            const int s = esrc[k];
            if (s > -3) dst[32 * k] = s >= 0 ? srec[s * TP + t] : (s == -2 ? 1.0 : 0.0);
        }
    };

    if (expander) {
        expand(T - 1);
    } else if (tid == 0) {
        double xT[X], sc[D::NSCs], Vx[X], Vxx[X * X];
#pragma unroll
        for (int i = 0; i < X; ++i) xT[i] = sx[i * TP + T];
#pragma unroll
        for (int j = 0; j < NSC; ++j) sc[j] = ssc[j * TP + T];
        M::end_derivatives(P, xT, sc, double(T), dt, Vx, Vxx);
#pragma unroll
        for (int i = 0; i < X; ++i) scratch[S::VX + i] = Vx[i];
#pragma unroll
        for (int i = 0; i < X * X; ++i) scratch[S::VXX + i] = Vxx[i];
    }
    named_barrier(2, kCoopThreads);

    for (int t = T - 1; t >= 0; --t) {
        if (expander) {
            if (t > 0) expand(t - 1);
        } else {
            const double* dense = scratch + ((t & 1) ? S::DENSE1 : S::DENSE0);
            solo_dot<X>(scratch, dense, j1);
            named_barrier(1, kCoopLanes);
            solo_dot<X>(scratch, dense, j2);
            named_barrier(1, kCoopLanes);
            // gains, box limits and the value function of this stage (optim.c:243-291, 950-984) without a
            // barrier in between: every lane inverts Quu + mu I and forms k itself (identical values),
            // then only the columns of K its own output needs.  Vxx on warps 0-1, Vx on warp 2.
            double Quu[U][U], Qu[U], Mi[U][U], k[U];
            bool open_row[U];                        // false: the feed-forward step hit a box limit, K row is 0
#pragma unroll
            for (int i = 0; i < U; ++i) {
                Qu[i] = scratch[S::QU + i];
#pragma unroll
                for (int j = 0; j < U; ++j) Quu[i][j] = scratch[S::QUU + i * U + j];
            }
            gain_inverse<U, double>(Quu, mu, Mi);
#pragma unroll
            for (int d = 0; d < U; ++d) {
                k[d] = gain_row<U, double>(Mi, d, Qu);
                const double ub = su[d * TP + t], hib = shi[d * TP + t], lob = slo[d * TP + t];
                const double cand = ub + k[d];
                open_row[d] = true;
                if (cand > hib) {
                    k[d] = hib - ub;
                    open_row[d] = false;
                }
                if (cand < lob) {
                    k[d] = lob - ub;
                    open_row[d] = false;
                }
            }
            const double* sQux = scratch + S::QUX;
            // column `col` of the gain matrix K
            auto gain_column = [&](int col, double (&Kc)[U]) {
                double q[U];
#pragma unroll
                for (int c = 0; c < U; ++c) q[c] = sQux[c * X + col];
#pragma unroll
                for (int d = 0; d < U; ++d) Kc[d] = open_row[d] ? gain_row<U, double>(Mi, d, q) : 0.0;
            };
            if (tid < U) sk[tid * TP + t] = k[tid];
            if (tid >= 64 && tid < 64 + X) {         // warp 2 also stores the gains of the stage
                double Kc[U];
                gain_column(tid - 64, Kc);
#pragma unroll
                for (int d = 0; d < U; ++d) sK[(d * X + tid - 64) * TP + t] = Kc[d];
            }
            if (tid < 64) {
                const int oc = tid < X * X ? tid : 0;
                const int i = oc / X, j = oc % X;
                double Ki[U], Kj[U], kq_ij = 0.0, kq_ji = 0.0, kquu[U];
                gain_column(i, Ki);
                gain_column(j, Kj);
#pragma unroll
                for (int c = 0; c < U; ++c) kq_ij += Ki[c] * sQux[c * X + j];
#pragma unroll
                for (int c = 0; c < U; ++c) kq_ji += Kj[c] * sQux[c * X + i];
#pragma unroll
                for (int jj = 0; jj < U; ++jj) {
                    double acc = 0.0;
#pragma unroll
                    for (int c = 0; c < U; ++c) acc += Ki[c] * Quu[c][jj];
                    kquu[jj] = acc;
                }
                double v = kq_ji + kq_ij;
#pragma unroll
                for (int c = 0; c < U; ++c) v += kquu[c] * Kj[c];
                v += scratch[S::QXX + oc];
                // (nothing in this step reads Vxx / Vx: they are overwritten in place)
                if (tid < X * X) scratch[S::VXX + tid] = v;
            } else {
                const int l = tid - 64, i = l < X ? l : 0;
                double Ki[U], kquu[U];
                gain_column(i, Ki);
#pragma unroll
                for (int jj = 0; jj < U; ++jj) {
                    double acc = 0.0;
#pragma unroll
                    for (int c = 0; c < U; ++c) acc += Ki[c] * Quu[c][jj];
                    kquu[jj] = acc;
                }
                double v = 0.0;
#pragma unroll
                for (int c = 0; c < U; ++c) v += kquu[c] * k[c];
#pragma unroll
                for (int c = 0; c < U; ++c) v += Ki[c] * Qu[c];
#pragma unroll
                for (int c = 0; c < U; ++c) v += sQux[c * X + i] * k[c];
                v += scratch[S::QX + i];
                if (l < X) scratch[S::VX + i] = v;
            }
        }
        named_barrier(2, kCoopThreads);
    }
}

template <typename M, int kScheme>
__global__ void __launch_bounds__(kSoloThreads, 1)
solo_update_kernel(const __grid_constant__ tplb_batch q, Workspace ws) {
    using D = Dims<M>;
    using L = SoloLayout<M>;
    constexpr int X = D::X, U = D::U, C = D::C, NSC = D::NSC, NC = D::COMPACT;
    constexpr bool kCoop = X * X >= 25 && X * X <= 64;   // narrower recursions stay on one lane
#ifdef TPLB_SOLO_ROLLOUT_LANES
    constexpr int kRolloutLanes = TPLB_SOLO_ROLLOUT_LANES;
#else
    constexpr int kRolloutLanes = M::DYNAMICS_LOOKUPS > 0 ? 2 : 8;
#endif
    extern __shared__ __align__(16) double solo_smem[];
    __shared__ SoloStatus st;
    __shared__ int8_t s_slot[D::DENSE];

    const int b = blockIdx.x, tid = threadIdx.x, n = blockDim.x, B = q.batch;
    const int T = q.horizons ? q.horizons[b] : q.horizon;
    const int TP = T + 1;
    double* sx = solo_smem;
    double* su = sx + X * TP;
    double* sk = su + U * TP;
    double* sK = sk + U * TP;
    double* shi = sK + U * X * TP;
    double* slo = shi + U * TP;
    double* slam = slo + U * TP;
    double* ssc = slam + D::Cs * TP;
    double* srec = ssc + D::NSCs * TP;
    double* scx = srec + NC * TP;                    // [8][X][TP]
    double* scu = scx + kAlphas * X * TP;            // [8][U][TP]
    double* sct = scu + kAlphas * U * TP;            // [8][TP]
    double* scratch = sct + kAlphas * TP;

    // parameters on chip: scalars in registers, the scene's array rows staged behind the stage slots
    const int scene = __ldg(q.scene_index + b);
    StagedParamView<double, M::NUM_SCALARS, M::NUM_ARRAYS> P;
    {
        const ParamView<double> G = param_view_scene<double>(q, scene);
#pragma unroll
        for (int i = 0; i < M::NUM_SCALARS; ++i) P.cached[i] = G.scalar(i);
        double* dst = scratch + L::SCRATCH;
#pragma unroll
        for (int a = 0; a < M::NUM_ARRAYS; ++a) {
            const int len = q.array_len[a];
            const double* src = G.row(a);
            for (int i = tid; i < len; i += n) dst[i] = __ldg(src + i);
            P.rows[a] = dst;
            P.len[a] = len;
            P.cols[a] = q.array_cols[a];
            dst += len;
        }
    }
    __syncthreads();                                 // the stage constants below look the rows up
    const double dt = q.dt;
    double w[D::Cs], lim[D::Cs];
#pragma unroll
    for (int c = 0; c < C; ++c) {
        w[c] = q.barrier_weight[(size_t)c * B + b];
        lim[c] = q.lg_mult_limit[(size_t)c * B + b];
    }

#ifdef TPLB_SOLO_TIMING
    long long tick_ = clock64();
#endif
    // ---- load the problem --------------------------------------------------------------
    for (int e = tid; e < D::DENSE; e += n) s_slot[e] = (int8_t)M::deriv_slot(e);
    for (int i = tid; i < X; i += n) sx[i * TP] = q.x[(size_t)i * B + b];
    for (int idx = tid; idx < T * U; idx += n) {
        const int t = idx / U, d = idx % U;
        const size_t g = ((size_t)t * U + d) * B + b;
        su[d * TP + t] = q.u[g];
        shi[d * TP + t] = q.u_max[g];
        slo[d * TP + t] = q.u_min[g];
    }
    for (int idx = tid; idx < T * C; idx += n) {
        const int t = idx / C, c = idx % C;
        slam[c * TP + t] = q.lagrange_multiplier[((size_t)t * C + c) * B + b];
    }
    for (int t = tid; t <= T; t += n) {
        double sc[D::NSCs];
        M::stage_constants(P, double(t), dt, sc);
#pragma unroll
        for (int j = 0; j < NSC; ++j) ssc[j * TP + t] = sc[j];
    }
    if (tid == 0) {
        st.traj_costs = q.traj_costs[b];
        st.alpha = q.alpha[b];
        st.mu = q.mu[b];
        st.mu_step = q.mu_step[b];
        st.iterations = q.iterations[b];
        st.lg_iterations = q.lg_iterations[b];
        st.trajectory_changed = q.trajectory_changed[b];
        st.improved = q.improved[b];
        st.termination_condition = q.termination_condition[b];
        st.running = 0;
        st.winner = -1;
        st.last_tried = ws.last_tried[b];
        st.c_lin = 0;
        st.c_bwd = 0;
        st.c_roll = 1;                               // the initial rollout
        st.lam_zero = 0;
    }
    __syncthreads();

    // stage t (t < T) or end cost (t == T) of trajectory (xs, us), multipliers as stored
    auto cost_term = [&](const double* xs, const double* us, int t) {
        double x[X], sc[D::NSCs], c;
#pragma unroll
        for (int i = 0; i < X; ++i) x[i] = xs[i * TP + t];
#pragma unroll
        for (int j = 0; j < NSC; ++j) sc[j] = ssc[j * TP + t];
        if (t < T) {
            double u[U], lam[D::Cs];
#pragma unroll
            for (int d = 0; d < U; ++d) u[d] = us[d * TP + t];
#pragma unroll
            for (int cc = 0; cc < C; ++cc) lam[cc] = slam[cc * TP + t];
            M::stage_cost(P, x, u, lam, w, sc, double(t), dt, &c);
        } else {
            M::end_cost(P, x, sc, double(T), dt, &c);
        }
        return c;
    };

    // ---- initial rollout and its cost (optim.c:1096-1111) ---------------------------------
    if (tid == 0) {
        double xn[X];
#pragma unroll
        for (int i = 0; i < X; ++i) xn[i] = sx[i * TP];
        for (int t = 0; t < T; ++t) {
            double un[U], sc[D::NSCs], xnext[X];
#pragma unroll
            for (int d = 0; d < U; ++d) un[d] = su[d * TP + t];
#pragma unroll
            for (int j = 0; j < NSC; ++j) sc[j] = ssc[j * TP + t];
            step_state<M, kScheme>(P, xn, un, sc, double(t), dt, xnext);
#pragma unroll
            for (int i = 0; i < X; ++i) {
                xn[i] = xnext[i];
                sx[i * TP + t + 1] = xnext[i];
            }
        }
    }
    __syncthreads();
    for (int t = tid; t <= T; t += n) sct[t] = cost_term(sx, su, t);
    __syncthreads();
    if (tid == 0) {
        double total = 0.0;
        for (int t = 0; t <= T; ++t) total += sct[t];
        st.traj_costs = total;
    }

    SOLO_TICK(0);
    int lg = 0;
    for (; lg < q.max_lg_iterations; ++lg) {
        // ---- multiplier update, per-outer-iteration reset (optim.c:1115-1136) -------------
        __syncthreads();
        if (tid == 0) {
            st.trajectory_changed = 1;
            st.improved = 0;
            st.iterations = 0;
            st.running = 1;
            bool zero = C > 0;
#pragma unroll
            for (int c = 0; c < C; ++c) zero = zero && lim[c] == 0.0;
            st.lam_zero = zero;
        }
        if (C > 0) {
            for (int t = tid; t < T; t += n) {
                double x[X], u[U], lam[D::Cs], g[D::Cs], sc[D::NSCs];
#pragma unroll
                for (int i = 0; i < X; ++i) x[i] = sx[i * TP + t];
#pragma unroll
                for (int d = 0; d < U; ++d) u[d] = su[d * TP + t];
#pragma unroll
                for (int c = 0; c < C; ++c) {
                    lam[c] = slam[c * TP + t];
                    g[c] = 0.0;
                }
#pragma unroll
                for (int j = 0; j < NSC; ++j) sc[j] = ssc[j * TP + t];
                M::constraints(P, x, u, lam, w, sc, double(t), dt, g);
#pragma unroll
                for (int c = 0; c < C; ++c) {
                    double v = lam[c] + w[c] * g[c];
                    v = (0.0 > v) ? 0.0 : v;
                    slam[c * TP + t] = (lim[c] < v) ? lim[c] : v;
                }
            }
        }
        __syncthreads();

        for (int s = 0; s < q.max_iterations; ++s) {
            if (!st.running) break;                  // block-uniform (read after a barrier)
            // ---- linearisation of every stage (optim.c:896-912) ---------------------------
            if (st.trajectory_changed) {
                for (int t = tid; t < T; t += n) {
                    double x[X], u[U], lam[D::Cs], sc[D::NSCs], rec[NC];
#pragma unroll
                    for (int i = 0; i < X; ++i) x[i] = sx[i * TP + t];
#pragma unroll
                    for (int d = 0; d < U; ++d) u[d] = su[d * TP + t];
#pragma unroll
                    for (int c = 0; c < C; ++c) lam[c] = slam[c * TP + t];
#pragma unroll
                    for (int j = 0; j < NSC; ++j) sc[j] = ssc[j * TP + t];
                    stage_record<M, double>(P, x, u, lam, w, sc, double(t), dt, rec);
#pragma unroll
                    for (int e = 0; e < M::DERIV_COMPACT; ++e) srec[e * TP + t] = rec[e];
                }
            }
            __syncthreads();
            SOLO_TICK(1);
            // ---- backward Riccati sweep (optim.c:914-985) ------------------------------------
            if (tid == 0) {
                st.iterations = s + 1;               // optim.c:894
                st.c_lin += st.trajectory_changed ? 1 : 0;
                st.c_bwd += 1;
                st.trajectory_changed = 0;           // optim.c:911
            }
            if constexpr (kCoop) {
                if (tid < kCoopThreads)
                    solo_backward_coop<M>(P, st.mu, dt, T, TP, sx, su, shi, slo, ssc, srec, sk, sK, scratch, s_slot);
            } else {
                if (tid == 0) solo_backward_serial<M>(P, st.mu, dt, T, TP, sx, su, shi, slo, ssc, srec, sk, sK);
            }
            __syncthreads();
            SOLO_TICK(2);
            // ---- line search: 8 step sizes on 8 lanes (optim.c:732-775, 859-873) ------------
            // kRolloutLanes step sizes per warp: models whose dynamics branch (lookups, fmod) would
            // serialise eight diverging lanes; their step sizes spread over the four schedulers instead
            if ((tid & 31) < kRolloutLanes && tid < 32 * (kAlphas / kRolloutLanes)) {
                const int a = (tid >> 5) * kRolloutLanes + (tid & 31);
                double tens = 1.0;
                for (int i = 0; i < a; ++i) tens *= 10.0;
                const double alpha = 1.0 / tens;
                double* cx = scx + a * X * TP;
                double* cu = scu + a * U * TP;
                double xn[X];
#pragma unroll
                for (int i = 0; i < X; ++i) {
                    xn[i] = sx[i * TP];
                    cx[i * TP] = xn[i];
                }
                for (int t = 0; t < T; ++t) {
                    double un[U], sc[D::NSCs], xnext[X];
#pragma unroll
                    for (int d = 0; d < U; ++d) {
                        const double ud = su[d * TP + t];
                        double v = sk[d * TP + t] * alpha + ud;
#pragma unroll
                        for (int j = 0; j < X; ++j) v += sK[(d * X + j) * TP + t] * (xn[j] - sx[j * TP + t]);
                        const double hi = shi[d * TP + t], lo = slo[d * TP + t];
                        const double capped = (hi < v) ? hi : v;              // optim.c:755-758
                        un[d] = (lo > capped) ? lo : capped;
                        cu[d * TP + t] = un[d];
                    }
#pragma unroll
                    for (int j = 0; j < NSC; ++j) sc[j] = ssc[j * TP + t];
                    step_state<M, kScheme>(P, xn, un, sc, double(t), dt, xnext);
#pragma unroll
                    for (int i = 0; i < X; ++i) {
                        xn[i] = xnext[i];
                        cx[i * TP + t + 1] = xnext[i];
                    }
                }
            }
            __syncthreads();
            SOLO_TICK(3);
            for (int idx = tid; idx < kAlphas * TP; idx += n) {
                const int a = idx / TP, t = idx - a * TP;
                sct[idx] = cost_term(scx + a * X * TP, scu + a * U * TP, t);
            }
            __syncthreads();
            SOLO_TICK(4);
            if (tid < kAlphas) {
                double total = 0.0;
                for (int t = 0; t <= T; ++t) total += sct[tid * TP + t];
                st.total[tid] = total;
            }
            __syncthreads();
            if (tid == 0) {
                const double before = st.traj_costs;
                double now = before;
                int win = -1;
                for (int i = kAlphas - 1; i >= 0; --i) {
                    const double c = st.total[i];
                    if (c < before && isfinite(c) && c >= 0.0) {             // optim.c:842
                        win = i;
                        now = c;
                    }
                }
                const SearchStatus ss{st.traj_costs, st.alpha, st.mu, st.mu_step, st.trajectory_changed,
                                      st.improved, st.termination_condition, st.running, st.winner,
                                      st.last_tried, st.c_roll};
                conclude_core(ss, true, q.min_rel_cost_change, win, before, now);
            }
            __syncthreads();
            // ---- the accepted step becomes the trajectory (optim.c:844-848) -----------------
            if (st.winner >= 0) {
                const double* cx = scx + st.winner * X * TP;
                const double* cu = scu + st.winner * U * TP;
                for (int idx = tid; idx < X * TP; idx += n) {
                    if (q.keep_previous) {
                        const int i = idx / TP, t = idx - i * TP;
                        q.prev_x[((size_t)t * X + i) * B + b] = sx[idx];
                    }
                    sx[idx] = cx[idx];
                }
                for (int idx = tid; idx < U * TP; idx += n) {
                    const int d = idx / TP, t = idx - d * TP;
                    if (t < T) {
                        if (q.keep_previous) q.prev_k[((size_t)t * U + d) * B + b] = sk[idx];
                        su[idx] = cu[idx];
                    }
                }
            }
            __syncthreads();
            SOLO_TICK(5);
        }
    }
    __syncthreads();

    // ---- results -----------------------------------------------------------------------------
    for (int idx = tid; idx < X * TP; idx += n) {
        const int i = idx / TP, t = idx - i * TP;
        q.x[((size_t)t * X + i) * B + b] = sx[idx];
    }
    for (int idx = tid; idx < U * TP; idx += n) {
        const int d = idx / TP, t = idx - d * TP;
        if (t < T) {
            const size_t g = ((size_t)t * U + d) * B + b;
            q.u[g] = su[idx];
            q.k[g] = sk[idx];
        }
    }
    for (int idx = tid; idx < U * X * TP; idx += n) {
        const int e = idx / TP, t = idx - e * TP;
        if (t < T) q.K[((size_t)t * U * X + e) * B + b] = sK[idx];
    }
    for (int idx = tid; idx < C * TP; idx += n) {
        const int c = idx / TP, t = idx - c * TP;
        if (t < T) q.lagrange_multiplier[((size_t)t * C + c) * B + b] = slam[idx];
    }
    for (int idx = tid; idx < NC * TP; idx += n) {                 // records of the last linearisation
        const int e = idx / TP, t = idx - e * TP;
        if (t < T && e < M::DERIV_COMPACT) ws.deriv[((size_t)t * NC + e) * B + b] = srec[idx];
    }
    {                                                                // next_x / next_u of the last search
        const int a = st.last_tried;
        if (a >= 0 && st.c_bwd > 0) {
            double* gx = ws.cand_x + (size_t)a * (q.t_max + 1) * X * B + b;
            double* gu = ws.cand_u + (size_t)a * q.t_max * U * B + b;
            for (int idx = tid; idx < X * TP; idx += n) {
                const int i = idx / TP, t = idx - i * TP;
                gx[((size_t)t * X + i) * B] = scx[a * X * TP + idx];
            }
            for (int idx = tid; idx < U * TP; idx += n) {
                const int d = idx / TP, t = idx - d * TP;
                if (t < T) gu[((size_t)t * U + d) * B] = scu[a * U * TP + idx];
            }
        }
    }
    if (tid == 0) {
        if (st.iterations == q.max_iterations) st.termination_condition = 1;   // optim.c:1147-1149
        q.traj_costs[b] = st.traj_costs;
        q.alpha[b] = st.alpha;
        q.mu[b] = st.mu;
        q.mu_step[b] = st.mu_step;
        q.iterations[b] = st.iterations;
        q.lg_iterations[b] = lg;
        q.trajectory_changed[b] = st.trajectory_changed;
        q.improved[b] = st.improved;
        q.termination_condition[b] = st.termination_condition;
        ws.running[b] = st.running;
        ws.winner[b] = -1;                           // already installed
        ws.last_tried[b] = st.last_tried;
        ws.counters[b] = st.c_lin;
        ws.counters[(size_t)B + b] = st.c_bwd;
        ws.counters[(size_t)2 * B + b] = st.c_roll;
        ws.lam_zero[b] = st.lam_zero;
        if (b == 0) *ws.records_f32 = 0;
    }
    SOLO_TICK(6);
}

}  // namespace tplb
