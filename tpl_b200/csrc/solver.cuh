// Batched iLQR kernels for sm_100a.  One instantiation per generated `Model`.
//
// Restates, for B independent problems, the solver of
// /root/reference/library/tpl/optim/templates/optim.c ("optim.c"):
//
//   stage_constants_kernel  per-(scene, stage) interpolation lookups, once per update()
//   rollout_kernel<Init>    optim.c:1096-1107    x[t+1] = F(x[t], u[t])            [problem parallel]
//   rollout_kernel<Search>  optim.c:732-775      step sizes alpha_i = 10^-i side by side:
//                                                u' = clip(u + alpha k + K (x' - x)), x' = F(x', u');
//                                                <kCost> also adds up the stage costs (:776-789)
//   stage_cost*_kernel      optim.c:776-789, 1105-1111   cost terms of every (candidate, stage)
//                                                                                   [stage parallel]
//   init_cost_kernel        optim.c:1106-1111    ordered sum -> trajCosts
//   multiplier_kernel       optim.c:1115-1136    multiplier update, outer-iteration reset
//   linearize_kernel        optim.c:896-912      derivative records of every stage  [stage parallel]
//   backward_kernel         optim.c:914-985      Riccati sweep, gains, box limits   [problem parallel]
//   select_kernel<round>    optim.c:840-873, 987-1006   ordered cost sums, first improving step
//                                                wins, mu schedule, relative-change stop.  Round 1
//                                                looks at alpha = 1, 0.1; only problems where both
//                                                fail go to round 2 (alpha = 1e-2 .. 1e-7)
//   accept_kernel           optim.c:844-848      copy the winning candidate         [stage parallel]
//                                                (folded into the next linearize_kernel while
//                                                launches are latency-bound)
//   finalize_kernel         optim.c:1145-1149    termination flag
//
// cabi.cu strings them together in one of two sequences with identical results: one for a
// batch that cannot fill the GPU (fewest launches, all 8 step sizes rolled out at once) and
// one for a full GPU (fewest bytes: two-round rollouts that sum their own costs).
//
// Data layout: structure of arrays, problem index fastest — a warp of consecutive
// problems reads/writes 256 contiguous bytes for every (stage, component).
//
// What is sequential stays sequential per problem (dynamics chain, Riccati recursion,
// cost summation order t = 0..T then end cost, first-hit line search); everything that
// the reference evaluates stage by stage without a dependency is evaluated stage
// parallel.  Derivative entries that are identically 0 or 1 for the model are neither
// stored nor multiplied: skipping `+ 0*v` and `1*v` leaves every sum bit-identical.
#pragma once

#include <cstdint>
#include <type_traits>
#include <cuda_runtime.h>

#include "../../include/tplb200.h"
#include "device_math.cuh"

namespace tplb {

constexpr int kAlphas = TPLB_LINE_SEARCH_STEPS;
constexpr int kRound1 = 2;      // step sizes of the first line-search round (alpha = 1, 0.1)

// Horizon of problem b: its own (tplb_batch.horizons, optim.c:508-622 — every reference object owns
// its T) or the batch's.  Stage-parallel kernels are launched for the largest horizon of the batch
// (q.horizon) and rows past a problem's own horizon return at once.
__device__ __forceinline__ int horizon_of(const tplb_batch& q, int b) {
    return q.horizons ? __ldg(q.horizons + b) : q.horizon;
}

// Problem handled by work item `idx`: the item itself, or an entry of the pending list.
__device__ __forceinline__ int problem_of(const int32_t* list, const int32_t* count, int idx, int B) {
    if (!list) return idx < B ? idx : -1;
    return idx < *count ? list[idx] : -1;
}

// The derivative records never leave the solver and are stored in the type they are computed
// in: fp64, or fp32 in TPLB_FP32 mode (the buffer keeps its fp64 size; fp32 uses the first half).
// Everything else — x, u, k, K, multipliers, line-search candidates, cost terms and sums — is fp64
// in both modes.
template <typename R> struct Scratch { using type = double; };
template <> struct Scratch<float> { using type = float; };
template <typename R> using scratch_t = typename Scratch<R>::type;
template <typename S> __host__ __device__ __forceinline__ S* scratch(double* p) { return reinterpret_cast<S*>(p); }

template <typename M>
struct Dims {
    static constexpr int X = M::X, U = M::U, C = M::C;
    static constexpr int Cs = C > 0 ? C : 1;
    static constexpr int NSC = M::NUM_STAGE_CONSTS;
    static constexpr int NSCs = NSC > 0 ? NSC : 1;
    // dense derivative record of one stage (the layout of the fx..lux views)
    static constexpr int OFF_FX = 0;
    static constexpr int OFF_FU = OFF_FX + X * X;
    static constexpr int OFF_LX = OFF_FU + X * U;
    static constexpr int OFF_LU = OFF_LX + X;
    static constexpr int OFF_LXX = OFF_LU + U;
    static constexpr int OFF_LUU = OFF_LXX + X * X;
    static constexpr int OFF_LUX = OFF_LUU + U * U;
    static constexpr int DENSE = OFF_LUX + U * X;
    static constexpr int COMPACT = M::DERIV_COMPACT > 0 ? M::DERIV_COMPACT : 1;
    static_assert(DENSE == M::DERIV_DENSE, "generated layout does not match the solver's");
};

// Scratch carved out of tplb_batch.workspace.
struct Workspace {
    double* stage_consts;  // [t_max+1][NSC][S]
    double* deriv;         // [t_max][COMPACT][B]   compact derivative records
    double* cand_x;        // [8][t_max+1][X][B]    line-search candidates
    double* cand_u;        // [8][t_max][U][B]
    double* cost_terms;    // [8][t_max+1][B]       stage costs (and end cost) of every candidate
    double* cand_cost;     // [8][B]
    int32_t* winner;       // [B]  index of the accepted alpha, -1 = none
    int32_t* running;      // [B]  inner loop still active
    int32_t* counters;     // [3][B] work the reference would have done in this update():
                           //        linearisations, backward sweeps, sequential rollouts
    int32_t* pending;      // [B]  problems whose round-1 step sizes all failed (unordered list)
    int32_t* pending_count;// [1]
    int32_t* records_f32;  // [1]  1: the derivative records were written as fp32
    int32_t* last_tried;   // [B]  candidate the last line search of the problem ended on (the reference's
                           //      next_x / next_u): the accepted one, or alpha = 1e-7 after a failed search; -1: none yet
    int32_t* lam_zero;     // [B]  1: every multiplier limit of the problem is 0, so after the
                           //      multiplier update lambda is identically 0 (pure penalty — the setting of
                           //      all shipped callers) and need not be read
};

__host__ __device__ inline size_t align_up(size_t v) { return (v + 255) & ~size_t(255); }

template <typename M>
__host__ __device__ inline Workspace carve(void* base, int B, int S, int t_max, size_t* total = nullptr) {
    using D = Dims<M>;
    char* p = static_cast<char*>(base);
    size_t off = 0;
    Workspace w;
    auto take = [&](size_t bytes) { char* r = p + off; off += align_up(bytes); return r; };
    w.stage_consts = reinterpret_cast<double*>(take(sizeof(double) * (size_t)(t_max + 1) * D::NSCs * S));
    w.deriv = reinterpret_cast<double*>(take(sizeof(double) * (size_t)t_max * D::COMPACT * B));
    w.cand_x = reinterpret_cast<double*>(take(sizeof(double) * (size_t)kAlphas * (t_max + 1) * D::X * B));
    w.cand_u = reinterpret_cast<double*>(take(sizeof(double) * (size_t)kAlphas * t_max * D::U * B));
    w.cost_terms = reinterpret_cast<double*>(take(sizeof(double) * (size_t)kAlphas * (t_max + 1) * B));
    w.cand_cost = reinterpret_cast<double*>(take(sizeof(double) * (size_t)kAlphas * B));
    w.winner = reinterpret_cast<int32_t*>(take(sizeof(int32_t) * (size_t)B));
    w.running = reinterpret_cast<int32_t*>(take(sizeof(int32_t) * (size_t)B));
    w.counters = reinterpret_cast<int32_t*>(take(sizeof(int32_t) * (size_t)3 * B));
    w.pending = reinterpret_cast<int32_t*>(take(sizeof(int32_t) * (size_t)B));
    w.pending_count = reinterpret_cast<int32_t*>(take(sizeof(int32_t)));
    w.records_f32 = reinterpret_cast<int32_t*>(take(sizeof(int32_t)));
    w.lam_zero = reinterpret_cast<int32_t*>(take(sizeof(int32_t) * (size_t)B));
    w.last_tried = reinterpret_cast<int32_t*>(take(sizeof(int32_t) * (size_t)B));
    if (total) *total = off;
    return w;
}

template <typename R>
__device__ __forceinline__ ParamView<R> param_view_scene(const tplb_batch& q, int scene) {
    ParamView<R> P;
    P.scalars = q.scalars;
    P.arrays = q.arrays;
    P.len = q.array_len;
    P.cols = q.array_cols;
    P.num_scenes = q.scenes;
    P.scene = scene;
    return P;
}

// 8-byte asynchronous copy global -> shared (LDGSTS).  Completion is tracked per thread by
// commit / wait groups, not by the register scoreboard, so a consumer of the previous
// batch never waits for the batch that was just issued.
__device__ __forceinline__ void async_copy8(double* smem_dst, const double* gmem_src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)),
                 "l"(gmem_src));
}
// the same from `gmem_base` + a literal byte offset (folded into the instruction)
template <long long kByteOffset>
__device__ __forceinline__ void async_copy8_at(double* smem_dst, const double* gmem_base) {
    asm volatile("cp.async.ca.shared.global [%0], [%1+%2], 8;" ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)),
                 "l"(gmem_base), "n"(kByteOffset));
}
__device__ __forceinline__ void async_commit() { asm volatile("cp.async.commit_group;"); }
__device__ __forceinline__ void async_wait_all_but_one() { asm volatile("cp.async.wait_group 1;" ::: "memory"); }

template <typename M, typename R>
__device__ __forceinline__ void load_stage_consts(const tplb_batch& q, const Workspace& ws, int scene, int t,
                                                  R* sc) {
    constexpr int NSC = M::NUM_STAGE_CONSTS;
#pragma unroll
    for (int j = 0; j < NSC; ++j) sc[j] = R(__ldg(ws.stage_consts + ((size_t)t * NSC + j) * q.scenes + scene));
}

// ---------------------------------------------------------------------------------
// stage constants: thread per (scene, stage)
// ---------------------------------------------------------------------------------
template <typename M>
__global__ void stage_constants_kernel(const __grid_constant__ tplb_batch q, Workspace ws) {
    constexpr int NSC = M::NUM_STAGE_CONSTS;
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    const int t = blockIdx.y;                                   // 0..T
    if (s >= q.scenes) return;
    if (NSC == 0) return;
    const ParamView<double> P = param_view_scene<double>(q, s);
    double sc[Dims<M>::NSCs];
    M::stage_constants(P, (double)t, q.dt, sc);
#pragma unroll
    for (int j = 0; j < NSC; ++j) ws.stage_consts[((size_t)t * NSC + j) * q.scenes + s] = sc[j];
}

// ---------------------------------------------------------------------------------
// integrators (optim.c:657-730); ctDynamics sees the same (t, dt) at all sub-stages
// ---------------------------------------------------------------------------------
template <typename M, int kScheme, typename R, typename PV>
__device__ __forceinline__ void step_state(const PV& P, const R* x, const R* u, const R* sc, R t, R h, R* out) {
    constexpr int X = M::X;
    R k1[X], k2[X], y[X];
    M::ct_dynamics(P, x, u, sc, t, h, k1);
    if constexpr (kScheme == TPLB_EULER) {
#pragma unroll
        for (int i = 0; i < X; ++i) out[i] = x[i] + k1[i] * h;
    } else if constexpr (kScheme == TPLB_HEUN) {
#pragma unroll
        for (int i = 0; i < X; ++i) y[i] = x[i] + k1[i] * h;
        M::ct_dynamics(P, y, u, sc, t, h, k2);
        const R hh = h / R(2);
#pragma unroll
        for (int i = 0; i < X; ++i) out[i] = x[i] + (k1[i] + k2[i]) * hh;
    } else {
        R k3[X], k4[X];
        const R hh = h / R(2);
#pragma unroll
        for (int i = 0; i < X; ++i) y[i] = x[i] + k1[i] * hh;
        M::ct_dynamics(P, y, u, sc, t, h, k2);
#pragma unroll
        for (int i = 0; i < X; ++i) y[i] = x[i] + k2[i] * hh;
        M::ct_dynamics(P, y, u, sc, t, h, k3);
#pragma unroll
        for (int i = 0; i < X; ++i) y[i] = x[i] + k3[i] * h;
        M::ct_dynamics(P, y, u, sc, t, h, k4);
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
// rollouts — the only part of the forward pass that is a chain in t.
//   kInit   : x[t+1] = F(x[t], u[t]) written in place (optim.c:1101-1107); block = PB problems
//             (reads q.x[0] and q.u only, so the in-place write of x[t+1] is safe under __ldg)
//   !kInit  : line search, block = PB problems x 8 step sizes; threadIdx.x = problem
//             (coalesced), threadIdx.y = i with alpha_i = 1 / 10^i (optim.c:863);
//             candidates go to ws.cand_x / ws.cand_u (the reference's next_x / next_u)
// Dynamic shared memory: rollout_smem_bytes<M, kInit, kCost>(PB, NA) (input staging, shared by a problem's candidates).
// ---------------------------------------------------------------------------------
// number of doubles one rollout stage reads per thread
template <typename M, bool kInit, bool kCost = false>
struct RolloutInputs {
    static constexpr int X = M::X, U = M::U, NSC = M::NUM_STAGE_CONSTS;
    static constexpr int O_U = 0, O_K = O_U + U, O_HI = O_K + (kInit ? 0 : U), O_LO = O_HI + (kInit ? 0 : U),
                         O_KK = O_LO + (kInit ? 0 : U), O_X = O_KK + (kInit ? 0 : U * X),
                         O_SC = O_X + (kInit ? 0 : X), O_LAM = O_SC + NSC,
                         COUNT = O_LAM + (kCost ? M::C : 0);   // kCost: the multipliers of the stage cost
};

// Batch sizes that get their own copy of the two hot kernel bodies with the batch size as a literal
// (BASELINE.json's 4096, 16384, 32768 and 65536, and the 8192 of its scene-sharded multi-start):
// every component stride i * B * 8 then is an immediate of the load / store / copy instruction
// instead of 64-bit address arithmetic (sweep: 1461 -> 1190 instructions per stage, first rollout
// round: 823 -> 658).  Any other size runs the general body; the results are bit-identical
// (tests/test_gpu_parity.py::test_literal_batch_bodies_match_the_general_body).
constexpr int kSpecialBatch0 = 4096, kSpecialBatch1 = 32768, kSpecialBatch2 = 65536, kSpecialBatch3 = 8192,
              kSpecialBatch4 = 16384;
// Slots of the staging ring.  Three: the copy of stage t+2 can never overwrite what a slower warp
// still reads for stage t.  Two, with a second barrier at the end of every stage, for the first
// line-search round of the throughput sequence (2 step sizes = 2 warps per block, where the barrier
// is cheap): that kernel runs 8 blocks per multiprocessor, and 8 x 23 KB of staging would leave the
// co-resident sweep too little L1 for its loads in flight (measured: -10 %; with two slots +5 %).
__host__ __device__ constexpr int rollout_slots(int step_sizes, bool with_cost) {
    return (step_sizes == 2 && with_cost) ? 2 : 3;
}

template <typename M, bool kInit, bool kCost = false>
__host__ __device__ inline size_t rollout_smem_bytes(int problems_per_block, int step_sizes) {
    using RI = RolloutInputs<M, kInit, kCost>;
    return sizeof(double) * rollout_slots(step_sizes, kCost) * RI::COUNT * problems_per_block;
}

// global address of staged item E at stage t for problem b (scene for the stage constants)
template <typename M, bool kInit, bool kCost, int E>
__device__ __forceinline__ const double* rollout_source(const tplb_batch& q, const Workspace& ws, size_t t,
                                                        int b, int scene) {
    using RI = RolloutInputs<M, kInit, kCost>;
    constexpr int X = M::X, U = M::U, NSC = M::NUM_STAGE_CONSTS, C = M::C;
    const size_t B = q.batch;
    if constexpr (E < RI::O_K) return q.u + (t * U + (E - RI::O_U)) * B + b;
    else if constexpr (E < RI::O_HI) return q.k + (t * U + (E - RI::O_K)) * B + b;
    else if constexpr (E < RI::O_LO) return q.u_max + (t * U + (E - RI::O_HI)) * B + b;
    else if constexpr (E < RI::O_KK) return q.u_min + (t * U + (E - RI::O_LO)) * B + b;
    else if constexpr (E < RI::O_X) return q.K + (t * (U * X) + (E - RI::O_KK)) * B + b;
    else if constexpr (E < RI::O_SC) return q.x + (t * X + (E - RI::O_X)) * B + b;
    else if constexpr (E < RI::O_LAM) return ws.stage_consts + (t * NSC + (E - RI::O_SC)) * q.scenes + scene;
    else return q.lagrange_multiplier + (t * C + (E - RI::O_LAM)) * B + b;
}

// The same address split into the first item of E's array at stage t (one computation per array,
// stage and thread) and E's position behind it in units of the batch stride: with a literal batch
// size KB the position is a literal byte offset of the copy instruction.  The stage constants
// are strided by the number of scenes and keep the general form.
template <typename M, bool kInit, bool kCost, int E>
struct RolloutItem {
    using RI = RolloutInputs<M, kInit, kCost>;
    static constexpr bool kBatchStrided = !(E >= RI::O_SC && E < RI::O_LAM);
    static constexpr int kFirst = E < RI::O_K ? RI::O_U : E < RI::O_HI ? RI::O_K : E < RI::O_LO ? RI::O_HI
                                : E < RI::O_KK ? RI::O_LO : E < RI::O_X ? RI::O_KK : E < RI::O_SC ? RI::O_X
                                : E < RI::O_LAM ? RI::O_SC : RI::O_LAM;
    static constexpr int kIndex = E - kFirst;
};

// copies items [E0, E1) of stage t into the ring slot `dst` (this problem's column)
// (`skip_lam`: the multipliers are known to be 0 and their ring entries were zeroed once)
template <typename M, bool kInit, bool kCost, int PB, int E0, int E1, int KB = 0>
__device__ __forceinline__ void rollout_fetch_range(const tplb_batch& q, const Workspace& ws, size_t t, int b,
                                                    int scene, double* dst, bool skip_lam) {
    if constexpr (E0 < E1) {
        using Item = RolloutItem<M, kInit, kCost, E0>;
        if (E0 < RolloutInputs<M, kInit, kCost>::O_LAM || !skip_lam) {
            if constexpr (KB > 0 && Item::kBatchStrided)
                async_copy8_at<(long long)Item::kIndex * KB * 8>(
                    dst + E0 * PB, rollout_source<M, kInit, kCost, Item::kFirst>(q, ws, t, b, scene));
            else
                async_copy8(dst + E0 * PB, rollout_source<M, kInit, kCost, E0>(q, ws, t, b, scene));
        }
        rollout_fetch_range<M, kInit, kCost, PB, E0 + 1, E1, KB>(q, ws, t, b, scene, dst, skip_lam);
    }
}

// candidate R of NA copies the R-th chunk of the items; `yy` is uniform in a warp, so the
// chain of comparisons is a uniform jump and every warp issues only its own copies
template <typename M, bool kInit, bool kCost, int PB, int NA, int KB = 0, int R = 0>
__device__ __forceinline__ void rollout_fetch(const tplb_batch& q, const Workspace& ws, size_t t, int b, int scene,
                                              double* dst, int yy, bool skip_lam) {
    using RI = RolloutInputs<M, kInit, kCost>;
    constexpr int CH = (RI::COUNT + NA - 1) / NA;
    if constexpr (R < NA) {
        if (yy == R) {
            constexpr int E0 = R * CH, E1 = (R + 1) * CH < RI::COUNT ? (R + 1) * CH : RI::COUNT;
            rollout_fetch_range<M, kInit, kCost, PB, (E0 < RI::COUNT ? E0 : RI::COUNT), E1, KB>(q, ws, t, b, scene,
                                                                                             dst, skip_lam);
        } else {
            rollout_fetch<M, kInit, kCost, PB, NA, KB, R + 1>(q, ws, t, b, scene, dst, yy, skip_lam);
        }
    }
}

// One thread = one (problem, step size); block = PB problems (threadIdx.x) x NA step sizes.
// All step sizes of a problem read the same inputs (u, k, bounds, K, x, stage constants of
// stage t), so the block stages them ONCE per problem: `smem` is a ring of rollout_slots()
// stages, [slot][COUNT][PB] doubles.  The candidates of a problem share the copies of a stage
// (each a contiguous chunk of the items), asynchronously and one stage ahead; one
// __syncthreads per stage publishes them.  With three slots the copy of stage t+2 can never
// overwrite what a slower warp still reads for stage t; with two, a second barrier ends the stage.  `live` == false: the thread only
// keeps the barriers.
// kCost: the thread also evaluates the stage costs of its candidate and adds them up in the
// reference's order (optim.c:773-790) -> ws.cand_cost; used when the GPU is full, where
// re-reading the candidates in a separate cost kernel costs more than the longer chain.
template <typename M, typename R, int PB, int NA, bool kInit, int kScheme, bool kCost, int KB = 0>
__device__ __forceinline__ void dev_rollout(const tplb_batch& q, const Workspace& ws, int b, int ai, bool live,
                                            unsigned char* smem) {
    if constexpr (KB > 0) __builtin_assume(q.batch == KB);
    using D = Dims<M>;
    using RI = RolloutInputs<M, kInit, kCost>;
    constexpr int kSlots = rollout_slots(NA, kCost);
    constexpr int X = D::X, U = D::U, NSC = D::NSC;
    const int B = q.batch;
    const int px = threadIdx.x, yy = threadIdx.y;
    double* stage_in = reinterpret_cast<double*>(smem) + px;

    if (kInit) {
        if (live) {
            ws.counters[b] = 0;
            ws.counters[(size_t)B + b] = 0;
            ws.counters[(size_t)2 * B + b] = 1;                  // the initial rollout itself
        }
    } else {
        live = live && ws.running[b];
    }
    const int T = q.horizon;                                 // block-uniform loop bound (barriers inside)
    const int Tb = live ? horizon_of(q, b) : 0;              // this problem's horizon
    const int scene = live ? __ldg(q.scene_index + b) : 0;
    CachedParamView<R, M::NUM_SCALARS> P;                    // scalar parameters in registers for the whole chain
    P.preload(param_view_scene<R>(q, scene));
    const bool second_order = q.use_quadratic_terms != 0;

    double tens = 1.0;
    for (int i = 0; i < ai; ++i) tens *= 10.0;
    const R alpha = R(1.0 / tens);

    using XS = std::conditional_t<kInit, double, scratch_t<R>>;   // the initial rollout writes q.x itself
    XS* cx = kInit ? reinterpret_cast<XS*>(q.x) + b
                   : scratch<XS>(ws.cand_x) + (size_t)ai * (q.t_max + 1) * X * B + b;
    XS* cu = kInit ? nullptr : scratch<XS>(ws.cand_u) + (size_t)ai * q.t_max * U * B + b;
    const int iB = B;                                        // component stride inside a stage

    // ring slot `buf` of this problem: item e at ring(buf)[e * PB]
    auto ring = [&](int buf) { return stage_in + buf * (RI::COUNT * PB); };
    const bool skip_lam = kCost && D::C > 0 && live && ws.lam_zero[b] != 0;
    auto fetch = [&](int t, int buf) {
        rollout_fetch<M, kInit, kCost, PB, NA, KB>(q, ws, (size_t)t, b, scene, ring(buf), yy, skip_lam);
    };
    if (kCost && skip_lam && yy == 0) {                      // the first barrier of the loop publishes the zeros
#pragma unroll
        for (int sl = 0; sl < kSlots; ++sl)
#pragma unroll
            for (int cc = 0; cc < D::C; ++cc) ring(sl)[(RI::O_LAM + cc) * PB] = 0.0;
    }

    R xn[X];
    if (live) {
#pragma unroll
        for (int i = 0; i < X; ++i) {
            xn[i] = R(q.x[(size_t)i * B + b]);
            if (!kInit) __stcs(cx + (size_t)i * B, (XS)xn[i]);
        }
    }
    double total = 0.0;
    R weight[D::Cs];
    if (kCost && live) {
#pragma unroll
        for (int cc = 0; cc < D::C; ++cc) weight[cc] = R(q.barrier_weight[(size_t)cc * B + b]);
    }

    if (live) fetch(0, 0);
    async_commit();
    int buf = 0;
    for (int t = 0; t < T; ++t) {
        const int nxt = buf + 1 == kSlots ? 0 : buf + 1;
        if (live && t + 1 < Tb) fetch(t + 1, nxt);
        async_commit();                                      // (possibly empty) group of stage t+1
        async_wait_all_but_one();                            // this thread's share of stage t has landed
        __syncthreads();                                     // ... and everybody else's

        if (live && t < Tb) {
            const double* in = ring(buf);                    // stage t of this problem, item e at in[e * PB]
            R un[U], xnext[X], sc[D::NSCs];
#pragma unroll
            for (int d = 0; d < U; ++d) {
                const R ud = R(in[(RI::O_U + d) * PB]);
                if (kInit) {
                    un[d] = ud;
                } else if (second_order) {
                    R v = R(in[(RI::O_K + d) * PB]) * alpha + ud;
#pragma unroll
                    for (int j = 0; j < X; ++j)
                        v += R(in[(RI::O_KK + d * X + j) * PB]) * (xn[j] - R(in[(RI::O_X + j) * PB]));
                    const R hi = R(in[(RI::O_HI + d) * PB]), lo = R(in[(RI::O_LO + d) * PB]);
                    const R capped = (hi < v) ? hi : v;               // optim.c:755-758
                    un[d] = (lo > capped) ? lo : capped;
                } else {
                    un[d] = ud - R(in[(RI::O_K + d) * PB]) * alpha;          // optim.c:803-804
                }
            }
#pragma unroll
            for (int j = 0; j < NSC; ++j) sc[j] = R(in[(RI::O_SC + j) * PB]);
            if (!kInit) {
                XS* cut = cu + (size_t)t * U * B;
#pragma unroll
                for (int d = 0; d < U; ++d) __stcs(cut + d * iB, (XS)un[d]);       // streaming: keep K, k, x, u in L2
            }
            if (kCost) {
                R lam[D::Cs], c;
#pragma unroll
                for (int cc = 0; cc < D::C; ++cc) lam[cc] = R(in[(RI::O_LAM + cc) * PB]);
                M::stage_cost(P, xn, un, lam, weight, sc, R(t), R(q.dt), &c);
                total += (double)c;
            }
            step_state<M, kScheme>(P, xn, un, sc, R(t), R(q.dt), xnext);
            XS* cxt = cx + (size_t)(t + 1) * X * B;
#pragma unroll
            for (int i = 0; i < X; ++i) {
                xn[i] = xnext[i];
                if (kInit) cxt[i * iB] = (XS)xnext[i];
                else __stcs(cxt + i * iB, (XS)xnext[i]);
            }
        }
        buf = nxt;
        if constexpr (kSlots == 2) __syncthreads();   // two slots: nobody may still read the slot the next fetch fills
    }
    if (kCost && live) {
        R sc[D::NSCs], c;
        load_stage_consts<M, R>(q, ws, scene, Tb, sc);
        M::end_cost(P, xn, sc, R(Tb), R(q.dt), &c);
        total += (double)c;
        ws.cand_cost[(size_t)ai * B + b] = total;
    }
}

// NA: step sizes per problem in this launch = blockDim.y (1 initial rollout, 8 all at once, 2 / 6 the
// two rounds of the throughput sequence)
// The first line-search round of the throughput sequence is built for 8 blocks of 64 threads per
// multiprocessor (128 registers; no spills up to 6 states): measured best together with the
// two-slot ring — 6 blocks (158 registers) and 10 blocks (96 registers) are both slower.
template <typename M, typename R, int PB, int NA, bool kInit, int kScheme, bool kCost = false>
__global__ void __launch_bounds__(PB * NA, (NA == 2 && kCost && M::X <= 6) ? 8 : 1)
rollout_kernel(const __grid_constant__ tplb_batch q, Workspace ws, int a_begin, const int32_t* list) {
    extern __shared__ __align__(16) unsigned char rollout_smem[];
    const int ai = kInit ? 0 : a_begin + threadIdx.y;
    if (list && blockIdx.x * PB >= *ws.pending_count) return;   // block-uniform: nothing pending here
    const int b = problem_of(list, ws.pending_count, blockIdx.x * PB + threadIdx.x, q.batch);
    const int bb = b < 0 ? 0 : b;
    if (kCost && NA == 2 && q.batch == kSpecialBatch0)          // literal batch size, see kSpecialBatch0
        dev_rollout<M, R, PB, NA, kInit, kScheme, kCost, kSpecialBatch0>(q, ws, bb, ai, b >= 0, rollout_smem);
    else if (kCost && NA == 2 && q.batch == kSpecialBatch1)
        dev_rollout<M, R, PB, NA, kInit, kScheme, kCost, kSpecialBatch1>(q, ws, bb, ai, b >= 0, rollout_smem);
    else if (kCost && NA == 2 && q.batch == kSpecialBatch2)
        dev_rollout<M, R, PB, NA, kInit, kScheme, kCost, kSpecialBatch2>(q, ws, bb, ai, b >= 0, rollout_smem);
    else if (kCost && NA == 2 && q.batch == kSpecialBatch3)
        dev_rollout<M, R, PB, NA, kInit, kScheme, kCost, kSpecialBatch3>(q, ws, bb, ai, b >= 0, rollout_smem);
    else if (kCost && NA == 2 && q.batch == kSpecialBatch4)
        dev_rollout<M, R, PB, NA, kInit, kScheme, kCost, kSpecialBatch4>(q, ws, bb, ai, b >= 0, rollout_smem);
    else
        dev_rollout<M, R, PB, NA, kInit, kScheme, kCost>(q, ws, bb, ai, b >= 0, rollout_smem);
}

// ---------------------------------------------------------------------------------
// cost terms of every (candidate, stage): stage cost for t < T, end cost for t == T.
// grid (ceil(B/128), T+1, candidates of this launch).  Candidate a of problem b is read from
// xs + a*x_stride, us + a*u_stride (the initial rollout passes q.x / q.u, 1 candidate).
// `list` != NULL: work items are entries of the pending list (round 2 of the line search).
// ---------------------------------------------------------------------------------
// S: element type of xs / us (double for q.x / q.u, scratch_t<R> for the candidates).
template <typename M, typename R, typename S>
__device__ __forceinline__ void dev_stage_cost(const tplb_batch& q, const Workspace& ws,
                                               const S* xs, const S* us, size_t x_stride,
                                               size_t u_stride, int check_running, int b, int t, int a) {
    using D = Dims<M>;
    constexpr int X = D::X, U = D::U, C = D::C;
    const int B = q.batch;
    if (check_running && !ws.running[b]) return;
    const int T = horizon_of(q, b);
    if (t > T) return;
    const int scene = __ldg(q.scene_index + b);
    const ParamView<R> P = param_view_scene<R>(q, scene);
    const S* xa = xs + a * x_stride + b;
    const S* ua = us + a * u_stride + b;

    R x[X], sc[D::NSCs], c;
#pragma unroll
    for (int i = 0; i < X; ++i) x[i] = R(xa[((size_t)t * X + i) * B]);
    load_stage_consts<M, R>(q, ws, scene, t, sc);
    if (t < T) {
        R u[U], lam[D::Cs], w[D::Cs];
#pragma unroll
        for (int i = 0; i < U; ++i) u[i] = R(ua[((size_t)t * U + i) * B]);
#pragma unroll
        for (int cc = 0; cc < C; ++cc) {
            lam[cc] = R(q.lagrange_multiplier[((size_t)t * C + cc) * B + b]);
            w[cc] = R(q.barrier_weight[(size_t)cc * B + b]);
        }
        M::stage_cost(P, x, u, lam, w, sc, R(t), R(q.dt), &c);
    } else {
        M::end_cost(P, x, sc, R(T), R(q.dt), &c);
    }
    ws.cost_terms[((size_t)a * (q.t_max + 1) + t) * B + b] = c;
}

// One thread = one (problem, stage, candidate).  With a pending list the x-blocks stride over
// the list, so the grid can stay small when few problems are pending.
template <typename M, typename R, typename S>
__global__ void stage_cost_kernel(const __grid_constant__ tplb_batch q, Workspace ws,
                                  const S* xs, const S* us, size_t x_stride, size_t u_stride,
                                  int check_running, int a_begin, const int32_t* list) {
    const int n = list ? *ws.pending_count : q.batch;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int b = list ? list[i] : i;
        dev_stage_cost<M, R, S>(q, ws, xs, us, x_stride, u_stride, check_running, b, blockIdx.y, a_begin + blockIdx.z);
    }
}

// Round 1 of the line search: one thread evaluates BOTH first-round candidates of its
// (problem, stage), so the multipliers, weights and stage constants are read once and the
// two evaluations interleave.  grid (ceil(B/128), T+1).
template <typename M, typename R>
__global__ void stage_cost_round1_kernel(const __grid_constant__ tplb_batch q, Workspace ws) {
    using D = Dims<M>;
    constexpr int X = D::X, U = D::U, C = D::C;
    const int B = q.batch, t = blockIdx.y;
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B || !ws.running[b]) return;
    const int T = horizon_of(q, b);
    if (t > T) return;
    const size_t x_stride = (size_t)(q.t_max + 1) * X * B, u_stride = (size_t)q.t_max * U * B;
    const scratch_t<R>* cand_x = scratch<scratch_t<R>>(ws.cand_x);
    const scratch_t<R>* cand_u = scratch<scratch_t<R>>(ws.cand_u);
    const int scene = __ldg(q.scene_index + b);
    const ParamView<R> P = param_view_scene<R>(q, scene);
    R sc[D::NSCs], x[kRound1][X], c[kRound1];
    load_stage_consts<M, R>(q, ws, scene, t, sc);
#pragma unroll
    for (int a = 0; a < kRound1; ++a)
#pragma unroll
        for (int i = 0; i < X; ++i) x[a][i] = R(cand_x[a * x_stride + ((size_t)t * X + i) * B + b]);
    if (t < T) {
        R u[kRound1][U], lam[D::Cs], w[D::Cs];
#pragma unroll
        for (int a = 0; a < kRound1; ++a)
#pragma unroll
            for (int i = 0; i < U; ++i) u[a][i] = R(cand_u[a * u_stride + ((size_t)t * U + i) * B + b]);
#pragma unroll
        for (int cc = 0; cc < C; ++cc) {
            lam[cc] = R(q.lagrange_multiplier[((size_t)t * C + cc) * B + b]);
            w[cc] = R(q.barrier_weight[(size_t)cc * B + b]);
        }
#pragma unroll
        for (int a = 0; a < kRound1; ++a) M::stage_cost(P, x[a], u[a], lam, w, sc, R(t), R(q.dt), &c[a]);
    } else {
#pragma unroll
        for (int a = 0; a < kRound1; ++a) M::end_cost(P, x[a], sc, R(T), R(q.dt), &c[a]);
    }
#pragma unroll
    for (int a = 0; a < kRound1; ++a) ws.cost_terms[((size_t)a * (q.t_max + 1) + t) * B + b] = c[a];
}

// trajCosts of the initial rollout: sum in the reference's order (optim.c:1099-1111)
__device__ __forceinline__ void dev_init_cost(const tplb_batch& q, const Workspace& ws, int b) {
    const int B = q.batch;
    double total = 0.0;
    const int T = horizon_of(q, b);
    for (int t = 0; t <= T; ++t) total += ws.cost_terms[(size_t)t * B + b];
    q.traj_costs[b] = total;
}

__global__ void init_cost_kernel(const __grid_constant__ tplb_batch q, Workspace ws) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= q.batch) return;
    dev_init_cost(q, ws, b);
}

// ---------------------------------------------------------------------------------
// multiplier update lambda <- min(limit, max(0, lambda + w g)) for every stage, and
// the per-outer-iteration reset of the solver flags (row t == 0 does it).
// ---------------------------------------------------------------------------------
template <typename M, typename R>
__device__ __forceinline__ void dev_multiplier(const tplb_batch& q, const Workspace& ws, int b, int t) {
    using D = Dims<M>;
    constexpr int X = D::X, U = D::U, C = D::C;
    const int B = q.batch;
    if (t == 0) {
        q.trajectory_changed[b] = 1;
        q.improved[b] = 0;
        q.iterations[b] = 0;
        ws.running[b] = 1;
        ws.winner[b] = -1;                      // nothing to install in front of the first sweep
        bool zero = C > 0;
#pragma unroll
        for (int c = 0; c < C; ++c) zero = zero && q.lg_mult_limit[(size_t)c * B + b] == 0.0;
        ws.lam_zero[b] = zero;
    }
    if (C == 0 || t >= horizon_of(q, b)) return;
    const int scene = __ldg(q.scene_index + b);
    const ParamView<R> P = param_view_scene<R>(q, scene);
    R x[X], u[U], lam[D::Cs], w[D::Cs], g[D::Cs], sc[D::NSCs];
#pragma unroll
    for (int i = 0; i < X; ++i) x[i] = R(q.x[((size_t)t * X + i) * B + b]);
#pragma unroll
    for (int i = 0; i < U; ++i) u[i] = R(q.u[((size_t)t * U + i) * B + b]);
#pragma unroll
    for (int c = 0; c < C; ++c) {
        lam[c] = R(q.lagrange_multiplier[((size_t)t * C + c) * B + b]);
        w[c] = R(q.barrier_weight[(size_t)c * B + b]);
        g[c] = R(0);
    }
    load_stage_consts<M, R>(q, ws, scene, t, sc);
    M::constraints(P, x, u, lam, w, sc, R(t), R(q.dt), g);
#pragma unroll
    for (int c = 0; c < C; ++c) {
        R v = lam[c] + w[c] * g[c];
        v = (R(0) > v) ? R(0) : v;
        const R lim = R(q.lg_mult_limit[(size_t)c * B + b]);
        q.lagrange_multiplier[((size_t)t * C + c) * B + b] = (lim < v) ? lim : v;
    }
}

template <typename M, typename R>
__global__ void multiplier_kernel(const __grid_constant__ tplb_batch q, Workspace ws) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= q.batch) return;
    dev_multiplier<M, R>(q, ws, b, blockIdx.y);
}

// ---------------------------------------------------------------------------------
// linearisation / quadratisation of every stage — thread per (problem, stage).
// Only entries that are not identically 0/1 are stored (compact record).
// kForce: every problem (tplb_linearize); otherwise only running problems whose
// trajectory changed (optim.c:896).  kAccept: first install the step the previous
// line search accepted (the work of accept_kernel; grid has one extra row for x[T]).
// ---------------------------------------------------------------------------------
template <typename M, typename R, bool kForce, bool kAccept>
__device__ __forceinline__ void dev_linearize(const tplb_batch& q, const Workspace& ws, int b, int t) {
    using D = Dims<M>;
    constexpr int X = D::X, U = D::U, C = D::C;
    const int B = q.batch;
    const int T = horizon_of(q, b);
    if (t > T) return;

    R x[X], u[U];
    bool have = false;
    if (kAccept) {
        // the step accepted by the previous line search becomes the trajectory (optim.c:844-848)
        const int win = ws.winner[b];
        if (win >= 0) {
            const double* cx = ws.cand_x + (size_t)win * (q.t_max + 1) * X * B + b;
            const double* cu = ws.cand_u + (size_t)win * q.t_max * U * B + b;
#pragma unroll
            for (int i = 0; i < X; ++i) {
                const size_t idx = ((size_t)t * X + i) * B + b;
                const double xv = (double)cx[((size_t)t * X + i) * B];
                x[i] = R(xv);
                if (q.keep_previous) q.prev_x[idx] = q.x[idx];
                q.x[idx] = xv;
            }
            if (t < T) {
#pragma unroll
                for (int d = 0; d < U; ++d) {
                    const size_t idx = ((size_t)t * U + d) * B + b;
                    const double uv = (double)cu[((size_t)t * U + d) * B];
                    u[d] = R(uv);
                    if (q.keep_previous) q.prev_k[idx] = q.k[idx];
                    q.u[idx] = uv;
                }
            }
            have = true;
        }
    }
    if (t >= T) return;
    if (!kForce && !(ws.running[b] && q.trajectory_changed[b])) return;
    const int scene = __ldg(q.scene_index + b);
    const ParamView<R> P = param_view_scene<R>(q, scene);

    R lam[D::Cs], w[D::Cs], sc[D::NSCs];
    if (!have) {
#pragma unroll
        for (int i = 0; i < X; ++i) x[i] = R(q.x[((size_t)t * X + i) * B + b]);
#pragma unroll
        for (int i = 0; i < U; ++i) u[i] = R(q.u[((size_t)t * U + i) * B + b]);
    }
    // inside update() the multipliers of a pure-penalty problem are known to be 0 (ws.lam_zero)
    const bool lam_is_zero = !kForce && C > 0 && ws.lam_zero[b] != 0;
#pragma unroll
    for (int c = 0; c < C; ++c) {
        lam[c] = lam_is_zero ? R(0) : R(q.lagrange_multiplier[((size_t)t * C + c) * B + b]);
        w[c] = R(q.barrier_weight[(size_t)c * B + b]);
    }
    load_stage_consts<M, R>(q, ws, scene, t, sc);
    R blk[D::DENSE];
    if (q.use_quadratic_terms) {
        M::linearize(P, x, u, lam, w, sc, R(t), R(q.dt),
                     blk + D::OFF_FX, blk + D::OFF_FU, blk + D::OFF_LX, blk + D::OFF_LU,
                     blk + D::OFF_LXX, blk + D::OFF_LUU, blk + D::OFF_LUX);
    } else {
#pragma unroll
        for (int e = D::OFF_LXX; e < D::DENSE; ++e) blk[e] = R(0);
        M::dynamics_jacobians(P, x, u, sc, R(t), R(q.dt), blk + D::OFF_FX, blk + D::OFF_FU);
        M::cost_gradients(P, x, u, lam, w, sc, R(t), R(q.dt), blk + D::OFF_LX, blk + D::OFF_LU);
    }
    scratch_t<R>* out = scratch<scratch_t<R>>(ws.deriv) + (size_t)t * D::COMPACT * B + b;
#pragma unroll
    for (int e = 0; e < D::DENSE; ++e)
        if (M::deriv_owner(e)) out[(size_t)M::deriv_slot(e) * B] = blk[e];
    if (b == 0 && t == 0) *ws.records_f32 = sizeof(scratch_t<R>) == sizeof(float);
}

template <typename M, typename R, bool kForce, bool kAccept>
__global__ void __launch_bounds__(128, 4)
linearize_kernel(const __grid_constant__ tplb_batch q, Workspace ws) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= q.batch) return;
    dev_linearize<M, R, kForce, kAccept>(q, ws, b, blockIdx.y);   // rows 0..T-1 (0..T with kAccept)
}

// compact -> dense records for the fx..lux views (optim.c:1663-1669), thread per (problem, stage)
template <typename M>
__global__ void expand_derivatives_kernel(const __grid_constant__ tplb_batch q, Workspace ws, double* dense) {
    using D = Dims<M>;
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    const int t = blockIdx.y;
    const int B = q.batch;
    if (b >= B) return;
    const bool f32 = *ws.records_f32 != 0;                   // format the last linearisation wrote
    const double* in = ws.deriv + (size_t)t * D::COMPACT * B + b;
    const float* in32 = scratch<float>(ws.deriv) + (size_t)t * D::COMPACT * B + b;
    double* out = dense + (size_t)t * D::DENSE * B + b;
#pragma unroll
    for (int e = 0; e < D::DENSE; ++e) {
        const int s = M::deriv_slot(e);
        out[(size_t)e * B] = s >= 0 ? (f32 ? (double)in32[(size_t)s * B] : in[(size_t)s * B]) : (s == -2 ? 1.0 : 0.0);
    }
}

// ---------------------------------------------------------------------------------
// gains  k = -(Quu + mu I)^-1 Qu,  K = -(Quu + mu I)^-1 Qux   (optim.c:243-291)
// ---------------------------------------------------------------------------------
// -(Quu + mu I)^-1 for one or two controls (optim.c:243-291): U == 1 tests the un-regularised
// value and returns 0 gains otherwise; U == 2 is the closed-form inverse without a definiteness check
template <int U, typename R>
__device__ __forceinline__ void gain_inverse(const R (&Quu)[U][U], R mu, R (&Mi)[U][U]) {
    static_assert(U == 1 || U == 2, "more than two controls are not supported (genopt.py:420-425)");
    // -1 / x as the negated straight-line reciprocal (correctly rounded on every argument tried, so the
    // value of the reference's division; no slow-path call in the middle of the Riccati step)
    if constexpr (U == 1) {
        const R s = (Quu[0][0] > R(0)) ? -m_inv(Quu[0][0] + mu) : R(0);     // test on the un-regularised value
        Mi[0][0] = s;
    } else {
        const R a = Quu[0][0] + mu, bb = Quu[0][1], d = Quu[1][1] + mu;
        const R det = a * d - bb * bb;
        const R s = -m_inv(det);                          // no definiteness check
        Mi[0][0] = d * s;
        Mi[0][1] = -bb * s;
        Mi[1][0] = Mi[0][1];
        Mi[1][1] = a * s;
    }
}

// one row of  Mi * v  as the reference's matrix product forms it
template <int U, typename R>
__device__ __forceinline__ R gain_row(const R (&Mi)[U][U], int i, const R (&v)[U]) {
    if constexpr (U == 1) {
        return v[0] * Mi[0][0];
    } else {
        R acc = R(0);
#pragma unroll
        for (int c = 0; c < 2; ++c) acc += Mi[i][c] * v[c];
        return acc;
    }
}

template <int X, int U, typename R>
__device__ __forceinline__ void control_gains(const R (&Quu)[U][U], const R (&Qu)[U],
                                              const R (&Qux)[U][X], R mu,
                                              R (&k)[U], R (&K)[U][X]) {
    R Mi[U][U];
    gain_inverse<U, R>(Quu, mu, Mi);
#pragma unroll
    for (int i = 0; i < U; ++i) {
        k[i] = gain_row<U, R>(Mi, i, Qu);
#pragma unroll
        for (int j = 0; j < X; ++j) {
            R col[U];
#pragma unroll
            for (int c = 0; c < U; ++c) col[c] = Qux[c][j];
            K[i][j] = gain_row<U, R>(Mi, i, col);
        }
    }
}

// acc += a * v where `slot` is the compile-time structure of a: -1 -> a == 0, -2 -> a == 1
template <typename R>
__device__ __forceinline__ void madd(R& acc, int slot, R a, R v) {
    if (slot == -1) return;
    if (slot == -2) acc += v;
    else acc += a * v;
}

// ---------------------------------------------------------------------------------
// one stage of the Riccati recursion (optim.c:926-985): Q terms from the compact record `rec`
// of the stage and the value function of the next one, gains, box limits on the feed-forward
// step, value function of this stage.  Everything in registers; shared by the stand-alone
// backward sweep and the fused linearise+sweep, so both round identically.
// ---------------------------------------------------------------------------------
template <typename M, typename R>
__device__ __forceinline__ void riccati_stage(const R (&rec)[Dims<M>::COMPACT], R (&Vx)[M::X], R (&Vxx)[M::X][M::X],
                                              R mu, const R (&ub)[M::U], const R (&hib)[M::U],
                                              const R (&lob)[M::U], R (&k)[M::U], R (&K)[M::U][M::X]) {
    using D = Dims<M>;
    constexpr int X = D::X, U = D::U;
    // dense entry e of the record: stored value, or the constant the structure says
    auto val = [&](int e) {
        const int s = M::deriv_slot(e);
        return s >= 0 ? rec[s] : (s == -2 ? R(1) : R(0));
    };
#define A_(i, j) val(D::OFF_FX + (i) * X + (j))
#define SA_(i, j) M::deriv_slot(D::OFF_FX + (i) * X + (j))
#define B_(i, j) val(D::OFF_FU + (i) * U + (j))
#define SB_(i, j) M::deriv_slot(D::OFF_FU + (i) * U + (j))

    R Qx[X], Qu[U], Qxx[X][X], Quu[U][U], Qux[U][X];
    R VA[X][X], VB[X][U];

#pragma unroll
    for (int i = 0; i < X; ++i) {                        // Qx = lx + A' Vx
        R acc = R(0);
#pragma unroll
        for (int r = 0; r < X; ++r) madd(acc, SA_(r, i), A_(r, i), Vx[r]);
        Qx[i] = val(D::OFF_LX + i) + acc;
    }
#pragma unroll
    for (int i = 0; i < U; ++i) {                        // Qu = lu + B' Vx
        R acc = R(0);
#pragma unroll
        for (int r = 0; r < X; ++r) madd(acc, SB_(r, i), B_(r, i), Vx[r]);
        Qu[i] = val(D::OFF_LU + i) + acc;
    }
#pragma unroll
    for (int i = 0; i < X; ++i) {                        // VA = Vxx A, VB = Vxx B
#pragma unroll
        for (int j = 0; j < X; ++j) {
            R acc = R(0);
#pragma unroll
            for (int r = 0; r < X; ++r) madd(acc, SA_(r, j), A_(r, j), Vxx[i][r]);
            VA[i][j] = acc;
        }
#pragma unroll
        for (int j = 0; j < U; ++j) {
            R acc = R(0);
#pragma unroll
            for (int r = 0; r < X; ++r) madd(acc, SB_(r, j), B_(r, j), Vxx[i][r]);
            VB[i][j] = acc;
        }
    }
#pragma unroll
    for (int i = 0; i < X; ++i)                          // Qxx = lxx + A' VA (lower triangle mirrored)
#pragma unroll
        for (int j = 0; j <= i; ++j) {
            R acc = R(0);
#pragma unroll
            for (int r = 0; r < X; ++r) madd(acc, SA_(r, i), A_(r, i), VA[r][j]);
            Qxx[i][j] = acc;
            Qxx[j][i] = acc;
        }
#pragma unroll
    for (int i = 0; i < X; ++i)
#pragma unroll
        for (int j = 0; j < X; ++j) Qxx[i][j] = val(D::OFF_LXX + i * X + j) + Qxx[i][j];
#pragma unroll
    for (int i = 0; i < U; ++i)                          // Quu = luu + B' VB (lower triangle mirrored)
#pragma unroll
        for (int j = 0; j <= i; ++j) {
            R acc = R(0);
#pragma unroll
            for (int r = 0; r < X; ++r) madd(acc, SB_(r, i), B_(r, i), VB[r][j]);
            Quu[i][j] = acc;
            Quu[j][i] = acc;
        }
#pragma unroll
    for (int i = 0; i < U; ++i)
#pragma unroll
        for (int j = 0; j < U; ++j) Quu[i][j] = val(D::OFF_LUU + i * U + j) + Quu[i][j];
#pragma unroll
    for (int i = 0; i < U; ++i)                          // Qux = lux + B' VA
#pragma unroll
        for (int j = 0; j < X; ++j) {
            R acc = R(0);
#pragma unroll
            for (int r = 0; r < X; ++r) madd(acc, SB_(r, i), B_(r, i), VA[r][j]);
            Qux[i][j] = val(D::OFF_LUX + i * X + j) + acc;
        }
#undef A_
#undef SA_
#undef B_
#undef SB_

    control_gains<X, U, R>(Quu, Qu, Qux, mu, k, K);

    // box limits on the feed-forward step (optim.c:950-963): two independent tests on u + k, the
    // lower one wins if both fire; selects instead of branches keep the step one basic block
#pragma unroll
    for (int d = 0; d < U; ++d) {
        const R cand = ub[d] + k[d];
        const bool above = cand > hib[d], below = cand < lob[d];
        k[d] = above ? hib[d] - ub[d] : k[d];
        k[d] = below ? lob[d] - ub[d] : k[d];
#pragma unroll
        for (int j = 0; j < X; ++j) K[d][j] = (above || below) ? R(0) : K[d][j];
    }

    // value function (optim.c:965-984)
    R KtQux[X][X], KtQuu[X][U];
#pragma unroll
    for (int i = 0; i < X; ++i) {
#pragma unroll
        for (int j = 0; j < X; ++j) {
            R acc = R(0);
#pragma unroll
            for (int c = 0; c < U; ++c) acc += K[c][i] * Qux[c][j];
            KtQux[i][j] = acc;
        }
#pragma unroll
        for (int j = 0; j < U; ++j) {
            R acc = R(0);
#pragma unroll
            for (int c = 0; c < U; ++c) acc += K[c][i] * Quu[c][j];
            KtQuu[i][j] = acc;
        }
    }
#pragma unroll
    for (int i = 0; i < X; ++i)
#pragma unroll
        for (int j = 0; j < X; ++j) {
            R v = KtQux[j][i] + KtQux[i][j];
#pragma unroll
            for (int c = 0; c < U; ++c) v += KtQuu[i][c] * K[c][j];
            Vxx[i][j] = v + Qxx[i][j];
        }
#pragma unroll
    for (int i = 0; i < X; ++i) {
        R v = R(0);
#pragma unroll
        for (int c = 0; c < U; ++c) v += KtQuu[i][c] * k[c];
#pragma unroll
        for (int c = 0; c < U; ++c) v += K[c][i] * Qu[c];
#pragma unroll
        for (int c = 0; c < U; ++c) v += Qux[c][i] * k[c];
        Vx[i] = v + Qx[i];
    }
}

// ---------------------------------------------------------------------------------
// backward Riccati sweep — thread per problem, value function in registers.
// The next stage's compact record is prefetched while the current one is processed.
// ---------------------------------------------------------------------------------
template <typename M, typename R>
__device__ __forceinline__ void dev_backward(const tplb_batch& q, const Workspace& ws, int b, int iteration) {
    using D = Dims<M>;
    constexpr int X = D::X, U = D::U, NC = D::COMPACT;
    const int B = q.batch;
    if (!ws.running[b]) return;
    q.iterations[b] = iteration + 1;             // optim.c:894
    ws.counters[b] += q.trajectory_changed[b] ? 1 : 0;
    ws.counters[(size_t)B + b] += 1;
    q.trajectory_changed[b] = 0;                 // optim.c:911 (linearize_kernel ran just before)
    const int scene = __ldg(q.scene_index + b);
    const ParamView<R> P = param_view_scene<R>(q, scene);
    const int T = horizon_of(q, b);
    const double mu = q.mu[b];

    R Vx[X], Vxx[X][X];
    {
        R xT[X], sc[D::NSCs];
#pragma unroll
        for (int i = 0; i < X; ++i) xT[i] = R(q.x[((size_t)T * X + i) * B + b]);
        load_stage_consts<M, R>(q, ws, scene, T, sc);
        M::end_derivatives(P, xT, sc, R(T), R(q.dt), Vx, &Vxx[0][0]);
    }

    R rec[NC], nxt[NC], ub[U], hib[U], lob[U], nub[U], nhib[U], nlob[U];
    auto fetch = [&](int t, R* r, R* uu, R* hh, R* ll) {
        const scratch_t<R>* blk = scratch<scratch_t<R>>(ws.deriv) + (size_t)t * NC * B + b;
#pragma unroll
        for (int s = 0; s < M::DERIV_COMPACT; ++s) r[s] = R(__ldcs(blk + (size_t)s * B));   // read once
#pragma unroll
        for (int d = 0; d < U; ++d) {
            const size_t idx = ((size_t)t * U + d) * B + b;
            uu[d] = R(q.u[idx]);
            hh[d] = R(q.u_max[idx]);
            ll[d] = R(q.u_min[idx]);
        }
    };
    fetch(T - 1, nxt, nub, nhib, nlob);

    for (int t = T - 1; t >= 0; --t) {
#pragma unroll
        for (int s = 0; s < NC; ++s) rec[s] = nxt[s];
#pragma unroll
        for (int d = 0; d < U; ++d) { ub[d] = nub[d]; hib[d] = nhib[d]; lob[d] = nlob[d]; }
        if (t > 0) fetch(t - 1, nxt, nub, nhib, nlob);

        R k[U], K[U][X];
        riccati_stage<M, R>(rec, Vx, Vxx, R(mu), ub, hib, lob, k, K);
#pragma unroll
        for (int d = 0; d < U; ++d) {
            q.k[((size_t)t * U + d) * B + b] = k[d];
#pragma unroll
            for (int j = 0; j < X; ++j) q.K[((size_t)t * U * X + d * X + j) * B + b] = K[d][j];
        }
    }
}

template <typename M, typename R>
__global__ void __launch_bounds__(128)
backward_kernel(const __grid_constant__ tplb_batch q, Workspace ws, int iteration) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b == 0) *ws.pending_count = 0;           // the line search of this iteration starts empty
    if (b >= q.batch) return;
    dev_backward<M, R>(q, ws, b, iteration);
}

// ---------------------------------------------------------------------------------
// fused linearise + Riccati sweep — the throughput sequence's replacement for
// accept_kernel + linearize_kernel + backward_kernel.  Thread per problem, t = T-1 .. 0:
//   * the trajectory is read straight from the candidate the previous line search accepted
//     (ws.winner) and installed into x, u on the way (optim.c:844-848) — no separate copy pass
//     (q.keep_previous: keep_previous_kernel saves x, k of the winners in front of the sweep);
//   * the derivative record of stage t is evaluated in registers and consumed by the Riccati
//     step of the same stage (optim.c:896-985) — it never travels through HBM
//     (q.keep_records != 0: it is also stored for the fx..lux views);
//   * gains K, k are the only per-stage output.
// Per (problem, stage): x, u (8) + box limits (4) [+ multipliers] in, x, u (8) + K, k (14) out,
// instead of 32 (accept) + 37 (linearize) + 49 (backward) doubles.
// In the first iteration of an inner loop ws.winner is -1 (multiplier_kernel): nothing to install.
// A problem that stopped in the previous iteration only installs its last accepted step.
// Linearising again after a failed line search (trajectory_changed == 0, optim.c:896 skips
// it) reproduces the stored record bit for bit: same inputs, same code.
// ---------------------------------------------------------------------------------
template <typename M, typename R, typename PV>
__device__ __forceinline__ void stage_record(const PV& P, const R* x, const R* u, const R* lam, const R* w,
                                             const R* sc, R t, R dt, R (&rec)[Dims<M>::COMPACT]) {
    using D = Dims<M>;
    R blk[D::DENSE];
    M::linearize(P, x, u, lam, w, sc, t, dt, blk + D::OFF_FX, blk + D::OFF_FU, blk + D::OFF_LX, blk + D::OFF_LU,
                 blk + D::OFF_LXX, blk + D::OFF_LUU, blk + D::OFF_LUX);
#pragma unroll
    for (int e = 0; e < D::DENSE; ++e)
        if (M::deriv_owner(e)) rec[M::deriv_slot(e)] = blk[e];
}

template <typename M, typename R, int KB = 0>
__device__ __forceinline__ void dev_sweep(const tplb_batch& q, const Workspace& ws, int b, int iteration) {
    if constexpr (KB > 0) __builtin_assume(q.batch == KB);
    using D = Dims<M>;
    using S = double;                            // line-search candidates are fp64 in every mode
    using SR = scratch_t<R>;                     // derivative records: storage type of the compute precision
    constexpr int X = D::X, U = D::U, C = D::C, NC = D::COMPACT, NSC = D::NSC;
    const int B = q.batch, T = horizon_of(q, b);
    const int win = ws.winner[b];
    const bool run = ws.running[b] != 0;
    if (!run && win < 0) return;
    const bool install = win >= 0;
    const S* cx = scratch<S>(ws.cand_x) + (size_t)(install ? win : 0) * (q.t_max + 1) * X * B + b;
    const S* cu = scratch<S>(ws.cand_u) + (size_t)(install ? win : 0) * q.t_max * U * B + b;
    double* qx = q.x + b;
    double* qu = q.u + b;

    // component i of stage t of the trajectory this iteration linearises about
    auto load_x = [&](int t, int i) -> R {
        const size_t idx = ((size_t)t * X + i) * B;
        if constexpr (std::is_same<S, double>::value) return R((install ? cx : qx)[idx]);
        else return install ? R(cx[idx]) : R(qx[idx]);
    };
    auto load_u = [&](int t, int d) -> R {
        const size_t idx = ((size_t)t * U + d) * B;
        if constexpr (std::is_same<S, double>::value) return R((install ? cu : qu)[idx]);
        else return install ? R(cu[idx]) : R(qu[idx]);
    };
    // the accepted step becomes the trajectory (optim.c:844-848)
    auto install_x = [&](int t, const R* x) {
        if (!install) return;
#pragma unroll
        for (int i = 0; i < X; ++i) {
            const size_t idx = ((size_t)t * X + i) * B;
            qx[idx] = (double)x[i];
        }
    };
    auto install_u = [&](int t, const R* u) {
        if (!install) return;
#pragma unroll
        for (int d = 0; d < U; ++d) {
            const size_t idx = ((size_t)t * U + d) * B;
            qu[idx] = (double)u[d];
        }
    };

    if (!run) {                                  // stopped in the previous iteration: install only
        for (int t = 0; t <= T; ++t) {
            R x[X], u[U];
#pragma unroll
            for (int i = 0; i < X; ++i) x[i] = load_x(t, i);
            install_x(t, x);
            if (t < T) {
#pragma unroll
                for (int d = 0; d < U; ++d) u[d] = load_u(t, d);
                install_u(t, u);
            }
        }
        return;
    }

    q.iterations[b] = iteration + 1;             // optim.c:894
    ws.counters[b] += q.trajectory_changed[b] ? 1 : 0;
    ws.counters[(size_t)B + b] += 1;
    q.trajectory_changed[b] = 0;                 // optim.c:911
    const int scene = __ldg(q.scene_index + b);
    const ParamView<R> P = param_view_scene<R>(q, scene);
    const R mu = R(q.mu[b]);
    const R dt = R(q.dt);
    const bool lam_is_zero = C > 0 && ws.lam_zero[b] != 0;
    R w[D::Cs];
#pragma unroll
    for (int c = 0; c < C; ++c) w[c] = R(q.barrier_weight[(size_t)c * B + b]);

    R Vx[X], Vxx[X][X];
    {
        R xT[X], sc[D::NSCs];
#pragma unroll
        for (int i = 0; i < X; ++i) xT[i] = load_x(T, i);
        install_x(T, xT);
        load_stage_consts<M, R>(q, ws, scene, T, sc);
        M::end_derivatives(P, xT, sc, R(T), dt, Vx, &Vxx[0][0]);
    }

    // Software pipeline: the inputs of stage t-1 (trajectory, box limits, multipliers, stage constants)
    // are requested right after the derivative record of stage t has consumed the registers that
    // hold stage t's, i.e. in front of the long Riccati step that does not need them — their HBM /
    // L2 latency hides behind it and no prefetch register is live during the linearisation.
    R x[X], u[U], hi[U], lo[U], lam[D::Cs], sc[D::NSCs];
    auto fetch = [&](int t) {
#pragma unroll
        for (int i = 0; i < X; ++i) x[i] = load_x(t, i);
#pragma unroll
        for (int d = 0; d < U; ++d) {
            const size_t idx = ((size_t)t * U + d) * B + b;
            u[d] = load_u(t, d);
            hi[d] = R(q.u_max[idx]);
            lo[d] = R(q.u_min[idx]);
        }
#pragma unroll
        for (int c = 0; c < C; ++c)
            lam[c] = lam_is_zero ? R(0) : R(q.lagrange_multiplier[((size_t)t * C + c) * B + b]);
        load_stage_consts<M, R>(q, ws, scene, t, sc);
    };
    fetch(T - 1);

    for (int t = T - 1; t >= 0; --t) {
        install_x(t, x);
        install_u(t, u);

        R rec[NC];
        stage_record<M, R>(P, x, u, lam, w, sc, R(t), dt, rec);
        if (q.keep_records) {
            SR* out = scratch<SR>(ws.deriv) + (size_t)t * NC * B + b;
#pragma unroll
            for (int s = 0; s < M::DERIV_COMPACT; ++s) __stcs(out + (size_t)s * B, (SR)rec[s]);
        }
        R ub[U], hib[U], lob[U];
#pragma unroll
        for (int d = 0; d < U; ++d) { ub[d] = u[d]; hib[d] = hi[d]; lob[d] = lo[d]; }
        if (t > 0) fetch(t - 1);

        R k[U], K[U][X];
        riccati_stage<M, R>(rec, Vx, Vxx, mu, ub, hib, lob, k, K);
#pragma unroll
        for (int d = 0; d < U; ++d) {
            q.k[((size_t)t * U + d) * B + b] = k[d];
#pragma unroll
            for (int j = 0; j < X; ++j) q.K[((size_t)t * U * X + d * X + j) * B + b] = K[d][j];
        }
    }
    if (q.keep_records && b == 0) *ws.records_f32 = sizeof(SR) == sizeof(float);
}

// prev_x <- x, prev_k <- k for the problems whose last line search accepted a step (optim.c:844-845),
// in front of a sweep that is about to install that step and overwrite the gains.  Stage parallel;
// launched only when the caller wants prev_x / prev_k maintained (q.keep_previous).
template <typename M>
__global__ void keep_previous_kernel(const __grid_constant__ tplb_batch q, Workspace ws) {
    using D = Dims<M>;
    constexpr int X = D::X, U = D::U;
    const int b = blockIdx.x * blockDim.x + threadIdx.x, t = blockIdx.y, B = q.batch;
    if (b >= B || ws.winner[b] < 0) return;
    const int T = horizon_of(q, b);
    if (t > T) return;
#pragma unroll
    for (int i = 0; i < X; ++i) {
        const size_t idx = ((size_t)t * X + i) * B + b;
        q.prev_x[idx] = q.x[idx];
    }
    if (t < T) {
#pragma unroll
        for (int d = 0; d < U; ++d) {
            const size_t idx = ((size_t)t * U + d) * B + b;
            q.prev_k[idx] = q.k[idx];
        }
    }
}

template <typename M, typename R>
__global__ void __launch_bounds__(128)
sweep_kernel(const __grid_constant__ tplb_batch q, Workspace ws, int iteration) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b == 0) *ws.pending_count = 0;           // the line search of this iteration starts empty
    if (b >= q.batch) return;
    if (q.batch == kSpecialBatch0) dev_sweep<M, R, kSpecialBatch0>(q, ws, b, iteration);   // literal batch size
    else if (q.batch == kSpecialBatch1) dev_sweep<M, R, kSpecialBatch1>(q, ws, b, iteration);
    else if (q.batch == kSpecialBatch2) dev_sweep<M, R, kSpecialBatch2>(q, ws, b, iteration);
    else if (q.batch == kSpecialBatch3) dev_sweep<M, R, kSpecialBatch3>(q, ws, b, iteration);
    else if (q.batch == kSpecialBatch4) dev_sweep<M, R, kSpecialBatch4>(q, ws, b, iteration);
    else dev_sweep<M, R>(q, ws, b, iteration);
}

// gradient-only sweep (optim.c:1038-1076): costate recursion, clipped descent direction
template <typename M, typename R>
__device__ __forceinline__ void dev_backward_first_order(const tplb_batch& q, const Workspace& ws, int b,
                                                         int iteration) {
    using D = Dims<M>;
    constexpr int X = D::X, U = D::U;
    const int B = q.batch;
    if (!ws.running[b]) return;
    q.iterations[b] = iteration + 1;
    ws.counters[b] += q.trajectory_changed[b] ? 1 : 0;
    ws.counters[(size_t)B + b] += 1;
    q.trajectory_changed[b] = 0;
    const int scene = __ldg(q.scene_index + b);
    const ParamView<R> P = param_view_scene<R>(q, scene);
    const int T = horizon_of(q, b);
    R Vx[X];
    {
        R xT[X], Vxx[X * X], sc[D::NSCs];
#pragma unroll
        for (int i = 0; i < X; ++i) xT[i] = R(q.x[((size_t)T * X + i) * B + b]);
        load_stage_consts<M, R>(q, ws, scene, T, sc);
        M::end_derivatives(P, xT, sc, R(T), R(q.dt), Vx, Vxx);
    }
    for (int t = T - 1; t >= 0; --t) {
        const scratch_t<R>* blk = scratch<scratch_t<R>>(ws.deriv) + (size_t)t * D::COMPACT * B + b;
        auto val = [&](int e) {
            const int s = M::deriv_slot(e);
            return s >= 0 ? R(blk[(size_t)s * B]) : (s == -2 ? R(1) : R(0));
        };
        R Qx[X];
#pragma unroll
        for (int i = 0; i < X; ++i) {
            R acc = R(0);
#pragma unroll
            for (int r = 0; r < X; ++r)
                madd(acc, M::deriv_slot(D::OFF_FX + r * X + i), val(D::OFF_FX + r * X + i), Vx[r]);
            Qx[i] = val(D::OFF_LX + i) + acc;
        }
#pragma unroll
        for (int i = 0; i < U; ++i) {
            R acc = R(0);
#pragma unroll
            for (int r = 0; r < X; ++r)
                madd(acc, M::deriv_slot(D::OFF_FU + r * U + i), val(D::OFF_FU + r * U + i), Vx[r]);
            const R gq = val(D::OFF_LU + i) + acc;
            const size_t idx = ((size_t)t * U + i) * B + b;
            const R ui = R(q.u[idx]);
            R kk = gq;
            const R cand = ui - gq;
            if (cand > R(q.u_max[idx])) kk = ui - R(q.u_max[idx]);
            if (cand < R(q.u_min[idx])) kk = ui - R(q.u_min[idx]);
            if (q.g) q.g[idx] = gq;
            q.k[idx] = kk;
        }
#pragma unroll
        for (int i = 0; i < X; ++i) Vx[i] = Qx[i];
    }
}

template <typename M, typename R>
__global__ void backward_first_order_kernel(const __grid_constant__ tplb_batch q, Workspace ws, int iteration) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b == 0) *ws.pending_count = 0;
    if (b >= q.batch) return;
    dev_backward_first_order<M, R>(q, ws, b, iteration);
}

// ---------------------------------------------------------------------------------
// line-search decision in two rounds.  Costs are summed in the reference's
// order (t = 0..T-1, then the end cost; optim.c:741-789); the lowest i whose cost
// passes `testImprovement` wins, which is what the sequential early-exit loop of
// the reference selects (optim.c:861-869).  Then the regularisation schedule and the
// relative-change stop test (optim.c:987-1006).
// ---------------------------------------------------------------------------------
// everything the reference does once the search has ended for a problem (optim.c:849-851, 987-1006);
// the status words are handed in by reference: global arrays (batched kernels) or shared memory
// (one-launch kernel, solo.cuh)
struct SearchStatus {
    double& traj_costs;
    double& alpha;
    double& mu;
    int32_t& mu_step;
    int32_t& trajectory_changed;
    int32_t& improved;
    int32_t& termination_condition;
    int32_t& running;
    int32_t& winner;
    int32_t& last_tried;        // candidate the reference's next_x / next_u hold after this search
    int32_t& rollouts;          // rollouts a sequential search runs
};

__device__ __forceinline__ void conclude_core(const SearchStatus& s, bool second_order, double min_rel_cost_change,
                                              int win, double before, double now) {
    s.winner = win;
    s.last_tried = win >= 0 ? win : kAlphas - 1;
    s.rollouts += (win >= 0) ? win + 1 : kAlphas;
    {
        double tn = 1.0;                                     // alpha of the last step tried (optim.c:863)
        for (int i = 0; i < (win < 0 ? kAlphas - 1 : win); ++i) tn *= 10.0;
        s.alpha = 1.0 / tn;
    }
    if (win >= 0) {
        s.traj_costs = now;
        s.trajectory_changed = 1;
        s.improved = 1;
    }
    if (second_order) {                                      // regularisation schedule (optim.c:989-999)
        int ms = s.mu_step;
        ms = (win >= 0) ? (ms - 1 > 0 ? ms - 1 : 0) : (ms + 1 < 7 ? ms + 1 : 7);
        s.mu_step = ms;
        double m = 0.0;
        if (ms > 0) {
            m = 1.0;
            for (int i = 1; i < ms; ++i) m *= 10.0;          // 10^(ms-1), exact
        }
        s.mu = m;
    }
    const double rel = fabs(now - before) / now;             // optim.c:1001-1006
    if (rel < min_rel_cost_change) {
        s.termination_condition = 2;
        s.running = 0;
    }
}

__device__ __forceinline__ void conclude_line_search(const tplb_batch& q, const Workspace& ws, int b,
                                                     int win, double before, double now) {
    const int B = q.batch;
    const SearchStatus s{q.traj_costs[b], q.alpha[b], q.mu[b], q.mu_step[b], q.trajectory_changed[b],
                         q.improved[b], q.termination_condition[b], ws.running[b], ws.winner[b],
                         ws.last_tried[b], ws.counters[(size_t)2 * B + b]};
    conclude_core(s, q.use_quadratic_terms != 0, q.min_rel_cost_change, win, before, now);
}

// cost of candidate a of problem b: the stage terms added in the reference's order
__device__ __forceinline__ double dev_candidate_total(const tplb_batch& q, const Workspace& ws, int b, int a) {
    const int B = q.batch, T = horizon_of(q, b);
    const double* terms = ws.cost_terms + (size_t)a * (q.t_max + 1) * B + b;
    double total = 0.0;
    int t = 0;
    for (; t + 8 <= T + 1; t += 8) {                         // loads in flight together, adds in order
        double c[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) c[j] = terms[(size_t)(t + j) * B];
#pragma unroll
        for (int j = 0; j < 8; ++j) total += c[j];
    }
    for (; t <= T; ++t) total += terms[(size_t)t * B];
    ws.cand_cost[(size_t)a * B + b] = total;
    return total;
}

// lowest index in [a0, a0 + na) of s_total[.][lane] that passes testImprovement (optim.c:842)
template <int PB>
__device__ __forceinline__ int first_improving(const double (*s_total)[PB], int lane, int a0, int na,
                                               double before, double& now) {
    int win = -1;
    now = before;
    for (int i = na - 1; i >= 0; --i) {
        const double c = s_total[i][lane];
        if (c < before && isfinite(c) && c >= 0.0) {
            win = a0 + i;
            now = c;
        }
    }
    return win;
}

// kRound == 1: block = PB problems x kRound1 candidates, every problem of the batch.
// kRound == 2: block = PB pending problems x (8 - kRound1) candidates.
// kSummed: the rollouts already left the candidate totals in ws.cand_cost.
template <int PB, int kRound, bool kSummed = false>
__global__ void __launch_bounds__(PB * (kRound == 1 ? kRound1 : kAlphas - kRound1))
select_kernel(const __grid_constant__ tplb_batch q, Workspace ws) {
    constexpr int NA = kRound == 1 ? kRound1 : kAlphas - kRound1;
    constexpr int A0 = kRound == 1 ? 0 : kRound1;
    __shared__ double s_total[NA][PB];
    const int lane = threadIdx.x, a = A0 + threadIdx.y;
    const int B = q.batch;
    const int b = problem_of(kRound == 1 ? nullptr : ws.pending, ws.pending_count, blockIdx.x * PB + lane, B);
    const bool live = b >= 0 && ws.running[b];
    s_total[threadIdx.y][lane] = !live ? 0.0 : kSummed ? ws.cand_cost[(size_t)a * B + b] : dev_candidate_total(q, ws, b, a);
    __syncthreads();
    if (threadIdx.y != 0 || b < 0) return;
    if (!live) {                                             // stopped earlier: nothing to accept
        ws.winner[b] = -1;
        return;
    }
    const double before = q.traj_costs[b];
    double now;
    const int win = first_improving<PB>(s_total, lane, A0, NA, before, now);
    if (kRound == 1 && win < 0) {                            // alpha = 1 and 0.1 failed: try the rest
        ws.winner[b] = -1;
        ws.pending[atomicAdd(ws.pending_count, 1)] = b;
        return;
    }
    conclude_line_search(q, ws, b, win, before, now);
}

// copy the accepted candidate into x, u (and keep prev_x, prev_k) — thread per (problem, stage)
// S: element type of the candidates (scratch_t of the precision that rolled them out)
template <typename M, typename S>
__device__ __forceinline__ void dev_accept(const tplb_batch& q, const Workspace& ws, int b, int t) {
    using D = Dims<M>;
    constexpr int X = D::X, U = D::U;
    const int B = q.batch;
    const int win = ws.winner[b];
    const int T = horizon_of(q, b);
    if (win < 0 || t > T) return;
    const S* cx = scratch<S>(ws.cand_x) + (size_t)win * (q.t_max + 1) * X * B + b;
    const S* cu = scratch<S>(ws.cand_u) + (size_t)win * q.t_max * U * B + b;
#pragma unroll
    for (int i = 0; i < X; ++i) {
        const size_t idx = ((size_t)t * X + i) * B + b;
        if (q.keep_previous) q.prev_x[idx] = q.x[idx];
        q.x[idx] = (double)cx[((size_t)t * X + i) * B];
    }
    if (t < T) {
#pragma unroll
        for (int d = 0; d < U; ++d) {
            const size_t idx = ((size_t)t * U + d) * B + b;
            if (q.keep_previous) q.prev_k[idx] = q.k[idx];
            q.u[idx] = (double)cu[((size_t)t * U + d) * B];
        }
    }
}

template <typename M, typename S>
__global__ void accept_kernel(const __grid_constant__ tplb_batch q, Workspace ws) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= q.batch) return;
    dev_accept<M, S>(q, ws, b, blockIdx.y);                  // rows 0..T
}

// next_x / next_u of the reference (optim.c:1657-1659): the trajectory the last line search of each
// problem ended on — the accepted candidate, or the alpha = 1e-7 one after a failed search.
// Thread per (problem, stage); rows 0..T.
template <typename M>
__global__ void next_trajectory_kernel(const __grid_constant__ tplb_batch q, Workspace ws, double* next_x,
                                       double* next_u) {
    using D = Dims<M>;
    constexpr int X = D::X, U = D::U;
    const int b = blockIdx.x * blockDim.x + threadIdx.x, t = blockIdx.y, B = q.batch;
    if (b >= B) return;
    const int a = ws.last_tried[b];
    const int T = horizon_of(q, b);
    if (t > T) return;
    const bool have = a >= 0 && a < kAlphas;
    const double* cx = ws.cand_x + (size_t)(have ? a : 0) * (q.t_max + 1) * X * B + b;
    const double* cu = ws.cand_u + (size_t)(have ? a : 0) * q.t_max * U * B + b;
#pragma unroll
    for (int i = 0; i < X; ++i) {
        const size_t idx = ((size_t)t * X + i) * B;
        next_x[idx + b] = have ? cx[idx] : 0.0;
    }
    if (t < T) {
#pragma unroll
        for (int d = 0; d < U; ++d) {
            const size_t idx = ((size_t)t * U + d) * B;
            next_u[idx + b] = have ? cu[idx] : 0.0;
        }
    }
}

__device__ __forceinline__ void dev_finalize(const tplb_batch& q, int b, int lg_done) {
    q.lg_iterations[b] = lg_done;
    if (q.iterations[b] == q.max_iterations) q.termination_condition[b] = 1;   // optim.c:1147-1149
}

__global__ void finalize_kernel(const __grid_constant__ tplb_batch q, int lg_done) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= q.batch) return;
    dev_finalize(q, b, lg_done);
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
    const int T = horizon_of(q, b);
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
// point evaluations of the dynamics (optim.c:1512-1652); stage constants are
// evaluated on the fly for the requested (t, dt)
// ---------------------------------------------------------------------------------
template <typename M, typename R>
__global__ void dynamics_kernel(const __grid_constant__ tplb_batch q, const double* x_in, const double* u_in,
                                const int32_t* scene_of_point, int n, int t, double dt, int continuous,
                                double* x_out) {
    constexpr int X = M::X, U = M::U;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int scene = scene_of_point ? scene_of_point[i] : ((n == q.batch && q.scene_index) ? q.scene_index[i] : 0);
    const ParamView<R> P = param_view_scene<R>(q, scene);
    R x[X], u[U], out[X], sc[Dims<M>::NSCs];
#pragma unroll
    for (int j = 0; j < X; ++j) x[j] = R(x_in[(size_t)j * n + i]);
#pragma unroll
    for (int j = 0; j < U; ++j) u[j] = R(u_in[(size_t)j * n + i]);
    M::stage_constants(P, R(t), R(dt), sc);
    if (continuous) M::ct_dynamics(P, x, u, sc, R(t), R(dt), out);
    else if (q.integrator_type == TPLB_EULER) step_state<M, TPLB_EULER>(P, x, u, sc, R(t), R(dt), out);
    else if (q.integrator_type == TPLB_HEUN) step_state<M, TPLB_HEUN>(P, x, u, sc, R(t), R(dt), out);
    else step_state<M, TPLB_RK4>(P, x, u, sc, R(t), R(dt), out);
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
// self-test of the elementary functions the generated code uses (fast_math.cuh)
// fn: 0 sin, 1 cos, 2 tan, 3 1/x, 4 1/sqrt(x), 5 sqrt(x)
// ---------------------------------------------------------------------------------
__global__ void math_selftest_kernel(int fn, const double* x, int n, double* out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double v = x[i];
    double r = 0.0;
    switch (fn) {
        case 0: r = m_sin(v); break;
        case 1: r = m_cos(v); break;
        case 2: r = m_tan(v); break;
        case 3: r = m_inv(v); break;
        case 4: r = m_rsqrt(v); break;
        default: r = m_sqrt(v); break;
    }
    out[i] = r;
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
