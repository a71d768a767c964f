// extern "C" entry points declared in include/tplb200.h, compiled once per problem
// definition:  nvcc -gencode arch=compute_100a,code=sm_100a -DTPLB_MODEL_HEADER='"generated/<name>.cuh"'
//
// Host side of the reference's update() (optim.c:1091-1160): the outer
// augmented-Lagrangian loop and the inner iLQR loop are unrolled into a fixed
// sequence of kernel launches; per-problem progress (running / trajectory_changed /
// winner) lives on the device, so no launch depends on a device->host read.
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "solver.cuh"
#include "solo.cuh"

// Internal linkage: several solver libraries (one per problem definition) are loaded
// into the same process, and their `Model` types must not be merged by the dynamic
// linker (inline variables would otherwise get process-wide STB_GNU_UNIQUE binding).
namespace tplb {
namespace {
#include TPLB_MODEL_HEADER
}
}

using tplb::Model;
using Dm = tplb::Dims<Model>;

namespace {

thread_local char g_error[256] = "";

int fail(int code, const char* msg) {
    std::snprintf(g_error, sizeof g_error, "%s", msg);
    return code;
}

int check_launch(const char* what) {
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        std::snprintf(g_error, sizeof g_error, "%s: %s", what, cudaGetErrorString(e));
        return (int)e;
    }
    return 0;
}

int validate(const tplb_batch* q) {
    if (!q) return fail(TPLB_E_ARG, "batch is NULL");
    if (q->struct_bytes != (int32_t)sizeof(tplb_batch)) return fail(TPLB_E_ABI, "tplb_batch size mismatch");
    if (q->batch <= 0 || q->scenes <= 0) return fail(TPLB_E_ARG, "batch and scenes must be positive");
    if (q->t_max > TPLB_HORIZON_MAX || q->horizon < 1 || q->horizon > q->t_max)
        return fail(TPLB_E_HORIZON, "horizon must satisfy 1 <= horizon <= t_max <= 299");
    if (q->opt_start != 0) return fail(TPLB_E_UNSUPPORTED, "opt_start != 0 is not supported");
    if (q->integrator_type < TPLB_EULER || q->integrator_type > TPLB_RK4)
        return fail(TPLB_E_UNSUPPORTED, "integrator_type must be EULER, HEUN or RK4");
    if (q->precision != TPLB_FP64 && q->precision != TPLB_FP32)
        return fail(TPLB_E_UNSUPPORTED, "precision must be TPLB_FP64 or TPLB_FP32");
    if (q->line_search_rounds < 0 || q->line_search_rounds > 2)
        return fail(TPLB_E_ARG, "line_search_rounds must be 0, 1 or 2");
    if (q->single_launch < -1 || q->single_launch > 1 || q->reserved0 != 0)
        return fail(TPLB_E_ARG, "single_launch must be -1, 0 or 1");
    if (!q->x || !q->u || !q->k || !q->K || !q->u_min || !q->u_max || !q->traj_costs || !q->alpha ||
        !q->mu || !q->iterations || !q->lg_iterations || !q->mu_step || !q->trajectory_changed ||
        !q->improved || !q->termination_condition || !q->scene_index || !q->workspace)
        return fail(TPLB_E_ARG, "a required device pointer is NULL");
    if (Model::C > 0 && (!q->lagrange_multiplier || !q->barrier_weight || !q->lg_mult_limit))
        return fail(TPLB_E_ARG, "constraint buffers are NULL");
    if (Model::NUM_SCALARS > 0 && !q->scalars) return fail(TPLB_E_ARG, "scalars is NULL");
    for (int a = 0; a < Model::NUM_ARRAYS; ++a)
        if (q->array_len[a] > 0 && !q->arrays[a]) return fail(TPLB_E_ARG, "a parameter array is NULL");
    // the reference's check(ARR.ndim == ...) / check(ARR.dims[0] == XS.dims[0]) of the lookups (optim.c:372-486)
    for (int a = 0; a < Model::NUM_ARRAYS; ++a) {
        const int cols = q->array_cols[a];
        if (Model::ARRAY_NDIM[a] == 2 ? (cols <= 0 && q->array_len[a] > 0) || (cols > 0 && q->array_len[a] % cols != 0)
                                      : cols != 0)
            return fail(TPLB_E_ARG, "array_cols does not describe the array the problem definition reads (blerp: 2-D, others: 1-D)");
    }
    for (int p = 0; p < Model::NUM_WRAP_PAIRS; ++p)
        if (q->array_len[Model::WRAP_PAIRS[2 * p]] != q->array_len[Model::WRAP_PAIRS[2 * p + 1]])
            return fail(TPLB_E_ARG, "lerp_wrap: the sample positions and the samples differ in length");
    if (q->keep_previous && (!q->prev_x || !q->prev_k)) return fail(TPLB_E_ARG, "prev_x/prev_k are NULL");
    size_t need = 0;
    tplb::carve<Model>(nullptr, q->batch, q->scenes, q->t_max, &need);
    if (q->workspace_bytes < need) return fail(TPLB_E_ARG, "workspace too small");
    return 0;
}

// thread-per-problem kernels: one warp per block while the batch cannot fill the
// chip, so the resident warps spread over all 148 SMs
int problem_block_forced() {                               // TPLB_PROBLEM_BLOCK: block size of the sweep, for A/B runs
    static const int forced = [] { const char* e = std::getenv("TPLB_PROBLEM_BLOCK"); return e ? std::atoi(e) : 0; }();
    return forced;
}
int problem_block(int B) { return (B <= 148 * 32 * 8) ? 32 : 128; }

const char* const* names(const char* const* a) { return a; }

}  // namespace

extern "C" {

int32_t tplb_abi_version(void) { return TPLB_ABI_VERSION; }

const tplb_model_info* tplb_model(void) {
    static const tplb_model_info info = {
        TPLB_ABI_VERSION, Model::X, Model::U, Model::C,
        Model::NUM_SCALARS, Model::NUM_ARRAYS, Model::NUM_PARAMS,
        Model::NAME, Model::DEFINITION_SHA1,
        names(Model::STATE_NAMES), names(Model::ACTION_NAMES), names(Model::SCALAR_NAMES),
        names(Model::ARRAY_NAMES), names(Model::PARAM_ORDER),
        Dm::DENSE, Dm::OFF_FX, Dm::OFF_FU, Dm::OFF_LX, Dm::OFF_LU, Dm::OFF_LXX, Dm::OFF_LUU, Dm::OFF_LUX,
        Model::DERIV_COMPACT, Model::NUM_STAGE_CONSTS,
        Model::ARRAY_NDIM, Model::NUM_WRAP_PAIRS, Model::WRAP_PAIRS,
    };
    return &info;
}

const char* tplb_last_error(void) { return g_error; }

size_t tplb_workspace_bytes(int32_t batch, int32_t scenes, int32_t t_max) {
    size_t need = 0;
    tplb::carve<Model>(nullptr, batch, scenes, t_max, &need);
    return need;
}

void* tplb_workspace_counters(void* workspace, int32_t batch, int32_t scenes, int32_t t_max) {
    return tplb::carve<Model>(workspace, batch, scenes, t_max).counters;
}

}  // extern "C"

namespace {

// Per-kernel-class timing for tplb_update_profiled(): CUDA events between launches.
struct Profiler {
    cudaStream_t st;
    bool on;
    cudaEvent_t ev[2];
    float ms[TPLB_NUM_KERNEL_CLASSES];
    int launches[TPLB_NUM_KERNEL_CLASSES];
    explicit Profiler(cudaStream_t s, bool enable) : st(s), on(enable) {
        std::memset(ms, 0, sizeof ms);
        std::memset(launches, 0, sizeof launches);
        if (on) { cudaEventCreate(&ev[0]); cudaEventCreate(&ev[1]); }
    }
    ~Profiler() { if (on) { cudaEventDestroy(ev[0]); cudaEventDestroy(ev[1]); } }
    void before() { if (on) cudaEventRecord(ev[0], st); }
    void after(int cls) {
        launches[cls] += (cls == TPLB_K_ROLLOUT_INIT) ? 3 : 1;
        if (!on) return;
        cudaEventRecord(ev[1], st);
        cudaEventSynchronize(ev[1]);
        float t = 0.f;
        cudaEventElapsedTime(&t, ev[0], ev[1]);
        ms[cls] += t;
    }
};

int run_update(const tplb_batch* qp, void* stream_, Profiler& prof);

}  // namespace

extern "C" {

int32_t tplb_update(const tplb_batch* qp, void* stream_) {
    Profiler prof(static_cast<cudaStream_t>(stream_), false);
    return run_update(qp, stream_, prof);
}

int32_t tplb_update_profiled(const tplb_batch* qp, void* stream_, float* ms_by_class, int32_t* launches_by_class) {
    Profiler prof(static_cast<cudaStream_t>(stream_), true);
    const int rc = run_update(qp, stream_, prof);
    for (int i = 0; i < TPLB_NUM_KERNEL_CLASSES; ++i) {
        if (ms_by_class) ms_by_class[i] = prof.ms[i];
        if (launches_by_class) launches_by_class[i] = prof.launches[i];
    }
    return rc;
}

}  // extern "C"

namespace {

constexpr int PB = 32;                                     // problems per rollout / select block
// rollouts of candidates [a_begin, a_begin + NA) for every problem (list == NULL) or for the
// problems of the pending list
template <typename R, int NA, bool kInit, bool kCost = false>
void launch_rollout(const tplb_batch& q, const tplb::Workspace& ws, cudaStream_t st, int a_begin,
                    const int32_t* list) {
    const dim3 grid((q.batch + PB - 1) / PB), block(PB, NA);
    const size_t smem = tplb::rollout_smem_bytes<Model, kInit, kCost>(PB, NA);
#define TPLB_ROLLOUT(SCHEME)                                                                          \
    do {                                                                                              \
        auto kern = tplb::rollout_kernel<Model, R, PB, NA, kInit, SCHEME, kCost>;                     \
        if (smem > 48 * 1024)   /* opt in per launch: the attribute belongs to the current device */ \
            cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);       \
        kern<<<grid, block, smem, st>>>(q, ws, a_begin, list);                                        \
    } while (0)
    switch (q.integrator_type) {
        case TPLB_EULER: TPLB_ROLLOUT(TPLB_EULER); break;
        case TPLB_HEUN: TPLB_ROLLOUT(TPLB_HEUN); break;
        default: TPLB_ROLLOUT(TPLB_RK4); break;
    }
#undef TPLB_ROLLOUT
}

// tplb_batch.line_search_rounds: 2 selects the throughput sequence, 1 the latency sequence,
// 0 decides by batch size.  Measured on B200, one batch alone (scripts/plan_crossover.sh): with the
// fused sweep of round 2 the throughput sequence wins from about 8192 problems on for the models
// whose sweep fits the register file (6x2 bicycle: 4096 problems 4.0 vs 4.4 ms, 8192 5.5 vs 4.8 ms,
// 16384 8.8 vs 6.0 ms; lateral 2x1: even at 8192); the 7x2 model with lookups in its dynamics
// (its fused sweep spills) still prefers the latency sequence at 8192 (15.3 vs 19.2 ms).
bool throughput_sequence(const tplb_batch& q) {
    if (q.line_search_rounds != 0) return q.line_search_rounds == 2;
    constexpr bool lean_sweep = Model::X <= 6 && Model::DYNAMICS_LOOKUPS == 0;
    return q.batch >= (lean_sweep ? 8192 : 16384);
}

template <typename R>
int run_update_as(const tplb_batch* qp, void* stream_, Profiler& prof);

// precision 0: every kernel computes in fp64; 1: the SEARCH DIRECTION (derivative records, Riccati
// recursion, gains) is computed in fp32, while rollouts, stage costs, cost sums and every accept /
// stop decision stay fp64 — an inexact Newton direction changes the path of the iteration, not
// the point it converges to
// ---- one launch, one thread block per problem (solo.cuh) --------------------------------------
size_t solo_smem_limit() {
    static int limit = -1;                                 // of the first device used; B200 boxes are uniform
    if (limit < 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        if (cudaDeviceGetAttribute(&limit, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev) != cudaSuccess) limit = 0;
    }
    return (size_t)limit;
}

size_t solo_smem_bytes(const tplb_batch& q) {
    int samples = 0;
    for (int a = 0; a < Model::NUM_ARRAYS; ++a) samples += q.array_len[a];
    return tplb::SoloLayout<Model>::bytes(q.horizon, samples);
}

bool solo_eligible(const tplb_batch& q) {
    return q.precision == TPLB_FP64 && q.use_quadratic_terms && solo_smem_bytes(q) <= solo_smem_limit();
}

// automatic choice: as long as every problem gets its own SM (measured crossover on B200)
int solo_auto_batch() {
    static const int n = [] {
        const char* e = std::getenv("TPLB_SOLO_MAX_BATCH");
        return e ? std::atoi(e) : 296;
    }();
    return n;
}

int run_solo(const tplb_batch& q, cudaStream_t st, Profiler& prof) {
    const tplb::Workspace ws = tplb::carve<Model>(q.workspace, q.batch, q.scenes, q.t_max);
    const size_t smem = solo_smem_bytes(q);
    prof.before();
#define TPLB_SOLO(SCHEME)                                                                             \
    do {                                                                                              \
        auto kern = tplb::solo_update_kernel<Model, SCHEME>;                                          \
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);           \
        kern<<<q.batch, tplb::kSoloThreads, smem, st>>>(q, ws);                                       \
    } while (0)
    switch (q.integrator_type) {
        case TPLB_EULER: TPLB_SOLO(TPLB_EULER); break;
        case TPLB_HEUN: TPLB_SOLO(TPLB_HEUN); break;
        default: TPLB_SOLO(TPLB_RK4); break;
    }
#undef TPLB_SOLO
    prof.after(TPLB_K_SOLO);
    return check_launch("tplb_update (single launch)");
}

int run_update(const tplb_batch* qp, void* stream_, Profiler& prof) {
    if (int e = validate(qp)) return e;
    const bool solo = qp->single_launch > 0 || (qp->single_launch == 0 && qp->batch <= solo_auto_batch() && solo_eligible(*qp));
    if (solo) {
        if (!solo_eligible(*qp))
            return fail(TPLB_E_UNSUPPORTED, "single_launch needs fp64, use_quadratic_terms and a problem that fits shared memory");
        return run_solo(*qp, static_cast<cudaStream_t>(stream_), prof);
    }
    return qp->precision == TPLB_FP32 ? run_update_as<float>(qp, stream_, prof)
                                      : run_update_as<double>(qp, stream_, prof);
}

template <typename R>
int run_update_as(const tplb_batch* qp, void* stream_, Profiler& prof) {
    const tplb_batch q = *qp;
    cudaStream_t st = static_cast<cudaStream_t>(stream_);
    const tplb::Workspace ws = tplb::carve<Model>(q.workspace, q.batch, q.scenes, q.t_max);
    const int B = q.batch, T = q.horizon, S = q.scenes;
    const int pb = problem_block(B);
    const dim3 pgrid((B + pb - 1) / pb);
    const int sb = 128;                                    // stage-parallel kernels
    const unsigned sgx = (B + sb - 1) / sb;
    const size_t cx_stride = (size_t)(q.t_max + 1) * Model::X * B;
    const size_t cu_stride = (size_t)q.t_max * Model::U * B;
    constexpr int R1 = tplb::kRound1, R2 = tplb::kAlphas - tplb::kRound1;
    using SC = double;                                     // storage of the line-search candidates
    // Two launch sequences with identical results:
    //   latency    — the batch cannot fill the GPU: fewest launches.  All 8 step sizes roll out at
    //                once, stage costs of the candidates in stage-parallel kernels, the accepted
    //                step is installed by the next linearize.
    //   throughput — GPU full, every kernel streams from HBM: fewest bytes.  Two-round rollouts
    //                that add up their own stage costs, stand-alone accept (high occupancy, at
    //                the HBM roofline) in front of a plain linearize.
    const bool throughput = throughput_sequence(q);
    const bool fold_accept = !throughput;
    // throughput sequence of the second-order solver: one fused kernel per iteration in front of
    // the rollouts (TPLB_NO_FUSED_SWEEP=1 in the environment keeps the three separate kernels, for A/B runs)
    static const bool no_fused = std::getenv("TPLB_NO_FUSED_SWEEP") != nullptr;
    const bool fused_sweep = throughput && q.use_quadratic_terms && !no_fused;
    // the sweep of the throughput sequence runs with other batches' kernels around it: four warps per
    // block (measured +2 % over one-warp blocks with 24 batches in flight)
    const int sweep_block = problem_block_forced() > 0 ? problem_block_forced() : 128;
    const dim3 sweep_grid((B + sweep_block - 1) / sweep_block);

    prof.before();
    tplb::stage_constants_kernel<Model><<<dim3((S + sb - 1) / sb, T + 1), sb, 0, st>>>(q, ws);
    prof.after(TPLB_K_STAGE_CONSTS);

    prof.before();
    launch_rollout<double, 1, true>(q, ws, st, 0, nullptr);
    tplb::stage_cost_kernel<Model, double, double><<<dim3(sgx, T + 1, 1), sb, 0, st>>>(
        q, ws, (const double*)q.x, (const double*)q.u, 0, 0, 0, 0, nullptr);
    tplb::init_cost_kernel<<<sgx, sb, 0, st>>>(q, ws);
    prof.after(TPLB_K_ROLLOUT_INIT);

    int lg = 0;
    for (; lg < q.max_lg_iterations; ++lg) {
        prof.before();
        tplb::multiplier_kernel<Model, double><<<dim3(sgx, Model::C > 0 ? T : 1), sb, 0, st>>>(q, ws);
        prof.after(TPLB_K_MULTIPLIER);
        for (int s = 0; s < q.max_iterations; ++s) {
            if (fused_sweep) {
                // linearise + Riccati in one pass over the stages; installs the step the previous
                // line search accepted on the way (no accept / linearize / record traffic)
                if (s > 0 && q.keep_previous) {
                    prof.before();
                    tplb::keep_previous_kernel<Model><<<dim3(sgx, T + 1), sb, 0, st>>>(q, ws);
                    prof.after(TPLB_K_ACCEPT);
                }
                prof.before();
                tplb::sweep_kernel<Model, R><<<sweep_grid, sweep_block, 0, st>>>(q, ws, s);
                prof.after(TPLB_K_BACKWARD);
            } else {
                prof.before();
                if (s == 0 || !fold_accept) tplb::linearize_kernel<Model, R, false, false><<<dim3(sgx, T), sb, 0, st>>>(q, ws);
                else tplb::linearize_kernel<Model, R, false, true><<<dim3(sgx, T + 1), sb, 0, st>>>(q, ws);
                prof.after(TPLB_K_LINEARIZE);
                prof.before();
                if (q.use_quadratic_terms)
                    tplb::backward_kernel<Model, R><<<pgrid, pb, 0, st>>>(q, ws, s);
                else
                    tplb::backward_first_order_kernel<Model, R><<<pgrid, pb, 0, st>>>(q, ws, s);
                prof.after(TPLB_K_BACKWARD);
            }

            if (throughput) {
                // round 1: alpha = 1, 0.1; round 2: the other six for the problems still pending
                prof.before();
                launch_rollout<double, R1, false, true>(q, ws, st, 0, nullptr);
                prof.after(TPLB_K_ROLLOUT);
                prof.before();
                tplb::select_kernel<PB, 1, true><<<(B + PB - 1) / PB, dim3(PB, R1), 0, st>>>(q, ws);
                prof.after(TPLB_K_SELECT);
                prof.before();
                launch_rollout<double, R2, false, true>(q, ws, st, R1, ws.pending);
                prof.after(TPLB_K_ROLLOUT);
                prof.before();
                tplb::select_kernel<PB, 2, true><<<(B + PB - 1) / PB, dim3(PB, R2), 0, st>>>(q, ws);
                prof.after(TPLB_K_SELECT);
            } else {
                prof.before();
                launch_rollout<double, tplb::kAlphas, false>(q, ws, st, 0, nullptr);
                prof.after(TPLB_K_ROLLOUT);
                // round 1: alpha = 1, 0.1
                prof.before();
                tplb::stage_cost_round1_kernel<Model, double><<<dim3(sgx, T + 1), sb, 0, st>>>(q, ws);
                prof.after(TPLB_K_STAGE_COST);
                prof.before();
                tplb::select_kernel<PB, 1><<<(B + PB - 1) / PB, dim3(PB, R1), 0, st>>>(q, ws);
                prof.after(TPLB_K_SELECT);
                // round 2: alpha = 1e-2 .. 1e-7 for the problems still pending
                prof.before();
                tplb::stage_cost_kernel<Model, double, SC><<<dim3(sgx < 8 ? sgx : 8, T + 1, R2), sb, 0, st>>>(
                    q, ws, (const SC*)tplb::scratch<SC>(ws.cand_x), (const SC*)tplb::scratch<SC>(ws.cand_u), cx_stride,
                    cu_stride, 1, R1, ws.pending);
                prof.after(TPLB_K_STAGE_COST);
                prof.before();
                tplb::select_kernel<PB, 2><<<(B + PB - 1) / PB, dim3(PB, R2), 0, st>>>(q, ws);
                prof.after(TPLB_K_SELECT);
            }
            if (!fold_accept && !fused_sweep && s + 1 < q.max_iterations) {
                prof.before();
                tplb::accept_kernel<Model, SC><<<dim3(sgx, T + 1), sb, 0, st>>>(q, ws);
                prof.after(TPLB_K_ACCEPT);
            }
        }
        if (q.max_iterations > 0) {
            prof.before();
            tplb::accept_kernel<Model, SC><<<dim3(sgx, T + 1), sb, 0, st>>>(q, ws);
            prof.after(TPLB_K_ACCEPT);
        }
    }
    prof.before();
    tplb::finalize_kernel<<<sgx, sb, 0, st>>>(q, lg);
    prof.after(TPLB_K_FINALIZE);
    return check_launch("tplb_update");
}

}  // namespace

extern "C" {

int32_t tplb_expand_derivatives(const tplb_batch* qp, void* stream_) {
    if (int e = validate(qp)) return e;
    if (!qp->deriv_dense) return fail(TPLB_E_ARG, "deriv_dense is NULL");
    const tplb_batch q = *qp;
    const tplb::Workspace ws = tplb::carve<Model>(q.workspace, q.batch, q.scenes, q.t_max);
    const dim3 grid((q.batch + 127) / 128, q.horizon);
    tplb::expand_derivatives_kernel<Model><<<grid, 128, 0, static_cast<cudaStream_t>(stream_)>>>(q, ws, q.deriv_dense);
    return check_launch("tplb_expand_derivatives");
}

int32_t tplb_linearize(const tplb_batch* qp, void* stream_) {
    if (int e = validate(qp)) return e;
    if (!qp->deriv_dense) return fail(TPLB_E_ARG, "deriv_dense is NULL");
    const tplb_batch q = *qp;
    cudaStream_t st = static_cast<cudaStream_t>(stream_);
    const tplb::Workspace ws = tplb::carve<Model>(q.workspace, q.batch, q.scenes, q.t_max);
    const dim3 grid((q.batch + 127) / 128, q.horizon);
    tplb::stage_constants_kernel<Model><<<dim3((q.scenes + 127) / 128, q.horizon + 1), 128, 0, st>>>(q, ws);
    tplb::linearize_kernel<Model, double, true, false><<<grid, 128, 0, st>>>(q, ws);
    tplb::expand_derivatives_kernel<Model><<<grid, 128, 0, st>>>(q, ws, q.deriv_dense);
    return check_launch("tplb_linearize");
}

int32_t tplb_next_trajectory(const tplb_batch* qp, double* next_x, double* next_u, void* stream_) {
    if (int e = validate(qp)) return e;
    if (!next_x || !next_u) return fail(TPLB_E_ARG, "tplb_next_trajectory: output is NULL");
    const tplb_batch q = *qp;
    const tplb::Workspace ws = tplb::carve<Model>(q.workspace, q.batch, q.scenes, q.t_max);
    const dim3 grid((q.batch + 127) / 128, q.horizon + 1);
    tplb::next_trajectory_kernel<Model><<<grid, 128, 0, static_cast<cudaStream_t>(stream_)>>>(q, ws, next_x, next_u);
    return check_launch("tplb_next_trajectory");
}

int32_t tplb_shift(const tplb_batch* qp, int32_t amount, const int32_t* amounts, void* stream_) {
    if (int e = validate(qp)) return e;
    const tplb_batch q = *qp;
    tplb::shift_kernel<Model><<<(q.batch + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream_)>>>(q, amount, amounts);
    return check_launch("tplb_shift");
}

int32_t tplb_dynamics(const tplb_batch* qp, const double* x_in, const double* u_in,
                      const int32_t* scene_of_point, int32_t n, int32_t t, double dt,
                      int32_t continuous, double* x_out, void* stream_) {
    if (!qp || !x_in || !u_in || !x_out || n <= 0) return fail(TPLB_E_ARG, "tplb_dynamics: bad argument");
    if (qp->struct_bytes != (int32_t)sizeof(tplb_batch)) return fail(TPLB_E_ABI, "tplb_batch size mismatch");
    const tplb_batch q = *qp;
    if (q.precision == TPLB_FP32)
        tplb::dynamics_kernel<Model, float><<<(n + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream_)>>>(
            q, x_in, u_in, scene_of_point, n, t, dt, continuous, x_out);
    else
        tplb::dynamics_kernel<Model, double><<<(n + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream_)>>>(
            q, x_in, u_in, scene_of_point, n, t, dt, continuous, x_out);
    return check_launch("tplb_dynamics");
}

int32_t tplb_argmin_groups(const double* traj_costs, int32_t groups, int32_t per_group,
                           double* min_cost, int32_t* arg_min, void* stream_) {
    if (!traj_costs || !min_cost || !arg_min || groups <= 0 || per_group <= 0)
        return fail(TPLB_E_ARG, "tplb_argmin_groups: bad argument");
    const int threads = 128, warps_per_block = threads / 32;
    tplb::argmin_groups_kernel<<<(groups + warps_per_block - 1) / warps_per_block, threads, 0,
                                 static_cast<cudaStream_t>(stream_)>>>(traj_costs, groups, per_group, min_cost, arg_min);
    return check_launch("tplb_argmin_groups");
}

int32_t tplb_selftest_math(int32_t fn, const double* x, int32_t n, double* out, void* stream_) {
    if (!x || !out || n <= 0 || fn < 0 || fn > 5) return fail(TPLB_E_ARG, "tplb_selftest_math: bad argument");
    tplb::math_selftest_kernel<<<(n + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream_)>>>(fn, x, n, out);
    return check_launch("tplb_selftest_math");
}

#ifdef TPLB_SOLO_TIMING
// development aid: cycles per phase of solo_update_kernel (block 0), accumulated since the last call
__attribute__((visibility("default"))) int32_t tplb_debug_solo_cycles(long long* out) {
    long long zero[8] = {0};
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(out, tplb::g_solo_cycles, sizeof zero);
    cudaMemcpyToSymbol(tplb::g_solo_cycles, zero, sizeof zero);
    return 0;
}
#endif

double tplb_measure_fp64_tflops(int32_t repeats, void* stream_) {
    cudaStream_t st = static_cast<cudaStream_t>(stream_);
    int dev = 0, sms = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return -1.0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    double* sink = nullptr;
    if (cudaMalloc(&sink, sizeof(double)) != cudaSuccess) return -1.0;
    const int blocks = sms * 8, threads = 256, inner = 1 << 14;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    tplb::dfma_peak_kernel<<<blocks, threads, 0, st>>>(sink, inner);         // warm-up
    double best = -1.0;
    for (int r = 0; r < (repeats > 0 ? repeats : 3); ++r) {
        cudaEventRecord(e0, st);
        tplb::dfma_peak_kernel<<<blocks, threads, 0, st>>>(sink, inner);
        cudaEventRecord(e1, st);
        cudaEventSynchronize(e1);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        const double flops = 2.0 * 8.0 * (double)inner * blocks * threads;
        const double tf = flops / (ms * 1e-3) / 1e12;
        if (tf > best) best = tf;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(sink);
    return best;
}

}  // extern "C"
