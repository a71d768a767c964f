/*
 * tplb200.h — C ABI of the batched B200 iLQR solver (one shared library per problem
 * definition, e.g. tpl_b200/lib/libtplb200_trajectory_tracking_mpc_time.so).
 *
 * It replaces, for a batch of independent problems, the interface of the reference's
 * generated CPython extension type `genopt<sha1>.Optim`
 * (/root/reference/library/tpl/optim/templates/optim.c:1485-1892, "optim.c" below):
 *
 *   reference (one problem)                        this ABI (B problems)
 *   -----------------------------------------      -----------------------------------------
 *   Optim.update()        optim.c:1485-1494  ->    tplb_update()
 *   Optim.shift(n)        optim.c:1496-1510  ->    tplb_shift()
 *   Optim.dynamics()      optim.c:1512-1581  ->    tplb_dynamics(continuous = 0)
 *   Optim.ct_dynamics()   optim.c:1583-1652  ->    tplb_dynamics(continuous = 1)
 *   array getters fx..lux optim.c:1663-1669  ->    tplb_expand_derivatives() / tplb_linearize()
 *   struct Optim          optim.c:508-622    ->    tplb_batch (device pointers, SoA)
 *   struct Params/DynArray optim.c:297-328   ->    tplb_batch.scalars / .arrays
 *
 * Conventions
 *  - plain C, no exceptions, no host allocation inside any call; every call only
 *    enqueues kernels on `stream` (a cudaStream_t passed as void*) and returns.
 *  - every pointer inside tplb_batch is a DEVICE pointer owned by the caller (the
 *    Python host uses torch tensors purely as these buffers).
 *  - layout is structure-of-arrays with the PROBLEM INDEX FASTEST: element
 *    (stage t, component i, problem b) of a trajectory lives at [(t*N + i)*batch + b].
 *    Shapes below are written slowest-to-fastest.
 *  - all reals are stored as IEEE double; all integers int32.
 *  - return value: 0 on success, TPLB_E_* (negative) for rejected arguments, or a
 *    positive cudaError_t from a failed launch.  tplb_last_error() describes it.
 */
#ifndef TPLB200_H
#define TPLB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define TPLB_API __attribute__((visibility("default")))
#else
#define TPLB_API
#endif

#define TPLB_ABI_VERSION 3
#define TPLB_MAX_ARRAYS 16
#define TPLB_LINE_SEARCH_STEPS 8          /* alpha = 10^-i, i = 0..7, optim.c:861-863 */
#define TPLB_HORIZON_MAX 299              /* H_MAX - 1, optim.c:49, 1732 */

enum { TPLB_EULER = 0, TPLB_HEUN = 1, TPLB_RK4 = 2 };           /* optim.c:492-496 */
enum { TPLB_FP64 = 0, TPLB_FP32 = 1 };                          /* arithmetic of the kernels */
enum {
    TPLB_E_ARG = -1,          /* null pointer / non-positive size */
    TPLB_E_HORIZON = -2,      /* horizon outside 1..t_max or t_max > TPLB_HORIZON_MAX */
    TPLB_E_UNSUPPORTED = -3,  /* e.g. opt_start != 0 */
    TPLB_E_ABI = -4           /* struct_bytes does not match this library */
};

/* Static description of the problem this library was generated for
 * (what `Optim.__slots__` / `Params.__slots__` expose, optim.c:1736-1782, genopt.py:341-353). */
typedef struct {
    int32_t abi_version;
    int32_t X, U, C;                       /* state, control, constraint dimensions */
    int32_t num_scalars, num_arrays, num_params;
    const char* name;
    const char* definition_sha1;
    const char* const* state_names;        /* [X] */
    const char* const* action_names;       /* [U] */
    const char* const* scalar_names;       /* [num_scalars] */
    const char* const* array_names;        /* [num_arrays] */
    const char* const* param_order;        /* [num_params] declaration order */
    int32_t deriv_stride;                  /* doubles per stage of a DENSE derivative record */
    int32_t off_fx, off_fu, off_lx, off_lu, off_lxx, off_luu, off_lux;   /* offsets in a dense record */
    int32_t deriv_compact;                 /* doubles per stage actually stored by the solver */
    int32_t num_stage_consts;              /* per-(scene, stage) constants hoisted out of the kernels */
    const int32_t* array_ndim;             /* [num_arrays] 2: read by blerp (optim.c:483), 1: every other lookup */
    int32_t num_wrap_pairs;                /* lerp_wrap calls (optim.c:450): (xs, arr) array indices, equal lengths */
    const int32_t* wrap_pairs;             /* [2 * num_wrap_pairs] */
} tplb_model_info;

/* One batch of B independent problems sharing horizon and solver settings.
 * Field comments give the reference member each one replaces (optim.c:508-622). */
typedef struct {
    int32_t struct_bytes;          /* sizeof(tplb_batch), ABI check */
    int32_t batch;                 /* B */
    int32_t scenes;                /* S parameter sets; problem b uses scene_index[b] */
    int32_t horizon;               /* T */
    int32_t t_max;                 /* stage capacity of the buffers (>= horizon) */

    /* settings */
    int32_t opt_start;             /* optStart; only 0 is supported (no caller sets it) */
    int32_t max_iterations;        /* maxIterations */
    int32_t max_lg_iterations;     /* maxLgIterations */
    int32_t integrator_type;       /* integratorType */
    int32_t use_quadratic_terms;   /* useQuadraticTerms: 1 = iLQR, 0 = gradient-only ilr */
    int32_t keep_previous;         /* 1: maintain prev_x / prev_k on accepted steps */
    int32_t precision;             /* TPLB_FP64 (default, the reference's arithmetic) or TPLB_FP32: the search
                                      direction — derivative records, Riccati recursion, gains — is computed
                                      (and the records stored) in fp32; rollouts, stage costs, cost sums, the
                                      accept / stop decisions and every array of this struct stay fp64, so the
                                      iteration converges to the fp64 solution.  Needs locally centred
                                      coordinates. */
    int32_t line_search_rounds;    /* how the 8 step sizes of optim.c:859-873 are rolled out; the accepted
                                      step is the reference's in every mode (bit-identical results):
                                      1 = all 8 at once (lowest latency of one batch),
                                      2 = alpha = 1, 0.1 first, the other six only for the problems that
                                          need them (least work: best when the GPU is full, i.e. large
                                          batches or several batches in flight on different streams),
                                      0 = choose by batch size (2 from 8192 problems on; from 16384 on for problem
                                          definitions with more than 6 states or lookups in the dynamics). */
    int32_t keep_records;          /* 1: the derivative records of the last linearisation stay readable through
                                      tplb_expand_derivatives() after update().  The sequences for small
                                      batches always keep them; the fused sweep of the throughput sequence
                                      stores them only on request (29 extra doubles per stage and iteration
                                      for the bicycle model). */
    int32_t single_launch;         /* the whole update() in ONE launch, one thread block per problem, the problem
                                      resident in shared memory (csrc/solo.cuh) — the path for a single
                                      planning / control cycle or a handful of candidates:
                                      0 = automatic (small fp64 second-order batches that fit shared memory),
                                      1 = always (TPLB_E_UNSUPPORTED if the problem does not fit), -1 = never.
                                      Results are identical to the batched launch sequences. */
    int32_t reserved0;             /* must be 0 */
    double dt;                     /* dt ("step") */
    double min_rel_cost_change;    /* minRelCostChange */

    /* trajectories */
    double* x;                     /* x       [t_max+1][X][B] */
    double* u;                     /* u       [t_max][U][B]   */
    double* prev_x;                /* prev_x  [t_max+1][X][B] (may be NULL if !keep_previous) */
    double* prev_k;                /* prev_k  [t_max][U][B]   (may be NULL if !keep_previous) */
    double* k;                     /* k       [t_max][U][B]   */
    double* K;                     /* K       [t_max][U*X][B] */
    double* g;                     /* g       [t_max][U][B]   (gradient-only mode; may be NULL otherwise) */

    /* constraints handling */
    double* lagrange_multiplier;   /* lagrangeMultiplier [t_max][C][B] */
    const double* barrier_weight;  /* barrierWeight [C][B] */
    const double* lg_mult_limit;   /* lgMultLimit   [C][B] */
    const double* u_min;           /* uMin [t_max][U][B] */
    const double* u_max;           /* uMax [t_max][U][B] */

    /* per-problem status, sticky across calls like the reference struct */
    double* traj_costs;            /* trajCosts [B] */
    double* alpha;                 /* alpha     [B] */
    double* mu;                    /* mu        [B] */
    int32_t* iterations;           /* iterations [B] */
    int32_t* lg_iterations;        /* lgIterations [B] */
    int32_t* mu_step;              /* muStep [B] */
    int32_t* trajectory_changed;   /* trajectoryChanged [B] */
    int32_t* improved;             /* improved [B] */
    int32_t* termination_condition;/* terminationCondition [B] */

    /* parameters (struct Params, optim.c:297-328 + genopt.py:321-335) */
    const int32_t* scene_index;    /* [B] */
    const double* scalars;         /* [num_scalars][S] */
    const double* arrays[TPLB_MAX_ARRAYS];   /* array a: [S][array_len[a]] row-major */
    int32_t array_len[TPLB_MAX_ARRAYS];      /* samples per scene; rows * cols of a 2-D array */
    int32_t array_cols[TPLB_MAX_ARRAYS];     /* DynArray.dims[1] of a 2-D array (row length), 0 for 1-D arrays */

    /* scratch owned by the caller, sized by tplb_workspace_bytes():
     *   stage_consts [t_max+1][num_stage_consts][S]  interpolation lookups that depend on the stage only
     *   deriv        [t_max][deriv_compact][B]  the entries of fx,fu,lx,lu,lxx,luu,lux that are not
     *                                           identically 0 or 1 for this problem definition
     *   cand_x       [8][t_max+1][X][B]         line-search candidates (next_x of each alpha)
     *   cand_u       [8][t_max][U][B]
     *   cost_terms   [8][t_max+1][B]            stage / end cost of every candidate and stage
     *   cand_cost    [8][B], winner [B], running [B], counters [3][B], pending [B] + count  */
    void* workspace;
    size_t workspace_bytes;

    /* optional: dense derivative records [t_max][deriv_stride][B] in the reference's layout
     * fx|fu|lx|lu|lxx|luu|lux (row-major each); written only by tplb_expand_derivatives() */
    double* deriv_dense;

    /* optional per-problem horizon T_b [B], 1 <= T_b <= horizon (NULL: `horizon` for every problem).  Each
     * reference Optim object owns its T (optim.c:508-622; PathOptim sets it every cycle from the length of
     * the local map, path_optim.py:126); stages t >= T_b of problem b are neither read nor written. */
    const int32_t* horizons;
} tplb_batch;

TPLB_API int32_t tplb_abi_version(void);
TPLB_API const tplb_model_info* tplb_model(void);
TPLB_API const char* tplb_last_error(void);

/* Bytes of device scratch a batch of this shape needs. */
TPLB_API size_t tplb_workspace_bytes(int32_t batch, int32_t scenes, int32_t t_max);

/* Device address of the [3][B] int32 work counters of the last update(): linearisations,
 * backward sweeps and the rollouts a sequential line search would have executed
 * (accepted index + 1, or 8 on failure; + 1 for the initial rollout).  They feed the
 * algorithmic-flop count of the roofline. */
TPLB_API void* tplb_workspace_counters(void* workspace, int32_t batch, int32_t scenes, int32_t t_max);

/* One `update()` for every problem of the batch (optim.c:1091-1160): initial rollout
 * and cost, then max_lg_iterations x { multiplier update; up to max_iterations x
 * { linearise (if the trajectory changed); backward Riccati sweep; 8-step line
 * search; regularisation schedule; relative-change stop } }, termination flags. */
TPLB_API int32_t tplb_update(const tplb_batch* batch, void* stream);

/* Same launches as tplb_update() with a CUDA-event pair around every launch, so the
 * time spent in each kernel class and the launch counts can be attributed (measurement
 * aid: it synchronises after every launch).  Either output may be NULL. */
enum {
    TPLB_K_STAGE_CONSTS = 0,  /* stage_constants_kernel                                   */
    TPLB_K_ROLLOUT_INIT = 1,  /* rollout_kernel<init> + stage_cost_kernel + init_cost_kernel */
    TPLB_K_MULTIPLIER = 2,    /* multiplier_kernel                                         */
    TPLB_K_LINEARIZE = 3,     /* linearize_kernel                                          */
    TPLB_K_BACKWARD = 4,      /* backward_kernel, or sweep_kernel (linearise + Riccati fused) */
    TPLB_K_ROLLOUT = 5,       /* rollout_kernel<line search>: the step-size candidates     */
    TPLB_K_STAGE_COST = 6,    /* stage_cost_kernel on the candidates (rounds 1 and 2)      */
    TPLB_K_SELECT = 7,        /* select_kernel (rounds 1 and 2)                            */
    TPLB_K_ACCEPT = 8,        /* accept_kernel (after the last iteration; otherwise folded
                                 into the next linearize_kernel)                           */
    TPLB_K_FINALIZE = 9,      /* finalize_kernel                                           */
    TPLB_K_SOLO = 10,         /* solo_update_kernel: the whole update() in one launch      */
    TPLB_NUM_KERNEL_CLASSES = 11
};
TPLB_API int32_t tplb_update_profiled(const tplb_batch* batch, void* stream,
                                      float* ms_by_class, int32_t* launches_by_class);

/* Derivative records of the CURRENT trajectory (optim.c:896-912) for every problem
 * regardless of solver state, expanded into batch->deriv_dense. */
TPLB_API int32_t tplb_linearize(const tplb_batch* batch, void* stream);

/* Expand the records the last update() linearised (what the reference's fx..lux getters
 * show after update(), optim.c:1663-1669) into batch->deriv_dense. */
TPLB_API int32_t tplb_expand_derivatives(const tplb_batch* batch, void* stream);

/* next_x [t_max+1][X][B] / next_u [t_max][U][B] of the reference (optim.c:1657-1659): for every problem the
 * trajectory its last line search ended on — the accepted step, or the alpha = 1e-7 candidate after
 * a failed search; zeros before the first search. */
TPLB_API int32_t tplb_next_trajectory(const tplb_batch* batch, double* next_x, double* next_u, void* stream);

/* Warm-start shift by `amount` stages for all problems, or by amounts[b] when
 * `amounts` (device, [B]) is not NULL (optim.c:1162-1177). */
TPLB_API int32_t tplb_shift(const tplb_batch* batch, int32_t amount, const int32_t* amounts, void* stream);

/* n point evaluations of the discrete (integrator_type) or continuous dynamics:
 * x_in [X][n], u_in [U][n] -> x_out [X][n]; point i uses scene scene_of_point[i]
 * (NULL: batch->scene_index when n == batch, else scene 0).  optim.c:1512-1652. */
TPLB_API int32_t tplb_dynamics(const tplb_batch* batch, const double* x_in, const double* u_in,
                      const int32_t* scene_of_point, int32_t n, int32_t t, double dt,
                      int32_t continuous, double* x_out, void* stream);

/* Multi-start reduction: problems are grouped contiguously, `per_group` each; writes
 * the smallest finite traj_costs of each group and the index (within the batch) of
 * the problem attaining it (lowest index wins ties; -1 if none is finite). */
TPLB_API int32_t tplb_argmin_groups(const double* traj_costs, int32_t groups, int32_t per_group,
                           double* min_cost, int32_t* arg_min, void* stream);

/* Diagnostic: evaluates the straight-line elementary functions the generated model code uses
 * (csrc/fast_math.cuh) on n device values: fn 0 sin, 1 cos, 2 tan, 3 1/x, 4 1/sqrt(x), 5 sqrt(x). */
TPLB_API int32_t tplb_selftest_math(int32_t fn, const double* x, int32_t n, double* out, void* stream);

/* Peak of the FP64 pipe, measured with a register-resident DFMA loop on the current
 * device (the roofline denominator for this path); returns TFLOP/s, <0 on error. */
TPLB_API double tplb_measure_fp64_tflops(int32_t repeats, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* TPLB200_H */
