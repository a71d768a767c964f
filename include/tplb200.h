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
 *   array getters fx..lux optim.c:1663-1669  ->    tplb_linearize() fills the same blocks
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
 *  - all reals are IEEE double; all integers int32.
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

#define TPLB_ABI_VERSION 1
#define TPLB_MAX_ARRAYS 16
#define TPLB_LINE_SEARCH_STEPS 8          /* alpha = 10^-i, i = 0..7, optim.c:861-863 */
#define TPLB_HORIZON_MAX 299              /* H_MAX - 1, optim.c:49, 1732 */

enum { TPLB_EULER = 0, TPLB_HEUN = 1, TPLB_RK4 = 2 };           /* optim.c:492-496 */
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
    int32_t deriv_stride;                  /* doubles per stage in tplb_batch.deriv */
    int32_t off_fx, off_fu, off_lx, off_lu, off_lxx, off_luu, off_lux;   /* offsets in a stage block */
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
    int32_t reserved0;
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
    int32_t array_len[TPLB_MAX_ARRAYS];

    /* scratch owned by the caller, sized by tplb_workspace_bytes():
     *   deriv      [t_max][deriv_stride][B]  fx,fu,lx,lu,lxx,luu,lux of every stage
     *   cand_x     [8][t_max+1][X][B]        line-search candidates (next_x of each alpha)
     *   cand_u     [8][t_max][U][B]
     *   cand_cost  [8][B], winner [B], running [B]                                         */
    void* workspace;
    size_t workspace_bytes;
} tplb_batch;

TPLB_API int32_t tplb_abi_version(void);
TPLB_API const tplb_model_info* tplb_model(void);
TPLB_API const char* tplb_last_error(void);

/* Bytes of device scratch a batch of this shape needs. */
TPLB_API size_t tplb_workspace_bytes(int32_t batch, int32_t t_max);
/* Device address of the derivative blocks inside a workspace (for fx..lux views). */
TPLB_API void* tplb_workspace_deriv(void* workspace, int32_t batch, int32_t t_max);
/* Device address of the [8][B] candidate costs of the last line search. */
TPLB_API void* tplb_workspace_cand_cost(void* workspace, int32_t batch, int32_t t_max);

/* One `update()` for every problem of the batch (optim.c:1091-1160): initial rollout
 * and cost, then max_lg_iterations x { multiplier update; up to max_iterations x
 * { linearise (if the trajectory changed); backward Riccati sweep; 8-step line
 * search; regularisation schedule; relative-change stop } }, termination flags. */
TPLB_API int32_t tplb_update(const tplb_batch* batch, void* stream);

/* Derivative blocks of the CURRENT trajectory into workspace.deriv (optim.c:896-912),
 * for every problem regardless of solver state. */
TPLB_API int32_t tplb_linearize(const tplb_batch* batch, void* stream);

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

/* Peak of the FP64 pipe, measured with a register-resident DFMA loop on the current
 * device (the roofline denominator for this path); returns TFLOP/s, <0 on error. */
TPLB_API double tplb_measure_fp64_tflops(int32_t repeats, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* TPLB200_H */
