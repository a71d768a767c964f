/* tplb200_prep.h — C ABI of the batched profile shaping that runs right before the lateral and
 * velocity solves of tpl's path/velocity decomposition planner.
 *
 * Replaces, for B independent problems at once:
 *   tplb_rampify_velocity : rampify_profile(v0, a0, lim_v, a_min, a_max, j_min, j_max, v_min, step)
 *                           library/tpl/planning/utils.py:5-65 (called from
 *                           planning/path_vel_decomp/velocity_optim.py:219-224)
 *   tplb_rampify_lateral  : rampify_profile(step, horizon, evasion_sharpness, proj_distance, path,
 *                           gap, lower, upper)
 *                           library/tpl/planning/path_vel_decomp/path_optim.py:11-55 (called twice
 *                           per cycle, path_optim.py:262-278)
 *
 * The reference runs them under numba (fastmath) on one problem; both are sequential scans over
 * the N samples of a profile, so here one thread owns one problem and the batch is the parallel
 * dimension.  All arrays are DEVICE pointers, fp64, structure of arrays with the problem index
 * fastest: sample i of problem b is at [i*B + b].  Calls are asynchronous on `stream`
 * (a cudaStream_t).  Return value 0 or a negative TPLB_PREP_E_* code; tplb_prep_last_error() has the text.
 * One shared library: libtplb200_prep.so (independent of the per-model solver libraries). */
#ifndef TPLB200_PREP_H
#define TPLB200_PREP_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define TPLB_PREP_API __attribute__((visibility("default")))
#else
#define TPLB_PREP_API
#endif

#define TPLB_PREP_ABI_VERSION 2
enum { TPLB_PREP_E_ARG = -1 };

TPLB_PREP_API int32_t tplb_prep_abi_version(void);
TPLB_PREP_API const char* tplb_prep_last_error(void);

/* planning/utils.py:5-65.  lim_v [N][B]; v0, a0 [B] or NULL (= the reference's `None` for every
 * problem: start from the backward pass' value); profile [N][2][B] (velocity, acceleration). */
TPLB_PREP_API int32_t tplb_rampify_velocity(int32_t batch, int32_t n, const double* v0, const double* a0,
                                            const double* lim_v, double a_min, double a_max, double j_min,
                                            double j_max, double v_min, double step, double* profile,
                                            void* stream);

/* path_optim.py:11-55.  path_v = column 5 of the reference's `path` array, lower, upper [N][B];
 * proj_distance [B]; horizon <= N samples are shaped, the remaining samples of the output keep the
 * reference's initial value -10; d_offset [N][B]. */
TPLB_PREP_API int32_t tplb_rampify_lateral(int32_t batch, int32_t n, int32_t horizon, double step,
                                           double evasion_sharpness, const double* proj_distance,
                                           const double* path_v, double gap, const double* lower,
                                           const double* upper, double* d_offset, void* stream);

/* Warm start between planning cycles: resample a trajectory array on its own uniform grid
 * ss = i*step at ss + offset[b] (the arc length travelled since the last cycle).  Replaces
 * VelocityOptim.shift_interp = scipy.interpolate.interp1d(ss, arr, kind, axis=0,
 * fill_value="extrapolate")(ss + arc_len), planning/path_vel_decomp/velocity_optim.py:86-104,
 * applied to opt.x[:-1], opt.u and opt.lagrange_multiplier at :163-168 (scipy 1.x: `linear` =
 * searchsorted + two-point formula with linear extrapolation, `zero` = previous sample, clamped).
 * in, out: [n][rows][B] (the solver's layout for x, u, lambda); must not overlap. */
enum { TPLB_INTERP_LINEAR = 0, TPLB_INTERP_ZERO = 1 };
TPLB_PREP_API int32_t tplb_shift_interp(int32_t batch, int32_t n, int32_t rows, double step,
                                        const double* offset, int32_t kind, const double* in, double* out,
                                        void* stream);

/* Closed-loop simulator step of the ego vehicle for B vehicles at once: SimCore.update_ego
 * (library/tpl/simulation/core.py:91-134) — actuator dead time (the command applied now is the
 * oldest buffered one not older than the dead time), kinematic bicycle with characteristic
 * velocity, explicit Euler, util.normalize_angle (util.py:92-100), clamps.
 * State and command arrays are [B]; the command histories are [capacity][B] with a length per
 * vehicle, capacity >= floor(dead_time / dt) + 2. */
typedef struct {
    int32_t struct_bytes;
    int32_t batch;
    int32_t capacity;              /* rows of the history buffers */
    int32_t reserved0;
    double* x; double* y; double* yaw; double* v; double* a; double* steer_angle;     /* ego.* (in/out) */
    const double* control_acc;     /* ego.control_acc   */
    const double* control_steer;   /* ego.control_steer */
    double acc_dead_time, steer_dead_time, wheel_base, v_ch, max_v, min_v, max_steer_angle;
    double* acc_t; double* acc_value; int32_t* acc_len;            /* self.acc_buffer            */
    double* steer_t; double* steer_value; int32_t* steer_len;      /* self.steering_angle_buffer */
} tplb_ego;

TPLB_PREP_API int32_t tplb_update_ego(const tplb_ego* ego, double t, double dt, void* stream);

/* ---- row f1: reference-path preparation in front of every MPC solve ---------------------------------
 * util.project (library/src/utils.cpp:257-408; control/model_predictive_controller.py:188-189 takes
 * x0[5] = arc_len from it) for B (path, position) pairs.  paths [B][points][stride] rows as the reference
 * holds them (x, y first), position [B][2], out [12][B]: distance, arc_len, alpha, index, start, end,
 * point x, y, tangent x, y, angle, in_bounds (TPLB_PROJ_*). */
enum { TPLB_PROJ_DISTANCE = 0, TPLB_PROJ_ARC_LEN, TPLB_PROJ_ALPHA, TPLB_PROJ_INDEX, TPLB_PROJ_START, TPLB_PROJ_END,
       TPLB_PROJ_POINT_X, TPLB_PROJ_POINT_Y, TPLB_PROJ_TANGENT_X, TPLB_PROJ_TANGENT_Y, TPLB_PROJ_ANGLE,
       TPLB_PROJ_IN_BOUNDS, TPLB_PROJ_FIELDS };
TPLB_PREP_API int32_t tplb_project(int32_t batch, int32_t points, int32_t stride, const double* paths,
                                   const double* position, int32_t closed, double* out, void* stream);

/* util.resample_path = tplcpp.resample (library/src/utils.cpp:410-560) + interp_resampled_path
 * (library/tpl/util.py:134-191; model_predictive_controller.py:124-128, path_optim.py:307).
 * paths [B][points][6] (x, y, orientation, s, curvature, velocity); rs [6][B][steps]: every component is a
 * (B, steps) array that can be bound as a solver parameter (ref_x, ref_y, ...) without a copy;
 * ok [B]: 0 where the reference returns None (resampling failed) or all points coincide — rs is then
 * zero; start_index [B] or NULL (= 0); scratch: tplb_resample_scratch_doubles() doubles. */
TPLB_PREP_API size_t tplb_resample_scratch_doubles(int32_t batch, int32_t points, int32_t steps);
TPLB_PREP_API int32_t tplb_resample_path(int32_t batch, int32_t points, const double* paths, double step_size,
                                         int32_t steps, const int32_t* start_index, int32_t zero_vel_at_end,
                                         int32_t closed, double* rs, int32_t* ok, double* scratch, void* stream);

/* path_optim.py:303-305: the lateral solution back to Cartesian coordinates, in place on paths [B][n][6]:
 * x += -sin(phi) d, y += cos(phi) d, phi += atan(v_d) with d, v_d = states 0, 1 of the solver's
 * buffer xs [t][state_dims][B] (tplb_batch.x). */
TPLB_PREP_API int32_t tplb_frenet_to_cartesian(int32_t batch, int32_t n, int32_t state_dims, double* paths,
                                               const double* xs, void* stream);

#ifdef __cplusplus
}
#endif
#endif
