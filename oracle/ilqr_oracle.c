/*
 * TEST INFRASTRUCTURE — CPU restatement of the reference iLQR solver.
 *
 * This file is the checker for the CUDA path, never the product: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline leg may load it.
 *
 * It restates, in scalar fp64 C, the algorithm of
 *   /root/reference/library/tpl/optim/templates/optim.c:243-291, 330-408, 636-1177
 * ("optim.c" below) for ONE problem at a time.  The model routines (dynamics,
 * cost and their derivatives) come from a generated header, oracle/models/<name>.h,
 * produced by oracle/gen_models.py with one common-subexpression pass per
 * routine like the reference's own generator.  Compile one shared object per
 * model:  gcc -DTPLO_MODEL_HEADER='"models/<name>.h"' ilqr_oracle.c
 *
 * Pinning: tests/test_oracle_vs_golden.py checks this restatement against the
 * golden vectors in tests/golden/ that were recorded from the real reference
 * build (oracle/_ref, see oracle/build_ref.py and tests/golden/make_golden.py).
 */

#include <math.h>
#include <stdbool.h>
#include <stddef.h>
#include <stdint.h>
#include <string.h>

#define TPLO_MAX_ARRAYS 16
#define TPLO_MAX_SCALARS 64

typedef struct {
    double scalar[TPLO_MAX_SCALARS];
    const double* array[TPLO_MAX_ARRAYS];
    int64_t length[TPLO_MAX_ARRAYS];        /* dims[0] */
    int64_t cols[TPLO_MAX_ARRAYS];          /* dims[1] of a 2-D array, 0 otherwise */
} tplo_params;

/* ------------------------------------------------------------------ helpers used by generated code */

static inline double sq(double a) { return a * a; }
static inline double pow3h(double a) { return a * sqrt(a); }
static inline double ipow3(double a) { return a * a * a; }
static inline double ipow4(double a) { double b = a * a; return b * b; }
static inline double ipow5(double a) { double b = a * a; return b * b * a; }
static inline double ipow6(double a) { double b = a * a * a; return b * b; }
static inline double ipow7(double a) { double b = a * a * a; return b * b * a; }
static inline double ipow8(double a) { double b = a * a; b = b * b; return b * b; }

/* optim.c:347-355.  `(size_t)floor(q)` of a negative q is, on x86-64, a huge
 * unsigned number, so min(size-1, .) selects the LAST sample; (size_t)ceil(q)
 * of q in (-1,0) is 0.  The blend weight clamps to [0,1] afterwards. */
typedef struct { double w_hi, w_lo; size_t lo, hi; } tplo_cell;

static size_t tplo_index(double v, size_t n) {
    if (!(v >= 0.0)) return n - 1;          /* negative (or NaN) wraps to the last sample */
    if (v >= (double)n) return n - 1;
    return (size_t)v;
}

static tplo_cell tplo_locate(double x0, double dx, double x, size_t n) {
    tplo_cell c;
    const double q = (x - x0) / dx;
    c.lo = tplo_index(floor(q), n);
    c.hi = tplo_index(ceil(q), n);
    double w = q - (double)c.lo;
    w = (0.0 > w) ? 0.0 : w;
    w = (w < 1.0) ? w : 1.0;
    c.w_hi = w;
    c.w_lo = 1.0 - w;
    return c;
}

/* optim.c:374-388 */
static double tplo_lerp(const tplo_params* P, int a, double x0, double dx, double x) {
    const size_t n = (size_t)P->length[a];
    if (n == 0) return 0.0;
    const tplo_cell c = tplo_locate(x0, dx, x, n);
    return c.w_lo * P->array[a][c.lo] + c.w_hi * P->array[a][c.hi];
}

/* optim.c:332-338 */
static double tplo_short_angle(double from, double to) {
    const double turn = M_PI * 2;
    const double d = fmod(to - from, turn);
    return fmod(2 * d, turn) - d;
}

/* optim.c:392-406 */
static double tplo_lerp_angle(const tplo_params* P, int a, double x0, double dx, double x) {
    const size_t n = (size_t)P->length[a];
    if (n == 0) return 0.0;
    const tplo_cell c = tplo_locate(x0, dx, x, n);
    const double lo = P->array[a][c.lo];
    return lo + tplo_short_angle(lo, P->array[a][c.hi]) * c.w_hi;
}

/* optim.c:357-370 */
static double tplo_box_interp(const tplo_params* P, int a, double dx, double x) {
    const size_t n = (size_t)P->length[a];
    if (n == 0) return 0.0;
    return P->array[a][tplo_index(floor(x / dx), n)];
}

/* optim.c:330 */
static double tplo_array_value(const tplo_params* P, int a, double i) {
    return P->array[a][(size_t)i];
}

/* optim.c:410-448.  The reference leaves the indices unclamped: with gap <= 0 a position at or
 * beyond the last sample reads past the array.  The restatement clamps to the last sample there. */
static double tplo_lerp_wrap(const tplo_params* P, int axs, int a, double len, double dx, double x) {
    const size_t n = (size_t)P->length[a];
    if (n == 0) return 0.0;
    const double first = P->array[axs][0];
    const double last = first + (double)(n - 1) * dx;
    const double gap = len - (last - first);
    x = fmod(x - first, len);
    if (x < 0) x += len;
    x += first;
    double w;
    size_t lo, hi;
    if (x >= last && gap > 0) {
        w = (x - last) / gap;
        lo = n - 1;
        hi = 0;
    } else {
        const double q = (x - first) / dx;
        lo = tplo_index(floor(q), n);
        hi = tplo_index(ceil(q), n);
        w = q - (double)lo;
    }
    return (1.0 - w) * P->array[a][lo] + w * P->array[a][hi];
}

/* optim.c:457-481: rows = dims[0], cols = dims[1], row-major */
static double tplo_blerp(const tplo_params* P, int a, double x0, double y0, double dx, double dy,
                         double x, double y) {
    const size_t rows = (size_t)P->length[a], cols = (size_t)P->cols[a];
    if (rows == 0 || cols == 0) return 0.0;      /* the reference would index out of bounds */
    const tplo_cell cx = tplo_locate(x0, dx, x, cols);
    const tplo_cell cy = tplo_locate(y0, dy, y, rows);
    const double* v = P->array[a];
    const double p0 = cy.w_lo * v[cy.lo * cols + cx.lo] + cy.w_hi * v[cy.hi * cols + cx.lo];
    const double p1 = cy.w_lo * v[cy.lo * cols + cx.hi] + cy.w_hi * v[cy.hi * cols + cx.hi];
    return cx.w_lo * p0 + cx.w_hi * p1;
}

#ifndef TPLO_MODEL_HEADER
#error "compile with -DTPLO_MODEL_HEADER='\"models/<name>.h\"'"
#endif
#include TPLO_MODEL_HEADER

enum { NX = TPLO_X, NU = TPLO_U, NC = TPLO_C };
#define NCs (NC > 0 ? NC : 1)

/* ------------------------------------------------------------------ problem record (ctypes mirror in oracle.py) */

typedef struct {
    /* settings (optim.c:600-620) */
    int32_t T, opt_start, max_iterations, max_lg_iterations, integrator, use_quadratic_terms;
    double dt, min_rel_cost_change;
    /* solver status, sticky across calls (optim.c:563-594) */
    double traj_costs, alpha, mu;
    int32_t iterations, lg_iterations, mu_step, trajectory_changed, improved, termination_condition;
    /* trajectories, row-major [stage][component] */
    double *x, *u, *next_x, *next_u, *prev_x, *prev_k;
    /* derivative blocks and gains */
    double *fx, *fu, *lx, *lu, *lxx, *luu, *lux, *g, *k, *K;
    /* multipliers and limits */
    double *lam, *barrier_weight, *lg_mult_limit, *u_min, *u_max;
    tplo_params params;
} tplo_problem;

int tplo_dims(int* out) {
    out[0] = NX; out[1] = NU; out[2] = NC; out[3] = TPLO_NUM_SCALARS; out[4] = TPLO_NUM_ARRAYS;
    return (int)sizeof(tplo_problem);
}

/* ------------------------------------------------------------------ integrators (optim.c:657-730) */

static void step_state(const tplo_problem* p, const double* x, const double* u,
                       int t, double h, int scheme, double* out) {
    double k1[NX], k2[NX], k3[NX], k4[NX], y[NX];
    const double tt = (double)t;
    ctDynamics(&p->params, x, u, tt, h, k1);
    if (scheme == 0) {                                   /* explicit Euler */
        for (int i = 0; i < NX; ++i) out[i] = x[i] + k1[i] * h;
    } else if (scheme == 1) {                            /* Heun */
        for (int i = 0; i < NX; ++i) y[i] = x[i] + k1[i] * h;
        ctDynamics(&p->params, y, u, tt, h, k2);
        for (int i = 0; i < NX; ++i) out[i] = x[i] + (k1[i] + k2[i]) * (h / 2.0);
    } else {                                             /* classic RK4 */
        for (int i = 0; i < NX; ++i) y[i] = x[i] + k1[i] * (h / 2.0);
        ctDynamics(&p->params, y, u, tt, h, k2);
        for (int i = 0; i < NX; ++i) y[i] = x[i] + k2[i] * (h / 2.0);
        ctDynamics(&p->params, y, u, tt, h, k3);
        for (int i = 0; i < NX; ++i) y[i] = x[i] + k3[i] * h;
        ctDynamics(&p->params, y, u, tt, h, k4);
        for (int i = 0; i < NX; ++i) {
            /* optim.c:717-724: ((k4 + (2 k3 + (2 k2 + (k1 + 0)))) * h/6 */
            double acc = k1[i] + 0.0;
            acc = k2[i] * 2.0 + acc;
            acc = k3[i] * 2.0 + acc;
            acc = k4[i] + acc;
            out[i] = x[i] + acc * (h / 6.0);
        }
    }
}

void tplo_dynamics(const tplo_problem* p, const double* x, const double* u, int t, double h, double* out) {
    step_state(p, x, u, t, h, p->integrator, out);
}

void tplo_ct_dynamics(const tplo_problem* p, const double* x, const double* u, int t, double h, double* out) {
    ctDynamics(&p->params, x, u, (double)t, h, out);
}

/* ------------------------------------------------------------------ derivative blocks (optim.c:896-912) */

void tplo_linearize(tplo_problem* p) {
    const double h = p->dt;
    for (int t = 0; t < p->T; ++t) {
        const double* x = p->x + (size_t)t * NX;
        const double* u = p->u + (size_t)t * NU;
        const double* lam = p->lam + (size_t)t * NC;
        const double tt = (double)t;
        stateJacobian(&p->params, x, u, tt, h, p->fx + (size_t)t * NX * NX);
        actionJacobian(&p->params, x, u, tt, h, p->fu + (size_t)t * NX * NU);
        stateGradient(&p->params, x, u, lam, p->barrier_weight, tt, h, p->lx + (size_t)t * NX);
        actionGradient(&p->params, x, u, lam, p->barrier_weight, tt, h, p->lu + (size_t)t * NU);
        if (p->use_quadratic_terms) {
            stateStateHessian(&p->params, x, u, lam, p->barrier_weight, tt, h, p->lxx + (size_t)t * NX * NX);
            actionActionHessian(&p->params, x, u, lam, p->barrier_weight, tt, h, p->luu + (size_t)t * NU * NU);
            actionStateHessian(&p->params, x, u, lam, p->barrier_weight, tt, h, p->lux + (size_t)t * NU * NX);
        }
    }
}

/* ------------------------------------------------------------------ gains (optim.c:243-291) */

static void control_gains(double Quu[NU][NU], const double Qu[NU], double Qux[NU][NX],
                          double mu, double* k, double* K) {
#if TPLO_U == 1
    double s = 0.0;
    if (Quu[0][0] > 0.0) s = -1.0 / (Quu[0][0] + mu);     /* test on the un-regularised value */
    k[0] = Qu[0] * s;
    for (int j = 0; j < NX; ++j) K[j] = Qux[0][j] * s;
#elif TPLO_U == 2
    const double a = Quu[0][0] + mu, b = Quu[0][1], d = Quu[1][1] + mu;
    const double det = a * d - b * b;
    const double s = -1.0 / det;                           /* no definiteness check */
    double M[2][2];
    M[0][0] = d * s;
    M[0][1] = -b * s;
    M[1][0] = M[0][1];
    M[1][1] = a * s;
    for (int i = 0; i < 2; ++i) {
        double acc = 0.0;
        for (int c = 0; c < 2; ++c) acc += M[i][c] * Qu[c];
        k[i] = acc;
        for (int j = 0; j < NX; ++j) {
            double r = 0.0;
            for (int c = 0; c < 2; ++c) r += M[i][c] * Qux[c][j];
            K[i * NX + j] = r;
        }
    }
#else
#error "more than two controls are not supported (genopt.py:420-425)"
#endif
}

/* ------------------------------------------------------------------ backward Riccati sweep (optim.c:914-985) */

static void backward_sweep(tplo_problem* p) {
    double Vx[NX], Vxx[NX][NX];
    const int T = p->T;
    endGradient(&p->params, p->x + (size_t)T * NX, (double)T, p->dt, Vx);
    endHessian(&p->params, p->x + (size_t)T * NX, (double)T, p->dt, &Vxx[0][0]);

    for (int t = T - 1; t >= 0; --t) {
        const double (*A)[NX] = (const double (*)[NX])(p->fx + (size_t)t * NX * NX);
        const double (*B)[NU] = (const double (*)[NU])(p->fu + (size_t)t * NX * NU);
        const double* lx = p->lx + (size_t)t * NX;
        const double* lu = p->lu + (size_t)t * NU;
        const double (*lxx)[NX] = (const double (*)[NX])(p->lxx + (size_t)t * NX * NX);
        const double (*luu)[NU] = (const double (*)[NU])(p->luu + (size_t)t * NU * NU);
        const double (*lux)[NX] = (const double (*)[NX])(p->lux + (size_t)t * NU * NX);
        double* k = p->k + (size_t)t * NU;
        double* K = p->K + (size_t)t * NU * NX;
        const double* u = p->u + (size_t)t * NU;

        double Qx[NX], Qu[NU], Qxx[NX][NX], Quu[NU][NU], Qux[NU][NX];
        double VA[NX][NX], VB[NX][NU];

        for (int i = 0; i < NX; ++i) {                       /* Qx = lx + A' Vx */
            double acc = 0.0;
            for (int r = 0; r < NX; ++r) acc += A[r][i] * Vx[r];
            Qx[i] = lx[i] + acc;
        }
        for (int i = 0; i < NU; ++i) {                       /* Qu = lu + B' Vx */
            double acc = 0.0;
            for (int r = 0; r < NX; ++r) acc += B[r][i] * Vx[r];
            Qu[i] = lu[i] + acc;
        }
        for (int i = 0; i < NX; ++i) {                       /* VA = Vxx A, VB = Vxx B */
            for (int j = 0; j < NX; ++j) {
                double acc = 0.0;
                for (int r = 0; r < NX; ++r) acc += Vxx[i][r] * A[r][j];
                VA[i][j] = acc;
            }
            for (int j = 0; j < NU; ++j) {
                double acc = 0.0;
                for (int r = 0; r < NX; ++r) acc += Vxx[i][r] * B[r][j];
                VB[i][j] = acc;
            }
        }
        for (int i = 0; i < NX; ++i)                         /* Qxx = lxx + A' VA, lower triangle mirrored */
            for (int j = 0; j <= i; ++j) {
                double acc = 0.0;
                for (int r = 0; r < NX; ++r) acc += A[r][i] * VA[r][j];
                Qxx[i][j] = acc;
                Qxx[j][i] = acc;
            }
        for (int i = 0; i < NX; ++i)
            for (int j = 0; j < NX; ++j) Qxx[i][j] = lxx[i][j] + Qxx[i][j];
        for (int i = 0; i < NU; ++i)                         /* Quu = luu + B' VB, lower triangle mirrored */
            for (int j = 0; j <= i; ++j) {
                double acc = 0.0;
                for (int r = 0; r < NX; ++r) acc += B[r][i] * VB[r][j];
                Quu[i][j] = acc;
                Quu[j][i] = acc;
            }
        for (int i = 0; i < NU; ++i)
            for (int j = 0; j < NU; ++j) Quu[i][j] = luu[i][j] + Quu[i][j];
        for (int i = 0; i < NU; ++i)                         /* Qux = lux + B' VA */
            for (int j = 0; j < NX; ++j) {
                double acc = 0.0;
                for (int r = 0; r < NX; ++r) acc += B[r][i] * VA[r][j];
                Qux[i][j] = lux[i][j] + acc;
            }

        control_gains(Quu, Qu, Qux, p->mu, k, K);

        /* box limits on the feed-forward step (optim.c:950-963) */
        for (int d = 0; d < NU; ++d) {
            const double cand = u[d] + k[d];
            const double hi = p->u_max[(size_t)t * NU + d], lo = p->u_min[(size_t)t * NU + d];
            if (cand > hi) {
                k[d] = hi - u[d];
                for (int j = 0; j < NX; ++j) K[d * NX + j] = 0.0;
            }
            if (cand < lo) {
                k[d] = lo - u[d];
                for (int j = 0; j < NX; ++j) K[d * NX + j] = 0.0;
            }
        }

        /* value function update (optim.c:965-984) */
        double KtQux[NX][NX], KtQuu[NX][NU];
        for (int i = 0; i < NX; ++i) {
            for (int j = 0; j < NX; ++j) {
                double acc = 0.0;
                for (int c = 0; c < NU; ++c) acc += K[c * NX + i] * Qux[c][j];
                KtQux[i][j] = acc;
            }
            for (int j = 0; j < NU; ++j) {
                double acc = 0.0;
                for (int c = 0; c < NU; ++c) acc += K[c * NX + i] * Quu[c][j];
                KtQuu[i][j] = acc;
            }
        }
        for (int i = 0; i < NX; ++i)
            for (int j = 0; j < NX; ++j) {
                double v = KtQux[j][i] + KtQux[i][j];
                for (int c = 0; c < NU; ++c) v += KtQuu[i][c] * K[c * NX + j];
                Vxx[i][j] = v + Qxx[i][j];
            }
        for (int i = 0; i < NX; ++i) {
            double v = 0.0;
            for (int c = 0; c < NU; ++c) v += KtQuu[i][c] * k[c];
            for (int c = 0; c < NU; ++c) v += K[c * NX + i] * Qu[c];
            for (int c = 0; c < NU; ++c) v += Qux[c][i] * k[c];
            Vx[i] = v + Qx[i];
        }
    }
}

/* gradient-only sweep (optim.c:1038-1076) */
static void backward_sweep_first_order(tplo_problem* p) {
    double Vx[NX];
    const int T = p->T;
    endGradient(&p->params, p->x + (size_t)T * NX, (double)T, p->dt, Vx);
    for (int t = T - 1; t >= 0; --t) {
        const double (*A)[NX] = (const double (*)[NX])(p->fx + (size_t)t * NX * NX);
        const double (*B)[NU] = (const double (*)[NU])(p->fu + (size_t)t * NX * NU);
        double Qx[NX];
        for (int i = 0; i < NX; ++i) {
            double acc = 0.0;
            for (int r = 0; r < NX; ++r) acc += A[r][i] * Vx[r];
            Qx[i] = p->lx[(size_t)t * NX + i] + acc;
        }
        for (int i = 0; i < NU; ++i) {
            double acc = 0.0;
            for (int r = 0; r < NX; ++r) acc += B[r][i] * Vx[r];
            const double q = p->lu[(size_t)t * NU + i] + acc;
            p->g[(size_t)t * NU + i] = q;
            p->k[(size_t)t * NU + i] = q;
        }
        memcpy(Vx, Qx, sizeof Qx);
    }
    for (int t = T - 1; t >= 0; --t)
        for (int d = 0; d < NU; ++d) {
            const size_t i = (size_t)t * NU + d;
            const double cand = p->u[i] - p->g[i];
            if (cand > p->u_max[i]) p->k[i] = p->u[i] - p->u_max[i];
            if (cand < p->u_min[i]) p->k[i] = p->u[i] - p->u_min[i];
        }
}

/* ------------------------------------------------------------------ rollouts (optim.c:732-838) */

static double rollout(tplo_problem* p, double step, int second_order) {
    const int T = p->T;
    double total = 0.0, c = 0.0;
    memcpy(p->next_x, p->x, sizeof(double) * NX);
    for (int t = p->opt_start; t < T; ++t) {
        double* xn = p->next_x + (size_t)t * NX;
        double* un = p->next_u + (size_t)t * NU;
        const double* k = p->k + (size_t)t * NU;
        for (int d = 0; d < NU; ++d) {
            if (second_order) {
                double v = k[d] * step + p->u[(size_t)t * NU + d];
                for (int j = 0; j < NX; ++j)
                    v += p->K[((size_t)t * NU + d) * NX + j] * (xn[j] - p->x[(size_t)t * NX + j]);
                const double hi = p->u_max[(size_t)t * NU + d], lo = p->u_min[(size_t)t * NU + d];
                const double capped = (hi < v) ? hi : v;
                un[d] = (lo > capped) ? lo : capped;
            } else {
                un[d] = p->u[(size_t)t * NU + d] - k[d] * step;    /* no clamp here (optim.c:803-804) */
            }
        }
        step_state(p, xn, un, t, p->dt, p->integrator, xn + NX);
        costs(&p->params, xn, un, p->lam + (size_t)t * NC, p->barrier_weight, (double)t, p->dt, &c);
        total += c;
    }
    endCosts(&p->params, p->next_x + (size_t)T * NX, (double)T, p->dt, &c);
    total += c;
    return total;
}

/* optim.c:840-873 */
static int line_search(tplo_problem* p, int second_order) {
    static const double tens[8] = {1.0, 10.0, 100.0, 1000.0, 1e4, 1e5, 1e6, 1e7};
    const int T = p->T;
    for (int i = 0; i < 8; ++i) {
        p->alpha = 1.0 / tens[i];
        const double cand = rollout(p, p->alpha, second_order);
        if (cand < p->traj_costs && isfinite(cand) && cand >= 0.0) {
            memcpy(p->prev_x, p->x, sizeof(double) * (size_t)(T + 1) * NX);
            memcpy(p->prev_k, p->k, sizeof(double) * (size_t)T * NU);
            memcpy(p->x, p->next_x, sizeof(double) * (size_t)(T + 1) * NX);
            memcpy(p->u, p->next_u, sizeof(double) * (size_t)T * NU);
            p->traj_costs = cand;
            p->trajectory_changed = 1;
            p->improved = 1;
            return 1;
        }
    }
    return 0;
}

/* ------------------------------------------------------------------ inner loops (optim.c:875-1089) */

static void inner_iterations(tplo_problem* p) {
    static const double decades[8] = {0.0, 1.0, 10.0, 100.0, 1000.0, 1e4, 1e5, 1e6};
    for (int s = p->iterations; s < p->max_iterations; ++s) {
        p->iterations = s + 1;
        if (p->trajectory_changed) {
            tplo_linearize(p);
            p->trajectory_changed = 0;
        }
        const double before = p->traj_costs;
        if (p->use_quadratic_terms) {
            backward_sweep(p);
            if (line_search(p, 1)) p->mu_step = (p->mu_step - 1 > 0) ? p->mu_step - 1 : 0;
            else                   p->mu_step = (p->mu_step + 1 < 7) ? p->mu_step + 1 : 7;
            p->mu = decades[p->mu_step];                   /* 0 or 10^(mu_step-1) */
        } else {
            backward_sweep_first_order(p);
            line_search(p, 0);
        }
        const double rel = fabs(p->traj_costs - before) / p->traj_costs;
        if (rel < p->min_rel_cost_change) {
            p->termination_condition = 2;
            break;
        }
    }
}

/* ------------------------------------------------------------------ update (optim.c:1091-1160) */

void tplo_update(tplo_problem* p) {
    const int T = p->T;
    double c = 0.0;
    p->traj_costs = 0.0;
    for (int t = p->opt_start; t < T; ++t) {
        step_state(p, p->x + (size_t)t * NX, p->u + (size_t)t * NU, t, p->dt, p->integrator,
                   p->x + (size_t)(t + 1) * NX);
        costs(&p->params, p->x + (size_t)t * NX, p->u + (size_t)t * NU, p->lam + (size_t)t * NC,
              p->barrier_weight, (double)t, p->dt, &c);
        p->traj_costs += c;
    }
    endCosts(&p->params, p->x + (size_t)T * NX, (double)T, p->dt, &c);
    p->traj_costs += c;

    for (p->lg_iterations = 0; p->lg_iterations < p->max_lg_iterations; ++p->lg_iterations) {
        double g[NCs];
        for (int t = p->opt_start; t < T; ++t) {
            for (int i = 0; i < NC; ++i) g[i] = 0.0;
            constraints(&p->params, p->x + (size_t)t * NX, p->u + (size_t)t * NU, p->lam + (size_t)t * NC,
                        p->barrier_weight, (double)t, p->dt, g);
            for (int i = 0; i < NC; ++i) {
                double v = p->lam[(size_t)t * NC + i] + p->barrier_weight[i] * g[i];
                v = (0.0 > v) ? 0.0 : v;
                p->lam[(size_t)t * NC + i] = (p->lg_mult_limit[i] < v) ? p->lg_mult_limit[i] : v;
            }
        }
        p->trajectory_changed = 1;
        p->improved = 0;
        p->iterations = 0;
        inner_iterations(p);
    }
    if (p->iterations == p->max_iterations) p->termination_condition = 1;
}

/* ------------------------------------------------------------------ warm-start shift (optim.c:1162-1177) */

void tplo_shift(tplo_problem* p, int amount) {
    const int T = p->T;
    if (amount < 0) amount = 0;
    for (int t = p->opt_start; t < T + 1; ++t) {
        const int s = (t + amount < T) ? t + amount : T;
        memmove(p->x + (size_t)t * NX, p->x + (size_t)s * NX, sizeof(double) * NX);
    }
    for (int t = p->opt_start; t < T; ++t) {
        const int s = (t + amount < T - 1) ? t + amount : T - 1;
        memmove(p->u + (size_t)t * NU, p->u + (size_t)s * NU, sizeof(double) * NU);
        if (NC > 0) memmove(p->lam + (size_t)t * NC, p->lam + (size_t)s * NC, sizeof(double) * NC);
    }
}

/* stage cost / constraint probes for tests */
double tplo_stage_cost(const tplo_problem* p, const double* x, const double* u, const double* lam, int t) {
    double c = 0.0;
    costs(&p->params, x, u, lam, p->barrier_weight, (double)t, p->dt, &c);
    return c;
}
