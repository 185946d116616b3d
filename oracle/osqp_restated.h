/*
 * oracle/osqp_restated.h  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU (FP64, plain C) restatement of the OSQP 0.6.3 ADMM algorithm as the
 * reference drives it through osqp-cpp.  The reference pins
 *   osqp 0.6.3.bcr.2, qdldl 0.1.7.bcr.1, osqp-cpp 0.0.0-20231004-4343373.bcr.1
 * (reference MODULE.bazel:18,21; MODULE.bazel.lock:56-59,83-84); none of
 * those sources are under /root/reference or anywhere in this image, so the
 * algorithm below is restated from the published OSQP 0.6.3 sources
 * (src/osqp.c, src/auxil.c, src/scaling.c, src/lin_alg.c, lin_sys/direct/qdldl)
 * and anchored on the reference's call sites:
 *   solver.Init                               walter_sr/operational_space_controller.h:390,580
 *   UpdateObjectiveAndConstraintMatrices      :565
 *   SetObjectiveVector / SetBounds            :568-569
 *   SetWarmStart                              :583,600
 *   Solve / primal_solution / dual_solution   :591-593
 *
 * PARITY UNPINNED: the reference ships no tests, golden vectors or fixtures
 * for this path and its solver cannot be built or
 * imported here.  The pins this oracle is held to are the ones created in
 * tests/ (KKT certificates, scipy cross-checks, analytic cases, closed-form
 * QP matrices vs. a literal transcription of autogen.py).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may use anything under oracle/.
 *
 * Differences from the OSQP binary, by construction:
 *   - the quasi-definite KKT system is factorised by a dense LDL^T without
 *     the AMD permutation (ORC_LINSYS_KKT) or, for speed, through the
 *     mathematically identical reduced system (P+sigma I+A' diag(rho) A)
 *     with a dense Cholesky (ORC_LINSYS_REDUCED);
 *   - adaptive_rho_interval = 0 resolves to the non-PROFILING rule
 *     (4 x check_termination); the wall-clock rule of PROFILING builds is
 *     non-deterministic and must be requested as an explicit interval;
 *   - no printing, timing, ctrl-c, polishing.
 */
#ifndef ORC_OSQP_RESTATED_H
#define ORC_OSQP_RESTATED_H

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_INFTY 1e30
#define ORC_RHO_MIN 1e-06
#define ORC_RHO_MAX 1e06
#define ORC_RHO_EQ_OVER_RHO_INEQ 1e03
#define ORC_RHO_TOL 1e-04
#define ORC_MIN_SCALING 1e-04
#define ORC_MAX_SCALING 1e+04
#define ORC_ADAPTIVE_RHO_MULTIPLE_TERMINATION 4
#define ORC_ADAPTIVE_RHO_FIXED 100

/* status values follow OSQP's constants.h */
enum {
  ORC_DUAL_INFEASIBLE_INACCURATE = 4,
  ORC_PRIMAL_INFEASIBLE_INACCURATE = 3,
  ORC_SOLVED_INACCURATE = 2,
  ORC_SOLVED = 1,
  ORC_MAX_ITER_REACHED = -2,
  ORC_PRIMAL_INFEASIBLE = -3,
  ORC_DUAL_INFEASIBLE = -4,
  ORC_NON_CVX = -7,
  ORC_UNSOLVED = -10
};

enum { ORC_LINSYS_KKT = 0, ORC_LINSYS_REDUCED = 1 };

typedef struct {
  double rho, sigma, alpha;
  double eps_abs, eps_rel, eps_prim_inf, eps_dual_inf;
  double adaptive_rho_tolerance;
  int scaling;
  int adaptive_rho;
  int adaptive_rho_interval; /* 0 = auto (4 x check_termination) */
  int max_iter;
  int check_termination;
  int warm_start;
  int scaled_termination;
  int linsys; /* ORC_LINSYS_* (oracle-only knob) */
} orc_settings;

typedef struct {
  int n_rows, n_cols, nnz;
  int *p;    /* column pointers, n_cols+1 */
  int *i;    /* row indices */
  double *x; /* values */
} orc_csc;

typedef struct {
  int iter;
  int status;
  double pri_res, dua_res;
  double rho_estimate;
  int rho_updates;
  /* smallest relative distance of any thresholded decision (termination
   * test, rho-update test) from its threshold during the last solve; a
   * parity test may exclude samples whose margin is at round-off level. */
  double decision_margin;
} orc_info;

typedef struct orc_workspace orc_workspace;

void orc_default_settings(orc_settings *s);

/* Dense (column-major) -> CSC, dropping exact zeros like Eigen's sparseView()
 * (reference :558-559).  upper_only keeps i<=j (osqp-cpp takes
 * triangularView<Upper> of the objective matrix). */
orc_csc *orc_csc_from_dense(const double *colmajor, int rows, int cols, int upper_only);
void orc_csc_free(orc_csc *m);
int orc_csc_same_pattern(const orc_csc *a, const orc_csc *b);

/* osqp_setup (osqp-cpp Init): bounds are clipped to +-ORC_INFTY first. */
orc_workspace *orc_setup(const orc_csc *P_upper, const double *q, const orc_csc *A,
                         const double *l, const double *u, const orc_settings *s);
void orc_cleanup(orc_workspace *w);

/* osqp_update_P_A with full value arrays (same pattern required: returns 1 on
 * pattern mismatch without touching the workspace, like osqp-cpp's
 * UpdateObjectiveAndConstraintMatrices). */
int orc_update_P_A(orc_workspace *w, const orc_csc *P_upper, const orc_csc *A);
void orc_update_lin_cost(orc_workspace *w, const double *q_new);
int orc_update_bounds(orc_workspace *w, const double *l_new, const double *u_new);
void orc_warm_start(orc_workspace *w, const double *x, const double *y);
int orc_solve(orc_workspace *w);

const double *orc_solution_x(const orc_workspace *w);
const double *orc_solution_y(const orc_workspace *w);
const orc_info *orc_get_info(const orc_workspace *w);
double orc_get_rho(const orc_workspace *w);
/* scaled iterates and scaling, for state-level parity checks */
void orc_get_scaled_state(const orc_workspace *w, double *x, double *z, double *y,
                          double *D, double *E, double *c);

#ifdef __cplusplus
}
#endif
#endif
