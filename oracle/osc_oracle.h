/*
 * oracle/osc_oracle.h  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU (FP64) restatement of the reference's per-step control law
 *   OSCData + TaskspaceTargets + contact mask -> QP -> OSQP -> torques
 * following, line by line,
 *   walter_sr/autogen/autogen.py:54-60    (B), :62-93 (dynamics equality),
 *                                :95-133  (friction pyramid), :135-345 (objective)
 *   walter_sr/operational_space_controller.h:284-353 (bounds),
 *                                :355-392 (set_up_optimization),
 *                                :515-539 (update_optimization_data),
 *                                :541-587 (update_optimization),
 *                                :589-594 (solve_optimization), :631 (torque slice)
 *   unitree_go2/... : same code at :285-308, :311-348, :457-481, :483-529, :531-536, :573
 *
 * PARITY UNPINNED (see osqp_restated.h).  The QP matrices are closed forms of
 * what the CasADi-generated functions evaluate at design_vector == 0
 * (reference :278, never written): Aeq=[M,-B,-Jc], beq=-C, Aineq = friction
 * pyramid, bineq=0, H = hessian, f = gradient at 0.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may use anything under oracle/.
 */
#ifndef ORC_OSC_ORACLE_H
#define ORC_OSC_ORACLE_H

#include "osqp_restated.h"

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_MAX_SITES 32
#define ORC_MAX_NU 16

typedef struct {
  int nv, nu, nc, ns;            /* dofs, actuators, contact sites, all sites (contact sites last) */
  double w_trans[ORC_MAX_SITES]; /* weights_config[<site>_translational_tracking] */
  double w_rot[ORC_MAX_SITES];   /* weights_config[<site>_rotational_tracking]    */
  double w_torque, w_reg;        /* weights_config[torque], [regularization]       */
  double mu;                     /* friction_coefficient                           */
  double u_lb[ORC_MAX_NU], u_ub[ORC_MAX_NU]; /* reference :309-320 (W/WW), go2 :285-296 */
  double fz_max;                 /* big_number, reference :282 */
} orc_robot;

/* derived sizes */
static inline int orc_n(const orc_robot *r) { return r->nv + r->nu + 3 * r->nc; }
static inline int orc_m(const orc_robot *r) { return r->nv + 4 * r->nc + orc_n(r); }
static inline int orc_s(const orc_robot *r) { return 6 * r->ns; }

/* Closed-form QP of one environment.  Inputs are laid out exactly like the
 * reference's OSCData / TaskspaceTargets / State.contact_mask (row-major):
 *   M nv*nv, C nv, J (6 ns)*nv [Jp rows of all sites; Jr rows of all sites],
 *   bias 6 ns, targets ns*6, mask nc.
 * Outputs: H n*n col-major, f n, A m*n col-major ([Aeq;Aineq;I]), l m, u m. */
void orc_build_qp(const orc_robot *r, const double *M, const double *C, const double *J,
                  const double *bias, const double *targets, const double *mask, double *H,
                  double *f, double *A, double *l, double *u);

typedef struct orc_ctrl orc_ctrl;

orc_ctrl *orc_ctrl_create(const orc_robot *r, const orc_settings *s);
void orc_ctrl_destroy(orc_ctrl *c);
/* set_up_optimization(): build QP from the data and Init the solver. */
int orc_ctrl_setup(orc_ctrl *c, const double *M, const double *C, const double *J,
                   const double *bias, const double *targets, const double *mask);
/* one control_loop body after update_osc_data(): update_optimization_data,
 * update_optimization (fast path or re-Init + SetWarmStart), solve, slice.
 * Returns 0, or 1 when the pattern changed and the re-Init path was taken. */
int orc_ctrl_step(orc_ctrl *c, const double *M, const double *C, const double *J,
                  const double *bias, const double *targets, const double *mask,
                  double *torque, double *x, double *y, orc_info *info);
void orc_ctrl_reset(orc_ctrl *c); /* reset_optimization(): zero warm start */
orc_workspace *orc_ctrl_workspace(orc_ctrl *c);

/* batch drivers (OpenMP over environments); arrays are [env][...] */
typedef struct orc_batch orc_batch;
orc_batch *orc_batch_create(const orc_robot *r, const orc_settings *s, int n_envs);
void orc_batch_destroy(orc_batch *b);
int orc_batch_setup(orc_batch *b, const double *M, const double *C, const double *J,
                    const double *bias, const double *targets, const double *mask, int n_threads);
int orc_batch_step(orc_batch *b, const double *M, const double *C, const double *J,
                   const double *bias, const double *targets, const double *mask, int n_threads,
                   double *torque, double *x, double *y, int *iters, int *status,
                   double *pri_res, double *dua_res, double *rho, int *rho_updates,
                   double *margin);
/* scaled iterates / scaling of env e after the last step */
void orc_batch_scaled_state(orc_batch *b, int e, double *x, double *z, double *y, double *D,
                            double *E, double *c);
int orc_max_threads(void);

#ifdef __cplusplus
}
#endif
#endif
