/*
 * oracle/osc_oracle.c  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 * See osc_oracle.h for the reference lines each function follows.
 */
#include "osc_oracle.h"

#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <stdatomic.h>
#include <unistd.h>

/* ------------------------------------------------------------------ */
/* QP matrices (closed forms of the CasADi functions at q = 0)          */
/* ------------------------------------------------------------------ */
void orc_build_qp(const orc_robot *r, const double *M, const double *C, const double *J,
                  const double *bias, const double *targets, const double *mask, double *H,
                  double *f, double *A, double *l, double *u) {
  const int nv = r->nv, nu = r->nu, nc = r->nc, ns = r->ns;
  const int nz = 3 * nc, n = nv + nu + nz, m = nv + 4 * nc + n, s = 6 * ns;
  const double inf = ORC_INFTY;

  /* per-row weight and target of ddx = J dv + bias
   * (autogen.py:163 split p/r; :173-177 targets cols 0-2 / 3-5; :187-329 weights) */
  double wrow[6 * ORC_MAX_SITES], trow[6 * ORC_MAX_SITES];
  for (int i = 0; i < ns; i++)
    for (int k = 0; k < 3; k++) {
      wrow[3 * i + k] = r->w_trans[i];
      trow[3 * i + k] = targets[6 * i + k];
      wrow[3 * ns + 3 * i + k] = r->w_rot[i];
      trow[3 * ns + 3 * i + k] = targets[6 * i + 3 + k];
    }

  /* H = blkdiag(2 J'WJ, 0, 0) + 2 w_reg I + 2 w_torque I_u ; f = [2 J'W(bias - t); 0; 0] */
  memset(H, 0, sizeof(double) * n * n);
  memset(f, 0, sizeof(double) * n);
  for (int a = 0; a < nv; a++) {
    for (int b = 0; b <= a; b++) {
      double acc = 0.0;
      for (int k = 0; k < s; k++) acc += wrow[k] * J[k * nv + a] * J[k * nv + b];
      acc *= 2.0;
      H[(size_t)b * n + a] = acc;
      H[(size_t)a * n + b] = acc;
    }
    double g = 0.0;
    for (int k = 0; k < s; k++) g += wrow[k] * J[k * nv + a] * (bias[k] - trow[k]);
    f[a] = 2.0 * g;
  }
  for (int j = 0; j < n; j++) H[(size_t)j * n + j] += 2.0 * r->w_reg;
  for (int j = nv; j < nv + nu; j++) H[(size_t)j * n + j] += 2.0 * r->w_torque;

  /* A = [Aeq; Aineq; I] (reference :543-544), column-major m x n */
  memset(A, 0, sizeof(double) * m * n);
  for (int i = 0; i < nv; i++) {
    for (int j = 0; j < nv; j++) A[(size_t)j * m + i] = M[i * nv + j];
    /* -B, B = [0_(nv-nu) x nu ; I_nu]  (autogen.py:54-60) */
    if (i >= nv - nu) A[(size_t)(nv + (i - (nv - nu))) * m + i] = -1.0;
    /* -Jc, Jc = (last nz rows of Jp)'  (reference :497-503) */
    for (int k = 0; k < nz; k++) A[(size_t)(nv + nu + k) * m + i] = -J[(3 * ns - nz + k) * nv + i];
  }
  static const double sx[4] = {1.0, -1.0, 1.0, -1.0};
  static const double sy[4] = {1.0, 1.0, -1.0, -1.0};
  for (int c = 0; c < nc; c++)
    for (int k = 0; k < 4; k++) {
      int row = nv + 4 * c + k, col = nv + nu + 3 * c;
      A[(size_t)(col + 0) * m + row] = sx[k];
      A[(size_t)(col + 1) * m + row] = sy[k];
      A[(size_t)(col + 2) * m + row] = -r->mu;
    }
  for (int j = 0; j < n; j++) A[(size_t)j * m + (nv + 4 * nc + j)] = 1.0;

  /* bounds (reference :546-555, :284-353) */
  int row = 0;
  for (int i = 0; i < nv; i++, row++) l[row] = u[row] = -C[i];
  for (int i = 0; i < 4 * nc; i++, row++) {
    l[row] = -inf;
    u[row] = 0.0;
  }
  for (int i = 0; i < nv; i++, row++) {
    l[row] = -inf;
    u[row] = inf;
  }
  for (int i = 0; i < nu; i++, row++) {
    l[row] = r->u_lb[i];
    u[row] = r->u_ub[i];
  }
  for (int c = 0; c < nc; c++) {
    const double zl[3] = {-inf, -inf, 0.0}, zu[3] = {inf, inf, r->fz_max};
    for (int k = 0; k < 3; k++, row++) {
      l[row] = zl[k] * mask[c];
      u[row] = zu[k] * mask[c];
    }
  }
}

/* ------------------------------------------------------------------ */
/* controller                                                           */
/* ------------------------------------------------------------------ */
struct orc_ctrl {
  orc_robot robot;
  orc_settings settings0;
  orc_workspace *w;
  int n, m;
  double *H, *f, *A, *l, *u;
  double *solution, *dual_solution;
};

orc_ctrl *orc_ctrl_create(const orc_robot *r, const orc_settings *s) {
  orc_ctrl *c = (orc_ctrl *)calloc(1, sizeof(orc_ctrl));
  c->robot = *r;
  c->settings0 = *s;
  c->n = orc_n(r);
  c->m = orc_m(r);
  c->H = (double *)calloc((size_t)c->n * c->n, sizeof(double));
  c->f = (double *)calloc((size_t)c->n, sizeof(double));
  c->A = (double *)calloc((size_t)c->m * c->n, sizeof(double));
  c->l = (double *)calloc((size_t)c->m, sizeof(double));
  c->u = (double *)calloc((size_t)c->m, sizeof(double));
  c->solution = (double *)calloc((size_t)c->n, sizeof(double));
  c->dual_solution = (double *)calloc((size_t)c->m, sizeof(double));
  return c;
}
void orc_ctrl_destroy(orc_ctrl *c) {
  if (!c) return;
  orc_cleanup(c->w);
  free(c->H); free(c->f); free(c->A); free(c->l); free(c->u);
  free(c->solution); free(c->dual_solution);
  free(c);
}
orc_workspace *orc_ctrl_workspace(orc_ctrl *c) { return c->w; }

int orc_ctrl_setup(orc_ctrl *c, const double *M, const double *C, const double *J,
                   const double *bias, const double *targets, const double *mask) {
  orc_build_qp(&c->robot, M, C, J, bias, targets, mask, c->H, c->f, c->A, c->l, c->u);
  orc_csc *P = orc_csc_from_dense(c->H, c->n, c->n, 1);
  orc_csc *A = orc_csc_from_dense(c->A, c->m, c->n, 0);
  orc_cleanup(c->w);
  c->w = orc_setup(P, c->f, A, c->l, c->u, &c->settings0);
  orc_csc_free(P);
  orc_csc_free(A);
  return c->w ? 0 : -1;
}

int orc_ctrl_step(orc_ctrl *c, const double *M, const double *C, const double *J,
                  const double *bias, const double *targets, const double *mask,
                  double *torque, double *x, double *y, orc_info *info) {
  int reinit = 0;
  orc_build_qp(&c->robot, M, C, J, bias, targets, mask, c->H, c->f, c->A, c->l, c->u);
  orc_csc *P = orc_csc_from_dense(c->H, c->n, c->n, 1);
  orc_csc *A = orc_csc_from_dense(c->A, c->m, c->n, 0);
  int rc = c->w ? orc_update_P_A(c->w, P, A) : 1;
  if (rc == 0) {
    orc_update_lin_cost(c->w, c->f);
    orc_update_bounds(c->w, c->l, c->u);
  } else {
    /* sparsity changed: re-Init + SetWarmStart(solution, dual_solution) (:571-584) */
    reinit = 1;
    orc_cleanup(c->w);
    c->w = orc_setup(P, c->f, A, c->l, c->u, &c->settings0);
    if (c->w) orc_warm_start(c->w, c->solution, c->dual_solution);
  }
  orc_csc_free(P);
  orc_csc_free(A);
  if (!c->w) return -1;
  orc_solve(c->w);
  memcpy(c->solution, orc_solution_x(c->w), sizeof(double) * c->n);
  memcpy(c->dual_solution, orc_solution_y(c->w), sizeof(double) * c->m);
  if (torque)
    for (int i = 0; i < c->robot.nu; i++) torque[i] = c->solution[c->robot.nv + i];
  if (x) memcpy(x, c->solution, sizeof(double) * c->n);
  if (y) memcpy(y, c->dual_solution, sizeof(double) * c->m);
  if (info) *info = *orc_get_info(c->w);
  return reinit;
}

void orc_ctrl_reset(orc_ctrl *c) {
  memset(c->solution, 0, sizeof(double) * c->n);
  memset(c->dual_solution, 0, sizeof(double) * c->m);
  if (c->w) orc_warm_start(c->w, c->solution, c->dual_solution);
}

/* ------------------------------------------------------------------ */
/* batch                                                                */
/* ------------------------------------------------------------------ */
struct orc_batch {
  orc_robot robot;
  int n_envs;
  orc_ctrl **c;
};

int orc_max_threads(void) {
  long n = sysconf(_SC_NPROCESSORS_ONLN);
  return n > 0 ? (int)n : 1;
}

/* pthread work queue over environments (chunks of 8) */
typedef struct {
  orc_batch *b;
  int is_setup;
  const double *M, *C, *J, *bias, *targets, *mask;
  double *torque, *x, *y, *pri_res, *dua_res, *rho, *margin;
  int *iters, *status, *rho_updates;
  atomic_int next;
  atomic_int count; /* failed setups or re-Inits */
} orc_job;

static void orc_job_env(orc_job *j, int e);

static void *orc_worker(void *arg) {
  orc_job *j = (orc_job *)arg;
  for (;;) {
    int e0 = atomic_fetch_add(&j->next, 8);
    if (e0 >= j->b->n_envs) break;
    int e1 = e0 + 8 < j->b->n_envs ? e0 + 8 : j->b->n_envs;
    for (int e = e0; e < e1; e++) orc_job_env(j, e);
  }
  return NULL;
}

static void orc_run_job(orc_job *j, int n_threads) {
  int nt = n_threads > 0 ? n_threads : orc_max_threads();
  if (nt > j->b->n_envs) nt = j->b->n_envs;
  if (nt < 1) nt = 1;
  atomic_store(&j->next, 0);
  atomic_store(&j->count, 0);
  if (nt == 1) {
    orc_worker(j);
    return;
  }
  pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * nt);
  for (int t = 0; t < nt; t++) pthread_create(&th[t], NULL, orc_worker, j);
  for (int t = 0; t < nt; t++) pthread_join(th[t], NULL);
  free(th);
}

orc_batch *orc_batch_create(const orc_robot *r, const orc_settings *s, int n_envs) {
  orc_batch *b = (orc_batch *)calloc(1, sizeof(orc_batch));
  b->robot = *r;
  b->n_envs = n_envs;
  b->c = (orc_ctrl **)calloc((size_t)n_envs, sizeof(orc_ctrl *));
  for (int e = 0; e < n_envs; e++) b->c[e] = orc_ctrl_create(r, s);
  return b;
}
void orc_batch_destroy(orc_batch *b) {
  if (!b) return;
  for (int e = 0; e < b->n_envs; e++) orc_ctrl_destroy(b->c[e]);
  free(b->c);
  free(b);
}

static void orc_job_env(orc_job *j, int e) {
  orc_batch *b = j->b;
  const orc_robot *r = &b->robot;
  const int nv = r->nv, nu = r->nu, s = 6 * r->ns, ns = r->ns, nc = r->nc;
  const int n = orc_n(r), m = orc_m(r);
  const double *M = j->M + (size_t)e * nv * nv, *C = j->C + (size_t)e * nv;
  const double *J = j->J + (size_t)e * s * nv, *bias = j->bias + (size_t)e * s;
  const double *targets = j->targets + (size_t)e * ns * 6, *mask = j->mask + (size_t)e * nc;
  if (j->is_setup) {
    if (orc_ctrl_setup(b->c[e], M, C, J, bias, targets, mask)) atomic_fetch_add(&j->count, 1);
    return;
  }
  orc_info info;
  int rc = orc_ctrl_step(b->c[e], M, C, J, bias, targets, mask,
                         j->torque ? j->torque + (size_t)e * nu : NULL,
                         j->x ? j->x + (size_t)e * n : NULL, j->y ? j->y + (size_t)e * m : NULL,
                         &info);
  if (rc > 0) atomic_fetch_add(&j->count, 1);
  if (rc < 0) return;
  if (j->iters) j->iters[e] = info.iter;
  if (j->status) j->status[e] = info.status;
  if (j->pri_res) j->pri_res[e] = info.pri_res;
  if (j->dua_res) j->dua_res[e] = info.dua_res;
  if (j->rho) j->rho[e] = orc_get_rho(b->c[e]->w);
  if (j->rho_updates) j->rho_updates[e] = info.rho_updates;
  if (j->margin) j->margin[e] = info.decision_margin;
}

int orc_batch_setup(orc_batch *b, const double *M, const double *C, const double *J,
                    const double *bias, const double *targets, const double *mask,
                    int n_threads) {
  orc_job j;
  memset(&j, 0, sizeof(j));
  j.b = b; j.is_setup = 1;
  j.M = M; j.C = C; j.J = J; j.bias = bias; j.targets = targets; j.mask = mask;
  orc_run_job(&j, n_threads);
  return atomic_load(&j.count);
}

int orc_batch_step(orc_batch *b, const double *M, const double *C, const double *J,
                   const double *bias, const double *targets, const double *mask, int n_threads,
                   double *torque, double *x, double *y, int *iters, int *status,
                   double *pri_res, double *dua_res, double *rho, int *rho_updates,
                   double *margin) {
  orc_job j;
  memset(&j, 0, sizeof(j));
  j.b = b; j.is_setup = 0;
  j.M = M; j.C = C; j.J = J; j.bias = bias; j.targets = targets; j.mask = mask;
  j.torque = torque; j.x = x; j.y = y; j.iters = iters; j.status = status;
  j.pri_res = pri_res; j.dua_res = dua_res; j.rho = rho; j.rho_updates = rho_updates;
  j.margin = margin;
  orc_run_job(&j, n_threads);
  return atomic_load(&j.count);
}

void orc_batch_scaled_state(orc_batch *b, int e, double *x, double *z, double *y, double *D,
                            double *E, double *c) {
  orc_get_scaled_state(b->c[e]->w, x, z, y, D, E, c);
}
