/*
 * oracle/osqp_restated.c  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 * See osqp_restated.h for provenance ("parity unpinned") and scope.
 *
 * Function names in comments are the OSQP 0.6.3 functions each block restates.
 */
#include "osqp_restated.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#define c_max(a, b) (((a) > (b)) ? (a) : (b))
#define c_min(a, b) (((a) < (b)) ? (a) : (b))
#define c_absval(x) (((x) < 0) ? -(x) : (x))

struct orc_workspace {
  int n, m;
  orc_settings settings;
  /* data (scaled in place, like OSQP's work->data) */
  orc_csc *P; /* upper triangle */
  orc_csc *A;
  double *q, *l, *u;
  /* scaling */
  double c, cinv;
  double *D, *Dinv, *E, *Einv;
  double *D_temp, *D_temp_A, *E_temp;
  /* rho */
  double *rho_vec, *rho_inv_vec;
  int *constr_type;
  /* iterates */
  double *x, *y, *z, *xz_tilde, *x_prev, *z_prev;
  double *Ax, *Px, *Aty, *delta_y, *Atdelta_y, *delta_x, *Pdelta_x, *Adelta_x;
  /* solution */
  double *sol_x, *sol_y;
  orc_info info;
  /* linear system */
  int ls_dim;      /* n+m (KKT) or n (reduced) */
  double *ls_mat;  /* dense factor storage, row-major lower */
  double *ls_diag; /* D of LDL^T (KKT mode) */
  double *ls_work;
  /* CSR copy of A for the reduced form */
  int *Ar_p, *Ar_j;
  double *Ar_x;
};

/* ------------------------------------------------------------------ */
/* lin_alg.c                                                            */
/* ------------------------------------------------------------------ */
static double vec_norm_inf(const double *v, int l) {
  double max = 0.0, a;
  for (int i = 0; i < l; i++) {
    a = c_absval(v[i]);
    if (a > max) max = a;
  }
  return max;
}
static double vec_scaled_norm_inf(const double *S, const double *v, int l) {
  double max = 0.0, a;
  for (int i = 0; i < l; i++) {
    a = c_absval(S[i] * v[i]);
    if (a > max) max = a;
  }
  return max;
}
static double vec_mean(const double *a, int n) {
  double mean = 0.0;
  for (int i = 0; i < n; i++) mean += a[i];
  return mean / (double)n;
}
static double vec_prod(const double *a, const double *b, int n) {
  double prod = 0.0;
  for (int i = 0; i < n; i++) prod += a[i] * b[i];
  return prod;
}

/* y (+)= A x, CSC traversal (mat_vec) */
static void mat_vec(const orc_csc *A, const double *x, double *y, int plus_eq) {
  if (!plus_eq)
    for (int i = 0; i < A->n_rows; i++) y[i] = 0;
  if (A->p[A->n_cols] == 0) return;
  if (plus_eq == -1) {
    for (int j = 0; j < A->n_cols; j++)
      for (int i = A->p[j]; i < A->p[j + 1]; i++) y[A->i[i]] -= A->x[i] * x[j];
  } else {
    for (int j = 0; j < A->n_cols; j++)
      for (int i = A->p[j]; i < A->p[j + 1]; i++) y[A->i[i]] += A->x[i] * x[j];
  }
}
/* y (+)= A' x (mat_tpose_vec) */
static void mat_tpose_vec(const orc_csc *A, const double *x, double *y, int plus_eq,
                          int skip_diag) {
  if (!plus_eq)
    for (int i = 0; i < A->n_cols; i++) y[i] = 0;
  if (A->p[A->n_cols] == 0) return;
  for (int j = 0; j < A->n_cols; j++) {
    for (int k = A->p[j]; k < A->p[j + 1]; k++) {
      int i = A->i[k];
      if (skip_diag && i == j) continue;
      if (plus_eq == -1)
        y[j] -= A->x[k] * x[i];
      else
        y[j] += A->x[k] * x[i];
    }
  }
}
static void mat_premult_diag(orc_csc *A, const double *d) {
  for (int j = 0; j < A->n_cols; j++)
    for (int i = A->p[j]; i < A->p[j + 1]; i++) A->x[i] *= d[A->i[i]];
}
static void mat_postmult_diag(orc_csc *A, const double *d) {
  for (int j = 0; j < A->n_cols; j++)
    for (int i = A->p[j]; i < A->p[j + 1]; i++) A->x[i] *= d[j];
}
static void mat_mult_scalar(orc_csc *A, double sc) {
  int nnzA = A->p[A->n_cols];
  for (int i = 0; i < nnzA; i++) A->x[i] *= sc;
}
static void mat_inf_norm_cols(const orc_csc *M, double *E) {
  for (int j = 0; j < M->n_cols; j++) E[j] = 0.;
  for (int j = 0; j < M->n_cols; j++)
    for (int ptr = M->p[j]; ptr < M->p[j + 1]; ptr++)
      E[j] = c_max(c_absval(M->x[ptr]), E[j]);
}
static void mat_inf_norm_rows(const orc_csc *M, double *E) {
  for (int j = 0; j < M->n_rows; j++) E[j] = 0.;
  for (int j = 0; j < M->n_cols; j++)
    for (int ptr = M->p[j]; ptr < M->p[j + 1]; ptr++) {
      int i = M->i[ptr];
      E[i] = c_max(c_absval(M->x[ptr]), E[i]);
    }
}
static void mat_inf_norm_cols_sym_triu(const orc_csc *M, double *E) {
  for (int j = 0; j < M->n_cols; j++) E[j] = 0.;
  for (int j = 0; j < M->n_cols; j++)
    for (int ptr = M->p[j]; ptr < M->p[j + 1]; ptr++) {
      int i = M->i[ptr];
      double abs_x = c_absval(M->x[ptr]);
      E[j] = c_max(abs_x, E[j]);
      if (i != j) E[i] = c_max(abs_x, E[i]);
    }
}

/* ------------------------------------------------------------------ */
/* CSC helpers                                                          */
/* ------------------------------------------------------------------ */
orc_csc *orc_csc_from_dense(const double *cm, int rows, int cols, int upper_only) {
  int nnz = 0;
  for (int j = 0; j < cols; j++)
    for (int i = 0; i < rows; i++) {
      if (upper_only && i > j) continue;
      if (cm[(size_t)j * rows + i] != 0.0) nnz++;
    }
  orc_csc *m = (orc_csc *)calloc(1, sizeof(orc_csc));
  m->n_rows = rows;
  m->n_cols = cols;
  m->nnz = nnz;
  m->p = (int *)malloc(sizeof(int) * (cols + 1));
  m->i = (int *)malloc(sizeof(int) * (nnz > 0 ? nnz : 1));
  m->x = (double *)malloc(sizeof(double) * (nnz > 0 ? nnz : 1));
  int k = 0;
  for (int j = 0; j < cols; j++) {
    m->p[j] = k;
    for (int i = 0; i < rows; i++) {
      if (upper_only && i > j) continue;
      double v = cm[(size_t)j * rows + i];
      if (v != 0.0) {
        m->i[k] = i;
        m->x[k] = v;
        k++;
      }
    }
  }
  m->p[cols] = k;
  return m;
}
static orc_csc *csc_copy(const orc_csc *a) {
  orc_csc *m = (orc_csc *)calloc(1, sizeof(orc_csc));
  *m = *a;
  int nnz = a->p[a->n_cols];
  m->p = (int *)malloc(sizeof(int) * (a->n_cols + 1));
  m->i = (int *)malloc(sizeof(int) * (nnz > 0 ? nnz : 1));
  m->x = (double *)malloc(sizeof(double) * (nnz > 0 ? nnz : 1));
  memcpy(m->p, a->p, sizeof(int) * (a->n_cols + 1));
  memcpy(m->i, a->i, sizeof(int) * nnz);
  memcpy(m->x, a->x, sizeof(double) * nnz);
  return m;
}
void orc_csc_free(orc_csc *m) {
  if (!m) return;
  free(m->p);
  free(m->i);
  free(m->x);
  free(m);
}
int orc_csc_same_pattern(const orc_csc *a, const orc_csc *b) {
  if (a->n_rows != b->n_rows || a->n_cols != b->n_cols) return 0;
  int nnz = a->p[a->n_cols];
  if (nnz != b->p[b->n_cols]) return 0;
  if (memcmp(a->p, b->p, sizeof(int) * (a->n_cols + 1))) return 0;
  if (memcmp(a->i, b->i, sizeof(int) * nnz)) return 0;
  return 1;
}

void orc_default_settings(orc_settings *s) {
  /* osqp_set_default_settings (constants.h) */
  s->rho = 0.1;
  s->sigma = 1e-06;
  s->alpha = 1.6;
  s->eps_abs = 1e-3;
  s->eps_rel = 1e-3;
  s->eps_prim_inf = 1e-4;
  s->eps_dual_inf = 1e-4;
  s->adaptive_rho_tolerance = 5;
  s->scaling = 10;
  s->adaptive_rho = 1;
  s->adaptive_rho_interval = 0;
  s->max_iter = 4000;
  s->check_termination = 25;
  s->warm_start = 1;
  s->scaled_termination = 0;
  s->linsys = ORC_LINSYS_KKT;
}

/* ------------------------------------------------------------------ */
/* scaling.c                                                            */
/* ------------------------------------------------------------------ */
static void limit_scaling(double *D, int n) {
  for (int i = 0; i < n; i++) {
    D[i] = D[i] < ORC_MIN_SCALING ? 1.0 : D[i];
    D[i] = D[i] > ORC_MAX_SCALING ? ORC_MAX_SCALING : D[i];
  }
}
static void compute_inf_norm_cols_KKT(const orc_csc *P, const orc_csc *A, double *D,
                                      double *D_temp_A, double *E, int n) {
  mat_inf_norm_cols_sym_triu(P, D);
  mat_inf_norm_cols(A, D_temp_A);
  for (int i = 0; i < n; i++) D[i] = c_max(D[i], D_temp_A[i]);
  mat_inf_norm_rows(A, E);
}
static void scale_data(orc_workspace *w) {
  int n = w->n, m = w->m;
  double c_temp, inf_norm_q;
  w->c = 1.0;
  for (int i = 0; i < n; i++) w->D[i] = w->Dinv[i] = 1.;
  for (int i = 0; i < m; i++) w->E[i] = w->Einv[i] = 1.;
  for (int it = 0; it < w->settings.scaling; it++) {
    compute_inf_norm_cols_KKT(w->P, w->A, w->D_temp, w->D_temp_A, w->E_temp, n);
    limit_scaling(w->D_temp, n);
    limit_scaling(w->E_temp, m);
    for (int i = 0; i < n; i++) w->D_temp[i] = sqrt(w->D_temp[i]);
    for (int i = 0; i < m; i++) w->E_temp[i] = sqrt(w->E_temp[i]);
    for (int i = 0; i < n; i++) w->D_temp[i] = 1. / w->D_temp[i];
    for (int i = 0; i < m; i++) w->E_temp[i] = 1. / w->E_temp[i];
    mat_premult_diag(w->P, w->D_temp);
    mat_postmult_diag(w->P, w->D_temp);
    mat_premult_diag(w->A, w->E_temp);
    mat_postmult_diag(w->A, w->D_temp);
    for (int i = 0; i < n; i++) w->q[i] = w->D_temp[i] * w->q[i];
    for (int i = 0; i < n; i++) w->D[i] = w->D[i] * w->D_temp[i];
    for (int i = 0; i < m; i++) w->E[i] = w->E[i] * w->E_temp[i];
    /* cost normalisation */
    mat_inf_norm_cols_sym_triu(w->P, w->D_temp);
    c_temp = vec_mean(w->D_temp, n);
    inf_norm_q = vec_norm_inf(w->q, n);
    limit_scaling(&inf_norm_q, 1);
    c_temp = c_max(c_temp, inf_norm_q);
    limit_scaling(&c_temp, 1);
    c_temp = 1. / c_temp;
    mat_mult_scalar(w->P, c_temp);
    for (int i = 0; i < n; i++) w->q[i] *= c_temp;
    w->c *= c_temp;
  }
  w->cinv = 1. / w->c;
  for (int i = 0; i < n; i++) w->Dinv[i] = 1. / w->D[i];
  for (int i = 0; i < m; i++) w->Einv[i] = 1. / w->E[i];
  for (int i = 0; i < m; i++) w->l[i] = w->E[i] * w->l[i];
  for (int i = 0; i < m; i++) w->u[i] = w->E[i] * w->u[i];
}
static void unscale_data(orc_workspace *w) {
  int n = w->n, m = w->m;
  mat_mult_scalar(w->P, w->cinv);
  mat_premult_diag(w->P, w->Dinv);
  mat_postmult_diag(w->P, w->Dinv);
  for (int i = 0; i < n; i++) w->q[i] *= w->cinv;
  for (int i = 0; i < n; i++) w->q[i] = w->Dinv[i] * w->q[i];
  mat_premult_diag(w->A, w->Einv);
  mat_postmult_diag(w->A, w->Dinv);
  for (int i = 0; i < m; i++) w->l[i] = w->Einv[i] * w->l[i];
  for (int i = 0; i < m; i++) w->u[i] = w->Einv[i] * w->u[i];
}

/* ------------------------------------------------------------------ */
/* linear system (replaces lin_sys/direct/qdldl)                        */
/* ------------------------------------------------------------------ */
static void build_csr(orc_workspace *w) {
  const orc_csc *A = w->A;
  int m = w->m, nnz = A->p[A->n_cols];
  memset(w->Ar_p, 0, sizeof(int) * (m + 1));
  for (int k = 0; k < nnz; k++) w->Ar_p[A->i[k] + 1]++;
  for (int i = 0; i < m; i++) w->Ar_p[i + 1] += w->Ar_p[i];
  int *fill = (int *)malloc(sizeof(int) * m);
  memcpy(fill, w->Ar_p, sizeof(int) * m);
  for (int j = 0; j < A->n_cols; j++)
    for (int k = A->p[j]; k < A->p[j + 1]; k++) {
      int r = A->i[k], pos = fill[r]++;
      w->Ar_j[pos] = j;
      w->Ar_x[pos] = A->x[k];
    }
  free(fill);
}

/* factorise with the current P, A, rho_vec; returns 0 on success */
static int linsys_factor(orc_workspace *w) {
  int n = w->n, m = w->m;
  if (w->settings.linsys == ORC_LINSYS_KKT) {
    int N = n + m;
    double *K = w->ls_mat;
    memset(K, 0, sizeof(double) * N * N);
    /* lower triangle of [[P+sigma I, A'],[A, -diag(1/rho)]] (row-major) */
    for (int j = 0; j < n; j++)
      for (int k = w->P->p[j]; k < w->P->p[j + 1]; k++) {
        int i = w->P->i[k]; /* i<=j: entry (i,j) -> lower (j,i) */
        K[(size_t)j * N + i] += w->P->x[k];
      }
    for (int j = 0; j < n; j++) K[(size_t)j * N + j] += w->settings.sigma;
    for (int j = 0; j < n; j++)
      for (int k = w->A->p[j]; k < w->A->p[j + 1]; k++) {
        int i = w->A->i[k];
        K[(size_t)(n + i) * N + j] = w->A->x[k];
      }
    for (int i = 0; i < m; i++) K[(size_t)(n + i) * N + n + i] = -w->rho_inv_vec[i];
    /* dense LDL^T, no pivoting (quasi-definite => strongly factorisable) */
    double *d = w->ls_diag;
    for (int j = 0; j < N; j++) {
      double dj = K[(size_t)j * N + j];
      for (int k = 0; k < j; k++) {
        double ljk = K[(size_t)j * N + k];
        dj -= ljk * ljk * d[k];
      }
      if (dj == 0.0) return 1;
      d[j] = dj;
      for (int i = j + 1; i < N; i++) {
        double v = K[(size_t)i * N + j];
        for (int k = 0; k < j; k++) v -= K[(size_t)i * N + k] * K[(size_t)j * N + k] * d[k];
        K[(size_t)i * N + j] = v / dj;
      }
    }
    return 0;
  } else {
    double *K = w->ls_mat;
    memset(K, 0, sizeof(double) * n * n);
    for (int j = 0; j < n; j++)
      for (int k = w->P->p[j]; k < w->P->p[j + 1]; k++) K[(size_t)j * n + w->P->i[k]] += w->P->x[k];
    for (int j = 0; j < n; j++) K[(size_t)j * n + j] += w->settings.sigma;
    build_csr(w);
    for (int r = 0; r < m; r++) {
      double rho = w->rho_vec[r];
      for (int a = w->Ar_p[r]; a < w->Ar_p[r + 1]; a++) {
        double va = rho * w->Ar_x[a];
        int ja = w->Ar_j[a];
        for (int b = w->Ar_p[r]; b <= a; b++) {
          int jb = w->Ar_j[b]; /* jb <= ja since CSR built in column order */
          K[(size_t)ja * n + jb] += va * w->Ar_x[b];
        }
      }
    }
    /* Cholesky, lower, in place */
    for (int j = 0; j < n; j++) {
      double dj = K[(size_t)j * n + j];
      for (int k = 0; k < j; k++) dj -= K[(size_t)j * n + k] * K[(size_t)j * n + k];
      if (dj <= 0.0) return 1;
      dj = sqrt(dj);
      K[(size_t)j * n + j] = dj;
      for (int i = j + 1; i < n; i++) {
        double v = K[(size_t)i * n + j];
        for (int k = 0; k < j; k++) v -= K[(size_t)i * n + k] * K[(size_t)j * n + k];
        K[(size_t)i * n + j] = v / dj;
      }
    }
    return 0;
  }
}

/* b holds [rhs_x (n) ; rhs_z (m)] on entry, [x_tilde ; z_tilde] on exit
 * (solve_linsys_qdldl, non-embedded branch) */
static void linsys_solve(orc_workspace *w, double *b) {
  int n = w->n, m = w->m;
  if (w->settings.linsys == ORC_LINSYS_KKT) {
    int N = n + m;
    const double *L = w->ls_mat, *d = w->ls_diag;
    double *s = w->ls_work;
    memcpy(s, b, sizeof(double) * N);
    for (int i = 0; i < N; i++) {
      double v = s[i];
      for (int k = 0; k < i; k++) v -= L[(size_t)i * N + k] * s[k];
      s[i] = v;
    }
    for (int i = 0; i < N; i++) s[i] /= d[i];
    for (int i = N - 1; i >= 0; i--) {
      double v = s[i];
      for (int k = i + 1; k < N; k++) v -= L[(size_t)k * N + i] * s[k];
      s[i] = v;
    }
    for (int j = 0; j < n; j++) b[j] = s[j];
    for (int j = 0; j < m; j++) b[j + n] += w->rho_inv_vec[j] * s[j + n];
  } else {
    /* (P + sigma I + A' R A) x = rhs_x + A' R rhs_z ;  z_tilde = A x */
    const double *L = w->ls_mat;
    double *s = w->ls_work;
    for (int j = 0; j < n; j++) s[j] = b[j];
    for (int r = 0; r < m; r++) {
      double t = w->rho_vec[r] * b[n + r];
      for (int a = w->Ar_p[r]; a < w->Ar_p[r + 1]; a++) s[w->Ar_j[a]] += w->Ar_x[a] * t;
    }
    for (int i = 0; i < n; i++) {
      double v = s[i];
      for (int k = 0; k < i; k++) v -= L[(size_t)i * n + k] * s[k];
      s[i] = v / L[(size_t)i * n + i];
    }
    for (int i = n - 1; i >= 0; i--) {
      double v = s[i];
      for (int k = i + 1; k < n; k++) v -= L[(size_t)k * n + i] * s[k];
      s[i] = v / L[(size_t)i * n + i];
    }
    for (int j = 0; j < n; j++) b[j] = s[j];
    for (int r = 0; r < m; r++) {
      double t = 0.0;
      for (int a = w->Ar_p[r]; a < w->Ar_p[r + 1]; a++) t += w->Ar_x[a] * s[w->Ar_j[a]];
      b[n + r] = t;
    }
  }
}

/* ------------------------------------------------------------------ */
/* auxil.c                                                              */
/* ------------------------------------------------------------------ */
static void set_rho_vec(orc_workspace *w) {
  w->settings.rho = c_min(c_max(w->settings.rho, ORC_RHO_MIN), ORC_RHO_MAX);
  for (int i = 0; i < w->m; i++) {
    if ((w->l[i] < -ORC_INFTY * ORC_MIN_SCALING) && (w->u[i] > ORC_INFTY * ORC_MIN_SCALING)) {
      w->constr_type[i] = -1;
      w->rho_vec[i] = ORC_RHO_MIN;
    } else if (w->u[i] - w->l[i] < ORC_RHO_TOL) {
      w->constr_type[i] = 1;
      w->rho_vec[i] = ORC_RHO_EQ_OVER_RHO_INEQ * w->settings.rho;
    } else {
      w->constr_type[i] = 0;
      w->rho_vec[i] = w->settings.rho;
    }
    w->rho_inv_vec[i] = 1. / w->rho_vec[i];
  }
}
static int update_rho_vec(orc_workspace *w) {
  int changed = 0;
  for (int i = 0; i < w->m; i++) {
    if ((w->l[i] < -ORC_INFTY * ORC_MIN_SCALING) && (w->u[i] > ORC_INFTY * ORC_MIN_SCALING)) {
      if (w->constr_type[i] != -1) {
        w->constr_type[i] = -1;
        w->rho_vec[i] = ORC_RHO_MIN;
        w->rho_inv_vec[i] = 1. / ORC_RHO_MIN;
        changed = 1;
      }
    } else if (w->u[i] - w->l[i] < ORC_RHO_TOL) {
      if (w->constr_type[i] != 1) {
        w->constr_type[i] = 1;
        w->rho_vec[i] = ORC_RHO_EQ_OVER_RHO_INEQ * w->settings.rho;
        w->rho_inv_vec[i] = 1. / w->rho_vec[i];
        changed = 1;
      }
    } else {
      if (w->constr_type[i] != 0) {
        w->constr_type[i] = 0;
        w->rho_vec[i] = w->settings.rho;
        w->rho_inv_vec[i] = 1. / w->settings.rho;
        changed = 1;
      }
    }
  }
  if (changed) return linsys_factor(w);
  return 0;
}
static int osqp_update_rho(orc_workspace *w, double rho_new) {
  w->settings.rho = c_min(c_max(rho_new, ORC_RHO_MIN), ORC_RHO_MAX);
  for (int i = 0; i < w->m; i++) {
    if (w->constr_type[i] == 0) {
      w->rho_vec[i] = w->settings.rho;
      w->rho_inv_vec[i] = 1. / w->settings.rho;
    } else if (w->constr_type[i] == 1) {
      w->rho_vec[i] = ORC_RHO_EQ_OVER_RHO_INEQ * w->settings.rho;
      w->rho_inv_vec[i] = 1. / w->rho_vec[i];
    }
  }
  return linsys_factor(w);
}
static void note_margin(orc_workspace *w, double value, double threshold) {
  double den = c_max(c_absval(threshold), 1e-300);
  double mg = c_absval(value - threshold) / den;
  if (mg < w->info.decision_margin) w->info.decision_margin = mg;
}
static double compute_rho_estimate(orc_workspace *w) {
  int n = w->n, m = w->m;
  double pri_res = vec_norm_inf(w->z_prev, m);
  double dua_res = vec_norm_inf(w->x_prev, n);
  double pri_res_norm = vec_norm_inf(w->z, m);
  double temp = vec_norm_inf(w->Ax, m);
  pri_res_norm = c_max(pri_res_norm, temp);
  pri_res /= (pri_res_norm + 1e-10);
  double dua_res_norm = vec_norm_inf(w->q, n);
  temp = vec_norm_inf(w->Aty, n);
  dua_res_norm = c_max(dua_res_norm, temp);
  temp = vec_norm_inf(w->Px, n);
  dua_res_norm = c_max(dua_res_norm, temp);
  dua_res /= (dua_res_norm + 1e-10);
  double rho_estimate = w->settings.rho * sqrt(pri_res / (dua_res + 1e-10));
  rho_estimate = c_min(c_max(rho_estimate, ORC_RHO_MIN), ORC_RHO_MAX);
  return rho_estimate;
}
static int adapt_rho(orc_workspace *w) {
  int exitflag = 0;
  double rho_new = compute_rho_estimate(w);
  w->info.rho_estimate = rho_new;
  note_margin(w, rho_new, w->settings.rho * w->settings.adaptive_rho_tolerance);
  note_margin(w, rho_new, w->settings.rho / w->settings.adaptive_rho_tolerance);
  if ((rho_new > w->settings.rho * w->settings.adaptive_rho_tolerance) ||
      (rho_new < w->settings.rho / w->settings.adaptive_rho_tolerance)) {
    exitflag = osqp_update_rho(w, rho_new);
    w->info.rho_updates += 1;
  }
  return exitflag;
}
static void cold_start(orc_workspace *w) {
  memset(w->x, 0, sizeof(double) * w->n);
  memset(w->z, 0, sizeof(double) * w->m);
  memset(w->y, 0, sizeof(double) * w->m);
}
static void compute_rhs(orc_workspace *w) {
  for (int i = 0; i < w->n; i++) w->xz_tilde[i] = w->settings.sigma * w->x_prev[i] - w->q[i];
  for (int i = 0; i < w->m; i++) w->xz_tilde[i + w->n] = w->z_prev[i] - w->rho_inv_vec[i] * w->y[i];
}
static void update_xz_tilde(orc_workspace *w) {
  compute_rhs(w);
  linsys_solve(w, w->xz_tilde);
}
static void update_x(orc_workspace *w) {
  int n = w->n;
  double alpha = w->settings.alpha;
  for (int i = 0; i < n; i++) w->x[i] = alpha * w->xz_tilde[i] + ((double)1.0 - alpha) * w->x_prev[i];
  for (int i = 0; i < n; i++) w->delta_x[i] = w->x[i] - w->x_prev[i];
}
static void project(orc_workspace *w, double *z) {
  for (int i = 0; i < w->m; i++) z[i] = c_min(c_max(z[i], w->l[i]), w->u[i]);
}
static void update_z(orc_workspace *w) {
  int n = w->n, m = w->m;
  double alpha = w->settings.alpha;
  for (int i = 0; i < m; i++)
    w->z[i] = alpha * w->xz_tilde[i + n] + ((double)1.0 - alpha) * w->z_prev[i] +
              w->rho_inv_vec[i] * w->y[i];
  project(w, w->z);
}
static void update_y(orc_workspace *w) {
  int n = w->n, m = w->m;
  double alpha = w->settings.alpha;
  for (int i = 0; i < m; i++) {
    w->delta_y[i] = w->rho_vec[i] * (alpha * w->xz_tilde[i + n] +
                                     ((double)1.0 - alpha) * w->z_prev[i] - w->z[i]);
    w->y[i] += w->delta_y[i];
  }
}
static double compute_pri_res(orc_workspace *w, double *x, double *z) {
  mat_vec(w->A, x, w->Ax, 0);
  for (int i = 0; i < w->m; i++) w->z_prev[i] = w->Ax[i] - z[i];
  if (w->settings.scaling && !w->settings.scaled_termination)
    return vec_scaled_norm_inf(w->Einv, w->z_prev, w->m);
  return vec_norm_inf(w->z_prev, w->m);
}
static double compute_pri_tol(orc_workspace *w, double eps_abs, double eps_rel) {
  double max_rel_eps, temp_rel_eps;
  if (w->settings.scaling && !w->settings.scaled_termination) {
    max_rel_eps = vec_scaled_norm_inf(w->Einv, w->z, w->m);
    temp_rel_eps = vec_scaled_norm_inf(w->Einv, w->Ax, w->m);
    max_rel_eps = c_max(max_rel_eps, temp_rel_eps);
  } else {
    max_rel_eps = vec_norm_inf(w->z, w->m);
    temp_rel_eps = vec_norm_inf(w->Ax, w->m);
    max_rel_eps = c_max(max_rel_eps, temp_rel_eps);
  }
  return eps_abs + eps_rel * max_rel_eps;
}
static double compute_dua_res(orc_workspace *w, double *x, double *y) {
  int n = w->n;
  memcpy(w->x_prev, w->q, sizeof(double) * n);
  mat_vec(w->P, x, w->Px, 0);
  mat_tpose_vec(w->P, x, w->Px, 1, 1);
  for (int i = 0; i < n; i++) w->x_prev[i] = w->x_prev[i] + w->Px[i];
  if (w->m > 0) {
    mat_tpose_vec(w->A, y, w->Aty, 0, 0);
    for (int i = 0; i < n; i++) w->x_prev[i] = w->x_prev[i] + w->Aty[i];
  }
  if (w->settings.scaling && !w->settings.scaled_termination)
    return w->cinv * vec_scaled_norm_inf(w->Dinv, w->x_prev, n);
  return vec_norm_inf(w->x_prev, n);
}
static double compute_dua_tol(orc_workspace *w, double eps_abs, double eps_rel) {
  double max_rel_eps, temp_rel_eps;
  int n = w->n;
  if (w->settings.scaling && !w->settings.scaled_termination) {
    max_rel_eps = vec_scaled_norm_inf(w->Dinv, w->q, n);
    temp_rel_eps = vec_scaled_norm_inf(w->Dinv, w->Aty, n);
    max_rel_eps = c_max(max_rel_eps, temp_rel_eps);
    temp_rel_eps = vec_scaled_norm_inf(w->Dinv, w->Px, n);
    max_rel_eps = c_max(max_rel_eps, temp_rel_eps);
    max_rel_eps *= w->cinv;
  } else {
    max_rel_eps = vec_norm_inf(w->q, n);
    temp_rel_eps = vec_norm_inf(w->Aty, n);
    max_rel_eps = c_max(max_rel_eps, temp_rel_eps);
    temp_rel_eps = vec_norm_inf(w->Px, n);
    max_rel_eps = c_max(max_rel_eps, temp_rel_eps);
  }
  return eps_abs + eps_rel * max_rel_eps;
}
static int is_primal_infeasible(orc_workspace *w, double eps_prim_inf) {
  int m = w->m, n = w->n;
  double norm_delta_y, ineq_lhs = 0.0;
  for (int i = 0; i < m; i++) {
    if (w->u[i] > ORC_INFTY * ORC_MIN_SCALING) {
      if (w->l[i] < -ORC_INFTY * ORC_MIN_SCALING)
        w->delta_y[i] = 0.0;
      else
        w->delta_y[i] = c_min(w->delta_y[i], 0.0);
    } else if (w->l[i] < -ORC_INFTY * ORC_MIN_SCALING) {
      w->delta_y[i] = c_max(w->delta_y[i], 0.0);
    }
  }
  if (w->settings.scaling && !w->settings.scaled_termination) {
    for (int i = 0; i < m; i++) w->Adelta_x[i] = w->E[i] * w->delta_y[i];
    norm_delta_y = vec_norm_inf(w->Adelta_x, m);
  } else {
    norm_delta_y = vec_norm_inf(w->delta_y, m);
  }
  if (norm_delta_y > eps_prim_inf) {
    for (int i = 0; i < m; i++)
      ineq_lhs += w->u[i] * c_max(w->delta_y[i], 0) + w->l[i] * c_min(w->delta_y[i], 0);
    if (ineq_lhs < -eps_prim_inf * norm_delta_y) {
      mat_tpose_vec(w->A, w->delta_y, w->Atdelta_y, 0, 0);
      if (w->settings.scaling && !w->settings.scaled_termination)
        for (int i = 0; i < n; i++) w->Atdelta_y[i] = w->Dinv[i] * w->Atdelta_y[i];
      return vec_norm_inf(w->Atdelta_y, n) < eps_prim_inf * norm_delta_y;
    }
  }
  return 0;
}
static int is_dual_infeasible(orc_workspace *w, double eps_dual_inf) {
  int n = w->n, m = w->m;
  double norm_delta_x, cost_scaling;
  if (w->settings.scaling && !w->settings.scaled_termination) {
    norm_delta_x = vec_scaled_norm_inf(w->D, w->delta_x, n);
    cost_scaling = w->c;
  } else {
    norm_delta_x = vec_norm_inf(w->delta_x, n);
    cost_scaling = 1.0;
  }
  if (norm_delta_x > eps_dual_inf) {
    if (vec_prod(w->q, w->delta_x, n) < -cost_scaling * eps_dual_inf * norm_delta_x) {
      mat_vec(w->P, w->delta_x, w->Pdelta_x, 0);
      mat_tpose_vec(w->P, w->delta_x, w->Pdelta_x, 1, 1);
      if (w->settings.scaling && !w->settings.scaled_termination)
        for (int i = 0; i < n; i++) w->Pdelta_x[i] = w->Dinv[i] * w->Pdelta_x[i];
      if (vec_norm_inf(w->Pdelta_x, n) < cost_scaling * eps_dual_inf * norm_delta_x) {
        mat_vec(w->A, w->delta_x, w->Adelta_x, 0);
        if (w->settings.scaling && !w->settings.scaled_termination)
          for (int i = 0; i < m; i++) w->Adelta_x[i] = w->Einv[i] * w->Adelta_x[i];
        for (int i = 0; i < m; i++) {
          if (((w->u[i] < ORC_INFTY * ORC_MIN_SCALING) &&
               (w->Adelta_x[i] > eps_dual_inf * norm_delta_x)) ||
              ((w->l[i] > -ORC_INFTY * ORC_MIN_SCALING) &&
               (w->Adelta_x[i] < -eps_dual_inf * norm_delta_x)))
            return 0;
        }
        return 1;
      }
    }
  }
  return 0;
}
static void update_info(orc_workspace *w, int iter) {
  w->info.pri_res = (w->m == 0) ? 0. : compute_pri_res(w, w->x, w->z);
  w->info.dua_res = compute_dua_res(w, w->x, w->y);
  w->info.iter = iter;
}
static int check_termination(orc_workspace *w, int approximate) {
  double eps_prim, eps_dual;
  int exitflag = 0, prim_res_check = 0, dual_res_check = 0, prim_inf_check = 0,
      dual_inf_check = 0;
  double eps_abs = w->settings.eps_abs, eps_rel = w->settings.eps_rel;
  double eps_prim_inf = w->settings.eps_prim_inf, eps_dual_inf = w->settings.eps_dual_inf;
  if ((w->info.pri_res > ORC_INFTY) || (w->info.dua_res > ORC_INFTY)) {
    w->info.status = ORC_NON_CVX;
    return 1;
  }
  if (approximate) {
    eps_abs *= 10;
    eps_rel *= 10;
    eps_prim_inf *= 10;
    eps_dual_inf *= 10;
  }
  if (w->m == 0) {
    prim_res_check = 1;
  } else {
    eps_prim = compute_pri_tol(w, eps_abs, eps_rel);
    if (!approximate) note_margin(w, w->info.pri_res, eps_prim);
    if (w->info.pri_res < eps_prim)
      prim_res_check = 1;
    else
      prim_inf_check = is_primal_infeasible(w, eps_prim_inf);
  }
  eps_dual = compute_dua_tol(w, eps_abs, eps_rel);
  if (!approximate) note_margin(w, w->info.dua_res, eps_dual);
  if (w->info.dua_res < eps_dual)
    dual_res_check = 1;
  else
    dual_inf_check = is_dual_infeasible(w, eps_dual_inf);
  if (prim_res_check && dual_res_check) {
    w->info.status = approximate ? ORC_SOLVED_INACCURATE : ORC_SOLVED;
    exitflag = 1;
  } else if (prim_inf_check) {
    w->info.status = approximate ? ORC_PRIMAL_INFEASIBLE_INACCURATE : ORC_PRIMAL_INFEASIBLE;
    if (w->settings.scaling && !w->settings.scaled_termination)
      for (int i = 0; i < w->m; i++) w->delta_y[i] = w->E[i] * w->delta_y[i];
    exitflag = 1;
  } else if (dual_inf_check) {
    w->info.status = approximate ? ORC_DUAL_INFEASIBLE_INACCURATE : ORC_DUAL_INFEASIBLE;
    if (w->settings.scaling && !w->settings.scaled_termination)
      for (int i = 0; i < w->n; i++) w->delta_x[i] = w->D[i] * w->delta_x[i];
    exitflag = 1;
  }
  return exitflag;
}
static int has_solution(const orc_info *info) {
  return ((info->status != ORC_PRIMAL_INFEASIBLE) &&
          (info->status != ORC_PRIMAL_INFEASIBLE_INACCURATE) &&
          (info->status != ORC_DUAL_INFEASIBLE) &&
          (info->status != ORC_DUAL_INFEASIBLE_INACCURATE) && (info->status != ORC_NON_CVX));
}
static void store_solution(orc_workspace *w) {
  int n = w->n, m = w->m;
  if (has_solution(&w->info)) {
    memcpy(w->sol_x, w->x, sizeof(double) * n);
    memcpy(w->sol_y, w->y, sizeof(double) * m);
    if (w->settings.scaling) {
      for (int i = 0; i < n; i++) w->sol_x[i] = w->D[i] * w->sol_x[i];
      for (int i = 0; i < m; i++) w->sol_y[i] = w->E[i] * w->sol_y[i];
      for (int i = 0; i < m; i++) w->sol_y[i] *= w->cinv;
    }
  } else {
    for (int i = 0; i < n; i++) w->sol_x[i] = NAN;
    for (int i = 0; i < m; i++) w->sol_y[i] = NAN;
    cold_start(w);
  }
}
static void reset_info(orc_info *info) { info->status = ORC_UNSOLVED; }

/* ------------------------------------------------------------------ */
/* osqp.c                                                               */
/* ------------------------------------------------------------------ */
static double *dalloc(int n) { return (double *)calloc((size_t)(n > 0 ? n : 1), sizeof(double)); }

orc_workspace *orc_setup(const orc_csc *P_upper, const double *q, const orc_csc *A,
                         const double *l, const double *u, const orc_settings *s) {
  orc_workspace *w = (orc_workspace *)calloc(1, sizeof(orc_workspace));
  int n = P_upper->n_cols, m = A->n_rows;
  w->n = n;
  w->m = m;
  w->settings = *s;
  w->P = csc_copy(P_upper);
  w->A = csc_copy(A);
  w->q = dalloc(n);
  w->l = dalloc(m);
  w->u = dalloc(m);
  memcpy(w->q, q, sizeof(double) * n);
  for (int i = 0; i < m; i++) {
    /* osqp-cpp clips to +-OSQP_INFTY before osqp_setup */
    w->l[i] = c_max(l[i], -ORC_INFTY);
    w->u[i] = c_min(u[i], ORC_INFTY);
  }
  w->D = dalloc(n);
  w->Dinv = dalloc(n);
  w->E = dalloc(m);
  w->Einv = dalloc(m);
  w->D_temp = dalloc(n);
  w->D_temp_A = dalloc(n);
  w->E_temp = dalloc(m);
  w->rho_vec = dalloc(m);
  w->rho_inv_vec = dalloc(m);
  w->constr_type = (int *)calloc((size_t)m, sizeof(int));
  w->x = dalloc(n);
  w->y = dalloc(m);
  w->z = dalloc(m);
  w->xz_tilde = dalloc(n + m);
  w->x_prev = dalloc(n);
  w->z_prev = dalloc(m);
  w->Ax = dalloc(m);
  w->Px = dalloc(n);
  w->Aty = dalloc(n);
  w->delta_y = dalloc(m);
  w->Atdelta_y = dalloc(n);
  w->delta_x = dalloc(n);
  w->Pdelta_x = dalloc(n);
  w->Adelta_x = dalloc(m);
  w->sol_x = dalloc(n);
  w->sol_y = dalloc(m);
  int nnzA = A->p[A->n_cols];
  w->Ar_p = (int *)calloc((size_t)m + 1, sizeof(int));
  w->Ar_j = (int *)calloc((size_t)(nnzA > 0 ? nnzA : 1), sizeof(int));
  w->Ar_x = dalloc(nnzA);
  w->ls_dim = (s->linsys == ORC_LINSYS_KKT) ? n + m : n;
  w->ls_mat = dalloc(w->ls_dim * w->ls_dim);
  w->ls_diag = dalloc(n + m);
  w->ls_work = dalloc(n + m);

  if (w->settings.scaling) {
    scale_data(w);
  } else {
    w->c = w->cinv = 1.0;
    for (int i = 0; i < n; i++) w->D[i] = w->Dinv[i] = 1.;
    for (int i = 0; i < m; i++) w->E[i] = w->Einv[i] = 1.;
  }
  set_rho_vec(w);
  if (linsys_factor(w)) {
    orc_cleanup(w);
    return NULL;
  }
  cold_start(w);
  reset_info(&w->info);
  w->info.iter = 0;
  w->info.rho_updates = 0;
  w->info.rho_estimate = w->settings.rho;
  w->info.decision_margin = 1e300;
  /* adaptive_rho_interval == 0: resolved inside orc_solve like the
   * non-PROFILING branch of osqp_solve */
  return w;
}

void orc_cleanup(orc_workspace *w) {
  if (!w) return;
  orc_csc_free(w->P);
  orc_csc_free(w->A);
  free(w->q); free(w->l); free(w->u);
  free(w->D); free(w->Dinv); free(w->E); free(w->Einv);
  free(w->D_temp); free(w->D_temp_A); free(w->E_temp);
  free(w->rho_vec); free(w->rho_inv_vec); free(w->constr_type);
  free(w->x); free(w->y); free(w->z); free(w->xz_tilde); free(w->x_prev); free(w->z_prev);
  free(w->Ax); free(w->Px); free(w->Aty); free(w->delta_y); free(w->Atdelta_y);
  free(w->delta_x); free(w->Pdelta_x); free(w->Adelta_x);
  free(w->sol_x); free(w->sol_y);
  free(w->Ar_p); free(w->Ar_j); free(w->Ar_x);
  free(w->ls_mat); free(w->ls_diag); free(w->ls_work);
  free(w);
}

int orc_update_P_A(orc_workspace *w, const orc_csc *P_upper, const orc_csc *A) {
  if (!orc_csc_same_pattern(P_upper, w->P) || !orc_csc_same_pattern(A, w->A)) return 1;
  if (w->settings.scaling) unscale_data(w);
  memcpy(w->P->x, P_upper->x, sizeof(double) * P_upper->p[P_upper->n_cols]);
  memcpy(w->A->x, A->x, sizeof(double) * A->p[A->n_cols]);
  if (w->settings.scaling) scale_data(w);
  int ex = linsys_factor(w);
  reset_info(&w->info);
  return ex ? 2 : 0;
}
void orc_update_lin_cost(orc_workspace *w, const double *q_new) {
  memcpy(w->q, q_new, sizeof(double) * w->n);
  if (w->settings.scaling) {
    for (int i = 0; i < w->n; i++) w->q[i] = w->D[i] * w->q[i];
    for (int i = 0; i < w->n; i++) w->q[i] *= w->c;
  }
  reset_info(&w->info);
}
int orc_update_bounds(orc_workspace *w, const double *l_new, const double *u_new) {
  for (int i = 0; i < w->m; i++)
    if (l_new[i] > u_new[i]) return 1;
  for (int i = 0; i < w->m; i++) {
    w->l[i] = c_max(l_new[i], -ORC_INFTY); /* osqp-cpp SetBounds clips too */
    w->u[i] = c_min(u_new[i], ORC_INFTY);
  }
  if (w->settings.scaling) {
    for (int i = 0; i < w->m; i++) w->l[i] = w->E[i] * w->l[i];
    for (int i = 0; i < w->m; i++) w->u[i] = w->E[i] * w->u[i];
  }
  reset_info(&w->info);
  return update_rho_vec(w);
}
void orc_warm_start(orc_workspace *w, const double *x, const double *y) {
  if (!w->settings.warm_start) w->settings.warm_start = 1;
  memcpy(w->x, x, sizeof(double) * w->n);
  memcpy(w->y, y, sizeof(double) * w->m);
  if (w->settings.scaling) {
    for (int i = 0; i < w->n; i++) w->x[i] = w->Dinv[i] * w->x[i];
    for (int i = 0; i < w->m; i++) w->y[i] = w->Einv[i] * w->y[i];
    for (int i = 0; i < w->m; i++) w->y[i] *= w->c;
  }
  mat_vec(w->A, w->x, w->z, 0);
}

int orc_solve(orc_workspace *w) {
  int iter, can_check_termination = 0;
  int max_iter = w->settings.max_iter;
  double *tmp;
  w->info.decision_margin = 1e300;
  w->info.rho_updates = 0;
  w->info.status = ORC_UNSOLVED;
  if (!w->settings.warm_start) cold_start(w);

  for (iter = 1; iter <= max_iter; iter++) {
    tmp = w->x; w->x = w->x_prev; w->x_prev = tmp;
    tmp = w->z; w->z = w->z_prev; w->z_prev = tmp;

    update_xz_tilde(w);
    update_x(w);
    update_z(w);
    update_y(w);

    can_check_termination =
        w->settings.check_termination && (iter % w->settings.check_termination == 0);
    if (can_check_termination) {
      update_info(w, iter);
      if (check_termination(w, 0)) break;
    }

    /* non-PROFILING branch: fix the automatic interval */
    if (w->settings.adaptive_rho && !w->settings.adaptive_rho_interval) {
      if (w->settings.check_termination)
        w->settings.adaptive_rho_interval =
            ORC_ADAPTIVE_RHO_MULTIPLE_TERMINATION * w->settings.check_termination;
      else
        w->settings.adaptive_rho_interval = ORC_ADAPTIVE_RHO_FIXED;
    }
    if (w->settings.adaptive_rho && w->settings.adaptive_rho_interval &&
        (iter % w->settings.adaptive_rho_interval == 0)) {
      if (!can_check_termination) update_info(w, iter);
      if (adapt_rho(w)) return 1;
    }
  }
  if (iter > max_iter) iter = max_iter; /* loop ran to completion */
  if (!can_check_termination) {
    update_info(w, iter);
    check_termination(w, 0);
  }
  if (w->info.status == ORC_UNSOLVED) {
    if (!check_termination(w, 1)) w->info.status = ORC_MAX_ITER_REACHED;
  }
  w->info.rho_estimate = compute_rho_estimate(w);
  store_solution(w);
  return 0;
}

const double *orc_solution_x(const orc_workspace *w) { return w->sol_x; }
const double *orc_solution_y(const orc_workspace *w) { return w->sol_y; }
const orc_info *orc_get_info(const orc_workspace *w) { return &w->info; }
double orc_get_rho(const orc_workspace *w) { return w->settings.rho; }
void orc_get_scaled_state(const orc_workspace *w, double *x, double *z, double *y, double *D,
                          double *E, double *c) {
  if (x) memcpy(x, w->x, sizeof(double) * w->n);
  if (z) memcpy(z, w->z, sizeof(double) * w->m);
  if (y) memcpy(y, w->y, sizeof(double) * w->m);
  if (D) memcpy(D, w->D, sizeof(double) * w->n);
  if (E) memcpy(E, w->E, sizeof(double) * w->m);
  if (c) *c = w->c;
}
