/*
 * noc_oracle.c — CPU oracle (plain C99) for the batched LQR / SLQ hot path.
 * TEST INFRASTRUCTURE ONLY — see noc_oracle.h.  PARITY UNPINNED against KahnShi/NOC: the reference
 * snapshot has no solver source; every function below restates DESIGN.md §2, which
 * freezes SURVEY.md Appendix A.  The slot this code would occupy in the reference is the commented
 * library target lqr_control/CMakeLists.txt:124-126 (src/${PROJECT_NAME}/lqr_control.cpp, absent).
 *
 * Build flags matter (oracle/Makefile): -ffp-contract=off -fno-fast-math, explicit fma() only where
 * the spec says "fma chain".  Plain a*b+c expressions are two IEEE roundings on both CPU and GPU.
 */
#include "noc_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define NMAX 24
#define MMAX 8
#define NMMAX (NMAX + MMAX)

/* ---------------------------------------------------------------- sincos */
/* Cody–Waite reduction by pi/2 with two fma steps, then the classic degree-13 / degree-14
 * minimax polynomials (coefficients as published in Sun fdlibm k_sin.c / k_cos.c). */
static const double TWO_OVER_PI = 6.36619772367581382433e-01;
static const double PIO2_HI = 1.57079632679489655800e+00;
static const double PIO2_LO = 6.12323399573676603587e-17;
static const double S1 = -1.66666666666666324348e-01, S2 = 8.33333333332248946124e-03,
                    S3 = -1.98412698298579493134e-04, S4 = 2.75573137070700676789e-06,
                    S5 = -2.50507602534068634195e-08, S6 = 1.58969099521155010221e-10;
static const double C1 = 4.16666666666666019037e-02, C2 = -1.38888888888741095749e-03,
                    C3 = 2.48015872894767294178e-05, C4 = -2.75573143513906633035e-07,
                    C5 = 2.08757232129817482790e-09, C6 = -1.13596475577881948265e-11;

void noc_oracle_sincos(double x, double* s, double* c) {
  const double kd = rint(x * TWO_OVER_PI);
  double r = fma(-kd, PIO2_HI, x);
  r = fma(-kd, PIO2_LO, r);
  const double z = r * r;
  /* sin(r) = r + r*z*(S1 + z*(S2 + ... )) */
  double ps = fma(z, S6, S5);
  ps = fma(z, ps, S4);
  ps = fma(z, ps, S3);
  ps = fma(z, ps, S2);
  ps = fma(z, ps, S1);
  const double sr = fma(r * z, ps, r);
  /* cos(r) = 1 - z/2 + z*z*(C1 + z*(C2 + ...)) */
  double pc = fma(z, C6, C5);
  pc = fma(z, pc, C4);
  pc = fma(z, pc, C3);
  pc = fma(z, pc, C2);
  pc = fma(z, pc, C1);
  const double cr = fma(z * z, pc, fma(z, -0.5, 1.0));
  const long long q = (long long)kd;
  switch ((int)(q & 3)) {
    case 0: *s = sr;  *c = cr;  break;
    case 1: *s = cr;  *c = -sr; break;
    case 2: *s = -sr; *c = -cr; break;
    default: *s = -cr; *c = sr; break;
  }
}

/* ---------------------------------------------------------------- quadrotor */
#define Q_MASS 1.0
#define Q_JX 0.01
#define Q_JY 0.01
#define Q_JZ 0.02
#define Q_ARM 0.2
#define Q_YAWC 0.02
#define Q_G 9.81
/* reciprocal inertias: the spec multiplies by these compile-time constants */
#define Q_INV_JX (1.0 / Q_JX)
#define Q_INV_JY (1.0 / Q_JY)
#define Q_INV_JZ (1.0 / Q_JZ)
/* dual-UAV coupling */
#define D_KS 2.0
#define D_CS 0.5
static const double D_REST[3] = {1.0, 0.0, 0.0};

int noc_oracle_model_dims(int model, int* n, int* m) {
  if (model == NOC_O_MODEL_QUAD12) { *n = 12; *m = 4; return 0; }
  if (model == NOC_O_MODEL_DUAL24) { *n = 24; *m = 8; return 0; }
  return -1;
}

void noc_oracle_hover_input(int model, double* u_h) {
  int n, m;
  noc_oracle_model_dims(model, &n, &m);
  for (int a = 0; a < m; ++a) u_h[a] = (Q_MASS * Q_G) / 4.0;
}

typedef struct { double sph, cph, sth, cth, sps, cps; } trig_t;

static void quad_trig(const double* x, trig_t* t) {
  noc_oracle_sincos(x[6], &t->sph, &t->cph);
  noc_oracle_sincos(x[7], &t->sth, &t->cth);
  noc_oracle_sincos(x[8], &t->sps, &t->cps);
}

/* xdot of one vehicle; x[12], u[4] */
static void quad_f(const double* x, const double* u, double* xd) {
  trig_t t;
  quad_trig(x, &t);
  const double wx = x[9], wy = x[10], wz = x[11];
  const double T = ((u[0] + u[1]) + u[2]) + u[3];
  const double tau_x = Q_ARM * (u[1] - u[3]);
  const double tau_y = Q_ARM * (u[2] - u[0]);
  const double tau_z = Q_YAWC * (((u[0] - u[1]) + u[2]) - u[3]);
  const double a = T / Q_MASS;
  const double r13 = (t.cps * t.sth) * t.cph + t.sps * t.sph;
  const double r23 = (t.sps * t.sth) * t.cph - t.cps * t.sph;
  const double r33 = t.cth * t.cph;
  xd[0] = x[3]; xd[1] = x[4]; xd[2] = x[5];
  xd[3] = a * r13;
  xd[4] = a * r23;
  xd[5] = a * r33 - Q_G;
  const double sec = 1.0 / t.cth;
  const double tth = t.sth * sec;
  xd[6] = (wx + (t.sph * tth) * wy) + (t.cph * tth) * wz;
  xd[7] = t.cph * wy - t.sph * wz;
  xd[8] = (t.sph * wy + t.cph * wz) * sec;
  xd[9] = (tau_x - (Q_JZ - Q_JY) * (wy * wz)) * Q_INV_JX;
  xd[10] = (tau_y - (Q_JX - Q_JZ) * (wz * wx)) * Q_INV_JY;
  xd[11] = (tau_z - (Q_JY - Q_JX) * (wx * wy)) * Q_INV_JZ;
}

/* continuous Jacobians of one vehicle, dense row-major: Jx[12][12], Ju[12][4] (zero-filled first) */
static void quad_jac(const double* x, const double* u, double* Jx, double* Ju) {
  trig_t t;
  quad_trig(x, &t);
  memset(Jx, 0, 144 * sizeof(double));
  memset(Ju, 0, 48 * sizeof(double));
  const double wx = x[9], wy = x[10], wz = x[11];
  const double T = ((u[0] + u[1]) + u[2]) + u[3];
  const double a = T / Q_MASS;
  const double r13 = (t.cps * t.sth) * t.cph + t.sps * t.sph;
  const double r23 = (t.sps * t.sth) * t.cph - t.cps * t.sph;
  const double r33 = t.cth * t.cph;
  /* pdot = v */
  Jx[0 * 12 + 3] = 1.0; Jx[1 * 12 + 4] = 1.0; Jx[2 * 12 + 5] = 1.0;
  /* vdot = a * r3 - g e3 : d/d(phi,theta,psi) */
  Jx[3 * 12 + 6] = a * (t.sps * t.cph - (t.cps * t.sth) * t.sph);
  Jx[3 * 12 + 7] = a * ((t.cps * t.cth) * t.cph);
  Jx[3 * 12 + 8] = a * (t.cps * t.sph - (t.sps * t.sth) * t.cph);
  Jx[4 * 12 + 6] = a * (-((t.sps * t.sth) * t.sph) - t.cps * t.cph);
  Jx[4 * 12 + 7] = a * ((t.sps * t.cth) * t.cph);
  Jx[4 * 12 + 8] = a * ((t.cps * t.sth) * t.cph + t.sps * t.sph);
  Jx[5 * 12 + 6] = a * (-(t.cth * t.sph));
  Jx[5 * 12 + 7] = a * (-(t.sth * t.cph));
  /* vdot wrt u: every rotor adds thrust along r3 */
  for (int j = 0; j < 4; ++j) {
    Ju[3 * 4 + j] = r13 / Q_MASS;
    Ju[4 * 4 + j] = r23 / Q_MASS;
    Ju[5 * 4 + j] = r33 / Q_MASS;
  }
  /* Euler rates */
  const double sec = 1.0 / t.cth;
  const double tth = t.sth * sec;
  const double sw = t.sph * wy + t.cph * wz;  /* s_phi*wy + c_phi*wz */
  const double cw = t.cph * wy - t.sph * wz;  /* c_phi*wy - s_phi*wz */
  Jx[6 * 12 + 6] = cw * tth;
  Jx[6 * 12 + 7] = sw * (sec * sec);
  Jx[6 * 12 + 9] = 1.0;
  Jx[6 * 12 + 10] = t.sph * tth;
  Jx[6 * 12 + 11] = t.cph * tth;
  Jx[7 * 12 + 6] = -sw;
  Jx[7 * 12 + 10] = t.cph;
  Jx[7 * 12 + 11] = -t.sph;
  Jx[8 * 12 + 6] = cw * sec;
  Jx[8 * 12 + 7] = (sw * tth) * sec;
  Jx[8 * 12 + 10] = t.sph * sec;
  Jx[8 * 12 + 11] = t.cph * sec;
  /* body rates */
  Jx[9 * 12 + 10] = -((Q_JZ - Q_JY) * wz) * Q_INV_JX;
  Jx[9 * 12 + 11] = -((Q_JZ - Q_JY) * wy) * Q_INV_JX;
  Jx[10 * 12 + 9] = -((Q_JX - Q_JZ) * wz) * Q_INV_JY;
  Jx[10 * 12 + 11] = -((Q_JX - Q_JZ) * wx) * Q_INV_JY;
  Jx[11 * 12 + 9] = -((Q_JY - Q_JX) * wy) * Q_INV_JZ;
  Jx[11 * 12 + 10] = -((Q_JY - Q_JX) * wx) * Q_INV_JZ;
  Ju[9 * 4 + 1] = Q_ARM / Q_JX;   Ju[9 * 4 + 3] = -(Q_ARM / Q_JX);
  Ju[10 * 4 + 0] = -(Q_ARM / Q_JY); Ju[10 * 4 + 2] = Q_ARM / Q_JY;
  Ju[11 * 4 + 0] = Q_YAWC / Q_JZ;  Ju[11 * 4 + 1] = -(Q_YAWC / Q_JZ);
  Ju[11 * 4 + 2] = Q_YAWC / Q_JZ;  Ju[11 * 4 + 3] = -(Q_YAWC / Q_JZ);
}

static void model_f(int model, const double* x, const double* u, double* xd) {
  if (model == NOC_O_MODEL_QUAD12) { quad_f(x, u, xd); return; }
  quad_f(x, u, xd);
  quad_f(x + 12, u + 4, xd + 12);
  for (int i = 0; i < 3; ++i) {
    const double c = (D_KS * ((x[12 + i] - x[i]) - D_REST[i]) + D_CS * (x[15 + i] - x[3 + i])) / Q_MASS;
    xd[3 + i] = xd[3 + i] + c;
    xd[15 + i] = xd[15 + i] - c;
  }
}

void noc_oracle_step(int model, const double* x, const double* u, double dt, double* x_next) {
  int n, m;
  double xd[NMAX];
  noc_oracle_model_dims(model, &n, &m);
  model_f(model, x, u, xd);
  for (int i = 0; i < n; ++i) x_next[i] = x[i] + dt * xd[i];
}

void noc_oracle_linearize(int model, const double* x, const double* u, double dt, double* A, double* B) {
  int n, m;
  noc_oracle_model_dims(model, &n, &m);
  double Jx[144], Ju[48];
  memset(A, 0, (size_t)n * n * sizeof(double));
  memset(B, 0, (size_t)n * m * sizeof(double));
  const int nveh = (model == NOC_O_MODEL_DUAL24) ? 2 : 1;
  for (int v = 0; v < nveh; ++v) {
    quad_jac(x + 12 * v, u + 4 * v, Jx, Ju);
    for (int i = 0; i < 12; ++i) {
      for (int j = 0; j < 12; ++j) {
        const double d = (i == j) ? 1.0 : 0.0;
        A[(12 * v + i) * n + (12 * v + j)] = d + dt * Jx[i * 12 + j];
      }
      for (int j = 0; j < 4; ++j) B[(12 * v + i) * m + (4 * v + j)] = dt * Ju[i * 4 + j];
    }
  }
  if (model == NOC_O_MODEL_DUAL24) {
    const double ks = D_KS / Q_MASS, cs = D_CS / Q_MASS;
    for (int i = 0; i < 3; ++i) {
      /* vehicle 1 velocity rows 3+i */
      A[(3 + i) * n + i] = dt * (-ks);
      A[(3 + i) * n + (12 + i)] = dt * ks;
      A[(3 + i) * n + (3 + i)] = 1.0 + dt * (-cs);
      A[(3 + i) * n + (15 + i)] = dt * cs;
      /* vehicle 2 velocity rows 15+i */
      A[(15 + i) * n + i] = dt * ks;
      A[(15 + i) * n + (12 + i)] = dt * (-ks);
      A[(15 + i) * n + (3 + i)] = dt * cs;
      A[(15 + i) * n + (15 + i)] = 1.0 + dt * (-cs);
    }
  }
}

void noc_oracle_rollout_open_loop(int model, int N, double dt, const double* x0, const double* ubar,
                                  double* xbar) {
  int n, m;
  noc_oracle_model_dims(model, &n, &m);
  memcpy(xbar, x0, (size_t)n * sizeof(double));
  for (int k = 0; k < N; ++k) noc_oracle_step(model, xbar + (size_t)k * n, ubar + (size_t)k * m, dt, xbar + (size_t)(k + 1) * n);
}

static void pack_ab(int n, int m, const double* A, const double* B, int layout, double* rec) {
  if (layout == NOC_O_LAYOUT_A_THEN_B) {
    memcpy(rec, A, (size_t)n * n * sizeof(double));
    memcpy(rec + n * n, B, (size_t)n * m * sizeof(double));
  } else {
    for (int i = 0; i < n; ++i) {
      memcpy(rec + (size_t)i * (n + m), A + (size_t)i * n, (size_t)n * sizeof(double));
      memcpy(rec + (size_t)i * (n + m) + n, B + (size_t)i * m, (size_t)m * sizeof(double));
    }
  }
}

void noc_oracle_linearize_traj(int model, int N, double dt, const double* xbar, const double* ubar,
                               int layout, double* AB) {
  int n, m;
  noc_oracle_model_dims(model, &n, &m);
  double A[NMAX * NMAX], B[NMAX * MMAX];
  const size_t rec = (size_t)n * n + (size_t)n * m;
  for (int k = 0; k < N; ++k) {
    noc_oracle_linearize(model, xbar + (size_t)k * n, ubar + (size_t)k * m, dt, A, B);
    pack_ab(n, m, A, B, layout, AB + (size_t)k * rec);
  }
}

/* ---------------------------------------------------------------- Riccati step */
/* F[n][n+m] = [A | B] rows.  Unpack one record. */
static void unpack_F(int n, int m, const double* rec, int layout, double* F) {
  const int nm = n + m;
  if (layout == NOC_O_LAYOUT_ROWS_AB) { memcpy(F, rec, (size_t)n * nm * sizeof(double)); return; }
  for (int i = 0; i < n; ++i) {
    for (int j = 0; j < n; ++j) F[i * nm + j] = rec[i * n + j];
    for (int j = 0; j < m; ++j) F[i * nm + n + j] = rec[n * n + i * m + j];
  }
}

typedef struct {
  int n, m;
  double G[NMAX * NMMAX];   /* P*F */
  double Hxx[NMAX * NMAX];  /* upper triangle valid */
  double Hux[MMAX * NMAX];
  double Huu[MMAX * MMAX];  /* lower triangle valid */
  double L[MMAX * MMAX], D[MMAX], d[MMAX]; /* unit-lower L, pivots D, d = 1/D */
} step_ws;

/* G = P F ; Hxx (upper) = Q + A^T G_A ; Hux = B^T G_A ; Huu (lower) = R + B^T G_B (+mu on diag) */
static void form_H(step_ws* w, const double* P, const double* F, const double* Q, const double* R, double mu) {
  const int n = w->n, m = w->m, nm = n + m;
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < nm; ++j) {
      double acc = 0.0;
      for (int l = 0; l < n; ++l) acc = fma(P[i * n + l], F[l * nm + j], acc);
      w->G[i * nm + j] = acc;
    }
  for (int i = 0; i < n; ++i)
    for (int j = i; j < n; ++j) {
      double acc = Q[i * n + j];
      for (int l = 0; l < n; ++l) acc = fma(F[l * nm + i], w->G[l * nm + j], acc);
      w->Hxx[i * n + j] = acc;
    }
  for (int a = 0; a < m; ++a)
    for (int j = 0; j < n; ++j) {
      double acc = 0.0;
      for (int l = 0; l < n; ++l) acc = fma(F[l * nm + n + a], w->G[l * nm + j], acc);
      w->Hux[a * n + j] = acc;
    }
  for (int a = 0; a < m; ++a)
    for (int b = 0; b <= a; ++b) {
      double acc = R[a * m + b];
      for (int l = 0; l < n; ++l) acc = fma(F[l * nm + n + a], w->G[l * nm + n + b], acc);
      if (a == b) acc = acc + mu;
      w->Huu[a * m + b] = acc;
    }
}

/* Square-root-free Cholesky Huu = L D L^T (L unit lower, D diagonal), left-looking:
 *   w_l = L[j][l] * D[l]                       (plain products, l < j)
 *   D[j] = Huu[j][j] - sum_l L[j][l] w_l       (fma chain, ascending l); r[j] = 1 / D[j]
 *   L[i][j] = (Huu[i][j] - sum_l L[i][l] w_l) * r[j]
 * Huu is positive definite iff every pivot D[j] > 0; pivots outside (1e-280, 1e280) also count as failure,
 * so that 1/D[j] never leaves the normal range.  returns 0 or NOT_PD */
#define PIVOT_MIN 1e-280
#define PIVOT_MAX 1e280
static int chol_lower(step_ws* w) {
  const int m = w->m;
  double v[MMAX];
  for (int j = 0; j < m; ++j) {
    for (int l = 0; l < j; ++l) v[l] = w->L[j * m + l] * w->D[l];
    double s = w->Huu[j * m + j];
    for (int l = 0; l < j; ++l) s = fma(-w->L[j * m + l], v[l], s);
    if (!(s > PIVOT_MIN) || !(s < PIVOT_MAX)) return NOC_O_NOT_PD;
    w->D[j] = s;
    w->d[j] = 1.0 / s;
    w->L[j * m + j] = 1.0;
    for (int i = j + 1; i < m; ++i) {
      double t = w->Huu[i * m + j];
      for (int l = 0; l < j; ++l) t = fma(-w->L[i * m + l], v[l], t);
      w->L[i * m + j] = t * w->d[j];
    }
  }
  return NOC_O_OK;
}

/* out = -(L D L^T)^{-1} rhs for one right-hand side of length m (stride between elements = rs) */
static void chol_solve_neg(const step_ws* w, const double* rhs, int rs, double* out, int os) {
  const int m = w->m;
  double y[MMAX], z[MMAX];
  for (int i = 0; i < m; ++i) {
    double s = rhs[i * rs];
    for (int l = 0; l < i; ++l) s = fma(-w->L[i * m + l], y[l], s);
    y[i] = s;
  }
  for (int i = m - 1; i >= 0; --i) {
    double s = y[i] * w->d[i];
    for (int l = i + 1; l < m; ++l) s = fma(-w->L[l * m + i], z[l], s);
    z[i] = s;
  }
  for (int i = 0; i < m; ++i) out[i * os] = -z[i];
}

/* K = -Huu^{-1} Hux ; P' upper = Hxx + Hux^T K, mirrored */
static void gains_and_value(step_ws* w, double* K, double* Pn) {
  const int n = w->n, m = w->m;
  for (int j = 0; j < n; ++j) chol_solve_neg(w, w->Hux + j, n, K + j, n);
  for (int i = 0; i < n; ++i)
    for (int j = i; j < n; ++j) {
      double acc = w->Hxx[i * n + j];
      for (int a = 0; a < m; ++a) acc = fma(w->Hux[a * n + i], K[a * n + j], acc);
      Pn[i * n + j] = acc;
      Pn[j * n + i] = acc;
    }
}

/* ---- input box u_min <= u <= u_max: control-limited DDP (Tassa, Mansard, Todorov 2014), DESIGN.md §2.4b ----
 * Backward pass: the feed-forward k solves  min 0.5 k'Quu k + Qu'k  s.t.  lo <= k <= hi  (lo = u_min - ubar,
 * hi = u_max - ubar) by projected Newton; inputs clamped at the solution get zero feedback rows, the free ones
 * K_f = -Quu_ff^{-1} Qux_f.  Every sum is an ascending fma chain, so the CUDA kernel reproduces it bit for bit. */
#define QP_MAXIT 16
#define QP_MAXLS 12
#define QP_ARMIJO 0.1
#define QP_GTOL 1e-10

static double huu_sym(const step_ws* w, int a, int b) {
  return (b <= a) ? w->Huu[a * w->m + b] : w->Huu[b * w->m + a];
}

/* r = Huu x ; returns f(x) = sum_a x_a (0.5 r_a + g_a) */
static double qp_eval(const step_ws* w, const double* g, const double* x, double* r) {
  const int m = w->m;
  for (int a = 0; a < m; ++a) {
    double s = 0.0;
    for (int b = 0; b < m; ++b) s = fma(huu_sym(w, a, b), x[b], s);
    r[a] = s;
  }
  double f = 0.0;
  for (int a = 0; a < m; ++a) f = fma(x[a], fma(0.5, r[a], g[a]), f);
  return f;
}

/* factorisation of Huu with the rows / columns of the clamped inputs replaced by those of the identity */
static int chol_lower_masked(const step_ws* w, const int* clamped, step_ws* wm) {
  const int m = w->m;
  wm->n = w->n; wm->m = m;
  for (int a = 0; a < m; ++a)
    for (int b = 0; b <= a; ++b)
      wm->Huu[a * m + b] = (clamped[a] || clamped[b]) ? (a == b ? 1.0 : 0.0) : w->Huu[a * m + b];
  return chol_lower(wm);
}

static double clampd(double v, double lo, double hi) { return v < lo ? lo : (v > hi ? hi : v); }

/* x: in = start point, out = solution; clamped[a] = 1 for the inputs held at a bound by the solution */
static void box_qp(const step_ws* w, const double* g, const double* lo, const double* hi, double* x, int* clamped) {
  const int m = w->m;
  double r[MMAX], grad[MMAX], rhs[MMAX], dx[MMAX], xc[MMAX], rc[MMAX];
  step_ws wm;
  double gmax = 0.0;
  for (int a = 0; a < m; ++a) { x[a] = clampd(x[a], lo[a], hi[a]); if (fabs(g[a]) > gmax) gmax = fabs(g[a]); }
  double f = qp_eval(w, g, x, r);
  for (int it = 0; it < QP_MAXIT; ++it) {
    int nfree = 0;
    double gn = 0.0;
    for (int a = 0; a < m; ++a) {
      grad[a] = r[a] + g[a];
      clamped[a] = (x[a] <= lo[a] && grad[a] > 0.0) || (x[a] >= hi[a] && grad[a] < 0.0);
      if (!clamped[a]) { ++nfree; if (fabs(grad[a]) > gn) gn = fabs(grad[a]); }
    }
    if (nfree == 0 || gn < QP_GTOL * (1.0 + gmax)) return;
    if (chol_lower_masked(w, clamped, &wm) != NOC_O_OK) return;
    for (int a = 0; a < m; ++a) rhs[a] = clamped[a] ? 0.0 : grad[a];
    chol_solve_neg(&wm, rhs, 1, dx, 1);           /* Newton step on the free subspace, zero on the clamped one */
    double sdotg = 0.0;
    for (int a = 0; a < m; ++a) sdotg = fma(grad[a], dx[a], sdotg);
    if (!(sdotg < 0.0)) return;
    double step = 1.0;
    int ok = 0;
    for (int ls = 0; ls < QP_MAXLS; ++ls) {
      for (int a = 0; a < m; ++a) xc[a] = clampd(fma(step, dx[a], x[a]), lo[a], hi[a]);
      const double fc = qp_eval(w, g, xc, rc);
      if (fc - f <= QP_ARMIJO * (step * sdotg)) {
        for (int a = 0; a < m; ++a) { x[a] = xc[a]; r[a] = rc[a]; }
        f = fc;
        ok = 1;
        break;
      }
      step = step * 0.5;
    }
    if (!ok) return;
  }
  /* iteration cap: the clamped set of the last accepted point */
  for (int a = 0; a < m; ++a) {
    const double ga = r[a] + g[a];
    clamped[a] = (x[a] <= lo[a] && ga > 0.0) || (x[a] >= hi[a] && ga < 0.0);
  }
}

/* K rows of the free inputs = -Huu_ff^{-1} Hux_f, zero rows for the clamped ones; P' = Hxx + Hux^T K */
static int gains_and_value_box(step_ws* w, const int* clamped, double* K, double* Pn) {
  const int n = w->n, m = w->m;
  step_ws wm;
  if (chol_lower_masked(w, clamped, &wm) != NOC_O_OK) return NOC_O_NOT_PD;
  double rhs[MMAX];
  for (int j = 0; j < n; ++j) {
    for (int a = 0; a < m; ++a) rhs[a] = clamped[a] ? 0.0 : w->Hux[a * n + j];
    chol_solve_neg(&wm, rhs, 1, K + j, n);
  }
  for (int i = 0; i < n; ++i)
    for (int j = i; j < n; ++j) {
      double acc = w->Hxx[i * n + j];
      for (int a = 0; a < m; ++a) acc = fma(w->Hux[a * n + i], K[a * n + j], acc);
      Pn[i * n + j] = acc;
      Pn[j * n + i] = acc;
    }
  return NOC_O_OK;
}

/* test hook: the box QP alone.  Huu[m*m] (lower triangle read), g, lo, hi, x (in: start, out: solution), clamped[m] */
void noc_oracle_box_qp(int m, const double* Huu, const double* g, const double* lo, const double* hi, double* x,
                       int32_t* clamped) {
  step_ws w;
  w.n = 0; w.m = m;
  memcpy(w.Huu, Huu, (size_t)m * m * sizeof(double));
  int cl[MMAX];
  box_qp(&w, g, lo, hi, x, cl);
  for (int a = 0; a < m; ++a) clamped[a] = cl[a];
}

static double quad_form_half(int n, const double* P, const double* x) {
  double c = 0.0;
  for (int i = 0; i < n; ++i) {
    double r = 0.0;
    for (int j = 0; j < n; ++j) r = fma(P[i * n + j], x[j], r);
    c = fma(x[i], r, c);
  }
  return 0.5 * c;
}

static void fill_nan(double* p, size_t cnt) {
  for (size_t i = 0; i < cnt; ++i) p[i] = NAN;
}

int noc_oracle_lqr_sweep(int n, int m, int N, const double* AB, int layout, const double* Q,
                         const double* R, const double* Qf, double* K, double* P0,
                         const double* x0dev, double* cost) {
  step_ws w;
  w.n = n; w.m = m;
  double P[NMAX * NMAX], Pn[NMAX * NMAX], F[NMAX * NMMAX], Kk[MMAX * NMAX];
  const size_t rec = (size_t)n * n + (size_t)n * m;
  memcpy(P, Qf, (size_t)n * n * sizeof(double));
  for (int k = N - 1; k >= 0; --k) {
    unpack_F(n, m, AB + (size_t)k * rec, layout, F);
    form_H(&w, P, F, Q, R, 0.0);
    if (chol_lower(&w) != NOC_O_OK) {
      /* failure rule: every output of this and earlier steps is NaN */
      if (K) fill_nan(K, (size_t)(k + 1) * m * n);
      if (P0) fill_nan(P0, (size_t)n * n);
      if (cost) *cost = NAN;
      return NOC_O_NOT_PD;
    }
    gains_and_value(&w, Kk, Pn);
    if (K) memcpy(K + (size_t)k * m * n, Kk, (size_t)m * n * sizeof(double));
    memcpy(P, Pn, (size_t)n * n * sizeof(double));
  }
  if (P0) memcpy(P0, P, (size_t)n * n * sizeof(double));
  if (cost && x0dev) *cost = quad_form_half(n, P, x0dev);
  return NOC_O_OK;
}

int noc_oracle_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

void noc_oracle_lqr_batch(int n, int m, int N, int64_t batch, const double* AB, int layout,
                          const double* QRQf, int qr_per_instance, const double* x0dev,
                          double* K, double* K0, double* P0, double* cost, int32_t* status,
                          int threads) {
  const size_t rec = (size_t)n * n + (size_t)n * m;
  const size_t qsz = (size_t)2 * n * n + (size_t)m * m;
  (void)threads;
#pragma omp parallel for schedule(static) num_threads(threads > 0 ? threads : noc_oracle_max_threads())
  for (int64_t b = 0; b < batch; ++b) {
    const double* qr = QRQf + (qr_per_instance ? (size_t)b * qsz : 0);
    double* Kb = NULL;
    double* Ktmp = NULL;
    if (K) Kb = K + (size_t)b * N * m * n;
    else if (K0) { Ktmp = (double*)malloc((size_t)N * m * n * sizeof(double)); Kb = Ktmp; }
    double c = NAN;
    const int st = noc_oracle_lqr_sweep(n, m, N, AB + (size_t)b * N * rec, layout, qr, qr + n * n,
                                        qr + n * n + m * m, Kb, P0 ? P0 + (size_t)b * n * n : NULL,
                                        x0dev ? x0dev + (size_t)b * n : NULL, &c);
    if (K0) memcpy(K0 + (size_t)b * m * n, Kb, (size_t)m * n * sizeof(double));
    if (cost) cost[b] = c;
    if (status) status[b] = st;
    free(Ktmp);
  }
}

void noc_oracle_lqr_from_x0_batch(int model, int N, double dt, int64_t batch, const double* x0,
                                  const double* QRQf, double* K0, double* P0, double* cost,
                                  int32_t* status, int threads) {
  int n, m;
  noc_oracle_model_dims(model, &n, &m);
  const size_t rec = (size_t)n * n + (size_t)n * m;
#pragma omp parallel num_threads(threads > 0 ? threads : noc_oracle_max_threads())
  {
    double* xbar = (double*)malloc((size_t)(N + 1) * n * sizeof(double));
    double* ubar = (double*)malloc((size_t)N * m * sizeof(double));
    double* AB = (double*)malloc((size_t)N * rec * sizeof(double));
    double* K = (double*)malloc((size_t)N * m * n * sizeof(double));
    double uh[MMAX];
    noc_oracle_hover_input(model, uh);
    for (int k = 0; k < N; ++k) memcpy(ubar + (size_t)k * m, uh, (size_t)m * sizeof(double));
#pragma omp for schedule(static)
    for (int64_t b = 0; b < batch; ++b) {
      noc_oracle_rollout_open_loop(model, N, dt, x0 + (size_t)b * n, ubar, xbar);
      noc_oracle_linearize_traj(model, N, dt, xbar, ubar, NOC_O_LAYOUT_A_THEN_B, AB);
      double c = NAN;
      const int st = noc_oracle_lqr_sweep(n, m, N, AB, NOC_O_LAYOUT_A_THEN_B, QRQf, QRQf + n * n,
                                          QRQf + n * n + m * m, K, P0 ? P0 + (size_t)b * n * n : NULL,
                                          x0 + (size_t)b * n, &c);
      if (K0) memcpy(K0 + (size_t)b * m * n, K, (size_t)m * n * sizeof(double));
      if (cost) cost[b] = c;
      if (status) status[b] = st;
    }
    free(xbar); free(ubar); free(AB); free(K);
  }
}

/* synthetic A_k,B_k streams for a batch of initial states (hover-thrust nominal), OpenMP over instances */
void noc_oracle_make_ab_batch(int model, int N, double dt, int64_t batch, const double* x0, int layout,
                              double* AB, int threads) {
  int n, m;
  noc_oracle_model_dims(model, &n, &m);
  const size_t rec = (size_t)n * n + (size_t)n * m;
#pragma omp parallel num_threads(threads > 0 ? threads : noc_oracle_max_threads())
  {
    double* xbar = (double*)malloc((size_t)(N + 1) * n * sizeof(double));
    double* ubar = (double*)malloc((size_t)N * m * sizeof(double));
    double uh[MMAX];
    noc_oracle_hover_input(model, uh);
    for (int k = 0; k < N; ++k) memcpy(ubar + (size_t)k * m, uh, (size_t)m * sizeof(double));
#pragma omp for schedule(static)
    for (int64_t b = 0; b < batch; ++b) {
      noc_oracle_rollout_open_loop(model, N, dt, x0 + (size_t)b * n, ubar, xbar);
      noc_oracle_linearize_traj(model, N, dt, xbar, ubar, layout, AB + (size_t)b * N * rec);
    }
    free(xbar); free(ubar);
  }
}

/* ---------------------------------------------------------------- DARE fixed point */
int noc_oracle_dare(int n, int m, const double* A, const double* B, const double* Q, const double* R,
                    double tol, int max_iter, double* P, double* K) {
  step_ws w;
  w.n = n; w.m = m;
  double F[NMAX * NMMAX], Pn[NMAX * NMAX], rec[NMAX * NMMAX];
  pack_ab(n, m, A, B, NOC_O_LAYOUT_ROWS_AB, rec);
  unpack_F(n, m, rec, NOC_O_LAYOUT_ROWS_AB, F);
  memcpy(P, Q, (size_t)n * n * sizeof(double));
  for (int it = 1; it <= max_iter; ++it) {
    form_H(&w, P, F, Q, R, 0.0);
    if (chol_lower(&w) != NOC_O_OK) return -it;
    gains_and_value(&w, K, Pn);
    double dmax = 0.0, pmax = 0.0;
    for (int i = 0; i < n * n; ++i) {
      const double d = fabs(Pn[i] - P[i]);
      const double p = fabs(Pn[i]);
      if (d > dmax) dmax = d;
      if (p > pmax) pmax = p;
    }
    memcpy(P, Pn, (size_t)n * n * sizeof(double));
    if (dmax < tol * pmax) return it;
  }
  return max_iter;
}

/* ---------------------------------------------------------------- SLQ / iLQR */
static void mat_vec_chain(int rows, int cols, const double* M, const double* v, double* out) {
  for (int i = 0; i < rows; ++i) {
    double r = 0.0;
    for (int j = 0; j < cols; ++j) r = fma(M[i * cols + j], v[j], r);
    out[i] = r;
  }
}

static double stage_cost(int n, int m, const double* x, const double* u, const double* xg,
                         const double* uh, const double* Q, const double* R) {
  double ex[NMAX], eu[MMAX];
  for (int i = 0; i < n; ++i) ex[i] = x[i] - xg[i];
  for (int a = 0; a < m; ++a) eu[a] = u[a] - uh[a];
  double qx = 0.0, qu = 0.0;
  for (int i = 0; i < n; ++i) {
    double r = 0.0;
    for (int j = 0; j < n; ++j) r = fma(Q[i * n + j], ex[j], r);
    qx = fma(ex[i], r, qx);
  }
  for (int a = 0; a < m; ++a) {
    double r = 0.0;
    for (int b = 0; b < m; ++b) r = fma(R[a * m + b], eu[b], r);
    qu = fma(eu[a], r, qu);
  }
  return 0.5 * (qx + qu);
}

static double terminal_cost(int n, const double* x, const double* xg, const double* Qf) {
  double ex[NMAX];
  for (int i = 0; i < n; ++i) ex[i] = x[i] - xg[i];
  return quad_form_half(n, Qf, ex);
}

double noc_oracle_traj_cost(int model, int N, const double* x, const double* u, const double* xgoal,
                            const double* QRQf) {
  int n, m;
  noc_oracle_model_dims(model, &n, &m);
  double uh[MMAX];
  noc_oracle_hover_input(model, uh);
  const double *Q = QRQf, *R = QRQf + n * n, *Qf = QRQf + n * n + m * m;
  double J = 0.0;
  for (int k = 0; k < N; ++k) J = J + stage_cost(n, m, x + (size_t)k * n, u + (size_t)k * m, xgoal, uh, Q, R);
  J = J + terminal_cost(n, x + (size_t)N * n, xgoal, Qf);
  return J;
}

int noc_oracle_backward_affine(int model, int N, double dt, const double* xbar, const double* ubar,
                               const double* xgoal, const double* QRQf, double mu, double* K,
                               double* kff, double* dV) {
  return noc_oracle_backward_affine_box(model, N, dt, xbar, ubar, xgoal, QRQf, mu, -INFINITY, INFINITY, K, kff, dV);
}

int noc_oracle_backward_affine_box(int model, int N, double dt, const double* xbar, const double* ubar,
                                   const double* xgoal, const double* QRQf, double mu, double u_min,
                                   double u_max, double* K, double* kff, double* dV) {
  const int boxed = isfinite(u_min) || isfinite(u_max);
  int n, m;
  noc_oracle_model_dims(model, &n, &m);
  const int nm = n + m;
  const double *Q = QRQf, *R = QRQf + n * n, *Qf = QRQf + n * n + m * m;
  step_ws w;
  w.n = n; w.m = m;
  double uh[MMAX];
  noc_oracle_hover_input(model, uh);
  double P[NMAX * NMAX], Pn[NMAX * NMAX], F[NMAX * NMMAX], A[NMAX * NMAX], B[NMAX * MMAX];
  double Vx[NMAX], Vxn[NMAX], ex[NMAX], eu[MMAX], h[NMMAX], rec[NMAX * NMMAX];
  memcpy(P, Qf, (size_t)n * n * sizeof(double));
  for (int i = 0; i < n; ++i) ex[i] = xbar[(size_t)N * n + i] - xgoal[i];
  mat_vec_chain(n, n, Qf, ex, Vx);
  double dV1 = 0.0, dV2 = 0.0;
  for (int k = N - 1; k >= 0; --k) {
    const double* xk = xbar + (size_t)k * n;
    const double* uk = ubar + (size_t)k * m;
    noc_oracle_linearize(model, xk, uk, dt, A, B);
    pack_ab(n, m, A, B, NOC_O_LAYOUT_ROWS_AB, rec);
    unpack_F(n, m, rec, NOC_O_LAYOUT_ROWS_AB, F);
    /* h = [Q ex ; R eu] + F^T Vx */
    for (int i = 0; i < n; ++i) ex[i] = xk[i] - xgoal[i];
    for (int a = 0; a < m; ++a) eu[a] = uk[a] - uh[a];
    mat_vec_chain(n, n, Q, ex, h);
    mat_vec_chain(m, m, R, eu, h + n);
    for (int i = 0; i < nm; ++i) {
      double acc = h[i];
      for (int l = 0; l < n; ++l) acc = fma(F[l * nm + i], Vx[l], acc);
      h[i] = acc;
    }
    form_H(&w, P, F, Q, R, mu);
    if (chol_lower(&w) != NOC_O_OK) return NOC_O_NOT_PD;
    double* Kk = K + (size_t)k * m * n;
    double* kk = kff + (size_t)k * m;
    chol_solve_neg(&w, h + n, 1, kk, 1);
    int inside = 1;
    double lo[MMAX], hi[MMAX];
    if (boxed)
      for (int a = 0; a < m; ++a) {
        lo[a] = u_min - uk[a];
        hi[a] = u_max - uk[a];
        if (!(kk[a] > lo[a] && kk[a] < hi[a])) inside = 0;
      }
    if (inside) {
      gains_and_value(&w, Kk, Pn);
    } else {   /* some input would leave the box: constrained feed-forward, zero gains on the clamped inputs */
      int clamped[MMAX];
      box_qp(&w, h + n, lo, hi, kk, clamped);
      if (gains_and_value_box(&w, clamped, Kk, Pn) != NOC_O_OK) return NOC_O_NOT_PD;
    }
    for (int i = 0; i < n; ++i) {
      double acc = h[i];
      for (int a = 0; a < m; ++a) acc = fma(w.Hux[a * n + i], kk[a], acc);
      Vxn[i] = acc;
    }
    double t1 = 0.0, t2 = 0.0;
    for (int a = 0; a < m; ++a) t1 = fma(kk[a], h[n + a], t1);
    for (int a = 0; a < m; ++a) {
      double r = 0.0;
      for (int b = 0; b < m; ++b) {
        const double hab = (b <= a) ? w.Huu[a * m + b] : w.Huu[b * m + a];
        r = fma(hab, kk[b], r);
      }
      t2 = fma(kk[a], r, t2);
    }
    dV1 = dV1 + t1;
    dV2 = dV2 + 0.5 * t2;
    memcpy(P, Pn, (size_t)n * n * sizeof(double));
    memcpy(Vx, Vxn, (size_t)n * sizeof(double));
  }
  dV[0] = dV1;
  dV[1] = dV2;
  return NOC_O_OK;
}

double noc_oracle_forward(int model, int N, double dt, const double* xbar, const double* ubar,
                          const double* xgoal, const double* QRQf, const double* K, const double* kff,
                          double alpha, double* xnew, double* unew) {
  return noc_oracle_forward_box(model, N, dt, xbar, ubar, xgoal, QRQf, K, kff, alpha, -INFINITY, INFINITY, xnew, unew);
}

/* Forward pass with the input box: every component of u is projected onto [u_min, u_max] after the feedback law.
 * The comparisons are false for NaN, so a diverged rollout stays NaN, and infinite bounds change nothing. */
double noc_oracle_forward_box(int model, int N, double dt, const double* xbar, const double* ubar,
                              const double* xgoal, const double* QRQf, const double* K, const double* kff,
                              double alpha, double u_min, double u_max, double* xnew, double* unew) {
  int n, m;
  noc_oracle_model_dims(model, &n, &m);
  const double *Q = QRQf, *R = QRQf + n * n, *Qf = QRQf + n * n + m * m;
  double uh[MMAX], dx[NMAX];
  noc_oracle_hover_input(model, uh);
  memcpy(xnew, xbar, (size_t)n * sizeof(double));
  double J = 0.0;
  for (int k = 0; k < N; ++k) {
    double* xk = xnew + (size_t)k * n;
    double* uk = unew + (size_t)k * m;
    for (int j = 0; j < n; ++j) dx[j] = xk[j] - xbar[(size_t)k * n + j];
    for (int a = 0; a < m; ++a) {
      double t = 0.0;
      for (int j = 0; j < n; ++j) t = fma(K[(size_t)k * m * n + a * n + j], dx[j], t);
      uk[a] = clampd((ubar[(size_t)k * m + a] + alpha * kff[(size_t)k * m + a]) + t, u_min, u_max);
    }
    J = J + stage_cost(n, m, xk, uk, xgoal, uh, Q, R);
    noc_oracle_step(model, xk, uk, dt, xk + n);
  }
  J = J + terminal_cost(n, xnew + (size_t)N * n, xgoal, Qf);
  return J;
}

/* Adaptive Levenberg-Marquardt schedule (Tassa, Erez, Todorov 2012): delta0 = 2, mu_min = 1e-6.
 * increase: delta <- max(delta0, delta * delta0), mu <- max(mu_min, mu * delta)
 * decrease: delta <- min(1 / delta0, delta / delta0), mu <- mu * delta if that is >= mu_min, else 0 */
static void reg_increase(double* mu, double* delta) {
  const double d = *delta * 2.0;
  *delta = (d > 2.0) ? d : 2.0;
  const double v = *mu * *delta;
  *mu = (v > 1e-6) ? v : 1e-6;
}
static void reg_decrease(double* mu, double* delta) {
  const double d = *delta * 0.5;
  *delta = (d < 0.5) ? d : 0.5;
  const double v = *mu * *delta;
  *mu = (v >= 1e-6) ? v : 0.0;
}

int noc_oracle_slq(const noc_oracle_slq_opts* o, const double* x0, const double* xgoal,
                   const double* QRQf, double* x, double* u, double* K, double* kff, double* cost,
                   int32_t* iters, int32_t* alpha_idx, double* reg) {
  return noc_oracle_slq_warm(o, x0, xgoal, QRQf, NULL, x, u, K, kff, cost, iters, alpha_idx, reg);
}

/* as noc_oracle_slq, starting from the input sequence u_init[N*m] (MPC warm start) instead of hover thrust */
int noc_oracle_slq_warm(const noc_oracle_slq_opts* o, const double* x0, const double* xgoal,
                        const double* QRQf, const double* u_init, double* x, double* u, double* K, double* kff,
                        double* cost, int32_t* iters, int32_t* alpha_idx, double* reg) {
  int n, m;
  noc_oracle_model_dims(o->model, &n, &m);
  const int N = o->N;
  double uh[MMAX];
  noc_oracle_hover_input(o->model, uh);
  double* Kw = K ? K : (double*)malloc((size_t)N * m * n * sizeof(double));
  double* kw = kff ? kff : (double*)malloc((size_t)N * m * sizeof(double));
  double* xn = (double*)malloc((size_t)(N + 1) * n * sizeof(double));
  double* un = (double*)malloc((size_t)N * m * sizeof(double));
  if (u_init) memcpy(u, u_init, (size_t)N * m * sizeof(double));
  else for (int k = 0; k < N; ++k) memcpy(u + (size_t)k * m, uh, (size_t)m * sizeof(double));
  noc_oracle_rollout_open_loop(o->model, N, o->dt, x0, u, x);
  double J = noc_oracle_traj_cost(o->model, N, x, u, xgoal, QRQf);
  double mu = o->reg0;
  double delta = 1.0;   /* adaptive schedule: the factor mu was last scaled by */
  const int adaptive = o->reg_schedule == 1;
  int status = NOC_O_MAX_ITER;
  int it = 0;
  for (int i = 0; i < o->max_iter; ++i) alpha_idx[i] = -1;
  for (it = 0; it < o->max_iter; ++it) {
    double dV[2];
    int bad = 0;
    while (noc_oracle_backward_affine_box(o->model, N, o->dt, x, u, xgoal, QRQf, mu, o->u_min, o->u_max, Kw, kw, dV) != NOC_O_OK) {
      if (adaptive) reg_increase(&mu, &delta);
      else mu = (mu * 10.0 > 1e-6) ? mu * 10.0 : 1e-6;
      if (mu > 1e6) { bad = 1; break; }
    }
    if (bad) { status = NOC_O_REG_FAIL; ++it; break; }
    int accepted = -1;
    double Jn = J;
    double alpha = 1.0;
    for (int j = 0; j < o->n_alpha; ++j) {
      const double Ja = noc_oracle_forward_box(o->model, N, o->dt, x, u, xgoal, QRQf, Kw, kw, alpha, o->u_min, o->u_max, xn, un);
      const double expected = -(alpha * dV[0] + (alpha * alpha) * dV[1]);
      if (J - Ja >= o->armijo * expected) { accepted = j; Jn = Ja; break; }
      alpha = alpha * 0.5;
    }
    if (accepted < 0) {
      if (!adaptive) { status = NOC_O_LS_FAIL; ++it; break; }
      /* adaptive: no step length passed -> raise mu and repeat the iteration from the same nominal (it counts) */
      reg_increase(&mu, &delta);
      if (mu > 1e6) { status = NOC_O_LS_FAIL; ++it; break; }
      alpha_idx[it] = -2;
      continue;
    }
    alpha_idx[it] = accepted;
    if (adaptive) reg_decrease(&mu, &delta);
    memcpy(x, xn, (size_t)(N + 1) * n * sizeof(double));
    memcpy(u, un, (size_t)N * m * sizeof(double));
    const double denom = (J > 1e-12) ? J : 1e-12;
    const double rel = (J - Jn) / denom;
    J = Jn;
    if (rel < o->tol) { status = NOC_O_OK; ++it; break; }
  }
  *iters = it;
  *cost = J;
  if (reg) *reg = mu;
  if (!K) free(Kw);
  if (!kff) free(kw);
  free(xn); free(un);
  return status;
}

void noc_oracle_slq_batch(const noc_oracle_slq_opts* o, int64_t batch, const double* x0,
                          const double* xgoal, const double* QRQf, double* x, double* u, double* K,
                          double* cost, int32_t* iters, int32_t* alpha_idx, int32_t* status,
                          int threads) {
  noc_oracle_slq_batch_warm(o, batch, x0, xgoal, QRQf, NULL, x, u, K, cost, iters, alpha_idx, status, threads);
}

void noc_oracle_slq_batch_warm(const noc_oracle_slq_opts* o, int64_t batch, const double* x0,
                               const double* xgoal, const double* QRQf, const double* u_init, double* x, double* u,
                               double* K, double* cost, int32_t* iters, int32_t* alpha_idx, int32_t* status,
                               int threads) {
  int n, m;
  noc_oracle_model_dims(o->model, &n, &m);
  const int N = o->N;
#pragma omp parallel for schedule(dynamic, 4) num_threads(threads > 0 ? threads : noc_oracle_max_threads())
  for (int64_t b = 0; b < batch; ++b) {
    double* xb = x ? x + (size_t)b * (N + 1) * n : (double*)malloc((size_t)(N + 1) * n * sizeof(double));
    double* ub = u ? u + (size_t)b * N * m : (double*)malloc((size_t)N * m * sizeof(double));
    int32_t* ai = alpha_idx ? alpha_idx + (size_t)b * o->max_iter : (int32_t*)malloc((size_t)o->max_iter * sizeof(int32_t));
    double c; int32_t itn;
    const int st = noc_oracle_slq_warm(o, x0 + (size_t)b * n, xgoal + (size_t)b * n, QRQf,
                                       u_init ? u_init + (size_t)b * N * m : NULL, xb, ub,
                                       K ? K + (size_t)b * N * m * n : NULL, NULL, &c, &itn, ai, NULL);
    if (cost) cost[b] = c;
    if (iters) iters[b] = itn;
    if (status) status[b] = st;
    if (!x) free(xb);
    if (!u) free(ub);
    if (!alpha_idx) free(ai);
  }
}
