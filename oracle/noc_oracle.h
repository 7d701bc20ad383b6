/*
 * noc_oracle.h — CPU oracle for the batched LQR / SLQ hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * PARITY UNPINNED: the mounted reference (/root/reference, KahnShi/NOC) holds no solver source,
 * tests or golden vectors (four catkin boilerplate files; every target in
 * lqr_control/CMakeLists.txt:124-126, :136, :193 is commented out).  This oracle therefore restates
 * the specification frozen in DESIGN.md §2 (derived from SURVEY.md App. A), not reference code.
 * It is pinned instead against independent math (tests/test_oracle_*.py): scipy's
 * solve_discrete_are, a numpy longdouble Riccati sweep, closed-form scalar / double-integrator
 * answers, finite-difference Jacobians and libm sin/cos.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load
 * this library.  The shipped library (libnoc_b200.so) never links or calls it.
 *
 * All matrices are row-major fp64.  Operation order is part of the spec: every inner product is
 * a sequential fma chain over ascending index, so the CUDA kernels can match bit for bit.
 */
#ifndef NOC_ORACLE_H
#define NOC_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

enum { NOC_O_MODEL_QUAD12 = 0, NOC_O_MODEL_DUAL24 = 1 };
enum { NOC_O_LAYOUT_A_THEN_B = 0, NOC_O_LAYOUT_ROWS_AB = 1 };
/* per-instance status (same values as include/lqr_control/noc_b200.h) */
enum { NOC_O_OK = 0, NOC_O_NOT_PD = 1, NOC_O_MAX_ITER = 2, NOC_O_LS_FAIL = 3, NOC_O_REG_FAIL = 4 };

/* -- elementary functions ------------------------------------------------------------------ */
void noc_oracle_sincos(double x, double* s, double* c);

/* -- models ---------------------------------------------------------- */
int  noc_oracle_model_dims(int model, int* n, int* m);
void noc_oracle_hover_input(int model, double* u_h);
/* x_next = x + dt * f(x,u) */
void noc_oracle_step(int model, const double* x, const double* u, double dt, double* x_next);
/* A = I + dt*df/dx (n*n), B = dt*df/du (n*m), dense row-major */
void noc_oracle_linearize(int model, const double* x, const double* u, double dt, double* A, double* B);
/* xbar[(N+1)*n] from x0 and ubar[N*m] */
void noc_oracle_rollout_open_loop(int model, int N, double dt, const double* x0, const double* ubar,
                                  double* xbar);
/* AB[N][n*n+n*m] (layout 0: A then B; layout 1: 12 rows of [A row | B row]) along a trajectory */
void noc_oracle_linearize_traj(int model, int N, double dt, const double* xbar, const double* ubar,
                               int layout, double* AB);

/* -- finite-horizon LTV LQR sweep ------------------------------------------ */
/* K[N][m*n] (may be NULL), P0[n*n] (may be NULL), returns status. x0dev (may be NULL) -> *cost */
int noc_oracle_lqr_sweep(int n, int m, int N, const double* AB, int layout, const double* Q,
                         const double* R, const double* Qf, double* K, double* P0,
                         const double* x0dev, double* cost);
/* batch driver, OpenMP over instances; qr_per_instance: QRQf is [B][n*n+m*m+n*n] else shared.
 * ab_stride_instances: 0 => the same AB for every instance. K0[B][m*n] = gains of step 0. */
void noc_oracle_lqr_batch(int n, int m, int N, int64_t batch, const double* AB, int layout,
                          const double* QRQf, int qr_per_instance, const double* x0dev,
                          double* K, double* K0, double* P0, double* cost, int32_t* status,
                          int threads);
/* config-2/4 pipeline from initial states: hover-thrust nominal rollout -> linearise -> sweep.
 * x0[B][n]; x0dev = x0 - xgoal (xgoal may be NULL = origin). Scratch is internal per thread. */
void noc_oracle_lqr_from_x0_batch(int model, int N, double dt, int64_t batch, const double* x0,
                                  const double* QRQf, double* K0, double* P0, double* cost,
                                  int32_t* status, int threads);

/* AB[B][N][rec] along the hover-thrust nominal from each x0[B][n] (bench / test input generator) */
void noc_oracle_make_ab_batch(int model, int N, double dt, int64_t batch, const double* x0, int layout,
                              double* AB, int threads);

/* -- infinite horizon ------------------------------------------------------ */
/* iterate the sweep step from P=Q until max|P'-P|/max|P'| < tol. returns iterations used (<= max_iter),
 * negative on Cholesky failure. */
int noc_oracle_dare(int n, int m, const double* A, const double* B, const double* Q, const double* R,
                    double tol, int max_iter, double* P, double* K);

/* -- SLQ / iLQR ------------------------------------------------------------ */
typedef struct {
  int model, N, max_iter, n_alpha;
  double dt, tol, reg0, armijo;
  double u_min, u_max;   /* box on every input component; -INFINITY / INFINITY = unconstrained */
  int reg_schedule;      /* 0: mu x10 on a failed factorisation, never decreased; 1: adaptive (2.4c) */
} noc_oracle_slq_opts;
/* x[(N+1)*n], u[N*m], K[N*m*n] (may be NULL), kff[N*m] (may be NULL), alpha_idx[max_iter] (filled with -1
 * beyond iters), returns status; *iters, *cost, *reg (final mu) */
int noc_oracle_slq(const noc_oracle_slq_opts* o, const double* x0, const double* xgoal,
                   const double* QRQf, double* x, double* u, double* K, double* kff, double* cost,
                   int32_t* iters, int32_t* alpha_idx, double* reg);
void noc_oracle_slq_batch(const noc_oracle_slq_opts* o, int64_t batch, const double* x0,
                          const double* xgoal, const double* QRQf, double* x, double* u, double* K,
                          double* cost, int32_t* iters, int32_t* alpha_idx, int32_t* status,
                          int threads);
int noc_oracle_slq_warm(const noc_oracle_slq_opts* o, const double* x0, const double* xgoal,
                        const double* QRQf, const double* u_init, double* x, double* u, double* K, double* kff,
                        double* cost, int32_t* iters, int32_t* alpha_idx, double* reg);
void noc_oracle_slq_batch_warm(const noc_oracle_slq_opts* o, int64_t batch, const double* x0,
                               const double* xgoal, const double* QRQf, const double* u_init, double* x, double* u,
                               double* K, double* cost, int32_t* iters, int32_t* alpha_idx, int32_t* status,
                               int threads);
/* one backward pass (affine) for unit tests: returns 0 or NOT_PD */
int noc_oracle_backward_affine(int model, int N, double dt, const double* xbar, const double* ubar,
                               const double* xgoal, const double* QRQf, double mu, double* K,
                               double* kff, double* dV /*[2]*/);
/* one forward pass with step alpha; returns cost */
double noc_oracle_forward(int model, int N, double dt, const double* xbar, const double* ubar,
                          const double* xgoal, const double* QRQf, const double* K, const double* kff,
                          double alpha, double* xnew, double* unew);
int noc_oracle_backward_affine_box(int model, int N, double dt, const double* xbar, const double* ubar,
                                   const double* xgoal, const double* QRQf, double mu, double u_min,
                                   double u_max, double* K, double* kff, double* dV);
double noc_oracle_forward_box(int model, int N, double dt, const double* xbar, const double* ubar,
                              const double* xgoal, const double* QRQf, const double* K, const double* kff,
                              double alpha, double u_min, double u_max, double* xnew, double* unew);
void noc_oracle_box_qp(int m, const double* Huu, const double* g, const double* lo, const double* hi, double* x,
                       int32_t* clamped);
double noc_oracle_traj_cost(int model, int N, const double* x, const double* u, const double* xgoal,
                            const double* QRQf);

int noc_oracle_max_threads(void);
#ifdef __cplusplus
}
#endif
#endif
