/*
 * lqr_control/noc_b200.h — C-ABI of libnoc_b200.so, the batched LQR / SLQ solver for NVIDIA B200.
 *
 * Drop-in boundary for the hot path of KahnShi/NOC's `lqr_control` package.  The reference snapshot
 * exports nothing (catkin_package() has no INCLUDE_DIRS/LIBRARIES, lqr_control/CMakeLists.txt:105-110)
 * and defines no target; the only fixed names are the slots of its CMake template:
 *     library  ${PROJECT_NAME}            lqr_control/CMakeLists.txt:124-126  -> this library
 *     headers  include/${PROJECT_NAME}/   lqr_control/CMakeLists.txt:175-179  -> this directory
 *     node     ${PROJECT_NAME}_node       lqr_control/CMakeLists.txt:136      -> out of scope (no ROS)
 * Every entry point below says which SURVEY.md §8(a) row it implements; there is no reference
 * function to cite because none exists.  Arithmetic is specified in DESIGN.md §2.
 *
 * Conventions
 *   - extern "C", plain pointers and sizes; no C++ types, no exceptions cross this boundary.
 *   - every data pointer may be a HOST pointer (pageable or pinned) or a DEVICE pointer of the
 *     context's GPU; the library detects which (cudaPointerGetAttributes).  Device pointers are used
 *     in place (no copy); host pointers are staged through library-owned device buffers.
 *   - all matrices are row-major fp64.  Optional outputs may be NULL.
 *   - return value: 0 = NOC_OK, negative = noc_error (call failed, see noc_last_error()).
 *     Per-instance numerical outcomes never fail the call; they go to status[] (noc_status).
 *   - a noc_ctx is bound to one GPU and is single-caller; distinct contexts are independent.
 *     Calls are synchronous on return unless NOC_FLAG_ASYNC is set (device pointers only).
 *   - there is NO CPU fallback: without a usable CUDA device noc_create() fails with
 *     NOC_ERR_NO_DEVICE and nothing else can be called.
 */
#ifndef LQR_CONTROL_NOC_B200_H
#define LQR_CONTROL_NOC_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NOC_B200_VERSION_MAJOR 0
#define NOC_B200_VERSION_MINOR 1

typedef struct noc_ctx noc_ctx; /* opaque: owns streams, scratch device buffers */

typedef enum {
  NOC_OK = 0,
  NOC_ERR_BAD_ARG = -1,     /* NULL required pointer, bad sizes, unknown enum */
  NOC_ERR_UNSUPPORTED = -2, /* (n,m) / model / layout combination without a kernel */
  NOC_ERR_CUDA = -3,        /* a CUDA runtime call failed; text in noc_last_error() */
  NOC_ERR_NO_DEVICE = -4,   /* no CUDA device / not sm_100: there is no CPU fallback */
  NOC_ERR_OOM = -5,         /* device allocation failed */
  NOC_ERR_NCCL = -6         /* NCCL could not be loaded or a collective failed (multi-GPU plane only) */
} noc_error;

/* per-instance outcome, written to status[] */
typedef enum {
  NOC_STATUS_OK = 0,        /* LQR: sweep completed.  SLQ: converged (relative decrease < tol) */
  NOC_STATUS_NOT_PD = 1,    /* LQR: R + B'PB not positive definite at some step; outputs are NaN */
  NOC_STATUS_MAX_ITER = 2,  /* SLQ: max_iter reached */
  NOC_STATUS_LS_FAIL = 3,   /* SLQ: no step length in the alpha schedule passed the Armijo test */
  NOC_STATUS_REG_FAIL = 4   /* SLQ: regularisation exceeded 1e6 */
} noc_status;

typedef enum {
  NOC_MODEL_QUAD12 = 0,   /* 12-state / 4-input quadrotor, DESIGN.md §2.1 */
  NOC_MODEL_DUAL24 = 1,   /* two spring-coupled quadrotors, 24 / 8, DESIGN.md §2.6 */
  NOC_MODEL_LTV_USER = 2  /* caller supplies A_k, B_k (noc_lqr_solve_batch / noc_dare_solve_batch only) */
} noc_model;

typedef enum {
  NOC_LAYOUT_A_THEN_B = 0, /* record k = A_k (n*n, row-major) followed by B_k (n*m, row-major) */
  NOC_LAYOUT_ROWS_AB = 1   /* record k = n rows of [A_k row i (n) | B_k row i (m)]  ("F = [A B]") */
} noc_layout;

enum {
  NOC_FLAG_NONE = 0,
  NOC_FLAG_ASYNC = 1,      /* do not synchronise before returning (all pointers must be device pointers) */
  NOC_FLAG_TIME_PARALLEL = 2 /* noc_lqr_solve_batch / noc_lqr_solve_from_x0 / noc_lqr_solve_along, n=12, m=4, A_THEN_B records:
                              time-parallel Riccati sweep — the horizon is cut into chunks that are swept
                              concurrently and glued by an associative scan: the latency path for SMALL batches (a call is
                              ~N/6 dependent steps deep instead of N; one CTA per instance).  Different association of the
                              arithmetic: results agree with the sequential sweep to ~1e-12 relative, not bitwise. */
};

typedef struct {
  int32_t n, m, N;         /* state dim, input dim, horizon (steps) */
  int32_t model;           /* noc_model */
  int32_t layout;          /* noc_layout of AB records */
  int32_t max_iter;        /* SLQ iterations / DARE iterations cap */
  int32_t n_alpha;         /* SLQ: number of step lengths alpha_j = 2^-j tried */
  int32_t flags;           /* NOC_FLAG_* */
  double dt;               /* Euler step of the built-in models */
  double tol;              /* SLQ: relative cost decrease; DARE: max|dP|/max|P| */
  double reg0;             /* SLQ: initial Levenberg-Marquardt mu added to the diagonal of Quu */
  double armijo;           /* SLQ: sufficient-decrease constant (1e-4) */
  double u_min, u_max;     /* SLQ: box on every input component (control-limited DDP: box QP for the feed-forward,
                              zero feedback on clamped inputs, clamped rollouts).  -INFINITY / +INFINITY
                              (the defaults) = unconstrained.  Both built-in models. */
  int32_t reg_schedule;    /* SLQ: noc_reg_schedule */
  int32_t reserved_;       /* keep 0 */
} noc_problem_t;

/* SLQ regularisation schedules (the Levenberg-Marquardt mu added to the diagonal of Quu) */
typedef enum {
  NOC_REG_X10 = 0,       /* default: mu <- max(10 mu, 1e-6) when a factorisation fails, never decreased; a failed line
                            search ends the instance with NOC_STATUS_LS_FAIL */
  NOC_REG_ADAPTIVE = 1   /* Tassa-style: mu <- max(1e-6, mu*delta), delta <- max(2, 2 delta) on a failed factorisation AND
                            on a failed line search (the iteration is repeated from the same nominal and recorded as
                            alpha_idx = -2); delta <- min(1/2, delta/2), mu <- mu*delta (0 below 1e-6) after every
                            accepted step.  mu > 1e6 ends the instance (REG_FAIL / LS_FAIL) */
} noc_reg_schedule;

/* ---- context ------------------------------------------------------------------------------ */
int noc_create(noc_ctx** out, int device_id);
void noc_destroy(noc_ctx* ctx);
const char* noc_last_error(const noc_ctx* ctx);       /* never NULL */
const char* noc_version(void);
/* Launch on a caller-owned CUDA stream (cudaStream_t cast to void*); NULL restores the context's own. */
int noc_set_stream(noc_ctx* ctx, void* cuda_stream);
int noc_synchronize(noc_ctx* ctx);
/* Scheduler options of a context (defaults 0 = automatic).  They change how a batch is cut into chunks, never a result. */
typedef enum {
  NOC_OPT_SCRATCH_LIMIT_BYTES = 0, /* cap on the library-owned scratch one chunk of a call may use (0: 60 % of the free HBM);
                                      batches that need more are processed in several outer chunks */
  NOC_OPT_CHUNK_WAVES = 1,         /* from-x0 pipeline with host outputs: inner chunk = this many waves of the sweep kernel */
  NOC_OPT_SINGLE_STREAM = 2,       /* 1: inner chunks on one stream (default: two alternating streams) */
  NOC_OPT_TIME_PARALLEL_DFMA = 3,  /* 1: NOC_FLAG_TIME_PARALLEL sweeps form their 12x12x16 products with DFMA chains instead of
                                      FP64 tensor-core tiles (DMMA) — the A/B measurement of DESIGN.md 3.1d */
  NOC_OPT_TIME_PARALLEL_CHUNKS = 4,/* > 0: force that many horizon chunks (1 = plain warp-per-instance recursion); 0 = planned */
  NOC_OPT_NO_SMALL_BATCH_KERNEL = 5, /* 1: noc_lqr_solve_batch keeps the 8-instances-per-warp sweep also for batches of at most two
                                      instances per SM (default: those run one instance per warp, bit-identical, ~1.6x sooner) */
  NOC_OPT_NO_TENSOR_SWEEP = 6,     /* 1: the sweeps that form their products on FP64 tensor-core tiles fall back to DFMA chains — the
                                      n=12 LTV sweep over A_THEN_B records (k_ric12_mma -> k_ric12) and every n=24 path over A_THEN_B
                                      records or the built-in dual model: LTV sweep, from-x0, SLQ backward pass, DARE (k_ric24_mma ->
                                      k_ric24 / k_ric_generic).  All are bit-identical to the oracle — DMMA evaluates the same
                                      ascending fma chain (tools/ubench/dmma_order.cu) — this is the A/B switch */
  NOC_OPT_TENSOR_SWEEP_VARIANT = 7, /* measurements: 0 = default (records staged in shared memory by TMA, eight free-running warps per
                                      SM), 1 = + phase interlock of the two warps of a scheduler, 2 = no staging: fragments read
                                      straight from L2, twelve warps per SM.  All bit-identical; 0 is the fastest */
  NOC_OPT_NO_WAVE_PLAN = 8,        /* 1: SLQ backward sweeps use the fixed launch shapes of round 1 (seven warps per CTA above two
                                      instances-slots per SM, else one) instead of the balanced wave plan */
  NOC_OPT_DEFERRED_BACKTRACK = 9,  /* 1: the SLQ line search settles the rejected instances beside the next backward sweep, on the
                                      auxiliary stream, and they rejoin the work list one sweep later (same results; measured slower
                                      on the BASELINE configs — the extra sweeps of the stragglers cost more) */
  NOC_OPT_NO_SPECULATIVE_BACKTRACK = 10, /* 1: the SLQ line search tries alpha = 1 first for EVERY instance and backtracks the rejected
                                      ones in a second launch set.  Default: the (few) instances that needed a shorter step in their
                                      previous iteration run all their step lengths as candidates beside the others' first trial,
                                      on the auxiliary stream — same accepted step, one latency-bound launch set less per iteration
                                       */
  NOC_OPT_COUNT = 11
} noc_option;
int noc_set_option(noc_ctx* ctx, int option, int64_t value);
/* Device memory for callers that do not link the CUDA runtime themselves (the C++ facade, cgo / ctypes bindings):
 * buffers on the context's GPU, and a stream-ordered copy in any direction (host<->device, device<->device) on the
 * context's stream.  noc_memcpy is asynchronous: noc_synchronize() before reading a host destination.            */
int noc_device_alloc(noc_ctx* ctx, uint64_t bytes, void** out);
int noc_device_free(noc_ctx* ctx, void* p);
int noc_memcpy(noc_ctx* ctx, void* dst, const void* src, uint64_t bytes);
/* kernels launched by this context since creation (bench.py's gpu_launches) */
int64_t noc_launch_count(const noc_ctx* ctx);
/* Per-kernel device time, measured with CUDA events on the launching stream (bench.py's roofline).
 * Off by default; while on, every launch of the library is bracketed by an event pair. */
typedef enum {
  NOC_KERNEL_SETUP = 0,      /* weight packing */
  NOC_KERNEL_ROLLOUT = 1,    /* open-loop nominal rollout */
  NOC_KERNEL_LINEARIZE = 2,  /* time-parallel A_k,B_k */
  NOC_KERNEL_RICCATI = 3,    /* backward sweep (LQR, DARE and the SLQ backward pass) */
  NOC_KERNEL_FORWARD = 4,    /* SLQ closed-loop rollouts + line search */
  NOC_KERNEL_ARGMIN = 5,
  NOC_KERNEL_COUNT = 6
} noc_kernel_id;
int noc_profile_enable(noc_ctx* ctx, int on);
/* total device milliseconds and launch count of one kernel id since the last reset (synchronises) */
int noc_profile_read(noc_ctx* ctx, int kernel_id, double* total_ms, int64_t* launches);
int noc_profile_reset(noc_ctx* ctx);
/* fill defaults: tol 1e-8, max_iter 50, n_alpha 11, armijo 1e-4, dt 0.02, layout A_THEN_B, no input box, NOC_REG_X10 */
void noc_problem_defaults(noc_problem_t* p, int32_t model, int32_t N);

/* ---- SURVEY §8(a) a2+a3: backward Riccati recursion, Cholesky, gains -------------------------
 * Finite-horizon LTV LQR sweep for `batch` independent instances.
 *   AB        [batch][N][n*n+n*m]   A_k,B_k records (prob->layout); any 1<=n<=24, 1<=m<=8 ((12,4) and (24,8) are the tuned shapes); a device pointer
 *                                   must be 16-byte aligned (the records are fetched by TMA bulk copies)
 *   QRQf      [n*n + m*m + n*n]     Q, R, Qf shared by the batch, or [batch][...] if qr_per_instance ((12,4) and
 *                                   (24,8) with A_THEN_B records: the tensor-core sweeps with each instance's own
 *                                   weight fragments; other shapes / layouts: the thread-per-instance kernel).
 *                                   Read: Q[i][j] and Qf[i][j] for i <= j, R[a][b] for b <= a
 *   x0dev     [batch][n] or NULL    initial deviation; needed only for `cost`
 *   K         [batch][N][m*n] or NULL   feedback gains u = K x
 *   K0        [batch][m*n]   or NULL   gains of step 0 only (MPC use)
 *   P0        [batch][n*n]   or NULL   cost-to-go matrix at step 0
 *   cost      [batch]        or NULL   0.5 x0' P0 x0
 *   status    [batch]        or NULL   noc_status                                                  */
int noc_lqr_solve_batch(noc_ctx* ctx, const noc_problem_t* prob, int64_t batch, const double* AB,
                        const double* QRQf, int qr_per_instance, const double* x0dev, double* K,
                        double* K0, double* P0, double* cost, int32_t* status);

/* ---- time-parallel linearisation ---------------------------------------------
 *   xbar [batch][N+1][n], ubar [batch][N][m] (NULL = hover input everywhere) -> AB [batch][N][rec]     */
int noc_linearize_batch(noc_ctx* ctx, const noc_problem_t* prob, int64_t batch, const double* xbar,
                        const double* ubar, double* AB);
/* open-loop nominal rollout x_{k+1} = x_k + dt f(x_k, u_k): x0 [batch][n] -> xbar [batch][N+1][n]     */
int noc_rollout_open_loop(noc_ctx* ctx, const noc_problem_t* prob, int64_t batch, const double* x0,
                          const double* ubar, double* xbar);

/* ---- configs 2 / 4 end to end: x0 -> hover-thrust nominal -> linearise -> sweep -----------------
 * Same outputs as noc_lqr_solve_batch; x0dev = x0 (goal = origin).  The linearisation is fused into the
 * sweep: no A_k,B_k stream is materialised, only the nominal trajectory is staged in library scratch (chunked
 * to the free HBM).  With host output pointers and batch >= 8192 the batch is processed in chunks whose
 * device->host copies overlap the next chunk's kernels.  Device output pointers must be 16-byte aligned.  */
int noc_lqr_solve_from_x0(noc_ctx* ctx, const noc_problem_t* prob, int64_t batch, const double* x0,
                          const double* QRQf, double* K, double* K0, double* P0, double* cost,
                          int32_t* status);

/* Time-varying LQR along a nominal input sequence (tracking controller around a planned trajectory): as
 * noc_lqr_solve_from_x0, but the nominal is the open-loop rollout of ubar [batch][N][m] (NULL = hover thrust) from x0,
 * and it is returned in xbar [batch][N+1][n] if that pointer is given.  Linearisation is fused into the sweep.     */
int noc_lqr_solve_along(noc_ctx* ctx, const noc_problem_t* prob, int64_t batch, const double* x0, const double* ubar,
                        const double* QRQf, double* xbar, double* K, double* K0, double* P0, double* cost,
                        int32_t* status);

/* ---- infinite-horizon LQR by fixed-point iteration ----------------------------
 *   AB [batch][n*n+n*m] one LTI record per instance; QR [n*n+m*m] shared.
 *   Iterates the sweep step from P = Q until max|P'-P| < tol * max|P'| or max_iter.
 *   P [batch][n*n], K [batch][m*n], iters [batch], status [batch] (MAX_ITER if not converged)
 *   Any 1 <= n <= 24, 1 <= m <= 8 (NOC_MODEL_LTV_USER); (12,4) and (24,8) with A_THEN_B records run tuned kernels
 *   (k_ric12<RIC_DARE>; k_ric24_mma<RIC_DARE> on the FP64 tensor cores), everything else the thread-per-instance one.   */
int noc_dare_solve_batch(noc_ctx* ctx, const noc_problem_t* prob, int64_t batch, const double* AB,
                         const double* QR, double* P, double* K, int32_t* iters, int32_t* status);

/* The same iteration for the built-in quadrotor model linearised at given operating points (BASELINE.json config 1:
 * "hover linearisation"): xbar [batch][12], ubar [batch][4] or NULL (= hover thrust).  The record is built in the
 * kernel and the structural zeros of the Jacobian are skipped, so no AB array exists; results are bit-identical to
 * noc_linearize_batch + noc_dare_solve_batch.  prob->model must be NOC_MODEL_QUAD12.                            */
int noc_dare_solve_at(noc_ctx* ctx, const noc_problem_t* prob, int64_t batch, const double* xbar, const double* ubar,
                      const double* QR, double* P, double* K, int32_t* iters, int32_t* status);

/* ---- SLQ / iLQR with the built-in nonlinear models ---------------------------
 *   x0, xgoal [batch][n]; QRQf shared.
 *   x [batch][N+1][n], u [batch][N][m], K [batch][N][m*n] (NULL ok), cost [batch],
 *   iters [batch], alpha_idx [batch][max_iter] (accepted j per iteration, -1 beyond; -2 = iteration rejected and
 *   regularisation raised, NOC_REG_ADAPTIVE only), status [batch]                                     */
int noc_slq_solve_batch(noc_ctx* ctx, const noc_problem_t* prob, int64_t batch, const double* x0,
                        const double* xgoal, const double* QRQf, double* x, double* u, double* K,
                        double* cost, int32_t* iters, int32_t* alpha_idx, int32_t* status);
/* The same, started from the input sequence u_init [batch][N][m] instead of hover thrust (receding-horizon MPC: the
 * previous solution shifted by one step).  u_init may be the `u` output array itself.  NULL = hover thrust.        */
int noc_slq_solve_warm(noc_ctx* ctx, const noc_problem_t* prob, int64_t batch, const double* x0, const double* xgoal,
                       const double* QRQf, const double* u_init, double* x, double* u, double* K, double* cost,
                       int32_t* iters, int32_t* alpha_idx, int32_t* status);

/* ---- "pick the best rollout" on this GPU: index and value of the smallest finite cost ------------
 * Multi-CTA reduction; NaN / inf costs (failed instances) are skipped, ties go to the smallest index, no finite
 * cost -> index -1, cost NaN.  The host variant synchronises; the _async variant leaves {double cost; int64 index}
 * (16 bytes, `noc_cost_index_t`) in DEVICE memory at best_dev and returns without a host round trip.           */
typedef struct { double cost; int64_t index; } noc_cost_index_t;
int noc_argmin_cost(noc_ctx* ctx, const double* cost, int64_t batch, int64_t* best_index,
                    double* best_cost);
int noc_argmin_cost_async(noc_ctx* ctx, const double* cost_dev, int64_t batch, int64_t index_offset,
                          noc_cost_index_t* best_dev);

/* ---- the multi-GPU plane (shard, gather, pick the best) ----------------------------------
 * The batch shards over the GPUs of one box with no data-path collective; NCCL over NVLink is used only to gather
 * per-instance costs and gains and to pick the best rollout.  A noc_comm is one NCCL rank bound to one noc_ctx
 * (collectives run on the context's stream).  NCCL is loaded at run time (dlopen "libnccl.so.2"): the library has no
 * link-time NCCL dependency, and inside a PyTorch process it shares torch's NCCL.  All pointers are DEVICE pointers
 * of the comm's GPU; every rank contributes the same count.  Calls are asynchronous on the context's stream unless
 * a host result pointer is given.
 *   one process, several GPUs:  noc_comm_create_all(comms, ctxs, n)   (ncclCommInitAll; then one host thread per
 *                               comm, or any thread between noc_comm_group_begin / _end)
 *   one process per GPU:        rank 0 calls noc_comm_unique_id(), the launcher distributes the 128 bytes
 *                               (torch.distributed / MPI / a file), every rank calls noc_comm_create_rank()    */
typedef struct noc_comm noc_comm;
#define NOC_COMM_ID_BYTES 128
int noc_comm_unique_id(void* id128);
int noc_comm_create_rank(noc_comm** out, noc_ctx* ctx, const void* id128, int rank, int nranks);
int noc_comm_create_all(noc_comm** out, noc_ctx* const* ctxs, int n);
void noc_comm_destroy(noc_comm* comm);
int noc_comm_rank(const noc_comm* comm);
int noc_comm_size(const noc_comm* comm);
const char* noc_comm_last_error(const noc_comm* comm);
int noc_comm_group_begin(void);   /* ncclGroupStart / ncclGroupEnd for single-thread multi-GPU callers */
int noc_comm_group_end(void);
/* all-gather of `count` elements of `elem_bytes` each per rank: dst [nranks][count] on every rank (costs: elem 8;
 * first-step gains K0: elem 8*m*n; whole gain sequences: elem 8*N*m*n)                                        */
int noc_comm_allgather(noc_comm* comm, const void* src_dev, void* dst_dev, int64_t count, int64_t elem_bytes);
/* pick the best rollout over ALL ranks: multi-CTA local argmin (global index = index_offset + local index),
 * all-gather of the ranks' 16-byte (cost, index) pairs, final argmin.  best_dev (device, 16 B) receives the winner on
 * every rank; best_host (may be NULL) additionally receives it on the host, which synchronises the stream.     */
int noc_comm_best_rollout(noc_comm* comm, const double* cost_dev, int64_t count, int64_t index_offset,
                          noc_cost_index_t* best_dev, noc_cost_index_t* best_host);
/* Sharded configs 2 / 4 in one call: this rank's slice x0 [count][n] -> nominal -> fused linearise + sweep in
 * wave-sized chunks, and while chunk i+1 is swept, chunk i's costs and first-step gains are gathered over NVLink
 * on a second stream into cost_all [nranks][count] and K0_all [nranks][count][m*n] (either may be NULL) — the
 * gather hides under the sweep instead of following it.  Then the best rollout as above.  status [count] is local. */
int noc_comm_lqr_solve_from_x0(noc_comm* comm, const noc_problem_t* prob, int64_t count, const double* x0_dev,
                               const double* QRQf_dev, double* cost_all_dev, double* K0_all_dev,
                               int32_t* status_dev, noc_cost_index_t* best_dev, noc_cost_index_t* best_host);

/* The chunk plan a NOC_FLAG_TIME_PARALLEL sweep of horizon N uses: `chunks` warps per instance, chunks 0 .. chunks-2 take
 * `len` steps each, the last one the remaining N - (chunks-1)*len >= 1 (host-side function, needs no GPU).             */
int noc_plan_time_parallel(int32_t N, int32_t* chunks, int32_t* len);

/* ---- diagnostics -----------------------------------------------------------------------------------
 * The pivots' reciprocals inside the sweep use a branch-free correctly-rounded sequence valid on
 * (1e-280, 1e280); this runs it next to the IEEE division on n values so tests can compare the bits. */
int noc_selftest_rcp(noc_ctx* ctx, const double* x, double* y_fast, double* y_ieee, int64_t n);

#ifdef __cplusplus
}
#endif
#endif /* LQR_CONTROL_NOC_B200_H */
