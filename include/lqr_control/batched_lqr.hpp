// lqr_control/batched_lqr.hpp — C++ host-side facade over the C-ABI of libnoc_b200.so.
//
// This is the "controller class" slot of the reference package: KahnShi/NOC's lqr_control scaffolds a library
// built from src/${PROJECT_NAME}/lqr_control.cpp with headers in include/${PROJECT_NAME}/
// (lqr_control/CMakeLists.txt:124-126, :175-179) and defines nothing in it, so the names
// below are this repo's.  Header-only C++17; the only thing that crosses the library boundary is the C-ABI in
// noc_b200.h (exceptions are thrown on this side of it).
//
// Pointers follow the C-ABI rule: host or device, detected by the library; outputs may be nullptr.
#pragma once
#include <algorithm>
#include <cstdint>
#include <memory>
#include <stdexcept>
#include <string>
#include <thread>
#include <utility>
#include <vector>

#include "noc_b200.h"

namespace lqr_control {

class NocError : public std::runtime_error {
 public:
  NocError(int code, const std::string& what) : std::runtime_error(what), code_(code) {}
  int code() const { return code_; }

 private:
  int code_;
};

// One GPU context (noc_ctx).  Single caller; create one per device / per host thread.
class Context {
 public:
  explicit Context(int device = 0) : device_(device) {
    const int rc = noc_create(&h_, device);
    if (rc != NOC_OK)
      throw NocError(rc, "noc_create(device " + std::to_string(device) +
                             ") failed: no usable sm_100 CUDA device (libnoc_b200 has no CPU fallback)");
  }
  ~Context() { noc_destroy(h_); }
  Context(const Context&) = delete;
  Context& operator=(const Context&) = delete;
  noc_ctx* handle() const { return h_; }
  int device() const { return device_; }
  void check(int rc) const {
    if (rc != NOC_OK) throw NocError(rc, std::string("libnoc_b200: ") + noc_last_error(h_));
  }
  void setStream(void* cuda_stream) { check(noc_set_stream(h_, cuda_stream)); }
  void synchronize() { check(noc_synchronize(h_)); }
  int64_t launchCount() const { return noc_launch_count(h_); }

 private:
  noc_ctx* h_ = nullptr;
  int device_;
};

// Q (n*n), R (m*m), Qf (n*n) packed the way the C-ABI wants them
struct Weights {
  int n = 12, m = 4;
  std::vector<double> qrqf;
  Weights() = default;
  Weights(int n_, int m_, const double* Q, const double* R, const double* Qf) : n(n_), m(m_) {
    qrqf.insert(qrqf.end(), Q, Q + n * n);
    qrqf.insert(qrqf.end(), R, R + m * m);
    qrqf.insert(qrqf.end(), Qf, Qf + n * n);
  }
  // the benchmark weights: Q = diag(10,10,10,1,1,1,5,5,5,.1,.1,.1) per vehicle, R = 0.1 I, Qf = 10 Q
  static Weights quadrotorDefault(int vehicles = 1) {
    const double q[12] = {10, 10, 10, 1, 1, 1, 5, 5, 5, 0.1, 0.1, 0.1};
    const int n = 12 * vehicles, m = 4 * vehicles;
    std::vector<double> Q((size_t)n * n, 0.0), R((size_t)m * m, 0.0), Qf((size_t)n * n, 0.0);
    for (int i = 0; i < n; ++i) { Q[(size_t)i * n + i] = q[i % 12]; Qf[(size_t)i * n + i] = 10.0 * q[i % 12]; }
    for (int i = 0; i < m; ++i) R[(size_t)i * m + i] = 0.1;
    return Weights(n, m, Q.data(), R.data(), Qf.data());
  }
};

inline void checkWeightDims(const Weights& w, const noc_problem_t& p) {
  if (w.n != p.n || w.m != p.m || w.qrqf.size() != (size_t)(2 * w.n * w.n + w.m * w.m))
    throw NocError(NOC_ERR_BAD_ARG, "Weights are " + std::to_string(w.n) + "x" + std::to_string(w.m) + ", the problem is " +
                                        std::to_string(p.n) + "x" + std::to_string(p.m));
}

// Finite-horizon batched LQR.
class BatchedLqr {
 public:
  BatchedLqr(std::shared_ptr<Context> ctx, int horizon, int model = NOC_MODEL_QUAD12) : ctx_(std::move(ctx)) {
    noc_problem_defaults(&prob_, model, horizon);
    weights_ = Weights::quadrotorDefault(model == NOC_MODEL_DUAL24 ? 2 : 1);
  }
  noc_problem_t& problem() { return prob_; }
  // the weights' dimensions must be the problem's (the library reads Q | R | Qf with the problem's n, m)
  void setWeights(const Weights& w) { checkWeightDims(w, prob_); weights_ = w; }
  // time-parallel sweep for small batches (NOC_FLAG_TIME_PARALLEL; n=12, m=4): ~2x shorter calls, results agree with
  // the default sweep to ~1e-12 relative instead of bitwise
  void setTimeParallel(bool on) { prob_.flags = on ? (prob_.flags | NOC_FLAG_TIME_PARALLEL) : (prob_.flags & ~NOC_FLAG_TIME_PARALLEL); }
  void setTimeStep(double dt) { prob_.dt = dt; }
  int stateDim() const { return prob_.n; }
  int inputDim() const { return prob_.m; }
  const Weights& weights() const { return weights_; }

  // caller-supplied A_k,B_k records: AB[batch][N][n*n+n*m] in problem().layout
  void solve(int64_t batch, const double* AB, const double* x0dev, double* K, double* K0, double* P0, double* cost,
             int32_t* status) {
    noc_problem_t p = prob_;
    p.model = NOC_MODEL_LTV_USER;
    ctx_->check(noc_lqr_solve_batch(ctx_->handle(), &p, batch, AB, weights_.qrqf.data(), 0, x0dev, K, K0, P0, cost,
                                    status));
  }
  // the same with one Q | R | Qf block per instance: weights[batch][n*n + m*m + n*n] (host or device).  (12,4) and (24,8)
  // with A_THEN_B records run on the tensor-core sweeps; other shapes on the thread-per-instance kernel.
  void solveWithInstanceWeights(int64_t batch, const double* AB, const double* weights, const double* x0dev, double* K,
                                double* K0, double* P0, double* cost, int32_t* status) {
    noc_problem_t p = prob_;
    p.model = NOC_MODEL_LTV_USER;
    ctx_->check(noc_lqr_solve_batch(ctx_->handle(), &p, batch, AB, weights, 1, x0dev, K, K0, P0, cost, status));
  }
  // built-in model: x0 -> hover-thrust nominal -> linearise -> sweep
  void solveFromInitialStates(int64_t batch, const double* x0, double* K, double* K0, double* P0, double* cost,
                              int32_t* status) {
    ctx_->check(noc_lqr_solve_from_x0(ctx_->handle(), &prob_, batch, x0, weights_.qrqf.data(), K, K0, P0, cost,
                                      status));
  }
  // time-varying LQR along the open-loop rollout of the nominal inputs ubar [batch][N][m] (nullptr = hover thrust);
  // xbar [batch][N+1][n] receives that nominal if given
  void solveAlong(int64_t batch, const double* x0, const double* ubar, double* xbar, double* K, double* K0, double* P0,
                  double* cost, int32_t* status) {
    ctx_->check(noc_lqr_solve_along(ctx_->handle(), &prob_, batch, x0, ubar, weights_.qrqf.data(), xbar, K, K0, P0,
                                    cost, status));
  }
  void linearize(int64_t batch, const double* xbar, const double* ubar, double* AB) {
    ctx_->check(noc_linearize_batch(ctx_->handle(), &prob_, batch, xbar, ubar, AB));
  }
  void rolloutOpenLoop(int64_t batch, const double* x0, const double* ubar, double* xbar) {
    ctx_->check(noc_rollout_open_loop(ctx_->handle(), &prob_, batch, x0, ubar, xbar));
  }
  // "pick the best rollout" on this GPU
  std::pair<int64_t, double> argminCost(const double* cost, int64_t batch) {
    int64_t idx = -1;
    double val = 0.0;
    ctx_->check(noc_argmin_cost(ctx_->handle(), cost, batch, &idx, &val));
    return {idx, val};
  }

 private:
  std::shared_ptr<Context> ctx_;
  noc_problem_t prob_{};
  Weights weights_;
};

// Infinite-horizon LQR by fixed-point iteration (row a5): one LTI record [A|B] per instance.
class InfiniteHorizonLqr {
 public:
  explicit InfiniteHorizonLqr(std::shared_ptr<Context> ctx, double tol = 1e-12, int max_iter = 10000)
      : ctx_(std::move(ctx)) {
    noc_problem_defaults(&prob_, NOC_MODEL_LTV_USER, 1);
    prob_.tol = tol;
    prob_.max_iter = max_iter;
    weights_ = Weights::quadrotorDefault();
  }
  noc_problem_t& problem() { return prob_; }
  // caller-supplied LTI records: the weights fix the dimensions of the problem
  void setWeights(const Weights& w) {
    if (w.qrqf.size() != (size_t)(2 * w.n * w.n + w.m * w.m)) throw NocError(NOC_ERR_BAD_ARG, "Weights: size does not match n, m");
    weights_ = w; prob_.n = w.n; prob_.m = w.m;
  }
  void solve(int64_t batch, const double* AB, double* P, double* K, int32_t* iters, int32_t* status) {
    // the C-ABI takes Q | R only; they are the leading part of the packed weights
    ctx_->check(noc_dare_solve_batch(ctx_->handle(), &prob_, batch, AB, weights_.qrqf.data(), P, K, iters, status));
  }
  // the built-in quadrotor linearised at the operating points xbar [batch][12], ubar [batch][4] (nullptr = hover thrust)
  void solveAt(int64_t batch, const double* xbar, const double* ubar, double* P, double* K, int32_t* iters,
               int32_t* status) {
    if (weights_.n != 12 || weights_.m != 4) throw NocError(NOC_ERR_BAD_ARG, "solveAt: built-in quadrotor needs 12x4 weights");
    noc_problem_t p = prob_;
    p.model = NOC_MODEL_QUAD12; p.n = 12; p.m = 4;
    ctx_->check(noc_dare_solve_at(ctx_->handle(), &p, batch, xbar, ubar, weights_.qrqf.data(), P, K, iters, status));
  }

 private:
  std::shared_ptr<Context> ctx_;
  noc_problem_t prob_{};
  Weights weights_;
};

// SLQ / iLQR with the built-in nonlinear models (rows a1-a4).
class BatchedSlq {
 public:
  BatchedSlq(std::shared_ptr<Context> ctx, int horizon, int model = NOC_MODEL_QUAD12) : ctx_(std::move(ctx)) {
    noc_problem_defaults(&prob_, model, horizon);
    weights_ = Weights::quadrotorDefault(model == NOC_MODEL_DUAL24 ? 2 : 1);
  }
  noc_problem_t& problem() { return prob_; }
  void setWeights(const Weights& w) { checkWeightDims(w, prob_); weights_ = w; }
  // box on every input component (rotor thrust limits): control-limited backward pass + clamped rollouts
  void setInputBox(double u_min, double u_max) { prob_.u_min = u_min; prob_.u_max = u_max; }
  // Levenberg-Marquardt schedule: NOC_REG_X10 (default) or NOC_REG_ADAPTIVE (mu raised on failed factorisations and
  // failed line searches, lowered after accepted steps)
  void setRegularisation(int schedule, double mu0 = 0.0) { prob_.reg_schedule = schedule; prob_.reg0 = mu0; }
  void solve(int64_t batch, const double* x0, const double* xgoal, double* x, double* u, double* K, double* cost,
             int32_t* iters, int32_t* alpha_idx, int32_t* status) {
    ctx_->check(noc_slq_solve_batch(ctx_->handle(), &prob_, batch, x0, xgoal, weights_.qrqf.data(), x, u, K, cost,
                                    iters, alpha_idx, status));
  }
  // warm start from the input sequences u_init [batch][N][m] (e.g. the previous MPC solution shifted by one step)
  void solveWarm(int64_t batch, const double* x0, const double* xgoal, const double* u_init, double* x, double* u,
                 double* K, double* cost, int32_t* iters, int32_t* alpha_idx, int32_t* status) {
    ctx_->check(noc_slq_solve_warm(ctx_->handle(), &prob_, batch, x0, xgoal, weights_.qrqf.data(), u_init, x, u, K, cost,
                                   iters, alpha_idx, status));
  }

 private:
  std::shared_ptr<Context> ctx_;
  noc_problem_t prob_{};
  Weights weights_;
};

// contiguous slice [begin, end) of `batch` owned by shard g of G
inline std::pair<int64_t, int64_t> shardSlice(int64_t batch, int g, int G) {
  const int64_t base = batch / G, rem = batch % G;
  const int64_t begin = g * base + std::min<int64_t>(g, rem);
  return {begin, begin + base + (g < rem ? 1 : 0)};
}

// RAII device buffer on a Context's GPU (noc_device_alloc): lets host code keep data resident without the CUDA headers
class DeviceBuffer {
 public:
  DeviceBuffer() = default;
  DeviceBuffer(std::shared_ptr<Context> ctx, size_t bytes) : ctx_(std::move(ctx)) { reserve(bytes); }
  ~DeviceBuffer() { release(); }
  DeviceBuffer(const DeviceBuffer&) = delete;
  DeviceBuffer& operator=(const DeviceBuffer&) = delete;
  void bind(std::shared_ptr<Context> ctx) { release(); ctx_ = std::move(ctx); }
  void reserve(size_t bytes) {   // grows, never shrinks; contents are not preserved
    if (bytes <= cap_) return;
    release();
    ctx_->check(noc_device_alloc(ctx_->handle(), bytes, &p_));
    cap_ = bytes;
  }
  template <class T> T* as() const { return static_cast<T*>(p_); }
  size_t capacity() const { return cap_; }

 private:
  void release() {
    if (p_) noc_device_free(ctx_->handle(), p_);
    p_ = nullptr; cap_ = 0;
  }
  std::shared_ptr<Context> ctx_;
  void* p_ = nullptr;
  size_t cap_ = 0;
};

// One host process driving several GPUs of a box: the batch is cut into contiguous, EQUAL slices (the
// batch is padded up to a multiple of the shard count with copies of the last instance), one host thread and one
// Context per device, no data-path collective.  With distinct devices the shards form an NCCL communicator inside
// the library (noc_comm_create_all): every GPU's slice of costs and first-step gains is gathered over NVLink into
// device-resident arrays on every GPU while the next chunk is swept, and the best rollout is picked across GPUs
// on the device (local multi-CTA argmin -> all-gather of 16-byte pairs -> final argmin).  Host outputs, if given,
// are copied from each shard's own slice over its own PCIe link.  When the device list repeats a GPU (a slicing
// test on one GPU) NCCL cannot be used and the shards only share the host arrays.
class ShardedBatchedLqr {
 public:
  ShardedBatchedLqr(const std::vector<int>& devices, int horizon, int model = NOC_MODEL_QUAD12) {
    for (int d : devices) {
      ctxs_.push_back(std::make_shared<Context>(d));
      shards_.emplace_back(ctxs_.back(), horizon, model);
    }
    bool distinct = true;
    for (size_t i = 0; i < devices.size(); ++i)
      for (size_t j = 0; j < i; ++j) distinct = distinct && devices[i] != devices[j];
    if (distinct) {
      std::vector<noc_ctx*> h;
      for (auto& c : ctxs_) h.push_back(c->handle());
      comms_.assign(devices.size(), nullptr);
      const int rc = noc_comm_create_all(comms_.data(), h.data(), (int)h.size());
      if (rc != NOC_OK) throw NocError(rc, std::string("noc_comm_create_all: ") + noc_last_error(h[0]));
      for (size_t g = 0; g < devices.size(); ++g) {
        bufs_.push_back(std::make_unique<Bufs>());
        for (auto& b : bufs_[g]->b) b.bind(ctxs_[g]);
      }
    }
  }
  ~ShardedBatchedLqr() {
    for (auto* c : comms_) noc_comm_destroy(c);
    bufs_.clear();
  }
  size_t shards() const { return shards_.size(); }
  bool deviceGather() const { return !comms_.empty(); }
  void setWeights(const Weights& w) { for (auto& s : shards_) s.setWeights(w); }
  int stateDim() const { return shards_.front().stateDim(); }
  int inputDim() const { return shards_.front().inputDim(); }
  // slice length of the last call (padded batch / shards) and the device-resident gathered results on shard g:
  // costs [shards][slice], first-step gains [shards][slice][m*n]
  int64_t sliceLength() const { return slice_; }
  const double* deviceCosts(int g) const { return bufs_[g]->b[2].as<double>(); }
  const double* deviceK0(int g) const { return bufs_[g]->b[3].as<double>(); }

  // x0 [batch][n] on the host; K0 [batch][m*n], P0 [batch][n*n] (host-gather mode only), cost [batch], status [batch]
  // may be nullptr.  Returns (index, cost) of the best instance over all shards.
  std::pair<int64_t, double> solveFromInitialStates(int64_t batch, const double* x0, double* K0, double* P0,
                                                    double* cost, int32_t* status) {
    const int n = stateDim(), m = inputDim();
    const int G = (int)shards_.size();
    std::vector<std::thread> th;
    std::vector<std::string> err(G);
    if (!deviceGather()) {
      for (int g = 0; g < G; ++g)
        th.emplace_back([&, g] {
          const auto sl = shardSlice(batch, g, G);
          const int64_t cnt = sl.second - sl.first, o = sl.first;
          if (cnt == 0) return;
          try {
            shards_[g].solveFromInitialStates(cnt, x0 + o * n, nullptr, K0 ? K0 + o * m * n : nullptr,
                                              P0 ? P0 + o * n * n : nullptr, cost ? cost + o : nullptr,
                                              status ? status + o : nullptr);
          } catch (const std::exception& e) { err[g] = e.what(); }
        });
      for (auto& t : th) t.join();
      for (auto& e : err)
        if (!e.empty()) throw NocError(NOC_ERR_CUDA, e);
      int64_t best = -1;
      double bv = 0.0;
      if (cost)
        for (int64_t i = 0; i < batch; ++i)
          if (cost[i] == cost[i] && (best < 0 || cost[i] < bv)) { best = i; bv = cost[i]; }
      return {best, bv};
    }
    if (P0) throw NocError(NOC_ERR_UNSUPPORTED, "ShardedBatchedLqr: P0 is not gathered in device-gather mode");
    const int64_t cnt = (batch + G - 1) / G;   // equal slices; the tail of the last one repeats the last instance
    slice_ = cnt;
    std::vector<noc_cost_index_t> best(G);
    for (int g = 0; g < G; ++g)
      th.emplace_back([&, g] {
        try {
          noc_ctx* h = ctxs_[g]->handle();
          auto& B = bufs_[g]->b;
          const Weights& w = shards_[g].weights();
          B[0].reserve(sizeof(double) * cnt * n);                 // x0 slice
          B[1].reserve(sizeof(double) * w.qrqf.size());           // weights
          B[2].reserve(sizeof(double) * cnt * G);                 // gathered costs
          B[3].reserve(sizeof(double) * cnt * G * m * n);         // gathered first-step gains
          B[4].reserve(sizeof(int32_t) * cnt);                    // local status
          B[5].reserve(sizeof(noc_cost_index_t));                 // the winner
          const int64_t o = g * cnt, real = std::max<int64_t>(0, std::min<int64_t>(cnt, batch - o));
          if (real > 0) ctxs_[g]->check(noc_memcpy(h, B[0].as<double>(), x0 + o * n, sizeof(double) * real * n));
          for (int64_t i = real; i < cnt; ++i)                    // padding: copies of the batch's last instance
            ctxs_[g]->check(noc_memcpy(h, B[0].as<double>() + i * n, x0 + (batch - 1) * n, sizeof(double) * n));
          ctxs_[g]->check(noc_memcpy(h, B[1].as<double>(), w.qrqf.data(), sizeof(double) * w.qrqf.size()));
          ctxs_[g]->check(noc_comm_lqr_solve_from_x0(comms_[g], &shards_[g].problem(), cnt, B[0].as<double>(),
                                                     B[1].as<double>(), B[2].as<double>(), B[3].as<double>(),
                                                     B[4].as<int32_t>(), B[5].as<noc_cost_index_t>(), &best[g]));
          if (real > 0) {
            if (cost) ctxs_[g]->check(noc_memcpy(h, cost + o, B[2].as<double>() + g * cnt, sizeof(double) * real));
            if (K0) ctxs_[g]->check(noc_memcpy(h, K0 + o * m * n, B[3].as<double>() + g * cnt * m * n, sizeof(double) * real * m * n));
            if (status) ctxs_[g]->check(noc_memcpy(h, status + o, B[4].as<int32_t>(), sizeof(int32_t) * real));
          }
          ctxs_[g]->synchronize();
        } catch (const std::exception& e) { err[g] = e.what(); }
      });
    for (auto& t : th) t.join();
    for (auto& e : err)
      if (!e.empty()) throw NocError(NOC_ERR_CUDA, e);
    // every rank holds the same winner; a padded copy of the last instance maps back onto it
    int64_t bi = best[0].index;
    if (bi >= batch) bi = batch - 1;
    return {bi, best[0].cost};
  }

 private:
  struct Bufs { DeviceBuffer b[6]; };
  std::vector<std::shared_ptr<Context>> ctxs_;
  std::vector<BatchedLqr> shards_;
  std::vector<noc_comm*> comms_;
  std::vector<std::unique_ptr<Bufs>> bufs_;
  int64_t slice_ = 0;
};

}  // namespace lqr_control
