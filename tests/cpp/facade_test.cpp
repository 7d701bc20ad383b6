// C++ host-side test of the lqr_control facade (include/lqr_control/batched_lqr.hpp) against the CPU oracle.
// Built and run by tests/test_cpp_facade.py.  Exit codes: 0 = all checks passed, 3 = no usable GPU (the facade
// must fail loudly: there is no CPU fallback), 1 = mismatch.
#include <cmath>
#include <cstdio>
#include <cstring>
#include <vector>

#include "lqr_control/batched_lqr.hpp"
extern "C" {
#include "../../oracle/noc_oracle.h"
}

using namespace lqr_control;

static unsigned long long rng_state = 0x4E4F43ULL;
static double urand(double lo, double hi) {
  rng_state = rng_state * 6364136223846793005ULL + 1442695040888963407ULL;
  return lo + (hi - lo) * ((rng_state >> 11) * (1.0 / 9007199254740992.0));
}
static std::vector<double> sample_x0(int64_t B) {
  std::vector<double> x((size_t)B * 12);
  for (int64_t b = 0; b < B; ++b) {
    double* p = &x[(size_t)b * 12];
    for (int i = 0; i < 3; ++i) p[i] = urand(-2, 2);
    for (int i = 3; i < 6; ++i) p[i] = urand(-1, 1);
    p[6] = urand(-0.3, 0.3); p[7] = urand(-0.3, 0.3); p[8] = urand(-3.14, 3.14);
    for (int i = 9; i < 12; ++i) p[i] = urand(-0.2, 0.2);
  }
  return x;
}
static bool same(const std::vector<double>& a, const std::vector<double>& b) {
  return a.size() == b.size() && std::memcmp(a.data(), b.data(), a.size() * sizeof(double)) == 0;
}
#define CHECK(c)                                                   \
  do {                                                             \
    if (!(c)) { std::printf("CHECK FAILED line %d: %s\n", __LINE__, #c); return 1; } \
  } while (0)

int main() {
  std::shared_ptr<Context> ctx;
  try {
    ctx = std::make_shared<Context>(0);
  } catch (const NocError& e) {
    std::printf("NO DEVICE: %s (code %d)\n", e.what(), e.code());
    return 3;
  }
  const Weights w = Weights::quadrotorDefault();

  {  // finite-horizon LQR from initial states, single context and two shards
    const int64_t B = 45;
    const int N = 30;
    auto x0 = sample_x0(B);
    std::vector<double> K0(B * 48), P0(B * 144), cost(B), rK0(B * 48), rP0(B * 144), rcost(B);
    std::vector<int32_t> st(B, -1), rst(B, -1);
    BatchedLqr lqr(ctx, N);
    lqr.solveFromInitialStates(B, x0.data(), nullptr, K0.data(), P0.data(), cost.data(), st.data());
    noc_oracle_lqr_from_x0_batch(NOC_O_MODEL_QUAD12, N, 0.02, B, x0.data(), w.qrqf.data(), rK0.data(), rP0.data(),
                                 rcost.data(), rst.data(), 0);
    CHECK(same(K0, rK0) && same(P0, rP0) && same(cost, rcost) && st == rst);
    auto best = lqr.argminCost(cost.data(), B);
    int64_t rb = 0;
    for (int64_t i = 1; i < B; ++i) if (rcost[i] < rcost[rb]) rb = i;
    CHECK(best.first == rb && best.second == rcost[rb]);

    {  // the time-parallel sweep (NOC_FLAG_TIME_PARALLEL): same gains to north_star's 1e-9, not bitwise
      BatchedLqr tp(ctx, N);
      tp.setTimeParallel(true);
      std::vector<double> tK0(B * 48), tP0(B * 144), tcost(B);
      std::vector<int32_t> tst(B, -1);
      tp.solveFromInitialStates(B, x0.data(), nullptr, tK0.data(), tP0.data(), tcost.data(), tst.data());
      double worst = 0.0, scale = 0.0;
      for (size_t i = 0; i < tK0.size(); ++i) { worst = std::max(worst, std::fabs(tK0[i] - rK0[i])); scale = std::max(scale, std::fabs(rK0[i])); }
      CHECK(worst <= 1e-9 * scale && tst == rst);
      for (int64_t i = 0; i < B; ++i) CHECK(std::fabs(tcost[i] - rcost[i]) <= 1e-9 * std::fabs(rcost[i]));
    }

    ShardedBatchedLqr sh({0, 0, 0}, N);   // three shards (contexts) on one GPU: exercises the slicing
    std::vector<double> sK0(B * 48), sP0(B * 144), scost(B);
    std::vector<int32_t> sst(B, -1);
    auto sbest = sh.solveFromInitialStates(B, x0.data(), sK0.data(), sP0.data(), scost.data(), sst.data());
    CHECK(same(sK0, rK0) && same(sP0, rP0) && same(scost, rcost) && sst == rst);
    CHECK(sbest.first == rb && sbest.second == rcost[rb]);
    {  // the in-library multi-GPU plane on ONE device (an NCCL communicator of size 1): device-resident gather + pick
      ShardedBatchedLqr one({0}, N);
      CHECK(one.deviceGather());
      std::vector<double> tK0(B * 48), tcost(B), dcost(B), dK0(B * 48);
      std::vector<int32_t> tst(B, -1);
      auto tbest = one.solveFromInitialStates(B, x0.data(), tK0.data(), nullptr, tcost.data(), tst.data());
      CHECK(same(tK0, rK0) && same(tcost, rcost) && tst == rst && tbest.first == rb && tbest.second == rcost[rb]);
      // the gathered arrays stay on the device: read them back through the C-ABI copy
      CHECK(noc_memcpy(ctx->handle(), dcost.data(), one.deviceCosts(0), sizeof(double) * B) == NOC_OK);
      CHECK(noc_memcpy(ctx->handle(), dK0.data(), one.deviceK0(0), sizeof(double) * B * 48) == NOC_OK);
      ctx->synchronize();
      CHECK(same(dcost, rcost) && same(dK0, rK0));
      std::printf("comm plane (1 rank) OK\n");
    }
    {  // a second GPU, when the box has one: two NCCL ranks in this process, one host thread each; costs and first-step
       // gains of BOTH slices end up on BOTH devices without passing through host arrays
      bool two = false;
      try { Context probe(1); two = true; } catch (const NocError&) {}
      if (two) {
        ShardedBatchedLqr sh2({0, 1}, N);
        CHECK(sh2.deviceGather());
        for (int64_t Bt : {B, B - 1}) {   // B - 1: the second slice is padded with a copy of the last instance
          std::vector<double> tK0(Bt * 48), tcost(Bt);
          std::vector<int32_t> tst(Bt, -1);
          auto tbest = sh2.solveFromInitialStates(Bt, x0.data(), tK0.data(), nullptr, tcost.data(), tst.data());
          int64_t rbt = 0;
          for (int64_t i = 1; i < Bt; ++i) if (rcost[i] < rcost[rbt]) rbt = i;
          CHECK(std::memcmp(tK0.data(), rK0.data(), sizeof(double) * Bt * 48) == 0);
          CHECK(std::memcmp(tcost.data(), rcost.data(), sizeof(double) * Bt) == 0);
          CHECK(std::memcmp(tst.data(), rst.data(), sizeof(int32_t) * Bt) == 0);
          CHECK(tbest.first == rbt && tbest.second == rcost[rbt]);
          // device 1 holds device 0's slice as well (and vice versa): read the gathered arrays back from GPU 1
          const int64_t cnt = sh2.sliceLength();
          std::vector<double> g1(2 * cnt), k1(2 * cnt * 48);
          Context c1(1);
          CHECK(noc_memcpy(c1.handle(), g1.data(), sh2.deviceCosts(1), sizeof(double) * 2 * cnt) == NOC_OK);
          CHECK(noc_memcpy(c1.handle(), k1.data(), sh2.deviceK0(1), sizeof(double) * 2 * cnt * 48) == NOC_OK);
          c1.synchronize();
          CHECK(std::memcmp(g1.data(), rcost.data(), sizeof(double) * Bt) == 0);
          CHECK(std::memcmp(k1.data(), rK0.data(), sizeof(double) * Bt * 48) == 0);
        }
        std::printf("two-GPU comm plane OK\n");
      }
    }
    auto s0 = shardSlice(10, 0, 3), s1 = shardSlice(10, 1, 3), s2 = shardSlice(10, 2, 3);
    CHECK(s0.first == 0 && s0.second == 4 && s1.first == 4 && s1.second == 7 && s2.first == 7 && s2.second == 10);
  }
  {  // infinite-horizon LQR at hover
    double x[12] = {0}, u[4], A[144], Bm[48], P[144], K[48], rP[144], rK[48];
    noc_oracle_hover_input(NOC_O_MODEL_QUAD12, u);
    noc_oracle_linearize(NOC_O_MODEL_QUAD12, x, u, 0.02, A, Bm);
    std::vector<double> rec(192);
    std::memcpy(rec.data(), A, sizeof(A));
    std::memcpy(rec.data() + 144, Bm, sizeof(Bm));
    int32_t it = 0, st = -1;
    InfiniteHorizonLqr dare(ctx);
    dare.solve(1, rec.data(), P, K, &it, &st);
    const int rit = noc_oracle_dare(12, 4, A, Bm, w.qrqf.data(), w.qrqf.data() + 144, 1e-12, 10000, rP, rK);
    CHECK(st == 0 && it == rit && std::memcmp(P, rP, sizeof(P)) == 0 && std::memcmp(K, rK, sizeof(K)) == 0);
    std::printf("hover DARE: %d iterations\n", it);
    // the same through solveAt: the operating point instead of its linearisation
    double P2[144], K2[48];
    int32_t it2 = 0, st2 = -1;
    dare.solveAt(1, x, nullptr, P2, K2, &it2, &st2);
    CHECK(st2 == 0 && it2 == rit && std::memcmp(P2, rP, sizeof(P2)) == 0 && std::memcmp(K2, rK, sizeof(K2)) == 0);
  }
  {  // SLQ to waypoints
    const int64_t B = 6;
    const int N = 40;
    std::vector<double> x0(B * 12, 0.0), xg(B * 12, 0.0);
    for (int64_t b = 0; b < B; ++b) { xg[b * 12 + 0] = urand(-2, 2); xg[b * 12 + 1] = urand(-2, 2); xg[b * 12 + 2] = urand(-2, 2); }
    std::vector<double> x(B * (N + 1) * 12), u(B * N * 4), cost(B), rx(x.size()), ru(u.size()), rcost(B);
    std::vector<int32_t> iters(B), st(B), aidx(B * 50), riters(B), rst(B), raidx(B * 50);
    BatchedSlq slq(ctx, N);
    slq.solve(B, x0.data(), xg.data(), x.data(), u.data(), nullptr, cost.data(), iters.data(), aidx.data(), st.data());
    noc_oracle_slq_opts o{NOC_O_MODEL_QUAD12, N, 50, 11, 0.02, 1e-8, 0.0, 1e-4, -INFINITY, INFINITY};
    noc_oracle_slq_batch(&o, B, x0.data(), xg.data(), w.qrqf.data(), rx.data(), ru.data(), nullptr, rcost.data(),
                         riters.data(), raidx.data(), rst.data(), 0);
    CHECK(same(x, rx) && same(u, ru) && same(cost, rcost) && iters == riters && st == rst && aidx == raidx);
    // the same problems with rotor-thrust limits (control-limited SLQ): setInputBox
    slq.setInputBox(0.0, 4.0);
    slq.solve(B, x0.data(), xg.data(), x.data(), u.data(), nullptr, cost.data(), iters.data(), aidx.data(), st.data());
    noc_oracle_slq_opts ob{NOC_O_MODEL_QUAD12, N, 50, 11, 0.02, 1e-8, 0.0, 1e-4, 0.0, 4.0};
    noc_oracle_slq_batch(&ob, B, x0.data(), xg.data(), w.qrqf.data(), rx.data(), ru.data(), nullptr, rcost.data(),
                         riters.data(), raidx.data(), rst.data(), 0);
    CHECK(same(x, rx) && same(u, ru) && same(cost, rcost) && iters == riters && st == rst && aidx == raidx);
    bool in_box = true;
    for (double v : u) in_box = in_box && v >= 0.0 && v <= 4.0;
    CHECK(in_box);
  }
  for (int model = 0; model < 2; ++model) {  // one Q | R | Qf block per instance: the PERW tensor-core sweeps (n=12 and n=24)
    const int64_t B = 37;
    const int N = 11, n = model ? 24 : 12, m = model ? 8 : 4, rec = n * n + n * m, qlen = 2 * n * n + m * m;
    const Weights wm = Weights::quadrotorDefault(model ? 2 : 1);
    std::vector<double> x0((size_t)B * n);
    for (auto& v : x0) v = urand(-0.3, 0.3);
    std::vector<double> AB((size_t)B * N * rec), qr((size_t)B * qlen);
    noc_oracle_make_ab_batch(model ? NOC_O_MODEL_DUAL24 : NOC_O_MODEL_QUAD12, N, 0.02, B, x0.data(), 0, AB.data(), 0);
    for (int64_t b = 0; b < B; ++b)
      for (int e = 0; e < qlen; ++e) qr[(size_t)b * qlen + e] = wm.qrqf[e] * (0.5 + 0.04 * (double)b);
    std::vector<double> K((size_t)B * N * m * n), K0((size_t)B * m * n), P0((size_t)B * n * n), cost(B);
    std::vector<double> rK(K.size()), rK0(K0.size()), rP0(P0.size()), rcost(B);
    std::vector<int32_t> st(B, -1), rst(B, -1);
    BatchedLqr lqr(ctx, N, model ? NOC_MODEL_DUAL24 : NOC_MODEL_QUAD12);
    lqr.solveWithInstanceWeights(B, AB.data(), qr.data(), x0.data(), K.data(), K0.data(), P0.data(), cost.data(), st.data());
    noc_oracle_lqr_batch(n, m, N, B, AB.data(), 0, qr.data(), 1, x0.data(), rK.data(), rK0.data(), rP0.data(), rcost.data(),
                         rst.data(), 0);
    if (!(same(K, rK) && same(K0, rK0) && same(P0, rP0) && same(cost, rcost) && st == rst)) {
      std::printf("per-instance weights, model %d: K %d K0 %d P0 %d cost %d status %d\n", model, (int)same(K, rK), (int)same(K0, rK0),
                  (int)same(P0, rP0), (int)same(cost, rcost), (int)(st == rst));
      for (size_t i = 0; i < K.size(); ++i)
        if (std::memcmp(&K[i], &rK[i], 8)) { std::printf("  first K mismatch at %zu (instance %zu, step %zu): %.17g vs %.17g\n", i, i / ((size_t)N * m * n), (i / ((size_t)m * n)) % N, K[i], rK[i]); break; }
      for (int64_t b = 0; b < B; ++b) if (st[b] != rst[b]) { std::printf("  status[%lld] = %d vs %d\n", (long long)b, st[b], rst[b]); break; }
    }
    CHECK(same(K, rK) && same(K0, rK0) && same(P0, rP0) && same(cost, rcost) && st == rst);
  }
  {  // errors surface as exceptions with the library's message
    BatchedLqr lqr(ctx, 10);
    bool threw = false;
    try { lqr.solveFromInitialStates(4, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr); }
    catch (const NocError& e) { threw = e.code() == NOC_ERR_BAD_ARG; }
    CHECK(threw);
  }
  std::printf("FACADE OK\n");
  return 0;
}
