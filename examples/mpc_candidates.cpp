// examples/mpc_candidates.cpp — the use the lqr_control node slot is scaffolded for (lqr_control/CMakeLists.txt:136):
// at every control tick, solve a finite-horizon LQR around the hover-thrust nominal for a cloud of candidate
// initial states (measurement uncertainty / warm starts), pick the best rollout and apply its first-step gain.
// Build (after `make -C noc_b200/csrc`):
//   g++ -std=c++17 -O2 -I include examples/mpc_candidates.cpp -L noc_b200 -lnoc_b200
//     -Wl,-rpath,$PWD/noc_b200 -Wl,-rpath,/usr/local/cuda/lib64 -pthread -o examples/mpc_candidates
#include <chrono>
#include <cstdio>
#include <random>
#include <vector>

#include "lqr_control/batched_lqr.hpp"

int main(int argc, char** argv) {
  using namespace lqr_control;
  const int64_t B = argc > 1 ? std::atoll(argv[1]) : 4096;   // candidates per tick
  const int N = 50, ticks = 20;
  std::shared_ptr<Context> ctx;
  try {
    ctx = std::make_shared<Context>(0);
  } catch (const NocError& e) {
    std::fprintf(stderr, "%s\n", e.what());   // no GPU: there is deliberately no CPU fallback
    return 3;
  }
  BatchedLqr lqr(ctx, N);
  std::mt19937_64 rng(0x4E4F43);
  std::uniform_real_distribution<double> pos(-0.2, 0.2), att(-0.05, 0.05);
  std::vector<double> state(12, 0.0), x0(B * 12), K0(B * 48), cost(B);
  std::vector<int32_t> status(B);
  state[0] = 1.0; state[2] = -0.5; state[8] = 0.3;              // true state: 1 m off in x, yawed
  double total_ms = 0.0;
  for (int tick = 0; tick < ticks; ++tick) {
    for (int64_t b = 0; b < B; ++b)
      for (int i = 0; i < 12; ++i) x0[b * 12 + i] = state[i] + (i < 6 ? pos(rng) : att(rng)) * (b ? 1.0 : 0.0);
    const auto t0 = std::chrono::steady_clock::now();
    lqr.solveFromInitialStates(B, x0.data(), nullptr, K0.data(), nullptr, cost.data(), status.data());
    const auto best = lqr.argminCost(cost.data(), B);
    total_ms += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    // u = u_hover + K0 (x - 0): here we only propagate a crude "move towards the goal" to change the state
    for (int i = 0; i < 6; ++i) state[i] *= 0.9;
    if (tick % 5 == 0)
      std::printf("tick %2d: best candidate %lld cost %.4f (status %d)\n", tick, (long long)best.first, best.second,
                  status[best.first]);
  }
  std::printf("%lld candidates x N=%d per tick: %.3f ms per tick (host buffers, incl. copies)\n", (long long)B, N,
              total_ms / ticks);
  return 0;
}
