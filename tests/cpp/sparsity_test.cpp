// Host-side check of the structural-sparsity predicates the fused sweeps prune with (noc_b200/csrc/model.cuh):
// quad_jac_nz_A/B and dual_jac_nz_A/B must be true for EVERY entry the Jacobian emitters write (a missed entry
// would silently drop a product from the Riccati chain) and for nothing else (a spurious one only costs time).
// model.cuh is plain C++ on the host, so this needs no GPU and no CUDA toolchain: g++ -std=c++17.
#include <cstdio>
#include <cstdlib>
#include <set>
#include <utility>

#include "../../noc_b200/csrc/model.cuh"

using namespace noc::model;

int main() {
  int bad = 0;
  srand(7);
  auto rnd = [] { return 2.0 * rand() / RAND_MAX - 1.0; };
  std::set<std::pair<int, int>> seenA, seenB, seenA24, seenB24;
  for (int trial = 0; trial < 200; ++trial) {
    double x[24], u[8];
    for (double& v : x) v = rnd();
    for (double& v : u) v = 2.0 + rnd();
    quad_jac_discrete(
        x, u, 0.02,
        [&](int i, int j, double) { seenA.insert({i, j}); if (!quad_jac_nz_A(i, j)) { ++bad; printf("A(%d,%d) emitted, predicate false\n", i, j); } },
        [&](int i, int j, double) { seenB.insert({i, j}); if (!quad_jac_nz_B(i, j)) { ++bad; printf("B(%d,%d) emitted, predicate false\n", i, j); } });
    model_jac_discrete<1>(
        x, u, 0.02,
        [&](int i, int j, double) { seenA24.insert({i, j}); if (!dual_jac_nz_A(i, j)) { ++bad; printf("A24(%d,%d) emitted, predicate false\n", i, j); } },
        [&](int i, int j, double) { seenB24.insert({i, j}); if (!dual_jac_nz_B(i, j)) { ++bad; printf("B24(%d,%d) emitted, predicate false\n", i, j); } });
  }
  int nA = 0, nB = 0, nA24 = 0, nB24 = 0;
  for (int i = 0; i < 12; ++i) {
    for (int j = 0; j < 12; ++j) nA += quad_jac_nz_A(i, j);
    for (int j = 0; j < 4; ++j) nB += quad_jac_nz_B(i, j);
  }
  for (int i = 0; i < 24; ++i) {
    for (int j = 0; j < 24; ++j) nA24 += dual_jac_nz_A(i, j);
    for (int j = 0; j < 8; ++j) nB24 += dual_jac_nz_B(i, j);
  }
  printf("quad: A %d predicate / %zu emitted, B %d / %zu; dual: A %d / %zu, B %d / %zu\n", nA, seenA.size(), nB, seenB.size(),
         nA24, seenA24.size(), nB24, seenB24.size());
  if (nA != (int)seenA.size() || nB != (int)seenB.size() || nA24 != (int)seenA24.size() || nB24 != (int)seenB24.size()) ++bad;
  if (nA != 40 || nB != 20 || nA24 != 98 || nB24 != 40) ++bad;
  // the per-group masks phase 1 of k_ric12 uses must cover the per-entry predicate
  for (int i = 0; i < 12; ++i)
    for (int c = 0; c < 16; ++c) {
      const bool nz = c < 12 ? quad_jac_nz_A(i, c) : quad_jac_nz_B(i, c - 12);
      if (nz && !quad_jac_nz_group(i, (c / 4) * 4)) { ++bad; printf("group mask misses (%d,%d)\n", i, c); }
    }
  // ... and the two slot masks of k_ric24 (slot A: a column of vehicle 0, slot B: one of vehicle 1)
  for (int l = 0; l < 24; ++l) {
    const bool slotA = l < 12 || (l >= 15 && l < 18), slotB = l >= 12 || (l >= 3 && l < 6);
    for (int c = 0; c < 32; ++c) {
      const bool nz = c < 24 ? dual_jac_nz_A(l, c) : dual_jac_nz_B(l, c - 24);
      const bool veh0 = c < 24 ? c < 12 : c - 24 < 4;
      if (nz && !(veh0 ? slotA : slotB)) { ++bad; printf("slot mask misses (%d,%d)\n", l, c); }
    }
  }
  printf(bad ? "SPARSITY MISMATCH\n" : "SPARSITY OK\n");
  return bad ? 1 : 0;
}
