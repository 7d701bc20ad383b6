"""Synthetic problem generators (noc_b200/problems.py): the SURVEY §8d recipe is std::mt19937_64 + libstdc++'s
uniform_real_distribution; the numpy restatement is checked against a g++ program that uses <random> itself."""
import os
import subprocess
import tempfile

import numpy as np

from noc_b200 import problems as pb

CPP = r'''
#include <cstdio>
#include <random>
int main() {
  const double lo[12] = {-2,-2,-2,-1,-1,-1,-0.3,-0.3,-3.141592653589793,-0.5,-0.5,-0.5};
  for (unsigned long long b : {0ULL, 1ULL, 65535ULL, 1048575ULL}) {
    std::mt19937_64 g(0x4E4F43ULL + b);
    for (int i = 0; i < 12; ++i) { std::uniform_real_distribution<double> d(lo[i], -lo[i]); std::printf("%a ", d(g)); }
    std::printf("\n");
  }
  std::mt19937_64 g;   // default seed 5489
  std::printf("%llu\n", (unsigned long long)g());
}
'''


def test_survey_sampler_matches_libstdcxx():
    with tempfile.TemporaryDirectory() as d:
        src, exe = os.path.join(d, "mt.cpp"), os.path.join(d, "mt")
        open(src, "w").write(CPP)
        subprocess.check_call(["g++", "-O1", "-std=c++17", src, "-o", exe])
        out = subprocess.run([exe], capture_output=True, text=True, check=True).stdout.strip().split("\n")
    want = np.array([[float.fromhex(t) for t in line.split()] for line in out[:4]])
    got = np.vstack([pb.sample_x0_survey(1, first_instance=b) for b in (0, 1, 65535, 1048575)])
    assert np.array_equal(got, want)
    assert int(pb.mt19937_64_first_draws([5489], 1)[0, 0]) == int(out[4]) == 14514284786278117030
    # chunking and offsets do not change the stream of an instance
    big = pb.sample_x0_survey(20000)
    assert np.array_equal(big[[0, 1]], want[:2]) and np.array_equal(pb.sample_x0_survey(3, first_instance=19997), big[19997:])


def test_survey_sampler_ranges():
    x = pb.sample_x0_survey(4096)
    lim = np.array([2.0] * 3 + [1.0] * 3 + [0.3, 0.3, np.pi] + [0.5] * 3)
    assert np.all(np.abs(x) <= lim) and np.all(np.abs(x).max(axis=0) > 0.95 * lim)
