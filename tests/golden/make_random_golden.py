"""Generates tests/golden/random_ref.npz from the reference's own random number classes (oracle/_ref, built by `make -C oracle ref`
where /root/reference exists) and from the C++ standard library's seed_seq.  Run from the repository root:
    python tests/golden/make_random_golden.py
"""
import os
import subprocess
import sys
import tempfile

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
from oracle import pyoracle as po                      # noqa: E402
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from test_random_ref_cpu import CASES_FLIP, CASES_UNIFORM, GOLDEN   # noqa: E402

SEED_SEQ_CC = r"""
#include <cstdint>
#include <cstdio>
#include <random>
#include <vector>
int main() {                       // what Pipeline::Init does with its seed (dali/pipeline/pipeline.cc:303-308)
  std::vector<int64_t> seed(1024);
  std::seed_seq ss{int64_t(7)};
  ss.generate(seed.begin(), seed.end());
  for (int i = 0; i < 16; i++) std::printf("%lld\n", static_cast<long long>(seed[i]));
}
"""


def main():
    assert po.have_ref(), "build oracle/_ref first (make -C oracle ref)"
    out = {}
    for k, (seed, it, b, vol, p, dt) in enumerate(CASES_FLIP):
        out[f"flip{k}"] = po.ref_random_coin_flip(seed, it, b, vol, p, dt)
    for k, (seed, it, b, vol, rng, vals, dt) in enumerate(CASES_UNIFORM):
        out[f"uniform{k}"] = po.ref_random_uniform(seed, it, b, vol, rng, vals, dt)
    with tempfile.TemporaryDirectory() as d:
        src, exe = os.path.join(d, "s.cc"), os.path.join(d, "s")
        open(src, "w").write(SEED_SEQ_CC)
        subprocess.run([os.environ.get("CXX", "g++"), "-std=c++17", "-O1", src, "-o", exe], check=True)
        out["seed_table_7"] = np.array([int(v) for v in subprocess.run([exe], check=True, capture_output=True, text=True).stdout.split()], np.int64)
    np.savez_compressed(GOLDEN, **out)
    print("wrote", GOLDEN, {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
