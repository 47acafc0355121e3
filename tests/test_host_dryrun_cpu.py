"""Host dry run of the -m gpu tests: the CUDA runtime calls are replaced by tools/fuzz/cuda_stub.c (preloaded) and tests/conftest.py
maps "device" tensors to host memory (DALIB200_DRYRUN=1).  Kernels and copies do nothing, so value comparisons cannot hold -- an
AssertionError counts as passed -- but every host-side step of every GPU test runs here: argument validation, plan set-up, descriptor
build, staging of the encoded streams, the executor with its prefetch slots, the reader's read-ahead thread and page-locked arenas,
seed assignment, the iterator.  Any other exception (an argument the library now rejects, a C++ exception, a Python error) fails."""
import os
import re
import shutil
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_gpu_tests_run_their_host_side_without_a_gpu(tmp_path):
    if shutil.which("gcc") is None:
        pytest.skip("needs gcc for the CUDA runtime stub")
    stub = str(tmp_path / "cuda_stub.so")
    subprocess.run(["gcc", "-shared", "-fPIC", "-O1", "-o", stub, os.path.join(ROOT, "tools", "fuzz", "cuda_stub.c")], check=True)
    env = dict(os.environ, DALIB200_DRYRUN="1", LD_PRELOAD=stub)
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(ROOT, "tests"), "-q", "-m", "gpu", "-p", "no:cacheprovider"],
                       capture_output=True, text=True, timeout=1500, env=env, cwd=ROOT)
    tail = r.stdout[-3000:] + r.stderr[-2000:]
    assert r.returncode == 0, tail
    m = re.search(r"(\d+) passed", r.stdout)
    assert m and int(m.group(1)) >= 60 and "failed" not in r.stdout.splitlines()[-1], tail
