"""The host side of the JPEG decoder parses untrusted bytes (markers, tables, EXIF, restart-marker scan, ROI): tools/fuzz builds it
with AddressSanitizer, replaces the CUDA runtime calls it makes by a stub and feeds it mutated streams.  A short run here; longer
ones with tools/fuzz/run.sh <iterations> <seeds...>."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_jpeg_host_parser_survives_mutated_streams(tmp_path):
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    asan = subprocess.run(["gcc", "-print-file-name=libasan.so"], capture_output=True, text=True).stdout.strip() if shutil.which("gcc") else ""
    if not os.path.exists(nvcc) or not os.path.isabs(asan) or not os.path.exists(asan):
        pytest.skip("needs nvcc and the AddressSanitizer runtime")
    env = dict(os.environ, FUZZ_DIR=str(tmp_path))
    r = subprocess.run([os.path.join(ROOT, "tools", "fuzz", "run.sh"), "4000", "5"], capture_output=True, text=True, timeout=900, env=env)
    assert r.returncode == 0 and "AddressSanitizer" not in r.stderr, (r.stdout[-2000:], r.stderr[-4000:])
    assert "no sanitizer report" in r.stdout


def test_plan_setups_survive_adversarial_arguments(tmp_path):
    """Every ...PlanSetup entry point under ASAN + UBSAN with adversarial arguments (tools/fuzz/run_setups.sh)."""
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    asan = subprocess.run(["gcc", "-print-file-name=libasan.so"], capture_output=True, text=True).stdout.strip() if shutil.which("gcc") else ""
    if not os.path.exists(nvcc) or not os.path.isabs(asan) or not os.path.exists(asan):
        pytest.skip("needs nvcc and the sanitizer runtimes")
    env = dict(os.environ, FUZZ_DIR=str(tmp_path))
    r = subprocess.run([os.path.join(ROOT, "tools", "fuzz", "run_setups.sh"), "400", "3"], capture_output=True, text=True, timeout=1500, env=env)
    assert r.returncode == 0 and "runtime error" not in r.stderr and "AddressSanitizer" not in r.stderr, (r.stdout[-2000:], r.stderr[-4000:])
    assert r.stdout.count("no sanitizer report") == 2
