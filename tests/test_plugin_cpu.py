"""plugin_manager.load_library (dali/plugin/plugin_manager.cc:26-41, python/nvidia/dali/plugin_manager.py): a plugin built against the
host headers registers its operators on load and fn.* picks them up."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def build_plugin(tmpdir):
    out = os.path.join(str(tmpdir), "libdali_customdummy.so")
    libdir = os.path.join(ROOT, "dali_b200", "lib")
    cmd = ["g++", "-std=c++17", "-shared", "-fPIC", os.path.join(ROOT, "tests", "plugin_src", "custom_dummy.cc"), "-o", out,
           "-I" + os.path.join(ROOT, "dali_b200", "host"), "-I/usr/local/cuda/include", "-L" + libdir, "-ldali_b200_host",
           "-L/usr/local/cuda/lib64", "-lcudart", "-Wl,-rpath," + libdir, "-Wl,-rpath,/usr/local/cuda/lib64"]
    subprocess.check_call(cmd)
    return out


def test_load_library_registers_operator(tmp_path):
    lib = build_plugin(tmp_path)
    code = f"""
import sys
sys.path.insert(0, {ROOT!r})
import nvidia.dali as dali
from nvidia.dali import fn, plugin_manager
assert not hasattr(fn, "custom_dummy")
plugin_manager.load_library({lib!r})
assert hasattr(fn, "custom_dummy") and fn.custom_dummy.schema_name == "CustomDummy"
plugin_manager.load_library({lib!r})            # idempotent
try:
    plugin_manager.load_library("/nonexistent/libx.so")
except RuntimeError as e:
    assert "Failed to load library" in str(e)
else:
    raise AssertionError("expected RuntimeError")
print("ok")
"""
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and r.stdout.strip().endswith("ok"), r.stderr[-2000:]


@pytest.mark.gpu
def test_plugin_operator_runs_in_a_pipeline(tmp_path):
    import numpy as np
    from dali_b200 import fn, pipeline_def, plugin_manager
    plugin_manager.load_library(build_plugin(tmp_path))
    data = [np.arange(24, dtype=np.uint8).reshape(2, 4, 3) + i for i in range(3)]

    @pipeline_def(batch_size=3, num_threads=1, device_id=0)
    def pipe():
        return fn.custom_dummy(fn.external_source(source=lambda i: data, device="gpu", layout="HWC"))
    p = pipe()
    p.build()
    (out,) = p.run()
    for i, o in enumerate(out.as_cpu()):
        assert np.array_equal(o, data[i])
