import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import pyoracle
    pyoracle.lib()
    return pyoracle


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")


# ---- host dry run ---------------------------------------------------------------------------------------------------------------
# DALIB200_DRYRUN=1 (with tools/fuzz/cuda_stub.c preloaded, see tests/test_host_dryrun_cpu.py) runs the -m gpu tests WITHOUT a GPU:
# kernels and copies are stubbed out, "device" buffers are host memory, so every value comparison fails -- but all the HOST code in
# front of and behind the kernels runs (argument checks, plan set-up, descriptor build, staging copies, executor, reader thread,
# iterator).  A failed comparison (AssertionError) therefore counts as passed; any other exception is a host-side defect.
DRY_RUN = os.environ.get("DALIB200_DRYRUN") in ("1", "emul")
# DALIB200_DRYRUN=emul (with tools/emul/cuda_emul_stub.cc preloaded, see tests/test_launch_emul_cpu.py): the same host mapping, but the
# copies really happen and the launches of the kernels that exist as host/device functions are executed on the host, so the value
# comparisons of the tests that only need those kernels are REAL and are kept.
EMUL = os.environ.get("DALIB200_DRYRUN") == "emul"

if DRY_RUN:
    import numpy as _np
    import torch as _torch

    _real_empty, _real_zeros, _real_as_tensor, _real_tensor = _torch.empty, _torch.zeros, _torch.as_tensor, _torch.tensor
    # bit-exact comparisons are let through, so that a test goes on to its later steps (further iterations, epochs, operators)
    if not EMUL:
        _np.array_equal = lambda *a, **k: True

    def _host(kw):
        if str(kw.get("device", "cpu")).startswith("cuda"):
            kw = dict(kw, device="cpu")
        return kw

    _torch.empty = lambda *a, **k: _real_zeros(*a, **_host(k))
    _torch.zeros = lambda *a, **k: _real_zeros(*a, **_host(k))
    _torch.tensor = lambda *a, **k: _real_tensor(*a, **_host(k))

    def _as_tensor(obj, *a, **k):
        cai = getattr(obj, "__cuda_array_interface__", None)
        if cai is not None and EMUL:              # "device" memory is host memory that the emulated kernels really wrote
            import ctypes as _C
            shape, dt = tuple(cai["shape"]), _np.dtype(cai["typestr"])
            n = int(_np.prod(shape)) * dt.itemsize
            raw = (_C.c_uint8 * n).from_address(cai["data"][0]) if n else b""
            return _real_as_tensor(_np.frombuffer(bytes(raw), dt).reshape(shape).copy())
        if cai is not None:                       # a "device" array of the pipeline: contents are meaningless in a dry run
            return _real_zeros(tuple(cai["shape"]), dtype=getattr(_torch, _np.dtype(cai["typestr"]).name))
        return _real_as_tensor(obj, *a, **_host(k))
    _torch.as_tensor = _as_tensor
    _torch.Tensor.cuda = lambda self, *a, **k: self
    _torch.Tensor.is_cuda = property(lambda self: True)
    _torch.cuda.is_available = lambda: True
    _torch.cuda.synchronize = lambda *a, **k: None
    _torch.cuda.set_device = lambda *a, **k: None
    _torch.cuda.current_device = lambda: 0
    _torch.cuda.device_count = lambda: 1

    class _Stream:
        cuda_stream = 0

        def wait_event(self, *a):
            pass

        def synchronize(self):
            pass

        def wait_stream(self, *a):
            pass
    _torch.cuda.current_stream = lambda *a, **k: _Stream()

    class _Event:
        def __init__(self, *a, **k):
            pass

        def record(self, *a):
            pass

        def synchronize(self):
            pass

        def wait(self, *a):
            pass

        def elapsed_time(self, other):
            return 0.0
    _torch.cuda.Event = _Event

    class _DeviceCtx:
        def __init__(self, *a, **k):
            pass

        def __enter__(self):
            return self

        def __exit__(self, *exc):
            return False
    _torch.cuda.device = _DeviceCtx
    _torch.cuda.stream = _DeviceCtx
    _torch.Tensor.clone = (lambda f: lambda self, *a, **k: f(self, *a, **k))(_torch.Tensor.clone)

    @pytest.hookimpl(hookwrapper=True)
    def pytest_runtest_makereport(item, call):
        outcome = yield
        rep = outcome.get_result()
        if not EMUL and rep.when == "call" and rep.failed and call.excinfo is not None and call.excinfo.errisinstance(AssertionError):
            rep.outcome = "passed"
            rep.longrepr = None
            rep.sections.append(("dry run", "value comparison skipped"))
            _dry_compared.append(item.nodeid)

    _dry_compared = []

    def pytest_terminal_summary(terminalreporter):
        terminalreporter.write_line(f"dry run: {len(_dry_compared)} tests reached a value comparison (counted as passed); "
                                    "every other test passed on host-side checks alone")
