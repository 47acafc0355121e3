"""ctypes binding of libdali_b200_host.so (dali_b200/host/c_api.cc): OpSpec / schema registry / Pipeline.

Plays the role of the reference's pybind module (dali/python/backend_impl.cc: Pipeline :2475, OpSpec, schema access).
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
HOST_LIB_PATH = os.path.join(_HERE, "lib", "libdali_b200_host.so")
_lib = None


class BackendError(RuntimeError):
    pass


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(HOST_LIB_PATH):
            raise BackendError(f"{HOST_LIB_PATH} is missing: run `python -m dali_b200.build`")
        # the kernel library must be resolvable first (rpath $ORIGIN covers the in-tree layout)
        C.CDLL(os.path.join(_HERE, "lib", "libdali_b200.so"), mode=C.RTLD_GLOBAL)
        _lib = C.CDLL(HOST_LIB_PATH)
        _lib.dalihLastError.restype = C.c_char_p
    return _lib


def check(rc):
    if rc != 0:
        msg = lib().dalihLastError().decode("utf-8", "replace")
        raise BackendError(msg)


def _b(s):
    return s.encode("utf-8") if isinstance(s, str) else s


def schema_names():
    n = lib().dalihNumSchemas()
    buf = C.create_string_buffer(256)
    out = []
    for i in range(n):
        check(lib().dalihSchemaName(i, buf, 256))
        out.append(buf.value.decode())
    return out


def schema_args(name):
    """{arg: (tensor_ok, required)}"""
    buf = C.create_string_buffer(16384)
    check(lib().dalihSchemaArgs(_b(name), buf, 16384))
    res = {}
    for ln in buf.value.decode().splitlines():
        a, t, r = ln.split("|")
        res[a] = (t == "1", r == "1")
    return res


def schema_info(name):
    mi, ma, no = C.c_int(), C.c_int(), C.c_int()
    doc = C.create_string_buffer(4096)
    check(lib().dalihSchemaInfo(_b(name), C.byref(mi), C.byref(ma), C.byref(no), doc, 4096))
    return mi.value, ma.value, no.value, doc.value.decode()


def operator_registered(name, backend):
    return bool(lib().dalihOperatorRegistered(_b(name), _b(backend)))


class OpSpec:
    def __init__(self, schema):
        self.h = C.c_void_p()
        check(lib().dalihOpSpecCreate(C.byref(self.h), _b(schema)))

    def __del__(self):
        try:
            if self.h:
                lib().dalihOpSpecDestroy(self.h)
        except Exception:
            pass

    def add_arg(self, name, v):
        import numpy as np
        l, n = lib(), _b(name)
        if isinstance(v, bool):
            check(l.dalihOpSpecAddArgBool(self.h, n, int(v)))
        elif isinstance(v, (int, np.integer)):
            check(l.dalihOpSpecAddArgInt(self.h, n, C.c_int64(int(v))))
        elif isinstance(v, (float, np.floating)):
            check(l.dalihOpSpecAddArgFloat(self.h, n, C.c_double(float(v))))
        elif isinstance(v, str):
            check(l.dalihOpSpecAddArgString(self.h, n, _b(v)))
        elif isinstance(v, (list, tuple, np.ndarray)):
            f = np.ascontiguousarray(np.asarray(v), dtype=np.float32).ravel()
            check(l.dalihOpSpecAddArgFloatVec(self.h, n, f.ctypes.data_as(C.c_void_p), int(f.size)))
        else:
            raise TypeError(f"Unsupported value for argument '{name}': {type(v).__name__}")

    def add_input(self, name, device):
        check(lib().dalihOpSpecAddInput(self.h, _b(name), _b(device)))

    def add_output(self, name, device):
        check(lib().dalihOpSpecAddOutput(self.h, _b(name), _b(device)))

    def add_argument_input(self, arg, input_name):
        check(lib().dalihOpSpecAddArgumentInput(self.h, _b(arg), _b(input_name)))


class Pipeline:
    def __init__(self, max_batch, num_threads, device_id):
        self.h = C.c_void_p()
        check(lib().dalihPipelineCreate(C.byref(self.h), int(max_batch), int(num_threads), -1 if device_id is None else int(device_id)))

    def __del__(self):
        try:
            if self.h:
                lib().dalihPipelineDestroy(self.h)
                self.h = C.c_void_p()
        except Exception:
            pass

    def add_external_input(self, name, device, layout=""):
        check(lib().dalihPipelineAddExternalInput(self.h, _b(name), _b(device), _b(layout or "")))

    def add_operator(self, spec, inst_name):
        check(lib().dalihPipelineAddOperator(self.h, spec.h, _b(inst_name)))

    def set_outputs(self, outs):
        n = len(outs)
        names = (C.c_char_p * n)(*[_b(o[0]) for o in outs])
        devs = (C.c_char_p * n)(*[_b(o[1]) for o in outs])
        check(lib().dalihPipelineSetOutputs(self.h, n, names, devs))

    def build(self):
        check(lib().dalihPipelineBuild(self.h))

    def feed_input(self, name, ptrs, shapes, ndim, dtype, layout="", no_copy=False):
        import numpy as np
        n = len(ptrs)
        p = (C.c_void_p * n)(*ptrs)
        sh = np.ascontiguousarray(shapes, dtype=np.int64).reshape(-1)
        check(lib().dalihPipelineFeedInputEx(self.h, _b(name), n, p, sh.ctypes.data_as(C.c_void_p), int(ndim), int(dtype), _b(layout or ""),
                                             int(bool(no_copy))))

    def run(self):
        check(lib().dalihPipelineRun(self.h))

    def wait(self):
        check(lib().dalihPipelineWait(self.h))

    def num_outputs(self):
        return lib().dalihPipelineNumOutputs(self.h)

    def stream(self):
        s = C.c_void_p()
        check(lib().dalihPipelineStream(self.h, C.byref(s)))
        return s.value or 0

    def output(self, i):
        """(is_gpu, dtype, layout, contiguous, shapes[n][ndim], ptrs[n])"""
        import numpy as np
        g, n, nd, dt, cont = C.c_int(), C.c_int(), C.c_int(), C.c_int(), C.c_int()
        lay = C.create_string_buffer(16)
        check(lib().dalihPipelineOutputInfo(self.h, i, C.byref(g), C.byref(n), C.byref(nd), C.byref(dt), lay, C.byref(cont)))
        shapes = np.zeros((n.value, nd.value), np.int64)
        ptrs = (C.c_void_p * max(n.value, 1))()
        check(lib().dalihPipelineOutputData(self.h, i, shapes.ctypes.data_as(C.c_void_p), ptrs))
        return bool(g.value), dt.value, lay.value.decode(), bool(cont.value), shapes, [ptrs[k] or 0 for k in range(n.value)]
