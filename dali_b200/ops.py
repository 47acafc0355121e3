"""Legacy object API, `nvidia.dali.ops` (dali/python/nvidia/dali/ops/__init__.py): an operator is configured in its constructor and
applied to DataNodes by calling the instance,

    decode = ops.decoders.Image(device="mixed", output_type=types.RGB)
    images = decode(jpegs)

inside `Pipeline.define_graph()` or a `@pipeline_def` function.  Every class forwards to the `fn` wrapper generated from the same
schema, so both APIs build identical graphs; constructor and call arguments are merged (call arguments win), per-sample tensor
arguments (DataNodes) are passed at call time as in the reference.
"""
import sys
import types as _pytypes

from . import fn as _fn


class _Operator:
    _fn_name = None            # dotted path of the functional wrapper below dali_b200.fn

    def __init__(self, **kwargs):
        self._init_args = kwargs

    @classmethod
    def _functional(cls):
        f = _fn
        for part in cls._fn_name.split("."):
            f = getattr(f, part)
        return f

    def __call__(self, *inputs, **kwargs):
        merged = dict(self._init_args)
        merged.update(kwargs)
        return type(self)._functional()(*inputs, **merged)


def _camel(snake):
    return "".join(p[:1].upper() + p[1:] for p in snake.split("_"))


def _make(module, class_name, fn_path, doc=None):
    cls = type(class_name, (_Operator,), {"_fn_name": fn_path, "__doc__": doc or f"ops counterpart of fn.{fn_path}", "__module__": module.__name__})
    setattr(module, class_name, cls)
    return cls


def _submodule(name):
    this = sys.modules[__name__]
    sub = getattr(this, name, None)
    if sub is None:
        sub = _pytypes.ModuleType(f"{__name__}.{name}")
        setattr(this, name, sub)
        sys.modules[sub.__name__] = sub
    return sub


def _install():
    from . import backend
    this = sys.modules[__name__]
    for schema in backend.schema_names():
        parts = schema.split("__")
        mod = this
        for p in parts[:-1]:
            mod = _submodule(p)
        f = _fn
        for p in parts[:-1]:
            f = getattr(f, p)
        snake = _fn._to_snake_case(parts[-1])
        _make(mod, parts[-1], ".".join(parts[:-1] + [snake]), getattr(f, snake).__doc__)
    # host-side sources and their legacy top-level names (ops/__init__.py deprecated aliases)
    _make(_submodule("readers"), "File", "readers.file")
    _make(this, "FileReader", "readers.file")
    _make(_submodule("random"), "CoinFlip", "random.coin_flip")
    _make(_submodule("random"), "Uniform", "random.uniform")
    _make(this, "CoinFlip", "random.coin_flip")
    _make(this, "Uniform", "random.uniform")
    _make(this, "ExternalSource", "external_source")
    if hasattr(this, "decoders"):
        _make(this, "ImageDecoder", "decoders.image")
        _make(this, "ImageDecoderCrop", "decoders.image_crop")
        _make(this, "ImageDecoderRandomCrop", "decoders.image_random_crop")
        _make(this, "ImageDecoderSlice", "decoders.image_slice")


_install()
