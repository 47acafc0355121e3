"""nvidia.dali.fn for the hot-path operators: wrappers generated from the C++ schema registry, as the reference
generates them from its schemas (dali/python/nvidia/dali/fn/__init__.py:66-115, ops/_names.py:24-72:
`decoders__Image` -> fn.decoders.image, `CropMirrorNormalize` -> fn.crop_mirror_normalize).
"""
import re
import sys
import types as _pytypes

import numpy as np

from . import backend
from . import types as _types
from .pipeline import DataNode, _current, _ExternalSourceGroup


def _to_snake_case(name):
    out = name[0].lower()
    for i in range(1, len(name)):
        c = name[i]
        if c.isupper():
            prev, nxt = name[i - 1], name[i + 1] if i + 1 < len(name) else ""
            if not prev.isupper() or (nxt and nxt.islower()):
                out += "_"
            out += c.lower()
        else:
            out += c
    return out


_INPUT_DEVICE = {"mixed": "cpu", "cpu": "cpu", "gpu": "gpu"}


def _make_wrapper(schema):
    args_info = backend.schema_args(schema)
    min_in, max_in, num_out, doc = backend.schema_info(schema)
    has_mixed = backend.operator_registered(schema, "mixed")
    has_gpu = backend.operator_registered(schema, "gpu")

    def op(*inputs, device=None, name=None, **kwargs):
        pipe = _current()
        if pipe is None:
            raise RuntimeError(f"fn.{_to_snake_case(schema.split('__')[-1])} must be called inside a pipeline definition "
                               "(@pipeline_def function or `with pipeline:` block)")
        if not (min_in <= len(inputs) <= max_in):
            raise ValueError(f"Operator {schema} expects {min_in}..{max_in} inputs, got {len(inputs)}")
        for i in inputs:
            if not isinstance(i, DataNode):
                raise TypeError(f"Operator inputs must be DataNodes, got {type(i).__name__}. Wrap constants with fn.external_source.")
        if device is None:
            # as in the reference: the device follows the first input (gpu input -> gpu operator)
            device = "gpu" if (inputs and inputs[0].device == "gpu" and has_gpu) else ("mixed" if has_mixed and not has_gpu else "cpu")
        if device not in ("cpu", "gpu", "mixed"):
            raise ValueError(f"Invalid device \"{device}\". Valid options are \"cpu\", \"gpu\" or \"mixed\"")
        spec = backend.OpSpec(schema)
        spec.add_arg("device", device)
        in_dev = _INPUT_DEVICE[device]
        for i in inputs:
            if i.device != in_dev:
                if i.device == "cpu" and in_dev == "gpu":
                    i = i.gpu()
                else:
                    raise ValueError(f"Operator {schema} on device '{device}' expects {in_dev} inputs, got a {i.device} input")
            if in_dev == "cpu":
                i._consumed_cpu = True
            spec.add_input(i.name, i.device)
        for k, v in kwargs.items():
            if k not in args_info:
                raise TypeError(f"Operator {schema} got an unexpected '{k}' argument")   # same wording family as the reference
            if v is None:
                continue
            if isinstance(v, DataNode):
                if not args_info[k][0]:
                    raise TypeError(f"Argument '{k}' of operator {schema} does not accept a DataNode (per-sample tensor) input")
                if v.device != "cpu":
                    raise ValueError(f"Argument inputs must be CPU DataNodes ('{k}')")
                v._consumed_cpu = True
                spec.add_argument_input(k, v.name)
            else:
                if hasattr(v, "value") and not isinstance(v, (int, float)):
                    v = v.value
                if isinstance(v, (_types.DALIDataType, _types.DALIInterpType, _types.DALIImageType)):
                    v = int(v)
                spec.add_arg(k, v)
        inst = name or pipe._new_name(schema)
        out_dev = "cpu" if device == "cpu" else "gpu"
        outs = [DataNode(f"{inst}[{k}]" if num_out > 1 else inst, out_dev) for k in range(num_out)]
        for o in outs:
            spec.add_output(o.name, o.device)
        pipe._nodes.append((schema, inst, spec))
        return outs[0] if num_out == 1 else tuple(outs)

    op.__name__ = _to_snake_case(schema.split("__")[-1])
    op.__doc__ = doc + "\n\nArguments: " + ", ".join(sorted(a for a in args_info if a not in ("num_threads", "max_batch_size", "seed", "preserve", "bytes_per_sample_hint")))
    op.schema_name = schema
    return op


def external_source(source=None, num_outputs=None, *, cycle=None, name=None, device="cpu", layout=None, dtype=None, ndim=None,
                    batch=True, batch_info=False, parallel=False, no_copy=None, prefetch_queue_depth=None, cuda_stream=None,
                    use_copy_kernel=None, blocking=None, repeat_last=False, **_ignored):
    """fn.external_source (dali/python/nvidia/dali/external_source.py:1002-1150), subset: callables, iterables, feed_input."""
    pipe = _current()
    if pipe is None:
        raise RuntimeError("fn.external_source must be called inside a pipeline definition")
    if device not in ("cpu", "gpu"):
        raise ValueError("external_source device must be 'cpu' or 'gpu'")
    n = num_outputs or 1
    base = name or pipe._new_name("ExternalSource")
    g = _ExternalSourceGroup(source, [], batch, cycle, layout if isinstance(layout, str) or layout is None else layout[0], dtype, device,
                             batch_info)
    for k in range(n):
        node = DataNode(base if n == 1 else f"{base}[{k}]", device, source=g)
        g.outputs.append(node)
        if name:
            pipe._ext_names[name if n == 1 else f"{name}[{k}]"] = node.name
    pipe._externals.append(g)
    return g.outputs[0] if num_outputs is None else list(g.outputs)


def _install():
    this = sys.modules[__name__]
    for schema in backend.schema_names():
        parts = schema.split("__")
        mod = this
        for p in parts[:-1]:
            sub = getattr(mod, p, None)
            if sub is None:
                sub = _pytypes.ModuleType(f"{mod.__name__}.{p}")
                setattr(mod, p, sub)
                sys.modules[sub.__name__] = sub
            mod = sub
        setattr(mod, _to_snake_case(parts[-1]), _make_wrapper(schema))
    # deprecated aliases kept by the reference (mixed_decoder.cc:40-50)
    if hasattr(this, "decoders"):
        this.image_decoder = this.decoders.image


_install()
