"""`nvidia.dali` alias of dali_b200, so that existing pipelines import unchanged when this repository is on sys.path."""
import sys as _sys

import dali_b200 as _impl
from dali_b200 import fn, types, pipeline, backend, ops  # noqa: F401
from dali_b200.pipeline import Pipeline, pipeline_def, DataNode  # noqa: F401
import dali_b200.plugin.pytorch as _pt
from dali_b200 import plugin_manager  # noqa: F401

_sys.modules[__name__ + ".fn"] = fn
_sys.modules[__name__ + ".types"] = types
_sys.modules[__name__ + ".ops"] = ops
_sys.modules[__name__ + ".pipeline"] = pipeline
_sys.modules[__name__ + ".plugin"] = _impl.plugin
_sys.modules[__name__ + ".plugin.pytorch"] = _pt
_sys.modules[__name__ + ".plugin_manager"] = plugin_manager
__version__ = _impl.__version__
