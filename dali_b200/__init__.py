"""dali_b200 -- Blackwell-native implementation of the DALI image / audio preprocessing hot path.

Python surface mirrors nvidia.dali for the covered operators:
    from dali_b200 import fn, types, pipeline_def, Pipeline
    from dali_b200.plugin.pytorch import DALIGenericIterator
(`import nvidia.dali` resolves to this package when the repository root is on sys.path.)
The native libraries are built in-tree by `python -m dali_b200.build`; nothing falls back to the CPU.
"""
__version__ = "0.1.0"


def __getattr__(name):
    # lazy: importing the package must not require the native libraries (build() imports it before they exist)
    import importlib
    if name in ("fn", "types", "pipeline", "backend", "capi", "hotpath", "plugin", "sharding", "plugin_manager", "readers", "ops"):
        return importlib.import_module(f"{__name__}.{name}")
    if name in ("Pipeline", "pipeline_def", "DataNode"):
        return getattr(importlib.import_module(f"{__name__}.pipeline"), name)
    raise AttributeError(name)
