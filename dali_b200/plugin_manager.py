"""nvidia.dali.plugin_manager for dali_b200 (dali/python/nvidia/dali/plugin_manager.py:19-112, dali/plugin/plugin_manager.cc:26-41).

A plugin is a shared library whose static DALI_SCHEMA / DALI_REGISTER_OPERATOR objects register operators when it is loaded
(written against dali_b200/host/dali.h -- the same class and macro names as the reference's operator.h -- and linked against
libdali_b200_host.so, as a DALI plugin links against libdali.so).  `load_library` dlopens it (RTLD_LAZY | RTLD_LOCAL, like
PluginManager::LoadLibrary) and regenerates the fn.* wrappers from the schema registry (ops.Reload() in the reference).
"""
import ctypes as C
import glob
import os

from . import backend

_loaded = {}


def load_library(library_path, global_symbols=False):
    """Loads a plugin library and exposes the operators it registers under dali_b200.fn (fn.<snake_case name>)."""
    path = os.path.abspath(library_path)
    if not os.path.exists(path):
        raise RuntimeError(f"Failed to load library: {library_path}: no such file")
    if path in _loaded:
        return
    backend.lib()            # the host library (registries) must be resident first
    mode = (C.RTLD_GLOBAL if global_symbols else C.RTLD_LOCAL) | os.RTLD_LAZY
    try:
        _loaded[path] = C.CDLL(path, mode=mode)
    except OSError as e:
        raise RuntimeError(f"Failed to load library: {e}") from e
    from . import fn
    fn._install()


def load_directory(plugin_dir_path, global_symbols=False, ignore_errors=False):
    """Loads every libdali_*.so below a directory (plugin_manager.py:37-68)."""
    for path in sorted(glob.glob(os.path.join(plugin_dir_path, "**", "libdali_*.so"), recursive=True)):
        try:
            load_library(path, global_symbols=global_symbols)
        except RuntimeError:
            if not ignore_errors:
                raise


def load_plugins():
    """DALI_PRELOAD_PLUGINS: colon-separated libraries / directories loaded at import time (plugin_manager.py:98-112)."""
    for item in filter(None, os.environ.get("DALI_PRELOAD_PLUGINS", "").split(":")):
        if os.path.isdir(item):
            load_directory(item, ignore_errors=True)
        else:
            load_library(item)
