"""Builds the in-tree native libraries (sm_100a only).

  dali_b200/lib/libdali_b200.so        CUDA kernels + the C-ABI of include/dali_b200.h   (nvcc)

`python -m dali_b200.build` or __graft_entry__.build().  nvcc cross-compiles without a GPU.
"""
import concurrent.futures as cf
import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
OBJDIR = os.path.join(HERE, "build")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr", "-fmad=false", "-Xptxas", "-v"]


def _newer(src, dst, extra=()):
    if not os.path.exists(dst):
        return True
    t = os.path.getmtime(dst)
    return any(os.path.getmtime(s) > t for s in (src,) + tuple(extra))


def build_kernels(verbose=False, force=False):
    os.makedirs(LIBDIR, exist_ok=True)
    os.makedirs(OBJDIR, exist_ok=True)
    srcs = sorted(glob.glob(os.path.join(CSRC, "*.cu")))
    hdrs = tuple(glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(CSRC, "*.h")) +
                 glob.glob(os.path.join(HERE, "..", "include", "*.h")))
    objs = []

    def compile_one(src):
        obj = os.path.join(OBJDIR, os.path.basename(src) + ".o")
        if force or _newer(src, obj, hdrs):
            cmd = [NVCC] + NVCC_FLAGS + ["-c", src, "-o", obj]
            r = subprocess.run(cmd, capture_output=True, text=True)
            log = r.stdout + r.stderr
            with open(obj + ".log", "w") as f:
                f.write(log)
            if r.returncode != 0:
                raise RuntimeError(f"nvcc failed for {src}:\n{log[-4000:]}")
            if verbose:
                print(log)
        return obj

    with cf.ThreadPoolExecutor(max_workers=min(8, len(srcs) or 1)) as ex:
        objs = list(ex.map(compile_one, srcs))
    lib = os.path.join(LIBDIR, "libdali_b200.so")
    if force or any(_newer(o, lib) for o in objs):
        cmd = [NVCC, "-shared", "-o", lib] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-lcudart"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n" + r.stdout + r.stderr)
    return lib


def build_host(verbose=False, force=False):
    """libdali_b200_host.so: the dali:: operator boundary, the operators and the pipeline C API (g++; links the kernel library)."""
    hdir = os.path.join(HERE, "host")
    srcs = sorted(glob.glob(os.path.join(hdir, "*.cc")))
    hdrs = tuple(glob.glob(os.path.join(hdir, "*.h")) + glob.glob(os.path.join(HERE, "..", "include", "*.h")))
    os.makedirs(OBJDIR, exist_ok=True)
    cxx = os.environ.get("CXX", "g++")
    flags = ["-std=c++17", "-O2", "-fPIC", "-Wall", "-Wno-unused-function", "-I/usr/local/cuda/include"]

    def compile_one(src):
        obj = os.path.join(OBJDIR, "host_" + os.path.basename(src) + ".o")
        if force or _newer(src, obj, hdrs):
            r = subprocess.run([cxx] + flags + ["-c", src, "-o", obj], capture_output=True, text=True)
            if r.returncode != 0:
                raise RuntimeError(f"{cxx} failed for {src}:\n{(r.stdout + r.stderr)[-4000:]}")
            if verbose and r.stderr:
                print(r.stderr)
        return obj

    with cf.ThreadPoolExecutor(max_workers=4) as ex:
        objs = list(ex.map(compile_one, srcs))
    lib = os.path.join(LIBDIR, "libdali_b200_host.so")
    klib = os.path.join(LIBDIR, "libdali_b200.so")
    if force or any(_newer(o, lib) for o in objs) or _newer(klib, lib):
        cmd = [cxx, "-shared", "-o", lib] + objs + ["-L" + LIBDIR, "-ldali_b200", "-L/usr/local/cuda/lib64", "-lcudart",
                                                    "-Wl,-rpath,$ORIGIN", "-Wl,-rpath,/usr/local/cuda/lib64", "-Wl,--no-undefined"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("host link failed:\n" + r.stdout + r.stderr)
    return lib


def build_all(verbose=False, force=False):
    k = build_kernels(verbose=verbose, force=force)
    build_host(verbose=verbose, force=force)
    return k


if __name__ == "__main__":
    print(build_all(verbose="-v" in sys.argv, force="-f" in sys.argv))
