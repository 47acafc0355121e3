"""Experiment helper: builds dali_b200/lib/libdali_b200_<tag>.so with extra nvcc defines (A/B runs through DALIB200_LIB).
  python tools/build_variant.py nofence -DDALIB200_NO_RING_FENCE"""
import glob, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dali_b200 import build as b
tag, defs = sys.argv[1], sys.argv[2:]
objdir = os.path.join(b.HERE, "build", "variant_" + tag)
os.makedirs(objdir, exist_ok=True)
objs = []
for src in sorted(glob.glob(os.path.join(b.CSRC, "*.cu"))):
    obj = os.path.join(objdir, os.path.basename(src) + ".o")
    subprocess.check_call([b.NVCC] + [f for f in b.NVCC_FLAGS if f not in ("-Xptxas", "-v")] + defs + ["-c", src, "-o", obj])
    objs.append(obj)
lib = os.path.join(b.LIBDIR, f"libdali_b200_{tag}.so")
subprocess.check_call([b.NVCC, "-shared", "-o", lib] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-lcudart"])
print(lib)
