"""Build a variant of libdcrf_b200.so with extra -D flags into build/variants/<name>/ for A/B timing on the GPU box:
  [VARIANT_SRC=filter,filter_wide] python profiles/build_variant.py <name> [-DFOO=1 ...]   (sources rebuilt with the flags)
  LD_LIBRARY_PATH=build/variants/<name> python profiles/quick_bench.py <name>
(the binding finds the library through LD_LIBRARY_PATH before its RUNPATH)."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as g
name, defs = sys.argv[1], sys.argv[2:]
out = os.path.join(ROOT, "build", "variants", name)
os.makedirs(out, exist_ok=True)
objs = []
for n in ("lattice", "filter", "filter_wide", "matrix_tc", "matrix_tc_tma", "learn", "api"):
    src, obj = os.path.join(g.CSRC, n + ".cu"), os.path.join(out, n + ".o")
    if n in os.environ.get("VARIANT_SRC", "filter,filter_wide").split(",") and defs:
        subprocess.check_call(["nvcc"] + g.NVCC_FLAGS + defs + ["-c", src, "-o", obj])
    else:
        obj = os.path.join(g.BUILD, n + ".o")
    objs.append(obj)
subprocess.check_call(["nvcc", "-shared", "-cudart", "static", "-gencode", "arch=compute_100a,code=sm_100a",
                       "-Wno-deprecated-gpu-targets", "-o", os.path.join(out, "libdcrf_b200.so")] + objs)
print("built", out)
