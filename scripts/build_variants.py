"""Experiment builds of libpcl.so: headline kernels only (-DPCL_QUICK) plus extra -D flags.

    python scripts/build_variants.py m6:-DPCL_POLAR_MINB=6 m7:-DPCL_POLAR_MINB=7
writes _variants/libpcl_<name>.so (git-ignored, travels with gpurun); select one at run time with PCL_LIB=<path>.
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from polarcode_and_ldpc_b200 import _build  # noqa: E402

OUT = os.path.join(ROOT, "_variants")
os.makedirs(OUT, exist_ok=True)


def one(spec):
    name, _, flags = spec.partition(":")
    lib = os.path.join(OUT, f"libpcl_{name}.so")
    cmd = ["nvcc"] + _build.NVCC_FLAGS + ["-DPCL_QUICK"] + [f for f in flags.split(",") if f] + \
          ["-Xptxas", "-v", "-o", lib, os.path.join(_build.CSRC, "pcl_api.cu")]
    r = subprocess.run(cmd, capture_output=True, text=True)
    regs = [ln for ln in r.stderr.splitlines() if "registers" in ln]
    names = [ln for ln in r.stderr.splitlines() if "Compiling entry" in ln]
    info = {n.split("'")[1][:60]: g.split("Used ")[1].split(",")[0] for n, g in zip(names, regs)}
    return name, r.returncode, {k: v for k, v in info.items() if "fast" in k and "EfLi10" in k}, r.stderr[-500:] if r.returncode else ""


with ThreadPoolExecutor(8) as ex:
    for res in ex.map(one, sys.argv[1:]):
        print(res, flush=True)
