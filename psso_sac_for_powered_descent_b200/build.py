"""Builds libpd_b200.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

    python -m psso_sac_for_powered_descent_b200.build [--force] [-v]

The translation units (fp64 kernels, fp32 kernels, shared-actor kernel, C ABI) are compiled
in parallel and linked into psso_sac_for_powered_descent_b200/libpd_b200.so.  The .so is
git-ignored but travels with the working tree to the GPU box.
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
# PD_LIB_PATH / PD_BUILD_DIR: experiment builds (tools/, -D knobs via PD_EXTRA_NVCC_FLAGS) that live
# next to the production library instead of replacing it
OUT = os.environ.get("PD_LIB_PATH") or os.path.join(HERE, "libpd_b200.so")
BUILD = os.environ.get("PD_BUILD_DIR") or os.path.join(HERE, "build")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
FLAGS = ["-std=c++17", "-O3", "-lineinfo", "-Xcompiler", "-fPIC", "-Xptxas", "-v"] + \
    os.environ.get("PD_EXTRA_NVCC_FLAGS", "").split()
UNITS = ["pd_fp64_roll.cu", "pd_fp32_roll.cu", "pd_fp64.cu", "pd_fp32.cu", "pd_api.cu", "pd_actor.cu", "pd_pso.cu", "pd_peak.cu", "pd_patch.cu"]


def _sources():
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    deps.append(os.path.join(os.path.dirname(HERE), "include", "pd_b200.h"))
    return deps


STAMP = os.path.join(BUILD, "flags.txt")


def _flags_changed():
    """Objects and library are only valid for the flags they were compiled with (PD_EXTRA_NVCC_FLAGS)."""
    try:
        return open(STAMP).read() != " ".join(FLAGS)
    except OSError:
        return True


def needs_build():
    if not os.path.exists(OUT) or _flags_changed():
        return True
    t = os.path.getmtime(OUT)
    return any(os.path.getmtime(d) > t for d in _sources())


def _compile(unit):
    obj = os.path.join(BUILD, unit.replace(".cu", ".o"))
    src = os.path.join(CSRC, unit)
    deps = _sources()
    if os.path.exists(obj) and all(os.path.getmtime(d) <= os.path.getmtime(obj) for d in deps):
        return obj, ""
    cmd = [NVCC, *ARCH, *FLAGS, "-c", src, "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed for {unit}:\n{r.stdout}\n{r.stderr}")
    with open(obj + ".ptxas.log", "w") as f:
        f.write(r.stderr)
    return obj, r.stderr


def build(force=False, verbose=False):
    if not force and not needs_build():
        return OUT
    os.makedirs(BUILD, exist_ok=True)
    if _flags_changed():
        for f in os.listdir(BUILD):
            if f.endswith(".o"):
                os.remove(os.path.join(BUILD, f))
    units = [u for u in UNITS if os.path.exists(os.path.join(CSRC, u))]
    with ThreadPoolExecutor(max_workers=len(units)) as ex:
        res = list(ex.map(_compile, units))
    objs = [o for o, _ in res]
    cmd = [NVCC, *ARCH, "-shared", "-o", OUT, *objs, "-lcudart"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    with open(STAMP, "w") as f:
        f.write(" ".join(FLAGS))
    if verbose:
        for _, log in res:
            sys.stderr.write(log)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
