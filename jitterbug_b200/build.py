"""Builds ``jitterbug_b200/_lib/libjbsim.so`` (CUDA, sm_100a) in-tree with nvcc.

Seven objects: the fused quad-per-env step kernel in its two instantiations (shared model / per-env models), the C-ABI +
utility kernels, the observation encoder and the on-device model compiler; they compile in parallel.  ``python -m jitterbug_b200.build`` or ``build()``.
"""

from __future__ import annotations

import concurrent.futures
import hashlib
import os
import subprocess
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
LIBDIR = os.path.join(_HERE, "_lib")
LIB = os.path.join(LIBDIR, "libjbsim.so")
# (source, object stem, extra flags): the step kernel is instantiated twice, see the head of jb_sim_kernel.cu
UNITS = [("jb_sim_kernel.cu", "jb_sim_kernel", []), ("jb_sim_kernel.cu", "jb_sim_kernel_perenv", ["-DJB_PER_ENV_BUILD"]),
         ("jb_api.cu", "jb_api", []), ("jb_sort.cu", "jb_sort", []), ("jb_monitor.cu", "jb_monitor", []), ("jb_encoder.cu", "jb_encoder", []), ("jb_augment.cu", "jb_augment", [])]
# -prec-div/-prec-sqrt=false: 2-ulp reciprocal-based division instead of the IEEE slow path (bulky and not needed at the
# 1e-5 parity tolerance); -ftz=true: fp32 denormals flush to zero, which drops the denormal-scaling wrappers around the
# reciprocal / rsqrt sequences (-3 % instructions, +3 % throughput; no physical quantity here lives near 1e-38);
# sin/cos stay the accurate versions (no --use_fast_math).
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-prec-div=false", "-prec-sqrt=false", "-ftz=true", "-Xcompiler", "-fPIC", "-Xptxas", "-v"] + \
    os.environ.get("JB_NVCC_EXTRA", "").split()


def _sources():
    out = [os.path.join(_HERE, "..", "include", "jbsim.h")]
    for f in sorted(os.listdir(CSRC)):
        if f.endswith((".cu", ".cuh", ".h")):
            out.append(os.path.join(CSRC, f))
    return out


def _digest() -> str:
    h = hashlib.sha256()
    for s in _sources():
        with open(s, "rb") as f:
            h.update(f.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(LIBDIR, exist_ok=True)
    stamp = os.path.join(LIBDIR, "build.sha256")
    dig = _digest()
    if not force and os.path.exists(LIB) and os.path.exists(stamp) and open(stamp).read().strip() == dig:
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")

    def compile_one(spec):
        unit, stem, extra = spec
        obj = os.path.join(LIBDIR, stem + ".o")
        cmd = [nvcc] + NVCC_FLAGS + extra + ["-c", os.path.join(CSRC, unit), "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        log = os.path.join(LIBDIR, stem + ".ptxas.log")
        with open(log, "w") as f:
            f.write(r.stdout + r.stderr)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed on {unit}:\n{r.stdout}\n{r.stderr}")
        if verbose:
            sys.stderr.write(r.stderr)
        return obj

    with concurrent.futures.ThreadPoolExecutor(max_workers=len(UNITS)) as ex:
        objs = list(ex.map(compile_one, UNITS))
    r = subprocess.run([nvcc, "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a"],
                       capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    with open(stamp, "w") as f:
        f.write(dig)
    return LIB


def build_variant(name: str, extra_flags, units=("jb_sim_kernel",)) -> str:
    """Tuning helper: ``_lib/variants/<name>/libjbsim.so`` with ``extra_flags`` (e.g. ``-DJB_PHASE_SYNC=0x52``) applied to
    the listed object stems (default: the shared-model step kernel); select it at run time with ``JB_LIB=<path>``."""
    build()
    vdir = os.path.join(LIBDIR, "variants", name)
    os.makedirs(vdir, exist_ok=True)
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    objs = []
    for unit, stem, extra in UNITS:
        if stem in units:
            obj = os.path.join(vdir, stem + ".o")
            r = subprocess.run([nvcc] + NVCC_FLAGS + extra + list(extra_flags) + ["-c", os.path.join(CSRC, unit), "-o", obj],
                               capture_output=True, text=True)
            with open(os.path.join(vdir, stem + ".ptxas.log"), "w") as f:
                f.write(r.stdout + r.stderr)
            if r.returncode != 0:
                raise RuntimeError(f"nvcc failed on {unit}:\n{r.stdout}\n{r.stderr}")
        else:
            obj = os.path.join(LIBDIR, stem + ".o")
        objs.append(obj)
    lib = os.path.join(vdir, "libjbsim.so")
    r = subprocess.run([nvcc, "-shared", "-o", lib] + objs + ["-gencode", "arch=compute_100a,code=sm_100a"],
                       capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return lib


if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "--variant":      # python -m jitterbug_b200.build --variant NAME -DFOO=1 ...
        print(build_variant(sys.argv[2], sys.argv[3:]))
    else:
        print(build(force="--force" in sys.argv, verbose=True))
