"""Builds libfacevae_b200.so in-tree with nvcc for sm_100a (no torch headers: the library is a plain C ABI)."""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libfacevae_b200.so")
SOURCES = ["fv_host.cu", "fv_glue.cu", "fv_conv.cu", "fv_conv_ring.cu", "fv_conv_win.cu", "fv_wgrad.cu", "fv_wgrad_ring.cu", "fv_xrank.cu", "fv_pointwise.cu", "fv_outconv.cu", "fv_debug.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden", "-Xptxas", "-v"]
# --use_fast_math (approximate division / exp, flush-to-zero) only where nothing is contracted to 1e-4: the tensor-core
# translation units, whose epilogues round to bf16 anyway.  The fp32 glue (KL, losses, batch-norm finalize, Adam), the
# cross-rank exchange and the first-layer kernels are compiled with IEEE semantics.
FAST_MATH_SOURCES = {"fv_conv.cu", "fv_conv_ring.cu", "fv_conv_win.cu", "fv_wgrad.cu", "fv_wgrad_ring.cu", "fv_debug.cu"}
if os.environ.get("FV_TRACE"):                       # role-loop cycle counters in the ring kernels (debug only)
    NVCC_FLAGS = NVCC_FLAGS + ["-DFV_TRACE", "-rdc=false"]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: libfacevae_b200.so cannot be built")


def _fingerprint() -> str:
    h = hashlib.sha256()
    files = sorted(os.listdir(CSRC)) + [os.path.join("..", "..", "include", "facevae_b200.h")]
    for f in files:
        with open(os.path.join(CSRC, f), "rb") as fh:
            h.update(f.encode() + b"\0" + fh.read())
    h.update(" ".join(NVCC_FLAGS + sorted(FAST_MATH_SOURCES)).encode())
    return h.hexdigest()


def have_nvcc() -> bool:
    try:
        _nvcc()
        return True
    except RuntimeError:
        return False


def stale() -> bool:
    """True when the library's stamp does not match the current sources (or there is no stamp to compare with)."""
    stamp = LIB + ".stamp"
    return not (os.path.exists(LIB) and os.path.exists(stamp) and open(stamp).read().strip() == _fingerprint())


def build(force: bool = False, verbose: bool = False) -> str:
    stamp = LIB + ".stamp"
    fp = _fingerprint()
    if not force and os.path.exists(LIB) and os.path.exists(stamp) and open(stamp).read().strip() == fp:
        return LIB
    nvcc = _nvcc()
    objdir = os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    procs = []
    for src in SOURCES:
        obj = os.path.join(objdir, src.replace(".cu", ".o"))
        flags = NVCC_FLAGS + (["--use_fast_math"] if src in FAST_MATH_SOURCES else [])
        cmd = [nvcc, *flags, "-c", os.path.join(CSRC, src), "-o", obj]
        procs.append((src, obj, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    objs = []
    for src, obj, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode:
            sys.stderr.write(out)
        if p.returncode:
            raise RuntimeError(f"nvcc failed on {src}")
        with open(os.path.join(objdir, src + ".log"), "w") as fh:
            fh.write(out)
        objs.append(obj)
    cmd = [nvcc, "-shared", "-o", LIB, *objs, "-cudart", "static"]
    subprocess.run(cmd, check=True)
    with open(stamp, "w") as fh:
        fh.write(fp)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
