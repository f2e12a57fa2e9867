"""Builds libdana_b200.so (sm_100a) in-tree with nvcc.  No torch headers, no JIT cache:
the .so sits next to this file so it travels to the GPU box with the repo snapshot."""
import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libdana_b200.so")
STAMP = os.path.join(HERE, ".libdana_b200.stamp")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "--expt-relaxed-constexpr",
    "-Xcompiler", "-fPIC",
]
LINK_FLAGS = ["-shared", "-cudart", "static", "-gencode", "arch=compute_100a,code=sm_100a"]
# translation units: lib.cu (the forward: tcgen05 GEMM templates, ~2 min) and bwd.cu (backward layout kernels, seconds)
UNITS = ("lib.cu", "bwd.cu")
OBJ_DIR = os.path.join(HERE, ".obj")


def _nvcc():
    cand = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(cand):
        raise RuntimeError("nvcc not found; cannot build libdana_b200.so")
    return cand


def _source_hash():
    h = hashlib.sha256()
    files = sorted(os.listdir(CSRC)) + ["../../include/dana_b200.h"]
    for name in files:
        path = os.path.join(CSRC, name)
        if os.path.isfile(path):
            h.update(name.encode())
            with open(path, "rb") as f:
                h.update(f.read())
    h.update(" ".join(NVCC_FLAGS + LINK_FLAGS).encode())
    return h.hexdigest()


def is_fresh():
    if not (os.path.exists(LIB) and os.path.exists(STAMP)):
        return False
    with open(STAMP) as f:
        return f.read().strip() == _source_hash()


def _unit_hash(unit):
    """Hash of everything a unit can include: lib.cu sees every header, bwd.cu only api_common.cuh + the C header."""
    h = hashlib.sha256()
    names = sorted(os.listdir(CSRC)) if unit == "lib.cu" else [unit, "api_common.cuh"]
    for name in [n for n in names if not (unit == "lib.cu" and n == "bwd.cu")] + ["../../include/dana_b200.h"]:
        path = os.path.join(CSRC, name)
        if os.path.isfile(path):
            h.update(name.encode())
            with open(path, "rb") as f:
                h.update(f.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build(force=False, verbose=False):
    """Compile csrc/{lib,bwd}.cu -> libdana_b200.so (objects cached per unit under .obj/).  Returns the library path."""
    if not force and is_fresh():
        return LIB
    os.makedirs(OBJ_DIR, exist_ok=True)
    procs = []
    for unit in UNITS:
        obj = os.path.join(OBJ_DIR, unit + ".o")
        stamp = obj + ".stamp"
        want = _unit_hash(unit)
        if not force and os.path.exists(obj) and os.path.exists(stamp) and open(stamp).read().strip() == want:
            continue
        cmd = [_nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", "-o", obj, os.path.join(CSRC, unit)]
        procs.append((unit, stamp, want, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)))
    failed = False
    for unit, stamp, want, proc in procs:
        out, err = proc.communicate()
        if proc.returncode != 0:
            sys.stderr.write(out + err)
            failed = True
            continue
        if verbose:
            sys.stderr.write(err)
        with open(stamp, "w") as f:
            f.write(want)
    if failed:
        raise RuntimeError("nvcc failed building libdana_b200.so")
    cmd = [_nvcc()] + LINK_FLAGS + ["-o", LIB] + [os.path.join(OBJ_DIR, u + ".o") for u in UNITS]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        sys.stderr.write(proc.stdout + proc.stderr)
        raise RuntimeError("nvcc failed linking libdana_b200.so")
    with open(STAMP, "w") as f:
        f.write(_source_hash())
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
