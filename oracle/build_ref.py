"""Test infrastructure (not product code): compiles the reference's own CPU operator extension
(`model._C`: nms, roi_align_forward, ...) from the sources WHERE THEY LIE under /root/reference into
oracle/_ref/ref_C*.so.  Nothing is copied; `ref_compat.h` is force-included to bridge two
torch-1.x-only tokens.  Runs only where /root/reference exists (this container); the GPU box
uses the prebuilt .so that travels with the snapshot."""
import glob
import os
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
REF_CSRC = "/root/reference/lib/model/csrc"
OUT_DIR = os.path.join(HERE, "_ref")


def ref_so_path():
    hits = sorted(glob.glob(os.path.join(OUT_DIR, "ref_C*.so")))
    return hits[0] if hits else None


def build(force=False):
    if not os.path.isdir(REF_CSRC):
        return ref_so_path()
    if ref_so_path() and not force:
        return ref_so_path()
    import torch
    from torch.utils import cpp_extension
    os.makedirs(OUT_DIR, exist_ok=True)
    srcs = [os.path.join(REF_CSRC, "vision.cpp")] + sorted(glob.glob(os.path.join(REF_CSRC, "cpu", "*.cpp")))
    inc = cpp_extension.include_paths()
    out = os.path.join(OUT_DIR, "ref_C" + (sysconfig.get_config_var("EXT_SUFFIX") or ".so"))
    cmd = ["g++", "-O3", "-std=c++17", "-fPIC", "-shared", "-w", "-DTORCH_EXTENSION_NAME=ref_C",
           "-DTORCH_API_INCLUDE_EXTENSION_H", "-D_GLIBCXX_USE_CXX11_ABI=%d" % int(torch._C._GLIBCXX_USE_CXX11_ABI),
           "-include", os.path.join(HERE, "ref_compat.h"), "-I", REF_CSRC,
           "-I", sysconfig.get_paths()["include"]]
    for i in inc:
        cmd += ["-isystem", i]
    libdir = os.path.join(os.path.dirname(torch.__file__), "lib")
    cmd += srcs + ["-o", out, "-L", libdir, "-ltorch", "-ltorch_cpu", "-lc10", "-ltorch_python",
                   "-Wl,-rpath," + libdir]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        sys.stderr.write(proc.stderr[-4000:])
        raise RuntimeError("building the reference CPU extension failed")
    return out


def load():
    """Import the compiled reference extension as a module (None when unavailable)."""
    path = ref_so_path()
    if path is None:
        return None
    import importlib.util
    import torch  # noqa: F401  (libtorch must be loaded first)
    spec = importlib.util.spec_from_file_location("ref_C", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
