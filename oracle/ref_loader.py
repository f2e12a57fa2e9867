"""Test infrastructure: imports the UNMODIFIED reference python (/root/reference/lib) in the build
container so that golden vectors can be generated from it (oracle/make_golden.py) and the oracle
can be cross-checked live (tests that are skipped when /root/reference is absent, e.g. on the GPU box).

Shims: a 10-line `easydict` stand-in, and the reference's CPU operator
extension compiled in place by oracle/build_ref.py registered as `model._C`."""
import os
import sys
import types

REF_ROOT = "/root/reference"
REF_LIB = os.path.join(REF_ROOT, "lib")


def available():
    return os.path.isdir(REF_LIB)


class _EasyDict(dict):
    """Minimal attribute-dict with recursive wrapping (what the reference uses easydict for)."""

    def __init__(self, d=None, **kw):
        super().__init__()
        for k, v in dict(d or {}, **kw).items():
            self[k] = v

    def __setitem__(self, k, v):
        if isinstance(v, dict) and not isinstance(v, _EasyDict):
            v = _EasyDict(v)
        super().__setitem__(k, v)

    __setattr__ = __setitem__

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)


def load():
    """Returns the reference's `model` package with `model._C` bound to the compiled reference ops."""
    if not available():
        raise RuntimeError("/root/reference is not mounted here")
    if "easydict" not in sys.modules:
        m = types.ModuleType("easydict")
        m.EasyDict = _EasyDict
        sys.modules["easydict"] = m
    if REF_LIB not in sys.path:
        sys.path.insert(0, REF_LIB)
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import build_ref
    build_ref.build()
    ref_c = build_ref.load()
    import model  # the reference package
    model._C = ref_c
    sys.modules["model._C"] = ref_c
    return model
