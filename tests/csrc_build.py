"""Test helper: make sure the in-tree shared objects exist (built from source if nvcc is here)."""
import importlib.util
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def ensure_built():
    path = os.path.join(ROOT, "all-in-one-deflicker_b200", "csrc", "build.py")
    spec = importlib.util.spec_from_file_location("b200_build", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    have = os.path.exists(mod.LIB) and os.path.exists(mod.HOSTLIB)
    if os.path.exists("/usr/local/cuda/bin/nvcc"):
        try:
            mod.build()
        except Exception:
            if not have:
                raise
    elif not have:
        raise RuntimeError("libb200deflicker.so missing and nvcc not available")
    return mod.LIB
