"""TEST INFRASTRUCTURE ONLY -- compiles the one piece of native code the reference ships, its PHOC
extension (pythia/utils/phoc/src/cphoc.c, a single CPython C-API source file), from where it lies under
/root/reference into oracle/_ref/cphoc.so.  No reference source is copied into this repository; oracle/_ref/
is git-ignored (it still travels to the GPU box with gpurun, where tests use it when present and fall back to
the committed golden vectors otherwise).

    python oracle/build_ref.py          # or __graft_entry__.build()
"""
import importlib.util
import os
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = "/root/reference/pythia/utils/phoc/src/cphoc.c"
REF_WRAPPER = "/root/reference/pythia/utils/phoc/build_phoc.py"
OUT_DIR = os.path.join(HERE, "_ref")
OUT = os.path.join(OUT_DIR, "cphoc.so")


def build(force=False):
    """-> path of oracle/_ref/cphoc.so, or None when /root/reference is absent and nothing was prebuilt."""
    if os.path.exists(OUT) and not force and (
            not os.path.exists(REF_SRC) or os.path.getmtime(OUT) >= os.path.getmtime(REF_SRC)):
        return OUT
    if not os.path.exists(REF_SRC):
        return OUT if os.path.exists(OUT) else None
    os.makedirs(OUT_DIR, exist_ok=True)
    inc = sysconfig.get_paths()["include"]
    # -include string.h: the source calls strlen/memcmp without declaring them (an error for gcc >= 14,
    # a warning for 13); -O2 without -ffast-math keeps binary32 semantics (no contraction is possible: the
    # source has no multiply-add)
    cmd = ["gcc", "-O2", "-fPIC", "-shared", "-w", "-include", "string.h", "-I", inc, REF_SRC, "-o", OUT]
    subprocess.run(cmd, check=True)
    return OUT


def load():
    """-> the compiled reference module (has build_phoc(str) -> list of 604 floats), or None."""
    path = build()
    if path is None:
        return None
    spec = importlib.util.spec_from_file_location("cphoc", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def load_reference_build_phoc():
    """The reference's own python wrapper (build_phoc.py) bound to the compiled module; only available
    where /root/reference exists.  Used by tests/golden/make_phoc_golden.py."""
    import types
    mod = load()
    if mod is None or not os.path.exists(REF_WRAPPER):
        return None
    pkg = types.ModuleType("_refphoc")
    pkg.__path__ = []
    sys.modules["_refphoc"] = pkg
    sys.modules["_refphoc.cphoc"] = mod
    spec = importlib.util.spec_from_file_location("_refphoc.build_phoc", REF_WRAPPER)
    wrapper = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(wrapper)
    return wrapper.build_phoc


if __name__ == "__main__":
    print(build(force=True))
