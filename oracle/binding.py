"""TEST INFRASTRUCTURE ONLY: ctypes handle on the CPU oracle (oracle/liboracle.so).

Import this from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs only.  It reuses the prefix-generic binding class of asuna_b200.capi with
the ``oracle_`` prefix; the dependency points from the oracle to the product, never back.
"""
import ctypes as C
import os
import subprocess

from asuna_b200.capi import Context, Library

_HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_LIB = os.path.join(_HERE, "liboracle.so")


def build(force=False):
    if force or not os.path.exists(ORACLE_LIB):
        subprocess.check_call(["make", "-C", _HERE] + (["-B"] if force else []), stdout=subprocess.DEVNULL)
    return ORACLE_LIB


_lib = None


def library():
    global _lib
    if _lib is None:
        _lib = Library(build(), "oracle_")
    return _lib


class OracleContext(Context):
    def __init__(self, threads=0):
        super().__init__(library())
        if threads:
            self._call("set_threads", C.c_int(threads))

    def traversal_counters(self):
        out = (C.c_uint64 * 3)()
        self._call("traversal_counters", out)
        return {"node_visits": out[0], "tri_tests": out[1], "shadow_rays_nonzero": out[2]}


# ---- oracle/_ref: the reference's own GLSL compiled as C++ (oracle/refbuild/build_ref.py) ----
REF_LIB = os.path.join(_HERE, "_ref", "libref.so")
_ref = None


def build_ref(force=False):
    """Builds oracle/_ref/libref.so when the reference tree is present (this container); the GPU box only
    ever sees the prebuilt file.  Returns the path, or None when it neither exists nor can be built."""
    ref_root = os.environ.get("ASUNA_REFERENCE", "/root/reference")
    if (force or not os.path.exists(REF_LIB)) and os.path.isdir(os.path.join(ref_root, "src", "shaders")):
        subprocess.check_call(["python3", os.path.join(_HERE, "refbuild", "build_ref.py"), "--reference", ref_root],
                              stdout=subprocess.DEVNULL)
    return REF_LIB if os.path.exists(REF_LIB) else None


def ref_library():
    global _ref
    if _ref is None:
        path = build_ref()
        if path is None:
            raise FileNotFoundError("oracle/_ref/libref.so is not built and /root/reference is absent")
        _ref = Library(path, "oracle_")
        assert C.CDLL(path).oracle_is_reference_glsl() == 1
    return _ref


class RefContext(Context):
    """Same ABI as OracleContext; per-pixel arithmetic is the reference's GLSL text, not the restatement."""

    def __init__(self, threads=0):
        super().__init__(ref_library())
        if threads:
            self._call("set_threads", C.c_int(threads))
