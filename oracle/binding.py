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
